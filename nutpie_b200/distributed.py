"""Multi-GPU: chains shard across ranks, one process per GPU.

Chains are independent (each reference chain has its own RNG stream, adaptation
state and trace: src/wrapper.rs:1482-1492 returns one batch pair per chain), so
a rank samples a contiguous block of GLOBAL chain ids with no data-path
collective; the random streams are keyed by global chain id, which makes the
sharded run identical, chain for chain, to the single-GPU run.  The only
exchange is at trace collection: an all-gather of the per-rank sample-stats
(and, on request, draws) buffers over torch.distributed — NCCL on GPUs (NVLink),
gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard(n_chains_total: int, rank: int, world: int):
    """Contiguous block of global chain ids for `rank`: (n_local, offset).  The
    first (n_chains_total % world) ranks get one extra chain."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_chains_total), int(world))
    n_local = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return n_local, offset


def all_gather_chains(local, n_chains_total: int, group=None, chain_axis: int = 0):
    """All-gather per-chain arrays from every rank into global chain order along
    `chain_axis` (0 for chain-major [n_local, ...] arrays; 1 for the engine's row-major
    [row, n_local, ...] trace buffers).  Accepts a numpy array (gathered over the group's
    default device: CPU for gloo) or a torch tensor (gathered where it lives)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    is_np = isinstance(local, np.ndarray)
    if chain_axis != 0:
        moved = np.moveaxis(local, chain_axis, 0) if is_np else local.movedim(chain_axis, 0)
        out = all_gather_chains(moved, n_chains_total, group, 0)
        return np.moveaxis(out, 0, chain_axis) if is_np else out.movedim(0, chain_axis)
    t = torch.from_numpy(np.ascontiguousarray(local)) if is_np else local.contiguous()
    if dist.get_backend(group) == "nccl" and not t.is_cuda:
        t = t.cuda()
    sizes = [shard(n_chains_total, r, world)[0] for r in range(world)]
    assert t.shape[0] == sizes[rank], "local block does not match shard()"
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=group)
    else:  # ragged shards: pad to the largest block
        mx = max(sizes)
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out = torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)
    return out.cpu().numpy() if is_np else out


def sample_sharded(compiled_model, *, chains: int, gather_draws: bool = False, sampler_fn=None,
                   group=None, **kwargs):
    """nutpie_b200.sample for `chains` chains in total, this rank sampling its shard on
    its own GPU (LOCAL_RANK).  Returns (local PyTrace, gathered stats[, gathered draws]).
    `sampler_fn(n_local, offset, **kwargs) -> (draws, stats)` replaces the GPU sampler in
    the CPU tests."""
    import os

    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_local, offset = shard(chains, rank, world)
    if sampler_fn is None:
        from .sample import sample

        tr = sample(compiled_model, chains=n_local, chain_id_offset=offset,
                    device=int(os.environ.get("LOCAL_RANK", "0")), return_raw_trace=True, **kwargs)
        draws, stats = tr.draws, tr.stats
    else:
        tr = None
        draws, stats = sampler_fn(n_local, offset, **kwargs)
    out = [tr, all_gather_chains(stats, chains, group)]
    if gather_draws:
        out.append(all_gather_chains(draws, chains, group))
    return tuple(out)
