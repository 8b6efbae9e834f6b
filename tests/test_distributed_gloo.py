"""CPU test of the multi-rank path (world_size 2, gloo): chain sharding + the stats gather.
The per-rank sampler is the CPU oracle standing in for the GPU engine; both key their
random streams by GLOBAL chain id, so the gathered trace must equal the unsharded run."""
import os
import socket

import numpy as np
import pytest


def test_shard_partitions_exactly():
    from nutpie_b200.distributed import shard

    for total in (1, 7, 8, 1024, 1000):
        for world in (1, 2, 3, 8):
            blocks = [shard(total, r, world) for r in range(world)]
            assert sum(n for n, _ in blocks) == total
            off = 0
            for n, o in blocks:
                assert o == off
                off += n
    with pytest.raises(ValueError):
        shard(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nutpie_b200.distributed import sample_sharded
        from oracle import pyoracle as O

        model = O.Model("funnel", 5)

        def oracle_sampler(n_local, offset, **kw):
            s = O.default_settings(seed=21, num_tune=40, num_draws=30)
            r = O.sample(model, s, n_local, chain_id_offset=offset, n_threads=1)
            return r["draws"], r["stats"]

        _, stats, draws = sample_sharded(None, chains=total, gather_draws=True,
                                         sampler_fn=oracle_sampler)
        # the engine's own buffers are row-major [row][chain][...]: gather along axis 1
        from nutpie_b200.distributed import all_gather_chains, shard

        n_local, offset = shard(total, rank, world)
        local_rm = np.ascontiguousarray(stats[offset:offset + n_local].transpose(1, 0, 2))
        stats_rm = all_gather_chains(local_rm, total, chain_axis=1)
        if rank == 0:
            q.put((stats, draws, stats_rm))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [6, 7])
def test_sharded_run_equals_single_run(total):
    import torch.multiprocessing as mp

    from oracle import pyoracle as O

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats, draws, stats_rm = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = O.default_settings(seed=21, num_tune=40, num_draws=30)
    ref = O.sample(O.Model("funnel", 5), s, total)
    assert np.array_equal(stats, ref["stats"])
    assert np.array_equal(draws, ref["draws"])
    assert np.array_equal(stats_rm.transpose(1, 0, 2), ref["stats"])
