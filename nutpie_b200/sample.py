"""`nutpie.sample` for the B200 engine.

Mirrors python/nutpie/sample.py (reference): `sample()` (:823-1102) with the
same keyword surface, `_BackgroundSampler` (:481-725) with wait / pause /
resume / abort / cancel / inspect, and the trace assembly of `_arrow_to_arviz`
(:62-164) — except that the trace arrives as two dense arrays for ALL chains
([chain, draw, dim] draws and [chain, draw, stat] statistics, one D2H copy each)
instead of one Arrow RecordBatch pair per chain, so posterior / sample_stats
groups are built by slicing, not by per-chain padding loops.  `arviz` is not
installed in the build image; when it is importable the result is converted
with `arviz.from_dict`, otherwise a light `Trace` container with the same
groups is returned.
"""
from __future__ import annotations

import json
import os
import warnings
from dataclasses import dataclass, field
from typing import Any, Literal

import numpy as np

from . import _lib


@dataclass
class Trace:
    """Plain stand-in for arviz.InferenceData: groups of {name: array[chain, draw, ...]}."""

    posterior: dict = field(default_factory=dict)
    sample_stats: dict = field(default_factory=dict)
    warmup_posterior: dict = field(default_factory=dict)
    warmup_sample_stats: dict = field(default_factory=dict)
    unconstrained_posterior: dict = field(default_factory=dict)
    dims: dict = field(default_factory=dict)
    coords: dict = field(default_factory=dict)
    attrs: dict = field(default_factory=dict)

    def groups(self):
        return [g for g in ("posterior", "sample_stats", "warmup_posterior", "warmup_sample_stats",
                            "unconstrained_posterior") if getattr(self, g)]


_STAT_EXPORT = ["depth", "maxdepth_reached", "index_in_trajectory", "logp", "energy",
                "energy_error", "diverging", "step_size", "step_size_bar", "n_steps",
                "mean_tree_accept", "mean_tree_accept_sym"]


def _trace_to_groups(trace: _lib.PyTrace, compiled_model, settings, save_warmup, var_names=None,
                     store_unconstrained=False):
    """The job of _arrow_to_arviz (sample.py:62-164) on dense arrays."""
    num_tune = settings.num_tune if save_warmup else 0
    rows = trace.rows_filled.astype(np.int64)
    n_rows = int(rows.min()) if len(rows) else 0  # equal for finished runs
    draws = trace.draws[:, :n_rows]
    n_tune_rows = min(num_tune, n_rows)
    full = trace.expanded or draws.shape[-1] == compiled_model.n_dim
    if trace.expanded:  # expand_vector ran on the device
        values = compiled_model._split_expanded(draws)
    elif full:
        values = compiled_model._expand(draws)
    else:
        values = {"unconstrained_draw": draws}
    if var_names is not None:
        values = {k: v for k, v in values.items() if k in var_names}
    dims = dict(compiled_model.dims or {})
    if hasattr(compiled_model, "_model_for_expand"):  # host plug-in: dims of the variable table
        dims.update(compiled_model._model_for_expand()._variable_dims())
    out = Trace(dims={k: list(v) for k, v in dims.items()}, coords=dict(compiled_model.coords or {}))
    for name, arr in values.items():
        out.warmup_posterior[name] = arr[:, :n_tune_rows]
        out.posterior[name] = arr[:, n_tune_rows:]
    for name in _STAT_EXPORT:
        a = trace.stat(name)[:, :n_rows]
        out.warmup_sample_stats[name] = a[:, :n_tune_rows]
        out.sample_stats[name] = a[:, n_tune_rows:]
    vec_stats = [("gradient", trace.gradients)] + trace.mass_matrix_columns()
    if trace.divergences is not None:  # sample.py:641-646
        vec_stats += [(n, trace.divergences[:, :, k]) for k, n in enumerate(_lib.DIVERGENCE_COLUMNS)]
    for name, arr in vec_stats:
        if arr is not None:
            out.warmup_sample_stats[name] = arr[:, :n_tune_rows]
            out.sample_stats[name] = arr[:, n_tune_rows:n_rows]
            out.dims[name] = ["mass_matrix_eigvals_dim" if name == "mass_matrix_eigvals"
                              else "unconstrained_parameter"]
    if store_unconstrained and full and not trace.expanded:
        out.sample_stats["unconstrained_draw"] = draws[:, n_tune_rows:]
        out.warmup_sample_stats["unconstrained_draw"] = draws[:, :n_tune_rows]
        out.dims["unconstrained_draw"] = ["unconstrained_parameter"]
    if not save_warmup:
        out.warmup_posterior, out.warmup_sample_stats = {}, {}
    out.attrs = {"inference_library": "nutpie_b200", "_version": _lib.__version__,
                 "_settings": json.dumps(settings.as_dict())}  # sample.py:666-672
    return out


def _add_arrow_data(data_dict, max_length, batch, chain, n_chains, dims, skip_vars):
    """python/nutpie/sample.py:167-214 on the pyarrow API: write one chain's RecordBatch into
    [chain, draw, *shape] arrays padded to `max_length` draws (NaN / None / 0 by dtype)."""
    import pyarrow

    num_draws = batch.num_rows
    for name in batch.schema.names:
        if name in skip_vars:
            continue
        meta = batch.schema.field(name).metadata or {}
        item_dims = meta.get(b"dims", b"").decode("utf-8")
        item_dims = item_dims.split(",") if item_dims else []
        item_shape = meta.get(b"shape", b"").decode("utf-8")
        item_shape = [int(x) for x in item_shape.split(",")] if item_shape else []
        total_shape = [n_chains, max_length, *item_shape]
        col = batch.column(name)
        if isinstance(col, pyarrow.ChunkedArray):
            col = col.combine_chunks()
        is_null = col.is_null()
        while hasattr(col, "flatten") and pyarrow.types.is_nested(col.type):
            col = col.flatten()
        dtype = col.type.to_pandas_dtype()
        if name not in data_dict:
            if dtype in (np.float64, np.float32):
                data = np.full(total_shape, np.nan, dtype=dtype)
            elif dtype == np.dtype("O"):
                data = np.full(total_shape, None, dtype=dtype)
            else:
                data = np.zeros(total_shape, dtype=dtype)
            data_dict[name] = data
            dims[name] = item_dims
        values = col.to_numpy(zero_copy_only=False)
        if is_null.sum().as_py() == 0:
            data_dict[name][chain, :num_draws] = values.reshape((num_draws,) + tuple(item_shape))
        else:
            mask = ~is_null.to_numpy(zero_copy_only=False)
            if values.shape[0] == num_draws:
                values = values[mask]
            data_dict[name][chain, :num_draws][mask] = values.reshape((mask.sum(),) + tuple(item_shape))


def _arrow_to_groups(draw_batches, stat_batches, skip_vars=None, reparameterized_names=None,
                     keep_unconstrained_draw=False, coords=None, save_warmup=True, attrs=None):
    """`_arrow_to_arviz` (python/nutpie/sample.py:62-164) for per-chain Arrow RecordBatch pairs —
    the form `PyTrace.get_arrow_trace()` and the Rust shim hand over (src/wrapper.rs:1477-1494):
    split every chain at its own number of tuning rows (the `tuning` column), pad ragged chains,
    move reparameterized value variables to `unconstrained_posterior`.  Returns the `Trace`
    groups (converted by `_maybe_arviz` when arviz is importable)."""
    skip_vars = list(skip_vars or [])
    reparameterized_names = list(reparameterized_names or [])
    n_chains = len(draw_batches)
    assert n_chains == len(stat_batches)
    max_tuning = max_posterior = 0
    num_tuning = []
    for draw, stat in zip(draw_batches, stat_batches):
        n_tune = int(np.asarray(stat.column("tuning").to_numpy(zero_copy_only=False)).sum())
        assert draw.num_rows == stat.num_rows
        max_tuning = max(max_tuning, n_tune)
        max_posterior = max(max_posterior, draw.num_rows - n_tune)
        num_tuning.append(n_tune)
    data_tune, data_post, stats_tune, stats_post, dims = {}, {}, {}, {}, {}
    for i, draw in enumerate(draw_batches):
        _add_arrow_data(data_tune, max_tuning, draw.slice(0, num_tuning[i]), i, n_chains, dims, [])
        _add_arrow_data(data_post, max_posterior, draw.slice(num_tuning[i], draw.num_rows - num_tuning[i]),
                        i, n_chains, dims, [])
    for i, stat in enumerate(stat_batches):
        _add_arrow_data(stats_tune, max_tuning, stat.slice(0, num_tuning[i]), i, n_chains, dims, skip_vars)
        _add_arrow_data(stats_post, max_posterior, stat.slice(num_tuning[i], stat.num_rows - num_tuning[i]),
                        i, n_chains, dims, skip_vars)
    uc_post = {n: data_post.pop(n) for n in reparameterized_names if n in data_post}
    uc_tune = {n: data_tune.pop(n) for n in reparameterized_names if n in data_tune}
    out = Trace(posterior=data_post, sample_stats=stats_post, dims=dims, coords=dict(coords or {}),
                attrs=dict(attrs or {}))
    if save_warmup:
        out.warmup_posterior, out.warmup_sample_stats = data_tune, stats_tune
    if keep_unconstrained_draw and uc_post:
        out.unconstrained_posterior = uc_post
    return out


def _maybe_arviz(tr: Trace):
    try:
        import arviz  # noqa: F401
    except Exception:
        return tr
    import arviz

    kwargs = dict(dims=tr.dims, coords=tr.coords)
    groups = {"posterior": tr.posterior, "sample_stats": tr.sample_stats}
    if tr.warmup_posterior:
        groups["warmup_posterior"] = tr.warmup_posterior
        groups["warmup_sample_stats"] = tr.warmup_sample_stats
    if tr.unconstrained_posterior:  # python/nutpie/sample.py:147-162
        groups["unconstrained_posterior"] = tr.unconstrained_posterior
    try:
        return arviz.from_dict(groups, **kwargs)  # arviz >= 1.0
    except TypeError:
        return arviz.from_dict(**groups, **kwargs)


class _BackgroundSampler:
    """sample.py:481-725 — owns the running PySampler and turns its trace into a result."""

    def __init__(self, compiled_model, settings, init_mean, cores, *, progress_bar=True,
                 progress_callback=None, save_warmup=True, return_raw_trace=False,
                 progress_rate=100, var_names=None, store_unconstrained=False, device=0,
                 chain_id_offset=0, trace_buffers=None, **sampler_kw):
        self._settings = settings
        self._compiled_model = compiled_model
        self._save_warmup = save_warmup
        self._return_raw_trace = return_raw_trace
        self._var_names = var_names
        self._store_unconstrained = store_unconstrained
        self._trace_buffers = trace_buffers
        settings._c.save_warmup = 1 if save_warmup else 0
        if progress_callback is not None:
            progress_type = _lib.ProgressType("callback", progress_rate, progress_callback)
        else:
            progress_type = _lib.ProgressType.none()
        # the reference's call (python/nutpie/sample.py:586-594) plus this engine's keywords
        self._sampler = compiled_model._make_sampler(
            settings, init_mean, cores, progress_type, progress_callback, progress_rate,
            _lib.PyStorage.arrow(), device=device, chain_id_offset=chain_id_offset,
            trace_buffers=trace_buffers, **sampler_kw)
        self._html = None

    def wait(self, *, timeout=None):
        """Wait until sampling is finished and return the trace (sample.py:596-608).
        Raises TimeoutError if `timeout` seconds pass first."""
        self._sampler.wait(timeout)
        results = self._sampler.take_results(self._trace_buffers)
        return self._extract(results)

    def inspect(self):
        """Snapshot of the draws so far, sampling continues (sample.py:688-691)."""
        results = self._sampler.inspect()
        return self._extract(results)

    def _extract(self, results: _lib.PyTrace):
        if self._return_raw_trace:
            return results
        tr = _trace_to_groups(results, self._compiled_model, self._settings, self._save_warmup,
                              self._var_names, self._store_unconstrained)
        return _maybe_arviz(tr)

    def pause(self):
        self._sampler.pause()

    def resume(self):
        self._sampler.resume()

    @property
    def is_finished(self):
        return self._sampler.is_finished()

    def abort(self):
        """Stop sampling and return the partial trace (sample.py:699-707)."""
        self._sampler.abort()
        return self._extract(self._sampler.inspect())

    def cancel(self):
        """Stop sampling and discard progress (sample.py:709-715)."""
        self._sampler.abort()
        self._sampler.close()

    def kernel_ms(self):
        return self._sampler.kernel_ms()

    def close(self):
        self._sampler.close()

    def __del__(self):  # sample.py:717-721
        try:
            self._sampler.close()
        except Exception:
            pass


def sample(
    compiled_model,
    *,
    draws: int | None = None,
    tune: int | None = None,
    chains: int | None = None,
    cores: int | None = None,
    seed: int | None = None,
    save_warmup: bool = True,
    progress_bar: bool = True,
    sampler: Literal["nuts", "mclmc"] = "nuts",
    adaptation: Literal["diag", "draw_diag", "low_rank", "flow"] = "diag",
    init_mean: np.ndarray | None = None,
    return_raw_trace: bool = False,
    blocking: bool = True,
    progress_callback: Any | None = None,
    progress_template: str | None = None,
    progress_style: str | None = None,
    progress_rate: int = 100,
    zarr_store=None,
    store_unconstrained: bool = False,
    var_names=None,
    device: int = 0,
    devices=None,
    chain_id_offset: int = 0,
    trace_buffers=None,
    expand_on_device: bool | None = None,
    **kwargs,
):
    """Sample the posterior of a compiled device model on a B200.

    Same keyword surface as `nutpie.sample` (python/nutpie/sample.py:823-1102);
    `cores` is accepted and ignored (chains map to warps/CTAs, not host threads).
    Extra keywords: `device` (CUDA device index), `devices` (an int n = the first n GPUs, a
    list of device indices, or "all": ONE call samples over several GPUs of this process —
    chains are split into contiguous blocks, one persistent kernel and one host thread per
    device, the role `cores` plays in the reference), `chain_id_offset` (global id
    of this process's first chain when a run is sharded over GPUs),
    `trace_buffers` (dict(draws=, stats=) of preallocated — ideally pinned —
    host arrays to receive the trace)."""
    if zarr_store is not None:
        raise NotImplementedError("zarr_store is not supported by the B200 engine")
    _use_grad_based = None
    for _old, _new in (("low_rank_modified_mass_matrix", "low_rank"), ("transform_adapt", "flow")):
        if _old in kwargs:
            if kwargs.pop(_old):
                warnings.warn(f"`{_old}` is deprecated. Use `adaptation='{_new}'` instead.",
                              FutureWarning, stacklevel=2)
                if adaptation == "diag":
                    adaptation = _new
                else:
                    raise ValueError(f"`{_old}` is deprecated and cannot be combined with the "
                                     "`adaptation` argument.")
    if "use_grad_based_mass_matrix" in kwargs:
        _use_grad_based = kwargs.pop("use_grad_based_mass_matrix")
        warnings.warn("`use_grad_based_mass_matrix` is deprecated. Use `adaptation='draw_diag'` "
                      "instead of `use_grad_based_mass_matrix=False`.", FutureWarning, stacklevel=2)

    if sampler == "nuts":
        if adaptation == "low_rank":
            if not _lib.LOW_RANK_SUPPORTED:
                raise NotImplementedError(
                    "adaptation='low_rank' is not implemented by the B200 engine yet")
            settings = _lib.PyNutsSettings.LowRank(seed)
        elif adaptation == "flow":
            settings = _lib.PyNutsSettings.Flow(seed)
        elif adaptation in ("diag", "draw_diag"):
            settings = _lib.PyNutsSettings.Diag(seed)
            if adaptation == "draw_diag" or _use_grad_based is False:
                settings.use_grad_based_mass_matrix = False
        else:
            raise ValueError(f"Unknown adaptation strategy '{adaptation}'. "
                             "Expected one of: 'diag', 'draw_diag', 'low_rank', 'flow'.")
    elif sampler == "mclmc":
        settings = _lib.PyMclmcSettings.Diag(seed)
    else:
        raise ValueError(f"Unknown sampler '{sampler}'. Expected one of: 'nuts', 'mclmc'.")

    sampler_kw = {}
    if devices is not None:
        sampler_kw["devices"] = list(range(_lib.device_count())) if devices == "all" else devices
    for k in ("q0", "z_tape", "draws_per_launch"):
        if k in kwargs:
            sampler_kw[k] = kwargs.pop(k)
    updates = dict(kwargs)
    if tune is not None:
        updates["num_tune"] = tune
    if draws is not None:
        updates["num_draws"] = draws
    if chains is not None:
        updates["num_chains"] = chains
    settings.update(updates)
    if store_unconstrained:
        settings.store_unconstrained = True
    # the trace holds expanded draws (constrained values + deterministics) computed on the
    # device, like the reference's posterior; unconstrained draws on request
    settings._c.expand_draws = 0 if (store_unconstrained or expand_on_device is False) else 1
    if init_mean is None:
        init_mean = np.zeros(compiled_model.n_dim)

    bg = _BackgroundSampler(
        compiled_model, settings, init_mean, cores, progress_bar=progress_bar,
        progress_callback=progress_callback, save_warmup=save_warmup,
        return_raw_trace=return_raw_trace, progress_rate=progress_rate, var_names=var_names,
        store_unconstrained=store_unconstrained, device=device, chain_id_offset=chain_id_offset,
        trace_buffers=trace_buffers, **sampler_kw)
    if not blocking:
        return bg
    try:
        result = bg.wait()
    except KeyboardInterrupt:
        result = bg.abort()
    except BaseException:
        bg.cancel()
        raise
    bg.close()
    return result
