"""Posterior diagnostics used by the benchmark (ESS/s is part of BASELINE.json's metric).

arviz is not installed here, so bulk effective sample size is implemented after
Vehtari et al. 2021 / Stan: split chains, FFT autocovariance per chain,
Geyer's initial monotone sequence on the chain-averaged autocorrelation.  Rank
normalisation is skipped (all targets here are close to Gaussian on the
unconstrained scale); the estimator is otherwise the one behind arviz.ess.
torch is used only as an FFT library (on the GPU when present).
"""
from __future__ import annotations

import numpy as np


def ess(draws: np.ndarray, max_chains: int | None = 256, device: str | None = None) -> np.ndarray:
    """draws: [chain, draw, param] -> ess[param] (total over the chains given).
    If max_chains is set, the estimate uses the first max_chains chains and is
    scaled to the full chain count (chains are exchangeable)."""
    import torch

    n_chains_all = draws.shape[0]
    x = draws[:max_chains] if max_chains else draws
    if device is None:
        device = "cuda" if torch.cuda.is_available() else "cpu"
    x = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float64, device=device)
    c, n, p = x.shape
    half = n // 2
    x = torch.cat([x[:, :half], x[:, half:2 * half]], dim=0)  # split chains
    m, n = x.shape[0], half
    x = x.permute(2, 0, 1).contiguous()  # [param, chain, draw]
    mean = x.mean(dim=2, keepdim=True)
    xc = x - mean
    nfft = 1 << int(np.ceil(np.log2(2 * n)))
    f = torch.fft.rfft(xc, n=nfft, dim=2)
    acov = torch.fft.irfft(f * f.conj(), n=nfft, dim=2)[..., :n] / n  # biased autocovariance
    chain_var = acov[..., 0] * n / (n - 1.0)
    W = chain_var.mean(dim=1)                       # within-chain variance   [param]
    B_over_n = mean.squeeze(2).var(dim=1, unbiased=True)
    var_plus = W * (n - 1.0) / n + B_over_n
    rho = 1.0 - (W[:, None] - acov.mean(dim=1)) / var_plus[:, None]  # [param, lag]
    rho[:, 0] = 1.0
    # Geyer: sums of adjacent pairs, truncated at the first negative pair, made monotone
    npairs = n // 2
    pair = rho[:, 0:2 * npairs:2] + rho[:, 1:2 * npairs:2]
    neg = pair < 0
    first_neg = torch.where(neg.any(dim=1), neg.float().argmax(dim=1),
                            torch.full((p,), npairs, device=pair.device))
    idx = torch.arange(npairs, device=pair.device)[None, :]
    pair = torch.where(idx < first_neg[:, None], pair, torch.zeros_like(pair))
    pair = torch.cummin(pair, dim=1).values
    tau = -1.0 + 2.0 * pair.sum(dim=1)
    tau = torch.clamp(tau, min=1.0 / np.log10(m * n))
    ess_sub = (m * n) / tau
    return (ess_sub * (n_chains_all / c)).cpu().numpy()
