"""Synthetic stand-ins for the datasets the reference examples download.

`radon.csv` (README.md:45-51, via pm.get_data) is not reachable offline, so the
benchmark uses a seeded synthetic data set of the same shape: J = 85 counties,
N = 919 observations, skewed county sizes (a few large counties, many with
<= 5 houses), ~17 % basement-less (`floor = 1`) houses, log-radon generated from
the hierarchical model itself.
"""
from __future__ import annotations

import numpy as np


def make_radon_data(n_county: int = 85, n_obs: int = 919, seed: int = 20240925):
    """Return dict(y float64[N], county int32[N], floor uint8[N], n_county, truth)."""
    rng = np.random.default_rng(seed)
    # skewed county sizes: every county has at least one house, the rest is
    # allocated with Dirichlet(0.6) weights (max ~ 100+, many <= 5)
    w = rng.dirichlet(np.full(n_county, 0.6))
    sizes = 1 + rng.multinomial(n_obs - n_county, w)
    county = np.repeat(np.arange(n_county, dtype=np.int32), sizes)
    # file order in the real csv is not sorted by county: shuffle
    perm = rng.permutation(n_obs)
    county = np.ascontiguousarray(county[perm])
    floor = (rng.random(n_obs) < 0.17).astype(np.uint8)
    truth = dict(intercept=1.5, floor_effect=-0.6, county_sd=0.3, county_floor_sd=0.2, sigma=0.75)
    a = rng.normal(0.0, truth["county_sd"], n_county)
    b = rng.normal(0.0, truth["county_floor_sd"], n_county)
    mu = truth["intercept"] + a[county] + floor * (truth["floor_effect"] + b[county])
    y = rng.normal(mu, truth["sigma"])
    return dict(y=np.ascontiguousarray(y, dtype=np.float64), county=county, floor=floor,
                n_county=n_county, truth=truth)
