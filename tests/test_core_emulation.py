"""CPU tests: the CUDA sampler core (nutpie_b200/csrc/nuts_core.cuh) compiled as host
C++ with a one-thread group must reproduce the recursive oracle BIT FOR BIT on
densities whose arithmetic is order-independent for one thread — this checks the
iterative tree / slot pool / shared-memory tier / resume logic without a GPU."""
import numpy as np
import pytest

from oracle import pyoracle as O
from tests.emul import pyemul as E

CASES = [
    ("normal", 1, dict(mu=0.0, sigma=1.0)),
    ("normal", 10, dict(mu=3.0, sigma=2.0)),
    ("normal", 33, dict(mu=-1.0, sigma=0.5)),
    ("funnel", 9, {}),
]


@pytest.mark.parametrize("kind,dim,kw", CASES)
@pytest.mark.parametrize("smem_slots,max_per_launch", [(0, 0), (3, 0), (6, 29), (64, 7)])
def test_emulated_core_equals_oracle(kind, dim, kw, smem_slots, max_per_launch):
    s = O.default_settings(seed=11, num_tune=150, num_draws=80, store_mass_matrix=1, store_gradient=1)
    a = O.sample(O.Model(kind, dim, **kw), s, 3)
    b = E.sample(kind, dim, s, 3, smem_slots=smem_slots, max_per_launch=max_per_launch, **kw)
    for k in ("draws", "stats", "gradients", "mass_matrix_inv"):
        assert np.array_equal(a[k], b[k]), k
    assert a["total_steps"] == b["total_steps"]


@pytest.mark.parametrize("opts", [
    dict(maxdepth=3), dict(maxdepth=12, mindepth=2), dict(check_turning=0, maxdepth=4),
    dict(use_grad_based_estimate=0), dict(step_size_method=2, fixed_step_size=0.3),
    dict(save_warmup=0), dict(store_dims=2), dict(target_accept=0.95, max_energy_error=5.0),
    dict(num_tune=0, num_draws=50), dict(mass_matrix_switch_freq=20, early_mass_matrix_switch_freq=5),
])
def test_emulated_core_options(opts):
    base = dict(seed=4, num_tune=120, num_draws=60)
    base.update(opts)
    s = O.default_settings(**base)
    a = O.sample(O.Model("funnel", 6), s, 2)
    b = E.sample("funnel", 6, s, 2, smem_slots=4, max_per_launch=50)
    assert np.array_equal(a["draws"], b["draws"]) and np.array_equal(a["stats"], b["stats"])


def test_emulated_core_tape_and_q0():
    rng = np.random.default_rng(0)
    s = O.default_settings(seed=8, num_tune=40, num_draws=30)
    z = rng.normal(size=(2, 70, 5))
    q0 = rng.normal(size=(2, 5))
    a = O.sample(O.Model("normal", 5), s, 2, q0=q0, z_tape=z)
    b = E.sample("normal", 5, s, 2, q0=q0, z_tape=z)
    assert np.array_equal(a["draws"], b["draws"])
    c = O.sample(O.Model("normal", 5), s, 2, q0=q0)
    assert not np.array_equal(a["draws"], c["draws"])  # the tape is really consumed


def test_emulated_radon_tracks_oracle(radon_data):
    """The radon density sums observations in a different order on the device
    (county runs) than the oracle (file order): agreement is to rounding, and the
    trajectories stay together for the first draws before chaos separates them."""
    d = radon_data
    J = d["n_county"]; D = 2 * J + 5
    s = O.default_settings(seed=3, num_tune=60, num_draws=20, init_radius=1.0)
    a = O.sample(O.Model("radon", D, y=d["y"], county=d["county"], floor=d["floor"], n_county=J), s, 3)
    b = E.sample("radon", D, s, 3, smem_slots=4, y=d["y"], county=d["county"], floor=d["floor"], n_county=J)
    dd = np.abs(a["draws"] - b["draws"]).max(axis=(0, 2))
    assert dd[0] < 1e-12 and dd[:5].max() < 1e-8
    assert abs(a["total_steps"] - b["total_steps"]) / a["total_steps"] < 0.25  # 3 short chaotic chains
