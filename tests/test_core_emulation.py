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


# ---- the same core run by T cooperative lanes per chain (GroupLanes, tests/emul/emul.cpp): the
# ---- per-thread loops (NIT), shared-memory tier, reductions and density layouts of the GPU geometry

STAT = {n: i for i, n in enumerate(O.STAT_NAMES)}


@pytest.mark.parametrize("kind,dim,kw,tpc", [
    ("normal", 37, dict(mu=1.0, sigma=2.0), 32),   # NIT = 2, predicated tail
    ("normal", 100, {}, 32),                      # run-time trip count
    ("normal", 40, {}, 64),                       # two warps per chain: cross-warp reduction
])
def test_lane_emulation_reproduces_oracle_trees(kind, dim, kw, tpc):
    """A warp (or two) per chain: every lane owns dimensions tid, tid + T, ...; sums are taken in
    the butterfly order of GroupCuda::warp_reduce.  The trees must be the oracle's, draw for
    draw; positions differ by the rounding of the differently ordered sums."""
    s = O.default_settings(seed=5, num_tune=80, num_draws=40)
    a = O.sample(O.Model(kind, dim, **kw), s, 2)
    b = E.sample_lanes(kind, dim, s, 2, threads_per_chain=tpc, smem_slots=3, **kw)
    for k in ("depth", "n_steps", "index_in_trajectory", "diverging", "maxdepth_reached"):
        assert np.array_equal(a["stats"][..., STAT[k]], b["stats"][..., STAT[k]]), k
    np.testing.assert_allclose(b["draws"][:, :3], a["draws"][:, :3], rtol=0, atol=1e-10)
    np.testing.assert_allclose(b["draws"], a["draws"], rtol=0, atol=1e-5)  # rounding growth over 120 draws
    np.testing.assert_allclose(b["stats"][..., STAT["step_size"]], a["stats"][..., STAT["step_size"]], rtol=1e-6)


def test_lane_emulation_pause_resume_is_bit_identical():
    s = O.default_settings(seed=9, num_tune=60, num_draws=30)
    a = E.sample_lanes("funnel", 9, s, 2, threads_per_chain=32, smem_slots=5)
    b = E.sample_lanes("funnel", 9, s, 2, threads_per_chain=32, smem_slots=5, max_per_launch=17)
    assert np.array_equal(a["draws"], b["draws"]) and np.array_equal(a["stats"], b["stats"])
    c = O.sample(O.Model("funnel", 9), s, 2)
    np.testing.assert_allclose(a["draws"][:, :3], c["draws"][:, :3], rtol=0, atol=1e-10)


@pytest.mark.parametrize("tpc,slots", [(32, 3), (64, 0)])
def test_lane_emulation_radon_production_geometry(radon_data, tpc, slots):
    """The radon density exactly as the B200 runs it — observations cut into one contiguous range
    per lane, prefix sums published at group ends, per-county gradients from the (county, floor)
    piece lists, 6 unrolled dimensions per lane, momentum tier in 'shared memory' — against the
    oracle's file-order evaluation: equal to rounding on the first draws."""
    d = radon_data
    J = d["n_county"]; D = 2 * J + 5
    s = O.default_settings(seed=3, num_tune=40, num_draws=10, init_radius=1.0)
    a = O.sample(O.Model("radon", D, y=d["y"], county=d["county"], floor=d["floor"], n_county=J), s, 2)
    b = E.sample_lanes("radon", D, s, 2, threads_per_chain=tpc, smem_slots=slots, y=d["y"],
                       county=d["county"], floor=d["floor"], n_county=J)
    dd = np.abs(a["draws"] - b["draws"]).max(axis=(0, 2))
    # (global-prefix differences: sums of ~900 residuals cancel down to one pair's sum, which costs
    # a few more ulps than adding the pair's own residuals as the oracle does)
    assert dd[0] < 1e-11 and dd[:5].max() < 1e-8, dd[:5]
    assert np.array_equal(a["stats"][:, :5, STAT["n_steps"]], b["stats"][:, :5, STAT["n_steps"]])
    assert abs(a["total_steps"] - b["total_steps"]) / a["total_steps"] < 0.25


@pytest.mark.parametrize("kind,dim,tpc,slots", [("radon", 175, 32, 3), ("radon", 175, 64, 2), ("funnel", 9, 32, 4),
                                                 ("normal", 100, 32, 3), ("normal", 40, 64, 3)])
def test_lane_schedule_independence(radon_data, kind, dim, tpc, slots):
    """Race check of the core.  Between two barriers the emulation runs the lanes of a chain one
    after another; ascending, descending and freshly shuffled orders after every barrier are all
    schedules independent thread scheduling allows, so a core with every needed barrier gives
    bit-identical results under each.  (Negative control below: one barrier removed.)"""
    kw = {}
    if kind == "radon":
        d = radon_data
        kw = dict(y=d["y"], county=d["county"], floor=d["floor"], n_county=d["n_county"])
    s = O.default_settings(seed=3, num_tune=40, num_draws=10, init_radius=1.0)
    runs = [E.sample_lanes(kind, dim, s, 2, threads_per_chain=tpc, smem_slots=slots, lane_order=o,
                           max_per_launch=23, **kw) for o in (0, 1, 2, 7)]
    for r in runs[1:]:
        assert np.array_equal(runs[0]["draws"], r["draws"]) and np.array_equal(runs[0]["stats"], r["stats"])


def test_lane_schedule_check_detects_a_missing_barrier():
    """The same check on a build of the core without the barrier that publishes a new leaf's
    scalars (NB200_EMUL_DROP_BARRIER): the results now depend on the lane order."""
    s = O.default_settings(seed=3, num_tune=40, num_draws=10)
    a = E.sample_lanes("normal", 37, s, 2, threads_per_chain=32, smem_slots=3, lane_order=0, drop_barrier=True)
    b = E.sample_lanes("normal", 37, s, 2, threads_per_chain=32, smem_slots=3, lane_order=1, drop_barrier=True)
    assert not np.array_equal(a["draws"], b["draws"])


def _radon_edge_cases():
    rng = np.random.default_rng(1)
    J = 85
    return [
        ("few_observations", [0, 3, 3, 84, 84], [0, 1, 0, 0, 1], J),          # most lanes own nothing
        ("one_observation", [42], [1], J),
        ("one_county_holds_everything", [7] * 300, rng.integers(0, 2, 300), J),  # a pair cut into 32 pieces
        ("all_floor_0", rng.integers(0, J, 200), [0] * 200, J),
        ("all_floor_1", rng.integers(0, J, 200), [1] * 200, J),
        ("empty_counties", rng.integers(0, J // 2, 500) * 2, rng.integers(0, 2, 500), J),
        ("n_33", rng.integers(0, J, 33), rng.integers(0, 2, 33), J),            # one more than the lanes
        ("n_1000_unsorted", rng.integers(0, J, 1000), rng.integers(0, 2, 1000), J),
        ("three_counties", rng.integers(0, 3, 40), rng.integers(0, 2, 40), 3),   # D = 11: one dimension per lane
    ]


@pytest.mark.parametrize("name,county,floor,J", _radon_edge_cases(), ids=[c[0] for c in _radon_edge_cases()])
def test_lane_emulation_radon_edge_layouts(name, county, floor, J):
    """Ragged and degenerate inputs of the radon density's host 'compile' step (radon_layout.hpp)
    at the GPU's lane geometry: lanes without observations, a (county, floor) pair cut into a piece
    per lane, counties without observations, a single observation.  Same trees as the oracle,
    positions equal to rounding."""
    rng = np.random.default_rng(7)
    county = np.asarray(county, dtype=np.int32)
    floor = np.asarray(floor, dtype=np.uint8)
    y = rng.normal(1.0, 0.8, size=len(county))
    D = 2 * J + 5
    s = O.default_settings(seed=2, num_tune=30, num_draws=10, init_radius=0.5)
    kw = dict(y=y, county=county, floor=floor, n_county=J)
    a = O.sample(O.Model("radon", D, **kw), s, 2)
    b = E.sample_lanes("radon", D, s, 2, threads_per_chain=32, smem_slots=2, **kw)
    dd = np.abs(a["draws"] - b["draws"]).max(axis=(0, 2))
    assert dd[0] < 1e-11 and dd[:5].max() < 1e-7, dd[:5]  # rounding grows along the chain
    assert np.array_equal(a["stats"][:, :5, STAT["n_steps"]], b["stats"][:, :5, STAT["n_steps"]])


def test_core_is_clean_under_address_and_ub_sanitizers():
    """Serial and lane emulation of the core from a library built with
    -fsanitize=address,undefined: every index into the slot pool, the shared-memory tier, the
    front buffer and the density scratch is in bounds for the geometries the GPU runs (radon at
    32 / 64 lanes with 0, 3 and all slots on chip, maxdepth 12, chunked launches)."""
    import os
    import subprocess
    import sys
    from pathlib import Path

    here = Path(E.__file__).resolve().parent
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not asan or not Path(asan).exists():
        pytest.skip("libasan not available")
    subprocess.run(["make", "-C", str(here), "libnuts_emul_asan.so"], check=True, stdout=subprocess.PIPE,
                   stderr=subprocess.STDOUT)
    env = dict(os.environ, LD_PRELOAD=asan,
               ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=1")
    r = subprocess.run([sys.executable, str(here / "sanitize_run.py"), str(here / "libnuts_emul_asan.so")],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and "SANITIZE-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
    assert "runtime error" not in r.stderr and "ERROR: AddressSanitizer" not in r.stderr, r.stderr[-4000:]


def test_lane_emulation_device_expand_matches_host_expand(radon_data):
    """expand_vector (src/pymc.rs:217-286) as the kernel runs it at draw write-out
    (RadonModel::expand, every lane writing its share of the 4J + 5 values) against the host twin
    CompiledDeviceModel._expand applied to the unconstrained draws of the same run."""
    import nutpie_b200

    d = radon_data
    J = d["n_county"]; D = 2 * J + 5
    kw = dict(y=d["y"], county=d["county"], floor=d["floor"], n_county=J)
    s = O.default_settings(seed=6, num_tune=20, num_draws=10, init_radius=1.0)
    raw = E.sample_lanes("radon", D, s, 2, threads_per_chain=32, smem_slots=3, **kw)
    exp = E.sample_lanes("radon", D, s, 2, threads_per_chain=32, smem_slots=3, expand=True, **kw)
    assert exp["draws"].shape[-1] == 4 * J + 5 and np.array_equal(raw["stats"], exp["stats"])
    cm = nutpie_b200.radon_model(d["y"], d["county"], d["floor"], J)
    host = cm._expand(raw["draws"])
    dev = cm._split_expanded(exp["draws"])
    assert set(host) == set(dev)
    for name in host:
        np.testing.assert_allclose(dev[name], host[name], rtol=1e-15, atol=0, err_msg=name)
