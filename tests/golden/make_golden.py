"""Generates tests/golden/*.json (run in the build container, where /root/reference exists).

The reference's golden arrays pin seeded nuts-rs bit streams we cannot reproduce
(SURVEY.md §8c); what travels to the GPU box is their summary, used as a
distributional pin.  Usage: python tests/golden/make_golden.py
"""
import json
from pathlib import Path

import numpy as np

REF = Path("/root/reference/tests/reference")
OUT = Path(__file__).resolve().parent

x = np.loadtxt(REF / "test_deterministic_sampling_numba.txt")
y = np.loadtxt(REF / "test_deterministic_sampling_jax.txt")
summary = {
    "source": "tests/reference/test_deterministic_sampling_numba.txt (pymc HalfNormal('a'), "
              "seed=123, draws=100, tune=100, 2 chains; tests/test_pymc.py:533-541)",
    "n": int(x.size), "mean": float(x.mean()), "std": float(x.std()),
    "min": float(x.min()), "max": float(x.max()),
    "n_repeated": int((x[1:] == x[:-1]).sum()),
    "identical_to_jax_file": bool(np.array_equal(x, y)),
    "first5": x[:5].tolist(),
}
(OUT / "halfnormal_reference_summary.json").write_text(json.dumps(summary, indent=1))
# the 200 values themselves (test OUTPUT data of the reference, chain-major 2 x 100): the
# replicate-distribution test needs per-chain autocorrelations, not only the moments
np.savetxt(OUT / "halfnormal_reference_values.txt", x, fmt="%.9g")
print(summary)
