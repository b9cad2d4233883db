"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle): the oracle
must keep reproducing them bit for bit (CPU), and the CUDA path must match them — integer cell indices and collision
flags exactly, the B x T tensors exactly, costs to 1e-4 relative, commands / control sequences to 1e-5 (GPU)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as G  # noqa: E402

NAMES = sorted(G.scenarios().keys())


def load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(name):
    from oracle.binding import OracleOptimizer
    got = G.run(OracleOptimizer, G.scenarios()[name])
    want = load(name)
    assert sorted(got) == sorted(want)
    for k in want:
        np.testing.assert_array_equal(got[k], want[k], err_msg=f"{name}:{k}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_matches_golden(name):
    import navigation2_b200 as nb
    sc = G.scenarios()[name]
    got = G.run(nb.Optimizer, sc)
    want = load(name)
    for k in ("cost_cells_0", "collisions_0", "x_0", "yaw_0"):
        np.testing.assert_array_equal(got[k], want[k], err_msg=f"{name}:{k}")
    c, w = got["costs_0"].astype(np.float64), want["costs_0"].astype(np.float64)
    assert np.max(np.abs(c - w) / np.maximum(np.abs(w), 1e-6)) <= 1e-4
    for cyc in range(sc["cycles"]):
        # free-running: later cycles start from control sequences that differ in the last ulps
        tol = 1e-5 if cyc == 0 else 5e-5
        np.testing.assert_allclose(got[f"u_{cyc}"], want[f"u_{cyc}"], rtol=0, atol=tol, err_msg=f"{name}: control sequence {cyc}")
        np.testing.assert_allclose(got[f"cmd_{cyc}"], want[f"cmd_{cyc}"], rtol=0, atol=tol)
        np.testing.assert_allclose(got[f"traj_{cyc}"], want[f"traj_{cyc}"], rtol=0, atol=5e-5)
