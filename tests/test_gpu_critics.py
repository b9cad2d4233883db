"""The reference's critic known-answer tests (critics_tests.cpp) through the CUDA scoring pass
(mppi_score_tensors -> k_rollout_score<FROM_TENSORS> + k_path_score), and a randomised comparison
of every critic against the oracle on the same tensors."""
import numpy as np
import pytest

import navigation2_b200 as nb
from navigation2_b200 import params as P, scenarios as S

from test_oracle_critics import CriticKATs, B, T, make_engine, cycle, score
from util import ALL_CRITICS, SQUARE_FOOTPRINT, assert_costs_close
from oracle.binding import OracleOptimizer

pytestmark = pytest.mark.gpu


class TestGpuCritics(CriticKATs):
    engine_cls = nb.Optimizer


@pytest.mark.parametrize("model", ["diff_drive", "omni", "ackermann"])
@pytest.mark.parametrize("critic", ALL_CRITICS)
def test_each_critic_matches_oracle_on_random_tensors(critic, model):
    """One critic at a time on random but plausible tensors over an inflated map."""
    rng = np.random.default_rng(hash((critic, model)) % (2 ** 32))
    fp = critic in ("CostCritic", "ObstaclesCritic")
    kw = dict(footprint=SQUARE_FOOTPRINT) if fp else dict(robot_radius=0.22)
    p = {"batch_size": 512, "time_steps": 40, "model_dt": 0.1, "controller_frequency": 10.0, "motion_model": model,
         "visualize": 1}
    ov = {critic: {"consider_footprint": True} if fp else {}}
    if critic == "VelocityDeadbandCritic":
        ov[critic] = {"deadband_velocities": [0.1, 0.1, 0.2]}
    cfg = P.make_config(p, [critic], ov, inflation={"cost_scaling_factor": 3.0, "inflation_radius": 0.7}, **kw)
    cells = S.block_costmap(200, 200, 0.05, seed=11, keep_clear=(5.0, 5.0), n_blocks=60, clear_radius=0.3,
                            inscribed_radius=cfg.inscribed_radius)
    g, c = nb.Optimizer(cfg), OracleOptimizer(cfg)
    for e in (g, c):
        e.setCostmap(cells, 0.05, 0.0, 0.0)
    Tn, Bn = 40, 512
    vx = rng.normal(0.2, 0.3, (Tn, Bn)).astype(np.float32)
    vy = rng.normal(0.0, 0.2, (Tn, Bn)).astype(np.float32)
    wz = rng.normal(0.0, 0.8, (Tn, Bn)).astype(np.float32)
    yaw = np.cumsum(wz * 0.1, axis=0).astype(np.float32)
    x = (5.0 + np.cumsum(vx * np.cos(yaw) * 0.1, axis=0) + rng.normal(0, 0.3, (1, Bn))).astype(np.float32)
    y = (5.0 + np.cumsum(vx * np.sin(yaw) * 0.1, axis=0) + rng.normal(0, 0.3, (1, Bn))).astype(np.float32)
    path = S.arc_path((5.0, 5.0, 0.0), n_points=60, curvature=0.3)
    for lpl in (None, 0.3):
        inp = nb.CycleInput(pose=(5.0, 5.0, 0.1), speed=(0.1, 0.0, 0.0), path=path, local_path_length=lpl,
                            goal_checker_xy_tolerance=0.25)
        st = {"vx": vx, "vy": vy, "wz": wz}
        tr = {"x": x, "y": y, "yaws": yaw}
        gc, gcoll, gfail, gf = g.scoreTrajectories(inp, st, tr)
        cc, ccoll, cfail, cf = c.scoreTrajectories(inp, st, tr)
        assert gf == cf, f"furthest {gf} vs {cf}"
        assert gfail == cfail
        assert np.array_equal(gcoll, ccoll), "collision flags"
        assert_costs_close(gc, cc, label=f"{critic}/{model}/lpl={lpl}")
    g.shutdown()
    c.shutdown()


def test_exact_division_and_sqrt_sequences_match_ieee_operators():
    """The fused kernel's straight-line division / square root against `/` and sqrtf on the
    device: world-to-map quotients near cell boundaries, random magnitudes, zeros, tiny and huge values."""
    import ctypes as C
    from navigation2_b200 import capi
    lib = capi.load_library()
    rng = np.random.default_rng(123)
    n = 1 << 22
    res = rng.choice(np.array([0.05, 0.1, 0.025, 0.02, 0.2, 1.0 / 3.0, 0.0499999], dtype=np.float32), size=n)
    cells = rng.integers(0, 4000, size=n).astype(np.float32)
    # numerators a hair's breadth around exact cell boundaries, plus log-uniform magnitudes
    a = (cells * res).astype(np.float32)
    a = np.nextafter(a, np.where(rng.random(n) < 0.5, np.float32(np.inf), np.float32(-np.inf))).astype(np.float32)
    k = n // 4
    a[:k] = np.exp(rng.uniform(-40, 40, size=k)).astype(np.float32)
    res[:k // 2] = np.exp(rng.uniform(-30, 30, size=k // 2)).astype(np.float32)
    a[k:k + 64] = 0.0
    a[k + 64:k + 128] = np.float32(1e-30)
    a[k + 128:k + 192] = np.float32(3e38)
    mism, flagged = C.c_uint(0), C.c_uint(0)
    rc = lib.mppi_selftest_exact_math(a.ctypes.data_as(C.c_void_p), res.ctypes.data_as(C.c_void_p), n, C.byref(mism), C.byref(flagged))
    assert rc == 0, lib.mppi_last_error(None)
    assert mism.value == 0
    assert 0 < flagged.value < n // 100      # only the planted extremes are declined
