"""Pins the oracle's optimizer / motion-model / utils restatement against the reference's own
unit-test expectations (transcribed; file:line cited per test, paths under
/root/reference/nav2_mppi_controller/test/)."""
import math

import numpy as np
import pytest

import navigation2_b200 as nb
from navigation2_b200 import capi, params as P
from oracle.binding import OracleOptimizer, load


def engine(params, critics=(), **kw):
    p = {"controller_frequency": 30.0}
    p.update(params)
    o = OracleOptimizer(P.make_config(p, critics, **kw))
    o.setCostmap(np.zeros((50, 50), dtype=np.uint8), 0.1, 0.0, 0.0)
    return o


def const(B, T, v):
    return np.full((T, B), v, dtype=np.float32)


# optimizer_unit_tests.cpp:205-250 testupdateStateVels (via TEST updateStateVelocitiesTests)
def test_predict_clamps_like_reference():
    o = engine({"batch_size": 1000, "time_steps": 50, "motion_model": "omni"})
    s = o.settings()
    dt = 0.05
    st = {"vx": const(1000, 50, 0.0), "vy": const(1000, 50, 0.0), "wz": const(1000, 50, 0.0),
          "cvx": const(1000, 50, 0.75), "cvy": const(1000, 50, 0.5), "cwz": const(1000, 50, 0.1)}
    st["vx"][0], st["vy"][0], st["wz"][0] = 5.0, 1.0, 6.0     # updateInitialStateVelocities
    out = o.predict(st)
    clamp = lambda v, lo, hi: min(max(v, lo), hi)
    assert abs(out["vx"][0, 0] - 5.0) < 1e-6 and abs(out["vy"][0, 0] - 1.0) < 1e-6 and abs(out["wz"][0, 0] - 6.0) < 1e-6
    assert abs(out["vx"][1, 0] - clamp(0.75, 5.0 + s["ax_min"] * dt, 5.0 + s["ax_max"] * dt)) < 1e-6
    assert abs(out["vy"][1, 0] - clamp(0.5, 1.0 + s["ay_min"] * dt, 1.0 + s["ay_max"] * dt)) < 1e-6
    assert abs(out["wz"][1, 0] - clamp(0.1, 6.0 - s["az_max"] * dt, 6.0 + s["az_max"] * dt)) < 1e-6
    st = {"vx": const(1000, 50, 0.0), "vy": const(1000, 50, 0.0), "wz": const(1000, 50, 0.0),
          "cvx": const(1000, 50, -0.75), "cvy": const(1000, 50, -0.5), "cwz": const(1000, 50, -0.1)}
    st["vx"][0], st["vy"][0], st["wz"][0] = -5.0, -1.0, -6.0
    out = o.predict(st)
    assert abs(out["vx"][1, 0] - clamp(-0.75, -5.0 + s["ax_min"] * dt, -5.0 + s["ax_max"] * dt)) < 1e-6
    assert abs(out["vy"][1, 0] - clamp(-0.5, -1.0 - s["ay_max"] * dt, -1.0 + s["ay_max"] * dt)) < 1e-6
    assert abs(out["wz"][1, 0] - clamp(-0.1, -6.0 - s["az_max"] * dt, -6.0 + s["az_max"] * dt)) < 1e-6


# optimizer_unit_tests.cpp:488-513 shiftControlSequenceTests
def test_shift_control_sequence():
    o = engine({"batch_size": 10, "time_steps": 100, "motion_model": "omni"})
    u = np.zeros((3, 100), dtype=np.float32)
    u[:, 0], u[:, 1], u[:, 2] = 9999, 6, 888
    o.setControlSequence(u)
    o.shiftControlSequence()
    v = o.getOptimalControlSequence()
    for ax in range(3):
        assert v[ax, 0] == 6 and v[ax, 1] == 888 and v[ax, 2] == 0


# optimizer_unit_tests.cpp:515-555 SpeedLimitTests
def test_speed_limit_api():
    o = engine({"batch_size": 1000, "time_steps": 50, "retry_attempt_limit": 2})
    s = o.settings()
    assert s["vx_max"] == np.float32(0.5) and s["vx_min"] == np.float32(-0.35)
    o.setSpeedLimit(0, False)
    s = o.settings()
    assert s["vx_max"] == np.float32(0.5) and s["vx_min"] == np.float32(-0.35)
    o.setSpeedLimit(50.0, True)
    s = o.settings()
    assert abs(s["vx_max"] - 0.25) < 1e-3 and abs(s["vx_min"] + 0.175) < 1e-3
    o.setSpeedLimit(0, True)
    s = o.settings()
    assert s["vx_max"] == np.float32(0.5) and s["vx_min"] == np.float32(-0.35)
    o.setSpeedLimit(0.75, False)
    s = o.settings()
    assert abs(s["vx_max"] - 0.75) < 1e-3 and abs(s["vx_min"] + 0.5249) < 1e-2


# optimizer_unit_tests.cpp:557-627 applyControlSequenceConstraintsTests
def test_apply_control_sequence_constraints():
    o = engine({"batch_size": 1000, "time_steps": 50, "vx_max": 1.0, "vx_min": -1.0, "vy_max": 0.75, "wz_max": 2.0,
                "motion_model": "omni"})
    for sign, inp in ((1.0, (1.0, 0.75, 2.0)), (1.0, (5.0, 5.0, 5.0)), (-1.0, (-5.0, -5.0, -5.0))):
        u = np.zeros((3, 50), dtype=np.float32)
        u[0], u[1], u[2] = inp
        o.setControlSequence(u)
        o.applyControlSequenceConstraints((sign * 1.0, sign * 0.75, sign * 2.0))
        v = o.getOptimalControlSequence()
        assert np.allclose(v[0], sign * 1.0) and np.allclose(v[1], sign * 0.75) and np.allclose(v[2], sign * 2.0)


# optimizer_unit_tests.cpp:707-777 integrateStateVelocitiesTests
def test_integrate_state_velocities_closed_forms():
    o = engine({"batch_size": 1000, "time_steps": 50, "model_dt": 0.1, "motion_model": "omni"})
    B, T = 1000, 50
    vx = const(B, T, 0.1)
    vx[0] = 0.0
    x, y, yaw = o.integrate((0, 0, 0), vx, const(B, T, 0.0), const(B, T, 0.0))
    assert np.allclose(y, 0.0) and np.allclose(yaw, 0.0)
    for i in range(T):
        assert abs(x[i, 1] - i * 0.1 * 0.1) < 1e-3
    vy = const(B, T, 0.2)
    vy[0] = 0.0
    x, y, yaw = o.integrate((0, 0, 0), vx, vy, const(B, T, 0.0))
    assert np.allclose(yaw, 0.0)
    for i in range(T):
        assert abs(x[i, 1] - i * 0.1 * 0.1) < 1e-3 and abs(y[i, 1] - i * 0.2 * 0.1) < 1e-3
    wz = const(B, T, 0.2)
    wz[0] = 0.0
    x, y, yaw = o.integrate((0, 0, 0), vx, const(B, T, 0.0), wz)
    ex, ey = np.float32(0), np.float32(0)
    for i in range(1, T):  # the unicycle recurrence of :765-774, to 1e-6
        ex = np.float32(ex + np.float32(np.float32(0.1) * np.float32(math.cos(0.2 * 0.1 * (i - 1)))) * np.float32(0.1))
        ey = np.float32(ey + np.float32(np.float32(0.1) * np.float32(math.sin(0.2 * 0.1 * (i - 1)))) * np.float32(0.1))
        assert abs(x[i, 1] - ex) < 1e-6 and abs(y[i, 1] - ey) < 1e-6


# optimizer_unit_tests.cpp:929-1054 applyControlSequenceInterIterationConstraintsTests
def test_inter_iteration_constraints_exact_values():
    base = {"model_dt": 0.05, "batch_size": 100, "time_steps": 10, "ax_max": 2.0, "ax_min": -1.0, "ay_max": 2.0,
            "ay_min": -1.0, "az_max": 2.0, "open_loop": True, "motion_model": "omni"}
    o = engine(dict(base, controller_frequency=20.0))
    assert o.settings()["shift_control_sequence"] == 1.0

    def run(o, speed, u0):
        u = np.zeros((3, 10), dtype=np.float32)
        u[:, 0] = u0
        o.setControlSequence(u)
        o.applyControlSequenceConstraints(speed, inter_iteration=True)
        return o.getOptimalControlSequence()[:, 0]

    v = run(o, (0.1, 0.3, 0.2), 5.0)     # shifting: u(0) is pinned to the current speed
    assert v[0] == np.float32(0.1) and v[1] == np.float32(0.3) and v[2] == np.float32(0.2)
    v = run(o, (0.0, 0.0, 0.0), 5.0)
    assert v[0] == 0.0 and v[1] == 0.0 and v[2] == 0.0
    # not shifting (controller at 40 Hz): u(0) is accel-clamped with the controller period
    o2 = engine(dict(base, controller_frequency=40.0))
    assert o2.settings()["shift_control_sequence"] == 0.0
    v = run(o2, (0.0, 0.0, 0.0), 5.0)
    assert np.allclose(v, 0.05, atol=1e-6)
    u = np.zeros((3, 10), dtype=np.float32)
    u[0, 0] = -5.0
    o2.setControlSequence(u)
    o2.applyControlSequenceConstraints((0.5, 0.0, 0.0), inter_iteration=True)
    assert abs(o2.getOptimalControlSequence()[0, 0] - 0.475) < 1e-6   # hard decel: 0.5 - 0.025 * 1.0
    u[0, 0] = 5.0
    o2.setControlSequence(u)
    o2.applyControlSequenceConstraints((0.5, 0.0, 0.0), inter_iteration=True)
    assert abs(o2.getOptimalControlSequence()[0, 0] - 0.55) < 1e-6    # 0.5 + 0.025 * 2.0


# optimizer_unit_tests.cpp:803-877 Omni_openLoopMppiTest: with open_loop the measured robot speed (absurd here) is
# ignored in favour of the last command, and at 30 Hz (no shifting) every command stays within one controller
# period of acceleration from the previous one
def test_omni_open_loop_commands_are_acceleration_bounded():
    cfg = P.make_config({"controller_frequency": 30.0, "batch_size": 1000, "model_dt": 0.05, "time_steps": 80,
                         "ax_max": 0.5, "az_max": 0.5, "ay_max": 0.5, "open_loop": True, "motion_model": "omni"})
    o = OracleOptimizer(cfg)
    o.setCostmap(np.zeros((50, 50), dtype=np.uint8), 0.1, 0.0, 0.0)
    period = o.settings()["controller_period"]
    assert o.settings()["shift_control_sequence"] == 0.0
    path = np.zeros((17, 3), dtype=np.float64)
    inp = nb.CycleInput(pose=(999.0, 0.0, 0.0), speed=(100.0, 100.0, 100.0), path=path, goal=(0.0, 0.0, 0.0))
    c1 = o.evalControl(inp).cmd
    assert abs(c1[0]) <= period * 0.5 and abs(c1[1]) <= period * 0.5 and abs(c1[2]) <= period * 0.5
    c2 = o.evalControl(inp).cmd
    for a, b in zip(c1, c2):
        assert abs(b - a) <= period * 0.5


# optimizer_unit_tests.cpp:377-381 setOffset: controller period > model_dt is a ControllerException
def test_set_offset_rejects_slow_controller():
    with pytest.raises(Exception):
        OracleOptimizer(P.make_config({"model_dt": 0.05, "controller_frequency": 5.0}))
    OracleOptimizer(P.make_config({"model_dt": 0.05, "controller_frequency": 50.0}))


UNBOUNDED = {"vx_max": 100, "vx_min": -100, "vy_max": 100, "wz_max": 100, "ax_max": 1000, "ax_min": -1000,
             "ay_min": -1000, "ay_max": 1000, "az_max": 1000}


# motion_model_tests.cpp:263-302 DelayReplayAllAxes
def test_delay_replay_all_axes():
    o = engine(dict(UNBOUNDED, batch_size=8, time_steps=20, model_dt=0.05, model_delay_vx=0.10, model_delay_vy=0.10,
                    model_delay_wz=0.15, motion_model="omni", controller_frequency=20.0))
    o.pushCommandHistory(1.0, 2.0, 3.0)
    o.pushCommandHistory(5.0, 6.0, 7.0)
    o.pushCommandHistory(9.0, 8.0, 4.0)
    z = lambda: np.zeros((20, 8), dtype=np.float32)
    out = o.predict({"vx": z(), "vy": z(), "wz": z(), "cvx": z(), "cvy": z(), "cwz": z()})
    assert np.all(out["vx"][0] == 0) and np.all(out["vy"][0] == 0) and np.all(out["wz"][0] == 0)
    assert np.all(out["vx"][1] == 9.0) and np.all(out["vy"][1] == 8.0) and np.all(out["wz"][1] == 7.0)
    assert np.all(out["wz"][2] == 4.0)
    for k, cols in (("vx", (2, 3)), ("vy", (2, 3)), ("wz", (3,))):
        for c in cols:
            assert np.all(out[k][c] == 0.0)


# motion_model_tests.cpp:304-349 DelayVyIgnoredOnDiffDrive / zero delay is a no-op
def test_zero_delay_keeps_rollout():
    o = engine(dict(UNBOUNDED, batch_size=4, time_steps=10, model_dt=0.05, motion_model="diff_drive",
                    controller_frequency=20.0))
    o.pushCommandHistory(123.0, 0.0, 0.0)
    st = {k: np.zeros((10, 4), dtype=np.float32) for k in ("vx", "vy", "wz", "cvx", "cvy", "cwz")}
    st["vx"][0] = 2.0
    st["cvx"][:] = 2.0
    out = o.predict(st)
    assert np.all(out["vx"] == 2.0)


# motion_model_tests.cpp:351-381 ClampRawControlsRewritesRawCommand
def test_clamp_raw_controls_rewrites_raw_command():
    o = engine({"vx_max": 100, "vx_min": -100, "vy_max": 100, "wz_max": 100, "ax_max": 1.0, "ax_min": -1.0,
                "ay_min": -1.0, "ay_max": 1.0, "az_max": 1.0, "batch_size": 1, "time_steps": 3, "model_dt": 1.0,
                "clamp_raw_controls": True, "motion_model": "omni", "controller_frequency": 1.0})
    st = {k: np.zeros((3, 1), dtype=np.float32) for k in ("vx", "vy", "wz", "cvx", "cvy", "cwz")}
    st["cvx"][0:2] = 5.0
    st["cwz"][0:2] = 5.0
    out = o.predict(st)
    assert out["vx"][1, 0] == 1.0 and out["vx"][2, 0] == 2.0 and out["wz"][1, 0] == 1.0 and out["wz"][2, 0] == 2.0
    assert out["cvx"][0, 0] == 1.0 and out["cvx"][1, 0] == 2.0 and out["cwz"][0, 0] == 1.0 and out["cwz"][1, 0] == 2.0


# motion_model_tests.cpp:124-252 Ackermann applyConstraints: |vx| / |wz| >= min_turning_r afterwards
def test_ackermann_apply_constraints():
    o = engine({"batch_size": 10, "time_steps": 30, "motion_model": "ackermann", "vx_max": 10.0, "wz_max": 100.0,
                "ax_max": 1000.0, "ax_min": -1000.0, "az_max": 1000.0})
    rng = np.random.default_rng(0)
    u = np.zeros((3, 30), dtype=np.float32)
    u[0] = rng.uniform(0.05, 0.5, 30)
    u[2] = rng.uniform(-5, 5, 30)
    o.setControlSequence(u)
    o.applyControlSequenceConstraints((float(u[0, 0]), 0.0, float(u[2, 0])))
    v = o.getOptimalControlSequence()
    ok = np.abs(v[0, 1:]) / np.maximum(np.abs(v[2, 1:]), 1e-9) >= 0.2 - 1e-5
    assert ok.all()


# utils_test.cpp:128-178 AnglesTests
def test_angle_helpers():
    lib = load()
    for i in range(100):
        a = float(i * i) * (-1.0 if i % 2 == 0 else 1.0)
        n = lib.oracle_normalize_angle(a)
        assert -math.pi - 1e-6 <= n <= math.pi + 1e-6
        d = lib.oracle_shortest_angular_distance(a, 0.0)
        assert -math.pi - 1e-6 <= d <= math.pi + 1e-6
    ppa = lib.oracle_pose_point_angle
    assert abs(ppa(0, 0, 0, 1.0, 0.0, 1)) < 1e-6
    assert abs(ppa(0, 0, 0, 1.0, 0.0, 0)) < 1e-6
    assert abs(ppa(0, 0, 0, -1.0, 0.0, 0)) < 1e-6
    assert abs(ppa(0, 0, 0, -1.0, 0.0, 1) - math.pi) < 1e-6
    ppy = lib.oracle_pose_point_angle_yaw
    assert abs(ppy(0, 0, 0, 1.0, 0.0, 0.0)) < 1e-6
    assert abs(ppy(0, 0, 0, 1.0, 0.0, math.pi) - math.pi) < 1e-6
    assert abs(ppy(0, 0, 0, 1.0, 0.0, 0.1)) < 1e-3
    assert abs(ppy(0, 0, 0, 1.0, 0.0, 3.04159) - math.pi) < 1e-3


# utils_test.cpp:180-226 FurthestAndClosestReachedPoint: == 5; :230-267 U-turn: <= 6
def test_find_path_furthest_reached_point():
    cfg = P.make_config({"batch_size": 100, "time_steps": 2, "model_dt": 0.1, "controller_frequency": 10.0},
                        ["PathAngleCritic"])
    o = OracleOptimizer(cfg)
    o.setCostmap(np.zeros((50, 50), dtype=np.uint8), 0.1, 0.0, 0.0)
    x = np.zeros((2, 100), dtype=np.float32)
    x[1] = 1.0
    z = np.zeros((2, 100), dtype=np.float32)
    path = np.zeros((10, 3))
    path[:, 0] = 0.2 * np.arange(10)
    inp = nb.CycleInput(pose=(0, 0, 0), speed=(0, 0, 0), path=path, local_path_length=5.0)
    _, _, _, furthest = o.scoreTrajectories(inp, {"vx": z, "vy": z, "wz": z}, {"x": x, "y": z, "yaws": z})
    assert furthest == 5

    cfg = P.make_config({"batch_size": 10, "time_steps": 4, "model_dt": 0.1, "controller_frequency": 10.0},
                        ["PathAngleCritic"])
    o = OracleOptimizer(cfg)
    o.setCostmap(np.zeros((50, 50), dtype=np.uint8), 0.1, 0.0, 0.0)
    px = [0, 1, 2, 3, 4, 5, 6, 7, 8, 8, 7, 6, 5, 4, 3, 2, 1, 0]
    py = [0] * 9 + [1] * 9
    path = np.stack([px, py, np.zeros(18)], axis=1).astype(np.float64)
    x = np.zeros((4, 10), dtype=np.float32)
    y = np.zeros((4, 10), dtype=np.float32)
    for k in range(1, 4):
        x[k] = np.float32(5.0 * k / 3.0)
        y[k] = np.float32(1.0 * k / 3.0)
    z = np.zeros((4, 10), dtype=np.float32)
    inp = nb.CycleInput(pose=(0, 0, 0), speed=(0, 0, 0), path=path, local_path_length=17.0)
    _, _, _, furthest = o.scoreTrajectories(inp, {"vx": z, "vy": z, "wz": z}, {"x": x, "y": y, "yaws": z})
    assert furthest <= 6


# utils_test.cpp:394-467 shiftColumnsByOnePlace is exercised through shiftControlSequence above;
# utils_test.cpp:336-391 SmootherTest: the Savitzky-Golay filter smooths a noisy sequence
def test_savitsky_golay_reduces_noise():
    o = engine({"batch_size": 10, "time_steps": 30, "motion_model": "omni", "controller_frequency": 20.0})
    rng = np.random.default_rng(3)
    u = np.zeros((3, 30), dtype=np.float32)
    u[0] = 1.0 + rng.normal(0, 0.1, 30)
    u[1] = 1.0 + rng.normal(0, 0.1, 30)
    u[2] = 1.0 + rng.normal(0, 0.1, 30)
    hist = np.ones((4, 3), dtype=np.float32)
    o.setControlSequence(u)
    new_hist = o.savitskyGolay(hist)
    v = o.getOptimalControlSequence()
    for ax in range(3):
        assert np.abs(np.diff(v[ax, :29])).sum() < np.abs(np.diff(u[ax, :29])).sum()
        assert v[ax, 29] == u[ax, 29]           # the last element is never filtered (utils.hpp:478-501)
    assert np.all(new_hist[0:3] == 1.0)          # history shifts left ...
    assert np.allclose(new_hist[3], v[:, 1])     # ... and takes the filtered u(offset = 1) (utils.hpp:534-541)
    # too short to smooth: untouched (utils.hpp:478-481)
    o2 = engine({"batch_size": 10, "time_steps": 20, "controller_frequency": 20.0})
    u2 = rng.normal(0, 1, (3, 20)).astype(np.float32)
    o2.setControlSequence(u2)
    o2.savitskyGolay(hist)
    assert np.array_equal(o2.getOptimalControlSequence(), u2)


# optimizer_smoke_test.cpp:38-112: every motion model x critic set produces a command
@pytest.mark.parametrize("model", ["diff_drive", "omni", "ackermann"])
def test_optimizer_smoke(model):
    critics = ["GoalCritic", "GoalAngleCritic", "ObstaclesCritic", "PathAngleCritic", "PathFollowCritic",
               "PreferForwardCritic", "CostCritic", "ConstraintCritic", "TwirlingCritic", "PathAlignCritic",
               "VelocityDeadbandCritic"]
    cfg = P.make_config({"batch_size": 400, "time_steps": 15, "motion_model": model, "controller_frequency": 20.0},
                        critics, robot_radius=0.15)
    o = OracleOptimizer(cfg)
    cells = np.zeros((40, 40), dtype=np.uint8)
    cells[16:24, 16:24] = 250
    o.setCostmap(cells, 0.1, 0.0, 0.0)
    path = np.stack([np.linspace(1.0, 3.0, 17), np.linspace(1.0, 3.0, 17), np.zeros(17)], axis=1)
    r = o.evalControl(nb.CycleInput(pose=(1.0, 1.0, 0.785), speed=(0, 0, 0), path=path, goal_checker_xy_tolerance=0.25))
    assert r.status == 0 and np.isfinite(r.control_sequence).all() and np.isfinite(r.optimal_trajectory).all()


# optimizer_unit_tests.cpp:395-433 fallbackTests: retries up to retry_attempt_limit, then NoValidControl
def test_fallback_and_retry_limit():
    cfg = P.make_config({"batch_size": 64, "time_steps": 30, "retry_attempt_limit": 2, "controller_frequency": 20.0},
                        ["CostCritic"], robot_radius=0.1)
    o = OracleOptimizer(cfg)
    o.setCostmap(np.full((40, 40), 254, dtype=np.uint8), 0.1, 0.0, 0.0)   # every trajectory collides
    path = np.stack([np.linspace(1, 2, 10), np.ones(10), np.zeros(10)], axis=1)
    r = o.evalControl(nb.CycleInput(pose=(1.0, 1.0, 0.0), speed=(0, 0, 0), path=path), raise_on_failure=False)
    assert r.status == capi.MPPI_ERR_NO_VALID_CONTROL
    assert r.fail_flag and r.retries == 2
