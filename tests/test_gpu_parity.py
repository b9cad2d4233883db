"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs with the same injected noise (BASELINE.json north_star):
  - integer costmap indices and collision flags bit-exact,
  - per-trajectory costs within 1e-4 relative,
  - cmd_vel and the control sequence within 1e-5 absolute.
"""
import numpy as np
import pytest

from navigation2_b200 import capi, params as P, scenarios as S
import navigation2_b200 as nb

from util import ALL_CRITICS, SQUARE_FOOTPRINT, Pair, assert_costs_close, assert_cycle_close, build_map

pytestmark = pytest.mark.gpu

TENSORS = [("vx", capi.TENSOR_VX), ("vy", capi.TENSOR_VY), ("wz", capi.TENSOR_WZ),
           ("cvx", capi.TENSOR_CVX), ("cvy", capi.TENSOR_CVY), ("cwz", capi.TENSOR_CWZ),
           ("x", capi.TENSOR_X), ("y", capi.TENSOR_Y), ("yaw", capi.TENSOR_YAW)]


def compare_state(pair, label, cells_cost=True, cells_obs=False, exact=True):
    """exact=True: both sides started the iteration from bit-identical state, so every tensor and
    every cell index must match bit for bit.  exact=False (second iteration of one call: the
    control sequences already differ in the last ulp): tensors to 2e-5, cell indices may differ
    only for the handful of samples that sit within that distance of a cell boundary."""
    for name, which in TENSORS:
        g = pair.gpu.getTensor(which)
        c = pair.cpu.getTensor(which)
        if exact:
            # same arithmetic, same order, no FMA contraction on either side: bit-identical
            bad = np.flatnonzero(g.view(np.uint32) != c.view(np.uint32))
            assert bad.size == 0, f"{label} tensor {name}: {bad.size} differing elements, first at {bad[:5]}: " \
                                  f"{g.ravel()[bad[:5]]} vs {c.ravel()[bad[:5]]}"
        else:
            # positions are ~10 m (float ulp 1e-6): allow a few ulps once the sequences have drifted apart
            np.testing.assert_allclose(g, c, rtol=0, atol=2e-5, err_msg=f"{label} tensor {name}")
    for on, which, what in ((cells_cost, capi.TENSOR_COST_CELLS, "CostCritic"),
                            (cells_obs, capi.TENSOR_OBSTACLE_CELLS, "ObstaclesCritic")):
        if not on:
            continue
        g = pair.gpu.getTensor(which)
        c = pair.cpu.getTensor(which)
        if exact:
            assert np.array_equal(g, c), f"{label} {what} cell indices differ at {np.argwhere(g != c)[:5]}"
        else:
            assert np.count_nonzero(g != c) <= max(2, g.size // 20000), f"{label} {what} cell indices diverge"
    if not exact:
        return
    g = pair.gpu.getTensor(capi.TENSOR_COLLISIONS)
    c = pair.cpu.getTensor(capi.TENSOR_COLLISIONS)
    assert np.array_equal(g, c), f"{label} collision flags differ"
    assert_costs_close(pair.gpu.getTensor(capi.TENSOR_COSTS), pair.cpu.getTensor(capi.TENSOR_COSTS), label=label)
    gc = pair.gpu.getTensor(capi.TENSOR_CRITIC_COSTS)
    cc = pair.cpu.getTensor(capi.TENSOR_CRITIC_COSTS)
    for k in range(gc.shape[0]):
        assert_costs_close(gc[k], cc[k], rel=1e-4, label=f"{label} critic {k}")


def drive(pair, path, pose, speed, cycles, label, goal_tol=0.25, resync=True, **cmp):
    """Closed loop: integrate the commanded twist like a perfect robot so later cycles see new poses.
    resync=True hands the oracle's carried state (control sequence, SG history, last command) to
    the GPU after every cycle, so each cycle is an independent bit-exact comparison; resync=False
    lets the two implementations run free and checks they stay within tolerance."""
    x, y, yaw = pose
    vx, vy, wz = speed
    dt = pair.cfg.model_dt
    for i in range(cycles):
        inp = nb.CycleInput(pose=(x, y, yaw), speed=(vx, vy, wz), path=path, goal=tuple(path[-1]),
                            goal_checker_xy_tolerance=goal_tol)
        g, c = pair.step(inp)
        assert_cycle_close(g, c, label=f"{label} cycle {i}")
        compare_state(pair, f"{label} cycle {i}", **cmp)
        if resync:
            pair.gpu.setState(pair.cpu.getState())
            pair.gpu.setControlSequence(pair.cpu.getOptimalControlSequence())
        vx, vy, wz = c.cmd
        x += (vx * np.cos(yaw) - vy * np.sin(yaw)) * dt
        y += (vx * np.sin(yaw) + vy * np.cos(yaw)) * dt
        yaw += wz * dt


def test_c1_diffdrive_default_critics_injected_noise():
    """BASELINE.json configs[1]: DiffDrive 2000x56, default critics (yaml weights), 400x400 @0.05 m."""
    cfg = P.nav2_bringup_config(materialize_tensors=1, visualize=1, debug_cells=1)
    cells = build_map("blocks", seed=1)
    pair = Pair(cfg, cells)
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)
    drive(pair, path, (10.0, 10.0, 0.0), (0.2, 0.0, 0.0), cycles=12, label="C1")
    pair.close()


def test_c1_two_iterations_reference_fixture():
    """The reference harness's own fixture: 40x40 @0.1 m, 8x8 block of cost 250, iteration_count 2
    (benchmark/optimizer_benchmark.cpp:49-99)."""
    cells, res, origin, path = S.reference_benchmark_fixture()
    cfg = P.make_config(
        {"batch_size": 2000, "time_steps": 56, "iteration_count": 2, "motion_model": "diff_drive",
         "materialize_tensors": 1, "visualize": 1, "debug_cells": 1},
        P.NAV2_BRINGUP_CRITICS, robot_radius=0.15)
    pair = Pair(cfg, cells, resolution=res, origin=origin)
    drive(pair, path, (2.0, 2.0, 0.785), (0.0, 0.0, 0.0), cycles=6, label="fixture-2iter", exact=False)
    pair.close()


def test_c1_free_running_stays_within_tolerance():
    """No resynchronisation for 30 cycles: ulp-level differences in the softmax sums must not grow."""
    cfg = P.nav2_bringup_config(materialize_tensors=1, visualize=1, debug_cells=1)
    cells = build_map("blocks", seed=2)
    pair = Pair(cfg, cells)
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)
    drive(pair, path, (10.0, 10.0, 0.0), (0.0, 0.0, 0.0), cycles=30, label="free", resync=False, exact=False)
    pair.close()


@pytest.mark.parametrize("model", ["omni", "ackermann"])
@pytest.mark.parametrize("footprint", [False, True])
def test_c3_models_full_critics(model, footprint):
    """BASELINE.json configs[2] at a size the oracle finishes in seconds: Omni / Ackermann, the full
    critic set incl. ObstaclesCritic + inflation, point and footprint collision checking."""
    overrides = {"CostCritic": {"consider_footprint": footprint}, "ObstaclesCritic": {"consider_footprint": footprint},
                 "VelocityDeadbandCritic": {"deadband_velocities": [0.05, 0.05, 0.05]}}
    kw = dict(footprint=SQUARE_FOOTPRINT) if footprint else dict(robot_radius=0.22)
    cfg = P.make_config(
        {"batch_size": 1024, "time_steps": 100, "motion_model": model, "materialize_tensors": 1, "visualize": 1,
         "debug_cells": 1, "controller_frequency": 20.0},
        ALL_CRITICS, overrides, inflation={"cost_scaling_factor": 3.0, "inflation_radius": 0.70},
        validator={"consider_footprint": footprint}, **kw)
    inscribed = cfg.inscribed_radius
    cells = build_map("blocks", seed=3, inscribed=inscribed, n_blocks=40, clear_radius=0.6)
    pair = Pair(cfg, cells)
    path = S.arc_path((10.0, 10.0, 0.3), n_points=120, curvature=-0.2)
    drive(pair, path, (10.0, 10.0, 0.3), (0.1, 0.0, 0.0), cycles=5, label=f"C3-{model}-fp{footprint}",
          cells_obs=True)
    pair.close()


def _lean_pair_check(cfg, cells, path, pose, speed, cycles, label, resolution=0.05):
    """Lean configuration (no materialised tensors / debug dumps): this is what production runs and
    what selects the fused one-cluster-per-robot kernel for batches <= 2048.  With resync every
    cycle starts bit-identical, so costs must agree to 1e-4 relative and controls to 1e-5."""
    pair = Pair(cfg, cells, resolution=resolution)
    x, y, yaw = pose
    vx, vy, wz = speed
    for i in range(cycles):
        goal = tuple(path[-1]) if len(path) else (0.0, 0.0, 0.0)
        inp = nb.CycleInput(pose=(x, y, yaw), speed=(vx, vy, wz), path=path, goal=goal,
                            goal_checker_xy_tolerance=0.25)
        g, c = pair.step(inp)
        assert_cycle_close(g, c, label=f"{label} cycle {i}")
        assert_costs_close(pair.gpu.getTensor(capi.TENSOR_COSTS), pair.cpu.getTensor(capi.TENSOR_COSTS),
                           label=f"{label} cycle {i}")
        gw, cw = pair.gpu.getTensor(capi.TENSOR_SOFTMAX), pair.cpu.getTensor(capi.TENSOR_SOFTMAX)
        np.testing.assert_allclose(gw, cw, rtol=2e-4, atol=1e-7, err_msg=f"{label} softmax weights")
        pair.gpu.setState(pair.cpu.getState())
        pair.gpu.setControlSequence(pair.cpu.getOptimalControlSequence())
        vx, vy, wz = c.cmd
        dt = cfg.model_dt
        x += (vx * np.cos(yaw) - vy * np.sin(yaw)) * dt
        y += (vx * np.sin(yaw) + vy * np.cos(yaw)) * dt
        yaw += wz * dt
    launches = pair.gpu.launchCount()
    pair.close()
    return launches


def test_c1_lean_fused_cluster_kernel():
    """BASELINE.json configs[1] exactly as benchmarked: one launch per optimize() iteration."""
    cfg = P.nav2_bringup_config()
    cells = build_map("blocks", seed=1)
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)
    launches = _lean_pair_check(cfg, cells, path, (10.0, 10.0, 0.0), (0.2, 0.0, 0.0), 15, "C1-lean")
    assert launches <= 1 + 15 + 2, f"expected the fused single-launch path, saw {launches} launches"


@pytest.mark.parametrize("model", ["omni", "ackermann"])
@pytest.mark.parametrize("footprint", [False, True])
@pytest.mark.parametrize("batch", [1000, 2048, 4096])
def test_lean_models_full_critics(model, footprint, batch):
    """Full critic set, lean configuration: fused kernel up to 2048 trajectories, general path above."""
    overrides = {"CostCritic": {"consider_footprint": footprint}, "ObstaclesCritic": {"consider_footprint": footprint},
                 "VelocityDeadbandCritic": {"deadband_velocities": [0.05, 0.05, 0.05]}}
    kw = dict(footprint=SQUARE_FOOTPRINT) if footprint else dict(robot_radius=0.22)
    cfg = P.make_config({"batch_size": batch, "time_steps": 60, "motion_model": model, "controller_frequency": 20.0,
                         "iteration_count": 2},
                        ALL_CRITICS, overrides, inflation={"cost_scaling_factor": 3.0, "inflation_radius": 0.70},
                        validator={"consider_footprint": footprint}, **kw)
    cells = build_map("blocks", seed=3, inscribed=cfg.inscribed_radius, n_blocks=40, clear_radius=0.6)
    path = S.arc_path((10.0, 10.0, 0.3), n_points=120, curvature=-0.2)
    _lean_pair_check(cfg, cells, path, (10.0, 10.0, 0.3), (0.1, 0.0, 0.0), 4, f"lean-{model}-fp{footprint}-B{batch}")


def test_near_goal_and_empty_path_lean():
    """Goal / GoalAngle active (path shorter than their thresholds), then an empty plan (SURVEY.md §9 hazard)."""
    cfg = P.nav2_bringup_config()
    cells = build_map("blocks", seed=4)
    short = S.arc_path((10.0, 10.0, 0.0), n_points=8)            # 0.35 m: Goal + GoalAngle on, path critics off
    _lean_pair_check(cfg, cells, short, (10.0, 10.0, 0.2), (0.1, 0.0, 0.0), 3, "near-goal")
    empty = np.zeros((0, 3))
    _lean_pair_check(cfg, cells, empty, (10.0, 10.0, 0.2), (0.1, 0.0, 0.0), 2, "empty-path")


@pytest.mark.parametrize("model", ["diff_drive", "omni"])
def test_input_delay_and_clamp_raw_controls(model):
    """MotionModel::predict's input-delay shift (motion_models.hpp:160-216, per-axis offsets from
    model_delay_*) and clamp_raw_controls write-back (:131-156), which also changes what the softmax
    update averages (optimizer.cpp:610-640)."""
    cfg = P.make_config(
        {"batch_size": 1500, "time_steps": 56, "motion_model": model, "controller_frequency": 20.0,
         "model_delay_vx": 0.10, "model_delay_vy": 0.15, "model_delay_wz": 0.20, "clamp_raw_controls": True,
         "materialize_tensors": 1, "visualize": 1, "debug_cells": 1},
        P.NAV2_BRINGUP_CRITICS, P.NAV2_BRINGUP_CRITIC_PARAMS, robot_radius=0.22,
        inflation={"cost_scaling_factor": 3.0, "inflation_radius": 0.70})
    cells = build_map("blocks", seed=6)
    pair = Pair(cfg, cells)
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)
    drive(pair, path, (10.0, 10.0, 0.0), (0.2, 0.0, 0.1), cycles=8, label=f"delay-{model}")
    pair.close()


def test_speed_limit_reset_and_reconfigure():
    """Optimizer::setSpeedLimit (optimizer.cpp:691-719), reset(), and a dynamic parameter update that
    changes batch_size / time_steps (ParametersHandler post-callback -> reset, optimizer.cpp:155)."""
    cfg = P.nav2_bringup_config()
    cells = build_map("blocks", seed=7)
    pair = Pair(cfg, cells)
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)
    inp = nb.CycleInput(pose=(10.0, 10.0, 0.0), speed=(0.3, 0.0, 0.0), path=path, goal_checker_xy_tolerance=0.25)
    for _ in range(3):
        g, c = pair.step(inp)
        assert_cycle_close(g, c, label="before limit")
        pair.gpu.setState(pair.cpu.getState()); pair.gpu.setControlSequence(pair.cpu.getOptimalControlSequence())
    for eng in (pair.gpu, pair.cpu):
        eng.setSpeedLimit(40.0, True)                      # 40 % of the maximum speeds
    for _ in range(4):
        g, c = pair.step(inp)
        assert_cycle_close(g, c, label="limited")
        assert np.all(g.control_sequence[0, 1:] <= 0.5 * 0.4 + 1e-6)
        pair.gpu.setState(pair.cpu.getState()); pair.gpu.setControlSequence(pair.cpu.getOptimalControlSequence())
    for eng in (pair.gpu, pair.cpu):
        eng.setSpeedLimit(0.0, False)                      # NO_SPEED_LIMIT
        eng.reset()
    assert np.all(pair.gpu.getOptimalControlSequence() == 0.0)
    # the noise tensors were injected, so they survive reset() on both sides
    g, c = pair.step(inp)
    assert_cycle_close(g, c, label="after reset")
    # dynamic reconfiguration: new shapes, fresh device buffers
    cfg2 = P.nav2_bringup_config(batch_size=768, time_steps=40)
    pair.gpu.setConfig(cfg2)
    cpu2 = type(pair.cpu)(cfg2)
    cpu2.setCostmap(cells, 0.05, 0.0, 0.0)
    nvx, nvy, nwz = S.seeded_noise(768, 40, (cfg2.vx_std, cfg2.vy_std, cfg2.wz_std), seed=9)
    pair.gpu.setNoise(nvx, nvy, nwz)
    cpu2.setNoise(nvx, nvy, nwz)
    g = pair.gpu.evalControl(inp)
    c = cpu2.evalControl(inp)
    assert g.control_sequence.shape == (3, 40)
    assert_cycle_close(g, c, label="reconfigured")
    cpu2.shutdown()
    pair.close()


def test_fallback_retries_and_no_valid_control():
    """fallback() (optimizer.cpp:281-298): every trajectory collides -> reset, retry, NoValidControl after
    retry_attempt_limit; the next call on a free map works again."""
    cfg = P.make_config({"batch_size": 512, "time_steps": 30, "retry_attempt_limit": 2, "controller_frequency": 20.0},
                        ["ConstraintCritic", "CostCritic", "PreferForwardCritic"], robot_radius=0.1)
    lethal = np.full((100, 100), 254, dtype=np.uint8)
    free = np.zeros((100, 100), dtype=np.uint8)
    pair = Pair(cfg, lethal, resolution=0.1)
    path = np.stack([np.linspace(5, 6, 10), np.full(10, 5.0), np.zeros(10)], axis=1)
    inp = nb.CycleInput(pose=(5.0, 5.0, 0.0), speed=(0.0, 0.0, 0.0), path=path)
    g, c = pair.step(inp, raise_on_failure=False)
    assert g.status == c.status == capi.MPPI_ERR_NO_VALID_CONTROL
    assert g.retries == c.retries == 2 and g.fail_flag and c.fail_flag
    with pytest.raises(nb.NoValidControl):
        pair.gpu.evalControl(inp)
    for eng in (pair.gpu, pair.cpu):
        eng.setCostmap(free, 0.1, 0.0, 0.0)
    g, c = pair.step(inp)
    assert g.status == 0 and not g.fail_flag
    assert_cycle_close(g, c, label="recovered")
    pair.close()


def test_validator_soft_reset():
    """OptimalTrajectoryValidator (optimal_trajectory_validator.hpp:117-158): an inscribed cell right in front of
    the robot with no collision critic configured -> SOFT_RESET -> retries -> NoValidControl, same as the oracle."""
    cfg = P.make_config({"batch_size": 256, "time_steps": 30, "retry_attempt_limit": 1, "controller_frequency": 20.0},
                        ["ConstraintCritic", "PreferForwardCritic", "PathFollowCritic"], robot_radius=0.1)
    cells = np.zeros((100, 100), dtype=np.uint8)
    cells[48:53, 50:60] = 253
    pair = Pair(cfg, cells, resolution=0.1)
    path = np.stack([np.linspace(5, 8, 40), np.full(40, 5.0), np.zeros(40)], axis=1)
    inp = nb.CycleInput(pose=(4.9, 5.0, 0.0), speed=(0.3, 0.0, 0.0), path=path)
    g, c = pair.step(inp, raise_on_failure=False)
    assert g.status == c.status
    assert g.validation == c.validation
    assert g.retries == c.retries
    pair.close()


def test_philox_noise_statistics_and_shard_invariance():
    """NoiseGenerator::generateNoisedControls replacement: N(0, sigma^2) per axis (the reference's own check is
    |noise| bounds at sigma = 0.1, noise_generator_test.cpp:97-100), zero vy for non-holonomic models, and the
    same global tensor regardless of how the batch is sharded (counter = global trajectory index)."""
    cfg = P.nav2_bringup_config(batch_size=4096, time_steps=56)
    opt = nb.Optimizer(cfg)
    nvx, nvy, nwz = (opt.getTensor(t) for t in (capi.TENSOR_NOISE_VX, capi.TENSOR_NOISE_VY, capi.TENSOR_NOISE_WZ))
    assert abs(nvx.mean()) < 3e-3 and abs(nvx.std() - 0.2) < 3e-3
    assert abs(nwz.mean()) < 6e-3 and abs(nwz.std() - 0.4) < 6e-3
    assert np.all(nvy == 0.0)
    assert abs(np.corrcoef(nvx.ravel(), nwz.ravel())[0, 1]) < 0.01
    assert np.abs(nvx).max() < 0.2 * 6.5
    halves = []
    for rank in range(2):
        c2 = P.nav2_bringup_config(batch_size=4096, time_steps=56, shard_world=2, shard_rank=rank)
        o2 = nb.Optimizer(c2)
        halves.append(o2.getTensor(capi.TENSOR_NOISE_VX))
        o2.shutdown()
    assert np.array_equal(np.concatenate(halves, axis=1), nvx)
    opt.reset()                                   # reset() draws a fresh tensor (noise_generator.cpp:77-96)
    assert not np.array_equal(opt.getTensor(capi.TENSOR_NOISE_VX), nvx)
    opt.shutdown()


def test_fleet_matches_single_robot_handles():
    """Fleet mode: n_robots independent robots in one handle give exactly what separate handles give."""
    R = 3
    cfgR = P.nav2_bringup_config(n_robots=R, batch_size=1000)
    cfg1 = P.nav2_bringup_config(batch_size=1000)
    fleet = nb.Optimizer(cfgR)
    singles = [nb.Optimizer(cfg1) for _ in range(R)]
    inputs = []
    for r in range(R):
        cells = build_map("blocks", seed=20 + r, size=200, pose=(5.0, 5.0))
        nvx, nvy, nwz = S.seeded_noise(1000, 56, (0.2, 0.2, 0.4), seed=100 + r)
        fleet.setCostmap(cells, 0.05, 0.0, 0.0, robot=r)
        fleet.setNoise(nvx, nvy, nwz, robot=r)
        singles[r].setCostmap(cells, 0.05, 0.0, 0.0)
        singles[r].setNoise(nvx, nvy, nwz)
        path = S.arc_path((5.0, 5.0, 0.1 * r), n_points=80, curvature=0.1 * (r - 1))
        inputs.append(nb.CycleInput(pose=(5.0, 5.0, 0.1 * r), speed=(0.1 * r, 0.0, 0.0), path=path, goal_checker_xy_tolerance=0.25))
    for cycle in range(3):
        res = fleet.evalControl(inputs)
        for r in range(R):
            one = singles[r].evalControl(inputs[r])
            np.testing.assert_array_equal(res[r].control_sequence, one.control_sequence, err_msg=f"robot {r} cycle {cycle}")
            assert res[r].cmd == one.cmd
    fleet.shutdown()
    for s_ in singles:
        s_.shutdown()


def _lean_engine(env=None):
    """A C1 handle (lean configuration -> single-launch cycles) with seeded injected noise; `env` is set only
    while the handle is created (the library reads its switches in mppi_create)."""
    import os
    env = env or {}
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        cfg = P.nav2_bringup_config()
        opt = nb.Optimizer(cfg)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    nvx, nvy, nwz = S.seeded_noise(2000, 56, (cfg.vx_std, cfg.vy_std, cfg.wz_std), seed=5)
    opt.setNoise(nvx, nvy, nwz)
    return opt


def test_in_place_costmap_matches_upload_then_optimize():
    """mppi_optimize_in_place (costmap read where it lies, window-only upload) against mppi_upload_costmap +
    mppi_optimize, bit for bit, over cycles with a moving robot and a changing costmap."""
    two_call, in_place = _lean_engine(), _lean_engine()
    pose = [10.0, 10.0, 0.0]
    speed = (0.2, 0.0, 0.0)
    for cycle in range(8):
        cells = build_map("blocks", seed=30 + cycle // 3)        # the costmap changes every third cycle
        path = S.arc_path(tuple(pose), n_points=100)
        inp = nb.CycleInput(pose=tuple(pose), speed=speed, path=path, goal_checker_xy_tolerance=0.25)
        two_call.setCostmap(cells, 0.05, 0.0, 0.0)
        a = two_call.evalControl(inp)
        b = in_place.evalControl(inp, costmaps=(cells, 0.05, 0.0, 0.0))
        np.testing.assert_array_equal(a.control_sequence, b.control_sequence, err_msg=f"cycle {cycle}")
        np.testing.assert_array_equal(a.optimal_trajectory, b.optimal_trajectory)
        assert a.cmd == b.cmd and a.furthest_reached_path_point == b.furthest_reached_path_point
        pose[0] += 0.05 * a.cmd[0] * np.cos(pose[2]); pose[1] += 0.05 * a.cmd[0] * np.sin(pose[2]); pose[2] += 0.05 * a.cmd[2]
        speed = (a.cmd[0], 0.0, a.cmd[2])
    assert in_place.windowMissCount() == 0
    # the borrowed costmap is gone after the call: the two-call entry point needs its own upload again
    with pytest.raises(Exception):
        in_place.evalControl(inp)
    two_call.shutdown(); in_place.shutdown()


def test_window_miss_repeats_the_cycle_with_the_whole_map():
    """A window too small for the trajectories (forced): the kernel reports the miss, commits nothing, the cycle
    runs again with the whole map, and every result equals a handle with the regular window."""
    regular = _lean_engine()
    tiny = _lean_engine({"MPPI_B200_WINDOW_HALF": "3"})        # 3 cells = 0.15 m either side of the robot
    cells = build_map("blocks", seed=33)
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)
    inp = nb.CycleInput(pose=(10.0, 10.0, 0.0), speed=(0.3, 0.0, 0.0), path=path, goal_checker_xy_tolerance=0.25)
    for cycle in range(4):
        a = regular.evalControl(inp, costmaps=(cells, 0.05, 0.0, 0.0))
        b = tiny.evalControl(inp, costmaps=(cells, 0.05, 0.0, 0.0))
        np.testing.assert_array_equal(a.control_sequence, b.control_sequence, err_msg=f"cycle {cycle}")
        assert a.cmd == b.cmd
    assert regular.windowMissCount() == 0
    misses = tiny.windowMissCount()
    assert misses >= 3        # the first cycle starts from a zero control sequence and can stay inside 0.15 m
    # same through the two-call entry point (staged map, window-only upload, miss, whole map)
    tiny.setCostmap(cells, 0.05, 0.0, 0.0); regular.setCostmap(cells, 0.05, 0.0, 0.0)
    a = regular.evalControl(inp); b = tiny.evalControl(inp)
    np.testing.assert_array_equal(a.control_sequence, b.control_sequence)
    assert tiny.windowMissCount() == misses + 1
    regular.shutdown(); tiny.shutdown()


def test_zero_copy_inputs_match_the_copied_ones():
    """MPPI_B200_ZERO_COPY=1 (opt-in, measured slower): the single-launch kernel pulls the cycle block, the path
    arrays and the packed window rows from the pinned host arena itself instead of a cudaMemcpyAsync ahead of it.
    Same bytes, same kernel: every cycle must end bit-identically, misses included."""
    regular = _lean_engine()
    pulled = _lean_engine({"MPPI_B200_ZERO_COPY": "1"})
    tiny = _lean_engine({"MPPI_B200_ZERO_COPY": "1", "MPPI_B200_WINDOW_HALF": "3"})
    cells = build_map("blocks", seed=35)
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)
    inp = nb.CycleInput(pose=(10.0, 10.0, 0.0), speed=(0.3, 0.0, 0.0), path=path, goal_checker_xy_tolerance=0.25)
    for cycle in range(5):
        a = regular.evalControl(inp, costmaps=(cells, 0.05, 0.0, 0.0))
        for other in (pulled, tiny):
            b = other.evalControl(inp, costmaps=(cells, 0.05, 0.0, 0.0))
            np.testing.assert_array_equal(a.control_sequence, b.control_sequence, err_msg=f"cycle {cycle}")
            np.testing.assert_array_equal(a.optimal_trajectory, b.optimal_trajectory, err_msg=f"cycle {cycle}")
            assert a.cmd == b.cmd
    assert pulled.windowMissCount() == 0 and tiny.windowMissCount() >= 3
    assert pulled.h2dByteCount() == regular.h2dByteCount()      # the pulled bytes are counted like the copied ones
    for e in (regular, pulled, tiny):
        e.shutdown()


def test_visualize_on_the_single_launch_path_matches_the_general_kernels():
    """`visualize` (bring-up default, nav2_params.yaml:157): per-critic cost arrays (critic_manager.cpp:79-117),
    collision flags and the strided trajectory samples TrajectoryVisualizer::add draws
    (trajectory_visualizer.cpp:144,177).  The single-launch kernel writes them itself; the general kernels (forced
    with materialize_tensors) are the cross-check, and both are compared with the oracle's tensors."""
    base = dict(batch_size=2000, time_steps=56, visualize=1, vis_trajectory_step=5, vis_time_step=3)
    cfg_fused = P.nav2_bringup_config(**base)
    cfg_general = P.nav2_bringup_config(materialize_tensors=1, **base)
    cells = build_map("blocks", seed=41)
    cells[180:230, 204:208] = 254                            # a wall 0.2 m ahead: some candidates collide
    nvx, nvy, nwz = S.seeded_noise(2000, 56, (cfg_fused.vx_std, cfg_fused.vy_std, cfg_fused.wz_std), seed=8)
    from oracle.binding import OracleOptimizer
    engines = [nb.Optimizer(cfg_fused), nb.Optimizer(cfg_general)]
    cpu = OracleOptimizer(cfg_general)
    for e in engines + [cpu]:
        e.setCostmap(cells, 0.05, 0.0, 0.0)
        e.setNoise(nvx, nvy, nwz)
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)
    inp = nb.CycleInput(pose=(10.0, 10.0, 0.0), speed=(0.3, 0.0, 0.0), path=path, goal_checker_xy_tolerance=0.25)
    for cycle in range(3):
        res = [e.evalControl(inp) for e in engines]
        ref = cpu.evalControl(inp)
        # the two paths reduce the softmax sums in a different order: last-ulp differences in the update
        np.testing.assert_allclose(res[0].control_sequence, res[1].control_sequence, rtol=0, atol=1e-6)
        f, g = engines
        for t in (capi.TENSOR_COSTS, capi.TENSOR_CRITIC_COSTS, capi.TENSOR_COLLISIONS, capi.TENSOR_VIS_XY):
            np.testing.assert_array_equal(f.getTensor(t), g.getTensor(t), err_msg=f"tensor {t} cycle {cycle}")
        vis = f.getTensor(capi.TENSOR_VIS_XY)
        assert vis.shape == (400, 19, 2)
        x, y = cpu.getTensor(capi.TENSOR_X), cpu.getTensor(capi.TENSOR_Y)        # (T, B)
        np.testing.assert_array_equal(vis[:, :, 0], x[::3, ::5].T)
        np.testing.assert_array_equal(vis[:, :, 1], y[::3, ::5].T)
        coll = f.getTensor(capi.TENSOR_COLLISIONS)
        np.testing.assert_array_equal(coll, cpu.getTensor(capi.TENSOR_COLLISIONS))
        assert 0 < coll.sum() < 2000
        np.testing.assert_allclose(f.getTensor(capi.TENSOR_CRITIC_COSTS), cpu.getTensor(capi.TENSOR_CRITIC_COSTS), rtol=1e-4, atol=1e-5)
        for e in engines:
            e.setState(cpu.getState()); e.setControlSequence(cpu.getOptimalControlSequence())
    assert f.launchCount() < g.launchCount()                 # one launch per cycle against four
    for e in engines + [cpu]:
        e.shutdown()


def test_regenerate_noises_prefetch_matches_inline_generation():
    """regenerate_noises (bring-up default, nav2_params.yaml:159): the next cycle's Philox tensors are drawn on the
    side stream into the second set while the host is between cycles (the reference's noise thread,
    noise_generator.cpp:98-106).  Same tensors, same commands as drawing them at the start of each cycle."""
    import os
    cfg = P.nav2_bringup_config(regenerate_noises=1)
    prefetching = nb.Optimizer(cfg)
    os.environ["MPPI_B200_NO_NOISE_PREFETCH"] = "1"
    try:
        inline = nb.Optimizer(cfg)
    finally:
        os.environ.pop("MPPI_B200_NO_NOISE_PREFETCH")
    cells = build_map("blocks", seed=12)
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)
    inp = nb.CycleInput(pose=(10.0, 10.0, 0.0), speed=(0.2, 0.0, 0.0), path=path, goal_checker_xy_tolerance=0.25)
    previous = None
    for cycle in range(6):
        a = prefetching.evalControl(inp, costmaps=(cells, 0.05, 0.0, 0.0))
        b = inline.evalControl(inp, costmaps=(cells, 0.05, 0.0, 0.0))
        na, nb_ = prefetching.getTensor(capi.TENSOR_NOISE_VX), inline.getTensor(capi.TENSOR_NOISE_VX)
        np.testing.assert_array_equal(na, nb_, err_msg=f"noise of cycle {cycle}")
        np.testing.assert_array_equal(a.control_sequence, b.control_sequence, err_msg=f"cycle {cycle}")
        assert previous is None or not np.array_equal(previous, na)     # a fresh tensor every cycle
        previous = na
        if cycle == 3:
            prefetching.reset(); inline.reset()                          # reset draws again on both
    prefetching.shutdown(); inline.shutdown()


def test_in_place_fleet_and_arena_growth():
    """Three robots in one handle through mppi_optimize_in_place (one window packet per robot, one copy for all),
    against the two-call sequence; then a larger costmap and a longer path arrive mid-run (the pinned / device
    arenas are reallocated, staged maps carried over) and the results still agree."""
    R = 3
    cfg = P.nav2_bringup_config(n_robots=R, batch_size=1000)
    a, b = nb.Optimizer(cfg), nb.Optimizer(cfg)
    noises = [S.seeded_noise(1000, 56, (0.2, 0.2, 0.4), seed=300 + r) for r in range(R)]
    for r in range(R):
        a.setNoise(*noises[r], robot=r); b.setNoise(*noises[r], robot=r)

    def cycle(size, n_points, seed):
        maps, inputs = [], []
        for r in range(R):
            cells = build_map("blocks", seed=seed + r, size=size, pose=(5.0, 5.0))
            maps.append((cells, 0.05, 0.0, 0.0))
            path = S.arc_path((5.0, 5.0, 0.1 * r), n_points=n_points, curvature=0.05 * (r - 1))
            inputs.append(nb.CycleInput(pose=(5.0, 5.0, 0.1 * r), speed=(0.1 * (r + 1), 0.0, 0.0), path=path,
                                        goal_checker_xy_tolerance=0.25))
            a.setCostmap(cells, 0.05, 0.0, 0.0, robot=r)
        ra = a.evalControl(inputs)
        rb = b.evalControl(inputs, costmaps=maps)
        for r in range(R):
            np.testing.assert_array_equal(ra[r].control_sequence, rb[r].control_sequence, err_msg=f"robot {r}")
            assert ra[r].cmd == rb[r].cmd

    for k in range(2):
        cycle(200, 80, 40 + 10 * k)
    cycle(260, 80, 70)        # bigger costmap: the arenas grow
    cycle(260, 300, 80)       # longer path than the cycle block's capacity: they grow again
    cycle(200, 80, 90)        # and smaller inputs again
    assert a.windowMissCount() == 0 and b.windowMissCount() == 0
    a.shutdown(); b.shutdown()


@pytest.mark.parametrize("pose", [(0.3, 0.3, 0.8), (9.9, 0.2, 3.0), (9.85, 9.75, -2.3), (-0.4, 5.0, 0.0), (5.0, 10.6, -1.57)])
def test_window_upload_at_map_edges_and_odd_sizes(pose):
    """The packed costmap window at the borders of a 203 x 197 map (row pitch padded to 208, window clipped, right edge
    rounded past size_x) and with the robot just outside the map: single-launch path through both entry points
    against the oracle."""
    from oracle.binding import OracleOptimizer
    cfg = P.nav2_bringup_config(batch_size=1024, time_steps=40)
    rng = np.random.default_rng(17)
    cells = (rng.random((197, 203)) < 0.02).astype(np.uint8) * 254
    cells[rng.random((197, 203)) < 0.05] = 120
    cx, cy_ = int(np.clip(pose[0] / 0.05, 0, 202)), int(np.clip(pose[1] / 0.05, 0, 196))
    cells[max(0, cy_ - 6):cy_ + 7, max(0, cx - 6):cx + 7] = 0          # free around the robot
    nvx, nvy, nwz = S.seeded_noise(1024, 40, (cfg.vx_std, cfg.vy_std, cfg.wz_std), seed=3)
    gpu_a, gpu_b, cpu = nb.Optimizer(cfg), nb.Optimizer(cfg), OracleOptimizer(cfg)
    for e in (gpu_a, gpu_b, cpu):
        e.setNoise(nvx, nvy, nwz)
    gpu_a.setCostmap(cells, 0.05, 0.0, 0.0); cpu.setCostmap(cells, 0.05, 0.0, 0.0)
    path = S.arc_path(pose, n_points=60, curvature=0.3)
    inp = nb.CycleInput(pose=pose, speed=(0.2, 0.0, 0.1), path=path, goal_checker_xy_tolerance=0.25)
    for cycle in range(3):
        a = gpu_a.evalControl(inp, raise_on_failure=False)
        b = gpu_b.evalControl(inp, raise_on_failure=False, costmaps=(cells, 0.05, 0.0, 0.0))
        c = cpu.evalControl(inp, raise_on_failure=False)
        assert a.status == b.status == c.status and a.retries == b.retries == c.retries
        np.testing.assert_array_equal(a.control_sequence, b.control_sequence)
        if c.status == 0:
            np.testing.assert_allclose(a.control_sequence, c.control_sequence, rtol=0, atol=1e-5, err_msg=f"cycle {cycle}")
            np.testing.assert_allclose(np.array(a.cmd), np.array(c.cmd), rtol=0, atol=1e-5)
        for e in (gpu_a, gpu_b):
            e.setState(cpu.getState()); e.setControlSequence(cpu.getOptimalControlSequence())
    for e in (gpu_a, gpu_b, cpu):
        e.shutdown()


def test_measurement_entry_points():
    """mppi_time_resident (event-timed replays, hot and with a flushed L2), mppi_h2d_byte_count (what bench.py reports
    as h2d_bytes_per_step) and the launch counter."""
    cfg = P.nav2_bringup_config()
    opt = _lean_engine()
    cells = build_map("blocks", seed=2)
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)
    inp = nb.CycleInput(pose=(10.0, 10.0, 0.0), speed=(0.2, 0.0, 0.0), path=path, goal_checker_xy_tolerance=0.25)
    opt.evalControl(inp, costmaps=(cells, 0.05, 0.0, 0.0))
    b0, l0 = opt.h2dByteCount(), opt.launchCount()
    for _ in range(4):
        opt.evalControl(inp, costmaps=(cells, 0.05, 0.0, 0.0))
    per_cycle = (opt.h2dByteCount() - b0) / 4
    assert 4000 < per_cycle < 40000          # cycle block + packed window, not the 160 KB map
    assert opt.launchCount() - l0 == 4       # one kernel per cycle
    opt.setCostmap(cells, 0.05, 0.0, 0.0)
    opt.setResidentCapture(True)
    opt.evalControl(inp)
    hot = opt.timeResident(20, 3, 0)
    cold = opt.timeResident(20, 3, 256 << 20)
    assert hot.shape == (20,) and np.all(hot > 0.005) and np.all(hot < 5.0)
    assert np.median(cold) > np.median(hot)
    opt.shutdown()


@pytest.mark.parametrize("model", ["omni", "diff_drive"])
def test_resident_replay_recomputes_the_captured_cycle(model):
    """mppi_enqueue_resident (what bench.py's `value` times) must run the SAME kernels on the same inputs as the live
    cycle it captured: identical per-trajectory costs, with every critic of the list active (a replay planned from the
    finished cycle's cleared run flags once picked an instantiation without ObstaclesCritic)."""
    cfg = P.make_config({"batch_size": 12000, "time_steps": 60, "motion_model": model, "controller_frequency": 20.0},
                        ALL_CRITICS, {"VelocityDeadbandCritic": {"deadband_velocities": [0.05, 0.05, 0.05]}}, robot_radius=0.22,
                        inflation={"cost_scaling_factor": 3.0, "inflation_radius": 0.70})
    opt = nb.Optimizer(cfg)
    cells = build_map("blocks", seed=3, inscribed=cfg.inscribed_radius, n_blocks=40, clear_radius=0.6)
    opt.setCostmap(cells, 0.05, 0.0, 0.0)
    path = S.arc_path((10.0, 10.0, 0.3), n_points=120, curvature=-0.2)
    inp = nb.CycleInput(pose=(10.0, 10.0, 0.3), speed=(0.2, 0.0, 0.0), path=path, goal_checker_xy_tolerance=0.25)
    for _ in range(6):
        opt.evalControl(inp)
    opt.setResidentCapture(True)
    opt.evalControl(inp)
    live = opt.getTensor(capi.TENSOR_COSTS)
    for _ in range(2):
        opt.enqueueResident(None)
        opt.synchronize()
        replay = opt.getTensor(capi.TENSOR_COSTS)
        assert np.isfinite(replay).all()
        np.testing.assert_array_equal(replay, live)
    opt.shutdown()
