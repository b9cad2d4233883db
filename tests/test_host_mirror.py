"""The C++ host mirror of the reference's plugin classes (navigation2_b200/host): built against the
in-tree ROS type shim, driven like controller_server drives the reference plugin
(configure -> activate -> computeVelocityCommands -> deactivate -> cleanup; the reference's
controller_state_transition_test.cpp and optimizer_smoke_test.cpp)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "navigation2_b200", "host", "test_host")


def test_host_mirror_configuration_logic():
    """ROS parameter names -> critic parameter blocks, ControllerException on the reference's config errors."""
    out = subprocess.run([BIN, "config"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "config ok" in out.stdout


def test_plugin_manifests_name_the_reference_classes():
    host = os.path.join(ROOT, "navigation2_b200", "host")
    assert "nav2_mppi_controller::MPPIController" in open(os.path.join(host, "mppic.xml")).read()
    critics = open(os.path.join(host, "critics.xml")).read()
    for n in ["ConstraintCritic", "CostCritic", "GoalCritic", "GoalAngleCritic", "ObstaclesCritic", "PathAlignCritic",
              "PathAngleCritic", "PathFollowCritic", "PreferForwardCritic", "TwirlingCritic", "VelocityDeadbandCritic"]:
        assert f"mppi::critics::{n}" in critics
    mm = open(os.path.join(host, "motion_models.xml")).read()
    for n in ["DiffDriveMotionModel", "OmniMotionModel", "AckermannMotionModel"]:
        assert f"mppi::{n}" in mm


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["diff_drive", "omni", "ackermann"])
def test_controller_lifecycle_on_gpu_matches_python_binding(model):
    """The C++ MPPIController and the Python Optimizer sit on the same C ABI: fed the same inputs
    (Philox noise from the same seed) they must publish the same commands."""
    import navigation2_b200 as nb
    from navigation2_b200 import params as P
    cycles = 8
    out = subprocess.run([BIN, "run", model, str(cycles)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "run ok" in out.stdout
    cmds = np.array([[float(v) for v in ln.split()[1:]] for ln in out.stdout.splitlines() if ln.startswith("cmd ")])
    assert cmds.shape == (cycles, 3) and np.isfinite(cmds).all()

    cfg = P.make_config({"batch_size": 2000, "time_steps": 56, "iteration_count": 2, "motion_model": model,
                         "controller_frequency": 20.0}, P.NAV2_BRINGUP_CRITICS,
                        {"PathAlignCritic": {"cost_weight": 14.0}}, robot_radius=0.15)
    opt = nb.Optimizer(cfg)
    cells = np.zeros((40, 40), dtype=np.uint8)
    cells[16:24, 16:24] = 250
    path = np.array([[0.5 + 0.05 * i, 0.5 + 0.02 * i, 0.38] for i in range(50)])
    x, y, yaw, speed = 0.5, 0.5, 0.3, (0.0, 0.0, 0.0)
    for i in range(cycles):
        opt.setCostmap(cells, 0.1, 0.0, 0.0)
        # the C++ side converts yaw -> quaternion -> tf2::getYaw; do the same round trip here
        qz, qw = np.sin(yaw * 0.5), np.cos(yaw * 0.5)
        pyaw = np.arctan2(2.0 * qw * qz, qw * qw - qz * qz)
        pq = np.arctan2(2.0 * np.cos(0.19) * np.sin(0.19), np.cos(0.19) ** 2 - np.sin(0.19) ** 2)
        p = path.copy()
        p[:, 2] = pq
        r = opt.evalControl(nb.CycleInput(pose=(x, y, pyaw), speed=speed, path=p, goal=tuple(p[-1]),
                                          goal_checker_xy_tolerance=0.25))
        np.testing.assert_allclose(np.array(r.cmd), cmds[i], rtol=0, atol=2e-6, err_msg=f"cycle {i}")
        speed = tuple(float(v) for v in cmds[i])
        x += (speed[0] * np.cos(yaw) - speed[1] * np.sin(yaw)) * 0.05
        y += (speed[0] * np.sin(yaw) + speed[1] * np.cos(yaw)) * 0.05
        yaw += speed[2] * 0.05
    opt.shutdown()


@pytest.mark.gpu
def test_inflation_layer_mirror_on_gpu_matches_oracle():
    """navigation2_b200/host/include/nav2_costmap_2d_b200/inflation_layer.hpp driven like the reference's
    inflation_tests.cpp; every printed grid is compared with the CPU oracle."""
    from oracle import inflation_binding as ob
    out = subprocess.run([BIN, "inflate"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "inflate ok" in out.stdout
    lines = out.stdout.splitlines()
    grids = {}
    i = 0
    while i < len(lines):
        if lines[i].startswith("grid "):
            _, name, sx, sy = lines[i].split()
            rows = [[int(v) for v in lines[i + 1 + r].split()] for r in range(int(sy))]
            grids[name] = np.array(rows, dtype=np.uint8)
            assert grids[name].shape == (int(sy), int(sx))
            i += int(sy)
        i += 1
    seed = np.zeros((10, 10), np.uint8); seed[0, 0] = 254
    assert np.array_equal(grids["corner"], ob.update_costs(ob.config(1.0, 2.1, inflation_radius=4.1, cost_scaling_factor=1.0), seed))
    seed = np.zeros((100, 100), np.uint8); seed[50, 50] = 254
    assert np.array_equal(grids["centre"], ob.update_costs(ob.config(1.0, 5.0, inflation_radius=10.5, cost_scaling_factor=1.0), seed))
    seed = np.zeros((120, 200), np.uint8)
    for k in range(40):
        seed[(k * 53) % 120, (k * 37) % 200] = 254
    want = ob.update_costs(ob.config(0.05, 0.22, inflation_radius=0.70, cost_scaling_factor=3.0), seed, 20, 10, 180, 100)
    assert np.array_equal(grids["window"], want)


def test_feasible_path_handler_mirror():
    """nav2_controller::FeasiblePathHandler mirror: the known-answer cases of the reference's
    nav2_controller/plugins/test/path_handler.cpp plus prune / costmap-border cases, no GPU needed."""
    out = subprocess.run([BIN, "pathhandler"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "pathhandler ok" in out.stdout


def test_trajectory_visualizer_mirror_reference_cases():
    """mppi::TrajectoryVisualizer mirror (host half): the four cases of the reference's
    test/trajectory_visualizer_tests.cpp (state transition, optimal trajectory markers, candidate trajectories at the
    default strides, optimal path with the command's stamp) plus collision colouring and the no-subscriber early
    return, no GPU needed."""
    out = subprocess.run([BIN, "visunit"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "visunit ok" in out.stdout


def test_parameters_handler_mirror_reference_cases():
    """mppi::ParametersHandler mirror: the reference's test/parameter_handler_test.cpp (namespaced getters, static
    parameters reject dynamic updates with a reason, pre-callbacks run before and post-callbacks after an accepted
    update, foreign parameters are left alone), no GPU needed."""
    out = subprocess.run([BIN, "params"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "params ok" in out.stdout


def _host_fixture_cfg(**extra):
    from navigation2_b200 import params as P
    p = {"batch_size": 2000, "time_steps": 56, "iteration_count": 1, "motion_model": "diff_drive", "controller_frequency": 20.0}
    p.update(extra)
    return P.make_config(p, P.NAV2_BRINGUP_CRITICS, {"PathAlignCritic": {"cost_weight": 14.0}}, robot_radius=0.15)


def _tf_round_trip(yaw):
    qz, qw = np.sin(yaw * 0.5), np.cos(yaw * 0.5)
    return float(np.arctan2(2.0 * qw * qz, qw * qw - qz * qz))


@pytest.mark.gpu
def test_visualizer_and_critics_stats_match_the_oracle():
    """`visualize: true` through the C++ mirror: the CriticsStats message (critic_manager.cpp:79-120) and the
    TrajectoryVisualizer markers (trajectory_visualizer.cpp:110-187) of the first control cycle against the CPU oracle
    fed the same Philox tensors: per-critic cost sums, the strided candidate trajectories point for point, the cost
    colours, collision colouring."""
    import navigation2_b200 as nb
    from navigation2_b200 import capi
    from oracle.binding import OracleOptimizer
    out = subprocess.run([BIN, "visualize", "1"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "visualize ok" in out.stdout
    stats = [ln.split() for ln in out.stdout.splitlines() if ln.startswith("stat ")]
    trajs = [[float(v) for v in ln.split()[1:]] for ln in out.stdout.splitlines() if ln.startswith("traj ")]
    assert [s[1] for s in stats] == ["ConstraintCritic", "CostCritic", "GoalCritic", "GoalAngleCritic", "PathAlignCritic",
                                     "PathFollowCritic", "PathAngleCritic", "PreferForwardCritic"]

    cfg = _host_fixture_cfg(visualize=1, materialize_tensors=1, vis_trajectory_step=50, vis_time_step=5)
    gpu = nb.Optimizer(cfg)                       # same seed as the C++ handle: the same Philox tensors
    cpu = OracleOptimizer(cfg)
    cpu.setNoise(*[gpu.getTensor(t) for t in (capi.TENSOR_NOISE_VX, capi.TENSOR_NOISE_VY, capi.TENSOR_NOISE_WZ)])
    cells = np.zeros((40, 40), dtype=np.uint8)
    cells[16:24, 16:24] = 250
    cpu.setCostmap(cells, 0.1, 0.0, 0.0)
    path = np.array([[0.5 + 0.05 * i, 0.5 + 0.02 * i, _tf_round_trip(0.38)] for i in range(50)])
    inp = nb.CycleInput(pose=(0.5, 0.5, _tf_round_trip(0.3)), speed=(0.0, 0.0, 0.0), path=path, goal=tuple(path[-1]),
                        goal_checker_xy_tolerance=0.25)
    cpu.evalControl(inp)
    critic_costs = cpu.getTensor(capi.TENSOR_CRITIC_COSTS)
    for k, s in enumerate(stats):
        want = float(critic_costs[k].astype(np.float64).sum())
        assert abs(float(s[3]) - want) <= 1e-4 * max(abs(want), 1.0), (s, want)
        assert int(s[2]) == int(want != 0.0)
    x, y = cpu.getTensor(capi.TENSOR_X), cpu.getTensor(capi.TENSOR_Y)      # (T, B)
    costs = cpu.getTensor(capi.TENSOR_COSTS)
    coll = cpu.getTensor(capi.TENSOR_COLLISIONS).astype(bool)
    ok = costs[~coll] if (~coll).any() else costs
    lo, hi = float(ok.min()), float(ok.max())
    assert len(trajs) == 40
    for k, row in enumerate(trajs):
        b = 50 * k
        pts = np.array(row[5:]).reshape(-1, 2)
        np.testing.assert_array_equal(pts[:, 0].astype(np.float32), x[::5, b])
        np.testing.assert_array_equal(pts[:, 1].astype(np.float32), y[::5, b])
        r, g, bl, a = row[1:5]
        if coll[b]:
            assert (r, g, bl, a) == (1.0, 0.0, 1.0, 0.6)
        else:
            n = (float(costs[b]) - lo) / (hi - lo) if hi > lo else 0.0
            want = (2 * n, 1.0) if n < 0.5 else (1.0, 2 * (1 - n))
            assert abs(r - want[0]) < 2e-3 and abs(g - want[1]) < 2e-3 and bl == 0.0 and abs(a - 0.8) < 1e-6
    gpu.shutdown(); cpu.shutdown()


def _prune_like_the_path_handler(path, pose_xy, prune_distance=2.0, map_size=4.0):
    """FeasiblePathHandler::findPlanSegment + transformLocalPlan on an identity transform (feasible_path_handler.cpp:159-252)."""
    d = np.hypot(path[:, 0] - pose_xy[0], path[:, 1] - pose_xy[1])
    closest = 0
    for i in range(1, len(path)):          # min_by: `<=`
        if d[i] <= d[closest]:
            closest = i
    if len(path) > 1 and closest == len(path) - 1:
        closest -= 1
    end, dist = len(path), 0.0
    for i in range(closest, len(path) - 1):
        dist += float(np.hypot(*(path[i + 1, :2] - path[i, :2])))
        if dist > prune_distance:
            end = i + 1
            break
    local = []
    for i in range(closest, end):
        if not (0.0 <= path[i, 0] and 0.0 <= path[i, 1] and int(path[i, 0] / 0.1) < int(map_size / 0.1) and int(path[i, 1] / 0.1) < int(map_size / 0.1)):
            break
        local.append(path[i])
    return closest, np.array(local)


@pytest.mark.gpu
def test_fleet_controller_server_matches_the_python_fleet():
    """FleetControllerServer (host/include/nav2_controller_b200): 5 robots with their own costmaps, plans and poses,
    pruned / transformed by the FeasiblePathHandler mirror and optimised in ONE call per cycle, against the Python
    fleet Optimizer fed the same pruned plans."""
    import navigation2_b200 as nb
    R, cycles = 5, 6
    out = subprocess.run([BIN, "fleet", str(R), str(cycles)], capture_output=True, text=True, timeout=180)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "fleet ok" in out.stdout
    rows = [ln.split() for ln in out.stdout.splitlines() if ln.startswith("fleet ") and len(ln.split()) == 7]
    assert len(rows) == R * cycles
    cfg = _host_fixture_cfg(n_robots=R)
    opt = nb.Optimizer(cfg)
    maps, paths, poses, yaws = [], [], [], []
    for r in range(R):
        cells = np.zeros((40, 40), dtype=np.uint8)
        cells[16:24, 16:24] = 250
        for i in range(6):
            for j in range(6):
                cells[(11 * r + 5 * j) % 40, (7 * r + 3 * i) % 40] = 120        # setCost(mx, my): cells[my, mx]
        maps.append((cells, 0.1, 0.0, 0.0))
        yaw_p = _tf_round_trip(float(np.arctan2(0.01 * (r + 1), 0.05)))
        paths.append(np.array([[0.4 + 0.05 * i, 0.4 + 0.01 * (r + 1) * i, yaw_p] for i in range(60)]))
        poses.append([0.5, 0.5 + 0.02 * r])
        yaws.append(0.1 * r)
    speeds = [(0.0, 0.0, 0.0)] * R
    k = 0
    for c in range(cycles):
        inputs = []
        for r in range(R):
            first, local = _prune_like_the_path_handler(paths[r], poses[r])
            paths[r] = paths[r][first:]                                           # prunePlan
            assert len(local) == int(rows[k + r][6]), f"local plan length, robot {r} cycle {c}"
            inputs.append(nb.CycleInput(pose=(poses[r][0], poses[r][1], _tf_round_trip(yaws[r])), speed=speeds[r], path=local,
                                        goal=tuple(paths[r][-1]), goal_checker_xy_tolerance=0.25))
        res = opt.evalControl(inputs, costmaps=maps)
        for r in range(R):
            got = np.array([float(v) for v in rows[k + r][3:6]])
            assert rows[k + r][2] == "1"
            np.testing.assert_allclose(np.array(res[r].cmd), got, rtol=0, atol=2e-6, err_msg=f"robot {r} cycle {c}")
            speeds[r] = tuple(float(v) for v in got)
            poses[r][0] += speeds[r][0] * np.cos(yaws[r]) * 0.05
            poses[r][1] += speeds[r][0] * np.sin(yaws[r]) * 0.05
            yaws[r] += speeds[r][2] * 0.05
        k += R
    opt.shutdown()


@pytest.mark.gpu
def test_critic_function_score_bridge():
    """CriticFunction::score(CriticData&) through the C++ mirror (critic_function.hpp:95; the call shape of
    test/critics_tests.cpp): closed-form expectations for CostCritic and PreferForwardCritic on caller tensors."""
    out = subprocess.run([BIN, "score"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "score ok" in out.stdout
