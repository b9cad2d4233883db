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
