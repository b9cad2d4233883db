"""Pins the CPU oracle against the reference's own closed-form critic expectations.

Every test is a transcription of a TEST in
/root/reference/nav2_mppi_controller/test/critics_tests.cpp (line numbers cited per test): same
tensors (1000 x 30, model_dt 0.1), same parameters, same expected numbers and tolerances.  The
reference's tests build a CriticData by hand and call one critic's score(); `scoreTrajectories`
is that interface.  The default test costmap is Costmap2DROS("dummy_costmap") after on_configure:
50 x 50 cells @0.1 m, origin (0, 0), all free.

The same scenarios run on the GPU in tests/test_gpu_critics.py (-m gpu).
"""
import math

import numpy as np
import pytest

import navigation2_b200 as nb
from navigation2_b200 import params as P
from oracle.binding import OracleOptimizer

B, T = 1000, 30


def make_engine(critic, overrides=None, model="diff_drive", params=None, engine_cls=OracleOptimizer, **kw):
    p = {"batch_size": B, "time_steps": T, "model_dt": 0.1, "controller_frequency": 10.0, "motion_model": model}
    p.update(params or {})
    cfg = P.make_config(p, [critic], {critic: overrides or {}}, **kw)
    eng = engine_cls(cfg)
    eng.setCostmap(np.zeros((50, 50), dtype=np.uint8), 0.1, 0.0, 0.0)
    return eng


def zeros():
    return np.zeros((T, B), dtype=np.float32)


def path10(**points):
    """models::Path::reset(10): ten zero poses, then individual entries set like the tests do."""
    p = np.zeros((10, 3), dtype=np.float64)
    for k, v in points.items():
        idx = int(k[1:])
        p[idx] = v
    return p


def cycle(path, pose_x=0.0, lpl=None, goal_checker_tol=None, pose_y=0.0, pose_yaw=0.0):
    return nb.CycleInput(pose=(pose_x, pose_y, pose_yaw), speed=(0.0, 0.0, 0.0), path=path,
                         local_path_length=lpl, goal_checker_xy_tolerance=goal_checker_tol)


def score(eng, inp, state=None, traj=None, costs=None, furthest=-1):
    state = state or {}
    traj = traj or {}
    st = {"vx": state.get("vx", zeros()), "vy": state.get("vy", zeros()), "wz": state.get("wz", zeros())}
    tr = {"x": traj.get("x", zeros()), "y": traj.get("y", zeros()), "yaws": traj.get("yaws", zeros())}
    return eng.scoreTrajectories(inp, st, tr, costs=costs, furthest_reached_path_point=furthest)


class CriticKATs:
    """The known-answer tests, parametrised by engine so the GPU suite can reuse them."""
    engine_cls = OracleOptimizer

    def eng(self, *a, **k):
        return make_engine(*a, engine_cls=self.engine_cls, **k)

    # critics_tests.cpp:59-203 ConstraintsCritic
    def test_constraint_critic(self):
        inp = cycle(path10())
        e = self.eng("ConstraintCritic")
        vx = np.full((T, B), 0.40, np.float32)
        wz = np.ones((T, B), np.float32)
        costs, *_ = score(e, inp, {"vx": vx, "wz": wz})
        assert abs(costs.sum()) < 1e-6
        vx[:, 999] = 0.60  # 4.0 weight * 0.1 model_dt * 0.1 error * 30 timesteps = 1.2
        costs, *_ = score(e, inp, {"vx": vx, "wz": wz})
        assert costs.sum() > 0 and abs(costs[999] - 1.2) < 0.01
        vx[:, 1] = -0.45
        costs, *_ = score(e, inp, {"vx": vx, "wz": wz})
        assert abs(costs[1] - 1.2) < 0.01
        e2 = self.eng("ConstraintCritic", {"cost_power": 2})
        costs, *_ = score(e2, inp, {"vx": vx, "wz": wz})
        assert abs(costs[1] - 1.44) < 0.01  # 1.2^2
        # Ackermann: in constraint, then violating min turning radius 0.2
        ea = self.eng("ConstraintCritic", model="ackermann")
        vx = np.full((T, B), 0.40, np.float32)
        wz = np.full((T, B), 1.5, np.float32)
        costs, *_ = score(ea, inp, {"vx": vx, "wz": wz})
        assert abs(costs.sum()) < 1e-6
        wz[:] = 2.5  # 4.0 * 0.1 * (0.2 - 0.4/2.5) * 30 = 0.48
        costs, *_ = score(ea, inp, {"vx": vx, "wz": wz})
        assert abs(costs[1] - 0.48) < 0.01
        ea2 = self.eng("ConstraintCritic", {"cost_power": 2}, model="ackermann")
        costs, *_ = score(ea2, inp, {"vx": vx, "wz": wz})
        assert abs(costs[1] - 0.23) < 0.01
        # Holonomic with mppi.vy_max 0.3
        eo = self.eng("ConstraintCritic", model="omni", params={"vy_max": 0.3})
        vx, vy, wz = zeros(), zeros(), zeros()
        vx[:, 999] = 0.60
        costs, *_ = score(eo, inp, {"vx": vx, "vy": vy, "wz": wz})
        assert abs(costs[999] - 1.2) < 0.01
        vx[:] = 0.0
        vy[:, 999] = 0.50  # 4.0 * 0.1 * 0.2 * 30 = 2.4
        costs, *_ = score(eo, inp, {"vx": vx, "vy": vy, "wz": wz})
        assert abs(costs[999] - 2.4) < 0.01
        vx[:, 999] = 0.6
        vy[:, 999] = -0.5
        costs, *_ = score(eo, inp, {"vx": vx, "vy": vy, "wz": wz})
        assert abs(costs[999] - 3.6) < 0.01
        eo2 = self.eng("ConstraintCritic", {"cost_power": 2}, model="omni", params={"vy_max": 0.3})
        costs, *_ = score(eo2, inp, {"vx": vx, "vy": vy, "wz": wz})
        assert abs(costs[999] - 12.96) < 0.01

    # critics_tests.cpp:285-346 GoalAngleCritic
    def test_goal_angle_critic(self):
        e = self.eng("GoalAngleCritic")
        path = path10(p9=(10.0, 0.0, 3.14))
        costs, *_ = score(e, cycle(path, pose_x=1.0, lpl=abs(1.0 - 10.0)))
        assert abs(costs.sum()) < 1e-6          # far away: no cost
        costs, *_ = score(e, cycle(path, pose_x=9.2, lpl=abs(9.2 - 10.0)))
        assert abs(costs.sum()) < 1e-6          # 0.8 m > 0.5 m threshold
        costs, *_ = score(e, cycle(path, pose_x=9.7, lpl=abs(9.7 - 10.0)))
        assert costs.sum() > 0
        assert abs(costs[0] - 9.42) < 0.02      # (3.14 - 0.0) * 3.0 weight

    # critics_tests.cpp:348-406 GoalAngleCriticSymmetric
    def test_goal_angle_critic_symmetric(self):
        e = self.eng("GoalAngleCritic", {"symmetric_yaw_tolerance": True})
        lpl = abs(9.7 - 10.0)
        costs, *_ = score(e, cycle(path10(p9=(10.0, 0.0, 3.14)), pose_x=9.7, lpl=lpl))
        assert costs.sum() > 0 and abs(costs[0]) < 0.02
        costs, *_ = score(e, cycle(path10(p9=(10.0, 0.0, 0.0)), pose_x=9.7, lpl=lpl))
        assert abs(costs[0]) < 0.02
        costs, *_ = score(e, cycle(path10(p9=(10.0, 0.0, 1.57)), pose_x=9.7, lpl=lpl))
        assert abs(costs[0] - 4.71) < 0.02      # (1.57 - 0.0) * 3.0 weight

    # critics_tests.cpp:408-461 GoalCritic
    def test_goal_critic(self):
        e = self.eng("GoalCritic")
        costs, *_ = score(e, cycle(path10(p9=(10.0, 0.0, 0.0)), pose_x=1.0, lpl=abs(1.0 - 10.0)))
        assert abs(costs[2]) < 1e-6 and abs(costs.sum()) < 1e-6
        costs, *_ = score(e, cycle(path10(p9=(0.5, 0.0, 0.0)), pose_x=1.0, lpl=abs(1.0 - 0.5)))
        assert abs(costs[2] - 2.5) < 1e-6       # sqrt(0.5^2) * 5.0 weight
        assert abs(costs.sum() - 2500.0) < 1e-3

    # critics_tests.cpp:463-589 PathAngleCritic (all three modes)
    def test_path_angle_critic(self):
        for mode, expected in ((0, 3.9947), (1, 2.9167), (2, 2.9167)):
            e = self.eng("PathAngleCritic", {"mode": mode})
            # close to the goal: nothing
            costs, *_ = score(e, cycle(path10(p9=(0.15, 0.0, 0.0)), lpl=0.15))
            assert abs(costs.sum()) < 1e-6
            # furthest_reached_path_point = 2 -> offset point 6; angle 0 < max_angle_to_furthest
            costs, *_ = score(e, cycle(path10(p9=(0.95, 0.0, 0.0), p6=(1.0, 0.0, 0.0)), lpl=0.95), furthest=2)
            assert abs(costs.sum()) < 1e-6
            if mode in (1, 2):
                # pointing straight backwards is fine without a directional preference
                costs, *_ = score(e, cycle(path10(p9=(0.95, 0.0, 0.0), p6=(-1.0, 0.0, 0.0)), lpl=0.95), furthest=2)
                assert abs(costs.sum()) < 1e-6
            costs, *_ = score(e, cycle(path10(p9=(0.95, 0.0, 0.0), p6=(-1.0, 4.0, 0.0)), lpl=0.95), furthest=2)
            assert costs.sum() > 0
            assert abs(costs[0] - expected) < 1e-2  # mode 0: atan2(4,-1) [1.81] * 2.2 weight

    # critics_tests.cpp:591-641 PreferForwardCritic
    def test_prefer_forward_critic(self):
        e = self.eng("PreferForwardCritic")
        costs, *_ = score(e, cycle(path10(p9=(10.0, 0.0, 0.0)), pose_x=1.0, lpl=9.0))
        assert abs(costs.sum()) < 1e-6
        costs, *_ = score(e, cycle(path10(p9=(0.15, 0.0, 0.0)), pose_x=1.0, lpl=0.85), {"vx": np.ones((T, B), np.float32)})
        assert abs(costs.sum()) < 1e-6
        costs, *_ = score(e, cycle(path10(p9=(0.15, 0.0, 0.0)), pose_x=1.0, lpl=0.85),
                          {"vx": np.full((T, B), -1.0, np.float32)})
        assert costs.sum() > 0
        assert abs(costs[0] - 15.0) < 1e-3      # 1.0 * 0.1 model_dt * 5.0 weight * 30 length

    # critics_tests.cpp:643-709 TwirlingCritic
    def test_twirling_critic(self):
        e = self.eng("TwirlingCritic")
        far = cycle(path10(p9=(10.0, 0.0, 0.0)), pose_x=1.0, lpl=9.0, goal_checker_tol=0.25)
        costs, *_ = score(e, far)
        assert abs(costs.sum()) < 1e-6
        near = cycle(path10(p9=(0.15, 0.0, 0.0)), pose_x=1.0, lpl=0.85, goal_checker_tol=0.25)
        costs, *_ = score(e, near)
        assert abs(costs.sum()) < 1e-6
        wz = zeros()
        wz[:, 0] = 10.0
        costs, *_ = score(e, near, {"wz": wz})
        assert abs(costs[0] - 100.0) < 1e-4     # mean(10.0) * 10.0 weight
        rng = np.random.default_rng(0)
        wz[:, 0] = (rng.standard_normal(T) * 0.5).astype(np.float32)
        costs, *_ = score(e, near, {"wz": wz})
        assert abs(costs[0] - 2.581) < 2.0      # mean |N(0, 0.5)| * 10 (reference tolerance 4e-1 on its own stream)
        assert abs(costs[0] - 10.0 * np.abs(wz[:, 0]).mean()) < 1e-4

    # critics_tests.cpp:711-762 PathFollowCritic
    def test_path_follow_critic(self):
        e = self.eng("PathFollowCritic")
        p = np.zeros((6, 3))
        p[5] = (1.8, 0.0, 0.0)
        costs, *_ = score(e, cycle(p, pose_x=2.0, lpl=abs(2.0 - 1.8)))
        assert abs(costs.sum()) < 1e-6
        p[5] = (0.15, 0.0, 0.0)
        costs, *_ = score(e, cycle(p, pose_x=2.0, lpl=abs(2.0 - 0.15)))
        assert abs(costs.sum() - 750.0) < 1e-2  # 0.15 * 5 weight * 1000

    # critics_tests.cpp:764-878 PathAlignCritic
    def test_path_align_critic(self):
        e = self.eng("PathAlignCritic")
        costs, *_ = score(e, cycle(path10(p9=(0.85, 0.0, 0.0)), pose_x=1.0, lpl=0.15))
        assert abs(costs.sum()) < 1e-6          # too close to the goal
        costs, *_ = score(e, cycle(path10(p9=(0.15, 0.0, 0.0)), pose_x=1.0, lpl=0.85))
        assert abs(costs.sum()) < 1e-6          # furthest point < offset_from_furthest
        costs, *_ = score(e, cycle(path10(p9=(0.15, 0.0, 0.0)), pose_x=1.0, lpl=0.85), furthest=21)
        assert abs(costs.sum()) < 1e-6
        # a real 22-point path, every trajectory point at x = 0.66
        p = np.zeros((22, 3))
        for i in range(22):
            p[i, 0] = min(0.1 * i, 0.9)
        x = np.full((T, B), 0.66, np.float32)
        costs, *_ = score(e, cycle(p, pose_x=0.0, lpl=0.9), traj={"x": x}, furthest=21)
        assert abs(costs.sum() - 6600.0) < 1e-2
        # path blocked by a lethal square: critic stands down
        cells = np.zeros((50, 50), dtype=np.uint8)
        cells[11:31, 11:31] = 254
        e.setCostmap(cells, 0.1, 0.0, 0.0)
        p2 = np.zeros((22, 3))
        p2[:, 0] = 1.5
        p2[:, 1] = 1.5
        costs, *_ = score(e, cycle(p2, pose_x=0.0, lpl=1.5), traj={"x": x}, furthest=21)
        assert abs(costs.sum()) < 1e-6

    # critics_tests.cpp:880-930 VelocityDeadbandCritic
    def test_velocity_deadband_critic(self):
        e = self.eng("VelocityDeadbandCritic", {"deadband_velocities": [0.08, 0.08, 0.08]}, model="omni")
        inp = cycle(path10())
        st = {"vx": np.full((T, B), 0.80, np.float32), "vy": np.full((T, B), 0.60, np.float32),
              "wz": np.full((T, B), 0.80, np.float32)}
        costs, *_ = score(e, inp, st)
        assert abs(costs.sum()) < 1e-6
        st = {"vx": np.full((T, B), 0.01, np.float32), "vy": np.full((T, B), 0.02, np.float32),
              "wz": np.full((T, B), 0.021, np.float32)}
        costs, *_ = score(e, inp, st)
        assert abs(costs[1] - 19.845) < 0.01    # 35 * 0.1 * 30 * ((0.08-0.01) + (0.08-0.02) + (0.08-0.021))

    # critics_tests.cpp:206-283: Cost / Obstacles critics throw on consider_footprint with a radius costmap
    def test_collision_critics_configuration_mismatch(self):
        for critic in ("CostCritic", "ObstaclesCritic"):
            with pytest.raises(Exception):
                self.eng(critic, {"consider_footprint": True}, robot_radius=0.1)
            self.eng(critic, {"consider_footprint": True}, footprint=[(0.2, 0.2), (0.2, -0.2), (-0.2, -0.2), (-0.2, 0.2)])


class TestOracleCritics(CriticKATs):
    engine_cls = OracleOptimizer
