#!/usr/bin/env python
"""Generates tests/golden/*.npz: small seeded scenarios run through the oracle (oracle/, the CPU restatement pinned
against the reference's own test expectations in tests/test_oracle_*.py).  The reference itself cannot be built or
imported here (C++ / ROS 2), so these are oracle outputs, committed so that
  * the oracle cannot drift unnoticed (tests/test_golden.py::test_oracle_reproduces_golden, CPU), and
  * the CUDA path is checked against fixed vectors as well as against the live oracle (…::test_gpu_matches_golden).
Re-run only when a deliberate change of the restatement is made:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from navigation2_b200 import capi, params as P, scenarios as S  # noqa: E402
from oracle.binding import OracleOptimizer  # noqa: E402

SQUARE = [(0.25, 0.2), (0.25, -0.2), (-0.25, -0.2), (-0.25, 0.2)]
ALL_CRITICS = ["ConstraintCritic", "CostCritic", "GoalCritic", "GoalAngleCritic", "ObstaclesCritic", "PathAlignCritic",
               "PathFollowCritic", "PathAngleCritic", "PreferForwardCritic", "TwirlingCritic", "VelocityDeadbandCritic"]


def scenarios():
    """name -> (config, costmap seed, path, pose, speed, cycles)"""
    out = {}
    out["c1_diffdrive_default"] = dict(
        cfg=P.nav2_bringup_config(batch_size=512, time_steps=56, visualize=1, debug_cells=1, materialize_tensors=1),
        map_seed=11, n_blocks=20, clear=1.0, path=S.arc_path((10.0, 10.0, 0.0), n_points=100), pose=(10.0, 10.0, 0.0),
        speed=(0.3, 0.0, 0.05), cycles=3)
    for model, fp in (("omni", False), ("ackermann", True)):
        over = {"CostCritic": {"consider_footprint": fp}, "ObstaclesCritic": {"consider_footprint": fp},
                "VelocityDeadbandCritic": {"deadband_velocities": [0.05, 0.05, 0.05]}}
        kw = dict(footprint=SQUARE) if fp else dict(robot_radius=0.22)
        cfg = P.make_config({"batch_size": 256, "time_steps": 40, "motion_model": model, "materialize_tensors": 1,
                             "visualize": 1, "debug_cells": 1, "controller_frequency": 20.0},
                            ALL_CRITICS, over, inflation={"cost_scaling_factor": 3.0, "inflation_radius": 0.70},
                            validator={"consider_footprint": fp}, **kw)
        out[f"c3_{model}_full_critics"] = dict(
            cfg=cfg, map_seed=5, n_blocks=40, clear=0.6, path=S.arc_path((10.0, 10.0, 0.3), n_points=120, curvature=-0.2),
            pose=(10.0, 10.0, 0.3), speed=(0.1, 0.0, 0.0), cycles=3)
    return out


def run(engine_cls, sc):
    """Free-running cycles from a zero control sequence; returns the per-cycle results and last-cycle tensors."""
    cfg = sc["cfg"]
    cells = S.block_costmap(400, 400, 0.05, seed=sc["map_seed"], keep_clear=sc["pose"][:2],
                            inscribed_radius=cfg.inscribed_radius if sc["n_blocks"] == 40 else None,
                            n_blocks=sc["n_blocks"], clear_radius=sc["clear"])
    holo = cfg.motion_model == capi.MPPI_MODEL_OMNI
    nvx, nvy, nwz = S.seeded_noise(cfg.batch_size, cfg.time_steps, (cfg.vx_std, cfg.vy_std, cfg.wz_std), seed=77, holonomic=holo)
    eng = engine_cls(cfg)
    eng.setCostmap(cells, 0.05, 0.0, 0.0)
    eng.setNoise(nvx, nvy, nwz)
    import navigation2_b200 as nb
    inp = nb.CycleInput(pose=sc["pose"], speed=sc["speed"], path=sc["path"], goal=tuple(sc["path"][-1]),
                        goal_checker_xy_tolerance=0.25)
    res = {}
    for c in range(sc["cycles"]):
        r = eng.evalControl(inp)
        res[f"u_{c}"] = r.control_sequence
        res[f"cmd_{c}"] = np.array(r.cmd, dtype=np.float64)
        res[f"traj_{c}"] = r.optimal_trajectory
        if c == 0:      # the first cycle starts from identical state on every engine: bit-comparable tensors
            res["costs_0"] = eng.getTensor(capi.TENSOR_COSTS)
            res["collisions_0"] = eng.getTensor(capi.TENSOR_COLLISIONS)
            res["cost_cells_0"] = eng.getTensor(capi.TENSOR_COST_CELLS)
            res["x_0"] = eng.getTensor(capi.TENSOR_X)
            res["yaw_0"] = eng.getTensor(capi.TENSOR_YAW)
    eng.shutdown()
    return res


if __name__ == "__main__":
    for name, sc in scenarios().items():
        res = run(OracleOptimizer, sc)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
        print(name, {k: v.shape for k, v in res.items()})
