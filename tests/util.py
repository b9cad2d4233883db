"""Shared helpers of the test-suite: paired (GPU product, CPU oracle) scenarios."""
from __future__ import annotations

import numpy as np

import navigation2_b200 as nb
from navigation2_b200 import capi, params as P, scenarios as S
from oracle.binding import OracleOptimizer

ALL_CRITICS = ["ConstraintCritic", "CostCritic", "GoalCritic", "GoalAngleCritic", "ObstaclesCritic",
               "PathAlignCritic", "PathFollowCritic", "PathAngleCritic", "PreferForwardCritic",
               "TwirlingCritic", "VelocityDeadbandCritic"]

SQUARE_FOOTPRINT = [(0.15, 0.15), (0.15, -0.15), (-0.15, -0.15), (-0.15, 0.15)]


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def std_of(cfg):
    return (cfg.vx_std, cfg.vy_std, cfg.wz_std)


def build_map(kind: str = "blocks", seed: int = 1, size: int = 400, resolution: float = 0.05,
              pose=(10.0, 10.0), inscribed=None, n_blocks: int = 20, clear_radius: float = 1.0):
    if kind == "blocks":
        return S.block_costmap(size, size, resolution, seed=seed, keep_clear=pose, inscribed_radius=inscribed,
                               n_blocks=n_blocks, clear_radius=clear_radius)
    if kind == "empty":
        return np.zeros((size, size), dtype=np.uint8)
    raise ValueError(kind)


class Pair:
    """The product (GPU) and the oracle (CPU) configured identically and fed identical inputs."""

    def __init__(self, cfg, cells, resolution=0.05, origin=(0.0, 0.0), noise_seed=42, inject=True):
        self.cfg = cfg
        self.gpu = nb.Optimizer(cfg)
        self.cpu = OracleOptimizer(cfg)
        self.gpu.setCostmap(cells, resolution, origin[0], origin[1])
        self.cpu.setCostmap(cells, resolution, origin[0], origin[1])
        if inject:
            holo = cfg.motion_model == capi.MPPI_MODEL_OMNI
            nvx, nvy, nwz = S.seeded_noise(cfg.batch_size, cfg.time_steps, std_of(cfg), seed=noise_seed, holonomic=holo)
            self.gpu.setNoise(nvx, nvy, nwz)
            self.cpu.setNoise(nvx, nvy, nwz)
        else:
            nvx = self.gpu.getTensor(capi.TENSOR_NOISE_VX)
            nvy = self.gpu.getTensor(capi.TENSOR_NOISE_VY)
            nwz = self.gpu.getTensor(capi.TENSOR_NOISE_WZ)
            self.gpu.setNoise(nvx, nvy, nwz)
            self.cpu.setNoise(nvx, nvy, nwz)

    def step(self, inp, raise_on_failure=True):
        g = self.gpu.evalControl(inp, raise_on_failure=raise_on_failure)
        c = self.cpu.evalControl(inp, raise_on_failure=raise_on_failure)
        return g, c

    def close(self):
        self.gpu.shutdown()
        self.cpu.shutdown()


def assert_cycle_close(g, c, ctrl_tol=1e-5, traj_tol=1e-4, label=""):
    assert g.status == c.status, f"{label} status {g.status} vs {c.status}"
    assert g.fail_flag == c.fail_flag, f"{label} fail_flag"
    assert g.validation == c.validation, f"{label} validation"
    assert g.retries == c.retries, f"{label} retries"
    if g.status != 0:
        return
    assert g.furthest_reached_path_point == c.furthest_reached_path_point, \
        f"{label} furthest {g.furthest_reached_path_point} vs {c.furthest_reached_path_point}"
    np.testing.assert_allclose(np.array(g.cmd), np.array(c.cmd), rtol=0, atol=ctrl_tol, err_msg=f"{label} cmd_vel")
    np.testing.assert_allclose(g.control_sequence, c.control_sequence, rtol=0, atol=ctrl_tol,
                               err_msg=f"{label} control sequence")
    np.testing.assert_allclose(g.optimal_trajectory, c.optimal_trajectory, rtol=0, atol=traj_tol,
                               err_msg=f"{label} optimal trajectory")


def assert_costs_close(g_costs, c_costs, rel=1e-4, label=""):
    """per-trajectory costs within 1e-4 relative (BASELINE.json north_star)."""
    g_costs = np.asarray(g_costs, dtype=np.float64)
    c_costs = np.asarray(c_costs, dtype=np.float64)
    denom = np.maximum(np.abs(c_costs), 1e-6)
    err = np.abs(g_costs - c_costs) / denom
    worst = int(np.argmax(err))
    assert err[worst] <= rel, f"{label} cost[{worst}]: gpu {g_costs[worst]!r} cpu {c_costs[worst]!r} rel {err[worst]:.3e}"
