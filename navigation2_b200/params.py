"""ROS-parameter-shaped configuration of the optimiser.

Parameter names are the reference's ROS parameter names (nav2_mppi_controller
src/optimizer.cpp:105-159 and each critic's initialize()).  `make_config` starts from the code
defaults the C library holds and applies a flat `{name: value}` dict plus per-critic dicts, the
way ParametersHandler::getParamGetter(ns) resolves "<controller>.<Critic>.<param>".
"""
from __future__ import annotations

import math
from typing import Any, Iterable, Mapping

from . import capi

_CONTROLLER_FIELDS = {
    "batch_size", "time_steps", "iteration_count", "model_dt", "model_delay_vx", "model_delay_vy",
    "model_delay_wz", "clamp_raw_controls", "temperature", "gamma", "vx_max", "vx_min", "vy_max",
    "wz_max", "ax_max", "ax_min", "ay_max", "ay_min", "az_max", "vx_std", "vy_std", "wz_std",
    "retry_attempt_limit", "open_loop", "sgf_order", "controller_frequency", "regenerate_noises",
    "visualize", "noise_seed", "materialize_tensors", "vis_trajectory_step", "vis_time_step", "debug_cells", "n_robots", "shard_rank",
    "shard_world", "min_turning_r",
}
_MOTION_MODEL_NAMES = {  # "motion_model" parameter values (optimizer.cpp:150, motion_models.xml)
    "diff_drive": capi.MPPI_MODEL_DIFF_DRIVE, "DiffDrive": capi.MPPI_MODEL_DIFF_DRIVE,
    "omni": capi.MPPI_MODEL_OMNI, "Omni": capi.MPPI_MODEL_OMNI,
    "ackermann": capi.MPPI_MODEL_ACKERMANN, "Ackermann": capi.MPPI_MODEL_ACKERMANN,
}

# nav2_bringup/params/nav2_params.yaml:133-237 — the "default critics with yaml weights" of
# BASELINE.json configs[0]/[1].  visualize / regenerate_noises are left at their code defaults
# (false) as the benchmark harness does (benchmark/optimizer_benchmark.cpp).
NAV2_BRINGUP_CONTROLLER = {
    "time_steps": 56, "model_dt": 0.05, "batch_size": 2000, "ax_max": 3.0, "ax_min": -3.0,
    "ay_max": 3.0, "ay_min": -3.0, "az_max": 3.5, "vx_std": 0.2, "vy_std": 0.2, "wz_std": 0.4,
    "vx_max": 0.5, "vx_min": -0.35, "vy_max": 0.5, "wz_max": 1.9, "iteration_count": 1,
    "temperature": 0.3, "gamma": 0.015, "motion_model": "diff_drive", "controller_frequency": 20.0,
}
NAV2_BRINGUP_CRITICS = [
    "ConstraintCritic", "CostCritic", "GoalCritic", "GoalAngleCritic", "PathAlignCritic",
    "PathFollowCritic", "PathAngleCritic", "PreferForwardCritic",
]
NAV2_BRINGUP_CRITIC_PARAMS = {
    "ConstraintCritic": {"cost_power": 1, "cost_weight": 4.0},
    "GoalCritic": {"cost_power": 1, "cost_weight": 5.0, "threshold_to_consider": 1.4},
    "GoalAngleCritic": {"cost_power": 1, "cost_weight": 3.0, "threshold_to_consider": 0.5},
    "PreferForwardCritic": {"cost_power": 1, "cost_weight": 5.0, "threshold_to_consider": 0.5},
    "CostCritic": {"cost_power": 1, "cost_weight": 3.81, "near_collision_cost": 253, "critical_cost": 300.0,
                   "consider_footprint": False, "collision_cost": 1000000.0, "near_goal_distance": 1.0,
                   "trajectory_point_step": 2},
    "PathAlignCritic": {"cost_power": 1, "cost_weight": 14.0, "max_path_occupancy_ratio": 0.05,
                        "trajectory_point_step": 4, "threshold_to_consider": 0.5, "offset_from_furthest": 20,
                        "use_path_orientations": False},
    "PathFollowCritic": {"cost_power": 1, "cost_weight": 5.0, "offset_from_furthest": 5,
                         "threshold_to_consider": 1.4},
    "PathAngleCritic": {"cost_power": 1, "cost_weight": 2.0, "offset_from_furthest": 4,
                        "threshold_to_consider": 0.5, "max_angle_to_furthest": 1.0, "mode": 0},
}
# nav2_bringup/params/nav2_params.yaml:250,261-262 local costmap
NAV2_BRINGUP_COSTMAP = {"robot_radius": 0.22, "cost_scaling_factor": 3.0, "inflation_radius": 0.70,
                        "resolution": 0.05}


def critic_params(name: str, overrides: Mapping[str, Any] | None = None) -> capi.MppiCriticParams:
    """Code defaults of mppi::critics::<name> with `<name>.<param>` overrides applied."""
    lib = capi.load_config_library()
    short = name.split("::")[-1]
    ctype = lib.mppi_critic_type_from_name(short.encode())
    if ctype < 0:
        raise ValueError(f"unknown critic '{name}' (critics.xml lists: {capi.CRITIC_NAMES})")
    p = capi.MppiCriticParams()
    rc = lib.mppi_default_critic_params(ctype, p)
    if rc != capi.MPPI_OK:
        raise RuntimeError(lib.mppi_status_string(rc).decode())
    fields = {f for f, _ in capi.MppiCriticParams._fields_}
    for k, v in (overrides or {}).items():
        if k == "deadband_velocities":
            for i in range(3):
                p.deadband_velocities[i] = float(v[i])
        elif k in fields and k != "type":
            setattr(p, k, type(getattr(p, k))(v))
        else:
            raise ValueError(f"{short} has no parameter '{k}'")
    return p


def polygon_radii(footprint: Iterable[tuple[float, float]]) -> tuple[float, float]:
    """nav2_costmap_2d::calculateMinAndMaxDistances (nav2_costmap_2d/src/footprint.cpp:43-72)."""
    pts = list(footprint)
    if len(pts) <= 2:
        return (float("inf"), 0.0)

    def dist_to_line(px, py, x0, y0, x1, y1):
        a, b, c, d = px - x0, py - y0, x1 - x0, y1 - y0
        dot, len_sq = a * c + b * d, c * c + d * d
        param = dot / len_sq if len_sq != 0 else -1.0
        if param < 0:
            xx, yy = x0, y0
        elif param > 1:
            xx, yy = x1, y1
        else:
            xx, yy = x0 + param * c, y0 + param * d
        return math.hypot(px - xx, py - yy)

    mn, mx = float("inf"), 0.0
    for i in range(len(pts)):
        x0, y0 = pts[i]
        x1, y1 = pts[(i + 1) % len(pts)]
        vd = math.hypot(x0, y0)
        ed = dist_to_line(0.0, 0.0, x0, y0, x1, y1)
        mn = min(mn, vd, ed)
        mx = max(mx, vd, ed)
    return (mn, mx)


def radius_footprint(radius: float) -> list[tuple[float, float]]:
    """nav2_costmap_2d::makeFootprintFromRadius: a 16-gon (nav2_costmap_2d/src/footprint.cpp:158-174)."""
    return [(math.cos(i * 2 * math.pi / 16) * radius, math.sin(i * 2 * math.pi / 16) * radius) for i in range(16)]


def make_config(
    params: Mapping[str, Any] | None = None,
    critics: Iterable[str] = (),
    critic_overrides: Mapping[str, Mapping[str, Any]] | None = None,
    *,
    robot_radius: float | None = None,
    footprint: Iterable[tuple[float, float]] | None = None,
    inflation: Mapping[str, float] | None = None,
    track_unknown: bool = False,
    validator: Mapping[str, Any] | None = None,
) -> capi.MppiConfig:
    """Build an MppiConfig.

    params            controller-level ROS parameters ("batch_size", "motion_model", ...)
    critics           the "critics" list, evaluation order
    critic_overrides  {"CostCritic": {"cost_weight": ...}, ...}
    robot_radius / footprint   what Costmap2DROS was configured with (exactly one may be given)
    inflation         {"cost_scaling_factor", "inflation_radius"} of the costmap's inflation layer, or None
    validator         TrajectoryValidator.{enabled, collision_lookahead_time, consider_footprint}
    """
    lib = capi.load_config_library()
    cfg = capi.MppiConfig()
    rc = lib.mppi_default_config(cfg)
    if rc != capi.MPPI_OK:
        raise RuntimeError(lib.mppi_status_string(rc).decode())
    for k, v in (params or {}).items():
        if k == "motion_model":
            if v not in _MOTION_MODEL_NAMES:
                raise ValueError(f"Failed to load motion model plugin '{v}'")
            cfg.motion_model = _MOTION_MODEL_NAMES[v]
        elif k in ("AckermannConstraints.min_turning_r", "ackermann.min_turning_r"):
            cfg.min_turning_r = float(v)
        elif k in _CONTROLLER_FIELDS:
            setattr(cfg, k, type(getattr(cfg, k))(v))
        else:
            raise ValueError(f"unknown controller parameter '{k}'")
    names = list(critics)
    if len(names) > capi.MPPI_MAX_CRITICS:
        raise ValueError("too many critics")
    cfg.n_critics = len(names)
    for i, n in enumerate(names):
        cfg.critics[i] = critic_params(n, (critic_overrides or {}).get(n.split("::")[-1]))
    if footprint is not None and robot_radius is not None:
        raise ValueError("give robot_radius or footprint, not both")
    if footprint is not None:
        pts = list(footprint)
        cfg.use_radius = 0
        cfg.n_footprint = len(pts)
        for i, (x, y) in enumerate(pts):
            cfg.footprint_x[i], cfg.footprint_y[i] = x, y
        cfg.inscribed_radius, cfg.circumscribed_radius = polygon_radii(pts)
    else:
        r = 0.1 if robot_radius is None else float(robot_radius)
        pts = radius_footprint(r)
        cfg.use_radius = 1
        cfg.n_footprint = len(pts)
        for i, (x, y) in enumerate(pts):
            cfg.footprint_x[i], cfg.footprint_y[i] = x, y
        # LayeredCostmap::setFootprint computes the radii from the 16-gon, not from the radius itself
        cfg.inscribed_radius, cfg.circumscribed_radius = polygon_radii(pts)
    if inflation:
        cfg.has_inflation_layer = 1
        cfg.inflation_cost_scaling_factor = float(inflation["cost_scaling_factor"])
        cfg.inflation_radius = float(inflation["inflation_radius"])
    cfg.track_unknown = 1 if track_unknown else 0
    for k, v in (validator or {}).items():
        if k == "enabled":
            cfg.validator_enabled = 1 if v else 0
        elif k == "collision_lookahead_time":
            cfg.collision_lookahead_time = float(v)
        elif k == "consider_footprint":
            cfg.validator_consider_footprint = 1 if v else 0
        else:
            raise ValueError(f"unknown TrajectoryValidator parameter '{k}'")
    return cfg


def nav2_bringup_config(**overrides: Any) -> capi.MppiConfig:
    """BASELINE.json configs[0]/[1]: DiffDrive 2000x56, model_dt 0.05, default critics, yaml weights."""
    p = dict(NAV2_BRINGUP_CONTROLLER)
    p.update(overrides)
    return make_config(
        p, NAV2_BRINGUP_CRITICS, NAV2_BRINGUP_CRITIC_PARAMS,
        robot_radius=NAV2_BRINGUP_COSTMAP["robot_radius"],
        inflation={"cost_scaling_factor": NAV2_BRINGUP_COSTMAP["cost_scaling_factor"],
                   "inflation_radius": NAV2_BRINGUP_COSTMAP["inflation_radius"]})
