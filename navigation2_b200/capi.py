"""ctypes binding of include/mppi_b200.h (the C ABI of libmppi_b200.so).

This is the Python-side stub a maintainer would write against the header; it carries no logic.
The structures mirror the header field by field; `check_abi()` compares sizes with the library.
"""
from __future__ import annotations

import ctypes as C
import os

MPPI_ABI_VERSION = 1
MPPI_MAX_CRITICS = 16
MPPI_MAX_FOOTPRINT = 32
MPPI_MAX_DELAY_STEPS = 16

# MppiStatus
MPPI_OK = 0
MPPI_ERR_INVALID_ARGUMENT = -1
MPPI_ERR_CUDA = -2
MPPI_ERR_NO_VALID_CONTROL = -3
MPPI_ERR_CONFIG = -4
MPPI_ERR_NOT_READY = -5
MPPI_ERR_UNSUPPORTED = -6

# MppiMotionModel
MPPI_MODEL_DIFF_DRIVE = 0
MPPI_MODEL_OMNI = 1
MPPI_MODEL_ACKERMANN = 2
MOTION_MODELS = {"DiffDrive": 0, "Omni": 1, "Ackermann": 2}

# MppiCriticType (critics.xml order)
CRITIC_NAMES = [
    "ConstraintCritic", "CostCritic", "GoalCritic", "GoalAngleCritic", "ObstaclesCritic",
    "PathAlignCritic", "PathAngleCritic", "PathFollowCritic", "PreferForwardCritic",
    "TwirlingCritic", "VelocityDeadbandCritic",
]
CRITIC_TYPES = {n: i for i, n in enumerate(CRITIC_NAMES)}

# MppiValidation
MPPI_VALIDATION_SUCCESS = 0
MPPI_VALIDATION_SOFT_RESET = 1
MPPI_VALIDATION_FAILURE = 2

# MppiTensor
TENSOR_VX, TENSOR_VY, TENSOR_WZ = 0, 1, 2
TENSOR_CVX, TENSOR_CVY, TENSOR_CWZ = 3, 4, 5
TENSOR_X, TENSOR_Y, TENSOR_YAW = 6, 7, 8
TENSOR_NOISE_VX, TENSOR_NOISE_VY, TENSOR_NOISE_WZ = 9, 10, 11
TENSOR_COSTS = 12
TENSOR_CRITIC_COSTS = 13
TENSOR_COLLISIONS = 14
TENSOR_COST_CELLS = 15
TENSOR_OBSTACLE_CELLS = 16
TENSOR_SOFTMAX = 17
TENSOR_PHASE_CLOCKS = 18
TENSOR_VIS_XY = 19


class MppiCriticParams(C.Structure):
    _fields_ = [
        ("type", C.c_int32),
        ("enabled", C.c_int32),
        ("cost_power", C.c_uint32),
        ("cost_weight", C.c_float),
        ("threshold_to_consider", C.c_float),
        ("consider_footprint", C.c_int32),
        ("critical_cost", C.c_float),
        ("near_collision_cost", C.c_uint32),
        ("collision_cost", C.c_float),
        ("near_goal_distance", C.c_float),
        ("trajectory_point_step", C.c_uint32),
        ("repulsion_weight", C.c_float),
        ("critical_weight", C.c_float),
        ("collision_margin_distance", C.c_float),
        ("symmetric_yaw_tolerance", C.c_int32),
        ("max_path_occupancy_ratio", C.c_float),
        ("offset_from_furthest", C.c_uint32),
        ("use_path_orientations", C.c_int32),
        ("max_angle_to_furthest", C.c_float),
        ("mode", C.c_int32),
        ("deadband_velocities", C.c_float * 3),
    ]


class MppiConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("n_robots", C.c_uint32),
        ("batch_size", C.c_uint32),
        ("time_steps", C.c_uint32),
        ("iteration_count", C.c_uint32),
        ("model_dt", C.c_float),
        ("model_delay_vx", C.c_float),
        ("model_delay_vy", C.c_float),
        ("model_delay_wz", C.c_float),
        ("clamp_raw_controls", C.c_int32),
        ("temperature", C.c_float),
        ("gamma", C.c_float),
        ("vx_max", C.c_float), ("vx_min", C.c_float), ("vy_max", C.c_float), ("wz_max", C.c_float),
        ("ax_max", C.c_float), ("ax_min", C.c_float), ("ay_max", C.c_float), ("ay_min", C.c_float),
        ("az_max", C.c_float),
        ("vx_std", C.c_float), ("vy_std", C.c_float), ("wz_std", C.c_float),
        ("retry_attempt_limit", C.c_uint32),
        ("open_loop", C.c_int32),
        ("sgf_order", C.c_uint32),
        ("controller_frequency", C.c_double),
        ("motion_model", C.c_int32),
        ("min_turning_r", C.c_float),
        ("regenerate_noises", C.c_int32),
        ("noise_seed", C.c_uint64),
        ("n_critics", C.c_uint32),
        ("critics", MppiCriticParams * MPPI_MAX_CRITICS),
        ("visualize", C.c_int32),
        ("materialize_tensors", C.c_int32),
        ("vis_trajectory_step", C.c_uint32),
        ("vis_time_step", C.c_uint32),
        ("validator_enabled", C.c_int32),
        ("collision_lookahead_time", C.c_float),
        ("validator_consider_footprint", C.c_int32),
        ("use_radius", C.c_int32),
        ("n_footprint", C.c_uint32),
        ("footprint_x", C.c_double * MPPI_MAX_FOOTPRINT),
        ("footprint_y", C.c_double * MPPI_MAX_FOOTPRINT),
        ("inscribed_radius", C.c_double),
        ("circumscribed_radius", C.c_double),
        ("has_inflation_layer", C.c_int32),
        ("inflation_cost_scaling_factor", C.c_double),
        ("inflation_radius", C.c_double),
        ("track_unknown", C.c_int32),
        ("shard_rank", C.c_uint32),
        ("shard_world", C.c_uint32),
        ("debug_cells", C.c_int32),
    ]


class MppiCostmapView(C.Structure):
    _fields_ = [
        ("cells", C.POINTER(C.c_uint8)),
        ("size_x", C.c_uint32), ("size_y", C.c_uint32),
        ("resolution", C.c_double), ("origin_x", C.c_double), ("origin_y", C.c_double),
    ]


class MppiCycleIn(C.Structure):
    _fields_ = [
        ("pose_x", C.c_double), ("pose_y", C.c_double), ("pose_yaw", C.c_double),
        ("speed_vx", C.c_double), ("speed_vy", C.c_double), ("speed_wz", C.c_double),
        ("path_x", C.POINTER(C.c_double)),
        ("path_y", C.POINTER(C.c_double)),
        ("path_yaw", C.POINTER(C.c_double)),
        ("n_path", C.c_uint32),
        ("goal_x", C.c_double), ("goal_y", C.c_double), ("goal_yaw", C.c_double),
        ("has_goal_checker", C.c_int32),
        ("goal_checker_xy_tolerance", C.c_double),
        ("has_local_path_length", C.c_int32),
        ("local_path_length", C.c_double),
    ]


class MppiCycleOut(C.Structure):
    _fields_ = [
        ("status", C.c_int32),
        ("cmd_vx", C.c_float), ("cmd_vy", C.c_float), ("cmd_wz", C.c_float),
        ("fail_flag", C.c_int32),
        ("validation", C.c_int32),
        ("retries", C.c_uint32),
        ("furthest_reached_path_point", C.c_int32),
        ("control_vx", C.POINTER(C.c_float)),
        ("control_vy", C.POINTER(C.c_float)),
        ("control_wz", C.POINTER(C.c_float)),
        ("optimal_trajectory", C.POINTER(C.c_float)),
    ]


class MppiOptimizerState(C.Structure):
    _fields_ = [
        ("last_command_vel", C.c_double * 3),
        ("control_history", C.c_float * 12),
        ("cmd_history_vx", C.c_float * MPPI_MAX_DELAY_STEPS),
        ("cmd_history_vy", C.c_float * MPPI_MAX_DELAY_STEPS),
        ("cmd_history_wz", C.c_float * MPPI_MAX_DELAY_STEPS),
        ("fallback_counter", C.c_uint32),
    ]


# every symbol include/mppi_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_F = C.POINTER(C.c_float)
_U8 = C.POINTER(C.c_uint8)
SYMBOLS = {
    "mppi_default_config": (C.c_int, [C.POINTER(MppiConfig)]),
    "mppi_default_critic_params": (C.c_int, [C.c_int32, C.POINTER(MppiCriticParams)]),
    "mppi_critic_type_from_name": (C.c_int, [C.c_char_p]),
    "mppi_critic_name": (C.c_char_p, [C.c_int32]),
    "mppi_status_string": (C.c_char_p, [C.c_int]),
    "mppi_sizeof_config": (C.c_size_t, []),
    "mppi_sizeof_critic_params": (C.c_size_t, []),
    "mppi_sizeof_cycle_in": (C.c_size_t, []),
    "mppi_sizeof_cycle_out": (C.c_size_t, []),
    "mppi_sizeof_optimizer_state": (C.c_size_t, []),
    "mppi_sizeof_costmap_view": (C.c_size_t, []),
    "mppi_inflation_compute_cost": (C.c_uint8, [C.c_double, C.c_double, C.c_double, C.c_double]),
    "mppi_create": (C.c_int, [C.POINTER(MppiConfig), C.POINTER(_P)]),
    "mppi_destroy": (C.c_int, [_P]),
    "mppi_set_config": (C.c_int, [_P, C.POINTER(MppiConfig)]),
    "mppi_reset": (C.c_int, [_P, C.c_uint32, C.c_int]),
    "mppi_set_speed_limit": (C.c_int, [_P, C.c_double, C.c_int]),
    "mppi_last_error": (C.c_char_p, [_P]),
    "mppi_upload_costmap": (C.c_int, [_P, C.c_uint32, _U8, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_double]),
    "mppi_set_noise": (C.c_int, [_P, C.c_uint32, _F, _F, _F]),
    "mppi_generate_noise": (C.c_int, [_P, C.c_uint32]),
    "mppi_optimize": (C.c_int, [_P, C.POINTER(MppiCycleIn), C.POINTER(MppiCycleOut)]),
    "mppi_score_tensors": (C.c_int, [_P, C.c_uint32, C.POINTER(MppiCycleIn), _F, _F, _F, _F, _F, _F, C.c_int32,
                                     _F, _U8, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mppi_get_tensor": (C.c_int, [_P, C.c_uint32, C.c_int32, _P, C.c_size_t]),
    "mppi_get_control_sequence": (C.c_int, [_P, C.c_uint32, _F, _F, _F]),
    "mppi_set_control_sequence": (C.c_int, [_P, C.c_uint32, _F, _F, _F]),
    "mppi_get_state": (C.c_int, [_P, C.c_uint32, C.POINTER(MppiOptimizerState)]),
    "mppi_set_state": (C.c_int, [_P, C.c_uint32, C.POINTER(MppiOptimizerState)]),
    "mppi_enqueue_resident": (C.c_int, [_P, _P]),
    "mppi_time_resident": (C.c_int, [_P, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_float)]),
    "mppi_enqueue_score_resident": (C.c_int, [_P, _P]),
    "mppi_set_resident_capture": (C.c_int, [_P, C.c_int]),
    "mppi_time_e2e": (C.c_int, [_P, _U8, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_double,
                                C.POINTER(MppiCycleIn), C.POINTER(MppiCycleOut), C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "mppi_set_costmap_device": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, C.c_uint32, C.c_size_t, C.c_double, C.c_double, C.c_double, _P]),
    "mppi_get_critics_evaluated": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_uint32)]),
    "mppi_time_e2e_views": (C.c_int, [_P, C.POINTER(MppiCostmapView), C.POINTER(MppiCycleIn), C.POINTER(MppiCycleOut), C.c_int,
                                      C.POINTER(C.c_double)]),
    "mppi_optimize_in_place": (C.c_int, [_P, C.POINTER(MppiCostmapView), C.POINTER(MppiCycleIn), C.POINTER(MppiCycleOut)]),
    "mppi_window_miss_count": (C.c_uint64, [_P]),
    "mppi_h2d_byte_count": (C.c_uint64, [_P]),
    "mppi_launch_count": (C.c_uint64, [_P]),
    "mppi_set_kernel_timing": (C.c_int, [_P, C.c_int]),
    "mppi_kernel_time_ms": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "mppi_kernel_c_time_ms": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "mppi_stream": (_P, [_P]),
    "mppi_set_stream": (C.c_int, [_P, _P]),
    "mppi_device_synchronize": (C.c_int, [_P]),
    "mppi_selftest_exact_math": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "mppi_shard_buffers": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_size_t), C.POINTER(_P), C.POINTER(C.c_size_t), C.POINTER(_P)]),
    "mppi_shard_begin": (C.c_int, [_P, C.POINTER(MppiCycleIn)]),
    "mppi_shard_rollout": (C.c_int, [_P]),
    "mppi_shard_score": (C.c_int, [_P]),
    "mppi_shard_update": (C.c_int, [_P, C.c_int]),
    "mppi_shard_end": (C.c_int, [_P, C.POINTER(MppiCycleOut)]),
    "mppi_shard_ipc_handle": (C.c_int, [_P, C.POINTER(C.c_ubyte)]),
    "mppi_shard_open_peers": (C.c_int, [_P, C.POINTER(C.c_ubyte)]),
    "mppi_shard_cycle": (C.c_int, [_P, C.POINTER(MppiCycleIn), C.POINTER(MppiCycleOut)]),
    "mppi_time_shard_cycles": (C.c_int, [_P, C.POINTER(MppiCycleIn), C.POINTER(MppiCycleOut), C.c_int, C.POINTER(C.c_double)]),
}

_LIB = None


def library_path() -> str:
    """The in-tree library; MPPI_B200_LIBRARY points experiments at another build of the same sources."""
    override = os.environ.get("MPPI_B200_LIBRARY")
    if override:
        return override
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmppi_b200.so")


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libmppi_b200.so and type every symbol.  Raises if the library was not built:
    there is no Python or CPU implementation to fall back to."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or library_path()
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: build it with `python -m navigation2_b200.build` "
            "(the optimiser is CUDA-only; there is no fallback path)")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _LIB = lib
    return lib


CONFIG_SYMBOLS = [
    "mppi_default_config", "mppi_default_critic_params", "mppi_critic_type_from_name", "mppi_critic_name",
    "mppi_status_string", "mppi_sizeof_config", "mppi_sizeof_critic_params", "mppi_sizeof_cycle_in",
    "mppi_sizeof_cycle_out", "mppi_sizeof_optimizer_state", "mppi_sizeof_costmap_view", "mppi_inflation_compute_cost",
]
_CONFIG_LIB = None


def load_config_library() -> C.CDLL:
    """The host-only configuration helpers of the C ABI (csrc/mppi_config.cpp: parameter defaults, critic names,
    status strings, struct sizes), built on their own as libmppi_config.so without any CUDA code.  `params` builds
    MppiConfig structures through it, so a process that only describes a workload — bench.py's reference arm, the
    CPU oracle's tests — never maps the product library."""
    global _CONFIG_LIB
    if _CONFIG_LIB is not None:
        return _CONFIG_LIB
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmppi_config.so")
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is missing: build it with `python -m navigation2_b200.build`")
    lib = C.CDLL(p)
    for name in CONFIG_SYMBOLS:
        res, args = SYMBOLS[name]
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _CONFIG_LIB = lib
    return lib


def check_abi(lib: C.CDLL | None = None) -> None:
    lib = lib or load_library()
    pairs = [
        (lib.mppi_sizeof_config(), C.sizeof(MppiConfig), "MppiConfig"),
        (lib.mppi_sizeof_critic_params(), C.sizeof(MppiCriticParams), "MppiCriticParams"),
        (lib.mppi_sizeof_cycle_in(), C.sizeof(MppiCycleIn), "MppiCycleIn"),
        (lib.mppi_sizeof_cycle_out(), C.sizeof(MppiCycleOut), "MppiCycleOut"),
        (lib.mppi_sizeof_optimizer_state(), C.sizeof(MppiOptimizerState), "MppiOptimizerState"),
        (lib.mppi_sizeof_costmap_view(), C.sizeof(MppiCostmapView), "MppiCostmapView"),
    ]
    for native, here, name in pairs:
        if native != here:
            raise RuntimeError(f"ABI mismatch for {name}: library {native} bytes, binding {here} bytes")
