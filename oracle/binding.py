"""ctypes binding of the CPU oracle (oracle/mppi_oracle.cpp).  TEST INFRASTRUCTURE.

Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs.  It reuses the POD structure definitions of navigation2_b200.capi (the oracle consumes the
same MppiConfig / MppiCycleIn / MppiCycleOut as the product); the product never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from navigation2_b200 import capi
from navigation2_b200.optimizer import CycleInput, CycleResult

HERE = os.path.dirname(os.path.abspath(__file__))
_P = C.c_void_p
_F = C.POINTER(C.c_float)
_U8 = C.POINTER(C.c_uint8)
_I = C.POINTER(C.c_int)

_SYMS = {
    "oracle_create": (C.c_int, [C.POINTER(capi.MppiConfig), C.POINTER(_P)]),
    "oracle_destroy": (C.c_int, [_P]),
    "oracle_reset": (C.c_int, [_P, C.c_int]),
    "oracle_set_speed_limit": (C.c_int, [_P, C.c_double, C.c_int]),
    "oracle_upload_costmap": (C.c_int, [_P, _U8, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_double]),
    "oracle_set_noise": (C.c_int, [_P, _F, _F, _F]),
    "oracle_optimize": (C.c_int, [_P, C.POINTER(capi.MppiCycleIn), C.POINTER(capi.MppiCycleOut)]),
    "oracle_time_optimize": (C.c_int, [_P, C.POINTER(capi.MppiCycleIn), C.c_int, C.POINTER(C.c_double)]),
    "oracle_score_tensors": (C.c_int, [_P, C.POINTER(capi.MppiCycleIn), _F, _F, _F, _F, _F, _F, C.c_int32, _F, _U8,
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "oracle_get_tensor": (C.c_int, [_P, C.c_int32, _P, C.c_size_t]),
    "oracle_get_control_sequence": (C.c_int, [_P, _F, _F, _F]),
    "oracle_set_control_sequence": (C.c_int, [_P, _F, _F, _F]),
    "oracle_last_error": (C.c_char_p, [_P]),
    "oracle_get_state": (C.c_int, [_P, C.POINTER(capi.MppiOptimizerState)]),
    "oracle_set_state": (C.c_int, [_P, C.POINTER(capi.MppiOptimizerState)]),
    "oracle_predict": (C.c_int, [_P, _F, _F, _F, _F, _F, _F]),
    "oracle_push_command_history": (C.c_int, [_P, C.c_float, C.c_float, C.c_float]),
    "oracle_integrate": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, _F, _F, _F, _F, _F, _F]),
    "oracle_apply_constraints": (C.c_int, [_P, C.c_double, C.c_double, C.c_double, C.c_int]),
    "oracle_shift_control_sequence": (C.c_int, [_P]),
    "oracle_savitsky_golay": (C.c_int, [_P, _F]),
    "oracle_get_settings": (C.c_int, [_P, _F]),
    "oracle_pose_point_angle": (C.c_float, [C.c_double] * 5 + [C.c_int]),
    "oracle_pose_point_angle_yaw": (C.c_float, [C.c_double] * 6),
    "oracle_normalize_angle": (C.c_float, [C.c_float]),
    "oracle_shortest_angular_distance": (C.c_float, [C.c_float, C.c_float]),
    "oracle_angles_normalize_angle": (C.c_double, [C.c_double]),
    "oracle_sincosf": (None, [C.c_float, _F, _F]),
    "oracle_find_closest_path_pt": (C.c_uint, [_F, C.c_uint, C.c_float, C.c_uint]),
    "oracle_line_cells": (C.c_int, [C.c_int] * 4 + [_I, _I, C.c_int]),
    "oracle_footprint_cost_at_pose": (C.c_double, [_U8, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_double,
                                                   C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double),
                                                   C.POINTER(C.c_double), C.c_uint32]),
    "oracle_line_cost": (C.c_double, [_U8, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int]),
    "oracle_compute_cost": (C.c_uint8, [C.c_double] * 4),
    "oracle_find_path_costs": (C.c_int, [_U8, C.c_uint32, C.c_uint32, C.c_double, C.c_double, C.c_double, _F, _F, C.c_int,
                                        C.c_int, _U8]),
    "oracle_world_to_map": (C.c_int, [C.c_double] * 5 + [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
}

_LIBS: dict[str, C.CDLL] = {}


def load(fast: bool = False) -> C.CDLL:
    """fast=False: parity build (-ffp-contract=off); fast=True: timing build (-O3 -mavx2 -mfma)."""
    name = "libmppi_oracle_fast.so" if fast else "libmppi_oracle.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(HERE, "_build", name)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
    lib = C.CDLL(path)
    for sym, (res, args) in _SYMS.items():
        fn = getattr(lib, sym)
        fn.restype = res
        fn.argtypes = args
    _LIBS[name] = lib
    return lib


def load_ref_line_iterator():
    """The reference's own nav2_util::LineIterator, compiled from /root/reference (oracle/_ref)."""
    path = os.path.join(HERE, "_ref", "libref_line_iterator.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_line_cells.restype = C.c_int
    lib.ref_line_cells.argtypes = [C.c_int] * 4 + [_I, _I, C.c_int]
    return lib


def _arr(a, shape):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.shape == shape, (a.shape, shape)
    return a


def _p(a):
    return a.ctypes.data_as(_F) if a is not None else None


class OracleOptimizer:
    """Same surface as navigation2_b200.Optimizer (one robot), backed by the CPU restatement."""

    def __init__(self, cfg: capi.MppiConfig, fast: bool = False):
        self._lib = load(fast)
        self.cfg = cfg
        self.T = int(cfg.time_steps)
        self.B = int(cfg.batch_size)
        self.n_critics = int(cfg.n_critics)
        self._h = _P()
        rc = self._lib.oracle_create(C.byref(cfg), C.byref(self._h))
        if rc != capi.MPPI_OK:
            self._h = None
            raise RuntimeError(f"oracle_create failed: {rc}")

    def shutdown(self):
        if getattr(self, "_h", None):
            self._lib.oracle_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.shutdown()
        except Exception:
            pass

    def reset(self, reset_dynamic_speed_limits=True):
        self._lib.oracle_reset(self._h, 1 if reset_dynamic_speed_limits else 0)

    def setSpeedLimit(self, speed_limit, percentage):
        self._lib.oracle_set_speed_limit(self._h, float(speed_limit), 1 if percentage else 0)

    def setCostmap(self, cells, resolution, origin_x, origin_y):
        cells = np.ascontiguousarray(cells, dtype=np.uint8)
        sy, sx = cells.shape
        self._lib.oracle_upload_costmap(self._h, cells.ctypes.data_as(_U8), sx, sy, float(resolution),
                                        float(origin_x), float(origin_y))

    def setNoise(self, nvx, nvy, nwz):
        s = (self.T, self.B)
        a, b, c = _arr(nvx, s), _arr(nvy, s), _arr(nwz, s)
        self._lib.oracle_set_noise(self._h, _p(a), _p(b), _p(c))

    def evalControl(self, inp: CycleInput, raise_on_failure=True) -> CycleResult:
        c_in = inp.to_c()
        c_out = capi.MppiCycleOut()
        T = self.T
        u = np.zeros((3, T), dtype=np.float32)
        traj = np.zeros((3, T), dtype=np.float32)
        c_out.control_vx, c_out.control_vy, c_out.control_wz = _p(u[0]), _p(u[1]), _p(u[2])
        c_out.optimal_trajectory = _p(traj)
        rc = self._lib.oracle_optimize(self._h, C.byref(c_in), C.byref(c_out))
        if rc != capi.MPPI_OK and raise_on_failure:
            raise RuntimeError(f"oracle_optimize: status {rc} {self._lib.oracle_last_error(self._h).decode()}")
        return CycleResult(status=c_out.status if rc == capi.MPPI_OK else rc,
                           cmd=(c_out.cmd_vx, c_out.cmd_vy, c_out.cmd_wz), control_sequence=u,
                           optimal_trajectory=traj.T.copy(), fail_flag=bool(c_out.fail_flag),
                           validation=c_out.validation, retries=c_out.retries,
                           furthest_reached_path_point=c_out.furthest_reached_path_point)

    def timeOptimize(self, inp: CycleInput, cycles: int) -> float:
        c_in = inp.to_c()
        sec = C.c_double(0.0)
        rc = self._lib.oracle_time_optimize(self._h, C.byref(c_in), int(cycles), C.byref(sec))
        if rc != capi.MPPI_OK:
            raise RuntimeError(f"oracle_time_optimize: status {rc}")
        return sec.value

    def scoreTrajectories(self, inp: CycleInput, state, trajectories, costs=None, furthest_reached_path_point=-1):
        s = (self.T, self.B)
        vx, vy, wz = _arr(state.get("vx"), s), _arr(state.get("vy"), s), _arr(state.get("wz"), s)
        x, y, yaw = _arr(trajectories.get("x"), s), _arr(trajectories.get("y"), s), _arr(trajectories.get("yaws"), s)
        out = np.zeros(self.B, dtype=np.float32) if costs is None else np.array(costs, dtype=np.float32, copy=True)
        coll = np.zeros(self.B, dtype=np.uint8)
        fail, furthest = C.c_int32(0), C.c_int32(-1)
        c_in = inp.to_c()
        rc = self._lib.oracle_score_tensors(
            self._h, C.byref(c_in), _p(vx), _p(vy), _p(wz), _p(x), _p(y), _p(yaw),
            int(furthest_reached_path_point), _p(out), coll.ctypes.data_as(_U8), C.byref(fail), C.byref(furthest))
        if rc != capi.MPPI_OK:
            raise RuntimeError(f"oracle_score_tensors: status {rc}")
        return out, coll, bool(fail.value), furthest.value

    def getTensor(self, which):
        B, T = self.B, self.T
        if which in (capi.TENSOR_COSTS, capi.TENSOR_SOFTMAX):
            out = np.zeros(B, dtype=np.float32)
        elif which == capi.TENSOR_CRITIC_COSTS:
            out = np.zeros((self.n_critics, B), dtype=np.float32)
        elif which == capi.TENSOR_COLLISIONS:
            out = np.zeros(B, dtype=np.uint8)
        elif which == capi.TENSOR_COST_CELLS:
            step = 1
            for k in range(self.n_critics):
                if self.cfg.critics[k].type == capi.CRITIC_TYPES["CostCritic"]:
                    step = int(self.cfg.critics[k].trajectory_point_step)
            out = np.zeros(((T - 1) // step + 1, B), dtype=np.uint32)
        elif which == capi.TENSOR_OBSTACLE_CELLS:
            out = np.zeros((T, B), dtype=np.uint32)
        else:
            out = np.zeros((T, B), dtype=np.float32)
        rc = self._lib.oracle_get_tensor(self._h, which, out.ctypes.data_as(_P), out.nbytes)
        if rc != capi.MPPI_OK:
            raise RuntimeError(f"oracle_get_tensor({which}): status {rc}")
        return out

    def getOptimalControlSequence(self):
        u = np.zeros((3, self.T), dtype=np.float32)
        self._lib.oracle_get_control_sequence(self._h, _p(u[0]), _p(u[1]), _p(u[2]))
        return u

    def setControlSequence(self, u):
        u = np.ascontiguousarray(u, dtype=np.float32).reshape(3, self.T)
        self._lib.oracle_set_control_sequence(self._h, _p(u[0]), _p(u[1]), _p(u[2]))

    def getState(self):
        st = capi.MppiOptimizerState()
        self._lib.oracle_get_state(self._h, C.byref(st))
        return st

    def setState(self, st):
        self._lib.oracle_set_state(self._h, C.byref(st))

    # --- unit-level entry points (known-answer tests transcribed from the reference) ---
    def predict(self, state):
        s = (self.T, self.B)
        arrs = [np.array(_arr(state[k], s), copy=True) for k in ("vx", "vy", "wz", "cvx", "cvy", "cwz")]
        self._lib.oracle_predict(self._h, *[_p(a) for a in arrs])
        return dict(zip(("vx", "vy", "wz", "cvx", "cvy", "cwz"), arrs))

    def pushCommandHistory(self, vx, vy, wz):
        self._lib.oracle_push_command_history(self._h, vx, vy, wz)

    def integrate(self, pose, vx, vy, wz):
        s = (self.T, self.B)
        a, b, c = _arr(vx, s), _arr(vy, s), _arr(wz, s)
        x, y, yaw = (np.zeros(s, dtype=np.float32) for _ in range(3))
        self._lib.oracle_integrate(self._h, pose[0], pose[1], pose[2], _p(a), _p(b), _p(c), _p(x), _p(y), _p(yaw))
        return x, y, yaw

    def applyControlSequenceConstraints(self, speed, inter_iteration=False):
        self._lib.oracle_apply_constraints(self._h, speed[0], speed[1], speed[2], 1 if inter_iteration else 0)

    def shiftControlSequence(self):
        self._lib.oracle_shift_control_sequence(self._h)

    def savitskyGolay(self, history):
        h = np.ascontiguousarray(history, dtype=np.float32).reshape(12).copy()
        self._lib.oracle_savitsky_golay(self._h, _p(h))
        return h.reshape(4, 3)

    def settings(self):
        v = np.zeros(16, dtype=np.float32)
        self._lib.oracle_get_settings(self._h, _p(v))
        keys = ["vx_max", "vx_min", "vy", "wz", "ax_max", "ax_min", "ay_max", "ay_min", "az_max",
                "shift_control_sequence", "controller_period", "traj_samples_to_evaluate"]
        return dict(zip(keys, v.tolist()))
