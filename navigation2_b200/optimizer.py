"""Python host mirror of mppi::Optimizer on top of the C ABI.

Method names and argument meaning follow the reference class
(nav2_mppi_controller/include/nav2_mppi_controller/optimizer.hpp:59-333); ROS message types are
replaced by plain tuples/arrays.  All compute happens in libmppi_b200.so on the GPU; if the
library or a device is missing the constructor raises — there is no fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from . import capi


class ControllerException(RuntimeError):
    """nav2_core::ControllerException (nav2_core/controller_exceptions.hpp)."""


class NoValidControl(ControllerException):
    """nav2_core::NoValidControl."""


class CudaUnavailable(RuntimeError):
    """The CUDA path could not run; nothing else exists to run instead."""


def _raise_for(lib, handle, rc: int) -> None:
    if rc == capi.MPPI_OK:
        return
    msg = (lib.mppi_last_error(handle) or b"").decode()
    text = f"{lib.mppi_status_string(rc).decode()}: {msg}"
    if rc == capi.MPPI_ERR_NO_VALID_CONTROL:
        raise NoValidControl(text)
    if rc == capi.MPPI_ERR_CONFIG:
        raise ControllerException(text)
    if rc == capi.MPPI_ERR_CUDA:
        raise CudaUnavailable(text)
    raise RuntimeError(text)


@dataclass
class CycleInput:
    """Arguments of Optimizer::evalControl for one robot (optimizer.hpp:91-96)."""
    pose: tuple[float, float, float]              # robot_pose: x, y, yaw
    speed: tuple[float, float, float]             # robot_speed: vx, vy, wz
    path: np.ndarray                              # plan: (n, 3) float64 x, y, yaw
    goal: tuple[float, float, float] = (0.0, 0.0, 0.0)
    goal_checker_xy_tolerance: float | None = None  # None = goal_checker == nullptr
    local_path_length: float | None = None          # None = calculate_path_length(plan)
    _keep: list = field(default_factory=list, repr=False)

    def to_c(self) -> capi.MppiCycleIn:
        path = np.ascontiguousarray(np.asarray(self.path, dtype=np.float64).reshape(-1, 3))
        px = np.ascontiguousarray(path[:, 0]); py = np.ascontiguousarray(path[:, 1])
        pyaw = np.ascontiguousarray(path[:, 2])
        self._keep = [px, py, pyaw]
        c = capi.MppiCycleIn()
        c.pose_x, c.pose_y, c.pose_yaw = (float(v) for v in self.pose)
        c.speed_vx, c.speed_vy, c.speed_wz = (float(v) for v in self.speed)
        dp = C.POINTER(C.c_double)
        c.path_x = px.ctypes.data_as(dp); c.path_y = py.ctypes.data_as(dp); c.path_yaw = pyaw.ctypes.data_as(dp)
        c.n_path = path.shape[0]
        c.goal_x, c.goal_y, c.goal_yaw = (float(v) for v in self.goal)
        c.has_goal_checker = 0 if self.goal_checker_xy_tolerance is None else 1
        c.goal_checker_xy_tolerance = float(self.goal_checker_xy_tolerance or 0.0)
        c.has_local_path_length = 0 if self.local_path_length is None else 1
        c.local_path_length = float(self.local_path_length or 0.0)
        return c


@dataclass
class CycleResult:
    """What evalControl returns plus the state the controller publishes."""
    status: int
    cmd: tuple[float, float, float]               # TwistStamped linear.x, linear.y, angular.z
    control_sequence: np.ndarray                  # (3, T): vx, vy, wz (before the shift)
    optimal_trajectory: np.ndarray                # (T, 3): x, y, yaw
    fail_flag: bool
    validation: int
    retries: int
    furthest_reached_path_point: int


class EngineBase:
    """Shared marshalling for anything that exposes the mppi_* function shapes."""

    _lib = None
    _prefix = "mppi_"
    _has_robot_arg = True

    def _fn(self, name):
        return getattr(self._lib, self._prefix + name)


class Optimizer(EngineBase):
    """mppi::Optimizer for n_robots robots (1 = the reference's controller)."""

    def __init__(self, cfg: capi.MppiConfig):
        self._lib = capi.load_library()
        capi.check_abi(self._lib)
        self.cfg = cfg
        self.R = int(cfg.n_robots)
        self.T = int(cfg.time_steps)
        self.B = int(cfg.batch_size) // max(1, int(cfg.shard_world))
        self.n_critics = int(cfg.n_critics)
        self._h = C.c_void_p()
        rc = self._lib.mppi_create(C.byref(cfg), C.byref(self._h))
        if rc != capi.MPPI_OK:
            self._h = None
            _raise_for(self._lib, None, rc)

    # -- lifecycle ---------------------------------------------------------------------------
    def shutdown(self) -> None:
        if getattr(self, "_h", None):
            self._lib.mppi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.shutdown()
        except Exception:
            pass

    def reset(self, reset_dynamic_speed_limits: bool = True, robot: int | None = None) -> None:
        r = 0xFFFFFFFF if robot is None else robot
        _raise_for(self._lib, self._h, self._lib.mppi_reset(self._h, r, 1 if reset_dynamic_speed_limits else 0))

    def setSpeedLimit(self, speed_limit: float, percentage: bool) -> None:
        _raise_for(self._lib, self._h, self._lib.mppi_set_speed_limit(self._h, float(speed_limit), 1 if percentage else 0))

    def setConfig(self, cfg: capi.MppiConfig) -> None:
        _raise_for(self._lib, self._h, self._lib.mppi_set_config(self._h, C.byref(cfg)))
        self.cfg = cfg
        self.T = int(cfg.time_steps)
        self.B = int(cfg.batch_size) // max(1, int(cfg.shard_world))
        self.n_critics = int(cfg.n_critics)

    # -- per-cycle data ----------------------------------------------------------------------
    def setCostmap(self, cells: np.ndarray, resolution: float, origin_x: float, origin_y: float, robot: int = 0) -> None:
        """cells: (size_y, size_x) uint8, row-major (Costmap2D::getCharMap)."""
        cells = np.ascontiguousarray(cells, dtype=np.uint8)
        sy, sx = cells.shape
        rc = self._lib.mppi_upload_costmap(
            self._h, robot, cells.ctypes.data_as(C.POINTER(C.c_uint8)), sx, sy, float(resolution),
            float(origin_x), float(origin_y))
        _raise_for(self._lib, self._h, rc)

    def setCostmapDevice(self, device_ptr: int, size_x: int, size_y: int, resolution: float, origin_x: float, origin_y: float,
                         robot: int = 0, row_pitch: int = 0, producer_stream: int | None = None) -> None:
        """The costmap where a device-side producer left it (e.g. a torch uint8 tensor's data_ptr() after
        InflationLayer.updateCostsDevice): device-to-device hand-over, no H2D (mppi_set_costmap_device)."""
        rc = self._lib.mppi_set_costmap_device(
            self._h, robot, C.c_void_p(device_ptr), int(size_x), int(size_y), int(row_pitch or size_x), float(resolution),
            float(origin_x), float(origin_y), C.c_void_p(producer_stream) if producer_stream else None)
        _raise_for(self._lib, self._h, rc)

    def setNoise(self, nvx: np.ndarray, nvy: np.ndarray | None, nwz: np.ndarray, robot: int = 0) -> None:
        """Inject noise tensors, each (T, B) float32 C-contiguous == column-major B x T."""
        fp = C.POINTER(C.c_float)
        arrs = []
        for a in (nvx, nvy, nwz):
            if a is None:
                arrs.append(None)
                continue
            a = np.ascontiguousarray(a, dtype=np.float32)
            if a.shape != (self.T, self.B):
                raise ValueError(f"noise must have shape (T={self.T}, B={self.B}), got {a.shape}")
            arrs.append(a)
        ptr = [a.ctypes.data_as(fp) if a is not None else None for a in arrs]
        _raise_for(self._lib, self._h, self._lib.mppi_set_noise(self._h, robot, ptr[0], ptr[1], ptr[2]))

    def generateNoise(self, robot: int = 0) -> None:
        _raise_for(self._lib, self._h, self._lib.mppi_generate_noise(self._h, robot))

    def evalControl(self, inputs: CycleInput | Sequence[CycleInput], raise_on_failure: bool = True, costmaps=None):
        """Optimizer::evalControl for every robot; returns one CycleResult (or a list for fleets).

        costmaps: None (the costmaps were sent with setCostmap), or one (cells, resolution, origin_x, origin_y)
        tuple (a list of them for fleets) read in place for this cycle (mppi_optimize_in_place)."""
        single = isinstance(inputs, CycleInput)
        ins = [inputs] if single else list(inputs)
        if len(ins) != self.R:
            raise ValueError(f"expected {self.R} CycleInput, got {len(ins)}")
        c_in = (capi.MppiCycleIn * self.R)(*[i.to_c() for i in ins])
        c_out = (capi.MppiCycleOut * self.R)()
        T = self.T
        u = np.zeros((self.R, 3, T), dtype=np.float32)
        traj = np.zeros((self.R, 3, T), dtype=np.float32)
        fp = C.POINTER(C.c_float)
        for r in range(self.R):
            c_out[r].control_vx = u[r, 0].ctypes.data_as(fp)
            c_out[r].control_vy = u[r, 1].ctypes.data_as(fp)
            c_out[r].control_wz = u[r, 2].ctypes.data_as(fp)
            c_out[r].optimal_trajectory = traj[r].ctypes.data_as(fp)
        if costmaps is None:
            rc = self._fn("optimize")(self._h, c_in, c_out)
        else:
            maps = [costmaps] if isinstance(costmaps, tuple) else list(costmaps)
            if len(maps) != self.R:
                raise ValueError(f"expected {self.R} costmaps, got {len(maps)}")
            keep = [np.ascontiguousarray(m[0], dtype=np.uint8) for m in maps]
            views = (capi.MppiCostmapView * self.R)()
            for r, (m, cells) in enumerate(zip(maps, keep)):
                views[r].cells = cells.ctypes.data_as(C.POINTER(C.c_uint8))
                views[r].size_x, views[r].size_y = cells.shape[1], cells.shape[0]
                views[r].resolution, views[r].origin_x, views[r].origin_y = float(m[1]), float(m[2]), float(m[3])
            rc = self._fn("optimize_in_place")(self._h, views, c_in, c_out)
        _raise_for(self._lib, self._h, rc)
        results = []
        for r in range(self.R):
            o = c_out[r]
            results.append(CycleResult(
                status=o.status, cmd=(o.cmd_vx, o.cmd_vy, o.cmd_wz), control_sequence=u[r].copy(),
                optimal_trajectory=traj[r].T.copy(), fail_flag=bool(o.fail_flag), validation=o.validation,
                retries=o.retries, furthest_reached_path_point=o.furthest_reached_path_point))
        if raise_on_failure:
            for r, res in enumerate(results):
                if res.status != capi.MPPI_OK:
                    _raise_for(self._lib, self._h, res.status)
        return results[0] if single else results

    def scoreTrajectories(self, inp: CycleInput, state: dict, trajectories: dict, costs: np.ndarray | None = None,
                          furthest_reached_path_point: int = -1, robot: int = 0):
        """CriticManager::evalTrajectoriesScores on caller tensors.

        state: {"vx","vy","wz"} and trajectories: {"x","y","yaws"}, each (T, B) float32.
        Returns (costs, collisions, fail_flag, furthest)."""
        fp = C.POINTER(C.c_float)

        def arr(a):
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=np.float32)
            if a.shape != (self.T, self.B):
                raise ValueError(f"tensor must have shape (T={self.T}, B={self.B}), got {a.shape}")
            return a

        vx, vy, wz = arr(state.get("vx")), arr(state.get("vy")), arr(state.get("wz"))
        x, y, yaw = arr(trajectories.get("x")), arr(trajectories.get("y")), arr(trajectories.get("yaws"))
        out = np.zeros(self.B, dtype=np.float32) if costs is None else np.array(costs, dtype=np.float32, copy=True)
        coll = np.zeros(self.B, dtype=np.uint8)
        fail = C.c_int32(0)
        furthest = C.c_int32(-1)
        c_in = inp.to_c()
        p = lambda a: a.ctypes.data_as(fp) if a is not None else None
        rc = self._fn("score_tensors")(
            self._h, robot, C.byref(c_in), p(vx), p(vy), p(wz), p(x), p(y), p(yaw),
            int(furthest_reached_path_point), out.ctypes.data_as(fp),
            coll.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(fail), C.byref(furthest))
        _raise_for(self._lib, self._h, rc)
        return out, coll, bool(fail.value), furthest.value

    # -- accessors ---------------------------------------------------------------------------
    def getTensor(self, which: int, robot: int = 0) -> np.ndarray:
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
        elif which == capi.TENSOR_VIS_XY:
            ts, tt = max(1, int(self.cfg.vis_trajectory_step)), max(1, int(self.cfg.vis_time_step))
            out = np.zeros(((B + ts - 1) // ts, (T + tt - 1) // tt, 2), dtype=np.float32)
        else:
            out = np.zeros((T, B), dtype=np.float32)
        rc = self._fn("get_tensor")(self._h, robot, which, out.ctypes.data_as(C.c_void_p), out.nbytes)
        _raise_for(self._lib, self._h, rc)
        return out

    def getGeneratedTrajectories(self, robot: int = 0) -> dict:
        """models::Trajectories of the last optimize() (needs materialize_tensors)."""
        return {"x": self.getTensor(capi.TENSOR_X, robot), "y": self.getTensor(capi.TENSOR_Y, robot),
                "yaws": self.getTensor(capi.TENSOR_YAW, robot)}

    def getOptimalControlSequence(self, robot: int = 0) -> np.ndarray:
        u = np.zeros((3, self.T), dtype=np.float32)
        fp = C.POINTER(C.c_float)
        rc = self._fn("get_control_sequence")(self._h, robot, u[0].ctypes.data_as(fp), u[1].ctypes.data_as(fp),
                                              u[2].ctypes.data_as(fp))
        _raise_for(self._lib, self._h, rc)
        return u

    def setControlSequence(self, u: np.ndarray, robot: int = 0) -> None:
        u = np.ascontiguousarray(u, dtype=np.float32).reshape(3, self.T)
        fp = C.POINTER(C.c_float)
        rc = self._fn("set_control_sequence")(self._h, robot, u[0].ctypes.data_as(fp), u[1].ctypes.data_as(fp),
                                              u[2].ctypes.data_as(fp))
        _raise_for(self._lib, self._h, rc)

    def getState(self, robot: int = 0) -> capi.MppiOptimizerState:
        st = capi.MppiOptimizerState()
        _raise_for(self._lib, self._h, self._lib.mppi_get_state(self._h, robot, C.byref(st)))
        return st

    def setState(self, st: capi.MppiOptimizerState, robot: int = 0) -> None:
        _raise_for(self._lib, self._h, self._lib.mppi_set_state(self._h, robot, C.byref(st)))

    # -- measurement -------------------------------------------------------------------------
    def setResidentCapture(self, enabled: bool) -> None:
        _raise_for(self._lib, self._h, self._lib.mppi_set_resident_capture(self._h, 1 if enabled else 0))

    def windowMissCount(self) -> int:
        return int(self._lib.mppi_window_miss_count(self._h))

    def h2dByteCount(self) -> int:
        return int(self._lib.mppi_h2d_byte_count(self._h))

    def timeE2E(self, cells: np.ndarray, resolution: float, origin, inputs, cycles: int, in_place: bool = False) -> np.ndarray:
        """Per-call seconds of `cycles` control cycles timed inside the C library: (upload costmap; optimize)
        pairs, or mppi_optimize_in_place calls."""
        ins = [inputs] if isinstance(inputs, CycleInput) else list(inputs)
        c_in = (capi.MppiCycleIn * self.R)(*[i.to_c() for i in ins])
        c_out = (capi.MppiCycleOut * self.R)()
        cells = np.ascontiguousarray(cells, dtype=np.uint8)
        t = np.zeros(cycles, dtype=np.float64)
        rc = self._lib.mppi_time_e2e(
            self._h, cells.ctypes.data_as(C.POINTER(C.c_uint8)), cells.shape[1], cells.shape[0], float(resolution),
            float(origin[0]), float(origin[1]), c_in, c_out, int(cycles), 1 if in_place else 0,
            t.ctypes.data_as(C.POINTER(C.c_double)))
        _raise_for(self._lib, self._h, rc)
        self.last_timed_status = [(int(c_out[r].status), int(c_out[r].retries)) for r in range(self.R)]
        return t

    def timeE2EViews(self, costmaps, inputs, cycles: int) -> np.ndarray:
        """Per-call seconds of `cycles` mppi_optimize_in_place calls with one (cells, resolution, origin_x, origin_y) per
        robot, timed inside the C library."""
        ins = [inputs] if isinstance(inputs, CycleInput) else list(inputs)
        maps = [costmaps] if isinstance(costmaps, tuple) else list(costmaps)
        if len(ins) != self.R or len(maps) != self.R:
            raise ValueError(f"expected {self.R} inputs and costmaps")
        c_in = (capi.MppiCycleIn * self.R)(*[i.to_c() for i in ins])
        c_out = (capi.MppiCycleOut * self.R)()
        keep = [np.ascontiguousarray(m[0], dtype=np.uint8) for m in maps]
        views = (capi.MppiCostmapView * self.R)()
        for r, (m, cells) in enumerate(zip(maps, keep)):
            views[r].cells = cells.ctypes.data_as(C.POINTER(C.c_uint8))
            views[r].size_x, views[r].size_y = cells.shape[1], cells.shape[0]
            views[r].resolution, views[r].origin_x, views[r].origin_y = float(m[1]), float(m[2]), float(m[3])
        t = np.zeros(cycles, dtype=np.float64)
        rc = self._lib.mppi_time_e2e_views(self._h, views, c_in, c_out, int(cycles), t.ctypes.data_as(C.POINTER(C.c_double)))
        _raise_for(self._lib, self._h, rc)
        self.last_timed_status = [(int(c_out[r].status), int(c_out[r].retries)) for r in range(self.R)]
        return t

    def enqueueResident(self, cuda_stream: int | None = None) -> None:
        _raise_for(self._lib, self._h, self._lib.mppi_enqueue_resident(self._h, cuda_stream))

    def timeResident(self, steps: int, warmup: int = 3, flush_bytes: int = 0) -> np.ndarray:
        """Device milliseconds of `steps` resident replays, CUDA events inside the library (no binding overhead
        between the events); flush_bytes > 0 flushes L2 before every replay."""
        ms = np.zeros(steps, dtype=np.float32)
        rc = self._lib.mppi_time_resident(self._h, int(steps), int(warmup), int(flush_bytes), ms.ctypes.data_as(C.POINTER(C.c_float)))
        _raise_for(self._lib, self._h, rc)
        return ms.astype(np.float64)

    def enqueueScoreResident(self, cuda_stream: int | None = None) -> None:
        _raise_for(self._lib, self._h, self._lib.mppi_enqueue_score_resident(self._h, cuda_stream))

    def synchronize(self) -> None:
        _raise_for(self._lib, self._h, self._lib.mppi_device_synchronize(self._h))

    def stream(self) -> int:
        return int(self._lib.mppi_stream(self._h) or 0)

    def launchCount(self) -> int:
        return int(self._lib.mppi_launch_count(self._h))

    def setKernelTiming(self, enabled: bool) -> None:
        _raise_for(self._lib, self._h, self._lib.mppi_set_kernel_timing(self._h, 1 if enabled else 0))

    def kernelCTimeMs(self, reset: bool = True) -> tuple[float, int]:
        ms = C.c_double(0.0)
        n = C.c_uint64(0)
        _raise_for(self._lib, self._h, self._lib.mppi_kernel_c_time_ms(self._h, 1 if reset else 0, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def kernelTimeMs(self, reset: bool = True) -> tuple[float, int]:
        ms = C.c_double(0.0)
        n = C.c_uint64(0)
        _raise_for(self._lib, self._h, self._lib.mppi_kernel_time_ms(self._h, 1 if reset else 0, C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)
