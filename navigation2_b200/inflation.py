"""Host-side mirror of nav2_costmap_2d::InflationLayer's cost pass on top of the C ABI
(include/costmap_inflation_b200.h, kernel in csrc/costmap_inflation.cu).

Names follow the reference (nav2_costmap_2d/plugins/inflation_layer.cpp): `InflationLayer.updateCosts(master,
min_i, min_j, max_i, max_j)` updates a row-major uint8 master grid in place, `computeCost(distance)` is the cost
function of inflation_layer.hpp:121-135 read back from the library's table.  CUDA only: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from navigation2_b200 import capi

NO_INFORMATION, LETHAL_OBSTACLE, INSCRIBED_INFLATED_OBSTACLE, FREE_SPACE = 255, 254, 253, 0  # cost_values.hpp:70-74
MAX_CELL_RADIUS = 120
ERR_RADIUS_TOO_LARGE = -3


class CostmapInflationConfig(C.Structure):
    _fields_ = [
        ("resolution", C.c_double),
        ("inflation_radius", C.c_double),
        ("inscribed_radius", C.c_double),
        ("cost_scaling_factor", C.c_double),
        ("inflate_unknown", C.c_int32),
        ("inflate_around_unknown", C.c_int32),
    ]


_P = C.c_void_p
_U8 = C.POINTER(C.c_uint8)

SYMBOLS = {
    "costmap_inflation_create": (C.c_int, [C.POINTER(CostmapInflationConfig), C.POINTER(_P)]),
    "costmap_inflation_destroy": (C.c_int, [_P]),
    "costmap_inflation_last_error": (C.c_char_p, [_P]),
    "costmap_inflation_cell_radius": (C.c_uint32, [_P]),
    "costmap_inflation_lut": (C.c_int, [_P, _U8, C.c_int]),
    "costmap_inflation_update_costs": (C.c_int, [_P, _U8, C.c_uint32, C.c_uint32] + [C.c_int32] * 4),
    "costmap_inflation_update_costs_device": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t]
                                              + [C.c_int32] * 4 + [_P]),
    "costmap_inflation_time_resident": (C.c_int, [_P, _U8, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int,
                                                  C.c_size_t, C.POINTER(C.c_float)]),
    "costmap_inflation_launch_count": (C.c_uint64, [_P]),
}

_LIB = None


def load_library() -> C.CDLL:
    """The inflation entry points live in libmppi_b200.so next to the optimiser's."""
    global _LIB
    if _LIB is None:
        lib = C.CDLL(capi.library_path()) if not capi._LIB else capi._LIB
        if not hasattr(lib, "costmap_inflation_create"):
            raise RuntimeError("libmppi_b200.so has no costmap inflation entry points: rebuild it "
                               "(`python -m navigation2_b200.build`); there is no fallback path")
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB


class InflationLayer:
    """inflation_layer.cpp: parameters `inflation_radius`, `cost_scaling_factor`, `inflate_unknown`,
    `inflate_around_unknown` (:82-100, defaults 0.55 / 10.0 / false / false); `resolution` and `inscribed_radius`
    are what matchSize (:139-146) and onFootprintChanged (:207-219) take from the layered costmap."""

    def __init__(self, resolution: float, inscribed_radius: float, inflation_radius: float = 0.55,
                 cost_scaling_factor: float = 10.0, inflate_unknown: bool = False,
                 inflate_around_unknown: bool = False):
        self._lib = load_library()
        self.config = CostmapInflationConfig(resolution, inflation_radius, inscribed_radius, cost_scaling_factor,
                                             int(inflate_unknown), int(inflate_around_unknown))
        self._h = _P()
        rc = self._lib.costmap_inflation_create(C.byref(self.config), C.byref(self._h))
        if rc == ERR_RADIUS_TOO_LARGE:
            raise ValueError(f"inflation radius above {MAX_CELL_RADIUS} cells is not supported")
        if rc != 0:
            raise RuntimeError(f"costmap_inflation_create failed ({rc}): a CUDA device is required, there is no CPU path")

    def close(self) -> None:
        if self._h:
            self._lib.costmap_inflation_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError(f"costmap inflation failed ({rc}): {self._lib.costmap_inflation_last_error(self._h).decode()}")

    @property
    def cell_inflation_radius(self) -> int:
        return int(self._lib.costmap_inflation_cell_radius(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.costmap_inflation_launch_count(self._h))

    def cost_lut(self) -> np.ndarray:
        n = self._lib.costmap_inflation_lut(self._h, None, 0)
        out = np.zeros(max(n, 0), dtype=np.uint8)
        if n > 0:
            self._lib.costmap_inflation_lut(self._h, out.ctypes.data_as(_U8), n)
        return out

    def updateCosts(self, master: np.ndarray, min_i: int, min_j: int, max_i: int, max_j: int) -> None:
        """InflationLayer::updateCosts on a (size_y, size_x) uint8 array in host memory, in place."""
        if master.dtype != np.uint8 or master.ndim != 2 or not master.flags.c_contiguous:
            raise ValueError("master grid must be a C-contiguous (size_y, size_x) uint8 array")
        size_y, size_x = master.shape
        self._check(self._lib.costmap_inflation_update_costs(
            self._h, master.ctypes.data_as(_U8), size_x, size_y, min_i, min_j, max_i, max_j))

    def updateCostsDevice(self, device_ptr: int, size_x: int, size_y: int, n_maps: int = 1, map_stride: int = 0,
                          window=None, stream: int = 0) -> None:
        """The same on grids resident in device memory (e.g. a torch uint8 tensor's data_ptr()); not synchronised."""
        w = window or (0, 0, size_x, size_y)
        self._check(self._lib.costmap_inflation_update_costs_device(
            self._h, _P(device_ptr), size_x, size_y, n_maps, map_stride or size_x * size_y, *w, _P(stream)))

    def time_resident(self, maps: np.ndarray, steps: int, warmup: int, flush_bytes: int = 0) -> float:
        """Mean kernel milliseconds per whole-map pass over `maps` ((n, size_y, size_x) uint8) kept in HBM."""
        maps = np.ascontiguousarray(maps, dtype=np.uint8)
        n, size_y, size_x = maps.shape
        ms = C.c_float()
        self._check(self._lib.costmap_inflation_time_resident(
            self._h, maps.ctypes.data_as(_U8), size_x, size_y, n, steps, warmup, flush_bytes, C.byref(ms)))
        return float(ms.value)
