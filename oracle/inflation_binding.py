"""ctypes binding of the CPU inflation oracle (oracle/inflation_oracle.cpp) and, when present, of the reference's
own distance transform (oracle/_ref/libref_distance_transform.so).  TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg; the product never imports this module."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_U8 = C.POINTER(C.c_uint8)
_F = C.POINTER(C.c_float)


class OracleInflationConfig(C.Structure):
    _fields_ = [("resolution", C.c_double), ("inflation_radius", C.c_double), ("inscribed_radius", C.c_double),
                ("cost_scaling_factor", C.c_double), ("inflate_unknown", C.c_int32), ("inflate_around_unknown", C.c_int32)]


_LIB = None
_REF = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        p = os.path.join(HERE, "_build", "libinflation_oracle.so")
        if not os.path.exists(p):
            raise RuntimeError(f"{p} missing: run `make -C oracle`")
        l = C.CDLL(p)
        cfg = C.POINTER(OracleInflationConfig)
        l.oracle_inflation_cell_radius.restype = C.c_uint
        l.oracle_inflation_cell_radius.argtypes = [cfg]
        l.oracle_inflation_lut.restype = C.c_int
        l.oracle_inflation_lut.argtypes = [cfg, _U8, C.c_int]
        l.oracle_distance_map.restype = None
        l.oracle_distance_map.argtypes = [_U8, C.c_int, C.c_int, _F]
        l.oracle_inflation_update_costs.restype = C.c_int
        l.oracle_inflation_update_costs.argtypes = [cfg, _U8, C.c_uint, C.c_uint] + [C.c_int] * 4
        l.oracle_inflation_time.restype = C.c_double
        l.oracle_inflation_time.argtypes = [cfg, _U8, C.c_uint, C.c_uint, C.c_int]
        _LIB = l
    return _LIB


def reference_distance_transform():
    """The reference's DistanceTransform::distanceTransform2D compiled from its own header, or None."""
    global _REF
    if _REF is None:
        p = os.path.join(HERE, "_ref", "libref_distance_transform.so")
        if not os.path.exists(p):
            return None
        r = C.CDLL(p)
        r.ref_distance_map.restype = None
        r.ref_distance_map.argtypes = [_U8, C.c_int, C.c_int, _F]
        _REF = r
    return _REF


def config(resolution, inscribed_radius, inflation_radius=0.55, cost_scaling_factor=10.0, inflate_unknown=False,
           inflate_around_unknown=False) -> OracleInflationConfig:
    return OracleInflationConfig(resolution, inflation_radius, inscribed_radius, cost_scaling_factor,
                                 int(inflate_unknown), int(inflate_around_unknown))


def cell_radius(cfg) -> int:
    return int(lib().oracle_inflation_cell_radius(C.byref(cfg)))


def cost_lut(cfg) -> np.ndarray:
    n = lib().oracle_inflation_lut(C.byref(cfg), None, 0)
    out = np.zeros(n, dtype=np.uint8)
    if n:
        lib().oracle_inflation_lut(C.byref(cfg), out.ctypes.data_as(_U8), n)
    return out


def compute_cost(cfg, distance_cells: float) -> int:
    """InflationLayer::computeCost through the oracle's table (exact for multiples of 1/100 cell)."""
    lut = cost_lut(cfg)
    return int(lut[min(len(lut) - 1, int(round(distance_cells * 100)))])


def distance_map(mask: np.ndarray, ref: bool = False) -> np.ndarray:
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    out = np.zeros(mask.shape, dtype=np.float32)
    fn = reference_distance_transform().ref_distance_map if ref else lib().oracle_distance_map
    fn(mask.ctypes.data_as(_U8), mask.shape[0], mask.shape[1], out.ctypes.data_as(_F))
    return out


def update_costs(cfg, master: np.ndarray, min_i=None, min_j=None, max_i=None, max_j=None) -> np.ndarray:
    """Returns the updated copy of `master` ((size_y, size_x) uint8)."""
    out = np.ascontiguousarray(master, dtype=np.uint8).copy()
    sy, sx = out.shape
    w = (0 if min_i is None else min_i, 0 if min_j is None else min_j, sx if max_i is None else max_i,
         sy if max_j is None else max_j)
    lib().oracle_inflation_update_costs(C.byref(cfg), out.ctypes.data_as(_U8), sx, sy, *w)
    return out


def time_update(cfg, master: np.ndarray, repeats: int) -> float:
    m = np.ascontiguousarray(master, dtype=np.uint8)
    return float(lib().oracle_inflation_time(C.byref(cfg), m.ctypes.data_as(_U8), m.shape[1], m.shape[0], repeats))
