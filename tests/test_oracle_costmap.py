"""Pins the costmap read side of the oracle: FootprintCollisionChecker known answers
(/root/reference/nav2_costmap_2d/test/unit/footprint_collision_checker_test.cpp:23-160), the
Bresenham restatement against the REFERENCE's own nav2_util::LineIterator compiled from
/root/reference (oracle/_ref), InflationLayer::computeCost and Costmap2D::worldToMap."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle.binding import load, load_ref_line_iterator

U8 = C.POINTER(C.c_uint8)
DP = C.POINTER(C.c_double)


def footprint_cost(cells, res, x, y, theta, fp):
    lib = load()
    cells = np.ascontiguousarray(cells, dtype=np.uint8)
    fx = np.array([p[0] for p in fp], dtype=np.float64)
    fy = np.array([p[1] for p in fp], dtype=np.float64)
    return lib.oracle_footprint_cost_at_pose(
        cells.ctypes.data_as(U8), cells.shape[1], cells.shape[0], res, 0.0, 0.0, x, y, theta,
        fx.ctypes.data_as(DP), fy.ctypes.data_as(DP), len(fp))


# footprint_collision_checker_test.cpp:23-50 test_basic
def test_footprint_basic():
    cells = np.zeros((100, 100), dtype=np.uint8)
    diamond = [(-0.5, 0.0), (0.0, 0.5), (0.5, 0.0), (0.0, -0.5)]
    assert abs(footprint_cost(cells, 0.1, 5.0, 5.0, 0.0, diamond)) < 1e-3


# :52-62 test_point_cost and :64-81 test_world_to_map
def test_point_cost_and_world_to_map():
    lib = load()
    mx, my = C.c_uint32(), C.c_uint32()
    assert lib.oracle_world_to_map(1.0, 1.0, 0.0, 0.0, 0.1, 100, 100, C.byref(mx), C.byref(my)) == 1
    assert (mx.value, my.value) == (10, 10)
    assert lib.oracle_world_to_map(5.0, 5.0, 0.0, 0.0, 0.1, 100, 100, C.byref(mx), C.byref(my)) == 1
    assert (mx.value, my.value) == (50, 50)
    # coordinate_transform_test.cpp: outside the map is rejected
    assert lib.oracle_world_to_map(-0.01, 1.0, 0.0, 0.0, 0.1, 100, 100, C.byref(mx), C.byref(my)) == 0
    assert lib.oracle_world_to_map(10.0, 1.0, 0.0, 0.0, 0.1, 100, 100, C.byref(mx), C.byref(my)) == 0
    assert lib.oracle_world_to_map(9.999, 9.999, 0.0, 0.0, 0.1, 100, 100, C.byref(mx), C.byref(my)) == 1


SQUARE = [(-1.0, 1.0), (1.0, 1.0), (1.0, -1.0), (-1.0, -1.0)]


# :83-121 test_footprint_at_pose_with_movement
def test_footprint_at_pose_with_movement():
    cells = np.full((100, 100), 254, dtype=np.uint8)
    cells[40:61, 40:61] = 0
    assert abs(footprint_cost(cells, 0.1, 5.0, 5.0, 0.0, SQUARE)) < 1e-3
    assert abs(footprint_cost(cells, 0.1, 5.0, 4.9, 0.0, SQUARE) - 254.0) < 1e-3
    assert abs(footprint_cost(cells, 0.1, 5.0, 5.2, 0.0, SQUARE) - 254.0) < 1e-3


# :123-160 test_point_and_line_cost
def test_point_and_line_cost():
    cells = np.zeros((100, 100), dtype=np.uint8)
    cells[50, 62] = 254   # setCost(62, 50): mx = 62, my = 50
    cells[60, 39] = 254   # setCost(39, 60)
    assert abs(footprint_cost(cells, 0.1, 5.0, 5.0, 0.0, SQUARE)) < 1e-3
    assert abs(footprint_cost(cells, 0.1, 4.9, 5.0, 0.0, SQUARE) - 254.0) < 1e-3
    assert abs(footprint_cost(cells, 0.1, 5.2, 5.0, 0.0, SQUARE) - 254.0) < 1e-3


def test_footprint_off_map_is_lethal():
    """footprint_collision_checker.cpp:56-58,67-69: any vertex off the map -> LETHAL_OBSTACLE."""
    cells = np.zeros((100, 100), dtype=np.uint8)
    assert footprint_cost(cells, 0.1, 0.5, 5.0, 0.0, SQUARE) == 254.0
    assert footprint_cost(cells, 0.1, 5.0, 9.5, 0.3, SQUARE) == 254.0


def test_line_iterator_matches_reference_header():
    """oracle::LineIterator vs the reference's nav2_util/line_iterator.hpp compiled from where it lies."""
    ref = load_ref_line_iterator()
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box and no prebuilt copy)")
    lib = load()
    rng = np.random.default_rng(5)
    cap = 4096
    xs, ys = (C.c_int * cap)(), (C.c_int * cap)()
    rx, ry = (C.c_int * cap)(), (C.c_int * cap)()
    cases = [(0, 0, 0, 0), (0, 0, 10, 0), (0, 0, 0, 10), (5, 5, -5, -5), (3, 7, 3, 7), (0, 0, 7, 3), (0, 0, 3, 7),
             (10, 2, 1, 9), (-4, 6, 9, -2)]
    cases += [tuple(int(v) for v in rng.integers(-300, 300, 4)) for _ in range(500)]
    for x0, y0, x1, y1 in cases:
        n = lib.oracle_line_cells(x0, y0, x1, y1, xs, ys, cap)
        m = ref.ref_line_cells(x0, y0, x1, y1, rx, ry, cap)
        assert n == m == max(abs(x1 - x0), abs(y1 - y0)) + 1
        assert list(xs[:n]) == list(rx[:n]) and list(ys[:n]) == list(ry[:n]), (x0, y0, x1, y1)
        assert (xs[0], ys[0]) == (x0, y0) and (xs[n - 1], ys[n - 1]) == (x1, y1)


# nav2_costmap_2d/test/integration/inflation_tests.cpp:330-400 computeCost correctness
def test_inflation_compute_cost():
    lib = load()
    res, r_ins, k = 0.05, 0.2, 3.0
    assert lib.oracle_compute_cost(0.0, res, r_ins, k) == 254
    assert lib.oracle_compute_cost(2.0, res, r_ins, k) == 253            # 0.1 m <= inscribed radius
    assert lib.oracle_compute_cost(4.0, res, r_ins, k) == 253            # exactly the inscribed radius
    for cells in (5.0, 8.0, 13.0, 21.5):
        expect = int(252 * math.exp(-k * (cells * res - r_ins)))
        assert lib.oracle_compute_cost(cells, res, r_ins, k) == expect
    prev = 255
    for cells in np.arange(0.0, 30.0, 0.5):                              # monotone non-increasing
        c = lib.oracle_compute_cost(float(cells), res, r_ins, k)
        assert c <= prev
        prev = c
    # the product's helper is the same function
    from navigation2_b200 import capi
    plib = capi.load_library()
    for cells in np.arange(0.0, 30.0, 0.37):
        assert plib.mppi_inflation_compute_cost(float(cells), res, r_ins, k) == lib.oracle_compute_cost(float(cells), res, r_ins, k)


def test_find_closest_path_pt():
    """utils::findClosestPathPt (tools/utils.hpp:550-567) including its quirks."""
    lib = load()
    vec = np.array([0.0, 0.1, 0.2, 0.3, 0.4], dtype=np.float32)
    F = C.POINTER(C.c_float)
    f = lambda d, init=0: lib.oracle_find_closest_path_pt(vec.ctypes.data_as(F), len(vec), d, init)
    assert f(0.0) == 0          # first point when nothing has been travelled
    assert f(0.04) == 0 and f(0.06) == 1
    assert f(0.26) == 3 and f(0.24) == 2
    assert f(5.0) == 4          # beyond the table: last index
    assert f(0.26, 3) == 3      # never moves backwards: the search starts after `init` ...
    assert f(0.36, 3) == 4 and f(0.31, 3) == 3   # ... and compares against vec[init] when init != 0


# utils_test.cpp:269-329 findPathCosts: a path point off the costmap (index 1) or on a lethal cell (index 10) is
# invalid; the point at (4.2, 4.2) maps to cell (42, 42), beside the strip of 253 the fixture writes at row 45, and
# stays valid like every untouched point
def test_find_path_costs_reference_fixture():
    lib = load()
    cells = np.zeros((50, 50), dtype=np.uint8)           # 5 m x 5 m @ 0.1 m, origin (0, 0)
    cells[10:31, 10:31] = 254
    cells[45, 40:46] = 253
    n = 50
    px = np.zeros(n, dtype=np.float32)
    py = np.zeros(n, dtype=np.float32)
    px[1] = py[1] = 9999999.0                            # off the costmap
    px[10] = py[10] = 1.5                                # in lethal
    px[20] = py[20] = 4.2                                # near the inflated strip, on a free cell
    F = C.POINTER(C.c_float)
    valid = np.zeros(n, dtype=np.uint8)
    m = lib.oracle_find_path_costs(cells.ctypes.data_as(U8), 50, 50, 0.1, 0.0, 0.0, px.ctypes.data_as(F),
                                   py.ctypes.data_as(F), n, 0, valid.ctypes.data_as(U8))
    assert m == n - 1
    for i in range(n - 1):
        assert bool(valid[i]) == (i not in (1, 10)), i
    # the cases the reference's switch distinguishes (utils.hpp:368-381): inscribed is invalid, unknown follows
    # track_unknown_space, anything below inscribed is valid
    px[:] = 0.05
    py[:] = 0.05
    for i, c in enumerate([0, 1, 252, 253, 254, 255]):
        cells[0, i] = c
        px[i] = 0.1 * i + 0.05
    for tracking, expect in [(0, [1, 1, 1, 0, 0, 0]), (1, [1, 1, 1, 0, 0, 1])]:
        lib.oracle_find_path_costs(cells.ctypes.data_as(U8), 50, 50, 0.1, 0.0, 0.0, px.ctypes.data_as(F),
                                   py.ctypes.data_as(F), n, tracking, valid.ctypes.data_as(U8))
        assert list(valid[:6]) == expect
