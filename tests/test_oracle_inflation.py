"""Pins the CPU inflation oracle (oracle/inflation_oracle.cpp): against the reference's own distance transform
compiled from its header (oracle/_ref), against a brute-force Euclidean transform, and against the closed-form
expectations of nav2_costmap_2d/test/integration/inflation_tests.cpp.  No GPU."""
import math
import os

import numpy as np
import pytest

from oracle import inflation_binding as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NO_INFORMATION, LETHAL, INSCRIBED, FREE = 255, 254, 253, 0


def brute_force_distance(mask):
    ys, xs = np.nonzero(mask)
    h, w = mask.shape
    if len(ys) == 0:
        return np.full((h, w), np.inf)
    gy, gx = np.mgrid[0:h, 0:w]
    d2 = (gy[..., None] - ys) ** 2 + (gx[..., None] - xs) ** 2
    return np.sqrt(d2.min(axis=-1).astype(np.float32))


@pytest.mark.parametrize("seed,shape,density", [(1, (40, 57), 0.02), (2, (64, 64), 0.2), (3, (1, 90), 0.05),
                                                (4, (77, 1), 0.05), (5, (33, 120), 0.001), (6, (50, 50), 0.9)])
def test_distance_map_is_the_exact_euclidean_transform(seed, shape, density):
    rng = np.random.default_rng(seed)
    mask = (rng.random(shape) < density).astype(np.uint8)
    if mask.sum() == 0:
        mask[shape[0] // 2, shape[1] // 2] = 1
    got = ob.distance_map(mask)
    assert np.array_equal(got, brute_force_distance(mask).astype(np.float32))


def test_distance_map_matches_the_reference_header_bit_for_bit():
    if ob.reference_distance_transform() is None:
        pytest.skip("oracle/_ref/libref_distance_transform.so not built (no /root/reference on this box)")
    rng = np.random.default_rng(11)
    for shape, density in [((60, 83), 0.01), ((128, 128), 0.05), ((1, 200), 0.03), ((300, 7), 0.1), ((400, 400), 0.002),
                           ((97, 131), 0.0)]:
        mask = (rng.random(shape) < density).astype(np.uint8)
        assert np.array_equal(ob.distance_map(mask), ob.distance_map(mask, ref=True)), shape


def test_distance_map_matches_reference_golden():
    """tests/golden/inflation_reference_dt.npz holds outputs of the reference's own distanceTransform2D
    (tests/golden/make_inflation_golden.py): the pin that travels to boxes without /root/reference."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "inflation_reference_dt.npz"))
    names = sorted({k.split("__")[0] for k in gold.files})
    assert len(names) == 6
    for n in names:
        assert np.array_equal(ob.distance_map(gold[f"{n}__mask"]), gold[f"{n}__distance"]), n


def test_cell_distance_and_cost_table():
    # costmap_2d.cpp:254-258 and inflation_layer.cpp:351-366 / inflation_layer.hpp:121-135
    cfg = ob.config(resolution=0.05, inscribed_radius=0.22, inflation_radius=0.70, cost_scaling_factor=3.0)
    assert ob.cell_radius(cfg) == 14
    lut = ob.cost_lut(cfg)
    assert len(lut) == 14 * 100 + 2
    assert lut[0] == LETHAL and lut[1] == INSCRIBED
    for i in (1, 50, 440, 441, 700, 1400, 1401):
        d = i / 100.0
        want = INSCRIBED if d * 0.05 <= 0.22 else int(252 * math.exp(-3.0 * (d * 0.05 - 0.22)))
        assert lut[i] == want
    assert ob.cell_radius(ob.config(1.0, 2.1, inflation_radius=4.1)) == 5
    assert len(ob.cost_lut(ob.config(1.0, 2.1, inflation_radius=0.0))) == 0


def validate_point_inflation(grid, mx, my, cfg, inflation_radius, resolution):
    # inflation_tests.cpp:112-166 validatePointInflation
    sy, sx = grid.shape
    cells = int(math.ceil(inflation_radius / resolution))
    for dy in range(-cells, cells + 1):
        for dx in range(-cells, cells + 1):
            x, y = mx + dx, my + dy
            if x < 0 or y < 0 or x >= sx or y >= sy:
                continue
            dist = math.hypot(dx, dy)
            if dist <= inflation_radius / resolution:
                # computeCost at an arbitrary distance: evaluate the cost function itself
                if dist == 0:
                    expected = LETHAL
                elif dist * cfg.resolution <= cfg.inscribed_radius:
                    expected = INSCRIBED
                else:
                    expected = int(252 * math.exp(-cfg.cost_scaling_factor * (dist * cfg.resolution - cfg.inscribed_radius)))
                assert grid[y, x] >= expected
                if dist < inflation_radius / resolution - 0.5:
                    assert grid[y, x] > FREE


def test_inflation_should_not_create_unknowns():
    # inflation_tests.cpp:233-258: 10x10 @1 m, inscribed radius 2.1, inflation radius 4.1, obstacle at (0, 0)
    cfg = ob.config(1.0, 2.1, inflation_radius=4.1, cost_scaling_factor=1.0)
    grid = np.zeros((10, 10), np.uint8)
    grid[0, 0] = LETHAL
    out = ob.update_costs(cfg, grid)
    assert (out == NO_INFORMATION).sum() == 0
    validate_point_inflation(out, 0, 0, cfg, 4.1, 1.0)


def test_inflation_in_unknown():
    # inflation_tests.cpp:260-293: 9x9 unknown map, obstacle at the centre, inflate_unknown: only the 4 corners stay unknown
    cfg = ob.config(1.0, 2.1, inflation_radius=4.1, cost_scaling_factor=1.0, inflate_unknown=True)
    grid = np.full((9, 9), NO_INFORMATION, np.uint8)
    grid[4, 4] = LETHAL
    out = ob.update_costs(cfg, grid)
    assert (out == NO_INFORMATION).sum() == 4
    assert all(out[y, x] == NO_INFORMATION for y in (0, 8) for x in (0, 8))
    # without inflate_unknown only costs >= INSCRIBED replace unknown cells (inflation_layer.cpp:270-273)
    out2 = ob.update_costs(ob.config(1.0, 2.1, inflation_radius=4.1, cost_scaling_factor=1.0), grid)
    assert set(np.unique(out2)) == {INSCRIBED, LETHAL, NO_INFORMATION}
    assert (out2 == INSCRIBED).sum() == 12  # cells with 0 < distance <= 2.1


def test_inflation_around_unknown():
    # inflation_tests.cpp:295-325: a single NO_INFORMATION cell is inflated like an obstacle
    cfg = ob.config(1.0, 2.1, inflation_radius=4.1, cost_scaling_factor=1.0, inflate_around_unknown=True)
    grid = np.zeros((10, 10), np.uint8)
    grid[4, 4] = NO_INFORMATION
    out = ob.update_costs(cfg, grid, 0, 0, 10, 10)
    validate_point_inflation(out, 4, 4, cfg, 4.1, 1.0)
    assert out[4, 4] == LETHAL  # distance 0 looks up LETHAL and replaces the unknown cell (:270-273)


def test_cost_function_correctness():
    # inflation_tests.cpp:330-381: 100x100 @1 m, inscribed radius 5, inflation radius 10.5, obstacle at (50, 50)
    cfg = ob.config(1.0, 5.0, inflation_radius=10.5, cost_scaling_factor=1.0)
    grid = np.zeros((100, 100), np.uint8)
    grid[50, 50] = LETHAL
    out = ob.update_costs(cfg, grid)
    for i in range(0, 6):
        assert out[50, 50 + i] >= INSCRIBED and out[50, 50 - i] >= INSCRIBED
        assert out[50 + i, 50] >= INSCRIBED and out[50 - i, 50] >= INSCRIBED
    for i in range(6, 12):
        assert out[50, 50 + i] == ob.compute_cost(cfg, float(i))
        assert out[50, 50 + i] == int(252 * math.exp(-1.0 * (i - 5.0)))
    assert out[50, 50 + 12] == FREE


def test_update_window_and_merge_rules():
    rng = np.random.default_rng(5)
    cfg = ob.config(0.05, 0.22, inflation_radius=0.70, cost_scaling_factor=3.0)
    grid = np.where(rng.random((120, 150)) < 0.004, LETHAL, 0).astype(np.uint8)
    grid[rng.random(grid.shape) < 0.05] = 200  # pre-existing costs: max() keeps the larger
    whole = ob.update_costs(cfg, grid)
    part = ob.update_costs(cfg, grid, 30, 20, 90, 70)
    assert np.array_equal(part[20:70, 30:90], whole[20:70, 30:90])  # padding makes the window self-sufficient (:303-311)
    untouched = np.ones(grid.shape, bool)
    untouched[20:70, 30:90] = False
    assert np.array_equal(part[untouched], grid[untouched])
    assert (whole >= grid).all()
    assert np.array_equal(ob.update_costs(cfg, whole), whole)  # idempotent
    # window clipped to the map, empty windows and a zero radius are no-ops (:289, :297-300)
    assert np.array_equal(ob.update_costs(cfg, grid, -50, -50, 1000, 1000), whole)
    assert np.array_equal(ob.update_costs(cfg, grid, 40, 40, 40, 80), grid)
    assert np.array_equal(ob.update_costs(ob.config(0.05, 0.22, inflation_radius=0.0), grid), grid)
