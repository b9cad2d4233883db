"""GPU parity of the costmap inflation kernel (include/costmap_inflation_b200.h) against the CPU oracle, through the
C ABI: every cell bit-exact.  Cases follow nav2_costmap_2d/test/integration/inflation_tests.cpp plus ragged sizes,
windows, batches of device-resident maps and the largest supported radius."""
import math

import numpy as np
import pytest

from navigation2_b200 import inflation as infl
from oracle import inflation_binding as ob

pytestmark = pytest.mark.gpu

NO_INFORMATION, LETHAL, INSCRIBED, FREE = 255, 254, 253, 0


def make_layer(resolution, inscribed_radius, **kw):
    return infl.InflationLayer(resolution, inscribed_radius, **kw), ob.config(resolution, inscribed_radius, **kw)


def random_grid(rng, shape, lethal=0.004, unknown=0.0, costs=0.0):
    grid = np.zeros(shape, np.uint8)
    r = rng.random(shape)
    grid[r < lethal] = LETHAL
    if unknown:
        grid[rng.random(shape) < unknown] = NO_INFORMATION
    if costs:
        m = (rng.random(shape) < costs) & (grid == 0)
        grid[m] = rng.integers(1, 253, size=int(m.sum()), dtype=np.uint8)
    return grid


@pytest.mark.parametrize("shape", [(400, 400), (200, 200), (1, 300), (300, 1), (97, 131), (65, 33), (32, 64), (33, 65), (5, 5)])
@pytest.mark.parametrize("flags", [(False, False), (True, False), (False, True), (True, True)])
def test_whole_map_matches_oracle(shape, flags):
    rng = np.random.default_rng(hash((shape, flags)) % 2**32)
    layer, cfg = make_layer(0.05, 0.22, inflation_radius=0.70, cost_scaling_factor=3.0,
                            inflate_unknown=flags[0], inflate_around_unknown=flags[1])
    assert layer.cell_inflation_radius == ob.cell_radius(cfg) == 14
    assert np.array_equal(layer.cost_lut(), ob.cost_lut(cfg))
    grid = random_grid(rng, shape, lethal=0.004, unknown=0.02, costs=0.05)
    want = ob.update_costs(cfg, grid)
    got = grid.copy()
    layer.updateCosts(got, 0, 0, shape[1], shape[0])
    assert np.array_equal(got, want)
    assert layer.launch_count == 1


@pytest.mark.parametrize("radius_m,resolution", [(0.05, 0.05), (0.55, 0.05), (0.12, 0.05), (1.0, 0.025), (3.0, 0.025), (4.1, 1.0), (2.0, 0.05)])
def test_radii_from_one_cell_to_the_maximum(radius_m, resolution):
    rng = np.random.default_rng(int(radius_m * 1000))
    layer, cfg = make_layer(resolution, 0.3 * radius_m, inflation_radius=radius_m, cost_scaling_factor=2.0)
    assert layer.cell_inflation_radius == ob.cell_radius(cfg) <= infl.MAX_CELL_RADIUS
    grid = random_grid(rng, (300, 417), lethal=0.0008, unknown=0.01, costs=0.02)
    want = ob.update_costs(cfg, grid)
    got = grid.copy()
    layer.updateCosts(got, 0, 0, 417, 300)
    assert np.array_equal(got, want)


def test_dense_and_empty_maps():
    layer, cfg = make_layer(0.05, 0.22, inflation_radius=0.70, cost_scaling_factor=3.0)
    rng = np.random.default_rng(3)
    for lethal in (0.0, 0.3, 1.0):
        grid = random_grid(rng, (130, 190), lethal=lethal)
        got = grid.copy()
        layer.updateCosts(got, 0, 0, 190, 130)
        assert np.array_equal(got, ob.update_costs(cfg, grid)), lethal


def test_update_windows():
    rng = np.random.default_rng(8)
    layer, cfg = make_layer(0.05, 0.22, inflation_radius=0.70, cost_scaling_factor=3.0, inflate_around_unknown=True)
    grid = random_grid(rng, (260, 310), lethal=0.003, unknown=0.002, costs=0.05)
    for w in [(30, 20, 90, 70), (0, 0, 310, 260), (-40, -40, 1000, 1000), (300, 250, 310, 260), (0, 100, 310, 101),
              (150, 0, 151, 260), (40, 40, 40, 80), (100, 100, 50, 50), (63, 31, 129, 65)]:
        got = grid.copy()
        layer.updateCosts(got, *w)
        assert np.array_equal(got, ob.update_costs(cfg, grid, *w)), w


def test_reference_integration_cases():
    # inflation_tests.cpp:233-258, :260-293, :295-325, :330-381 through the CUDA path
    layer, cfg = make_layer(1.0, 2.1, inflation_radius=4.1, cost_scaling_factor=1.0)
    grid = np.zeros((10, 10), np.uint8)
    grid[0, 0] = LETHAL
    layer.updateCosts(grid, 0, 0, 10, 10)
    assert (grid == NO_INFORMATION).sum() == 0 and grid[0, 0] == LETHAL and grid[2, 0] == INSCRIBED

    layer, cfg = make_layer(1.0, 2.1, inflation_radius=4.1, cost_scaling_factor=1.0, inflate_unknown=True)
    grid = np.full((9, 9), NO_INFORMATION, np.uint8)
    grid[4, 4] = LETHAL
    layer.updateCosts(grid, 0, 0, 9, 9)
    assert (grid == NO_INFORMATION).sum() == 4 and all(grid[y, x] == NO_INFORMATION for y in (0, 8) for x in (0, 8))

    layer, cfg = make_layer(1.0, 2.1, inflation_radius=4.1, cost_scaling_factor=1.0, inflate_around_unknown=True)
    grid = np.zeros((10, 10), np.uint8)
    grid[4, 4] = NO_INFORMATION
    layer.updateCosts(grid, 0, 0, 10, 10)
    assert grid[4, 4] == LETHAL and grid[4, 6] == INSCRIBED
    assert grid[4, 8] == int(252 * math.exp(-1.0 * (4.0 - 2.1)))

    layer, cfg = make_layer(1.0, 5.0, inflation_radius=10.5, cost_scaling_factor=1.0)
    grid = np.zeros((100, 100), np.uint8)
    grid[50, 50] = LETHAL
    layer.updateCosts(grid, 0, 0, 100, 100)
    for i in range(0, 6):
        assert grid[50, 50 + i] >= INSCRIBED and grid[50, 50 - i] >= INSCRIBED
        assert grid[50 + i, 50] >= INSCRIBED and grid[50 - i, 50] >= INSCRIBED
    for i in range(6, 12):
        assert grid[50, 50 + i] == int(252 * math.exp(-1.0 * (i - 5.0)))
    assert grid[50, 62] == FREE


def test_fleet_of_device_resident_maps():
    import torch
    rng = np.random.default_rng(21)
    layer, cfg = make_layer(0.05, 0.22, inflation_radius=0.70, cost_scaling_factor=3.0)
    maps = np.stack([random_grid(rng, (200, 200), lethal=0.002 * (k % 4), costs=0.03) for k in range(12)])
    dev = torch.from_numpy(maps).cuda()
    layer.updateCostsDevice(dev.data_ptr(), 200, 200, n_maps=12, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = dev.cpu().numpy()
    for k in range(12):
        assert np.array_equal(got[k], ob.update_costs(cfg, maps[k])), k
    assert layer.launch_count == 1
    # a window on the resident batch
    dev = torch.from_numpy(maps).cuda()
    layer.updateCostsDevice(dev.data_ptr(), 200, 200, n_maps=12, window=(50, 60, 150, 120),
                            stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = dev.cpu().numpy()
    for k in range(12):
        assert np.array_equal(got[k], ob.update_costs(cfg, maps[k], 50, 60, 150, 120)), k


def test_large_map_matches_oracle_and_is_idempotent():
    rng = np.random.default_rng(4)
    layer, cfg = make_layer(0.05, 0.22, inflation_radius=0.70, cost_scaling_factor=3.0)
    grid = random_grid(rng, (2048, 2048), lethal=0.001, unknown=0.001, costs=0.01)
    once = grid.copy()
    layer.updateCosts(once, 0, 0, 2048, 2048)
    assert np.array_equal(once, ob.update_costs(cfg, grid))
    twice = once.copy()
    layer.updateCosts(twice, 0, 0, 2048, 2048)
    assert np.array_equal(twice, once)
    # 4096 x 4096: size-independent properties (the update only raises costs, keeps seeds, and is idempotent)
    big = random_grid(rng, (4096, 4096), lethal=0.0005)
    out = big.copy()
    layer.updateCosts(out, 0, 0, 4096, 4096)
    assert (out >= big).all() and np.array_equal(out == LETHAL, big == LETHAL)
    again = out.copy()
    layer.updateCosts(again, 0, 0, 4096, 4096)
    assert np.array_equal(again, out)
    assert np.array_equal(out[1000:1400, 2000:2400], ob.update_costs(cfg, big[900:1500, 1900:2500])[100:500, 100:500])


@pytest.mark.parametrize("radius_m", [0.55, 0.70, 1.0, 1.6, 2.0])
def test_large_grids_take_the_small_cta_path(radius_m):
    # >= 2368 tiles: 128 x 64 tiles up to 32 cells of radius (compiled-in 11 and 14, generic 20 and 32), 128-thread CTAs beyond
    rng = np.random.default_rng(int(radius_m * 100))
    layer, cfg = make_layer(0.05, 0.22, inflation_radius=radius_m, cost_scaling_factor=3.0, inflate_around_unknown=True)
    grid = random_grid(rng, (2048, 3072), lethal=0.0006, unknown=0.0003, costs=0.01)
    got = grid.copy()
    layer.updateCosts(got, 0, 0, 3072, 2048)
    assert np.array_equal(got, ob.update_costs(cfg, grid))
    # an odd row length takes the byte-load path on the same grid size
    grid = random_grid(rng, (2048, 3071), lethal=0.0006, costs=0.01)
    got = grid.copy()
    layer.updateCosts(got, 0, 0, 3071, 2048)
    assert np.array_equal(got, ob.update_costs(cfg, grid))


def test_kernel_agrees_with_reference_golden():
    """Distances produced by the REFERENCE's own distance transform (tests/golden/inflation_reference_dt.npz) decide
    the expected cost of every cell through the reference's table chain; the kernel must reproduce them."""
    import os
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inflation_reference_dt.npz"))
    layer, cfg = make_layer(0.05, 0.22, inflation_radius=0.70, cost_scaling_factor=3.0)
    lut = layer.cost_lut()
    radius = float(layer.cell_inflation_radius)
    for n in sorted({k.split("__")[0] for k in gold.files}):
        mask, dist = gold[f"{n}__mask"], gold[f"{n}__distance"]
        grid = np.where(mask != 0, LETHAL, 0).astype(np.uint8)
        got = grid.copy()
        layer.updateCosts(got, 0, 0, grid.shape[1], grid.shape[0])
        # applyInflation (inflation_layer.cpp:259-276) on the reference's distances, old cost 0 or LETHAL
        near = np.minimum(dist, np.float32(radius + 1.0))  # cells beyond the radius are skipped anyway (:259-261)
        slot = np.minimum(len(lut) - 1, (near * np.float32(100) + np.float32(0.5)).astype(np.uint32))
        want = np.where(dist > radius, grid, np.maximum(grid, lut[slot]))
        assert np.array_equal(got, want), n


def test_errors_and_degenerate_radius():
    with pytest.raises(ValueError):
        infl.InflationLayer(0.01, 0.2, inflation_radius=5.0)  # 500 cells
    layer, cfg = make_layer(0.05, 0.22, inflation_radius=0.0)
    grid = np.zeros((20, 20), np.uint8)
    grid[3, 3] = LETHAL
    before = grid.copy()
    layer.updateCosts(grid, 0, 0, 20, 20)  # cell_inflation_radius_ == 0: inflation_layer.cpp:289
    assert np.array_equal(grid, before) and layer.launch_count == 0
    with pytest.raises(ValueError):
        layer.updateCosts(np.zeros((4, 4), np.float32), 0, 0, 4, 4)


def test_resident_timing_entry_point():
    rng = np.random.default_rng(2)
    layer, _ = make_layer(0.05, 0.22, inflation_radius=0.70, cost_scaling_factor=3.0)
    maps = np.stack([random_grid(rng, (200, 200), lethal=0.003) for _ in range(4)])
    ms = layer.time_resident(maps, steps=5, warmup=2, flush_bytes=0)
    assert 0.0 < ms < 50.0 and layer.launch_count == 7
