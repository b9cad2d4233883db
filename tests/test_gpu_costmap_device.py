"""SURVEY.md section 8f row 3, the device-resident hand-off: the costmap is inflated on the GPU
(costmap_inflation_update_costs_device; reference producer nav2_costmap_2d/plugins/inflation_layer.cpp:283-348) and
consumed by the optimiser where it lies (mppi_set_costmap_device; reference consumer cost_critic.cpp:138-144) — no byte
of the map crosses PCIe.  Checked against the two-call host path (download the inflated map, mppi_upload_costmap) bit
for bit, and against the CPU oracle."""
import numpy as np
import pytest

import navigation2_b200 as nb
from navigation2_b200 import capi, inflation, params as P, scenarios as S

from util import assert_cycle_close

pytestmark = pytest.mark.gpu


def _raw_seeds(size_y, size_x, seed, clear=(100, 100)):
    rng = np.random.default_rng(seed)
    raw = np.zeros((size_y, size_x), dtype=np.uint8)
    for _ in range(25):
        w, h = rng.integers(3, 10, size=2)
        x0, y0 = int(rng.integers(0, size_x - w)), int(rng.integers(0, size_y - h))
        if abs(x0 - clear[0]) < 22 and abs(y0 - clear[1]) < 22:
            continue
        raw[y0:y0 + h, x0:x0 + w] = 254
    return raw


@pytest.mark.parametrize("batch,size", [(2000, (200, 200)), (2000, (197, 203)), (16384, (200, 200))])
def test_inflated_on_device_and_consumed_in_place(batch, size):
    """2000 trajectories: the single-launch kernel reads its window from the resident map; 16384: kernels A-D.
    197 x 203: rows are re-pitched to 208 bytes by the hand-over.  H2D traffic per cycle is the cycle block only."""
    import torch
    from oracle.binding import OracleOptimizer
    size_y, size_x = size
    cfg = P.nav2_bringup_config(batch_size=batch)
    layer = inflation.InflationLayer(0.05, cfg.inscribed_radius, inflation_radius=0.70, cost_scaling_factor=3.0)
    on_device, two_call, cpu = nb.Optimizer(cfg), nb.Optimizer(cfg), OracleOptimizer(cfg)
    nvx, nvy, nwz = S.seeded_noise(batch, 56, (cfg.vx_std, cfg.vy_std, cfg.wz_std), seed=11)
    for e in (on_device, two_call, cpu):
        e.setNoise(nvx, nvy, nwz)
    stream = torch.cuda.Stream()
    pose = [5.0, 5.0, 0.2]
    speed = (0.2, 0.0, 0.0)
    h2d = []
    for cycle in range(5):
        raw = _raw_seeds(size_y, size_x, seed=40 + cycle // 2)                 # the map changes every second cycle
        d_map = torch.from_numpy(raw).cuda()
        stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(stream):
            layer.updateCostsDevice(d_map.data_ptr(), size_x, size_y, stream=stream.cuda_stream)
        on_device.setCostmapDevice(d_map.data_ptr(), size_x, size_y, 0.05, 0.0, 0.0, producer_stream=stream.cuda_stream)
        path = S.arc_path(tuple(pose), n_points=90, curvature=0.1)
        inp = nb.CycleInput(pose=tuple(pose), speed=speed, path=path, goal=tuple(path[-1]), goal_checker_xy_tolerance=0.25)
        b0 = on_device.h2dByteCount()
        a = on_device.evalControl(inp)
        h2d.append(on_device.h2dByteCount() - b0)
        inflated = d_map.cpu().numpy()
        assert (inflated != raw).sum() > 500, "the inflation pass must have run before the optimiser read the map"
        two_call.setCostmap(inflated, 0.05, 0.0, 0.0)
        cpu.setCostmap(inflated, 0.05, 0.0, 0.0)
        b = two_call.evalControl(inp)
        c = cpu.evalControl(inp)
        np.testing.assert_array_equal(a.control_sequence, b.control_sequence, err_msg=f"cycle {cycle}")
        np.testing.assert_array_equal(a.optimal_trajectory, b.optimal_trajectory)
        assert a.cmd == b.cmd and a.furthest_reached_path_point == b.furthest_reached_path_point
        np.testing.assert_array_equal(on_device.getTensor(capi.TENSOR_COSTS), two_call.getTensor(capi.TENSOR_COSTS))
        assert_cycle_close(a, c, label=f"device hand-over vs oracle, cycle {cycle}")
        for e in (on_device, two_call):
            e.setState(cpu.getState()); e.setControlSequence(cpu.getOptimalControlSequence())
        pose[0] += 0.05 * c.cmd[0] * np.cos(pose[2]); pose[1] += 0.05 * c.cmd[0] * np.sin(pose[2]); pose[2] += 0.05 * c.cmd[2]
        speed = (c.cmd[0], 0.0, c.cmd[2])
    assert max(h2d) < 16384, f"only the cycle block may go up, saw {h2d} bytes per cycle"
    for e in (on_device, two_call, cpu):
        e.shutdown()
    layer.close()


def test_device_map_survives_arena_growth_and_is_replaced_by_a_host_map():
    """A longer path than the cycle block's capacity reallocates the arena: the device-only map moves with it.  A later
    mppi_upload_costmap replaces it."""
    import torch
    cfg = P.nav2_bringup_config(batch_size=1024)
    a, b = nb.Optimizer(cfg), nb.Optimizer(cfg)
    nvx, nvy, nwz = S.seeded_noise(1024, 56, (cfg.vx_std, cfg.vy_std, cfg.wz_std), seed=3)
    for e in (a, b):
        e.setNoise(nvx, nvy, nwz)
    cells = S.block_costmap(200, 200, 0.05, seed=8, keep_clear=(5.0, 5.0))
    d_map = torch.from_numpy(cells).cuda()
    torch.cuda.synchronize()
    a.setCostmapDevice(d_map.data_ptr(), 200, 200, 0.05, 0.0, 0.0)
    b.setCostmap(cells, 0.05, 0.0, 0.0)
    for n_points in (80, 700, 90):
        path = S.arc_path((5.0, 5.0, 0.0), n_points=n_points, spacing=0.01 if n_points > 500 else 0.05)
        inp = nb.CycleInput(pose=(5.0, 5.0, 0.0), speed=(0.1, 0.0, 0.0), path=path, goal_checker_xy_tolerance=0.25)
        ra, rb = a.evalControl(inp), b.evalControl(inp)
        np.testing.assert_array_equal(ra.control_sequence, rb.control_sequence, err_msg=f"n_points {n_points}")
    other = S.block_costmap(200, 200, 0.05, seed=9, keep_clear=(5.0, 5.0))
    a.setCostmap(other, 0.05, 0.0, 0.0); b.setCostmap(other, 0.05, 0.0, 0.0)
    ra, rb = a.evalControl(inp), b.evalControl(inp)
    np.testing.assert_array_equal(ra.control_sequence, rb.control_sequence)
    a.shutdown(); b.shutdown()
