"""Synthetic workloads of SURVEY.md §8(d): inflated costmaps, arc paths, seeded noise.

Used by bench.py and the tests.  Pure numpy/scipy host code; nothing here is on the hot path.
"""
from __future__ import annotations

import math

import numpy as np

from . import params as P


def inflate(lethal: np.ndarray, resolution: float, inscribed_radius: float, cost_scaling_factor: float,
            inflation_radius: float) -> np.ndarray:
    """Costmap cells from a lethal mask: Euclidean distance -> InflationLayer::computeCost
    (nav2_costmap_2d/include/nav2_costmap_2d/inflation_layer.hpp:121-135), zero beyond
    inflation_radius (plugins/inflation_layer.cpp:237-348, without its 1/100-cell LUT quantisation)."""
    from scipy import ndimage

    dist = ndimage.distance_transform_edt(~lethal)  # in cells, 0 on lethal cells
    out = np.zeros(lethal.shape, dtype=np.uint8)
    d_m = dist * resolution
    factor = np.exp(-1.0 * cost_scaling_factor * (d_m - inscribed_radius))
    decay = np.floor(252.0 * factor).clip(0, 252).astype(np.uint8)
    out[:] = decay
    out[d_m <= inscribed_radius] = 253
    out[dist == 0] = 254
    out[d_m > inflation_radius] = 0
    return out


def block_costmap(size_x: int, size_y: int, resolution: float, seed: int, n_blocks: int = 20,
                  keep_clear=None, clear_radius: float = 1.0, inscribed_radius: float | None = None,
                  cost_scaling_factor: float = 3.0, inflation_radius: float = 0.70) -> np.ndarray:
    """~n_blocks axis-aligned lethal blocks (4-12 cells) at least clear_radius from `keep_clear`
    (world x, y with origin 0,0), inflated with the bring-up inflation parameters."""
    rng = np.random.default_rng(seed)
    lethal = np.zeros((size_y, size_x), dtype=bool)
    if inscribed_radius is None:
        inscribed_radius = P.polygon_radii(P.radius_footprint(P.NAV2_BRINGUP_COSTMAP["robot_radius"]))[0]
    placed = 0
    tries = 0
    while placed < n_blocks and tries < 10000:
        tries += 1
        w, h = rng.integers(4, 13, size=2)
        x0 = int(rng.integers(0, max(1, size_x - w)))
        y0 = int(rng.integers(0, max(1, size_y - h)))
        if keep_clear is not None:
            cx, cy = (x0 + w / 2) * resolution, (y0 + h / 2) * resolution
            if math.hypot(cx - keep_clear[0], cy - keep_clear[1]) < clear_radius + max(w, h) * resolution:
                continue
        lethal[y0:y0 + h, x0:x0 + w] = True
        placed += 1
    return inflate(lethal, resolution, inscribed_radius, cost_scaling_factor, inflation_radius)


def arc_path(start, n_points: int = 80, spacing: float = 0.05, curvature: float = 0.15) -> np.ndarray:
    """n_points poses every `spacing` metres along a gentle arc from `start` (x, y, yaw); yaw = tangent."""
    x, y, yaw = (float(v) for v in start)
    out = np.zeros((n_points, 3), dtype=np.float64)
    for i in range(n_points):
        out[i] = (x, y, yaw)
        x += spacing * math.cos(yaw)
        y += spacing * math.sin(yaw)
        yaw += curvature * spacing
    return out


def seeded_noise(batch: int, steps: int, std, seed: int = 42, holonomic: bool = False):
    """N(0, sigma^2) noise tensors, each (T, B) float32 (== column-major B x T), vx then wz then vy
    like NoiseGenerator::generateNoisedControls (src/noise_generator.cpp:108-119)."""
    rng = np.random.default_rng(seed)
    nvx = (rng.standard_normal((steps, batch)) * std[0]).astype(np.float32)
    nwz = (rng.standard_normal((steps, batch)) * std[2]).astype(np.float32)
    nvy = (rng.standard_normal((steps, batch)) * std[1]).astype(np.float32) if holonomic else None
    return nvx, nvy, nwz


def reference_benchmark_fixture():
    """The reference's own harness fixture (benchmark/optimizer_benchmark.cpp:53-79): 40x40 map @0.1 m,
    one 8x8 block of cost 250 at the centre, 50 path poses stepping (+0.1, +0.1) from the map centre."""
    cells = np.zeros((40, 40), dtype=np.uint8)
    cells[16:24, 16:24] = 250
    start = (2.0, 2.0)
    path = np.zeros((50, 3), dtype=np.float64)
    for i in range(50):
        path[i] = (start[0] + 0.1 * i, start[1] + 0.1 * i, 0.0)
    return cells, 0.1, (0.0, 0.0), path
