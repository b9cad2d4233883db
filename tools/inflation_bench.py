"""Measurement of the costmap inflation kernel on one GPU (run on the B200 box; writes one JSON document).

Per workload: event-timed resident kernel passes (L2 flushed before every pass), the host-buffer call
(`InflationLayer.updateCosts`: H2D + kernel + D2H inside the timed region) and the CPU oracle on the same map
(single thread, like the reference built without OpenMP).  Algorithmic bytes: every cell read once + every
changed cell written once."""
import json
import sys
import time

import numpy as np

from navigation2_b200 import inflation as infl
from oracle import inflation_binding as ob

LETHAL = 254


def grid(rng, shape, lethal):
    g = np.zeros(shape, np.uint8)
    g[rng.random(shape) < lethal] = LETHAL
    return g


def blocks_grid(rng, shape, n_blocks):
    """SURVEY.md section 8d style map: axis-aligned lethal blocks of 4-12 cells."""
    g = np.zeros(shape, np.uint8)
    for _ in range(n_blocks):
        h, w = rng.integers(4, 13, size=2)
        y, x = rng.integers(0, shape[0] - h), rng.integers(0, shape[1] - w)
        g[y:y + h, x:x + w] = LETHAL
    return g


def main():
    peaks = {}
    try:
        peaks = json.load(open("MEASURED_PEAKS.json"))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6535.4))
    rng = np.random.default_rng(1)
    kw = dict(inflation_radius=0.70, cost_scaling_factor=3.0)
    out = {"peak_hbm_gbs": hbm, "workloads": []}
    cases = [
        ("C1 map 400x400, 20 lethal blocks", blocks_grid(rng, (400, 400), 20)[None], 200),
        ("fleet 512 x 200x200, 8 blocks each", np.stack([blocks_grid(rng, (200, 200), 8) for _ in range(512)]), 50),
        ("4096x4096, 2000 blocks", blocks_grid(rng, (4096, 4096), 2000)[None], 30),
        ("8192x8192, 8000 blocks", blocks_grid(rng, (8192, 8192), 8000)[None], 20),
        ("8192x8192, 2% random lethal cells (every tile active)", grid(rng, (8192, 8192), 0.02)[None], 20),
    ]
    for name, maps, steps in cases:
        layer = infl.InflationLayer(0.05, 0.22, **kw)
        cfg = ob.config(0.05, 0.22, **kw)
        n, sy, sx = maps.shape
        ms = layer.time_resident(maps, steps=steps, warmup=3, flush_bytes=256 << 20)
        ms_hot = layer.time_resident(maps, steps=steps, warmup=3, flush_bytes=0)
        want0 = ob.update_costs(cfg, maps[0])
        changed = int((want0 != maps[0]).sum()) * n if n == 1 else sum(
            int((ob.update_costs(cfg, m) != m).sum()) for m in maps[:8]) * n // 8
        cells = n * sy * sx
        alg_bytes = cells + changed
        # host-buffer call on one map (what InflationLayer::updateCosts replaces)
        work = maps[0].copy()
        layer.updateCosts(work, 0, 0, sx, sy)
        assert np.array_equal(work, want0), name
        lat = []
        for _ in range(10):
            work = maps[0].copy()
            t0 = time.perf_counter()
            layer.updateCosts(work, 0, 0, sx, sy)
            lat.append(time.perf_counter() - t0)
        cpu_s = ob.time_update(cfg, maps[0], 3)
        out["workloads"].append({
            "workload": name, "cells": cells, "cell_radius": layer.cell_inflation_radius,
            "kernel_ms_cold_l2": ms, "kernel_ms_hot_l2": ms_hot,
            "cells_per_s": cells / (ms * 1e-3),
            "algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": alg_bytes / (ms * 1e-3) / 1e9 / hbm,
            "host_buffer_call_ms_one_map": float(np.median(lat)) * 1e3,
            "cpu_oracle_ms_one_map": cpu_s * 1e3,
            "cpu_cells_per_s": sy * sx / cpu_s,
        })
        layer.close()
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
