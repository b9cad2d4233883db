"""Two launches of the inflation kernel on an 8192 x 8192 map for `ncu` (blocks map, then 2% random seeds)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from navigation2_b200 import inflation as infl
from run_inflation_bench import blocks_grid, grid  # noqa: E402

rng = np.random.default_rng(1)
layer = infl.InflationLayer(0.05, 0.22, inflation_radius=0.70, cost_scaling_factor=3.0)
for maps in (blocks_grid(rng, (8192, 8192), 8000)[None], grid(rng, (8192, 8192), 0.02)[None]):
    print(layer.time_resident(maps, steps=1, warmup=0, flush_bytes=0))
