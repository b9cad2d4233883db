"""Ad-hoc: host-side phase breakdown of one control cycle at C1 (two-call and in-place entry points)."""
import os, sys
os.environ["MPPI_B200_HOST_TIMERS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
wl = bench.workload_c1()
for in_place in (False, True):
    opt = bench.setup_engine(wl)
    opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 20, in_place=in_place)
    t = opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 500, in_place=in_place)
    print("in_place" if in_place else "two_call", "e2e p50 us", np.median(t) * 1e6, "mean", t.mean() * 1e6, "p10", np.percentile(t, 10) * 1e6)
    opt.shutdown()
