"""Ad-hoc: host-side phase breakdown of one (upload costmap + optimize) cycle at C1."""
import os, sys
os.environ["MPPI_B200_HOST_TIMERS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
wl = bench.workload_c1()
opt = bench.setup_engine(wl)
opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 20)
t = opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 500)
print("e2e p50 us", np.median(t) * 1e6, "mean", t.mean() * 1e6)
opt.shutdown()
