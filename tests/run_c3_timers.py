import os, sys
os.environ["MPPI_B200_HOST_TIMERS"] = "1"
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
wl = bench.workload_c3()
for in_place in (False, True):
    opt = bench.setup_engine(wl)
    opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 10, in_place=in_place)
    t = opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 200, in_place=in_place)
    print("in_place" if in_place else "two_call", "e2e p50 us", np.median(t) * 1e6)
    opt.shutdown()
