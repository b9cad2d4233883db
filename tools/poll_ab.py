"""Ad-hoc: result-stamp polling on / off on the same box: end-to-end p50 and resident hot p50."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
wl = bench.workload_c1()
for rep in range(2):
    for poll in ("1", "0"):
        os.environ["MPPI_B200_NO_POLL"] = "0" if poll == "1" else "1"
        opt = bench.setup_engine(wl)
        opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 20, in_place=True)
        t = opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 1000, in_place=True)
        print(f"poll={poll}  e2e p50 {np.median(t) * 1e6:6.2f} us  p10 {np.percentile(t, 10) * 1e6:6.2f}")
        opt.shutdown()
