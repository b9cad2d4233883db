"""Ad-hoc: end-to-end latency of the bring-up parameter set (visualize + regenerate_noises, nav2_params.yaml:157-159)
next to the code defaults the headline benchmark uses."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from navigation2_b200 import params as P
wl = bench.workload_c1()
for label, kw in (("code defaults", {}), ("visualize", {"visualize": 1}), ("regenerate_noises", {"regenerate_noises": 1}),
                  ("bring-up: visualize + regenerate_noises", {"visualize": 1, "regenerate_noises": 1})):
    cfg = P.nav2_bringup_config(**kw)
    opt = bench.setup_engine(wl, cfg=cfg)
    if kw.get("regenerate_noises"):
        opt.reset()                       # drop the injected noise: Philox from here on
        opt._lib.mppi_generate_noise(opt._h, 0)
    opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 20, in_place=True)
    n0 = opt.launchCount()
    t = opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 500, in_place=True)
    print(f"{label:45s} e2e p50 {np.median(t) * 1e6:7.1f} us   launches/cycle {(opt.launchCount() - n0) / 500:.1f}")
    opt.shutdown()
