"""Ad-hoc: per-phase SM clocks of the fused kernel with a hot and with a flushed L2 (where do the cold microseconds go)."""
import os, sys, ctypes as C
os.environ["MPPI_B200_PHASE_CLOCKS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from navigation2_b200 import capi
wl = bench.workload_c1()
opt = bench.setup_engine(wl)
opt.setResidentCapture(True)
opt.evalControl(wl["inp"])
order = [(0, "start"), (17, "noise copies issued"), (18, "param copies issued"), (19, "window issued + landed"),
         (1, "sync; u0 fix"), (3, "P1 velocity chains"), (4, "P2 sincos"), (5, "P3 chains"), (15, "P3b cell loop"),
         (21, "P4 scans"), (6, "furthest"), (7, "reduce + cluster.sync"), (8, "block_b"), (9, "P5 path critics"),
         (10, "cluster.sync"), (11, "P6 softmax + sums"), (12, "cluster.sync"), (13, "combine"), (14, "finalize end")]
res = {}
for label, flush in (("hot", 0), ("cold", 256 << 20), ("hot2", 0), ("cold2", 256 << 20)):
    ms = opt.timeResident(3, 2, flush)
    out = np.zeros(16 * 16 * 32, dtype=np.int64)
    rc = opt._lib.mppi_get_tensor(opt._h, 0, capi.TENSOR_PHASE_CLOCKS, out.ctypes.data_as(C.c_void_p), out.nbytes)
    clk = out.reshape(16, 16, 32)
    t0 = clk[:, :, 0].min()
    res[label] = ([int(clk[0, 0, s] - t0) for s, _ in order], float(np.median(ms)) * 1e3, int(clk[0, 1, 31] - clk[0, 1, 30]))
print(f"{'phase (rank 0 warp 0 arrival, cycles)':40s}" + "".join(f"{k:>10s}" for k in res))
for i, (_, name) in enumerate(order):
    print(f"{name:40s}" + "".join(f"{res[k][0][i]:10d}" for k in res))
print(f"{'event-timed step (us)':40s}" + "".join(f"{res[k][1]:10.1f}" for k in res))
print(f"{'in-kernel wall (ns)':40s}" + "".join(f"{res[k][2]:10d}" for k in res))
