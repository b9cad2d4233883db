"""Ad-hoc: per-phase SM-clock breakdown of the fused kernel at C1 (MPPI_B200_PHASE_CLOCKS=1)."""
import os, sys, ctypes as C
os.environ["MPPI_B200_PHASE_CLOCKS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from navigation2_b200 import capi
wl = bench.workload_c1()
opt = bench.setup_engine(wl)
for _ in range(5):
    opt.evalControl(wl["inp"])
out = np.zeros(16 * 32, dtype=np.int64)
rc = opt._lib.mppi_get_tensor(opt._h, 0, capi.TENSOR_PHASE_CLOCKS, out.ctypes.data_as(C.c_void_p), out.nbytes)
assert rc == 0, rc
clk = out.reshape(16, 32)
names = ["start", "issued loads", "loads landed", "P1 vel chains", "P2 sincos", "P3 xy chains", "P4 critics", "xch1 sync",
         "block_b", "P5 path+cost", "xch2 sync", "P6 softmax+wsum", "xch3 sync", "combine+sync", "finalize"]
c = clk[0]
print("fine marks rank 0 (relative to start):", {k: int(c[k] - c[0]) for k in range(16, 22) if c[k]}, "mark1", int(c[1]-c[0]), "mark5", int(c[5]-c[0]), "mark6", int(c[6]-c[0]))
print("finalize marks (delta cycles):", [int(c[k] - c[k-1]) for k in range(23, 32) if c[k] and c[k-1]], "first", int(c[22] - c[13]) if c[22] else None)
for rank in (0,):
    c = clk[rank]
    print("rank", rank, "total cycles", c[13] - c[0], "(+finalize", c[14] - c[13] if c[14] else 0, ")")
    for i in range(1, 15):
        if c[i]:
            print(f"   {names[i]:18s} {c[i] - c[i-1]:8d} cyc")
