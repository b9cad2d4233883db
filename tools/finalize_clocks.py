"""Ad-hoc: SM-clock breakdown of k_finalize (the general path's kernel D) at a given batch (MPPI_B200_PHASE_CLOCKS=1)."""
import os, sys, ctypes as C
os.environ["MPPI_B200_PHASE_CLOCKS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from navigation2_b200 import capi
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
wl = bench.workload_c1()
cfg = bench.large_batch_config(batch)
opt = bench.setup_engine(wl, cfg=cfg)
for _ in range(15):
    opt.evalControl(wl["inp"])
out = np.zeros(16 * 16 * 32, dtype=np.int64)
rc = opt._lib.mppi_get_tensor(opt._h, 0, capi.TENSOR_PHASE_CLOCKS, out.ctypes.data_as(C.c_void_p), out.nbytes)
assert rc == 0, rc
row = out.reshape(16, 16, 32)[0, 0]
names = {0: "kernel start", 1: "params staged + partials reduced", 2: "mean control sequence ready", 3: "window staged",
         22: "tail: start", 23: "tail: filter", 24: "tail: constraint chains (+ heading)", 25: "tail: trajectory integrated",
         26: "tail: validator", 27: "tail: header published", 28: "tail: state committed", 14: "kernel end"}
t0, prev = row[0], row[0]
for slot in (0, 1, 2, 3, 22, 23, 24, 25, 26, 27, 28, 14):
    if row[slot]:
        print(f"{names[slot]:40s} {int(row[slot] - t0):8d}  [+{int(row[slot] - prev)}]")
        prev = row[slot]
