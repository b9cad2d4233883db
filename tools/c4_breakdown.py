"""Where a large-batch cycle spends its time on one GPU: per-kernel CUDA-event times of kernels A and C in live cycles
(first cycle from a zero control sequence vs steady state), the resident replay, and the whole mppi_shard_cycle call.
  python tools/c4_breakdown.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402
from navigation2_b200.sharded import ShardedOptimizer  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20)
wl = bench.workload_c1()
cfg = bench.large_batch_config(batch)
opt = ShardedOptimizer(cfg)
opt.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
inp = wl["inp"]


def kernels(label, n=10):
    opt.setKernelTiming(True)
    t = opt.timeCycles(inp, n)
    a, na = opt.kernelTimeMs()
    c, nc = opt.kernelCTimeMs()
    opt.setKernelTiming(False)
    print(f"{label}: kernel A {a * 1e3:.1f} us ({na} launches)  kernel C {c * 1e3:.1f} us  whole call (ungraphed, with events) "
          f"{np.median(t) * 1e6:.1f} us", flush=True)


opt.evalControl(inp)
opt.reset()
opt.setKernelTiming(True)
opt.evalControl(inp)
a, _ = opt.kernelTimeMs()
c, _ = opt.kernelCTimeMs()
opt.setKernelTiming(False)
print(f"first cycle after reset: kernel A {a * 1e3:.1f} us  kernel C {c * 1e3:.1f} us")
for _ in range(12):
    opt.evalControl(inp)
kernels("steady state")
t = opt.timeCycles(inp, 50)
print(f"steady state, graphed cycle: p50 {np.median(t) * 1e6:.1f} us  mean {np.mean(t) * 1e6:.1f} us")
opt.setResidentCapture(True)
opt.evalControl(inp)
for flush in (0, 256 << 20):
    ms = opt.timeResident(20, 3, flush)
    print(f"resident replay of a steady-state cycle, flush {flush >> 20} MiB: mean {np.mean(ms) * 1e3:.1f} us")
opt.shutdown()
