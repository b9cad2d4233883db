"""How a C3 control cycle's device time evolves as the control sequence converges (robot held in place):
kernel A / kernel C CUDA-event times and the whole call, every 100 cycles.  python tools/c3_breakdown.py [model]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402

wl = bench.workload_c3(sys.argv[1] if len(sys.argv) > 1 else "omni")
opt = bench.setup_engine(wl)
inp = wl["inp"]
maps = (wl["cells"], wl["resolution"], wl["origin"][0], wl["origin"][1])
for block in range(8):
    for _ in range(100 if block else 12):
        r = opt.evalControl(inp, costmaps=maps)
    opt.setKernelTiming(True)
    for _ in range(5):
        opt.evalControl(inp, costmaps=maps)
    a, _ = opt.kernelTimeMs()
    c, _ = opt.kernelCTimeMs()
    opt.setKernelTiming(False)
    t = opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], inp, 20, in_place=True)
    print(f"after ~{12 + 125 * block:4d} cycles: kernel A {a * 1e3:7.1f} us  kernel C {c * 1e3:6.1f} us  whole call p50 {np.median(t) * 1e6:7.1f} us  "
          f"cmd {r.cmd[0]:.3f} {r.cmd[1]:.3f} {r.cmd[2]:.3f} retries {r.retries} |u|max {np.abs(r.control_sequence).max():.3f}", flush=True)
opt.shutdown()
opt = bench.setup_engine(wl)
for _ in range(12):
    opt.evalControl(inp)
opt.setResidentCapture(True)
opt.evalControl(inp)
opt.setKernelTiming(True)
ms = opt.timeResident(10, 3, 0)
a, na = opt.kernelTimeMs()
c, _ = opt.kernelCTimeMs()
opt.setKernelTiming(False)
print(f"resident replay: whole {np.mean(ms) * 1e3:.1f} us  kernel A {a * 1e3:.1f} us ({na})  kernel C {c * 1e3:.1f} us")
opt.setResidentCapture(False)
opt.setKernelTiming(True)
for _ in range(5):
    opt.evalControl(inp)
a, na = opt.kernelTimeMs()
opt.setKernelTiming(False)
print(f"live (staged map, same handle): kernel A {a * 1e3:.1f} us ({na})")
opt.shutdown()
