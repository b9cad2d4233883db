"""Ad-hoc: the 2^20 x 56 batch (BASELINE.json configs[3] on one GPU) for ncu captures and quick timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
batch = int(sys.argv[2]) if len(sys.argv) > 2 else (1 << 20)
wl = bench.workload_c1()
cfg = bench.large_batch_config(batch)
opt = bench.setup_engine(wl, cfg=cfg)
if os.environ.get("LARGE_BATCH_STEADY", "1") == "1":     # converged control sequence: PathAlign active, like bench.py
    for _ in range(bench.STEADY_STATE_CYCLES):
        opt.evalControl(wl["inp"])
opt.setResidentCapture(True)
opt.evalControl(wl["inp"])
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
opt.setKernelTiming(True)
ms = bench.time_resident(opt, steps, 2, flush, stream)
ka, n = opt.kernelTimeMs()
alg = bench.algorithmic_bytes_kernel_a(cfg, 0)
print(f"step_ms mean {ms.mean():.4f}  kernelA_ms {ka:.4f} ({n} launches)  kernelA GB/s {alg / ka / 1e6:.1f}  steps/s {batch * 56 / ms.mean() * 1e3:.3e}")

# the scoring pass alone on materialised tensors (the reference's critics' inputs), same batch
if len(sys.argv) > 3 and sys.argv[3] == "score":
    opt.shutdown()
    cfg2 = bench.large_batch_config(batch)
    cfg2.materialize_tensors = 1
    opt2 = bench.setup_engine(wl, cfg=cfg2)
    opt2.setResidentCapture(True)
    opt2.evalControl(wl["inp"])
    opt2.setKernelTiming(True)
    for i in range(steps + 2):
        flush.zero_()
        opt2.enqueueScoreResident(stream.cuda_stream)
        if i == 1:
            torch.cuda.synchronize(); opt2.kernelTimeMs(reset=True)
    torch.cuda.synchronize()
    ks, n = opt2.kernelTimeMs()
    T = 56
    alg12 = batch * T * 12 + batch * 4 + 5 * batch   # x, y, vx all columns + yaw last column + cost/flag
    print(f"scoring pass kernelA(from tensors)_ms {ks:.4f} ({n} launches)  GB/s at 12 B/step {alg12 / ks / 1e6:.1f}  frac {alg12 / ks / 1e6 / 6535.4:.3f}")
