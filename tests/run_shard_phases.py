"""Ad-hoc (one GPU): host time of each mppi_shard_* phase call of a 2^20 x 56 cycle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ctypes as C
import torch
import bench
from navigation2_b200 import sharded, capi
wl = bench.workload_c1()
cfg = bench.large_batch_config(1 << 20)
opt = sharded.ShardedOptimizer(cfg)
opt.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
lib, h = opt._lib, opt._h
for _ in range(5):
    opt.evalControl(wl["inp"])
c_in = (capi.MppiCycleIn * 1)(wl["inp"].to_c())
c_out = (capi.MppiCycleOut * 1)()
u = np.zeros((3, opt.T), dtype=np.float32); traj = np.zeros((3, opt.T), dtype=np.float32)
fp = C.POINTER(C.c_float)
c_out[0].control_vx = u[0].ctypes.data_as(fp); c_out[0].control_vy = u[1].ctypes.data_as(fp)
c_out[0].control_wz = u[2].ctypes.data_as(fp); c_out[0].optimal_trajectory = traj.ctypes.data_as(fp)
acc = {}
n = 30
for _ in range(n):
    torch.cuda.synchronize()
    marks = [("start", time.perf_counter())]
    lib.mppi_shard_begin(h, c_in); marks.append(("begin", time.perf_counter()))
    lib.mppi_shard_rollout(h); marks.append(("rollout enqueue", time.perf_counter()))
    lib.mppi_shard_score(h); marks.append(("score enqueue", time.perf_counter()))
    opt._slot_all.copy_(opt._slot); marks.append(("slot copy", time.perf_counter()))
    lib.mppi_shard_update(h, 1); marks.append(("update enqueue", time.perf_counter()))
    lib.mppi_shard_end(h, c_out); marks.append(("end (sync + finish)", time.perf_counter()))
    for (a, ta), (b, tb) in zip(marks[:-1], marks[1:]):
        acc[b] = acc.get(b, 0.0) + (tb - ta)
print({k: round(v / n * 1e6, 1) for k, v in acc.items()}, "us; total", round(sum(acc.values()) / n * 1e6, 1))
