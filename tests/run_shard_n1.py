"""Ad-hoc (one GPU): mppi_shard_cycle timed inside the library, on the handle's own stream and on a torch stream,
next to the graph-replayed mppi_optimize of the same batch."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
import navigation2_b200 as nb
from navigation2_b200 import capi
wl = bench.workload_c1()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20)
for label in ("own stream", "torch stream"):
    cfg = bench.large_batch_config(batch)
    opt = nb.Optimizer(cfg)
    if label == "torch stream":
        st = torch.cuda.Stream()
        opt._lib.mppi_set_stream(opt._h, st.cuda_stream)
    opt.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
    c_in = (capi.MppiCycleIn * 1)(wl["inp"].to_c())
    c_out = (capi.MppiCycleOut * 1)()
    u = np.zeros((3, opt.T), dtype=np.float32); traj = np.zeros((3, opt.T), dtype=np.float32)
    fp = C.POINTER(C.c_float)
    c_out[0].control_vx = u[0].ctypes.data_as(fp); c_out[0].control_vy = u[1].ctypes.data_as(fp)
    c_out[0].control_wz = u[2].ctypes.data_as(fp); c_out[0].optimal_trajectory = traj.ctypes.data_as(fp)
    t = np.zeros(40)
    assert opt._lib.mppi_time_shard_cycles(opt._h, c_in, c_out, 5, t.ctypes.data_as(C.POINTER(C.c_double))) == 0
    assert opt._lib.mppi_time_shard_cycles(opt._h, c_in, c_out, 40, t.ctypes.data_as(C.POINTER(C.c_double))) == 0
    print(f"{label:14s} mppi_shard_cycle p50 {np.median(t) * 1e3:.3f} ms")
    if label == "own stream":
        t2 = opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 40, in_place=False)
        print(f"{'':14s} mppi_upload_costmap + mppi_optimize (graph replay) p50 {np.median(t2) * 1e3:.3f} ms")
    opt.shutdown()
