"""Ad-hoc (torchrun, N >= 2): the sharded cycle with exchanges over peer memory against the same cycle with
torch.distributed collectives between the kernels: identical results on every rank, and the time per cycle of both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
from navigation2_b200 import sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
batch = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20)
cycles = 6
wl = bench.workload_c1()
res = {}
for mode in ("nccl", "peer"):
    os.environ["MPPI_B200_NCCL_EXCHANGE"] = "1" if mode == "nccl" else "0"
    cfg = bench.large_batch_config(batch)
    cfg.shard_world, cfg.shard_rank = world, rank
    opt = sharded.ShardedOptimizer(cfg)
    opt.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
    u = np.stack([opt.evalControl(wl["inp"]).control_sequence.copy() for _ in range(cycles)])
    t = torch.tensor(u, device="cuda")
    allu = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allu, t)
    ranks_agree = all(torch.equal(allu[0], a) for a in allu)
    dist.barrier(); torch.cuda.synchronize()
    n = 50
    t0 = time.perf_counter()
    for _ in range(n):
        opt.evalControl(wl["inp"])
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / n
    lib_dt = float(np.median(opt.timeCycles(wl["inp"], n))) if mode == "peer" else float("nan")
    res[mode] = (u, dt, ranks_agree, lib_dt)
    dist.barrier()
    opt.shutdown()
if rank == 0:
    print(f"world {world} batch {batch}")
    for mode in res:
        same = np.array_equal(res[mode][0], res["nccl"][0])
        print(f"  {mode:5s} ranks agree: {res[mode][2]}  equals nccl result: {same}  {res[mode][1] * 1e3:.3f} ms/cycle through Python"
              f"  {res[mode][3] * 1e3:.3f} ms/cycle inside the library (p50)")
dist.destroy_process_group()
