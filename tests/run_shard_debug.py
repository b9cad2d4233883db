"""Ad-hoc (torchrun N=2): one or two sharded cycles per mode, everything printed per rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
from navigation2_b200 import sharded, capi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
wl = bench.workload_c1()
for mode in ("nccl", "peer"):
    os.environ["MPPI_B200_NCCL_EXCHANGE"] = "1" if mode == "nccl" else "0"
    cfg = bench.large_batch_config(batch)
    cfg.shard_world, cfg.shard_rank = world, rank
    opt = sharded.ShardedOptimizer(cfg)
    opt.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
    for c in range(3):
        r = opt.evalControl(wl["inp"], raise_on_failure=False)
        print(f"[{mode} rank {rank} cycle {c}] status {r.status} retries {r.retries} fail {r.fail_flag} furthest {r.furthest_reached_path_point} "
              f"cmd {r.cmd[0]:.9f} {r.cmd[2]:.9f} u_sum {float(np.abs(r.control_sequence).sum()):.9f}", flush=True)
    dist.barrier()
    opt.shutdown()
dist.destroy_process_group()
