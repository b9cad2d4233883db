"""world_size-2 test of the batch-sharding exchanges on CPU (gloo): each rank owns half of the
trajectories of one oracle rollout and runs the host restatement of what kernels A / D2 do;
the all-reduce (MAX over encoded words) and all-gather (rescaled softmax slots) must reproduce
the unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import navigation2_b200 as nb
from navigation2_b200 import capi, params as P, scenarios as S
from navigation2_b200.sharded import combine_slots, decode_min, encode_min, local_slot
from oracle.binding import OracleOptimizer


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, costs, cv, per_traj_furthest, rep, inv_temp, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = costs.shape[0]
    lo, hi = rank * B // world, (rank + 1) * B // world
    # AR-1: MAX over (furthest index, encoded repulsion min, any-free flag)
    words = torch.tensor([int(per_traj_furthest[lo:hi].max()), int(encode_min(np.array([rep[lo:hi].min()]))[0]),
                          int((costs[lo:hi] < 1e5).any())], dtype=torch.int32)
    dist.all_reduce(words, op=dist.ReduceOp.MAX)
    # AR-2: all-gather of the local softmax slots
    slot = torch.from_numpy(local_slot(costs[lo:hi], cv[:, :, lo:hi], inv_temp))
    gathered = [torch.empty_like(slot) for _ in range(world)]
    dist.all_gather(gathered, slot)
    u = combine_slots(torch.stack(gathered).numpy(), inv_temp)
    if rank == 0:
        np.savez(out_path, words=words.numpy(), u=u)
    dist.barrier()
    dist.destroy_process_group()


def test_encode_min_roundtrip_and_order():
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.normal(0, 100, 1000), [0.0, -0.0, 1e-30, -1e-30, 3.4e38, -3.4e38]]).astype(np.float32)
    assert np.array_equal(decode_min(encode_min(v)), v)
    nz = v[~((v == 0) & np.signbit(v))]                          # -0.0 orders below +0.0 in the code, equal as floats
    order = np.argsort(nz, kind="stable")
    codes = encode_min(nz)[order]
    assert np.all(np.diff(codes.astype(np.int64)) <= 0)          # larger float -> smaller code
    assert decode_min(np.array([encode_min(v).max()]))[0] == v.min()


def test_two_rank_exchange_reproduces_unsharded_update(tmp_path):
    cfg = P.nav2_bringup_config(batch_size=256, materialize_tensors=1)
    o = OracleOptimizer(cfg)
    o.setCostmap(S.block_costmap(200, 200, 0.05, seed=2, keep_clear=(5.0, 5.0)), 0.05, 0.0, 0.0)
    path = S.arc_path((5.0, 5.0, 0.0), n_points=80)
    o.evalControl(nb.CycleInput(pose=(5.0, 5.0, 0.0), speed=(0.2, 0.0, 0.0), path=path))
    costs = o.getTensor(capi.TENSOR_COSTS)
    cv = np.stack([o.getTensor(capi.TENSOR_CVX), o.getTensor(capi.TENSOR_CVY), o.getTensor(capi.TENSOR_CWZ)])
    inv_temp = 1.0 / cfg.temperature
    rng = np.random.default_rng(1)
    furthest = rng.integers(0, 60, size=costs.shape[0])
    rep = rng.normal(0, 3, size=costs.shape[0]).astype(np.float32)
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(2, _free_port(), costs, cv, furthest, rep, inv_temp, out), nprocs=2, join=True)
    got = np.load(out)
    assert got["words"][0] == furthest.max()
    assert decode_min(got["words"][1:2])[0] == rep.min()
    assert got["words"][2] == int((costs < 1e5).any())
    # unsharded softmax-weighted mean, in double (optimizer.cpp:629-640)
    w = np.exp(-inv_temp * (costs.astype(np.float64) - costs.min()))
    expect = ((cv.astype(np.float64) * w[None, None, :]).sum(axis=2) / w.sum()).reshape(-1)
    np.testing.assert_allclose(got["u"], expect, rtol=0, atol=2e-6)
    # and it is what the oracle's own softmax weights give
    sw = o.getTensor(capi.TENSOR_SOFTMAX).astype(np.float64)
    expect2 = ((cv.astype(np.float64) * sw[None, None, :]).sum(axis=2) / sw.sum()).reshape(-1)
    np.testing.assert_allclose(got["u"], expect2, rtol=0, atol=2e-6)
