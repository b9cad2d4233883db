"""torchrun worker behind tests/test_gpu_sharded.py: the batch-sharded optimiser on WORLD_SIZE GPUs against the
unsharded optimiser on one GPU and against the CPU oracle (reference semantics
src/optimizer.cpp:605-645).  Every rank drives its shard through mppi_shard_cycle (exchanges over peer memory,
k_shard_exchange) and, with --nccl, through the mppi_shard_* phases with torch.distributed collectives between them.
Rank 0 additionally runs the same global batch on a plain handle and on the oracle, fed the GPU's own Philox tensors.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_parity_worker.py [batch]

Prints `SHARD_PARITY_OK ...` on rank 0 when everything agrees; any mismatch raises on the rank that saw it.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import navigation2_b200 as nb  # noqa: E402
from navigation2_b200 import capi, params as P, scenarios as S, sharded  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 20)
    cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    cells = S.block_costmap(400, 400, 0.05, seed=1, keep_clear=(10.0, 10.0))
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)

    engines = {}
    for mode in ("peer", "nccl"):
        os.environ["MPPI_B200_NCCL_EXCHANGE"] = "1" if mode == "nccl" else "0"
        cfg = P.nav2_bringup_config(batch_size=batch, shard_world=world, shard_rank=rank)
        engines[mode] = sharded.ShardedOptimizer(cfg)
        engines[mode].setCostmap(cells, 0.05, 0.0, 0.0)
    os.environ.pop("MPPI_B200_NCCL_EXCHANGE")
    assert engines["peer"]._peer_exchange and not engines["nccl"]._peer_exchange

    plain = cpu = None
    if rank == 0:
        from oracle.binding import OracleOptimizer
        from util import assert_costs_close, assert_cycle_close
        cfg1 = P.nav2_bringup_config(batch_size=batch)
        plain = nb.Optimizer(cfg1)
        cpu = OracleOptimizer(cfg1)
        noise = [plain.getTensor(t) for t in (capi.TENSOR_NOISE_VX, capi.TENSOR_NOISE_VY, capi.TENSOR_NOISE_WZ)]
        # Philox counters are keyed by the global trajectory index: this rank's shard is a column block of the global tensor
        mine = engines["peer"].getTensor(capi.TENSOR_NOISE_VX)
        assert np.array_equal(mine, noise[0][:, : batch // world]), "shard noise is not the global tensor's column block"
        cpu.setNoise(*noise)
        plain.setNoise(*noise)
        for e in (plain, cpu):
            e.setCostmap(cells, 0.05, 0.0, 0.0)
        del noise

    # rank 0 has just spent seconds on the oracle's set-up: an exchange waits 2 s for its peers, so start together
    dist.barrier()
    pose, speed = [10.0, 10.0, 0.0], (0.2, 0.0, 0.0)
    worst = 0.0
    for cycle in range(cycles):
        inp = nb.CycleInput(pose=tuple(pose), speed=speed, path=path, goal=tuple(path[-1]), goal_checker_xy_tolerance=0.25)
        res = {mode: eng.evalControl(inp) for mode, eng in engines.items()}
        # peer-memory exchange and NCCL exchange run the same kernels on the same payloads
        assert np.array_equal(res["peer"].control_sequence, res["nccl"].control_sequence), f"peer vs nccl, cycle {cycle}"
        # every rank ends a cycle with the same control sequence (kernel D2 runs redundantly on all ranks)
        t = torch.tensor(res["peer"].control_sequence, device="cuda")
        allu = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allu, t)
        assert all(torch.equal(allu[0], a) for a in allu), f"ranks disagree in cycle {cycle}"
        state = [None]
        if rank == 0:
            c = cpu.evalControl(inp)
            g = plain.evalControl(inp)
            assert_cycle_close(g, c, label=f"unsharded vs oracle, cycle {cycle}")
            for mode in engines:
                assert_cycle_close(res[mode], c, label=f"sharded x{world} ({mode}) vs oracle, cycle {cycle}")
                # sharded vs unsharded: same kernels, different reduction tree for the softmax sums only
                np.testing.assert_allclose(res[mode].control_sequence, g.control_sequence, rtol=0, atol=2e-6,
                                           err_msg=f"sharded ({mode}) vs unsharded, cycle {cycle}")
                worst = max(worst, float(np.max(np.abs(res[mode].control_sequence - c.control_sequence))))
            lo_costs = cpu.getTensor(capi.TENSOR_COSTS)[: batch // world]
            assert_costs_close(engines["peer"].getTensor(capi.TENSOR_COSTS), lo_costs, label=f"shard 0 costs, cycle {cycle}")
            st = cpu.getState()
            state = [(bytes(st), cpu.getOptimalControlSequence(), c.cmd)]
            plain.setState(st)
            plain.setControlSequence(state[0][1])
        dist.broadcast_object_list(state, src=0)
        st = capi.MppiOptimizerState.from_buffer_copy(state[0][0])
        for eng in engines.values():
            eng.setState(st)
            eng.setControlSequence(state[0][1])
        cmd = state[0][2]
        pose[0] += 0.05 * cmd[0] * np.cos(pose[2]); pose[1] += 0.05 * cmd[0] * np.sin(pose[2]); pose[2] += 0.05 * cmd[2]
        speed = (cmd[0], 0.0, cmd[2])

    # dynamic reconfiguration of a live sharded handle: weights change, shapes stay
    cfg2 = P.nav2_bringup_config(batch_size=batch, shard_world=world, shard_rank=rank, temperature=0.25)
    inp = nb.CycleInput(pose=tuple(pose), speed=speed, path=path, goal=tuple(path[-1]), goal_checker_xy_tolerance=0.25)
    after = {}
    dist.barrier()
    for mode, eng in engines.items():
        eng.setConfig(cfg2)
        after[mode] = eng.evalControl(inp)
    assert np.array_equal(after["peer"].control_sequence, after["nccl"].control_sequence), "peer vs nccl after setConfig"
    if rank == 0:
        from oracle.binding import OracleOptimizer
        cfg2g = P.nav2_bringup_config(batch_size=batch, temperature=0.25)
        cpu2 = OracleOptimizer(cfg2g)
        plain.setConfig(cfg2g)           # same seed, epoch advanced identically on every handle by the reset
        noise = [plain.getTensor(t) for t in (capi.TENSOR_NOISE_VX, capi.TENSOR_NOISE_VY, capi.TENSOR_NOISE_WZ)]
        assert np.array_equal(engines["peer"].getTensor(capi.TENSOR_NOISE_VX), noise[0][:, : batch // world])
        cpu2.setNoise(*noise)
        cpu2.setCostmap(cells, 0.05, 0.0, 0.0)
        c = cpu2.evalControl(inp)
        assert_cycle_close(after["peer"], c, label="after setConfig")
        cpu2.shutdown()
        print(f"SHARD_PARITY_OK world {world} batch {batch} cycles {cycles} worst |u - oracle| {worst:.2e}", flush=True)
    dist.barrier()
    for eng in engines.values():
        eng.shutdown()
    if rank == 0:
        plain.shutdown(); cpu.shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
