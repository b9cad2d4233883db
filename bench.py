#!/usr/bin/env python
"""bench.py — MPPI trajectory-steps scored per second (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c3|c4|c5]

A *step* is one Optimizer::optimize() (iteration_count iterations of rollout -> score -> update)
over one batch of synthetic input.  Headline workload: BASELINE.json configs[1], DiffDrive 2000 x 56,
model_dt 0.05, default critics with the bring-up yaml weights, 400x400 costmap @0.05 m, seeded injected
noise.  N>1 (torchrun, one rank per GPU): every rank drives its own independent controller on the same
workload ("fleets of independent controllers are partitioned one robot-set per GPU", no data-path
collective, weak scaling).

The same JSON line carries, at EVERY N, the two multi-GPU workloads north_star names:
  c4_sharded  BASELINE.json configs[3]: ONE 2^20 x 56 batch sharded along the trajectory axis over the N ranks
              (strong scaling; the two exchanges of an iteration run over peer memory inside the kernels),
              whole mppi_shard_cycle calls with host buffers, max over ranks, parity-checked against the oracle
  c5_fleet    BASELINE.json configs[4]: 512 robots per GPU, own 200x200 costmap + path each (weak scaling),
              device-timed and end to end, parity-checked against the oracle
`--workload c3|c4|c5` makes one of them the headline line instead (development runs).

Printed JSON (one line, rank 0):
  value      trajectory-steps/s, inputs resident in HBM, CUDA-event timed, L2 flushed between steps
  e2e        the same metric through the C ABI (mppi_optimize_in_place) with HOST buffers: costmap window +
             cycle inputs H2D from pinned memory and the command / control sequence D2H inside the timed region
  roofline   dominant kernel against the measured HBM peak
  cpu_baseline  the CPU oracle (a restatement, `kind: port`) on this box's host cores: 1 thread like the
             reference's control path, plus an all-cores figure (one independent controller per core)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "MPPI trajectory-steps scored/sec at batch 2000\u00d756; p50 optimize() latency"   # BASELINE.json's metric, character for character
UNIT = "trajectory-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1", choices=["c1", "c3", "c4", "c5"])
    ap.add_argument("--robots", type=int, default=512, help="robots per GPU for --workload c5")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large-batch", action="store_true", help="skip roofline_large_batch / c4_sharded / c5_fleet")
    return ap.parse_args()


# ----------------------------------------------------------------------------- workloads

def workload_c1():
    """BASELINE.json configs[0]/[1]."""
    from navigation2_b200 import params as P, scenarios as S
    import navigation2_b200 as nb
    cfg = P.nav2_bringup_config()
    cells = S.block_costmap(400, 400, 0.05, seed=1, keep_clear=(10.0, 10.0))
    path = S.arc_path((10.0, 10.0, 0.0), n_points=100)
    inp = nb.CycleInput(pose=(10.0, 10.0, 0.0), speed=(0.2, 0.0, 0.0), path=path, goal=tuple(path[-1]),
                        goal_checker_xy_tolerance=0.25)
    return dict(name="C1 DiffDrive 2000x56, default critics (yaml weights), 400x400 costmap @0.05 m, injected seeded noise",
                cfg=cfg, cells=cells, resolution=0.05, origin=(0.0, 0.0), inp=inp, holo=False)


def workload_c3(model="omni"):
    """BASELINE.json configs[2]: Omni (or Ackermann) 16384 x 100, full critic set incl. Obstacles + inflation."""
    from navigation2_b200 import params as P, scenarios as S
    import navigation2_b200 as nb
    critics = ["ConstraintCritic", "CostCritic", "GoalCritic", "GoalAngleCritic", "ObstaclesCritic", "PathAlignCritic",
               "PathFollowCritic", "PathAngleCritic", "PreferForwardCritic", "TwirlingCritic", "VelocityDeadbandCritic"]
    cfg = P.make_config({"batch_size": 16384, "time_steps": 100, "motion_model": model, "controller_frequency": 20.0},
                        critics, P.NAV2_BRINGUP_CRITIC_PARAMS, robot_radius=0.22,
                        inflation={"cost_scaling_factor": 3.0, "inflation_radius": 0.70})
    cells = S.block_costmap(400, 400, 0.05, seed=1, keep_clear=(10.0, 10.0))
    path = S.arc_path((10.0, 10.0, 0.0), n_points=140)
    inp = nb.CycleInput(pose=(10.0, 10.0, 0.0), speed=(0.2, 0.0, 0.0), path=path, goal=tuple(path[-1]),
                        goal_checker_xy_tolerance=0.25)
    return dict(name=f"C3 {model} 16384x100, full critic set incl. Obstacles+inflation, 400x400 costmap",
                cfg=cfg, cells=cells, resolution=0.05, origin=(0.0, 0.0), inp=inp, holo=model == "omni")


def workload_reference_harness():
    """The reference's own harness (benchmark/optimizer_benchmark.cpp:49-99): DiffDrive 2000 x 56, iteration_count 2,
    40x40 costmap @0.1 m with one 8x8 block of cost 250, 50 path poses stepping (+0.1, +0.1)."""
    from navigation2_b200 import params as P, scenarios as S
    import navigation2_b200 as nb
    cells, res, origin, path = S.reference_benchmark_fixture()
    cfg = P.make_config({"batch_size": 2000, "time_steps": 56, "iteration_count": 2, "motion_model": "diff_drive"},
                        P.NAV2_BRINGUP_CRITICS, robot_radius=0.15)
    inp = nb.CycleInput(pose=(2.0, 2.0, 0.785), speed=(0.0, 0.0, 0.0), path=path, goal=tuple(path[-1]))
    return dict(name="reference harness: DiffDrive 2000x56, iteration_count 2, 40x40 costmap @0.1 m (optimizer_benchmark.cpp)",
                cfg=cfg, cells=cells, resolution=res, origin=origin, inp=inp, holo=False)


def bench_config(wl, cfg):
    """The `config` object of the JSON line.  Both arms build it here from the workload alone, so the driver's
    same_config check compares like with like; arm-specific detail (L2 handling, parallelism) is spelled out for both
    arms in the same strings."""
    return {"workload": wl["name"], "batch_size": int(cfg.batch_size), "time_steps": int(cfg.time_steps),
            "iteration_count": int(cfg.iteration_count),
            "l2": "GPU arm: L2 flushed before every timed step (a 256 MiB buffer overwritten, mppi_time_resident); "
                  "reference arm: CPU, caches as they fall",
            "parallelism": "GPU arm: one independent controller per GPU, no collective (weak scaling), plus the c4_sharded / "
                           "c5_fleet sections; reference arm: rank 0 only, single thread like the reference's control path"}


def large_batch_config(batch, world=1, rank=0):
    from navigation2_b200 import params as P
    return P.nav2_bringup_config(batch_size=batch, shard_world=world, shard_rank=rank)


# ----------------------------------------------------------------------------- helpers

class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks and throttle reasons DURING the timed region."""

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture, or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):   # the latest committed capture that has the kernel
        try:
            e = json.load(open(os.path.join(ROOT, "profiles", name)))[key]
            return int(e["dram_bytes_read"]) + int(e["dram_bytes_write"])
        except Exception:
            continue
    return None


def setup_engine(wl, cfg=None, seed=42):
    import navigation2_b200 as nb
    from navigation2_b200 import scenarios as S
    cfg = cfg or wl["cfg"]
    opt = nb.Optimizer(cfg)  # raises if the CUDA library / device is missing: no fallback
    R = int(cfg.n_robots)
    B = int(cfg.batch_size) // int(cfg.shard_world)
    for r in range(R):
        opt.setCostmap(wl["cells"], wl["resolution"], wl["origin"][0], wl["origin"][1], robot=r)
    if B * int(cfg.time_steps) * R <= (1 << 24):
        for r in range(R):
            nvx, nvy, nwz = S.seeded_noise(B, int(cfg.time_steps), (cfg.vx_std, cfg.vy_std, cfg.wz_std),
                                           seed=seed + r, holonomic=wl["holo"])
            opt.setNoise(nvx, nvy, nwz, robot=r)
    # else: keep the Philox noise the library drew at reset (seed + global trajectory index)
    return opt


def algorithmic_bytes_kernel_a(cfg, win_bytes):
    """noise columns read once (4 B x axes per trajectory-step) + 5 B per trajectory
    (cost + collision flag) + the staged costmap window once."""
    from navigation2_b200 import capi
    axes = 3 if cfg.motion_model == capi.MPPI_MODEL_OMNI else 2
    B = int(cfg.batch_size) // int(cfg.shard_world)
    R = int(cfg.n_robots)
    return R * (B * int(cfg.time_steps) * 4 * axes + 5 * B + win_bytes)


FLUSH_BYTES = 256 << 20   # > 126 MB of L2
# control cycles run before anything is timed or captured (converged control sequence: every default critic active);
# MPPI_BENCH_STEADY_CYCLES=0 reproduces round 1's measurement (first cycles after a reset, PathAlign not yet active)
STEADY_STATE_CYCLES = int(os.environ.get("MPPI_BENCH_STEADY_CYCLES", "12"))


def time_resident(opt, steps, warmup, flush, stream):
    """K resident steps, each bracketed by CUDA events on the launching stream (inside the library:
    mppi_time_resident), L2 flushed before every step when `flush` is not None."""
    return opt.timeResident(steps, warmup, FLUSH_BYTES if flush is not None else 0)


def time_resident_python(opt, steps, warmup, flush, stream):
    """The same loop driven from Python with torch events (kept for cross-checking the in-library loop)."""
    import torch
    for _ in range(warmup):
        if flush is not None:
            flush.zero_()
        opt.enqueueResident(stream.cuda_stream)
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    stops = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    for i in range(steps):
        if flush is not None:
            flush.zero_()  # > L2 (126 MB): the next step starts cold
        starts[i].record(stream)
        opt.enqueueResident(stream.cuda_stream)
        stops[i].record(stream)
    torch.cuda.synchronize()
    return np.array([s.elapsed_time(e) for s, e in zip(starts, stops)])  # ms


# ----------------------------------------------------------------------------- reference arm

def run_reference(args):
    """The reference's own CPU implementation of the path cannot be built here (no ROS 2 / Eigen);
    this arm times the oracle port (oracle/, -O3 -mavx2 -mfma, the reference's flags) on the host
    cores.  The reference control path is single-threaded, so is this.  One "step" of this arm is a
    bounded SAMPLE of the workload: as many back-to-back evalControl cycles as fit in ~0.25 s
    (calibrated once), so a short --steps run still times seconds of CPU work, not milliseconds.
    Nothing here maps the product library (navigation2_b200.params builds the configuration through
    the host-only libmppi_config.so)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.binding import OracleOptimizer
    from navigation2_b200 import scenarios as S
    wl = workload_c1() if args.workload in ("c1", "c4", "c5") else workload_c3()
    cfg = wl["cfg"]
    o = OracleOptimizer(cfg, fast=True)
    o.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
    nvx, nvy, nwz = S.seeded_noise(int(cfg.batch_size), int(cfg.time_steps), (cfg.vx_std, cfg.vy_std, cfg.wz_std),
                                   holonomic=wl["holo"])
    o.setNoise(nvx, nvy, nwz)
    units = int(cfg.batch_size) * int(cfg.time_steps) * int(cfg.iteration_count)
    for _ in range(max(args.warmup, 3)):
        o.evalControl(wl["inp"])
    one = o.timeOptimize(wl["inp"], 5)                        # seconds per cycle
    per_step = int(min(max(round(0.25 / max(one, 1e-6)), 1), 1000))
    n_steps = min(args.steps, 400)                            # bounded: <= ~100 s of CPU work
    singles = []
    for _ in range(min(50, max(10, per_step))):                # single-call latencies for the p50
        t0 = time.perf_counter()
        o.evalControl(wl["inp"])
        singles.append(time.perf_counter() - t0)
    times = np.array([o.timeOptimize(wl["inp"], per_step) * per_step for _ in range(n_steps)])   # seconds per sample
    total = float(times.sum())
    value = units * per_step * n_steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": n_steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / n_steps, "cycles_per_step": per_step,
        "ms_per_optimize": 1e3 * total / (n_steps * per_step), "p50_ms": float(np.median(singles) * 1e3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(wl, cfg),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{n_steps} samples of {per_step} back-to-back evalControl cycles of the workload (oracle port, "
                                   "g++ -O3 -mavx2 -mfma, 1 thread; the reference's Eigen build needs ROS 2 and cannot be compiled here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host_cpu": cpu_model(), "host_cores": os.cpu_count(),
    }
    emit(line)


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ----------------------------------------------------------------------------- our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    from navigation2_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # the ranks of one box share its host cores: each library instance packs its fleet's inputs on its share of them
    # (torchrun exports OMP_NUM_THREADS=1; the library sizes its parallel regions from this variable instead)
    os.environ.setdefault("MPPI_B200_HOST_THREADS", str(max(1, (os.cpu_count() or 1) // world)))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the optimiser has no CPU path to measure")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if args.workload == "c4":
        return run_c4(args, world, rank, local_rank)

    wl = workload_c1() if args.workload in ("c1", "c5") else workload_c3()
    cfg = wl["cfg"]
    if args.workload == "c5":
        from navigation2_b200 import params as P, scenarios as S
        cfg = P.nav2_bringup_config(n_robots=args.robots)
        wl = dict(wl)
        wl["cfg"] = cfg
        wl["cells"] = S.block_costmap(200, 200, 0.05, seed=1 + rank, keep_clear=(5.0, 5.0))
        import navigation2_b200 as nb
        path = S.arc_path((5.0, 5.0, 0.0), n_points=100)
        wl["inp"] = nb.CycleInput(pose=(5.0, 5.0, 0.0), speed=(0.2, 0.0, 0.0), path=path, goal=tuple(path[-1]),
                                  goal_checker_xy_tolerance=0.25)
        wl["name"] = f"C5 fleet: {args.robots} independent robots per GPU, own 200x200 costmap + path each, batch 2000x56"
    R = int(cfg.n_robots)
    B, T, iters = int(cfg.batch_size), int(cfg.time_steps), int(cfg.iteration_count)
    units_per_step = R * B * T * iters

    opt = setup_engine(wl, seed=42 + 1000 * rank)
    inputs = wl["inp"] if R == 1 else [wl["inp"]] * R
    # a real (non-default) stream: the library treats a NULL stream argument as "use my own", and
    # CUDA events only see the stream they are recorded on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    flush = True  # L2 flushed before every timed step (mppi_time_resident, FLUSH_BYTES)

    # Cycles through the public API until the control sequence has converged (the robot stands still, the sequence
    # settles on driving along the path), then one more whose inputs the resident path replays: `value` times a
    # steady-state optimize(), like the reference arm's closed loop, not the first cycle from a zero sequence.
    for _ in range(STEADY_STATE_CYCLES):
        opt.evalControl(inputs)
    opt.setResidentCapture(True)
    opt.evalControl(inputs)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- value: resident inputs, CUDA events on the launching stream, L2 flushed between steps ----
    launches0 = opt.launchCount()
    opt.setKernelTiming(False)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    step_ms = time_resident(opt, args.steps, args.warmup, flush, stream)
    if distributed:
        dist.barrier()
    torch.cuda.synchronize()
    launches = opt.launchCount() - launches0
    launches_timed = launches * args.steps // (args.steps + args.warmup)
    total_ms = float(step_ms.sum())
    if distributed:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * units_per_step * args.steps / (total_ms * 1e-3)

    # hot-L2 variant (no flush) for context: what a 20 Hz controller loop actually sees
    hot_ms = time_resident(opt, args.steps, 3, None, stream)

    # ---- dominant kernel (rollout+score) timed alone with CUDA events inside the library ----
    opt.setKernelTiming(True)
    time_resident(opt, min(args.steps, 50), 3, flush, stream)
    ka_ms, ka_n = opt.kernelTimeMs(reset=True)
    opt.setKernelTiming(False)

    # ---- e2e: the C ABI with HOST buffers, timed inside the library with steady_clock per call (mppi_time_e2e),
    # i.e. what a C++ controller_server thread sees.  Primary: mppi_optimize_in_place, the reference's own shape of
    # the call (evalControl reads the costmap where it lies, under the costmap mutex): pack the costmap window +
    # cycle inputs, ONE H2D copy, ONE kernel, result in mapped pinned memory, host sync.  Also reported: the
    # two-call sequence (mppi_upload_costmap stages the whole map first) and the same loop through Python. ----
    opt.setResidentCapture(False)

    def timed_e2e(in_place):
        opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], inputs, args.warmup, in_place=in_place)
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()
        b0 = opt.h2dByteCount()
        t = opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], inputs, args.steps, in_place=in_place)
        torch.cuda.synchronize()
        return t, (opt.h2dByteCount() - b0) / args.steps

    # the e2e loops start from a device and a host that have been running control cycles for a while (clocks, caches,
    # page tables settled): a controller runs continuously, the first ~100 calls after a pause are 15 % slower
    t_settle = time.perf_counter()
    while time.perf_counter() - t_settle < 0.3:
        opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], inputs, 50, in_place=True)
    two_call_times, two_call_h2d = timed_e2e(False)
    e2e_times, h2d = timed_e2e(True)
    e2e_status = getattr(opt, "last_timed_status", None)
    e2e_total = float(e2e_times.sum())
    py_times = []
    maps = [(wl["cells"], wl["resolution"], wl["origin"][0], wl["origin"][1])] * R
    for _ in range(min(args.steps, 50)):
        t0 = time.perf_counter()
        opt.evalControl(inputs, costmaps=maps if R > 1 else maps[0])
        py_times.append(time.perf_counter() - t0)
    py_times = np.array(py_times)
    if distributed:
        t = torch.tensor([e2e_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())
    e2e_value = world * units_per_step * args.steps / e2e_total
    d2h = 2 * R * (48 + 6 * T * 4)   # header + control sequence + optimal trajectory, every 32-bit word with its 32-bit sequence tag
    window_misses = opt.windowMissCount()

    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peak_gbs()
    win = 0
    try:
        half = min(max(int(np.ceil((0.5 * T * cfg.model_dt * 1.25 + cfg.circumscribed_radius + 4 * wl["resolution"]) / wl["resolution"])), 8), 96)
        win = ((2 * half + 1 + 15) // 16 * 16) * (2 * half + 1)
    except Exception:
        pass
    alg_bytes = algorithmic_bytes_kernel_a(cfg, win)
    achieved = alg_bytes / (ka_ms * 1e-3) / 1e9 if ka_ms > 0 else 0.0
    fused = launches_timed <= args.steps * iters + 1
    roofline = {"kernel": "k_fused_cluster (rollout + all critics + softmax update, one launch)" if fused else "k_rollout_score",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": measured_traffic("k_fused_cluster@C1") if fused and args.workload == "c1" else None,
                "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms": ka_ms, "launches_timed": ka_n, "peak_source": peak_src,
                "note": "C1 moves <1 MB per launch: it is latency-bound (0.9 MB at HBM speed is 0.14 us), not bandwidth-bound "
                        "(SURVEY.md §8d); roofline_large_batch has the HBM-regime kernels on a batch larger than L2"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(wl, cfg), "n_robots_per_gpu": R,
        "state": f"steady state: {STEADY_STATE_CYCLES} control cycles before the captured / timed ones",
        "p50_optimize_us": float(np.median(step_ms) * 1e3),
        "p50_optimize_us_hot_l2": float(np.median(hot_ms) * 1e3),
        "value_hot_l2": world * units_per_step / (float(np.mean(hot_ms)) * 1e-3),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "p50_latency_us": float(np.median(e2e_times) * 1e6), "mean_latency_us": float(np.mean(e2e_times) * 1e6),
                "timed": "mppi_time_e2e: steady_clock around mppi_optimize_in_place (costmap read in the caller's memory) "
                         "inside the C library",
                "window_misses": int(window_misses),
                "last_cycle_status_and_retries": e2e_status,
                "settle": "0.3 s of untimed control cycles before the e2e loops",
                "two_call": {"p50_latency_us": float(np.median(two_call_times) * 1e6), "h2d_bytes_per_step": int(two_call_h2d),
                             "timed": "mppi_upload_costmap (whole map staged) + mppi_optimize"},
                "p50_latency_us_python_binding": float(np.median(py_times) * 1e6)},
        "gpu_launches": int(launches_timed),
        "roofline": roofline,
        "clocks": clocks,
    }

    opt.shutdown()
    # the two multi-GPU workloads of north_star, at every N (every rank takes part; rank 0 gets the numbers)
    if not args.no_large_batch and args.workload == "c1":
        c4 = section_c4_sharded(args, world, rank)
        c5 = section_c5_fleet(args, world, rank, stream)
        if rank == 0:
            line["c4_sharded"] = c4
            line["c5_fleet"] = c5
    # single-GPU extras (rank 0 at N = 1 only): the HBM-regime kernels, the inflation pass, the reference harness, the CPU baseline
    if rank == 0 and world == 1 and not args.no_large_batch and args.workload == "c1":
        line["roofline_large_batch"] = large_batch_roofline(peak, peak_src, stream, flush)
        line["costmap_inflation"] = inflation_pass(wl, peak, peak_src)
        line["reference_harness"] = reference_harness_section(stream, not args.no_cpu_baseline)
        line["costmap_device_handoff"] = device_handoff_section(wl)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(wl)
    if rank == 0:
        emit(line)
    if distributed:
        dist.destroy_process_group()


def inflation_pass(wl, peak, peak_src):
    """SURVEY.md section 8f row 3, reported beside the headline: the costmap inflation kernel (k_inflate) on the
    workload's own 400x400 map reduced to its lethal cells, and on one GPU's share of the fleet (512 maps of
    200x200).  Kernel time from CUDA events inside the library, L2 flushed before every pass; algorithmic bytes =
    every cell read once + every changed cell written once.  A failure is reported, not hidden."""
    try:
        from navigation2_b200 import inflation, scenarios as S
        kw = dict(inflation_radius=0.70, cost_scaling_factor=3.0)
        layer = inflation.InflationLayer(wl["resolution"], 0.22, **kw)
        out = {"kernel": "k_inflate<14>", "bound": "issue (ncu: issue slots ~80% busy; HBM fraction reported for scale)",
               "peak": peak, "unit": "GB/s", "peak_source": peak_src, "cases": []}
        cells = np.asarray(wl["cells"], dtype=np.uint8)
        one = np.where(cells == 254, 254, 0).astype(np.uint8)[None]
        fleet = np.stack([np.where(S.block_costmap(200, 200, 0.05, seed=100 + k, n_blocks=8) == 254, 254, 0).astype(np.uint8)
                          for k in range(8)])
        fleet = np.concatenate([fleet] * 64)  # 512 maps
        for name, maps, steps in (("C1 map 400x400", one, 200), ("fleet share 512 x 200x200", fleet, 30)):
            before = maps.copy()
            ms = layer.time_resident(maps, steps=steps, warmup=3, flush_bytes=256 << 20)
            probe = before[0].copy()
            layer.updateCosts(probe, 0, 0, probe.shape[1], probe.shape[0])
            changed = int((probe != before[0]).sum())
            if maps.shape[0] > 1:
                changed = 0
                for k in range(8):
                    probe = before[k].copy()
                    layer.updateCosts(probe, 0, 0, probe.shape[1], probe.shape[0])
                    changed += int((probe != before[k]).sum())
                changed *= maps.shape[0] // 8
            alg = int(maps.size) + changed
            out["cases"].append({"workload": name, "kernel_ms": ms, "cells_per_s": maps.size / (ms * 1e-3),
                                 "algorithmic_bytes_per_launch": alg, "achieved": alg / (ms * 1e-3) / 1e9,
                                 "frac": alg / (ms * 1e-3) / 1e9 / peak})
        out["gpu_launches"] = layer.launch_count
        layer.close()
        return out
    except Exception as exc:  # the headline line must still be printed
        return {"error": f"{type(exc).__name__}: {exc}"}


def device_handoff_section(wl):
    """SURVEY.md section 8f row 3 end to end: the costmap never leaves the device.  Per cycle: restore the raw (lethal cells
    only) map, inflate it on the GPU (k_inflate), hand it to the optimiser device-to-device (mppi_set_costmap_device)
    and run the control cycle; against the host route on the same handle (inflated map in host memory,
    mppi_optimize_in_place).  Both loops go through the Python binding, so they carry the same call overhead."""
    try:
        import torch
        from navigation2_b200 import inflation
        cfg = wl["cfg"]
        cells = np.asarray(wl["cells"], dtype=np.uint8)
        raw = np.where(cells == 254, 254, 0).astype(np.uint8)
        sy, sx = raw.shape
        layer = inflation.InflationLayer(wl["resolution"], cfg.inscribed_radius, inflation_radius=0.70, cost_scaling_factor=3.0)
        opt = setup_engine(wl)
        d_raw = torch.from_numpy(raw).cuda()
        d_map = torch.empty_like(d_raw)
        s = torch.cuda.current_stream()
        n = 300

        def device_cycle():
            d_map.copy_(d_raw)
            layer.updateCostsDevice(d_map.data_ptr(), sx, sy, stream=s.cuda_stream)
            opt.setCostmapDevice(d_map.data_ptr(), sx, sy, wl["resolution"], wl["origin"][0], wl["origin"][1],
                                 producer_stream=s.cuda_stream)
            return opt.evalControl(wl["inp"])

        for _ in range(STEADY_STATE_CYCLES):
            device_cycle()
        b0 = opt.h2dByteCount()
        t_dev = []
        for _ in range(n):
            t0 = time.perf_counter()
            device_cycle()
            t_dev.append(time.perf_counter() - t0)
        h2d_dev = (opt.h2dByteCount() - b0) / n
        inflated = d_map.cpu().numpy()
        maps = (inflated, wl["resolution"], wl["origin"][0], wl["origin"][1])
        for _ in range(STEADY_STATE_CYCLES):
            opt.evalControl(wl["inp"], costmaps=maps)
        b0 = opt.h2dByteCount()
        t_host = []
        for _ in range(n):
            t0 = time.perf_counter()
            opt.evalControl(wl["inp"], costmaps=maps)
            t_host.append(time.perf_counter() - t0)
        h2d_host = (opt.h2dByteCount() - b0) / n
        out = {"workload": wl["name"] + "; costmap inflated on the device every cycle",
               "device_route": {"p50_latency_us": float(np.median(t_dev) * 1e6), "h2d_bytes_per_cycle": int(h2d_dev),
                                "steps": "D2D restore of the raw map + k_inflate + mppi_set_costmap_device + mppi_optimize"},
               "host_route": {"p50_latency_us": float(np.median(t_host) * 1e6), "h2d_bytes_per_cycle": int(h2d_host),
                              "steps": "mppi_optimize_in_place on the already inflated map in host memory (no inflation timed)"},
               "timed": "time.perf_counter around the Python calls (both routes carry the binding's overhead)"}
        opt.shutdown()
        layer.close()
        return out
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"}


def large_batch_roofline(peak, peak_src, stream, flush):
    """The same rollout+score kernel on one GPU's worth of BASELINE.json configs[3] (2^20 x 56, Philox
    noise): working set (2 x 235 MB of noise) >> L2, so this is the HBM-bound regime."""
    import torch
    import navigation2_b200 as nb
    wl = workload_c1()
    cfg = large_batch_config(1 << 20)
    opt = setup_engine(wl, cfg=cfg)
    for _ in range(STEADY_STATE_CYCLES):
        opt.evalControl(wl["inp"])
    opt.setResidentCapture(True)
    opt.evalControl(wl["inp"])
    opt.setKernelTiming(True)
    step_ms = time_resident(opt, 10, 3, flush, stream)
    ka_ms, ka_n = opt.kernelTimeMs(reset=True)
    kc_ms, kc_n = opt.kernelCTimeMs(reset=True)
    win = 0
    alg = algorithmic_bytes_kernel_a(cfg, win)
    achieved = alg / (ka_ms * 1e-3) / 1e9
    units = int(cfg.batch_size) * int(cfg.time_steps)
    # kernel C re-reads the noise columns once (4 B x axes per trajectory-step) + cost and weight per trajectory
    alg_c = units * 8 + int(cfg.batch_size) * 12
    achieved_c = alg_c / (kc_ms * 1e-3) / 1e9 if kc_ms > 0 else 0.0
    out = {"workload": "DiffDrive 1048576x56 (configs[3] on one GPU), default critics, Philox noise",
           "kernel": "k_rollout_lean (kernel A: rollout + per-trajectory critic terms)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
           "frac": achieved / peak, "traffic": measured_traffic("k_rollout_score@2^20x56"),
           "algorithmic_bytes_per_launch": alg, "kernel_ms": ka_ms, "launches_timed": ka_n, "peak_source": peak_src,
           "limiter": "instruction issue and the latency of the per-trajectory serial chain (ncu, profiles/r02_kernels_c4_raw.json: issue "
                      "slots 75% busy at 120 thread-instructions per trajectory-step, 64 registers, 32 of 64 warps resident), not HBM "
                      "(dram throughput 31%)",
           "kernel_c": {"kernel": "k_softmax_partial", "bound": "hbm", "achieved": achieved_c, "peak": peak, "unit": "GB/s",
                        "frac": achieved_c / peak, "traffic": measured_traffic("k_softmax_partial@2^20x56"),
                        "algorithmic_bytes_per_launch": alg_c, "kernel_ms": kc_ms, "launches_timed": kc_n},
           "step_ms": float(np.mean(step_ms)), "value": units / (float(np.mean(step_ms)) * 1e-3), "unit_value": UNIT}
    opt.shutdown()
    return out


def cpu_baseline(wl):
    """The oracle port (kind "port": the reference's Eigen build needs ROS 2) on one host core,
    ~10-20 s of the same workload."""
    from oracle.binding import OracleOptimizer
    from navigation2_b200 import scenarios as S
    cfg = wl["cfg"]
    B = int(cfg.batch_size)
    if int(cfg.n_robots) != 1:
        from navigation2_b200 import params as P
        cfg = P.nav2_bringup_config()
        B = int(cfg.batch_size)
    o = OracleOptimizer(cfg, fast=True)
    o.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
    nvx, nvy, nwz = S.seeded_noise(B, int(cfg.time_steps), (cfg.vx_std, cfg.vy_std, cfg.wz_std), holonomic=wl["holo"])
    o.setNoise(nvx, nvy, nwz)
    sec = o.timeOptimize(wl["inp"], 20)
    cycles = int(min(max(12.0 / max(sec, 1e-6), 20), 20000))
    sec = o.timeOptimize(wl["inp"], cycles)
    units = B * int(cfg.time_steps) * int(cfg.iteration_count)
    o.shutdown()
    out = {"value": units / sec, "unit": UNIT, "cores": 1, "kind": "port", "ms_per_optimize": sec * 1e3,
           "sample": f"{cycles} evalControl cycles of the same workload (~{cycles * sec:.0f} s), oracle/ C++ restatement, "
                     "g++ -O3 -mavx2 -mfma, single thread like the reference's control path",
           "host_cpu": cpu_model(), "host_cores": os.cpu_count()}
    # Courtesy upper bound: every host core runs its own independent controller on the same
    # workload (the reference has no intra-optimize() threading; a fleet of controllers is what more cores buy).
    try:
        n_thr = max(1, len(os.sched_getaffinity(0)))
        per = max(3, int(4.0 / max(sec, 1e-6)))

        def worker(res, i):
            ow = OracleOptimizer(cfg, fast=True)
            ow.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
            ow.setNoise(nvx, nvy, nwz)
            ow.timeOptimize(wl["inp"], 3)
            res[i] = ow.timeOptimize(wl["inp"], per) * per      # the C call releases the GIL
            ow.shutdown()

        res = [0.0] * n_thr
        threads = [threading.Thread(target=worker, args=(res, i)) for i in range(n_thr)]
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        wall = max(max(res), 1e-9)
        out["all_cores"] = {"value": units * per * n_thr / wall, "unit": UNIT, "cores": n_thr, "kind": "port",
                            "sample": f"{n_thr} threads x {per} evalControl cycles, one independent controller per thread "
                                      f"(wall {time.perf_counter() - t0:.1f} s incl. set-up)"}
    except Exception as exc:
        out["all_cores"] = {"error": f"{type(exc).__name__}: {exc}"}
    return out


def _parity_errors(g, c, g_costs, c_costs):
    """Worst deviations of one GPU cycle from the oracle's: (control sequence abs, cmd_vel abs, cost relative)."""
    u_err = float(np.max(np.abs(np.asarray(g.control_sequence, dtype=np.float64) - c.control_sequence)))
    cmd_err = float(np.max(np.abs(np.array(g.cmd, dtype=np.float64) - np.array(c.cmd))))
    gc, cc = np.asarray(g_costs, dtype=np.float64), np.asarray(c_costs, dtype=np.float64)
    cost_err = float(np.max(np.abs(gc - cc) / np.maximum(np.abs(cc), 1e-6)))
    return u_err, cmd_err, cost_err


def section_c4_sharded(args, world, rank):
    """BASELINE.json configs[3]: ONE 2^20 x 56 batch sharded along the trajectory axis over the ranks (strong scaling).
    A cycle is one mppi_shard_cycle call per rank: cycle inputs H2D, four kernels per iteration with the two
    exchanges (MAX of the reduction words, gather of the softmax slots) done inside them over peer memory, result
    block read from mapped memory.  Timed inside the library around whole calls, max over ranks.  Before anything is
    timed, the first cycle is checked against the CPU oracle fed the GPU's own Philox tensors (rank 0).
    At N > 1 rank 0 also times the unsharded batch on its own GPU (`n1_ms`), so the line carries both ends of the
    strong-scaling ratio from the same box and process."""
    import torch
    import torch.distributed as dist
    from navigation2_b200 import capi
    from navigation2_b200.sharded import ShardedOptimizer
    out = {"workload": "C4 DiffDrive 1048576x56 (BASELINE.json configs[3]) sharded along the trajectory axis, Philox noise "
                       "keyed by global trajectory id", "scaling": "strong", "n_gpus": world, "unit": UNIT}
    try:
        wl = workload_c1()
        batch = 1 << 20
        cfg = large_batch_config(batch, world=world, rank=rank)
        opt = ShardedOptimizer(cfg)
        opt.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
        peer = bool(opt._peer_exchange)
        units = batch * int(cfg.time_steps) * int(cfg.iteration_count)
        one = cpu = None
        if rank == 0:
            from oracle.binding import OracleOptimizer
            cfg1 = large_batch_config(batch)
            one = opt if world == 1 else ShardedOptimizer(cfg1)      # unsharded: the global Philox tensors and the N = 1 time
            if one is not opt:
                one.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
            cpu = OracleOptimizer(cfg1)
            cpu.setNoise(*[one.getTensor(t) for t in (capi.TENSOR_NOISE_VX, capi.TENSOR_NOISE_VY, capi.TENSOR_NOISE_WZ)])
            cpu.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
        if world > 1:
            dist.barrier()               # an exchange waits 2 s for its peers: start the cycle together
        g = opt.evalControl(wl["inp"])
        if rank == 0:
            c = cpu.evalControl(wl["inp"])
            lo = opt.getTensor(capi.TENSOR_COSTS)
            u_err, cmd_err, cost_err = _parity_errors(g, c, lo, cpu.getTensor(capi.TENSOR_COSTS)[: lo.shape[0]])
            out["parity_checked"] = bool(u_err <= 1e-5 and cmd_err <= 1e-5 and cost_err <= 1e-4)
            out["parity"] = {"against": "CPU oracle, same Philox tensors, first cycle", "control_sequence_abs": u_err,
                             "cmd_vel_abs": cmd_err, "cost_rel": cost_err, "tolerances": [1e-5, 1e-5, 1e-4]}
            cpu.shutdown()
        if world > 1:
            dist.barrier()
        for _ in range(STEADY_STATE_CYCLES):
            opt.evalControl(wl["inp"])
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        n = int(min(max(args.steps, 20), 200))
        l0 = opt.launchCount()
        t = opt.timeCycles(wl["inp"], n)
        torch.cuda.synchronize()
        launches = opt.launchCount() - l0
        total = torch.tensor([float(t.sum())], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        ms = float(total.item()) * 1e3 / n
        out.update({"ms_per_cycle": ms, "value": units / (ms * 1e-3), "cycles_timed": n,
                    "p50_ms_rank0": float(np.median(t) * 1e3),
                    "exchange": ("peer" if peer else "nccl") if world > 1 else "none (one rank)",
                    "gpu_launches_per_cycle": launches / n,
                    "timed": "steady_clock around whole mppi_shard_cycle calls inside the library (host preparation, H2D of "
                             "the cycle block, kernels + exchanges, result block), sum over cycles, max over ranks; "
                             f"after {STEADY_STATE_CYCLES} warm-up cycles; working set 470 MB of noise / N per GPU"})
        if rank == 0:
            if one is opt:
                out["n1_ms"] = ms
            else:
                for _ in range(STEADY_STATE_CYCLES + 1):
                    one.evalControl(wl["inp"])
                out["n1_ms"] = float(one.timeCycles(wl["inp"], 30).sum()) * 1e3 / 30
                one.shutdown()
            out["strong_scaling_speedup"] = out["n1_ms"] / ms
        if world > 1:
            dist.barrier()
        opt.shutdown()
    except Exception as exc:  # the headline line must still be printed
        out["error"] = f"{type(exc).__name__}: {exc}"
        out["parity_checked"] = False
    return out


def fleet_workload(n_robots, rank):
    """BASELINE.json configs[4] for one GPU's robot-set: every robot its own 200x200 costmap (16 distinct seeded maps
    cycled over the fleet, each robot its own array), its own pose heading and path."""
    from navigation2_b200 import scenarios as S
    import navigation2_b200 as nb
    rng = np.random.default_rng(1000 + rank)
    base = [S.block_costmap(200, 200, 0.05, seed=31 + 16 * rank + k, keep_clear=(5.0, 5.0), n_blocks=10) for k in range(16)]
    maps, inputs = [], []
    for r in range(n_robots):
        maps.append((base[r % 16].copy(), 0.05, 0.0, 0.0))
        yaw = float(rng.uniform(-3.1, 3.1))
        path = S.arc_path((5.0, 5.0, yaw), n_points=100, curvature=float(rng.uniform(-0.25, 0.25)))
        inputs.append(nb.CycleInput(pose=(5.0, 5.0, yaw), speed=(0.2, 0.0, 0.0), path=path, goal=tuple(path[-1]),
                                    goal_checker_xy_tolerance=0.25))
    return maps, inputs


def section_c5_fleet(args, world, rank, stream):
    """BASELINE.json configs[4]: a fleet of independent robots, 2000 x 56 each, one robot-set per GPU, no collective
    (weak scaling).  device_value: resident replays of a steady-state fleet cycle, CUDA events, L2 flushed;
    e2e_value: mppi_optimize_in_place with every robot's costmap and path in host memory.  Two robots of rank 0's
    set are checked against their own CPU oracle first."""
    import torch
    import torch.distributed as dist
    import navigation2_b200 as nb
    from navigation2_b200 import capi, params as P
    R = int(args.robots)
    out = {"workload": f"C5 fleet (BASELINE.json configs[4]): {R} independent robots per GPU, own 200x200 costmap + path "
                       "each, DiffDrive 2000x56", "scaling": "weak", "n_gpus": world, "robots_per_gpu": R, "unit": UNIT}
    try:
        cfg = P.nav2_bringup_config(n_robots=R)
        cfg1 = P.nav2_bringup_config()
        maps, inputs = fleet_workload(R, rank)
        opt = nb.Optimizer(cfg)
        units = R * int(cfg.batch_size) * int(cfg.time_steps) * int(cfg.iteration_count)
        first = opt.evalControl(inputs, costmaps=maps)
        if rank == 0:
            from oracle.binding import OracleOptimizer
            worst = [0.0, 0.0, 0.0]
            for r in sorted({0, R - 1}):
                cpu = OracleOptimizer(cfg1)
                cpu.setNoise(*[opt.getTensor(t, r) for t in (capi.TENSOR_NOISE_VX, capi.TENSOR_NOISE_VY, capi.TENSOR_NOISE_WZ)])
                cpu.setCostmap(*maps[r])
                c = cpu.evalControl(inputs[r])
                errs = _parity_errors(first[r], c, opt.getTensor(capi.TENSOR_COSTS, r), cpu.getTensor(capi.TENSOR_COSTS))
                worst = [max(a, b) for a, b in zip(worst, errs)]
                cpu.shutdown()
            out["parity_checked"] = bool(worst[0] <= 1e-5 and worst[1] <= 1e-5 and worst[2] <= 1e-4)
            out["parity"] = {"against": f"CPU oracle per robot (robots 0 and {R - 1}), same Philox tensors, first cycle",
                             "control_sequence_abs": worst[0], "cmd_vel_abs": worst[1], "cost_rel": worst[2]}
        for _ in range(STEADY_STATE_CYCLES):
            opt.evalControl(inputs, costmaps=maps)
        # device: stage the maps, capture a steady-state cycle, replay it resident
        for r in range(R):
            opt.setCostmap(*maps[r], robot=r)
        opt.setResidentCapture(True)
        opt.evalControl(inputs)
        n_dev = int(min(max(args.steps, 10), 50))
        l0 = opt.launchCount()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        dev_ms = time_resident(opt, n_dev, 3, True, stream)
        launches = (opt.launchCount() - l0) * n_dev // (n_dev + 3)
        opt.setResidentCapture(False)
        n_e2e = int(min(max(args.steps, 10), 40))
        opt.timeE2EViews(maps, inputs, 3)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        b0 = opt.h2dByteCount()
        e2e_t = opt.timeE2EViews(maps, inputs, n_e2e)
        h2d = (opt.h2dByteCount() - b0) / n_e2e
        tot = torch.tensor([float(dev_ms.sum()) * 1e-3, float(e2e_t.sum())], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = float(tot[0].item()) / n_dev, float(tot[1].item()) / n_e2e
        out.update({"device_ms_per_fleet_cycle": dev_s * 1e3, "device_value": world * units / dev_s,
                    "e2e_ms_per_fleet_cycle": e2e_s * 1e3, "e2e_value": world * units / e2e_s,
                    "h2d_bytes_per_cycle": int(h2d), "d2h_bytes_per_cycle": int(2 * R * (48 + 6 * int(cfg.time_steps) * 4)),
                    "gpu_launches": int(launches), "window_misses": int(opt.windowMissCount()),
                    "host_threads_per_rank": int(os.environ.get("MPPI_B200_HOST_THREADS", "0")), "host_cores": os.cpu_count(),
                    "timed": "device: mppi_time_resident (CUDA events, L2 flushed before every replay), max over ranks; e2e: "
                             "steady_clock around mppi_optimize_in_place calls with per-robot host costmaps inside the library"})
        opt.shutdown()
    except Exception as exc:
        out["error"] = f"{type(exc).__name__}: {exc}"
        out["parity_checked"] = False
    return out


def reference_harness_section(stream, with_cpu):
    """The reference's own benchmark configuration (benchmark/optimizer_benchmark.cpp:49-99: iteration_count 2, 40x40
    fixture), timed the same three ways as the headline: device-resident, end to end, CPU oracle on one core."""
    try:
        wl = workload_reference_harness()
        cfg = wl["cfg"]
        opt = setup_engine(wl)
        units = int(cfg.batch_size) * int(cfg.time_steps) * int(cfg.iteration_count)
        for _ in range(STEADY_STATE_CYCLES):
            opt.evalControl(wl["inp"])
        opt.setResidentCapture(True)
        opt.evalControl(wl["inp"])
        dev_ms = time_resident(opt, 200, 5, True, stream)
        opt.setResidentCapture(False)
        opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 10, in_place=True)
        e2e = opt.timeE2E(wl["cells"], wl["resolution"], wl["origin"], wl["inp"], 300, in_place=True)
        out = {"workload": wl["name"], "unit": UNIT, "value": units / (float(np.mean(dev_ms)) * 1e-3),
               "p50_optimize_us": float(np.median(dev_ms) * 1e3),
               "e2e": {"value": units * len(e2e) / float(e2e.sum()), "p50_latency_us": float(np.median(e2e) * 1e6)}}
        opt.shutdown()
        if with_cpu:
            from oracle.binding import OracleOptimizer
            from navigation2_b200 import scenarios as S
            o = OracleOptimizer(cfg, fast=True)
            o.setCostmap(wl["cells"], wl["resolution"], *wl["origin"])
            o.setNoise(*S.seeded_noise(int(cfg.batch_size), int(cfg.time_steps), (cfg.vx_std, cfg.vy_std, cfg.wz_std)))
            o.timeOptimize(wl["inp"], 5)
            sec = o.timeOptimize(wl["inp"], 500)
            o.shutdown()
            out["cpu_baseline"] = {"value": units / sec, "unit": UNIT, "cores": 1, "kind": "port", "ms_per_optimize": sec * 1e3,
                                   "sample": "500 evalControl cycles, oracle port, 1 thread"}
        return out
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"}


def run_c4(args, world, rank, local_rank):
    """`--workload c4`: the sharded 2^20 x 56 batch as the headline line (development runs; the default line carries the
    same numbers in its c4_sharded section)."""
    import torch.distributed as dist
    sec = section_c4_sharded(args, world, rank)
    if rank == 0:
        if "error" in sec:
            raise SystemExit("c4: " + sec["error"])
        wl = workload_c1()
        cfg = large_batch_config(1 << 20)
        emit({
            "metric": METRIC, "value": sec["value"], "unit": UNIT, "n_gpus": world, "steps": sec["cycles_timed"],
            "warmup": STEADY_STATE_CYCLES, "ms_per_step": sec["ms_per_cycle"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(dict(wl, name=sec["workload"]), cfg), "c4_sharded": sec,
            "gpu_launches": int(round(sec["gpu_launches_per_cycle"] * sec["cycles_timed"]))})
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line: dict) -> None:
    """The one JSON line of the contract, on the process's real stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    # Native libraries write to stdout too (NCCL announces its version there at initialisation); the driver wants
    # exactly one JSON line on stdout.  Everything else is sent to stderr: fd 1 is pointed at fd 2 for the whole
    # run, the result line goes to a duplicate of the original stdout.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    from navigation2_b200 import build
    build.build_all()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
