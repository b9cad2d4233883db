"""Batch sharding across GPUs: one process per GPU, torch.distributed (NCCL) for the
two small exchanges per iteration.  The trajectory axis is split evenly; costmap, path and control
sequence are replicated; Philox noise is keyed by the global trajectory index, so the result does
not depend on the number of ranks.

  kernel A (rollout + per-trajectory critics)
  all-reduce MAX over the 8 reduction words      furthest path point, -min repulsion, any-free flags
  kernels B, C, D1 (path critics, local softmax partials, local (min, sum, W) slot)
  all-gather of the (2 + 3T)-float slots          rescaled by exp(-(m - m_g)/temperature) <= 1 in kernel D2 (m: the rank's min, m_g: the global min)
  kernel D2 (combine, Savitzky-Golay filter, constraints, validator) — run redundantly on every rank
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from .optimizer import CycleInput, CycleResult, Optimizer, _raise_for


class _DeviceBuffer:
    """Zero-copy view of a raw device pointer for torch.as_tensor."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class ShardedOptimizer(Optimizer):
    def __init__(self, cfg: capi.MppiConfig, group=None):
        import torch
        import torch.distributed as dist
        if cfg.n_robots != 1:
            raise ValueError("batch sharding drives one robot; fleets are partitioned per GPU without collectives")
        super().__init__(cfg)
        self._torch = torch
        self._dist = dist
        self._group = group
        self.world = int(cfg.shard_world)
        self._distributed = self.world > 1
        if self._distributed and (not dist.is_initialized() or dist.get_world_size(group) != self.world):
            raise RuntimeError("shard_world does not match the torch.distributed world size")
        # Exchanges over peer memory (folded into the kernels, KArgs::xch_*): every rank exports its mailbox as a CUDA IPC
        # handle, the handles are all-gathered once, and from then on a cycle is mppi_optimize's own captured graph of
        # four kernels per iteration on the handle's own stream.  MPPI_B200_NCCL_EXCHANGE=1 (or a non-NCCL backend)
        # keeps torch.distributed collectives between the mppi_shard_* phases instead (A/B measurements).
        self._peer_exchange = not self._distributed   # one rank: nothing to exchange
        if self._distributed and os.environ.get("MPPI_B200_NCCL_EXCHANGE", "0") != "1" and dist.get_backend(group) == "nccl":
            mine = (C.c_ubyte * 64)()
            _raise_for(self._lib, self._h, self._lib.mppi_shard_ipc_handle(self._h, mine))
            t = torch.tensor(list(mine), dtype=torch.uint8, device="cuda")
            gathered = torch.empty(64 * self.world, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(gathered, t, group=group)
            handles = (C.c_ubyte * (64 * self.world))(*gathered.cpu().tolist())
            _raise_for(self._lib, self._h, self._lib.mppi_shard_open_peers(self._h, handles))
            dist.barrier(group=group)      # every mailbox is open everywhere before the first exchange
            self._peer_exchange = True
        self._stream = None
        if not self._peer_exchange:
            # Kernels and torch.distributed collectives must be ordered on ONE stream.  The legacy default stream has
            # handle 0, which mppi_set_stream reads as "the handle's own stream": kernels and collectives would then
            # run unordered (ranks diverge).  So a dedicated torch stream is used unless the caller made one current.
            cur = torch.cuda.current_stream()
            self._stream = cur if cur.cuda_stream != 0 else torch.cuda.Stream()
            _raise_for(self._lib, self._h, self._lib.mppi_set_stream(self._h, self._stream.cuda_stream))
            self._bind_exchange_buffers()

    def _bind_exchange_buffers(self) -> None:
        """Views of the library's exchange buffers for the torch.distributed collectives.  The library re-allocates
        them whenever the configuration is set (mppi_set_config -> allocate), so the views are rebuilt then."""
        torch = self._torch
        ar1, slot, slot_all = C.c_void_p(), C.c_void_p(), C.c_void_p()
        n1, n2 = C.c_size_t(), C.c_size_t()
        _raise_for(self._lib, self._h, self._lib.mppi_shard_buffers(
            self._h, C.byref(ar1), C.byref(n1), C.byref(slot), C.byref(n2), C.byref(slot_all)))
        self._ar1 = torch.as_tensor(_DeviceBuffer(ar1.value, n1.value, "<i4"), device="cuda")
        self._slot = torch.as_tensor(_DeviceBuffer(slot.value, n2.value, "<f4"), device="cuda")
        self._slot_all = torch.as_tensor(_DeviceBuffer(slot_all.value, n2.value * self.world, "<f4"), device="cuda")

    def setConfig(self, cfg: capi.MppiConfig) -> None:
        """Dynamic parameter update of a live sharded handle.  The shard layout (shard_world, shard_rank, time_steps)
        is fixed once the peers' mailboxes are open: the library rejects a change (MPPI_ERR_UNSUPPORTED)."""
        if int(cfg.n_robots) != 1 or int(cfg.shard_world) != self.world:
            raise ValueError("n_robots / shard_world cannot change on a live sharded handle")
        super().setConfig(cfg)
        if not self._peer_exchange:
            self._bind_exchange_buffers()

    def evalControl(self, inp: CycleInput, raise_on_failure: bool = True) -> CycleResult:
        lib, h = self._lib, self._h
        c_in = (capi.MppiCycleIn * 1)(inp.to_c())
        c_out = (capi.MppiCycleOut * 1)()
        T = self.T
        u = np.zeros((3, T), dtype=np.float32)
        traj = np.zeros((3, T), dtype=np.float32)
        fp = C.POINTER(C.c_float)
        c_out[0].control_vx = u[0].ctypes.data_as(fp)
        c_out[0].control_vy = u[1].ctypes.data_as(fp)
        c_out[0].control_wz = u[2].ctypes.data_as(fp)
        c_out[0].optimal_trajectory = traj.ctypes.data_as(fp)
        if self._peer_exchange:
            _raise_for(lib, h, lib.mppi_shard_cycle(h, c_in, c_out))
            return self._result(c_out[0], u, traj, raise_on_failure)
        with self._torch.cuda.stream(self._stream):   # the collectives order against the stream the kernels run on
            _raise_for(lib, h, lib.mppi_shard_begin(h, c_in))
            iters = int(self.cfg.iteration_count)
            while True:
                for it in range(iters):
                    _raise_for(lib, h, lib.mppi_shard_rollout(h))
                    if self._distributed:
                        self._dist.all_reduce(self._ar1, op=self._dist.ReduceOp.MAX, group=self._group)
                    _raise_for(lib, h, lib.mppi_shard_score(h))
                    if self._distributed:
                        self._dist.all_gather_into_tensor(self._slot_all, self._slot, group=self._group)
                    else:
                        self._slot_all.copy_(self._slot)
                    _raise_for(lib, h, lib.mppi_shard_update(h, 1 if it == iters - 1 else 0))
                rc = lib.mppi_shard_end(h, c_out)
                if rc == 1:
                    continue  # fallback(): every rank saw the same flags and retries together
                _raise_for(lib, h, rc)
                break
        return self._result(c_out[0], u, traj, raise_on_failure)

    def timeCycles(self, inp: CycleInput, cycles: int) -> np.ndarray:
        """Per-call seconds of `cycles` sharded cycles with the exchanges over peer memory, timed inside the library
        (every rank calls it)."""
        if not self._peer_exchange:
            raise RuntimeError("timeCycles needs the peer-memory exchange (NCCL backend)")
        c_in = (capi.MppiCycleIn * 1)(inp.to_c())
        c_out = (capi.MppiCycleOut * 1)()
        T = self.T
        u = np.zeros((3, T), dtype=np.float32)
        traj = np.zeros((3, T), dtype=np.float32)
        fp = C.POINTER(C.c_float)
        c_out[0].control_vx = u[0].ctypes.data_as(fp)
        c_out[0].control_vy = u[1].ctypes.data_as(fp)
        c_out[0].control_wz = u[2].ctypes.data_as(fp)
        c_out[0].optimal_trajectory = traj.ctypes.data_as(fp)
        t = np.zeros(cycles, dtype=np.float64)
        _raise_for(self._lib, self._h, self._lib.mppi_time_shard_cycles(self._h, c_in, c_out, int(cycles), t.ctypes.data_as(C.POINTER(C.c_double))))
        return t

    def _result(self, o, u, traj, raise_on_failure):
        res = CycleResult(status=o.status, cmd=(o.cmd_vx, o.cmd_vy, o.cmd_wz), control_sequence=u,
                          optimal_trajectory=traj.T.copy(), fail_flag=bool(o.fail_flag), validation=o.validation,
                          retries=o.retries, furthest_reached_path_point=o.furthest_reached_path_point)
        if raise_on_failure and res.status != capi.MPPI_OK:
            _raise_for(self._lib, self._h, res.status)
        return res


# ---- host restatement of the two exchanges (what kernels A / D2 do on the device), used by the
# ---- world_size-2 gloo tests on CPU and as executable documentation of the protocol ----

def encode_min(values: np.ndarray) -> np.ndarray:
    """float32 -> int32 such that MAX over the codes is MIN over the floats (kernel: enc_min)."""
    i = np.asarray(values, dtype=np.float32).view(np.int32).astype(np.int64)
    ordered = np.where(i >= 0, i, i ^ 0x7FFFFFFF)
    return (~ordered).astype(np.int32)


def decode_min(codes: np.ndarray) -> np.ndarray:
    o = (~np.asarray(codes, dtype=np.int32).astype(np.int64))
    bits = np.where(o >= 0, o, o ^ 0x7FFFFFFF).astype(np.int32)
    return bits.view(np.float32)


def local_slot(costs: np.ndarray, cv: np.ndarray, inv_temp: float) -> np.ndarray:
    """One rank's AR-2 slot: (min cost, sum of exp(-(c - min)/T), weighted control sums [3T]).
    costs: (B_local,), cv: (3, T, B_local)."""
    m = np.float32(costs.min())
    w = np.exp(-np.float32(inv_temp) * (costs.astype(np.float32) - m)).astype(np.float32)
    W = (cv.astype(np.float64) * w.astype(np.float64)[None, None, :]).sum(axis=2)
    return np.concatenate([[m, w.astype(np.float64).sum()], W.reshape(-1)]).astype(np.float32)


def combine_slots(slots: np.ndarray, inv_temp: float) -> np.ndarray:
    """All-gathered slots (world, 2 + 3T) -> softmax-weighted mean control sequence (3T,),
    rescaling every rank by exp(-(m - m_g)/temperature) <= 1, m the rank's own minimum and m_g the
    global one (kernel D2)."""
    slots = np.asarray(slots, dtype=np.float32)
    m = slots[:, 0].min()
    scale = np.exp(-np.float32(inv_temp) * (slots[:, 0] - m)).astype(np.float64)
    S = (slots[:, 1].astype(np.float64) * scale).sum()
    W = (slots[:, 2:].astype(np.float64) * scale[:, None]).sum(axis=0)
    return (W / S).astype(np.float32)
