"""Ad-hoc: per-phase, per-warp SM-clock breakdown of the fused kernel at C1 (MPPI_B200_PHASE_CLOCKS=1)."""
import os, sys, ctypes as C
os.environ["MPPI_B200_PHASE_CLOCKS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from navigation2_b200 import capi
wl = bench.workload_c1()
opt = bench.setup_engine(wl)
maps = (wl["cells"], wl["resolution"], wl["origin"][0], wl["origin"][1])
for _ in range(int(os.environ.get("PHASE_CLOCK_CYCLES", "15"))):
    try:
        opt.evalControl(wl["inp"], costmaps=maps)
    except Exception as e:      # experiments that corrupt results on purpose still leave the clocks
        print("evalControl:", type(e).__name__)
out = np.zeros(16 * 16 * 32, dtype=np.int64)
rc = opt._lib.mppi_get_tensor(opt._h, 0, capi.TENSOR_PHASE_CLOCKS, out.ctypes.data_as(C.c_void_p), out.nbytes)
assert rc == 0, rc
clk = out.reshape(16, 16, 32)          # [rank][warp][slot]
order = [(0, "start"), (16, "smem carve"), (17, "noise copies issued"), (18, "param copies issued"), (19, "window issued + landed"),
         (1, "sync; u0 fix"), (2, "sync"), (3, "P1 velocity chains + sync"), (4, "P2 sincos + sync"), (5, "P3 x/y chains + sync"),
         (15, "P3b cell loop (no sync)"), (20, "sync after P3b"), (21, "P4 serial scans (no sync)"), (6, "furthest search"),
         (7, "reductions + cluster.sync"), (8, "block_b + sync"), (9, "P5 path critics + min"), (10, "cluster.sync"),
         (11, "P6 softmax + weighted sums"), (12, "cluster.sync"), (13, "combine + cluster.sync")]
order += [(k, f"finalize step {k - 22}") for k in range(22, 32)] + [(14, "finalize end")]
t0 = clk[:, :, 0].min()
print(f"{'phase':34s} {'rank0 w0':>9s} {'r0 min':>8s} {'r0 max':>8s} {'all min':>8s} {'all max':>8s}   (cycles since kernel start; delta of rank0 warp0 in brackets)")
prev = 0
for slot, name in order:
    c = clk[:, :, slot]
    if not c.any():
        continue
    r0 = c[0][c[0] > 0] - t0
    al = c[c > 0] - t0
    w0 = int(clk[0, 0, slot] - t0) if clk[0, 0, slot] else -1
    print(f"{name:34s} {w0:9d} {int(r0.min()):8d} {int(r0.max()):8d} {int(al.min()):8d} {int(al.max()):8d}   [{w0 - prev:+d}]")
    if w0 >= 0:
        prev = w0
print("per-warp arrival at 'P3b cell loop' end, rank 0:", [int(v - t0) for v in clk[0, :, 15]])
print("per-warp arrival at 'P4 scans' end, rank 0:    ", [int(v - t0) for v in clk[0, :, 21]])
print("per-warp arrival at 'P1' end (before sync) n/a; P2 end slot 4, rank 0:", [int(v - t0) for v in clk[0, :, 4]])
print("per-rank warp0 at P6 end:", [int(v - t0) for v in clk[:, 0, 11]])
for slot in (2, 3, 4, 5, 15, 20, 21, 6, 8, 9, 11):
    print(f"slot {slot:2d} rank 0 per warp:", [int(v - t0) for v in clk[0, :, slot]])
print("slot 5 rank 3 per warp:", [int(v - t0) for v in clk[3, :, 5]])
print("slot 15 rank 3 per warp:", [int(v - t0) for v in clk[3, :, 15]])
for slot in (22, 23):
    print(f"DEBUG slot {slot:2d} rank 1 per warp:", [int(v - t0) if v else 0 for v in clk[1, :, slot]])
print(f"DEBUG slot 4/5 rank 1 per warp:", [int(v - t0) for v in clk[1, :, 4]], [int(v - t0) for v in clk[1, :, 5]])
for slot in (24, 25, 26):
    print(f"DEBUG slot {slot:2d} rank 1 per warp:", [int(v - t0) if v else 0 for v in clk[1, :, slot]])
print("DEBUG redo counts (slot 27) rank 1 per warp:", [int(v) for v in clk[1, :, 27]], "rank 0:", [int(v) for v in clk[0, :, 27]])
for slot in (5, 28, 29, 30, 31, 15):
    print(f"DEBUG 3b slot {slot:2d} rank 1 per warp:", [int(v - t0) if v else 0 for v in clk[1, :, slot]])
w0 = clk[:, 1, 30]; w1 = clk[:, 1, 31]
print("WALL kernel entry spread over ranks (ns):", int(w0.max() - w0.min()), " rank0 in-kernel ns:", int(w1[0] - w0[0]),
      " other ranks in-kernel ns:", [int(b - a) for a, b in zip(w0[1:], w1[1:])][:5], " first entry -> last exit ns:", int(w1.max() - w0.min()))
print("WALL rank0 cycles:", int(clk[0, 0, 14] - clk[0, 0, 0]), "=> effective MHz:", (clk[0, 0, 14] - clk[0, 0, 0]) / max(1, (w1[0] - w0[0])) * 1e3)

print("P5 detail, rank 1 per warp (cycles since kernel start): start of P5 (slot 8), after the PathAlign pre-pass (22), after the critic loop (23), before the copy wait (24), end (9)")
for slot in (8, 22, 23, 24, 9):
    print(f"  slot {slot:2d}:", [int(v - t0) if v else 0 for v in clk[1, :, slot]])
