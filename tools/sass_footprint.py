#!/usr/bin/env python
"""Joins an ncu `--page source --csv` export (SASS view, one kernel) with `nvdisasm -g` line info of the same build:
executed code footprint (distinct SASS instructions that ran at least once), warp instructions and stall samples per
source line and per line range.  Instruction fetch is what a once-per-launch, latency-bound kernel pays for every
distinct instruction, so the footprint per phase is the number to shrink.

usage: sass_footprint.py <source.csv[.gz]> <nvdisasm -g text of the kernel> [ranges: name:lo-hi,...]"""
import csv, gzip, re, sys, collections

def load_ncu(path):
    op = gzip.open if path.endswith(".gz") else open
    rows = list(csv.reader(op(path, "rt")))
    hdr = rows[1]
    # an export may hold several launches one after the other: the first one is taken
    nxt = [i for i, r in enumerate(rows) if i > 1 and r and r[0] == "Kernel Name"]
    return hdr, rows[2:(nxt[0] if nxt else len(rows))]

def load_lines(path):
    out = []
    cur = ("?", 0)
    pat_line = re.compile(r'//## File "([^"]+)", line (\d+)')
    pat_ins = re.compile(r'^\s*/\*([0-9a-f]+)\*/\s+(.*);')
    for ln in open(path):
        m = pat_line.search(ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = pat_ins.match(ln)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out

def main():
    hdr, data = load_ncu(sys.argv[1])
    ins = load_lines(sys.argv[2])
    assert len(ins) == len(data), (len(ins), len(data))
    iI, iS, iN = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("stall_no_inst")
    per = collections.defaultdict(lambda: [0, 0, 0, 0])   # footprint, warp-instr, samples, no_inst
    for (addr, (f, l), txt), r in zip(ins, data):
        n = int(r[iI])
        e = per[(f, l)]
        if n > 0:
            e[0] += 1
        e[1] += n; e[2] += int(r[iS]); e[3] += int(r[iN])
    tot = [sum(v[k] for v in per.values()) for k in range(4)]
    print(f"total: executed footprint {tot[0]} instr ({tot[0] * 16} B), warp-instr {tot[1]}, samples {tot[2]}, no_inst {tot[3]}")
    if len(sys.argv) > 3:
        for spec in sys.argv[3].split(","):
            name, rng = spec.split(":")
            lo, hi = map(int, rng.split("-"))
            acc = [0, 0, 0, 0]
            for (f, l), v in per.items():
                if f == "mppi_kernels.cu" and lo <= l <= hi:
                    for k in range(4):
                        acc[k] += v[k]
            print(f"{name:28s} lines {lo:5d}-{hi:5d}: footprint {acc[0]:5d} ({acc[0] * 16:6d} B)  warp-instr {acc[1]:8d}  samples {acc[2]:4d}  no_inst {acc[3]:3d}")
    byfile = collections.defaultdict(lambda: [0, 0, 0, 0])
    for (f, l), v in per.items():
        for k in range(4):
            byfile[f][k] += v[k]
    print("by file:")
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:40s} footprint {v[0]:5d}  warp-instr {v[1]:8d}  samples {v[2]:4d}  no_inst {v[3]:3d}")
    print("top source lines by footprint:")
    for (f, l), v in sorted(per.items(), key=lambda kv: -kv[1][0])[:60]:
        print(f"  {f}:{l:5d}  footprint {v[0]:4d}  warp-instr {v[1]:7d}  samples {v[2]:3d}  no_inst {v[3]:2d}")

main()
