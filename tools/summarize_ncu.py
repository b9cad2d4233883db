#!/usr/bin/env python
"""Turns the ncu artefacts of a gpurun call (gpurun_out/*.csv, *.ncu-rep) into the small text / JSON
summaries committed under profiles/.  Usage: python tools/summarize_ncu.py <round-tag>"""
import csv
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

RAW_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def launches(tag, name):
    path = os.path.join(SRC, name)
    if not os.path.exists(path):
        return None
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        try:
            d[r[ik]].append(float(r[iv].replace(",", "")))
        except ValueError:
            pass
    total = sum(sum(v) for v in d.values())
    lines = [f"# {name}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)",
             f"{'kernel':90s} {'launches':>8s} {'mean_us':>10s} {'share':>7s}"]
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        lines.append(f"{k[:90]:90s} {len(v):8d} {sum(v) / len(v) / 1e3:10.2f} {100 * sum(v) / total:6.1f}%")
    open(os.path.join(OUT, f"{tag}_{name.replace('.csv', '')}_summary.txt"), "w").write("\n".join(lines) + "\n")
    return lines


def raw(tag, rep):
    path = os.path.join(SRC, rep)
    if not os.path.exists(path):
        return None
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    res = []
    for r in rows[2:]:
        entry = {"kernel": r[ik]}
        for m in RAW_METRICS:
            if m in hdr:
                entry[m] = f"{r[hdr.index(m)]} {units[hdr.index(m)]}".strip()
        res.append(entry)
    json.dump(res, open(os.path.join(OUT, f"{tag}_{rep.replace('.ncu-rep', '')}_raw.json"), "w"), indent=1)
    return res


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    for f in sorted(os.listdir(SRC)):
        if f.startswith(tag) and f.endswith(".csv"):
            for ln in launches(tag, f) or []:
                print(ln)
        if f.startswith(tag) and f.endswith(".ncu-rep"):
            for e in raw(tag, f) or []:
                print(json.dumps(e))
        if f.startswith(tag) and f.endswith(".txt"):
            open(os.path.join(OUT, f), "w").write(open(os.path.join(SRC, f)).read())
