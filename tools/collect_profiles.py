#!/usr/bin/env python
"""Turns what tools/r02_run.sh left in a directory (default gpurun_out/) into the summaries committed under profiles/:
bench lines, the ncu launch list as shares per kernel, the `ncu --set full` raw pages as a few metrics per kernel,
dram traffic per launch for bench.py's `roofline.traffic`, the executed code footprint of the fused kernel and the
phase clocks.  Usage: python tools/collect_profiles.py <tag> [dir]"""
import csv
import gzip
import json
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")

RAW_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "smsp__average_warp_latency_per_inst_issued.ratio",
]


def launches(src, tag, name):
    rows = [r for r in csv.reader(open(os.path.join(src, name))) if len(r) > 5]
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
    open(os.path.join(OUT, name.replace(".csv", "_summary.txt")), "w").write("\n".join(lines) + "\n")


def raw(src, name):
    rows = list(csv.reader(open(os.path.join(src, name))))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        e = {"kernel": r[hdr.index("Kernel Name")]}
        for m in RAW_METRICS:
            if m in hdr:
                e[m] = f"{r[hdr.index(m)]} {units[hdr.index(m)]}".strip()
        res.append(e)
    json.dump(res, open(os.path.join(OUT, name.replace(".csv", ".json")), "w"), indent=1)
    return res


def num(s):
    v, _, unit = s.partition(" ")
    scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(unit.strip(), 1.0)
    return float(v) * scale


def main():
    tag = sys.argv[1]
    src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
    os.makedirs(OUT, exist_ok=True)
    traffic = {"_source": f"ncu --set full --clock-control none captures of {tag} (profiles/{tag}_fused_c1_raw.json, "
                          f"profiles/{tag}_kernels_c4_raw.json): dram__bytes_read.sum + dram__bytes_write.sum per launch"}
    for f in sorted(os.listdir(src)):
        if not f.startswith(tag):
            continue
        p = os.path.join(src, f)
        if f.endswith("_raw.csv"):
            for e in raw(src, f):
                k = e["kernel"]
                key = None
                if "k_fused_cluster" in k:
                    key = "k_fused_cluster@C1"
                elif "k_rollout" in k:
                    key = "k_rollout_score@2^20x56"
                elif "k_softmax_partial" in k:
                    key = "k_softmax_partial@2^20x56"
                if key and key not in traffic:
                    traffic[key] = {"kernel": k, "dram_bytes_read": int(num(e["dram__bytes_read.sum"])),
                                    "dram_bytes_write": int(num(e["dram__bytes_write.sum"]))}
        elif f.endswith(".csv"):
            launches(src, tag, f)
        elif f.endswith((".txt", ".json")) and os.path.getsize(p) > 0:
            shutil.copy(p, os.path.join(OUT, f))
    if len(traffic) > 1:
        old = os.path.join(OUT, "r01_traffic.json")
        if os.path.exists(old):      # the inflation captures are not re-taken every round
            for k, v in json.load(open(old)).items():
                if k.startswith("k_inflate"):
                    traffic[k] = v
        json.dump(traffic, open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)
    # executed code footprint of the fused kernel: needs the cubin of the SAME build that ran
    srcz = os.path.join(src, f"{tag}_fused_c1_source.csv.gz")
    lib = os.path.join(ROOT, "navigation2_b200", "libmppi_b200.so")
    if os.path.exists(srcz) and os.path.exists(lib):
        tmp = "/tmp/collect_sass"
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp)
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, "mppi_kernels.sm_100a.cubin")], capture_output=True, text=True).stdout
        lines = sass.splitlines()
        starts = [i for i, ln in enumerate(lines) if ln.startswith(".text.")]
        for n, i in enumerate(starts):
            if "k_fused_clusterILb0" in lines[i]:
                end = starts[n + 1] if n + 1 < len(starts) else len(lines)
                open(os.path.join(tmp, "fused.sass"), "w").write("\n".join(lines[i:end]) + "\n")
        rng = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_footprint.py"), srcz, os.path.join(tmp, "fused.sass")],
                             capture_output=True, text=True)
        if rng.returncode == 0:
            open(os.path.join(OUT, f"{tag}_fused_c1_code_footprint.txt"), "w").write(
                "# executed code footprint of k_fused_cluster<false> at C1 (tools/sass_footprint.py: ncu SASS page x nvdisasm -g)\n" +
                "\n".join(rng.stdout.splitlines()[:45]) + "\n")
        else:
            print("footprint join failed:", rng.stderr[-400:])


if __name__ == "__main__":
    main()
