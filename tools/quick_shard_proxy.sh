# one rank's share of the 8-GPU sharded batch (2^20 / 8 trajectories) on one GPU: cycle time and per-kernel durations
# with warm caches (ncu --cache-control none), to see what does not shrink with the shard
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
for b in 131072 262144 1048576; do python tools/large_batch.py 10 $b 2>&1 | tail -1; done
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 120 --csv --log-file gpurun_out/proxy_launches_131072.csv python tools/large_batch.py 5 131072 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/proxy_launches_131072.csv")) if len(r)>5]
hdr=rows[0]; ik,iv=hdr.index("Kernel Name"),hdr.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[1:]:
    try: d[r[ik]].append(float(r[iv].replace(",","")))
    except ValueError: pass
for k,v in sorted(d.items(), key=lambda kv:-sum(kv[1])):
    tail=v[len(v)//2:]
    print(f"{k[:80]:80s} n={len(v):3d} mean_us={sum(v)/len(v)/1e3:8.2f} last-half mean_us={sum(tail)/len(tail)/1e3:8.2f}")
PY
for t in 4 8 16; do echo "== C5 fleet, MPPI_B200_HOST_THREADS=$t"; MPPI_B200_HOST_THREADS=$t MPPI_B200_HOST_TIMERS=1 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --no-large-batch 2>&1 | grep -o '"e2e": {[^}]*}\|host timers[^$]*' | cut -c1-400; done
