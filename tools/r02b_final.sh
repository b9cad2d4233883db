cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/r02b_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02b_gpu_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02b_smoke.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench_c1_n1.json 2> gpurun_out/r02b_bench_c1_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02b_bench_c1_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "hot us", d.get("p50_optimize_us_hot_l2"), "e2e", d["e2e"]["value"], d["e2e"]["p50_latency_us"], "cpu", d.get("cpu_baseline",{}).get("value"))
for k in ("roofline_large_batch","c4_sharded","c5_fleet","reference_harness"):
    if k in d: print(k, json.dumps(d[k])[:400])
PY
