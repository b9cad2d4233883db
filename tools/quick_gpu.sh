# quick GPU check: the GPU test suite, phase clocks of the fused kernel, a short bench without the CPU legs
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
TAG=${TAG:-quick}
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -15 gpurun_out/${TAG}_gpu_tests.log
PHASE_CLOCK_CYCLES=15 python tools/phase_clocks.py > gpurun_out/${TAG}_phase_clocks.txt 2>&1; head -36 gpurun_out/${TAG}_phase_clocks.txt | cut -c1-110
python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo bench rc=$?
python - <<'PY'
import json,os
d=json.loads(open(f"gpurun_out/{os.environ.get('TAG','quick')}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "hot us", d.get("p50_optimize_us_hot_l2"), "e2e us", d["e2e"]["p50_latency_us"], "two_call", d["e2e"].get("two_call",{}).get("p50_latency_us"))
for k in ("roofline_large_batch","c4_sharded","c5_fleet"):
    if k in d: print(k, {kk: d[k][kk] for kk in d[k] if kk in ("kernel_ms","frac","step_ms","ms_per_cycle","device_ms_per_fleet_cycle","e2e_ms_per_fleet_cycle","parity")})
PY
