# asynchronous cluster exchanges of the fused kernel: parity of the single-launch path, phase clocks, a short bench
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
TAG=${TAG:-xch}
timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_shapes.py -x -q -m gpu -k "c1 or lean or open_loop or fleet or fallback or window or non_shift" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
timeout 60 python tools/phase_clocks.py > gpurun_out/${TAG}_clocks.txt 2>&1; echo "clocks rc=$?"; head -33 gpurun_out/${TAG}_clocks.txt | cut -c1-100
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-large-batch > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json,os
d=json.loads(open(f"gpurun_out/{os.environ.get('TAG','xch')}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "hot us", d.get("p50_optimize_us_hot_l2"), "e2e us", d["e2e"]["p50_latency_us"])
PY
