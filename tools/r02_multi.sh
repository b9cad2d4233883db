# Round-2 multi-GPU pass on N GPUs of one box (gpurun --gpus N): the world-2 sharded parity tests and the bench
# line (headline + c4_sharded + c5_fleet sections) at every N up to the GPUs visible.
set -x
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
TAG=${TAG:-r02}
NG=$(python -c "import torch; print(torch.cuda.device_count())")
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests/test_gpu_sharded.py -x -q -m gpu > gpurun_out/${TAG}_gpu_sharded_tests_n${NG}.log 2>&1; tail -5 gpurun_out/${TAG}_gpu_sharded_tests_n${NG}.log; fi
for N in ${NS:-2 4 8}; do
  if [ $N -le $NG ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c1_n${N}.json 2> gpurun_out/${TAG}_bench_c1_n${N}.err; echo "bench N=$N rc=$?"
    tail -c 600 gpurun_out/${TAG}_bench_c1_n${N}.err
  fi
done
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/*_bench_c1_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    c4, c5 = d.get("c4_sharded", {}), d.get("c5_fleet", {})
    print(os.path.basename(f), "value %.3e" % d["value"], "e2e %.3e" % d["e2e"]["value"],
          "| c4 ms/cycle", c4.get("ms_per_cycle"), "exchange", c4.get("exchange"), "parity", c4.get("parity_checked"),
          "| c5 device %.3e e2e %.3e" % (c5.get("device_value", 0), c5.get("e2e_value", 0)))
PY
