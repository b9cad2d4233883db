cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
for i in 1 2; do python tools/large_batch.py 10 2>&1 | tail -1; done
python tools/large_batch.py 10 131072 2>&1 | tail -1
python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py tests/test_gpu_critics.py -x -q -m gpu 2>&1 | tail -3
python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --no-large-batch 2>/dev/null | grep -o '"ms_per_step": [0-9.]*'
