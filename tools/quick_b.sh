cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
for b in 131072 1048576; do python tools/large_batch.py 10 $b 2>&1 | tail -1; done
python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu-baseline --no-large-batch 2>/dev/null | grep -o '"ms_per_step": [0-9.]*'
python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --no-large-batch 2>/dev/null | grep -o '"ms_per_step": [0-9.]*'
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
