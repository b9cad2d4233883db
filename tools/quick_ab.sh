# A/B on one box: kernel A at 2^20 x 56, graph replay against plain launches end to end, the baseline-shape parity tests
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
for i in 1 2; do python tools/large_batch.py 10 2>&1 | tail -1; done
python -m pytest tests/test_gpu_baseline_shapes.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for g in 0 1 0 1; do echo "== MPPI_B200_NO_GRAPH=$g"; MPPI_B200_NO_GRAPH=$g python tools/e2e_timers.py 2>&1 | tail -2; done
