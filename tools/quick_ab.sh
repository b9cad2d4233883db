# A/B of the opt-in switches on one box: GPU tests, zero-copy on/off end to end, the 2^20 x 56 batch
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
TAG=${TAG:-ab}
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; tail -4 gpurun_out/${TAG}_tests.log
for z in 0 1 0 1; do
echo "== MPPI_B200_ZERO_COPY=$z"; MPPI_B200_ZERO_COPY=$z python tools/e2e_timers.py 2>&1 | tail -4
done
MPPI_B200_ZERO_COPY=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
python tools/large_batch.py 10 2>&1 | tail -2
python tools/c4_breakdown.py 2>&1 | tail -6
