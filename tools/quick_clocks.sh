# phase clocks of the fused kernel with and without debug flags, plus the fused-path GPU tests
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
TAG=${TAG:-qc}
python -m pytest tests/test_gpu_parity.py tests/test_host_mirror.py -x -q -m gpu > gpurun_out/${TAG}_tests.log 2>&1; tail -4 gpurun_out/${TAG}_tests.log
python tools/phase_clocks.py > gpurun_out/${TAG}_clocks_plain.txt 2>&1; head -36 gpurun_out/${TAG}_clocks_plain.txt | cut -c1-100
for f in ${FLAGS:-1}; do
MPPI_B200_DEBUG_FLAGS=$f python tools/phase_clocks.py > gpurun_out/${TAG}_clocks_flags$f.txt 2>&1; echo "== flags $f"; sed -n 20,36p gpurun_out/${TAG}_clocks_flags$f.txt | cut -c1-100
done
python tools/e2e_timers.py > gpurun_out/${TAG}_e2e.txt 2>&1; tail -12 gpurun_out/${TAG}_e2e.txt
