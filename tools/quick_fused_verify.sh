# everything that runs the single-launch (fused cluster) kernel: parity, host mirror, device costmap hand-off, smoke
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
TAG=${TAG:-fv}
timeout 120 python -m pytest tests/test_gpu_parity.py tests/test_host_mirror.py tests/test_gpu_costmap_device.py tests/test_gpu_baseline_shapes.py -x -q -m gpu -k "not c4_million and not c3_full_shape and not 65536 and not c5_fleet_general" > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${TAG}_tests.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/${TAG}_smoke.log
