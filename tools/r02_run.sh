# Round-2 measurement pass on one B200: GPU tests, smoke, bench (both arms), ncu launch list, full captures of the
# fused kernel (C1) and of the large-batch kernels (2^20 x 56), phase clocks.  The .ncu-rep files are turned into CSV
# pages on the box and removed (gpurun_out/ is capped at 64 MiB).
set -x
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
TAG=${TAG:-r02}
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -5 gpurun_out/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -4 gpurun_out/${TAG}_smoke.log
fi
python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_c1_n1.json 2> gpurun_out/${TAG}_bench_c1_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_reference_n1.json 2> gpurun_out/${TAG}_bench_reference_n1.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused_cluster -c 1 --launch-skip 5 -f -o /tmp/${TAG}_fused_c1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-large-batch > gpurun_out/${TAG}_ncu_fused.log 2>&1
ncu -i /tmp/${TAG}_fused_c1.ncu-rep --page raw --csv > gpurun_out/${TAG}_fused_c1_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_fused_c1.ncu-rep --page source --csv > gpurun_out/${TAG}_fused_c1_source.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_rollout|k_softmax_partial|k_path_score|k_finalize' -c 4 --launch-skip 8 -f -o /tmp/${TAG}_kernels_c4 python tools/large_batch.py 2 > gpurun_out/${TAG}_ncu_c4.log 2>&1
ncu -i /tmp/${TAG}_kernels_c4.ncu-rep --page raw --csv > gpurun_out/${TAG}_kernels_c4_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_kernels_c4.ncu-rep --page source --csv --kernel-name regex:k_rollout > gpurun_out/${TAG}_rollout_c4_source.csv 2>/dev/null
python tools/phase_clocks.py > gpurun_out/${TAG}_phase_clocks_c1.txt 2>&1
python tools/c4_breakdown.py > gpurun_out/${TAG}_c4_breakdown.txt 2>&1
python tools/c3_breakdown.py > gpurun_out/${TAG}_c3_breakdown.txt 2>&1
gzip -f gpurun_out/*_source.csv
ls -la gpurun_out | tail -25
