set -x
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
python -m pytest tests -x -q -m gpu > gpurun_out/r01_final_gpu_tests.log 2>&1; tail -3 gpurun_out/r01_final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r01_final_smoke.log 2>&1; tail -4 gpurun_out/r01_final_smoke.log
python bench.py > gpurun_out/r01_plain_bench.json 2> gpurun_out/r01_plain_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_ref_n1.json 2> gpurun_out/final_ref_n1.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file gpurun_out/r01_launches_inflation.csv python tools/inflation_profile.py > gpurun_out/r01_ncu_inflation_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inflate -f -o gpurun_out/r01_inflate python tools/inflation_profile.py > gpurun_out/r01_ncu_inflate.log 2>&1
