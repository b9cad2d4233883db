# full ncu capture of kernel A at 2^20 x 56 (steady state) with the SASS page for stall analysis
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rollout -c 1 --launch-skip 14 -f -o /tmp/ka python tools/large_batch.py 2 > gpurun_out/ka_ncu.log 2>&1
ncu -i /tmp/ka.ncu-rep --page raw --csv > gpurun_out/ka_raw.csv 2>/dev/null
ncu -i /tmp/ka.ncu-rep --page source --csv > gpurun_out/ka_source.csv 2>/dev/null
gzip -f gpurun_out/ka_source.csv
tail -2 gpurun_out/ka_ncu.log
