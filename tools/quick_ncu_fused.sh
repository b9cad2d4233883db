# warm-cache SASS-level stall samples of the fused kernel, accumulated over several launches (steady state)
cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
timeout 400 ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_fused_cluster -c 12 --launch-skip 30 -f -o /tmp/kf python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-large-batch > gpurun_out/kf_ncu.log 2>&1
ncu -i /tmp/kf.ncu-rep --page source --csv > gpurun_out/kf_source.csv 2>/dev/null
ncu -i /tmp/kf.ncu-rep --page raw --csv > gpurun_out/kf_raw.csv 2>/dev/null
gzip -f gpurun_out/kf_source.csv
tail -2 gpurun_out/kf_ncu.log
