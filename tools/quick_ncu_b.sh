cd $GRAFT_REPO_ROOT
export PYTHONPATH=.
python tools/large_batch.py 5 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_path_score|k_finalize' -c 2 --launch-skip 26 -f -o /tmp/kb python tools/large_batch.py 2 > gpurun_out/kb_ncu.log 2>&1
ncu -i /tmp/kb.ncu-rep --page raw --csv > gpurun_out/kb_raw.csv 2>/dev/null
ncu -i /tmp/kb.ncu-rep --page source --csv --kernel-name regex:k_path_score > gpurun_out/kb_path_source.csv 2>/dev/null
ncu -i /tmp/kb.ncu-rep --page source --csv --kernel-name regex:k_finalize > gpurun_out/kb_fin_source.csv 2>/dev/null
gzip -f gpurun_out/kb_*_source.csv
tail -3 gpurun_out/kb_ncu.log
