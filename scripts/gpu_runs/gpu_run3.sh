set -x
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tsweep -s 1 -c 1 -o gpurun_out/r2_ts_raw python scripts/profile_target.py 4096 2 0 raw > gpurun_out/ncu_ts_raw.log 2>&1; tail -3 gpurun_out/ncu_ts_raw.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tsweep -s 1 -c 1 -o gpurun_out/r2_ts_cond python scripts/profile_target.py 4096 2 1 cond > gpurun_out/ncu_ts_cond.log 2>&1; tail -3 gpurun_out/ncu_ts_cond.log
ls -la gpurun_out/*.ncu-rep
