cd $GRAFT_REPO_ROOT
PYDEM_B200_TS_DEBUG=2 timeout 600 python scripts/sweep_ab.py 4096 tile=0 tile=3 tile=1 > gpurun_out/r2_ab7.log 2>&1; grep -E '^\{' gpurun_out/r2_ab7.log; grep "CTA-time" gpurun_out/r2_ab7.log | awk 'NR%16==2'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tsweep -s 1 -c 1 -o gpurun_out/r2_ts_raw7 python scripts/profile_target.py 4096 2 0 raw > gpurun_out/ncu_ts_raw7.log 2>&1; tail -2 gpurun_out/ncu_ts_raw7.log
