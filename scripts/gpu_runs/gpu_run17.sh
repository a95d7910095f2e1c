cd $GRAFT_REPO_ROOT
PYDEM_B200_SWEEP=tile PYDEM_B200_TS_TILE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tsweep -s 1 -c 1 -o gpurun_out/r2_ts_cond17 python scripts/profile_target.py 4096 2 1 cond > gpurun_out/ncu_ts_cond17.log 2>&1; tail -3 gpurun_out/ncu_ts_cond17.log
