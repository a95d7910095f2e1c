cd $GRAFT_REPO_ROOT
PYDEM_B200_TS_DEBUG=2 timeout 600 python scripts/sweep_ab.py 4096 tile=0 tile=1 tile=11 > gpurun_out/r2_ab5.log 2>&1; grep -E '^\{' gpurun_out/r2_ab5.log; grep "timeline" gpurun_out/r2_ab5.log | awk 'NR%16==2'
