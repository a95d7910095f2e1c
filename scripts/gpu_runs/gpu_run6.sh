cd $GRAFT_REPO_ROOT
PYDEM_B200_TS_DEBUG=2 timeout 600 python scripts/sweep_ab.py 4096 tile=0 tile=3 > gpurun_out/r2_ab6.log 2>&1; grep -E '^\{' gpurun_out/r2_ab6.log; grep "CTA-time" gpurun_out/r2_ab6.log | awk 'NR%16==2'
