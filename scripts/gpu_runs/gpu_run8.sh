cd $GRAFT_REPO_ROOT
PYDEM_B200_TS_DEBUG=2 timeout 600 python scripts/sweep_ab.py 4096 tile=0 tile=6 tile=0,tocc=1 > gpurun_out/r2_ab8.log 2>&1; grep -E '^\{' gpurun_out/r2_ab8.log; grep "CTA-time\|river" gpurun_out/r2_ab8.log | awk 'NR%32==3 || NR%32==4'
