cd $GRAFT_REPO_ROOT
PYDEM_B200_TS_DEBUG=45 timeout 900 python scripts/sweep_ab.py 4096 tile=1 tile=0 tile=3 tile=4 > gpurun_out/r2_ab14.log 2>&1; grep -E '^\{|rror' gpurun_out/r2_ab14.log | grep cond; grep "late visits\|\[ts\] kernel" gpurun_out/r2_ab14.log | awk 'NR%64==5 || NR%64==6'
