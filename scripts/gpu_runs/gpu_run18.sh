cd $GRAFT_REPO_ROOT
PYDEM_B200_TS_DEBUG=45 timeout 900 python scripts/sweep_ab.py 4096 sweep=tile,tile=1 sweep=tile,tile=0 > gpurun_out/r2_ab18.log 2>&1; grep -E '^\{|rror' gpurun_out/r2_ab18.log | grep cond; grep "late visits\|busy turns\|\[ts\] kernel" gpurun_out/r2_ab18.log | awk 'NR%128>=9 && NR%128<=11'
