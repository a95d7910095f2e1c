cd $GRAFT_REPO_ROOT
timeout 300 python tests/tools/gpu_check.py > gpurun_out/r2_check21.log 2>&1; grep -c "^OK" gpurun_out/r2_check21.log; grep -v "^OK" gpurun_out/r2_check21.log | tail -4
PYDEM_B200_TS_DEBUG=45 timeout 900 python scripts/sweep_ab.py 4096 sweep=worklist sweep=tile,tile=1 sweep=tile,tile=0 sweep=tile,tile=2 sweep=tile,tile=4 sweep=tile,tile=6 > gpurun_out/r2_ab21.log 2>&1; grep -E '^\{|rror|^cond|^raw' gpurun_out/r2_ab21.log | cut -c1-250; grep "late visits\|\[ts\] kernel" gpurun_out/r2_ab21.log | awk 'NR%128>=9 && NR%128<=11' | cut -c1-300
