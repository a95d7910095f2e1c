cd $GRAFT_REPO_ROOT
timeout 300 python tests/tools/gpu_check.py > gpurun_out/r2_check25.log 2>&1; grep -c "^OK" gpurun_out/r2_check25.log; grep -v "^OK" gpurun_out/r2_check25.log | tail -4
PYDEM_B200_TS_DEBUG=2 timeout 900 python scripts/sweep_ab.py 4096 sweep=worklist sweep=tile,tile=1 sweep=tile,tile=0 sweep=tile,tile=2 sweep=tile,tile=4 > gpurun_out/r2_ab25.log 2>&1; grep -E '^\{|rror|^cond|^raw' gpurun_out/r2_ab25.log | cut -c1-250; grep "\[ts\] kernel" gpurun_out/r2_ab25.log | awk 'NR%32==5 || NR%32==6' | cut -c1-200
grep timeline gpurun_out/r2_ab25.log | sed -n 3p | cut -c1-900
