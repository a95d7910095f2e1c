cd $GRAFT_REPO_ROOT
timeout 300 python tests/tools/gpu_check.py > gpurun_out/r2_check13.log 2>&1; grep -c "^OK" gpurun_out/r2_check13.log; grep -v "^OK" gpurun_out/r2_check13.log | tail -8
PYDEM_B200_TS_DEBUG=2 timeout 900 python scripts/sweep_ab.py 4096 legacy=1 tile=0 tile=1 tile=2 tile=3 tile=4 tile=5 tile=6 tile=7 > gpurun_out/r2_ab13.log 2>&1; grep -E '^\{|^cond|rror' gpurun_out/r2_ab13.log; grep "CTA-time\|\[ts\] kernel" gpurun_out/r2_ab13.log | awk 'NR%32==3 || NR%32==4'
grep timeline gpurun_out/r2_ab13.log | sed -n 3p
