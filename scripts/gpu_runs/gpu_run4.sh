set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tests/tools/gpu_check.py > gpurun_out/r2_check4.log 2>&1; tail -1 gpurun_out/r2_check4.log
PYDEM_B200_TS_DEBUG=1 timeout 1200 python scripts/sweep_ab.py 4096 legacy=1 tile=0 tile=1 tile=3 tile=4 tile=5 tile=6 tile=7 tile=8 tile=9 tile=10 tile=11 tile=12 > gpurun_out/r2_ab4.log 2>&1; grep -E '^\{|^cond|^raw' gpurun_out/r2_ab4.log; grep "ts\]" gpurun_out/r2_ab4.log | awk 'NR%16==2'
