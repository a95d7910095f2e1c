set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tests/tools/gpu_check.py > gpurun_out/r2_check2.log 2>&1; tail -3 gpurun_out/r2_check2.log
PYDEM_B200_TS_DEBUG=1 timeout 900 python scripts/sweep_ab.py 4096 legacy=1 tile=0 tile=1 tile=2 tile=3 tile=0,tocc=3 > gpurun_out/r2_ab2.log 2>&1; grep -E '^\{|^cond|^raw' gpurun_out/r2_ab2.log; grep "ts\]" gpurun_out/r2_ab2.log | awk 'NR%16==2'
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_benchmark_regime.py -x -q > gpurun_out/r2_pytest2.log 2>&1; tail -15 gpurun_out/r2_pytest2.log
