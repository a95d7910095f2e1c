cd $GRAFT_REPO_ROOT
timeout 300 python tests/tools/gpu_check.py > gpurun_out/r2_check9.log 2>&1; grep -c "^OK" gpurun_out/r2_check9.log; grep -v "^OK" gpurun_out/r2_check9.log | tail -5
PYDEM_B200_TS_DEBUG=2 timeout 900 python scripts/sweep_ab.py 4096 legacy=1 tile=0 tile=3 tile=1 tile=4 tile=2 tile=13 tile=7 > gpurun_out/r2_ab9.log 2>&1; grep -E '^\{|^cond|^raw' gpurun_out/r2_ab9.log; grep "CTA-time\|river" gpurun_out/r2_ab9.log | awk 'NR%32==3 || NR%32==4'
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_benchmark_regime.py -x -q 2>&1 | tail -5
