cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_benchmark_regime.py tests/test_gpu_sharded.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err; tail -c 1500 gpurun_out/r2_bench_n1_a.json; tail -3 gpurun_out/r2_bench_n1_a.err
