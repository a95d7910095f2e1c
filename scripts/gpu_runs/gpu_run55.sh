cd $GRAFT_REPO_ROOT
timeout 250 python tests/tools/run_configs.py --configs 3 --size3 16384 --out gpurun_out/r2_configs.json --timeout 230 2>&1 | tail -5 | cut -c1-600
