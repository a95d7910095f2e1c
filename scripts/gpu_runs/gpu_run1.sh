set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python tests/tools/gpu_check.py > gpurun_out/r2_check1.log 2>&1; tail -40 gpurun_out/r2_check1.log
PYDEM_B200_TS_DEBUG=1 timeout 900 python scripts/sweep_ab.py 4096 legacy=1 tile=0 tile=1 tile=2 tile=3 > gpurun_out/r2_ab1.log 2>&1; cat gpurun_out/r2_ab1.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; tail -15 gpurun_out/r2_pytest1.log
