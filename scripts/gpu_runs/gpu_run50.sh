cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_sharded.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/dist_check.py 4096 2048 pits 20 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|UserWarning\|warnings.warn" | tail -2 | cut -c1-200
