cd $GRAFT_REPO_ROOT
timeout 600 python scripts/stress.py 2048 300 cond,raw 0 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/dist_check.py 4096 2048 pits 100 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$\|UserWarning\|warnings.warn" | tail -4 | cut -c1-500
