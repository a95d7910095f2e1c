cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/run_mosaic_resident.py 8192 4 2 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -4 | cut -c1-1500
