cd $GRAFT_REPO_ROOT
nvidia-smi topo -m 2>&1 | head -8
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py 2048 1024 2>&1 | tail -5
PYDEM_B200_SHARD_P2P=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_check.py 2048 1024 2>&1 | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_p2p.json 2> gpurun_out/r2_bench_n2_p2p.err; tail -c 2500 gpurun_out/r2_bench_n2_p2p.json; tail -5 gpurun_out/r2_bench_n2_p2p.err
PYDEM_B200_SHARD_P2P=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_rounds.json 2> gpurun_out/r2_bench_n2_rounds.err; tail -c 1500 gpurun_out/r2_bench_n2_rounds.json; tail -5 gpurun_out/r2_bench_n2_rounds.err
