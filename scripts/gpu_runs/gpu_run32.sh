cd $GRAFT_REPO_ROOT
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/dist_check.py 2048 1024 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -6 | cut -c1-600
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 scripts/dist_check.py 4096 4096 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -6 | cut -c1-600
PDM_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_wl.json 2> gpurun_out/r2_bench_n2_wl.err; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2_wl.json") if l.startswith("{")][0])
    print("value", d["value"], "ms", d["ms_per_step"], "parity", d["parity"]["ok"], "stages", d["stages"])
    print("config4", json.dumps(d.get("config4"))[:1500])
except Exception as e:
    print("ERR", e)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n2_wl.err | tail -12 | cut -c1-300
