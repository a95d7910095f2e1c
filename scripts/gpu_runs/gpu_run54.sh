cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 --no-config4 > gpurun_out/r2_bench_n4_final.json 2> gpurun_out/r2_bench_n4_final.err; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n4_final.json") if l.startswith("{")][0])
    print("N=4 value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "parity", d["parity"]["ok"], "stages", d["stages"])
except Exception as e:
    print("ERR", e)
PY
