cd $GRAFT_REPO_ROOT
PDM_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8_wl.json 2> gpurun_out/r2_bench_n8_wl.err; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n8_wl.json") if l.startswith("{")][0])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"]["ok"], "stages", d["stages"])
    print("config4", json.dumps(d.get("config4"))[:1500])
except Exception as e:
    print("ERR", e)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n8_wl.err | tail -8 | cut -c1-300
