cd $GRAFT_REPO_ROOT
PDM_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-config4 > gpurun_out/r2_bench_n2_e2e.json 2> gpurun_out/r2_bench_n2_e2e.err; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2_e2e.json") if l.startswith("{")][0])
    print("N=2 value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", d["e2e"], "parity", d["parity"]["ok"], "stages", d["stages"])
except Exception as e:
    print("ERR", e)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n2_e2e.err | tail -3 | cut -c1-300
