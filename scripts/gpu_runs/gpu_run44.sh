cd $GRAFT_REPO_ROOT
PDM_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8_final.json 2> gpurun_out/r2_bench_n8_final.err; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n8_final.json") if l.startswith("{")][0])
    print("N=8 value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "parity", d["parity"]["ok"], "stages", d["stages"])
    print("config4", round(d["config4"]["value"]), d["config4"]["ms_per_step"], d["config4"]["stages"], d["config4"]["checks"]["ok"])
except Exception as e:
    print("ERR", e)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n8_final.err | tail -3 | cut -c1-300
