cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n1_final.json") if l.startswith("{")][0])
print("N=1 value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "all", round(d["e2e"]["all_outputs"]["value"]), "parity", d["parity"]["ok"], "sweep ms", round(d["roofline"]["ms_per_launch"],3), "frac", round(d["roofline"]["frac"],4))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2_final.json") if l.startswith("{")][0])
    print("N=2 value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "parity", d["parity"]["ok"], "sweep", d["stages"]["ms_sweep_first"], "config4", round(d["config4"]["value"]), d["config4"]["checks"]["ok"])
except Exception as e:
    print("ERR", e)
PY
