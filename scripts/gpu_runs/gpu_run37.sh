cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in 0 1; do
PYDEM_B200_CHUNKED_UPLOAD=$v timeout 600 python bench.py --steps 30 > gpurun_out/r2_bench_n1_chunk$v.json 2> gpurun_out/r2_bench_n1_chunk$v.err; tail -2 gpurun_out/r2_bench_n1_chunk$v.err | cut -c1-300; python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n1_chunk$v.json") if l.startswith("{")][0])
print("chunk=$v value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), "all", round(d["e2e"]["all_outputs"]["value"]), "parity", d["parity"]["ok"], "sweep ms", round(d["roofline"]["ms_per_launch"],3))
PY
done
