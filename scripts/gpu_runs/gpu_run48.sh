cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 10 --config4 --no-parity > gpurun_out/r2_bench_n1_config4.json 2> gpurun_out/r2_bench_n1_config4.err; tail -3 gpurun_out/r2_bench_n1_config4.err | cut -c1-300; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n1_config4.json") if l.startswith("{")][0])
print("value", round(d["value"]), "config4", json.dumps(d.get("config4"))[:1200])
PY
