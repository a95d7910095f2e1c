cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err; tail -2 gpurun_out/r2_bench_n1_final.err | cut -c1-300; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n1_final.json") if l.startswith("{")][0])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), "all", round(d["e2e"]["all_outputs"]["value"]), "parity", d["parity"]["ok"], d["parity"]["uca_rel"], "sweep ms", round(d["roofline"]["ms_per_launch"],3), "frac", round(d["roofline"]["frac"],4), "cpu", round(d["cpu_baseline"]["value"],2), "launches", d["gpu_launches"], "clocks", d["clocks"])
print({k: (round(v["ms"], 3), round(v["frac_of_hbm_peak"], 3)) for k, v in d["per_stage"].items()})
PY
