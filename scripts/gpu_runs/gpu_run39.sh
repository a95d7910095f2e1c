cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 30 > gpurun_out/r2_bench_n1_burst.json 2> gpurun_out/r2_bench_n1_burst.err; tail -2 gpurun_out/r2_bench_n1_burst.err | cut -c1-300; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n1_burst.json") if l.startswith("{")][0])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), "all", round(d["e2e"]["all_outputs"]["value"]), "parity", d["parity"]["ok"], d["parity"]["uca_rel"], "sweep ms", round(d["roofline"]["ms_per_launch"],3), "frac", d["roofline"]["frac"])
print({k: (round(v["ms"], 3), round(v["frac_of_hbm_peak"], 3)) for k, v in d["per_stage"].items()})
PY
timeout 300 python scripts/c_abi_shard_check.py 2 2048 1024 pits 2>&1 | tail -3 | cut -c1-300
PDM_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_burst.json 2> gpurun_out/r2_bench_n2_burst.err; python - <<'PY'
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2_burst.json") if l.startswith("{")][0])
    print("N=2 value", d["value"], "ms", d["ms_per_step"], "parity", d["parity"]["ok"], "stages", d["stages"])
    print("config4", d["config4"]["value"], d["config4"]["ms_per_step"], d["config4"]["stages"], d["config4"]["checks"])
except Exception as e:
    print("ERR", e)
PY
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_bench_n2_burst.err | tail -4 | cut -c1-300
