cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -3 gpurun_out/r2_bench_n1.err; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_n1.json") if l.startswith("{")][0])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "all", d["e2e"]["all_outputs"]["value"], "parity", d["parity"]["ok"], "roofline", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], "launches", d["gpu_launches"])
print({k: (round(v["ms"], 3), round(v["frac_of_hbm_peak"], 3)) for k, v in d["per_stage"].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_4096.csv python bench.py --steps 2 --warmup 1 --no-parity > gpurun_out/r2_b_ncu.log 2>&1; tail -2 gpurun_out/r2_b_ncu.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_worklist|k_slopes_tiled|k_links|k_indeg|k_twi|k_uca_finalize|k_flats_extend|k_ccl" -s 12 -c 12 -o gpurun_out/r2_n1_kernels python scripts/profile_target.py 4096 3 1 cond > gpurun_out/ncu_n1_kernels.log 2>&1; tail -2 gpurun_out/ncu_n1_kernels.log
