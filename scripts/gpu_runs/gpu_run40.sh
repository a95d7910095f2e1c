cd $GRAFT_REPO_ROOT
timeout 900 python scripts/sweep_ab.py 4096 sweep=worklist,burst=1 sweep=worklist,burst=0 > gpurun_out/r2_ab40.log 2>&1
grep -E '^\{|rror|assert|Trace|^cond|^raw' gpurun_out/r2_ab40.log | cut -c1-330
for b in 1; do
PYDEM_B200_SWEEP_BURST=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$b bench.py --gpus 2 --steps 5 --warmup 3 --no-parity > gpurun_out/r2_bench_n2_b$b.json 2> gpurun_out/r2_bench_n2_b$b.err; python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2_bench_n2_b$b.json") if l.startswith("{")][0])
    print("burst=$b N=2 value", round(d["value"]), "ms", round(d["ms_per_step"],3), "sweep", d["stages"]["ms_sweep_first"])
    print("config4", round(d["config4"]["value"]), round(d["config4"]["ms_per_step"],2), d["config4"]["stages"]["ms_sweep_first"], d["config4"]["stages"]["n_queue_items"])
except Exception as e:
    print("ERR", e)
PY
done
