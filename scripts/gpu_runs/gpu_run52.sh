cd $GRAFT_REPO_ROOT
timeout 900 python scripts/sweep_ab.py 4096 sweep=worklist,backoff=1600 sweep=worklist,backoff=3200 sweep=worklist,backoff=6400 sweep=worklist,backoff=12800 sweep=worklist > gpurun_out/r2_ab52.log 2>&1
grep -E '^\{|rror|assert|Trace' gpurun_out/r2_ab52.log | cut -c1-330
