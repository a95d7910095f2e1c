cd $GRAFT_REPO_ROOT
timeout 900 python scripts/sweep_ab.py 4096 sweep=worklist,burst=1 sweep=worklist,burst=0 > gpurun_out/r2_ab38.log 2>&1
grep -E '^\{|rror|assert|Trace|^cond|^raw' gpurun_out/r2_ab38.log | cut -c1-330
