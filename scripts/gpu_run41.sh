cd $GRAFT_REPO_ROOT
timeout 900 python scripts/sweep_ab.py 4096 sweep=worklist,burst=1 sweep=worklist,burst=0 sweep=worklist,bprof=1,lib=pydem_b200/libpydem_b200_prof.so > gpurun_out/r2_ab41.log 2>&1
grep -E '^\{|rror|assert|Trace|^cond|^raw' gpurun_out/r2_ab41.log | cut -c1-330
grep '^\[bp\]' gpurun_out/r2_ab41.log | head -2 | cut -c1-400
