cd $GRAFT_REPO_ROOT
timeout 900 python scripts/sweep_ab.py 4096 sweep=worklist sweep=worklist,lib=pydem_b200/libpydem_b200_nopf.so sweep=worklist sweep=worklist,lib=pydem_b200/libpydem_b200_nopf.so > gpurun_out/r2_ab41.log 2>&1
grep -E '^\{|rror|assert|Trace' gpurun_out/r2_ab41.log | cut -c1-330
