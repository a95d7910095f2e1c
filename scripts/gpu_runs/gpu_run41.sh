cd $GRAFT_REPO_ROOT
timeout 900 python scripts/sweep_ab.py 4096 sweep=worklist sweep=worklist,lib=pydem_b200/libpydem_b200_norb.so sweep=worklist sweep=worklist,lib=pydem_b200/libpydem_b200_norb.so > gpurun_out/r2_ab41.log 2>&1
grep -E '^\{|rror|assert|Trace|^cond|^raw' gpurun_out/r2_ab41.log | cut -c1-330
