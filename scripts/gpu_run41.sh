cd $GRAFT_REPO_ROOT
timeout 900 python scripts/sweep_ab.py 4096 sweep=worklist,lib=pydem_b200/libpydem_b200_a.so sweep=worklist,lib=pydem_b200/libpydem_b200_b.so sweep=worklist,lib=pydem_b200/libpydem_b200_c.so sweep=worklist,lib=pydem_b200/libpydem_b200_d.so sweep=worklist,lib=pydem_b200/libpydem_b200_e.so sweep=worklist,lib=pydem_b200/libpydem_b200_a.so > gpurun_out/r2_ab41.log 2>&1
grep -E '^\{|rror|assert|Trace' gpurun_out/r2_ab41.log | cut -c1-330
