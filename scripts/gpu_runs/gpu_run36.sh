cd $GRAFT_REPO_ROOT
timeout 600 python scripts/sweep_ab.py 4096 sweep=worklist,dbg=1 > gpurun_out/r2_wl_dbg.log 2>&1
grep -n '^{' gpurun_out/r2_wl_dbg.log | cut -c1-300
grep '^\[wl\]' gpurun_out/r2_wl_dbg.log | head -4 | cut -c1-500
