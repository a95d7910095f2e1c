cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_4096_final.csv python bench.py --steps 2 --warmup 1 --no-parity > gpurun_out/r2_b_ncu.log 2>&1; tail -2 gpurun_out/r2_b_ncu.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_worklist|k_slopes_tiled|k_links|k_indeg|k_twi|k_uca_finalize|k_flats_extend|k_ccl" -s 12 -c 12 -o gpurun_out/r2_n1_kernels_final python scripts/profile_target.py 4096 3 1 cond > gpurun_out/ncu_n1_kernels.log 2>&1; tail -2 gpurun_out/ncu_n1_kernels.log
ls -la gpurun_out/*.ncu-rep | tail -3
