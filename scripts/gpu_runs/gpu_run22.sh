cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -4
timeout 300 python scripts/stencil_ab.py 4096 stencil=v1 stencil=tiled 2>&1 | tail -6
timeout 300 python scripts/stencil_ab.py 4095 stencil=v1 stencil=tiled 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_slopes --csv python scripts/profile_target.py 4096 2 0 raw 2>&1 | grep -i "k_slopes" | cut -c1-200 | tail -3
