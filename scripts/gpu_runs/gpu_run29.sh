cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_process_manager.py -x -q -k "resident" 2>&1 | tail -25
