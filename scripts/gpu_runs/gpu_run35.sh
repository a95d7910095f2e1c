cd $GRAFT_REPO_ROOT
timeout 300 python scripts/c_abi_shard_check.py 2 2048 1024 2>&1 | grep -v "^$" | tail -6 | cut -c1-400
timeout 300 python scripts/c_abi_shard_check.py 2 2048 1024 pits 2>&1 | grep -v "^$" | tail -6 | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
