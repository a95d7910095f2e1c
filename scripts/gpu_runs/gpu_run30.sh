cd $GRAFT_REPO_ROOT
timeout 300 python scripts/run_mosaic_resident.py --cond 4096 2>&1 | tail -2 | cut -c1-1200
timeout 300 python scripts/run_mosaic_resident.py --cond 4096 --host 2>&1 | tail -2 | cut -c1-1200
timeout 300 python scripts/run_mosaic_resident.py 2048 4 2 2>&1 | tail -2 | cut -c1-1200
