cd $GRAFT_REPO_ROOT
timeout 600 python scripts/stencil_ab.py 4096 stencil=tiled9 stencil=tiled10 stencil=tiled11 stencil=tiled12 stencil=tiled13 stencil=tiled14 2>&1 | tail -11 | grep -v identical
