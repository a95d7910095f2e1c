"""Design study for the multi-GPU path: how many exchange rounds does the
row-sharded sweep need?  Test tooling (takes the drainage graph from the oracle).  Simulates the
shipped protocol structurally: every shard drains to quiescence, cross-boundary pushes are delivered
between rounds; rounds = 1 + the largest number of shard-boundary crossings of a flow path.

    python tests/tools/proto_shard_rounds.py [n=4096] [periodic_blocks=1] [shards: 2 4 8]
periodic_blocks = k stacks the row-periodic benchmark block k times (the weak-scaling workload of
bench.py); the measured 14 rounds at 2 and 4 GPUs are the check of this simulator."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from pydem_b200 import synth
from oracle.oracle import OracleDEMProcessor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 1
shards = [int(a) for a in sys.argv[3:]] or [2, 4, 8]
B = synth.conditioned_fractal_dem(n, 0, wrap_rows=(blocks > 1))
E = np.vstack([B] * blocks)
R, C = E.shape
dp = OracleDEMProcessor(E, dX=30.0, dY=30.0, fill_flats=False, drain_pits_path=False, drain_pits=False)
dp.calc_slopes_directions()
g, sec = dp._graph()
cptr, cidx, cdat, rptr, ridx = g.export()
N = R * C
indeg0 = (rptr[1:] - rptr[:-1]).astype(np.int64)
print("DEM %dx%d (%d periodic block(s) of %d rows), %d edges, %d sources" % (R, C, blocks, n, cidx.size, int((indeg0 == 0).sum())))
src_of_edge = np.repeat(np.arange(N, dtype=np.int64), np.diff(cptr))
for S in shards:
    rows = -(-R // S)
    shard_of = (np.arange(N, dtype=np.int64) // C) // rows
    cross = shard_of[src_of_edge] != shard_of[cidx]            # edges that cross a shard boundary
    # crossings-so-far of the best (max) path into each cell, by Kahn order with NumPy frontiers
    indeg = indeg0.copy()
    depth = np.zeros(N, np.int32)
    frontier = np.nonzero(indeg == 0)[0]
    t0 = time.time()
    while frontier.size:
        # expand all out-edges of the frontier
        lo, hi = cptr[frontier], cptr[frontier + 1]
        cnt = hi - lo
        if cnt.sum() == 0:
            break
        eidx = np.repeat(lo, cnt) + (np.arange(cnt.sum()) - np.repeat(np.cumsum(cnt) - cnt, cnt))
        srcs = np.repeat(frontier, cnt)
        dst = cidx[eidx]
        d = depth[srcs] + cross[eidx].astype(np.int32)
        np.maximum.at(depth, dst, d)
        np.subtract.at(indeg, dst, 1)
        u = np.unique(dst)
        frontier = u[indeg[u] == 0]
    print("%d shards of %d rows: %d boundary-crossing edges, max crossings on a flow path %d -> %d rounds  [%.0f s]"
          % (S, rows, int(cross.sum()), int(depth.max()), int(depth.max()) + 1, time.time() - t0))
