"""Design study for DESIGN.md section 8 item 1 (flow paths in shared memory): how would a
tile-resident sweep behave on the benchmark DEM?  Test tooling -- it takes the drainage graph from
the oracle (hence under tests/) and simulates, structure only, the schedule

    repeat until nothing is ready:
        every tile that holds ready cells drains, inside the tile, everything that becomes ready
        (a tile's cells live in shared memory: a step costs ~t_smem instead of ~1 us of L2 traffic);
        pushes that leave the tile are delivered at the end of the round (double-buffered in-boxes)

and reports rounds, tile visits, and the longest in-tile dependency depth per round, from which
the sweep time is estimated as  sum over rounds of (t_round + depth_max * t_smem).

    python tests/tools/proto_tile_sweep.py [n=1024] [tile sizes: 32 64 128]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from pydem_b200 import synth
from oracle.oracle import OracleDEMProcessor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tiles = [int(a) for a in sys.argv[2:]] or [32, 64, 128]
E = synth.conditioned_fractal_dem(n, 0)
dp = OracleDEMProcessor(E, dX=30.0, dY=30.0, fill_flats=False, drain_pits_path=False)
dp.calc_slopes_directions()
g, sec = dp._graph()
cptr, cidx, cdat, rptr, ridx = g.export()
N = n * n
indeg0 = (rptr[1:] - rptr[:-1]).astype(np.int64)
# reference depth: level-synchronous rounds of the plain sweep
dp.calc_uca()
print("DEM %dx%d conditioned: %d cells, %d edges, %d sources, %d dependency levels" % (n, n, N, cidx.size, int((indeg0 == 0).sum()), dp.stats["rounds"]))
cptr_l = cptr.tolist(); cidx_l = cidx.tolist()
T_SMEM, T_ROUND, T_GLOBAL = 0.1e-6, 5e-6, 1.0e-6      # s: in-tile step, per-round fixed cost (grid sync + tile load/store), today's step
for T in tiles:
    ntc = (n + T - 1) // T
    ii, jj = np.divmod(np.arange(N), n)
    tile_of = ((ii // T) * ntc + (jj // T)).tolist()
    indeg = indeg0.tolist()
    ready = {}
    for c in np.nonzero(indeg0 == 0)[0].tolist():
        ready.setdefault(tile_of[c], []).append(c)
    rounds = visits = drained = 0
    est = 0.0
    depth_hist = []
    t0 = time.time()
    while ready:
        rounds += 1
        inbox = {}
        dmax = 0
        for tl, cells in ready.items():
            visits += 1
            frontier = cells; depth = 0
            while frontier:                      # level-synchronous INSIDE the tile: depth = dependent steps
                depth += 1
                nxt = []
                for c in frontier:
                    drained += 1
                    for e in range(cptr_l[c], cptr_l[c + 1]):
                        r = cidx_l[e]
                        if tile_of[r] == tl:
                            indeg[r] -= 1
                            if indeg[r] == 0: nxt.append(r)
                        else:
                            inbox[r] = inbox.get(r, 0) + 1
                frontier = nxt
            dmax = max(dmax, depth)
        depth_hist.append(dmax)
        est += T_ROUND + dmax * T_SMEM
        ready = {}
        for r, k in inbox.items():
            indeg[r] -= k
            if indeg[r] == 0:
                ready.setdefault(tile_of[r], []).append(r)
    assert drained == N, (drained, N)
    print("tile %3d: %5d rounds, %7d tile visits (%.1f per tile), in-tile depth per round: mean %.1f max %d | estimated sweep %.2f ms "
          "(levels x %.1f us today: %.2f ms)  [%.0f s simulated]" % (T, rounds, visits, visits / (ntc * ntc), float(np.mean(depth_hist)), max(depth_hist),
          est * 1e3, T_GLOBAL * 1e6, dp.stats["rounds"] * T_GLOBAL * 1e3, time.time() - t0))
