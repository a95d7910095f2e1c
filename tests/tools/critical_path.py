"""Test tooling (uses the oracle, hence under tests/): the dependency structure of the benchmark's drainage graph --
the analysis behind the chain bursts of the sweep.  Builds the graph of the conditioned
4096 x 4096 benchmark DEM with the oracle, computes the dependency level of every cell (Kahn), the cells per level,
and walks the longest flow path back from the deepest cell: out-degree / in-degree of its cells, weights of its edges,
how many donors of a path cell sit exactly one level above it.
    python tests/tools/critical_path.py [n]
Observed for n = 4096: 3895 levels; 100 k of 16.8 M cells above level 500, 20 k above level 1000 (median 5 per level);
3896 path cells, 3674 with one receiver, 94 % of the path's edges with weight exactly 1; 1931 of the first 1999 path
cells have exactly one donor on the level above."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from pydem_b200 import synth
from oracle import oracle as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
E = synth.conditioned_fractal_dem(n, 0, wrap_rows=True)
dp = orc.OracleDEMProcessor(E, dX=30.0, dY=30.0, fill_flats=False, drain_pits_path=False, drain_pits=True)
dp.calc_slopes_directions()
g, sec = dp._graph()
cptr, cidx, cdat, rptr, ridx = g.export()   # CSC: column i -> receivers cidx[cptr[i]:cptr[i+1]] ; CSR rows: donors
N = n * n
outdeg = np.diff(cptr); indeg = np.diff(rptr)
print("cells", N, "edges", cidx.size, "outdeg hist", np.bincount(outdeg)[:6], "indeg hist", np.bincount(indeg)[:10])
# levels by Kahn
t = time.time()
level = np.zeros(N, np.int32)
deg = indeg.copy()
front = np.nonzero(deg == 0)[0]
lv = 0
counts = []
parent = np.full(N, -1, np.int64)
while front.size:
    counts.append(front.size)
    # receivers of front
    starts = cptr[front]; lens = cptr[front + 1] - starts
    tot = int(lens.sum())
    if tot == 0: break
    idx = np.repeat(starts - np.concatenate(([0], np.cumsum(lens)[:-1])), lens) + np.arange(tot)
    rec = cidx[idx]
    src = np.repeat(front, lens)
    np.subtract.at(deg, rec, 1)
    lv += 1
    cand = np.unique(rec)
    nf = cand[deg[cand] == 0]
    level[nf] = lv
    # remember one donor of max level (= lv-1 donors are in front): pick any src that is in front
    parent[rec] = src     # last writer among current front: it has level lv-1, valid for cells that become ready now
    front = nf
print("levels", lv, "time %.1f" % (time.time() - t))
counts = np.array(counts)
print("cells per level: first 10", counts[:10], "median over levels>1000:", np.median(counts[1000:]), "sum over levels>500", counts[500:].sum(), ">1000", counts[1000:].sum(), ">2000", counts[2000:].sum())
# critical path: from deepest cell back via parent
c = int(np.argmax(level))
path = [c]
while parent[path[-1]] >= 0 and level[path[-1]] > 0:
    path.append(int(parent[path[-1]]))
path = np.array(path[::-1])
print("critical path cells", path.size)
od = outdeg[path]; idg = indeg[path]
print("path outdeg hist", np.bincount(od)[:5], "path indeg hist", np.bincount(idg)[:10])
# weights of the path edge
w = []
for a, b in zip(path[:-1], path[1:]):
    s, e = cptr[a], cptr[a + 1]
    k = np.nonzero(cidx[s:e] == b)[0]
    w.append(cdat[s + k[0]] if k.size else np.nan)
w = np.array(w)
print("path edge weight quantiles", np.nanquantile(w, [0.05, 0.25, 0.5, 0.75, 0.95]), "frac w>0.99", np.mean(w > 0.99), "frac w==1", np.mean(w == 1.0))
# how many donors of a path cell have level == level-1 (i.e. arrive at the same time: the 'ladder')
lad = []
for cnode in path[1:2000:1]:
    d = ridx[rptr[cnode]:rptr[cnode + 1]]
    lad.append(int((level[d] == level[cnode] - 1).sum()))
print("donors at level-1 per path cell hist", np.bincount(lad))
