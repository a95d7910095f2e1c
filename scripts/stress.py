"""Stress the work-list engine: many repeated sweeps (single tile and sharded LocalGroup), every
run must drain every cell and reproduce the first run's uca to fp64 re-association.
    python scripts/stress.py [n] [reps] [kinds: cond,raw] [sharded 0/1]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pydem_b200 import synth, sharded, tile as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
kinds = sys.argv[3].split(",") if len(sys.argv) > 3 else ["cond", "raw"]
do_sharded = int(sys.argv[4]) if len(sys.argv) > 4 else 1
for kind in kinds:
    E = synth.conditioned_fractal_dem(n, 0, wrap_rows=True) if kind == "cond" else synth.fractal_dem(n, 0)
    dt = T.DeviceTile(n, n); dt.set_spacing(30.0, 30.0); dt.upload(T.F_ELEV, E); dt.slopes_directions()
    ref = None; t0 = time.time()
    for r in range(reps):
        dt.slopes_directions()       # calc_uca drains pits in place (mag/flats), so every run starts from the stencil
        try:
            st = dt.uca(drain_pits=1 if kind == "cond" else 0)
        except Exception as e:
            print(kind, "rep", r, "FAILED:", e, flush=True)
            raise
        assert st["n_drained"] == n * n and st["n_undone"] == 0, st
        if r % 25 == 0:
            u = dt.download(T.F_UCA)
            if ref is None: ref = u
            assert np.allclose(u, ref, rtol=1e-10, equal_nan=True)
    print(kind, "single tile", reps, "sweeps ok, %.1f ms each" % ((time.time() - t0) / reps * 1e3), flush=True)
    dt.close()
    if not do_sharded:
        continue
    world = 3
    S = np.vstack([E] * world) if kind == "cond" else synth.fractal_dem(0, 1, shape=(n * world, n))
    first = None; t0 = time.time()
    for r in range(max(reps // 10, 5)):
        out = sharded.run_local(S, world, dX=30.0, dY=30.0)
        assert sum(s["n_drained"] for s in out["stats"]) == S.size
        if first is None: first = out["uca"]
        assert np.allclose(out["uca"], first, rtol=1e-10, equal_nan=True)
    print(kind, "sharded x3", max(reps // 10, 5), "runs ok, rounds", out["stats"][0]["sweep_rounds"], "%.0f ms each" % ((time.time() - t0) / max(reps // 10, 5) * 1e3), flush=True)
print("stress ok")
