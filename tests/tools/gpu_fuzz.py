"""Randomised GPU-vs-oracle parity (test tooling; run on a GPU box):
    python tests/tools/gpu_fuzz.py [n_cases=100] [first_seed=0]
The cases are tests/helpers.fuzz_case(seed) -- the ones tests/test_oracle_fuzz.py pins against the
live reference on the CPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
from oracle.oracle import OracleDEMProcessor
from pydem_b200 import DEMProcessor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
bad = 0
for seed in range(s0, s0 + n):
    E, kw, kind = helpers.fuzz_case(seed)
    ref = helpers.run_all(lambda e, **k: OracleDEMProcessor(e, **k), E, kw)
    got = helpers.run_all(lambda e, **k: DEMProcessor(elev=e, **k), E, kw)
    r = helpers.compare(ref, got)
    ok = np.array_equal(ref["elev"], got["elev"], equal_nan=True) and all(r[k + "_neq"] == 0 for k in ("flats0", "flats", "edge_todo", "edge_done")) and \
        all(r[k + "_nanpat"] == 0 for k in ("mag0", "mag", "dir", "uca", "twi")) and r["mag_rel"] <= 1e-12 and r["dir_abs"] <= 1e-12 and \
        r["uca_rel"] <= 1e-9 and r["twi_abs"] <= 1e-8
    if not ok:
        bad += 1
        print("MISMATCH seed", seed, "kind", kind, E.shape, r)
    # update mode with random neighbour strips
    kwu = dict(kw, fill_flats=False, drain_pits_path=False)
    u1, t1, d1 = helpers.update_call(lambda e, **k: OracleDEMProcessor(e, **k), E, kwu, seed)
    u2, t2, d2 = helpers.update_call(lambda e, **k: DEMProcessor(elev=e, **k), E, kwu, seed)
    with np.errstate(invalid="ignore"):
        rel = float(np.nanmax(np.abs(u1 - u2) / np.abs(u1))) if np.isfinite(u1).any() else 0.0
    if not (np.array_equal(np.isnan(u1), np.isnan(u2)) and rel <= 1e-9 and np.array_equal(t1, t2) and np.array_equal(d1, d2)):
        bad += 1
        print("UPDATE MISMATCH seed", seed, "kind", kind, E.shape, "rel", rel, int((t1 != t2).sum()), int((d1 != d2).sum()))
print("cases", n, "mismatches", bad)
sys.exit(1 if bad else 0)
