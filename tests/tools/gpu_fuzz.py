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
print("cases", n, "mismatches", bad)
sys.exit(1 if bad else 0)
