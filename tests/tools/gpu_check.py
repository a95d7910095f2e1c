"""Exploratory parity run on the GPU box: every catalogue case, GPU vs oracle, one line each."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers
from oracle.oracle import OracleDEMProcessor
from pydem_b200 import DEMProcessor

only = sys.argv[1:]
fails = 0
for name, (E, kw) in helpers.cases().items():
    if only and name not in only:
        continue
    t0 = time.time()
    try:
        ref = helpers.run(lambda e, **k: OracleDEMProcessor(e, **k), E, kw)
        t1 = time.time()
        got = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), E, kw)
        t2 = time.time()
        r = helpers.compare(ref, got)
        try:
            helpers.assert_parity(r, name); ok = "OK  "
        except AssertionError:
            ok = "FAIL"; fails += 1
        keys = ["mag0_neq", "mag0_rel", "dir_abs", "flats0_neq", "flats_neq", "uca_rel", "uca_nanpat", "edge_todo_neq",
                "edge_done_neq", "mag_rel", "twi_abs"]
        print(ok, name, E.shape, "orc %.2fs gpu %.2fs" % (t1 - t0, t2 - t1), {k: r[k] for k in keys}, flush=True)
    except Exception as e:
        import traceback; traceback.print_exc()
        print("ERR ", name, repr(e), flush=True); fails += 1
print("failures:", fails)
