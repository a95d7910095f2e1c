import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from pydem_b200 import synth, sharded, DEMProcessor
n = int(sys.argv[1]); world = int(sys.argv[2])
B = synth.conditioned_fractal_dem(n, 0, wrap_rows=True)
E = np.vstack([B] * world)
t = time.time()
out = sharded.run_local(E, world, dX=30.0, dY=30.0)
print("sharded done %.2fs" % (time.time() - t), {k: out["stats"][0][k] for k in ("sweep_rounds", "label_rounds", "n_drained")}, flush=True)
dp = DEMProcessor(elev=E, dX=30.0, dY=30.0, fill_flats=False, drain_pits_path=False, drain_pits=False)
dp.calc_twi()
print("uca equal:", np.allclose(out["uca"], dp.uca, rtol=1e-9, equal_nan=True), "mag eq", np.array_equal(out["mag"], dp.mag), flush=True)
