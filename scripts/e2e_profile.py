import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pydem_b200 import synth, DEMProcessor, _pinned, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
_lib.init(0)
E = _pinned.pinned_copy(synth.fractal_dem(n, 0))
kw = dict(dX=30.0, dY=30.0, fill_flats=False, drain_pits_path=False, drain_pits=False)
def step():
    dp = DEMProcessor(elev=E, **kw); t = dp.calc_twi(); x = float(t[n // 2, n // 2]); dp._free_tile(); return x
for _ in range(3): step()
t0 = time.perf_counter()
for _ in range(5): step()
print("e2e ms/step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
