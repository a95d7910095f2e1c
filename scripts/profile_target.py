"""Minimal device-resident pass for ncu: python scripts/profile_target.py N [reps] [drain_pits] [raw|cond]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pydem_b200 import synth, tile as T
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dp = int(sys.argv[3]) if len(sys.argv) > 3 else 0
kind = sys.argv[4] if len(sys.argv) > 4 else "raw"
E = synth.conditioned_fractal_dem(n, 0, wrap_rows=True) if kind == "cond" else synth.fractal_dem(n, 0)
dt = T.DeviceTile(n, n)
dt.set_spacing(30.0, 30.0)
dt.upload(T.F_ELEV, E)
for r in range(reps):
    dt.slopes_directions()
    st = dt.uca(drain_pits=dp)
    dt.twi()
dt.sync()
print(st)
