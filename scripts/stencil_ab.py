"""A/B of library builds on the stencil stage: python scripts/stencil_ab.py [n] lib1.so lib2.so ...
(one subprocess per library, PYDEM_B200_LIB).  Prints best / median ms of slopes_directions and checks
mag / direction / flats bit for bit against the first library."""
import sys, os, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    import torch
    from pydem_b200 import tile as T
    n = int(sys.argv[2]); tag = sys.argv[3]
    E = np.load("/tmp/sab_%d.npy" % n)
    dt = T.DeviceTile(n, n, stream=torch.cuda.current_stream().cuda_stream)
    dt.set_spacing(30.0, 30.0); dt.upload(T.F_ELEV, E)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ms = []
    for rep in range(25):
        torch.cuda.synchronize(); ev[0].record(); dt.slopes_directions(); ev[1].record(); torch.cuda.synchronize()
        if rep >= 3: ms.append(ev[0].elapsed_time(ev[1]))
    out = {f: dt.download(getattr(T, "F_" + f)) for f in ("MAG", "DIR", "FLATS")}
    np.savez("/tmp/sab_out_%s.npz" % tag, **out)
    print(json.dumps(dict(lib=os.environ.get("PYDEM_B200_LIB", "default"), stencil=os.environ.get("PYDEM_B200_STENCIL", "tiled"), n=n, ms_best=round(min(ms), 4), ms_median=round(float(np.median(ms)), 4))), flush=True)
    sys.exit(0)
from pydem_b200 import synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
libs = sys.argv[2:] or ["default"]
np.save("/tmp/sab_%d.npy" % n, synth.conditioned_fractal_dem(n, 0))
for rnd in range(2):
    for k, lib in enumerate(libs):
        env = dict(os.environ)
        if lib.startswith("stencil="): env["PYDEM_B200_STENCIL"] = lib.split("=")[1]      # stencil=v1 | parity | tiled
        elif lib != "default": env["PYDEM_B200_LIB"] = os.path.abspath(lib)
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(n), str(k)], env=env, check=False)
ref = np.load("/tmp/sab_out_0.npz")
for k in range(1, len(libs)):
    o = np.load("/tmp/sab_out_%d.npz" % k)
    print(libs[k], "bit-identical to", libs[0], all(np.array_equal(ref[f], o[f], equal_nan=True) for f in ("MAG", "DIR", "FLATS")))
