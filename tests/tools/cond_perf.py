"""Time the elevation-conditioning stages on the device (config-3 style input: quantised fractal
with lakes) and the oracle restatement on a crop, and check them against each other on the crop.
    python tests/tools/cond_perf.py [n] [crop]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from pydem_b200 import synth, tile as T
from oracle import conditioning as oc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
crop = int(sys.argv[2]) if len(sys.argv) > 2 else 384


def lakes_dem(n, seed, quant=None):
    E = synth.fractal_dem(n, seed).copy()
    if quant:
        E = np.round(E / quant) * quant + 1
    yy, xx = np.mgrid[:n, :n]
    for bi in range(0, n, 1024):
        for bj in range(0, n, 1024):
            ci, cj, r = bi + min(512, n // 2), bj + min(512, n // 2), min(64, n // 8)
            m = (yy - ci) ** 2 + (xx - cj) ** 2 <= r * r
            E[m] = E[m].min()
    return E


def run(E, label):
    R = E.shape[0]
    t = T.DeviceTile(*E.shape); t.set_spacing(30.0, 30.0)
    out = {}
    for stage in ("fill_flats", "pit_drain_paths"):
        times = []
        for rep in range(2):
            t.upload(T.F_ELEV, E if stage == "fill_flats" else out["fill_flats"]); t.sync()
            t0 = time.perf_counter(); st = t.condition(stage); t.sync(); times.append(time.perf_counter() - t0)
        out[stage] = t.download(T.F_ELEV)
        print("%s %s %dx%d: %.1f ms  (%.1f Mcells/s)  %s" % (label, stage, R, E.shape[1], min(times) * 1e3, E.size / min(times) / 1e6,
                                                           {k: v for k, v in st.items() if v}), flush=True)
    t.close()
    return out


for label, E in (("lakes", lakes_dem(n, 1)), ("lakes+quantised", lakes_dem(n, 1, quant=0.5))):
    run(E, label)
    sub = np.ascontiguousarray(E[n // 2 - crop // 2:n // 2 + crop // 2, n // 2 - crop // 2:n // 2 + crop // 2])
    got = run(sub, label + " crop")
    t0 = time.perf_counter(); f = oc.fill_flats(sub); t1 = time.perf_counter()
    p = oc.pit_drain_paths(f, np.full(crop - 1, 30.0), np.full(crop - 1, 30.0))[0]; t2 = time.perf_counter()
    print("%s crop oracle (1 core): fill_flats %.0f ms, pit_drain_paths %.0f ms; device == oracle: %s %s" % (
        label, (t1 - t0) * 1e3, (t2 - t1) * 1e3, np.array_equal(got["fill_flats"], f), np.array_equal(got["pit_drain_paths"], p)), flush=True)
