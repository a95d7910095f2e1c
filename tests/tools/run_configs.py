"""Test tooling (uses the oracle as the checker, hence under tests/).
BASELINE.json configs[0..2] on one B200 through the public operator (DEMProcessor): per-stage wall
times incl. host<->device copies, device-resident stage times, and a parity check of every config
against the oracle (full size where the oracle finishes in seconds, cropped windows above).
    python tests/tools/run_configs.py [--configs 1,2,3] [--size3 16384] [--out gpurun_out/configs.json]
Each config runs in its own subprocess under a timeout so that one slow stage cannot eat the box."""
import argparse, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def lakes_dem(n, seed):
    """config 3: fractal with one r=64 px lake (disc of constant elevation) per 1024^2 block."""
    from pydem_b200 import synth
    E = synth.fractal_dem(n, seed).copy()
    r = min(64, n // 8)
    yy, xx = np.mgrid[-r:r + 1, -r:r + 1]
    disc = yy * yy + xx * xx <= r * r
    for bi in range(0, n, 1024):
        for bj in range(0, n, 1024):
            ci, cj = bi + min(512, n // 2), bj + min(512, n // 2)
            w = E[ci - r:ci + r + 1, cj - r:cj + r + 1]
            w[disc] = w[disc].min()
    return E


def timed(f):
    import torch
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize()
    return r, (time.perf_counter() - t) * 1e3


def run_one(cfg, size3):
    import helpers
    from pydem_b200 import synth, DEMProcessor, _pinned
    from oracle.oracle import OracleDEMProcessor
    if cfg == 1:
        E = synth.cone_dem(256); kw = dict(dX=1.0, dY=1.0, fill_flats=True, drain_pits_path=True); name = "256x256 cone, default flags"
        win = None
    elif cfg == 2:
        E = synth.conditioned_fractal_dem(4096, 0); kw = dict(dX=30.0, dY=30.0, fill_flats=False, drain_pits_path=False)
        name = "4096x4096 conditioned fractal, fill_flats=False, drain_pits_path=False, drain_pits=True"
        win = (slice(1024, 1024 + 768), slice(2048, 2048 + 768))
    elif cfg == 3:
        # pit carving (drain_pits_path) visits the pits one after the other by definition (each sees the
        # previous carvings); the raw fractal has ~0.33 pits per 16 cells, so it is off here and
        # measured separately at 2048^2 (config 4 of this script)
        E = lakes_dem(size3, 1); kw = dict(dX=30.0, dY=30.0, fill_flats=True, drain_pits_path=False)
        name = "%dx%d fractal with lakes, fill_flats=True, drain_pits_path=False, drain_pits=True" % (size3, size3)
        c = min(512, size3 // 2)
        win = (slice(c - 320, c + 320), slice(c - 320, c + 320))     # around the first lake
    else:
        E = lakes_dem(2048, 1); kw = dict(dX=30.0, dY=30.0, fill_flats=True, drain_pits_path=True)
        name = "2048x2048 fractal with lakes, default flags (fill_flats=True, drain_pits_path=True, drain_pits=True)"
        win = (slice(512 - 192, 512 + 192), slice(512 - 192, 512 + 192))
    n_cells = E.size
    Eh = _pinned.pinned_copy(E)
    res = dict(config=cfg, workload=name, cells=int(n_cells))
    # warm-up (library load, pinned pools, tile allocation), then the measured pass
    for rep in range(1 if cfg >= 3 else 2):
        dp = DEMProcessor(elev=Eh, **kw)
        _, t_sd = timed(dp.calc_slopes_directions)
        _, t_uca = timed(dp.calc_uca)
        _, t_twi = timed(dp.calc_twi)
        st = dict(dp.uca_stats or {})
        out = dict(mag=dp.mag, direction=dp.direction, flats=dp.flats, uca=dp.uca, twi=dp.twi, elev=dp.elev)
        dp._free_tile()
    res.update(ms_calc_slopes_directions=t_sd, ms_calc_uca=t_uca, ms_calc_twi=t_twi,
               Mcells_s_slopes=n_cells / t_sd / 1e3, Mcells_s_uca=n_cells / t_uca / 1e3, Mcells_s_twi=n_cells / t_twi / 1e3,
               Mcells_s_total=n_cells / (t_sd + t_uca + t_twi) / 1e3,
               device={k: st.get(k) for k in ("ms_graph", "ms_sweep", "n_sources", "n_pits", "n_pit_edges", "n_undone")})
    res["cond_stats"] = {k: int(v) for k, v in (getattr(dp, "cond_stats", None) or {}).items()}
    # one chained call (what bench.py's e2e times)
    dp = DEMProcessor(elev=Eh, **kw)
    _, t_all = timed(dp.calc_twi)
    dp._free_tile()
    res.update(ms_calc_twi_chained=t_all, Mcells_s_chained=n_cells / t_all / 1e3)
    # parity against the oracle
    t0 = time.perf_counter()
    if win is None:
        ref = helpers.run(lambda e, **k: OracleDEMProcessor(e, **k), E, kw)
        got = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), E, kw)
        r = helpers.compare(ref, got); helpers.assert_parity(r, "config %d" % cfg)
        res["parity"] = dict(scope="full size", **{k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in r.items()})
    else:
        # the oracle on a window: interior cells whose whole upstream area lies inside the window agree
        # with the full-size run; stencil outputs agree everywhere but the window's rim
        Ew = np.ascontiguousarray(E[win])
        ref = helpers.run(lambda e, **k: OracleDEMProcessor(e, **k), Ew, kw)
        gw = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), Ew, kw)
        r = helpers.compare(ref, gw); helpers.assert_parity(r, "config %d window" % cfg)
        res["parity"] = dict(scope="window %s of the same DEM, GPU vs oracle on identical input" % str([(s.start, s.stop) for s in win]),
                             **{k: (float(v) if isinstance(v, (float, np.floating)) else v) for k, v in r.items()})
        if cfg == 2:
            inner = (slice(win[0].start + 2, win[0].stop - 2), slice(win[1].start + 2, win[1].stop - 2))
            m_full = out["mag"][inner]; m_win = gw["mag"][2:-2, 2:-2]
            res["parity"]["full_vs_window_mag_equal"] = bool(np.array_equal(m_full, m_win, equal_nan=True))
    res["parity_s"] = time.perf_counter() - t0
    # size-independent properties at full size
    u = out["uca"]; fl = out["flats"]
    area = (kw["dX"] * kw["dY"])
    res["properties"] = dict(uca_min_is_cell_area=bool(np.nanmin(u) == area), uca_nan_only_on_flats=bool(np.array_equal(np.isnan(u), fl)),
                             undone=int(st.get("n_undone", 0) or 0),
                             mass_out_over_cells=float(np.nansum(u[0]) + np.nansum(u[-1]) + np.nansum(u[1:-1, 0]) + np.nansum(u[1:-1, -1])) / area / n_cells)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3")
    ap.add_argument("--size3", type=int, default=16384)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    ap.add_argument("--child", type=int, default=0)
    ap.add_argument("--timeout", type=int, default=420)
    a = ap.parse_args()
    if a.child:
        print("RESULT " + json.dumps(run_one(a.child, a.size3)), flush=True)
        sys.exit(0)
    results = []
    for c in [int(x) for x in a.configs.split(",")]:
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(c), "--size3", str(a.size3)],
                               capture_output=True, text=True, timeout=a.timeout)
            lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if lines:
                results.append(json.loads(lines[-1][7:]))
            else:
                results.append(dict(config=c, error=(p.stderr or p.stdout)[-1500:]))
        except subprocess.TimeoutExpired:
            results.append(dict(config=c, error="timeout after %d s" % a.timeout))
        print(json.dumps(results[-1]), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(results, open(a.out, "w"), indent=1)
