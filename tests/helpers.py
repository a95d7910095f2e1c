"""Shared helpers of the parity tests: case catalogue, oracle/GPU runners, comparison."""
import io
import contextlib
import warnings

import numpy as np

from pydem_b200 import synth

# tolerances of the parity bar (north_star: "stated float tolerance, integers bit-exact")
DIR_ATOL = 1e-12     # radians; CUDA atan2 vs host libm/NumPy differ by <= 2 ulp
MAG_RTOL = 1e-13     # sqrt/div are IEEE-exact on both sides; only pit mean-slopes re-associate
UCA_RTOL = 1e-9      # fp64 atomics re-associate the reference's (level, index) summation order
TWI_ATOL = 1e-9

HOT = dict(fill_flats=False, drain_pits_path=False)


def cases():
    """name -> (elev, kwargs).  Sizes the oracle finishes in seconds."""
    out = {}
    card = np.repeat(np.arange(1.0, 6.0)[:, None], 5, 1)
    diag = np.add.outer(np.arange(5.0), np.arange(5.0)) + 1.0
    for nm, E in (("card", card), ("diag", diag)):
        out[nm] = (E, {})
        out[nm + "_rev"] = (E[::-1].copy(), {})
        out[nm + "_T"] = (E.T.copy(), {})
        out[nm + "_Trev"] = (E[::-1, ::-1].T.copy(), {})
    out["cone32"] = (synth.cone_dem(32), {})
    out["cone256"] = (synth.cone_dem(256), {})
    out["cone256_nopits"] = (synth.cone_dem(256), dict(drain_pits=False))
    out["frac128"] = (synth.fractal_dem(128, 1), {})
    out["frac128_nopits"] = (synth.fractal_dem(128, 2), dict(drain_pits=False))
    out["frac256_dx30"] = (synth.fractal_dem(256, 3), dict(dX=30.0, dY=30.0))
    R = 200
    out["frac_rect_vardx"] = (synth.fractal_dem(0, 4, shape=(R, 150)),
                              dict(dX=np.linspace(20, 30, R - 1), dY=np.full(R - 1, 27.3),
                                   dX2=np.linspace(20, 30, R), dY2=np.full(R, 27.3)))
    E = synth.fractal_dem(128, 5); E[40:44, 50:53] = np.nan; E[80, 90] = np.nan; E[0, 5] = np.nan
    out["nan_holes"] = (E, {})
    out["quantized"] = (np.round(synth.fractal_dem(128, 6) / 20) * 20, {})
    E = synth.fractal_dem(160, 7)
    yy, xx = np.mgrid[0:160, 0:160]
    for (cy, cx, r) in ((40, 40, 9), (100, 120, 14), (120, 30, 6)):
        m = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        E[m] = E[m].min()
    out["lakes"] = (E, {})
    out["lakes_minborder"] = (E, dict(drain_pits_min_border=True))
    out["frac96_maxdist4"] = (synth.fractal_dem(96, 8), dict(drain_pits_max_dist=4, drain_pits_max_iter=20))
    out["frac96_xy"] = (synth.fractal_dem(96, 9), dict(dX=30.0, dY=30.0, drain_pits_max_dist_XY=70.0))
    out["odd_cols"] = (synth.fractal_dem(0, 10, shape=(67, 131)), dict(dX=10.0, dY=12.5))
    out["tiny3"] = (np.array([[3.0, 2.0, 3.0], [2.0, 1.0, 2.0], [3.0, 2.5, 3.0]]), {})
    # circular_ref_maxcount only acts through the loop condition of dem_processing.py:951-952 (the drainage graph
    # is acyclic): <= 1 switches the accumulation sweep off, 2 allows exactly the one call that is ever needed
    out["frac96_maxcount1"] = (synth.fractal_dem(96, 13), dict(circular_ref_maxcount=1))
    out["frac96_maxcount2"] = (synth.fractal_dem(96, 13), dict(circular_ref_maxcount=2, drain_pits=False))
    out["frac512_limits"] = (synth.fractal_dem(512, 11), dict(dX=30.0, dY=30.0, apply_uca_limit_edges=True,
                                                            apply_twi_limits=True, apply_twi_limits_on_uca=True))
    return out


def golden_cases():
    """Small catalogue whose reference outputs are committed under tests/golden/."""
    out = {k: v for k, v in cases().items() if k.startswith(("card", "diag")) or k in ("cone32", "tiny3")}
    out["frac48"] = (synth.fractal_dem(48, 31), {})
    out["frac48_nopits"] = (synth.fractal_dem(48, 32), dict(drain_pits=False))
    R = 40
    out["rect_vardx"] = (synth.fractal_dem(0, 33, shape=(R, 30)),
                         dict(dX=np.linspace(20, 30, R - 1), dY=np.full(R - 1, 27.3),
                              dX2=np.linspace(20, 30, R), dY2=np.full(R, 27.3)))
    E = synth.fractal_dem(48, 34); E[20:23, 30:32] = np.nan; E[40, 10] = np.nan; E[0, 5] = np.nan
    out["nan48"] = (E, {})
    out["quant48"] = (np.round(synth.fractal_dem(48, 35) / 20) * 20, {})
    E = synth.fractal_dem(64, 36)
    yy, xx = np.mgrid[0:64, 0:64]
    for (cy, cx, r) in ((20, 20, 6), (45, 40, 8)):
        m = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        E[m] = E[m].min()
    out["lakes64"] = (E, {})
    out["lakes64_minborder"] = (E, dict(drain_pits_min_border=True))
    out["maxdist40"] = (synth.fractal_dem(40, 37), dict(drain_pits_max_dist=4, drain_pits_max_iter=20))
    out["xy40"] = (synth.fractal_dem(40, 38), dict(dX=30.0, dY=30.0, drain_pits_max_dist_XY=70.0))
    out["odd37x51"] = (synth.fractal_dem(0, 39, shape=(37, 51)), dict(dX=10.0, dY=12.5))
    out["limits48"] = (synth.fractal_dem(48, 40), dict(dX=30.0, dY=30.0, apply_uca_limit_edges=True,
                                                       apply_twi_limits=True, apply_twi_limits_on_uca=True))
    return out


def run(make, E, kw):
    """Run the three stages on a processor factory; returns a dict of outputs."""
    k = dict(HOT); k.update(kw)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        dp = make(E, **k)
        mag, direction = dp.calc_slopes_directions()
        out = dict(mag0=np.array(mag), dir=np.array(direction), flats0=np.array(dp.flats))
        out["uca"] = np.array(dp.calc_uca())
        out["edge_todo"] = np.array(dp.edge_todo); out["edge_done"] = np.array(dp.edge_done)
        out["mag"] = np.array(dp.mag); out["flats"] = np.array(dp.flats)
        out["twi"] = np.array(dp.calc_twi()); out["twi10"] = np.array(dp.twi)
        out["twi_min_area"] = dp.twi_min_area
    return out


def _nanmax(x):
    x = x[np.isfinite(x)]
    return float(x.max()) if x.size else 0.0


def compare(ref, got):
    """dict of error measures between two output dicts (ref = oracle / reference)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        r = {}
        for k in ("mag0", "mag", "dir", "uca", "twi", "twi10"):
            a, b = np.asarray(ref[k], float), np.asarray(got[k], float)
            r[k + "_nanpat"] = int((np.isnan(a) != np.isnan(b)).sum())
            r[k + "_abs"] = _nanmax(np.abs(a - b))
            r[k + "_rel"] = _nanmax(np.abs(a - b) / np.abs(a))
            r[k + "_neq"] = int(((a != b) & ~(np.isnan(a) & np.isnan(b))).sum())
        for k in ("flats0", "flats", "edge_todo", "edge_done"):
            r[k + "_neq"] = int((np.asarray(ref[k], bool) != np.asarray(got[k], bool)).sum())
        r["min_area_eq"] = bool(ref["twi_min_area"] == got["twi_min_area"])
    return r


def assert_parity(r, name=""):
    msg = "%s: %s" % (name, r)
    for k in ("flats0", "flats", "edge_todo", "edge_done"):
        assert r[k + "_neq"] == 0, msg                       # integer/bool outputs bit-exact
    for k in ("mag0", "mag", "dir", "uca", "twi", "twi10"):
        assert r[k + "_nanpat"] == 0, msg
    assert r["mag0_rel"] <= MAG_RTOL and r["mag_rel"] <= 1e-12, msg
    assert r["dir_abs"] <= DIR_ATOL, msg
    assert r["uca_rel"] <= UCA_RTOL, msg
    assert r["twi_abs"] <= TWI_ATOL, msg
    assert r["min_area_eq"], msg


def conditioning_cases():
    """name -> (elev, dX, dY): inputs of the elevation-conditioning parity tests (fill_flats,
    pit artifacts, pit drain paths).  Small enough for the reference/oracle Python loops."""
    rng = np.random.default_rng(5)
    out = {}
    out["fractal96"] = synth.fractal_dem(96, 3)
    out["quant96"] = np.round(synth.fractal_dem(96, 3) / 8.0) + 1        # coarse quantisation: flats, artifacts, tied pits
    out["quant_fine80"] = np.round(synth.fractal_dem(80, 4) / 2.0) + 1
    g = synth.fractal_dem(128, 7).copy()
    yy, xx = np.mgrid[:128, :128]
    for (ci, cj, r) in ((40, 40, 12), (90, 70, 9), (20, 100, 6), (0, 64, 10), (127, 20, 8)):   # two lakes cross the array edge
        m = (yy - ci) ** 2 + (xx - cj) ** 2 <= r * r
        g[m] = g[m].min()
    out["lakes128"] = g
    out["noise_int"] = np.round(rng.random((40, 56)) * 6) + 1
    p = synth.fractal_dem(64, 9).copy()
    p[20:30, 20:30] = p.max() + 5; p[40:44, 10:30] = p.max() + 7               # plateaus: the "peak" branch
    out["peaks64"] = p
    out["rect50x90"] = np.round(synth.fractal_dem(0, 11, shape=(50, 90)) / 5.0) + 2
    s = synth.fractal_dem(72, 12).copy(); s[s < np.percentile(s, 20)] = 0.0    # sea (elev == 0) is masked out
    out["sea72"] = s
    a = np.round(synth.fractal_dem(64, 13) / 3.0) + 5                           # planted quantisation artifacts
    for (ci, cj, h, w) in ((10, 10, 1, 1), (20, 30, 2, 3), (40, 15, 5, 6), (45, 45, 6, 6), (30, 50, 1, 4), (55, 8, 3, 3)):
        v = a[ci, cj]
        a[ci - 1:ci + h + 1, cj - 1:cj + w + 1] = v + 1
        a[ci:ci + h, cj:cj + w] = v
    out["artifacts64"] = a
    res = {}
    for k, E in out.items():
        R = E.shape[0]
        res[k] = (E, 30.0 + np.linspace(0, 3, R - 1), np.full(R - 1, 28.0))
    return res


def pm_cases():
    """name -> (DEM, nx_grid, ny_grid, overlap, dem_proc_kwargs) for the tile-orchestrator tests:
    the reference's own multi-file tilings of the 32x32 cone (test_end_to_end.py:86-149) and a
    rough fractal."""
    x, y = np.mgrid[-1:1:32j, -1:1:32j]
    cone = 1 - np.sqrt(x ** 2 + y ** 2) / np.sqrt(2)              # utils_test_pydem.py:98-103
    frac = synth.fractal_dem(64, 3)
    c = {}
    for (nx, ny, ov) in ((3, 3, 2), (5, 4, 2), (5, 4, 3), (3, 3, 1), (3, 4, 1)):
        c["cone_%dx%d_%doverlap" % (nx, ny, ov)] = (cone, nx, ny, ov, {})
    c["fractal_3x3_2overlap"] = (frac, 3, 3, 2, {})
    c["fractal_2x3_1overlap"] = (frac, 2, 3, 1, {})
    # beyond the reference's own tilings: no overlap at all, single rows / columns of tiles, wide overlap, flags
    c["cone_3x3_0overlap"] = (cone, 3, 3, 0, {})
    c["fractal_2x2_0overlap"] = (frac, 2, 2, 0, {})
    c["fractal_4x1_2overlap"] = (frac, 4, 1, 2, {})
    c["fractal_1x3_4overlap"] = (frac, 1, 3, 4, {})
    c["fractal_3x3_2overlap_nopits"] = (frac, 3, 3, 2, dict(drain_pits=False))
    c["fractal_2x2_2overlap_limits"] = (frac, 2, 2, 2, dict(apply_uca_limit_edges=True, apply_twi_limits=True, apply_twi_limits_on_uca=True))
    return c


# one more orchestrator case with the spacing the reference derives from the rasters instead of its
# tests' dX = dY = 1: a 64x64 fractal spanning 1 degree of latitude and 1.5 of longitude
PM_SPACING_CASE = "fractal_3x3_2overlap_spacing"
PM_SPACING_GEO = dict(lat=(46.0, 45.0), lon=(-73.0, -71.5))


def pm_spacing(shape, boxes):
    """per-tile spacing of PM_SPACING_CASE, with the arithmetic of the reference's raster writer
    (utils.mk_geotiff_obj :196-199 on the tile's own corner coordinates) and reader (utils.py:132-137):
    dX = pixel width, dY = |pixel height| of each tile's raster"""
    la = np.linspace(PM_SPACING_GEO["lat"][0], PM_SPACING_GEO["lat"][1], shape[0])
    lo = np.linspace(PM_SPACING_GEO["lon"][0], PM_SPACING_GEO["lon"][1], shape[1])
    out = []
    for (te, be, le, re) in boxes:
        ph = abs(la[te] - la[be - 1]) / ((be - te) - 1.0)
        pw = abs(lo[le] - lo[re - 1]) / ((re - le) - 1.0)
        out.append(dict(dX=pw, dY=ph, dX2=pw, dY2=ph))
    return out


def pm_compare(pm, G, name):
    """Our ProcessManager's per-tile arrays against the reference's side-by-side store (golden
    dict G).  Returns the worst deviation per field (bool fields: number of differing cells)."""
    worst = {}
    gs = G[name + "_grid_slice"]
    for i, t in enumerate(pm.tiles):
        sl = (slice(gs[i][0], gs[i][1]), slice(gs[i][2], gs[i][3]))
        for key in ("elev", "aspect", "slope", "uca", "uca_edges", "edge_todo", "edge_done", "twi"):
            a = G["%s_%s" % (name, key)][sl]; b = getattr(t, key)
            if a.dtype == bool:
                d = float((a != b).sum())
            else:
                with np.errstate(invalid="ignore"):
                    d = np.where(np.isnan(a) & np.isnan(b), 0.0, np.where(np.isnan(a) != np.isnan(b), np.inf, np.abs(a - b)))
                    if key in ("uca", "uca_edges"):
                        d = d / np.maximum(1.0, np.abs(np.nan_to_num(a)))
                d = float(d.max())
            worst[key] = max(worst.get(key, 0.0), d)
    return worst


def fuzz_case(seed):
    """A small random DEM (plain / quantised / lakes / NaN holes / tilted plane + noise / integer
    elevations) with random spacing and flags.  Returns (elev, kwargs, kind)."""
    rng = np.random.default_rng(seed)
    R = int(rng.integers(12, 56)); C = int(rng.integers(12, 56))
    E = synth.fractal_dem(0, seed + 1000, shape=(R, C))
    kind = rng.integers(0, 6)
    if kind == 1:
        E = np.round(E / rng.choice([5, 20, 50])) * rng.choice([5, 20, 50]) + 1
    elif kind == 2:
        yy, xx = np.mgrid[:R, :C]
        for _ in range(int(rng.integers(1, 4))):
            cy, cx, r = rng.integers(0, R), rng.integers(0, C), rng.integers(2, 8)
            m = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
            E[m] = E[m].min()
    elif kind == 3:
        for _ in range(int(rng.integers(1, 4))):
            i, j = rng.integers(0, R - 2), rng.integers(0, C - 2)
            E[i:i + rng.integers(1, 3), j:j + rng.integers(1, 3)] = np.nan
    elif kind == 4:
        E = E * 0.01 + np.add.outer(np.arange(R) * rng.uniform(-2, 2), np.arange(C) * rng.uniform(-2, 2)) + 500
    elif kind == 5:
        E = np.round(E)  # integer elevations: many ties
    kw = {}
    if rng.random() < 0.5:
        kw.update(dX=float(rng.uniform(5, 40)), dY=float(rng.uniform(5, 40)))
    if rng.random() < 0.3:
        kw.update(dX=np.linspace(20, 30, R - 1) * rng.uniform(0.5, 2), dY=np.full(R - 1, float(rng.uniform(10, 40))))
        kw.update(dX2=np.linspace(20, 30, R), dY2=np.full(R, 27.0))
    kw["drain_pits"] = bool(rng.random() < 0.7)
    if rng.random() < 0.3: kw["drain_pits_min_border"] = True
    if rng.random() < 0.3: kw.update(drain_pits_max_dist=int(rng.integers(2, 10)), drain_pits_max_iter=int(rng.integers(5, 40)))
    if rng.random() < 0.3: kw.update(apply_uca_limit_edges=True, apply_twi_limits=True, apply_twi_limits_on_uca=True)
    cond = rng.random() < 0.4 and kind != 3
    kw["fill_flats"] = bool(cond); kw["drain_pits_path"] = bool(cond and rng.random() < 0.7)
    return E, kw, kind


def run_all(make, E, kw):
    """Conditioning (per flags) + the three stages on a processor factory, every output incl. elev."""
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        dp = make(E.copy(), **kw)
        mag, direction = dp.calc_slopes_directions()
        out = dict(elev=np.array(dp.elev), mag0=np.array(mag), dir=np.array(direction), flats0=np.array(dp.flats))
        out["uca"] = np.array(dp.calc_uca()); out["edge_todo"] = np.array(dp.edge_todo); out["edge_done"] = np.array(dp.edge_done)
        out["mag"] = np.array(dp.mag); out["flats"] = np.array(dp.flats)
        out["twi"] = np.array(dp.calc_twi()); out["twi10"] = np.array(dp.twi); out["twi_min_area"] = dp.twi_min_area
    return out


def update_call(make, E, kw, seed):
    """full calc_uca, then calc_uca(uca_init, edge_init_data) with RANDOM neighbour strips (values,
    done flags, a random subset of the tile's own todo edges): dem_processing.py:719-744, 778-862"""
    R, C = E.shape
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        dp = make(E.copy(), **kw)
        dp.calc_slopes_directions(); dp.calc_uca()
        uca0 = np.array(dp.uca); todo0 = np.array(dp.edge_todo)
        r2 = np.random.default_rng(88_000 + seed)
        data = {"left": r2.uniform(1, 500, R), "right": r2.uniform(1, 500, R), "top": r2.uniform(1, 500, C), "bottom": r2.uniform(1, 500, C)}
        done = {k: r2.random(v.size) < 0.6 for k, v in data.items()}
        todo = {"left": todo0[:, 0] & (r2.random(R) < 0.9), "right": todo0[:, -1] & (r2.random(R) < 0.9),
                "top": todo0[0] & (r2.random(C) < 0.9), "bottom": todo0[-1] & (r2.random(C) < 0.9)}
        dp2 = make(E.copy(), direction=np.array(dp.direction), mag=np.array(dp.mag), **kw)
        dp2.find_flats()
        dp2.calc_uca(uca_init=uca0.copy(), edge_init_data=[data, done, todo])
        return np.array(dp2.uca), np.array(dp2.edge_todo), np.array(dp2.edge_done)
