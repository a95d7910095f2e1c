"""Randomised pinning of the tile orchestrator against the live, unmodified reference
ProcessManager (this container only): random DEMs (fractal, cone, tilted plane + noise, quantised),
grids of 1..3 x 1..3 tiles, overlaps 0..4, random conditioning / pit flags.  Elevation and edge
masks bit for bit, the other arrays within the oracle-vs-reference noise, the same order of
corrections and the same tile boxes.  (300 such cases were checked once; 36 run here.)"""
import contextlib
import io
import warnings

import numpy as np
import pytest

import helpers
from oracle import ref_harness
from oracle.oracle import OracleDEMProcessor
from pydem_b200.process_manager import ProcessManager, split_mosaic

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present (GPU box)")


def _case(seed):
    rng = np.random.default_rng(10_000 + seed)
    R, C = int(rng.integers(24, 60)), int(rng.integers(24, 60))
    kind = int(rng.integers(0, 4))
    E = helpers.synth.fractal_dem(0, 500 + seed, shape=(R, C))
    if kind == 1:
        y, x = np.mgrid[-1:1:R * 1j, -1:1:C * 1j]
        E = 1 - np.sqrt(x ** 2 + y ** 2) / np.sqrt(2) + 0.001
    elif kind == 2:
        E = E * 0.002 + np.add.outer(np.arange(R) * rng.uniform(0.5, 2), np.arange(C) * rng.uniform(0.5, 2)) + 5
    elif kind == 3:
        E = np.round(E / 10) * 10 + 1
    ny, nx, ov = int(rng.integers(1, 4)), int(rng.integers(1, 4)), int(rng.integers(0, 5))
    kw = {}
    if rng.random() < 0.3: kw["drain_pits"] = False
    if rng.random() < 0.3: kw["fill_flats"] = False
    if rng.random() < 0.3: kw["drain_pits_path"] = False
    return E, ny, nx, ov, kw


@pytest.mark.parametrize("block", range(6))
def test_orchestrator_equals_reference_on_random_mosaics(block):
    from oracle import ref_pm_harness as H
    for seed in range(block * 6, block * 6 + 6):
        E, ny, nx, ov, kw = _case(seed)
        orig = np.argsort

        def stable(a, *args, **kws):        # the reference's pit order is only defined up to ties (test_oracle_fuzz.py)
            if "kind" not in kws and len(args) < 2:
                kws["kind"] = "stable"
            return orig(a, *args, **kws)
        np.argsort = stable
        try:
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r = H.run_reference_pm(E, nx, ny, ov, "fuzz%d" % seed, dem_proc_kwargs=kw)
        finally:
            np.argsort = orig
        boxes = r["boxes"]
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pm = ProcessManager([E[b[0]:b[1], b[2]:b[3]] for b in boxes], boxes, dem_proc_kwargs=kw,
                                dem_processor=lambda **k: OracleDEMProcessor(k.pop("elev"), **k))
            pm.process_twi()
        G = {"fz_" + k: v for k, v in r.items() if isinstance(v, np.ndarray)}
        G["fz_grid_slice"] = np.array(r["grid_slice"])
        w = helpers.pm_compare(pm, G, "fz")
        msg = "seed %d shape %s grid %dx%d overlap %d %s: %s" % (seed, E.shape, ny, nx, ov, kw, w)
        assert r["success"].all(), msg
        assert sorted(split_mosaic(E.shape, ny, nx, ov)) == sorted(boxes), msg
        assert pm.correction_log == r["correction_order"], msg
        assert w["elev"] == 0 and w["edge_todo"] == 0 and w["edge_done"] == 0, msg
        assert w["slope"] <= 1e-12 and w["aspect"] <= 1e-12 and w["uca"] <= 1e-9 and w["uca_edges"] <= 1e-9 and w["twi"] <= 1e-7, msg


@pytest.mark.parametrize("drop", [(4,), (0,), (8,), (5,), (1, 7)])
def test_mosaics_with_missing_tiles(drop):
    """The reference tolerates holes in the tile grid (grid_id2i == -1, process_manager.py:527, 613-655)."""
    from oracle import ref_pm_harness as H
    E = helpers.synth.fractal_dem(64, 3)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = H.run_reference_pm(E, 3, 3, 2, "hole" + "_".join(map(str, drop)), drop=drop)
        boxes = r["boxes"]
        pm = ProcessManager([E[b[0]:b[1], b[2]:b[3]] for b in boxes], boxes,
                            dem_processor=lambda **k: OracleDEMProcessor(k.pop("elev"), **k))
        pm.process_twi()
    assert r["success"].all() and len(boxes) == 9 - len(drop)
    G = {"fz_" + k: v for k, v in r.items() if isinstance(v, np.ndarray)}
    G["fz_grid_slice"] = np.array(r["grid_slice"])
    w = helpers.pm_compare(pm, G, "fz")
    assert pm.correction_log == r["correction_order"], w
    assert w["elev"] == 0 and w["edge_todo"] == 0 and w["edge_done"] == 0 and w["uca"] <= 1e-9 and w["uca_edges"] <= 1e-9 and w["aspect"] <= 1e-12, w
