"""Tile orchestrator (pydem_b200.process_manager) on the CPU:
* the orchestrator driving the ORACLE operator reproduces the unmodified reference
  ProcessManager array by array (golden fixture tests/golden/ref_pm.npz, made by
  tests/golden/make_golden_pm.py) including the order in which tiles were corrected;
* where the reference tree is present: the reference's ProcessManager runs live over the
  in-memory store, (a) with its own DEMProcessor and (b) with the oracle operator swapped in at
  ``pydem.process_manager.DEMProcessor`` -- the drop-in point INTEGRATION.md names."""
import contextlib
import io
import os
import warnings

import numpy as np
import pytest

import helpers
from oracle.oracle import OracleDEMProcessor
from oracle import ref_harness
from pydem_b200.process_manager import ProcessManager, split_mosaic

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pm.npz"))
CASES = helpers.pm_cases()
# oracle operator vs reference operator: direction 8.9e-16, uca rel 8e-15, masks exact
TOL = dict(elev=0.0, slope=1e-13, aspect=1e-12, uca=1e-9, uca_edges=1e-9, edge_todo=0, edge_done=0, twi=1e-8)


def oracle_factory(**k):
    return OracleDEMProcessor(k.pop("elev"), **k)


def run_pm(E, boxes, kw, factory):
    tiles = [E[b[0]:b[1], b[2]:b[3]] for b in boxes]
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        pm = ProcessManager(tiles, boxes, dem_proc_kwargs=kw, dem_processor=factory)
        pm.process_twi()
    return pm


@pytest.mark.parametrize("name", sorted(CASES))
def test_split_matches_reference_generator(name):
    E, nx, ny, ov, kw = CASES[name]
    assert sorted(split_mosaic(E.shape, ny, nx, ov)) == sorted(map(tuple, G[name + "_boxes"].tolist()))


@pytest.mark.parametrize("name", sorted(CASES))
def test_orchestrator_equals_reference_process_manager(name):
    E, nx, ny, ov, kw = CASES[name]
    np.testing.assert_array_equal(E, G[name + "_E"])
    boxes = [tuple(b) for b in G[name + "_boxes"].tolist()]
    pm = run_pm(E, boxes, kw, oracle_factory)
    worst = helpers.pm_compare(pm, G, name)
    for k, tol in TOL.items():
        assert worst[k] <= tol, (name, worst)
    assert pm.correction_log == G[name + "_order"].tolist(), name          # same scheduling decisions
    assert pm.success.all()
    m = pm.mosaic("uca"); c = G[name + "_compact_uca"]
    assert m.shape == c.shape
    np.testing.assert_allclose(m, c, rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(pm.mosaic("twi"), G[name + "_compact_twi"], atol=1e-8, equal_nan=True)


def test_orchestrator_with_raster_derived_spacing():
    """Non-unit, anisotropic spacing (what the reference derives from a projected raster) through the
    whole orchestration: pixel 1.5/63 deg wide, 1/63 deg high."""
    name = helpers.PM_SPACING_CASE
    E = G[name + "_E"]
    boxes = [tuple(b) for b in G[name + "_boxes"].tolist()]
    tiles = [E[b[0]:b[1], b[2]:b[3]] for b in boxes]
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        pm = ProcessManager(tiles, boxes, spacing=helpers.pm_spacing(E.shape, boxes), dem_processor=oracle_factory)
        pm.process_twi()
    worst = helpers.pm_compare(pm, G, name)
    for k, tol in TOL.items():
        assert worst[k] <= tol, (name, worst)
    assert pm.correction_log == G[name + "_order"].tolist()


@pytest.mark.parametrize("name", ["fractal_3x3_2overlap", "fractal_2x3_1overlap", "cone_5x4_3overlap"])
def test_overview_pyramids_equal_reference(name):
    """process_overviews (:933-991) on the compact mosaics: block means with the reference's chunk
    arithmetic and end-of-chunk leftovers."""
    from pydem_b200.process_manager import overview_pyramid
    E, nx, ny, ov, kw = CASES[name]
    chunks = [E.shape[0] // ny, E.shape[1] // nx]
    # the arithmetic itself, on the reference's own mosaics: bit for bit
    for key in ("elev", "uca"):
        mine = overview_pyramid(G["%s_compact_%s" % (name, key)], chunks, (3, 9))
        for o in (3, 9):
            np.testing.assert_array_equal(mine[o], G["%s_overview_%s_%d" % (name, key, o)])
    # and through the orchestrator
    pm = run_pm(E, [tuple(b) for b in G[name + "_boxes"].tolist()], kw, oracle_factory)
    res = pm.process_overviews(keys=("elev", "uca"), overviews=(3, 9))
    for o in (3, 9):
        np.testing.assert_array_equal(res["elev"][o], G["%s_overview_elev_%d" % (name, o)])
        np.testing.assert_allclose(res["uca"][o], G["%s_overview_uca_%d" % (name, o)], rtol=1e-9, equal_nan=True)


def test_reference_criterion_on_the_cone():
    """test_end_to_end.py:96: the mosaic's uca equals the single-tile uca away from the rim."""
    E, nx, ny, ov, kw = CASES["cone_5x4_2overlap"]
    pm = run_pm(E, split_mosaic(E.shape, ny, nx, ov), kw, oracle_factory)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        dp = OracleDEMProcessor(E.copy(), dX=1.0, dY=1.0)
        dp.calc_twi()
    np.testing.assert_array_almost_equal(dp.uca[1:-1, 1:-1], pm.mosaic("uca")[1:-1, 1:-1])


def test_rejects_inconsistent_tiles():
    E = CASES["cone_3x3_2overlap"][0]
    with pytest.raises(ValueError):
        ProcessManager([E[:10, :10]], [(0, 12, 0, 10)], dem_processor=oracle_factory)
    with pytest.raises(ValueError):
        ProcessManager([E[:10, :10], E[:12, 8:20]], [(0, 10, 0, 10), (0, 12, 8, 20)], dem_processor=oracle_factory)


@pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", ["cone_3x3_1overlap", "fractal_3x3_2overlap"])
def test_operator_drops_in_under_the_reference_process_manager(name):
    """The UNMODIFIED reference ProcessManager with this repo's operator interface swapped in at
    pydem.process_manager.DEMProcessor (the oracle stands in for the CUDA operator on the CPU;
    tests/test_gpu_process_manager.py runs the CUDA operator through the same orchestration)."""
    from oracle import ref_pm_harness as H
    E, nx, ny, ov, kw = CASES[name]

    class Adapter(OracleDEMProcessor):
        def __init__(self, **k):
            k.pop("bounds", None); k.pop("transform", None)
            super().__init__(k.pop("elev"), **k)

    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        r = H.run_reference_pm(E, nx, ny, ov, "dropin_" + name, dem_processor=Adapter, dem_proc_kwargs=kw)
    assert r["success"].all()
    assert r["correction_order"] == G[name + "_order"].tolist()
    for key in ("elev", "slope", "edge_todo", "edge_done"):
        np.testing.assert_array_equal(r[key], G["%s_%s" % (name, key)])
    np.testing.assert_allclose(r["aspect"], G[name + "_aspect"], atol=1e-12)
    np.testing.assert_allclose(r["uca"] + r["uca_edges"], G[name + "_uca"] + G[name + "_uca_edges"], rtol=1e-9, equal_nan=True)


@pytest.mark.parametrize("stop_after", [0, 1, 2, "edges", 3])
def test_directory_store_resumes_like_the_reference(tmp_path, stop_after):
    """out_path: every stage writes the per-tile arrays it produced and marks the tile in `success` (the reference's
    zarr arrays + success array, process_manager.py:362-381, 998-1007, 1251-1288).  A manager created on the store of
    an interrupted run does not repeat a finished stage and ends with the same arrays, the same mosaic and the same
    order of corrections as an uninterrupted run."""
    name = sorted(CASES)[0]
    E, nx, ny, ov, kw = CASES[name]
    boxes = [tuple(b) for b in G[name + "_boxes"].tolist()]
    tiles = [E[b[0]:b[1], b[2]:b[3]] for b in boxes]
    calls = []

    def counting_factory(**k):
        calls.append(1)
        return oracle_factory(**k)

    store = str(tmp_path / "store")
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        ref = run_pm(E, boxes, kw, oracle_factory)                         # uninterrupted, no store
        pm1 = ProcessManager(tiles, boxes, dem_proc_kwargs=kw, dem_processor=counting_factory, out_path=store)
        pm1.process_elevation()
        if stop_after != 0:
            pm1.process_aspect_slope()
        if stop_after in (2, "edges", 3):
            pm1.process_uca()
        if stop_after in ("edges", 3):
            pm1.process_uca_edges()
        if stop_after == 3:
            pm1.process_twi()
        n1 = len(calls)
        del pm1                                                            # "the process died here"
        pm2 = ProcessManager(tiles, boxes, dem_proc_kwargs=kw, dem_processor=counting_factory, out_path=store)
        done_cols = {0: 1, 1: 2, 2: 3, "edges": 3, 3: 4}[stop_after]
        assert pm2.success[:, :done_cols].all() and not pm2.success[:, done_cols:].any()
        pm2.process_twi()
        n2 = len(calls) - n1
        pm_fresh_calls = []
        pm3 = ProcessManager(tiles, boxes, dem_proc_kwargs=kw, dem_processor=lambda **k: (pm_fresh_calls.append(1), oracle_factory(**k))[1])
        pm3.process_twi()
    assert n1 + n2 == len(pm_fresh_calls)                                  # nothing was computed twice
    if stop_after == 3:
        assert n2 == 0
    assert pm2.success.all() and pm2.correction_log == ref.correction_log
    for key in ("elev", "aspect", "slope", "uca", "uca_edges", "edge_todo", "edge_done", "twi"):
        for a, b in zip(pm2.tiles, ref.tiles):
            np.testing.assert_array_equal(np.asarray(getattr(a, key)), np.asarray(getattr(b, key)), err_msg=key)
    np.testing.assert_array_equal(pm2.mosaic("uca"), ref.mosaic("uca"))
    with pytest.raises(ValueError):
        ProcessManager(tiles[:-1], boxes[:-1], dem_proc_kwargs=kw, dem_processor=oracle_factory, out_path=store)
