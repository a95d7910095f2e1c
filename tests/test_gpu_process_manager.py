"""Tile orchestrator with the CUDA operator (pydem_b200.DEMProcessor) on the GPU: every per-tile
array, the scheduling order and the mosaic against the unmodified reference ProcessManager
(tests/golden/ref_pm.npz), and against the same orchestration driven by the oracle on a rough
4x4 mosaic (BASELINE.json configs[4] in miniature)."""
import contextlib
import io
import os
import warnings

import numpy as np
import pytest

import helpers
from oracle.oracle import OracleDEMProcessor
from pydem_b200.process_manager import ProcessManager, split_mosaic

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pm.npz"))
CASES = helpers.pm_cases()
TOL = dict(elev=0.0, slope=helpers.MAG_RTOL, aspect=helpers.DIR_ATOL, uca=helpers.UCA_RTOL, uca_edges=helpers.UCA_RTOL,
           edge_todo=0, edge_done=0, twi=10 * helpers.TWI_ATOL)


def run_pm(E, boxes, kw, factory=None, spacing=None):
    tiles = [E[b[0]:b[1], b[2]:b[3]] for b in boxes]
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        pm = ProcessManager(tiles, boxes, spacing=spacing, dem_proc_kwargs=kw, dem_processor=factory)
        pm.process_twi()
    return pm


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_orchestration_equals_reference_process_manager(name):
    E, nx, ny, ov, kw = CASES[name]
    boxes = [tuple(b) for b in G[name + "_boxes"].tolist()]
    pm = run_pm(E, boxes, kw)
    worst = helpers.pm_compare(pm, G, name)
    for k, tol in TOL.items():
        assert worst[k] <= tol, (name, worst)
    assert pm.correction_log == G[name + "_order"].tolist(), name
    np.testing.assert_allclose(pm.mosaic("uca"), G[name + "_compact_uca"], rtol=helpers.UCA_RTOL, equal_nan=True)


def test_reference_criterion_on_the_cone_cuda():
    """test_end_to_end.py:96 with the CUDA operator on both sides."""
    from pydem_b200 import DEMProcessor
    E, nx, ny, ov, kw = CASES["cone_5x4_3overlap"]
    pm = run_pm(E, split_mosaic(E.shape, ny, nx, ov), kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dp = DEMProcessor(elev=E.copy(), dX=1.0, dY=1.0)
        dp.calc_twi()
    np.testing.assert_array_almost_equal(dp.uca[1:-1, 1:-1], pm.mosaic("uca")[1:-1, 1:-1])


def test_rough_mosaic_cuda_equals_oracle_orchestration():
    E = helpers.synth.conditioned_fractal_dem(384, 7)
    boxes = split_mosaic(E.shape, 4, 4, 2)
    kw = dict(fill_flats=False, drain_pits_path=False)
    sp = dict(dX=30.0, dY=30.0)
    a = run_pm(E, boxes, kw, spacing=sp)
    b = run_pm(E, boxes, kw, factory=lambda **k: OracleDEMProcessor(k.pop("elev"), **k), spacing=sp)
    assert a.correction_log == b.correction_log and len(a.correction_log) >= len(boxes)
    for ta, tb in zip(a.tiles, b.tiles):
        np.testing.assert_array_equal(ta.edge_todo, tb.edge_todo)
        np.testing.assert_array_equal(ta.edge_done, tb.edge_done)
        np.testing.assert_allclose(ta.uca + ta.uca_edges, tb.uca + tb.uca_edges, rtol=helpers.UCA_RTOL, equal_nan=True)
    np.testing.assert_allclose(a.mosaic("twi"), b.mosaic("twi"), atol=10 * helpers.TWI_ATOL, equal_nan=True)


@pytest.mark.parametrize("name", ["cone_5x4_2overlap", "fractal_3x3_2overlap", "fractal_2x3_1overlap"])
def test_cuda_correction_rounds_equal_reference(name):
    """Correction rounds (the tile-parallel schedule) with the CUDA operator vs the reference's serial loop."""
    E, nx, ny, ov, kw = CASES[name]
    boxes = [tuple(b) for b in G[name + "_boxes"].tolist()]
    tiles = [E[b[0]:b[1], b[2]:b[3]] for b in boxes]
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        pm = ProcessManager(tiles, boxes, dem_proc_kwargs=kw)
        pm.process_twi(rounds=True)
    m = pm.mosaic("uca")
    np.testing.assert_allclose(m[1:-1, 1:-1], G[name + "_compact_uca"][1:-1, 1:-1], rtol=helpers.UCA_RTOL, equal_nan=True)


@pytest.mark.parametrize("name", sorted(CASES))
def test_resident_orchestration_equals_reference_process_manager(name):
    """The device-resident manager (tiles stay in HBM, only edge rings and strips move, the graph of a tile's
    full sweep is kept for its corrections) against the unmodified reference ProcessManager: same order of
    corrections, same arrays."""
    from pydem_b200.process_manager import ResidentProcessManager
    E, nx, ny, ov, kw = CASES[name]
    boxes = [tuple(b) for b in G[name + "_boxes"].tolist()]
    tiles = [E[b[0]:b[1], b[2]:b[3]] for b in boxes]
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        pm = ResidentProcessManager(tiles, boxes, dem_proc_kwargs=kw)
        pm.process_twi()
    assert pm.correction_log == G[name + "_order"].tolist(), name
    gs = G[name + "_grid_slice"]
    for i, t in enumerate(pm.tiles):
        sl = (slice(gs[i][0], gs[i][1]), slice(gs[i][2], gs[i][3]))
        np.testing.assert_allclose(pm.full(t, "aspect"), G[name + "_aspect"][sl], atol=helpers.DIR_ATOL, equal_nan=True)
        np.testing.assert_allclose(pm.full(t, "uca"), G[name + "_uca"][sl] + G[name + "_uca_edges"][sl], rtol=helpers.UCA_RTOL, equal_nan=True)
        np.testing.assert_array_equal(pm.full(t, "edge_todo"), G[name + "_edge_todo"][sl])
        np.testing.assert_array_equal(pm.full(t, "edge_done"), G[name + "_edge_done"][sl])
        np.testing.assert_allclose(pm.full(t, "twi"), G[name + "_twi"][sl], atol=10 * helpers.TWI_ATOL, equal_nan=True)
    np.testing.assert_allclose(pm.mosaic("uca"), G[name + "_compact_uca"], rtol=helpers.UCA_RTOL, equal_nan=True)
    pm.close()


def test_resident_rough_mosaic_equals_host_orchestration():
    from pydem_b200.process_manager import ResidentProcessManager
    E = helpers.synth.conditioned_fractal_dem(384, 7)
    boxes = split_mosaic(E.shape, 4, 4, 2)
    kw = dict(fill_flats=False, drain_pits_path=False)
    sp = dict(dX=30.0, dY=30.0)
    a = run_pm(E, boxes, kw, spacing=sp)
    tiles = [E[b[0]:b[1], b[2]:b[3]] for b in boxes]
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        b = ResidentProcessManager(tiles, boxes, spacing=sp, dem_proc_kwargs=kw)
        b.process_twi()
    assert a.correction_log == b.correction_log
    np.testing.assert_allclose(a.mosaic("uca"), b.mosaic("uca"), rtol=helpers.UCA_RTOL, equal_nan=True)
    np.testing.assert_allclose(a.mosaic("twi"), b.mosaic("twi"), atol=10 * helpers.TWI_ATOL, equal_nan=True)
    for ta, tb in zip(a.tiles, b.tiles):
        np.testing.assert_array_equal(ta.edge_done, b.full(tb, "edge_done"))
    assert b.bytes_moved["h2d"] < 1.5 * sum(t.size for t in tiles) * 8        # elevation once + rings and strips
    b.close()
