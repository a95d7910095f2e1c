"""Tile-parallel mode of the orchestrator (pydem_b200.process_manager with group=...) on the CPU:
* the edge rings remote tiles are reduced to (_RingArray);
* two and three real processes over gloo (oracle operator): every rank ends with the same mosaic and
  the same order of corrections as the unmodified reference ProcessManager (golden fixture) and as
  the one-rank run bit for bit, and keeps full arrays only for its own tiles;
* the experimental correction rounds (not the default): equal to the reference on the pinned
  tilings, never two adjacent tiles in a round."""
import contextlib
import io
import os
import socket
import sys
import warnings

import numpy as np
import pytest

import helpers
from oracle.oracle import OracleDEMProcessor
from pydem_b200.process_manager import ProcessManager, _RingArray, split_mosaic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pm.npz"))
CASES = helpers.pm_cases()


def oracle_factory(**k):
    return OracleDEMProcessor(k.pop("elev"), **k)


def run_rounds(E, boxes, kw, group=None, rounds=True):
    tiles = [E[b[0]:b[1], b[2]:b[3]] for b in boxes]
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        pm = ProcessManager(tiles, boxes, dem_proc_kwargs=kw, dem_processor=oracle_factory, group=group)
        pm.process_twi(rounds=rounds)
    return pm


def test_ring_array_matches_full_array():
    rng = np.random.default_rng(0)
    a = rng.random((9, 7))
    r = _RingArray.of(a, 2)
    for key in ((slice(0, None), 0), (slice(0, None), -1), (0, slice(0, None)), (-1, slice(0, None)), (0, 0), (-1, -1), (1, slice(0, None)),
                (slice(0, None), 5), (7, 6)):
        np.testing.assert_array_equal(r[key], a[key])
    with pytest.raises(IndexError):
        r[4, slice(0, None)]
    # writes keep the four strips consistent (corner cells live in two of them)
    r[(slice(0, None), 0)] = np.arange(9.0); a[:, 0] = np.arange(9.0)
    r[(-1, slice(0, None))] = -np.arange(7.0); a[-1, :] = -np.arange(7.0)
    r[(0, -1)] = 42.0; a[0, -1] = 42.0
    b = _RingArray.of(a, 2)
    for s1, s2 in zip(r.strips(), b.strips()):
        np.testing.assert_array_equal(s1, s2)


@pytest.mark.parametrize("name", [n for n in sorted(CASES) if "0overlap" not in n])
def test_rounds_equal_the_reference_serial_loop(name):
    """Correction rounds (ties for the best metric, non-adjacent) against the unmodified reference
    ProcessManager's one-tile-at-a-time result, away from the rim of the mosaic."""
    E, nx, ny, ov, kw = CASES[name]
    boxes = [tuple(b) for b in G[name + "_boxes"].tolist()]
    pm = run_rounds(E, boxes, kw)
    assert all(isinstance(r, tuple) and len(r) >= 1 for r in pm.correction_log)
    m = pm.mosaic("uca")
    np.testing.assert_allclose(m[1:-1, 1:-1], G[name + "_compact_uca"][1:-1, 1:-1], rtol=1e-9, equal_nan=True)
    np.testing.assert_allclose(pm.mosaic("twi")[1:-1, 1:-1], G[name + "_compact_twi"][1:-1, 1:-1], atol=1e-8, equal_nan=True)
    if name.startswith("cone"):
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            dp = OracleDEMProcessor(E.copy(), dX=1.0, dY=1.0); dp.calc_twi()
        np.testing.assert_array_almost_equal(dp.uca[1:-1, 1:-1], m[1:-1, 1:-1])             # test_end_to_end.py:96


def test_rounds_resolve_every_edge_on_rough_terrain():
    E, nx, ny, ov, kw = CASES["fractal_3x3_2overlap"]
    pm = run_rounds(E, split_mosaic(E.shape, ny, nx, ov), kw)
    assert sum(int(t.edge_todo.sum()) for t in pm.tiles) == 0
    assert np.isfinite(pm.mosaic("twi")).mean() > 0.9
    # rounds never put two adjacent tiles together
    for rnd in pm.correction_log:
        for a in rnd:
            for b in rnd:
                if a != b:
                    ta, tb = pm.tiles[a], pm.tiles[b]
                    assert max(abs(ta.gi - tb.gi), abs(ta.gj - tb.gj)) > 1


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q, names):
    try:
        sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from pydem_b200.process_manager import TorchGroup
        g = TorchGroup()
        out = {}
        for name in names:
            E, nx, ny, ov, kw = CASES[name]
            boxes = [tuple(b) for b in G[name + "_boxes"].tolist()]
            pm = run_rounds(E, boxes, kw, group=g, rounds=False)       # the default: the reference's serial decisions
            own_full = all(isinstance(t.uca, np.ndarray) == (t.owner == rank) for t in pm.tiles)
            out[name] = (pm.mosaic("uca"), pm.mosaic("twi"), pm.mosaic("aspect"), list(pm.correction_log), own_full)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", out))
    except Exception:
        import traceback
        q.put((rank, "error", traceback.format_exc()))


@pytest.mark.parametrize("world", [2, 3])
def test_tile_parallel_over_gloo_equals_one_rank(world):
    import multiprocessing as mp
    names = ["cone_3x3_1overlap", "fractal_3x3_2overlap", "fractal_2x3_1overlap", "cone_3x3_0overlap", "fractal_1x3_4overlap"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q, names)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    for rank, status, payload in res:
        assert status == "ok", payload
    for name in names:
        E, nx, ny, ov, kw = CASES[name]
        one = run_rounds(E, [tuple(b) for b in G[name + "_boxes"].tolist()], kw, rounds=False)
        for rank, status, payload in res:
            uca, twi, aspect, log, own_full = payload[name]
            np.testing.assert_array_equal(uca, one.mosaic("uca"))              # = one rank, bit for bit
            np.testing.assert_array_equal(twi, one.mosaic("twi"))
            np.testing.assert_array_equal(aspect, one.mosaic("aspect"))
            assert log == list(one.correction_log) == G[name + "_order"].tolist()   # = the reference's scheduling
            np.testing.assert_allclose(uca, G[name + "_compact_uca"], rtol=1e-9, equal_nan=True)
            assert own_full
