"""GPU suite (-m gpu): elevation conditioning on the device (pdm_tile_fill_pit_artifacts /
pdm_tile_fill_flats / pdm_tile_pit_drain_paths through the DEMProcessor mirror) against the
fixtures produced by the unmodified reference and against the pinned oracle restatement.
Everything here is bit-exact: conditioning only moves, compares, adds and divides f64."""
import os
import warnings

import numpy as np
import pytest

import helpers
from oracle import conditioning as oc
from oracle.oracle import OracleDEMProcessor

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_conditioning.npz"))
CASES = helpers.conditioning_cases()


def golden(name, key, base):
    out = base.copy()
    out.ravel()[GOLD["%s_%s_idx" % (name, key)]] = GOLD["%s_%s_val" % (name, key)]
    return out


def gpu(E, dX, dY, **kw):
    from pydem_b200 import DEMProcessor
    return DEMProcessor(elev=E.copy(), dX=dX, dY=dY, **kw)


@pytest.mark.parametrize("name", sorted(CASES))
def test_fill_pit_artifacts(cuda_lib, name):
    E, dX, dY = CASES[name]
    dp = gpu(E, dX, dY)
    dp.calc_fill_pit_artifacts()
    np.testing.assert_array_equal(dp.elev, golden(name, "art", E))


@pytest.mark.parametrize("name", sorted(CASES))
def test_fill_flats(cuda_lib, name):
    E, dX, dY = CASES[name]
    dp = gpu(E, dX, dY)
    dp.calc_fill_flats()
    np.testing.assert_array_equal(dp.elev, golden(name, "fill", E))


@pytest.mark.parametrize("name", sorted(CASES))
def test_pit_drain_paths(cuda_lib, name):
    E, dX, dY = CASES[name]
    filled = golden(name, "fill", E)
    dp = gpu(filled, dX, dY)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dp.calc_pit_drain_paths()
    want, n_bad, maxit = oc.pit_drain_paths(filled, dX, dY)            # raster-order ties, like the device
    np.testing.assert_array_equal(dp.elev, want)
    assert dp.cond_stats.get("n_pits_undrained", 0) == n_bad
    assert dp.cond_stats.get("path_max_iter", 0) == maxit
    if bool(GOLD[name + "_paths_tiefree"]):
        np.testing.assert_array_equal(dp.elev, golden(name, "paths", filled))   # the reference itself


@pytest.mark.parametrize("kw", [dict(fill_flats_peaks=False), dict(fill_flats_pits=False), dict(fill_flats_source_tol=3),
                                dict(maximum_pit_area=0), dict(maximum_pit_area=4.0), dict(fill_flats_below_sea=True)])
def test_fill_flats_flags(cuda_lib, kw):
    for name in ("lakes128", "peaks64", "artifacts64", "sea72"):
        E, dX, dY = CASES[name]
        dp = gpu(E, dX, dY, **kw)
        dp.calc_fill_flats()
        want = oc.fill_flats(E, kw.get("fill_flats_below_sea", False), kw.get("fill_flats_source_tol", 1),
                             kw.get("fill_flats_peaks", True), kw.get("fill_flats_pits", True), kw.get("maximum_pit_area", 32.0))
        np.testing.assert_array_equal(dp.elev, want, err_msg="%s %s" % (name, kw))


@pytest.mark.parametrize("kw", [dict(drain_pits_max_dist=3), dict(drain_pits_max_iter=5), dict(drain_pits_max_dist_XY=80.0),
                                dict(drain_pits_max_dist=None)])
def test_pit_drain_path_flags(cuda_lib, kw):
    for name in ("fractal96", "quant96", "lakes128"):
        E, dX, dY = CASES[name]
        dp = gpu(E, dX, dY, **kw)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dp.calc_pit_drain_paths()
        want = oc.pit_drain_paths(E, dX, dY, False, kw.get("drain_pits_max_iter", 300), kw.get("drain_pits_max_dist", 32),
                                  kw.get("drain_pits_max_dist_XY"))[0]
        np.testing.assert_array_equal(dp.elev, want, err_msg="%s %s" % (name, kw))


@pytest.mark.parametrize("name", ["fractal96", "lakes128", "peaks64", "sea72", "quant_fine80", "artifacts64"])
def test_default_flags_end_to_end(cuda_lib, name):
    """The reference's DEFAULT configuration (fill_flats=True, drain_pits_path=True, drain_pits=True):
    conditioning + slopes + UCA + TWI on the device against the oracle."""
    E, dX, dY = CASES[name]
    R = E.shape[0]
    kw = dict(dX=dX, dY=dY, dX2=np.full(R, 30.0), dY2=np.full(R, 28.0), fill_flats=True, drain_pits_path=True)
    from pydem_b200 import DEMProcessor
    ref = helpers.run(lambda e, **k: OracleDEMProcessor(e, **k), E, kw)
    got = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), E, kw)
    helpers.assert_parity(helpers.compare(ref, got), name)


def test_conditioning_512_lakes_and_quantised(cuda_lib):
    """A larger mixed case: lakes (r up to 40), quantised terraces and planted artifacts."""
    from pydem_b200 import synth
    E = np.round(synth.fractal_dem(512, 21) / 4.0) + 3
    yy, xx = np.mgrid[:512, :512]
    for (ci, cj, r) in ((100, 100, 40), (300, 380, 25), (420, 90, 17), (511, 300, 30), (60, 400, 12)):
        m = (yy - ci) ** 2 + (xx - cj) ** 2 <= r * r
        E[m] = E[m].min()
    dX = np.full(511, 30.0); dY = np.full(511, 30.0)
    dp = gpu(E, dX, dY)
    dp.calc_fill_flats()
    want = oc.fill_flats(E)
    np.testing.assert_array_equal(dp.elev, want)
    assert dp.cond_stats["n_flat_regions"] > 1000 and dp.cond_stats["distance_sweeps"] >= 40
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dp.calc_pit_drain_paths()
    np.testing.assert_array_equal(dp.elev, oc.pit_drain_paths(want, dX, dY)[0])


def test_conditioning_refused_on_a_shard(cuda_lib):
    from pydem_b200 import tile as T, _lib
    t = T.DeviceTile(34, 40)
    t.set_spacing(30.0, 30.0)
    t.upload(T.F_ELEV, np.random.default_rng(0).random((34, 40)) + 1)
    t.set_window(31, 100, 1, 33)
    import ctypes as ct
    with pytest.raises(_lib.PdmError):
        _lib.check(t.L.pdm_tile_fill_flats(t.h, None, None))
    t.close()
