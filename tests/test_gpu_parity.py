"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same
seeded inputs, against the committed reference fixtures, and -- at benchmark sizes -- through
size-independent properties."""
import ctypes as ct
import os

import numpy as np
import pytest

import helpers
import tiling
from oracle.oracle import OracleDEMProcessor

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_gpu(e, **k):
    from pydem_b200 import DEMProcessor
    return DEMProcessor(elev=e, **k)


def make_oracle(e, **k):
    return OracleDEMProcessor(e, **k)


def test_reference_known_answers(cuda_lib):
    """The reference's own 5x5 known-answer test (test_end_to_end.py:152-287) on the GPU path."""
    g = np.load(os.path.join(GOLD, "ref_known_answers.npz"))
    for nm in ("cardinal", "diagonal"):
        elev, ang, mag, uca = (g["%s_%s" % (nm, k)] for k in ("elev", "ang", "mag", "uca"))
        views = [(lambda a: a, True), (lambda a: a[::-1], False), (lambda a: a.T, False),
                 (lambda a: a[::-1, ::-1].T, False)]
        for view, check_ang in views:
            dp = make_gpu(np.ascontiguousarray(view(elev)), fill_flats=False, drain_pits_path=False)
            m, a = dp.calc_slopes_directions()
            np.testing.assert_array_almost_equal(m, view(mag))
            if check_ang:
                np.testing.assert_array_almost_equal(a, view(ang))
            np.testing.assert_array_almost_equal(dp.calc_uca(), view(uca))


@pytest.mark.parametrize("name", sorted(helpers.golden_cases()))
def test_gpu_matches_reference_fixture(cuda_lib, name):
    """GPU vs outputs of the unmodified reference (tests/golden/ref_cases.npz)."""
    g = np.load(os.path.join(GOLD, "ref_cases.npz"))
    E, kw = helpers.golden_cases()[name]
    ref = {k.split("__", 1)[1]: g[k] for k in g.files if k.startswith(name + "__")}
    ref["twi_min_area"] = float(ref["twi_min_area"])
    got = helpers.run(make_gpu, E, kw)
    helpers.assert_parity(helpers.compare(ref, got), name)
    k = dict(helpers.HOT); k.update(kw)
    dp = make_gpu(E, **k); dp.calc_slopes_directions()
    np.testing.assert_array_equal(dp.calc_section_proportion(), ref["section"])   # facet index: bit-exact


@pytest.mark.parametrize("name", sorted(helpers.cases()))
def test_gpu_matches_oracle(cuda_lib, name):
    E, kw = helpers.cases()[name]
    ref = helpers.run(make_oracle, E, kw)
    got = helpers.run(make_gpu, E, kw)
    helpers.assert_parity(helpers.compare(ref, got), name)


def test_gpu_update_mode_fixture(cuda_lib):
    """calc_uca(uca_init, edge_init_data) against every call recorded from the reference."""
    g = np.load(os.path.join(GOLD, "ref_update.npz"))
    E = g["elev"]
    for c in range(int(g["n_calls"])):
        f = lambda k: g["call%d__%s" % (c, k)]
        t, b = f("block")
        dp = make_gpu(np.ascontiguousarray(E[t:b]), direction=f("direction").copy(), mag=f("mag").copy(),
                      drain_pits=False, **helpers.HOT)
        dp.find_flats()
        data, done, todo = ({s: f("%s_%s" % (nm, s)) for s in ("left", "right", "top", "bottom")}
                            for nm in ("data", "done", "todo"))
        uca = dp.calc_uca(uca_init=f("uca_init"), edge_init_data=[data, done, todo])
        np.testing.assert_allclose(uca, f("out_uca"), rtol=helpers.UCA_RTOL, equal_nan=True)
        np.testing.assert_array_equal(dp.edge_todo, f("out_todo"))
        np.testing.assert_array_equal(dp.edge_done, f("out_done"))


@pytest.mark.parametrize("case", ["domed", "rough", "pits"])
def test_gpu_update_sequence_matches_oracle(cuda_lib, case):
    """Whole cross-tile edge-resolution runs (same scheduling decisions, same per-block state)."""
    if case == "domed":
        E = helpers.synth.fractal_dem(96, 7) * 0.05 + helpers.synth.cone_dem(96) * 300 + 1
        kw = dict(helpers.HOT, drain_pits=False)
    elif case == "rough":
        E = helpers.synth.fractal_dem(128, 8)
        kw = dict(helpers.HOT, drain_pits=False)
    else:
        E = helpers.synth.fractal_dem(96, 9)
        kw = dict(helpers.HOT, drain_pits=True)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        so, lo, _ = tiling.tiled_rows(make_oracle, E, 4, 2, kw)
        sg, lg, _ = tiling.tiled_rows(make_gpu, E, 4, 2, kw)
    assert lo == lg and len(lo) > 0
    for a, b in zip(so, sg):
        np.testing.assert_allclose(a["uca0"] + a["edges"], b["uca0"] + b["edges"], rtol=helpers.UCA_RTOL, equal_nan=True)
        np.testing.assert_array_equal(a["todo"], b["todo"])
        np.testing.assert_array_equal(a["done"], b["done"])


def test_gpu_matches_oracle_2048(cuda_lib):
    E = helpers.synth.fractal_dem(2048, 100)
    kw = dict(dX=30.0, dY=30.0, drain_pits=False)
    ref = helpers.run(make_oracle, E, kw)
    got = helpers.run(make_gpu, E, kw)
    helpers.assert_parity(helpers.compare(ref, got), "frac2048")


def test_one_shot_c_abi_calls(cuda_lib):
    """The host-buffer entry points of include/pydem_b200.h, straight through ctypes."""
    L = cuda_lib.load()
    E = helpers.synth.fractal_dem(0, 3, shape=(90, 70))
    R, C = E.shape
    ref = helpers.run(make_oracle, E, dict(dX=30.0, dY=25.0, drain_pits=False))
    dX = np.full(R - 1, 30.0); dY = np.full(R - 1, 25.0); dX2 = np.full(R, 30.0); dY2 = np.full(R, 25.0)
    mag = np.empty_like(E); direction = np.empty_like(E); flats = np.empty(E.shape, np.uint8)
    P = cuda_lib.ptr
    cuda_lib.check(L.pdm_slopes_directions(P(E), R, C, P(dX), P(dY), None, None, P(mag), P(direction), P(flats)))
    np.testing.assert_array_equal(flats.astype(bool), ref["flats0"])
    np.testing.assert_allclose(mag, ref["mag0"], rtol=helpers.MAG_RTOL)
    np.testing.assert_allclose(direction, ref["dir"], atol=helpers.DIR_ATOL)
    p = cuda_lib.UcaParams(); L.pdm_default_uca_params(ct.byref(p)); p.drain_pits = 0
    st = cuda_lib.UcaStats()
    uca = np.empty_like(E); todo = np.empty(E.shape, np.uint8); done = np.empty(E.shape, np.uint8)
    cuda_lib.check(L.pdm_uca(P(E), P(direction), P(mag), P(flats), R, C, P(dX), P(dY), P(dX2), P(dY2), None, None,
                             ct.byref(p), P(uca), P(todo), P(done), ct.byref(st)))
    np.testing.assert_allclose(uca, ref["uca"], rtol=helpers.UCA_RTOL, equal_nan=True)
    np.testing.assert_array_equal(todo.astype(bool), ref["edge_todo"])
    np.testing.assert_array_equal(done.astype(bool), ref["edge_done"])
    assert st.n_cells == R * C and st.n_drained == R * C and st.n_undone == 0
    q = cuda_lib.TwiParams(); L.pdm_default_twi_params(ct.byref(q)); q.twi_min_area = st.min_area
    twi = np.empty_like(E)
    cuda_lib.check(L.pdm_twi(P(uca), P(mag), R * C, ct.byref(q), P(twi)))
    np.testing.assert_allclose(twi, ref["twi"], atol=helpers.TWI_ATOL, equal_nan=True)
    # error behaviour: bad shape -> ValueError, message available
    with pytest.raises(ValueError):
        cuda_lib.check(L.pdm_slopes_directions(P(E), 2, 2, P(dX), P(dY), None, None, P(mag), P(direction), P(flats)))


def test_properties_at_4096(cuda_lib):
    """BASELINE.json config 2 size: properties that do not need the oracle.
    * linearity: doubling dX2*dY2 doubles uca
    * mass balance: every cell's area leaves through cells without a kept receiver
    * rerun determinism of the integer outputs, uca reproducible to fp64 re-association
    * every cell drained exactly once (acyclic graph)"""
    E = helpers.synth.fractal_dem(4096, 0)
    kw = dict(dX=30.0, dY=30.0, drain_pits=False, **helpers.HOT)
    dp = make_gpu(E, **kw)
    dp.calc_twi()
    st = dp.uca_stats
    assert st["n_drained"] == E.size and st["n_undone"] == 0
    uca = dp.uca.copy()
    assert np.isnan(uca).sum() == dp.flats.sum()
    assert np.nanmin(uca) >= 900.0 - 1e-9
    dp2 = make_gpu(E, dX=30.0, dY=30.0, dX2=np.full(4096, 60.0), dY2=np.full(4096, 30.0), drain_pits=False, **helpers.HOT)
    dp2.calc_twi()
    np.testing.assert_array_equal(dp2.flats, dp.flats)
    np.testing.assert_array_equal(dp2.edge_todo, dp.edge_todo)
    np.testing.assert_array_equal(dp2.edge_done, dp.edge_done)
    np.testing.assert_allclose(dp2.uca, 2.0 * uca, rtol=1e-10, equal_nan=True)
    np.testing.assert_array_equal(dp2.mag, dp.mag)
    # twi consistent with its definition on the host
    with np.errstate(invalid="ignore", divide="ignore"):
        np.testing.assert_allclose(dp.twi, 10 * np.log(uca / (dp.mag + 1e-3)), rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("name", ["card", "diag", "cone256", "frac256_dx30", "frac_rect_vardx", "nan_holes", "quantized",
                                  "lakes", "odd_cols", "tiny3", "frac512_limits"])
def test_fast_stencil_equals_literal_stencil(cuda_lib, name):
    """The division-/atan2-saving interior stencil must reproduce the literal 8 x (3 div + atan2)
    formulation bit for bit (mag, direction, flats)."""
    from pydem_b200 import tile as T
    E, kw = helpers.cases()[name]
    R, C = E.shape
    out = []
    for parity in (0, 1):
        dt = T.DeviceTile(R, C)
        dt.set_spacing(kw.get("dX", 1.0), kw.get("dY", 1.0), kw.get("dX2"), kw.get("dY2"))
        dt.set_stencil_parity(parity)
        dt.upload(T.F_ELEV, E)
        dt.slopes_directions()
        out.append((dt.download(T.F_MAG), dt.download(T.F_DIR), dt.download(T.F_FLATS)))
        dt.close()
    for a, b in zip(*out):
        np.testing.assert_array_equal(a, b)


def test_fast_stencil_extreme_values(cuda_lib):
    """Tiny / huge differences, exact ties and strongly anisotropic cells: the fast path must either
    agree bit for bit or hand the cell to the literal path."""
    from pydem_b200 import tile as T
    rng = np.random.default_rng(3)
    base = helpers.synth.fractal_dem(96, 50)
    cases = [(base * 1e-160, 1.0, 1.0), (base * 1e140, 1.0, 1.0), (np.round(base), 1.0, 1.0), (np.round(base), 30.0, 10.0),
             (base, 1e-3, 1e3), (base, 1.0, 1e-7), (np.floor(base / 7) + rng.integers(0, 2, base.shape) * 1e-300, 2.0, 3.0)]
    for E, dX, dY in cases:
        res = []
        for parity in (0, 1):
            dt = T.DeviceTile(*E.shape)
            dt.set_spacing(dX, dY)
            dt.set_stencil_parity(parity)
            dt.upload(T.F_ELEV, E)
            dt.slopes_directions()
            res.append((dt.download(T.F_MAG), dt.download(T.F_DIR)))
            dt.close()
        np.testing.assert_array_equal(res[0][0], res[1][0])
        np.testing.assert_array_equal(res[0][1], res[1][1])


def test_reciprocal_division_is_ieee_exact(cuda_lib):
    """The stencil's 5-operation division (Markstein sequence on a correctly rounded reciprocal)
    against __ddiv_rn on 2e9 operand pairs incl. extreme mantissas/exponents: bit-identical."""
    bad = ct.c_ulonglong(123)
    cuda_lib.check(cuda_lib.load().pdm_selftest_division(2024, 2_000_000_000, ct.byref(bad)))
    assert bad.value == 0
