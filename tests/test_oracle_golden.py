"""CPU suite: the oracle (oracle/pdm_oracle.c + oracle/oracle.py) against the committed
fixtures that tests/golden/make_golden.py produced from the unmodified reference."""
import os

import numpy as np
import pytest

import helpers
import tiling
from oracle.oracle import OracleDEMProcessor

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_oracle(e, **k):
    return OracleDEMProcessor(e, **k)


def test_reference_known_answers():
    """The reference's own 5x5 vectors (test_end_to_end.py:152-182, 220-251), four orientations,
    checked the way the reference checks them (6 decimals)."""
    g = np.load(os.path.join(GOLD, "ref_known_answers.npz"))
    for nm in ("cardinal", "diagonal"):
        elev, ang, mag, uca = (g["%s_%s" % (nm, k)] for k in ("elev", "ang", "mag", "uca"))
        views = [(lambda a: a, True), (lambda a: a[::-1], False), (lambda a: a.T, False),
                 (lambda a: a[::-1, ::-1].T, False)]
        for view, check_ang in views:
            dp = OracleDEMProcessor(view(elev), fill_flats=False, drain_pits_path=False)
            m, a = dp.calc_slopes_directions()
            np.testing.assert_array_almost_equal(m, view(mag))
            if check_ang:
                np.testing.assert_array_almost_equal(a, view(ang))
            np.testing.assert_array_almost_equal(dp.calc_uca(), view(uca))


@pytest.mark.parametrize("name", sorted(helpers.golden_cases()))
def test_oracle_matches_reference_fixture(name):
    g = np.load(os.path.join(GOLD, "ref_cases.npz"))
    E, kw = helpers.golden_cases()[name]
    np.testing.assert_array_equal(np.nan_to_num(E, nan=-7.0), np.nan_to_num(g[name + "__elev"], nan=-7.0))
    ref = {k.split("__", 1)[1]: g[k] for k in g.files if k.startswith(name + "__")}
    ref["twi_min_area"] = float(ref["twi_min_area"])
    got = helpers.run(make_oracle, E, kw)
    helpers.assert_parity(helpers.compare(ref, got), name)
    # integer facet index of every cell (DEMProcessor.section), bit-exact
    from oracle import oracle as orc
    k = dict(helpers.HOT); k.update(kw)
    dp = OracleDEMProcessor(E, **k); dp.calc_slopes_directions()
    sec, prop = orc.section_proportion(dp.direction, dp.flats, dp.dX, dp.dY)
    np.testing.assert_array_equal(sec, ref["section"])


def test_oracle_update_mode_fixture():
    """Every calc_uca(uca_init, edge_init_data) call of a row-tiled run through the reference."""
    g = np.load(os.path.join(GOLD, "ref_update.npz"))
    E = g["elev"]
    n = int(g["n_calls"])
    assert n >= 3
    for c in range(n):
        f = lambda k: g["call%d__%s" % (c, k)]
        t, b = f("block")
        dp = OracleDEMProcessor(E[t:b], direction=f("direction").copy(), mag=f("mag").copy(), drain_pits=False,
                                **helpers.HOT)
        dp.find_flats()
        data, done, todo = ({s: f("%s_%s" % (nm, s)) for s in ("left", "right", "top", "bottom")}
                            for nm in ("data", "done", "todo"))
        uca = dp.calc_uca(uca_init=f("uca_init"), edge_init_data=[data, done, todo])
        np.testing.assert_allclose(uca, f("out_uca"), rtol=1e-12, equal_nan=True)
        np.testing.assert_array_equal(dp.edge_todo, f("out_todo"))
        np.testing.assert_array_equal(dp.edge_done, f("out_done"))


def test_oracle_tiled_equals_single_tile_on_cone():
    """The reference's own multi-tile criterion (test_end_to_end.py:96): tiled uca[1:-1,1:-1]
    equals the single-tile result on the analytic cone."""
    E = helpers.synth.cone_dem(48) * 100 + 1
    kw = dict(helpers.HOT, drain_pits=False)
    single = OracleDEMProcessor(E, **kw); single.calc_slopes_directions(); u = single.calc_uca()
    st, log, _ = tiling.tiled_rows(make_oracle, E, 3, 2, kw)
    np.testing.assert_array_almost_equal(tiling.stitch(st, *E.shape)[1:-1, 1:-1], u[1:-1, 1:-1])


def test_numpy_pairwise_sum_restatement():
    from oracle import oracle as orc
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 7, 8, 9, 15, 16, 17, 31, 33, 127, 128, 129, 130, 257, 1000):
        a = rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n)
        assert orc.lib().orc_np_sum(np.ascontiguousarray(a), n) == float(np.sum(a))


@pytest.mark.parametrize("name", ["frac128", "frac128_nopits", "nan_holes", "quantized", "lakes", "lakes_minborder", "frac_rect_vardx",
                                  "frac96_maxdist4", "odd_cols", "cone256"])
def test_drainage_graph_is_acyclic(name):
    """The circular-reference restart of the reference (dem_processing.py:951-964) never fires: one call of
    drain_area finishes every cell, because an edge needs elev[j] <= elev[i] AND a weight > 1e-8, and a facet
    neighbour at the centre's own elevation always gets weight 0.  The CUDA path relies on it
    (cells left over after a sweep are an error there, not a case to replay)."""
    import warnings
    from oracle.oracle import OracleDEMProcessor
    E, kw = helpers.cases()[name]
    k = dict(helpers.HOT); k.update(kw)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dp = OracleDEMProcessor(E, **k)
        dp.calc_slopes_directions()
        dp.calc_uca()
    assert dp.done.all() and dp.stats["restarts"] == 0
