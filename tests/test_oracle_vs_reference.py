"""Oracle against the live reference (only where /root/reference exists, i.e. the build
container; the GPU box skips this module)."""
import numpy as np
import pytest

import helpers
import tiling
from oracle import ref_harness as rh
from oracle.oracle import OracleDEMProcessor

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")

NAMES = ["cone256", "frac128", "frac128_nopits", "frac256_dx30", "frac_rect_vardx", "nan_holes", "quantized",
         "lakes", "lakes_minborder", "frac96_maxdist4", "frac96_xy", "odd_cols", "tiny3", "frac96_maxcount1",
         "frac96_maxcount2"]


@pytest.mark.parametrize("name", NAMES)
def test_oracle_equals_reference(name):
    E, kw = helpers.cases()[name]
    ref = helpers.run(lambda e, **k: rh.ref_processor(e, **k), E, kw)
    got = helpers.run(lambda e, **k: OracleDEMProcessor(e, **k), E, kw)
    helpers.assert_parity(helpers.compare(ref, got), name)


def test_oracle_update_sequence_equals_reference():
    E = helpers.synth.fractal_dem(64, 8)
    kw = dict(helpers.HOT, drain_pits=False)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        sr, lr, _ = tiling.tiled_rows(lambda e, **k: rh.ref_processor(e, **k), E, 4, 2, kw)
    so, lo, _ = tiling.tiled_rows(lambda e, **k: OracleDEMProcessor(e, **k), E, 4, 2, kw)
    assert lr == lo and len(lr) > 0
    for a, b in zip(sr, so):
        np.testing.assert_allclose(a["uca0"] + a["edges"], b["uca0"] + b["edges"], rtol=1e-12, equal_nan=True)
        np.testing.assert_array_equal(a["todo"], b["todo"])
        np.testing.assert_array_equal(a["done"], b["done"])


def test_reference_cyutils_matches_oracle_sweep():
    """The reference's compiled Cython drain_area (oracle/_ref) against the oracle's frontier-list
    restatement on the same CSC/CSR matrix."""
    from oracle import oracle as orc
    cy = rh.load_ref_cyutils()
    E = helpers.synth.fractal_dem(96, 12)
    dp = OracleDEMProcessor(E, drain_pits=False, **helpers.HOT)
    dp.calc_slopes_directions()
    g, sec = dp._graph()
    cptr, cidx, cdat, rptr, ridx = g.export()
    inflow, _, _ = g.sums()
    R, C = E.shape
    for skip in (False,):
        area1 = np.ones(R * C); done1 = np.ascontiguousarray(inflow == 0, np.uint8); ids1 = done1.copy()
        area2 = area1.copy(); done2 = done1.copy().astype(bool); ids2 = ids1.copy().astype(bool)
        g.drain_area(area1, done1, ids1, R, C)
        cy.drain_area(area2, done2, ids2, cptr.astype(np.int32), cidx.astype(np.int32), cdat,
                      rptr.astype(np.int32), ridx.astype(np.int32), R, C)
        np.testing.assert_array_equal(area1, area2)     # same order of additions -> bit-identical
        np.testing.assert_array_equal(done1.astype(bool), done2)
