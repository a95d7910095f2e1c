"""Randomised pinning of the oracle against the live, unmodified reference (this container only):
small random DEMs (plain, quantised, lakes, NaN holes, tilted planes, integer elevations) with
random spacings and flags, including the conditioning flags.  Every output is compared --
conditioned elevation and the masks bit for bit.

The reference visits pits of equal elevation in np.argsort's (unstable, platform-defined) order;
the oracle and the CUDA path use raster order.  For the comparison the reference is run with a
stable sort, which is the only way to make its result well defined on such inputs."""
import numpy as np
import pytest

import helpers
from oracle import ref_harness
from oracle.oracle import OracleDEMProcessor

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="reference tree not present (GPU box)")


def _reference_stable(E, kw):
    orig = np.argsort

    def stable(a, *args, **kws):
        if "kind" not in kws and len(args) < 2:
            kws["kind"] = "stable"
        return orig(a, *args, **kws)
    np.argsort = stable
    try:
        return helpers.run_all(lambda e, **k: ref_harness.ref_processor(e, **k), E, kw)
    finally:
        np.argsort = orig


@pytest.mark.parametrize("block", range(6))
def test_oracle_equals_reference_on_random_cases(block):
    for seed in range(block * 12, block * 12 + 12):
        E, kw, kind = helpers.fuzz_case(seed)
        ref = _reference_stable(E, kw)
        got = helpers.run_all(lambda e, **k: OracleDEMProcessor(e, **k), E, kw)
        r = helpers.compare(ref, got)
        msg = "seed %d kind %d shape %s: %s" % (seed, kind, E.shape, r)
        assert np.array_equal(ref["elev"], got["elev"], equal_nan=True), msg
        for k in ("flats0", "flats", "edge_todo", "edge_done"):
            assert r[k + "_neq"] == 0, msg
        for k in ("mag0", "mag", "dir", "uca", "twi"):
            assert r[k + "_nanpat"] == 0, msg
        assert r["mag_rel"] <= 1e-12 and r["dir_abs"] <= 1e-12 and r["uca_rel"] <= 1e-9 and r["twi_abs"] <= 1e-8, msg


@pytest.mark.parametrize("block", range(3))
def test_oracle_update_mode_equals_reference_on_random_edge_data(block):
    for seed in range(block * 10, block * 10 + 10):
        E, kw, kind = helpers.fuzz_case(seed)
        kw = dict(kw, fill_flats=False, drain_pits_path=False)
        u1, t1, d1 = helpers.update_call(lambda e, **k: ref_harness.ref_processor(e, **k), E, kw, seed)
        u2, t2, d2 = helpers.update_call(lambda e, **k: OracleDEMProcessor(e, **k), E, kw, seed)
        msg = "seed %d kind %d shape %s" % (seed, kind, E.shape)
        assert np.array_equal(np.isnan(u1), np.isnan(u2)), msg
        np.testing.assert_allclose(u2, u1, rtol=1e-9, equal_nan=True, err_msg=msg)
        assert np.array_equal(t1, t2) and np.array_equal(d1, d2), msg
