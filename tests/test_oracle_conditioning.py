"""CPU suite: the conditioning restatement (oracle/conditioning.py) against the fixtures that
tests/golden/make_golden.py produced from the unmodified reference (calc_fill_pit_artifacts,
calc_fill_flats, calc_pit_drain_paths; dem_processing.py:396-585)."""
import os

import numpy as np
import pytest

import helpers
from oracle import conditioning as oc

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_conditioning.npz"))
CASES = helpers.conditioning_cases()


def golden(name, key, base):
    out = base.copy()
    out.ravel()[GOLD["%s_%s_idx" % (name, key)]] = GOLD["%s_%s_val" % (name, key)]
    return out


@pytest.mark.parametrize("name", sorted(CASES))
def test_fill_pit_artifacts_matches_reference(name):
    E = CASES[name][0]
    np.testing.assert_array_equal(oc.fill_pit_artifacts(E), golden(name, "art", E))


@pytest.mark.parametrize("name", sorted(CASES))
def test_fill_flats_matches_reference(name):
    E = CASES[name][0]
    np.testing.assert_array_equal(oc.fill_flats(E), golden(name, "fill", E))


@pytest.mark.parametrize("name", sorted(CASES))
def test_pit_drain_paths_matches_reference(name):
    E, dX, dY = CASES[name]
    filled = golden(name, "fill", E)
    want = golden(name, "paths", filled)
    got, undrained, maxit = oc.pit_drain_paths(filled, dX, dY)
    if bool(GOLD[name + "_paths_tiefree"]):
        np.testing.assert_array_equal(got, want)       # bit-exact, including the carved ramps
    else:
        # pits of equal elevation interact: the reference's visiting order among them is np.argsort's
        # platform-defined tie order, so only the cells no tied pit touches can be compared
        assert np.mean(got == want) > 0.9
    assert (got <= filled).sum() > 0


def test_fixture_covers_every_branch():
    """The fixture set exercises artifacts, peaks, edge-crossing lakes and pit paths."""
    assert sum(len(GOLD[k]) for k in GOLD.files if k.endswith("_art_idx")) > 50
    assert sum(len(GOLD[k]) for k in GOLD.files if k.endswith("_fill_idx")) > 1500
    assert sum(len(GOLD[k]) for k in GOLD.files if k.endswith("_paths_idx")) > 2000
