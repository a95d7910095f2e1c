"""GPU suite: the chain bursts of the accumulation sweep (csrc/drain_op.cuh, chain_warp) on terrain made for them --
long meandering valleys whose thalweg is a run of single-receiver cells with a tributary joining at every cell,
edge_todo taint entering at the top border and travelling down the river, non-unit weights where a receiver was
filtered, per-row spacing -- against the oracle, and against the strictly ordered sweep (which never bursts)."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def valley_dem(R, C, seed, n_valleys=3, noise=0.0):
    rng = np.random.default_rng(seed)
    i = np.arange(R)[:, None].astype("float64"); j = np.arange(C)[None, :].astype("float64")
    E = 2.0 * (R - i) + 50.0                                   # tilted plane: everything drains towards the last row
    cross = np.full((R, C), np.inf)
    for k in range(n_valleys):
        centre = (k + 0.5) * C / n_valleys + (C / (4.0 * n_valleys)) * np.sin(i / (25.0 + 7 * k) + k)
        cross = np.minimum(cross, np.abs(j - centre))
    E = E + 0.9 * cross                                        # V-shaped valleys with meandering thalwegs
    if noise:
        E = E + noise * rng.random((R, C))
    return E


CASES = {
    "valleys": (valley_dem(768, 640, 1), dict(dX=30.0, dY=30.0)),
    "valleys_noise": (valley_dem(512, 512, 2, noise=0.3), dict(dX=30.0, dY=30.0)),
    "valleys_vardx": (valley_dem(400, 300, 3, n_valleys=2, noise=0.05),
                      dict(dX=np.linspace(20, 30, 399), dY=np.full(399, 27.3), dX2=np.linspace(20, 30, 400), dY2=np.full(400, 27.3))),
    "valleys_nopits": (valley_dem(512, 384, 4, noise=0.2), dict(dX=10.0, dY=12.5, drain_pits=False)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_rivers_equal_oracle(cuda_lib, name):
    from pydem_b200 import DEMProcessor
    from oracle.oracle import OracleDEMProcessor
    E, kw = CASES[name]
    ref = helpers.run(lambda e, **k: OracleDEMProcessor(e, **k), E, kw)
    got = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), E, kw)
    helpers.assert_parity(helpers.compare(ref, got), name)
    # the terrain does what it was made for: a river collects a large share of the tile, tainted from the top border
    assert np.nanmax(ref["uca"]) > 0.05 * E.size * np.nanmin(ref["uca"])
    assert (~np.asarray(ref["edge_done"], bool)).sum() > E.shape[0] // 2


def test_bursts_equal_strict_sweep(cuda_lib):
    """strict=True holds every count-off back until its adds have returned and never uses the bursts: same sums up to
    re-association, same masks."""
    import pydem_b200
    from pydem_b200 import DEMProcessor
    E, kw = CASES["valleys"]
    a = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), E, kw)
    pydem_b200.set_sweep("worklist", strict=True)
    try:
        b = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), E, kw)
    finally:
        pydem_b200.set_sweep("env")
    np.testing.assert_allclose(a["uca"], b["uca"], rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(a["edge_done"], b["edge_done"])
    np.testing.assert_array_equal(a["edge_todo"], b["edge_todo"])
