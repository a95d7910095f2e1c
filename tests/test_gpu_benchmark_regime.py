"""GPU suite: parity at the benchmark's own regime (BASELINE.json configs[1], primary variant): the
4096 x 4096 priority-flood conditioned fractal with the default drain_pits=True -- 3.9 k dependency
levels, rivers thousands of cells long -- compared with the oracle at FULL size (the oracle needs
~10 s for it); the sweep's determinism; the legacy (L2-atomic) sweep in its strictly ordered mode as
an independent second implementation; and the row-sharded path on a conditioned DEM."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 4096
KW = dict(dX=30.0, dY=30.0, fill_flats=False, drain_pits_path=False, drain_pits=True)


@pytest.fixture(scope="module")
def bench_dem():
    return helpers.synth.conditioned_fractal_dem(N, 0, wrap_rows=True)       # exactly bench.py's DEM


@pytest.fixture(scope="module")
def gpu_out(cuda_lib, bench_dem):
    from pydem_b200 import DEMProcessor
    return helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), bench_dem, KW)


def test_conditioned_4096_equals_oracle_at_full_size(gpu_out, bench_dem):
    from oracle.oracle import OracleDEMProcessor
    ref = helpers.run(lambda e, **k: OracleDEMProcessor(e, **k), bench_dem, KW)
    r = helpers.compare(ref, gpu_out)
    helpers.assert_parity(r, "conditioned4096")
    assert np.isfinite(gpu_out["uca"]).sum() > 0.99 * N * N
    assert np.nanmax(gpu_out["uca"]) > 1e5 * 900.0          # rivers: >1e5 cells drain through one outlet


def test_sweep_is_deterministic(cuda_lib, gpu_out, bench_dem):
    """A cell's sum has a fixed order (pull over its donors), so two runs agree bit for bit --
    unlike a sweep built on floating-point atomics.  (Only a cell that receives from two or more
    pits adds those few terms in arrival order; none of the cells of a run without pit drains.)"""
    from pydem_b200 import DEMProcessor
    again = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), bench_dem, KW)
    for k in ("mag", "dir"):
        np.testing.assert_array_equal(again[k], gpu_out[k])
    np.testing.assert_allclose(again["uca"], gpu_out["uca"], rtol=1e-13, equal_nan=True)
    kw = dict(KW, drain_pits=False)
    a = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), bench_dem, kw)
    b = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), bench_dem, kw)
    for k in ("uca", "twi"):
        np.testing.assert_array_equal(a[k], b[k])
    for k in ("edge_todo", "edge_done", "flats"):
        np.testing.assert_array_equal(again[k], gpu_out[k])


_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import helpers
from pydem_b200 import DEMProcessor
E = np.load(sys.argv[1])
out = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), E, dict(dX=30.0, dY=30.0, fill_flats=False, drain_pits_path=False, drain_pits=True))
np.savez(sys.argv[2], uca=out["uca"], edge_todo=out["edge_todo"], edge_done=out["edge_done"], twi=out["twi"])
"""


@pytest.mark.parametrize("env", [dict(PYDEM_B200_SWEEP_LEGACY="1", PYDEM_B200_SWEEP_STRICT="1"),
                                 dict(PYDEM_B200_TS_TILE="2")])
def test_independent_sweeps_agree(gpu_out, bench_dem, tmp_path, env):
    """The legacy cell-by-cell sweep with every decrement held back until its adds have returned
    (an independent implementation, different summation order) and the tile sweep with another
    tile shape must give the default sweep's result: masks exact, uca to re-association."""
    fn_in, fn_out = str(tmp_path / "E.npy"), str(tmp_path / "out.npz")
    np.save(fn_in, bench_dem)
    e = dict(os.environ); e.update(env)
    subprocess.run([sys.executable, "-c", _CHILD % (ROOT, os.path.join(ROOT, "tests")), fn_in, fn_out], env=e, check=True,
                   timeout=600)
    o = np.load(fn_out)
    np.testing.assert_array_equal(o["edge_todo"], gpu_out["edge_todo"])
    np.testing.assert_array_equal(o["edge_done"], gpu_out["edge_done"])
    np.testing.assert_allclose(o["uca"], gpu_out["uca"], rtol=1e-12, equal_nan=True)
    if "PYDEM_B200_TS_TILE" in env:
        np.testing.assert_allclose(o["uca"], gpu_out["uca"], rtol=1e-13, equal_nan=True)   # same pull order, any tile shape


def test_sharded_conditioned_2048_world8(cuda_lib):
    from pydem_b200 import DEMProcessor, sharded
    E = helpers.synth.conditioned_fractal_dem(2048, 3)
    dp = DEMProcessor(elev=E, dX=30.0, dY=30.0, drain_pits=False, **helpers.HOT)
    dp.calc_twi()
    out = sharded.run_local(E, 8, dX=30.0, dY=30.0)
    for k in ("mag", "direction", "flats", "edge_todo", "edge_done"):
        np.testing.assert_array_equal(out[k], getattr(dp, k))
    np.testing.assert_array_equal(out["uca"], dp.uca)                      # fixed summation order: bit-identical
    assert out["stats"][0]["sweep_rounds"] >= 3
