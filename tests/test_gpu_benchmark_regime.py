"""GPU suite: parity at the benchmark's own regime (BASELINE.json configs[1], primary variant): the
4096 x 4096 priority-flood conditioned fractal with the default drain_pits=True -- 3.9 k dependency
levels, rivers thousands of cells long -- compared with the oracle at FULL size (the oracle needs
~10 s for it); the default work-list sweep against its strictly ordered mode and against the tile
sweep (an independent, bit-reproducible second implementation); and the row-sharded path on a
conditioned DEM."""
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


@pytest.fixture()
def tile_sweep(cuda_lib):
    """Run the accumulation with the tile sweep (the engine of the sharded path) instead of the default work-list."""
    import pydem_b200
    pydem_b200.set_sweep("tile")
    yield
    pydem_b200.set_sweep("env")


def test_tile_sweep_is_deterministic_and_equals_worklist(cuda_lib, gpu_out, bench_dem, tile_sweep):
    """The tile sweep sums a cell's donors in a fixed order (pull), so two runs agree bit for bit --
    unlike the work-list's floating-point atomics -- and it is an independent second implementation of
    the default engine: masks exact, uca to re-association.  (Only a cell that receives from two or
    more pits adds those few terms in arrival order; none of the cells of a run without pit drains.)"""
    from pydem_b200 import DEMProcessor, get_sweep
    assert get_sweep() == "tile"
    t1 = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), bench_dem, KW)
    for k in ("edge_todo", "edge_done", "flats"):
        np.testing.assert_array_equal(t1[k], gpu_out[k])
    np.testing.assert_allclose(t1["uca"], gpu_out["uca"], rtol=1e-12, equal_nan=True)
    kw = dict(KW, drain_pits=False)
    a = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), bench_dem, kw)
    b = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), bench_dem, kw)
    for k in ("uca", "twi", "edge_done"):
        np.testing.assert_array_equal(a[k], b[k])


def test_strict_worklist_equals_default(cuda_lib, gpu_out, bench_dem):
    """The default work-list does not wait for a receiver's adds before counting it off (same-sector
    program order at L2, csrc/drain_op.cuh).  With strict=True every count-off is held back until the
    adds have returned: must give the same masks and the same uca up to re-association."""
    import pydem_b200
    from pydem_b200 import DEMProcessor
    pydem_b200.set_sweep("worklist", strict=True)
    try:
        o = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), bench_dem, KW)
    finally:
        pydem_b200.set_sweep("env")
    np.testing.assert_array_equal(o["edge_todo"], gpu_out["edge_todo"])
    np.testing.assert_array_equal(o["edge_done"], gpu_out["edge_done"])
    np.testing.assert_allclose(o["uca"], gpu_out["uca"], rtol=1e-12, equal_nan=True)


_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import helpers
from pydem_b200 import DEMProcessor
E = np.load(sys.argv[1])
out = helpers.run(lambda e, **k: DEMProcessor(elev=e, **k), E, dict(dX=30.0, dY=30.0, fill_flats=False, drain_pits_path=False, drain_pits=True))
np.savez(sys.argv[2], uca=out["uca"], edge_todo=out["edge_todo"], edge_done=out["edge_done"], twi=out["twi"])
"""


def test_tile_shapes_agree_bit_for_bit(cuda_lib, bench_dem, tmp_path):
    """Same pull order whatever the tile shape: two tile sweeps with different tiles (the shape is read
    from the environment once per process) give identical bits."""
    fn_in = str(tmp_path / "E.npy")
    np.save(fn_in, bench_dem)
    outs = []
    for shape in ("0", "2"):
        fn_out = str(tmp_path / ("out%s.npz" % shape))
        e = dict(os.environ, PYDEM_B200_SWEEP="tile", PYDEM_B200_TS_TILE=shape)
        subprocess.run([sys.executable, "-c", _CHILD % (ROOT, os.path.join(ROOT, "tests")), fn_in, fn_out], env=e, check=True,
                       timeout=600)
        outs.append(np.load(fn_out))
    for k in ("uca", "edge_todo", "edge_done", "twi"):
        np.testing.assert_array_equal(outs[0][k], outs[1][k])


def test_sharded_conditioned_2048_world8(cuda_lib, tile_sweep):
    from pydem_b200 import DEMProcessor, sharded
    E = helpers.synth.conditioned_fractal_dem(2048, 3)
    dp = DEMProcessor(elev=E, dX=30.0, dY=30.0, drain_pits=False, **helpers.HOT)
    dp.calc_twi()
    out = sharded.run_local(E, 8, dX=30.0, dY=30.0)
    for k in ("mag", "direction", "flats", "edge_todo", "edge_done"):
        np.testing.assert_array_equal(out[k], getattr(dp, k))
    np.testing.assert_array_equal(out["uca"], dp.uca)                      # fixed summation order: bit-identical
    assert out["stats"][0]["sweep_rounds"] >= 3
