"""GPU suite: the row-sharded path (all shards on one GPU through LocalGroup) must reproduce
the single-tile result -- integer/bool fields bit-exact, mag/direction bit-exact (same kernel,
same neighbourhood), uca to fp64 re-association."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def single(E, **kw):
    from pydem_b200 import DEMProcessor
    dp = DEMProcessor(elev=E, drain_pits=False, **helpers.HOT, **kw)
    dp.calc_twi()
    return dp


def check(E, world, **kw):
    from pydem_b200 import sharded
    dp = single(E, **kw)
    out = sharded.run_local(E, world, dX=kw.get("dX", 1.0), dY=kw.get("dY", 1.0), dX2=kw.get("dX2"), dY2=kw.get("dY2"))
    np.testing.assert_array_equal(out["mag"], dp.mag)
    np.testing.assert_array_equal(out["direction"], dp.direction)
    np.testing.assert_array_equal(out["flats"], dp.flats)
    np.testing.assert_array_equal(out["edge_todo"], dp.edge_todo)
    np.testing.assert_array_equal(out["edge_done"], dp.edge_done)
    np.testing.assert_allclose(out["uca"], dp.uca, rtol=helpers.UCA_RTOL, equal_nan=True)
    np.testing.assert_allclose(out["twi"] * 10, dp.twi, rtol=1e-9, atol=1e-9, equal_nan=True)
    assert sum(s["n_drained"] for s in out["stats"]) == E.size
    return out


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_sharded_equals_single_tile_fractal(cuda_lib, world):
    out = check(helpers.synth.fractal_dem(0, 41, shape=(203, 160)), world, dX=30.0, dY=30.0)
    if world > 1:
        assert out["stats"][0]["sweep_rounds"] >= 2      # flow really crosses shard boundaries


def test_sharded_cone_and_variable_spacing(cuda_lib):
    check(helpers.synth.cone_dem(96) * 100 + 1, 4)
    R = 150
    check(helpers.synth.fractal_dem(0, 42, shape=(R, 90)), 3, dX=np.linspace(20, 30, R - 1), dY=np.full(R - 1, 27.3),
          dX2=np.linspace(20, 30, R), dY2=np.full(R, 27.3))


def test_sharded_flat_regions_across_boundaries(cuda_lib):
    """Lakes and quantised terraces that span several row blocks: the region labels (and with them
    the one-pixel flat extension) need the cross-rank label rounds."""
    E = helpers.synth.fractal_dem(192, 43)
    yy, xx = np.mgrid[0:192, 0:192]
    for (cy, cx, r) in ((48, 60, 20), (96, 120, 30), (140, 40, 12)):
        m = (yy - cy) ** 2 + (xx - cx) ** 2 <= r * r
        E[m] = E[m].min()
    out = check(E, 4)
    assert max(s["label_rounds"] for s in out["stats"]) >= 1
    check(np.round(helpers.synth.fractal_dem(160, 44) / 25) * 25, 5)
    # a snake-shaped flat that crosses every boundary several times
    E = helpers.synth.fractal_dem(128, 45) + 50
    for j in range(8, 120, 16):
        E[4:124, j:j + 3] = 7.0
        E[(4 if (j // 16) % 2 else 121):(7 if (j // 16) % 2 else 124), j:j + 19] = 7.0
    check(E, 4)


def test_sharded_nan_holes(cuda_lib):
    E = helpers.synth.fractal_dem(128, 46)
    E[30:34, 50:53] = np.nan; E[63:66, 10:40] = np.nan; E[0, 5] = np.nan; E[127, 100] = np.nan
    check(E, 4)


def test_sharded_2048_many_rounds(cuda_lib):
    E = helpers.synth.value_noise_dem(0, 2048, 1024, seed=3)
    out = check(E, 8, dX=30.0, dY=30.0)
    assert out["stats"][0]["sweep_rounds"] >= 2


def _n_gpus(cuda_lib):
    import ctypes as ct
    n = ct.c_int(0)
    cuda_lib.check(cuda_lib.load().pdm_device_count(ct.byref(n)))
    return n.value


@pytest.mark.parametrize("pits", [False, True])
def test_c_abi_sharded_pass_on_two_gpus(cuda_lib, pits):
    """Two processes, one per GPU, drive the sharded pass through the C ABI alone (pdm_comm_* over NCCL for the
    halo rows, one work-list sweep across the GPUs over CUDA-IPC peer memory; with pits: pit regions and drains
    across the shard boundary) and compare their rows with the single-tile result -- scripts/c_abi_shard_check.py.
    Needs a box with >= 2 GPUs; the driver's N > 1 bench runs check the same path (bench.py, "parity")."""
    if _n_gpus(cuda_lib) < 2:
        pytest.skip("needs 2 GPUs")
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "scripts", "c_abi_shard_check.py"), "2", "1024", "768"] + (["pits"] if pits else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=400)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "c_abi_shard_check OK" in r.stdout and "torch_loaded False" in r.stdout
