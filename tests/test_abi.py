"""CPU suite: the C-ABI shared library builds for sm_100a, loads, exports every symbol that
include/pydem_b200.h declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as ct
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pydem_b200 import build, _lib
    build.build()
    return _lib


def test_library_exports_header_symbols(lib):
    hdr = open(os.path.join(ROOT, "include", "pydem_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pdm_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = lib.load()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(lib.EXPORTS)
    assert L.pdm_abi_version() == 2


def test_struct_layouts_match_defaults(lib):
    L = lib.load()
    p = lib.UcaParams(); L.pdm_default_uca_params(ct.byref(p))
    assert (p.drain_pits, p.drain_pits_max_iter, p.drain_pits_max_dist, p.circular_ref_maxcount) == (1, 300, 32, 50)
    assert p.uca_saturation_limit == 32.0 and p.drain_pits_max_dist_xy == 0.0
    q = lib.TwiParams(); L.pdm_default_twi_params(ct.byref(q))
    assert q.twi_min_slope == 1e-3 and q.uca_saturation_limit == 32.0 and np.isinf(q.twi_min_area)


def test_sass_is_sm100a(lib):
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % lib.LIB_PATH).read()
    assert "sm_100a" in out, out


def _has_gpu(lib):
    n = ct.c_int(0)
    return lib.load().pdm_device_count(ct.byref(n)) == 0 and n.value > 0


def test_no_cpu_fallback(lib):
    """Without a device the operator raises; nothing is computed on the host."""
    if _has_gpu(lib):
        pytest.skip("a GPU is present")
    from pydem_b200 import DEMProcessor
    dp = DEMProcessor(elev=np.ones((8, 8)), fill_flats=False, drain_pits_path=False)
    with pytest.raises(RuntimeError):
        dp.calc_slopes_directions()
    with pytest.raises(RuntimeError):
        dp.calc_twi()
    assert dp.mag is None and dp.uca is None


def test_operator_surface_matches_reference():
    """Same method names / argument names as pydem.dem_processing.DEMProcessor
    (reference dem_processing.py:204, 305, 587, 682, 1647)."""
    from pydem_b200 import DEMProcessor
    assert list(inspect.signature(DEMProcessor.__init__).parameters)[:2] == ["self", "elev_fn"]
    assert list(inspect.signature(DEMProcessor.calc_slopes_directions).parameters) == ["self", "plotflag"]
    assert list(inspect.signature(DEMProcessor.calc_uca).parameters) == ["self", "plotflag", "edge_init_data", "uca_init"]
    assert list(inspect.signature(DEMProcessor.calc_twi).parameters) == ["self"]
    dp = DEMProcessor(elev=np.zeros((5, 7)), dX=2.0, dY=3.0)
    assert dp.dX.shape == (4,) and dp.dX2.shape == (5,) and dp.dY[0] == 3.0 and dp.dX2[0] == 2.0
    for flag, default in (("fill_flats", True), ("drain_pits", True), ("drain_pits_path", True),
                          ("drain_pits_max_iter", 300), ("drain_pits_max_dist", 32), ("uca_saturation_limit", 32),
                          ("twi_min_slope", 1e-3), ("circular_ref_maxcount", 50), ("apply_twi_limits", False)):
        assert getattr(dp, flag) == default
    with pytest.raises(TypeError):
        DEMProcessor(elev=np.zeros((5, 5)), not_a_flag=1)
    for meth in ("calc_fill_flats", "calc_fill_pit_artifacts", "calc_pit_drain_paths", "find_flats"):
        assert list(inspect.signature(getattr(DEMProcessor, meth)).parameters) == ["self"]
    with pytest.raises(RuntimeError):                                   # no device here, and no CPU fallback
        DEMProcessor(elev=np.ones((5, 5))).calc_slopes_directions()      # default flags: conditioning runs first


def test_header_is_plain_c_and_the_c_host_links(lib, tmp_path):
    """include/pydem_b200.h must be usable from C (the boundary is a C ABI): examples/shard_host.c -- the sharded pass
    driven from plain C, INTEGRATION.md section 5 -- compiles as C99 with -Wall -Wextra and links against the library."""
    import os, shutil, subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "shard_host")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"),
                        os.path.join(root, "examples", "shard_host.c"), "-L", os.path.join(root, "pydem_b200"),
                        "-lpydem_b200", "-lm", "-o", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert os.path.exists(out)
