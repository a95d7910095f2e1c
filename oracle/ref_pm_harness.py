"""Run the UNMODIFIED reference ``pydem/process_manager.py`` in this container, in memory.

Test infrastructure (see ``oracle/__init__.py``); extends ``ref_harness``.  The reference's tile
orchestrator talks to two packages that are not installed here -- ``zarr`` (its result store) and
``rasterio`` (its GeoTIFF reader).  Neither does arithmetic, so both are replaced by in-memory
stand-ins registered in ``sys.modules`` before ``pydem.process_manager`` is imported from where it
lies:

* ``zarr.open(path, ...)`` -> arrays/groups of a process-wide dict keyed by path; like zarr,
  indexing an array returns a COPY and only ``arr[key] = value`` writes to the store;
* ``rasterio.open(fn)`` -> a registered in-memory raster with ``bounds``, ``transform``, ``shape``,
  ``crs`` and ``read(1)``, laid out exactly as ``utils.mk_geotiff_obj`` (utils.py:178-206) would
  have written it (pixel-centred lat/lon, WGS84 extents), so that ``ProcessManager.compute_grid``
  / ``compute_grid_overlaps`` see the same geometry as in the reference's own multi-file tests
  (``utils_test_pydem.mk_test_multifile`` :359-408, ``test_end_to_end.py`` :86-149).

``/root/reference`` does not exist on the GPU box; the outputs of this harness travel as
``tests/golden/ref_pm.npz`` (script: ``tests/golden/make_golden_pm.py``).
"""
import collections
import importlib
import os
import sys
import types
import warnings

import numpy as np

from . import ref_harness

_STORE = {}      # path -> ndarray          (the "zarr" store)
_CHUNKS = {}     # path -> chunk shape given when the array was created
_RASTERS = {}    # file name -> _Raster     (the "GeoTIFF" files)


class _MemArray(object):
    def __init__(self, a, chunks=None):
        self._a = a
        self.chunks = list(chunks) if chunks is not None else list(a.shape)

    shape = property(lambda self: list(self._a.shape))
    dtype = property(lambda self: self._a.dtype)

    def __getitem__(self, key):
        v = self._a[key]
        return np.array(v) if isinstance(v, np.ndarray) else v

    def __setitem__(self, key, value):
        self._a[key] = value

    def __array__(self, dtype=None, copy=None):
        return np.array(self._a, dtype=dtype)


class _MemGroup(object):
    def __init__(self, path):
        self._p = path

    def __getitem__(self, name):
        p = os.path.normpath(os.path.join(self._p, name))
        return _MemArray(_STORE[p], _CHUNKS.get(p))

    def __contains__(self, name):
        return os.path.normpath(os.path.join(self._p, name)) in _STORE


def _zarr_open(path, mode="a", shape=None, chunks=None, dtype=None, fill_value=None, **kw):
    p = os.path.normpath(path)
    if shape is not None and p not in _STORE:
        _STORE[p] = np.full(tuple(int(s) for s in shape), 0 if fill_value is None else fill_value, dtype=dtype or "float64")
        _CHUNKS[p] = [int(c) for c in chunks] if chunks is not None else None
    if p in _STORE:
        return _MemArray(_STORE[p], _CHUNKS.get(p))
    return _MemGroup(p)


_Bounds = collections.namedtuple("BoundingBox", "left bottom right top")
_Affine = collections.namedtuple("Affine", "a b c d e f")


class _Crs(object):
    is_projected = True      # -> dX = transform.a, dY = |transform.e| (utils.py:132-137)


class _Raster(object):
    def __init__(self, data, lat, lon):
        # utils.mk_geotiff_obj: lat = [north, south], lon = [west, east] are PIXEL CENTRES
        ni, nj = data.shape
        ph = -abs(lat[0] - lat[1]) / (ni - 1.0)
        pw = abs(lon[0] - lon[1]) / (nj - 1.0)
        top = max(lat) - ph / 2
        left = min(lon) - pw / 2
        self._data = np.array(data, dtype="float64")
        self.transform = _Affine(pw, 0.0, left, 0.0, ph, top)
        self.bounds = _Bounds(left, top + ni * ph, left + nj * pw, top)
        self.shape = data.shape
        self.crs = _Crs()

    def read(self, band=1):
        return np.array(self._data)


def _rasterio_open(fn, mode="r", **kw):
    return _RASTERS[fn]


def _install():
    ref_harness._install_shims()
    z = sys.modules["zarr"]
    z.open = _zarr_open
    r = sys.modules["rasterio"]
    r.open = _rasterio_open


_PM = None


def load_process_manager():
    """The reference's ``pydem.process_manager`` module, imported in place."""
    global _PM
    if _PM is None:
        ref_harness.load_reference()
        _install()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _PM = importlib.import_module("pydem.process_manager")
    return _PM


def chunk_edges(nn, n_grid, overlap):
    """utils_test_pydem.mk_test_multifile._get_chunk_edges (:372-380)."""
    size = int(np.ceil(nn / n_grid))
    lo = np.arange(0, nn - overlap, size)
    lo[1:] -= overlap // 2
    hi = np.arange(0, nn - overlap, size)
    hi[:-1] = hi[1:] + int(np.ceil(overlap / 2))
    hi[-1] = nn
    return lo, np.minimum(hi, nn)


def register_tiles(E, nx_grid, ny_grid, overlap, tag, lat=(46.0, 45.0), lon=(-73.0, -72.0)):
    """Cut ``E`` like mk_test_multifile and register the pieces as in-memory rasters.
    Returns (file names, [(te, be, le, re)])."""
    ni, nj = E.shape
    te_, be_ = chunk_edges(ni, ny_grid, overlap)
    le_, re_ = chunk_edges(nj, nx_grid, overlap)
    la = np.linspace(lat[0], lat[1], ni)
    lo = np.linspace(lon[0], lon[1], nj)
    names, boxes = [], []
    for te, be in zip(te_, be_):
        for le, re in zip(le_, re_):
            fn = "mem://%s/chunks/tile_%04d_%04d.tif" % (tag, te, le)
            _RASTERS[fn] = _Raster(E[te:be, le:re], [la[te], la[be - 1]], [lo[le], lo[re - 1]])
            names.append(fn); boxes.append((int(te), int(be), int(le), int(re)))
    order = np.argsort(names)      # ProcessManager sorts its source files (process_manager.py:470)
    return [names[k] for k in order], [boxes[k] for k in order]


def run_reference_pm(E, nx_grid, ny_grid, overlap, tag, dem_processor=None, dem_proc_kwargs=None, debug_spacing=True,
                     lat=(46.0, 45.0), lon=(-73.0, -72.0), overviews=None, drop=()):
    """ProcessManager.process_twi() + save_non_overlap_data() of the reference on in-memory tiles.
    dem_processor: class to put in place of ``pydem.process_manager.DEMProcessor`` (None = the
    reference's own).  debug_spacing=False keeps the spacing the reference derives from the rasters
    (projected CRS: dX = pixel width, dY = pixel height, utils.py:132-137).  Returns a dict of the global arrays and the compact (non-overlapping) ones."""
    pm_mod = load_process_manager()
    names, boxes = register_tiles(E, nx_grid, ny_grid, overlap, tag, lat=lat, lon=lon)
    if drop:                                  # a mosaic with holes: the reference allows missing tiles (grid_id2i == -1)
        names = [n for k, n in enumerate(names) if k not in drop]; boxes = [b for k, b in enumerate(boxes) if k not in drop]
    out_path = "mem://%s/results.zarr" % tag
    for k in [k for k in _STORE if k.startswith(os.path.normpath("mem://%s" % tag))]:
        del _STORE[k]
    old_dp, old_dbg, old_ec = pm_mod.DEMProcessor, pm_mod.DEBUG, pm_mod.calc_uca_ec
    if dem_processor is not None:
        pm_mod.DEMProcessor = dem_processor
    pm_mod.DEBUG = bool(debug_spacing)        # dX = dY = dX2 = dY2 = 1, as in test_end_to_end.py:52
    order = []

    def logged_ec(**kw):                      # records which tile process_uca_edges corrects, in order
        order.append(names.index(kw["fn"]))
        return old_ec(**kw)
    pm_mod.calc_uca_ec = logged_ec
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pm = pm_mod.ProcessManager(in_path="mem://%s/chunks" % tag, out_path=out_path, elev_source_files=list(names),
                                       n_workers=1, dem_proc_kwargs=dict(dem_proc_kwargs or {}))
            pm.process_twi()
            if not drop:                      # the reference's compact store cannot be laid out around a hole
                pm.save_non_overlap_data()
            if overviews:
                pm.process_overviews(pm.out_path_noverlap, keys=["elev", "uca"], overviews=list(overviews))
    finally:
        pm_mod.DEMProcessor, pm_mod.DEBUG, pm_mod.calc_uca_ec = old_dp, old_dbg, old_ec
    res = {"correction_order": order}
    base = os.path.normpath(out_path)
    for key in ("elev", "aspect", "slope", "uca", "uca_edges", "edge_todo", "edge_done", "twi", "success"):
        res[key] = np.array(_STORE[os.path.normpath(os.path.join(base, key))])
    compact = os.path.normpath(pm.out_path_noverlap) if not drop else "\0none"
    for key in ("elev", "uca", "aspect", "slope", "twi"):
        if not drop:
            res["compact_" + key] = np.array(_STORE[os.path.normpath(os.path.join(compact, key))])
    for k in _STORE:
        if k.startswith(compact + os.sep) and os.path.basename(k).split("_")[-1].isdigit():
            res["overview_" + os.path.basename(k)] = np.array(_STORE[k])
    res["compact_chunks"] = _CHUNKS.get(os.path.normpath(os.path.join(compact, "uca")))
    res["grid_slice"] = [(s[0].start, s[0].stop, s[1].start, s[1].stop) for s in pm.grid_slice]
    res["grid_slice_unique"] = [(s[0].start, s[0].stop, s[1].start, s[1].stop) for s in pm.grid_slice_unique]
    res["edge_data"] = pm.edge_data
    res["boxes"] = boxes
    return res
