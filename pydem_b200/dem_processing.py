"""Drop-in ``DEMProcessor`` for pyDEM's hot path, backed by ``libpydem_b200.so``.

Mirrors the operator surface of the reference ``pydem.dem_processing.DEMProcessor``
(reference dem_processing.py:98-258): same constructor keywords, same attributes
(``elev dX dY dX2 dY2 direction mag flats uca twi edge_todo edge_done twi_min_area`` and the
behaviour flags of lines 105-154), same methods

    calc_slopes_directions() -> (mag, direction)      reference :587
    find_flats()                                      reference :305
    calc_uca(plotflag=False, edge_init_data=None, uca_init=None) -> uca   reference :682
    calc_twi() -> twi (un-scaled; self.twi = 10*twi)  reference :1647

including the reference's in-place side effects (``mag``/``flats`` updated at drained pits,
``twi_min_area`` lowered by ``calc_uca``).  Arrays are NumPy on the host side, exactly like
the reference; all arithmetic happens on the GPU.  There is no CPU fallback: a missing
library or device raises ``RuntimeError`` (the reference does the same when its Cython
module is absent, dem_processing.py:714-715).

Host<->device traffic: results stay in HBM and the array attributes (``mag direction flats uca
edge_todo edge_done twi`` and a conditioned ``elev``) are *lazy*: the first read downloads the
field (the object keeps its device tile alive until then), only the value a ``calc_*`` call
returns is copied eagerly.  ``calc_twi()`` on a fresh object therefore moves the DEM in and the
index out -- not the six intermediate fields nobody asked for.  Callers of the reference mutate
the NumPy attributes between calls (``dp.dX[:] = 1`` in its tests, edge rows of ``direction`` in
``process_manager.calc_uca``): once an attribute has been read (or assigned) the host array is
the truth again and the next ``calc_*`` call uploads it, exactly as before.
``PYDEM_B200_EAGER=1`` restores the eager behaviour (every output downloaded by the call that
produced it).
"""
import ctypes as ct
import os
import warnings

import numpy as np

from . import _lib, _pinned

FLAT_ID = np.nan
FLAT_ID_INT = -1


class DEMProcessor(object):
    # behaviour flags and their reference defaults (dem_processing.py:105-154)
    _DEFAULTS = dict(
        fill_flats=True, fill_flats_below_sea=False, fill_flats_source_tol=1, fill_flats_peaks=True,
        fill_flats_pits=True, fill_flats_max_iter=10,
        drain_pits=True, drain_pits_path=True, drain_pits_min_border=False, drain_pits_spill=False,
        drain_flats=False, drain_pits_max_iter=300, drain_pits_max_dist=32, drain_pits_max_dist_XY=None,
        apply_uca_limit_edges=False, apply_twi_limits=False, apply_twi_limits_on_uca=False,
        uca_saturation_limit=32.0, twi_min_slope=1e-3, twi_min_area=np.inf, circular_ref_maxcount=50,
        maximum_pit_area=32.0, plotflag=False,
    )
    _ARRAYS = ("elev", "direction", "mag", "uca", "twi", "flats", "done", "dX", "dY", "dX2", "dY2")

    def __init__(self, elev_fn=None, **kwargs):
        self._tile = None
        if elev_fn:
            kwds = _raster_kwargs(elev_fn)
            kwds.update(kwargs)
            kwargs = kwds
        if kwargs.get("elev") is None:
            raise ValueError("DEMProcessor needs `elev` (or an elevation file name)")
        nrow = np.shape(kwargs["elev"])[0]
        # scalar spacing -> per-row arrays (dem_processing.py:233-240)
        if not isinstance(kwargs.get("dX"), np.ndarray):
            if "dX2" not in kwargs:
                kwargs["dX2"] = np.ones(nrow) * kwargs.get("dX", 1)
            kwargs["dX"] = np.ones(nrow - 1) * kwargs.get("dX", 1)
        if not isinstance(kwargs.get("dY"), np.ndarray):
            if "dY2" not in kwargs:
                kwargs["dY2"] = np.ones(nrow) * kwargs.get("dY", 1)
            kwargs["dY"] = np.ones(nrow - 1) * kwargs.get("dY", 1)
        for k, v in self._DEFAULTS.items():
            setattr(self, k, kwargs.pop(k, v))
        self._host = {}          # lazy attributes: host copies (None = not on the host)
        self._ondev = set()      # lazy attributes whose only valid copy lives on the device tile
        self._eager = bool(int(os.environ.get("PYDEM_B200_EAGER", "0")))
        for k in self._ARRAYS:
            v = kwargs.pop(k, None)
            setattr(self, k, None if v is None else np.asarray(v))
        if self.dX2 is None:
            self.dX2 = np.ones(nrow)                                    # dem_processing.py:252-258
        if self.dY2 is None:
            self.dY2 = np.ones(nrow)
        self.bounds = kwargs.pop("bounds", [])
        self.transform = kwargs.pop("transform", [])
        self.device = kwargs.pop("device", None)
        if kwargs:
            raise TypeError("DEMProcessor got unexpected keyword arguments %s" % sorted(kwargs))
        self.section = None
        self.proportion = None
        self.A = None            # the reference's sparse matrix is never materialised here
        self.uca_stats = None    # pdm_uca_stats of the last calc_uca
        self._tile = None
        self._tile_shape = None
        self._chain = 0          # >0 while stages are chained inside one public call
        self._resident = set()   # fields valid in HBM during a chain

    # ------------------------------------------------------------------------------------
    # lazy array attributes
    # ------------------------------------------------------------------------------------
    _LAZY = {"elev": _lib.F_ELEV, "mag": _lib.F_MAG, "direction": _lib.F_DIR, "flats": _lib.F_FLATS, "uca": _lib.F_UCA,
             "edge_todo": _lib.F_EDGE_TODO, "edge_done": _lib.F_EDGE_DONE}
    _BOOL = ("flats", "edge_todo", "edge_done")

    def _lazy_get(self, name):
        v = self._host.get(name)
        if v is None and name in self._ondev:
            v = self._fetch(name)
        return v

    def _lazy_set(self, name, value):
        self._host[name] = None if value is None else np.asarray(value)
        self._ondev.discard(name)
        if name == "elev" and value is not None:
            self._shape = tuple(np.shape(value))

    def _has(self, name):
        return self._host.get(name) is not None or name in self._ondev

    def _fetch(self, name):
        """first read of a device-only attribute: download it.  From here on the host array is the
        truth (the caller may mutate it), so the device copy is no longer trusted."""
        out = _pinned.empty(self._tile_shape, _lib.FIELD_DTYPE[self._LAZY[name]])
        _lib.check(_lib.load().pdm_tile_download(self._tile, self._LAZY[name], _lib.ptr(out)))
        if name in self._BOOL:
            out = out.view(np.bool_)
        self._host[name] = out
        self._ondev.discard(name)
        return out

    def _produced(self, *names):
        """fields a stage has just written on the device"""
        for n in names:
            self._host[n] = None
            self._ondev.add(n)
        if self._eager:
            for n in names:
                self._fetch(n)

    @property
    def twi(self):
        """10 * (un-scaled index), dem_processing.py:1674"""
        if self._twi10 is None and self._twi_raw is not None:
            self._twi10 = self._twi_raw * 10.0        # same IEEE product the device would form
        return self._twi10

    @twi.setter
    def twi(self, value):
        self._twi10 = None if value is None else np.asarray(value)
        self._twi_raw = None

    # ------------------------------------------------------------------------------------
    # device plumbing
    # ------------------------------------------------------------------------------------
    def _lib(self):
        L = _lib.load()
        _lib.init(self.device)
        return L

    # one parked device tile per shape: ProcessManager-style callers build a fresh DEMProcessor
    # per tile and stage (process_manager.py:54-315), all of the same shape
    _TILE_CACHE = {}

    def _get_tile(self):
        L = self._lib()
        shape = tuple(self._shape)
        if self._tile is not None and self._tile_shape != shape:
            self._free_tile()
        if self._tile is None:
            if len(shape) != 2:
                raise ValueError("elev must be 2-D")
            key = (shape, _lib._device)
            h = DEMProcessor._TILE_CACHE.pop(key, None)
            if h is None:
                h = ct.c_void_p()
                _lib.check(L.pdm_tile_create(shape[0], shape[1], None, ct.byref(h)))
            self._tile, self._tile_shape = h, shape
        return self._tile

    def _free_tile(self, keep_results=False):
        """Park the device tile for the next DEMProcessor of this shape.  keep_results: download the
        attributes that only live on the device first (otherwise they are dropped)."""
        if self._tile is not None:
            if keep_results:
                for n in list(self._ondev):
                    self._fetch(n)
            self._ondev = set()
            try:
                key = (self._tile_shape, _lib._device)
                if key not in DEMProcessor._TILE_CACHE:
                    DEMProcessor._TILE_CACHE[key] = self._tile
                else:
                    _lib.load().pdm_tile_destroy(self._tile)
            except Exception:
                pass
            self._tile = None
            self._resident = set()

    @classmethod
    def release_device_memory(cls):
        """Free the parked device tiles."""
        for h in cls._TILE_CACHE.values():
            try:
                _lib.load().pdm_tile_destroy(h)
            except Exception:
                pass
        cls._TILE_CACHE.clear()

    def __del__(self):
        self._free_tile()

    def _up(self, field, name):
        """host -> HBM unless the field is already resident from a chained stage or has never
        left the device."""
        if name in self._ondev or (self._chain and field in self._resident):
            return
        a = np.asarray(self._host[name])
        if a.dtype == np.bool_ and _lib.FIELD_DTYPE[field] == np.uint8:
            a = a.view(np.uint8)                      # same bytes, no copy
        a = np.ascontiguousarray(a, dtype=_lib.FIELD_DTYPE[field])
        if a.shape != self._tile_shape:
            raise ValueError("field %d has shape %s, expected %s" % (field, a.shape, self._tile_shape))
        _lib.check(_lib.load().pdm_tile_upload(self._get_tile(), field, _lib.ptr(a)))
        self._resident.add(field)

    def _down(self, field):
        """HBM -> page-locked host array.  The copy is queued behind the stage that produced the
        field and overlaps the following stages; `_end()` of the outermost public call waits."""
        out = _pinned.empty(self._tile_shape, _lib.FIELD_DTYPE[field])
        _lib.check(_lib.load().pdm_tile_download_async(self._get_tile(), field, _lib.ptr(out)))
        return out

    @staticmethod
    def _patched(arr, cells, values, dtype):
        """arr.flat[cells] = values, in place when arr is a writable C-contiguous array of the right
        dtype (the reference mutates the caller's array), else on a converted copy."""
        a = arr
        if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.writeable and a.flags.c_contiguous):
            a = np.array(arr, dtype=dtype, order="C")
        a.reshape(-1)[cells] = values
        return a

    def _spacing(self):
        R = self._shape[0]
        dX = np.ascontiguousarray(self.dX, "float64"); dY = np.ascontiguousarray(self.dY, "float64")
        dX2 = np.ascontiguousarray(self.dX2, "float64"); dY2 = np.ascontiguousarray(self.dY2, "float64")
        if dX.shape != (R - 1,) or dY.shape != (R - 1,) or dX2.shape != (R,) or dY2.shape != (R,):
            raise ValueError("dX, dY need %d entries and dX2, dY2 %d" % (R - 1, R))
        # facet angles with NumPy's arctan2, as the reference computes them (dem_processing.py:1936)
        thA = np.ascontiguousarray(np.arctan2(dY, dX)); thB = np.ascontiguousarray(np.arctan2(dX, dY))
        _lib.check(_lib.load().pdm_tile_set_spacing(self._get_tile(), _lib.ptr(dX), _lib.ptr(dY), _lib.ptr(dX2),
                                                    _lib.ptr(dY2), _lib.ptr(thA), _lib.ptr(thB)))

    def _begin(self):
        if self._chain == 0:
            self._resident = set()
        self._chain += 1

    def _end(self):
        self._chain -= 1
        if self._chain == 0:
            self._resident = set()
            if self._tile is not None:
                _lib.check(_lib.load().pdm_tile_sync(self._tile))      # all queued downloads have landed

    # ------------------------------------------------------------------------------------
    # reference API
    # ------------------------------------------------------------------------------------
    def find_flats(self):
        """dem_processing.py:305-306"""
        self.flats = self.mag == FLAT_ID_INT

    def _cond_params(self):
        p = _lib.CondParams()
        _lib.load().pdm_default_cond_params(ct.byref(p))
        p.fill_flats_below_sea = int(bool(self.fill_flats_below_sea))
        p.fill_flats_source_tol = int(self.fill_flats_source_tol)
        p.fill_flats_peaks = int(bool(self.fill_flats_peaks))
        p.fill_flats_pits = int(bool(self.fill_flats_pits))
        p.maximum_pit_area = float(self.maximum_pit_area or 0.0)
        p.drain_pits_max_iter = int(self.drain_pits_max_iter)
        p.drain_pits_max_dist = int(self.drain_pits_max_dist or 0)
        p.drain_pits_max_dist_xy = float(self.drain_pits_max_dist_XY or 0.0)
        return p

    def _condition(self, stages):
        """Run conditioning stages on the device copy of ``elev`` (uploaded once) and bring the
        conditioned elevation back: the reference rebinds ``self.elev`` (425, 548, 580)."""
        self._begin()
        try:
            L = self._lib()
            t = self._get_tile()
            self._spacing()
            self._up(_lib.F_ELEV, "elev")
            p = self._cond_params()
            stats = {}
            for name in stages:
                st = _lib.CondStats()
                _lib.check(getattr(L, "pdm_tile_" + name)(t, ct.byref(p), ct.byref(st)))
                stats.update({k: getattr(st, k) for k, _ in st._fields_ if getattr(st, k)})
            self._resident = {_lib.F_ELEV}            # every derived field of the tile is stale now
            self._produced("elev")                    # the reference rebinds self.elev
            self.cond_stats = stats
            if stats.get("n_pits_undrained"):
                warnings.warn("Warning %d pits had no place to drain to in this chunk" % stats["n_pits_undrained"])
        finally:
            self._end()
        return self.elev

    def calc_fill_pit_artifacts(self):
        """dem_processing.py:396-426"""
        self._condition(["fill_pit_artifacts"])

    def calc_fill_flats(self):
        """dem_processing.py:551-585 (includes calc_fill_pit_artifacts when maximum_pit_area is set)"""
        self._condition(["fill_flats"])

    def calc_pit_drain_paths(self):
        """dem_processing.py:428-548"""
        return self._condition(["pit_drain_paths"])

    def calc_slopes_directions(self, plotflag=False):
        """Magnitude and direction of slopes -> self.mag, self.direction, self.flats
        (dem_processing.py:587-619), after the elevation conditioning the flags ask for (601-609)."""
        self._begin()
        try:
            stages = (["fill_flats"] if self.fill_flats else []) + (["pit_drain_paths"] if self.drain_pits_path else [])
            if stages:
                self._condition(stages)
        except Exception:
            self._end()
            raise
        try:
            L = self._lib()
            t = self._get_tile()
            self._spacing()
            if ("elev" in self._ondev or (self._chain and _lib.F_ELEV in self._resident)
                    or os.environ.get("PYDEM_B200_CHUNKED_UPLOAD", "1") == "0"):
                self._up(_lib.F_ELEV, "elev")
                _lib.check(L.pdm_tile_slopes_directions(t))
            else:
                # elevation still on the host: upload it in row chunks and run the stencil of a chunk while the next one
                # is on its way (same results as upload + stencil)
                a = np.ascontiguousarray(np.asarray(self._host["elev"]), dtype=np.float64)
                if a.shape != self._tile_shape:
                    raise ValueError("elev has shape %s, expected %s" % (a.shape, self._tile_shape))
                self._upload_keepalive = a
                _lib.check(L.pdm_tile_upload_slopes_directions(t, _lib.ptr(a)))
                self._resident.add(_lib.F_ELEV)
            self._resident.update((_lib.F_MAG, _lib.F_DIR, _lib.F_FLATS))
            self._produced("mag", "direction", "flats")
            outer = self._chain == 1
        finally:
            self._end()
        if outer:
            return self.mag, self.direction       # the call's return value: downloaded now
        return None

    def _uca_params(self):
        p = _lib.UcaParams()
        _lib.load().pdm_default_uca_params(ct.byref(p))
        if self.drain_flats and not self.drain_pits:
            raise NotImplementedError("drain_flats is deprecated in the reference (README) and not accelerated")
        if self.drain_pits_spill and not self.drain_pits:
            raise NotImplementedError("drain_pits_spill ('not a great option', dem_processing.py:115) is not accelerated")
        p.drain_pits = int(bool(self.drain_pits))
        p.drain_pits_min_border = int(bool(self.drain_pits_min_border))
        p.drain_pits_max_iter = int(self.drain_pits_max_iter)
        p.drain_pits_max_dist = int(self.drain_pits_max_dist or 0)
        p.drain_pits_max_dist_xy = float(self.drain_pits_max_dist_XY or 0.0)
        p.apply_uca_limit_edges = int(bool(self.apply_uca_limit_edges))
        p.circular_ref_maxcount = int(self.circular_ref_maxcount)
        p.uca_saturation_limit = float(self.uca_saturation_limit)
        return p

    def calc_uca(self, plotflag=False, edge_init_data=None, uca_init=None):
        """Upstream contributing area (dem_processing.py:682-776).

        Full mode (``uca_init is None``) computes the tile's own UCA plus ``edge_todo`` /
        ``edge_done``.  Update mode propagates only the difference between the neighbours'
        finished edge values (``edge_init_data = [data, done, todo]``, dicts keyed
        left/right/top/bottom) and ``uca_init``."""
        self._begin()
        try:
            if not self._has("direction"):
                self.calc_slopes_directions()
            L = self._lib()
            t = self._get_tile()
            self._spacing()
            self._up(_lib.F_ELEV, "elev")
            self._up(_lib.F_DIR, "direction")
            self._up(_lib.F_MAG, "mag")
            if not self._has("flats"):
                raise ValueError("flats is not set: call calc_slopes_directions() or find_flats() first")
            self._up(_lib.F_FLATS, "flats")
            p = self._uca_params()
            st = _lib.UcaStats()
            if uca_init is None:
                _lib.check(L.pdm_tile_uca(t, ct.byref(p), ct.byref(st)))
                # dem_processing.py:898-899
                self.twi_min_area = min(self.twi_min_area, st.min_area)
            else:
                R, C = self._tile_shape
                strips = _pack_edges(edge_init_data, R, C)
                self._resident.discard(_lib.F_UCA)
                self.uca = np.asarray(uca_init).astype("float64")                   # 744
                self._up(_lib.F_UCA, "uca")
                args = [_lib.ptr(a) for a in strips]
                _lib.check(L.pdm_tile_uca_update(t, ct.byref(p), *args, ct.byref(st)))
            self._resident.update((_lib.F_UCA, _lib.F_MAG, _lib.F_FLATS))
            self.uca_stats = st.as_dict()
            self._produced("uca", "edge_todo", "edge_done")
            if p.drain_pits and st.n_pits:
                # _mk_connectivity_pits updates mag / flats of drained pits in place (1370-1371).  It
                # changes them only at the examined pits: host arrays the caller already holds are
                # patched in place (like the reference) from a sparse readback; copies that only live
                # on the device are already up to date
                host_mag, host_fl = self._host.get("mag"), self._host.get("flats")
                if host_mag is not None or host_fl is not None:
                    _lib.check(L.pdm_tile_sync(t))
                    npit = int(st.n_pits)
                    cells = np.empty(npit, np.int32); pmag = np.empty(npit, np.float64); pfl = np.empty(npit, np.uint8)
                    got = ct.c_int64(0)
                    _lib.check(L.pdm_tile_pit_updates(t, npit, _lib.ptr(cells), _lib.ptr(pmag), _lib.ptr(pfl), ct.byref(got)))
                    if host_mag is not None:
                        self._host["mag"] = self._patched(host_mag, cells, pmag, np.float64)
                    if host_fl is not None:
                        self._host["flats"] = self._patched(host_fl, cells, pfl.view(np.bool_), np.bool_)
                if st.n_pits_undrained:
                    warnings.warn("Warning %d pits had no place to drain to in this chunk" % st.n_pits_undrained)
            if st.n_undone:
                # only reachable with circular_ref_maxcount <= 1 (the reference's sweep loop, dem_processing.py:951-952,
                # then never runs); otherwise the library has raised
                warnings.warn("circular_ref_maxcount=%d: the accumulation sweep did not run, %d cells keep their own area"
                              % (int(self.circular_ref_maxcount), st.n_undone))
            outer = self._chain == 1
        finally:
            self._end()
        return self.uca if outer else None

    def calc_twi(self):
        """Topographic wetness index (dem_processing.py:1647-1677): returns the un-scaled
        index and stores ``self.twi = 10 * twi``."""
        self._begin()
        try:
            if not self._has("uca"):
                self.calc_uca()
            L = self._lib()
            t = self._get_tile()
            self._up(_lib.F_UCA, "uca")
            self._up(_lib.F_MAG, "mag")
            p = _lib.TwiParams()
            L.pdm_default_twi_params(ct.byref(p))
            p.twi_min_slope = float(self.twi_min_slope)
            p.twi_min_area = float(self.twi_min_area)
            p.uca_saturation_limit = float(self.uca_saturation_limit)
            p.apply_twi_limits = int(bool(self.apply_twi_limits))
            p.apply_twi_limits_on_uca = int(bool(self.apply_twi_limits_on_uca))
            _lib.check(L.pdm_tile_twi(t, ct.byref(p)))
            twi = self._down(_lib.F_TWI)
            self._twi10 = None                                              # 1674: self.twi = 10 * twi, formed on first read
            self._twi_raw = twi
        finally:
            self._end()
        return twi

    # debugging attributes of the reference (dem_processing.py:137-139)
    def calc_section_proportion(self):
        """Fill ``self.section`` / ``self.proportion`` (reference _calc_uca_section_proportion :1021)
        from the current direction / flats.  Test hook; not needed for calc_uca."""
        self._begin()
        try:
            L = self._lib()
            t = self._get_tile()
            self._spacing()
            self._up(_lib.F_ELEV, "elev")
            self._up(_lib.F_DIR, "direction")
            self._up(_lib.F_MAG, "mag")
            self._up(_lib.F_FLATS, "flats")
            self.section = self._down(_lib.F_SECTION)    # (synchronous: computed on demand)
        finally:
            self._end()
        return self.section


def _install_lazy():
    for _name in DEMProcessor._LAZY:
        setattr(DEMProcessor, _name, property(lambda self, _n=_name: self._lazy_get(_n),
                                              lambda self, v, _n=_name: self._lazy_set(_n, v)))


_install_lazy()


def _pack_edges(edge_init_data, R, C):
    """[data, done, todo] dicts keyed left/right/top/bottom -> 12 contiguous strips in the
    order the C ABI takes them (data x4, done x4, todo x4)."""
    keys = ("left", "right", "top", "bottom")
    lens = {"left": R, "right": R, "top": C, "bottom": C}
    if edge_init_data is None:
        data = {k: np.zeros(lens[k]) for k in keys}
        done = {k: np.zeros(lens[k], bool) for k in keys}
        todo = {k: np.zeros(lens[k], bool) for k in keys}
    else:
        data, done, todo = edge_init_data
    out = []
    for d, dt in ((data, np.float64), (done, np.uint8), (todo, np.uint8)):
        for k in keys:
            a = np.ascontiguousarray(np.asarray(d[k]).reshape(-1), dtype=dt)
            if a.size != lens[k]:
                raise ValueError("edge strip %r has %d entries, expected %d" % (k, a.size, lens[k]))
            out.append(a)
    return out


def _raster_kwargs(fn):
    """Elevation + per-row spacing from a raster file (reference utils.py:46-51), read by this
    package's own baseline-GeoTIFF parser (pydem_b200/raster_io.py; rasterio / geopy are not needed)."""
    from .raster_io import dem_processor_from_raster_kwargs
    return dem_processor_from_raster_kwargs(fn)
