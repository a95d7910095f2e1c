"""CPU oracle for the pyDEM hot path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``OracleDEMProcessor`` restates ``pydem.dem_processing.DEMProcessor``'s
``calc_slopes_directions / calc_uca / calc_twi`` (reference dem_processing.py:587, 682, 1647)
on top of ``oracle/pdm_oracle.c`` (per-cell C restatement of the array kernels and of the
Cython sweep) plus a little NumPy glue for the per-tile bookkeeping of
``_calc_uca_chunk`` (864-987) and ``_calc_uca_chunk_update`` (778-862).

Parity status: PINNED -- checked against the unmodified reference imported in place
(tests/test_oracle_vs_reference.py) and the committed fixtures made from it (tests/golden).

The elevation-conditioning flags (``fill_flats``, ``drain_pits_path``; SURVEY.md section 8(f)
rank 1, outside the hot path) are not restated here; the oracle treats ``elev`` as already
conditioned.
"""
import ctypes as ct
import os
import subprocess
import warnings

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "pdm_oracle.c")
_SO = os.path.join(_HERE, "libpdm_oracle.so")

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i8p = np.ctypeslib.ndpointer(dtype=np.int8, flags="C_CONTIGUOUS")
_i64 = ct.c_int64


class _PitParams(ct.Structure):
    _fields_ = [("max_iter", ct.c_int64), ("max_dist", ct.c_int64),
                ("max_dist_xy", ct.c_double), ("min_border", ct.c_int)]


def build(force=False):
    """gcc the C restatement (strict IEEE: no FMA contraction)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared",
                               "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = ct.CDLL(build())
    L.orc_slopes_directions.argtypes = [_f64p, _i64, _i64, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p]
    L.orc_flats.argtypes = [_f64p, _f64p, _i64, _i64, _u8p]
    L.orc_section_proportion.argtypes = [_f64p, _u8p, _f64p, _i64, _i64, _i8p, _f64p]
    L.orc_receivers.argtypes = [_i8p, _i64, _i64, _i64p, _i64p]
    L.orc_pits.argtypes = [_f64p, _u8p, _f64p, _f64p, _f64p, _i64, _i64, ct.POINTER(_PitParams),
                           ct.POINTER(_i64), ct.POINTER(ct.POINTER(_i64)), ct.POINTER(ct.POINTER(_i64)),
                           ct.POINTER(ct.POINTER(ct.c_double)), ct.POINTER(_i64)]
    L.orc_free.argtypes = [ct.c_void_p]
    L.orc_graph_build.restype = ct.c_void_p
    L.orc_graph_build.argtypes = [_f64p, _i64, _i64p, _i64p, _f64p, _i64, _i64p, _i64p, _f64p]
    L.orc_graph_nnz.restype = _i64
    L.orc_graph_nnz.argtypes = [ct.c_void_p]
    L.orc_graph_export.argtypes = [ct.c_void_p, _i64p, _i64p, _f64p, _i64p, _i64p]
    L.orc_graph_sums.argtypes = [ct.c_void_p, _f64p, _f64p, _i64p]
    L.orc_graph_free.argtypes = [ct.c_void_p]
    L.orc_drain_area.restype = _i64
    L.orc_drain_area.argtypes = [ct.c_void_p, _f64p, _u8p, _u8p, _i64, _i64, ct.c_void_p, ct.c_int,
                                 ct.POINTER(_i64)]
    L.orc_drain_connections.restype = _i64
    L.orc_drain_connections.argtypes = [ct.c_void_p, _u8p, _u8p, ct.c_int]
    L.orc_twi.argtypes = [_f64p, _f64p, _i64, ct.c_double, ct.c_double, ct.c_double, ct.c_int, ct.c_int, _f64p]
    L.orc_np_sum.restype = ct.c_double
    L.orc_np_sum.argtypes = [_f64p, _i64]
    _lib = L
    return L


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def row_thetas(dX, dY):
    """theta of the two facet families per fence, as the reference computes them
    (np.arctan2(d2, d1), dem_processing.py:1936)."""
    dX = np.asarray(dX, "float64"); dY = np.asarray(dY, "float64")
    return np.arctan2(dY, dX), np.arctan2(dX, dY)


# --------------------------------------------------------------------------------------
# kernels
# --------------------------------------------------------------------------------------
def slopes_directions(elev, dX, dY):
    """a1 -> (mag, direction) before flats are stamped."""
    E = _c(elev, "float64"); R, C = E.shape
    dX = _c(dX, "float64"); dY = _c(dY, "float64")
    thA, thB = row_thetas(dX, dY)
    mag = np.empty_like(E); direction = np.empty_like(E)
    rc = lib().orc_slopes_directions(E, R, C, dX, dY, _c(thA, "float64"), _c(thB, "float64"), mag, direction)
    if rc:
        raise ValueError("oracle: grid must be at least 3x3")
    return mag, direction


def find_flats_edges(elev, mag):
    """a2 -> bool flats."""
    E = _c(elev, "float64"); R, C = E.shape
    f = np.zeros(E.shape, np.uint8)
    rc = lib().orc_flats(E, _c(mag, "float64"), R, C, f)
    assert rc == 0
    return f.astype(bool)


def section_proportion(direction, flats, dX, dY):
    """a3 -> (section int8, proportion f64)."""
    D = _c(direction, "float64"); R, C = D.shape
    thA, _ = row_thetas(dX, dY)
    sec = np.empty(D.shape, np.int8); prop = np.empty(D.shape, np.float64)
    rc = lib().orc_section_proportion(D, _c(flats, np.uint8), _c(thA, "float64"), R, C, sec, prop)
    if rc:
        raise IndexError("oracle: section outside [-8, 7] (the reference raises IndexError too)")
    return sec, prop


def receivers(section):
    sec = _c(section, np.int8); R, C = sec.shape
    j1 = np.empty(sec.shape, np.int64); j2 = np.empty(sec.shape, np.int64)
    lib().orc_receivers(sec, R, C, j1, j2)
    return j1, j2


def pit_edges(elev, flats, mag, dX, dY, max_iter=300, max_dist=32, max_dist_xy=None, min_border=False):
    """a5.  Mutates ``flats`` (uint8/bool array) and ``mag`` in place like the reference.
    Returns (pit_i, pit_j, pit_prop, n_warn)."""
    E = _c(elev, "float64"); R, C = E.shape
    f8 = _c(flats, np.uint8)
    m = mag if (mag.dtype == np.float64 and mag.flags.c_contiguous) else _c(mag, "float64")
    pp = _PitParams(int(max_iter), int(max_dist or 0), float(max_dist_xy or 0.0), int(bool(min_border)))
    ne = _i64(0); nw = _i64(0)
    pi = ct.POINTER(_i64)(); pj = ct.POINTER(_i64)(); pw = ct.POINTER(ct.c_double)()
    lib().orc_pits(E, f8, m, _c(dX, "float64"), _c(dY, "float64"), R, C, ct.byref(pp),
                   ct.byref(ne), ct.byref(pi), ct.byref(pj), ct.byref(pw), ct.byref(nw))
    n = ne.value
    pit_i = np.ctypeslib.as_array(pi, (max(n, 1),))[:n].copy()
    pit_j = np.ctypeslib.as_array(pj, (max(n, 1),))[:n].copy()
    pit_w = np.ctypeslib.as_array(pw, (max(n, 1),))[:n].copy()
    for p in (pi, pj, pw):
        lib().orc_free(ct.cast(p, ct.c_void_p))
    flats[...] = f8.astype(flats.dtype)
    if m is not mag:
        mag[...] = m
    return pit_i, pit_j, pit_w, nw.value


class Graph(object):
    """Drainage matrix A of _mk_adjacency_matrix (1072-1153) as CSC+CSR."""

    def __init__(self, elev, j1, j2, prop, pit_i=None, pit_j=None, pit_prop=None):
        E = _c(elev, "float64").ravel()
        self.N = E.size
        if pit_i is None:
            pit_i = np.zeros(0, np.int64); pit_j = np.zeros(0, np.int64); pit_prop = np.zeros(0, np.float64)
        self._h = lib().orc_graph_build(E, self.N, _c(j1, np.int64).ravel(), _c(j2, np.int64).ravel(),
                                        _c(prop, "float64").ravel(), len(pit_i), _c(pit_i, np.int64),
                                        _c(pit_j, np.int64), _c(pit_prop, "float64"))
        self.nnz = lib().orc_graph_nnz(self._h)

    def export(self):
        cptr = np.empty(self.N + 1, np.int64); rptr = np.empty(self.N + 1, np.int64)
        cidx = np.empty(self.nnz, np.int64); ridx = np.empty(self.nnz, np.int64)
        cdat = np.empty(self.nnz, np.float64)
        lib().orc_graph_export(self._h, cptr, cidx, cdat, rptr, ridx)
        return cptr, cidx, cdat, rptr, ridx

    def sums(self):
        inflow = np.empty(self.N); outflow = np.empty(self.N); indeg = np.empty(self.N, np.int64)
        lib().orc_graph_sums(self._h, inflow, outflow, indeg)
        return inflow, outflow, indeg

    def drain_area(self, area, done, ids, R, C, edge_todo=None, skip_edge=False):
        """cyutils.drain_area; arrays are 1-D and mutated in place.  Returns (rounds, cells drained)."""
        nd = _i64(0)
        et = None if edge_todo is None else edge_todo.ctypes.data_as(ct.c_void_p)
        rounds = lib().orc_drain_area(self._h, area, done, ids, R, C, et, int(bool(skip_edge)), ct.byref(nd))
        return rounds, nd.value

    def drain_connections(self, arr, ids, set_to):
        return lib().orc_drain_connections(self._h, arr, ids, int(bool(set_to)))

    def __del__(self):
        try:
            if self._h:
                lib().orc_graph_free(self._h); self._h = None
        except Exception:
            pass


def twi(uca, mag, min_slope, min_area, sat_limit=32.0, limit_uca=False, limit_twi=False):
    u = _c(uca, "float64"); m = _c(mag, "float64")
    out = np.empty_like(u)
    lib().orc_twi(u.ravel(), m.ravel(), u.size, float(min_slope), float(min_area), float(sat_limit),
                  int(bool(limit_uca)), int(bool(limit_twi)), out.ravel())
    return out


# --------------------------------------------------------------------------------------
# the operator
# --------------------------------------------------------------------------------------
class OracleDEMProcessor(object):
    """Same attribute / method surface as the reference DEMProcessor for the hot path."""

    _FLAGS = dict(fill_flats=True, fill_flats_below_sea=False, fill_flats_source_tol=1, fill_flats_peaks=True,
                  fill_flats_pits=True, maximum_pit_area=32.0, drain_pits=True, drain_pits_path=True, drain_pits_min_border=False,
                  drain_pits_max_iter=300, drain_pits_max_dist=32, drain_pits_max_dist_XY=None,
                  apply_uca_limit_edges=False, apply_twi_limits=False, apply_twi_limits_on_uca=False,
                  uca_saturation_limit=32.0, twi_min_slope=1e-3, twi_min_area=np.inf,
                  circular_ref_maxcount=50)

    def __init__(self, elev, dX=None, dY=None, dX2=None, dY2=None, **kwargs):
        self.elev = np.array(elev, dtype="float64", copy=True)
        R = self.elev.shape[0]
        # scalar spacing broadcast: dem_processing.py:233-240
        if not isinstance(dX, np.ndarray):
            if dX2 is None:
                dX2 = np.ones(R) * (1 if dX is None else dX)
            dX = np.ones(R - 1) * (1 if dX is None else dX)
        if not isinstance(dY, np.ndarray):
            if dY2 is None:
                dY2 = np.ones(R) * (1 if dY is None else dY)
            dY = np.ones(R - 1) * (1 if dY is None else dY)
        self.dX = np.asarray(dX, "float64"); self.dY = np.asarray(dY, "float64")
        self.dX2 = np.ones(R) if dX2 is None else np.asarray(dX2, "float64")
        self.dY2 = np.ones(R) if dY2 is None else np.asarray(dY2, "float64")
        for k, v in self._FLAGS.items():
            setattr(self, k, kwargs.pop(k, v))
        for k in ("direction", "mag", "uca", "twi", "flats"):
            setattr(self, k, kwargs.pop(k, None))
        if kwargs:
            raise TypeError("unknown arguments: %s" % sorted(kwargs))
        self.edge_todo = self.edge_done = None
        self.stats = {}

    def find_flats(self):
        self.flats = self.mag == -1

    def calc_fill_flats(self):
        from . import conditioning
        self.elev = conditioning.fill_flats(self.elev, self.fill_flats_below_sea, self.fill_flats_source_tol,
                                            self.fill_flats_peaks, self.fill_flats_pits, self.maximum_pit_area)

    def calc_pit_drain_paths(self):
        from . import conditioning
        self.elev, n_bad, _ = conditioning.pit_drain_paths(self.elev, self.dX, self.dY, self.fill_flats_below_sea,
                                                           self.drain_pits_max_iter, self.drain_pits_max_dist,
                                                           self.drain_pits_max_dist_XY)
        self.stats["pits_without_drain_path"] = n_bad
        return self.elev

    def calc_slopes_directions(self, plotflag=False):
        # conditioning (601-609): restated in oracle/conditioning.py
        if self.fill_flats:
            self.calc_fill_flats()
        if self.drain_pits_path:
            self.calc_pit_drain_paths()
        mag, direction = slopes_directions(self.elev, self.dX, self.dY)
        flats = find_flats_edges(self.elev, mag)
        direction[flats] = -1; mag[flats] = -1                          # 611-612
        self.mag, self.direction, self.flats = mag, direction, flats
        return self.mag, self.direction

    # -- graph shared by full and update mode (787-793 / 873-879) --------------------------
    def _graph(self):
        sec, prop = section_proportion(self.direction, self.flats, self.dX, self.dY)
        j1, j2 = receivers(sec)
        pit = (None, None, None)
        if self.drain_pits:
            pi, pj, pw, nwarn = pit_edges(self.elev, self.flats, self.mag, self.dX, self.dY,
                                          self.drain_pits_max_iter, self.drain_pits_max_dist,
                                          self.drain_pits_max_dist_XY, self.drain_pits_min_border)
            self.stats["pit_edges"] = len(pi); self.stats["pits_undrained"] = nwarn
            pit = (pi, pj, pw)
        self.section, self.proportion = sec, prop
        return Graph(self.elev, j1, j2, prop, *pit), sec

    def calc_uca(self, plotflag=False, edge_init_data=None, uca_init=None):
        if self.direction is None:
            self.calc_slopes_directions()
        if uca_init is None:
            return self._uca_full()
        return self._uca_update(edge_init_data, uca_init)

    def _uca_full(self):
        """_calc_uca_chunk dem_processing.py:864-987."""
        E = self.elev; R, C = E.shape
        g, sec = self._graph()
        flats = self.flats                                              # after pit mutation
        inflow, outflow, _ = g.sums()
        ids = (inflow == 0)                                             # 882-883
        rowarea = self.dX2 * self.dY2                                   # 885
        self.twi_min_area = min(self.twi_min_area, np.nanmin(rowarea))  # 898-899
        min_area = np.nanmin(rowarea)
        area = np.repeat(rowarea.reshape(R, 1), C, 1).astype("float64") # 901
        done = ids.copy().reshape(R, C)
        outflow2 = outflow.reshape(R, C); inflow2 = inflow.reshape(R, C)
        TOL = 1e-2
        todo = np.zeros((R, C), bool)                                   # 909-930
        todo[:, 0] = (outflow2[:, 0] > TOL) & np.isin(sec[:, 0], [6, 7, 0, 1])
        todo[:, -1] = (outflow2[:, -1] > TOL) & np.isin(sec[:, -1], [2, 3, 4, 5])
        todo[0, :] = (outflow2[0, :] > TOL) & np.isin(sec[0, :], [4, 5, 6, 7])
        todo[-1, :] = (outflow2[-1, :] > TOL) & np.isin(sec[-1, :], [0, 1, 2, 3])
        for (a, b) in ((0, 0), (0, -1), (-1, 0), (-1, -1)):
            todo[a, b] |= (outflow2[a, b] > TOL) | (inflow2[a, b] < TOL)
        todo[np.isnan(E)] = False                                       # 935
        todo_i = todo.copy()                                            # 937
        area_ = area.ravel(); done_ = np.ascontiguousarray(done.ravel(), np.uint8)
        ids_ = np.ascontiguousarray(ids, np.uint8)
        taint = todo.astype("float64").ravel()
        e_ = E.ravel()
        count = 1; done_sum = 0; rounds = 0; drained = 0
        while (not done_.all()) and count < self.circular_ref_maxcount and done_sum != int(done_.sum()):  # 951-952
            done_sum = int(done_.sum()); count += 1
            r, d = g.drain_area(area_, done_, ids_, R, C, taint, skip_edge=False)   # 956-960
            rounds += r; drained += d
            with np.errstate(invalid="ignore", divide="ignore"):
                und = e_ * (done_ == 0)                                 # 963
                max_elev = und.max()
                ids_ = np.ascontiguousarray((und - max_elev) / max_elev > -0.01, np.uint8)  # 964
        self.stats.update(rounds=rounds, drained=drained, restarts=count - 2)
        area = area_.reshape(R, C)
        area[flats] = np.nan                                            # 972
        edge_done = ~(taint.reshape(R, C).astype(bool))                 # 969, 974
        edge_done[np.isnan(E)] = True                                   # 975
        if self.apply_uca_limit_edges:                                  # 977-980
            with np.errstate(invalid="ignore"):
                edge_done[area > self.uca_saturation_limit * 2 * min_area] = True
        self.uca, self.edge_todo, self.edge_done = area, todo_i, edge_done
        self.done = done_.reshape(R, C).astype(bool)
        return self.uca

    def _uca_update(self, edge_init_data, uca_init):
        """calc_uca edge packing 719-744 + _calc_uca_chunk_update 778-862."""
        E = self.elev; R, C = E.shape
        e_init = np.zeros((R, C)); e_done = np.zeros((R, C), bool); e_todo = np.zeros((R, C), bool)
        sl = {"left": (slice(None), slice(0, 1)), "right": (slice(None), slice(-1, None)),
              "top": (slice(0, 1), slice(None)), "bottom": (slice(-1, None), slice(None))}
        if edge_init_data is not None:
            data, done_d, todo_d = edge_init_data
            for k, v in sl.items():                                     # 730-737
                shp = e_init[v].shape
                e_done[v] = e_done[v] | np.asarray(done_d[k]).reshape(shp)
                e_init[v] += (np.asarray(data[k]) * np.asarray(done_d[k])).reshape(shp)
                e_todo[v] = e_todo[v] | np.asarray(todo_d[k]).reshape(shp)
            for k, v in sl.items():                                     # 738-739
                e_init[v][~e_done[v]] = 0
        uca0 = np.asarray(uca_init).astype("float64")
        g, _ = self._graph()
        flats = self.flats
        ids = (e_done & e_todo)                                         # 798
        e_todo = e_todo & ~e_done                                       # 799
        area = np.zeros((R, C))
        # 806-809 (left, right, bottom, top -- later assignments win at the corners)
        area[e_done[:, 0], 0] = e_init[e_done[:, 0], 0] - uca0[e_done[:, 0], 0]
        area[e_done[:, -1], -1] = e_init[e_done[:, -1], -1] - uca0[e_done[:, -1], -1]
        area[-1, e_done[-1, :]] = e_init[-1, e_done[-1, :]] - uca0[-1, e_done[-1, :]]
        area[0, e_done[0, :]] = e_init[0, e_done[0, :]] - uca0[0, e_done[0, :]]
        ids_ = np.ascontiguousarray(ids.ravel(), np.uint8)
        area[flats] = np.nan                                            # 815
        todo_i = e_todo.copy()                                          # 817
        done = np.ones(R * C, np.uint8); done[ids_ != 0] = 0            # 820-821
        g.drain_connections(done, ids_, False)                          # 823-824
        done[np.isnan(E).ravel()] = 1                                   # 827
        done[ids_ != 0] = 1                                             # 831
        area_ = area.ravel()
        g.drain_area(area_, done, ids_, R, C, None, skip_edge=False)    # 836-842
        t8 = np.ascontiguousarray(e_todo.ravel(), np.uint8)
        g.drain_connections(t8, t8.copy(), True)                        # 848-853
        e_todo = t8.reshape(R, C).astype(bool)
        area = area_.reshape(R, C)
        area[flats] = np.nan                                            # 855
        self.uca = uca0 + area                                          # 769
        self.edge_todo = todo_i                                         # 770
        self.edge_done = ~e_todo                                        # 771, 856
        return self.uca

    def calc_twi(self):
        if self.uca is None:
            self.calc_uca()
        t = twi(self.uca, self.mag, self.twi_min_slope, self.twi_min_area, self.uca_saturation_limit,
                self.apply_twi_limits_on_uca, self.apply_twi_limits)
        self.twi = t * 10                                               # 1674
        return t
