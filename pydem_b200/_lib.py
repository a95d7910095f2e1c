"""ctypes binding of ``libpydem_b200.so`` (C ABI: include/pydem_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is usable,
the first compute call raises ``RuntimeError`` -- the same way the reference raises when
its Cython extension is not compiled (dem_processing.py:714-715).
"""
import ctypes as ct
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYDEM_B200_LIB") or os.path.join(_HERE, "libpydem_b200.so")   # override: A/B of library builds

# pdm_field
(F_ELEV, F_MAG, F_DIR, F_FLATS, F_UCA, F_TWI, F_EDGE_TODO, F_EDGE_DONE, F_SECTION, F_TWI10, F_RESERVED, F_FLAT0,
 F_LINK, F_CELL) = range(14)
FIELD_DTYPE = {F_ELEV: np.float64, F_MAG: np.float64, F_DIR: np.float64, F_FLATS: np.uint8,
               F_UCA: np.float64, F_TWI: np.float64, F_EDGE_TODO: np.uint8, F_EDGE_DONE: np.uint8,
               F_SECTION: np.int8, F_TWI10: np.float64, F_FLAT0: np.uint8, F_LINK: np.uint8,
               F_CELL: np.dtype((np.void, 32))}


class UcaParams(ct.Structure):
    _fields_ = [("drain_pits", ct.c_int32), ("drain_pits_min_border", ct.c_int32),
                ("drain_pits_max_iter", ct.c_int64), ("drain_pits_max_dist", ct.c_int64),
                ("drain_pits_max_dist_xy", ct.c_double), ("apply_uca_limit_edges", ct.c_int32),
                ("circular_ref_maxcount", ct.c_int32), ("uca_saturation_limit", ct.c_double)]


class UcaStats(ct.Structure):
    _fields_ = [("n_cells", ct.c_int64), ("n_sources", ct.c_int64), ("n_drained", ct.c_int64),
                ("n_undone", ct.c_int64), ("n_pits", ct.c_int64), ("n_pit_edges", ct.c_int64),
                ("n_pits_undrained", ct.c_int64), ("n_queue_items", ct.c_int64), ("n_restarts", ct.c_int64),
                ("n_edge_todo", ct.c_int64), ("min_area", ct.c_double), ("ms_graph", ct.c_float),
                ("ms_sweep", ct.c_float), ("ms_total", ct.c_float), ("ms_sweep_scan", ct.c_float),
                ("ms_sweep_kernel", ct.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class TwiParams(ct.Structure):
    _fields_ = [("twi_min_slope", ct.c_double), ("twi_min_area", ct.c_double),
                ("uca_saturation_limit", ct.c_double), ("apply_twi_limits", ct.c_int32),
                ("apply_twi_limits_on_uca", ct.c_int32)]


class CondParams(ct.Structure):
    _fields_ = [("fill_flats_below_sea", ct.c_int32), ("fill_flats_source_tol", ct.c_int32),
                ("fill_flats_peaks", ct.c_int32), ("fill_flats_pits", ct.c_int32),
                ("maximum_pit_area", ct.c_double), ("drain_pits_max_iter", ct.c_int32),
                ("drain_pits_max_dist", ct.c_int32), ("drain_pits_max_dist_xy", ct.c_double)]


class CondStats(ct.Structure):
    _fields_ = [(k, ct.c_int64) for k in ("n_artifact_regions", "n_artifacts_filled", "n_flat_regions",
                                          "distance_sweeps", "n_pits", "n_pits_undrained", "path_max_iter")]


# every symbol include/pydem_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "pdm_abi_version", "pdm_last_error", "pdm_init", "pdm_device_count", "pdm_default_uca_params",
    "pdm_default_twi_params", "pdm_launch_count", "pdm_shard_p2p_export", "pdm_shard_p2p_connect", "pdm_shard_p2p_connect_all", "pdm_shard_p2p_disconnect", "pdm_set_sweep_mode", "pdm_get_sweep_mode", "pdm_host_alloc", "pdm_host_free", "pdm_tile_create", "pdm_tile_destroy", "pdm_tile_set_spacing",
    "pdm_tile_upload", "pdm_tile_download", "pdm_tile_download_async", "pdm_tile_device_ptr", "pdm_tile_mark_resident", "pdm_tile_set_stencil_parity", "pdm_tile_sync", "pdm_selftest_division",
    "pdm_tile_slopes_directions", "pdm_tile_upload_slopes_directions", "pdm_tile_find_flats", "pdm_tile_uca", "pdm_tile_pit_updates", "pdm_tile_uca_update",
    "pdm_tile_twi", "pdm_tile_set_keep_graph", "pdm_tile_set_window", "pdm_shard_slopes", "pdm_shard_ccl", "pdm_shard_label_pack",
    "pdm_shard_label_unpack", "pdm_shard_flats_extend", "pdm_shard_links", "pdm_tile_set_global_spacing", "pdm_shard_pits", "pdm_shard_pit_in_apply", "pdm_shard_indeg", "pdm_shard_sweep",
    "pdm_shard_sweep_sent", "pdm_shard_finalize",
    "pdm_comm_unique_id", "pdm_comm_init", "pdm_comm_finalize", "pdm_comm_barrier", "pdm_shard_connect", "pdm_shard_disconnect", "pdm_shard_run",
    "pdm_slopes_directions", "pdm_uca", "pdm_uca_update", "pdm_twi",
    "pdm_default_cond_params", "pdm_tile_fill_pit_artifacts", "pdm_tile_fill_flats", "pdm_tile_pit_drain_paths",
]

_lib = None
_vp = ct.c_void_p
_i64 = ct.c_int64


def load():
    """dlopen the library and declare prototypes.  Raises RuntimeError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "pydem_b200: CUDA library %s is not built (run `python -m pydem_b200.build`); "
            "there is no CPU fallback" % LIB_PATH)
    L = ct.CDLL(LIB_PATH)
    L.pdm_abi_version.restype = ct.c_int
    L.pdm_last_error.restype = ct.c_char_p
    L.pdm_init.argtypes = [ct.c_int]
    L.pdm_device_count.argtypes = [ct.POINTER(ct.c_int)]
    L.pdm_launch_count.restype = ct.c_ulonglong
    L.pdm_set_sweep_mode.argtypes = [ct.c_int, ct.c_int]
    L.pdm_host_alloc.argtypes = [ct.c_size_t, ct.POINTER(_vp)]
    L.pdm_host_free.argtypes = [_vp]
    L.pdm_default_uca_params.argtypes = [ct.POINTER(UcaParams)]
    L.pdm_default_uca_params.restype = None
    L.pdm_default_twi_params.argtypes = [ct.POINTER(TwiParams)]
    L.pdm_default_twi_params.restype = None
    L.pdm_tile_create.argtypes = [_i64, _i64, _vp, ct.POINTER(_vp)]
    L.pdm_tile_destroy.argtypes = [_vp]
    L.pdm_tile_set_spacing.argtypes = [_vp] + [_vp] * 6
    L.pdm_tile_upload.argtypes = [_vp, ct.c_int, _vp]
    L.pdm_tile_download.argtypes = [_vp, ct.c_int, _vp]
    L.pdm_tile_download_async.argtypes = [_vp, ct.c_int, _vp]
    L.pdm_tile_device_ptr.argtypes = [_vp, ct.c_int, ct.POINTER(_vp)]
    L.pdm_tile_mark_resident.argtypes = [_vp, ct.c_int]
    L.pdm_tile_set_stencil_parity.argtypes = [_vp, ct.c_int]
    L.pdm_tile_sync.argtypes = [_vp]
    L.pdm_selftest_division.argtypes = [ct.c_ulonglong, ct.c_longlong, ct.POINTER(ct.c_ulonglong)]
    L.pdm_tile_slopes_directions.argtypes = [_vp]
    L.pdm_tile_upload_slopes_directions.argtypes = [_vp, _vp]
    L.pdm_tile_find_flats.argtypes = [_vp]
    L.pdm_tile_uca.argtypes = [_vp, ct.POINTER(UcaParams), ct.POINTER(UcaStats)]
    L.pdm_tile_pit_updates.argtypes = [_vp, _i64, _vp, _vp, _vp, ct.POINTER(_i64)]
    L.pdm_tile_uca_update.argtypes = [_vp, ct.POINTER(UcaParams)] + [_vp] * 12 + [ct.POINTER(UcaStats)]
    L.pdm_tile_twi.argtypes = [_vp, ct.POINTER(TwiParams)]
    L.pdm_tile_set_keep_graph.argtypes = [_vp, ct.c_int]
    L.pdm_default_cond_params.argtypes = [ct.POINTER(CondParams)]
    L.pdm_default_cond_params.restype = None
    for nm in ("pdm_tile_fill_pit_artifacts", "pdm_tile_fill_flats", "pdm_tile_pit_drain_paths"):
        getattr(L, nm).argtypes = [_vp, ct.POINTER(CondParams), ct.POINTER(CondStats)]
    L.pdm_tile_set_window.argtypes = [_vp, _i64, _i64, _i64, _i64, _vp]
    for nm in ("pdm_shard_slopes", "pdm_shard_ccl", "pdm_shard_flats_extend", "pdm_shard_indeg"):
        getattr(L, nm).argtypes = [_vp]
    L.pdm_shard_label_pack.argtypes = [_vp, _i64, _vp, _vp]
    L.pdm_shard_label_unpack.argtypes = [_vp, _i64, _vp, _vp, _vp]
    L.pdm_shard_links.argtypes = [_vp, ct.POINTER(UcaParams)]
    L.pdm_tile_set_global_spacing.argtypes = [_vp, _vp, _vp, _i64]
    L.pdm_shard_pits.argtypes = [_vp, ct.POINTER(UcaParams), _vp, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _i64]
    L.pdm_shard_pit_in_apply.argtypes = [_vp, _vp, _vp, _i64]
    L.pdm_shard_sweep.argtypes = [_vp, ct.c_int]
    L.pdm_shard_p2p_export.argtypes = [_vp, _vp, ct.POINTER(_i64)]
    L.pdm_shard_p2p_connect.argtypes = [_vp, _vp, _vp, _vp, ct.c_int, ct.c_int]
    L.pdm_shard_p2p_connect_all.argtypes = [_vp, _vp, ct.c_int, ct.c_int]
    L.pdm_shard_p2p_disconnect.argtypes = [_vp]
    L.pdm_shard_sweep_sent.argtypes = [_vp, _vp]
    L.pdm_shard_finalize.argtypes = [_vp, ct.POINTER(UcaParams), ct.POINTER(UcaStats)]
    L.pdm_comm_unique_id.argtypes = [_vp]
    L.pdm_comm_init.argtypes = [ct.c_int, ct.c_int, _vp]
    L.pdm_comm_barrier.argtypes = [_vp]
    L.pdm_shard_connect.argtypes = [_vp]
    L.pdm_shard_disconnect.argtypes = [_vp]
    L.pdm_shard_run.argtypes = [_vp, ct.POINTER(UcaParams), ct.POINTER(TwiParams), ct.POINTER(UcaStats), ct.POINTER(ct.c_int)]
    L.pdm_slopes_directions.argtypes = [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.pdm_uca.argtypes = [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                          ct.POINTER(UcaParams), _vp, _vp, _vp, ct.POINTER(UcaStats)]
    L.pdm_uca_update.argtypes = [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                 ct.POINTER(UcaParams)] + [_vp] * 12 + [_vp, _vp]
    L.pdm_twi.argtypes = [_vp, _vp, _i64, ct.POINTER(TwiParams), _vp]
    if L.pdm_abi_version() != 2:
        raise RuntimeError("pydem_b200: ABI version mismatch (%d)" % L.pdm_abi_version())
    _lib = L
    return L


class PdmError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().pdm_last_error().decode(errors="replace")
        if rc == 4:
            raise IndexError(msg)  # the reference raises IndexError for out-of-range sections
        if rc == 1:
            raise ValueError(msg)
        raise PdmError("pydem_b200 [status %d]: %s" % (rc, msg))


_device = None


def init(device=None):
    """Bind this process to a CUDA device (lazy; call again after fork in a new worker)."""
    global _device
    L = load()
    if device is None:
        if _device is not None:
            return _device
        device = int(os.environ.get("PYDEM_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    check(L.pdm_init(int(device)))
    _device = int(device)
    return _device


def ptr(a):
    return None if a is None else ct.c_void_p(a.ctypes.data)


SWEEP_MODES = {"worklist": 0, "tile": 1, "env": -1}


def set_sweep(mode="env", strict=None):
    """Engine of the UCA accumulation on stand-alone tiles: "worklist" (default, fastest), "tile"
    (pull-based, bit-reproducible; the row-sharded path always uses it) or "env"
    (PYDEM_B200_SWEEP).  strict=True makes the work-list hold every count-off back until its
    adds have returned (cross-check)."""
    check(load().pdm_set_sweep_mode(SWEEP_MODES[mode], -1 if strict is None else int(bool(strict))))


def get_sweep():
    return "worklist" if load().pdm_get_sweep_mode() == 0 else "tile"
