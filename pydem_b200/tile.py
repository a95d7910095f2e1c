"""``DeviceTile``: thin Python handle over the ``pdm_tile_*`` C ABI -- fields stay in HBM
between stages.  Used by bench.py (device-resident throughput), the sharded driver and the
tests; ``DEMProcessor`` is the reference-compatible operator built on the same calls."""
import ctypes as ct

import numpy as np

from . import _lib
from ._lib import (F_CELL, F_DIR, F_EDGE_DONE, F_EDGE_TODO, F_ELEV, F_FLAT0, F_FLATS, F_LINK, F_MAG,  # noqa: F401
                   F_SECTION, F_TWI, F_UCA)


class DeviceTile(object):
    def __init__(self, R, C, device=None, stream=None):
        self.L = _lib.load()
        _lib.init(device)
        self.R, self.C = int(R), int(C)
        self.shape = (self.R, self.C)
        h = ct.c_void_p()
        _lib.check(self.L.pdm_tile_create(self.R, self.C, ct.c_void_p(stream or 0), ct.byref(h)))
        self.h = h
        self.min_area = None

    def close(self):
        if getattr(self, "h", None) is not None:
            self.L.pdm_tile_destroy(self.h)
            self.h = None

    __del__ = close

    def set_spacing(self, dX, dY, dX2=None, dY2=None):
        R = self.R
        dX = np.ascontiguousarray(np.broadcast_to(np.asarray(dX, "float64"), (R - 1,)))
        dY = np.ascontiguousarray(np.broadcast_to(np.asarray(dY, "float64"), (R - 1,)))
        dX2 = np.ascontiguousarray(np.broadcast_to(np.asarray(dX[0] if dX2 is None else dX2, "float64"), (R,)))
        dY2 = np.ascontiguousarray(np.broadcast_to(np.asarray(dY[0] if dY2 is None else dY2, "float64"), (R,)))
        thA = np.ascontiguousarray(np.arctan2(dY, dX)); thB = np.ascontiguousarray(np.arctan2(dX, dY))
        P = _lib.ptr
        _lib.check(self.L.pdm_tile_set_spacing(self.h, P(dX), P(dY), P(dX2), P(dY2), P(thA), P(thB)))
        self.min_area = float(np.nanmin(dX2 * dY2))

    def upload(self, field, arr):
        a = np.ascontiguousarray(arr, dtype=_lib.FIELD_DTYPE[field])
        assert a.shape == self.shape, (a.shape, self.shape)
        _lib.check(self.L.pdm_tile_upload(self.h, field, _lib.ptr(a)))

    def download(self, field, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=_lib.FIELD_DTYPE[field])
        _lib.check(self.L.pdm_tile_download(self.h, field, _lib.ptr(out)))
        return out

    def device_ptr(self, field):
        p = ct.c_void_p()
        _lib.check(self.L.pdm_tile_device_ptr(self.h, field, ct.byref(p)))
        return p.value

    def as_torch(self, field):
        """Zero-copy torch view of a field (for NCCL send/recv of halo rows)."""
        import torch

        class _CAI(object):
            pass
        o = _CAI()
        dt = np.dtype(_lib.FIELD_DTYPE[field])
        shape, typestr = self.shape, dt.str
        if dt.kind == "V":                       # 32-byte sweep records: rows of raw bytes
            shape, typestr = (self.R, self.C * dt.itemsize), "|u1"
        o.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(self.device_ptr(field), False),
                                          version=2, strides=None)
        return torch.as_tensor(o, device="cuda")

    # ---- row shards (include/pydem_b200.h, "row shards") ------------------------------------
    def set_window(self, row_off, R_global, own_lo, own_hi, th_row=None):
        a = None if th_row is None else np.ascontiguousarray(th_row, "float64")
        assert a is None or a.shape == (self.R,)
        _lib.check(self.L.pdm_tile_set_window(self.h, int(row_off), int(R_global), int(own_lo), int(own_hi), _lib.ptr(a)))
        self.own = (int(own_lo), int(own_hi))

    def shard_stage(self, name, *args):
        _lib.check(getattr(self.L, "pdm_shard_" + name)(self.h, *args))

    def shard_links(self, **flags):
        p = _lib.UcaParams()
        self.L.pdm_default_uca_params(ct.byref(p))
        p.drain_pits = 0
        for k, v in flags.items():
            setattr(p, k, v)
        self._shard_params = p
        _lib.check(self.L.pdm_shard_links(self.h, ct.byref(p)))

    def shard_finalize(self):
        st = _lib.UcaStats()
        _lib.check(self.L.pdm_shard_finalize(self.h, ct.byref(self._shard_params), ct.byref(st)))
        return st.as_dict()

    def mark_resident(self, field):
        _lib.check(self.L.pdm_tile_mark_resident(self.h, field))

    def set_stencil_parity(self, on):
        _lib.check(self.L.pdm_tile_set_stencil_parity(self.h, int(bool(on))))

    def sync(self):
        _lib.check(self.L.pdm_tile_sync(self.h))

    def slopes_directions(self):
        _lib.check(self.L.pdm_tile_slopes_directions(self.h))

    def find_flats(self):
        _lib.check(self.L.pdm_tile_find_flats(self.h))

    def condition(self, stage, **flags):
        """stage: 'fill_pit_artifacts' | 'fill_flats' | 'pit_drain_paths' (ELEV modified in HBM)."""
        p = _lib.CondParams()
        self.L.pdm_default_cond_params(ct.byref(p))
        for k, v in flags.items():
            setattr(p, k, v)
        st = _lib.CondStats()
        _lib.check(getattr(self.L, "pdm_tile_" + stage)(self.h, ct.byref(p), ct.byref(st)))
        return {k: getattr(st, k) for k, _ in st._fields_}

    def uca(self, **flags):
        p = _lib.UcaParams()
        self.L.pdm_default_uca_params(ct.byref(p))
        for k, v in flags.items():
            setattr(p, k, v)
        st = _lib.UcaStats()
        _lib.check(self.L.pdm_tile_uca(self.h, ct.byref(p), ct.byref(st)))
        return st.as_dict()

    def uca_update(self, strips, **flags):
        """Edge-update mode on the resident UCA (= uca_init): strips = the 12 arrays of
        dem_processing._pack_edges (data x4, done x4, todo x4; left, right, top, bottom)."""
        p = _lib.UcaParams()
        self.L.pdm_default_uca_params(ct.byref(p))
        for k, v in flags.items():
            setattr(p, k, v)
        st = _lib.UcaStats()
        _lib.check(self.L.pdm_tile_uca_update(self.h, ct.byref(p), *[_lib.ptr(a) for a in strips], ct.byref(st)))
        return st.as_dict()

    def set_keep_graph(self, on=True):
        _lib.check(self.L.pdm_tile_set_keep_graph(self.h, int(bool(on))))

    def twi(self, twi_min_area=None, **flags):
        p = _lib.TwiParams()
        self.L.pdm_default_twi_params(ct.byref(p))
        p.twi_min_area = self.min_area if twi_min_area is None else twi_min_area
        for k, v in flags.items():
            setattr(p, k, v)
        _lib.check(self.L.pdm_tile_twi(self.h, ct.byref(p)))
