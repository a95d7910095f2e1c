"""In-memory tile orchestrator for mosaics: the caller of the DEM hot path.

Mirrors the semantics of the reference's ``pydem.process_manager.ProcessManager`` -- grid and
overlap geometry (process_manager.py:517-740), the per-tile workers ``calc_elev_cond`` (:54),
``calc_aspect_slope`` (:73), ``calc_uca`` (:94, including the one-pixel-overlap patch of edge
directions), ``calc_uca_ec_metrics`` (:199), ``calc_uca_ec`` (:224, including the corner
double-count rule) and ``calc_twi`` (:296), and the serial scheduling loop of
``process_uca_edges`` (:1090-1189) -- with two differences of form, not of result:

* no zarr store and no GeoTIFF files: tiles are NumPy arrays handed in with their pixel box in
  the mosaic; every tile keeps its own result arrays, and a neighbour's edge strip is read from
  the neighbour's arrays at the pixel the reference's ``edge_data`` slice addresses;
* the ``DEMProcessor`` is this package's CUDA operator (any class with the reference's operator
  surface can be passed, which is how the CPU tests run the same orchestrator on the oracle).

Tile-parallel mode (``group=``): the tiles are dealt round-robin to the ranks of a
``torch.distributed`` job (one GPU each).  A rank keeps full arrays only for its own tiles; of the
other tiles it keeps the edge rings (the strips within overlap distance of a tile edge -- all a
neighbour ever reads, see ``_RingArray``), refreshed with one ``all_gather_object`` per stage /
correction round.  The one-pixel-overlap patch of ``calc_uca`` is order dependent and only touches
edge cells, so every rank replays it for all tiles on the rings (bit-identical to the serial
order); the correction loop takes the reference's serial decisions on every rank (the metrics
only need rings) and the owner of the chosen tile executes it, so the result is the one-rank
result by construction and only the stages around the corrections run in parallel.

The reference's own orchestrator, run unmodified over in-memory stand-ins for zarr/rasterio
(``oracle/ref_pm_harness.py``), produced ``tests/golden/ref_pm.npz``; ``tests/test_process_manager.py``
compares this module against it array by array (aspect, slope, uca, uca_edges, edge masks, twi,
and the order in which tiles were corrected).
"""
import warnings

import json
import os

import numpy as np

SIDES = ("left", "right", "top", "bottom")
CORNERS = ("top-left", "top-right", "bottom-left", "bottom-right")
# where a tile's own edge lives inside its arrays (process_manager.py:41-50)
EDGE = {
    "left": (slice(0, None), 0), "right": (slice(0, None), -1), "top": (0, slice(0, None)), "bottom": (-1, slice(0, None)),
    "top-left": (0, 0), "top-right": (0, -1), "bottom-left": (-1, 0), "bottom-right": (-1, -1),
}
# direction ranges (radians CCW from east) that leave the tile through a side (process_manager.py:102-107)
DOWNSTREAM = {"left": (np.pi / 2, 3 * np.pi / 2), "right": (2 * np.pi - np.pi / 2, np.pi / 2), "top": (0.0, np.pi), "bottom": (np.pi, 2 * np.pi)}


def split_mosaic(shape, ny_grid, nx_grid, overlap):
    """Pixel boxes (r0, r1, c0, c1) of an ny_grid x nx_grid tiling of a ``shape`` raster whose
    neighbouring tiles share ``overlap`` pixels, row-major.  Same cut as the reference's test
    generator (utils_test_pydem.mk_test_multifile :359-408): interior seams are centred on the
    regular grid lines."""
    def cuts(n, k):
        step = -(-n // k)
        starts = [s for s in range(0, n - overlap, step)]
        lo = [starts[0]] + [s - overlap // 2 for s in starts[1:]]
        hi = [s + (overlap + 1) // 2 for s in starts[1:]] + [n]
        return list(zip(lo, [min(h, n) for h in hi]))
    return [(r0, r1, c0, c1) for (r0, r1) in cuts(shape[0], ny_grid) for (c0, c1) in cuts(shape[1], nx_grid)]


class _RingArray(object):
    """The part of a remote tile's result array that neighbours read: the first / last `w` rows and
    columns.  Supports exactly the accesses of the orchestrator (a row or column of the ring, a
    corner), reading and writing; corner cells live in a row strip and a column strip and are kept
    consistent on writes."""

    def __init__(self, shape, w, dtype):
        self.shape = tuple(shape); self.w = int(w); self.dtype = np.dtype(dtype)
        R, C = self.shape
        w = min(self.w, R, C)
        self.w = w
        self.top = np.zeros((w, C), self.dtype); self.bottom = np.zeros((w, C), self.dtype)
        self.left = np.zeros((R, w), self.dtype); self.right = np.zeros((R, w), self.dtype)

    @classmethod
    def of(cls, a, w):
        r = cls(a.shape, w, a.dtype)
        w = r.w
        r.top[:] = a[:w]; r.bottom[:] = a[-w:]; r.left[:] = a[:, :w]; r.right[:] = a[:, -w:]
        return r

    def strips(self):
        return (self.top, self.bottom, self.left, self.right)

    def set_strips(self, st):
        self.top[:], self.bottom[:], self.left[:], self.right[:] = st

    def _norm(self, key):
        r, c = key
        R, C = self.shape
        if isinstance(r, (int, np.integer)) and r < 0: r += R
        if isinstance(c, (int, np.integer)) and c < 0: c += C
        return r, c

    def _row(self, r):
        R = self.shape[0]
        if r < self.w: return self.top, r
        if r >= R - self.w: return self.bottom, r - (R - self.w)
        raise IndexError("row %d is not in the edge ring (width %d)" % (r, self.w))

    def _col(self, c):
        C = self.shape[1]
        if c < self.w: return self.left, c
        if c >= C - self.w: return self.right, c - (C - self.w)
        raise IndexError("column %d is not in the edge ring (width %d)" % (c, self.w))

    def __getitem__(self, key):
        r, c = self._norm(key)
        if isinstance(r, slice):
            a, k = self._col(c)
            return a[r, k]
        a, k = self._row(r)
        return a[k, c]

    def __setitem__(self, key, value):
        r, c = self._norm(key)
        R, C = self.shape
        w = self.w
        if isinstance(r, slice):                      # a whole column of the ring
            a, k = self._col(c)
            a[r, k] = value
            col = a[:, k]
            self.top[:, c] = col[:w]; self.bottom[:, c] = col[R - w:]
        elif isinstance(c, slice):                    # a whole row
            a, k = self._row(r)
            a[k, c] = value
            row = a[k]
            self.left[r, :] = row[:w]; self.right[r, :] = row[C - w:]
        else:                                         # one cell (a corner of the tile)
            a, k = self._row(r); a[k, c] = value
            a, k = self._col(c); a[r, k] = value


class _Tile(object):
    __slots__ = ("box", "gi", "gj", "elev", "dX", "dY", "dX2", "dY2", "aspect", "slope", "uca", "uca_edges",
                 "edge_todo", "edge_done", "twi", "trim", "edge_src", "owner", "shape")


class ProcessManager(object):
    """tiles: list of elevation arrays; boxes: their (r0, r1, c0, c1) pixel boxes in the mosaic, in
    processing order (the reference processes its source files in sorted-name order).
    spacing: dict(dX=, dY=, dX2=, dY2=) of scalars applied to every tile (the reference derives
    per-row arrays from each GeoTIFF; its tests force 1 everywhere), or a list of such dicts of
    per-tile arrays.  dem_proc_kwargs: flags forwarded to every DEMProcessor (reference trait of
    the same name).  dem_processor: operator class (default: the CUDA DEMProcessor)."""

    def __init__(self, tiles, boxes, spacing=None, dem_proc_kwargs=None, dem_processor=None, n_workers=1, group=None, out_path=None):
        """group: None (one process) or an object with `rank`, `world` and `all_gather(obj) -> list`
        (see `TorchGroup`); tile k belongs to rank k % world.  Every rank passes the same `boxes`;
        a rank may pass None for the elevation of tiles it does not own."""
        self.group = group
        self.rank = 0 if group is None else int(group.rank)
        self.world = 1 if group is None else int(group.world)
        if dem_processor is None:
            from .dem_processing import DEMProcessor as dem_processor
        self.DEMProcessor = dem_processor
        self.dem_proc_kwargs = dict(dem_proc_kwargs or {})
        self.n_workers = int(n_workers)
        if len(tiles) != len(boxes) or not tiles:
            raise ValueError("need one pixel box per tile")
        self.tiles = []
        for k, (e, b) in enumerate(zip(tiles, boxes)):
            t = _Tile()
            t.box = tuple(int(v) for v in b)
            t.owner = k % self.world
            t.shape = (t.box[1] - t.box[0], t.box[3] - t.box[2])
            if t.owner == self.rank:
                t.elev = np.array(e, dtype="float64")
                if t.elev.shape != t.shape:
                    raise ValueError("tile %d: shape %s does not match its box %s" % (k, t.elev.shape, t.box))
            else:
                t.elev = None
            sp = (spacing[k] if isinstance(spacing, (list, tuple)) else spacing) or {}
            R = t.shape[0]
            t.dX = np.ones(R - 1) * sp.get("dX", 1.0); t.dY = np.ones(R - 1) * sp.get("dY", 1.0)     # scalars or per-row arrays
            t.dX2 = np.ones(R) * sp.get("dX2", sp.get("dX", 1.0)) if not isinstance(sp.get("dX2"), np.ndarray) else sp["dX2"]
            t.dY2 = np.ones(R) * sp.get("dY2", sp.get("dY", 1.0)) if not isinstance(sp.get("dY2"), np.ndarray) else sp["dY2"]
            for nm in ("aspect", "slope", "uca", "uca_edges", "edge_todo", "edge_done", "twi"):
                setattr(t, nm, None)
            self.tiles.append(t)
        self.n_inputs = len(self.tiles)
        self.uca_edge_metrics = np.zeros((self.n_inputs, 2))
        self.success = np.zeros((self.n_inputs, 4), bool)
        self.correction_log = []       # tiles in the order process_uca_edges corrected them
        # out_path: a directory store of the per-tile results with the reference's resume semantics (its zarr
        # arrays + the `success` array, process_manager.py:362-381, 998-1007): every stage writes what it produced
        # and marks the tile; a manager created on an existing store skips what is already there
        self.out_path = out_path
        self.edges_done = False
        self.georef = None             # set by from_directory: extreme bounds of the inputs
        self.compute_grid()
        self.compute_grid_overlaps()
        # ring width: the farthest a neighbour reaches into a tile (the larger overlap, at least 1)
        self.ring_w = 1
        for t in self.tiles:
            for (di, dj, ax) in ((0, -1, 1), (0, 1, 1), (-1, 0, 0), (1, 0, 0)):
                o = self._nb(t, di, dj)
                if o is not None:
                    shared = (o.box[3] - t.box[2]) if dj < 0 else (t.box[3] - o.box[2]) if dj > 0 else \
                             (o.box[1] - t.box[0]) if di < 0 else (t.box[1] - o.box[0])
                    self.ring_w = max(self.ring_w, int(shared))

        if self.out_path is not None:
            self._store_load()

    # ------------------------------------------------------------------------------------
    # directory store + resume (the reference keeps every result in zarr arrays and a `success` array of
    # n_inputs x 4 stages, and queue_processes skips what is marked: process_manager.py:998-1007, 1251-1288)
    # ------------------------------------------------------------------------------------
    STAGE_KEYS = (("elev",), ("aspect", "slope"), ("aspect", "slope", "uca", "uca_edges", "edge_todo", "edge_done"), ("twi",))

    def _store_file(self, key, i):
        return os.path.join(self.out_path, key, "%05d.npy" % i)

    def _store_save(self, i, keys):
        t = self.tiles[i]
        if not self._mine(t):
            return
        for key in keys:
            a = getattr(t, key)
            if a is None or isinstance(a, _RingArray):
                continue
            fn = self._store_file(key, i)
            os.makedirs(os.path.dirname(fn), exist_ok=True)
            with open(fn + ".tmp", "wb") as f:
                np.save(f, np.asarray(a))
            os.replace(fn + ".tmp", fn)               # a reader never sees a half-written array

    def _store_mark(self):
        if self.rank == 0 or self.group is None:
            with open(os.path.join(self.out_path, "success.npy.tmp"), "wb") as f:
                np.save(f, self.success)
            os.replace(os.path.join(self.out_path, "success.npy.tmp"), os.path.join(self.out_path, "success.npy"))
            with open(os.path.join(self.out_path, "edges.json.tmp"), "w") as f:
                json.dump({"edges_done": bool(self.edges_done), "correction_log": [int(k) for k in self.correction_log]}, f)
            os.replace(os.path.join(self.out_path, "edges.json.tmp"), os.path.join(self.out_path, "edges.json"))

    def _store_load(self):
        os.makedirs(self.out_path, exist_ok=True)
        fn = os.path.join(self.out_path, "success.npy")
        if not os.path.exists(fn):
            return
        ok = np.load(fn)
        if ok.shape != self.success.shape:
            raise ValueError("%s holds the results of another tiling (%s tiles)" % (self.out_path, ok.shape[0]))
        for i, t in enumerate(self.tiles):
            for col in range(4):
                if not ok[i, col]:
                    break                                   # a stage needs the ones before it
                if self._mine(t):
                    for key in self.STAGE_KEYS[col]:
                        f = self._store_file(key, i)
                        if not os.path.exists(f):
                            ok[i, col:] = False
                            break
                        a = np.load(f)
                        if a.shape != t.shape:
                            raise ValueError("%s: shape %s does not match tile %d %s" % (f, a.shape, i, t.shape))
                        setattr(t, key, a)
                    else:
                        continue
                    break
        self.success = ok
        ej = os.path.join(self.out_path, "edges.json")
        if os.path.exists(ej) and self.success[:, 2].all():
            with open(ej) as f:
                d = json.load(f)
            self.edges_done = bool(d.get("edges_done"))
            self.correction_log = [int(k) for k in d.get("correction_log", [])]

    INPUT_FILE_TYPES = ("tif", "tiff")      # of the reference's _INPUT_FILE_TYPES (:462-463), what raster_io reads

    @classmethod
    def from_directory(cls, in_path, unit_spacing=False, grid_round_decimals=2, **kwargs):
        """The reference's entry point: a directory of elevation GeoTIFFs (ProcessManager(in_path=...),
        :461-472, compute_index :484-509).  Files are taken in sorted order; a tile's pixel box in
        the mosaic follows from its bounds and pixel size, its spacing from its georeferencing
        (utils.mk_dx_dy_from_geotif_layer via pydem_b200.raster_io).  unit_spacing=True forces
        dX = dY = 1 like the reference's DEBUG switch (process_manager.py:52, used by its tests)."""
        import os
        from . import raster_io
        files = sorted(os.path.join(in_path, f) for f in os.listdir(in_path)
                       if os.path.splitext(f)[-1].replace(".", "").lower() in cls.INPUT_FILE_TYPES)
        if not files:
            raise ValueError("no elevation rasters (%s) in %s" % (", ".join(cls.INPUT_FILE_TYPES), in_path))
        rs = [raster_io.read_geotiff(f) for f in files]
        a = rs[0]["transform"].a; e = rs[0]["transform"].e
        for f, r in zip(files, rs):
            if abs(r["transform"].a - a) > 1e-9 * abs(a) or abs(r["transform"].e - e) > 1e-9 * abs(e):
                raise ValueError("%s: tiles of different resolution are not supported by the in-memory orchestrator" % f)
        left = min(r["bounds"][0] for r in rs); top = max(r["bounds"][3] for r in rs)
        tiles, boxes, spacing = [], [], []
        for r in rs:
            H, W = r["elev"].shape
            c0 = int(round((r["bounds"][0] - left) / a)); r0 = int(round((top - r["bounds"][3]) / abs(e)))
            tiles.append(np.asarray(r["elev"], dtype="float64")); boxes.append((r0, r0 + H, c0, c0 + W))
            if unit_spacing:
                spacing.append(dict(dX=1.0, dY=1.0, dX2=1.0, dY2=1.0))
            else:
                dX, dY, dX2, dY2 = raster_io.mk_dx_dy(r["transform"], H, r["is_projected"], r["ellipsoid"])
                spacing.append(dict(dX=dX, dY=dY, dX2=dX2, dY2=dY2))
        pm = cls(tiles, boxes, spacing=spacing, **kwargs)
        pm.elev_source_files = files
        pm.georef = dict(left=left, top=top, right=max(r["bounds"][2] for r in rs), bottom=min(r["bounds"][1] for r in rs),
                         projected=rs[0]["is_projected"])
        return pm

    def save_geotiff(self, filename, key, dtype="float64", rescale=None):
        """The non-overlapping mosaic of `key` as one GeoTIFF (save_geotiff :862-931: transform from the
        extreme bounds of the inputs over the mosaic size, optional
        (data - rescale[0]) / (rescale[1] - rescale[0]) * rescale[2]).  Uncompressed, 512 x 512 tiles
        (the reference writes LZW through rasterio).  Needs a manager made by from_directory."""
        from . import raster_io
        if getattr(self, "georef", None) is None:
            raise RuntimeError("save_geotiff needs the georeferencing of the inputs: build the manager with from_directory()")
        m = self.mosaic(key)
        g = self.georef
        dlat = (g["bottom"] - g["top"]) / m.shape[0]; dlon = (g["right"] - g["left"]) / m.shape[1]
        if rescale:
            m = (m - rescale[0]) / (rescale[1] - rescale[0]) * rescale[2]
        if self.rank == 0:
            raster_io.write_geotiff(filename, m.astype(dtype), raster_io.Affine((dlon, 0.0, g["left"], 0.0, dlat, g["top"])),
                                    projected=g["projected"], tile=512 if max(m.shape) > 512 else None)
        return filename

    # ------------------------------------------------------------------------------------
    # geometry (process_manager.py:517-740)
    # ------------------------------------------------------------------------------------
    def compute_grid(self):
        """Place every tile in the (row, column) grid of the mosaic: tiles that start on the same
        mosaic row / column form a grid row / column (the reference does this with the rounded
        top/left bounds of the files)."""
        r0s = sorted(set(t.box[0] for t in self.tiles)); c0s = sorted(set(t.box[2] for t in self.tiles))
        self.grid_shape = (len(r0s), len(c0s))
        self.grid_id2i = -np.ones(self.grid_shape, dtype=int)
        for i, t in enumerate(self.tiles):
            t.gi, t.gj = r0s.index(t.box[0]), c0s.index(t.box[2])
            if self.grid_id2i[t.gi, t.gj] >= 0:
                raise ValueError("two tiles start at mosaic pixel (%d, %d)" % (t.box[0], t.box[2]))
            self.grid_id2i[t.gi, t.gj] = i
        for i, t in enumerate(self.tiles):      # a grid row has one height, a grid column one width (:537-545)
            for o in self.tiles:
                if o.gi == t.gi and o.shape[0] != t.shape[0] or o.gj == t.gj and o.shape[1] != t.shape[1]:
                    raise ValueError("tiles of one grid row / column must have the same number of rows / columns")

    def _nb(self, t, di, dj):
        i, j = t.gi + di, t.gj + dj
        if i < 0 or j < 0 or i >= self.grid_shape[0] or j >= self.grid_shape[1]:
            return None
        k = self.grid_id2i[i, j]
        return None if k < 0 else self.tiles[k]

    @staticmethod
    def _overlap(shared, tie):
        """(pixels this tile gives up on that side, distance of the neighbour's copy of my edge
        from the neighbour's own edge) for `shared` common pixels (_calc_overlap :567-599)."""
        return int(np.round((shared + tie - 0.01) / 2)), max(int(np.round(shared)), 1)

    def compute_grid_overlaps(self):
        for t in self.tiles:
            R, C = t.shape
            lf, rt, up, dn = self._nb(t, 0, -1), self._nb(t, 0, 1), self._nb(t, -1, 0), self._nb(t, 1, 0)
            c_lo, c_lo_e = self._overlap(lf.box[3] - t.box[2], 0) if lf else (0, 0)
            c_hi, c_hi_e = self._overlap(t.box[3] - rt.box[2], 1) if rt else (0, 0)
            r_lo, r_lo_e = self._overlap(up.box[1] - t.box[0], 0) if up else (0, 0)
            r_hi, r_hi_e = self._overlap(t.box[1] - dn.box[0], 1) if dn else (0, 0)
            t.trim = (r_lo, r_hi, c_lo, c_hi)
            # Where the values for my edges come from: (tile, row index/slice, column index/slice).
            # The reference addresses its side-by-side store one step outside the tile's block
            # (edge_data :700-711); with no neighbour on a side that step is zero and the edge
            # data are the tile's own edge.
            def src(di, dj, rows, cols):
                o = self._nb(t, di, dj) if (di or dj) else t
                return (o if o is not None else None, rows, cols)
            tl = 1 if (t.gj > 0 and t.gi > 0) else 0
            bl = 1 if (t.gj > 0 and t.gi < self.grid_shape[0] - 1) else 0
            tr = 1 if (t.gj < self.grid_shape[1] - 1 and t.gi > 0) else 0
            br = 1 if (t.gj < self.grid_shape[1] - 1 and t.gi < self.grid_shape[0] - 1) else 0

            def corner(on, di, dj, re, ce, own_r, own_c):
                # the block one step up/down by `re` and left/right by `ce` (each may be 0)
                sdi = di if (on and re) else 0; sdj = dj if (on and ce) else 0
                o = self._nb(t, sdi, sdj) if (sdi or sdj) else t
                if o is None:
                    return (None, 0, 0)
                oR, oC = o.shape
                rr = (oR - re if di < 0 else re - 1) if sdi else own_r
                cc = (oC - ce if dj < 0 else ce - 1) if sdj else own_c
                return (o, rr, cc)
            t.edge_src = {
                "left": src(0, -1 if c_lo_e else 0, slice(0, None), (lf.shape[1] - c_lo_e) if c_lo_e else 0),
                "right": src(0, 1 if c_hi_e else 0, slice(0, None), (c_hi_e - 1) if c_hi_e else C - 1),
                "top": src(-1 if r_lo_e else 0, 0, (up.shape[0] - r_lo_e) if r_lo_e else 0, slice(0, None)),
                "bottom": src(1 if r_hi_e else 0, 0, (r_hi_e - 1) if r_hi_e else R - 1, slice(0, None)),
                "top-left": corner(tl, -1, -1, r_lo_e, c_lo_e, 0, 0),
                "top-right": corner(tr, -1, 1, r_lo_e, c_hi_e, 0, C - 1),
                "bottom-left": corner(bl, 1, -1, r_hi_e, c_lo_e, R - 1, 0),
                "bottom-right": corner(br, 1, 1, r_hi_e, c_hi_e, R - 1, C - 1),
            }
            t_one = {}
            # check_1overlap (:286-294): is the source exactly one step outside my block?
            for key, (o, rr, cc) in t.edge_src.items():
                steps = []
                if key in ("left", "top-left", "bottom-left"): steps.append(c_lo_e if (key == "left" or (tl if key == "top-left" else bl)) else 0)
                if key in ("right", "top-right", "bottom-right"): steps.append(c_hi_e if (key == "right" or (tr if key == "top-right" else br)) else 0)
                if key in ("top", "top-left", "top-right"): steps.append(r_lo_e if (key == "top" or (tl if key == "top-left" else tr)) else 0)
                if key in ("bottom", "bottom-left", "bottom-right"): steps.append(r_hi_e if (key == "bottom" or (bl if key == "bottom-left" else br)) else 0)
                t_one[key] = any(s == 1 for s in steps)
            t.edge_src["_one"] = t_one

    def _edge(self, t, key, field):
        """The reference's ``store[field][edge_slice[key]]`` for tile t (a copy)."""
        o, rr, cc = t.edge_src[key]
        if o is None:
            return np.zeros((), dtype=getattr(t, field).dtype)[()] if key in CORNERS else np.zeros(
                t.shape[0] if key in ("left", "right") else t.shape[1], dtype=getattr(t, field).dtype)
        a = getattr(o, field)
        v = a[rr, cc]
        return np.array(v) if isinstance(v, np.ndarray) else v

    def _dp(self, t, **kw):
        k = dict(elev=t.elev.copy(), dX=t.dX.copy(), dY=t.dY.copy(), dX2=t.dX2.copy(), dY2=t.dY2.copy())
        k.update(kw)
        k.update(self.dem_proc_kwargs)
        return self.DEMProcessor(**k)

    # ------------------------------------------------------------------------------------
    # workers
    # ------------------------------------------------------------------------------------
    def _elev_cond(self, t):                                    # calc_elev_cond :54-71
        dp = self._dp(t)
        dp.calc_fill_flats()
        dp.calc_pit_drain_paths()
        t.elev = np.array(dp.elev, dtype="float64")

    def _aspect_slope(self, t):                                 # calc_aspect_slope :73-92
        dp = self._dp(t, fill_flats=False)
        dp.calc_slopes_directions()
        t.aspect = np.array(dp.direction, dtype="float64"); t.slope = np.array(dp.mag, dtype="float64")

    @staticmethod
    def _leaves(d, key):
        lo, hi = DOWNSTREAM[key]
        if key == "right":
            return ((d >= lo) | (d <= hi)) & (d >= 0)
        return (d >= lo) & (d <= hi)

    @staticmethod
    def _east_west(d):
        return ((d < 1e-6) & (d >= 0)) | (np.abs(d - np.pi * 2) < 1e-6)

    def _patch_edges(self, t):
        """The one-pixel-overlap patch of calc_uca (:109-180): where this tile's edge is also the
        neighbour's edge, cells whose flow LEAVES through that edge (and flats) take the neighbour's
        direction and magnitude, which were computed with the cells beyond the edge in view.  Works
        in place on t.aspect / t.slope -- full arrays for own tiles, edge rings for remote ones --
        in the reference's order: four sides, then four corners."""
        A, S = t.aspect, t.slope
        one = t.edge_src["_one"]
        for key in SIDES:
            if not one[key]:
                continue
            d = np.array(A[EDGE[key]]); m = np.array(S[EDGE[key]])
            ids = self._leaves(d, key)                          # only edges the flow leaves through
            if key in ("top", "bottom"):
                ids = ids | self._east_west(d)
            ids = ids | (d == -1)                               # and flats
            other = ("top", "bottom") if key in ("left", "right") else ("left", "right")
            # the two end cells belong to the corner rule below when they also leave through the adjacent side
            ids[0] = ids[0] & (not bool(self._leaves(d[0:1], other[0])[0]))
            ids[-1] = ids[-1] & (not bool(self._leaves(d[-1:], other[1])[0]))
            nb_a = self._edge(t, key, "aspect"); nb_s = self._edge(t, key, "slope")
            d[ids] = nb_a[ids]; m[ids] = nb_s[ids]
            A[EDGE[key]] = d; S[EDGE[key]] = m                  # the stored fields are patched (:149-150)
        for key in CORNERS:
            if not one[key]:
                continue
            ktb, klr = key.split("-")
            da = np.array([A[EDGE[key]]])
            if not bool((self._leaves(da, klr) & (self._leaves(da, ktb) | self._east_west(da)))[0]):
                continue
            A[EDGE[key]] = self._edge(t, key, "aspect"); S[EDGE[key]] = self._edge(t, key, "slope")

    def _uca(self, t):                                          # calc_uca :94-197 (after _patch_edges)
        dp = self._dp(t, direction=t.aspect.copy(), mag=t.slope.copy(), fill_flats=False)
        dp.find_flats()
        dp.calc_uca()
        t.uca = np.array(dp.uca, dtype="float64")
        t.edge_todo = np.array(dp.edge_todo, dtype=bool); t.edge_done = np.array(dp.edge_done, dtype=bool)
        if t.uca_edges is None:
            t.uca_edges = np.zeros_like(t.uca)

    def _metrics(self, t):                                      # calc_uca_ec_metrics :199-221
        n_done = 0; p_done = 0
        for key in SIDES + ("top-left", "top-right", "bottom-left", "bottom-right"):
            et = t.edge_todo[EDGE[key]]
            edn = self._edge(t, key, "edge_done")
            n_done += np.sum(et & edn); p_done += np.sum(et)
        return (n_done / (1e-16 + p_done), n_done)

    def _uca_ec(self, t):                                       # calc_uca_ec :224-284
        dp = self._dp(t, direction=t.aspect.copy(), mag=t.slope.copy(), fill_flats=False)
        dp.find_flats()
        uca_init = t.uca.copy(); edges_init = t.uca_edges.copy()
        data = {k: self._edge(t, k, "uca") + self._edge(t, k, "uca_edges") for k in SIDES}
        done = {k: self._edge(t, k, "edge_done") for k in SIDES}
        todo = {k: np.array(t.edge_todo[EDGE[k]]) for k in SIDES}
        todo_nb = {k: self._edge(t, k, "edge_todo") for k in SIDES}
        # a corner cell sits on two strips; when both neighbours have finished it, one of them is
        # dropped so that its inflow is not counted twice, and the diagonal neighbour's value is
        # preferred where that neighbour really is one step away and done (:257-270)
        for key in ("top-left", "bottom-right", "top-right", "bottom-left"):
            ktb, klr = key.split("-")
            ir, ic = EDGE[key]
            if done[ktb][ic] & done[klr][ir]:
                done[ktb][ic] = False
                if t.edge_src["_one"][key] and self._edge(t, key, "edge_done"):
                    v = self._edge(t, key, "uca") + self._edge(t, key, "uca_edges")
                    data[klr][ir] = v; data[ktb][ic] = v
        todo = {k: v & (todo_nb[k] == False) for k, v in todo.items()}  # noqa: E712 (:274)
        dp.calc_uca(uca_init=uca_init + edges_init, edge_init_data=[data, done, todo])
        t.uca_edges = np.array(dp.uca, dtype="float64") - uca_init
        t.edge_todo = np.array(dp.edge_todo, dtype=bool); t.edge_done = np.array(dp.edge_done, dtype=bool)

    def _twi(self, t):                                          # calc_twi :296-315
        dp = self._dp(t, direction=t.aspect.copy(), mag=t.slope.copy(), fill_flats=False, uca=t.uca + t.uca_edges)
        dp.find_flats()
        dp.calc_twi()
        t.twi = np.array(dp.twi, dtype="float64")

    # ------------------------------------------------------------------------------------
    # stages (process_manager.py:993-1316)
    # ------------------------------------------------------------------------------------
    def _mine(self, t):
        return t.owner == self.rank

    def _sync(self, fields, index=None):
        """Refresh the edge rings of the tiles in `index` (default: all) on the ranks that do not own
        them: one all_gather of the owners' strips."""
        if self.group is None:
            return
        index = range(self.n_inputs) if index is None else index
        mine = {}
        for i in index:
            t = self.tiles[i]
            if self._mine(t):
                mine[i] = {f: _RingArray.of(getattr(t, f), self.ring_w).strips() for f in fields}
        for part in self.group.all_gather(mine):
            for i, fs in part.items():
                t = self.tiles[i]
                if self._mine(t):
                    continue
                for f, st in fs.items():
                    ra = getattr(t, f)
                    if not isinstance(ra, _RingArray):
                        ra = _RingArray(t.shape, self.ring_w, st[0].dtype)
                        setattr(t, f, ra)
                    ra.set_strips(st)

    def _stage(self, fn, col):
        for i, t in enumerate(self.tiles):
            if self._mine(t) and not self.success[i, col]:
                fn(t)
                if self.out_path is not None:
                    self._store_save(i, self.STAGE_KEYS[col])
            self.success[i, col] = True
        if self.out_path is not None:
            self._store_mark()
        return self.success[:, col].copy()

    def process_elevation(self):
        return self._stage(self._elev_cond, 0)

    def process_aspect_slope(self):
        r = self._stage(self._aspect_slope, 1)
        self._sync(("aspect", "slope"))
        return r

    def process_uca(self):
        # the edge patch is order dependent (a tile reads what earlier tiles patched) and only touches
        # edge cells: every rank replays it for all tiles, on full arrays or rings
        if not self.success[:, 2].all():
            for t in self.tiles:
                self._patch_edges(t)
        r = self._stage(self._uca, 2)
        self._sync(("uca", "uca_edges", "edge_todo", "edge_done"))
        return r

    def update_uca_edge_metrics(self, index=None):
        for i in (range(self.n_inputs) if index is None else index):
            self.uca_edge_metrics[i] = self._metrics(self.tiles[i])
        return self.uca_edge_metrics.copy()

    def _order(self, mets, mets_type):
        if mets.shape[0] == 1:
            return np.zeros(1, int)
        return np.argpartition(-mets[:, mets_type], min(self.n_workers * 2, mets.shape[0] - 1))

    def process_uca_edges(self, mets_type=0, max_count=100000):
        """The reference's correction loop (:1090-1189, n_workers == 1): correct the tile whose inflow
        edges are most complete, refresh the metrics of that tile and its four neighbours, stop when
        the ranking no longer changes.  On several ranks every rank takes the same decisions (the
        metrics only need edge rings, which all ranks hold); the owner of the chosen tile corrects
        it and its rings are refreshed -- the result is the serial loop's by construction.  (This
        stage does not parallelise under the reference's semantics, see process_uca_edges_rounds.)"""
        mets = self.update_uca_edge_metrics()
        if self.edges_done:                 # resumed from a store whose corrections had finished
            return mets
        I = self._order(mets, mets_type)
        I_old = np.zeros_like(I)
        count = 0
        while np.any(I_old != I) and count < max_count:
            count += 1
            k = int(I[0])
            if self._mine(self.tiles[k]):
                self._uca_ec(self.tiles[k])
            self._sync(("uca_edges", "edge_todo", "edge_done"), [k])
            self.correction_log.append(k)
            I_old = I.copy()
            t = self.tiles[k]
            chk = {k}
            for (di, dj) in ((0, -1), (0, 1), (-1, 0), (1, 0)):
                o = self._nb(t, di, dj)
                if o is not None:
                    chk.add(self.tiles.index(o))
            mets = self.update_uca_edge_metrics(sorted(chk))
            I = self._order(mets, mets_type)
        self.edges_done = True
        if self.out_path is not None:
            for i in range(self.n_inputs):
                self._store_save(i, ("uca_edges", "edge_todo", "edge_done"))
            self._store_mark()
        return mets

    def process_uca_edges_rounds(self, max_rounds=100000):
        """EXPERIMENTAL alternative to the reference's one-tile-at-a-time loop, not used by default:
        every round corrects the tiles that TIE for the best metric (share of inflow edges whose
        neighbour is done), have an edge a neighbour can resolve, and are pairwise non-adjacent
        (8-neighbourhood), so that no corrected tile reads another one of the same round.

        The reference's correction scheme is order dependent, and this schedule is NOT equivalent
        to its loop in general: it reproduces the reference on the 11 pinned tilings, but on 200
        random mosaics 35 differed -- mostly because the reference's loop also corrects a tile when
        no neighbour edge is done yet (which flips todo flags and gets the iteration going), and
        once through a genuine order effect.  Looser rounds ("every tile with a resolvable edge",
        "within half of the best metric") changed up to 168 interior cells even on the pinned
        cases.  Even where it is exact it saves little (19 rounds for 24 corrections on the 5x4
        cone), so the tile-parallel mode keeps the serial loop and parallelises the stages around it."""
        rounds = 0
        while rounds < max_rounds:
            mets = self.update_uca_edge_metrics()
            cand = [int(k) for k in np.argsort(-mets[:, 0], kind="stable") if mets[k, 1] > 0]
            if not cand:
                break
            best = mets[cand[0], 0]
            cand = [k for k in cand if mets[k, 0] >= best - 1e-12]
            chosen, blocked = [], set()
            for k in cand:
                if k in blocked:
                    continue
                chosen.append(k)
                t = self.tiles[k]
                for di in (-1, 0, 1):
                    for dj in (-1, 0, 1):
                        o = self._nb(t, di, dj) if (di or dj) else t
                        if o is not None:
                            blocked.add(self.tiles.index(o))
            for k in chosen:
                if self._mine(self.tiles[k]):
                    self._uca_ec(self.tiles[k])
            self._sync(("uca_edges", "edge_todo", "edge_done"), chosen)
            self.correction_log.append(tuple(chosen))
            rounds += 1
        return self.update_uca_edge_metrics()

    def process_twi(self, rounds=False):
        """The whole pipeline (process_twi :1290-1316).  rounds=True swaps the reference's correction
        loop for the experimental rounds schedule (process_uca_edges_rounds)."""
        self.process_elevation()
        self.process_aspect_slope()
        self.process_uca()
        if rounds:
            self.process_uca_edges_rounds()
        else:
            self.process_uca_edges()
        return self._stage(self._twi, 3)

    # ------------------------------------------------------------------------------------
    # results
    # ------------------------------------------------------------------------------------
    def mosaic(self, key):
        """The non-overlapping mosaic of a result (save_non_overlap_data :742-784): every tile
        contributes the part it does not share, `uca` includes the edge corrections.  On several
        ranks the pieces are gathered and every rank returns the whole mosaic."""
        R = max(t.box[1] for t in self.tiles); C = max(t.box[3] for t in self.tiles)
        pieces = []
        for t in self.tiles:
            if not self._mine(t):
                continue
            a = getattr(t, key)
            if key == "uca":
                a = a + t.uca_edges
            r_lo, r_hi, c_lo, c_hi = t.trim
            nr, nc = t.shape
            pieces.append(((t.box[0] + r_lo, t.box[1] - r_hi, t.box[2] + c_lo, t.box[3] - c_hi), a[r_lo:nr - r_hi, c_lo:nc - c_hi]))
        if self.group is not None:
            pieces = [p for part in self.group.all_gather(pieces) for p in part]
        out = np.full((R, C), np.nan)
        for (r0, r1, c0, c1), a in pieces:
            out[r0:r1, c0:c1] = a
        return out

    def process_overviews(self, keys=("elev", "uca", "aspect", "slope", "twi"), overviews=(3, 9, 27, 81, 243, 729, 2187)):
        """Overview pyramids of the non-overlapping mosaics (process_overviews :933-941 applied to the
        compact store, whose chunks are the mosaic size over the number of grid rows / columns,
        save_non_overlap_data :746-748).  Returns {key: {ov: array}}."""
        res = {}
        for key in keys:
            m = self.mosaic(key)
            chunks = [m.shape[0] // self.grid_shape[0], m.shape[1] // self.grid_shape[1]]
            res[key] = overview_pyramid(m, chunks, overviews)
        return res


def block_means(data, factor):
    """One overview step of a chunk (process_manager.calc_overview :317-353): means over
    factor x factor blocks; the columns / rows / corner left over at the end of the chunk are
    averaged into one extra output column / row / cell.  None for an all-zero chunk (the reference
    skips those)."""
    data = np.asarray(data, dtype="float64")
    if data.size == 0 or np.all(data == 0):
        return None
    R, C = data.shape
    r, c = R // factor, C // factor
    out = np.zeros((-(-R // factor), -(-C // factor)))
    try:
        with np.errstate(invalid="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out[:r, :c] = data[:r * factor, :c * factor].reshape(r, factor, c, factor).transpose(0, 2, 1, 3).reshape(r, c, factor * factor).mean(axis=-1)
            out[:r, c:] = data[:r * factor, c * factor:].reshape(r, -1).mean(axis=1)[:, None]
            out[r:, :c] = data[r * factor:, :c * factor].reshape(-1, c).mean(axis=0)[None, :]
            out[r:, c:] = data[r * factor:, c * factor:].mean()
    except ValueError:
        # a chunk narrower than the factor cannot be reshaped; the reference's worker fails on it too
        # (its exception is swallowed, :349-350) and the chunk stays zero
        return None
    return out


def overview_pyramid(base, chunks, overviews=(3, 9, 27, 81, 243, 729, 2187)):
    """The reference's overview pyramid of one array (process_overviews / _calc_overview :933-991):
    level `ov` is built from the previous level with factor ov // last_ov, chunk by chunk with the
    reference's chunk arithmetic, until a level would be no larger than the factor.
    Returns {ov: array}."""
    out = {}
    last, last_ov = np.asarray(base, dtype="float64"), 1
    chunks = [int(c) for c in chunks]
    for ov in overviews:
        factor = ov // last_ov
        shape = list(last.shape)
        if any(c == 0 for c in chunks):
            chunks = list(shape)
        n_files = [np.ceil(s / c) for s, c in zip(shape, chunks)]
        new_shape = [int(np.ceil(s / factor)) for s in shape]
        new_chunks = [int(max(factor, np.ceil(ns / n))) for ns, n in zip(new_shape, n_files)]
        if any(n <= factor for n in new_shape):
            break
        new = np.zeros(new_shape)
        row = 0
        while (row - 1) * new_chunks[0] < new_shape[0]:
            col = 0
            while (col - 1) * new_chunks[1] < new_shape[1]:
                blk = block_means(last[row * new_chunks[0] * factor:(row + 1) * new_chunks[0] * factor,
                                       col * new_chunks[1] * factor:(col + 1) * new_chunks[1] * factor], factor)
                if blk is not None:
                    tgt = new[row * new_chunks[0]:(row + 1) * new_chunks[0], col * new_chunks[1]:(col + 1) * new_chunks[1]]
                    tgt[...] = blk[:tgt.shape[0], :tgt.shape[1]] if blk.shape != tgt.shape else blk
                col += 1
            row += 1
        out[ov] = new
        last, last_ov, chunks = new, ov, new_chunks
    return out


class TorchGroup(object):
    """The ranks of a torch.distributed job (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self):
        import torch.distributed as dist
        self._dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()

    def all_gather(self, obj):
        out = [None] * self.world
        self._dist.all_gather_object(out, obj)
        return out


class ResidentProcessManager(ProcessManager):
    """The same orchestrator with the tiles RESIDENT on the GPU (BASELINE.json
    configs[4]): a tile's elevation goes to HBM once, slope / aspect / flats / UCA / edge masks stay there
    through all stages, and only the edge rings -- the strips a neighbour can read, `_RingArray` -- live on
    the host, where the reference's scheduling decisions, its one-pixel-overlap patch and its corner rules
    are taken exactly as in the base class.  An edge correction (calc_uca_ec, process_manager.py:224-284)
    then moves four strips of (value, done, todo) to the device, runs the update sweep on the resident UCA
    with the graph of the tile's full sweep (pdm_tile_set_keep_graph; the reference rebuilds it per call,
    dem_processing.py:787-793) and reads the three rings back -- instead of uploading elevation, direction
    and magnitude and downloading every array per call like the host-side manager above.

    Needs the CUDA operator (it drives `pydem_b200.tile.DeviceTile` directly) and a torch CUDA build for the
    zero-copy views of the tile fields.  Elevation conditioning (stage 0) runs through the base class."""

    UCA_FLAGS = ("drain_pits", "drain_pits_min_border", "drain_pits_max_iter", "drain_pits_max_dist",
                 "apply_uca_limit_edges", "uca_saturation_limit", "circular_ref_maxcount")

    def __init__(self, *a, **kw):
        if kw.get("dem_processor") is not None:
            raise ValueError("ResidentProcessManager drives the CUDA tiles directly; dem_processor cannot be replaced")
        if kw.get("out_path") is not None:
            raise ValueError("ResidentProcessManager keeps the results in HBM; use ProcessManager(out_path=...) for a store that can resume")
        ProcessManager.__init__(self, *a, **kw)
        import torch
        from . import tile as T
        self._T, self._torch = T, torch
        self._dev = {}
        self._mag0 = {}
        flags = {}
        for k in self.UCA_FLAGS:
            if k in self.dem_proc_kwargs and self.dem_proc_kwargs[k] is not None:
                flags[k] = self.dem_proc_kwargs[k]
        if "drain_pits_max_dist_XY" in self.dem_proc_kwargs and self.dem_proc_kwargs["drain_pits_max_dist_XY"]:
            flags["drain_pits_max_dist_xy"] = float(self.dem_proc_kwargs["drain_pits_max_dist_XY"])
        self._uca_flags = {k: (int(v) if isinstance(v, (bool, np.bool_)) else v) for k, v in flags.items()}
        self.bytes_moved = dict(h2d=0, d2h=0)

    # ---- device tiles and rings
    def _tile(self, t):
        k = id(t)
        if k not in self._dev:
            T = self._T
            dt = T.DeviceTile(t.shape[0], t.shape[1], stream=self._torch.cuda.current_stream().cuda_stream)
            dt.set_spacing(t.dX, t.dY, t.dX2, t.dY2)
            dt.upload(T.F_ELEV, t.elev)
            dt.set_keep_graph(True)
            self.bytes_moved["h2d"] += t.elev.nbytes
            self._dev[k] = dt
        return self._dev[k]

    def _ring_down(self, t, field, as_bool=False):
        """edge ring of a resident field -> host `_RingArray`"""
        v = self._tile(t).as_torch(field)
        w = min(self.ring_w, t.shape[0], t.shape[1])
        st = [v[:w], v[-w:], v[:, :w], v[:, -w:]]
        st = [x.contiguous().cpu().numpy() for x in st]
        if as_bool:
            st = [x.astype(bool) for x in st]
        ra = _RingArray(t.shape, self.ring_w, st[0].dtype)
        ra.set_strips(st)
        self.bytes_moved["d2h"] += sum(x.nbytes for x in st)
        return ra

    def _ring_up(self, t, field, ra):
        """host ring -> the resident field's edge strips"""
        v = self._tile(t).as_torch(field)
        w = ra.w
        dev = v.device
        for dst, src in ((v[:w], ra.top), (v[-w:], ra.bottom), (v[:, :w], ra.left), (v[:, -w:], ra.right)):
            dst.copy_(self._torch.from_numpy(np.ascontiguousarray(src)).to(dev))
            self.bytes_moved["h2d"] += src.nbytes

    def full(self, t, key):
        """a whole result array of an own tile, read back from the device"""
        T = self._T
        f = {"aspect": T.F_DIR, "slope": T.F_MAG, "uca": T.F_UCA, "twi": T.F_TWI, "edge_todo": T.F_EDGE_TODO,
             "edge_done": T.F_EDGE_DONE, "elev": T.F_ELEV}[key]
        a = self._tile(t).download(f)
        self.bytes_moved["d2h"] += a.nbytes
        if key == "twi":
            a = a * 10.0                        # DEMProcessor.twi stores 10 * twi (dem_processing.py:1674)
        return a.astype(bool) if key in ("edge_todo", "edge_done") else a

    # ---- workers on resident tiles
    def _aspect_slope(self, t):
        T = self._T
        dt = self._tile(t)
        dt.slopes_directions()
        t.aspect = self._ring_down(t, T.F_DIR); t.slope = self._ring_down(t, T.F_MAG)

    def _uca(self, t):
        T = self._T
        dt = self._tile(t)
        self._ring_up(t, T.F_DIR, t.aspect); self._ring_up(t, T.F_MAG, t.slope)      # the one-pixel-overlap patch
        dt.mark_resident(T.F_DIR); dt.mark_resident(T.F_MAG)
        # the pit drains of calc_uca patch mag in place (dem_processing.py:1370-1371), but the reference's TWI stage
        # reads the slope as stored BEFORE calc_uca (process_manager.py:296-315): keep that copy on the device
        self._mag0[id(t)] = dt.as_torch(T.F_MAG).clone()
        dt.find_flats()
        st = dt.uca(**self._uca_flags)
        if st.get("n_pits_undrained"):
            warnings.warn("Warning %d pits had no place to drain to in this chunk" % st["n_pits_undrained"])
        t.uca = self._ring_down(t, T.F_UCA)
        t.edge_todo = self._ring_down(t, T.F_EDGE_TODO, True); t.edge_done = self._ring_down(t, T.F_EDGE_DONE, True)
        t.uca_edges = _RingArray(t.shape, self.ring_w, np.float64)

    def _uca_ec(self, t):
        from .dem_processing import _pack_edges
        T = self._T
        data = {k: self._edge(t, k, "uca") + self._edge(t, k, "uca_edges") for k in SIDES}
        done = {k: self._edge(t, k, "edge_done") for k in SIDES}
        todo = {k: np.array(t.edge_todo[EDGE[k]]) for k in SIDES}
        todo_nb = {k: self._edge(t, k, "edge_todo") for k in SIDES}
        for key in ("top-left", "bottom-right", "top-right", "bottom-left"):        # the corner rule (:257-270), as in the base class
            ktb, klr = key.split("-")
            ir, ic = EDGE[key]
            if done[ktb][ic] & done[klr][ir]:
                done[ktb][ic] = False
                if t.edge_src["_one"][key] and self._edge(t, key, "edge_done"):
                    v = self._edge(t, key, "uca") + self._edge(t, key, "uca_edges")
                    data[klr][ir] = v; data[ktb][ic] = v
        todo = {k: v & (todo_nb[k] == False) for k, v in todo.items()}  # noqa: E712 (:274)
        strips = _pack_edges([data, done, todo], t.shape[0], t.shape[1])
        self.bytes_moved["h2d"] += sum(a.nbytes for a in strips)
        self._tile(t).uca_update(strips, **self._uca_flags)        # the resident UCA is uca + uca_edges = uca_init of the call
        total = self._ring_down(t, T.F_UCA)
        edges = _RingArray(t.shape, self.ring_w, np.float64)
        edges.set_strips([a - b for a, b in zip(total.strips(), t.uca.strips())])
        t.uca_edges = edges
        t.edge_todo = self._ring_down(t, T.F_EDGE_TODO, True); t.edge_done = self._ring_down(t, T.F_EDGE_DONE, True)

    def _twi(self, t):
        kw = {k: self.dem_proc_kwargs[k] for k in ("twi_min_slope", "uca_saturation_limit", "apply_twi_limits", "apply_twi_limits_on_uca")
              if k in self.dem_proc_kwargs}
        kw = {k: (int(v) if isinstance(v, (bool, np.bool_)) else v) for k, v in kw.items()}
        # a fresh DEMProcessor(uca=...) of the base class carries twi_min_area = inf unless the caller set it
        dt = self._tile(t)
        if id(t) in self._mag0:
            dt.as_torch(self._T.F_MAG).copy_(self._mag0.pop(id(t)))
        dt.twi(twi_min_area=float(self.dem_proc_kwargs.get("twi_min_area", np.inf)), **kw)
        t.twi = True                                # resident; read with full(t, "twi")

    def _sync(self, fields, index=None):
        """rings of own tiles are rings already"""
        if self.group is None:
            return
        index = range(self.n_inputs) if index is None else index
        mine = {}
        for i in index:
            t = self.tiles[i]
            if self._mine(t):
                mine[i] = {f: getattr(t, f).strips() for f in fields}
        for part in self.group.all_gather(mine):
            for i, fs in part.items():
                t = self.tiles[i]
                if self._mine(t):
                    continue
                for f, st in fs.items():
                    ra = getattr(t, f)
                    if not isinstance(ra, _RingArray):
                        ra = _RingArray(t.shape, self.ring_w, st[0].dtype)
                        setattr(t, f, ra)
                    ra.set_strips(st)

    def mosaic(self, key):
        R = max(t.box[1] for t in self.tiles); C = max(t.box[3] for t in self.tiles)
        pieces = []
        for t in self.tiles:
            if not self._mine(t):
                continue
            a = self.full(t, key)                   # the resident UCA already includes the edge corrections
            r_lo, r_hi, c_lo, c_hi = t.trim
            nr, nc = t.shape
            pieces.append(((t.box[0] + r_lo, t.box[1] - r_hi, t.box[2] + c_lo, t.box[3] - c_hi), a[r_lo:nr - r_hi, c_lo:nc - c_hi]))
        if self.group is not None:
            pieces = [p for part in self.group.all_gather(pieces) for p in part]
        out = np.full((R, C), np.nan)
        for (r0, r1, c0, c1), a in pieces:
            out[r0:r1, c0:c1] = a
        return out

    def close(self):
        for dt in self._dev.values():
            dt.close()
        self._dev = {}
