"""GeoTIFF in, spacing out -- the file side of ``DEMProcessor(elev_fn)`` without rasterio / geopy.

Mirrors ``pydem.utils.dem_processor_from_raster_kwargs`` (utils.py:46-51) and
``mk_dx_dy_from_geotif_layer`` (utils.py:127-174): elevation of band 1, ``bounds``, ``transform`` and
the per-row cell spacings ``dX, dY`` (length rows-1, between cell centres) and ``dX2, dY2`` (length
rows).  rasterio and geopy are not installed in this environment, so

* the reader is a small baseline-TIFF parser: uncompressed strips or tiles, one band, little or
  big endian, 8/16/32/64-bit integer or float samples, GeoTIFF tags ModelPixelScale /
  ModelTiepoint / GeoKeyDirectory -- what the reference's own fixtures and ``utils.save_raster``
  outputs use.  Anything else raises ``ValueError`` (use rasterio and pass ``elev=`` etc.);
* geodesic distances (geographic CRS) come from Vincenty's inverse formula on the file's
  ellipsoid instead of geopy's Karney geodesic.  PARITY UNPINNED for this piece: geopy is absent,
  so the two cannot be compared here; both are sub-millimetre accurate for the ~10..100 m
  distances involved (checked against closed forms on the equator and along a meridian).
  Projected CRS: spacing = pixel size, exactly as the reference.
"""
import math
import struct

import numpy as np

_TYPES = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d", 16: "Q", 17: "q"}
ELLIPSOIDS = {               # semi-major axis [km], inverse flattening (geopy.distance.ELLIPSOIDS)
    "WGS-84": (6378.137, 298.257223563),
    "GRS-80": (6378.137, 298.257222101),
    "Airy (1830)": (6377.563396, 299.3249646),
    "Intl 1924": (6378.388, 297.0),
    "Clarke (1880)": (6378.249145, 293.465),
    "GRS-67": (6378.1600, 298.25),
}


class Affine(tuple):
    """(a, b, c, d, e, f): x = a*col + b*row + c, y = d*col + e*row + f (rasterio's order)."""
    a = property(lambda s: s[0]); b = property(lambda s: s[1]); c = property(lambda s: s[2])
    d = property(lambda s: s[3]); e = property(lambda s: s[4]); f = property(lambda s: s[5])


def _ifd(buf, bo, off):
    n, = struct.unpack_from(bo + "H", buf, off)
    tags = {}
    for k in range(n):
        tag, typ, cnt, raw = struct.unpack_from(bo + "HHI4s", buf, off + 2 + 12 * k)
        fmt = _TYPES.get(typ)
        if fmt is None:
            continue
        per = struct.calcsize("=" + fmt)
        size = per * cnt
        src = raw if size <= 4 else buf[struct.unpack(bo + "I", raw)[0]:][:size]
        if typ == 2:
            tags[tag] = bytes(src[:cnt]).split(b"\0")[0].decode("latin1")
        else:
            vals = struct.unpack(bo + fmt * cnt, bytes(src[:size]))
            if typ in (5, 10):
                vals = tuple(vals[i] / vals[i + 1] if vals[i + 1] else 0.0 for i in range(0, len(vals), 2))
            tags[tag] = vals
    return tags


def read_geotiff(fn, projected=None):
    """-> dict(elev, transform, bounds, is_projected, ellipsoid).  ``projected``: what to assume when a
    georeferenced file carries no GTModelTypeGeoKey (None: raise)."""
    buf = open(fn, "rb").read()
    if buf[:2] == b"II":
        bo = "<"
    elif buf[:2] == b"MM":
        bo = ">"
    else:
        raise ValueError("%s is not a TIFF file" % fn)
    magic, off = struct.unpack_from(bo + "HI", buf, 2)
    if magic != 42:
        raise ValueError("%s: BigTIFF / unknown TIFF flavour (magic %d) is not supported here" % (fn, magic))
    t = _ifd(buf, bo, off)
    W, H = t[256][0], t[257][0]
    bits = t.get(258, (1,))[0]
    if t.get(259, (1,))[0] != 1:
        raise ValueError("%s: compressed TIFF (compression %d) is not supported here; use rasterio" % (fn, t[259][0]))
    if t.get(277, (1,))[0] != 1:
        raise ValueError("%s: %d bands; only single-band rasters are supported here" % (fn, t[277][0]))
    sf = t.get(339, (1,))[0]
    kind = {1: "u", 2: "i", 3: "f"}.get(sf)
    if kind is None or bits not in (8, 16, 32, 64) or (kind == "f" and bits < 32):
        raise ValueError("%s: unsupported sample format %d with %d bits" % (fn, sf, bits))
    dt = np.dtype(bo + kind + str(bits // 8))
    data = np.empty((H, W), dt)
    if 324 in t:                                    # tiles
        tw, th = t[322][0], t[323][0]
        offs, k = t[324], 0
        for i in range(0, H, th):
            for j in range(0, W, tw):
                tile = np.frombuffer(buf, dt, tw * th, offs[k]).reshape(th, tw)
                data[i:i + th, j:j + tw] = tile[:min(th, H - i), :min(tw, W - j)]
                k += 1
    else:                                           # strips
        rps = min(t.get(278, (H,))[0], H)
        offs = t[273]
        for k, i in enumerate(range(0, H, rps)):
            rows = min(rps, H - i)
            data[i:i + rows] = np.frombuffer(buf, dt, rows * W, offs[k]).reshape(rows, W)
    # georeferencing: pixel scale + tiepoint (raster (i, j, k) -> model (x, y, z))
    if 33550 in t and 33922 in t:
        sx, sy = t[33550][0], t[33550][1]
        i0, j0, _, x0, y0, _ = t[33922][:6]
        transform = Affine((sx, 0.0, x0 - i0 * sx, 0.0, -sy, y0 + j0 * sy))
    elif 34264 in t:                                # ModelTransformation 4x4
        m = t[34264]
        transform = Affine((m[0], m[1], m[3], m[4], m[5], m[7]))
    else:
        transform = Affine((1.0, 0.0, 0.0, 0.0, -1.0, 0.0))
    keys = t.get(34735, ())
    geokeys = {keys[i]: keys[i + 3] for i in range(4, len(keys) - 3, 4) if keys[i + 1] == 0}
    # The reference asks rasterio for crs.is_projected and for the spheroid named in the WKT and fails when the file has no
    # CRS (utils.py:132-151).  Here the same two facts come from the geokeys; anything that cannot be determined
    # raises instead of silently producing geodesic spacings for metre coordinates (or the wrong ellipsoid).
    georef = 33550 in t or 34264 in t
    model = geokeys.get(1024)
    if model is None and georef and projected is not None:
        model = 1 if projected else 2
        geokeys.setdefault(2048, 4326)
    if model is None:
        if georef:
            raise ValueError("%s: georeferenced TIFF without a GTModelTypeGeoKey: cannot tell projected from geographic "
                             "coordinates (pass projected=True/False)" % fn)
        model = 1                                   # no georeferencing at all: pixel coordinates (unit spacing)
    if model not in (1, 2):
        raise ValueError("%s: GTModelTypeGeoKey %d (geocentric / user defined) is not supported" % (fn, model))
    ellipsoid = "WGS-84"
    if model == 2:
        known = {4326: "WGS-84", 4269: "GRS-80", 4258: "GRS-80", 4277: "Airy (1830)"}
        gcs = geokeys.get(2048)
        if gcs not in known:
            raise ValueError("%s: geographic CRS code %r: ellipsoid unknown (supported: EPSG %s)"
                             % (fn, gcs, ", ".join(str(k) for k in sorted(known))))
        ellipsoid = known[gcs]
    left, top = transform.c, transform.f
    bounds = (left, top + H * transform.e, left + W * transform.a, top)          # left, bottom, right, top
    return dict(elev=data.astype(data.dtype.newbyteorder("=")), transform=transform, bounds=bounds,
                is_projected=(model == 1), ellipsoid=ellipsoid)


def vincenty_km(p1, p2, ellipsoid="WGS-84"):
    """Geodesic distance [km] between (lat, lon) points in degrees (Vincenty 1975, inverse problem)."""
    a, rf = ELLIPSOIDS[ellipsoid]
    f = 1.0 / rf
    b = a * (1.0 - f)
    (la1, lo1), (la2, lo2) = p1, p2
    U1 = math.atan((1 - f) * math.tan(math.radians(la1))); U2 = math.atan((1 - f) * math.tan(math.radians(la2)))
    L = math.radians(lo2 - lo1)
    sU1, cU1, sU2, cU2 = math.sin(U1), math.cos(U1), math.sin(U2), math.cos(U2)
    lam = L
    for _ in range(200):
        sl, cl = math.sin(lam), math.cos(lam)
        ss = math.hypot(cU2 * sl, cU1 * sU2 - sU1 * cU2 * cl)
        if ss == 0.0:
            return 0.0
        cs = sU1 * sU2 + cU1 * cU2 * cl
        sig = math.atan2(ss, cs)
        sa = cU1 * cU2 * sl / ss
        c2a = 1.0 - sa * sa
        c2sm = cs - 2.0 * sU1 * sU2 / c2a if c2a else 0.0
        Cc = f / 16.0 * c2a * (4.0 + f * (4.0 - 3.0 * c2a))
        new = L + (1.0 - Cc) * f * sa * (sig + Cc * ss * (c2sm + Cc * cs * (-1.0 + 2.0 * c2sm * c2sm)))
        if abs(new - lam) < 1e-14:
            lam = new
            break
        lam = new
    u2 = c2a * (a * a - b * b) / (b * b)
    A = 1.0 + u2 / 16384.0 * (4096.0 + u2 * (-768.0 + u2 * (320.0 - 175.0 * u2)))
    B = u2 / 1024.0 * (256.0 + u2 * (-128.0 + u2 * (74.0 - 47.0 * u2)))
    ds = B * ss * (c2sm + B / 4.0 * (cs * (-1.0 + 2.0 * c2sm * c2sm) - B / 6.0 * c2sm * (-3.0 + 4.0 * ss * ss) * (-3.0 + 4.0 * c2sm * c2sm)))
    return b * A * (sig - ds)


def mk_dx_dy(transform, nrows, is_projected, ellipsoid="WGS-84"):
    """utils.mk_dx_dy_from_geotif_layer (127-174) -> dX, dY (nrows-1), dX2, dY2 (nrows), metres."""
    if is_projected:
        dX = np.ones(nrows - 1) * transform.a; dX2 = np.ones(nrows) * transform.a
        dY = np.abs(np.ones(nrows - 1) * transform.e); dY2 = np.abs(np.ones(nrows) * transform.e)
        return dX, dY, dX2, dY2
    clip = lambda v: min(90.0, max(-90.0, v))
    m = lambda p, q: vincenty_km(p, q, ellipsoid) * 1000.0
    dx, dy = transform.a, transform.e
    lon = transform.d + dx / 2          # as in the reference (it reads the rotation term, utils.py:154): only lat matters
    lat = transform.f + dy / 2
    dX = np.array([m((clip(lat + dy * (j + 1)), lon + dx), (clip(lat + dy * (j + 1)), lon)) for j in range(nrows - 1)])
    dY = np.array([m((clip(lat + dy * i), lon), (clip(lat + dy * (i + 1)), lon)) for i in range(nrows - 1)])
    lon = transform.d + dx
    lat = transform.f + dy
    dX2 = np.array([m((clip(lat + dy * (j + 1)), lon + dx), (clip(lat + dy * (j + 1)), lon)) for j in range(nrows)])
    dY2 = np.array([m((clip(lat + dy * i), lon), (clip(lat + dy * (i + 1)), lon)) for i in range(nrows)])
    return dX, dY, dX2, dY2


def dem_processor_from_raster_kwargs(fn):
    """utils.dem_processor_from_raster_kwargs (46-51)."""
    r = read_geotiff(fn)
    dX, dY, dX2, dY2 = mk_dx_dy(r["transform"], r["elev"].shape[0], r["is_projected"], r["ellipsoid"])
    return dict(dX=dX, dY=dY, elev=r["elev"], bounds=r["bounds"], transform=r["transform"], dX2=dX2, dY2=dY2)


def write_geotiff(fn, data, transform, projected=False, tile=None, gcs=4326, model=None):
    """Minimal writer (uncompressed, one band, little endian; strips or `tile` x `tile` tiles) --
    what utils.save_raster produces for the reference's tests.  Used by this repo's tests."""
    data = np.ascontiguousarray(data)
    H, W = data.shape
    kind = {"f": 3, "i": 2, "u": 1}[data.dtype.kind]
    entries = []            # (tag, type, count, payload bytes)
    def ent(tag, typ, vals):
        fmt = _TYPES[typ]
        entries.append((tag, typ, len(vals), struct.pack("<" + fmt * len(vals), *vals)))
    ent(256, 4, [W]); ent(257, 4, [H]); ent(258, 3, [data.dtype.itemsize * 8]); ent(259, 3, [1]); ent(262, 3, [1])
    ent(277, 3, [1]); ent(284, 3, [1]); ent(339, 3, [kind])
    if tile:
        blocks = []
        for i in range(0, H, tile):
            for j in range(0, W, tile):
                blk = np.zeros((tile, tile), data.dtype)
                part = data[i:i + tile, j:j + tile]
                blk[:part.shape[0], :part.shape[1]] = part
                blocks.append(blk.astype("<" + data.dtype.str[1:]).tobytes())
        ent(322, 4, [tile]); ent(323, 4, [tile])
        off_tag, cnt_tag = 324, 325
    else:
        blocks = [data.astype("<" + data.dtype.str[1:]).tobytes()]
        ent(278, 4, [H])
        off_tag, cnt_tag = 273, 279
    ent(33550, 12, [transform.a, -transform.e, 0.0])
    ent(33922, 12, [0.0, 0.0, 0.0, transform.c, transform.f, 0.0])
    # (gcs / model: tests write files whose CRS the reader must refuse)
    ent(34735, 3, [1, 1, 0, 2, 1024, 0, 1, (1 if projected else 2) if model is None else model, 2048, 0, 1, gcs])
    n = len(entries) + 2
    head = 8 + 2 + 12 * n + 4
    extra = b""
    offs_of = {}
    for tag, typ, cnt, payload in entries:
        if len(payload) > 4:
            offs_of[tag] = head + len(extra)
            extra += payload + (b"\0" if len(payload) % 2 else b"")
    data_off = head + len(extra) + 4 * len(blocks) * 2
    boffs, pos = [], data_off
    for b in blocks:
        boffs.append(pos); pos += len(b)
    tables = struct.pack("<%dI" % len(blocks), *boffs) + struct.pack("<%dI" % len(blocks), *[len(b) for b in blocks])
    t_off = head + len(extra)
    if len(blocks) == 1:
        entries.append((off_tag, 4, 1, struct.pack("<I", boffs[0]))); entries.append((cnt_tag, 4, 1, struct.pack("<I", len(blocks[0]))))
    else:
        entries.append((off_tag, 4, len(blocks), None)); offs_of[off_tag] = t_off
        entries.append((cnt_tag, 4, len(blocks), None)); offs_of[cnt_tag] = t_off + 4 * len(blocks)
    entries.sort(key=lambda e: e[0])
    out = b"II" + struct.pack("<HI", 42, 8) + struct.pack("<H", len(entries))
    for tag, typ, cnt, payload in entries:
        if tag in offs_of:
            val = struct.pack("<I", offs_of[tag])
        else:
            val = payload + b"\0" * (4 - len(payload))
        out += struct.pack("<HHI", tag, typ, cnt) + val
    out += struct.pack("<I", 0) + extra + tables + b"".join(blocks)
    with open(fn, "wb") as f:
        f.write(out)
