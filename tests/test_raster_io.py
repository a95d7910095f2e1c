"""File side of DEMProcessor(elev_fn): the baseline-GeoTIFF reader and the spacing arithmetic of
pydem_b200/raster_io.py (reference utils.py:46-51, 127-174)."""
import os

import numpy as np
import pytest

from oracle import ref_harness
from pydem_b200 import raster_io as rio

FIXTURE = os.path.join(ref_harness.REFERENCE_ROOT, "pydem", "test", "test_NN032_033_elev.tif")


@pytest.mark.parametrize("dtype", ["float64", "float32", "int16", "uint8"])
@pytest.mark.parametrize("tile", [None, 16])
def test_write_read_round_trip(tmp_path, dtype, tile):
    rng = np.random.default_rng(1)
    a = (rng.random((37, 53)) * 200).astype(dtype)
    tr = rio.Affine((0.25, 0.0, -73.0, 0.0, -0.5, 46.0))
    fn = str(tmp_path / "t.tif")
    rio.write_geotiff(fn, a, tr, projected=(tile is None), tile=tile)
    r = rio.read_geotiff(fn)
    np.testing.assert_array_equal(r["elev"], a)
    assert r["elev"].dtype == np.dtype(dtype)
    assert tuple(r["transform"]) == tuple(tr)
    assert r["bounds"] == (-73.0, 46.0 - 37 * 0.5, -73.0 + 53 * 0.25, 46.0)
    assert r["is_projected"] == (tile is None)


def test_projected_spacing_is_the_pixel_size(tmp_path):
    fn = str(tmp_path / "p.tif")
    rio.write_geotiff(fn, np.ones((5, 4)), rio.Affine((30.0, 0.0, 5e5, 0.0, -25.0, 4e6)), projected=True)
    kw = rio.dem_processor_from_raster_kwargs(fn)
    np.testing.assert_array_equal(kw["dX"], np.full(4, 30.0)); np.testing.assert_array_equal(kw["dY"], np.full(4, 25.0))
    np.testing.assert_array_equal(kw["dX2"], np.full(5, 30.0)); np.testing.assert_array_equal(kw["dY2"], np.full(5, 25.0))


def test_vincenty_against_closed_forms():
    a, rf = rio.ELLIPSOIDS["WGS-84"]
    f = 1 / rf; e2 = f * (2 - f)
    assert abs(rio.vincenty_km((0, 10), (0, 11)) - a * np.pi / 180) < 1e-9            # along the equator: a * dlon
    phi = np.linspace(np.radians(45), np.radians(45.03), 200001)                      # meridian arc by quadrature
    M = a * (1 - e2) / (1 - e2 * np.sin(phi) ** 2) ** 1.5
    assert abs(rio.vincenty_km((45, -73), (45.03, -73)) - np.trapezoid(M, phi)) < 1e-9
    N = a / np.sqrt(1 - e2 * np.sin(np.radians(45)) ** 2)                              # short east-west step: N cos(lat) dlon
    assert abs(rio.vincenty_km((45, -73), (45, -73 + 1e-3)) - N * np.cos(np.radians(45)) * np.radians(1e-3)) < 1e-8
    assert rio.vincenty_km((12.5, 7.0), (12.5, 7.0)) == 0.0


def test_geographic_spacing_shapes_and_monotony(tmp_path):
    fn = str(tmp_path / "g.tif")
    rio.write_geotiff(fn, np.ones((6, 3)), rio.Affine((1 / 31.0, 0.0, -73.0, 0.0, -1 / 31.0, 46.0)), projected=False)
    kw = rio.dem_processor_from_raster_kwargs(fn)
    assert kw["dX"].shape == (5,) and kw["dY"].shape == (5,) and kw["dX2"].shape == (6,) and kw["dY2"].shape == (6,)
    assert np.all(np.diff(kw["dX"]) > 0)             # rows go south from 46 N: parallels get longer
    assert np.all(np.abs(kw["dY"] - 3585.5) < 1.0)   # ~1/31 degree of latitude in metres


def test_rejects_what_it_cannot_read(tmp_path):
    fn = str(tmp_path / "x.tif")
    open(fn, "wb").write(b"not a tiff")
    with pytest.raises(ValueError):
        rio.read_geotiff(fn)


@pytest.mark.skipif(not os.path.isfile(FIXTURE), reason="reference tree not present (GPU box)")
def test_reads_the_reference_fixture():
    """pydem/test/test_NN032_033_elev.tif is the 32x32 cone of utils_test_pydem.py:98-103, written by the
    reference with the pixel-centred WGS84 layout of utils.mk_geotiff_obj (:178-206)."""
    r = rio.read_geotiff(FIXTURE)
    x, y = np.mgrid[-1:1:32j, -1:1:32j]
    np.testing.assert_array_equal(r["elev"], 1 - np.sqrt(x ** 2 + y ** 2) / np.sqrt(2))
    ph = pw = 1.0 / 31.0
    np.testing.assert_allclose(r["transform"], (pw, 0, -73 - pw / 2, 0, -ph, 46 + ph / 2), rtol=0, atol=1e-12)
    assert r["is_projected"] is False and r["ellipsoid"] == "WGS-84"
    # and the operator accepts the file name, like the reference's DEMProcessor(elev_fn)
    from pydem_b200 import DEMProcessor
    dp = DEMProcessor(FIXTURE)
    assert dp.elev.shape == (32, 32) and dp.dX.shape == (31,) and dp.dY2.shape == (32,)


def test_orchestrator_from_a_directory_of_geotiffs(tmp_path):
    """ProcessManager.from_directory on tiles written the way the reference's multi-file generator lays
    them out (utils_test_pydem.mk_test_multifile :359-408: pixel-centred lat/lon per chunk) reproduces
    the reference ProcessManager's result for the same tiling (golden fixture), scheduling included."""
    import contextlib, io, warnings
    import helpers
    from oracle.oracle import OracleDEMProcessor
    from pydem_b200.process_manager import ProcessManager
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pm.npz"))
    name = "cone_5x4_3overlap"
    E = G[name + "_E"]
    boxes = [tuple(b) for b in G[name + "_boxes"].tolist()]
    la = np.linspace(46.0, 45.0, E.shape[0]); lo = np.linspace(-73.0, -72.0, E.shape[1])
    for (te, be, le, re) in boxes:
        ph = -abs(la[te] - la[be - 1]) / ((be - te) - 1.0); pw = abs(lo[le] - lo[re - 1]) / ((re - le) - 1.0)
        tr = rio.Affine((pw, 0.0, lo[le] - pw / 2, 0.0, ph, la[te] - ph / 2))
        rio.write_geotiff(str(tmp_path / ("tile_%04d_%04d.tif" % (te, le))), E[te:be, le:re], tr)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        pm = ProcessManager.from_directory(str(tmp_path), unit_spacing=True,
                                           dem_processor=lambda **k: OracleDEMProcessor(k.pop("elev"), **k))
        assert [t.box for t in pm.tiles] == boxes
        pm.process_twi()
    worst = helpers.pm_compare(pm, G, name)
    assert worst["elev"] == 0 and worst["edge_todo"] == 0 and worst["edge_done"] == 0 and worst["uca"] <= 1e-9, worst
    assert pm.correction_log == G[name + "_order"].tolist()
    # and back to a file: the mosaic with the transform of save_geotiff (:889-895)
    out = str(tmp_path / "uca_mosaic.tif")
    pm.save_geotiff(out, "uca")
    r = rio.read_geotiff(out)
    np.testing.assert_array_equal(r["elev"], pm.mosaic("uca"))
    pw = 1.0 / 31.0
    np.testing.assert_allclose(r["transform"], ((-72 + pw / 2 - (-73 - pw / 2)) / 32, 0, -73 - pw / 2, 0, -((46 + pw / 2) - (45 - pw / 2)) / 32, 46 + pw / 2),
                               rtol=0, atol=1e-12)


def test_big_endian_multi_strip_tiff(tmp_path):
    """A hand-built big-endian TIFF with 3-row strips and a ModelTransformation tag instead of
    pixel scale + tiepoint: the reader's other code paths."""
    import struct
    H, W, rps = 7, 5, 3
    a = (np.arange(H * W, dtype=">f4").reshape(H, W) * 1.5).astype(">f4")
    strips = [a[i:i + rps].tobytes() for i in range(0, H, rps)]
    n_tags = 10
    ifd_off = 8
    extra_off = ifd_off + 2 + 12 * n_tags + 4
    offs_tbl = extra_off
    cnts_tbl = offs_tbl + 4 * len(strips)
    mt_off = cnts_tbl + 4 * len(strips)
    data_off = mt_off + 16 * 8
    soffs, pos = [], data_off
    for st in strips:
        soffs.append(pos); pos += len(st)
    M = [2.0, 0, 0, 100.0, 0, -3.0, 0, 50.0, 0, 0, 0, 0, 0, 0, 0, 1]
    def ent(tag, typ, cnt, val):
        return struct.pack(">HHI", tag, typ, cnt) + val
    body = b"MM" + struct.pack(">HI", 42, ifd_off) + struct.pack(">H", n_tags)
    body += ent(256, 3, 1, struct.pack(">HH", W, 0)) + ent(257, 3, 1, struct.pack(">HH", H, 0)) + ent(258, 3, 1, struct.pack(">HH", 32, 0))
    body += ent(259, 3, 1, struct.pack(">HH", 1, 0)) + ent(273, 4, len(strips), struct.pack(">I", offs_tbl)) + ent(277, 3, 1, struct.pack(">HH", 1, 0))
    body += ent(278, 3, 1, struct.pack(">HH", rps, 0)) + ent(279, 4, len(strips), struct.pack(">I", cnts_tbl)) + ent(339, 3, 1, struct.pack(">HH", 3, 0))
    body += ent(34264, 12, 16, struct.pack(">I", mt_off)) + struct.pack(">I", 0)
    body += struct.pack(">%dI" % len(strips), *soffs) + struct.pack(">%dI" % len(strips), *[len(st) for st in strips])
    body += struct.pack(">16d", *M) + b"".join(strips)
    fn = str(tmp_path / "be.tif")
    open(fn, "wb").write(body)
    with pytest.raises(ValueError, match="GTModelTypeGeoKey"):
        rio.read_geotiff(fn)        # georeferenced, but nothing says whether the coordinates are metres or degrees
    r = rio.read_geotiff(fn, projected=True)
    assert r["is_projected"]
    np.testing.assert_array_equal(r["elev"], a.astype("f4"))
    assert tuple(r["transform"]) == (2.0, 0.0, 100.0, 0.0, -3.0, 50.0)
    assert r["bounds"] == (100.0, 50.0 - 3.0 * H, 100.0 + 2.0 * W, 50.0)


def test_from_directory_keeps_the_rasters_spacing(tmp_path):
    """Projected rasters, no unit-spacing override: every tile's dX / dY are its own pixel size
    (utils.py:132-137) -- the golden run of the reference ProcessManager without its DEBUG switch."""
    import contextlib, io, warnings
    import helpers
    from oracle.oracle import OracleDEMProcessor
    from pydem_b200.process_manager import ProcessManager
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_pm.npz"))
    name = helpers.PM_SPACING_CASE
    E = G[name + "_E"]
    boxes = [tuple(b) for b in G[name + "_boxes"].tolist()]
    la = np.linspace(helpers.PM_SPACING_GEO["lat"][0], helpers.PM_SPACING_GEO["lat"][1], E.shape[0])
    lo = np.linspace(helpers.PM_SPACING_GEO["lon"][0], helpers.PM_SPACING_GEO["lon"][1], E.shape[1])
    for (te, be, le, re) in boxes:
        ph = -abs(la[te] - la[be - 1]) / ((be - te) - 1.0); pw = abs(lo[le] - lo[re - 1]) / ((re - le) - 1.0)
        rio.write_geotiff(str(tmp_path / ("tile_%04d_%04d.tif" % (te, le))), E[te:be, le:re],
                          rio.Affine((pw, 0.0, lo[le] - pw / 2, 0.0, ph, la[te] - ph / 2)), projected=True)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        pm = ProcessManager.from_directory(str(tmp_path), dem_processor=lambda **k: OracleDEMProcessor(k.pop("elev"), **k))
        assert [t.box for t in pm.tiles] == boxes
        pm.process_twi()
    worst = helpers.pm_compare(pm, G, name)
    assert worst["elev"] == 0 and worst["edge_todo"] == 0 and worst["edge_done"] == 0, worst
    assert worst["slope"] <= 1e-13 * 1e6 and worst["aspect"] <= 1e-12 and worst["uca"] <= 1e-9, worst
    assert pm.correction_log == G[name + "_order"].tolist()


def test_undeterminable_crs_is_refused(tmp_path):
    """The reader must not guess: a geographic file whose ellipsoid is unknown, or a model type that is neither
    projected nor geographic, raises instead of silently producing WGS-84 geodesic spacings (the reference fails too
    when rasterio cannot give it `crs.is_projected` / a spheroid name, utils.py:132-151)."""
    a = np.arange(12, dtype="f4").reshape(3, 4)
    tr = rio.Affine((0.01, 0.0, -72.0, 0.0, -0.01, 45.0))
    ok = str(tmp_path / "ok.tif"); rio.write_geotiff(ok, a, tr)
    assert rio.read_geotiff(ok)["ellipsoid"] == "WGS-84" and not rio.read_geotiff(ok)["is_projected"]
    nad83 = str(tmp_path / "nad83.tif"); rio.write_geotiff(nad83, a, tr, gcs=4269)
    assert rio.read_geotiff(nad83)["ellipsoid"] == "GRS-80"
    bad_gcs = str(tmp_path / "gcs.tif"); rio.write_geotiff(bad_gcs, a, tr, gcs=4267)           # NAD27: Clarke 1866, not in the table
    with pytest.raises(ValueError, match="ellipsoid unknown"):
        rio.read_geotiff(bad_gcs)
    geocentric = str(tmp_path / "geoc.tif"); rio.write_geotiff(geocentric, a, tr, model=3)
    with pytest.raises(ValueError, match="not supported"):
        rio.read_geotiff(geocentric)
    projected = str(tmp_path / "proj.tif"); rio.write_geotiff(projected, a, rio.Affine((30.0, 0.0, 5e5, 0.0, -30.0, 4e6)), projected=True, gcs=4267)
    assert rio.read_geotiff(projected)["is_projected"]            # metres: the ellipsoid does not matter
