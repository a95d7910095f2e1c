"""Generate tests/golden/ref_pm.npz: the UNMODIFIED reference ``ProcessManager.process_twi()`` +
``save_non_overlap_data()`` (process_manager.py:1290-1316, 742-784) run in memory over the stand-in
store of ``oracle/ref_pm_harness.py`` (build container only, where /root/reference exists):

    python tests/golden/make_golden_pm.py

Cases (tests/helpers.pm_cases()): the reference's own five multi-file tilings of the 32x32 cone
(test_end_to_end.py:86-149) and two tilings of a rough 64x64 fractal.  Stored per case: the input
DEM, the tile boxes in processing order, the reference's side-by-side result arrays (elev, aspect,
slope, uca, uca_edges, edge_todo, edge_done, twi) with every tile's block, the compact mosaic,
and the order in which process_uca_edges corrected the tiles.  Before anything is stored the
reference has to meet its own criterion on the cone: mosaic uca[1:-1,1:-1] == single-tile uca.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from oracle import ref_harness as rh, ref_pm_harness as H  # noqa: E402


OVERVIEW_CASES = ("fractal_3x3_2overlap", "fractal_2x3_1overlap", "cone_5x4_3overlap")


def main():
    out = {}
    ref = rh.load_reference()
    for name, (E, nx, ny, ov, kw) in helpers.pm_cases().items():
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = H.run_reference_pm(E, nx, ny, ov, name, dem_proc_kwargs=kw, overviews=[3, 9] if name in OVERVIEW_CASES else None)
            if name.startswith("cone") and ov >= 1:      # the reference's own tilings all overlap
                dp = ref.DEMProcessor(elev=E.copy(), **kw)
                dp.dX[:] = 1; dp.dY[:] = 1; dp.dX2[:] = 1; dp.dY2[:] = 1
                dp.calc_twi()
                np.testing.assert_array_almost_equal(dp.uca[1:-1, 1:-1], r["compact_uca"][1:-1, 1:-1])   # test_end_to_end.py:96
        assert r["success"].all(), name
        out[name + "_E"] = E
        out[name + "_boxes"] = np.array(r["boxes"])
        out[name + "_grid_slice"] = np.array(r["grid_slice"])
        out[name + "_order"] = np.array(r["correction_order"])
        for k in ("elev", "aspect", "slope", "uca", "uca_edges", "edge_todo", "edge_done", "twi", "compact_uca", "compact_twi"):
            out["%s_%s" % (name, k)] = r[k]
        if name in OVERVIEW_CASES:        # process_overviews(out_path_noverlap, keys=[elev, uca], overviews=[3, 9])
            out[name + "_compact_elev"] = r["compact_elev"]
            for k, v in r.items():
                if k.startswith("overview_"):
                    out["%s_%s" % (name, k)] = v
        print(name, "tiles", len(r["boxes"]), "corrections", len(r["correction_order"]))
    # the spacing case: the reference keeps the spacing it derives from the rasters (no DEBUG override)
    name = helpers.PM_SPACING_CASE
    E = helpers.pm_cases()["fractal_3x3_2overlap"][0]
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = H.run_reference_pm(E, 3, 3, 2, name, debug_spacing=False, **helpers.PM_SPACING_GEO)
    assert r["success"].all(), name
    out[name + "_E"] = E
    out[name + "_boxes"] = np.array(r["boxes"]); out[name + "_grid_slice"] = np.array(r["grid_slice"]); out[name + "_order"] = np.array(r["correction_order"])
    for k in ("elev", "aspect", "slope", "uca", "uca_edges", "edge_todo", "edge_done", "twi", "compact_uca", "compact_twi"):
        out["%s_%s" % (name, k)] = r[k]
    print(name, "tiles", len(r["boxes"]), "corrections", len(r["correction_order"]))
    np.savez_compressed(os.path.join(HERE, "ref_pm.npz"), **out)
    print("wrote ref_pm.npz, %.0f kB" % (os.path.getsize(os.path.join(HERE, "ref_pm.npz")) / 1e3))


if __name__ == "__main__":
    main()
