"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container,
where /root/reference exists):

    python tests/golden/make_golden.py

* ref_known_answers.npz : the reference's own 5x5 known-answer vectors, read from
  /root/reference/pydem/test/test_end_to_end.py:152-182 and 220-251 (elev, ang, mag, uca).
* ref_cases.npz         : inputs and reference outputs (mag, direction, flats, uca, edge_todo,
  edge_done, twi, section, the post-uca mag/flats) for the small cases of tests/helpers.golden_cases().
* ref_update.npz        : a row-tiled edge-update sequence (calc_uca(uca_init, edge_init_data))
  driven through the reference, every call's inputs and outputs.
* ref_conditioning.npz  : reference calc_fill_pit_artifacts / calc_fill_flats /
  calc_pit_drain_paths outputs for tests/helpers.conditioning_cases() (stored as sparse
  differences from the input).  `<case>_paths_tiefree` records whether the reference's pit order
  is free of elevation ties that change the result (np.argsort's tie order is platform-defined).
"""
import contextlib
import importlib.util
import io
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import helpers  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402


def known_answers():
    mod = rh.load_reference()
    pkg = sys.modules["pydem"]
    for nm in ("process_manager",):
        m = types.ModuleType("pydem." + nm); sys.modules["pydem." + nm] = m; setattr(pkg, nm, m)
    import pydem.utils as u, pydem.utils_test_pydem as utp
    pkg.utils = u; pkg.utils_test_pydem = utp; pkg.DEMProcessor = mod.DEMProcessor
    spec = importlib.util.spec_from_file_location("ref_test_e2e", os.path.join(rh.REFERENCE_ROOT, "pydem/test/test_end_to_end.py"))
    t = importlib.util.module_from_spec(spec); spec.loader.exec_module(t)
    out = {}
    for nm, cls in (("cardinal", t.TestEndtoEndCardinal), ("diagonal", t.TestEndtoEndDiagonal)):
        for k in ("elev", "ang", "mag", "uca"):
            out["%s_%s" % (nm, k)] = np.asarray(getattr(cls, k), float)
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cls().test_simple_slope()   # the reference must pass its own test here
    np.savez_compressed(os.path.join(HERE, "ref_known_answers.npz"), **out)


def case_outputs():
    out = {}
    for name, (E, kw) in helpers.golden_cases().items():
        res = helpers.run(lambda e, **k: rh.ref_processor(e, **k), E, kw)
        # section as the reference computes it from the pre-uca state
        k = dict(helpers.HOT); k.update(kw)
        with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            dp = rh.ref_processor(E, **k)
            dp.calc_slopes_directions()
            sec, prop = dp._calc_uca_section_proportion(dp.elev, dp.dX, dp.dY, dp.direction, dp.flats)
        res["section"] = sec.astype(np.int8); res["proportion"] = prop
        out[name + "__elev"] = E
        for kk, v in res.items():
            out[name + "__" + kk] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "ref_cases.npz"), **out)


def update_sequence():
    from tiling import tiled_rows
    E = helpers.synth.fractal_dem(48, 21) * 0.05 + helpers.synth.cone_dem(48) * 300 + 1
    kw = dict(helpers.HOT, drain_pits=False)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        blocks, log, calls = tiled_rows(lambda e, **k: rh.ref_processor(e, **k), E, 3, 2, kw, record=True)
    out = {"elev": E, "n_calls": np.array(len(calls))}
    for n, c in enumerate(calls):
        for kk, v in c.items():
            out["call%d__%s" % (n, kk)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "ref_update.npz"), **out)


def _sparse(out, key, new, old):
    idx = np.flatnonzero(new.ravel() != old.ravel())
    out[key + "_idx"] = idx.astype("int32"); out[key + "_val"] = new.ravel()[idx]


def conditioning():
    from oracle import conditioning as oc
    out = {}
    for nm, (E, dX, dY) in helpers.conditioning_cases().items():
        def run(meth, elev):
            dp = rh.ref_processor(elev, dX=dX, dY=dY)
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                getattr(dp, meth)()
            return np.array(dp.elev)
        art = run("calc_fill_pit_artifacts", E)
        fill = run("calc_fill_flats", E)
        paths = run("calc_pit_drain_paths", fill)
        _sparse(out, nm + "_art", art, E); _sparse(out, nm + "_fill", fill, E); _sparse(out, nm + "_paths", paths, fill)
        mine = oc.pit_drain_paths(fill, dX, dY)[0]
        out[nm + "_paths_tiefree"] = np.array(np.array_equal(mine, paths))
        print(nm, "art", len(out[nm + "_art_idx"]), "fill", len(out[nm + "_fill_idx"]), "paths", len(out[nm + "_paths_idx"]),
              "tiefree", bool(out[nm + "_paths_tiefree"]))
    np.savez_compressed(os.path.join(HERE, "ref_conditioning.npz"), **out)


if __name__ == "__main__":
    known_answers(); case_outputs(); update_sequence(); conditioning()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
