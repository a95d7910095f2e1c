"""Load the UNMODIFIED reference (``/root/reference/pydem``) in this container.

Test infrastructure (see ``oracle/__init__.py``).  The reference needs a few pure
Python packages that are not installed here (traitlets, traittypes, rasterio, zarr,
geopy); none of them does arithmetic on the hot path, so small stand-ins are
registered in ``sys.modules`` before the reference is imported *from where it lies*
(nothing is copied into this repo).  The one native piece of the reference,
``pydem/cyfuncs/cyutils.pyx``, is cythonized+compiled by ``oracle/build_ref.py`` into the
git-ignored ``oracle/_ref/`` and injected as ``pydem.cyfuncs.cyutils``.

``/root/reference`` does not exist on the GPU box: this module is only used here, to
validate the restatement (``oracle/pdm_oracle.c`` + ``oracle/oracle.py``) and to generate
``tests/golden/*.npz`` (script: ``tests/golden/make_golden.py``).
"""
import importlib
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("PYDEM_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_BUILD_DIR = os.path.join(_HERE, "_ref")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pydem", "dem_processing.py"))


# ----------------------------------------------------------------------------
# stand-ins for traitlets / traittypes (configuration containers, no arithmetic)
# ----------------------------------------------------------------------------
_NOTSET = object()


class _Trait(object):
    def __init__(self, default_value=_NOTSET, *args, **kwargs):
        self.default_value = default_value
        self.allow_none = kwargs.get("allow_none", False)
        self.name = None

    def __set_name__(self, owner, name):
        self.name = name

    def make_default(self, obj):
        dv = self.default_value
        return None if dv is _NOTSET else dv

    def coerce(self, value):
        return value

    def __get__(self, obj, objtype=None):
        if obj is None:
            return self
        d = obj.__dict__.setdefault("_trait_values", {})
        if self.name not in d:
            maker = None
            for klass in type(obj).__mro__:
                for attr in vars(klass).values():
                    if getattr(attr, "_default_for", None) == self.name:
                        maker = attr
                        break
                if maker is not None:
                    break
            d[self.name] = self.coerce(maker(obj)) if maker is not None else self.make_default(obj)
        return d[self.name]

    def __set__(self, obj, value):
        obj.__dict__.setdefault("_trait_values", {})[self.name] = self.coerce(value)


class _Container(_Trait):
    def __init__(self, default_value=_NOTSET, *args, **kwargs):
        # tl.List(SomeTrait(), default) form: first arg may be a trait instance
        if isinstance(default_value, _Trait):
            default_value = args[0] if args else kwargs.get("default_value", _NOTSET)
        super().__init__(default_value, **kwargs)

    def make_default(self, obj):
        dv = self.default_value
        if dv is _NOTSET or dv is None:
            return self._empty()
        return type(self._empty())(dv)


class _List(_Container):
    def _empty(self):
        return []


class _Dict(_Container):
    def _empty(self):
        return {}


class _Instance(_Trait):
    def __init__(self, klass=None, default_value=None, *args, **kwargs):
        super().__init__(None, **kwargs)


class _Enum(_Trait):
    def __init__(self, values=None, default_value=_NOTSET, **kwargs):
        super().__init__(default_value, **kwargs)


class _Array(_Trait):
    def make_default(self, obj):
        dv = self.default_value
        return None if dv is _NOTSET or dv is None else np.asarray(dv)

    def coerce(self, value):
        return None if value is None else np.asarray(value)


class _HasTraits(object):
    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)


def _default(name):
    def deco(fn):
        fn._default_for = name
        return fn
    return deco


def _install_shims():
    if "traitlets" not in sys.modules:
        tl = types.ModuleType("traitlets")
        tl.HasTraits = _HasTraits
        for nm in ("Bool", "Int", "Float", "Unicode", "Any", "Tuple"):
            setattr(tl, nm, type(nm, (_Trait,), {}))
        tl.List, tl.Dict, tl.Instance, tl.Enum = _List, _Dict, _Instance, _Enum
        tl.default = _default
        tl.observe = lambda *a, **k: (lambda fn: fn)
        tl.validate = lambda *a, **k: (lambda fn: fn)
        sys.modules["traitlets"] = tl
    if "traittypes" not in sys.modules:
        tt = types.ModuleType("traittypes")
        tt.Array = _Array
        sys.modules["traittypes"] = tt
    for nm in ("rasterio", "zarr", "geopy"):
        if nm not in sys.modules:
            sys.modules[nm] = types.ModuleType(nm)
    if "geopy.distance" not in sys.modules:
        gd = types.ModuleType("geopy.distance")
        gd.distance = None
        sys.modules["geopy.distance"] = gd
        sys.modules["geopy"].distance = gd


def load_ref_cyutils():
    """The reference's own Cython kernels, built by oracle/build_ref.py (travels to the GPU box)."""
    import glob
    hits = sorted(glob.glob(os.path.join(REF_BUILD_DIR, "cyutils*.so")))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("pydem.cyfuncs.cyutils", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_REF = None


def load_reference():
    """Return the reference's ``pydem.dem_processing`` module (imported in place)."""
    global _REF
    if _REF is not None:
        return _REF
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_shims()
    cy = load_ref_cyutils()
    if cy is None:
        from . import build_ref
        build_ref.build()
        cy = load_ref_cyutils()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # stub the top-level package so pydem/__init__.py (which pulls process_manager ->
    # real zarr/rasterio) is not executed; submodules are imported from the real files.
    pkg = types.ModuleType("pydem")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "pydem")]
    sys.modules["pydem"] = pkg
    sys.modules["pydem.cyfuncs.cyutils"] = cy
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _REF = importlib.import_module("pydem.dem_processing")
    assert _REF.CYTHON, "reference cyutils failed to load"
    return _REF


def ref_processor(elev, **kwargs):
    """``DEMProcessor`` of the reference on a private copy of ``elev``."""
    mod = load_reference()
    return mod.DEMProcessor(elev=np.array(elev, dtype="float64", copy=True), **kwargs)
