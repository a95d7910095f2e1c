"""Build the reference's only native component into ``oracle/_ref/`` (git-ignored).

``/root/reference/pydem/cyfuncs/cyutils.pyx`` is cythonized *from where it lies* (no
reference source is copied into the repo; only the generated C++ and the compiled
extension land in ``oracle/_ref/``), with the reference's own flags
(reference ``setup.py:27-44``: C++, ``-O3 -march=x86-64``).
The resulting ``cyutils*.so`` travels to the GPU box with the snapshot and is what
``bench.py --impl reference`` / ``cpu_baseline`` time for the UCA sweep.
"""
import os
import subprocess
import sys
import sysconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(_HERE, "_ref")
PYX = os.path.join(os.environ.get("PYDEM_REFERENCE_ROOT", "/root/reference"),
                   "pydem", "cyfuncs", "cyutils.pyx")


def build(force=False):
    import numpy as np
    os.makedirs(OUT, exist_ok=True)
    so = os.path.join(OUT, "cyutils" + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(so) and not force:
        return so
    if not os.path.isfile(PYX):
        raise RuntimeError("reference cyutils.pyx not found: %s" % PYX)
    cpp = os.path.join(OUT, "cyutils.cpp")
    subprocess.check_call([sys.executable, "-m", "cython", "--cplus", "-3", PYX, "-o", cpp])
    inc = sysconfig.get_paths()["include"]
    cmd = ["g++", "-O3", "-march=x86-64", "-shared", "-fPIC", "-w",
           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
           "-I", inc, "-I", np.get_include(), cpp, "-o", so]
    subprocess.check_call(cmd)
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
