"""Build ``libpydem_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pydem_b200.build [--force]

The shared library is a plain C-ABI object (include/pydem_b200.h): no torch, no pybind.
``-fmad=false`` keeps every double-precision expression in the reference's IEEE operation
order (NumPy never fuses a*b+c), which is what makes ``mag`` and the integer outputs
bit-identical to the reference.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libpydem_b200.so")
SOURCES = ["slopes.cu", "flats.cu", "graph.cu", "pits.cu", "sweep.cu", "tsweep.cu", "update.cu", "shard.cu", "comm.cu", "condition.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false",
              "-std=c++17", "-Xcompiler", "-fPIC", "-cudart", "static"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    mt = os.path.getmtime(target)
    return any(os.path.getmtime(d) > mt for d in deps)


def build(force=False, verbose=False):
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(_HERE, "..", "include", "pydem_b200.h"))
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [_nvcc()] + NVCC_FLAGS + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out.decode(errors="replace")))
    if force or procs or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-cudart", "static", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    from . import synth
    synth._host_lib()       # synthetic-data helper (gcc)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
