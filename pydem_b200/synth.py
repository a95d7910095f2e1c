"""Synthetic DEMs for tests and benchmarks (no dataset ships with the reference and there is
no network): the analytic cone of the reference's own fixtures and seeded fractal terrain.

reference: pydem/utils_test_pydem.py:98-124 (case_cone), SURVEY.md section 8(d) (configs).
"""
import ctypes as ct
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def cone_dem(n=256):
    """``case_cone`` of the reference's test catalogue (utils_test_pydem.py:98-103, 422):
    a cone that drains radially outward, elevation in [0, 1]."""
    x, y = np.mgrid[-1:1:complex(0, n), -1:1:complex(0, n)]
    return 1.0 - np.sqrt(x ** 2 + y ** 2) / np.sqrt(2.0)


def fractal_dem(n, seed=0, hurst=0.8, lo=1.0, hi=1001.0, shape=None):
    """Spectral-synthesis fractal terrain (random-phase power law ``k^-(H+1)``), rescaled to
    [lo, hi] metres so every cell is above sea level."""
    rows, cols = (n, n) if shape is None else shape
    rng = np.random.default_rng(seed)
    ky = np.fft.fftfreq(rows)[:, None]
    kx = np.fft.fftfreq(cols)[None, :]
    k = np.sqrt(kx * kx + ky * ky)
    k[0, 0] = 1.0
    spec = (rng.standard_normal((rows, cols)) + 1j * rng.standard_normal((rows, cols))) * k ** (-(hurst + 1.0))
    spec[0, 0] = 0.0
    z = np.fft.ifft2(spec).real
    z -= z.min()
    z *= (hi - lo) / z.max()
    z += lo
    return np.ascontiguousarray(z)


def value_noise_dem(row0, nrows, ncols, seed=2, octaves=8, base=1024, lo=1.0, hi=1001.0):
    """Octave-summed seeded lattice noise that can be evaluated for any row block
    independently (the FFT generator cannot be sharded): rows [row0, row0+nrows) of an
    unbounded terrain.  Used for the row-sharded configurations."""
    ii = (np.arange(row0, row0 + nrows, dtype=np.float64))[:, None]
    jj = (np.arange(ncols, dtype=np.float64))[None, :]
    out = np.zeros((nrows, ncols))
    amp, norm = 1.0, 0.0
    for o in range(octaves):
        cell = max(base >> o, 1)
        y = ii / cell
        x = jj / cell
        y0 = np.floor(y); x0 = np.floor(x)
        fy = y - y0; fx = x - x0
        sy = fy * fy * (3 - 2 * fy); sx = fx * fx * (3 - 2 * fx)

        def lat(a, b):
            h = (a.astype(np.int64) * 73856093) ^ (b.astype(np.int64) * 19349663) ^ np.int64(seed * 83492791 + o * 2654435761)
            h = (h ^ (h >> 13)) * np.int64(1274126177)
            h = h ^ (h >> 16)
            return (h & 0xFFFFFF).astype(np.float64) / float(0xFFFFFF)

        v00 = lat(y0, x0); v01 = lat(y0, x0 + 1); v10 = lat(y0 + 1, x0); v11 = lat(y0 + 1, x0 + 1)
        out += amp * ((v00 * (1 - sx) + v01 * sx) * (1 - sy) + (v10 * (1 - sx) + v11 * sx) * sy)
        norm += amp
        amp *= 0.55
    out /= norm
    return lo + (hi - lo) * out


def value_noise_dem_torch(out, row0, seed=2, octaves=8, base=1024, lo=1.0, hi=1001.0, chunk=256):
    """``value_noise_dem`` evaluated on the device, row chunk by row chunk, straight into ``out`` (a
    float64 torch tensor of shape (nrows, ncols), e.g. a view of a tile's elevation field): the same
    IEEE operations and the same 64-bit integer hash as the NumPy version, so both give identical
    bits (tests/test_sharded_gloo.py).  A 8192 x 65536 shard (config 4) takes about a second."""
    import torch
    nrows, ncols = out.shape
    dev = out.device
    jj = torch.arange(ncols, dtype=torch.float64, device=dev)[None, :]
    for a in range(0, nrows, chunk):
        b = min(a + chunk, nrows)
        ii = torch.arange(row0 + a, row0 + b, dtype=torch.float64, device=dev)[:, None]
        acc = torch.zeros((b - a, ncols), dtype=torch.float64, device=dev)
        amp, norm = 1.0, 0.0
        for o in range(octaves):
            cell = max(base >> o, 1)
            y = ii / cell
            x = jj / cell
            y0 = torch.floor(y); x0 = torch.floor(x)
            fy = y - y0; fx = x - x0
            sy = fy * fy * (3 - 2 * fy); sx = fx * fx * (3 - 2 * fx)
            salt = (seed * 83492791 + o * 2654435761)

            def lat(p, q):
                h = (p.to(torch.int64) * 73856093) ^ (q.to(torch.int64) * 19349663) ^ salt
                h = (h ^ (h >> 13)) * 1274126177
                h = h ^ (h >> 16)
                return (h & 0xFFFFFF).to(torch.float64) / float(0xFFFFFF)

            v00 = lat(y0, x0); v01 = lat(y0, x0 + 1); v10 = lat(y0 + 1, x0); v11 = lat(y0 + 1, x0 + 1)
            acc += amp * ((v00 * (1 - sx) + v01 * sx) * (1 - sy) + (v10 * (1 - sx) + v11 * sx) * sy)
            norm += amp
            amp *= 0.55
        acc /= norm
        out[a:b] = lo + (hi - lo) * acc
    return out


def _host_lib():
    src = os.path.join(_HERE, "csrc_host", "priority_flood.c")
    so = os.path.join(_HERE, "libpdm_synth.so")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src])
    L = ct.CDLL(so)
    L.pdm_priority_flood_eps2.restype = ct.c_int64
    L.pdm_priority_flood_eps2.argtypes = [np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), ct.c_int64, ct.c_int64,
                                          ct.c_double, ct.c_int]
    return L


def priority_flood(elev, eps=1e-3, wrap_rows=False):
    """Priority-flood + epsilon conditioning (in a copy): every cell gets a strictly descending
    path to the border (no interior pits / flats).  ``wrap_rows``: rows wrap around and only the
    left/right columns are outlets, so the block can be stacked vertically into one seamless,
    conditioned DEM."""
    E = np.array(elev, dtype=np.float64, order="C", copy=True)
    n = _host_lib().pdm_priority_flood_eps2(E, E.shape[0], E.shape[1], float(eps), int(bool(wrap_rows)))
    if n < 0:
        raise MemoryError("priority_flood")
    return E


def conditioned_fractal_dem(n, seed=0, eps=1e-3, wrap_rows=False, **kw):
    """The fractal DEM after hydrological conditioning: long river networks draining to the border
    (BASELINE.json configs[1] primary variant)."""
    return priority_flood(fractal_dem(n, seed, **kw), eps, wrap_rows)
