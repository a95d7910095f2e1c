"""CPU restatement of pyDEM's elevation conditioning.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): nothing under ``pydem_b200/`` imports this.

Three passes, all run by the reference inside ``calc_slopes_directions`` when the default flags
are on (dem_processing.py:601-609):

* ``fill_pit_artifacts``  -- calc_fill_pit_artifacts, dem_processing.py:396-426
* ``fill_flats``          -- calc_fill_flats 551-585 -> _fill_flat 308-394, with
                             utils.get_border_mask 342-370, utils.get_distance 374-402,
                             utils.find_centroid 450-468, utils.grow_obj 430-448
* ``pit_drain_paths``     -- calc_pit_drain_paths 428-548, with utils.get_border_index 313-340 and
                             _get_dX_mean 1993-1997

The restatement is written region-by-region on explicit index lists (no ``find_objects`` windows):
what the reference computes on a grown bounding-box window only ever looks at a region's cells and
their 8-neighbours, so the window disappears except where it leaks into the arithmetic (the
``dmax = window size`` start value of get_distance, the window-relative centroid).

Stated limits (also in DESIGN.md):
* NaN-free elevation (``scipy.ndimage.minimum_filter`` has no defined NaN behaviour).
* ``pit_drain_paths`` visits pits in ascending elevation; the reference uses ``np.argsort`` with the
  default (unstable, platform-dependent) sort, so the relative order of pits with EQUAL elevation is
  not defined by the reference.  This restatement (and the CUDA path) break ties by raster index.
* The recursive re-fill at the end of ``_fill_flat`` (378-394) writes into a private copy that is
  dropped (``out = out2`` rebinds a local name), so it has no effect on the result and is not
  restated; only its "maximum iterations" warning is lost.
"""
import numpy as np
from scipy import ndimage as ndi

_K8 = np.ones((3, 3), bool)
_SQRT2 = np.sqrt(2.0)
_NB8 = ((-1, 0), (0, -1), (0, 1), (1, 0), (-1, -1), (-1, 1), (1, -1), (1, 1))


def _min3x3(a):
    """3x3 moving minimum with scipy's default 'reflect' border (== edge replication for a 1-px reach)."""
    p = np.pad(a, 1, mode="edge")
    R, C = a.shape
    out = a.copy()
    for di in (0, 1, 2):
        for dj in (0, 1, 2):
            np.minimum(out, p[di:di + R, dj:dj + C], out=out)
    return out


def _sea_mask(elev, below_sea):
    return (elev != 0) if below_sea else (elev > 0)         # 401-404, 562-563, 444-445


def _regions(mask):
    """8-connected components of mask: list of (flat index array in raster order)."""
    lab, n = ndi.label(mask, structure=_K8)
    if n == 0:
        return []
    idx = np.flatnonzero(lab)
    order = np.argsort(lab.ravel()[idx], kind="stable")
    idx = idx[order]
    cuts = np.flatnonzero(np.diff(lab.ravel()[idx])) + 1
    return np.split(idx, cuts)


# ----------------------------------------------------------------------------------------------
def fill_pit_artifacts(elev, below_sea=False, maximum_pit_area=32.0):
    """Raise by 1 every small closed depression whose whole rim is exactly one unit higher
    (quantisation artifacts).  dem_processing.py:396-426."""
    elev = np.asarray(elev, "float64")
    R, C = elev.shape
    out = elev.copy()
    flat = (_min3x3(elev) >= elev) & _sea_mask(elev, below_sea)                      # 405
    for cells in _regions(flat):
        ii, jj = np.divmod(cells, C)
        # the reference skips regions whose grown window was clipped, i.e. that touch the array edge (410-412)
        if ii.min() == 0 or jj.min() == 0 or ii.max() == R - 1 or jj.max() == C - 1:
            continue
        if len(cells) > maximum_pit_area:                                            # 417
            continue
        inreg = np.zeros((R, C), bool)
        inreg.ravel()[cells] = True
        e = elev.ravel()[cells[0]]                                                   # 415
        ok = True
        for di, dj in _NB8:                                                          # rim = dilation minus region (414)
            ni, nj = ii + di, jj + dj
            rim = ~inreg[ni, nj]
            if not np.all(elev[ni[rim], nj[rim]] - 1 == e):                          # 416
                ok = False
                break
        if ok:
            out.ravel()[cells] += 1                                                  # 422
    return out


# ----------------------------------------------------------------------------------------------
def _window(ii, jj, R, C):
    """Bounding box grown by one cell, clipped to the array (utils.grow_obj 430-448)."""
    i0, i1 = max(0, ii.min() - 1), min(R, ii.max() + 2)
    j0, j1 = max(0, jj.min() - 1), min(C, jj.max() + 2)
    return i0, i1, j0, j1


def _region_distance(region, src):
    """utils.get_distance 374-402: Jacobi relaxation of the 1/sqrt2 chamfer distance, started at
    dmax = window size, stopped after the first sweep that leaves no region cell at dmax (so the
    values are "cheapest path of at most k steps", not converged distances)."""
    dmax = float(region.size)
    d = np.full(region.shape, dmax)
    d[src] = 0
    m, n = region.shape
    for _ in range(region.size):
        p = np.pad(d, 1, mode="edge")
        c = p[1:-1, 1:-1]
        orth = np.minimum(np.minimum(np.minimum(p[:-2, 1:-1], p[2:, 1:-1]), np.minimum(p[1:-1, :-2], p[1:-1, 2:])), c) + 1
        full = c.copy()
        for di in (0, 1, 2):
            for dj in (0, 1, 2):
                full = np.minimum(full, p[di:di + m, dj:dj + n])
        full = full + _SQRT2
        new = np.minimum(np.minimum(orth, full), d)
        d[region] = new[region]
        if (d[region] < dmax).all():
            break
    return d


def _centroid(region):
    """utils.find_centroid 450-468: region cell nearest the window-relative centre of mass, first in raster order."""
    w = np.argwhere(region)
    n = len(w)
    x = float(w[:, 0].sum()) / n
    y = float(w[:, 1].sum()) / n
    dist = np.sqrt((w[:, 0] - x) ** 2 + (w[:, 1] - y) ** 2)
    k = int(np.argmin(dist))
    return int(w[k, 0]), int(w[k, 1])


def _dilate_minus(region):
    p = np.pad(region, 1, mode="constant")
    m, n = region.shape
    grown = np.zeros_like(region)
    for di in (0, 1, 2):
        for dj in (0, 1, 2):
            grown |= p[di:di + m, dj:dj + n]
    return grown & ~region


def _fill_one_flat(roi, out, region, edge, source_tol, peaks, pits):
    """_fill_flat 308-376 for one region (roi = untouched elevation window, out = window of the result)."""
    e = roi[region][0]
    if roi.size <= 9 and region.sum() == 1:                                          # 311-326
        higher = roi > e
        n = int(higher.sum())
        if n == roi.size - 1:
            pass                                                                    # a one-cell pit: left to the pit passes
        elif n > 0:
            out[region] += min(1.0, (roi[higher].min() - e)) - 0.01
        elif peaks:
            out[region] += 0.5
        return
    border = _dilate_minus(region)                                                   # 329
    drain = border & (roi == e)
    source = border & (roi > e)
    replace = None
    if source.any():                                                                 # 345-349
        e_source = roi[source].min()
        eH = min(e + 1.0, e_source)
        source &= (roi <= e_source + source_tol)
    elif peaks:                                                                      # 350-356
        eH = e + 0.5
        c = _centroid(region)
        out[c] = eH
        source[c] = True
        replace = source
    else:
        return
    if drain.any():                                                                  # 361-363
        pass
    elif (region & edge).any():                                                      # 364-368
        replace = drain = region & edge
        if not (region & ~drain).any():
            return
    elif pits:                                                                       # 369-373
        c = _centroid(region)
        drain[c] = True
        replace = drain
    else:
        return
    dH = _region_distance(region, source)                                            # 378-379
    dL = _region_distance(region, drain)
    interp = region if replace is None else (region & ~replace)
    out[interp] = (eH * dL[interp] ** 2 + e * dH[interp] ** 2) / (dL[interp] ** 2 + dH[interp] ** 2)   # 382


def fill_flats(elev, below_sea=False, source_tol=1, peaks=True, pits=True, maximum_pit_area=32.0):
    """calc_fill_flats 551-585 (including its call of calc_fill_pit_artifacts)."""
    elev = np.asarray(elev, "float64")
    if maximum_pit_area:                                                             # 557-558
        elev = fill_pit_artifacts(elev, below_sea, maximum_pit_area)
    data = elev
    R, C = data.shape
    filled = data.copy()
    edge = np.ones((R, C), bool)
    edge[1:-1, 1:-1] = False
    flat = (_min3x3(data) >= data) & _sea_mask(data, below_sea)                      # 568
    flat[0, 0] = flat[-1, 0] = flat[0, -1] = flat[-1, -1] = False                    # 570-573
    for cells in _regions(flat):
        ii, jj = np.divmod(cells, C)
        i0, i1, j0, j1 = _window(ii, jj, R, C)
        region = np.zeros((i1 - i0, j1 - j0), bool)
        region[ii - i0, jj - j0] = True
        _fill_one_flat(data[i0:i1, j0:j1], filled[i0:i1, j0:j1], region, edge[i0:i1, j0:j1], source_tol, peaks, pits)
    return filled


# ----------------------------------------------------------------------------------------------
def _np_mean(a):
    return a.mean()


def pit_drain_paths(elev, dX, dY, below_sea=False, max_iter=300, max_dist=32, max_dist_xy=None, tie_order="raster"):
    """calc_pit_drain_paths 428-548: carve a monotone path from every strict local minimum to the
    nearest lower cell found by growing the pit along its lowest rim.  Returns (elev, n_undrained, max_it).
    Sequential by construction: every pit sees the carving of the pits before it."""
    elev = np.array(elev, dtype="float64", copy=True)
    R, C = elev.shape
    e = elev.ravel()
    dX = np.asarray(dX, "float64"); dY = np.asarray(dY, "float64")
    # strict local minima of the 8 neighbours; the reflected border puts the centre itself into
    # the footprint on the array edge, so edge cells never qualify (446-448)
    p = np.pad(elev, 1, mode="edge")
    nbmin = np.full((R, C), np.inf)
    for di in (0, 1, 2):
        for dj in (0, 1, 2):
            if di == 1 and dj == 1:
                continue
            nbmin = np.minimum(nbmin, p[di:di + R, dj:dj + C])
    pits = np.flatnonzero(((nbmin > elev) & _sea_mask(elev, below_sea)).ravel())
    # 451.  tie_order="numpy" reproduces the reference call literally (platform-dependent ties):
    # used only to show that ties are the sole difference when pinning against the reference
    pits = pits[np.argsort(e[pits]) if tie_order == "numpy" else np.argsort(e[pits], kind="stable")]
    in_area = np.zeros(R * C, bool)
    undrained = 0
    maxit = 0

    def neighbours(cells):
        ii, jj = np.divmod(cells, C)
        out = []
        for di, dj in _NB8:
            ni, nj = ii + di, jj + dj
            ok = (ni >= 0) & (ni < R) & (nj >= 0) & (nj < C)
            out.append(ni[ok] * C + nj[ok])
        return np.concatenate(out)

    for pit in pits:
        area = [int(pit)]
        in_area[pit] = True
        epit = e[pit]                                                                # 460
        path = [int(pit)]
        border = np.setdiff1d(neighbours(np.array([pit])), [pit])
        drain = None
        it_used = 0
        for it in range(max_iter):                                                   # 462-476
            it_used = it
            if border.size == 0:
                break
            eb = e[border]
            emin = eb.min()
            lowest = border[eb == emin]
            if emin < epit:
                drain = lowest
                break
            path += lowest.tolist()
            area += lowest.tolist()
            in_area[lowest] = True
            grown = neighbours(lowest)
            border = np.union1d(border[eb != emin], grown[~in_area[grown]])
        in_area[area] = False
        if drain is None:                                                            # 478-480
            undrained += 1
            continue
        maxit = max(maxit, it_used + 1)
        ip, jp = divmod(int(pit), C)
        Id, Jd = np.divmod(drain, C)
        if max_dist:                                                                 # 486-494
            keep = np.sqrt((ip - Id) ** 2 + (jp - Jd) ** 2) <= max_dist
            if not keep.any():
                undrained += 1
                continue
            drain, Id, Jd = drain[keep], Id[keep], Jd[keep]
        dx = np.empty(len(drain)); dy = np.empty(len(drain))                         # 497-500
        for k, (idr, jdr) in enumerate(zip(Id, Jd)):
            if ip == idr:
                mean_dx = dX[min(ip, dX.size - 1)]
            else:
                mean_dx = _np_mean(dX[min(ip, idr):max(ip, idr)])
            dx[k] = mean_dx * (jp - jdr)
            dy[k] = dY[min(ip, idr):max(ip, idr)].sum()
        dxy = np.sqrt(dx ** 2 + dy ** 2)
        if max_dist_xy:                                                              # 503-509
            keep = dxy <= max_dist_xy
            if not keep.any():
                undrained += 1
                continue
            drain, dxy = drain[keep], dxy[keep]
        if drain.size > 1:                                                           # 512-514
            drain = drain[dxy == dxy.min()]
        drain = int(drain[0])
        path.append(drain)
        # walk back from the drain, dropping every cell that does not touch the cell kept after it;
        # the pit itself (position 0) is never tested (517-533)
        kept = [path[-1]]
        for cell in reversed(path[1:-1]):
            ci, cj = divmod(cell, C)
            ki, kj = divmod(kept[-1], C)
            if abs(ci - ki) <= 1 and abs(cj - kj) <= 1:
                kept.append(cell)
        kept.append(path[0])
        path = kept[::-1]
        if e[pit] < e[drain]:                                                        # 536-537
            ep = e[path]
            e[pit] = ep[ep > e[drain]].min()
        si = e[drain] - e[pit]                                                       # 539-540
        n = len(path)
        lin = np.arange(n, dtype="float64") * (1.0 / (n - 1))                        # np.linspace(0, 1, n)
        lin[-1] = 1.0
        e[path] = e[pit] + lin * si
    return elev, undrained, maxit
