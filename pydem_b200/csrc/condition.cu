// condition.cu -- elevation conditioning on the device.
//
// The reference runs three passes over `elev` inside calc_slopes_directions when the default
// flags are on (dem_processing.py:601-609):
//   calc_fill_pit_artifacts 396-426   small closed depressions whose whole rim is exactly one
//                                     unit higher are raised by one
//   calc_fill_flats 551-585 -> _fill_flat 308-394
//                                     every 8-connected region of locally minimal cells is
//                                     re-interpolated between its higher ("source") and equal
//                                     ("drain") rim using two in-region chamfer distances
//                                     (utils.get_distance 374-402)
//   calc_pit_drain_paths 428-548      every strict local minimum grows along its lowest rim
//                                     until a lower cell is found, then the path pit -> drain is
//                                     carved into a monotone ramp; pits are visited in ascending
//                                     elevation and each sees the carving of the previous ones
//
// Device formulation
//   * regions: union-find labelling (flats.cu) of the mask, one Region record per root, all
//     per-region quantities (size, window, rim minimum, centre of mass ...) by atomics on exact
//     integer/ordered-float quantities -> deterministic;
//   * the reference evaluates each region on a bounding-box window grown by one cell; only the
//     region's cells and their 8-neighbours are ever read, so the window survives only in the two
//     places where it leaks into the arithmetic: dmax = window size (get_distance's start value)
//     and the window-relative centre of mass (find_centroid, utils.py:450-468);
//   * get_distance is a Jacobi relaxation that stops after the first sweep that reaches every
//     region cell (so its values are "cheapest path of at most k steps", not converged
//     distances): all regions relax together, one launch per sweep, each region with its own
//     stopping sweep; rim cells are never stored -- whether a rim cell is a source/drain is
//     recomputed from its elevation;
//   * pit drain paths are inherently sequential (each pit reads what the earlier ones carved), so
//     one thread block walks the sorted pit list and does each pit's rim search, candidate
//     selection and ramp cooperatively.  Pits of equal elevation are visited in raster order
//     (the reference's np.argsort tie order is platform-defined).
// Arithmetic is IEEE-exact (no fma contraction, correctly rounded div/sqrt), results are
// bit-identical to the reference on NaN-free elevation.
#include "pdm_internal.cuh"
#include "np_sum.cuh"

#include <cub/cub.cuh>
#include <limits.h>

namespace {

enum : unsigned {
    RG_EDGE = 1u,        // region has a cell on the array's outer ring
    RG_SRC = 2u,         // rim has a higher cell
    RG_DRAIN = 4u,       // rim has a cell of the region's elevation
    RG_BAD = 8u,         // artifact test failed
    RG_SINGLE = 16u,     // one-cell region (special cases 311-326)
    RG_SKIP = 32u,       // _fill_flat returns before interpolating
    RG_SRC_CENT = 64u,   // peak: the centroid is the only source (350-356)
    RG_DRN_EDGE = 128u,  // drain = region cells on the array ring (364-368)
    RG_DRN_CENT = 256u,  // closed depression: the centroid is the only drain (369-373)
    RG_NEED_CENT = 512u,
};

struct Region {
    int32_t root, count, nedge;
    int32_t imin, imax, jmin, jmax;
    int32_t wi0, wj0;              // origin of the grown, clipped window
    unsigned flags;
    int32_t centroid;
    int32_t reachH, reachL;        // cells whose distance is below dmax
    int32_t stopH, stopL;          // sweep after which the relaxation stopped (-1: still running)
    double e, eH, srcmax, dmax, cx, cy;
    unsigned long long esrc;       // ordered bits of the lowest higher rim cell
    unsigned long long si, sj;     // sums of row / column indices (centre of mass)
    unsigned long long best;       // bits of the smallest distance to the centre of mass
};

__device__ __forceinline__ unsigned long long ord_bits(double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double ord_value(unsigned long long o)
{
    const unsigned long long b = (o >> 63) ? (o & 0x7fffffffffffffffULL) : ~o;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ bool sea_ok(double e, int below_sea) { return below_sea ? (e != 0.0) : (e > 0.0); }

// mask = (3x3 minimum >= centre) & sea mask [& not an array corner]; label := own index.
// scipy's 'reflect' border only repeats cells that are in the window anyway.
__global__ void __launch_bounds__(256)
k_locmin(const double *__restrict__ E, uint8_t *__restrict__ mask, int32_t *__restrict__ label, int64_t R, int64_t C,
         int below_sea, int skip_corners)
{
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= R || j >= C) return;
    const int64_t n = i * C + j;
    const double e = E[n];
    bool ok = sea_ok(e, below_sea);
    for (int di = -1; di <= 1 && ok; di++)
        for (int dj = -1; dj <= 1; dj++) {
            const int64_t ni = i + di, nj = j + dj;
            if (ni < 0 || ni >= R || nj < 0 || nj >= C) continue;
            if (E[ni * C + nj] < e) { ok = false; break; }
        }
    if (skip_corners && (i == 0 || i == R - 1) && (j == 0 || j == C - 1)) ok = false;   // 570-573
    mask[n] = ok ? 1 : 0;
    label[n] = (int32_t)n;
}

__global__ void __launch_bounds__(256)
k_count_roots(const uint8_t *__restrict__ mask, const int32_t *__restrict__ label, int64_t N, unsigned long long *ctr)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool root = n < N && mask[n] && label[n] == (int32_t)n;
    const unsigned b = __ballot_sync(0xffffffffu, root);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(&ctr[CT_TMP0], (unsigned long long)__popc(b));
}

__global__ void __launch_bounds__(256)
k_reg_init(const double *__restrict__ E, const uint8_t *__restrict__ mask, const int32_t *__restrict__ label, int64_t N,
           int32_t *__restrict__ regid, Region *__restrict__ reg, unsigned long long *ctr)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N || !mask[n] || label[n] != (int32_t)n) return;
    const int32_t id = (int32_t)atomicAdd(&ctr[CT_TMP1], 1ULL);
    regid[n] = id;
    Region g;
    memset(&g, 0, sizeof(g));
    g.root = (int32_t)n;
    g.imin = g.jmin = INT_MAX;
    g.imax = g.jmax = -1;
    g.centroid = INT_MAX;
    g.stopH = g.stopL = -1;
    g.e = E[n];
    g.esrc = ~0ULL;
    g.best = ~0ULL;
    reg[id] = g;
}

// per mask cell: size, window, centre-of-mass sums, rim classification
template <int ARTIFACT>
__global__ void __launch_bounds__(256)
k_reg_stats(const double *__restrict__ E, const uint8_t *__restrict__ mask, const int32_t *__restrict__ label,
            const int32_t *__restrict__ regid, Region *__restrict__ reg, int64_t R, int64_t C)
{
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= R || j >= C) return;
    const int64_t n = i * C + j;
    if (!mask[n]) return;
    const int32_t root = label[n];
    Region *g = reg + regid[root];
    const double e = g->e;
    atomicAdd(&g->count, 1);
    atomicMin(&g->imin, (int32_t)i); atomicMax(&g->imax, (int32_t)i);
    atomicMin(&g->jmin, (int32_t)j); atomicMax(&g->jmax, (int32_t)j);
    unsigned fl = 0;
    if (i == 0 || j == 0 || i == R - 1 || j == C - 1) { fl |= RG_EDGE; atomicAdd(&g->nedge, 1); }
    if (!ARTIFACT) {
        atomicAdd(&g->si, (unsigned long long)i);
        atomicAdd(&g->sj, (unsigned long long)j);
    }
    unsigned long long lowest = ~0ULL;
    for (int di = -1; di <= 1; di++)
        for (int dj = -1; dj <= 1; dj++) {
            const int64_t ni = i + di, nj = j + dj;
            if ((di == 0 && dj == 0) || ni < 0 || ni >= R || nj < 0 || nj >= C) continue;
            const int64_t m = ni * C + nj;
            if (mask[m] && label[m] == root) continue;
            const double v = E[m];
            if (ARTIFACT) {
                if (!(__dsub_rn(v, 1.0) == e)) fl |= RG_BAD;                         // 416
            } else {
                if (v == e) fl |= RG_DRAIN;                                          // 330
                if (v > e) { fl |= RG_SRC; const unsigned long long o = ord_bits(v); if (o < lowest) lowest = o; }   // 331
            }
        }
    if (fl) atomicOr(&g->flags, fl);
    if (lowest != ~0ULL) atomicMin(&g->esrc, lowest);
}

__global__ void __launch_bounds__(256)
k_art_apply(double *__restrict__ E, const uint8_t *__restrict__ mask, const int32_t *__restrict__ label,
            const int32_t *__restrict__ regid, const Region *__restrict__ reg, int64_t N, double max_area,
            unsigned long long *ctr)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N || !mask[n]) return;
    const Region *g = reg + regid[label[n]];
    // regions whose grown window was clipped (they touch the array ring) are skipped (410-412)
    if ((g->flags & (RG_EDGE | RG_BAD)) || (double)g->count > max_area) return;
    E[n] = __dadd_rn(E[n], 1.0);                                                    // 422
    if (g->root == (int32_t)n) atomicAdd(&ctr[CT_TMP0], 1ULL);
}

// the decision tree of _fill_flat (345-376) per region
__global__ void __launch_bounds__(128)
k_reg_decide(Region *__restrict__ reg, int64_t nreg, int64_t R, int64_t C, int tol, int peaks, int pits)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nreg) return;
    Region g = reg[k];
    const int64_t wi0 = g.imin > 0 ? g.imin - 1 : 0, wi1 = (g.imax + 2 < R) ? g.imax + 2 : R;   // utils.grow_obj 430-448
    const int64_t wj0 = g.jmin > 0 ? g.jmin - 1 : 0, wj1 = (g.jmax + 2 < C) ? g.jmax + 2 : C;
    g.wi0 = (int32_t)wi0; g.wj0 = (int32_t)wj0;
    g.dmax = (double)((wi1 - wi0) * (wj1 - wj0));                                   // utils.py:392
    g.cx = __ddiv_rn((double)(g.si - (unsigned long long)g.count * (unsigned long long)wi0), (double)g.count);
    g.cy = __ddiv_rn((double)(g.sj - (unsigned long long)g.count * (unsigned long long)wj0), (double)g.count);
    unsigned f = g.flags;
    if (g.count == 1) {
        f |= RG_SINGLE;
        g.stopH = g.stopL = 0;
    } else {
        bool go = true;
        if (f & RG_SRC) {
            const double es = ord_value(g.esrc);
            const double e1 = __dadd_rn(g.e, 1.0);
            g.eH = e1 <= es ? e1 : es;                                              // 347
            g.srcmax = __dadd_rn(es, (double)tol);                                  // 348
        } else if (peaks) {
            g.eH = __dadd_rn(g.e, 0.5);                                             // 351
            f |= RG_SRC_CENT | RG_NEED_CENT;
        } else {
            go = false;
        }
        if (go) {
            if (f & RG_DRAIN) {
            } else if (f & RG_EDGE) {
                f |= RG_DRN_EDGE;
                if (g.nedge == g.count) go = false;                                 // 367-368
            } else if (pits) {
                f |= RG_DRN_CENT | RG_NEED_CENT;
            } else {
                go = false;
            }
        }
        if (!go) { f |= RG_SKIP; g.stopH = g.stopL = 0; }
    }
    g.flags = f;
    reg[k] = g;
}

// utils.find_centroid 450-468: the region cell nearest the window-relative centre of mass
// (first in raster order among equals), in two passes of atomic minima
template <int PASS>
__global__ void __launch_bounds__(256)
k_reg_centroid(const uint8_t *__restrict__ mask, const int32_t *__restrict__ label, const int32_t *__restrict__ regid,
               Region *__restrict__ reg, int64_t N, int64_t C)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N || !mask[n]) return;
    Region *g = reg + regid[label[n]];
    if (!(g->flags & RG_NEED_CENT)) return;
    const int64_t i = n / C, j = n - i * C;
    const double dx = __dsub_rn((double)(i - g->wi0), g->cx), dy = __dsub_rn((double)(j - g->wj0), g->cy);
    const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);       // d >= 0: bits are ordered
    if (PASS == 0) atomicMin(&g->best, b);
    else if (b == g->best) atomicMin(&g->centroid, (int32_t)n);
}

struct DistBufs { double *H[2], *L[2]; };

__global__ void __launch_bounds__(256)
k_dist_init(const uint8_t *__restrict__ mask, const int32_t *__restrict__ label, const int32_t *__restrict__ regid,
            Region *__restrict__ reg, int64_t N, int64_t R, int64_t C, DistBufs B, double *__restrict__ out)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N || !mask[n]) return;
    Region *g = reg + regid[label[n]];
    const unsigned f = g->flags;
    if (f & RG_SINGLE) return;
    const bool cent = g->centroid == (int32_t)n;
    if ((f & RG_SRC_CENT) && cent) out[n] = g->eH;                                  // 353
    if (f & RG_SKIP) return;
    const int64_t i = n / C, j = n - i * C;
    const bool ring = (i == 0 || j == 0 || i == R - 1 || j == C - 1);
    const double dH = ((f & RG_SRC_CENT) && cent) ? 0.0 : g->dmax;
    const double dL = (((f & RG_DRN_CENT) && cent) || ((f & RG_DRN_EDGE) && ring)) ? 0.0 : g->dmax;
    B.H[0][n] = dH; B.L[0][n] = dL;
    if (dH < g->dmax) atomicAdd(&g->reachH, 1);
    if (dL < g->dmax) atomicAdd(&g->reachL, 1);
}

// one Jacobi sweep of utils.get_distance 394-400 for every region that is still relaxing
__global__ void __launch_bounds__(256)
k_dist_step(const double *__restrict__ E, const uint8_t *__restrict__ mask, const int32_t *__restrict__ label,
            const int32_t *__restrict__ regid, Region *__restrict__ reg, int64_t R, int64_t C, DistBufs B, int sweep)
{
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= R || j >= C) return;
    const int64_t n = i * C + j;
    if (!mask[n]) return;
    const int32_t root = label[n];
    Region *g = reg + regid[root];
    const bool doH = g->stopH < 0, doL = g->stopL < 0;
    if (!doH && !doL) return;
    const unsigned f = g->flags;
    const double e = g->e, dmax = g->dmax, srcmax = g->srcmax;
    const double *curH = B.H[sweep & 1], *curL = B.L[sweep & 1];
    double oH = curH[n], oL = curL[n];          // the footprints contain the centre
    double aH = oH, aL = oL;
    const double SQRT2 = 1.4142135623730951;
    for (int di = -1; di <= 1; di++)
        for (int dj = -1; dj <= 1; dj++) {
            const int64_t ni = i + di, nj = j + dj;
            if ((di == 0 && dj == 0) || ni < 0 || ni >= R || nj < 0 || nj >= C) continue;
            const int64_t m = ni * C + nj;
            double vH, vL;
            if (mask[m] && label[m] == root) {
                vH = doH ? curH[m] : dmax;
                vL = doL ? curL[m] : dmax;
            } else {
                const double v = E[m];
                vH = (!(f & RG_SRC_CENT) && v > e && v <= srcmax) ? 0.0 : dmax;   // 331, 348
                vL = ((f & RG_DRAIN) && v == e) ? 0.0 : dmax;                     // 330
            }
            if (di == 0 || dj == 0) { if (vH < oH) oH = vH; if (vL < oL) oL = vL; }
            if (vH < aH) aH = vH;
            if (vL < aL) aL = vL;
        }
    if (doH) {
        const double cur = curH[n];
        const double a = __dadd_rn(oH, 1.0), b = __dadd_rn(aH, SQRT2);
        double v = a < b ? a : b;
        if (cur < v) v = cur;
        B.H[(sweep + 1) & 1][n] = v;
        if (!(cur < dmax) && v < dmax) atomicAdd(&g->reachH, 1);
    }
    if (doL) {
        const double cur = curL[n];
        const double a = __dadd_rn(oL, 1.0), b = __dadd_rn(aL, SQRT2);
        double v = a < b ? a : b;
        if (cur < v) v = cur;
        B.L[(sweep + 1) & 1][n] = v;
        if (!(cur < dmax) && v < dmax) atomicAdd(&g->reachL, 1);
    }
}

__global__ void __launch_bounds__(128)
k_dist_check(Region *__restrict__ reg, int64_t nreg, int sweep, unsigned long long *ctr)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nreg) return;
    Region *g = reg + k;
    bool active = false;
    if (g->stopH < 0) { if (g->reachH == g->count) g->stopH = sweep + 1; else active = true; }   // 401-402
    if (g->stopL < 0) { if (g->reachL == g->count) g->stopL = sweep + 1; else active = true; }
    if (active) atomicAdd(&ctr[CT_TMP0], 1ULL);
}

__global__ void __launch_bounds__(256)
k_fill_apply(const double *__restrict__ E, const uint8_t *__restrict__ mask, const int32_t *__restrict__ label,
             const int32_t *__restrict__ regid, const Region *__restrict__ reg, int64_t R, int64_t C, DistBufs B,
             double *__restrict__ out, int peaks)
{
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= R || j >= C) return;
    const int64_t n = i * C + j;
    if (!mask[n]) return;
    const Region *g = reg + regid[label[n]];
    const unsigned f = g->flags;
    const double e = g->e;
    if (f & RG_SINGLE) {                                                             // 311-326
        int nwin = 0, nhigh = 0;
        double lowest = INFINITY;
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++) {
                const int64_t ni = i + di, nj = j + dj;
                if ((di == 0 && dj == 0) || ni < 0 || ni >= R || nj < 0 || nj >= C) continue;
                nwin++;
                const double v = E[ni * C + nj];
                if (v > e) { nhigh++; if (v < lowest) lowest = v; }
            }
        if (nhigh == nwin) return;                                                  // a one-cell pit
        if (nhigh > 0) {
            double d = __dsub_rn(lowest, e);
            if (!(d < 1.0)) d = 1.0;                                                // min(1.0, d)
            out[n] = __dadd_rn(e, __dsub_rn(d, 0.01));
        } else if (peaks) {
            out[n] = __dadd_rn(e, 0.5);
        }
        return;
    }
    if (f & RG_SKIP) return;
    const bool cent = g->centroid == (int32_t)n;
    const bool ring = (i == 0 || j == 0 || i == R - 1 || j == C - 1);
    bool replaced;                                                                   // the last assignment of `replace` wins
    if (f & RG_DRN_EDGE) replaced = ring;
    else if (f & RG_DRN_CENT) replaced = cent;
    else if (f & RG_SRC_CENT) replaced = cent;
    else replaced = false;
    if (replaced) return;
    const double dH = B.H[g->stopH & 1][n], dL = B.L[g->stopL & 1][n];
    const double l2 = __dmul_rn(dL, dL), h2 = __dmul_rn(dH, dH);
    out[n] = __ddiv_rn(__dadd_rn(__dmul_rn(g->eH, l2), __dmul_rn(e, h2)), __dadd_rn(l2, h2));   // 382
}

// ---------------------------------------------------------------------------------------------
// pit drain paths
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_strict_minima(const double *__restrict__ E, uint8_t *__restrict__ flag, int64_t R, int64_t C, int below_sea)
{
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= R || j >= C) return;
    const int64_t n = i * C + j;
    // on the array ring the reflected border puts the centre itself under the footprint: never a pit (446-448)
    bool ok = i > 0 && j > 0 && i < R - 1 && j < C - 1;
    if (ok) {
        const double e = E[n];
        ok = sea_ok(e, below_sea);
        for (int di = -1; di <= 1 && ok; di++)
            for (int dj = -1; dj <= 1; dj++) {
                if (di == 0 && dj == 0) continue;
                if (!(E[n + di * C + dj] > e)) { ok = false; break; }
            }
    }
    flag[n] = ok ? 1 : 0;
}

__global__ void __launch_bounds__(256)
k_pit_keys(const double *__restrict__ E, const int32_t *__restrict__ cells, int64_t n, unsigned long long *__restrict__ keys)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) keys[k] = ord_bits(E[cells[k]]);
}

struct PathArgs {
    double *E;
    const double *dX, *dY;
    int64_t R, C;
    const int32_t *pits;
    int64_t npits;
    uint8_t *in_area, *in_border;   // zeroed byte maps
    int32_t *path, *kept;           // [N + 1] each
    int max_iter, max_dist;
    double max_dist_xy;
    unsigned long long *ctr;        // CT_TMP0: pits without a drain, CT_TMP1: max iterations
};

#define PP_THREADS 256

// one block, all pits in order
__global__ void __launch_bounds__(PP_THREADS) k_pit_paths(PathArgs a)
{
    typedef cub::BlockScan<int, PP_THREADS> Scan;
    typedef cub::BlockReduce<double, PP_THREADS> RedD;
    __shared__ typename Scan::TempStorage scan_tmp;
    __shared__ typename RedD::TempStorage red_tmp;
    __shared__ double s_emin, s_best;
    __shared__ int s_i0, s_i1, s_j0, s_j1;     // bounding box of the pit area (inclusive)
    __shared__ int s_n, s_bestpos, s_nk;
    __shared__ unsigned long long s_undrained, s_maxit;
    const int tid = threadIdx.x;
    const int64_t R = a.R, C = a.C;
    if (tid == 0) { s_undrained = 0; s_maxit = 0; }
    __syncthreads();

    for (int64_t p = 0; p < a.npits; p++) {
        const int32_t pit = a.pits[p];
        const int ip = (int)(pit / C), jp = (int)(pit - (int64_t)ip * C);
        const double epit = __ldcg(a.E + pit);                                       // 460
        if (tid == 0) { s_i0 = s_i1 = ip; s_j0 = s_j1 = jp; a.path[0] = pit; a.in_area[pit] = 1; }
        if (tid < 9 && tid != 4) {
            const int ni = ip + tid / 3 - 1, nj = jp + tid % 3 - 1;
            if (ni >= 0 && ni < R && nj >= 0 && nj < C) a.in_border[(int64_t)ni * C + nj] = 1;
        }
        __syncthreads();
        int npath = 1, ndrain = 0, it_used = 0;
        bool found = false;
        for (int it = 0; it < a.max_iter; it++) {                                    // 462-476
            it_used = it;
            const int wi0 = max(s_i0 - 1, 0), wi1 = min(s_i1 + 1, (int)R - 1);
            const int wj0 = max(s_j0 - 1, 0), wj1 = min(s_j1 + 1, (int)C - 1);
            const int ww = wj1 - wj0 + 1, wn = (wi1 - wi0 + 1) * ww;
            // lowest rim elevation
            double mn = INFINITY;
            for (int k = tid; k < wn; k += PP_THREADS) {
                const int64_t m = (int64_t)(wi0 + k / ww) * C + (wj0 + k % ww);
                if (a.in_border[m]) { const double v = __ldcg(a.E + m); if (v < mn) mn = v; }
            }
            mn = RedD(red_tmp).Reduce(mn, cub::Min());
            if (tid == 0) s_emin = mn;
            __syncthreads();
            const double emin = s_emin;
            if (emin == INFINITY) break;                                             // empty rim (466-467)
            // the rim cells at that elevation, in ascending index order, appended to the path
            int base = npath;
            for (int k0 = 0; k0 < wn; k0 += PP_THREADS) {
                const int k = k0 + tid;
                int64_t m = -1;
                int flag = 0;
                if (k < wn) {
                    m = (int64_t)(wi0 + k / ww) * C + (wj0 + k % ww);
                    flag = (a.in_border[m] && __ldcg(a.E + m) == emin) ? 1 : 0;
                }
                int pos, total;
                Scan(scan_tmp).ExclusiveSum(flag, pos, total);
                if (flag) a.path[base + pos] = (int32_t)m;
                base += total;
                __syncthreads();
            }
            const int nlow = base - npath;
            if (emin < epit) { found = true; ndrain = nlow; break; }                 // 471-473
            // the pit area swallows them (475-476); the rim is rebuilt around the new cells
            for (int k = tid; k < nlow; k += PP_THREADS) {
                const int32_t c = a.path[npath + k];
                a.in_area[c] = 1; a.in_border[c] = 0;
                const int ci = c / (int)C, cj = c - ci * (int)C;
                atomicMin(&s_i0, ci); atomicMax(&s_i1, ci); atomicMin(&s_j0, cj); atomicMax(&s_j1, cj);
            }
            __syncthreads();
            for (int k = tid; k < nlow; k += PP_THREADS) {
                const int32_t c = a.path[npath + k];
                const int ci = c / (int)C, cj = c - ci * (int)C;
                for (int di = -1; di <= 1; di++)
                    for (int dj = -1; dj <= 1; dj++) {
                        const int ni = ci + di, nj = cj + dj;
                        if ((di == 0 && dj == 0) || ni < 0 || ni >= R || nj < 0 || nj >= C) continue;
                        const int64_t m = (int64_t)ni * C + nj;
                        if (!a.in_area[m]) a.in_border[m] = 1;
                    }
            }
            npath += nlow;
            __syncthreads();
        }
        if (found && tid == 0 && (unsigned long long)(it_used + 1) > s_maxit) s_maxit = (unsigned long long)(it_used + 1);   // 481
        // ---- choose the drain (484-514)
        bool ok = found;
        if (ok) {
            double best = INFINITY;
            int bestpos = INT_MAX;
            for (int k = tid; k < ndrain; k += PP_THREADS) {
                const int32_t d = a.path[npath + k];
                const int id = d / (int)C, jd = d - id * (int)C;
                const long long di = ip - id, dj = jp - jd;
                if (a.max_dist && di * di + dj * dj > (long long)a.max_dist * a.max_dist) continue;   // 486-494
                const int lo = min(ip, id), hi = max(ip, id);
                double dxm, dy;
                if (ip == id) {                                                      // _get_dX_mean 1993-1997
                    dxm = a.dX[min((int64_t)ip, R - 2)];
                    dy = 0.0;
                } else {
                    dxm = __ddiv_rn(np_sum(a.dX + lo, hi - lo), (double)(hi - lo));
                    dy = np_sum(a.dY + lo, hi - lo);
                }
                const double dx = __dmul_rn(dxm, (double)(jp - jd));
                const double dxy = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                if (a.max_dist_xy > 0.0 && !(dxy <= a.max_dist_xy)) continue;        // 503-509
                if (dxy < best) { best = dxy; bestpos = k; }
            }
            const double bmin = RedD(red_tmp).Reduce(best, cub::Min());
            if (tid == 0) { s_best = bmin; s_bestpos = INT_MAX; }
            __syncthreads();
            if (best == s_best && bestpos != INT_MAX) atomicMin(&s_bestpos, bestpos);   // first of the closest (512-514)
            __syncthreads();
            ok = s_bestpos != INT_MAX;
        }
        if (ok) {
            const int32_t drain = a.path[npath + s_bestpos];
            // walk back from the drain: a cell stays if it touches the cell kept after it; the pit
            // itself is never tested (517-533)
            if (tid == 0) {
                int nk = 0;
                a.kept[nk++] = drain;
                int li = drain / (int)C, lj = drain - li * (int)C;
                for (int k = npath - 1; k >= 1; k--) {
                    const int32_t c = a.path[k];
                    const int ci = c / (int)C, cj = c - ci * (int)C;
                    if (abs(ci - li) <= 1 && abs(cj - lj) <= 1) { a.kept[nk++] = c; li = ci; lj = cj; }
                }
                a.kept[nk++] = pit;
                double e0 = __ldcg(a.E + pit);
                const double ed = __ldcg(a.E + drain);
                if (e0 < ed) {                                                       // 536-537
                    double mn = INFINITY;
                    for (int k = 0; k < nk; k++) { const double v = __ldcg(a.E + a.kept[k]); if (v > ed && v < mn) mn = v; }
                    a.E[pit] = mn;
                }
                s_nk = nk;
            }
            __syncthreads();
            const int nk = s_nk;
            const double e0 = __ldcg(a.E + pit), si = __dsub_rn(__ldcg(a.E + drain), e0);   // 539
            const double step = __ddiv_rn(1.0, (double)(nk - 1));                    // np.linspace(0, 1, nk)
            __syncthreads();
            for (int k = tid; k < nk; k += PP_THREADS) {
                const double lin = (k == nk - 1) ? 1.0 : __dmul_rn((double)k, step);
                a.E[a.kept[nk - 1 - k]] = __dadd_rn(e0, __dmul_rn(lin, si));         // 540
            }
        } else if (tid == 0) {
            s_undrained++;                                                           // 478-480, 491-493
        }
        // ---- wipe the maps
        {
            const int wi0 = max(s_i0 - 1, 0), wi1 = min(s_i1 + 1, (int)R - 1);
            const int wj0 = max(s_j0 - 1, 0), wj1 = min(s_j1 + 1, (int)C - 1);
            const int ww = wj1 - wj0 + 1, wn = (wi1 - wi0 + 1) * ww;
            __syncthreads();
            for (int k = tid; k < wn; k += PP_THREADS) {
                const int64_t m = (int64_t)(wi0 + k / ww) * C + (wj0 + k % ww);
                a.in_area[m] = 0; a.in_border[m] = 0;
            }
        }
        __syncthreads();
    }
    if (tid == 0) { a.ctr[CT_TMP0] = s_undrained; a.ctr[CT_TMP1] = s_maxit; }
}

int require_standalone(const pdm_tile *t, const char *who)
{
    if (!t) { pdm_set_error("%s: NULL tile", who); return PDM_ERR_ARG; }
    if (!t->have_elev) { pdm_set_error("%s: no elevation uploaded", who); return PDM_ERR_STATE; }
    if (t->win.lo != 0 || t->win.hi != t->R || t->win.Rg != t->R) {
        pdm_set_error("%s: elevation conditioning is not available on a row shard (regions and pit paths cross shard boundaries); "
                      "condition the DEM before it is sharded", who);
        return PDM_ERR_STATE;
    }
    return PDM_OK;
}

int read_ctr(pdm_tile *t)
{
    PDM_CUDA(cudaMemcpyAsync(t->h_counters, t->d_counters, CT_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    return PDM_OK;
}

// label the mask in t->flat0 and build the Region table; *reg_out is cudaMalloc'ed (caller frees)
int build_regions(pdm_tile *t, Region **reg_out, int64_t *nreg_out)
{
    *reg_out = nullptr; *nreg_out = 0;
    int rc = pdm_launch_ccl(t);
    if (rc) return rc;
    PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_TMP0, 0, 2 * sizeof(unsigned long long), t->stream));
    k_count_roots<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>(t->flat0, t->label, t->N, t->d_counters);
    PDM_LAUNCHED();
    if ((rc = read_ctr(t))) return rc;
    const int64_t nreg = (int64_t)t->h_counters[CT_TMP0];
    if (nreg == 0) return PDM_OK;
    PDM_CUDA(cudaMalloc(reg_out, (size_t)nreg * sizeof(Region)));
    k_reg_init<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>(t->elev, t->flat0, t->label, t->N, t->queue, *reg_out,
                                                                     t->d_counters);
    PDM_LAUNCHED();
    *nreg_out = nreg;
    return PDM_OK;
}

void invalidate(pdm_tile *t)
{
    t->have_slopes = t->have_flats = t->have_graph = t->have_uca = false;
    t->queue_ready = false;   // the queue array served as scratch
}

}  // namespace

extern "C" {

void pdm_default_cond_params(pdm_cond_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->fill_flats_below_sea = 0;      // dem_processing.py:106-110
    p->fill_flats_source_tol = 1;
    p->fill_flats_peaks = 1;
    p->fill_flats_pits = 1;
    p->maximum_pit_area = 32.0;       // 154
    p->drain_pits_max_iter = 300;     // 117-119
    p->drain_pits_max_dist = 32;
    p->drain_pits_max_dist_xy = 0.0;
}

// calc_fill_pit_artifacts, dem_processing.py:396-426
int pdm_tile_fill_pit_artifacts(pdm_tile *t, const pdm_cond_params *p, pdm_cond_stats *st)
{
    int rc = require_standalone(t, "pdm_tile_fill_pit_artifacts");
    if (rc) return rc;
    pdm_cond_params def;
    if (!p) { pdm_default_cond_params(&def); p = &def; }
    PDM_CUDA(cudaSetDevice(t->device));
    invalidate(t);
    dim3 block(32, 8), grid((unsigned)((t->C + 31) / 32), (unsigned)((t->R + 7) / 8));
    k_locmin<<<grid, block, 0, t->stream>>>(t->elev, t->flat0, t->label, t->R, t->C, p->fill_flats_below_sea, 0);
    PDM_LAUNCHED();
    Region *reg = nullptr;
    int64_t nreg = 0;
    if ((rc = build_regions(t, &reg, &nreg))) return rc;
    if (st) { st->n_artifact_regions = nreg; st->n_artifacts_filled = 0; }
    if (nreg == 0) return PDM_OK;
    k_reg_stats<1><<<grid, block, 0, t->stream>>>(t->elev, t->flat0, t->label, t->queue, reg, t->R, t->C);
    PDM_LAUNCHED();
    PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_TMP0, 0, sizeof(unsigned long long), t->stream));
    k_art_apply<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>(t->elev, t->flat0, t->label, t->queue, reg, t->N,
                                                                      p->maximum_pit_area, t->d_counters);
    PDM_LAUNCHED();
    rc = read_ctr(t);
    cudaFree(reg);
    if (rc) return rc;
    if (st) st->n_artifacts_filled = (int64_t)t->h_counters[CT_TMP0];
    return PDM_OK;
}

// calc_fill_flats, dem_processing.py:551-585 (runs calc_fill_pit_artifacts first when maximum_pit_area != 0)
int pdm_tile_fill_flats(pdm_tile *t, const pdm_cond_params *p, pdm_cond_stats *st)
{
    int rc = require_standalone(t, "pdm_tile_fill_flats");
    if (rc) return rc;
    pdm_cond_params def;
    if (!p) { pdm_default_cond_params(&def); p = &def; }
    if (st) memset(st, 0, sizeof(*st));
    if (p->maximum_pit_area != 0.0 && (rc = pdm_tile_fill_pit_artifacts(t, p, st))) return rc;   // 557-558
    PDM_CUDA(cudaSetDevice(t->device));
    invalidate(t);
    dim3 block(32, 8), grid((unsigned)((t->C + 31) / 32), (unsigned)((t->R + 7) / 8));
    const unsigned lin = (unsigned)((t->N + 255) / 256);
    k_locmin<<<grid, block, 0, t->stream>>>(t->elev, t->flat0, t->label, t->R, t->C, p->fill_flats_below_sea, 1);
    PDM_LAUNCHED();
    Region *reg = nullptr;
    int64_t nreg = 0;
    if ((rc = build_regions(t, &reg, &nreg))) return rc;
    if (st) st->n_flat_regions = nreg;
    if (nreg == 0) return PDM_OK;
    const unsigned rgrid = (unsigned)((nreg + 127) / 128);
    // scratch: four distance planes in mag / dir / twi / the Cell array, the result in uca
    DistBufs B;
    B.H[0] = t->mag; B.H[1] = t->dir; B.L[0] = t->twi; B.L[1] = (double *)t->cell;
    double *out = t->uca;
    auto body = [&]() -> int {
        k_reg_stats<0><<<grid, block, 0, t->stream>>>(t->elev, t->flat0, t->label, t->queue, reg, t->R, t->C);
        PDM_LAUNCHED();
        k_reg_decide<<<rgrid, 128, 0, t->stream>>>(reg, nreg, t->R, t->C, p->fill_flats_source_tol, p->fill_flats_peaks,
                                                    p->fill_flats_pits);
        PDM_LAUNCHED();
        k_reg_centroid<0><<<lin, 256, 0, t->stream>>>(t->flat0, t->label, t->queue, reg, t->N, t->C);
        PDM_LAUNCHED();
        k_reg_centroid<1><<<lin, 256, 0, t->stream>>>(t->flat0, t->label, t->queue, reg, t->N, t->C);
        PDM_LAUNCHED();
        PDM_CUDA(cudaMemcpyAsync(out, t->elev, (size_t)t->N * 8, cudaMemcpyDeviceToDevice, t->stream));   // filled = data.copy() (561)
        k_dist_init<<<lin, 256, 0, t->stream>>>(t->flat0, t->label, t->queue, reg, t->N, t->R, t->C, B, out);
        PDM_LAUNCHED();
        int sweep = 0;
        for (;;) {
            // a few sweeps per host round trip; finished regions ignore the extra ones
            for (int k = 0; k < 4; k++, sweep++) {
                if (k == 3) PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_TMP0, 0, sizeof(unsigned long long), t->stream));
                k_dist_step<<<grid, block, 0, t->stream>>>(t->elev, t->flat0, t->label, t->queue, reg, t->R, t->C, B, sweep);
                PDM_LAUNCHED();
                k_dist_check<<<rgrid, 128, 0, t->stream>>>(reg, nreg, sweep, t->d_counters);
                PDM_LAUNCHED();
            }
            int rc2 = read_ctr(t);
            if (rc2) return rc2;
            if (t->h_counters[CT_TMP0] == 0) break;
            if ((int64_t)sweep > t->N) { pdm_set_error("pdm_tile_fill_flats: distance relaxation did not stop"); return PDM_ERR_STATE; }
        }
        if (st) st->distance_sweeps = sweep;
        k_fill_apply<<<grid, block, 0, t->stream>>>(t->elev, t->flat0, t->label, t->queue, reg, t->R, t->C, B, out,
                                                    p->fill_flats_peaks);
        PDM_LAUNCHED();
        PDM_CUDA(cudaMemcpyAsync(t->elev, out, (size_t)t->N * 8, cudaMemcpyDeviceToDevice, t->stream));
        PDM_CUDA(cudaStreamSynchronize(t->stream));
        return PDM_OK;
    };
    rc = body();
    cudaFree(reg);
    return rc;
}

// calc_pit_drain_paths, dem_processing.py:428-548
int pdm_tile_pit_drain_paths(pdm_tile *t, const pdm_cond_params *p, pdm_cond_stats *st)
{
    int rc = require_standalone(t, "pdm_tile_pit_drain_paths");
    if (rc) return rc;
    if (!t->have_spacing) { pdm_set_error("pdm_tile_pit_drain_paths: set the spacing first"); return PDM_ERR_STATE; }
    pdm_cond_params def;
    if (!p) { pdm_default_cond_params(&def); p = &def; }
    PDM_CUDA(cudaSetDevice(t->device));
    invalidate(t);
    dim3 block(32, 8), grid((unsigned)((t->C + 31) / 32), (unsigned)((t->R + 7) / 8));
    k_strict_minima<<<grid, block, 0, t->stream>>>(t->elev, t->flat0, t->R, t->C, p->fill_flats_below_sea);
    PDM_LAUNCHED();
    // pits in raster order (ordered compaction), then a stable sort by elevation
    int32_t *cells = t->label;
    int *d_n = (int *)(t->d_counters + CT_TMP0);
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cub::CountingInputIterator<int32_t> idx(0);
    PDM_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, idx, t->flat0, cells, d_n, (int)t->N, t->stream));
    PDM_CUDA(cudaMalloc(&tmp, tmp_bytes));
    cudaError_t e = cub::DeviceSelect::Flagged(tmp, tmp_bytes, idx, t->flat0, cells, d_n, (int)t->N, t->stream);
    g_pdm_launches++;
    if (e == cudaSuccess) e = cudaStreamSynchronize(t->stream);
    int npits = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&npits, d_n, sizeof(int), cudaMemcpyDeviceToHost);
    cudaFree(tmp);
    if (e != cudaSuccess) return pdm_cuda_fail(e, "cub::DeviceSelect::Flagged", __FILE__, __LINE__);
    if (st) { st->n_pits = npits; st->n_pits_undrained = 0; st->path_max_iter = 0; }
    if (npits == 0) return PDM_OK;
    unsigned long long *keys = nullptr, *keys2 = nullptr;
    int32_t *cells2 = nullptr, *kept = nullptr;
    auto body = [&]() -> int {
        PDM_CUDA(cudaMalloc(&keys, (size_t)npits * 8));
        PDM_CUDA(cudaMalloc(&keys2, (size_t)npits * 8));
        PDM_CUDA(cudaMalloc(&cells2, (size_t)npits * 4));
        PDM_CUDA(cudaMalloc(&kept, (size_t)(t->N + 1) * 4));
        k_pit_keys<<<(unsigned)((npits + 255) / 256), 256, 0, t->stream>>>(t->elev, cells, npits, keys);
        PDM_LAUNCHED();
        size_t sb = 0;
        PDM_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sb, keys, keys2, cells, cells2, npits, 0, 64, t->stream));
        void *stmp = nullptr;
        PDM_CUDA(cudaMalloc(&stmp, sb));
        cudaError_t e2 = cub::DeviceRadixSort::SortPairs(stmp, sb, keys, keys2, cells, cells2, npits, 0, 64, t->stream);   // stable (451)
        g_pdm_launches++;
        if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(t->stream);
        cudaFree(stmp);
        if (e2 != cudaSuccess) return pdm_cuda_fail(e2, "cub::DeviceRadixSort::SortPairs", __FILE__, __LINE__);
        PDM_CUDA(cudaMemsetAsync(t->flats, 0, (size_t)t->N, t->stream));
        PDM_CUDA(cudaMemsetAsync(t->link, 0, (size_t)t->N, t->stream));
        PathArgs a;
        a.E = t->elev; a.dX = t->dX; a.dY = t->dY; a.R = t->R; a.C = t->C;
        a.pits = cells2; a.npits = npits;
        a.in_area = t->flats; a.in_border = t->link;
        a.path = t->queue; a.kept = kept;
        a.max_iter = p->drain_pits_max_iter; a.max_dist = p->drain_pits_max_dist; a.max_dist_xy = p->drain_pits_max_dist_xy;
        a.ctr = t->d_counters;
        k_pit_paths<<<1, PP_THREADS, 0, t->stream>>>(a);
        PDM_LAUNCHED();
        int rc2 = read_ctr(t);
        if (rc2) return rc2;
        if (st) { st->n_pits_undrained = (int64_t)t->h_counters[CT_TMP0]; st->path_max_iter = (int64_t)t->h_counters[CT_TMP1]; }
        return PDM_OK;
    };
    rc = body();
    cudaFree(keys); cudaFree(keys2); cudaFree(cells2); cudaFree(kept);
    return rc;
}

}  // extern "C"
