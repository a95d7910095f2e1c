// slopes.cu -- K1: D-infinity slope magnitude / flow direction (Tarboton 1997) for every cell.
//
// Reference behaviour: _tarboton_slopes_directions (dem_processing.py:1753-1903),
// _calc_direction (1942-1991), _get_d1_d2 (1905-1938), tables 173-193.
//
// One thread per cell, one pass over the grid: the reference's 8 whole-array facet
// passes, its 4 copy-from-interior passes and its 24 edge/corner facet passes collapse
// into "evaluate the in-bounds facets of this cell in ascending order, keep the strictly
// largest squared magnitude"; a border cell first evaluates the one interior cell the
// copy passes (1782-1795) would have copied from.  All arithmetic is IEEE double in the
// reference's operation order (compiled with -fmad=false; the sums that matter use
// explicit _rn intrinsics), so mag is bit-identical and direction differs only by the
// ulp-level difference between CUDA's atan2 and the host libm.
//
// Shards: the kernel runs over the owned rows of a window (Win); rows above/below come from
// halo rows, and "border" means the border of the global grid, so a row-sharded run
// computes exactly what one big tile would.
//
// Roofline: 8 B read (elev) + 16 B written (mag, direction) + 1 B (flat0) per cell.
#include <stdlib.h>
#include <string.h>

#include "pdm_internal.cuh"

namespace {

struct Geom {
    const double *dX, *dY, *dg, *thA, *thB;   // indexed by local fence (between local rows f, f+1)
    const double *rdX, *rdY, *rdg;            // correctly rounded reciprocals of dX, dY, dg
};

// a / b, correctly rounded, from the correctly rounded reciprocal rb = RN(1/b) (Markstein's
// sequence: q <- q + (a - b q) rb, twice; the first pass makes q faithful, the second one is
// then exact-residual + one rounding = IEEE quotient).  5 fp64 operations instead of the ~28 of
// the generic division.  Preconditions (checked by the caller): a == 0 or 1e-150 <= |a| <= 1e150,
// 1e-150 <= b <= 1e150, so nothing over/underflows.  pdm_selftest_division() compares it with
// __ddiv_rn on the device.
__device__ __forceinline__ double mdiv(double a, double b, double rb)
{
    double q = __dmul_rn(a, rb);
    double r = __fma_rn(-q, b, a);
    q = __fma_rn(r, rb, q);
    r = __fma_rn(-q, b, a);
    return __fma_rn(r, rb, q);
}

// facet tables (dem_processing.py:173-193): cardinal neighbour e1, diagonal neighbour e2,
// direction = r * a + q * pi/2
__constant__ int c_e1r[8] = {0, -1, -1, 0, 0, 1, 1, 0};
__constant__ int c_e1c[8] = {1, 0, 0, -1, -1, 0, 0, 1};
__constant__ int c_e2r[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
__constant__ int c_e2c[8] = {1, 1, -1, -1, -1, -1, 1, 1};

__device__ __forceinline__ void facet_update(double e0, double e1, double e2, double d1, double d2,
                                             double dg, double th, double qpi2, double a,
                                             double &m2, double &dr)
{
    double s1 = __ddiv_rn(__dsub_rn(e0, e1), d1);
    double s2 = __ddiv_rn(__dsub_rn(e1, e2), d2);
    double s1_2 = __dmul_rn(s1, s1);
    double sd = __ddiv_rn(__dsub_rn(e0, e2), dg);
    double r = atan2(s2, s1);
    double rad2 = __dadd_rn(s1_2, __dmul_rn(s2, s2));
    bool s1le = s1 <= 0.0, s2le = s2 <= 0.0, s1gt = s1 > 0.0, s2gt = s2 > 0.0;
    if ((s1le && s2gt) || (r > th)) { rad2 = __dmul_rn(sd, sd); r = th; }   // 1973-1976
    if ((s1gt && s2le) || (r < 0.0)) { rad2 = s1_2; r = 0.0; }              // 1978-1981
    if (s1le && (s2le || (s2gt && (sd <= 0.0)))) rad2 = -1.0;               // 1983-1984
    if (rad2 > m2) {                                                        // 1986-1989
        m2 = rad2;
        dr = __dadd_rn(__dmul_rn(r, a), qpi2);
    }
}

// All in-grid facets of the cell at local row i, column j, ascending facet index, strict max
// (m2, dr in/out).  Upper facets (0-3) use fence i-1, lower facets (4-7) fence i
// (_get_d1_d2 1913-1934).
template <bool INTERIOR>
__device__ __forceinline__ void cell_facets(const double *__restrict__ E, const Win &w,
                                            int64_t i, int64_t j, const Geom &g, double &m2, double &dr)
{
    const int64_t C = w.C;
    const double e0 = __ldg(E + i * C + j);
    const double HALF_PI_Q[8] = {0.0 * PDM_PI / 2, 1.0 * PDM_PI / 2, 1.0 * PDM_PI / 2, 2.0 * PDM_PI / 2,
                                 2.0 * PDM_PI / 2, 3.0 * PDM_PI / 2, 3.0 * PDM_PI / 2, 4.0 * PDM_PI / 2};
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int64_t i1 = i + c_e1r[k], j1 = j + c_e1c[k], i2 = i + c_e2r[k], j2 = j + c_e2c[k];
        if (!INTERIOR) {
            if (!w.row_in_grid(i1) || j1 < 0 || j1 >= C || !w.row_in_grid(i2) || j2 < 0 || j2 >= C) continue;
        }
        const int64_t f = (k < 4) ? i - 1 : i;
        const bool typeA = (k == 0 || k == 3 || k == 4 || k == 7);
        const double dx = __ldg(g.dX + f), dy = __ldg(g.dY + f);
        const double d1 = typeA ? dx : dy, d2 = typeA ? dy : dx;
        const double th = typeA ? __ldg(g.thA + f) : __ldg(g.thB + f);
        const double dgf = __ldg(g.dg + f);
        const double e1 = __ldg(E + i1 * C + j1), e2 = __ldg(E + i2 * C + j2);
        facet_update(e0, e1, e2, d1, d2, dgf, th, HALF_PI_Q[k], (k & 1) ? -1.0 : 1.0, m2, dr);
    }
}

// ------------------------------------------------------------------------------------------
// Fast interior path.  Same result as cell_facets<true> bit for bit, with ~4x fewer fp64
// instructions: the parity formulation spends 3 divisions, 1 atan2 and a dozen compares on
// every one of the 8 facets, but
//   * for positive spacings the signs of s1, s2, sd are the signs of the elevation
//     differences, so a facet's case (unclamped / clamped to the cardinal / clamped to the
//     diagonal / uphill) is known before any division and only the quotients that case needs
//     are formed (uphill facets: none);
//   * "r > theta" is decided by cross-multiplying (n2*d1^2 vs n1*d2^2); inside a 1e-8 relative
//     guard band (exact ties on planar test surfaces) the whole cell takes the parity path, so
//     the decision always agrees with comparing the computed angles;
//   * the direction needs atan2 only for the facet that finally wins.
// Every quotient that is formed uses the same IEEE operations as the parity path, so the
// strict-greater tie-breaking between facets sees identical numbers.  Cells with a non-finite
// neighbour, a tiny non-zero difference (quotient could underflow) or unusual spacing take the
// parity path.
// ------------------------------------------------------------------------------------------
struct FacetBest {
    double m2, dr;      // best squared magnitude and its direction (dr valid unless `lazy`)
    double s1, s2;      // slopes of the best facet when its angle is still to be computed
    double a, qpi2;
    bool lazy;
};

// One facet, branch-free (lanes of a warp are in different cases, so per-case branches would
// serialise).  n1 = e0-e1, n2 = e1-e2, nd = e0-e2, all finite; spacings > 0.  d1sq/d2sq = d1^2, d2^2.
// Returns false when the "r > theta" test falls inside the guard band (the cell then takes the
// parity path).
__device__ __forceinline__ bool facet_fast(double n1, double n2, double nd, double d1, double d2, double dg,
                                           double r1, double r2, double rg,
                                           double d1sq, double d2sq, double th, double qpi2, double a, FacetBest &b)
{
    const bool p1 = n1 > 0.0, p2 = n2 > 0.0;
    // r > theta  <=>  s2/s1 > d2/d1  <=>  n2*d1^2 > n1*d2^2   (s1, s2 > 0)
    const double u = __dmul_rn(n2, d1sq), v = __dmul_rn(n1, d2sq);
    const bool A = p1 && p2;
    const bool band = A && !(fabs(u - v) > 1e-8 * v);
    const bool over = u > v;
    const bool a_under = A && !over;              // unclamped: rad2 = s1^2 + s2^2, r = atan2(s2, s1)
    const bool use_s1 = a_under || (p1 && !p2);   // ... or clamped to the cardinal: rad2 = s1^2, r = 0
    const bool use_sd = (A && over) || (!p1 && p2 && nd > 0.0);   // clamped to the diagonal: rad2 = sd^2, r = theta
    const double q1 = mdiv(use_s1 ? n1 : nd, use_s1 ? d1 : dg, use_s1 ? r1 : rg);   // s1 or sd
    const double q2 = mdiv(n2, d2, r2);                                              // s2 (only used when unclamped)
    double rad2 = __dmul_rn(q1, q1);
    if (a_under) rad2 = __dadd_rn(rad2, __dmul_rn(q2, q2));
    if ((use_s1 || use_sd) && rad2 > b.m2) {
        b.m2 = rad2;
        b.lazy = a_under;
        b.s1 = q1; b.s2 = q2; b.a = a; b.qpi2 = qpi2;
        b.dr = __dadd_rn(__dmul_rn(use_sd ? th : 0.0, a), qpi2);
    }
    return !band;
}

// returns false when the cell must take the parity path
__device__ __forceinline__ bool cell_fast(const double *__restrict__ E, const Win &w, int64_t i, int64_t j, const Geom &g,
                                          double &m2, double &dr)
{
    const int64_t C = w.C;
    const double *row = E + i * C + j;
    const double c = __ldg(row), eE = __ldg(row + 1), eW = __ldg(row - 1);
    const double eN = __ldg(row - C), eNE = __ldg(row - C + 1), eNW = __ldg(row - C - 1);
    const double eS = __ldg(row + C), eSE = __ldg(row + C + 1), eSW = __ldg(row + C - 1);
    const double dXu = __ldg(g.dX + i - 1), dYu = __ldg(g.dY + i - 1), dgu = __ldg(g.dg + i - 1);
    const double dXl = __ldg(g.dX + i), dYl = __ldg(g.dY + i), dgl = __ldg(g.dg + i);
    const double tAu = __ldg(g.thA + i - 1), tBu = __ldg(g.thB + i - 1), tAl = __ldg(g.thA + i), tBl = __ldg(g.thB + i);
    const double rXu = __ldg(g.rdX + i - 1), rYu = __ldg(g.rdY + i - 1), rgu = __ldg(g.rdg + i - 1);
    const double rXl = __ldg(g.rdX + i), rYl = __ldg(g.rdY + i), rgl = __ldg(g.rdg + i);
    // spacing sanity (warp-uniform): 1e-150 < d < 1e150, aspect angle away from 0 and pi/2
    const double lo = 1e-5, hi = PDM_PI / 2 - 1e-5;
    bool ok = dXu > 1e-150 && dYu > 1e-150 && dXl > 1e-150 && dYl > 1e-150 && dgu < 1e150 && dgl < 1e150 &&
              tAu > lo && tAu < hi && tBu > lo && tBu < hi && tAl > lo && tAl < hi && tBl > lo && tBl < hi;
    // elevations: each is 0 or 1e-130 <= |e| <= 1e140 (exponent test on the integer pipe).  Then
    // every non-zero difference of two of them is >= 2^-53 * 1e-130 > 1e-150 and < 1e150: no
    // quotient below can underflow, overflow or see a NaN.
    {
        const double ev[9] = {c, eE, eW, eN, eNE, eNW, eS, eSE, eSW};
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const unsigned hi32 = (unsigned)__double2hiint(ev[k]) & 0x7fffffffu;
            const bool zero = (hi32 | (unsigned)__double2loint(ev[k])) == 0u;
            ok = ok && (zero || (hi32 >= 0x25000000u && hi32 <= 0x5D000000u));   // 2^-431 .. 2^465
        }
    }
    if (!ok) return false;
    // differences: cardinal (c - e1), cardinal->diagonal (e1 - e2), diagonal (c - e2)
    const double cE = __dsub_rn(c, eE), cN = __dsub_rn(c, eN), cW = __dsub_rn(c, eW), cS = __dsub_rn(c, eS);
    const double cNE = __dsub_rn(c, eNE), cNW = __dsub_rn(c, eNW), cSW = __dsub_rn(c, eSW), cSE = __dsub_rn(c, eSE);
    const double n2[8] = {__dsub_rn(eE, eNE), __dsub_rn(eN, eNE), __dsub_rn(eN, eNW), __dsub_rn(eW, eNW),
                          __dsub_rn(eW, eSW), __dsub_rn(eS, eSW), __dsub_rn(eS, eSE), __dsub_rn(eE, eSE)};
    const double n1[8] = {cE, cN, cN, cW, cW, cS, cS, cE};
    const double ndv[8] = {cNE, cNE, cNW, cNW, cSW, cSW, cSE, cSE};
    FacetBest b;
    b.m2 = -1.0; b.dr = -1.0; b.s1 = 0.0; b.s2 = 0.0; b.a = 1.0; b.qpi2 = 0.0; b.lazy = false;
    const double H = PDM_PI;
    const double xu2 = __dmul_rn(dXu, dXu), yu2 = __dmul_rn(dYu, dYu), xl2 = __dmul_rn(dXl, dXl), yl2 = __dmul_rn(dYl, dYl);
    bool sure = true;
    sure &= facet_fast(n1[0], n2[0], ndv[0], dXu, dYu, dgu, rXu, rYu, rgu, xu2, yu2, tAu, 0.0 * H / 2, 1.0, b);
    sure &= facet_fast(n1[1], n2[1], ndv[1], dYu, dXu, dgu, rYu, rXu, rgu, yu2, xu2, tBu, 1.0 * H / 2, -1.0, b);
    sure &= facet_fast(n1[2], n2[2], ndv[2], dYu, dXu, dgu, rYu, rXu, rgu, yu2, xu2, tBu, 1.0 * H / 2, 1.0, b);
    sure &= facet_fast(n1[3], n2[3], ndv[3], dXu, dYu, dgu, rXu, rYu, rgu, xu2, yu2, tAu, 2.0 * H / 2, -1.0, b);
    sure &= facet_fast(n1[4], n2[4], ndv[4], dXl, dYl, dgl, rXl, rYl, rgl, xl2, yl2, tAl, 2.0 * H / 2, 1.0, b);
    sure &= facet_fast(n1[5], n2[5], ndv[5], dYl, dXl, dgl, rYl, rXl, rgl, yl2, xl2, tBl, 3.0 * H / 2, -1.0, b);
    sure &= facet_fast(n1[6], n2[6], ndv[6], dYl, dXl, dgl, rYl, rXl, rgl, yl2, xl2, tBl, 3.0 * H / 2, 1.0, b);
    sure &= facet_fast(n1[7], n2[7], ndv[7], dXl, dYl, dgl, rXl, rYl, rgl, xl2, yl2, tAl, 4.0 * H / 2, -1.0, b);
    if (!sure) return false;
    m2 = b.m2;
    dr = b.lazy ? __dadd_rn(__dmul_rn(atan2(b.s2, b.s1), b.a), b.qpi2) : b.dr;
    return true;
}

// ------------------------------------------------------------------------------------------
// Tiled kernel (default).  A block stages the elevation tile + halo in shared memory with bulk
// asynchronous copies (TMA, one cp.async.bulk per row, mbarrier completion) and the per-fence geometry of
// its rows once; a thread then computes FOUR consecutive cells of a row from registers (3 x 6 window read
// with 16-byte shared loads, outputs written with 16-byte stores).  Against the one-cell-per-thread kernel
// (9 elevation + 22 geometry loads per cell through L1) that is 18 shared loads per four cells, and the
// facet bookkeeping keeps only (m2, facet code, s1, s2) per candidate instead of seven doubles: the
// stencil is bound by the fp64 pipe, and every integer / select instruction saved frees an issue slot for it.
// Same arithmetic as cell_fast, bit for bit (tests/test_gpu_parity.py::test_fast_stencil_*).
// ------------------------------------------------------------------------------------------
struct FenceGeom {          // geometry of one fence (between two rows), shared memory
    double dX, dY, dg, thA, thB, rdX, rdY, rdg, x2, y2;
    int ok, pad;
};

// one facet on values; best = (m2, code, s1, s2) with code = facet | lazy << 3 | clamped-to-diagonal << 4
__device__ __forceinline__ bool facet_lean(int k, double n1, double n2, double nd, double d1, double d2, double dg,
                                           double r1, double r2, double rg, double d1sq, double d2sq,
                                           double &m2, int &code, double &bs1, double &bs2)
{
    const bool p1 = n1 > 0.0, p2 = n2 > 0.0;
    const double u = __dmul_rn(n2, d1sq), v = __dmul_rn(n1, d2sq);
    const bool A = p1 && p2;
    const bool band = A && !(fabs(u - v) > 1e-8 * v);
    const bool over = u > v;
    const bool a_under = A && !over;
    const bool use_s1 = a_under || (p1 && !p2);
    const bool use_sd = (A && over) || (!p1 && p2 && nd > 0.0);
    const double q1 = mdiv(use_s1 ? n1 : nd, use_s1 ? d1 : dg, use_s1 ? r1 : rg);
    const double q2 = mdiv(n2, d2, r2);
    double rad2 = __dmul_rn(q1, q1);
    if (a_under) rad2 = __dadd_rn(rad2, __dmul_rn(q2, q2));
    if ((use_s1 || use_sd) && rad2 > m2) {
        m2 = rad2;
        code = k | (a_under ? 8 : 0) | (use_sd ? 16 : 0);
        bs1 = q1; bs2 = q2;
    }
    return !band;
}

// returns false when the cell must take the parity path.  u / l: the fences above / below the cell's row.
__device__ __forceinline__ bool cell_lean(double c, double eE, double eW, double eN, double eNE, double eNW,
                                          double eS, double eSE, double eSW, const FenceGeom &u, const FenceGeom &l,
                                          double &m2_out, double &dr_out)
{
    bool ok = true;
    {
        const double ev[9] = {c, eE, eW, eN, eNE, eNW, eS, eSE, eSW};
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const unsigned hi32 = (unsigned)__double2hiint(ev[k]) & 0x7fffffffu;
            const bool zero = (hi32 | (unsigned)__double2loint(ev[k])) == 0u;
            ok = ok && (zero || (hi32 >= 0x25000000u && hi32 <= 0x5D000000u));   // 2^-431 .. 2^465 (see cell_fast)
        }
    }
    if (!ok) return false;
    const double cE = __dsub_rn(c, eE), cN = __dsub_rn(c, eN), cW = __dsub_rn(c, eW), cS = __dsub_rn(c, eS);
    const double cNE = __dsub_rn(c, eNE), cNW = __dsub_rn(c, eNW), cSW = __dsub_rn(c, eSW), cSE = __dsub_rn(c, eSE);
    double m2 = -1.0, bs1 = 0.0, bs2 = 0.0;
    int code = -1;
    bool sure = true;
    sure &= facet_lean(0, cE, __dsub_rn(eE, eNE), cNE, u.dX, u.dY, u.dg, u.rdX, u.rdY, u.rdg, u.x2, u.y2, m2, code, bs1, bs2);
    sure &= facet_lean(1, cN, __dsub_rn(eN, eNE), cNE, u.dY, u.dX, u.dg, u.rdY, u.rdX, u.rdg, u.y2, u.x2, m2, code, bs1, bs2);
    sure &= facet_lean(2, cN, __dsub_rn(eN, eNW), cNW, u.dY, u.dX, u.dg, u.rdY, u.rdX, u.rdg, u.y2, u.x2, m2, code, bs1, bs2);
    sure &= facet_lean(3, cW, __dsub_rn(eW, eNW), cNW, u.dX, u.dY, u.dg, u.rdX, u.rdY, u.rdg, u.x2, u.y2, m2, code, bs1, bs2);
    sure &= facet_lean(4, cW, __dsub_rn(eW, eSW), cSW, l.dX, l.dY, l.dg, l.rdX, l.rdY, l.rdg, l.x2, l.y2, m2, code, bs1, bs2);
    sure &= facet_lean(5, cS, __dsub_rn(eS, eSW), cSW, l.dY, l.dX, l.dg, l.rdY, l.rdX, l.rdg, l.y2, l.x2, m2, code, bs1, bs2);
    sure &= facet_lean(6, cS, __dsub_rn(eS, eSE), cSE, l.dY, l.dX, l.dg, l.rdY, l.rdX, l.rdg, l.y2, l.x2, m2, code, bs1, bs2);
    sure &= facet_lean(7, cE, __dsub_rn(eE, eSE), cSE, l.dX, l.dY, l.dg, l.rdX, l.rdY, l.rdg, l.x2, l.y2, m2, code, bs1, bs2);
    if (!sure) return false;
    m2_out = m2;
    if (code < 0) { dr_out = -1.0; return true; }
    // direction of the winning facet: r * a + q * pi/2 (tables dem_processing.py:185-193)
    const int k = code & 7;
    const double a = (k & 1) ? -1.0 : 1.0;
    const double qpi2 = __dmul_rn((double)((k + 1) >> 1), PDM_PI / 2);
    const bool typeA = (0x99 >> k) & 1;                      // facets 0, 3, 4, 7
    const FenceGeom &f = k < 4 ? u : l;
    const double th = typeA ? f.thA : f.thB;
    const double r = (code & 8) ? atan2(bs2, bs1) : ((code & 16) ? th : 0.0);
    dr_out = __dadd_rn(__dmul_rn(r, a), qpi2);
    return true;
}

#define ST_TX 32            // threads along x
#define ST_TY 8             // rows of a block

// ST_CPT cells per thread along x; ST_W columns per block; staged columns: 2 left + 2 right (keeps every row
// copy 16-byte aligned)
template <int ST_CPT, int MINB>
__global__ void __launch_bounds__(ST_TX * ST_TY, MINB)
k_slopes_tiled(const double *__restrict__ E, Win w, Geom g, int use_tma,
               double *__restrict__ mag, double *__restrict__ dir, uint8_t *__restrict__ flat0,
               int32_t *__restrict__ label)
{
    constexpr int ST_W = ST_TX * ST_CPT, ST_SW = ST_W + 4;
    __shared__ __align__(128) double tile[(ST_TY + 2) * ST_SW];
    __shared__ FenceGeom fg[ST_TY + 1];
    __shared__ __align__(8) unsigned long long mbar;
    const int64_t C = w.C;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * ST_TX + tx;
    const int64_t c0 = (int64_t)blockIdx.x * ST_W, r0 = w.lo + (int64_t)blockIdx.y * ST_TY;
    // staged rows r0-1 .. r0+ST_TY, columns c0-2 .. c0+ST_W+1, clipped to the local array
    const int64_t ia = max(r0 - 1, (int64_t)0), ib = min(r0 + ST_TY + 1, w.R);
    const int64_t ja = max(c0 - 2, (int64_t)0), jb = min(c0 + ST_W + 2, C);
    const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (use_tma) {
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)((ib - ia) * (jb - ja) * 8);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        }
        if (tid < (int)(ib - ia)) {
            const int64_t gi = ia + tid;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tile[(gi - (r0 - 1)) * ST_SW + (ja - (c0 - 2))]);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst), "l"(E + gi * C + ja), "r"((uint32_t)((jb - ja) * 8)), "r"(mb) : "memory");
        }
    } else {
        for (int idx = tid; idx < (int)((ib - ia) * (jb - ja)); idx += ST_TX * ST_TY) {
            const int64_t gi = ia + idx / (int)(jb - ja), gj = ja + idx % (int)(jb - ja);
            tile[(gi - (r0 - 1)) * ST_SW + (gj - (c0 - 2))] = __ldg(E + gi * C + gj);
        }
    }
    // geometry of the fences r0-1 .. r0+ST_TY-1 (fence f lies between local rows f and f+1)
    if (tid < ST_TY + 1) {
        const int64_t f = r0 - 1 + tid;
        FenceGeom q;
        q.ok = 0; q.pad = 0;
        q.dX = q.dY = q.dg = q.thA = q.thB = q.rdX = q.rdY = q.rdg = q.x2 = q.y2 = 0.0;
        if (f >= 0 && f < w.R - 1) {
            q.dX = __ldg(g.dX + f); q.dY = __ldg(g.dY + f); q.dg = __ldg(g.dg + f);
            q.thA = __ldg(g.thA + f); q.thB = __ldg(g.thB + f);
            q.rdX = __ldg(g.rdX + f); q.rdY = __ldg(g.rdY + f); q.rdg = __ldg(g.rdg + f);
            q.x2 = __dmul_rn(q.dX, q.dX); q.y2 = __dmul_rn(q.dY, q.dY);
            const double lo = 1e-5, hi = PDM_PI / 2 - 1e-5;     // spacing sanity of the fast path (see cell_fast)
            q.ok = q.dX > 1e-150 && q.dY > 1e-150 && q.dg < 1e150 && q.thA > lo && q.thA < hi && q.thB > lo && q.thB < hi;
        }
        fg[tid] = q;
    }
    if (use_tma) {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mb), "r"(0) : "memory");
        }
    }
    __syncthreads();
    const int64_t i = r0 + ty, j0 = c0 + (int64_t)tx * ST_CPT;
    if (i >= w.hi || j0 >= C) return;
    const bool top = w.top(i), bot = w.bottom(i);
    const FenceGeom &gu = fg[ty], &gl = fg[ty + 1];
    const bool row_fast = !top && !bot && gu.ok && gl.ok;
    // 3 x (CPT+2) window of this thread's cells: columns j0-1 .. j0+CPT = staged columns CPT*tx+1 .. CPT*tx+CPT+2
    double win[3][ST_CPT + 2];
    if (row_fast) {
#pragma unroll
        for (int rr = 0; rr < 3; rr++) {
            const double *src = &tile[(ty + rr) * ST_SW + ST_CPT * tx + 1];
#pragma unroll
            for (int q = 0; q < ST_CPT + 2; q++) win[rr][q] = src[q];
        }
    }
    double om[ST_CPT], od[ST_CPT];
    unsigned flats = 0;
#pragma unroll
    for (int u = 0; u < ST_CPT; u++) {
        const int64_t j = j0 + u;
        double m2 = -1.0, dr = -1.0;
        if (j < C) {
            const bool border = top | bot | (j == 0) | (j == C - 1);
            if (!border) {
                bool done = false;
                if (row_fast)
                    done = cell_lean(win[1][u + 1], win[1][u + 2], win[1][u], win[0][u + 1], win[0][u + 2], win[0][u],
                                     win[2][u + 1], win[2][u + 2], win[2][u], gu, gl, m2, dr);
                if (!done) {
                    m2 = -1.0; dr = -1.0;
                    cell_facets<true>(E, w, i, j, g, m2, dr);
                }
            } else {
                // copy-from-interior passes 1782-1795, resolved per border cell (see k_slopes)
                const double hp = PDM_PI / 2, thp = 3 * PDM_PI / 2, tp = 2 * PDM_PI;
                const int64_t si = top ? i + 1 : (bot ? i - 1 : i);
                const int64_t sj = (j == 0) ? 1 : (j == C - 1 ? C - 2 : j);
                double sm = -1.0, sd_ = -1.0;
                cell_facets<true>(E, w, si, sj, g, sm, sd_);
                bool take = true;
                if (j == 0) take = take && (sd_ > hp && sd_ < thp);
                if (j == C - 1) take = take && (sd_ < hp || sd_ > thp);
                if (top) take = take && (sd_ > 0.0 && sd_ < PDM_PI);
                if (bot) take = take && (sd_ > PDM_PI && sd_ < tp);
                if (take) { m2 = sm; dr = sd_; }
                cell_facets<false>(E, w, i, j, g, m2, dr);
            }
        }
        const bool fl = (m2 == -1.0);
        om[u] = (m2 > 0.0) ? sqrt(m2) : m2;                                  // 1901
        od[u] = dr;
        flats |= (fl ? 1u : 0u) << (8 * u);
    }
    const int64_t n = i * C + j0;
    if (ST_CPT == 4 && j0 + ST_CPT <= C && (C & 3) == 0) {
        *reinterpret_cast<double2 *>(mag + n) = make_double2(om[0], om[1]);
        *reinterpret_cast<double2 *>(mag + n + 2) = make_double2(om[2 % ST_CPT], om[3 % ST_CPT]);
        *reinterpret_cast<double2 *>(dir + n) = make_double2(od[0], od[1]);
        *reinterpret_cast<double2 *>(dir + n + 2) = make_double2(od[2 % ST_CPT], od[3 % ST_CPT]);
        *reinterpret_cast<uint32_t *>(flat0 + n) = flats;
    } else if (ST_CPT == 2 && j0 + ST_CPT <= C && (C & 1) == 0) {
        *reinterpret_cast<double2 *>(mag + n) = make_double2(om[0], om[1 % ST_CPT]);
        *reinterpret_cast<double2 *>(dir + n) = make_double2(od[0], od[1 % ST_CPT]);
        *reinterpret_cast<uint16_t *>(flat0 + n) = (uint16_t)((flats & 1u) | ((flats >> 8) & 1u) << 8);
    } else {
#pragma unroll
        for (int u = 0; u < ST_CPT; u++)
            if (j0 + u < C) { mag[n + u] = om[u]; dir[n + u] = od[u]; flat0[n + u] = (uint8_t)((flats >> (8 * u)) & 1u); }
    }
#pragma unroll
    for (int u = 0; u < ST_CPT; u++)
        if (j0 + u < C && ((flats >> (8 * u)) & 1u)) label[n + u] = (int32_t)(n + u);   // union-find root of the flat-region labelling (flats.cu)
}

// 3 blocks per SM (80 registers, 136 B of spills) beats the unconstrained 114 registers / 2 blocks:
// 0.92 -> 0.84 ms at 4096^2 (scripts/stencil_ab.py), same bits
__global__ void __launch_bounds__(256, 3)
k_slopes(const double *__restrict__ E, Win w, Geom g, int fast,
         double *__restrict__ mag, double *__restrict__ dir, uint8_t *__restrict__ flat0,
         int32_t *__restrict__ label)
{
    const int64_t C = w.C;
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = w.lo + (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= w.hi || j >= C) return;
    double m2 = -1.0, dr = -1.0;
    const bool top = w.top(i), bot = w.bottom(i);
    const bool border = top | bot | (j == 0) | (j == C - 1);
    if (!border) {
        if (!(fast && cell_fast(E, w, i, j, g, m2, dr))) {
            m2 = -1.0; dr = -1.0;
            cell_facets<true>(E, w, i, j, g, m2, dr);
        }
    } else {
        // copy-from-interior passes 1782-1795, resolved per border cell: the four
        // sequential whole-row/column copies mean an edge cell looks at its inward
        // neighbour and a corner at its diagonal interior neighbour through two tests.
        const double hp = PDM_PI / 2, thp = 3 * PDM_PI / 2, tp = 2 * PDM_PI;
        const int64_t si = top ? i + 1 : (bot ? i - 1 : i);
        const int64_t sj = (j == 0) ? 1 : (j == C - 1 ? C - 2 : j);
        double sm = -1.0, sd_ = -1.0;
        cell_facets<true>(E, w, si, sj, g, sm, sd_);
        bool take = true;
        if (j == 0) take = take && (sd_ > hp && sd_ < thp);
        if (j == C - 1) take = take && (sd_ < hp || sd_ > thp);
        // a column copy that did not happen leaves -1 in row 1 / R-2, which fails the row tests
        if (top) take = take && (sd_ > 0.0 && sd_ < PDM_PI);
        if (bot) take = take && (sd_ > PDM_PI && sd_ < tp);
        if (take) { m2 = sm; dr = sd_; }
        cell_facets<false>(E, w, i, j, g, m2, dr);
    }
    const int64_t n = i * C + j;
    const bool fl = (m2 == -1.0);
    mag[n] = (m2 > 0.0) ? sqrt(m2) : m2;                                     // 1901
    dir[n] = dr;
    flat0[n] = fl ? 1 : 0;
    if (fl) label[n] = (int32_t)n;  // union-find root of the flat-region labelling (flats.cu)
}

__global__ void k_geometry(const double *__restrict__ dX, const double *__restrict__ dY, int64_t R,
                           double *__restrict__ dg, double *__restrict__ rdX, double *__restrict__ rdY,
                           double *__restrict__ rdg)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f < R - 1) {
        const double g = sqrt(__dadd_rn(__dmul_rn(dX[f], dX[f]), __dmul_rn(dY[f], dY[f])));  // 1962
        dg[f] = g;
        rdX[f] = __drcp_rn(dX[f]); rdY[f] = __drcp_rn(dY[f]); rdg[f] = __drcp_rn(g);
    }
}

// device self-test of mdiv against the IEEE division: pseudo-random operands over the whole
// admissible exponent range plus mantissas near 1, 2 and all-ones
__global__ void k_selftest_div(unsigned long long seed, long long per_thread, unsigned long long *mismatch)
{
    unsigned long long x = seed ^ (0x9E3779B97F4A7C15ULL * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1));
    unsigned long long bad = 0;
    for (long long k = 0; k < per_thread; k++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        unsigned long long ma = x & 0xFFFFFFFFFFFFFULL;
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        unsigned long long mb = x & 0xFFFFFFFFFFFFFULL;
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        const int mode = (int)(x & 7);
        if (mode == 1) ma = 0; if (mode == 2) mb = 0xFFFFFFFFFFFFFULL; if (mode == 3) mb = 0; if (mode == 4) ma = 0xFFFFFFFFFFFFFULL;
        if (mode == 5) mb &= 0xFFULL; if (mode == 6) ma &= 0xFFFULL;
        const long long ea = 1023 - 480 + (long long)((x >> 8) % 960), eb = 1023 - 480 + (long long)((x >> 24) % 960);
        const unsigned long long sa = (x >> 40) & 1;
        const double a = __longlong_as_double((long long)((sa << 63) | ((unsigned long long)ea << 52) | ma));
        const double b = __longlong_as_double((long long)(((unsigned long long)eb << 52) | mb));
        const double q0 = __ddiv_rn(a, b), q1 = mdiv(a, b, __drcp_rn(b));
        if (__double_as_longlong(q0) != __double_as_longlong(q1)) bad++;
    }
    if (bad) atomicAdd(mismatch, bad);
}

}  // namespace

int pdm_launch_geometry(pdm_tile *t)
{
    int64_t n = t->R;
    k_geometry<<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>(t->dX, t->dY, t->R, t->dg, t->rdX, t->rdY, t->rdg);
    PDM_LAUNCHED();
    return PDM_OK;
}

int pdm_launch_slopes(pdm_tile *t) { return pdm_launch_slopes_rows(t, t->win.lo, t->win.hi); }

// the stencil on the owned rows [row_lo, row_hi) only (rows row_lo - 1 and row_hi of ELEV must be in place): lets
// the upload of a tile overlap its stencil, chunk by chunk (pdm_tile_upload_slopes_directions)
int pdm_launch_slopes_rows(pdm_tile *t, int64_t row_lo, int64_t row_hi)
{
    Geom g{t->dX, t->dY, t->dg, t->thA, t->thB, t->rdX, t->rdY, t->rdg};
    Win w = t->win;
    if (row_lo < w.lo || row_hi > w.hi || row_lo >= row_hi) { pdm_set_error("pdm_launch_slopes_rows: bad row range"); return PDM_ERR_ARG; }
    w.lo = row_lo; w.hi = row_hi;
    dim3 block(32, 8);
    dim3 grid((unsigned)((w.C + 31) / 32), (unsigned)((w.hi - w.lo + 7) / 8));
    // PYDEM_B200_STENCIL=parity runs the literal 8 x (3 div + atan2) formulation everywhere (tests
    // compare both bit for bit)
    // PYDEM_B200_STENCIL=v1: the one-cell-per-thread fast kernel of round 1 (A/B runs)
    static int fast = -1;
    if (fast < 0) { const char *e = getenv("PYDEM_B200_STENCIL"); fast = (e && !strcmp(e, "parity")) ? 0 : ((e && !strcmp(e, "v1")) ? 1 : ((e && !strncmp(e, "tiled", 5) && e[5]) ? atoi(e + 5) : 2)); }
    if (fast >= 2 && !t->stencil_parity) {
        dim3 b2(ST_TX, ST_TY);
        // bulk copies need 16-byte aligned rows: an even number of columns (the arrays are 256-byte aligned)
        const int use_tma = (w.C % 2 == 0) ? 1 : 0;
        const unsigned gy = (unsigned)((w.hi - w.lo + ST_TY - 1) / ST_TY);
#define ST_LAUNCH(CPT, MINB) k_slopes_tiled<CPT, MINB><<<dim3((unsigned)((w.C + ST_TX * CPT - 1) / (ST_TX * CPT)), gy), b2, 0, t->stream>>>( \
            t->elev, w, g, use_tma, t->mag, t->dir, t->flat0, t->label)
        switch (fast) {
            case 3: ST_LAUNCH(4, 3); break;
            case 4: ST_LAUNCH(2, 2); break;
            case 5: ST_LAUNCH(2, 3); break;
            case 6: ST_LAUNCH(1, 3); break;
            case 7: ST_LAUNCH(1, 2); break;
            case 8: ST_LAUNCH(1, 4); break;
            case 9: ST_LAUNCH(2, 4); break;
            case 10: ST_LAUNCH(2, 5); break;
            case 11: ST_LAUNCH(2, 6); break;
            case 12: ST_LAUNCH(1, 5); break;
            case 13: ST_LAUNCH(1, 6); break;
            case 14: ST_LAUNCH(4, 4); break;
            case 15: ST_LAUNCH(4, 2); break;
            default: ST_LAUNCH(2, 4); break;      // best of the A/B on B200 (scripts/stencil_ab.py): 64 registers, 4 blocks / SM
        }
        PDM_LAUNCHED();
        return PDM_OK;
    }
    k_slopes<<<grid, block, 0, t->stream>>>(t->elev, w, g, t->stencil_parity ? 0 : fast, t->mag, t->dir, t->flat0, t->label);
    PDM_LAUNCHED();
    return PDM_OK;
}

// number of (a, b) pairs out of blocks*256*per_thread for which mdiv(a,b) != a/b (must be 0)
int pdm_launch_selftest_div(unsigned long long seed, int blocks, long long per_thread, unsigned long long *mismatch_host)
{
    unsigned long long *d = nullptr;
    PDM_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
    PDM_CUDA(cudaMemset(d, 0, sizeof(unsigned long long)));
    k_selftest_div<<<blocks, 256>>>(seed, per_thread, d);
    PDM_LAUNCHED();
    PDM_CUDA(cudaMemcpy(mismatch_host, d, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    PDM_CUDA(cudaFree(d));
    return PDM_OK;
}
