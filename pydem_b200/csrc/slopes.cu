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
#include "pdm_internal.cuh"

namespace {

struct Geom {
    const double *dX, *dY, *dg, *thA, *thB;   // indexed by local fence (between local rows f, f+1)
};

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

__global__ void __launch_bounds__(256)
k_slopes(const double *__restrict__ E, Win w, Geom g,
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
        cell_facets<true>(E, w, i, j, g, m2, dr);
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
                           double *__restrict__ dg)
{
    int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f < R - 1) dg[f] = sqrt(__dadd_rn(__dmul_rn(dX[f], dX[f]), __dmul_rn(dY[f], dY[f])));  // 1962
}

}  // namespace

int pdm_launch_geometry(pdm_tile *t)
{
    int64_t n = t->R;
    k_geometry<<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>(t->dX, t->dY, t->R, t->dg);
    PDM_LAUNCHED();
    return PDM_OK;
}

int pdm_launch_slopes(pdm_tile *t)
{
    Geom g{t->dX, t->dY, t->dg, t->thA, t->thB};
    const Win &w = t->win;
    dim3 block(32, 8);
    dim3 grid((unsigned)((w.C + 31) / 32), (unsigned)((w.hi - w.lo + 7) / 8));
    k_slopes<<<grid, block, 0, t->stream>>>(t->elev, w, g, t->mag, t->dir, t->flat0, t->label);
    PDM_LAUNCHED();
    return PDM_OK;
}
