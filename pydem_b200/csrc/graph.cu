// graph.cu -- K3: drainage graph of a tile, never materialised as a sparse matrix.
//
// Reference behaviour: _calc_uca_section_proportion (dem_processing.py:1021-1070),
// _mk_connectivity (1155-1267), the edge filter of _mk_adjacency_matrix (1136-1141), the
// source / inflow-border bookkeeping of _calc_uca_chunk (882-937).
//
// Per cell the reference's CSC/CSR matrix rows reduce to one byte (facet index + which of
// the two receivers survive the filter) and one double (share of the cardinal receiver);
// the in-degree (CSR row length) is counted by looking at the 8 neighbours' bytes.
//
// Algorithmic traffic: link pass reads dir 8 + flats 1 + elev 8 (+2 neighbour elevations from
// L1/L2), writes link 1 + prop 8 (scratch); in-degree pass reads link 1 + prop 8 and writes the
// cell's 32-byte sweep record (Cell).
#include "pdm_internal.cuh"

namespace {

__constant__ int g_e1r[8] = {0, -1, -1, 0, 0, 1, 1, 0};
__constant__ int g_e1c[8] = {1, 0, 0, -1, -1, 0, 0, 1};
__constant__ int g_e2r[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
__constant__ int g_e2c[8] = {1, 1, -1, -1, -1, -1, 1, 1};

// section / proportion of one cell, exactly the reference's operation order.
// Returns section in [-8, 7] or 127 if outside (the reference raises IndexError there).
__device__ __forceinline__ int section_prop(double d, bool flat, double th, double &p_out)
{
    const double hp = PDM_PI / 2.0;
    double qd = floor(__dmul_rn(__ddiv_rn(d, PDM_PI), 2.0));               // 1035
    int q = (qd >= -128.0 && qd <= 127.0) ? (int)qd : 60;                   // int8 cast; NaN/out of range -> bad
    double quad = __dsub_rn(d, __dmul_rn(hp, (double)q));                   // 1037
    const double cth = __dsub_rn(hp, th);
    const int odd = q & 1;                                                  // numpy modulo on int8
    int sec = q * 2 + ((quad > th) && !odd) + ((quad > cth) && odd);        // 1040-1043
    double p = __longlong_as_double(0x7ff8000000000000LL);
    const bool I1 = (sec == 0) | (sec == 1) | (sec == 4) | (sec == 5);      // 1050
    if (I1 && quad <= th) p = __ddiv_rn(quad, th);                          // 1052-1053
    if (I1 && quad > th) p = __ddiv_rn(__dsub_rn(quad, th), cth);           // 1054-1056
    if (!I1 && quad <= cth) p = __ddiv_rn(quad, cth);                       // 1057-1059
    if (!I1 && quad > cth) p = __ddiv_rn(__dsub_rn(quad, cth), th);         // 1060-1062
    if (flat) { sec = -1; p = __longlong_as_double(0x7ff8000000000000LL); } // 1064-1065
    if (sec == 8) sec = 0;                                                  // 1067
    if (sec < -8 || sec > 7) { p_out = p; return 127; }
    const int idx = sec < 0 ? sec + 8 : sec;                                // python negative indexing
    const double a = (idx & 1) ? -1.0 : 1.0;
    p_out = __dsub_rn((1.0 + a) / 2.0, __dmul_rn(a, p));                    // 1068
    return sec;
}

__global__ void __launch_bounds__(256)
k_links(const double *__restrict__ E, const double *__restrict__ dir, const uint8_t *__restrict__ flats,
        const double *__restrict__ th_row, Win w,
        uint8_t *__restrict__ link, double *__restrict__ prop, uint8_t *__restrict__ pitmask,
        unsigned long long *counters)
{
    const int64_t C = w.C;
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = w.lo + (int64_t)blockIdx.y * 8 + threadIdx.y;
    bool pit = false;
    if (i < w.hi && j < C) {
        const int64_t n = i * C + j;
        const bool fl = flats[n] != 0;
        const double e0 = E[n];
        double p;
        int sec = section_prop(dir[n], fl, __ldg(th_row + i), p);
        if (sec == 127) { atomicAdd(&counters[CT_BADSEC], 1ULL); sec = -1; }
        uint8_t lk = LK_NOSEC;
        if (sec >= 0) {
            lk = (uint8_t)sec;
            const int64_t i1 = i + g_e1r[sec], j1 = j + g_e1c[sec];
            const int64_t i2 = i + g_e2r[sec], j2 = j + g_e2c[sec];
            // _mk_connectivity: a receiver exists iff it is inside the grid; filter 1136-1137:
            // weight not NaN, > 1e-8, and the receiver is not higher than the source
            if (w.row_in_grid(i1) && j1 >= 0 && j1 < C) {
                if (p > 1e-8 && __ldg(E + i1 * C + j1) <= e0) lk |= LK_KEEP1;
            }
            if (w.row_in_grid(i2) && j2 >= 0 && j2 < C) {
                const double w2 = __dsub_rn(1.0, p);
                if (w2 > 1e-8 && __ldg(E + i2 * C + j2) <= e0) lk |= LK_KEEP2;
            }
        }
        link[n] = lk;
        prop[n] = p;
        pit = fl && (e0 > 0.0);                                             // pits_bool, 1284
        pitmask[n] = pit ? 1 : 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, pit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[CT_NPITS], (unsigned long long)__popc(m));
}

// export of DEMProcessor.section (int8) for tests / debugging
__global__ void __launch_bounds__(256)
k_section_export(const double *__restrict__ dir, const uint8_t *__restrict__ flats,
                 const double *__restrict__ th_row, int64_t R, int64_t C, int8_t *__restrict__ section)
{
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= R || j >= C) return;  // all local rows
    double p;
    int sec = section_prop(dir[i * C + j], flats[i * C + j] != 0, __ldg(th_row + i), p);
    section[i * C + j] = (int8_t)sec;
}

// does neighbour byte `lk` at relative position (di,dj) drain into the centre cell?
__device__ __forceinline__ int drains_in(uint8_t lk, uint8_t keepbit, uint32_t secmask)
{
    return ((lk & keepbit) && !(lk & (LK_NOSEC | LK_PIT)) && ((secmask >> (lk & LK_SEC_MASK)) & 1u)) ? 1 : 0;
}

// in-degree of every owned cell (+ pit in-edges counted by the pit kernel), source flag, and
// the initial sweep record of the cell: area = dX2*dY2 of the row, taint = 0.  Halo rows of a
// shard (the out-boxes of the sweep) start at zero.
__global__ void __launch_bounds__(256)
k_indeg(uint8_t *__restrict__ link, Win w, const double *__restrict__ row_area, const double *__restrict__ prop,
        const int32_t *__restrict__ pit_in, Cell *__restrict__ cell, unsigned long long *counters)
{
    const int64_t C = w.C;
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;   // all local rows
    const bool in = (i < w.R && j < C);
    const bool own = in && i >= w.lo && i < w.hi;
    int cnt = 0;
    int64_t n = 0;
    if (in) {
        n = i * C + j;
        Cell rec;
        rec.taint = 0.0; rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
        rec.link = link[n];
        if (own) {
            const bool up = w.row_in_grid(i - 1), dn = w.row_in_grid(i + 1), lf = j > 0, rt = j < C - 1;
            // all eight neighbour bytes are loaded unconditionally (a missing neighbour reads the
            // cell's own byte and is masked out), so the loads are independent and in flight together
            // instead of eight load -> test -> branch round trips (0.44 -> 0.2x ms at 4096^2)
            const int64_t dW = lf ? -1 : 0, dE = rt ? 1 : 0, dN = up ? -C : 0, dS = dn ? C : 0;
            const uint8_t bW = link[n + dW], bE = link[n + dE], bN = link[n + dN], bS = link[n + dS];
            const uint8_t bNW = link[n + dN + dW], bNE = link[n + dN + dE], bSW = link[n + dS + dW], bSE = link[n + dS + dE];
            cnt += lf ? drains_in(bW, LK_KEEP1, 0x81u) : 0;                     // W neighbour: e1 = (0,+1)
            cnt += rt ? drains_in(bE, LK_KEEP1, 0x18u) : 0;                     // E: e1 = (0,-1)
            cnt += up ? drains_in(bN, LK_KEEP1, 0x60u) : 0;                     // N: e1 = (+1,0)
            cnt += dn ? drains_in(bS, LK_KEEP1, 0x06u) : 0;                     // S: e1 = (-1,0)
            cnt += (up && lf) ? drains_in(bNW, LK_KEEP2, 0xC0u) : 0;            // NW: e2 = (+1,+1)
            cnt += (up && rt) ? drains_in(bNE, LK_KEEP2, 0x30u) : 0;            // NE: e2 = (+1,-1)
            cnt += (dn && lf) ? drains_in(bSW, LK_KEEP2, 0x03u) : 0;            // SW: e2 = (-1,+1)
            cnt += (dn && rt) ? drains_in(bSE, LK_KEEP2, 0x0Cu) : 0;            // SE: e2 = (-1,-1)
            if (pit_in) cnt += pit_in[n];
            rec.indeg = cnt;
            rec.area = __ldg(row_area + i);                                     // 885, 901
            rec.prop = prop[n];
        } else {
            rec.indeg = 0; rec.area = 0.0; rec.prop = 0.0;
        }
        cell[n] = rec;
    }
    const bool src = own && cnt == 0;
    if (src) link[n] |= LK_SOURCE;                                          // 882-883
    // one add per block: a fifth of all cells are sources, and a per-warp add kept ~300 k atomics
    // queueing on one address
    const int nsrc = __syncthreads_count(src);
    if (threadIdx.x == 0 && threadIdx.y == 0 && nsrc) atomicAdd(&counters[CT_SOURCES], (unsigned long long)nsrc);
}

__device__ __forceinline__ bool sec_in(int sec, uint32_t mask) { return sec >= 0 && ((mask >> sec) & 1u); }

// inflow-border mask of _calc_uca_chunk 909-937 for the owned cells on the border of the global
// grid.  Thread t < C: top-row candidate, t < 2C: bottom-row candidate, then two per owned row
// (left / right column).  Also seeds taint (edge_todo as float, 944).
// LEGACY: proportions / taint live in the Cell records of the work-list sweep; else the proportion
// plane is read and only the mask is written (the tile sweep seeds taint from the mask).
template <bool LEGACY>
__global__ void __launch_bounds__(256)
k_border_todo(const double *__restrict__ E, const uint8_t *__restrict__ link, Cell *__restrict__ cell,
              const double *__restrict__ prop, Win w, const int32_t *__restrict__ pit_beg, const int32_t *__restrict__ pit_end, const double *__restrict__ pit_w,
              const int32_t *__restrict__ pit_dst, int64_t n_pit_edges,
              uint8_t *__restrict__ edge_todo, unsigned long long *counters)
{
    const int64_t C = w.C;
    const int64_t nown = w.hi - w.lo;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * C + 2 * nown) return;
    int64_t i, j;
    if (t < C) { i = -w.row_off; j = t; }                                   // global row 0
    else if (t < 2 * C) { i = w.Rg - 1 - w.row_off; j = t - C; }            // global last row
    else {
        const int64_t u = t - 2 * C;
        i = w.lo + (u >> 1); j = (u & 1) ? C - 1 : 0;
        if (w.top(i) || w.bottom(i)) return;                                // corners belong to the row threads
    }
    if (i < w.lo || i >= w.hi) return;                                      // that border row lives on another rank
    const int64_t n = i * C + j;
    const uint8_t lk = link[n];
    const double TOL = 1e-2;
    // column sum of A = total outflow weight that survived the filter (913-919)
    double outflow = 0.0;
    int sec = (lk & LK_NOSEC) ? -1 : (lk & LK_SEC_MASK);
    if (lk & LK_PIT) {
        const int64_t slot = __double_as_longlong(LEGACY ? cell[n].prop : prop[n]);
        for (int32_t e = pit_beg[slot]; e < pit_end[slot]; e++) outflow += pit_w[e];
    } else {
        const double p = LEGACY ? cell[n].prop : prop[n];
        // scipy sums a column's entries in row-index order; two terms commute
        if (lk & LK_KEEP1) outflow += p;
        if (lk & LK_KEEP2) outflow += __dsub_rn(1.0, p);
    }
    bool todo = false;
    const bool big = outflow > TOL;
    const bool top = w.top(i), bot = w.bottom(i);
    if (j == 0) todo = big && sec_in(sec, 0xC3u);                           // left: 6,7,0,1
    if (j == C - 1) todo = big && sec_in(sec, 0x3Cu);                       // right: 2,3,4,5 (later assignment wins)
    if (top) todo = big && sec_in(sec, 0xF0u);                              // top: 4,5,6,7
    if (bot) todo = big && sec_in(sec, 0x0Fu);                              // bottom: 0,1,2,3
    if ((top || bot) && (j == 0 || j == C - 1)) {
        // corners 924-930: |= outflow > TOL  |  inflow < TOL (row sum of A)
        double inflow = 0.0;
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++) {
                const int64_t mi = i + di, mj = j + dj;
                if ((di == 0 && dj == 0) || !w.row_in_grid(mi) || mj < 0 || mj >= C) continue;
                const uint8_t ml = link[mi * C + mj];
                if (ml & (LK_NOSEC | LK_PIT)) continue;
                const int ms = ml & LK_SEC_MASK;
                const double mp = LEGACY ? cell[mi * C + mj].prop : prop[mi * C + mj];
                if ((ml & LK_KEEP1) && g_e1r[ms] == -di && g_e1c[ms] == -dj) inflow += mp;
                if ((ml & LK_KEEP2) && g_e2r[ms] == -di && g_e2c[ms] == -dj) inflow += __dsub_rn(1.0, mp);
            }
        for (int64_t e = 0; e < n_pit_edges; e++)
            if (pit_dst[e] == (int32_t)n) inflow += pit_w[e];
        todo = todo || big || (inflow < TOL);
    }
    const double e0 = E[n];
    if (e0 != e0) todo = false;                                             // 935
    edge_todo[n] = todo ? 1 : 0;
    if (todo) { if (LEGACY) cell[n].taint = 1.0; atomicAdd(&counters[CT_EDGE_TODO], 1ULL); }
}

}  // namespace


// section/proportion + receivers + filter (k_links), then the pit drains.  Leaves indeg holding
// only the pit in-edges.  Shared by the full sweep and the update mode (the reference rebuilds
// both on every calc_uca call, dem_processing.py:787-793 / 873-879).
int pdm_graph_links_pits(pdm_tile *t, const pdm_uca_params *p)
{
    const Win &w = t->win;
    dim3 block(32, 8);
    dim3 grid((unsigned)((w.C + 31) / 32), (unsigned)((w.hi - w.lo + 7) / 8));
    // stage counters only: the work-list counters below CT_SOURCES remember which queue slots are dirty
    PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_SOURCES, 0, (CT_N - CT_SOURCES) * sizeof(unsigned long long), t->stream));
    // proportion goes to scratch (the TWI buffer is free until calc_twi); k_indeg / the update
    // mode assemble the 32-byte sweep records from it
    k_links<<<grid, block, 0, t->stream>>>(t->elev, t->dir, t->flats, t->th_row, w, t->link, t->twi,
                                           t->flat0, t->d_counters);
    PDM_LAUNCHED();
    t->n_pits = 0; t->n_pit_edges = 0;
    t->shard_pits_wanted = false; t->shard_pits_done = false;
    if (p->drain_pits) {
        if (w.lo != 0 || w.hi != t->R || w.Rg != t->R) {
            // row shard: the search needs the neighbours' strips of elevation and pit mask -> pdm_shard_pits (shard.cu)
            t->shard_pits_wanted = true;
            return PDM_OK;
        }
        int rc = pdm_launch_pits(t, p);
        if (rc) return rc;
    }
    return PDM_OK;
}

// in-degree + source flags + sweep state + inflow-border mask (after the neighbours' link rows
// are in place on a shard)
int pdm_launch_indeg_todo(pdm_tile *t)
{
    const Win &w = t->win;
    dim3 block(32, 8);
    dim3 grid((unsigned)((w.C + 31) / 32), (unsigned)((t->R + 7) / 8));
    PDM_CUDA(cudaMemsetAsync(t->edge_todo, 0, (size_t)t->N, t->stream));
    k_indeg<<<grid, block, 0, t->stream>>>(t->link, w, t->row_area, t->twi, (t->n_pits || t->shard_pits_done) ? t->label : nullptr, t->cell,
                                           t->d_counters);
    PDM_LAUNCHED();
    const int64_t per = 2 * w.C + 2 * (w.hi - w.lo);
    k_border_todo<true><<<(unsigned)((per + 255) / 256), 256, 0, t->stream>>>(
        t->elev, t->link, t->cell, t->twi, w, t->pit_beg, t->pit_end, t->pit_w, t->pit_dst, t->n_pit_edges,
        t->edge_todo, t->d_counters);
    PDM_LAUNCHED();
    t->legacy_graph = true;
    return PDM_OK;
}

// tile sweep: only the inflow-border mask is needed (in-degrees are counted per tile visit from the
// neighbours' link bytes; the neighbours' link AND proportion halo rows must be in place on a shard)
int pdm_launch_border_todo(pdm_tile *t)
{
    const Win &w = t->win;
    PDM_CUDA(cudaMemsetAsync(t->edge_todo, 0, (size_t)t->N, t->stream));
    const int64_t per = 2 * w.C + 2 * (w.hi - w.lo);
    k_border_todo<false><<<(unsigned)((per + 255) / 256), 256, 0, t->stream>>>(
        t->elev, t->link, nullptr, t->twi, w, t->pit_beg, t->pit_end, t->pit_w, t->pit_dst, t->n_pit_edges,
        t->edge_todo, t->d_counters);
    PDM_LAUNCHED();
    t->legacy_graph = false;
    return PDM_OK;
}

int pdm_launch_graph(pdm_tile *t, const pdm_uca_params *p, pdm_uca_stats *st)
{
    (void)st;
    int rc = pdm_graph_links_pits(t, p);
    if (rc) return rc;
    // circular_ref_maxcount <= 1: the reference's sweep loop never runs (see pdm_launch_sweep_full): the epilogue
    // needs the initial records of the work-list graph
    return (pdm_sweep_legacy() || p->circular_ref_maxcount <= 1) ? pdm_launch_indeg_todo(t) : pdm_launch_border_todo(t);
}

int pdm_launch_section_export(pdm_tile *t)
{
    if (!t->section) PDM_CUDA(cudaMalloc(&t->section, (size_t)t->N));
    dim3 block(32, 8);
    dim3 grid((unsigned)((t->C + 31) / 32), (unsigned)((t->R + 7) / 8));
    k_section_export<<<grid, block, 0, t->stream>>>(t->dir, t->flats, t->th_row, t->R, t->C, t->section);
    PDM_LAUNCHED();
    return PDM_OK;
}
