// update.cu -- a8: edge-update mode of calc_uca.
//
// Reference behaviour: calc_uca edge packing (dem_processing.py:719-744, 769-771) and
// _calc_uca_chunk_update (778-862) with cyutils.drain_connections (cyutils.pyx:35-72) and
// cyutils.drain_area(skip_edge=False).
//
// Given the neighbours' copies of the four border strips {uca value, done, todo}:
//   start   = done & todo            (my inflow cells whose upstream value became final)
//   delta   = value - uca  on border cells with done, 0 elsewhere, NaN on flats
//   cone    = everything downstream of start            (flood 1; also counts, per cone cell,
//                                                         its in-edges from start/cone cells)
//   sweep   = push delta through the cone in dependency order, never into a start cell
//   reach   = downstream closure of the remaining todo cells (flood 2) -> edge_done = ~reach
//   uca    += delta
// All three traversals run on the barrier-free work-list engine (worklist.cuh); the
// reference rebuilds the sparse matrix and rescans all cells every round.
#include "drain_op.cuh"

namespace {

// sweep records of the update mode: .area is the delta (0, NaN on flats: 802, 815), .indeg the
// restricted in-degree counted by flood 1
__global__ void __launch_bounds__(256)
k_upd_init(const uint8_t *__restrict__ flats, const uint8_t *__restrict__ link, const double *__restrict__ prop,
           int64_t N, Cell *__restrict__ cell, uint8_t *__restrict__ st, uint8_t *__restrict__ edge_todo)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    Cell rec;
    rec.area = flats[n] ? __longlong_as_double(0x7ff8000000000000LL) : 0.0;
    rec.taint = 0.0; rec.prop = prop[n]; rec.indeg = 0; rec.link = link[n];
    rec.pad[0] = rec.pad[1] = rec.pad[2] = 0;
    cell[n] = rec;
    st[n] = 0;
    edge_todo[n] = 0;
}

// strips layout in the staging buffers: [left R][right R][top C][bottom C]
__global__ void __launch_bounds__(256)
k_upd_border(const double *__restrict__ sdata, const uint8_t *__restrict__ sdone, const uint8_t *__restrict__ stodo,
             const double *__restrict__ uca0, const uint8_t *__restrict__ flats, int64_t R, int64_t C,
             Cell *__restrict__ cell, uint8_t *__restrict__ st, uint8_t *__restrict__ edge_todo,
             unsigned long long *ctr)
{
    const int64_t per = 2 * C + 2 * (R - 2);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= per) return;
    int64_t i, j;
    if (t < C) { i = 0; j = t; }
    else if (t < 2 * C) { i = R - 1; j = t - C; }
    else { const int64_t u = t - 2 * C; i = 1 + (u >> 1); j = (u & 1) ? C - 1 : 0; }
    const int64_t n = i * C + j;
    // 730-737 in the reference's key order left, right, top, bottom; corners see two strips
    bool done = false, todo = false;
    double init = 0.0;
    if (j == 0)     { const int64_t k = i;             done |= sdone[k] != 0; init = __dadd_rn(init, __dmul_rn(sdata[k], (double)(sdone[k] != 0))); todo |= stodo[k] != 0; }
    if (j == C - 1) { const int64_t k = R + i;         done |= sdone[k] != 0; init = __dadd_rn(init, __dmul_rn(sdata[k], (double)(sdone[k] != 0))); todo |= stodo[k] != 0; }
    if (i == 0)     { const int64_t k = 2 * R + j;     done |= sdone[k] != 0; init = __dadd_rn(init, __dmul_rn(sdata[k], (double)(sdone[k] != 0))); todo |= stodo[k] != 0; }
    if (i == R - 1) { const int64_t k = 2 * R + C + j; done |= sdone[k] != 0; init = __dadd_rn(init, __dmul_rn(sdata[k], (double)(sdone[k] != 0))); todo |= stodo[k] != 0; }
    if (!done) init = 0.0;                                                   // 738-739
    const bool start = done && todo;                                         // 798
    todo = todo && !done;                                                    // 799
    if (done && !flats[n]) cell[n].area = __dsub_rn(init, uca0[n]);         // 806-809 (flats stay NaN, 815)
    st[n] = (start ? ST_START : 0) | (todo ? (ST_TODOSEED | ST_REACH) : 0);
    edge_todo[n] = todo ? 1 : 0;                                             // 817 (edge_todo_i)
    if (start) atomicAdd(&ctr[CT_SOURCES], 1ULL);
    if (todo) atomicAdd(&ctr[CT_EDGE_TODO], 1ULL);
}

// flood over the drainage graph: visit every cell downstream of the seeds once.
//   WHICH 0: seeds ST_START, marks ST_CONE, skips start cells, counts restricted in-degree
//   WHICH 1: seeds ST_TODOSEED, marks ST_REACH
template <int WHICH>
struct FloodOp {
    static constexpr bool P2P = false;
    Cell *cell;
    uint8_t *st;
    int32_t C;
    const int32_t *pit_beg;
    const int32_t *pit_end;
    const int32_t *pit_dst;

    __device__ __forceinline__ bool is_seed(int32_t c) const
    {
        return (st[c] & (WHICH == 0 ? ST_START : ST_TODOSEED)) != 0;
    }
    // returns true when r was reached for the first time
    __device__ __forceinline__ bool visit(int32_t r) const
    {
        if (WHICH == 0) {
            if (st[r] & ST_START) return false;  // start bits are fixed before the flood
            atomicAdd(&cell[r].indeg, 1);
            return !(st_fetch_or(st, r, ST_CONE) & ST_CONE);
        }
        return !(st_fetch_or(st, r, ST_REACH) & ST_REACH);
    }
    // single-lane chain (worklist.cuh): runs until the chain ends or forks
    __device__ __forceinline__ unsigned long long chain(int32_t &cur, int32_t &other, const wl::Queue &q) const
    {
        unsigned long long n = 0;
        int32_t i = cur;
        while (i >= 0 && other < 0) { n++; i = process(i, q, other); }
        cur = i;
        return n;
    }
    __device__ __forceinline__ int32_t chase_offset(uint8_t) const { return 0; }
    // (no warp-wide bursts for the floods: the holder lane runs the ordinary chain)
    __device__ __forceinline__ unsigned long long chain_warp(int32_t &cur, int32_t &other, const wl::Queue &q, int holder, int32_t *, int &, int &, const int32_t *) const
    {
        return (int)(threadIdx.x & 31) == holder ? chain(cur, other, q) : 0ULL;
    }
    __device__ __forceinline__ int32_t process(int32_t i, const wl::Queue &q, int32_t &defer) const
    {
        const longlong2 pm = *reinterpret_cast<const longlong2 *>(&cell[i].prop);   // prop | indeg, link (link, prop fixed)
        const uint8_t lk = (uint8_t)((unsigned long long)pm.y >> 32);
        int32_t nxt = -1;
        if (lk & LK_PIT) {
            const int64_t slot = pm.x;
            for (int32_t e = pit_beg[slot]; e < pit_end[slot]; e++) {
                const int32_t r = pit_dst[e];
                if (visit(r)) { if (nxt < 0) nxt = r; else q.push(r); }
            }
            return nxt;
        }
        const int sec = lk & LK_SEC_MASK;
        if (lk & LK_KEEP1) {
            const int32_t r = i + wl::off_e1(sec, C);
            if (visit(r)) nxt = r;
        }
        if (lk & LK_KEEP2) {
            const int32_t r = i + wl::off_e2(sec, C);
            if (visit(r)) { if (nxt < 0) nxt = r; else defer = r; }
        }
        return nxt;
    }
};

__global__ void __launch_bounds__(256)
k_upd_finalize(const Cell *__restrict__ cell, const uint8_t *__restrict__ st, int64_t N,
               double *__restrict__ uca, uint8_t *__restrict__ edge_done)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    uca[n] = __dadd_rn(uca[n], cell[n].area);                                // 769
    edge_done[n] = (st[n] & ST_REACH) ? 0 : 1;                               // 856
}

}  // namespace


static int g_blocks_flood0 = 0, g_blocks_flood1 = 0, g_blocks_drain1 = 0;

int pdm_launch_update(pdm_tile *t, const pdm_uca_params *p, const double *const data[4],
                      const uint8_t *const done[4], const uint8_t *const todo[4], pdm_uca_stats *st)
{
    const int64_t R = t->R, C = t->C, per_all = 2 * R + 2 * C;
    int rc;
    if (!g_blocks_flood0) {
        if ((rc = wl::grid_for(wl::k_worklist<FloodOp<0>, wl::DomainBorder>, &g_blocks_flood0))) return rc;
        if ((rc = wl::grid_for(wl::k_worklist<FloodOp<1>, wl::DomainBorder>, &g_blocks_flood1))) return rc;
        if ((rc = wl::grid_for(wl::k_worklist<DrainOp<1>, wl::DomainBorder>, &g_blocks_drain1))) return rc;
    }
    // the reference rebuilds section/proportion and the matrix on every call (787-793).  A tile that stays
    // resident between the corrections of a mosaic (pdm_tile_set_keep_graph) keeps the graph of its full sweep:
    // link bytes, proportions and pit edge lists are not touched by an update, and the pits must NOT be searched
    // again -- their drains have already patched mag / flats in place (1370-1371)
    PDM_CUDA(cudaEventRecord(t->ev[0], t->stream));
    if (!(t->keep_graph && t->have_graph && t->legacy_graph) && (rc = pdm_graph_links_pits(t, p))) return rc;   // (the tile sweep recycles the proportion plane)
    PDM_CUDA(cudaEventRecord(t->ev[1], t->stream));
    // stage the strips
    if (!t->edge_buf_d) {
        PDM_CUDA(cudaMalloc(&t->edge_buf_d, (size_t)per_all * 8));
        PDM_CUDA(cudaMalloc(&t->edge_buf_b, (size_t)per_all * 2));
    }
    const int64_t len[4] = {R, R, C, C};
    int64_t off = 0;
    for (int k = 0; k < 4; k++) {
        PDM_CUDA(cudaMemcpyAsync(t->edge_buf_d + off, data[k], (size_t)len[k] * 8, cudaMemcpyHostToDevice, t->stream));
        PDM_CUDA(cudaMemcpyAsync(t->edge_buf_b + off, done[k], (size_t)len[k], cudaMemcpyHostToDevice, t->stream));
        PDM_CUDA(cudaMemcpyAsync(t->edge_buf_b + per_all + off, todo[k], (size_t)len[k], cudaMemcpyHostToDevice, t->stream));
        off += len[k];
    }
    uint8_t *stt = t->flat0;
    PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_SOURCES, 0, sizeof(unsigned long long), t->stream));
    PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_EDGE_TODO, 0, sizeof(unsigned long long), t->stream));
    k_upd_init<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>(t->flats, t->link, t->twi, t->N, t->cell, stt, t->edge_todo);
    PDM_LAUNCHED();
    const int64_t per = 2 * C + 2 * (R - 2);
    k_upd_border<<<(unsigned)((per + 255) / 256), 256, 0, t->stream>>>(t->edge_buf_d, t->edge_buf_b, t->edge_buf_b + per_all,
                                                                       t->uca, t->flats, R, C, t->cell, stt, t->edge_todo,
                                                                       t->d_counters);
    PDM_LAUNCHED();
    const wl::Queue q{t->queue, t->d_counters, (long long)t->N, t->d_counters + CT_SOURCES};       // seeds: start cells
    const wl::Queue q2{t->queue, t->d_counters, (long long)t->N, t->d_counters + CT_EDGE_TODO};    // seeds: remaining todo cells
    const wl::DomainBorder dom{R, C};
    // flood 1: cone + restricted in-degree (820-831)
    if ((rc = wl::reset_queue(t))) return rc;
    wl::k_worklist<<<g_blocks_flood0, 256, 0, t->stream>>>(
        FloodOp<0>{t->cell, stt, (int32_t)C, t->pit_beg, t->pit_end, t->pit_dst}, dom, q);
    PDM_LAUNCHED();
    // sweep of the deltas (836-842)
    if ((rc = wl::reset_queue(t))) return rc;
    wl::k_worklist<<<g_blocks_drain1, 256, 0, t->stream>>>(
        DrainOp<1>{t->link, t->cell, stt, (int32_t)C, t->pit_beg, t->pit_end, t->pit_dst, t->pit_w, 0},
        dom, q);
    PDM_LAUNCHED();
    PDM_CUDA(cudaMemcpyAsync(t->h_counters + CT_DRAINED, t->d_counters + CT_DRAINED, sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, t->stream));
    // flood 2: remaining todo cells taint everything downstream (848-853)
    if ((rc = wl::reset_queue(t))) return rc;
    wl::k_worklist<<<g_blocks_flood1, 256, 0, t->stream>>>(
        FloodOp<1>{t->cell, stt, (int32_t)C, t->pit_beg, t->pit_end, t->pit_dst}, dom, q2);
    PDM_LAUNCHED();
    k_upd_finalize<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>(t->cell, stt, t->N, t->uca, t->edge_done);
    PDM_LAUNCHED();
    PDM_CUDA(cudaEventRecord(t->ev[2], t->stream));
    PDM_CUDA(cudaMemcpyAsync(t->h_counters + CT_SOURCES, t->d_counters + CT_SOURCES, sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaMemcpyAsync(t->h_counters + CT_EDGE_TODO, t->d_counters + CT_EDGE_TODO, sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaMemcpyAsync(t->h_counters + CT_WATCHDOG, t->d_counters + CT_WATCHDOG, sizeof(unsigned long long),
                             cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    if (t->h_counters[CT_WATCHDOG]) { pdm_set_error("pdm_tile_uca_update: work-list watchdog fired"); return PDM_ERR_STATE; }
    if (st) {
        st->n_sources = (int64_t)t->h_counters[CT_SOURCES];
        st->n_drained = (int64_t)t->h_counters[CT_DRAINED];
        st->n_edge_todo = (int64_t)t->h_counters[CT_EDGE_TODO];
        st->n_pits = t->n_pits; st->n_pit_edges = t->n_pit_edges;
        cudaEventElapsedTime(&st->ms_graph, t->ev[0], t->ev[1]);
        cudaEventElapsedTime(&st->ms_sweep, t->ev[1], t->ev[2]);
        cudaEventElapsedTime(&st->ms_total, t->ev[0], t->ev[2]);
    }
    return PDM_OK;
}

// Circular references: the reference's restart loop (dem_processing.py:951-964) is unreachable -- the drainage
// graph is acyclic by construction (sweep.cu explains) -- so cells left over after a full sweep mean a broken
// graph, not a case to replay: fail loudly.
int pdm_restart_rounds(pdm_tile *t, const pdm_uca_params *p, pdm_uca_stats *st)
{
    (void)t;
    st->n_restarts = 0;
    if (p->circular_ref_maxcount <= 1) return PDM_OK;     // no sweep was asked for (the loop condition of 951-952)
    pdm_set_error("pdm_tile_uca: %lld cells were not drained by a full sweep; the drainage graph of a DEM is acyclic "
                  "(elev[j] <= elev[i] on every edge, zero weight between equal elevations), so this is an internal error",
                  (long long)st->n_undone);
    return PDM_ERR_STATE;
}
