// shard.cu -- row-sharded execution of the hot path: one tile per GPU holds a block of rows of
// one big DEM plus one halo row per interior side.  The host driver (pydem_b200/sharded.py)
// moves halo rows between neighbouring ranks with NCCL send/recv and calls the stages below.
//
// UCA across shards is the same dependency-ordered accumulation as on one tile.  A cell whose
// receiver lives on the neighbouring rank pushes into the halo row, which acts as an out-box:
// area and taint accumulate there and the in-degree counter counts the decrements (it goes
// negative).  When the local work-list is quiescent the out-boxes are exchanged, added into the
// neighbour's boundary row, and the cells whose in-degree reached zero seed the next local pass.
// The loop ends when no rank sent anything: #rounds = 1 + the largest number of shard boundaries
// any flow path crosses.  Because every owned cell sees its true 3x3 neighbourhood and "border"
// means the border of the global grid, the result equals the single-tile result (this is the
// gating-free form of pyDEM's cross-tile edge resolution, process_manager.py:1090-1249, where
// tiles re-run calc_uca on edge deltas until nothing changes).
#include <string.h>

#include "pdm_internal.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_outbox_pack(Cell *cell, int64_t row, int64_t C,
              double *__restrict__ out_a, double *__restrict__ out_t, int32_t *__restrict__ out_c, long long *nonzero)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool nz = false;
    if (j < C) {
        Cell &x = cell[row * C + j];
        const int32_t c = -x.indeg;
        out_a[j] = x.area; out_t[j] = x.taint; out_c[j] = c;
        x.area = 0.0; x.taint = 0.0; x.indeg = 0;
        nz = c != 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, nz);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd((unsigned long long *)nonzero, (unsigned long long)__popc(m));
}

__global__ void __launch_bounds__(256)
k_inbox_apply(Cell *cell, int64_t row, int64_t C,
              const double *__restrict__ in_a, const double *__restrict__ in_t, const int32_t *__restrict__ in_c,
              int32_t *__restrict__ seeds, unsigned long long *ctr)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= C) return;
    const int32_t c = in_c[j];
    if (c == 0) return;
    const int64_t n = row * C + j;
    Cell &x = cell[n];
    x.area = __dadd_rn(x.area, in_a[j]);
    x.taint = __dadd_rn(x.taint, in_t[j]);
    const int32_t left = x.indeg - c;
    x.indeg = left;
    if (left == 0) seeds[atomicAdd(&ctr[CT_TMP1], 1ULL)] = (int32_t)n;
}

}  // namespace

int pdm_launch_ccl(pdm_tile *t);
int pdm_launch_flats_extend(pdm_tile *t);
int pdm_launch_label_pack(pdm_tile *t, int64_t row, long long *out_l, double *out_e);
int pdm_launch_label_unpack(pdm_tile *t, int64_t row, const long long *in_l, const double *in_e);
int pdm_launch_indeg_todo(pdm_tile *t);
int pdm_launch_sweep_first(pdm_tile *t);
int pdm_launch_sweep_resume(pdm_tile *t);
int pdm_launch_uca_finalize(pdm_tile *t, const pdm_uca_params *p);

static int read_ctr(pdm_tile *t)
{
    PDM_CUDA(cudaMemcpyAsync(t->h_counters, t->d_counters, CT_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    return PDM_OK;
}

extern "C" {

int pdm_tile_set_window(pdm_tile *t, int64_t row_off, int64_t R_global, int64_t own_lo, int64_t own_hi,
                        const double *th_row)
{
    if (!t) { pdm_set_error("NULL tile"); return PDM_ERR_ARG; }
    if (own_lo < 0 || own_hi > t->R || own_hi - own_lo < 2 || row_off + own_lo < 0 || row_off + own_hi > R_global ||
        (own_lo > 0) != (row_off + own_lo > 0) || (own_hi < t->R) != (row_off + own_hi < R_global) || own_lo > 1 ||
        t->R - own_hi > 1) {
        pdm_set_error("pdm_tile_set_window: need >= 2 owned rows, exactly one halo row on every side that has a "
                      "neighbour and none on the grid border (R=%lld own=[%lld,%lld) row_off=%lld Rg=%lld)",
                      (long long)t->R, (long long)own_lo, (long long)own_hi, (long long)row_off, (long long)R_global);
        return PDM_ERR_ARG;
    }
    t->win.row_off = row_off; t->win.Rg = R_global; t->win.lo = own_lo; t->win.hi = own_hi;
    if (th_row) PDM_CUDA(cudaMemcpyAsync(t->th_row, th_row, (size_t)t->R * 8, cudaMemcpyHostToDevice, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    t->have_graph = false;
    return PDM_OK;
}

// a1 on the owned rows (elev halo rows must be in place).  Out: MAG, DIR, flat0.
int pdm_shard_slopes(pdm_tile *t)
{
    if (!t || !t->have_elev || !t->have_spacing) { pdm_set_error("pdm_shard_slopes: ELEV / spacing missing"); return PDM_ERR_STATE; }
    int rc = pdm_launch_slopes(t);
    if (!rc) t->have_slopes = true;
    return rc;
}

// region labelling over all local rows (flat0 halo rows must be in place)
int pdm_shard_ccl(pdm_tile *t)
{
    if (!t || !t->have_slopes) { pdm_set_error("pdm_shard_ccl: run pdm_shard_slopes first"); return PDM_ERR_STATE; }
    return pdm_launch_ccl(t);
}

int pdm_shard_label_pack(pdm_tile *t, int64_t row, void *out_labels, void *out_elev)
{
    if (!t || !t->glabel || row < 0 || row >= t->R) { pdm_set_error("pdm_shard_label_pack: bad state/row"); return PDM_ERR_ARG; }
    return pdm_launch_label_pack(t, row, (long long *)out_labels, (double *)out_elev);
}

// changed (device pointer, int64) accumulates the number of regions whose label went down
int pdm_shard_label_unpack(pdm_tile *t, int64_t row, const void *in_labels, const void *in_elev, void *changed)
{
    if (!t || !t->glabel || row < 0 || row >= t->R) { pdm_set_error("pdm_shard_label_unpack: bad state/row"); return PDM_ERR_ARG; }
    PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_FLAG, 0, sizeof(unsigned long long), t->stream));
    int rc = pdm_launch_label_unpack(t, row, (const long long *)in_labels, (const double *)in_elev);
    if (rc) return rc;
    if (changed) {
        rc = read_ctr(t);
        if (rc) return rc;
        long long add = (long long)t->h_counters[CT_FLAG], cur = 0;
        PDM_CUDA(cudaMemcpy(&cur, changed, sizeof(cur), cudaMemcpyDeviceToHost));
        cur += add;
        PDM_CUDA(cudaMemcpy(changed, &cur, sizeof(cur), cudaMemcpyHostToDevice));
    }
    return PDM_OK;
}

int pdm_shard_flats_extend(pdm_tile *t)
{
    if (!t || !t->have_slopes) { pdm_set_error("pdm_shard_flats_extend: bad state"); return PDM_ERR_STATE; }
    int rc = pdm_launch_flats_extend(t);
    if (!rc) t->have_flats = true;
    return rc;
}

// a3/a4 on the owned rows.  Out: link bytes + proportion (send the boundary link rows next).
int pdm_shard_links(pdm_tile *t, const pdm_uca_params *p_in)
{
    if (!t || !t->have_flats || !t->have_slopes) { pdm_set_error("pdm_shard_links: bad state"); return PDM_ERR_STATE; }
    pdm_uca_params p;
    if (p_in) p = *p_in; else { pdm_default_uca_params(&p); p.drain_pits = 0; }
    return pdm_graph_links_pits(t, &p);
}

// in-degree, sources, sweep state, inflow-border mask (link halo rows must be in place)
int pdm_shard_indeg(pdm_tile *t)
{
    if (!t) return PDM_ERR_ARG;
    return pdm_launch_indeg_todo(t);
}

// one local accumulation pass to quiescence; first != 0: seeds are the sources
int pdm_shard_sweep(pdm_tile *t, int first)
{
    if (!t) return PDM_ERR_ARG;
    return first ? pdm_launch_sweep_first(t) : pdm_launch_sweep_resume(t);
}

// side 0: out-box toward the rank above (halo row lo-1), side 1: below (halo row hi).
// nonzero (device int64) += number of cells with something to deliver.
int pdm_shard_outbox_pack(pdm_tile *t, int side, void *out_area, void *out_taint, void *out_count, void *nonzero)
{
    if (!t) return PDM_ERR_ARG;
    const Win &w = t->win;
    const int64_t row = side == 0 ? w.lo - 1 : w.hi;
    if (row < 0 || row >= t->R) { pdm_set_error("pdm_shard_outbox_pack: no halo row on side %d", side); return PDM_ERR_ARG; }
    k_outbox_pack<<<(unsigned)((t->C + 255) / 256), 256, 0, t->stream>>>(t->cell, row, t->C,
                                                                         (double *)out_area, (double *)out_taint,
                                                                         (int32_t *)out_count, (long long *)nonzero);
    PDM_LAUNCHED();
    return PDM_OK;
}

// reset the seed list before the in-boxes of a round are applied
int pdm_shard_inbox_begin(pdm_tile *t)
{
    if (!t) return PDM_ERR_ARG;
    PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_TMP1, 0, sizeof(unsigned long long), t->stream));
    return PDM_OK;
}

// side 0: what the rank above delivered to my first owned row, side 1: below -> last owned row
int pdm_shard_inbox_apply(pdm_tile *t, int side, const void *in_area, const void *in_taint, const void *in_count)
{
    if (!t) return PDM_ERR_ARG;
    const Win &w = t->win;
    const int64_t row = side == 0 ? w.lo : w.hi - 1;
    k_inbox_apply<<<(unsigned)((t->C + 255) / 256), 256, 0, t->stream>>>(t->cell, row, t->C,
                                                                         (const double *)in_area, (const double *)in_taint,
                                                                         (const int32_t *)in_count, t->label, t->d_counters);
    PDM_LAUNCHED();
    return PDM_OK;
}

int pdm_shard_finalize(pdm_tile *t, const pdm_uca_params *p_in, pdm_uca_stats *stats)
{
    if (!t) return PDM_ERR_ARG;
    pdm_uca_params p;
    if (p_in) p = *p_in; else pdm_default_uca_params(&p);
    int rc = pdm_launch_uca_finalize(t, &p);
    if (rc) return rc;
    rc = read_ctr(t);
    if (rc) return rc;
    if (t->h_counters[CT_WATCHDOG]) {
        pdm_set_error("pdm_shard_finalize: work-list watchdog fired (QTAIL=%llu QHEAD=%llu QDONE=%llu PHASE1=%llu DRAINED=%llu)",
                      t->h_counters[CT_QTAIL], t->h_counters[CT_QHEAD], t->h_counters[CT_QDONE], t->h_counters[CT_PHASE1],
                      t->h_counters[CT_DRAINED]);
        return PDM_ERR_STATE;
    }
    t->have_uca = true; t->have_graph = true;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_cells = (t->win.hi - t->win.lo) * t->C;
        stats->n_sources = (int64_t)t->h_counters[CT_SOURCES];
        stats->n_drained = (int64_t)t->h_counters[CT_DRAINED];
        stats->n_undone = (int64_t)t->h_counters[CT_UNDONE];
        stats->n_edge_todo = (int64_t)t->h_counters[CT_EDGE_TODO];
        stats->min_area = t->min_area;
    }
    return PDM_OK;
}

}  // extern "C"
