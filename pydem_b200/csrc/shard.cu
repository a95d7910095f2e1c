// shard.cu -- row-sharded execution of the hot path: one tile per GPU holds a block of rows of
// one big DEM plus one halo row per interior side.  The host driver (pydem_b200/sharded.py)
// moves halo rows between neighbouring ranks with NCCL send/recv and calls the stages below.
//
// UCA across shards is the same tile-resident, pull-based accumulation as on one tile (tsweep.cu).
// Tiles cover the owned rows; the halo rows are ring cells: a boundary cell pulls the contributions
// of its donors on the neighbouring rank from the halo row once they are final there.  After a
// local sweep has come to rest, every rank sends its boundary rows of UCA / taint (final values, or
// the "not done" pattern) into the neighbours' halo rows and resumes the boundary tiles whose ring
// received new donors.  The loop ends when no rank completed a boundary cell with a receiver across
// the boundary: #rounds = 1 + the largest number of shard boundaries any flow path crosses.
// Because every owned cell sees its true 3x3 neighbourhood and "border" means the border of the
// global grid, the result equals the single-tile result -- bit for bit, since a cell's sum has a
// fixed order (this is the gating-free form of pyDEM's cross-tile edge resolution,
// process_manager.py:1090-1249, where tiles re-run calc_uca on edge deltas until nothing changes).
#include <string.h>

#include "pdm_internal.cuh"

#include "tsweep.cuh"

namespace {

// pit edges that start on a neighbouring rank and end in my rows next to the boundary: add their counts to the
// per-cell pit in-degree (t->label) before the in-degrees are built
__global__ void __launch_bounds__(256)
k_pit_in_apply(int32_t *__restrict__ pit_in, const int32_t *__restrict__ strip, int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { const int32_t v = strip[k]; if (v) pit_in[k] += v; }
}

__global__ void k_add_sent(const unsigned long long *ctr, long long *out) { *out += (long long)ctr[ts::TC_SENT]; }
__global__ void k_add_counter(const unsigned long long *ctr, long long *out) { *out += (long long)*ctr; }

}  // namespace

int pdm_launch_ccl(pdm_tile *t);
int pdm_launch_flats_extend(pdm_tile *t);
int pdm_launch_label_pack(pdm_tile *t, int64_t row, long long *out_l, double *out_e);
int pdm_launch_label_unpack(pdm_tile *t, int64_t row, const long long *in_l, const double *in_e);
int pdm_ts_p2p_export(pdm_tile *t, ts::P2PExport *e);
int pdm_ts_p2p_connect(pdm_tile *t, const ts::P2PExport *up, const ts::P2PExport *down, const ts::P2PExport *root, int world, int rank);
int pdm_ts_p2p_connect_ranks(pdm_tile *t, const ts::P2PExport *const *ex, int world, int rank);

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

// changed (device pointer, int64) accumulates the number of regions whose label went down (on the device: no host
// synchronisation; the driver all-reduces it)
int pdm_shard_label_unpack(pdm_tile *t, int64_t row, const void *in_labels, const void *in_elev, void *changed)
{
    if (!t || !t->glabel || row < 0 || row >= t->R) { pdm_set_error("pdm_shard_label_unpack: bad state/row"); return PDM_ERR_ARG; }
    PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_FLAG, 0, sizeof(unsigned long long), t->stream));
    int rc = pdm_launch_label_unpack(t, row, (const long long *)in_labels, (const double *)in_elev);
    if (rc) return rc;
    if (changed) {
        k_add_counter<<<1, 1, 0, t->stream>>>(t->d_counters + CT_FLAG, (long long *)changed);
        PDM_LAUNCHED();
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

// fences of the whole grid (host arrays of n = R_global - 1 doubles): pit drains of a row shard reach across
// the shard boundary, their distances (_get_dX_mean / make_slice, dem_processing.py:1346-1349) need rows this rank does not own
int pdm_tile_set_global_spacing(pdm_tile *t, const double *dX, const double *dY, int64_t n)
{
    if (!t || !dX || !dY || n != t->win.Rg - 1) { pdm_set_error("pdm_tile_set_global_spacing: need R_global - 1 = %lld fences (set the window first)", t ? (long long)(t->win.Rg - 1) : 0LL); return PDM_ERR_ARG; }
    if (t->dXg) { cudaFree(t->dXg); cudaFree(t->dYg); t->dXg = t->dYg = nullptr; }
    PDM_CUDA(cudaMalloc(&t->dXg, (size_t)n * 8));
    PDM_CUDA(cudaMalloc(&t->dYg, (size_t)n * 8));
    PDM_CUDA(cudaMemcpyAsync(t->dXg, dX, (size_t)n * 8, cudaMemcpyHostToDevice, t->stream));
    PDM_CUDA(cudaMemcpyAsync(t->dYg, dY, (size_t)n * 8, cudaMemcpyHostToDevice, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    return PDM_OK;
}

// a5 on a row shard (_mk_connectivity_pits, dem_processing.py:1269-1382), between pdm_shard_links and pdm_shard_indeg.
// E_up / P_up: elevation and pit mask (pdm_tile field FLAT0 after pdm_shard_links) of the Hu rows above the owned rows,
// i.e. the last Hu owned rows of the rank above, in order; E_dn / P_dn: the first Hd owned rows of the rank below (device
// pointers; 0 rows and NULL at the ends of the grid).  A region may grow drain_pits_max_iter + 1 rows beyond a pit, so the
// strips must be that deep (or the search fails loudly).  Pit edges that end on a neighbour are counted into in_up / in_dn
// (int32 [Hin][C], the neighbour's rows next to the boundary, Hin >= drain_pits_max_dist): send them to the neighbours and
// hand what arrives to pdm_shard_pit_in_apply.  The sweep then pushes along such an edge straight into the neighbour's
// record, so this needs the multi-GPU work-list sweep (pdm_shard_p2p_connect_all).
int pdm_shard_pits(pdm_tile *t, const pdm_uca_params *p_in, const void *E_up, const void *P_up, int64_t Hu,
                   const void *E_dn, const void *P_dn, int64_t Hd, void *in_up, void *in_dn, int64_t Hin)
{
    if (!t || !p_in) { pdm_set_error("pdm_shard_pits: NULL argument"); return PDM_ERR_ARG; }
    if (!t->shard_pits_wanted) { pdm_set_error("pdm_shard_pits: call pdm_shard_links with drain_pits set first"); return PDM_ERR_STATE; }
    if (!pdm_shard_worklist_p2p(t)) {
        pdm_set_error("drain_pits=True on a row shard needs the multi-GPU work-list sweep (pdm_shard_p2p_connect_all): a pit may drain into the neighbouring rank's cells");
        return PDM_ERR_STATE;
    }
    if (!t->dXg) { pdm_set_error("pdm_shard_pits: pdm_tile_set_global_spacing first"); return PDM_ERR_STATE; }
    const Win &w = t->win;
    const bool up = w.lo > 0, dn = w.hi < t->R;
    if ((up && (!E_up || !P_up || Hu < 1 || (Hin > 0 && !in_up))) || (dn && (!E_dn || !P_dn || Hd < 1 || (Hin > 0 && !in_dn))) || Hin < 0) {
        pdm_set_error("pdm_shard_pits: strips of the neighbouring ranks are missing");
        return PDM_ERR_ARG;
    }
    pdm_pit_shard sh;
    memset(&sh, 0, sizeof(sh));
    if (up) { sh.E_up = (const double *)E_up; sh.P_up = (const uint8_t *)P_up; sh.Hu = Hu; sh.in_up = (int32_t *)in_up; sh.peer_row[0] = t->p2p.hi[0] - w.lo; }
    if (dn) { sh.E_dn = (const double *)E_dn; sh.P_dn = (const uint8_t *)P_dn; sh.Hd = Hd; sh.in_dn = (int32_t *)in_dn; sh.peer_row[1] = t->p2p.lo[1] - w.hi; }
    sh.Hin = Hin; sh.dXg = t->dXg; sh.dYg = t->dYg;
    const int64_t W = p_in->drain_pits_max_iter + 1;
    if ((w.hi - w.lo + 2 * W) * w.C >= ((int64_t)1 << 31) || (up && t->p2p.hi[0] * w.C >= ((int64_t)1 << 30)) ||
        (dn && (t->p2p.lo[1] + Hin) * w.C >= ((int64_t)1 << 30))) {
        pdm_set_error("pdm_shard_pits: shard too large for the 32-bit extended cell index of the pit search");
        return PDM_ERR_ARG;
    }
    PDM_CUDA(cudaMemsetAsync(t->label, 0, (size_t)t->N * sizeof(int32_t), t->stream));
    if (Hin > 0 && up) PDM_CUDA(cudaMemsetAsync(in_up, 0, (size_t)Hin * w.C * 4, t->stream));
    if (Hin > 0 && dn) PDM_CUDA(cudaMemsetAsync(in_dn, 0, (size_t)Hin * w.C * 4, t->stream));
    int rc = pdm_launch_pits(t, p_in, &sh);
    if (rc) return rc;
    t->shard_pits_done = true;
    return PDM_OK;
}

// from_up: what the rank above counted for my first Hin owned rows (its in_dn), from_dn: what the rank below counted
// for my last Hin owned rows (its in_up)
int pdm_shard_pit_in_apply(pdm_tile *t, const void *from_up, const void *from_dn, int64_t Hin)
{
    if (!t || !t->shard_pits_done) { pdm_set_error("pdm_shard_pit_in_apply: run pdm_shard_pits first"); return PDM_ERR_STATE; }
    const Win &w = t->win;
    if (Hin < 0 || Hin > w.hi - w.lo) { pdm_set_error("pdm_shard_pit_in_apply: Hin out of range"); return PDM_ERR_ARG; }
    const int64_t n = Hin * w.C;
    if (n == 0) return PDM_OK;
    if (from_up && w.lo > 0) {
        k_pit_in_apply<<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>(t->label + w.lo * w.C, (const int32_t *)from_up, n);
        PDM_LAUNCHED();
    }
    if (from_dn && w.hi < t->R) {
        k_pit_in_apply<<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>(t->label + (w.hi - Hin) * w.C, (const int32_t *)from_dn, n);
        PDM_LAUNCHED();
    }
    return PDM_OK;
}

// inflow-border mask + fresh sweep state (link AND proportion halo rows must be in place)
int pdm_shard_indeg(pdm_tile *t)
{
    if (!t) return PDM_ERR_ARG;
    if (t->shard_pits_wanted && !t->shard_pits_done) {
        pdm_set_error("pdm_shard_indeg: drain_pits was requested in pdm_shard_links; run pdm_shard_pits (+ pdm_shard_pit_in_apply) first");
        return PDM_ERR_STATE;
    }
    if (pdm_shard_worklist_p2p(t)) return pdm_launch_indeg_todo(t);    // Cell records of the work-list sweep
    int rc = pdm_launch_border_todo(t);
    if (rc) return rc;
    return pdm_ts_reset_state(t);
}

// one local accumulation pass until it comes to rest; first != 0: every tile, else the boundary
// tiles whose halo row received new donors
int pdm_shard_sweep(pdm_tile *t, int first)
{
    if (!t) return PDM_ERR_ARG;
    if (t->legacy_graph && pdm_shard_worklist_p2p(t)) {
        if (!first) { pdm_set_error("pdm_shard_sweep: a multi-GPU sweep has no resume rounds"); return PDM_ERR_STATE; }
        return pdm_launch_sweep_p2p(t);
    }
    return pdm_launch_tsweep(t, first);
}

// *sent (device int64) += boundary cells completed by the last pdm_shard_sweep whose receiver lives
// on a neighbouring rank (0 on every rank = the sweep is over)
int pdm_shard_sweep_sent(pdm_tile *t, void *sent)
{
    if (!t || !sent || !t->ts_ctr) { pdm_set_error("pdm_shard_sweep_sent: bad argument / no sweep yet"); return PDM_ERR_ARG; }
    k_add_sent<<<1, 1, 0, t->stream>>>(t->ts_ctr, (long long *)sent);
    PDM_LAUNCHED();
    return PDM_OK;
}

// ---- one sweep across GPUs: the row neighbours' boundary records, tile state words and queues as mapped
//      peer memory (CUDA IPC over NVLink), the termination counter on rank 0.
// *size: in: room at buf, out: bytes of the export blob (send it to the ranks above / below and, from rank 0,
// to every rank -- e.g. with an all-gather)
int pdm_shard_p2p_export(pdm_tile *t, void *buf, int64_t *size)
{
    if (!t || !size) { pdm_set_error("pdm_shard_p2p_export: NULL argument"); return PDM_ERR_ARG; }
    const int64_t need = (int64_t)sizeof(ts::P2PExport);
    if (!buf || *size < need) { *size = need; if (!buf) return PDM_OK; pdm_set_error("pdm_shard_p2p_export: need %lld bytes", (long long)need); return PDM_ERR_ARG; }
    *size = need;
    return pdm_ts_p2p_export(t, reinterpret_cast<ts::P2PExport *>(buf));
}

int pdm_shard_p2p_connect(pdm_tile *t, const void *up, const void *down, const void *root, int world, int rank)
{
    if (!t || world < 1 || rank < 0 || rank >= world) { pdm_set_error("pdm_shard_p2p_connect: bad argument"); return PDM_ERR_ARG; }
    if ((t->win.lo > 0) != (up != nullptr) || (t->win.hi < t->R) != (down != nullptr) || (rank == 0) != (root == nullptr)) {
        pdm_set_error("pdm_shard_p2p_connect: need the export of the rank above iff the tile has a halo row above, the one below "
                      "iff it has one below, and rank 0's export on every other rank");
        return PDM_ERR_ARG;
    }
    return pdm_ts_p2p_connect(t, reinterpret_cast<const ts::P2PExport *>(up), reinterpret_cast<const ts::P2PExport *>(down),
                              reinterpret_cast<const ts::P2PExport *>(root), world, rank);
}

// every rank's export, world blobs of pdm_shard_p2p_export's size back to back in rank order (e.g. the result of
// an all-gather): maps the row neighbours' records and every rank's control block.  With all ranks known the
// accumulation runs on the work-list engine as ONE sweep across the GPUs (PYDEM_B200_SHARD_SWEEP=tile keeps
// the tile sweep).
int pdm_shard_p2p_connect_all(pdm_tile *t, const void *blobs, int world, int rank)
{
    if (!t || !blobs || world < 1 || world > PDM_MAX_WORLD || rank < 0 || rank >= world) { pdm_set_error("pdm_shard_p2p_connect_all: bad argument"); return PDM_ERR_ARG; }
    if ((t->win.lo > 0) != (rank > 0) || (t->win.hi < t->R) != (rank + 1 < world)) {
        pdm_set_error("pdm_shard_p2p_connect_all: the tile's halo rows do not match rank %d of %d", rank, world);
        return PDM_ERR_ARG;
    }
    const ts::P2PExport *ex[PDM_MAX_WORLD];
    for (int r = 0; r < world; r++) ex[r] = reinterpret_cast<const ts::P2PExport *>(blobs) + r;
    return pdm_ts_p2p_connect_ranks(t, ex, world, rank);
}

// unmap the peers' memory (collective use: every rank disconnects, barrier, then tiles may be destroyed)
int pdm_shard_p2p_disconnect(pdm_tile *t)
{
    if (!t) { pdm_set_error("NULL tile"); return PDM_ERR_ARG; }
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    pdm_ts_p2p_close(t);
    return PDM_OK;
}

int pdm_shard_finalize(pdm_tile *t, const pdm_uca_params *p_in, pdm_uca_stats *stats)
{
    if (!t) return PDM_ERR_ARG;
    pdm_uca_params p;
    if (p_in) p = *p_in; else pdm_default_uca_params(&p);
    const bool wl = t->legacy_graph;     // the work-list engine ran (one sweep across the GPUs)
    int rc = wl ? pdm_launch_uca_finalize(t, &p) : pdm_launch_ts_finalize(t, &p);
    if (rc) return rc;
    rc = read_ctr(t);
    if (rc) return rc;
    rc = pdm_ts_read_counters(t);
    if (rc) return rc;
    const unsigned long long *wlc = t->ts_hctr + ts::TC_WLC;
    if (wl && wlc[CT_WATCHDOG]) {
        pdm_set_error("pdm_shard_finalize: multi-GPU work-list sweep gave up (code %llu: 1 = no progress, 2 = in-box overrun, 3 = lost in-box slot; "
                      "QTAIL=%llu QHEAD=%llu QDONE=%llu INBOX=%llu DRAINED=%llu)", wlc[CT_WATCHDOG], wlc[CT_QTAIL], wlc[CT_QHEAD],
                      wlc[CT_QDONE], wlc[CT_INBOX_TAIL], wlc[CT_DRAINED]);
        return PDM_ERR_STATE;
    }
    t->have_uca = true; t->have_graph = true;
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->n_cells = (t->win.hi - t->win.lo) * t->C;
        stats->n_sources = wl ? (int64_t)t->h_counters[CT_SOURCES] : (int64_t)t->ts_hctr[ts::TC_SOURCES];
        stats->n_drained = wl ? (int64_t)wlc[CT_DRAINED] : (int64_t)t->ts_hctr[ts::TC_CELLS];
        stats->n_queue_items = wl ? (int64_t)wlc[CT_QTAIL] : (int64_t)t->ts_hctr[ts::TC_VISITS];
        if (wl) stats->ms_sweep_kernel = (float)((double)(wlc[CT_T_END] - wlc[CT_T_START]) * 1e-6);
        stats->n_undone = (int64_t)t->h_counters[CT_UNDONE];
        stats->n_edge_todo = (int64_t)t->h_counters[CT_EDGE_TODO];
        stats->min_area = t->min_area;
        stats->n_pits = t->n_pits; stats->n_pit_edges = t->n_pit_edges;
        stats->n_pits_undrained = t->n_pits ? (int64_t)t->h_counters[CT_PITS_UNDRAINED] : 0;
    }
    return PDM_OK;
}

}  // extern "C"
