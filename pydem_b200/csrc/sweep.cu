// sweep.cu -- K5: upstream-contributing-area accumulation (the "drain" sweep) and TWI.
//
// Reference behaviour: cyutils.drain_area (cyutils.pyx:78-187) driven by
// _calc_uca_chunk (dem_processing.py:939-980).  The reference is level-synchronous and
// rescans every cell 3-4x per level; here the same dependency order (a cell drains when
// every cell draining into it has drained) is enforced by live in-degree counters:
//
//   * every cell with in-degree 0 is a source; a persistent grid scans for sources,
//   * a thread drains a cell by pushing area*weight (and the edge_todo "taint") to its
//     receivers with fp64 L2 atomics, then decrements the receivers' in-degree; the
//     thread that brings a counter to 0 owns that receiver and continues with it
//     (chain following, no level barrier).  A second ready receiver is handed to idle
//     lanes through a global queue.
//   * ordering: a receiver's area adds and its in-degree decrement hit the same 32-byte
//     sector in program order from one thread and are performed in that order at L2
//     (no MEMBAR on the critical path; drain_op.cuh explains, PYDEM_B200_SWEEP_STRICT=1
//     cross-checks with a true data dependency).
//
// For an acyclic graph this visits every cell exactly once and yields the reference's
// sums up to fp64 re-association (the reference adds in (level, source index) order).
// Circular references (dem_processing.py:951-964, "should never occur"): the reference restarts its
// sweep from the highest undone cells up to circular_ref_maxcount times.  That path is unreachable:
// an edge i -> j only survives the filter (1136-1137) if elev[j] <= elev[i] and its weight is > 1e-8,
// so a cycle would have to consist of cells of equal elevation -- but a facet neighbour at the centre's
// elevation always gets weight exactly 0 (s = 0 clamps the facet angle onto the other neighbour,
// 1942-1991), and pit edges only go to strictly lower cells or to non-pit cells that drain strictly
// downhill (1312-1320).  The graph is a DAG, one sweep finishes it, and the flag only matters through
// the loop condition: circular_ref_maxcount <= 1 means "no sweep at all" (honoured in
// pdm_launch_sweep_full, pinned against the reference in tests).  Should cells be left over after a
// sweep anyway, pdm_tile_uca fails loudly instead of returning partial sums.
//
// Traffic per cell (sweep): one 32-byte sector for the cell's own record and one per receiver
// (fp64 atomic on .area + int atomic on .indeg of the same sector); the kernel is bound by
// sector-granular DRAM access / L2 atomic latency, not by streaming bandwidth.
#include <stdlib.h>

#include "drain_op.cuh"
#include "tsweep.cuh"

namespace {

// multi-GPU work-list sweep, set-up: the seed count goes where the peers can read it, then this rank
// arrives at the start barrier on rank 0 (everything queued before this kernel -- records, queue and
// in-box resets -- is complete and, after the fence, visible to the peers)
__global__ void k_wl_p2p_arrive(unsigned long long *wlc, const unsigned long long *nseeds, unsigned long long *arrived)
{
    wlc[CT_SOURCES] = *nseeds;
    wlc[CT_INBOX_TAIL] = 0; wlc[CT_INBOX_HEAD] = 0; wlc[CT_WATCHDOG] = 0;
    __threadfence_system();
    atomicAdd_system(arrived, 1ULL);
}

// a6 epilogue: dem_processing.py:966-980 (owned cells [n0, n1)): sweep records -> uca, edge_done
__global__ void __launch_bounds__(256)
k_uca_finalize(const double *__restrict__ E, const uint8_t *__restrict__ flats, const Cell *__restrict__ cell,
               double *__restrict__ uca, uint8_t *__restrict__ edge_done,
               int64_t n0, int64_t n1, int limit_edges, double limit_area, unsigned long long *counters)
{
    const int64_t n = n0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool undone = false;
    if (n < n1) {
        const double e = E[n];
        const double2 at = *reinterpret_cast<const double2 *>(&cell[n].area);
        double u = at.x;
        if (flats[n]) u = __longlong_as_double(0x7ff8000000000000LL);                   // 972
        uca[n] = u;
        bool ed = !(at.y != 0.0);                                                        // 969, 974
        if (e != e) ed = true;                                                           // 975
        if (limit_edges && u > limit_area) ed = true;                                    // 977-980
        edge_done[n] = ed ? 1 : 0;
        undone = cell[n].indeg > 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, undone);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[CT_UNDONE], (unsigned long long)__popc(m));
}

// a9: calc_twi dem_processing.py:1647-1677 (un-scaled)
__global__ void __launch_bounds__(256)
k_twi(const double *__restrict__ uca, const double *__restrict__ mag, double *__restrict__ twi,
      double *__restrict__ twi10, int64_t N,
      double min_slope, double cap, int limit_uca, int limit_twi, double twi_sat)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    double u = uca[n];
    if (limit_uca && u > cap) u = cap;                                       // 1663-1665
    double t = log(__ddiv_rn(u, __dadd_rn(mag[n], min_slope)));              // 1667
    if (limit_twi && t > twi_sat) t = twi_sat;                               // 1669-1672
    twi[n] = t;
    if (twi10) twi10[n] = __dmul_rn(t, 10.0);                                // 1674: DEMProcessor.twi
}

}  // namespace

static int g_sweep_blocks = 0;

// Which engine runs the accumulation on a stand-alone tile (pdm_set_sweep_mode; default from the
// environment: PYDEM_B200_SWEEP=worklist|tile, default worklist):
//   0 work-list (worklist.cuh / drain_op.cuh): flow paths followed cell by cell through L2 atomics --
//     the fastest on one tile; sums in arrival order (fp64 atomics: equal up to re-association from
//     run to run), ordering of a receiver's adds against its count-off rests on same-sector program
//     order at L2 (drain_op.cuh; `strict` removes that assumption at the price of a round trip per cell);
//   1 tile sweep (tsweep.cu): pull-based, tile-resident, no floating-point atomics, ordered by the PTX
//     memory model only, bit-reproducible -- the engine of the row-sharded path and the cross-check
//     of the work-list in the GPU tests.
static int g_sweep_mode = -1, g_sweep_strict = -1;

bool pdm_sweep_legacy()
{
    if (g_sweep_mode < 0) {
        const char *e = getenv("PYDEM_B200_SWEEP"), *l = getenv("PYDEM_B200_SWEEP_LEGACY");
        int v = 0;
        if (e) v = (e[0] == 't' || e[0] == 'T' || e[0] == '1') ? 1 : 0;
        else if (l) v = atoi(l) ? 0 : 1;
        g_sweep_mode = v;
    }
    return g_sweep_mode == 0;
}

static int sweep_burst()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("PYDEM_B200_SWEEP_BURST"); v = (e && !atoi(e)) ? 0 : 1; }
    return v;
}

static int sweep_strict()
{
    if (g_sweep_strict < 0) { const char *e = getenv("PYDEM_B200_SWEEP_STRICT"); g_sweep_strict = (e && atoi(e)) ? 1 : 0; }
    return g_sweep_strict;
}

extern "C" int pdm_set_sweep_mode(int mode, int strict)
{
    if (mode < -1 || mode > 1) { pdm_set_error("pdm_set_sweep_mode: mode must be -1 (environment), 0 (work-list) or 1 (tile sweep)"); return PDM_ERR_ARG; }
    g_sweep_mode = mode;
    g_sweep_strict = strict < 0 ? -1 : (strict ? 1 : 0);
    return PDM_OK;
}

extern "C" int pdm_get_sweep_mode(void) { return pdm_sweep_legacy() ? 0 : 1; }

// first pass: every owned cell nobody drains into is a seed
int pdm_launch_sweep_first(pdm_tile *t)
{
    if (!g_sweep_blocks) {
        int rc = wl::grid_for(wl::k_worklist<DrainOp<0>, wl::DomainPatches>, &g_sweep_blocks);
        if (rc) return rc;
    }
    int rc = wl::reset_queue(t);
    if (rc) return rc;
    const Win &w = t->win;
    DrainOp<0> op{t->link, t->cell, nullptr, (int32_t)t->C, t->pit_beg, t->pit_end, t->pit_dst, t->pit_w, sweep_strict(), sweep_strict() ? 0 : sweep_burst(), (int32_t)t->R};
    wl::k_worklist<<<g_sweep_blocks, 256, 0, t->stream>>>(op, wl::DomainPatches{w.lo, w.hi - w.lo, w.C},
                                                          wl::tuned(wl::Queue{t->queue, t->d_counters, (long long)t->N, t->d_counters + CT_SOURCES}));
    PDM_LAUNCHED();
    return PDM_OK;
}

// Can the work-list sweep of this row shard run as one sweep across the GPUs?  (every rank's control block
// mapped: pdm_shard_p2p_connect_all; PYDEM_B200_SHARD_SWEEP=tile keeps the tile sweep)
bool pdm_shard_worklist_p2p(const pdm_tile *t)
{
    static int want = -1;
    if (want < 0) { const char *e = getenv("PYDEM_B200_SHARD_SWEEP"); want = (e && (e[0] == 't' || e[0] == 'T')) ? 0 : 1; }
    if (!want || !t->p2p.on || t->p2p.world < 2 || t->p2p.world > PDM_MAX_WORLD) return false;
    for (int r = 0; r < t->p2p.world; r++) if (!t->p2p.all_ctl[r]) return false;
    return true;
}

// ONE work-list sweep across the row shards of all GPUs (every rank launches it once per step): seeds =
// this rank's sources, pushes across a shard boundary go straight into the neighbour's records and in-box
// over NVLink, the end is detected on the counters of all ranks (worklist.cuh, p2p_service).
static int g_sweep_blocks_p2p = 0;
int pdm_launch_sweep_p2p(pdm_tile *t)
{
    if (!pdm_shard_worklist_p2p(t)) { pdm_set_error("pdm_launch_sweep_p2p: the tile is not connected to every rank"); return PDM_ERR_STATE; }
    if (!g_sweep_blocks_p2p) {
        int rc = wl::grid_for(wl::k_worklist<DrainOp<3>, wl::DomainPatches>, &g_sweep_blocks_p2p);
        if (rc) return rc;
    }
    pdm_tile::P2P &pp = t->p2p;
    const Win &w = t->win;
    unsigned long long *wlc = t->ts_ctr + ts::TC_WLC;
    // queue slots the previous run used -> -1 (its QTAIL lives in wlc), queue counters -> 0, in-box -> -1
    if (!t->queue_ready || t->queue_dirty_ctr != wlc) {
        PDM_CUDA(cudaMemsetAsync(t->queue, 0xFF, (size_t)(t->N + 1) * sizeof(int32_t), t->stream));
        t->queue_ready = true; t->queue_dirty_ctr = wlc;
    } else {
        wl::k_queue_clean<<<296, 256, 0, t->stream>>>(t->queue, wlc, (long long)t->N);
        PDM_LAUNCHED();
    }
    wl::k_queue_zero<<<1, 1, 0, t->stream>>>(wlc, 0);
    PDM_LAUNCHED();
    PDM_CUDA(cudaMemsetAsync(t->ts_slots, 0xFF, (size_t)t->ts_cap * sizeof(int32_t), t->stream));
    pp.launches++;
    unsigned long long *root = reinterpret_cast<unsigned long long *>(pp.all_ctl[0]);
    k_wl_p2p_arrive<<<1, 1, 0, t->stream>>>(wlc, t->d_counters + CT_SOURCES, root + ts::TC_ARRIVED);
    PDM_LAUNCHED();
    DrainOp<3> op{t->link, t->cell, nullptr, (int32_t)t->C, t->pit_beg, t->pit_end, t->pit_dst, t->pit_w, 0, sweep_burst(), (int32_t)t->R};
    op.own_lo = (int32_t)(w.lo * w.C); op.own_hi = (int32_t)(w.hi * w.C);
    for (int side = 0; side < 2; side++) {
        op.peer_cell[side] = nullptr; op.peer_ctr[side] = nullptr; op.peer_inbox[side] = nullptr; op.peer_cell0[side] = 0;
        if (!pp.rec[side]) continue;
        // above: the neighbour's last owned row; below: its first owned row
        const long long prow = side == 0 ? pp.hi[0] - 1 : pp.lo[1];
        op.peer_cell0[side] = (int32_t)(prow * w.C);
        op.peer_cell[side] = reinterpret_cast<Cell *>(const_cast<void *>(pp.rec[side])) + prow * w.C;
        char *ctl = reinterpret_cast<char *>(pp.ctl[side]);
        op.peer_ctr[side] = reinterpret_cast<unsigned long long *>(ctl) + ts::TC_WLC;
        op.peer_inbox[side] = reinterpret_cast<int32_t *>(ctl + pp.off_slots[side]);
    }
    wl::Queue q{t->queue, wlc, (long long)t->N, wlc + CT_SOURCES};
    q = wl::tuned(q);
    q.inbox = t->ts_slots; q.inbox_cap = (long long)t->ts_cap;
    q.epoch = pp.launches;
    q.arrived = root + ts::TC_ARRIVED; q.start_target = (unsigned long long)pp.world * pp.launches;
    q.world = pp.world;
    for (int r = 0; r < pp.world; r++) q.all_ctr[r] = reinterpret_cast<unsigned long long *>(pp.all_ctl[r]) + ts::TC_WLC;
    wl::k_worklist<<<g_sweep_blocks_p2p, 256, 0, t->stream>>>(op, wl::DomainPatches{w.lo, w.hi - w.lo, w.C}, q);
    PDM_LAUNCHED();
    return PDM_OK;
}

int pdm_launch_uca_finalize(pdm_tile *t, const pdm_uca_params *p)
{
    const Win &w = t->win;
    const double limit_area = p->uca_saturation_limit * 2 * t->min_area;
    const int64_t n0 = w.lo * w.C, n1 = w.hi * w.C;
    k_uca_finalize<<<(unsigned)((n1 - n0 + 255) / 256), 256, 0, t->stream>>>(
        t->elev, t->flats, t->cell, t->uca, t->edge_done, n0, n1, p->apply_uca_limit_edges, limit_area,
        t->d_counters);
    PDM_LAUNCHED();
    return PDM_OK;
}

int pdm_launch_sweep_full(pdm_tile *t, const pdm_uca_params *p, pdm_uca_stats *st)
{
    (void)st;
    if (t->legacy_graph) {
        // dem_processing.py:951-952: `while np.any(~done_) and count < self.circular_ref_maxcount ...` with count = 1:
        // for circular_ref_maxcount <= 1 the reference never calls drain_area -- every cell keeps its own area, only
        // the cells nobody drains into count as done.  One call (maxcount >= 2) drains a whole acyclic graph.
        if (p->circular_ref_maxcount > 1) {
            int rc = pdm_launch_sweep_first(t);
            if (rc) return rc;
        } else {
            PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_T_START, 0, 3 * sizeof(unsigned long long), t->stream));
        }
        return pdm_launch_uca_finalize(t, p);
    }
    int rc = pdm_ts_reset_state(t);
    if (rc) return rc;
    rc = pdm_launch_tsweep(t, 1);
    if (rc) return rc;
    return pdm_launch_ts_finalize(t, p);
}

int pdm_launch_twi(pdm_tile *t, const pdm_twi_params *p)
{
    const double cap = p->uca_saturation_limit * p->twi_min_area;
    const double twi_sat = log(p->uca_saturation_limit * p->twi_min_area / p->twi_min_slope);
    k_twi<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>(t->uca, t->mag, t->twi, t->twi10, t->N, p->twi_min_slope, cap,
                                                                 p->apply_twi_limits_on_uca, p->apply_twi_limits,
                                                                 twi_sat);
    PDM_LAUNCHED();
    return PDM_OK;
}
