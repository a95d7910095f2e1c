// tsweep.cu -- a7: the UCA accumulation sweep, tile-resident and pull-based.
//
// Reference behaviour: cyutils.drain_area (cyutils.pyx:78-187) driven by _calc_uca_chunk
// (dem_processing.py:939-964): a cell drains once every cell draining into it has drained; its
// area (and the edge_todo "taint", cyutils.pyx:163) is split between its receivers.
//
// The reference is level-synchronous over the whole grid (O(cells x depth)).  Round 1 followed the
// flow paths cell by cell through L2 atomics: ~1 us of dependent L2 traffic per level, so a
// conditioned 4096^2 DEM (3897 levels) took 5 ms, and the ordering of its un-fenced fp64 adds
// against the in-degree decrement rested on hardware behaviour.  This sweep keeps the dependent
// chain in shared memory and has no floating-point atomics at all:
//
//   * sweep state = one 32-byte record per cell (TRec: area, taint, proportion, link byte, donor
//     mask, flags), built once per graph (k_ts_init_records);
//   * the grid is cut into TW x TH tiles; a CTA that owns a tile stages the tile and a one-cell
//     ring of its neighbours in shared memory (one 32-byte load per cell);
//   * every undone cell counts its donors that are not done yet (popcount over its donor mask);
//     cells at zero form the first frontier; a ready cell PULLS: area = own cell area + the
//     contributions of its donors in ascending neighbour order (W, E, N, S, NW, NE, SW, SE) --
//     deterministic, bit-reproducible sums; it then decrements its in-tile receivers' counters
//     (shared-memory integer atomics) and receivers reaching zero form the next frontier.  Small
//     frontiers (rivers) are advanced by one warp with warp-level synchronisation only, and a
//     level is a few dozen instructions: the critical path of a river runs at shared-memory speed;
//   * completed cells are written back; a cell with a receiver in another tile marks that tile
//     for a (re)visit.  Tiles are scheduled asynchronously through a ticket queue in global
//     memory by a persistent grid; there is no grid-wide barrier.  A per-tile state word
//     {pending, running, complete} guarantees that a tile is run by one CTA at a time and is run
//     again whenever a neighbour published new donors after it was loaded.  A CTA that finishes
//     a tile continues with the neighbour it has just made runnable (no queue round trip on a
//     river's path);
//   * publication follows the PTX memory model: plain stores, CTA barrier, fence.acq_rel.gpu by
//     one thread, then the notifying atomic; the consumer's atomic on the tile state is followed
//     by a fence and a CTA barrier before anything is loaded (L2 loads, ld.cg).
//
// "done" is encoded in the values: area and taint hold a signalling-NaN pattern no arithmetic
// produces until the cell's sums are final.  They are two independent 8-byte words, each written
// exactly once (NOT_DONE -> final), and a cell counts as done only when both are final: a tile
// that is loaded while its neighbour is still writing can never pair a final area with a stale
// taint, without any ordering between the two stores.
//
// Long-range pit edges (_mk_connectivity_pits, dem_processing.py:1269-1382) are rare and keep a
// push form: a drained pit adds into per-receiver accumulators with global atomics, fences, and
// decrements the receiver's pit counter; the receiver is gated on that counter and adds the
// accumulator into its pull sum.
//
// Row shards: tiles cover the owned rows; halo rows are ring cells only.  A completed boundary
// cell whose receiver belongs to the neighbouring rank is recorded (TC_SENT); the host exchanges
// the boundary rows of records and resumes the tiles whose ring received new donors, until no rank
// completed such a cell (sharded.py).
#include <stdlib.h>

#include "tsweep.cuh"

namespace ts {

#define ST_NEW 0x01    // completed during this visit (store phase)

#define TF_PENDING 1u
#define TF_RUNNING 2u
#define TF_COMPLETE 4u
#define TF_VISITED 8u

#define RD_NONE 0xFFFFu    // receiver descriptor: no receiver in that slot
#define RD_OUT 0x8000u     // receiver outside the tile: low bits = bit of the notify mask (9 = other rank)

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int32_t ld_volatile_i32(const int32_t *p)
{
    int32_t v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// release / acquire fence at device scope (the default __threadfence() is the sequentially consistent
// flavour, MEMBAR.SC, and nothing here needs SC)
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ bool is_done(double a) { return (unsigned long long)__double_as_longlong(a) != TS_NOT_DONE; }

// neighbour q of a cell in ring coordinates: 0 W, 1 E, 2 N, 3 S, 4 NW, 5 NE, 6 SW, 7 SE
template <int HW>
__device__ __forceinline__ int nbr_off(int q)
{
    // rows {0,0,-1,+1,-1,-1,+1,+1} + 1 and cols {-1,+1,0,0,-1,+1,-1,+1} + 1, two bits each (q = 0 lowest)
    constexpr unsigned DR = 1u | (1u << 2) | (0u << 4) | (2u << 6) | (0u << 8) | (0u << 10) | (2u << 12) | (2u << 14);
    constexpr unsigned DC = 0u | (2u << 2) | (1u << 4) | (1u << 6) | (0u << 8) | (2u << 10) | (0u << 12) | (2u << 14);
    const int dr = (int)((DR >> (2 * q)) & 3u) - 1;
    const int dc = (int)((DC >> (2 * q)) & 3u) - 1;
    return dr * HW + dc;
}

// e1 / e2 offsets of a facet as (row, col) steps (facets table dem_processing.py:173-182)
__device__ __forceinline__ void off_e1(int sec, int &dr, int &dc)
{
    dr = ((0x6941 >> (2 * sec)) & 3) - 1;
    dc = ((0x9416 >> (2 * sec)) & 3) - 1;
}
__device__ __forceinline__ void off_e2(int sec, int &dr, int &dc)
{
    dr = ((sec >> 1) & 2) - 1;
    dc = 1 - (((sec + 2) >> 1) & 2);
}

// Shared memory of one tile = one warp.  The records stay in their 32-byte global layout: rows of the
// tile + ring arrive as bulk asynchronous copies (TMA, cp.async.bulk) signalled through an mbarrier.
template <int TW_, int TH_>
struct Smem {
    static constexpr int TW = TW_, TH = TH_, HW = TW_ + 2, HH = TH_ + 2, HN = HW * HH, TN = TW_ * TH_;
    static constexpr int CPR = TW_ / 32;          // cells per lane and row
    static constexpr int CPL = CPR * TH_;         // cells per lane: one bit each in the lane's masks
    TRec rec[HN];
    unsigned long long mbar;                      // mbarrier of the bulk copies
    double rowa[TH_];                             // cell area of the tile's rows (dX2 * dY2, dem_processing.py:885)
    uint32_t cand[32][2];                         // per lane: own cells to re-examine in the next pass (bit = owner_bit)
    uint32_t claim[TN / 32];                      // per own cell: a chain has taken it
    unsigned long long x_ph[8], x_t;              // debug: ns per phase
};

// Which lane owns cell (y, x) of the tile, and which bit of that lane's masks.  A lane owns CPR cells of
// every row; consecutive rows are rotated by 5 columns so that rows, columns and most diagonal runs of
// a river spread over different lanes.
#define TS_ROT 5
template <class S> __device__ __forceinline__ int owner_lane(int y, int x) { return ((x & 31) + TS_ROT * y) & 31; }
template <class S> __device__ __forceinline__ int owner_bit(int y, int x) { return y * S::CPR + (x >> 5); }
template <class S> __device__ __forceinline__ void cell_of(int lane, int bit, int &y, int &x)
{
    y = bit / S::CPR;
    x = ((lane - TS_ROT * y) & 31) + 32 * (bit - y * S::CPR);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void queue_push(const Args &a, int32_t tile)
{
    const unsigned long long t = atomicAdd(&a.ctr[TC_TAIL], 1ULL);
    const int32_t old = atomicExch(a.slots + (t & a.cap_mask), tile);
    if (old >= 0) atomicExch(&a.ctr[TC_ABORT], 2ULL);   // ring overrun: cannot happen (cap >= 4 x tiles), reported if it does
}

// "tile Y has a new donor": make it pending; whoever finds it idle owns the duty to schedule it
__device__ __forceinline__ bool notify_tile(const Args &a, int32_t y)
{
    const uint32_t old = atomicOr(&a.flag[y], TF_PENDING);
    if ((old & (TF_PENDING | TF_RUNNING | TF_COMPLETE)) == 0) {
        atomicAdd(&a.ctr[TC_INFLIGHT], 1ULL);
        return true;
    }
    return false;
}

// lane 0: wait for a queue item (ticket = fetch-and-add, never retries).  -1: the sweep is over.
__device__ int32_t acquire_tile(const Args &a)
{
    const unsigned long long h = atomicAdd(&a.ctr[TC_HEAD], 1ULL);
    int32_t *slot = a.slots + (h & a.cap_mask);
    unsigned ns = 32, polls = 0;
    unsigned long long t_last = 0, v_last = 0;
    for (;;) {
        // only the holder of this ticket consumes this slot in this lap: take it with one exchange
        const int32_t v = atomicExch(slot, -1);
        if (v >= 0) return v;
        if ((++polls & 7u) == 0) {
            // in-flight tiles only reach 0 when everything is done (a tile is counted from the moment
            // it is made pending until its run has ended and scheduled its successors)
            if (ld_volatile_u64(&a.ctr[TC_INFLIGHT]) == 0ULL) return -1;
            if (ld_volatile_u64(&a.ctr[TC_ABORT])) return -1;
            if ((polls & 1023u) == 0) {
                // watchdog (never fires in a correct run): no tile visit anywhere for 4 s
                const unsigned long long now = globaltimer_ns(), vis = ld_volatile_u64(&a.ctr[TC_VISITS]);
                if (t_last == 0 || vis != v_last) { t_last = now; v_last = vis; }
                else if (now - t_last > 4000000000ULL) { atomicExch(&a.ctr[TC_ABORT], 1ULL); return -1; }
            }
        }
        __nanosleep(ns);
        if (ns < 512) ns <<= 1;
    }
}

template <class S> __device__ __forceinline__ void mark_candidate(S &s, int y, int x)
{
    const int b = owner_bit<S>(y, x);
    atomicOr(&s.cand[owner_lane<S>(y, x)][b >> 5], 1u << (b & 31));
}

// both words of a staged record final?  (cells completed during this visit are written by other lanes
// with one 16-byte store; testing both words makes the test independent of how that store is performed)
template <class S> __device__ __forceinline__ bool cell_done(const S &s, int k)
{
    const double2 v = *reinterpret_cast<const double2 *>(&s.rec[k].area);
    return is_done(v.x) && is_done(v.y);
}

// all donors of ring cell k done (and, for a receiver of pit edges, its pit counter at zero)?
template <class S>
__device__ __forceinline__ bool cell_ready(const S &s, const Args &a, int k, int64_t n)
{
    const unsigned long long w = *reinterpret_cast<const unsigned long long *>(&s.rec[k].link);
    unsigned dm = (unsigned)(w >> 8) & 0xffu;
    while (dm) {
        const int q = __ffs(dm) - 1;
        dm &= dm - 1;
        if (!cell_done(s, k + nbr_off<S::HW>(q))) return false;
    }
    if ((w & LK_PITIN) && a.has_pits) {
        if (ld_volatile_i32(a.pit_cnt + n) != 0) return false;
        fence_acq_rel_gpu();       // counter seen at zero: the pit accumulators of this cell are final
    }
    return true;
}

// drained pit: push along its long-range edges (rare; plain fence ordering)
template <class S>
__device__ __forceinline__ void pit_push(S &s, const Args &a, double slot_bits, double ar, double tt, int rows_valid, int cols_valid,
                                         int64_t r0, int64_t c0, int my_tile)
{
    const int64_t slot = __double_as_longlong(slot_bits);
    const int32_t e0 = a.pit_beg[slot], e1 = a.pit_end[slot];
    for (int32_t e = e0; e < e1; e++) {
        const int32_t r = a.pit_dst[e];
        const double w = a.pit_w[e];
        atomicAdd(&a.pit_acc_a[r], __dmul_rn(ar, w));
        if (tt != 0.0) atomicAdd(&a.pit_acc_t[r], __dmul_rn(tt, w));
    }
    fence_acq_rel_gpu();
    for (int32_t e = e0; e < e1; e++) {
        const int32_t r = a.pit_dst[e];
        if (atomicSub(&a.pit_cnt[r], 1) != 1) continue;
        fence_acq_rel_gpu();   // the last decrement has observed all others: pass their accumulator adds on
        const int64_t ri = r / a.w.C, rj = r - ri * a.w.C;
        if (ri < a.w.lo || ri >= a.w.hi) continue;       // (pit edges never cross a shard: refused at graph build)
        const int y = (int)((ri - a.w.lo) / S::TH) * a.ntx + (int)(rj / S::TW);
        if (y == my_tile) {
            const int ry = (int)(ri - r0), rx = (int)(rj - c0);
            if (ry >= 0 && ry < rows_valid && rx >= 0 && rx < cols_valid) mark_candidate(s, ry, rx);
        } else if (notify_tile(a, y)) {
            queue_push(a, y);
        }
    }
}

#define TS_MARK(idx) if (a.dbg && lane == 0) { const unsigned long long now = globaltimer_ns(); s.x_ph[idx] += now - s.x_t; s.x_t = now; }

// One warp = one tile at a time.
//
// A visit: claim the tile, stage it (TMA), then PASSES until nothing moves: in a pass every lane looks at
// its own undone cells that got a new done donor ("candidates"; all undone cells in the first pass) and
// tests whether all donors are done (phase 1); after a warp barrier the ready cells pull their sums in
// ascending neighbour order, are written to shared and global memory, and nominate their in-tile
// receivers as candidates of the next pass (phase 2).  A pass is one dependency level inside the tile;
// its critical path is two shared-memory loads and the fp64 adds -- no lists, no counters, no atomics
// with a result.  Both phases read only what was written before the previous barrier: no races.
template <int TW, int TH>
__global__ void __launch_bounds__(32) k_tsweep(const Args a)
{
    typedef Smem<TW, TH> S;
    constexpr int HW = S::HW, CPL = S::CPL;
    static_assert(CPL <= 64, "a lane's cells must fit one 64-bit mask");
    extern __shared__ __align__(128) unsigned char ts_raw[];
    S &s = *reinterpret_cast<S *>(ts_raw);
    const int lane = threadIdx.x;
    const int64_t C = a.w.C;
    const unsigned FULL = 0xffffffffu;
    const uint32_t mbar = smem_u32(&s.mbar);
    unsigned long long x_cells = 0, x_levels = 0, x_sources = 0, x_rerun = 0, x_defer = 0, x_sent = 0, x_pass_cyc = 0;
    s.cand[lane][0] = 0; s.cand[lane][1] = 0;
    if (lane == 0) {
        for (int q = 0; q < 8; q++) s.x_ph[q] = 0;
        s.x_t = globaltimer_ns();
        atomicMin(&a.ctr[TC_T_START], s.x_t);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t mphase = 0;
    int tile = 0;
    if (lane == 0) tile = acquire_tile(a);
    tile = __shfl_sync(FULL, tile, 0);
    // mode 0: tile came from the queue (pending -> claim it); 1: handed over by the previous tile of
    // this warp (already claimed); 2: same tile again in place (ring reload only)
    int mode = 0, first = 0;
    while (tile >= 0) {
        if (lane == 0) {
            if (mode == 0) {
                // claim: pending -> running.  Everything published before the notification that made
                // the tile pending is visible after this atomic + fence + warp barrier.
                const uint32_t old = atomicExch(&a.flag[tile], TF_RUNNING | TF_VISITED);
                first = (old & TF_VISITED) ? 0 : 1;
            }
            fence_acq_rel_gpu();
        }
        first = __shfl_sync(FULL, first, 0);
        TS_MARK(0)
        const int ty = tile / a.ntx, tx = tile - ty * a.ntx;
        const int64_t r0 = a.w.lo + (int64_t)ty * TH, c0 = (int64_t)tx * TW;
        const int rows_valid = (int)min((int64_t)TH, a.w.hi - r0), cols_valid = (int)min((int64_t)TW, C - c0);
        const int64_t j0 = max(c0 - 1, (int64_t)0), j1 = min(c0 + cols_valid + 1, C);      // columns of the ring that exist
        const int64_t i0 = max(r0 - 1, (int64_t)0), i1 = min(r0 + rows_valid + 1, a.w.R);  // rows of the ring that exist
        // ---- stage the tile
        if (mode != 2) {
            // shared memory was read / written through the generic proxy during the previous visit
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const uint32_t bytes = (uint32_t)((i1 - i0) * (j1 - j0) * (int64_t)sizeof(TRec));
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
            }
            __syncwarp();
            for (int64_t gi = i0 + lane; gi < i1; gi += 32) {
                const int hy = (int)(gi - (r0 - 1));
                const uint32_t dst = smem_u32(&s.rec[hy * HW + (int)(j0 - (c0 - 1))]);
                const uint32_t bytes = (uint32_t)((j1 - j0) * (int64_t)sizeof(TRec));
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(a.rec + gi * C + j0), "r"(bytes), "r"(mbar) : "memory");
            }
            for (int y = lane; y < rows_valid; y += 32) s.rowa[y] = __ldg(a.row_area + r0 + y);
            // ring cells outside the grid: done, no receivers (nothing else ever reads beyond the ring)
            for (int idx = lane; idx < 2 * (cols_valid + 2) + 2 * rows_valid; idx += 32) {
                int hy, hx;
                if (idx < cols_valid + 2) { hy = 0; hx = idx; }
                else if (idx < 2 * (cols_valid + 2)) { hy = rows_valid + 1; hx = idx - (cols_valid + 2); }
                else { const int u = idx - 2 * (cols_valid + 2); hy = 1 + (u >> 1); hx = (u & 1) ? cols_valid + 1 : 0; }
                const int64_t gi = r0 - 1 + hy, gj = c0 - 1 + hx;
                if (gi < 0 || gi >= a.w.R || gj < 0 || gj >= C) {
                    TRec z; z.area = 0.0; z.taint = 0.0; z.prop = 0.0; z.link = LK_NOSEC; z.dmask = 0; z.flags = 0;
                    s.rec[hy * HW + hx] = z;
                }
            }
            // wait for the bytes
            {
                uint32_t done = 0;
                while (!done) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(mbar), "r"(mphase) : "memory");
                }
                mphase ^= 1u;
            }
            __syncwarp();
        }
        // ring cells: (re)load in place for a repeated visit; a cell counts as done only when both words are final
        for (int idx = lane; idx < 2 * (cols_valid + 2) + 2 * rows_valid; idx += 32) {
            int hy, hx;
            if (idx < cols_valid + 2) { hy = 0; hx = idx; }
            else if (idx < 2 * (cols_valid + 2)) { hy = rows_valid + 1; hx = idx - (cols_valid + 2); }
            else { const int u = idx - 2 * (cols_valid + 2); hy = 1 + (u >> 1); hx = (u & 1) ? cols_valid + 1 : 0; }
            const int64_t gi = r0 - 1 + hy, gj = c0 - 1 + hx;
            if (gi < 0 || gi >= a.w.R || gj < 0 || gj >= C) continue;
            TRec &rr = s.rec[hy * HW + hx];
            double2 v;
            if (mode == 2) v = __ldcg(reinterpret_cast<const double2 *>(&a.rec[gi * C + gj].area));
            else v = *reinterpret_cast<const double2 *>(&rr.area);
            if (!is_done(v.x) || !is_done(v.y)) { v.x = __longlong_as_double((long long)TS_NOT_DONE); v.y = 0.0; }
            *reinterpret_cast<double2 *>(&rr.area) = v;
        }
        TS_MARK(1)
        // ---- this lane's undone cells: the candidates of the first pass
        uint32_t cand0 = 0, cand1 = 0;
        int nsrc = 0, undone_at_load = 0;
        for (int i = lane; i < S::TN / 32; i += 32) s.claim[i] = 0;
#pragma unroll 8
        for (int b = 0; b < CPL; b++) {
            int y, x;
            cell_of<S>(lane, b, y, x);
            if (y < rows_valid && x < cols_valid) {
                const TRec &rr = s.rec[(y + 1) * HW + x + 1];
                if (!is_done(rr.area)) {
                    if (b < 32) cand0 |= 1u << b; else cand1 |= 1u << (b - 32);
                    undone_at_load++;
                    if (first && rr.dmask == 0 && !(rr.link & LK_PITIN)) nsrc++;
                }
            }
        }
        __syncwarp();
        TS_MARK(2)
        // ---- passes.  Phase 1: every lane tests its candidate cells (own cells that may have become
        //      ready).  Phase 2: a ready cell pulls its sum and the lane FOLLOWS THE FLOW PATH: it tests
        //      the cell's receivers (any lane's cells) and continues with one that is ready and that it
        //      can claim -- a river crosses the tile on one lane, one cell after the other, with no
        //      warp-wide step in between.  A receiver that is not ready yet is left as a candidate for
        //      its owner: whoever completes its last donor finds it ready, and should two lanes
        //      complete the last two donors at the same moment and both see the other's cell
        //      unfinished, the owner's next pass picks it up.
        unsigned notify = 0;
        int lvl = 0, completed = 0;
        const long long c_in = a.dbg ? clock64() : 0;
        for (;;) {
            uint32_t ready0 = 0, ready1 = 0;
#pragma unroll
            for (int h = 0; h < (CPL > 32 ? 2 : 1); h++) {
                for (uint32_t mm = h ? cand1 : cand0; mm; mm &= mm - 1) {
                    const int b = __ffs(mm) - 1 + 32 * h;
                    int y, x;
                    cell_of<S>(lane, b, y, x);
                    const int k = (y + 1) * HW + x + 1;
                    if (is_done(s.rec[k].area)) continue;
                    if (cell_ready(s, a, k, (r0 + y) * C + (c0 + x))) { if (h) ready1 |= 1u << (b - 32); else ready0 |= 1u << b; }
                }
            }
            if (!__any_sync(FULL, (ready0 | ready1) != 0)) break;
            __syncwarp();
#pragma unroll
            for (int h = 0; h < (CPL > 32 ? 2 : 1); h++) {
                for (uint32_t mm = h ? ready1 : ready0; mm; mm &= mm - 1) {
                    const int b = __ffs(mm) - 1 + 32 * h;
                    int y, x;
                    cell_of<S>(lane, b, y, x);
                    int spare = -1;          // a second receiver that became ready at a fork
                    for (;;) {
                        // ---- pull (cyutils.pyx:161-163 seen from the receiver), ascending neighbour order
                        const int k = (y + 1) * HW + x + 1;
                        const unsigned long long w = *reinterpret_cast<const unsigned long long *>(&s.rec[k].link);
                        const uint8_t lk = (uint8_t)w;
                        double ar = s.rowa[y];                                                   // dem_processing.py:885, 901
                        double tt = ((w >> 16) & TR_TODO) ? 1.0 : 0.0;                           // 944
                        unsigned dm = (unsigned)(w >> 8) & 0xffu;
                        while (dm) {
                            const int q = __ffs(dm) - 1;
                            dm &= dm - 1;
                            const TRec &d = s.rec[k + nbr_off<HW>(q)];
                            const double p = d.prop;
                            const double wgt = q < 4 ? p : __dsub_rn(1.0, p);                    // dem_processing.py:1082
                            ar = __dadd_rn(ar, __dmul_rn(d.area, wgt));                          // cyutils.pyx:161
                            tt = __dadd_rn(tt, __dmul_rn(d.taint, wgt));                         // cyutils.pyx:163
                        }
                        const int64_t n = (r0 + y) * C + (c0 + x);
                        if ((lk & LK_PITIN) && a.has_pits) {
                            ar = __dadd_rn(ar, __ldcg(a.pit_acc_a + n));
                            tt = __dadd_rn(tt, __ldcg(a.pit_acc_t + n));
                        }
                        *reinterpret_cast<double2 *>(&s.rec[k].area) = make_double2(ar, tt);
                        *reinterpret_cast<double2 *>(&a.rec[n].area) = make_double2(ar, tt);
                        completed++;
                        // ---- the receivers
                        int ny = -1, nx = -1;
                        if (lk & LK_PIT) {
                            pit_push(s, a, s.rec[k].prop, ar, tt, rows_valid, cols_valid, r0, c0, tile);
                        } else if (!(lk & LK_NOSEC)) {
                            const int sec = lk & LK_SEC_MASK;
#pragma unroll
                            for (int e = 0; e < 2; e++) {
                                if (!(lk & (e == 0 ? LK_KEEP1 : LK_KEEP2))) continue;
                                int dr, dc;
                                if (e == 0) off_e1(sec, dr, dc); else off_e2(sec, dr, dc);
                                const int ry = y + dr, rx = x + dc;
                                if (ry >= 0 && ry < rows_valid && rx >= 0 && rx < cols_valid) {
                                    const int rk = (ry + 1) * HW + rx + 1;
                                    bool mine = false;
                                    if (cell_ready(s, a, rk, (r0 + ry) * C + (c0 + rx))) {
                                        const int ro = ry * TW + rx;
                                        mine = !(atomicOr(&s.claim[ro >> 5], 1u << (ro & 31)) & (1u << (ro & 31)));
                                    } else {
                                        mark_candidate(s, ry, rx);
                                    }
                                    if (mine) {
                                        if (ny < 0) { ny = ry; nx = rx; }
                                        else if (spare < 0) spare = ry * TW + rx;
                                        else mark_candidate(s, ry, rx);      // claimed but parked: its owner takes it in the next pass
                                    }
                                } else {
                                    const int64_t gr = r0 + ry;
                                    if (gr < a.w.lo || gr >= a.w.hi) notify |= 1u << 9;           // receiver on the neighbouring rank
                                    else notify |= 1u << ((ry < 0 ? 0 : (ry >= rows_valid ? 2 : 1)) * 3 + (rx < 0 ? 0 : (rx >= cols_valid ? 2 : 1)));
                                }
                            }
                        }
                        if (ny < 0 && spare >= 0) { ny = spare / TW; nx = spare - ny * TW; spare = -1; }
                        if (ny < 0) break;
                        y = ny; x = nx;
                    }
                }
            }
            __syncwarp();
            cand0 = s.cand[lane][0];
            if (cand0) s.cand[lane][0] = 0;
            if (CPL > 32) { cand1 = s.cand[lane][1]; if (cand1) s.cand[lane][1] = 0; }
            lvl++;
        }
        if (a.dbg) x_pass_cyc += (unsigned long long)(clock64() - c_in);
        TS_MARK(3)
        // ---- totals of the visit
        int completed_all = completed, undone_all = undone_at_load, nsrc_all = nsrc;
        for (int d = 16; d > 0; d >>= 1) {
            completed_all += __shfl_xor_sync(FULL, completed_all, d);
            undone_all += __shfl_xor_sync(FULL, undone_all, d);
            nsrc_all += __shfl_xor_sync(FULL, nsrc_all, d);
            notify |= __shfl_xor_sync(FULL, notify, d);
        }
        // ---- schedule.  The global stores of all lanes are ordered before the fence of lane 0 by the warp
        //      barrier; the notifying atomics follow the fence.  Lane b < 9 looks after neighbour tile b of
        //      the 3x3 block, lane 9 after this tile's own state word: one round trip for all of them.
        __syncwarp();
        if (lane == 0) fence_acq_rel_gpu();
        __syncwarp();
        {
            const unsigned nm = notify;
            const int y2 = ty + lane / 3 - 1, x2 = tx + lane % 3 - 1;
            const bool mine = lane < 9 && lane != 4 && ((nm >> lane) & 1u) && y2 >= 0 && y2 < a.nty && x2 >= 0 && x2 < a.ntx;
            const int32_t nb = y2 * a.ntx + x2;
            uint32_t old = TF_PENDING;
            if (mine) old = atomicOr(&a.flag[nb], TF_PENDING);
            const bool complete = completed_all == undone_all;
            uint32_t own_old = 0;
            if (lane == 9) own_old = atomicCAS(&a.flag[tile], TF_RUNNING | TF_VISITED, TF_VISITED | (complete ? TF_COMPLETE : 0u));
            own_old = __shfl_sync(FULL, own_old, 9);
            const bool released = own_old == (TF_RUNNING | TF_VISITED);
            const bool idle = mine && (old & (TF_PENDING | TF_RUNNING | TF_COMPLETE)) == 0;   // I made it pending: I schedule it
            const unsigned im = __ballot_sync(FULL, idle);
            const int n_idle = __popc(im);
            // a released warp continues with its first idle successor itself (no queue round trip on the
            // critical path of a river); everything else goes to the queue
            const int take_lane = (released && im) ? __ffs(im) - 1 : -1;
            // in-flight accounting: +1 per tile made pending, -1 for this tile if released.  The handed-over
            // tile inherits this tile's count, so a river moves on without touching the counter.  Increments
            // are performed before the tiles become visible in the queue, the decrement comes last.
            const int n_push = n_idle - (take_lane >= 0 ? 1 : 0);
            int requeue_self = 0;
            if (!released) {
                // a neighbour published donors while this tile was loaded.  While the queue holds other
                // work the tile goes to the back (several neighbours' updates are then served by one visit);
                // once the queue is empty -- the river phase -- it is run again at once, ring reload only.
                if (lane == 0) {
                    const long long backlog = (long long)ld_volatile_u64(&a.ctr[TC_TAIL]) - (long long)ld_volatile_u64(&a.ctr[TC_HEAD]);
                    requeue_self = backlog > 0;
                }
                requeue_self = __shfl_sync(FULL, requeue_self, 0);
            }
            if (lane == 0 && n_push > 0) (void)atomicAdd(&a.ctr[TC_INFLIGHT], (unsigned long long)n_push);
            if (n_push > 0) { if (lane == 0) fence_acq_rel_gpu(); __syncwarp(); }
            if (idle && lane != take_lane) queue_push(a, nb);
            int next_tile = -1, next_mode = 0, next_first = 0;
            if (released) {
                if (take_lane >= 0) {
                    next_tile = __shfl_sync(FULL, nb, take_lane);
                    next_first = (__shfl_sync(FULL, old, take_lane) & TF_VISITED) ? 0 : 1;
                    next_mode = 1;
                    if (lane == take_lane) atomicExch(&a.flag[nb], TF_RUNNING | TF_VISITED);     // claim (lane 0 fences at the top of the visit)
                } else if (lane == 0) {
                    atomicAdd(&a.ctr[TC_INFLIGHT], ~0ULL);   // -1
                }
            } else if (requeue_self) {
                if (lane == 0) {
                    atomicExch(&a.flag[tile], TF_PENDING | TF_VISITED);     // still counted in flight
                    queue_push(a, tile);
                    x_defer++;
                }
            } else {
                if (lane == 0) { atomicExch(&a.flag[tile], TF_RUNNING | TF_VISITED); x_rerun++; }
                next_tile = tile; next_mode = 2;
            }
            if (lane == 0) {
                atomicAdd(&a.ctr[TC_VISITS], 1ULL);      // heartbeat of the watchdog
                if (a.dbg) {
                    unsigned long long b = (globaltimer_ns() - ld_volatile_u64(&a.ctr[TC_T_START])) / 100000ULL;
                    if (b > 63) b = 63;
                    atomicAdd(&a.ctr[TC_HIST + 2 * b], 1ULL);
                    atomicAdd(&a.ctr[TC_HIST + 2 * b + 1], (unsigned long long)completed_all);
                }
                x_cells += (unsigned long long)completed_all; x_levels += (unsigned long long)lvl;
                x_sources += (unsigned long long)nsrc_all; x_sent += (unsigned long long)((nm >> 9) & 1u);
                if (next_tile < 0) { next_tile = acquire_tile(a); next_mode = 0; }
            }
            tile = __shfl_sync(FULL, next_tile, 0);
            mode = __shfl_sync(FULL, next_mode, 0);
            first = next_mode == 1 ? next_first : 0;
        }
        TS_MARK(5)
    }
    if (lane == 0) {
        if (x_cells) atomicAdd(&a.ctr[TC_CELLS], x_cells);
        if (x_levels) atomicAdd(&a.ctr[TC_LEVELS], x_levels);
        if (x_sources) atomicAdd(&a.ctr[TC_SOURCES], x_sources);
        if (x_rerun) atomicAdd(&a.ctr[TC_REQUEUE], x_rerun);
        if (x_defer) atomicAdd(&a.ctr[TC_DEFER], x_defer);
        if (x_sent) atomicAdd(&a.ctr[TC_SENT], x_sent);
        atomicMax(&a.ctr[TC_T_END], globaltimer_ns());
        if (a.dbg) {
            for (int q = 0; q < 6; q++) atomicAdd(&a.ctr[TC_PHASE + q], s.x_ph[q]);
            atomicAdd(&a.ctr[TC_PHASE + 6], x_pass_cyc);
            atomicAdd(&a.ctr[TC_PHASE + 7], x_levels);
        }
    }
}

// queue / tile-state set-up of one launch.  mode 0: every tile pending (first pass).
// mode 1 (shard resume): the tiles of the first / last tile row whose ring row received a donor
// since the last launch (seen[] remembers, per column of the two halo rows, what was already done).
__global__ void __launch_bounds__(256)
k_ts_fill(Args a, int mode, uint8_t *seen)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= (int64_t)a.cap_mask) a.slots[i] = (mode == 0 && i < a.ntiles) ? (int32_t)i : -1;
    if (mode == 0) {
        if (i < a.ntiles) a.flag[i] = TF_PENDING;
        if (seen && i < 2 * a.w.C) seen[i] = 0;
    }
}

__device__ __forceinline__ bool rec_done(const TRec *r) { return is_done(r->area) && is_done(r->taint); }

__global__ void __launch_bounds__(256)
k_ts_setup(Args a, int mode, int TW, uint8_t *seen)
{
    const int tid = threadIdx.x;
    __shared__ int s_count;
    if (tid == 0) s_count = 0;
    __syncthreads();
    const int64_t C = a.w.C;
    if (mode == 0) {
        if (tid == 0) s_count = a.ntiles;
    } else {
        for (int side = 0; side < 2; side++) {
            const int64_t hrow = side == 0 ? a.w.lo - 1 : a.w.hi;
            if (hrow < 0 || hrow >= a.w.R) continue;
            const int ty = side == 0 ? 0 : a.nty - 1;
            for (int tx = tid; tx < a.ntx; tx += blockDim.x) {
                bool fresh = false;
                const int64_t j0 = max((int64_t)0, (int64_t)tx * TW - 1), j1 = min(C, (int64_t)(tx + 1) * TW + 1);
                for (int64_t j = j0; j < j1; j++)
                    if (rec_done(a.rec + hrow * C + j) && !seen[side * C + j]) { fresh = true; break; }
                if (!fresh) continue;
                const int32_t y = ty * a.ntx + tx;
                const uint32_t old = atomicOr(&a.flag[y], TF_PENDING);
                if ((old & (TF_PENDING | TF_COMPLETE)) == 0) a.slots[atomicAdd(&s_count, 1)] = y;
            }
        }
        __syncthreads();
        for (int side = 0; side < 2; side++) {
            const int64_t hrow = side == 0 ? a.w.lo - 1 : a.w.hi;
            if (hrow < 0 || hrow >= a.w.R) continue;
            for (int64_t j = tid; j < C; j += blockDim.x) seen[side * C + j] = rec_done(a.rec + hrow * C + j) ? 1 : 0;
        }
    }
    __syncthreads();
    if (tid == 0) {
        a.ctr[TC_HEAD] = 0; a.ctr[TC_TAIL] = (unsigned long long)s_count; a.ctr[TC_INFLIGHT] = (unsigned long long)s_count;
        a.ctr[TC_ABORT] = 0; a.ctr[TC_SENT] = 0;
        a.ctr[TC_T_START] = ~0ULL; a.ctr[TC_T_END] = 0;
        if (mode == 0) { a.ctr[TC_VISITS] = 0; a.ctr[TC_CELLS] = 0; a.ctr[TC_SOURCES] = 0; a.ctr[TC_LEVELS] = 0; a.ctr[TC_REQUEUE] = 0; a.ctr[TC_DEFER] = 0; }
        a.ctr[TC_QUEUED] = (unsigned long long)s_count;
        for (int b = 0; b < 128; b++) a.ctr[TC_HIST + b] = 0;
        for (int q = 0; q < 8; q++) a.ctr[TC_PHASE + q] = 0;
    }
}

// does neighbour byte lk drain into the centre?  (keepbit: the receiver slot the centre occupies --
// cardinal e1 or diagonal e2; secmask: the facets whose slot points at it)
__device__ __forceinline__ int drains_in(uint8_t lk, uint8_t keepbit, uint32_t secmask)
{
    return ((lk & keepbit) && !(lk & (LK_NOSEC | LK_PIT)) && ((secmask >> (lk & LK_SEC_MASK)) & 1u)) ? 1 : 0;
}

// sweep records of the owned rows of a fresh graph (the CSR row structure of the reference's matrix
// B = A.tocsr(), dem_processing.py:879, reduced to one byte per cell): donor mask from the eight
// neighbours' link bytes, proportion, link, inflow-border flag; area / taint "not done".  Halo rows
// of a shard get their records from the neighbouring rank.
__global__ void __launch_bounds__(256)
k_ts_init_records(const uint8_t *__restrict__ link, const double *__restrict__ prop, const uint8_t *__restrict__ edge_todo,
                  Win w, TRec *__restrict__ rec)
{
    const int64_t C = w.C;
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = w.lo + (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= w.hi || j >= C) return;
    const int64_t n = i * C + j;
    const bool up = w.row_in_grid(i - 1), dn = w.row_in_grid(i + 1), lf = j > 0, rt = j < C - 1;
    // all eight neighbour bytes are loaded unconditionally (a missing neighbour reads the cell's own
    // byte and is masked out), so the loads are independent and in flight together
    const int64_t dW = lf ? -1 : 0, dE = rt ? 1 : 0, dN = up ? -C : 0, dS = dn ? C : 0;
    const uint8_t bW = link[n + dW], bE = link[n + dE], bN = link[n + dN], bS = link[n + dS];
    const uint8_t bNW = link[n + dN + dW], bNE = link[n + dN + dE], bSW = link[n + dS + dW], bSE = link[n + dS + dE];
    unsigned m = 0;
    m |= (lf ? drains_in(bW, LK_KEEP1, 0x81u) : 0) << 0;                    // W neighbour: its e1 = (0,+1) for facets 0, 7
    m |= (rt ? drains_in(bE, LK_KEEP1, 0x18u) : 0) << 1;                    // E: e1 = (0,-1) for facets 3, 4
    m |= (up ? drains_in(bN, LK_KEEP1, 0x60u) : 0) << 2;                    // N: e1 = (+1,0) for facets 5, 6
    m |= (dn ? drains_in(bS, LK_KEEP1, 0x06u) : 0) << 3;                    // S: e1 = (-1,0) for facets 1, 2
    m |= ((up && lf) ? drains_in(bNW, LK_KEEP2, 0xC0u) : 0) << 4;           // NW: e2 = (+1,+1) for facets 6, 7
    m |= ((up && rt) ? drains_in(bNE, LK_KEEP2, 0x30u) : 0) << 5;           // NE: e2 = (+1,-1) for facets 4, 5
    m |= ((dn && lf) ? drains_in(bSW, LK_KEEP2, 0x03u) : 0) << 6;           // SW: e2 = (-1,+1) for facets 0, 1
    m |= ((dn && rt) ? drains_in(bSE, LK_KEEP2, 0x0Cu) : 0) << 7;           // SE: e2 = (-1,-1) for facets 2, 3
    const unsigned long long word = (unsigned long long)link[n] | ((unsigned long long)m << 8) |
                                    ((unsigned long long)(edge_todo[n] ? TR_TODO : 0) << 16);
    double2 *out = reinterpret_cast<double2 *>(rec + n);
    out[0] = make_double2(__longlong_as_double((long long)TS_NOT_DONE), __longlong_as_double((long long)TS_NOT_DONE));
    out[1] = make_double2(prop[n], __longlong_as_double((long long)word));
}

// a6 epilogue: dem_processing.py:966-980 on the owned cells [n0, n1)
__global__ void __launch_bounds__(256)
k_ts_finalize(const double *__restrict__ E, const uint8_t *__restrict__ flats, const TRec *__restrict__ rec,
              double *__restrict__ uca, uint8_t *__restrict__ edge_done, int64_t n0, int64_t n1, int limit_edges,
              double limit_area, unsigned long long *counters)
{
    const int64_t n = n0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool undone = false;
    if (n < n1) {
        const double2 at = *reinterpret_cast<const double2 *>(&rec[n].area);
        double u = at.x;
        undone = !is_done(at.x) || !is_done(at.y);
        if (!undone) {
            if (flats[n]) u = __longlong_as_double(0x7ff8000000000000LL);                // 972
            uca[n] = u;
            bool ed = !(at.y != 0.0);                                                    // 969, 974
            const double e = E[n];
            if (e != e) ed = true;                                                       // 975
            if (limit_edges && u > limit_area) ed = true;                                // 977-980
            edge_done[n] = ed ? 1 : 0;
        } else {
            uca[n] = __longlong_as_double((long long)TS_NOT_DONE);      // cells on circular references: left to the restart pass
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, undone);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counters[CT_UNDONE], (unsigned long long)__popc(m));
}

// mark the receivers of pit edges (their pull adds the pit accumulator and waits for the pit counter)
__global__ void __launch_bounds__(256)
k_pit_mark(const int32_t *__restrict__ pit_dst, int64_t n_edges, uint8_t *link)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int32_t r = pit_dst[e];
    atomicOr(reinterpret_cast<unsigned int *>(link) + (r >> 2), (unsigned)LK_PITIN << ((r & 3) * 8));
}

struct Variant { int tw, th, nt; size_t smem; void (*kernel)(const Args); int blocks; };

#define TS_VARIANT(tw, th) {tw, th, 32, sizeof(Smem<tw, th>), k_tsweep<tw, th>, 0}
static Variant g_variants[] = {
    TS_VARIANT(32, 32), TS_VARIANT(32, 16), TS_VARIANT(64, 16), TS_VARIANT(64, 32), TS_VARIANT(32, 64), TS_VARIANT(32, 8),
};

static int pick_variant()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PYDEM_B200_TS_TILE");
        v = e ? atoi(e) : 0;
        if (v < 0 || v >= (int)(sizeof(g_variants) / sizeof(g_variants[0]))) v = 0;
    }
    return v;
}

}  // namespace ts

using namespace ts;

static int ts_prepare(pdm_tile *t, Variant **out)
{
    Variant &v = g_variants[pick_variant()];
    if (!v.blocks) {
        int dev = 0, sms = 0, occ = 0;
        PDM_CUDA(cudaGetDevice(&dev));
        PDM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        PDM_CUDA(cudaFuncSetAttribute(v.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem));
        PDM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, v.kernel, v.nt, v.smem));
        if (occ < 1) { pdm_set_error("tile sweep kernel does not fit on an SM"); return PDM_ERR_CUDA; }
        const char *e = getenv("PYDEM_B200_TS_OCC");
        if (e && atoi(e) > 0 && atoi(e) < occ) occ = atoi(e);
        v.blocks = sms * occ;
    }
    const Win &w = t->win;
    const int64_t ntx = (w.C + v.tw - 1) / v.tw, nty = (w.hi - w.lo + v.th - 1) / v.th;
    const int64_t ntiles = ntx * nty;
    int64_t cap = 1024;
    while (cap < 4 * ntiles) cap <<= 1;
    if (t->ts_cap < cap || t->ts_ntiles_cap < ntiles) {
        if (t->ts_slots) { cudaFree(t->ts_slots); t->ts_slots = nullptr; }
        if (t->ts_flag) { cudaFree(t->ts_flag); t->ts_flag = nullptr; }
        PDM_CUDA(cudaMalloc(&t->ts_slots, (size_t)cap * 4));
        PDM_CUDA(cudaMalloc(&t->ts_flag, (size_t)ntiles * 4));
        t->ts_cap = cap; t->ts_ntiles_cap = ntiles;
    }
    if (!t->ts_ctr) {
        PDM_CUDA(cudaMalloc(&t->ts_ctr, TC_N * sizeof(unsigned long long)));
        PDM_CUDA(cudaMallocHost((void **)&t->ts_hctr, TC_N * sizeof(unsigned long long)));
        PDM_CUDA(cudaMemsetAsync(t->ts_ctr, 0, TC_N * sizeof(unsigned long long), t->stream));
    }
    if (!t->ts_seen) PDM_CUDA(cudaMalloc(&t->ts_seen, (size_t)2 * w.C));
    *out = &v;
    return PDM_OK;
}

static Args ts_args(pdm_tile *t, const Variant &v)
{
    const Win &w = t->win;
    Args a;
    a.rec = pdm_trec(t);
    a.row_area = t->row_area;
    a.pit_beg = t->pit_beg; a.pit_end = t->pit_end; a.pit_dst = t->pit_dst; a.pit_w = t->pit_w;
    // pit accumulators: UCA (written by the epilogue only) and the proportion scratch (consumed by
    // k_ts_init_records) are free while the sweep runs
    a.pit_cnt = t->label; a.pit_acc_a = t->uca; a.pit_acc_t = t->twi;
    a.w = w;
    a.ntx = (int32_t)((w.C + v.tw - 1) / v.tw); a.nty = (int32_t)((w.hi - w.lo + v.th - 1) / v.th);
    a.ntiles = a.ntx * a.nty;
    a.slots = t->ts_slots; a.cap_mask = (uint32_t)(t->ts_cap - 1); a.flag = t->ts_flag; a.ctr = t->ts_ctr;
    static int dbg = -1;
    if (dbg < 0) { const char *e = getenv("PYDEM_B200_TS_DEBUG"); dbg = e ? atoi(e) : 0; }
    a.dbg = dbg;
    a.has_pits = t->n_pit_edges > 0 ? 1 : 0;
    return a;
}

// the sweep state of a fresh graph: records of the owned rows (link halo rows in place on a shard),
// pit receivers marked, pit accumulators at zero
int pdm_ts_reset_state(pdm_tile *t)
{
    const Win &w = t->win;
    if (t->n_pit_edges > 0) {
        k_pit_mark<<<(unsigned)((t->n_pit_edges + 255) / 256), 256, 0, t->stream>>>(t->pit_dst, t->n_pit_edges, t->link);
        PDM_LAUNCHED();
    }
    dim3 block(32, 8);
    dim3 grid((unsigned)((w.C + 31) / 32), (unsigned)((w.hi - w.lo + 7) / 8));
    k_ts_init_records<<<grid, block, 0, t->stream>>>(t->link, t->twi, t->edge_todo, w, pdm_trec(t));
    PDM_LAUNCHED();
    if (t->n_pit_edges > 0) {
        PDM_CUDA(cudaMemsetAsync(t->uca, 0, (size_t)t->N * 8, t->stream));
        PDM_CUDA(cudaMemsetAsync(t->twi, 0, (size_t)t->N * 8, t->stream));
    }
    return PDM_OK;
}

// first != 0: every tile is pending; else (shard resume) the boundary tiles that received donors
int pdm_launch_tsweep(pdm_tile *t, int first)
{
    Variant *v = nullptr;
    int rc = ts_prepare(t, &v);
    if (rc) return rc;
    const Args a = ts_args(t, *v);
    {
        int64_t nfill = (int64_t)a.cap_mask + 1;
        if (first && 2 * a.w.C > nfill) nfill = 2 * a.w.C;
        k_ts_fill<<<(unsigned)((nfill + 255) / 256), 256, 0, t->stream>>>(a, first ? 0 : 1, t->ts_seen);
        PDM_LAUNCHED();
    }
    k_ts_setup<<<1, 256, 0, t->stream>>>(a, first ? 0 : 1, v->tw, t->ts_seen);
    PDM_LAUNCHED();
    int blocks = v->blocks;
    if (blocks > a.ntiles) blocks = a.ntiles;
    if (blocks < 1) blocks = 1;
    v->kernel<<<blocks, v->nt, v->smem, t->stream>>>(a);
    PDM_LAUNCHED();
    return PDM_OK;
}

int pdm_ts_read_counters(pdm_tile *t)
{
    PDM_CUDA(cudaMemcpyAsync(t->ts_hctr, t->ts_ctr, TC_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    if (t->ts_hctr[TC_ABORT]) {
        pdm_set_error("tile sweep aborted (%s): visits=%llu cells=%llu inflight=%llu head=%llu tail=%llu",
                      t->ts_hctr[TC_ABORT] == 2 ? "queue overrun" : "watchdog: no progress for 4 s", t->ts_hctr[TC_VISITS],
                      t->ts_hctr[TC_CELLS], t->ts_hctr[TC_INFLIGHT], t->ts_hctr[TC_HEAD], t->ts_hctr[TC_TAIL]);
        return PDM_ERR_STATE;
    }
    return PDM_OK;
}

int pdm_launch_ts_finalize(pdm_tile *t, const pdm_uca_params *p)
{
    const Win &w = t->win;
    const double limit_area = p->uca_saturation_limit * 2 * t->min_area;
    const int64_t n0 = w.lo * w.C, n1 = w.hi * w.C;
    k_ts_finalize<<<(unsigned)((n1 - n0 + 255) / 256), 256, 0, t->stream>>>(
        t->elev, t->flats, pdm_trec(t), t->uca, t->edge_done, n0, n1, p->apply_uca_limit_edges, limit_area, t->d_counters);
    PDM_LAUNCHED();
    return PDM_OK;
}
