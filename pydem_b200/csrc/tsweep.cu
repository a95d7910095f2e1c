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
#include <string.h>
#include <unistd.h>

#include <cub/device/device_radix_sort.cuh>

#include "tsweep.cuh"

namespace ts {

#define ST_NEW 0x01    // completed during this visit (store phase)

#define TF_PENDING 1u
#define TF_RUNNING 2u
#define TF_COMPLETE 4u
#define TF_VISITED 8u

#define TS_MARK_DONE 0x40000000u   // counter word of a cell completed in this visit (written to global memory when the visit ends)
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

// Shared memory of one CTA = one tile.  The records stay in their 32-byte global layout: rows of the
// tile + ring arrive as bulk asynchronous copies (TMA, cp.async.bulk) signalled through an mbarrier.
template <int TW_, int TH_>
struct Smem {
    static constexpr int TW = TW_, TH = TH_, HW = TW_ + 2, HH = TH_ + 2, HN = HW * HH, TN = TW_ * TH_;
    TRec rec[HN];
    unsigned long long mbar;                      // mbarrier of the bulk copies
    double rowa[TH_];                             // cell area of the tile's rows (dX2 * dY2, dem_processing.py:885)
    uint32_t cnt[HN];                             // per staged cell: own cells: donors that are not final yet (TS_MARK_DONE once completed in this visit); ring cells: bit 31
    uint32_t rcv[HN];                             // own cells that are not final: ring indices of the two receivers (low / high half, 0xffff = none)
    uint16_t fr[TN];                              // the cells in the order they became ready (a topological order of this visit)
    int n_fr;                                     // entries of fr = cells that became ready (tickets of the producers)
    int head;                                     // tickets of the consumers
    int live;                                     // ready cells not finished yet (queued or held by a lane); 0 = the visit is over
    int completed;                                // cells completed in this visit
    int undone;                                   // own cells not final when the tile was staged
    int nsrc;                                     // of those: cells nobody drains into (first visit)
    unsigned notify;                              // neighbour tiles that got a new donor (bit = 3 * dy + dx; 9: the neighbouring rank)
    int late;                                     // a pit of this tile released a receiver inside this tile: run again
    int next_tile, next_mode, next_first;
    unsigned long long x_ph[8], x_t;              // debug: ns per phase
    unsigned long long x_late[8];                 // the same for visits that start after dbg x 100 us (dbg >= 10)
    int x_is_late;
    unsigned long long x_dbg2[4];
    int x_busy_it;                                // debug: iterations in which the busiest warp held a cell
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Tile state words, queue slots and the in-flight counter are also written by the neighbouring GPUs when one
// sweep spans several row shards (a.p2p): every access to them is then a system-scope atomic.
__device__ __forceinline__ unsigned long long add_u64(const Args &a, unsigned long long *p, unsigned long long v)
{
    return a.p2p ? atomicAdd_system(p, v) : atomicAdd(p, v);
}
__device__ __forceinline__ uint32_t or_u32(const Args &a, uint32_t *p, uint32_t v) { return a.p2p ? atomicOr_system(p, v) : atomicOr(p, v); }
__device__ __forceinline__ uint32_t exch_u32(const Args &a, uint32_t *p, uint32_t v) { return a.p2p ? atomicExch_system(p, v) : atomicExch(p, v); }
__device__ __forceinline__ int32_t exch_i32(const Args &a, int32_t *p, int32_t v) { return a.p2p ? atomicExch_system(p, v) : atomicExch(p, v); }
__device__ __forceinline__ uint32_t cas_u32(const Args &a, uint32_t *p, uint32_t c, uint32_t v) { return a.p2p ? atomicCAS_system(p, c, v) : atomicCAS(p, c, v); }
// release / acquire fence of a visit: device scope, system scope when peers read this GPU's records
__device__ __forceinline__ void fence_visit(const Args &a)
{
    if (a.p2p) asm volatile("fence.acq_rel.sys;" ::: "memory");
    else fence_acq_rel_gpu();
}

// push into a queue given by its counters / slots (this rank's or a peer's)
__device__ __forceinline__ void queue_push_to(const Args &a, unsigned long long *ctr, int32_t *slots, uint32_t cap_mask, int32_t tile)
{
    const unsigned long long t = add_u64(a, &ctr[TC_TAIL], 1ULL);
    const int32_t old = exch_i32(a, slots + (t & cap_mask), tile);
    if (old >= 0) atomicExch(&a.ctr[TC_ABORT], 2ULL);   // ring overrun: cannot happen (cap >= 4 x tiles), reported if it does
}
__device__ __forceinline__ void queue_push(const Args &a, int32_t tile) { queue_push_to(a, a.ctr, a.slots, a.cap_mask, tile); }

// "tile Y has a new donor": make it pending; whoever finds it idle owns the duty to schedule it
__device__ __forceinline__ bool notify_tile(const Args &a, int32_t y)
{
    const uint32_t old = or_u32(a, &a.flag[y], TF_PENDING);
    if ((old & (TF_PENDING | TF_RUNNING | TF_COMPLETE)) == 0) {
        add_u64(a, a.inflight, 1ULL);
        return true;
    }
    return false;
}

// one thread: wait for a queue item (ticket = fetch-and-add, never retries).  -1: the sweep is over.
__device__ int32_t acquire_tile(const Args &a)
{
    const unsigned long long h = atomicAdd(&a.ctr[TC_HEAD], 1ULL);
    int32_t *slot = a.slots + (h & a.cap_mask);
    unsigned ns = 32, polls = 0;
    unsigned long long t_last = 0, v_last = 0;
    const unsigned every = a.p2p ? 31u : 7u;      // the counter of a multi-GPU sweep lives on rank 0: poll it over NVLink less often
    for (;;) {
        // only the holder of this ticket consumes this slot in this lap: take it with one exchange
        const int32_t v = exch_i32(a, slot, -1);
        if (v >= 0) return v;
        if ((++polls & every) == 0) {
            // in-flight tiles only reach 0 when everything is done (a tile is counted from the moment
            // it is made pending until its run has ended and scheduled its successors)
            if (ld_volatile_u64(a.inflight) == 0ULL) return -1;
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

// drained pit: push along its long-range edges (rare; plain fence ordering).  A receiver released inside
// this tile makes the tile run again (s.late); one in another tile makes that tile pending.
template <class S>
__device__ __noinline__ void pit_push(S &s, const Args &a, double slot_bits, double ar, double tt, int my_tile)
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
            s.late = 1;
        } else if (notify_tile(a, y)) {
            queue_push(a, y);
        }
    }
}

#define TS_MARK(idx) if (a.dbg && tid == 0) { const unsigned long long now = globaltimer_ns(); s.x_ph[idx] += now - s.x_t; if (s.x_is_late) s.x_late[idx] += now - s.x_t; s.x_t = now; }

#ifdef TS_NO_CTA_FENCE
__device__ __forceinline__ void fence_cta() { asm volatile("" ::: "memory"); }
#else
__device__ __forceinline__ void fence_cta() { asm volatile("fence.acq_rel.cta;" ::: "memory"); }
#endif
__device__ __forceinline__ int ld_volatile_shared_i32(const int *p)
{
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_volatile_shared_f64x2(const double *p)
{
    double2 v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_shared_f64x2(double *p, double x, double y)
{
    asm volatile("st.volatile.shared.v2.f64 [%0], {%1, %2};" ::"r"(smem_u32(p)), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ unsigned ld_volatile_shared_u16(const uint16_t *p)
{
    unsigned short v;
    asm volatile("ld.volatile.shared.u16 %0, [%1];" : "=h"(v) : "r"(smem_u32(p)) : "memory");
    return (unsigned)v;
}

// One ready cell: PULL its sums (cyutils.pyx:161-163 seen from the receiver) over its donors, publish
// them in shared and global memory, and count the cell off at its receivers.  The sum has a fixed shape
// -- own area + (((W + E) + (N + S)) + ((NW + NE) + (SW + SE))) with exact zeros for neighbours that are
// not donors -- so a cell's value does not depend on the schedule, the tile shape or the sharding.
// Written for the dependent path of a river (one lane, nothing to overlap with): all eight neighbours are
// staged, so their loads are unconditional and in flight together with the cell's own word (~30 cycles);
// the sums are a depth-4 tree (~8 cycles a level); the two count-offs are in flight together; measured
// ~150 cycles a cell in isolation (scripts/ubench/smem_lat.cu).  r1 / r2: in-tile receivers that became
// ready (ring index) or -1.
template <class S>
__device__ __forceinline__ void drain_cell(S &s, const Args &a, int k, int rows_valid, int cols_valid, int64_t r0, int64_t c0,
                                           int tile, unsigned &notify, int &r1, int &r2)
{
    constexpr int HW = S::HW;
    const TRec *rk = &s.rec[k];
    const unsigned long long w = *reinterpret_cast<const unsigned long long *>(&rk->link);
    const uint32_t rv = s.rcv[k];
    const int hy = k / HW;                          // ring row: own cells 1..TH
    unsigned dm = (unsigned)(w >> 8) & 0xffu;
    const unsigned lk = (unsigned)w & 0xffu;
#ifndef TS_DRAIN_LOOP
    double2 v[8];
    double p[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const TRec *d = rk + nbr_off<HW>(q);
        v[q] = *reinterpret_cast<const double2 *>(&d->area);
        p[q] = d->prop;
    }
    const double base = s.rowa[hy - 1];                                          // dem_processing.py:885, 901
    double ca[8], ct[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const double wgt = q < 4 ? p[q] : __dsub_rn(1.0, p[q]);                  // dem_processing.py:1082
        const bool on = (dm >> q) & 1u;
        ca[q] = on ? __dmul_rn(v[q].x, wgt) : 0.0;                               // cyutils.pyx:161
        ct[q] = on ? __dmul_rn(v[q].y, wgt) : 0.0;                               // cyutils.pyx:163
    }
    double ar = __dadd_rn(__dadd_rn(__dadd_rn(ca[0], ca[1]), __dadd_rn(ca[2], ca[3])), __dadd_rn(__dadd_rn(ca[4], ca[5]), __dadd_rn(ca[6], ca[7])));
    double tt = __dadd_rn(__dadd_rn(__dadd_rn(ct[0], ct[1]), __dadd_rn(ct[2], ct[3])), __dadd_rn(__dadd_rn(ct[4], ct[5]), __dadd_rn(ct[6], ct[7])));
    ar = __dadd_rn(base, ar);
    tt = __dadd_rn(((w >> 16) & TR_TODO) ? 1.0 : 0.0, tt);                       // 944
#else
    // donors in ascending neighbour order (W, E, N, S, NW, NE, SW, SE), a few instructions each: a cell has
    // two or three donors, and a single lane on a river pays for every instruction of the step
    double ar = s.rowa[hy - 1];                                                  // dem_processing.py:885, 901
    double tt = ((w >> 16) & TR_TODO) ? 1.0 : 0.0;                               // 944
    while (dm) {
        const int q = __ffs(dm) - 1;
        dm &= dm - 1;
        const TRec *d = rk + nbr_off<HW>(q);
        const double2 v = *reinterpret_cast<const double2 *>(&d->area);
        const double p = d->prop;
        const double wgt = q < 4 ? p : __dsub_rn(1.0, p);                        // dem_processing.py:1082
        ar = __dadd_rn(ar, __dmul_rn(v.x, wgt));                                 // cyutils.pyx:161
        tt = __dadd_rn(tt, __dmul_rn(v.y, wgt));                                 // cyutils.pyx:163
    }
#endif
    if ((lk & LK_PITIN) && a.has_pits) {
        const int64_t n = (r0 + hy - 1) * a.w.C + (c0 + (k - hy * HW) - 1);
        ar = __dadd_rn(ar, __ldcg(a.pit_acc_a + n));
        tt = __dadd_rn(tt, __ldcg(a.pit_acc_t + n));
    }
    *reinterpret_cast<double2 *>(&s.rec[k].area) = make_double2(ar, tt);
    s.cnt[k] = TS_MARK_DONE;                       // (nobody counts a completed cell off any more)
    fence_cta();                                   // the sums are visible in the CTA before the cell is counted off
    const int k1 = (int)(rv & 0xffffu), k2 = (int)(rv >> 16);
    uint32_t o1 = 0, o2 = 0;
    if (k1 != 0xffff) o1 = atomicSub(&s.cnt[k1], 1u);
    if (k2 != 0xffff) o2 = atomicSub(&s.cnt[k2], 1u);
    r1 = o1 == 1u ? k1 : -1;
    r2 = o2 == 1u ? k2 : -1;
    if (r1 >= 0 || r2 >= 0) fence_cta();           // the last count-off has observed the others: their sums are visible
    if ((o1 | o2) & 0x80000000u) {
        // a receiver in the ring: another tile (or the neighbouring rank) has a new donor
#pragma unroll
        for (int e = 0; e < 2; e++) {
            if (!((e ? o2 : o1) & 0x80000000u)) continue;
            const int kk = e ? k2 : k1;
            const int ry = kk / HW, rx = kk - ry * HW;
            const int64_t gr = r0 + ry - 1;
            if (gr < a.w.lo || gr >= a.w.hi)     // receiver on the neighbouring rank: bit 9 + the peer's tile (10..12 above, 13..15 below)
                notify |= (1u << 9) | (1u << ((gr < a.w.lo ? 10 : 13) + (rx < 1 ? 0 : (rx > cols_valid ? 2 : 1))));
            else notify |= 1u << ((ry < 1 ? 0 : (ry > rows_valid ? 2 : 1)) * 3 + (rx < 1 ? 0 : (rx > cols_valid ? 2 : 1)));
        }
    }
    if ((lk & LK_PIT) && a.has_pits) pit_push(s, a, s.rec[k].prop, ar, tt, tile);
}

// One CTA = one tile at a time.
//
// A visit: claim the tile, stage it and its ring (TMA), count for every own cell that is not final the
// donors that are not final (ring cells included), and queue the cells at zero.  Then every lane of the
// CTA follows flow paths: it takes a ready cell (ticket queue in shared memory), pulls its sums, writes
// them to shared and global memory, counts the cell off at its in-tile receivers and continues with a
// receiver that reached zero (a second one goes to the queue).  No level barrier: hill slopes are worked
// by all lanes at once, a river crossing the tile advances on one lane at shared-memory latency.
template <int TW, int TH, int NT>
__global__ void __launch_bounds__(NT, 1) k_tsweep(const __grid_constant__ Args a)
{
    typedef Smem<TW, TH> S;
    constexpr int HW = S::HW, TN = S::TN;
    static_assert(TW % 4 == 0, "own cells are counted four to a word");
    extern __shared__ __align__(128) unsigned char ts_raw[];
    S &s = *reinterpret_cast<S *>(ts_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t C = a.w.C;
    const unsigned FULL = 0xffffffffu;
    const uint32_t mbar = smem_u32(&s.mbar);
    unsigned long long x_cells = 0, x_levels = 0, x_sources = 0, x_rerun = 0, x_defer = 0, x_sent = 0;
    if (tid == 0) {
        for (int q = 0; q < 8; q++) { s.x_ph[q] = 0; s.x_late[q] = 0; }
        for (int q = 0; q < 4; q++) s.x_dbg2[q] = 0;
        s.x_is_late = 0;
        s.x_t = globaltimer_ns();
        atomicMin(&a.ctr[TC_T_START], s.x_t);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (a.p2p) {
            // one sweep over several GPUs: nobody starts before every rank has queued its tiles and
            // counted them in (else the global in-flight counter could touch zero too early)
            const unsigned long long t0 = globaltimer_ns();
            unsigned polls = 0;
            while (ld_volatile_u64(a.arrived) < a.start_target) {
                __nanosleep(200);
                if ((++polls & 255u) == 0 && globaltimer_ns() - t0 > 4000000000ULL) { atomicExch(&a.ctr[TC_ABORT], 3ULL); break; }
            }
            asm volatile("fence.acq_rel.sys;" ::: "memory");
        }
        s.next_tile = ld_volatile_u64(&a.ctr[TC_ABORT]) ? -1 : acquire_tile(a); s.next_mode = 0; s.next_first = 0;
    }
    __syncthreads();
    uint32_t mphase = 0;
    int tile = s.next_tile;
    // mode 0: tile came from the queue (pending -> claim it); 1: handed over by the previous tile of
    // this CTA (already claimed); 2: same tile again in place (ring reload only)
    int mode = 0, first = 0;
    while (tile >= 0) {
        if (tid == 0) {
            if (mode == 0) {
                // claim: pending -> running.  Everything published before the notification that made
                // the tile pending is visible after this atomic + fence + CTA barrier.
                const uint32_t old = exch_u32(a, &a.flag[tile], TF_RUNNING | TF_VISITED);
                s.next_first = (old & TF_VISITED) ? 0 : 1;
            }
            fence_visit(a);
            s.n_fr = 0; s.undone = 0; s.nsrc = 0; s.notify = 0; s.late = 0;
            if (a.dbg >= 10) s.x_is_late = (globaltimer_ns() - ld_volatile_u64(&a.ctr[TC_T_START])) > (unsigned long long)(a.dbg % 1000) * 100000ULL;
        }
        TS_MARK(0)
        const int ty = tile / a.ntx, tx = tile - ty * a.ntx;
        const int64_t r0 = a.w.lo + (int64_t)ty * TH, c0 = (int64_t)tx * TW;
        const int rows_valid = (int)min((int64_t)TH, a.w.hi - r0), cols_valid = (int)min((int64_t)TW, C - c0);
        const int64_t j0 = max(c0 - 1, (int64_t)0), j1 = min(c0 + cols_valid + 1, C);      // columns of the ring that exist
        const int64_t i0 = max(r0 - 1, (int64_t)0), i1 = min(r0 + rows_valid + 1, a.w.R);  // rows of the ring that exist
        const int nring = 2 * (cols_valid + 2) + 2 * rows_valid;
        // ---- stage the tile
        if (mode != 2) {
            // shared memory was read / written through the generic proxy during the previous visit
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) {
                const uint32_t bytes = (uint32_t)((i1 - i0) * (j1 - j0) * (int64_t)sizeof(TRec));
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
            }
            for (int64_t gi = i0 + tid; gi < i1; gi += NT) {
                const int hy = (int)(gi - (r0 - 1));
                const uint32_t dst = smem_u32(&s.rec[hy * HW + (int)(j0 - (c0 - 1))]);
                const uint32_t bytes = (uint32_t)((j1 - j0) * (int64_t)sizeof(TRec));
                // a halo row of a multi-GPU sweep is read where it lives: the neighbour's boundary row, over NVLink
                const TRec *src = a.rec + gi * C + j0;
                if (a.p2p) {
                    if (gi < a.w.lo && (a.p2p & 1)) src = a.halo_src[0] + j0;
                    else if (gi >= a.w.hi && (a.p2p & 2)) src = a.halo_src[1] + j0;
                }
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
            }
            for (int y = tid; y < rows_valid; y += NT) s.rowa[y] = __ldg(a.row_area + r0 + y);
            // ring cells outside the grid: done, no receivers (nothing else ever reads beyond the ring)
            for (int idx = tid; idx < nring; idx += NT) {
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
        } else {
            __syncthreads();           // thread 0 has reset the visit's counters
        }
        first = s.next_first;
        for (int i = tid; i < TN / 2; i += NT) reinterpret_cast<uint32_t *>(s.fr)[i] = 0xffffffffu;      // empty queue slots
        for (int i = tid; i < S::HN; i += NT) s.cnt[i] = 0x80000000u;                                      // ring cells never become ready (bit 31 survives their few count-offs)
        // ring cells: (re)load in place for a repeated visit; a cell counts as done only when both words are final
        for (int idx = tid; idx < nring; idx += NT) {
            int hy, hx;
            if (idx < cols_valid + 2) { hy = 0; hx = idx; }
            else if (idx < 2 * (cols_valid + 2)) { hy = rows_valid + 1; hx = idx - (cols_valid + 2); }
            else { const int u = idx - 2 * (cols_valid + 2); hy = 1 + (u >> 1); hx = (u & 1) ? cols_valid + 1 : 0; }
            const int64_t gi = r0 - 1 + hy, gj = c0 - 1 + hx;
            if (gi < 0 || gi >= a.w.R || gj < 0 || gj >= C) continue;
            TRec &rr = s.rec[hy * HW + hx];
            double2 v;
            if (mode == 2) {
                const TRec *src = a.rec + gi * C + gj;
                if (a.p2p && gi < a.w.lo && (a.p2p & 1)) src = a.halo_src[0] + gj;
                else if (a.p2p && gi >= a.w.hi && (a.p2p & 2)) src = a.halo_src[1] + gj;
                v.x = __longlong_as_double((long long)ld_volatile_u64(reinterpret_cast<const unsigned long long *>(&src->area)));
                v.y = __longlong_as_double((long long)ld_volatile_u64(reinterpret_cast<const unsigned long long *>(&src->taint)));
            }
            else v = *reinterpret_cast<const double2 *>(&rr.area);
            if (!is_done(v.x) || !is_done(v.y)) { v.x = __longlong_as_double((long long)TS_NOT_DONE); v.y = 0.0; }
            *reinterpret_cast<double2 *>(&rr.area) = v;
        }
        __syncthreads();
        TS_MARK(1)
        // ---- count: per own cell that is not final, its donors that are not final; cells at zero are ready
        {
            int undone_here = 0, nsrc_here = 0;
            for (int g = tid; g < TN / 4; g += NT) {
                const int y = g / (TW / 4), x4 = (g - y * (TW / 4)) * 4;
                if (y < rows_valid) {
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int x = x4 + u;
                        if (x >= cols_valid) continue;
                        const int k = (y + 1) * HW + x + 1;
                        unsigned c = 0;
                        if (!is_done(s.rec[k].area)) {
                            undone_here++;
                            const unsigned long long w = *reinterpret_cast<const unsigned long long *>(&s.rec[k].link);
                            unsigned dm = (unsigned)(w >> 8) & 0xffu;
                            if (first && dm == 0 && !(w & LK_PITIN)) nsrc_here++;
                            while (dm) {
                                const int q = __ffs(dm) - 1;
                                dm &= dm - 1;
                                c += is_done(s.rec[k + nbr_off<HW>(q)].area) ? 0u : 1u;
                            }
                            if ((w & LK_PITIN) && a.has_pits) {
                                if (ld_volatile_i32(a.pit_cnt + (r0 + y) * C + (c0 + x)) != 0) c += 1;       // pit edges still to arrive: not in this visit
                                else fence_acq_rel_gpu();       // counter seen at zero: the pit accumulators of this cell are final
                            }
                            if (c == 0) s.fr[atomicAdd(&s.n_fr, 1)] = (uint16_t)k;      // (published by the CTA barrier below)
                            // the two receivers as ring indices (facets table dem_processing.py:173-182; a receiver in the
                            // ring belongs to another tile: its counter word only tells that)
                            const unsigned lk = (unsigned)w & 0xffu;
                            uint32_t rv = 0xffffffffu;
                            if (!(lk & (LK_NOSEC | LK_PIT))) {
                                const int sec = lk & LK_SEC_MASK;
                                int dr1, dc1, dr2, dc2;
                                off_e1(sec, dr1, dc1); off_e2(sec, dr2, dc2);
                                const uint32_t k1 = (lk & LK_KEEP1) ? (uint32_t)(k + dr1 * HW + dc1) : 0xffffu;
                                const uint32_t k2 = (lk & LK_KEEP2) ? (uint32_t)(k + dr2 * HW + dc2) : 0xffffu;
                                rv = k1 | (k2 << 16);
                            }
                            s.rcv[k] = rv;
                        }
                        s.cnt[k] = c;
                    }
                }
            }
            if (undone_here) atomicAdd(&s.undone, undone_here);
            if (nsrc_here) atomicAdd(&s.nsrc, nsrc_here);
        }
        __syncthreads();
        TS_MARK(2)
        // ---- flow paths
        unsigned notify = 0;
        if (tid == 0) { s.live = s.n_fr; s.head = 0; s.completed = 0; s.x_busy_it = 0; }
        __syncthreads();
        {
            // Every lane follows flow paths.  A warp runs in lock step: lanes holding a cell drain it, lanes
            // without one look for work in the queue -- only every fourth turn while a lane of the warp is on
            // a path (the look costs ~60 cycles on that lane's dependent chain).  A warp leaves, all lanes
            // together, when nothing is queued or held anywhere in the CTA.
            int k = -1, ticket = -1, mine = 0, busy_it = 0;
            long long c_turn = 0, c_drain = 0, c_hand = 0, c_look = 0, c0t = 0;
            for (int it = 0;; it++) {
                const bool any_busy = __any_sync(FULL, k >= 0);
                busy_it += any_busy ? 1 : 0;
                if (a.dbg) c0t = clock64();
                if (!any_busy || (it & 3) == 0) {
                    if (k < 0) {
                        if (ticket < 0) ticket = atomicAdd(&s.head, 1);
                        if (ticket < TN) {
                            const unsigned v = ld_volatile_shared_u16(&s.fr[ticket]);
                            if (v != 0xffffu) { k = (int)v; ticket = -1; fence_cta(); }
                        }
                    }
                    if (!any_busy) {
                        const int live = ld_volatile_shared_i32(&s.live);
                        if (__all_sync(FULL, k < 0)) {
                            if (live == 0) break;
                            __nanosleep(100);      // a warp without work leaves the issue slots to the others
                        }
                    }
                }
                long long c1t = 0, c2t = 0;
                if (a.dbg) c1t = clock64();
                int spare = -1;
                if (k >= 0) {
                    int r1, r2;
                    drain_cell(s, a, k, rows_valid, cols_valid, r0, c0, tile, notify, r1, r2);
                    mine++;
                    if (r1 >= 0 && r2 >= 0) { k = r1; spare = r2; }
                    else if (r1 >= 0) k = r1;
                    else if (r2 >= 0) k = r2;
                    else { k = -1; atomicSub(&s.live, 1); }
                }
                // forks stay in the warp: a D-infinity flow path is a band two or three cells wide whose cells
                // fork and join again on every level; handing the second branch to the queue would make it wait
                // for a lane to look there (hundreds of cycles, on every level of a river).  The j-th forking
                // lane gives its second receiver to the j-th lane without a cell; the queue takes the rest.
                if (a.dbg) c2t = clock64();
                unsigned fm = __ballot_sync(FULL, spare >= 0);
                if (fm) {
                    unsigned im = __ballot_sync(FULL, k < 0);
                    while (fm) {
                        const int src = __ffs(fm) - 1;
                        fm &= fm - 1;
                        const int val = __shfl_sync(FULL, spare, src);
                        if (im) {
                            const int dst = __ffs(im) - 1;
                            im &= im - 1;
                            if (lane == dst) k = val;      // (a ticket it may hold stays valid: it looks there again when this path has ended)
                            if (lane == src) atomicAdd(&s.live, 1);
                        } else if (lane == src) {
                            atomicAdd(&s.live, 1);
                            *reinterpret_cast<volatile uint16_t *>(&s.fr[atomicAdd(&s.n_fr, 1)]) = (uint16_t)val;
                        }
                    }
                    __syncwarp();      // the receiving lanes read what the handing lanes have observed
                }
                if (a.dbg && any_busy) { const long long c3t = clock64(); c_turn += c3t - c0t; c_drain += c2t - c1t; c_hand += c3t - c2t; c_look += c1t - c0t; }
            }
            if (mine) atomicAdd(&s.completed, mine);
            if (a.dbg && lane == 0) { atomicMax(&s.x_busy_it, busy_it); atomicAdd(&s.x_dbg2[0], (unsigned long long)c_turn); atomicAdd(&s.x_dbg2[1], (unsigned long long)c_drain); atomicAdd(&s.x_dbg2[2], (unsigned long long)c_hand); atomicAdd(&s.x_dbg2[3], (unsigned long long)busy_it); }
        }
        __syncthreads();
        // ---- write the cells completed in this visit to global memory
        for (int g = tid; g < TN; g += NT) {
            const int y = g / TW, x = g - y * TW;
            const int k = (y + 1) * HW + x + 1;
            if (s.cnt[k] == TS_MARK_DONE)
                *reinterpret_cast<double2 *>(&a.rec[(r0 + y) * C + (c0 + x)].area) = *reinterpret_cast<const double2 *>(&s.rec[k].area);
        }
        TS_MARK(3)
        // ---- schedule.  The global stores of all threads are ordered before the fence of thread 0 by the CTA
        //      barrier; the notifying atomics follow the fence.  Lane b < 9 of warp 0 looks after neighbour
        //      tile b of the 3x3 block, lane 9 after this tile's own state word: one round trip for all.
        if (notify) atomicOr(&s.notify, notify);
        __syncthreads();
        if (warp == 0) {
            if (lane == 0) fence_visit(a);
            __syncwarp();
            const unsigned nm = s.notify;
            const int completed_all = s.completed, undone_all = s.undone;
            // lanes 0..8: the 3x3 block of this rank's tiles; lanes 10..12 / 13..15: the three tiles of the
            // neighbouring rank above / below (one sweep across GPUs: their state words are peer memory)
            const int y2 = ty + lane / 3 - 1, x2 = tx + lane % 3 - 1;
            const bool mine_local = lane < 9 && lane != 4 && ((nm >> lane) & 1u) && y2 >= 0 && y2 < a.nty && x2 >= 0 && x2 < a.ntx;
            const int side = lane >= 13 ? 1 : 0, x3 = tx + (lane - 10) % 3 - 1;
            const bool mine_peer = lane >= 10 && lane < 16 && ((nm >> lane) & 1u) && ((a.p2p >> side) & 1) && x3 >= 0 && x3 < a.ntx;
            const bool mine = mine_local || mine_peer;
            const int32_t nb = mine_peer ? a.peer_tile0[side] + x3 : y2 * a.ntx + x2;
            uint32_t old = TF_PENDING;
            if (mine_local) old = or_u32(a, &a.flag[nb], TF_PENDING);
            if (mine_peer) old = atomicOr_system(&a.peer_flag[side][nb], TF_PENDING);
            const bool complete = completed_all == undone_all;
            uint32_t own_old = 0;
            if (lane == 9) {
                if (s.late) or_u32(a, &a.flag[tile], TF_PENDING);       // a pit released a receiver in this tile: not released
                own_old = cas_u32(a, &a.flag[tile], TF_RUNNING | TF_VISITED, TF_VISITED | (complete ? TF_COMPLETE : 0u));
            }
            own_old = __shfl_sync(FULL, own_old, 9);
            const bool released = own_old == (TF_RUNNING | TF_VISITED);
            const bool idle = mine && (old & (TF_PENDING | TF_RUNNING | TF_COMPLETE)) == 0;   // I made it pending: I schedule it
            const unsigned im = __ballot_sync(FULL, idle);
            const int n_idle = __popc(im);
            // a released CTA continues with its first idle successor on this rank itself (no queue round trip
            // on the critical path of a river); everything else goes to the queue (a peer's tile: the peer's)
            const int take_lane = (released && (im & 0x1ffu)) ? __ffs(im & 0x1ffu) - 1 : -1;
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
            if (lane == 0 && n_push > 0) (void)add_u64(a, a.inflight, (unsigned long long)n_push);
            if (n_push > 0) { if (lane == 0) fence_visit(a); __syncwarp(); }
            if (idle && lane != take_lane) {
                if (mine_peer) queue_push_to(a, a.peer_ctr[side], a.peer_slots[side], a.peer_cap_mask[side], nb);
                else queue_push(a, nb);
            }
            int next_tile = -1, next_mode = 0, next_first = 0;
            if (released) {
                if (take_lane >= 0) {
                    next_tile = __shfl_sync(FULL, nb, take_lane);
                    next_first = (__shfl_sync(FULL, old, take_lane) & TF_VISITED) ? 0 : 1;
                    next_mode = 1;
                    if (lane == take_lane) exch_u32(a, &a.flag[nb], TF_RUNNING | TF_VISITED);     // claim (thread 0 fences at the top of the visit)
                } else if (lane == 0) {
                    add_u64(a, a.inflight, ~0ULL);   // -1
                }
            } else if (requeue_self) {
                if (lane == 0) {
                    exch_u32(a, &a.flag[tile], TF_PENDING | TF_VISITED);     // still counted in flight
                    queue_push(a, tile);
                    x_defer++;
                }
            } else {
                if (lane == 0) { exch_u32(a, &a.flag[tile], TF_RUNNING | TF_VISITED); x_rerun++; }
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
                if (a.dbg && s.x_is_late) { s.x_late[6] += 1; s.x_late[7] += (unsigned long long)completed_all; s.x_late[4] += (unsigned long long)s.x_busy_it; }
                if (a.dbg && !s.x_is_late) { for (int q = 0; q < 4; q++) s.x_dbg2[q] = 0; }
                x_cells += (unsigned long long)completed_all; 
                x_sources += (unsigned long long)s.nsrc; x_sent += (unsigned long long)((nm >> 9) & 1u);
                if (next_tile < 0) { next_tile = acquire_tile(a); next_mode = 0; }
                s.next_tile = next_tile; s.next_mode = next_mode;
                if (next_mode == 1) s.next_first = next_first; else if (next_mode == 2) s.next_first = 0;
            }
        }
        __syncthreads();
        tile = s.next_tile; mode = s.next_mode;
        TS_MARK(5)
    }
    if (tid == 0) {
        if (x_cells) atomicAdd(&a.ctr[TC_CELLS], x_cells);
        if (x_levels) atomicAdd(&a.ctr[TC_LEVELS], x_levels);
        if (x_sources) atomicAdd(&a.ctr[TC_SOURCES], x_sources);
        if (x_rerun) atomicAdd(&a.ctr[TC_REQUEUE], x_rerun);
        if (x_defer) atomicAdd(&a.ctr[TC_DEFER], x_defer);
        if (x_sent) atomicAdd(&a.ctr[TC_SENT], x_sent);
        atomicMax(&a.ctr[TC_T_END], globaltimer_ns());
        if (a.dbg) {
            for (int q = 0; q < 6; q++) atomicAdd(&a.ctr[TC_PHASE + q], s.x_ph[q]);
            for (int q = 0; q < 8; q++) atomicAdd(&a.ctr[TC_LATE + q], s.x_late[q]);
            for (int q = 0; q < 4; q++) atomicAdd(&a.ctr[TC_DBG2 + q], s.x_dbg2[q]);
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

// Order of the first visits: highest tiles first.  Water runs downhill, so a tile visited after its higher
// neighbours finds most of its ring donors final and completes most of its cells in one visit (simulated on the
// benchmark DEM: 4.0 instead of 6.4 visits per tile, tests/tools/proto_tile_sweep.py).  One warp per tile: key =
// the tile's highest elevation as an order-preserving integer, inverted (ascending radix sort = descending height).
__global__ void __launch_bounds__(256)
k_ts_tile_keys(const double *__restrict__ E, Win w, int tw, int th, int ntx, int ntiles, uint32_t *__restrict__ keys, int32_t *__restrict__ ids)
{
    const int tile = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (tile >= ntiles) return;
    const int ty = tile / ntx, tx = tile - ty * ntx;
    const int64_t r0 = w.lo + (int64_t)ty * th, c0 = (int64_t)tx * tw;
    float m = -3.0e38f;
    for (int y = 0; y < th && r0 + y < w.hi; y += 4)          // every fourth row is plenty for an ordering hint
        for (int x = lane; x < tw && c0 + x < w.C; x += 32) {
            const float e = (float)__ldg(E + (r0 + y) * w.C + c0 + x);
            if (e == e && e > m) m = e;
        }
    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
    if (lane == 0) {
        uint32_t u = __float_as_uint(m);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);      // order-preserving map of floats to unsigned integers
        keys[tile] = ~u;
        ids[tile] = tile;
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
        a.ctr[TC_HEAD] = 0; a.ctr[TC_TAIL] = (unsigned long long)s_count;
        a.ctr[TC_ABORT] = 0; a.ctr[TC_SENT] = 0;
        a.ctr[TC_T_START] = ~0ULL; a.ctr[TC_T_END] = 0;
        if (mode == 0) { a.ctr[TC_VISITS] = 0; a.ctr[TC_CELLS] = 0; a.ctr[TC_SOURCES] = 0; a.ctr[TC_LEVELS] = 0; a.ctr[TC_REQUEUE] = 0; a.ctr[TC_DEFER] = 0; }
        a.ctr[TC_QUEUED] = (unsigned long long)s_count;
        for (int b = 0; b < 128; b++) a.ctr[TC_HIST + b] = 0;
        for (int q = 0; q < 8; q++) { a.ctr[TC_PHASE + q] = 0; a.ctr[TC_LATE + q] = 0; a.ctr[TC_DBG2 + q] = 0; }
        if (a.p2p) {
            // one sweep across GPUs: count this rank's tiles into the global counter (it is back at zero when a
            // sweep ends; rank 0 must add, not set: other ranks may already have counted in), make the queue and
            // the records written by the kernels before this one visible to the peers, then arrive at the start barrier
            atomicAdd_system(a.inflight, (unsigned long long)s_count);
            __threadfence_system();
            atomicAdd_system(a.arrived, 1ULL);
        } else {
            *a.inflight = (unsigned long long)s_count;
        }
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

#define TS_VARIANT(tw, th, nt) {tw, th, nt, sizeof(Smem<tw, th>), k_tsweep<tw, th, nt>, 0}
static Variant g_variants[] = {
    TS_VARIANT(32, 32, 128), TS_VARIANT(32, 32, 64), TS_VARIANT(64, 32, 128), TS_VARIANT(64, 32, 256), TS_VARIANT(64, 64, 256),
    TS_VARIANT(64, 16, 64), TS_VARIANT(32, 16, 64), TS_VARIANT(32, 32, 32),
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
    while (cap < 4 * ntiles || cap < 2 * w.C + 64) cap <<= 1;   // (the slots also serve as the in-box of a multi-GPU work-list sweep: <= 2 C cells)
    if (t->ts_cap < cap || t->ts_ntiles_cap < ntiles) {
        if (t->p2p.on) { pdm_set_error("tile sweep: the control block is shared with peer GPUs and cannot grow; connect after the first sweep geometry is known"); return PDM_ERR_STATE; }
        if (t->ts_ctl) { cudaFree(t->ts_ctl); t->ts_ctl = nullptr; }
        // one allocation [counters | tile flags | queue slots], at least 4 MiB so that it is a cudaMalloc
        // allocation of its own (CUDA IPC shares whole allocations)
        const size_t off_flag = (size_t)TC_N * 8, off_slots = (off_flag + (size_t)ntiles * 4 + 255) & ~(size_t)255;
        size_t bytes = off_slots + (size_t)cap * 4;
        if (bytes < ((size_t)4 << 20)) bytes = (size_t)4 << 20;
        PDM_CUDA(cudaMalloc(&t->ts_ctl, bytes));
        PDM_CUDA(cudaMemsetAsync(t->ts_ctl, 0, bytes, t->stream));
        t->ts_ctr = reinterpret_cast<unsigned long long *>(t->ts_ctl);
        t->ts_flag = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(t->ts_ctl) + off_flag);
        t->ts_slots = reinterpret_cast<int32_t *>(reinterpret_cast<char *>(t->ts_ctl) + off_slots);
        t->ts_cap = cap; t->ts_ntiles_cap = ntiles;
    }
    if (!t->ts_hctr) PDM_CUDA(cudaMallocHost((void **)&t->ts_hctr, TC_N * sizeof(unsigned long long)));
    if (!t->ts_seen) PDM_CUDA(cudaMalloc(&t->ts_seen, (size_t)2 * w.C));
    *out = &v;
    return PDM_OK;
}

static Args ts_args(pdm_tile *t, const Variant &v)
{
    const Win &w = t->win;
    Args a;
    memset(&a, 0, sizeof(a));
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
    a.inflight = t->ts_ctr + TC_INFLIGHT;
    a.arrived = t->ts_ctr + TC_ARRIVED;
    return a;
}

// Args of one sweep across the GPUs of the row shards: the neighbours' boundary rows, tile state words and
// queues are mapped peer memory, the in-flight counter and the start barrier live on rank 0
static void ts_args_p2p(pdm_tile *t, const Variant &v, Args &a)
{
    const pdm_tile::P2P &q = t->p2p;
    a.p2p = 0;
    for (int side = 0; side < 2; side++) {
        if (!q.rec[side]) continue;
        a.p2p |= 1 << side;
        // above: the peer's last owned row; below: its first owned row
        const long long prow = side == 0 ? q.hi[0] - 1 : q.lo[1];
        a.halo_src[side] = reinterpret_cast<const TRec *>(q.rec[side]) + prow * t->win.C;
        char *ctl = reinterpret_cast<char *>(q.ctl[side]);
        a.peer_ctr[side] = reinterpret_cast<unsigned long long *>(ctl);
        a.peer_flag[side] = reinterpret_cast<uint32_t *>(ctl + q.off_flag[side]);
        a.peer_slots[side] = reinterpret_cast<int32_t *>(ctl + q.off_slots[side]);
        a.peer_cap_mask[side] = q.cap_mask[side];
        a.peer_tile0[side] = side == 0 ? (q.nty[0] - 1) * a.ntx : 0;
    }
    unsigned long long *root = reinterpret_cast<unsigned long long *>(q.root_ctl ? q.root_ctl : t->ts_ctl);
    a.inflight = root + TC_INFLIGHT;
    a.arrived = root + TC_ARRIVED;
    a.start_target = (unsigned long long)q.world * q.launches;
    if (!a.p2p) a.p2p = 4;      // a single rank of a p2p group still uses the system-scope protocol
    (void)v;
}

// ---- CUDA IPC plumbing of the multi-GPU sweep (called through pdm_shard_p2p_* in shard.cu)
static int alloc_offset(const void *p, long long *off)
{
    // offset of p inside its cudaMalloc allocation (IPC handles name whole allocations); the driver entry
    // point is fetched through the runtime, the library does not link libcuda
    typedef int (*fn_t)(unsigned long long *, size_t *, unsigned long long);
    static fn_t fn = nullptr;
    if (!fn) {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        PDM_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &qr));
        if (!f) { pdm_set_error("cuMemGetAddressRange not available"); return PDM_ERR_CUDA; }
        fn = (fn_t)f;
    }
    unsigned long long base = 0; size_t size = 0;
    if (fn(&base, &size, (unsigned long long)(uintptr_t)p) != 0) { pdm_set_error("cuMemGetAddressRange failed"); return PDM_ERR_CUDA; }
    *off = (long long)((unsigned long long)(uintptr_t)p - base);
    return PDM_OK;
}

int pdm_ts_p2p_export(pdm_tile *t, P2PExport *e)
{
    Variant *v = nullptr;
    int rc = ts_prepare(t, &v);
    if (rc) return rc;
    memset(e, 0, sizeof(*e));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    PDM_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(e->h_rec), t->cell));
    PDM_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(e->h_ctl), t->ts_ctl));
    if ((rc = alloc_offset(t->cell, &e->off_rec)) || (rc = alloc_offset(t->ts_ctl, &e->off_ctl))) return rc;
    e->off_flag = (long long)(reinterpret_cast<char *>(t->ts_flag) - reinterpret_cast<char *>(t->ts_ctl));
    e->off_slots = (long long)(reinterpret_cast<char *>(t->ts_slots) - reinterpret_cast<char *>(t->ts_ctl));
    e->lo = t->win.lo; e->hi = t->win.hi; e->C = t->win.C;
    e->ntx = (int)((t->win.C + v->tw - 1) / v->tw); e->nty = (int)((t->win.hi - t->win.lo + v->th - 1) / v->th);
    e->cap_mask = (unsigned)(t->ts_cap - 1);
    e->device = t->device;
    e->proc_tag = (long long)getpid();
    e->raw_rec = t->cell; e->raw_ctl = t->ts_ctl;
    return PDM_OK;
}

static int open_peer(const unsigned char *handle, long long off, void **map, void **ptr)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    PDM_CUDA(cudaIpcOpenMemHandle(map, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr = reinterpret_cast<char *>(*map) + off;
    return PDM_OK;
}

// ex[r]: the export of rank r (nullptr = not known; ex[rank] is ignored).  The row neighbours' records and
// control blocks and rank 0's control block are needed; with every rank's control block mapped the work-list
// sweep can span the GPUs as well (sweep.cu, pdm_launch_sweep_p2p).
int pdm_ts_p2p_connect_ranks(pdm_tile *t, const P2PExport *const *ex, int world, int rank)
{
    pdm_tile::P2P &q = t->p2p;
    if (q.on) { pdm_set_error("pdm_shard_p2p_connect: already connected"); return PDM_ERR_STATE; }
    if (world > PDM_MAX_WORLD) { pdm_set_error("pdm_shard_p2p_connect: at most %d ranks", PDM_MAX_WORLD); return PDM_ERR_ARG; }
    Variant *v = nullptr;
    int rc = ts_prepare(t, &v);
    if (rc) return rc;
    memset(&q, 0, sizeof(q));
    const long long me = (long long)getpid();
    for (int r = 0; r < world; r++) {
        if (r == rank) { q.all_ctl[r] = t->ts_ctl; continue; }
        const P2PExport *e = ex[r];
        if (!e) continue;
        if (e->proc_tag == me) q.all_ctl[r] = e->raw_ctl;
        else if ((rc = open_peer(e->h_ctl, e->off_ctl, &q.map_all[r], &q.all_ctl[r]))) return rc;   // (an allocation is opened once per process)
    }
    for (int side = 0; side < 2; side++) {
        const int r = side == 0 ? rank - 1 : rank + 1;
        if (r < 0 || r >= world) continue;
        const P2PExport *e = ex[r];
        if (!e) { pdm_set_error("pdm_shard_p2p_connect: the export of rank %d (a row neighbour) is missing", r); return PDM_ERR_ARG; }
        if (e->C != t->win.C) { pdm_set_error("pdm_shard_p2p_connect: the neighbour has %lld columns, this tile %lld", e->C, (long long)t->win.C); return PDM_ERR_ARG; }
        if (e->proc_tag == me) q.rec[side] = e->raw_rec;
        else {
            void *p = nullptr;
            if ((rc = open_peer(e->h_rec, e->off_rec, &q.map_rec[side], &p))) return rc;
            q.rec[side] = p;
        }
        q.ctl[side] = q.all_ctl[r];
        q.off_flag[side] = e->off_flag; q.off_slots[side] = e->off_slots;
        q.lo[side] = e->lo; q.hi[side] = e->hi; q.nty[side] = e->nty; q.cap_mask[side] = e->cap_mask;
    }
    if (rank > 0) {
        if (!q.all_ctl[0]) { pdm_set_error("pdm_shard_p2p_connect: the export of rank 0 is missing"); return PDM_ERR_ARG; }
        q.root_ctl = q.all_ctl[0];
    }
    q.world = world; q.rank = rank; q.launches = 0; q.on = 1;
    return PDM_OK;
}

// up / down: the exports of the ranks holding the rows above / below (nullptr at the ends of the grid);
// root: rank 0's export (nullptr on rank 0 itself)
int pdm_ts_p2p_connect(pdm_tile *t, const P2PExport *up, const P2PExport *down, const P2PExport *root, int world, int rank)
{
    if (world > PDM_MAX_WORLD) { pdm_set_error("pdm_shard_p2p_connect: at most %d ranks", PDM_MAX_WORLD); return PDM_ERR_ARG; }
    const P2PExport *ex[PDM_MAX_WORLD];
    for (int r = 0; r < PDM_MAX_WORLD; r++) ex[r] = nullptr;
    if (rank > 0) ex[0] = root;
    if (rank > 0) ex[rank - 1] = up;        // (rank 1: rank 0 is the neighbour above; the same export)
    if (rank + 1 < world) ex[rank + 1] = down;
    return pdm_ts_p2p_connect_ranks(t, ex, world, rank);
}

void pdm_ts_p2p_close(pdm_tile *t)
{
    pdm_tile::P2P &q = t->p2p;
    for (int side = 0; side < 2; side++)
        if (q.map_rec[side]) cudaIpcCloseMemHandle(q.map_rec[side]);
    for (int r = 0; r < PDM_MAX_WORLD; r++)
        if (q.map_all[r]) cudaIpcCloseMemHandle(q.map_all[r]);
    memset(&q, 0, sizeof(q));
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

// first != 0: every tile is pending; else (shard resume) the boundary tiles that received donors.
// On a tile connected to its row neighbours (pdm_shard_p2p_connect) the one launch with first != 0 is the
// whole multi-GPU sweep: every rank launches it once, the kernels run until the global counter is at zero.
int pdm_launch_tsweep(pdm_tile *t, int first)
{
    Variant *v = nullptr;
    int rc = ts_prepare(t, &v);
    if (rc) return rc;
    Args a = ts_args(t, *v);
    if (t->p2p.on) {
        if (!first) { pdm_set_error("tile sweep: a multi-GPU sweep has no resume rounds"); return PDM_ERR_STATE; }
        t->p2p.launches++;
        ts_args_p2p(t, *v, a);
    }
    {
        int64_t nfill = (int64_t)a.cap_mask + 1;
        if (first && 2 * a.w.C > nfill) nfill = 2 * a.w.C;
        k_ts_fill<<<(unsigned)((nfill + 255) / 256), 256, 0, t->stream>>>(a, first ? 0 : 1, t->ts_seen);
        PDM_LAUNCHED();
    }
    static int ordered = -1;
    if (ordered < 0) { const char *e = getenv("PYDEM_B200_TS_ORDER"); ordered = (e && !atoi(e)) ? 0 : 1; }
    if (first && ordered && a.ntiles > 1) {
        // queue the tiles from the highest to the lowest (keys, ids, sorted keys + CUB scratch live in the idle work-list queue buffer)
        uint32_t *keys = reinterpret_cast<uint32_t *>(t->queue);
        int32_t *ids = t->queue + a.ntiles;
        uint32_t *keys2 = reinterpret_cast<uint32_t *>(t->queue + 2 * (int64_t)a.ntiles);
        void *tmp = t->queue + 3 * (int64_t)a.ntiles;
        size_t tmp_bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, keys2, ids, t->ts_slots, a.ntiles, 0, 32, t->stream);
        if ((size_t)(3 * (int64_t)a.ntiles) * 4 + tmp_bytes <= (size_t)(t->N + 1) * 4) {
            k_ts_tile_keys<<<(unsigned)(((int64_t)a.ntiles * 32 + 255) / 256), 256, 0, t->stream>>>(t->elev, a.w, v->tw, v->th, a.ntx, a.ntiles, keys, ids);
            PDM_LAUNCHED();
            PDM_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys2, ids, t->ts_slots, a.ntiles, 0, 32, t->stream));
            g_pdm_launches += 4;
            t->queue_ready = false;   // the work-list's queue array served as scratch
        }
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
