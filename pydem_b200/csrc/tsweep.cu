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
//   * the grid is cut into TW x TH tiles; a CTA that owns a tile stages the tile and a one-cell
//     ring of its neighbours (area, taint, proportion, link byte) in shared memory;
//   * every undone cell counts its donors that are not done yet; cells at zero form the first
//     frontier; a ready cell PULLS: area = own cell area + sum over its done donors, in a fixed
//     neighbour order (W, E, N, S, NW, NE, SW, SE) -- deterministic, bit-reproducible sums; it
//     then decrements its in-tile receivers' counters (shared-memory integer atomics), and
//     receivers reaching zero form the next frontier.  Small frontiers (rivers) are advanced by
//     one warp with warp-level synchronisation only: a level costs shared-memory latency;
//   * completed cells are written back; a cell with a receiver in another tile marks that tile
//     for a (re)visit.  Tiles are scheduled asynchronously through a ticket queue in global
//     memory by a persistent grid; there is no grid-wide barrier.  A per-tile state word
//     {pending, running, complete} guarantees that a tile is run by one CTA at a time and is run
//     again whenever a neighbour published new donors after it was loaded;
//   * publication follows the PTX memory model: plain stores, __threadfence() by every storing
//     thread, CTA barrier, then the notifying atomic; the consumer's atomic on the tile state is
//     followed by a fence and a CTA barrier before anything is loaded (L2 loads, ld.cg).
//
// "done" is encoded in the area itself: UCA holds a signalling-NaN pattern no arithmetic produces
// until the cell's sum is final, so a cell's state is one 8-byte word and the sweep writes its
// result straight into the output field.
//
// Long-range pit edges (_mk_connectivity_pits, dem_processing.py:1269-1382) are rare and keep a
// push form: a drained pit adds into per-receiver accumulators with global atomics, fences, and
// decrements the receiver's pit counter; the receiver is gated on that counter and adds the
// accumulator into its pull sum.
//
// Row shards: tiles cover the owned rows; halo rows are ring cells only.  A completed boundary
// cell whose receiver belongs to the neighbouring rank is counted (TC_SENT); the host exchanges
// boundary rows of UCA / taint and resumes the tiles whose ring received new donors, until no rank
// completed such a cell (sharded.py).
#include <stdlib.h>

#include "tsweep.cuh"

namespace ts {

#define ST_NEW 0x01    // completed during this visit (store phase)
#define ST_TODO 0x02   // inflow-border cell: initial taint 1 (dem_processing.py:909-944)

#define TF_PENDING 1u
#define TF_RUNNING 2u
#define TF_COMPLETE 4u
#define TF_VISITED 8u

__device__ __forceinline__ int32_t ld_volatile_i32(const int32_t *p)
{
    int32_t v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ bool is_done(double a) { return (unsigned long long)__double_as_longlong(a) != TS_NOT_DONE; }

// does a neighbour with link byte lk drain into the centre cell?  (keepbit: the receiver slot the
// centre occupies -- cardinal e1 or diagonal e2; secmask: the facets whose slot points at it)
__device__ __forceinline__ bool drains_in(uint8_t lk, uint8_t keepbit, uint32_t secmask)
{
    return (lk & keepbit) && !(lk & (LK_NOSEC | LK_PIT)) && ((secmask >> (lk & LK_SEC_MASK)) & 1u);
}

template <int TW_, int TH_>
struct Smem {
    static constexpr int TW = TW_, TH = TH_, HW = TW_ + 2, HH = TH_ + 2, HN = HW * HH, TN = TW_ * TH_;
    double area[HN];
    double taint[HN];
    double prop[HN];
    uint32_t cnt[TN / 4];        // per own cell: donors not done yet (+1 while its pit counter is non-zero), packed bytes
    uint16_t list[2][TN];        // frontier double buffer (ring-cell indices)
    uint8_t link[(HN + 15) / 16 * 16];
    uint8_t st[TN];
    int n[3];                    // rotating frontier counters: level L reads n[L%3], appends to n[(L+1)%3], clears n[(L+2)%3]
    int lvl;
    int tile;
    unsigned notify;             // 3x3 bit mask of neighbour tiles that received a new donor
    int sent;                    // completed cells whose receiver lives on the neighbouring rank
    int undone;                  // own cells not done when the tile was loaded
    int nsrc;                    // own cells nobody drains into (first visit)
    int completed;
};

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

// thread 0: wait for a queue item (ticket = fetch-and-add, never retries).  -1: the sweep is over.
__device__ int32_t acquire_tile(const Args &a)
{
    const unsigned long long h = atomicAdd(&a.ctr[TC_HEAD], 1ULL);
    int32_t *slot = a.slots + (h & a.cap_mask);
    unsigned ns = 32, polls = 0;
    unsigned long long t_last = 0, v_last = 0;
    for (;;) {
        int32_t v = ld_volatile_i32(slot);
        if (v >= 0) {
            v = atomicExch(slot, -1);
            if (v >= 0) return v;
        }
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

// drained pit: push along its long-range edges (rare; plain fence ordering)
template <class S>
__device__ __noinline__ void pit_push(S &s, const Args &a, int k, double ar, double tt, int rows_valid, int cols_valid,
                                      int64_t r0, int64_t c0, int my_tile, uint16_t *nxt, int *nn)
{
    const int64_t slot = __double_as_longlong(s.prop[k]);
    const int32_t e0 = a.pit_beg[slot], e1 = a.pit_end[slot];
    for (int32_t e = e0; e < e1; e++) {
        const int32_t r = a.pit_dst[e];
        const double w = a.pit_w[e];
        atomicAdd(&a.pit_acc_a[r], __dmul_rn(ar, w));
        if (tt != 0.0) atomicAdd(&a.pit_acc_t[r], __dmul_rn(tt, w));
    }
    __threadfence();
    for (int32_t e = e0; e < e1; e++) {
        const int32_t r = a.pit_dst[e];
        if (atomicSub(&a.pit_cnt[r], 1) != 1) continue;
        __threadfence();   // the last decrement has observed all others: pass their accumulator adds on
        const int64_t ri = r / a.w.C, rj = r - ri * a.w.C;
        if (ri < a.w.lo || ri >= a.w.hi) continue;       // (pit edges never cross a shard: refused at graph build)
        const int y = (int)((ri - a.w.lo) / S::TH) * a.ntx + (int)(rj / S::TW);
        if (y == my_tile) {
            const int ry = (int)(ri - r0) + 1, rx = (int)(rj - c0) + 1;
            if (ry >= 1 && ry <= rows_valid && rx >= 1 && rx <= cols_valid) {
                const int o = (ry - 1) * S::TW + (rx - 1);
                const int sh = (o & 3) * 8;
                const uint32_t old = atomicSub(&s.cnt[o >> 2], 1u << sh);
                if (((old >> sh) & 0xffu) == 1u) nxt[atomicAdd(nn, 1)] = (uint16_t)(ry * S::HW + rx);
            }
        } else if (notify_tile(a, y)) {
            queue_push(a, y);
        }
    }
}

// one ready cell: pull the donors' contributions, publish in shared memory, release the receivers
template <class S>
__device__ __forceinline__ void process_cell(S &s, const Args &a, int k, int rows_valid, int cols_valid, int64_t r0,
                                             int64_t c0, int my_tile, uint16_t *nxt, int *nn)
{
    constexpr int HW = S::HW, TW = S::TW;
    const int hy = k / HW, hx = k - hy * HW;
    const int o = (hy - 1) * TW + (hx - 1);
    const int64_t gi = r0 - 1 + hy;
    const uint8_t lk = s.link[k];
    double ar = __ldg(a.row_area + gi);                                      // dem_processing.py:885, 901
    double tt = (s.st[o] & ST_TODO) ? 1.0 : 0.0;                             // 944
#define TS_PULL(dk, keepbit, secmask, diag)                                                   \
    {                                                                                         \
        const uint8_t b = s.link[k + (dk)];                                                   \
        if (drains_in(b, keepbit, secmask)) {                                                 \
            const double p = s.prop[k + (dk)];                                                \
            const double wgt = (diag) ? __dsub_rn(1.0, p) : p;     /* dem_processing.py:1082 */ \
            ar = __dadd_rn(ar, __dmul_rn(s.area[k + (dk)], wgt));  /* cyutils.pyx:161 */      \
            const double td = s.taint[k + (dk)];                                              \
            if (td != 0.0) tt = __dadd_rn(tt, __dmul_rn(td, wgt)); /* cyutils.pyx:163 */      \
        }                                                                                     \
    }
    TS_PULL(-1, LK_KEEP1, 0x81u, false)          // W neighbour: its e1 = (0,+1) for facets 0, 7
    TS_PULL(+1, LK_KEEP1, 0x18u, false)          // E: e1 = (0,-1) for facets 3, 4
    TS_PULL(-HW, LK_KEEP1, 0x60u, false)         // N: e1 = (+1,0) for facets 5, 6
    TS_PULL(+HW, LK_KEEP1, 0x06u, false)         // S: e1 = (-1,0) for facets 1, 2
    TS_PULL(-HW - 1, LK_KEEP2, 0xC0u, true)      // NW: e2 = (+1,+1) for facets 6, 7
    TS_PULL(-HW + 1, LK_KEEP2, 0x30u, true)      // NE: e2 = (+1,-1) for facets 4, 5
    TS_PULL(+HW - 1, LK_KEEP2, 0x03u, true)      // SW: e2 = (-1,+1) for facets 0, 1
    TS_PULL(+HW + 1, LK_KEEP2, 0x0Cu, true)      // SE: e2 = (-1,-1) for facets 2, 3
#undef TS_PULL
    if (lk & LK_PITIN) {
        const int64_t n = gi * a.w.C + (c0 - 1 + hx);
        ar = __dadd_rn(ar, __ldcg(a.pit_acc_a + n));
        const double td = __ldcg(a.pit_acc_t + n);
        if (td != 0.0) tt = __dadd_rn(tt, td);
    }
    s.area[k] = ar;
    s.taint[k] = tt;
    s.st[o] |= ST_NEW;
    if (lk & LK_PIT) { pit_push(s, a, k, ar, tt, rows_valid, cols_valid, r0, c0, my_tile, nxt, nn); return; }
    if (lk & LK_NOSEC) return;
    const int sec = lk & LK_SEC_MASK;
#pragma unroll
    for (int e = 0; e < 2; e++) {
        if (!(lk & (e == 0 ? LK_KEEP1 : LK_KEEP2))) continue;
        int dr, dc;
        if (e == 0) off_e1(sec, dr, dc); else off_e2(sec, dr, dc);
        const int ry = hy + dr, rx = hx + dc;
        if (ry >= 1 && ry <= rows_valid && rx >= 1 && rx <= cols_valid) {
            const int ro = (ry - 1) * TW + (rx - 1);
            const int sh = (ro & 3) * 8;
            const uint32_t old = atomicSub(&s.cnt[ro >> 2], 1u << sh);
            if (((old >> sh) & 0xffu) == 1u) nxt[atomicAdd(nn, 1)] = (uint16_t)(ry * HW + rx);
        } else {
            const int64_t gr = r0 - 1 + ry;
            if (gr < a.w.lo || gr >= a.w.hi) atomicAdd(&s.sent, 1);      // receiver on the neighbouring rank
            else {
                const int tdy = ry < 1 ? 0 : (ry > rows_valid ? 2 : 1), tdx = rx < 1 ? 0 : (rx > cols_valid ? 2 : 1);
                atomicOr(&s.notify, 1u << (tdy * 3 + tdx));
            }
        }
    }
}

template <int TW, int TH, int NT>
__global__ void __launch_bounds__(NT) k_tsweep(const Args a)
{
    typedef Smem<TW, TH> S;
    constexpr int HW = S::HW, HN = S::HN, TN = S::TN;
    extern __shared__ __align__(16) unsigned char ts_raw[];
    S &s = *reinterpret_cast<S *>(ts_raw);
    const int tid = threadIdx.x;
    const int64_t C = a.w.C;
    if (tid == 0) {
        atomicMin(&a.ctr[TC_T_START], globaltimer_ns());
        s.tile = acquire_tile(a);
    }
    __syncthreads();
    int tile = s.tile;
    bool fresh = true;     // tile came from the queue / a neighbour: its state word still has to be claimed
    while (tile >= 0) {
        // ---- claim: pending -> running.  Everything published before the notification that made the
        //      tile pending is visible after this atomic + fence + barrier.
        if (tid == 0) {
            s.n[0] = s.n[1] = s.n[2] = 0;
            s.notify = 0; s.sent = 0; s.undone = 0; s.nsrc = 0; s.completed = 0; s.lvl = 0;
            uint32_t old = TF_VISITED;
            if (fresh) old = atomicExch(&a.flag[tile], TF_RUNNING | TF_VISITED);
            s.tile = (old & TF_VISITED) ? 1 : 0;      // (reused as "visited before" until the end of the visit)
            __threadfence();
        }
        __syncthreads();
        const bool first_visit = s.tile == 0;
        const int ty = tile / a.ntx, tx = tile - ty * a.ntx;
        const int64_t r0 = a.w.lo + (int64_t)ty * TH, c0 = (int64_t)tx * TW;
        const int rows_valid = (int)min((int64_t)TH, a.w.hi - r0), cols_valid = (int)min((int64_t)TW, C - c0);
        // ---- load the tile and its ring
        for (int k = tid; k < HN; k += NT) {
            const int hy = k / HW, hx = k - hy * HW;
            const int64_t gi = r0 - 1 + hy, gj = c0 - 1 + hx;
            const bool inside = gi >= 0 && gi < a.w.R && gj >= 0 && gj < C && hy <= rows_valid + 1 && hx <= cols_valid + 1;
            uint8_t lk = LK_NOSEC;
            double ar = 0.0, tt = 0.0, p = 0.0;
            if (inside) {
                const int64_t n = gi * C + gj;
                lk = a.link[n];
                p = a.prop[n];
                ar = __ldcg(a.area + n);
                if (is_done(ar)) {
                    // area and taint are two independent 8-byte words, each written exactly once
                    // (NOT_DONE -> final): a cell counts as done only when both are final, so a tile
                    // that is loaded while its neighbour is still writing never pairs a final area
                    // with a stale taint -- without ordering the two stores
                    tt = __ldcg(a.taint + n);
                    if (!is_done(tt)) { ar = tt; tt = 0.0; }
                }
            }
            s.link[k] = lk; s.area[k] = ar; s.taint[k] = tt; s.prop[k] = p;
        }
        __syncthreads();
        // ---- donors not done yet; the first frontier
        {
            int undone = 0, nsrc = 0;
            for (int o = tid; o < TN; o += NT) {
                const int y = o / TW, x = o - y * TW;
                const int k = (y + 1) * HW + (x + 1);
                uint8_t st = 0;
                int cnt = 0;
                if (y < rows_valid && x < cols_valid && !is_done(s.area[k])) {
                    const int64_t gi = r0 + y, gj = c0 + x;
                    int donors = 0;
#define TS_CNT(dk, keepbit, secmask)                                         \
    if (drains_in(s.link[k + (dk)], keepbit, secmask)) {                     \
        donors++;                                                            \
        if (!is_done(s.area[k + (dk)])) cnt++;                               \
    }
                    TS_CNT(-1, LK_KEEP1, 0x81u) TS_CNT(+1, LK_KEEP1, 0x18u) TS_CNT(-HW, LK_KEEP1, 0x60u) TS_CNT(+HW, LK_KEEP1, 0x06u)
                    TS_CNT(-HW - 1, LK_KEEP2, 0xC0u) TS_CNT(-HW + 1, LK_KEEP2, 0x30u) TS_CNT(+HW - 1, LK_KEEP2, 0x03u)
                    TS_CNT(+HW + 1, LK_KEEP2, 0x0Cu)
#undef TS_CNT
                    const uint8_t lk = s.link[k];
                    if (lk & LK_PITIN) {
                        donors++;
                        if (ld_volatile_i32(a.pit_cnt + gi * C + gj) != 0) cnt++;
                    }
                    const bool border = gj == 0 || gj == C - 1 || a.w.top(gi) || a.w.bottom(gi);
                    if (border && a.edge_todo[gi * C + gj]) st = ST_TODO;
                    undone++;
                    if (donors == 0) nsrc++;
                    if (cnt == 0) s.list[0][atomicAdd(&s.n[0], 1)] = (uint16_t)k;
                }
                reinterpret_cast<uint8_t *>(s.cnt)[o] = (uint8_t)cnt;
                s.st[o] = st;
            }
            for (int d = 16; d > 0; d >>= 1) {
                undone += __shfl_down_sync(0xffffffffu, undone, d);
                nsrc += __shfl_down_sync(0xffffffffu, nsrc, d);
            }
            if ((tid & 31) == 0) {
                if (undone) atomicAdd(&s.undone, undone);
                if (nsrc && first_visit) atomicAdd(&s.nsrc, nsrc);
            }
        }
        // the pit counters were read before this CTA drains any pit: a gate found closed is opened
        // either by this CTA (pit_push, local) or by another one, which then re-notifies this tile
        __threadfence_block();
        // ---- frontier loop.  Level L: cells of list[L&1] (n[L%3] of them) pull and release.
        int lvl = 0;
        for (;;) {
            __syncthreads();
            const int n = s.n[lvl % 3];
            if (n == 0) break;
            if (n <= 32) {
                // river mode: one warp advances the frontier with warp-level synchronisation until it
                // empties or widens; the other warps wait at the barrier below
                if (tid < 32) {
                    int l = lvl, m = n;
                    while (m > 0 && m <= 32) {
                        if (tid == 0) s.n[(l + 2) % 3] = 0;
                        __syncwarp();
                        if (tid < m) process_cell(s, a, s.list[l & 1][tid], rows_valid, cols_valid, r0, c0, tile,
                                                  s.list[(l + 1) & 1], &s.n[(l + 1) % 3]);
                        __syncwarp();
                        l++;
                        m = *(volatile int *)&s.n[l % 3];
                    }
                    if (tid == 0) s.lvl = l;
                }
                __syncthreads();
                lvl = s.lvl;
                continue;
            }
            if (tid == 0) s.n[(lvl + 2) % 3] = 0;
            const uint16_t *cur = s.list[lvl & 1];
            for (int idx = tid; idx < n; idx += NT)
                process_cell(s, a, cur[idx], rows_valid, cols_valid, r0, c0, tile, s.list[(lvl + 1) & 1], &s.n[(lvl + 1) % 3]);
            lvl++;
        }
        // ---- write the completed cells back
        {
            int completed = 0;
            for (int o = tid; o < TN; o += NT) {
                if (s.st[o] & ST_NEW) {
                    const int y = o / TW, x = o - y * TW;
                    const int k = (y + 1) * HW + (x + 1);
                    const int64_t n = (r0 + y) * C + (c0 + x);
                    a.taint[n] = s.taint[k];
                    a.area[n] = s.area[k];
                    completed++;
                }
            }
            for (int d = 16; d > 0; d >>= 1) completed += __shfl_down_sync(0xffffffffu, completed, d);
            if ((tid & 31) == 0 && completed) atomicAdd(&s.completed, completed);
        }
        __threadfence();
        __syncthreads();
        // ---- schedule: neighbours that received donors, then this tile's own state
        if (tid == 0) {
            int next = -1;
            const unsigned nm = s.notify;
            if (nm) {
                for (int b = 0; b < 9; b++) {
                    if (!((nm >> b) & 1u) || b == 4) continue;
                    const int y2 = ty + b / 3 - 1, x2 = tx + b % 3 - 1;
                    if (y2 < 0 || y2 >= a.nty || x2 < 0 || x2 >= a.ntx) continue;
                    const int32_t nb = y2 * a.ntx + x2;
                    if (notify_tile(a, nb)) {
                        if (next < 0) next = nb; else queue_push(a, nb);
                    }
                }
            }
            atomicAdd(&a.ctr[TC_VISITS], 1ULL);
            if (s.completed) atomicAdd(&a.ctr[TC_CELLS], (unsigned long long)s.completed);
            if (s.sent) atomicAdd(&a.ctr[TC_SENT], (unsigned long long)s.sent);
            if (s.nsrc) atomicAdd(&a.ctr[TC_SOURCES], (unsigned long long)s.nsrc);
            if (lvl) atomicAdd(&a.ctr[TC_LEVELS], (unsigned long long)lvl);
            const bool complete = s.completed == s.undone;
            const uint32_t idle = TF_VISITED | (complete ? TF_COMPLETE : 0u);
            const uint32_t old = atomicCAS(&a.flag[tile], TF_RUNNING | TF_VISITED, idle);
            if (old == (TF_RUNNING | TF_VISITED)) {
                // released.  Successors were counted in-flight above, so the count cannot touch 0 early.
                atomicAdd(&a.ctr[TC_INFLIGHT], ~0ULL);   // -1
                s.tile = next >= 0 ? next : acquire_tile(a);
                s.lvl = 1;   // fresh
            } else {
                // a neighbour published donors while this tile was loaded: run it again
                atomicExch(&a.flag[tile], TF_RUNNING | TF_VISITED);
                atomicAdd(&a.ctr[TC_REQUEUE], 1ULL);
                if (next >= 0) queue_push(a, next);
                s.tile = tile;
                s.lvl = 0;
            }
        }
        __syncthreads();
        tile = s.tile;
        fresh = s.lvl != 0;
        __syncthreads();
    }
    if (tid == 0) atomicMax(&a.ctr[TC_T_END], globaltimer_ns());
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
                    if (is_done(a.area[hrow * C + j]) && !seen[side * C + j]) { fresh = true; break; }
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
            for (int64_t j = tid; j < C; j += blockDim.x) seen[side * C + j] = is_done(a.area[hrow * C + j]) ? 1 : 0;
        }
    }
    __syncthreads();
    if (tid == 0) {
        a.ctr[TC_HEAD] = 0; a.ctr[TC_TAIL] = (unsigned long long)s_count; a.ctr[TC_INFLIGHT] = (unsigned long long)s_count;
        a.ctr[TC_ABORT] = 0; a.ctr[TC_SENT] = 0;
        a.ctr[TC_T_START] = ~0ULL; a.ctr[TC_T_END] = 0;
        if (mode == 0) { a.ctr[TC_VISITS] = 0; a.ctr[TC_CELLS] = 0; a.ctr[TC_SOURCES] = 0; a.ctr[TC_LEVELS] = 0; a.ctr[TC_REQUEUE] = 0; }
        a.ctr[TC_QUEUED] = (unsigned long long)s_count;
    }
}

// a6 epilogue: dem_processing.py:966-980 on the owned cells [n0, n1)
__global__ void __launch_bounds__(256)
k_ts_finalize(const double *__restrict__ E, const uint8_t *__restrict__ flats, const double *__restrict__ taint,
              double *__restrict__ uca, uint8_t *__restrict__ edge_done, int64_t n0, int64_t n1, int limit_edges,
              double limit_area, unsigned long long *counters)
{
    const int64_t n = n0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool undone = false;
    if (n < n1) {
        double u = uca[n];
        undone = !is_done(u);
        // cells on circular references were never reached: they keep their own cell area until the
        // restart pass (restart.cu) has dealt with them
        const double tt = undone ? 0.0 : taint[n];
        if (!undone) {
            if (flats[n]) { u = __longlong_as_double(0x7ff8000000000000LL); uca[n] = u; }        // 972
            bool ed = !(tt != 0.0);                                                      // 969, 974
            const double e = E[n];
            if (e != e) ed = true;                                                       // 975
            if (limit_edges && u > limit_area) ed = true;                                // 977-980
            edge_done[n] = ed ? 1 : 0;
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

__global__ void __launch_bounds__(256) k_fill_u64(unsigned long long *p, int64_t n, unsigned long long v)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct Variant { int tw, th, nt; size_t smem; void (*kernel)(const Args); int blocks; };

static Variant g_variants[] = {
    {32, 32, 256, sizeof(Smem<32, 32>), k_tsweep<32, 32, 256>, 0},
    {64, 32, 256, sizeof(Smem<64, 32>), k_tsweep<64, 32, 256>, 0},
    {64, 64, 512, sizeof(Smem<64, 64>), k_tsweep<64, 64, 512>, 0},
    {32, 16, 128, sizeof(Smem<32, 16>), k_tsweep<32, 16, 128>, 0},
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
    a.link = t->link; a.prop = t->twi; a.area = t->uca; a.taint = pdm_taint(t); a.edge_todo = t->edge_todo;
    a.row_area = t->row_area;
    a.pit_beg = t->pit_beg; a.pit_end = t->pit_end; a.pit_dst = t->pit_dst; a.pit_w = t->pit_w;
    a.pit_cnt = t->label; a.pit_acc_a = pdm_taint(t) + t->N; a.pit_acc_t = pdm_taint(t) + 2 * t->N;
    a.w = w;
    a.ntx = (int32_t)((w.C + v.tw - 1) / v.tw); a.nty = (int32_t)((w.hi - w.lo + v.th - 1) / v.th);
    a.ntiles = a.ntx * a.nty;
    a.slots = t->ts_slots; a.cap_mask = (uint32_t)(t->ts_cap - 1); a.flag = t->ts_flag; a.ctr = t->ts_ctr;
    return a;
}

// the sweep state of a fresh graph: every local cell "not done" (halo rows included), pit
// accumulators at zero, pit receivers marked
int pdm_ts_reset_state(pdm_tile *t)
{
    k_fill_u64<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>((unsigned long long *)t->uca, t->N, TS_NOT_DONE);
    PDM_LAUNCHED();
    k_fill_u64<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>((unsigned long long *)pdm_taint(t), t->N, TS_NOT_DONE);
    PDM_LAUNCHED();
    if (t->n_pit_edges > 0) {
        PDM_CUDA(cudaMemsetAsync(pdm_taint(t) + t->N, 0, (size_t)t->N * 16, t->stream));
        k_pit_mark<<<(unsigned)((t->n_pit_edges + 255) / 256), 256, 0, t->stream>>>(t->pit_dst, t->n_pit_edges, t->link);
        PDM_LAUNCHED();
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
        t->elev, t->flats, pdm_taint(t), t->uca, t->edge_done, n0, n1, p->apply_uca_limit_edges, limit_area, t->d_counters);
    PDM_LAUNCHED();
    return PDM_OK;
}
