// drain_op.cuh -- the per-cell "drain" step of the accumulation sweep (cyutils.pyx:149-185)
// as a work-list operator (worklist.cuh).
//
//   MODE 0: full sweep of _calc_uca_chunk: seeds = cells nobody drains into, the
//           edge_todo taint travels with the area (cyutils.pyx:163).
//   MODE 1: update sweep of _calc_uca_chunk_update (dem_processing.py:836-842): seeds = the
//           border cells whose neighbour value became final (ST_START), `area` is the delta
//           array, and pushes into a start cell are dropped -- the reference's
//           "(skip_edge or done[r]) and r on the tile edge" rule (cyutils.pyx:157-159), which
//           in update mode fires exactly for the start cells.
#pragma once
#include "worklist.cuh"

// per-cell state byte of the update mode (t->flat0 is reused for it)
#define ST_START 0x01     // ids: edge_done & edge_todo (dem_processing.py:798)
#define ST_CONE 0x02      // downstream of a start cell (drain_connections, 823-824)
#define ST_TODOSEED 0x04  // edge_todo & ~edge_done (799)
#define ST_REACH 0x08     // downstream closure of the remaining todo cells (848-853)

__device__ __forceinline__ uint32_t st_fetch_or(uint8_t *st, int32_t cell, uint32_t bits)
{
    uint32_t *w = reinterpret_cast<uint32_t *>(st) + (cell >> 2);
    const int sh = (cell & 3) * 8;
    const uint32_t old = atomicOr(w, bits << sh);
    return (old >> sh) & 0xffu;
}

//   MODE 2: resumed full sweep of a shard: like MODE 0, but the seeds come from an explicit list
//           (cells whose last upstream contribution arrived from a neighbouring rank).
//   MODE 3: full sweep of a row shard as part of ONE sweep across several GPUs: like MODE 0, but a
//           receiver in a halo row is the neighbouring rank's boundary cell -- the push goes into that
//           rank's record over NVLink (push_peer) and a cell it makes ready into that rank's in-box
//           (worklist.cuh, p2p_service).
// Predicated atomics without branches.  Written as inline PTX so that the step below stays
// straight-line code: with `if (k1) o1 = atomicSub(..); if (k2) o2 = atomicSub(..);` the compiler
// wrapped each atomic in its own branch region and placed the test of the first result inside it,
// i.e. the second decrement was only issued after the first had returned (seen in SASS: ATOMG, ISETP
// on its result, then the next ATOMG) -- three dependent round trips per cell instead of two.
__device__ __forceinline__ void red_add_f64_if(bool on, double *addr, double v)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %0, 0;\n\t@q red.global.add.f64 [%1], %2;\n\t}"
                 :: "r"((int)on), "l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ int atom_add_s32_if(bool on, int32_t *addr, int v)
{
    int old = 0;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t@q atom.global.add.s32 %0, [%2], %3;\n\t}"
                 : "+r"(old) : "r"((int)on), "l"(addr), "r"(v) : "memory");
    return old;
}

#ifndef PDM_BURST_MAX
#define PDM_BURST_MAX 31
#endif
#ifndef PDM_FORK_STEPS
#define PDM_FORK_STEPS 3     // cells of a side branch drained in place at a fork on a river
#endif
#ifndef PDM_BURST_MIN
#define PDM_BURST_MIN 4
#endif

template <int MODE>
struct DrainOp {
    static constexpr bool P2P = MODE == 3;
    const uint8_t *link;   // SoA copy of the link bytes: only the seed scan reads it
    Cell *cell;            // 32-byte sweep records
    const uint8_t *st;
    int32_t C;
    const int32_t *pit_beg;
    const int32_t *pit_end;
    const int32_t *pit_dst;
    const double *pit_w;
    int32_t strict;        // hold the decrements back until the adds have returned (see process())
    int32_t burst;         // the express chain advances by warp-wide bursts along single-receiver runs (chain_warp)
    int32_t nrows;         // rows of the tile (chain_warp's L1 warm-up stays inside the link plane)
    // MODE 3 only
    int32_t own_lo, own_hi;              // owned cells [own_lo, own_hi); the rows next to them are halo rows
    Cell *peer_cell[2];                  // the neighbour's record behind column 0 of the halo row above / below
    unsigned long long *peer_ctr[2];     // its counters and in-box
    int32_t *peer_inbox[2];
    int32_t peer_cell0[2];               // the neighbour's own index of that record

    __device__ __forceinline__ bool is_seed(int32_t c) const
    {
        return (MODE == 0 || MODE == 3) ? (link[c] & LK_SOURCE) != 0 : (MODE == 2 ? true : (st[c] & ST_START) != 0);
    }
    // MODE 3: receiver r lives on the neighbouring GPU.  The adds must be part of the record before the
    // decrement that may publish it: over NVLink that order is enforced with a system-scope fence (rare: only
    // cells of a boundary row come here).  The decrement's return value is back before the chain goes on,
    // so the in-box push below is counted before this chain can be counted as finished.
    __device__ __forceinline__ void push_peer(int32_t r, double da, double dt, bool tt) const
    {
        const int side = r < own_lo ? 0 : 1;
        const int32_t col = r - (side == 0 ? own_lo - C : own_hi);
        push_peer_at(side, peer_cell0[side] + col, da, dt, tt);
    }
    // idx: the cell's index on the neighbouring rank (boundary row: push_peer; any row: pit drains across the boundary)
    __device__ __forceinline__ void push_peer_at(int side, int32_t idx, double da, double dt, bool tt) const
    {
        Cell *P = peer_cell[side] + (idx - peer_cell0[side]);
        atomicAdd_system(&P->area, da);
        if (tt) atomicAdd_system(&P->taint, dt);
        __threadfence_system();
        if (atomicAdd_system(&P->indeg, -1) == 1) {
            const unsigned long long s = atomicAdd_system(peer_ctr[side] + CT_INBOX_TAIL, 1ULL);
            wl::st_volatile_i32(peer_inbox[side] + s, idx);
        }
    }
    __device__ __forceinline__ bool skip(int32_t r) const { return MODE == 1 && (st[r] & ST_START); }

    // Drain one ready cell; returns the receiver this lane continues with (or -1).
    __device__ __forceinline__ int32_t process(int32_t i, const wl::Queue &q, int32_t &defer) const
    {
        // the cell's whole sweep state is one 32-byte sector: two 16-byte L2 loads
        const double2 at = __ldcg(reinterpret_cast<const double2 *>(&cell[i].area));      // area, taint
        const longlong2 pm = __ldcg(reinterpret_cast<const longlong2 *>(&cell[i].prop));  // prop | indeg, link
        return drain(i, at.x, MODE != 1 ? at.y : 0.0, __longlong_as_double(pm.x), (uint8_t)((unsigned long long)pm.y >> 32), q, defer);
    }

    // the same with the cell's record (final area ai, taint ti, proportion p, link byte lk) already in registers
    __device__ __forceinline__ int32_t drain(int32_t i, double ai, double ti, double p, uint8_t lk, const wl::Queue &q, int32_t &defer) const
    {
        int32_t nxt = -1;
        if (lk & LK_PIT) {
            // long-range pit edges (_mk_connectivity_pits): rare, plain fence ordering
            const int64_t slot = __double_as_longlong(p);
            const int32_t e0 = pit_beg[slot], e1 = pit_end[slot];
            for (int32_t e = e0; e < e1; e++) {
                const int32_t r = pit_dst[e];
                const double w = pit_w[e];
                if (MODE == 3 && r < 0) {
                    // the drain is a cell of the neighbouring rank (pits.cu, k_pit_search<true>): ~r = side << 30 | its index there
                    push_peer_at((~r) >> 30, (~r) & 0x3fffffff, __dmul_rn(ai, w), __dmul_rn(ti, w), ti != 0.0);
                    continue;
                }
                if (skip(r)) continue;
                atomicAdd(&cell[r].area, __dmul_rn(ai, w));
                if (MODE != 1 && ti != 0.0) atomicAdd(&cell[r].taint, __dmul_rn(ti, w));
            }
            __threadfence();
            for (int32_t e = e0; e < e1; e++) {
                const int32_t r = pit_dst[e];
                if (MODE == 3 && r < 0) continue;
                if (skip(r)) continue;
                if (atomicSub(&cell[r].indeg, 1) == 1) {
                    if (nxt < 0) nxt = r; else q.push(r);
                }
            }
            return nxt;
        }
        bool k1 = lk & LK_KEEP1, k2 = lk & LK_KEEP2;
        if (!(k1 || k2)) return -1;
        const int sec = lk & LK_SEC_MASK;
        const int32_t r1 = i + wl::off_e1(sec, C), r2 = i + wl::off_e2(sec, C);
        if (MODE == 1) {
            if (k1 && skip(r1)) k1 = false;
            if (k2 && skip(r2)) k2 = false;
            if (!(k1 || k2)) return -1;
        }
        const double w2 = __dsub_rn(1.0, p);                                    // dem_processing.py:1082
        if (MODE == 3) {
            const bool x1 = k1 && (r1 < own_lo || r1 >= own_hi), x2 = k2 && (r2 < own_lo || r2 >= own_hi);
            if (x1 || x2) {
                if (x1) push_peer(r1, __dmul_rn(ai, p), __dmul_rn(ti, p), ti != 0.0);
                if (x2) push_peer(r2, __dmul_rn(ai, w2), __dmul_rn(ti, w2), ti != 0.0);
                k1 = k1 && !x1; k2 = k2 && !x2;
                if (!(k1 || k2)) return -1;
            }
        }
        // Ordering: a receiver's area must contain this contribution before the decrement that may
        // publish it.  The adds and the decrement of one receiver go to the SAME 32-byte sector and
        // are issued in program order by one thread, so they travel the same SM -> L2-slice path
        // and are performed in issue order; no fence, no waiting for the adds (a MEMBAR.GPU here
        // cost 20-160 % of the sweep).  That is how the hardware behaves (scripts/stress.py: 1200
        // sweeps + sharded runs, every result compared), not a PTX memory-model guarantee, so
        // `strict` (PYDEM_B200_SWEEP_STRICT=1) makes the decrement's operand depend on the adds'
        // return values -- a comparison ptxas cannot fold -- for cross-checking: one more round
        // trip per cell.
        int one = 1;
        int o1 = 0, o2 = 0;
        if (!strict) {
            // straight-line: four adds without a result, then both decrements, then the tests
            const bool tt = MODE != 1 && ti != 0.0;                                         // cyutils.pyx:163-164
            red_add_f64_if(k1, &cell[r1].area, __dmul_rn(ai, p));                           // cyutils.pyx:161
            red_add_f64_if(k2, &cell[r2].area, __dmul_rn(ai, w2));
            red_add_f64_if(k1 && tt, &cell[r1].taint, __dmul_rn(ti, p));
            red_add_f64_if(k2 && tt, &cell[r2].taint, __dmul_rn(ti, w2));
            o1 = atom_add_s32_if(k1, &cell[r1].indeg, -1);
            o2 = atom_add_s32_if(k2, &cell[r2].indeg, -1);
        } else {
            // cross-check: every decrement waits for the adds' return values (a comparison ptxas
            // cannot fold); kept apart from the normal path on purpose -- sharing the adds made the
            // compiler touch their return values in the normal path as well
            double a1 = 0.0, a2 = 0.0, u1 = 0.0, u2 = 0.0;
            if (k1) a1 = atomicAdd(&cell[r1].area, __dmul_rn(ai, p));
            if (k2) a2 = atomicAdd(&cell[r2].area, __dmul_rn(ai, w2));
            if (MODE != 1 && ti != 0.0) {
                if (k1) u1 = atomicAdd(&cell[r1].taint, __dmul_rn(ti, p));
                if (k2) u2 = atomicAdd(&cell[r2].taint, __dmul_rn(ti, w2));
            }
            const int NEVER = 0x7ff4dead;   // high word of a signalling NaN no sum produces
            const int one = 1 + ((__double2hiint(a1) == NEVER) | (__double2hiint(a2) == NEVER) | (__double2hiint(u1) == NEVER) |
                                 (__double2hiint(u2) == NEVER));
            if (k1) o1 = atomicSub(&cell[r1].indeg, one);
            if (k2) o2 = atomicSub(&cell[r2].indeg, one);
        }
        const bool rdy1 = k1 && o1 == 1, rdy2 = k2 && o2 == 1;
        if (rdy1 && rdy2) {
            // follow the larger share, hand the other receiver to an idle lane
            if (p >= 0.5) { nxt = r1; defer = r2; } else { nxt = r2; defer = r1; }
        } else if (rdy1) nxt = r1;
        else if (rdy2) nxt = r2;
        return nxt;
    }

    // Express chain (single lane, the critical path of the sweep): follow the flow path from `cur`
    // until it ends (cur = -1) or forks, i.e. both receivers became ready at once (cur = the
    // receiver with the larger share, other = the second one).  The caller gives `other` to an idle
    // lane of the same warp so that both branches advance together.  Returns the cells drained.
    //
    // The chain pays dependent memory round trips only: load the cell's record, then add + decrement
    // on the receivers.  scripts/ubench/lat.cu on B200: L2-hit ld.cg 209 ns, returning atomic
    // 260-310 ns, +130-230 ns when the sector comes from DRAM (river cells were last touched by
    // their tributaries, long before); measured here 1.0 us per cell on the 4096^2 conditioned DEM.
    // Ideas to shorten it that were built, measured at 4096^2 and dropped (all bit-compatible):
    // (1) a look-ahead walker on one-byte link hints prefetching the receivers' records into L2 --
    //     its own dependent loads miss L2 and cost more than the DRAM latency they save
    //     (1.03 -> 1.41 us per cell);
    // (2) taking the final area from the add's return value when the receiver's in-degree is
    //     already 1 -- needs that in-degree first, still two round trips;
    // (3) reading both receivers' records with one 256-bit load each right behind the decrement and
    //     carrying the next cell's record in registers (one round trip per cell).  In isolation
    //     this halves the latency (scripts/ubench/chain.cu, cold synthetic river: 932 -> 485 ns per
    //     cell); inside the sweep the chain still took 1.08 us per cell (the ~100 instructions of a
    //     step, executed by one lane with nothing to overlap them, weigh as much as a round trip)
    //     and the 4096^2 conditioned sweep was slower, 5.4-5.9 against 5.0-5.2 ms (builds
    //     alternated on one GPU): its 16 extra registers cost the bulk phase a resident block;
    // (4) the same for the lanes of a team (bounded chains): slower (6.5 -> 8.3 ms); draining a whole
    //     front of up to four ready cells per iteration from one lane (local-memory front): 12+ ms;
    // (5) keeping the receivers' link bytes in the record and prefetching the receivers' receivers
    //     one cell ahead without a dependent load: no measurable change.
    // What is left is the number of dependent L2 operations per level; the next step is to keep a
    // flow path's cells in shared memory ("what comes next").
    __device__ __forceinline__ unsigned long long chain(int32_t &cur, int32_t &other, const wl::Queue &q) const
    {
        unsigned long long n = 0;
        int32_t i = cur;
        while (i >= 0 && other < 0) {
            n++;
            i = process(i, q, other);
        }
        cur = i;
        return n;
    }

    // ---- Chain bursts: the express chain of a warp whose other lanes are idle, advanced by the WHOLE warp.
    //
    // The critical path of a conditioned DEM is a river of single-receiver cells (on the 4096^2 benchmark DEM 94 % of
    // the 3896 cells of the longest flow path have one receiver, mostly with weight exactly 1) whose tributaries
    // finished long ago: every cell on it waits for nothing but its predecessor.  Followed cell by cell it costs two
    // dependent L2 operations + ~100 single-lane instructions per cell (~1.07 us); but along such a run the sweep is a
    // linear recurrence  A[j] = S[j] + w[j] * A[j-1]  (S[j]: what the cell's record has accumulated = its own area +
    // its finished tributaries, cyutils.pyx:161), which a warp evaluates for up to 31 cells at once:
    //   1. the holder lane follows the static link bytes from its ready cell: R1, R2, ... (one L1/L2-hit byte load per
    //      hop -- the link plane is read-only during the sweep) while every next cell has a single receiver;
    //   2. lane t loads the record of R[t+1] with ONE 256-bit load.  A record with in-degree 1 is waiting for its
    //      chain predecessor only (the predecessor is undone and counts, so nobody else is pending, and every donor
    //      that has counted off did its adds before -- the ordering the sweep relies on anyway, see drain()); the
    //      leading lanes with in-degree 1 form the run;
    //   3. a warp scan composes the affine maps x -> S + w x over the run (area and taint), the lanes store the final
    //      records (in-degree 0 = drained), and the holder continues from the last cell of the run with its record in
    //      registers.  Nobody reads a record of the run again (all its donors have pushed; only the epilogue does).
    // A run that fails at once costs one hop and one load and falls back to the ordinary step; the speculation length
    // adapts (4..31).  Sums re-associate like the fp64 atomics of the bulk phase do (uca within the parity bar).
    // Measured without effect and dropped: following a copy of the link bytes laid out in 8 x 16-cell tiles of 128 B
    // (so that a hop mostly stays in one L1 line), and copying the 32 x 32 window of link bytes around the cell into
    // shared memory with all lanes before the chase (2.82-2.90 vs 2.72 ms).  A cycle-accounting build (clock64 per
    // phase, holder lanes summed) gave per burst: own record + chase ~5400 cycles, records ~1900, scan + stores ~800,
    // and ~2600 per ordinary step at the end of a run.
    // PYDEM_B200_SWEEP_BURST=0 switches it off; the strict cross-check mode never uses it.
    struct Rec32 { double area, taint, prop; int32_t indeg; uint8_t link; };
    static __device__ __forceinline__ Rec32 ld_rec(const Cell *c)
    {
        unsigned long long a, b, pq, d;
        asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(pq), "=l"(d) : "l"(c) : "memory");
        Rec32 r;
        r.area = __longlong_as_double((long long)a); r.taint = __longlong_as_double((long long)b);
        r.prop = __longlong_as_double((long long)pq);
        r.indeg = (int32_t)(d & 0xffffffffULL); r.link = (uint8_t)(d >> 32);
        return r;
    }
    // the only receiver of a cell with link byte lk at index c, or -1 (no / two receivers, pit drain, outside the owned rows)
    __device__ __forceinline__ int32_t single_receiver(int32_t c, uint8_t lk) const
    {
        if (lk & (LK_PIT | LK_NOSEC)) return -1;
        const bool k1 = lk & LK_KEEP1, k2 = lk & LK_KEEP2;
        if (k1 == k2) return -1;
        const int sec = lk & LK_SEC_MASK;
        const int32_t r = c + (k1 ? wl::off_e1(sec, C) : wl::off_e2(sec, C));
        if (MODE == 3 && (r < own_lo || r >= own_hi)) return -1;
        return r;
    }

    // offset from a cell with link byte lk to its only receiver, 0 if it has none or two (table s_off of k_worklist:
    // the holder's chase is one dependent instruction chain, and the table turns ~25 ALU instructions per hop into one
    // shared-memory load)
    __device__ __forceinline__ int32_t chase_offset(uint8_t lk) const
    {
        if (lk & (LK_PIT | LK_NOSEC)) return 0;
        const bool k1 = lk & LK_KEEP1, k2 = lk & LK_KEEP2;
        if (k1 == k2) return 0;
        const int sec = lk & LK_SEC_MASK;
        return k1 ? wl::off_e1(sec, C) : wl::off_e2(sec, C);
    }

    // all 32 lanes call this; `holder` is the one lane with work (cur >= 0).  sh: 32 ints of shared memory of this warp.
    // Returns the cells drained (on the holder lane; 0 elsewhere).
    __device__ __forceinline__ unsigned long long chain_warp(int32_t &cur, int32_t &other, const wl::Queue &q, int holder, int32_t *sh, int &L, int &score, const int32_t *s_off) const
    {
        const unsigned full = 0xffffffffu;
        const int lane = threadIdx.x & 31;
        const bool H = lane == holder;
        if ((MODE != 0 && MODE != 3) || strict || !burst) {
            return H ? chain(cur, other, q) : 0ULL;
        }
        unsigned long long n = 0;
        int32_t i = H ? cur : -1;
        bool have = false;                 // holder: the record of i is in the registers below
        double ai = 0.0, ti = 0.0, p = 0.0;
        uint8_t lk = 0;
        for (;;) {
            const int go = __shfl_sync(full, (H && i >= 0 && other < 0) ? 1 : 0, holder);
            if (!go) break;
            int32_t r1 = -1;
            if (H) {
                if (!have) {
                    const double2 at = __ldcg(reinterpret_cast<const double2 *>(&cell[i].area));
                    const longlong2 pm = __ldcg(reinterpret_cast<const longlong2 *>(&cell[i].prop));
                    ai = at.x; ti = at.y; p = __longlong_as_double(pm.x); lk = (uint8_t)((unsigned long long)pm.y >> 32);
                    have = true;
                }
                r1 = single_receiver(i, lk);
            }
            int run = 0;
            const unsigned long long x_t0 = q.dbg ? wl::globaltimer_ns() : 0;
            // score (per warp, persists across calls, -64..64): how productive this warp's bursts have been.  On terrain
            // without long single-receiver runs (a raw fractal: runs of ~1 cell) the bursts would only cost time: with a
            // negative score only every 16th step tries one (L <= 0 counts the ordinary steps in between)
            if (L <= 0) { L = L < -1 ? L + 1 : 4; r1 = -1; }
            if (__shfl_sync(full, r1, holder) >= 0) {
                // 1. follow the link bytes (holder): sh[t] = R[t+1]  (pulling the link bytes around the cell into L1 with
                //    all lanes first made no difference: measured, dropped)
                int nch = 0;
                if (H) {
                    int32_t c = r1;
                    while (nch < L) {
                        sh[nch++] = c;
#ifndef PDM_NO_PREFETCH
                        asm volatile("prefetch.global.L2 [%0];" :: "l"(cell + c));      // the lanes read these records next
#endif
                        const int32_t o = s_off[__ldg(link + c) & 0x7f];
                        if (o == 0) break;
                        c += o;
                        if (MODE == 3 && (c < own_lo || c >= own_hi)) break;
                    }
                }
                __syncwarp(full);
                if (q.dbg && H) atomicAdd(&q.ctr[CT_X_CHASE_NS], wl::globaltimer_ns() - x_t0);
                nch = __shfl_sync(full, nch, holder);
                // 2. the records of the chased cells, one lane each
                const bool mine = lane < nch;
                const int32_t c = mine ? sh[lane] : 0;
                Rec32 rec;
                rec.area = 0.0; rec.taint = 0.0; rec.prop = 0.0; rec.indeg = 0; rec.link = 0;
                if (mine) rec = ld_rec(&cell[c]);
                const unsigned vm = __ballot_sync(full, mine && rec.indeg == 1);
                run = __ffs(~vm) - 1;                       // leading lanes waiting for their predecessor only
                if (run < 0) run = 32;
                // (the speculation length follows the runs: grow after a full hit, shrink after a miss)
                if (run == nch && nch == L) { L = L * 2 < PDM_BURST_MAX ? L * 2 : PDM_BURST_MAX; }
                else if (2 * run < nch) { L = L / 2 > PDM_BURST_MIN ? L / 2 : PDM_BURST_MIN; }
                score += run - 2;
                score = score > 64 ? 64 : (score < -64 ? -64 : score);
                if (score < 0 && run < 2) L = -16;
                if (run > 0) {
                    // 3. weight of the edge into my cell: its predecessor's proportion (the holder's cell for lane 0)
                    const double hp = __shfl_sync(full, p, holder);
                    const int hl = __shfl_sync(full, (int)lk, holder);
                    double pp = __shfl_up_sync(full, rec.prop, 1);
                    int pl = __shfl_up_sync(full, (int)rec.link, 1);
                    if (lane == 0) { pp = hp; pl = hl; }
                    double m = (pl & LK_KEEP1) ? pp : __dsub_rn(1.0, pp);       // dem_processing.py:1082
                    double a = rec.area, t = rec.taint;
                    // inclusive scan of the affine maps x -> a + m x (area) / t + m x (taint) over the lanes [0, run)
                    for (int d = 1; d < 32; d <<= 1) {
                        const double m2 = __shfl_up_sync(full, m, d), a2 = __shfl_up_sync(full, a, d), t2 = __shfl_up_sync(full, t, d);
                        if (lane >= d) {
                            a = __dadd_rn(a, __dmul_rn(m, a2));
                            t = __dadd_rn(t, __dmul_rn(m, t2));
                            m = __dmul_rn(m, m2);
                        }
                    }
                    const double a0 = __shfl_sync(full, ai, holder), t0 = __shfl_sync(full, ti, holder);
                    const double A = __dadd_rn(a, __dmul_rn(m, a0));                                  // cyutils.pyx:161
                    const double T = (t0 != 0.0) ? __dadd_rn(t, __dmul_rn(m, t0)) : t;                // cyutils.pyx:163-164
                    // the last cell of the run becomes the holder's current cell (record in registers; it is the holder
                    // who stores it: whoever gets that cell later got it from the holder)
                    const int last = run - 1;
                    const double la = __shfl_sync(full, A, last), lt = __shfl_sync(full, T, last), lp = __shfl_sync(full, rec.prop, last);
                    const int ll = __shfl_sync(full, (int)rec.link, last);
                    const int32_t lc = __shfl_sync(full, c, last);
                    if (lane < last) {
                        // drained: final sums, in-degree 0
                        asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" :: "l"(&cell[c].area), "d"(A), "d"(T) : "memory");
                        asm volatile("st.global.cg.s32 [%0], %1;" :: "l"(&cell[c].indeg), "r"(0) : "memory");
                    }
                    if (H) {
                        asm volatile("st.global.cg.v2.f64 [%0], {%1, %2};" :: "l"(&cell[lc].area), "d"(la), "d"(lt) : "memory");
                        asm volatile("st.global.cg.s32 [%0], %1;" :: "l"(&cell[lc].indeg), "r"(0) : "memory");
                        i = lc; ai = la; ti = lt; p = lp; lk = (uint8_t)ll;
                        n += (unsigned long long)run;
                    }
                }
                __syncwarp(full);
                if (q.dbg) {
                    // diagnostics straight into the counters (no registers held across the loop)
                    const int fi = __shfl_sync(full, rec.indeg, run < 31 ? run : 31);      // in-degree of the record that ended the run
                    if (H) {
                        atomicAdd(&q.ctr[CT_X_BURSTS], 1ULL); atomicAdd(&q.ctr[CT_X_BURST_CELLS], (unsigned long long)run);
                        if (run == 0) atomicAdd(&q.ctr[CT_X_BURST_FAILS], 1ULL);
                        atomicAdd(&q.ctr[CT_X_BURST_NS], wl::globaltimer_ns() - x_t0);
                        atomicAdd(&q.ctr[76], (unsigned long long)nch);
                        if (run == nch) atomicAdd(&q.ctr[CT_DBG_DEALT], 1ULL);
                        else if (fi == 2) atomicAdd(&q.ctr[77], 1ULL);
                        else if (fi >= 3) atomicAdd(&q.ctr[78], 1ULL);
                        else atomicAdd(&q.ctr[79], 1ULL);
                    }
                }
            }
            if (run == 0 && H) {
                // the ordinary step (record in registers)
                n++;
                // (reading the receivers' records behind their decrements here, to start the next step without a load,
                //  was measured again in this path: 2.41-2.43 vs 2.34-2.35 ms -- dropped, as in round 1)
                i = drain(i, ai, MODE != 1 ? ti : 0.0, p, lk, q, other);
                have = false;
                if (other >= 0 && score > 0) {
                    // fork on a river (this warp's bursts have been productive): both receivers became ready at once.  The side branch of a D-infinity flow path usually ends
                    // one cell later in a cell of the main path, so it is drained right here (a few ordinary steps) and
                    // the main path goes on in bursts -- instead of leaving for the team loop of the warp.  What is left
                    // of a longer side branch goes to the queue.
                    int32_t sb = other, o2 = -1;
                    other = -1;
                    for (int k = 0; k < PDM_FORK_STEPS && sb >= 0 && o2 < 0; k++) { n++; sb = process(sb, q, o2); }
                    if (sb >= 0) q.push(sb);
                    if (o2 >= 0) q.push(o2);
                }
            }
        }
        if (H) cur = i;
        return n;
    }
};
