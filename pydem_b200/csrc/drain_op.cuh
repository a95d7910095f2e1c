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
        const double ai = at.x;
        const double ti = MODE != 1 ? at.y : 0.0;
        const double p = __longlong_as_double(pm.x);
        const uint8_t lk = (uint8_t)((unsigned long long)pm.y >> 32);
        int32_t nxt = -1;
        if (lk & LK_PIT) {
            // long-range pit edges (_mk_connectivity_pits): rare, plain fence ordering
            const int64_t slot = pm.x;
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
};
