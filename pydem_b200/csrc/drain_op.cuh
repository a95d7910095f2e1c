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
// one 32-byte sweep record read with a single 256-bit load (one point in time for all fields)
struct Rec32 { double area, taint, prop; int32_t indeg; uint8_t link; };
__device__ __forceinline__ Rec32 ld_rec(const Cell *c)
{
    unsigned long long a, b, p, d;
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(p), "=l"(d) : "l"(c) : "memory");
    Rec32 r;
    r.area = __longlong_as_double((long long)a); r.taint = __longlong_as_double((long long)b);
    r.prop = __longlong_as_double((long long)p);
    r.indeg = (int32_t)(d & 0xffffffffULL); r.link = (uint8_t)(d >> 32);
    return r;
}

template <int MODE>
struct DrainOp {
    const uint8_t *link;   // SoA copy of the link bytes: only the seed scan reads it
    Cell *cell;            // 32-byte sweep records
    const uint8_t *st;
    int32_t C;
    const int32_t *pit_beg;
    const int32_t *pit_end;
    const int32_t *pit_dst;
    const double *pit_w;
    int32_t strict;        // hold the decrements back until the adds have returned (see process())
    int32_t spec;          // the tail reads the receivers' records behind the decrement (see step())

    __device__ __forceinline__ bool is_seed(int32_t c) const
    {
        return MODE == 0 ? (link[c] & LK_SOURCE) != 0 : (MODE == 2 ? true : (st[c] & ST_START) != 0);
    }
    __device__ __forceinline__ bool skip(int32_t r) const { return MODE == 1 && (st[r] & ST_START); }

    // What a lane carries from one step to the next: the record of the cell it is about to drain,
    // when it was read behind the decrement that made that cell ready (see step()).
    struct State {
        int32_t cell;     // the cell the fields below belong to (-1: nothing carried)
        double area, taint, prop;
        uint8_t link;
        __device__ State() : cell(-1), area(0.0), taint(0.0), prop(0.0), link(0) {}
    };

    // Drain one ready cell; returns the receiver this lane continues with (or -1).
    //
    // read_behind (the latency-bound tail of the sweep; off in the bandwidth-bound bulk phase): both
    // receivers' records are read with one 256-bit load each, issued right BEHIND their
    // decrements.  Adds, decrement and load of one receiver target the same 32-byte sector and are
    // performed in issue order (the ordering the sweep relies on anyway, see below), so when a
    // decrement turns out to have been the last one, the record read behind it already holds the
    // receiver's final area and taint, and the next step starts without loading it: ONE dependent
    // round trip per cell instead of two.  scripts/ubench/chain.cu (cold synthetic river, B200):
    // 932 -> 485 ns per cell ("snapshot before the decrement" 520 ns and "preload one level
    // ahead" 554 ns are the measured alternatives).  PYDEM_B200_SWEEP_SPEC=0 turns it off.
    // Tried and dropped: a look-ahead walker on one-byte link hints prefetching records into L2 (its
    // own dependent loads cost more than the DRAM latency they save) and prefetching the
    // receivers' receivers from link bytes kept in the record (no measurable change).
    __device__ __forceinline__ int32_t step(int32_t i, const wl::Queue &q, int32_t &defer, State &st, bool read_behind) const
    {
        double ai, ti, p;
        uint8_t lk;
        if (st.cell == i) {
            ai = st.area; ti = st.taint; p = st.prop; lk = st.link;
        } else {
            // the cell's whole sweep state is one 32-byte sector
            const Rec32 me = ld_rec(&cell[i]);
            ai = me.area; ti = me.taint; p = me.prop; lk = me.link;
        }
        st.cell = -1;
        if (MODE == 1) ti = 0.0;
        int32_t nxt = -1;
        if (lk & LK_PIT) {
            // long-range pit edges (_mk_connectivity_pits): rare, plain fence ordering
            const int64_t slot = __double_as_longlong(p);
            const int32_t e0 = pit_beg[slot], e1 = pit_end[slot];
            for (int32_t e = e0; e < e1; e++) {
                const int32_t r = pit_dst[e];
                if (skip(r)) continue;
                const double w = pit_w[e];
                atomicAdd(&cell[r].area, __dmul_rn(ai, w));
                if (MODE != 1 && ti != 0.0) atomicAdd(&cell[r].taint, __dmul_rn(ti, w));
            }
            __threadfence();
            for (int32_t e = e0; e < e1; e++) {
                const int32_t r = pit_dst[e];
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
        // Ordering: a receiver's area must contain this contribution before the decrement that may
        // publish it.  The adds and the decrement of one receiver go to the SAME 32-byte sector and
        // are issued in program order by one thread, so they travel the same SM -> L2-slice path
        // and are performed in issue order; no fence, no waiting for the adds (a MEMBAR.GPU here
        // cost 20-160 % of the sweep).  That is how the hardware behaves (scripts/stress.py: 1200
        // sweeps + sharded runs, every result compared), not a PTX memory-model guarantee, so
        // `strict` (PYDEM_B200_SWEEP_STRICT=1) makes the decrement's operand depend on the adds'
        // return values -- a comparison ptxas cannot fold -- for cross-checking: one more round
        // trip per cell.
        double a1 = 0.0, a2 = 0.0, u1 = 0.0, u2 = 0.0;
        if (k1) a1 = atomicAdd(&cell[r1].area, __dmul_rn(ai, p));               // cyutils.pyx:161
        if (k2) a2 = atomicAdd(&cell[r2].area, __dmul_rn(ai, w2));
        if (MODE != 1 && ti != 0.0) {                                           // cyutils.pyx:163-164
            if (k1) u1 = atomicAdd(&cell[r1].taint, __dmul_rn(ti, p));
            if (k2) u2 = atomicAdd(&cell[r2].taint, __dmul_rn(ti, w2));
        }
        int one = 1;
        if (strict) {
            const int NEVER = 0x7ff4dead;   // high word of a signalling NaN no sum produces
            one += (__double2hiint(a1) == NEVER) | (__double2hiint(a2) == NEVER) | (__double2hiint(u1) == NEVER) |
                   (__double2hiint(u2) == NEVER);
        }
        int o1 = 0, o2 = 0;
        if (k1) o1 = atomicSub(&cell[r1].indeg, one);
        if (k2) o2 = atomicSub(&cell[r2].indeg, one);
        Rec32 s1, s2;
        const bool rb = MODE != 1 && read_behind && spec && !strict;
        if (rb) {
            if (k1) s1 = ld_rec(&cell[r1]);     // behind the decrement of the same sector
            if (k2) s2 = ld_rec(&cell[r2]);
        }
        const bool rdy1 = k1 && o1 == 1, rdy2 = k2 && o2 == 1;
        if (rdy1 && rdy2) {
            // follow the larger share, hand the other receiver to an idle lane
            if (p >= 0.5) { nxt = r1; defer = r2; } else { nxt = r2; defer = r1; }
        } else if (rdy1) nxt = r1;
        else if (rdy2) nxt = r2;
        if (rb && nxt >= 0) {
            const Rec32 &s = (nxt == r1) ? s1 : s2;
            st.cell = nxt; st.area = s.area; st.taint = s.taint; st.prop = s.prop; st.link = s.link;
        }
        return nxt;
    }

    // bulk phase: one cell, nothing carried, no read-behind
    __device__ __forceinline__ int32_t process(int32_t i, const wl::Queue &q, int32_t &defer) const
    {
        State st;
        return step(i, q, defer, st, false);
    }

    // Express chain (single lane, the critical path of the sweep): follow the flow path from `cur`
    // until it ends (cur = -1) or forks, i.e. both receivers became ready at once (cur = the
    // receiver with the larger share, other = the second one; the caller gives `other` to an idle
    // lane of the same warp so that both branches advance together).  The next cell's record is
    // carried in registers (step(), read-behind).  Returns the cells drained.
    // Measured and dropped at 4096^2: running the lanes of a team through bounded chains of this
    // kind to give them the carried record too (6.5 -> 8.3 ms), and draining a whole front of up
    // to four ready cells per iteration from one lane (front kept in local memory: 12+ ms).
    __device__ __forceinline__ unsigned long long chain(int32_t &cur, int32_t &other, const wl::Queue &q) const
    {
        unsigned long long n = 0;
        int32_t i = cur;
        State st;
        while (i >= 0 && other < 0) {
            n++;
            i = step(i, q, other, st, true);
        }
        cur = i;
        return n;
    }
};
