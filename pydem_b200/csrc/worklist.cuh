// worklist.cuh -- persistent, barrier-free work-list engine shared by the UCA sweep and the
// update-mode floods.
//
// A grid of resident warps (one launch, sized by the occupancy API so that every block is
// co-resident) runs until global quiescence:
//   * seeds are found by scanning a domain (owned cells, a seed list, or just the tile
//     perimeter) for Op::is_seed(cell): chunks of 32 consecutive entries are handed to warps by a
//     global counter, tested with one coalesced load and dealt out to the warp's idle lanes,
//   * Op::process(cell, queue, defer) handles one ready cell and returns the cell this lane
//     continues with (chain following) or -1; one more ready cell can be returned in `defer`
//     (pushed to the global queue with one warp-aggregated fetch-and-add), further ones are
//     pushed directly (rare: pit drains),
//   * termination: the number of seeds S is known before the launch (the kernels that flag
//     the seeds count them), every chain -- whether it started at a seed or at a queue item --
//     bumps CT_QDONE when it has completely ended, and every queue push bumps CT_QTAIL.  Always
//     QDONE <= seeds dealt + QTAIL <= S + QTAIL, all monotonic; so reading QDONE first and QTAIL
//     second and finding QDONE == S + QTAIL proves that at the first read every seed had been
//     dealt and every started chain had ended: nothing can be produced any more.
// Every cell may enter the queue at most once per run, so a queue of N+1 slots suffices.
#pragma once
#include <stdlib.h>

#include "pdm_internal.cuh"

namespace wl {

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
__device__ __forceinline__ void st_volatile_i32(int32_t *p, int32_t v)
{
    asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// zero derived from x.  NOTE: ptxas folds "x & 0", so this does NOT hold the consumer back until x
// has arrived (checked in SASS); it only keeps the source order explicit.  Nothing relies on it:
// the termination test is confirmed with fenced read-modify-writes, and a chain's queue pushes
// are complete before its end is counted because the slot store consumed the push counter's
// return value (in-order issue).
__device__ __forceinline__ unsigned long long dep_zero_u64(unsigned long long x)
{
    asm volatile("and.b64 %0, %0, 0;" : "+l"(x));
    return x;
}

struct Queue {
    int32_t *slots;
    unsigned long long *ctr;
    long long cap;  // number of slots (a cell is queued at most once per run: N suffices)
    const unsigned long long *nseeds;  // device counter: exact number of seeds in the scan domain
    int32_t backoff_max;   // longest sleep of an idle warp between polls, ns (0: default)
    int32_t dbg;           // collect the CT_X_* diagnostics (PYDEM_B200_WL_DEBUG)
    // ---- one sweep across the row shards of several GPUs (Op::P2P; see the block comment at p2p_service)
    int32_t *inbox;                       // cells a peer GPU made ready, in arrival order (slots are -1 until written)
    long long inbox_cap;
    unsigned long long epoch;             // number of this multi-GPU sweep (CT_GTERM >= epoch: it is over everywhere)
    const unsigned long long *arrived;    // start barrier on rank 0: ranks whose state is ready, summed over all sweeps
    unsigned long long start_target;
    int32_t world;
    unsigned long long *all_ctr[PDM_MAX_WORLD];   // every rank's counters (own included), mapped peer memory
    __device__ __forceinline__ void push(int32_t cell) const
    {
        const unsigned long long slot = atomicAdd(&ctr[CT_QTAIL], 1ULL);
        st_volatile_i32(slots + slot, cell);
    }
};

// scan domains: index -> cell
#ifndef PDM_CHUNK_BATCH
#define PDM_CHUNK_BATCH 16     // chunks a warp reserves at a time = rows of a patch of DomainPatches
#endif
// (the seed scan walks a domain in chunks of 32 entries: nchunks() chunks, chunk_cell(chunk, lane) = the lane's cell or -1)
#define WL_LINEAR_CHUNKS                                                                                          \
    __device__ __forceinline__ long long nchunks() const { return (long long)((size() + 31) >> 5); }              \
    __device__ __forceinline__ int32_t chunk_cell(long long ch, int lane) const { const int64_t t = (ch << 5) + lane; return t < size() ? cell(t) : -1; }
struct DomainAll {
    int64_t N;
    __device__ __forceinline__ int64_t size() const { return N; }
    __device__ __forceinline__ int32_t cell(int64_t t) const { return (int32_t)t; }
    WL_LINEAR_CHUNKS
};
struct DomainRange {   // cells [first, first + n): the owned rows of a shard
    int64_t first, n;
    __device__ __forceinline__ int64_t size() const { return n; }
    __device__ __forceinline__ int32_t cell(int64_t t) const { return (int32_t)(first + t); }
    WL_LINEAR_CHUNKS
};
// the owned rows [row0, row0 + nrows) of a tile with C columns, walked in patches of 16 rows x 32 columns: a batch of 16 (PDM_CHUNK_BATCH)
// chunks (what a warp reserves at a time) is one patch, so the cells a warp drains back to back are neighbours in both
// directions and their receivers' records are still in L2 when the next row of the patch pushes into them
struct DomainPatches {
    int64_t row0, nrows, C;
    __device__ __forceinline__ long long tpr() const { return (long long)((C + 31) >> 5); }
    __device__ __forceinline__ long long nchunks() const { return ((nrows + PDM_CHUNK_BATCH - 1) / PDM_CHUNK_BATCH) * tpr() * PDM_CHUNK_BATCH; }
    __device__ __forceinline__ int32_t chunk_cell(long long ch, int lane) const
    {
        const long long patch = ch / PDM_CHUNK_BATCH, t = tpr();
        const long long row = (patch / t) * PDM_CHUNK_BATCH + (ch % PDM_CHUNK_BATCH), col = (patch % t) * 32 + lane;
        return (row < nrows && col < C) ? (int32_t)((row0 + row) * C + col) : -1;
    }
};
struct DomainList {    // explicit seed list; its length lives in device memory
    const int32_t *cells;
    const unsigned long long *count;
    __device__ __forceinline__ int64_t size() const { return (int64_t)*count; }
    __device__ __forceinline__ int32_t cell(int64_t t) const { return cells[t]; }
    WL_LINEAR_CHUNKS
};
struct DomainBorder {
    int64_t R, C;
    __device__ __forceinline__ int64_t size() const { return 2 * C + 2 * (R - 2); }
    __device__ __forceinline__ int32_t cell(int64_t t) const
    {
        if (t < C) return (int32_t)t;
        if (t < 2 * C) return (int32_t)((R - 1) * C + (t - C));
        const int64_t u = t - 2 * C;
        return (int32_t)((1 + (u >> 1)) * C + ((u & 1) ? C - 1 : 0));
    }
    WL_LINEAR_CHUNKS
};

// One sweep across the row shards of several GPUs (Op::P2P, drain_op.cuh MODE 3).  A cell whose receiver
// belongs to the neighbouring rank adds into that rank's record and counts its in-degree down with
// system-scope atomics over NVLink; the cell that brings a peer's counter to zero goes into the peer's
// IN-BOX (CT_INBOX_TAIL + slot store).  Warp 0 of block 0 of every rank is the SERVICE warp: it
//   * forwards in-box cells to the local queue (single consumer: no claim traffic, 32 cells per turn).
//     Accounting: the in-box item counts as produced at CT_INBOX_TAIL (by the peer) and as finished at
//     CT_QDONE once it has been re-pushed (CT_QTAIL), so that on every rank
//     QDONE <= seeds dealt + QTAIL + INBOX_TAIL, all monotonic;
//   * detects the end: when this rank looks quiescent it reads every rank's QDONE (first wave), then every
//     rank's QTAIL, INBOX_TAIL and seed count (second wave, issued after the first has returned), each with a
//     system-scope read-modify-write performed at the counter's home.  sum QDONE == sum (seeds + QTAIL +
//     INBOX_TAIL) proves that at a moment between the waves nothing was running or queued anywhere (a push
//     into a peer's in-box has returned before the pushing chain is counted as finished).  It then writes the
//     sweep number into every rank's CT_GTERM -- which is never reset, so a rank that already prepares the
//     next sweep cannot be confused by a late writer, and a rank whose own check reads a neighbour's freshly
//     reset counters still leaves through its flag.
// All other warps leave when their own CT_GTERM says so.
static __device__ __forceinline__ void p2p_service(const Queue &q, unsigned long long nseeds)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned long long head = 0, t_idle = 0;
    unsigned polls = 0;
    for (;;) {
        unsigned long long tail = 0;
        if (lane == 0) tail = ld_volatile_u64(q.ctr + CT_INBOX_TAIL);
        tail = __shfl_sync(full, tail, 0);
        if (tail > (unsigned long long)q.inbox_cap) { if (lane == 0) atomicExch(&q.ctr[CT_WATCHDOG], 2ULL); tail = head; }
        while (head < tail) {
            const int n = (int)(tail - head < 32ULL ? tail - head : 32ULL);
            int32_t v = -1;
            if (lane < n) {
                // the peer bumps the counter first and stores the slot right after: a short wait at most
                unsigned spins = 0;
                while ((v = ld_volatile_i32(q.inbox + head + lane)) < 0 && ++spins < (1u << 24)) __nanosleep(50);
            }
            const unsigned ok = __ballot_sync(full, v >= 0);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&q.ctr[CT_QTAIL], (unsigned long long)__popc(ok));
            base = __shfl_sync(full, base, 0);
            if (v >= 0) st_volatile_i32(q.slots + base + __popc(ok & ((1u << lane) - 1u)), v);
            __syncwarp(full);
            // finished only after the re-push has been counted (base has returned: in-order issue)
            if (lane == 0) atomicAdd(&q.ctr[CT_QDONE], (unsigned long long)n + dep_zero_u64(base));
            if (__popc(ok) != n && lane == 0) atomicExch(&q.ctr[CT_WATCHDOG], 3ULL);   // a slot never arrived (cannot happen)
            head += n;
            t_idle = 0;
        }
        int term = 0;
        if (lane == 0) {
            if (ld_volatile_u64(q.ctr + CT_GTERM) >= q.epoch) term = 1;
            else if (ld_volatile_u64(q.ctr + CT_WATCHDOG) != 0) term = 1;
        }
        if (__shfl_sync(full, term, 0)) break;
        // cheap local filter (plain loads, QDONE first), then the two waves over all ranks
        int quiet = 0;
        if (lane == 0) {
            const unsigned long long d = ld_volatile_u64(q.ctr + CT_QDONE);
            const unsigned long long t = ld_volatile_u64(q.ctr + CT_QTAIL + dep_zero_u64(d));
            const unsigned long long it = ld_volatile_u64(q.ctr + CT_INBOX_TAIL + dep_zero_u64(t));
            quiet = (d == nseeds + t + it && it == head) ? 1 : 0;
        }
        quiet = __shfl_sync(full, quiet, 0);
        if (quiet) {
            unsigned long long d = 0, t = 0;
            unsigned long long *c = lane < q.world ? q.all_ctr[lane] : nullptr;
            if (c) d = atomicAdd_system(c + CT_QDONE, 0ULL);
            for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(full, d, o);
            __threadfence_system();
            if (c) {
                // second wave: its addresses depend on the first wave's sum (and the shuffles above were a
                // convergence point), so it is issued after every QDONE has come back
                const unsigned long long z = dep_zero_u64(d);
                t = atomicAdd_system(c + CT_QTAIL + z, 0ULL) + atomicAdd_system(c + CT_INBOX_TAIL + z, 0ULL) +
                    atomicAdd_system(c + CT_SOURCES + z, 0ULL);
            }
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(full, t, o);
            if (d == t) {
                if (c) atomicMax_system(c + CT_GTERM, q.epoch);
                break;
            }
            __nanosleep(1000);
        } else {
            __nanosleep(100);
        }
        // watchdog (never fires in a correct run): nothing arrived and nobody finished for 4 s
        if ((++polls & 1023u) == 0 && lane == 0) {
            const unsigned long long now = globaltimer_ns();
            if (t_idle == 0) t_idle = now;
            else if (now - t_idle > 8000000000ULL) atomicExch(&q.ctr[CT_WATCHDOG], 1ULL);
        }
    }
}

template <class Op, class Domain>
__global__ void __launch_bounds__(256, 4) k_worklist(Op op, Domain dom, Queue q)
{
    __shared__ int32_t s_chain[8][32];   // per warp: the cells of a chain burst (Op::chain_warp)
    __shared__ int32_t s_off[128];       // link byte -> offset of the cell's only receiver (0: none or two); Op::chase_offset
    if (threadIdx.x < 128) s_off[threadIdx.x] = op.chase_offset((uint8_t)threadIdx.x);
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const long long nchunks = dom.nchunks();
    // seed scan state (warp-uniform): the scan domain is cut into chunks of 32 consecutive
    // entries; a warp reserves a batch of CHUNK_BATCH chunks with one fetch-and-add (a per-chunk
    // counter was the hottest address of the kernel), tests a chunk with one coalesced load,
    // keeps the mask of seeds not yet given to a lane, and moves on when it is empty
    const int CHUNK_BATCH = PDM_CHUNK_BATCH;
    // Critical path.  All lanes of a warp advance in lock step, so a chain that shares its warp
    // with 31 other busy lanes pays the slowest lane's memory latency three times per cell
    // (measured ~7 us per cell against ~1 us alone).  Long flow paths (rivers of a conditioned
    // DEM: thousands of cells) therefore run in EXPRESS mode: one chain per warp, lane 0 only.
    //   * warp 0 of every block is express from the start (it never scans for seeds);
    //   * every other warp becomes express when its seed scan is exhausted;
    //   * a scanning warp hands a chain longer than HANDOFF cells to the queue, a team one longer than TEAM_HANDOFF.
// (thresholds A/B-ed on one B200 at 4096^2, conditioned / raw sweep in ms: team 4, scan 16: 2.70 / 1.12; 8: 2.58 / 1.00; 16: 2.50 / 0.99;
//  16, 32: 2.44 / 0.98; 32, 64: 2.35 / 0.99; 64, 64: 2.37 / 1.00; 32, 128: 2.35 / 0.97; no team hand-off at all: 3.1-4.1, unstable)
#ifndef PDM_HANDOFF
#define PDM_HANDOFF 64
#endif
#ifndef PDM_TEAM_HANDOFF
#define PDM_TEAM_HANDOFF 32
#endif
    const int HANDOFF = PDM_HANDOFF, TEAM_HANDOFF = PDM_TEAM_HANDOFF;
    const bool express_only = (threadIdx.x >> 5) == 0 && gridDim.x * (blockDim.x >> 5) > 8;
    bool scanning = nchunks > 0 && !express_only;
    int32_t ch_next = 0, ch_end = 0;     // current batch [ch_next, ch_end) (chunks of 32 cells: < 2^26 per tile)
    unsigned pend_mask = 0;  // lanes whose pend_cell is an unassigned seed
    int32_t pend_cell = -1;
    int32_t cur = -1;
    int32_t stash = -1, stash2 = -1, stash3 = -1;   // ready receivers kept for after the current chain (see below)
    int chain_len = 0;
    int burst_len = 16;      // speculation length of the chain bursts (warp-uniform, adapts to the runs; Op::chain_warp)
    int burst_score = 8;     // how productive this warp's bursts have been (warp-uniform)
    int32_t ticket = -1;     // queue slot this lane is entitled to (fetch-and-add ticket, never fails; the queue has <= 2^31 slots)
    bool active = false;     // this lane owns an unfinished chain (current cell and/or stash)
    bool scan_stamped = false;
    const unsigned long long nseeds = *q.nseeds;
    unsigned processed = 0;              // cells this lane drained (a tile has < 2^31 cells)
    unsigned idle_polls = 0, iters = 0;
    unsigned long long push_dep = 0;
    unsigned done_acc = 0;   // chains this warp has finished but not yet added to CT_QDONE (warp-uniform)
    unsigned long long idle_since = 0;   // watchdog: a warp that sees no progress for WATCHDOG_NS raises CT_WATCHDOG
    const unsigned long long WATCHDOG_NS = 4000000000ULL;
    unsigned backoff = 200;
    // (A/B at 4096^2 with the chain bursts, conditioned / raw sweep in ms: 200 ns: 2.33 / 0.97; 400: 2.31 / 0.97; 1600: 2.28 / 0.95;
    //  3200: 2.26 / 0.95; 6400: 2.32 / 0.97; 12800: 2.49 / 1.01)
    const unsigned backoff_max = q.backoff_max > 0 ? (unsigned)q.backoff_max : 2000u;
    if constexpr (Op::P2P) {
        // start barrier: nobody touches a peer's records or in-box before every rank has reset its own
        // (each rank arrives on rank 0's counter after its set-up kernels, pdm_launch_sweep_p2p)
        if (threadIdx.x == 0) {
            while (ld_volatile_u64(q.arrived) < q.start_target) __nanosleep(200);
            __threadfence_system();
        }
        __syncthreads();
    }
    if (lane == 0) atomicMin(&q.ctr[CT_T_START], globaltimer_ns());
    if constexpr (Op::P2P) {
        if (blockIdx.x == 0 && (threadIdx.x >> 5) == 0) {
            p2p_service(q, nseeds);
            if (lane == 0) atomicMax(&q.ctr[CT_T_END], globaltimer_ns());
            return;
        }
    }

    for (;;) {
        // ---- a finished chain first continues with the lane's own stashed cell
        if (cur < 0 && stash >= 0) { cur = stash; stash = stash2; stash2 = stash3; stash3 = -1; chain_len = 0; }
        // ---- acquire work, seeds first
        unsigned idle_mask = __ballot_sync(full, cur < 0);
        while (scanning && idle_mask) {
            if (pend_mask == 0) {
                if (ch_next >= ch_end) {
                    long long b = 0;
                    if (lane == 0) b = (long long)atomicAdd(&q.ctr[CT_CHUNK], (unsigned long long)CHUNK_BATCH);
                    b = __shfl_sync(full, b, 0);
                    if (b >= nchunks) { scanning = false; break; }
                    ch_next = (int32_t)b;
                    ch_end = (int32_t)(b + CHUNK_BATCH < nchunks ? b + CHUNK_BATCH : nchunks);
                }
                pend_cell = dom.chunk_cell((long long)ch_next, lane);
                ch_next++;
                pend_mask = __ballot_sync(full, pend_cell >= 0 && op.is_seed(pend_cell));
                if (pend_mask == 0) continue;
            }
            // k-th idle lane takes the k-th pending seed
            const int k = __popc(idle_mask & lt_mask);
            const int npend = __popc(pend_mask);
            const bool take = (cur < 0) && k < npend;
            const int src = take ? (int)__fns(pend_mask, 0, k + 1) : 0;
            const int32_t c = __shfl_sync(full, pend_cell, src);
            if (take) { cur = c; active = true; chain_len = 0; }
            int ntake = __popc(idle_mask);
            if (ntake > npend) ntake = npend;
#ifdef PDM_WL_DEBUG   // diagnostics (hot address: costs ~40 % of a seed-heavy sweep)
            if (lane == 0) atomicAdd(&q.ctr[CT_DBG_DEALT], (unsigned long long)ntake);
#endif
            for (int i = 0; i < ntake; i++) pend_mask &= pend_mask - 1;   // drop the seeds just handed out
            idle_mask = __ballot_sync(full, cur < 0);
        }
        // ---- no more seeds for this warp: idle lanes take a queue ticket.  Tickets are handed out
        //      with fetch-and-add, so claiming never retries (a CAS-claimed queue serialises at one
        //      claim per L2 round trip); a ticket whose slot is still empty waits for its producer.
        // express: one queue item per warp at a time (its forks stay in the warp, see below)
        const bool warp_idle = __ballot_sync(full, cur >= 0 || stash >= 0) == 0;
        const bool need_ticket = warp_idle && !scanning && ticket < 0 && lane == 0;
        const unsigned need_mask = __ballot_sync(full, need_ticket);
        if (need_mask) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&q.ctr[CT_QHEAD], (unsigned long long)__popc(need_mask));
            base = __shfl_sync(full, base, 0);
            if (need_ticket) { const unsigned long long tk = base + __popc(need_mask & lt_mask); ticket = tk < (unsigned long long)q.cap ? (int32_t)tk : 0x7fffffff; }
        }
        if (cur < 0 && ticket >= 0 && ticket != 0x7fffffff) {
            const int32_t v = ld_volatile_i32(q.slots + ticket);
            if (v >= 0) {
                cur = v; active = true; ticket = -1; chain_len = 0;
#ifdef PDM_WL_DEBUG
                atomicAdd(&q.ctr[CT_DBG_TAKEN], 1ULL);
#endif
            }
        }
        // ---- statistics only: when did the last warp run out of seeds
        if (!scanning && !scan_stamped) {
            if (lane == 0) atomicMax(&q.ctr[CT_T_SCAN], globaltimer_ns());
            if (q.dbg) atomicAdd(&q.ctr[CT_X_SCAN_CELLS], (unsigned long long)processed);
            scan_stamped = true;
        }
        // ---- nothing to do in this warp: terminate on global quiescence, else back off
        const unsigned work_mask = __ballot_sync(full, cur >= 0);
        if (work_mask == 0) {
            // finished chains are counted lazily: one add when the warp runs dry instead of one per
            // step (every seed-born chain counts, and a per-step add to the one QDONE address cost
            // ~40 % of a source-rich sweep).  QDONE only has to be complete at quiescence, and a warp
            // that still holds uncounted chains is by construction not idle yet.
            if (done_acc) {
                if (lane == 0) atomicAdd(&q.ctr[CT_QDONE], (unsigned long long)done_acc + dep_zero_u64(push_dep));
                done_acc = 0;
            }
            if (!scanning) {
                // idle warps mostly watch their own ticket slots (distinct addresses); the shared
                // counters are read only every 4th round so that polling does not slow the
                // producers' atomics on the same lines
                int term = 0;
                if (Op::P2P) {
                    // a sweep across GPUs ends when the service warp (p2p_service) has seen -- or been told --
                    // that every rank is quiescent
                    if ((++idle_polls & 3u) == 0) {
                        if (lane == 0) term = ld_volatile_u64(q.ctr + CT_GTERM) >= q.epoch ? 1 : 0;
                        term = __shfl_sync(full, term, 0);
                    }
                } else if ((++idle_polls & 3u) == 0) {
                    if (lane == 0) {
                        // the proof needs QDONE to be read before QTAIL.  The cheap filter reads both
                        // with plain volatile loads (the second address depends on the first value);
                        // the two lines live in different L2 slices -- on different dies -- and a
                        // load may be served from a copy that lags the slice the atomics are performed
                        // in (seen once per ~2000 sweeps: a stale QTAIL made 37 warps leave early and
                        // orphaned a ticket).  So an apparent quiescence is CONFIRMED with two
                        // read-modify-writes, which are performed at the counters themselves, each
                        // issued only after the previous one returned.
                        const unsigned long long d = ld_volatile_u64(q.ctr + CT_QDONE);
                        const unsigned long long t = ld_volatile_u64(q.ctr + CT_QTAIL + dep_zero_u64(d));
                        if (d == nseeds + t) {
                            __threadfence();
                            const unsigned long long d2 = atomicAdd(&q.ctr[CT_QDONE], 0ULL);
                            __threadfence();
                            const unsigned long long t2 = atomicAdd(&q.ctr[CT_QTAIL + dep_zero_u64(d2)], 0ULL);
                            term = (d2 == nseeds + t2) ? 1 : 0;
                        }
                    }
                    term = __shfl_sync(full, term, 0);
                }
                if (term) { if (lane == 0) atomicAdd(&q.ctr[CT_DBG_EXITS], 1ULL); break; }
                // watchdog (never fires in a correct run): instead of spinning forever on a lost
                // item, give up loudly -- the host turns CT_WATCHDOG into an error
                if ((idle_polls & 63u) == 0) {
                    int quit = 0;
                    if (lane == 0) {
                        const unsigned long long now = globaltimer_ns();
                        if (idle_since == 0) idle_since = now;
                        else if (now - idle_since > WATCHDOG_NS) atomicExch(&q.ctr[CT_WATCHDOG], 1ULL);
                        quit = ld_volatile_u64(q.ctr + CT_WATCHDOG) != 0;
                    }
                    if (__shfl_sync(full, quit, 0)) break;
                }
                // idle warps back off: thousands of them polling the same few lines slow the L2
                // slices the working chains' atomics go through
                __nanosleep(backoff);
                if (backoff < backoff_max) backoff <<= 1;
                if (q.dbg && lane == 0) atomicAdd(&q.ctr[CT_X_POLLS], 1ULL);
            }
            continue;
        }
        idle_since = 0;
        backoff = 200;
        if ((++iters & 4095u) == 0) {
            int quit = 0;
            if (lane == 0) quit = ld_volatile_u64(q.ctr + CT_WATCHDOG) != 0;
            if (__shfl_sync(full, quit, 0)) break;
        }
        // ---- one step per working lane
        bool finished_q = false;
        int32_t defer = -1;  // second ready receiver: dealt to an idle lane of this warp or handed to the queue
        bool handoff = false;   // defer is this lane's own long chain on its way to an express warp
        if (!scanning && (work_mask & (work_mask - 1)) == 0) {
            // express fast path: one lane has work and the warp has no seeds left to deal -> it
            // follows the chain in a tight single-lane loop (no warp collectives between cells:
            // the critical path pays memory round trips only) until the chain ends or forks
            // (all lanes go in: the idle ones help the chain along single-receiver runs, drain_op.cuh chain_warp)
            const int holder = __ffs(work_mask) - 1;
            const unsigned long long t0 = q.dbg ? globaltimer_ns() : 0;
            const unsigned long long nc = op.chain_warp(cur, defer, q, holder, s_chain[threadIdx.x >> 5], burst_len, burst_score, s_off);
            if (lane == holder) {
                processed += (unsigned)nc;
                if (q.dbg) { atomicAdd(&q.ctr[CT_X_CHAIN_NS], globaltimer_ns() - t0); atomicAdd(&q.ctr[CT_X_CHAIN_CELLS], nc); atomicAdd(&q.ctr[CT_X_CHAIN_CALLS], 1ULL); }
                if (cur < 0 && defer >= 0) { cur = defer; defer = -1; }
                if (cur < 0 && stash < 0) { finished_q = active; active = false; }
            }
            __syncwarp(full);
        } else if (cur >= 0) {
            processed++;
            chain_len++;
            const unsigned long long t0 = (q.dbg && !scanning) ? globaltimer_ns() : 0;
            const int32_t nxt = op.process(cur, q, defer);
            if (q.dbg && !scanning) { atomicAdd(&q.ctr[CT_X_TEAM_NS], globaltimer_ns() - t0); atomicAdd(&q.ctr[CT_X_TEAM_LANES], 1ULL); }
            cur = nxt;
            // a second ready receiver waits in the lane's stash (no queue traffic) unless the
            // stash is taken; a long chain hands its stash to the queue so that it cannot sit on
            // the critical path behind it
            if (defer >= 0 && cur < 0) { cur = defer; defer = -1; }
            else if (defer >= 0 && scanning && stash3 < 0) {
                // bulk phase: a second ready receiver waits in the lane's stash (no queue traffic)
                if (stash < 0) stash = defer; else if (stash2 < 0) stash2 = defer; else stash3 = defer;
                defer = -1;
            }
            // bulk phase: a long chain moves to an express warp; its stash follows one cell per step.  The same in a
            // team (several chains in lock step, after the scan): a chain that keeps going is a river, and an idle warp
            // that takes it from the queue advances it in bursts (Op::chain_warp) instead of ~3 us per cell here
            if (defer < 0 && chain_len > (scanning ? HANDOFF : TEAM_HANDOFF)) {
                if (cur >= 0) { defer = cur; cur = -1; chain_len = 0; handoff = true; }
            }
            if (cur < 0 && stash < 0) { finished_q = active; active = false; }
        }
        // ---- express team: a flow path that forks (both receivers became ready at once) keeps both
        //      branches in this warp -- the k-th forking lane gives its second receiver to the k-th
        //      idle lane, which drains it in lock step.  Through the global queue the branch would
        //      wait for an idle warp to notice its ticket (~1 us, more than draining two cells),
        //      and the branches of a D-infinity flow path re-join a few cells later, so that wait
        //      sat on the critical path of nearly every level.  A dealt cell is not a queue item:
        //      it belongs to the giving chain's accounting, which is only reported when the whole
        //      warp has run dry.
        if (!scanning) {
            const unsigned dm = __ballot_sync(full, defer >= 0 && !handoff);      // (a handed-off chain goes to the queue)
            if (dm) {
                const bool idle = cur < 0 && stash < 0 && ticket < 0;
                const unsigned im = __ballot_sync(full, idle);
                const int nd = __popc(dm), ni = __popc(im);
                const int k = __popc(im & lt_mask);
                const bool take = idle && k < nd;
                const int src = take ? (int)__fns(dm, 0, k + 1) : 0;
                const int32_t c = __shfl_sync(full, defer, src);
                if (defer >= 0 && !handoff && __popc(dm & lt_mask) < ni) defer = -1;   // given away
                if (take) { cur = c; chain_len = 0; }
            }
        }
        const unsigned pm = __ballot_sync(full, defer >= 0);
        unsigned long long base = 0;
        if (pm) {
            if (lane == 0) base = atomicAdd(&q.ctr[CT_QTAIL], (unsigned long long)__popc(pm));
            base = __shfl_sync(full, base, 0);
            if (defer >= 0) st_volatile_i32(q.slots + base + __popc(pm & lt_mask), defer);
        }
        const unsigned fq = __ballot_sync(full, finished_q);
        // counted later (see done_acc).  A chain's pushes are performed before its end is counted: the
        // push counter add is returning and its value addressed the slot store above, so it was
        // back before anything after it issued (a fence here cost 20-160 % of the sweep)
        done_acc += __popc(fq);
        if (pm) push_dep = base;
    }
    unsigned long long processed_w = processed;
    for (int o = 16; o > 0; o >>= 1) processed_w += __shfl_down_sync(full, processed_w, o);
    if (lane == 0) {
        if (processed_w) atomicAdd(&q.ctr[CT_DRAINED], processed_w);
        atomicMax(&q.ctr[CT_T_END], globaltimer_ns());
    }
}

// facet -> flat-index offsets of the cardinal (e1) and diagonal (e2) receiver
// (facets table dem_processing.py:173-182)
__device__ __forceinline__ int32_t off_e1(int sec, int32_t C)
{
    // rows {0,-1,-1,0,0,1,1,0}, columns {1,0,0,-1,-1,0,0,1} as two packed 2-bit tables (value + 1)
    const int r = ((0x6941 >> (2 * sec)) & 3) - 1;
    const int c = ((0x9416 >> (2 * sec)) & 3) - 1;
    return r * C + c;
}
__device__ __forceinline__ int32_t off_e2(int sec, int32_t C)
{
    const int r = ((sec >> 1) & 2) - 1;             // -1 for facets 0..3, +1 for 4..7
    const int c = 1 - (((sec + 2) >> 1) & 2);       // +1 for facets 0,1,6,7, -1 for 2..5
    return r * C + c;
}
// host: tunables / diagnostics of a run (environment, read once)
inline Queue tuned(Queue q)
{
    static int backoff = -1, dbg = 0;
    if (backoff < 0) {
        const char *e = getenv("PYDEM_B200_WL_BACKOFF");
        backoff = e ? atoi(e) : 0;
        const char *d = getenv("PYDEM_B200_WL_DEBUG");
        dbg = d ? atoi(d) : 0;
    }
    q.backoff_max = backoff;
    q.dbg = dbg;
    return q;
}
inline int max_blocks_per_sm()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("PYDEM_B200_WL_OCC"); v = e ? atoi(e) : 8; if (v < 1) v = 1; }
    return v;
}

// host: persistent grid size (all blocks co-resident) for a given instantiation
template <class K>
int grid_for(K kernel, int *blocks_out)
{
    int dev = 0, sms = 0, occ = 0;
    PDM_CUDA(cudaGetDevice(&dev));
    PDM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PDM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, 0));
    if (occ < 1) { pdm_set_error("work-list kernel does not fit on an SM"); return PDM_ERR_CUDA; }
    if (occ > max_blocks_per_sm()) occ = max_blocks_per_sm();
    *blocks_out = sms * occ;
    return PDM_OK;
}

// wipe the slots the previous run used (they are the only ones that differ from -1)
static __global__ void __launch_bounds__(256) k_queue_clean(int32_t *slots, const unsigned long long *ctr, long long cap)
{
    long long n = (long long)ctr[CT_QTAIL];
    if (n > cap) n = cap;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        slots[i] = -1;
}
static __global__ void k_queue_zero(unsigned long long *ctr, int keep_drained)
{
    ctr[CT_QTAIL] = 0; ctr[CT_QHEAD] = 0; ctr[CT_QDONE] = 0; ctr[CT_PHASE1] = 0; ctr[CT_CHUNK] = 0;   // (CT_GTERM is never reset)
    if (!keep_drained) ctr[CT_DRAINED] = 0;
    for (int k = CT_X_FIRST; k <= CT_X_LAST; k++) ctr[k] = 0;
    for (int k = 76; k < 80; k++) ctr[k] = 0;
    ctr[CT_T_START] = ~0ULL; ctr[CT_T_SCAN] = 0; ctr[CT_T_END] = 0; ctr[CT_DBG_DEALT] = 0; ctr[CT_DBG_TAKEN] = 0; ctr[CT_DBG_EXITS] = 0;
}

// host: after a run's counters were read back, turn a fired watchdog into an error
inline int check_watchdog(pdm_tile *t)
{
    const unsigned long long *h = t->h_counters;
    if (h[CT_WATCHDOG]) {
        pdm_set_error("work-list watchdog fired: no progress for 4 s (QTAIL=%llu QHEAD=%llu QDONE=%llu PHASE1=%llu CHUNK=%llu "
                      "DRAINED=%llu); results are incomplete", h[CT_QTAIL], h[CT_QHEAD], h[CT_QDONE], h[CT_PHASE1], h[CT_CHUNK],
                      h[CT_DRAINED]);
        return PDM_ERR_STATE;
    }
    return PDM_OK;
}

// host: reset the queue state before a run (slots to -1, the queue counters to 0)
inline int reset_queue(pdm_tile *t, int keep_drained = 0)
{
    if (!t->queue_ready || (t->queue_dirty_ctr && t->queue_dirty_ctr != t->d_counters)) {
        PDM_CUDA(cudaMemsetAsync(t->queue, 0xFF, (size_t)(t->N + 1) * sizeof(int32_t), t->stream));
        t->queue_ready = true; t->queue_dirty_ctr = t->d_counters;
    } else {
        k_queue_clean<<<296, 256, 0, t->stream>>>(t->queue, t->d_counters, (long long)t->N);
        PDM_LAUNCHED();
    }
    k_queue_zero<<<1, 1, 0, t->stream>>>(t->d_counters, keep_drained);
    PDM_LAUNCHED();
    return PDM_OK;
}

}  // namespace wl
