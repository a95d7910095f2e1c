// worklist.cuh -- persistent, barrier-free work-list engine shared by the UCA sweep and the
// update-mode floods.
//
// A grid of resident warps (one launch, sized by the occupancy API so that every block is
// co-resident) runs until global quiescence:
//   * seeds are found by scanning a domain (all cells, or just the tile perimeter) for
//     Op::is_seed(cell),
//   * Op::process(cell, queue, defer) handles one ready cell and returns the cell this lane
//     continues with (chain following) or -1; one more ready cell can be returned in `defer`
//     (pushed to the global queue with one warp-aggregated fetch-and-add), further ones are
//     pushed directly (rare: pit drains),
//   * termination: every warp has finished its scan and scan-born chains (CT_PHASE1 ==
//     #warps) and every queued item has been completely processed (CT_QDONE == CT_QTAIL).
//     Reading QDONE, then PHASE1, then QTAIL makes the test race-free: both counters are
//     monotonic and QDONE <= QTAIL always, so equality observed in that order means there
//     was an instant with no active chain and an empty queue, after which nothing can be
//     produced.
// Every cell may enter the queue at most once per run, so a queue of N+1 slots suffices.
#pragma once
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

struct Queue {
    int32_t *slots;
    unsigned long long *ctr;
    long long cap;  // number of slots (a cell is queued at most once per run: N suffices)
    __device__ __forceinline__ void push(int32_t cell) const
    {
        const unsigned long long slot = atomicAdd(&ctr[CT_QTAIL], 1ULL);
        st_volatile_i32(slots + slot, cell);
    }
};

// scan domains: index -> cell
struct DomainAll {
    int64_t N;
    __device__ __forceinline__ int64_t size() const { return N; }
    __device__ __forceinline__ int32_t cell(int64_t t) const { return (int32_t)t; }
};
struct DomainRange {   // cells [first, first + n): the owned rows of a shard
    int64_t first, n;
    __device__ __forceinline__ int64_t size() const { return n; }
    __device__ __forceinline__ int32_t cell(int64_t t) const { return (int32_t)(first + t); }
};
struct DomainList {    // explicit seed list; its length lives in device memory
    const int32_t *cells;
    const unsigned long long *count;
    __device__ __forceinline__ int64_t size() const { return (int64_t)*count; }
    __device__ __forceinline__ int32_t cell(int64_t t) const { return cells[t]; }
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
};

template <class Op, class Domain>
__global__ void __launch_bounds__(256) k_worklist(Op op, Domain dom, Queue q)
{
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const unsigned long long nwarps = (unsigned long long)(nthreads >> 5);
    const int64_t dsize = dom.size();
    int64_t scan = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool scanning = scan < dsize;
    int32_t cur = -1;
    long long ticket = -1;  // queue slot this lane is entitled to (fetch-and-add ticket, never fails)
    int origin = 0;         // 1: chain started at a scanned seed, 2: at a queue item
    bool p1_reported = false;
    unsigned long long processed = 0;
    if (lane == 0) atomicMin(&q.ctr[CT_T_START], globaltimer_ns());

    for (;;) {
        // ---- acquire work: the next seed of this lane's scan stripe, or -- once the stripe is
        //      exhausted -- a queue ticket.  Tickets are handed out with fetch-and-add, so
        //      claiming never retries (a CAS-claimed queue serialises at one claim per L2 round
        //      trip); a ticket whose slot is still empty simply waits for its producer.
        if (cur < 0 && scanning) {
            for (int s = 0; s < 8 && scan < dsize; s++) {
                const int32_t c = dom.cell(scan);
                scan += nthreads;
                if (op.is_seed(c)) { cur = c; origin = 1; break; }
            }
            if (scan >= dsize) scanning = false;
        }
        const bool need_ticket = cur < 0 && !scanning && ticket < 0;
        const unsigned need_mask = __ballot_sync(full, need_ticket);
        if (need_mask) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&q.ctr[CT_QHEAD], (unsigned long long)__popc(need_mask));
            base = __shfl_sync(full, base, 0);
            if (need_ticket) ticket = (long long)(base + __popc(need_mask & ((1u << lane) - 1u)));
        }
        if (cur < 0 && ticket >= 0 && ticket < q.cap) {
            const int32_t v = ld_volatile_i32(q.slots + ticket);
            if (v >= 0) { cur = v; origin = 2; ticket = -1; }
        }
        // ---- scan accounting: a warp reports once its scan and scan-born chains have ended
        if (!p1_reported) {
            const unsigned busy1 = __ballot_sync(full, scanning || (cur >= 0 && origin == 1));
            if (busy1 == 0) {
                if (lane == 0) {
                    atomicAdd(&q.ctr[CT_PHASE1], 1ULL);
                    atomicMax(&q.ctr[CT_T_SCAN], globaltimer_ns());
                }
                p1_reported = true;
            }
        }
        // ---- nothing to do in this warp: terminate on global quiescence, else back off
        const unsigned work_mask = __ballot_sync(full, cur >= 0);
        if (work_mask == 0) {
            const unsigned still_scanning = __ballot_sync(full, scanning);
            if (still_scanning == 0) {
                int term = 0;
                if (lane == 0) {
                    const unsigned long long d = ld_volatile_u64(q.ctr + CT_QDONE);
                    const unsigned long long p1 = ld_volatile_u64(q.ctr + CT_PHASE1);
                    const unsigned long long t = ld_volatile_u64(q.ctr + CT_QTAIL);
                    term = (p1 == nwarps && d == t) ? 1 : 0;
                }
                term = __shfl_sync(full, term, 0);
                if (term) break;
                __nanosleep(100);
            }
            continue;
        }
        // ---- one step per working lane
        bool finished_q = false;
        int32_t defer = -1;  // second ready receiver: handed to the queue, warp-aggregated below
        if (cur >= 0) {
            processed++;
            const int32_t nxt = op.process(cur, q, defer);
            cur = nxt;
            if (cur < 0) { finished_q = (origin == 2); origin = 0; }
        }
        const unsigned pm = __ballot_sync(full, defer >= 0);
        if (pm) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&q.ctr[CT_QTAIL], (unsigned long long)__popc(pm));
            base = __shfl_sync(full, base, 0);
            if (defer >= 0) st_volatile_i32(q.slots + base + __popc(pm & ((1u << lane) - 1u)), defer);
        }
        const unsigned fq = __ballot_sync(full, finished_q);
        if (fq && lane == 0) atomicAdd(&q.ctr[CT_QDONE], (unsigned long long)__popc(fq));
    }
    for (int o = 16; o > 0; o >>= 1) processed += __shfl_down_sync(full, processed, o);
    if (lane == 0) {
        if (processed) atomicAdd(&q.ctr[CT_DRAINED], processed);
        atomicMax(&q.ctr[CT_T_END], globaltimer_ns());
    }
}

// facet -> flat-index offsets of the cardinal (e1) and diagonal (e2) receiver
// (facets table dem_processing.py:173-182)
__device__ __forceinline__ int32_t off_e1(int sec, int32_t C)
{
    const int r = (sec == 1 || sec == 2) ? -1 : ((sec == 5 || sec == 6) ? 1 : 0);
    const int c = (sec == 0 || sec == 7) ? 1 : ((sec == 3 || sec == 4) ? -1 : 0);
    return r * C + c;
}
__device__ __forceinline__ int32_t off_e2(int sec, int32_t C)
{
    const int r = (sec < 4) ? -1 : 1;
    const int c = (sec <= 1 || sec >= 6) ? 1 : -1;
    return r * C + c;
}
// opaque zero that depends on x (a true data dependency the compiler cannot fold)
__device__ __forceinline__ int dep_zero(double x)
{
    int z = __double2hiint(x);
    asm volatile("and.b32 %0, %0, 0;" : "+r"(z));
    return z;
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
    if (occ > 4) occ = 4;
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
    ctr[CT_QTAIL] = 0; ctr[CT_QHEAD] = 0; ctr[CT_QDONE] = 0; ctr[CT_PHASE1] = 0;
    if (!keep_drained) ctr[CT_DRAINED] = 0;
    ctr[CT_T_START] = ~0ULL; ctr[CT_T_SCAN] = 0; ctr[CT_T_END] = 0;
}

// host: reset the queue state before a run (slots to -1, the queue counters to 0)
inline int reset_queue(pdm_tile *t, int keep_drained = 0)
{
    if (!t->queue_ready) {
        PDM_CUDA(cudaMemsetAsync(t->queue, 0xFF, (size_t)(t->N + 1) * sizeof(int32_t), t->stream));
        t->queue_ready = true;
    } else {
        k_queue_clean<<<296, 256, 0, t->stream>>>(t->queue, t->d_counters, (long long)t->N);
        PDM_LAUNCHED();
    }
    k_queue_zero<<<1, 1, 0, t->stream>>>(t->d_counters, keep_drained);
    PDM_LAUNCHED();
    return PDM_OK;
}

}  // namespace wl
