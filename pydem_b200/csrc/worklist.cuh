// worklist.cuh -- persistent, barrier-free work-list engine shared by the UCA sweep and the
// update-mode floods.
//
// A grid of resident warps (one launch, sized by the occupancy API so that every block is
// co-resident) runs until global quiescence:
//   * seeds are found by scanning a domain (all cells, or just the tile perimeter) for
//     Op::is_seed(cell),
//   * Op::process(cell, push) handles one ready cell and returns the cell this lane
//     continues with (chain following) or -1; additional ready cells are handed to idle
//     lanes through a global queue (push),
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

struct Queue {
    int32_t *slots;
    unsigned long long *ctr;
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
    int origin = 0;  // 1: chain started at a scanned seed, 2: at a queue item
    bool p1_reported = false;
    unsigned long long processed = 0;
    unsigned iter = 0;

    for (;;) {
        iter++;
        // ---- acquire work: queue first (items on the critical path), then the next seed
        bool want = cur < 0;
        const unsigned want_mask = __ballot_sync(full, want);
        const unsigned scan_mask = __ballot_sync(full, scanning);
        if (want_mask) {
            if (scan_mask == 0 || (iter & 3u) == 0) {
                int32_t base = 0, got = 0;
                if (lane == 0) {
                    const unsigned long long h = ld_volatile_u64(q.ctr + CT_QHEAD);
                    const unsigned long long t = ld_volatile_u64(q.ctr + CT_QTAIL);
                    if (t > h) {
                        unsigned long long k = (unsigned long long)__popc(want_mask);
                        if (k > t - h) k = t - h;
                        if (atomicCAS(&q.ctr[CT_QHEAD], h, h + k) == h) { base = (int32_t)h; got = (int32_t)k; }
                    }
                }
                base = __shfl_sync(full, base, 0);
                got = __shfl_sync(full, got, 0);
                if (got) {
                    const int rank = __popc(want_mask & ((1u << lane) - 1u));
                    if (want && rank < got) {
                        int32_t v;
                        do { v = ld_volatile_i32(q.slots + base + rank); } while (v < 0);  // slot reserved, store in flight
                        cur = v; origin = 2; want = false;
                    }
                }
            }
            if (want && scanning) {
                for (int s = 0; s < 8 && scan < dsize; s++) {
                    const int32_t c = dom.cell(scan);
                    scan += nthreads;
                    if (op.is_seed(c)) { cur = c; origin = 1; break; }
                }
                if (scan >= dsize) scanning = false;
            }
        }
        // ---- scan accounting: a warp reports once its scan and scan-born chains have ended
        if (!p1_reported) {
            const unsigned busy1 = __ballot_sync(full, scanning || (cur >= 0 && origin == 1));
            if (busy1 == 0) {
                if (lane == 0) atomicAdd(&q.ctr[CT_PHASE1], 1ULL);
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
                __nanosleep(200);
            }
            continue;
        }
        // ---- one step per working lane
        bool finished_q = false;
        if (cur >= 0) {
            processed++;
            const int32_t nxt = op.process(cur, q);
            cur = nxt;
            if (cur < 0) { finished_q = (origin == 2); origin = 0; }
        }
        const unsigned fq = __ballot_sync(full, finished_q);
        if (fq && lane == 0) atomicAdd(&q.ctr[CT_QDONE], (unsigned long long)__popc(fq));
    }
    for (int o = 16; o > 0; o >>= 1) processed += __shfl_down_sync(full, processed, o);
    if (lane == 0 && processed) atomicAdd(&q.ctr[CT_DRAINED], processed);
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

// host: reset the queue state before a run (slots to -1, the queue counters to 0)
inline int reset_queue(pdm_tile *t)
{
    PDM_CUDA(cudaMemsetAsync(t->queue, 0xFF, (size_t)(t->N + 1) * sizeof(int32_t), t->stream));
    PDM_CUDA(cudaMemsetAsync(t->d_counters, 0, 5 * sizeof(unsigned long long), t->stream));  // QTAIL..DRAINED
    return PDM_OK;
}

}  // namespace wl
