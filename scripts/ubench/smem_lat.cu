// smem_lat.cu -- dependent-chain latencies inside one SM that bound a flow path followed in shared memory
// (tsweep.cu): LDS chase, returning ATOMS, fence.acq_rel.cta, fp64 add/mul, and the combination of a drain step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_lat smem_lat.cu && ./smem_lat
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define N 2048
template <int mode>
__global__ void k(int iters, long long *out, double *sink)
{
    __shared__ uint32_t nxt[N];
    __shared__ double val[N];
    __shared__ uint32_t cnt[N];
    for (int i = threadIdx.x; i < N; i += blockDim.x) { nxt[i] = (i * 1664525u + 1013904223u) & (N - 1); val[i] = 1.0 + i * 1e-9; cnt[i] = 1000000; }
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint32_t c = 0;
    double acc = 0.0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (mode == 0) { c = nxt[c]; }                                                         // LDS chase
        else if (mode == 1) { c = (atomicSub(&cnt[c], 1) & 0) + nxt[c]; }                       // returning ATOMS + LDS (parallel)
        else if (mode == 2) { const uint32_t o = atomicSub(&cnt[c], 1); c = nxt[(c + (o & 0)) & (N - 1)]; }   // ATOMS -> LDS dependent
        else if (mode == 3) { c = nxt[c]; asm volatile("fence.acq_rel.cta;" ::: "memory"); }    // LDS + fence
        else if (mode == 4) { acc = __dadd_rn(acc, val[c & 1]); c += (acc == -1.0); }           // DADD chain (+LDS)
        else if (mode == 5) { acc = __dadd_rn(acc, __dmul_rn(acc, 1e-30)); }                    // DMUL -> DADD chain
        else if (mode == 6) {                                                                   // a drain step: LDS meta, 2 donors, STS, fence, 2 ATOMS, next
            const uint32_t m = nxt[c];
            double a = val[m], b = val[(m + 7) & (N - 1)];
            acc = __dadd_rn(__dadd_rn(acc * 1e-30, __dmul_rn(a, 0.5)), __dmul_rn(b, 0.5));
            val[c] = acc;
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            const uint32_t o1 = atomicSub(&cnt[m], 1), o2 = atomicSub(&cnt[(m + 1) & (N - 1)], 1);
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            c = (o1 == 1u) ? ((m + 1) & (N - 1)) : (o2 == 1u ? ((m + 2) & (N - 1)) : m);
        } else if (mode == 7) {                                                                 // the same without fences
            const uint32_t m = nxt[c];
            double a = val[m], b = val[(m + 7) & (N - 1)];
            acc = __dadd_rn(__dadd_rn(acc * 1e-30, __dmul_rn(a, 0.5)), __dmul_rn(b, 0.5));
            val[c] = acc;
            const uint32_t o1 = atomicSub(&cnt[m], 1), o2 = atomicSub(&cnt[(m + 1) & (N - 1)], 1);
            c = (o1 == 1u) ? ((m + 1) & (N - 1)) : (o2 == 1u ? ((m + 2) & (N - 1)) : m);
        } else if (mode == 8) { __nanosleep(100); }
        else if (mode == 9) { __syncwarp(); c = nxt[c]; }
        else if (mode == 10) { asm volatile("" ::: "memory"); c += it; }
        else if (mode == 11) {
            const uint32_t m = nxt[c];
            double a = val[m], b = val[(m + 7) & (N - 1)];
            acc = __dadd_rn(__dadd_rn(acc * 1e-30, __dmul_rn(a, 0.5)), __dmul_rn(b, 0.5));
            val[c] = acc;
            sink[1 + (c & 1023) * 4] = acc;
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            const uint32_t o1 = atomicSub(&cnt[m], 1), o2 = atomicSub(&cnt[(m + 1) & (N - 1)], 1);
            asm volatile("fence.acq_rel.cta;" ::: "memory");
            c = (o1 == 1u) ? ((m + 1) & (N - 1)) : (o2 == 1u ? ((m + 2) & (N - 1)) : m);
        } else if (mode == 12) { c = nxt[c]; sink[1 + (c & 1023) * 4] = acc; asm volatile("fence.acq_rel.cta;" ::: "memory"); }
    }
    const long long t1 = clock64();
    out[mode] = t1 - t0;
    sink[0] = acc + c;
}

int main()
{
    long long *out; double *sink;
    cudaMallocManaged(&out, 64 * sizeof(long long)); cudaMalloc(&sink, 8 * 8192);
    const char *names[] = {"LDS chase", "ATOMS ret || LDS", "ATOMS ret -> LDS", "LDS + fence.acq_rel.cta", "LDS + DADD chain", "DMUL->DADD chain",
                           "drain step with fences", "drain step without fences", "nanosleep(100)", "syncwarp + LDS", "empty loop (volatile asm)", "drain step, fences + STG", "LDS + STG + fence", "drain step, fences, 31 lanes polling"};
    const int iters = 20000;
    for (int threads = 32; threads <= 128; threads *= 4)
    {
#define RUN(m) k<m><<<1, threads>>>(iters, out, sink); cudaDeviceSynchronize(); printf("threads %3d  %-28s %7.1f cycles / iteration\n", threads, names[m], (double)out[m] / iters);
        RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
    }
    return 0;
}
