// lat.cu -- dependent-chain latencies on B200 that bound the sweep's critical path:
// L2/DRAM loads, returning fp64/int atomics, with and without an L2 prefetch running ahead.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

struct __align__(32) Rec { double area, taint, prop; int indeg; int pad; };

__device__ __forceinline__ unsigned long long gt()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int dz(double x)
{
    int z = __double2hiint(x);
    asm volatile("and.b32 %0, %0, 0;" : "+r"(z));
    return z;
}
__device__ __forceinline__ int dzi(int z)
{
    asm volatile("and.b32 %0, %0, 0;" : "+r"(z));
    return z;
}
__device__ __forceinline__ uint32_t nxt(uint32_t c, uint32_t mask) { return (c * 1664525u + 1013904223u) & mask; }

// Every step's address comes out of the previous step's returned value (ptxas folds "x & 0", so
// the chain has to be a real pointer chase): rec[i].area = (double)next(i), rec[i].indeg = next(i).
// mode 0: ld.cg chain; 1: atomicAdd f64 (+0.0) chain; 2: atomicAdd s32 (+0) chain;
// 3: the sweep's step: ld record -> (fire-and-forget adds) -> returning sub on the receiver;
// 4: mode 1 with an L2 prefetch running `pf` steps ahead; 5: mode 3 with prefetch
__global__ void k_init(Rec *rec, uint32_t mask)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > mask) return;
    const uint32_t n = nxt(i, mask);
    rec[i].area = (double)n; rec[i].taint = 0.0; rec[i].prop = 0.5; rec[i].indeg = (int)n; rec[i].pad = 0;
}

__global__ void k_lat(Rec *rec, uint32_t mask, int steps, int mode, int pf, unsigned long long *out, uint32_t start)
{
    if (threadIdx.x != 0) return;
    uint32_t c = start & mask;
    uint32_t ahead = c;
    for (int k = 0; k < pf; k++) ahead = nxt(ahead, mask);
    const unsigned long long t0 = gt();
    for (int s = 0; s < steps; s++) {
        if (pf) { asm volatile("prefetch.global.L2 [%0];" ::"l"(&rec[ahead])); ahead = nxt(ahead, mask); }
        if (mode == 0) { c = (uint32_t)__ldcg(&rec[c].area); }
        else if (mode == 1 || mode == 4) { c = (uint32_t)atomicAdd(&rec[c].area, 0.0); }
        else if (mode == 2) { c = (uint32_t)atomicAdd(&rec[c].indeg, 0); }
        else {
            const double2 at = __ldcg(reinterpret_cast<const double2 *>(&rec[c].area));
            const uint32_t r = (uint32_t)at.x;                       // the "receiver"
            atomicAdd(&rec[r].taint, at.y);                          // fire and forget
            c = (uint32_t)atomicAdd(&rec[r].indeg, 0);               // returning; then on to the receiver's receiver
        }
    }
    const unsigned long long t1 = gt();
    out[0] = t1 - t0;
    out[1] = c;
}

// background traffic: other blocks hammer random records with atomics (like the bulk of the sweep)
__global__ void k_noise(Rec *rec, uint32_t mask, int iters)
{
    uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u & mask;
    for (int i = 0; i < iters; i++) { atomicAdd(&rec[c].taint, 1.0); c = nxt(c, mask); }
}

int main()
{
    const uint32_t NBIG = 1u << 24, NSMALL = 1u << 18;   // 512 MB / 8 MB of records
    Rec *rec; unsigned long long *out, h[2];
    cudaMalloc(&rec, (size_t)NBIG * sizeof(Rec));
    cudaMemset(rec, 0, (size_t)NBIG * sizeof(Rec));
    cudaMalloc(&out, 16);
    const char *names[] = {"ld.cg", "atomicAdd f64 (returning)", "atomicAdd s32 (returning)", "ld -> red + atom (sweep step)",
                           "atom f64 + prefetch 16 ahead", "sweep step + prefetch 16 ahead"};
    const int steps = 20000;
    for (int big = 0; big < 2; big++) {
        const uint32_t n = big ? NBIG : NSMALL, mask = n - 1;
        for (int noise = 0; noise < 2; noise++)
        for (int mode = 0; mode < 6; mode++) {
            const int pf = (mode >= 4) ? 16 : 0;
            k_init<<<(n + 255) / 256, 256>>>(rec, mask);
            if (big) k_init<<<(NBIG + 255) / 256, 256>>>(rec + 0, NBIG - 1);     // (re)write everything: the walk starts cold-ish
            else k_lat<<<1, 32>>>(rec, mask, steps, 0, 0, out, 12345u);            // warm L2 along the same walk
            cudaDeviceSynchronize();
            cudaStream_t s2; cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
            if (noise) k_noise<<<148 * 4, 256, 0, s2>>>(rec, NBIG - 1, 4000);
            k_lat<<<1, 32>>>(rec, mask, steps, mode >= 4 ? (mode == 4 ? 1 : 3) : mode, pf, out, 12345u);
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            cudaDeviceSynchronize();
            cudaStreamDestroy(s2);
            printf("%-6s %-10s %-34s %7.1f ns/step\n", big ? "512MB" : "8MB", noise ? "busy" : "quiet", names[mode], (double)h[0] / steps);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
