// river.cu -- what does one cell of a river cost inside a shared-memory tile?  A synthetic river of
// `width` columns runs down a 32 x 32 tile; every cell has the donors N and NW/NE (two-wide band with
// cross links, like a D-infinity flow path) and the receivers S and SE/SW.  Variants:
//   0  warp-synchronous levels (ballot compaction, shared-memory atomics)  -- what round-2's first sweep did
//   1  one lane, level buffers, plain loads/stores (no atomics, no ballots, no warp sync)
//   2  like 1, software-pipelined by hand: receivers' counters and donors' values loaded together
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false river.cu -o river
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int TW = 32, TH = 32, HW = TW + 2, HN = HW * (TH + 2), TN = TW * TH;
struct S {
    double area[HN], taint[HN], prop[HN];
    uint32_t rdesc[TN], cnt[TN / 4];
    uint16_t list[2][TN];
    uint8_t dmask[TN], flags[TN], st[TN];
    double rowa[TH + 2];
    int n[3];
    unsigned notify;
};
__device__ __forceinline__ int nbr_off(int q)
{
    constexpr unsigned DR = 1u | (1u << 2) | (0u << 4) | (2u << 6) | (0u << 8) | (0u << 10) | (2u << 12) | (2u << 14);
    constexpr unsigned DC = 0u | (2u << 2) | (1u << 4) | (1u << 6) | (0u << 8) | (2u << 10) | (0u << 12) | (2u << 14);
    return ((int)((DR >> (2 * q)) & 3u) - 1) * HW + (int)((DC >> (2 * q)) & 3u) - 1;
}
__device__ void setup(S &s, int tid, int width)
{
    for (int i = tid; i < HN; i += 32) { s.area[i] = 1.0; s.taint[i] = 0.0; s.prop[i] = 0.75; }
    for (int o = tid; o < TN; o += 32) {
        const int y = o / TW, x = o % TW;
        const int xd = (x & 1) ? x - 1 : x + 1;        // diagonal partner column
        const bool in = x < width && xd < width;
        s.dmask[o] = y > 0 ? (uint8_t)(0x04 | (in ? ((x & 1) ? 0x10 : 0x20) : 0)) : 0;      // N, and NW (odd x) / NE (even x)
        s.flags[o] = 0; s.st[o] = 0;
        reinterpret_cast<uint8_t *>(s.cnt)[o] = y > 0 ? (in ? 2 : 1) : 0;
        const uint32_t r1 = y + 1 < TH ? (uint32_t)((y + 1) * TW + x) : 0xFFFFu;
        const uint32_t r2 = (y + 1 < TH && in) ? (uint32_t)((y + 1) * TW + xd) : 0xFFFFu;
        s.rdesc[o] = r1 | (r2 << 16);
    }
    for (int i = tid; i < TH + 2; i += 32) s.rowa[i] = 900.0;
    if (tid < width) s.list[0][tid] = (uint16_t)tid;
    if (tid == 0) { s.n[0] = width; s.n[1] = s.n[2] = 0; s.notify = 0; }
    __syncwarp();
}
__device__ __forceinline__ void pull(S &s, int o, double &ar, double &tt)
{
    const int y = o / TW, x = o - y * TW;
    const int k = (y + 1) * HW + x + 1;
    ar = s.rowa[y + 1];
    tt = (s.flags[o] & 1) ? 1.0 : 0.0;
    unsigned m = s.dmask[o];
    while (m) {
        const int q = __ffs(m) - 1;
        m &= m - 1;
        const int kk = k + nbr_off(q);
        const double p = s.prop[kk];
        const double wgt = q < 4 ? p : __dsub_rn(1.0, p);
        ar = __dadd_rn(ar, __dmul_rn(s.area[kk], wgt));
        tt = __dadd_rn(tt, __dmul_rn(s.taint[kk], wgt));
    }
    s.area[k] = ar; s.taint[k] = tt; s.st[o] = 1;
}
template <int VARIANT>
__global__ void k(int width, long long *out)
{
    extern __shared__ unsigned char raw[];
    S &s = *reinterpret_cast<S *>(raw);
    const int tid = threadIdx.x;
    if (tid >= 32) { __syncthreads(); return; }      // extra warps wait at a barrier like the tile sweep's idle warps
    setup(s, tid, width);
    const unsigned lt = (1u << tid) - 1u;
    int l = 0, m = width, cells = 0;
    const long long c0 = clock64();
    if (VARIANT == 0) {
        while (m > 0) {
            const bool act = tid < m;
            const int o = act ? (int)s.list[l & 1][tid] : 0;
            uint16_t *nxt = s.list[(l + 1) & 1];
            int r1 = -1, r2 = -1;
            if (act) {
                double ar, tt;
                pull(s, o, ar, tt);
                const uint32_t d = s.rdesc[o];
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const uint32_t c = e == 0 ? (d & 0xffffu) : (d >> 16);
                    if (c == 0xFFFFu) continue;
                    const int sh = (c & 3) * 8;
                    const uint32_t old = atomicSub(&s.cnt[c >> 2], 1u << sh);
                    if (((old >> sh) & 0xffu) == 1u) { if (e == 0) r1 = (int)c; else r2 = (int)c; }
                }
            }
            const unsigned m1 = __ballot_sync(0xffffffffu, r1 >= 0), m2 = __ballot_sync(0xffffffffu, r2 >= 0);
            const int n1 = __popc(m1);
            if (r1 >= 0) nxt[__popc(m1 & lt)] = (uint16_t)r1;
            if (r2 >= 0) nxt[n1 + __popc(m2 & lt)] = (uint16_t)r2;
            cells += m;
            m = n1 + __popc(m2);
            if (tid == 0) { s.n[(l + 1) % 3] = m; s.n[(l + 2) % 3] = 0; }
            __syncwarp();
            l++;
        }
    } else if (tid == 0) {
        uint8_t *cnt8 = reinterpret_cast<uint8_t *>(s.cnt);
        while (m > 0) {
            const uint16_t *cur = s.list[l & 1];
            uint16_t *nxt = s.list[(l + 1) & 1];
            int mm = 0;
            for (int i = 0; i < m; i++) {
                const int o = cur[i];
                if (VARIANT == 1) {
                    double ar, tt;
                    pull(s, o, ar, tt);
                    const uint32_t d = s.rdesc[o];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const uint32_t c = e == 0 ? (d & 0xffffu) : (d >> 16);
                        if (c == 0xFFFFu) continue;
                        const uint8_t b = cnt8[c] - 1;
                        cnt8[c] = b;
                        if (b == 0) nxt[mm++] = (uint16_t)c;
                    }
                } else {
                    // all loads of the cell up front: descriptor + donor mask, then counters and donor values
                    const uint32_t d = s.rdesc[o];
                    const unsigned dm = s.dmask[o];
                    const int y = o / TW, x = o - y * TW;
                    const int k0 = (y + 1) * HW + x + 1;
                    const uint32_t c1 = d & 0xffffu, c2 = d >> 16;
                    const bool h1 = c1 != 0xFFFFu, h2 = c2 != 0xFFFFu;
                    const uint8_t b1 = h1 ? cnt8[c1] : 2, b2 = h2 ? cnt8[c2] : 2;
                    double ar = s.rowa[y + 1], tt = (s.flags[o] & 1) ? 1.0 : 0.0;
                    unsigned q = dm;
                    while (q) {
                        const int qq = __ffs(q) - 1;
                        q &= q - 1;
                        const int kk = k0 + nbr_off(qq);
                        const double p = s.prop[kk];
                        const double wgt = qq < 4 ? p : __dsub_rn(1.0, p);
                        ar = __dadd_rn(ar, __dmul_rn(s.area[kk], wgt));
                        tt = __dadd_rn(tt, __dmul_rn(s.taint[kk], wgt));
                    }
                    s.area[k0] = ar; s.taint[k0] = tt; s.st[o] = 1;
                    if (h1) { cnt8[c1] = b1 - 1; if (b1 == 1) nxt[mm++] = (uint16_t)c1; }
                    if (h2) { const uint8_t bb = (h1 && c2 == c1) ? b1 - 1 : b2; cnt8[c2] = bb - 1; if (bb == 1) nxt[mm++] = (uint16_t)c2; }
                }
            }
            cells += m;
            m = mm;
            l++;
        }
    }
    const long long c1 = clock64();
    if (blockDim.x > 32) __syncthreads();
    if (tid == 0 && blockIdx.x == 0) { out[0] = c1 - c0; out[1] = l; out[2] = cells; out[3] = (long long)(s.area[(TH) * HW + 1] * 1000); }
}
template <int V> void run(const char *name, int width, int blocks = 1, int threads = 32)
{
    long long *d, h[4];
    cudaMalloc(&d, 32);
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S));
    for (int rep = 0; rep < 2; rep++) k<V><<<blocks, threads, sizeof(S)>>>(width, d);
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("%-46s width %2d: %3lld levels %4lld cells: %6.0f cycles per level, %5.0f per cell  chk %lld (%s)\n", name, width, h[1], h[2],
           (double)h[0] / (double)h[1], (double)h[0] / (double)h[2], h[3], cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}
int main()
{
    for (int b : {1, 148, 148 * 3, 148 * 5}) {
        char nm[64];
        snprintf(nm, sizeof nm, "warp-sync, %d blocks x 32 threads", b); run<0>(nm, 2, b, 32);
        snprintf(nm, sizeof nm, "warp-sync, %d blocks x 256 threads", b); run<0>(nm, 2, b, 256);
    }
    for (int w : {1, 2, 4, 8, 16}) {
        run<0>("warp-synchronous levels, ATOMS + ballots", w);
        run<1>("one lane, plain loads/stores", w);
        run<2>("one lane, loads hoisted", w);
    }
    return 0;
}
