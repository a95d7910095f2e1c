// chain.cu -- which formulation of the sweep's express chain has the shortest per-cell latency?
// A synthetic river (one cell per row, wandering +-1 column) through 512 MB of 32-byte records,
// cold in L2, drained by ONE thread with the sweep's memory operations:
//   V0  today's step: load own record -> add (no return) + returning decrement on both receivers
//   V1  snapshot-before: 256-bit loads of both receivers issued BEFORE the adds/decrements of the
//       same level (same sector, performed first); when the snapshot shows in-degree 1 the final
//       area is snapshot + own share, no reload: one round trip per cell
//   V2  read-behind: adds, decrements, then loads of the receivers behind them (one round trip)
//   V3  preload-ahead: the receivers' records of the NEXT cell are loaded one level ahead
//       (different sectors than this level's atomics); one round trip per cell, DRAM hidden
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain chain.cu && ./chain
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

struct __align__(32) Rec { double area, taint; int off1, off2, indeg, pad; };
struct Snap { double area, taint; int off1, off2, indeg, pad; };

__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ Snap ld256(const Rec *p)
{
    unsigned long long a, b, c, d;
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p) : "memory");
    Snap s;
    s.area = __longlong_as_double(a); s.taint = __longlong_as_double(b);
    s.off1 = (int)(c & 0xffffffffu); s.off2 = (int)(c >> 32); s.indeg = (int)(d & 0xffffffffu); s.pad = (int)(d >> 32);
    return s;
}

const int C = 4096;
__host__ __device__ inline unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// river: row k -> column col(k); receivers of the path cell in row k: the path cell of row k+1 (off1) and its
// right-hand neighbour (off2, in-degree 2: never becomes ready)
__global__ void k_init(Rec *rec, int rows, int col0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    rec[i].area = 1.0; rec[i].taint = 0.0; rec[i].off1 = 0; rec[i].off2 = 0; rec[i].indeg = 2; rec[i].pad = 0;
}
__global__ void k_path(Rec *rec, int rows, int col0)
{
    int col = col0;
    for (int k = 0; k + 1 < rows; k++) {
        const int d = (int)(hash(k) % 3) - 1;
        int ncol = col + d; if (ncol < 8) ncol = 8; if (ncol > C - 8) ncol = C - 8;
        Rec &r = rec[(size_t)k * C + col];
        r.off1 = C + (ncol - col); r.off2 = r.off1 + 1; r.indeg = 1;
        col = ncol;
    }
    rec[(size_t)(rows - 1) * C + col].indeg = 1;
}
__global__ void k_evict(double *buf, size_t n) { for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = 0.0; }

__global__ void k_chain(Rec *rec, int start, int steps, int variant, unsigned long long *out)
{
    if (threadIdx.x) return;
    int c = start;
    const double p = 0.6, w2 = 0.4;
    unsigned long long t0 = gt();
    int n = 0;
    if (variant == 0) {
        while (n < steps) {
            const double2 at = __ldcg(reinterpret_cast<const double2 *>(&rec[c].area));
            const int4 m = __ldcg(reinterpret_cast<const int4 *>(&rec[c].off1));
            if (m.x == 0) break;
            const int r1 = c + m.x, r2 = c + m.y;
            atomicAdd(&rec[r1].area, at.x * p); atomicAdd(&rec[r2].area, at.x * w2);
            const int o1 = atomicSub(&rec[r1].indeg, 1), o2 = atomicSub(&rec[r2].indeg, 1);
            n++;
            if (o1 == 1) c = r1; else if (o2 == 1) c = r2; else break;
        }
    }
    out[0] = gt() - t0; out[1] = n; out[2] = (unsigned long long)c;
    out[3] = (unsigned long long)__double_as_longlong(rec[c].area);
}

__global__ void k_chain2(Rec *rec, int start, int steps, int variant, unsigned long long *out)
{
    if (threadIdx.x) return;
    int c = start, n = 0;
    const double p = 0.6, w2 = 0.4;
    const unsigned long long t0 = gt();
    Snap me = ld256(&rec[c]);
    if (variant == 1) {            // snapshot-before
        while (n < steps && me.off1 != 0) {
            const int r1 = c + me.off1, r2 = c + me.off2;
            const double c1 = me.area * p, c2 = me.area * w2;
            const Snap s1 = ld256(&rec[r1]), s2 = ld256(&rec[r2]);
            atomicAdd(&rec[r1].area, c1); atomicAdd(&rec[r2].area, c2);
            const int o1 = atomicSub(&rec[r1].indeg, 1), o2 = atomicSub(&rec[r2].indeg, 1);
            n++;
            if (o1 == 1) { c = r1; me = s1; if (s1.indeg == 1) me.area = s1.area + c1; else me = ld256(&rec[c]); }
            else if (o2 == 1) { c = r2; me = s2; if (s2.indeg == 1) me.area = s2.area + c2; else me = ld256(&rec[c]); }
            else break;
        }
    } else if (variant == 2) {     // read-behind
        while (n < steps && me.off1 != 0) {
            const int r1 = c + me.off1, r2 = c + me.off2;
            atomicAdd(&rec[r1].area, me.area * p); atomicAdd(&rec[r2].area, me.area * w2);
            const int o1 = atomicSub(&rec[r1].indeg, 1), o2 = atomicSub(&rec[r2].indeg, 1);
            const Snap s1 = ld256(&rec[r1]), s2 = ld256(&rec[r2]);
            n++;
            if (o1 == 1) { c = r1; me = s1; } else if (o2 == 1) { c = r2; me = s2; } else break;
        }
    } else if (variant == 3) {     // preload-ahead
        int r1 = c + me.off1, r2 = c + me.off2;
        Snap s1 = ld256(&rec[r1]), s2 = ld256(&rec[r2]);     // snapshots of this cell's receivers (before any of my atomics)
        while (n < steps && me.off1 != 0) {
            const double c1 = me.area * p, c2 = me.area * w2;
            atomicAdd(&rec[r1].area, c1); atomicAdd(&rec[r2].area, c2);
            const int o1 = atomicSub(&rec[r1].indeg, 1), o2 = atomicSub(&rec[r2].indeg, 1);
            // predicted next cell: the receiver waiting only for me; preload ITS receivers now
            const bool pred1 = s1.indeg == 1;
            const Snap &sp = pred1 ? s1 : s2;
            const int cp = pred1 ? r1 : r2;
            Snap q1, q2;
            const bool more = sp.off1 != 0;
            if (more) { q1 = ld256(&rec[cp + sp.off1]); q2 = ld256(&rec[cp + sp.off2]); }
            n++;
            int nx = -1; double cn = 0;
            if (o1 == 1) { nx = r1; cn = c1; } else if (o2 == 1) { nx = r2; cn = c2; }
            if (nx < 0) break;
            if (nx == cp && sp.indeg == 1) {
                c = nx; me = sp; me.area = sp.area + cn;
                if (!more) break;
                r1 = c + me.off1; r2 = c + me.off2; s1 = q1; s2 = q2;
            } else {
                c = nx; me = ld256(&rec[c]);
                if (me.off1 == 0) break;
                r1 = c + me.off1; r2 = c + me.off2; s1 = ld256(&rec[r1]); s2 = ld256(&rec[r2]);
            }
        }
    }
    out[0] = gt() - t0; out[1] = n; out[2] = (unsigned long long)c;
    out[3] = (unsigned long long)__double_as_longlong(me.area);
}

int main()
{
    const int rows = 4096;
    const size_t N = (size_t)rows * C;
    Rec *rec; double *ev; unsigned long long *out, h[4];
    cudaMalloc(&rec, N * sizeof(Rec)); cudaMalloc(&ev, (size_t)1 << 29); cudaMalloc(&out, 64);
    const char *names[] = {"V0 today (load own -> add + returning dec)", "V1 snapshot-before (ld256 + add + dec, one trip)",
                           "V2 read-behind (add + dec + ld256, one trip)", "V3 preload-ahead (next level's records one level early)"};
    for (int rep = 0; rep < 2; rep++)
        for (int v = 0; v < 4; v++) {
            k_init<<<(unsigned)((N + 255) / 256), 256>>>(rec, rows, 2000);
            k_path<<<1, 1>>>(rec, rows, 2000);
            k_evict<<<592, 256>>>(ev, ((size_t)1 << 29) / 8);
            cudaDeviceSynchronize();
            if (v == 0) k_chain<<<1, 32>>>(rec, 2000, rows, 0, out); else k_chain2<<<1, 32>>>(rec, 2000, rows, v, out);
            cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost);
            double area; memcpy(&area, &h[3], 8);
            printf("%-58s %7.1f ns/cell  (%llu cells, last cell %llu, area %.12g)\n", names[v], (double)h[0] / (double)h[1], h[1], h[2], area);
        }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
