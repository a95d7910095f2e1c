// flats.cu -- K2: flat mask with one-pixel extension.
//
// Reference behaviour: _find_flats_edges (dem_processing.py:657-680) with
// scipy.ndimage.label (8-connectivity) and utils.get_adjacent_index (utils.py:270-311),
// then the stamping direction[flats] = mag[flats] = -1 (611-613).
//
// The reference walks the labelled regions in label order and overwrites
// flat[J] = (elev[J] == elev[first cell of region]) for every 8-neighbour J of a region
// cell, so for each J the adjacent region with the HIGHEST label decides.  scipy labels in
// raster order of a region's first cell, so "highest label" == "largest minimum index".
// Here: union-find over flat0 = (mag == -1) with the minimum cell index as root
// (k_ccl_merge / k_ccl_flatten), then one gather pass (k_flats_extend).
#include "pdm_internal.cuh"

namespace {

__device__ __forceinline__ int32_t uf_find(int32_t *label, int32_t x)
{
    int32_t p = label[x];
    while (p != x) {
        int32_t gp = label[p];
        if (gp != p) label[x] = gp;  // path halving (benign race: only ever moves toward the root)
        x = p;
        p = gp;
    }
    return x;
}

__device__ __forceinline__ void uf_union(int32_t *label, int32_t a, int32_t b)
{
    for (;;) {
        a = uf_find(label, a);
        b = uf_find(label, b);
        if (a == b) return;
        if (a < b) { int32_t s = a; a = b; b = s; }  // a > b: hang the larger root under the smaller
        int32_t old = atomicMin(&label[a], b);
        if (old == a) return;
        a = old;
    }
}

// each flat0 cell merges with its already-visited 8-neighbours (W, NW, N, NE)
__global__ void __launch_bounds__(256)
k_ccl_merge(const uint8_t *__restrict__ flat0, int32_t *label, int64_t R, int64_t C)
{
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= R || j >= C) return;
    const int64_t n = i * C + j;
    if (!flat0[n]) return;
    if (j > 0 && flat0[n - 1]) uf_union(label, (int32_t)n, (int32_t)(n - 1));
    if (i > 0) {
        if (j > 0 && flat0[n - C - 1]) uf_union(label, (int32_t)n, (int32_t)(n - C - 1));
        if (flat0[n - C]) uf_union(label, (int32_t)n, (int32_t)(n - C));
        if (j < C - 1 && flat0[n - C + 1]) uf_union(label, (int32_t)n, (int32_t)(n - C + 1));
    }
}

__global__ void __launch_bounds__(256)
k_ccl_flatten(const uint8_t *__restrict__ flat0, int32_t *label, int64_t N)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N || !flat0[n]) return;
    label[n] = uf_find(label, (int32_t)n);
}

__global__ void __launch_bounds__(256)
k_flats_extend(const double *__restrict__ E, const uint8_t *__restrict__ flat0,
               const int32_t *__restrict__ label, int64_t R, int64_t C,
               uint8_t *__restrict__ flats, double *__restrict__ mag, double *__restrict__ dir)
{
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= R || j >= C) return;
    const int64_t n = i * C + j;
    int32_t best = -1;
#pragma unroll
    for (int di = -1; di <= 1; di++) {
        const int64_t ni = i + di;
        if (ni < 0 || ni >= R) continue;
#pragma unroll
        for (int dj = -1; dj <= 1; dj++) {
            const int64_t nj = j + dj;
            if ((di == 0 && dj == 0) || nj < 0 || nj >= C) continue;
            const int64_t m = ni * C + nj;
            if (flat0[m]) best = max(best, label[m]);
        }
    }
    bool f = flat0[n] != 0;
    if (best >= 0) f = (E[n] == E[best]);                                    // 677
    flats[n] = f ? 1 : 0;
    if (f) { mag[n] = -1.0; dir[n] = -1.0; }                                 // 611-612
}

__global__ void __launch_bounds__(256)
k_find_flats(const double *__restrict__ mag, uint8_t *__restrict__ flats, int64_t N)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) flats[n] = (mag[n] == -1.0) ? 1 : 0;                          // 305-306
}

}  // namespace

int pdm_launch_flats(pdm_tile *t)
{
    dim3 block(32, 8);
    dim3 grid((unsigned)((t->C + 31) / 32), (unsigned)((t->R + 7) / 8));
    k_ccl_merge<<<grid, block, 0, t->stream>>>(t->flat0, t->label, t->R, t->C);
    PDM_LAUNCHED();
    k_ccl_flatten<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>(t->flat0, t->label, t->N);
    PDM_LAUNCHED();
    k_flats_extend<<<grid, block, 0, t->stream>>>(t->elev, t->flat0, t->label, t->R, t->C, t->flats, t->mag, t->dir);
    PDM_LAUNCHED();
    return PDM_OK;
}

int pdm_launch_find_flats(pdm_tile *t)
{
    k_find_flats<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>(t->mag, t->flats, t->N);
    PDM_LAUNCHED();
    return PDM_OK;
}
