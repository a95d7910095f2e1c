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
//
// Shards: regions can span row blocks.  Each local root carries the GLOBAL minimum cell
// index of its region and that cell's elevation (glabel / glelev); neighbouring ranks
// exchange these for their shared rows until nothing changes (pdm_tile_labels_*), after
// which the gather pass sees exactly the ordering one big tile would.
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

// read-only find for the flatten pass: concurrent threads only ever store roots there, so a
// walk that never writes cannot undo another thread's final label (path halving could: it may
// store a stale grandparent over a root written in the meantime)
__device__ __forceinline__ int32_t uf_find_ro(const int32_t *label, int32_t x)
{
    int32_t p = label[x];
    while (p != x) { x = p; p = label[x]; }
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

// each flat0 cell of local rows [r0, r1) merges with its already-visited 8-neighbours (W, NW, N, NE)
__global__ void __launch_bounds__(256)
k_ccl_merge(const uint8_t *__restrict__ flat0, int32_t *label, int64_t r0, int64_t r1, int64_t C)
{
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = r0 + (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= r1 || j >= C) return;
    const int64_t n = i * C + j;
    if (!flat0[n]) return;
    if (j > 0 && flat0[n - 1]) uf_union(label, (int32_t)n, (int32_t)(n - 1));
    if (i > r0) {
        if (j > 0 && flat0[n - C - 1]) uf_union(label, (int32_t)n, (int32_t)(n - C - 1));
        if (flat0[n - C]) uf_union(label, (int32_t)n, (int32_t)(n - C));
        if (j < C - 1 && flat0[n - C + 1]) uf_union(label, (int32_t)n, (int32_t)(n - C + 1));
    }
}

// label := root; roots of a shard also get their global identity
__global__ void __launch_bounds__(256)
k_ccl_flatten(const uint8_t *__restrict__ flat0, int32_t *label, int64_t n0, int64_t n1,
              const double *__restrict__ E, long long *__restrict__ gl, double *__restrict__ glE, long long goff)
{
    const int64_t n = n0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n1 || !flat0[n]) return;
    const int32_t r = uf_find_ro(label, (int32_t)n);
    label[n] = r;
    if (gl && r == (int32_t)n) { gl[n] = (long long)n + goff; glE[n] = E[n]; }
}

__global__ void __launch_bounds__(256)
k_flats_extend(const double *__restrict__ E, const uint8_t *__restrict__ flat0,
               const int32_t *__restrict__ label, Win w, const long long *__restrict__ gl,
               const double *__restrict__ glE,
               uint8_t *__restrict__ flats, double *__restrict__ mag, double *__restrict__ dir)
{
    const int64_t C = w.C;
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t i = w.lo + (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= w.hi || j >= C) return;
    const int64_t n = i * C + j;
    long long best = -1;
    int32_t best_root = -1;
#pragma unroll
    for (int di = -1; di <= 1; di++) {
        const int64_t ni = i + di;
        if (!w.row_in_grid(ni)) continue;
#pragma unroll
        for (int dj = -1; dj <= 1; dj++) {
            const int64_t nj = j + dj;
            if ((di == 0 && dj == 0) || nj < 0 || nj >= C) continue;
            const int64_t m = ni * C + nj;
            if (flat0[m]) {
                const int32_t r = label[m];
                const long long key = gl ? gl[r] : (long long)r;
                if (key > best) { best = key; best_root = r; }
            }
        }
    }
    bool f = flat0[n] != 0;
    if (best_root >= 0) f = (E[n] == (gl ? glE[best_root] : E[best_root]));  // 677
    flats[n] = f ? 1 : 0;
    if (f) { mag[n] = -1.0; dir[n] = -1.0; }                                 // 611-612
}

__global__ void __launch_bounds__(256)
k_find_flats(const double *__restrict__ mag, uint8_t *__restrict__ flats, int64_t N)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) flats[n] = (mag[n] == -1.0) ? 1 : 0;                          // 305-306
}

// ---- shard support -------------------------------------------------------------------------
// halo rows received from a neighbour: every flat0 cell starts as its own root
__global__ void __launch_bounds__(256)
k_label_init_rows(const uint8_t *__restrict__ flat0, int32_t *label, int64_t n0, int64_t n1)
{
    const int64_t n = n0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < n1 && flat0[n]) label[n] = (int32_t)n;
}

// one row -> (global label, elevation of that region's first cell); -1 where not flat0
__global__ void __launch_bounds__(256)
k_label_pack(const uint8_t *__restrict__ flat0, const int32_t *__restrict__ label, const long long *__restrict__ gl,
             const double *__restrict__ glE, int64_t row, int64_t C, long long *__restrict__ out_l,
             double *__restrict__ out_e)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= C) return;
    const int64_t n = row * C + j;
    if (flat0[n]) { const int32_t r = label[n]; out_l[j] = gl[r]; out_e[j] = glE[r]; }
    else { out_l[j] = -1; out_e[j] = 0.0; }
}

// the neighbour's view of my halo row: lower the region's global label where theirs is smaller
__global__ void __launch_bounds__(256)
k_label_unpack_min(const uint8_t *__restrict__ flat0, const int32_t *__restrict__ label, long long *gl, int64_t row,
                   int64_t C, const long long *__restrict__ in_l, unsigned long long *ctr)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= C) return;
    const int64_t n = row * C + j;
    if (!flat0[n] || in_l[j] < 0) return;
    const int32_t r = label[n];
    if (in_l[j] < gl[r]) {
        atomicMin(&gl[r], in_l[j]);
        atomicAdd(&ctr[CT_FLAG], 1ULL);
    }
}
__global__ void __launch_bounds__(256)
k_label_unpack_elev(const uint8_t *__restrict__ flat0, const int32_t *__restrict__ label, const long long *__restrict__ gl,
                    double *glE, int64_t row, int64_t C, const long long *__restrict__ in_l,
                    const double *__restrict__ in_e)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= C) return;
    const int64_t n = row * C + j;
    if (!flat0[n] || in_l[j] < 0) return;
    const int32_t r = label[n];
    if (in_l[j] == gl[r]) glE[r] = in_e[j];
}

}  // namespace

static bool sharded(const pdm_tile *t) { return t->win.lo != 0 || t->win.hi != t->R || t->win.Rg != t->R; }

// union-find over every local row whose flat0 is valid (a shard passes all rows after the
// flat0 halo exchange), then flatten
int pdm_launch_ccl(pdm_tile *t)
{
    const Win &w = t->win;
    dim3 block(32, 8);
    dim3 grid((unsigned)((w.C + 31) / 32), (unsigned)((t->R + 7) / 8));
    if (sharded(t)) {
        if (!t->glabel) {
            PDM_CUDA(cudaMalloc(&t->glabel, (size_t)t->N * sizeof(long long)));
            PDM_CUDA(cudaMalloc(&t->glelev, (size_t)t->N * sizeof(double)));
        }
        // halo rows hold flat0 bytes received from the neighbours
        if (w.lo > 0) {
            k_label_init_rows<<<(unsigned)((w.lo * w.C + 255) / 256), 256, 0, t->stream>>>(t->flat0, t->label, 0, w.lo * w.C);
            PDM_LAUNCHED();
        }
        if (w.hi < t->R) {
            k_label_init_rows<<<(unsigned)(((t->R - w.hi) * w.C + 255) / 256), 256, 0, t->stream>>>(
                t->flat0, t->label, w.hi * w.C, t->N);
            PDM_LAUNCHED();
        }
    }
    k_ccl_merge<<<grid, block, 0, t->stream>>>(t->flat0, t->label, 0, t->R, w.C);
    PDM_LAUNCHED();
    k_ccl_flatten<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>(
        t->flat0, t->label, 0, t->N, t->elev, sharded(t) ? t->glabel : nullptr, t->glelev, (long long)(w.row_off * w.C));
    PDM_LAUNCHED();
    return PDM_OK;
}

int pdm_launch_flats_extend(pdm_tile *t)
{
    const Win &w = t->win;
    dim3 block(32, 8);
    dim3 grid((unsigned)((w.C + 31) / 32), (unsigned)((w.hi - w.lo + 7) / 8));
    k_flats_extend<<<grid, block, 0, t->stream>>>(t->elev, t->flat0, t->label, w, sharded(t) ? t->glabel : nullptr,
                                                  t->glelev, t->flats, t->mag, t->dir);
    PDM_LAUNCHED();
    return PDM_OK;
}

int pdm_launch_flats(pdm_tile *t)
{
    int rc = pdm_launch_ccl(t);
    if (rc) return rc;
    return pdm_launch_flats_extend(t);
}

int pdm_launch_find_flats(pdm_tile *t)
{
    k_find_flats<<<(unsigned)((t->N + 255) / 256), 256, 0, t->stream>>>(t->mag, t->flats, t->N);
    PDM_LAUNCHED();
    return PDM_OK;
}

// shard label exchange: pack one local row for the neighbour / merge the neighbour's row
int pdm_launch_label_pack(pdm_tile *t, int64_t row, long long *out_l, double *out_e)
{
    k_label_pack<<<(unsigned)((t->C + 255) / 256), 256, 0, t->stream>>>(t->flat0, t->label, t->glabel, t->glelev, row,
                                                                        t->C, out_l, out_e);
    PDM_LAUNCHED();
    return PDM_OK;
}

int pdm_launch_label_unpack(pdm_tile *t, int64_t row, const long long *in_l, const double *in_e)
{
    k_label_unpack_min<<<(unsigned)((t->C + 255) / 256), 256, 0, t->stream>>>(t->flat0, t->label, t->glabel, row, t->C,
                                                                              in_l, t->d_counters);
    PDM_LAUNCHED();
    k_label_unpack_elev<<<(unsigned)((t->C + 255) / 256), 256, 0, t->stream>>>(t->flat0, t->label, t->glabel, t->glelev,
                                                                               row, t->C, in_l, in_e);
    PDM_LAUNCHED();
    return PDM_OK;
}
