// pits.cu -- K4: drains for pits / flats that the stencil left without a receiver.
//
// Reference behaviour: _mk_connectivity_pits (dem_processing.py:1269-1382) with
// utils.get_border_index (utils.py:313-340), _get_dX_mean (1993-1997), make_slice
// (utils.py:404-408).  Every cell of flats & elev>0 grows a region along its lowest
// border until a lower cell shows up, then drains to all lower border cells within
// drain_pits_max_dist, weighted by slope; mag[pit] = mean slope, flats[pit] = False.
//
// The reference visits pits one after the other in a Python loop, but a search only
// reads elev and the pre-loop pit mask, so pits are independent: one thread block per
// pit, region membership as a bitmap of the (2*max_iter+3)^2 window in shared memory,
// the region border as a compacted list that is rebuilt incrementally each iteration
// (block-wide min reductions instead of np.setdiff1d + min over Python arrays).
#include "pdm_internal.cuh"
#include <string.h>

#include "np_sum.cuh"

#define PIT_THREADS 128
#define PIT_CAP 16384  // border-list capacity per block (cells)

namespace {

struct PitArgs {
    const double *E;
    const uint8_t *pitmask;   // flats & elev > 0 before any pit is drained (1284)
    const double *dX, *dY;
    int64_t R, C;
    int64_t npits;
    const int32_t *pit_cell;
    int32_t *pit_beg, *pit_end, *pit_dst;
    double *pit_w;
    int64_t edge_cap;
    int32_t *scratch_i;       // [blocks][2][PIT_CAP]
    double *scratch_d;        // [blocks][PIT_CAP]
    uint8_t *flats, *link;
    double *mag, *prop;       // prop: proportion scratch (pit cells store their edge-list slot there)
    int32_t *indeg;           // pit in-edges per receiving cell
    unsigned long long *ctr;
    int max_iter, max_dist, min_border;
    double max_dist_xy;
    int W;                    // window radius = max_iter + 1
    // ---- row shard (k_pit_search<true>): the search also sees H rows of the neighbouring ranks (strips of
    //      elevation and pit mask received before the search); cells are addressed by an EXTENDED index
    //      (local row - lo + H) * C + column, dX / dY are the fences of the whole grid
    int64_t lo, hi, row_off, Rg;          // owned local rows, global row of local row 0, rows of the grid
    int64_t H, Hu, Hd;                    // coordinate offset; strip rows available above / below
    const double *E_up, *E_dn;
    const uint8_t *P_up, *P_dn;
    int32_t *in_up, *in_dn;               // pit edges arriving at the neighbours' cells: [Hin][C] each (rows next to the boundary)
    int64_t Hin;
    int64_t peer_row[2];                  // neighbour's local row = my local row + peer_row[side]
};

// cell access: plain local index on a stand-alone tile, extended index on a shard
template <bool SHARD> __device__ __forceinline__ int64_t x_row(const PitArgs &a, int32_t c) { return SHARD ? (int64_t)c / a.C - a.H + a.lo : (int64_t)c / a.C; }
template <bool SHARD> __device__ __forceinline__ int32_t x_idx(const PitArgs &a, int64_t i, int64_t j) { return SHARD ? (int32_t)((i - a.lo + a.H) * a.C + j) : (int32_t)(i * a.C + j); }
template <bool SHARD> __device__ __forceinline__ double x_elev(const PitArgs &a, int32_t c)
{
    if (!SHARD) return a.E[c];
    const int64_t i = x_row<true>(a, c), j = (int64_t)c % a.C;
    if (i < a.lo) return a.E_up[(i - (a.lo - a.Hu)) * a.C + j];
    if (i >= a.hi) return a.E_dn[(i - a.hi) * a.C + j];
    return a.E[i * a.C + j];
}
template <bool SHARD> __device__ __forceinline__ bool x_pit(const PitArgs &a, int32_t c)
{
    if (!SHARD) return a.pitmask[c] != 0;
    const int64_t i = x_row<true>(a, c), j = (int64_t)c % a.C;
    if (i < a.lo) return a.P_up[(i - (a.lo - a.Hu)) * a.C + j] != 0;
    if (i >= a.hi) return a.P_dn[(i - a.hi) * a.C + j] != 0;
    return a.pitmask[i * a.C + j] != 0;
}
// is (ni, nj) a cell of the grid?  On a shard *beyond = true when it is, but lies outside the rows this rank can see
template <bool SHARD> __device__ __forceinline__ bool x_inside(const PitArgs &a, int64_t ni, int64_t nj, bool *beyond)
{
    if (nj < 0 || nj >= a.C) return false;
    if (!SHARD) return ni >= 0 && ni < a.R;
    if (ni + a.row_off < 0 || ni + a.row_off >= a.Rg) return false;
    if (ni < a.lo - a.Hu || ni >= a.hi + a.Hd) { *beyond = true; return false; }
    return true;
}

// NaN-propagating min (np.min) over a block; every thread passes its partial
// (value, saw_nan, count) and gets the block result.
struct MinAcc {
    double v;
    int nan;
    int cnt;
};
__device__ __forceinline__ void acc_add(MinAcc &a, double e)
{
    if (e != e) a.nan = 1;
    else if (a.cnt == 0 || e < a.v) a.v = e;
    a.cnt++;
}
__device__ __forceinline__ void acc_merge(MinAcc &a, const MinAcc &b)
{
    if (b.cnt) {
        if (!a.cnt || b.v < a.v) a.v = b.v;  // v only meaningful when some non-NaN value was seen
    }
    a.nan |= b.nan;
    a.cnt += b.cnt;
}

template <bool SHARD>
__global__ void __launch_bounds__(PIT_THREADS)
k_pit_search(PitArgs a)
{
    extern __shared__ uint32_t seen[];  // window bitmap: cell is in the region or on its border list
    __shared__ int s_nb[2];             // border list lengths (double buffered)
    __shared__ MinAcc s_red[3][PIT_THREADS / 32];
    __shared__ int s_flag;
    const int tid = threadIdx.x;
    const int64_t C = a.C;
    const int64_t Rg = SHARD ? a.Rg : a.R;      // rows of the whole grid
    const int W = a.W, WW = 2 * W + 1;
    const int nwords = (WW * WW + 31) / 32;
    int32_t *B[2] = {a.scratch_i + (size_t)blockIdx.x * 2 * PIT_CAP, a.scratch_i + (size_t)blockIdx.x * 2 * PIT_CAP + PIT_CAP};
    double *S = a.scratch_d + (size_t)blockIdx.x * PIT_CAP;
    const int DR[8] = {-1, 0, 0, 1, -1, -1, 1, 1};
    const int DC[8] = {0, -1, 1, 0, -1, 1, -1, 1};

    for (int64_t slot = blockIdx.x; slot < a.npits; slot += gridDim.x) {
        const int32_t pit = a.pit_cell[slot];
        const int64_t ip = pit / C, jp = pit % C;
        for (int w = tid; w < nwords; w += PIT_THREADS) seen[w] = 0;
        if (tid == 0) { s_nb[0] = 0; s_nb[1] = 0; s_flag = 0; }
        __syncthreads();
        // region := {pit}; border := its in-bounds 8-neighbours (utils.py:313-340)
        if (tid == 0) {
            const int bit = W * WW + W;
            seen[bit >> 5] |= 1u << (bit & 31);
        }
        __syncthreads();
        if (tid < 8) {
            const int64_t ni = ip + DR[tid], nj = jp + DC[tid];
            bool beyond = false;
            if (x_inside<SHARD>(a, ni, nj, &beyond)) {
                const int bit = (int)(ni - ip + W) * WW + (int)(nj - jp + W);
                atomicOr(&seen[bit >> 5], 1u << (bit & 31));
                B[0][atomicAdd(&s_nb[0], 1)] = x_idx<SHARD>(a, ni, nj);
            }
            if (beyond) s_flag = 2;
        }
        __syncthreads();
        const double epit = a.E[pit];
        double epit_border = epit;
        int cur = 0, mode = 0;
        if (a.min_border) {                                                  // 1292-1297
            MinAcc m = {0.0, 0, 0};
            for (int t = tid; t < s_nb[0]; t += PIT_THREADS) acc_add(m, x_elev<SHARD>(a, B[0][t]));
            for (int o = 16; o > 0; o >>= 1) {
                MinAcc b = {__shfl_down_sync(0xffffffffu, m.v, o), __shfl_down_sync(0xffffffffu, m.nan, o), __shfl_down_sync(0xffffffffu, m.cnt, o)};
                acc_merge(m, b);
            }
            if ((tid & 31) == 0) s_red[0][tid >> 5] = m;
            __syncthreads();
            MinAcc tot = s_red[0][0];
            for (int w = 1; w < PIT_THREADS / 32; w++) acc_merge(tot, s_red[0][w]);
            if (tot.cnt) epit_border = tot.nan ? __longlong_as_double(0x7ff8000000000000LL) : tot.v;
            __syncthreads();
        }
        for (int it = 0; it < a.max_iter; it++) {                            // 1300
            const int nb = s_nb[cur];
            if (nb == 0) break;                                              // 1304-1305
            // min over the whole border, over its non-pit part and over its pit part
            MinAcc mall = {0.0, 0, 0}, mnp = {0.0, 0, 0}, mp = {0.0, 0, 0};
            for (int t = tid; t < nb; t += PIT_THREADS) {
                const int32_t c = B[cur][t];
                const double e = x_elev<SHARD>(a, c);
                acc_add(mall, e);
                if (x_pit<SHARD>(a, c)) acc_add(mp, e); else acc_add(mnp, e);
            }
            MinAcc *ms[3] = {&mall, &mnp, &mp};
            for (int k = 0; k < 3; k++) {
                MinAcc &m = *ms[k];
                for (int o = 16; o > 0; o >>= 1) {
                    MinAcc b = {__shfl_down_sync(0xffffffffu, m.v, o), __shfl_down_sync(0xffffffffu, m.nan, o), __shfl_down_sync(0xffffffffu, m.cnt, o)};
                    acc_merge(m, b);
                }
                if ((tid & 31) == 0) s_red[k][tid >> 5] = m;
            }
            __syncthreads();
            MinAcc tall = s_red[0][0], tnp = s_red[1][0], tp = s_red[2][0];
            for (int w = 1; w < PIT_THREADS / 32; w++) { acc_merge(tall, s_red[0][w]); acc_merge(tnp, s_red[1][w]); acc_merge(tp, s_red[2][w]); }
            __syncthreads();
            if (tnp.cnt > 0 && !tnp.nan && tnp.v < epit_border) { mode = 1; break; }   // 1312-1316
            if (tp.cnt > 0 && !tp.nan && tp.v < epit) { mode = 2; break; }             // 1317-1320
            if (tall.nan) break;  // emin is NaN: nothing equals it, the region can never grow again
            const double emin = tall.v;                                      // 1307
            // region += border cells at emin (1322-1323); border' = border - those + their new neighbours
            const int nxt = cur ^ 1;
            if (tid == 0) s_nb[nxt] = 0;
            __syncthreads();
            for (int t = tid; t < nb; t += PIT_THREADS) {
                const int32_t c = B[cur][t];
                if (x_elev<SHARD>(a, c) == emin) {
                    const int64_t ci = x_row<SHARD>(a, c), cj = c % C;
                    for (int k = 0; k < 8; k++) {
                        const int64_t ni = ci + DR[k], nj = cj + DC[k];
                        bool beyond = false;
                        if (!x_inside<SHARD>(a, ni, nj, &beyond)) { if (beyond) s_flag = 2; continue; }
                        const int bit = (int)(ni - ip + W) * WW + (int)(nj - jp + W);
                        const uint32_t mask = 1u << (bit & 31);
                        if (atomicOr(&seen[bit >> 5], mask) & mask) continue;
                        const int pos = atomicAdd(&s_nb[nxt], 1);
                        if (pos < PIT_CAP) B[nxt][pos] = x_idx<SHARD>(a, ni, nj); else s_flag = 1;
                    }
                } else {
                    const int pos = atomicAdd(&s_nb[nxt], 1);
                    if (pos < PIT_CAP) B[nxt][pos] = c; else s_flag = 1;
                }
            }
            __syncthreads();
            if (s_flag) break;
            cur = nxt;
        }
        __syncthreads();
        if (s_flag) {  // border list overflow (1) / region beyond the neighbours' strips (2): fail loudly on the host
            if (tid == 0) atomicAdd(&a.ctr[CT_ABORT], s_flag == 2 ? (1ULL << 32) : 1ULL);
            __syncthreads();
            continue;
        }
        if (mode == 0) {                                                     // 1327-1329
            if (tid == 0) { atomicAdd(&a.ctr[CT_PITS_UNDRAINED], 1ULL); a.pit_beg[slot] = 0; a.pit_end[slot] = 0; }
            __syncthreads();
            continue;
        }
        // drains := qualifying border cells, ascending index (setdiff1d order), distance filtered
        const int nb = s_nb[cur];
        const int oth = cur ^ 1;
        if (tid == 0) s_nb[oth] = 0;
        __syncthreads();
        for (int t = tid; t < nb; t += PIT_THREADS) {
            const int32_t c = B[cur][t];
            const double e = x_elev<SHARD>(a, c);
            const bool isp = x_pit<SHARD>(a, c);
            bool ok = (mode == 1) ? (!isp && e < epit_border) : (isp && e < epit);
            if (ok && a.max_dist > 0) {                                      // 1335-1343
                const int64_t di = ip - x_row<SHARD>(a, c), dj = jp - c % C;
                ok = sqrt((double)(di * di + dj * dj)) <= (double)a.max_dist;
            }
            if (ok) B[oth][atomicAdd(&s_nb[oth], 1)] = c;
        }
        __syncthreads();
        const int nd = s_nb[oth];
        // ascending index: small lists, rank sort by the block
        for (int t = tid; t < nd; t += PIT_THREADS) {
            const int32_t c = B[oth][t];
            int rank = 0;
            for (int u = 0; u < nd; u++) rank += (B[oth][u] < c);
            B[cur][rank] = c;
        }
        __syncthreads();
        if (tid == 0) {
            int32_t *D = B[cur];
            int n = nd;
            if (n == 0) {                                                    // 1338-1340
                atomicAdd(&a.ctr[CT_PITS_UNDRAINED], 1ULL);
                a.pit_beg[slot] = 0; a.pit_end[slot] = 0;
            } else {
                // real distances 1346-1349
                // (rows of the whole grid: on a shard dX / dY are the grid's fences)
                const int64_t gip = SHARD ? ip + a.row_off : ip;
                for (int t = 0; t < n; t++) {
                    const int64_t id = x_row<SHARD>(a, D[t]) + (SHARD ? a.row_off : 0), jd = D[t] % C;
                    double dxm, dy;
                    if (gip == id) { dxm = a.dX[gip < Rg - 2 ? gip : Rg - 2]; dy = 0.0; }
                    else {
                        const int64_t lo = gip < id ? gip : id, hi = gip < id ? id : gip;
                        dxm = __ddiv_rn(np_sum(a.dX + lo, hi - lo), (double)(hi - lo));
                        dy = np_sum(a.dY + lo, hi - lo);
                    }
                    const double dx = __dmul_rn(dxm, (double)(jp - jd));
                    S[t] = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                }
                if (a.max_dist_xy > 0.0) {                                   // 1352-1358
                    int k = 0;
                    for (int t = 0; t < n; t++)
                        if (S[t] <= a.max_dist_xy) { D[k] = D[t]; S[k] = S[t]; k++; }
                    n = k;
                }
                if (n == 0) {
                    atomicAdd(&a.ctr[CT_PITS_UNDRAINED], 1ULL);
                    a.pit_beg[slot] = 0; a.pit_end[slot] = 0;
                } else {
                    for (int t = 0; t < n; t++) S[t] = __ddiv_rn(fabs(__dsub_rn(epit, x_elev<SHARD>(a, D[t]))), S[t]);  // 1361
                    const double ssum = np_sum(S, n);
                    // edges that survive the matrix filter (1136-1137): weight > 1e-8, not NaN
                    int kept = 0;
                    for (int t = 0; t < n; t++) {
                        const double w = __ddiv_rn(S[t], ssum);
                        if (w > 1e-8 && x_elev<SHARD>(a, D[t]) <= epit) kept++;
                    }
                    const unsigned long long base = atomicAdd(&a.ctr[CT_NPITEDGES], (unsigned long long)kept);
                    if ((int64_t)(base + kept) > a.edge_cap) {
                        atomicAdd(&a.ctr[CT_FLAG], 1ULL);  // edge buffer too small: host grows it and reruns
                        a.pit_beg[slot] = 0; a.pit_end[slot] = 0;
                    } else {
                        int k = 0;
                        for (int t = 0; t < n; t++) {
                            const double w = __ddiv_rn(S[t], ssum);                      // 1367
                            if (w > 1e-8 && x_elev<SHARD>(a, D[t]) <= epit) {
                                int32_t dst = D[t];
                                if (SHARD) {
                                    // own cell: local index; a neighbour's cell: ~(side << 30 | its own index there), and
                                    // the edge is counted in the strip that travels to that rank before its in-degrees are built
                                    const int64_t di = x_row<true>(a, dst), dj = dst % C;
                                    if (di < a.lo) {
                                        dst = ~(int32_t)((di + a.peer_row[0]) * C + dj);
                                        if (di >= a.lo - a.Hin) atomicAdd(a.in_up + (di - (a.lo - a.Hin)) * C + dj, 1); else atomicAdd(&a.ctr[CT_ABORT], 1ULL << 48);
                                    } else if (di >= a.hi) {
                                        dst = ~(int32_t)((1 << 30) | (int32_t)((di + a.peer_row[1]) * C + dj));
                                        if (di < a.hi + a.Hin) atomicAdd(a.in_dn + (di - a.hi) * C + dj, 1); else atomicAdd(&a.ctr[CT_ABORT], 1ULL << 48);
                                    } else {
                                        dst = (int32_t)(di * C + dj);
                                        atomicAdd(a.indeg + dst, 1);
                                    }
                                } else {
                                    atomicAdd(a.indeg + dst, 1);
                                }
                                a.pit_dst[base + k] = dst;
                                a.pit_w[base + k] = w;
                                k++;
                            }
                        }
                        a.pit_beg[slot] = (int32_t)base;
                        a.pit_end[slot] = (int32_t)(base + kept);
                        a.link[pit] = LK_NOSEC | LK_PIT;
                        a.prop[pit] = __longlong_as_double((long long)slot);
                    }
                    a.mag[pit] = __ddiv_rn(ssum, (double)n);                             // 1370 (np.mean)
                    a.flats[pit] = 0;                                                    // 1371
                }
            }
        }
        __syncthreads();
    }
}

// compact the pit mask into a list (order is irrelevant to the result)
__global__ void __launch_bounds__(256)
k_pit_compact(const uint8_t *__restrict__ pitmask, int64_t n0, int64_t N, int32_t *__restrict__ pit_cell, unsigned long long *ctr)
{
    const int64_t n = n0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // owned cells [n0, N)
    const bool is = n < N && pitmask[n];
    const unsigned m = __ballot_sync(0xffffffffu, is);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&ctr[CT_TMP0], (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (is) pit_cell[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)n;
}

// after the search: what changed at each examined pit (for sparse host-side updates)
__global__ void __launch_bounds__(256)
k_pit_gather(const int32_t *__restrict__ pit_cell, int64_t npits, const double *__restrict__ mag,
             const uint8_t *__restrict__ flats, double *__restrict__ out_mag, uint8_t *__restrict__ out_flat)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= npits) return;
    const int32_t c = pit_cell[k];
    out_mag[k] = mag[c];
    out_flat[k] = flats[c];
}

}  // namespace

// cells examined by the last pit search with their current mag / flats values (host buffers of
// t->n_pits entries each)
int pdm_launch_pit_readback(pdm_tile *t, int32_t *cells, double *mag, uint8_t *flats)
{
    const int64_t n = t->n_pits;
    if (n == 0) return PDM_OK;
    double *d_mag = nullptr; uint8_t *d_fl = nullptr;
    PDM_CUDA(cudaMalloc(&d_mag, (size_t)n * 8));
    PDM_CUDA(cudaMalloc(&d_fl, (size_t)n));
    k_pit_gather<<<(unsigned)((n + 255) / 256), 256, 0, t->stream>>>(t->pit_cell, n, t->mag, t->flats, d_mag, d_fl);
    PDM_LAUNCHED();
    PDM_CUDA(cudaMemcpyAsync(cells, t->pit_cell, (size_t)n * 4, cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaMemcpyAsync(mag, d_mag, (size_t)n * 8, cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaMemcpyAsync(flats, d_fl, (size_t)n, cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    cudaFree(d_mag); cudaFree(d_fl);
    return PDM_OK;
}

static int read_ctr(pdm_tile *t)
{
    PDM_CUDA(cudaMemcpyAsync(t->h_counters, t->d_counters, CT_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    return PDM_OK;
}

// Called after k_links has written the pit mask into t->flat0 and counted it in CT_NPITS.
// sh != nullptr: the tile is a row shard and the search also sees the neighbours' strips (pdm_shard_pits).
int pdm_launch_pits(pdm_tile *t, const pdm_uca_params *p, const pdm_pit_shard *sh)
{
    int rc = read_ctr(t);
    if (rc) return rc;
    const int64_t npits = (int64_t)t->h_counters[CT_NPITS];
    t->n_pits = npits;
    t->n_pit_edges = 0;
    if (npits == 0) return PDM_OK;
    const int W = (int)p->drain_pits_max_iter + 1, WW = 2 * W + 1;
    const size_t smem = (size_t)((WW * WW + 31) / 32) * 4;
    if (p->drain_pits_max_iter < 1 || smem > 200 * 1024) {
        pdm_set_error("drain_pits_max_iter=%lld outside the supported range [1, 630] of the GPU pit search",
                      (long long)p->drain_pits_max_iter);
        return PDM_ERR_ARG;
    }
    void (*kernel)(PitArgs) = sh ? k_pit_search<true> : k_pit_search<false>;
    PDM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, occ = 0;
    PDM_CUDA(cudaGetDevice(&dev));
    PDM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PDM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, PIT_THREADS, smem));
    if (occ < 1) occ = 1;
    int64_t blocks = (int64_t)sms * occ;
    if (blocks > npits) blocks = npits;
    if (npits > t->pit_cap) {
        if (t->pit_cell) { cudaFree(t->pit_cell); cudaFree(t->pit_beg); cudaFree(t->pit_end); t->pit_cell = t->pit_beg = t->pit_end = nullptr; }
        t->pit_cap = npits + npits / 4 + 64;
        PDM_CUDA(cudaMalloc(&t->pit_cell, (size_t)t->pit_cap * 4));
        PDM_CUDA(cudaMalloc(&t->pit_beg, (size_t)t->pit_cap * 4));
        PDM_CUDA(cudaMalloc(&t->pit_end, (size_t)t->pit_cap * 4));
    }
    if (blocks > t->pit_scratch_blocks) {
        if (t->pit_scratch_i) { cudaFree(t->pit_scratch_i); cudaFree(t->pit_scratch_d); t->pit_scratch_i = nullptr; t->pit_scratch_d = nullptr; }
        t->pit_scratch_blocks = (int64_t)sms * occ;
        PDM_CUDA(cudaMalloc(&t->pit_scratch_i, (size_t)t->pit_scratch_blocks * 2 * PIT_CAP * 4));
        PDM_CUDA(cudaMalloc(&t->pit_scratch_d, (size_t)t->pit_scratch_blocks * PIT_CAP * 8));
    }
    if (t->pit_edge_cap < 4 * npits + 1024) {
        if (t->pit_dst) { cudaFree(t->pit_dst); cudaFree(t->pit_w); t->pit_dst = nullptr; t->pit_w = nullptr; }
        t->pit_edge_cap = 4 * npits + 1024;
        PDM_CUDA(cudaMalloc(&t->pit_dst, (size_t)t->pit_edge_cap * 4));
        PDM_CUDA(cudaMalloc(&t->pit_w, (size_t)t->pit_edge_cap * 8));
    }
    // t->label (free after the flat labelling) counts the pit edges arriving at each cell
    PDM_CUDA(cudaMemsetAsync(t->label, 0, (size_t)t->N * sizeof(int32_t), t->stream));
    PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_TMP0, 0, sizeof(unsigned long long), t->stream));
    {
        const int64_t n0 = t->win.lo * t->C, n1 = t->win.hi * t->C;    // (halo rows of flat0 do not hold the pit mask)
        k_pit_compact<<<(unsigned)((n1 - n0 + 255) / 256), 256, 0, t->stream>>>(t->flat0, n0, n1, t->pit_cell, t->d_counters);
        PDM_LAUNCHED();
    }
    for (int attempt = 0; attempt < 8; attempt++) {
        PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_NPITEDGES, 0, sizeof(unsigned long long), t->stream));
        PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_PITS_UNDRAINED, 0, sizeof(unsigned long long), t->stream));
        PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_FLAG, 0, sizeof(unsigned long long), t->stream));
        PDM_CUDA(cudaMemsetAsync(t->d_counters + CT_ABORT, 0, sizeof(unsigned long long), t->stream));
        PitArgs a;
        memset(&a, 0, sizeof(a));
        if (sh) {
            a.lo = t->win.lo; a.hi = t->win.hi; a.row_off = t->win.row_off; a.Rg = t->win.Rg;
            a.H = W; a.Hu = sh->Hu; a.Hd = sh->Hd;
            a.E_up = sh->E_up; a.E_dn = sh->E_dn; a.P_up = sh->P_up; a.P_dn = sh->P_dn;
            a.in_up = sh->in_up; a.in_dn = sh->in_dn; a.Hin = sh->Hin;
            a.peer_row[0] = sh->peer_row[0]; a.peer_row[1] = sh->peer_row[1];
            if (sh->Hin > 0) {
                if (sh->in_up) PDM_CUDA(cudaMemsetAsync(sh->in_up, 0, (size_t)sh->Hin * t->C * 4, t->stream));
                if (sh->in_dn) PDM_CUDA(cudaMemsetAsync(sh->in_dn, 0, (size_t)sh->Hin * t->C * 4, t->stream));
            }
        }
        a.E = t->elev; a.pitmask = t->flat0; a.dX = sh ? sh->dXg : t->dX; a.dY = sh ? sh->dYg : t->dY; a.R = t->R; a.C = t->C; a.npits = npits;
        a.pit_cell = t->pit_cell; a.pit_beg = t->pit_beg; a.pit_end = t->pit_end; a.pit_dst = t->pit_dst; a.pit_w = t->pit_w;
        a.edge_cap = t->pit_edge_cap; a.scratch_i = t->pit_scratch_i; a.scratch_d = t->pit_scratch_d;
        a.flats = t->flats; a.link = t->link; a.mag = t->mag; a.prop = t->twi; a.indeg = t->label; a.ctr = t->d_counters;
        a.max_iter = (int)p->drain_pits_max_iter; a.max_dist = (int)p->drain_pits_max_dist;
        a.min_border = p->drain_pits_min_border; a.max_dist_xy = p->drain_pits_max_dist_xy; a.W = W;
        kernel<<<(unsigned)blocks, PIT_THREADS, smem, t->stream>>>(a);
        PDM_LAUNCHED();
        rc = read_ctr(t);
        if (rc) return rc;
        if (t->h_counters[CT_ABORT] >> 32) {
            pdm_set_error("pit search on a row shard: %llu pit region(s) reach beyond the %lld / %lld rows of the neighbouring ranks this rank "
                          "can see, %llu drain(s) lie farther than %lld rows across the boundary (shards need >= drain_pits_max_iter + 1 rows)",
                          (t->h_counters[CT_ABORT] >> 32) & 0xffffULL, (long long)(sh ? sh->Hu : 0), (long long)(sh ? sh->Hd : 0),
                          t->h_counters[CT_ABORT] >> 48, (long long)(sh ? sh->Hin : 0));
            return PDM_ERR_ARG;
        }
        if (t->h_counters[CT_ABORT]) {
            pdm_set_error("pit search: region border of %llu pit(s) exceeded %d cells", t->h_counters[CT_ABORT], PIT_CAP);
            return PDM_ERR_NOMEM;
        }
        if (!t->h_counters[CT_FLAG]) {
            t->n_pit_edges = (int64_t)t->h_counters[CT_NPITEDGES];
            return PDM_OK;
        }
        // edge buffer overflow: the search is idempotent except for indeg (+= per edge); undo and retry bigger
        const int64_t need = (int64_t)t->h_counters[CT_NPITEDGES];
        cudaFree(t->pit_dst); cudaFree(t->pit_w); t->pit_dst = nullptr; t->pit_w = nullptr;
        t->pit_edge_cap = need + need / 4 + 1024;
        PDM_CUDA(cudaMalloc(&t->pit_dst, (size_t)t->pit_edge_cap * 4));
        PDM_CUDA(cudaMalloc(&t->pit_w, (size_t)t->pit_edge_cap * 8));
        PDM_CUDA(cudaMemsetAsync(t->label, 0, (size_t)t->N * sizeof(int32_t), t->stream));
    }
    pdm_set_error("pit search: edge buffer kept overflowing");
    return PDM_ERR_NOMEM;
}
