// api.cu -- C ABI of libpydem_b200 (include/pydem_b200.h): tile lifetime, host<->device
// movement and the stage drivers.  Host code only; the kernels live in slopes.cu, flats.cu,
// graph.cu, pits.cu, sweep.cu, update.cu.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "tsweep.cuh"

static thread_local char g_err[512] = "";
unsigned long long g_pdm_launches = 0;

void pdm_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int pdm_cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    pdm_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return PDM_ERR_CUDA;
}

static size_t field_elem_size(int field)
{
    switch (field) {
        case PDM_F_FLATS: case PDM_F_EDGE_TODO: case PDM_F_EDGE_DONE: case PDM_F_SECTION:
        case PDM_F_FLAT0: case PDM_F_LINK: return 1;
        case PDM_F_CELL: return 32;
        default: return 8;
    }
}

static void *field_ptr(pdm_tile *t, int field)
{
    switch (field) {
        case PDM_F_ELEV: return t->elev;
        case PDM_F_MAG: return t->mag;
        case PDM_F_DIR: return t->dir;
        case PDM_F_FLATS: return t->flats;
        case PDM_F_UCA: return t->uca;
        case PDM_F_TWI: return t->twi;
        case PDM_F_EDGE_TODO: return t->edge_todo;
        case PDM_F_EDGE_DONE: return t->edge_done;
        case PDM_F_SECTION: return t->section;
        case PDM_F_TWI10: return t->twi10;
        case PDM_F_FLAT0: return t->flat0;
        case PDM_F_LINK: return t->link;
        case PDM_F_CELL: return t->cell;
        default: return nullptr;
    }
}

static bool standalone(const pdm_tile *t) { return t->win.lo == 0 && t->win.hi == t->R && t->win.Rg == t->R && t->win.row_off == 0; }

extern "C" {

int pdm_abi_version(void) { return PDM_ABI_VERSION; }
const char *pdm_last_error(void) { return g_err; }

unsigned long long pdm_launch_count(void) { return g_pdm_launches; }

int pdm_host_alloc(size_t bytes, void **out)
{
    if (!out) { pdm_set_error("pdm_host_alloc: NULL out"); return PDM_ERR_ARG; }
    PDM_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return PDM_OK;
}

int pdm_host_free(void *p)
{
    if (p) PDM_CUDA(cudaFreeHost(p));
    return PDM_OK;
}

int pdm_device_count(int *count)
{
    if (!count) { pdm_set_error("pdm_device_count: NULL"); return PDM_ERR_ARG; }
    PDM_CUDA(cudaGetDeviceCount(count));
    return PDM_OK;
}

int pdm_init(int device)
{
    int n = 0;
    PDM_CUDA(cudaGetDeviceCount(&n));
    if (n <= 0) { pdm_set_error("pdm_init: no CUDA device (this library has no CPU fallback)"); return PDM_ERR_CUDA; }
    if (device < 0 || device >= n) { pdm_set_error("pdm_init: device %d out of range [0,%d)", device, n); return PDM_ERR_ARG; }
    PDM_CUDA(cudaSetDevice(device));
    PDM_CUDA(cudaFree(0));
    return PDM_OK;
}

void pdm_default_uca_params(pdm_uca_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->drain_pits = 1;
    p->drain_pits_min_border = 0;
    p->drain_pits_max_iter = 300;
    p->drain_pits_max_dist = 32;
    p->drain_pits_max_dist_xy = 0.0;
    p->apply_uca_limit_edges = 0;
    p->circular_ref_maxcount = 50;
    p->uca_saturation_limit = 32.0;
}

void pdm_default_twi_params(pdm_twi_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->twi_min_slope = 1e-3;
    p->twi_min_area = INFINITY;
    p->uca_saturation_limit = 32.0;
}

int pdm_tile_create(int64_t R, int64_t C, void *stream, pdm_tile **out)
{
    if (!out) { pdm_set_error("pdm_tile_create: NULL out"); return PDM_ERR_ARG; }
    *out = nullptr;
    if (R < 3 || C < 3) { pdm_set_error("pdm_tile_create: tile must be at least 3x3 (got %lld x %lld)", (long long)R, (long long)C); return PDM_ERR_ARG; }
    if (R * C >= (int64_t)2147483000) { pdm_set_error("pdm_tile_create: %lld cells exceed the 32-bit cell index of one tile; shard the grid", (long long)(R * C)); return PDM_ERR_ARG; }
    pdm_tile *t = new pdm_tile();
    memset(t, 0, sizeof(*t));
    t->R = R; t->C = C; t->N = R * C;
    t->win.R = R; t->win.C = C; t->win.row_off = 0; t->win.Rg = R; t->win.lo = 0; t->win.hi = R;
    t->stream = (cudaStream_t)stream;
    cudaError_t e = cudaGetDevice(&t->device);
    if (e != cudaSuccess) { delete t; return pdm_cuda_fail(e, "cudaGetDevice", __FILE__, __LINE__); }
    const size_t N = (size_t)t->N;
    struct { void **p; size_t bytes; } allocs[] = {
        {(void **)&t->elev, N * 8}, {(void **)&t->mag, N * 8}, {(void **)&t->dir, N * 8}, {(void **)&t->uca, N * 8},
        {(void **)&t->cell, N * sizeof(Cell)}, {(void **)&t->twi, N * 8},
        {(void **)&t->flats, N}, {(void **)&t->flat0, N + 8}, {(void **)&t->link, N}, {(void **)&t->edge_todo, N},
        {(void **)&t->edge_done, N},
        {(void **)&t->label, N * 4}, {(void **)&t->queue, (N + 1) * 4},
        {(void **)&t->dX, (size_t)R * 8}, {(void **)&t->dY, (size_t)R * 8}, {(void **)&t->dg, (size_t)R * 8},
        {(void **)&t->thA, (size_t)R * 8}, {(void **)&t->thB, (size_t)R * 8}, {(void **)&t->th_row, (size_t)R * 8},
        {(void **)&t->row_area, (size_t)R * 8}, {(void **)&t->rdX, (size_t)R * 8}, {(void **)&t->rdY, (size_t)R * 8},
        {(void **)&t->rdg, (size_t)R * 8},
        {(void **)&t->d_counters, CT_N * sizeof(unsigned long long)},
    };
    for (auto &a : allocs) {
        e = cudaMalloc(a.p, a.bytes);
        if (e != cudaSuccess) {
            pdm_cuda_fail(e, "cudaMalloc(tile field)", __FILE__, __LINE__);
            pdm_tile_destroy(t);
            return e == cudaErrorMemoryAllocation ? PDM_ERR_NOMEM : PDM_ERR_CUDA;
        }
    }
    e = cudaMallocHost((void **)&t->h_counters, CT_N * sizeof(unsigned long long));
    if (e != cudaSuccess) { pdm_cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__); pdm_tile_destroy(t); return PDM_ERR_CUDA; }
    for (int k = 0; k < 4; k++) {
        e = cudaEventCreate(&t->ev[k]);
        if (e != cudaSuccess) { pdm_cuda_fail(e, "cudaEventCreate", __FILE__, __LINE__); pdm_tile_destroy(t); return PDM_ERR_CUDA; }
    }
    cudaMemsetAsync(t->d_counters, 0, CT_N * sizeof(unsigned long long), t->stream);
    e = cudaStreamCreateWithFlags(&t->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&t->copy_ev, cudaEventDisableTiming);
    if (e != cudaSuccess) { pdm_cuda_fail(e, "copy stream", __FILE__, __LINE__); pdm_tile_destroy(t); return PDM_ERR_CUDA; }
    t->min_area = INFINITY;
    *out = t;
    return PDM_OK;
}

int pdm_tile_destroy(pdm_tile *t)
{
    if (!t) return PDM_OK;
    void *ptrs[] = {t->elev, t->mag, t->dir, t->uca, t->cell, t->twi, t->flats, t->flat0, t->link,
                    t->edge_todo, t->edge_done, t->section, t->label, t->queue, t->dX, t->dY, t->dg,
                    t->thA, t->thB, t->th_row, t->row_area, t->d_counters, t->pit_cell, t->pit_beg, t->pit_end,
                    t->pit_dst, t->pit_w, t->pit_scratch_i, t->pit_scratch_d, t->edge_buf_d, t->edge_buf_b,
                    t->glabel, t->glelev, t->twi10, t->rdX, t->rdY, t->rdg, t->ts_ctl, t->ts_seen, t->dXg, t->dYg};
    pdm_ts_p2p_close(t);
    pdm_comm_scratch_free(t);
    for (void *p : ptrs) if (p) cudaFree(p);
    if (t->h_counters) cudaFreeHost(t->h_counters);
    if (t->ts_hctr) cudaFreeHost(t->ts_hctr);
    for (int k = 0; k < 4; k++) if (t->ev[k]) cudaEventDestroy(t->ev[k]);
    if (t->copy_stream) { cudaStreamSynchronize(t->copy_stream); cudaStreamDestroy(t->copy_stream); }
    if (t->copy_ev) cudaEventDestroy(t->copy_ev);
    for (int k = 0; k < 16; k++) if (t->up_ev[k]) cudaEventDestroy(t->up_ev[k]);
    delete t;
    return PDM_OK;
}

int pdm_tile_set_spacing(pdm_tile *t, const double *dX, const double *dY, const double *dX2, const double *dY2,
                         const double *thA, const double *thB)
{
    if (!t || !dX || !dY || !dX2 || !dY2) { pdm_set_error("pdm_tile_set_spacing: NULL argument"); return PDM_ERR_ARG; }
    const int64_t R = t->R;
    std::vector<double> a((size_t)R), b((size_t)R), area((size_t)R), thr((size_t)R);
    for (int64_t f = 0; f < R - 1; f++) {
        a[f] = thA ? thA[f] : atan2(dY[f], dX[f]);   // dem_processing.py:1936, facets 0,3,4,7
        b[f] = thB ? thB[f] : atan2(dX[f], dY[f]);   // facets 1,2,5,6
    }
    // theta of _calc_uca_section_proportion for a stand-alone tile: facet-0 theta of fences
    // 0..R-3 padded with its first and last entry (dem_processing.py:1031-1033); a shard gets
    // its slice of the global array through pdm_tile_set_window
    for (int64_t i = 0; i < R; i++) thr[i] = a[i == 0 ? 0 : (i == R - 1 ? R - 3 : i - 1)];
    double mn = NAN;
    for (int64_t i = 0; i < R; i++) {
        area[i] = dX2[i] * dY2[i];                    // 885
        if (area[i] == area[i] && !(mn <= area[i])) mn = area[i];  // nanmin (898)
    }
    t->min_area = mn;
    PDM_CUDA(cudaMemcpyAsync(t->dX, dX, (size_t)(R - 1) * 8, cudaMemcpyHostToDevice, t->stream));
    PDM_CUDA(cudaMemcpyAsync(t->dY, dY, (size_t)(R - 1) * 8, cudaMemcpyHostToDevice, t->stream));
    PDM_CUDA(cudaMemcpyAsync(t->thA, a.data(), (size_t)(R - 1) * 8, cudaMemcpyHostToDevice, t->stream));
    PDM_CUDA(cudaMemcpyAsync(t->thB, b.data(), (size_t)(R - 1) * 8, cudaMemcpyHostToDevice, t->stream));
    PDM_CUDA(cudaMemcpyAsync(t->row_area, area.data(), (size_t)R * 8, cudaMemcpyHostToDevice, t->stream));
    PDM_CUDA(cudaMemcpyAsync(t->th_row, thr.data(), (size_t)R * 8, cudaMemcpyHostToDevice, t->stream));
    int rc = pdm_launch_geometry(t);
    if (rc) return rc;
    PDM_CUDA(cudaStreamSynchronize(t->stream));  // the staging vectors die here
    t->have_spacing = true;
    t->have_graph = false;
    return PDM_OK;
}

int pdm_tile_upload(pdm_tile *t, int field, const void *host)
{
    if (!t || !host) { pdm_set_error("pdm_tile_upload: NULL argument"); return PDM_ERR_ARG; }
    void *d = field_ptr(t, field);
    if (!d || field == PDM_F_SECTION) { pdm_set_error("pdm_tile_upload: field %d cannot be uploaded", field); return PDM_ERR_ARG; }
    PDM_CUDA(cudaMemcpyAsync(d, host, (size_t)t->N * field_elem_size(field), cudaMemcpyHostToDevice, t->stream));
    if (field == PDM_F_ELEV) { t->have_elev = true; t->have_graph = false; }
    if (field == PDM_F_MAG || field == PDM_F_DIR) { t->have_slopes = true; t->have_graph = false; }
    if (field == PDM_F_FLATS) { t->have_flats = true; t->have_graph = false; }
    if (field == PDM_F_UCA) t->have_uca = true;
    return PDM_OK;
}

int pdm_tile_download(pdm_tile *t, int field, void *host)
{
    if (!t || !host) { pdm_set_error("pdm_tile_download: NULL argument"); return PDM_ERR_ARG; }
    if (field == PDM_F_SECTION) {
        if (!t->have_slopes || !t->have_flats || !t->have_spacing) { pdm_set_error("pdm_tile_download(SECTION): needs direction, flats and spacing"); return PDM_ERR_STATE; }
        int rc = pdm_launch_section_export(t);
        if (rc) return rc;
    }
    void *d = field_ptr(t, field);
    if (!d) { pdm_set_error("pdm_tile_download: unknown field %d", field); return PDM_ERR_ARG; }
    PDM_CUDA(cudaMemcpyAsync(host, d, (size_t)t->N * field_elem_size(field), cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    return PDM_OK;
}

// device -> host copy of one field, ordered after everything queued on the tile so far but not
// blocking the tile's stream: the copy overlaps the next stage.  The host buffer (page-locked
// for a truly asynchronous copy) is valid after pdm_tile_sync().
int pdm_tile_download_async(pdm_tile *t, int field, void *host)
{
    if (!t || !host) { pdm_set_error("pdm_tile_download_async: NULL argument"); return PDM_ERR_ARG; }
    if (field == PDM_F_SECTION) return pdm_tile_download(t, field, host);
    void *d = field_ptr(t, field);
    if (!d) { pdm_set_error("pdm_tile_download_async: unknown field %d", field); return PDM_ERR_ARG; }
    PDM_CUDA(cudaEventRecord(t->copy_ev, t->stream));
    PDM_CUDA(cudaStreamWaitEvent(t->copy_stream, t->copy_ev, 0));
    PDM_CUDA(cudaMemcpyAsync(host, d, (size_t)t->N * field_elem_size(field), cudaMemcpyDeviceToHost, t->copy_stream));
    return PDM_OK;
}

int pdm_tile_device_ptr(pdm_tile *t, int field, void **dev)
{
    if (!t || !dev) { pdm_set_error("pdm_tile_device_ptr: NULL argument"); return PDM_ERR_ARG; }
    void *d = field_ptr(t, field);
    if (!d) { pdm_set_error("pdm_tile_device_ptr: field %d not available", field); return PDM_ERR_ARG; }
    *dev = d;
    return PDM_OK;
}

int pdm_tile_mark_resident(pdm_tile *t, int field)
{
    // a field written through pdm_tile_device_ptr (e.g. by a torch tensor view) is declared valid here
    if (!t) return PDM_ERR_ARG;
    if (field == PDM_F_ELEV) { t->have_elev = true; t->have_graph = false; }
    if (field == PDM_F_MAG || field == PDM_F_DIR) { t->have_slopes = true; t->have_graph = false; }
    if (field == PDM_F_FLATS) { t->have_flats = true; t->have_graph = false; }
    if (field == PDM_F_UCA) t->have_uca = true;
    return PDM_OK;
}

// Device-resident mosaic tiles: keep the drainage graph (link bytes, proportions, pit edge lists) of the last
// pdm_tile_uca for the following pdm_tile_uca_update calls instead of rebuilding it per call.
int pdm_tile_set_keep_graph(pdm_tile *t, int on)
{
    if (!t) { pdm_set_error("NULL tile"); return PDM_ERR_ARG; }
    t->keep_graph = on != 0;
    return PDM_OK;
}

int pdm_selftest_division(unsigned long long seed, long long n_pairs, unsigned long long *mismatches)
{
    if (!mismatches || n_pairs <= 0) { pdm_set_error("pdm_selftest_division: bad argument"); return PDM_ERR_ARG; }
    const int blocks = 592;
    long long per_thread = (n_pairs + blocks * 256 - 1) / (blocks * 256);
    return pdm_launch_selftest_div(seed, blocks, per_thread, mismatches);
}

int pdm_tile_set_stencil_parity(pdm_tile *t, int on)
{
    if (!t) return PDM_ERR_ARG;
    t->stencil_parity = on != 0;
    return PDM_OK;
}

int pdm_tile_sync(pdm_tile *t)
{
    if (!t) return PDM_ERR_ARG;
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->copy_stream));
    return PDM_OK;
}

int pdm_tile_slopes_directions(pdm_tile *t)
{
    if (!t) { pdm_set_error("NULL tile"); return PDM_ERR_ARG; }
    if (!t->have_elev || !t->have_spacing) { pdm_set_error("pdm_tile_slopes_directions: upload ELEV and set spacing first"); return PDM_ERR_STATE; }
    if (!standalone(t)) { pdm_set_error("pdm_tile_slopes_directions: tile is a row shard; use the pdm_shard_* stages"); return PDM_ERR_STATE; }
    int rc = pdm_launch_slopes(t);
    if (rc) return rc;
    rc = pdm_launch_flats(t);
    if (rc) return rc;
    t->have_slopes = true; t->have_flats = true; t->have_graph = false;
    return PDM_OK;
}

// ELEV from the host and calc_slopes_directions in one call: the elevation goes up in row chunks on the copy stream
// and the stencil of a chunk starts as soon as its rows (and the first row of the next chunk) have arrived, so the
// stencil hides behind the host -> device copy (the copy is the longer of the two: ~55 GB/s PCIe vs ~0.2 TB/s stencil).
// Same results as pdm_tile_upload + pdm_tile_slopes_directions.
int pdm_tile_upload_slopes_directions(pdm_tile *t, const void *host_elev)
{
    if (!t || !host_elev) { pdm_set_error("pdm_tile_upload_slopes_directions: NULL argument"); return PDM_ERR_ARG; }
    if (!t->have_spacing) { pdm_set_error("pdm_tile_upload_slopes_directions: set the spacing first"); return PDM_ERR_STATE; }
    if (!standalone(t)) { pdm_set_error("pdm_tile_upload_slopes_directions: tile is a row shard; use the pdm_shard_* stages"); return PDM_ERR_STATE; }
    const int64_t R = t->R, C = t->C;
    int K = 8;
    if (R < 64 * K) K = (int)(R / 64 > 0 ? R / 64 : 1);
    const char *host = reinterpret_cast<const char *>(host_elev);
    // the copy stream starts behind whatever still reads the old ELEV on the tile's stream
    PDM_CUDA(cudaEventRecord(t->copy_ev, t->stream));
    PDM_CUDA(cudaStreamWaitEvent(t->copy_stream, t->copy_ev, 0));
    int64_t edge[17];
    for (int k = 0; k <= K; k++) edge[k] = R * k / K;
    for (int k = 0; k < K; k++) {
        if (!t->up_ev[k]) PDM_CUDA(cudaEventCreateWithFlags(&t->up_ev[k], cudaEventDisableTiming));
        PDM_CUDA(cudaMemcpyAsync(t->elev + edge[k] * C, host + (size_t)edge[k] * C * 8, (size_t)(edge[k + 1] - edge[k]) * C * 8,
                                 cudaMemcpyHostToDevice, t->copy_stream));
        PDM_CUDA(cudaEventRecord(t->up_ev[k], t->copy_stream));
    }
    t->have_elev = true; t->have_graph = false;
    // chunk k computes rows [edge[k] - (k > 0), edge[k + 1] - (k < K - 1)): its last row needs the row below it
    for (int k = 0; k < K; k++) {
        PDM_CUDA(cudaStreamWaitEvent(t->stream, t->up_ev[k], 0));
        const int64_t a = k > 0 ? edge[k] - 1 : 0, b = k < K - 1 ? edge[k + 1] - 1 : R;
        if (b > a) { int rc = pdm_launch_slopes_rows(t, a, b); if (rc) return rc; }
    }
    int rc = pdm_launch_flats(t);
    if (rc) return rc;
    t->have_slopes = true; t->have_flats = true;
    return PDM_OK;
}

int pdm_tile_find_flats(pdm_tile *t)
{
    if (!t) { pdm_set_error("NULL tile"); return PDM_ERR_ARG; }
    if (!t->have_slopes) { pdm_set_error("pdm_tile_find_flats: MAG not present"); return PDM_ERR_STATE; }
    int rc = pdm_launch_find_flats(t);
    if (rc) return rc;
    t->have_flats = true; t->have_graph = false;
    return PDM_OK;
}

static int read_counters(pdm_tile *t)
{
    PDM_CUDA(cudaMemcpyAsync(t->h_counters, t->d_counters, CT_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    return PDM_OK;
}


int pdm_tile_uca(pdm_tile *t, const pdm_uca_params *p_in, pdm_uca_stats *stats)
{
    if (!t) { pdm_set_error("NULL tile"); return PDM_ERR_ARG; }
    if (!t->have_elev || !t->have_spacing || !t->have_slopes || !t->have_flats) {
        pdm_set_error("pdm_tile_uca: needs ELEV, spacing, DIR/MAG and FLATS on the tile");
        return PDM_ERR_STATE;
    }
    if (!standalone(t)) { pdm_set_error("pdm_tile_uca: tile is a row shard; use the pdm_shard_* stages"); return PDM_ERR_STATE; }
    pdm_uca_params p;
    if (p_in) p = *p_in; else pdm_default_uca_params(&p);
    pdm_uca_stats st;
    memset(&st, 0, sizeof(st));
    st.n_cells = t->N;
    st.min_area = t->min_area;
    PDM_CUDA(cudaEventRecord(t->ev[0], t->stream));
    int rc = pdm_launch_graph(t, &p, &st);
    if (rc) return rc;
    PDM_CUDA(cudaEventRecord(t->ev[1], t->stream));
    rc = pdm_launch_sweep_full(t, &p, &st);
    if (rc) return rc;
    PDM_CUDA(cudaEventRecord(t->ev[2], t->stream));
    rc = read_counters(t);
    if (rc) return rc;
    if (!t->legacy_graph) {
        rc = pdm_ts_read_counters(t);
        if (rc) return rc;
    }
    if (t->h_counters[CT_WATCHDOG]) {
        pdm_set_error("pdm_tile_uca: work-list watchdog fired (QTAIL=%llu QHEAD=%llu QDONE=%llu PHASE1=%llu CHUNK=%llu DRAINED=%llu "
                      "SOURCES=%llu DEALT=%llu TAKEN=%llu EXITS=%llu)",
                      t->h_counters[CT_QTAIL], t->h_counters[CT_QHEAD], t->h_counters[CT_QDONE], t->h_counters[CT_PHASE1],
                      t->h_counters[CT_CHUNK], t->h_counters[CT_DRAINED], t->h_counters[CT_SOURCES], t->h_counters[CT_DBG_DEALT],
                      t->h_counters[CT_DBG_TAKEN], t->h_counters[CT_DBG_EXITS]);
        return PDM_ERR_STATE;
    }
    if (t->h_counters[CT_BADSEC]) {
        pdm_set_error("pdm_tile_uca: %llu cells have a direction outside the 8 sections (reference raises IndexError)",
                      t->h_counters[CT_BADSEC]);
        return PDM_ERR_SECTION;
    }
    st.n_sources = (int64_t)t->h_counters[CT_SOURCES];
    st.n_drained = (int64_t)t->h_counters[CT_DRAINED];
    st.n_undone = (int64_t)t->h_counters[CT_UNDONE];
    st.n_queue_items = (int64_t)t->h_counters[CT_QTAIL];
    st.n_edge_todo = (int64_t)t->h_counters[CT_EDGE_TODO];
    st.n_pits = t->n_pits;
    st.n_pit_edges = t->n_pit_edges;
    st.n_pits_undrained = (int64_t)t->h_counters[CT_PITS_UNDRAINED];
    st.ms_sweep_scan = (float)((double)(t->h_counters[CT_T_SCAN] - t->h_counters[CT_T_START]) * 1e-6);
    st.ms_sweep_kernel = (float)((double)(t->h_counters[CT_T_END] - t->h_counters[CT_T_START]) * 1e-6);
    if (!t->legacy_graph) {
        const unsigned long long *h = t->ts_hctr;
        st.n_sources = (int64_t)h[ts::TC_SOURCES];
        st.n_drained = (int64_t)h[ts::TC_CELLS];
        st.n_queue_items = (int64_t)h[ts::TC_VISITS];       // tile visits
        st.ms_sweep_scan = 0.0f;
        st.ms_sweep_kernel = (float)((double)(h[ts::TC_T_END] - h[ts::TC_T_START]) * 1e-6);
        if (getenv("PYDEM_B200_TS_DEBUG"))
            fprintf(stderr, "[ts] kernel %.3f ms | visits %llu (repeated in place %llu, deferred %llu) | cells %llu | levels %llu | sources %llu\n",
                    st.ms_sweep_kernel, h[ts::TC_VISITS], h[ts::TC_REQUEUE], h[ts::TC_DEFER], h[ts::TC_CELLS], h[ts::TC_LEVELS], h[ts::TC_SOURCES]);
        if (getenv("PYDEM_B200_TS_DEBUG") && atoi(getenv("PYDEM_B200_TS_DEBUG")) > 1) {
            fprintf(stderr, "[ts] CTA-time per phase, ms summed over CTAs: claim %.2f load %.2f count %.2f levels %.2f store %.2f schedule %.2f\n",
                    h[ts::TC_PHASE] * 1e-6, h[ts::TC_PHASE + 1] * 1e-6, h[ts::TC_PHASE + 2] * 1e-6, h[ts::TC_PHASE + 3] * 1e-6,
                    h[ts::TC_PHASE + 4] * 1e-6, h[ts::TC_PHASE + 5] * 1e-6);
            fprintf(stderr, "[ts] passes: %llu, %.0f cycles per pass\n", h[ts::TC_PHASE + 7],
                    h[ts::TC_PHASE + 7] ? (double)h[ts::TC_PHASE + 6] / (double)h[ts::TC_PHASE + 7] : 0.0);
            fprintf(stderr, "[ts] late visits: %llu visits, %llu cells; us per visit: claim %.2f load %.2f count %.2f flow %.2f busy turns of the busiest warp %.1f schedule+wait %.2f\n",
                    h[ts::TC_LATE + 6], h[ts::TC_LATE + 7], h[ts::TC_LATE] * 1e-3 / (double)(h[ts::TC_LATE + 6] + 1), h[ts::TC_LATE + 1] * 1e-3 / (double)(h[ts::TC_LATE + 6] + 1),
                    h[ts::TC_LATE + 2] * 1e-3 / (double)(h[ts::TC_LATE + 6] + 1), h[ts::TC_LATE + 3] * 1e-3 / (double)(h[ts::TC_LATE + 6] + 1),
                    h[ts::TC_LATE + 4] * 1.0 / (double)(h[ts::TC_LATE + 6] + 1), h[ts::TC_LATE + 5] * 1e-3 / (double)(h[ts::TC_LATE + 6] + 1));
            fprintf(stderr, "[ts] late visits, all warps' busy turns: cycles total %llu in drain step %llu hand-over %llu | busy turns summed over warps %llu\n",
                    h[ts::TC_DBG2], h[ts::TC_DBG2 + 1], h[ts::TC_DBG2 + 2], h[ts::TC_DBG2 + 3]);
            fprintf(stderr, "[ts] timeline (100 us buckets) visits/kcells:");
            for (int b = 0; b < 64; b++)
                if (h[ts::TC_HIST + 2 * b]) fprintf(stderr, " %d:%llu/%llu", b, h[ts::TC_HIST + 2 * b], h[ts::TC_HIST + 2 * b + 1] / 1000);
            fprintf(stderr, "\n");
        }
    }
    if (getenv("PYDEM_B200_WL_DEBUG")) {
        const unsigned long long *h = t->h_counters;
        fprintf(stderr, "[wl] kernel %.3f ms scan %.3f ms | scan-phase cells %llu | chains: calls %llu cells %llu avg %.0f ns/cell | "
                "team: lane-steps %llu avg %.0f ns | idle polls %llu | queue items %llu\n",
                st.ms_sweep_kernel, st.ms_sweep_scan, h[CT_X_SCAN_CELLS], h[CT_X_CHAIN_CALLS], h[CT_X_CHAIN_CELLS],
                h[CT_X_CHAIN_CELLS] ? (double)h[CT_X_CHAIN_NS] / (double)h[CT_X_CHAIN_CELLS] : 0.0, h[CT_X_TEAM_LANES],
                h[CT_X_TEAM_LANES] ? (double)h[CT_X_TEAM_NS] / (double)h[CT_X_TEAM_LANES] : 0.0, h[CT_X_POLLS], h[CT_QTAIL]);
        fprintf(stderr, "[wl] chased %llu (avg %.1f) | runs ended by the chase %llu, by in-degree 2: %llu, >= 3: %llu, <= 0: %llu\n", h[76],
                h[CT_X_BURSTS] ? (double)h[76] / (double)h[CT_X_BURSTS] : 0.0, h[CT_DBG_DEALT], h[77], h[78], h[79]);
        fprintf(stderr, "[wl] bursts %llu (failed %llu) cells %llu avg run %.1f | avg ns per burst %.0f of which chase %.0f\n", h[CT_X_BURSTS], h[CT_X_BURST_FAILS],
                h[CT_X_BURST_CELLS], h[CT_X_BURSTS] ? (double)h[CT_X_BURST_CELLS] / (double)h[CT_X_BURSTS] : 0.0,
                h[CT_X_BURSTS] ? (double)h[CT_X_BURST_NS] / (double)h[CT_X_BURSTS] : 0.0, h[CT_X_BURSTS] ? (double)h[CT_X_CHASE_NS] / (double)h[CT_X_BURSTS] : 0.0);
    }
    if (st.n_undone > 0) {
        // dem_processing.py:951-964: unreachable for a DAG unless circular_ref_maxcount <= 1 switched the sweep off
        rc = pdm_restart_rounds(t, &p, &st);
        if (rc) return rc;
    }
    PDM_CUDA(cudaEventRecord(t->ev[3], t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    cudaEventElapsedTime(&st.ms_graph, t->ev[0], t->ev[1]);
    cudaEventElapsedTime(&st.ms_sweep, t->ev[1], t->ev[2]);
    cudaEventElapsedTime(&st.ms_total, t->ev[0], t->ev[3]);
    t->have_graph = true; t->have_uca = true;
    if (stats) *stats = st;
    return PDM_OK;
}

int pdm_tile_uca_update(pdm_tile *t, const pdm_uca_params *p_in,
                        const double *data_left, const double *data_right, const double *data_top, const double *data_bottom,
                        const uint8_t *done_left, const uint8_t *done_right, const uint8_t *done_top, const uint8_t *done_bottom,
                        const uint8_t *todo_left, const uint8_t *todo_right, const uint8_t *todo_top, const uint8_t *todo_bottom,
                        pdm_uca_stats *stats)
{
    if (!t) { pdm_set_error("NULL tile"); return PDM_ERR_ARG; }
    if (!t->have_elev || !t->have_spacing || !t->have_slopes || !t->have_flats || !t->have_uca) {
        pdm_set_error("pdm_tile_uca_update: needs ELEV, spacing, DIR/MAG, FLATS and UCA (= uca_init) on the tile");
        return PDM_ERR_STATE;
    }
    if (!standalone(t)) { pdm_set_error("pdm_tile_uca_update: tile is a row shard"); return PDM_ERR_STATE; }
    pdm_uca_params p;
    if (p_in) p = *p_in; else pdm_default_uca_params(&p);
    pdm_uca_stats st;
    memset(&st, 0, sizeof(st));
    st.n_cells = t->N; st.min_area = t->min_area;
    const double *data[4] = {data_left, data_right, data_top, data_bottom};
    const uint8_t *done[4] = {done_left, done_right, done_top, done_bottom};
    const uint8_t *todo[4] = {todo_left, todo_right, todo_top, todo_bottom};
    for (int k = 0; k < 4; k++)
        if (!data[k] || !done[k] || !todo[k]) { pdm_set_error("pdm_tile_uca_update: NULL edge strip"); return PDM_ERR_ARG; }
    int rc = pdm_launch_update(t, &p, data, done, todo, &st);
    if (rc) return rc;
    if (stats) *stats = st;
    return PDM_OK;
}

int pdm_tile_pit_updates(pdm_tile *t, int64_t capacity, int32_t *cells, double *mag, uint8_t *flats, int64_t *n)
{
    if (!t || !n) { pdm_set_error("pdm_tile_pit_updates: NULL argument"); return PDM_ERR_ARG; }
    *n = t->n_pits;
    if (t->n_pits == 0) return PDM_OK;
    if (capacity < t->n_pits || !cells || !mag || !flats) { pdm_set_error("pdm_tile_pit_updates: need room for %lld pits", (long long)t->n_pits); return PDM_ERR_ARG; }
    return pdm_launch_pit_readback(t, cells, mag, flats);
}

int pdm_tile_twi(pdm_tile *t, const pdm_twi_params *p_in)
{
    if (!t) { pdm_set_error("NULL tile"); return PDM_ERR_ARG; }
    if (!t->twi10) PDM_CUDA(cudaMalloc(&t->twi10, (size_t)t->N * 8));
    if (!t->have_uca || !t->have_slopes) { pdm_set_error("pdm_tile_twi: needs UCA and MAG"); return PDM_ERR_STATE; }
    pdm_twi_params p;
    if (p_in) p = *p_in; else pdm_default_twi_params(&p);
    return pdm_launch_twi(t, &p);
}

// ---- one-shot host-buffer calls ----------------------------------------------------------

int pdm_slopes_directions(const double *elev, int64_t R, int64_t C, const double *dX, const double *dY,
                          const double *thA, const double *thB, double *mag, double *direction, uint8_t *flats)
{
    if (!elev || !dX || !dY || !mag || !direction) { pdm_set_error("pdm_slopes_directions: NULL argument"); return PDM_ERR_ARG; }
    pdm_tile *t = nullptr;
    int rc = pdm_tile_create(R, C, nullptr, &t);
    if (rc) return rc;
    std::vector<double> ones((size_t)R, 1.0);
    rc = pdm_tile_set_spacing(t, dX, dY, ones.data(), ones.data(), thA, thB);
    if (!rc) rc = pdm_tile_upload(t, PDM_F_ELEV, elev);
    if (!rc) rc = pdm_tile_slopes_directions(t);
    if (!rc) rc = pdm_tile_download(t, PDM_F_MAG, mag);
    if (!rc) rc = pdm_tile_download(t, PDM_F_DIR, direction);
    if (!rc && flats) rc = pdm_tile_download(t, PDM_F_FLATS, flats);
    pdm_tile_destroy(t);
    return rc;
}

int pdm_uca(const double *elev, const double *direction, double *mag, uint8_t *flats, int64_t R, int64_t C,
            const double *dX, const double *dY, const double *dX2, const double *dY2, const double *thA,
            const double *thB, const pdm_uca_params *p, double *uca, uint8_t *edge_todo, uint8_t *edge_done,
            pdm_uca_stats *stats)
{
    if (!elev || !direction || !mag || !flats || !dX || !dY || !dX2 || !dY2 || !uca) {
        pdm_set_error("pdm_uca: NULL argument");
        return PDM_ERR_ARG;
    }
    pdm_tile *t = nullptr;
    int rc = pdm_tile_create(R, C, nullptr, &t);
    if (rc) return rc;
    rc = pdm_tile_set_spacing(t, dX, dY, dX2, dY2, thA, thB);
    if (!rc) rc = pdm_tile_upload(t, PDM_F_ELEV, elev);
    if (!rc) rc = pdm_tile_upload(t, PDM_F_DIR, direction);
    if (!rc) rc = pdm_tile_upload(t, PDM_F_MAG, mag);
    if (!rc) rc = pdm_tile_upload(t, PDM_F_FLATS, flats);
    if (!rc) rc = pdm_tile_uca(t, p, stats);
    if (!rc) rc = pdm_tile_download(t, PDM_F_UCA, uca);
    if (!rc) rc = pdm_tile_download(t, PDM_F_MAG, mag);
    if (!rc) rc = pdm_tile_download(t, PDM_F_FLATS, flats);
    if (!rc && edge_todo) rc = pdm_tile_download(t, PDM_F_EDGE_TODO, edge_todo);
    if (!rc && edge_done) rc = pdm_tile_download(t, PDM_F_EDGE_DONE, edge_done);
    pdm_tile_destroy(t);
    return rc;
}

int pdm_uca_update(const double *elev, const double *direction, double *mag, uint8_t *flats, double *uca,
                   int64_t R, int64_t C, const double *dX, const double *dY, const double *dX2, const double *dY2,
                   const double *thA, const double *thB, const pdm_uca_params *p,
                   const double *data_left, const double *data_right, const double *data_top, const double *data_bottom,
                   const uint8_t *done_left, const uint8_t *done_right, const uint8_t *done_top, const uint8_t *done_bottom,
                   const uint8_t *todo_left, const uint8_t *todo_right, const uint8_t *todo_top, const uint8_t *todo_bottom,
                   uint8_t *edge_todo, uint8_t *edge_done)
{
    if (!elev || !direction || !mag || !flats || !uca || !dX || !dY || !dX2 || !dY2) {
        pdm_set_error("pdm_uca_update: NULL argument");
        return PDM_ERR_ARG;
    }
    pdm_tile *t = nullptr;
    int rc = pdm_tile_create(R, C, nullptr, &t);
    if (rc) return rc;
    rc = pdm_tile_set_spacing(t, dX, dY, dX2, dY2, thA, thB);
    if (!rc) rc = pdm_tile_upload(t, PDM_F_ELEV, elev);
    if (!rc) rc = pdm_tile_upload(t, PDM_F_DIR, direction);
    if (!rc) rc = pdm_tile_upload(t, PDM_F_MAG, mag);
    if (!rc) rc = pdm_tile_upload(t, PDM_F_FLATS, flats);
    if (!rc) rc = pdm_tile_upload(t, PDM_F_UCA, uca);
    if (!rc) rc = pdm_tile_uca_update(t, p, data_left, data_right, data_top, data_bottom, done_left, done_right,
                                      done_top, done_bottom, todo_left, todo_right, todo_top, todo_bottom, nullptr);
    if (!rc) rc = pdm_tile_download(t, PDM_F_UCA, uca);
    if (!rc) rc = pdm_tile_download(t, PDM_F_MAG, mag);
    if (!rc) rc = pdm_tile_download(t, PDM_F_FLATS, flats);
    if (!rc && edge_todo) rc = pdm_tile_download(t, PDM_F_EDGE_TODO, edge_todo);
    if (!rc && edge_done) rc = pdm_tile_download(t, PDM_F_EDGE_DONE, edge_done);
    pdm_tile_destroy(t);
    return rc;
}

int pdm_twi(const double *uca, const double *mag, int64_t n, const pdm_twi_params *p_in, double *twi)
{
    if (!uca || !mag || !twi || n <= 0) { pdm_set_error("pdm_twi: bad argument"); return PDM_ERR_ARG; }
    pdm_twi_params p;
    if (p_in) p = *p_in; else pdm_default_twi_params(&p);
    double *d_u = nullptr, *d_m = nullptr, *d_t = nullptr;
    int rc = PDM_OK;
    cudaError_t e;
    if ((e = cudaMalloc(&d_u, (size_t)n * 8)) != cudaSuccess || (e = cudaMalloc(&d_m, (size_t)n * 8)) != cudaSuccess ||
        (e = cudaMalloc(&d_t, (size_t)n * 8)) != cudaSuccess) {
        rc = pdm_cuda_fail(e, "cudaMalloc(twi)", __FILE__, __LINE__);
    } else {
        pdm_tile fake;
        memset(&fake, 0, sizeof(fake));
        fake.N = n; fake.uca = d_u; fake.mag = d_m; fake.twi = d_t; fake.stream = nullptr;
        e = cudaMemcpy(d_u, uca, (size_t)n * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_m, mag, (size_t)n * 8, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) rc = pdm_cuda_fail(e, "cudaMemcpy(twi in)", __FILE__, __LINE__);
        if (!rc) rc = pdm_launch_twi(&fake, &p);
        if (!rc) {
            e = cudaMemcpy(twi, d_t, (size_t)n * 8, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) rc = pdm_cuda_fail(e, "cudaMemcpy(twi out)", __FILE__, __LINE__);
        }
    }
    cudaFree(d_u); cudaFree(d_m); cudaFree(d_t);
    return rc;
}

}  // extern "C"
