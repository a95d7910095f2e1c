// comm.cu -- the row-sharded hot path driven entirely through the C ABI: NCCL for the halo rows (pdm_comm_*),
// CUDA-IPC peer memory for the accumulation sweep, and pdm_shard_run = the whole pass
//
//     elev halo -> slope/aspect (a1) -> flat0 halo -> region labels (+ label rounds) -> flats (a2) -> links (a3/a4)
//     -> [pit drains with the neighbours' strips (a5)] -> link halo -> in-degrees -> ONE sweep across the GPUs (a6/a7)
//     -> finalize -> TWI (a9)
//
// which is what pydem_b200/sharded.py does with torch.distributed; a C / Fortran / ctypes-only host needs no torch.
// Semantics of pyDEM's cross-tile edge resolution (process_manager.py:1090-1249) in its gating-free form; the
// sharded result equals the single-tile result (shard.cu).
//
// NCCL is loaded at run time (dlopen "libnccl.so.2": the system library, or the copy another component of the
// process -- e.g. torch -- has already loaded), so the library has no link-time dependency on it and single-GPU
// users never touch it.  Only types come from <nccl.h>.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "pdm_internal.cuh"
#include "tsweep.cuh"

namespace {

struct Nccl {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclComm_t comm;
    int rank, world;
};
Nccl g_nccl;

int nccl_load()
{
    if (g_nccl.lib) return PDM_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { pdm_set_error("pdm_comm: cannot load libnccl.so.2 (%s)", dlerror()); return PDM_ERR_STATE; }
#define SYM(field, name) \
    *(void **)(&g_nccl.field) = dlsym(h, name); \
    if (!g_nccl.field) { pdm_set_error("pdm_comm: %s not found in libnccl", name); dlclose(h); return PDM_ERR_STATE; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
    SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = h;
    return PDM_OK;
}

#define PDM_NCCL(call)                                                                                       \
    do {                                                                                                     \
        ncclResult_t r__ = (call);                                                                           \
        if (r__ != ncclSuccess) { pdm_set_error("NCCL: %s failed: %s", #call, g_nccl.GetErrorString(r__)); return PDM_ERR_CUDA; } \
    } while (0)

int need_comm()
{
    if (!g_nccl.comm) { pdm_set_error("pdm_comm_init has not been called"); return PDM_ERR_STATE; }
    return PDM_OK;
}

// rows travel between row neighbours: my send_up lands in the recv_down of the rank above, my send_down in the
// recv_up of the rank below (one NCCL group per exchange)
int exchange(pdm_tile *t, const void *send_up, void *recv_up, const void *send_dn, void *recv_dn, size_t bytes)
{
    const int r = g_nccl.rank, w = g_nccl.world;
    if (w == 1 || bytes == 0) return PDM_OK;
    PDM_NCCL(g_nccl.GroupStart());
    if (r > 0) {
        PDM_NCCL(g_nccl.Send(send_up, bytes, ncclChar, r - 1, g_nccl.comm, t->stream));
        PDM_NCCL(g_nccl.Recv(recv_up, bytes, ncclChar, r - 1, g_nccl.comm, t->stream));
    }
    if (r < w - 1) {
        PDM_NCCL(g_nccl.Send(send_dn, bytes, ncclChar, r + 1, g_nccl.comm, t->stream));
        PDM_NCCL(g_nccl.Recv(recv_dn, bytes, ncclChar, r + 1, g_nccl.comm, t->stream));
    }
    PDM_NCCL(g_nccl.GroupEnd());
    return PDM_OK;
}

// one halo exchange of a field with `esz` bytes per cell: first / last owned row out, halo rows in
int halo(pdm_tile *t, void *field, size_t esz)
{
    const Win &w = t->win;
    char *f = reinterpret_cast<char *>(field);
    const size_t row = (size_t)w.C * esz;
    return exchange(t, f + (size_t)w.lo * row, f + (size_t)(w.lo - 1) * row, f + (size_t)(w.hi - 1) * row, f + (size_t)w.hi * row, row);
}

// device scratch of pdm_shard_run, kept with the tile
struct Scratch {
    long long *lab_l[4];     // packed region labels: send_up, send_down, recv_up, recv_down
    double *lab_e[4];
    long long *flag;         // [2]: changed regions (mine, all ranks)
    int64_t H, Hin;
    double *E[2]; uint8_t *P[2]; int32_t *in[2], *from[2];   // pit strips: above / below
};

int scratch_get(pdm_tile *t, Scratch **out)
{
    if (!t->comm_scratch) {
        Scratch *s = new Scratch();
        memset(s, 0, sizeof(*s));
        for (int k = 0; k < 4; k++) {
            PDM_CUDA(cudaMalloc(&s->lab_l[k], (size_t)t->C * 8));
            PDM_CUDA(cudaMalloc(&s->lab_e[k], (size_t)t->C * 8));
        }
        PDM_CUDA(cudaMalloc(&s->flag, 16));
        t->comm_scratch = s;
    }
    *out = reinterpret_cast<Scratch *>(t->comm_scratch);
    return PDM_OK;
}

int scratch_pits(pdm_tile *t, Scratch *s, int64_t H, int64_t Hin)
{
    if (s->H == H && s->Hin == Hin) return PDM_OK;
    for (int k = 0; k < 2; k++) {
        if (s->E[k]) { cudaFree(s->E[k]); cudaFree(s->P[k]); cudaFree(s->in[k]); cudaFree(s->from[k]); }
        s->E[k] = nullptr; s->P[k] = nullptr; s->in[k] = nullptr; s->from[k] = nullptr;
    }
    const bool has[2] = {t->win.lo > 0, t->win.hi < t->R};
    for (int k = 0; k < 2; k++) {
        if (!has[k]) continue;
        PDM_CUDA(cudaMalloc(&s->E[k], (size_t)H * t->C * 8));
        PDM_CUDA(cudaMalloc(&s->P[k], (size_t)H * t->C));
        PDM_CUDA(cudaMalloc(&s->in[k], (size_t)(Hin > 0 ? Hin : 1) * t->C * 4));
        PDM_CUDA(cudaMalloc(&s->from[k], (size_t)(Hin > 0 ? Hin : 1) * t->C * 4));
    }
    s->H = H; s->Hin = Hin;
    return PDM_OK;
}

}  // namespace

void pdm_comm_scratch_free(pdm_tile *t)
{
    Scratch *s = reinterpret_cast<Scratch *>(t->comm_scratch);
    if (!s) return;
    for (int k = 0; k < 4; k++) { if (s->lab_l[k]) cudaFree(s->lab_l[k]); if (s->lab_e[k]) cudaFree(s->lab_e[k]); }
    if (s->flag) cudaFree(s->flag);
    for (int k = 0; k < 2; k++) {
        if (s->E[k]) cudaFree(s->E[k]);
        if (s->P[k]) cudaFree(s->P[k]);
        if (s->in[k]) cudaFree(s->in[k]);
        if (s->from[k]) cudaFree(s->from[k]);
    }
    delete s;
    t->comm_scratch = nullptr;
}

extern "C" {

// rank 0: a fresh NCCL unique id (PDM_COMM_ID_BYTES = 128 bytes) to hand to every rank by any means
int pdm_comm_unique_id(void *id_out)
{
    if (!id_out) { pdm_set_error("pdm_comm_unique_id: NULL argument"); return PDM_ERR_ARG; }
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    PDM_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "PDM_COMM_ID_BYTES");
    memcpy(id_out, &id, sizeof(id));
    return PDM_OK;
}

// one process per GPU (pdm_init first): join the communicator of `nranks` row shards; rank r holds the r-th row block
int pdm_comm_init(int rank, int nranks, const void *id)
{
    if (!id || nranks < 1 || rank < 0 || rank >= nranks || nranks > PDM_MAX_WORLD) { pdm_set_error("pdm_comm_init: bad argument"); return PDM_ERR_ARG; }
    if (g_nccl.comm) { pdm_set_error("pdm_comm_init: already initialised"); return PDM_ERR_STATE; }
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    PDM_NCCL(g_nccl.CommInitRank(&g_nccl.comm, nranks, uid, rank));
    g_nccl.rank = rank; g_nccl.world = nranks;
    return PDM_OK;
}

int pdm_comm_finalize(void)
{
    if (g_nccl.comm) { g_nccl.CommDestroy(g_nccl.comm); g_nccl.comm = nullptr; }
    return PDM_OK;
}

// barrier of all ranks on the tile's stream (returns when every rank has reached it)
int pdm_comm_barrier(pdm_tile *t)
{
    int rc = need_comm();
    if (rc) return rc;
    Scratch *s = nullptr;
    if ((rc = scratch_get(t, &s))) return rc;
    PDM_CUDA(cudaMemsetAsync(s->flag, 0, 16, t->stream));
    PDM_NCCL(g_nccl.AllReduce(s->flag, s->flag + 1, 1, ncclInt64, ncclSum, g_nccl.comm, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    return PDM_OK;
}

// collective, once after pdm_tile_set_window: all-gather the ranks' CUDA-IPC exports over NCCL and map the
// neighbours' records / every rank's control block (pdm_shard_p2p_connect_all)
int pdm_shard_connect(pdm_tile *t)
{
    int rc = need_comm();
    if (rc) return rc;
    if (!t) { pdm_set_error("NULL tile"); return PDM_ERR_ARG; }
    int64_t size = 0;
    if ((rc = pdm_shard_p2p_export(t, nullptr, &size))) return rc;
    std::string mine((size_t)size, '\0'), all((size_t)size * g_nccl.world, '\0');
    if ((rc = pdm_shard_p2p_export(t, &mine[0], &size))) return rc;
    char *d_mine = nullptr, *d_all = nullptr;
    PDM_CUDA(cudaMalloc(&d_mine, (size_t)size));
    PDM_CUDA(cudaMalloc(&d_all, (size_t)size * g_nccl.world));
    PDM_CUDA(cudaMemcpyAsync(d_mine, mine.data(), (size_t)size, cudaMemcpyHostToDevice, t->stream));
    PDM_NCCL(g_nccl.AllGather(d_mine, d_all, (size_t)size, ncclChar, g_nccl.comm, t->stream));
    PDM_CUDA(cudaMemcpyAsync(&all[0], d_all, (size_t)size * g_nccl.world, cudaMemcpyDeviceToHost, t->stream));
    PDM_CUDA(cudaStreamSynchronize(t->stream));
    cudaFree(d_mine); cudaFree(d_all);
    if ((rc = pdm_shard_p2p_connect_all(t, all.data(), g_nccl.world, g_nccl.rank))) return rc;
    return pdm_comm_barrier(t);
}

// collective: unmap the peers' memory on every rank, then a barrier; after it tiles may be destroyed
int pdm_shard_disconnect(pdm_tile *t)
{
    int rc = pdm_shard_p2p_disconnect(t);
    if (rc) return rc;
    return g_nccl.comm ? pdm_comm_barrier(t) : PDM_OK;
}

// collective: one pass of the hot path over the row shards (ELEV of the owned rows resident on every rank;
// pdm_shard_connect done).  twi NULL: stop after UCA.  Out: MAG, DIR, FLATS, UCA, EDGE_TODO, EDGE_DONE, TWI of the
// owned rows on every rank.  *label_rounds (may be NULL): exchange rounds the cross-rank flat regions took.
int pdm_shard_run(pdm_tile *t, const pdm_uca_params *p_in, const pdm_twi_params *twi, pdm_uca_stats *stats, int *label_rounds)
{
    int rc = need_comm();
    if (rc) return rc;
    if (!t) { pdm_set_error("NULL tile"); return PDM_ERR_ARG; }
    if (g_nccl.world > 1 && !pdm_shard_worklist_p2p(t)) { pdm_set_error("pdm_shard_run: pdm_shard_connect first"); return PDM_ERR_STATE; }
    pdm_uca_params p;
    if (p_in) p = *p_in; else pdm_default_uca_params(&p);
    const Win &w = t->win;
    const int64_t C = t->C;
    Scratch *s = nullptr;
    if ((rc = scratch_get(t, &s))) return rc;
    const bool up = w.lo > 0, dn = w.hi < t->R;
    // a1
    if ((rc = halo(t, t->elev, 8)) || (rc = pdm_shard_slopes(t))) return rc;
    // a2: regions across ranks agree on their global minimum label
    if ((rc = halo(t, t->flat0, 1)) || (rc = pdm_shard_ccl(t))) return rc;
    int rounds = 0;
    while (g_nccl.world > 1) {
        if (up && (rc = pdm_shard_label_pack(t, w.lo, s->lab_l[0], s->lab_e[0]))) return rc;
        if (dn && (rc = pdm_shard_label_pack(t, w.hi - 1, s->lab_l[1], s->lab_e[1]))) return rc;
        if ((rc = exchange(t, s->lab_l[0], s->lab_l[2], s->lab_l[1], s->lab_l[3], (size_t)C * 8))) return rc;
        if ((rc = exchange(t, s->lab_e[0], s->lab_e[2], s->lab_e[1], s->lab_e[3], (size_t)C * 8))) return rc;
        PDM_CUDA(cudaMemsetAsync(s->flag, 0, 16, t->stream));
        if (up && (rc = pdm_shard_label_unpack(t, w.lo - 1, s->lab_l[2], s->lab_e[2], s->flag))) return rc;
        if (dn && (rc = pdm_shard_label_unpack(t, w.hi, s->lab_l[3], s->lab_e[3], s->flag))) return rc;
        PDM_NCCL(g_nccl.AllReduce(s->flag, s->flag + 1, 1, ncclInt64, ncclSum, g_nccl.comm, t->stream));
        long long changed = 0;
        PDM_CUDA(cudaMemcpyAsync(&changed, s->flag + 1, 8, cudaMemcpyDeviceToHost, t->stream));
        PDM_CUDA(cudaStreamSynchronize(t->stream));
        rounds++;
        if (changed == 0) break;
    }
    if (label_rounds) *label_rounds = rounds;
    if ((rc = pdm_shard_flats_extend(t))) return rc;
    // a3/a4 (+ a5 across the shard boundaries)
    if ((rc = pdm_shard_links(t, &p))) return rc;
    if (p.drain_pits && g_nccl.world > 1) {
        const int64_t H = p.drain_pits_max_iter + 1;
        const int64_t Hin = p.drain_pits_max_dist > 0 ? (p.drain_pits_max_dist < H ? p.drain_pits_max_dist : H) : H;
        if (w.hi - w.lo < H) { pdm_set_error("drain_pits on row shards needs at least drain_pits_max_iter + 1 = %lld rows per rank", (long long)H); return PDM_ERR_ARG; }
        if ((rc = scratch_pits(t, s, H, Hin))) return rc;
        if ((rc = exchange(t, t->elev + w.lo * C, s->E[0], t->elev + (w.hi - H) * C, s->E[1], (size_t)H * C * 8))) return rc;
        if ((rc = exchange(t, t->flat0 + w.lo * C, s->P[0], t->flat0 + (w.hi - H) * C, s->P[1], (size_t)H * C))) return rc;
        if ((rc = pdm_shard_pits(t, &p, s->E[0], s->P[0], up ? H : 0, s->E[1], s->P[1], dn ? H : 0, s->in[0], s->in[1], Hin))) return rc;
        if ((rc = exchange(t, s->in[0], s->from[0], s->in[1], s->from[1], (size_t)Hin * C * 4))) return rc;
        if ((rc = pdm_shard_pit_in_apply(t, s->from[0], s->from[1], Hin))) return rc;
    }
    if ((rc = halo(t, t->link, 1)) || (rc = pdm_shard_indeg(t))) return rc;
    // a6/a7: one sweep across the GPUs, then the epilogue; a9
    if ((rc = pdm_shard_sweep(t, 1)) || (rc = pdm_shard_finalize(t, &p, stats))) return rc;
    if (twi && (rc = pdm_tile_twi(t, twi))) return rc;
    return PDM_OK;
}

}  // extern "C"
