// pdm_internal.cuh -- shared declarations of libpydem_b200 (not part of the public ABI).
//
// Data layout in HBM (per tile, N = R*C cells, C-order, 32-bit cell indices):
//   elev, mag, dir, uca, twi : f64[N]
//   flats, flat0, link, edge_todo, edge_done : u8[N]
//   cell : Cell[N] -- the sweep state of a cell packed into ONE 32-byte sector (area, taint,
//          proportion, live in-degree, link byte): the sweep visits cells in flow order, i.e. at
//          random, and every visit then costs one DRAM sector instead of five
//   label (i32[N], flat-region union-find, later pit in-degree scratch / seed list)
//   queue (i32[N], deferred work items of the sweep)
//   per-row geometry: dX,dY,dg,thA,thB : f64[R-1] (fences);  th_row,row_area : f64[R]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/pydem_b200.h"

#define PDM_PI 3.141592653589793
#define PDM_MAX_WORLD 16   // ranks of one multi-GPU sweep (one box)

// link byte written by the graph kernel (reference: section of
// _calc_uca_section_proportion + the keep filter of _mk_adjacency_matrix 1136-1137)
#define LK_SEC_MASK 0x07   // facet index 0..7 of the two receivers
#define LK_KEEP1 0x08      // edge to the cardinal receiver e1 survives the filter
#define LK_KEEP2 0x10      // edge to the diagonal receiver e2 survives the filter
#define LK_PIT 0x20        // cell drains through a pit edge list (prop holds the slot)
#define LK_NOSEC 0x40      // section outside 0..7 (flat / undefined): no receivers
#define LK_SOURCE 0x80     // legacy work-list sweep: nobody drains into this cell (initial frontier)
#define LK_PITIN 0x80      // tile sweep: the cell receives pit edges (same bit: the two sweeps never share a graph)

// tile sweep: UCA / taint hold this signalling-NaN pattern until the cell's sum is final (no
// arithmetic produces it: operations quiet a signalling NaN)
#define TS_NOT_DONE 0x7ff4dead7ff4deadULL

// Row window of a tile inside a (possibly larger, row-sharded) grid.  A stand-alone tile has
// row_off = 0, Rg = R, lo = 0, hi = R.  A shard keeps halo rows around its owned rows [lo, hi);
// "border" always means the border of the GLOBAL grid (rows 0 / Rg-1, columns 0 / C-1).
struct Win {
    int64_t R, C;        // local rows (owned + halo), columns
    int64_t row_off;     // global row index of local row 0
    int64_t Rg;          // rows of the global grid
    int64_t lo, hi;      // owned local rows
    __host__ __device__ __forceinline__ bool top(int64_t li) const { return li + row_off == 0; }
    __host__ __device__ __forceinline__ bool bottom(int64_t li) const { return li + row_off == Rg - 1; }
    // is local row li (possibly one step outside [0,R)) inside the global grid?
    __host__ __device__ __forceinline__ bool row_in_grid(int64_t li) const { return li + row_off >= 0 && li + row_off < Rg; }
};

// sweep state of one cell: exactly one 32-byte DRAM sector
struct __align__(32) Cell {
    double area;      // accumulated upstream area (cyutils.pyx:161); update mode: the delta
    double taint;     // accumulated edge_todo weight (cyutils.pyx:163)
    double prop;      // share of the cardinal receiver; for pit cells: slot of the pit edge list
    int32_t indeg;    // receivers still waiting for this many sources; halo rows: -(decrements)
    uint8_t link;     // LK_* byte (copy of link[])
    uint8_t pad[3];
};

struct pdm_tile {
    int64_t R, C, N;
    Win win;
    int device;
    cudaStream_t stream;
    // fields
    double *elev, *mag, *dir, *uca, *twi;
    double *twi10;       // 10 * twi (what DEMProcessor.twi stores), allocated on first download request
    Cell *cell;
    uint8_t *flats, *flat0, *link, *edge_todo, *edge_done;
    int8_t *section;  // only materialised on download of PDM_F_SECTION
    int32_t *label, *queue;
    long long *glabel;   // sharded mode: global minimum cell index of each flat region (at its local root)
    double *glelev;      //               elevation of that cell
    // geometry
    double *dX, *dY, *dg, *thA, *thB, *th_row, *row_area;
    double *rdX, *rdY, *rdg;   // correctly rounded reciprocals of dX, dY, dg (fast stencil divisions)
    double *dXg, *dYg;         // row shard: fences of the whole grid (pit drains across the shard boundary)
    double min_area;
    bool have_spacing, have_elev, have_slopes, have_flats, have_graph, have_uca;
    bool stencil_parity; // run the literal (slow) stencil formulation on this tile (tests)
    bool queue_ready;    // queue slots are all -1 except those the last work-list run used
    const unsigned long long *queue_dirty_ctr;   // the counters whose CT_QTAIL counts those slots (d_counters, or the shared block of a multi-GPU sweep)
    bool keep_graph;     // update mode reuses the graph of the last full sweep (device-resident mosaic tiles)
    // pit edge lists (device)
    int32_t *pit_cell;     // [pit_cap] cells examined by the pit search (flats & elev > 0)
    int32_t *pit_beg;      // [pit_cap] first / one-past-last edge of each pit in pit_dst/pit_w
    int32_t *pit_end;
    int32_t *pit_dst;      // [pit_edge_cap] receiving cell of each pit edge
    double *pit_w;         // [pit_edge_cap] share of the pit's area
    int32_t *pit_scratch_i;  // per-block border double buffers of the pit search
    double *pit_scratch_d;
    int64_t n_pits, n_pit_edges, pit_cap, pit_edge_cap, pit_scratch_blocks;
    // staging of the update mode's edge strips: [left R][right R][top C][bottom C]
    double *edge_buf_d;
    uint8_t *edge_buf_b;
    // small device scratch for counters (pinned host mirror for readback)
    unsigned long long *d_counters;  // [32]
    unsigned long long *h_counters;  // pinned
    cudaEvent_t ev[4];
    cudaStream_t copy_stream;   // device->host copies that overlap the next stage (pdm_tile_download_async)
    cudaEvent_t copy_ev;
    cudaEvent_t up_ev[16];      // chunked upload of ELEV overlapping the stencil (pdm_tile_upload_slopes_directions)
    // tile sweep (tsweep.cu): ticket queue, per-tile state words, counters (+ pinned mirror), the
    // "already seen" bytes of the two halo rows of a shard
    void *ts_ctl;               // one allocation: [counters | tile flags | queue slots] (shared with peer GPUs through CUDA IPC)
    int32_t *ts_slots;
    uint32_t *ts_flag;
    unsigned long long *ts_ctr, *ts_hctr;
    uint8_t *ts_seen;
    int64_t ts_cap, ts_ntiles_cap;
    // one sweep across GPUs (shard.cu, pdm_shard_p2p_*): mapped peer memory of the row neighbours and of rank 0
    struct P2P {
        int on, world, rank;
        unsigned long long launches;       // start barrier target = world * launches
        void *map_rec[2], *map_ctl[2], *map_root;     // what cudaIpcOpenMemHandle returned (to close)
        void *map_all[PDM_MAX_WORLD];                 // control blocks of the ranks that are neither neighbours nor rank 0
        void *all_ctl[PDM_MAX_WORLD];                 // every rank's control block (own included); all set = the work-list sweep may span the GPUs
        const void *rec[2]; void *ctl[2]; void *root_ctl;
        long long off_flag[2], off_slots[2], lo[2], hi[2];
        int nty[2];
        unsigned cap_mask[2];
    } p2p;
    void *comm_scratch;         // device buffers of pdm_shard_run (comm.cu)
    bool legacy_graph;          // the graph on the tile was built for the legacy work-list sweep
    bool shard_pits_wanted, shard_pits_done;   // row shard with drain_pits: pdm_shard_pits must run between pdm_shard_links and pdm_shard_indeg
};

// sweep state of one cell in the tile sweep (tsweep.cu): exactly one 32-byte DRAM sector, in the
// same array the legacy sweep and the update mode use for their Cell records
struct __align__(32) TRec {
    double area;      // accumulated upstream area (cyutils.pyx:161); TS_NOT_DONE until final
    double taint;     // accumulated edge_todo weight (cyutils.pyx:163); TS_NOT_DONE until final
    double prop;      // share of the cardinal receiver; for pit cells: slot of the pit edge list
    uint8_t link;     // LK_* byte
    uint8_t dmask;    // which of the 8 neighbours (W E N S NW NE SW SE) drain into this cell
    uint8_t flags;    // TR_*
    uint8_t pad[5];
};
#define TR_TODO 0x01   // inflow-border cell: its taint starts at 1 (dem_processing.py:909-944)
static inline TRec *pdm_trec(const pdm_tile *t) { return reinterpret_cast<TRec *>(t->cell); }

void pdm_set_error(const char *fmt, ...);
int pdm_cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define PDM_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e__ = (call);                                             \
        if (e__ != cudaSuccess) return pdm_cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// every kernel launch goes through this: error check + launch accounting (pdm_launch_count)
extern unsigned long long g_pdm_launches;
#define PDM_LAUNCHED()                       \
    do {                                     \
        PDM_CUDA(cudaGetLastError());        \
        g_pdm_launches++;                    \
    } while (0)

// kernels' launchers (each returns a pdm_status)
int pdm_launch_geometry(pdm_tile *t);
int pdm_launch_slopes(pdm_tile *t);
int pdm_launch_slopes_rows(pdm_tile *t, int64_t row_lo, int64_t row_hi);
int pdm_launch_flats(pdm_tile *t);
int pdm_launch_find_flats(pdm_tile *t);
int pdm_launch_ccl(pdm_tile *t);   // union-find labels of the mask in flat0 (label[] must hold own indices)
int pdm_launch_graph(pdm_tile *t, const pdm_uca_params *p, pdm_uca_stats *st);
int pdm_launch_sweep_full(pdm_tile *t, const pdm_uca_params *p, pdm_uca_stats *st);
int pdm_launch_update(pdm_tile *t, const pdm_uca_params *p, const double *const data[4],
                      const uint8_t *const done[4], const uint8_t *const todo[4], pdm_uca_stats *st);
int pdm_launch_twi(pdm_tile *t, const pdm_twi_params *p);
int pdm_launch_section_export(pdm_tile *t);
int pdm_launch_pit_readback(pdm_tile *t, int32_t *cells, double *mag, uint8_t *flats);
int pdm_launch_selftest_div(unsigned long long seed, int blocks, long long per_thread, unsigned long long *mismatch_host);
int pdm_restart_rounds(pdm_tile *t, const pdm_uca_params *p, pdm_uca_stats *st);
int pdm_graph_links_pits(pdm_tile *t, const pdm_uca_params *p);
// row shard: what the pit search sees of the neighbouring ranks (pdm_shard_pits, shard.cu)
struct pdm_pit_shard {
    const double *E_up, *E_dn;      // elevation of the Hu rows above / Hd rows below the owned rows (device)
    const uint8_t *P_up, *P_dn;     // their pit mask
    int64_t Hu, Hd;
    int32_t *in_up, *in_dn;         // out: pit edges arriving at the neighbours' cells, [Hin][C] each
    int64_t Hin;
    int64_t peer_row[2];            // neighbour's local row = my local row + peer_row[side]
    const double *dXg, *dYg;        // fences of the whole grid (device)
};
int pdm_launch_pits(pdm_tile *t, const pdm_uca_params *p, const pdm_pit_shard *sh = nullptr);
// tile sweep (tsweep.cu)
int pdm_ts_reset_state(pdm_tile *t);
int pdm_launch_tsweep(pdm_tile *t, int first);
int pdm_ts_read_counters(pdm_tile *t);
int pdm_launch_ts_finalize(pdm_tile *t, const pdm_uca_params *p);
void pdm_ts_p2p_close(pdm_tile *t);
int pdm_launch_border_todo(pdm_tile *t);
int pdm_launch_indeg_todo(pdm_tile *t);
bool pdm_sweep_legacy();
void pdm_comm_scratch_free(pdm_tile *t);
bool pdm_shard_worklist_p2p(const pdm_tile *t);
int pdm_launch_sweep_p2p(pdm_tile *t);
int pdm_launch_sweep_first(pdm_tile *t);
int pdm_launch_uca_finalize(pdm_tile *t, const pdm_uca_params *p);

// counters slots.  The queue counters are hammered by every warp (fetch-and-add tickets,
// pushes, completion counts, termination polls): each lives on its own 128-byte line so the
// L2 serialises them independently.
enum {
    CT_QTAIL = 0,      // sweep queue: items produced
    CT_QHEAD = 16,     // tickets handed out
    CT_QDONE = 32,     // chains (seed- or queue-born) that have completely ended
    CT_PHASE1 = 48,    // (unused)
    CT_DRAINED = 64,   // cells drained
    CT_CHUNK = 72,     // next 32-entry chunk of the seed scan

    CT_SOURCES = 80,
    CT_UNDONE = 81,
    CT_BADSEC = 82,
    CT_NPITS = 83,
    CT_NPITEDGES = 96,
    CT_PITS_UNDRAINED = 97,
    CT_EDGE_TODO = 84,
    CT_FLAG = 98,
    CT_TMP0 = 99,
    CT_TMP1 = 100,
    CT_ABORT = 101,
    CT_DBG_DEALT = 73,  // diagnostics: seeds dealt / queue items consumed / warps that left through the termination test
    CT_DBG_TAKEN = 74,
    CT_DBG_EXITS = 75,
    CT_WATCHDOG = 104, // set by a warp that saw no progress for seconds; sticky until the next graph build
    CT_T_START = 112,  // work-list timing (globaltimer ns): first warp in
    CT_T_SCAN = 113,   // last warp finished its seed scan (+ scan-born chains)
    CT_T_END = 114,    // last warp out
    CT_X_FIRST = 116,  // diagnostics of a work-list run (PYDEM_B200_WL_DEBUG)
    CT_X_CHAIN_CALLS = 116, CT_X_CHAIN_CELLS = 117, CT_X_CHAIN_NS = 118, CT_X_TEAM_LANES = 119, CT_X_TEAM_NS = 120,
    CT_X_POLLS = 121, CT_X_SCAN_CELLS = 122,
    CT_X_BURSTS = 123, CT_X_BURST_CELLS = 124, CT_X_BURST_FAILS = 125, CT_X_CHASE_NS = 126, CT_X_BURST_NS = 127,   // chain bursts (drain_op.cuh)
    CT_X_LAST = 127,
    // one work-list sweep across the row shards of several GPUs (drain_op.cuh MODE 3): the counters then live in the
    // control block the peers have mapped (tsweep.cuh, TC_WLC)
    CT_INBOX_TAIL = 128,  // cells made ready by a peer GPU (it pushes them into this rank's in-box)
    CT_INBOX_HEAD = 144,  // in-box tickets handed out
    CT_GTERM = 160,       // number of the last multi-GPU sweep known to be over everywhere (monotonic, never reset)
    CT_N = 176
};

