/*
 * pydem_b200.h -- C ABI of the B200-native pyDEM hot path (libpydem_b200.so).
 *
 * Drop-in boundary for pyDEM's per-tile engine: the D-infinity slope/aspect stencil, the
 * upstream-contributing-area (UCA) sweep and the TWI map, i.e. what
 * pydem.dem_processing.DEMProcessor.calc_slopes_directions / calc_uca / calc_twi compute
 * (reference dem_processing.py:587, 682, 1647) and what its only native seam,
 * pydem/cyfuncs/cyutils.pyx (drain_area :78, drain_connections :35), accelerates.
 *
 * Plain C: pointers and sizes only, no torch / numpy types.  All grids are C-order
 * [R][C] (row = latitude, row 0 = north), float64 unless stated, flat index i*C+j.
 * Every function returns 0 on success and a non-zero pdm_status otherwise; the message
 * is available from pdm_last_error() (thread-local).  There is NO CPU fallback: without
 * a usable CUDA device every compute entry point fails with PDM_ERR_CUDA.
 *
 * Two layers:
 *   1. tile handles (pdm_tile_*): fields live in HBM across calls, so the three stages
 *      chain without host round trips and the graph is built once per tile;
 *   2. one-shot host-buffer calls (pdm_slopes_directions, pdm_uca, pdm_uca_update,
 *      pdm_twi): upload -> compute -> download, the exact shape of the reference calls.
 */
#ifndef PYDEM_B200_H
#define PYDEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDM_ABI_VERSION 2

typedef enum {
    PDM_OK = 0,
    PDM_ERR_ARG = 1,      /* bad argument (NULL, shape < 3x3, unknown field ...) */
    PDM_ERR_CUDA = 2,     /* CUDA runtime / no device */
    PDM_ERR_STATE = 3,    /* stage called before its inputs exist on the tile */
    PDM_ERR_SECTION = 4,  /* direction outside [-pi/2, 2pi]: the reference raises IndexError
                             (dem_processing.py:1068) */
    PDM_ERR_NOMEM = 5
} pdm_status;

/* Fields of a tile (device-resident).  dtype in brackets. */
typedef enum {
    PDM_F_ELEV = 0,      /* [f64]  DEMProcessor.elev */
    PDM_F_MAG = 1,       /* [f64]  .mag      (-1 on flats) */
    PDM_F_DIR = 2,       /* [f64]  .direction (radians CCW from east, -1 on flats) */
    PDM_F_FLATS = 3,     /* [u8]   .flats */
    PDM_F_UCA = 4,       /* [f64]  .uca      (NaN on flats) */
    PDM_F_TWI = 5,       /* [f64]  un-scaled twi (DEMProcessor.twi is 10x this) */
    PDM_F_EDGE_TODO = 6, /* [u8]   .edge_todo */
    PDM_F_EDGE_DONE = 7, /* [u8]   .edge_done */
    PDM_F_SECTION = 8,   /* [i8]   DEMProcessor.section (debug attribute, dem_processing.py:137) */
    PDM_F_TWI10 = 9,     /* [f64]  10 * twi: what DEMProcessor.twi stores (dem_processing.py:1674) */
    /* 10: reserved */
    PDM_F_FLAT0 = 11,    /* [u8]   mag == -1 before the one-pixel extension (shard halo exchange) */
    PDM_F_LINK = 12,     /* [u8]   facet index + kept-receiver bits of each cell (shard halo exchange) */
    PDM_F_CELL = 13,     /* [32 B] sweep record of a cell: area, taint, proportion, link, donor mask (shard halo exchange) */
    PDM_F_COUNT_ = 14
} pdm_field;

/* Flags of DEMProcessor that act on the hot path (dem_processing.py:105-154). */
typedef struct {
    int32_t drain_pits;             /* default 1 */
    int32_t drain_pits_min_border;  /* default 0 */
    int64_t drain_pits_max_iter;    /* default 300 */
    int64_t drain_pits_max_dist;    /* default 32; 0 = no index-distance filter */
    double  drain_pits_max_dist_xy; /* default 0 (= None) */
    int32_t apply_uca_limit_edges;  /* default 0 */
    int32_t circular_ref_maxcount;  /* default 50 */
    double  uca_saturation_limit;   /* default 32 */
} pdm_uca_params;

typedef struct {
    int64_t n_cells;
    int64_t n_sources;       /* cells nobody drains into (dem_processing.py:882-883) */
    int64_t n_drained;       /* cells whose area was pushed downstream */
    int64_t n_undone;        /* cells never reached (circular references) */
    int64_t n_pits;          /* flats & elev>0 examined by the pit search */
    int64_t n_pit_edges;
    int64_t n_pits_undrained;
    int64_t n_queue_items;   /* deferred (second-receiver) work items */
    int64_t n_restarts;      /* circular-reference restarts (dem_processing.py:951-964) */
    int64_t n_edge_todo;     /* border inflow cells */
    double  min_area;        /* nanmin(dX2*dY2), dem_processing.py:898 */
    float   ms_graph;        /* device time: section/proportion + receivers + pits + in-degree */
    float   ms_sweep;        /* device time: accumulation sweep */
    float   ms_total;
    float   ms_sweep_scan;   /* sweep kernel: until the last warp finished its source stripe */
    float   ms_sweep_kernel; /* sweep kernel: first warp in to last warp out (globaltimer) */
} pdm_uca_stats;

typedef struct {
    double  twi_min_slope;            /* default 1e-3 */
    double  twi_min_area;             /* DEMProcessor.twi_min_area after calc_uca */
    double  uca_saturation_limit;     /* default 32 */
    int32_t apply_twi_limits;         /* default 0 */
    int32_t apply_twi_limits_on_uca;  /* default 0 */
} pdm_twi_params;

/* elevation conditioning (the reference's fill_flats* / drain_pits_path traits, dem_processing.py:105-119, 154) */
typedef struct {
    int32_t fill_flats_below_sea;   /* default 0: cells with elev <= 0 are sea and never conditioned */
    int32_t fill_flats_source_tol;  /* default 1 */
    int32_t fill_flats_peaks;       /* default 1 */
    int32_t fill_flats_pits;        /* default 1 */
    double  maximum_pit_area;       /* default 32 cells; 0 disables calc_fill_pit_artifacts */
    int32_t drain_pits_max_iter;    /* default 300 */
    int32_t drain_pits_max_dist;    /* default 32 cells; 0 = no limit */
    double  drain_pits_max_dist_xy; /* default 0 = None */
} pdm_cond_params;

typedef struct {
    int64_t n_artifact_regions;     /* locally-minimal regions examined by the artifact pass */
    int64_t n_artifacts_filled;     /* regions raised by one unit */
    int64_t n_flat_regions;         /* regions handed to _fill_flat */
    int64_t distance_sweeps;        /* Jacobi sweeps of utils.get_distance (all regions together) */
    int64_t n_pits;                 /* strict local minima visited by calc_pit_drain_paths */
    int64_t n_pits_undrained;       /* the reference's "pits had no place to drain to" warning count */
    int64_t path_max_iter;          /* the reference's printed maxiter */
} pdm_cond_stats;

typedef struct pdm_tile pdm_tile;

/* ---- library ------------------------------------------------------------------------- */
int         pdm_abi_version(void);
const char *pdm_last_error(void);
/* Select the CUDA device for the calling thread's tiles (lazy context creation; safe to
 * call after fork() in a ProcessManager worker, process_manager.py:1267). */
int         pdm_init(int device);
int         pdm_device_count(int *count);
/* number of CUDA kernels this library has launched in this process (bench accounting) */
unsigned long long pdm_launch_count(void);

/* Engine of the UCA accumulation (a7) on stand-alone tiles, process-wide.
 *   mode 0: work-list sweep (L2 atomics, fastest on one tile; default), 1: tile sweep (pull-based,
 *   tile-resident, bit-reproducible; always used by the pdm_shard_* path), -1: take it from the
 *   environment (PYDEM_B200_SWEEP=worklist|tile).
 *   strict 1: the work-list holds every count-off back until the adds it publishes have returned
 *   (cross-check of its ordering assumption), 0: off, -1: PYDEM_B200_SWEEP_STRICT.
 * Replaces nothing in the reference (cyutils.drain_area, cyutils.pyx:78-187, has one engine). */
int pdm_set_sweep_mode(int mode, int strict);
int pdm_get_sweep_mode(void);
/* page-locked host buffers for the host-buffer entry points (full-rate PCIe copies) */
int         pdm_host_alloc(size_t bytes, void **out);
int         pdm_host_free(void *p);
void        pdm_default_uca_params(pdm_uca_params *p);
void        pdm_default_twi_params(pdm_twi_params *p);
void        pdm_default_cond_params(pdm_cond_params *p);

/* ---- tile handles ---------------------------------------------------------------------- */
/* R x C tile on the current device.  stream: a cudaStream_t passed as void* (NULL = the
 * legacy default stream); every kernel and copy of the tile is issued on it. */
int pdm_tile_create(int64_t R, int64_t C, void *stream, pdm_tile **out);
int pdm_tile_destroy(pdm_tile *t);
/* Per-row spacing (DEMProcessor.dX, dY: R-1 "fence" values; dX2, dY2: R "post" values).
 * thA/thB: optional (NULL -> computed with host atan2) per-fence facet angles
 * atan2(dY,dX) / atan2(dX,dY) (dem_processing.py:1936) so a NumPy host can pass the
 * values its own arctan2 produces. */
int pdm_tile_set_spacing(pdm_tile *t, const double *dX, const double *dY,
                         const double *dX2, const double *dY2,
                         const double *thA, const double *thB);
/* host <-> device for one field (sizes implied by the tile shape and the field dtype) */
int pdm_tile_upload(pdm_tile *t, int field, const void *host);
int pdm_tile_download(pdm_tile *t, int field, void *host);
/* like pdm_tile_download, but the copy runs on a side stream after the work queued so far and
 * overlaps later stages; the host buffer (page-locked: pdm_host_alloc) is valid after pdm_tile_sync */
int pdm_tile_download_async(pdm_tile *t, int field, void *host);
/* raw device pointer of a field (for zero-copy interop, e.g. NCCL halo exchange) */
int pdm_tile_device_ptr(pdm_tile *t, int field, void **dev);
/* declare a field valid after writing it through pdm_tile_device_ptr (device-side producers) */
int pdm_tile_mark_resident(pdm_tile *t, int field);
/* on != 0: use the literal 8 x (3 divisions + atan2) stencil formulation instead of the
 * division-/atan2-saving one; both give bit-identical MAG/DIR (checked by the GPU tests) */
int pdm_tile_set_stencil_parity(pdm_tile *t, int on);
int pdm_tile_sync(pdm_tile *t);
/* device self-test: the stencil's reciprocal-based division against IEEE division on n_pairs
 * pseudo-random operand pairs; *mismatches must come back 0 */
int pdm_selftest_division(unsigned long long seed, long long n_pairs, unsigned long long *mismatches);

/* a1+a2: _tarboton_slopes_directions (dem_processing.py:1753-1903) + _find_flats_edges
 * (657-680) + the stamping of flats (611-613).  In: ELEV.  Out: MAG, DIR, FLATS. */
int pdm_tile_slopes_directions(pdm_tile *t);
/* DEMProcessor.find_flats (305-306): FLATS = (MAG == -1). */
/* pdm_tile_upload(ELEV) + pdm_tile_slopes_directions in one call: the stencil of a row chunk runs while the next
 * chunk is still on its way over PCIe (same results) */
int pdm_tile_upload_slopes_directions(pdm_tile *t, const void *host_elev);
int pdm_tile_find_flats(pdm_tile *t);
/* a3..a7: _calc_uca_chunk (864-987) incl. _calc_uca_section_proportion (1021),
 * _mk_adjacency_matrix (1072), _mk_connectivity_pits (1269) and cyutils.drain_area.
 * In: ELEV, DIR, MAG, FLATS.  Out: UCA, EDGE_TODO, EDGE_DONE; MAG/FLATS updated at
 * drained pits (1370-1371).  stats may be NULL. */
int pdm_tile_uca(pdm_tile *t, const pdm_uca_params *p, pdm_uca_stats *stats);
/* The cells the last pdm_tile_uca examined as pits (flats & elev > 0) with their MAG / FLATS values
 * after the search -- the only cells of MAG / FLATS that calc_uca changes (dem_processing.py:1370-1371),
 * so a host mirror can patch its arrays in place without downloading whole fields.  Call with
 * capacity 0 to learn *n (= pdm_uca_stats.n_pits). */
int pdm_tile_pit_updates(pdm_tile *t, int64_t capacity, int32_t *cells, double *mag, uint8_t *flats, int64_t *n);
/* a8: calc_uca(uca_init, edge_init_data) (719-744, 769-771) + _calc_uca_chunk_update
 * (778-862).  UCA must hold uca_init.  Edge strips: left/right have R entries, top/bottom
 * C entries; data f64, done/todo u8.  Out: UCA += delta, EDGE_TODO, EDGE_DONE. */
int pdm_tile_uca_update(pdm_tile *t, const pdm_uca_params *p,
                        const double *data_left, const double *data_right,
                        const double *data_top, const double *data_bottom,
                        const uint8_t *done_left, const uint8_t *done_right,
                        const uint8_t *done_top, const uint8_t *done_bottom,
                        const uint8_t *todo_left, const uint8_t *todo_right,
                        const uint8_t *todo_top, const uint8_t *todo_bottom,
                        pdm_uca_stats *stats);
/* on != 0: pdm_tile_uca_update reuses the graph (receivers, proportions, pit drains) built by the tile's last
 * pdm_tile_uca instead of rebuilding it on every call like the reference does (dem_processing.py:787-793) --
 * for tiles that stay in HBM between the edge corrections of a mosaic (process_manager.py:224-284), whose
 * direction / magnitude do not change in between.  The caller must not upload DIR / MAG / FLATS in between. */
int pdm_tile_set_keep_graph(pdm_tile *t, int on);
/* a9: calc_twi (1647-1677).  In: UCA, MAG.  Out: TWI (un-scaled, the return value) and TWI10
 * (= 10 * twi, the attribute). */
int pdm_tile_twi(pdm_tile *t, const pdm_twi_params *p);

/* ---- elevation conditioning: what calc_slopes_directions runs first when
 * fill_flats / drain_pits_path are on (dem_processing.py:601-609).  In/out: ELEV, modified in
 * place on the device; every derived field of the tile is invalidated.  Bit-identical to the
 * reference on NaN-free elevation (pits of equal elevation are visited in raster order; the
 * reference's np.argsort tie order is platform-defined).  Not available on a row shard. */
/* calc_fill_pit_artifacts (396-426) */
int pdm_tile_fill_pit_artifacts(pdm_tile *t, const pdm_cond_params *p, pdm_cond_stats *stats);
/* calc_fill_flats (551-585) -> _fill_flat (308-394); runs the artifact pass first when
 * maximum_pit_area != 0, as the reference does */
int pdm_tile_fill_flats(pdm_tile *t, const pdm_cond_params *p, pdm_cond_stats *stats);
/* calc_pit_drain_paths (428-548); needs the spacing (dX, dY) */
int pdm_tile_pit_drain_paths(pdm_tile *t, const pdm_cond_params *p, pdm_cond_stats *stats);

/* ---- row shards (one tile per GPU = a block of rows of one big DEM) --------------------------
 * The tile has one halo row on every side where another rank continues the grid.  The host
 * driver fills halo rows (elev, then flat0, then link) from the neighbouring ranks -- NCCL
 * send/recv on the row views obtained with pdm_tile_device_ptr -- and calls the stages in this
 * order: slopes, ccl, {label_pack / label_unpack until no rank changed}, flats_extend, links,
 * (LINK halo rows), indeg, (CELL halo rows), {sweep, sweep_sent, CELL halo rows until no rank sent},
 * finalize.
 * Equivalent of pyDEM's cross-tile edge resolution (process_manager.py:1090-1249) with true
 * halo stencils, so the sharded result equals the single-tile result. */
int pdm_tile_set_window(pdm_tile *t, int64_t row_off, int64_t R_global, int64_t own_lo, int64_t own_hi,
                        const double *th_row);
int pdm_shard_slopes(pdm_tile *t);
int pdm_shard_ccl(pdm_tile *t);
int pdm_shard_label_pack(pdm_tile *t, int64_t row, void *out_labels, void *out_elev);
int pdm_shard_label_unpack(pdm_tile *t, int64_t row, const void *in_labels, const void *in_elev, void *changed);
int pdm_shard_flats_extend(pdm_tile *t);
int pdm_shard_links(pdm_tile *t, const pdm_uca_params *p);
/* drain_pits on row shards (_mk_connectivity_pits, dem_processing.py:1269-1382; needs pdm_shard_p2p_connect_all):
 * pdm_tile_set_global_spacing once (host arrays of R_global - 1 fences); per pass, between links and indeg:
 * exchange strips of ELEV and FLAT0 (= the pit mask after links) -- the Hu / Hd = drain_pits_max_iter + 1 owned rows
 * next to each boundary -- then pdm_shard_pits (device pointers to the received strips; in_up / in_dn: int32
 * [Hin][C] out, pit edges that end on the neighbour's Hin rows next to the boundary, Hin >= drain_pits_max_dist),
 * exchange in_up / in_dn, pdm_shard_pit_in_apply with what arrived. */
int pdm_tile_set_global_spacing(pdm_tile *t, const double *dX, const double *dY, int64_t n);
int pdm_shard_pits(pdm_tile *t, const pdm_uca_params *p, const void *E_up, const void *P_up, int64_t Hu,
                   const void *E_dn, const void *P_dn, int64_t Hd, void *in_up, void *in_dn, int64_t Hin);
int pdm_shard_pit_in_apply(pdm_tile *t, const void *from_up, const void *from_dn, int64_t Hin);
int pdm_shard_indeg(pdm_tile *t);
int pdm_shard_sweep(pdm_tile *t, int first);
int pdm_shard_sweep_sent(pdm_tile *t, void *sent);
/* One sweep across the GPUs of the row shards instead of exchange rounds: every rank exports CUDA IPC
 * handles of its sweep records and of its tile queue (pdm_shard_p2p_export: *size in = room at buf, out =
 * bytes written; buf NULL asks for the size), the host driver hands each rank the blobs of the ranks above /
 * below and of rank 0 (NULL where there is none / on rank 0), and from then on pdm_shard_sweep(t, 1) on
 * every rank is the whole accumulation: boundary tiles read the neighbour's boundary row and wake its
 * tiles over NVLink peer memory, one in-flight counter on rank 0 ends all kernels together.
 * Semantics as pyDEM's process_uca_edges (process_manager.py:1090-1249) without its rounds. */
int pdm_shard_p2p_export(pdm_tile *t, void *buf, int64_t *size);
int pdm_shard_p2p_connect(pdm_tile *t, const void *up, const void *down, const void *root, int world, int rank);
/* the same with every rank's export (world blobs back to back in rank order, e.g. from an all-gather): the
 * accumulation then runs on the work-list engine as one sweep across the GPUs */
int pdm_shard_p2p_connect_all(pdm_tile *t, const void *blobs, int world, int rank);
int pdm_shard_p2p_disconnect(pdm_tile *t);   /* unmap the peers' memory: all ranks, then a barrier, before any tile is destroyed */
int pdm_shard_finalize(pdm_tile *t, const pdm_uca_params *p, pdm_uca_stats *stats);

/* ---- the row-sharded path without any host-side framework (pdm_comm_init) ---------
 * One process per GPU.  NCCL (loaded at run time) moves the halo rows, CUDA-IPC peer memory carries the
 * accumulation sweep.  Call order on every rank:
 *   pdm_init(device); rank 0: pdm_comm_unique_id(id) and hand the 128 bytes to every rank;
 *   pdm_comm_init(rank, nranks, id); pdm_tile_create / pdm_tile_set_spacing / pdm_tile_set_window
 *   (+ pdm_tile_set_global_spacing for drain_pits); pdm_shard_connect(t);
 *   per pass: upload ELEV of the owned rows, pdm_shard_run(t, ...), download the owned rows of the results;
 *   pdm_shard_disconnect(t); pdm_tile_destroy(t); pdm_comm_finalize().
 * pdm_shard_run is the collective equivalent of DEMProcessor.calc_slopes_directions + calc_uca + calc_twi on the
 * whole grid (pydem/dem_processing.py:587-776, 1647-1677): the result equals the single-tile result. */
#define PDM_COMM_ID_BYTES 128
int pdm_comm_unique_id(void *id_out);
int pdm_comm_init(int rank, int nranks, const void *id);
int pdm_comm_finalize(void);
int pdm_comm_barrier(pdm_tile *t);
int pdm_shard_connect(pdm_tile *t);
int pdm_shard_disconnect(pdm_tile *t);
int pdm_shard_run(pdm_tile *t, const pdm_uca_params *p, const pdm_twi_params *twi, pdm_uca_stats *stats, int *label_rounds);

/* ---- one-shot host-buffer calls ---------------------------------------------------------- */
/* DEMProcessor.calc_slopes_directions with the conditioning flags off. */
int pdm_slopes_directions(const double *elev, int64_t R, int64_t C,
                          const double *dX, const double *dY,
                          const double *thA, const double *thB,
                          double *mag, double *direction, uint8_t *flats);
/* DEMProcessor.calc_uca() full mode.  mag/flats are in-out (pit updates). */
int pdm_uca(const double *elev, const double *direction, double *mag, uint8_t *flats,
            int64_t R, int64_t C, const double *dX, const double *dY,
            const double *dX2, const double *dY2, const double *thA, const double *thB,
            const pdm_uca_params *p,
            double *uca, uint8_t *edge_todo, uint8_t *edge_done, pdm_uca_stats *stats);
/* DEMProcessor.calc_uca(uca_init=..., edge_init_data=[data, done, todo]) update mode.
 * uca is in-out: uca_init on entry, uca_init + propagated edge deltas on return. */
int pdm_uca_update(const double *elev, const double *direction, double *mag, uint8_t *flats, double *uca,
                   int64_t R, int64_t C, const double *dX, const double *dY,
                   const double *dX2, const double *dY2, const double *thA, const double *thB,
                   const pdm_uca_params *p,
                   const double *data_left, const double *data_right,
                   const double *data_top, const double *data_bottom,
                   const uint8_t *done_left, const uint8_t *done_right,
                   const uint8_t *done_top, const uint8_t *done_bottom,
                   const uint8_t *todo_left, const uint8_t *todo_right,
                   const uint8_t *todo_top, const uint8_t *todo_bottom,
                   uint8_t *edge_todo, uint8_t *edge_done);
/* DEMProcessor.calc_twi(). */
int pdm_twi(const double *uca, const double *mag, int64_t n, const pdm_twi_params *p, double *twi);

#ifdef __cplusplus
}
#endif
#endif /* PYDEM_B200_H */
