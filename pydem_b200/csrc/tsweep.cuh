// tsweep.cuh -- shared declarations of the tile-resident sweep (tsweep.cu, tsweep_p2p.cu).
#pragma once
#include "pdm_internal.cuh"

namespace ts {

// counters of a sweep (u64 slots; the hammered ones live on their own 128-byte lines)
enum {
    TC_HEAD = 0,        // queue tickets handed out
    TC_TAIL = 16,       // queue items produced
    TC_INFLIGHT = 32,   // tiles pending or running (0 = the sweep is over)
    TC_ABORT = 48,      // 1: watchdog, 2: queue overrun
    TC_VISITS = 64,     // tile visits
    TC_CELLS = 65,      // cells completed
    TC_SENT = 66,       // completed cells whose receiver lives on a neighbouring rank (this launch)
    TC_SOURCES = 67,    // cells nobody drains into
    TC_LEVELS = 68,     // in-tile frontier levels, summed over visits
    TC_REQUEUE = 69,    // visits repeated because a neighbour published while the tile was loaded
    TC_T_START = 70,    // globaltimer ns: first CTA in / last CTA out
    TC_T_END = 71,
    TC_QUEUED = 72,     // tiles queued by the set-up of this launch
    TC_DEFER = 73,      // visits deferred to the back of the queue (neighbour published, queue busy)
    TC_ARRIVED = 76,    // multi-GPU sweep, on rank 0: ranks whose queues are set up, summed over all sweeps so far (never reset)
    TC_PHASE = 80,      // debug: ns summed over CTAs per phase {claim, load, count, levels, store, schedule}
    TC_HIST = 96,       // PYDEM_B200_TS_DEBUG timeline: per 100 us bucket {visits, cells completed} x 64
    TC_LATE = 96 + 128, // debug: phases {claim, load, count, flow, store, schedule} ns, visits, cells of the visits after dbg x 100 us
    TC_DBG2 = 96 + 128 + 8, // debug: cycles {busy turns total, inside the drain step, in the hand-over, looking for work} of the busiest-warp turns of late visits
    TC_WLC = 96 + 128 + 16,  // counters (CT_*, pdm_internal.cuh) of a work-list sweep that spans the GPUs (sweep.cu, pdm_launch_sweep_p2p)
    TC_N = 96 + 128 + 16 + CT_N
};

struct Args {
    TRec *rec;                // sweep records (pdm_internal.cuh)
    const double *row_area;
    const int32_t *pit_beg, *pit_end, *pit_dst;
    const double *pit_w;
    int32_t *pit_cnt;         // pit edges still to arrive at a cell
    double *pit_acc_a, *pit_acc_t;
    Win w;
    int32_t ntx, nty, ntiles;
    int32_t *slots;
    uint32_t cap_mask;
    uint32_t *flag;
    unsigned long long *ctr;
    int32_t dbg;
    int32_t has_pits;         // the graph has pit edges: pit receivers are gated on their pit counters
    // ---- one sweep across the row shards of several GPUs (peer memory over NVLink; shard.cu):
    unsigned long long *inflight;   // tiles pending or running, all ranks together (rank 0's counter; = ctr + TC_INFLIGHT on one GPU)
    unsigned long long *arrived;    // start barrier: ranks whose queues are set up (rank 0's counter, monotonic)
    unsigned long long start_target;
    int32_t p2p;                    // bit 0: the rows above [lo] live on a peer, bit 1: the rows below [hi)
    const TRec *halo_src[2];        // the peer's boundary row of records (its last / first owned row), column 0
    uint32_t *peer_flag[2];         // the peer's tile state words, queue and counters
    int32_t *peer_slots[2];
    unsigned long long *peer_ctr[2];
    uint32_t peer_cap_mask[2];
    int32_t peer_tile0[2];          // first tile of the peer's boundary tile row
};

// what a rank publishes to its neighbours (pdm_shard_p2p_export): CUDA IPC handles of the record array and of
// the control block [counters | tile flags | queue slots], offsets of the pointers inside their allocations
struct P2PExport {
    unsigned char h_rec[64], h_ctl[64];
    long long off_rec, off_ctl;
    long long off_flag, off_slots;      // relative to the control block
    long long lo, hi, C;                // owned local rows, columns
    int ntx, nty;
    unsigned cap_mask;
    int device;
    long long proc_tag;                 // same process: pointers are used as they are (tests on one GPU)
    void *raw_rec, *raw_ctl;
};

}  // namespace ts
