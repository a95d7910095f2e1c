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
    TC_PHASE = 80,      // debug: ns summed over CTAs per phase {claim, load, count, levels, store, schedule}
    TC_HIST = 96,       // PYDEM_B200_TS_DEBUG timeline: per 100 us bucket {visits, cells completed} x 64
    TC_N = 96 + 128
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
};

}  // namespace ts
