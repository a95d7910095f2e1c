#!/usr/bin/env python
"""bench.py -- pyDEM hot path on B200: Mcells/s for slope+aspect -> UCA -> TWI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size 4096]

One "step" = one pass of the hot path (calc_slopes_directions + calc_uca + calc_twi, conditioning
flags off) over one synthetic fractal DEM.  N = 1 runs BASELINE.json configs[1] (4096 x 4096,
f64, dX = dY = 30 m); N > 1 runs the row-sharded single-DEM path (pydem_b200.sharded: 4096 rows x
4096 columns per GPU, halo rows exchanged over NCCL), weak scaling.

Printed JSON line (rank 0):
  value        whole-job Mcells/s with the elevation already resident in HBM (device-timed)
  e2e          the same metric through the reference-facing operator
               (DEMProcessor(elev=<pinned host array>).calc_twi()): H2D of the DEM and D2H of
               mag/direction/uca/twi/flats/edge masks inside the timed region
  roofline     dominant kernel (the UCA sweep) against the measured HBM peak
  cpu_baseline the oracle port on one host core on the same DEM (whole DEM up to 4096 x 4096)
  --impl reference: the reference's CPU path on all host cores (see reference_arm()).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mcells/s slope+aspect and UCA (one pass of the D-infinity hot path: slope+aspect -> UCA -> TWI, f64; per-stage values in per_stage)"
SPACING = 30.0
BYTES_PER_CELL = 24.0   # read elev 8 + direction 8, write uca 8 (also 24 for the stencil)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self):
        """wall-clock stamp (nvidia-smi prints local time with ms)"""
        import datetime
        return datetime.datetime.now()

    def stop(self, t0=None, t1=None):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.1)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f")
                rows.append((ts, float(c[1]), float(c[2]), c))
            except ValueError:
                continue
        inside = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
        out["window"] = "timed region" if inside else "whole run (no sample fell inside the timed region)"
        for ts, a, b, c in (inside or rows):
            sm.append(a); mx.append(b)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            out.update({"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                        "samples": len(sm)})
        return out


# ----------------------------------------------------------------------------------------------
# CPU legs (oracle = checker; only this file's cpu_baseline / --impl reference legs may time it)
# ----------------------------------------------------------------------------------------------
def _cpu_hot_path(E, use_ref_sweep, drain_pits=True, keep=False):
    """One pass of the hot path on one host core.  use_ref_sweep: run the accumulation with the
    reference's own compiled Cython kernel (oracle/_ref/cyutils*.so, built from
    /root/reference/pydem/cyfuncs/cyutils.pyx) instead of the oracle's restatement of it.
    keep: return the output arrays (the parity check of the GPU arm) instead of the cell count."""
    from oracle import oracle as orc
    dp = orc.OracleDEMProcessor(E, dX=SPACING, dY=SPACING, fill_flats=False, drain_pits_path=False, drain_pits=drain_pits)
    dp.calc_slopes_directions()
    if not use_ref_sweep:
        dp.calc_uca()
    else:
        from oracle import ref_harness
        cy = ref_harness.load_ref_cyutils()
        g, sec = dp._graph()
        cptr, cidx, cdat, rptr, ridx = g.export()
        inflow, _, _ = g.sums()
        R, C = E.shape
        area = np.full(R * C, SPACING * SPACING)
        done = np.ascontiguousarray(inflow == 0)
        ids = done.copy()
        todo = np.zeros(R * C)
        cy.drain_area(area, done, ids, cptr.astype(np.int32), cidx.astype(np.int32), cdat, rptr.astype(np.int32),
                      ridx.astype(np.int32), R, C, todo, todo.copy())
        area[dp.flats.ravel()] = np.nan
        dp.uca = area.reshape(R, C)
        dp.twi_min_area = SPACING * SPACING
    twi = dp.calc_twi()
    if keep:
        return dict(mag=np.asarray(dp.mag), direction=np.asarray(dp.direction), flats=np.asarray(dp.flats),
                    uca=np.asarray(dp.uca), edge_todo=np.asarray(dp.edge_todo), edge_done=np.asarray(dp.edge_done),
                    twi=np.asarray(twi))
    return E.size


# the reference arm's unit of work: the benchmark DEM cut 4 x 4 into tiles with a one-pixel overlap, one
# tile per process -- how the reference's ProcessManager spreads a DEM over cores
# (process_manager.py:1251-1288, tiles from compute_grid with overlap, :560-600)
REF_GRID = 4


def _ref_tiles(n):
    t = n // REF_GRID
    out = []
    for a in range(REF_GRID):
        for b in range(REF_GRID):
            out.append((max(a * t - 1, 0), min((a + 1) * t + 1, n), max(b * t - 1, 0), min((b + 1) * t + 1, n)))
    return out


_REF_DEM = {}


def _ref_dem(n, variant):
    key = (n, variant)
    if key not in _REF_DEM:
        from pydem_b200 import synth
        _REF_DEM[key] = (synth.conditioned_fractal_dem(n, 0, wrap_rows=True) if variant == "conditioned"
                         else synth.fractal_dem(n, 0))
    return _REF_DEM[key]


def _ref_worker(args):
    box, n, use_ref, variant = args
    E = np.ascontiguousarray(_ref_dem(n, variant)[box[0]:box[1], box[2]:box[3]])     # inherited from the parent (fork)
    t = time.perf_counter()
    cells = _cpu_hot_path(E, use_ref, drain_pits=(variant == "conditioned"))
    return cells, time.perf_counter() - t


def cpu_baseline_leg(E_full, window=4096, drain_pits=True):
    """Oracle port, one core, on the benchmark DEM itself (top-left window when the DEM is larger than
    `window`): ~10-20 s of CPU work at 4096 x 4096.  Returns (baseline dict, oracle arrays or None)."""
    w = min(window, E_full.shape[0], E_full.shape[1])
    E = np.ascontiguousarray(E_full[:w, :w])
    _cpu_hot_path(E[:256, :256].copy(), False, drain_pits)   # builds/loads the oracle library
    t = time.perf_counter()
    arrays = _cpu_hot_path(E, False, drain_pits, keep=True)
    dt = time.perf_counter() - t
    whole = w == E_full.shape[0] == E_full.shape[1]
    what = "the whole benchmark DEM" if whole else "the %dx%d top-left window of the benchmark DEM" % (w, w)
    return ({"value": E.size / dt / 1e6, "unit": "Mcells/s", "cores": 1, "kind": "port",
             "sample": "oracle/pdm_oracle.c (C restatement, 1 core) on %s (%dx%d), slope+aspect + UCA + TWI, drain_pits=%s, %.1f s"
                       % (what, w, w, bool(drain_pits), dt)}, arrays if whole else None)


# parity bar of tests/helpers.py (north_star: stated float tolerance, integer / bool outputs bit-exact)
PARITY_BAR = {"mag_rel": 1e-12, "direction_abs": 1e-12, "uca_rel": 1e-9, "twi_abs": 1e-9}


def parity_block(ref, got):
    """Error measures of the GPU arm's outputs against the oracle's on the same DEM, and the verdict."""
    out = {}
    with np.errstate(invalid="ignore", divide="ignore"):
        for k in ("mag", "direction", "uca", "twi"):
            a, b = np.asarray(ref[k], float), np.asarray(got[k], float)
            d = np.abs(a - b)
            fin = np.isfinite(d)
            out[k + "_nan_mismatch"] = int((np.isnan(a) != np.isnan(b)).sum())
            out[k + "_abs"] = float(d[fin].max()) if fin.any() else 0.0
            rel = d / np.abs(a)
            fin = np.isfinite(rel)
            out[k + "_rel"] = float(rel[fin].max()) if fin.any() else 0.0
        for k in ("flats", "edge_todo", "edge_done"):
            out[k + "_mismatch_bits"] = int((np.asarray(ref[k], bool) != np.asarray(got[k], bool)).sum())
    ok = all(out[k + "_mismatch_bits"] == 0 for k in ("flats", "edge_todo", "edge_done"))
    ok = ok and all(out[k + "_nan_mismatch"] == 0 for k in ("mag", "direction", "uca", "twi"))
    ok = ok and out["mag_rel"] <= PARITY_BAR["mag_rel"] and out["direction_abs"] <= PARITY_BAR["direction_abs"]
    ok = ok and out["uca_rel"] <= PARITY_BAR["uca_rel"] and out["twi_abs"] <= PARITY_BAR["twi_abs"]
    out["bar"] = PARITY_BAR
    out["ok"] = bool(ok)
    return out


def reference_arm(args):
    """Reference CPU path on all host cores, like for like with the GPU arm: the SAME seed-0 benchmark
    DEM with the same flags, cut 4 x 4 into tiles with a one-pixel overlap, one tile per process -- the
    reference has no intra-tile threading, its ProcessManager runs one tile per process
    (process_manager.py:1251-1288).  A step = the 16 tiles of the DEM.  Only the reference's native piece
    can travel to the GPU box (oracle/_ref: cyutils.pyx compiled with the reference's flags): it does the
    UCA sweep, ~90% of the reference's time; the NumPy layers around it run as the oracle's C port
    (faster than the reference's NumPy, i.e. optimistic for the reference) -- hence kind "port+ref-sweep".
    A second stated baseline: the reference's compiled sweep on the WHOLE DEM as one tile on one core
    (one timed run, what a user without the tiling gets)."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc, ref_harness
    orc.build()
    use_ref = ref_harness.load_ref_cyutils() is not None
    cores = os.cpu_count() or 1
    n = args.size
    variant = args.variant
    pits = variant == "conditioned"
    E = _ref_dem(n, variant)
    tiles = _ref_tiles(n)
    procs = min(cores, len(tiles))
    ctx = mp.get_context("fork")
    whole = None
    with ctx.Pool(procs) as pool:
        jobs = [(box, n, use_ref, variant) for box in tiles]
        if args.warmup > 0:
            pool.map(_ref_worker, [((0, 258, 0, 258), n, use_ref, variant)] * procs)
        # second baseline, concurrently on one more process: the whole DEM as ONE tile on one core
        if not args.no_whole_dem:
            whole_pool = ctx.Pool(1)
            whole_job = whole_pool.apply_async(_ref_worker, (((0, n, 0, n), n, use_ref, variant),))
        t0 = time.perf_counter()
        cells = 0
        for s in range(args.steps):
            res = pool.map(_ref_worker, jobs, chunksize=1)
            cells += n * n          # the DEM's cells (overlap pixels are computed twice, counted once)
        dt = time.perf_counter() - t0
        if not args.no_whole_dem:
            c_w, t_w = whole_job.get()
            whole_pool.close()
            whole = {"value": c_w / t_w / 1e6, "unit": "Mcells/s", "cores": 1, "seconds": t_w,
                     "sample": "the whole %dx%d benchmark DEM as one tile, one core, one run (ran beside the tiled steps)" % (n, n)}
    val = cells / dt / 1e6
    kind = "port+ref-sweep" if use_ref else "port"
    flags_txt = "fill_flats=False, drain_pits_path=False, drain_pits=%s" % pits
    line = {"metric": METRIC, "value": val, "unit": "Mcells/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": _workload_text(n, variant, flags_txt),
                       "parallelism": "%d processes, the DEM cut %dx%d into %dx%d tiles (+1 pixel overlap), one tile per "
                                      "process (reference ProcessManager style)" % (procs, REF_GRID, REF_GRID, n // REF_GRID, n // REF_GRID)},
            "cpu_baseline": {"value": val, "unit": "Mcells/s", "cores": procs, "kind": kind,
                             "sample": ("UCA sweep = the reference's compiled cyutils.drain_area (oracle/_ref); " if use_ref
                                        else "UCA sweep = oracle restatement; ") +
                                       "stencil/flats/graph/pit drains/TWI = oracle C port; %d steps x 16 tiles of the same "
                                       "seed-0 DEM the GPU arm runs" % args.steps,
                             "whole_dem_one_core": whole},
            "e2e": {"value": val, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _workload_text(n, variant, flags_txt):
    return ("%dx%d fractal DEM (spectral synthesis, seed 0, H=0.8, 1..1001 m%s), dX=dY=30 m, "
            "slope+aspect + UCA + TWI, %s (BASELINE.json configs[1], %s variant)"
            % (n, n, ", priority-flood+eps conditioned" if variant == "conditioned" else "", flags_txt,
               "primary" if variant == "conditioned" else "'sinks'"))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def _note(msg):
    if os.environ.get("PDM_BENCH_VERBOSE"):
        print("[bench rank %s %.1fs] %s" % (os.environ.get("RANK", "0"), time.perf_counter() - _T0, msg), file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from pydem_b200 import synth, _lib, _pinned, tile as T, DEMProcessor

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    _lib.init(local)
    L = _lib.load()
    n = args.size
    peak_gbs, peak_src = measured_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    variant = args.variant
    if variant == "conditioned":
        # BASELINE.md primary variant: conditioned once on the host (priority-flood + eps; rows wrap so
        # that the same block tiles vertically for the N>1 runs) -> long river networks, no interior pits
        E = synth.conditioned_fractal_dem(n, 0, wrap_rows=True)
    else:
        E = synth.fractal_dem(n, 0)                      # secondary 'sinks' variant: raw fractal
    # the reference's default drain_pits=True on the conditioned variant; on row shards it needs the multi-GPU
    # work-list sweep over peer memory (pits drain across shard boundaries), i.e. not PYDEM_B200_SHARD_P2P=0
    p2p_on = os.environ.get("PYDEM_B200_SHARD_P2P", "1") != "0" and not os.environ.get("PYDEM_B200_SHARD_SWEEP", "w").lower().startswith("t")
    pits_flag = 1 if (variant == "conditioned" and (world == 1 or p2p_on)) else 0
    flags_txt = ("fill_flats=False, drain_pits_path=False, drain_pits=%s" % bool(pits_flag))
    if world == 1:
        stream = torch.cuda.current_stream().cuda_stream
        dt = T.DeviceTile(n, n, stream=stream)
        dt.set_spacing(SPACING, SPACING)
        dt.upload(T.F_ELEV, E)
        stats = {}

        sev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        stage_ms = {"slopes": [], "uca": [], "twi": []}

        def step():
            sev[0].record()
            dt.slopes_directions()
            sev[1].record()
            st = dt.uca(drain_pits=pits_flag)      # synchronises the stream (reads the sweep's counters)
            sev[2].record()
            dt.twi()
            sev[3].record()
            stats.update(st)
            stage_ms["_pending"] = True
            return st

        def collect_stage_times():
            # called after the step's work has completed (the next uca() call or the final barrier syncs)
            if stage_ms.pop("_pending", False):
                sev[3].synchronize()
                stage_ms["slopes"].append(sev[0].elapsed_time(sev[1]))
                stage_ms["uca"].append(sev[1].elapsed_time(sev[2]))
                stage_ms["twi"].append(sev[2].elapsed_time(sev[3]))
        cells_per_step = n * n
        workload = _workload_text(n, variant, flags_txt)
        parallelism = "1 GPU, single tile"
    else:
        from pydem_b200 import sharded
        sh = sharded.ShardedDEM(rows_per_rank=n, cols=n, spacing=SPACING, seed=0, profile=True, block=E)
        sh.uca_flags = {"drain_pits": pits_flag}
        stats = {}

        def step():
            st = sh.step()
            stats.update(st)
            return st
        cells_per_step = n * n * world
        workload = ("%dx%d DEM = the %dx%d block of the 1-GPU run (%s) repeated %d times vertically (periodic, "
                    "seamless), row-sharded over %d GPUs (%d rows each, halo rows over NCCL send/recv, UCA as one work-list sweep across the GPUs over NVLink peer memory), dX=dY=30 m, "
                    "slope+aspect + UCA + TWI, %s"
                    % (n * world, n, n, n, "conditioned fractal" if variant == "conditioned" else "raw fractal", world,
                       world, n, flags_txt))
        parallelism = "row-block x%d" % world

    shard_parity = None
    if world > 1 and not args.no_parity:
        shard_parity = sharded_check(sh, E, world, rank, pits_flag)
    sampler = ClockSampler(local) if rank == 0 else None
    _note("setup done")
    for _ in range(args.warmup):
        step()
    barrier()
    _note("warmup done")
    t_begin = sampler.mark() if sampler else None
    launches0 = L.pdm_launch_count()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    sweep_ms, sweep_kernel_ms, slopes_ms = [], [], []
    ev0.record()
    for _ in range(args.steps):
        st = step()
        if world == 1:
            collect_stage_times()      # waits for the step's last (0.1 ms) kernel: steps are serial anyway
        sweep_ms.append(st.get("ms_sweep", 0.0) or (st.get("ms_sweep_first", 0.0) + st.get("ms_sweep_resume", 0.0)))
        sweep_kernel_ms.append(st.get("ms_sweep_kernel", 0.0))
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    _note("timed loop done: %.1f ms" % ms)
    launches = L.pdm_launch_count() - launches0
    clocks = sampler.stop(t_begin, sampler.mark()) if sampler else None
    if world > 1:
        tms = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    ms_per_step = ms / args.steps
    value = cells_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- e2e through the reference-facing operator (host buffers, copies inside the timed region)
    e2e = None
    if world == 1:
        Eh = _pinned.pinned_copy(E)
        kw = dict(dX=SPACING, dY=SPACING, fill_flats=False, drain_pits_path=False, drain_pits=bool(pits_flag))

        def e2e_step(all_outputs=False):
            dp = DEMProcessor(elev=Eh, **kw)
            twi = dp.calc_twi()                # what the reference call returns (dem_processing.py:1647-1677)
            chk = float(twi[n // 2, n // 2])   # the step's result is read on the host
            if all_outputs:                    # every array attribute the reference object carries after the call
                for name in ("mag", "direction", "uca", "flats", "edge_todo", "edge_done"):
                    chk += float(getattr(dp, name)[n // 2, n // 2])
            dp._free_tile()
            return chk
        gpu_out = None
        if rank == 0 and not args.no_parity:
            gpu_out = dict(mag=dt.download(T.F_MAG), direction=dt.download(T.F_DIR), flats=dt.download(T.F_FLATS),
                           uca=dt.download(T.F_UCA), edge_todo=dt.download(T.F_EDGE_TODO), edge_done=dt.download(T.F_EDGE_DONE),
                           twi=dt.download(T.F_TWI))
        dt.close()

        def time_e2e(all_outputs):
            for _ in range(max(1, min(args.warmup, 2))):
                e2e_step(all_outputs)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step(all_outputs)
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / args.steps
        t_e2e = time_e2e(False)
        t_all = time_e2e(True)
        # result-only: DEMProcessor(elev=<pinned host>).calc_twi() returns twi; the other arrays stay on the device
        # until somebody reads them (lazy attributes).  all_outputs: mag, direction, uca, twi + 3 masks are read too.
        e2e = {"value": cells_per_step / t_e2e / 1e6, "unit": "Mcells/s", "ms_per_step": t_e2e * 1e3,
               "h2d_bytes_per_step": int(n * n * 8 + 4 * n * 8),
               "d2h_bytes_per_step": int(n * n * 8),
               "what": "DEMProcessor(elev=<pinned host array>).calc_twi() -> twi on the host",
               "all_outputs": {"value": cells_per_step / t_all / 1e6, "unit": "Mcells/s", "ms_per_step": t_all * 1e3,
                               "h2d_bytes_per_step": int(n * n * 8 + 4 * n * 8),
                               "d2h_bytes_per_step": int(n * n * (8 * 4 + 3)),
                               "what": "the same call + reading mag, direction, uca, flats, edge_todo, edge_done on the host"}}
    else:
        e2e = sh.e2e(args.steps)
    _note("e2e done")

    config4 = None
    if world == 1 and args.config4:
        dt.close()
        config4 = config4_single(args)
    if world > 1 and not args.no_config4:
        sh.close()
        config4 = config4_leg(args, world, rank)
        _note("config 4 done")
    if rank != 0:
        return
    ms_sweep = float(np.mean(sweep_kernel_ms)) if sweep_kernel_ms and sweep_kernel_ms[0] else float(np.mean(sweep_ms))
    cells_rank = cells_per_step / world
    achieved = cells_rank * BYTES_PER_CELL / (ms_sweep * 1e-3) / 1e9 if ms_sweep else None
    line = {
        "metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload, "parallelism": parallelism,
                   "l2": "inputs larger than L2 (elev %.0f MB, every intermediate field >= %.0f MB vs 126 MB L2)"
                         % (n * n * 8 / 1e6, n * n / 1e6)},
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": ("wl::k_worklist<DrainOp<0>> (UCA accumulation sweep)" if world == 1 else
                                "wl::k_worklist<DrainOp<3>> (UCA accumulation sweep, one sweep across the GPUs over NVLink peer memory)"),
                     "bound": "hbm",
                     "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                     "frac": (achieved / peak_gbs) if achieved else None,
                     "traffic": PROFILE_TRAFFIC.get(n), "peak_source": peak_src,
                     "algorithmic_bytes_per_cell": BYTES_PER_CELL, "cells_per_launch": int(cells_rank),
                     "ms_per_launch": ms_sweep,
                     "note": "dependency/latency-bound graph sweep; traffic = dram bytes of one ncu --set full capture"},
        "stages": {k: stats.get(k) for k in ("ms_slopes", "ms_flats", "ms_graph", "ms_sweep", "ms_sweep_scan",
                                              "ms_sweep_kernel", "ms_sweep_first", "ms_sweep_resume", "ms_exchange",
                                              "ms_finalize_twi", "sweep_rounds", "label_rounds", "n_sources",
                                              "n_queue_items", "n_drained", "n_undone", "n_edge_todo") if k in stats},
    }
    if world == 1:
        ms_sl, ms_u, ms_t = (float(np.mean(stage_ms[k])) for k in ("slopes", "uca", "twi"))
        # BASELINE.json's metric names the stages separately: slope+aspect (calc_slopes_directions incl. the flat
        # mask), UCA (graph + pit drains + sweep), TWI -- device-resident, CUDA events in the timed loop
        line["per_stage"] = {
            "slope_aspect": {"Mcells_s": cells_per_step / ms_sl / 1e3, "ms": ms_sl,
                             "hbm_GBs_algorithmic": cells_per_step * 25.0 / (ms_sl * 1e-3) / 1e9,
                             "frac_of_hbm_peak": cells_per_step * 25.0 / (ms_sl * 1e-3) / 1e9 / peak_gbs,
                             "note": "25 B/cell: read elev 8, write mag 8 + direction 8 + flats 1; k_slopes is fp64-issue bound"},
            "uca": {"Mcells_s": cells_per_step / ms_u / 1e3, "ms": ms_u,
                    "hbm_GBs_algorithmic": cells_per_step * BYTES_PER_CELL / (ms_u * 1e-3) / 1e9,
                    "frac_of_hbm_peak": cells_per_step * BYTES_PER_CELL / (ms_u * 1e-3) / 1e9 / peak_gbs},
            "twi": {"Mcells_s": cells_per_step / ms_t / 1e3, "ms": ms_t,
                    "hbm_GBs_algorithmic": cells_per_step * 32.0 / (ms_t * 1e-3) / 1e9,
                    "frac_of_hbm_peak": cells_per_step * 32.0 / (ms_t * 1e-3) / 1e9 / peak_gbs,
                    "note": "32 B/cell: read uca 8 + mag 8, write twi 8 + 10*twi 8"},
        }
        line["cpu_baseline"], oracle_out = cpu_baseline_leg(E, drain_pits=bool(pits_flag))
        if gpu_out is not None and oracle_out is not None:
            # the oracle has just computed the whole benchmark DEM for the baseline: check the GPU arm against it
            line["parity"] = parity_block(oracle_out, gpu_out)
            line["parity"]["against"] = "oracle (oracle/pdm_oracle.c) on the whole benchmark DEM, same flags"
    else:
        line["parity"] = shard_parity
    if config4 is not None:
        line["config4"] = config4
    print(json.dumps(line), flush=True)
    if line.get("parity") and not line["parity"].get("ok", True):
        print("bench.py: PARITY FAILURE: %s" % json.dumps(line["parity"]), file=sys.stderr, flush=True)
        sys.exit(3)


def config4_single(args):
    """The per-GPU share of config 4 (one shard's worth: config4_rows x config4_cols of the same value-noise terrain,
    rows 0..) as ONE stand-alone tile on one GPU -- the N = 1 point of that shape's weak scaling (opt-in: --config4)."""
    import torch
    from pydem_b200 import synth, tile as T
    rows, cols = args.config4_rows, args.config4_cols
    t0 = time.perf_counter()
    dt = T.DeviceTile(rows, cols, stream=torch.cuda.current_stream().cuda_stream)
    dt.set_spacing(SPACING, SPACING)
    synth.value_noise_dem_torch(dt.as_torch(T.F_ELEV), 0, seed=11)
    dt.mark_resident(T.F_ELEV)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t0
    stats = {}

    def step():
        dt.slopes_directions()
        stats.update(dt.uca(drain_pits=0))
        dt.twi()
    step()
    torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.config4_steps):
        step()
    ev1.record(); torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / args.config4_steps
    uca = dt.as_torch(T.F_UCA); fl = dt.as_torch(T.F_FLATS)
    bad = [int(stats.get("n_undone", 0)), int((torch.isnan(uca) != (fl != 0)).sum().item()),
           int((uca[~torch.isnan(uca)] < SPACING * SPACING).sum().item())]
    out = {"workload": "%dx%d value-noise DEM (8 octaves, seed 11; the first shard of the config-4 terrain) as one stand-alone tile, "
                       "dX=dY=30 m, slope+aspect + UCA + TWI, fill_flats=False, drain_pits_path=False, drain_pits=False" % (rows, cols),
           "value": rows * cols / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": ms, "steps": args.config4_steps, "n_gpus": 1,
           "cells": rows * cols, "seconds_setup": t_gen,
           "stages": {k: stats.get(k) for k in ("ms_graph", "ms_sweep", "ms_sweep_kernel", "ms_sweep_scan", "n_sources", "n_queue_items")},
           "checks": {"undone_cells": bad[0], "nan_pattern_vs_flats_mismatches": bad[1], "uca_below_one_cell": bad[2], "ok": sum(bad) == 0}}
    dt.close()
    return out


def config4_leg(args, world, rank):
    """BASELINE.json configs[3] at its own shape: ONE non-periodic fractal (value-noise) DEM of
    (8192 * N) x 65536 cells, 8192 x 65536 per GPU (N = 8: 65536 x 65536), row-sharded, UCA as one sweep across
    the GPUs.  Device-resident timing of a few steps + checks that need no second implementation at that size:
    a 1024 x 1024 window of every rank's rows against the oracle (stencil outputs are local: mag bit-exact,
    direction to 1e-12; flats exact away from the window edge), every cell drained, NaN exactly on flats,
    min(uca) = one cell.  The reference cannot run this size at all (32-bit indices, cyutils.pyx:27)."""
    import torch
    import torch.distributed as dist
    from pydem_b200 import sharded, tile as T
    rows, cols = args.config4_rows, args.config4_cols
    t0 = time.perf_counter()
    sh = sharded.ShardedDEM(rows_per_rank=rows, cols=cols, spacing=SPACING, seed=11, profile=True, noise=True)
    t_gen = time.perf_counter() - t0
    stats = {}
    for _ in range(1):
        stats.update(sh.step())
    dist.barrier(); torch.cuda.synchronize()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    steps = args.config4_steps
    ev0.record()
    for _ in range(steps):
        stats.update(sh.step())
    ev1.record()
    dist.barrier(); torch.cuda.synchronize()
    tms = torch.tensor([ev0.elapsed_time(ev1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item()) / steps
    # ---- checks
    s = sh.spec
    tile = sh.engine.tile
    uca = sh.engine.rows(T.F_UCA)[s.lo:s.hi]
    fl = sh.engine.rows(T.F_FLATS)[s.lo:s.hi]
    bad = torch.zeros(4, dtype=torch.int64, device="cuda")
    bad[0] = int(stats.get("n_undone", 0))
    bad[1] = int((torch.isnan(uca) != (fl != 0)).sum().item())
    bad[2] = int((uca[~torch.isnan(uca)] < SPACING * SPACING).sum().item())
    wn = min(1024, s.r1 - s.r0 - 2, cols)
    wi, wj = s.lo + (s.r1 - s.r0 - wn) // 2, (cols - wn) // 2
    Ew = sh.engine.rows(T.F_ELEV)[wi:wi + wn, wj:wj + wn].cpu().numpy()
    from oracle import oracle as orc
    dp = orc.OracleDEMProcessor(np.ascontiguousarray(Ew), dX=SPACING, dY=SPACING, fill_flats=False, drain_pits_path=False, drain_pits=False)
    dp.calc_slopes_directions()
    mg = sh.engine.rows(T.F_MAG)[wi:wi + wn, wj:wj + wn].cpu().numpy()[2:-2, 2:-2]
    dr = sh.engine.rows(T.F_DIR)[wi:wi + wn, wj:wj + wn].cpu().numpy()[2:-2, 2:-2]
    # interior of the window only; pits the oracle would drain are excluded by comparing where both are defined
    om, od = np.asarray(dp.mag)[2:-2, 2:-2], np.asarray(dp.direction)[2:-2, 2:-2]
    both = (om > 0) & (mg > 0)
    with np.errstate(invalid="ignore"):
        bad[3] = int((np.abs(mg[both] - om[both]) > 1e-12 * np.abs(om[both])).sum() + (np.abs(dr[both] - od[both]) > 1e-12).sum())
    frac_checked = float(both.mean())
    dist.all_reduce(bad)
    cells = rows * cols * world
    out = {"workload": "one %dx%d value-noise DEM (8 octaves, seed 11, non-periodic, generated on the device), %d rows x %d columns per GPU, "
                       "dX=dY=30 m, slope+aspect + UCA + TWI, fill_flats=False, drain_pits_path=False, drain_pits=False"
                       % (rows * world, cols, rows, cols),
           "value": cells / (ms * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": ms, "steps": steps, "n_gpus": world,
           "cells": cells, "bytes_resident_per_gpu": int(93 * rows * cols), "seconds_setup": t_gen,
           "stages": {k: stats.get(k) for k in ("ms_slopes", "ms_flats", "ms_graph", "ms_sweep_first", "ms_finalize_twi", "sweep_rounds",
                                                 "label_rounds", "n_sources", "n_queue_items") if k in stats},
           "checks": {"undone_cells": int(bad[0]), "nan_pattern_vs_flats_mismatches": int(bad[1]), "uca_below_one_cell": int(bad[2]),
                      "window_1024_vs_oracle_stencil_mismatches": int(bad[3]), "window_fraction_compared": frac_checked,
                      "ok": bool(int(bad.sum()) == 0)}}
    sh.close()
    return out


def sharded_check(sh, block, world, rank, pits_flag):
    """One-off check before the timed loop: the NCCL-sharded result of this rank's rows against the
    single-tile result of the whole (stacked) DEM computed on this rank's own GPU; all ranks must agree."""
    import torch
    import torch.distributed as dist
    from pydem_b200 import DEMProcessor, tile as T
    sh.step()
    s = sh.spec
    full = np.concatenate([block] * world, axis=0)
    dp = DEMProcessor(elev=full, dX=SPACING, dY=SPACING, fill_flats=False, drain_pits_path=False, drain_pits=bool(pits_flag))
    dp.calc_twi()
    out = {}
    ok = True
    for name, f, exact in (("mag", T.F_MAG, True), ("direction", T.F_DIR, True), ("flats", T.F_FLATS, True),
                           ("edge_todo", T.F_EDGE_TODO, True), ("edge_done", T.F_EDGE_DONE, True), ("uca", T.F_UCA, False)):
        a = sh.engine.tile.download(f)[s.lo:s.hi]
        b = np.asarray(getattr(dp, name))[s.r0:s.r1]
        if exact:
            bad = int((a.astype(bool) != b).sum()) if a.dtype.kind != "f" else int(((a != b) & ~(np.isnan(a) & np.isnan(b))).sum())
        else:
            with np.errstate(invalid="ignore", divide="ignore"):
                rel = np.abs(a - b) / np.abs(b)
            bad = int((rel[np.isfinite(rel)] > PARITY_BAR["uca_rel"]).sum()) + int((np.isnan(a) != np.isnan(b)).sum())
        out[name + "_mismatches"] = bad
        ok = ok and bad == 0
    dp._free_tile()
    tot = torch.tensor([out[k] for k in sorted(out)], device="cuda", dtype=torch.int64)
    dist.all_reduce(tot)
    res = {k: int(v) for k, v in zip(sorted(out), tot.tolist())}
    res["ok"] = all(v == 0 for v in res.values())
    res["against"] = ("single-tile run of the whole %dx%d DEM on every rank's own GPU (uca rtol %g, everything else bit-exact), "
                      "summed over %d ranks" % (full.shape[0], full.shape[1], PARITY_BAR["uca_rel"], world))
    return res


# dram__bytes_read.sum + dram__bytes_write.sum of the sweep kernel, one launch (profiles/, per size)
PROFILE_TRAFFIC = {4096: 1.3952e9}   # profiles/r2_n1_kernels_4096_conditioned_final_ncu_full.csv: 868.0 MB read + 527.2 MB written


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=4096, help="rows (per GPU) = columns of the DEM")
    ap.add_argument("--variant", default="conditioned", choices=["conditioned", "sinks"],
                    help="conditioned: priority-flood conditioned fractal, default flags (BASELINE.md primary); "
                         "sinks: raw fractal with drain_pits=False")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle / single-tile check of the GPU arm's outputs")
    ap.add_argument("--no-config4", action="store_true", help="N > 1: skip the extra run of BASELINE configs[3] at its own shape")
    ap.add_argument("--config4", action="store_true", help="N = 1: also run one GPU's share of config 4 as a stand-alone tile")
    ap.add_argument("--config4-rows", type=int, default=8192, help="rows per GPU of the config-4 run")
    ap.add_argument("--config4-cols", type=int, default=65536)
    ap.add_argument("--config4-steps", type=int, default=3)
    ap.add_argument("--no-whole-dem", action="store_true", help="reference arm: skip the one-core whole-DEM run")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
