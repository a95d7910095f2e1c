"""Row-sharded execution of the hot path over the GPUs of one box.

One DEM of ``R x C`` cells is cut into contiguous row blocks, one per rank (one process per
GPU, ``torch.distributed`` / NCCL).  Every rank keeps one halo row per interior side; halo rows
travel between neighbouring ranks with NCCL send/recv (``Group.exchange``), everything else is
local CUDA work through the ``pdm_shard_*`` C ABI (csrc/shard.cu):

    elev halo -> slope/aspect (a1) -> flat0 halo -> region labels (+ label rounds) -> flats (a2)
    -> links (a3/a4) -> link halo -> inflow mask, sweep records -> record halo
    -> { local sweep until it rests ; exchange boundary rows of sweep records } until nobody sent
    -> finalize -> TWI

This is pyDEM's cross-tile UCA edge resolution (reference process_manager.py:1090-1249:
tiles exchange edge strips and re-run calc_uca on the deltas until nothing changes) in its
gating-free form: a boundary cell pulls the contributions of its donors on the neighbouring
rank from the halo row once they are final there ("not done" travels as a NaN bit pattern).
Because every owned cell is computed from its true 3x3 neighbourhood, in a fixed summation
order, and "border" means the border of the whole DEM, the sharded result equals the
single-tile result bit for bit.

``Group`` hides where the ranks live: ``DistGroup`` = this process is one rank of a
torch.distributed job; ``LocalGroup`` = all ranks live in this process (tests on one GPU, and
the CPU tests of the exchange logic).
"""
import numpy as np

from . import synth


# ----------------------------------------------------------------------------------------------
# partition
# ----------------------------------------------------------------------------------------------
def row_blocks(R, world):
    """Contiguous, near-equal row blocks [r0, r1) per rank (every rank gets >= 2 rows)."""
    if R < 2 * world:
        raise ValueError("need at least 2 rows per rank")
    edges = [(R * k) // world for k in range(world + 1)]
    return [(edges[k], edges[k + 1]) for k in range(world)]


def global_row_theta(dX, dY):
    """theta of _calc_uca_section_proportion for the whole grid (dem_processing.py:1031-1033)."""
    th = np.arctan2(np.asarray(dY, "float64"), np.asarray(dX, "float64"))
    R = th.size + 1
    idx = np.clip(np.arange(R) - 1, 0, R - 3)
    return th[idx]


class ShardSpec(object):
    """Geometry of one rank's tile: owned global rows [r0, r1), local rows incl. halos."""

    def __init__(self, R, C, rank, world):
        self.R, self.C, self.rank, self.world = R, C, rank, world
        self.r0, self.r1 = row_blocks(R, world)[rank]
        self.halo_top = 1 if rank > 0 else 0
        self.halo_bot = 1 if rank < world - 1 else 0
        self.row_off = self.r0 - self.halo_top          # global row of local row 0
        self.Rl = (self.r1 - self.r0) + self.halo_top + self.halo_bot
        self.lo = self.halo_top
        self.hi = self.lo + (self.r1 - self.r0)

    def local_slice(self):
        """global rows held locally (owned + halo)"""
        return slice(self.row_off, self.row_off + self.Rl)


# ----------------------------------------------------------------------------------------------
# groups: neighbour exchange + sum-allreduce
# ----------------------------------------------------------------------------------------------
class LocalGroup(object):
    """All ranks in this process.  ``exchange`` takes one dict per rank."""

    def __init__(self, world):
        self.world = world
        self.local_ranks = list(range(world))

    def exchange(self, bufs):
        """bufs[k] = dict(send_up, send_down, recv_up, recv_down) of equally shaped tensors/arrays
        (None where there is no neighbour).  Rank k's send_down lands in rank k+1's recv_up."""
        for k in range(self.world - 1):
            _copy(bufs[k + 1]["recv_up"], bufs[k]["send_down"])
            _copy(bufs[k]["recv_down"], bufs[k + 1]["send_up"])

    def allreduce_sum(self, values):
        tot = sum(int(v) for v in values)
        return [tot] * len(values)


def _copy(dst, src):
    if dst is None or src is None:
        raise ValueError("exchange buffers missing on an interior boundary")
    if hasattr(dst, "copy_"):
        dst.copy_(src)
    else:
        dst[...] = src


class DistGroup(object):
    """This process is one rank of a torch.distributed job (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.local_ranks = [self.rank]

    def exchange(self, bufs):
        dist = self.dist
        b = bufs[0]
        ops = []
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, b["send_up"], self.rank - 1))
            ops.append(dist.P2POp(dist.irecv, b["recv_up"], self.rank - 1))
        if self.rank < self.world - 1:
            ops.append(dist.P2POp(dist.isend, b["send_down"], self.rank + 1))
            ops.append(dist.P2POp(dist.irecv, b["recv_down"], self.rank + 1))
        if ops:
            for r in dist.batch_isend_irecv(ops):   # one ncclGroupStart/End with both neighbours
                r.wait()

    def allreduce_sum(self, values):
        import torch
        v = values[0]
        t = v if hasattr(v, "device") else torch.tensor([int(v)], dtype=torch.int64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(t.item())]

    def connect_p2p(self, engine):
        """Map the row neighbours' sweep records and tile queues (CUDA IPC over NVLink) so that the UCA
        accumulation is ONE sweep across all ranks instead of exchange rounds (csrc/shard.cu,
        pdm_shard_p2p_*).  Collective: every rank calls it once after its tile window is set."""
        import torch
        blob = engine.p2p_export()
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
        allb = [torch.empty_like(mine) for _ in range(self.world)]
        self.dist.all_gather(allb, mine)
        blobs = [bytes(b.cpu().numpy().tobytes()) for b in allb]
        engine.p2p_connect_all(blobs, self.world, self.rank)
        self.dist.barrier()


# ----------------------------------------------------------------------------------------------
# GPU engine of one rank
# ----------------------------------------------------------------------------------------------
class ShardEngine(object):
    """One rank's tile on one GPU."""

    def __init__(self, spec, dX, dY, dX2, dY2, device=None, stream=None):
        import torch
        from . import tile as T
        self.T, self.torch, self.spec = T, torch, spec
        s = spec
        self.tile = T.DeviceTile(s.Rl, s.C, device=device, stream=stream)
        fence = slice(s.row_off, s.row_off + s.Rl - 1)
        self.tile.set_spacing(np.asarray(dX)[fence], np.asarray(dY)[fence], np.asarray(dX2)[s.local_slice()],
                              np.asarray(dY2)[s.local_slice()])
        self.tile.set_window(s.row_off, s.R, s.lo, s.hi, global_row_theta(dX, dY)[s.local_slice()])
        self.tile.min_area = float(np.nanmin(np.asarray(dX2) * np.asarray(dY2)))   # of the whole DEM (898)
        dev = torch.device("cuda", torch.cuda.current_device())
        C = s.C
        mk = lambda dt: torch.zeros(C, dtype=dt, device=dev)
        # packed exchange buffers of the region labels (int64 + f64)
        self.lab = {k: (mk(torch.int64), mk(torch.float64)) for k in ("send_up", "send_down", "recv_up", "recv_down")}
        self.flag = torch.zeros(1, dtype=torch.int64, device=dev)
        self.views = {}
        self.p2p = False
        self._dXY = (np.ascontiguousarray(dX, "float64"), np.ascontiguousarray(dY, "float64"))   # fences of the whole grid
        self._pit = None

    # ---- one sweep across GPUs (pdm_shard_p2p_*)
    def p2p_export(self):
        import ctypes as ct
        L, h = self.tile.L, self.tile.h
        size = ct.c_int64(0)
        self.T._lib.check(L.pdm_shard_p2p_export(h, None, ct.byref(size)))
        buf = (ct.c_ubyte * size.value)()
        self.T._lib.check(L.pdm_shard_p2p_export(h, buf, ct.byref(size)))
        return bytes(buf)

    def p2p_connect(self, up, down, root, world, rank):
        import ctypes as ct
        keep = [ct.create_string_buffer(b, len(b)) if b is not None else None for b in (up, down, root)]
        self.T._lib.check(self.tile.L.pdm_shard_p2p_connect(self.tile.h, *keep, int(world), int(rank)))
        self.p2p = True

    def p2p_connect_all(self, blobs, world, rank):
        """every rank's export: the work-list engine can then run the accumulation as one sweep across the GPUs"""
        import ctypes as ct
        joined = b"".join(blobs)
        buf = ct.create_string_buffer(joined, len(joined))
        self.T._lib.check(self.tile.L.pdm_shard_p2p_connect_all(self.tile.h, buf, int(world), int(rank)))
        self.p2p = True

    # ---- drain_pits on a row shard (pdm_shard_pits): strips of the neighbours' elevation / pit mask in, counts of the
    #      pit edges that end on a neighbour out
    def pit_setup(self):
        import ctypes as ct
        p = self.tile._shard_params
        H = int(p.drain_pits_max_iter) + 1
        Hin = min(H, int(p.drain_pits_max_dist)) if int(p.drain_pits_max_dist) > 0 else H
        if self._pit is not None and self._pit["H"] == H and self._pit["Hin"] == Hin:
            return self._pit
        s, torch = self.spec, self.torch
        if s.r1 - s.r0 < H:
            raise ValueError("drain_pits on row shards needs at least drain_pits_max_iter + 1 = %d rows per rank (have %d)"
                             % (H, s.r1 - s.r0))
        if self._pit is None:
            dX, dY = self._dXY
            self.T._lib.check(self.tile.L.pdm_tile_set_global_spacing(self.tile.h, self.T._lib.ptr(dX), self.T._lib.ptr(dY),
                                                                      ct.c_int64(dX.size)))
        dev = torch.device("cuda", torch.cuda.current_device())
        mk = lambda rows, dt, ok: torch.zeros((rows, s.C), dtype=dt, device=dev) if ok else None
        b = dict(H=H, Hin=Hin)
        for side, ok in (("up", s.halo_top), ("dn", s.halo_bot)):
            b["E_" + side] = mk(H, torch.float64, ok)
            b["P_" + side] = mk(H, torch.uint8, ok)
            b["in_" + side] = mk(Hin, torch.int32, ok)
            b["from_" + side] = mk(Hin, torch.int32, ok)
        self._pit = b
        return b

    def pit_bufs(self, what):
        """exchange buffers: 'E' / 'P' = my owned rows next to each boundary -> the neighbours' strips; 'in' = edge counts"""
        s, b = self.spec, self._pit
        if what == "in":
            return dict(send_up=b["in_up"], recv_up=b["from_up"], send_down=b["in_dn"], recv_down=b["from_dn"])
        v = self.rows(self.T.F_ELEV if what == "E" else self.T.F_FLAT0)
        H = b["H"]
        return dict(send_up=v[s.lo:s.lo + H] if s.halo_top else None, recv_up=b[what + "_up"],
                    send_down=v[s.hi - H:s.hi] if s.halo_bot else None, recv_down=b[what + "_dn"])

    def shard_pits(self):
        import ctypes as ct
        b = self._pit
        dp = lambda x: ct.c_void_p(x.data_ptr()) if x is not None else None
        Hu = b["H"] if b["E_up"] is not None else 0
        Hd = b["H"] if b["E_dn"] is not None else 0
        self.T._lib.check(self.tile.L.pdm_shard_pits(self.tile.h, ct.byref(self.tile._shard_params), dp(b["E_up"]), dp(b["P_up"]), Hu,
                                                     dp(b["E_dn"]), dp(b["P_dn"]), Hd, dp(b["in_up"]), dp(b["in_dn"]), b["Hin"]))

    def pit_in_apply(self):
        import ctypes as ct
        b = self._pit
        dp = lambda x: ct.c_void_p(x.data_ptr()) if x is not None else None
        self.T._lib.check(self.tile.L.pdm_shard_pit_in_apply(self.tile.h, dp(b["from_up"]), dp(b["from_dn"]), b["Hin"]))

    def rows(self, field):
        if field not in self.views:
            self.views[field] = self.tile.as_torch(field)
        return self.views[field]

    def halo_bufs(self, field):
        """Row views for a plain halo exchange of one field: I send my first/last owned row and
        receive into my halo rows."""
        s, v = self.spec, self.rows(field)
        return dict(send_up=v[s.lo] if s.halo_top else None, recv_up=v[s.lo - 1] if s.halo_top else None,
                    send_down=v[s.hi - 1] if s.halo_bot else None, recv_down=v[s.hi] if s.halo_bot else None)

    # packed label rows: [int64 labels | f64 elevations] travel as two tensors -> two exchanges
    def label_pack(self):
        s = self.spec
        if s.halo_top:
            self.tile.shard_stage("label_pack", s.lo, self.lab["send_up"][0].data_ptr(), self.lab["send_up"][1].data_ptr())
        if s.halo_bot:
            self.tile.shard_stage("label_pack", s.hi - 1, self.lab["send_down"][0].data_ptr(), self.lab["send_down"][1].data_ptr())

    def label_unpack(self):
        s = self.spec
        self.flag.zero_()
        if s.halo_top:
            self.tile.shard_stage("label_unpack", s.lo - 1, self.lab["recv_up"][0].data_ptr(), self.lab["recv_up"][1].data_ptr(),
                                  self.flag.data_ptr())
        if s.halo_bot:
            self.tile.shard_stage("label_unpack", s.hi, self.lab["recv_down"][0].data_ptr(), self.lab["recv_down"][1].data_ptr(),
                                  self.flag.data_ptr())
        return self.flag

    def sweep_sent(self):
        """device int64: boundary cells the last sweep completed whose receiver lives on a neighbour"""
        self.flag.zero_()
        self.tile.shard_stage("sweep_sent", self.flag.data_ptr())
        return self.flag

    def lab_bufs(self, which):
        s = self.spec
        g = lambda k, ok: self.lab[k][which] if ok else None
        return dict(send_up=g("send_up", s.halo_top), recv_up=g("recv_up", s.halo_top),
                    send_down=g("send_down", s.halo_bot), recv_down=g("recv_down", s.halo_bot))


class _Timer(object):
    """CUDA-event stage timer (only when profiling is requested)."""

    def __init__(self, on):
        self.on, self.marks = on, []
        if on:
            import torch
            self.torch = torch
            self.mark("start")

    def mark(self, name):
        if self.on:
            ev = self.torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def result(self):
        if not self.on:
            return {}
        self.torch.cuda.synchronize()
        out = {}
        for (n0, e0), (n1, e1) in zip(self.marks[:-1], self.marks[1:]):
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
        return out


def run_hot_path(engines, group, twi=True, profile=False, **uca_flags):
    """slope/aspect -> UCA -> TWI over all shards.  ``engines``: the engines of the ranks that
    live in this process (group.local_ranks).  Returns per-engine stats dicts."""
    T = engines[0].T
    tm = _Timer(profile)
    # a1: elevation halo, stencil
    group.exchange([e.halo_bufs(T.F_ELEV) for e in engines])
    for e in engines:
        e.tile.shard_stage("slopes")
    tm.mark("ms_slopes")
    # a2: flat0 halo, region labels (+ cross-rank label rounds), one-pixel extension
    group.exchange([e.halo_bufs(T.F_FLAT0) for e in engines])
    for e in engines:
        e.tile.shard_stage("ccl")
    label_rounds = 0
    if group.world > 1:
        while True:
            for e in engines:
                e.label_pack()
            group.exchange([e.lab_bufs(0) for e in engines])
            group.exchange([e.lab_bufs(1) for e in engines])
            changed = group.allreduce_sum([e.label_unpack() for e in engines])[0]
            label_rounds += 1
            if changed == 0:
                break
    for e in engines:
        e.tile.shard_stage("flats_extend")
    tm.mark("ms_flats")
    # a3/a4: links, link halo, in-degree + inflow-border mask
    for e in engines:
        e.tile.shard_links(**uca_flags)
    if uca_flags.get("drain_pits") and group.world > 1:
        # a5 across shards: a pit's search region / drains may lie on the neighbouring rank
        for e in engines:
            e.pit_setup()
        group.exchange([e.pit_bufs("E") for e in engines])
        group.exchange([e.pit_bufs("P") for e in engines])
        for e in engines:
            e.shard_pits()
        group.exchange([e.pit_bufs("in") for e in engines])
        for e in engines:
            e.pit_in_apply()
    group.exchange([e.halo_bufs(T.F_LINK) for e in engines])
    for e in engines:
        e.tile.shard_stage("indeg")
    p2p = all(e.p2p for e in engines)
    if not p2p:
        group.exchange([e.halo_bufs(T.F_CELL) for e in engines])      # the neighbours' boundary records (static part + "not done")
    tm.mark("ms_graph")
    # a6/a7: local sweeps + exchanges of the boundary rows -- or, with the neighbours' records mapped as peer
    # memory, ONE sweep across all ranks: no rounds, no host synchronisation
    rounds, first = 0, 1
    while True:
        if p2p:
            for e in engines:
                e.tile.shard_stage("sweep", 1)
            tm.mark("ms_sweep_first")
            rounds = 1
            break
        for e in engines:
            e.tile.shard_stage("sweep", first)
        tm.mark("ms_sweep_first" if first else "ms_sweep_resume")
        first = 0
        rounds += 1
        if group.world == 1:
            break
        sent = group.allreduce_sum([e.sweep_sent() for e in engines])[0]
        if sent == 0:
            tm.mark("ms_exchange")
            break
        group.exchange([e.halo_bufs(T.F_CELL) for e in engines])
        tm.mark("ms_exchange")
    stats = []
    for e in engines:
        st = e.tile.shard_finalize()
        st.update(sweep_rounds=rounds, label_rounds=label_rounds)
        stats.append(st)
        if twi:
            e.tile.twi()
    tm.mark("ms_finalize_twi")
    prof = tm.result()
    for st in stats:
        st.update(prof)
    return stats


# ----------------------------------------------------------------------------------------------
# convenience drivers
# ----------------------------------------------------------------------------------------------
def run_local(E, world, dX=1.0, dY=1.0, dX2=None, dY2=None, **uca_flags):
    """All ``world`` shards of ``E`` on the current GPU, one after the other (test helper and
    single-GPU fallback for DEMs that do not fit one 32-bit-indexed tile).  Returns a dict of
    assembled full-size arrays plus per-shard stats."""
    import torch
    from . import tile as T
    R, C = E.shape
    dXa = np.broadcast_to(np.asarray(dX, "float64"), (R - 1,)); dYa = np.broadcast_to(np.asarray(dY, "float64"), (R - 1,))
    dX2a = np.broadcast_to(np.asarray(dXa[0] if dX2 is None else dX2, "float64"), (R,))
    dY2a = np.broadcast_to(np.asarray(dYa[0] if dY2 is None else dY2, "float64"), (R,))
    group = LocalGroup(world)
    stream = torch.cuda.current_stream().cuda_stream
    engines = []
    for k in range(world):
        spec = ShardSpec(R, C, k, world)
        e = ShardEngine(spec, dXa, dYa, dX2a, dY2a, stream=stream)
        loc = np.full((spec.Rl, C), np.nan)
        loc[spec.lo:spec.hi] = E[spec.r0:spec.r1]            # halo rows arrive through the exchange
        e.tile.upload(T.F_ELEV, loc)
        engines.append(e)
    stats = run_hot_path(engines, group, **uca_flags)
    out = {}
    for name, f in (("mag", T.F_MAG), ("direction", T.F_DIR), ("flats", T.F_FLATS), ("uca", T.F_UCA), ("twi", T.F_TWI),
                    ("edge_todo", T.F_EDGE_TODO), ("edge_done", T.F_EDGE_DONE)):
        full = np.empty((R, C), dtype=T._lib.FIELD_DTYPE[f])
        for e in engines:
            s = e.spec
            full[s.r0:s.r1] = e.tile.download(f)[s.lo:s.hi]
        out[name] = full.astype(bool) if full.dtype == np.uint8 else full
    out["stats"] = stats
    for e in engines:
        e.tile.close()
    return out


class ShardedDEM(object):
    """bench.py's N>1 workload: a value-noise DEM of (rows_per_rank * world) x cols, one row block
    per rank of the torch.distributed job; ``step()`` runs the whole hot path once."""

    def __init__(self, rows_per_rank, cols, spacing=30.0, seed=0, profile=False, block=None, noise=False):
        """block: the rows every rank holds (periodic block, weak scaling with identical work per GPU);
        noise=True: rank r holds rows [r * rows_per_rank, ...) of ONE non-periodic value-noise DEM,
        generated on the device (BASELINE.json configs[3]: 65536 x 65536 over 8 GPUs = 8192 x 65536 each)."""
        import torch
        from . import tile as T
        self.T, self.torch = T, torch
        self.group = DistGroup()
        w, r = self.group.world, self.group.rank
        R = rows_per_rank * w
        self.spec = ShardSpec(R, cols, r, w)
        s = self.spec
        d = np.full(R - 1, float(spacing)); d2 = np.full(R, float(spacing))
        self.engine = ShardEngine(s, d, d, d2, d2, stream=torch.cuda.current_stream().cuda_stream)
        import os
        if w > 1 and os.environ.get("PYDEM_B200_SHARD_P2P", "1") != "0":
            self.group.connect_p2p(self.engine)
        self.profile = profile
        self.uca_flags = {}     # e.g. drain_pits=1 (needs the p2p connection)
        self.cells = (s.r1 - s.r0) * cols
        if noise:
            view = self.engine.rows(T.F_ELEV)                     # zero-copy torch view of the tile's elevation field
            view.fill_(float("nan"))                              # halo rows arrive through the exchange
            synth.value_noise_dem_torch(view[s.lo:s.hi], s.r0, seed=seed)
            self.engine.tile.mark_resident(T.F_ELEV)
            self.host_elev = None
            torch.cuda.synchronize()
            return
        loc = np.full((s.Rl, cols), np.nan)
        # every rank holds the same periodic block (spectral synthesis, conditioned with wrapping
        # rows): stacked vertically the blocks join seamlessly, so per-GPU work is identical (weak
        # scaling) and rivers really cross the shard boundaries
        if block is None:
            block = synth.conditioned_fractal_dem(rows_per_rank, seed, shape=(rows_per_rank, cols), wrap_rows=True)
        loc[s.lo:s.hi] = block
        self.host_elev = loc
        self.engine.tile.upload(T.F_ELEV, loc)

    def close(self):
        """collective: unmap the peers' memory on every rank before any rank frees its tile"""
        if self.engine.p2p:
            self.T._lib.check(self.engine.tile.L.pdm_shard_p2p_disconnect(self.engine.tile.h))
            self.engine.p2p = False
            self.torch.cuda.synchronize()
            self.group.dist.barrier()
        self.engine.tile.close()

    def step(self):
        return run_hot_path([self.engine], self.group, profile=self.profile, **self.uca_flags)[0]

    def e2e(self, steps):
        """Same metric end to end: the rank's rows uploaded from pinned host memory and the RESULT (twi, what
        DEMProcessor.calc_twi returns -- the 1-GPU e2e figure) read back inside the timed region; `all_outputs` adds
        mag, direction, uca, flats and the edge masks, like the 1-GPU line's variant of the same name."""
        import time
        from . import _pinned
        torch, T, s = self.torch, self.T, self.spec
        Eh = _pinned.pinned_copy(self.host_elev)
        outs = {f: _pinned.empty((s.Rl, s.C), T._lib.FIELD_DTYPE[f])
                for f in (T.F_TWI, T.F_MAG, T.F_DIR, T.F_UCA, T.F_FLATS, T.F_EDGE_TODO, T.F_EDGE_DONE)}
        dist = self.group.dist

        def one(fields):
            self.engine.tile.upload(T.F_ELEV, Eh)
            run_hot_path([self.engine], self.group, **self.uca_flags)
            for f in fields:
                self.engine.tile.download(f, outs[f])

        def timed(fields, n):
            one(fields)
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n):
                one(fields)
            torch.cuda.synchronize(); dist.barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / n], dtype=torch.float64, device="cuda")
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item())
        w = self.group.world
        cells = self.cells * w
        dt = timed([T.F_TWI], steps)
        dt_all = timed(list(outs), max(2, steps // 2))
        return {"value": cells / dt / 1e6, "unit": "Mcells/s", "ms_per_step": dt * 1e3,
                "h2d_bytes_per_step": int(s.Rl * s.C * 8 * w), "d2h_bytes_per_step": int(s.Rl * s.C * 8 * w),
                "what": "per rank: rows uploaded from pinned host memory -> sharded pass -> twi on the host",
                "all_outputs": {"value": cells / dt_all / 1e6, "unit": "Mcells/s", "ms_per_step": dt_all * 1e3,
                                "h2d_bytes_per_step": int(s.Rl * s.C * 8 * w),
                                "d2h_bytes_per_step": int(s.Rl * s.C * (8 * 4 + 3) * w),
                                "what": "the same + mag, direction, uca, flats, edge_todo, edge_done on the host"}}
