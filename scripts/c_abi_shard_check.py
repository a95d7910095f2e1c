"""The row-sharded hot path through the C ABI alone (no torch, no torch.distributed): N processes, one per GPU,
drive pdm_comm_* / pdm_shard_connect / pdm_shard_run with ctypes and compare their rows with the single-tile
result of the whole DEM.  This is the host a C / Fortran program would write.

    python scripts/c_abi_shard_check.py [world] [rows] [cols] [pits]

The parent only spawns the ranks and hands rank 0's NCCL id to the others through a file."""
import ctypes as ct
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def child(rank, world, R, C, pits, idfile):
    assert "torch" not in sys.modules
    from pydem_b200 import _lib, synth, sharded
    L = _lib.load()
    _lib.init(rank)            # pdm_init(rank): this process drives GPU `rank`
    ident = (ct.c_ubyte * 128)()
    if rank == 0:
        _lib.check(L.pdm_comm_unique_id(ident))
        with open(idfile + ".tmp", "wb") as f:
            f.write(bytes(ident))
        os.rename(idfile + ".tmp", idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 60:
                raise RuntimeError("no NCCL id from rank 0")
            time.sleep(0.05)
        ident = (ct.c_ubyte * 128).from_buffer_copy(open(idfile, "rb").read())
    _lib.check(L.pdm_comm_init(rank, world, ident))
    E = synth.value_noise_dem(0, R, C, seed=7)
    E[R // 2 - 3:R // 2 + 3, 100:400] = E[R // 2 - 3:R // 2 + 3, 100:400].min()     # a lake on a shard boundary (world 2)
    spec = sharded.ShardSpec(R, C, rank, world)
    d = np.full(R - 1, 30.0); d2 = np.full(R, 30.0)
    h = ct.c_void_p()
    _lib.check(L.pdm_tile_create(spec.Rl, C, None, ct.byref(h)))
    fence = slice(spec.row_off, spec.row_off + spec.Rl - 1)
    loc = spec.local_slice()
    dXl, dYl, dX2l, dY2l = d[fence].copy(), d[fence].copy(), d2[loc].copy(), d2[loc].copy()
    thA, thB = np.ascontiguousarray(np.arctan2(dYl, dXl)), np.ascontiguousarray(np.arctan2(dXl, dYl))
    _lib.check(L.pdm_tile_set_spacing(h, _lib.ptr(dXl), _lib.ptr(dYl), _lib.ptr(dX2l), _lib.ptr(dY2l), _lib.ptr(thA), _lib.ptr(thB)))
    th = sharded.global_row_theta(d, d)[loc].copy()
    _lib.check(L.pdm_tile_set_window(h, spec.row_off, R, spec.lo, spec.hi, _lib.ptr(th)))
    _lib.check(L.pdm_tile_set_global_spacing(h, _lib.ptr(d), _lib.ptr(d), ct.c_int64(d.size)))
    _lib.check(L.pdm_shard_connect(h))
    elev = np.full((spec.Rl, C), np.nan)
    elev[spec.lo:spec.hi] = E[spec.r0:spec.r1]
    _lib.check(L.pdm_tile_upload(h, _lib.F_ELEV, _lib.ptr(elev)))
    p = _lib.UcaParams(); L.pdm_default_uca_params(ct.byref(p)); p.drain_pits = int(pits)
    tw = _lib.TwiParams(); L.pdm_default_twi_params(ct.byref(tw))
    tw.twi_min_area = 900.0
    st = _lib.UcaStats(); rounds = ct.c_int(0)
    ms = []
    for _ in range(3):
        _lib.check(L.pdm_comm_barrier(h))
        t0 = time.perf_counter()
        _lib.check(L.pdm_shard_run(h, ct.byref(p), ct.byref(tw), ct.byref(st), ct.byref(rounds)))
        _lib.check(L.pdm_tile_sync(h))
        ms.append((time.perf_counter() - t0) * 1e3)
    got = {}
    for name, f in (("mag", _lib.F_MAG), ("direction", _lib.F_DIR), ("flats", _lib.F_FLATS), ("uca", _lib.F_UCA),
                    ("edge_todo", _lib.F_EDGE_TODO), ("edge_done", _lib.F_EDGE_DONE), ("twi", _lib.F_TWI)):
        a = np.empty((spec.Rl, C), _lib.FIELD_DTYPE[f])
        _lib.check(L.pdm_tile_download(h, f, _lib.ptr(a)))
        got[name] = a[spec.lo:spec.hi]
    _lib.check(L.pdm_shard_disconnect(h))
    _lib.check(L.pdm_tile_destroy(h))
    # the single-tile answer on this rank's own GPU (public operator; it uses the C ABI too)
    import warnings
    from pydem_b200 import DEMProcessor
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dp = DEMProcessor(elev=E, dX=30.0, dY=30.0, fill_flats=False, drain_pits_path=False, drain_pits=bool(pits))
        ref_twi = dp.calc_twi()
    ok = True
    for name, exact in (("mag", True), ("direction", True), ("flats", True), ("edge_todo", True), ("edge_done", True),
                        ("uca", False), ("twi", False)):
        b = np.asarray(ref_twi if name == "twi" else getattr(dp, name))[spec.r0:spec.r1]
        a = got[name]
        if exact:
            good = np.array_equal(a, b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a.astype(bool), b)
        else:
            good = np.allclose(a, b, rtol=1e-9, atol=0, equal_nan=True)
        if not good:
            print("rank", rank, "MISMATCH", name, flush=True)
        ok = ok and good
    _lib.check(L.pdm_comm_finalize())
    print("c_abi_shard_check rank %d/%d %s rows %d cols %d drain_pits %s label_rounds %d ms/pass %.2f drained %d undone %d torch_loaded %s"
          % (rank, world, "OK" if ok else "FAILED", R, C, bool(pits), rounds.value, min(ms), st.n_drained, st.n_undone,
             "torch" in sys.modules), flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), sys.argv[7])
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    C = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    pits = 1 if (len(sys.argv) > 4 and sys.argv[4] == "pits") else 0
    idfile = "/tmp/pdm_nccl_id_%d" % os.getpid()
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--child", str(r), str(world), str(R), str(C), str(pits), idfile])
             for r in range(world)]
    rcs = []
    for pr in procs:
        try:
            rcs.append(pr.wait(timeout=240))
        except subprocess.TimeoutExpired:
            pr.kill(); rcs.append(-9)
    if os.path.exists(idfile):
        os.remove(idfile)
    print("c_abi_shard_check", "OK" if all(r == 0 for r in rcs) else "FAILED %s" % rcs, flush=True)
    sys.exit(0 if all(r == 0 for r in rcs) else 1)
