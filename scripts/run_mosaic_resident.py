"""BASELINE.json configs[4] at its own size: a ny x nx mosaic of T x T tiles (default 4 x 4 of 8192 x 8192,
2 px overlap) through the device-resident orchestrator (pydem_b200.process_manager.ResidentProcessManager: the
reference ProcessManager's stages and its serial process_uca_edges decisions, tiles resident in HBM, only edge
rings / strips moving).  The tiles are cut from one value-noise terrain, every rank generates its own tiles on
its GPU.  `--host` runs the host-side manager (every call round-trips the tile) for comparison, `--cond N` cuts
the tiles from the N x N conditioned fractal instead (the round-1 miniature: 4096, 4 x 4).
    python scripts/run_mosaic_resident.py [tile=8192] [grid=4] [overlap=2] [--host] [--cond 4096]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_mosaic_resident.py
Writes gpurun_out/mosaic_resident[_nN].json."""
import json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pydem_b200 import synth, _lib
from pydem_b200.process_manager import ProcessManager, ResidentProcessManager, TorchGroup, split_mosaic

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
_lib.init(local)
group = None
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = TorchGroup()
pos = [a for a in sys.argv[1:] if not a.startswith("--") and not (sys.argv[sys.argv.index(a) - 1] == "--cond")]
T_ = int(pos[0]) if len(pos) > 0 else 8192
grid = int(pos[1]) if len(pos) > 1 else 4
ov = int(pos[2]) if len(pos) > 2 else 2
host = "--host" in sys.argv
cond = int(sys.argv[sys.argv.index("--cond") + 1]) if "--cond" in sys.argv else 0
n = cond if cond else T_ * grid
boxes = split_mosaic((n, n), grid, grid, ov)
kw = dict(fill_flats=False, drain_pits_path=False)
sp = dict(dX=30.0, dY=30.0)
t0 = time.perf_counter()
tiles = []
Efull = synth.conditioned_fractal_dem(cond, 0) if cond else None
for k, b in enumerate(boxes):
    if k % world != rank:
        tiles.append(None)
    elif cond:
        tiles.append(np.ascontiguousarray(Efull[b[0]:b[1], b[2]:b[3]]))
    else:
        # rows b[0]..b[1] of the terrain, all columns of the tile: the generator takes a row offset; the column offset is a shift of the seed lattice
        out = torch.empty((b[1] - b[0], n), dtype=torch.float64, device="cuda") if n <= 16384 else None
        if out is not None:
            synth.value_noise_dem_torch(out, b[0], seed=5)
            tiles.append(out[:, b[2]:b[3]].contiguous().cpu().numpy())
        else:
            rows = []
            for a in range(b[0], b[1], 1024):
                o = torch.empty((min(1024, b[1] - a), n), dtype=torch.float64, device="cuda")
                synth.value_noise_dem_torch(o, a, seed=5)
                rows.append(o[:, b[2]:b[3]].contiguous().cpu().numpy())
            tiles.append(np.concatenate(rows, axis=0))
        del out
torch.cuda.empty_cache()
t_gen = time.perf_counter() - t0
res = dict(workload="%dx%d mosaic of %d tiles of ~%dx%d (%d px overlap) cut from %s, dX=dY=30 m, fill_flats=False, drain_pits_path=False, "
                    "drain_pits=True; %s orchestrator" % (grid, grid, len(boxes), boxes[0][1] - boxes[0][0], boxes[0][3] - boxes[0][2], ov,
                                                          ("the %dx%d conditioned fractal" % (cond, cond)) if cond else ("one %dx%d value-noise terrain" % (n, n)),
                                                          "host-side" if host else "device-resident"),
           cells=int(n) * int(n), n_gpus=world, seconds_generate=t_gen)
PM = ProcessManager if host else ResidentProcessManager
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    pm = PM(tiles, boxes, spacing=sp, dem_proc_kwargs=kw, group=group)
    pm.success[:, 0] = True     # elevation is taken as it is (conditioning stage off)
    t = {}
    for name, fn in (("aspect_slope", pm.process_aspect_slope), ("uca", pm.process_uca), ("uca_edges", pm.process_uca_edges),
                     ("twi", lambda: pm._stage(pm._twi, 3))):
        if group is not None: dist.barrier()
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        if group is not None: dist.barrier()
        t[name] = (time.perf_counter() - t0) * 1e3
res["ms"] = t
res["ms_total"] = sum(t.values())
res["Mcells_s"] = res["cells"] / res["ms_total"] / 1e3
res["corrections"] = len(pm.correction_log)
if not host:
    res["bytes_moved"] = pm.bytes_moved
if cond and world == 1:
    from pydem_b200 import DEMProcessor
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        dp = DEMProcessor(elev=Efull, dX=30.0, dY=30.0, **kw)
        dp.calc_twi()
    m = pm.mosaic("uca")
    with np.errstate(invalid="ignore", divide="ignore"):
        rel = np.abs(m - dp.uca) / np.abs(dp.uca)
    res["mosaic_vs_single_dem_share_within_1e-9"] = float((rel[np.isfinite(rel)] <= 1e-9).mean())
if rank == 0:
    out = os.path.join(ROOT, "gpurun_out", "mosaic_%s%s%s.json" % ("host" if host else "resident", "_cond%d" % cond if cond else "_%dx%dx%d" % (grid, grid, T_),
                                                                       "" if world == 1 else "_n%d" % world))
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res))
if group is not None:
    dist.barrier()
    dist.destroy_process_group()
