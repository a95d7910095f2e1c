"""Multi-GPU correctness under torchrun: the NCCL-sharded hot path vs the single-tile result.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/dist_check.py [rows cols]
Every rank computes the single-tile answer for the whole DEM on its own GPU and compares its
owned rows of the sharded run against it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from pydem_b200 import _lib, sharded, synth, tile as T, DEMProcessor
_lib.init(local)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
REPS = int(sys.argv[4]) if len(sys.argv) > 4 else 1     # stress: repeat the sharded pass, every result must equal the first
PITS = len(sys.argv) > 3 and sys.argv[3] == "pits"     # drain_pits=True (the reference default): pits search and drain across the shard boundaries
g = sharded.DistGroup()
E = synth.value_noise_dem(0, R, C, seed=7)
E[R // 2 - 3:R // 2 + 3, 100:400] = E[R // 2 - 3:R // 2 + 3, 100:400].min()     # a lake on a shard boundary
spec = sharded.ShardSpec(R, C, g.rank, g.world)
d = np.full(R - 1, 30.0); d2 = np.full(R, 30.0)
eng = sharded.ShardEngine(spec, d, d, d2, d2, stream=torch.cuda.current_stream().cuda_stream)
if g.world > 1 and os.environ.get("PYDEM_B200_SHARD_P2P", "1") != "0":
    g.connect_p2p(eng)      # one sweep across the GPUs (peer memory) instead of exchange rounds
loc = np.full((spec.Rl, C), np.nan); loc[spec.lo:spec.hi] = E[spec.r0:spec.r1]
eng.tile.upload(T.F_ELEV, loc)
st = sharded.run_hot_path([eng], g, drain_pits=1)[0] if PITS else sharded.run_hot_path([eng], g)[0]
first = eng.tile.download(T.F_UCA)[spec.lo:spec.hi].copy()
stress_bad = 0
for rep in range(1, REPS):
    eng.tile.upload(T.F_ELEV, loc)
    st = sharded.run_hot_path([eng], g, drain_pits=1)[0] if PITS else sharded.run_hot_path([eng], g)[0]
    u = eng.tile.download(T.F_UCA)[spec.lo:spec.hi]
    if st["n_undone"] != 0 or not np.allclose(u, first, rtol=1e-10, equal_nan=True):
        stress_bad += 1
dp = DEMProcessor(elev=E, dX=30.0, dY=30.0, fill_flats=False, drain_pits_path=False, drain_pits=PITS)
dp.calc_twi()
ok = stress_bad == 0
if stress_bad:
    print("rank", g.rank, "STRESS: %d of %d repeated passes differ" % (stress_bad, REPS - 1), flush=True)
for name, f, exact in (("mag", T.F_MAG, True), ("direction", T.F_DIR, True), ("flats", T.F_FLATS, True),
                       ("edge_todo", T.F_EDGE_TODO, True), ("edge_done", T.F_EDGE_DONE, True), ("uca", T.F_UCA, False)):
    a = eng.tile.download(f)[spec.lo:spec.hi]
    b = np.asarray(getattr(dp, name))[spec.r0:spec.r1]
    if exact:
        good = np.array_equal(a.astype(b.dtype), b, equal_nan=True) if a.dtype.kind == "f" else np.array_equal(a.astype(bool), b)
    else:
        good = np.allclose(a, b, rtol=1e-9, equal_nan=True)
    ok = ok and good
    if not good:
        print("rank", g.rank, "MISMATCH", name, flush=True)
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
if g.rank == 0:
    print("dist_check", "OK" if flag.item() == 0 else "FAILED", "world", g.world, "rows", R, "cols", C, "drain_pits", PITS, "passes", REPS, st, flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 0 else 1)
