"""Stage timings on the GPU box (device-resident), with sweep statistics.
    python scripts/gpu_perf.py [raw|cond] sizes..."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pydem_b200 import synth, tile as T

args = sys.argv[1:]
kind = "raw"
if args and args[0] in ("raw", "cond"):
    kind = args.pop(0)
sizes = [int(a) for a in args] or [4096]
stream = torch.cuda.current_stream().cuda_stream
for n in sizes:
    t0 = time.time()
    E = synth.fractal_dem(n, 0) if kind == "raw" else synth.conditioned_fractal_dem(n, 0)
    tg = time.time() - t0
    for drain_pits in (0, 1):
        dt = T.DeviceTile(n, n, stream=stream)
        dt.set_spacing(30.0, 30.0)
        dt.upload(T.F_ELEV, E)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for rep in range(3):
            torch.cuda.synchronize()
            ev[0].record(); dt.slopes_directions(); ev[1].record()
            st = dt.uca(drain_pits=drain_pits); ev[2].record()
            dt.twi(); ev[3].record()
            torch.cuda.synchronize()
            res = dict(kind=kind, n=n, drain_pits=drain_pits, gen_s=round(tg, 2), ms_slopes=ev[0].elapsed_time(ev[1]), ms_uca=ev[1].elapsed_time(ev[2]),
                       ms_twi=ev[2].elapsed_time(ev[3]), **{k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
            res["Mcells_s_slopes"] = round(n * n / res["ms_slopes"] / 1e3, 1)
            res["Mcells_s_uca"] = round(n * n / res["ms_uca"] / 1e3, 1)
            print(json.dumps(res), flush=True)
        dt.close()
