"""In-memory emulation of the reference's cross-tile UCA edge resolution over row blocks
(process_manager.calc_uca :94-197 and calc_uca_ec :224-284 without zarr), used to exercise
``calc_uca(uca_init=..., edge_init_data=...)`` with realistic inputs.  ``make(elev, **kw)`` is
any DEMProcessor-compatible factory (reference, oracle or the CUDA drop-in)."""
import numpy as np


def tiled_rows(make, E, nblocks, overlap, kw, max_rounds=60, record=False):
    R, C = E.shape
    edges = np.linspace(0, R, nblocks + 1).astype(int)
    spans = [(max(0, edges[k] - overlap), min(R, edges[k + 1] + overlap)) for k in range(nblocks)]
    st = []
    for (t, b) in spans:
        dp = make(E[t:b], **kw)
        dp.calc_slopes_directions(); dp.find_flats(); dp.calc_uca()
        st.append(dict(t=t, b=b, uca0=np.array(dp.uca), edges=np.zeros_like(dp.uca), todo=np.array(dp.edge_todo),
                       done=np.array(dp.edge_done), dir=np.array(dp.direction), mag=np.array(dp.mag)))
    log, calls = [], []
    for _ in range(max_rounds):
        best, bm = None, 0
        for k, s in enumerate(st):          # process_manager.calc_uca_ec_metrics :199-221
            m = 0
            for side, nk in (("top", k - 1), ("bottom", k + 1)):
                if nk < 0 or nk >= nblocks:
                    continue
                n = st[nk]
                myrow = s["t"] if side == "top" else s["b"] - 1
                m += int((s["todo"][0 if side == "top" else -1] & n["done"][myrow - n["t"]]).sum())
            if m > bm:
                best, bm = k, m
        if best is None:
            break
        s = st[best]
        nr = s["b"] - s["t"]
        data = {"left": np.zeros(nr), "right": np.zeros(nr), "top": np.zeros(C), "bottom": np.zeros(C)}
        done = {k: np.zeros(v.shape, bool) for k, v in data.items()}
        todo = {"left": s["todo"][:, 0].copy(), "right": s["todo"][:, -1].copy(),
                "top": s["todo"][0].copy(), "bottom": s["todo"][-1].copy()}
        for side, nk in (("top", best - 1), ("bottom", best + 1)):
            if nk < 0 or nk >= nblocks:
                continue
            n = st[nk]
            myrow = s["t"] if side == "top" else s["b"] - 1
            data[side] = (n["uca0"] + n["edges"])[myrow - n["t"]].copy()
            done[side] = n["done"][myrow - n["t"]].copy()
            todo[side] = todo[side] & ~n["todo"][myrow - n["t"]]      # process_manager.py:274
        dp = make(E[s["t"]:s["b"]], direction=s["dir"].copy(), mag=s["mag"].copy(), **kw)
        dp.find_flats()
        uca_init = s["uca0"] + s["edges"]
        dp.calc_uca(uca_init=uca_init, edge_init_data=[data, done, todo])
        if record:
            c = dict(block=np.array([s["t"], s["b"]]), direction=s["dir"], mag=s["mag"], uca_init=uca_init,
                     out_uca=np.array(dp.uca), out_todo=np.array(dp.edge_todo), out_done=np.array(dp.edge_done))
            for nm, d in (("data", data), ("done", done), ("todo", todo)):
                for side in ("left", "right", "top", "bottom"):
                    c["%s_%s" % (nm, side)] = np.array(d[side])
            calls.append(c)
        s["edges"] = np.array(dp.uca) - s["uca0"]; s["todo"] = np.array(dp.edge_todo); s["done"] = np.array(dp.edge_done)
        log.append((best, bm))
    return st, log, calls


def stitch(st, R, C):
    """Each physical row from the block where it is farthest from a block edge."""
    out = np.full((R, C), np.nan)
    depth = np.full(R, -1)
    for s in st:
        u = s["uca0"] + s["edges"]
        for i in range(s["t"], s["b"]):
            d = min(i - s["t"], s["b"] - 1 - i)
            if d > depth[i]:
                depth[i] = d; out[i] = u[i - s["t"]]
    return out
