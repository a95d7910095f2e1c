"""CPU suite for the N>1 path: two real processes over gloo.

* the neighbour-exchange / all-reduce plumbing of pydem_b200.sharded.DistGroup,
* the boundary-row protocol itself (a cell pulls its donors' final values; "not done" travels as
  NaN in the boundary rows that neighbours exchange after every local sweep; stop when no rank
  completed a cell whose receiver lives across a boundary): simulated in NumPy on the oracle's
  drainage graph, it must reproduce the single-tile oracle UCA."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from pydem_b200 import sharded, synth
        from oracle.oracle import OracleDEMProcessor
        g = sharded.DistGroup()
        # --- plumbing
        C = 17
        b = dict(send_up=torch.full((C,), 10.0 * rank + 1) if rank > 0 else None,
                 recv_up=torch.zeros(C) if rank > 0 else None,
                 send_down=torch.full((C,), 10.0 * rank + 2) if rank < world - 1 else None,
                 recv_down=torch.zeros(C) if rank < world - 1 else None)
        g.exchange([b])
        if rank > 0:
            assert float(b["recv_up"][0]) == 10.0 * (rank - 1) + 2
        if rank < world - 1:
            assert float(b["recv_down"][0]) == 10.0 * (rank + 1) + 1
        assert g.allreduce_sum([torch.tensor([rank + 1])])[0] == world * (world + 1) // 2
        # --- boundary-row protocol on the oracle's graph
        E = synth.fractal_dem(0, 5, shape=(60, 40)) * 0.05 + synth.cone_dem(60)[:, :40] * 200 + 1
        R, Cc = E.shape
        dp = OracleDEMProcessor(E, fill_flats=False, drain_pits_path=False, drain_pits=False)
        dp.calc_slopes_directions(); ref = dp.calc_uca().copy()
        dp2 = OracleDEMProcessor(E, fill_flats=False, drain_pits_path=False, drain_pits=False)
        dp2.calc_slopes_directions(); gr, sec = dp2._graph()
        cptr, cidx, cdat, rptr, ridx = gr.export()
        wgt = {}
        for i in range(R * Cc):
            for e in range(cptr[i], cptr[i + 1]):
                wgt[(i, int(cidx[e]))] = cdat[e]
        r0, r1 = sharded.row_blocks(R, world)[rank]
        own = np.zeros(R * Cc, bool); own[r0 * Cc:r1 * Cc] = True
        area = np.full(R * Cc, np.nan)          # NaN = not done (the device uses a signalling-NaN pattern)
        todo = set(int(i) for i in np.nonzero(own)[0])
        rounds = 0
        while True:
            sent = 0
            progress = True
            while progress:                    # local sweep until it rests
                progress = False
                for r in sorted(todo):
                    srcs = [int(x) for x in ridx[rptr[r]:rptr[r + 1]]]
                    if all(not np.isnan(area[x]) for x in srcs):
                        area[r] = 1.0 + sum(area[x] * wgt[(x, r)] for x in srcs)
                        todo.discard(r); progress = True
                        sent += sum(1 for e in range(cptr[r], cptr[r + 1]) if not own[cidx[e]])
            rounds += 1
            if g.allreduce_sum([torch.tensor([sent])])[0] == 0:
                break
            A2 = area.reshape(R, Cc)
            bufs = dict(send_up=torch.from_numpy(A2[r0].copy()) if rank > 0 else None,
                        recv_up=torch.zeros(Cc, dtype=torch.float64) if rank > 0 else None,
                        send_down=torch.from_numpy(A2[r1 - 1].copy()) if rank < world - 1 else None,
                        recv_down=torch.zeros(Cc, dtype=torch.float64) if rank < world - 1 else None)
            g.exchange([bufs])
            if rank > 0:
                A2[r0 - 1] = bufs["recv_up"].numpy()
            if rank < world - 1:
                A2[r1] = bufs["recv_down"].numpy()
        indeg = np.zeros(R * Cc, np.int64)
        indeg[list(todo)] = 1
        mine = area.reshape(R, Cc)[r0:r1]
        refm = ref[r0:r1]
        ok = np.allclose(np.where(np.isnan(refm), 0, mine), np.nan_to_num(refm), rtol=1e-12) and (indeg[own] == 0).all()
        q.put((rank, bool(ok), rounds))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, False, traceback.format_exc()))


def test_row_blocks_and_spec():
    from pydem_b200 import sharded
    assert sharded.row_blocks(10, 3) == [(0, 3), (3, 6), (6, 10)]
    with pytest.raises(ValueError):
        sharded.row_blocks(3, 2)
    s = sharded.ShardSpec(100, 7, 1, 4)
    assert (s.r0, s.r1, s.row_off, s.Rl, s.lo, s.hi) == (25, 50, 24, 27, 1, 26)
    s = sharded.ShardSpec(100, 7, 0, 4)
    assert (s.row_off, s.Rl, s.lo, s.hi) == (0, 26, 0, 25)
    s = sharded.ShardSpec(100, 7, 3, 4)
    assert (s.row_off, s.Rl, s.lo, s.hi) == (74, 26, 1, 26)
    th = sharded.global_row_theta(np.array([1.0, 2.0, 3.0, 4.0]), np.ones(4))
    ref = np.arctan2(np.ones(4), np.array([1.0, 2.0, 3.0, 4.0]))
    np.testing.assert_array_equal(th, ref[[0, 0, 1, 2, 2]])


def test_local_group_exchange():
    from pydem_b200 import sharded
    g = sharded.LocalGroup(3)
    bufs = []
    for k in range(3):
        bufs.append(dict(send_up=np.full(4, k + 0.1) if k > 0 else None, recv_up=np.zeros(4) if k > 0 else None,
                         send_down=np.full(4, k + 0.2) if k < 2 else None, recv_down=np.zeros(4) if k < 2 else None))
    g.exchange(bufs)
    assert bufs[1]["recv_up"][0] == 0.2 and bufs[1]["recv_down"][0] == 2.1 and bufs[2]["recv_up"][0] == 1.2
    assert g.allreduce_sum([1, 2, 3]) == [6, 6, 6]


def test_value_noise_is_shard_consistent():
    from pydem_b200 import synth
    full = synth.value_noise_dem(0, 96, 64, seed=2)
    part = synth.value_noise_dem(37, 20, 64, seed=2)
    np.testing.assert_array_equal(full[37:57], part)
    assert full.min() >= 1.0 and np.isfinite(full).all()


@pytest.mark.parametrize("world", [2, 3])     # 3: a middle rank talks to two neighbours
def test_gloo_protocol(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, info in res:
        assert ok, "rank %d: %s" % (rank, info)
    assert all(r[2] >= 2 for r in res)


def test_value_noise_generators_agree():
    """The device-side generator of the config-4 DEM (torch, row chunks) and the NumPy one give identical bits."""
    import torch
    from pydem_b200 import synth
    a = synth.value_noise_dem(4000, 130, 517, seed=11)
    out = torch.empty((130, 517), dtype=torch.float64)
    synth.value_noise_dem_torch(out, 4000, seed=11, chunk=48)
    assert np.array_equal(a, out.numpy())
