"""world_size-2 gloo test of the N>1 path on CPU: contiguous shards, all-gather of the per-OCP [loss | dL/dtheta] rows,
fixed-tree reduction on every rank — bit-identical to the single-process result.  The per-OCP numbers
come from the host-emulated kernels (tests/emu), driven through the same C ABI and COCSys host class as on the GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _batch(B):
    rng = np.random.default_rng(11)
    theta = np.stack([rng.uniform(0.8, 2.0, B), rng.uniform(0.5, 1.5, B), rng.uniform(0.5, 1.5, B)], 1)
    taus = np.array([0.25, 0.55, 0.8])
    wp = rng.uniform(0.0, 2.0, size=(B, 3, 1))
    return np.zeros((B, 2)), theta, taus, wp


def _rows(oc, x0, theta, taus, wp):
    oc.aux_mode = oc.MODE_RK45
    sol = oc.cocSolverBatch(x0, 1.0, theta)
    aux = oc.auxSysSolverBatch(sol, taus, wp, [0])
    assert (np.asarray(sol["status"]) == 1).all() and (np.asarray(aux["aux_status"]) == 0).all()
    return np.concatenate([np.asarray(aux["loss"])[:, None], np.asarray(aux["dtheta"])], 1)


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.emu.support import emu_oc
    from lfsd_b200 import synthetic
    oc = emu_oc("pendulum")
    x0, theta, taus, wp = _batch(B)
    lo, hi = synthetic.shard_bounds(B, world, rank)
    mine = torch.from_numpy(_rows(oc, x0[lo:hi], theta[lo:hi], taus, wp[lo:hi]))
    gathered = torch.empty((B, mine.shape[1]), dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, mine)
    g = gathered.numpy()
    red = oc.reduceBatch(np.ascontiguousarray(g[:, 0]), np.ascontiguousarray(g[:, 1:]))
    q.put((rank, lo, hi, g.copy(), np.asarray(red).copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allgather_matches_single_process():
    B, world = 8, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    from tests.emu.support import emu_oc
    from lfsd_b200 import synthetic
    oc = emu_oc("pendulum")
    x0, theta, taus, wp = _batch(B)
    full = _rows(oc, x0, theta, taus, wp)
    red_full = np.asarray(oc.reduceBatch(np.ascontiguousarray(full[:, 0]), np.ascontiguousarray(full[:, 1:])))
    assert [(r[1], r[2]) for r in res] == [synthetic.shard_bounds(B, world, r) for r in range(world)] == [(0, 4), (4, 8)]
    for rank, lo, hi, g, red in res:
        assert np.array_equal(g, full)              # shard assignment and gathering change nothing, bit for bit
        assert np.array_equal(red, red_full)        # every rank holds the same reduced [loss | dL/dtheta]
