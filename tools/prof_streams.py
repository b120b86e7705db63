#!/usr/bin/env python3
"""Developer tool: does splitting the batch over two CUDA streams (fixed Newton rounds, no host sync) hide kernel tails?
Times one gradient iteration of the 4096-OCP quadrotor batch: (a) adaptive rounds, (b) fixed rounds, (c) fixed rounds,
`chunks` chunks on separate streams with separate workspaces."""
import argparse, copy, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--rounds", type=int, default=8)
ap.add_argument("--chunks", type=int, default=2)
a = ap.parse_args()
import torch
import lfsd_b200  # noqa
from lfsd_b200 import standard, synthetic

B = a.batch
qb = synthetic.quad_batch(B)
dev = torch.device("cuda", 0)
res = {k: torch.as_tensor(np.asarray(qb[k], dtype=float)).to(dev) for k in ("x0", "goal", "wp")}
ocs = []
for c in range(a.chunks):
    oc = standard.quadrotor_oc(n_grid=50); oc.build(name=oc.lib_name); oc.aux_mode = oc.MODE_BDF
    ocs.append(oc)
oc = ocs[0]
streams = [torch.cuda.Stream() for _ in range(a.chunks)]


def single(rounds):
    red, sol, aux = oc.gradIterBatch(res["x0"], 1.0, qb["theta"], qb["taus"], res["wp"], qb["sel"], pdata=res["goal"], rounds=rounds)
    return red


def chunked(rounds):
    main = torch.cuda.current_stream()
    rows = []
    per = B // a.chunks
    for c, (o, s) in enumerate(zip(ocs, streams)):
        s.wait_stream(main)
        with torch.cuda.stream(s):
            lo, hi = c * per, (c + 1) * per
            sol = o.cocSolverBatch(res["x0"][lo:hi], 1.0, qb["theta"], pdata=res["goal"][lo:hi], rounds=rounds)
            aux = o.auxSysSolverBatch(sol, qb["taus"], res["wp"][lo:hi], qb["sel"])
            rows.append((aux["loss"], aux["dtheta"]))
    for s in streams:
        main.wait_stream(s)
    loss = torch.cat([r[0] for r in rows]); dth = torch.cat([r[1] for r in rows])
    return oc.reduceBatch(loss, dth)


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


t_a, r_a = timed(lambda: single(0))
t_b, r_b = timed(lambda: single(a.rounds))
t_c, r_c = timed(lambda: chunked(a.rounds))
print(json.dumps({"adaptive_ms": t_a, "fixed_rounds_ms": t_b, "chunked_ms": t_c, "chunks": a.chunks, "rounds": a.rounds,
                  "same_result_fixed": bool(torch.equal(r_a, r_b)), "rel_diff_chunked": float((r_c - r_a).abs().max() / r_a.abs().max())}))
