#!/usr/bin/env python3
"""Profiling driver (developer tool, not part of the product path): runs the quadrotor benchmark batch through the
forward solve once, then the backward Riccati sweep `--reps` times, printing CUDA-event times and the distribution of
the per-OCP work counters.  Meant to be run under ncu with `-k regex:k_riccati` for the detailed capture."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=592)
    ap.add_argument("--n-grid", type=int, default=50)
    ap.add_argument("--mode", default="bdf")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--phases", type=int, default=1)
    ap.add_argument("--name", default=None, help="library name (variants built under other names)")
    ap.add_argument("--build-only", action="store_true")
    a = ap.parse_args()
    import lfsd_b200  # noqa: F401
    from lfsd_b200 import standard, synthetic
    oc = standard.quadrotor_oc(n_grid=a.n_grid)
    oc.build(name=a.name or oc.lib_name)
    if a.build_only:
        return
    oc.aux_mode = oc.MODE_BDF if a.mode == "bdf" else oc.MODE_RK45
    import torch
    qb = synthetic.quad_batch(a.batch)
    sol = oc.cocSolverBatch(qb["x0"], 1.0, qb["theta"], pdata=qb["goal"])
    torch.cuda.synchronize()
    times = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        aux = oc.auxSysSolverBatch(sol, qb["taus"], qb["wp"], qb["sel"], phases=a.phases)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    cnt = aux["counters"].cpu().numpy()
    names = ["back_rhs", "back_steps", "fwd_rhs", "fwd_steps", "back_lu", "back_jac"]
    out = {"batch": a.batch, "mode": a.mode, "ms": times}
    for i, nm in enumerate(names):
        c = cnt[:, i]
        out[nm] = dict(min=int(c.min()), p50=float(np.median(c)), mean=float(c.mean()), p99=float(np.percentile(c, 99)), max=int(c.max()))
    out["aux_failed"] = int((aux["aux_status"] != 0).sum().item())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
