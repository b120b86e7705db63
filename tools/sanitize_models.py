#!/usr/bin/env python3
"""Developer tool: one small gradient iteration (BDF + RK45 modes) of each standard model, meant to be run under
compute-sanitizer (memcheck / racecheck) on a GPU box."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lfsd_b200  # noqa
from lfsd_b200 import standard, synthetic

cases = {
    "pendulum": dict(n_grid=4, x0=np.zeros((2, 2)), T=1.0, th=np.array([1.0, 0.5, 1.5]), taus=np.array([0.4]), wp=np.full((2, 1, 1), 1.0), sel=[0]),
    "cartpole": dict(n_grid=4, x0=np.tile([0.0, 0.3, 0.0, 0.0], (2, 1)), T=1.0, th=np.array([1.5, 0.5, 1.0, 0.2, 0.3]), taus=np.array([0.5]), wp=np.full((2, 1, 2), 0.5), sel=[0, 1]),
}
ab = synthetic.robotarm_batch(2)
cases["robotarm"] = dict(n_grid=6, x0=ab["x0"], T=1.0, th=ab["theta"], taus=ab["taus"], wp=ab["wp"], sel=ab["sel"])
rb = synthetic.rocket_batch(2)
cases["rocket"] = dict(n_grid=4, x0=rb["x0"], T=3.0, th=rb["theta0"], taus=np.array([0.75, 2.0]), wp=np.zeros((2, 2, 7)), sel=rb["sel"])
qb = synthetic.quad_batch(2)      # (its node rows are a multiple of 16 bytes: the forward sweep's bulk-copy ring is active)
cases["quadrotor"] = dict(n_grid=4, x0=qb["x0"], T=1.0, th=qb["theta"], taus=qb["taus"], wp=qb["wp"], sel=qb["sel"], pdata=qb["goal"])
only = sys.argv[1:]
for name, c in cases.items():
    if only and name not in only:
        continue
    oc = standard.STANDARD[name](n_grid=c["n_grid"])
    oc.build(name=oc.lib_name)
    for mode in (oc.MODE_BDF, oc.MODE_RK45):
        red, sol, aux = oc.gradIterBatch(c["x0"], c["T"], c["th"], c["taus"], c["wp"], c["sel"], mode=mode, pdata=c.get("pdata"))
        torch.cuda.synchronize()
        print(name, "mode", mode, "status", sol["status"].tolist(), "aux", aux["aux_status"].tolist(), "counters", aux["counters"][0].tolist())
