#!/usr/bin/env python3
"""Developer tool: throughput of the secondary BASELINE configs (robot arm B=256, rocket B=1024) through the public API,
as-shipped integrators, CUDA events.  Prints one JSON line per config."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lfsd_b200  # noqa
from lfsd_b200 import standard, synthetic


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


oc = standard.STANDARD["robotarm"](n_grid=30); oc.build(name=oc.lib_name); oc.aux_mode = oc.MODE_BDF
ab = synthetic.robotarm_batch(256)
ms = timed(lambda: oc.gradIterBatch(ab["x0"], 1.0, ab["theta"], ab["taus"], ab["wp"], ab["sel"]))
print(json.dumps({"config": "robotarm_random B=256 n_grid 30 (BDF/RK45 as shipped)", "ms_per_iter": ms, "ocp_grad_iters_per_s": 256 / ms * 1e3}))
oc = standard.STANDARD["rocket"](n_grid=15); oc.build(name=oc.lib_name); oc.aux_mode = oc.MODE_BDF
rb = synthetic.rocket_batch(1024)
demo = oc.cocSolverBatch(rb["x0"], 3.0, rb["theta_true"])
Xd = demo["X"].cpu().numpy()
wp = Xd[:, rb["tau_idx"]][:, :, rb["sel"]]
taus = np.asarray(demo["time_grid"])[rb["tau_idx"]]
ms = timed(lambda: oc.gradIterBatch(rb["x0"], 3.0, rb["theta0"], taus, wp, rb["sel"]))
print(json.dumps({"config": "rocket_groundtruth B=1024 n_grid 15 T=3 (BDF/RK45 as shipped)", "ms_per_iter": ms, "ocp_grad_iters_per_s": 1024 / ms * 1e3}))
