#!/usr/bin/env python3
"""Developer tool: replays the reference's stored quadrotor learning run (100 Nesterov iterations) through the CUDA path
and prints the deviation of parameter_trace / loss_trace per iteration."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lfsd_b200  # noqa
from lfsd_b200 import standard
from lfsd_b200.optim import Learner, cpdp_grad_fn
g = np.load(os.path.join(ROOT, "tests", "golden", "quad_run.npz"))
oc = standard.quadrotor_oc(n_grid=25); oc.build(name=oc.lib_name)
oc.aux_mode = oc.MODE_BDF if (len(sys.argv) < 2 or sys.argv[1] == "bdf") else oc.MODE_RK45
P, lr, mu = g["parameter_trace"], float(g["learning_rate"]), float(g["mu"])
fn = cpdp_grad_fn(oc, g["ini_state"].reshape(1, 13), 1.0, g["time_grid"], g["waypoints"].reshape(1, -1, 3), [0, 1, 2], pdata=g["goal_position"].reshape(1, 3))
L = Learner(fn, 7)
L.load_optimization_function({"learning_rate": lr, "iter_num": 100, "method": "Nesterov", "mu": mu, "true_loss_print_flag": False})
L.run(P[0])
got = np.array(L.parameter_trace)
rel = np.linalg.norm(got - P[:len(got)], axis=1) / np.linalg.norm(P[:len(got)], axis=1)
lrel = np.abs(np.array(L.loss_trace) - g["loss_trace"][:len(L.loss_trace)]) / g["loss_trace"][:len(L.loss_trace)]
print(json.dumps({"iterations": len(L.loss_trace), "theta_rel_err_at": {str(j): float(rel[j]) for j in (1, 5, 10, 20, 50, 75, len(got) - 1) if j < len(got)},
                  "theta_rel_err_max": float(rel.max()), "loss_rel_err_max": float(lrel.max()), "final_loss": L.loss_trace[-1], "stored_final_loss": float(g["loss_trace"][-1])}))
