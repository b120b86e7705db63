"""Driver of tests/test_emu_tsan.py: one small problem through every kernel of the thread-emulated library built with
-fsanitize=thread (run with libtsan preloaded).  TEST INFRASTRUCTURE."""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.emu.support import emu_oc
name = sys.argv[1]
oc = emu_oc(name, tsan=True)
if name == "pendulum":
    th = np.array([1.0, 0.5, 1.5]); x0 = np.zeros((1, 2)); taus = np.array([0.4]); wp = np.array([[[1.0]]]); sel = [0]
    sol = oc.cocSolverBatch(x0, 1.0, th)
else:
    from lfsd_b200 import synthetic
    qb = synthetic.quad_batch(1); oc.setIntegrator(n_grid=4)
    sol = oc.cocSolverBatch(qb["x0"], 1.0, qb["theta"], pdata=qb["goal"]); taus, wp, sel = qb["taus"], qb["wp"], qb["sel"]
for mode in (oc.MODE_BDF, oc.MODE_RK45):
    oc.aux_mode = mode
    aux = oc.auxSysSolverBatch(sol, taus, wp, sel)
    print(name, "mode", mode, "status", aux["aux_status"], "counters", aux["counters"][0][:6])
