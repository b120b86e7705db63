import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import lfsd_b200
from lfsd_b200 import standard, synthetic
oc = standard.quadrotor_oc(n_grid=50)
oc.build(name=oc.lib_name)
oc.aux_mode = oc.MODE_BDF
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
qb = synthetic.quad_batch(B)
sol = oc.cocSolverBatch(qb["x0"], 1.0, qb["theta"], pdata=qb["goal"])
aux = oc.auxSysSolverBatch(sol, qb["taus"], qb["wp"], qb["sel"], phases=1)
torch.cuda.synchronize()
st = aux["aux_status"].cpu().numpy(); cnt = aux["counters"].cpu().numpy()
bad = np.flatnonzero(st != 0)
print("failing", bad.tolist())
print("status", st[bad].tolist())
print("counters", cnt[bad].tolist())
if len(bad) and os.environ.get("CPDP_EXTRA_NVCC_FLAGS"):
    Xa = aux["Xa"].cpu().numpy()
    np.save(os.path.join(ROOT, "gpurun_out", "dump.npy"), Xa[bad[0]].ravel()[:5 * 169 + 1])
    print("dumped", bad[0])
