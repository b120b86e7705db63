#!/usr/bin/env python3
"""Developer tool: times the forward solve (cocSolverBatch) of the quadrotor benchmark batch for a given library name
(variants built with different CPDP_EXTRA_NVCC_FLAGS under different names)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--name", default="quadrotor")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--build-only", action="store_true")
ap.add_argument("--no-build", action="store_true", help="load lib/libcpdp_<name>.so as it is (A/B runs against a library built from another revision)")
a = ap.parse_args()
import lfsd_b200  # noqa
from lfsd_b200 import standard, synthetic
oc = standard.quadrotor_oc(n_grid=50)
if a.no_build:
    from lfsd_b200 import _capi
    oc._lib = _capi.CpdpLib(os.path.join(_capi.LIB_DIR, "libcpdp_%s.so" % a.name))
else:
    oc.build(name=a.name)
if a.build_only:
    sys.exit(0)
import torch
qb = synthetic.quad_batch(a.batch)
ts = []
for _ in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sol = oc.cocSolverBatch(qb["x0"], 1.0, qb["theta"], pdata=qb["goal"])
    e1.record(); torch.cuda.synchronize()
    ts.append(round(e0.elapsed_time(e1), 2))
print(json.dumps({"name": a.name, "ms": ts, "conv": int((sol["status"] == 1).sum().item())}))
