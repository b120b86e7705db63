#!/usr/bin/env python3
"""Developer tool: backward-sweep time per problem against the number of problems resident per SM (the dynamic shared memory of
k_riccati_bdf is padded to limit residency; the batch is one full wave at each setting).  Separates per-warp latency from the
slow-down warps inflict on each other (instruction cache, shared pipes)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import lfsd_b200  # noqa: F401
    from lfsd_b200 import standard, synthetic, _capi, codegen
    oc = standard.quadrotor_oc(n_grid=50)
    text, info = codegen.generate_model_header(oc.lib_name, oc.state, oc.control, oc.auxvar, oc.dyn, oc.path_cost, oc.final_cost, oc.pvar)
    so = os.path.join(_capi.LIB_DIR, "libcpdp_quadrotor_occ.so")
    if "--build-only" in sys.argv or not os.path.exists(so):
        _capi.EXTRA_NVCC_FLAGS = ["-DCPDP_BDF_OCCUPANCY_KNOB"]
        so = _capi.build_model_library("quadrotor_occ", text, force=True)
        _capi.EXTRA_NVCC_FLAGS = []
    if "--build-only" in sys.argv:
        return
    import torch
    smem_total = 227 * 1024
    base = 28 * 1024
    for occ in (1, 2, 3, 4, 6, 8):
        pad = 0 if occ == 8 else max(0, smem_total // occ - base - 2048)
        os.environ["CPDP_BDF_SMEM_PAD"] = str(pad)
        B = 148 * occ
        qb = synthetic.quad_batch(B)
        ref = standard.quadrotor_oc(n_grid=50)
        ref.build(name=ref.lib_name)
        sol = ref.cocSolverBatch(qb["x0"], 1.0, qb["theta"], pdata=qb["goal"])
        oc2 = standard.quadrotor_oc(n_grid=50)
        oc2._lib = _capi.CpdpLib(so)
        oc2.aux_mode = oc2.MODE_BDF
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            aux = oc2.auxSysSolverBatch(sol, qb["taus"], qb["wp"], qb["sel"], phases=1)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(json.dumps({"resident_per_sm": occ, "batch": B, "smem_pad": pad, "ms": [round(t, 2) for t in ts],
                          "failed": int((aux["aux_status"] != 0).sum().item())}), flush=True)


if __name__ == "__main__":
    main()
