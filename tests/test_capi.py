"""The shipped CUDA libraries load and export every symbol include/cpdp.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

import lfsd_b200  # noqa: F401
from lfsd_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "cpdp.h")).read()
    return sorted(set(re.findall(r"\b(cpdp_[a-z_0-9]+)\s*\(", txt)))


@pytest.mark.parametrize("model,dims", [("pendulum", (2, 1, 3, 0)), ("robotarm", (4, 2, 5, 0)),
                                        ("rocket", (13, 3, 12, 0)), ("quadrotor", (13, 4, 7, 3)),
                                        ("cartpole", (4, 1, 5, 0)), ("pendulum_tw2", (2, 1, 4, 0)),
                                        ("robotarm_wd", (4, 2, 5, 0)), ("quadrotor_cost1", (13, 4, 5, 3)),
                                        ("quadrotor_cost2", (13, 4, 11, 3)), ("rocket_cost1", (13, 3, 6, 0))])
def test_library_exports(model, dims):
    so = os.path.join(_capi.LIB_DIR, "libcpdp_%s.so" % model)
    if not os.path.exists(so):
        pytest.skip("library not built yet (run __graft_entry__.build())")
    lib = _capi.CpdpLib(so)
    raw = ctypes.CDLL(so)
    syms = _declared_symbols()
    assert "cpdp_solve" in syms and "cpdp_aux" in syms and "cpdp_reduce" in syms
    assert "cpdp_pack_rows" in syms and "cpdp_reduce_rows" in syms and "cpdp_optim_step" in syms
    for s in syms:
        assert hasattr(raw, s), "missing export %s" % s
    assert (lib.n, lib.m, lib.r, lib.q) == dims
    assert lib.workspace_bytes(8, 10, 4) > 0
    assert lib.workspace_bytes(0, 10, 4) == 0
    # argument errors are reported as negative codes before anything touches the device
    assert raw.cpdp_reduce(None, None, 0, None, None, None) == -1


def test_sass_is_sm100a_fp64():
    """The quadrotor library holds sm_100 SASS with DFMA instructions, bulk-copy (TMA) + mbarrier instructions and no tensor-core
    (HMMA/UTC*MMA) code."""
    import shutil
    import subprocess
    so = os.path.join(_capi.LIB_DIR, "libcpdp_quadrotor.so")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(so) or not os.path.exists(cuobjdump):
        pytest.skip("library or cuobjdump missing")
    out = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True).stdout
    assert "sm_100" in out
    assert "DFMA" in out
    assert "HMMA" not in out and "UTCHMMA" not in out
    # the forward sweep's node-row ring is fed by the TMA engine: bulk copy global -> shared, completion on an mbarrier
    assert "UBLKCP" in out and "SYNCS.PHASECHK" in out and "SYNCS.ARRIVE.TRANS64" in out
