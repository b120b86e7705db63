"""Error and edge behaviour of the C ABI and of the host class, exercised on the thread-emulated kernel library (the same
cpdp_api.inl and kernels as the shipped CUDA build).  The reference reports API misuse with assertions
(/root/reference/CPDP/CPDP.py:93-97) and ignores IPOPT's status (:183); here misuse returns a negative code /
raises, and numerical outcomes are per-problem statuses."""
import ctypes

import numpy as np
import pytest

from tests.emu.support import emu_oc
from lfsd_b200 import _capi


@pytest.fixture(scope="module")
def pend():
    return emu_oc("pendulum")


def _solve_args(lib, B=1, N=10, S=4, ws_bytes=None, theta_stride=0, x0=None, theta=None):
    n, m, r = lib.n, lib.m, lib.r
    nbytes = lib.workspace_bytes(B, N, S)
    ws = np.zeros(nbytes, dtype=np.uint8)
    x0 = np.zeros((B, n)) if x0 is None else x0
    th = np.array([1.0, 0.5, 1.5]) if theta is None else theta
    out = dict(X=np.zeros((B, N + 1, n)), U=np.zeros((B, N + 1, m)), Lam=np.zeros((B, N + 1, n)),
               status=np.zeros(B, dtype=np.int32), iters=np.zeros(B, dtype=np.int32))
    keep = (ws, x0, th, out)
    args = [ws.ctypes.data, nbytes if ws_bytes is None else ws_bytes, B, N, S, 1.0, x0.ctypes.data, th.ctypes.data, theta_stride, 0,
            1e-10, 200, 0, out["X"].ctypes.data, out["U"].ctypes.data, out["Lam"].ctypes.data, out["status"].ctypes.data,
            out["iters"].ctypes.data, 0, 0, 0]
    return args, keep


def test_argument_error_codes(pend):
    lib = pend.build()
    L = lib.L
    args, keep = _solve_args(lib)
    assert L.cpdp_solve(*args) == 0 and int(keep[3]["status"][0]) == 1
    bad = list(args); bad[0] = 0                                   # no workspace
    assert L.cpdp_solve(*bad) == -1
    bad = list(args); bad[2] = 0                                   # B = 0 (empty batch)
    assert L.cpdp_solve(*bad) == -1
    bad = list(args); bad[8] = 2                                   # theta_stride neither 0 nor r
    assert L.cpdp_solve(*bad) == -2
    bad = list(args); bad[1] = 16                                  # workspace too small
    assert L.cpdp_solve(*bad) == -3
    assert lib.workspace_bytes(0, 10, 4) == 0 and lib.workspace_bytes(4, 0, 4) == 0
    with pytest.raises(_capi.CpdpError, match="code -3"):
        lib.check(-3, "cpdp_solve")
    import os
    so = os.path.join(_capi.LIB_DIR, "libcpdp_pendulum.so")        # the shipped CUDA library names its codes
    if os.path.exists(so):
        shipped = _capi.CpdpLib(so)
        assert b"workspace" in shipped.L.cpdp_error_string(-3) and b"phases" in shipped.L.cpdp_error_string(-9)
    with pytest.raises(_capi.CpdpError, match="not found"):
        _capi.CpdpLib("/nonexistent/libcpdp_nothing.so")           # no CPU fallback: a missing library is an error


def test_aux_argument_errors_and_modes(pend):
    sol = pend.cocSolverBatch(np.zeros((1, 2)), 1.0, np.array([1.0, 0.5, 1.5]))
    with pytest.raises(_capi.CpdpError, match="code -7"):
        pend.auxSysSolverBatch(sol, np.array([0.5]), np.array([[[1.0]]]), [0], mode=5)        # unknown integrator
    with pytest.raises(_capi.CpdpError, match="code -6"):
        pend.auxSysSolverBatch(sol, np.array([0.5]), np.array([[[1.0]]]), [7])                # observed index out of range
    with pytest.raises(_capi.CpdpError, match="code -9"):
        pend.auxSysSolverBatch(sol, np.array([0.5]), np.array([[[1.0]]]), [0], phases=4)
    # no waypoints at all is legal: sensitivities only, loss 0
    aux = pend.auxSysSolverBatch(sol)
    assert int(aux["aux_status"][0]) == 0 and float(aux["loss"][0]) == 0.0 and np.all(aux["dtheta"] == 0.0)


def test_per_problem_status_reporting(pend):
    """One batch holding a normal problem, one cut off by max_iter and one with a non-finite initial state: each gets its
    own status; the auxiliary sweep skips the non-solution (aux_status 3) and zeroes its loss row."""
    pend.max_iter = 3
    try:
        x0 = np.array([[0.0, 0.0], [0.0, 0.0], [np.nan, 0.0]])
        th = np.array([[1.0, 0.5, 1.5], [2.0, 1.0, 1.0], [1.0, 0.5, 1.5]])
        sol = pend.cocSolverBatch(x0, 1.0, th)
        st = [int(v) for v in sol["status"]]
        assert st[1] == 2 and int(sol["iters"][1]) == 3               # ST_MAXITER (needs 7 iterations)
        assert st[2] == 4                                              # ST_NUMERIC
        assert st[0] in (1, 2)
        aux = pend.auxSysSolverBatch(sol, np.array([0.5]), np.full((3, 1, 1), 1.0), [0], mode=pend.MODE_RK45)
        assert int(aux["aux_status"][2]) == 3 and float(aux["loss"][2]) == 0.0 and np.all(aux["dtheta"][2] == 0.0)
        assert int(aux["aux_status"][1]) == 0 and np.isfinite(aux["dtheta"][1]).all()   # max-iter exits are integrated as they are
        red = pend.reduceBatch(aux["loss"], aux["dtheta"])
        assert np.isfinite(red).all()
    finally:
        pend.max_iter = 200


def test_ragged_waypoint_times(pend):
    """Per-problem waypoint times ([B, W]) and shared ones ([W]) give the same rows when they coincide; a waypoint exactly
    at t = 0 or t = T is legal (interp1d's closed interval)."""
    th = np.array([1.0, 0.5, 1.5])
    sol = pend.cocSolverBatch(np.zeros((2, 2)), 1.0, th)
    wp = np.array([[[0.2], [1.5]], [[0.2], [1.5]]])
    pend.aux_mode = pend.MODE_RK45
    a1 = pend.auxSysSolverBatch(sol, np.array([0.0, 1.0]), wp, [0])
    a2 = pend.auxSysSolverBatch(sol, np.array([[0.0, 1.0], [0.0, 1.0]]), wp, [0])
    assert np.array_equal(a1["dtheta"], a2["dtheta"]) and np.array_equal(a1["loss"], a2["loss"])
    assert np.array_equal(a1["dtheta"][0], a1["dtheta"][1])


def test_failures_travel_with_the_sum(pend):
    """gradIterBatch reports the number of failed OCPs next to the reduced row (device-side, cpdp_pack_rows +
    cpdp_reduce_rows); a fixed number of Newton rounds that is too small leaves problems `running` and must not pass as a
    valid (all-zero) gradient; cpdp_grad_fn raises on it."""
    from lfsd_b200.optim import cpdp_grad_fn
    th = np.array([1.0, 0.5, 1.5])
    x0 = np.zeros((3, 2))
    wp = np.full((3, 1, 1), 1.0)
    pend.aux_mode = pend.MODE_RK45
    red, sol, aux = pend.gradIterBatch(x0, 1.0, th, np.array([0.5]), wp, [0])
    assert float(aux["n_failed"][0]) == 0.0 and red.shape == (4,)
    full = pend.reduceRows(pend.packRows(sol, aux))
    assert np.array_equal(full[:4], red) and np.array_equal(red, pend.reduceBatch(aux["loss"], aux["dtheta"]))
    red2, sol2, aux2 = pend.gradIterBatch(x0, 1.0, th, np.array([0.5]), wp, [0], rounds=2)     # needs 6-7 rounds
    assert [int(v) for v in sol2["status"]] == [0, 0, 0]
    assert float(aux2["n_failed"][0]) == 3.0 and np.all(red2 == 0.0)
    # an out-of-range waypoint time is an error status, not an extrapolation (scipy's interp1d raises in the reference)
    aux3 = pend.auxSysSolverBatch(sol, np.array([[0.5], [1.2], [-0.1]]), wp, [0])
    assert [int(v) for v in aux3["aux_status"]] == [0, 5, 5]
    assert float(aux3["loss"][1]) == 0.0 and np.all(aux3["dtheta"][1:] == 0.0)
    assert float(pend.reduceRows(pend.packRows(sol, aux3))[4]) == 2.0
    fn = cpdp_grad_fn(pend, x0, 1.0, np.array([1.5]), wp, [0], mode=pend.MODE_RK45)
    with pytest.raises(FloatingPointError, match="failed for 3 OCP"):
        fn(th)
    # max_iter beyond the filter capacity is an argument error, not a silent clamp
    args, keep = _solve_args(pend.build())
    bad = list(args); bad[11] = 256
    assert pend.build().L.cpdp_solve(*bad) == -10


def test_reference_shaped_aux_rejects_what_it_cannot_integrate(pend):
    th = np.array([1.0, 0.5, 1.5])
    tg, opt_sol = pend.cocSolver([0.0, 0.0], 1, th, interplation_level=2)
    with pytest.raises(NotImplementedError, match="linear"):
        pend.auxSysSolver(tg, opt_sol, th)
    tg, opt_sol = pend.cocSolver([0.0, 0.0], 1, th)
    tg2 = tg.copy(); tg2[3] += 0.01
    with pytest.raises(NotImplementedError, match="uniform"):
        pend.auxSysSolver(tg2, opt_sol, th)
