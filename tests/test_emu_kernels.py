"""CPU tests of the CUDA kernel LOGIC: the kernels of csrc/*.cuh are compiled with g++ against the thread-per-CUDA-
thread emulation in tests/emu/ and driven through the same C ABI and the same COCSys host class as on the GPU.
(The GPU parity tests proper are in test_gpu_parity.py, marker `gpu`.)"""
import os

import numpy as np
import pytest

from tests.emu.support import emu_oc
from oracle import models
from oracle.cpdp_oracle import Oracle, TIGHT

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def pend():
    return emu_oc("pendulum")


@pytest.fixture(scope="module")
def quad():
    return emu_oc("quadrotor")


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def test_pendulum_solve_and_aux_vs_oracle(pend):
    orc = Oracle(models.pendulum(), n_grid=10)
    th = np.array([1.0, 0.5, 1.5])
    tg, opt_sol = pend.cocSolver([0.0, 0.0], 1, th)
    tgo, X, U, Lam, info = orc.solve([0.0, 0.0], 1.0, th, return_info=True)
    assert pend.last_status == 1 and pend.last_iters == info["iters"]
    assert np.array_equal(tg, tgo)
    nodes = opt_sol(tg)
    assert np.abs(nodes[:, :2] - X).max() < 1e-9 * max(1.0, np.abs(X).max())
    assert np.abs(nodes[:, 2:3] - U).max() < 1e-9 * max(1.0, np.abs(U).max())
    assert np.abs(nodes[:, 3:] - Lam).max() < 1e-8 * max(1.0, np.abs(Lam).max())
    # RK45 / RK45 at scipy defaults and tight: identical step sequences -> agreement to rounding
    for (rb, ab, back, fwd) in ((1e-3, 1e-6, {}, {}), (1e-10, 1e-12, TIGHT, TIGHT)):
        pend.aux_mode = pend.MODE_RK45
        pend.rtol_back, pend.atol_back, pend.rtol_fwd, pend.atol_fwd = rb, ab, rb, ab
        aux_sol = pend.auxSysSolver(tg, opt_sol, th)
        Xa, Ua, PW, cnt = orc.aux(tgo, X, U, Lam, th, back=back, fwd=fwd, return_counts=True)
        got = aux_sol(tg)
        assert pend.last_aux_status == 0
        assert pend.last_aux_counters[0] == cnt["back_rhs"] and pend.last_aux_counters[2] == cnt["fwd_rhs"]
        assert _rel(got[:, :6], Xa) < 1e-9 and _rel(got[:, 6:], Ua) < 1e-9


def test_pendulum_long_line_search(pend):
    """theta* = [2,1,1]: swing-up with a long merit plateau (J*=11.4644386014)."""
    orc = Oracle(models.pendulum(), n_grid=10)
    th = np.array([2.0, 1.0, 1.0])
    sol = pend.cocSolverBatch(np.zeros((1, 2)), 1.0, th)
    _, X, U, Lam, info = orc.solve([0.0, 0.0], 1.0, th, return_info=True)
    assert int(sol["status"][0]) == 1 and int(sol["iters"][0]) == info["iters"]
    assert abs(float(sol["cost"][0]) - 11.4644386014) < 1e-8
    assert np.abs(sol["X"][0] - X).max() < 1e-8


def test_quad_k2_stored_optimum(quad):
    """Emulated CUDA solve vs the reference's stored IPOPT optimum (KAT K2)."""
    g = np.load(os.path.join(HERE, "golden", "quad_run.npz"))
    quad.setIntegrator(n_grid=25)
    sol = quad.cocSolverBatch(g["ini_state"].reshape(1, 13), 1.0, g["parameter_trace"][-1],
                              pdata=g["goal_position"].reshape(1, 3))
    assert int(sol["status"][0]) == 1 and int(sol["iters"][0]) <= 7
    assert np.abs(sol["X"][0] - g["opt_state_traj"][::4]).max() < 1e-10
    assert np.abs(sol["U"][0] - g["opt_control_traj"][::4]).max() < 1e-10
    assert np.array_equal(sol["U"][0][-1], sol["U"][0][-2])          # CPDP.py:191


def test_quad_batch_grad_iter_rk45_vs_oracle(quad):
    """Two different OCPs in one batch (different start/goal/waypoints), shared theta, RK45 sweeps at defaults."""
    from lfsd_b200 import synthetic
    qb = synthetic.quad_batch(2)
    quad.setIntegrator(n_grid=10)
    quad.aux_mode = quad.MODE_RK45
    quad.rtol_back, quad.atol_back, quad.rtol_fwd, quad.atol_fwd = 1e-3, 1e-6, 1e-3, 1e-6
    red, sol, aux = quad.gradIterBatch(qb["x0"], 1.0, qb["theta"], qb["taus"], qb["wp"], qb["sel"], pdata=qb["goal"])
    orc = Oracle(models.quadrotor(), n_grid=10)
    tot = np.zeros(8)
    for b in range(2):
        orc.pd = qb["goal"][b]
        loss, dl, ex = orc.grad_iter(qb["x0"][b], 1.0, qb["theta"], qb["taus"], qb["wp"][b], back={}, fwd={})
        assert int(sol["iters"][b]) == ex["info"]["iters"]
        assert np.abs(sol["X"][b] - ex["X"]).max() < 1e-8
        assert abs(aux["loss"][b] - loss) < 1e-9 * loss
        assert _rel(aux["dtheta"][b], dl) < 1e-8
        tot += np.concatenate([[loss], dl])
    assert _rel(red, tot) < 1e-12
    assert np.array_equal(red, np.concatenate([[aux["loss"][0] + aux["loss"][1]], aux["dtheta"][0] + aux["dtheta"][1]]))


def test_reduce_tree_is_sharding_invariant(quad):
    """The canonical tree sum of B rows equals the tree sum of the tree sums of aligned power-of-two shards."""
    rng = np.random.default_rng(5)
    B = 64
    loss = rng.normal(size=B) * 10 ** rng.uniform(-3, 3, size=B)
    dth = rng.normal(size=(B, 7)) * 10 ** rng.uniform(-3, 3, size=(B, 1))
    full = quad.reduceBatch(loss, dth)
    for G in (2, 4, 8):
        parts = np.stack([quad.reduceBatch(loss[g * B // G:(g + 1) * B // G], dth[g * B // G:(g + 1) * B // G]) for g in range(G)])
        again = quad.reduceBatch(parts[:, 0].copy(), parts[:, 1:].copy())
        assert np.array_equal(again, full)
    # and it is a correct sum
    assert np.allclose(full, np.concatenate([[loss.sum()], dth.sum(0)]), rtol=1e-12)


def test_pendulum_bdf_asshipped_vs_oracle(pend):
    """Mode 1 (scipy-BDF control logic with the closed-form Jacobian) against the oracle calling scipy's own BDF
    exactly as the reference does (CPDP.py:333-336): same accepted steps, node tables equal to ~1e-9."""
    from scipy.integrate import BDF
    orc = Oracle(models.pendulum(), n_grid=10)
    th = np.array([1.0, 0.5, 1.5])
    tg, opt_sol = pend.cocSolver([0.0, 0.0], 1, th)
    tgo, X, U, Lam = orc.solve([0.0, 0.0], 1.0, th)
    pend.aux_mode = pend.MODE_BDF
    pend.rtol_back, pend.atol_back, pend.rtol_fwd, pend.atol_fwd = 1e-3, 1e-6, 1e-3, 1e-6
    aux_sol = pend.auxSysSolver(tg, opt_sol, th)
    assert pend.last_aux_status == 0
    Xa, Ua, PW = orc.aux(tgo, X, U, Lam, th)          # as shipped: BDF backward, RK45 forward
    got = aux_sol(tg)
    assert _rel(got[:, :6], Xa) < 1e-8 and _rel(got[:, 6:], Ua) < 1e-8
    # step / LU / rhs counters against scipy's BDF class driven interval by interval
    from scipy.interpolate import interp1d
    osol = interp1d(tgo, np.concatenate((X, U, Lam), axis=1), axis=0)

    def rhs(t, y):
        v = osol(t)
        Pd, Wd = orc.riccati_rhs(v[:2], v[2:3], v[3:], th, y[:4].reshape(2, 2), y[4:].reshape(2, 3))
        return np.concatenate((Pd.ravel(), Wd.ravel()))
    nsteps = nlu = 0
    y = PW[-1].copy()
    for k in range(10, 0, -1):
        s = BDF(rhs, tgo[k], y, tgo[k - 1])
        while s.status == "running":
            s.step()
            nsteps += 1
        nlu += s.nlu
        y = s.y
    c = pend.last_aux_counters
    assert (c[1], c[4]) == (nsteps, nlu), (c, nsteps, nlu)


def test_quad_k4_gradient_bdf_vs_stored_run(quad):
    """KAT K4 through the emulated CUDA path in as-shipped mode: dL/dtheta(theta0) implied by the reference's stored
    parameter_trace, tolerance 1e-5 relative (north_star); measured 1.8e-7."""
    g = np.load(os.path.join(HERE, "golden", "quad_run.npz"))
    quad.setIntegrator(n_grid=25)
    quad.aux_mode = quad.MODE_BDF
    quad.rtol_back, quad.atol_back, quad.rtol_fwd, quad.atol_fwd = 1e-3, 1e-6, 1e-3, 1e-6
    P, lr = g["parameter_trace"], float(g["learning_rate"])
    sol = quad.cocSolverBatch(g["ini_state"].reshape(1, 13), 1.0, P[0], pdata=g["goal_position"].reshape(1, 3))
    aux = quad.auxSysSolverBatch(sol, g["time_grid"], g["waypoints"].reshape(1, -1, 3), [0, 1, 2])
    assert int(aux["aux_status"][0]) == 0
    assert abs(aux["loss"][0] - g["loss_trace"][0]) / g["loss_trace"][0] < 1e-8
    assert _rel(aux["dtheta"][0], (P[0] - P[1]) / lr) < 1e-5


def test_pendulum_timewarping_vs_oracle():
    """COCSys_TimeVarying (reference CPDP.py:394-787, Examples/pendulum_timewarping.py with v(t) = beta1 + 2 beta2 t):
    frozen-time RK4 map + linspace grid in the solve, RK45 for both sweeps, the time reaching every model function."""
    oc = emu_oc("pendulum_tw2")
    assert type(oc).__name__ == "COCSys_TimeVarying" and oc.aux_mode == oc.MODE_RK45
    orc = Oracle(models.pendulum_timewarp(2), n_grid=10)
    th, T = np.array([2.5, 0.7, 0.8, 1.3]), 0.2
    tg, opt_sol = oc.cocSolver([0.0, 0.0], T, th)
    tgo, X, U, Lam, info = orc.solve([0.0, 0.0], T, th, return_info=True)
    assert oc.last_status == 1 and oc.last_iters == info["iters"]
    assert np.array_equal(tg, tgo) and np.array_equal(tg, np.linspace(0, T, 11))
    nodes = opt_sol(tg)
    assert np.abs(nodes[:, :2] - X).max() < 1e-10 and np.abs(nodes[:, 2:3] - U).max() < 1e-9 and np.abs(nodes[:, 3:] - Lam).max() < 1e-9
    aux_sol = oc.auxSysSolver(tg, opt_sol, th)
    Xa, Ua, PW, cnt = orc.aux(tgo, X, U, Lam, th, return_counts=True)          # default for the class: RK45 / RK45
    got = aux_sol(tg)
    assert oc.last_aux_status == 0
    assert (oc.last_aux_counters[0], oc.last_aux_counters[2]) == (cnt["back_rhs"], cnt["fwd_rhs"])
    assert _rel(got[:, :8], Xa) < 1e-9 and _rel(got[:, 8:], Ua) < 1e-9
    # the time really enters: beta2 moves the solution
    tg2, sol2 = oc.cocSolver([0.0, 0.0], T, np.array([2.5, 0.0, 0.8, 1.3]))
    assert np.abs(sol2(tg2) - nodes).max() > 1e-3


def test_cartpole_through_codegen_vs_oracle():
    """a JinEnv definition no example script uses (CartPole) goes through the same code generation and
    kernels; checked against the oracle's independent restatement of the model."""
    cp = emu_oc("cartpole")
    orc = Oracle(models.cartpole(), n_grid=20)
    th = np.array([1.5, 0.5, 1.0, 0.2, 0.3])
    x0 = np.array([0.0, 0.3, 0.0, 0.0])
    sol = cp.cocSolverBatch(x0.reshape(1, 4), 1.0, th)
    tg, X, U, Lam, info = orc.solve(x0, 1.0, th, return_info=True)
    assert int(sol["status"][0]) == 1 and int(sol["iters"][0]) == info["iters"]
    assert np.abs(sol["X"][0] - X).max() < 1e-8 * max(1.0, np.abs(X).max())
    assert np.abs(sol["Lam"][0] - Lam).max() < 1e-7 * max(1.0, np.abs(Lam).max())
    taus, wp = np.array([0.25, 0.7]), np.array([[[0.1, 1.0], [0.0, 2.5]]])
    for mode, back in ((cp.MODE_RK45, {}), (cp.MODE_BDF, {'method': 'BDF', 'jac': 'closed'})):
        cp.aux_mode = mode
        cp.rtol_back, cp.atol_back, cp.rtol_fwd, cp.atol_fwd = 1e-3, 1e-6, 1e-3, 1e-6
        aux = cp.auxSysSolverBatch(sol, taus, wp, [0, 1])
        assert int(aux["aux_status"][0]) == 0
        Xa, Ua, PW = orc.aux(tg, X, U, Lam, th, back=back, fwd={})
        loss, dl = orc.loss_grad(taus, wp[0], tg, X, Xa, sel=[0, 1])
        assert abs(aux["loss"][0] - loss) < 1e-9 * max(1.0, loss)
        assert _rel(aux["dtheta"][0], dl) < 1e-6, (mode, aux["dtheta"][0], dl)
