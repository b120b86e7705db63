"""the JinEnv cost definitions no example script of the reference uses -- RobotArm.initCost_WeightedDistance
(JinEnv.py:239-285), Quadrotor.initCost / initCost2 (:755-884), Rocket.initCost / initCost_Ex (:1328-1399, 1475-1551) -- through
the product's JinEnv layer, the code generator and the kernels, against independent restatements in oracle/models.py."""
import numpy as np
import pytest
import sympy as sp

import lfsd_b200  # noqa: F401
from lfsd_b200 import standard, synthetic
from lfsd_b200.sx import _to_matrix
from oracle import models
from oracle.cpdp_oracle import Oracle


def cases():
    qb = synthetic.quad_batch(1)
    rb = synthetic.rocket_batch(1)
    return {
        "robotarm_wd": (lambda: models.robotarm(cost='weighted_distance'), 20, [3.0, 60., 40., 5., 5.], np.array([-np.pi / 2, 0, 0, 0.]), None, 1.0,
                        np.array([0.3, 0.8]), np.array([[-0.6, 0.9], [1.2, 0.1]])),
        "quadrotor_cost1": (lambda: models.quadrotor(cost='weighted'), 15, [2.0, 1.5, 8., 60., 6.], qb["x0"][0], qb["goal"][0], 1.0,
                            qb["taus"], qb["wp"][0]),
        "quadrotor_cost2": (lambda: models.quadrotor(cost='per_axis'), 15, [2.0, 1., 2., 1.5, 8., 9., 10., 5., 6., 7., 60.], qb["x0"][0],
                            qb["goal"][0], 1.0, qb["taus"], qb["wp"][0]),
        "rocket_cost1": (lambda: models.rocket(cost='scalar'), 15, [1.5, 1., 1.2, 2., 0.8, 1.1], rb["x0"][0], None, 3.0,
                         np.array([0.6, 1.8, 2.6]), np.tile(np.array([[8.0, -6.0, 2.0, 0.9, 0.0, -0.3, 0.3]]), (3, 1)) * np.array([[1.0], [0.5], [0.2]])),
    }


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


@pytest.mark.parametrize("name", ["robotarm_wd", "quadrotor_cost1", "quadrotor_cost2", "rocket_cost1"])
def test_variant_models_equal_oracle_restatements(name):
    oc = standard.VARIANTS[name]()
    om = cases()[name][0]()
    prod = [oc.dyn, oc.path_cost, oc.final_cost]
    orac = [om.dyn, sp.Matrix([om.path]), sp.Matrix([om.final])]
    psyms = sorted(set().union(*[_to_matrix(e).free_symbols for e in prod]), key=lambda s: s.name)
    osyms = sorted(set().union(*[e.free_symbols for e in orac]), key=lambda s: s.name)
    assert [s.name for s in psyms] == [s.name for s in osyms]
    assert [s.name for s in list(_to_matrix(oc.auxvar))] == [s.name for s in om.theta]       # the order that defines theta
    rng = np.random.default_rng(1)
    for _ in range(3):
        vals = rng.normal(size=len(psyms))
        for pe, oe in zip(prod, orac):
            a = np.array(_to_matrix(pe).subs(dict(zip(psyms, vals))).evalf(30), dtype=float)
            b = np.array(oe.subs(dict(zip(osyms, vals))).evalf(30), dtype=float)
            assert np.allclose(a, b, rtol=1e-13, atol=1e-13)


def test_rocket_initcost_ex_is_rejected_like_in_the_reference():
    """initCost_Ex puts the control-dependent side-thrust term into the final cost (JinEnv.py:1548-1551); the reference's
    setFinalCost wraps it in a CasADi Function of (state, auxvar) only, which raises on the free control symbols.  Here the code
    generator refuses it."""
    from lfsd_b200 import JinEnv
    env = JinEnv.Rocket()
    env.initDyn(Jx=1, Jy=1, Jz=1, mass=1, l=1)
    env.initCost_Ex(wthrust=0.1)
    assert [str(s) for s in _to_matrix(env.cost_auxvar)][-2:] == ['wtilt', 'wsidethrust']
    oc = standard._wrap(env, "rocket_costex", 15)
    with pytest.raises(AssertionError, match="final cost must not depend on the control"):
        oc.build()


def _check(oc, name, grad_tol):
    mk, N, th, x0, pd, T, taus, wp = cases()[name]
    orc = Oracle(mk(), n_grid=N)
    if pd is not None:
        orc.pd = np.asarray(pd, dtype=float)
    th = np.array(th)
    sol = oc.cocSolverBatch(np.asarray(x0).reshape(1, -1), T, th, pdata=None if pd is None else np.asarray(pd).reshape(1, -1))
    tg, X, U, Lam, info = orc.solve(x0, T, th, return_info=True)
    tonp = (lambda t: t.detach().cpu().numpy()) if hasattr(sol["X"], "detach") else np.asarray
    assert int(tonp(sol["status"])[0]) == 1 and int(tonp(sol["iters"])[0]) == info["iters"]
    assert _rel(tonp(sol["X"])[0], X) < 1e-6 and _rel(tonp(sol["U"])[0], U) < 1e-6 and _rel(tonp(sol["Lam"])[0], Lam) < 1e-6
    for mode, back in ((oc.MODE_RK45, {}), (oc.MODE_BDF, {'method': 'BDF', 'jac': 'closed'})):
        oc.aux_mode = mode
        oc.rtol_back, oc.atol_back, oc.rtol_fwd, oc.atol_fwd = 1e-3, 1e-6, 1e-3, 1e-6
        aux = oc.auxSysSolverBatch(sol, taus, wp.reshape(1, len(taus), -1), oc.sel)
        assert int(tonp(aux["aux_status"])[0]) == 0
        Xa, Ua, PW = orc.aux(tg, X, U, Lam, th, back=back, fwd={})
        loss, dl = orc.loss_grad(taus, wp, tg, X, Xa, sel=oc.sel)
        assert abs(tonp(aux["loss"])[0] - loss) < 1e-8 * max(1.0, loss)
        assert _rel(tonp(aux["dtheta"])[0], dl) < grad_tol, (name, mode, tonp(aux["dtheta"])[0], dl)


@pytest.mark.parametrize("name", ["robotarm_wd", "quadrotor_cost1"])
def test_variants_through_emulated_kernels(name):
    from tests.emu.support import emu_oc
    _check(emu_oc(name), name, 1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["robotarm_wd", "quadrotor_cost1", "quadrotor_cost2", "rocket_cost1"])
def test_variants_on_gpu_vs_live_oracle(name):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    oc = standard.VARIANTS[name]()
    oc.build(name=oc.lib_name)
    _check(oc, name, 1e-5)
