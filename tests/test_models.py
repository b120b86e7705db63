"""Cross-checks the product's JinEnv/sx layer against the oracle's independent sympy restatement, and the generated
C against sympy (compiled for the host with gcc through the emulation build)."""
import numpy as np
import pytest
import sympy as sp

import lfsd_b200  # noqa: F401
from lfsd_b200 import standard
from lfsd_b200.sx import _to_matrix
from oracle import models


@pytest.mark.parametrize("name", ["pendulum", "robotarm", "rocket", "quadrotor", "cartpole", "pendulum_tw2"])
def test_product_models_equal_oracle_models(name):
    oc = standard.STANDARD[name]()
    om = models.pendulum_timewarp(2) if name == "pendulum_tw2" else getattr(models, name)()
    rng = np.random.default_rng(0)
    prod = [oc.dyn, oc.path_cost, oc.final_cost]
    orac = [om.dyn, sp.Matrix([om.path]), sp.Matrix([om.final])]
    psyms = sorted(set().union(*[_to_matrix(e).free_symbols for e in prod]), key=lambda s: s.name)
    osyms = sorted(set().union(*[e.free_symbols for e in orac]), key=lambda s: s.name)
    assert [s.name for s in psyms] == [s.name for s in osyms]
    assert [s.name for s in list(_to_matrix(oc.auxvar))] == [s.name for s in om.theta]
    assert [s.name for s in list(_to_matrix(oc.state))] == [s.name for s in om.x]
    for _ in range(5):
        vals = rng.normal(size=len(psyms))
        for pe, oe in zip(prod, orac):
            a = np.array(_to_matrix(pe).subs(dict(zip(psyms, vals))).evalf(30), dtype=float)
            b = np.array(oe.subs(dict(zip(osyms, vals))).evalf(30), dtype=float)
            assert np.allclose(a, b, rtol=1e-13, atol=1e-13)


def test_reference_parameter_orders():
    """cost_auxvar orders that define the meaning of theta (JinEnv.py:307-320, 911-933, 1403-1435)."""
    assert [str(s) for s in _to_matrix(standard.robotarm_oc().env.cost_auxvar)] == ['w_q1_sq', 'w_q1', 'w_q2_sq', 'w_q2']
    assert [str(s) for s in _to_matrix(standard.quadrotor_oc().env.cost_auxvar)] == \
        ['w_xsq', 'w_x', 'w_ysq', 'w_y', 'w_zsq', 'w_z']
    r = [str(s) for s in _to_matrix(standard.rocket_oc().env.cost_auxvar)]
    assert r[-2:] == ['wsidethrust', 'wtilt'] and len(r) == 11


def test_incomplete_definition_raises():
    """Same precondition asserts as the reference (CPDP.py:93-97)."""
    from lfsd_b200.CPDP import COCSys
    oc = COCSys()
    with pytest.raises(AssertionError, match="state variable"):
        oc.cocSolver([0.0], 1.0, [1.0])
