"""The optimal-control systems of the reference's example scripts, built exactly as the scripts build them
(time-warping wrapper: dyn = beta*f, path = beta*c, final cost unscaled, theta = [beta; cost_auxvar]).
Each function returns a COCSys with a stable library name so that ``__graft_entry__.build()`` can prebuild it.
"""
import math

from . import JinEnv
from .CPDP import COCSys, COCSys_TimeVarying
from .sx import SX, vertcat


def _wrap(env, name, n_grid):
    """/root/reference/Examples/pendulum_groundtruth.py:21-30 (identical in every example)."""
    oc = COCSys(name)
    beta = SX.sym('beta')
    oc.setAuxvarVariable(vertcat(beta, env.cost_auxvar))
    oc.setStateVariable(env.X)
    oc.setControlVariable(env.U)
    oc.setDyn(beta * env.f)
    oc.setPathCost(beta * env.path_cost)
    oc.setFinalCost(env.final_cost)
    oc.setIntegrator(n_grid=n_grid)
    oc.env = env
    oc.lib_name = name
    return oc


def pendulum_oc(n_grid=10):
    """Examples/pendulum_groundtruth.py:16-18 ; observed: q (x[0])."""
    env = JinEnv.SinglePendulum()
    env.initDyn(l=1, m=1, damping_ratio=0.1)
    env.initCost(wu=.01)
    oc = _wrap(env, "pendulum", n_grid)
    oc.sel = [0]
    return oc


def robotarm_oc(n_grid=30):
    """Examples/robotarm_random.py:14-29 ; observed: q1, q2."""
    env = JinEnv.RobotArm()
    env.initDyn(l1=1, m1=1, l2=1, m2=1, g=0)
    env.initCost_Polynomial(wu=.5)
    oc = _wrap(env, "robotarm", n_grid)
    oc.sel = [0, 1]
    return oc


def cartpole_oc(n_grid=20):
    """JinEnv.CartPole (JinEnv.py:499-574) behind the same time-warping wrapper; no example script of the reference uses
    it, constants as in the upstream PDP examples (mc = mp = 0.5, l = 1, wu = 0.1); observed: x, q."""
    env = JinEnv.CartPole()
    env.initDyn(mc=0.5, mp=0.5, l=1)
    env.initCost(wu=0.1)
    oc = _wrap(env, "cartpole", n_grid)
    oc.sel = [0, 1]
    return oc


def rocket_oc(n_grid=15):
    """Examples/rocket_groundtruth.py:15-30 ; observed: position and quaternion."""
    env = JinEnv.Rocket()
    env.initDyn(Jx=1, Jy=1, Jz=1, mass=1, l=1)
    env.initCost2(wthrust=0.1)
    oc = _wrap(env, "rocket", n_grid)
    oc.sel = [0, 1, 2, 6, 7, 8, 9]
    return oc


def quadrotor_oc(n_grid=25, goal_position=None):
    """lib/QuadAlgorithm.py:74-104 with QuadPara(1,1,1, mass=1, l=1, c=0.02) (Examples/quad_example.py:26).
    With goal_position=None the goal position is per-problem data (3 constants), which is what the batched
    benchmark needs; a numeric goal is baked into the model like the reference does."""
    env = JinEnv.Quadrotor()
    env.initDyn(Jx=1.0, Jy=1.0, Jz=1.0, mass=1.0, l=1.0, c=0.02)
    goal = JinEnv.QuadStates()
    pvar = None
    if goal_position is None:
        pvar = vertcat(SX.sym('goal_x'), SX.sym('goal_y'), SX.sym('goal_z'))
        goal.position = pvar
        name = "quadrotor"
    else:
        goal.position = list(goal_position)
        name = "quadrotor_fixedgoal"
    env.initCost_Polynomial(goal, w_thrust=0.1)
    oc = _wrap(env, name, n_grid)
    if pvar is not None:
        oc.setProblemVariable(pvar)
    oc.sel = [0, 1, 2]
    return oc


def pendulum_timewarp_oc(n_grid=10, order=2):
    """Examples/pendulum_timewarping.py:15-56: COCSys_TimeVarying with the polynomial time-warping speed
    v(t) = beta1 + 2 beta2 t + ... of the given order (the script ships order 1 active and orders 2-4 commented out; order 2
    is the default here because it is the smallest one in which the time enters the model); observed: q."""
    env = JinEnv.SinglePendulum()
    env.initDyn(l=1, m=1, damping_ratio=0.1)
    env.initCost(wu=.01)
    name = "pendulum_tw%d" % order
    oc = COCSys_TimeVarying(name)
    oc.setStateVariable(env.X)
    oc.setControlVariable(env.U)
    t = SX.sym('t')
    oc.setTimeVariable(t)
    betas = [SX.sym('beta%d' % (i + 1)) for i in range(order)]
    oc.setAuxvarVariable(vertcat(*betas, env.cost_auxvar))
    v = betas[0]
    for i in range(1, order):
        v = v + (i + 1) * betas[i] * t ** i
    oc.setDyn(v * env.f)
    oc.setPathCost(v * env.path_cost)
    oc.setFinalCost(env.final_cost)
    oc.setIntegrator(n_grid=n_grid)
    oc.env = env
    oc.lib_name = name
    oc.sel = [0]
    return oc


# ---- the JinEnv cost definitions no example script uses, behind the same wrapper ----------------------
def robotarm_wd_oc(n_grid=20):
    """RobotArm.initCost_WeightedDistance (JinEnv.py:239-285): theta = [beta, wq1, wq2, wdq1, wdq2]."""
    env = JinEnv.RobotArm()
    env.initDyn(l1=1, m1=1, l2=1, m2=1, g=0)
    env.initCost_WeightedDistance(wu=.5)
    oc = _wrap(env, "robotarm_wd", n_grid)
    oc.sel = [0, 1]
    return oc


def _quad_env():
    env = JinEnv.Quadrotor()
    env.initDyn(Jx=1.0, Jy=1.0, Jz=1.0, mass=1.0, l=1.0, c=0.02)
    goal = JinEnv.QuadStates()
    pvar = vertcat(SX.sym('goal_x'), SX.sym('goal_y'), SX.sym('goal_z'))
    goal.position = pvar
    return env, goal, pvar


def quadrotor_cost1_oc(n_grid=15):
    """Quadrotor.initCost (JinEnv.py:755-815): theta = [beta, wr, wv, wq, ww]; goal position per problem."""
    env, goal, pvar = _quad_env()
    env.initCost(goal, wthrust=0.1)
    oc = _wrap(env, "quadrotor_cost1", n_grid)
    oc.setProblemVariable(pvar)
    oc.sel = [0, 1, 2]
    return oc


def quadrotor_cost2_oc(n_grid=15):
    """Quadrotor.initCost2 (JinEnv.py:817-884): theta = [beta, wrx..wwz (9), wq]; goal position per problem."""
    env, goal, pvar = _quad_env()
    env.initCost2(goal, wthrust=0.1)
    oc = _wrap(env, "quadrotor_cost2", n_grid)
    oc.setProblemVariable(pvar)
    oc.sel = [0, 1, 2]
    return oc


def rocket_cost1_oc(n_grid=15):
    """Rocket.initCost (JinEnv.py:1328-1399): theta = [beta, wr, wv, wtilt, wsidethrust, ww].
    (Rocket.initCost_Ex, :1475-1551, puts the control-dependent side-thrust term into the FINAL cost: the reference's own
    setFinalCost builds a CasADi Function of (state, auxvar) only and rejects it; here codegen asserts the same.)"""
    env = JinEnv.Rocket()
    env.initDyn(Jx=1, Jy=1, Jz=1, mass=1, l=1)
    env.initCost(wthrust=0.1)
    oc = _wrap(env, "rocket_cost1", n_grid)
    oc.sel = [0, 1, 2, 6, 7, 8, 9]
    return oc


VARIANTS = {"robotarm_wd": robotarm_wd_oc, "quadrotor_cost1": quadrotor_cost1_oc, "quadrotor_cost2": quadrotor_cost2_oc,
            "rocket_cost1": rocket_cost1_oc}

STANDARD = {"pendulum": pendulum_oc, "pendulum_tw2": pendulum_timewarp_oc, "robotarm": robotarm_oc, "rocket": rocket_oc, "quadrotor": quadrotor_oc,
            "cartpole": cartpole_oc}


def build_all(verbose=False):
    libs = {}
    for name, fn in list(STANDARD.items()) + list(VARIANTS.items()):
        oc = fn()
        libs[name] = oc.build(name=oc.lib_name, verbose=verbose).path
    return libs
