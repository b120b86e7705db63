"""JinEnv dynamics and cost definitions (symbolic), with the reference's class / method / attribute
names so that example scripts read the same (``/root/reference/JinEnv/JinEnv.py``).

Only the symbolic definitions are on the CPDP hot path; the matplotlib
animation / plotting facilities of the reference are out of scope and are not provided.
Each ``initDyn`` / ``initCost*`` keeps the reference convention: an argument left as ``None`` becomes
a learnable symbol appended (in argument order) to ``dyn_auxvar`` / ``cost_auxvar``.
"""
import math
from dataclasses import dataclass, field

import numpy as np

from .sx import SX, vertcat, vcat, horzcat, mtimes, dot, pinv, diag, transpose, trace, sin, cos


class _Params:
    """Collects ``None`` arguments as fresh symbols (the reference's 'parameter' lists)."""

    def __init__(self):
        self.syms = []

    def __call__(self, name, value):
        if value is None:
            s = SX.sym(name)
            self.syms.append(s)
            return s
        return value

    def new(self, name):
        return self(name, None)

    def vec(self):
        return vcat(self.syms)


@dataclass
class QuadStates:
    """Plain container, ``/root/reference/lib/QuadStates.py:5-14``."""
    position: list = field(default_factory=lambda: [0, 0, 0])
    velocity: list = field(default_factory=lambda: [0, 0, 0])
    attitude_quaternion: list = field(default_factory=lambda: [1, 0, 0, 0])
    angular_velocity: list = field(default_factory=lambda: [0, 0, 0])


# ------------------------------------------------------------------------------------------------
# attitude helpers shared by Quadrotor and Rocket (JinEnv.py:1182-1212, 1552-1576)
# ------------------------------------------------------------------------------------------------
def _dir_cosine(q):
    """Direction cosine matrix, inertial -> body, scalar-first quaternion."""
    q0, q1, q2, q3 = q[0], q[1], q[2], q[3]
    return vertcat(
        horzcat(1 - 2 * (q2 ** 2 + q3 ** 2), 2 * (q1 * q2 + q0 * q3), 2 * (q1 * q3 - q0 * q2)),
        horzcat(2 * (q1 * q2 - q0 * q3), 1 - 2 * (q1 ** 2 + q3 ** 2), 2 * (q2 * q3 + q0 * q1)),
        horzcat(2 * (q1 * q3 + q0 * q2), 2 * (q2 * q3 - q0 * q1), 1 - 2 * (q1 ** 2 + q2 ** 2)))


def _skew(v):
    return vertcat(horzcat(0, -v[2], v[1]),
                   horzcat(v[2], 0, -v[0]),
                   horzcat(-v[1], v[0], 0))


def _omega(w):
    return vertcat(horzcat(0, -w[0], -w[1], -w[2]),
                   horzcat(w[0], 0, w[2], -w[1]),
                   horzcat(w[1], -w[2], 0, w[0]),
                   horzcat(w[2], w[1], -w[0], 0))


def _as_vec(v):
    """list / ndarray / SX -> SX column; entries may be numbers or symbols (a symbolic goal becomes per-problem
    data of the batched solver, see COCSys.setProblemVariable)."""
    if isinstance(v, SX):
        return v
    return vertcat(*list(v))


def _sq_err(vec, goal):
    d = vec - _as_vec(goal)
    return dot(d, d)


class _RigidBody6DoF:
    """State layout [r_I(3), v_I(3), q(4, scalar first), w_B(3)] (JinEnv.py:666-678, 1252-1264)."""

    def _declare_state(self):
        self.r_I = vertcat(SX.sym('rx'), SX.sym('ry'), SX.sym('rz'))
        self.v_I = vertcat(SX.sym('vx'), SX.sym('vy'), SX.sym('vz'))
        self.q = vertcat(SX.sym('q0'), SX.sym('q1'), SX.sym('q2'), SX.sym('q3'))
        self.w_B = vertcat(SX.sym('wx'), SX.sym('wy'), SX.sym('wz'))

    dir_cosine = staticmethod(_dir_cosine)
    skew = staticmethod(_skew)
    omega = staticmethod(_omega)

    @staticmethod
    def quaternion_mul(p, q):
        return vertcat(p[0] * q[0] - p[1] * q[1] - p[2] * q[2] - p[3] * q[3],
                       p[0] * q[1] + p[1] * q[0] + p[2] * q[3] - p[3] * q[2],
                       p[0] * q[2] - p[1] * q[3] + p[2] * q[0] + p[3] * q[1],
                       p[0] * q[3] + p[1] * q[2] - p[2] * q[1] + p[3] * q[0])

    def _rigid_body_ode(self, force_I_over_m, moment_B):
        """ṙ=v, v̇=F/m+g, q̇=½Ω(w)q, ẇ=J⁻¹(M − w×Jw)."""
        dq = 1 / 2 * mtimes(self.omega(self.w_B), self.q)
        dw = mtimes(pinv(self.J_B), moment_B - mtimes(mtimes(self.skew(self.w_B), self.J_B), self.w_B))
        self.X = vertcat(self.r_I, self.v_I, self.q, self.w_B)
        self.f = vertcat(self.v_I, force_I_over_m + self.g_I, dq, dw)


# ------------------------------------------------------------------------------------------------
class SinglePendulum:
    """n=2, m=1 (JinEnv.py:40-107)."""

    def __init__(self, project_name='single pendlumn system'):
        self.project_name = project_name

    def initDyn(self, l=None, m=None, damping_ratio=None):
        g = 10
        p = _Params()
        self.l, self.m, self.damping_ratio = p('l', l), p('m', m), p('damping_ratio', damping_ratio)
        self.dyn_auxvar = p.vec()
        self.q, self.dq = SX.sym('q'), SX.sym('dq')
        self.X = vertcat(self.q, self.dq)
        self.U = SX.sym('u')
        inertia = 1 / 3 * self.m * self.l * self.l
        self.f = vertcat(self.dq,
                         (self.U - self.m * g * self.l * sin(self.q) - self.damping_ratio * self.dq) / inertia)

    def initCost(self, wq=None, wdq=None, wu=0.001):
        p = _Params()
        self.wq, self.wdq = p('wq', wq), p('wdq', wdq)
        self.cost_auxvar = p.vec()
        self.cost_q = (self.q - math.pi) ** 2
        self.cost_dq = (self.dq - 0) ** 2
        self.cost_u = dot(self.U, self.U)
        self.final_cost = self.wq * self.cost_q + self.wdq * self.cost_dq
        self.path_cost = self.final_cost + wu * self.cost_u


class RobotArm:
    """Two-link arm, n=4 ([q1,q2,dq1,dq2]), m=2 (JinEnv.py:178-326)."""

    def __init__(self, project_name='two-link robot arm'):
        self.project_name = project_name

    def initDyn(self, l1=None, m1=None, l2=None, m2=None, g=10):
        p = _Params()
        self.l1, self.m1, self.l2, self.m2 = p('l1', l1), p('m1', m1), p('l2', l2), p('m2', m2)
        self.dyn_auxvar = p.vec()
        self.q1, self.dq1, self.q2, self.dq2 = SX.sym('q1'), SX.sym('dq1'), SX.sym('q2'), SX.sym('dq2')
        self.X = vertcat(self.q1, self.q2, self.dq1, self.dq2)
        self.U = vertcat(SX.sym('u1'), SX.sym('u2'))
        l1_, m1_, l2_, m2_ = self.l1, self.m1, self.l2, self.m2
        r1, r2 = l1_ / 2, l2_ / 2
        I1, I2 = l1_ * l1_ * m1_ / 12, l2_ * l2_ * m2_ / 12
        c2 = cos(self.q2)
        M11 = m1_ * r1 * r1 + I1 + m2_ * (l1_ * l1_ + r2 * r2 + 2 * l1_ * r2 * c2) + I2
        M12 = m2_ * (r2 * r2 + l1_ * r2 * c2) + I2
        M22 = m2_ * r2 * r2 + I2
        M = vertcat(horzcat(M11, M12), horzcat(M12, M22))
        h = m2_ * l1_ * r2 * sin(self.q2)
        C = vertcat(-h * self.dq2 * self.dq2 - 2 * h * self.dq1 * self.dq2, h * self.dq1 * self.dq1)
        c12 = cos(self.q1 + self.q2)
        G = vertcat(m1_ * r1 * g * cos(self.q1) + m2_ * g * (r2 * c12 + l1_ * cos(self.q1)),
                    m2_ * g * r2 * c12)
        ddq = mtimes(pinv(M), -C - G + self.U)
        self.f = vertcat(self.dq1, self.dq2, ddq)

    def _goal_costs(self):
        goal = [math.pi / 2, 0, 0, 0]
        return [(s - g_) ** 2 for s, g_ in zip((self.q1, self.q2, self.dq1, self.dq2), goal)]

    def initCost_WeightedDistance(self, wq1=None, wq2=None, wdq1=None, wdq2=None, wu=0.1):
        p = _Params()
        self.wq1, self.wq2, self.wdq1, self.wdq2 = p('wq1', wq1), p('wq2', wq2), p('wdq1', wdq1), p('wdq2', wdq2)
        self.cost_auxvar = p.vec()
        self.cost_q1, self.cost_q2, self.cost_dq1, self.cost_dq2 = self._goal_costs()
        self.cost_u = dot(self.U, self.U)
        self.final_cost = self.wq1 * self.cost_q1 + self.wq2 * self.cost_q2 + \
            self.wdq1 * self.cost_dq1 + self.wdq2 * self.cost_dq2
        self.path_cost = self.final_cost + wu * self.cost_u

    def initCost_Polynomial(self, wu=0.1):
        p = _Params()
        self.cost_goal_q1, self.cost_goal_q2, self.cost_goal_dq1, self.cost_goal_dq2 = self._goal_costs()
        self.cost_u = dot(self.U, self.U)
        # parameter order [w_q1_sq, w_q1, w_q2_sq, w_q2] (JinEnv.py:307-320)
        self.w_q1_sq = p.new('w_q1_sq'); self.feature_q1_sq = 0.5 * self.q1 * self.q1
        self.w_q1 = p.new('w_q1'); self.feature_q1 = self.q1
        self.w_q2_sq = p.new('w_q2_sq'); self.feature_q2_sq = 0.5 * self.q2 * self.q2
        self.w_q2 = p.new('w_q2'); self.feature_q2 = self.q2
        self.path_cost = self.w_q1 * self.feature_q1 + self.w_q1_sq * self.feature_q1_sq + \
            self.w_q2 * self.feature_q2 + self.w_q2_sq * self.feature_q2_sq + wu * self.cost_u
        self.final_cost = 100 * self.cost_goal_q1 + 100 * self.cost_goal_q2 + \
            100 * self.cost_goal_dq1 + 100 * self.cost_goal_dq2
        self.cost_auxvar = p.vec()


class CartPole:
    """n=4 ([x,q,dx,dq]), m=1 (JinEnv.py:499-574)."""

    def __init__(self, project_name='cart-pole-system'):
        self.project_name = project_name

    def initDyn(self, mc=None, mp=None, l=None):
        g = 10
        p = _Params()
        self.mc, self.mp, self.l = p('mc', mc), p('mp', mp), p('l', l)
        self.dyn_auxvar = p.vec()
        self.x, self.q, self.dx, self.dq = SX.sym('x'), SX.sym('q'), SX.sym('dx'), SX.sym('dq')
        self.X = vertcat(self.x, self.q, self.dx, self.dq)
        self.U = SX.sym('u')
        sq, cq = sin(self.q), cos(self.q)
        ddx = (self.U + self.mp * sq * (self.l * self.dq * self.dq + g * cq)) / (self.mc + self.mp * sq * sq)
        ddq = (-self.U * cq - self.mp * self.l * self.dq * self.dq * sq * cq - (self.mc + self.mp) * g * sq) / \
              (self.l * self.mc + self.l * self.mp * sq * sq)
        self.f = vertcat(self.dx, self.dq, ddx, ddq)

    def initCost(self, wx=None, wq=None, wdx=None, wdq=None, wu=0.001):
        p = _Params()
        self.wx, self.wq, self.wdx, self.wdq = p('wx', wx), p('wq', wq), p('wdx', wdx), p('wdq', wdq)
        self.cost_auxvar = p.vec()
        goal = [0.0, math.pi, 0.0, 0.0]
        self.final_cost = self.wx * (self.x - goal[0]) ** 2 + self.wq * (self.q - goal[1]) ** 2 + \
            self.wdx * (self.dx - goal[2]) ** 2 + self.wdq * (self.dq - goal[3]) ** 2
        self.path_cost = self.final_cost + wu * (self.U * self.U)


class Quadrotor(_RigidBody6DoF):
    """n=13, m=4 rotor thrusts (JinEnv.py:662-953)."""

    def __init__(self, project_name='my UAV'):
        self.project_name = 'my uav'
        self._declare_state()
        self.T_B = vertcat(SX.sym('f1'), SX.sym('f2'), SX.sym('f3'), SX.sym('f4'))

    def initDyn(self, Jx=None, Jy=None, Jz=None, mass=None, l=None, c=None):
        g = 9.81
        p = _Params()
        self.Jx, self.Jy, self.Jz = p('Jx', Jx), p('Jy', Jy), p('Jz', Jz)
        self.mass, self.l, self.c = p('mass', mass), p('l', l), p('c', c)
        self.dyn_auxvar = p.vec()
        self.J_B = diag(vertcat(self.Jx, self.Jy, self.Jz))
        self.g_I = vertcat(0, 0, -g)
        self.m = self.mass
        f = self.T_B
        self.thrust_B = vertcat(0, 0, f[0] + f[1] + f[2] + f[3])
        self.M_B = vertcat(-f[1] * self.l / 2 + f[3] * self.l / 2,
                           -f[0] * self.l / 2 + f[2] * self.l / 2,
                           (f[0] - f[1] + f[2] - f[3]) * self.c)
        C_I_B = transpose(self.dir_cosine(self.q))
        self._rigid_body_ode(1 / self.m * mtimes(C_I_B, self.thrust_B), self.M_B)
        self.U = self.T_B

    def _goal_terms(self, goal: QuadStates):
        self._goal_r = _as_vec(goal.position)
        self._goal_v = _as_vec(goal.velocity)
        self._goal_w = _as_vec(goal.angular_velocity)
        goal_R = self.dir_cosine(_as_vec(goal.attitude_quaternion))
        att = trace(np.identity(3) - mtimes(transpose(goal_R), self.dir_cosine(self.q)))
        return att

    def initCost(self, QuadDesiredStates: QuadStates, wr=None, wv=None, wq=None, ww=None, wthrust=0.1):
        p = _Params()
        self.wr, self.wv, self.wq, self.ww = p('wr', wr), p('wv', wv), p('wq', wq), p('ww', ww)
        self.cost_auxvar = p.vec()
        self.cost_q = self._goal_terms(QuadDesiredStates)
        self.cost_r_I = _sq_err(self.r_I, self._goal_r)
        self.cost_v_I = _sq_err(self.v_I, self._goal_v)
        self.cost_w_B = _sq_err(self.w_B, self._goal_w)
        self.cost_thrust = dot(self.T_B, self.T_B)
        self.final_cost = self.wr * self.cost_r_I + self.wv * self.cost_v_I + \
            self.ww * self.cost_w_B + self.wq * self.cost_q
        self.path_cost = self.final_cost + wthrust * self.cost_thrust

    def initCost2(self, QuadDesiredStates: QuadStates, wthrust=0.1):
        p = _Params()
        # parameter order [wrx,wry,wrz,wvx,wvy,wvz,wwx,wwy,wwz,wq] (JinEnv.py:827-853)
        names = ['wrx', 'wry', 'wrz', 'wvx', 'wvy', 'wvz', 'wwx', 'wwy', 'wwz', 'wq']
        for nm in names:
            setattr(self, nm, p.new(nm))
        self.cost_auxvar = p.vec()
        self.cost_q = self._goal_terms(QuadDesiredStates)
        cost = 0
        for k, ax in enumerate('xyz'):
            cr = (self.r_I[k] - self._goal_r[k]) ** 2
            cv = (self.v_I[k] - self._goal_v[k]) ** 2
            cw = (self.w_B[k] - self._goal_w[k]) ** 2
            setattr(self, 'cost_r_I_' + ax, cr)
            setattr(self, 'cost_v_I_' + ax, cv)
            setattr(self, 'cost_w_B_' + ax, cw)
            cost = cost + getattr(self, 'wr' + ax) * cr + getattr(self, 'wv' + ax) * cv + \
                getattr(self, 'ww' + ax) * cw
        self.cost_thrust = dot(self.T_B, self.T_B)
        self.final_cost = cost + self.wq * self.cost_q
        self.path_cost = self.final_cost + wthrust * self.cost_thrust

    def initCost_Polynomial(self, QuadDesiredStates: QuadStates, w_thrust=0.1):
        p = _Params()
        self.cost_goal_q = self._goal_terms(QuadDesiredStates)
        self.cost_goal_r = _sq_err(self.r_I, self._goal_r)
        self.cost_goal_v = _sq_err(self.v_I, self._goal_v)
        self.cost_goal_w = _sq_err(self.w_B, self._goal_w)
        self.cost_thrust = dot(self.T_B, self.T_B)
        # parameter order [w_xsq,w_x,w_ysq,w_y,w_zsq,w_z] (JinEnv.py:911-933)
        path = 0
        for k, ax in enumerate('xyz'):
            wsq, wl = p.new('w_%ssq' % ax), p.new('w_%s' % ax)
            fsq, fl = 0.5 * self.r_I[k] * self.r_I[k], self.r_I[k]
            setattr(self, 'w_%ssq' % ax, wsq); setattr(self, 'w_%s' % ax, wl)
            setattr(self, 'feature_%ssq' % ax, fsq); setattr(self, 'feature_%s' % ax, fl)
            path = path + wsq * fsq + wl * fl
        self.path_cost = path + w_thrust * self.cost_thrust
        self.final_cost = 1 * self.cost_goal_r + 11 * self.cost_goal_v + \
            100 * self.cost_goal_q + 10 * self.cost_goal_w
        self.cost_auxvar = p.vec()


class Rocket(_RigidBody6DoF):
    """6-DoF powered landing, n=13, m=3 (JinEnv.py:1248-1551)."""

    def __init__(self, project_name='rocket powered landing'):
        self.project_name = project_name
        self._declare_state()
        self.T_B = vertcat(SX.sym('ux'), SX.sym('uy'), SX.sym('uz'))

    def initDyn(self, Jx=None, Jy=None, Jz=None, mass=None, l=None):
        g = 10
        p = _Params()
        self.Jx, self.Jy, self.Jz = p('Jx', Jx), p('Jy', Jy), p('Jz', Jz)
        self.mass, self.l = p('mass', mass), p('l', l)
        self.dyn_auxvar = p.vec()
        self.J_B = diag(vertcat(self.Jx, self.Jy, self.Jz))
        self.g_I = vertcat(-g, 0, 0)
        self.r_T_B = vertcat(-self.l / 2, 0, 0)
        self.m = self.mass
        C_I_B = transpose(self.dir_cosine(self.q))
        self._rigid_body_ode(1 / self.m * mtimes(C_I_B, self.T_B), mtimes(self.skew(self.r_T_B), self.T_B))
        self.U = self.T_B

    def _common_terms(self):
        C_I_B = transpose(self.dir_cosine(self.q))
        body_x_in_I = mtimes(C_I_B, np.array([1., 0., 0.]))
        proj_ny = dot(np.array([0., 1., 0.]), body_x_in_I)
        proj_nz = dot(np.array([0., 0., 1.]), body_x_in_I)
        self.cost_tilt = proj_ny ** 2 + proj_nz ** 2
        self.cost_side_thrust = self.T_B[1] ** 2 + self.T_B[2] ** 2
        self.cost_thrust = dot(self.T_B, self.T_B)

    def initCost(self, wr=None, wv=None, wtilt=None, ww=None, wsidethrust=None, wthrust=1.0):
        p = _Params()
        # symbol order [wr, wv, wtilt, wsidethrust, ww] (JinEnv.py:1331-1362)
        self.wr, self.wv, self.wtilt = p('wr', wr), p('wv', wv), p('wtilt', wtilt)
        self.wsidethrust, self.ww = p('wsidethrust', wsidethrust), p('ww', ww)
        self.cost_auxvar = p.vec()
        self._common_terms()
        self.cost_r_I = _sq_err(self.r_I, [0, 0, 0])
        self.cost_v_I = _sq_err(self.v_I, [0, 0, 0])
        self.cost_w_B = _sq_err(self.w_B, [0, 0, 0])
        self.final_cost = self.wr * self.cost_r_I + self.wv * self.cost_v_I + \
            self.ww * self.cost_w_B + self.wtilt * self.cost_tilt
        self.path_cost = self.final_cost + self.wsidethrust * self.cost_side_thrust + wthrust * self.cost_thrust

    def _per_axis(self, p, tail_names):
        for nm in ['wrx', 'wry', 'wrz', 'wvx', 'wvy', 'wvz', 'wwx', 'wwy', 'wwz'] + tail_names:
            setattr(self, nm, p.new(nm))
        self.cost_auxvar = p.vec()
        self._common_terms()
        cost = 0
        for k, ax in enumerate('xyz'):
            cr, cv, cw = self.r_I[k] ** 2, self.v_I[k] ** 2, self.w_B[k] ** 2
            setattr(self, 'cost_r_I_' + ax, cr)
            setattr(self, 'cost_v_I_' + ax, cv)
            setattr(self, 'cost_w_B_' + ax, cw)
            cost = cost + getattr(self, 'wr' + ax) * cr + getattr(self, 'wv' + ax) * cv + \
                getattr(self, 'ww' + ax) * cw
        return cost

    def initCost2(self, wthrust=0.1):
        """θ_cost = [wrx..wwz (9), wsidethrust, wtilt]; side-thrust only in the path cost (JinEnv.py:1401-1473)."""
        cost = self._per_axis(_Params(), ['wsidethrust', 'wtilt'])
        self.final_cost = cost + self.wtilt * self.cost_tilt
        self.path_cost = self.final_cost + self.wsidethrust * self.cost_side_thrust + wthrust * self.cost_thrust

    def initCost_Ex(self, wthrust=0.1):
        """θ_cost = [wrx..wwz (9), wtilt, wsidethrust]; side-thrust also in the final cost (JinEnv.py:1475-1551)."""
        cost = self._per_axis(_Params(), ['wtilt', 'wsidethrust'])
        self.final_cost = cost + self.wtilt * self.cost_tilt + self.wsidethrust * self.cost_side_thrust
        self.path_cost = self.final_cost + wthrust * self.cost_thrust


# ------------------------------------------------------------------------------------------------
def toQuaternion(angle, dir):
    """(angle, axis) -> scalar-first unit quaternion as a list (JinEnv.py:1730-1737)."""
    d = np.asarray(dir, dtype=float)
    d = d / np.linalg.norm(d)
    return [math.cos(angle / 2)] + (math.sin(angle / 2) * d).tolist()


def normalizeVec(vec):
    v = np.asarray(vec, dtype=float)
    return v / np.linalg.norm(v)


def quaternion_conj(q):
    return [q[0], -q[1], -q[2], -q[3]]
