"""TEST INFRASTRUCTURE — CPU oracle, model definitions.  Not part of the product path.

Independent sympy restatement of the reference's JinEnv models and of the time-warping wrapper every
example puts around them (dyn = beta*f, path = beta*c, final = h unscaled).  Written straight from the
reference formulas, NOT through the product's ``lfsd_b200.sx`` / ``lfsd_b200.JinEnv`` layer, so that
``tests/test_models.py`` can cross-check the two derivations numerically.

Follows:
  * SinglePendulum  /root/reference/JinEnv/JinEnv.py:44-107
  * RobotArm        /root/reference/JinEnv/JinEnv.py:183-237, 287-326
  * Quadrotor       /root/reference/JinEnv/JinEnv.py:680-753, 886-953, 1182-1205
  * Rocket          /root/reference/JinEnv/JinEnv.py:1266-1326, 1401-1473
  * CartPole        /root/reference/JinEnv/JinEnv.py:504-574
  * beta wrapper    /root/reference/Examples/pendulum_groundtruth.py:21-30 (same in every example,
                    /root/reference/lib/QuadAlgorithm.py:89-98)
"""
import math

import numpy as np
import sympy as sp


class OracleModel:
    """x (n), u (m), theta (r) symbols and the three expressions dyn (n), path (scalar), final (scalar).
    ``sel`` lists the state indices observed by the loss closure of the matching example."""

    def __init__(self, name, x, u, theta, dyn, path, final, sel, pdata=()):
        self.name = name
        self.x, self.u, self.theta = list(x), list(u), list(theta)
        self.n, self.m, self.r = len(self.x), len(self.u), len(self.theta)
        self.dyn = sp.Matrix(dyn)
        self.path = sp.sympify(path)
        self.final = sp.sympify(final)
        self.sel = list(sel)
        self.pdata = list(pdata)      # per-problem constants (not learnable), e.g. a symbolic goal position
        self.time = None              # explicit time symbol of a COCSys_TimeVarying model (CPDP.py:434)


def _syms(names):
    return [sp.Symbol(s, real=True) for s in names.split()]


def _dcm(q):
    q0, q1, q2, q3 = q
    return sp.Matrix([
        [1 - 2 * (q2 ** 2 + q3 ** 2), 2 * (q1 * q2 + q0 * q3), 2 * (q1 * q3 - q0 * q2)],
        [2 * (q1 * q2 - q0 * q3), 1 - 2 * (q1 ** 2 + q3 ** 2), 2 * (q2 * q3 + q0 * q1)],
        [2 * (q1 * q3 + q0 * q2), 2 * (q2 * q3 - q0 * q1), 1 - 2 * (q1 ** 2 + q2 ** 2)]])


def _quat_rate(q, w):
    wx, wy, wz = w
    Om = sp.Matrix([[0, -wx, -wy, -wz], [wx, 0, wz, -wy], [wy, -wz, 0, wx], [wz, wy, -wx, 0]])
    return Om * sp.Matrix(q) / 2


def _euler(J, M, w):
    Jm = sp.diag(*J)
    wv = sp.Matrix(w)
    return Jm.inv() * (sp.Matrix(M) - wv.cross(Jm * wv))


def pendulum(l=1.0, m=1.0, damping=0.1, wu=0.01):
    q, dq = _syms('q dq')
    u, = _syms('u')
    beta, wq, wdq = _syms('beta wq wdq')
    g = 10
    inertia = sp.Rational(1, 3) * m * l * l
    f = sp.Matrix([dq, (u - m * g * l * sp.sin(q) - damping * dq) / inertia])
    h = wq * (q - math.pi) ** 2 + wdq * dq ** 2
    c = h + wu * u ** 2
    return OracleModel('pendulum', [q, dq], [u], [beta, wq, wdq], beta * f, beta * c, h, sel=[0])


def pendulum_timewarp(order=2, l=1.0, m=1.0, damping=0.1, wu=0.01):
    """/root/reference/Examples/pendulum_timewarping.py:15-56: the pendulum under COCSys_TimeVarying with the polynomial
    time-warping speed v(t) = beta1 + 2 beta2 t + 3 beta3 t^2 + ... (order 1 is the script's active variant, 2-4 its
    commented ones); dyn = v f, path = v c, final = h."""
    q, dq = _syms('q dq')
    u, = _syms('u')
    t, = _syms('t')
    betas = _syms(' '.join('beta%d' % (i + 1) for i in range(order)))
    wq, wdq = _syms('wq wdq')
    g = 10
    inertia = sp.Rational(1, 3) * m * l * l
    f = sp.Matrix([dq, (u - m * g * l * sp.sin(q) - damping * dq) / inertia])
    h = wq * (q - math.pi) ** 2 + wdq * dq ** 2
    c = h + wu * u ** 2
    v = sum((i + 1) * betas[i] * t ** i for i in range(order))
    mdl = OracleModel('pendulum_tw%d' % order, [q, dq], [u], list(betas) + [wq, wdq], v * f, v * c, h, sel=[0])
    mdl.time = t
    return mdl


def robotarm(l1=1.0, m1=1.0, l2=1.0, m2=1.0, g=0.0, wu=0.5, cost='polynomial'):
    q1, q2, dq1, dq2 = _syms('q1 q2 dq1 dq2')
    u1, u2 = _syms('u1 u2')
    beta, w1s, w1, w2s, w2 = _syms('beta w_q1_sq w_q1 w_q2_sq w_q2')
    r1, r2 = l1 / 2, l2 / 2
    I1, I2 = l1 * l1 * m1 / 12, l2 * l2 * m2 / 12
    M11 = m1 * r1 * r1 + I1 + m2 * (l1 * l1 + r2 * r2 + 2 * l1 * r2 * sp.cos(q2)) + I2
    M12 = m2 * (r2 * r2 + l1 * r2 * sp.cos(q2)) + I2
    M22 = m2 * r2 * r2 + I2
    Mm = sp.Matrix([[M11, M12], [M12, M22]])
    hh = m2 * l1 * r2 * sp.sin(q2)
    C = sp.Matrix([-hh * dq2 * dq2 - 2 * hh * dq1 * dq2, hh * dq1 * dq1])
    G = sp.Matrix([m1 * r1 * g * sp.cos(q1) + m2 * g * (r2 * sp.cos(q1 + q2) + l1 * sp.cos(q1)),
                   m2 * g * r2 * sp.cos(q1 + q2)])
    det = M11 * M22 - M12 * M12
    Minv = sp.Matrix([[M22, -M12], [-M12, M11]]) / det
    ddq = Minv * (-C - G + sp.Matrix([u1, u2]))
    f = sp.Matrix([dq1, dq2, ddq[0], ddq[1]])
    if cost == 'weighted_distance':      # JinEnv.py:239-285 (initCost_WeightedDistance), goal [pi/2, 0, 0, 0]
        wq1, wq2, wdq1, wdq2 = _syms('wq1 wq2 wdq1 wdq2')
        h = wq1 * (q1 - math.pi / 2) ** 2 + wq2 * (q2 - 0) ** 2 + wdq1 * (dq1 - 0) ** 2 + wdq2 * (dq2 - 0) ** 2
        c = h + wu * (u1 * u1 + u2 * u2)
        return OracleModel('robotarm_wd', [q1, q2, dq1, dq2], [u1, u2], [beta, wq1, wq2, wdq1, wdq2],
                           beta * f, beta * c, h, sel=[0, 1])
    c = w1 * q1 + w1s * q1 * q1 / 2 + w2 * q2 + w2s * q2 * q2 / 2 + wu * (u1 ** 2 + u2 ** 2)
    h = 100 * ((q1 - math.pi / 2) ** 2 + q2 ** 2 + dq1 ** 2 + dq2 ** 2)
    return OracleModel('robotarm', [q1, q2, dq1, dq2], [u1, u2], [beta, w1s, w1, w2s, w2],
                       beta * f, beta * c, h, sel=[0, 1])


def quadrotor(goal_r=None, goal_v=(0., 0., 0.), goal_q=(1., 0., 0., 0.), goal_w=(0., 0., 0.),
              J=(1.0, 1.0, 1.0), mass=1.0, l=1.0, c=0.02, w_thrust=0.1, cost='polynomial'):
    r = _syms('rx ry rz'); v = _syms('vx vy vz'); q = _syms('q0 q1 q2 q3'); w = _syms('wx wy wz')
    f_ = _syms('f1 f2 f3 f4')
    beta, wxs, wx_, wys, wy_, wzs, wz_ = _syms('beta w_xsq w_x w_ysq w_y w_zsq w_z')
    pdata = []
    if goal_r is None:            # symbolic goal position -> per-problem data
        pdata = _syms('goal_x goal_y goal_z')
        goal_r = pdata
    thrust = sp.Matrix([0, 0, sum(f_)])
    Mb = [(-f_[1] + f_[3]) * l / 2, (-f_[0] + f_[2]) * l / 2, (f_[0] - f_[1] + f_[2] - f_[3]) * c]
    C_I_B = _dcm(q).T
    dv = C_I_B * thrust / mass + sp.Matrix([0, 0, -9.81])
    f = sp.Matrix([*v, *dv, *_quat_rate(q, w), *_euler(J, Mb, w)])
    path = (wxs * r[0] ** 2 / 2 + wx_ * r[0] + wys * r[1] ** 2 / 2 + wy_ * r[1] + wzs * r[2] ** 2 / 2 + wz_ * r[2]
            + w_thrust * sum(fi ** 2 for fi in f_))
    att = (sp.eye(3) - _dcm(goal_q).T * _dcm(q)).trace()
    thr = sum(fi ** 2 for fi in f_)
    if cost == 'weighted':               # JinEnv.py:755-815 (initCost): scalar weights [wr, wv, wq, ww]
        wr, wv, wq, ww = _syms('wr wv wq ww')
        h = (wr * sum((a - b) ** 2 for a, b in zip(r, goal_r)) + wv * sum((a - b) ** 2 for a, b in zip(v, goal_v))
             + ww * sum((a - b) ** 2 for a, b in zip(w, goal_w)) + wq * att)
        return OracleModel('quadrotor_cost1', r + v + q + w, f_, [beta, wr, wv, wq, ww], beta * f, beta * (h + w_thrust * thr), h,
                           sel=[0, 1, 2], pdata=pdata)
    if cost == 'per_axis':               # JinEnv.py:817-884 (initCost2): [wrx wry wrz wvx wvy wvz wwx wwy wwz wq]
        ws = _syms('wrx wry wrz wvx wvy wvz wwx wwy wwz wq')
        h = (sum(ws[i] * (r[i] - goal_r[i]) ** 2 for i in range(3)) + sum(ws[3 + i] * (v[i] - goal_v[i]) ** 2 for i in range(3))
             + sum(ws[6 + i] * (w[i] - goal_w[i]) ** 2 for i in range(3)) + ws[9] * att)
        return OracleModel('quadrotor_cost2', r + v + q + w, f_, [beta] + ws, beta * f, beta * (h + w_thrust * thr), h,
                           sel=[0, 1, 2], pdata=pdata)
    h = (1 * sum((a - b) ** 2 for a, b in zip(r, goal_r)) + 11 * sum((a - b) ** 2 for a, b in zip(v, goal_v))
         + 100 * att + 10 * sum((a - b) ** 2 for a, b in zip(w, goal_w)))
    return OracleModel('quadrotor', r + v + q + w, f_, [beta, wxs, wx_, wys, wy_, wzs, wz_],
                       beta * f, beta * path, h, sel=[0, 1, 2], pdata=pdata)


def rocket(J=(1.0, 1.0, 1.0), mass=1.0, l=1.0, wthrust=0.1, cost='cost2'):
    r = _syms('rx ry rz'); v = _syms('vx vy vz'); q = _syms('q0 q1 q2 q3'); w = _syms('wx wy wz')
    u = _syms('ux uy uz')
    names = 'wrx wry wrz wvx wvy wvz wwx wwy wwz wsidethrust wtilt'
    beta, = _syms('beta')
    wts = _syms(names)
    wr, wv, ww, wside, wtilt = wts[0:3], wts[3:6], wts[6:9], wts[9], wts[10]
    C_I_B = _dcm(q).T
    T_B = sp.Matrix(u)
    dv = C_I_B * T_B / mass + sp.Matrix([-10, 0, 0])
    r_T = sp.Matrix([-l / 2, 0, 0])
    f = sp.Matrix([*v, *dv, *_quat_rate(q, w), *_euler(J, r_T.cross(T_B), w)])
    bx = C_I_B * sp.Matrix([1, 0, 0])
    tilt = bx[1] ** 2 + bx[2] ** 2
    side, thr = u[1] ** 2 + u[2] ** 2, sum(ui ** 2 for ui in u)
    if cost == 'scalar':                 # JinEnv.py:1328-1399 (initCost): [wr, wv, wtilt, wsidethrust, ww]; side thrust in the path cost only
        swr, swv, swtilt, swside, sww = _syms('wr wv wtilt wsidethrust ww')
        hs = (swr * sum(a ** 2 for a in r) + swv * sum(a ** 2 for a in v) + sww * sum(a ** 2 for a in w) + swtilt * tilt)
        return OracleModel('rocket_cost1', r + v + q + w, u, [beta, swr, swv, swtilt, swside, sww], beta * f,
                           beta * (hs + swside * side + wthrust * thr), hs, sel=[0, 1, 2, 6, 7, 8, 9])
    if cost == 'ex':                     # JinEnv.py:1475-1551 (initCost_Ex): per-axis weights, tilt BEFORE side thrust in theta,
        ex = _syms('wrx wry wrz wvx wvy wvz wwx wwy wwz wtilt wsidethrust')      # and the side-thrust term also in the final cost
        hx = (sum(a * b ** 2 for a, b in zip(ex[0:3], r)) + sum(a * b ** 2 for a, b in zip(ex[3:6], v))
              + sum(a * b ** 2 for a, b in zip(ex[6:9], w)) + ex[9] * tilt + ex[10] * side)
        return OracleModel('rocket_costex', r + v + q + w, u, [beta] + ex, beta * f, beta * (hx + wthrust * thr), hx,
                           sel=[0, 1, 2, 6, 7, 8, 9])
    h = (sum(a * b ** 2 for a, b in zip(wr, r)) + sum(a * b ** 2 for a, b in zip(wv, v))
         + sum(a * b ** 2 for a, b in zip(ww, w)) + wtilt * tilt)
    c = h + wside * (u[1] ** 2 + u[2] ** 2) + wthrust * sum(ui ** 2 for ui in u)
    return OracleModel('rocket', r + v + q + w, u, [beta] + wts, beta * f, beta * c, h,
                       sel=[0, 1, 2, 6, 7, 8, 9])


def cartpole(mc=0.5, mp=0.5, l=1.0, wu=0.1):
    """JinEnv.py:504-574 with numeric mc, mp, l; weights [wx, wq, wdx, wdq] learnable; goal [0, pi, 0, 0]; g = 10."""
    x, q, dx, dq = _syms('x q dx dq')
    u, = _syms('u')
    beta, wx, wq, wdx, wdq = _syms('beta wx wq wdx wdq')
    g = 10
    ddx = (u + mp * sp.sin(q) * (l * dq * dq + g * sp.cos(q))) / (mc + mp * sp.sin(q) * sp.sin(q))
    ddq = (-u * sp.cos(q) - mp * l * dq * dq * sp.sin(q) * sp.cos(q) - (mc + mp) * g * sp.sin(q)) / \
        (l * mc + l * mp * sp.sin(q) * sp.sin(q))
    f = sp.Matrix([dx, dq, ddx, ddq])
    h = wx * (x - 0.0) ** 2 + wq * (q - math.pi) ** 2 + wdx * (dx - 0.0) ** 2 + wdq * (dq - 0.0) ** 2
    c = h + wu * (u * u)
    return OracleModel('cartpole', [x, q, dx, dq], [u], [beta, wx, wq, wdx, wdq], beta * f, beta * c, h, sel=[0, 1])


def to_quaternion(angle, axis):
    """/root/reference/JinEnv/JinEnv.py:1730-1737"""
    d = np.asarray(axis, dtype=float)
    d = d / np.linalg.norm(d)
    return [math.cos(angle / 2)] + (math.sin(angle / 2) * d).tolist()
