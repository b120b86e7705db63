"""TEST INFRASTRUCTURE — CPU oracle for the CPDP gradient iteration.  Not part of the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.

A numpy/scipy restatement of the reference algorithm, following it step for step:

  * ``solve``        ← ``COCSys.cocSolver``      /root/reference/CPDP/CPDP.py:92-198
                       RK4 multiple shooting (:111-124), NLP layout/seed/bounds (:127-175), unpacking and
                       costate convention (:187-193).  IPOPT (un-vendored, README pins coinor-libipopt 3.11.9 /
                       the casadi wheel's bundled build) is replaced by its published algorithm restricted to
                       this problem class (no inequalities => no barrier): full-space Newton on the KKT system
                       with exact Lagrangian Hessian, IPOPT's inertia-correction schedule (Waechter & Biegler
                       2006, Alg. IC: 1e-4, x100 first time, x8 afterwards, /3 on re-entry) and IPOPT's filter
                       line search (same paper, Sec. 2.3, default constants; no second-order correction, no
                       restoration phase).  Pinned by the reference's stored run (tests/golden, KAT K2/K3).
  * ``aux_asshipped``← ``COCSys.auxSysSolver``   /root/reference/CPDP/CPDP.py:301-381, RHS formulas :253-298,
                       derivative set :201-248.  Calls the installed ``scipy.integrate.solve_ivp`` exactly as the
                       reference does (BDF backward / default RK45 forward, default tolerances, one call per grid
                       interval, ``t_eval=[end]``) and ``scipy.interpolate.interp1d`` for every interpolant.
  * ``loss_grad``    ← the ``getloss_*corrections`` closures, e.g. /root/reference/lib/QuadAlgorithm.py:616-639,
                       /root/reference/Examples/rocket_groundtruth.py:45-70 (note: dl_dy = y - wp, no factor 2).

Parity status: PINNED against the reference's only stored outputs
(/root/reference/data/uav_results_random_20210308113016.mat → tests/golden/quad_run.npz, KATs K1-K5 of SURVEY.md
§8c).  The reference itself cannot run here (no casadi / IPOPT in the image).
"""
import numpy as np
import sympy as sp
import scipy.linalg as sla
from scipy.integrate import solve_ivp
from scipy.interpolate import interp1d


# ---------------------------------------------------------------------------------------------
# lambdified model functions
# ---------------------------------------------------------------------------------------------
class _Fns:
    """Numeric callables derived symbolically from an OracleModel (CPDP.py:201-248 derivative set)."""

    def __init__(self, model):
        self.model = model
        n, m, r = model.n, model.m, model.r
        x, u, th = model.x, model.u, model.theta
        pd = list(getattr(model, 'pdata', [])) + ([model.time] if getattr(model, 'time', None) is not None else [])
        z = x + u
        mu = [sp.Symbol('mu%d' % i, real=True) for i in range(n)]
        wc = sp.Symbol('wc', real=True)
        f, c, h = model.dyn, model.path, model.final
        args = [x, u, th, pd]

        def lam(a, outs):
            flat = sp.Matrix([e for o in outs for e in (list(o) if isinstance(o, sp.MatrixBase) else [o])])
            shapes = [(o.shape if isinstance(o, sp.MatrixBase) else ()) for o in outs]
            fn = sp.lambdify(a, list(flat), modules='math', cse=True)

            def call(*vals):
                v = np.array(fn(*vals), dtype=float)
                res, k = [], 0
                for s in shapes:
                    cnt = int(np.prod(s)) if s else 1
                    res.append(v[k:k + cnt].reshape(s) if s else v[k])
                    k += cnt
                return res
            return call

        fvec = sp.Matrix(1, n, list(f))   # row vector => 1-D after squeeze below
        _fc = lam(args, [fvec, c])
        self.fc = lambda *a: (lambda o: (o[0].ravel(), o[1]))(_fc(*a))
        Hm = wc * c + (sp.Matrix(mu).T * f)[0, 0]
        fz = f.jacobian(z)
        cz = sp.Matrix([c]).jacobian(z)
        Hz = sp.Matrix([Hm]).jacobian(z)
        Hzz = Hz.jacobian(z)
        _s1 = lam(args, [fvec, c, fz, cz])
        self.stage1 = lambda *a: (lambda o: (o[0].ravel(), o[1], o[2], o[3]))(_s1(*a))
        self.stage2 = lam(args + [mu, wc], [Hzz])
        # PMP set with H = c + f' lam
        lm = mu
        Hp = c + (sp.Matrix(lm).T * f)[0, 0]
        Hx = sp.Matrix([Hp]).jacobian(x)
        Hu = sp.Matrix([Hp]).jacobian(u)
        self.pmp = lam([x, u, lm, th, pd], [f.jacobian(x), f.jacobian(u), f.jacobian(th),
                                        Hx.jacobian(x), Hx.jacobian(u), Hx.jacobian(th),
                                        Hu.jacobian(u), Hu.jacobian(th)])
        hx = sp.Matrix([h]).jacobian(x)
        self.term = lam([x, th, pd], [h, hx, hx.jacobian(x), hx.jacobian(th)])


class OracleIntegrationError(RuntimeError):
    pass


class Oracle:
    def __init__(self, model, n_grid=10, steps_per_grid=4):
        self.model = model
        self.n, self.m, self.r = model.n, model.m, model.r
        self.N, self.S = int(n_grid), int(steps_per_grid)
        self.fn = _Fns(model)
        self.pd = np.zeros(len(getattr(model, 'pdata', [])))   # per-problem constants (e.g. goal position)
        self.tv = getattr(model, 'time', None) is not None      # COCSys_TimeVarying (CPDP.py:394-787): explicit time in f, c, h

    def _pd(self, t):
        """model constants for an evaluation at time t: [pdata | t] for a time-varying model"""
        return np.concatenate((self.pd, [float(t)])) if self.tv else self.pd

    # -----------------------------------------------------------------------------------------
    # RK4 interval map (CPDP.py:111-124)
    # -----------------------------------------------------------------------------------------
    def interval(self, x, u, th, DT, tk=0.0):
        """(x,u) -> (x_end, integral of path cost) over one grid interval: S classical RK4 steps.
        Time-varying models: every stage of every sub-step is evaluated at the interval's start time tk (CPDP.py:512-519)."""
        X = np.array(x, dtype=float)
        Q = 0.0
        fc = self.fn.fc
        pd = self._pd(tk)
        for _ in range(self.S):
            k1, q1 = fc(X, u, th, pd)
            k2, q2 = fc(X + DT / 2 * k1, u, th, pd)
            k3, q3 = fc(X + DT / 2 * k2, u, th, pd)
            k4, q4 = fc(X + DT * k3, u, th, pd)
            X = X + DT / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
            Q = Q + DT / 6 * (q1 + 2 * q2 + 2 * q3 + q4)
        return X, Q

    def interval_derivs(self, x, u, th, lam_next, DT, tk=0.0):
        """Value, first derivatives and Hessian of q + lam_next' F for one interval.

        First order by forward sensitivities through the 4*S stages, second order by the discrete adjoint:
        Hess = sum_s Sz_s' * Hess_z(mu_s' f + w_s c)(z_s) * Sz_s  with mu_s, w_s the adjoints of stage s."""
        n, m = self.n, self.m
        nz = n + m
        pd = self._pd(tk)
        E = np.zeros((m, nz)); E[:, n:] = np.eye(m)
        X = np.array(x, dtype=float)
        Sx = np.zeros((n, nz)); Sx[:, :n] = np.eye(n)
        Q = 0.0
        dQ = np.zeros(nz)
        rec = []  # per substep: list of (xs, fz, cz, Sz) for the 4 stages
        bco = [1.0 / 6, 2.0 / 6, 2.0 / 6, 1.0 / 6]
        aco = [0.0, 0.5, 0.5, 1.0]
        for _ in range(self.S):
            stages = []
            kprev = None; dkprev = None
            Xn = X.copy(); Sn = Sx.copy()
            for s in range(4):
                xs = X if s == 0 else X + aco[s] * DT * kprev
                Ss = Sx if s == 0 else Sx + aco[s] * DT * dkprev
                f, c, fz, cz = self.fn.stage1(xs, u, th, pd)
                Sz = np.vstack([Ss, E])
                dk = fz @ Sz
                stages.append((xs, fz, cz.ravel(), Sz))
                Xn = Xn + bco[s] * DT * f
                Sn = Sn + bco[s] * DT * dk
                Q += bco[s] * DT * c
                dQ += bco[s] * DT * (cz.ravel() @ Sz)
                kprev, dkprev = f, dk
            rec.append(stages)
            X, Sx = Xn, Sn
        A, B = Sx[:, :n], Sx[:, n:]
        grad = dQ + lam_next @ Sx          # d(q + lam'F)/dz
        # adjoint sweep
        H = np.zeros((nz, nz))
        a = np.array(lam_next, dtype=float)
        for stages in reversed(rec):
            ax = np.zeros(n)
            kap_next_xi = None
            for s in (3, 2, 1, 0):
                kap = bco[s] * DT * a
                if s < 3:
                    kap = kap + aco[s + 1] * DT * kap_next_xi
                w = bco[s] * DT
                xs, fz, cz, Sz = stages[s]
                Hzz, = self.fn.stage2(xs, u, th, pd, kap, w)
                H += Sz.T @ Hzz @ Sz
                xi = fz[:, :n].T @ kap + w * cz[:n]
                ax += xi
                kap_next_xi = xi
            a = a + ax
        return X, Q, A, B, dQ, grad, 0.5 * (H + H.T)

    # -----------------------------------------------------------------------------------------
    # forward solve (CPDP.py:92-198)
    # -----------------------------------------------------------------------------------------
    def _riccati_step(self, Ak, Bk, Hk, gLk, dfc, lam, hx, hxx, delta):
        """Newton step of the equality-constrained NLP through the stage structure of its KKT matrix (block LDL' = Riccati
        recursion): the linear algebra a banded / structure-exploiting solver (MUMPS inside IPOPT) performs, in O(N n^3)
        instead of the dense O((N n)^3) factorisation of `solve`'s default path.  Same step as the dense path (checked in
        tests/test_oracle_kat.py); the inertia is correct iff every Quu block is positive definite.
        Returns (ok, dX [N+1,n], dU [N,m], lam_new [N+1,n])."""
        n, m, N = self.n, self.m, self.N
        V = hxx + delta * np.eye(n)
        v = hx.copy()
        Vs, vs, Ks, ks = [None] * (N + 1), [None] * (N + 1), [None] * N, [None] * N
        Vs[N], vs[N] = V, v
        for k in range(N - 1, -1, -1):
            AB = np.hstack([Ak[k], Bk[k]])
            vt = v + V @ dfc[k + 1]
            gq = gLk[k] - AB.T @ lam[k + 1]
            Q = 0.5 * (Hk[k] + Hk[k].T) + delta * np.eye(n + m) + AB.T @ (V @ AB)
            qv = gq + AB.T @ vt
            Quu = 0.5 * (Q[n:, n:] + Q[n:, n:].T)
            try:
                L = np.linalg.cholesky(Quu)
            except np.linalg.LinAlgError:
                return False, None, None, None
            Qux = 0.5 * (Q[n:, :n] + Q[:n, n:].T)
            sol = -sla.cho_solve((L, True), np.hstack([Qux, qv[n:, None]]))
            K, kf = sol[:, :n], sol[:, n]
            Qxx = 0.5 * (Q[:n, :n] + Q[:n, :n].T)
            V = Qxx + 0.5 * (Qux.T @ K + K.T @ Qux)
            v = qv[:n] + Qux.T @ kf
            Vs[k], vs[k], Ks[k], ks[k] = V, v, K, kf
        dX = np.zeros((N + 1, n)); dU = np.zeros((N, m)); ln = np.zeros((N + 1, n))
        dX[0] = dfc[0]
        for k in range(N):
            dU[k] = ks[k] + Ks[k] @ dX[k]
            ln[k] = vs[k] + Vs[k] @ dX[k]
            dX[k + 1] = dfc[k + 1] + Ak[k] @ dX[k] + Bk[k] @ dU[k]
        ln[N] = vs[N] + Vs[N] @ dX[N]
        return True, dX, dU, ln

    def solve(self, x0, horizon, theta, tol=1e-10, max_iter=200, verbose=False, return_info=False, linear_solver='dense'):
        """linear_solver: 'dense' (default: assembled KKT matrix, LDL' inertia count, dense solve -- the independent check of the
        kernels' structured recursion) or 'riccati' (stage-structured recursion, used by bench.py's CPU baseline so that the port
        is not charged for a dense factorisation no sparse NLP solver would perform)."""
        n, m, N = self.n, self.m, self.N
        nz = n + m
        th = np.asarray(theta, dtype=float)
        x0 = np.asarray(x0, dtype=float).ravel()
        DT = horizon / N / self.S
        # node times: [T/N*k] (CPDP.py:192); numpy.linspace for the time-varying class (CPDP.py:544)
        tgrid = np.linspace(0, horizon, N + 1) if self.tv else np.array([horizon / N * k for k in range(N + 1)])
        nw = (N + 1) * n + N * m
        ng = (N + 1) * n
        w = np.zeros(nw)          # seed: all zeros (CPDP.py:139,155,167)
        lam = np.zeros(ng)
        ox = lambda k: k * nz
        ou = lambda k: k * nz + n

        def evaluate(wv):
            """objective and constraint vector only (line search)."""
            J = 0.0
            g = np.zeros(ng)
            g[:n] = x0 - wv[:n]
            for k in range(N):
                xe, q = self.interval(wv[ox(k):ox(k) + n], wv[ou(k):ou(k) + m], th, DT, tgrid[k])
                J += q
                g[(k + 1) * n:(k + 2) * n] = xe - wv[ox(k + 1):ox(k + 1) + n]
            J += self.fn.term(wv[ox(N):ox(N) + n], th, self._pd(tgrid[-1]))[0]
            return J, g

        filt = []                 # IPOPT filter: list of (theta, phi) corners
        theta_min = theta_max = None
        delta_last = 0.0
        info = dict(iters=0, status='max_iter', reg=[], alphas=[])
        for it in range(max_iter + 1):
            structured = (linear_solver == 'riccati')
            gradJ = np.zeros(nw)
            W = None if structured else np.zeros((nw, nw))
            Ag = None if structured else np.zeros((ng, nw))
            g = np.zeros(ng)
            J = 0.0
            g[:n] = x0 - w[:n]
            if not structured:
                Ag[:n, :n] = -np.eye(n)
            Aks, Bks, Hks, gLs = [], [], [], []
            gradL = np.zeros(nw)                       # gradient of the Lagrangian, assembled stage-wise
            gradL[:n] -= lam[:n]
            for k in range(N):
                xk, uk = w[ox(k):ox(k) + n], w[ou(k):ou(k) + m]
                lk1 = lam[(k + 1) * n:(k + 2) * n]
                xe, q, A, B, dQ, gLk, H = self.interval_derivs(xk, uk, th, lk1, DT, tgrid[k])
                J += q
                g[(k + 1) * n:(k + 2) * n] = xe - w[ox(k + 1):ox(k + 1) + n]
                gradJ[ox(k):ox(k) + nz] += dQ
                gradL[ox(k):ox(k) + nz] += gLk
                gradL[ox(k + 1):ox(k + 1) + n] -= lk1
                if structured:
                    Aks.append(A); Bks.append(B); Hks.append(H); gLs.append(gLk)
                else:
                    W[ox(k):ox(k) + nz, ox(k):ox(k) + nz] += H
                    Ag[(k + 1) * n:(k + 2) * n, ox(k):ox(k) + n] = A
                    Ag[(k + 1) * n:(k + 2) * n, ou(k):ou(k) + m] = B
                    Ag[(k + 1) * n:(k + 2) * n, ox(k + 1):ox(k + 1) + n] = -np.eye(n)
            hv, hx, hxx, _ = self.fn.term(w[ox(N):ox(N) + n], th, self._pd(tgrid[-1]))
            J += hv
            gradJ[ox(N):ox(N) + n] += hx.ravel()
            gradL[ox(N):ox(N) + n] += hx.ravel()
            if not structured:
                W[ox(N):ox(N) + n, ox(N):ox(N) + n] += hxx
            kkt_err = max(np.abs(gradL).max() if structured else np.abs(gradJ + Ag.T @ lam).max(), np.abs(g).max())
            if verbose:
                print('it %3d  J=%.10f  |gradL|=%.3e |g|=%.3e' % (it, J, np.abs(gradJ + Ag.T @ lam).max(), np.abs(g).max()))
            if kkt_err < tol:
                info['status'] = 'converged'
                break
            if it == max_iter:
                break
            # ---- inertia-corrected Newton step
            delta = 0.0
            first = True
            while True:
                if structured:
                    ok_, dXs, dUs, lns = self._riccati_step(Aks, Bks, Hks, gLs, g.reshape(N + 1, n), lam.reshape(N + 1, n),
                                                            hx.ravel(), hxx, delta)
                    if ok_:
                        break
                else:
                    K = np.zeros((nw + ng, nw + ng))
                    K[:nw, :nw] = W + delta * np.eye(nw)
                    K[:nw, nw:] = Ag.T
                    K[nw:, :nw] = Ag
                    if _inertia_ok(K, nw, ng):
                        break
                if first:
                    delta = 1e-4 if delta_last == 0.0 else max(1e-20, delta_last / 3.0)
                    first = False
                else:
                    delta = delta * (100.0 if delta_last == 0.0 else 8.0)
                if delta > 1e40:
                    raise FloatingPointError('inertia correction failed')
            if delta > 0:
                delta_last = delta
            info['reg'].append(delta)
            if structured:
                d = np.concatenate([np.hstack([dXs[:N], dUs]).ravel(), dXs[N]])
                lam_new = lns.ravel()
            else:
                sol = np.linalg.solve(K, -np.concatenate([gradJ, g]))
                d, lam_new = sol[:nw], sol[nw:]
            # ---- IPOPT filter line search (Waechter & Biegler 2006, Sec. 2.3; no barrier term: phi = J,
            #      theta = |g|_1; constants = IPOPT defaults gamma_theta 1e-5, gamma_phi 1e-8, delta 1, s_theta 1.1,
            #      s_phi 2.3, eta_phi 1e-8, theta_min/max = 1e-4/1e4 * max(1, theta(x0)); no second-order correction
            #      and no restoration phase: more than 30 halvings -> 'linesearch_fail')
            th0 = np.abs(g).sum()
            if theta_min is None:
                theta_min = 1e-4 * max(1.0, th0)
                theta_max = 1e4 * max(1.0, th0)
            gd = gradJ @ d
            alpha = 1.0
            accepted = ftype = False
            for _ls in range(31):
                try:
                    with np.errstate(all='ignore'):
                        Jt, gt = evaluate(w + alpha * d)
                    tht = np.abs(gt).sum()
                except (ValueError, OverflowError, FloatingPointError):   # e.g. math.cos(inf) at a wild trial point
                    Jt = tht = np.inf
                ok = np.isfinite(Jt) and np.isfinite(tht) and tht <= theta_max
                if ok:
                    for (tf, pf) in filt:
                        if tht >= tf and Jt >= pf:
                            ok = False
                            break
                if ok:
                    switching = gd < 0 and alpha * (-gd) ** 2.3 > th0 ** 1.1
                    if switching and th0 <= theta_min:
                        if Jt <= J + 1e-8 * alpha * gd:
                            accepted = ftype = True
                    elif tht <= (1 - 1e-5) * th0 or Jt <= J - 1e-8 * th0:
                        accepted = True
                if accepted:
                    break
                alpha *= 0.5
            if not accepted:
                info['status'] = 'linesearch_fail'
                break
            if not ftype:
                filt.append(((1 - 1e-5) * th0, J - 1e-8 * th0))
            info['alphas'].append(alpha)
            w = w + alpha * d
            lam = lam + alpha * (lam_new - lam)
            info['iters'] = it + 1
        info['J'] = J
        info['kkt'] = kkt_err
        # ---- unpack (CPDP.py:187-193)
        sc = np.concatenate([w, np.zeros(m)]).reshape(-1, nz)
        X = sc[:, :n].copy()
        U = sc[:, n:].copy()
        U[-1, :] = U[-2, :]
        time_grid = tgrid
        Lam = lam.reshape(-1, n).copy()
        if return_info:
            return time_grid, X, U, Lam, info
        return time_grid, X, U, Lam

    # -----------------------------------------------------------------------------------------
    # auxiliary system (CPDP.py:253-381)
    # -----------------------------------------------------------------------------------------
    def _coeffs(self, x, u, lam, th, t=0.0):
        fx, fu, fe, Hxx, Hxu, Hxe, Huu, Hue = self.fn.pmp(x, u, lam, th, self._pd(t))
        invHuu = np.linalg.inv(Huu)
        return fx, fu, fe, Hxx, Hxu, Hxe, Huu, Hue, invHuu

    def riccati_rhs(self, x, u, lam, th, P, W, reassoc=False, t=0.0):
        """CPDP.py:262-274 (time-varying: :650-667)"""
        fx, fu, fe, Hxx, Hxu, Hxe, Huu, Hue, invHuu = self._coeffs(x, u, lam, th, t)
        G = fu @ invHuu
        HxuinvHuu = Hxu @ invHuu
        A = fx - G @ Hxu.T
        R = G @ fu.T
        Q = Hxx - HxuinvHuu @ Hxu.T
        r = fe - G @ Hue
        q = Hxe - HxuinvHuu @ Hue
        if reassoc:      # same mathematics, different floating-point association (see ``aux``)
            P_dot = -(Q + (A.T @ P + P @ A) - P @ (R @ P))
            W_dot = P @ (R @ W) - (A.T @ W + P @ r) - q
            return P_dot, W_dot
        P_dot = -(Q + A.T @ P + P @ A - P @ R @ P)
        W_dot = P @ R @ W - A.T @ W - P @ r - q
        return P_dot, W_dot

    def riccati_jac(self, x, u, lam, th, P, W, t=0.0):
        """Closed-form Jacobian of ``riccati_rhs`` w.r.t. the row-major state [vec P | vec W] (the matrix scipy's BDF
        approximates by finite differences when the reference calls it without ``jac``, CPDP.py:335):
        d(Pdot)[dP] = -((A'-PR) dP + dP (A-RP)),  d(Wdot)[dP, dW] = dP (RW - r_) + (PR - A') dW."""
        fx, fu, fe, Hxx, Hxu, Hxe, Huu, Hue, invHuu = self._coeffs(x, u, lam, th, t)
        n, r = P.shape[0], W.shape[1]
        G = fu @ invHuu
        A = fx - G @ Hxu.T
        R = G @ fu.T
        r_ = fe - G @ Hue
        In = np.eye(n)
        J = np.zeros((n * n + n * r, n * n + n * r))
        J[:n * n, :n * n] = -(np.kron(A.T - P @ R, In) + np.kron(In, (A - R @ P).T))
        J[n * n:, :n * n] = np.kron(In, (R @ W - r_).T)
        J[n * n:, n * n:] = np.kron(P @ R - A.T, np.eye(r))
        return J

    def aux_controller(self, x, u, lam, th, P, W, Xa, t=0.0):
        """CPDP.py:294-295"""
        fx, fu, fe, Hxx, Hxu, Hxe, Huu, Hue, invHuu = self._coeffs(x, u, lam, th, t)
        return -invHuu @ ((Hxu.T + fu.T @ P) @ Xa + fu.T @ W + Hue)

    def aux_rhs(self, x, u, lam, th, P, W, Xa, t=0.0):
        """CPDP.py:295-297"""
        fx, fu, fe, Hxx, Hxu, Hxe, Huu, Hue, invHuu = self._coeffs(x, u, lam, th, t)
        Ua = -invHuu @ ((Hxu.T + fu.T @ P) @ Xa + fu.T @ W + Hue)
        return fx @ Xa + fu @ Ua + fe

    def aux(self, time_grid, X, U, Lam, theta, back=None, fwd=None, return_counts=False):
        """CPDP.py:301-381.  ``back`` / ``fwd`` are the keyword dictionaries handed to solve_ivp for the
        backward Riccati sweep and the forward sweep; the as-shipped reference is back={'method':'BDF'}, fwd={}.
        Two oracle-only keys of ``back``: ``jac='closed'`` hands scipy's BDF the closed-form Jacobian instead of its
        finite-difference one (same solver, same control logic); ``reassoc=True`` evaluates the Riccati right-hand
        side with a different association of the same matrix products (measures how far roundoff alone moves the
        as-shipped result through the finite-difference Jacobian)."""
        # as shipped: BDF for COCSys (CPDP.py:335), solve_ivp's default RK45 for COCSys_TimeVarying (CPDP.py:740)
        back = dict(({} if self.tv else {'method': 'BDF'}) if back is None else back)
        reassoc = back.pop('reassoc', False)
        fwd = {} if fwd is None else fwd
        n, m, r, N = self.n, self.m, self.r, self.N
        th = np.asarray(theta, dtype=float)
        opt_sol = interp1d(time_grid, np.concatenate((X, U, Lam), axis=1), axis=0)
        counts = dict(back_rhs=0, fwd_rhs=0)

        def split(t):
            v = opt_sol(t)
            return v[:n], v[n:n + m], v[n + m:]

        def vec_PW_ode(t, vec_PW):
            counts['back_rhs'] += 1
            P = vec_PW[:n * n].reshape(n, n)
            W = vec_PW[n * n:].reshape(n, -1)
            x, u, lam = split(t)
            Pd, Wd = self.riccati_rhs(x, u, lam, th, P, W, reassoc=reassoc, t=t)
            return np.concatenate((Pd.flatten(), Wd.flatten()))

        if back.get('jac') == 'closed':
            def vec_PW_jac(t, vec_PW):
                x, u, lam = split(t)
                return self.riccati_jac(x, u, lam, th, vec_PW[:n * n].reshape(n, n), vec_PW[n * n:].reshape(n, -1), t=t)
            back['jac'] = vec_PW_jac

        xT = opt_sol(float(time_grid[-1]))[:n]
        _, _, hxx, hxe = self.fn.term(xT, th, self._pd(time_grid[-1]))
        PW = np.zeros((N + 1, n * n + n * r))
        PW[-1, :] = np.concatenate((hxx.flatten(), hxe.flatten()))
        for k in range(N, 0, -1):
            t_span = [time_grid[k], time_grid[k - 1]]
            sol = solve_ivp(vec_PW_ode, t_span, PW[k, :], t_eval=[t_span[1]], **back)
            if sol.status != 0:       # the reference would crash on the empty sol.y here (CPDP.py:336)
                raise OracleIntegrationError('backward sweep failed on interval %d: %s' % (k, sol.message))
            PW[k - 1, :] = sol.y.flatten()
        PW_sol = interp1d(time_grid, PW, axis=0)

        def PWat(t):
            v = PW_sol(t)
            return v[:n * n].reshape(n, n), v[n * n:].reshape(n, -1)

        def vec_aux_ode(t, vecX):
            counts['fwd_rhs'] += 1
            Xa = vecX.reshape(n, r)
            x, u, lam = split(t)
            P, W = PWat(t)
            return self.aux_rhs(x, u, lam, th, P, W, Xa, t=t).flatten()

        Xa = np.zeros((N + 1, n * r))
        Ua = np.zeros((N + 1, m * r))
        x, u, lam = split(0)
        P, W = PWat(0)
        Ua[0, :] = self.aux_controller(x, u, lam, th, P, W, Xa[0].reshape(n, r), t=0.0).flatten()
        for k in range(N):
            t_span = [time_grid[k], time_grid[k + 1]]
            sol = solve_ivp(vec_aux_ode, t_span, Xa[k, :], t_eval=[time_grid[k + 1]], **fwd)
            if sol.status != 0:
                raise OracleIntegrationError('forward sweep failed on interval %d: %s' % (k, sol.message))
            Xa[k + 1, :] = sol.y.flatten()
            x, u, lam = split(float(time_grid[k + 1]))
            P, W = PWat(float(time_grid[k + 1]))
            Ua[k + 1, :] = self.aux_controller(x, u, lam, th, P, W, Xa[k + 1].reshape(n, r), t=time_grid[k + 1]).flatten()
        if return_counts:
            return Xa, Ua, PW, counts
        return Xa, Ua, PW

    # -----------------------------------------------------------------------------------------
    # loss closure (QuadAlgorithm.py:616-639 et al.)
    # -----------------------------------------------------------------------------------------
    def loss_grad(self, taus, waypoints, time_grid, X, Xa, sel=None):
        n, r = self.n, self.r
        sel = self.model.sel if sel is None else sel
        opt_x = interp1d(time_grid, X, axis=0)
        aux_x = interp1d(time_grid, Xa, axis=0)
        loss = 0.0
        dl = np.zeros(r)
        waypoints = np.atleast_2d(waypoints)
        for k, t in enumerate(np.atleast_1d(taus)):
            y = opt_x(t)[sel]
            loss += np.linalg.norm(waypoints[k, :] - y) ** 2
            dl_dy = y - waypoints[k, :]
            dx_dp = aux_x(t).reshape(n, r)
            dl += dl_dy @ dx_dp[sel, :]
        return loss, dl

    # -----------------------------------------------------------------------------------------
    def grad_iter(self, x0, horizon, theta, taus, waypoints, back=None, fwd=None, tol=1e-10, linear_solver='dense'):
        """One CPDP gradient iteration for one OCP: (loss, dL/dtheta, extras)."""
        tg, X, U, Lam, info = self.solve(x0, horizon, theta, tol=tol, return_info=True, linear_solver=linear_solver)
        Xa, Ua, PW = self.aux(tg, X, U, Lam, theta, back=back, fwd=fwd)
        loss, dl = self.loss_grad(taus, waypoints, tg, X, Xa)
        return loss, dl, dict(time_grid=tg, X=X, U=U, Lam=Lam, Xa=Xa, Ua=Ua, PW=PW, info=info)


TIGHT = dict(rtol=1e-10, atol=1e-12)


def _inertia_ok(K, nw, ng):
    """True iff K has exactly nw positive, ng negative and no zero eigenvalues (via LDL')."""
    try:
        _, D, _ = sla.ldl(K, lower=True, hermitian=True, overwrite_a=False, check_finite=False)
    except Exception:
        return False
    npos = nneg = 0
    i, nn = 0, K.shape[0]
    while i < nn:
        if i + 1 < nn and D[i + 1, i] != 0.0:
            a, b, c = D[i, i], D[i + 1, i], D[i + 1, i + 1]
            tr, det = a + c, a * c - b * b
            if det < 0:
                npos += 1; nneg += 1
            elif det > 0:
                if tr > 0: npos += 2
                else: nneg += 2
            else:
                return False
            i += 2
        else:
            if D[i, i] > 0: npos += 1
            elif D[i, i] < 0: nneg += 1
            else: return False
            i += 1
    return npos == nw and nneg == ng
