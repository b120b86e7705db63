"""Work counters of scipy's BDF on the backward Riccati sweep of every oracle fixture problem: accepted steps, right-hand
side evaluations, LU factorisations and Jacobian evaluations, summed over the grid intervals exactly as
COCSys.auxSysSolver restarts the solver on each of them (/root/reference/CPDP/CPDP.py:333-336).  scipy's own `BDF` class is
stepped (solve_ivp is the same loop), with the closed-form Jacobian the CUDA kernel uses (`asshipped_cj`).
Run from the repo root:  python tests/golden/make_counter_fixture.py   ->  tests/golden/oracle_counters.npz
"""
import os
import sys
from multiprocessing import Pool

import numpy as np
from scipy.integrate import BDF
from scipy.interpolate import interp1d

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import models  # noqa: E402
from oracle.cpdp_oracle import Oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"pendulum": "pendulum", "robotarm": "robotarm", "rocket": "rocket", "quadkat": "quadrotor", "quad50": "quadrotor"}


def count_problem(args):
    kind, n_grid, T, theta, pd, X, U, Lam = args
    orc = Oracle(getattr(models, kind)(), n_grid=n_grid)
    if pd is not None:
        orc.pd = np.asarray(pd, dtype=float)
    n, m, r = orc.n, orc.m, orc.r
    tg = np.array([T / n_grid * k for k in range(n_grid + 1)])
    osol = interp1d(tg, np.concatenate((X, U, Lam), axis=1), axis=0)
    th = np.asarray(theta, dtype=float)

    def split(t):
        v = osol(t)
        return v[:n], v[n:n + m], v[n + m:]

    def rhs(t, y):
        x, u, lam = split(t)
        Pd, Wd = orc.riccati_rhs(x, u, lam, th, y[:n * n].reshape(n, n), y[n * n:].reshape(n, -1))
        return np.concatenate((Pd.ravel(), Wd.ravel()))

    def jac(t, y):
        x, u, lam = split(t)
        return orc.riccati_jac(x, u, lam, th, y[:n * n].reshape(n, n), y[n * n:].reshape(n, -1))

    _, _, hxx, hxe = orc.fn.term(osol(float(tg[-1]))[:n], th, orc.pd)
    y = np.concatenate((hxx.flatten(), hxe.flatten()))
    steps = nfev = nlu = njev = 0
    for k in range(n_grid, 0, -1):
        s = BDF(rhs, tg[k], y, tg[k - 1], jac=jac)
        while s.status == "running":
            s.step()
            steps += 1
        if s.status != "finished":
            return np.array([-1, -1, -1, -1])
        nfev += s.nfev; nlu += s.nlu; njev += s.njev
        y = s.y
    return np.array([nfev, steps, nlu, njev])


def main():
    jobs, index = [], []
    for case, kind in CASES.items():
        f = np.load(os.path.join(HERE, "oracle_%s.npz" % case))
        B = f["x0"].shape[0]
        for b in range(B):
            pd = f["pdata"][b] if "pdata" in f.files else None
            jobs.append((kind, int(f["n_grid"]), float(f["T"]), f["theta"][b], pd, f["X"][b], f["U"][b], f["Lam"][b]))
            index.append(case)
    with Pool(min(8, os.cpu_count())) as pool:
        res = pool.map(count_problem, jobs, chunksize=1)
    out = {}
    for case in CASES:
        out[case] = np.stack([r for c, r in zip(index, res) if c == case])
        print(case, out[case].tolist())
    np.savez_compressed(os.path.join(HERE, "oracle_counters.npz"), **out)


if __name__ == "__main__":
    main()
