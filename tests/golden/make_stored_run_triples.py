"""The 100 (theta, loss, dL/dtheta) triples the reference's stored learning run pins (KAT K5), and the oracle's
values at every one of them.

The stored run (/root/reference/data/uav_results_random_20210308113016.mat -> tests/golden/quad_run.npz) is 100 Nesterov
iterations (/root/reference/lib/QuadAlgorithm.py:469-494, true_loss_print_flag off): the loss / gradient of iteration j were
evaluated at the look-ahead point  theta_j + mu v_j  (v_0 = 0, v_{j+1} = theta_{j+1} - theta_j), so
    g_j = (mu v_j - v_{j+1}) / lr     and     loss_trace[j]
are reference outputs at known inputs -- non-accumulating, unlike a replay of the recurrence.
For each point this script runs the CPU oracle once (forward solve) and the auxiliary system twice:
    cj : scipy's BDF handed the closed-form Jacobian (what k_riccati_bdf implements)
    fd : scipy's BDF with its own finite-difference Jacobian (the as-shipped call, CPDP.py:335)
Run from the repo root:  python tests/golden/make_stored_run_triples.py  ->  tests/golden/stored_run_triples.npz
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import models  # noqa: E402
from oracle.cpdp_oracle import Oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def triples(g):
    P, lr, mu = g["parameter_trace"], float(g["learning_rate"]), float(g["mu"])
    v = np.vstack([np.zeros((1, P.shape[1])), np.diff(P, axis=0)])          # v[j] = velocity entering iteration j
    look = P[:-1] + mu * v[:-1]
    grads = (mu * v[:-1] - v[1:]) / lr
    return look, g["loss_trace"].ravel(), grads


_ORC = None


def one(job):
    global _ORC
    g = np.load(os.path.join(HERE, "quad_run.npz"))
    if _ORC is None:
        _ORC = Oracle(models.quadrotor(), n_grid=int(g["n_grid"]), steps_per_grid=int(g["steps_per_grid"]))
        _ORC.pd = g["goal_position"].astype(float)
    theta = job
    tg, X, U, Lam, info = _ORC.solve(g["ini_state"], 1.0, theta, return_info=True)
    out = [info["iters"]]
    for back in ({"method": "BDF", "jac": "closed"}, {"method": "BDF"}):
        Xa, Ua, PW = _ORC.aux(tg, X, U, Lam, theta, back=back, fwd={})
        loss, dl = _ORC.loss_grad(g["time_grid"], g["waypoints"], tg, X, Xa)
        out += [loss, dl]
    return out


def main():
    g = np.load(os.path.join(HERE, "quad_run.npz"))
    look, losses, grads = triples(g)
    with Pool(min(8, os.cpu_count())) as pool:
        res = pool.map(one, list(look), chunksize=1)
    iters = np.array([r[0] for r in res])
    loss_cj = np.array([r[1] for r in res]); dl_cj = np.stack([r[2] for r in res])
    loss_fd = np.array([r[3] for r in res]); dl_fd = np.stack([r[4] for r in res])
    rel = lambda a, b: np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)
    print("max rel err of dL/dtheta vs the stored run: cj %.3e   fd %.3e" % (rel(dl_cj, grads).max(), rel(dl_fd, grads).max()))
    print("fd points above 1e-5:", [(int(j), float(e)) for j, e in enumerate(rel(dl_fd, grads)) if e > 1e-5])
    np.savez_compressed(os.path.join(HERE, "stored_run_triples.npz"), theta=look, loss_ref=losses, grad_ref=grads, iters=iters,
                        loss_cj=loss_cj, dl_cj=dl_cj, loss_fd=loss_fd, dl_fd=dl_fd)


if __name__ == "__main__":
    main()
