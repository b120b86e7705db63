"""Host-side learner (lfsd_b200.optim) against the reference's stored run and against independent restatements of
the update rules (/root/reference/lib/QuadAlgorithm.py:239-257, 454-578)."""
import os

import numpy as np
import pytest

import lfsd_b200  # noqa: F401
from lfsd_b200.optim import Learner

HERE = os.path.dirname(os.path.abspath(__file__))


def test_nesterov_recurrence_reproduces_stored_parameter_trace():
    """Feeding the gradients implied by the stored parameter_trace (KAT K5 algebra: g_j = (mu v_j - v_{j+1}) / lr) back
    through Learner.Nesterov must give the stored trace, including the projection (never active in this run)."""
    g = np.load(os.path.join(HERE, "golden", "quad_run.npz"))
    P, lr, mu = g["parameter_trace"], float(g["learning_rate"]), float(g["mu"])
    v = np.diff(P, axis=0)
    vprev = np.vstack([np.zeros(7), v[:-1]])
    grads = (mu * vprev - v) / lr
    calls = []

    def grad_fn(theta):
        j = len(calls)
        calls.append(np.array(theta))
        return float(g["loss_trace"][j]), grads[j]
    L = Learner(grad_fn, 7)
    L.load_optimization_function({"learning_rate": lr, "iter_num": 100, "method": "Nesterov", "mu": mu, "true_loss_print_flag": False})
    L.run(P[0])
    got = np.array(L.parameter_trace)
    assert got.shape == P.shape
    assert np.abs(got - P).max() < 1e-12
    assert np.allclose(L.loss_trace, g["loss_trace"])
    # the look-ahead points the gradients were requested at (theta_1 + 0.9 v_1)
    assert np.allclose(calls[1], [1.11134476, 0.1018448, 0.09903501, 0.10055924, 0.09941804, 0.09791882, -1.00165903], atol=5e-9)


def _quad(theta):
    A = np.diag([3.0, 1.0, 0.5])
    b = np.array([1.0, -2.0, 0.5])
    return float(0.5 * theta @ A @ theta - b @ theta) + 10.0, A @ theta - b


def test_adam_matches_torch_and_others_match_restatements():
    import torch
    th0 = np.array([1.0, 2.0, -1.0])
    # Adam == torch.optim.Adam (same bias-corrected form)
    L = Learner(_quad, 3)
    L.load_optimization_function({"learning_rate": 0.05, "iter_num": 20, "method": "Adam", "beta_1": 0.9, "beta_2": 0.999, "epsilon": 1e-8})
    L.run(th0)
    t = torch.tensor(th0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([t], lr=0.05, betas=(0.9, 0.999), eps=1e-8)
    for _ in range(20):
        opt.zero_grad()
        t.grad = torch.tensor(_quad(t.detach().numpy())[1])
        opt.step()
    assert np.allclose(L.parameter_trace[-1], t.detach().numpy(), rtol=1e-10, atol=1e-12)
    # Vanilla / AMSGrad / Nadam against line-by-line restatements
    for method in ("Vanilla", "AMSGrad", "Nadam"):
        L = Learner(_quad, 3)
        L.load_optimization_function({"learning_rate": 0.05, "iter_num": 15, "method": method, "beta_1": 0.9, "beta_2": 0.999, "epsilon": 1e-8})
        L.run(th0)
        th, m, v, vh = th0.copy(), np.zeros(3), np.zeros(3), np.zeros(3)
        for j in range(15):
            gk = _quad(th)[1]
            if method == "Vanilla":
                th = th - 0.05 * gk
            else:
                m = 0.9 * m + 0.1 * gk
                v = 0.999 * v + 0.001 * gk ** 2
                if method == "AMSGrad":
                    vh = np.maximum(vh, v)
                    th = th - 0.05 * m / (np.sqrt(vh) + 1e-8)
                else:
                    mh, vhh = m / (1 - 0.9 ** (j + 1)), v / (1 - 0.999 ** (j + 1))
                    th = th - 0.05 * (0.9 * mh + (0.1 / (1 - 0.9 ** (j + 1))) * gk) / (np.sqrt(vhh) + 1e-8)
            th[0] = max(th[0], 1e-8)
        assert np.allclose(L.parameter_trace[-1], th, rtol=1e-12, atol=1e-14), method


def test_projection_stop_rule_and_errors():
    L = Learner(lambda th: (5.0, np.array([100.0, 0.0])), 2)
    L.load_optimization_function({"learning_rate": 1.0, "iter_num": 3, "method": "Vanilla"})
    L.run([1.0, 1.0])
    assert L.parameter_trace[1][0] == 1e-8                       # projected (QuadAlgorithm.py:250)
    L = Learner(lambda th: (0.5, np.array([1.0, 0.0])), 2)        # loss below 0.9 after the first evaluation -> stop
    L.load_optimization_function({"learning_rate": 0.1, "iter_num": 10, "method": "Vanilla"})
    L.run([1.0, 1.0])
    assert len(L.loss_trace) == 1
    with pytest.raises(Exception, match="Wrong optimization method type!"):
        Learner(_quad, 3).load_optimization_function({"learning_rate": 0.1, "iter_num": 1, "method": "SGD"})


@pytest.mark.parametrize("method", ["Vanilla", "Nesterov", "NesterovTrue", "Adam", "Nadam", "AMSGrad"])
def test_device_learner_matches_host_learner(method):
    """DeviceLearner (update rule, projection, stop rule and traces in k_optim_* kernels around gradIterBatch; here on the
    thread-emulated library) against the host Learner fed by the same CUDA-path gradients: identical traces, bit for bit."""
    from tests.emu.support import emu_oc
    from lfsd_b200.optim import DeviceLearner, cpdp_grad_fn
    oc = emu_oc("pendulum")
    oc.aux_mode = oc.MODE_RK45
    x0 = np.zeros((2, 2))
    taus, wp, sel = np.array([0.3, 0.8]), np.array([[[1.4], [2.9]], [[1.5], [2.7]]]), [0]
    para = {"learning_rate": 0.02, "iter_num": 2, "method": method.replace("True", ""), "mu": 0.9,
            "true_loss_print_flag": method.endswith("True"), "beta_1": 0.9, "beta_2": 0.999, "epsilon": 1e-8}
    H = Learner(cpdp_grad_fn(oc, x0, 1.0, taus, wp, sel), 3)
    H.load_optimization_function(para)
    th_h = H.run([1.0, 0.5, 1.5])
    D = DeviceLearner(oc, x0, 1.0, taus, wp, sel)
    D.load_optimization_function(para)
    th_d = D.run([1.0, 0.5, 1.5])
    assert len(D.loss_trace) == len(H.loss_trace) == 2, (D.loss_trace, H.loss_trace)
    assert np.array_equal(np.array(D.parameter_trace), np.array(H.parameter_trace)), method
    assert np.array_equal(np.array(D.loss_trace), np.array(H.loss_trace))
    assert np.array_equal(th_d, th_h)


def test_device_learner_stop_rule_and_projection():
    from tests.emu.support import emu_oc
    from lfsd_b200.optim import DeviceLearner
    oc = emu_oc("pendulum")
    oc.aux_mode = oc.MODE_RK45
    # waypoint on the optimum's own trajectory start: loss 0 < 0.9 -> the rule stops after the first iteration
    D = DeviceLearner(oc, np.zeros((1, 2)), 1.0, np.array([0.0]), np.array([[[0.0]]]), [0])
    D.load_optimization_function({"learning_rate": 0.02, "iter_num": 3, "method": "Vanilla"})
    D.run([1.0, 0.5, 1.5])
    assert len(D.loss_trace) == 1 and len(D.parameter_trace) == 2 and D.stopped_early
    # a huge learning rate drives theta[0] negative: projected to 1e-8 (QuadAlgorithm.py:250); stop thresholds disabled
    D = DeviceLearner(oc, np.zeros((1, 2)), 1.0, np.array([0.5]), np.array([[[3.0]]]), [0])
    D.load_optimization_function({"learning_rate": 1e3, "iter_num": 1, "method": "Vanilla"})
    D.run([1.0, 0.5, 1.5], loss_stop=-1.0, grad_stop=-1.0)
    assert D.parameter_trace[1][0] == 1e-8 or D.parameter_trace[1][0] > 1.0
    with pytest.raises(Exception, match="Wrong optimization method type!"):
        D.load_optimization_function({"learning_rate": 0.1, "iter_num": 1, "method": "SGD"})
