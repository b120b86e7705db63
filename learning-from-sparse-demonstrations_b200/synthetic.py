"""Synthetic problem batches for the benchmark and the parity tests.

Headline batch: B quadrotor OCPs, RNG ``numpy.random.default_rng(20210308)``, draws problem-major in the order
start position ~ U(box), goal position ~ U(box), 5 x N(0, 0.1^2 I3) waypoint noise; box = the reference's
``config.json`` lab limits with the z floor raised to 0.2; start/goal velocity 0, attitude [1,0,0,0], body rate 0;
waypoints at tau_i = i/6 on the straight line start->goal plus the noise.  Shared theta = [1,.1,.1,.1,.1,.1,-1]
(``lib/QuadAlgorithm.py:235``), horizon normalised to 1 (``:222``).
"""
import math

import numpy as np

BOX_LO = np.array([-3.2, -1.6, 0.2])
BOX_HI = np.array([3.2, 1.6, 2.2])
QUAD_THETA0 = np.array([1.0, 0.1, 0.1, 0.1, 0.1, 0.1, -1.0])


def quad_batch(B, seed=20210308, n_waypoints=5):
    rng = np.random.default_rng(seed)
    x0 = np.zeros((B, 13))
    goal = np.zeros((B, 3))
    taus = np.array([(i + 1) / (n_waypoints + 1) for i in range(n_waypoints)])
    wp = np.zeros((B, n_waypoints, 3))
    for b in range(B):
        start = rng.uniform(BOX_LO, BOX_HI)
        g = rng.uniform(BOX_LO, BOX_HI)
        noise = rng.normal(0.0, 0.1, size=(n_waypoints, 3))
        x0[b, 0:3] = start
        x0[b, 6] = 1.0
        goal[b] = g
        wp[b] = start[None, :] + taus[:, None] * (g - start)[None, :] + noise
    return dict(x0=x0, goal=goal, taus=taus, wp=wp, theta=QUAD_THETA0.copy(), horizon=1.0, sel=[0, 1, 2])


def shard_bounds(B, world_size, rank):
    """Contiguous block sharding: rank g owns problems [g*B/G, (g+1)*B/G)."""
    assert B % world_size == 0, "batch must divide evenly over the ranks"
    per = B // world_size
    return rank * per, (rank + 1) * per


def robotarm_batch(B, seed=1):
    """B random initial parameters for Examples/robotarm_random.py: beta~U(1,6), weights~U(.5,1.5)."""
    rng = np.random.default_rng(seed)
    theta = np.concatenate([rng.uniform(1.0, 6.0, size=(B, 1)), rng.uniform(0.5, 1.5, size=(B, 4))], axis=1)
    x0 = np.tile(np.array([-math.pi / 2, 0.0, 0.0, 0.0]), (B, 1))
    taus = np.array([0.3])
    wp = np.tile(np.array([[[-math.pi / 4, 2 * math.pi / 3]]]), (B, 1, 1))
    return dict(x0=x0, theta=theta, taus=taus, wp=wp, horizon=1.0, sel=[0, 1])


def rocket_batch(B, seed=2):
    """B demos for Examples/rocket_groundtruth.py: initial positions [10,-8,3] + N(0,1), rest as :34-38."""
    rng = np.random.default_rng(seed)
    ang, axis = 1.0, np.array([0.0, -1.0, 1.0])
    axis = axis / np.linalg.norm(axis)
    quat = np.concatenate([[math.cos(ang / 2)], math.sin(ang / 2) * axis])
    x0 = np.zeros((B, 13))
    x0[:, 0:3] = np.array([10.0, -8.0, 3.0]) + rng.normal(0.0, 1.0, size=(B, 3))
    x0[:, 3:6] = [0.1, 0.0, 0.0]
    x0[:, 6:10] = quat
    return dict(x0=x0, theta_true=np.array([2.0] + [1.0] * 11), theta0=np.array([1.0] + [0.5] * 11), horizon=3.0,
                tau_idx=[1, 3, 6, 10, 13], sel=[0, 1, 2, 6, 7, 8, 9])
