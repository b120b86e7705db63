"""What the reference's learner does after the last iteration: one more solve at the learned
parameter, the trajectory sampled on 101 points, and the two result files.

Mirrors ``/root/reference/lib/QuadAlgorithm.py``:
  * final solve + sampling   :299-312   ``time_steps = linspace(0, horizon, 101)``, ``opt_sol(time_steps)`` split into
                                          ``opt_state_traj`` (101 x n) and ``opt_control_traj`` (101 x m)
  * ``.mat`` file            :318-333   ``sio.savemat(name + '.mat', {'results': save_data})`` with the ten keys below
  * ``.csv`` file            :335-343   rows = [time_steps; first six states], ``np.savetxt(..., delimiter=",")``
and ``/root/reference/lib/InputWaypoints.py:212-228`` (``generate_time``).  The solve itself runs on the CUDA path
(``COCSys.cocSolver``); the writers only format its node table.
"""
import numpy as np

RESULT_KEYS = ('parameter_trace', 'loss_trace', 'learning_rate', 'waypoints', 'time_grid', 'time_steps',
               'opt_state_traj', 'opt_control_traj', 'horizon', 'T')


def final_trajectory(oc, ini_state, horizon, parameter, num=100 + 1):
    """QuadAlgorithm.py:299-312: (time_steps, opt_state_traj, opt_control_traj) at the learned parameter."""
    _, opt_sol = oc.cocSolver(ini_state, horizon, parameter)
    time_steps = np.linspace(0, horizon, num=num)
    opt_traj = opt_sol(time_steps)
    n, m = oc.n_state, oc.n_control
    return time_steps, opt_traj[:, :n], opt_traj[:, n:n + m]


def results_dict(parameter_trace, loss_trace, learning_rate, waypoints, time_list_sparse, time_steps, opt_state_traj,
                 opt_control_traj, horizon, T):
    """The ``save_data`` dictionary of QuadAlgorithm.py:320-329 (same keys, same order)."""
    vals = (parameter_trace, loss_trace, learning_rate, waypoints, time_list_sparse, time_steps, opt_state_traj,
            opt_control_traj, horizon, T)
    return dict(zip(RESULT_KEYS, vals))


def save_mat(path, save_data):
    """QuadAlgorithm.py:333 — MATLAB struct ``results``."""
    import scipy.io as sio
    sio.savemat(path, {'results': save_data})


def csv_array(time_steps, opt_state_traj):
    """QuadAlgorithm.py:338-342: 7 x num array [time; x; y; z; vx; vy; vz]."""
    posi_velo = np.transpose(np.array(opt_state_traj)[:, 0:6])
    return np.concatenate((np.array([time_steps]), posi_velo), axis=0)


def save_csv(path, time_steps, opt_state_traj):
    np.savetxt(path, csv_array(time_steps, opt_state_traj), delimiter=",")


def generate_time(waypoints_output, start_position, goal_position, quad_average_speed):
    """InputWaypoints.py:212-228: cumulative time stamps of [start] + waypoints + [goal], each segment
    ``round(distance / average_speed, 2)`` long; returns the list including 0.0 for the start."""
    waypoints_all = [start_position] + list(waypoints_output) + [goal_position]
    time_list_all = [0.0]
    for i in range(1, len(waypoints_all)):
        distance_current = np.linalg.norm(np.array(waypoints_all[i]) - np.array(waypoints_all[i - 1]))
        time_segment = round(distance_current / quad_average_speed, 2)
        time_list_all.append(time_segment + time_list_all[i - 1])
    return time_list_all
