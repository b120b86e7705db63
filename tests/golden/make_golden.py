"""Extracts the reference's only stored numerical results into a small fixture.

Source : /root/reference/data/uav_results_random_20210308113016.mat  (struct 'results', written by
         /root/reference/lib/QuadAlgorithm.py:322-338) — a data file produced BY the reference, i.e. genuine
         reference outputs.  /root/reference does not exist on the GPU box, hence this committed copy.
Output : tests/golden/quad_run.npz   (run: python tests/golden/make_golden.py)
"""
import os
import numpy as np
import scipy.io as sio

SRC = '/root/reference/data/uav_results_random_20210308113016.mat'
CSV = '/root/reference/trajectories/20210308113016.csv'
here = os.path.dirname(os.path.abspath(__file__))

r = sio.loadmat(SRC, squeeze_me=True, struct_as_record=False)['results']
csv = np.loadtxt(CSV, delimiter=',')
out = dict(parameter_trace=np.asarray(r.parameter_trace, dtype=float),
           loss_trace=np.asarray(r.loss_trace, dtype=float),
           learning_rate=float(r.learning_rate),
           waypoints=np.asarray(r.waypoints, dtype=float),
           time_grid=np.asarray(r.time_grid, dtype=float),
           time_steps=np.asarray(r.time_steps, dtype=float),
           opt_state_traj=np.asarray(r.opt_state_traj, dtype=float),
           opt_control_traj=np.asarray(r.opt_control_traj, dtype=float),
           horizon=float(r.horizon), T=float(r.T),
           csv=csv)
# run configuration recovered in SURVEY.md §8c (Examples/quad_example_human_input.py:35-71)
out.update(ini_state=np.array([-2, -1, 0.6, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0], dtype=float),
           goal_position=np.array([2.5, 1.0, 1.5]), n_grid=25, steps_per_grid=4, mu=0.9)
np.savez_compressed(os.path.join(here, 'quad_run.npz'), **out)
print('wrote quad_run.npz', {k: np.shape(v) for k, v in out.items()})
