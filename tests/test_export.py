"""Result-file writers and the waypoint-time heuristic (lfsd_b200.export; /root/reference/lib/QuadAlgorithm.py:299-343,
/root/reference/lib/InputWaypoints.py:212-228) against the reference's stored run."""
import os

import numpy as np

import lfsd_b200  # noqa: F401
from lfsd_b200 import export

HERE = os.path.dirname(os.path.abspath(__file__))


def test_mat_and_csv_layout_match_the_stored_files(tmp_path):
    import scipy.io as sio
    g = np.load(os.path.join(HERE, "golden", "quad_run.npz"))
    d = export.results_dict(g["parameter_trace"].tolist(), g["loss_trace"].tolist(), float(g["learning_rate"]), g["waypoints"],
                            g["time_grid"], g["time_steps"], g["opt_state_traj"], g["opt_control_traj"], float(g["horizon"]), float(g["T"]))
    assert tuple(d) == export.RESULT_KEYS
    export.save_mat(str(tmp_path / "r.mat"), d)
    back = sio.loadmat(str(tmp_path / "r.mat"))["results"][0, 0]
    assert set(back.dtype.names) == set(export.RESULT_KEYS)            # field names of data/uav_results_random_*.mat
    assert back["parameter_trace"].shape == (101, 7) and back["loss_trace"].shape == (1, 100)
    assert back["opt_state_traj"].shape == (101, 13) and back["time_steps"].shape == (1, 101)
    assert np.array_equal(back["opt_control_traj"], g["opt_control_traj"])
    export.save_csv(str(tmp_path / "t.csv"), g["time_steps"], g["opt_state_traj"])
    csv = np.loadtxt(str(tmp_path / "t.csv"), delimiter=",")
    assert csv.shape == (7, 101) and np.array_equal(csv, g["csv"])        # trajectories/20210308113016.csv


def test_generate_time():
    t = export.generate_time([[1.0, 0.0, 0.0], [1.0, 2.0, 0.0]], [0.0, 0.0, 0.0], [1.0, 2.0, 1.5], 0.7)
    assert t == [0.0, round(1 / 0.7, 2), round(1 / 0.7, 2) + round(2 / 0.7, 2), round(1 / 0.7, 2) + round(2 / 0.7, 2) + round(1.5 / 0.7, 2)]
