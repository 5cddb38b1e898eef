"""Truth-based accuracy: OSPA / position error of a whole run on the reference's bundled Ackerman
scene against the simulation truth of matlab/simData2_ackerman.mat (tests/golden/truth_ackerman.npz), the scoring of
python/batch_analyze.py:16-40.  The CPU test drives the oracle, the GPU test the CUDA path; with parity both produce
the same numbers."""
import os

import numpy as np
import pytest

import phdslam_b200 as P
from phdslam_b200 import accuracy as A
from conftest import DATA, GOLDEN


def scene_cfg(n_particles):
    """sensor and vehicle parameters the scene was generated with (matlab/SynthSetup2.m:27-33,59-66, cfg/config.cfg.bak);
    the controls in the truth file are noise free, so the process noise is a tuning choice"""
    tr = A.Truth(os.path.join(GOLDEN, "truth_ackerman.npz"))
    cfg = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    cfg.set(n_particles=n_particles, max_components=512, seed="7", initial_x=float(tr.traj[0, 0]), initial_y=float(tr.traj[0, 1]),
            initial_yaw=float(tr.traj[0, 2]), max_range=10.0, std_range=1.0, std_bearing=0.0349, l=2.83, h=0.76, a=3.78, b=0.5,
            std_encoder=0.2, std_alpha=0.03, dt=1.0, birth_weight=0.005, min_feature_weight=1e-5)
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_ackerman.txt"))
    U = np.load(os.path.join(GOLDEN, "truth_ackerman.npz"))["controls"]
    return cfg, tr, Z, U


def test_ospa_known_answers():
    X = np.array([[0.0, 0.0], [1.0, 0.0], [4.0, 4.0]])
    assert A.ospa_distance(X, X) == (0.0, 0.0, 0.0)
    assert A.ospa_distance(np.zeros((0, 2)), np.zeros((0, 2))) == (0.0, 0.0, 0.0)
    assert A.ospa_distance(X, np.zeros((0, 2)), c=5.0) == (5.0, 0.0, 5.0)          # ospa.py:227-228
    # one point 1 m off, one missing: (1 + c)/3 with p = 1
    Y = np.array([[0.0, 1.0], [1.0, 0.0]])
    o, loc, cn = A.ospa_distance(X, Y, p=1.0, c=5.0)
    assert abs(o - (1.0 + 5.0) / 3.0) < 1e-12 and abs(loc - 1.0 / 3.0) < 1e-12 and abs(cn - 5.0 / 3.0) < 1e-12
    # the cut-off: a point 100 m away costs c; the assignment is optimal, not greedy
    o, _, _ = A.ospa_distance(np.array([[0.0, 0.0], [3.0, 0.0]]), np.array([[2.0, 0.0], [100.0, 0.0]]), c=5.0)
    assert abs(o - (1.0 + 5.0) / 2.0) < 1e-12
    # symmetric
    assert A.ospa_distance(X, Y) == A.ospa_distance(Y, X)


def test_map_extraction_and_log_parsing(tmp_path):
    w = np.array([0.9, 0.2, 0.95, 0.04])
    m = np.array([[0, 0], [1, 1], [2, 2], [3, 3]], dtype=float)
    est = A.extract_map_means(w, m)                        # round(2.09) = 2 heaviest
    assert est.tolist() == [[2.0, 2.0], [0.0, 0.0]]
    g = np.zeros(2, dtype=P.GAUSSIAN_DTYPE)
    g["weight"] = [0.7, 0.6]
    g["mean"] = [[1.5, -2.0], [4.0, 4.0]]
    g["cov"] = [[1, 0, 0, 1], [2, 0, 0, 2]]
    poses = np.zeros(3, dtype=P.POSE_DTYPE)
    path = str(tmp_path / "state_estimate00007.log")
    P.write_log(path, 0, np.float32([1, 2, 0.5, 0, 0, 0]), g, np.log(np.float32([0.5, 0.25, 0.25])), poses)
    r = A.parse_log(path)
    np.testing.assert_allclose(r["pose"][:3], [1, 2, 0.5])
    np.testing.assert_allclose(r["weights"], [0.7, 0.6], rtol=1e-5)
    np.testing.assert_allclose(r["means"], [[1.5, -2.0], [4.0, 4.0]])
    np.testing.assert_allclose(np.exp(r["log_weights"]).sum(), 1.0, rtol=1e-5)


def check(summary):
    assert summary["steps"] == 331
    assert summary["pose_rmse"] < 1.0, summary           # metres, over the whole 331-step run
    assert summary["ospa_final"] < 2.0, summary          # OSPA(p=1, c=5) of the final map
    assert abs(summary["n_est_final"] - summary["n_true_final"]) <= 8, summary


def test_oracle_tracks_the_ackerman_scene():
    from oracle import oracle as O
    cfg, tr, Z, U = scene_cfg(256)
    rows = A.run_and_score(O.Oracle(cfg, threads=os.cpu_count() or 1), tr, Z, U)
    check(A.summary(rows))


@pytest.mark.gpu
def test_cuda_tracks_the_ackerman_scene_and_scores_like_the_oracle():
    from oracle import oracle as O
    cfg, tr, Z, U = scene_cfg(256)
    rg = A.run_and_score(P.PhdSlam(cfg), tr, Z, U)
    check(A.summary(rg))
    ro = A.run_and_score(O.Oracle(cfg, threads=os.cpu_count() or 1), tr, Z, U)
    for a, b in zip(rg, ro):                               # same filter, same draws (counter-based RNG): same scores
        assert a["n_est"] == b["n_est"], a["step"]
        assert abs(a["pose_err"] - b["pose_err"]) < 1e-3 and abs(a["ospa"] - b["ospa"]) < 1e-3, a["step"]
    # BASELINE configs[0]/[1] particle counts on the device
    cfg4k, _, _, _ = scene_cfg(4096)
    check(A.summary(A.run_and_score(P.PhdSlam(cfg4k), tr, Z, U)))
