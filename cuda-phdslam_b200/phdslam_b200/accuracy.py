"""Truth-based accuracy of a filter run: vehicle position error and OSPA map error per time step.

This is the scoring the reference does offline in python/batch_analyze.py:16-40 (Python 2 + a Cython Munkres,
python/ospa.py:221-268) and matlab/computeBatchResults.m:67-98, restated on numpy/scipy:

* position error  = |true (x, y) - expected (x, y)|                                  (batch_analyze.py:32, :43-44)
* map estimate    = the round(sum of weights) heaviest components of the map line    (batch_analyze.py:27-31)
* OSPA(p=1, c=5)  between the true landmarks seen so far and those means             (batch_analyze.py:33, ospa.py:221-268)
* nEff            = 1 / sum exp(2 w)                                                  (batch_analyze.py:38)

It scores either the `state_estimateNNNNN.log` files of a run directory (README:31-39, five lines per file) or a
filter object driven step by step (`run_and_score`).  The truth comes from matlab/simData2_*.mat, extracted into
tests/golden/truth_*.npz by tests/golden/make_truth.py.

    python -m phdslam_b200.accuracy RUN_DIR tests/golden/truth_ackerman.npz
"""
import glob
import os
import sys

import numpy as np
from scipy.optimize import linear_sum_assignment


def ospa_distance(X, Y, p=1.0, c=5.0):
    """(ospa, localisation part, cardinality part) of two point sets [n][2]; ospa.py:221-268 including its
    conventions for empty sets ((0,0,0) and (c,0,c))."""
    X = np.asarray(X, dtype=np.float64).reshape(-1, 2)
    Y = np.asarray(Y, dtype=np.float64).reshape(-1, 2)
    if len(X) == 0 and len(Y) == 0:
        return 0.0, 0.0, 0.0
    if len(X) == 0 or len(Y) == 0:
        return float(c), 0.0, float(c)
    if len(X) > len(Y):
        X, Y = Y, X
    m, n = len(X), len(Y)
    d = np.minimum(np.sqrt(((X[:, None, :] - Y[None, :, :]) ** 2).sum(2)), c)     # cut-off distance (munkres_step4.pyx compute_cost)
    rows, cols = linear_sum_assignment(d ** p)
    total_loc = float((d[rows, cols] ** p).sum())
    err_cn = (c ** p * (n - m) / n) ** (1.0 / p)
    err_loc = (total_loc / n) ** (1.0 / p)
    return ((total_loc + (n - m) * c ** p) / n) ** (1.0 / p), err_loc, err_cn


def extract_map_means(weights, means):
    """the round(sum w) heaviest components (batch_analyze.py:27-31)"""
    weights = np.asarray(weights, dtype=np.float64)
    means = np.asarray(means, dtype=np.float64).reshape(-1, 2)
    if len(weights) == 0:
        return means
    k = int(round(float(weights.sum())))
    order = np.argsort(weights)[::-1]
    return means[order[:max(k, 0)]]


class Truth(object):
    def __init__(self, path):
        z = np.load(path)
        self.traj, self.loc, self.off = z["traj"], z["loc"], z["off"]

    @property
    def n_steps(self):
        return len(self.traj)

    def map_at(self, k):
        return self.loc[self.off[k]:self.off[k + 1]]


def score_step(truth, k, pose_xy, weights, means, log_weights=None):
    est = extract_map_means(weights, means)
    o, ol, oc = ospa_distance(truth.map_at(k), est)
    row = {"step": k, "pose_err": float(np.hypot(*(truth.traj[k, :2] - np.asarray(pose_xy, dtype=np.float64)))),
           "ospa": o, "ospa_loc": ol, "ospa_cn": oc, "n_est": len(est), "n_true": len(truth.map_at(k))}
    if log_weights is not None:
        row["n_eff"] = float(1.0 / np.sum(np.exp(np.asarray(log_weights, dtype=np.float64)) ** 2))
    return row


def parse_log(path):
    """README:31-39: pose / map (7 numbers per Gaussian: w mx my c0 c1 c2 c3) / log weights / particle poses / cardinality"""
    with open(path) as f:
        lines = f.read().split("\n")
    nums = [np.array(ln.split(), dtype=np.float64) for ln in lines[:5]]
    g = nums[1].reshape(-1, 7) if nums[1].size else np.zeros((0, 7))
    return {"pose": nums[0], "weights": g[:, 0], "means": g[:, 1:3], "log_weights": nums[2]}


def score_run_dir(run_dir, truth):
    rows = []
    for path in sorted(glob.glob(os.path.join(run_dir, "state_estimate*.log"))):
        k = int(os.path.basename(path)[len("state_estimate"):-len(".log")])
        if k >= truth.n_steps:
            break
        r = parse_log(path)
        rows.append(score_step(truth, k, r["pose"][:2], r["weights"], r["means"], r["log_weights"]))
    return rows


def run_and_score(filt, truth, Z, U=None, n_steps=None, which_map=1):
    """Drives `filt` (PhdSlam or the oracle: same interface) through run_synth's loop (src/main.cpp:1231-1297)
    and scores every step against the truth; the map estimate is the MAP particle's map (which_map=1) or the EAP map (2)."""
    n = min(n_steps or len(Z), len(Z), truth.n_steps)
    rows = []
    for k in range(n):
        u = None if (U is None or k == 0) else U[k - 1]
        est = filt.step_filter(k, u, Z[k])
        m = filt.map_estimate(which_map)          # where recoverSlamState runs: before resampling (src/main.cpp:1274)
        filt.step_resample(len(Z[k]), est)
        rows.append(score_step(truth, k, est.pose[:2], m["weight"], m["mean"]))
    return rows


def summary(rows):
    a = lambda key: np.array([r[key] for r in rows], dtype=np.float64)
    return {"steps": len(rows), "pose_rmse": float(np.sqrt(np.mean(a("pose_err") ** 2))), "pose_err_final": float(a("pose_err")[-1]),
            "ospa_mean": float(a("ospa").mean()), "ospa_final": float(a("ospa")[-1]),
            "n_est_final": int(a("n_est")[-1]), "n_true_final": int(a("n_true")[-1])}


if __name__ == "__main__":
    rows_ = score_run_dir(sys.argv[1], Truth(sys.argv[2]))
    for r_ in rows_:
        print("%(step)5d pose_err %(pose_err)8.4f ospa %(ospa)7.4f loc %(ospa_loc)7.4f cn %(ospa_cn)7.4f n %(n_est)d/%(n_true)d" % r_)
    print(summary(rows_))
