"""Cases of the MIXED feature model (feature_model = 2: static + constant-velocity features) shared by
the golden generator (tests/golden/make_ref_mixed_golden.py), the CPU pin (tests/test_mixed_ref_pin.py) and the GPU parity
tests (tests/test_mixed_gpu.py).  One particle per case: the reference's phdUpdateKernelMixed reads the predicted weights
without the particle's offset (src/phdfilter.cu:2411,2437), so only particle 0 of a launch is computed as intended."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from phdslam_b200 import scene as S  # noqa: E402

G2, G4, POSE = O.GAUSSIAN_DTYPE, O.GAUSSIAN4_DTYPE, O.POSE_DTYPE

# name: (static in-range features, dynamic in-range features, measurements, labelled, particle weighting, seed)
MIXED_CASES = {
    "small": (5, 3, 4, False, 0, 1),
    "vo_weighting": (12, 7, 9, False, 1, 2),
    "dynamic_only": (0, 4, 3, False, 0, 3),
    "static_only": (6, 0, 5, False, 0, 4),
    "labelled": (20, 10, 16, True, 0, 5),
    "labelled_vo": (8, 5, 6, True, 1, 6),
    "wide": (40, 24, 30, False, 0, 7),
}

MIXED_OVERRIDES = dict(feature_model=2, std_ax_features=0.5, std_ay_features=0.4, cov_vx_birth=0.25, cov_vy_birth=0.36, tau=0.3,
                       beta=4.0, ps=0.97, birth_weight=0.01, max_components_dynamic=64)


def mixed_config(n_particles, M, labelled=False, weighting=0, **kw):
    ov = dict(MIXED_OVERRIDES)
    ov.update(particle_weighting=weighting, labeled_measurements=int(labelled), measurement_fields=3 if labelled else 2)
    ov.update(kw)
    return S.scene_config_py(n_particles, 8, M, max_components=kw.pop("max_components", 128), **ov)


def _spd4(rng, ps=0.05, vs=0.3):
    a = rng.normal(0, 1, (4, 4))
    q = a @ a.T / 4 + np.eye(4) * 0.5
    sc = np.array([ps, ps, vs, vs])
    return q * sc[:, None] * sc[None, :]


def _ring(rng, n, r_hi=13.0):
    r = np.sqrt(rng.uniform(1, r_hi ** 2, n))
    a = rng.uniform(-3, 3, n)
    return np.stack([r * np.cos(a), r * np.sin(a)], 1)


def static_features(rng, n, r_hi=13.0):
    sm = np.zeros(n, G2)
    pos = _ring(rng, n, r_hi)
    for i in range(n):
        b = rng.normal(0, 1, (2, 2))
        pm = (b @ b.T / 2 + np.eye(2) * 0.5) * 0.04
        sm["cov"][i] = pm.T.reshape(-1)
        sm["mean"][i] = pos[i]
        sm["weight"][i] = rng.uniform(0.2, 1)
    return sm


def dynamic_features(rng, n, r_hi=13.0):
    dm = np.zeros(n, G4)
    pos = _ring(rng, n, r_hi)
    for i in range(n):
        dm["cov"][i] = _spd4(rng).T.reshape(-1)
        dm["mean"][i, :2] = pos[i]
        dm["mean"][i, 2:] = rng.normal(0, 0.5, 2)
        dm["weight"][i] = rng.uniform(0.2, 1)
    return dm


def measurements(rng, pose, targets, M, labelled, n_static=None):
    """Range-bearing measurements of random targets (70 %) and clutter; label 0 = static, 1 = dynamic."""
    Z = []
    for _ in range(M):
        lab = float(rng.integers(2))
        if len(targets) and rng.uniform() < 0.7:
            k = int(rng.integers(len(targets)))
            q = targets[k]
            if n_static is not None:
                lab = 0.0 if k < n_static else 1.0
            dx, dy = q[0] - pose["px"], q[1] - pose["py"]
            r = np.hypot(dx, dy) + rng.normal(0, 0.25)
            b = np.arctan2(dy, dx) - pose["ptheta"] + rng.normal(0, 0.0087)
        else:
            r = rng.uniform(1, 14)
            b = rng.uniform(-3, 3)
        Z.append([r, b] + ([lab] if labelled else []))
    return np.array(Z, np.float32)


def build_case(name):
    ns, nd, M, labelled, weighting, seed = MIXED_CASES[name]
    rng = np.random.default_rng(1000 + seed)
    cfg = mixed_config(1, M, labelled, weighting)
    pose = np.zeros(1, POSE)
    pose["px"], pose["py"], pose["ptheta"] = rng.normal(0, 0.1), rng.normal(0, 0.1), rng.normal(0, 0.05)
    sm = static_features(rng, ns)
    dm = dynamic_features(rng, nd)
    targets = np.concatenate([sm["mean"], dm["mean"][:, :2]]) if ns + nd else np.zeros((0, 2))
    Z = measurements(rng, pose[0], targets, M, labelled, n_static=ns)
    return cfg, pose, sm, dm, Z


def reference_case(name):
    """Everything the reference's own kernels produce for a case (needs oracle/_ref/libphd_ref.so)."""
    from oracle import ref as R
    cfg, pose, sm, dm, Z = build_case(name)
    R.set_config(cfg)
    st, dt, fs, fd, dlogw = R.mixed_update_terms(pose, sm, dm, Z)
    pred, jump = R.predict_features4(dm)
    s_cand = st[fs == 0]
    d_cand = dt[fd == 0]
    out = dict(pose=pose, smap=sm, dmap=dm, Z=Z, s_terms=st, d_terms=dt, s_flags=fs, d_flags=fd, dlogw=np.float32(dlogw),
               predicted=pred, d_merged=R.merge4(d_cand), s_merged=R.merge([s_cand])[0] if len(s_cand) else s_cand)
    nd = len(dm)
    out["mahal"] = np.array([[R.mahalanobis4(dm[i], dm[j]) for j in range(nd)] for i in range(nd)], np.float32).reshape(nd, nd)
    return out


def relerr(a, b, floor):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0


def assert_gaussians_close(a, b, what, wtol=1e-4, wfloor=1e-7, mfloor=1e-2, cfloor=1e-4):
    """1e-4 relative (north_star's tolerance); means against a floor of 1e-2 (velocities near zero), covariances 1e-4."""
    assert len(a) == len(b), "%s: %d vs %d components" % (what, len(a), len(b))
    assert relerr(a["weight"], b["weight"], wfloor) < wtol, (what, "weight", relerr(a["weight"], b["weight"], wfloor))
    assert relerr(a["mean"], b["mean"], mfloor) < 1e-4, (what, "mean", relerr(a["mean"], b["mean"], mfloor))
    assert relerr(a["cov"], b["cov"], cfloor) < 1e-4, (what, "cov", relerr(a["cov"], b["cov"], cfloor))
