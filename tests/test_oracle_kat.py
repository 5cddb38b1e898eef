"""Known-answer tests that pin the CPU oracle (oracle/phd_oracle.cpp) to closed-form results,
independent float64 numpy restatements and the invariants of the GM-PHD recursion.
The reference ships no tests or golden vectors; see also test_ref_pin.py (the reference's own kernels run on the CPU)."""
import math

import numpy as np

import phdslam_b200 as P
from phdslam_b200 import scene as S
from oracle import oracle as O

G = P.GAUSSIAN_DTYPE


def gauss(w, mean, cov):
    g = np.zeros(1, dtype=G)
    g["weight"] = w
    g["mean"] = [mean]
    g["cov"] = [[cov[0][0], cov[1][0], cov[0][1], cov[1][1]]]   # column-major (src/slamtypes.h:123-127)
    return g


def ekf_numpy(pose, mean, cov, z, R):
    """textbook range-bearing EKF update in float64"""
    dx, dy = mean[0] - pose[0], mean[1] - pose[1]
    r2 = dx * dx + dy * dy
    r = math.sqrt(r2)
    b = math.atan2(dy, dx) - pose[2]
    H = np.array([[dx / r, dy / r], [-dy / r2, dx / r2]])
    Sg = H @ cov @ H.T + R
    K = cov @ H.T @ np.linalg.inv(Sg)
    nu = np.array([z[0] - r, (z[1] - b + math.pi) % (2 * math.pi) - math.pi])
    m = mean + K @ nu
    I_KH = np.eye(2) - K @ H
    Pn = I_KH @ cov @ I_KH.T + K @ R @ K.T
    loglik = -0.5 * nu @ np.linalg.inv(Sg) @ nu - math.log(2 * math.pi) - 0.5 * math.log(np.linalg.det(Sg))
    return m, Pn, loglik, r, b


def one_particle_filter(cfg, pose, comps):
    o = O.Oracle(cfg)
    p = np.zeros(1, dtype=P.POSE_DTYPE)
    p["px"], p["py"], p["ptheta"] = pose
    o.poses = p
    o.set_maps([len(comps)], comps)
    return o


def test_single_component_update_closed_form():
    cfg = S.scene_config(1, 1, 1, max_components=32)
    pose = (0.3, -0.2, 0.1)
    mean = np.array([4.0, 3.0])
    cov = np.array([[0.2, 0.05], [0.05, 0.1]])
    w0 = 0.8
    o = one_particle_filter(cfg, pose, gauss(w0, mean, cov))
    z = np.array([[5.1, 0.55]], dtype=np.float32)
    terms, nin, dlw = o.update_terms(z)
    assert nin[0] == 1 and len(terms) == 1 * 2 + 1
    R = np.diag([cfg.std_range ** 2, cfg.std_bearing ** 2])
    m, Pn, ll, r, b = ekf_numpy(pose, mean, cov, z[0], R)
    nd, det, birth = terms[0], terms[1], terms[2]
    # non-detection term: w*(1-pd), same Gaussian (src/phdfilter.cu:2137-2141)
    assert abs(nd["weight"] - w0 * (1 - cfg.pd)) < 1e-7
    np.testing.assert_allclose(nd["mean"], mean, rtol=1e-6)
    # detection term
    np.testing.assert_allclose(det["mean"], m, rtol=2e-5)
    np.testing.assert_allclose(det["cov"].reshape(2, 2).T, Pn, rtol=2e-3, atol=2e-6)
    lik = cfg.pd * w0 * math.exp(ll)
    kappa, wb = cfg.clutter_density, cfg.birth_weight
    norm = lik + kappa + wb
    assert abs(det["weight"] - lik / norm) < 2e-4 * lik / norm
    assert abs(birth["weight"] - wb / norm) < 2e-4 * wb / norm
    # particle log-weight increment, scheme 0: sum_m log(norm_m) - (sum pd*w + M*w_b) (:2258-2262)
    assert abs(dlw[0] - (math.log(norm) - (cfg.pd * w0 + wb))) < 1e-4


def test_birth_covariance_matches_numeric_jacobian():
    # host birth loop (src/phdfilter.cu:3468-3507): cov = J diag(var) J^T with J = d(x,y)/d(r,b)
    cfg = S.scene_config(1, 0, 1, max_components=32, birth_noise_factor=1.5)
    pose = (1.0, 2.0, 0.3)
    o = one_particle_filter(cfg, pose, np.zeros(0, dtype=G))
    z = np.array([[6.0, -0.4]], dtype=np.float32)
    terms, nin, _ = o.update_terms(z)
    assert nin[0] == 0 and len(terms) == 1
    bt = terms[0]
    th = pose[2] + z[0, 1]
    np.testing.assert_allclose(bt["mean"], [pose[0] + 6 * math.cos(th), pose[1] + 6 * math.sin(th)], rtol=1e-6)
    J = np.array([[math.cos(th), -6 * math.sin(th)], [math.sin(th), 6 * math.cos(th)]])
    V = np.diag([(cfg.std_range * 1.5) ** 2, (cfg.std_bearing * 1.5) ** 2])
    np.testing.assert_allclose(bt["cov"].reshape(2, 2).T, J @ V @ J.T, rtol=1e-5, atol=1e-9)
    # empty map: normaliser = clutterDensity + birthWeight (:2211-2215)
    assert abs(bt["weight"] - cfg.birth_weight / (cfg.birth_weight + cfg.clutter_density)) < 1e-7


def test_update_weight_normalisation_invariant():
    # for every measurement: sum_j w_detect(j,m) + w_birth(m) = 1 - kappa/norm_m  (Vo & Ma 2006)
    P_, C_, M_ = 3, 40, 12
    cfg = S.scene_config(P_, C_, M_)
    sc = S.make_scene(P_, C_, M_, seed=5)
    o = O.Oracle(cfg)
    S.load_scene(o, sc)
    terms, nin, dlw = o.update_terms(sc["Z"])
    assert (nin == C_).all()
    T = C_ * (M_ + 1) + M_
    t = terms.reshape(P_, T)
    for p in range(P_):
        det = t[p, C_:C_ + M_ * C_]["weight"].reshape(M_, C_).astype(np.float64)
        bw = t[p, C_ + M_ * C_:]["weight"].astype(np.float64)
        norm = cfg.birth_weight / bw
        np.testing.assert_allclose(det.sum(1) + bw, 1 - cfg.clutter_density / norm, rtol=2e-5, atol=5e-7)
        nd = t[p, :C_]["weight"]
        np.testing.assert_allclose(nd, sc["maps"].reshape(P_, -1)[p]["weight"] * np.float32(1 - cfg.pd), rtol=1e-6)


def test_in_range_split_classes():
    # computeInRangeKernel (src/phdfilter.cu:1328-1346)
    cfg = S.scene_config(1, 1, 1, max_components=32, min_range=1.0, max_bearing=1.0)
    pts = [(5, 0, 1), (14.9, 0, 1), (15.1, 0, 2), (17.9, 0, 2), (18.1, 0, 0), (0.9, 0, 2), (0.7, 0, 0),
           (5 * math.cos(1.1), 5 * math.sin(1.1), 2), (5 * math.cos(1.3), 5 * math.sin(1.3), 0), (-5, 0, 0)]
    comps = np.concatenate([gauss(0.5, (x, y), [[0.1, 0], [0, 0.1]]) for x, y, _ in pts])
    o = one_particle_filter(cfg, (0, 0, 0), comps)
    z = np.array([[5.0, 0.0]], dtype=np.float32)
    _, nin, _ = o.update_terms(z, want_terms=False)
    assert nin[0] == sum(1 for p in pts if p[2] == 1)
    o.phdUpdateSynth(z)
    sizes, m = o.get_maps()
    # far components (class 0) are re-appended unchanged at the end, in order (:3311-3318)
    far = [p for p in pts if p[2] == 0]
    np.testing.assert_allclose(m["mean"][-len(far):], np.float32([[p[0], p[1]] for p in far]), rtol=1e-6)
    assert (m["weight"][-len(far):] == np.float32(0.5)).all()


def test_merge_two_identical_gaussians():
    cfg = S.scene_config(1, 1, 1)
    g = gauss(0.4, (1.0, 2.0), [[0.3, 0.1], [0.1, 0.2]])
    lib = O.load()
    inp = np.concatenate([g, g, gauss(0.2, (10.0, 10.0), [[0.1, 0], [0, 0.1]])])
    out = np.zeros(3, dtype=G)
    import ctypes
    n = lib.oracle_merge(ctypes.byref(cfg), inp.ctypes.data, 3, out.ctypes.data)
    assert n == 2
    assert abs(out[0]["weight"] - 0.8) < 1e-7
    np.testing.assert_allclose(out[0]["mean"], [1.0, 2.0], rtol=1e-7)
    np.testing.assert_allclose(out[0]["cov"], [0.3, 0.1, 0.1, 0.2], rtol=1e-6)
    assert abs(out[1]["weight"] - 0.2) < 1e-7      # output ordered by descending seed weight
    assert lib.oracle_mahalanobis(g.ctypes.data, g.ctypes.data) == 0.0


def test_merge_moment_matching_against_numpy():
    cfg = S.scene_config(1, 1, 1, min_separation=1e9)   # everything merges into one
    rng = np.random.default_rng(0)
    n = 7
    inp = np.zeros(n, dtype=G)
    inp["weight"] = rng.uniform(0.1, 1, n)
    inp["mean"] = rng.normal(0, 0.3, (n, 2))
    for i in range(n):
        a = rng.normal(0, 1, (2, 2))
        c = a @ a.T + 0.5 * np.eye(2)
        inp["cov"][i] = [c[0, 0], c[1, 0], c[0, 1], c[1, 1]]
    out = np.zeros(n, dtype=G)
    import ctypes
    assert O.load().oracle_merge(ctypes.byref(cfg), inp.ctypes.data, n, out.ctypes.data) == 1
    w = inp["weight"].astype(np.float64)
    mu = (w[:, None] * inp["mean"]).sum(0) / w.sum()
    cov = np.zeros((2, 2))
    for i in range(n):
        d = mu - inp["mean"][i]
        cov += w[i] * (inp["cov"][i].reshape(2, 2).T + np.outer(d, d))
    cov /= w.sum()
    assert abs(out[0]["weight"] - w.sum()) < 1e-6
    np.testing.assert_allclose(out[0]["mean"], mu, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out[0]["cov"].reshape(2, 2).T, cov, rtol=1e-5)
    assert out[0]["cov"][1] == out[0]["cov"][2]     # force_symmetric_covariance


def test_weights_normalised_after_update():
    P_, C_, M_ = 16, 20, 8
    cfg = S.scene_config(P_, C_, M_)
    sc = S.make_scene(P_, C_, M_, seed=1)
    o = O.Oracle(cfg)
    S.load_scene(o, sc)
    o.phdUpdateSynth(sc["Z"])
    w = o.log_weights.astype(np.float64)
    assert abs(np.exp(w).sum() - 1) < 1e-5
    e = o.recoverSlamState()
    poses = o.poses
    np.testing.assert_allclose(e.px, (np.exp(w) * poses["px"]).sum(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(e.ptheta, (np.exp(w) * poses["ptheta"]).sum(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(e.neff, 1.0 / np.exp(2 * w).sum() / P_, rtol=1e-5)
    assert e.map_particle == int(np.argmax(o.log_weights))


def test_resample_equal_weights_identity_and_literal_agreement():
    n = 64
    cfg = S.scene_config(n, 1, 1)
    o = O.Oracle(cfg)
    u = np.full(n + 1, 0.5)
    anc = o.resampleParticles(u)
    assert (anc == np.arange(n)).all()
    assert (o.resample_idx == np.arange(n)).all()
    np.testing.assert_allclose(o.log_weights, -math.log(n), rtol=1e-6)
    rng = np.random.default_rng(7)
    for trial in range(20):
        o1, o2 = O.Oracle(cfg), O.Oracle(cfg)
        w = rng.normal(0, 2, n)
        w = (w - np.log(np.exp(w).sum())).astype(np.float32)
        poses = np.zeros(n, dtype=P.POSE_DTYPE)
        poses["px"] = np.arange(n)
        for o_ in (o1, o2):
            o_.log_weights = w
            o_.poses = poses
        u = rng.uniform(0, 1, n + 1)
        a1 = o1.resampleParticles(u, literal=False)
        a2 = o2.resampleParticles(u, literal=True)    # reference's sequential double CDF walk (main.cpp:461-499)
        assert (a1 == a2).all()
        assert (np.diff(a1) >= 0).all()
        assert (o1.poses["px"] == a1).all()           # copy_particles (slamtypes.h:313-333)


def test_resample_systematic_mode():
    n = 32
    cfg = S.scene_config(n, 1, 1, resample_mode=1)
    o = O.Oracle(cfg)
    w = np.full(n, -20.0, dtype=np.float32)
    w[5] = 0.0
    o.log_weights = w
    anc = o.resampleParticles(np.full(n + 1, 0.3))
    assert (anc == 5).all()


def test_reduce_mixture_eap():
    # reduceGaussianMixture (src/gm_reduce.cpp:57-134)
    a = gauss(0.5, (0, 0), [[0.1, 0], [0, 0.1]])
    b = gauss(0.3, (0.05, 0), [[0.1, 0], [0, 0.1]])
    c = gauss(0.9, (5, 5), [[0.1, 0], [0, 0.1]])
    inp = np.concatenate([a, b, c])
    out = np.zeros(3, dtype=G)
    n = O.load().oracle_reduce_mixture(inp.ctypes.data, 3, 5.0, out.ctypes.data)
    assert n == 2
    assert abs(out[0]["weight"] - 0.9) < 1e-7 and abs(out[1]["weight"] - 0.8) < 1e-7
    np.testing.assert_allclose(out[1]["mean"], [0.3 * 0.05 / 0.8, 0], atol=1e-7)


def test_predict_ackerman_and_cv_against_numpy():
    n = 8
    cfg = S.scene_config(n, 1, 1)
    o = O.Oracle(cfg)
    poses = np.zeros(n, dtype=P.POSE_DTYPE)
    poses["ptheta"] = np.linspace(-3, 3, n)
    o.poses = poses
    draws = np.random.default_rng(3).normal(0, 1, 2 * n)
    u = np.float32([2.5, 0.1])
    o.phdPredict(u, draws)
    got = o.poses
    for i in range(n):   # phdPredictKernelAckerman (src/phdfilter.cu:785-825)
        ve = u[0] + cfg.std_encoder * draws[2 * i + 1]
        al = u[1] + cfg.std_alpha * draws[2 * i]
        vc = ve / (1 - math.tan(al) * cfg.h / cfg.l)
        th = poses["ptheta"][i]
        thd = vc * math.tan(al) / cfg.l
        x = cfg.dt * (vc * math.cos(th) - thd * (cfg.a * math.sin(th) + cfg.b * math.cos(th)))
        y = cfg.dt * (vc * math.sin(th) + thd * (cfg.a * math.cos(th) - cfg.b * math.sin(th)))
        assert abs(got["px"][i] - x) < 1e-5 and abs(got["py"][i] - y) < 1e-5
        t = th + cfg.dt * thd
        assert abs(math.remainder(got["ptheta"][i] - t, 2 * math.pi)) < 1e-5
    cfg2 = S.scene_config(n, 1, 1, motion_type=0, acc_x=0.5, acc_y=0.2, acc_yaw=0.05, initial_vx=2.0, initial_vyaw=0.2)
    o2 = O.Oracle(cfg2)
    d3 = np.random.default_rng(4).normal(0, 1, 3 * n)
    o2.phdPredict(None, d3)
    g2 = o2.poses
    dt = cfg2.dt
    for i in range(n):   # phdPredictKernel (src/phdfilter.cu:827-859), noise = 3*sigma*randn (:1115-1117)
        ax, ay, at = 3 * 0.5 * d3[3 * i], 3 * 0.2 * d3[3 * i + 1], 3 * 0.05 * d3[3 * i + 2]
        assert abs(g2["px"][i] - (dt * 2.0 + 0.5 * dt * dt * ax)) < 1e-5
        assert abs(g2["py"][i] - (0.5 * dt * dt * ay)) < 1e-5
        assert abs(g2["vx"][i] - (2.0 + dt * ax)) < 1e-5
        assert abs(g2["vtheta"][i] - (0.2 + dt * at)) < 1e-5


def test_multi_step_filter_tracks_truth():
    """statistical sanity on the bundled Ackerman scene: the oracle filter runs 40 steps and stays finite"""
    import os
    from conftest import DATA, GOLDEN
    cfg = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    cfg.set(n_particles=32, max_components=256)
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_ackerman.txt"))
    U = P.load_controls(os.path.join(DATA, "controls_synth.txt"))
    o = O.Oracle(cfg)
    for k in range(25):
        e, res = o.step(k, U[k - 1] if k > 0 else None, Z[k])
        assert np.isfinite(e.pose).all() and 0 < e.neff <= 1.0 + 1e-5
    assert o.map_sizes.max() < 256 and o.map_sizes.min() > 3
