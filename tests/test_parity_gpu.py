"""Parity of the sm_100a path (through the C-ABI of libphdslam.so) against the CPU oracle on identical
seeded inputs.  Bar (BASELINE.json north_star): component counts and ancestor indices bit-exact;
poses, map means, covariances and log-weights within 1e-4 relative.  Because both sides implement the
same canonical fp32 arithmetic (include/phd_detmath.h) the float outputs are in fact expected to be
bit-identical; tests assert the 1e-4 bar and report bit-exactness separately."""
import os

import numpy as np
import pytest

import phdslam_b200 as P
from phdslam_b200 import scene as S
from oracle import oracle as O
from conftest import DATA, GOLDEN

pytestmark = pytest.mark.gpu

RTOL = 1e-4   # north_star tolerance for floating-point outputs


def close(a, b, what, rtol=RTOL, atol=1e-7):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), "%s: max rel err %.3g" % (what, np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)))


def assert_gaussians(g, o, what):
    assert len(g) == len(o), what
    for f in ("weight", "mean", "cov"):
        close(g[f], o[f], what + "." + f)


def pair(cfg, sc=None):
    g, o = P.PhdSlam(cfg), O.Oracle(cfg)
    if sc is not None:
        S.load_scene(g, sc)
        S.load_scene(o, sc)
    return g, o


def assert_state_equal(g, o, what=""):
    gs, gm = g.get_maps()
    os_, om = o.get_maps()
    assert (gs == os_).all(), what + " component counts differ"           # bit-exact
    assert_gaussians(gm, om, what + " maps")
    close(g.log_weights, o.log_weights, what + " log-weights", atol=1e-6)
    gp, op = g.poses, o.poses
    for f in gp.dtype.names:
        close(gp[f], op[f], what + " pose." + f, atol=1e-6)
    return (gm.tobytes() == om.tobytes()) and (g.log_weights.tobytes() == o.log_weights.tobytes())


@pytest.mark.parametrize("Pn,C,M,near,far", [(8, 1, 1, 0, 0), (32, 48, 24, 7, 9), (16, 100, 37, 3, 0), (4, 0, 5, 0, 4)])
def test_dense_update_terms(Pn, C, M, near, far):
    """rows 3-6 of SURVEY 8(a): in-range split, births, EKF pre-update, PHD update, particle log-weight"""
    cfg = S.scene_config(Pn, C, M, max_components=256)
    sc = S.make_scene(Pn, C, M, seed=11, n_near=near, n_far=far)
    g, o = pair(cfg, sc)
    gt, gn, gd = g.update_terms(sc["Z"])
    ot, on, od = o.update_terms(sc["Z"])
    assert (gn == on).all() and (gn == C).all()
    assert_gaussians(gt, ot, "update terms")
    close(gd, od, "particle log-weight increment", atol=1e-5)
    # prune decisions (w < min_feature_weight) must agree exactly
    assert ((gt["weight"] < cfg.min_feature_weight) == (ot["weight"] < cfg.min_feature_weight)).all()
    assert gt.tobytes() == ot.tobytes(), "dense terms are expected to be bit-identical"
    assert gd.tobytes() == od.tobytes()


@pytest.mark.parametrize("weighting,metric", [(0, 0), (1, 0), (0, 1)])
def test_full_update_prune_merge_weights(weighting, metric):
    """rows 7-9: prune, merge (Mahalanobis / Hellinger), weight update + normalisation"""
    Pn, C, M = 48, 60, 20
    cfg = S.scene_config(Pn, C, M, max_components=256, particle_weighting=weighting, distance_metric=metric,
                         min_separation=10.0 if metric == 0 else 0.6)
    sc = S.make_scene(Pn, C, M, seed=3, n_near=5, n_far=6)
    g, o = pair(cfg, sc)
    g.phdUpdateSynth(sc["Z"])
    o.phdUpdateSynth(sc["Z"])
    assert assert_state_equal(g, o, "after update"), "state expected bit-identical"
    # second update on the merged maps (maps now contain merged + re-appended far components)
    Z2 = S.make_scene(Pn, C, M, seed=4)["Z"]
    g.phdUpdateSynth(Z2)
    o.phdUpdateSynth(Z2)
    assert assert_state_equal(g, o, "after second update")
    ge, oe = g.recoverSlamState(), o.recoverSlamState()
    close(ge.pose, oe.pose, "expected pose", atol=1e-6)
    close(ge.neff, oe.neff, "nEff")
    assert ge.map_particle == oe.map_particle


def test_batched_update_matches_single_batch():
    Pn, C, M = 64, 40, 16
    sc = S.make_scene(Pn, C, M, seed=8, n_near=2, n_far=2)
    cfg1 = S.scene_config(Pn, C, M, max_components=128)
    per_particle = (C * (M + 1) + M + 8) * 28
    cfg2 = S.scene_config(Pn, C, M, max_components=128, update_buffer_bytes=str(per_particle * 10))
    g1, g2 = P.PhdSlam(cfg1), P.PhdSlam(cfg2)
    for g in (g1, g2):
        S.load_scene(g, sc)
        g.phdUpdateSynth(sc["Z"])
    s1, m1 = g1.get_maps()
    s2, m2 = g2.get_maps()
    assert (s1 == s2).all() and m1.tobytes() == m2.tobytes()
    assert g1.log_weights.tobytes() == g2.log_weights.tobytes()


@pytest.mark.parametrize("motion", [1, 0])
def test_predict(motion):
    """rows 1-2: Ackerman / constant-velocity prediction with injected draws and with the Philox RNG"""
    n = 1000
    cfg = S.scene_config(n, 1, 1, motion_type=motion, acc_x=0.5, acc_y=0.2, acc_yaw=0.05, initial_vx=2.0,
                         initial_vyaw=0.2, seed="1234")
    g, o = pair(cfg)
    rng = np.random.default_rng(0)
    poses = np.zeros(n, dtype=P.POSE_DTYPE)
    for f in poses.dtype.names:
        poses[f] = rng.normal(0, 2, n)
    g.poses = poses
    o.poses = poses
    u = np.float32([2.2, -0.15])
    draws = rng.normal(0, 1, n * (2 if motion == 1 else 3))
    g.phdPredict(u, draws)
    o.phdPredict(u, draws)
    for f in poses.dtype.names:
        close(g.poses[f], o.poses[f], "injected " + f, atol=1e-6)
    assert g.poses.tobytes() == o.poses.tobytes()
    for k in range(3):   # counter-based RNG: same stream on both sides
        g.phdPredict(u)
        o.phdPredict(u)
    assert g.poses.tobytes() == o.poses.tobytes()
    assert np.abs(g.poses["px"] - poses["px"]).max() > 1e-3


@pytest.mark.parametrize("mode", [0, 1])
def test_resample_ancestors_bit_exact(mode):
    """row 12: ancestor indices are bit-exact; maps, poses and cardinalities follow their ancestors"""
    Pn, C, M = 512, 12, 6
    cfg = S.scene_config(Pn, C, M, max_components=64, resample_mode=mode, seed="99")
    sc = S.make_scene(Pn, C, M, seed=21)
    g, o = pair(cfg, sc)
    g.phdUpdateSynth(sc["Z"])
    o.phdUpdateSynth(sc["Z"])
    u = np.random.default_rng(5).uniform(0, 1, Pn + 1)
    ga = g.resampleParticles(u)
    oa = o.resampleParticles(u)
    assert (ga == oa).all()
    assert len(np.unique(ga)) < Pn          # the scene's weights are not uniform
    assert assert_state_equal(g, o, "after resample")
    assert (g.resample_idx == o.resample_idx).all()
    # second round with the on-device Philox uniforms
    g.phdUpdateSynth(sc["Z"])
    o.phdUpdateSynth(sc["Z"])
    ga = g.resampleParticles()
    oa = o.resampleParticles()
    assert (ga == oa).all()
    assert assert_state_equal(g, o, "after philox resample")


def test_step_loop_on_bundled_ackerman_data():
    """config #1 shape: run_synth loop (predict, update, estimate, nEff, resample) on the bundled Ackerman
    scene; every step's counts, ancestors and estimates must match the oracle"""
    cfg = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    cfg.set(n_particles=256, max_components=256, seed="7")
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_ackerman.txt"))
    U = P.load_controls(os.path.join(DATA, "controls_synth.txt"))
    g, o = pair(cfg)
    n_res = 0
    for k in range(40):
        u = U[k - 1] if k > 0 else None
        ge, gr = g.step(k, u, Z[k])
        oe, orr = o.step(k, u, Z[k])
        assert gr == orr, "resampling decision differs at step %d" % k
        n_res += gr
        close(ge.pose, oe.pose, "expected pose step %d" % k, atol=1e-6)
        close(ge.neff, oe.neff, "nEff step %d" % k)
        assert (g.map_sizes == o.map_sizes).all(), "component counts differ at step %d" % k
        assert (g.resample_idx == o.resample_idx).all(), "ancestors differ at step %d" % k
    assert n_res > 0
    assert_state_equal(g, o, "after 40 steps")


def test_step_loop_configs0_complete_run():
    """BASELINE configs[0] in full: all 331 steps of the bundled Ackerman scene at 256 particles -- every resampling decision,
    every component count and every ancestor index of the run equal the oracle's"""
    cfg = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    cfg.set(n_particles=256, max_components=256, seed="7")
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_ackerman.txt"))
    U = P.load_controls(os.path.join(DATA, "controls_synth.txt"))
    g = P.PhdSlam(cfg)
    o = O.Oracle(cfg, threads=os.cpu_count() or 1)
    n_res = 0
    for k in range(len(Z)):
        u = U[k - 1] if k > 0 else None
        ge, gr = g.step(k, u, Z[k])
        oe, orr = o.step(k, u, Z[k])
        assert gr == orr, "resampling decision differs at step %d" % k
        n_res += gr
        assert (g.map_sizes == o.map_sizes).all(), "component counts differ at step %d" % k
        assert (g.resample_idx == o.resample_idx).all(), "ancestors differ at step %d" % k
    assert len(Z) == 331 and n_res > 20
    assert_state_equal(g, o, "after the complete run")


def test_step_loop_constant_velocity_data():
    """config #2 shape (reduced particle count for the CPU oracle): CV motion model on measurements_synth_cv"""
    cfg = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    cfg.set(motion_type=0, n_particles=128, max_components=256, initial_vx=2.0, initial_vyaw=0.2, acc_x=0.5, acc_y=0.5,
            acc_yaw=0.087, dt=0.02, max_range=10.0, std_range=1.0, std_bearing=0.0349, seed="3")
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_cv.txt"))
    g, o = pair(cfg)
    for k in range(20):
        ge, gr = g.step(k, None, Z[k])
        oe, orr = o.step(k, None, Z[k])
        assert gr == orr
        assert (g.map_sizes == o.map_sizes).all()
    assert_state_equal(g, o, "after 20 CV steps")


def test_step_loop_configs1_at_full_size():
    """BASELINE configs[1] at its literal size: constant-velocity text data, 4096 particles, the first 120 steps of
    run_synth's loop (predict, update, estimate, nEff test, resampling) against the oracle (all host threads): resampling
    decisions, component counts and ancestor indices bit-exact at EVERY step, floats within 1e-4 at the end"""
    cfg = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    cfg.set(motion_type=0, n_particles=4096, max_components=256, initial_vx=2.0, initial_vyaw=0.2, acc_x=0.5, acc_y=0.5,
            acc_yaw=0.087, dt=0.02, max_range=10.0, std_range=1.0, std_bearing=0.0349, seed="3")
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_cv.txt"))
    g = P.PhdSlam(cfg)
    o = O.Oracle(cfg, threads=os.cpu_count() or 1)
    n_res = 0
    for k in range(120):
        ge, gr = g.step(k, None, Z[k])
        oe, orr = o.step(k, None, Z[k])
        assert gr == orr, "resampling decision differs at step %d" % k
        n_res += gr
        assert (g.map_sizes == o.map_sizes).all(), "component counts differ at step %d" % k
        assert (g.resample_idx == o.resample_idx).all(), "ancestors differ at step %d" % k
        close(ge.pose, oe.pose, "expected pose step %d" % k, atol=1e-6)
        close(ge.neff, oe.neff, "nEff step %d" % k)
    assert n_res >= 3
    assert_state_equal(g, o, "after 120 CV steps at 4096 particles")


def test_edge_cases():
    # empty measurement set: no-op (main.cpp:1258)
    cfg = S.scene_config(8, 4, 2, max_components=64)
    sc = S.make_scene(8, 4, 2, seed=1)
    g, o = pair(cfg, sc)
    g.phdUpdateSynth(np.zeros((0, 2), np.float32))
    assert assert_state_equal(g, o, "empty Z")
    # empty maps: only births survive (weight = w_b/(w_b+kappa))
    g2, o2 = pair(cfg)
    g2.phdUpdateSynth(sc["Z"])
    o2.phdUpdateSynth(sc["Z"])
    assert assert_state_equal(g2, o2, "empty map")
    assert (g2.map_sizes == 2).all()
    # labelled measurements (fields=3) with labeled_measurements=1: dynamic labels give zero weight
    cfg3 = S.scene_config(8, 4, 2, max_components=64, labeled_measurements=1, measurement_fields=3)
    g3, o3 = pair(cfg3, sc)
    Z3 = np.concatenate([sc["Z"], np.float32([[0], [1]])], 1)
    g3.phdUpdateSynth(Z3)
    o3.phdUpdateSynth(Z3)
    assert assert_state_equal(g3, o3, "labelled")
    # more than 256 measurements are truncated (src/phdfilter.cu:3390-3394)
    Zbig = S.make_scene(8, 4, 300, seed=2)["Z"]
    cfg4 = S.scene_config(8, 4, 256, max_components=512)
    g4, o4 = pair(cfg4, sc)
    g4.phdUpdateSynth(Zbig)
    o4.phdUpdateSynth(Zbig)
    assert assert_state_equal(g4, o4, "M > 256")


def test_capacity_overflow_is_reported():
    cfg = S.scene_config(4, 30, 40, max_components=32)
    sc = S.make_scene(4, 30, 40, seed=1)
    g = P.PhdSlam(cfg)
    S.load_scene(g, sc)
    with pytest.raises(P.PhdSlamError) as ei:
        g.phdUpdateSynth(sc["Z"])      # 30 surviving components + 40 births > 32
    assert ei.value.code == -3


def test_properties_at_scale():
    """size-independent properties on a scene too large for the oracle: weights normalised, counts bounded,
    dense-term invariant sum_j w_detect + w_birth = 1 - kappa/norm sampled, resampling ancestors sorted"""
    Pn, C, M = 4096, 64, 32
    cfg = S.scene_config(Pn, C, M, max_components=256)
    sc = S.make_scene(Pn, C, M, seed=2)
    g = P.PhdSlam(cfg)
    S.load_scene(g, sc)
    terms, nin, dlw = g.update_terms(sc["Z"])
    assert (nin == C).all()
    T = C * (M + 1) + M
    t = terms.reshape(Pn, T)[::257]
    det = t[:, C:C + M * C]["weight"].reshape(-1, M, C).astype(np.float64).sum(2)
    bw = t[:, C + M * C:]["weight"].astype(np.float64)
    np.testing.assert_allclose(det + bw, 1 - cfg.clutter_density / (cfg.birth_weight / bw), rtol=1e-4, atol=1e-6)
    g.phdUpdateSynth(sc["Z"])
    w = g.log_weights.astype(np.float64)
    assert abs(np.exp(w).sum() - 1) < 1e-4
    sizes = g.map_sizes
    assert sizes.min() >= C and sizes.max() <= C + M
    anc = g.resampleParticles()
    assert (np.diff(anc) >= 0).all() and anc.min() >= 0 and anc.max() < Pn
    assert (g.map_sizes == sizes[anc]).all()
    np.testing.assert_allclose(g.log_weights, -np.log(Pn), rtol=1e-6)


def test_map_estimates_map_and_eap():
    """row 10: MAP map = map of the heaviest particle; EAP map = computeExpectedMap + reduceGaussianMixture
    (main.cpp:290-316, gm_reduce.cpp:57-134) on the device: component count bit-exact, values within 1e-4"""
    Pn, C, M = 24, 30, 12
    cfg = S.scene_config(Pn, C, M, max_components=128, map_estimate=3)
    sc = S.make_scene(Pn, C, M, seed=6, n_near=3, n_far=3)
    g, o = pair(cfg, sc)
    g.phdUpdateSynth(sc["Z"])
    o.phdUpdateSynth(sc["Z"])
    gm, om = g.map_estimate(1), o.map_estimate(1)
    assert gm.tobytes() == om.tobytes() and len(gm) > 0
    ge, oe = g.map_estimate(2, cap=8192), o.map_estimate(2)
    assert len(ge) == len(oe) and len(ge) >= C          # bit-exact component count
    close(ge["weight"], oe["weight"], "EAP weight", atol=1e-9)
    close(ge["mean"], oe["mean"], "EAP mean", atol=1e-5)
    close(ge["cov"], oe["cov"], "EAP cov", atol=2e-6)
    # the expected map conserves the expected number of features
    w = np.exp(o.log_weights.astype(np.float64))
    sizes, maps = o.get_maps()
    exp_n = float((np.repeat(w, sizes) * maps["weight"]).sum())
    assert abs(ge["weight"].sum() - exp_n) < 1e-3 * exp_n


def test_cli_logs_match_oracle(tmp_path):
    """rows 13-14 + the process interface: `phdslam <cfg> synth` on the bundled Ackerman data writes
    state_estimateNNNNN.log files (README 5-line layout) identical to the oracle's, text for text.  The logged state is
    the one run_synth looks at: after recoverSlamState, before resampleParticles (src/main.cpp:1274-1279) -- the map is the
    maximum-weight particle's, the weights are the update's."""
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "cuda-phdslam_b200", "phdslam")
    assert os.path.exists(exe), "CLI not built"
    n_steps = 12
    out = tmp_path / "run"
    cmd = [exe, os.path.join(GOLDEN, "config_ackerman.cfg"), "synth", "--measurements",
           os.path.join(DATA, "measurements_synth_ackerman.txt"), "--controls", os.path.join(DATA, "controls_synth.txt"),
           "--out", str(out), "--steps", str(n_steps), "--set", "n_particles=64", "--set", "map_estimate=1", "--set", "seed=21",
           "--set", "resample_threshold=0.85", "--quiet"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    cfg = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    cfg.set(n_particles=64, map_estimate=1, seed="21", resample_threshold=0.85)
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_ackerman.txt"))
    U = P.load_controls(os.path.join(DATA, "controls_synth.txt"))
    o = O.Oracle(cfg)
    resampled_any = False
    for k in range(n_steps):
        e = o.step_filter(k, U[k - 1] if k > 0 else np.float32([0, 0]), Z[k])
        ref = tmp_path / ("ref%05d.log" % k)
        lw = o.log_weights
        P.write_log(str(ref), 0, e.pose, o.map_estimate(1), lw, o.poses, n_card=cfg.max_cardinality + 1)
        if k > 0:
            resampled_any |= o.step_resample(len(Z[k]), e)
            assert lw.max() > lw.min(), "the logged weights are the weighted (pre-resampling) ones"
        else:
            o.step_resample(len(Z[k]), e)
        got = open(out / ("state_estimate%05d.log" % k)).read()
        assert got == open(ref).read(), "log of step %d differs" % k
        assert len(got.split("\n")) == 6
    assert resampled_any, "the run must cover a resampling step"


# ---- prune + merge: the shared-memory kernel (merge_fast_kernel), its overflow queue and merge_kernel agree ----
def _with_cap(monkeypatch, cap):
    if cap is None:
        monkeypatch.delenv("PHDSLAM_MERGE_CAP", raising=False)
    else:
        monkeypatch.setenv("PHDSLAM_MERGE_CAP", cap)


@pytest.mark.parametrize("cap", [None, "0", "96", "160"])
def test_merge_kernels_match_oracle(monkeypatch, cap):
    """cap None: adaptive capacity (every particle in merge_fast_kernel); "0": merge_kernel only; "96"/"160": the
    particles above the capacity go through the overflow queue, the others stay in shared memory"""
    _with_cap(monkeypatch, cap)
    Pn, C, M = 40, 60, 20
    cfg = S.scene_config(Pn, C, M, max_components=256)
    sc = S.make_scene(Pn, C, M, seed=21, n_near=5, n_far=6)
    # uneven maps so that a fixed capacity splits the particles between the two kernels
    sizes = sc["sizes"].copy()
    maps = sc["maps"].reshape(Pn, -1).copy()
    for p in range(Pn):
        sizes[p] = maps.shape[1] - (p % 5) * 9
    g, o = P.PhdSlam(cfg), O.Oracle(cfg)
    for f in (g, o):
        f.poses = sc["poses"]
        f.log_weights = sc["log_weights"]
        f.set_maps(sizes, np.concatenate([maps[p, :sizes[p]] for p in range(Pn)]))
    for k in range(3):
        Z = S.make_scene(Pn, C, M, seed=21 + k)["Z"]
        g.phdUpdateSynth(Z)
        o.phdUpdateSynth(Z)
        assert assert_state_equal(g, o, "cap=%s update %d" % (cap, k)), "state expected bit-identical"


@pytest.mark.parametrize("cap", [None, "0"])
def test_merge_dense_clusters_and_ties(monkeypatch, cap):
    """many mutually overlapping components (long gate queues, one grid cell), exact weight ties, zero weights"""
    _with_cap(monkeypatch, cap)
    Pn, C, M = 12, 90, 12
    cfg = S.scene_config(Pn, C, M, max_components=256, min_separation=25.0)
    sc = S.make_scene(Pn, C, M, seed=5)
    maps = sc["maps"].reshape(Pn, C).copy()
    rng = np.random.Generator(np.random.Philox(77))
    centres = rng.uniform(-8, 8, (6, 2)).astype(np.float32)
    for p in range(Pn):
        which = rng.integers(0, 6, C)
        maps["mean"][p] = centres[which] + rng.normal(0, 0.15, (C, 2)).astype(np.float32)
        maps["weight"][p, ::3] = 0.25            # exact ties
        maps["weight"][p, 5] = 0.0
    sc["maps"] = maps.reshape(-1)
    g, o = pair(cfg, sc)
    for k in range(2):
        Z = S.make_scene(Pn, C, M, seed=40 + k)["Z"]
        g.phdUpdateSynth(Z)
        o.phdUpdateSynth(Z)
        assert assert_state_equal(g, o, "dense clusters, update %d" % k), "state expected bit-identical"


def test_merge_kernels_agree_at_scale(monkeypatch):
    """4096 particles: merge_fast_kernel (adaptive capacity, two steps so that it adapts) vs merge_kernel, bit for bit"""
    Pn, C, M = 4096, 96, 32
    cfg = S.scene_config(Pn, C, M, max_components=256)
    sc = S.make_scene(Pn, C, M, seed=9, n_near=4, n_far=3)
    res = []
    for cap in (None, "0"):
        _with_cap(monkeypatch, cap)
        g = P.PhdSlam(cfg)
        S.load_scene(g, sc)
        for k in range(2):
            g.phdUpdateSynth(S.make_scene(Pn, C, M, seed=9 + k)["Z"])
        sizes, maps = g.get_maps()
        res.append((sizes.copy(), maps.tobytes(), g.log_weights.tobytes()))
    assert (res[0][0] == res[1][0]).all()
    assert res[0][1] == res[1][1] and res[0][2] == res[1][2]


def test_update_merge_overlap_gives_identical_state():
    """phdslam_set_overlap(1): merge on a second stream under the next sub-batch's update -- same maps, same weights"""
    Pn, C, M = 6000, 48, 20
    cfg = S.scene_config(Pn, C, M, max_components=128)
    sc = S.make_scene(Pn, C, M, seed=12, n_near=2, n_far=2)
    res = []
    for on in (0, 1):
        g = P.PhdSlam(cfg)
        g.set_overlap(on)
        S.load_scene(g, sc)
        for k in range(2):
            g.phdUpdateSynth(S.make_scene(Pn, C, M, seed=12 + k)["Z"])
        sizes, maps = g.get_maps()
        res.append((sizes.copy(), maps.tobytes(), g.log_weights.tobytes()))
    assert (res[0][0] == res[1][0]).all() and res[0][1] == res[1][1] and res[0][2] == res[1][2]


def test_full_size_step_matches_oracle_on_a_particle_subset():
    """BASELINE configs[2] at full size (65 536 particles x 256 components x 64 measurements): particles are independent
    through predict / update / prune / merge, so the oracle run on 24 of them (same poses, same maps, the same Philox
    counters are NOT needed: the control noise is injected) must reproduce the device maps of those particles bit for
    bit; the cross-particle outputs are checked through their invariants."""
    Pn, C, M = 65536, 256, 64
    cfg = S.scene_config(Pn, C, M, max_components=384)
    sc = S.make_scene(Pn, C, M, seed=0)
    g = P.PhdSlam(cfg)
    S.load_scene(g, sc)
    g.phdUpdateSynth(sc["Z"])
    sizes, maps = g.get_maps()
    w = g.log_weights.astype(np.float64)
    assert abs(np.exp(w).sum() - 1.0) < 1e-4
    assert sizes.min() > C // 2 and sizes.max() <= C + M      # neighbouring landmarks of this dense scene merge
    off = np.concatenate([[0], np.cumsum(sizes)])
    pick = np.r_[0:8, 30000:30008, Pn - 8:Pn]
    sub_cfg = S.scene_config(len(pick), C, M, max_components=384)
    o = O.Oracle(sub_cfg, threads=os.cpu_count() or 1)
    o.poses = sc["poses"][pick]
    o.log_weights = sc["log_weights"][pick]
    n_all = len(sc["maps"]) // Pn
    o.set_maps(sc["sizes"][pick], np.concatenate([sc["maps"][p * n_all:(p + 1) * n_all] for p in pick]))
    o.phdUpdateSynth(sc["Z"])
    os_, om = o.get_maps()
    assert (os_ == sizes[pick]).all()
    gm = np.concatenate([maps[off[p]:off[p + 1]] for p in pick])
    assert gm.tobytes() == om.tobytes(), "maps of the sampled particles are expected to be bit-identical"
    # the particle log-weight increments: un-normalised differences agree with the oracle's
    ow = o.log_weights.astype(np.float64)
    gw = w[pick]
    np.testing.assert_allclose((gw - gw[0]), (ow - ow[0]), rtol=1e-4, atol=2e-4)
    anc = g.resampleParticles()
    assert (np.diff(anc) >= 0).all() and anc.min() >= 0 and anc.max() < Pn
    assert (g.map_sizes == sizes[anc]).all()


@pytest.mark.parametrize("Pn,C,M,cphd", [(131072, 128, 50, True), (65536, 128, 100, False)])
def test_full_size_shapes_of_configs3_and_configs4_on_a_particle_subset(Pn, C, M, cphd):
    """BASELINE configs[3]'s per-GPU shard (131 072 x 128 x 50, CPHD with the cardinality distribution) and configs[4]'s
    C x M shape (128 x 100, a quarter of its per-GPU particles at N = 8) at device scale: the oracle on 24 sampled
    particles must reproduce their device maps (and cardinality rows) bit for bit, as in the configs[2] test above."""
    extra = dict(filter_type=1, max_cardinality=255) if cphd else {}
    cfg = S.scene_config(Pn, C, M, max_components=256, **extra)
    sc = S.make_scene(Pn, C, M, seed=0)
    g = P.PhdSlam(cfg)
    S.load_scene(g, sc)
    g.phdUpdateSynth(sc["Z"])
    sizes, maps = g.get_maps()
    w = g.log_weights.astype(np.float64)
    assert abs(np.exp(w).sum() - 1.0) < 1e-4
    off = np.concatenate([[0], np.cumsum(sizes)])
    pick = np.r_[0:8, Pn // 2:Pn // 2 + 8, Pn - 8:Pn]
    sub_cfg = S.scene_config(len(pick), C, M, max_components=256, **extra)
    o = O.Oracle(sub_cfg, threads=os.cpu_count() or 1)
    o.poses = sc["poses"][pick]
    o.log_weights = sc["log_weights"][pick]
    n_all = len(sc["maps"]) // Pn
    o.set_maps(sc["sizes"][pick], np.concatenate([sc["maps"][p * n_all:(p + 1) * n_all] for p in pick]))
    o.phdUpdateSynth(sc["Z"])
    os_, om = o.get_maps()
    assert (os_ == sizes[pick]).all()
    gm = np.concatenate([maps[off[p]:off[p + 1]] for p in pick])
    assert gm.tobytes() == om.tobytes(), "maps of the sampled particles are expected to be bit-identical"
    if cphd:
        assert g.cardinalities[pick].tobytes() == o.cardinalities.tobytes()
    ow = o.log_weights.astype(np.float64)
    gw = w[pick]
    np.testing.assert_allclose((gw - gw[0]), (ow - ow[0]), rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("cphd", [False, True])
def test_update_modes_agree(cphd):
    """update_mode = 1 (the fused update emits only the prune survivors; production mode, bench.py's `production` key) gives
    the same maps, weights and cardinalities as update_mode = 0 (the reference-equivalent dense update terms are
    materialised in HBM first) -- bit for bit -- and both match the oracle"""
    Pn, C, M = 96, 70, 24
    extra = dict(filter_type=1, max_cardinality=127) if cphd else {}
    cfg = S.scene_config(Pn, C, M, max_components=256, **extra)
    sc = S.make_scene(Pn, C, M, seed=31, n_near=3, n_far=2)
    out = []
    for mode in (0, 1):
        cfg.set(update_mode=mode)
        g = P.PhdSlam(cfg)
        S.load_scene(g, sc)
        g.phdPredict(np.float32([1.0, 0.05]))
        g.phdUpdateSynth(sc["Z"])
        sizes, maps = g.get_maps()
        out.append((sizes, maps.tobytes(), g.log_weights.tobytes(), g.cardinalities.tobytes() if cphd else b""))
    assert (out[0][0] == out[1][0]).all() and out[0][1:] == out[1][1:]
    cfg.set(update_mode=0)
    o = O.Oracle(cfg)
    S.load_scene(o, sc)
    o.phdPredict(np.float32([1.0, 0.05]))
    o.phdUpdateSynth(sc["Z"])
    assert_state_equal(g, o, "update_mode 1")


def test_particle_dump_npz(tmp_path):
    """the reference's per-step particle dump (writeParticlesMat, src/main.cpp:594-713) as NPZ: every field round-trips"""
    Pn, C, M = 12, 20, 8
    cfg = S.scene_config(Pn, C, M, max_components=128, map_estimate=3)
    sc = S.make_scene(Pn, C, M, seed=2, n_near=1, n_far=1)
    g = P.PhdSlam(cfg)
    S.load_scene(g, sc)
    g.phdUpdateSynth(sc["Z"])
    path = g.save_particles(str(tmp_path), 7)
    assert os.path.basename(path) == "particles00007.npz"
    d = np.load(path)
    sizes, maps = g.get_maps()
    assert (np.diff(d["map_offsets"]) == sizes).all()
    assert d["maps_static.weights"].tobytes() == maps["weight"].tobytes() and d["maps_static.covs"].shape == (len(maps), 4)
    assert d["maps_static.means"].tobytes() == maps["mean"].tobytes()
    assert d["weights"].tobytes() == g.log_weights.tobytes() and d["states"].shape == (Pn, 6)
    assert (d["resample_idx"] == np.arange(Pn)).all()
    assert d["max_map_static.weights"].tobytes() == g.map_estimate(1)["weight"].tobytes()
    assert len(d["exp_map_static.weights"]) == len(g.map_estimate(2, cap=8192))
