"""n_predict_particles > 1 ("shotgun" prediction, reference src/phdfilter.cu:1091, 796-797, 1185-1238) and the
N > 5 n_particles down-sampling trigger of the loop (src/main.cpp:1286-1289): SURVEY 8(f) rank 3."""
import os

import numpy as np
import pytest

import phdslam_b200 as P
from conftest import DATA, GOLDEN


def _cfg(k, n=40, **kw):
    cfg = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    cfg.set(n_particles=n, n_predict_particles=k, max_components=128, seed="13", resample_threshold=0.05, **kw)
    return cfg


def _inputs():
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_ackerman.txt"))
    U = P.load_controls(os.path.join(DATA, "controls_synth.txt"))
    return Z, U


def test_oracle_particle_count_follows_the_reference_loop():
    from oracle import oracle as O
    Z, U = _inputs()
    o = O.Oracle(_cfg(3))
    counts, resampled = [], []
    for k in range(7):
        n_before = o.n
        e, r = o.step(k, U[k - 1] if k else None, Z[k])
        counts.append((n_before, o.n))
        resampled.append(r)
    # 40 -> (no predict at step 0) 40 -> 120 -> 360 > 200: resample to 40 -> 120 -> ...
    assert counts[0] == (40, 40) and counts[1] == (40, 120) and counts[2] == (120, 40) and resampled[2]
    assert counts[3] == (40, 120) and counts[4] == (120, 40)
    w = o.log_weights.astype(np.float64)
    assert abs(np.exp(w).sum() - 1.0) < 1e-4


def test_fan_out_duplicates_maps_and_scales_weights():
    from oracle import oracle as O
    from phdslam_b200 import scene as S
    cfg = S.scene_config(5, 6, 3, max_components=32, n_predict_particles=4, seed="3")
    sc = S.make_scene(5, 6, 3, seed=2)
    o = O.Oracle(cfg)
    S.load_scene(o, sc)
    w0 = o.log_weights.copy()
    s0, m0 = o.get_maps()
    o.phdPredict(np.float32([1.0, 0.05]))
    assert o.n == 20
    s1, m1 = o.get_maps()
    assert (s1 == np.repeat(s0, 4)).all()
    assert m1.tobytes() == np.concatenate([m0.reshape(5, 6)[i // 4] for i in range(20)]).tobytes()
    np.testing.assert_allclose(o.log_weights, np.repeat(w0, 4) - np.log(np.float32(4)), rtol=1e-6)
    p = o.poses
    assert len(np.unique(p["px"])) == 20               # every prediction has its own noise draw


@pytest.mark.gpu
@pytest.mark.parametrize("k,sub", [(3, 1), (2, 1), (6, 1), (2, 2)])
def test_cuda_shotgun_loop_matches_oracle(k, sub):
    from oracle import oracle as O
    Z, U = _inputs()
    cfg = _cfg(k, subdivide_predict=sub)
    g, o = P.PhdSlam(cfg), O.Oracle(cfg)
    seen = set()
    for step in range(9):
        u = U[step - 1] if step else None
        ge, gr = g.step(step, u, Z[step])
        oe, orr = o.step(step, u, Z[step])
        assert gr == orr and g.n_local == o.n, step
        seen.add((g.n_local, gr))
        np.testing.assert_allclose(ge.pose, oe.pose, rtol=1e-4, atol=1e-6)
        assert (g.map_sizes == o.map_sizes).all(), step
        assert (g.resample_idx == o.resample_idx).all(), step
        np.testing.assert_allclose(g.log_weights, o.log_weights, rtol=1e-4, atol=1e-6)
    assert len(seen) > 1
    gs, gm = g.get_maps()
    os_, om = o.get_maps()
    for f in ("weight", "mean", "cov"):
        np.testing.assert_allclose(gm[f], om[f], rtol=1e-4, atol=1e-7)


@pytest.mark.gpu
def test_cuda_resample_to_a_smaller_count():
    from oracle import oracle as O
    from phdslam_b200 import scene as S
    cfg = S.scene_config(60, 8, 4, max_components=64, n_predict_particles=2, seed="4")
    sc = S.make_scene(60, 8, 4, seed=6)
    g, o = P.PhdSlam(cfg), O.Oracle(cfg)
    for f in (g, o):
        S.load_scene(f, sc)
        f.phdPredict(np.float32([1.0, 0.0]))
        f.phdUpdateSynth(sc["Z"])
    assert g.n_local == 120
    anc = g.resampleParticles(n_new=60)
    oa = o.resampleParticles(n_new=60)
    assert g.n_local == 60 and o.n == 60
    assert (anc == oa).all()
    assert (g.map_sizes == o.map_sizes).all()
