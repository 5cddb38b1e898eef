"""Does the mixed feature model do what it is for?  A vehicle at rest sees five fixed landmarks and one target moving at
constant velocity, through range-bearing measurements with noise, missed detections and clutter.  After a few steps the
heaviest component of the dynamic map must sit on the target with its velocity, and the static map must hold the landmarks
(the truth-based check SURVEY 8(f) rank 1 asks for, applied to rank 4).  CPU: the oracle; GPU: the same run through
libphdslam.so, bit-identical."""
import numpy as np
import pytest

from oracle import oracle as O
import mixed_cases as MC

LANDMARKS = np.array([[5, 2], [8, -4], [-6, 3], [2, 9], [-3, -7]], float)
T0, TV = np.array([-8.0, -2.0]), np.array([0.8, 0.5])
STEPS = 40


def config(n=8):
    return MC.mixed_config(n, 8, motion_type=0, ax=0.0, ay=0.0, ayaw=0.0, dt=0.5, std_ax_features=0.3, std_ay_features=0.3,
                           cov_vx_birth=1.0, cov_vy_birth=1.0, tau=0.25, beta=8.0, ps=0.99, birth_weight=0.02, clutter_rate=2.0,
                           min_feature_weight=1e-5, min_separation=5.0, resample_threshold=0.0, max_components_dynamic=128)


def measurement_sets(dt):
    rng = np.random.default_rng(0)
    out = []
    for k in range(STEPS):
        t = T0 + TV * dt * k
        Z = []
        for q in list(LANDMARKS) + [t]:
            if rng.uniform() < 0.95:
                Z.append([np.hypot(*q) + rng.normal(0, 0.25), np.arctan2(q[1], q[0]) + rng.normal(0, 0.0087)])
        for _ in range(rng.poisson(2)):
            Z.append([rng.uniform(1, 15), rng.uniform(-3.1, 3.1)])
        out.append(np.array(Z, np.float32))
    return out


def run(filt, cfg):
    """returns per step (dynamic MAP map, static MAP map)"""
    hist = []
    for k, Z in enumerate(measurement_sets(cfg.dt)):
        if k:
            filt.phdPredict(None)
        filt.phdUpdateSynth(Z)
        e = filt.recoverSlamState()
        ds, dm = filt.get_maps_dynamic()
        ss, sm = filt.get_maps()
        lo, slo = int(ds[:e.map_particle].sum()), int(ss[:e.map_particle].sum())
        hist.append((dm[lo:lo + ds[e.map_particle]].copy(), sm[slo:slo + ss[e.map_particle]].copy()))
    return hist


def check_tracking(hist, dt):
    seen = np.zeros(len(LANDMARKS), int)
    tracked = 0
    for k in range(8, STEPS):
        d, s = hist[k]
        top = d[np.argmax(d["weight"])]
        truth = T0 + TV * dt * k
        tracked += (0.8 < top["weight"] < 1.2 and np.hypot(*(top["mean"][:2] - truth)) < 0.5 and
                    np.hypot(*(top["mean"][2:] - TV)) < 0.35)
        heavy = s[s["weight"] > 0.5]
        for i, q in enumerate(LANDMARKS):
            seen[i] += (np.hypot(heavy["mean"][:, 0] - q[0], heavy["mean"][:, 1] - q[1]) < 0.5).any()
    # every landmark is a static component of weight > 0.5 within 0.5 m, except in the step after a missed detection
    # (pd = 0.95: its weight drops to 1 - pd until it is seen again)
    assert (seen >= 0.85 * (STEPS - 8)).all(), seen
    # the heaviest dynamic component is the target: weight ~ 1, within 0.5 m and 0.35 m/s of the truth
    assert tracked >= 0.85 * (STEPS - 8), tracked


def test_moving_target_is_tracked_by_the_oracle():
    cfg = config()
    check_tracking(run(O.Oracle(cfg), cfg), cfg.dt)


@pytest.mark.gpu
def test_moving_target_is_tracked_on_the_gpu():
    import phdslam_b200 as P
    cfg = config()
    hg = run(P.PhdSlam(cfg), cfg)
    check_tracking(hg, cfg.dt)
    ho = run(O.Oracle(cfg), cfg)
    for k in range(STEPS):
        assert hg[k][0].tobytes() == ho[k][0].tobytes() and hg[k][1].tobytes() == ho[k][1].tobytes(), k
