"""Synthetic scenes for the parity tests and the benchmark.

Every particle gets exactly C in-range map components and every step exactly M measurements, so the
number of GM-PHD updates per step is exactly P*C*M.  Sensor and filter parameters are those of the
reference's cfg/config.cfg (:59-64, :139-152); the measurement model (true detections with Gaussian
range/bearing noise + uniform clutter) mirrors python/RangeBearingMeasurementModel.py:33-55.
"""
import numpy as np

from . import GAUSSIAN_DTYPE, POSE_DTYPE, Config, default_config


def scene_config(P, C, M, max_components=None, **overrides):
    """Config of the synthetic scene: cfg/config.cfg sensor + filter values."""
    kw = dict(
        motion_type=1, n_particles=P, filter_type=0, feature_model=0,
        max_range=15.0, max_bearing=3.141593, min_range=0.0, std_range=0.25, std_bearing=0.008727, pd=0.95,
        clutter_rate=20.0, birth_weight=1e-4, birth_noise_factor=1.0, min_feature_weight=1e-6, min_separation=10.0,
        particle_weighting=0, distance_metric=0, resample_threshold=0.5, map_estimate=0,
        l=1.415, h=0.38, a=1.89, b=0.5, std_encoder=1.0, std_alpha=0.034907, dt=0.1,
        max_components=max_components if max_components is not None else max(64, 2 * C + M),
    )
    kw.update(overrides)
    return default_config(**kw)


def scene_config_py(P, C, M, max_components=None, **overrides):
    """scene_config() WITHOUT libphdslam.so: the same phdslam_config_t image built field by field in Python (defaults of
    phdslam_config_defaults restated; tests/test_bench_host.py checks it is byte-identical to scene_config()).
    For bench.py's --impl reference arm, whose process must not load the product library."""
    f32 = np.float32
    c = Config()
    d = dict(motion_type=1, ax=0.5, ay=0.0, ayaw=0.0087, dt=0.1, max_bearing=float(f32(np.pi)), min_range=0.0, max_range=20.0,
             std_bearing=0.0524, std_range=1.0, clutter_rate=15.0, pd=0.98, n_particles=512, n_predict_particles=1,
             resample_threshold=0.15, subdivide_predict=1, birth_weight=0.05, birth_noise_factor=1.5, min_separation=5.0,
             min_feature_weight=0.00001, particle_weighting=1, max_cardinality=256, filter_type=1, map_estimate=1,
             max_steps=10000, n_steps=-1, data_directory=b"data/", measurement_fields=2, max_components=256,
             update_buffer_bytes=32 << 30, ps=0.98, tau=0.0, beta=1.0, max_components_dynamic=64)
    kw = dict(
        motion_type=1, n_particles=P, filter_type=0, feature_model=0,
        max_range=15.0, max_bearing=3.141593, min_range=0.0, std_range=0.25, std_bearing=0.008727, pd=0.95,
        clutter_rate=20.0, birth_weight=1e-4, birth_noise_factor=1.0, min_feature_weight=1e-6, min_separation=10.0,
        particle_weighting=0, distance_metric=0, resample_threshold=0.5, map_estimate=0,
        l=1.415, h=0.38, a=1.89, b=0.5, std_encoder=1.0, std_alpha=0.034907, dt=0.1,
        max_components=max_components if max_components is not None else max(64, 2 * C + M),
    )
    kw.update(overrides)
    d.update(kw)
    names = {n for n, _ in Config._fields_}
    for k, v in d.items():
        if k not in names:
            raise KeyError("scene_config_py: %r is not a plain field of phdslam_config_t" % k)
        setattr(c, k, int(v) if k == "seed" else v)
    # config.clutterDensity = config.clutterRate / (2*config.maxBearing*config.maxRange) (main.cpp:1065), in fp32
    c.clutter_density = float(f32(c.clutter_rate) / (f32(2.0) * f32(c.max_bearing) * f32(c.max_range)))
    return c


def make_scene(P, C, M, seed=0, n_far=0, n_near=0, max_range=15.0, std_range=0.25, std_bearing=0.008727, particle_seed=None):
    """Returns dict(poses, log_weights, sizes, maps, Z).

    n_far / n_near add components per particle outside the field of view ("far", class 0) and in the
    "nearly in range" shell (class 2) so that the in-range split is exercised.
    """
    rng = np.random.Generator(np.random.Philox(seed))
    # per-particle randomness (poses, map jitter) may come from its own stream so that several ranks can share
    # one landmark scene and one measurement set while holding different particles
    prng = rng if particle_seed is None else np.random.Generator(np.random.Philox([seed, 1000003 + particle_seed]))
    poses = np.zeros(P, dtype=POSE_DTYPE)
    poses["px"] = prng.normal(0, 0.1, P)
    poses["py"] = prng.normal(0, 0.1, P)
    poses["ptheta"] = prng.normal(0, 0.01, P)

    def ring(n, r_lo, r_hi):
        r = np.sqrt(rng.uniform(r_lo * r_lo, r_hi * r_hi, n))
        a = rng.uniform(-np.pi, np.pi, n)
        return np.stack([r * np.cos(a), r * np.sin(a)], 1)

    lm = ring(C, 1.0, 0.95 * max_range)
    near = ring(n_near, 1.06 * max_range, 1.14 * max_range)
    far = ring(n_far, 1.5 * max_range, 3.0 * max_range)
    allm = np.concatenate([lm, near, far], 0)
    kind = np.concatenate([np.zeros(C, int), np.ones(n_near, int), 2 * np.ones(n_far, int)])
    n_all = len(allm)
    # interleave classes so that the stable compaction has work to do
    perm = rng.permutation(n_all)
    allm, kind = allm[perm], kind[perm]
    phi = rng.uniform(0, np.pi, n_all)
    e1 = rng.uniform(0.01, 0.25, n_all)
    e2 = rng.uniform(0.01, 0.25, n_all)
    c, s = np.cos(phi), np.sin(phi)
    pxx = c * c * e1 + s * s * e2
    pxy = c * s * (e1 - e2)
    pyy = s * s * e1 + c * c * e2
    w = rng.uniform(0.1, 1.0, n_all)

    maps = np.zeros((P, n_all), dtype=GAUSSIAN_DTYPE)
    jit = prng.normal(0, 0.05, (P, n_all, 2)).astype(np.float32)
    maps["mean"] = (allm[None, :, :] + jit).astype(np.float32)
    maps["cov"][:, :, 0] = pxx[None, :]
    maps["cov"][:, :, 1] = pxy[None, :]
    maps["cov"][:, :, 2] = pxy[None, :]
    maps["cov"][:, :, 3] = pyy[None, :]
    maps["weight"] = w[None, :]
    sizes = np.full(P, n_all, dtype=np.int32)

    # measurements: round(0.6*M) detections of random in-range landmarks + clutter
    n_det = min(int(round(0.6 * M)), C)
    lm_in = allm[kind == 0]
    pick = rng.choice(len(lm_in), n_det, replace=False) if n_det > 0 else np.zeros(0, int)
    zr = np.hypot(lm_in[pick, 0], lm_in[pick, 1]) + rng.normal(0, std_range, n_det)
    zb = np.arctan2(lm_in[pick, 1], lm_in[pick, 0]) + rng.normal(0, std_bearing, n_det)
    cr = rng.uniform(0, max_range, M - n_det)
    cb = rng.uniform(-np.pi, np.pi, M - n_det)
    Z = np.stack([np.concatenate([zr, cr]), np.concatenate([zb, cb])], 1).astype(np.float32)
    Z[:, 0] = np.maximum(Z[:, 0], 0.05)
    Z = Z[rng.permutation(M)]
    logw = np.full(P, -np.log(np.float32(P)), dtype=np.float32)
    return dict(poses=poses, log_weights=logw, sizes=sizes, maps=maps.reshape(-1), Z=Z)


def load_scene(filt, sc):
    """Push a scene into a PhdSlam or Oracle instance."""
    filt.poses = sc["poses"]
    filt.log_weights = sc["log_weights"]
    filt.set_maps(sc["sizes"], sc["maps"])


# ---- mixed feature model (feature_model = 2): a dynamic map per particle next to the static scene ----
MIXED_KEYS = dict(feature_model=2, std_ax_features=0.5, std_ay_features=0.4, cov_vx_birth=0.25, cov_vy_birth=0.36, tau=0.3, beta=4.0,
                  ps=0.97)


def make_dynamic_maps(n, Cd, seed=1, particle_seed=None, max_range=15.0):
    """Cd constant-velocity features (Gaussian4D) in the field of view, the same for every particle up to a per-particle
    jitter of the means (as make_scene jitters the static maps).  Returns (sizes[n], maps[n * Cd], the Cd base features)."""
    from . import GAUSSIAN4_DTYPE
    rng = np.random.Generator(np.random.Philox(seed))
    prng = rng if particle_seed is None else np.random.Generator(np.random.Philox([seed, 2000003 + particle_seed]))
    base = np.zeros(Cd, GAUSSIAN4_DTYPE)
    r = np.sqrt(rng.uniform(1.0, (0.9 * max_range) ** 2, Cd))
    a = rng.uniform(-3.0, 3.0, Cd)
    for i in range(Cd):
        q = rng.normal(0, 1, (4, 4))
        q = q @ q.T / 4 + np.eye(4) * 0.5
        sc = np.array([0.05, 0.05, 0.3, 0.3])
        base["cov"][i] = (q * sc[:, None] * sc[None, :]).T.reshape(-1)
        base["mean"][i] = (r[i] * np.cos(a[i]), r[i] * np.sin(a[i]), rng.normal(0, 0.5), rng.normal(0, 0.5))
        base["weight"][i] = rng.uniform(0.2, 1.0)
    maps = np.tile(base, n)
    maps["mean"] += prng.normal(0, 0.05, maps["mean"].shape).astype(np.float32)
    return np.full(n, Cd, np.int32), maps, base


def mix_dynamic_measurements(Z, base, seed=2, every=3):
    """every `every`-th measurement of the static scene is replaced by a noisy observation of a dynamic feature"""
    Z = np.array(Z, np.float32).reshape(len(Z), -1).copy()
    rng = np.random.Generator(np.random.Philox(seed))
    for m in range(0, len(Z), every):
        q = base["mean"][int(rng.integers(len(base))), :2]
        Z[m, 0] = np.hypot(q[0], q[1]) + rng.normal(0, 0.25)
        Z[m, 1] = np.arctan2(q[1], q[0]) + rng.normal(0, 0.0087)
    return Z
