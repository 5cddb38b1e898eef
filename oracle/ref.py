"""ctypes loader for oracle/_ref/libphd_ref.so -- the REFERENCE's own kernels compiled for the CPU.

TEST INFRASTRUCTURE ONLY (see oracle/ref_build.sh, oracle/ref_harness.cpp).  The library exists only where
oracle/ref_build.sh has run (a container that has /root/reference); `available()` says whether it does.
Used by tests/test_ref_pin.py and tests/golden/make_ref_golden.py.
"""
import ctypes as C
import os

import numpy as np

from .oracle import GAUSSIAN_DTYPE, POSE_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHD_REF_LIB") or os.path.join(_HERE, "_ref", "libphd_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        for n in ("ref_wrap_angle", "ref_safe_log"):
            getattr(lib, n).restype = C.c_float
            getattr(lib, n).argtypes = [C.c_float]
        for n in ("ref_mahalanobis", "ref_hellinger"):
            getattr(lib, n).restype = C.c_float
            getattr(lib, n).argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_log_sum_exp.restype = C.c_float
        lib.ref_log_sum_exp.argtypes = [C.c_void_p, C.c_int]
        lib.ref_sum_by_reduction.restype = C.c_float
        lib.ref_sum_by_reduction.argtypes = [C.c_void_p]
        lib.ref_set_config.argtypes = [C.c_void_p]
        lib.ref_predict.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_in_range.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_birth_device.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_void_p]
        lib.ref_update_terms.restype = C.c_size_t
        lib.ref_update_terms.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        lib.ref_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.ref_update.restype = C.c_size_t
        lib.ref_update.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_size_t]
        lib.ref_resample.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_recover.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.ref_neff.restype = C.c_float
        lib.ref_neff.argtypes = [C.c_void_p, C.c_int]
        lib.ref_mahalanobis4.restype = C.c_float
        lib.ref_mahalanobis4.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_predict_features4.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.ref_mixed_update_terms.restype = C.c_float
        lib.ref_mixed_update_terms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_merge4.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib = lib
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def set_config(cfg):
    load().ref_set_config(C.byref(cfg))


def wrap_angle(x):
    lib = load()
    return np.array([lib.ref_wrap_angle(float(v)) for v in _f32(x)], dtype=np.float32)


def safe_log(x):
    lib = load()
    return np.array([lib.ref_safe_log(float(v)) for v in _f32(x)], dtype=np.float32)


def mahalanobis(a, b):
    a, b = np.ascontiguousarray(a, GAUSSIAN_DTYPE), np.ascontiguousarray(b, GAUSSIAN_DTYPE)
    return load().ref_mahalanobis(a.ctypes.data, b.ctypes.data)


def hellinger(a, b):
    a, b = np.ascontiguousarray(a, GAUSSIAN_DTYPE), np.ascontiguousarray(b, GAUSSIAN_DTYPE)
    return load().ref_hellinger(a.ctypes.data, b.ctypes.data)


def log_sum_exp(w):
    w = _f32(w)
    return load().ref_log_sum_exp(w.ctypes.data, len(w))


def sum_by_reduction(v):
    x = np.zeros(256, dtype=np.float32)
    x[:len(v)] = v
    return load().ref_sum_by_reduction(x.ctypes.data)


def predict(poses, control, noise):
    """control = (v_encoder, alpha); noise [n][2] {n_alpha, n_encoder} (Ackerman) or [n][3] {ax, ay, atheta} (CV)"""
    p = np.ascontiguousarray(poses, POSE_DTYPE)
    out = np.zeros_like(p)
    c, nz = _f32(control), _f32(noise)
    load().ref_predict(p.ctypes.data, len(p), c.ctypes.data, nz.ctypes.data, out.ctypes.data)
    return out


def in_range(maps, sizes, poses):
    maps = np.ascontiguousarray(maps, GAUSSIAN_DTYPE)
    sizes = np.ascontiguousarray(sizes, np.int32)
    p = np.ascontiguousarray(poses, POSE_DTYPE)
    cls = np.zeros(max(len(maps), 1), dtype=np.int8)
    n_in = np.zeros(len(p), np.int32)
    n_near = np.zeros(len(p), np.int32)
    load().ref_in_range(maps.ctypes.data, sizes.ctypes.data, len(p), p.ctypes.data, cls.ctypes.data, n_in.ctypes.data,
                        n_near.ctypes.data)
    return cls[:len(maps)], n_in, n_near


def birth_device(pose, r, b, label=0):
    p = np.ascontiguousarray(pose, POSE_DTYPE)
    out = np.zeros(1, GAUSSIAN_DTYPE)
    load().ref_birth_device(p.ctypes.data, float(r), float(b), int(label), out.ctypes.data)
    return out[0]


def update_terms(poses, features, n_in, Z):
    """features: concatenated in-range components.  Returns (terms, prune_flags, particle_log_weight_increments)."""
    p = np.ascontiguousarray(poses, POSE_DTYPE)
    f = np.ascontiguousarray(features, GAUSSIAN_DTYPE)
    f = np.concatenate([f, np.zeros(1, GAUSSIAN_DTYPE)])
    n_in = np.ascontiguousarray(n_in, np.int32)
    z = _f32(Z).reshape(len(Z), -1)
    M = z.shape[0]
    n_upd = int(n_in.sum()) * (M + 1) + len(p) * M
    terms = np.zeros(n_upd, GAUSSIAN_DTYPE)
    flags = np.zeros(n_upd, np.int8)
    pw = np.zeros(len(p), np.float32)
    n = load().ref_update_terms(p.ctypes.data, len(p), f.ctypes.data, n_in.ctypes.data, z.ctypes.data, M, z.shape[1],
                                terms.ctypes.data, flags.ctypes.data, pw.ctypes.data)
    assert n == n_upd
    return terms, flags, pw


def merge(cands):
    """cands: list of candidate arrays (one per particle).  Returns list of merged arrays."""
    offs = np.zeros(len(cands) + 1, np.int32)
    offs[1:] = np.cumsum([len(c) for c in cands])
    cat = np.concatenate([np.ascontiguousarray(c, GAUSSIAN_DTYPE) for c in cands] + [np.zeros(1, GAUSSIAN_DTYPE)])
    out = np.zeros(len(cat), GAUSSIAN_DTYPE)
    sizes = np.zeros(len(cands), np.int32)
    load().ref_merge(cat.ctypes.data, offs.ctypes.data, len(cands), out.ctypes.data, sizes.ctypes.data)
    return [out[offs[i]:offs[i] + sizes[i]].copy() for i in range(len(cands))]


def update(poses, sizes, maps, log_weights, Z):
    """whole static-map phdUpdateSynth.  Returns (sizes, maps, log_weights)."""
    p = np.ascontiguousarray(poses, POSE_DTYPE)
    sizes = np.ascontiguousarray(sizes, np.int32)
    maps = np.ascontiguousarray(maps, GAUSSIAN_DTYPE)
    maps = np.concatenate([maps, np.zeros(1, GAUSSIAN_DTYPE)])
    lw = _f32(log_weights).copy()
    z = _f32(Z).reshape(len(Z), -1)
    M = z.shape[0]
    cap = int(sizes.sum()) * (M + 1) + len(p) * M + 1
    out = np.zeros(cap, GAUSSIAN_DTYPE)
    out_sizes = np.zeros(len(p), np.int32)
    n = load().ref_update(p.ctypes.data, len(p), sizes.ctypes.data, maps.ctypes.data, lw.ctypes.data, z.ctypes.data, M,
                          z.shape[1], out_sizes.ctypes.data, out.ctypes.data, cap)
    return out_sizes, out[:n].copy(), lw


def resample(log_weights, uniforms, n_new=-1):
    """uniforms: every randu01() the reference consumes, in order (n_new + 1 values)."""
    lw = _f32(log_weights)
    u = np.ascontiguousarray(uniforms, np.float64)
    nn = len(lw) if n_new < 0 else n_new
    assert len(u) >= nn + 1
    idx = np.zeros(nn, np.int32)
    nlw = np.zeros(nn, np.float32)
    load().ref_resample(lw.ctypes.data, len(lw), n_new, u.ctypes.data, idx.ctypes.data, nlw.ctypes.data)
    return idx, nlw


def recover(log_weights, poses):
    lw = _f32(log_weights)
    p = np.ascontiguousarray(poses, POSE_DTYPE)
    e = np.zeros(1, POSE_DTYPE)
    k = C.c_int()
    load().ref_recover(lw.ctypes.data, p.ctypes.data, len(lw), e.ctypes.data, C.byref(k))
    return e[0], k.value


def cardinality_predict(prior, births):
    """cardinalityPredictKernel (src/phdfilter.cu:867-888): prior [P][N1], births [N1] log-probabilities -> predicted [P][N1].
    set_config() must carry max_cardinality = N1 - 1."""
    lib = load()
    prior = np.ascontiguousarray(prior, np.float32)
    births = np.ascontiguousarray(births, np.float32)
    assert prior.ndim == 2 and births.shape == (prior.shape[1],)
    out = np.zeros_like(prior)
    lib.ref_cardinality_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.ref_cardinality_predict(prior.ctypes.data, births.ctypes.data, prior.shape[0], out.ctypes.data)
    return out


def expected_map(log_weights, sizes, maps, cap=65536):
    """EAP map: recoverSlamState with mapEstimate = 2 -> computeExpectedMap (src/main.cpp:290-316) + the reference's own
    reduceGaussianMixture (src/gm_reduce.cpp:57-134, compiled over the Eigen stand-in of oracle/ref_shim/eigen3)"""
    lib = load()
    lw = _f32(log_weights)
    sizes = np.ascontiguousarray(sizes, np.int32)
    maps = np.ascontiguousarray(maps)
    out = np.zeros(cap, dtype=maps.dtype)
    lib.ref_expected_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.ref_expected_map.restype = C.c_int
    n = lib.ref_expected_map(lw.ctypes.data, sizes.ctypes.data, maps.ctypes.data, len(lw), out.ctypes.data, cap)
    assert n <= cap
    return out[:n].copy()


def write_log(directory, expected_pose, map_est, log_weights, poses, resample_idx, cardinality, t):
    """writeLog (src/main.cpp:848-954) run inside `directory`: appends to state_estimate%05d.log there.
    set_config() decides maxCardinality / filterType / nPredictParticles."""
    lib = load()
    e = np.ascontiguousarray(expected_pose)
    m = np.ascontiguousarray(map_est)
    w = _f32(log_weights)
    p = np.ascontiguousarray(poses)
    ri = np.ascontiguousarray(resample_idx, np.int32)
    cn = _f32(cardinality)
    lib.ref_write_log.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_int]
    rc = lib.ref_write_log(os.fsencode(directory), e.ctypes.data, m.ctypes.data if len(m) else None, len(m), w.ctypes.data,
                           p.ctypes.data, len(w), ri.ctypes.data, cn.ctypes.data, len(cn), int(t))
    assert rc == 0
    return os.path.join(directory, "state_estimate%05d.log" % t)


def load_timestamps(path, cap=1 << 16):
    lib = load()
    out = np.zeros(cap, np.float32)
    lib.ref_load_timestamps.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    n = lib.ref_load_timestamps(os.fsencode(path), out.ctypes.data, cap)
    return out[:n].copy()


def load_controls(path, cap=1 << 16):
    lib = load()
    out = np.zeros((cap, 2), np.float32)
    lib.ref_load_controls.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    n = lib.ref_load_controls(os.fsencode(path), out.ctypes.data, cap)
    return out[:n].copy()


def load_measurements(path, cap_vals=1 << 18, cap_sets=1 << 14):
    """loadMeasurements<measurementSet> (src/main.cpp:221-245) with HEAD's `r b label` parser (:192-208): list of (M_k, 3)"""
    lib = load()
    out = np.zeros((cap_vals, 3), np.float32)
    counts = np.zeros(cap_sets, np.int32)
    lib.ref_load_measurements.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    n = lib.ref_load_measurements(os.fsencode(path), out.ctypes.data, cap_vals, counts.ctypes.data, cap_sets)
    o = np.concatenate([[0], np.cumsum(counts[:n])])
    return [out[o[i]:o[i + 1]].copy() for i in range(n)]


def load_trajectory(path, cap=1 << 16):
    lib = load()
    out = np.zeros(cap, dtype=np.dtype([(k, "f4") for k in ("px", "py", "ptheta", "vx", "vy", "vtheta")]))
    lib.ref_load_trajectory.argtypes = [C.c_char_p, C.c_void_p, C.c_int]
    n = lib.ref_load_trajectory(os.fsencode(path), out.ctypes.data, cap)
    return out[:n].copy()


def plan_events(measurement_times, control_times):
    """the has_timestamps branch of run_synth's loop (src/main.cpp:1188-1230): (z_idx, c_idx, dt) per event"""
    lib = load()
    mt, ct = _f32(measurement_times), _f32(control_times)
    cap = len(mt) + len(ct) + 1
    z, c, dt = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float32)
    lib.ref_plan_events.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    n = lib.ref_plan_events(mt.ctypes.data, len(mt), ct.ctypes.data, len(ct), z.ctypes.data, c.ctypes.data, dt.ctypes.data, cap)
    return z[:n].copy(), c[:n].copy(), dt[:n].copy()


def neff(log_weights):
    lw = _f32(log_weights)
    return load().ref_neff(lw.ctypes.data, len(lw))


def cphd_update(poses, features, n_in, Z, cn_predict, variant="head"):
    """The reference's CPHD multi-object update through the emulator: cphdConstantsKernel, cphdPreUpdateKernel,
    computeEsfKernel, computePsiKernel, cphdUpdateKernel with the .bak wrapper's launch sequence (.bak:2503-2544).
    variant "head": HEAD's commented-out kernels (src/phdfilter.cu:543-607,701-748,1430-1822) made live;
    variant "bak": the live kernels of src/phdfilter.cu.bak:369-415,1058-1478.
    features: concatenated in-range components (n_in[p] per particle); cn_predict [P][256] log predicted cardinality
    (set_config() must carry max_cardinality = 255: the kernels' reductions are 256 wide).
    Returns a dict: detect [sum C][M] and nondetect [sum C] gaussians (weights final), flags, cn_update [P][256],
    ip0 / ip1 [P] (log<Psi0,p>, log<Psi1,p>), ip1d [P][M], esf [P][M+1], esfd [P][M][M]."""
    lib = load()
    p = np.ascontiguousarray(poses, POSE_DTYPE)
    f = np.concatenate([np.ascontiguousarray(features, GAUSSIAN_DTYPE), np.zeros(1, GAUSSIAN_DTYPE)])
    n_in = np.ascontiguousarray(n_in, np.int32)
    z = _f32(Z).reshape(len(Z), -1)
    M, P = z.shape[0], len(p)
    cn = np.ascontiguousarray(cn_predict, np.float32)
    assert cn.shape == (P, 256)
    ntot = int(n_in.sum())
    n_upd = ntot * (M + 1)
    terms = np.zeros(max(n_upd, 1), GAUSSIAN_DTYPE)
    flags = np.zeros(max(n_upd, 1), np.int8)
    cn_up = np.zeros((P, 256), np.float32)
    ip0, ip1, ip1d = np.zeros(P, np.float32), np.zeros(P, np.float32), np.zeros((P, M), np.float32)
    esf, esfd = np.zeros((P, M + 1), np.float32), np.zeros((P, M, M), np.float32)
    lib.ref_cphd_update.restype = C.c_int
    lib.ref_cphd_update.argtypes = [C.c_int] + [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + \
        [C.c_void_p] * 10
    wp = np.zeros((max(ntot, 1), M), np.float32)
    n = lib.ref_cphd_update({"head": 0, "bak": 1}[variant], p.ctypes.data, P, f.ctypes.data, n_in.ctypes.data, z.ctypes.data,
                            M, z.shape[1], cn.ctypes.data, terms.ctypes.data, flags.ctypes.data, cn_up.ctypes.data,
                            ip0.ctypes.data, ip1.ctypes.data, ip1d.ctypes.data, esf.ctypes.data, esfd.ctypes.data, wp.ctypes.data)
    assert n == n_upd, n
    det, nd, dfl, nfl = [], [], [], []
    off = 0
    for c in n_in:                                 # per particle: C*M detection terms (feature-major), then C non-detection
        blk = terms[off * (M + 1):(off + c) * (M + 1)]
        fl = flags[off * (M + 1):(off + c) * (M + 1)]
        det.append(blk[:c * M].reshape(c, M)); dfl.append(fl[:c * M].reshape(c, M))
        nd.append(blk[c * M:]); nfl.append(fl[c * M:])
        off += c
    return dict(detect=np.concatenate(det) if det else np.zeros((0, M), GAUSSIAN_DTYPE), nondetect=np.concatenate(nd),
                detect_flags=np.concatenate(dfl), nondetect_flags=np.concatenate(nfl), cn_update=cn_up, ip0=ip0, ip1=ip1,
                ip1d=ip1d, esf=esf, esfd=esfd, w_partial=wp[:ntot])


# ---- mixed feature model (featureModel = MIXED_MODEL): the reference's own device code through the emulator ----
GAUSSIAN4_DTYPE = np.dtype([("cov", "f4", (16,)), ("mean", "f4", (4,)), ("weight", "f4")])


def mahalanobis4(a, b):
    a = np.ascontiguousarray(a, GAUSSIAN4_DTYPE)
    b = np.ascontiguousarray(b, GAUSSIAN4_DTYPE)
    return float(load().ref_mahalanobis4(a.ctypes.data, b.ctypes.data))


def predict_features4(feats):
    """predictMapKernelMixed: (predicted dynamic features, the jump features the host wrapper discards)"""
    f = np.ascontiguousarray(feats, GAUSSIAN4_DTYPE)
    out = np.zeros(max(len(f), 1), GAUSSIAN4_DTYPE)
    jump = np.zeros(max(len(f), 1), GAUSSIAN_DTYPE)
    load().ref_predict_features4(f.ctypes.data, len(f), out.ctypes.data, jump.ctypes.data)
    return out[:len(f)].copy(), jump[:len(f)].copy()


def mixed_update_terms(pose, s_in, d_in, Z):
    """phdUpdateKernelMixed on one particle: (static terms, dynamic terms, static prune flags, dynamic prune flags, dlogw)"""
    p = np.ascontiguousarray(pose, POSE_DTYPE).reshape(1)
    s_in = np.ascontiguousarray(s_in, GAUSSIAN_DTYPE)
    d_in = np.ascontiguousarray(d_in, GAUSSIAN4_DTYPE)
    z = _f32(Z).reshape(len(Z), -1)
    M = z.shape[0]
    st = np.zeros(len(s_in) * (M + 1) + M, GAUSSIAN_DTYPE)
    dt = np.zeros(len(d_in) * (M + 1) + M, GAUSSIAN4_DTYPE)
    fs = np.zeros(len(st), np.int8)
    fd = np.zeros(len(dt), np.int8)
    sp = np.concatenate([s_in, np.zeros(1, GAUSSIAN_DTYPE)])
    dp = np.concatenate([d_in, np.zeros(1, GAUSSIAN4_DTYPE)])
    pw = load().ref_mixed_update_terms(p.ctypes.data, sp.ctypes.data, len(s_in), dp.ctypes.data, len(d_in), z.ctypes.data, M,
                                       z.shape[1], st.ctypes.data, dt.ctypes.data, fs.ctypes.data, fd.ctypes.data)
    return st, dt, fs, fd, float(pw)


def merge4(cand):
    c = np.ascontiguousarray(cand, GAUSSIAN4_DTYPE)
    out = np.zeros(max(len(c), 1), GAUSSIAN4_DTYPE)
    padded = np.concatenate([c, np.zeros(1, GAUSSIAN4_DTYPE)])
    n = load().ref_merge4(padded.ctypes.data, len(c), out.ctypes.data)
    return out[:n].copy()
