"""ctypes loader for the CPU ORACLE (oracle/libphd_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libphd_oracle.so")

POSE_DTYPE = np.dtype([("px", "f4"), ("py", "f4"), ("ptheta", "f4"), ("vx", "f4"), ("vy", "f4"), ("vtheta", "f4")])
GAUSSIAN_DTYPE = np.dtype([("cov", "f4", (4,)), ("mean", "f4", (2,)), ("weight", "f4")])
GAUSSIAN4_DTYPE = np.dtype([("cov", "f4", (16,)), ("mean", "f4", (4,)), ("weight", "f4")])   # Gaussian4D, 84 B

_lib = None


LIB_USED = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def build_native(timeout=180):
    """-O3 -march=native build for the host this process runs on (the timed CPU baseline).  Returns its path or None."""
    path = os.path.join(_HERE, "libphd_oracle_native.so")
    try:
        subprocess.check_call(["make", "-s", "-B", "-C", _HERE, "native"], timeout=timeout, stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
        C.CDLL(path)          # loads and relocates: a binary this CPU cannot run fails here, not in the timed loop
        return path
    except Exception:
        return None


def load(path=None):
    """path: another build of the same source (build_native()); must be chosen before the first load()."""
    global _lib, LIB_USED
    if _lib is None:
        path = path or os.environ.get("PHD_ORACLE_LIB") or LIB_PATH
        if path == LIB_PATH and not os.path.exists(LIB_PATH):
            build()
        LIB_USED = path
        lib = C.CDLL(path)
        lib.oracle_create.restype = C.c_void_p
        lib.oracle_create.argtypes = [C.c_void_p]
        lib.oracle_mahalanobis.restype = C.c_float
        lib.oracle_hellinger.restype = C.c_float
        lib.oracle_warp_sum.restype = C.c_float
        lib.oracle_update_terms.restype = C.c_size_t
        lib.oracle_update_terms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        for n in ("oracle_destroy", "oracle_n_particles"):
            getattr(lib, n).argtypes = [C.c_void_p]
        for n in ("oracle_set_config", "oracle_get_poses", "oracle_set_poses", "oracle_get_log_weights", "oracle_set_log_weights",
                  "oracle_get_map_sizes", "oracle_get_maps", "oracle_get_resample_idx", "oracle_get_cardinalities",
                  "oracle_set_cardinalities", "oracle_estimate"):
            getattr(lib, n).argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        for n in ("oracle_get_map_sizes_dynamic", "oracle_get_maps_dynamic"):
            getattr(lib, n).argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_set_maps_dynamic.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_mahalanobis4.restype = C.c_float
        lib.oracle_mahalanobis4.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_merge4.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_predict_feature4.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_mixed_terms.restype = C.c_float
        lib.oracle_mixed_terms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                           C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_set_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.oracle_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.oracle_map_estimate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.oracle_resample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.oracle_step_filter.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.oracle_step_resample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.oracle_mahalanobis.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_hellinger.argtypes = [C.c_void_p, C.c_void_p]
        lib.oracle_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_reduce_mixture.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        lib.oracle_warp_sum.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_detmath.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.oracle_philox.argtypes = [C.c_uint] * 6 + [C.c_void_p]
        lib.oracle_esf.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.oracle_cphd_factors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


class Estimate(C.Structure):
    _fields_ = [("px", C.c_float), ("py", C.c_float), ("ptheta", C.c_float), ("vx", C.c_float), ("vy", C.c_float),
                ("vtheta", C.c_float), ("map_particle", C.c_int), ("neff", C.c_float), ("max_log_weight", C.c_float)]

    @property
    def pose(self):
        return np.array([self.px, self.py, self.ptheta, self.vx, self.vy, self.vtheta], dtype=np.float32)


class Oracle(object):
    """CPU restatement of the reference filter; same method names as phdslam_b200.PhdSlam."""

    def __init__(self, cfg, threads=1):
        self.lib = load()
        self.cfg = cfg
        self._h = C.c_void_p(self.lib.oracle_create(C.byref(cfg)))
        self.lib.oracle_set_threads(self._h, threads)

    def close(self):
        if self._h:
            self.lib.oracle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def setDeviceConfig(self, cfg):
        self.lib.oracle_set_config(self._h, C.byref(cfg))
        self.cfg = cfg

    @property
    def n(self):
        return self.lib.oracle_n_particles(self._h)

    n_local = n

    def phdPredict(self, control=None, draws=None):
        c = None if control is None else np.ascontiguousarray(control, dtype=np.float32)
        d = None if draws is None else np.ascontiguousarray(draws, dtype=np.float64)
        self.lib.oracle_predict(self._h, _ptr(c), _ptr(d))

    def phdUpdateSynth(self, Z):
        z = np.ascontiguousarray(Z, dtype=np.float32)
        if z.size == 0:
            return
        z = z.reshape(len(z), -1)
        self.lib.oracle_update(self._h, z.ctypes.data, z.shape[0], z.shape[1])

    def recoverSlamState(self):
        e = Estimate()
        self.lib.oracle_estimate(self._h, C.byref(e))
        return e

    def resampleParticles(self, uniforms=None, literal=False, n_new=-1):
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        anc = np.empty(self.n if n_new < 0 else n_new, dtype=np.int32)
        self.lib.oracle_resample(self._h, n_new, _ptr(u), int(literal), anc.ctypes.data)
        return anc

    def step(self, step_index, control, Z):
        c = None if control is None else np.ascontiguousarray(control, dtype=np.float32)
        z = np.ascontiguousarray(Z, dtype=np.float32)
        M = 0 if z.size == 0 else len(z)
        fields = 2 if M == 0 else z.reshape(M, -1).shape[1]
        e, res = Estimate(), C.c_int()
        self.lib.oracle_step(self._h, step_index, _ptr(c), z.ctypes.data if M else None, M, fields, C.byref(e), C.byref(res))
        return e, bool(res.value)

    def step_filter(self, step_index, control, Z):
        c = None if control is None else np.ascontiguousarray(control, dtype=np.float32)
        z = np.ascontiguousarray(Z, dtype=np.float32)
        M = 0 if z.size == 0 else len(z)
        fields = 2 if M == 0 else z.reshape(M, -1).shape[1]
        e = Estimate()
        self.lib.oracle_step_filter(self._h, step_index, _ptr(c), z.ctypes.data if M else None, M, fields, C.byref(e))
        return e

    def step_resample(self, M, est):
        res = C.c_int()
        self.lib.oracle_step_resample(self._h, int(M), C.byref(est), C.byref(res))
        return bool(res.value)

    def map_estimate(self, which=1, cap=65536):
        out = np.zeros(cap, dtype=GAUSSIAN_DTYPE)
        n = self.lib.oracle_map_estimate(self._h, which, out.ctypes.data, cap)
        return out[:n].copy()

    def update_terms(self, Z, want_terms=True):
        z = np.ascontiguousarray(Z, dtype=np.float32)
        z = z.reshape(len(z), -1)
        n, M = self.n, z.shape[0]
        nin = np.zeros(n, dtype=np.int32)
        dlw = np.zeros(n, dtype=np.float32)
        cap = int(self.map_sizes.sum()) * (M + 1) + M * n
        terms = np.zeros(cap if want_terms else 0, dtype=GAUSSIAN_DTYPE)
        tot = self.lib.oracle_update_terms(self._h, z.ctypes.data, M, z.shape[1], _ptr(terms) if want_terms else None, cap,
                                           nin.ctypes.data, dlw.ctypes.data)
        return (terms[:tot] if want_terms else None), nin, dlw

    @property
    def poses(self):
        out = np.zeros(self.n, dtype=POSE_DTYPE)
        self.lib.oracle_get_poses(self._h, out.ctypes.data)
        return out

    @poses.setter
    def poses(self, v):
        v = np.ascontiguousarray(v, dtype=POSE_DTYPE)
        assert len(v) == self.n
        self.lib.oracle_set_poses(self._h, v.ctypes.data)

    @property
    def log_weights(self):
        out = np.zeros(self.n, dtype=np.float32)
        self.lib.oracle_get_log_weights(self._h, out.ctypes.data)
        return out

    @log_weights.setter
    def log_weights(self, v):
        v = np.ascontiguousarray(v, dtype=np.float32)
        assert len(v) == self.n
        self.lib.oracle_set_log_weights(self._h, v.ctypes.data)

    @property
    def map_sizes(self):
        out = np.zeros(self.n, dtype=np.int32)
        self.lib.oracle_get_map_sizes(self._h, out.ctypes.data)
        return out

    def get_maps(self):
        sizes = self.map_sizes
        out = np.zeros(int(sizes.sum()), dtype=GAUSSIAN_DTYPE)
        self.lib.oracle_get_maps(self._h, out.ctypes.data)
        return sizes, out

    def set_maps(self, sizes, maps):
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        maps = np.ascontiguousarray(maps, dtype=GAUSSIAN_DTYPE)
        assert len(sizes) == self.n and int(sizes.sum()) == len(maps)
        self.lib.oracle_set_maps(self._h, sizes.ctypes.data, maps.ctypes.data)

    @property
    def map_sizes_dynamic(self):
        out = np.zeros(self.n, dtype=np.int32)
        self.lib.oracle_get_map_sizes_dynamic(self._h, out.ctypes.data)
        return out

    def get_maps_dynamic(self):
        sizes = self.map_sizes_dynamic
        out = np.zeros(int(sizes.sum()), dtype=GAUSSIAN4_DTYPE)
        self.lib.oracle_get_maps_dynamic(self._h, out.ctypes.data)
        return sizes, out

    def set_maps_dynamic(self, sizes, maps):
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        maps = np.ascontiguousarray(maps, dtype=GAUSSIAN4_DTYPE)
        assert len(sizes) == self.n and int(sizes.sum()) == len(maps)
        self.lib.oracle_set_maps_dynamic(self._h, sizes.ctypes.data, maps.ctypes.data)

    @property
    def resample_idx(self):
        out = np.zeros(self.n, dtype=np.int32)
        self.lib.oracle_get_resample_idx(self._h, out.ctypes.data)
        return out

    @property
    def cardinalities(self):
        out = np.zeros((self.n, self.cfg.max_cardinality + 1), dtype=np.float32)
        self.lib.oracle_get_cardinalities(self._h, out.ctypes.data)
        return out

    @cardinalities.setter
    def cardinalities(self, v):
        v = np.ascontiguousarray(v, dtype=np.float32)
        assert v.shape == (self.n, self.cfg.max_cardinality + 1)
        self.lib.oracle_set_cardinalities(self._h, v.ctypes.data)


def mahalanobis4(a, b):
    a = np.ascontiguousarray(a, GAUSSIAN4_DTYPE)
    b = np.ascontiguousarray(b, GAUSSIAN4_DTYPE)
    return float(load().oracle_mahalanobis4(a.ctypes.data, b.ctypes.data))


def merge4(cfg, cand):
    cand = np.ascontiguousarray(cand, GAUSSIAN4_DTYPE)
    out = np.zeros(max(len(cand), 1), GAUSSIAN4_DTYPE)
    n = load().oracle_merge4(C.byref(cfg), cand.ctypes.data, len(cand), out.ctypes.data)
    return out[:n].copy()


def predict_features4(cfg, feats):
    feats = np.ascontiguousarray(feats, GAUSSIAN4_DTYPE)
    out = np.zeros(len(feats), GAUSSIAN4_DTYPE)
    for i in range(len(feats)):
        load().oracle_predict_feature4(C.byref(cfg), feats[i:i + 1].ctypes.data, out[i:i + 1].ctypes.data)
    return out


def mixed_terms(cfg, pose, smap, dmap, Z):
    """Dense update terms of ONE particle of the mixed feature model: (static terms, dynamic terms, log-weight increment)."""
    pose = np.ascontiguousarray(pose, POSE_DTYPE).reshape(1)
    smap = np.ascontiguousarray(smap, GAUSSIAN_DTYPE)
    dmap = np.ascontiguousarray(dmap, GAUSSIAN4_DTYPE)
    z = np.ascontiguousarray(Z, np.float32).reshape(len(Z), -1)
    M = z.shape[0]
    st = np.zeros(len(smap) * (M + 1) + M, GAUSSIAN_DTYPE)
    dt = np.zeros(len(dmap) * (M + 1) + M, GAUSSIAN4_DTYPE)
    ns, nd = C.c_int(), C.c_int()
    dl = load().oracle_mixed_terms(C.byref(cfg), pose.ctypes.data, smap.ctypes.data, len(smap), dmap.ctypes.data, len(dmap),
                                   z.ctypes.data, M, z.shape[1], st.ctypes.data, C.byref(ns), dt.ctypes.data, C.byref(nd))
    return st[:ns.value].copy(), dt[:nd.value].copy(), float(dl)


def cphd_factors(cfg, w, pd, S, prior):
    """CPHD multi-object terms of one particle: returns (D[M], ND, inc, card_out[N1])"""
    lib = load()
    w = np.ascontiguousarray(w, np.float32)
    pd = np.ascontiguousarray(pd, np.float32)
    S = np.ascontiguousarray(S, np.float32)
    prior = np.ascontiguousarray(prior, np.float32)
    D = np.zeros(len(S), np.float32)
    card = np.zeros(len(prior), np.float32)
    nd, inc = C.c_float(), C.c_float()
    lib.oracle_cphd_factors(C.byref(cfg), w.ctypes.data, pd.ctypes.data, len(w), S.ctypes.data, len(S), prior.ctypes.data,
                            len(prior), D.ctypes.data, C.byref(nd), C.byref(inc), card.ctypes.data)
    return D, nd.value, inc.value, card


def cphd_predict_cardinality(cfg, prior, M):
    """(pb[M+1], pm[N1]): binomial birth cardinality and predicted cardinality of one particle (log domain)"""
    lib = load()
    prior = np.ascontiguousarray(prior, np.float32)
    pb = np.zeros(M + 1, np.float32)
    pm = np.zeros(len(prior), np.float32)
    lib.oracle_cphd_predict_cardinality.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.oracle_cphd_predict_cardinality(C.byref(cfg), prior.ctypes.data, len(prior), int(M), pb.ctypes.data, pm.ctypes.data)
    return pb, pm


def detmath(fn, x, y=None):
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y if y is not None else np.zeros_like(x), dtype=np.float32)
    out = np.zeros_like(x)
    out2 = np.zeros_like(x)
    code = {"exp": 0, "log": 1, "atan2": 2, "sincos": 3, "wrap": 4, "tan": 5, "safe_log": 6}[fn]
    lib.oracle_detmath(code, x.ctypes.data, y.ctypes.data, out.ctypes.data, out2.ctypes.data, len(x))
    return (out, out2) if fn == "sincos" else out
