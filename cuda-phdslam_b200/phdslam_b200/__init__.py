"""phdslam_b200 -- Python host-side mirror of the reference's filter interface over libphdslam.so.

The product is the C-ABI shared library (include/phdslam.h, CUDA kernels for sm_100a).  This module is
a thin ctypes binding used by the tests, bench.py and anyone driving the filter from Python; it mirrors
the reference's names (src/phdfilter.h:10-34, src/main.cpp):

    phdPredict(particles, control)      -> PhdSlam.phdPredict(control)
    phdUpdateSynth(particles, Z)        -> PhdSlam.phdUpdateSynth(Z)
    recoverSlamState(particles, ...)    -> PhdSlam.recoverSlamState()
    resampleParticles(particles, N)     -> PhdSlam.resampleParticles()
    setDeviceConfig(config)             -> PhdSlam.setDeviceConfig(cfg)

There is NO CPU fallback: constructing PhdSlam without the built extension or without a GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libphdslam.so")

POSE_DTYPE = np.dtype([("px", "f4"), ("py", "f4"), ("ptheta", "f4"), ("vx", "f4"), ("vy", "f4"), ("vtheta", "f4")])
GAUSSIAN_DTYPE = np.dtype([("cov", "f4", (4,)), ("mean", "f4", (2,)), ("weight", "f4")])
GAUSSIAN4_DTYPE = np.dtype([("cov", "f4", (16,)), ("mean", "f4", (4,)), ("weight", "f4")])   # Gaussian4D (src/slamtypes.h:135-139)
assert POSE_DTYPE.itemsize == 24 and GAUSSIAN_DTYPE.itemsize == 28


class Config(C.Structure):
    """ctypes image of phdslam_config_t (include/phdslam.h); keys of cfg/config.cfg."""

    _fields_ = [
        ("x0", C.c_float), ("y0", C.c_float), ("yaw0", C.c_float), ("vx0", C.c_float), ("vy0", C.c_float),
        ("vyaw0", C.c_float),
        ("motion_type", C.c_int),
        ("ax", C.c_float), ("ay", C.c_float), ("ayaw", C.c_float),
        ("dt", C.c_float),
        ("min_range", C.c_float), ("max_range", C.c_float), ("max_bearing", C.c_float),
        ("std_range", C.c_float), ("std_bearing", C.c_float),
        ("clutter_rate", C.c_float), ("clutter_density", C.c_float), ("pd", C.c_float),
        ("n_particles", C.c_int), ("n_predict_particles", C.c_int), ("subdivide_predict", C.c_int),
        ("resample_threshold", C.c_float), ("birth_weight", C.c_float), ("birth_noise_factor", C.c_float),
        ("min_separation", C.c_float), ("min_feature_weight", C.c_float),
        ("particle_weighting", C.c_int), ("distance_metric", C.c_int), ("max_cardinality", C.c_int),
        ("filter_type", C.c_int), ("map_estimate", C.c_int), ("feature_model", C.c_int),
        ("l", C.c_float), ("h", C.c_float), ("a", C.c_float), ("b", C.c_float), ("std_encoder", C.c_float),
        ("std_alpha", C.c_float),
        ("labeled_measurements", C.c_int), ("follow_trajectory", C.c_int), ("max_steps", C.c_int),
        ("n_steps", C.c_int),
        ("data_directory", C.c_char * 1024),
        ("measurement_fields", C.c_int), ("max_components", C.c_int), ("resample_mode", C.c_int),
        ("log_layout", C.c_int),
        ("seed", C.c_ulonglong),
        ("update_mode", C.c_int),
        ("update_buffer_bytes", C.c_ulonglong),
        ("ps", C.c_float), ("tau", C.c_float), ("beta", C.c_float),
        ("std_ax_features", C.c_float), ("std_ay_features", C.c_float),
        ("cov_vx_birth", C.c_float), ("cov_vy_birth", C.c_float),
        ("max_components_dynamic", C.c_int),
    ]

    def set(self, **kw):
        """Set options by cfg-file key (goes through the library's parser so derived fields stay in sync)."""
        lib = load_library()
        for k, v in kw.items():
            if isinstance(v, str):
                sv = v
            elif isinstance(v, (bool, int, np.integer)):
                sv = str(int(v))
            else:
                sv = repr(float(v))
            rc = lib.phdslam_config_set(C.byref(self), k.encode(), sv.encode())
            if rc != 0:
                raise KeyError("unknown config key %r" % k)
        return self


class Estimate(C.Structure):
    _fields_ = [("px", C.c_float), ("py", C.c_float), ("ptheta", C.c_float), ("vx", C.c_float), ("vy", C.c_float),
                ("vtheta", C.c_float), ("map_particle", C.c_int), ("neff", C.c_float), ("max_log_weight", C.c_float)]

    @property
    def pose(self):
        return np.array([self.px, self.py, self.ptheta, self.vx, self.vy, self.vtheta], dtype=np.float32)


class Timings(C.Structure):
    _fields_ = [("predict_ms", C.c_float), ("update_ms", C.c_float), ("merge_ms", C.c_float), ("weights_ms", C.c_float),
                ("estimate_ms", C.c_float), ("resample_ms", C.c_float), ("launches", C.c_ulonglong),
                ("migrated_in", C.c_ulonglong), ("h2d_bytes", C.c_ulonglong), ("d2h_bytes", C.c_ulonglong),
                ("dynamic_ms", C.c_float)]


# every symbol include/phdslam.h declares (tests/test_abi.py checks the built library exports all of them)
ABI_SYMBOLS = [
    "phdslam_config_defaults", "phdslam_config_load", "phdslam_config_set", "phdslam_last_error", "phdslam_version",
    "phdslam_create", "phdslam_destroy", "phdslam_set_config", "phdslam_get_config", "phdslam_dist_unique_id",
    "phdslam_dist_init", "phdslam_plan_migration", "phdslam_resample_threshold", "phdslam_predict", "phdslam_update", "phdslam_estimate", "phdslam_map_estimate",
    "phdslam_resample", "phdslam_step", "phdslam_step_filter", "phdslam_step_resample", "phdslam_set_particle_count",
    "phdslam_particle_capacity", "phdslam_particle_checksums", "phdslam_import_tiled", "phdslam_n_local", "phdslam_local_offset", "phdslam_get_poses",
    "phdslam_set_poses", "phdslam_get_log_weights", "phdslam_set_log_weights", "phdslam_get_map_sizes",
    "phdslam_get_maps", "phdslam_set_maps", "phdslam_get_resample_idx", "phdslam_get_cardinalities",
    "phdslam_set_cardinalities", "phdslam_update_terms", "phdslam_get_timings", "phdslam_stream",
    "phdslam_synchronize", "phdslam_set_overlap", "phdslam_dist_p2p", "phdslam_snapshot", "phdslam_restore", "phdslam_load_measurements",
    "phdslam_load_controls", "phdslam_load_timestamps", "phdslam_load_trajectory", "phdslam_plan_events", "phdslam_free",
    "phdslam_write_log", "phdslam_write_log_mixed",
    "phdslam_get_map_sizes_dynamic", "phdslam_get_maps_dynamic", "phdslam_set_maps_dynamic", "phdslam_map_estimate_dynamic",
]

_lib = None


def load_library(path=None):
    """Load libphdslam.so.  Raises if the extension has not been built: there is no fallback path."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError("libphdslam.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C cuda-phdslam_b200`; there is no CPU fallback" % p)
    lib = C.CDLL(p)
    lib.phdslam_last_error.restype = C.c_char_p
    lib.phdslam_version.restype = C.c_char_p
    lib.phdslam_stream.restype = C.c_void_p
    lib.phdslam_stream.argtypes = [C.c_void_p]
    lib.phdslam_config_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    lib.phdslam_config_load.argtypes = [C.c_char_p, C.c_void_p]
    lib.phdslam_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
    lib.phdslam_destroy.argtypes = [C.c_void_p]
    lib.phdslam_destroy.restype = None
    for name in ("phdslam_set_config", "phdslam_get_config"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p]
    lib.phdslam_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.phdslam_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    lib.phdslam_estimate.argtypes = [C.c_void_p, C.c_void_p]
    lib.phdslam_map_estimate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.phdslam_resample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.phdslam_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.phdslam_step_filter.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.phdslam_step_resample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.phdslam_set_particle_count.argtypes = [C.c_void_p, C.c_int]
    lib.phdslam_particle_capacity.argtypes = [C.c_void_p]
    lib.phdslam_particle_checksums.argtypes = [C.c_void_p, C.c_void_p]
    lib.phdslam_import_tiled.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    for name in ("phdslam_n_local", "phdslam_local_offset", "phdslam_synchronize", "phdslam_set_overlap", "phdslam_snapshot", "phdslam_restore"):
        getattr(lib, name).argtypes = [C.c_void_p]
    for name in ("phdslam_get_poses", "phdslam_set_poses", "phdslam_get_log_weights", "phdslam_set_log_weights",
                 "phdslam_get_map_sizes", "phdslam_get_resample_idx", "phdslam_get_cardinalities",
                 "phdslam_set_cardinalities", "phdslam_get_timings"):
        getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p]
    lib.phdslam_set_overlap.argtypes = [C.c_void_p, C.c_int]
    lib.phdslam_get_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.phdslam_set_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.phdslam_get_map_sizes_dynamic.argtypes = [C.c_void_p, C.c_void_p]
    lib.phdslam_get_maps_dynamic.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.phdslam_set_maps_dynamic.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.phdslam_map_estimate_dynamic.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.phdslam_update_terms.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    lib.phdslam_dist_unique_id.argtypes = [C.c_void_p]
    lib.phdslam_dist_init.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.phdslam_plan_migration.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint, C.c_ulonglong, C.c_void_p]
    lib.phdslam_resample_threshold.argtypes = [C.c_int, C.c_int, C.c_ulonglong, C.c_void_p, C.c_int, C.c_uint, C.c_ulonglong]
    lib.phdslam_resample_threshold.restype = C.c_ulonglong
    lib.phdslam_load_measurements.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_float)),
                                              C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_int)]
    lib.phdslam_load_controls.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_int)]
    lib.phdslam_load_timestamps.argtypes = [C.c_char_p, C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int)]
    lib.phdslam_load_trajectory.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]
    lib.phdslam_plan_events.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.phdslam_free.argtypes = [C.c_void_p]
    lib.phdslam_free.restype = None
    lib.phdslam_write_log.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    if path is None:
        _lib = lib
    return lib


class PhdSlamError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, "phdslam status %d: %s" % (code, msg))
        self.code = code


def _check(rc):
    if rc != 0:
        raise PhdSlamError(rc, load_library().phdslam_last_error().decode(errors="replace"))


def default_config(**kw):
    cfg = Config()
    load_library().phdslam_config_defaults(C.byref(cfg))
    return cfg.set(**kw) if kw else cfg


def load_config(path):
    """loadConfig (reference src/main.cpp:956-1073)."""
    cfg = Config()
    _check(load_library().phdslam_config_load(os.fsencode(path), C.byref(cfg)))
    return cfg


def load_measurements(path, fields=2, has_header=1):
    """loadMeasurements (src/main.cpp:220-244): list of (M_k, fields) float32 arrays, one per time step."""
    lib = load_library()
    data, offs, n = C.POINTER(C.c_float)(), C.POINTER(C.c_int)(), C.c_int()
    _check(lib.phdslam_load_measurements(os.fsencode(path), fields, has_header, C.byref(data), C.byref(offs), C.byref(n)))
    o = np.ctypeslib.as_array(offs, shape=(n.value + 1,)).copy()
    tot = int(o[-1])
    d = np.ctypeslib.as_array(data, shape=(max(tot * fields, 1),)).copy()[: tot * fields].reshape(tot, fields)
    lib.phdslam_free(data)
    lib.phdslam_free(offs)
    return [d[o[i]:o[i + 1]].copy() for i in range(n.value)]


def load_controls(path):
    """loadControls (src/main.cpp:169-190): (n, 2) array of {v_encoder, alpha}."""
    lib = load_library()
    data, n = C.POINTER(C.c_float)(), C.c_int()
    _check(lib.phdslam_load_controls(os.fsencode(path), C.byref(data), C.byref(n)))
    d = np.ctypeslib.as_array(data, shape=(max(n.value * 2, 1),)).copy()[: n.value * 2].reshape(n.value, 2)
    lib.phdslam_free(data)
    return d


def load_timestamps(path):
    """loadTimestamps (src/main.cpp:147-167); a missing file gives an empty array (run without time stamps)."""
    lib = load_library()
    data, n = C.POINTER(C.c_double)(), C.c_int()
    _check(lib.phdslam_load_timestamps(os.fsencode(path), C.byref(data), C.byref(n)))
    d = np.ctypeslib.as_array(data, shape=(max(n.value, 1),)).copy()[: n.value] if n.value else np.zeros(0)
    lib.phdslam_free(data)
    return d


def load_trajectory(path):
    """loadTrajectory (src/main.cpp:247-264): POSE_DTYPE array."""
    lib = load_library()
    data, n = C.c_void_p(), C.c_int()
    _check(lib.phdslam_load_trajectory(os.fsencode(path), C.byref(data), C.byref(n)))
    buf = (C.c_char * (max(n.value, 1) * POSE_DTYPE.itemsize)).from_address(data.value)
    d = np.frombuffer(buf, dtype=POSE_DTYPE, count=n.value).copy()
    lib.phdslam_free(data)
    return d


EVENT_DTYPE = np.dtype([("z_idx", "i4"), ("c_idx", "i4"), ("dt", "f4")])


def plan_events(measurement_times, control_times):
    """The per-step input schedule of run_synth for time-stamped streams (src/main.cpp:1187-1230): EVENT_DTYPE array."""
    zt = np.ascontiguousarray(measurement_times, dtype=np.float64)
    ct = np.ascontiguousarray(control_times, dtype=np.float64)
    ev = np.zeros(len(zt) + len(ct) + 1, dtype=EVENT_DTYPE)
    n = load_library().phdslam_plan_events(zt.ctypes.data, len(zt), ct.ctypes.data, len(ct), ev.ctypes.data, len(ev))
    if n < 0:
        _check(n)
    return ev[:n].copy()


def write_log(path, layout, expected_pose, map_est, log_weights, poses, resample_idx=None, cardinality=None, n_card=1,
              filter_type=0, map_dynamic=None):
    """writeLog (src/main.cpp:848-954) / README 5-line layout.  map_dynamic: the dynamic map estimate of the mixed feature
    model (line 3 of the 7-line layout)."""
    e = np.ascontiguousarray(expected_pose, dtype=np.float32)
    m = np.ascontiguousarray(map_est, dtype=GAUSSIAN_DTYPE)
    d = np.zeros(0, GAUSSIAN4_DTYPE) if map_dynamic is None else np.ascontiguousarray(map_dynamic, dtype=GAUSSIAN4_DTYPE)
    w = np.ascontiguousarray(log_weights, dtype=np.float32)
    p = np.ascontiguousarray(poses, dtype=POSE_DTYPE)
    ri = None if resample_idx is None else np.ascontiguousarray(resample_idx, dtype=np.int32)
    cd = None if cardinality is None else np.ascontiguousarray(cardinality, dtype=np.float32)
    lib = load_library()
    lib.phdslam_write_log_mixed.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    _check(lib.phdslam_write_log_mixed(
        os.fsencode(path), layout, e.ctypes.data, m.ctypes.data if len(m) else None, len(m), d.ctypes.data if len(d) else None,
        len(d), w.ctypes.data, p.ctypes.data, len(w), None if ri is None else ri.ctypes.data,
        None if cd is None else cd.ctypes.data, n_card, filter_type))


def _ptr(a):
    return None if a is None else a.ctypes.data


def dist_unique_id():
    buf = np.zeros(128, dtype=np.uint8)
    _check(load_library().phdslam_dist_unique_id(buf.ctypes.data))
    return buf.tobytes()


def plan_migration(totals, n_new, uniforms=None, resample_mode=0, call=0, seed=0):
    """Host-side planning of the global resampling exchange (phdslam_plan_migration): bounds[r] = first offspring
    whose ancestor lives on a rank >= r."""
    t = np.ascontiguousarray(totals, dtype=np.uint64)
    u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
    bounds = np.zeros(len(t) + 1, dtype=np.int32)
    _check(load_library().phdslam_plan_migration(len(t), t.ctypes.data, n_new, _ptr(u), resample_mode, call, seed, bounds.ctypes.data))
    return bounds


def resample_threshold(j, n_new, total, uniforms=None, resample_mode=0, call=0, seed=0):
    u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
    return int(load_library().phdslam_resample_threshold(j, n_new, total, _ptr(u), resample_mode, call, seed))


class PhdSlam(object):
    """One filter instance on one GPU (opaque phdslam_t handle with persistent device state)."""

    def __init__(self, cfg, device=0):
        self.lib = load_library()
        self.cfg = cfg
        self._h = C.c_void_p()
        _check(self.lib.phdslam_create(C.byref(cfg), device, C.byref(self._h)))

    def close(self):
        if self._h:
            self.lib.phdslam_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def dist_init(self, rank, world, unique_id=None):
        """Shard the particles over `world` ranks (one process per GPU).  unique_id: 128-byte ncclUniqueId made by
        rank 0 with dist_unique_id(); when None it is created on rank 0 and broadcast through torch.distributed."""
        if world == 1:
            return
        if unique_id is None:
            import torch
            import torch.distributed as dist
            buf = np.zeros(128, dtype=np.uint8)
            if rank == 0:
                _check(self.lib.phdslam_dist_unique_id(buf.ctypes.data))
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            t = torch.from_numpy(buf).to(dev)
            dist.broadcast(t, src=0)
            unique_id = t.cpu().numpy().tobytes()
        idb = (C.c_char * 128).from_buffer_copy(unique_id)
        _check(self.lib.phdslam_dist_init(self._h, rank, world, idb))

    @property
    def local_offset(self):
        return self.lib.phdslam_local_offset(self._h)

    # ---- reference-named operations -------------------------------------------------------------
    def setDeviceConfig(self, cfg):
        _check(self.lib.phdslam_set_config(self._h, C.byref(cfg)))
        self.cfg = cfg

    def phdPredict(self, control=None, draws=None):
        c = None if control is None else np.ascontiguousarray(control, dtype=np.float32)
        d = None if draws is None else np.ascontiguousarray(draws, dtype=np.float64)
        _check(self.lib.phdslam_predict(self._h, _ptr(c), _ptr(d)))

    def phdUpdateSynth(self, Z):
        z = np.ascontiguousarray(Z, dtype=np.float32)
        if z.size == 0:
            return
        z = z.reshape(len(z), -1)
        _check(self.lib.phdslam_update(self._h, z.ctypes.data, z.shape[0], z.shape[1]))

    def recoverSlamState(self):
        e = Estimate()
        _check(self.lib.phdslam_estimate(self._h, C.byref(e)))
        return e

    def resampleParticles(self, uniforms=None, n_new=-1):
        """uniforms: n_new+1 injected draws (the same array on every rank), or None for the counter-based RNG.
        n_new != current count (single GPU only): the down-sampling of run_synth after "shotgun" predictions."""
        n = self.n_local if n_new < 0 else n_new
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        anc = np.empty(n, dtype=np.int32)
        _check(self.lib.phdslam_resample(self._h, n_new, _ptr(u), anc.ctypes.data))
        return anc

    def step(self, step_index, control, Z):
        c = None if control is None else np.ascontiguousarray(control, dtype=np.float32)
        z = np.ascontiguousarray(Z, dtype=np.float32)
        M = 0 if z.size == 0 else len(z)
        fields = 2 if M == 0 else z.reshape(M, -1).shape[1]
        e, res = Estimate(), C.c_int()
        _check(self.lib.phdslam_step(self._h, step_index, _ptr(c), z.ctypes.data if M else None, M, fields, C.byref(e),
                                     C.byref(res)))
        return e, bool(res.value)

    def save_particles(self, directory, t, map_estimates=True):
        """The reference's per-step particle dump (writeParticlesMat, src/main.cpp:594-713: `particles%05d.mat` with the
        struct fields states, weights, resample_idx, maps_static {weights, means, covs}, max_map_static, exp_map_static)
        as `particles%05d.npz`.  Maps are concatenated particle after particle; `map_offsets[i]:map_offsets[i+1]` is
        particle i's map.  (The reference never writes the covariances -- its `ptr_covs` stays NULL, :553 -- this does;
        with feature_model = 2 the dynamic maps follow as maps_dynamic.* / max_map_dynamic.* with `map_offsets_dynamic`;
        exp_map_dynamic is not computed.)"""
        sizes, maps = self.get_maps()
        p = self.poses
        out = {"states": np.stack([p[f] for f in POSE_DTYPE.names], 1), "weights": self.log_weights,
               "resample_idx": self.resample_idx, "map_offsets": np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64),
               "maps_static.weights": maps["weight"], "maps_static.means": maps["mean"], "maps_static.covs": maps["cov"]}
        if map_estimates:
            for which, name in ((1, "max_map_static"), (2, "exp_map_static")):
                if self.cfg.map_estimate & which:
                    m = self.map_estimate(which, cap=1 << 16)
                    out[name + ".weights"], out[name + ".means"], out[name + ".covs"] = m["weight"], m["mean"], m["cov"]
        if self.cfg.filter_type == 1:
            out["cardinalities"] = self.cardinalities
        if self.cfg.feature_model == 2:
            dsz, dm = self.get_maps_dynamic()
            out["map_offsets_dynamic"] = np.concatenate([[0], np.cumsum(dsz)]).astype(np.int64)
            out["maps_dynamic.weights"], out["maps_dynamic.means"], out["maps_dynamic.covs"] = dm["weight"], dm["mean"], dm["cov"]
            if map_estimates and (self.cfg.map_estimate & 1):
                m = self.map_estimate_dynamic()
                out["max_map_dynamic.weights"], out["max_map_dynamic.means"], out["max_map_dynamic.covs"] = (
                    m["weight"], m["mean"], m["cov"])
        path = os.path.join(directory, "particles%05d.npz" % int(t))
        np.savez_compressed(path, **out)
        return path

    def particle_checksums(self):
        """one uint64 per local particle over pose, map and cardinality: an offspring's equals its ancestor's"""
        out = np.zeros(self.n_local, dtype=np.uint64)
        _check(self.lib.phdslam_particle_checksums(self._h, out.ctypes.data))
        return out

    def import_tiled(self, sc):
        """loads the particles of scene `sc` and fills the rest of the local particles with copies of them (on the device)"""
        p = np.ascontiguousarray(sc["poses"], dtype=POSE_DTYPE)
        w = np.ascontiguousarray(sc["log_weights"], dtype=np.float32)
        sizes = np.ascontiguousarray(sc["sizes"], dtype=np.int32)
        maps = np.ascontiguousarray(sc["maps"], dtype=GAUSSIAN_DTYPE)
        assert len(p) == len(w) == len(sizes) <= self.n_local and int(sizes.sum()) == len(maps)
        _check(self.lib.phdslam_import_tiled(self._h, len(p), p.ctypes.data, w.ctypes.data, sizes.ctypes.data, maps.ctypes.data))

    def step_filter(self, step_index, control, Z):
        """predict + update + estimate: the state run_synth looks at (and logs) is the one after this half"""
        c = None if control is None else np.ascontiguousarray(control, dtype=np.float32)
        z = np.ascontiguousarray(Z, dtype=np.float32)
        M = 0 if z.size == 0 else len(z)
        fields = 2 if M == 0 else z.reshape(M, -1).shape[1]
        e = Estimate()
        _check(self.lib.phdslam_step_filter(self._h, step_index, _ptr(c), z.ctypes.data if M else None, M, fields, C.byref(e)))
        return e

    def step_resample(self, M, est):
        res = C.c_int()
        _check(self.lib.phdslam_step_resample(self._h, int(M), C.byref(est), C.byref(res)))
        return bool(res.value)

    def map_estimate(self, which=1, cap=4096):
        out = np.zeros(cap, dtype=GAUSSIAN_DTYPE)
        n = C.c_int()
        _check(self.lib.phdslam_map_estimate(self._h, which, out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    def update_terms(self, Z, want_terms=True):
        """Dense GM-PHD update terms in the reference's features_update order (no state change)."""
        z = np.ascontiguousarray(Z, dtype=np.float32)
        z = z.reshape(len(z), -1)
        n = self.n_local
        nin = np.zeros(n, dtype=np.int32)
        dlw = np.zeros(n, dtype=np.float32)
        M = z.shape[0]
        if want_terms:
            sizes = self.map_sizes
            cap = int(sizes.sum()) * (M + 1) + M * n
            terms = np.zeros(cap, dtype=GAUSSIAN_DTYPE)
            _check(self.lib.phdslam_update_terms(self._h, z.ctypes.data, M, z.shape[1], terms.ctypes.data, cap, nin.ctypes.data,
                                                 dlw.ctypes.data))
            tot = int((nin.astype(np.int64) * (M + 1) + M).sum())
            return terms[:tot], nin, dlw
        _check(self.lib.phdslam_update_terms(self._h, z.ctypes.data, M, z.shape[1], None, 0, nin.ctypes.data, dlw.ctypes.data))
        return None, nin, dlw

    # ---- state ----------------------------------------------------------------------------------
    @property
    def n_local(self):
        return self.lib.phdslam_n_local(self._h)

    @property
    def poses(self):
        out = np.zeros(self.n_local, dtype=POSE_DTYPE)
        _check(self.lib.phdslam_get_poses(self._h, out.ctypes.data))
        return out

    @poses.setter
    def poses(self, v):
        v = np.ascontiguousarray(v, dtype=POSE_DTYPE)
        assert len(v) == self.n_local
        _check(self.lib.phdslam_set_poses(self._h, v.ctypes.data))

    @property
    def log_weights(self):
        out = np.zeros(self.n_local, dtype=np.float32)
        _check(self.lib.phdslam_get_log_weights(self._h, out.ctypes.data))
        return out

    @log_weights.setter
    def log_weights(self, v):
        v = np.ascontiguousarray(v, dtype=np.float32)
        assert len(v) == self.n_local
        _check(self.lib.phdslam_set_log_weights(self._h, v.ctypes.data))

    @property
    def map_sizes(self):
        out = np.zeros(self.n_local, dtype=np.int32)
        _check(self.lib.phdslam_get_map_sizes(self._h, out.ctypes.data))
        return out

    def get_maps(self):
        """(sizes, concatenated GAUSSIAN_DTYPE array) -- SynthSLAM::maps_static."""
        sizes = self.map_sizes
        out = np.zeros(int(sizes.sum()), dtype=GAUSSIAN_DTYPE)
        _check(self.lib.phdslam_get_maps(self._h, out.ctypes.data, len(out)))
        return sizes, out

    def set_maps(self, sizes, maps):
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        maps = np.ascontiguousarray(maps, dtype=GAUSSIAN_DTYPE)
        assert len(sizes) == self.n_local and int(sizes.sum()) == len(maps)
        _check(self.lib.phdslam_set_maps(self._h, sizes.ctypes.data, maps.ctypes.data))

    # ---- mixed feature model (feature_model = 2): SynthSLAM::maps_dynamic ----
    @property
    def map_sizes_dynamic(self):
        out = np.zeros(self.n_local, dtype=np.int32)
        _check(self.lib.phdslam_get_map_sizes_dynamic(self._h, out.ctypes.data))
        return out

    def get_maps_dynamic(self):
        sizes = self.map_sizes_dynamic
        out = np.zeros(int(sizes.sum()), dtype=GAUSSIAN4_DTYPE)
        _check(self.lib.phdslam_get_maps_dynamic(self._h, out.ctypes.data, len(out)))
        return sizes, out

    def set_maps_dynamic(self, sizes, maps):
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        maps = np.ascontiguousarray(maps, dtype=GAUSSIAN4_DTYPE)
        assert len(sizes) == self.n_local and int(sizes.sum()) == len(maps)
        _check(self.lib.phdslam_set_maps_dynamic(self._h, sizes.ctypes.data, maps.ctypes.data))

    def map_estimate_dynamic(self, cap=4096):
        """MAP estimate of the dynamic map (recoverSlamState: max_map_dynamic)."""
        out = np.zeros(cap, dtype=GAUSSIAN4_DTYPE)
        n = C.c_int()
        _check(self.lib.phdslam_map_estimate_dynamic(self._h, out.ctypes.data, cap, C.byref(n)))
        return out[:n.value].copy()

    @property
    def resample_idx(self):
        out = np.zeros(self.n_local, dtype=np.int32)
        _check(self.lib.phdslam_get_resample_idx(self._h, out.ctypes.data))
        return out

    @property
    def cardinalities(self):
        """CPHD: (n_local, max_cardinality+1) log cardinality distributions (SynthSLAM::cardinalities)"""
        out = np.zeros((self.n_local, self.cfg.max_cardinality + 1), dtype=np.float32)
        _check(self.lib.phdslam_get_cardinalities(self._h, out.ctypes.data))
        return out

    @cardinalities.setter
    def cardinalities(self, v):
        v = np.ascontiguousarray(v, dtype=np.float32)
        assert v.shape == (self.n_local, self.cfg.max_cardinality + 1)
        _check(self.lib.phdslam_set_cardinalities(self._h, v.ctypes.data))

    def timings(self):
        t = Timings()
        _check(self.lib.phdslam_get_timings(self._h, C.byref(t)))
        return t

    def snapshot(self):
        _check(self.lib.phdslam_snapshot(self._h))

    def restore(self):
        _check(self.lib.phdslam_restore(self._h))

    def set_overlap(self, on):
        """update / merge overlap on two streams (default off: no gain on B200); same results either way"""
        _check(self.lib.phdslam_set_overlap(self._h, int(bool(on))))

    @property
    def dist_p2p(self):
        """True when the resampling exchange pushes particles into the peers' buffers over NVLink (mapped with CUDA IPC)"""
        self.lib.phdslam_dist_p2p.argtypes = [C.c_void_p]
        return bool(self.lib.phdslam_dist_p2p(self._h))

    def synchronize(self):
        _check(self.lib.phdslam_synchronize(self._h))

    @property
    def stream(self):
        return self.lib.phdslam_stream(self._h)
