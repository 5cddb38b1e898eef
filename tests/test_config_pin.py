"""cfg/config.cfg is part of the drop-in interface (BASELINE.json north_star): every option the reference registers with
boost::program_options in loadConfig (src/main.cpp:960-1048) must be accepted by phdslam_config_set under the same key, and
the options of the static-map path must have the reference's default value.  The option table is read from the
reference's source at test time (CPU container only); nothing is copied."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

import phdslam_b200 as P

REF_MAIN = "/root/reference/src/main.cpp"
needs_ref = pytest.mark.skipif(not os.path.exists(REF_MAIN), reason="reference sources not present (GPU box)")

OPTION = re.compile(r'^\s*\("([a-z_0-9]+)",\s*value<([^>]+)>\(&([A-Za-z_.0-9]+)\)->default_value\(([^)]*)\)')


def reference_options():
    out = []
    for line in open(REF_MAIN, encoding="latin-1").read().split("\n")[959:1049]:
        m = OPTION.match(line)
        if m:
            out.append((m.group(1), m.group(2).strip(), m.group(3), m.group(4).strip()))
    return out


def fields_of(cfg):
    vals = {}
    for name, _ in cfg._fields_:
        v = getattr(cfg, name)
        vals[name] = bytes(v) if isinstance(v, (bytes, C.Array)) else v
    return vals


@needs_ref
def test_every_reference_option_is_accepted_and_defaults_match():
    opts = reference_options()
    assert len(opts) > 80 and ("min_separation", "REAL", "config.minSeparation", "5") in opts
    lib = P.load_library()
    lib.phdslam_config_set.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    base = P.default_config()
    base_fields = fields_of(base)
    mapped, ignored = {}, []
    for key, typ, target, default in opts:
        if key == "data_directory":
            probe = b"some/dir/"
        elif typ == "bool":
            probe = b"true" if default == "false" else b"false"
        elif typ == "int":
            probe = b"7"
        else:
            probe = b"0.3125"
        cfg = P.default_config()
        rc = lib.phdslam_config_set(C.byref(cfg), key.encode(), probe)
        assert rc == 0, "reference option '%s' is rejected" % key
        changed = [n for n, v in fields_of(cfg).items() if v != base_fields[n] and n != "clutter_density"]
        if changed:
            assert len(changed) == 1, (key, changed)
            mapped[key] = changed[0]
        else:
            ignored.append(key)
    # every option of the path this library replaces is live (not merely swallowed)
    live = {"initial_x", "initial_y", "initial_yaw", "initial_vx", "initial_vyaw", "follow_trajectory", "motion_type", "acc_x",
            "acc_y", "acc_yaw", "dt", "max_bearing", "min_range", "max_range", "std_bearing", "std_range", "clutter_rate", "pd",
            "n_particles", "n_predict_particles", "resample_threshold", "subdivide_predict", "birth_weight", "birth_noise_factor",
            "feature_model", "min_separation", "min_feature_weight", "particle_weighting", "max_cardinality", "filter_type",
            "map_estimate", "distance_metric", "h", "l", "a", "b", "std_encoder", "std_alpha", "labeled_measurements",
            "data_directory", "max_time_steps", "n_steps"}
    assert live <= set(mapped), sorted(live - set(mapped))
    # defaults of the live options = the reference's default_value(...)
    for key, typ, target, default in opts:
        if key not in mapped or key == "data_directory":
            continue
        ours = base_fields[mapped[key]]
        if typ == "bool":
            want = 1 if default == "true" else 0
        elif default == "M_PI":
            want = np.float32(math.pi)
        else:
            want = int(default) if typ == "int" else np.float32(float(default))
        assert ours == want, "default of '%s': %r, the reference has %s" % (key, ours, default)
    assert base_fields[mapped["data_directory"]].rstrip(b"\0") == b"data/"
    # the derived clutter density (main.cpp:1065)
    assert base.clutter_density == np.float32(base.clutter_rate) / (np.float32(2) * np.float32(base.max_bearing) * np.float32(base.max_range))
    print("live:", len(mapped), "accepted and ignored:", sorted(ignored))


@needs_ref
def test_reference_cfg_files_load_unchanged(tmp_path):
    """the cfg files the reference ships parse without a rejected line"""
    import glob
    files = sorted(glob.glob("/root/reference/cfg/*.cfg*"))
    assert files
    for f in files:
        cfg = P.load_config(f)
        assert cfg.n_particles > 0 and cfg.max_range > 0
