"""C-ABI surface and host-side logic (no GPU): the library loads, exports every declared symbol,
and the config / loader / log-writer mirror the reference's behaviour."""
import ctypes as C
import math
import os
import re

import numpy as np
import pytest

from conftest import DATA, GOLDEN, ROOT

import phdslam_b200 as P


def test_library_exports_every_header_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "phdslam.h")).read()
    declared = set(re.findall(r"\b(phdslam_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"phdslam_config_t", "phdslam_t"}
    assert declared == set(P.ABI_SYMBOLS), (declared ^ set(P.ABI_SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), name


def test_config_struct_size_matches(lib):
    # the ctypes mirror must have the C layout: poke the last field through the C setter
    cfg = P.default_config(update_buffer_bytes="12345", seed="77", max_components=96)
    assert cfg.update_buffer_bytes == 12345 and cfg.seed == 77 and cfg.max_components == 96


def test_defaults_are_loadConfig_defaults():
    # reference src/main.cpp:961-1048
    c = P.default_config()
    assert c.motion_type == 1 and c.n_particles == 512 and c.filter_type == 1 and c.particle_weighting == 1
    assert c.max_cardinality == 256 and c.map_estimate == 1 and c.n_predict_particles == 1
    assert abs(c.pd - 0.98) < 1e-7 and abs(c.birth_weight - 0.05) < 1e-7 and abs(c.birth_noise_factor - 1.5) < 1e-7
    assert abs(c.max_bearing - math.pi) < 1e-6 and c.max_range == 20 and c.min_range == 0
    assert abs(c.min_feature_weight - 1e-5) < 1e-10 and c.min_separation == 5
    assert abs(c.clutter_density - 15.0 / (2 * math.pi * 20)) < 1e-7


def test_load_reference_style_cfg():
    c = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    assert c.n_particles == 200 and c.filter_type == 0 and c.particle_weighting == 0 and c.map_estimate == 0
    assert c.max_cardinality == 255 and c.feature_model == 0  # inline comment stripped
    assert abs(c.std_alpha - 0.034907) < 1e-8 and abs(c.l - 1.415) < 1e-7
    # clutterDensity = clutterRate/(2*maxBearing*maxRange) (main.cpp:1065)
    assert abs(c.clutter_density - 20.0 / (2 * 3.141593 * 15.0)) < 1e-7
    assert c.data_directory == b"data/"


def test_unknown_key_is_reported_not_fatal(tmp_path, capfd):
    p = tmp_path / "c.cfg"
    p.write_text("n_particles = 7\nno_such_key = 1\nmin_separation = 3\n")
    c = P.load_config(str(p))
    assert c.n_particles == 7 and c.min_separation == 3
    assert "no_such_key" in capfd.readouterr().err


def test_initial_vy_quirk(tmp_path):
    # initial_vz overrides initial_vy in the reference parser (main.cpp:969-970)
    p = tmp_path / "c.cfg"
    p.write_text("initial_vy = 2\ninitial_vz = 0.5\n")
    c = P.load_config(str(p))
    assert c.vy0 == 0.5


def test_load_bundled_measurements_and_controls():
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_ackerman.txt"))
    assert len(Z) == 331
    n = [len(z) for z in Z]
    assert min(n) == 14 and max(n) == 44 and Z[0].shape[1] == 2
    assert abs(Z[0][0, 0] - 9.476848) < 1e-6 and abs(Z[0][0, 1] + 2.299487) < 1e-6
    U = P.load_controls(os.path.join(DATA, "controls_synth.txt"))
    assert U.shape[1] == 2 and len(U) in (330, 999)
    assert abs(U[0, 0] - 2.77796) < 1e-5 and abs(U[0, 1] + 0.186915) < 1e-6
    Zcv = P.load_measurements(os.path.join(DATA, "measurements_synth_cv.txt"))
    assert len(Zcv) == 1000


def test_measurement_parser_edge_cases(tmp_path):
    p = tmp_path / "m.txt"
    p.write_text("% header\n1 0.5 2 -0.5 \n\n3 0.1 0\n4 0.2")
    Z = P.load_measurements(str(p), fields=2)
    assert [len(z) for z in Z] == [2, 0, 1, 1]        # trailing blank: no garbage measurement; "3 0.1 0" -> odd token dropped
    Z3 = P.load_measurements(str(p), fields=3)
    assert [len(z) for z in Z3] == [1, 0, 1, 0]
    p2 = tmp_path / "nohdr.txt"
    p2.write_text("1 2\n3 4\n")
    assert len(P.load_measurements(str(p2), has_header=-1)) == 2
    assert len(P.load_measurements(str(p2), has_header=1)) == 1   # reference always skips line 1 (main.cpp:230)


def test_write_log_five_line_layout(tmp_path):
    # README:31-39 -- 5 lines, 6 significant digits, trailing blank before newline
    poses = np.zeros(2, dtype=P.POSE_DTYPE)
    poses["px"] = [1.5, 2.25]
    m = np.zeros(1, dtype=P.GAUSSIAN_DTYPE)
    m["weight"] = 0.75
    m["mean"] = [[3.0, 4.0]]
    m["cov"] = [[0.1, 0.01, 0.01, 0.2]]
    path = str(tmp_path / "state_estimate00000.log")
    P.write_log(path, 0, [1.23456789, 2, 3, 0, 0, 0], m, [-0.693147, -0.693147], poses, n_card=3)
    lines = open(path).read().split("\n")
    assert len(lines) == 6 and lines[5] == ""
    assert lines[0] == "1.23457 2 3 0 0 0 "
    assert lines[1] == "0.75 3 4 0.1 0.01 0.01 0.2 "
    assert lines[2] == "-0.693147 -0.693147 "
    assert lines[3].split() == ["1.5", "0", "0", "0", "0", "0", "2.25", "0", "0", "0", "0", "0"]
    assert lines[4] == "0 0 0 "
    P.write_log(path, 1, [0] * 6, m[:0], [-0.1], poses[:1], resample_idx=[0], n_card=1)
    assert len(open(path).read().split("\n")) == 8     # 7-line writeLog layout (main.cpp:848-954)


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        return
    cfg = P.default_config(filter_type=0, n_particles=4)
    try:
        P.PhdSlam(cfg)
    except P.PhdSlamError as e:
        assert e.code == -1
    else:
        raise AssertionError("PhdSlam must fail loudly without a CUDA device")


def test_header_is_plain_c99_and_links(tmp_path):
    """The drop-in boundary is a C ABI: include/phdslam.h compiles as C99 (-pedantic, no C++), a C program links against
    libphdslam.so and calls an entry point that needs no device."""
    import subprocess
    src = tmp_path / "cabi.c"
    src.write_text('#include "phdslam.h"\n'
                   'int main(void) { phdslam_config_t c; phdslam_config_defaults(&c);\n'
                   '  return (c.n_particles == 512 && phdslam_version() != 0) ? 0 : 1; }\n')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "cuda-phdslam_b200")
    exe = str(tmp_path / "cabi")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src),
                           "-L", lib_dir, "-lphdslam", "-Wl,-rpath," + lib_dir, "-o", exe])
    assert subprocess.call([exe]) == 0


def test_product_path_fails_loudly_without_a_device(tmp_path):
    """No CPU fallback: on a machine without a GPU phdslam_create reports PHDSLAM_ERR_CUDA and the CLI exits non-zero
    (skipped where a device is present)."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cfg = P.default_config(n_particles=8)
    with pytest.raises(P.PhdSlamError) as e:
        P.PhdSlam(cfg, device=0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([os.path.join(root, "cuda-phdslam_b200", "phdslam"), os.path.join(GOLDEN, "config_ackerman.cfg"), "synth",
                        "--measurements", os.path.join(DATA, "measurements_synth_ackerman.txt"),
                        "--controls", os.path.join(DATA, "controls_synth.txt"), "--out", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
