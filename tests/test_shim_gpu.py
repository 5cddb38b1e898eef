"""The drop-in boundary, EXECUTED: the reference's own run_synth loop -- input loading, particle initialisation, the
time-step loop with its host-side resampleParticles, recoverSlamState and writeLog, all cut verbatim from src/main.cpp by
oracle/ref_build.sh -- linked with cuda-phdslam_b200/shim/phdfilter_b200.cpp (phdPredict / phdUpdateSynth /
setDeviceConfig over the C-ABI) into oracle/_ref/shim_replay, against the `phdslam` CLI on the same inputs.

The run crosses several resampling steps: the reference resamples on the HOST and replaces `particles`
(src/main.cpp:1286-1289), so the shim has to notice that the host copy changed and upload it again -- the defect of the
first-round shim, which kept predicting the un-resampled device particles.

Both programs write the reference's 7-line state_estimateNNNNN.log (writeLog, src/main.cpp:848-954).  Resample indices
and map sizes must be equal; numbers agree to 1e-4 relative (the reference's recoverSlamState sums exp(w) * pose in fp32 in
particle order, the library in fixed point; its resampler walks a double CDF, the library an integer one)."""
import os
import subprocess

import numpy as np
import pytest

import phdslam_b200 as P
from conftest import DATA, GOLDEN, ROOT

pytestmark = pytest.mark.gpu
REPLAY = os.path.join(ROOT, "oracle", "_ref", "shim_replay")
CLI = os.path.join(ROOT, "cuda-phdslam_b200", "phdslam")


def _write_data(d, n_steps):
    """the bundled Ackerman scene in the formats HEAD's parsers define: `r b label` triples (src/main.cpp:192-208)"""
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_ackerman.txt"))[:n_steps]
    U = P.load_controls(os.path.join(DATA, "controls_synth.txt"))[:n_steps + 2]
    with open(os.path.join(d, "measurements.txt"), "w") as f:
        f.write("% range bearing label\n")
        for z in Z:
            f.write(" ".join("%.7g %.7g 0" % (r, b) for r, b in z) + "\n")
    with open(os.path.join(d, "controls.txt"), "w") as f:
        f.write("% v_encoder alpha\n")
        for u in U:
            f.write("%.7g %.7g\n" % (u[0], u[1]))


def _parse(path):
    with open(path) as f:
        return [np.array(line.split(), dtype=np.float64) for line in f.read().split("\n")[:7]]


@pytest.mark.skipif(not os.path.exists(REPLAY), reason="oracle/_ref/shim_replay not built (needs /root/reference at build time)")
@pytest.mark.parametrize("filter_type", [0, 1])
def test_reference_run_synth_over_the_shim_matches_the_cli(tmp_path, filter_type):
    n_steps = 30
    d = tmp_path / "data"
    d.mkdir()
    _write_data(str(d), n_steps)
    sets = ["data_directory=%s/" % d, "n_particles=48", "map_estimate=1", "seed=21", "n_steps=%d" % n_steps,
            "resample_threshold=0.75", "max_components=256", "filter_type=%d" % filter_type, "max_cardinality=63",
            "measurement_fields=3"]
    cfg = os.path.join(GOLDEN, "config_ackerman.cfg")
    out_s, out_c = tmp_path / "shim", tmp_path / "cli"
    out_s.mkdir()
    r = subprocess.run([REPLAY, cfg, str(out_s)] + sets, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    args = [CLI, cfg, "synth", "--out", str(out_c), "--quiet", "--set", "log_layout=1"]
    for kv in sets:
        args += ["--set", kv]
    r = subprocess.run(args, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    resampled = 0
    for k in range(n_steps):
        a, b = _parse(out_s / ("state_estimate%05d.log" % k)), _parse(out_c / ("state_estimate%05d.log" % k))
        assert len(a) == len(b) == 7
        # line 6: resample indices (bit-exact); a step whose indices are not the identity follows a resampling.  At step 0
        # there is no previous resampling: the reference's vector is still value-initialised (all zeros,
        # src/slamtypes.h:282-283), the CLI writes the identity.
        if k == 0:
            assert (a[5] == 0).all() and (b[5] == np.arange(len(b[5]))).all()
        else:
            assert a[5].shape == b[5].shape and (a[5] == b[5]).all(), "step %d: resample indices differ" % k
            resampled += int((a[5] != np.arange(len(a[5]))).any())
        for i, what in ((0, "expected pose"), (1, "map estimate"), (3, "log-weights"), (4, "particle poses"), (6, "cardinality")):
            assert a[i].shape == b[i].shape, "step %d: %s has another size" % (k, what)
            np.testing.assert_allclose(a[i], b[i], rtol=1e-4, atol=2e-5, err_msg="step %d: %s" % (k, what))
    assert resampled >= 3, "the run must cross several host-side resampling steps (saw %d)" % resampled
    assert not os.path.exists(out_s / ("state_estimate%05d.log" % n_steps))
