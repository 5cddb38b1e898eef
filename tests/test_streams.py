"""Time-stamped asynchronous measurement / control streams and follow_trajectory (SURVEY 8(f) rank 2;
reference src/main.cpp:147-167, 247-264, 1087-1127, 1187-1243)."""
import os
import subprocess

import numpy as np
import pytest

import phdslam_b200 as P
from conftest import DATA, GOLDEN, ROOT


def literal_schedule(zt, ct):
    """run_synth's loop head, restated line by line (main.cpp:1187-1230) with REAL = float"""
    current_time = np.float32(0)
    z_idx = c_idx = 0
    out = []
    for _ in range(len(zt) + len(ct)):
        if z_idx >= len(zt) or c_idx >= len(ct):
            break
        tz, tc = np.float32(zt[z_idx]), np.float32(ct[c_idx])
        last_time = current_time
        current_time = tc
        dt = np.float32(current_time - last_time)
        if tz < tc:
            out.append((z_idx, -1, dt)); z_idx += 1
        elif tz == tc:
            out.append((z_idx, c_idx, dt)); z_idx += 1; c_idx += 1
        else:
            out.append((-1, c_idx, dt)); c_idx += 1
    return out


def test_loaders(tmp_path):
    p = tmp_path / "measurement_times.txt"
    p.write_text("0.1\n0.25\n0.4\n\n")
    np.testing.assert_allclose(P.load_timestamps(str(p)), [0.1, 0.25, 0.4])
    assert len(P.load_timestamps(str(tmp_path / "missing.txt"))) == 0          # no file: no time stamps (main.cpp:1092)
    t = tmp_path / "traj.txt"
    t.write_text("% x y theta vx vy vtheta\n1 2 0.5 0.1 0 0.01\n3 4 -0.5 0.2 0 0.02\n")
    tr = P.load_trajectory(str(t))
    assert len(tr) == 2 and tr["px"].tolist() == [1.0, 3.0] and abs(tr["vtheta"][1] - 0.02) < 1e-7
    with pytest.raises(P.PhdSlamError):
        P.load_trajectory(str(tmp_path / "missing.txt"))


@pytest.mark.parametrize("seed", range(4))
def test_event_schedule_is_the_reference_loop(seed):
    rng = np.random.default_rng(seed)
    ct = np.round(np.cumsum(rng.uniform(0.01, 0.05, 60)), 3)
    zt = np.sort(np.concatenate([rng.choice(ct, 8, replace=False), np.round(rng.uniform(0, ct[-1], 20), 3)]))
    ev = P.plan_events(zt, ct)
    lit = literal_schedule(zt, ct)
    assert len(ev) == len(lit) and len(ev) > 20
    for e, (zi, ci, dt) in zip(ev, lit):
        assert (e["z_idx"], e["c_idx"]) == (zi, ci) and e["dt"] == dt
    kinds = {(e["z_idx"] >= 0, e["c_idx"] >= 0) for e in ev}
    assert kinds == {(True, False), (True, True), (False, True)}              # all three branches exercised


def _run_cli(args, timeout=300):
    exe = os.path.join(ROOT, "cuda-phdslam_b200", "phdslam")
    assert os.path.exists(exe), "CLI not built"
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


@pytest.mark.gpu
def test_cli_timestamped_streams_match_oracle(tmp_path):
    """controls at 50 Hz, a measurement set every third control (equal time stamps) plus two sets between controls"""
    from oracle import oracle as O
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_ackerman.txt"))[:9]
    U = P.load_controls(os.path.join(DATA, "controls_synth.txt"))[:24]
    ct = 0.02 * (1 + np.arange(len(U)))
    zt = np.array([ct[0], ct[3], ct[6], ct[8] + 0.01, ct[9], ct[12], ct[15] + 0.005, ct[18], ct[21]])
    d = tmp_path / "data"
    d.mkdir()
    with open(d / "measurements.txt", "w") as f:
        f.write("% header\n")
        for z in Z:
            f.write(" ".join("%.6f" % v for v in z.reshape(-1)) + "\n")
    with open(d / "controls.txt", "w") as f:
        f.write("% header\n")
        for u in U:
            f.write("%.6f %.6f\n" % (u[0], u[1]))
    np.savetxt(d / "measurement_times.txt", zt, fmt="%.6f")
    np.savetxt(d / "control_times.txt", ct, fmt="%.6f")
    out = tmp_path / "run"
    _run_cli([os.path.join(GOLDEN, "config_ackerman.cfg"), "synth", "--out", str(out), "--set", "data_directory=%s/" % d,
              "--set", "n_particles=48", "--set", "map_estimate=1", "--set", "seed=5", "--quiet"])
    # the same loop on the oracle
    Zt = P.load_measurements(str(d / "measurements.txt"))
    Ut = P.load_controls(str(d / "controls.txt"))
    ev = P.plan_events(P.load_timestamps(str(d / "measurement_times.txt")), P.load_timestamps(str(d / "control_times.txt")))
    cfg = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    cfg.set(n_particles=48, map_estimate=1, seed="5")
    o = O.Oracle(cfg)
    u = np.float32([0, 0])
    assert len(ev) >= 20 and (ev["z_idx"] >= 0).sum() == 9
    for k, e in enumerate(ev):
        if e["c_idx"] >= 0:
            u = Ut[e["c_idx"]]
        cfg.set(dt=float(e["dt"]))
        o.setDeviceConfig(cfg)
        zk = Zt[e["z_idx"]] if e["z_idx"] >= 0 else np.zeros((0, 2), np.float32)
        est = o.step_filter(k, u, zk)
        ref = tmp_path / ("ref%05d.log" % k)       # the log is the state before resampling (main.cpp:1274-1279)
        P.write_log(str(ref), 0, est.pose, o.map_estimate(1), o.log_weights, o.poses, n_card=cfg.max_cardinality + 1)
        o.step_resample(len(zk), est)
        assert open(out / ("state_estimate%05d.log" % k)).read() == open(ref).read(), "log of event %d differs" % k
    assert not os.path.exists(out / ("state_estimate%05d.log" % len(ev)))


@pytest.mark.gpu
def test_cli_follow_trajectory_matches_oracle(tmp_path):
    """follow_trajectory: one particle pinned to traj.txt, mapping only (main.cpp:1122-1127, 1239-1243)"""
    from oracle import oracle as O
    tr = np.load(os.path.join(GOLDEN, "truth_ackerman.npz"))["traj"][:15]
    t = tmp_path / "traj.txt"
    with open(t, "w") as f:
        f.write("% x y theta vx vy vtheta\n")
        for p in tr:
            f.write("%.6f %.6f %.6f 0 0 0\n" % (p[0], p[1], p[2]))
    out = tmp_path / "run"
    sets = ["follow_trajectory=1", "map_estimate=1", "max_range=10", "std_range=1.0", "std_bearing=0.0349", "birth_weight=0.005",
            "min_feature_weight=0.00001", "seed=3"]
    args = [os.path.join(GOLDEN, "config_ackerman.cfg"), "synth", "--measurements", os.path.join(DATA, "measurements_synth_ackerman.txt"),
            "--controls", os.path.join(DATA, "controls_synth.txt"), "--trajectory", str(t), "--out", str(out), "--quiet"]
    for kv in sets:
        args += ["--set", kv]
    _run_cli(args)
    cfg = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    cfg.set(follow_trajectory=1, map_estimate=1, max_range=10.0, std_range=1.0, std_bearing=0.0349, birth_weight=0.005,
            min_feature_weight=1e-5, seed="3", n_particles=1)
    Z = P.load_measurements(os.path.join(DATA, "measurements_synth_ackerman.txt"))
    traj = P.load_trajectory(str(t))
    o = O.Oracle(cfg)
    for k in range(len(traj)):
        o.poses = traj[k:k + 1]
        est = o.step_filter(0, np.float32([0, 0]), Z[k])
        ref = tmp_path / ("ref%05d.log" % k)
        P.write_log(str(ref), 0, est.pose, o.map_estimate(1), o.log_weights, o.poses, n_card=cfg.max_cardinality + 1)
        o.step_resample(len(Z[k]), est)
        assert open(out / ("state_estimate%05d.log" % k)).read() == open(ref).read(), "log of step %d differs" % k
    assert not os.path.exists(out / ("state_estimate%05d.log" % len(traj)))
    # mapping with the true poses: the landmarks seen so far are in the map
    m = o.map_estimate(1)
    assert (m["weight"] > 0.5).sum() >= 10
