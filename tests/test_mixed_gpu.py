"""The MIXED feature model (feature_model = 2: static + constant-velocity features) on the GPU, through
the C-ABI: bit-identical to the CPU oracle (the arithmetic of both is include/phd_mixed_math.h + phd_detmath.h), and within
1e-4 of the REFERENCE's own kernels on the golden cases (tests/golden/ref_mixed_golden.npz, tests/test_mixed_ref_pin.py)."""
import os

import numpy as np
import pytest

import phdslam_b200 as P
from oracle import oracle as O
from conftest import GOLDEN
import mixed_cases as MC

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(GOLDEN, "ref_mixed_golden.npz"))


def gold(name):
    pre = name + "/"
    return {k[len(pre):]: GOLD[k] for k in GOLD.files if k.startswith(pre)}


def scene(n, ns, nd, M, labelled, seed, n_far_dynamic=2, n_far_static=3):
    """n particles around one truth pose; every particle its own jittered copy of a static and a dynamic map (different
    sizes per particle), some components out of range."""
    rng = np.random.default_rng(seed)
    poses = np.zeros(n, MC.POSE)
    poses["px"], poses["py"], poses["ptheta"] = rng.normal(0, 0.1, n), rng.normal(0, 0.1, n), rng.normal(0, 0.02, n)
    base_s = np.concatenate([MC.static_features(rng, ns), MC.static_features(rng, n_far_static, r_hi=13.0)])
    base_s["mean"][ns:] *= 3.0                                           # beyond max_range: class 0 / 2
    base_d = np.concatenate([MC.dynamic_features(rng, nd), MC.dynamic_features(rng, n_far_dynamic)])
    base_d["mean"][nd:, :2] *= 3.0
    ssz, dsz, sm, dm = [], [], [], []
    for p in range(n):
        ks = rng.permutation(len(base_s))[:rng.integers(max(len(base_s) - 3, 0), len(base_s) + 1)]
        kd = rng.permutation(len(base_d))[:rng.integers(max(len(base_d) - 2, 0), len(base_d) + 1)]
        s = base_s[np.sort(ks)].copy()
        d = base_d[np.sort(kd)].copy()
        s["mean"] += rng.normal(0, 0.05, s["mean"].shape).astype(np.float32)
        d["mean"] += rng.normal(0, 0.05, d["mean"].shape).astype(np.float32)
        ssz.append(len(s)); dsz.append(len(d)); sm.append(s); dm.append(d)
    targets = np.concatenate([base_s["mean"][:ns], base_d["mean"][:nd, :2]])
    Z = MC.measurements(rng, np.zeros(1, MC.POSE)[0], targets, M, labelled, n_static=ns)
    return dict(poses=poses, ssz=np.array(ssz, np.int32), smaps=np.concatenate(sm), dsz=np.array(dsz, np.int32),
                dmaps=np.concatenate(dm), Z=Z, targets=targets, ns=ns)


def load(f, sc):
    f.poses = sc["poses"]
    f.set_maps(sc["ssz"], sc["smaps"])
    f.set_maps_dynamic(sc["dsz"], sc["dmaps"])


def assert_same_state(g, o, what):
    gs, gm = g.get_maps()
    os_, om = o.get_maps()
    gd, gdm = g.get_maps_dynamic()
    od, odm = o.get_maps_dynamic()
    assert (gs == os_).all(), what + ": static component counts"
    assert (gd == od).all(), what + ": dynamic component counts"
    assert gm.tobytes() == om.tobytes(), what + ": static maps"
    assert gdm.tobytes() == odm.tobytes(), what + ": dynamic maps"
    assert g.log_weights.tobytes() == o.log_weights.tobytes(), what + ": log-weights"
    assert g.poses.tobytes() == o.poses.tobytes(), what + ": poses"


@pytest.mark.parametrize("name", list(MC.MIXED_CASES))
def test_golden_cases_against_the_reference_kernels(name):
    cfg, pose, sm, dm, Z = MC.build_case(name)
    gd = gold(name)
    g = P.PhdSlam(cfg)
    g.poses = pose
    g.set_maps([len(sm)], sm)
    g.set_maps_dynamic([len(dm)], dm)
    g.phdUpdateSynth(Z)
    _, smap = g.get_maps()
    _, dmap = g.get_maps_dynamic()
    assert len(dmap) == len(gd["d_merged"]) and len(smap) == len(gd["s_merged"])          # bit-exact counts
    MC.assert_gaussians_close(dmap, gd["d_merged"], name + " dynamic map")
    MC.assert_gaussians_close(smap, gd["s_merged"], name + " static map")
    # and the map prediction
    g2 = P.PhdSlam(cfg)
    g2.poses = pose
    g2.set_maps_dynamic([len(dm)], dm)
    g2.phdPredict(np.float32([0.0, 0.0]))
    _, pred = g2.get_maps_dynamic()
    MC.assert_gaussians_close(pred, gd["predicted"], name + " predicted", wfloor=1e-9)


@pytest.mark.parametrize("labelled,weighting,metric,mode", [(False, 0, 0, 0), (True, 0, 0, 0), (False, 1, 0, 0), (True, 1, 0, 0),
                                                            (False, 0, 1, 0), (False, 0, 0, 1), (True, 1, 0, 1), (False, 0, 0, 2)])
def test_filter_steps_are_bit_identical_to_the_oracle(labelled, weighting, metric, mode):
    """predict (poses + dynamic map), mixed update, estimate, resampling with injected and counter-based draws, over several
    steps and 96 particles with different map sizes.  mode 1: the fused update (update_mode = 1, no dense terms);
    mode 2: the merge of one sub-batch overlapping the update of the next (phdslam_set_overlap)."""
    n, M = 96, 14
    cfg = MC.mixed_config(n, M, labelled, weighting, distance_metric=metric, resample_threshold=1.1, seed="5",
                          update_mode=1 if mode == 1 else 0)
    sc = scene(n, 18, 6, M, labelled, seed=21 + weighting)
    g, o = P.PhdSlam(cfg), O.Oracle(cfg)
    if mode == 2:
        g.set_overlap(True)
    load(g, sc)
    load(o, sc)
    rng = np.random.default_rng(3)
    for k in range(5):
        u = np.float32([1.0, 0.02 * k])
        if k:
            g.phdPredict(u)
            o.phdPredict(u)
            assert_same_state(g, o, "step %d predict" % k)
        Z = MC.measurements(rng, np.zeros(1, MC.POSE)[0], sc["targets"], M - k, labelled, n_static=sc["ns"])
        g.phdUpdateSynth(Z)
        o.phdUpdateSynth(Z)
        assert_same_state(g, o, "step %d update" % k)
        eg, eo = g.recoverSlamState(), o.recoverSlamState()
        assert eg.map_particle == eo.map_particle and eg.pose.tobytes() == eo.pose.tobytes()
        mg = g.map_estimate_dynamic()
        _, odm = o.get_maps_dynamic()
        od = o.map_sizes_dynamic
        lo = int(od[:eo.map_particle].sum())
        assert mg.tobytes() == odm[lo:lo + od[eo.map_particle]].tobytes()
        uu = rng.uniform(0, 1, n + 1) if k % 2 == 0 else None
        pre = g.particle_checksums()
        ag, ao = g.resampleParticles(uu), o.resampleParticles(uu)
        assert (ag == ao).all()
        assert (g.particle_checksums() == pre[ag]).all()       # every offspring carries its ancestor's maps, both of them
        assert_same_state(g, o, "step %d resample" % k)
    sizes = g.map_sizes_dynamic
    assert sizes.max() > 0 and len(np.unique(sizes)) >= 1


def test_step_loop_and_snapshot_restore():
    n, M = 64, 10
    cfg = MC.mixed_config(n, M, seed="9", resample_threshold=0.6)
    sc = scene(n, 12, 5, M, False, seed=40)
    g, o = P.PhdSlam(cfg), O.Oracle(cfg)
    load(g, sc)
    load(o, sc)
    g.snapshot()
    rng = np.random.default_rng(8)
    zs = [MC.measurements(rng, np.zeros(1, MC.POSE)[0], sc["targets"], M, False) for _ in range(4)]
    for k, Z in enumerate(zs):
        eg, rg = g.step(k, np.float32([1.0, 0.01]), Z)
        eo, ro = o.step(k, np.float32([1.0, 0.01]), Z)
        assert rg == ro and eg.neff == eo.neff
        assert_same_state(g, o, "step %d" % k)
    first = g.get_maps_dynamic()[1].tobytes()
    g.restore()
    o2 = O.Oracle(cfg)
    load(o2, sc)
    assert_same_state(g, o2, "restored")
    for k, Z in enumerate(zs):
        g.step(k, np.float32([1.0, 0.01]), Z)
    assert g.get_maps_dynamic()[1].tobytes() == first


def test_unsupported_combinations_fail_loudly():
    for kw in (dict(feature_model=1), dict(filter_type=1, max_cardinality=63), dict(n_predict_particles=2)):
        ov = dict(MC.MIXED_OVERRIDES)
        ov.update(kw)
        cfg = MC.S.scene_config_py(8, 8, 4, **ov)
        with pytest.raises(Exception):
            P.PhdSlam(cfg)
    # static feature model: no dynamic maps
    with pytest.raises(Exception):
        P.PhdSlam(MC.S.scene_config_py(4, 8, 4)).get_maps_dynamic()


def test_dynamic_map_capacity_is_reported():
    cfg = MC.mixed_config(2, 40, max_components_dynamic=8, min_separation=1e-6)
    rng = np.random.default_rng(2)
    g = P.PhdSlam(cfg)
    Z = MC.measurements(rng, np.zeros(1, MC.POSE)[0], np.zeros((0, 2)), 40, False)     # 40 births, nothing merges
    with pytest.raises(Exception) as e:
        g.phdUpdateSynth(Z)
    assert "dynamic" in str(e.value)


def test_cli_writes_the_dynamic_map_line(tmp_path):
    """The process interface with feature_model = 2 and the 7-line log layout of writeLog (src/main.cpp:848-954): line 3 is
    the dynamic map of the maximum-weight particle, 21 numbers per component; the whole file equals the oracle's, text for
    text, over steps that include resampling."""
    import subprocess
    from conftest import ROOT
    import test_mixed_tracking as T
    exe = os.path.join(ROOT, "cuda-phdslam_b200", "phdslam")
    assert os.path.exists(exe), "CLI not built"
    cfg = T.config(16)
    zs = T.measurement_sets(cfg.dt)[:12]
    data = tmp_path / "data"
    data.mkdir()
    with open(data / "measurements.txt", "w") as f:
        f.write("header\n")
        for Z in zs:
            f.write(" ".join("%.9g" % v for v in Z.reshape(-1)) + "\n")
    with open(data / "controls.txt", "w") as f:
        f.write("header\n")
        for _ in zs:
            f.write("0 0\n")
    cfgfile = tmp_path / "mixed.cfg"
    keys = dict(feature_model=2, n_particles=16, motion_type=0, acc_x=0.2, acc_y=0.2, acc_yaw=0.01, dt=cfg.dt, max_range=cfg.max_range,
                max_bearing=cfg.max_bearing, min_range=0, std_range=cfg.std_range, std_bearing=cfg.std_bearing, pd=cfg.pd,
                clutter_rate=cfg.clutter_rate, birth_weight=cfg.birth_weight, birth_noise_factor=cfg.birth_noise_factor,
                min_feature_weight=cfg.min_feature_weight, min_separation=cfg.min_separation, particle_weighting=0,
                filter_type=0, map_estimate=1, resample_threshold=0.9, std_ax_features=0.3, std_ay_features=0.3,
                cov_vx_birth=1.0, cov_vy_birth=1.0, tau=0.25, beta=8.0, ps=0.99, max_components=cfg.max_components,
                max_components_dynamic=128, log_layout="extended", seed=4, data_directory=str(data) + "/")
    with open(cfgfile, "w") as f:
        for k, v in keys.items():
            f.write("%s = %s\n" % (k, v))
    out = tmp_path / "run"
    r = subprocess.run([exe, str(cfgfile), "synth", "--out", str(out), "--steps", str(len(zs)), "--quiet"], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ocfg = P.load_config(str(cfgfile))
    o = O.Oracle(ocfg)
    resampled_any = False
    U = np.zeros((len(zs), 2), np.float32)
    for k, Z in enumerate(zs):
        e = o.step_filter(k, U[k - 1] if k else np.float32([0, 0]), Z)
        ds, dm = o.get_maps_dynamic()
        lo = int(ds[:e.map_particle].sum())
        ref = tmp_path / ("ref%05d.log" % k)
        P.write_log(str(ref), 1, e.pose, o.map_estimate(1), o.log_weights, o.poses, resample_idx=o.resample_idx,
                    n_card=ocfg.max_cardinality + 1, map_dynamic=dm[lo:lo + ds[e.map_particle]])
        resampled_any |= o.step_resample(len(Z), e)
        got = open(out / ("state_estimate%05d.log" % k)).read()
        assert got == open(ref).read(), "log of step %d differs" % k
        lines = got.split("\n")
        assert len(lines) == 8 and len(lines[2].split()) == 21 * int(ds[e.map_particle])
    assert resampled_any


# ---- particle sharding: one process per GPU; the dynamic maps travel with their particles at resampling ----
def _sharded_worker(rank, world, uid, q, labelled, n):
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import phdslam_b200 as P2
    M = 12
    cfg = MC.mixed_config(n, M, labelled, 0, resample_threshold=1.1, seed="6")
    sc = scene(n, 14, 5, M, labelled, seed=33)
    g = P2.PhdSlam(cfg, device=rank)
    g.dist_init(rank, world, unique_id=uid)
    lo, k = g.local_offset, g.n_local
    so, do = np.concatenate([[0], np.cumsum(sc["ssz"])]), np.concatenate([[0], np.cumsum(sc["dsz"])])
    g.poses = sc["poses"][lo:lo + k]
    g.log_weights = np.full(k, -np.log(n), np.float32) + np.float32(np.linspace(-3, 0, n)[lo:lo + k])   # skewed: offspring cross GPUs
    g.set_maps(sc["ssz"][lo:lo + k], sc["smaps"][so[lo]:so[lo + k]])
    g.set_maps_dynamic(sc["dsz"][lo:lo + k], sc["dmaps"][do[lo]:do[lo + k]])
    out = {}
    rng = np.random.default_rng(12)
    for step in range(3):
        if step:
            g.phdPredict(np.float32([1.0, 0.03]))
        Z = MC.measurements(rng, np.zeros(1, MC.POSE)[0], sc["targets"], M, labelled, n_static=sc["ns"])
        g.phdUpdateSynth(Z)
        e = g.recoverSlamState()
        out["est%d" % step] = (e.pose.copy(), e.map_particle, g.map_estimate_dynamic())
        out["pre%d" % step] = g.particle_checksums()
        out["anc%d" % step] = g.resampleParticles(np.random.default_rng(50 + step).uniform(0, 1, n + 1) if step < 2 else None)
        out["post%d" % step] = g.particle_checksums()
        out["dyn%d" % step] = g.get_maps_dynamic()
        out["sta%d" % step] = g.get_maps()
        out["w%d" % step] = g.log_weights
    out["migrated"] = g.timings().migrated_in
    q.put((rank, lo, k, out))


@pytest.mark.parametrize("world,p2p,labelled", [(2, 1, False), (2, 0, True), (4, 1, False), (3, 0, False)])
def test_sharding_of_the_mixed_model_matches_oracle(world, p2p, labelled, monkeypatch):
    """2 GPUs: one shift of the dynamic-map ring; 3 and 4 GPUs (a particle count the world size does not divide): the general
    ring.  Both exchange modes of the static maps."""
    import torch
    n = 75
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (run with gpurun --gpus %d)" % (world, world))
    monkeypatch.setenv("PHDSLAM_P2P", str(p2p))
    import torch.multiprocessing as mp
    uid = P.dist_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, uid, q, labelled, n)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    M = 12
    cfg = MC.mixed_config(n, M, labelled, 0, resample_threshold=1.1, seed="6")
    sc = scene(n, 14, 5, M, labelled, seed=33)
    o = O.Oracle(cfg)
    load(o, sc)
    o.log_weights = np.full(n, -np.log(n), np.float32) + np.float32(np.linspace(-3, 0, n))
    rng = np.random.default_rng(12)
    crossed = 0
    for step in range(3):
        if step:
            o.phdPredict(np.float32([1.0, 0.03]))
        Z = MC.measurements(rng, np.zeros(1, MC.POSE)[0], sc["targets"], M, labelled, n_static=sc["ns"])
        o.phdUpdateSynth(Z)
        e = o.recoverSlamState()
        ds, dm = o.get_maps_dynamic()
        lo = int(ds[:e.map_particle].sum())
        owner = [r for r in res if r[1] <= e.map_particle < r[1] + r[2]][0]
        for r in res:
            pose, mp_, dyn = r[3]["est%d" % step]
            assert pose.tobytes() == e.pose.tobytes() and mp_ == e.map_particle
            if r is owner:
                assert dyn.tobytes() == dm[lo:lo + ds[e.map_particle]].tobytes()
            else:
                assert len(dyn) == 0
        anc = o.resampleParticles(np.random.default_rng(50 + step).uniform(0, 1, n + 1) if step < 2 else None)
        assert (np.concatenate([r[3]["anc%d" % step] for r in res]) == anc).all()
        assert (np.concatenate([r[3]["post%d" % step] for r in res]) == np.concatenate([r[3]["pre%d" % step] for r in res])[anc]).all()
        own = lambda i: np.searchsorted([r[1] for r in res], i, side="right") - 1
        crossed += int((own(np.arange(n)) != own(anc)).sum())
        ods, odm = o.get_maps_dynamic()
        oss, osm = o.get_maps()
        assert (np.concatenate([r[3]["dyn%d" % step][0] for r in res]) == ods).all()
        assert np.concatenate([r[3]["dyn%d" % step][1] for r in res]).tobytes() == odm.tobytes()
        assert (np.concatenate([r[3]["sta%d" % step][0] for r in res]) == oss).all()
        assert np.concatenate([r[3]["sta%d" % step][1] for r in res]).tobytes() == osm.tobytes()
        assert np.concatenate([r[3]["w%d" % step] for r in res]).tobytes() == o.log_weights.tobytes()
    assert crossed > 0 and sum(r[3]["migrated"] for r in res) > 0       # offspring (and their dynamic maps) crossed GPUs
