"""Particle sharding over 2 GPUs (one process per GPU, NCCL inside libphdslam.so): weight normalisation,
state estimate and global resampling with migration reproduce the single-process oracle bit-for-bit."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cfg(S, N, C, M, cphd):
    extra = dict(filter_type=1, max_cardinality=63) if cphd else {}
    return S.scene_config(N, C, M, max_components=64, seed="11", **extra)


def _scene(S, N, C, M, skew):
    sc = S.make_scene(N, C, M, seed=4, n_near=2, n_far=2)
    if skew:
        # skewed prior weights: the first quarter of the particles (the first ranks) carries almost everything, a block in
        # the middle next to nothing -- most offspring descend from ancestors on OTHER ranks, some ranks serve nobody
        rng = np.random.default_rng(77)
        w = rng.uniform(0.5, 1.5, N)
        w[N // 4:] *= 1e-3
        w[N // 2: (5 * N) // 8] *= 1e-6
        sc["log_weights"] = np.log(w / w.sum()).astype(np.float32)
    return sc


def _worker(rank, world, uid, q, cphd=False, N=301, skew=False):
    sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
    sys.path.insert(0, ROOT)
    import phdslam_b200 as P
    from phdslam_b200 import scene as S
    C, M = 16, 8
    cfg = _cfg(S, N, C, M, cphd)
    sc = _scene(S, N, C, M, skew)
    g = P.PhdSlam(cfg, device=rank)
    g.dist_init(rank, world, unique_id=uid)
    lo, n = g.local_offset, g.n_local
    per = C + 4
    g.poses = sc["poses"][lo:lo + n]
    g.log_weights = sc["log_weights"][lo:lo + n]
    g.set_maps(sc["sizes"][lo:lo + n], sc["maps"][lo * per:(lo + n) * per])
    out = {}
    g.phdPredict(np.float32([1.0, 0.05]))
    g.phdUpdateSynth(sc["Z"])
    e = g.recoverSlamState()
    out["est"] = (e.pose.copy(), e.neff, e.map_particle)
    out["w1"] = g.log_weights
    u = np.random.default_rng(9).uniform(0, 1, N + 1)
    out["pre1"] = g.particle_checksums()
    out["anc1"] = g.resampleParticles(u)
    out["post1"] = g.particle_checksums()
    out["sizes1"], out["maps1"] = g.get_maps()
    out["poses1"] = g.poses
    g.phdUpdateSynth(sc["Z"])
    out["anc2"] = g.resampleParticles()          # counter-based RNG
    out["sizes2"], out["maps2"] = g.get_maps()
    out["w2"] = g.log_weights
    if cphd:
        out["card2"] = g.cardinalities
    out["migrated"] = g.timings().migrated_in
    out["p2p"] = g.dist_p2p
    q.put((rank, lo, n, out))


@pytest.mark.parametrize("p2p,cphd", [(1, False), (0, False), (1, True), (0, True)])
def test_two_gpu_sharding_matches_oracle(p2p, cphd, monkeypatch):
    """p2p = 1: the resampling exchange pushes the offspring into the peer's buffers over NVLink (CUDA IPC window, one
    fused gather kernel); p2p = 0: the NCCL send/recv ring.  Both must reproduce the single-process oracle bit for bit."""
    _run_sharded(2, 301, False, p2p, cphd, monkeypatch)


@pytest.mark.parametrize("world,N,p2p,cphd", [(8, 1003, 1, False), (8, 1003, 0, False), (8, 1003, 1, True), (4, 1003, 1, False),
                                              (3, 301, 1, False)])
def test_many_gpu_sharding_skewed_weights_matches_oracle(world, N, p2p, cphd, monkeypatch):
    """The GPU-side analogue of tests/test_dist_cpu.py's N = 3..8 exchange-plan test: a particle count the world size does
    not divide, prior weights so skewed that most offspring descend from another rank's ancestors and some ranks serve
    nobody.  Needs `world` GPUs (gpurun --gpus 8)."""
    _run_sharded(world, N, True, p2p, cphd, monkeypatch)


def _run_sharded(world, N, skew, p2p, cphd, monkeypatch):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (run with gpurun --gpus %d)" % (world, world))
    monkeypatch.setenv("PHDSLAM_P2P", str(p2p))       # inherited by the spawned workers
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
    import phdslam_b200 as P
    from phdslam_b200 import scene as S
    from oracle import oracle as O
    uid = P.dist_unique_id()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, uid, q, cphd, N, skew)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    C, M = 16, 8
    cfg = _cfg(S, N, C, M, cphd)
    sc = _scene(S, N, C, M, skew)
    o = O.Oracle(cfg)
    S.load_scene(o, sc)
    o.phdPredict(np.float32([1.0, 0.05]))
    o.phdUpdateSynth(sc["Z"])
    oe = o.recoverSlamState()
    cat = lambda k: np.concatenate([r[3][k] for r in res])
    assert cat("w1").tobytes() == o.log_weights.tobytes()
    for r in res:
        pose, neff, mp_ = r[3]["est"]
        assert pose.tobytes() == oe.pose.tobytes() and neff == oe.neff and mp_ == oe.map_particle
    u = np.random.default_rng(9).uniform(0, 1, N + 1)
    oa = o.resampleParticles(u)
    assert (cat("anc1") == oa).all()
    # every offspring carries its ancestor's pose, map and cardinality, whichever GPU the ancestor lived on
    assert (cat("post1") == cat("pre1")[oa]).all()
    bounds = [N * r // world for r in range(world + 1)]
    own = lambda i: np.searchsorted(bounds, i, side="right") - 1
    crossed = int((own(np.arange(N)) != own(oa)).sum())
    if skew:
        assert crossed > N // 3, "the skewed weights must make most offspring cross GPUs (%d of %d)" % (crossed, N)
    os_, om = o.get_maps()
    assert (cat("sizes1") == os_).all() and cat("maps1").tobytes() == om.tobytes()
    assert cat("poses1").tobytes() == o.poses.tobytes()
    o.phdUpdateSynth(sc["Z"])
    oa2 = o.resampleParticles()
    assert (cat("anc2") == oa2).all()
    os2, om2 = o.get_maps()
    assert (cat("sizes2") == os2).all() and cat("maps2").tobytes() == om2.tobytes()
    assert cat("w2").tobytes() == o.log_weights.tobytes()
    if cphd:                                           # the cardinality distributions travel with the particles
        assert cat("card2").tobytes() == o.cardinalities.tobytes()
    assert sum(r[3]["migrated"] for r in res) > 0      # some offspring really crossed GPUs
    modes = {r[3]["p2p"] for r in res}
    assert len(modes) == 1                             # the ranks agree on the exchange path
    if p2p == 0:
        assert modes == {False}
    else:
        print("NVLink peer window mapped:", modes)
