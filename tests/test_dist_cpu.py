"""N > 1 host logic on CPU (gloo, world_size 2): the planning of the global-resampling exchange
(phdslam_plan_migration / phdslam_resample_threshold, pure host code in libphdslam.so) reproduces the
single-process canonical resampler of the oracle when particles are sharded over ranks, and the order-independent
fixed-point weight statistics combine to the same bits for any rank count."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, seed, mode, q):
    sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import phdslam_b200 as P
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        w = rng.normal(0, 3, n)
        w = (w - np.log(np.exp(w).sum())).astype(np.float32)
        payload = np.arange(n, dtype=np.float32) * 10           # stands for the particle state
        u = rng.uniform(0, 1, n + 1)
        lo, hi = n * rank // world, n * (rank + 1) // world
        # local integer weights and totals (Q40 of the canonical exp)
        q40 = np.rint(O.detmath("exp", w[lo:hi]).astype(np.float64) * float(1 << 40)).astype(np.uint64)
        tot = torch.tensor([int(q40.sum())], dtype=torch.int64)
        gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, tot)
        totals = np.array([int(t[0]) for t in gathered], dtype=np.uint64)
        bounds = P.plan_migration(totals, n, uniforms=u, resample_mode=mode)
        total = int(totals.sum())
        base = int(totals[:rank].sum())
        cdf = base + np.cumsum(q40.astype(object))               # inclusive, python ints
        # serve: offspring whose ancestor is mine, grouped by destination rank
        out = {}
        for d in range(world):
            dlo, dhi = n * d // world, n * (d + 1) // world
            a, b = max(bounds[rank], dlo), min(bounds[rank + 1], dhi)
            anc = []
            for j in range(a, b):
                R = P.resample_threshold(j, n, total, uniforms=u, resample_mode=mode)
                i = int(np.searchsorted(np.array(cdf, dtype=object), R, side="right"))
                assert 0 <= i < hi - lo, "plan sent an offspring to the wrong rank"
                anc.append(lo + i)
            out[d] = (a, np.array(anc, dtype=np.int64))
        new_anc = np.full(hi - lo, -1, dtype=np.int64)
        new_payload = np.zeros(hi - lo, dtype=np.float32)
        a, anc = out[rank]
        new_anc[a - lo:a - lo + len(anc)] = anc
        new_payload[a - lo:a - lo + len(anc)] = payload[anc]
        for k in range(1, world):
            d, s = (rank + k) % world, (rank - k) % world
            a, anc = out[d]
            cnt_in = max(min(bounds[s + 1], hi) - max(bounds[s], lo), 0)
            reqs = []
            if len(anc):
                reqs.append(dist.isend(torch.from_numpy(anc.copy()), d))
                reqs.append(dist.isend(torch.from_numpy(payload[anc].copy()), d))
            if cnt_in:
                ra = torch.zeros(cnt_in, dtype=torch.int64)
                rp = torch.zeros(cnt_in, dtype=torch.float32)
                dist.recv(ra, s)
                dist.recv(rp, s)
                first = max(bounds[s], lo) - lo
                new_anc[first:first + cnt_in] = ra.numpy()
                new_payload[first:first + cnt_in] = rp.numpy()
            for r in reqs:
                r.wait()
        # order-independent weight statistics: per-rank fixed-point partial sums add up to the global bits
        q36 = np.rint(O.detmath("exp", (w[lo:hi] - w.max()).astype(np.float32)).astype(np.float64) * float(1 << 36)).astype(np.uint64)
        s36 = torch.tensor([int(q36.sum())], dtype=torch.int64)
        dist.all_reduce(s36)
        q.put((rank, new_anc, new_payload, int(s36[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", [0, 1])
def test_sharded_resampling_matches_single_process(mode):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
    import phdslam_b200 as P
    from phdslam_b200 import scene as S
    from oracle import oracle as O
    n, seed, world = 203, 17 + mode, 2                           # odd count: uneven shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000) + mode
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, seed, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    anc = np.concatenate([r[1] for r in res])
    pay = np.concatenate([r[2] for r in res])
    # single-process oracle on the same weights / uniforms
    rng = np.random.default_rng(seed)
    w = rng.normal(0, 3, n)
    w = (w - np.log(np.exp(w).sum())).astype(np.float32)
    u = rng.uniform(0, 1, n + 1)
    cfg = S.scene_config(n, 1, 1, resample_mode=mode)
    o = O.Oracle(cfg)
    o.log_weights = w
    poses = np.zeros(n, dtype=P.POSE_DTYPE)
    poses["px"] = np.arange(n, dtype=np.float32) * 10
    o.poses = poses
    oa = o.resampleParticles(u)
    assert (anc == oa).all()                                     # bit-exact ancestors for any rank count
    assert (pay == o.poses["px"]).all()
    # fixed-point logsumexp partial sums: same integer on every rank, equal to the single-process sum
    q36 = np.rint(O.detmath("exp", (w - w.max()).astype(np.float32)).astype(np.float64) * float(1 << 36)).astype(np.uint64)
    assert res[0][3] == res[1][3] == int(q36.sum())


def test_plan_bounds_properties():
    sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
    import phdslam_b200 as P
    totals = np.array([0, 5 << 38, 0, 3 << 38], dtype=np.uint64)   # ranks with zero weight own no ancestors
    b = P.plan_migration(totals, 1000, uniforms=None, seed=5)
    assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all()
    assert b[1] == 0 and b[2] == b[3]                              # empty ranks get empty intervals
    assert abs(int(b[2]) - 625) <= 2                               # 5/8 of the offspring descend from rank 1


@pytest.mark.parametrize("world,n,mode,skew", [(3, 101, 0, 0), (4, 1000, 0, 1), (8, 997, 0, 0), (8, 1024, 1, 2), (5, 64, 0, 2)])
def test_exchange_plan_for_larger_worlds(world, n, mode, skew):
    """The exchange the GPUs run at N = 4 / 8 (one interval of offspring per (serving rank, owning rank) pair, derived by
    both ends from the all-gathered totals alone) simulated in one process with the library's own planning functions:
    every offspring slot is served exactly once, by the rank that owns its ancestor, and the ancestors equal the
    single-process oracle's.  skew 1: almost all the weight on one rank; skew 2: some ranks carry no weight at all."""
    sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
    sys.path.insert(0, ROOT)
    import phdslam_b200 as P
    from phdslam_b200 import scene as S
    from oracle import oracle as O
    rng = np.random.default_rng(1000 * world + n + mode)
    w = rng.normal(0, 2, n)
    lo = [n * r // world for r in range(world + 1)]
    if skew == 1:
        w[lo[1]:lo[2]] += 12.0
    if skew == 2:
        for r in range(0, world, 2):
            w[lo[r]:lo[r + 1]] = -200.0                          # exp underflows to an integer weight of 0
    w = (w - np.log(np.exp(w).sum())).astype(np.float32)
    u = rng.uniform(0, 1, n + 1)
    q40 = np.rint(O.detmath("exp", w).astype(np.float64) * float(1 << 40)).astype(np.uint64)
    totals = np.array([int(q40[lo[r]:lo[r + 1]].astype(object).sum()) for r in range(world)], dtype=np.uint64)
    total = int(totals.astype(object).sum())
    bounds = P.plan_migration(totals, n, uniforms=u, resample_mode=mode)
    assert bounds[0] == 0 and bounds[-1] == n and (np.diff(bounds) >= 0).all()
    served = np.zeros(n, dtype=np.int64)
    anc = np.full(n, -1, dtype=np.int64)
    for me in range(world):                                      # what rank `me` does (phdslam_resample, world > 1)
        base = int(totals[:me].astype(object).sum())
        cdf = base + np.cumsum(q40[lo[me]:lo[me + 1]].astype(object))
        for k in range(world):                                   # k = 0: its own offspring; k > 0: the ring of shifts
            d = (me + k) % world
            a, b = max(bounds[me], lo[d]), min(bounds[me + 1], lo[d + 1])
            for j in range(a, b):
                R = P.resample_threshold(j, n, total, uniforms=u, resample_mode=mode)
                i = int(np.searchsorted(np.array(cdf, dtype=object), R, side="right"))
                assert 0 <= i < lo[me + 1] - lo[me], "offspring %d planned onto rank %d, which does not own its ancestor" % (j, me)
                served[j] += 1
                anc[j] = lo[me] + i
    assert (served == 1).all()
    cfg = S.scene_config(n, 1, 1, resample_mode=mode)
    o = O.Oracle(cfg)
    o.log_weights = w
    assert (anc == o.resampleParticles(u)).all()
