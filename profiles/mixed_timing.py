#!/usr/bin/env python
"""Timing of the mixed feature model (feature_model = 2) at a synthetic shape: P particles x Cs static x Cd dynamic in-range
components x M measurements, the static side exactly the scene of bench.py's workloads (phdslam_b200.scene.make_scene).
Prints one JSON line: the step time and its phases (CUDA events inside libphdslam.so; `dynamic` = dyn_pre_kernel +
dyn_update_kernel), the same step on the CPU oracle for a particle subset, and a parity check of that subset (bit-exact).

  python profiles/mixed_timing.py [--particles 65536] [--static 128] [--dynamic 16] [--meas 50] [--steps 10] [--warmup 3]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mixed_cases as MC  # noqa: E402
import phdslam_b200 as P  # noqa: E402
from phdslam_b200 import scene as S  # noqa: E402


def dynamic_maps(n, nd, seed):
    rng = np.random.default_rng(seed)
    base = MC.dynamic_features(rng, nd)
    dm = np.tile(base, n)
    dm["mean"] += rng.normal(0, 0.05, dm["mean"].shape).astype(np.float32)
    return np.full(n, nd, np.int32), dm, base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=65536)
    ap.add_argument("--static", type=int, default=128)
    ap.add_argument("--dynamic", type=int, default=16)
    ap.add_argument("--meas", type=int, default=50)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--oracle-particles", type=int, default=512)
    a = ap.parse_args()
    n, Cs, Cd, M = a.particles, a.static, a.dynamic, a.meas
    kw = dict(MC.MIXED_OVERRIDES)
    kw.update(max_components_dynamic=max(64, (Cd + M + 37) & ~7), seed="3")   # the map grows by the births of a step
    cfg = S.scene_config(n, Cs, M, max_components=2 * Cs, **kw)
    sc = S.make_scene(n, Cs, M, seed=0)
    dsz, dm, base = dynamic_maps(n, Cd, 1)
    # a third of the measurements come from the dynamic features
    Z = np.array(sc["Z"], np.float32).reshape(M, -1)
    rng = np.random.default_rng(2)
    for m in range(0, M, 3):
        q = base["mean"][rng.integers(Cd), :2]
        Z[m, 0], Z[m, 1] = np.hypot(*q) + rng.normal(0, 0.25), np.arctan2(q[1], q[0]) + rng.normal(0, 0.0087)
    g = P.PhdSlam(cfg)
    S.load_scene(g, sc)
    g.set_maps_dynamic(dsz, dm)
    g.snapshot()
    u = np.float32([1.0, 0.02])
    phases = {k: 0.0 for k in ("predict", "update", "merge", "dynamic", "weights", "estimate", "resample")}
    wall = []
    for k in range(a.warmup + a.steps):
        g.restore()
        g.synchronize()
        t0 = time.perf_counter()
        g.step(1, u, Z)
        g.synchronize()
        t1 = time.perf_counter()
        if k >= a.warmup:
            t = g.timings()
            wall.append((t1 - t0) * 1e3)
            for name in phases:
                phases[name] += getattr(t, name + "_ms") / a.steps
    dyn_sizes = g.map_sizes_dynamic
    # parity + CPU time on a subset
    from oracle import oracle as O
    ns = min(a.oracle_particles, n)
    cfg_s = S.scene_config(ns, Cs, M, max_components=2 * Cs, **kw)
    sub = dict(poses=sc["poses"][:ns], log_weights=np.full(ns, -np.log(ns), np.float32), sizes=sc["sizes"][:ns],
               maps=sc["maps"][:int(sc["sizes"][:ns].sum())])
    o = O.Oracle(cfg_s, threads=os.cpu_count())
    S.load_scene(o, sub)
    o.set_maps_dynamic(dsz[:ns], dm[:int(dsz[:ns].sum())])
    t0 = time.perf_counter()
    o.phdPredict(u)
    o.phdUpdateSynth(Z)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    gs = P.PhdSlam(cfg_s)
    S.load_scene(gs, sub)
    gs.set_maps_dynamic(dsz[:ns], dm[:int(dsz[:ns].sum())])
    gs.phdPredict(u)
    gs.phdUpdateSynth(Z)
    same = (gs.get_maps_dynamic()[1].tobytes() == o.get_maps_dynamic()[1].tobytes() and
            gs.get_maps()[1].tobytes() == o.get_maps()[1].tobytes() and gs.log_weights.tobytes() == o.log_weights.tobytes())
    upd = float(n) * (Cs + Cd) * M
    print(json.dumps({
        "workload": "synthetic_%dx(%d+%d)x%d_mixed" % (n, Cs, Cd, M), "steps": a.steps, "warmup": a.warmup,
        "ms_per_step_wall": float(np.mean(wall)), "phase_ms": {k: round(v, 4) for k, v in phases.items()},
        "updates_per_s": upd / (float(np.mean(wall)) * 1e-3), "unit": "GM-PHD updates/s (static + dynamic terms)",
        "dynamic_map_sizes_after": {"mean": float(dyn_sizes.mean()), "max": int(dyn_sizes.max())},
        "cpu_oracle": {"particles": ns, "ms": cpu_ms, "threads": os.cpu_count(),
                       "updates_per_s": float(ns) * (Cs + Cd) * M / (cpu_ms * 1e-3)},
        "parity_subset": {"particles": ns, "bit_identical_to_oracle": bool(same)},
    }))


if __name__ == "__main__":
    main()
