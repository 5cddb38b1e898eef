#!/usr/bin/env python
"""BASELINE.json configs[0] / configs[1] in full: the run_synth loop over the bundled text data
(tests/golden/data = the reference's matlab/*.txt), timed per step.

    python profiles/text_configs.py --impl oracle [--threads N]      # the CPU oracle
    python profiles/text_configs.py --impl ours                      # the CUDA library (needs a GPU)
    python profiles/text_configs.py --impl both                      # both arms on the same box, one after the other

configs[0]: measurements_synth_ackerman.txt + controls_synth.txt, cfg/config.cfg sensor / filter values, 256 particles
configs[1]: measurements_synth_cv.txt, constant-velocity motion, 4096 particles
One JSON line per (config, arm): steps, mean / median ms per step, steps/s, (component, measurement) pairs/s (all map
components x measurements of each step: an upper bound of the GM-PHD updates, whose exact count needs the in-range split)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import phdslam_b200 as P  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(GOLDEN, "data")


def configs():
    c0 = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    c0.set(n_particles=256, max_components=256, seed="7")
    z0 = P.load_measurements(os.path.join(DATA, "measurements_synth_ackerman.txt"))
    u0 = P.load_controls(os.path.join(DATA, "controls_synth.txt"))
    c1 = P.load_config(os.path.join(GOLDEN, "config_ackerman.cfg"))
    c1.set(motion_type=0, n_particles=4096, max_components=256, initial_vx=2.0, initial_vyaw=0.2, acc_x=0.5, acc_y=0.5,
           acc_yaw=0.087, dt=0.02, max_range=10.0, std_range=1.0, std_bearing=0.0349, seed="3")
    z1 = P.load_measurements(os.path.join(DATA, "measurements_synth_cv.txt"))
    return [("configs[0] ackerman text data, 256 particles", c0, z0, u0),
            ("configs[1] constant-velocity text data, 4096 particles", c1, z1, None)]


def run(name, cfg, Z, U, impl, threads, max_steps):
    if impl == "ours":
        f = P.PhdSlam(cfg, device=0)
    else:
        from oracle import oracle as O
        if O._lib is None:
            O.load(O.build_native())           # -O3 -march=native on the host that is timed
        f = O.Oracle(cfg, threads=threads)
    n_steps = min(len(Z), max_steps) if U is None else min(len(Z), len(U) + 1, max_steps)
    ms, pairs = [], 0
    for k in range(n_steps):
        u = U[k - 1] if (U is not None and k > 0) else None
        # (component, measurement) pairs of this step, counted on ALL map components (in range or not): an upper bound of
        # the GM-PHD updates, read outside the timed region
        pairs += int(f.map_sizes.sum()) * len(Z[k])
        t0 = time.perf_counter()
        f.step(k, u, Z[k])
        if impl == "ours":
            f.synchronize()
        ms.append((time.perf_counter() - t0) * 1e3)
    line = {"config": name, "impl": impl, "threads": threads if impl == "oracle" else None, "steps": n_steps,
            "particles": int(cfg.n_particles), "ms_per_step_mean": float(np.mean(ms)), "ms_per_step_median": float(np.median(ms)),
            "steps_per_s": float(1e3 / np.mean(ms)),
            "component_measurement_pairs_per_s": float(pairs / (np.sum(ms) * 1e-3)), "final_mean_map_size": float(np.mean(f.map_sizes)),
            "measurements_per_step_mean": float(np.mean([len(z) for z in Z[:n_steps]])), "host_cores": os.cpu_count(),
            "slowest_steps": [(int(k), round(float(ms[k]), 3), len(Z[k])) for k in np.argsort(ms)[::-1][:8]],
            "ms_per_step_mean_without_step0": float(np.mean(ms[1:])) if len(ms) > 1 else None}
    print(json.dumps(line))
    sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="oracle", choices=["oracle", "ours", "both"])
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--max-steps", type=int, default=1 << 30)
    ap.add_argument("--only", type=int, default=-1)
    a = ap.parse_args()
    for i, (name, cfg, Z, U) in enumerate(configs()):
        if a.only >= 0 and i != a.only:
            continue
        for impl in (("ours", "oracle") if a.impl == "both" else (a.impl,)):
            run(name, cfg, Z, U, impl, a.threads, a.max_steps)
            if impl == "oracle" and a.impl == "both" and a.threads > 1:
                run(name, cfg, Z, U, impl, 1, min(a.max_steps, 100))      # single-thread figure on a prefix of the run


if __name__ == "__main__":
    main()
