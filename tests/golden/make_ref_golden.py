#!/usr/bin/env python
"""Generates tests/golden/ref_golden.npz from the REFERENCE's own kernels.

Run in a container that has /root/reference:   bash oracle/ref_build.sh && python tests/golden/make_ref_golden.py
oracle/_ref/libphd_ref.so = the reference's phdPredictKernel(Ackerman), computeInRangeKernel, preUpdateSynthKernel,
phdUpdateKernel, phdUpdateMergeKernel, computeMahalDist/Hellinger, resampleParticles and recoverSlamState, cut
verbatim from /root/reference/src and run through a CUDA execution-model emulator (oracle/ref_shim/cuda_emul.h).
The .npz stores the inputs as well, so that the checks do not depend on numpy's random streams.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cuda-phdslam_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref_cases as RC  # noqa: E402
from oracle import ref as R  # noqa: E402
from phdslam_b200 import scene as S  # noqa: E402


def main():
    out = {}
    names = list(RC.UPDATE_CASES) + [RC.LABELED_CASE[0]]
    for name in names:
        cfg, sc = RC.build_case(name)
        R.set_config(cfg)
        cls, n_in, n_near = R.in_range(sc["maps"], sc["sizes"], sc["poses"])
        terms, flags, dlw = R.update_terms(sc["poses"], sc["maps"][cls == 1], n_in, sc["Z"])
        sizes, maps, lw = R.update(sc["poses"], sc["sizes"], sc["maps"], sc["log_weights"], sc["Z"])
        e, k = R.recover(lw, sc["poses"])
        for key, val in (("in_poses", sc["poses"]), ("in_sizes", sc["sizes"]), ("in_maps", sc["maps"]),
                         ("in_logw", sc["log_weights"]), ("Z", sc["Z"]), ("cls", cls), ("n_in", n_in), ("n_near", n_near),
                         ("terms", terms), ("prune_flags", flags), ("dlogw", dlw), ("out_sizes", sizes), ("out_maps", maps),
                         ("out_logw", lw), ("expected_pose", np.array([e])), ("map_particle", np.int32(k)),
                         ("neff", np.float32(R.neff(lw)))):
            out["%s/%s" % (name, key)] = val
        print("%-14s particles %d  in-range %s  merged sizes %s" % (name, len(sizes), n_in.tolist(), sizes.tolist()))

    # predict: both motion models
    rng = np.random.Generator(np.random.Philox(99))
    n = 64
    poses = np.zeros(n, RC.P.POSE_DTYPE)
    for f in poses.dtype.names:
        poses[f] = rng.normal(0, 2.0, n)
    poses["ptheta"] = rng.uniform(-np.pi, np.pi, n)
    for mt, nd in ((1, 2), (0, 3)):
        cfg = RC.predict_config(mt, n)
        R.set_config(cfg)
        draws = rng.normal(size=(n, nd))
        noise = RC.ackerman_noise(cfg, draws) if mt == 1 else RC.cv_noise(cfg, draws)
        control = np.float32([3.1, -0.12])
        out["predict%d/poses" % mt] = poses
        out["predict%d/draws" % mt] = draws
        out["predict%d/control" % mt] = control
        out["predict%d/out" % mt] = R.predict(poses, control, noise)

    # resampling: HEAD's stratified resampleParticles on a few weight vectors
    for i, (n, conc) in enumerate(((64, 0.2), (257, 1.0), (1000, 0.05), (16, 5.0))):
        lw = np.log(np.maximum(rng.dirichlet(np.ones(n) * conc), 1e-30)).astype(np.float32)
        lw = (lw - np.log(np.sum(np.exp(lw.astype(np.float64))))).astype(np.float32)
        u = rng.uniform(size=n + 1)
        idx, nlw = R.resample(lw, u)
        out["resample%d/logw" % i] = lw
        out["resample%d/uniforms" % i] = u
        out["resample%d/idx" % i] = idx
        out["resample%d/new_logw" % i] = nlw
    path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
