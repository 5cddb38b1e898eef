#!/usr/bin/env python
"""Extracts the simulation truth of the reference's bundled scenes into small fixtures that can travel to the GPU box
(/root/reference does not exist there).  Source: matlab/simData2_{ackerman,cv}.mat (`sim.traj`, `sim.groundTruth(k).loc`,
the structures matlab/computeBatchResults.m:67-74 and python/batch_analyze.py:16-40 score against).

    python tests/golden/make_truth.py        # writes tests/golden/truth_{ackerman,cv}.npz

traj   [n_steps][3 or 6] float32  true vehicle state per step
loc    [sum n_k][2]      float32  landmarks in the field of view history at step k, concatenated over k
off    [n_steps+1]       int32    loc[off[k]:off[k+1]] is the true map of step k
controls [n_steps-1][2], control_dt [n_steps-1]   (Ackerman scene) the noise-free controls of the trajectory
"""
import os

import numpy as np
import scipy.io as sio

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PHD_REFERENCE", "/root/reference")

for name in ("ackerman", "cv"):
    sim = sio.loadmat(os.path.join(REF, "matlab", "simData2_%s.mat" % name), squeeze_me=True, struct_as_record=False)["sim"]
    traj = np.asarray(sim.traj, dtype=np.float32).T
    locs = [np.atleast_2d(np.asarray(g.loc, dtype=np.float32)).reshape(2, -1).T for g in sim.groundTruth]
    off = np.zeros(len(locs) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(x) for x in locs])
    extra = {}
    if name == "ackerman":   # the controls the trajectory was driven with: sim.control(k).u = (speed, steering angle), .dt
        extra["controls"] = np.array([np.asarray(c.u, dtype=np.float32) for c in sim.control], dtype=np.float32)
        extra["control_dt"] = np.array([float(c.dt) for c in sim.control], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "truth_%s.npz" % name), traj=traj, loc=np.concatenate(locs, 0), off=off, **extra)
    print(name, traj.shape, off[-1])
