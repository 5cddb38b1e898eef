"""Pins the MIXED feature model (static + constant-velocity features) of the CPU oracle to the
REFERENCE's own device code, run on the CPU.

oracle/ref_build.sh cuts predictMapKernelMixed (src/phdfilter.cu:910-963), computeBirth / computePreUpdate on Gaussian4D
(:244-521) and phdUpdateKernelMixed (:2323-2635) out of the reference, verbatim (one barrier added for the emulator, as for
phdUpdateKernel); computeMahalDist(Gaussian4D) / invert_matrix4 / ConstantVelocityMotionModel come with
src/device_math.cuh and the merge is the Gaussian4D instance of the phdUpdateMergeKernel template.  Their outputs on the cases
of tests/mixed_cases.py are committed as tests/golden/ref_mixed_golden.npz (generator:
tests/golden/make_ref_mixed_golden.py).  One particle per case: the reference kernel drops the particle's offset when it
reads the predicted weights (:2411,2437), so only particle 0 of a launch is what the authors meant.

Tolerance: 1e-4 relative (mixed_cases.assert_gaussians_close); discrete outputs -- the prune flags and the merged
component counts -- exactly.  Not reproduced, as in the reference: the "jump" features of predictMapKernelMixed are thrown
away by its host wrapper (:1015-1021) and so never reach the static map.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref as R
from conftest import GOLDEN
import mixed_cases as MC

GOLD = np.load(os.path.join(GOLDEN, "ref_mixed_golden.npz"))
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libphd_ref.so not built (needs /root/reference)")
CASES = list(MC.MIXED_CASES)


def gold(name):
    pre = name + "/"
    return {k[len(pre):]: GOLD[k] for k in GOLD.files if k.startswith(pre)}


@pytest.mark.parametrize("name", CASES)
def test_golden_inputs_reproducible(name):
    _, pose, sm, dm, Z = MC.build_case(name)
    g = gold(name)
    assert pose.tobytes() == g["pose"].tobytes() and sm.tobytes() == g["smap"].tobytes()
    assert dm.tobytes() == g["dmap"].tobytes() and Z.tobytes() == g["Z"].tobytes()


@needs_ref
@pytest.mark.parametrize("name", ["small", "labelled_vo"])
def test_golden_is_what_the_reference_kernels_produce(name):
    r = MC.reference_case(name)
    g = gold(name)
    for k in ("s_terms", "d_terms", "s_flags", "d_flags", "predicted", "d_merged", "mahal"):
        assert np.asarray(r[k]).tobytes() == g[k].tobytes(), k


@pytest.mark.parametrize("name", CASES)
def test_update_terms_of_both_maps(name):
    """phdUpdateKernelMixed: non-detection, detection and birth terms of the static and of the dynamic map, normalised by
    the shared per-measurement normaliser, and the particle's log-weight increment (schemes 0 and 1)."""
    cfg, pose, sm, dm, Z = MC.build_case(name)
    g = gold(name)
    st, dt, dlogw = O.mixed_terms(cfg, pose, sm, dm, Z)
    MC.assert_gaussians_close(st, g["s_terms"], name + " static terms")
    MC.assert_gaussians_close(dt, g["d_terms"], name + " dynamic terms")
    assert abs(dlogw - float(g["dlogw"])) <= 1e-4 * max(1.0, abs(float(g["dlogw"])))
    # prune flags (:2612-2633): the same terms survive, except where a weight sits within 1e-4 of the threshold
    for terms, flags in ((st, g["s_flags"]), (dt, g["d_flags"])):
        mine = terms["weight"] < cfg.min_feature_weight
        near = np.abs(terms["weight"] - cfg.min_feature_weight) < 1e-4 * cfg.min_feature_weight
        assert ((mine == (flags != 0)) | near).all()


@pytest.mark.parametrize("name", CASES)
def test_map_prediction(name):
    """predictMapKernelMixed + ConstantVelocityMotionModel::compute_prediction: constant-velocity mean and covariance,
    weight scaled by ps and the jump-Markov sigmoid of the speed."""
    cfg, _, _, dm, _ = MC.build_case(name)
    g = gold(name)
    MC.assert_gaussians_close(O.predict_features4(cfg, dm), g["predicted"], name + " predicted", wfloor=1e-9)


@pytest.mark.parametrize("name", CASES)
def test_mahalanobis_4d(name):
    """computeMahalDist(Gaussian4D, Gaussian4D): the reference inverts by cofactors (invert_matrix4), the oracle factors."""
    _, _, _, dm, _ = MC.build_case(name)
    g = gold(name)
    n = len(dm)
    mine = np.array([[O.mahalanobis4(dm[i], dm[j]) for j in range(n)] for i in range(n)], np.float32).reshape(n, n)
    assert MC.relerr(mine, g["mahal"], 1e-3) < 1e-4


@pytest.mark.parametrize("name", CASES)
def test_merge_of_the_dynamic_map(name):
    """phdUpdateMergeKernel<Gaussian4D> on the reference's own prune survivors: same clusters, same moments."""
    cfg, _, _, _, _ = MC.build_case(name)
    g = gold(name)
    cand = g["d_terms"][g["d_flags"] == 0]
    merged = O.merge4(cfg, cand)
    assert len(merged) == len(g["d_merged"])
    MC.assert_gaussians_close(merged, g["d_merged"], name + " merged dynamic map")


@pytest.mark.parametrize("name", CASES)
def test_whole_update_of_one_particle(name):
    """oracle_update with feature_model = 2 = the reference's kernels chained as phdUpdateSynth chains them (:3449-3462,
    :3703-3726): static map = merge of the static survivors, dynamic map = merge of the dynamic survivors (out-of-range
    dynamic features dropped)."""
    cfg, pose, sm, dm, Z = MC.build_case(name)
    g = gold(name)
    o = O.Oracle(cfg)
    o.poses = pose
    o.set_maps([len(sm)], sm)
    o.set_maps_dynamic([len(dm)], dm)
    o.phdUpdateSynth(Z)
    _, smap = o.get_maps()
    _, dmap = o.get_maps_dynamic()
    MC.assert_gaussians_close(dmap, g["d_merged"], name + " dynamic map after the update")
    MC.assert_gaussians_close(smap, g["s_merged"], name + " static map after the update")
    assert o.log_weights[0] == 0.0          # one particle: normalised


def test_out_of_range_dynamic_features_are_dropped_and_static_ones_kept():
    """:3713-3719 ("hack to kill off out-of-range dynamic features") against :3311-3318 for the static map."""
    cfg, pose, sm, dm, Z = MC.build_case("small")
    far_s = sm[:1].copy()
    far_s["mean"][0] = (60.0, 0.0)
    far_d = dm[:1].copy()
    far_d["mean"][0, :2] = (0.0, -70.0)
    o = O.Oracle(cfg)
    o.poses = pose
    o.set_maps([len(sm) + 1], np.concatenate([sm, far_s]))
    o.set_maps_dynamic([len(dm) + 1], np.concatenate([far_d, dm]))
    o.phdUpdateSynth(Z)
    _, smap = o.get_maps()
    _, dmap = o.get_maps_dynamic()
    g = gold("small")
    assert len(smap) == len(g["s_merged"]) + 1 and smap[-1].tobytes() == far_s[0].tobytes()
    assert len(dmap) == len(g["d_merged"]) and not (np.abs(dmap["mean"][:, 1] + 70.0) < 1.0).any()


def test_dynamic_feature_without_velocity_coupling_updates_like_a_static_one():
    """Known answer: a 4-D component whose position block equals a 2-D component's, with no position-velocity covariance,
    has the same innovation, likelihood and position update as the 2-D component (computePreUpdate 2-D vs 4-D)."""
    cfg, pose, sm, _, Z = MC.build_case("small")
    dm = np.zeros(len(sm), MC.G4)
    for i in range(len(sm)):
        c4 = np.zeros((4, 4), np.float32)
        c4[:2, :2] = sm["cov"][i].reshape(2, 2).T
        c4[2, 2], c4[3, 3] = 0.3, 0.2
        dm["cov"][i] = c4.T.reshape(-1)
        dm["mean"][i, :2] = sm["mean"][i]
        dm["weight"][i] = sm["weight"][i]
    st, dt, _ = O.mixed_terms(cfg, pose, sm, dm, Z)
    assert MC.relerr(dt["weight"], st["weight"], 1e-7) < 2e-5
    assert MC.relerr(dt["mean"][:, :2], st["mean"], 1e-3) < 2e-5
    n = len(sm)
    assert (dt["mean"][n:n + n * len(Z), 2:] == 0).all()          # no coupling: the velocity stays where it was
    pos_cov = dt["cov"].reshape(-1, 4, 4)[:, :2, :2].transpose(0, 2, 1).reshape(-1, 4)
    births = slice(n + n * len(Z), None)
    assert MC.relerr(pos_cov[:births.start], st["cov"][:births.start], 1e-5) < 1e-5
    assert (dt["cov"][births][:, 10] == cfg.cov_vx_birth).all() and (dt["cov"][births][:, 15] == cfg.cov_vy_birth).all()


def test_prediction_known_answers():
    cfg = MC.mixed_config(1, 4, std_ax_features=0.0, std_ay_features=0.0, tau=0.0, beta=1.0, ps=1.0)
    f = np.zeros(1, MC.G4)
    f["cov"][0] = np.eye(4, dtype=np.float32).reshape(-1)
    f["mean"][0] = (1.0, 2.0, 3.0, -4.0)
    f["weight"][0] = 0.5
    p = O.predict_features4(cfg, f)[0]
    dt = cfg.dt
    assert np.allclose(p["mean"], (1 + 3 * dt, 2 - 4 * dt, 3, -4), rtol=1e-6)
    c = p["cov"].reshape(4, 4)
    assert np.allclose(c[0, 0], 1 + dt * dt) and np.allclose(c[0, 2], dt) and np.allclose(c[2, 2], 1.0) and c[0, 1] == 0
    assert np.isclose(p["weight"], 0.5 / (1 + np.exp(-5.0)), rtol=1e-6)     # |v| = 5, tau = 0, beta = 1, ps = 1


def test_mahalanobis_4d_known_answer():
    a = np.zeros(1, MC.G4)
    b = np.zeros(1, MC.G4)
    a["cov"][0] = (np.eye(4) * 2).reshape(-1)
    b["cov"][0] = (np.eye(4) * 4).reshape(-1)
    b["mean"][0] = (3, 0, 0, 4)
    assert np.isclose(O.mahalanobis4(a[0], b[0]), 25.0 / 3.0, rtol=1e-6)
