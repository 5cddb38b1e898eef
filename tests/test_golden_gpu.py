"""The sm_100a path (through the C-ABI) against tests/golden/ref_golden.npz -- outputs of the REFERENCE's own
kernels (compiled for the CPU by oracle/ref_build.sh; generator: tests/golden/make_ref_golden.py).
Bar: component counts, in-range counts and ancestor indices bit-exact; floats within 1e-4 relative (ref_cases.py
states the absolute floors used next to the relative bound)."""
import os

import numpy as np
import pytest

import phdslam_b200 as P
from phdslam_b200 import scene as S
from conftest import GOLDEN
import ref_cases as RC

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(GOLDEN, "ref_golden.npz"))
ALL_CASES = list(RC.UPDATE_CASES) + [RC.LABELED_CASE[0]]


def gold(name, key):
    return GOLD["%s/%s" % (name, key)]


def filter_for(name):
    cfg, _ = RC.build_case(name)
    if name == RC.LABELED_CASE[0]:
        cfg.set(measurement_fields=3)
    g = P.PhdSlam(cfg)
    g.poses = gold(name, "in_poses")
    g.log_weights = gold(name, "in_logw")
    g.set_maps(gold(name, "in_sizes"), gold(name, "in_maps"))
    return cfg, g


@pytest.mark.parametrize("name", ALL_CASES)
def test_cuda_update_terms_vs_reference_kernels(name):
    cfg, g = filter_for(name)
    terms, n_in, dlw = g.update_terms(gold(name, "Z"))
    assert (n_in == gold(name, "n_in")).all()
    RC.assert_gaussians_close(terms, gold(name, "terms"), name + " terms")
    RC.close(dlw, gold(name, "dlogw"), name + " particle log-weight increment", atol=2e-5)


@pytest.mark.parametrize("name", ALL_CASES)
def test_cuda_full_update_vs_reference_kernels(name):
    cfg, g = filter_for(name)
    g.phdUpdateSynth(gold(name, "Z"))
    sizes, maps = g.get_maps()
    RC.assert_maps_close(sizes, maps, gold(name, "out_sizes"), gold(name, "out_maps"), name)
    RC.close(g.log_weights, gold(name, "out_logw"), name + " log-weights", atol=2e-5)
    e = g.recoverSlamState()
    ge = gold(name, "expected_pose")[0]
    RC.close(e.pose, np.array([ge[f] for f in ge.dtype.names]), name + " expected pose", atol=1e-6)
    assert e.map_particle == int(gold(name, "map_particle"))
    RC.close(e.neff, gold(name, "neff"), name + " nEff")


@pytest.mark.parametrize("mt", [1, 0])
def test_cuda_predict_vs_reference_kernels(mt):
    cfg = RC.predict_config(mt)
    g = P.PhdSlam(cfg)
    g.poses = GOLD["predict%d/poses" % mt]
    g.phdPredict(GOLD["predict%d/control" % mt], draws=GOLD["predict%d/draws" % mt])
    got, want = g.poses, GOLD["predict%d/out" % mt]
    for f in got.dtype.names:
        RC.close(got[f], want[f], "predict pose." + f, atol=RC.ATOL_POS)


@pytest.mark.parametrize("i", range(4))
def test_cuda_resample_vs_reference(i):
    lw, u, idx = GOLD["resample%d/logw" % i], GOLD["resample%d/uniforms" % i], GOLD["resample%d/idx" % i]
    cfg = S.scene_config(len(lw), 1, 1)
    g = P.PhdSlam(cfg)
    g.log_weights = lw
    anc = g.resampleParticles(uniforms=u)
    assert (anc == idx).all()
    RC.close(g.log_weights, GOLD["resample%d/new_logw" % i], "weights after resampling")


# ---- CPHD: the sm_100a path against the reference's own CPHD kernels (tests/golden/ref_cphd_golden.npz; what they are
# and how they were run: tests/test_cphd_ref_pin.py, tests/ref_cases.py CPHD_CASES) ----
CPHD_GOLD = np.load(os.path.join(GOLDEN, "ref_cphd_golden.npz"))


@pytest.mark.parametrize("name", list(RC.CPHD_CASES))
def test_cuda_cphd_update_vs_reference_kernels(name):
    cfg, sc, _ = RC.build_cphd_case(name)
    pre = name + "/"
    ref = {k[len(pre):]: CPHD_GOLD[k] for k in CPHD_GOLD.files if k.startswith(pre)}
    g = P.PhdSlam(cfg)
    g.poses = ref["in_poses"]
    g.log_weights = sc["log_weights"]
    g.set_maps(ref["in_sizes"], ref["in_maps"])
    g.cardinalities = ref["cn_predict"]
    terms, n_in, dlw = g.update_terms(ref["Z"])
    lw0 = g.log_weights.astype(np.float64)
    g.phdUpdateSynth(ref["Z"])
    RC.assert_cphd_matches_reference(ref, terms, n_in, dlw, g.cardinalities, name)
    w = lw0 + ref["ip0"].astype(np.float64)
    w -= np.log(np.sum(np.exp(w - w.max()))) + w.max()
    RC.close(g.log_weights, w, name + " particle log-weights", atol=2e-5)
