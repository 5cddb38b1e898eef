"""CPHD on the sm_100a path against the CPU oracle on identical inputs.
Component counts bit-exact, floats within 1e-4 relative, log cardinality distributions within 1e-4 absolute
(both sides implement the same canonical arithmetic, so bit-identity is expected and reported)."""
import math

import numpy as np
import pytest
from scipy.special import gammaln

import phdslam_b200 as P
from phdslam_b200 import scene as S
from oracle import oracle as O
import ref_cases as RC

pytestmark = pytest.mark.gpu


def poisson_card(lam, n1, Pn):
    n = np.arange(n1)
    return np.tile((n * math.log(lam) - lam - gammaln(n + 1.0)).astype(np.float32), (Pn, 1))


def pair(cfg, sc, card=None):
    g, o = P.PhdSlam(cfg), O.Oracle(cfg)
    for f in (g, o):
        S.load_scene(f, sc)
        if card is not None:
            f.cardinalities = card
    return g, o


@pytest.mark.parametrize("Pn,C,M,N1,near,far", [(8, 24, 9, 64, 2, 2), (4, 70, 50, 256, 3, 0), (6, 0, 5, 32, 0, 3),
                                                (3, 130, 100, 257, 0, 0), (5, 10, 1, 16, 1, 1), (2, 40, 37, 1000, 0, 0)])
def test_cphd_update_matches_oracle(Pn, C, M, N1, near, far):
    cfg = S.scene_config(Pn, C, M, max_components=512, filter_type=1, max_cardinality=N1 - 1)
    sc = S.make_scene(Pn, C, M, seed=41 + C, n_near=near, n_far=far)
    lam = max(float(sc["maps"]["weight"][:C + near + far].sum()), 0.5)
    g, o = pair(cfg, sc, poisson_card(min(lam, 0.8 * N1), N1, Pn))
    # dense terms (query: state not advanced)
    gt, gn, gd = g.update_terms(sc["Z"])
    ot, on, od = o.update_terms(sc["Z"])
    assert (gn == on).all()
    RC.assert_gaussians_close(gt, ot, "CPHD dense terms")
    RC.close(gd, od, "CPHD particle log-weight increment", atol=1e-5)
    assert g.cardinalities.tobytes() == o.cardinalities.tobytes(), "query advanced the cardinality"
    # full update
    g.phdUpdateSynth(sc["Z"])
    o.phdUpdateSynth(sc["Z"])
    gs, gm = g.get_maps()
    os_, om = o.get_maps()
    RC.assert_maps_close(gs, gm, os_, om, "CPHD maps")
    RC.close(g.log_weights, o.log_weights, "log-weights", atol=1e-6)
    gc, oc = g.cardinalities, o.cardinalities
    live = oc > -80.0
    assert np.abs(gc[live] - oc[live]).max() <= 1e-4
    assert np.allclose(np.exp(gc.astype(np.float64)).sum(1), 1.0, atol=3e-3)
    print("bit-identical: maps %s, cardinality %s" % (gm.tobytes() == om.tobytes(), gc.tobytes() == oc.tobytes()))


def test_cphd_step_loop_with_resampling():
    """several CPHD steps incl. predict and resampling (the cardinality rows travel with their particle)"""
    Pn, C, M = 64, 20, 8
    cfg = S.scene_config(Pn, C, M, max_components=256, filter_type=1, max_cardinality=63, resample_threshold=1.0, seed="9")
    sc = S.make_scene(Pn, C, M, seed=77, n_near=1, n_far=1)
    g, o = pair(cfg, sc, poisson_card(12.0, 64, Pn))
    for k in range(4):
        ge, gr = g.step(k, np.float32([1.0, 0.02]), sc["Z"])
        oe, orr = o.step(k, np.float32([1.0, 0.02]), sc["Z"])
        assert gr == orr
        assert (g.map_sizes == o.map_sizes).all(), "step %d" % k
        assert (g.resample_idx == o.resample_idx).all(), "step %d" % k
        RC.close(ge.neff, oe.neff, "nEff")
    gc, oc = g.cardinalities, o.cardinalities
    live = oc > -80.0
    assert np.abs(gc[live] - oc[live]).max() <= 1e-4
    gs, gm = g.get_maps()
    os_, om = o.get_maps()
    RC.assert_maps_close(gs, gm, os_, om, "CPHD maps after loop")
