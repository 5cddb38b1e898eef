"""Known-answer tests of the oracle's CPHD multi-object terms.

CPHD is dead code in the reference (every CPHD kernel is commented out at HEAD, the live older pipeline is
in the un-built src/phdfilter.cu.bak), so there is nothing to run it against: parity for this row is pinned by
  (1) a float64 numpy transcription of the LITERAL formulas of the commented kernels (computeEsfKernel
      src/phdfilter.cu:1524-1618 via numpy.poly, computePsiKernel :1626-1769 with its O(N M^2) Psi1d loop,
      cphdUpdateKernel :1780-1822, cardinalityPredictKernel :867-888 + binomial births .bak:779-791),
  (2) properties of the CPHD recursion (Vo, Vo & Cantoni 2007): the updated cardinality is a distribution; with a
      Poisson predicted cardinality the weights reduce to the PHD update.
"""
import math

import numpy as np
import pytest
from scipy.special import gammaln, logsumexp

from phdslam_b200 import scene as S
from oracle import oracle as O


def literal_cphd(cfg, w, pd, Sm, prior, do_predict=True):
    """float64, literal: Psi0/Psi1/Psi1d per n exactly as computePsiKernel loops them"""
    w, pd, Sm, prior = (np.asarray(a, np.float64) for a in (w, pd, Sm, prior))
    M, N1 = len(Sm), len(prior)
    N = N1 - 1
    wb, cr, cd = float(cfg.birth_weight), float(cfg.clutter_rate), float(cfg.clutter_density)
    lf = gammaln(np.arange(max(N, M) + 2) + 1.0)
    logC = lambda n, k: lf[n] - lf[k] - lf[n - k] if 0 <= k <= n else -np.inf   # noqa: E731
    if do_predict:
        pb = np.array([logC(M, k) + k * math.log(wb) + (M - k) * math.log1p(-wb) for k in range(M + 1)])
        pm = np.full(N1, -np.inf)
        for n in range(N1):
            terms = [pb[n - j] + prior[j] for j in range(n + 1) if n - j <= M]
            pm[n] = logsumexp(terms)
    else:
        pm = prior
    lam = (Sm + wb) * cr / cd
    esf = np.abs(np.poly(lam))                       # |coefficients| = e_0..e_M (Vieta)
    esfd = [np.atleast_1d(np.abs(np.poly(np.delete(lam, m)))) for m in range(M)]
    q = np.sum(w * (1 - pd))
    W = np.sum(w) + M * wb
    lq = math.log(q) if q > 0 else -np.inf
    lW = math.log(W)
    fact = lf                                        # dev_factorial (log)
    clutter = lambda k: k * math.log(cr) - cr - lf[k]   # dev_cn_clutter, :735-737  # noqa: E731

    def pw(k, x):
        return 0.0 if k == 0 else k * x

    psi0 = np.full(N1, -np.inf)
    psi1 = np.full(N1, -np.inf)
    psi1d = np.full((M, N1), -np.inf)
    for n in range(N1):
        t0, t1 = [], []
        for j in range(min(n, M) + 1):
            aux = fact[M - j] + clutter(M - j) + math.log(esf[j]) - pw(n, lW)
            t0.append(aux + logC(n, j) + fact[j] + pw(n - j, lq))
            if j + 1 <= n:
                t1.append(aux + logC(n, j + 1) + fact[j + 1] + pw(n - j - 1, lq))
        psi0[n] = logsumexp(t0)
        if t1:
            psi1[n] = logsumexp(t1)
        for m in range(M):
            td = []
            for j in range(min(M - 1, n) + 1):
                if j + 1 <= n:
                    aux = fact[M - 1 - j] + clutter(M - 1 - j) + math.log(esfd[m][j]) - pw(n, lW)
                    td.append(aux + logC(n, j + 1) + fact[j + 1] + pw(n - j - 1, lq))
            if td:
                psi1d[m, n] = logsumexp(td)
    ip0 = logsumexp(psi0 + pm)
    ip1 = logsumexp(psi1 + pm)
    ip1d = np.array([logsumexp(psi1d[m] + pm) for m in range(M)])
    card = pm + psi0 - ip0
    D = ip1d - ip0 + math.log(cr) - math.log(cd)
    return D, ip1 - ip0, ip0, card, pm


def scenario(C, M, N1, seed, pd=0.95, q_zero=False):
    rng = np.random.default_rng(seed)
    cfg = S.scene_config(1, C, M, filter_type=1, max_cardinality=N1 - 1, pd=pd)
    w = rng.uniform(0.1, 1.0, C).astype(np.float32)
    pdv = np.full(C, 1.0 if q_zero else pd, np.float32)
    # likelihood masses: some measurements are strong detections, some pure clutter
    Sm = np.where(rng.uniform(size=M) < 0.6, rng.uniform(0.5, 60.0, M), rng.uniform(0, 1e-6, M)).astype(np.float32)
    lam = max(float(w.sum()), 0.5)
    n = np.arange(N1)
    prior = (n * math.log(lam) - lam - gammaln(n + 1.0)).astype(np.float32)
    return cfg, w, pdv, Sm, prior


@pytest.mark.parametrize("C,M,N1,seed", [(5, 3, 16, 1), (12, 8, 40, 2), (40, 20, 64, 3), (64, 50, 256, 4), (3, 10, 32, 5),
                                         (0, 4, 16, 6), (30, 1, 48, 7)])
def test_cphd_factors_vs_literal_float64(C, M, N1, seed):
    cfg, w, pdv, Sm, prior = scenario(C, M, N1, seed)
    D, ND, inc, card = O.cphd_factors(cfg, w, pdv, Sm, prior)
    D2, ND2, inc2, card2, _ = literal_cphd(cfg, w, pdv, Sm, prior)
    # fp32 log domain: n*log<1,w> etc. are O(1e3), so the log-domain results carry ~1e-3 absolute error at most
    np.testing.assert_allclose(D, D2, rtol=0, atol=3e-3)
    assert abs(ND - ND2) < 3e-3
    assert abs(inc - inc2) < 3e-3 + 1e-5 * abs(inc2)
    keep = card2 > -60
    np.testing.assert_allclose(card[keep], card2[keep], rtol=0, atol=5e-3)
    assert abs(np.sum(np.exp(card.astype(np.float64))) - 1.0) < 2e-3      # a probability distribution


@pytest.mark.parametrize("kind", ["all_equal", "half_equal", "wide_spread", "one_dominant", "two_clusters"])
def test_cphd_leave_one_out_deflation_adversarial_roots(kind):
    """The oracle obtains the leave-one-out elementary symmetric functions by composite deflation of the full ones
    (forward below the crossover, backward above it) instead of recomputing them per measurement.  Root sets that break a
    one-directional deflation -- repeated roots (clutter-only measurements share one lambda), widely spread magnitudes,
    one dominant root -- must still match the literal formulas, which rebuild every polynomial with numpy.poly."""
    C, M, N1 = 40, 48, 128
    cfg, w, pdv, Sm, prior = scenario(C, M, N1, 21)
    rng = np.random.default_rng(5)
    if kind == "all_equal":
        Sm = np.zeros(M, np.float32)                                   # every lambda = w_b * area
    elif kind == "half_equal":
        Sm = np.concatenate([np.zeros(M // 2), rng.uniform(0.5, 60.0, M - M // 2)]).astype(np.float32)
    elif kind == "wide_spread":
        Sm = (10.0 ** rng.uniform(-6, 3, M)).astype(np.float32)
    elif kind == "one_dominant":
        Sm = np.concatenate([np.full(M - 1, 1e-5), [500.0]]).astype(np.float32)
    else:
        Sm = np.concatenate([np.full(M // 2, 3.0), np.full(M - M // 2, 0.02)]).astype(np.float32)
    D, ND, inc, card = O.cphd_factors(cfg, w, pdv, Sm, prior)
    D2, ND2, inc2, card2, _ = literal_cphd(cfg, w, pdv, Sm, prior)
    np.testing.assert_allclose(D, D2, rtol=0, atol=3e-3)
    assert abs(ND - ND2) < 3e-3 and abs(inc - inc2) < 3e-3 + 1e-5 * abs(inc2)


def test_cphd_large_measurement_set_is_finite():
    """256 measurements: the linear fp32 recursion of the reference overflows; scaled double does not"""
    cfg, w, pdv, Sm, prior = scenario(100, 256, 257, 9)
    D, ND, inc, card = O.cphd_factors(cfg, w, pdv, Sm, prior)
    assert np.isfinite(D).all() and np.isfinite(ND) and np.isfinite(inc)
    assert abs(np.sum(np.exp(card.astype(np.float64))) - 1.0) < 5e-3


@pytest.mark.parametrize("C,M,N1,wscale,pd", [(2, 5, 345, 0.005, 0.95), (200, 60, 345, 1.0, 0.95), (30, 12, 345, 1.0, 1.0),
                                             (0, 3, 345, 1.0, 0.95), (50, 256, 300, 1.0, 0.5), (120, 40, 1024, 1.0, 0.95)])
def test_cphd_linear_domain_tables_stay_in_range(C, M, N1, wscale, pd):
    """Psi0 / A1 are evaluated as double-precision convolutions scaled by s = n_c/<1,w> (n_c = 128 up to 345 bins):
    nearly empty and heavy maps (<1,w> from 0.01 to ~110), q_D = 0 and 256 measurements must all give finite factors, a
    normalised posterior cardinality and agreement with the literal float64 formulas where those are representable.
    (The random likelihood masses of `scenario` pull the posterior towards ~40 objects whatever the map weight is; for
    <1,w> above ~150 the Poisson prior is below e^-87 there and the reference's fp32 `exp(birth + prior)` prediction
    (src/phdfilter.cu:880-887), which the oracle restates, flushes it to zero -- a property of the reference's arithmetic,
    pinned in tests/test_ref_pin.py, so such inconsistent inputs are not compared with the float64 formulas here.)"""
    rng = np.random.default_rng(C + M)
    cfg = S.scene_config(1, max(C, 1), M, filter_type=1, max_cardinality=N1 - 1, pd=pd)
    w = (rng.uniform(0.1, 1.0, C) * wscale).astype(np.float32)
    pdv = np.full(C, pd, np.float32)
    Sm = np.where(rng.uniform(size=M) < 0.5, rng.uniform(0.5, 60.0, M), rng.uniform(0, 1e-6, M)).astype(np.float32)
    lam = max(float(w.sum()), 0.3)
    n = np.arange(N1)
    prior = (n * math.log(lam) - lam - gammaln(n + 1.0)).astype(np.float32)
    D, ND, inc, card = O.cphd_factors(cfg, w, pdv, Sm, prior)
    assert np.isfinite(D).all() and np.isfinite(ND) and np.isfinite(inc)
    assert abs(np.sum(np.exp(card.astype(np.float64))) - 1.0) < 5e-3
    D2, ND2, inc2, card2, _ = literal_cphd(cfg, w, pdv, Sm, prior)
    ok = np.isfinite(D2)
    np.testing.assert_allclose(D[ok], D2[ok], rtol=0, atol=5e-3)
    if np.isfinite(ND2):
        assert abs(ND - ND2) < 5e-3
    keep = np.isfinite(card2) & (card2 > -60)
    np.testing.assert_allclose(card[keep], card2[keep], rtol=0, atol=1e-2)


@pytest.mark.parametrize("C,N1,pd", [(700, 1024, 0.08), (400, 690, 0.15), (300, 346, 0.2)])
def test_cphd_large_cardinalities_use_a_larger_table_scale(C, N1, pd):
    """Hundreds of objects (a low detection probability makes that consistent with 60 measurements): the scale of the
    linear-domain tables grows with the bin count (phd_cphd_log_nc: 128 up to 345 bins, 256 up to 690, 512 beyond), so that
    n!/n_c^n stays below 1 over the whole cardinality range.  With the fixed scale of 128 the posterior of the 700-object
    case came out as LOG0."""
    M = 60
    rng = np.random.default_rng(C + M)
    cfg = S.scene_config(1, C, M, filter_type=1, max_cardinality=N1 - 1, pd=pd)
    w = rng.uniform(0.5, 1.5, C).astype(np.float32)
    pdv = np.full(C, pd, np.float32)
    Sm = rng.uniform(0.5, 60.0, M).astype(np.float32)
    n = np.arange(N1)
    lam = float(C)
    prior = (n * math.log(lam) - lam - gammaln(n + 1.0)).astype(np.float32)
    D, ND, inc, card = O.cphd_factors(cfg, w, pdv, Sm, prior)
    D2, ND2, inc2, card2, _ = literal_cphd(cfg, w, pdv, Sm, prior)
    np.testing.assert_allclose(D, D2, rtol=0, atol=3e-3)
    assert abs(ND - ND2) < 3e-3 and abs(inc - inc2) < 3e-3 + 1e-5 * abs(inc2)
    keep = card2 > -40
    assert abs(int(card.argmax()) - int(card2.argmax())) <= 1 and abs(int(card2.argmax()) - C) < 0.05 * C
    np.testing.assert_allclose(card[keep], card2[keep], rtol=0, atol=1e-2)


def test_cphd_poisson_cardinality_reduces_to_phd():
    """With a Poisson predicted cardinality of mean <1,w> the CPHD update is the PHD update:
    detection factor exp(D_m)*area... = 1/(kappa + S_m + w_b), non-detection factor 1 (Vo et al. 2007, sec. IV)."""
    C, M, N1 = 20, 6, 200
    cfg, w, pdv, Sm, _ = scenario(C, M, N1, 11)
    W = float(w.sum()) + M * cfg.birth_weight
    n = np.arange(N1)
    pm = n * math.log(W) - W - gammaln(n + 1.0)
    D2, ND2, inc2, card2, _ = literal_cphd(cfg, w, pdv, Sm, pm, do_predict=False)
    kappa = cfg.clutter_density
    np.testing.assert_allclose(np.exp(D2), 1.0 / (kappa + Sm + cfg.birth_weight), rtol=1e-6)
    assert abs(ND2) < 1e-6
    # the oracle convolves the prior with Binomial(M, w_b) births first; feed it a prior whose prediction is ~Poisson(W):
    # Poisson(W - M w_b) * Binomial(M, w_b) ~ Poisson(W) up to O(M w_b^2)
    W0 = W - M * cfg.birth_weight
    prior = (n * math.log(W0) - W0 - gammaln(n + 1.0)).astype(np.float32)
    D, ND, inc, card = O.cphd_factors(cfg, w, pdv, Sm, prior)
    np.testing.assert_allclose(np.exp(D.astype(np.float64)), 1.0 / (kappa + Sm + cfg.birth_weight), rtol=2e-3)
    assert abs(ND) < 2e-3


def test_cphd_cardinality_predict_is_a_convolution():
    cfg, w, pdv, Sm, prior = scenario(8, 5, 32, 13)
    _, _, _, _, pm = literal_cphd(cfg, w, pdv, Sm, prior)
    pb = np.array([math.comb(5, k) * cfg.birth_weight ** k * (1 - cfg.birth_weight) ** (5 - k) for k in range(6)])
    conv = np.convolve(np.exp(prior.astype(np.float64)), pb)[:32]
    np.testing.assert_allclose(np.exp(pm), conv, rtol=1e-9)


def test_cphd_filter_step_runs_and_keeps_cardinality_normalised():
    Pn, C, M = 6, 24, 9
    cfg = S.scene_config(Pn, C, M, max_components=256, filter_type=1, max_cardinality=63)
    sc = S.make_scene(Pn, C, M, seed=21, n_near=2, n_far=2)
    o = O.Oracle(cfg)
    S.load_scene(o, sc)
    n = np.arange(64)
    lam = float(sc["maps"]["weight"][:C + 4].sum())
    o.cardinalities = np.tile((n * math.log(lam) - lam - gammaln(n + 1.0)).astype(np.float32), (Pn, 1))
    o.phdUpdateSynth(sc["Z"])
    card = o.cardinalities
    assert np.allclose(np.exp(card.astype(np.float64)).sum(1), 1.0, atol=2e-3)
    sizes, maps = o.get_maps()
    assert (sizes > 0).all() and np.isfinite(maps["weight"]).all() and (maps["weight"] >= 0).all()
    assert np.isfinite(o.log_weights).all()
    # the expected number of targets of the updated intensity tracks the cardinality mean (CPHD consistency)
    mean_card = (np.exp(card.astype(np.float64)) * n).sum(1)
    off = 0
    for p in range(Pn):
        mass = float(maps["weight"][off:off + sizes[p]].sum())
        off += sizes[p]
        # in-range mass only is governed by the cardinality; the far/near components bypass the update
        assert abs(mass - mean_card[p]) < 0.35 * max(mean_card[p], 1.0) + 4 * 0.6
