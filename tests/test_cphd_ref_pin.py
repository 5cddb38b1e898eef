"""Pins the CPHD row of the CPU oracle to the REFERENCE's own CPHD kernels, run on the CPU.

The reference's CPHD update is dead code at HEAD -- commented out line by line (src/phdfilter.cu:543-607,701-748,
1430-1822) -- and live, in an older form, in src/phdfilter.cu.bak:369-415,1058-1478.  oracle/ref_build.sh makes both
executable (HEAD: the leading `//` of every line removed; .bak: verbatim) and oracle/ref_harness.cpp launches them through
the CUDA emulator with the .bak wrapper's launch sequence (.bak:2503-2544).  Their outputs on the cases of
ref_cases.CPHD_CASES are committed as tests/golden/ref_cphd_golden.npz (generator: tests/golden/make_ref_cphd_golden.py).

What is compared: detection and non-detection weights, means and covariances of every updated component, the particle
log-weight increment log<Psi0,p-> and the posterior cardinality distribution.  Tolerances (ref_cases.py): weights 1e-4
relative, log-domain scalars 1e-4 absolute -- what the reference's fp32 log-domain sums over up to 256 x 31 terms resolve.
What the reference's own defects force (ref_cases.py, CPHD_CASES): HEAD's fp32 linear-domain ESF recursion overflows
beyond about a dozen detected landmarks, so the M = 30 cases run the .bak's log-domain kernels, whose detection side
(leave-one-out ESFs through fabs() of a difference; wrong maximum in <Psi1d,p>) is replaced by the identity
Psi1d[w,Z](m) = Psi1[w, Z \\ {z_m}] evaluated with the same kernels -- an identity this file first checks on HEAD's kernels.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref as R
from conftest import GOLDEN
import ref_cases as RC

GOLD = np.load(os.path.join(GOLDEN, "ref_cphd_golden.npz"))
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libphd_ref.so not built (needs /root/reference)")


def gold(name):
    pre = name + "/"
    return {k[len(pre):]: GOLD[k] for k in GOLD.files if k.startswith(pre)}


def oracle_for(name):
    cfg, sc, _ = RC.build_cphd_case(name)
    g = gold(name)
    o = O.Oracle(cfg)
    o.poses = g["in_poses"]
    o.log_weights = sc["log_weights"]
    o.set_maps(g["in_sizes"], g["in_maps"])
    o.cardinalities = g["cn_predict"]
    return cfg, o, g


@pytest.mark.parametrize("name", list(RC.CPHD_CASES))
def test_golden_inputs_reproducible(name):
    _, sc, _ = RC.build_cphd_case(name)
    g = gold(name)
    assert sc["maps"].tobytes() == g["in_maps"].tobytes() and sc["Z"].tobytes() == g["Z"].tobytes()
    assert sc["cn_predict"].tobytes() == g["cn_predict"].tobytes()


@pytest.mark.parametrize("name", list(RC.CPHD_CASES))
def test_oracle_cphd_update_vs_reference_kernels(name):
    cfg, o, g = oracle_for(name)
    terms, n_in, dlw = o.update_terms(g["Z"])
    lw0 = o.log_weights.astype(np.float64)
    o.phdUpdateSynth(g["Z"])
    RC.assert_cphd_matches_reference(g, terms, n_in, dlw, o.cardinalities, name)
    # the particle weights carry the increment (.bak:2666, 2695): w_i += log<Psi0,p->, then normalisation
    w = lw0 + g["ip0"].astype(np.float64)
    w -= np.log(np.sum(np.exp(w - w.max()))) + w.max()
    RC.close(o.log_weights, w, name + " particle log-weights", atol=2e-5)


@needs_ref
def test_golden_is_current():
    for name in RC.CPHD_CASES:
        r, g = RC.reference_cphd(name), gold(name)
        for k in r:
            assert np.asarray(r[k]).tobytes() == g[k].tobytes(), (name, k)


@needs_ref
@pytest.mark.parametrize("name", ["cphd_m8", "cphd_m8_lowpd"])
def test_reference_identities_on_head_kernels(name):
    """On HEAD's (finite, correct) kernels: (1) <Psi1d_m, p> equals <Psi1, p> of the measurement set without z_m -- the
    identity the M = 30 cases use; (2) the .bak's kernels agree with HEAD's on everything except the detection side."""
    cfg, sc, _ = RC.build_cphd_case(name)
    R.set_config(cfg)
    n_in = sc["sizes"]
    h = R.cphd_update(sc["poses"], sc["maps"], n_in, sc["Z"], sc["cn_predict"], "head")
    b = R.cphd_update(sc["poses"], sc["maps"], n_in, sc["Z"], sc["cn_predict"], "bak")
    for m in range(len(sc["Z"])):
        hm = R.cphd_update(sc["poses"], sc["maps"], n_in, np.delete(sc["Z"], m, axis=0), sc["cn_predict"], "head")
        RC.close(hm["ip1"], h["ip1d"][:, m], "Psi1d(m) = Psi1(Z without m)", rtol=1e-6, atol=5e-5)
    for k in ("ip0", "ip1"):
        RC.close(b[k], h[k], ".bak vs HEAD " + k, rtol=1e-6, atol=5e-5)
    ok = h["esf"] > -60.0            # HEAD's linear-domain coefficients underflow to log(0) where the .bak's do not
    RC.close(b["esf"][ok], h["esf"][ok], ".bak vs HEAD esf", rtol=1e-6, atol=5e-5)
    live = h["cn_update"] > -80
    RC.close(b["cn_update"][live], h["cn_update"][live], ".bak vs HEAD cardinality", rtol=1e-6, atol=5e-5)
    RC.close(b["nondetect"]["weight"], h["nondetect"]["weight"], ".bak vs HEAD non-detection weights", rtol=2e-5)
    assert b["w_partial"].tobytes() == h["w_partial"].tobytes()
    # and the two documented defects of the .bak's detection side are real: its detection weights are far off
    assert np.max(np.abs(np.log(b["detect"]["weight"] + 1e-30) - np.log(h["detect"]["weight"] + 1e-30))) > 1.0


@needs_ref
def test_head_linear_esf_overflows_at_m30():
    """why the M = 30 cases cannot use HEAD's kernels: their fp32 linear-domain ESF recursion is not finite there"""
    cfg, sc, _ = RC.build_cphd_case("cphd_m30")
    R.set_config(cfg)
    h = R.cphd_update(sc["poses"], sc["maps"], sc["sizes"], sc["Z"], sc["cn_predict"], "head")
    assert not np.isfinite(h["ip0"]).all()


@needs_ref
@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_random_cphd_scenes_vs_reference_kernels(seed):
    """beyond the fixtures: random small scenes, HEAD's kernels, run live"""
    rng = np.random.default_rng(seed)
    Pn, C, M = int(rng.integers(1, 5)), int(rng.integers(1, 30)), int(rng.integers(1, 10))
    RC.CPHD_CASES["_tmp"] = (Pn, C, M, 100 + seed, "head", dict(pd=float(rng.uniform(0.5, 0.99)),
                                                                 clutter_rate=float(rng.uniform(1.0, 30.0))))
    try:
        r = RC.reference_cphd("_tmp")
        cfg, sc, _ = RC.build_cphd_case("_tmp")
    finally:
        del RC.CPHD_CASES["_tmp"]
    if not np.isfinite(r["ip0"]).all():
        pytest.skip("HEAD's fp32 linear-domain ESF overflowed on this scene")
    o = O.Oracle(cfg)
    o.poses, o.log_weights = sc["poses"], sc["log_weights"]
    o.set_maps(sc["sizes"], sc["maps"])
    o.cardinalities = sc["cn_predict"]
    terms, n_in, dlw = o.update_terms(sc["Z"])
    o.phdUpdateSynth(sc["Z"])
    RC.assert_cphd_matches_reference(r, terms, n_in, dlw, o.cardinalities, "random CPHD scene %d" % seed)
