"""Cases shared by tests/golden/make_ref_golden.py (which runs the REFERENCE's own kernels, compiled for the CPU
by oracle/ref_build.sh, and stores their outputs in tests/golden/ref_golden.npz), tests/test_ref_pin.py (oracle vs
reference kernels / golden) and tests/test_golden_gpu.py (CUDA path vs golden).

Each case is a full static-map phdUpdateSynth (src/phdfilter.cu:3336-3761) on a small seeded scene.
"""
import numpy as np

import phdslam_b200 as P
from phdslam_b200 import scene as S

# name -> (P, C, M, n_near, n_far, seed, cfg overrides)
UPDATE_CASES = {
    "base": (6, 40, 12, 4, 5, 3, dict(min_range=1.0)),
    "many_terms": (3, 72, 20, 3, 2, 7, dict()),                       # > 256 merge candidates: several reduction chunks
    "vo_weighting": (5, 24, 9, 2, 2, 11, dict(particle_weighting=1)),
    "hellinger": (4, 20, 8, 2, 0, 13, dict(distance_metric=1, min_separation=0.5)),
    "empty_maps": (5, 0, 7, 0, 0, 17, dict()),                        # first step: births only
    "far_only": (4, 0, 6, 3, 4, 19, dict()),                          # nothing in range, class-2/0 bypass
    "narrow_fov": (5, 30, 10, 0, 0, 23, dict(max_bearing=1.2, min_range=2.0, max_range=12.0)),
    "single": (1, 1, 1, 0, 0, 29, dict()),
    "low_pd_tight": (4, 36, 16, 2, 2, 31, dict(pd=0.6, min_separation=4.0, min_feature_weight=1e-4, birth_weight=0.05)),
    # BASELINE shapes per particle (the reference's kernels through the emulator take a few seconds per particle here)
    "headline_shape": (2, 256, 64, 0, 0, 41, dict()),                 # configs[2]: 256 components x 64 measurements
    "configs4_shape": (2, 128, 100, 3, 2, 43, dict()),                # configs[4]: 128 components x 100 measurements
}
LABELED_CASE = ("labeled", (4, 16, 8, 1, 1, 37, dict(labeled_measurements=1)))

RTOL = 1e-4   # BASELINE.json north_star: floats within 1e-4 relative; counts / indices bit-exact
# Coordinates (map means, poses) live on a ~15 m scene: a coordinate that happens to be near 0 carries the absolute
# rounding of the O(10) terms it is the difference of (1 ulp(8.0) = 9.5e-7), so positions get an absolute floor of
# 1e-4 * 0.2 m = 2e-5 m next to the relative bound.  Covariances are compared norm-wise: every entry within
# 1e-4 of the largest entry of that 2x2 matrix (an off-diagonal near 0 is a difference of O(trace) terms).
# Component weights are not in north_star's list (poses, map means, covariances, log-weights).  A detection weight
# is exp(-d^2/2 + ...) with d^2 = nu' S nu, S ~ 1/std_bearing^2 = 1.3e4: the three products of the quadratic form are
# O(1e3) and cancel to O(10), so fp32 rounding alone (fused or not, summation shape) moves d^2 by ~1e-4 absolute and the
# weight by ~1e-4 relative.  They are held to 5e-4.
ATOL_POS = 2e-5
ATOL_COV = 2e-7
RTOL_COMPONENT_WEIGHT = 5e-4


def build_case(name):
    if name == LABELED_CASE[0]:
        Pn, C, M, near, far, seed, over = LABELED_CASE[1]
    else:
        Pn, C, M, near, far, seed, over = UPDATE_CASES[name]
    cfg = S.scene_config(Pn, C, M, max_components=512, **over)
    sc = S.make_scene(Pn, C, M, seed=seed, n_near=near, n_far=far, max_range=float(over.get("max_range", 15.0)))
    rng = np.random.Generator(np.random.Philox(1000 + seed))
    sc["log_weights"] = np.log(rng.dirichlet(np.ones(Pn) * 2.0)).astype(np.float32) if Pn > 1 else np.zeros(1, np.float32)
    if name == LABELED_CASE[0]:
        lab = (rng.uniform(size=M) < 0.3).astype(np.float32)
        sc["Z"] = np.concatenate([sc["Z"], lab[:, None]], 1).astype(np.float32)
    return cfg, sc


def close(a, b, what, rtol=RTOL, atol=1e-7):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), "%s: max rel err %.3g, max abs err %.3g" % (
        what, np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)), np.max(np.abs(a - b)))


def close_cov(a, b, what, rtol=RTOL):
    """norm-wise: |a - b| <= rtol * max|b| per 2x2 matrix"""
    a = np.asarray(a, dtype=np.float64).reshape(-1, 4)
    b = np.asarray(b, dtype=np.float64).reshape(-1, 4)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = np.abs(b).max(axis=1, keepdims=True)
    err = np.abs(a - b) - (1e-12 + rtol * scale)
    assert (err <= 0).all(), "%s: max norm-wise rel err %.3g" % (what, np.max(np.abs(a - b) / np.maximum(scale, 1e-30)))


def assert_gaussians_close(g, ref, what):
    assert len(g) == len(ref), what
    close(g["weight"], ref["weight"], what + ".weight", rtol=RTOL_COMPONENT_WEIGHT, atol=1e-12)
    close(g["mean"], ref["mean"], what + ".mean", atol=ATOL_POS)
    close_cov(g["cov"], ref["cov"], what + ".cov")


def assert_maps_close(sizes, maps, ref_sizes, ref_maps, what, allow_near_tie_reorder=False):
    """Component counts bit-exact, components in the same order and within tolerance.
    allow_near_tie_reorder: output order = descending seed weight; two seeds whose weights agree to ~1 ulp (typical:
    clutter births, w_b/(kappa + w_b + 1e-12)) are ordered by rounding noise, so for long loops a particle whose
    in-order comparison fails is re-compared after matching components by nearest mean (must be a bijection)."""
    sizes, ref_sizes = np.asarray(sizes), np.asarray(ref_sizes)
    assert (sizes == ref_sizes).all(), what + ": component counts differ (must be bit-exact)"
    if not allow_near_tie_reorder:
        return assert_gaussians_close(maps, ref_maps, what)
    off = 0
    for p, n in enumerate(sizes):
        a, b = maps[off:off + n], ref_maps[off:off + n]
        off += n
        try:
            assert_gaussians_close(a, b, "%s particle %d" % (what, p))
        except AssertionError:
            d = np.linalg.norm(a["mean"][:, None, :].astype(np.float64) - b["mean"][None, :, :], axis=2)
            match = d.argmin(axis=1)
            assert len(set(match.tolist())) == n, "%s particle %d: not a permutation of the reference's components" % (what, p)
            moved = np.nonzero(match != np.arange(n))[0]
            assert len(moved) <= max(4, n // 4), "%s particle %d: %d components out of order" % (what, p, len(moved))
            assert_gaussians_close(a, b[match], "%s particle %d (matched)" % (what, p))


def cv_noise(cfg, draws):
    """noiseVector of phdPredict's CV branch (src/phdfilter.cu:1113-1117): ax = 3*config.ax*randn(), stored as float"""
    d = np.asarray(draws, np.float64).reshape(-1, 3)
    s = np.array([np.float32(3.0) * np.float32(cfg.ax), np.float32(3.0) * np.float32(cfg.ay),
                  np.float32(3.0) * np.float32(cfg.ayaw)], np.float32).astype(np.float64)
    return (d * s[None, :]).astype(np.float32)


def ackerman_noise(cfg, draws):
    """(:1148-1152): n_alpha = config.stdAlpha*randn(); n_encoder = config.stdEncoder*randn(), stored as float"""
    d = np.asarray(draws, np.float64).reshape(-1, 2)
    s = np.array([cfg.std_alpha, cfg.std_encoder], np.float32).astype(np.float64)
    return (d * s[None, :]).astype(np.float32)


def predict_config(motion_type, n=64):
    return S.scene_config(n, 1, 1, motion_type=motion_type, acc_x=0.5, acc_y=0.2, acc_yaw=0.1, subdivide_predict=2)


# ---- CPHD: the reference's multi-object update (HEAD's commented-out kernels made live / the .bak's live kernels, both
# through the emulator: oracle/ref_build.sh).  The reference creates births BEFORE the update (addBirths, .bak:794-900)
# and this implementation, like HEAD's PHD path, inside it; with birth_weight = 0 the two coincide, and the predicted
# cardinality handed to computePsiKernel is the prior.  max_cardinality = 255: the reference's reductions are 256 wide.
# name -> (P, C, M, seed, variant, cfg overrides)
#   variant "head": every output of HEAD's kernels is usable while its fp32 LINEAR-domain ESF recursion stays finite
#                   (roots lambda_m ~ 1e3 for a detected landmark: about a dozen of them overflow 3e38);
#   variant "bak" : log-domain kernels, finite at any M, but (1) the leave-one-out ESF recursion takes fabs() of a
#                   DIFFERENCE of exponentials at every step (.bak:1264-1266: not an elementary symmetric function) and
#                   (2) <Psi1d_m,p> is normalised with the wrong maximum (.bak:1417 `exp(val-max_val0)`).  Both defects
#                   sit on the detection side only; the detection factor of measurement m is therefore taken from the
#                   same kernels run on Z \ {z_m}: Psi1d[w,Z](m) = Psi1[w, Z \ {z_m}] (Vo, Vo & Cantoni 2007, eq. 21-22),
#                   which only exercises the full-ESF code.  tests/test_ref_pin.py checks that identity on HEAD's kernels.
CPHD_CASES = {
    "cphd_m1": (3, 10, 1, 1, "head", dict()),
    "cphd_m8": (3, 12, 8, 2, "head", dict()),
    "cphd_m8_lowpd": (4, 5, 8, 5, "head", dict(pd=0.6, clutter_rate=5.0)),
    "cphd_m30": (2, 20, 30, 3, "bak", dict()),
    "cphd_m30_sparse": (3, 6, 30, 7, "bak", dict(pd=0.8)),
}
CPHD_RTOL_WEIGHT = 1e-4     # component weights against the reference kernels' (fp32 log-domain sums of up to 256 x 31 terms)
CPHD_ATOL_LOG = 1e-4        # log-domain scalars (log<Psi0,p>, log cardinality rows): |a - b| <= 1e-4 + 1e-6 |b|


def build_cphd_case(name):
    Pn, C, M, seed, variant, over = CPHD_CASES[name]
    cfg = S.scene_config(Pn, C, M, max_components=512, filter_type=1, max_cardinality=255, birth_weight=0.0, **over)
    sc = S.make_scene(Pn, C, M, seed=seed, n_near=0, n_far=0)
    # predicted cardinality: Poisson(sum of the map weights), as the .bak wrapper substitutes (.bak:2473-2499)
    lf = np.concatenate([[0.0], np.cumsum(np.log(np.arange(1, 256, dtype=np.float64)))])
    cn = np.zeros((Pn, 256), np.float32)
    off = 0
    for p in range(Pn):
        ws = float(sc["maps"]["weight"][off:off + sc["sizes"][p]].astype(np.float64).sum())
        off += sc["sizes"][p]
        cn[p] = (np.arange(256) * np.log(ws) - ws - lf).astype(np.float32)
    sc["cn_predict"] = cn
    return cfg, sc, variant


def reference_cphd(name):
    """Runs the reference's CPHD kernels (needs oracle/_ref).  Returns the dict stored in tests/golden/ref_cphd_golden.npz."""
    from oracle import ref as R
    cfg, sc, variant = build_cphd_case(name)
    R.set_config(cfg)
    cls, n_in, _ = R.in_range(sc["maps"], sc["sizes"], sc["poses"])
    assert (cls == 1).all()
    r = R.cphd_update(sc["poses"], sc["maps"], n_in, sc["Z"], sc["cn_predict"], variant)
    M = len(sc["Z"])
    ip1d = r["ip1d"].copy()
    detect = r["detect"].copy()
    if variant == "bak":
        for m in range(M):                                   # Psi1d(m) = Psi1 on Z without z_m: full-ESF code only
            rm = R.cphd_update(sc["poses"], sc["maps"], n_in, np.delete(sc["Z"], m, axis=0), sc["cn_predict"], variant)
            ip1d[:, m] = rm["ip1"]
        lcr, lcd = np.log(np.float32(cfg.clutter_rate)), np.log(np.float32(cfg.clutter_density))
        part = np.repeat(np.arange(len(n_in)), n_in)
        # cphdUpdateKernel's detection weight (.bak:1449-1452) with the repaired factor, in the kernel's fp32 order
        t = r["w_partial"] + ip1d[part, :] - r["ip0"][part, None] + lcr - lcd
        detect["weight"] = np.exp(t.astype(np.float32))
    return dict(in_poses=sc["poses"], in_sizes=sc["sizes"], in_maps=sc["maps"], Z=sc["Z"], cn_predict=sc["cn_predict"],
                n_in=n_in, detect=detect, nondetect=r["nondetect"], cn_update=r["cn_update"], ip0=r["ip0"], ip1=r["ip1"],
                ip1d=ip1d, esf=r["esf"])


def split_cphd_terms(terms, n_in, M):
    """this implementation's dense layout [non-detect C | detect m-major M*C | birth M] per particle ->
    (nondetect [sum C], detect [sum C][M] feature-major as the reference stores them, births [P][M])"""
    nd, det, births = [], [], []
    off = 0
    for c in n_in:
        t = terms[off:off + c * (M + 1) + M]
        off += c * (M + 1) + M
        nd.append(t[:c])
        det.append(t[c:c + M * c].reshape(M, c).T)
        births.append(t[c + M * c:])
    return np.concatenate(nd), np.concatenate(det), np.stack(births)


def assert_cphd_matches_reference(ref, terms, n_in, dlogw, card, what):
    """terms / n_in / dlogw: update_terms() of the implementation under test on the case's inputs; card: its cardinality
    rows after the update."""
    M = len(ref["Z"])
    assert (np.asarray(n_in) == ref["n_in"]).all(), what
    nd, det, births = split_cphd_terms(terms, n_in, M)
    assert (births["weight"] == 0).all(), what + ": birth_weight = 0"
    close(nd["weight"], ref["nondetect"]["weight"], what + " non-detection weights", rtol=CPHD_RTOL_WEIGHT, atol=1e-12)
    close(det["weight"], ref["detect"]["weight"], what + " detection weights", rtol=CPHD_RTOL_WEIGHT, atol=1e-30)
    close(det["mean"], ref["detect"]["mean"], what + " detection means", atol=ATOL_POS)
    close_cov(det["cov"], ref["detect"]["cov"], what + " detection covariances")
    close(dlogw, ref["ip0"], what + " particle increment log<Psi0,p>", rtol=1e-6, atol=CPHD_ATOL_LOG)
    live = ref["cn_update"] > -80.0
    close(np.asarray(card)[live], ref["cn_update"][live], what + " posterior cardinality", rtol=1e-6, atol=CPHD_ATOL_LOG)
