/*
 * ref_harness.cpp -- C entry points around the REFERENCE's own kernels, compiled for the CPU.
 *
 * TEST INFRASTRUCTURE (part of the oracle; see oracle/ref_build.sh for how and why it is built).  The kernels
 * and helpers themselves are #included from oracle/_ref/gen/*.inc, which ref_build.sh cuts verbatim out of
 * /root/reference/src/{phdfilter.cu,main.cpp,device_math.cuh,gm_reduce.cpp}; this file only
 *   - supplies the globals those kernels expect (`dev_config`, `config`, constant `Z[256]`, randu01()),
 *   - marshals flat C arrays into the kernels' argument lists, using the same buffer layouts the reference's
 *     host wrapper builds (each site cites the wrapper lines it mirrors), and launches them through the
 *     fiber emulator with the reference's launch shape (256 threads per block),
 *   - restates the thin host glue BETWEEN kernels of phdUpdateSynth (stable prune, recombination order,
 *     particle-weight normalisation) so a whole update can be compared.
 * It is used by tests/test_ref_pin.py to pin oracle/phd_oracle.cpp, and to generate tests/golden/ref_*.npz.
 */
#define PHD_CUDA_EMUL_IMPL
#include "cuda_emul.h"

class MotionModel; /* src/slamtypes.h:335 names it without declaring it */

#include "device_math_syncwarp.cuh" /* = src/device_math.cuh (+ __syncwarp), pulls in src/slamtypes.h */

#include "phdslam.h"

/* globals of src/phdfilter.cu:118-119 and src/main.cpp */
RangeBearingMeasurement Z[256];
SlamConfig dev_config;
SlamConfig config;

#define DEBUG_MSG(x)
#define DEBUG_VAL(x)
#define checkCudaErrors(x) (x)

/* src/rng.h:13-25: extern "C" double randu01() -- here: injected draws */
static const double* g_uniforms = nullptr;
static size_t g_uniform_pos = 0;
extern "C" double randu01() { return g_uniforms[g_uniform_pos++]; }

/* src/gm_reduce.cpp:8-134, the reference's own EAP map reduction, over the Eigen stand-in of oracle/ref_shim/eigen3 */
#include <deque>
#include <algorithm>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <unistd.h>
#include "eigen3/Eigen/Core"
#include "eigen3/Eigen/Cholesky"
#include "eigen3/Eigen/LU"
#include "ref_gm_reduce.inc"

/* src/phdfilter.cu:114 `extern __shared__ REAL shmem[]`: the dynamic shared memory of cardinalityPredictKernel */
static REAL shmem[4096];

#include "ref_kernels.inc"
#include "ref_host.inc"

typedef phdslam_gaussian2d_t G2;
typedef phdslam_pose_t Pose;
static_assert(sizeof(G2) == sizeof(Gaussian2D), "layout");
static_assert(sizeof(Pose) == sizeof(ConstantVelocityState), "layout");

static const int kThreads = 256;

extern "C" void ref_set_config(const phdslam_config_t* c) {
  SlamConfig s;
  memset(&s, 0, sizeof(s));
  s.x0 = c->x0; s.y0 = c->y0; s.yaw0 = c->yaw0; s.vx0 = c->vx0; s.vy0 = c->vy0; s.vyaw0 = c->vyaw0;
  s.ax = c->ax; s.ay = c->ay; s.ayaw = c->ayaw; s.dt = c->dt;
  s.minRange = c->min_range; s.maxRange = c->max_range; s.maxBearing = c->max_bearing;
  s.stdRange = c->std_range; s.stdBearing = c->std_bearing;
  s.clutterRate = c->clutter_rate; s.clutterDensity = c->clutter_density; s.pd = c->pd;
  s.n_particles = c->n_particles; s.nPredictParticles = c->n_predict_particles; s.subdividePredict = c->subdivide_predict;
  s.resampleThresh = c->resample_threshold; s.birthWeight = c->birth_weight; s.birthNoiseFactor = c->birth_noise_factor;
  s.minSeparation = c->min_separation; s.minFeatureWeight = c->min_feature_weight;
  s.particleWeighting = c->particle_weighting; s.distanceMetric = c->distance_metric;
  s.maxCardinality = c->max_cardinality; s.filterType = c->filter_type; s.mapEstimate = c->map_estimate;
  s.featureModel = c->feature_model; s.motionType = c->motion_type; s.labeledMeasurements = c->labeled_measurements != 0;
  s.l = c->l; s.h = c->h; s.a = c->a; s.b = c->b; s.stdAlpha = c->std_alpha; s.stdEncoder = c->std_encoder;
  s.ps = c->ps; s.tau = c->tau; s.beta = c->beta; s.stdAxMap = c->std_ax_features; s.stdAyMap = c->std_ay_features;
  s.covVxBirth = c->cov_vx_birth; s.covVyBirth = c->cov_vy_birth;
  s.nSamples = 256;
  dev_config = s;   /* setDeviceConfig, src/phdfilter.cu:3885-3890 */
  config = s;
}

static void set_measurements(const float* z, int M, int fields) {
  for (int m = 0; m < M; ++m) {
    Z[m].range = z[m * fields];
    Z[m].bearing = z[m * fields + 1];
    Z[m].label = fields > 2 ? (int)z[m * fields + 2] : 0;
  }
}

/* ---- device_math.cuh helpers ---- */
extern "C" float ref_wrap_angle(float a) { return wrapAngle(a); }
extern "C" float ref_safe_log(float x) { return safeLog(x); }
extern "C" float ref_mahalanobis(const G2* a, const G2* b) {
  Gaussian2D x, y;
  memcpy(&x, a, sizeof(x)); memcpy(&y, b, sizeof(y));
  return computeMahalDist(x, y);
}
extern "C" float ref_hellinger(const G2* a, const G2* b) {
  Gaussian2D x, y;
  memcpy(&x, a, sizeof(x)); memcpy(&y, b, sizeof(y));
  return computeHellingerDist(x, y);
}
extern "C" float ref_log_sum_exp(const float* w, int n) { return logSumExp(std::vector<float>(w, w + n)); }
/* sumByReduction through the emulator: the 256-wide shared-memory tree itself */
extern "C" float ref_sum_by_reduction(const float* v256) {
  static float sdata[256];
  emul_launch(1, kThreads, [&] { sumByReduction(sdata, v256[threadIdx.x], threadIdx.x); });
  return sdata[0];
}

/* ---- CPHD cardinality prediction: cardinalityPredictKernel (src/phdfilter.cu:867-888), the one CPHD kernel that is
 * compiled at HEAD: one block per particle, one thread per cardinality n (its launch is commented out at HEAD together
 * with the rest of the CPHD update; the .bak pipeline launches it with <<<nParticles, maxCardinality+1>>>, .bak:592).
 * prior [n_particles][N+1], births [N+1] log-probabilities; out [n_particles][N+1]. */
extern "C" void ref_cardinality_predict(const float* prior, const float* births, int n_particles, float* out) {
  const int n1 = dev_config.maxCardinality + 1;
  emul_launch(n_particles, n1, [&] { cardinalityPredictKernel((REAL*)prior, (REAL*)births, nullptr, out); });
}

/* ---- predict: phdPredictKernelAckerman / phdPredictKernel, launched as src/phdfilter.cu:1119-1170 does.
 * noise = what the host wrapper stores in noiseVector (float), [n][2] {n_alpha, n_encoder} or [n][3]. */
extern "C" void ref_predict(const Pose* in, int n, const float* control_venc_alpha, const float* noise, Pose* out) {
  int nb = (n + kThreads - 1) / kThreads;
  ConstantVelocityState* pin = (ConstantVelocityState*)in;
  ConstantVelocityState* pout = (ConstantVelocityState*)out;
  if (config.motionType == ACKERMAN_MOTION) {
    AckermanControl u;
    u.v_encoder = control_venc_alpha[0];
    u.alpha = control_venc_alpha[1];
    emul_launch(nb, kThreads, [&] { phdPredictKernelAckerman(pin, u, (AckermanNoise*)noise, pout, n); });
  } else {
    emul_launch(nb, kThreads, [&] { phdPredictKernel(pin, (ConstantVelocityNoise*)noise, pout, n); });
  }
}

/* ---- in-range classification: computeInRangeKernel launched as prepareUpdateInputs does (:2950-2990) ---- */
extern "C" void ref_in_range(const G2* features, const int* map_sizes, int n_particles, const Pose* poses, char* in_range,
                             int* n_in, int* n_nearly) {
  emul_launch(std::min(n_particles, 65535), kThreads, [&] {
    computeInRangeKernel((Gaussian2D*)features, (int*)map_sizes, n_particles, (ConstantVelocityState*)poses, in_range, n_in,
                         n_nearly);
  });
}

/* births: the host loop of phdUpdateSynth, src/phdfilter.cu:3468-3510, verbatim */
static void births_host(SynthSLAM& particles, measurementSet& measurements, int n_particles, int n_measure,
                        vector<Gaussian2D>& births) {
#include "ref_births.inc"
}
/* and its __device__ twin computeBirth (:205-242) */
extern "C" void ref_birth_device(const Pose* pose, float r, float b, int label, G2* out) {
  RangeBearingMeasurement z;
  z.range = r; z.bearing = b; z.label = label;
  Gaussian2D g;
  computeBirth(*(const ConstantVelocityState*)pose, z, g);
  memcpy(out, &g, sizeof(g));
}

/* ---- preUpdateSynthKernel + phdUpdateKernel on already-split in-range maps (src/phdfilter.cu:3463-3585).
 * features: concatenated in-range components, n_in[p] per particle.  Outputs use the reference layout
 * [nondetect C | detect m-major M*C | birth M] per particle at update_offset = off_p*(M+1) + p*M. */
extern "C" size_t ref_update_terms(const Pose* poses, int n_particles, const G2* features, const int* n_in, const float* z,
                                   int M, int fields, G2* terms_out, char* prune_flags_out, float* particle_weights_out) {
  set_measurements(z, M, fields);
  vector<int> offsets(n_particles + 1, 0), pose_idx;
  for (int p = 0; p < n_particles; ++p) {
    offsets[p + 1] = offsets[p] + n_in[p];
    pose_idx.insert(pose_idx.end(), n_in[p], p);                       /* :3532-3534 */
  }
  int n_total = offsets[n_particles];
  size_t n_update = (size_t)n_total * (M + 1) + (size_t)n_particles * M;
  SynthSLAM particles(n_particles);
  for (int p = 0; p < n_particles; ++p) memcpy(&particles.states[p], &poses[p], sizeof(Pose));
  measurementSet meas(Z, Z + M);
  vector<Gaussian2D> births((size_t)n_particles * M);
  births_host(particles, meas, n_particles, M, births);
  vector<Gaussian2D> preupdate((size_t)std::max(n_total, 1) * M), update(n_update);
  vector<float> pd(std::max(n_total, 1)), likelihoods((size_t)std::max(n_total, 1) * M), pw(n_particles, 0.0f);
  vector<char> flags(n_update, 0);
  pose_idx.push_back(0);
  if (n_total > 0) {
    int nb = std::min((int)ceil(n_total / 256.0), 65535);              /* :3555 */
    emul_launch(nb, kThreads, [&] {
      preUpdateSynthKernel((ConstantVelocityState*)poses, pose_idx.data(), (Gaussian2D*)features, pd.data(), n_total, M,
                           likelihoods.data(), preupdate.data());
    });
  }
  emul_launch(std::min(n_particles, 65535), kThreads, [&] {            /* :3573-3578 */
    phdUpdateKernel<Gaussian2D>((Gaussian2D*)features, pd.data(), preupdate.data(), births.data(), offsets.data(), n_particles,
                                M, update.data(), (bool*)flags.data(), pw.data());
  });
  if (terms_out) memcpy(terms_out, update.data(), n_update * sizeof(G2));
  if (prune_flags_out) memcpy(prune_flags_out, flags.data(), n_update);
  if (particle_weights_out) memcpy(particle_weights_out, pw.data(), n_particles * sizeof(float));
  return n_update;
}

/* ---- phdUpdateMergeKernel on one or more candidate lists (mergeAndCopyMaps, src/phdfilter.cu:3269-3276) ---- */
extern "C" void ref_merge(const G2* cand, const int* offsets /* n_particles+1 */, int n_particles, G2* merged_out,
                          int* merged_sizes) {
  int n = offsets[n_particles];
  vector<Gaussian2D> in(std::max(n, 1)), out(std::max(n, 1));
  memcpy(in.data(), cand, (size_t)n * sizeof(G2));
  vector<char> flags(std::max(n, 1), 0);
  emul_launch(n_particles, kThreads, [&] {
    phdUpdateMergeKernel<Gaussian2D>(in.data(), out.data(), merged_sizes, (bool*)flags.data(), (int*)offsets, n_particles);
  });
  memcpy(merged_out, out.data(), (size_t)n * sizeof(G2));   /* particle p's merged map starts at offsets[p] */
}

/* ---- mixed feature model (featureModel = MIXED_MODEL): the reference's own device code, src/phdfilter.cu:244-521,
 * 910-963, 2323-2635, plus computeMahalDist(Gaussian4D) / ConstantVelocityMotionModel of src/device_math.cuh and the
 * Gaussian4D instance of phdUpdateMergeKernel ---- */
#include "ref_mixed.inc"
typedef phdslam_gaussian4d_t G4;
static_assert(sizeof(G4) == sizeof(Gaussian4D), "layout");

extern "C" float ref_mahalanobis4(const G4* a, const G4* b) {
  Gaussian4D x, y;
  memcpy(&x, a, sizeof(x)); memcpy(&y, b, sizeof(y));
  return computeMahalDist(x, y);
}
/* predictMapMixed (src/phdfilter.cu:965-1035): launch shape and motion-model set-up of the host wrapper; the jump
 * features come back too (the wrapper discards them) */
extern "C" void ref_predict_features4(const G4* in, int n, G4* out, G2* jump_out) {
  vector<Gaussian4D> prior(std::max(n, 1)), pred(std::max(n, 1));
  vector<Gaussian2D> jump(std::max(n, 1));
  memcpy(prior.data(), in, (size_t)n * sizeof(G4));
  ConstantVelocityMotionModel motion_model;
  motion_model.std_accx = config.stdAxMap;
  motion_model.std_accy = config.stdAyMap;
  int n_blocks = (n + 255) / 256;
  emul_launch(n_blocks, kThreads, [&] { predictMapKernelMixed(prior.data(), motion_model, n, pred.data(), jump.data()); });
  memcpy(out, pred.data(), (size_t)n * sizeof(G4));
  if (jump_out) memcpy(jump_out, jump.data(), (size_t)n * sizeof(G2));
}
/* phdUpdateKernelMixed on ONE particle (the kernel reads the predicted weights without the particle's offset, :2411,2437,
 * so only particle 0 of a launch is computed as intended): the in-range static and dynamic features in, the dense update
 * terms of both maps ([non-detect | detect m-major | birth], :2344-2356) and the particle's log-weight increment out. */
extern "C" float ref_mixed_update_terms(const Pose* pose, const G2* s_in, int ns, const G4* d_in, int nd, const float* z, int M,
                                        int fields, G2* s_terms, G4* d_terms, char* s_flags, char* d_flags) {
  set_measurements(z, M, fields);
  vector<Gaussian2D> sp(std::max(ns, 1)), su((size_t)ns * (M + 1) + M + 1);
  vector<Gaussian4D> dp(std::max(nd, 1)), du((size_t)nd * (M + 1) + M + 1);
  memcpy(sp.data(), s_in, (size_t)ns * sizeof(G2));
  memcpy(dp.data(), d_in, (size_t)nd * sizeof(G4));
  int off_s[2] = {0, ns}, off_d[2] = {0, nd};
  vector<char> fs(su.size(), 0), fd(du.size(), 0);
  ConstantVelocityState ps;
  memcpy(&ps, pose, sizeof(ps));
  REAL pw = 0;
  emul_launch(1, kThreads, [&] {
    phdUpdateKernelMixed(&ps, sp.data(), dp.data(), off_s, off_d, 1, M, su.data(), du.data(), (bool*)fs.data(), (bool*)fd.data(),
                         &pw);
  });
  const size_t nst = (size_t)ns * (M + 1) + M, ndt = (size_t)nd * (M + 1) + M;
  memcpy(s_terms, su.data(), nst * sizeof(G2));
  memcpy(d_terms, du.data(), ndt * sizeof(G4));
  if (s_flags) memcpy(s_flags, fs.data(), nst);
  if (d_flags) memcpy(d_flags, fd.data(), ndt);
  return pw;
}
extern "C" int ref_merge4(const G4* cand, int n, G4* merged_out) {
  vector<Gaussian4D> in(std::max(n, 1)), out(std::max(n, 1));
  memcpy(in.data(), cand, (size_t)n * sizeof(G4));
  vector<char> flags(std::max(n, 1), 0);
  int offsets[2] = {0, n}, size = 0;
  emul_launch(1, kThreads, [&] {
    phdUpdateMergeKernel<Gaussian4D>(in.data(), out.data(), &size, (bool*)flags.data(), offsets, 1);
  });
  memcpy(merged_out, out.data(), (size_t)size * sizeof(G4));
  return size;
}

/* ---- a whole static-map phdUpdateSynth (src/phdfilter.cu:3336-3761): reference kernels + restated glue ---- */
extern "C" size_t ref_update(const Pose* poses, int n_particles, const int* map_sizes, const G2* maps, float* log_weights,
                             const float* z, int M, int fields, int* out_sizes, G2* out_maps, size_t out_cap) {
  if (M > 256) M = 256;                                                /* :3390-3394 */
  int n_total = 0;
  for (int p = 0; p < n_particles; ++p) n_total += map_sizes[p];
  vector<char> cls(std::max(n_total, 1), 0);
  vector<int> n_in(n_particles), n_near(n_particles);
  ref_in_range(maps, map_sizes, n_particles, poses, cls.data(), n_in.data(), n_near.data());
  /* host 3-way split, order preserved within each class (prepareUpdateInputs :3030-3070) */
  vector<G2> f_in, f_out1, f_out2;
  vector<int> n_out1(n_particles, 0), n_out2(n_particles, 0);
  {
    size_t k = 0;
    for (int p = 0; p < n_particles; ++p)
      for (int i = 0; i < map_sizes[p]; ++i, ++k) {
        if (cls[k] == 1) f_in.push_back(maps[k]);
        else if (cls[k] == 2) { f_out2.push_back(maps[k]); n_out2[p]++; }
        else { f_out1.push_back(maps[k]); n_out1[p]++; }
      }
  }
  size_t n_update = (size_t)f_in.size() * (M + 1) + (size_t)n_particles * M;
  vector<G2> terms(n_update);
  vector<char> flags(n_update);
  vector<float> pw(n_particles);
  f_in.push_back(G2());
  ref_update_terms(poses, n_particles, f_in.data(), n_in.data(), z, M, fields, terms.data(), flags.data(), pw.data());
  /* pruneMap (:3120-3174): stable remove_copy_if + per-particle recount; recombination with the nearly-in-range
   * features per particle (:3227-3257) */
  vector<G2> combined;
  vector<int> offsets(n_particles + 1, 0);
  {
    size_t k = 0, k2 = 0;
    for (int p = 0; p < n_particles; ++p) {
      size_t np = (size_t)n_in[p] * (M + 1) + M;
      for (size_t i = 0; i < np; ++i, ++k)
        if (!flags[k]) combined.push_back(terms[k]);
      for (int i = 0; i < n_out2[p]; ++i) combined.push_back(f_out2[k2++]);
      offsets[p + 1] = (int)combined.size();
    }
  }
  vector<G2> merged(std::max<size_t>(combined.size(), 1));
  vector<int> msizes(n_particles, 0);
  combined.push_back(G2());
  ref_merge(combined.data(), offsets.data(), n_particles, merged.data(), msizes.data());
  /* copy out + re-append the far features (:3296-3318) */
  size_t w = 0, k1 = 0;
  for (int p = 0; p < n_particles; ++p) {
    out_sizes[p] = msizes[p] + n_out1[p];
    for (int i = 0; i < msizes[p]; ++i, ++w)
      if (w < out_cap) out_maps[w] = merged[offsets[p] + i];
    for (int i = 0; i < n_out1[p]; ++i, ++w, ++k1)
      if (w < out_cap) out_maps[w] = f_out1[k1];
  }
  /* particle weights (:3735-3755) */
  vector<float> lw(log_weights, log_weights + n_particles);
  if (config.particleWeighting != 2)
    for (int p = 0; p < n_particles; ++p) lw[p] += pw[p];
  float s = logSumExp(lw);
  for (int p = 0; p < n_particles; ++p) log_weights[p] = lw[p] - s;
  return w;
}

/* ---- CPHD: the reference's multi-object update, HEAD's commented-out kernels (src/phdfilter.cu:701-748,1430-1822, made
 * live by ref_build.sh) and the live older kernels of src/phdfilter.cu.bak:369-415,1058-1478, each set in its own
 * namespace.  The launch sequence and shapes are the .bak wrapper's (phdUpdate, .bak:2503-2544; initCphdConstants,
 * .bak:418-448): cphdConstantsKernel<<<N+1, N+1>>>, cphdPreUpdateKernel<<<ceil(n_update/128), 128>>>,
 * computeEsfKernel<<<particles, M>>>, computePsiKernel<<<particles, N+1>>>, cphdUpdateKernel<<<particles, M>>>.
 * The kernels' 256-wide shared-memory reductions need max_cardinality = 255 (as the reference's default run has).
 * Layout of the update terms (cphdPreUpdateKernel): per particle at (M+1)*off_p: C*M detection terms, FEATURE-major
 * (feature j, measurement m at j*M + m), then the C non-detection terms. ---- */
namespace cphd_head {
#include "ref_cphd_head.inc"
}
namespace cphd_bak {
#include "ref_cphd_bak.inc"
}

#define CPHD_PIPELINE(NS)                                                                                              \
  emul_launch(N1, N1, [&] { NS::cphdConstantsKernel(lfact.data(), Ctab.data(), cn_clutter.data()); });                 \
  if (n_total > 0)                                                                                                     \
    emul_launch((n_update + 127) / 128, 128, [&] {                                                                     \
      NS::cphdPreUpdateKernel((Gaussian2D*)features, offsets.data(), n_particles, M, (ConstantVelocityState*)poses,    \
                              upd.data(), w_partial.data(), qdw.data());                                               \
    });                                                                                                                \
  emul_launch(n_particles, M, [&] { NS::computeEsfKernel(w_partial.data(), offsets.data(), M, esf.data(), esfd.data()); }); \
  emul_launch(n_particles, N1, [&] {                                                                                   \
    NS::computePsiKernel((Gaussian2D*)features, (REAL*)cn_predict, esf.data(), esfd.data(), offsets.data(), M, qdw.data(), \
                         lfact.data(), Ctab.data(), cn_clutter.data(), cn_update_out, ip0_out, ip1_out, ip1d_out);     \
  });                                                                                                                  \
  emul_launch(n_particles, M, [&] {                                                                                    \
    NS::cphdUpdateKernel(offsets.data(), M, ip0_out, ip1_out, ip1d_out, (bool*)flags.data(), upd.data());              \
  });

/* variant 0 = HEAD (uncommented), 1 = .bak.  features: concatenated in-range components, n_in[p] per particle;
 * cn_predict [n_particles][N+1] log predicted cardinality.  Outputs: terms_out / flags_out [(M+1) * sum n_in],
 * cn_update_out [n_particles][N+1], ip0/ip1 [n_particles] = log<Psi0,p>, log<Psi1,p>, ip1d [n_particles][M],
 * esf_out [n_particles][M+1] (log e_j), esfd_out [n_particles][M][M] (log leave-one-out e_j, M-1 used per row),
 * lambda is not exported by the kernels.  Returns the number of update terms. */
extern "C" int ref_cphd_update(int variant, const Pose* poses, int n_particles, const G2* features, const int* n_in,
                               const float* z, int M, int fields, const float* cn_predict, G2* terms_out, char* flags_out,
                               float* cn_update_out, float* ip0_out, float* ip1_out, float* ip1d_out, float* esf_out,
                               float* esfd_out, float* w_partial_out /* [sum n_in][M] partial log-weights */) {
  const int N1 = dev_config.maxCardinality + 1;
  if (N1 != 256 || M < 1 || M > 256) return -1;
  set_measurements(z, M, fields);
  vector<int> offsets(n_particles + 1, 0);
  for (int p = 0; p < n_particles; ++p) offsets[p + 1] = offsets[p] + n_in[p];
  const int n_total = offsets[n_particles];
  const int n_update = n_total * (M + 1);
  /* initCphdConstants (.bak:418-448, HEAD :751-782 commented): log-factorials on the host, tables by the kernel */
  vector<REAL> lfact(N1), Ctab((size_t)N1 * N1), cn_clutter(N1);
  lfact[0] = 0;
  for (int n = 1; n < N1; n++) lfact[n] = lfact[n - 1] + safeLog((REAL)n);
  vector<Gaussian2D> upd(std::max(n_update, 1));
  vector<REAL> w_partial((size_t)std::max(n_total, 1) * M), qdw(std::max(n_total, 1));
  vector<REAL> esf((size_t)n_particles * (M + 1)), esfd((size_t)n_particles * M * M, 0.0f);
  vector<char> flags(std::max(n_update, 1), 0);
  if (variant == 0) { CPHD_PIPELINE(cphd_head) } else { CPHD_PIPELINE(cphd_bak) }
  if (terms_out) memcpy(terms_out, upd.data(), (size_t)n_update * sizeof(G2));
  if (flags_out) memcpy(flags_out, flags.data(), (size_t)n_update);
  if (esf_out) memcpy(esf_out, esf.data(), esf.size() * sizeof(float));
  if (esfd_out) memcpy(esfd_out, esfd.data(), esfd.size() * sizeof(float));
  if (w_partial_out) memcpy(w_partial_out, w_partial.data(), (size_t)n_total * M * sizeof(float));
  return n_update;
}

/* ---- resampleParticles<SynthSLAM> (src/main.cpp:452-501); uniforms = every randu01() it consumes, in order ---- */
extern "C" void ref_resample(const float* log_weights, int n, int n_new, const double* uniforms, int* idx_out,
                             float* new_log_weights_out) {
  SynthSLAM p(n);
  p.weights.assign(log_weights, log_weights + n);
  g_uniforms = uniforms;
  g_uniform_pos = 0;
  SynthSLAM q = resampleParticles(p, n_new);
  for (int j = 0; j < q.n_particles; ++j) idx_out[j] = q.resample_idx[j];
  if (new_log_weights_out)
    for (int j = 0; j < q.n_particles; ++j) new_log_weights_out[j] = q.weights[j];
}

/* ---- recoverSlamState (src/main.cpp:318-388) with mapEstimate = 1; the MAP particle is identified by giving
 * particle i a one-component map whose weight is i ---- */
extern "C" void ref_recover(const float* log_weights, const Pose* poses, int n, Pose* expected, int* map_particle) {
  SynthSLAM p(n);
  p.weights.assign(log_weights, log_weights + n);
  for (int i = 0; i < n; ++i) {
    memcpy(&p.states[i], &poses[i], sizeof(Pose));
    Gaussian2D g;
    memset(&g, 0, sizeof(g));
    g.weight = (float)i;
    p.maps_static[i].assign(1, g);
  }
  int saved = config.mapEstimate;
  config.mapEstimate = 1;
  ConstantVelocityState e;
  vector<REAL> cn;
  recoverSlamState(p, e, cn);
  config.mapEstimate = saved;
  memcpy(expected, &e, sizeof(e));
  *map_particle = (int)p.max_map_static[0].weight;
}

/* ---- recoverSlamState with mapEstimate = 2: the EAP map, computeExpectedMap (src/main.cpp:290-316) +
 * reduceGaussianMixture (src/gm_reduce.cpp:57-134).  maps: concatenated, map_sizes[i] components of particle i. ---- */
extern "C" int ref_expected_map(const float* log_weights, const int* map_sizes, const G2* maps, int n, G2* out, int cap) {
  SynthSLAM p(n);
  p.weights.assign(log_weights, log_weights + n);
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    memset(&p.states[i], 0, sizeof(ConstantVelocityState));
    p.maps_static[i].assign((const Gaussian2D*)maps + off, (const Gaussian2D*)maps + off + map_sizes[i]);
    off += (size_t)map_sizes[i];
  }
  int saved = config.mapEstimate;
  config.mapEstimate = 2;
  ConstantVelocityState e;
  vector<REAL> cn;
  recoverSlamState(p, e, cn);
  config.mapEstimate = saved;
  const int m = (int)p.exp_map_static.size();
  for (int i = 0; i < m && i < cap; ++i) memcpy(&out[i], &p.exp_map_static[i], sizeof(G2));
  return m;
}

/* ---- writeLog (src/main.cpp:848-954): writes ./state_estimateNNNNN.log (append mode) -- run inside `dir` ---- */
extern "C" int ref_write_log(const char* dir, const Pose* expected, const G2* map, int n_map, const float* log_weights,
                             const Pose* poses, int n, const int* idx_resample, const float* cn, int n_cn, int t) {
  char cwd[4096];
  if (!getcwd(cwd, sizeof(cwd)) || chdir(dir) != 0) return -1;
  SynthSLAM p(n);
  p.weights.assign(log_weights, log_weights + n);
  for (int i = 0; i < n; ++i) memcpy(&p.states[i], &poses[i], sizeof(Pose));
  ConstantVelocityState e;
  memcpy(&e, expected, sizeof(e));
  vector<Gaussian2D> ms((const Gaussian2D*)map, (const Gaussian2D*)map + n_map);
  vector<Gaussian4D> md;
  vector<int> idx(idx_resample, idx_resample + n);
  vector<REAL> cnv(cn, cn + n_cn);
  writeLog(p, e, ms, md, idx, cnv, t);
  return chdir(cwd);
}

/* ---- text inputs: loadTimestamps / loadControls / loadMeasurements / loadTrajectory (src/main.cpp:147-264) ---- */
extern "C" int ref_load_timestamps(const char* path, float* out, int cap) {
  vector<REAL> t = loadTimestamps(path);
  for (size_t i = 0; i < t.size() && (int)i < cap; ++i) out[i] = t[i];
  return (int)t.size();
}
extern "C" int ref_load_controls(const char* path, float* out /* [cap][2] v_encoder, alpha */, int cap) {
  vector<AckermanControl> u = loadControls(path);
  for (size_t i = 0; i < u.size() && (int)i < cap; ++i) { out[2 * i] = u[i].v_encoder; out[2 * i + 1] = u[i].alpha; }
  return (int)u.size();
}
extern "C" int ref_load_measurements(const char* path, float* out /* [cap_vals][3] */, int cap_vals, int* counts, int cap_sets) {
  vector<measurementSet> all;
  loadMeasurements(path, all);
  size_t k = 0;
  for (size_t s = 0; s < all.size(); ++s) {
    if ((int)s < cap_sets) counts[s] = (int)all[s].size();
    for (size_t i = 0; i < all[s].size(); ++i, ++k)
      if ((int)k < cap_vals) { out[3 * k] = all[s][i].range; out[3 * k + 1] = all[s][i].bearing; out[3 * k + 2] = (float)all[s][i].label; }
  }
  return (int)all.size();
}
extern "C" int ref_load_trajectory(const char* path, Pose* out, int cap) {
  vector<ConstantVelocityState> t;
  loadTrajectory(path, t);
  for (size_t i = 0; i < t.size() && (int)i < cap; ++i) memcpy(&out[i], &t[i], sizeof(Pose));
  return (int)t.size();
}

/* ---- the input schedule of a time-stamped run: the has_timestamps branch of run_synth's loop (src/main.cpp:1188-1230),
 * verbatim, inside a loop with the variables run_synth declares (:1161-1168, globals :83-84).  Output per event: the
 * measurement set taken (-1 none), the control taken (-1: the current one is kept), config.dt. ---- */
static REAL current_time = 0, last_time = 0;
static void setDeviceConfig(const SlamConfig&) {}
extern "C" int ref_plan_events(const float* mt, int nz, const float* ct, int nc, int* z_out, int* c_out, float* dt_out, int cap) {
  vector<REAL> measurement_times(mt, mt + nz), control_times(ct, ct + nc);
  vector<measurementSet> allMeasurements(nz);
  vector<AckermanControl> all_controls(nc);
  measurementSet ZZ;
  AckermanControl current_control;
  current_control.alpha = 0; current_control.v_encoder = 0;
  bool do_predict = false;
  REAL dt = 0;
  unsigned int z_idx = 0;
  int c_idx = 0;
  current_time = last_time = 0;
  const int nSteps = nz + nc; /* :1116 */
  int k = 0;
  for (int n = 0; n < nSteps; n++) {
    const unsigned int z0 = z_idx;
    const int c0 = c_idx;
#include "ref_events.inc"
    if (k < cap) {
      z_out[k] = (z_idx > z0) ? (int)z0 : -1;
      c_out[k] = (c_idx > c0) ? c0 : -1;
      dt_out[k] = dt;
    }
    ++k;
  }
  (void)do_predict;
  return k;
}

/* nEff exactly as run_synth spells it (src/main.cpp:1281-1284) -- three lines, restated */
extern "C" float ref_neff(const float* log_weights, int n) {
  REAL nEff = 0;
  for (int i = 0; i < n; i++) nEff += exp(2 * log_weights[i]);
  nEff = 1.0 / nEff / n;
  return nEff;
}
