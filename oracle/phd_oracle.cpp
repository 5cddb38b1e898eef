/*
 * phd_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).
 * See phd_oracle.h for the role of this file and how parity is pinned.
 *
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -mfma -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off is REQUIRED: every a*b+c below must round twice exactly as the
 * kernels (nvcc -fmad=false) do; fused operations are written as explicit fmaf().
 */
#include "phd_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/phd_detmath.h"
#include "../include/phd_mixed_math.h"

#ifdef _OPENMP
#include <omp.h>
#endif

typedef phdslam_gaussian2d_t G2;
typedef phdslam_gaussian4d_t G4;
typedef phdslam_pose_t Pose;

struct phd_oracle {
  phdslam_config_t cfg;
  std::vector<Pose> states;                 /* ParticleSLAM::states   (src/slamtypes.h:279) */
  std::vector<float> weights;               /* ParticleSLAM::weights  (log domain) */
  std::vector<std::vector<G2>> maps;        /* SynthSLAM::maps_static (src/slamtypes.h:290) */
  std::vector<std::vector<G4>> maps_dyn;    /* SynthSLAM::maps_dynamic (:291), mixed feature model only */
  std::vector<int> resample_idx;
  std::vector<std::vector<float>> card;     /* SynthSLAM::cardinalities (log domain) */
  unsigned predict_calls = 0;               /* Philox counter words */
  unsigned resample_calls = 0;
  int threads = 1;
};

/* ------------------------------------------------------------------------- */
/* helpers                                                                    */
/* ------------------------------------------------------------------------- */

/* Largest float <= d and smallest float >= d: lets a reference comparison made in double
 * (`r <= 1.2*dev_config.maxRange`, src/phdfilter.cu:1340-1342) be done exactly in fp32. */
static float float_floor(double d) {
  float f = (float)d;
  if ((double)f > d) f = nextafterf(f, -INFINITY);
  return f;
}
static float float_ceil(double d) {
  float f = (float)d;
  if ((double)f < d) f = nextafterf(f, INFINITY);
  return f;
}

/* Canonical reduction shape of the kernels: each of the 32 lanes owns two adjacent elements of every
 * 64-element chunk and keeps one sequential partial sum per element slot (a packed fp32x2 accumulator),
 * i.e. 64 strided partials p[i] = sum_k v[64k+i]; the two slots of a lane are added, then an
 * xor-butterfly (offsets 16,8,4,2,1) runs over the lanes.  Stands in for the reference's 256-wide
 * shared-memory tree (sumByReduction, src/device_math.cuh:452-472). */
static float warp_sum(const float* v, int n) {
  float p[64];
  for (int l = 0; l < 64; ++l) {
    float acc = 0.0f;
    for (int i = l; i < n; i += 64) acc = acc + v[i];
    p[l] = acc;
  }
  float q[32];
  for (int l = 0; l < 32; ++l) q[l] = p[2 * l] + p[2 * l + 1];
  for (int off = 16; off >= 1; off >>= 1) {
    float t[32];
    for (int l = 0; l < 32; ++l) t[l] = q[l] + q[l ^ off];
    memcpy(q, t, sizeof(q));
  }
  return q[0];
}

extern "C" float oracle_warp_sum(const float* v, int n) { return warp_sum(v, n); }

/* ------------------------------------------------------------------------- */
/* lifetime / state access                                                    */
/* ------------------------------------------------------------------------- */

extern "C" phd_oracle_t* oracle_create(const phdslam_config_t* cfg) {
  phd_oracle* o = new phd_oracle();
  o->cfg = *cfg;
  int n = cfg->n_particles;
  /* run_synth initialisation, src/main.cpp:1129-1144 */
  Pose p0 = {cfg->x0, cfg->y0, cfg->yaw0, cfg->vx0, cfg->vy0, cfg->vyaw0};
  o->states.assign(n, p0);
  o->weights.assign(n, -phd_logf((float)n)); /* -log(float(n_particles)), main.cpp:1144 */
  o->maps.assign(n, std::vector<G2>());
  o->maps_dyn.assign(n, std::vector<G4>());
  o->resample_idx.resize(n);
  for (int i = 0; i < n; ++i) o->resample_idx[i] = i;
  o->card.assign(n, std::vector<float>());
  if (cfg->filter_type == 1) {
    int nc = cfg->max_cardinality + 1;
    for (int i = 0; i < n; ++i) o->card[i].assign(nc, -phd_logf((float)nc)); /* main.cpp:1140-1143 */
  }
  return o;
}
extern "C" void oracle_destroy(phd_oracle_t* o) { delete o; }
extern "C" void oracle_set_config(phd_oracle_t* o, const phdslam_config_t* cfg) { o->cfg = *cfg; }
extern "C" void oracle_set_threads(phd_oracle_t* o, int n) { o->threads = n < 1 ? 1 : n; }
extern "C" int oracle_n_particles(const phd_oracle_t* o) { return (int)o->states.size(); }
extern "C" void oracle_get_poses(const phd_oracle_t* o, Pose* out) { memcpy(out, o->states.data(), o->states.size() * sizeof(Pose)); }
extern "C" void oracle_set_poses(phd_oracle_t* o, const Pose* in) { memcpy(o->states.data(), in, o->states.size() * sizeof(Pose)); }
extern "C" void oracle_get_log_weights(const phd_oracle_t* o, float* out) { memcpy(out, o->weights.data(), o->weights.size() * 4); }
extern "C" void oracle_set_log_weights(phd_oracle_t* o, const float* in) { memcpy(o->weights.data(), in, o->weights.size() * 4); }
extern "C" void oracle_get_map_sizes(const phd_oracle_t* o, int* out) {
  for (size_t i = 0; i < o->maps.size(); ++i) out[i] = (int)o->maps[i].size();
}
extern "C" void oracle_get_maps(const phd_oracle_t* o, G2* out) {
  size_t k = 0;
  for (auto& m : o->maps)
    for (auto& g : m) out[k++] = g;
}
extern "C" void oracle_set_maps(phd_oracle_t* o, const int* sizes, const G2* in) {
  size_t k = 0;
  for (size_t i = 0; i < o->maps.size(); ++i) {
    o->maps[i].assign(in + k, in + k + sizes[i]);
    k += sizes[i];
  }
}
extern "C" void oracle_get_map_sizes_dynamic(const phd_oracle_t* o, int* out) {
  for (size_t i = 0; i < o->maps_dyn.size(); ++i) out[i] = (int)o->maps_dyn[i].size();
}
extern "C" void oracle_get_maps_dynamic(const phd_oracle_t* o, G4* out) {
  size_t k = 0;
  for (const auto& m : o->maps_dyn)
    for (const G4& g : m) out[k++] = g;
}
extern "C" void oracle_set_maps_dynamic(phd_oracle_t* o, const int* sizes, const G4* in) {
  size_t k = 0;
  for (size_t i = 0; i < o->maps_dyn.size(); ++i) {
    o->maps_dyn[i].assign(in + k, in + k + sizes[i]);
    k += sizes[i];
  }
}
extern "C" void oracle_get_resample_idx(const phd_oracle_t* o, int* out) { memcpy(out, o->resample_idx.data(), o->resample_idx.size() * 4); }
extern "C" void oracle_get_cardinalities(const phd_oracle_t* o, float* out) {
  int nc = o->cfg.max_cardinality + 1;
  for (size_t i = 0; i < o->card.size(); ++i)
    if ((int)o->card[i].size() == nc) memcpy(out + i * nc, o->card[i].data(), nc * 4);
}
extern "C" void oracle_set_cardinalities(phd_oracle_t* o, const float* in) {
  int nc = o->cfg.max_cardinality + 1;
  for (size_t i = 0; i < o->card.size(); ++i) o->card[i].assign(in + i * nc, in + (i + 1) * nc);
}

/* ------------------------------------------------------------------------- */
/* predict                                                                    */
/* ------------------------------------------------------------------------- */

/* phdPredict host wrapper (src/phdfilter.cu:1080-1257) + phdPredictKernelAckerman (:785-825)
 * + phdPredictKernel (:827-859).  One call = one sub-step; dt = config.dt / subdividePredict.
 * Noise draws: injected (reference call order, :1113-1117 / :1148-1152) or Philox4x32-10. */
extern "C" void oracle_predict(phd_oracle_t* o, const float* control, const double* draws) {
  const phdslam_config_t& c = o->cfg;
  if (c.n_predict_particles > 1) {
    /* "shotgun" prediction (:1091, :796-797, :1185-1238): every particle spawns nPredictParticles predictions
     * (prediction j descends from particle j / k); maps, cardinalities and resample indices are duplicated and the
     * weights scaled down by k */
    const int k = c.n_predict_particles, n0 = (int)o->states.size();
    std::vector<Pose> ns((size_t)n0 * k);
    std::vector<float> nw((size_t)n0 * k);
    std::vector<std::vector<G2>> nm((size_t)n0 * k);
    std::vector<std::vector<G4>> nd((size_t)n0 * k);
    std::vector<std::vector<float>> nc((size_t)n0 * k);
    std::vector<int> nr((size_t)n0 * k);
    const float logk = phd_safe_log((float)k);
    for (int j = 0; j < n0 * k; ++j) {
      const int i = j / k;
      ns[j] = o->states[i];
      nw[j] = o->weights[i] - logk;              /* :1208 */
      nm[j] = o->maps[i];
      nd[j] = o->maps_dyn[i];
      nc[j] = o->card[i];
      nr[j] = o->resample_idx[i];
    }
    o->maps_dyn.swap(nd);
    o->states.swap(ns); o->weights.swap(nw); o->maps.swap(nm); o->card.swap(nc); o->resample_idx.swap(nr);
  }
  int n = (int)o->states.size();
  float dt = c.dt / (float)c.subdivide_predict; /* REAL dt = dev_config.dt/dev_config.subdividePredict */
  unsigned call = o->predict_calls++;
  for (int i = 0; i < n; ++i) {
    Pose s = o->states[i];
    Pose ns;
    float sn, cs;
    phd_sincosf(s.ptheta, &sn, &cs);
    if (c.motion_type == 1) {
      double d_alpha, d_enc;
      if (draws) {
        d_alpha = draws[2 * i];
        d_enc = draws[2 * i + 1];
      } else {
        phd_philox4_t r = phd_philox4x32_10((uint32_t)i, call, PHD_STREAM_PREDICT, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32));
        float z0, z1;
        phd_box_muller(r.v[0], r.v[1], &z0, &z1);
        d_alpha = (double)z0;
        d_enc = (double)z1;
      }
      /* noiseVector[i].n_alpha = config.stdAlpha * randn() : float*double, stored as float (:1150-1151) */
      float n_alpha = (float)((double)c.std_alpha * d_alpha);
      float n_enc = (float)((double)c.std_encoder * d_enc);
      float v_enc = control ? control[0] : 0.0f, alpha = control ? control[1] : 0.0f;
      float ve = v_enc + n_enc;                       /* :804 */
      float al = alpha + n_alpha;                     /* :805 */
      float ta = phd_tanf(al);
      float vc = ve / (1.0f - ta * c.h / c.l);        /* :806 */
      float xc_dot = vc * cs;                         /* :807 */
      float yc_dot = vc * sn;                         /* :808 */
      float thc_dot = vc * ta / c.l;                  /* :809 */
      ns.px = s.px + dt * (xc_dot - thc_dot * (c.a * sn + c.b * cs));   /* :811-814 */
      ns.py = s.py + dt * (yc_dot + thc_dot * (c.a * cs - c.b * sn));   /* :815-818 */
      ns.ptheta = phd_wrap_angle(s.ptheta + dt * thc_dot);              /* :819 */
      ns.vx = 0.0f; ns.vy = 0.0f; ns.vtheta = 0.0f;                     /* :820-822 */
    } else {
      double d0, d1, d2;
      if (draws) {
        d0 = draws[3 * i]; d1 = draws[3 * i + 1]; d2 = draws[3 * i + 2];
      } else {
        phd_philox4_t r = phd_philox4x32_10((uint32_t)i, call, PHD_STREAM_PREDICT, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32));
        float z0, z1, z2, z3;
        phd_box_muller(r.v[0], r.v[1], &z0, &z1);
        phd_box_muller(r.v[2], r.v[3], &z2, &z3);
        d0 = z0; d1 = z1; d2 = z2;
      }
      /* noiseVector[i].ax = 3*config.ax * randn() (:1115-1117; note the factor 3) */
      float nax = (float)((double)(3.0f * c.ax) * d0);
      float nay = (float)((double)(3.0f * c.ay) * d1);
      float nat = (float)((double)(3.0f * c.ayaw) * d2);
      float hdt2 = dt * dt * 0.5f;
      ns.px = s.px + dt * (s.vx * cs - s.vy * sn) + hdt2 * (nax * cs - nay * sn);      /* :843-847 */
      ns.py = s.py + dt * (s.vx * sn + s.vy * cs) + hdt2 * (nax * sn + nay * cs);      /* :848-852 */
      ns.ptheta = phd_wrap_angle(s.ptheta + dt * s.vtheta + hdt2 * nat);               /* :853-855 */
      ns.vx = s.vx + dt * nax;                                                         /* :856-858 */
      ns.vy = s.vy + dt * nay;
      ns.vtheta = s.vtheta + dt * nat;
    }
    o->states[i] = ns;
  }
  /* map prediction of the dynamic features (:1240-1242 -> predictMapMixed :965-1035 -> predictMapKernelMixed :910-963).
   * Reference quirks kept: the kernel steps by the WHOLE dev_config.dt on every phdPredict call (so subdividePredict = k
   * predicts the map k times by dt); the "jump" features (a dynamic feature turning static) are computed and thrown away
   * (the insertion into maps_static is commented out, :1015-1021), so the static maps are untouched. */
  if (c.feature_model == 2) {
    const float var_x = c.std_ax_features * c.std_ax_features, var_y = c.std_ay_features * c.std_ay_features;
    for (int i = 0; i < n; ++i)
      for (G4& g : o->maps_dyn[i]) {
        G4 q;
        phd_g4_predict(&g, c.dt, var_x, var_y, c.ps, c.beta, c.tau, &q);
        g = q;
      }
  }
}

/* ------------------------------------------------------------------------- */
/* GM-PHD update                                                              */
/* ------------------------------------------------------------------------- */

/* computeInRangeKernel classification (src/phdfilter.cu:1328-1346):
 * 1 = in the field of view, 2 = "nearly" (bypasses the update, joins the merge), 0 = far. */
static int classify(const phdslam_config_t& c, const Pose& pose, const G2& f) {
  float dx = f.mean[0] - pose.px;
  float dy = f.mean[1] - pose.py;
  float r2 = dx * dx + dy * dy;
  float r = sqrtf(r2);
  float bearing = phd_wrap_angle(phd_atan2f(dy, dx) - pose.ptheta);
  if (r >= c.min_range && r <= c.max_range && fabsf(bearing) <= c.max_bearing) return 1;
  /* 0.8*minRange etc. are double products in the reference; exact fp32 equivalents: */
  float lo2 = float_ceil(0.8 * (double)c.min_range);
  float hi2 = float_floor(1.2 * (double)c.max_range);
  float hb2 = float_floor(1.2 * (double)c.max_bearing);
  if (r >= lo2 && r <= hi2 && fabsf(bearing) <= hb2) return 2;
  return 0;
}

/* birth Gaussian from one measurement (host loop, src/phdfilter.cu:3468-3507) */
static G2 birth_term(const phdslam_config_t& c, const Pose& pose, float zr, float zb, int label) {
  G2 b;
  float theta = pose.ptheta + zb;
  float sn, cs;
  phd_sincosf(theta, &sn, &cs);
  float dx = zr * cs;
  float dy = zr * sn;
  b.mean[0] = pose.px + dx;
  b.mean[1] = pose.py + dy;
  float J0 = dx / zr, J1 = dy / zr, J2 = -dy, J3 = dx;
  float sr = c.std_range * c.birth_noise_factor;
  float sb = c.std_bearing * c.birth_noise_factor;
  float var_range = sr * sr;      /* pow(config.stdRange*config.birthNoiseFactor,2) */
  float var_bearing = sb * sb;
  b.cov[0] = J0 * J0 * var_range + J2 * J2 * var_bearing;
  b.cov[1] = J0 * J1 * var_range + J2 * J3 * var_bearing;
  b.cov[2] = b.cov[1];
  b.cov[3] = J1 * J1 * var_range + J3 * J3 * var_bearing;
  if (label == 0 || !c.labeled_measurements)
    b.weight = phd_safe_log(c.birth_weight);
  else
    b.weight = phd_safe_log(0.0f);
  return b;
}

struct UpdateOut {
  std::vector<G2> terms;  /* [nondetect C | detect m-major M*C | birth M] */
  int n_in = 0;
  float dlogw = 0.0f;
};

/* preUpdateSynthKernel (src/phdfilter.cu:1824-1925) followed by phdUpdateKernel (:2083-2321)
 * for ONE particle.  `in` are the in-range (class 1) components in map order. */
static void cphd_factors(const phdslam_config_t& c, const float* w, const float* pd, int C, const float* S, int M,
                         const float* prior, int N1, float* D, float* ND, float* inc, float* card_out);

/* Mixed feature model (phdUpdateKernelMixed, src/phdfilter.cu:2323-2608): the static and the dynamic features of a
 * particle share the per-measurement normaliser and the predicted cardinality; the dynamic side hands these in. */
struct MixHook {
  const float* dsum;   /* [M] sum over the in-range dynamic features of exp(partial log-weight) (:2470-2471) */
  float nhat_dyn;      /* sum of pd * w over the in-range dynamic features (:2424-2446) */
  float* L_out;        /* [M] the log normalisers, for the dynamic terms */
};

static void update_particle(const phdslam_config_t& c, const Pose& pose, const std::vector<G2>& in, const float* z,
                            int M, int fields, UpdateOut& out, std::vector<float>* card = nullptr,
                            const MixHook* mix = nullptr) {
  const int C = (int)in.size();
  const int T = C * (M + 1) + M;
  out.n_in = C;
  out.terms.assign(T, G2());
  std::vector<float> pdv(C), logdet(C), rj(C), bj(C);
  std::vector<float> K(4 * C), S(4 * C), covu(4 * C);
  const float var_r = c.std_range * c.std_range;      /* pow(dev_config.stdRange,2) */
  const float var_b = c.std_bearing * c.std_bearing;
  for (int i = 0; i < C; ++i) {
    const G2& f = in[i];
    float dx = f.mean[0] - pose.px;
    float dy = f.mean[1] - pose.py;
    float r2 = dx * dx + dy * dy;
    float r = sqrtf(r2);
    float bearing = phd_wrap_angle(phd_atan2f(dy, dx) - pose.ptheta);
    float pd = 0.0f;
    if (r <= c.max_range && fabsf(bearing) <= c.max_bearing) pd = c.pd;   /* :1841-1844 (no minRange test) */
    float J[4];
    J[0] = dx / r; J[2] = dy / r; J[1] = -dy / r2; J[3] = dx / r2;          /* :1847-1851 */
    const float* P = f.cov;
    float sigma[4];                                                        /* :1857-1861 */
    sigma[0] = (P[0] * J[0] + J[2] * P[1]) * J[0] + (J[0] * P[2] + P[3] * J[2]) * J[2] + var_r;
    sigma[1] = (P[0] * J[1] + J[3] * P[1]) * J[0] + (J[1] * P[2] + P[3] * J[3]) * J[2];
    sigma[2] = (P[0] * J[0] + J[2] * P[1]) * J[1] + (J[0] * P[2] + P[3] * J[2]) * J[3];
    sigma[3] = (P[0] * J[1] + J[3] * P[1]) * J[1] + (J[1] * P[2] + P[3] * J[3]) * J[3] + var_b;
    sigma[1] = (sigma[1] + sigma[2]) / 2.0f;                               /* :1864-1865 */
    sigma[2] = sigma[1];
    float det = sigma[0] * sigma[3] - sigma[1] * sigma[2];                 /* :1867 */
    float* Si = &S[4 * i];
    Si[0] = sigma[3] / det; Si[1] = -sigma[1] / det; Si[2] = -sigma[2] / det; Si[3] = sigma[0] / det;
    float* Ki = &K[4 * i];                                                 /* :1877-1881 */
    Ki[0] = Si[0] * (P[0] * J[0] + P[2] * J[2]) + Si[1] * (P[0] * J[1] + P[2] * J[3]);
    Ki[1] = Si[0] * (P[1] * J[0] + P[3] * J[2]) + Si[1] * (P[1] * J[1] + P[3] * J[3]);
    Ki[2] = Si[2] * (P[0] * J[0] + P[2] * J[2]) + Si[3] * (P[0] * J[1] + P[2] * J[3]);
    Ki[3] = Si[2] * (P[1] * J[0] + P[3] * J[2]) + Si[3] * (P[1] * J[1] + P[3] * J[3]);
    /* Joseph-form covariance, Maple-expanded in the reference (:1883-1887).  a,b,cc,d = I - K*J */
    float a = 1.0f - Ki[0] * J[0] - Ki[2] * J[1];
    float b = -Ki[0] * J[2] - Ki[2] * J[3];
    float cc = -Ki[1] * J[0] - Ki[3] * J[1];
    float d = 1.0f - Ki[1] * J[2] - Ki[3] * J[3];
    float* cu = &covu[4 * i];
    cu[0] = (a * P[0] + b * P[1]) * a + (a * P[2] + b * P[3]) * b + Ki[0] * Ki[0] * var_r + Ki[2] * Ki[2] * var_b;
    cu[2] = (a * P[0] + b * P[1]) * cc + (a * P[2] + b * P[3]) * d + Ki[0] * var_r * Ki[1] + Ki[2] * var_b * Ki[3];
    cu[1] = (cc * P[0] + d * P[1]) * a + (cc * P[2] + d * P[3]) * b + Ki[0] * var_r * Ki[1] + Ki[2] * var_b * Ki[3];
    cu[3] = (cc * P[0] + d * P[1]) * cc + (cc * P[2] + d * P[3]) * d + Ki[1] * Ki[1] * var_r + Ki[3] * Ki[3] * var_b;
    pdv[i] = pd; logdet[i] = phd_safe_log(det); rj[i] = r; bj[i] = bearing;
    /* non-detection term (:2137-2141) */
    G2& nd = out.terms[i];
    nd = f;
    nd.weight = f.weight * (1.0f - pd);
  }
  /* detection terms, measurement-major (:1898-1923, copied by :2144-2151).  Canonical arithmetic of the
   * per-(component, measurement) inner loop: every multiply-add is an explicit fused multiply-add (the
   * reference's own build contracts them too: nvcc defaults to -fmad=true); <= 1 ulp from the unfused text. */
  for (int m = 0; m < M; ++m) {
    float zr = z[m * fields], zb = z[m * fields + 1];
    int label = fields > 2 ? (int)z[m * fields + 2] : 0;
    for (int i = 0; i < C; ++i) {
      const G2& f = in[i];
      const float* Ki = &K[4 * i];
      const float* Si = &S[4 * i];
      float innov0 = zr - rj[i];
      float innov1 = phd_wrap_angle(zb - bj[i]);
      G2& t = out.terms[C + m * C + i];
      t.mean[0] = fmaf(Ki[2], innov1, fmaf(Ki[0], innov0, f.mean[0]));
      t.mean[1] = fmaf(Ki[3], innov1, fmaf(Ki[1], innov0, f.mean[1]));
      for (int n = 0; n < 4; ++n) t.cov[n] = covu[4 * i + n];
      float dist = (innov0 * innov0) * Si[0];
      dist = fmaf(innov0 * innov1, Si[1] + Si[2], dist);
      dist = fmaf(innov1 * innov1, Si[3], dist);
      float g = fmaf(dist, -0.5f, -PHD_LOG_2PI_F) + (-(0.5f * logdet[i]));        /* :1911 */
      if (label == 0 || !c.labeled_measurements)
        t.weight = (phd_safe_log(pdv[i]) + phd_safe_log(f.weight)) + g;            /* :1916-1917, log domain */
      else
        t.weight = phd_safe_log(0.0f);
    }
  }
  /* births (:3468-3507), log(birthWeight) */
  for (int m = 0; m < M; ++m) {
    int label = fields > 2 ? (int)z[m * fields + 2] : 0;
    out.terms[C + M * C + m] = birth_term(c, pose, z[m * fields], z[m * fields + 1], label);
  }
  if (card) {
    /* CPHD: cphdPreUpdateKernel (:1430-1511) computes the same partial log-weights; the multi-object terms replace
     * the per-measurement normaliser (cphd_factors), cphdUpdateKernel (:1780-1822) applies them */
    std::vector<float> ev(std::max(C, 1)), S(M), D(M), wv(std::max(C, 1));
    for (int m = 0; m < M; ++m) {
      G2* det = &out.terms[C + m * C];
      for (int i = 0; i < C; ++i) ev[i] = phd_expf(det[i].weight);
      S[m] = (C > 0) ? warp_sum(ev.data(), C) : 0.0f;
    }
    for (int i = 0; i < C; ++i) wv[i] = in[i].weight;
    float ND = 0.0f, inc = 0.0f;
    std::vector<float> card_out(card->size());
    cphd_factors(c, wv.data(), pdv.data(), C, S.data(), M, card->data(), (int)card->size(), D.data(), &ND, &inc,
                 card_out.data());
    for (int m = 0; m < M; ++m) {
      G2* det = &out.terms[C + m * C];
      for (int i = 0; i < C; ++i) det[i].weight = phd_expf(det[i].weight + D[m]);
      G2& bt = out.terms[C + M * C + m];
      bt.weight = phd_expf(bt.weight + D[m]);
    }
    for (int i = 0; i < C; ++i)                                                  /* :1810-1813 */
      out.terms[i].weight = phd_expf((phd_safe_log(in[i].weight) + ND) + phd_safe_log(1.0f - pdv[i]));
    out.dlogw = inc;
    card->swap(card_out);
    return;
  }
  /* predicted cardinality: sum of pd*w over features then birthWeight per measurement (:2133-2186) */
  std::vector<float> tmp(C + M);
  for (int i = 0; i < C; ++i) tmp[i] = pdv[i] * in[i].weight;
  for (int m = 0; m < M; ++m) tmp[C + m] = c.birth_weight;
  /* mixed model: the birth terms do not enter (val stays 0 for them, :2398-2405,2427-2436), the dynamic features do */
  float cardinality_predict = mix ? (warp_sum(tmp.data(), C) + mix->nhat_dyn) : warp_sum(tmp.data(), C + M);
  /* per-measurement normaliser and final weights (:2190-2252) */
  float particle_weight = 0.0f;
  std::vector<float> ev(C), dsum(M);
  for (int m = 0; m < M; ++m) {
    G2* det = &out.terms[C + m * C];
    for (int i = 0; i < C; ++i) ev[i] = phd_expf(det[i].weight);
    float sum = (C > 0) ? warp_sum(ev.data(), C) : 0.0f;
    if (mix) sum = sum + mix->dsum[m];
    sum = sum + c.clutter_density;
    sum = sum + c.birth_weight;
    if (mix && !c.labeled_measurements) sum = sum + c.birth_weight;   /* "we get 2 birth terms when measurements are unlabeled", :2478-2480 */
    float log_normalizer = phd_safe_log(sum);
    if (mix) mix->L_out[m] = log_normalizer;
    /* exp(logw - L) of the reference (:2222-2229), evaluated as exp(logw) * exp(-L): the first factor is the term that
     * was just summed, so the kernel does not evaluate a second exponential per update term (canonical, <= 2 ulp) */
    const float scale = phd_expf(-log_normalizer);
    for (int i = 0; i < C; ++i) {
      det[i].weight = ev[i] * scale;
      ev[i] = det[i].weight;
    }
    G2& bt = out.terms[C + M * C + m];
    bt.weight = phd_expf(bt.weight - log_normalizer);
    dsum[m] = ((C > 0) ? warp_sum(ev.data(), C) : 0.0f) + bt.weight;
    particle_weight = particle_weight + log_normalizer;                     /* :2250 */
  }
  if (c.particle_weighting == 0) {
    out.dlogw = particle_weight - cardinality_predict;                      /* :2260-2262 */
  } else if (c.particle_weighting == 1) {
    /* Vo empty-map weighting (:2264-2279).  The reference sums serially on thread 0; canonical order:
     * warp_sum per block of terms, blocks accumulated in term order. */
    std::vector<float> w(std::max(C, 1));
    for (int i = 0; i < C; ++i) w[i] = in[i].weight;
    float cn_predict = warp_sum(w.data(), C);
    for (int i = 0; i < C; ++i) w[i] = out.terms[i].weight;
    float cn_update = warp_sum(w.data(), C);
    for (int m = 0; m < M; ++m) cn_update = cn_update + dsum[m];
    out.dlogw = (float)M * c.clutter_density + cn_update - cn_predict - c.clutter_rate;
  } else {
    out.dlogw = 0.0f; /* scheme 2 is unfinished in the reference's device code (:2281-2305) */
  }
}

/* computeMahalDist (src/device_math.cuh:309-325) with invert_matrix2 (:61-70).  Canonical deviation: the
 * four divisions by det are one reciprocal and four products (<= 1 ulp per entry). */
static float mahal(const G2& a, const G2& b) {
  float sigma[4], inv[4];
  for (int i = 0; i < 4; ++i) sigma[i] = (a.cov[i] + b.cov[i]) / 2.0f;
  float det = sigma[0] * sigma[3] - sigma[2] * sigma[1];
  float rdet = 1.0f / det;
  inv[0] = sigma[3] * rdet; inv[1] = -sigma[1] * rdet; inv[2] = -sigma[2] * rdet; inv[3] = sigma[0] * rdet;
  float i0 = a.mean[0] - b.mean[0];
  float i1 = a.mean[1] - b.mean[1];
  return i0 * i0 * inv[0] + i0 * i1 * (inv[1] + inv[2]) + i1 * i1 * inv[3];
}

/* computeHellingerDist (src/device_math.cuh:380-413) */
static float hellinger(const G2& a, const G2& b) {
  float innov0 = a.mean[0] - b.mean[0], innov1 = a.mean[1] - b.mean[1];
  float s[4] = {a.cov[0] + b.cov[0], a.cov[1] + b.cov[1], a.cov[2] + b.cov[2], a.cov[3] + b.cov[3]};
  float det = s[0] * s[3] - s[2] * s[1];
  float inv[4] = {1.0f, 0.0f, 0.0f, 1.0f};
  if (det > FLT_MIN) {
    inv[0] = s[3] / det; inv[1] = -s[1] / det; inv[2] = -s[2] / det; inv[3] = s[0] / det;
  }
  float eps = -0.25f * (innov0 * innov0 * inv[0] + innov0 * innov1 * (inv[1] + inv[2]) + innov1 * innov1 * inv[3]);
  det = det / 4.0f;
  float dist = 1.0f / det;
  float p[4];
  p[0] = a.cov[0] * b.cov[0] + a.cov[2] * b.cov[1];
  p[1] = a.cov[1] * b.cov[0] + a.cov[3] * b.cov[1];
  p[2] = a.cov[0] * b.cov[2] + a.cov[2] * b.cov[3];
  p[3] = a.cov[1] * b.cov[2] + a.cov[3] * b.cov[3];
  float detp = p[0] * p[3] - p[2] * p[1];
  dist = dist * sqrtf(detp);
  dist = 1.0f - sqrtf(dist) * phd_expf(eps);
  return dist;
}

extern "C" float oracle_mahalanobis(const G2* a, const G2* b) { return mahal(*a, *b); }
extern "C" float oracle_hellinger(const G2* a, const G2* b) { return hellinger(*a, *b); }

/* phdUpdateMergeKernel (src/phdfilter.cu:2707-2898): greedy Gaussian-mixture reduction.
 * Arg-max ties are broken exactly as the reference's reduction tree breaks them (merge_tie_key).
 * Canonical choice where the reference's rounding is shape-dependent: cluster sums accumulate sequentially in
 * ascending index order (the reference: 256 strided partials + shared-memory tree).
 * Canonical gate (Mahalanobis metric only): candidate b can join the cluster of seed a only if
 *     |mu_a - mu_b|^2 <= (0.515625 * minSeparation) * (lam_a + lam_b),
 * lam = largest eigenvalue of the component's covariance.  For positive semi-definite covariances
 * d_M^2 >= |d|^2 / lam_max((Pa+Pb)/2) >= |d|^2 / ((lam_a+lam_b)/2), so a candidate outside the gate has
 * d_M^2 > 1.03*minSeparation and the reference would not merge it either; the gate (with its 3 % margin
 * against fp32 rounding) only changes results for numerically degenerate covariances.  It lets the kernel
 * look at a 3x3 neighbourhood of a uniform grid (cell size >= the largest gate radius) instead of at every
 * candidate. */
static float merge_lambda_max(const G2& g) {
  float t = g.cov[0] + g.cov[3];
  float det = g.cov[0] * g.cov[3] - g.cov[1] * g.cov[2];
  float disc = fmaxf(t * t - 4.0f * det, 0.0f);
  return 0.5f * (t + sqrtf(disc));
}

/* Arg-max tie rule of the reference's reduction (:2751-2787), which is deterministic: thread `tid` scans the
 * candidates tid, tid+256, ... and keeps the first maximum (strict <); the 256-slot tree then prefers the
 * lower slot at every level s = 128, 64, ..., 1, i.e. slots are compared by their bit-reversed number.
 * Among equal weights the winner is therefore the candidate with the smallest (bitrev8(i mod 256), i div 256).
 * (Exact ties are common: every clutter-born component has the weight w_b/(kappa + w_b).) */
static inline unsigned merge_tie_key(int i) {
  unsigned t = (unsigned)i & 255u, r = 0;
  for (int b = 0; b < 8; ++b) r |= ((t >> b) & 1u) << (7 - b);
  return (r << 24) | ((unsigned)i >> 8);
}

static void merge_mixture(const phdslam_config_t& c, const std::vector<G2>& cand, std::vector<G2>& out) {
  const int n = (int)cand.size();
  const bool gated = (c.distance_metric == 0);
  const float gk = 0.515625f * c.min_separation;
  std::vector<float> lam(n);
  for (int i = 0; i < n; ++i) lam[i] = merge_lambda_max(cand[i]);
  std::vector<char> merged(n, 0);
  std::vector<int> members;
  while (true) {
    int best = -1;
    for (int i = 0; i < n; ++i) {                                                     /* :2751-2787 */
      if (merged[i]) continue;
      if (best < 0 || cand[best].weight < cand[i].weight ||
          (cand[best].weight == cand[i].weight && merge_tie_key(i) < merge_tie_key(best)))
        best = i;
    }
    if (best < 0) break;
    const G2& mx = cand[best];
    members.clear();
    for (int i = 0; i < n; ++i) {
      if (merged[i]) continue;
      float gx = mx.mean[0] - cand[i].mean[0], gy = mx.mean[1] - cand[i].mean[1];
      if (gated && !(gx * gx + gy * gy <= gk * (lam[best] + lam[i]))) continue;
      float dist = (c.distance_metric == 0) ? mahal(mx, cand[i]) : hellinger(mx, cand[i]);  /* :2802-2805 */
      if (dist < c.min_separation) members.push_back(i);
    }
    G2 mg;
    memset(&mg, 0, sizeof(mg));
    float wsum = 0.0f, m0 = 0.0f, m1 = 0.0f;
    for (int i : members) {                                             /* :2808-2811 */
      wsum = wsum + cand[i].weight;
      m0 = m0 + cand[i].weight * cand[i].mean[0];
      m1 = m1 + cand[i].weight * cand[i].mean[1];
    }
    if (wsum == 0.0f) break;                                            /* :2821-2822 */
    mg.weight = wsum;
    const float rw = 1.0f / wsum;      /* canonical: one reciprocal instead of the reference's six divisions */
    mg.mean[0] = m0 * rw;                                               /* :2823-2829 */
    mg.mean[1] = m1 * rw;
    float cv[4] = {0, 0, 0, 0};
    for (int i : members) {                                             /* :2836-2872 */
      float d[2] = {mg.mean[0] - cand[i].mean[0], mg.mean[1] - cand[i].mean[1]};
      for (int j = 0; j < 2; ++j)
        for (int k = 0; k < 2; ++k) cv[j * 2 + k] = cv[j * 2 + k] + cand[i].weight * (cand[i].cov[j * 2 + k] + d[j] * d[k]);
      merged[i] = 1;
    }
    for (int j = 0; j < 4; ++j) mg.cov[j] = cv[j] * rw;                 /* :2874-2881 */
    mg.cov[1] = (mg.cov[1] + mg.cov[2]) / 2.0f;                         /* force_symmetric_covariance, device_math.cuh:710-725 */
    mg.cov[2] = mg.cov[1];
    out.push_back(mg);
  }
}

extern "C" int oracle_merge(const phdslam_config_t* cfg, const G2* in, int n, G2* out) {
  std::vector<G2> cand(in, in + n), res;
  merge_mixture(*cfg, cand, res);
  memcpy(out, res.data(), res.size() * sizeof(G2));
  return (int)res.size();
}

/* prepareUpdateInputs split (src/phdfilter.cu:3030-3070) */
static void split_map(const phdslam_config_t& c, const Pose& pose, const std::vector<G2>& map, std::vector<G2>& in,
                      std::vector<G2>& out2, std::vector<G2>& out1) {
  for (const G2& f : map) {
    int cls = classify(c, pose, f);
    if (cls == 1) in.push_back(f);
    else if (cls == 2) out2.push_back(f);
    else out1.push_back(f);
  }
}

extern "C" size_t oracle_update_terms(phd_oracle_t* o, const float* z, int M, int fields, G2* terms_out, size_t cap,
                                      int* n_in_range_out, float* dlogw_out) {
  if (M > 256) M = 256;
  size_t k = 0;
  for (size_t p = 0; p < o->states.size(); ++p) {
    std::vector<G2> in, out2, out1;
    split_map(o->cfg, o->states[p], o->maps[p], in, out2, out1);
    UpdateOut u;
    std::vector<float> card_copy;
    if (o->cfg.filter_type == 1) card_copy = o->card[p];          /* the dense terms are a query: state is not advanced */
    update_particle(o->cfg, o->states[p], in, z, M, fields, u, (o->cfg.filter_type == 1) ? &card_copy : nullptr);
    if (n_in_range_out) n_in_range_out[p] = u.n_in;
    if (dlogw_out) dlogw_out[p] = u.dlogw;
    if (terms_out && k + u.terms.size() <= cap) memcpy(terms_out + k, u.terms.data(), u.terms.size() * sizeof(G2));
    k += u.terms.size();
  }
  return k;
}

/* ------------------------------------------------------------------------- */
/* mixed feature model: the dynamic (constant-velocity) features of a particle  */
/* ------------------------------------------------------------------------- */

/* First half (before the static update): pre-update constants of the in-range dynamic features, the sum of their
 * likelihood terms per measurement and their share of the predicted cardinality. */
static void dyn_pre(const phdslam_config_t& c, const Pose& pose, const std::vector<G4>& in, const float* z, int M, int fields,
                    std::vector<phd_g4_pre_t>& pre, std::vector<float>& dsum, float* nhat) {
  const int C = (int)in.size();
  const float var_r = c.std_range * c.std_range, var_b = c.std_bearing * c.std_bearing;
  pre.resize(C);
  std::vector<float> tmp(std::max(C, 1));
  for (int i = 0; i < C; ++i) {
    phd_g4_preupdate(pose.px, pose.py, pose.ptheta, &in[i], c.max_range, c.max_bearing, c.pd, var_r, var_b, &pre[i]);
    tmp[i] = pre[i].pd * in[i].weight;
  }
  *nhat = (C > 0) ? warp_sum(tmp.data(), C) : 0.0f;
  dsum.assign(M, 0.0f);
  for (int m = 0; m < M; ++m) {
    const int label = fields > 2 ? (int)z[m * fields + 2] : 0;
    const int dead = c.labeled_measurements && label != 1;          /* DYNAMIC_MEASUREMENT, :504 */
    for (int i = 0; i < C; ++i)
      tmp[i] = phd_expf(phd_g4_detect(&pre[i], &in[i], z[m * fields], z[m * fields + 1], dead, nullptr));
    dsum[m] = (C > 0) ? warp_sum(tmp.data(), C) : 0.0f;
  }
}

/* Second half (after the static update fixed the normalisers L): the dense dynamic update terms in the reference's order
 * [non-detect C | detect m-major M*C | birth M] (:2414-2446, :2490-2528).  upd_sum / prior_sum: the sums Vo's empty-map
 * weighting takes over them (:2546-2563). */
static void dyn_terms(const phdslam_config_t& c, const Pose& pose, const std::vector<G4>& in,
                      const std::vector<phd_g4_pre_t>& pre, const float* z, int M, int fields, const float* L,
                      std::vector<G4>& terms, float* upd_sum, float* prior_sum) {
  const int C = (int)in.size();
  terms.assign((size_t)C * (M + 1) + M, G4());
  std::vector<float> tmp(std::max(C, 1));
  for (int i = 0; i < C; ++i) {
    terms[i] = in[i];
    terms[i].weight = in[i].weight * (1.0f - pre[i].pd);
    tmp[i] = terms[i].weight;
  }
  float upd = (C > 0) ? warp_sum(tmp.data(), C) : 0.0f;
  const float sr = c.std_range * c.birth_noise_factor, sb = c.std_bearing * c.birth_noise_factor;
  for (int m = 0; m < M; ++m) {
    const int label = fields > 2 ? (int)z[m * fields + 2] : 0;
    const int dead = c.labeled_measurements && label != 1;
    for (int i = 0; i < C; ++i) {
      G4& t = terms[C + (size_t)m * C + i];
      const float lw = phd_g4_detect(&pre[i], &in[i], z[m * fields], z[m * fields + 1], dead, t.mean);
      memcpy(t.cov, pre[i].cov, sizeof(t.cov));
      t.weight = phd_expf(lw - L[m]);
      tmp[i] = t.weight;
    }
    G4& b = terms[C + (size_t)M * C + m];
    phd_g4_birth(pose.px, pose.py, pose.ptheta, z[m * fields], z[m * fields + 1], sr * sr, sb * sb, c.cov_vx_birth,
                 c.cov_vy_birth, &b);
    b.weight = phd_expf((dead ? PHD_LOG0 : phd_safe_log(c.birth_weight)) - L[m]);
    upd = upd + (((C > 0) ? warp_sum(tmp.data(), C) : 0.0f) + b.weight);
  }
  for (int i = 0; i < C; ++i) tmp[i] = in[i].weight;
  *prior_sum = (C > 0) ? warp_sum(tmp.data(), C) : 0.0f;
  *upd_sum = upd;
}

/* phdUpdateMergeKernel<Gaussian4D> (src/phdfilter.cu:2707-2898): as merge_mixture.  The Hellinger metric has no 4-D form
 * in the reference (the template returns 0, src/device_math.cuh:366-371), so with distance_metric = 1 every candidate
 * joins the first cluster.  Canonical gate (Mahalanobis metric), the 4-D form of merge_mixture's: candidate b can join
 * the cluster of seed a only if, for the position pair AND for the velocity pair of the state,
 *     |d|^2 <= (0.515625 * minSeparation) * (tr P_a + tr P_b)      (d, P: that pair's difference and 2 x 2 block).
 * The Mahalanobis distance of a sub-vector under its marginal never exceeds that of the whole vector, and the trace
 * bounds the largest eigenvalue of a positive semi-definite block, so a candidate outside either gate has d_M^2 > 1.03 *
 * minSeparation and the reference would not merge it either.  It spares the kernel a 4 x 4 factorisation for most pairs. */
static void merge_mixture4(const phdslam_config_t& c, const std::vector<G4>& cand, std::vector<G4>& out) {
  const int n = (int)cand.size();
  const bool gated = (c.distance_metric == 0);
  const float gk = 0.515625f * c.min_separation;
  std::vector<float> trp(n), trv(n);
  for (int i = 0; i < n; ++i) {
    trp[i] = cand[i].cov[0] + cand[i].cov[5];
    trv[i] = cand[i].cov[10] + cand[i].cov[15];
  }
  std::vector<char> merged(n, 0);
  std::vector<int> members;
  while (true) {
    int best = -1;
    for (int i = 0; i < n; ++i) {
      if (merged[i]) continue;
      if (best < 0 || cand[best].weight < cand[i].weight ||
          (cand[best].weight == cand[i].weight && merge_tie_key(i) < merge_tie_key(best)))
        best = i;
    }
    if (best < 0) break;
    members.clear();
    for (int i = 0; i < n; ++i) {
      if (merged[i]) continue;
      if (i == best) {                       /* the seed opens its own cluster (distance 0) */
        members.push_back(i);
        continue;
      }
      if (gated) {
        const float g0 = cand[best].mean[0] - cand[i].mean[0], g1 = cand[best].mean[1] - cand[i].mean[1];
        const float g2 = cand[best].mean[2] - cand[i].mean[2], g3 = cand[best].mean[3] - cand[i].mean[3];
        if (!(g0 * g0 + g1 * g1 <= gk * (trp[best] + trp[i]))) continue;
        if (!(g2 * g2 + g3 * g3 <= gk * (trv[best] + trv[i]))) continue;
      }
      const float dist = (c.distance_metric == 0) ? phd_g4_mahal(&cand[best], &cand[i]) : 0.0f;
      if (dist < c.min_separation) members.push_back(i);
    }
    G4 mg;
    if (!phd_g4_moment_match(cand.data(), members.data(), (int)members.size(), &mg)) break;
    for (int i : members) merged[i] = 1;
    out.push_back(mg);
  }
}

extern "C" int oracle_merge4(const phdslam_config_t* cfg, const G4* in, int n, G4* out) {
  std::vector<G4> cand(in, in + n), res;
  merge_mixture4(*cfg, cand, res);
  memcpy(out, res.data(), res.size() * sizeof(G4));
  return (int)res.size();
}
extern "C" float oracle_mahalanobis4(const G4* a, const G4* b) { return phd_g4_mahal(a, b); }
extern "C" void oracle_predict_feature4(const phdslam_config_t* c, const G4* in, G4* out) {
  phd_g4_predict(in, c->dt, c->std_ax_features * c->std_ax_features, c->std_ay_features * c->std_ay_features, c->ps, c->beta,
                 c->tau, out);
}

static void split_map4(const phdslam_config_t& c, const Pose& pose, const std::vector<G4>& map, std::vector<G4>& in) {
  for (const G4& f : map) {
    G2 q;
    memset(&q, 0, sizeof(q));
    q.mean[0] = f.mean[0];
    q.mean[1] = f.mean[1];
    if (classify(c, pose, q) == 1) in.push_back(f);   /* everything else is dropped, :3713-3719 */
  }
}

/* One particle of phdUpdateSynth with featureModel = MIXED_MODEL (:3449-3462, :3703-3726).  The dense terms of both maps
 * come back when asked for (the pin against the reference's kernel). */
static float update_particle_mixed(const phdslam_config_t& c, const Pose& pose, std::vector<G2>& smap, std::vector<G4>& dmap,
                                   const float* z, int M, int fields, std::vector<G2>* s_terms = nullptr,
                                   std::vector<G4>* d_terms = nullptr, bool advance = true) {
  std::vector<G2> in, out2, out1;
  split_map(c, pose, smap, in, out2, out1);
  std::vector<G4> din;
  split_map4(c, pose, dmap, din);
  std::vector<phd_g4_pre_t> pre;
  std::vector<float> dsum, L(M);
  float nhat = 0.0f;
  dyn_pre(c, pose, din, z, M, fields, pre, dsum, &nhat);
  MixHook hook{dsum.data(), nhat, L.data()};
  UpdateOut u;
  update_particle(c, pose, in, z, M, fields, u, nullptr, &hook);
  std::vector<G4> dt;
  float upd = 0.0f, prior = 0.0f;
  dyn_terms(c, pose, din, pre, z, M, fields, L.data(), dt, &upd, &prior);
  float dlogw = u.dlogw;
  if (c.particle_weighting == 1) dlogw = dlogw + ((upd - prior) - (float)M * c.birth_weight);   /* :2546-2566 */
  if (s_terms) *s_terms = u.terms;
  if (d_terms) *d_terms = dt;
  if (!advance) return dlogw;
  std::vector<G2> cand;
  for (const G2& t : u.terms)
    if (!(t.weight < c.min_feature_weight)) cand.push_back(t);
  cand.insert(cand.end(), out2.begin(), out2.end());
  std::vector<G2> mergedv;
  merge_mixture(c, cand, mergedv);
  mergedv.insert(mergedv.end(), out1.begin(), out1.end());
  smap.swap(mergedv);
  std::vector<G4> dcand, dmerged;
  for (const G4& t : dt)
    if (!(t.weight < c.min_feature_weight)) dcand.push_back(t);
  merge_mixture4(c, dcand, dmerged);
  dmap.swap(dmerged);
  return dlogw;
}

/* dense update terms of ONE particle (maps given as they are, classified inside); returns the log-weight increment */
extern "C" float oracle_mixed_terms(const phdslam_config_t* cfg, const Pose* pose, const G2* smap, int ns, const G4* dmap,
                                    int nd, const float* z, int M, int fields, G2* s_terms_out, int* n_s_terms, G4* d_terms_out,
                                    int* n_d_terms) {
  std::vector<G2> sm(smap, smap + ns), st;
  std::vector<G4> dm(dmap, dmap + nd), dt;
  const float dl = update_particle_mixed(*cfg, *pose, sm, dm, z, M, fields, &st, &dt, false);
  if (s_terms_out) memcpy(s_terms_out, st.data(), st.size() * sizeof(G2));
  if (d_terms_out) memcpy(d_terms_out, dt.data(), dt.size() * sizeof(G4));
  *n_s_terms = (int)st.size();
  *n_d_terms = (int)dt.size();
  return dl;
}

/* Canonical order-independent log-sum-exp: fixed-point accumulation of exp(w - max).
 * Stands in for the host logSumExp (src/device_math.cuh:549-558), a sequential fp32 sum. */
static float log_sum_exp_fx(const std::vector<float>& w) {
  float mx = -FLT_MAX;
  for (float v : w) mx = (v > mx) ? v : mx;
  uint64_t acc = 0;
  for (float v : w) acc += phd_fx_from_unit(phd_expf(v - mx), PHD_FX_WEIGHT_BITS);
  float sumf = (float)((double)acc * (1.0 / (double)((uint64_t)1 << PHD_FX_WEIGHT_BITS)));
  return phd_safe_log(sumf) + mx;
}

/* phdUpdateSynth (src/phdfilter.cu:3336-3761) */
extern "C" void oracle_update(phd_oracle_t* o, const float* z, int M, int fields) {
  const phdslam_config_t& c = o->cfg;
  if (M <= 0) return;           /* main.cpp:1258 */
  if (M > 256) M = 256;         /* :3390-3394 */
  const int n = (int)o->states.size();
  std::vector<float> dlogw(n, 0.0f);
#pragma omp parallel for schedule(dynamic, 8) num_threads(o->threads)
  for (int p = 0; p < n; ++p) {
    if (c.feature_model == 2) {
      dlogw[p] = update_particle_mixed(c, o->states[p], o->maps[p], o->maps_dyn[p], z, M, fields);
      continue;
    }
    std::vector<G2> in, out2, out1;
    split_map(c, o->states[p], o->maps[p], in, out2, out1);
    UpdateOut u;
    update_particle(c, o->states[p], in, z, M, fields, u, (c.filter_type == 1) ? &o->card[p] : nullptr);
    dlogw[p] = u.dlogw;
    /* pruneMap (:3120-3174): stable removal of terms with weight < minFeatureWeight (flags :2308-2319) */
    std::vector<G2> cand;
    for (const G2& t : u.terms)
      if (!(t.weight < c.min_feature_weight)) cand.push_back(t);
    /* recombine with the nearly-in-range features (:3227-3257) and merge (:3276) */
    cand.insert(cand.end(), out2.begin(), out2.end());
    std::vector<G2> mergedv;
    merge_mixture(c, cand, mergedv);
    /* re-append the far features (:3311-3318) */
    mergedv.insert(mergedv.end(), out1.begin(), out1.end());
    o->maps[p].swap(mergedv);
  }
  /* particle weights (:3735-3755) */
  if (c.particle_weighting != 2)
    for (int p = 0; p < n; ++p) o->weights[p] = o->weights[p] + dlogw[p];
  float lse = log_sum_exp_fx(o->weights);
  for (int p = 0; p < n; ++p) o->weights[p] = o->weights[p] - lse;
}

/* ------------------------------------------------------------------------- */
/* state extraction                                                           */
/* ------------------------------------------------------------------------- */

/* recoverSlamState (src/main.cpp:318-388) + nEff (main.cpp:1281-1284).  Sums over particles use the
 * fixed-point encodings of phd_detmath.h so they do not depend on summation order / GPU count. */
extern "C" void oracle_estimate(phd_oracle_t* o, phdslam_estimate_t* out) {
  const int n = (int)o->states.size();
  int64_t acc[6] = {0, 0, 0, 0, 0, 0};
  uint64_t e2 = 0;
  float maxw = -FLT_MAX;
  int max_idx = -1;
  for (int i = 0; i < n; ++i) {
    float w = o->weights[i];
    float ew = phd_expf(w);                                         /* REAL exp_weight = exp(weights[i]) */
    const float* s = &o->states[i].px;
    for (int k = 0; k < 6; ++k) acc[k] += phd_fx_from_prod(ew, s[k], PHD_FX_POSE_BITS);
    e2 += phd_fx_from_unit(phd_expf(2.0f * w), PHD_FX_NEFF_BITS);   /* nEff += exp(2*weights[i]) */
    if (w > maxw) { maxw = w; max_idx = i; }                        /* main.cpp:349-356, strict > */
  }
  float* e = &out->expected_pose.px;
  const double inv = 1.0 / (double)((uint64_t)1 << PHD_FX_POSE_BITS);
  if (n > 1) {
    for (int k = 0; k < 6; ++k) e[k] = (float)((double)acc[k] * inv);
  } else {
    out->expected_pose = o->states[0];                              /* main.cpp:381-387 */
    max_idx = 0;
  }
  out->map_particle = max_idx;
  out->max_log_weight = maxw;
  double s2 = (double)e2 * (1.0 / (double)((uint64_t)1 << PHD_FX_NEFF_BITS));
  out->neff = (float)(1.0 / s2 / (double)n);                        /* nEff = 1.0/nEff/n_particles */
}

/* 2x2 Cholesky-based Mahalanobis of gm_reduce.cpp:31-38: L L^T = (Pa+Pb)/2, x = L^-1 d, |x|^2 */
static float mahal_llt(const G2& a, const G2& b) {
  float s00 = 0.5f * (a.cov[0] + b.cov[0]);
  float s10 = 0.5f * (a.cov[1] + b.cov[1]);
  float s11 = 0.5f * (a.cov[3] + b.cov[3]);
  float d0 = a.mean[0] - b.mean[0], d1 = a.mean[1] - b.mean[1];
  float l00 = sqrtf(s00);
  float l10 = s10 / l00;
  float l11 = sqrtf(s11 - l10 * l10);
  float x0 = d0 / l00;
  float x1 = (d1 - l10 * x0) / l11;
  return x0 * x0 + x1 * x1;
}

/* reduceGaussianMixture<Gaussian2D> (src/gm_reduce.cpp:57-134).  std::sort is not stable in the
 * reference; canonical order: weight descending, ties by original index. */
static void reduce_mixture(const std::vector<G2>& in, float min_distance, std::vector<G2>& out) {
  const int n = (int)in.size();
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return in[a].weight > in[b].weight; });
  std::vector<char> gone(n, 0);
  std::vector<int> mergel;
  for (int oi = 0; oi < n; ++oi) {
    int s = order[oi];
    if (gone[s]) continue;
    gone[s] = 1;
    const G2& mx = in[s];
    mergel.clear();
    for (int oj = oi + 1; oj < n; ++oj) {                            /* :86-101 */
      int t = order[oj];
      if (gone[t]) continue;
      if (mahal_llt(mx, in[t]) < min_distance) { mergel.push_back(t); gone[t] = 1; }
    }
    G2 mg = mx;                                                      /* :104-110 */
    float w = mx.weight;
    float m0 = mx.mean[0] * mx.weight, m1 = mx.mean[1] * mx.weight;
    for (int t : mergel) {
      m0 = m0 + in[t].weight * in[t].mean[0];
      m1 = m1 + in[t].weight * in[t].mean[1];
      w = w + in[t].weight;
    }
    mg.weight = w;
    mg.mean[0] = m0 / w;
    mg.mean[1] = m1 / w;
    float d0 = mg.mean[0] - mx.mean[0], d1 = mg.mean[1] - mx.mean[1];   /* :111-113 */
    float cv[4];
    cv[0] = mx.weight * (mx.cov[0] + d0 * d0);
    cv[1] = mx.weight * (mx.cov[1] + d1 * d0);
    cv[2] = mx.weight * (mx.cov[2] + d0 * d1);
    cv[3] = mx.weight * (mx.cov[3] + d1 * d1);
    for (int t : mergel) {                                            /* :115-119 */
      d0 = mg.mean[0] - in[t].mean[0];
      d1 = mg.mean[1] - in[t].mean[1];
      cv[0] = cv[0] + in[t].weight * (in[t].cov[0] + d0 * d0);
      cv[1] = cv[1] + in[t].weight * (in[t].cov[1] + d1 * d0);
      cv[2] = cv[2] + in[t].weight * (in[t].cov[2] + d0 * d1);
      cv[3] = cv[3] + in[t].weight * (in[t].cov[3] + d1 * d1);
    }
    for (int k = 0; k < 4; ++k) mg.cov[k] = cv[k] / w;               /* :120 */
    out.push_back(mg);
  }
}

extern "C" int oracle_reduce_mixture(const G2* in, int n, float min_distance, G2* out) {
  std::vector<G2> v(in, in + n), r;
  reduce_mixture(v, min_distance, r);
  memcpy(out, r.data(), r.size() * sizeof(G2));
  return (int)r.size();
}

/* which = 1: MAP map (main.cpp:344-361); which = 2: EAP map, computeExpectedMap (main.cpp:290-316) */
extern "C" int oracle_map_estimate(phd_oracle_t* o, int which, G2* out, int cap) {
  std::vector<G2> res;
  if (which == 1) {
    phdslam_estimate_t e;
    oracle_estimate(o, &e);
    res = o->maps[e.map_particle];
  } else {
    std::vector<G2> concat;
    for (size_t p = 0; p < o->maps.size(); ++p) {
      float ew = phd_expf(o->weights[p]);
      for (G2 g : o->maps[p]) {
        g.weight = g.weight * ew;                                    /* map[i].weight *= exp(weights[n]) */
        concat.push_back(g);
      }
    }
    if (!concat.empty()) reduce_mixture(concat, o->cfg.min_separation, res);
  }
  int n = (int)std::min<size_t>(res.size(), (size_t)cap);
  memcpy(out, res.data(), n * sizeof(G2));
  return (int)res.size();
}

/* ------------------------------------------------------------------------- */
/* resampling                                                                 */
/* ------------------------------------------------------------------------- */

/* resampleParticles (src/main.cpp:453-501) + SynthSLAM::copy_particles (src/slamtypes.h:313-333).
 * HEAD is *stratified*: one uniform is drawn and discarded, then a fresh one per offspring.
 *   literal   : the reference's sequential double CDF walk, verbatim semantics.
 *   canonical : integer CDF of Q40 weights; ancestor(j) = min{ i : C_i > floor(r_j * C_total) }.
 *               Order-independent, hence identical for any GPU count; agrees with `literal`
 *               except when r_j falls within ~1e-12 of a CDF step (tests/test_oracle_kat.py). */
extern "C" void oracle_resample(phd_oracle_t* o, int n_new, const double* uniforms, int literal, int* ancestors_out) {
  const phdslam_config_t& c = o->cfg;
  const int n = (int)o->states.size();
  if (n_new < 0) n_new = n;
  std::vector<int> idx(n_new, 0);
  const double interval = 1.0 / (double)n_new;
  unsigned call = o->resample_calls++;
  auto draw = [&](int j) -> double {
    if (uniforms) return (c.resample_mode == 1) ? uniforms[0] : uniforms[1 + j];
    uint32_t ctr = (c.resample_mode == 1) ? 0u : (uint32_t)j;
    phd_philox4_t r = phd_philox4x32_10(ctr, call, PHD_STREAM_RESAMPLE, 0u, (uint32_t)c.seed, (uint32_t)(c.seed >> 32));
    return phd_u01d(r.v[0], r.v[1]);
  };
  if (literal) {
    double cdf = (double)expf(o->weights[0]);
    int i = 0;
    for (int j = 0; j < n_new; ++j) {
      double r = j * interval + draw(j) * interval;
      while (r > cdf) {
        i++;
        if (i >= n) {                                                /* :475-494 overflow guard */
          double mw = -1;
          int mi = -1;
          for (int k = 0; k < n; ++k)
            if ((double)expf(o->weights[k]) > mw) { mw = (double)expf(o->weights[k]); mi = k; }
          i = mi;
          cdf = 2;
          break;
        }
        cdf += (double)expf(o->weights[i]);
      }
      idx[j] = i;
    }
  } else {
    std::vector<uint64_t> cdf(n);
    uint64_t run = 0;
    for (int i = 0; i < n; ++i) {
      run += phd_fx_from_unit(phd_expf(o->weights[i]), PHD_FX_CDF_BITS);
      cdf[i] = run;
    }
    const uint64_t total = run;
    for (int j = 0; j < n_new; ++j) {
      double r = (double)j * interval + draw(j) * interval;
      double t = floor(r * (double)total);
      uint64_t R = (t <= 0.0) ? 0 : (uint64_t)t;
      if (R >= total) R = total - 1;
      idx[j] = (int)(std::upper_bound(cdf.begin(), cdf.end(), R) - cdf.begin());
    }
  }
  /* copy_particles: deep copy, every weight = -log(N_new), resample_idx = indices */
  std::vector<Pose> ns(n_new);
  std::vector<std::vector<G2>> nm(n_new);
  std::vector<std::vector<float>> nc(n_new);
  std::vector<std::vector<G4>> nd(n_new);
  for (int j = 0; j < n_new; ++j) {
    ns[j] = o->states[idx[j]];
    nm[j] = o->maps[idx[j]];
    nd[j] = o->maps_dyn[idx[j]];
    nc[j] = o->card[idx[j]];
  }
  o->maps_dyn.swap(nd);
  o->states.swap(ns);
  o->maps.swap(nm);
  o->card.swap(nc);
  o->weights.assign(n_new, -phd_logf((float)n_new));
  o->resample_idx = idx;
  if (ancestors_out) memcpy(ancestors_out, idx.data(), n_new * sizeof(int));
}

/* run_synth loop body (src/main.cpp:1231-1297), in the two halves either side of the point where the reference looks at
 * the state (recoverSlamState outputs, writeParticlesMat: :1274-1279) */
extern "C" void oracle_step_filter(phd_oracle_t* o, int step_index, const float* control, const float* z, int M, int fields,
                                   phdslam_estimate_t* est_out) {
  const phdslam_config_t& c = o->cfg;
  if (step_index > 0)
    for (int i = 0; i < c.subdivide_predict; ++i) oracle_predict(o, control, nullptr);   /* :1244-1255 */
  if (M > 0) oracle_update(o, z, M, fields);                                             /* :1258-1272 */
  oracle_estimate(o, est_out);                                                           /* :1274, :1281-1284 */
}
extern "C" void oracle_step_resample(phd_oracle_t* o, int M, const phdslam_estimate_t* e, int* resampled_out) {
  const phdslam_config_t& c = o->cfg;
  int n = (int)o->states.size();
  int res = 0;
  if ((e->neff <= c.resample_threshold && M > 0) || n > 5 * c.n_particles) {             /* :1286 */
    oracle_resample(o, c.n_particles, nullptr, 0, nullptr);
    res = 1;
  } else {
    for (int i = 0; i < n; ++i) o->resample_idx[i] = i;                                  /* :1293-1296 */
  }
  if (resampled_out) *resampled_out = res;
}
extern "C" void oracle_step(phd_oracle_t* o, int step_index, const float* control, const float* z, int M, int fields,
                            phdslam_estimate_t* est_out, int* resampled_out) {
  phdslam_estimate_t e;
  oracle_step_filter(o, step_index, control, z, M, fields, &e);
  oracle_step_resample(o, M, &e, resampled_out);
  if (est_out) *est_out = e;
}

/* ------------------------------------------------------------------------- */
/* CPHD (spec: commented kernels src/phdfilter.cu:701-748,1430-1822; live older forms in       */
/* src/phdfilter.cu.bak:369-448,518-545,779-791,1058-1504,2473-2544)                            */
/* ------------------------------------------------------------------------- */

/* Elementary symmetric functions e_0..e_n of `roots` (Vieta recursion of computeEsfKernel,
 * src/phdfilter.cu:1553-1576, in double). */
extern "C" void oracle_esf(const double* roots, int n, double* out) {
  out[0] = 1.0;
  for (int i = 1; i <= n; ++i) out[i] = 0.0;
  for (int m = 0; m < n; ++m)
    for (int k = m + 1; k >= 1; --k) out[k] = fma(roots[m], out[k - 1], out[k]);   /* one rounding per fold, as the kernel's DFMA */
}

/* k * x with the convention 0 * x = 0 even for x = LOG0 (the reference relies on exp() of huge negatives) */
static inline float mulk(int k, float x) { return k == 0 ? 0.0f : (float)k * x; }
static inline float lclamp(float t) { return (t < PHD_LOG0) ? PHD_LOG0 : t; }
/* log of a non-negative double, as float: exact frexp, float log of the mantissa */
static inline float logd(double v) {
  if (!(v > 0.0)) return PHD_LOG0;
  int ex;
  double mant = frexp(v, &ex);
  return phd_logf((float)mant) + (float)ex * 0.693147182f;
}
/* exp of a double argument as a double with float accuracy and double RANGE: t = k ln2 + r, exp(t) = 2^k * expf(r).
 * Only IEEE double operations, the deterministic float exp and an exact scaling: identical in the kernel. */
static inline double expd(double t) {
  if (t != t) return t;
  if (!(t > -700.0)) return 0.0;
  if (t > 700.0) t = 700.0;
  const double kd = rint(t * 1.4426950408889634);
  const double r = fma(-kd, 0.6931471805599453, t);
  return ldexp((double)phd_expf((float)r), (int)kd);
}

/* log-sum-exp with the kernels' warp reduction shape: max, then warp_sum of exp(t - max) */
static float lse_warp(float* t, int n) {
  if (n <= 0) return PHD_LOG0;
  float mx = t[0];
  for (int i = 1; i < n; ++i) mx = (t[i] > mx) ? t[i] : mx;
  for (int i = 0; i < n; ++i) t[i] = phd_expf(t[i] - mx);
  return phd_safe_log(warp_sum(t, n)) + mx;
}

/* Birth cardinality Binomial(M, w_b) (src/phdfilter.cu.bak:779-791) and the predicted cardinality of
 * cardinalityPredictKernel (src/phdfilter.cu:867-888), p-(n) = log sum_j exp(birth(n-j) + prior(j)): the reference's plain
 * sum of exponentials, canonically evaluated as the convolution of the two pmfs (N1 + M + 1 exponentials instead of
 * N1 (M + 1)); terms with n-j > M are exactly 0.  lf: log-factorials 0..max(N, M); out: pb[M+1], pm[N1]. */
static void cphd_predict_cardinality(const phdslam_config_t& c, const float* prior, int N1, int M, const std::vector<float>& lf,
                                     std::vector<float>& pb, std::vector<float>& pm) {
  const int N = N1 - 1;
  const float lwb = phd_safe_log(c.birth_weight), l1wb = phd_safe_log(1.0f - c.birth_weight);
  pb.assign(M + 1, 0.0f);
  for (int k = 0; k <= M; ++k) {
    float t = lf[M] - lf[k];
    t = t - lf[M - k];
    t = t + mulk(k, lwb);
    t = t + mulk(M - k, l1wb);
    pb[k] = t;
  }
  pm.assign(N1, 0.0f);
  std::vector<float> epb(M + 1), epr(N1);
  for (int k = 0; k <= M; ++k) epb[k] = phd_expf(pb[k]);
  for (int n = 0; n <= N; ++n) epr[n] = phd_expf(prior[n]);
  for (int n = 0; n <= N; ++n) {
    float sum = 0.0f;
    for (int j = std::max(0, n - M); j <= n; ++j) sum = fmaf(epb[n - j], epr[j], sum);
    pm[n] = (sum != 0.0f) ? phd_safe_log(sum) : PHD_LOG0;
  }
}

static std::vector<float> cphd_log_factorials(int nlf) {
  std::vector<float> lf(nlf);
  lf[0] = 0.0f;
  for (int k = 1; k < nlf; ++k) lf[k] = lf[k - 1] + phd_safe_log((float)k);           /* .bak:2474-2479 */
  return lf;
}

/* test entry: pb_out[M+1] birth cardinality, pm_out[N1] predicted cardinality (pinned against the reference's
 * cardinalityPredictKernel in tests/test_ref_pin.py) */
extern "C" void oracle_cphd_predict_cardinality(const phdslam_config_t* cfg, const float* prior, int N1, int M, float* pb_out,
                                                float* pm_out) {
  std::vector<float> lf = cphd_log_factorials(std::max(N1 - 1, M) + 1), pb, pm;
  cphd_predict_cardinality(*cfg, prior, N1, M, lf, pb, pm);
  memcpy(pb_out, pb.data(), pb.size() * sizeof(float));
  memcpy(pm_out, pm.data(), pm.size() * sizeof(float));
}

/*
 * CPHD multi-object terms for ONE particle (Vo, Vo & Cantoni 2007), following the reference's kernels:
 *   cardinalityPredictKernel (src/phdfilter.cu:867-888, live) with the binomial birth cardinality of
 *   birthsKernel (src/phdfilter.cu.bak:779-791); computeEsfKernel (:1524-1618, commented at HEAD);
 *   computePsiKernel (:1626-1769); cphdUpdateKernel (:1780-1822).
 * Canonical evaluation: fp32 log domain as in the reference, except
 *   - the elementary symmetric functions run in double on roots scaled by the largest one (the reference's fp32
 *     linear recursion overflows beyond a handful of measurements; its .bak log-domain form needs |.|);
 *   - the sums over n and j of <Psi1, p>, <Psi1d_m, p> are exchanged (A1[j] below), which turns the reference's
 *     O(N M^2) into O(N M + M^2); identical in exact arithmetic;
 *   - the (N+1) x (M+1) tables Psi0(n), A1[j] and the predicted cardinality are convolutions and are evaluated in the
 *     linear domain (double, one fused multiply-add per term; float products for the cardinality prediction) instead of
 *     one exponential of a log-domain sum per term; identical in exact arithmetic, float accuracy;
 *   - <Psi1d_m, p> takes the leave-one-out elementary symmetric functions from the full ones by composite deflation
 *     (O(M) per measurement) instead of recomputing the recursion per measurement (O(M^2));
 *   - factorial[k] + cn_clutter[k] is spelled k*log(clutterRate) - clutterRate (:735-737, :1691-1692).
 * HEAD creates the birth terms at update time (one per measurement, weight w_b, always detected), so measurement m's
 * likelihood mass is S_m + w_b (as in the PHD normaliser, :2213-2214), <1,w> includes M*w_b and <q_D,w> does not.
 *   in : w[C], pd[C] in-range components; S[M] = sum_j pd_j w_j g_j(z_m); prior[N1] log cardinality
 *   out: D[M] log factor of measurement m's detection and birth terms; *ND log factor of the non-detection terms;
 *        *inc particle log-weight increment log<Psi0, p>; card_out[N1] updated log cardinality
 */
static void cphd_factors(const phdslam_config_t& c, const float* w, const float* pd, int C, const float* S, int M,
                         const float* prior, int N1, float* D, float* ND, float* inc, float* card_out) {
  const int N = N1 - 1;
  const int nlf = std::max(N, M) + 1;
  const std::vector<float> lf = cphd_log_factorials(nlf);
  std::vector<float> pb, pm;
  cphd_predict_cardinality(c, prior, N1, M, lf, pb, pm);
  const float lcr = phd_safe_log(c.clutter_rate), lcd = phd_safe_log(c.clutter_density);
  const float larea = lcr - lcd;
  /* log lambda_m (:1539-1552) with the birth mass of measurement m */
  std::vector<float> llam(M);
  float lmax = PHD_LOG0;
  for (int m = 0; m < M; ++m) {
    llam[m] = phd_safe_log(S[m] + c.birth_weight) + larea;
    lmax = (llam[m] > lmax) ? llam[m] : lmax;
  }
  std::vector<double> x(M);
  for (int m = 0; m < M; ++m) x[m] = (double)phd_expf(llam[m] - lmax);
  /* full and leave-one-out elementary symmetric functions (:1553-1616) */
  std::vector<double> e(M + 1);
  oracle_esf(x.data(), M, e.data());
  std::vector<float> le(M + 1);
  for (int j = 0; j <= M; ++j) le[j] = logd(e[j]) + mulk(j, lmax);
  /* <q_D, w> and <1, w> (:1649-1683) */
  std::vector<float> tmp(std::max(C, 1));
  for (int j = 0; j < C; ++j) tmp[j] = w[j] * (1.0f - pd[j]);
  const float q = warp_sum(tmp.data(), C);
  const float Wsum = warp_sum(w, C) + (float)M * c.birth_weight;
  const float lq = phd_safe_log(q);
  const float lW = (Wsum > 0.0f) ? phd_logf(Wsum) : 0.0f;
  std::vector<float> cK(M + 1);
  for (int k = 0; k <= M; ++k) cK[k] = mulk(k, lcr) - c.clutter_rate;
  /* The two (N+1) x (M+1) tables of the update -- Psi0(n) and A1[j] below -- are sums of products n!/(n-j)! q^(n-j) ...,
   * i.e. CONVOLUTIONS with d[k] = (q s)^k / k!.  The reference evaluates every term as one exponential of a log-domain
   * sum (:1686-1703, :1706-1764); canonically they are evaluated in the linear domain in double, one fused multiply-add
   * per term and 3 (N+1) + M + 1 exponentials in all.  s = n_c / <1,w> centres k!/n_c^k so that every factor stays
   * inside the double range for any map weight (k <= 1023); a term that still underflows is below e^-700 of the sum. */
  const double lnc = phd_cphd_log_nc(N1);                                                /* log of the cardinality scale */
  const double lsd = lnc - (double)lW;                                           /* log s */
  const double lqs = (double)lq + lsd;
  std::vector<double> cc(N1), dd(N1), aa(M + 1);
  for (int n = 0; n <= N; ++n) {
    cc[n] = expd(((double)pm[n] + (double)lf[n]) - (double)n * lnc);             /* p-(n) n! / (s <1,w>)^n */
    dd[n] = (n == 0) ? 1.0 : expd((double)n * lqs - (double)lf[n]);                      /* (q s)^n / n! */
  }
  /* Psi0(n) (:1686-1703) and the updated cardinality (:1767-1768) */
  std::vector<float> psi0(N1), v(N1), t(std::max(M + 1, N1));
  float amax = PHD_LOG0;
  for (int j = 0; j <= M; ++j) {
    t[j] = lclamp(cK[M - j] + le[j]);
    const float u = (float)((double)t[j] + (double)j * lsd);
    amax = (u > amax) ? u : amax;
  }
  for (int j = 0; j <= M; ++j) aa[j] = expd(((double)t[j] + (double)j * lsd) - (double)amax);
  for (int n = 0; n <= N; ++n) {
    int stop = std::min(n, M);
    double sum = 0.0;
    for (int j = 0; j <= stop; ++j) sum = fma(aa[j], dd[n - j], sum);
    psi0[n] = lclamp((float)((((double)logd(sum) + (double)amax) + (double)lf[n]) - (double)n * lnc));
    v[n] = lclamp(psi0[n] + pm[n]);
  }
  for (int n = 0; n <= N; ++n) t[n] = v[n];
  const float ip0 = lse_warp(t.data(), N1);                                             /* :1717-1722 */
  for (int n = 0; n <= N; ++n) card_out[n] = lclamp((pm[n] + psi0[n]) - ip0);
  /* A1[j] = log sum_{n>j} p(n) P(n,j+1) <q_D,w>^(n-j-1) / <1,w>^n = log(s^(j+1) sum_n c[n] d[n-j-1]):
   * four strided partial sums (n = j+1+p, step 4), combined (p0+p1)+(p2+p3) */
  std::vector<float> A1(M + 1);
  for (int j = 0; j <= M; ++j) {
    double part[4] = {0.0, 0.0, 0.0, 0.0};
    for (int n = j + 1; n <= N; ++n) part[(n - j - 1) & 3] = fma(cc[n], dd[n - j - 1], part[(n - j - 1) & 3]);
    const double sum = (part[0] + part[1]) + (part[2] + part[3]);
    A1[j] = lclamp((float)((double)logd(sum) + (double)(j + 1) * lsd));
  }
  for (int j = 0; j <= M; ++j) t[j] = lclamp((cK[M - j] + le[j]) + A1[j]);
  const float ip1 = lse_warp(t.data(), M + 1);                                          /* <Psi1, p>, :1706-1735 */
  *ND = ip1 - ip0;
  /* <Psi1d_m, p> (:1738-1764) = sum_j g[j] e_j(roots without m), g[j] = exp(cK[M-1-j] + j lmax + A1[j]): a linear
   * functional of the leave-one-out elementary symmetric functions.  The reference recomputes the whole O(M^2) Vieta
   * recursion for every m (:1577-1616, O(M^3) in all); canonically the leave-one-out coefficients come from the full ones
   * by deflation, e'_k = e_k - x_m e'_(k-1), in O(M) per m: forward up to the crossover ks = first k with
   * e_(k+1) <= x_m e_k (the peak of e_k x_m^-k: ESF sequences are log-concave), backward from the top beyond it
   * (Peters-Wilkinson composite deflation: each direction is used where it damps rounding errors; ~1e-15 relative for
   * coefficients above 1e-100, tests/test_cphd_kat.py), and the products with g[j] are accumulated on the fly. */
  {
    float gmax = PHD_LOG0;
    for (int j = 0; j < M; ++j) {
      const float u = lclamp((cK[M - 1 - j] + le[j]) + A1[j]);
      gmax = (u > gmax) ? u : gmax;
    }
    std::vector<double> g(std::max(M, 1));
    for (int j = 0; j < M; ++j)
      g[j] = expd((((double)cK[M - 1 - j] + (double)j * (double)lmax) + (double)A1[j]) - (double)gmax);
    for (int m = 0; m < M; ++m) {
      const double xm = x[m];
      int ks = M;
      if (xm > 0.0)
        for (int k = 0; k < M; ++k)
          if (e[k + 1] <= xm * e[k]) { ks = k; break; }
      double accf = 0.0, accb = 0.0, f = 1.0;
      for (int k = 0; k < ks; ++k) {                      /* forward: e'_0 = 1, e'_k = e_k - x_m e'_(k-1) */
        if (k > 0) f = fma(-xm, f, e[k]);
        accf = fma(g[k], f, accf);
      }
      if (ks < M) {                                       /* backward: e'_(M-1) = e_M / x_m, e'_(k-1) = (e_k - e'_k) / x_m */
        const double r = 1.0 / xm;
        double b = e[M] * r;
        for (int k = M - 1; k >= ks; --k) {
          if (k < M - 1) b = (e[k + 1] - b) * r;
          accb = fma(g[k], b, accb);
        }
      }
      const double acc = accf + accb;                     /* the two directions are independent chains in the kernel */
      const float ip1d = lclamp((float)((double)logd(acc) + (double)gmax));
      D[m] = ((ip1d - ip0) + lcr) - lcd;                                                  /* :1796-1798 */
    }
  }
  *inc = ip0;                                                                           /* .bak:2666 */
}

extern "C" void oracle_cphd_factors(const phdslam_config_t* cfg, const float* w, const float* pd, int C, const float* S,
                                    int M, const float* prior, int N1, float* D, float* ND, float* inc, float* card_out) {
  cphd_factors(*cfg, w, pd, C, S, M, prior, N1, D, ND, inc, card_out);
}

/* ------------------------------------------------------------------------- */
/* test hooks for the shared deterministic math                               */
/* ------------------------------------------------------------------------- */
extern "C" void oracle_detmath(int fn, const float* x, const float* y, float* out, float* out2, int n) {
  for (int i = 0; i < n; ++i) {
    switch (fn) {
      case 0: out[i] = phd_expf(x[i]); break;
      case 1: out[i] = phd_logf(x[i]); break;
      case 2: out[i] = phd_atan2f(y[i], x[i]); break;
      case 3: phd_sincosf(x[i], &out[i], &out2[i]); break;
      case 4: out[i] = phd_wrap_angle(x[i]); break;
      case 5: out[i] = phd_tanf(x[i]); break;
      case 6: out[i] = phd_safe_log(x[i]); break;
      default: out[i] = 0.0f;
    }
  }
}
extern "C" void oracle_philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out4) {
  phd_philox4_t r = phd_philox4x32_10(c0, c1, c2, c3, k0, k1);
  for (int i = 0; i < 4; ++i) out4[i] = r.v[i];
}
