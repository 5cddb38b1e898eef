/*
 * phd_mixed_math.h -- canonical fp32 arithmetic of the MIXED feature model (static + constant-velocity
 * features): the 4-D Gaussian components of the reference's `maps_dynamic`.
 *
 * Same contract as phd_detmath.h: only correctly rounded IEEE-754 operations and the transcendental
 * functions of phd_detmath.h, so the sm_100a kernels (-fmad=false) and the CPU oracle (-ffp-contract=off)
 * produce the same bits.  Each function states the reference code it restates; the expressions are written
 * as small matrix loops, not as the reference's Maple-expanded polynomials, and agree with those to rounding
 * (tests/test_mixed_ref_pin.py compares them with the reference's own kernels run through the emulator).
 *
 * Layouts (reference: Gaussian4D, src/slamtypes.h:135-139): state (x, y, vx, vy); cov[16] column-major,
 * cov[r + 4 c]; 84 bytes.
 */
#ifndef PHD_MIXED_MATH_H
#define PHD_MIXED_MATH_H

#include "phd_detmath.h"
#include "phdslam.h"

/* Survival / jump-Markov weight factor of a dynamic feature (predictMapKernelMixed, src/phdfilter.cu:919-952, the
 * MIXED_MODEL branch): p_jmm = 1 / (1 + exp(beta (tau - |v|))), predicted weight = p_jmm * ps * w.
 * Constant-velocity prediction (ConstantVelocityMotionModel::compute_prediction, src/device_math.cuh:612-660, scale 1):
 * mean' = F mean, P' = F P F^T + Q with F = [I, dt I; 0, I] and the white-acceleration Q of var_x, var_y. */
PHD_HD void phd_g4_predict(const phdslam_gaussian4d_t* p, float dt, float var_x, float var_y, float ps, float beta,
                           float tau, phdslam_gaussian4d_t* o) {
  const float vx = p->mean[2], vy = p->mean[3];
  const float vmag = sqrtf(vx * vx + vy * vy);
  const float p_jmm = 1.0f / (1.0f + phd_expf(beta * (tau - vmag)));
  o->mean[0] = p->mean[0] + dt * vx;
  o->mean[1] = p->mean[1] + dt * vy;
  o->mean[2] = vx;
  o->mean[3] = vy;
  float T[16]; /* T = F P : rows 0,1 gain dt * rows 2,3 */
  for (int c = 0; c < 4; ++c) {
    T[0 + 4 * c] = p->cov[0 + 4 * c] + dt * p->cov[2 + 4 * c];
    T[1 + 4 * c] = p->cov[1 + 4 * c] + dt * p->cov[3 + 4 * c];
    T[2 + 4 * c] = p->cov[2 + 4 * c];
    T[3 + 4 * c] = p->cov[3 + 4 * c];
  }
  for (int r = 0; r < 4; ++r) { /* P' = T F^T : columns 0,1 gain dt * columns 2,3 */
    o->cov[r + 0] = T[r + 0] + dt * T[r + 8];
    o->cov[r + 4] = T[r + 4] + dt * T[r + 12];
    o->cov[r + 8] = T[r + 8];
    o->cov[r + 12] = T[r + 12];
  }
  const float dt2 = dt * dt;
  const float q4 = (dt2 * dt2) / 4.0f, q3 = (dt2 * dt) / 2.0f;
  o->cov[0] = o->cov[0] + q4 * var_x;
  o->cov[2] = o->cov[2] + q3 * var_x;
  o->cov[8] = o->cov[8] + q3 * var_x;
  o->cov[10] = o->cov[10] + dt2 * var_x;
  o->cov[5] = o->cov[5] + q4 * var_y;
  o->cov[7] = o->cov[7] + q3 * var_y;
  o->cov[13] = o->cov[13] + q3 * var_y;
  o->cov[15] = o->cov[15] + dt2 * var_y;
  o->weight = (p_jmm * ps) * p->weight;
}

/* Per-component constants of the EKF update of a dynamic feature: computePreUpdate(.., Gaussian4D, ..)
 * (src/phdfilter.cu:395-521) up to its measurement loop. */
typedef struct phd_g4_pre {
  float pd, r, bearing;
  float S0, S12, S3; /* innovation information S[0], S[1] + S[2], S[3] */
  float nhl;         /* -(log det Sigma) / 2 */
  float base;        /* log pd + log w */
  float K[8];        /* Kalman gain, column-major 4 x 2 */
  float cov[16];     /* Joseph-form updated covariance (not re-symmetrised, as in the reference) */
} phd_g4_pre_t;

PHD_HD void phd_g4_preupdate(float px, float py, float pth, const phdslam_gaussian4d_t* f, float max_range,
                             float max_bearing, float pd_cfg, float var_r, float var_b, phd_g4_pre_t* o) {
  const float dx = f->mean[0] - px;
  const float dy = f->mean[1] - py;
  const float r2 = dx * dx + dy * dy;
  const float r = sqrtf(r2);
  const float bearing = phd_wrap_angle(phd_atan2f(dy, dx) - pth);
  float pd = 0.0f;
  if (r <= max_range && fabsf(bearing) <= max_bearing) pd = pd_cfg; /* :407-409 */
  /* H = d(range, bearing)/d(x, y), H[row][col]; the velocity columns are zero (:414-417) */
  const float H00 = dx / r, H01 = dy / r, H10 = -dy / r2, H11 = dx / r2;
  const float* P = f->cov;
  /* innovation covariance Sigma = H P H^T + R over the position block (:424-430) */
  const float a0 = P[0] * H00 + P[4] * H01, a1 = P[1] * H00 + P[5] * H01; /* P_pos H_0^T */
  const float b0 = P[0] * H10 + P[4] * H11, b1 = P[1] * H10 + P[5] * H11; /* P_pos H_1^T */
  float sg0 = H00 * a0 + H01 * a1 + var_r;
  float sg1 = H10 * a0 + H11 * a1;
  float sg2 = H00 * b0 + H01 * b1;
  float sg3 = H10 * b0 + H11 * b1 + var_b;
  sg1 = (sg1 + sg2) / 2.0f; /* :432-434 */
  sg2 = sg1;
  const float det = sg0 * sg3 - sg1 * sg2;
  const float S0 = sg3 / det, S1 = -sg1 / det, S2 = -sg2 / det, S3 = sg0 / det;
  /* K = P H^T S (:445-461): K[i][0] = P[i][0] (H00 S0 + H10 S1) + P[i][1] (H01 S0 + H11 S1), second column with S2, S3 */
  const float g00 = H00 * S0 + H10 * S1, g10 = H01 * S0 + H11 * S1;
  const float g01 = H00 * S2 + H10 * S3, g11 = H01 * S2 + H11 * S3;
  for (int i = 0; i < 4; ++i) {
    o->K[i] = P[i] * g00 + P[i + 4] * g10;
    o->K[i + 4] = P[i] * g01 + P[i + 4] * g11;
  }
  /* Joseph form P+ = A P A^T + K R K^T, A = I - K H (:464-480).  A differs from I in its first two columns only. */
  float A[16];
  for (int i = 0; i < 4; ++i) {
    A[i + 0] = ((i == 0) ? 1.0f : 0.0f) - o->K[i] * H00 - o->K[i + 4] * H10;
    A[i + 4] = ((i == 1) ? 1.0f : 0.0f) - o->K[i] * H01 - o->K[i + 4] * H11;
    A[i + 8] = (i == 2) ? 1.0f : 0.0f;
    A[i + 12] = (i == 3) ? 1.0f : 0.0f;
  }
  float T[16]; /* T = P A^T : T[a][j] = sum_b P[a][b] A[j][b] = P[a][0] A[j][0] + P[a][1] A[j][1] (+ P[a][j] for j >= 2) */
  for (int a = 0; a < 4; ++a)
    for (int j = 0; j < 4; ++j) {
      float t = P[a] * A[j] + P[a + 4] * A[j + 4];
      if (j >= 2) t = t + P[a + 4 * j];
      T[a + 4 * j] = t;
    }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float t = A[i] * T[0 + 4 * j] + A[i + 4] * T[1 + 4 * j];
      if (i >= 2) t = t + T[i + 4 * j];
      t = t + (o->K[i] * var_r) * o->K[j];
      t = t + (o->K[i + 4] * var_b) * o->K[j + 4];
      o->cov[i + 4 * j] = t;
    }
  o->pd = pd;
  o->r = r;
  o->bearing = bearing;
  o->S0 = S0;
  o->S12 = S1 + S2;
  o->S3 = S3;
  o->nhl = -(0.5f * phd_safe_log(det));
  o->base = phd_safe_log(pd) + phd_safe_log(f->weight);
}

/* One detection term: updated mean and partial log-weight (:483-515).  `dead`: labelled measurements and the label is
 * not DYNAMIC_MEASUREMENT. */
PHD_HD float phd_g4_detect(const phd_g4_pre_t* c, const phdslam_gaussian4d_t* f, float zr, float zb, int dead,
                           float* mean_out /* 4, may be null */) {
  const float i0 = zr - c->r;
  const float i1 = phd_wrap_angle(zb - c->bearing);
  if (mean_out)
    for (int i = 0; i < 4; ++i) mean_out[i] = fmaf(c->K[i + 4], i1, fmaf(c->K[i], i0, f->mean[i]));
  float dist = (i0 * i0) * c->S0;
  dist = fmaf(i0 * i1, c->S12, dist);
  dist = fmaf(i1 * i1, c->S3, dist);
  if (dead) return PHD_LOG0;
  return c->base + (fmaf(dist, -0.5f, -PHD_LOG_2PI_F) + c->nhl);
}

/* Birth term of a measurement in the dynamic map: computeBirth(.., Gaussian4D&) (src/phdfilter.cu:244-295).
 * Zero velocity, velocity variances from the configuration; the weight is set by the caller. */
PHD_HD void phd_g4_birth(float px, float py, float pth, float zr, float zb, float bvar_r, float bvar_b, float cov_vx,
                         float cov_vy, phdslam_gaussian4d_t* o) {
  float sn, cs;
  phd_sincosf(pth + zb, &sn, &cs);
  const float dx = zr * cs, dy = zr * sn;
  const float J0 = dx / zr, J1 = dy / zr, J2 = -dy, J3 = dx;
  for (int i = 0; i < 16; ++i) o->cov[i] = 0.0f;
  o->cov[0] = J0 * J0 * bvar_r + J2 * J2 * bvar_b;
  o->cov[1] = J0 * J1 * bvar_r + J2 * J3 * bvar_b;
  o->cov[4] = o->cov[1];
  o->cov[5] = J1 * J1 * bvar_r + J3 * J3 * bvar_b;
  o->cov[10] = cov_vx;
  o->cov[15] = cov_vy;
  o->mean[0] = px + dx;
  o->mean[1] = py + dy;
  o->mean[2] = 0.0f;
  o->mean[3] = 0.0f;
}

/* computeMahalDist(Gaussian4D, Gaussian4D) (src/device_math.cuh:346-363): d^T [(Pa + Pb)/2]^-1 d.  The reference inverts
 * with a cofactor expansion (invert_matrix4, :88-106); here the symmetric part of the averaged covariance is factored
 * L D L^T and the quadratic form is sum_j y_j^2 / D_j with L y = d -- equal up to rounding for the positive definite
 * matrices the filter produces. */
PHD_HD float phd_g4_mahal(const phdslam_gaussian4d_t* a, const phdslam_gaussian4d_t* b) {
  float s[4][4];
  for (int i = 0; i < 4; ++i) {
    s[i][i] = (a->cov[i + 4 * i] + b->cov[i + 4 * i]) * 0.5f;
    for (int j = 0; j < i; ++j)
      s[i][j] = ((a->cov[i + 4 * j] + b->cov[i + 4 * j]) + (a->cov[j + 4 * i] + b->cov[j + 4 * i])) * 0.25f;
  }
  float L[4][4], D[4], y[4];
  float dist = 0.0f;
  for (int j = 0; j < 4; ++j) {
    float dj = s[j][j];
    for (int k = 0; k < j; ++k) dj = dj - (L[j][k] * L[j][k]) * D[k];
    D[j] = dj;
    const float rd = 1.0f / dj;
    for (int i = j + 1; i < 4; ++i) {
      float t = s[i][j];
      for (int k = 0; k < j; ++k) t = t - (L[i][k] * L[j][k]) * D[k];
      L[i][j] = t * rd;
    }
    float yj = a->mean[j] - b->mean[j];
    for (int k = 0; k < j; ++k) yj = yj - L[j][k] * y[k];
    y[j] = yj;
    dist = dist + (yj * yj) * rd;
  }
  return dist;
}

/* Moment-matched merge of the cluster `members` (ascending candidate indices) of `cand`: the accumulation of
 * phdUpdateMergeKernel<Gaussian4D> (src/phdfilter.cu:2808-2886) in the canonical sequential order, followed by
 * force_symmetric_covariance (src/device_math.cuh:710-725).  Returns 0 when the weights sum to zero (:2821-2822). */
PHD_HD int phd_g4_moment_match(const phdslam_gaussian4d_t* cand, const int* members, int n_members,
                               phdslam_gaussian4d_t* o) {
  float wsum = 0.0f, m[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  for (int q = 0; q < n_members; ++q) {
    const phdslam_gaussian4d_t* g = cand + members[q];
    wsum = wsum + g->weight;
    for (int j = 0; j < 4; ++j) m[j] = m[j] + g->weight * g->mean[j];
  }
  if (wsum == 0.0f) return 0;
  const float rw = 1.0f / wsum;
  for (int j = 0; j < 4; ++j) o->mean[j] = m[j] * rw;
  float cv[16];
  for (int j = 0; j < 16; ++j) cv[j] = 0.0f;
  for (int q = 0; q < n_members; ++q) {
    const phdslam_gaussian4d_t* g = cand + members[q];
    float d[4];
    for (int j = 0; j < 4; ++j) d[j] = o->mean[j] - g->mean[j];
    for (int j = 0; j < 4; ++j)
      for (int k = 0; k < 4; ++k) cv[j * 4 + k] = cv[j * 4 + k] + g->weight * (g->cov[j * 4 + k] + d[j] * d[k]);
  }
  for (int j = 0; j < 16; ++j) o->cov[j] = cv[j] * rw;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < i; ++j) {
      const float s = (o->cov[i + 4 * j] + o->cov[j + 4 * i]) / 2.0f;
      o->cov[i + 4 * j] = s;
      o->cov[j + 4 * i] = s;
    }
  o->weight = wsum;
  return 1;
}

#endif
