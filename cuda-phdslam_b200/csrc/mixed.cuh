/*
 * mixed.cuh -- the dynamic (constant-velocity) features of the MIXED feature model (feature_model = 2; SURVEY
 * section 8(f) rank 4).  Reference: predictMapKernelMixed (src/phdfilter.cu:910-963), computeBirth / computePreUpdate on
 * Gaussian4D (:244-299, :397-521), phdUpdateKernelMixed (:2323-2635), phdUpdateMergeKernel<Gaussian4D> (:2707-2898) and the
 * MIXED_MODEL branches of phdUpdateSynth (:3412-3462, :3703-3726).  Oracle: oracle/phd_oracle.cpp (dyn_pre, dyn_terms,
 * merge_mixture4); the arithmetic is include/phd_mixed_math.h on both sides.
 *
 * A particle's static and dynamic features are coupled only through the per-measurement normaliser and the predicted
 * cardinality, so the static map keeps the tuned update / merge kernels of kernels.cuh (update_mixed_kernel) and the
 * dynamic map takes the three kernels below:
 *   dyn_predict_kernel  one thread per dynamic component, in place;
 *   dyn_pre_kernel      BEFORE the static update: sum over the in-range dynamic components of exp(partial log-weight) per
 *                       measurement, and sum of pd * w  -> mix_dsum, mix_nhat;
 *   dyn_update_kernel   AFTER it (it needs the normalisers the static kernel wrote to mix_L): the update terms in the
 *                       reference's order, prune, greedy 4-D Mahalanobis merge, the new dynamic map.
 * Dynamic maps are small (moving objects in the field of view: tens of components), so these kernels are one 64-thread
 * block per particle; the merge rounds of dyn_update_kernel run on one warp (DESIGN.md section 11 has the measurements).
 * The layout is plane-SoA like the static map: [particle][21][Dmax].  Resampling: dyn_gather_kernel (local copies and the
 * packing of what migrates to another rank).
 */
#ifndef PHD_MIXED_CUH
#define PHD_MIXED_CUH

#include "../../include/phd_mixed_math.h"

#define DYN_PLANES 21          /* Gaussian4D: cov[16], mean[4], weight */
#define DYN_THREADS 64
#define DYN_WARPS (DYN_THREADS / 32)

struct DynArgs {
  const float* dmap_in; const int* dcount_in;   /* [n][21][Dmax], [n] */
  float* dmap_out; int* dcount_out;
  const float* pose;                            /* [6][n] */
  const float* z;                               /* [3][PHD_MAX_MEAS] */
  int M, n, Dmax, Sd;
  float* dsum; float* nhat;                     /* [n][PHD_MAX_MEAS], [n] */
  const float* L;                               /* [n][PHD_MAX_MEAS] log normalisers of the static update kernel */
  float* dlogw;                                 /* [n] (Vo's weighting adds the dynamic sums) */
  phdslam_gaussian4d_t* cand;                   /* [n][Sd] prune survivors */
  Reductions* red;
  float dt, var_x, var_y, ps, beta, tau, cov_vx, cov_vy;
  DevCfg c;
};

__device__ __forceinline__ void dyn_load(const float* mp, int Dmax, int j, phdslam_gaussian4d_t* g) {
  float* f = reinterpret_cast<float*>(g);
#pragma unroll
  for (int k = 0; k < DYN_PLANES; ++k) f[k] = mp[(size_t)k * Dmax + j];
}
__device__ __forceinline__ void dyn_store(float* mp, int Dmax, int j, const phdslam_gaussian4d_t* g) {
  const float* f = reinterpret_cast<const float*>(g);
#pragma unroll
  for (int k = 0; k < DYN_PLANES; ++k) mp[(size_t)k * Dmax + j] = f[k];
}

/* Dynamic shared memory of dyn_pre_kernel / dyn_update_kernel: the stage buffers (pre-update constants, index list and one
 * exponential array per warp for Dmax components, three per-measurement arrays), or -- they are dead by then -- the merge
 * arrays of dyn_update_kernel (17 bytes per candidate) with at least the candidate capacity Sd. */
__host__ __device__ static inline size_t dyn_smem_bytes(int Dmax, int M, int Sd) {
  const size_t stage = (size_t)Dmax * (sizeof(phd_g4_pre_t) + sizeof(int) + DYN_WARPS * sizeof(float)) +
                       3 * (size_t)((M + 3) & ~3) * sizeof(float);
  const size_t merge = 17 * (size_t)Sd + 64;
  return ((stage > merge ? stage : merge) + 15) & ~(size_t)15;
}

/* predictMapKernelMixed: every component of every particle, in place */
__global__ void dyn_predict_kernel(float* __restrict__ dmap, const int* __restrict__ dcount, int n, int Dmax, float dt,
                                   float var_x, float var_y, float ps, float beta, float tau) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * Dmax) return;
  const int p = (int)(i / Dmax), j = (int)(i - (long long)p * Dmax);
  if (j >= dcount[p]) return;
  float* mp = dmap + (size_t)p * DYN_PLANES * Dmax;
  phdslam_gaussian4d_t g, q;
  dyn_load(mp, Dmax, j, &g);
  phd_g4_predict(&g, dt, var_x, var_y, ps, beta, tau, &q);
  dyn_store(mp, Dmax, j, &q);
}

/* Shared staging of both update kernels: the particle's in-range (class 1) dynamic components in map order and their
 * pre-update constants.  Returns their number. */
__device__ __forceinline__ int dyn_stage(const DynArgs& a, int p, int* s_idx, phd_g4_pre_t* s_pre, int* s_wcnt) {
  const DevCfg& c = a.c;
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int cnt = a.dcount_in[p];
  const float px = a.pose[0 * (size_t)a.n + p], py = a.pose[1 * (size_t)a.n + p], pth = a.pose[2 * (size_t)a.n + p];
  const float* mp = a.dmap_in + (size_t)p * DYN_PLANES * a.Dmax;
  int C = 0;
  for (int base = 0; base < cnt; base += DYN_THREADS) {
    const int i = base + tid;
    bool in = false;
    if (i < cnt) {   /* computeInRangeKernel on the position (:1328-1346); class 2 and 0 are dropped (:3713-3719) */
      const float dx = mp[16 * (size_t)a.Dmax + i] - px, dy = mp[17 * (size_t)a.Dmax + i] - py;
      const float r = sqrtf(dx * dx + dy * dy);
      const float ab = fabsf(phd_wrap_angle(phd_atan2f(dy, dx) - pth));
      in = (r >= c.min_range && r <= c.max_range && ab <= c.max_bearing);
    }
    const unsigned bal = __ballot_sync(FULL_MASK, in);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < DYN_WARPS; ++w) {
      if (w < warp) woff += s_wcnt[w];
      tot += s_wcnt[w];
    }
    if (in) s_idx[C + woff + __popc(bal & ((1u << lane) - 1u))] = i;
    C += tot;
    __syncthreads();
  }
  for (int j = tid; j < C; j += DYN_THREADS) {
    phdslam_gaussian4d_t g;
    dyn_load(mp, a.Dmax, s_idx[j], &g);
    phd_g4_preupdate(px, py, pth, &g, c.max_range, c.max_bearing, c.pd, c.var_r, c.var_b, &s_pre[j]);
  }
  __syncthreads();
  return C;
}

__global__ void __launch_bounds__(DYN_THREADS) dyn_pre_kernel(DynArgs a) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  phd_g4_pre_t* s_pre = reinterpret_cast<phd_g4_pre_t*>(dyn_smem);
  int* s_idx = reinterpret_cast<int*>(s_pre + a.Dmax);
  float* s_ev = reinterpret_cast<float*>(s_idx + a.Dmax);            /* [DYN_WARPS][Dmax] */
  __shared__ int s_wcnt[DYN_WARPS];
  const DevCfg& c = a.c;
  const int p = blockIdx.x, lane = lane_id(), warp = warp_id();
  const float* mp = a.dmap_in + (size_t)p * DYN_PLANES * a.Dmax;
  const int C = dyn_stage(a, p, s_idx, s_pre, s_wcnt);
  float* ev = s_ev + (size_t)warp * a.Dmax;
  if (warp == 0) {   /* share of the predicted cardinality: sum pd * w (:2424-2446) */
    for (int j = lane; j < C; j += 32) ev[j] = s_pre[j].pd * mp[20 * (size_t)a.Dmax + s_idx[j]];
    __syncwarp();
    const float s = warp_sum_array(ev, C);
    if (lane == 0) a.nhat[p] = s;
    __syncwarp();
  }
  for (int m = warp; m < a.M; m += DYN_WARPS) {
    const float zr = a.z[m], zb = a.z[PHD_MAX_MEAS + m];
    const int dead = c.labeled && (a.z[2 * PHD_MAX_MEAS + m] != 1.0f);   /* not DYNAMIC_MEASUREMENT (:504) */
    for (int j = lane; j < C; j += 32) ev[j] = phd_expf(phd_g4_detect(&s_pre[j], nullptr, zr, zb, dead, nullptr));
    __syncwarp();
    const float s = warp_sum_array(ev, C);
    if (lane == 0) a.dsum[(size_t)p * PHD_MAX_MEAS + m] = s;
    __syncwarp();
  }
}

/* tie rule of the reference's arg-max reduction (oracle: merge_tie_key) */
__device__ __forceinline__ unsigned dyn_tie_key(int i) { return (__brev((unsigned)i & 255u) >> 24 << 24) | ((unsigned)i >> 8); }

/* The greedy rounds of the dynamic map's merge, run by ONE warp.  Every sum runs over the members in ascending candidate
 * order, as phd_g4_moment_match (the oracle) does; the 21 sums of a cluster are independent, so 5 lanes take the weight and
 * the mean sums and 16 lanes the covariance sums.  ALL_SH: every candidate is staged in shared memory (the usual case).
 * Returns the number of clusters. */
template <bool ALL_SH>
__device__ __forceinline__ int dyn_merge_rounds(const DevCfg& c, int n, int n_sh, const phdslam_gaussian4d_t* s_cand,
                                                const phdslam_gaussian4d_t* cand, unsigned char* s_merged, int* s_members,
                                                const float* s_cw, const float* s_trp, const float* s_trv, float* mo, int Dmax) {
  __shared__ phdslam_gaussian4d_t s_mg;
  const int lane = lane_id();
  const float gk = 0.515625f * c.min_sep;
  auto at = [&](int i) -> const phdslam_gaussian4d_t* { return (ALL_SH || i < n_sh) ? (s_cand + i) : (cand + i); };
  int nout = 0;
  for (;;) {
    /* arg-max of the unmerged weights, ties as the reference's reduction tree breaks them: weights are non-negative, so
     * their bit patterns order like the values, and the complement of the tie key makes "smallest key" a maximum too */
    unsigned hi = 0, lo = 0;
    for (int i = lane; i < n; i += 32) {
      if (s_merged[i]) continue;
      const unsigned h = __float_as_uint(s_cw[i]), l = ~dyn_tie_key(i);
      if (h > hi || (h == hi && l > lo)) { hi = h; lo = l; }
    }
    const unsigned mh = __reduce_max_sync(FULL_MASK, hi);
    const unsigned ml = __reduce_max_sync(FULL_MASK, (hi == mh) ? lo : 0u);
    if (ml == 0u) break;                                  /* nothing left (a live candidate has a non-zero complement) */
    const unsigned tk = ~ml;
    const int best = (int)(((tk & 0xffffffu) << 8) | (__brev(tk >> 24) >> 24));
    const phdslam_gaussian4d_t* sp = at(best);
    const float s0 = sp->mean[0], s1 = sp->mean[1], s2 = sp->mean[2], s3 = sp->mean[3];
    const float tp = s_trp[best], tv = s_trv[best];
    /* members: unmerged candidates closer than minSeparation (:2797-2812), listed in ascending order */
    int nm = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      bool mem = false, need = false;
      if (i < n && !s_merged[i]) {
        if (i == best) {
          mem = true;                       /* the seed opens its own cluster (distance 0) */
        } else if (c.distance_metric == 0) {
          /* canonical gates (oracle: merge_mixture4): outside them the Mahalanobis distance exceeds 1.03 minSeparation */
          const phdslam_gaussian4d_t* gp = at(i);
          const float g0 = s0 - gp->mean[0], g1 = s1 - gp->mean[1], g2 = s2 - gp->mean[2], g3 = s3 - gp->mean[3];
          need = (g0 * g0 + g1 * g1 <= gk * (tp + s_trp[i])) && (g2 * g2 + g3 * g3 <= gk * (tv + s_trv[i]));
        } else {
          mem = (0.0f < c.min_sep);         /* no 4-D Hellinger distance in the reference: 0 */
        }
      }
      if (__any_sync(FULL_MASK, need)) {    /* most clusters are singletons: no lane needs the 4 x 4 factorisation */
        if (need) {
          const phdslam_gaussian4d_t sd = *sp, g = *at(i);
          mem = (phd_g4_mahal(&sd, &g) < c.min_sep);
        }
      }
      const unsigned bal = __ballot_sync(FULL_MASK, mem);
      if (mem) s_members[nm + __popc(bal & ((1u << lane) - 1u))] = i;
      nm += __popc(bal);
    }
    __syncwarp();
    if (lane < 5) {     /* weight sum and the four weighted mean sums (:2808-2829) */
      float acc = 0.0f;
      for (int q = 0; q < nm; ++q) {
        const phdslam_gaussian4d_t* g = at(s_members[q]);
        acc = (lane == 0) ? acc + g->weight : acc + g->weight * g->mean[lane - 1];
      }
      if (lane == 0) s_mg.weight = acc;
      else s_mg.mean[lane - 1] = acc;
    }
    __syncwarp();
    const float wsum = s_mg.weight;
    if (wsum == 0.0f) break;                                         /* :2821-2822 */
    const float rw = 1.0f / wsum;
    if (lane < 16) {    /* covariance sums (:2836-2881), element j * 4 + k */
      const int j4 = lane >> 2, k4 = lane & 3;
      const float mj = s_mg.mean[j4] * rw, mk = s_mg.mean[k4] * rw;
      float acc = 0.0f;
      for (int q = 0; q < nm; ++q) {
        const phdslam_gaussian4d_t* g = at(s_members[q]);
        const float dj = mj - g->mean[j4], dk = mk - g->mean[k4];
        acc = acc + g->weight * (g->cov[lane] + dj * dk);
      }
      s_mg.cov[lane] = acc * rw;
    }
    __syncwarp();
    if (lane < DYN_PLANES) {
      float v;
      if (lane < 16) {                     /* force_symmetric_covariance: off-diagonal pairs averaged */
        const int r4 = lane & 3, c4 = lane >> 2;
        v = (r4 == c4) ? s_mg.cov[lane] : (s_mg.cov[lane] + s_mg.cov[c4 + 4 * r4]) / 2.0f;
      } else if (lane < 20) {
        v = s_mg.mean[lane - 16] * rw;
      } else {
        v = wsum;
      }
      if (nout < Dmax) mo[(size_t)lane * Dmax + nout] = v;
    }
    for (int q = lane; q < nm; q += 32) s_merged[s_members[q]] = 1;
    __syncwarp();
    nout++;
  }
  return nout;
}

__global__ void __launch_bounds__(DYN_THREADS, 16) dyn_update_kernel(DynArgs a) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  phd_g4_pre_t* s_pre = reinterpret_cast<phd_g4_pre_t*>(dyn_smem);
  int* s_idx = reinterpret_cast<int*>(s_pre + a.Dmax);
  float* s_ev = reinterpret_cast<float*>(s_idx + a.Dmax);            /* [DYN_WARPS][Dmax] */
  const int Mpad = (a.M + 3) & ~3;
  float* s_L = s_ev + (size_t)DYN_WARPS * a.Dmax;                     /* [Mpad] log normalisers */
  float* s_wb = s_L + Mpad;                                           /* birth weights */
  float* s_dm = s_wb + Mpad;                                          /* Vo: detection + birth weight sum per measurement */
  __shared__ int s_wcnt[DYN_WARPS];
  __shared__ int s_ncand;
  const DevCfg& c = a.c;
  const int p = blockIdx.x, tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int M = a.M, Dmax = a.Dmax, Sd = a.Sd;
  const float px = a.pose[0 * (size_t)a.n + p], py = a.pose[1 * (size_t)a.n + p], pth = a.pose[2 * (size_t)a.n + p];
  const float* mp = a.dmap_in + (size_t)p * DYN_PLANES * Dmax;
  phdslam_gaussian4d_t* cand = a.cand + (size_t)p * Sd;
  const int C = dyn_stage(a, p, s_idx, s_pre, s_wcnt);
  for (int m = tid; m < M; m += DYN_THREADS) {
    const float L = a.L[(size_t)p * PHD_MAX_MEAS + m];
    const bool dead = c.labeled && (a.z[2 * PHD_MAX_MEAS + m] != 1.0f);
    s_L[m] = L;
    s_wb[m] = phd_expf((dead ? PHD_LOG0 : c.log_birth_weight) - L);   /* :2266-2271, :2503-2528 */
  }
  if (tid == 0) s_ncand = 0;
  __syncthreads();

  /* Vo's empty-map weighting sums every dynamic update weight (:2546-2563); canonical order as in the oracle's dyn_terms */
  if (c.particle_weighting == 1) {
    float* ev = s_ev + (size_t)warp * Dmax;
    for (int m = warp; m < M; m += DYN_WARPS) {
      const float zr = a.z[m], zb = a.z[PHD_MAX_MEAS + m];
      const int dead = c.labeled && (a.z[2 * PHD_MAX_MEAS + m] != 1.0f);
      for (int j = lane; j < C; j += 32) ev[j] = phd_expf(phd_g4_detect(&s_pre[j], nullptr, zr, zb, dead, nullptr) - s_L[m]);
      __syncwarp();
      const float s = warp_sum_array(ev, C);
      if (lane == 0) s_dm[m] = s + s_wb[m];
      __syncwarp();
    }
    __syncthreads();
    if (warp == 0) {
      float* e0 = s_ev;
      for (int j = lane; j < C; j += 32) e0[j] = mp[20 * (size_t)Dmax + s_idx[j]] * (1.0f - s_pre[j].pd);
      __syncwarp();
      float upd = warp_sum_array(e0, C);
      __syncwarp();
      for (int j = lane; j < C; j += 32) e0[j] = mp[20 * (size_t)Dmax + s_idx[j]];
      __syncwarp();
      const float prior = warp_sum_array(e0, C);
      if (lane == 0) {
        for (int m = 0; m < M; ++m) upd = upd + s_dm[m];
        a.dlogw[p] = a.dlogw[p] + ((upd - prior) - (float)M * c.birth_weight);
      }
    }
    __syncthreads();
  }

  /* ---- update terms in the reference's order [non-detect C | detect m-major | birth M]; the prune survivors
   * (!(w < minFeatureWeight), :2612-2633) keep that order (pruneMap :3120-3174) ---- */
  const int T = C * (M + 1) + M;
  for (int t0 = 0; t0 < T; t0 += DYN_THREADS) {
    const int t = t0 + tid;
    float w = 0.0f;
    int kind = -1, j = 0, m = 0;
    if (t < T) {
      if (t < C) {
        kind = 0; j = t;
        w = mp[20 * (size_t)Dmax + s_idx[j]] * (1.0f - s_pre[j].pd);
      } else if (t < C + M * C) {
        kind = 1; m = (t - C) / C; j = (t - C) - m * C;
        const int dead = c.labeled && (a.z[2 * PHD_MAX_MEAS + m] != 1.0f);
        w = phd_expf(phd_g4_detect(&s_pre[j], nullptr, a.z[m], a.z[PHD_MAX_MEAS + m], dead, nullptr) - s_L[m]);
      } else {
        kind = 2; m = t - C - M * C;
        w = s_wb[m];
      }
    }
    const bool keep = (kind >= 0) && !(w < c.min_w);
    const unsigned bal = __ballot_sync(FULL_MASK, keep);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < DYN_WARPS; ++q) {
      if (q < warp) woff += s_wcnt[q];
      tot += s_wcnt[q];
    }
    const int slot = s_ncand + woff + __popc(bal & ((1u << lane) - 1u));
    if (keep && slot < Sd) {
      phdslam_gaussian4d_t g;
      if (kind == 2) {
        phd_g4_birth(px, py, pth, a.z[m], a.z[PHD_MAX_MEAS + m], c.bvar_r, c.bvar_b, a.cov_vx, a.cov_vy, &g);
      } else {
        dyn_load(mp, Dmax, s_idx[j], &g);
        if (kind == 1) {
          float mean[4];
          phd_g4_detect(&s_pre[j], &g, a.z[m], a.z[PHD_MAX_MEAS + m], 0, mean);
#pragma unroll
          for (int k = 0; k < 4; ++k) g.mean[k] = mean[k];
#pragma unroll
          for (int k = 0; k < 16; ++k) g.cov[k] = s_pre[j].cov[k];
        }
      }
      g.weight = w;
      cand[slot] = g;
    }
    __syncthreads();
    if (tid == 0) s_ncand += tot;
    __syncthreads();
  }
  int n = s_ncand;
  if (n > Sd) {
    if (tid == 0) atomicOr(&a.red->err_flag, 8);
    n = Sd;
  }
  __threadfence_block();
  __syncthreads();

  /* ---- greedy merge (phdUpdateMergeKernel<Gaussian4D>).  The stage buffers are dead: their shared memory now holds, sized
   * by the candidate count, the merged flags, the member list of the current cluster, the candidate weights and traces,
   * and as many candidates as fit behind them (an L2 round trip per access otherwise: the rounds are latency bound);
   * the others stay where the update wrote them. ---- */
  unsigned char* s_merged = reinterpret_cast<unsigned char*>(dyn_smem);                 /* n bytes */
  int* s_members = reinterpret_cast<int*>(dyn_smem + ((n + 15) & ~15));                 /* n ints */
  float* s_cw = reinterpret_cast<float*>(s_members + n);                                /* n floats */
  float* s_trp = s_cw + n;                                                              /* trace of the position block */
  float* s_trv = s_trp + n;                                                             /* trace of the velocity block */
  phdslam_gaussian4d_t* s_cand = reinterpret_cast<phdslam_gaussian4d_t*>(s_trv + n);
  const int n_sh = min(n, (int)((dyn_smem_bytes(Dmax, M, Sd) - (size_t)(reinterpret_cast<unsigned char*>(s_cand) - dyn_smem)) /
                                sizeof(phdslam_gaussian4d_t)));
  for (int i = tid; i < n; i += DYN_THREADS) {
    s_merged[i] = 0;
    s_cw[i] = cand[i].weight;
    s_trp[i] = cand[i].cov[0] + cand[i].cov[5];
    s_trv[i] = cand[i].cov[10] + cand[i].cov[15];
  }
  for (int i = tid; i < n_sh * DYN_PLANES; i += DYN_THREADS)
    reinterpret_cast<float*>(s_cand)[i] = reinterpret_cast<const float*>(cand)[i];
  __syncthreads();
  /* The rounds are short and strictly sequential: ONE warp runs them (warp barriers only, no block barrier per round; the
   * other warp leaves its issue slots to the other blocks of the SM). */
  if (warp == 0) {
    float* mo = a.dmap_out + (size_t)p * DYN_PLANES * Dmax;
    int nout;
    if (n_sh == n)
      nout = dyn_merge_rounds<true>(c, n, n_sh, s_cand, cand, s_merged, s_members, s_cw, s_trp, s_trv, mo, Dmax);
    else
      nout = dyn_merge_rounds<false>(c, n, n_sh, s_cand, cand, s_merged, s_members, s_cw, s_trp, s_trv, mo, Dmax);
    if (lane == 0) {
      if (nout > Dmax) atomicOr(&a.red->err_flag, 8);
      a.dcount_out[p] = min(nout, Dmax);
    }
  }
}

/* Resampling: offspring j takes the dynamic map of ancestor anc[j] - anc_offset of THIS rank; offspring whose ancestor
 * lives elsewhere (index outside [0, n_src)) are left alone -- the exchange below fills them.  The same kernel packs the
 * maps an outgoing interval of offspring needs into the staging buffer that is sent to their owner. */
__global__ void dyn_gather_kernel(const int* __restrict__ anc, int n_off, int anc_offset, int n_src,
                                  const float* __restrict__ dmap_in, const int* __restrict__ dcount_in,
                                  float* __restrict__ dmap_out, int* __restrict__ dcount_out, int Dmax) {
  const int j = blockIdx.x;
  if (j >= n_off) return;
  const int a = anc[j] - anc_offset;
  if (a < 0 || a >= n_src) return;
  const int cnt = dcount_in[a];
  if (threadIdx.x == 0) dcount_out[j] = cnt;
  const float* src = dmap_in + (size_t)a * DYN_PLANES * Dmax;
  float* dst = dmap_out + (size_t)j * DYN_PLANES * Dmax;
  for (int i = threadIdx.x; i < DYN_PLANES * cnt; i += blockDim.x) {
    const int k = i / cnt, q = i - k * cnt;
    dst[(size_t)k * Dmax + q] = src[(size_t)k * Dmax + q];
  }
}

/* phdslam_particle_checksums: the dynamic map's share of a particle's checksum (size and live words, position-mixed like
 * the static words of particle_checksum_kernel), added to what that kernel wrote */
__global__ void dyn_checksum_add_kernel(const int* __restrict__ dcount, const float* __restrict__ dmap, int n, int Dmax,
                                        unsigned long long* __restrict__ out) {
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= n) return;
  const int lane = lane_id();
  const int cnt = dcount[p];
  unsigned long long acc = (lane == 0) ? checksum_mix((2ull << 32), (unsigned)cnt) : 0ull;
  const float* m = dmap + (size_t)p * DYN_PLANES * Dmax;
  for (int k = 0; k < DYN_PLANES; ++k)
    for (int i = lane; i < cnt; i += 32)
      acc += checksum_mix((2ull << 32) + 16ull + (unsigned long long)i * DYN_PLANES + k, __float_as_uint(m[(size_t)k * Dmax + i]));
  acc = warp_sum_u64(acc);
  if (lane == 0) out[p] += acc;
}

#endif
