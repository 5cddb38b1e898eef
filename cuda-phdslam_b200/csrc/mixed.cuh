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
 * Dynamic maps are small (moving objects in the field of view: tens of components), so these kernels are one 128-thread
 * block per particle and not tuned further; the layout is plane-SoA like the static map: [particle][21][Dmax].
 */
#ifndef PHD_MIXED_CUH
#define PHD_MIXED_CUH

#include "../../include/phd_mixed_math.h"

#define DYN_PLANES 21          /* Gaussian4D: cov[16], mean[4], weight */
#define DYN_THREADS 128
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

__host__ __device__ static inline size_t dyn_smem_bytes(int Dmax) {
  return (size_t)Dmax * (sizeof(phd_g4_pre_t) + sizeof(int) + DYN_WARPS * sizeof(float)) + 3 * PHD_MAX_MEAS * sizeof(float) +
         PHD_MAX_MEAS * sizeof(float);
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

__global__ void __launch_bounds__(DYN_THREADS) dyn_update_kernel(DynArgs a) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  phd_g4_pre_t* s_pre = reinterpret_cast<phd_g4_pre_t*>(dyn_smem);
  int* s_idx = reinterpret_cast<int*>(s_pre + a.Dmax);
  float* s_ev = reinterpret_cast<float*>(s_idx + a.Dmax);            /* [DYN_WARPS][Dmax] */
  float* s_L = s_ev + (size_t)DYN_WARPS * a.Dmax;                     /* [PHD_MAX_MEAS] */
  float* s_wb = s_L + PHD_MAX_MEAS;                                   /* birth weights */
  float* s_dm = s_wb + PHD_MAX_MEAS;                                  /* Vo: detection + birth weight sum per measurement */
  __shared__ int s_wcnt[DYN_WARPS];
  __shared__ int s_ncand, s_nout, s_best, s_done;
  __shared__ float s_bw[DYN_WARPS];
  __shared__ unsigned s_bk[DYN_WARPS];
  __shared__ int s_bi[DYN_WARPS];
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
  if (tid == 0) { s_ncand = 0; s_nout = 0; }
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

  /* ---- greedy merge (phdUpdateMergeKernel<Gaussian4D>); the stage buffers are dead: s_ev becomes the merged flags,
   * s_idx / s_pre the member list of the current cluster ---- */
  unsigned char* s_merged = reinterpret_cast<unsigned char*>(s_pre);     /* Sd bytes: fits (Sd <= Dmax * 128) */
  int* s_members = reinterpret_cast<int*>(s_merged + ((Sd + 15) & ~15)); /* Sd ints */
  for (int i = tid; i < n; i += DYN_THREADS) s_merged[i] = 0;
  __syncthreads();
  float* mo = a.dmap_out + (size_t)p * DYN_PLANES * Dmax;
  for (;;) {
    /* arg-max of the unmerged weights, ties as the reference's reduction tree breaks them */
    float bw = -1.0f;
    unsigned bk = 0xffffffffu;
    int bi = -1;
    for (int i = tid; i < n; i += DYN_THREADS) {
      if (s_merged[i]) continue;
      const float w = cand[i].weight;
      const unsigned k = dyn_tie_key(i);
      if (bi < 0 || bw < w || (bw == w && k < bk)) { bw = w; bk = k; bi = i; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const float ow = __shfl_xor_sync(FULL_MASK, bw, off);
      const unsigned ok = __shfl_xor_sync(FULL_MASK, bk, off);
      const int oi = __shfl_xor_sync(FULL_MASK, bi, off);
      if (oi >= 0 && (bi < 0 || bw < ow || (bw == ow && ok < bk))) { bw = ow; bk = ok; bi = oi; }
    }
    if (lane == 0) { s_bw[warp] = bw; s_bk[warp] = bk; s_bi[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
      for (int q = 1; q < DYN_WARPS; ++q)
        if (s_bi[q] >= 0 && (bi < 0 || bw < s_bw[q] || (bw == s_bw[q] && s_bk[q] < bk))) { bw = s_bw[q]; bk = s_bk[q]; bi = s_bi[q]; }
      s_best = bi;
      s_done = 0;
    }
    __syncthreads();
    const int best = s_best;
    if (best < 0) break;
    const phdslam_gaussian4d_t seed = cand[best];
    /* members: unmerged candidates closer than minSeparation (:2797-2812); 2 = member */
    for (int i = tid; i < n; i += DYN_THREADS) {
      if (s_merged[i]) continue;
      const phdslam_gaussian4d_t g = cand[i];
      const float dist = (c.distance_metric == 0) ? phd_g4_mahal(&seed, &g) : 0.0f;
      if (dist < c.min_sep) s_merged[i] = 2;
    }
    __syncthreads();
    if (tid == 0) {
      int nm = 0;
      for (int i = 0; i < n; ++i)
        if (s_merged[i] == 2) s_members[nm++] = i;
      phdslam_gaussian4d_t mg;
      if (!phd_g4_moment_match(cand, s_members, nm, &mg)) {
        s_done = 1;                                   /* :2821-2822 */
      } else {
        for (int q = 0; q < nm; ++q) s_merged[s_members[q]] = 1;
        if (s_nout < Dmax) dyn_store(mo, Dmax, s_nout, &mg);
        else atomicOr(&a.red->err_flag, 8);
        s_nout++;
      }
    }
    __syncthreads();
    if (s_done) break;
  }
  if (tid == 0) a.dcount_out[p] = min(s_nout, Dmax);
}

/* resampling: offspring j of this rank takes the dynamic map of its (local) ancestor */
__global__ void dyn_gather_kernel(const int* __restrict__ anc, int n_off, int n_src, const float* __restrict__ dmap_in,
                                  const int* __restrict__ dcount_in, float* __restrict__ dmap_out, int* __restrict__ dcount_out,
                                  int Dmax) {
  const int j = blockIdx.x;
  if (j >= n_off) return;
  const int a = anc[j];
  const int cnt = dcount_in[a];
  if (threadIdx.x == 0) dcount_out[j] = cnt;
  const float* src = dmap_in + (size_t)a * DYN_PLANES * Dmax;
  float* dst = dmap_out + (size_t)j * DYN_PLANES * Dmax;
  for (int i = threadIdx.x; i < DYN_PLANES * cnt; i += blockDim.x) {
    const int k = i / cnt, q = i - k * cnt;
    dst[(size_t)k * Dmax + q] = src[(size_t)k * Dmax + q];
  }
}

#endif
