/*
 * kernels.cuh -- hand-written sm_100a kernels of the RB-PHD-SLAM filter step.
 *
 * Compile with -fmad=false: the arithmetic below must round exactly as written (see
 * include/phd_detmath.h); fused multiply-adds appear only as explicit fmaf().
 *
 * Layouts (all fp32, resident in HBM between steps):
 *   pose   [6][N]            SoA planes px,py,ptheta,vx,vy,vtheta
 *   map    [N][6][Cmax]      per particle one contiguous block of 6 planes {w,mx,my,pxx,pxy,pyy};
 *                            a CTA reads its particle's block with fully coalesced 128-byte requests
 *   dense  per particle p, at float offset 7*toff[p], its update terms in the reference's features_update order
 *          [non-detect C | detect m-major M*C | birth M] (src/phdfilter.cu:2123-2124,2137-2166), stored in blocks
 *          of 64 terms; each block holds the 7 planes {c0,c1,c2,c3,mx,my,w} back to back (7 x 256 B), see
 *          dense_index()
 */
#ifndef PHD_KERNELS_CUH
#define PHD_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/phd_detmath.h"
#include "phdslam_internal.h"

#define FULL_MASK 0xffffffffu

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

/* Canonical reduction: the caller has accumulated lane-strided sequential partials; this is the
 * xor butterfly (16,8,4,2,1).  Every lane ends with the same bits (fp add is commutative). */
__device__ __forceinline__ float warp_butterfly_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(FULL_MASK, v, off);
  return v;
}

/* canonical warp_sum over a shared/global array (oracle: warp_sum) -- must be called by a full warp:
 * 64 strided partials (two per lane), pair add, xor butterfly */
__device__ __forceinline__ float warp_sum_array(const float* v, int n) {
  float a0 = 0.0f, a1 = 0.0f;
  for (int i = 2 * lane_id(); i < n; i += 64) {
    a0 = a0 + v[i];
    if (i + 1 < n) a1 = a1 + v[i + 1];
  }
  return warp_butterfly_sum(a0 + a1);
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(FULL_MASK, v, off);
  return v;
}
__device__ __forceinline__ long long warp_sum_i64(long long v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(FULL_MASK, v, off);
  return v;
}

/* order-preserving maps float -> int / uint (for atomicMax) */
__device__ __forceinline__ int float_to_ordered_int(float f) {
  int i = __float_as_int(f);
  return (i >= 0) ? i : (i ^ 0x7fffffff);
}
__device__ __host__ __forceinline__ float ordered_int_to_float(int i) {
  int j = (i >= 0) ? i : (i ^ 0x7fffffff);
#ifdef __CUDA_ARCH__
  return __int_as_float(j);
#else
  float f;
  memcpy(&f, &j, 4);
  return f;
#endif
}
__device__ __forceinline__ uint32_t float_to_ordered_uint(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
/* inverse; key 0 is never produced by a real float and stands for "no value" */
__device__ __host__ __forceinline__ float ordered_uint_to_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}

/* streaming stores for the write-once dense buffer: keep it out of L1 */
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

/* =========================================================================================== */
/* predict: phdPredictKernelAckerman / phdPredictKernel (reference src/phdfilter.cu:785-859)     */
/* one thread per particle, in place on the SoA pose planes                                     */
/* =========================================================================================== */
__global__ void predict_kernel(float* __restrict__ pose, int n, int offset, float v_enc, float alpha,
                               const double* __restrict__ draws, unsigned call, DevCfg c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float px = pose[0 * (size_t)n + i], py = pose[1 * (size_t)n + i], th = pose[2 * (size_t)n + i];
  float sn, cs;
  phd_sincosf(th, &sn, &cs);
  const float dt = c.dt_sub;
  if (c.motion_type == 1) {
    double d_alpha, d_enc;
    if (draws) {
      d_alpha = draws[2 * (size_t)i];
      d_enc = draws[2 * (size_t)i + 1];
    } else {
      phd_philox4_t r = phd_philox4x32_10((uint32_t)(offset + i), call, PHD_STREAM_PREDICT, 0u, c.seed_lo, c.seed_hi);
      float z0, z1;
      phd_box_muller(r.v[0], r.v[1], &z0, &z1);
      d_alpha = (double)z0;
      d_enc = (double)z1;
    }
    float n_alpha = (float)((double)c.std_alpha * d_alpha);
    float n_enc = (float)((double)c.std_enc * d_enc);
    float ve = v_enc + n_enc;
    float al = alpha + n_alpha;
    float ta = phd_tanf(al);
    float vc = ve / (1.0f - ta * c.h / c.l);
    float xc_dot = vc * cs;
    float yc_dot = vc * sn;
    float thc_dot = vc * ta / c.l;
    pose[0 * (size_t)n + i] = px + dt * (xc_dot - thc_dot * (c.a * sn + c.b * cs));
    pose[1 * (size_t)n + i] = py + dt * (yc_dot + thc_dot * (c.a * cs - c.b * sn));
    pose[2 * (size_t)n + i] = phd_wrap_angle(th + dt * thc_dot);
    pose[3 * (size_t)n + i] = 0.0f;
    pose[4 * (size_t)n + i] = 0.0f;
    pose[5 * (size_t)n + i] = 0.0f;
  } else {
    float vx = pose[3 * (size_t)n + i], vy = pose[4 * (size_t)n + i], vth = pose[5 * (size_t)n + i];
    double d0, d1, d2;
    if (draws) {
      d0 = draws[3 * (size_t)i];
      d1 = draws[3 * (size_t)i + 1];
      d2 = draws[3 * (size_t)i + 2];
    } else {
      phd_philox4_t r = phd_philox4x32_10((uint32_t)(offset + i), call, PHD_STREAM_PREDICT, 0u, c.seed_lo, c.seed_hi);
      float z0, z1, z2, z3;
      phd_box_muller(r.v[0], r.v[1], &z0, &z1);
      phd_box_muller(r.v[2], r.v[3], &z2, &z3);
      d0 = z0; d1 = z1; d2 = z2;
    }
    float nax = (float)((double)c.ax3 * d0);
    float nay = (float)((double)c.ay3 * d1);
    float nat = (float)((double)c.ayaw3 * d2);
    float hdt2 = dt * dt * 0.5f;
    pose[0 * (size_t)n + i] = px + dt * (vx * cs - vy * sn) + hdt2 * (nax * cs - nay * sn);
    pose[1 * (size_t)n + i] = py + dt * (vx * sn + vy * cs) + hdt2 * (nax * sn + nay * cs);
    pose[2 * (size_t)n + i] = phd_wrap_angle(th + dt * vth + hdt2 * nat);
    pose[3 * (size_t)n + i] = vx + dt * nax;
    pose[4 * (size_t)n + i] = vy + dt * nay;
    pose[5 * (size_t)n + i] = vth + dt * nat;
  }
}

/* =========================================================================================== */
/* in-range classification: computeInRangeKernel (reference src/phdfilter.cu:1279-1358)          */
/* one warp per particle; reads only the two mean planes; also emits the padded dense term       */
/* count of the particle so that the dense buffer can be packed exactly                          */
/* =========================================================================================== */
__global__ void classify_kernel(const float* __restrict__ map, const int* __restrict__ count,
                                const float* __restrict__ pose, int n, int M, uint8_t* __restrict__ cls,
                                int* __restrict__ n_in, unsigned long long* __restrict__ tpad, Reductions* red,
                                DevCfg c) {
  int p = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (p >= n) return;
  const int lane = lane_id();
  const int cnt = count[p];
  const float px = pose[0 * (size_t)n + p], py = pose[1 * (size_t)n + p], th = pose[2 * (size_t)n + p];
  const float* mx = map + (size_t)p * PHD_MAP_PLANES * c.Cmax + 1 * c.Cmax;
  const float* my = mx + c.Cmax;
  uint8_t* cl = cls + (size_t)p * c.Cmax;
  int nin = 0;
  for (int base = 0; base < cnt; base += 32) {
    int i = base + lane;
    int k = 0;
    if (i < cnt) {
      float dx = mx[i] - px;
      float dy = my[i] - py;
      float r2 = dx * dx + dy * dy;
      float r = sqrtf(r2);
      float bearing = phd_wrap_angle(phd_atan2f(dy, dx) - th);
      float ab = fabsf(bearing);
      if (r >= c.min_range && r <= c.max_range && ab <= c.max_bearing)
        k = 1;
      else if (r >= c.lo2 && r <= c.hi2 && ab <= c.hb2)
        k = 2;
      cl[i] = (uint8_t)k;
    }
    nin += __popc(__ballot_sync(FULL_MASK, k == 1));
  }
  if (lane == 0) {
    n_in[p] = nin;
    unsigned long long t = (unsigned long long)nin * (unsigned)(M + 1) + (unsigned)M;
    t = (t + 63ull) & ~63ull; /* dense terms are stored in blocks of 64 (see dense_index) */
    tpad[p] = t;
    atomicMax(&red->max_terms, (int)t);
  }
}

/* =========================================================================================== */
/* exclusive scan of uint64 (three small kernels; used for dense offsets and the resampling CDF) */
/* =========================================================================================== */
#define SCAN_THREADS 512
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ unsigned long long block_exclusive_scan_u64(unsigned long long v, unsigned long long* total) {
  /* v: this thread's value; returns exclusive prefix over the block in thread order */
  __shared__ unsigned long long s_w[SCAN_THREADS / 32];
  unsigned long long inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    unsigned long long t = __shfl_up_sync(FULL_MASK, inc, off);
    if (lane_id() >= off) inc += t;
  }
  if (lane_id() == 31) s_w[warp_id()] = inc;
  __syncthreads();
  if (warp_id() == 0) {
    unsigned long long w = (lane_id() < SCAN_THREADS / 32) ? s_w[lane_id()] : 0ull;
    unsigned long long winc = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned long long t = __shfl_up_sync(FULL_MASK, winc, off);
      if (lane_id() >= off) winc += t;
    }
    if (lane_id() < SCAN_THREADS / 32) s_w[lane_id()] = winc - w; /* exclusive warp offsets */
    if (lane_id() == SCAN_THREADS / 32 - 1) *total = winc;
  }
  __syncthreads();
  unsigned long long res = s_w[warp_id()] + inc - v;
  return res;
}

__global__ void scan_tile_sums_kernel(const unsigned long long* __restrict__ in, int n, unsigned long long* __restrict__ tile_sums) {
  __shared__ unsigned long long s_total;
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < (size_t)n) s += in[base + k];
  block_exclusive_scan_u64(s, &s_total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = s_total;
}
/* single block: exclusive scan of the tile sums in place; writes the grand total to *grand */
__global__ void scan_tile_offsets_kernel(unsigned long long* __restrict__ tile_sums, int n_tiles, unsigned long long* grand) {
  __shared__ unsigned long long s_total;
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_tiles; base += SCAN_THREADS) {
    int i = base + threadIdx.x;
    unsigned long long v = (i < n_tiles) ? tile_sums[i] : 0ull;
    unsigned long long ex = block_exclusive_scan_u64(v, &s_total);
    unsigned long long carry = s_carry;
    if (i < n_tiles) tile_sums[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand = s_carry;
}
/* out[i] = exclusive prefix; out[n] = total */
__global__ void scan_apply_kernel(const unsigned long long* __restrict__ in, int n, const unsigned long long* __restrict__ tile_offsets,
                                  unsigned long long* __restrict__ out) {
  __shared__ unsigned long long s_total;
  size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  unsigned long long v[SCAN_ITEMS];
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < (size_t)n) ? in[base + k] : 0ull;
    s += v[k];
  }
  unsigned long long ex = block_exclusive_scan_u64(s, &s_total) + tile_offsets[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < (size_t)n) out[base + k] = ex;
    ex += v[k];
    if (base + k == (size_t)n - 1) out[n] = ex;
  }
}

/* =========================================================================================== */
/* GM-PHD update: preUpdateSynthKernel + phdUpdateKernel + host birth loop                      */
/* (reference src/phdfilter.cu:1824-1925, 2083-2321, 3468-3507) fused into one pass.             */
/*                                                                                               */
/* One CTA (8 warps) per particle.  The particle's in-range components are compacted into shared */
/* memory and the per-component EKF constants are computed once.  Warp w then owns measurements  */
/* m = w, w+8, ...; each lane owns two ADJACENT components of every 64-component chunk and does  */
/* the per-(component, measurement) arithmetic on both at once with Blackwell's packed fp32x2    */
/* instructions (FADD2/FMUL2/FFMA2 via __fadd2_rn/__fmul2_rn/__ffma2_rn), 64-bit shared loads     */
/* and 64-bit streaming stores.  The per-measurement normaliser is the canonical reduction       */
/* (packed lane partials -> pair add -> xor butterfly) mirrored by the oracle's warp_sum.         */
/* Outputs: (DENSE) the reference's features_update array, plane-SoA, 128-byte coalesced          */
/* streaming stores; (always) the terms that survive the prune as 32-byte records for the merge.  */
/* Algorithmic traffic per particle: read 24*C + 32 B, write 28*(C*(M+1)+M) + 4 B.               */
/* =========================================================================================== */
#define UPD_THREADS 256
#define UPD_WARPS (UPD_THREADS / 32)
#define UPD_FLOATS_PER_COMP 25
#define UPD_NF 17            /* per-component constants kept for the (component, measurement) loop */
#define UPD_MAXCH 32         /* 64-component chunks per particle the fused mode's chunk skipping covers (Cmax <= 2048) */

/* field order of the shared-memory constant records: one record per PAIR of adjacent components, UPD_NF float2
 * each (136 bytes; lane-consecutive records are conflict-free for 64-bit shared loads) */
enum { F_NR = 0, F_NB, F_S0, F_S12, F_S3, F_NHL, F_BASE, F_K0, F_K1, F_K2, F_K3, F_MX, F_MY, F_CU0, F_CU1, F_CU2, F_CU3 };

/* The update kernel emits the prune survivors in SEGMENTS of consecutive terms: 32 non-detection terms (one warp
 * ballot), 64 detection terms (one warp iteration of one measurement), one birth term.  One atomicAdd reserves the slots
 * of a segment, and inside it the survivors take consecutive candidate slots in term order.  merge_fast_kernel rebuilds
 * (first slot, count) of every segment from the term indices the records carry and restores the reference's term order
 * (pruneMap keeps it, :3120-3174) with one prefix sum over the segments instead of a radix sort of the term indices. */
/* segment of term t of a particle with C in-range components and M measurements */
__device__ __forceinline__ int upd_segment_of(int t, int C, int M) {
  if (t < C) return t >> 5;
  const int u = t - C;
  const int nd = (C + 31) >> 5, nch = (C + 63) >> 6;
  if (u < M * C) {
    const int m = u / C;
    return nd + m * nch + ((u - m * C) >> 6);
  }
  return nd + M * nch + (u - M * C);
}
__host__ __device__ static inline int upd_nseg_nd(int C) { return (C + 31) >> 5; }
__host__ __device__ static inline int upd_nchunk(int C) { return (C + 63) >> 6; }
__host__ __device__ static inline int upd_nseg(int C, int M) { return upd_nseg_nd(C) + M * upd_nchunk(C) + M; }

struct UpdArgs {
  const float* map; const int* count; const uint8_t* cls; const float* pose;
  const float* z;                      /* [3][PHD_MAX_MEAS] range, bearing, label */
  int M, n, p0;
  const unsigned long long* toff;      /* exclusive scan of padded term counts (global over local particles) */
  unsigned long long tbase;            /* toff of the first particle of this batch */
  float* dense;
  const int* n_in;
  float* dlogw;
  float4* cand;                        /* [batch][Smax][2] surviving terms, unordered, term index in .w of the 2nd half */
  int* n_cand;                         /* [n] */
  int Smax;
  /* CPHD (filter_type 1) */
  const float* lfact;                  /* log-factorial table, PHD_LF_MAX entries */
  float* card;                         /* [n][n_card] log cardinality, updated in place */
  int write_card;                      /* 0 for the dense-terms query (state is not advanced) */
  /* mixed feature model (update_mixed_kernel): what the dynamic features of the particle add to the per-measurement
   * normaliser and to the predicted cardinality (dyn_pre_kernel), and the log normalisers handed back to them */
  const float* mix_dsum;               /* [n][PHD_MAX_MEAS] */
  const float* mix_nhat;               /* [n] */
  float* mix_L;                        /* [n][PHD_MAX_MEAS] */
  DevCfg c;
};

__host__ __device__ static inline size_t update_smem_bytes(int Cmax) {
  return ((size_t)UPD_FLOATS_PER_COMP * Cmax + 11 * PHD_MAX_MEAS + 64) * sizeof(float);
}

/* =========================================================================================== */
/* CPHD multi-object terms for the particle of this CTA (oracle: cphd_factors).                   */
/* Reference: cardinalityPredictKernel (src/phdfilter.cu:867-888) + binomial births               */
/* (src/phdfilter.cu.bak:779-791), computeEsfKernel (:1524-1618), computePsiKernel (:1626-1769),  */
/* cphdUpdateKernel (:1780-1822) -- all commented out at HEAD.                         */
/* 256 threads: thread-per-n for the cardinality-indexed terms, warp-per-job for the (M+1)        */
/* elementary-symmetric-function recursions (scaled, double) and the measurement-indexed          */
/* log-sum-exps.  Every reduction has the oracle's shape.                                         */
/* =========================================================================================== */
#define PHD_LF_MAX 1025            /* log-factorials 0..1024: max_cardinality <= 1023, M <= 256 */
#define CPHD_E_STRIDE 264          /* doubles per warp ESF array (M + 1 <= 257) */
#define CPHD_KREG 9                /* ceil(257 / 32) */

/* shared-memory layout of cphd_block, sized by the cardinality bins AND the measurement count of the step (the arrays
 * indexed by measurement / ESF degree take M + 1 entries, not 257): 47 KB instead of 62 KB per particle at N = 255,
 * M = 50, Cmax = 256, which is the difference between 3 and 4 resident CTAs per SM.
 * floats: lf[nlf] | pm[N1] | psi[N1] (prior pmf, then Psi0) | pb, A1, le, cK, llam, ip1d (Mp each) | 16 scalars
 * doubles: x, e (full ESF), a, g (Mp each) | c[N1] | d[N1] */
__host__ __device__ static inline int cphd_mp(int M) { return (M + 4) & ~3; }
__host__ __device__ static inline int cphd_nlf(int n_card, int M) { return ((n_card - 1 > M ? n_card - 1 : M) + 4) & ~3; }
__host__ __device__ static inline size_t cphd_float_count(int n_card, int M) {
  return (size_t)cphd_nlf(n_card, M) + 2 * (((size_t)n_card + 3) & ~(size_t)3) + 6 * (size_t)cphd_mp(M) + 16;
}
__host__ __device__ static inline size_t cphd_smem_bytes(int n_card, int M) {
  return cphd_float_count(n_card, M) * sizeof(float) + (4 * (size_t)cphd_mp(M) + 2 * (size_t)n_card) * sizeof(double);
}

__host__ __device__ static inline size_t cphd_chunkmax_bytes(int M, int Cmax) { return (size_t)M * ((Cmax + 63) >> 6) * sizeof(float); }

__device__ __forceinline__ float cphd_mulk(int k, float x) { return k == 0 ? 0.0f : (float)k * x; }
__device__ __forceinline__ float cphd_clamp(float t) { return (t < PHD_LOG0) ? PHD_LOG0 : t; }
/* exp of a double argument with float accuracy and double range (oracle: expd): t = k ln2 + r, exp(t) = 2^k expf(r) */
__device__ __forceinline__ double cphd_expd(double t) {
  if (t != t) return t;
  if (!(t > -700.0)) return 0.0;
  if (t > 700.0) t = 700.0;
  const double kd = rint(t * 1.4426950408889634);
  const double r = __fma_rn(-kd, 0.6931471805599453, t);
  const double m = (double)phd_expf((float)r);
  return m * __longlong_as_double((long long)((int)kd + 1023) << 52);      /* exact: the result is a normal number */
}
__device__ __forceinline__ float cphd_logd(double v) {
  if (!(v > 0.0)) return PHD_LOG0;
  int ex;
  double mant = frexp(v, &ex);
  return phd_logf((float)mant) + (float)ex * 0.693147182f;
}
/* log-sum-exp over i in [0, n) of f(i) with the canonical warp shape (oracle: lse_warp); full warp */
#define CPHD_LSE_R 5               /* terms of up to 64 * CPHD_LSE_R elements are evaluated once and kept in registers */
template <class F>
__device__ __forceinline__ float cphd_lse_warp(int n, F f) {
  if (n <= 0) return PHD_LOG0;
  const int lane = lane_id();
  if (n <= 64 * CPHD_LSE_R) {
    /* same maximum, same per-lane summation order (elements 2 lane + 64 r and the one after it) as the loops below */
    float v0[CPHD_LSE_R], v1[CPHD_LSE_R];
    float mxc = -INFINITY;
#pragma unroll
    for (int r = 0; r < CPHD_LSE_R; ++r) {
      v0[r] = -INFINITY;
      v1[r] = -INFINITY;
      if (64 * r < n) {
        const int i = 2 * lane + 64 * r;
        if (i < n) v0[r] = f(i);
        if (i + 1 < n) v1[r] = f(i + 1);
        mxc = fmaxf(mxc, fmaxf(v0[r], v1[r]));
      }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mxc = fmaxf(mxc, __shfl_xor_sync(FULL_MASK, mxc, off));
    float c0 = 0.0f, c1 = 0.0f;
#pragma unroll
    for (int r = 0; r < CPHD_LSE_R; ++r) {
      if (64 * r < n) {
        const int i = 2 * lane + 64 * r;
        if (i < n) c0 = c0 + phd_expf(v0[r] - mxc);
        if (i + 1 < n) c1 = c1 + phd_expf(v1[r] - mxc);
      }
    }
    return phd_safe_log(warp_butterfly_sum(c0 + c1)) + mxc;
  }
  float mx = -INFINITY;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, f(i));
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, off));
  float a0 = 0.0f, a1 = 0.0f;
  for (int i = 2 * lane; i < n; i += 64) {
    a0 = a0 + phd_expf(f(i) - mx);
    if (i + 1 < n) a1 = a1 + phd_expf(f(i + 1) - mx);
  }
  return phd_safe_log(warp_butterfly_sum(a0 + a1)) + mx;
}

/* Elementary symmetric functions of the roots x[0..M) without root `skip` (-1: none), one warp, coefficients in
 * REGISTERS: lane L holds the KR consecutive coefficients e[KR L .. KR L + KR - 1], KR = ceil((M + 1) / 32).  Folding
 * one root is e[k] += x * e[k-1] on the old values: inside a lane from the top coefficient down, and the lane's lowest
 * coefficient takes e[k-1] from the lane below with one shuffle.  Same operations as the oracle's esf(): every
 * coefficient is an independent chain of double fused multiply-adds, and the terms beyond the current degree are exact
 * zeros.  Result to the warp's shared array Es[0..M]. */
template <int KR>
__device__ __forceinline__ void cphd_esf_warp_t(const double* __restrict__ x, int M, int skip, double* __restrict__ Es, int lane) {
  double E[KR];
#pragma unroll
  for (int i = 0; i < KR; ++i) E[i] = 0.0;
  if (lane == 0) E[0] = 1.0;
  auto fold = [&](double xn) {
    double below = __shfl_up_sync(FULL_MASK, E[KR - 1], 1);
    if (lane == 0) below = 0.0;
#pragma unroll
    for (int i = KR - 1; i >= 1; --i) E[i] = __fma_rn(xn, E[i - 1], E[i]);
    E[0] = __fma_rn(xn, below, E[0]);
  };
  const int n_skip = (skip < 0) ? 0 : skip;
  for (int n = 0; n < n_skip; ++n) fold(x[n]);
  for (int n = skip + 1; n < M; ++n) fold(x[n]);
#pragma unroll
  for (int i = 0; i < KR; ++i) {
    const int k = KR * lane + i;
    if (k <= M) Es[k] = E[i];
  }
  __syncwarp();
}
__device__ __forceinline__ void cphd_esf_warp(const double* __restrict__ x, int M, int skip, double* __restrict__ Es, int lane) {
  switch ((M + 32) >> 5) {
    case 1: cphd_esf_warp_t<1>(x, M, skip, Es, lane); break;
    case 2: cphd_esf_warp_t<2>(x, M, skip, Es, lane); break;
    case 3: cphd_esf_warp_t<3>(x, M, skip, Es, lane); break;
    case 4: cphd_esf_warp_t<4>(x, M, skip, Es, lane); break;
    case 5: cphd_esf_warp_t<5>(x, M, skip, Es, lane); break;
    case 6: cphd_esf_warp_t<6>(x, M, skip, Es, lane); break;
    case 7: cphd_esf_warp_t<7>(x, M, skip, Es, lane); break;
    case 8: cphd_esf_warp_t<8>(x, M, skip, Es, lane); break;
    default: cphd_esf_warp_t<CPHD_KREG>(x, M, skip, Es, lane); break;
  }
}

/* in: s_w[C] weights, s_qd[C] = w*(1-pd), s_S[M] likelihood masses (overwritten by D[m], the log factor of
 * measurement m's detection and birth terms).  out: scal[0] = ND (log factor of the non-detection terms),
 * scal[1] = log<Psi0,p> (particle log-weight increment); card updated in place when write_card. */
__device__ void cphd_block(const DevCfg& c, int C, int M, const float* __restrict__ lfact, float* __restrict__ card,
                           int write_card, const float* s_w, const float* s_qd, float* s_S, float* scal_out,
                           unsigned char* smem_cphd) {
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int N1 = c.n_card, N = N1 - 1;
  const int Mp = cphd_mp(M), N1p = (N1 + 3) & ~3;
  float* s_lf = reinterpret_cast<float*>(smem_cphd);
  float* s_pm = s_lf + cphd_nlf(N1, M);
  float* s_psi = s_pm + N1p;
  float* s_pb = s_psi + N1p;
  float* s_A1 = s_pb + Mp;
  float* s_le = s_A1 + Mp;
  float* s_cK = s_le + Mp;
  float* s_llam = s_cK + Mp;
  float* s_ip1d = s_llam + Mp;
  float* s_sc = s_ip1d + Mp;             /* 16 scalars: 0 lq, 1 lW, 2 lmax, 3 ip0, 4 ip1, 5 amax, 6 gmax */
  double* s_x = reinterpret_cast<double*>(reinterpret_cast<float*>(smem_cphd) + cphd_float_count(N1, M));
  double* s_ef = s_x + Mp;                   /* elementary symmetric functions e_0..e_M of all the (scaled) roots */
  double* s_a = s_ef + Mp;                   /* a[j] of Psi0 */
  double* s_g = s_a + Mp;                    /* g[j] of <Psi1d_m, p> */
  double* s_c = s_g + Mp;                    /* c[n] = p-(n) n! / (s <1,w>)^n, n = 0..N */
  double* s_d = s_c + N1;                    /* d[k] = (q s)^k / k!, k = 0..N */

  const int nlf = max(N, M) + 1;
  const float wb = c.birth_weight;
  const float lwb = phd_safe_log(wb), l1wb = phd_safe_log(1.0f - wb);
  const float lcr = phd_safe_log(c.clutter_rate), lcd = phd_safe_log(c.clutter_density);
  const float larea = lcr - lcd;

  __shared__ int s_kmax;                 /* last birth cardinality whose probability is not exactly zero */
  if (tid == 0) s_kmax = 0;
  for (int k = tid; k < nlf; k += UPD_THREADS) s_lf[k] = lfact[k];
  for (int n = tid; n < N1; n += UPD_THREADS) s_psi[n] = phd_expf(card[n]); /* prior pmf (linear); becomes Psi0 below */
  for (int m = tid; m < M; m += UPD_THREADS) s_llam[m] = phd_safe_log(s_S[m] + wb) + larea;   /* :1539-1552 */
  if (warp == 0) {                                                            /* <q_D,w>, <1,w> (:1649-1683) */
    float q = warp_sum_array(s_qd, C);
    float Wsum = warp_sum_array(s_w, C) + (float)M * wb;
    if (lane == 0) {
      s_sc[0] = phd_safe_log(q);
      s_sc[1] = (Wsum > 0.0f) ? phd_logf(Wsum) : 0.0f;
    }
  }
  __syncthreads();
  /* birth cardinality Binomial(M, w_b) (.bak:779-791), clutter term k*log(cr) - cr (:735-737, :1691-1692) */
  for (int k = tid; k <= M; k += UPD_THREADS) {
    float t = s_lf[M] - s_lf[k];
    t = t - s_lf[M - k];
    t = t + cphd_mulk(k, lwb);
    t = t + cphd_mulk(M - k, l1wb);
    const float pbk = phd_expf(t);
    s_pb[k] = pbk;                                                          /* birth pmf (linear) */
    if (pbk != 0.0f) atomicMax(&s_kmax, k);
    s_cK[k] = cphd_mulk(k, lcr) - c.clutter_rate;
  }
  if (warp == 0) {
    float mx = PHD_LOG0;
    for (int m = lane; m < M; m += 32) mx = fmaxf(mx, s_llam[m]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, off));
    if (lane == 0) s_sc[2] = mx;
  }
  __syncthreads();
  const float lq = s_sc[0], lW = s_sc[1], lmax = s_sc[2];
  /* predicted cardinality (:880-887): the reference's plain sum of exp(birth(n-j) + prior(j)), evaluated as the
   * convolution of the two pmfs (N1 + M + 1 exponentials instead of N1 (M + 1)); the terms with n-j > M are exactly 0 */
  /* Binomial(M, w_b) underflows to an exact zero a few births beyond its mode (k > 12 at w_b = 1e-4, M = 50): those terms
   * add fma(0, p, sum) = sum, so leaving them out changes no bit of the sum (the oracle runs the full loop) */
  const int kmax = s_kmax;
  for (int n = tid; n < N1; n += UPD_THREADS) {
    float sum = 0.0f;
    for (int j = max(0, n - kmax); j <= n; ++j) sum = fmaf(s_pb[n - j], s_psi[j], sum);
    s_pm[n] = phd_safe_log(sum);
  }
  for (int m = tid; m < M; m += UPD_THREADS) s_x[m] = (double)phd_expf(s_llam[m] - lmax);
  __syncthreads();
  /* The two (N+1) x (M+1) tables of the update -- A1[j] here and Psi0(n) below -- are convolutions with
   * d[k] = (q s)^k / k!; they are evaluated in the linear domain in double, one fused multiply-add per term (the
   * reference: one exponential of a log-domain sum per term, :1686-1764).  s = n_c / <1,w> keeps every factor inside
   * the double range (oracle: cphd_factors). */
  const double lnc = phd_cphd_log_nc(N1);              /* log of the cardinality scale (phd_detmath.h) */
  const double lsd = lnc - (double)lW;
  const double lqs = (double)lq + lsd;
  for (int n = tid; n < N1; n += UPD_THREADS) {
    s_c[n] = cphd_expd(((double)s_pm[n] + (double)s_lf[n]) - (double)n * lnc);
    s_d[n] = (n == 0) ? 1.0 : cphd_expd((double)n * lqs - (double)s_lf[n]);
  }
  __syncthreads();
  if (warp == UPD_WARPS - 1) {
    /* elementary symmetric functions of the scaled roots (:1553-1576), by the last warp while the others do A1 */
    cphd_esf_warp(s_x, M, -1, s_ef, lane);
    for (int j = lane; j <= M; j += 32) s_le[j] = cphd_logd(s_ef[j]) + cphd_mulk(j, lmax);
  } else {
    /* A1[j] = log(s^(j+1) sum_{n>j} c[n] d[n-j-1]): four lanes per j, n = j+1+p step 4, combined (p0+p1)+(p2+p3) */
    for (int j0 = 0; j0 <= M; j0 += (UPD_THREADS - 32) / 4) {
      const int j = j0 + (tid >> 2), p = tid & 3;
      double part = 0.0;
      if (j <= M)
        for (int n = j + 1 + p; n <= N; n += 4) part = __fma_rn(s_c[n], s_d[n - j - 1], part);
      part = part + __shfl_xor_sync(FULL_MASK, part, 1);
      part = part + __shfl_xor_sync(FULL_MASK, part, 2);
      if (j <= M && p == 0) s_A1[j] = cphd_clamp((float)((double)cphd_logd(part) + (double)(j + 1) * lsd));
    }
  }
  __syncthreads();
  /* normalisers of the two linear functionals below */
  if (warp == 0) {
    float mx = PHD_LOG0;
    for (int j = lane; j <= M; j += 32)
      mx = fmaxf(mx, (float)((double)cphd_clamp(s_cK[M - j] + s_le[j]) + (double)j * lsd));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, off));
    if (lane == 0) s_sc[5] = mx;
  } else if (warp == 1) {
    float mx = PHD_LOG0;
    for (int j = lane; j < M; j += 32) mx = fmaxf(mx, cphd_clamp((s_cK[M - 1 - j] + s_le[j]) + s_A1[j]));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, off));
    if (lane == 0) s_sc[6] = mx;
  }
  __syncthreads();
  const float amax = s_sc[5], gmax = s_sc[6];
  for (int j = tid; j <= M; j += UPD_THREADS) {
    s_a[j] = cphd_expd(((double)cphd_clamp(s_cK[M - j] + s_le[j]) + (double)j * lsd) - (double)amax);
    if (j < M)
      s_g[j] = cphd_expd((((double)s_cK[M - 1 - j] + (double)j * (double)lmax) + (double)s_A1[j]) - (double)gmax);
  }
  __syncthreads();
  /* <Psi1d_m, p> (:1738-1764) = sum_j g[j] e_j(roots without m): the leave-one-out coefficients come from the full ones
   * by composite deflation, e'_k = e_k - x_m e'_(k-1) forward up to the crossover ks (first k with e_(k+1) <= x_m e_k),
   * backward from the top beyond it, O(M) per measurement instead of the reference's O(M^2) recomputation (:1577-1616);
   * one thread per measurement, the products with g[j] accumulated on the fly (oracle: cphd_factors, the same operations) */
  /* The deflation is one serial recursion per measurement and keeps two warps busy for ~M dependent double FMAs; everything
   * else that is left -- Psi0(n), <Psi0,p>, <Psi1,p> -- runs UNDER it on the other six warps (their own named barrier between
   * the Psi0 table and the log-sum-exp that reads it) instead of behind it. */
  if (warp < 2) {
    for (int m = tid; m < M; m += 64) {
      const double xm = s_x[m];
      const double r = (xm > 0.0) ? 1.0 / xm : 0.0;
      int ks = M;                                   /* first k with e_(k+1) <= x_m e_k; no early exit: the loads pipeline */
      if (xm > 0.0)
        for (int k = M - 1; k >= 0; --k)
          if (s_ef[k + 1] <= xm * s_ef[k]) ks = k;
      /* forward (k < ks) and backward (k >= ks) recursions are independent: one loop, two dependency chains */
      double accf = 0.0, accb = 0.0, f = 1.0, b = s_ef[M] * r;
      const int nb = M - ks, nmax = max(ks, nb);
      for (int i = 0; i < nmax; ++i) {
        if (i < ks) {
          if (i > 0) f = __fma_rn(-xm, f, s_ef[i]);
          accf = __fma_rn(s_g[i], f, accf);
        }
        if (i < nb) {
          const int k = M - 1 - i;
          if (i > 0) b = (s_ef[k + 1] - b) * r;
          accb = __fma_rn(s_g[k], b, accb);
        }
      }
      const double acc = accf + accb;
      s_ip1d[m] = cphd_clamp((float)((double)cphd_logd(acc) + (double)gmax));
    }
  } else {
    /* Psi0(n) (:1686-1703) = n! / (s <1,w>)^n * sum_j a[j] d[n-j], a[j] = exp(cK[M-j] + le[j] + j log s - amax) */
    for (int n = tid - 64; n < N1; n += UPD_THREADS - 64) {
      const int stop = min(n, M);
      double sum = 0.0;
      for (int j = 0; j <= stop; ++j) sum = __fma_rn(s_a[j], s_d[n - j], sum);
      s_psi[n] = cphd_clamp((float)((((double)cphd_logd(sum) + (double)amax) + (double)s_lf[n]) - (double)n * lnc));
    }
    asm volatile("bar.sync 1, %0;" ::"n"(UPD_THREADS - 64) : "memory");   /* the six Psi0 warps only */
    if (warp == 2) {        /* <Psi0, p> (:1717-1722) */
      float v = cphd_lse_warp(N1, [&](int n) { return cphd_clamp(s_psi[n] + s_pm[n]); });
      if (lane == 0) s_sc[3] = v;
    } else if (warp == 3) { /* <Psi1, p> (:1706-1735) */
      float v = cphd_lse_warp(M + 1, [&](int j) { return cphd_clamp((s_cK[M - j] + s_le[j]) + s_A1[j]); });
      if (lane == 0) s_sc[4] = v;
    }
  }
  __syncthreads();
  const float ip0 = s_sc[3], ip1 = s_sc[4];
  if (write_card)
    for (int n = tid; n < N1; n += UPD_THREADS) card[n] = cphd_clamp((s_pm[n] + s_psi[n]) - ip0);    /* :1767-1768 */
  for (int m = tid; m < M; m += UPD_THREADS) s_S[m] = ((s_ip1d[m] - ip0) + lcr) - lcd;                /* :1796-1798 */
  if (tid == 0) {
    scal_out[0] = ip1 - ip0;
    scal_out[1] = ip0;
  }
  __syncthreads();
}

/* log-factorial table, sequential as the reference builds it (src/phdfilter.cu.bak:2474-2479) */
__global__ void lfact_kernel(float* lf, int n) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float acc = 0.0f;
    lf[0] = 0.0f;
    for (int k = 1; k < n; ++k) {
      acc = acc + phd_safe_log((float)k);
      lf[k] = acc;
    }
  }
}

/* dense layout: terms in blocks of 64; a block stores its 7 planes back to back (7 x 256 bytes), so one
 * warp iteration of the update kernel writes one contiguous 1792-byte block with immediate plane offsets */
__host__ __device__ __forceinline__ size_t dense_index(size_t t) { return (t >> 6) * (PHD_NPLANES * 64) + (t & 63); }

__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }

/* phd_expf on two values at once; bit-identical per element to the scalar function */
__device__ __forceinline__ float2 phd_expf2(float2 x) {
  const float2 nm = __ffma2_rn(x, splat2(1.44269504088896341f), splat2(12582912.0f));
  const float2 n = __fadd2_rn(nm, splat2(-12582912.0f));
  float2 r = __ffma2_rn(n, splat2(-0.693359375f), x);
  r = __ffma2_rn(n, splat2(2.12194440e-4f), r);
  const float2 z = __fmul2_rn(r, r);
  float2 p = splat2(1.9875691500e-4f);
  p = __ffma2_rn(p, r, splat2(1.3981999507e-3f));
  p = __ffma2_rn(p, r, splat2(8.3334519073e-3f));
  p = __ffma2_rn(p, r, splat2(4.1665795894e-2f));
  p = __ffma2_rn(p, r, splat2(1.6666665459e-1f));
  p = __ffma2_rn(p, r, splat2(5.0000001201e-1f));
  p = __ffma2_rn(p, z, r);
  p = __fadd2_rn(p, splat2(1.0f));
  float2 sc;
  sc.x = __uint_as_float(((__float_as_uint(nm.x) - 0x4B400000u) + 127u) << 23);
  sc.y = __uint_as_float(((__float_as_uint(nm.y) - 0x4B400000u) + 127u) << 23);
  float2 v = __fmul2_rn(p, sc);
  /* x < -87.3 -> 0, x > 88 -> +inf, NaN propagates through v (same results as the scalar early returns) */
  v.x = (x.x < -87.3f) ? 0.0f : v.x;
  v.y = (x.y < -87.3f) ? 0.0f : v.y;
  v.x = (x.x > 88.0f) ? INFINITY : v.x;
  v.y = (x.y > 88.0f) ? INFINITY : v.y;
  return v;
}

/* phd_wrap_angle for |a| < 2*pi: r = a, or (a -+ 2pi) +- err when |a| >= float(pi) */
__device__ __forceinline__ float2 wrap_small2(float2 a) {
  const unsigned sx = __float_as_uint(a.x) & 0x80000000u, sy = __float_as_uint(a.y) & 0x80000000u;
  const bool px = fabsf(a.x) >= PHD_PI_F, py = fabsf(a.y) >= PHD_PI_F;
  float2 o1, o2;
  const unsigned tp = __float_as_uint(PHD_TWO_PI_F), er = __float_as_uint(PHD_TWO_PI_ERR);
  o1.x = px ? __uint_as_float((sx ^ 0x80000000u) | tp) : 0.0f;   /* -sign(a) * float(2*pi) */
  o1.y = py ? __uint_as_float((sy ^ 0x80000000u) | tp) : 0.0f;
  o2.x = px ? __uint_as_float(sx | er) : 0.0f;                   /* sign(a) * (float(2*pi) - 2*pi) */
  o2.y = py ? __uint_as_float(sy | er) : 0.0f;
  return __fadd2_rn(__fadd2_rn(a, o1), o2);
}

__device__ __forceinline__ float2 lds2(const float2* p) { return *p; }

/* One 64-component chunk of one measurement for one warp: lane owns components jr, jr+1.
 * PASS 1 accumulates exp(log-weight); PASS 2 normalises, writes the dense terms and emits the prune survivors. */
/* STASH (PHD): pass 1 leaves exp(log-weight) of the pair in the warp's shared-memory stash and pass 2 normalises it with
 * one multiplication, w = e * exp(-L_m) (oracle: the same product), instead of evaluating the exponential a second time;
 * nL2 is then exp(-L_m).  Without STASH (CPHD: all first passes run before any second pass) pass 2 recomputes
 * w = exp(log-weight + nL2). */
/* cmk (fused mode, pass 1 only; nullptr otherwise): receives the chunk's largest exp(log-weight) (STASH) or largest
 * log-weight, so that pass 2 can leave out chunks that cannot hold a survivor of the prune (upd_measurement). */
template <int PASS, bool TAIL, bool FAST, bool DENSE, bool STASH>
__device__ __forceinline__ void upd_chunk(const float2* __restrict__ rec, int jr, int C, float2 zr2, float2 zb2, bool dead,
                                          float2 nL2, float2& acc, float* __restrict__ Dm, bool even, float min_w, int lane,
                                          int* s_ncand, float4* __restrict__ cand, int Smax, int tbase_m, float2* __restrict__ stash,
                                          float* __restrict__ cmk = nullptr) {
  bool v0 = true, v1 = true;
  if (TAIL) {
    v0 = jr < C;
    v1 = jr + 1 < C;
    if (!v0) rec -= (size_t)(jr >> 1) * UPD_NF;        /* masked lanes read record 0 */
  }
  const float2 i0 = __fadd2_rn(zr2, lds2(rec + F_NR));
  float2 a1 = __fadd2_rn(zb2, lds2(rec + F_NB));
  float2 i1;
  if (FAST) {
    i1 = wrap_small2(a1);
  } else {
    i1.x = phd_wrap_angle(a1.x);
    i1.y = phd_wrap_angle(a1.y);
  }
  /* NOTE ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false, so the canonical
   * arithmetic of this loop spells every multiply-add as an explicit fused multiply-add (oracle: fmaf). */
  float2 lw = splat2(0.0f);
  if (PASS == 1 || !STASH) {
    float2 d = __fmul2_rn(__fmul2_rn(i0, i0), lds2(rec + F_S0));
    d = __ffma2_rn(__fmul2_rn(i0, i1), lds2(rec + F_S12), d);
    d = __ffma2_rn(__fmul2_rn(i1, i1), lds2(rec + F_S3), d);
    const float2 g = __fadd2_rn(__ffma2_rn(d, splat2(-0.5f), splat2(-PHD_LOG_2PI_F)), lds2(rec + F_NHL));
    lw = __fadd2_rn(lds2(rec + F_BASE), g);
    if (dead) lw = splat2(PHD_LOG0);
  }
  if (PASS == 1) {
    float2 e = phd_expf2(lw);
    if (TAIL) {
      if (!v0) e.x = 0.0f;
      if (!v1) e.y = 0.0f;
    }
    if (STASH) stash[jr >> 1] = e;
    acc = __fadd2_rn(acc, e);
    if (!DENSE && cmk) {
      if (STASH) {     /* e >= +0 (or NaN, which is the largest bit pattern): unsigned order = float order */
        const unsigned mx = __reduce_max_sync(FULL_MASK, max(__float_as_uint(e.x), __float_as_uint(e.y)));
        if (lane == 0) *cmk = __uint_as_float(mx);
      } else {
        unsigned k0 = float_to_ordered_uint(lw.x), k1 = float_to_ordered_uint(lw.y);
        if (TAIL) {
          if (!v0) k0 = 0u;
          if (!v1) k1 = 0u;
        }
        const unsigned mx = __reduce_max_sync(FULL_MASK, max(k0, k1));
        if (lane == 0) *cmk = ordered_uint_to_float(mx);
      }
    }
  } else {
    float2 wt = STASH ? __fmul2_rn(stash[jr >> 1], nL2) : phd_expf2(__fadd2_rn(lw, nL2));
    if (TAIL) {
      if (!v0) wt.x = 0.0f;
      if (!v1) wt.y = 0.0f;
    }
    /* survivors of the prune (w >= minFeatureWeight, flags :2308-2319) go to the merge as 32-byte records; their slots
     * are requested here and used after the dense stores, so that the round trip of the atomic is hidden */
    const bool k0 = v0 && !(wt.x < min_w), k1 = v1 && !(wt.y < min_w);
    const unsigned b0 = __ballot_sync(FULL_MASK, k0), b1 = __ballot_sync(FULL_MASK, k1);
    int slot0 = 0;
    if ((b0 | b1) && lane == 0) slot0 = atomicAdd(s_ncand, __popc(b0) + __popc(b1));
    const float2 m0 = __ffma2_rn(lds2(rec + F_K2), i1, __ffma2_rn(lds2(rec + F_K0), i0, lds2(rec + F_MX)));
    const float2 m1 = __ffma2_rn(lds2(rec + F_K3), i1, __ffma2_rn(lds2(rec + F_K1), i0, lds2(rec + F_MY)));
    const float2 c0 = lds2(rec + F_CU0), c1 = lds2(rec + F_CU1), c2 = lds2(rec + F_CU2), c3 = lds2(rec + F_CU3);
    const int t = tbase_m + jr;                 /* term index of the first component of the pair */
    if (DENSE && v0) {
      if (even) {                               /* t is even: the pair sits in one 64-term block, 8-byte aligned */
        float2* q = reinterpret_cast<float2*>(Dm + dense_index((size_t)t));
        __stcs(q, c0); __stcs(q + 32, c1); __stcs(q + 64, c2); __stcs(q + 96, c3);
        __stcs(q + 128, m0); __stcs(q + 160, m1); __stcs(q + 192, wt);
      } else {
        float* q = Dm + dense_index((size_t)t);
        __stcs(q, c0.x); __stcs(q + 64, c1.x); __stcs(q + 128, c2.x); __stcs(q + 192, c3.x);
        __stcs(q + 256, m0.x); __stcs(q + 320, m1.x); __stcs(q + 384, wt.x);
        if (v1) {
          q = Dm + dense_index((size_t)t + 1);
          __stcs(q, c0.y); __stcs(q + 64, c1.y); __stcs(q + 128, c2.y); __stcs(q + 192, c3.y);
          __stcs(q + 256, m0.y); __stcs(q + 320, m1.y); __stcs(q + 384, wt.y);
        }
      }
    }
    acc = __fadd2_rn(acc, wt);
    if (b0 | b1) {
      const unsigned lt_mask = (1u << lane) - 1u;
      slot0 = __shfl_sync(FULL_MASK, slot0, 0);
      /* slots in TERM order: components 2 lane, 2 lane + 1 follow those of the lower lanes */
      const int pos = __popc(b0 & lt_mask) + __popc(b1 & lt_mask);
      if (k0) {
        int slot = slot0 + pos;
        if (slot < Smax) {
          cand[2 * slot] = make_float4(c0.x, c1.x, c2.x, c3.x);
          cand[2 * slot + 1] = make_float4(m0.x, m1.x, wt.x, __int_as_float(t));
        }
      }
      if (k1) {
        int slot = slot0 + pos + (k0 ? 1 : 0);
        if (slot < Smax) {
          cand[2 * slot] = make_float4(c0.y, c1.y, c2.y, c3.y);
          cand[2 * slot + 1] = make_float4(m0.y, m1.y, wt.y, __int_as_float(t + 1));
        }
      }
    }
  }
}

/* Fused mode (no dense output) only -- cm != nullptr: pass 1 records per 64-component chunk the largest exp(log-weight)
 * (PHD) or the largest log-weight (CPHD); pass 2 skips a chunk none of whose terms can survive the prune:
 *   PHD : w = e * sc is monotone in e, so fmul(max e, sc) < min_w is EXACTLY "no survivor in this chunk";
 *   CPHD: w = exp(lw + D): skipped when max lw + D < skip_thr = log(min_w) - 0.01 (the margin covers the 2 ulp of the
 *         deterministic exp; a chunk inside the margin is simply processed).
 * Nothing else reads the detection weights in that mode (the sum of the detection weights is needed by Vo's empty-map
 * weighting only, for which the caller passes cm = nullptr), so the maps, weights and cardinalities are bit-identical. */
template <int PASS, bool FAST, bool DENSE, bool STASH>
__device__ __forceinline__ float2 upd_measurement(const float2* __restrict__ rec0, int C, float2 zr2, float2 zb2, bool dead,
                                                  float2 nL2, float* __restrict__ Dm, bool even, float min_w, int lane,
                                                  int* s_ncand, float4* __restrict__ cand, int Smax, int tbase_m,
                                                  float2* __restrict__ stash, float* __restrict__ cm = nullptr,
                                                  float skip_thr = 0.0f) {
  float2 acc = make_float2(0.0f, 0.0f);
  const int nfull = C >> 6;
  const float2* rec = rec0 + (size_t)lane * UPD_NF;
  int jr = 2 * lane;
  for (int k = 0; k < nfull; ++k) {
    bool skip = false;
    if (!DENSE && PASS == 2 && cm) skip = STASH ? (__fmul_rn(cm[k], nL2.x) < min_w) : (cm[k] + nL2.x < skip_thr);
    if (!skip)
      upd_chunk<PASS, false, FAST, DENSE, STASH>(rec, jr, C, zr2, zb2, dead, nL2, acc, Dm, even, min_w, lane, s_ncand, cand, Smax, tbase_m, stash,
                                                 (!DENSE && PASS == 1 && cm) ? cm + k : nullptr);
    rec += 32 * UPD_NF;
    jr += 64;
  }
  if (C & 63) {
    bool skip = false;
    if (!DENSE && PASS == 2 && cm) skip = STASH ? (__fmul_rn(cm[nfull], nL2.x) < min_w) : (cm[nfull] + nL2.x < skip_thr);
    if (!skip)
      upd_chunk<PASS, true, FAST, DENSE, STASH>(rec, jr, C, zr2, zb2, dead, nL2, acc, Dm, even, min_w, lane, s_ncand, cand, Smax, tbase_m, stash,
                                                (!DENSE && PASS == 1 && cm) ? cm + nfull : nullptr);
  }
  return acc;
}

template <bool DENSE, bool CPHD, bool MIXED>
__device__ __forceinline__ void update_body(const UpdArgs& a) {
  extern __shared__ __align__(16) float smem[];
  const DevCfg& c = a.c;
  const int Cmax = c.Cmax;
  float2* s_rec = reinterpret_cast<float2*>(smem);          /* (Cmax/2) records of UPD_NF float2 */
  float* s_w = smem + (size_t)UPD_NF * Cmax;
  float* s_mx = s_w + Cmax;
  float* s_my = s_mx + Cmax;
  float* s_pxx = s_my + Cmax;
  float* s_pxy = s_pxx + Cmax;
  float* s_pyy = s_pxy + Cmax;
  float* s_nd = s_pyy + Cmax;           /* non-detection weights (scheme-1 particle weighting) */
  float* s_tmp = s_nd + Cmax;           /* Cmax + 256 : pd*w, then one birth weight per measurement */
  float* s_zr = s_tmp + Cmax + PHD_MAX_MEAS;
  float* s_zb = s_zr + PHD_MAX_MEAS;
  float* s_zl = s_zb + PHD_MAX_MEAS;
  float* s_L = s_zl + PHD_MAX_MEAS;
  float* s_ds = s_L + PHD_MAX_MEAS;
  float* s_bb = s_ds + PHD_MAX_MEAS;    /* 5 x PHD_MAX_MEAS: birth covariance (b0, b1, b3) and mean per measurement; the remaining slack covers even padding */
  __shared__ int s_wcnt[UPD_WARPS];
  __shared__ int s_ncand;
  __shared__ float s_cphd_scal[2];
  __shared__ float s_cmw[UPD_WARPS][UPD_MAXCH];   /* fused PHD: chunk maxima of the measurement a warp is working on */

  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int pl = a.p0 + blockIdx.x;     /* local particle index */
  const int M = a.M;
  const int n = a.n;
  const int cnt = a.count[pl];
  const float px = a.pose[0 * (size_t)n + pl], py = a.pose[1 * (size_t)n + pl], pth = a.pose[2 * (size_t)n + pl];
  const float* mp = a.map + (size_t)pl * PHD_MAP_PLANES * Cmax;
  const uint8_t* cl = a.cls + (size_t)pl * Cmax;
  float4* cand = a.cand + (size_t)blockIdx.x * a.Smax * 2;
  const int Smax = a.Smax;
  const unsigned lt_mask = (1u << lane) - 1u;

  if (tid == 0) s_ncand = 0;
  for (int m = tid; m < M; m += UPD_THREADS) {
    s_zr[m] = a.z[m];
    s_zb[m] = a.z[PHD_MAX_MEAS + m];
    s_zl[m] = a.z[2 * PHD_MAX_MEAS + m];
  }

  /* ---- phase 0: stable compaction of the class-1 components into shared memory ---- */
  int C = 0;
  for (int base = 0; base < cnt; base += UPD_THREADS) {
    int i = base + tid;
    bool in = (i < cnt) && (cl[i] == 1);
    unsigned bal = __ballot_sync(FULL_MASK, in);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < UPD_WARPS; ++w) {
      int cw = s_wcnt[w];
      if (w < warp) woff += cw;
      tot += cw;
    }
    if (in) {
      int pos = C + woff + __popc(bal & lt_mask);
      s_w[pos] = mp[0 * Cmax + i];
      s_mx[pos] = mp[1 * Cmax + i];
      s_my[pos] = mp[2 * Cmax + i];
      s_pxx[pos] = mp[3 * Cmax + i];
      s_pxy[pos] = mp[4 * Cmax + i];
      s_pyy[pos] = mp[5 * Cmax + i];
    }
    C += tot;
    __syncthreads();
  }

  float* D = nullptr;
  if (DENSE) D = a.dense + (a.toff[pl] - a.tbase) * PHD_NPLANES;

  /* ---- phase 1: per-component EKF constants (preUpdateSynthKernel :1835-1894) + non-detection terms ---- */
  for (int j0 = 0; j0 < C; j0 += UPD_THREADS) {
    const int j = j0 + tid;
    bool keep = false;
    float P0 = 0, P1 = 0, P3 = 0, fx = 0, fy = 0, wnd = 0;
    if (j < C) {
      const float w = s_w[j];
      fx = s_mx[j]; fy = s_my[j];
      P0 = s_pxx[j]; P1 = s_pxy[j]; P3 = s_pyy[j];
      const float P2 = P1;
      float dx = fx - px;
      float dy = fy - py;
      float r2 = dx * dx + dy * dy;
      float r = sqrtf(r2);
      float bearing = phd_wrap_angle(phd_atan2f(dy, dx) - pth);
      float pd = 0.0f;
      if (r <= c.max_range && fabsf(bearing) <= c.max_bearing) pd = c.pd;
      float J0 = dx / r, J2 = dy / r, J1 = -dy / r2, J3 = dx / r2;
      float sg0 = (P0 * J0 + J2 * P1) * J0 + (J0 * P2 + P3 * J2) * J2 + c.var_r;
      float sg1 = (P0 * J1 + J3 * P1) * J0 + (J1 * P2 + P3 * J3) * J2;
      float sg2 = (P0 * J0 + J2 * P1) * J1 + (J0 * P2 + P3 * J2) * J3;
      float sg3 = (P0 * J1 + J3 * P1) * J1 + (J1 * P2 + P3 * J3) * J3 + c.var_b;
      sg1 = (sg1 + sg2) / 2.0f;
      sg2 = sg1;
      float det = sg0 * sg3 - sg1 * sg2;
      float S0 = sg3 / det, S1 = -sg1 / det, S2 = -sg2 / det, S3 = sg0 / det;
      float K0 = S0 * (P0 * J0 + P2 * J2) + S1 * (P0 * J1 + P2 * J3);
      float K1 = S0 * (P1 * J0 + P3 * J2) + S1 * (P1 * J1 + P3 * J3);
      float K2 = S2 * (P0 * J0 + P2 * J2) + S3 * (P0 * J1 + P2 * J3);
      float K3 = S2 * (P1 * J0 + P3 * J2) + S3 * (P1 * J1 + P3 * J3);
      float qa = 1.0f - K0 * J0 - K2 * J1;
      float qb = -K0 * J2 - K2 * J3;
      float qc = -K1 * J0 - K3 * J1;
      float qd = 1.0f - K1 * J2 - K3 * J3;
      float cu0 = (qa * P0 + qb * P1) * qa + (qa * P2 + qb * P3) * qb + K0 * K0 * c.var_r + K2 * K2 * c.var_b;
      float cu2 = (qa * P0 + qb * P1) * qc + (qa * P2 + qb * P3) * qd + K0 * c.var_r * K1 + K2 * c.var_b * K3;
      float cu1 = (qc * P0 + qd * P1) * qa + (qc * P2 + qd * P3) * qb + K0 * c.var_r * K1 + K2 * c.var_b * K3;
      float cu3 = (qc * P0 + qd * P1) * qc + (qc * P2 + qd * P3) * qd + K1 * K1 * c.var_r + K3 * K3 * c.var_b;
      float* rec = reinterpret_cast<float*>(s_rec + (size_t)(j >> 1) * UPD_NF) + (j & 1);
      rec[2 * F_NR] = -r; rec[2 * F_NB] = -bearing;
      rec[2 * F_K0] = K0; rec[2 * F_K1] = K1; rec[2 * F_K2] = K2; rec[2 * F_K3] = K3;
      rec[2 * F_S0] = S0; rec[2 * F_S12] = S1 + S2; rec[2 * F_S3] = S3;
      rec[2 * F_BASE] = phd_safe_log(pd) + phd_safe_log(w);
      rec[2 * F_NHL] = -(0.5f * phd_safe_log(det));
      rec[2 * F_CU0] = cu0; rec[2 * F_CU1] = cu1; rec[2 * F_CU2] = cu2; rec[2 * F_CU3] = cu3;
      rec[2 * F_MX] = fx; rec[2 * F_MY] = fy;
      s_tmp[j] = CPHD ? pd : pd * w;
      /* non-detection term (:2137-2141); CPHD scales it by <Psi1,p>/<Psi0,p> later (phase 2b) */
      wnd = w * (1.0f - pd);
      s_nd[j] = wnd;
      if (!CPHD) {
        if (DENSE) {
          float* q = D + dense_index((size_t)j);
          st_stream(q, P0); st_stream(q + 64, P1); st_stream(q + 128, P2); st_stream(q + 192, P3);
          st_stream(q + 256, fx); st_stream(q + 320, fy); st_stream(q + 384, wnd);
        }
        keep = !(wnd < c.min_w);
      }
    }
    /* survivors of the prune go to the merge as 32-byte records (warp-aggregated slot allocation) */
    unsigned bal = __ballot_sync(FULL_MASK, keep);
    if (bal) {
      int slot0 = 0;
      if (lane == 0) slot0 = atomicAdd(&s_ncand, __popc(bal));
      slot0 = __shfl_sync(FULL_MASK, slot0, 0);
      if (keep) {
        int slot = slot0 + __popc(bal & lt_mask);
        if (slot < Smax) {
          cand[2 * slot] = make_float4(P0, P1, P1, P3);
          cand[2 * slot + 1] = make_float4(fx, fy, wnd, __int_as_float(j));
        }
      }
    }
  }
  /* an odd component count leaves half a record: fill it with neutral values */
  if (tid == 0 && (C & 1)) {
    float* rec = reinterpret_cast<float*>(s_rec + (size_t)(C >> 1) * UPD_NF) + 1;
#pragma unroll
    for (int f = 0; f < UPD_NF; ++f) rec[2 * f] = 0.0f;
    rec[2 * F_BASE] = PHD_LOG0;
  }
  for (int m = tid; m < M; m += UPD_THREADS) {
    s_tmp[C + m] = c.birth_weight;
    /* birth term of measurement m (host loop :3468-3507): one thread per measurement, not one lane per warp round */
    const float zr = s_zr[m], zb = s_zb[m];
    float sn, cs;
    phd_sincosf(pth + zb, &sn, &cs);
    const float bdx = zr * cs;
    const float bdy = zr * sn;
    const float J0 = bdx / zr, J1 = bdy / zr, J2 = -bdy, J3 = bdx;
    s_bb[m] = J0 * J0 * c.bvar_r + J2 * J2 * c.bvar_b;
    s_bb[PHD_MAX_MEAS + m] = J0 * J1 * c.bvar_r + J2 * J3 * c.bvar_b;
    s_bb[2 * PHD_MAX_MEAS + m] = J1 * J1 * c.bvar_r + J3 * J3 * c.bvar_b;
    s_bb[3 * PHD_MAX_MEAS + m] = px + bdx;
    s_bb[4 * PHD_MAX_MEAS + m] = py + bdy;
  }
  __syncthreads();

  /* predicted cardinality (:2133-2186) and, for scheme 1, the prior / non-detect weight sums */
  float card_predict = 0.0f, cn_predict = 0.0f, nd_sum = 0.0f;
  if (warp == 0) {
    /* mixed model (phdUpdateKernelMixed, src/phdfilter.cu:2398-2446): the births do not count, the dynamic features do */
    card_predict = MIXED ? (warp_sum_array(s_tmp, C) + a.mix_nhat[pl]) : warp_sum_array(s_tmp, C + M);
    if (c.particle_weighting == 1) {
      cn_predict = warp_sum_array(s_w, C);
      nd_sum = warp_sum_array(s_nd, C);
    }
  }

  /* PHD: the per-warp stash of pass 1 (Cmax / 2 float2 per warp) overlays the phase-0/1 staging arrays s_w .. s_tmp,
   * which are dead once warp 0 has taken the sums above */
  float2* s_stash = reinterpret_cast<float2*>(s_w) + (size_t)warp * (Cmax >> 1);
  if (!CPHD) __syncthreads();
  const bool even = ((C & 1) == 0);      /* 64-bit stores need (C + m*C + j) even for every m */
  /* fused mode: chunk skipping in pass 2 (upd_measurement); not for Vo's weighting, which sums every detection weight */
  const int nch = (Cmax + 63) >> 6;
  const bool use_skip = !DENSE && nch <= UPD_MAXCH && (CPHD || c.particle_weighting != 1);
  float* s_cmm = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(smem) + update_smem_bytes(Cmax) +
                                          (CPHD ? cphd_smem_bytes(c.n_card, M) : 0));   /* CPHD: [M][nch] */
  const float skip_thr = (c.min_w > 0.0f) ? phd_logf(c.min_w) - 0.01f : -INFINITY;
  if (CPHD) {
    /* ---- CPHD phase 2a: likelihood mass S_m = sum_j exp(partial log-weight) of every measurement ---- */
    for (int m = warp; m < M; m += UPD_WARPS) {
      const float zr = s_zr[m], zb = s_zb[m];
      const bool dead = c.labeled && (s_zl[m] != 0.0f);
      const bool fast = fabsf(zb) < 3.14159f;
      float2 acc;
      float* cm = use_skip ? s_cmm + (size_t)m * nch : nullptr;
      if (fast)
        acc = upd_measurement<1, true, DENSE, false>(s_rec, C, splat2(zr), splat2(zb), dead, splat2(0.0f), D, even, c.min_w, lane, &s_ncand, cand, Smax, 0, nullptr, cm);
      else
        acc = upd_measurement<1, false, DENSE, false>(s_rec, C, splat2(zr), splat2(zb), dead, splat2(0.0f), D, even, c.min_w, lane, &s_ncand, cand, Smax, 0, nullptr, cm);
      float sum = warp_butterfly_sum(acc.x + acc.y);
      if (lane == 0) s_L[m] = sum;
    }
    __syncthreads();
    /* ---- multi-object terms: s_L[m] becomes the log factor D_m of measurement m ---- */
    cphd_block(c, C, M, a.lfact, a.card + (size_t)pl * c.n_card, a.write_card, s_w, s_nd, s_L, s_cphd_scal,
               reinterpret_cast<unsigned char*>(smem) + update_smem_bytes(Cmax));
    const float ND = s_cphd_scal[0];
    /* ---- CPHD phase 2b: non-detection terms (cphdUpdateKernel :1803-1820) ---- */
    for (int j0 = 0; j0 < C; j0 += UPD_THREADS) {
      const int j = j0 + tid;
      bool keep = false;
      float P0 = 0, P1 = 0, P3 = 0, fx = 0, fy = 0, wnd = 0;
      if (j < C) {
        fx = s_mx[j]; fy = s_my[j];
        P0 = s_pxx[j]; P1 = s_pxy[j]; P3 = s_pyy[j];
        wnd = phd_expf((phd_safe_log(s_w[j]) + ND) + phd_safe_log(1.0f - s_tmp[j]));
        if (DENSE) {
          float* q = D + dense_index((size_t)j);
          st_stream(q, P0); st_stream(q + 64, P1); st_stream(q + 128, P1); st_stream(q + 192, P3);
          st_stream(q + 256, fx); st_stream(q + 320, fy); st_stream(q + 384, wnd);
        }
        keep = !(wnd < c.min_w);
      }
      unsigned bal = __ballot_sync(FULL_MASK, keep);
      if (bal) {
        int slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(&s_ncand, __popc(bal));
        slot0 = __shfl_sync(FULL_MASK, slot0, 0);
        if (keep) {
          int slot = slot0 + __popc(bal & lt_mask);
          if (slot < Smax) {
            cand[2 * slot] = make_float4(P0, P1, P1, P3);
            cand[2 * slot + 1] = make_float4(fx, fy, wnd, __int_as_float(j));
          }
        }
      }
    }
  }
  /* ---- phase 2: detection terms; warp w owns measurements w, w+8, ... (:1898-1923, :2190-2252) ---- */
  for (int m = warp; m < M; m += UPD_WARPS) {
    const float zr = s_zr[m], zb = s_zb[m];
    const bool dead = c.labeled && (s_zl[m] != 0.0f);
    const float2 zr2 = splat2(zr), zb2 = splat2(zb);
    /* |zb| < 3.14159 and |bearing| <= float(pi) => |zb - bearing| < float(2*pi): wrap needs no fmod */
    const bool fast = fabsf(zb) < 3.14159f;
    const int tbase_m = C + m * C;
    float L;
    float2 wacc;
    if (CPHD) {
      L = -s_L[m];                        /* weights = exp(partial log-weight + D_m) (cphdUpdateKernel :1794-1799) */
      float* cm = use_skip ? s_cmm + (size_t)m * nch : nullptr;
      if (fast)
        wacc = upd_measurement<2, true, DENSE, false>(s_rec, C, zr2, zb2, dead, splat2(-L), D, even, c.min_w, lane, &s_ncand, cand, Smax, tbase_m, nullptr, cm, skip_thr);
      else
        wacc = upd_measurement<2, false, DENSE, false>(s_rec, C, zr2, zb2, dead, splat2(-L), D, even, c.min_w, lane, &s_ncand, cand, Smax, tbase_m, nullptr, cm, skip_thr);
    } else {
      float2 acc;
      float* cm = use_skip ? s_cmw[warp] : nullptr;
      if (fast)
        acc = upd_measurement<1, true, DENSE, true>(s_rec, C, zr2, zb2, dead, splat2(0.0f), D, even, c.min_w, lane, &s_ncand, cand, Smax, tbase_m, s_stash, cm);
      else
        acc = upd_measurement<1, false, DENSE, true>(s_rec, C, zr2, zb2, dead, splat2(0.0f), D, even, c.min_w, lane, &s_ncand, cand, Smax, tbase_m, s_stash, cm);
      __syncwarp();                       /* lane 0's chunk maxima are read by the whole warp in pass 2 */
      float sum = warp_butterfly_sum(acc.x + acc.y);
      if (MIXED) sum = sum + a.mix_dsum[(size_t)pl * PHD_MAX_MEAS + m];   /* :2461-2473 */
      sum = sum + c.clutter_density;
      sum = sum + c.birth_weight;
      if (MIXED && !c.labeled) sum = sum + c.birth_weight;                /* two birth terms per unlabelled measurement, :2478-2480 */
      L = phd_safe_log(sum);
      if (MIXED && lane == 0) a.mix_L[(size_t)pl * PHD_MAX_MEAS + m] = L;
      /* every lane reads back only what it stashed itself: no barrier between the passes */
      const float2 sc2 = splat2(phd_expf(-L));
      if (fast)
        wacc = upd_measurement<2, true, DENSE, true>(s_rec, C, zr2, zb2, dead, sc2, D, even, c.min_w, lane, &s_ncand, cand, Smax, tbase_m, s_stash, cm);
      else
        wacc = upd_measurement<2, false, DENSE, true>(s_rec, C, zr2, zb2, dead, sc2, D, even, c.min_w, lane, &s_ncand, cand, Smax, tbase_m, s_stash, cm);
      __syncwarp();                       /* before pass 1 of the warp's next measurement overwrites the maxima */
    }
    /* the sum of the detection weights is needed by Vo's empty-map particle weighting only (:2264-2279) */
    float dsum = 0.0f;
    if (c.particle_weighting == 1) dsum = warp_butterfly_sum(wacc.x + wacc.y);
    if (lane == 0) {
      if (!CPHD) s_L[m] = L;
      s_ds[m] = dsum;
    }
  }
  /* birth terms of this warp's measurements (host loop :3468-3507, normalised at :2232-2242): one lane per measurement
   * after the loop instead of lane 0 inside every round; each lane reads what lane 0 of its own warp wrote */
  __syncwarp();
  for (int m = warp + UPD_WARPS * lane; m < M; m += UPD_WARPS * 32) {
    const bool dead = c.labeled && (s_zl[m] != 0.0f);
    const float L = CPHD ? -s_L[m] : s_L[m];
    const float b0 = s_bb[m], b1 = s_bb[PHD_MAX_MEAS + m], b3 = s_bb[2 * PHD_MAX_MEAS + m];
    const float bmx = s_bb[3 * PHD_MAX_MEAS + m], bmy = s_bb[4 * PHD_MAX_MEAS + m];
    const float lb = dead ? PHD_LOG0 : c.log_birth_weight;
    const float wb = phd_expf(lb - L);
    const int t = C + M * C + m;
    if (DENSE) {
      float* q = D + dense_index((size_t)t);
      st_stream(q, b0); st_stream(q + 64, b1); st_stream(q + 128, b1); st_stream(q + 192, b3);
      st_stream(q + 256, bmx); st_stream(q + 320, bmy); st_stream(q + 384, wb);
    }
    if (!(wb < c.min_w)) {
      const int slot = atomicAdd(&s_ncand, 1);
      if (slot < Smax) {
        cand[2 * slot] = make_float4(b0, b1, b1, b3);
        cand[2 * slot + 1] = make_float4(bmx, bmy, wb, __int_as_float(t));
      }
    }
    s_ds[m] = s_ds[m] + wb;
  }
  __syncthreads();

  /* ---- phase 3: particle log-weight increment (:2258-2279) ---- */
  if (tid == 0) {
    float pw = 0.0f;
    for (int m = 0; m < M; ++m) pw = pw + s_L[m];
    float out;
    if (CPHD) {
      out = s_cphd_scal[1];              /* log <Psi0, p> (src/phdfilter.cu.bak:2666) */
    } else if (c.particle_weighting == 0) {
      out = pw - card_predict;
    } else if (c.particle_weighting == 1) {
      float cn_update = nd_sum;
      for (int m = 0; m < M; ++m) cn_update = cn_update + s_ds[m];
      out = (float)M * c.clutter_density + cn_update - cn_predict - c.clutter_rate;
    } else {
      out = 0.0f;
    }
    a.dlogw[pl] = out;
    a.n_cand[pl] = s_ncand;
  }
}

template <bool DENSE, bool CPHD>
__global__ void __launch_bounds__(UPD_THREADS) update_kernel(UpdArgs a) {
  update_body<DENSE, CPHD, false>(a);
}
/* the static map of the mixed feature model (feature_model = 2; PHD only) */
template <bool DENSE>
__global__ void __launch_bounds__(UPD_THREADS) update_mixed_kernel(UpdArgs a) {
  update_body<DENSE, false, true>(a);
}

/* =========================================================================================== */
/* prune + merge: pruneMap + mergeAndCopyMaps + phdUpdateMergeKernel                             */
/* (reference src/phdfilter.cu:3120-3333, 2707-2898; computeMahalDist device_math.cuh:309-325)    */
/*                                                                                               */
/* One WARP per particle, no block-level barriers.  The reference re-scans every update term of   */
/* the particle three times per output component; here                                            */
/*  A. the warp takes the records of the terms that survived the prune (emitted unordered by the  */
/*     update kernel), restores term order with a stable LSD radix sort on the term index,        */
/*     and appends the "nearly in range" components (:3243-3252);                                 */
/*  B. candidates are ranked once by (weight desc, index asc) with a radix sort -- weights of     */
/*     unmerged candidates never change, so every greedy arg-max is the next unmerged entry;      */
/*  C. candidates are binned into a uniform grid whose cell size is the canonical gate radius     */
/*     (oracle: merge_gate_radius2), so a seed only examines its 3x3 cell neighbourhood;           */
/*  D. each round evaluates gate + Mahalanobis/Hellinger distance for those few candidates, then  */
/*     accumulates the moment-matched merge over the members in ascending index order (the        */
/*     canonical order of the oracle).                                                            */
/* =========================================================================================== */
#define MRG_THREADS 128
#define MRG_WARPS (MRG_THREADS / 32)
#define MRG_GMAX 24          /* grid is at most 24 x 24 cells */
#define MRG_NCELL (MRG_GMAX * MRG_GMAX)
#define MRG_STAGE 8          /* members of one merge round staged in shared memory (larger clusters take the general path) */

struct MrgArgs {
  int M, n, p0, p1;
  const float* map_in; const int* count_in; const uint8_t* cls;
  float* map_out; int* count_out;
  const float4* cand_in;               /* [p1-p0][Smax][2] from the update kernel (unordered) */
  const int* n_cand;                   /* [n] */
  const int* n_in;                     /* [n] in-range components per particle */
  float4* cand;                        /* [p1-p0][Smax][2] scratch: candidates in canonical (term) order */
  Reductions* red;
  int Smax;
  /* fast kernel: shared-memory capacity (candidates per particle); particles above it are queued for merge_kernel */
  int Scap;
  int* ovf_list;                       /* [n] local particle indices queued by merge_fast_kernel */
  int use_list;                        /* merge_kernel: 1 = process red->ovf_n particles of ovf_list instead of [p0, p1) */
  DevCfg c;
};

__host__ __device__ static inline size_t merge_warp_smem_bytes(int Smax) {
  /* staging | cell ids (u16 Smax) | counters (1280 B) | cell starts (u16, 1168 B) | order, items (u16 Smax each) | alive + memb words */
  return (size_t)MRG_STAGE * 32 + 16 + (size_t)Smax * 2 + 1280 + 1168 + (size_t)Smax * 4 + (size_t)(Smax / 32) * 8;
}
static inline size_t merge_smem_bytes(int Smax) { return merge_warp_smem_bytes(Smax) * MRG_WARPS; }

__device__ __forceinline__ float dev_mahal(float ac0, float ac1, float ac2, float ac3, float am0, float am1,
                                           float bc0, float bc1, float bc2, float bc3, float bm0, float bm1) {
  float s0 = (ac0 + bc0) / 2.0f, s1 = (ac1 + bc1) / 2.0f, s2 = (ac2 + bc2) / 2.0f, s3 = (ac3 + bc3) / 2.0f;
  float det = s0 * s3 - s2 * s1;
  float rdet = 1.0f / det;
  float v0 = s3 * rdet, v1 = -s1 * rdet, v2 = -s2 * rdet, v3 = s0 * rdet;
  float i0 = am0 - bm0;
  float i1 = am1 - bm1;
  return i0 * i0 * v0 + i0 * i1 * (v1 + v2) + i1 * i1 * v3;
}

__device__ __forceinline__ float dev_hellinger(float ac0, float ac1, float ac2, float ac3, float am0, float am1,
                                               float bc0, float bc1, float bc2, float bc3, float bm0, float bm1) {
  float innov0 = am0 - bm0, innov1 = am1 - bm1;
  float s0 = ac0 + bc0, s1 = ac1 + bc1, s2 = ac2 + bc2, s3 = ac3 + bc3;
  float det = s0 * s3 - s2 * s1;
  float v0 = 1.0f, v1 = 0.0f, v2 = 0.0f, v3 = 1.0f;
  if (det > FLT_MIN) {
    v0 = s3 / det; v1 = -s1 / det; v2 = -s2 / det; v3 = s0 / det;
  }
  float eps = -0.25f * (innov0 * innov0 * v0 + innov0 * innov1 * (v1 + v2) + innov1 * innov1 * v3);
  det = det / 4.0f;
  float dist = 1.0f / det;
  float p0 = ac0 * bc0 + ac2 * bc1;
  float p1 = ac1 * bc0 + ac3 * bc1;
  float p2 = ac0 * bc2 + ac2 * bc3;
  float p3 = ac1 * bc2 + ac3 * bc3;
  float detp = p0 * p3 - p2 * p1;
  dist = dist * sqrtf(detp);
  dist = 1.0f - sqrtf(dist) * phd_expf(eps);
  return dist;
}

/* largest eigenvalue of a candidate covariance (oracle: merge_lambda_max) */
__device__ __forceinline__ float dev_lambda_max(float4 cv) {
  float t = cv.x + cv.w;
  float det = cv.x * cv.w - cv.y * cv.z;
  float disc = fmaxf(t * t - 4.0f * det, 0.0f);
  return 0.5f * (t + sqrtf(disc));
}

/* exclusive scan of 256 u32 counters held 8 per lane; returns the 8 exclusive prefixes in ex[] */
__device__ __forceinline__ void warp_scan256(const unsigned* cnt, unsigned ex[8], int lane) {
  unsigned loc[8], sum = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    loc[k] = cnt[lane * 8 + k];
    sum += loc[k];
  }
  unsigned inc = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    unsigned t = __shfl_up_sync(FULL_MASK, inc, off);
    if (lane >= off) inc += t;
  }
  unsigned run = inc - sum;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    ex[k] = run;
    run += loc[k];
  }
}

/* One stable LSD radix pass (8-bit digit) over the index list `in` -> `out`, keyed by
 * key_of(idx) = bits of recs[2*idx+1].{w|z} (optionally transformed).  Warp-synchronous. */
template <int MODE> /* 0: key = term index (record .w); 1: key = ~ordered(weight) (record .z); 2: bitrev8(idx mod 256) */
__device__ __forceinline__ unsigned radix_key(const float4* recs, unsigned idx) {
  if (MODE == 2) return __brev(idx) >> 24;
  float4 r1 = recs[2 * idx + 1];
  if (MODE == 0) return __float_as_uint(r1.w);
  return ~float_to_ordered_uint(r1.z);
}
template <int MODE>
__device__ __forceinline__ void radix_pass(const float4* recs, const unsigned short* in, unsigned short* out, int n,
                                           unsigned* hist, int shift, int lane) {
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int i = lane; i < 256; i += 32) hist[i] = 0;
  __syncwarp();
  for (int i = lane; i < n; i += 32) atomicAdd(&hist[(radix_key<MODE>(recs, in[i]) >> shift) & 255u], 1u);
  __syncwarp();
  unsigned ex[8];
  warp_scan256(hist, ex, lane);
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 8; ++k) hist[lane * 8 + k] = ex[k];
  __syncwarp();
  for (int b0 = 0; b0 < n; b0 += 32) {
    int i = b0 + lane;
    unsigned idx = 0, d = 0x80000000u | (unsigned)lane;   /* inactive lanes: unique keys */
    if (i < n) {
      idx = in[i];
      d = (radix_key<MODE>(recs, idx) >> shift) & 255u;
    }
    unsigned same = __match_any_sync(FULL_MASK, d);
    unsigned before = (i < n) ? hist[d] : 0u;
    __syncwarp();
    if (i < n) {
      out[before + __popc(same & lt_mask)] = (unsigned short)idx;
      if ((same & lt_mask) == 0) hist[d] = before + __popc(same);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(MRG_THREADS) merge_kernel(MrgArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DevCfg& c = a.c;
  const int Smax = a.Smax;
  const int lane = lane_id(), warp = warp_id();
  int pl = a.p0 + blockIdx.x * MRG_WARPS + warp;
  if (a.use_list) {         /* the particles merge_fast_kernel could not hold in shared memory */
    const int k = blockIdx.x * MRG_WARPS + warp;
    if (k >= a.red->ovf_n) return;
    pl = a.ovf_list[k];
  }
  if (pl >= a.p1) return;   /* warps are independent: no block barrier below */

  unsigned char* base0 = smem_raw + (size_t)warp * merge_warp_smem_bytes(Smax);
  float4* s_stage = (float4*)base0;                                    /* MRG_STAGE staged member records */
  unsigned char* s_perm = base0 + MRG_STAGE * 32;                      /* rank -> staging slot */
  unsigned char* base = base0 + MRG_STAGE * 32 + 16;
  unsigned short* s_cell = (unsigned short*)base;                      /* cy * MRG_GMAX + cx per candidate */
  unsigned* s_cnt = (unsigned*)(base + Smax * 2);                      /* 256 radix counters / 576 packed u16 cell counters */
  unsigned short* s_start = (unsigned short*)(base + Smax * 2 + 1280); /* MRG_NCELL + 1 cell starts */
  unsigned short* s_order = (unsigned short*)(base + Smax * 2 + 1280 + 1168);
  unsigned short* s_items = s_order + Smax;
  unsigned* s_alive = (unsigned*)(s_items + Smax);
  unsigned* s_memb = s_alive + Smax / 32;

  const int Cmax = c.Cmax;
  const int cnt = a.count_in[pl];
  const float* mp = a.map_in + (size_t)pl * PHD_MAP_PLANES * Cmax;
  const uint8_t* cl = a.cls + (size_t)pl * Cmax;
  float* mo = a.map_out + (size_t)pl * PHD_MAP_PLANES * Cmax;
  const float4* cin = a.cand_in + (size_t)(pl - a.p0) * Smax * 2;
  float4* cand = a.cand + (size_t)(pl - a.p0) * Smax * 2;
  const unsigned lt_mask = (1u << lane) - 1u;

  /* ---- A. survivors of the prune (flags :2308-2319, pruneMap :3120-3174) back into term order ---- */
  int n = a.n_cand[pl];
  if (n > Smax) {
    if (lane == 0) atomicOr(&a.red->err_flag, 1);
    n = Smax;
  }
  {
    const int T = a.M + (int)cnt * (a.M + 1);              /* upper bound of the term index */
    int bits = 32 - __clz(max(T, 1));
    for (int i = lane; i < n; i += 32) s_order[i] = (unsigned short)i;
    __syncwarp();
    unsigned short* src = s_order;
    unsigned short* dst = s_items;
    for (int shift = 0; shift < bits; shift += 8) {
      radix_pass<0>(cin, src, dst, n, s_cnt, shift, lane);
      unsigned short* t = src; src = dst; dst = t;
    }
    for (int i = lane; i < n; i += 32) {
      unsigned sidx = src[i];
      float4 r0 = cin[2 * sidx], r1 = cin[2 * sidx + 1];
      r1.w = dev_lambda_max(r0);          /* the term index is no longer needed: keep the gate eigenvalue there */
      cand[2 * i] = r0;
      cand[2 * i + 1] = r1;
    }
  }
  float tmax = 0.0f, xmin = FLT_MAX, xmax = -FLT_MAX, ymin = FLT_MAX, ymax = -FLT_MAX;
  /* nearly-in-range components in map order (:3243-3252) */
  for (int b0 = 0; b0 < cnt; b0 += 32) {
    int i = b0 + lane;
    bool k2 = (i < cnt) && (cl[i] == 2);
    unsigned bal = __ballot_sync(FULL_MASK, k2);
    if (k2) {
      int pos = n + __popc(bal & lt_mask);
      if (pos < Smax) {
        float pxy = mp[4 * Cmax + i];
        float4 r0 = make_float4(mp[3 * Cmax + i], pxy, pxy, mp[5 * Cmax + i]);
        cand[2 * pos] = r0;
        cand[2 * pos + 1] = make_float4(mp[1 * Cmax + i], mp[2 * Cmax + i], mp[0 * Cmax + i], dev_lambda_max(r0));
      }
    }
    n += __popc(bal);
  }
  if (n > Smax) {
    if (lane == 0) atomicOr(&a.red->err_flag, 1);
    n = Smax;
  }
  __syncwarp();   /* candidate records written by other lanes are read below */
  for (int i = lane; i < n; i += 32) {
    float4 r1 = cand[2 * i + 1];
    tmax = fmaxf(tmax, r1.w);
    xmin = fminf(xmin, r1.x); xmax = fmaxf(xmax, r1.x);
    ymin = fminf(ymin, r1.y); ymax = fmaxf(ymax, r1.y);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    tmax = fmaxf(tmax, __shfl_xor_sync(FULL_MASK, tmax, off));
    xmin = fminf(xmin, __shfl_xor_sync(FULL_MASK, xmin, off));
    xmax = fmaxf(xmax, __shfl_xor_sync(FULL_MASK, xmax, off));
    ymin = fminf(ymin, __shfl_xor_sync(FULL_MASK, ymin, off));
    ymax = fmaxf(ymax, __shfl_xor_sync(FULL_MASK, ymax, off));
  }

  int nout = 0;
  if (n > 0) {
    /* ---- B. rank by (weight desc, reference tie rule): stable LSD radix.  Equal weights are ordered as the
     * reference's 256-slot arg-max tree orders them: by (bitrev8(i mod 256), i div 256) (oracle merge_tie_key).
     * With n <= 256 the first pass alone realises that order; otherwise i div 256 is already ascending within
     * each residue because the pass is stable over the identity order. ---- */
    for (int i = lane; i < n; i += 32) s_items[i] = (unsigned short)i;
    __syncwarp();
    radix_pass<2>(cand, s_items, s_order, n, s_cnt, 0, lane);
    radix_pass<1>(cand, s_order, s_items, n, s_cnt, 0, lane);
    radix_pass<1>(cand, s_items, s_order, n, s_cnt, 8, lane);
    radix_pass<1>(cand, s_order, s_items, n, s_cnt, 16, lane);
    radix_pass<1>(cand, s_items, s_order, n, s_cnt, 24, lane);

    /* ---- C. uniform grid over candidate means; cell size >= the largest gate radius ---- */
    const bool gated = (c.distance_metric == 0);
    const float gk = 0.515625f * c.min_sep;                 /* pair gate: |d|^2 <= gk * (lam_a + lam_b) */
    int G = 1;
    float cs = 1.0f;
    if (gated) {
      float ext = fmaxf(xmax - xmin, ymax - ymin);
      float rg = sqrtf(gk * (tmax + tmax)) * 1.0001f;    /* tmax = largest eigenvalue among the candidates */
      if (rg >= 0.0f && rg < ext && ext < FLT_MAX) {     /* false for inf / NaN radius or degenerate extent */
        int g = (int)(ext / rg) + 1;
        if (g > MRG_GMAX) g = MRG_GMAX;
        G = g;
        cs = fmaxf(rg, (ext / (float)g) * 1.0001f);
      }
    }
    for (int i = lane; i < MRG_NCELL / 2; i += 32) s_cnt[i] = 0;        /* two u16 counters per word */
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      float4 r1 = cand[2 * i + 1];
      int cx = 0, cy = 0;
      if (G > 1) {
        cx = (int)((r1.x - xmin) / cs);
        cy = (int)((r1.y - ymin) / cs);
        cx = min(max(cx, 0), G - 1);
        cy = min(max(cy, 0), G - 1);
      }
      int cid = cy * MRG_GMAX + cx;
      s_cell[i] = (unsigned short)cid;
      atomicAdd(&s_cnt[cid >> 1], (cid & 1) ? 0x10000u : 1u);
    }
    __syncwarp();
    {
      /* exclusive scan of the MRG_NCELL (= 576) counters, 18 per lane */
      const unsigned short* c16 = (const unsigned short*)s_cnt;
      unsigned loc[MRG_NCELL / 32], sum = 0;
#pragma unroll
      for (int k = 0; k < MRG_NCELL / 32; ++k) {
        loc[k] = c16[lane * (MRG_NCELL / 32) + k];
        sum += loc[k];
      }
      unsigned inc = sum;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        unsigned t = __shfl_up_sync(FULL_MASK, inc, off);
        if (lane >= off) inc += t;
      }
      unsigned run = inc - sum;
      __syncwarp();
#pragma unroll
      for (int k = 0; k < MRG_NCELL / 32; ++k) {
        s_start[lane * (MRG_NCELL / 32) + k] = (unsigned short)run;
        run += loc[k];
      }
      if (lane == 31) s_start[MRG_NCELL] = (unsigned short)n;
    }
    __syncwarp();
    for (int i = lane; i < MRG_NCELL / 2; i += 32) s_cnt[i] = 0;        /* becomes the running fill count */
    __syncwarp();
    /* stable scatter in index order */
    for (int b0 = 0; b0 < n; b0 += 32) {
      int i = b0 + lane;
      unsigned cid = (i < n) ? (unsigned)s_cell[i] : (0x80000000u | (unsigned)lane);
      unsigned same = __match_any_sync(FULL_MASK, cid);
      unsigned short* f16 = (unsigned short*)s_cnt;
      unsigned before = (i < n) ? f16[cid] : 0u;
      __syncwarp();
      if (i < n) {
        s_items[s_start[cid] + before + __popc(same & lt_mask)] = (unsigned short)i;
        if ((same & lt_mask) == 0) f16[cid] = (unsigned short)(before + __popc(same));
      }
      __syncwarp();
    }
    for (int i = lane; i < Smax / 32; i += 32) {
      int lo = i * 32;
      unsigned w = 0;
      if (lo < n) w = (n - lo >= 32) ? 0xffffffffu : ((1u << (n - lo)) - 1u);
      s_alive[i] = w;
      s_memb[i] = 0;
    }
    __syncwarp();

    /* ---- D. greedy merge rounds (:2739-2894) ---- */
    const int nwords = (n + 31) / 32;
    int pos = 0;
    bool stop = false;
    while (!stop) {
      int seed = -1;
      while (pos < n) {
        int i = pos + lane;
        unsigned ci = (i < n) ? (unsigned)s_order[i] : 0u;
        bool al = (i < n) && ((s_alive[ci >> 5] >> (ci & 31)) & 1u);
        unsigned bal = __ballot_sync(FULL_MASK, al);
        if (bal) {
          int first = __ffs(bal) - 1;
          seed = (int)__shfl_sync(FULL_MASK, ci, first);
          pos += first;
          break;
        }
        pos += 32;
      }
      if (seed < 0) break;
      const float4 A0 = cand[2 * seed], A1 = cand[2 * seed + 1];
      const int scell = s_cell[seed];
      const int scy = scell / MRG_GMAX, scx = scell - scy * MRG_GMAX;
      const int cx0 = max(scx - 1, 0), cx1 = min(scx + 1, G - 1);
      int beg[3], len[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        int row = scy + d - 1;
        if (row >= 0 && row < G) {
          beg[d] = s_start[row * MRG_GMAX + cx0];
          len[d] = (int)s_start[row * MRG_GMAX + cx1 + 1] - beg[d];
        } else {
          beg[d] = 0;
          len[d] = 0;
        }
      }
      const int tot = len[0] + len[1] + len[2];
      float wsum = 0.0f, m0 = 0.0f, m1 = 0.0f;
      float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f, v3 = 0.0f, mm0 = 0.0f, mm1 = 0.0f, rw = 0.0f;
      /* candidates of the 3x3 neighbourhood, 32 per iteration; members stage their record in shared memory */
      int nm = 0;
      for (int q0 = 0; q0 < tot; q0 += 32) {
        int q = q0 + lane;
        bool memb = false;
        unsigned it = 0;
        float4 B0 = make_float4(0.f, 0.f, 0.f, 0.f), B1 = B0;
        if (q < tot) {
          int src = (q < len[0]) ? beg[0] + q : ((q < len[0] + len[1]) ? beg[1] + (q - len[0]) : beg[2] + (q - len[0] - len[1]));
          it = s_items[src];
          if ((s_alive[it >> 5] >> (it & 31)) & 1u) {
            B1 = cand[2 * it + 1];
            B0 = cand[2 * it];
            float gx = A1.x - B1.x, gy = A1.y - B1.y;
            if (!gated || (gx * gx + gy * gy <= gk * (A1.w + B1.w))) {
              float dist = (c.distance_metric == 0)
                               ? dev_mahal(A0.x, A0.y, A0.z, A0.w, A1.x, A1.y, B0.x, B0.y, B0.z, B0.w, B1.x, B1.y)
                               : dev_hellinger(A0.x, A0.y, A0.z, A0.w, A1.x, A1.y, B0.x, B0.y, B0.z, B0.w, B1.x, B1.y);
              memb = dist < c.min_sep;
            }
          }
        }
        const unsigned mb = __ballot_sync(FULL_MASK, memb);
        if (memb) {
          atomicOr(&s_memb[it >> 5], 1u << (it & 31));
          int slot = nm + __popc(mb & lt_mask);
          if (slot < MRG_STAGE) {
            s_stage[2 * slot] = B0;
            s_stage[2 * slot + 1] = make_float4(B1.x, B1.y, B1.z, __uint_as_float(it));
          }
        }
        nm += __popc(mb);
      }
      __syncwarp();
      if (nm <= MRG_STAGE) {
        /* ---- staged path: rank the (few) members by candidate index, then accumulate from shared memory ---- */
        if (lane < nm) {
          unsigned my = __float_as_uint(s_stage[2 * lane + 1].w);
          int rank = 0;
          for (int k = 0; k < nm; ++k) rank += (__float_as_uint(s_stage[2 * k + 1].w) < my) ? 1 : 0;
          s_perm[rank] = (unsigned char)lane;
          atomicAnd(&s_alive[my >> 5], ~(1u << (my & 31)));     /* retire */
          s_memb[my >> 5] = 0;
        }
        __syncwarp();
        for (int r = 0; r < nm; ++r) {
          float4 B1 = s_stage[2 * s_perm[r] + 1];
          wsum = wsum + B1.z;
          m0 = m0 + B1.z * B1.x;
          m1 = m1 + B1.z * B1.y;
        }
        if (wsum == 0.0f) {      /* :2821-2822: the reference abandons the remaining components */
          stop = true;
        } else {
          rw = 1.0f / wsum;
          mm0 = m0 * rw;
          mm1 = m1 * rw;
          for (int r = 0; r < nm; ++r) {
            int sl = s_perm[r];
            float4 B0 = s_stage[2 * sl], B1 = s_stage[2 * sl + 1];
            float d0 = mm0 - B1.x, d1 = mm1 - B1.y;
            v0 = v0 + B1.z * (B0.x + d0 * d0);
            v1 = v1 + B1.z * (B0.y + d0 * d1);
            v2 = v2 + B1.z * (B0.z + d1 * d0);
            v3 = v3 + B1.z * (B0.w + d1 * d1);
          }
        }
      } else {
        /* ---- general path: members walked in ascending index order through the membership bitmask ---- */
        for (int wg = 0; wg < nwords; wg += 32) {
          const unsigned mword = (wg + lane < nwords) ? s_memb[wg + lane] : 0u;
          for (unsigned hb = __ballot_sync(FULL_MASK, mword != 0u); hb; hb &= hb - 1) {
            int wd = __ffs(hb) - 1;
            unsigned mb = __shfl_sync(FULL_MASK, mword, wd);
            for (; mb; mb &= mb - 1) {
              int i = (wg + wd) * 32 + (__ffs(mb) - 1);
              float4 B1 = cand[2 * i + 1];
              wsum = wsum + B1.z;
              m0 = m0 + B1.z * B1.x;
              m1 = m1 + B1.z * B1.y;
            }
          }
        }
        if (wsum == 0.0f) {
          stop = true;
        } else {
          rw = 1.0f / wsum;
          mm0 = m0 * rw;
          mm1 = m1 * rw;
          for (int wg = 0; wg < nwords; wg += 32) {
            const unsigned mword = (wg + lane < nwords) ? s_memb[wg + lane] : 0u;
            for (unsigned hb = __ballot_sync(FULL_MASK, mword != 0u); hb; hb &= hb - 1) {
              int wd = __ffs(hb) - 1;
              unsigned mb = __shfl_sync(FULL_MASK, mword, wd);
              for (; mb; mb &= mb - 1) {
                int i = (wg + wd) * 32 + (__ffs(mb) - 1);
                float4 B0 = cand[2 * i], B1 = cand[2 * i + 1];
                float d0 = mm0 - B1.x, d1 = mm1 - B1.y;
                v0 = v0 + B1.z * (B0.x + d0 * d0);
                v1 = v1 + B1.z * (B0.y + d0 * d1);
                v2 = v2 + B1.z * (B0.z + d1 * d0);
                v3 = v3 + B1.z * (B0.w + d1 * d1);
              }
            }
            if (wg + lane < nwords) {     /* retire the members of this word group */
              s_alive[wg + lane] &= ~mword;
              s_memb[wg + lane] = 0;
            }
          }
        }
      }
      if (!stop) {
        v0 = v0 * rw; v1 = v1 * rw; v2 = v2 * rw; v3 = v3 * rw;
        v1 = (v1 + v2) / 2.0f;                       /* force_symmetric_covariance */
        if (nout < Cmax) {
          if (lane < PHD_MAP_PLANES) {       /* lane f stores plane f of the merged component */
            float val = (lane == 0) ? wsum : (lane == 1) ? mm0 : (lane == 2) ? mm1 : (lane == 3) ? v0 : (lane == 4) ? v1 : v3;
            mo[lane * Cmax + nout] = val;
          }
        } else if (lane == 0) {
          atomicOr(&a.red->err_flag, 2);
        }
        nout++;
      }
      __syncwarp();
    }
  }

  /* ---- E. re-append the far (class 0) components (:3311-3318) and publish the map size ---- */
  int pos0 = nout;
  for (int b0 = 0; b0 < cnt; b0 += 32) {
    int i = b0 + lane;
    bool k0 = (i < cnt) && (cl[i] == 0);
    unsigned bal = __ballot_sync(FULL_MASK, k0);
    if (k0) {
      int pos = pos0 + __popc(bal & lt_mask);
      if (pos < Cmax) {
#pragma unroll
        for (int f = 0; f < PHD_MAP_PLANES; ++f) mo[f * Cmax + pos] = mp[f * Cmax + i];
      }
    }
    pos0 += __popc(bal);
  }
  if (lane == 0) {
    if (pos0 > Cmax) {
      atomicOr(&a.red->err_flag, 2);
      pos0 = Cmax;
    }
    a.count_out[pl] = pos0;
  }
}

/* =========================================================================================== */
/* merge_fast_kernel: the same prune + merge (same candidates, same arithmetic, same output order   */
/* as merge_kernel and the oracle's merge_mixture), restructured so that nothing is serial in the   */
/* number of output components.                                                                     */
/*                                                                                                  */
/* The greedy reduction "take the heaviest unmerged candidate, absorb every unmerged candidate      */
/* within the distance threshold, repeat" is a priority maximal-independent-set problem: with the   */
/* candidates ranked by (weight desc, reference tie rule), candidate r is a SEED iff no seed of      */
/* lower rank is near it, and otherwise it belongs to the lowest-ranked seed near it (every member   */
/* of a seed's cluster has a higher rank than the seed, because the seed was the arg-max of what     */
/* was left).  One CTA of MF_WARPS warps per particle:                                               */
/*  1. block-wide stable radix sorts put the prune survivors back in term order and rank the          */
/*     candidates; the gate data {x, y, w, lambda_max} of every candidate sits in shared memory in   */
/*     RANK order, the full records go to the (L2-resident) scratch buffer in rank order;             */
/*  2. candidates are binned in the uniform grid of merge_kernel.  Cell by cell (cells dealt to the   */
/*     warps), the lanes hold the candidates of the cell's 3x3 neighbourhood and the candidates of    */
/*     the cell are broadcast against them: the cheap Euclidean gate runs on all lanes, pairs that    */
/*     pass are queued and the Mahalanobis distance is evaluated 32 queued pairs at a time -- no       */
/*     divergence on the expensive part.  A near pair (lower rank, higher rank) is appended to the    */
/*     higher-ranked candidate's linked list (nodes from a per-particle pool);                        */
/*  3. ownership is resolved from the lists in rank order, 32 ranks at a time (ballots);              */
/*  4. a counting sort by output slot, stable in candidate index, gives every cluster its members    */
/*     in ascending index order (the canonical accumulation order), and one THREAD per cluster       */
/*     accumulates the moment-matched merge; stores are coalesced across clusters.                   */
/* Particles with more candidates than the shared-memory capacity Scap (adapted by the host from     */
/* the previous step), or with more than 2.5 * Scap near pairs, are queued for merge_kernel.      */
/* Mahalanobis metric only.                                                                           */
/* =========================================================================================== */
#define MF_WARPS 4
#define MF_THREADS (MF_WARPS * 32)
#define MF_POOL2 5           /* near-pair pool: MF_POOL2 / 2 list nodes per candidate of capacity (a particle that needs more takes merge_kernel) */
#define MF_NONE 0xffffu
#define MF_CELLS 592         /* MRG_NCELL + 1, padded */
#define MF_QUEUE 128         /* gate survivors queued per warp (ring buffer) */
#define MF_SEG 8             /* cell-order positions per gate segment (pair phase); 16 measured slower (7.47 vs 7.00 ms) */
#define MF_ACH 8             /* A candidates gated per compaction step (up to MF_ACH * 32 new queue entries) */

__host__ __device__ static inline size_t merge_fast_smem_bytes(int S) {
  /* gate records 16 B | four u16 arrays (two of them double as the list heads) | near-pair pool | per-warp radix
   * histograms (the grid's cell ends alias them) | pair queues | self-threshold bits */
  return (size_t)S * (16 + 8 + 2 * MF_POOL2) + (size_t)MF_ACH * 16 + (size_t)MF_WARPS * 512 + (size_t)MF_WARPS * MF_QUEUE * 4 + (size_t)S / 8 + 64;
}

/* arr[idx] += v on a 4-byte aligned u16 array (no carry into the neighbour: counts stay < 65536); returns the old value */
__device__ __forceinline__ unsigned atomic_add_u16(unsigned short* arr, int idx, unsigned v) {
  unsigned old = atomicAdd(reinterpret_cast<unsigned*>(arr) + (idx >> 1), (idx & 1) ? (v << 16) : v);
  return (idx & 1) ? (old >> 16) : (old & 0xffffu);
}

/* in-place exclusive prefix sum of n u16 entries by one warp (lane-contiguous blocks) */
__device__ __forceinline__ void warp_exclusive_scan_u16(unsigned short* arr, int n, int lane) {
  const int per = (n + 31) >> 5;
  const int lo = min(lane * per, n), hi = min(lo + per, n);
  unsigned sum = 0;
  for (int i = lo; i < hi; ++i) sum += arr[i];
  unsigned inc = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    unsigned t = __shfl_up_sync(FULL_MASK, inc, off);
    if (lane >= off) inc += t;
  }
  unsigned run = inc - sum;
  __syncwarp();
  for (int i = lo; i < hi; ++i) {
    unsigned v = arr[i];
    arr[i] = (unsigned short)run;
    run += v;
  }
  __syncwarp();
}

/* Block-wide stable LSD radix pass (8-bit digit) over the index list in -> out, keys in shared memory.
 * Warp w owns a contiguous slice of the input; per-warp digit histograms make the scatter stable.
 * TIE: the key is bitrev8(idx mod 256) (oracle: merge_tie_key).  Ends with a block barrier. */
#define MF_RADIX_R 6         /* a warp's slice of up to 32 * MF_RADIX_R elements is held in registers across a pass */
template <bool TIE>
__device__ __forceinline__ void block_radix_pass(const unsigned* keys, const unsigned short* in, unsigned short* out, int n,
                                                 unsigned short* hist /* [MF_WARPS][256] */, int shift) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned* h32 = reinterpret_cast<unsigned*>(hist);
  for (int i = tid; i < MF_WARPS * 128; i += MF_THREADS) h32[i] = 0;
  const int per = (((n + MF_WARPS - 1) / MF_WARPS) + 31) & ~31;
  const int lo = min(warp * per, n), hi = min(lo + per, n);
  unsigned short* hw = hist + warp * 256;
  /* (index, digit) of the slice's elements: loaded once, back to back, and used by the count AND the scatter */
  const bool inreg = per <= 32 * MF_RADIX_R;
  unsigned pk[MF_RADIX_R];
  if (inreg) {
#pragma unroll
    for (int r = 0; r < MF_RADIX_R; ++r) {
      const int i = lo + lane + 32 * r;
      pk[r] = (i < hi) ? (unsigned)in[i] : 0xffffffffu;
    }
#pragma unroll
    for (int r = 0; r < MF_RADIX_R; ++r)
      if (pk[r] != 0xffffffffu) pk[r] |= (TIE ? (__brev(pk[r]) >> 24) : ((keys[pk[r]] >> shift) & 255u)) << 16;
  }
  __syncthreads();
  if (inreg) {
#pragma unroll
    for (int r = 0; r < MF_RADIX_R; ++r)
      if (pk[r] != 0xffffffffu) atomic_add_u16(hw, (int)(pk[r] >> 16), 1u);
  } else {
    for (int i = lo + lane; i < hi; i += 32) {
      const unsigned idx = in[i];
      const unsigned d = TIE ? (__brev(idx) >> 24) : ((keys[idx] >> shift) & 255u);
      atomic_add_u16(hw, (int)d, 1u);      /* result unused: a fire-and-forget shared-memory reduction (measured: aggregating
                                              equal digits with match_any first is slower -- a dependent load/store chain) */
    }
  }
  __syncthreads();
  if (warp == 0) {
    unsigned tot[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned t = 0;
#pragma unroll
      for (int w = 0; w < MF_WARPS; ++w) t += hist[w * 256 + lane * 8 + k];
      tot[k] = t;
      sum += t;
    }
    unsigned inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned t = __shfl_up_sync(FULL_MASK, inc, off);
      if (lane >= off) inc += t;
    }
    unsigned run = inc - sum;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned o = run;
#pragma unroll
      for (int w = 0; w < MF_WARPS; ++w) {
        const unsigned cw = hist[w * 256 + lane * 8 + k];
        hist[w * 256 + lane * 8 + k] = (unsigned short)o;
        o += cw;
      }
      run += tot[k];
    }
  }
  __syncthreads();
  if (inreg) {
#pragma unroll
    for (int r = 0; r < MF_RADIX_R; ++r) {
      if (lo + 32 * r >= hi) break;                       /* warp-uniform */
      const bool valid = pk[r] != 0xffffffffu;
      const unsigned idx = pk[r] & 0xffffu;
      const unsigned d = valid ? (pk[r] >> 16) : (0x80000000u | (unsigned)lane);   /* inactive lanes: unique keys */
      const unsigned same = __match_any_sync(FULL_MASK, d);
      const unsigned before = valid ? hw[d] : 0u;
      __syncwarp();
      if (valid) {
        out[before + __popc(same & lt_mask)] = (unsigned short)idx;
        if ((same & lt_mask) == 0) hw[d] = (unsigned short)(before + __popc(same));
      }
      __syncwarp();
    }
  } else {
    for (int b0 = lo; b0 < hi; b0 += 32) {
      const int i = b0 + lane;
      unsigned idx = 0, d = 0x80000000u | (unsigned)lane;   /* inactive lanes: unique keys */
      if (i < hi) {
        idx = in[i];
        d = TIE ? (__brev(idx) >> 24) : ((keys[idx] >> shift) & 255u);
      }
      const unsigned same = __match_any_sync(FULL_MASK, d);
      const unsigned before = (i < hi) ? hw[d] : 0u;
      __syncwarp();
      if (i < hi) {
        out[before + __popc(same & lt_mask)] = (unsigned short)idx;
        if ((same & lt_mask) == 0) hw[d] = (unsigned short)(before + __popc(same));
      }
      __syncwarp();
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(MF_THREADS, 8) merge_fast_kernel(MrgArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ float s_red[MF_WARPS][5];
  __shared__ int s_cnt2[MF_WARPS];
  __shared__ int s_scan[MF_WARPS];
  __shared__ int s_flag;                 /* a near list overflowed: the particle goes to merge_kernel */
  __shared__ int s_kzero;                /* first cluster with zero weight (:2821-2822) */
  __shared__ int s_nseeds, s_klimit, s_n, s_pool_n, s_stopr;
  const DevCfg& c = a.c;
  const int S = a.Scap;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pl = a.p0 + blockIdx.x;
  const unsigned lt_mask = (1u << lane) - 1u;

  float4* Gc = reinterpret_cast<float4*>(smem_raw);                    /* [S] gate records {mx, my, lambda_max, rank} in CELL order */
  unsigned short* P1 = reinterpret_cast<unsigned short*>(Gc + S + MF_ACH);   /* [S] source slot per candidate; later rank per candidate
                                                                                (the gate records are padded: the A side is read MF_ACH at a time) */
  unsigned short* Abuf = P1 + S;
  unsigned short* Bbuf = Abuf + S;
  unsigned short* own = Bbuf + S;                                      /* [S] grid cell, then rank of the owning seed, then output slot */
  unsigned* HD = reinterpret_cast<unsigned*>(Abuf);                    /* [S] near list of a candidate (by rank): head node, MF_NONE = empty
                                                                          (over the two sort buffers, free between the record load and the output slots) */
  unsigned* pool = reinterpret_cast<unsigned*>(own + S);               /* [MF_POOL2 * S / 2] list nodes: lower rank | next node << 16 */
  unsigned short* hist = reinterpret_cast<unsigned short*>(pool + (MF_POOL2 * S) / 2);   /* [MF_WARPS][256]; the grid's cell ends alias it */
  unsigned* queue = reinterpret_cast<unsigned*>(hist + MF_WARPS * 256);/* [MF_WARPS][MF_QUEUE] */
  unsigned* selfbits = queue + MF_WARPS * MF_QUEUE;                    /* [S / 32] candidate (by rank) is within its own threshold */
  unsigned* K = reinterpret_cast<unsigned*>(Gc);                       /* sort keys (before the gate records are built) */
  unsigned* Wtmp = pool;                                               /* weight keys by emission slot (before the lists exist) */
  unsigned short* cend = hist;

  const int Cmax = c.Cmax;
  const int cnt = a.count_in[pl];
  const float* mp = a.map_in + (size_t)pl * PHD_MAP_PLANES * Cmax;
  const uint8_t* cl = a.cls + (size_t)pl * Cmax;
  float* mo = a.map_out + (size_t)pl * PHD_MAP_PLANES * Cmax;
  const float4* cin = a.cand_in + (size_t)(pl - a.p0) * a.Smax * 2;
  float4* crec = a.cand + (size_t)(pl - a.p0) * a.Smax * 2;            /* full records in rank order */

  /* ---- A. does the particle fit? ---- */
  const int n1 = a.n_cand[pl];
  {
    int c2 = 0;
    for (int b0 = warp * 32; b0 < cnt; b0 += MF_THREADS) {
      const int i = b0 + lane;
      c2 += __popc(__ballot_sync(FULL_MASK, (i < cnt) && (cl[i] == 2)));
    }
    if (lane == 0) s_cnt2[warp] = c2;
    if (tid == 0) {
      s_flag = 0;
      s_kzero = 0x7fffffff;
      s_pool_n = 0;
    }
  }
  __syncthreads();
  int n2 = 0;
#pragma unroll
  for (int w = 0; w < MF_WARPS; ++w) n2 += s_cnt2[w];
  if (tid == 0) atomicMax(&a.red->max_cand, min(n1, a.Smax) + n2);
  if (n1 > a.Smax || n1 + n2 > S) {
    if (tid == 0) a.ovf_list[atomicAdd(&a.red->ovf_n, 1)] = pl;
    return;
  }

  /* ---- B. survivors of the prune back into term order (pruneMap :3120-3174); weight keys on the way.  The update kernel
   * emitted them segment by segment (upd_nseg): the slots of a segment are consecutive and in term order, so the first
   * and last slot of every segment follow from comparing the segment of each record with its slot neighbours', an
   * exclusive prefix sum of the segment sizes in term order gives every segment its first position. ---- */
  const int Cin = a.n_in[pl];
  const int nseg = upd_nseg(Cin, a.M);
  unsigned short* seg_first = reinterpret_cast<unsigned short*>(K + S);   /* [nseg] (the rest of the gate-record area) */
  unsigned short* seg_last = seg_first + nseg;
  if ((size_t)nseg * 4 > (size_t)S * 12) {          /* no room for the table: the general kernel sorts instead */
    if (tid == 0) a.ovf_list[atomicAdd(&a.red->ovf_n, 1)] = pl;
    return;
  }
  for (int i = tid; i < n1; i += MF_THREADS) {
    const float4 r1 = cin[2 * i + 1];
    Abuf[i] = (unsigned short)upd_segment_of(__float_as_int(r1.w), Cin, a.M);
    Wtmp[i] = ~float_to_ordered_uint(r1.z);
  }
  for (int sg = tid; sg < nseg; sg += MF_THREADS) {
    seg_first[sg] = 1;                               /* empty: last - first + 1 = 0 */
    seg_last[sg] = 0;
  }
  __syncthreads();
  for (int i = tid; i < n1; i += MF_THREADS) {
    const unsigned sg = Abuf[i];
    if (i == 0 || Abuf[i - 1] != sg) seg_first[sg] = (unsigned short)i;
    if (i == n1 - 1 || Abuf[i + 1] != sg) seg_last[sg] = (unsigned short)i;
  }
  __syncthreads();
  {
    const int per = (nseg + MF_THREADS - 1) / MF_THREADS;
    const int s_lo = min(tid * per, nseg), s_hi = min(s_lo + per, nseg);
    int sum = 0;
    for (int sg = s_lo; sg < s_hi; ++sg) sum += (int)seg_last[sg] - (int)seg_first[sg] + 1;
    int inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int t = __shfl_up_sync(FULL_MASK, inc, off);
      if (lane >= off) inc += t;
    }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    int start = inc - sum;
#pragma unroll
    for (int w = 0; w < MF_WARPS; ++w)
      if (w < warp) start += s_scan[w];
    for (int sg = s_lo; sg < s_hi; ++sg) {            /* seg_last becomes the first POSITION of the segment */
      const int cnt = (int)seg_last[sg] - (int)seg_first[sg] + 1;
      seg_last[sg] = (unsigned short)start;
      start += cnt;
    }
  }
  __syncthreads();
  for (int i = tid; i < n1; i += MF_THREADS) {        /* slot i -> its position in term order */
    const unsigned sg = Abuf[i];
    const int pos = (int)seg_last[sg] + (i - (int)seg_first[sg]);
    if (pos < n1) {
      P1[pos] = (unsigned short)i;
      K[pos] = Wtmp[i];
    }
  }
  /* nearly-in-range components in map order (:3243-3252) */
  if (warp == 0) {
    int n = n1;
    for (int b0 = 0; b0 < cnt; b0 += 32) {
      const int i = b0 + lane;
      const bool k2 = (i < cnt) && (cl[i] == 2);
      const unsigned bal = __ballot_sync(FULL_MASK, k2);
      if (k2) {
        const int pos = n + __popc(bal & lt_mask);
        P1[pos] = (unsigned short)i;
        K[pos] = ~float_to_ordered_uint(mp[i]);
      }
      n += __popc(bal);
    }
    if (lane == 0) s_n = n;
  }
  __syncthreads();
  const int n = s_n;

  int nout = 0;
  if (n > 0) {
    /* ---- C. rank by (weight desc, reference tie rule), see merge_kernel step B ---- */
    for (int i = tid; i < n; i += MF_THREADS) Abuf[i] = (unsigned short)i;
    __syncthreads();
    unsigned short* src = Abuf;
    unsigned short* dst = Bbuf;
    block_radix_pass<true>(K, src, dst, n, hist, 0);
    { unsigned short* t = src; src = dst; dst = t; }
    for (int shift = 0; shift < 32; shift += 8) {
      block_radix_pass<false>(K, src, dst, n, hist, shift);
      unsigned short* t = src; src = dst; dst = t;
    }
    unsigned short* ord = src;       /* rank -> candidate index; later the cluster ends */
    unsigned short* items = dst;     /* output slot of a seed */

    /* ---- D. candidate records in rank order to the scratch buffer (stays in L1/L2).  While a record is in registers: is
     * the candidate within the threshold of itself (it is, unless its covariance is degenerate)?  Its mean is staged in
     * shared memory (the tail of the near-pair pool, free until the pair phase) for the grid of phase E. ---- */
    const float gk = 0.515625f * c.min_sep;                   /* pair gate: |d|^2 <= gk * (lam_a + lam_b) */
    float2* st_xy = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(pool) + 2 * (size_t)S);   /* [S], behind cellr */
    float tmax = 0.0f, xmin = FLT_MAX, xmax = -FLT_MAX, ymin = FLT_MAX, ymax = -FLT_MAX;
    for (int rb = warp * 32; rb < n; rb += MF_THREADS) {      /* a warp covers 32 consecutive ranks: one selfbits word */
      const int r = rb + lane;
      bool self = false;
      if (r < n) {
        const int i = ord[r];
        const int s = P1[i];
        float4 r0, r1;
        if (i < n1) {
          r0 = cin[2 * s];
          r1 = cin[2 * s + 1];
        } else {
          const float pxy = mp[4 * Cmax + s];
          r0 = make_float4(mp[3 * Cmax + s], pxy, pxy, mp[5 * Cmax + s]);
          r1 = make_float4(mp[1 * Cmax + s], mp[2 * Cmax + s], mp[s], 0.0f);
        }
        r1.w = dev_lambda_max(r0);
        crec[2 * r] = r0;
        crec[2 * r + 1] = r1;
        P1[i] = (unsigned short)r;     /* candidate index -> rank */
        st_xy[r] = make_float2(r1.x, r1.y);
        tmax = fmaxf(tmax, r1.w);
        xmin = fminf(xmin, r1.x); xmax = fmaxf(xmax, r1.x);
        ymin = fminf(ymin, r1.y); ymax = fmaxf(ymax, r1.y);
        self = 0.0f <= gk * (r1.w + r1.w) &&
               dev_mahal(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r0.x, r0.y, r0.z, r0.w, r1.x, r1.y) < c.min_sep;
      }
      const unsigned sb = __ballot_sync(FULL_MASK, self);
      if (lane == 0) selfbits[rb >> 5] = sb;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      tmax = fmaxf(tmax, __shfl_xor_sync(FULL_MASK, tmax, off));
      xmin = fminf(xmin, __shfl_xor_sync(FULL_MASK, xmin, off));
      xmax = fmaxf(xmax, __shfl_xor_sync(FULL_MASK, xmax, off));
      ymin = fminf(ymin, __shfl_xor_sync(FULL_MASK, ymin, off));
      ymax = fmaxf(ymax, __shfl_xor_sync(FULL_MASK, ymax, off));
    }
    if (lane == 0) {
      s_red[warp][0] = tmax; s_red[warp][1] = xmin; s_red[warp][2] = xmax; s_red[warp][3] = ymin; s_red[warp][4] = ymax;
    }
    for (int i = tid; i < MF_CELLS / 2; i += MF_THREADS) reinterpret_cast<unsigned*>(cend)[i] = 0;
    for (int i = ((n + 31) >> 5) + tid; i < (S >> 5); i += MF_THREADS) selfbits[i] = 0;  /* words beyond the last rank */
    __syncthreads();
#pragma unroll
    for (int w = 0; w < MF_WARPS; ++w) {
      tmax = fmaxf(tmax, s_red[w][0]);
      xmin = fminf(xmin, s_red[w][1]); xmax = fmaxf(xmax, s_red[w][2]);
      ymin = fminf(ymin, s_red[w][3]); ymax = fmaxf(ymax, s_red[w][4]);
    }

    /* ---- E. uniform grid over the candidate means, cell size >= the largest gate radius; gate records in cell order ---- */
    int G = 1;
    float cs = 1.0f;
    {
      const float ext = fmaxf(xmax - xmin, ymax - ymin);
      const float rg = sqrtf(gk * (tmax + tmax)) * 1.0001f;
      if (rg >= 0.0f && rg < ext && ext < FLT_MAX) {
        int g = (int)(ext / rg) + 1;
        if (g > MRG_GMAX) g = MRG_GMAX;
        G = g;
        cs = fmaxf(rg, (ext / (float)g) * 1.0001f);
      }
    }
    unsigned short* cellr = reinterpret_cast<unsigned short*>(pool);   /* cell of every candidate (the pool is free until the pair phase) */
    /* lambda_max of this thread's candidates for the second loop, requested now: the loads are in flight under the first
     * loop, the cell scan and two barriers (up to MF_RADIX_R rounds; larger capacities load inside the loop) */
    const bool lam_inreg = n <= MF_THREADS * MF_RADIX_R;
    float lam[MF_RADIX_R];
    if (lam_inreg) {
#pragma unroll
      for (int q = 0; q < MF_RADIX_R; ++q) {
        const int r = tid + q * MF_THREADS;
        lam[q] = (r < n) ? crec[2 * r + 1].w : 0.0f;
      }
    }
    for (int r = tid; r < n; r += MF_THREADS) {
      const float2 xy = st_xy[r];
      int cx = 0, cy = 0;
      if (G > 1) {
        cx = (int)((xy.x - xmin) / cs);
        cy = (int)((xy.y - ymin) / cs);
        cx = min(max(cx, 0), G - 1);
        cy = min(max(cy, 0), G - 1);
      }
      const int cid = cy * MRG_GMAX + cx;
      cellr[r] = (unsigned short)cid;
      atomic_add_u16(cend, cid, 1u);
    }
    __syncthreads();
    if (warp == 0) warp_exclusive_scan_u16(cend, MRG_NCELL, lane);        /* cell starts */
    __syncthreads();
    if (lam_inreg) {
#pragma unroll
      for (int q = 0; q < MF_RADIX_R; ++q) {
        const int r = tid + q * MF_THREADS;
        if (r < n) {
          const float2 xy = st_xy[r];
          const unsigned cid = cellr[r];
          const unsigned pos = atomic_add_u16(cend, (int)cid, 1u);
          Gc[pos] = make_float4(xy.x, xy.y, lam[q], (float)r);
          own[pos] = (unsigned short)cid;      /* the cell of every cell-order position, for the pair phase */
          HD[r] = MF_NONE;
        }
      }
    } else {
      for (int r = tid; r < n; r += MF_THREADS) {
        const float2 xy = st_xy[r];
        const unsigned cid = cellr[r];
        const unsigned pos = atomic_add_u16(cend, (int)cid, 1u);
        Gc[pos] = make_float4(xy.x, xy.y, crec[2 * r + 1].w, (float)r);
        own[pos] = (unsigned short)cid;
        HD[r] = MF_NONE;
      }
    }
    __syncthreads();                                       /* cend[c] is now the END of cell c; st_xy is dead: the pool is free */

    /* ---- F. near pairs (:2802-2806).  Every unordered pair of candidates in neighbouring cells is gated once by the
     * Euclidean test, one lane per candidate (see the loop below).  Pairs that pass are compacted into a per-warp ring
     * and the Mahalanobis distance is evaluated 32 queued pairs at a time. ---- */
    {
      unsigned* q = queue + warp * MF_QUEUE;   /* ring of pairs of cell-order positions */
      const int pool_cap = (MF_POOL2 * S) / 2;
      int qh = 0, qn = 0;
      /* seed = the lower-ranked candidate, the argument order of the reference.  A near pair becomes a node of the
       * higher-ranked candidate's list. */
      auto evaluate = [&](int cntq) {
        bool near = false;
        unsigned ar = 0, br = 0;
        if (lane < cntq) {
          const unsigned e = q[(qh + lane) & (MF_QUEUE - 1)];
          float4 A1 = Gc[e & 0xffffu], B1 = Gc[e >> 16];
          if (A1.w < B1.w) {           /* A = the higher rank (the candidate), B = the lower rank (the seed) */
            const float4 t = A1; A1 = B1; B1 = t;
          }
          ar = (unsigned)A1.w;
          br = (unsigned)B1.w;
          const float4 A0 = crec[2 * ar], B0 = crec[2 * br];
          near = dev_mahal(B0.x, B0.y, B0.z, B0.w, B1.x, B1.y, A0.x, A0.y, A0.z, A0.w, A1.x, A1.y) < c.min_sep;
        }
        const unsigned nb = __ballot_sync(FULL_MASK, near);
        if (nb) {
          int base = 0;
          if (lane == 0) base = atomicAdd(&s_pool_n, __popc(nb));
          base = __shfl_sync(FULL_MASK, base, 0);
          if (near) {
            const int k = base + __popc(nb & lt_mask);
            if (k < pool_cap) pool[k] = br | (atomicExch(&HD[ar], (unsigned)k) << 16);
            else s_flag = 1;
          }
        }
      };
      /* The A side of a B candidate is two contiguous runs of cell-order positions: the row above, cells cx-1..cx+1, and
       * its own row from cell cx-1 up to itself, so every unordered pair of candidates in neighbouring cells is gated
       * exactly once.  Neighbourhood sizes differ a lot between the 32 candidates of a warp round (clusters), so the runs
       * are cut into segments of MF_SEG positions and the SEGMENTS are dealt to the lanes (prefix sum over the round's
       * candidates, owner found by a shuffle binary search): ~80 % of the gate's lane slots do useful work instead of ~25 %. */
      for (int b0 = warp * 32; b0 < n; b0 += MF_THREADS) {
        const int pb = b0 + lane;
        int beg1 = 0, len1 = 0, beg2 = 0, tot = 0;
        if (pb < n) {
          const int cid = own[pb];                                     /* the cell of this position (phase E) */
          const int cy = cid / MRG_GMAX, cx = cid - cy * MRG_GMAX;
          const int left = cid - ((cx > 0) ? 1 : 0);
          beg1 = (left > 0) ? cend[left - 1] : 0;
          len1 = pb - beg1;
          int len2 = 0;
          if (cy > 0) {
            const int up = left - MRG_GMAX;
            beg2 = (up > 0) ? cend[up - 1] : 0;
            len2 = (int)cend[cid - MRG_GMAX + ((cx + 1 < G) ? 1 : 0)] - beg2;
          }
          tot = len1 + len2;
        }
        const int nseg = (tot + MF_SEG - 1) / MF_SEG;
        int incl = nseg;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(FULL_MASK, incl, off);
          if (lane >= off) incl += t;
        }
        const int T = __shfl_sync(FULL_MASK, incl, 31);
        const int excl = incl - nseg;
        for (int s0 = 0; s0 < T; s0 += 32) {
          const int sidx = s0 + lane;
          int o = 0;                                                   /* owner: the last candidate with excl <= sidx */
#pragma unroll
          for (int step = 16; step >= 1; step >>= 1) {
            const int e = __shfl_sync(FULL_MASK, excl, o + step);
            if (e <= sidx) o += step;
          }
          const int ob1 = __shfl_sync(FULL_MASK, beg1, o), ol1 = __shfl_sync(FULL_MASK, len1, o);
          const int ob2 = __shfl_sync(FULL_MASK, beg2, o) - ol1;
          const int t0 = (sidx - __shfl_sync(FULL_MASK, excl, o)) * MF_SEG;
          const int otot = __shfl_sync(FULL_MASK, tot, o);             /* every lane takes part in the shuffle */
          const int tend = (sidx < T) ? otot : 0;
          const int pbo = min(b0 + o, n - 1);
          const float4 B1 = Gc[pbo];
          const unsigned pbs = (unsigned)pbo << 16;
          unsigned mask = 0u;                                         /* bit k: position t0 + k passes the gate */
          /* consecutive lanes usually hold consecutive segments of one run, MF_SEG records = 128 bytes apart: read in step
           * they would all hit the same four banks (an 8-way conflict on every quarter-warp of the 128-bit loads, 58 % of
           * this kernel's shared-memory wavefronts were excess ones); each lane therefore starts at its own offset */
#pragma unroll
          for (int kk = 0; kk < MF_SEG; ++kk) {
            const int k = (kk + lane) & (MF_SEG - 1);
            const int t = t0 + k;
            if (t < tend) {
              const float4 A1 = Gc[((t < ol1) ? ob1 : ob2) + t];
              const float gx = B1.x - A1.x, gy = B1.y - A1.y;
              if (gx * gx + gy * gy <= gk * (B1.z + A1.z)) mask |= 1u << k;
            }
          }
          /* compaction: every lane appends its pairs to the ring */
          const int cntl = __popc(mask);
          int inc = cntl;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(FULL_MASK, inc, off);
            if (lane >= off) inc += t;
          }
          const int total = __shfl_sync(FULL_MASK, inc, 31);
          if (total == 0) continue;
          if (qn + total <= MF_QUEUE) {
            int slot = qh + qn + inc - cntl;
            for (unsigned m = mask; m; m &= m - 1) {
              const int t = t0 + (__ffs(m) - 1);
              q[slot & (MF_QUEUE - 1)] = (unsigned)(((t < ol1) ? ob1 : ob2) + t) | pbs;
              ++slot;
            }
            qn += total;
            __syncwarp();
            while (qn >= 32) {
              evaluate(32);
              qh = (qh + 32) & (MF_QUEUE - 1);
              qn -= 32;
            }
            __syncwarp();
          } else {
            /* a very dense neighbourhood: one pair per lane and round */
            while (__any_sync(FULL_MASK, mask != 0u)) {
              const bool has = mask != 0u;
              const unsigned hb = __ballot_sync(FULL_MASK, has);
              if (has) {
                const int t = t0 + (__ffs(mask) - 1);
                q[(qh + qn + __popc(hb & lt_mask)) & (MF_QUEUE - 1)] = (unsigned)(((t < ol1) ? ob1 : ob2) + t) | pbs;
                mask &= mask - 1;
              }
              qn += __popc(hb);
              __syncwarp();
              if (qn >= 32) {
                evaluate(32);
                qh = (qh + 32) & (MF_QUEUE - 1);
                qn -= 32;
              }
              __syncwarp();
            }
          }
        }
      }
      __syncwarp();
      evaluate(qn);
    }
    __syncthreads();
    for (int r = tid; r < n; r += MF_THREADS) own[r] = (unsigned short)MF_NONE;   /* held the cells until here */
    __syncthreads();

    /* ---- G. ownership from the near lists: candidate r is owned by the lowest-ranked SEED among its lower-ranked near
     * neighbours, and is a seed itself if there is none.  A candidate can decide once no undecided neighbour ranks
     * below its lowest-ranked seed neighbour; block-wide sweeps until every candidate has decided (the lowest-ranked
     * undecided candidate always can). ---- */
    if (!s_flag) {
      /* A candidate's list holds LOWER-ranked neighbours only, so the ranks are resolved in blocks of MF_THREADS, in
       * ascending order: every neighbour in an earlier block has decided, and the few neighbours inside the same block
       * are waited for by iterating on the block (reads of an iteration, barrier, writes, barrier: no thread reads a
       * state another thread is writing).  A candidate decides once no undecided neighbour ranks below its lowest-ranked
       * seed neighbour -- the lowest-ranked undecided candidate of the block always can, so every iteration makes progress. */
      for (int r0 = 0; r0 < n; r0 += MF_THREADS) {
        const int r = r0 + tid;
        bool todo = r < n;
        for (;;) {
          unsigned dec = MF_NONE;
          if (todo) {
            unsigned minseed = MF_NONE, minund = MF_NONE;
            for (unsigned node = HD[r]; node != MF_NONE;) {
              const unsigned nd = pool[node];
              const unsigned e = nd & 0xffffu;
              node = nd >> 16;
              const unsigned st = own[e];
              if (st == MF_NONE) minund = min(minund, e);
              else if (st == e) minseed = min(minseed, e);
            }
            if (minund == MF_NONE || minseed < minund) dec = (minseed != MF_NONE) ? minseed : (unsigned)r;
          }
          __syncthreads();
          if (dec != MF_NONE) {
            own[r] = (unsigned short)dec;
            todo = false;
          }
          if (!__syncthreads_or(todo ? 1 : 0)) break;
        }
      }
      /* output slots of the seeds, in rank order (= descending seed weight) */
      if (warp == 0) {
        unsigned stopr = MF_NONE;    /* first seed that is not within the threshold of itself */
        int nseeds = 0;
        for (int rb = 0; rb < n; rb += 32) {
          const int r = rb + lane;
          const bool isseed = (r < n) && (own[r] == r);
          const unsigned bal = __ballot_sync(FULL_MASK, isseed);
          if (isseed) items[r] = (unsigned short)(nseeds + __popc(bal & lt_mask));
          const unsigned sb = bal & ~selfbits[rb >> 5];
          if (sb && stopr == MF_NONE) stopr = (unsigned)(rb + __ffs(sb) - 1);
          nseeds += __popc(bal);
        }
        __syncwarp();
        if (lane == 0) {
          s_nseeds = nseeds;
          s_stopr = (int)stopr;
          /* a seed outside its own threshold ends the reduction after its cluster (the oracle picks it again, finds
           * nothing and stops, :2821-2822); it is not a member of its own cluster */
          s_klimit = (stopr != MF_NONE) ? (int)items[stopr] + 1 : nseeds;
        }
      }
    }
    __syncthreads();
    if (s_flag) {
      if (tid == 0) a.ovf_list[atomicAdd(&a.red->ovf_n, 1)] = pl;
      return;
    }
    const int nseeds = s_nseeds, klimit = s_klimit;
    unsigned short* slotA = items;
    unsigned short* csz = ord;       /* cluster sizes -> starts -> ends (the rank -> index table is no longer needed) */
    unsigned short* memb = reinterpret_cast<unsigned short*>(pool);   /* member lists (the near lists are no longer needed) */
    unsigned short* keyA = memb + S;                                  /* output slot of every candidate's cluster */
    const unsigned stopr = (unsigned)s_stopr;
    for (int k = tid; k < ((nseeds + 2) >> 1); k += MF_THREADS) reinterpret_cast<unsigned*>(csz)[k] = 0;
    __syncthreads();
    for (int r = tid; r < n; r += MF_THREADS) {
      unsigned key = slotA[own[r]];
      if ((unsigned)r == stopr) key = MF_NONE;
      keyA[r] = (unsigned short)key;
      if (key != MF_NONE) atomic_add_u16(csz, (int)key, 1u);
    }
    __syncthreads();
    if (warp == 0) {
      warp_exclusive_scan_u16(csz, nseeds, lane);
      /* members of every cluster in ascending candidate index: stable scatter over the candidate order */
      for (int ib = 0; ib < n; ib += 32) {
        const int i = ib + lane;
        unsigned r = 0, key = MF_NONE;
        if (i < n) {
          r = P1[i];
          key = keyA[r];
        }
        const unsigned mk = (key != MF_NONE) ? key : (0x80000000u | (unsigned)lane);
        const unsigned same = __match_any_sync(FULL_MASK, mk);
        const unsigned before = (key != MF_NONE) ? csz[key] : 0u;
        __syncwarp();
        if (key != MF_NONE) {
          memb[before + __popc(same & lt_mask)] = (unsigned short)r;
          if ((same & lt_mask) == 0) csz[key] = (unsigned short)(before + __popc(same));
        }
        __syncwarp();
      }
    }
    __syncthreads();

    /* ---- H. moment-matched merge, one thread per cluster (:2808-2881).  The members' records are first staged in shared
     * memory in MEMBER order -- all threads, independent L2 loads in flight -- over arrays that are dead by now (gate
     * records, candidate -> rank table, seed slots: 20 bytes per candidate in one piece; the tail of the near-pair pool: 8
     * more), so that the per-cluster loops, which are serial in the members, read shared memory instead of waiting for
     * one L2 round trip per member. ---- */
    {
      const int mtot = (nseeds > 0) ? (int)csz[nseeds - 1] : 0;
      float* st_mx = reinterpret_cast<float*>(Gc);            /* five planes of S floats: Gc | P1 | items  (20 S + 128 bytes) */
      float* st_my = st_mx + S;
      float* st_w = st_my + S;
      float* st_c0 = st_w + S;
      float* st_c3 = st_c0 + S;
      float2* st_c12 = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(pool) + 2 * (size_t)S);   /* behind memb */
      for (int j = tid; j < mtot; j += MF_THREADS) {
        const int rr = memb[j];
        const float4 B0 = crec[2 * rr], B1 = crec[2 * rr + 1];
        st_mx[j] = B1.x; st_my[j] = B1.y; st_w[j] = B1.z;
        st_c0[j] = B0.x; st_c3[j] = B0.w;
        st_c12[j] = make_float2(B0.y, B0.z);
      }
      __syncthreads();
      for (int k = tid; k < klimit; k += MF_THREADS) {
        const int beg = (k > 0) ? csz[k - 1] : 0, end = csz[k];
        float wsum = 0.0f, m0 = 0.0f, m1 = 0.0f;
        for (int j = beg; j < end; ++j) {
          const float w = st_w[j];
          wsum = wsum + w;
          m0 = m0 + w * st_mx[j];
          m1 = m1 + w * st_my[j];
        }
        if (wsum == 0.0f) {                                 /* :2821-2822: the reference stops here */
          atomicMin(&s_kzero, k);
        } else if (k < Cmax) {
          const float rw = 1.0f / wsum;
          const float mm0 = m0 * rw, mm1 = m1 * rw;
          float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
          for (int j = beg; j < end; ++j) {
            const float w = st_w[j];
            const float2 c12 = st_c12[j];
            const float d0 = mm0 - st_mx[j], d1 = mm1 - st_my[j];
            v0 = v0 + w * (st_c0[j] + d0 * d0);
            v1 = v1 + w * (c12.x + d0 * d1);
            v2 = v2 + w * (c12.y + d1 * d0);
            v3 = v3 + w * (st_c3[j] + d1 * d1);
          }
          v0 = v0 * rw; v1 = v1 * rw; v2 = v2 * rw; v3 = v3 * rw;
          v1 = (v1 + v2) / 2.0f;                             /* force_symmetric_covariance */
          mo[0 * Cmax + k] = wsum;
          mo[1 * Cmax + k] = mm0;
          mo[2 * Cmax + k] = mm1;
          mo[3 * Cmax + k] = v0;
          mo[4 * Cmax + k] = v1;
          mo[5 * Cmax + k] = v3;
        }
      }
    }
    __syncthreads();
    nout = min(klimit, s_kzero);
    if (nout > Cmax && tid == 0) atomicOr(&a.red->err_flag, 2);
  }

  /* ---- I. re-append the far (class 0) components (:3311-3318) and publish the map size ---- */
  if (warp == 0) {
    int pos0 = nout;
    for (int b0 = 0; b0 < cnt; b0 += 32) {
      const int i = b0 + lane;
      const bool k0 = (i < cnt) && (cl[i] == 0);
      const unsigned bal = __ballot_sync(FULL_MASK, k0);
      if (k0) {
        const int pos = pos0 + __popc(bal & lt_mask);
        if (pos < Cmax) {
#pragma unroll
          for (int f = 0; f < PHD_MAP_PLANES; ++f) mo[f * Cmax + pos] = mp[f * Cmax + i];
        }
      }
      pos0 += __popc(bal);
    }
    if (lane == 0) {
      if (pos0 > Cmax) {
        atomicOr(&a.red->err_flag, 2);
        pos0 = Cmax;
      }
      a.count_out[pl] = pos0;
    }
  }
}

/* =========================================================================================== */
/* particle weights: w += dw; w -= logsumexp(w)  (reference src/phdfilter.cu:3735-3755)           */
/* and state extraction sums (src/main.cpp:324-356, 1281-1284).  All cross-particle sums are       */
/* integer (fixed point), hence independent of summation order and of the number of GPUs.          */
/* =========================================================================================== */
__global__ void weights_add_max_kernel(float* __restrict__ logw, const float* __restrict__ dlogw, int n, Reductions* red) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float w = -FLT_MAX;
  bool isnan_ = false;
  if (i < n) {
    w = logw[i];
    if (dlogw) {
      w = w + dlogw[i];
      logw[i] = w;
    }
    isnan_ = (w != w);
    if (isnan_) w = -FLT_MAX;
  }
  unsigned key = float_to_ordered_uint(w);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) key = max(key, __shfl_xor_sync(FULL_MASK, key, off));
  unsigned nanb = __ballot_sync(FULL_MASK, isnan_);
  if (lane_id() == 0) {
    atomicMax(&red->max_key, key);
    if (nanb) atomicAdd(&red->nan_count, (unsigned long long)__popc(nanb));
  }
}
__global__ void weights_sum_kernel(const float* __restrict__ logw, int n, Reductions* red) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float mx = ordered_uint_to_float(red->max_key);
  unsigned long long v = 0;
  if (i < n) {
    float w = logw[i];
    if (w == w) v = phd_fx_from_unit(phd_expf(w - mx), PHD_FX_WEIGHT_BITS);
  }
  v = warp_sum_u64(v);
  if (lane_id() == 0 && v) atomicAdd(&red->sum_fx, v);
  if (i == 0 && red->err_flag) atomicAdd(&red->err_ranks, 1ull);   /* travels with the sum: every rank learns of the error */
}
__global__ void weights_normalise_kernel(float* __restrict__ logw, int n, const Reductions* red) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float mx = ordered_uint_to_float(red->max_key);
  float sumf = (float)((double)red->sum_fx * (1.0 / (double)(1ull << PHD_FX_WEIGHT_BITS)));
  float lse = phd_safe_log(sumf) + mx;
  logw[i] = logw[i] - lse;
}
__global__ void estimate_kernel(const float* __restrict__ logw, const float* __restrict__ pose, int n, int offset, Reductions* red) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  long long acc[6] = {0, 0, 0, 0, 0, 0};
  unsigned long long e2 = 0, key = 0, cdf = 0;
  if (i < n) {
    float w = logw[i];
    if (w == w) {
      float ew = phd_expf(w);
#pragma unroll
      for (int k = 0; k < 6; ++k) acc[k] = phd_fx_from_prod(ew, pose[(size_t)k * n + i], PHD_FX_POSE_BITS);
      e2 = phd_fx_from_unit(phd_expf(2.0f * w), PHD_FX_NEFF_BITS);
      cdf = phd_fx_from_unit(ew, PHD_FX_CDF_BITS);        /* = resample_weights_kernel: the rank's resampling CDF total */
      if (w > -FLT_MAX)
        key = ((unsigned long long)float_to_ordered_uint(w) << 32) | (unsigned long long)(0xffffffffu - (unsigned)(offset + i));
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k] = warp_sum_i64(acc[k]);
  e2 = warp_sum_u64(e2);
  cdf = warp_sum_u64(cdf);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    unsigned long long o = __shfl_xor_sync(FULL_MASK, key, off);
    key = (o > key) ? o : key;
  }
  if (lane_id() == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (acc[k]) atomicAdd((unsigned long long*)&red->pose_fx[k], (unsigned long long)acc[k]);
    if (e2) atomicAdd(&red->neff_fx, e2);
    if (cdf) atomicAdd(&red->cdf_total, cdf);
    atomicMax(&red->argmax_key, key);
  }
}

/* =========================================================================================== */
/* resampling: resampleParticles + copy_particles (reference src/main.cpp:453-501,               */
/* src/slamtypes.h:313-333) on the canonical integer CDF                                          */
/* =========================================================================================== */
__global__ void resample_weights_kernel(const float* __restrict__ logw, int n, unsigned long long* __restrict__ q) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float w = logw[i];
  q[i] = (w == w) ? phd_fx_from_unit(phd_expf(w), PHD_FX_CDF_BITS) : 0ull;
}

/* ancestor(j) = min{ i : C_i > floor(r_j * total) }, C = inclusive CDF = excl[i+1] (+ cdf_base for this rank).
 * Offspring j (global) are [j0, j0+n_off); ancestors searched in the local CDF [0,n) shifted by cdf_base. */
/* ancestor of offspring j (global index) if it lives on this rank (local index), else -1 */
__device__ __forceinline__ int resample_ancestor(const unsigned long long* __restrict__ excl, int n, unsigned long long cdf_base,
                                                 unsigned long long total, int n_new, int j, const double* __restrict__ uniforms,
                                                 int systematic, unsigned call, uint32_t seed_lo, uint32_t seed_hi) {
  const double interval = 1.0 / (double)n_new;
  double u;
  if (uniforms) {
    u = systematic ? uniforms[0] : uniforms[1 + (size_t)j];
  } else {
    phd_philox4_t r = phd_philox4x32_10(systematic ? 0u : (uint32_t)j, call, PHD_STREAM_RESAMPLE, 0u, seed_lo, seed_hi);
    u = phd_u01d(r.v[0], r.v[1]);
  }
  double r = (double)j * interval + u * interval;
  double t = floor(r * (double)total);
  unsigned long long R = (t <= 0.0) ? 0ull : (unsigned long long)t;
  if (R >= total) R = total - 1;
  /* is the ancestor on this rank?  local inclusive CDF range is (cdf_base, cdf_base + excl[n]] */
  if (R < cdf_base || R >= cdf_base + excl[n]) return -1;
  unsigned long long Rl = R - cdf_base;
  int lo = 0, hi = n - 1; /* smallest i with excl[i+1] > Rl */
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (excl[mid + 1] > Rl) hi = mid; else lo = mid + 1;
  }
  return lo;
}
__global__ void resample_search_kernel(const unsigned long long* __restrict__ excl, int n, unsigned long long cdf_base,
                                       unsigned long long total, int n_new, int j0, int n_off, int anc_offset,
                                       const double* __restrict__ uniforms, int systematic, unsigned call, uint32_t seed_lo,
                                       uint32_t seed_hi, int* __restrict__ anc) {
  int jj = blockIdx.x * blockDim.x + threadIdx.x;
  if (jj >= n_off) return;
  const int a = resample_ancestor(excl, n, cdf_base, total, n_new, j0 + jj, uniforms, systematic, call, seed_lo, seed_hi);
  anc[jj] = (a < 0) ? -1 : anc_offset + a;
}

/* one warp per offspring: copy pose, map block (only `count` live components per plane), cardinality.
 * Local resampling: the destination is this GPU's back buffer (dst_first = 0, anc_out = nullptr).
 * NVLink exchange (world > 1, peer window mapped): the same kernel PUSHES the offspring a peer owns straight into that
 * peer's back buffer -- pose_out / count_out / map_out / card_out / anc_out are peer pointers, dst_first the first slot
 * of the interval, n_dst the peer's particle count -- so the gather and the transfer are one kernel and nothing is
 * packed, staged or unpacked (coalesced 128-byte stores over NVLink). */
__global__ void resample_gather_kernel(const int* __restrict__ anc, int n_off, int anc_offset, int n_src, int n_dst,
                                       const float* __restrict__ pose_in, float* __restrict__ pose_out,
                                       const int* __restrict__ count_in, int* __restrict__ count_out,
                                       const float* __restrict__ map_in, float* __restrict__ map_out,
                                       const float* __restrict__ card_in, float* __restrict__ card_out, int Cmax, int n_card,
                                       int dst_first, int* __restrict__ anc_out) {
  int j = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (j >= n_off) return;
  const int lane = lane_id();
  const int a = anc[j] - anc_offset;
  if (a < 0 || a >= n_src) return;
  const size_t jd = (size_t)dst_first + j;
  if (lane < 6) pose_out[(size_t)lane * n_dst + jd] = pose_in[(size_t)lane * n_src + a];
  const int cnt = count_in[a];
  if (lane == 0) {
    count_out[jd] = cnt;
    if (anc_out) anc_out[jd] = anc[j];
  }
  const float* src = map_in + (size_t)a * PHD_MAP_PLANES * Cmax;
  float* dst = map_out + jd * PHD_MAP_PLANES * Cmax;
  for (int f = 0; f < PHD_MAP_PLANES; ++f)
    for (int k = lane; k < cnt; k += 32) dst[f * Cmax + k] = src[f * Cmax + k];
  if (n_card > 0 && card_in)
    for (int k = lane; k < n_card; k += 32) card_out[jd * n_card + k] = card_in[(size_t)a * n_card + k];
}

/* NVLink exchange in ONE launch: the offspring intervals every peer owns whose ancestors live here (both ends derive them
 * from the ranks' CDF totals, nothing is negotiated).  One warp per offspring: ancestor search on the local CDF, then the
 * particle is written straight into the owner's back buffer through the mapped peer window. */
#define PHD_MAX_PEERS 8
struct PushDst {
  float* pose; int* count; float* map; float* card; int* anc_in;   /* the peer's back buffers (mapped) */
  int n_dst;       /* the peer's particle count */
  int dst_first;   /* first local slot (at the peer) of the interval */
  int j0;          /* first global offspring index of the interval */
  int first;       /* exclusive prefix of the interval sizes */
};
struct PushArgs {
  PushDst d[PHD_MAX_PEERS];
  int n_dst_entries, total;
  const unsigned long long* excl; int n; unsigned long long cdf_base, cdf_total; int n_new, anc_offset;
  const double* uniforms; int systematic; unsigned call; uint32_t seed_lo, seed_hi;
  const float* pose_in; const int* count_in; const float* map_in; const float* card_in; int Cmax, n_card;
};
__global__ void resample_push_kernel(PushArgs a) {
  const int jj = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (jj >= a.total) return;
  const int lane = lane_id();
  int e = 0;
#pragma unroll
  for (int k = 1; k < PHD_MAX_PEERS; ++k)
    if (k < a.n_dst_entries && jj >= a.d[k].first) e = k;
  const PushDst& D = a.d[e];
  const int off = jj - D.first;
  const int j = D.j0 + off;
  const int al = resample_ancestor(a.excl, a.n, a.cdf_base, a.cdf_total, a.n_new, j, a.uniforms, a.systematic, a.call, a.seed_lo,
                                   a.seed_hi);
  if (al < 0) return;     /* cannot happen: the interval was planned from the same totals */
  const size_t jd = (size_t)D.dst_first + off;
  if (lane < 6) D.pose[(size_t)lane * D.n_dst + jd] = a.pose_in[(size_t)lane * a.n + al];
  const int cnt = a.count_in[al];
  if (lane == 0) {
    D.count[jd] = cnt;
    D.anc_in[jd] = a.anc_offset + al;
  }
  const float* src = a.map_in + (size_t)al * PHD_MAP_PLANES * a.Cmax;
  float* dst = D.map + jd * PHD_MAP_PLANES * a.Cmax;
  for (int f = 0; f < PHD_MAP_PLANES; ++f)
    for (int k = lane; k < cnt; k += 32) dst[f * a.Cmax + k] = src[f * a.Cmax + k];
  if (a.n_card > 0 && a.card_in)
    for (int k = lane; k < a.n_card; k += 32) D.card[jd * a.n_card + k] = a.card_in[(size_t)al * a.n_card + k];
}

/* Mailbox all-gather over the NVLink peer window: every rank writes a record of up to MBOX_PAYLOAD 64-bit words into its
 * slot of EVERY rank's mailbox (payload, system fence, then the sequence number as the flag), waits until all ranks'
 * records of this sequence number have arrived in its own mailbox, and combines them in rank order (word by word: sum or
 * max) -- one small kernel per exchange, a few microseconds, instead of an NCCL collective per statistic.  Sequence
 * numbers advance in lock step on all ranks (the same call sequence), MBOX_SLOTS records per source are in flight at
 * most (a rank can be one exchange ahead of the slowest one).  A rank that never arrives (it left the step with an
 * error) is noticed after MBOX_TIMEOUT_NS: err_flag bit 4 instead of a hang. */
#define MBOX_SLOTS 4
#define MBOX_WORDS 16
#define MBOX_PAYLOAD 15
#define MBOX_TIMEOUT_NS 120000000000ull   /* 120 s: ranks may reach their first exchange seconds apart (scene loading) */
struct MboxPeers { unsigned long long* box[PHD_MAX_PEERS]; };   /* mailbox region of every rank (own and mapped) */
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__global__ void mbox_exchange_kernel(MboxPeers pp, int me, int W, unsigned long long seq, unsigned long long* __restrict__ words,
                                     int n_words, unsigned max_mask, unsigned long long* __restrict__ gathered /* [W][MBOX_WORDS] or null */,
                                     int* err_flag) {
  __shared__ unsigned long long s_in[PHD_MAX_PEERS][MBOX_WORDS];
  const int w = warp_id(), lane = lane_id();
  const size_t rec = ((size_t)(seq & (MBOX_SLOTS - 1)) * PHD_MAX_PEERS);
  if (w < W) {
    /* my record into rank w's mailbox */
    volatile unsigned long long* out = pp.box[w] + (rec + me) * MBOX_WORDS;
    if (lane < n_words) out[lane] = words[lane];
    __syncwarp();
    if (lane == 0) {
      __threadfence_system();
      out[MBOX_PAYLOAD] = seq;
    }
    /* rank w's record in my mailbox */
    volatile unsigned long long* in = pp.box[me] + (rec + w) * MBOX_WORDS;
    int ok = 1;
    if (lane == 0) {
      const unsigned long long t0 = global_timer_ns();
      while (in[MBOX_PAYLOAD] != seq) {
        if (global_timer_ns() - t0 > MBOX_TIMEOUT_NS) { ok = 0; break; }
      }
      __threadfence_system();
      if (!ok) atomicOr(err_flag, 4);
    }
    __syncwarp();
    if (lane < MBOX_WORDS) s_in[w][lane] = (lane < n_words) ? in[lane] : ((lane == MBOX_PAYLOAD) ? seq : 0ull);
  }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < n_words) {
    unsigned long long acc = s_in[0][t];
    for (int r = 1; r < W; ++r) {
      const unsigned long long v = s_in[r][t];
      acc = ((max_mask >> t) & 1u) ? (v > acc ? v : acc) : acc + v;
    }
    words[t] = acc;
  }
  if (gathered && t < W * MBOX_WORDS) gathered[t] = s_in[t / MBOX_WORDS][t % MBOX_WORDS];
}

/* after the exchange: offspring whose ancestor lives on a peer (-1 from the local search) take the index the peer pushed */
__global__ void resample_take_pushed_kernel(int* __restrict__ anc, const int* __restrict__ anc_in, int n) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n && anc[j] < 0) anc[j] = anc_in[j];
}

/* migration: pack the ancestors of `cnt` remote offspring into contiguous staging (one warp per record) */
__global__ void migrate_pack_kernel(const int* __restrict__ anc, int cnt, int anc_offset, int n_src,
                                    const float* __restrict__ pose_in, const int* __restrict__ count_in,
                                    const float* __restrict__ map_in, const float* __restrict__ card_in, int Cmax, int n_card,
                                    float* __restrict__ pose_out /* [cnt][6] */, int* __restrict__ count_out,
                                    float* __restrict__ map_out /* [cnt][6*Cmax] */, float* __restrict__ card_out) {
  int j = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (j >= cnt) return;
  const int lane = lane_id();
  const int a = anc[j] - anc_offset;
  if (lane < 6) pose_out[(size_t)j * 6 + lane] = pose_in[(size_t)lane * n_src + a];
  const int c = count_in[a];
  if (lane == 0) count_out[j] = c;
  const float* src = map_in + (size_t)a * PHD_MAP_PLANES * Cmax;
  float* dst = map_out + (size_t)j * PHD_MAP_PLANES * Cmax;
  for (int f = 0; f < PHD_MAP_PLANES; ++f)
    for (int k = lane; k < c; k += 32) dst[f * Cmax + k] = src[f * Cmax + k];
  if (n_card > 0 && card_in)
    for (int k = lane; k < n_card; k += 32) card_out[(size_t)j * n_card + k] = card_in[(size_t)a * n_card + k];
}
/* migration: received AoS poses -> SoA planes at offspring positions [first, first+cnt) */
__global__ void migrate_unpack_pose_kernel(const float* __restrict__ pose_in /* [cnt][6] */, int cnt, int first, int n_dst,
                                           float* __restrict__ pose_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cnt * 6) return;
  int j = i / 6, k = i - j * 6;
  pose_out[(size_t)k * n_dst + first + j] = pose_in[i];
}

/* "shotgun" prediction (reference src/phdfilter.cu:796-797, 1185-1238): prediction j descends from particle j / k */
__global__ void fanout_kernel(int n_out, int k, float logk, const float* __restrict__ logw_in, float* __restrict__ logw_out,
                              const int* __restrict__ ridx_in, int* __restrict__ ridx_out, int* __restrict__ anc) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  const int i = j / k;
  anc[j] = i;
  logw_out[j] = logw_in[i] - logk;
  ridx_out[j] = ridx_in[i];
}

/* One 64-bit checksum per particle over everything a resampled copy must carry (src/slamtypes.h:313-333: state, map,
 * cardinality): pose (6 words), map size, the size x 6 live map words and the cardinality row.  Every word is mixed with
 * its position (splitmix64 finaliser) and the mixes are summed, so the warp can hash in parallel and the result still
 * depends on the order of the words.  Used by bench.py's exchange_check and the multi-GPU tests: an offspring's checksum
 * must equal its ancestor's, whichever GPU the ancestor lived on. */
__device__ __forceinline__ unsigned long long checksum_mix(unsigned long long idx, unsigned word) {
  unsigned long long x = (idx + 1ull) * 0x9E3779B97F4A7C15ull ^ (unsigned long long)word;
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
__global__ void particle_checksum_kernel(const float* __restrict__ pose, const int* __restrict__ count, const float* __restrict__ map,
                                         const float* __restrict__ card, int n, int Cmax, int n_card,
                                         unsigned long long* __restrict__ out) {
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= n) return;
  const int lane = lane_id();
  const int cnt = count[p];
  unsigned long long acc = 0;
  if (lane < 6) acc += checksum_mix(lane, __float_as_uint(pose[(size_t)lane * n + p]));
  if (lane == 6) acc += checksum_mix(6, (unsigned)cnt);
  const float* m = map + (size_t)p * PHD_MAP_PLANES * Cmax;
  for (int k = 0; k < PHD_MAP_PLANES; ++k)
    for (int i = lane; i < cnt; i += 32) acc += checksum_mix(16ull + (unsigned long long)i * PHD_MAP_PLANES + k, __float_as_uint(m[(size_t)k * Cmax + i]));
  for (int i = lane; i < n_card; i += 32) acc += checksum_mix((1ull << 32) + i, __float_as_uint(card[(size_t)p * n_card + i]));
  acc = warp_sum_u64(acc);
  if (lane == 0) out[p] = acc;
}

/* ancestors[j] = j mod n_src; log-weights follow their source */
__global__ void tile_index_kernel(int* __restrict__ anc, int n, int n_src, float* logw) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  anc[j] = j % n_src;
  if (j >= n_src) logw[j] = logw[j % n_src];
}
__global__ void fill_kernel(float* p, int n, float v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void iota_kernel(int* p, int n, int base) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = base + i;
}

/* =========================================================================================== */
/* EAP map: computeExpectedMap (reference src/main.cpp:290-316) + reduceGaussianMixture            */
/* (src/gm_reduce.cpp:57-134, host Eigen code in the reference) on the device.                    */
/* All particles' components, weights scaled by exp(particle log-weight), are concatenated; each   */
/* round takes the heaviest remaining component (ties: lowest concat index), gathers every         */
/* remaining component within the Cholesky Mahalanobis distance and moment-matches the cluster.    */
/* Cluster sums are accumulated in double (order independent to < 1 fp32 ulp).                     */
/* =========================================================================================== */
struct EapAcc {                 /* per-round accumulators */
  unsigned long long key;       /* arg-max key: ordered(weight) << 32 | ~global concat index */
  double w, wx, wy, s0, s1, s2, s3;
  float seed[8];                /* seed record: c0..c3, x, y, w, valid */
  int n_out;
  int pad;
};

__global__ void eap_concat_kernel(const float* __restrict__ map, const int* __restrict__ count, const float* __restrict__ logw,
                                  const unsigned long long* __restrict__ off, int n, int Cmax, float4* __restrict__ rec) {
  int p = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (p >= n) return;
  const int lane = lane_id();
  const float ew = phd_expf(logw[p]);                 /* map[i].weight *= exp(weights[n]) (main.cpp:301-302) */
  const float* mp = map + (size_t)p * PHD_MAP_PLANES * Cmax;
  const int c = count[p];
  float4* o = rec + 2 * off[p];
  for (int i = lane; i < c; i += 32) {
    float pxy = mp[4 * Cmax + i];
    o[2 * i] = make_float4(mp[3 * Cmax + i], pxy, pxy, mp[5 * Cmax + i]);
    o[2 * i + 1] = make_float4(mp[1 * Cmax + i], mp[2 * Cmax + i], mp[0 * Cmax + i] * ew, 1.0f);
  }
}

__global__ void eap_argmax_kernel(const float4* __restrict__ rec, unsigned long long n_tot, unsigned long long gbase, EapAcc* acc) {
  unsigned long long key = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_tot;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    float4 r1 = rec[2 * i + 1];
    if (r1.w != 0.0f) {
      unsigned long long k = ((unsigned long long)float_to_ordered_uint(r1.z) << 32) | (unsigned long long)(0xffffffffu - (unsigned)(gbase + i));
      key = (k > key) ? k : key;
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    unsigned long long o = __shfl_xor_sync(FULL_MASK, key, off);
    key = (o > key) ? o : key;
  }
  if (lane_id() == 0 && key) atomicMax(&acc->key, key);
}

/* the rank that owns the global arg-max publishes the seed record (others contribute zeros to the sum) */
__global__ void eap_seed_kernel(const float4* __restrict__ rec, unsigned long long n_tot, unsigned long long gbase, EapAcc* acc) {
  if (threadIdx.x != 0) return;
  for (int k = 0; k < 8; ++k) acc->seed[k] = 0.0f;
  acc->w = acc->wx = acc->wy = acc->s0 = acc->s1 = acc->s2 = acc->s3 = 0.0;
  unsigned long long key = acc->key;
  if (!key) return;
  unsigned long long g = (unsigned long long)(0xffffffffu - (unsigned)(key & 0xffffffffull));
  if (g >= gbase && g < gbase + n_tot) {
    float4 r0 = rec[2 * (g - gbase)], r1 = rec[2 * (g - gbase) + 1];
    acc->seed[0] = r0.x; acc->seed[1] = r0.y; acc->seed[2] = r0.z; acc->seed[3] = r0.w;
    acc->seed[4] = r1.x; acc->seed[5] = r1.y; acc->seed[6] = r1.z; acc->seed[7] = 1.0f;
  }
}

/* mahalanobisDistance of gm_reduce.cpp:31-38: L L^T = (Pa+Pb)/2, x = L^-1 d, |x|^2 (oracle: mahal_llt) */
__device__ __forceinline__ float dev_mahal_llt(const float* a, float4 b0, float4 b1) {
  float s00 = 0.5f * (a[0] + b0.x);
  float s10 = 0.5f * (a[1] + b0.y);
  float s11 = 0.5f * (a[3] + b0.w);
  float d0 = a[4] - b1.x, d1 = a[5] - b1.y;
  float l00 = sqrtf(s00);
  float l10 = s10 / l00;
  float l11 = sqrtf(s11 - l10 * l10);
  float x0 = d0 / l00;
  float x1 = (d1 - l10 * x0) / l11;
  return x0 * x0 + x1 * x1;
}

__global__ void eap_cluster_kernel(float4* __restrict__ rec, unsigned long long n_tot, unsigned long long gbase, float min_distance,
                                   EapAcc* acc) {
  __shared__ float sd[8];
  if (threadIdx.x < 8) sd[threadIdx.x] = acc->seed[threadIdx.x];
  __syncthreads();
  const unsigned long long key = acc->key;
  if (!key) return;
  const unsigned long long gseed = (unsigned long long)(0xffffffffu - (unsigned)(key & 0xffffffffull));
  double w = 0, wx = 0, wy = 0, s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_tot;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    float4 r1 = rec[2 * i + 1];
    if (r1.w == 0.0f) continue;
    float4 r0 = rec[2 * i];
    bool memb = (gbase + i == gseed) || (dev_mahal_llt(sd, r0, r1) < min_distance);
    if (memb) {
      double ww = (double)r1.z, x = (double)r1.x, y = (double)r1.y;
      w += ww; wx += ww * x; wy += ww * y;
      s0 += ww * ((double)r0.x + x * x); s1 += ww * ((double)r0.y + y * x);
      s2 += ww * ((double)r0.z + x * y); s3 += ww * ((double)r0.w + y * y);
      r1.w = 0.0f;
      rec[2 * i + 1] = r1;
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    w += __shfl_xor_sync(FULL_MASK, w, off); wx += __shfl_xor_sync(FULL_MASK, wx, off); wy += __shfl_xor_sync(FULL_MASK, wy, off);
    s0 += __shfl_xor_sync(FULL_MASK, s0, off); s1 += __shfl_xor_sync(FULL_MASK, s1, off);
    s2 += __shfl_xor_sync(FULL_MASK, s2, off); s3 += __shfl_xor_sync(FULL_MASK, s3, off);
  }
  if (lane_id() == 0 && w != 0.0) {
    atomicAdd(&acc->w, w); atomicAdd(&acc->wx, wx); atomicAdd(&acc->wy, wy);
    atomicAdd(&acc->s0, s0); atomicAdd(&acc->s1, s1); atomicAdd(&acc->s2, s2); atomicAdd(&acc->s3, s3);
  }
}

/* moment-matched merge (gm_reduce.cpp:104-130): mean = sum(w mu)/W, cov = sum w (P + (m-mu)(m-mu)^T)/W */
__global__ void eap_finalize_kernel(EapAcc* acc, phdslam_gaussian2d_t* __restrict__ out, int cap, unsigned long long* key_out) {
  if (threadIdx.x != 0) return;
  *key_out = acc->key;
  if (acc->key && acc->w != 0.0) {
    double W = acc->w, mx = acc->wx / W, my = acc->wy / W;
    phdslam_gaussian2d_t g;
    g.weight = (float)W;
    g.mean[0] = (float)mx; g.mean[1] = (float)my;
    g.cov[0] = (float)(acc->s0 / W - mx * mx); g.cov[1] = (float)(acc->s1 / W - my * mx);
    g.cov[2] = (float)(acc->s2 / W - mx * my); g.cov[3] = (float)(acc->s3 / W - my * my);
    if (acc->n_out < cap) out[acc->n_out] = g;
    acc->n_out++;
  }
  acc->key = 0;
}

/* blocked dense planes -> reference AoS order (tests / phdslam_update_terms export only) */
__global__ void dense_export_kernel(const float* __restrict__ dense, const unsigned long long* __restrict__ toff,
                                    unsigned long long tbase, const int* __restrict__ n_in, int M, int p0, int np,
                                    const unsigned long long* __restrict__ out_off, phdslam_gaussian2d_t* __restrict__ out) {
  int pl = p0 + blockIdx.x;
  int C = n_in[pl];
  unsigned long long T = (unsigned long long)C * (unsigned)(M + 1) + (unsigned)M;
  const float* D = dense + (toff[pl] - tbase) * PHD_NPLANES;
  phdslam_gaussian2d_t* o = out + out_off[pl];
  for (unsigned long long t = threadIdx.x; t < T; t += blockDim.x) {
    const float* q = D + dense_index((size_t)t);
    phdslam_gaussian2d_t g;
    g.cov[0] = q[0]; g.cov[1] = q[64]; g.cov[2] = q[128]; g.cov[3] = q[192];
    g.mean[0] = q[256]; g.mean[1] = q[320]; g.weight = q[384];
    o[t] = g;
  }
  (void)np;
}

#endif /* PHD_KERNELS_CUH */
