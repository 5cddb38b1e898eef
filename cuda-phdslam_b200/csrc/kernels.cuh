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
 *   dense  per particle p a block of 7 planes {c0,c1,c2,c3,mx,my,w} of Tpad_p terms each, at float
 *          offset 7*toff[p]; term order inside a plane is the reference's features_update order
 *          [non-detect C | detect m-major M*C | birth M] (src/phdfilter.cu:2123-2124,2137-2166)
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

/* canonical warp_sum over a shared/global array (oracle: warp_sum) -- must be called by a full warp */
__device__ __forceinline__ float warp_sum_array(const float* v, int n) {
  float acc = 0.0f;
  for (int i = lane_id(); i < n; i += 32) acc = acc + v[i];
  return warp_butterfly_sum(acc);
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
    t = (t + 7ull) & ~7ull; /* planes start on 32-byte sectors */
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
/* GM-PHD update, dense output: preUpdateSynthKernel + phdUpdateKernel + host birth loop        */
/* (reference src/phdfilter.cu:1824-1925, 2083-2321, 3468-3507) fused into one pass.             */
/*                                                                                               */
/* One CTA (8 warps) per particle.  The particle's in-range components are compacted into shared */
/* memory, the per-component EKF constants are computed once, then warp w owns measurements      */
/* m = w, w+8, ...: lanes stride the components, the per-measurement normaliser is a lane-strided */
/* partial sum + xor butterfly (the canonical reduction the oracle mirrors), and the normalised   */
/* terms leave as 128-byte coalesced streaming stores, one plane at a time.                      */
/* Algorithmic traffic per particle: read 24*C + 32 B, write 28*(C*(M+1)+M) + 4 B.               */
/* =========================================================================================== */
#define UPD_THREADS 256
#define UPD_WARPS (UPD_THREADS / 32)
#define UPD_FLOATS_PER_COMP 23

struct UpdArgs {
  const float* map; const int* count; const uint8_t* cls; const float* pose;
  const float* z;                      /* [3][PHD_MAX_MEAS] range, bearing, label */
  int M, n, p0;
  const unsigned long long* toff;      /* exclusive scan of padded term counts (global over local particles) */
  unsigned long long tbase;            /* toff of the first particle of this batch */
  float* dense;
  const int* n_in;
  float* dlogw;
  DevCfg c;
};

static inline size_t update_smem_bytes(int Cmax) {
  return ((size_t)UPD_FLOATS_PER_COMP * Cmax + 6 * PHD_MAX_MEAS) * sizeof(float);
}

__global__ void __launch_bounds__(UPD_THREADS) update_dense_kernel(UpdArgs a) {
  extern __shared__ float smem[];
  const DevCfg& c = a.c;
  const int Cmax = c.Cmax;
  float* s_w = smem;
  float* s_mx = s_w + Cmax;
  float* s_my = s_mx + Cmax;
  float* s_pxx = s_my + Cmax;
  float* s_pxy = s_pxx + Cmax;
  float* s_pyy = s_pxy + Cmax;
  float* s_r = s_pyy + Cmax;
  float* s_b = s_r + Cmax;
  float* s_K0 = s_b + Cmax;
  float* s_K1 = s_K0 + Cmax;
  float* s_K2 = s_K1 + Cmax;
  float* s_K3 = s_K2 + Cmax;
  float* s_S0 = s_K3 + Cmax;
  float* s_S12 = s_S0 + Cmax;
  float* s_S3 = s_S12 + Cmax;
  float* s_base = s_S3 + Cmax;
  float* s_hl = s_base + Cmax;
  float* s_cu0 = s_hl + Cmax;
  float* s_cu1 = s_cu0 + Cmax;
  float* s_cu2 = s_cu1 + Cmax;
  float* s_cu3 = s_cu2 + Cmax;
  float* s_nd = s_cu3 + Cmax;           /* non-detection weights (scheme-1 particle weighting) */
  float* s_tmp = s_nd + Cmax;           /* Cmax + 256 : pd*w, then one birth weight per measurement */
  float* s_zr = s_tmp + Cmax + PHD_MAX_MEAS;
  float* s_zb = s_zr + PHD_MAX_MEAS;
  float* s_zl = s_zb + PHD_MAX_MEAS;
  float* s_L = s_zl + PHD_MAX_MEAS;
  float* s_ds = s_L + PHD_MAX_MEAS;
  __shared__ int s_wcnt[UPD_WARPS];

  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int pl = a.p0 + blockIdx.x;     /* local particle index */
  const int M = a.M;
  const int n = a.n;
  const int cnt = a.count[pl];
  const float px = a.pose[0 * (size_t)n + pl], py = a.pose[1 * (size_t)n + pl], pth = a.pose[2 * (size_t)n + pl];
  const float* mp = a.map + (size_t)pl * PHD_MAP_PLANES * Cmax;
  const uint8_t* cl = a.cls + (size_t)pl * Cmax;

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
      int pos = C + woff + __popc(bal & ((1u << lane) - 1u));
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

  const unsigned long long T = (unsigned long long)C * (unsigned)(M + 1) + (unsigned)M;
  const unsigned long long Tpad = (T + 7ull) & ~7ull;
  float* D = a.dense + (a.toff[pl] - a.tbase) * PHD_NPLANES;
  float* D0 = D;
  float* D1 = D + Tpad;
  float* D2 = D + 2 * Tpad;
  float* D3 = D + 3 * Tpad;
  float* D4 = D + 4 * Tpad;
  float* D5 = D + 5 * Tpad;
  float* D6 = D + 6 * Tpad;

  /* ---- phase 1: per-component EKF constants (preUpdateSynthKernel :1835-1894) + non-detection terms ---- */
  for (int j = tid; j < C; j += UPD_THREADS) {
    const float w = s_w[j], fx = s_mx[j], fy = s_my[j];
    const float P0 = s_pxx[j], P1 = s_pxy[j], P2 = s_pxy[j], P3 = s_pyy[j];
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
    s_r[j] = r; s_b[j] = bearing;
    s_K0[j] = K0; s_K1[j] = K1; s_K2[j] = K2; s_K3[j] = K3;
    s_S0[j] = S0; s_S12[j] = S1 + S2; s_S3[j] = S3;
    s_base[j] = phd_safe_log(pd) + phd_safe_log(w);
    s_hl[j] = 0.5f * phd_safe_log(det);
    s_cu0[j] = cu0; s_cu1[j] = cu1; s_cu2[j] = cu2; s_cu3[j] = cu3;
    s_tmp[j] = pd * w;
    /* non-detection term (:2137-2141) */
    float wnd = w * (1.0f - pd);
    s_nd[j] = wnd;
    st_stream(D0 + j, P0); st_stream(D1 + j, P1); st_stream(D2 + j, P2); st_stream(D3 + j, P3);
    st_stream(D4 + j, fx); st_stream(D5 + j, fy); st_stream(D6 + j, wnd);
  }
  for (int m = tid; m < M; m += UPD_THREADS) s_tmp[C + m] = c.birth_weight;
  __syncthreads();

  /* predicted cardinality (:2133-2186) and, for scheme 1, the prior / non-detect weight sums */
  float card_predict = 0.0f, cn_predict = 0.0f, nd_sum = 0.0f;
  if (warp == 0) {
    card_predict = warp_sum_array(s_tmp, C + M);
    if (c.particle_weighting == 1) {
      cn_predict = warp_sum_array(s_w, C);
      nd_sum = warp_sum_array(s_nd, C);
    }
  }

  /* ---- phase 2: detection terms; warp w owns measurements w, w+8, ... (:1898-1923, :2190-2252) ---- */
  for (int m = warp; m < M; m += UPD_WARPS) {
    const float zr = s_zr[m], zb = s_zb[m];
    const bool dead = c.labeled && (s_zl[m] != 0.0f);
    float acc = 0.0f;
    for (int j = lane; j < C; j += 32) {
      float i0 = zr - s_r[j];
      float i1 = phd_wrap_angle(zb - s_b[j]);
      float dist = i0 * i0 * s_S0[j] + i0 * i1 * s_S12[j] + i1 * i1 * s_S3[j];
      float g = -0.5f * dist - PHD_LOG_2PI_F - s_hl[j];
      float lw = dead ? PHD_LOG0 : (s_base[j] + g);
      acc = acc + phd_expf(lw);
    }
    float sum = warp_butterfly_sum(acc);
    sum = sum + c.clutter_density;
    sum = sum + c.birth_weight;
    const float L = phd_safe_log(sum);
    float* d0 = D0 + C + (size_t)m * C;
    float* d1 = D1 + C + (size_t)m * C;
    float* d2 = D2 + C + (size_t)m * C;
    float* d3 = D3 + C + (size_t)m * C;
    float* d4 = D4 + C + (size_t)m * C;
    float* d5 = D5 + C + (size_t)m * C;
    float* d6 = D6 + C + (size_t)m * C;
    float wacc = 0.0f;
    for (int j = lane; j < C; j += 32) {
      float i0 = zr - s_r[j];
      float i1 = phd_wrap_angle(zb - s_b[j]);
      float dist = i0 * i0 * s_S0[j] + i0 * i1 * s_S12[j] + i1 * i1 * s_S3[j];
      float g = -0.5f * dist - PHD_LOG_2PI_F - s_hl[j];
      float lw = dead ? PHD_LOG0 : (s_base[j] + g);
      float wt = phd_expf(lw - L);
      float m0 = s_mx[j] + s_K0[j] * i0 + s_K2[j] * i1;
      float m1 = s_my[j] + s_K1[j] * i0 + s_K3[j] * i1;
      st_stream(d0 + j, s_cu0[j]); st_stream(d1 + j, s_cu1[j]); st_stream(d2 + j, s_cu2[j]); st_stream(d3 + j, s_cu3[j]);
      st_stream(d4 + j, m0); st_stream(d5 + j, m1); st_stream(d6 + j, wt);
      wacc = wacc + wt;
    }
    float dsum = warp_butterfly_sum(wacc);
    /* birth term of measurement m (host loop :3468-3507, normalised at :2232-2242) */
    if (lane == 0) {
      float theta = pth + zb;
      float sn, cs;
      phd_sincosf(theta, &sn, &cs);
      float bdx = zr * cs;
      float bdy = zr * sn;
      float J0 = bdx / zr, J1 = bdy / zr, J2 = -bdy, J3 = bdx;
      float b0 = J0 * J0 * c.bvar_r + J2 * J2 * c.bvar_b;
      float b1 = J0 * J1 * c.bvar_r + J2 * J3 * c.bvar_b;
      float b3 = J1 * J1 * c.bvar_r + J3 * J3 * c.bvar_b;
      float lb = dead ? PHD_LOG0 : c.log_birth_weight;
      float wb = phd_expf(lb - L);
      size_t t = (size_t)C + (size_t)M * C + m;
      st_stream(D0 + t, b0); st_stream(D1 + t, b1); st_stream(D2 + t, b1); st_stream(D3 + t, b3);
      st_stream(D4 + t, px + bdx); st_stream(D5 + t, py + bdy); st_stream(D6 + t, wb);
      s_L[m] = L;
      s_ds[m] = dsum + wb;
    }
  }
  __syncthreads();

  /* ---- phase 3: particle log-weight increment (:2258-2279) ---- */
  if (tid == 0) {
    float pw = 0.0f;
    for (int m = 0; m < M; ++m) pw = pw + s_L[m];
    float out;
    if (c.particle_weighting == 0) {
      out = pw - card_predict;
    } else if (c.particle_weighting == 1) {
      float cn_update = nd_sum;
      for (int m = 0; m < M; ++m) cn_update = cn_update + s_ds[m];
      out = (float)M * c.clutter_density + cn_update - cn_predict - c.clutter_rate;
    } else {
      out = 0.0f;
    }
    a.dlogw[pl] = out;
  }
}

/* =========================================================================================== */
/* prune + merge: pruneMap + mergeAndCopyMaps + phdUpdateMergeKernel                             */
/* (reference src/phdfilter.cu:3120-3333, 2707-2898; computeMahalDist device_math.cuh:309-325)    */
/*                                                                                               */
/* One CTA (4 warps) per particle.  Survivors of the prune (weight >= minFeatureWeight) are       */
/* compacted from the dense weight plane into shared memory in term order, followed by the        */
/* "nearly in range" (class 2) components.  Candidates are ranked once by (weight desc, index      */
/* asc) with a bitonic sort -- weights of unmerged candidates never change, so the greedy arg-max  */
/* of every round is the next unmerged entry of that ranking.  Each round: all threads evaluate    */
/* the distance of their candidates to the seed and publish membership as ballot words; warp 0     */
/* then accumulates the moment-matched merge over the members in ascending index order (the        */
/* canonical order of the oracle).                                                                */
/* =========================================================================================== */
#define MRG_THREADS 128
#define MRG_WARPS (MRG_THREADS / 32)

struct MrgArgs {
  const float* dense; const unsigned long long* toff; unsigned long long tbase;
  const int* n_in; int M, n, p0;
  const float* map_in; const int* count_in; const uint8_t* cls;
  float* map_out; int* count_out;
  Reductions* red;
  int Smax;
  DevCfg c;
};

static inline size_t merge_smem_bytes(int Smax) {
  /* 7 candidate planes + 64-bit sort keys + membership/alive words */
  return (size_t)Smax * (7 * 4 + 8) + (size_t)(Smax / 32 + 1) * 2 * 4 + 64;
}

__device__ __forceinline__ float dev_mahal(float ac0, float ac1, float ac2, float ac3, float am0, float am1,
                                           float bc0, float bc1, float bc2, float bc3, float bm0, float bm1) {
  float s0 = (ac0 + bc0) / 2.0f, s1 = (ac1 + bc1) / 2.0f, s2 = (ac2 + bc2) / 2.0f, s3 = (ac3 + bc3) / 2.0f;
  float det = s0 * s3 - s2 * s1;
  float v0 = s3 / det, v1 = -s1 / det, v2 = -s2 / det, v3 = s0 / det;
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

__global__ void __launch_bounds__(MRG_THREADS) merge_kernel(MrgArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DevCfg& c = a.c;
  const int Smax = a.Smax;
  unsigned long long* s_key = (unsigned long long*)smem_raw;          /* Smax */
  float* s_c0 = (float*)(s_key + Smax);
  float* s_c1 = s_c0 + Smax;
  float* s_c2 = s_c1 + Smax;
  float* s_c3 = s_c2 + Smax;
  float* s_m0 = s_c3 + Smax;
  float* s_m1 = s_m0 + Smax;
  float* s_wt = s_m1 + Smax;
  unsigned* s_alive = (unsigned*)(s_wt + Smax);                       /* Smax/32+1 words */
  unsigned* s_memb = s_alive + (Smax / 32 + 1);
  __shared__ int s_wcnt[MRG_WARPS];
  __shared__ int s_n, s_seed, s_pos, s_stop, s_nout;

  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  const int pl = a.p0 + blockIdx.x;
  const int Cmax = c.Cmax;
  const int M = a.M;
  const int C = a.n_in[pl];
  const unsigned long long T = (unsigned long long)C * (unsigned)(M + 1) + (unsigned)M;
  const unsigned long long Tpad = (T + 7ull) & ~7ull;
  const float* D = a.dense + (a.toff[pl] - a.tbase) * PHD_NPLANES;
  const float* Dw = D + 6 * Tpad;
  const int cnt = a.count_in[pl];
  const float* mp = a.map_in + (size_t)pl * PHD_MAP_PLANES * Cmax;
  const uint8_t* cl = a.cls + (size_t)pl * Cmax;
  float* mo = a.map_out + (size_t)pl * PHD_MAP_PLANES * Cmax;

  /* ---- 1. prune: stable compaction of surviving dense terms (flags :2308-2319, pruneMap :3120-3174) ---- */
  /* pass A: each warp counts the survivors of its contiguous quarter of the term range */
  const int Ti = (int)T;
  const int seg = ((Ti + MRG_WARPS * 32 - 1) / (MRG_WARPS * 32)) * 32; /* per-warp segment, multiple of 32 */
  const int t_lo = warp * seg, t_hi = min(Ti, t_lo + seg);
  int mycount = 0;
  for (int t = t_lo + lane; t < t_lo + seg; t += 32) {
    bool keep = (t < t_hi) && !(Dw[t] < c.min_w);
    mycount += __popc(__ballot_sync(FULL_MASK, keep));
  }
  if (lane == 0) s_wcnt[warp] = mycount;
  __syncthreads();
  int woff = 0, nsurv = 0;
#pragma unroll
  for (int w = 0; w < MRG_WARPS; ++w) {
    if (w < warp) woff += s_wcnt[w];
    nsurv += s_wcnt[w];
  }
  /* pass B: place */
  int run = woff;
  for (int t = t_lo + lane; t < t_lo + seg; t += 32) {
    float wv = (t < t_hi) ? Dw[t] : 0.0f;
    bool keep = (t < t_hi) && !(wv < c.min_w);
    unsigned bal = __ballot_sync(FULL_MASK, keep);
    if (keep) {
      int pos = run + __popc(bal & ((1u << lane) - 1u));
      if (pos < Smax) {
        s_c0[pos] = D[t]; s_c1[pos] = D[Tpad + t]; s_c2[pos] = D[2 * Tpad + t]; s_c3[pos] = D[3 * Tpad + t];
        s_m0[pos] = D[4 * Tpad + t]; s_m1[pos] = D[5 * Tpad + t]; s_wt[pos] = wv;
      }
    }
    run += __popc(bal);
  }
  __syncthreads();
  /* ---- 2. append the nearly-in-range components in map order (:3243-3252) ---- */
  if (warp == 0) {
    int pos0 = nsurv;
    for (int base = 0; base < cnt; base += 32) {
      int i = base + lane;
      bool k2 = (i < cnt) && (cl[i] == 2);
      unsigned bal = __ballot_sync(FULL_MASK, k2);
      if (k2) {
        int pos = pos0 + __popc(bal & ((1u << lane) - 1u));
        if (pos < Smax) {
          s_wt[pos] = mp[0 * Cmax + i]; s_m0[pos] = mp[1 * Cmax + i]; s_m1[pos] = mp[2 * Cmax + i];
          s_c0[pos] = mp[3 * Cmax + i]; s_c1[pos] = mp[4 * Cmax + i]; s_c2[pos] = mp[4 * Cmax + i]; s_c3[pos] = mp[5 * Cmax + i];
        }
      }
      pos0 += __popc(bal);
    }
    if (lane == 0) {
      if (pos0 > Smax) {
        atomicOr(&a.red->err_flag, 1);
        pos0 = Smax;
      }
      s_n = pos0;
    }
  }
  __syncthreads();
  const int n = s_n;

  /* ---- 3. rank candidates: key = (~ordered(weight) << 32) | index, ascending bitonic sort ---- */
  int npow = 32;
  while (npow < n) npow <<= 1;
  for (int i = tid; i < npow; i += MRG_THREADS) {
    unsigned long long key = ~0ull;
    if (i < n) key = ((unsigned long long)(~float_to_ordered_uint(s_wt[i])) << 32) | (unsigned)i;
    s_key[i] = key;
  }
  for (int i = tid; i < (n + 31) / 32 + 1; i += MRG_THREADS) {
    int lo = i * 32;
    unsigned w = 0;
    if (lo < n) w = (n - lo >= 32) ? 0xffffffffu : ((1u << (n - lo)) - 1u);
    s_alive[i] = w;
    s_memb[i] = 0;
  }
  __syncthreads();
  for (int k = 2; k <= npow; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow; i += MRG_THREADS) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long ka = s_key[i], kb = s_key[ixj];
          bool up = ((i & k) == 0);
          if ((ka > kb) == up) {
            s_key[i] = kb;
            s_key[ixj] = ka;
          }
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    s_pos = 0;
    s_nout = 0;
    s_stop = 0;
    s_seed = (n > 0) ? (int)(s_key[0] & 0xffffffffu) : -1;
  }
  __syncthreads();

  /* ---- 4. greedy merge rounds (:2739-2894) ---- */
  const int nwords = (n + 31) / 32;
  while (true) {
    const int seed = s_seed;
    if (seed < 0 || s_stop) break;
    const float ac0 = s_c0[seed], ac1 = s_c1[seed], ac2 = s_c2[seed], ac3 = s_c3[seed];
    const float am0 = s_m0[seed], am1 = s_m1[seed];
    /* membership: warp w evaluates 32-candidate words w, w+4, ... */
    for (int wd = warp; wd < nwords; wd += MRG_WARPS) {
      unsigned alive = s_alive[wd];
      int i = wd * 32 + lane;
      bool memb = false;
      if ((alive >> lane) & 1u) {
        float dist = (c.distance_metric == 0)
                         ? dev_mahal(ac0, ac1, ac2, ac3, am0, am1, s_c0[i], s_c1[i], s_c2[i], s_c3[i], s_m0[i], s_m1[i])
                         : dev_hellinger(ac0, ac1, ac2, ac3, am0, am1, s_c0[i], s_c1[i], s_c2[i], s_c3[i], s_m0[i], s_m1[i]);
        memb = dist < c.min_sep;
      }
      unsigned bal = __ballot_sync(FULL_MASK, memb);
      if (lane == 0) s_memb[wd] = bal;
    }
    __syncthreads();
    if (warp == 0) {
      /* moment-matched merge over members in ascending index order; every lane computes the same values */
      float wsum = 0.0f, m0 = 0.0f, m1 = 0.0f;
      for (int wd = 0; wd < nwords; ++wd) {
        unsigned mb = s_memb[wd];
        while (mb) {
          int i = wd * 32 + (__ffs(mb) - 1);
          mb &= mb - 1;
          float wi = s_wt[i];
          wsum = wsum + wi;
          m0 = m0 + wi * s_m0[i];
          m1 = m1 + wi * s_m1[i];
        }
      }
      if (wsum == 0.0f) {
        if (lane == 0) s_stop = 1;
      } else {
        float mm0 = m0 / wsum, mm1 = m1 / wsum;
        float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f, v3 = 0.0f;
        for (int wd = 0; wd < nwords; ++wd) {
          unsigned mb = s_memb[wd];
          while (mb) {
            int i = wd * 32 + (__ffs(mb) - 1);
            mb &= mb - 1;
            float wi = s_wt[i];
            float d0 = mm0 - s_m0[i], d1 = mm1 - s_m1[i];
            v0 = v0 + wi * (s_c0[i] + d0 * d0);
            v1 = v1 + wi * (s_c1[i] + d0 * d1);
            v2 = v2 + wi * (s_c2[i] + d1 * d0);
            v3 = v3 + wi * (s_c3[i] + d1 * d1);
          }
        }
        v0 = v0 / wsum; v1 = v1 / wsum; v2 = v2 / wsum; v3 = v3 / wsum;
        v1 = (v1 + v2) / 2.0f;                       /* force_symmetric_covariance */
        const int slot = s_nout;
        if (lane == 0) {
          if (slot < Cmax) {
            mo[0 * Cmax + slot] = wsum; mo[1 * Cmax + slot] = mm0; mo[2 * Cmax + slot] = mm1;
            mo[3 * Cmax + slot] = v0; mo[4 * Cmax + slot] = v1; mo[5 * Cmax + slot] = v3;
          } else {
            atomicOr(&a.red->err_flag, 2);
          }
        }
        /* retire members */
        for (int wd = lane; wd < nwords; wd += 32) s_alive[wd] &= ~s_memb[wd];
        __syncwarp();
        /* next seed: first still-alive entry of the ranking */
        int pos = s_pos;
        int next = -1;
        while (pos < n) {
          int cand = (int)(s_key[pos] & 0xffffffffu);
          if ((s_alive[cand >> 5] >> (cand & 31)) & 1u) {
            next = cand;
            break;
          }
          ++pos;
        }
        if (lane == 0) {
          s_pos = pos;
          s_seed = next;
          s_nout = slot + 1;
        }
      }
    }
    __syncthreads();
  }

  /* ---- 5. re-append the far (class 0) components (:3311-3318) and publish the map size ---- */
  if (warp == 0) {
    int pos0 = s_nout;
    for (int base = 0; base < cnt; base += 32) {
      int i = base + lane;
      bool k0 = (i < cnt) && (cl[i] == 0);
      unsigned bal = __ballot_sync(FULL_MASK, k0);
      if (k0) {
        int pos = pos0 + __popc(bal & ((1u << lane) - 1u));
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
    if (nanb) atomicAdd(&red->nan_count, __popc(nanb));
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
  unsigned long long e2 = 0, key = 0;
  if (i < n) {
    float w = logw[i];
    if (w == w) {
      float ew = phd_expf(w);
#pragma unroll
      for (int k = 0; k < 6; ++k) acc[k] = phd_fx_from_prod(ew, pose[(size_t)k * n + i], PHD_FX_POSE_BITS);
      e2 = phd_fx_from_unit(phd_expf(2.0f * w), PHD_FX_NEFF_BITS);
      if (w > -FLT_MAX)
        key = ((unsigned long long)float_to_ordered_uint(w) << 32) | (unsigned long long)(0xffffffffu - (unsigned)(offset + i));
    }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) acc[k] = warp_sum_i64(acc[k]);
  e2 = warp_sum_u64(e2);
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
__global__ void resample_search_kernel(const unsigned long long* __restrict__ excl, int n, unsigned long long cdf_base,
                                       unsigned long long total, int n_new, int j0, int n_off, int anc_offset,
                                       const double* __restrict__ uniforms, int systematic, unsigned call, uint32_t seed_lo,
                                       uint32_t seed_hi, int* __restrict__ anc) {
  int jj = blockIdx.x * blockDim.x + threadIdx.x;
  if (jj >= n_off) return;
  int j = j0 + jj;
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
  if (R < cdf_base || R >= cdf_base + excl[n]) {
    anc[jj] = -1;
    return;
  }
  unsigned long long Rl = R - cdf_base;
  int lo = 0, hi = n - 1; /* smallest i with excl[i+1] > Rl */
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (excl[mid + 1] > Rl) hi = mid; else lo = mid + 1;
  }
  anc[jj] = anc_offset + lo;
}

/* one warp per offspring: copy pose, map block (only `count` live components per plane), cardinality */
__global__ void resample_gather_kernel(const int* __restrict__ anc, int n_off, int anc_offset, int n_src, int n_dst,
                                       const float* __restrict__ pose_in, float* __restrict__ pose_out,
                                       const int* __restrict__ count_in, int* __restrict__ count_out,
                                       const float* __restrict__ map_in, float* __restrict__ map_out,
                                       const float* __restrict__ card_in, float* __restrict__ card_out, int Cmax, int n_card) {
  int j = blockIdx.x * (blockDim.x >> 5) + warp_id();
  if (j >= n_off) return;
  const int lane = lane_id();
  const int a = anc[j] - anc_offset;
  if (a < 0 || a >= n_src) return;
  if (lane < 6) pose_out[(size_t)lane * n_dst + j] = pose_in[(size_t)lane * n_src + a];
  const int cnt = count_in[a];
  if (lane == 0) count_out[j] = cnt;
  const float* src = map_in + (size_t)a * PHD_MAP_PLANES * Cmax;
  float* dst = map_out + (size_t)j * PHD_MAP_PLANES * Cmax;
  for (int f = 0; f < PHD_MAP_PLANES; ++f)
    for (int k = lane; k < cnt; k += 32) dst[f * Cmax + k] = src[f * Cmax + k];
  if (n_card > 0 && card_in)
    for (int k = lane; k < n_card; k += 32) card_out[(size_t)j * n_card + k] = card_in[(size_t)a * n_card + k];
}

__global__ void fill_kernel(float* p, int n, float v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void iota_kernel(int* p, int n, int base) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = base + i;
}

/* dense planes -> reference AoS order (tests / phdslam_update_terms export only) */
__global__ void dense_export_kernel(const float* __restrict__ dense, const unsigned long long* __restrict__ toff,
                                    unsigned long long tbase, const int* __restrict__ n_in, int M, int p0, int np,
                                    const unsigned long long* __restrict__ out_off, phdslam_gaussian2d_t* __restrict__ out) {
  int pl = p0 + blockIdx.x;
  int C = n_in[pl];
  unsigned long long T = (unsigned long long)C * (unsigned)(M + 1) + (unsigned)M;
  unsigned long long Tpad = (T + 7ull) & ~7ull;
  const float* D = dense + (toff[pl] - tbase) * PHD_NPLANES;
  phdslam_gaussian2d_t* o = out + out_off[pl];
  for (unsigned long long t = threadIdx.x; t < T; t += blockDim.x) {
    phdslam_gaussian2d_t g;
    g.cov[0] = D[t]; g.cov[1] = D[Tpad + t]; g.cov[2] = D[2 * Tpad + t]; g.cov[3] = D[3 * Tpad + t];
    g.mean[0] = D[4 * Tpad + t]; g.mean[1] = D[5 * Tpad + t]; g.weight = D[6 * Tpad + t];
    o[t] = g;
  }
  (void)np;
}

#endif /* PHD_KERNELS_CUH */
