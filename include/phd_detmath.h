/*
 * phd_detmath.h -- the canonical fp32 arithmetic of the PHD-SLAM hot path.
 *
 * Every function here is built only from IEEE-754 correctly rounded operations
 * (+ - * / sqrt fma, float<->int conversion, bit casts), so a value computed by
 * the sm_100a kernels (compiled with -fmad=false) and by the CPU oracle
 * (compiled with -ffp-contract=off) is bit-identical.  That is what lets the
 * parity tests demand bit-exact component counts and ancestor indices: every
 * threshold test (prune w < min_feature_weight, merge dist < min_separation,
 * in-range tests, CDF search) sees the same bits on both sides.
 *
 * The reference uses libm/CUDA-libm (`exp`, `log`, `atan2f`, `cos`, `sin`,
 * `tan`, `fmod`; src/device_math.cuh:9-16,242-251, src/phdfilter.cu:785-859,
 * 1840-1923) whose results differ between platforms in the last ulp; the
 * functions below are within 2 ulp of those over the ranges the filter uses
 * (checked in tests/test_detmath.py against float64 libm).
 *
 * Usable from C++ (host) and CUDA (host+device).
 */
#ifndef PHD_DETMATH_H
#define PHD_DETMATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <float.h>

#ifdef __CUDACC__
#define PHD_HD __host__ __device__ __forceinline__
#else
#define PHD_HD static inline
#endif

/* reference: `#define LOG0 -FLT_MAX` (src/slamtypes.h:26) */
#define PHD_LOG0 (-FLT_MAX)

#define PHD_PI_F 3.14159274101257324219f       /* float(pi), the smallest float > pi */
#define PHD_TWO_PI_F 6.28318548202514648438f   /* float(2*pi) */
#define PHD_TWO_PI_ERR 1.74845553e-7f          /* float(2*pi) - 2*pi */
#define PHD_LOG_2PI_F 1.83787706640934548356f  /* log(2*pi) */

PHD_HD float phd_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
PHD_HD uint32_t phd_f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}

/* exp(x); results below FLT_MIN flush to 0, x > 88 returns +inf.
 * n = round(x*log2(e)) is taken with the 1.5*2^23 magic constant inside one fused multiply-add, so the
 * scalar function and the packed fp32x2 form used by the update kernel (FFMA2) are the same arithmetic. */
PHD_HD float phd_expf(float x) {
  if (!(x >= -87.3f)) return (x != x) ? x : 0.0f;
  if (x > 88.0f) return INFINITY;
  float nm = fmaf(x, 1.44269504088896341f, 12582912.0f);
  float n = nm - 12582912.0f;
  float r = fmaf(n, -0.693359375f, x);
  r = fmaf(n, 2.12194440e-4f, r);
  float z = r * r;
  float p = 1.9875691500e-4f;
  p = fmaf(p, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  p = fmaf(p, z, r);
  p = p + 1.0f;
  /* p in [0.7,1.5]; scale by 2^n, n in [-126,127] = low mantissa bits of nm */
  return p * phd_u2f(((phd_f2u(nm) - 0x4B400000u) + 127u) << 23);
}

/* log(x) for x > 0 (denormals handled); caller handles x <= 0 (phd_safe_log). */
PHD_HD float phd_logf(float x) {
  int e = 0;
  if (x < FLT_MIN) {
    x = x * 8388608.0f; /* 2^23, exact */
    e = -23;
  }
  uint32_t u = phd_f2u(x);
  e += (int)(u >> 23) - 126;
  float m = phd_u2f((u & 0x007fffffu) | 0x3f000000u); /* [0.5,1) */
  float f;
  if (m < 0.707106781186547524f) {
    e -= 1;
    f = (m + m) - 1.0f;
  } else {
    f = m - 1.0f;
  }
  float z = f * f;
  float y = 7.0376836292e-2f;
  y = fmaf(y, f, -1.1514610310e-1f);
  y = fmaf(y, f, 1.1676998740e-1f);
  y = fmaf(y, f, -1.2420140846e-1f);
  y = fmaf(y, f, 1.4249322787e-1f);
  y = fmaf(y, f, -1.6668057665e-1f);
  y = fmaf(y, f, 2.0000714765e-1f);
  y = fmaf(y, f, -2.4999993993e-1f);
  y = fmaf(y, f, 3.3333331174e-1f);
  y = (y * f) * z;
  float fe = (float)e;
  y = fmaf(fe, -2.12194440e-4f, y);
  y = fmaf(-0.5f, z, y);
  float r = f + y;
  r = fmaf(fe, 0.693359375f, r);
  return r;
}

/* reference safeLog (src/device_math.cuh:9-16): x <= 0 -> LOG0 */
PHD_HD float phd_safe_log(float x) { return (x <= 0.0f) ? PHD_LOG0 : phd_logf(x); }

/* atan(x), |err| < 2 ulp */
PHD_HD float phd_atanf(float x) {
  float ax = fabsf(x);
  float y0, t;
  if (ax > 2.414213562373095f) {
    y0 = 1.57079637050628662109f;
    t = -(1.0f / ax);
  } else if (ax > 0.4142135623730950f) {
    y0 = 0.78539818525314331055f;
    t = (ax - 1.0f) / (ax + 1.0f);
  } else {
    y0 = 0.0f;
    t = ax;
  }
  float z = t * t;
  float p = 8.05374449538e-2f;
  p = fmaf(p, z, -1.38776856032e-1f);
  p = fmaf(p, z, 1.99777106478e-1f);
  p = fmaf(p, z, -3.33329491539e-1f);
  p = (p * z) * t + t;
  float r = y0 + p;
  return (x < 0.0f) ? -r : r;
}

/* atan2(y,x) with the usual quadrant conventions (atan2(0,0) = 0). */
PHD_HD float phd_atan2f(float y, float x) {
  if (x == 0.0f) {
    if (y > 0.0f) return 1.57079637050628662109f;
    if (y < 0.0f) return -1.57079637050628662109f;
    return 0.0f;
  }
  float a = phd_atanf(y / x);
  if (x < 0.0f) a = (y < 0.0f) ? (a - PHD_PI_F) : (a + PHD_PI_F);
  return a;
}

/* sin and cos together, |x| < 8192 (Cody-Waite 3-constant reduction). */
PHD_HD void phd_sincosf(float x, float* s_out, float* c_out) {
  float ax = fabsf(x);
  int j = (int)(1.27323954473516f * ax);
  float y = (float)j;
  if (j & 1) {
    j += 1;
    y += 1.0f;
  }
  float r = fmaf(y, -0.78515625f, ax);
  r = fmaf(y, -2.4187564849853515625e-4f, r);
  r = fmaf(y, -3.77489497744594108e-8f, r);
  float z = r * r;
  /* cos polynomial on [-pi/4,pi/4] */
  float pc = 2.443315711809948e-5f;
  pc = fmaf(pc, z, -1.388731625493765e-3f);
  pc = fmaf(pc, z, 4.166664568298827e-2f);
  pc = (pc * z) * z;
  pc = fmaf(-0.5f, z, pc);
  pc = pc + 1.0f;
  /* sin polynomial */
  float ps = -1.9515295891e-4f;
  ps = fmaf(ps, z, 8.3321608736e-3f);
  ps = fmaf(ps, z, -1.6666654611e-1f);
  ps = (ps * z) * r + r;
  int q = (j >> 1) & 3; /* quadrant: angle = r + q*pi/2 */
  float s, c;
  switch (q) {
    case 0: s = ps; c = pc; break;
    case 1: s = pc; c = -ps; break;
    case 2: s = -ps; c = -pc; break;
    default: s = -pc; c = ps; break;
  }
  *s_out = (x < 0.0f) ? -s : s;
  *c_out = c;
}

PHD_HD float phd_tanf(float x) {
  float s, c;
  phd_sincosf(x, &s, &c);
  return s / c;
}

/*
 * reference wrapAngle (src/device_math.cuh:242-251):
 *   r = fmod(a, float(2*M_PI)); if (r > M_PI) r -= 2*M_PI; else if (r < -M_PI) r += 2*M_PI;
 * fmodf is exact; for |a| < 4*pi it reduces to one exact subtraction (Sterbenz).
 * `r > M_PI` (double compare) <=> r >= float(pi).  The double-precision
 * `r -= 2*M_PI` is reproduced to < 1 ulp with the float(2pi) representation
 * error added back.
 */
PHD_HD float phd_wrap_angle(float a) {
  float r = a;
  float aa = fabsf(a);
  if (aa >= PHD_TWO_PI_F) {
    if (aa < 2.0f * PHD_TWO_PI_F)
      r = (a < 0.0f) ? (a + PHD_TWO_PI_F) : (a - PHD_TWO_PI_F);
    else
      r = fmodf(a, PHD_TWO_PI_F);
  }
  if (r >= PHD_PI_F)
    r = (r - PHD_TWO_PI_F) + PHD_TWO_PI_ERR;
  else if (r <= -PHD_PI_F)
    r = (r + PHD_TWO_PI_F) - PHD_TWO_PI_ERR;
  return r;
}

/* ---------------------------------------------------------------------------
 * Philox4x32-10 counter-based RNG (Salmon et al. 2011).  Replaces the
 * reference's global boost::mt19937 seeded with time(0) (src/rng.cpp:10-35).
 * Counter layout used by the filter:
 *   ctr = { global_particle_index, step, stream, draw_block },  key = { seed_lo, seed_hi }
 * so the noise of particle i at step k is independent of how particles are
 * sharded over GPUs.
 * ------------------------------------------------------------------------- */
typedef struct { uint32_t v[4]; } phd_philox4_t;

PHD_HD uint32_t phd_mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

PHD_HD phd_philox4_t phd_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                       uint32_t k0, uint32_t k1) {
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = phd_mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = phd_mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  phd_philox4_t r;
  r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
  return r;
}

/* uniform in (0,1): 24 random bits, never 0 or 1 */
PHD_HD float phd_u01f(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f + 2.98023223876953125e-8f; }
/* uniform in [0,1) with 53 bits */
PHD_HD double phd_u01d(uint32_t hi, uint32_t lo) {
  uint64_t v = (((uint64_t)hi << 32) | (uint64_t)lo) >> 11;
  return (double)v * 1.1102230246251565404e-16;
}

/* two standard normals from two 32-bit words (Box-Muller on the det. functions) */
PHD_HD void phd_box_muller(uint32_t a, uint32_t b, float* z0, float* z1) {
  float u1 = phd_u01f(a), u2 = phd_u01f(b);
  float rad = sqrtf(-2.0f * phd_logf(u1));
  float s, c;
  phd_sincosf(PHD_TWO_PI_F * u2, &s, &c);
  *z0 = rad * c;
  *z1 = rad * s;
}

/* stream ids for the Philox counter word c2 */
#define PHD_STREAM_PREDICT 1u
#define PHD_STREAM_RESAMPLE 2u
#define PHD_STREAM_SCENE 3u

/* ---------------------------------------------------------------------------
 * Fixed-point encodings used for order-independent (hence GPU-count
 * independent) reductions across particles.
 * ------------------------------------------------------------------------- */
#define PHD_FX_WEIGHT_BITS 36 /* exp(w - wmax) in [0,1] -> Q36: exact sums for up to 2^27 particles */
#define PHD_FX_NEFF_BITS 60   /* exp(2w) with sum <= 1 */
#define PHD_FX_POSE_BITS 40   /* exp(w)*pose, |pose| < 2^20 */
#define PHD_FX_CDF_BITS 40    /* exp(w) for the resampling CDF */

PHD_HD uint64_t phd_fx_from_unit(float v, int bits) {
  /* v in [0,1]; round-to-nearest-even of v*2^bits (double mul is exact: 24-bit mantissa * power of two) */
  double d = (double)v * (double)((uint64_t)1 << bits);
  return (uint64_t)llrint(d);
}
PHD_HD int64_t phd_fx_from_prod(float w, float x, int bits) {
  double d = ((double)w * (double)x) * (double)((uint64_t)1 << bits);
  return (int64_t)llrint(d);
}

/* ---------------------------------------------------------------------------
 * CPHD: log of the cardinality scale n_c of the linear-domain tables (kernels.cuh cphd_block, oracle cphd_factors:
 * c[n] = p-(n) n! / n_c^n, d[k] = (q n_c / <1,w>)^k / k!).  n!/n_c^n <= 1 needs n_c >= n/e, d[k] <= e^n_c needs
 * n_c < 709: 128 up to 345 cardinality bins, 256 up to 690, 512 up to the maximum of 1024 bins.
 * One function for kernel and oracle: the value is a literal on both sides.
 * ------------------------------------------------------------------------- */
PHD_HD double phd_cphd_log_nc(int n_bins) {
  return (n_bins <= 345) ? 4.852030263919617 /* log 128 */ : (n_bins <= 690) ? 5.545177444479562 /* log 256 */
                                                                            : 6.238324625039508; /* log 512 */
}

#endif /* PHD_DETMATH_H */
