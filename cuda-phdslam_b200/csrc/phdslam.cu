/*
 * phdslam.cu -- C-ABI of libphdslam.so (see include/phdslam.h): persistent device state and the
 * launch sequence of the filter step.  One handle = one GPU = one host thread (one rank).
 *
 * Reference call sequence being replaced: run_synth loop body (src/main.cpp:1231-1297) ->
 * phdPredict (src/phdfilter.cu:1080) -> phdUpdateSynth (:3336) -> recoverSlamState (main.cpp:318)
 * -> resampleParticles (main.cpp:453).  The reference keeps all state on the host and pays
 * >= 20 cudaMalloc + >= 12 cudaMemcpy per step; here nothing but the M measurements goes up and
 * a 100-byte estimate comes down.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>

/* NCCL is bound at run time (dlopen), not at link time: a process that also hosts PyTorch must use the one
 * libnccl.so.2 PyTorch brought along, whichever of the two libraries is loaded first. */
namespace nccl_rt {
static void* lib = nullptr;
static ncclResult_t (*GetUniqueId)(ncclUniqueId*);
static ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
static ncclResult_t (*CommDestroy)(ncclComm_t);
static ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
static ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
static ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
static ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
static ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
static ncclResult_t (*GroupStart)();
static ncclResult_t (*GroupEnd)();
static const char* (*GetErrorString)(ncclResult_t);
static bool load() {
  if (lib) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return false;
#define SYM(name) *(void**)(&name) = dlsym(h, "nccl" #name); if (!name) return false;
  SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(AllReduce) SYM(AllGather) SYM(Broadcast) SYM(Send) SYM(Recv)
  SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
  lib = h;
  return true;
}
}  // namespace nccl_rt
#define ncclGetUniqueId nccl_rt::GetUniqueId
#define ncclCommInitRank nccl_rt::CommInitRank
#define ncclCommDestroy nccl_rt::CommDestroy
#define ncclAllReduce nccl_rt::AllReduce
#define ncclAllGather nccl_rt::AllGather
#define ncclBroadcast nccl_rt::Broadcast
#define ncclSend nccl_rt::Send
#define ncclRecv nccl_rt::Recv
#define ncclGroupStart nccl_rt::GroupStart
#define ncclGroupEnd nccl_rt::GroupEnd
#define ncclGetErrorString nccl_rt::GetErrorString

#include "kernels.cuh"
#include "mixed.cuh"
#include "phdslam_internal.h"

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      phdslam_set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                        std::to_string(__LINE__) + ")");                                             \
      return PHDSLAM_ERR_CUDA;                                                                       \
    }                                                                                                \
  } while (0)

#define CKN(call)                                                                                    \
  do {                                                                                               \
    ncclResult_t e__ = (call);                                                                       \
    if (e__ != ncclSuccess) {                                                                        \
      phdslam_set_error(std::string(#call) + ": " + ncclGetErrorString(e__) + " (" + __FILE__ + ":" + \
                        std::to_string(__LINE__) + ")");                                             \
      return PHDSLAM_ERR_NCCL;                                                                       \
    }                                                                                                \
  } while (0)

/* Every entry point that touches particle state: select the device and allocate the state on first use.  The allocation
 * is deferred from phdslam_create so that phdslam_dist_init can allocate this rank's SHARE only (16.7 M particles x 12 KB of
 * map buffers do not fit one GPU; a rank's share of them does). */
static int ensure_state(phdslam* h);
#define ENTER(h)                                  \
  do {                                            \
    CK(cudaSetDevice((h)->device));               \
    if (!(h)->state_ready) {                      \
      int rc__ = ensure_state(h);                 \
      if (rc__) return rc__;                      \
    }                                             \
  } while (0)

#define LAUNCH_CHECK(h)            \
  do {                             \
    (h)->launches++;               \
    CK(cudaGetLastError());        \
  } while (0)

/* every host<->device copy of the library goes through these four: the byte counters of phdslam_timings_t are counted
 * from the calls, not estimated (bench.py's e2e.h2d_bytes_per_step / d2h_bytes_per_step) */
static inline cudaError_t copy_h2d_async(phdslam* h, void* dst, const void* src, size_t n, cudaStream_t st) {
  h->tim.h2d_bytes += n;
  return cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, st);
}
static inline cudaError_t copy_d2h_async(phdslam* h, void* dst, const void* src, size_t n, cudaStream_t st) {
  h->tim.d2h_bytes += n;
  return cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, st);
}
static inline cudaError_t copy_h2d(phdslam* h, void* dst, const void* src, size_t n) {
  h->tim.h2d_bytes += n;
  return cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice);
}
static inline cudaError_t copy_d2h(phdslam* h, void* dst, const void* src, size_t n) {
  h->tim.d2h_bytes += n;
  return cudaMemcpy(dst, src, n, cudaMemcpyDeviceToHost);
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

static float float_floor_d(double d) {
  float f = (float)d;
  if ((double)f > d) f = nextafterf(f, -INFINITY);
  return f;
}
static float float_ceil_d(double d) {
  float f = (float)d;
  if ((double)f < d) f = nextafterf(f, INFINITY);
  return f;
}

/* host evaluation of the canonical log (phd_detmath.h is __host__ __device__) */
static void derive_devcfg(const phdslam_config_t& c, int Cmax, DevCfg* d) {
  memset(d, 0, sizeof(*d));
  d->min_range = c.min_range; d->max_range = c.max_range; d->max_bearing = c.max_bearing;
  d->lo2 = float_ceil_d(0.8 * (double)c.min_range);
  d->hi2 = float_floor_d(1.2 * (double)c.max_range);
  d->hb2 = float_floor_d(1.2 * (double)c.max_bearing);
  d->var_r = c.std_range * c.std_range;
  d->var_b = c.std_bearing * c.std_bearing;
  float sr = c.std_range * c.birth_noise_factor, sb = c.std_bearing * c.birth_noise_factor;
  d->bvar_r = sr * sr;
  d->bvar_b = sb * sb;
  d->pd = c.pd;
  d->log_pd = phd_safe_log(c.pd);
  d->clutter_density = c.clutter_density;
  d->clutter_rate = c.clutter_rate;
  d->log_clutter_rate = phd_safe_log(c.clutter_rate);
  d->birth_weight = c.birth_weight;
  d->log_birth_weight = phd_safe_log(c.birth_weight);
  d->min_sep = c.min_separation;
  d->min_w = c.min_feature_weight;
  d->distance_metric = c.distance_metric;
  d->particle_weighting = c.particle_weighting;
  d->labeled = c.labeled_measurements;
  d->filter_type = c.filter_type;
  d->motion_type = c.motion_type;
  d->dt_sub = c.dt / (float)c.subdivide_predict;
  d->l = c.l; d->h = c.h; d->a = c.a; d->b = c.b;
  d->std_alpha = c.std_alpha; d->std_enc = c.std_encoder;
  d->ax3 = 3.0f * c.ax; d->ay3 = 3.0f * c.ay; d->ayaw3 = 3.0f * c.ayaw;
  d->seed_lo = (uint32_t)c.seed; d->seed_hi = (uint32_t)(c.seed >> 32);
  d->Cmax = Cmax;
  d->n_card = (c.filter_type == 1) ? c.max_cardinality + 1 : 0;
}

static int validate_config(const phdslam_config_t* c) {
  if (c->n_particles < 1 || c->max_components < 1) {
    phdslam_set_error("n_particles and max_components must be >= 1");
    return PHDSLAM_ERR_INVALID;
  }
  if (c->feature_model != 0 && c->feature_model != 2) {
    /* DYNAMIC_MODEL: the reference's update launch for it is commented out (src/phdfilter.cu:3663-3670) */
    phdslam_set_error("feature_model must be 0 (static) or 2 (mixed: static + constant-velocity features)");
    return PHDSLAM_ERR_INVALID;
  }
  if (c->feature_model == 2 && (c->filter_type != 0 || c->n_predict_particles != 1 || c->max_components_dynamic < 1 ||
                                c->max_components_dynamic > 1024)) {
    phdslam_set_error("feature_model = 2 needs filter_type = 0 (phdUpdateKernelMixed is a PHD update), n_predict_particles = 1 "
                      "and 1 <= max_components_dynamic <= 1024");
    return PHDSLAM_ERR_INVALID;
  }
  if (c->n_predict_particles < 1 || c->n_predict_particles > 64) {
    phdslam_set_error("n_predict_particles must be in [1, 64]");
    return PHDSLAM_ERR_INVALID;
  }
  if (c->filter_type == 1 && (c->max_cardinality < 1 || c->max_cardinality + 1 >= PHD_LF_MAX)) {
    phdslam_set_error("CPHD: max_cardinality must be in [1, 1023]");
    return PHDSLAM_ERR_INVALID;
  }
  if (c->filter_type != 0 && c->filter_type != 1) {
    phdslam_set_error("filter_type must be 0 (PHD) or 1 (CPHD)");
    return PHDSLAM_ERR_INVALID;
  }
  if (c->particle_weighting < 0 || c->particle_weighting > 1) {
    /* scheme 2 (single-feature, src/phdfilter.cu:3600-3661) is an unfinished host fallback in the reference itself (mirrored
     * feature index, evalGaussianMixture with the wrong sign and no weights: DESIGN.md section 2); refusing it is better
     * than a run whose particle weights are silently never updated */
    phdslam_set_error("particle_weighting must be 0 (cluster process) or 1 (Vo empty map); scheme 2 is not reproduced");
    return PHDSLAM_ERR_INVALID;
  }
  if (c->subdivide_predict < 1) {
    phdslam_set_error("subdivide_predict must be >= 1");
    return PHDSLAM_ERR_INVALID;
  }
  return 0;
}

/* layout of the peer window for a rank with n particles: 256-byte header (magic), then the arrays below */
struct SlabLayout {
  size_t pose[2], count[2], map[2], card[2], anc_in, mbox, bytes;
};
static SlabLayout slab_layout(size_t n, size_t Cmax, size_t n_card) {
  SlabLayout L;
  size_t o = 256;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
  for (int b = 0; b < 2; ++b) {
    L.pose[b] = take(6 * n * sizeof(float));
    L.count[b] = take(n * sizeof(int));
    L.map[b] = take(n * PHD_MAP_PLANES * Cmax * sizeof(float));
    L.card[b] = take(n * n_card * sizeof(float));
  }
  L.anc_in = take(n * sizeof(int));
  L.mbox = take((size_t)MBOX_SLOTS * PHD_MAX_PEERS * MBOX_WORDS * sizeof(unsigned long long));
  L.bytes = std::max(o, (size_t)4 << 20);      /* >= 4 MB: an allocation of its own, never a sub-allocation */
  return L;
}

static inline int rank_offset_of(long long n_global, int world, int r) { return (int)(n_global * r / world); }

static void close_peers(phdslam* h) {
  if (h->peer_base) {
    for (int r = 0; r < h->world; ++r)
      if (r != h->rank && h->peer_base[r]) cudaIpcCloseMemHandle(h->peer_base[r]);
    delete[] h->peer_base;
    h->peer_base = nullptr;
  }
  h->p2p = 0;
}

static void free_state(phdslam* h) {
  close_peers(h);
  if (h->peer_slab) {
    cudaFree(h->peer_slab);
    h->peer_slab = nullptr;
    h->anc_in = nullptr;
    for (int b = 0; b < 2; ++b) { h->pose[b] = nullptr; h->count[b] = nullptr; h->map[b] = nullptr; h->card[b] = nullptr; }
  }
  cudaFree(h->barrier_dev);
  h->barrier_dev = nullptr;
  cudaFree(h->gath_dev);
  h->gath_dev = nullptr;
  if (h->gath_host) cudaFreeHost(h->gath_host);
  h->gath_host = nullptr;
  h->mbox = 0; h->mbox_base = nullptr; h->totals_valid = 0;
  for (int b = 0; b < 2; ++b) {
    cudaFree(h->pose[b]); cudaFree(h->count[b]); cudaFree(h->map[b]); cudaFree(h->card[b]);
  }
  cudaFree(h->logw); cudaFree(h->resample_idx);
  cudaFree(h->snap_pose); cudaFree(h->snap_count); cudaFree(h->snap_map); cudaFree(h->snap_card); cudaFree(h->snap_logw);
  cudaFree(h->cls); cudaFree(h->n_in); cudaFree(h->dlogw); cudaFree(h->tpad); cudaFree(h->toff); cudaFree(h->scan_tmp);
  cudaFree(h->dense); cudaFree(h->z_dev); cudaFree(h->draws_dev); cudaFree(h->q_fx); cudaFree(h->cdf_excl);
  cudaFree(h->ancestors); cudaFree(h->red); cudaFree(h->cand); cudaFree(h->cand_in); cudaFree(h->n_cand); cudaFree(h->ovf_list);
  cudaFree(h->mig_map); cudaFree(h->mig_pose); cudaFree(h->mig_count); cudaFree(h->mig_anc); cudaFree(h->mig_card);
  cudaFree(h->mig_pose_in); cudaFree(h->totals_dev); cudaFree(h->lfact); cudaFree(h->mig_anc2);
  for (int b = 0; b < 2; ++b) { cudaFree(h->dmap[b]); cudaFree(h->dcount[b]); h->dmap[b] = nullptr; h->dcount[b] = nullptr; }
  cudaFree(h->mix_dsum); cudaFree(h->mix_nhat); cudaFree(h->mix_L); cudaFree(h->dcand); cudaFree(h->snap_dmap); cudaFree(h->snap_dcount);
  cudaFree(h->dyn_mig_map); cudaFree(h->dyn_mig_count); cudaFree(h->dyn_mig_anc);
  h->dyn_mig_map = nullptr; h->dyn_mig_count = nullptr; h->dyn_mig_anc = nullptr; h->dyn_mig_cap = 0;
  h->mix_dsum = h->mix_nhat = h->mix_L = nullptr; h->dcand = nullptr; h->snap_dmap = nullptr; h->snap_dcount = nullptr;
  h->mig_pose_in = nullptr; h->totals_dev = nullptr; h->lfact = nullptr; h->mig_anc2 = nullptr; h->mig_anc_cap = 0;
  if (h->red_host) cudaFreeHost(h->red_host);
  if (h->z_host) cudaFreeHost(h->z_host);
  h->red_host = nullptr; h->z_host = nullptr;
}

/* Particle capacity.  With n_predict_particles = k > 1 every prediction multiplies the particle count by k and the loop
 * resamples back to n_particles once the count exceeds 5 n_particles (src/main.cpp:1286): the count before a prediction
 * is at most 5 n, so the largest reachable count is g^m n with g = k^subdivide_predict and g^(m-1) <= 5 < g^m. */
static int particle_capacity(const phdslam_config_t& c, int n) {
  if (c.n_predict_particles <= 1) return n;
  long long g = 1;                              /* growth per time step: every sub-step of the prediction fans out */
  for (int i = 0; i < std::max(c.subdivide_predict, 1); ++i) g *= c.n_predict_particles;
  long long f = 1;
  while (f <= 5) f *= g;
  return (int)std::min<long long>(f * n, 0x7fffffffLL);
}

static int alloc_state(phdslam* h) {
  h->n_cap = (h->world > 1) ? h->n_local : particle_capacity(h->cfg, h->n_local);
  const size_t n = (size_t)h->n_cap;
  const size_t C = (size_t)h->Cmax;
  if (h->world > 1) {
    /* one peer-mappable allocation (see phdslam_internal.h) */
    const SlabLayout L = slab_layout(n, C, (size_t)h->n_card);
    CK(cudaMalloc(&h->peer_slab, L.bytes));
    h->peer_slab_bytes = L.bytes;
    for (int b = 0; b < 2; ++b) {
      h->pose[b] = reinterpret_cast<float*>(h->peer_slab + L.pose[b]);
      h->count[b] = reinterpret_cast<int*>(h->peer_slab + L.count[b]);
      h->map[b] = reinterpret_cast<float*>(h->peer_slab + L.map[b]);
      h->card[b] = h->n_card ? reinterpret_cast<float*>(h->peer_slab + L.card[b]) : nullptr;
    }
    h->anc_in = reinterpret_cast<int*>(h->peer_slab + L.anc_in);
    h->mbox_base = reinterpret_cast<unsigned long long*>(h->peer_slab + L.mbox);
    CK(cudaMemset(h->mbox_base, 0, (size_t)MBOX_SLOTS * PHD_MAX_PEERS * MBOX_WORDS * sizeof(unsigned long long)));
    h->mbox_seq = 1;
    CK(cudaMalloc(&h->gath_dev, (size_t)PHD_MAX_PEERS * MBOX_WORDS * sizeof(unsigned long long)));
    CK(cudaMallocHost(&h->gath_host, (size_t)PHD_MAX_PEERS * MBOX_WORDS * sizeof(unsigned long long)));
    h->slab_sw = 0;
    CK(cudaMalloc(&h->barrier_dev, 2 * sizeof(int)));
    CK(cudaMemset(h->barrier_dev, 0, 2 * sizeof(int)));
  } else
  for (int b = 0; b < 2; ++b) {
    CK(cudaMalloc(&h->pose[b], 6 * n * sizeof(float)));
    CK(cudaMalloc(&h->count[b], n * sizeof(int)));
    CK(cudaMalloc(&h->map[b], n * PHD_MAP_PLANES * C * sizeof(float)));
    if (h->n_card) CK(cudaMalloc(&h->card[b], n * h->n_card * sizeof(float)));
  }
  CK(cudaMalloc(&h->logw, n * sizeof(float)));
  CK(cudaMalloc(&h->resample_idx, n * sizeof(int)));
  CK(cudaMalloc(&h->cls, n * C));
  CK(cudaMalloc(&h->n_in, n * sizeof(int)));
  CK(cudaMalloc(&h->dlogw, n * sizeof(float)));
  CK(cudaMalloc(&h->n_cand, n * sizeof(int)));
  CK(cudaMalloc(&h->ovf_list, n * sizeof(int)));
  CK(cudaMalloc(&h->tpad, n * sizeof(unsigned long long)));
  CK(cudaMalloc(&h->toff, (n + 1) * sizeof(unsigned long long)));
  CK(cudaMalloc(&h->scan_tmp, (size_t)(cdiv(n, SCAN_TILE) + 1) * sizeof(unsigned long long)));
  CK(cudaMalloc(&h->z_dev, 3 * PHD_MAX_MEAS * sizeof(float)));
  CK(cudaMalloc(&h->q_fx, n * sizeof(unsigned long long)));
  CK(cudaMalloc(&h->cdf_excl, (n + 1) * sizeof(unsigned long long)));
  CK(cudaMalloc(&h->ancestors, n * sizeof(int)));
  CK(cudaMalloc(&h->red, sizeof(Reductions)));
  CK(cudaMallocHost(&h->red_host, sizeof(Reductions)));
  CK(cudaMallocHost(&h->z_host, 3 * PHD_MAX_MEAS * sizeof(float)));
  CK(cudaMalloc(&h->lfact, PHD_LF_MAX * sizeof(float)));
  if (h->Dmax) {
    for (int b = 0; b < 2; ++b) {
      /* one particle of slack: the all-gather of a sharded run sends every rank's block padded to the largest share */
      CK(cudaMalloc(&h->dmap[b], (n + 1) * DYN_PLANES * (size_t)h->Dmax * sizeof(float)));
      CK(cudaMalloc(&h->dcount[b], (n + 1) * sizeof(int)));
      CK(cudaMemset(h->dcount[b], 0, (n + 1) * sizeof(int)));
    }
    CK(cudaMalloc(&h->mix_dsum, n * PHD_MAX_MEAS * sizeof(float)));
    CK(cudaMalloc(&h->mix_L, n * PHD_MAX_MEAS * sizeof(float)));
    CK(cudaMalloc(&h->mix_nhat, n * sizeof(float)));
    CK(cudaMalloc(&h->dcand, n * (size_t)h->Sd * sizeof(phdslam_gaussian4d_t)));
  }
  return 0;
}

static int init_particles(phdslam* h) {
  /* run_synth initialisation, src/main.cpp:1129-1144 */
  const int n = h->n_local;
  const phdslam_config_t& c = h->cfg;
  const float init[6] = {c.x0, c.y0, c.yaw0, c.vx0, c.vy0, c.vyaw0};
  for (int k = 0; k < 6; ++k) {
    fill_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->pose[h->cur] + (size_t)k * n, n, init[k]);
    LAUNCH_CHECK(h);
  }
  fill_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->logw, n, -phd_logf((float)h->n_global));
  LAUNCH_CHECK(h);
  CK(cudaMemsetAsync(h->count[0], 0, (size_t)n * sizeof(int), h->stream));
  CK(cudaMemsetAsync(h->count[1], 0, (size_t)n * sizeof(int), h->stream));
  iota_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->resample_idx, n, h->offset);
  LAUNCH_CHECK(h);
  if (h->n_card) {
    fill_kernel<<<cdiv((long long)n * h->n_card, 256), 256, 0, h->stream>>>(h->card[h->cur], n * h->n_card,
                                                                            -phd_logf((float)h->n_card));
    LAUNCH_CHECK(h);
  }
  if (h->Dmax) {
    CK(cudaMemsetAsync(h->dcount[0], 0, (size_t)n * sizeof(int), h->stream));
    CK(cudaMemsetAsync(h->dcount[1], 0, (size_t)n * sizeof(int), h->stream));
  }
  lfact_kernel<<<1, 32, 0, h->stream>>>(h->lfact, PHD_LF_MAX);
  LAUNCH_CHECK(h);
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

static int create_impl(phdslam* h, const phdslam_config_t* cfg, int device);

extern "C" int phdslam_create(const phdslam_config_t* cfg, int device, phdslam_t** out) {
  if (!cfg || !out) return PHDSLAM_ERR_INVALID;
  int rc = validate_config(cfg);
  if (rc) return rc;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    phdslam_set_error("no CUDA device: libphdslam has no CPU fallback");
    return PHDSLAM_ERR_CUDA;
  }
  CK(cudaSetDevice(device));
  phdslam* h = new phdslam();
  memset(h, 0, sizeof(*h));
  rc = create_impl(h, cfg, device);
  if (rc) {                      /* nothing leaks on a failed create: streams, events and whatever was allocated go */
    const std::string why = phdslam_last_error();
    phdslam_destroy(h);
    phdslam_set_error(why);
    return rc;
  }
  *out = h;
  return 0;
}

static int create_impl(phdslam* h, const phdslam_config_t* cfg, int device) {
  int rc;
  h->cfg = *cfg;
  h->device = device;
  h->rank = 0; h->world = 1;
  h->n_global = cfg->n_particles; h->n_local = cfg->n_particles; h->offset = 0;
  h->Cmax = (cfg->max_components + 31) & ~31;
  h->n_card = (cfg->filter_type == 1) ? cfg->max_cardinality + 1 : 0;
  int smax = 1024; /* merge candidates per particle: survivors of the prune + nearly-in-range components */
  while (smax < 2 * h->Cmax + PHD_MAX_MEAS && smax < 4096) smax <<= 1;
  h->Smax = smax;
  if (cfg->feature_model == 2) {
    h->Dmax = (cfg->max_components_dynamic + 7) & ~7;
    h->Sd = 4 * h->Dmax + PHD_MAX_MEAS;   /* prune survivors of one particle's dynamic update (more -> PHDSLAM_ERR_CAPACITY) */
  }
  derive_devcfg(h->cfg, h->Cmax, &h->dc);
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  for (int i = 0; i < 12; ++i) CK(cudaEventCreate(&h->ev[i]));
  for (int i = 0; i < 4; ++i) CK(cudaEventCreate(&h->ev_dyn[i]));
  {
    /* Optional (PHDSLAM_OVERLAP=1 / phdslam_set_overlap): the merge of sub-batch k runs on a higher-priority stream
     * while the update of sub-batch k+1 streams.  Measured on B200 at 65 536 x 256 x 64: no gain (15.8 ms vs 15.1 ms
     * serial) -- a merge CTA (32 KB of shared memory, 6 per SM) and an update CTA (45 KB, 4 per SM) do not share an SM
     * profitably: both kernels need their full occupancy to hide latency, so the SMs time-slice either way.  Off by
     * default. */
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&h->stream_m, cudaStreamNonBlocking, hi));
    for (int i = 0; i < PHD_MAX_SUB; ++i) CK(cudaEventCreateWithFlags(&h->ev_sub[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_merge_done, cudaEventDisableTiming));
    const char* e = getenv("PHDSLAM_OVERLAP");
    h->overlap = (e && atoi(e) != 0) ? 1 : 0;
  }
  CK(cudaFuncSetAttribute(update_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem_bytes(h->Cmax)));
  CK(cudaFuncSetAttribute(update_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem_bytes(h->Cmax)));
  if (h->n_card) {
    const size_t sm = update_smem_bytes(h->Cmax) + cphd_smem_bytes(h->n_card, PHD_MAX_MEAS) + cphd_chunkmax_bytes(PHD_MAX_MEAS, h->Cmax);
    if (sm > 227 * 1024) {
      phdslam_set_error("CPHD: max_components / max_cardinality need more than 227 KB of shared memory per particle");
      return PHDSLAM_ERR_INVALID;
    }
    CK(cudaFuncSetAttribute(update_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    CK(cudaFuncSetAttribute(update_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  }
  if (h->Dmax) {
    CK(cudaFuncSetAttribute(update_mixed_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem_bytes(h->Cmax)));
    CK(cudaFuncSetAttribute(update_mixed_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem_bytes(h->Cmax)));
    CK(cudaFuncSetAttribute(dyn_pre_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem_bytes(h->Dmax, PHD_MAX_MEAS, h->Sd)));
    CK(cudaFuncSetAttribute(dyn_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem_bytes(h->Dmax, PHD_MAX_MEAS, h->Sd)));
  }
  CK(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem_bytes(h->Smax)));
  {
    /* merge_fast_kernel keeps a particle's candidates in shared memory: capacity Scap (multiple of 32), adapted after
     * every update to the largest candidate count seen (+12.5 %); particles above it take merge_kernel.
     * PHDSLAM_MERGE_CAP pins the capacity (0 = always merge_kernel): a test / profiling knob. */
    h->Scap_max = std::min(h->Smax, 4096);
    while (merge_fast_smem_bytes(h->Scap_max) > 200 * 1024) h->Scap_max -= 32;
    h->Scap = std::min(h->Scap_max, std::max(256, (h->Cmax + h->Cmax / 4 + 31) & ~31));
    h->Scap_pinned = 0;
    if (const char* e = getenv("PHDSLAM_MERGE_CAP")) {
      int v = atoi(e);
      h->Scap = (v <= 0) ? 0 : std::min(h->Scap_max, std::max(32, (v + 31) & ~31));
      h->Scap_pinned = 1;
    }
    CK(cudaFuncSetAttribute(merge_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_fast_smem_bytes(h->Scap_max)));
  }
  (void)rc;
  return 0;       /* particle state: allocated and initialised on first use (ensure_state) or by phdslam_dist_init */
}

static int ensure_state(phdslam* h) {
  int rc = alloc_state(h);
  if (rc) return rc;
  rc = init_particles(h);
  if (rc) return rc;
  h->state_ready = 1;
  return 0;
}

extern "C" void phdslam_destroy(phdslam_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->stream_m) cudaStreamSynchronize(h->stream_m);
  if (h->nccl_comm) ncclCommDestroy((ncclComm_t)h->nccl_comm);
  free_state(h);
  for (int i = 0; i < 12; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
  for (int i = 0; i < 4; ++i) if (h->ev_dyn[i]) cudaEventDestroy(h->ev_dyn[i]);
  for (int i = 0; i < PHD_MAX_SUB; ++i) if (h->ev_sub[i]) cudaEventDestroy(h->ev_sub[i]);
  if (h->ev_merge_done) cudaEventDestroy(h->ev_merge_done);
  if (h->stream_m) cudaStreamDestroy(h->stream_m);
  if (h->stream) cudaStreamDestroy(h->stream);
  cudaGetLastError();
  delete h;
}

extern "C" int phdslam_set_config(phdslam_t* h, const phdslam_config_t* cfg) {
  if (cfg->n_particles != h->cfg.n_particles || ((cfg->max_components + 31) & ~31) != h->Cmax ||
      cfg->filter_type != h->cfg.filter_type || cfg->n_predict_particles != h->cfg.n_predict_particles ||
      cfg->feature_model != h->cfg.feature_model ||
      (cfg->feature_model == 2 && ((cfg->max_components_dynamic + 7) & ~7) != h->Dmax) ||
      (cfg->n_predict_particles > 1 && cfg->subdivide_predict != h->cfg.subdivide_predict)) {
    phdslam_set_error("n_particles / max_components / filter_type / n_predict_particles / feature_model are fixed at create time");
    return PHDSLAM_ERR_INVALID;
  }
  int rc = validate_config(cfg);
  if (rc) return rc;
  h->cfg = *cfg;
  derive_devcfg(h->cfg, h->Cmax, &h->dc);
  return 0;
}
extern "C" int phdslam_get_config(const phdslam_t* h, phdslam_config_t* cfg) { *cfg = h->cfg; return 0; }
extern "C" int phdslam_n_local(const phdslam_t* h) { return h->n_local; }
extern "C" int phdslam_local_offset(const phdslam_t* h) { return h->offset; }
extern "C" void* phdslam_stream(phdslam_t* h) { return (void*)h->stream; }
extern "C" int phdslam_synchronize(phdslam_t* h) { CK(cudaStreamSynchronize(h->stream)); return 0; }
extern "C" int phdslam_set_overlap(phdslam_t* h, int on) { h->overlap = on ? 1 : 0; return 0; }
extern "C" int phdslam_dist_p2p(const phdslam_t* h) { return h->p2p; }

extern "C" int phdslam_dist_unique_id(void* id128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (!nccl_rt::load()) {
    phdslam_set_error("libnccl.so.2 not found");
    return PHDSLAM_ERR_NCCL;
  }
  ncclUniqueId id;
  CKN(ncclGetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

/* Map every peer's window through CUDA IPC (one process per GPU on one node).  Any failure -- IPC unsupported, no peer
 * access between two devices, PHDSLAM_P2P=0 -- leaves p2p = 0 on EVERY rank (the decision is all-reduced) and the
 * resampling exchange uses the NCCL send/recv ring instead. */
static int map_peer_windows(phdslam* h) {
  const int W = h->world, me = h->rank;
  ncclComm_t comm = (ncclComm_t)h->nccl_comm;
  const char* env = getenv("PHDSLAM_P2P");
  int ok = (env && atoi(env) == 0) ? 0 : 1;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  const unsigned long long magic = 0x5048445f50325000ull + (unsigned)me;
  if (ok && cudaIpcGetMemHandle(&mine, h->peer_slab) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  CK(copy_h2d_async(h, h->peer_slab, &magic, sizeof(magic), h->stream));
  unsigned char* hbuf = nullptr;                          /* [W] handles, all-gathered */
  CK(cudaMalloc(&hbuf, (size_t)W * sizeof(mine)));
  CK(copy_h2d_async(h, hbuf + (size_t)me * sizeof(mine), &mine, sizeof(mine), h->stream));
  CKN(ncclAllGather(hbuf + (size_t)me * sizeof(mine), hbuf, sizeof(mine), ncclChar, comm, h->stream));
  std::vector<cudaIpcMemHandle_t> all(W);
  CK(copy_d2h_async(h, all.data(), hbuf, (size_t)W * sizeof(mine), h->stream));
  CK(cudaStreamSynchronize(h->stream));                   /* every rank's magic is in place: the all-gather completed */
  h->peer_base = new unsigned char*[W];
  for (int r = 0; r < W; ++r) h->peer_base[r] = nullptr;
  h->peer_base[me] = h->peer_slab;
  for (int r = 0; r < W && ok; ++r) {
    if (r == me) continue;
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    h->peer_base[r] = (unsigned char*)p;
    unsigned long long seen = 0;
    if (copy_d2h(h, &seen, p, sizeof(seen)) != cudaSuccess ||
        seen != 0x5048445f50325000ull + (unsigned)r) { cudaGetLastError(); ok = 0; }
  }
  /* unanimous decision */
  int* flag = reinterpret_cast<int*>(hbuf);
  CK(copy_h2d_async(h, flag, &ok, sizeof(int), h->stream));
  CKN(ncclAllReduce(flag, flag, 1, ncclInt32, ncclMin, comm, h->stream));
  CK(copy_d2h_async(h, &ok, flag, sizeof(int), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaFree(hbuf);
  if (!ok) close_peers(h);
  h->p2p = ok;
  {
    const char* e = getenv("PHDSLAM_MBOX");       /* 0: keep the NCCL collectives for the statistics (A/B, fallback) */
    h->mbox = (ok && W <= PHD_MAX_PEERS && !(e && atoi(e) == 0)) ? 1 : 0;
  }
  return 0;
}

/* One mailbox exchange on the handle's stream: words_dev[0..n_words) of every rank combined in place (bit t of max_mask:
 * word t by max, otherwise by sum); with want_gathered the W raw records land in gath_dev.  See mbox_exchange_kernel. */
static int mbox_exchange(phdslam* h, unsigned long long* words_dev, int n_words, unsigned max_mask, bool want_gathered) {
  MboxPeers pp;
  for (int r = 0; r < PHD_MAX_PEERS; ++r) pp.box[r] = nullptr;
  for (int r = 0; r < h->world; ++r) {
    const size_t nr = (size_t)(rank_offset_of(h->n_global, h->world, r + 1) - rank_offset_of(h->n_global, h->world, r));
    const SlabLayout L = slab_layout(nr, (size_t)h->Cmax, (size_t)h->n_card);
    pp.box[r] = reinterpret_cast<unsigned long long*>(h->peer_base[r] + L.mbox);
  }
  mbox_exchange_kernel<<<1, 32 * h->world, 0, h->stream>>>(pp, h->rank, h->world, h->mbox_seq, words_dev, n_words, max_mask,
                                                          want_gathered ? h->gath_dev : nullptr, &h->red->err_flag);
  LAUNCH_CHECK(h);
  h->mbox_seq++;
  return 0;
}

/* Shard the particle set: rank r owns global particles [r*N/world, (r+1)*N/world).  Re-allocates the device
 * state for the local share and joins the NCCL communicator used by the weight statistics (all-reduce) and by
 * the resampling migration (send/recv).  The reference is single-GPU (src/main.cpp:1445-1453). */
extern "C" int phdslam_dist_init(phdslam_t* h, int rank, int world, const void* id128) {
  if (world < 1 || rank < 0 || rank >= world) return PHDSLAM_ERR_INVALID;
  if (world == 1) return 0;
  if (h->n_global < world) {
    phdslam_set_error("fewer particles than ranks");
    return PHDSLAM_ERR_INVALID;
  }
  if (!nccl_rt::load()) {
    phdslam_set_error("libnccl.so.2 not found");
    return PHDSLAM_ERR_NCCL;
  }
  if (h->cfg.n_predict_particles > 1) {
    phdslam_set_error("n_predict_particles > 1 changes the particle count every step and is single-GPU only");
    return PHDSLAM_ERR_INVALID;
  }
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  CKN(ncclCommInitRank(&comm, world, id, rank));
  h->nccl_comm = (void*)comm;
  h->rank = rank;
  h->world = world;
  if (h->state_ready) free_state(h);
  h->state_ready = 0;
  h->dense = nullptr; h->dense_floats = 0; h->cand = h->cand_in = nullptr; h->cand_cap = 0;
  h->draws_dev = nullptr; h->draws_cap = 0;
  h->snap_pose = nullptr; h->snap_count = nullptr; h->snap_map = nullptr; h->snap_card = nullptr; h->snap_logw = nullptr;
  h->mig_map = nullptr; h->mig_cap = 0;
  const long long N = h->n_global;
  h->offset = (int)(N * rank / world);
  h->n_local = (int)(N * (rank + 1) / world) - h->offset;
  h->cur = 0;
  int rc = alloc_state(h);
  if (rc) return rc;
  rc = init_particles(h);
  if (rc) return rc;
  h->state_ready = 1;
  return map_peer_windows(h);
}

static inline int rank_offset(const phdslam* h, int r) { return (int)((long long)h->n_global * r / h->world); }

/* ---- exclusive scan helper ---- */
static int scan_u64(phdslam* h, const unsigned long long* in, int n, unsigned long long* out /* n+1 */,
                    unsigned long long* grand_dev) {
  int tiles = cdiv(n, SCAN_TILE);
  scan_tile_sums_kernel<<<tiles, SCAN_THREADS, 0, h->stream>>>(in, n, h->scan_tmp);
  LAUNCH_CHECK(h);
  scan_tile_offsets_kernel<<<1, SCAN_THREADS, 0, h->stream>>>(h->scan_tmp, tiles, grand_dev);
  LAUNCH_CHECK(h);
  scan_apply_kernel<<<tiles, SCAN_THREADS, 0, h->stream>>>(in, n, h->scan_tmp, out);
  LAUNCH_CHECK(h);
  return 0;
}

static int check_err_flag(phdslam* h) {
  /* red_host was filled by a preceding async copy + sync */
  if (h->red_host->err_flag & 4) {
    phdslam_set_error("peer exchange timed out: another rank did not reach the same step");
    return PHDSLAM_ERR_NCCL;
  }
  if (h->red_host->err_flag & 1) {
    phdslam_set_error("merge candidate buffer overflow: raise max_components (candidates after prune exceeded 2*max_components+256)");
    return PHDSLAM_ERR_CAPACITY;
  }
  if (h->red_host->err_flag & 2) {
    phdslam_set_error("a particle's map exceeded max_components");
    return PHDSLAM_ERR_CAPACITY;
  }
  if (h->red_host->err_flag & 8) {
    phdslam_set_error("a particle's dynamic map (or its prune survivors) exceeded max_components_dynamic");
    return PHDSLAM_ERR_CAPACITY;
  }
  if (h->red_host->err_ranks) {   /* all-reduced with the weight sum: every rank leaves the step with the same status */
    phdslam_set_error("another rank's update exceeded its map / candidate capacity");
    return PHDSLAM_ERR_CAPACITY;
  }
  return 0;
}

/* ---- predict ---- */
/* "shotgun" prediction (src/phdfilter.cu:1091, 796-797, 1185-1238): every particle becomes k particles (same map,
 * cardinality and resample index, log-weight - log k); the predict kernel then moves each of them with its own noise */
static int fan_out(phdslam* h) {
  const int k = h->cfg.n_predict_particles, n = h->n_local;
  const long long n_out = (long long)n * k;
  if (n_out > h->n_cap) {
    phdslam_set_error("n_predict_particles: particle count exceeds the capacity (resample before predicting again)");
    return PHDSLAM_ERR_CAPACITY;
  }
  const int b = h->cur;
  fanout_kernel<<<cdiv(n_out, 256), 256, 0, h->stream>>>((int)n_out, k, phd_safe_log((float)k), h->logw, h->dlogw, h->resample_idx,
                                                        h->n_in, h->ancestors);
  LAUNCH_CHECK(h);
  resample_gather_kernel<<<cdiv(n_out, 8), 256, 0, h->stream>>>(h->ancestors, (int)n_out, 0, n, (int)n_out, h->pose[b], h->pose[b ^ 1],
                                                              h->count[b], h->count[b ^ 1], h->map[b], h->map[b ^ 1],
                                                              h->card[b], h->card[b ^ 1], h->Cmax, h->n_card, 0, nullptr);
  LAUNCH_CHECK(h);
  std::swap(h->logw, h->dlogw);             /* both are n_cap floats; dlogw is per-update scratch */
  std::swap(h->resample_idx, h->n_in);      /* both are n_cap ints; n_in is per-update scratch */
  h->cur ^= 1;
  h->n_local = h->n_global = (int)n_out;
  h->totals_valid = 0;
  return 0;
}

extern "C" int phdslam_predict(phdslam_t* h, const float* control, const double* draws) {
  ENTER(h);
  if (h->cfg.n_predict_particles > 1) {
    int rcf = fan_out(h);
    if (rcf) return rcf;
  }
  const int n = h->n_local;
  const int per = (h->cfg.motion_type == 1) ? 2 : 3;
  const double* ddev = nullptr;
  if (draws) {
    size_t need = (size_t)n * per;
    if (h->draws_cap < need) {
      cudaFree(h->draws_dev);
      CK(cudaMalloc(&h->draws_dev, need * sizeof(double)));
      h->draws_cap = need;
    }
    CK(copy_h2d_async(h, h->draws_dev, draws, need * sizeof(double), h->stream));
    ddev = h->draws_dev;
  }
  float v_enc = control ? control[0] : 0.0f, alpha = control ? control[1] : 0.0f;
  CK(cudaEventRecord(h->ev[0], h->stream));
  predict_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->pose[h->cur], n, h->offset, v_enc, alpha, ddev, h->predict_calls, h->dc);
  LAUNCH_CHECK(h);
  if (h->Dmax) {
    /* predictMapMixed (src/phdfilter.cu:1240-1242, 965-1035): by the whole config.dt on every call, as the reference does */
    const phdslam_config_t& c = h->cfg;
    dyn_predict_kernel<<<cdiv((long long)n * h->Dmax, 256), 256, 0, h->stream>>>(
        h->dmap[h->dcur], h->dcount[h->dcur], n, h->Dmax, c.dt, c.std_ax_features * c.std_ax_features,
        c.std_ay_features * c.std_ay_features, c.ps, c.beta, c.tau);
    LAUNCH_CHECK(h);
  }
  CK(cudaEventRecord(h->ev[1], h->stream));
  h->predict_calls++;
  if (draws) CK(cudaStreamSynchronize(h->stream)); /* the caller may free `draws` on return */
  return 0;
}

/* ---- update ---- */
static int upload_measurements(phdslam* h, const float* z, int M, int fields) {
  /* pinned staging buffer owned by the handle: no synchronisation needed -- the previous update (the only other user) has
   * completed before this call (phdslam_update ends with a stream synchronisation) */
  float* zz = h->z_host;
  memset(zz, 0, 3 * PHD_MAX_MEAS * sizeof(float));
  for (int m = 0; m < M; ++m) {
    zz[m] = z[(size_t)m * fields];
    zz[PHD_MAX_MEAS + m] = z[(size_t)m * fields + 1];
    zz[2 * PHD_MAX_MEAS + m] = (fields > 2) ? z[(size_t)m * fields + 2] : 0.0f;
  }
  CK(copy_h2d_async(h, h->z_dev, zz, 3 * PHD_MAX_MEAS * sizeof(float), h->stream));
  return 0;
}

/* classification + dense-offset scan; leaves total/max padded term counts in red_host */
#define PHD_CAND_BUDGET (8ull << 30)
/* worst_case_ok: the caller can plan from an upper bound of the term counts (every particle with max_components in-range
 * components), so the counts need not come back to the host: red_host then holds that bound, not the measured values */
static int classify_and_scan(phdslam* h, int M, bool worst_case_ok = false) {
  const int n = h->n_local;
  CK(cudaMemsetAsync(h->red, 0, sizeof(Reductions), h->stream));
  classify_kernel<<<cdiv(n, 8), 256, 0, h->stream>>>(h->map[h->cur], h->count[h->cur], h->pose[h->cur], n, M, h->cls,
                                                     h->n_in, h->tpad, h->red, h->dc);
  LAUNCH_CHECK(h);
  int rc = scan_u64(h, h->tpad, n, h->toff, &h->red->total_terms);
  if (rc) return rc;
  if (worst_case_ok) {
    /* sized for the measurement count rounded up to a multiple of 64, so that a run whose sets grow from step to step
     * does not re-allocate the dense buffer every time */
    const unsigned Ma = (unsigned)std::min((M + 63) & ~63, PHD_MAX_MEAS);
    const unsigned long long tmax = (((unsigned long long)h->Cmax * (Ma + 1) + Ma) + 63ull) & ~63ull;
    const unsigned long long bound = tmax * (unsigned long long)n;
    const unsigned long long budget_terms = h->cfg.update_buffer_bytes / (PHD_NPLANES * 4);
    if (bound <= (1ull << 28) && bound <= budget_terms &&
        (unsigned long long)n * h->Smax * 64ull <= PHD_CAND_BUDGET) {   /* <= 7.5 GB of dense terms, one batch */
      memset(h->red_host, 0, sizeof(Reductions));
      h->red_host->total_terms = bound;
      h->red_host->max_terms = (int)tmax;
      return 0;
    }
  }
  CK(copy_d2h_async(h, h->red_host, h->red, sizeof(Reductions), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

static int ensure_cand(phdslam* h, size_t particles) {
  size_t need = particles * (size_t)h->Smax * 2;
  if (h->cand_cap >= need) return 0;
  cudaFree(h->cand);
  cudaFree(h->cand_in);
  h->cand = h->cand_in = nullptr;
  h->cand_cap = 0;
  CK(cudaMalloc(&h->cand, need * sizeof(float4)));
  CK(cudaMalloc(&h->cand_in, need * sizeof(float4)));
  h->cand_cap = need;
  return 0;
}

static int ensure_dense(phdslam* h, size_t floats) {
  if (h->dense_floats >= floats) return 0;
  cudaFree(h->dense);
  h->dense = nullptr;
  h->dense_floats = 0;
  CK(cudaMalloc(&h->dense, floats * sizeof(float)));
  h->dense_floats = floats;
  return 0;
}

/* Particles stream through the scratch buffers in batches [p0, p1): the dense update-term buffer (dense mode,
 * bounded by update_buffer_bytes) and the merge candidate records (bounded by PHD_CAND_BUDGET). */
static int plan_batches(phdslam* h, bool dense, std::vector<int>& bounds, size_t* max_batch_terms) {
  const int n = h->n_local;
  const unsigned long long total = h->red_host->total_terms;
  const unsigned long long budget_terms = std::max<unsigned long long>(h->cfg.update_buffer_bytes / (PHD_NPLANES * 4), 1);
  int per = (int)std::min<unsigned long long>((unsigned long long)n, std::max<unsigned long long>(PHD_CAND_BUDGET / ((size_t)h->Smax * 64), 1));
  *max_batch_terms = 0;
  if (dense) {
    const unsigned long long maxT = std::max(h->red_host->max_terms, 1);
    if (maxT > budget_terms) {
      phdslam_set_error("update_buffer_bytes is smaller than one particle's update terms");
      return PHDSLAM_ERR_INVALID;
    }
    if (total <= budget_terms && per >= n) {
      *max_batch_terms = (size_t)total;
    } else {
      per = (int)std::min<unsigned long long>((unsigned long long)per, std::max<unsigned long long>(budget_terms / maxT, 1));
      *max_batch_terms = (size_t)std::min<unsigned long long>((unsigned long long)per * maxT, total);
    }
  }
  bounds.clear();
  for (int p = 0; p < n; p += per) bounds.push_back(p);
  bounds.push_back(n);
  return 0;
}

/* cand_p0: local particle whose records sit at the start of the candidate buffers (p0 when the buffers hold one batch,
 * 0 when they hold every local particle) */
static int launch_update_batch(phdslam* h, int M, int p0, int p1, unsigned long long tbase, bool dense, int write_card = 1,
                               int cand_p0 = -1) {
  if (cand_p0 < 0) cand_p0 = p0;
  UpdArgs a;
  a.lfact = h->lfact; a.card = h->n_card ? h->card[h->cur] : nullptr; a.write_card = write_card;
  a.map = h->map[h->cur]; a.count = h->count[h->cur]; a.cls = h->cls; a.pose = h->pose[h->cur];
  a.z = h->z_dev; a.M = M; a.n = h->n_local; a.p0 = p0;
  a.toff = h->toff; a.tbase = tbase; a.dense = h->dense; a.n_in = h->n_in; a.dlogw = h->dlogw; a.c = h->dc;
  a.cand = h->cand_in + (size_t)(p0 - cand_p0) * h->Smax * 2; a.n_cand = h->n_cand; a.Smax = h->Smax;
  a.mix_dsum = h->mix_dsum; a.mix_nhat = h->mix_nhat; a.mix_L = h->mix_L;
  if (h->Dmax) {
    if (dense)
      update_mixed_kernel<true><<<p1 - p0, UPD_THREADS, update_smem_bytes(h->Cmax), h->stream>>>(a);
    else
      update_mixed_kernel<false><<<p1 - p0, UPD_THREADS, update_smem_bytes(h->Cmax), h->stream>>>(a);
  } else if (h->n_card) {
    /* the fused mode keeps one chunk maximum per (measurement, 64-component chunk) behind the multi-object tables */
    const size_t sm = update_smem_bytes(h->Cmax) + cphd_smem_bytes(h->n_card, M) + (dense ? 0 : cphd_chunkmax_bytes(M, h->Cmax));
    if (dense)
      update_kernel<true, true><<<p1 - p0, UPD_THREADS, sm, h->stream>>>(a);
    else
      update_kernel<false, true><<<p1 - p0, UPD_THREADS, sm, h->stream>>>(a);
  } else if (dense) {
    update_kernel<true, false><<<p1 - p0, UPD_THREADS, update_smem_bytes(h->Cmax), h->stream>>>(a);
  } else {
    update_kernel<false, false><<<p1 - p0, UPD_THREADS, update_smem_bytes(h->Cmax), h->stream>>>(a);
  }
  LAUNCH_CHECK(h);
  return 0;
}

static int launch_merge_batch(phdslam* h, int M, int p0, int p1, cudaStream_t st = nullptr, int cand_p0 = -1) {
  if (!st) st = h->stream;
  if (cand_p0 < 0) cand_p0 = p0;
  MrgArgs a;
  a.M = M; a.n = h->n_local; a.p0 = p0; a.n_cand = h->n_cand;
  a.cand_in = h->cand_in + (size_t)(p0 - cand_p0) * h->Smax * 2;
  a.map_in = h->map[h->cur]; a.count_in = h->count[h->cur]; a.cls = h->cls;
  a.map_out = h->map[h->cur ^ 1]; a.count_out = h->count[h->cur ^ 1];
  a.red = h->red; a.Smax = h->Smax; a.c = h->dc; a.p1 = p1;
  a.cand = h->cand + (size_t)(p0 - cand_p0) * h->Smax * 2;
  a.n_in = h->n_in;
  a.Scap = h->Scap; a.ovf_list = h->ovf_list; a.use_list = 0;
  if (h->Scap > 0 && h->dc.distance_metric == 0) {
    merge_fast_kernel<<<p1 - p0, MF_THREADS, merge_fast_smem_bytes(h->Scap), st>>>(a);
    LAUNCH_CHECK(h);
    a.use_list = 1;   /* whatever did not fit the shared-memory capacity (normally nothing: the warps exit at once) */
    merge_kernel<<<cdiv(p1 - p0, MRG_WARPS), MRG_THREADS, merge_smem_bytes(h->Smax), st>>>(a);
    LAUNCH_CHECK(h);
    CK(cudaMemsetAsync(&h->red->ovf_n, 0, sizeof(int), st));
    return 0;
  }
  merge_kernel<<<cdiv(p1 - p0, MRG_WARPS), MRG_THREADS, merge_smem_bytes(h->Smax), st>>>(a);
  LAUNCH_CHECK(h);
  return 0;
}

/* mixed feature model: arguments of the dynamic-map kernels (csrc/mixed.cuh) */
static DynArgs dyn_args(phdslam* h, int M) {
  const phdslam_config_t& c = h->cfg;
  DynArgs a;
  a.dmap_in = h->dmap[h->dcur]; a.dcount_in = h->dcount[h->dcur];
  a.dmap_out = h->dmap[h->dcur ^ 1]; a.dcount_out = h->dcount[h->dcur ^ 1];
  a.pose = h->pose[h->cur]; a.z = h->z_dev; a.M = M; a.n = h->n_local; a.Dmax = h->Dmax; a.Sd = h->Sd;
  a.dsum = h->mix_dsum; a.nhat = h->mix_nhat; a.L = h->mix_L; a.dlogw = h->dlogw; a.cand = h->dcand; a.red = h->red;
  a.dt = c.dt; a.var_x = c.std_ax_features * c.std_ax_features; a.var_y = c.std_ay_features * c.std_ay_features;
  a.ps = c.ps; a.beta = c.beta; a.tau = c.tau; a.cov_vx = c.cov_vx_birth; a.cov_vy = c.cov_vy_birth;
  a.c = h->dc;
  return a;
}

/* w += dw; normalise (src/phdfilter.cu:3735-3755) */
static int update_weights(phdslam* h, bool add) {
  const int n = h->n_local;
  /* only the weight accumulators are reset: err_flag / max_cand of the update and merge kernels stay for the host */
  CK(cudaMemsetAsync(h->red, 0, offsetof(Reductions, err_ranks) + sizeof(unsigned long long), h->stream));
  weights_add_max_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->logw, add ? h->dlogw : nullptr, n, h->red);
  LAUNCH_CHECK(h);
  if (h->mbox) { /* {max_key, pad0 = 0} as one 64-bit word, combined by max */
    int rc = mbox_exchange(h, reinterpret_cast<unsigned long long*>(&h->red->max_key), 1, 1u, false);
    if (rc) return rc;
  } else if (h->world > 1) /* global max of the log-weights (ordered-uint keys: max is exact and order independent) */
    CKN(ncclAllReduce(&h->red->max_key, &h->red->max_key, 1, ncclUint32, ncclMax, (ncclComm_t)h->nccl_comm, h->stream));
  weights_sum_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->logw, n, h->red);
  LAUNCH_CHECK(h);
  if (h->mbox) {
    int rc = mbox_exchange(h, &h->red->sum_fx, 3, 0u, false);
    if (rc) return rc;
  } else if (h->world > 1) /* integer (Q36) sum: identical on every rank for any GPU count; NaN and error counts ride along */
    CKN(ncclAllReduce(&h->red->sum_fx, &h->red->sum_fx, 3, ncclUint64, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
  weights_normalise_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->logw, n, h->red);
  LAUNCH_CHECK(h);
  return 0;
}

extern "C" int phdslam_update(phdslam_t* h, const float* z, int M, int fields) {
  ENTER(h);
  if (M <= 0) return 0;                 /* main.cpp:1258 */
  h->totals_valid = 0;
  if (fields != 2 && fields != 3) return PHDSLAM_ERR_INVALID;
  if (M > PHD_MAX_MEAS) M = PHD_MAX_MEAS; /* :3390-3394 */
  int rc = upload_measurements(h, z, M, fields);
  if (rc) return rc;
  CK(cudaEventRecord(h->ev[2], h->stream));
  rc = classify_and_scan(h, M, /*worst_case_ok=*/true);   /* small particle sets: no host round trip for the term counts */
  if (rc) return rc;
  std::vector<int> bounds;
  size_t max_terms = 0;
  const bool dense = (h->cfg.update_mode == 0);
  rc = plan_batches(h, dense, bounds, &max_terms);
  if (rc) return rc;
  if (dense) {
    rc = ensure_dense(h, max_terms * PHD_NPLANES);
    if (rc) return rc;
  }
  /* overlap needs candidate buffers for every local particle (the merge of one sub-batch reads them while the update of
   * the next one writes) */
  const bool overlap = h->overlap && ((unsigned long long)h->n_local * h->Smax * 64ull <= PHD_CAND_BUDGET);
  {
    int maxb = 0;
    for (size_t b = 0; b + 1 < bounds.size(); ++b) maxb = std::max(maxb, bounds[b + 1] - bounds[b]);
    rc = ensure_cand(h, overlap ? (size_t)h->n_local : (size_t)maxb);
    if (rc) return rc;

  }
  std::vector<unsigned long long> tb(bounds.size(), 0);
  if (bounds.size() > 2) {
    for (size_t b = 0; b + 1 < bounds.size(); ++b)
      CK(copy_d2h_async(h, &tb[b], h->toff + bounds[b], sizeof(unsigned long long), h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  CK(cudaMemsetAsync(h->red, 0, sizeof(Reductions), h->stream));
  if (h->Dmax) {   /* what the dynamic features add to the normalisers and the predicted cardinality of the static update */
    CK(cudaEventRecord(h->ev_dyn[0], h->stream));
    dyn_pre_kernel<<<h->n_local, DYN_THREADS, dyn_smem_bytes(h->Dmax, M, h->Sd), h->stream>>>(dyn_args(h, M));
    LAUNCH_CHECK(h);
    CK(cudaEventRecord(h->ev_dyn[1], h->stream));
  }
  float upd_ms = 0, mrg_ms = 0;
  bool multi = bounds.size() > 2;
  if (overlap) {
    /* sub-batches: the update of sub-batch k+1 (this stream, HBM-bound) runs under the merge of sub-batch k (the
     * higher-priority stream, issue-bound).  Sub-batches of one dense batch share its dense buffer. */
    const int n = h->n_local;
    const int n_dense = (int)bounds.size() - 1;
    int per_sub = std::max(2048, cdiv(n, PHD_MAX_SUB / 4));
    while (cdiv(n, per_sub) + n_dense > PHD_MAX_SUB) per_sub *= 2;
    CK(cudaEventRecord(h->ev[3], h->stream));
    CK(cudaStreamWaitEvent(h->stream_m, h->ev[3], 0));           /* the merge stream starts behind the classification */
    int k = 0;
    for (int b = 0; b < n_dense; ++b) {
      for (int p0 = bounds[b]; p0 < bounds[b + 1]; p0 += per_sub, ++k) {
        const int p1 = std::min(p0 + per_sub, bounds[b + 1]);
        rc = launch_update_batch(h, M, p0, p1, tb[b], dense, 1, 0);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev_sub[k], h->stream));
        CK(cudaStreamWaitEvent(h->stream_m, h->ev_sub[k], 0));
        rc = launch_merge_batch(h, M, p0, p1, h->stream_m, 0);
        if (rc) return rc;
      }
    }
    CK(cudaEventRecord(h->ev[4], h->stream));                     /* end of the last update kernel */
    CK(cudaEventRecord(h->ev_merge_done, h->stream_m));
    CK(cudaStreamWaitEvent(h->stream, h->ev_merge_done, 0));
    CK(cudaEventRecord(h->ev[5], h->stream));                     /* end of the last merge kernel */
    multi = false;                                                /* one pair of spans: update, then the exposed merge tail */
  } else
  for (size_t b = 0; b + 1 < bounds.size(); ++b) {
    CK(cudaEventRecord(h->ev[3], h->stream));
    rc = launch_update_batch(h, M, bounds[b], bounds[b + 1], tb[b], dense);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev[4], h->stream));
    rc = launch_merge_batch(h, M, bounds[b], bounds[b + 1]);
    if (rc) return rc;
    CK(cudaEventRecord(h->ev[5], h->stream));
    if (multi) {
      CK(cudaEventSynchronize(h->ev[5]));
      float t1, t2;
      cudaEventElapsedTime(&t1, h->ev[3], h->ev[4]);
      cudaEventElapsedTime(&t2, h->ev[4], h->ev[5]);
      upd_ms += t1; mrg_ms += t2;
    }
  }
  if (h->Dmax) {   /* the dynamic map: update terms, prune, merge (after the static update wrote the normalisers) */
    CK(cudaEventRecord(h->ev_dyn[2], h->stream));
    dyn_update_kernel<<<h->n_local, DYN_THREADS, dyn_smem_bytes(h->Dmax, M, h->Sd), h->stream>>>(dyn_args(h, M));
    LAUNCH_CHECK(h);
    CK(cudaEventRecord(h->ev_dyn[3], h->stream));
  }
  rc = update_weights(h, true);
  if (rc) return rc;
  CK(copy_d2h_async(h, h->red_host, h->red, sizeof(Reductions), h->stream));
  CK(cudaEventRecord(h->ev[6], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->Dmax) h->dcur ^= 1;
  h->cur ^= 1; /* merged maps become the front buffer; poses and weights are single-buffered in place */
  /* pose planes are not touched by the update: keep the front pose buffer consistent with `cur` */
  std::swap(h->pose[0], h->pose[1]);
  std::swap(h->card[0], h->card[1]);
  h->slab_sw ^= 1;   /* pose[i] / card[i] now live in slot i ^ slab_sw of the peer window (same on every rank) */
  if (!multi) {
    cudaEventElapsedTime(&upd_ms, h->ev[3], h->ev[4]);
    cudaEventElapsedTime(&mrg_ms, h->ev[4], h->ev[5]);
  }
  h->tim.update_ms = upd_ms;
  h->tim.merge_ms = mrg_ms;
  cudaEventElapsedTime(&h->tim.weights_ms, h->Dmax ? h->ev_dyn[3] : h->ev[5], h->ev[6]);
  if (h->Dmax) {
    float t1 = 0, t2 = 0;
    cudaEventElapsedTime(&t1, h->ev_dyn[0], h->ev_dyn[1]);
    cudaEventElapsedTime(&t2, h->ev_dyn[2], h->ev_dyn[3]);
    h->tim.dynamic_ms = t1 + t2;
  }
  if (!h->Scap_pinned) {
    const int mc = h->red_host->max_cand;
    h->Scap = std::min(h->Scap_max, std::max(64, (mc + mc / 16 + 8 + 31) & ~31));
  }
  rc = check_err_flag(h);
  if (rc) return rc;
  h->nan_seen = h->red_host->nan_count != 0;
  if (h->nan_seen) {              /* the reference notices through nEff at the end of the iteration (main.cpp:1307-1311) */
    phdslam_set_error("nan weights detected");
    return PHDSLAM_ERR_NAN;
  }
  return 0;
}

extern "C" int phdslam_update_terms(phdslam_t* h, const float* z, int M, int fields, phdslam_gaussian2d_t* terms_out,
                                    size_t cap, int* n_in_range_out, float* dlogw_out) {
  ENTER(h);
  if (h->Dmax) {
    phdslam_set_error("phdslam_update_terms: the dense-terms query covers the static feature model only");
    return PHDSLAM_ERR_INVALID;
  }
  if (M <= 0 || (fields != 2 && fields != 3)) return PHDSLAM_ERR_INVALID;
  if (M > PHD_MAX_MEAS) M = PHD_MAX_MEAS;
  int rc = upload_measurements(h, z, M, fields);
  if (rc) return rc;
  rc = classify_and_scan(h, M);
  if (rc) return rc;
  const int n = h->n_local;
  const size_t total = (size_t)h->red_host->total_terms;
  if (total * PHD_NPLANES * 4 > h->cfg.update_buffer_bytes) {
    phdslam_set_error("phdslam_update_terms needs the whole dense result to fit update_buffer_bytes");
    return PHDSLAM_ERR_INVALID;
  }
  rc = ensure_dense(h, total * PHD_NPLANES);
  if (rc) return rc;
  CK(cudaEventRecord(h->ev[3], h->stream));
  rc = ensure_cand(h, (size_t)n);
  if (rc) return rc;
  rc = launch_update_batch(h, M, 0, n, 0, true, /*write_card=*/0);   /* a query: the cardinality is not advanced */
  if (rc) return rc;
  CK(cudaEventRecord(h->ev[4], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->tim.update_ms, h->ev[3], h->ev[4]);
  std::vector<int> nin(n);
  CK(copy_d2h(h, nin.data(), h->n_in, (size_t)n * sizeof(int)));
  if (n_in_range_out) memcpy(n_in_range_out, nin.data(), (size_t)n * sizeof(int));
  if (dlogw_out) CK(copy_d2h(h, dlogw_out, h->dlogw, (size_t)n * sizeof(float)));
  if (terms_out) {
    std::vector<unsigned long long> ooff(n + 1, 0);
    for (int p = 0; p < n; ++p) ooff[p + 1] = ooff[p] + (unsigned long long)nin[p] * (M + 1) + M;
    if (ooff[n] > cap) {
      phdslam_set_error("terms_out too small");
      return PHDSLAM_ERR_INVALID;
    }
    unsigned long long* d_off = nullptr;
    phdslam_gaussian2d_t* d_out = nullptr;
    CK(cudaMalloc(&d_off, (n + 1) * sizeof(unsigned long long)));
    CK(cudaMalloc(&d_out, std::max<size_t>(ooff[n], 1) * sizeof(phdslam_gaussian2d_t)));
    CK(copy_h2d(h, d_off, ooff.data(), (n + 1) * sizeof(unsigned long long)));
    dense_export_kernel<<<n, 256, 0, h->stream>>>(h->dense, h->toff, 0, h->n_in, M, 0, n, d_off, d_out);
    LAUNCH_CHECK(h);
    CK(cudaStreamSynchronize(h->stream));
    CK(copy_d2h(h, terms_out, d_out, ooff[n] * sizeof(phdslam_gaussian2d_t)));
    cudaFree(d_off);
    cudaFree(d_out);
  }
  return 0;
}

/* ---- estimate ---- */
extern "C" int phdslam_estimate(phdslam_t* h, phdslam_estimate_t* out) {
  ENTER(h);
  const int n = h->n_local;
  CK(cudaEventRecord(h->ev[7], h->stream));
  CK(cudaMemsetAsync(h->red, 0, sizeof(Reductions), h->stream));
  estimate_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->logw, h->pose[h->cur], n, h->offset, h->red);
  LAUNCH_CHECK(h);
  h->totals_valid = 0;
  if (h->mbox) {
    /* neff_fx, pose_fx[6], argmax_key, cdf_total are nine adjacent 64-bit integers: ONE exchange (sums, the arg-max key by
     * max); the raw records also tell every rank every rank's resampling CDF total, so a resampling that follows needs
     * no exchange of its own before it can plan the migration */
    int rc = mbox_exchange(h, &h->red->neff_fx, 9, 1u << 7, true);
    if (rc) return rc;
    CK(copy_d2h_async(h, h->gath_host, h->gath_dev, (size_t)h->world * MBOX_WORDS * sizeof(unsigned long long), h->stream));
  } else if (h->world > 1) {
    /* neff_fx and pose_fx[6] are adjacent 64-bit integers: one exact integer all-reduce; arg-max key by max */
    CKN(ncclAllReduce(&h->red->neff_fx, &h->red->neff_fx, 7, ncclUint64, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
    CKN(ncclAllReduce(&h->red->argmax_key, &h->red->argmax_key, 1, ncclUint64, ncclMax, (ncclComm_t)h->nccl_comm, h->stream));
  }
  CK(copy_d2h_async(h, h->red_host, h->red, sizeof(Reductions), h->stream));
  CK(cudaEventRecord(h->ev[8], h->stream));
  CK(cudaStreamSynchronize(h->stream));
  cudaEventElapsedTime(&h->tim.estimate_ms, h->ev[7], h->ev[8]);
  if (h->red_host->err_flag & 4) {
    phdslam_set_error("peer exchange timed out: another rank did not reach the estimate");
    return PHDSLAM_ERR_NCCL;
  }
  if (h->mbox) {
    for (int r = 0; r < h->world; ++r) h->totals_host[r] = h->gath_host[(size_t)r * MBOX_WORDS + 8];
    h->totals_valid = 1;
  } else if (h->world == 1) {
    h->totals_host[0] = h->red_host->cdf_total;      /* the resampling CDF total of the current weights (estimate_kernel) */
    h->totals_valid = 1;
  }
  const Reductions& r = *h->red_host;
  const double inv = 1.0 / (double)(1ull << PHD_FX_POSE_BITS);
  float* e = &out->expected_pose.px;
  for (int k = 0; k < 6; ++k) e[k] = (float)((double)r.pose_fx[k] * inv);
  if (r.argmax_key) {
    out->map_particle = (int)(0xffffffffu - (unsigned)(r.argmax_key & 0xffffffffull));
    out->max_log_weight = ordered_uint_to_float((uint32_t)(r.argmax_key >> 32));
  } else {
    out->map_particle = -1;
    out->max_log_weight = -FLT_MAX;
  }
  double s2 = (double)r.neff_fx * (1.0 / (double)(1ull << PHD_FX_NEFF_BITS));
  out->neff = (float)(1.0 / s2 / (double)h->n_global);
  if (h->n_global == 1 && h->world == 1) { /* main.cpp:381-387 */
    std::vector<float> p(6);
    for (int k = 0; k < 6; ++k) CK(copy_d2h(h, &p[k], h->pose[h->cur] + k, sizeof(float)));
    memcpy(&out->expected_pose, p.data(), 6 * sizeof(float));
    out->map_particle = 0;
  }
  return 0;
}

/* ---- resample ---- */
static int ensure_migration(phdslam* h, size_t records) {
  if (!h->mig_pose_in) CK(cudaMalloc(&h->mig_pose_in, (size_t)h->n_local * 6 * sizeof(float)));
  if (!h->totals_dev) CK(cudaMalloc(&h->totals_dev, (size_t)h->world * sizeof(unsigned long long)));
  if (h->mig_cap >= records) return 0;
  cudaFree(h->mig_map); cudaFree(h->mig_pose); cudaFree(h->mig_count); cudaFree(h->mig_anc); cudaFree(h->mig_card);
  h->mig_map = nullptr; h->mig_pose = nullptr; h->mig_count = nullptr; h->mig_anc = nullptr; h->mig_card = nullptr;
  h->mig_cap = 0;
  const size_t row = (size_t)PHD_MAP_PLANES * h->Cmax;
  CK(cudaMalloc(&h->mig_map, records * row * sizeof(float)));
  CK(cudaMalloc(&h->mig_pose, records * 6 * sizeof(float)));
  CK(cudaMalloc(&h->mig_count, records * sizeof(int)));
  CK(cudaMalloc(&h->mig_anc, records * sizeof(int)));
  if (h->n_card) CK(cudaMalloc(&h->mig_card, records * h->n_card * sizeof(float)));
  h->mig_cap = records;
  return 0;
}

extern "C" int phdslam_resample(phdslam_t* h, int n_new, const double* uniforms, int* ancestors_out) {
  ENTER(h);
  const int n = h->n_local;
  if (n_new < 0) n_new = h->n_global;
  if (n_new != h->n_global && (h->world > 1 || n_new > h->n_cap || n_new < 1)) {
    phdslam_set_error("resampling to a different particle count needs a single GPU and n_new within the particle capacity");
    return PHDSLAM_ERR_INVALID;
  }
  const double* udev = nullptr;
  if (uniforms) {
    size_t need = (size_t)n_new + 1;
    if (h->draws_cap < need) {
      cudaFree(h->draws_dev);
      CK(cudaMalloc(&h->draws_dev, need * sizeof(double)));
      h->draws_cap = need;
    }
    CK(copy_h2d_async(h, h->draws_dev, uniforms, need * sizeof(double), h->stream));
    udev = h->draws_dev;
  }
  const int sysmode = (h->cfg.resample_mode == 1);
  CK(cudaEventRecord(h->ev[9], h->stream));
  resample_weights_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->logw, n, h->q_fx);
  LAUNCH_CHECK(h);
  CK(cudaMemsetAsync(h->red, 0, sizeof(Reductions), h->stream));
  int rc = scan_u64(h, h->q_fx, n, h->cdf_excl, &h->red->cdf_total);
  if (rc) return rc;
  const int b = h->cur;
  unsigned long long total = 0, base = 0;
  std::vector<int> bounds;
  std::vector<unsigned long long> totals_plan;
  if (h->world > 1) {
    /* every rank learns every rank's integer weight total: its CDF offset and the global total follow.  After an
     * estimate on the same weights the totals are already on the host (they rode on the estimate's exchange). */
    rc = ensure_migration(h, 1);
    if (rc) return rc;
    std::vector<unsigned long long> totals(h->world);
    if (h->mbox && h->totals_valid) {
      for (int r = 0; r < h->world; ++r) totals[r] = h->totals_host[r];
    } else {
      if (h->mbox) {
        rc = mbox_exchange(h, &h->red->cdf_total, 1, 0u, true);
        if (rc) return rc;
        CK(copy_d2h_async(h, h->gath_host, h->gath_dev, (size_t)h->world * MBOX_WORDS * sizeof(unsigned long long), h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int r = 0; r < h->world; ++r) totals[r] = h->gath_host[(size_t)r * MBOX_WORDS];
      } else {
        CKN(ncclAllGather(&h->red->cdf_total, h->totals_dev, 1, ncclUint64, (ncclComm_t)h->nccl_comm, h->stream));
        CK(copy_d2h_async(h, totals.data(), h->totals_dev, (size_t)h->world * sizeof(unsigned long long), h->stream));
        CK(cudaStreamSynchronize(h->stream));
      }
    }
    for (int r = 0; r < h->world; ++r) {
      if (r < h->rank) base += totals[r];
      total += totals[r];
    }
    if (total == 0) {
      phdslam_set_error("all particle weights are zero or NaN");
      return PHDSLAM_ERR_NAN;
    }
    totals_plan = totals;        /* the migration is planned below, on the host, while the local gather runs */
  } else {
    if (h->totals_valid) {
      total = h->totals_host[0];            /* known since the estimate: no host round trip */
    } else {
      CK(copy_d2h_async(h, h->red_host, h->red, sizeof(Reductions), h->stream));
      CK(cudaStreamSynchronize(h->stream));
      total = h->red_host->cdf_total;
    }
    if (total == 0) {
      phdslam_set_error("all particle weights are zero or NaN");
      return PHDSLAM_ERR_NAN;
    }
  }
  /* offspring of this rank whose ancestor is local (remote ones get -1 and are skipped by the gather) */
  /* offspring owned by this rank: n (= its share of n_new), or all n_new on a single GPU (which may differ from n:
   * the down-sampling after "shotgun" predictions, src/main.cpp:1286-1289) */
  const int n_off = (h->world > 1) ? n : n_new;
  resample_search_kernel<<<cdiv(n_off, 256), 256, 0, h->stream>>>(h->cdf_excl, n, base, total, n_new, h->offset, n_off, h->offset, udev,
                                                                sysmode, h->resample_calls, h->dc.seed_lo, h->dc.seed_hi, h->ancestors);
  LAUNCH_CHECK(h);
  resample_gather_kernel<<<cdiv(n_off, 8), 256, 0, h->stream>>>(h->ancestors, n_off, h->offset, n, n_off, h->pose[b], h->pose[b ^ 1],
                                                              h->count[b], h->count[b ^ 1], h->map[b], h->map[b ^ 1],
                                                              h->card[b], h->card[b ^ 1], h->Cmax, h->n_card, 0, nullptr);
  LAUNCH_CHECK(h);
  if (h->Dmax && h->world == 1) {   /* copy_particles carries maps_dynamic too (src/slamtypes.h:324) */
    dyn_gather_kernel<<<n_off, 64, 0, h->stream>>>(h->ancestors, n_off, 0, n, h->dmap[h->dcur], h->dcount[h->dcur],
                                                   h->dmap[h->dcur ^ 1], h->dcount[h->dcur ^ 1], h->Dmax);
    LAUNCH_CHECK(h);
  }
  if (h->world > 1) {
    bounds.resize(h->world + 1);
    rc = phdslam_plan_migration(h->world, totals_plan.data(), n_new, uniforms, h->cfg.resample_mode, h->resample_calls, h->cfg.seed,
                                bounds.data());
    if (rc) return rc;
  }
  if (h->world > 1 && h->p2p) {
    /* NVLink exchange: for every peer d, the offspring interval d owns whose ancestors live here (both ends derive it
     * from `bounds`, nothing is negotiated) is searched on the local CDF and pushed by the gather kernel straight into
     * d's back buffer through the mapped peer window.  d's back buffer is free: the all-gather above completed, so d
     * has finished its merge.  The all-reduce at the end is the barrier that makes every push visible to its owner. */
    const int me = h->rank, W = h->world;
    PushArgs pa;
    memset(&pa, 0, sizeof(pa));
    int n_e = 0, tot = 0;
    for (int k = 1; k < W; ++k) {
      const int d = (me + k) % W, sr = (me - k + W) % W;
      const int off_d = rank_offset(h, d), end_d = rank_offset(h, d + 1);
      const int out_lo = std::max(bounds[me], off_d), out_hi = std::min(bounds[me + 1], end_d);
      const int cnt_out = std::max(out_hi - out_lo, 0);
      const int in_lo = std::max(bounds[sr], h->offset), in_hi = std::min(bounds[sr + 1], h->offset + n);
      h->tim.migrated_in += (unsigned long long)std::max(in_hi - in_lo, 0);
      if (cnt_out <= 0) continue;
      const size_t nd = (size_t)(end_d - off_d);
      const SlabLayout L = slab_layout(nd, (size_t)h->Cmax, (size_t)h->n_card);
      unsigned char* pb = h->peer_base[d];
      PushDst& D = pa.d[n_e++];
      D.pose = reinterpret_cast<float*>(pb + L.pose[b ^ 1 ^ h->slab_sw]);
      D.count = reinterpret_cast<int*>(pb + L.count[b ^ 1]);
      D.map = reinterpret_cast<float*>(pb + L.map[b ^ 1]);
      D.card = h->n_card ? reinterpret_cast<float*>(pb + L.card[b ^ 1 ^ h->slab_sw]) : nullptr;
      D.anc_in = reinterpret_cast<int*>(pb + L.anc_in);
      D.n_dst = (int)nd; D.dst_first = out_lo - off_d; D.j0 = out_lo; D.first = tot;
      tot += cnt_out;
    }
    if (tot > 0) {          /* every peer's interval in ONE launch: ancestor search + push over NVLink */
      pa.n_dst_entries = n_e; pa.total = tot;
      pa.excl = h->cdf_excl; pa.n = n; pa.cdf_base = base; pa.cdf_total = total; pa.n_new = n_new; pa.anc_offset = h->offset;
      pa.uniforms = udev; pa.systematic = sysmode; pa.call = h->resample_calls; pa.seed_lo = h->dc.seed_lo; pa.seed_hi = h->dc.seed_hi;
      pa.pose_in = h->pose[b]; pa.count_in = h->count[b]; pa.map_in = h->map[b]; pa.card_in = h->card[b];
      pa.Cmax = h->Cmax; pa.n_card = h->n_card;
      resample_push_kernel<<<cdiv(tot, 8), 256, 0, h->stream>>>(pa);
      LAUNCH_CHECK(h);
    }
    if (h->mbox) {          /* "all pushes have landed": every rank's pushes precede its record (stream order + system fence) */
      rc = mbox_exchange(h, nullptr, 0, 0u, false);
      if (rc) return rc;
    } else
      CKN(ncclAllReduce(h->barrier_dev, h->barrier_dev + 1, 1, ncclInt32, ncclSum, (ncclComm_t)h->nccl_comm, h->stream));
    resample_take_pushed_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->ancestors, h->anc_in, n);
    LAUNCH_CHECK(h);
  } else if (h->world > 1) {
    /* migration: ring of shifts; in shift k this rank serves rank+k and is served by rank-k.  Both ends derive
     * the same offspring interval from `bounds`, so no counts are exchanged. */
    ncclComm_t comm = (ncclComm_t)h->nccl_comm;
    const size_t row = (size_t)PHD_MAP_PLANES * h->Cmax;
    const int me = h->rank, W = h->world;
    for (int k = 1; k < W; ++k) {
      const int d = (me + k) % W, sr = (me - k + W) % W;
      const int off_d = rank_offset(h, d), end_d = rank_offset(h, d + 1);
      const int out_lo = std::max(bounds[me], off_d), out_hi = std::min(bounds[me + 1], end_d);
      const int cnt_out = std::max(out_hi - out_lo, 0);
      const int in_lo = std::max(bounds[sr], h->offset), in_hi = std::min(bounds[sr + 1], h->offset + n);
      const int cnt_in = std::max(in_hi - in_lo, 0);
      if (cnt_out > 0) {
        rc = ensure_migration(h, (size_t)cnt_out);
        if (rc) return rc;
        resample_search_kernel<<<cdiv(cnt_out, 256), 256, 0, h->stream>>>(h->cdf_excl, n, base, total, n_new, out_lo, cnt_out, h->offset,
                                                                        udev, sysmode, h->resample_calls, h->dc.seed_lo,
                                                                        h->dc.seed_hi, h->mig_anc);
        LAUNCH_CHECK(h);
        migrate_pack_kernel<<<cdiv(cnt_out, 8), 256, 0, h->stream>>>(h->mig_anc, cnt_out, h->offset, n, h->pose[b], h->count[b],
                                                                    h->map[b], h->card[b], h->Cmax, h->n_card, h->mig_pose,
                                                                    h->mig_count, h->mig_map, h->mig_card);
        LAUNCH_CHECK(h);
      }
      if (cnt_out > 0 || cnt_in > 0) {
        CKN(ncclGroupStart());
        if (cnt_out > 0) {
          CKN(ncclSend(h->mig_pose, (size_t)cnt_out * 6, ncclFloat, d, comm, h->stream));
          CKN(ncclSend(h->mig_count, (size_t)cnt_out, ncclInt32, d, comm, h->stream));
          CKN(ncclSend(h->mig_anc, (size_t)cnt_out, ncclInt32, d, comm, h->stream));
          CKN(ncclSend(h->mig_map, (size_t)cnt_out * row, ncclFloat, d, comm, h->stream));
          if (h->n_card) CKN(ncclSend(h->mig_card, (size_t)cnt_out * h->n_card, ncclFloat, d, comm, h->stream));
        }
        if (cnt_in > 0) {
          const size_t first = (size_t)(in_lo - h->offset);
          CKN(ncclRecv(h->mig_pose_in, (size_t)cnt_in * 6, ncclFloat, sr, comm, h->stream));
          CKN(ncclRecv(h->count[b ^ 1] + first, (size_t)cnt_in, ncclInt32, sr, comm, h->stream));
          CKN(ncclRecv(h->ancestors + first, (size_t)cnt_in, ncclInt32, sr, comm, h->stream));
          CKN(ncclRecv(h->map[b ^ 1] + first * row, (size_t)cnt_in * row, ncclFloat, sr, comm, h->stream));
          if (h->n_card) CKN(ncclRecv(h->card[b ^ 1] + first * h->n_card, (size_t)cnt_in * h->n_card, ncclFloat, sr, comm, h->stream));
        }
        CKN(ncclGroupEnd());
        if (cnt_in > 0) {
          migrate_unpack_pose_kernel<<<cdiv((long long)cnt_in * 6, 256), 256, 0, h->stream>>>(h->mig_pose_in, cnt_in, in_lo - h->offset, n,
                                                                                           h->pose[b ^ 1]);
          LAUNCH_CHECK(h);
        }
      }
      h->tim.migrated_in += (unsigned long long)cnt_in;
    }
  }
  if (h->Dmax && h->world > 1) {
    /* Mixed feature model, sharded: the dynamic maps follow their offspring.  h->ancestors now holds the GLOBAL ancestor
     * of every local offspring (both exchange paths): offspring with a local ancestor copy its dynamic map here; for the
     * others the ring below repeats the shape of the static ring -- in shift k this rank serves rank+k and is served by
     * rank-k, both ends derive the offspring interval from `bounds`, the serving side searches the interval's ancestors
     * again, packs their dynamic maps and sends them; they land directly in the owner's back buffer.  (The dynamic maps are
     * not in the NVLink peer window: NCCL moves them in both exchange modes.) */
    ncclComm_t comm = (ncclComm_t)h->nccl_comm;
    const size_t per = (size_t)DYN_PLANES * h->Dmax;
    const int me = h->rank, W = h->world;
    const int db = h->dcur;
    dyn_gather_kernel<<<n_off, 64, 0, h->stream>>>(h->ancestors, n_off, h->offset, n, h->dmap[db], h->dcount[db], h->dmap[db ^ 1],
                                                   h->dcount[db ^ 1], h->Dmax);
    LAUNCH_CHECK(h);
    for (int k = 1; k < W; ++k) {
      const int d = (me + k) % W, sr = (me - k + W) % W;
      const int off_d = rank_offset(h, d), end_d = rank_offset(h, d + 1);
      const int out_lo = std::max(bounds[me], off_d), out_hi = std::min(bounds[me + 1], end_d);
      const int cnt_out = std::max(out_hi - out_lo, 0);
      const int in_lo = std::max(bounds[sr], h->offset), in_hi = std::min(bounds[sr + 1], h->offset + n);
      const int cnt_in = std::max(in_hi - in_lo, 0);
      if (cnt_out > 0) {
        if (h->dyn_mig_cap < (size_t)cnt_out) {
          cudaFree(h->dyn_mig_map); cudaFree(h->dyn_mig_count); cudaFree(h->dyn_mig_anc);
          h->dyn_mig_map = nullptr; h->dyn_mig_count = nullptr; h->dyn_mig_anc = nullptr; h->dyn_mig_cap = 0;
          const size_t cap = (size_t)cnt_out + (size_t)cnt_out / 4 + 64;
          CK(cudaMalloc(&h->dyn_mig_map, cap * per * sizeof(float)));
          CK(cudaMalloc(&h->dyn_mig_count, cap * sizeof(int)));
          CK(cudaMalloc(&h->dyn_mig_anc, cap * sizeof(int)));
          h->dyn_mig_cap = cap;
        }
        resample_search_kernel<<<cdiv(cnt_out, 256), 256, 0, h->stream>>>(h->cdf_excl, n, base, total, n_new, out_lo, cnt_out, h->offset,
                                                                        udev, sysmode, h->resample_calls, h->dc.seed_lo,
                                                                        h->dc.seed_hi, h->dyn_mig_anc);
        LAUNCH_CHECK(h);
        dyn_gather_kernel<<<cnt_out, 64, 0, h->stream>>>(h->dyn_mig_anc, cnt_out, h->offset, n, h->dmap[db], h->dcount[db],
                                                         h->dyn_mig_map, h->dyn_mig_count, h->Dmax);
        LAUNCH_CHECK(h);
      }
      if (cnt_out > 0 || cnt_in > 0) {
        CKN(ncclGroupStart());
        if (cnt_out > 0) {
          CKN(ncclSend(h->dyn_mig_count, (size_t)cnt_out, ncclInt32, d, comm, h->stream));
          CKN(ncclSend(h->dyn_mig_map, (size_t)cnt_out * per, ncclFloat, d, comm, h->stream));
        }
        if (cnt_in > 0) {
          const size_t first = (size_t)(in_lo - h->offset);
          CKN(ncclRecv(h->dcount[db ^ 1] + first, (size_t)cnt_in, ncclInt32, sr, comm, h->stream));
          CKN(ncclRecv(h->dmap[db ^ 1] + first * per, (size_t)cnt_in * per, ncclFloat, sr, comm, h->stream));
        }
        CKN(ncclGroupEnd());
      }
    }
  }
  fill_kernel<<<cdiv(n_off, 256), 256, 0, h->stream>>>(h->logw, n_off, -phd_logf((float)n_new));
  LAUNCH_CHECK(h);
  CK(cudaMemcpyAsync(h->resample_idx, h->ancestors, (size_t)n_off * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaEventRecord(h->ev[10], h->stream));
  if (ancestors_out) CK(copy_d2h_async(h, ancestors_out, h->ancestors, (size_t)n_off * sizeof(int), h->stream));
  if (h->mbox) CK(copy_d2h_async(h, &h->red_host->err_flag, &h->red->err_flag, sizeof(int), h->stream));
  h->resample_timed = 1;
  if (h->world > 1 || ancestors_out || uniforms) {
    /* the caller's buffers / the peers need the result now; on a single GPU with the counter-based draws nothing does:
     * the next call on the stream is ordered behind the gather, and phdslam_get_timings reads the events later */
    CK(cudaStreamSynchronize(h->stream));
    cudaEventElapsedTime(&h->tim.resample_ms, h->ev[9], h->ev[10]);
    h->resample_timed = 0;
  }
  h->cur ^= 1;
  if (h->Dmax) h->dcur ^= 1;
  h->resample_calls++;
  h->totals_valid = 0;
  if (h->mbox && (h->red_host->err_flag & 4)) {
    phdslam_set_error("peer exchange timed out: another rank did not reach the resampling");
    return PHDSLAM_ERR_NCCL;
  }
  if (h->world == 1) h->n_local = h->n_global = n_new;
  return 0;
}

/* ---- one loop iteration of run_synth (src/main.cpp:1231-1297), in the two halves either side of the point where the
 * reference looks at the state (recoverSlamState's outputs and writeParticlesMat, :1274-1279) ---- */
extern "C" int phdslam_step_filter(phdslam_t* h, int step_index, const float* control, const float* z, int M, int fields,
                                   phdslam_estimate_t* est_out) {
  int rc, nan_rc = 0;
  if (est_out) memset(est_out, 0, sizeof(*est_out));
  if (step_index > 0) {
    for (int i = 0; i < h->cfg.subdivide_predict; ++i) {
      rc = phdslam_predict(h, control, nullptr);
      if (rc) return rc;
    }
  }
  if (M > 0) {
    rc = phdslam_update(h, z, M, fields);
    if (rc == PHDSLAM_ERR_NAN) nan_rc = rc;   /* the reference finishes the iteration, then breaks (main.cpp:1307-1311) */
    else if (rc) return rc;
  }
  phdslam_estimate_t e;
  rc = phdslam_estimate(h, &e);
  if (rc) return rc;
  if (est_out) *est_out = e;
  if (nan_rc || e.neff != e.neff) {
    phdslam_set_error("nan weights detected");
    return PHDSLAM_ERR_NAN;
  }
  return 0;
}

extern "C" int phdslam_step_resample(phdslam_t* h, int M, const phdslam_estimate_t* est, int* resampled_out) {
  int res = 0;
  if ((est->neff <= h->cfg.resample_threshold && M > 0) || h->n_global > 5 * h->cfg.n_particles) {   /* :1286 */
    int rc = phdslam_resample(h, h->cfg.n_particles, nullptr, nullptr);
    if (rc) return rc;
    res = 1;
  } else {                                                                                           /* :1293-1296 */
    ENTER(h);
    iota_kernel<<<cdiv(h->n_local, 256), 256, 0, h->stream>>>(h->resample_idx, h->n_local, h->offset);
    LAUNCH_CHECK(h);
  }
  if (resampled_out) *resampled_out = res;
  return 0;
}

extern "C" int phdslam_step(phdslam_t* h, int step_index, const float* control, const float* z, int M, int fields,
                            phdslam_estimate_t* est_out, int* resampled_out) {
  phdslam_estimate_t e;
  if (resampled_out) *resampled_out = 0;
  int rc = phdslam_step_filter(h, step_index, control, z, M, fields, &e);
  if (est_out) *est_out = e;
  if (rc) return rc;          /* NaN weights included: nothing sensible to resample from */
  return phdslam_step_resample(h, M, &e, resampled_out);
}

/* Sets the number of live particles (single GPU, within the capacity fixed at create time): the caller then imports
 * poses / weights / maps of that many particles.  The reference's host loop changes the count itself (shotgun prediction,
 * resampling back to n_particles: src/phdfilter.cu:1185-1238, src/main.cpp:1286-1289); the drop-in shim needs this to
 * push a host particle set whose size differs from the device's. */
extern "C" int phdslam_set_particle_count(phdslam_t* h, int n) {
  if (h->world > 1 || n < 1 || n > phdslam_particle_capacity(h)) {
    phdslam_set_error("phdslam_set_particle_count: single GPU only, 1 <= n <= particle capacity");
    return PHDSLAM_ERR_INVALID;
  }
  ENTER(h);
  CK(cudaStreamSynchronize(h->stream));
  h->n_local = h->n_global = n;
  h->totals_valid = 0;
  return 0;
}
extern "C" int phdslam_particle_capacity(const phdslam_t* h) {
  if (h->state_ready) return h->n_cap;
  return (h->world > 1) ? h->n_local : particle_capacity(h->cfg, h->n_local);
}

/* ---- import / export ---- */
extern "C" int phdslam_get_poses(phdslam_t* h, phdslam_pose_t* out) {
  ENTER(h);
  const int n = h->n_local;
  std::vector<float> soa((size_t)6 * n);
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_d2h(h, soa.data(), h->pose[h->cur], soa.size() * sizeof(float)));
  for (int i = 0; i < n; ++i) {
    float* o = &out[i].px;
    for (int k = 0; k < 6; ++k) o[k] = soa[(size_t)k * n + i];
  }
  return 0;
}
extern "C" int phdslam_set_poses(phdslam_t* h, const phdslam_pose_t* in) {
  ENTER(h);
  const int n = h->n_local;
  std::vector<float> soa((size_t)6 * n);
  for (int i = 0; i < n; ++i) {
    const float* s = &in[i].px;
    for (int k = 0; k < 6; ++k) soa[(size_t)k * n + i] = s[k];
  }
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_h2d(h, h->pose[h->cur], soa.data(), soa.size() * sizeof(float)));
  return 0;
}
extern "C" int phdslam_get_log_weights(phdslam_t* h, float* out) {
  ENTER(h);
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_d2h(h, out, h->logw, (size_t)h->n_local * sizeof(float)));
  return 0;
}
extern "C" int phdslam_set_log_weights(phdslam_t* h, const float* in) {
  ENTER(h);
  h->totals_valid = 0;
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_h2d(h, h->logw, in, (size_t)h->n_local * sizeof(float)));
  return 0;
}
extern "C" int phdslam_get_map_sizes(phdslam_t* h, int* out) {
  ENTER(h);
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_d2h(h, out, h->count[h->cur], (size_t)h->n_local * sizeof(int)));
  return 0;
}
extern "C" int phdslam_get_maps(phdslam_t* h, phdslam_gaussian2d_t* out, size_t cap) {
  ENTER(h);
  const int n = h->n_local;
  const size_t C = h->Cmax;
  std::vector<int> cnt(n);
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_d2h(h, cnt.data(), h->count[h->cur], (size_t)n * sizeof(int)));
  std::vector<float> blk(PHD_MAP_PLANES * C);
  /* chunked download keeps host memory bounded for large particle counts */
  const size_t chunk = std::max<size_t>(1, (64u << 20) / (PHD_MAP_PLANES * C * 4));
  std::vector<float> buf(chunk * PHD_MAP_PLANES * C);
  size_t k = 0;
  for (size_t p0 = 0; p0 < (size_t)n; p0 += chunk) {
    size_t np = std::min(chunk, (size_t)n - p0);
    CK(copy_d2h(h, buf.data(), h->map[h->cur] + p0 * PHD_MAP_PLANES * C, np * PHD_MAP_PLANES * C * 4));
    for (size_t p = 0; p < np; ++p) {
      const float* b = buf.data() + p * PHD_MAP_PLANES * C;
      for (int i = 0; i < cnt[p0 + p]; ++i) {
        if (k >= cap) { phdslam_set_error("phdslam_get_maps: output too small"); return PHDSLAM_ERR_INVALID; }
        phdslam_gaussian2d_t g;
        g.weight = b[0 * C + i]; g.mean[0] = b[1 * C + i]; g.mean[1] = b[2 * C + i];
        g.cov[0] = b[3 * C + i]; g.cov[1] = b[4 * C + i]; g.cov[2] = b[4 * C + i]; g.cov[3] = b[5 * C + i];
        out[k++] = g;
      }
    }
  }
  return 0;
}
/* ---- mixed feature model: the dynamic maps (SynthSLAM::maps_dynamic), host-side transposition like the static ones ---- */
static int need_mixed(const phdslam* h) {
  if (h->Dmax) return 0;
  phdslam_set_error("dynamic maps exist only with feature_model = 2");
  return PHDSLAM_ERR_INVALID;
}
extern "C" int phdslam_get_map_sizes_dynamic(phdslam_t* h, int* out) {
  ENTER(h);
  if (need_mixed(h)) return PHDSLAM_ERR_INVALID;
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_d2h(h, out, h->dcount[h->dcur], (size_t)h->n_local * sizeof(int)));
  return 0;
}
extern "C" int phdslam_get_maps_dynamic(phdslam_t* h, phdslam_gaussian4d_t* out, size_t cap) {
  ENTER(h);
  if (need_mixed(h)) return PHDSLAM_ERR_INVALID;
  const int n = h->n_local;
  const size_t D = h->Dmax, per = (size_t)DYN_PLANES * D;
  std::vector<int> cnt(n);
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_d2h(h, cnt.data(), h->dcount[h->dcur], (size_t)n * sizeof(int)));
  const size_t chunk = std::max<size_t>(1, (64u << 20) / (per * 4));
  std::vector<float> buf(chunk * per);
  size_t k = 0;
  for (size_t p0 = 0; p0 < (size_t)n; p0 += chunk) {
    const size_t np = std::min(chunk, (size_t)n - p0);
    CK(copy_d2h(h, buf.data(), h->dmap[h->dcur] + p0 * per, np * per * 4));
    for (size_t p = 0; p < np; ++p) {
      const float* b = buf.data() + p * per;
      for (int i = 0; i < cnt[p0 + p]; ++i) {
        if (k >= cap) { phdslam_set_error("phdslam_get_maps_dynamic: output too small"); return PHDSLAM_ERR_INVALID; }
        float* g = reinterpret_cast<float*>(&out[k++]);
        for (int q = 0; q < DYN_PLANES; ++q) g[q] = b[(size_t)q * D + i];
      }
    }
  }
  return 0;
}
extern "C" int phdslam_set_maps_dynamic(phdslam_t* h, const int* sizes, const phdslam_gaussian4d_t* in) {
  ENTER(h);
  if (need_mixed(h)) return PHDSLAM_ERR_INVALID;
  const int n = h->n_local;
  const size_t D = h->Dmax, per = (size_t)DYN_PLANES * D;
  for (int p = 0; p < n; ++p)
    if (sizes[p] < 0 || (size_t)sizes[p] > D) {
      phdslam_set_error("phdslam_set_maps_dynamic: a map exceeds max_components_dynamic");
      return PHDSLAM_ERR_CAPACITY;
    }
  CK(cudaStreamSynchronize(h->stream));
  const size_t chunk = std::max<size_t>(1, (64u << 20) / (per * 4));
  std::vector<float> buf(chunk * per);
  size_t k = 0;
  for (size_t p0 = 0; p0 < (size_t)n; p0 += chunk) {
    const size_t np = std::min(chunk, (size_t)n - p0);
    std::fill(buf.begin(), buf.begin() + np * per, 0.0f);
    for (size_t p = 0; p < np; ++p) {
      float* b = buf.data() + p * per;
      for (int i = 0; i < sizes[p0 + p]; ++i) {
        const float* g = reinterpret_cast<const float*>(&in[k++]);
        for (int q = 0; q < DYN_PLANES; ++q) b[(size_t)q * D + i] = g[q];
      }
    }
    CK(copy_h2d(h, h->dmap[h->dcur] + p0 * per, buf.data(), np * per * 4));
  }
  CK(copy_h2d(h, h->dcount[h->dcur], sizes, (size_t)n * sizeof(int)));
  return 0;
}
/* recoverSlamState: particles.max_map_dynamic = particles.maps_dynamic[max_idx] (src/main.cpp:359); the arg-max particle is
 * the one the last phdslam_estimate found */
extern "C" int phdslam_map_estimate_dynamic(phdslam_t* h, phdslam_gaussian4d_t* out, int cap, int* n_out_p) {
  ENTER(h);
  if (need_mixed(h)) return PHDSLAM_ERR_INVALID;
  phdslam_estimate_t e;
  int rc = phdslam_estimate(h, &e);
  if (rc) return rc;
  const int p = e.map_particle - h->offset;        /* sharded: only the rank that owns the particle returns its map */
  *n_out_p = 0;
  if (p < 0 || p >= h->n_local) return 0;
  const size_t D = h->Dmax, per = (size_t)DYN_PLANES * D;
  int cnt = 0;
  CK(copy_d2h(h, &cnt, h->dcount[h->dcur] + p, sizeof(int)));
  std::vector<float> b(per);
  CK(copy_d2h(h, b.data(), h->dmap[h->dcur] + (size_t)p * per, per * 4));
  for (int i = 0; i < cnt && i < cap; ++i) {
    float* g = reinterpret_cast<float*>(&out[i]);
    for (int q = 0; q < DYN_PLANES; ++q) g[q] = b[(size_t)q * D + i];
  }
  *n_out_p = std::min(cnt, cap);
  return 0;
}

static int set_maps_prefix(phdslam_t* h, int n, const int* sizes, const phdslam_gaussian2d_t* in);
extern "C" int phdslam_set_maps(phdslam_t* h, const int* sizes, const phdslam_gaussian2d_t* in) {
  return set_maps_prefix(h, h->n_local, sizes, in);
}
/* maps of the first n local particles */
static int set_maps_prefix(phdslam_t* h, int n, const int* sizes, const phdslam_gaussian2d_t* in) {
  ENTER(h);
  const size_t C = h->Cmax;
  for (int p = 0; p < n; ++p)
    if (sizes[p] > h->Cmax || sizes[p] < 0) {
      phdslam_set_error("phdslam_set_maps: a map exceeds max_components");
      return PHDSLAM_ERR_CAPACITY;
    }
  CK(cudaStreamSynchronize(h->stream));
  const size_t chunk = std::max<size_t>(1, (64u << 20) / (PHD_MAP_PLANES * C * 4));
  std::vector<float> buf(chunk * PHD_MAP_PLANES * C);
  size_t k = 0;
  for (size_t p0 = 0; p0 < (size_t)n; p0 += chunk) {
    size_t np = std::min(chunk, (size_t)n - p0);
    std::fill(buf.begin(), buf.end(), 0.0f);
    for (size_t p = 0; p < np; ++p) {
      float* b = buf.data() + p * PHD_MAP_PLANES * C;
      for (int i = 0; i < sizes[p0 + p]; ++i) {
        const phdslam_gaussian2d_t& g = in[k++];
        b[0 * C + i] = g.weight; b[1 * C + i] = g.mean[0]; b[2 * C + i] = g.mean[1];
        b[3 * C + i] = g.cov[0]; b[4 * C + i] = g.cov[1]; b[5 * C + i] = g.cov[3];
      }
    }
    CK(copy_h2d(h, h->map[h->cur] + p0 * PHD_MAP_PLANES * C, buf.data(), np * PHD_MAP_PLANES * C * 4));
  }
  CK(copy_h2d(h, h->count[h->cur], sizes, (size_t)n * sizeof(int)));
  return 0;
}
extern "C" int phdslam_get_resample_idx(phdslam_t* h, int* out) {
  ENTER(h);
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_d2h(h, out, h->resample_idx, (size_t)h->n_local * sizeof(int)));
  return 0;
}
extern "C" int phdslam_get_cardinalities(phdslam_t* h, float* out) {
  if (!h->n_card) return PHDSLAM_ERR_INVALID;
  ENTER(h);
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_d2h(h, out, h->card[h->cur], (size_t)h->n_local * h->n_card * sizeof(float)));
  return 0;
}
extern "C" int phdslam_set_cardinalities(phdslam_t* h, const float* in) {
  if (!h->n_card) return PHDSLAM_ERR_INVALID;
  ENTER(h);
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_h2d(h, h->card[h->cur], in, (size_t)h->n_local * h->n_card * sizeof(float)));
  return 0;
}

/* local particles [n_src, n_local) become copies of the first n_src (cyclically): exactly what a resampling with ancestors
 * j mod n_src does -- gather into the back buffers, then flip */
static int tile_from_prefix(phdslam* h, int n_src) {
  const int n = h->n_local, b = h->cur;
  if (n_src >= n) return 0;
  tile_index_kernel<<<cdiv(n, 256), 256, 0, h->stream>>>(h->ancestors, n, n_src, h->logw);
  LAUNCH_CHECK(h);
  resample_gather_kernel<<<cdiv(n, 8), 256, 0, h->stream>>>(h->ancestors, n, 0, n, n, h->pose[b], h->pose[b ^ 1], h->count[b],
                                                          h->count[b ^ 1], h->map[b], h->map[b ^ 1], h->card[b], h->card[b ^ 1],
                                                          h->Cmax, h->n_card, 0, nullptr);
  LAUNCH_CHECK(h);
  CK(cudaStreamSynchronize(h->stream));
  h->cur ^= 1;
  return 0;
}

/* Imports n_src particles and fills the local particles [n_src, n_local) with copies of them (cyclically), on the device:
 * lets a benchmark build a 16 M-particle scene from one it can afford to generate and upload. */
extern "C" int phdslam_import_tiled(phdslam_t* h, int n_src, const phdslam_pose_t* poses, const float* logw, const int* sizes,
                                    const phdslam_gaussian2d_t* maps) {
  ENTER(h);
  const int n = h->n_local;
  if (n_src < 1 || n_src > n) return PHDSLAM_ERR_INVALID;
  if (h->Dmax) {
    phdslam_set_error("phdslam_import_tiled: static feature model only");
    return PHDSLAM_ERR_INVALID;
  }
  CK(cudaStreamSynchronize(h->stream));
  const int b = h->cur;
  std::vector<float> plane(n_src);
  for (int k = 0; k < 6; ++k) {
    for (int i = 0; i < n_src; ++i) plane[i] = (&poses[i].px)[k];
    CK(copy_h2d(h, h->pose[b] + (size_t)k * n, plane.data(), (size_t)n_src * sizeof(float)));
  }
  CK(copy_h2d(h, h->logw, logw, (size_t)n_src * sizeof(float)));
  int rc = set_maps_prefix(h, n_src, sizes, maps);
  if (rc) return rc;
  h->totals_valid = 0;
  h->tile_n = (n_src < n) ? n_src : 0;
  return tile_from_prefix(h, n_src);
}

extern "C" int phdslam_particle_checksums(phdslam_t* h, unsigned long long* out) {
  ENTER(h);
  const int n = h->n_local;
  /* q_fx is per-resampling scratch of n_cap 64-bit words */
  particle_checksum_kernel<<<cdiv(n, 8), 256, 0, h->stream>>>(h->pose[h->cur], h->count[h->cur], h->map[h->cur], h->card[h->cur], n,
                                                             h->Cmax, h->n_card, h->q_fx);
  LAUNCH_CHECK(h);
  if (h->Dmax) {
    dyn_checksum_add_kernel<<<cdiv(n, 8), 256, 0, h->stream>>>(h->dcount[h->dcur], h->dmap[h->dcur], n, h->Dmax, h->q_fx);
    LAUNCH_CHECK(h);
  }
  CK(copy_d2h_async(h, out, h->q_fx, (size_t)n * sizeof(unsigned long long), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return 0;
}

extern "C" int phdslam_map_estimate(phdslam_t* h, int which, phdslam_gaussian2d_t* out, int cap, int* n_out_p) {
  int* n_out = n_out_p;
  ENTER(h);
  if (which == 1) {
    phdslam_estimate_t e;
    int rc = phdslam_estimate(h, &e);
    if (rc) return rc;
    int lp = e.map_particle - h->offset;
    if (lp < 0 || lp >= h->n_local) { *n_out = 0; return 0; }
    const size_t C = h->Cmax;
    int cnt = 0;
    CK(copy_d2h(h, &cnt, h->count[h->cur] + lp, sizeof(int)));
    std::vector<float> b(PHD_MAP_PLANES * C);
    CK(copy_d2h(h, b.data(), h->map[h->cur] + (size_t)lp * PHD_MAP_PLANES * C, b.size() * 4));
    *n_out = cnt;
    for (int i = 0; i < cnt && i < cap; ++i) {
      phdslam_gaussian2d_t g;
      g.weight = b[0 * C + i]; g.mean[0] = b[1 * C + i]; g.mean[1] = b[2 * C + i];
      g.cov[0] = b[3 * C + i]; g.cov[1] = b[4 * C + i]; g.cov[2] = b[4 * C + i]; g.cov[3] = b[5 * C + i];
      out[i] = g;
    }
    return 0;
  }
  if (which != 2) return PHDSLAM_ERR_INVALID;
  /* ---- EAP map: computeExpectedMap + reduceGaussianMixture (main.cpp:290-316, gm_reduce.cpp:57-134) ---- */
  const int n = h->n_local;
  ncclComm_t comm = (ncclComm_t)h->nccl_comm;
  /* concat offsets = exclusive scan of the map sizes */
  std::vector<int> cnt(n);
  CK(cudaStreamSynchronize(h->stream));
  CK(copy_d2h(h, cnt.data(), h->count[h->cur], (size_t)n * sizeof(int)));
  std::vector<unsigned long long> off(n + 1, 0);
  for (int p = 0; p < n; ++p) off[p + 1] = off[p] + (unsigned long long)cnt[p];
  const unsigned long long n_tot = off[n];
  unsigned long long gbase = 0;
  if (h->world > 1) {
    if (!h->totals_dev) CK(cudaMalloc(&h->totals_dev, (size_t)h->world * sizeof(unsigned long long)));
    unsigned long long* d_one = nullptr;
    CK(cudaMalloc(&d_one, sizeof(unsigned long long)));
    CK(copy_h2d(h, d_one, &n_tot, sizeof(n_tot)));
    CKN(ncclAllGather(d_one, h->totals_dev, 1, ncclUint64, comm, h->stream));
    std::vector<unsigned long long> all(h->world);
    CK(copy_d2h_async(h, all.data(), h->totals_dev, (size_t)h->world * 8, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_one);
    for (int r = 0; r < h->rank; ++r) gbase += all[r];
  }
  float4* rec = nullptr;
  unsigned long long* d_off = nullptr;
  EapAcc* acc = nullptr;
  phdslam_gaussian2d_t* d_out = nullptr;
  unsigned long long* key_host = nullptr;
  CK(cudaMalloc(&rec, std::max<unsigned long long>(n_tot, 1) * 2 * sizeof(float4)));
  CK(cudaMalloc(&d_off, (size_t)(n + 1) * sizeof(unsigned long long)));
  CK(cudaMalloc(&acc, sizeof(EapAcc)));
  CK(cudaMalloc(&d_out, (size_t)std::max(cap, 1) * sizeof(phdslam_gaussian2d_t)));
  CK(cudaMallocHost(&key_host, sizeof(unsigned long long)));
  CK(copy_h2d_async(h, d_off, off.data(), (size_t)(n + 1) * sizeof(unsigned long long), h->stream));
  CK(cudaMemsetAsync(acc, 0, sizeof(EapAcc), h->stream));
  eap_concat_kernel<<<cdiv(n, 8), 256, 0, h->stream>>>(h->map[h->cur], h->count[h->cur], h->logw, d_off, n, h->Cmax, rec);
  LAUNCH_CHECK(h);
  const int blocks = (int)std::min<unsigned long long>(std::max<unsigned long long>((n_tot + 255) / 256, 1), 148 * 8);
  int rc = 0;
  for (long long round = 0;; ++round) {
    eap_argmax_kernel<<<blocks, 256, 0, h->stream>>>(rec, n_tot, gbase, acc);
    LAUNCH_CHECK(h);
    if (h->world > 1) CKN(ncclAllReduce(&acc->key, &acc->key, 1, ncclUint64, ncclMax, comm, h->stream));
    eap_seed_kernel<<<1, 32, 0, h->stream>>>(rec, n_tot, gbase, acc);
    LAUNCH_CHECK(h);
    if (h->world > 1) CKN(ncclAllReduce(acc->seed, acc->seed, 8, ncclFloat, ncclSum, comm, h->stream));
    eap_cluster_kernel<<<blocks, 256, 0, h->stream>>>(rec, n_tot, gbase, h->cfg.min_separation, acc);
    LAUNCH_CHECK(h);
    if (h->world > 1) CKN(ncclAllReduce(&acc->w, &acc->w, 7, ncclDouble, ncclSum, comm, h->stream));
    eap_finalize_kernel<<<1, 32, 0, h->stream>>>(acc, d_out, cap, key_host);
    LAUNCH_CHECK(h);
    CK(cudaStreamSynchronize(h->stream));
    if (*key_host == 0) break;
  }
  int n_eap = 0;
  CK(copy_d2h(h, &n_eap, &acc->n_out, sizeof(int)));
  *n_out_p = n_eap;
  if (n_eap > 0) CK(copy_d2h(h, out, d_out, (size_t)std::min(n_eap, cap) * sizeof(phdslam_gaussian2d_t)));
  cudaFree(rec); cudaFree(d_off); cudaFree(acc); cudaFree(d_out); cudaFreeHost(key_host);
  return rc;
}

/* ---- timings / snapshot ---- */
extern "C" int phdslam_get_timings(phdslam_t* h, phdslam_timings_t* out) {
  ENTER(h);
  CK(cudaStreamSynchronize(h->stream));
  float ms = 0;
  if (h->predict_calls && cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->tim.predict_ms = ms;
  if (h->resample_timed && cudaEventElapsedTime(&ms, h->ev[9], h->ev[10]) == cudaSuccess) h->tim.resample_ms = ms;
  h->resample_timed = 0;
  h->tim.launches = h->launches;
  *out = h->tim;
  return 0;
}

extern "C" int phdslam_snapshot(phdslam_t* h) {
  ENTER(h);
  /* a tiled particle set (phdslam_import_tiled) is snapshot as its tile_n distinct particles only */
  const size_t n = h->tile_n ? (size_t)h->tile_n : (size_t)h->n_local, C = h->Cmax;
  const size_t np = h->n_local;                 /* stride of the pose planes */
  CK(cudaStreamSynchronize(h->stream));
  if (!h->snap_pose) {
    const size_t nc = h->tile_n ? (size_t)h->tile_n : (size_t)h->n_cap;
    CK(cudaMalloc(&h->snap_pose, 6 * nc * sizeof(float)));
    CK(cudaMalloc(&h->snap_count, nc * sizeof(int)));
    CK(cudaMalloc(&h->snap_map, nc * PHD_MAP_PLANES * C * sizeof(float)));
    CK(cudaMalloc(&h->snap_logw, nc * sizeof(float)));
    if (h->n_card) CK(cudaMalloc(&h->snap_card, nc * h->n_card * sizeof(float)));
  }
  h->snap_n = (int)np;
  for (int k = 0; k < 6; ++k)
    CK(cudaMemcpy(h->snap_pose + (size_t)k * n, h->pose[h->cur] + (size_t)k * np, n * sizeof(float), cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(h->snap_count, h->count[h->cur], n * sizeof(int), cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(h->snap_map, h->map[h->cur], n * PHD_MAP_PLANES * C * sizeof(float), cudaMemcpyDeviceToDevice));
  CK(cudaMemcpy(h->snap_logw, h->logw, n * sizeof(float), cudaMemcpyDeviceToDevice));
  if (h->n_card) CK(cudaMemcpy(h->snap_card, h->card[h->cur], n * h->n_card * sizeof(float), cudaMemcpyDeviceToDevice));
  if (h->Dmax) {
    if (!h->snap_dmap) {
      CK(cudaMalloc(&h->snap_dmap, (size_t)h->n_cap * DYN_PLANES * h->Dmax * sizeof(float)));
      CK(cudaMalloc(&h->snap_dcount, (size_t)h->n_cap * sizeof(int)));
    }
    CK(cudaMemcpy(h->snap_dmap, h->dmap[h->dcur], n * DYN_PLANES * (size_t)h->Dmax * sizeof(float), cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(h->snap_dcount, h->dcount[h->dcur], n * sizeof(int), cudaMemcpyDeviceToDevice));
  }
  h->snap_predict_calls = h->predict_calls;
  h->snap_resample_calls = h->resample_calls;
  return 0;
}
extern "C" int phdslam_restore(phdslam_t* h) {
  ENTER(h);
  if (!h->snap_pose) return PHDSLAM_ERR_INVALID;
  h->totals_valid = 0;
  if (h->world == 1) h->n_local = h->n_global = h->snap_n;
  const size_t n = h->tile_n ? (size_t)h->tile_n : (size_t)h->n_local, C = h->Cmax;
  const size_t np = h->n_local;
  for (int k = 0; k < 6; ++k)
    CK(cudaMemcpyAsync(h->pose[h->cur] + (size_t)k * np, h->snap_pose + (size_t)k * n, n * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->count[h->cur], h->snap_count, n * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->map[h->cur], h->snap_map, n * PHD_MAP_PLANES * C * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  CK(cudaMemcpyAsync(h->logw, h->snap_logw, n * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  if (h->n_card)
    CK(cudaMemcpyAsync(h->card[h->cur], h->snap_card, n * h->n_card * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  if (h->Dmax && h->snap_dmap) {
    CK(cudaMemcpyAsync(h->dmap[h->dcur], h->snap_dmap, n * DYN_PLANES * (size_t)h->Dmax * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->dcount[h->dcur], h->snap_dcount, n * sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
  }
  h->predict_calls = h->snap_predict_calls;
  h->resample_calls = h->snap_resample_calls;
  if (h->tile_n) return tile_from_prefix(h, h->tile_n);     /* two kernels, outside any timed region of the caller */
  return 0;
}
