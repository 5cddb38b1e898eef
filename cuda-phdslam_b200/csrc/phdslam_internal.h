/* phdslam_internal.h -- handle layout and kernel-facing derived configuration (not installed). */
#ifndef PHDSLAM_INTERNAL_H
#define PHDSLAM_INTERNAL_H

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/phdslam.h"

void phdslam_set_error(const std::string& s);

#define PHD_MAX_MEAS 256   /* reference: __constant__ RangeBearingMeasurement Z[256] (src/phdfilter.cu:120) */
#define PHD_NPLANES 7      /* dense update term = {c0,c1,c2,c3,mx,my,w}, the field order of Gaussian2D */
#define PHD_MAX_SUB 64     /* sub-batches per update whose merge overlaps the next sub-batch's update */
#define PHD_MAP_PLANES 6   /* persistent map component = {w,mx,my,pxx,pxy,pyy} (covariance stored once) */

/* Constants derived on the host once per set_config and passed BY VALUE to every kernel
 * (the reference copies its whole 324-byte SlamConfig to __constant__ memory, src/phdfilter.cu:3885). */
struct DevCfg {
  float min_range, max_range, max_bearing;
  float lo2, hi2, hb2;            /* "nearly in range" bounds, exact fp32 images of 0.8*minR, 1.2*maxR, 1.2*maxB */
  float var_r, var_b;             /* stdRange^2, stdBearing^2 */
  float bvar_r, bvar_b;           /* (std*birthNoiseFactor)^2 */
  float pd, log_pd;
  float clutter_density, clutter_rate, birth_weight, log_birth_weight;
  float min_sep, min_w;
  int distance_metric, particle_weighting, labeled, filter_type;
  int motion_type;
  float dt_sub, l, h, a, b, std_alpha, std_enc, ax3, ay3, ayaw3;
  uint32_t seed_lo, seed_hi;
  int Cmax, n_card;
  float log_clutter_rate;
};

struct Reductions {       /* device-resident cross-particle accumulators (all order-independent) */
  unsigned max_key;       /* ordered-uint image of max log-weight (0 = none) */
  int pad0;
  /* the next three are adjacent 64-bit integers: ONE exact integer all-reduce when the particles are sharded */
  unsigned long long sum_fx;      /* sum exp(w-max) in Q36 */
  unsigned long long nan_count;   /* particles whose log-weight is NaN (reference: `if isnan(nEff) break`, main.cpp:1307) */
  unsigned long long err_ranks;   /* ranks whose update / merge raised err_flag: every rank returns the same status */
  unsigned long long neff_fx;     /* sum exp(2w) in Q60 */
  long long pose_fx[6];           /* sum exp(w)*pose in Q40 */
  unsigned long long argmax_key;  /* (ordered weight << 32) | ~global_index */
  unsigned long long cdf_total;   /* local sum of Q40 weights */
  int err_flag;                   /* capacity / overflow flags raised by kernels */
  int max_terms;                  /* max over particles of padded dense term count */
  unsigned long long total_terms; /* sum over particles of padded dense term count */
  int max_cand;                   /* max over particles of the merge candidate count (sizes the next step's merge) */
  int ovf_n;                      /* particles queued for the general merge kernel */
};

struct phdslam {
  phdslam_config_t cfg;
  DevCfg dc;
  int device;
  cudaStream_t stream;
  int rank, world;
  int n_global, n_local, offset;
  int state_ready;       /* particle state allocated and initialised (deferred from phdslam_create to the first use) */
  int n_cap;             /* particle capacity of every per-particle array (> n_particles only for n_predict_particles > 1) */
  int Cmax, n_card, Smax;
  int Scap_max, Scap_pinned;
  int Scap;              /* shared-memory candidate capacity of merge_fast_kernel, adapted from the last step */
  int* ovf_list;         /* [n_local] */
  /* persistent state, double buffered (front = cur) */
  int cur;
  float* pose[2];        /* [6][n_local] SoA */
  int* count[2];         /* [n_local] */
  float* map[2];         /* [n_local][6][Cmax] */
  float* card[2];        /* [n_local][n_card] (CPHD) */
  float* logw;           /* [n_local] */
  int* resample_idx;     /* [n_local] */
  /* snapshot for bench */
  float* snap_pose; int* snap_count; float* snap_map; float* snap_card; float* snap_logw;
  unsigned snap_predict_calls, snap_resample_calls;
  int snap_n;
  int tile_n;            /* > 0: the local particles are tile_n distinct ones repeated (phdslam_import_tiled); snapshot / restore
                            then keep only those and re-tile on the device */
  /* per-step scratch */
  uint8_t* cls;                      /* [n_local][Cmax] in-range class */
  int* n_in;                         /* [n_local] */
  float* dlogw;                      /* [n_local] */
  unsigned long long* tpad;          /* [n_local] padded term counts */
  unsigned long long* toff;          /* [n_local+1] exclusive scan */
  unsigned long long* scan_tmp;      /* block sums */
  float* dense; size_t dense_floats; /* dense update-term buffer */
  float4* cand; float4* cand_in; size_t cand_cap; /* merge candidate records (2 x float4 each): ordered scratch / update output */
  int* n_cand;                       /* [n_local] candidates emitted by the update kernel */
  float* z_dev;                      /* [3][256]: range, bearing, label */
  double* draws_dev; size_t draws_cap;
  unsigned long long* q_fx;          /* [n_local] Q40 weights */
  unsigned long long* cdf_excl;      /* [n_local+1] */
  int* ancestors;                    /* [n_local] */
  Reductions* red;                   /* device */
  Reductions* red_host;              /* pinned */
  float* z_host;                     /* pinned staging of the measurements, [3][256] */
  int resample_timed;                /* ev[9..10] of the last resampling have not been read yet */
  int nan_seen;                      /* the last update saw NaN particle weights (reported by phdslam_step after the estimate) */
  /* counters */
  unsigned predict_calls, resample_calls;
  unsigned long long launches;
  cudaEvent_t ev[12];
  /* update / merge overlap: merge kernels run on a second, higher-priority stream behind per-sub-batch events */
  cudaStream_t stream_m;
  cudaEvent_t ev_sub[PHD_MAX_SUB], ev_merge_done;
  int overlap;
  phdslam_timings_t tim;
  void* nccl_comm;
  /* resampling migration staging (sender side), grown on demand */
  float* mig_map; float* mig_pose; int* mig_count; int* mig_anc; float* mig_card; float* mig_pose_in; size_t mig_cap;
  unsigned long long* totals_dev; /* [world] all-gathered local CDF totals */
  /* NVLink peer window (world > 1): pose/count/map/card of both buffers and the pushed-ancestor array live in ONE
   * allocation that every peer maps through CUDA IPC, so that the resampling exchange is a gather kernel that writes
   * the offspring straight into the owner's back buffer (no packing, no send/recv).  p2p = 0: NCCL send/recv path. */
  unsigned char* peer_slab; size_t peer_slab_bytes;
  int* anc_in;                    /* [n_local] ancestors pushed by the peers */
  int p2p;                        /* 1: peers' slabs are mapped */
  int slab_sw;                    /* pose[i] and card[i] live in window slot i ^ slab_sw (the update swaps the pointers) */
  unsigned char** peer_base;      /* [world] mapped slab of every rank (own slab at [rank]) */
  int* barrier_dev;               /* 2 ints: end-of-exchange all-reduce */
  /* mailbox exchange through the peer window (kernels.cuh: mbox_exchange_kernel): replaces the tiny NCCL collectives of the
   * weight statistics, the estimate and the resampling barrier when the window is mapped */
  int mbox;                       /* 1: use it (p2p && PHDSLAM_MBOX != 0) */
  unsigned long long* mbox_base;  /* this rank's mailbox region inside the slab */
  unsigned long long mbox_seq;    /* sequence number of the next exchange (same on every rank) */
  unsigned long long* gath_dev;   /* [world][16] records of the last gathering exchange */
  unsigned long long* gath_host;  /* pinned copy */
  int totals_valid;               /* totals_host = every rank's resampling CDF total for the CURRENT weights (set by estimate) */
  unsigned long long totals_host[8];
  int* mig_anc2; size_t mig_anc_cap; /* ancestors of the offspring interval being pushed */
  float* lfact;                   /* log-factorial table for the CPHD terms (PHD_LF_MAX floats) */
  /* mixed feature model (feature_model = 2; csrc/mixed.cuh): the dynamic maps, double buffered like the static ones */
  int Dmax, Sd;                   /* per-particle capacity of the dynamic map / of its prune survivors */
  int dcur;
  float* dmap[2];                 /* [n_cap][21][Dmax] plane-SoA Gaussian4D components */
  int* dcount[2];                 /* [n_cap] */
  float* mix_dsum; float* mix_nhat; float* mix_L;   /* [n][256], [n], [n][256]: coupling of the two maps' updates */
  phdslam_gaussian4d_t* dcand;    /* [n][Sd] */
  float* snap_dmap; int* snap_dcount;
  cudaEvent_t ev_dyn[4];          /* around dyn_pre_kernel and dyn_update_kernel */
  float* dyn_mig_map; int* dyn_mig_count; int* dyn_mig_anc; size_t dyn_mig_cap;   /* sharded runs: staging of the dynamic maps sent to another rank */
};

#endif
