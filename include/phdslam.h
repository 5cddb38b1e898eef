/*
 * phdslam.h -- C-ABI of libphdslam.so, the B200 (sm_100a) drop-in for the
 * particle-parallel RB-PHD-SLAM filter step of cheesinglee/cuda-PHDSLAM.
 *
 * The reference's boundary between its host loop (src/main.cpp:1178-1312) and
 * its CUDA code is the set of C++ prototypes in src/phdfilter.h:10-34 (the
 * older src/phdfilter.cu.bak:54-82 declared the same set `extern "C"`), which
 * pass `SynthSLAM&` / `std::vector` objects and re-upload the whole map every
 * step.  This header replaces each of them with a plain-C entry point on an
 * opaque handle that owns *persistent device state*; every function below
 * names the reference interface it stands in for.
 *
 * All functions return 0 on success or a negative phdslam_status; none calls
 * exit() (the reference's checkCudaErrors does).  A handle is not thread-safe:
 * one host thread (one rank per GPU) drives it, as run_synth does.
 */
#ifndef PHDSLAM_H
#define PHDSLAM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum phdslam_status {
  PHDSLAM_OK = 0,
  PHDSLAM_ERR_CUDA = -1,        /* a CUDA runtime call failed; see phdslam_last_error */
  PHDSLAM_ERR_INVALID = -2,     /* bad argument / unsupported option */
  PHDSLAM_ERR_CAPACITY = -3,    /* a particle's map exceeded max_components (reference maps are unbounded) */
  PHDSLAM_ERR_NAN = -4,         /* NaN particle weights (reference: `if isnan(nEff) break`, main.cpp:1307) */
  PHDSLAM_ERR_NCCL = -5,
  PHDSLAM_ERR_IO = -6
} phdslam_status;

/* reference: ConstantVelocityState, src/slamtypes.h:44-51 (24 B) */
typedef struct phdslam_pose {
  float px, py, ptheta, vx, vy, vtheta;
} phdslam_pose_t;

/* reference: Gaussian2D, src/slamtypes.h:123-127 (28 B; cov column-major) */
typedef struct phdslam_gaussian2d {
  float cov[4];
  float mean[2];
  float weight;
} phdslam_gaussian2d_t;

/* reference: Gaussian4D, src/slamtypes.h:135-139 (84 B; state x, y, vx, vy; cov column-major) -- a component of the
 * dynamic map of the mixed feature model (feature_model = 2) */
typedef struct phdslam_gaussian4d {
  float cov[16];
  float mean[4];
  float weight;
} phdslam_gaussian4d_t;

/*
 * POD mirror of the live fields of the reference's SlamConfig
 * (src/slamtypes.h:142-250) with the defaults of loadConfig
 * (src/main.cpp:961-1048).  Keys of cfg/config.cfg map 1:1 (see
 * phdslam_config_load).  Fields after `--- extensions ---` do not exist in the
 * reference.
 */
typedef struct phdslam_config {
  float x0, y0, yaw0, vx0, vy0, vyaw0;           /* initial_x .. initial_vyaw */
  int motion_type;                               /* 0 = constant velocity, 1 = Ackerman */
  float ax, ay, ayaw;                            /* acc_x, acc_y, acc_yaw */
  float dt;
  float min_range, max_range, max_bearing;
  float std_range, std_bearing;
  float clutter_rate;
  float clutter_density;                         /* derived: clutter_rate/(2*max_bearing*max_range), main.cpp:1065 */
  float pd;
  int n_particles;
  int n_predict_particles;
  int subdivide_predict;
  float resample_threshold;
  float birth_weight;
  float birth_noise_factor;
  float min_separation;
  float min_feature_weight;
  int particle_weighting;                        /* 0 cluster-process, 1 Vo empty-map; 2 (single-feature, unfinished in the
                                                    reference: src/phdfilter.cu:3600-3661) -> PHDSLAM_ERR_INVALID */
  int distance_metric;                           /* 0 Mahalanobis, 1 Hellinger */
  int max_cardinality;
  int filter_type;                               /* 0 PHD, 1 CPHD */
  int map_estimate;                              /* bitmask: 1 = MAP map, 2 = EAP map (main.cpp:344,363) */
  int feature_model;                             /* 0 static; 2 mixed = static + constant-velocity features (the `--- mixed
                                                    feature model ---` fields below); 1 (dynamic only: the reference's update
                                                    launch for it is commented out, src/phdfilter.cu:3663-3670) -> INVALID */
  float l, h, a, b, std_encoder, std_alpha;      /* Ackerman vehicle */
  int labeled_measurements;
  int follow_trajectory;
  int max_steps;
  int n_steps;
  char data_directory[1024];
  /* --- extensions --- */
  int measurement_fields;                        /* 2 = "r b" pairs (README, bundled data), 3 = "r b label" (HEAD parser) */
  int max_components;                            /* per-particle map capacity of the device SoA */
  int resample_mode;                             /* 0 = stratified as HEAD main.cpp:461-499, 1 = systematic as .bak:3279-3326 */
  int log_layout;                                /* 0 = README 5-line, 1 = writeLog 7-line (main.cpp:848-954) */
  unsigned long long seed;                       /* Philox key; the reference seeds mt19937 with time(0) */
  int update_mode;                               /* 0 = dense (reference-equivalent materialised update terms), 1 = fused */
  unsigned long long update_buffer_bytes;        /* cap on the dense update-term buffer; particles stream through it */
  /* --- mixed feature model (feature_model = 2): reference keys ps, tau, beta, std_ax_features, std_ay_features,
   * cov_vx_birth, cov_vy_birth (src/main.cpp:990,1022-1025,1037-1038); appended here so that the offsets above stay --- */
  float ps;                                      /* survival probability of a dynamic feature */
  float tau, beta;                               /* jump-Markov sigmoid: p = 1/(1 + exp(beta (tau - |v|))) */
  float std_ax_features, std_ay_features;        /* white-acceleration noise of the constant-velocity feature model */
  float cov_vx_birth, cov_vy_birth;              /* velocity variances of a dynamic birth term */
  int max_components_dynamic;                    /* extension: per-particle capacity of the dynamic map */
} phdslam_config_t;

typedef struct phdslam phdslam_t;

/* ---- configuration (reference: loadConfig, src/main.cpp:956-1073; host only, no GPU needed) ---- */
void phdslam_config_defaults(phdslam_config_t* cfg);
/* Parses a boost::program_options style key=value file with the reference's key names.
 * Unknown keys are reported on stderr and skipped (the reference prints the exception and keeps defaults). */
int phdslam_config_load(const char* path, phdslam_config_t* cfg);
/* Sets one key from strings, as the INI parser does.  Returns 0, or PHDSLAM_ERR_INVALID for an unknown key. */
int phdslam_config_set(phdslam_config_t* cfg, const char* key, const char* value);

const char* phdslam_last_error(void);
const char* phdslam_version(void);

/* ---- lifetime ---- */
/* Creates the handle for cfg->n_particles particles on CUDA device `device`; the persistent device state is allocated on
 * the first call that touches it (or by phdslam_dist_init, for this rank's share only: a particle set that does not fit
 * one GPU can still be created and then sharded) and initialised as run_synth does (src/main.cpp:1129-1144): every pose = (x0..vyaw0),
 * log-weight = -log(N), empty maps, uniform CPHD cardinality.  Also stands in for
 * initRandomNumberGenerators() + setDeviceConfig() (src/phdfilter.cu:142,3885). */
int phdslam_create(const phdslam_config_t* cfg, int device, phdslam_t** out);
void phdslam_destroy(phdslam_t* h);
/* reference: setDeviceConfig(config) (src/phdfilter.cu:3885-3890); n_particles/max_components may not change. */
int phdslam_set_config(phdslam_t* h, const phdslam_config_t* cfg);
int phdslam_get_config(const phdslam_t* h, phdslam_config_t* cfg);

/* Particle sharding over ranks (one process per GPU).  Rank r of `world` owns the contiguous block of
 * global particle indices [r*N/world, (r+1)*N/world).  Must be called before any filter call when world > 1;
 * nccl_unique_id is the 128-byte ncclUniqueId created by rank 0 (phdslam_dist_unique_id) and broadcast by the caller. */
int phdslam_dist_unique_id(void* id128);
int phdslam_dist_init(phdslam_t* h, int rank, int world, const void* nccl_unique_id128);
/* Host-only planning of the global resampling exchange (no GPU, no NCCL; also used by the CPU tests).
 * totals[r] = rank r's sum of Q40 fixed-point weights (phd_detmath.h PHD_FX_CDF_BITS).  On return
 * bounds[r] (r = 0..world) is the first offspring index j whose ancestor lives on a rank >= r, so offspring
 * j has its ancestor on rank s iff bounds[s] <= j < bounds[s+1] (thresholds are monotone in j).
 * uniforms: NULL -> counter-based RNG keyed by (seed, call), else n_new+1 injected draws. */
int phdslam_plan_migration(int world, const unsigned long long* totals, int n_new, const double* uniforms,
                           int resample_mode, unsigned call, unsigned long long seed, int* bounds);
/* Offspring threshold R_j = floor(((j + u_j)/n_new) * total) of the canonical resampler (host evaluation). */
unsigned long long phdslam_resample_threshold(int j, int n_new, unsigned long long total, const double* uniforms,
                                              int resample_mode, unsigned call, unsigned long long seed);

/* ---- the filter step ---- */
/* reference: phdPredict(SynthSLAM&, ...) (src/phdfilter.cu:1080-1257) for ONE sub-step.
 * control = {v_encoder, alpha} (file order, main.cpp:183); ignored for motion_type 0, may be NULL.
 * draws: NULL -> noise from the counter-based RNG; else injected standard-normal draws in the
 * reference's call order (Ackerman: per particle n_alpha then n_encoder, :1148-1152; CV: ax, ay, atheta,
 * :1113-1117), 2 or 3 per LOCAL particle. */
int phdslam_predict(phdslam_t* h, const float* control, const double* draws);

/* reference: phdUpdateSynth(SynthSLAM&, measurementSet) (src/phdfilter.cu:3336-3761):
 * in-range split, births, GM-PHD (or CPHD) update, prune, merge, particle weight update + normalisation.
 * z = M records of `fields` floats: {range, bearing} or {range, bearing, label}.  M > 256 is truncated
 * to 256 as the reference does (:3390-3394).  M == 0 is a no-op (main.cpp:1258). */
int phdslam_update(phdslam_t* h, const float* z, int M, int fields);

typedef struct phdslam_estimate {
  phdslam_pose_t expected_pose;  /* sum_i exp(w_i) * state_i (main.cpp:324-340) */
  int map_particle;              /* global index of the max-weight particle (main.cpp:347-356), -1 if unused */
  float neff;                    /* 1/sum exp(2w)/N (main.cpp:1281-1284) */
  float max_log_weight;
} phdslam_estimate_t;

/* reference: recoverSlamState (src/main.cpp:318-388) + the nEff computation (main.cpp:1281-1284). */
int phdslam_estimate(phdslam_t* h, phdslam_estimate_t* out);
/* Map estimate selected by cfg.map_estimate: bit 1 -> MAP map (copy of the max-weight particle's map),
 * bit 2 -> EAP map (computeExpectedMap + reduceGaussianMixture, main.cpp:290-316, gm_reduce.cpp:57-134).
 * which = 1 or 2.  Writes up to cap gaussians, returns the count in *n. */
int phdslam_map_estimate(phdslam_t* h, int which, phdslam_gaussian2d_t* out, int cap, int* n);

/* reference: resampleParticles(particles, n_particles) (src/main.cpp:453-501) + copy_particles
 * (src/slamtypes.h:313-333).  uniforms: NULL -> counter-based RNG, else n_new+1 injected draws in [0,1)
 * in the reference's call order (one discarded, then one per offspring).  ancestors_out (may be NULL)
 * receives the n_new global ancestor indices of the LOCAL offspring block.  n_new must equal n_particles. */
int phdslam_resample(phdslam_t* h, int n_new, const double* uniforms, int* ancestors_out);

/* One iteration of the run_synth loop body (src/main.cpp:1231-1297): predict (subdivide_predict sub-steps,
 * skipped when step_index == 0), update when M > 0, estimate, nEff test, resample when triggered.
 * resampled_out: 1 if resampling happened. */
int phdslam_step(phdslam_t* h, int step_index, const float* control, const float* z, int M, int fields,
                 phdslam_estimate_t* est_out, int* resampled_out);
/* The same iteration in its two halves, for a caller that looks at the particle set where run_synth does -- after
 * recoverSlamState and before resampleParticles (src/main.cpp:1274-1279: the map estimate, the particle weights and
 * poses of the log are those of the weighted, not yet resampled particles; resample_idx still holds the previous
 * step's ancestors):
 *   phdslam_step_filter   = predict + update + estimate (main.cpp:1244-1274, 1281-1284).  Returns PHDSLAM_ERR_NAN
 *                           (with *est_out filled) when particle weights are NaN (main.cpp:1307-1311);
 *   phdslam_step_resample = the nEff test and resampleParticles, or resample_idx = identity (main.cpp:1286-1297). */
int phdslam_step_filter(phdslam_t* h, int step_index, const float* control, const float* z, int M, int fields,
                        phdslam_estimate_t* est_out);
int phdslam_step_resample(phdslam_t* h, int M, const phdslam_estimate_t* est, int* resampled_out);

/* ---- state import / export (tests, checkpoints, the log writer) ---- */
int phdslam_n_local(const phdslam_t* h);             /* particles owned by this rank */
/* Number of live particles (single GPU; 1 <= n <= phdslam_particle_capacity): for a host loop that changes the
 * particle count itself (src/phdfilter.cu:1185-1238, src/main.cpp:1286-1289) and then imports the new set. */
int phdslam_set_particle_count(phdslam_t* h, int n);
int phdslam_particle_capacity(const phdslam_t* h);
int phdslam_local_offset(const phdslam_t* h);        /* global index of local particle 0 */
int phdslam_get_poses(phdslam_t* h, phdslam_pose_t* out /* n_local */);
int phdslam_set_poses(phdslam_t* h, const phdslam_pose_t* in);
int phdslam_get_log_weights(phdslam_t* h, float* out);
int phdslam_set_log_weights(phdslam_t* h, const float* in);
int phdslam_get_map_sizes(phdslam_t* h, int* out);
/* maps concatenated particle after particle, sizes[] entries each */
int phdslam_get_maps(phdslam_t* h, phdslam_gaussian2d_t* out, size_t cap);
int phdslam_set_maps(phdslam_t* h, const int* sizes, const phdslam_gaussian2d_t* in);
int phdslam_get_resample_idx(phdslam_t* h, int* out);
/* CPHD cardinality distributions, (max_cardinality+1) log-probabilities per particle */
int phdslam_get_cardinalities(phdslam_t* h, float* out);
int phdslam_set_cardinalities(phdslam_t* h, const float* in);

/* Mixed feature model: the dynamic maps (reference: SynthSLAM::maps_dynamic, src/slamtypes.h:291), concatenated particle
 * after particle like the static ones.  PHDSLAM_ERR_INVALID unless feature_model = 2. */
int phdslam_get_map_sizes_dynamic(phdslam_t* h, int* out);
int phdslam_get_maps_dynamic(phdslam_t* h, phdslam_gaussian4d_t* out, size_t cap);
int phdslam_set_maps_dynamic(phdslam_t* h, const int* sizes, const phdslam_gaussian4d_t* in);
/* MAP estimate of the dynamic map: the max-weight particle's (recoverSlamState, src/main.cpp:359) */
int phdslam_map_estimate_dynamic(phdslam_t* h, phdslam_gaussian4d_t* out, int cap, int* n);

/* One 64-bit checksum per local particle over what a resampled copy carries (src/slamtypes.h:313-333): pose, map size,
 * map components, CPHD cardinality row, and the dynamic map of the mixed feature model.  An offspring's checksum equals its ancestor's, whichever GPU owned the ancestor
 * (bench.py's exchange_check, tests/test_dist_gpu.py). */
int phdslam_particle_checksums(phdslam_t* h, unsigned long long* out /* n_local */);

/* Dense GM-PHD update of the current state against z WITHOUT prune/merge or weight update:
 * the features_update array of phdUpdateKernel (src/phdfilter.cu:2083-2321) in the reference's order
 * [non-detect C | detect m-major M*C | birth M] per particle, plus the per-particle log-weight increment
 * and in-range counts.  terms_out holds sum_p (C_p*(M+1)+M) gaussians (may be NULL to only time it). */
int phdslam_update_terms(phdslam_t* h, const float* z, int M, int fields, phdslam_gaussian2d_t* terms_out,
                         size_t cap, int* n_in_range_out, float* dlogw_out);

/* ---- measurement ---- */
typedef struct phdslam_timings {
  float predict_ms, update_ms, merge_ms, weights_ms, estimate_ms, resample_ms;
  unsigned long long launches; /* kernels launched by this handle so far */
  unsigned long long migrated_in; /* particles received from other ranks by resampling so far */
  unsigned long long h2d_bytes;   /* bytes this handle copied host -> device so far (counted at every copy call) */
  unsigned long long d2h_bytes;   /* bytes this handle copied device -> host so far */
  float dynamic_ms;               /* mixed feature model: the dynamic-map kernels of the last update (not part of update_ms / merge_ms) */
} phdslam_timings_t;
/* CUDA-event timings of the most recent call of each phase (events on the handle's stream). */
int phdslam_get_timings(phdslam_t* h, phdslam_timings_t* out);
/* Raw stream handle (cudaStream_t) so a caller can record its own events on the stream kernels run on. */
void* phdslam_stream(phdslam_t* h);
int phdslam_synchronize(phdslam_t* h);
/* on = 1: the prune + merge of one sub-batch of particles runs (on a second, higher-priority stream) under the GM-PHD
 * update of the next sub-batch; the results are identical.  Off by default: on B200 the two kernels do not share an SM
 * profitably.  With overlap, phdslam_timings_t.update_ms is the span of the update kernels and
 * merge_ms the part of the merge that is still running when the last update kernel ends. */
int phdslam_set_overlap(phdslam_t* h, int on);
/* 1 when phdslam_dist_init mapped every peer's particle buffers over NVLink (CUDA IPC) and the global resampling
 * exchange is the fused gather-and-push kernel; 0 when it runs the NCCL send/recv ring (PHDSLAM_P2P=0, or no peer
 * access between the GPUs).  Both give bit-identical particles. */
int phdslam_dist_p2p(const phdslam_t* h);
/* Imports n_src particles (poses, log-weights, maps) as the first local particles and fills [n_src, n_local) with copies
 * of them (cyclically) on the device: a benchmark builds a 16 M-particle scene from one it can afford to generate. */
int phdslam_import_tiled(phdslam_t* h, int n_src, const phdslam_pose_t* poses, const float* log_weights, const int* sizes,
                         const phdslam_gaussian2d_t* maps);
/* Snapshot / restore of the whole device state inside the handle (bench: identical work every step). */
int phdslam_snapshot(phdslam_t* h);
int phdslam_restore(phdslam_t* h);

/* ---- host-only helpers shared by the CLI and the tests (no GPU) ---- */
/* reference: loadMeasurements/parseMeasurements (main.cpp:192-244), loadControls (:169-190).
 * Returns the number of steps; *data is malloc'd (free with phdslam_free): for measurements, offsets[n_steps+1]
 * index into floats of `fields` per record. */
int phdslam_load_measurements(const char* path, int fields, int has_header, float** data, int** offsets, int* n_steps);
int phdslam_load_controls(const char* path, float** data /* n x {v_encoder, alpha} */, int* n);
/* reference: loadTimestamps (src/main.cpp:147-167): one time per line.  A missing file is not an error: *n = 0
 * (the reference then runs without time stamps, main.cpp:1091-1093). */
int phdslam_load_timestamps(const char* path, double** data, int* n);
/* reference: loadTrajectory (src/main.cpp:247-264): lines starting with '%' skipped, six numbers per pose. */
int phdslam_load_trajectory(const char* path, phdslam_pose_t** data, int* n);

/* The input schedule of run_synth for time-stamped, asynchronous measurement and control streams
 * (src/main.cpp:1187-1230), one event per time step:
 *   measurement earlier than the next control -> update with that measurement set, the control is kept
 *   equal time stamps                         -> take the control and the measurement set
 *   otherwise                                 -> take the control, no measurements
 * Every event predicts with dt = (time stamp of control c_idx) - (time of the previous event) -- the reference reads the
 * CONTROL time stamp in all three branches -- and the loop ends when either stream is exhausted (:1189-1192). */
typedef struct phdslam_event {
  int z_idx;        /* measurement set to update with, -1 = none */
  int c_idx;        /* control to take, -1 = keep the current one */
  float dt;         /* config.dt of this step (setDeviceConfig, :1199-1200) */
} phdslam_event_t;
/* events[cap]; returns the number of events (<= nz + nc, main.cpp:1116), or a negative status */
int phdslam_plan_events(const double* measurement_times, int nz, const double* control_times, int nc, phdslam_event_t* events,
                        int cap);
void phdslam_free(void* p);
/* reference: writeLog (main.cpp:848-954) in the README 5-line layout (README:31-39) or the 7-line one. */
int phdslam_write_log(const char* path, int layout, const phdslam_pose_t* expected, const phdslam_gaussian2d_t* map,
                      int n_map, const float* log_weights, const phdslam_pose_t* poses, int n_particles,
                      const int* resample_idx, const float* cardinality, int n_card, int filter_type);
/* the same with the dynamic map estimate on line 3 of the 7-line layout (main.cpp:885-900; the 5-line layout has no such line) */
int phdslam_write_log_mixed(const char* path, int layout, const phdslam_pose_t* expected, const phdslam_gaussian2d_t* map,
                            int n_map, const phdslam_gaussian4d_t* map_dynamic, int n_map_dynamic, const float* log_weights,
                            const phdslam_pose_t* poses, int n_particles, const int* resample_idx, const float* cardinality,
                            int n_card, int filter_type);

#ifdef __cplusplus
}
#endif
#endif /* PHDSLAM_H */
