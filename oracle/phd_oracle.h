/*
 * phd_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A scalar C++ restatement of the reference's RB-PHD-SLAM filter step
 * (cheesinglee/cuda-PHDSLAM), function by function, each citing the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libphdslam.so) never does.
 *
 * PARITY PIN: the reference ships no tests, golden vectors or fixtures
 *  and its named CPU path (src/scphd_cpu.cpp) is an empty stub.
 * The pin is therefore (a) oracle/_ref: the reference's own CUDA kernels and
 * host functions of this path, cut verbatim out of /root/reference/src by
 * oracle/ref_build.sh and run on the CPU through a CUDA execution-model emulator
 * (oracle/ref_shim/cuda_emul.h; two documented barrier insertions, see the
 * recipe), compared with this oracle in tests/test_ref_pin.py, with the outputs
 * committed as tests/golden/ref_golden.npz (generator:
 * tests/golden/make_ref_golden.py); (b) closed-form known-answer tests
 * (tests/test_oracle_kat.py).  CPHD is dead code at the reference's HEAD
 * (commented out line by line) and live, in an older form, in
 * src/phdfilter.cu.bak; ref_build.sh makes both executable (HEAD: the leading
 * "//" of every line removed; .bak: verbatim) and tests/test_cphd_ref_pin.py
 * compares this oracle with their outputs (golden: tests/golden/
 * ref_cphd_golden.npz): detection / non-detection weights, means, covariances,
 * log<Psi0,p> and the posterior cardinality within 1e-4.  The mixed feature
 * model (static + constant-velocity features, feature_model = 2) is pinned the
 * same way: the reference's computeBirth / computePreUpdate on Gaussian4D,
 * predictMapKernelMixed, phdUpdateKernelMixed and the Gaussian4D instance of
 * phdUpdateMergeKernel run through the emulator (tests/test_mixed_ref_pin.py,
 * tests/golden/ref_mixed_golden.npz, generator make_ref_mixed_golden.py).
 *
 * Arithmetic: fp32 in the reference's operation order, with the transcendental
 * functions of include/phd_detmath.h so that results are bit-reproducible on the
 * GPU.  Deviations from the reference's literal arithmetic are listed in
 * DESIGN.md section "Canonical arithmetic".
 */
#ifndef PHD_ORACLE_H
#define PHD_ORACLE_H

#include "../include/phdslam.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct phd_oracle phd_oracle_t;

phd_oracle_t* oracle_create(const phdslam_config_t* cfg);
void oracle_destroy(phd_oracle_t* o);
void oracle_set_config(phd_oracle_t* o, const phdslam_config_t* cfg);
void oracle_set_threads(phd_oracle_t* o, int n_threads);

int oracle_n_particles(const phd_oracle_t* o);
void oracle_get_poses(const phd_oracle_t* o, phdslam_pose_t* out);
void oracle_set_poses(phd_oracle_t* o, const phdslam_pose_t* in);
void oracle_get_log_weights(const phd_oracle_t* o, float* out);
void oracle_set_log_weights(phd_oracle_t* o, const float* in);
void oracle_get_map_sizes(const phd_oracle_t* o, int* out);
void oracle_get_maps(const phd_oracle_t* o, phdslam_gaussian2d_t* out);
void oracle_set_maps(phd_oracle_t* o, const int* sizes, const phdslam_gaussian2d_t* in);
/* mixed feature model (feature_model = 2): SynthSLAM::maps_dynamic */
void oracle_get_map_sizes_dynamic(const phd_oracle_t* o, int* out);
void oracle_get_maps_dynamic(const phd_oracle_t* o, phdslam_gaussian4d_t* out);
void oracle_set_maps_dynamic(phd_oracle_t* o, const int* sizes, const phdslam_gaussian4d_t* in);
void oracle_get_resample_idx(const phd_oracle_t* o, int* out);
void oracle_get_cardinalities(const phd_oracle_t* o, float* out);
void oracle_set_cardinalities(phd_oracle_t* o, const float* in);

/* phdPredict, one sub-step (src/phdfilter.cu:1080-1257) */
void oracle_predict(phd_oracle_t* o, const float* control, const double* draws);
/* phdUpdateSynth (src/phdfilter.cu:3336-3761) */
void oracle_update(phd_oracle_t* o, const float* z, int M, int fields);
/* dense update terms only (preUpdateSynthKernel + phdUpdateKernel), no state change */
size_t oracle_update_terms(phd_oracle_t* o, const float* z, int M, int fields, phdslam_gaussian2d_t* terms_out,
                           size_t cap, int* n_in_range_out, float* dlogw_out);
/* recoverSlamState + nEff (src/main.cpp:318-388,1281-1284) */
void oracle_estimate(phd_oracle_t* o, phdslam_estimate_t* out);
int oracle_map_estimate(phd_oracle_t* o, int which, phdslam_gaussian2d_t* out, int cap);
/* resampleParticles (src/main.cpp:453-501); mode 0 = canonical integer CDF, 1 = literal double CDF walk */
void oracle_resample(phd_oracle_t* o, int n_new, const double* uniforms, int literal, int* ancestors_out);
/* run_synth loop body (src/main.cpp:1231-1297) */
void oracle_step(phd_oracle_t* o, int step_index, const float* control, const float* z, int M, int fields,
                 phdslam_estimate_t* est_out, int* resampled_out);

void oracle_step_filter(phd_oracle_t* o, int step_index, const float* control, const float* z, int M, int fields,
                        phdslam_estimate_t* est_out);
void oracle_step_resample(phd_oracle_t* o, int M, const phdslam_estimate_t* est, int* resampled_out);

/* pieces exposed for known-answer tests */
float oracle_mahalanobis(const phdslam_gaussian2d_t* a, const phdslam_gaussian2d_t* b);
float oracle_hellinger(const phdslam_gaussian2d_t* a, const phdslam_gaussian2d_t* b);
int oracle_merge(const phdslam_config_t* cfg, const phdslam_gaussian2d_t* in, int n, phdslam_gaussian2d_t* out);
int oracle_reduce_mixture(const phdslam_gaussian2d_t* in, int n, float min_distance, phdslam_gaussian2d_t* out);
/* mixed feature model: computeMahalDist(Gaussian4D,..), phdUpdateMergeKernel<Gaussian4D>, predictMapKernelMixed for one
 * feature, and the dense update terms of both maps of one particle (phdUpdateKernelMixed) */
float oracle_mahalanobis4(const phdslam_gaussian4d_t* a, const phdslam_gaussian4d_t* b);
int oracle_merge4(const phdslam_config_t* cfg, const phdslam_gaussian4d_t* in, int n, phdslam_gaussian4d_t* out);
void oracle_predict_feature4(const phdslam_config_t* cfg, const phdslam_gaussian4d_t* in, phdslam_gaussian4d_t* out);
float oracle_mixed_terms(const phdslam_config_t* cfg, const phdslam_pose_t* pose, const phdslam_gaussian2d_t* smap, int ns,
                         const phdslam_gaussian4d_t* dmap, int nd, const float* z, int M, int fields,
                         phdslam_gaussian2d_t* s_terms_out, int* n_s_terms, phdslam_gaussian4d_t* d_terms_out, int* n_d_terms);
float oracle_warp_sum(const float* v, int n);
void oracle_detmath(int fn, const float* x, const float* y, float* out, float* out2, int n);
void oracle_philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1, unsigned* out4);
void oracle_esf(const double* roots, int n, double* out /* n+1 */);
/* CPHD birth cardinality pb[M+1] and predicted cardinality pm[N1] (cardinalityPredictKernel, src/phdfilter.cu:867-888) */
void oracle_cphd_predict_cardinality(const phdslam_config_t* cfg, const float* prior, int N1, int M, float* pb_out, float* pm_out);
/* CPHD multi-object terms of one particle (see phd_oracle.cpp: cphd_factors) */
void oracle_cphd_factors(const phdslam_config_t* cfg, const float* w, const float* pd, int C, const float* S, int M,
                         const float* prior, int N1, float* D /* M */, float* ND, float* inc, float* card_out /* N1 */);

#ifdef __cplusplus
}
#endif
#endif
