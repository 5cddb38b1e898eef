/*
 * phdfilter_b200.cpp -- drop-in for the reference's src/phdfilter.cu.
 *
 * A maintainer of cheesinglee/cuda-PHDSLAM switches the static-map PHD / CPHD path to this library by compiling THIS file
 * instead of src/phdfilter.cu (next to the reference's own main.cpp, slamtypes.h, phdfilter.h) and linking -lphdslam.
 * It defines the functions of src/phdfilter.h:10-34 that src/phdfilter.cu defines and run_synth calls (initRandomNumberGenerators,
 * setDeviceConfig, phdPredict, phdUpdateSynth), with the reference's types; the device handle lives
 * across calls.  run_synth owns the particle set on the HOST and edits it between the calls -- resampleParticles replaces it
 * (src/main.cpp:1286-1289), follow_trajectory overwrites the poses (:1239-1243), the non-resampling branch resets
 * resample_idx (:1293-1296) -- so every entry point first compares the host particles with a shadow of what the device last
 * handed back (poses, weights, resample indices, particle count: O(N) words, no maps) and re-uploads the set, maps and
 * CPHD cardinalities included, when the host has touched it.  With the host resampler of the reference that is one map
 * upload per resampling step (the reference itself re-uploads every map on every call, src/phdfilter.cu:2901-3103);
 * resampleParticlesDevice() below avoids it.
 *
 * tests/test_shim_compiles.py compiles it against the reference's headers and links it against libphdslam.so;
 * oracle/ref_build.sh links it with the reference's OWN run_synth loop, resampleParticles, recoverSlamState and writeLog
 * (cut verbatim from src/main.cpp) into oracle/_ref/shim_replay, which tests/test_shim_gpu.py runs on the GPU against the CLI.
 */
class MotionModel; /* src/slamtypes.h:335 names it without declaring it */
#include "slamtypes.h"
#include "phdfilter.h"
#include <phdslam.h> /* include/phdslam.h of this repository */
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

extern SlamConfig config; /* src/main.cpp:80 */
static phdslam_t* g_h = nullptr;

static_assert(sizeof(Gaussian2D) == sizeof(phdslam_gaussian2d_t), "Gaussian2D layout (src/slamtypes.h:123-127)");
static_assert(sizeof(ConstantVelocityState) == sizeof(phdslam_pose_t), "ConstantVelocityState layout (src/slamtypes.h:44-51)");

static void check(int rc, const char* what) {
  if (rc != 0) { /* the reference's checkCudaErrors prints and exits (helper_cuda.h) */
    fprintf(stderr, "%s: %s\n", what, phdslam_last_error());
    exit(1);
  }
}

static phdslam_config_t to_cfg(const SlamConfig& c) { /* field for field, src/slamtypes.h:142-250 */
  phdslam_config_t k;
  phdslam_config_defaults(&k);
  k.x0 = c.x0; k.y0 = c.y0; k.yaw0 = c.yaw0; k.vx0 = c.vx0; k.vy0 = c.vy0; k.vyaw0 = c.vyaw0;
  k.motion_type = c.motionType; k.ax = c.ax; k.ay = c.ay; k.ayaw = c.ayaw; k.dt = c.dt;
  k.min_range = c.minRange; k.max_range = c.maxRange; k.max_bearing = c.maxBearing;
  k.std_range = c.stdRange; k.std_bearing = c.stdBearing;
  k.clutter_rate = c.clutterRate; k.clutter_density = c.clutterDensity; k.pd = c.pd;
  k.n_particles = c.n_particles; k.n_predict_particles = c.nPredictParticles; k.subdivide_predict = c.subdividePredict;
  k.resample_threshold = c.resampleThresh; k.birth_weight = c.birthWeight; k.birth_noise_factor = c.birthNoiseFactor;
  k.min_separation = c.minSeparation; k.min_feature_weight = c.minFeatureWeight; k.particle_weighting = c.particleWeighting;
  k.distance_metric = c.distanceMetric; k.max_cardinality = c.maxCardinality; k.filter_type = c.filterType;
  k.map_estimate = c.mapEstimate; k.feature_model = c.featureModel;
  k.l = c.l; k.h = c.h; k.a = c.a; k.b = c.b; k.std_encoder = c.stdEncoder; k.std_alpha = c.stdAlpha;
  k.labeled_measurements = c.labeledMeasurements;
  /* mixed feature model (featureModel = MIXED_MODEL) */
  k.ps = c.ps; k.tau = c.tau; k.beta = c.beta; k.std_ax_features = c.stdAxMap; k.std_ay_features = c.stdAyMap;
  k.cov_vx_birth = c.covVxBirth; k.cov_vy_birth = c.covVyBirth;
  return k;
}

/* extensions of phdslam_config_t that SlamConfig has no field for (Philox seed, map capacity): set before the first
 * setDeviceConfig by a host that wants other values than the defaults */
unsigned long long phdslam_shim_seed = 0;
int phdslam_shim_max_components = 0;

void initRandomNumberGenerators() {} /* src/phdfilter.cu:142: the library's Philox generator is counter based, nothing to seed per thread */

void setDeviceConfig(const SlamConfig& c) { /* src/phdfilter.cu:3885-3890 */
  phdslam_config_t k = to_cfg(c);
  k.seed = phdslam_shim_seed;
  if (phdslam_shim_max_components > 0) k.max_components = phdslam_shim_max_components;
  if (!g_h) check(phdslam_create(&k, 0, &g_h), "phdslam_create");
  else check(phdslam_set_config(g_h, &k), "phdslam_set_config");
}

/* ---- what the device last handed to the host: if the host particles still equal it, the device state is current ---- */
static struct {
  bool valid;
  std::vector<ConstantVelocityState> states;
  std::vector<REAL> weights;
  std::vector<int> resample_idx;
} g_shadow;

static bool host_untouched(const SynthSLAM& p) {
  const size_t n = g_shadow.states.size();
  return g_shadow.valid && (size_t)p.n_particles == n && p.states.size() == n && p.weights.size() == n &&
         p.resample_idx.size() == n && p.maps_static.size() == n &&
         memcmp(&p.states[0], &g_shadow.states[0], n * sizeof(ConstantVelocityState)) == 0 &&
         memcmp(&p.weights[0], &g_shadow.weights[0], n * sizeof(REAL)) == 0 &&
         memcmp(&p.resample_idx[0], &g_shadow.resample_idx[0], n * sizeof(int)) == 0;
}
static void remember(const SynthSLAM& p) {
  g_shadow.states = p.states;
  g_shadow.weights = p.weights;
  g_shadow.resample_idx = p.resample_idx;
  g_shadow.valid = true;
}

/* host -> device: the whole particle set (run_synth initialises it on the host, src/main.cpp:1129-1144, and replaces it
 * when it resamples) */
static void push(SynthSLAM& p) {
  const int n = p.n_particles;
  if (n != phdslam_n_local(g_h)) check(phdslam_set_particle_count(g_h, n), "phdslam_set_particle_count");
  std::vector<int> sizes;
  std::vector<phdslam_gaussian2d_t> flat;
  for (int i = 0; i < n; ++i) {
    sizes.push_back((int)p.maps_static[i].size());
    for (size_t j = 0; j < p.maps_static[i].size(); ++j) {
      phdslam_gaussian2d_t g;
      memcpy(&g, &p.maps_static[i][j], sizeof(g));
      flat.push_back(g);
    }
  }
  flat.push_back(phdslam_gaussian2d_t());
  check(phdslam_set_poses(g_h, (const phdslam_pose_t*)&p.states[0]), "phdslam_set_poses");
  check(phdslam_set_log_weights(g_h, &p.weights[0]), "phdslam_set_log_weights");
  check(phdslam_set_maps(g_h, sizes.data(), flat.data()), "phdslam_set_maps");
  if (config.featureModel == MIXED_MODEL) {   /* maps_dynamic travel like maps_static */
    static_assert(sizeof(phdslam_gaussian4d_t) == sizeof(Gaussian4D), "Gaussian4D layout");
    std::vector<int> dsizes;
    std::vector<phdslam_gaussian4d_t> dflat;
    for (int i = 0; i < n; ++i) {
      dsizes.push_back((int)p.maps_dynamic[i].size());
      for (size_t j = 0; j < p.maps_dynamic[i].size(); ++j) {
        phdslam_gaussian4d_t g;
        memcpy(&g, &p.maps_dynamic[i][j], sizeof(g));
        dflat.push_back(g);
      }
    }
    dflat.push_back(phdslam_gaussian4d_t());
    check(phdslam_set_maps_dynamic(g_h, dsizes.data(), dflat.data()), "phdslam_set_maps_dynamic");
  }
  if (config.filterType == CPHD_TYPE) {
    const size_t n1 = (size_t)config.maxCardinality + 1;
    std::vector<float> card((size_t)n * n1);
    for (int i = 0; i < n; ++i)
      for (size_t k = 0; k < n1 && k < p.cardinalities[i].size(); ++k) card[i * n1 + k] = p.cardinalities[i][k];
    check(phdslam_set_cardinalities(g_h, card.data()), "phdslam_set_cardinalities");
  }
}

static void sync_to_device(SynthSLAM& p) {
  if (!host_untouched(p)) push(p);
}

/* device -> host: what run_synth reads between calls (weights for nEff and the log, poses for the log, maps for
 * recoverSlamState / writeParticlesMat, cardinalities for the CPHD estimate) */
static void pull(SynthSLAM& p, bool take_resample_idx = false) {
  const int n = phdslam_n_local(g_h);
  /* resample_idx belongs to the host loop (resampleParticles sets it, the non-resampling branch resets it, src/main.cpp:
   * 1286-1297; phdPredict and phdUpdateSynth never touch it): the device's copy is taken only after a device resampling
   * or when the particle count changed (shotgun prediction: every copy inherits its parent's index, :1185-1238) */
  take_resample_idx = take_resample_idx || (int)p.resample_idx.size() != n;
  p.states.resize(n); p.weights.resize(n); p.maps_static.resize(n); p.resample_idx.resize(n);
  p.maps_dynamic.resize(n); p.cardinalities.resize(n); p.variances.resize(n);
  p.n_particles = n;
  check(phdslam_get_poses(g_h, (phdslam_pose_t*)&p.states[0]), "phdslam_get_poses");
  check(phdslam_get_log_weights(g_h, &p.weights[0]), "phdslam_get_log_weights");
  if (take_resample_idx) check(phdslam_get_resample_idx(g_h, &p.resample_idx[0]), "phdslam_get_resample_idx");
  std::vector<int> sizes(n);
  check(phdslam_get_map_sizes(g_h, sizes.data()), "phdslam_get_map_sizes");
  size_t total = 0;
  for (int i = 0; i < n; ++i) total += (size_t)sizes[i];
  std::vector<phdslam_gaussian2d_t> flat(total + 1);
  check(phdslam_get_maps(g_h, flat.data(), total + 1), "phdslam_get_maps");
  size_t k = 0;
  for (int i = 0; i < n; ++i) {
    p.maps_static[i].resize(sizes[i]);
    for (int j = 0; j < sizes[i]; ++j, ++k) memcpy(&p.maps_static[i][j], &flat[k], sizeof(Gaussian2D));
  }
  if (config.featureModel == MIXED_MODEL) {
    std::vector<int> dsizes(n);
    check(phdslam_get_map_sizes_dynamic(g_h, dsizes.data()), "phdslam_get_map_sizes_dynamic");
    size_t dtotal = 0;
    for (int i = 0; i < n; ++i) dtotal += (size_t)dsizes[i];
    std::vector<phdslam_gaussian4d_t> dflat(dtotal + 1);
    check(phdslam_get_maps_dynamic(g_h, dflat.data(), dtotal + 1), "phdslam_get_maps_dynamic");
    size_t q = 0;
    for (int i = 0; i < n; ++i) {
      p.maps_dynamic[i].resize(dsizes[i]);
      for (int j = 0; j < dsizes[i]; ++j, ++q) memcpy(&p.maps_dynamic[i][j], &dflat[q], sizeof(Gaussian4D));
    }
  }
  if (config.filterType == CPHD_TYPE) {
    const size_t n1 = (size_t)config.maxCardinality + 1;
    std::vector<float> card((size_t)n * n1);
    check(phdslam_get_cardinalities(g_h, card.data()), "phdslam_get_cardinalities");
    for (int i = 0; i < n; ++i) p.cardinalities[i].assign(card.begin() + i * n1, card.begin() + (i + 1) * n1);
  }
  remember(p);
}

void phdPredict(SynthSLAM& particles, ...) { /* src/phdfilter.cu:1080-1257 */
  sync_to_device(particles);
  float u[2] = {0.0f, 0.0f};
  if (config.motionType == ACKERMAN_MOTION) { /* the reference passes the control through C varargs (:1141-1144) */
    va_list ap;
    va_start(ap, particles);
    AckermanControl c = va_arg(ap, AckermanControl);
    va_end(ap);
    u[0] = c.v_encoder;
    u[1] = c.alpha;
  }
  check(phdslam_predict(g_h, u, nullptr), "phdslam_predict");
  pull(particles);
}

SynthSLAM phdUpdateSynth(SynthSLAM& particles, measurementSet Z) { /* src/phdfilter.cu:3336-3761 */
  sync_to_device(particles);
  SynthSLAM before = particles; /* the reference returns the pre-merge copy (particlesPreMerge, src/main.cpp:1268) */
  std::vector<float> z;
  for (size_t i = 0; i < Z.size(); ++i) {
    z.push_back(Z[i].range);
    z.push_back(Z[i].bearing);
    z.push_back((float)Z[i].label);
  }
  z.push_back(0.0f);
  int rc = phdslam_update(g_h, z.data(), (int)Z.size(), 3);
  if (rc != PHDSLAM_ERR_NAN) check(rc, "phdslam_update"); /* NaN weights: run_synth notices through nEff and breaks (:1307) */
  pull(particles);
  return before;
}

/* recoverSlamState(SynthSLAM&, ...) is declared in src/phdfilter.h but DEFINED in src/main.cpp:318-388, on the host copy of
 * the particles -- which pull() keeps current, so the reference's own definition keeps working unchanged (including its
 * host EAP reduction, src/gm_reduce.cpp).  This is the device version (fixed-point expected pose, MAP map copy, EAP map
 * reduced on the GPU); a maintainer calls it instead at src/main.cpp:1274 to skip the host O(n^2) reduction. */
void recoverSlamStateDevice(SynthSLAM& particles, ConstantVelocityState& expectedPose, vector<REAL>& cn_estimate) {
  sync_to_device(particles);
  phdslam_estimate_t e;
  check(phdslam_estimate(g_h, &e), "phdslam_estimate");
  memcpy(&expectedPose, &e.expected_pose, sizeof(expectedPose));
  const int cap = 1 << 16;
  std::vector<phdslam_gaussian2d_t> m(cap);
  for (int which = 1; which <= 2; which <<= 1) { /* bit 1: MAP map, bit 2: EAP map */
    if (!(config.mapEstimate & which)) continue;
    int n = 0;
    check(phdslam_map_estimate(g_h, which, m.data(), cap, &n), "phdslam_map_estimate");
    vector<Gaussian2D>& out = (which == 1) ? particles.max_map_static : particles.exp_map_static;
    out.resize(n);
    for (int i = 0; i < n; ++i) memcpy(&out[i], &m[i], sizeof(Gaussian2D));
  }
  if (config.filterType == CPHD_TYPE && e.map_particle >= 0) { /* cardinality distribution of the maximum-weight particle (:356-361) */
    const int n1 = config.maxCardinality + 1, np = phdslam_n_local(g_h);
    std::vector<float> all((size_t)np * n1);
    check(phdslam_get_cardinalities(g_h, all.data()), "phdslam_get_cardinalities");
    cn_estimate.assign(all.begin() + (size_t)e.map_particle * n1, all.begin() + (size_t)(e.map_particle + 1) * n1);
  }
}

/* Device twin of resampleParticles (src/main.cpp:453-501): the maps never leave the GPU.  A maintainer replaces
 *   particles = resampleParticles(particles, config.n_particles) ;         (src/main.cpp:1289)
 * with
 *   resampleParticlesDevice(particles, config.n_particles) ;
 * (same stratified draws, from the library's counter-based generator instead of rng.cpp). */
void resampleParticlesDevice(SynthSLAM& particles, int n_new) {
  sync_to_device(particles);
  check(phdslam_resample(g_h, n_new, nullptr, nullptr), "phdslam_resample");
  pull(particles, true);
}
