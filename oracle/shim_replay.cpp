/*
 * shim_replay.cpp -- the reference's OWN run_synth, linked against the drop-in shim.
 *
 * TEST INFRASTRUCTURE (built by oracle/ref_build.sh into oracle/_ref/shim_replay, only where /root/reference exists; run by
 * tests/test_shim_gpu.py on the GPU box).  It answers the question a maintainer of cheesinglee/cuda-PHDSLAM asks: "if I
 * compile cuda-phdslam_b200/shim/phdfilter_b200.cpp instead of src/phdfilter.cu, does my unmodified host loop still work?"
 *
 * From the reference, verbatim (cut by line range into oracle/_ref/gen/ by ref_build.sh, nothing is copied into the
 * repository): the body of run_synth (src/main.cpp:1076-1313: input loading, particle initialisation, the time-step loop
 * with its host-side resampleParticles and resample_idx bookkeeping), loadMeasurements / loadControls / loadTimestamps /
 * loadTrajectory, computeExpectedMap, recoverSlamState, resampleParticles, writeLog, and src/gm_reduce.cpp.
 * From this repository: the shim (phdPredict, phdUpdateSynth, setDeviceConfig, initRandomNumberGenerators over the C-ABI).
 * Ours, in this file: main() (loadConfig needs boost::program_options; the library's parser fills the same SlamConfig),
 * randu01() (the reference's is boost::mt19937 seeded with time(0), src/rng.cpp:23-35; here the library's counter-based
 * stratified draws, so that the host resampler of the reference and the device resampler of the CLI see the same
 * uniforms), and two stubs: boost::archive::binary_oarchive (the state100.bin dump of :1262-1269) and writeParticlesMat
 * (matio; DEBUG builds call it right after recoverSlamState, :1276-1279 -- the one place the loop exports its state --
 * so it is mapped onto the reference's own writeLog, which HEAD never calls).
 */
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

class MotionModel; /* src/slamtypes.h:335 names it without declaring it */
#include "slamtypes.h"
#include "phdfilter.h"
#include "phdslam.h"
#include "phd_detmath.h"

using namespace std;

/* globals of src/main.cpp:56-84 */
SlamConfig config;
size_t deviceMemLimit;
std::string config_filename;
std::string data_dir;
time_t rawtime;
struct tm* timeinfo;
timeval start, stop;
char timestamp[80];
REAL current_time = 0;
REAL last_time = 0;
int n_steps = -1;

#define DEBUG
#define DEBUG_MSG(x)
#define DEBUG_VAL(x)

/* the library's stratified resampling draws (include/phd_detmath.h; oracle_resample spells the same): the reference's
 * resampleParticles draws one uniform and discards it, then one per offspring (src/main.cpp:461-469) */
extern unsigned long long phdslam_shim_seed;
static unsigned g_resample_call = 0;
static long g_draw = -1;
extern "C" double randu01() {
  const long j = g_draw++;
  if (j < 0) return 0.5; /* the discarded draw */
  phd_philox4_t r = phd_philox4x32_10((uint32_t)j, g_resample_call, PHD_STREAM_RESAMPLE, 0u, (uint32_t)phdslam_shim_seed,
                                      (uint32_t)(phdslam_shim_seed >> 32));
  return phd_u01d(r.v[0], r.v[1]);
}

/* src/gm_reduce.cpp over the Eigen stand-in (as oracle/ref_harness.cpp) */
#include "eigen3/Eigen/Core"
#include "eigen3/Eigen/Cholesky"
#include "eigen3/Eigen/LU"
#include "ref_gm_reduce.inc"
#include "ref_host.inc"

/* state100.bin (src/main.cpp:1262-1269) needs boost::serialization: the dump is not part of the filter */
namespace boost { namespace archive {
struct binary_oarchive {
  explicit binary_oarchive(std::ofstream&) {}
  template <class T> binary_oarchive& operator<<(const T&) { return *this; }
};
struct binary_iarchive {
  explicit binary_iarchive(std::ifstream&) {}
  template <class T> binary_iarchive& operator>>(T&) { return *this; }
};
}}

static void replay_log(SynthSLAM& particles, int n, ConstantVelocityState& expectedPose, vector<REAL>& cn_estimate) {
  vector<Gaussian4D> none;
  vector<Gaussian2D>& map = (config.mapEstimate & 2) ? particles.exp_map_static : particles.max_map_static;
  writeLog(particles, expectedPose, map, none, particles.resample_idx, cn_estimate, n);
}
static SynthSLAM replay_resample(SynthSLAM& p, int n_new) {
  g_draw = -1;
  SynthSLAM q = resampleParticles(p, n_new);
  g_resample_call++;
  return q;
}
#define writeParticlesMat(p, ...) replay_log(p, n, expectedPose, cn_estimate)
#define resampleParticles(p, n_new) replay_resample(p, n_new)

void run_synth(bool profile_run) {
#include "ref_run_synth_body.inc" /* src/main.cpp:1076-1313, verbatim */
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <config.cfg> <out_dir> [key=value ...]\n", argv[0]);
    return 2;
  }
  phdslam_config_t k;
  if (phdslam_config_load(argv[1], &k) != 0) {
    fprintf(stderr, "%s\n", phdslam_last_error());
    return 2;
  }
  for (int i = 3; i < argc; ++i) {
    std::string kv = argv[i];
    size_t eq = kv.find('=');
    if (eq == std::string::npos || phdslam_config_set(&k, kv.substr(0, eq).c_str(), kv.substr(eq + 1).c_str()) != 0) return 2;
  }
  /* the SlamConfig loadConfig would have produced (src/main.cpp:956-1073) */
  memset(&config, 0, sizeof(config));
  config.x0 = k.x0; config.y0 = k.y0; config.yaw0 = k.yaw0; config.vx0 = k.vx0; config.vy0 = k.vy0; config.vyaw0 = k.vyaw0;
  config.motionType = k.motion_type; config.ax = k.ax; config.ay = k.ay; config.ayaw = k.ayaw; config.dt = k.dt;
  config.minRange = k.min_range; config.maxRange = k.max_range; config.maxBearing = k.max_bearing;
  config.stdRange = k.std_range; config.stdBearing = k.std_bearing;
  config.clutterRate = k.clutter_rate; config.clutterDensity = k.clutter_density; config.pd = k.pd;
  config.n_particles = k.n_particles; config.nPredictParticles = k.n_predict_particles; config.subdividePredict = k.subdivide_predict;
  config.resampleThresh = k.resample_threshold; config.birthWeight = k.birth_weight; config.birthNoiseFactor = k.birth_noise_factor;
  config.minSeparation = k.min_separation; config.minFeatureWeight = k.min_feature_weight;
  config.particleWeighting = k.particle_weighting; config.distanceMetric = k.distance_metric;
  config.maxCardinality = k.max_cardinality; config.filterType = k.filter_type; config.mapEstimate = k.map_estimate;
  config.featureModel = k.feature_model; config.labeledMeasurements = k.labeled_measurements != 0;
  config.l = k.l; config.h = k.h; config.a = k.a; config.b = k.b; config.stdAlpha = k.std_alpha; config.stdEncoder = k.std_encoder;
  config.followTrajectory = k.follow_trajectory != 0;
  config.savePrediction = false;
  config.nSamples = 256;
  data_dir = k.data_directory;
  n_steps = k.n_steps > 0 ? k.n_steps : k.max_steps;
  phdslam_shim_seed = k.seed;
  extern int phdslam_shim_max_components;
  phdslam_shim_max_components = k.max_components;
  if (chdir(argv[2]) != 0) {     /* writeLog and loopTime.log write into the working directory */
    perror(argv[2]);
    return 2;
  }
  initRandomNumberGenerators(); /* src/main.cpp:1464-1465 */
  setDeviceConfig(config);
  run_synth(false);             /* :1496 */
  return 0;
}
