/*
 * phdslam -- host driver over the C-ABI of libphdslam.so.
 *
 * Stands in for the reference's `./bin/cuda-PHDSLAM <config.cfg> [synth]` (src/main.cpp:1442-1514) and its
 * run_synth() time-step loop (src/main.cpp:1075-1312): same cfg keys, same input text files, and one
 * state_estimateNNNNN.log per step in the README's 5-line layout (README:31-39; `log_layout = extended` gives the
 * 7-line layout of writeLog, main.cpp:848-954).  All filter work happens on the GPU inside phdslam_step().
 *
 *   phdslam <config.cfg> [synth] [--measurements FILE] [--controls FILE] [--out DIR] [--steps N]
 *           [--set key=value ...] [--no-header] [--quiet]
 *
 * The reference hard-wires <data_directory>/measurements.txt and controls.txt (main.cpp:1079-1086); the two
 * overrides exist because the bundled files are named measurements_synth_*.txt.  When
 * <data_directory>/measurement_times.txt exists (with control_times.txt) the streams are asynchronous and every step is
 * one event of phdslam_plan_events (main.cpp:1187-1230) with its own dt; follow_trajectory pins the (single) particle to
 * <data_directory>/traj.txt (main.cpp:1122-1127,1239-1243).  Not built: the disparity mode.
 */
#include <sys/stat.h>
#include <sys/time.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/phdslam.h"

static double now_ms() {
  struct timeval tv;
  gettimeofday(&tv, nullptr);
  return tv.tv_sec * 1e3 + tv.tv_usec * 1e-3;
}

#define CHECK(call)                                                               \
  do {                                                                            \
    int rc__ = (call);                                                            \
    if (rc__ != 0) {                                                              \
      fprintf(stderr, "phdslam: %s failed (%d): %s\n", #call, rc__, phdslam_last_error()); \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <config.cfg> [synth] [--measurements FILE] [--controls FILE] [--measurement-times FILE] "
                    "[--control-times FILE] [--trajectory FILE] [--out DIR] [--steps N] [--set key=value] [--no-header] [--quiet]\n", argv[0]);
    return 1;
  }
  phdslam_config_t cfg;
  CHECK(phdslam_config_load(argv[1], &cfg));
  std::string meas_path, ctrl_path, mt_path, ct_path, traj_path, out_dir = ".";
  int max_steps = -1, has_header = 1;
  bool quiet = false;
  for (int i = 2; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "synth") continue;
    if (a == "disparity") {
      fprintf(stderr, "phdslam: the disparity / stereo mode is outside the hot path and not built\n");
      return 1;
    }
    if (a == "--measurements" && i + 1 < argc) meas_path = argv[++i];
    else if (a == "--controls" && i + 1 < argc) ctrl_path = argv[++i];
    else if (a == "--measurement-times" && i + 1 < argc) mt_path = argv[++i];
    else if (a == "--control-times" && i + 1 < argc) ct_path = argv[++i];
    else if (a == "--trajectory" && i + 1 < argc) traj_path = argv[++i];
    else if (a == "--out" && i + 1 < argc) out_dir = argv[++i];
    else if (a == "--steps" && i + 1 < argc) max_steps = atoi(argv[++i]);
    else if (a == "--no-header") has_header = 0;
    else if (a == "--quiet") quiet = true;
    else if (a == "--set" && i + 1 < argc) {
      std::string kv = argv[++i];
      size_t eq = kv.find('=');
      if (eq == std::string::npos || phdslam_config_set(&cfg, kv.substr(0, eq).c_str(), kv.substr(eq + 1).c_str()) != 0) {
        fprintf(stderr, "phdslam: bad --set %s\n", kv.c_str());
        return 1;
      }
    } else {
      fprintf(stderr, "phdslam: unknown argument %s\n", a.c_str());
      return 1;
    }
  }
  std::string dd = cfg.data_directory;
  if (!dd.empty() && dd.back() != '/') dd += '/';
  if (meas_path.empty()) meas_path = dd + "measurements.txt";   /* main.cpp:1079 */
  if (ctrl_path.empty()) ctrl_path = dd + "controls.txt";       /* main.cpp:1084 */
  if (mt_path.empty()) mt_path = dd + "measurement_times.txt";  /* main.cpp:1088 */
  if (ct_path.empty()) ct_path = dd + "control_times.txt";      /* main.cpp:1090 */
  if (traj_path.empty()) traj_path = dd + "traj.txt";           /* main.cpp:1123 */

  float* zdata = nullptr;
  int* zoff = nullptr;
  int n_meas_steps = 0;
  CHECK(phdslam_load_measurements(meas_path.c_str(), cfg.measurement_fields, has_header, &zdata, &zoff, &n_meas_steps));
  printf("Loaded %d measurements\n", n_meas_steps);
  /* time stamps (main.cpp:1087-1117) */
  double *ztimes = nullptr, *ctimes = nullptr;
  int n_zt = 0, n_ct = 0;
  CHECK(phdslam_load_timestamps(mt_path.c_str(), &ztimes, &n_zt));
  CHECK(phdslam_load_timestamps(ct_path.c_str(), &ctimes, &n_ct));
  const bool has_timestamps = n_zt > 0;                            /* :1092 */
  /* the reference loads controls.txt for every motion model (:1084-1086); the constant-velocity model only needs it for
   * the control time stamps */
  float* udata = nullptr;
  int n_controls = 0;
  if (cfg.motion_type == 1 || has_timestamps) {
    CHECK(phdslam_load_controls(ctrl_path.c_str(), &udata, &n_controls));
    printf("Loaded %d control inputs\n", n_controls);
  }
  std::vector<phdslam_event_t> events;
  int n_steps = n_meas_steps;                                     /* main.cpp:1097 */
  if (has_timestamps) {
    if (n_zt != n_meas_steps) {
      fprintf(stderr, "mismatched measurements and measurement timestamps!\n");      /* :1104-1108 */
      return 1;
    }
    if (n_ct != n_controls) {
      fprintf(stderr, "mismatched controls and controls timestamps!\n");             /* :1109-1113 */
      return 1;
    }
    events.resize((size_t)n_zt + n_ct + 1);
    n_steps = phdslam_plan_events(ztimes, n_zt, ctimes, n_ct, events.data(), (int)events.size());
    printf("Loaded %d + %d time stamps: %d events\n", n_zt, n_ct, n_steps);
  }
  if (cfg.n_steps > 0 && n_steps > cfg.n_steps) n_steps = cfg.n_steps;   /* :1118 */
  if (max_steps > 0 && n_steps > max_steps) n_steps = max_steps;
  phdslam_pose_t* traj = nullptr;
  int n_traj = 0;
  if (cfg.follow_trajectory) {                                    /* main.cpp:1122-1127 */
    CHECK(phdslam_load_trajectory(traj_path.c_str(), &traj, &n_traj));
    cfg.n_particles = 1;                                          /* only need 1 particle */
    if (n_steps > n_traj) n_steps = n_traj;
    printf("Following the trajectory of %s (%d poses)\n", traj_path.c_str(), n_traj);
  }
  mkdir(out_dir.c_str(), 0755);

  phdslam_t* h = nullptr;
  CHECK(phdslam_create(&cfg, 0, &h));
  const int n_card = cfg.max_cardinality + 1;
  std::vector<phdslam_pose_t> poses;
  std::vector<float> logw;
  std::vector<int> ridx;
  std::vector<float> card((size_t)(cfg.filter_type == 1 ? n_card : 1)), all_card;
  std::vector<phdslam_gaussian2d_t> map_est(65536);
  std::vector<phdslam_gaussian4d_t> dyn_est(1);
  float current_u[2] = {0.0f, 0.0f};                              /* main.cpp:1166-1168 */
  printf("STARTING SIMULATION\n");
  FILE* tf = fopen((out_dir + "/loopTime.log").c_str(), "w");     /* main.cpp:1300-1305 */
  for (int n = 0; n < n_steps; ++n) {
    double t0 = now_ms();
    if (!quiet) printf("****** Time Step [%d/%d] ******\n", n, n_steps);
    int zi = n;
    if (has_timestamps) {
      /* one event of the asynchronous streams (main.cpp:1187-1230): its dt goes to the device config */
      const phdslam_event_t& e = events[n];
      zi = e.z_idx;
      if (e.c_idx >= 0) {
        current_u[0] = udata[2 * (size_t)e.c_idx];
        current_u[1] = udata[2 * (size_t)e.c_idx + 1];
      }
      cfg.dt = e.dt;
      CHECK(phdslam_set_config(h, &cfg));
    } else if (cfg.motion_type == 1 && n > 0) {
      int ci = n - 1;                                             /* current_control = all_controls[n-1], main.cpp:1234 */
      if (ci >= n_controls) ci = n_controls - 1;
      if (ci >= 0) {
        current_u[0] = udata[2 * (size_t)ci];
        current_u[1] = udata[2 * (size_t)ci + 1];
      }
    }
    const int M = (zi >= 0) ? zoff[zi + 1] - zoff[zi] : 0;
    const float* z = (zi >= 0) ? zdata + (size_t)zoff[zi] * cfg.measurement_fields : nullptr;
    int step_index = n;
    if (cfg.follow_trajectory) {                                  /* main.cpp:1239-1243: the pose comes from the file, no prediction */
      CHECK(phdslam_set_poses(h, &traj[n]));
      step_index = 0;
    }
    /* the filter half of the iteration: predict, update, recoverSlamState (main.cpp:1244-1274) */
    phdslam_estimate_t est;
    memset(&est, 0, sizeof(est));
    est.map_particle = -1;
    int resampled = 0;
    int rc = phdslam_step_filter(h, step_index, current_u, z, M, cfg.measurement_fields, &est);
    if (rc != 0 && rc != PHDSLAM_ERR_NAN) {
      fprintf(stderr, "phdslam: step %d failed (%d): %s\n", n, rc, phdslam_last_error());
      return 1;
    }
    /* state export where the reference looks at the particle set: after recoverSlamState, BEFORE resampleParticles
     * (main.cpp:1274-1279) -- the weighted particles, the map of the maximum-weight particle, and the resample indices
     * of the previous step.  The particle count changes from step to step when n_predict_particles > 1. */
    const int P = phdslam_n_local(h);
    poses.resize(P); logw.resize(P); ridx.resize(P);
    CHECK(phdslam_get_poses(h, poses.data()));
    CHECK(phdslam_get_log_weights(h, logw.data()));
    CHECK(phdslam_get_resample_idx(h, ridx.data()));
    if (cfg.filter_type == 1 && est.map_particle >= 0 && est.map_particle < P) {
      /* cardinality distribution of the maximum-weight particle (recoverSlamState, main.cpp:357-361) */
      all_card.resize((size_t)P * n_card);
      CHECK(phdslam_get_cardinalities(h, all_card.data()));
      std::copy(all_card.begin() + (size_t)est.map_particle * n_card, all_card.begin() + (size_t)(est.map_particle + 1) * n_card,
                card.begin());
    }
    int n_map = 0;
    if (cfg.map_estimate & 3) {
      int which = (cfg.map_estimate & 2) ? 2 : 1;
      int mrc = phdslam_map_estimate(h, which, map_est.data(), (int)map_est.size(), &n_map);
      if (mrc != 0) {
        fprintf(stderr, "phdslam: map estimate failed: %s\n", phdslam_last_error());
        n_map = 0;
      }
    }
    /* mixed feature model: the dynamic map of the maximum-weight particle goes on line 3 of the 7-line layout
     * (writeLog, main.cpp:885-900; run_synth passes particles.max_map_dynamic) */
    int n_dyn = 0;
    if (cfg.feature_model == 2 && (cfg.map_estimate & 1)) {
      dyn_est.resize((size_t)cfg.max_components_dynamic + 8);
      if (phdslam_map_estimate_dynamic(h, dyn_est.data(), (int)dyn_est.size(), &n_dyn) != 0) {
        fprintf(stderr, "phdslam: dynamic map estimate failed: %s\n", phdslam_last_error());
        n_dyn = 0;
      }
    }
    char name[64];
    snprintf(name, sizeof(name), "/state_estimate%05d.log", n);
    CHECK(phdslam_write_log_mixed((out_dir + name).c_str(), cfg.log_layout, &est.expected_pose, map_est.data(), n_map,
                                  dyn_est.data(), n_dyn, logw.data(), poses.data(), P, ridx.data(),
                                  cfg.filter_type == 1 ? card.data() : nullptr, n_card, cfg.filter_type));
    /* the nEff test and resampleParticles (main.cpp:1281-1297) */
    if (rc == 0) {
      int rrc = phdslam_step_resample(h, M, &est, &resampled);
      if (rrc != 0) {
        fprintf(stderr, "phdslam: resampling of step %d failed (%d): %s\n", n, rrc, phdslam_last_error());
        return 1;
      }
    }
    double el = now_ms() - t0;
    if (tf) fprintf(tf, "%g\n", el);
    if (!quiet)
      printf("pose %.4f %.4f %.4f  nEff %.4f%s  %.2f ms\n", est.expected_pose.px, est.expected_pose.py, est.expected_pose.ptheta,
             est.neff, resampled ? "  [resampled]" : "", el);
    if (rc == PHDSLAM_ERR_NAN) {
      printf("nan weights detected! exiting...\n");                /* main.cpp:1307-1311 */
      break;
    }
  }
  if (tf) fclose(tf);
  phdslam_timings_t t;
  phdslam_get_timings(h, &t);
  printf("done: %d steps, %llu kernel launches\n", n_steps, t.launches);
  phdslam_destroy(h);
  phdslam_free(zdata);
  phdslam_free(zoff);
  phdslam_free(udata);
  phdslam_free(ztimes);
  phdslam_free(ctimes);
  phdslam_free(traj);
  return 0;
}
