/*
 * host_io.cpp -- host-only part of libphdslam.so: configuration, text loaders, log writer.
 * No CUDA calls; usable (and tested) on a machine without a GPU.
 *
 * Mirrors: loadConfig (reference src/main.cpp:956-1073), loadMeasurements / parseMeasurements
 * (:192-244), loadControls (:169-190), writeLog (:848-954) and the README's 5-line log contract
 * (README:31-39).
 */
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/phd_detmath.h"
#include "../../include/phdslam.h"

static thread_local std::string g_last_error;
void phdslam_set_error(const std::string& s) { g_last_error = s; }

extern "C" const char* phdslam_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* phdslam_version(void) { return "phdslam-b200 0.1 (sm_100a)"; }

/* Defaults of loadConfig's options_description (src/main.cpp:961-1048). */
extern "C" void phdslam_config_defaults(phdslam_config_t* c) {
  memset(c, 0, sizeof(*c));
  c->motion_type = 1;
  c->ax = 0.5f; c->ay = 0.0f; c->ayaw = 0.0087f;
  c->dt = 0.1f;
  c->max_bearing = (float)M_PI;
  c->min_range = 0.0f;
  c->max_range = 20.0f;
  c->std_bearing = 0.0524f;
  c->std_range = 1.0f;
  c->clutter_rate = 15.0f;
  c->pd = 0.98f;
  c->n_particles = 512;
  c->n_predict_particles = 1;
  c->resample_threshold = 0.15f;
  c->subdivide_predict = 1;
  c->birth_weight = 0.05f;
  c->birth_noise_factor = 1.5f;
  c->min_separation = 5.0f;
  c->min_feature_weight = 0.00001f;
  c->particle_weighting = 1;
  c->max_cardinality = 256;
  c->filter_type = 1;
  c->map_estimate = 1;
  c->distance_metric = 0;
  c->feature_model = 0;
  c->max_steps = 10000;
  c->n_steps = -1;
  strcpy(c->data_directory, "data/");
  /* extensions */
  c->measurement_fields = 2;
  c->max_components = 256;
  c->resample_mode = 0;
  c->log_layout = 0;
  c->seed = 0;
  c->update_mode = 0;
  c->update_buffer_bytes = 32ull << 30;
  /* mixed feature model (main.cpp:990,1022-1025,1037-1038) */
  c->ps = 0.98f;
  c->tau = 0.0f;
  c->beta = 1.0f;
  c->max_components_dynamic = 64;
  c->clutter_density = c->clutter_rate / (2 * c->max_bearing * c->max_range);
}

static void derive(phdslam_config_t* c) {
  /* config.clutterDensity = config.clutterRate/( 2*config.maxBearing*config.maxRange ) (main.cpp:1065) */
  c->clutter_density = c->clutter_rate / (2 * c->max_bearing * c->max_range);
}

static bool parse_bool(const char* v, int* out) {
  std::string s(v);
  for (auto& ch : s) ch = (char)tolower(ch);
  if (s == "1" || s == "true" || s == "yes" || s == "on") { *out = 1; return true; }
  if (s == "0" || s == "false" || s == "no" || s == "off") { *out = 0; return true; }
  return false;
}

/* Keys the reference parses that are not on this path (disparity camera, constant-position feature noise, dead options).
 * Accepted and ignored so that the reference's cfg files load unchanged. */
static const char* kIgnoredKeys[] = {
    "debug", "initial_z", "initial_roll", "initial_pitch", "acc_z", "acc_roll", "acc_pitch", "gate_births",
    "gate_measurements", "gate_threshold", "min_expected_feature_weight", "max_features", "daughter_mixture_type",
    "n_samples", "cphd_disttype", "nu", "std_vx_features", "std_vy_features", "std_u", "std_v", "disparity_birth", "image_width", "image_height", "std_d_birth",
    "fx", "fy", "u0", "v0", "particles_per_feature", "save_all_maps", "save_prediction", nullptr};

extern "C" int phdslam_config_set(phdslam_config_t* c, const char* key, const char* value) {
  std::string k(key);
  const char* v = value;
#define F(name, field) if (k == name) { c->field = strtof(v, nullptr); derive(c); return 0; }
#define I(name, field) if (k == name) { c->field = (int)strtol(v, nullptr, 10); return 0; }
#define B(name, field) if (k == name) { int b; if (!parse_bool(v, &b)) return PHDSLAM_ERR_INVALID; c->field = b; return 0; }
  F("initial_x", x0) F("initial_y", y0) F("initial_yaw", yaw0) F("initial_vx", vx0)
  /* Reference quirk (main.cpp:969-972): initial_vy AND initial_vz both bind config.vy0, and boost notifies
   * in key order, so initial_vz (default 0) always wins; initial_vroll/vpitch/vyaw all bind vyaw0 and
   * initial_vyaw wins.  Mirrored: initial_vy is accepted but has no effect. */
  if (k == "initial_vy") {
    if (strtof(v, nullptr) != 0.0f)
      fprintf(stderr, "phdslam: note: initial_vy is overridden by initial_vz in the reference parser (main.cpp:969-970); ignored\n");
    return 0;
  }
  F("initial_vz", vy0)
  if (k == "initial_vroll" || k == "initial_vpitch") return 0;
  F("initial_vyaw", vyaw0)
  B("follow_trajectory", follow_trajectory)
  I("motion_type", motion_type)
  F("acc_x", ax) F("acc_y", ay) F("acc_yaw", ayaw) F("dt", dt)
  F("max_bearing", max_bearing) F("min_range", min_range) F("max_range", max_range)
  F("std_bearing", std_bearing) F("std_range", std_range) F("clutter_rate", clutter_rate) F("pd", pd)
  I("n_particles", n_particles) I("n_predict_particles", n_predict_particles)
  F("resample_threshold", resample_threshold) I("subdivide_predict", subdivide_predict)
  F("birth_weight", birth_weight) F("birth_noise_factor", birth_noise_factor)
  I("feature_model", feature_model)
  F("min_separation", min_separation) F("min_feature_weight", min_feature_weight)
  I("particle_weighting", particle_weighting) I("max_cardinality", max_cardinality) I("filter_type", filter_type)
  I("map_estimate", map_estimate) I("distance_metric", distance_metric)
  F("h", h) F("l", l) F("a", a) F("b", b) F("std_encoder", std_encoder) F("std_alpha", std_alpha)
  B("labeled_measurements", labeled_measurements)
  I("max_time_steps", max_steps) I("n_steps", n_steps)
  if (k == "data_directory") {
    snprintf(c->data_directory, sizeof(c->data_directory), "%s", v);
    return 0;
  }
  /* extensions */
  I("measurement_fields", measurement_fields) I("max_components", max_components) I("resample_mode", resample_mode)
  I("update_mode", update_mode)
  /* mixed feature model */
  F("ps", ps) F("tau", tau) F("beta", beta) F("std_ax_features", std_ax_features) F("std_ay_features", std_ay_features)
  F("cov_vx_birth", cov_vx_birth) F("cov_vy_birth", cov_vy_birth) I("max_components_dynamic", max_components_dynamic)
  if (k == "log_layout") {
    c->log_layout = (std::string(v) == "extended" || std::string(v) == "1") ? 1 : 0;
    return 0;
  }
  if (k == "seed") { c->seed = strtoull(v, nullptr, 10); return 0; }
  if (k == "update_buffer_bytes") { c->update_buffer_bytes = strtoull(v, nullptr, 10); return 0; }
#undef F
#undef I
#undef B
  for (int i = 0; kIgnoredKeys[i]; ++i)
    if (k == kIgnoredKeys[i]) return 0;
  return PHDSLAM_ERR_INVALID;
}

static std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

/* boost::program_options::parse_config_file grammar: `key = value`, `#` starts a comment (also inline),
 * blank lines ignored, `[section]` prefixes keys with "section.". */
extern "C" int phdslam_config_load(const char* path, phdslam_config_t* c) {
  std::ifstream ifs(path);
  if (!ifs) {
    phdslam_set_error(std::string("Unable to open config file: ") + path);
    return PHDSLAM_ERR_IO;
  }
  phdslam_config_defaults(c);
  std::string line, section;
  int lineno = 0;
  while (std::getline(ifs, line)) {
    ++lineno;
    size_t hash = line.find('#');
    if (hash != std::string::npos) line = line.substr(0, hash);
    line = trim(line);
    if (line.empty()) continue;
    if (line.front() == '[' && line.back() == ']') {
      section = trim(line.substr(1, line.size() - 2)) + ".";
      continue;
    }
    size_t eq = line.find('=');
    if (eq == std::string::npos) {
      fprintf(stderr, "phdslam: %s:%d: not a key=value line, skipped\n", path, lineno);
      continue;
    }
    std::string key = section + trim(line.substr(0, eq));
    std::string val = trim(line.substr(eq + 1));
    if (phdslam_config_set(c, key.c_str(), val.c_str()) != 0)
      fprintf(stderr, "phdslam: %s:%d: unknown or invalid option '%s' skipped (the reference aborts the parse here, main.cpp:1068-1071)\n",
              path, lineno, key.c_str());
  }
  derive(c);
  return 0;
}

/* ---- text loaders --------------------------------------------------------------------------- */

static void split_floats(const std::string& line, std::vector<float>& out) {
  const char* p = line.c_str();
  while (*p) {
    while (*p && (isspace((unsigned char)*p) || *p == ',')) ++p;
    if (!*p) break;
    char* end = nullptr;
    float v = strtof(p, &end);
    if (end == p) break;
    out.push_back(v);
    p = end;
  }
}

/* One line per time step; README:22-24 "Range1 Bearing1 Range2 Bearing2 ..." (fields = 2) or the
 * HEAD parser's "range bearing label" triples (fields = 3, main.cpp:198-203).  has_header: 1 = skip the
 * first line as loadMeasurements does (main.cpp:230), 0 = no header, -1 = skip it only if it is not numeric.
 * Unlike the reference, trailing blanks do not create a garbage measurement and a missing final newline
 * does not drop the last step. */
extern "C" int phdslam_load_measurements(const char* path, int fields, int has_header, float** data, int** offsets,
                                         int* n_steps) {
  std::ifstream f(path);
  if (!f) {
    phdslam_set_error(std::string("could not open measurements file: ") + path);
    return PHDSLAM_ERR_IO;
  }
  if (fields != 2 && fields != 3) return PHDSLAM_ERR_INVALID;
  std::vector<float> all;
  std::vector<int> off(1, 0);
  std::string line;
  bool first = true;
  std::vector<std::string> lines;
  while (std::getline(f, line)) lines.push_back(line);
  while (!lines.empty() && trim(lines.back()).empty()) lines.pop_back();
  for (const std::string& ln : lines) {
    if (first) {
      first = false;
      std::string t = trim(ln);
      bool numeric = !t.empty() && (isdigit((unsigned char)t[0]) || t[0] == '-' || t[0] == '+' || t[0] == '.');
      if (has_header == 1 || (has_header < 0 && !numeric)) continue;
    }
    std::vector<float> v;
    split_floats(ln, v);
    size_t nrec = v.size() / fields;
    all.insert(all.end(), v.begin(), v.begin() + nrec * fields);
    off.push_back((int)(all.size() / fields));
  }
  *n_steps = (int)off.size() - 1;
  *data = (float*)malloc(std::max<size_t>(all.size(), 1) * sizeof(float));
  memcpy(*data, all.data(), all.size() * sizeof(float));
  *offsets = (int*)malloc(off.size() * sizeof(int));
  memcpy(*offsets, off.data(), off.size() * sizeof(int));
  return 0;
}

/* loadControls (main.cpp:169-190): header skipped, "v_encoder alpha" per line. */
extern "C" int phdslam_load_controls(const char* path, float** data, int* n) {
  std::ifstream f(path);
  if (!f) {
    phdslam_set_error(std::string("could not open controls file: ") + path);
    return PHDSLAM_ERR_IO;
  }
  std::vector<float> all;
  std::string line;
  bool first = true;
  while (std::getline(f, line)) {
    std::string t = trim(line);
    if (first) {
      first = false;
      bool numeric = !t.empty() && (isdigit((unsigned char)t[0]) || t[0] == '-' || t[0] == '+' || t[0] == '.');
      if (!numeric) continue;
    }
    if (t.empty()) continue;
    std::vector<float> v;
    split_floats(t, v);
    all.push_back(v.size() > 0 ? v[0] : 0.0f);
    all.push_back(v.size() > 1 ? v[1] : 0.0f);
  }
  *n = (int)(all.size() / 2);
  *data = (float*)malloc(std::max<size_t>(all.size(), 1) * sizeof(float));
  memcpy(*data, all.data(), all.size() * sizeof(float));
  return 0;
}

/* loadTimestamps (main.cpp:147-167).  The reference pushes one value per getline() and pops the last (the empty line
 * after the final newline); blank lines are skipped here instead. */
extern "C" int phdslam_load_timestamps(const char* path, double** data, int* n) {
  *n = 0;
  *data = nullptr;
  std::ifstream f(path);
  if (!f) return 0;                       /* no time stamps: run_synth steps once per measurement set (:1091-1098) */
  std::vector<double> all;
  std::string line;
  while (std::getline(f, line)) {
    std::string t = trim(line);
    if (t.empty()) continue;
    all.push_back(strtod(t.c_str(), nullptr));
  }
  *n = (int)all.size();
  *data = (double*)malloc(std::max<size_t>(all.size(), 1) * sizeof(double));
  memcpy(*data, all.data(), all.size() * sizeof(double));
  return 0;
}

/* loadTrajectory (main.cpp:247-264) */
extern "C" int phdslam_load_trajectory(const char* path, phdslam_pose_t** data, int* n) {
  std::ifstream f(path);
  if (!f) {
    phdslam_set_error(std::string("could not open trajectory file: ") + path);
    return PHDSLAM_ERR_IO;
  }
  std::vector<phdslam_pose_t> all;
  std::string line;
  while (std::getline(f, line)) {
    std::string t = trim(line);
    if (t.empty() || t[0] == '%') continue;
    std::vector<float> v;
    split_floats(t, v);
    phdslam_pose_t s;
    float* o = &s.px;
    for (int k = 0; k < 6; ++k) o[k] = (k < (int)v.size()) ? v[k] : 0.0f;
    all.push_back(s);
  }
  *n = (int)all.size();
  *data = (phdslam_pose_t*)malloc(std::max<size_t>(all.size(), 1) * sizeof(phdslam_pose_t));
  memcpy(*data, all.data(), all.size() * sizeof(phdslam_pose_t));
  return 0;
}

/* the event schedule of run_synth (main.cpp:1187-1230); REAL is float in the reference (slamtypes.h) */
extern "C" int phdslam_plan_events(const double* zt, int nz, const double* ct, int nc, phdslam_event_t* ev, int cap) {
  if (nz < 0 || nc < 0 || (cap > 0 && !ev)) return PHDSLAM_ERR_INVALID;
  float current_time = 0.0f, last_time = 0.0f;     /* main.cpp:83-84 */
  int z_idx = 0, c_idx = 0, n = 0;
  const int n_steps = nz + nc;                      /* :1116 */
  for (; n < n_steps && n < cap; ++n) {
    if (z_idx >= nz || c_idx >= nc) break;          /* :1189-1192 */
    const float tz = (float)zt[z_idx], tc = (float)ct[c_idx];
    phdslam_event_t e;
    last_time = current_time;
    current_time = tc;
    e.dt = current_time - last_time;
    if (tz < tc) {
      e.z_idx = z_idx++;
      e.c_idx = -1;
    } else if (tz == tc) {
      e.c_idx = c_idx++;
      e.z_idx = z_idx++;
    } else {
      e.c_idx = c_idx++;
      e.z_idx = -1;
    }
    ev[n] = e;
  }
  return n;
}

extern "C" void phdslam_free(void* p) { free(p); }

/* ---- log writer ----------------------------------------------------------------------------- */

/* `stateFile << float` with the default ostream state == printf("%g") (6 significant digits). */
static void put(FILE* f, float v) { fprintf(f, "%g ", (double)v); }

extern "C" int phdslam_write_log(const char* path, int layout, const phdslam_pose_t* e, const phdslam_gaussian2d_t* map,
                                 int n_map, const float* log_weights, const phdslam_pose_t* poses, int n_particles,
                                 const int* resample_idx, const float* cardinality, int n_card, int filter_type) {
  return phdslam_write_log_mixed(path, layout, e, map, n_map, nullptr, 0, log_weights, poses, n_particles, resample_idx,
                                 cardinality, n_card, filter_type);
}

extern "C" int phdslam_write_log_mixed(const char* path, int layout, const phdslam_pose_t* e, const phdslam_gaussian2d_t* map,
                                       int n_map, const phdslam_gaussian4d_t* map_dynamic, int n_map_dynamic,
                                       const float* log_weights, const phdslam_pose_t* poses, int n_particles,
                                       const int* resample_idx, const float* cardinality, int n_card, int filter_type) {
  FILE* f = fopen(path, "w");
  if (!f) {
    phdslam_set_error(std::string("cannot write ") + path);
    return PHDSLAM_ERR_IO;
  }
  /* line 1: expected pose (main.cpp:861-864) */
  put(f, e->px); put(f, e->py); put(f, e->ptheta); put(f, e->vx); put(f, e->vy); put(f, e->vtheta);
  fputc('\n', f);
  /* line 2: map, 7 numbers per Gaussian: weight mean[2] cov[4] (main.cpp:867-882) */
  for (int n = 0; n < n_map; ++n) {
    put(f, map[n].weight);
    put(f, map[n].mean[0]); put(f, map[n].mean[1]);
    for (int i = 0; i < 4; ++i) put(f, map[n].cov[i]);
  }
  fputc('\n', f);
  if (layout == 1) { /* dynamic map line of the 7-line layout, 21 numbers per Gaussian: weight mean[4] cov[16] (main.cpp:885-900) */
    for (int n = 0; n < n_map_dynamic; ++n) {
      put(f, map_dynamic[n].weight);
      for (int i = 0; i < 4; ++i) put(f, map_dynamic[n].mean[i]);
      for (int i = 0; i < 16; ++i) put(f, map_dynamic[n].cov[i]);
    }
    fputc('\n', f);
  }
  /* particle log-weights (main.cpp:913-920) */
  for (int n = 0; n < n_particles; ++n) put(f, log_weights[n]);
  fputc('\n', f);
  /* particle poses (main.cpp:923-936) */
  for (int n = 0; n < n_particles; ++n) {
    put(f, poses[n].px); put(f, poses[n].py); put(f, poses[n].ptheta);
    put(f, poses[n].vx); put(f, poses[n].vy); put(f, poses[n].vtheta);
  }
  fputc('\n', f);
  if (layout == 1) { /* resample indices (main.cpp:939-943) */
    for (int n = 0; n < n_particles; ++n) fprintf(f, "%d ", resample_idx ? resample_idx[n] : n);
    fputc('\n', f);
  }
  /* cardinality distribution (main.cpp:946-953): "0 " repeated for PHD */
  for (int n = 0; n < n_card; ++n) {
    if (filter_type == 1 && cardinality) put(f, cardinality[n]);
    else fputs("0 ", f);
  }
  fputc('\n', f);
  fclose(f);
  return 0;
}

/* ---- global resampling: which rank holds the ancestor of offspring j ----------------------------------- */

/* Same arithmetic as resample_search_kernel / oracle_resample (canonical integer CDF). */
extern "C" unsigned long long phdslam_resample_threshold(int j, int n_new, unsigned long long total, const double* uniforms,
                                                         int resample_mode, unsigned call, unsigned long long seed) {
  const double interval = 1.0 / (double)n_new;
  double u;
  if (uniforms) {
    u = (resample_mode == 1) ? uniforms[0] : uniforms[1 + (size_t)j];
  } else {
    phd_philox4_t r = phd_philox4x32_10((resample_mode == 1) ? 0u : (uint32_t)j, call, PHD_STREAM_RESAMPLE, 0u,
                                        (uint32_t)seed, (uint32_t)(seed >> 32));
    u = phd_u01d(r.v[0], r.v[1]);
  }
  double rr = (double)j * interval + u * interval;
  double t = floor(rr * (double)total);
  unsigned long long R = (t <= 0.0) ? 0ull : (unsigned long long)t;
  if (R >= total) R = total - 1;
  return R;
}

extern "C" int phdslam_plan_migration(int world, const unsigned long long* totals, int n_new, const double* uniforms,
                                      int resample_mode, unsigned call, unsigned long long seed, int* bounds) {
  if (world < 1 || n_new < 1) return PHDSLAM_ERR_INVALID;
  unsigned long long total = 0;
  for (int r = 0; r < world; ++r) total += totals[r];
  if (total == 0) return PHDSLAM_ERR_NAN;
  unsigned long long base = 0;
  bounds[0] = 0;
  for (int r = 1; r < world; ++r) {
    base += totals[r - 1];
    /* first j with R_j >= base (R_j is non-decreasing in j) */
    int lo = 0, hi = n_new;
    while (lo < hi) {
      int mid = lo + (hi - lo) / 2;
      if (phdslam_resample_threshold(mid, n_new, total, uniforms, resample_mode, call, seed) >= base) hi = mid; else lo = mid + 1;
    }
    bounds[r] = lo;
  }
  bounds[world] = n_new;
  return 0;
}
