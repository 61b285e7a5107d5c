// planner_common.h -- shared pieces of the batched planner hosts (sff_planner.cpp, rrt_planner.h): the XML configuration
// in the reference schema (README.md:45-274, src/main.cpp:40-437), the double-precision geometry helpers that restate
// src/primitives.h, and small batching utilities over the C ABI of include/sffg.h.
#pragma once
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "sffg.h"

namespace planner {

constexpr double kTol = 1e-9;          // TOLERANCE, src/primitives.h:45
constexpr double kSample = 0.1;        // Solver::collisionSampleSize, src/problemStruct.h:121

[[noreturn]] inline void die(const std::string &msg) {
  std::cout << "Problem loading error: " << msg << "\n";
  std::exit(1);
}
inline void check(int rc) {
  if (rc != SFFG_OK) {
    std::cout << "engine error: " << sffg_last_error() << "\n";
    std::exit(3);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// configuration (same XML schema as the reference, README.md:45-274)
// ---------------------------------------------------------------------------------------------------------------
struct Tag {
  std::string name;
  std::map<std::string, std::string> attr;
  std::string parent;   // name of the enclosing element ("" at top level): the reference reads children of named nodes only
};

inline std::vector<Tag> scan_tags(const std::string &t) {
  std::vector<Tag> out;
  std::vector<std::string> open;   // element nesting
  size_t i = 0;
  while ((i = t.find('<', i)) != std::string::npos) {
    ++i;
    if (i >= t.size()) break;
    if (t[i] == '?') continue;
    if (t[i] == '/') {
      if (!open.empty()) open.pop_back();
      continue;
    }
    if (t.compare(i, 3, "!--") == 0) {
      size_t e = t.find("-->", i);
      i = e == std::string::npos ? t.size() : e + 3;
      continue;
    }
    Tag tag;
    while (i < t.size() && !std::isspace((unsigned char)t[i]) && t[i] != '>' && t[i] != '/') tag.name += t[i++];
    while (i < t.size() && t[i] != '>') {
      while (i < t.size() && (std::isspace((unsigned char)t[i]) || t[i] == '/')) ++i;
      if (i >= t.size() || t[i] == '>') break;
      std::string key;
      while (i < t.size() && t[i] != '=' && !std::isspace((unsigned char)t[i]) && t[i] != '>') key += t[i++];
      while (i < t.size() && t[i] != '"' && t[i] != '\'' && t[i] != '>') ++i;
      if (i >= t.size() || t[i] == '>') break;
      const char q = t[i++];
      std::string val;
      while (i < t.size() && t[i] != q) val += t[i++];
      ++i;
      tag.attr[key] = val;
    }
    tag.parent = open.empty() ? std::string() : open.back();
    const bool self_closing = i > 0 && i < t.size() && t[i] == '>' && t[i - 1] == '/';
    if (!self_closing) open.push_back(tag.name);
    out.push_back(tag);
  }
  return out;
}

inline bool parse_point(const std::string &s, double scale, double out[3]) {
  // "[x; y; z]"  (Point<T>(const std::string&, T scale), src/primitives.h:104-114)
  size_t a = s.find('['), b = s.find(']');
  if (a == std::string::npos || b == std::string::npos) return false;
  std::string body = s.substr(a + 1, b - a - 1);
  for (char &c : body)
    if (c == ';') c = ' ';
  std::istringstream is(body);
  for (int i = 0; i < 3; ++i) {
    if (!(is >> out[i])) return false;
    out[i] *= scale;
  }
  return true;
}

struct MeshRef {
  std::string file;
  int is_obj = 0;   // SFFG_MESH_TRI / SFFG_MESH_OBJ / SFFG_MESH_OBJ_FIXED (extension attribute fix_mesh="true")
  double pos[3] = {0, 0, 0};
};

struct Config {
  std::string solver = "sff";
  bool optimize = false;
  bool smoothing = false;
  int dim = 6;              // 2 or 6 (Dimensions, src/primitives.h:76-79)
  double scale = 1;
  MeshRef robot;
  std::vector<MeshRef> obstacles;
  bool has_map = true;
  std::vector<std::array<double, 3>> roots;
  bool has_goal = false;      // <Goal coord=.../>: single-goal planning (src/main.cpp:290-305)
  double goal[3] = {0, 0, 0};
  bool auto_range = false;
  double range[6] = {0, 0, 0, 0, 0, 0};
  double dtree = 0, circum = 1;
  double priority_bias = 0;
  int threshold_misses = 3;   // DEFAULT_THRES_MISS
  long max_iterations = 0;
  std::string params_file, id;
};

inline Config load_config(const std::string &path) {
  std::ifstream f(path);
  if (!f.good()) die("cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  std::vector<Tag> tags = scan_tags(ss.str());
  Config c;
  bool seen_problem = false, seen_range = false, seen_dist = false, seen_iter = false, seen_robot = false, in_save = false;
  for (const Tag &t : tags) {
    auto get = [&](const char *k) -> const std::string * {
      auto it = t.attr.find(k);
      return it == t.attr.end() ? nullptr : &it->second;
    };
    if (t.name == "Problem") {
      seen_problem = true;
      if (auto v = get("solver")) c.solver = *v; else die("invalid solver attribute in Problem node!");
      if (auto v = get("optimize")) c.optimize = *v == "true"; else die("invalid optimize attribute in Problem node!");
      if (auto v = get("smoothing")) c.smoothing = *v == "true";
      if (auto v = get("scale")) c.scale = std::stod(*v);
      if (auto v = get("dim")) {
        if (*v == "2D" || *v == "2d") c.dim = 2;
        else if (*v == "3D" || *v == "3d") c.dim = 6;
        else die("invalid dim attribute!");
      }
    } else if (t.name == "Robot") {
      seen_robot = true;
      if (auto v = get("file")) c.robot.file = *v; else die("invalid file node in Robot node!");
      if (auto v = get("is_obj")) c.robot.is_obj = *v == "true";
      if (auto v = get("fix_mesh")) if (*v == "true" && c.robot.is_obj) c.robot.is_obj = SFFG_MESH_OBJ_FIXED;
    } else if (t.name == "Obstacle" && t.parent == "Environment") {   // src/main.cpp:241-258: children of <Environment>
      MeshRef m;
      if (auto v = get("file")) m.file = *v; else die("invalid file attribute in Obstacle node!");
      if (auto v = get("is_obj")) m.is_obj = *v == "true";
      if (auto v = get("fix_mesh")) if (*v == "true" && m.is_obj) m.is_obj = SFFG_MESH_OBJ_FIXED;
      if (auto v = get("position")) if (!parse_point(*v, 1.0, m.pos)) die("Unknown format of point");
      c.obstacles.push_back(m);
    } else if (t.name == "Point" && t.parent == "Points") {           // src/main.cpp:166-186: children of <Points>
      std::array<double, 3> p;
      auto v = get("coord");
      if (!v || !parse_point(*v, c.scale, p.data())) die("invalid coord attribute in Point node!");
      c.roots.push_back(p);
    } else if (t.name == "Goal") {
      auto v = get("coord");
      if (!v || !parse_point(*v, c.scale, c.goal)) die("invalid coord attribute in Goal node!");
      c.has_goal = true;
    } else if (t.name == "Range") {
      seen_range = true;
      if (auto v = get("autoDetect")) c.auto_range = *v == "true";
    } else if (t.name == "RangeX" || t.name == "RangeY" || t.name == "RangeZ") {
      const int ax = t.name[5] - 'X';
      auto lo = get("min"), hi = get("max");
      if (!lo || !hi) die("invalid min/max attribute in range node");
      c.range[2 * ax] = c.scale * std::stod(*lo);
      c.range[2 * ax + 1] = c.scale * std::stod(*hi);
    } else if (t.name == "Distances") {
      seen_dist = true;
      auto a = get("dtree"), b = get("circum");
      if (!a || !b) die("invalid Distances node!");
      c.dtree = c.scale * std::stod(*a);
      c.circum = c.scale * std::stod(*b);
    } else if (t.name == "Improvements") {
      if (auto v = get("priorityBias")) c.priority_bias = std::stod(*v);
    } else if (t.name == "Thresholds") {
      if (auto v = get("standard")) c.threshold_misses = std::stoi(*v);
    } else if (t.name == "MaxIterations") {
      seen_iter = true;
      if (auto v = get("value")) c.max_iterations = std::stol(*v); else die("invalid MaxIterations node!");
    } else if (t.name == "Save") {
      in_save = true;
    } else if (t.name == "Params" && in_save && t.parent == "Save") {   // src/main.cpp:394-399: a child of <Save>
      if (auto v = get("file")) c.params_file = *v;
      if (auto v = get("id")) c.id = *v;
    }
  }
  if (!seen_problem) die("invalid root node!");
  if (!seen_robot) die("invalid Robot node!");
  if (!seen_range) die("invalid range node");
  if (!seen_dist) die("invalid Distances node!");
  if (!seen_iter) die("invalid MaxIterations node!");
  if (c.roots.empty()) die("invalid Points node - insert at least one point!");
  if (c.solver != "sff" && c.solver != "rrt" && c.solver != "lazy")
    die("unknown solver: the batched hosts cover \"sff\" (SFF / SFF*), \"rrt\" (RRT / RRT* / Multi-T-RRT) and \"lazy\" (Lazy-TSP)");
  if (c.solver == "rrt") {   // the reference's own validation, src/main.cpp:286-288, :327-329
    if (c.optimize && c.roots.size() > 1) die("Multi-T-RRT* is undefined!");
    if (!c.has_goal && c.priority_bias != 0) die("Multi-T-RRT with bias is undefined!");
  }
  if (c.solver == "lazy") {   // src/main.cpp:292-293, :330-331
    if (c.has_goal) die("single point path planning not defined for Lazy solver (use RRT/RRT* solver instead)!");
    if (c.priority_bias != 0) die("priority bias for Lazy solver is not implemented!");
  }
  c.has_map = !c.obstacles.empty();
  return c;
}

// ---------------------------------------------------------------------------------------------------------------
// geometry helpers (double, same operation order as the reference)
// ---------------------------------------------------------------------------------------------------------------
inline double wrap(double a) {   // NormalizeAngle, src/primitives.h:277-286
  if (a < -M_PI) return a + 2 * M_PI;
  if (a >= M_PI) return a - 2 * M_PI;
  return a;
}
inline double dist6(const double *a, const double *b) {   // Point<T>::distance, src/primitives.h:224-235
  double sum = 0;
  for (int i = 0; i < 3; ++i) {
    const double d = a[i] - b[i];
    sum += d * d;
  }
  for (int i = 3; i < 6; ++i) {
    const double d = wrap(b[i] - a[i]);
    sum += d * d;
  }
  return std::sqrt(sum);
}

// robot + obstacle meshes -> one engine environment (parseFile, src/main.cpp:161, :254); resolves Range autoDetect from the
// obstacles' bounding boxes (Environment::processLimits, src/environment.h:46-53)
inline sffg_env *load_environment(Config &cfg) {
  double *tris = nullptr;
  int64_t n = 0;
  double bbox[6];
  const double zero[3] = {0, 0, 0};
  check(sffg_mesh_load(cfg.robot.file.c_str(), cfg.robot.is_obj, zero, cfg.scale, &tris, &n, bbox));
  std::vector<double> robot(tris, tris + 9 * n);
  sffg_free(tris);
  std::vector<double> obst;
  double lim[6] = {1e308, -1e308, 1e308, -1e308, 1e308, -1e308};
  for (const MeshRef &m : cfg.obstacles) {
    check(sffg_mesh_load(m.file.c_str(), m.is_obj, m.pos, cfg.scale, &tris, &n, bbox));
    obst.insert(obst.end(), tris, tris + 9 * n);
    sffg_free(tris);
    for (int k = 0; k < 3; ++k) {
      lim[2 * k] = std::min(lim[2 * k], bbox[2 * k]);
      lim[2 * k + 1] = std::max(lim[2 * k + 1], bbox[2 * k + 1]);
    }
  }
  if (cfg.auto_range) {
    for (int k = 0; k < 6; ++k) cfg.range[k] = lim[k];
    cfg.auto_range = false;
  }
  sffg_env *env = nullptr;
  check(sffg_env_create(obst.empty() ? nullptr : obst.data(), (int64_t)(obst.size() / 9), robot.data(), (int64_t)(robot.size() / 9), &env));
  return env;
}

struct EdgeBatch {
  std::vector<double> s, e;
  std::vector<uint8_t> free_flag;
  int add(const double *a, const double *b) {
    s.insert(s.end(), a, a + 6);
    e.insert(e.end(), b, b + 6);
    return (int)(s.size() / 6) - 1;
  }
  void run(sffg_env *env) {
    const int64_t m = (int64_t)(s.size() / 6);
    free_flag.assign((size_t)m, 1);
    if (m) check(sffg_check_edges(env, s.data(), e.data(), m, kSample, SFFG_ROT_REFERENCE, free_flag.data(), nullptr));
  }
  void clear() {
    s.clear();
    e.clear();
    free_flag.clear();
  }
};

// ---------------------------------------------------------------------------------------------------------------
// plans between roots: DistanceHolder / neighboringMatrix of the reference (src/primitives.h:597-655,
// src/problemStruct.h:138) plus the solver-independent post-processing on top of it
// ---------------------------------------------------------------------------------------------------------------
struct Link {
  int n1 = -1, n2 = -1;
  double distance = std::numeric_limits<double>::max();
  std::vector<int> plan;   // global node ids
  bool exists() const { return n1 >= 0; }
};

class PlanBook {
 public:
  // node accessors supplied by the solver: position (6 doubles) and the id of the root tree a node hangs under
  std::function<const double *(int)> pos;
  std::function<int(int)> origin;
  std::map<std::pair<int, int>, Link> links;
  long smoothed = 0, verified_segments = 0;

  Link &link(int i, int j) { return links[{std::min(i, j), std::max(i, j)}]; }
  const Link &link(int i, int j) const {
    static const Link none;
    auto it = links.find({std::min(i, j), std::max(i, j)});
    return it == links.end() ? none : it->second;
  }
  double plan_length(const std::vector<int> &plan) const {   // Solver::computeDistance, src/problemStruct.h:170-181
    double d = 0;
    for (size_t i = 1; i < plan.size(); ++i) d += dist6(pos(plan[i - 1]), pos(plan[i]));
    return d;
  }

  // Solver::getAllPaths (src/problemStruct.h:183-253): plans between trees that only meet through a third tree
  void compose(const std::vector<int> &ct) {
    for (int id3 : ct)
      for (int id1 : ct) {
        if (id1 == id3 || !link(id1, id3).exists()) continue;
        for (int id2 : ct) {
          if (id1 == id2 || id2 == id3 || !link(id2, id3).exists()) continue;
          const Link &h1 = link(id1, id3), &h2 = link(id2, id3);
          std::vector<int> p1 = h1.plan, p2 = h2.plan;
          if (origin(p1.front()) != id1) std::reverse(p1.begin(), p1.end());
          if (origin(p2.front()) != id2) std::reverse(p2.begin(), p2.end());
          int last = -1;
          while (!p1.empty() && !p2.empty() && p1.back() == p2.back()) {   // drop the shared tail towards root id3
            last = p1.back();
            p1.pop_back();
            p2.pop_back();
          }
          if (last < 0) continue;
          std::vector<int> plan = p1;
          plan.push_back(last);
          plan.insert(plan.end(), p2.rbegin(), p2.rend());
          const double d = plan_length(plan);
          Link &direct = link(id1, id2);
          if (d < direct.distance - kTol) {
            direct.n1 = plan.front();
            direct.n2 = plan.back();
            direct.distance = d;
            direct.plan = plan;
          }
        }
      }
  }

  // smoothPaths (src/forest.h:465-511, src/rrt.h:353-379): walking back from the far end of a plan, connect the current
  // target to the EARLIEST node of the plan that sees it and drop everything in between.  All L(L-1)/2 candidate
  // shortcuts of a plan are evaluated in one batched edge call; the greedy choice is replayed on the host.
  void smooth(sffg_env *env, bool has_map, long &calls, long &n_edges) {
    for (auto &kv : links) {
      Link &l = kv.second;
      if (!l.exists() || l.plan.size() < 3) continue;
      const int L = (int)l.plan.size();
      EdgeBatch eb;
      std::vector<int> eid((size_t)L * L, -1);
      for (int g = 2; g < L; ++g)
        for (int t = 0; t + 1 < g; ++t) eid[(size_t)t * L + g] = eb.add(pos(l.plan[t]), pos(l.plan[g]));
      if (has_map) {
        eb.run(env);
        ++calls;
      } else {
        eb.free_flag.assign(eb.s.size() / 6, 1);
      }
      n_edges += (long)(eb.s.size() / 6);
      std::vector<int> keep;   // built from the back
      int g = L - 1;
      keep.push_back(l.plan[g]);
      while (g > 0) {
        int t = g - 1;
        for (int c = 0; c + 1 < g; ++c)
          if (eb.free_flag[eid[(size_t)c * L + g]]) {
            t = c;
            break;
          }
        keep.push_back(l.plan[t]);
        g = t;
      }
      std::reverse(keep.begin(), keep.end());
      l.plan = keep;
      l.distance = plan_length(keep);
      ++smoothed;
    }
  }

  // Solver::checkDistances (src/problemStruct.h:370-389) as an always-on verifier: every segment of every reported plan
  // must pass the local planner (either direction: tree edges were validated child->parent or parent->child)
  void verify(sffg_env *env, bool has_map, long &calls) {
    EdgeBatch fwd, bwd;
    for (const auto &kv : links) {
      const Link &l = kv.second;
      for (size_t i = 1; i < l.plan.size(); ++i) {
        fwd.add(pos(l.plan[i - 1]), pos(l.plan[i]));
        bwd.add(pos(l.plan[i]), pos(l.plan[i - 1]));
      }
    }
    if (fwd.s.empty() || !has_map) return;
    fwd.run(env);
    bwd.run(env);
    calls += 2;
    for (size_t i = 0; i < fwd.free_flag.size(); ++i)
      if (!fwd.free_flag[i] && !bwd.free_flag[i]) {
        std::cout << "Error: a segment of a reported path is not collision free\n";
        std::exit(1);
      }
    verified_segments = (long)fwd.free_flag.size();
  }

  // every plan as "treeA treeB length n  x y z yaw pitch roll ..." so that tests can re-validate each segment
  void save_paths(const std::string &file) const {
    std::ofstream out(file);
    out.precision(17);
    for (const auto &kv : links) {
      const Link &l = kv.second;
      if (!l.exists()) continue;
      out << kv.first.first << " " << kv.first.second << " " << l.distance << " " << l.plan.size();
      for (int id : l.plan)
        for (int c = 0; c < 6; ++c) out << " " << pos(id)[c];
      out << "\n";
    }
  }

  // one row of the Params file in the reference's format (Solver::saveParams, src/problemStruct.h:391-429)
  void save_params(const Config &cfg, const std::string &run_id, long iterations, bool solved, const std::vector<int> &connected,
                   double elapsed) const {
    if (cfg.params_file.empty()) return;
    std::ofstream out(cfg.params_file, std::ios_base::app);
    if (!out.good()) {
      std::cout << "Cannot create file at: " << cfg.params_file << "\n";
      return;
    }
    out << cfg.id << "," << run_id << "," << iterations << "," << (solved ? "solved" : "unsolved") << ",[";
    for (size_t i = 0; i < connected.size(); ++i) out << connected[i] << (i + 1 != connected.size() ? ";" : "");
    out << "],[";
    const int n = (int)connected.size();
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j) {
        out << link(connected[i], connected[j]).distance / cfg.scale;
        if (i + 1 != n || j + 1 != i) out << ";";
      }
    out << "]," << elapsed << "\n";
  }
};

struct StageClock {
  double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::chrono::steady_clock::time_point last;
  void start() { last = std::chrono::steady_clock::now(); }
  void lap(int i) {
    auto now = std::chrono::steady_clock::now();
    t[i] += std::chrono::duration<double>(now - last).count();
    last = now;
  }
};

}  // namespace planner
