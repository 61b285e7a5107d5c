// writers.h -- the reference's output files (README.md "Save" section; Solver<T,R>::saveCities / saveTrees / savePaths /
// saveTsp, src/problemStruct.h:263-341, :431-527), byte-compatible so that downstream tooling (the TSP solver fed by the
// TSP file, the visualisation scripts fed by the tree / path files) keeps working on top of the batched hosts.
// Positions are divided by the scale (angles are not, Point::operator/, src/primitives.h:216-222); numbers use the
// default ostream formatting like the reference.
#pragma once
#include "planner_common.h"

namespace planner {

struct OutFile {   // FileStruct, src/primitives.h:663-666
  std::string name;
  bool is_obj = false;
  bool set() const { return !name.empty(); }
};

struct SaveOptions {   // <Save> node, src/main.cpp:359-423
  OutFile goals, tree, raw_path, smooth_path, tsp, frontiers;
  long tree_every = 0, frontiers_every = 0;   // everyIteration attributes (src/main.cpp:372-376, :419-423)
};

// prefixFileName, src/primitives.h:698-709: the prefix goes behind the last '/' of the path
inline OutFile prefixed(const OutFile &f, const std::string &prefix) {
  OutFile r = f;
  const size_t pos = r.name.find_last_of('/');
  r.name.insert(pos == std::string::npos ? 0 : pos + 1, prefix);
  return r;
}

// getFile, src/main.cpp:439-465: "_<run>" goes in front of the extension when the run id is not 0
inline OutFile out_file(const Tag &t, const std::string &run_id) {
  OutFile f;
  auto it = t.attr.find("file");
  if (it == t.attr.end()) return f;
  f.name = it->second;
  bool nonzero = false;
  try {
    nonzero = std::stoi(run_id) != 0;
  } catch (...) {
    nonzero = false;
  }
  if (nonzero) {
    const size_t dot = f.name.find_last_of('.');
    if (dot != std::string::npos) f.name.insert(dot, "_" + std::to_string(std::stoi(run_id)));
  }
  auto o = t.attr.find("is_obj");
  f.is_obj = o != t.attr.end() && o->second == "true";
  return f;
}

inline SaveOptions load_save_options(const std::string &xml_path, const std::string &run_id, bool smoothing, bool is_sff) {
  std::ifstream f(xml_path);
  std::stringstream ss;
  ss << f.rdbuf();
  SaveOptions so;
  bool in_save = false;
  for (const Tag &t : scan_tags(ss.str())) {
    if (t.name == "Save") in_save = true;
    if (!in_save) continue;
    if (t.name == "Goals") so.goals = out_file(t, run_id);
    else if (t.name == "Tree") {
      so.tree = out_file(t, run_id);
      auto e = t.attr.find("everyIteration");
      if (so.tree.set() && e != t.attr.end()) so.tree_every = std::stol(e->second);
    } else if (t.name == "Frontiers") {
      so.frontiers = out_file(t, run_id);
      if (so.frontiers.set() && !is_sff) die("frontiers output is defined only for SFF-based solvers!");
      auto e = t.attr.find("everyIteration");
      if (so.frontiers.set() && e != t.attr.end()) so.frontiers_every = std::stol(e->second);
    }
    else if (t.name == "RawPath") so.raw_path = out_file(t, run_id);
    else if (t.name == "SmoothPath") so.smooth_path = out_file(t, run_id);
    else if (t.name == "TSP") so.tsp = out_file(t, run_id);
  }
  if (so.smooth_path.set() && !smoothing) die("smoothing is disabled, therefore \"SmoothPath\" parameter might not be defined!");
  return so;
}

// what the writers need to know about the solver's nodes
struct NodeView {
  int n_nodes = 0, n_trees = 0;
  double scale = 1;
  std::function<const double *(int)> pos;
  std::function<int(int)> parent;       // -1 for roots
  std::function<int(int)> tree;         // Node::Root id
  std::function<long(int)> age;         // iteration of creation
  std::function<bool(int)> is_root;     // DistanceToRoot == 0
};

inline bool open_out(std::ofstream &out, const OutFile &f) {
  out.open(f.name.c_str());
  if (!out.good()) {
    std::cout << "Cannot create file at: " << f.name << "\n";
    return false;
  }
  return true;
}
inline void put_pos(std::ostream &o, const double *p, double scale) { o << p[0] / scale << " " << p[1] / scale << " " << p[2] / scale; }
inline void put_point(std::ostream &o, const double *p, double scale) {   // operator<<(Point), src/primitives.h:273-275
  put_pos(o, p, scale);
  o << " " << p[3] << " " << p[4] << " " << p[5];
}

// Solver::saveCities, src/problemStruct.h:263-294 (called before the search: the trees hold their roots / the goal)
inline void save_goals(const OutFile &f, const NodeView &v, const std::vector<int> &root_nodes) {
  std::ofstream out;
  if (!f.set() || !open_out(out, f)) return;
  if (f.is_obj) out << "o Points\n";
  for (int id : root_nodes) {
    if (f.is_obj) out << "v ";
    put_point(out, v.pos(id), v.scale);
    out << "\n";
  }
}

// Solver::saveTrees, src/problemStruct.h:296-341
inline void save_trees(const OutFile &f, const NodeView &v) {
  std::ofstream out;
  if (!f.set() || !open_out(out, f)) return;
  if (f.is_obj) {
    out << "o Trees\n";
    for (int i = 0; i < v.n_nodes; ++i) {
      out << "v ";
      put_pos(out, v.pos(i), v.scale);
      out << "\n";
    }
    for (int t = 0; t < v.n_trees; ++t)
      for (int i = 0; i < v.n_nodes; ++i)
        if (v.tree(i) == t && !v.is_root(i) && v.parent(i) >= 0) out << "l " << i + 1 << " " << v.parent(i) + 1 << "\n";
  } else {
    out << "#X1 Y1 Z1 Yaw1 Pitch1 Roll1 X2 Y2 Z2 Yaw2 Pitch2 Roll2 TreeID IterationOfCreation\n";
    for (int t = 0; t < v.n_trees; ++t)
      for (int i = 0; i < v.n_nodes; ++i)
        if (v.tree(i) == t && !v.is_root(i) && v.parent(i) >= 0) {
          put_point(out, v.pos(i), v.scale);
          out << " ";
          put_point(out, v.pos(v.parent(i)), v.scale);
          out << " " << t << " " << v.age(i) << "\n";
        }
  }
}

// SpaceForest::saveFrontiers, src/forest.h:513-566: the open nodes (the priority variant prints positions only in OBJ files)
inline void save_frontiers(const OutFile &f, const NodeView &v, const std::vector<int> &open_nodes, bool priority) {
  std::ofstream out;
  if (!f.set() || !open_out(out, f)) return;
  if (f.is_obj) out << "o Open nodes\n";
  for (int id : open_nodes) {
    if (f.is_obj) {
      out << "v ";
      if (priority) put_pos(out, v.pos(id), v.scale);
      else put_point(out, v.pos(id), v.scale);
      out << "\n";
    } else {
      put_point(out, v.pos(id), v.scale);
      out << " 1\n";   // "the one in the end is just for plotting purposes"
    }
  }
}

// Solver::savePaths, src/problemStruct.h:470-527
inline void save_paths(const OutFile &f, const NodeView &v, const PlanBook &book) {
  std::ofstream out;
  if (!f.set() || !open_out(out, f)) return;
  if (f.is_obj) {
    out << "o Paths\n";
    for (int i = 0; i < v.n_nodes; ++i) {
      out << "v ";
      put_pos(out, v.pos(i), v.scale);
      out << "\n";
    }
  }
  for (int i = 0; i < v.n_trees; ++i)
    for (int j = i + 1; j < v.n_trees; ++j) {
      const Link &l = book.link(i, j);
      if (!l.exists()) continue;
      for (size_t k = 0; k + 1 < l.plan.size(); ++k) {
        if (f.is_obj) {
          out << "l " << l.plan[k] + 1 << " " << l.plan[k + 1] + 1 << "\n";
        } else {
          put_point(out, v.pos(l.plan[k]), v.scale);
          out << " ";
          put_point(out, v.pos(l.plan[k + 1]), v.scale);
          out << "\n";
        }
      }
      if (!f.is_obj) out << "\n";
    }
}

// Solver::saveTsp, src/problemStruct.h:431-468: TSPLIB explicit lower-diagonal matrix over the connected trees
inline void save_tsp(const OutFile &f, const Config &cfg, const PlanBook &book, const std::vector<int> &connected) {
  std::ofstream out;
  if (!f.set() || !open_out(out, f)) return;
  out << "NAME: " << cfg.id << "\n";
  out << "COMMENT: ";
  for (size_t i = 0; i < connected.size(); ++i) out << connected[i] << (i + 1 != connected.size() ? " " : "");
  out << "\n";
  out << "TYPE: TSP\n";
  out << "DIMENSION: " << connected.size() << "\n";
  out << "EDGE_WEIGHT_TYPE : EXPLICIT\n";
  out << "EDGE_WEIGHT_FORMAT : LOWER_DIAG_ROW\n";
  out << "EDGE_WEIGHT_SECTION\n";
  for (size_t i = 0; i < connected.size(); ++i) {
    for (size_t j = 0; j < i; ++j) out << book.link(connected[i], connected[j]).distance / cfg.scale << " ";
    out << "0\n";
  }
}

}  // namespace planner
