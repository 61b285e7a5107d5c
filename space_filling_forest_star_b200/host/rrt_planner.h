// rrt_planner.h -- restructured host of the reference's RRT / RRT* / Multi-T-RRT solver on top of the engine's C ABI.
//
// SURVEY.md section 8(f) row 1 names the RRT* loop next to SFF*: RapidExpTree<T,R>::Solve / expandNode (reference
// src/rrt.h:85-322) issue one nearest-neighbour query, one pose check and one edge check per iteration, then (RRT*) one
// k-NN query with up to 2k lazy edge checks and one 1-NN query + edge check per other tree.  Here the same rules run over
// *rounds* of up to B iterations:
//   1. B samples are drawn (tree choice src/rrt.h:94, goal bias :130-134, randomPointInSpace src/randGen.h:123-146)
//   2. one sffg_knn_multi(k = 1) finds every sample's nearest node in its tree                           (rrt.h:143)
//   3. steering (Point::getStateInDistance, src/primitives.h:237-250), then ONE pose call and ONE edge call     (:148-151)
//   4. RRT*: one sffg_knn_multi(k = 2e*log10(#nodes)) for the survivors                                     (:157-166)
//   5. one sffg_knn_multi(k = 1) of the survivors against every other tree                               (:219-229)
//   6. ONE edge call with every edge the parent-choice, rewire and tree-link rules could ask for  (:172, :185, :232)
//   7. the accept / choose-parent / rewire / link / merge rules are replayed on the host in sample order.
// A sample whose nearest node would have been a node created earlier in the same round is carried over to the next
// round with the same random point (the Voronoi bias of RRT is preserved exactly); k-NN and link queries see the
// tree as it was at the start of the round.  Deliberate differences to the reference: exact neighbours with the
// intended metric, all six coordinates are copied on a tree merge (the reference copies two,
// src/rrt.h:243-245), stepped angles are normalised into [-pi, pi), a link's length is the true length of its plan,
// the RNG is seedable.
#pragma once
#include <deque>

#include "planner_common.h"
#include "writers.h"

namespace planner {

class RrtPlanner {
 public:
  RrtPlanner(const Config &cfg, uint64_t seed, int batch, bool quiet) : cfg_(cfg), rng_(seed), batch_(batch), quiet_(quiet) {
    book_.pos = [this](int id) { return nodes_[id].p; };
    book_.origin = [this](int id) { return origin_of(id); };
  }

  // an environment owned by the caller (the Lazy-TSP host runs many searches in one map)
  void use_env(sffg_env *env) {
    env_ = env;
    own_env_ = false;
  }

  void load() {
    if (!env_) env_ = load_environment(cfg_);
    // one tree + one index per root (rrt.h:47-61); the goal is a tree of its own that is never expanded (:64-81)
    const int R = (int)cfg_.roots.size();
    for (int t = 0; t < R + (cfg_.has_goal ? 1 : 0); ++t) {
      Tree tr;
      check(sffg_index_create(cfg_.dim, &tr.idx));
      trees_.push_back(tr);
      RNode nd{};
      const double *src = t < R ? cfg_.roots[t].data() : cfg_.goal;
      for (int k = 0; k < 3; ++k) nd.p[k] = src[k];
      nd.tree = t;
      nd.parent = -1;
      const int id = add_node(nd);
      if (t == R) goal_node_ = id;
      frontier_.push_back(t);
    }
    flush_index_appends();
  }

  void solve() {
    {
      std::vector<int> roots;
      for (size_t i = 0; i < nodes_.size(); ++i)
        if (nodes_[i].parent < 0) roots.push_back((int)i);
      save_goals(save_.goals, view(), roots);   // rrt.h:86-88
    }
    const auto t0 = std::chrono::steady_clock::now();
    while (!solved_ && iter_ < cfg_.max_iterations) {
      const long iter_before = iter_;
      run_round();
      ++rounds_;
      if (save_.tree_every > 0) {   // Solver::saveIterCheck, src/problemStruct.h:255-261 (see sff_planner.cpp: dump_periodic)
        const long last = iter_ / save_.tree_every * save_.tree_every;
        if (last > iter_before && last > 0) save_trees(prefixed(save_.tree, "iter_" + std::to_string(last) + "_"), view());
      }
    }
    elapsed_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    connected_trees();
    save_trees(save_.tree, view());
    build_paths();
    save_paths(save_.raw_path, view(), book_);
    if (cfg_.smoothing) {
      book_.smooth(env_, cfg_.has_map, calls_, n_edges_);
      save_paths(save_.smooth_path, view(), book_);
    }
    book_.verify(env_, cfg_.has_map, calls_);
    save_tsp(save_.tsp, cfg_, book_, connected_);
  }

  void set_save(const SaveOptions &so) { save_ = so; }
  NodeView view() const {
    NodeView v;
    v.n_nodes = (int)nodes_.size();
    v.n_trees = (int)trees_.size();
    v.scale = cfg_.scale;
    v.pos = [this](int i) { return nodes_[i].p; };
    v.parent = [this](int i) { return nodes_[i].parent; };
    v.tree = [this](int i) { return origin_of(i); };
    v.age = [this](int i) { return nodes_[i].generation; };
    v.is_root = [this](int i) { return nodes_[i].parent < 0; };
    return v;
  }

  void save_params(const std::string &run_id) const { book_.save_params(cfg_, run_id, iter_, solved_, connected_, elapsed_); }
  void dump_plans(const std::string &file) const { book_.save_paths(file); }

  void report() const {
    if (quiet_) return;
    std::cout << "nodes " << nodes_.size() << ", iterations " << iter_ << ", rounds " << rounds_ << ", "
              << (solved_ ? "solved" : "unsolved") << ", connected trees " << connected_.size() << ", elapsed " << elapsed_
              << " s, engine calls " << calls_ << ", poses " << n_poses_ << ", edges " << n_edges_ << ", queries " << n_queries_
              << ", carried samples " << carried_ << ", rewires " << rewires_
              << "\nstage seconds: sample+nearest " << clk_.t[0] << ", pose+edge " << clk_.t[1] << ", k-NN " << clk_.t[2]
              << ", tree links " << clk_.t[3] << ", edges " << clk_.t[4] << ", replay " << clk_.t[5] << ", index append " << clk_.t[6]
              << "\nsmoothed plans " << book_.smoothed << ", verified segments " << book_.verified_segments << "\n";
  }

  ~RrtPlanner() {
    for (Tree &t : trees_)
      if (t.idx) sffg_index_destroy(t.idx);
    if (env_ && own_env_) sffg_env_destroy(env_);
  }

  bool solved() const { return solved_; }
  long iterations() const { return iter_; }
  long calls() const { return calls_; }
  const PlanBook &book() const { return book_; }
  const double *node_pos(int id) const { return nodes_[id].p; }

 private:
  struct RNode {
    double p[6];
    int tree;                // tree that currently owns the node (ExpandedRoot, src/primitives.h:450)
    int parent;              // global id, -1 for a root (Closest)
    double d_parent, d_root;
    std::vector<int> children;
    long generation = 0;
  };
  struct Tree {
    sffg_index *idx = nullptr;
    std::vector<int> members;                    // index id -> global node id (Tree::nodes order)
    std::vector<std::pair<int, int>> links;      // Tree::links, src/primitives.h:508
    std::vector<int> eaten;                      // Tree::eaten
    int eaten_by = -1;
  };
  struct Sample {
    int tree;
    double rnd[6];
  };
  struct Cand {
    Sample s;
    int nearest = -1;
    double p[6] = {0, 0, 0, 0, 0, 0};
    bool alive = false;
    double step = 0;
    int e_first = -1;
    std::vector<int> knn, e_parent, e_rewire;
    std::vector<int> link_tree, link_nb, e_link;   // per other frontier tree: neighbour (global id) and its edge or -1
  };

  int origin_of(int id) const {   // Node::Root: the root tree the node's parent chain ends in
    while (nodes_[id].parent >= 0) id = nodes_[id].parent;
    return root_tree_[id];
  }
  int goal_tree() const { return (int)cfg_.roots.size(); }
  int owner(int t) const {
    while (trees_[t].eaten_by >= 0) t = trees_[t].eaten_by;
    return t;
  }

  int add_node(const RNode &nd) {
    const int id = (int)nodes_.size();
    nodes_.push_back(nd);
    root_tree_.push_back(nd.parent < 0 ? nd.tree : -1);
    trees_[nd.tree].members.push_back(id);
    pending_.push_back(id);
    return id;
  }

  void flush_index_appends() {
    if (pending_.empty()) return;
    const int dim = cfg_.dim;
    std::vector<sffg_index *> idx;
    std::vector<int64_t> per;
    std::vector<float> rows;
    for (size_t t = 0; t < trees_.size(); ++t) {
      int64_t cnt = 0;
      for (int id : pending_)
        if (nodes_[id].tree == (int)t) {
          for (int c = 0; c < dim; ++c) rows.push_back((float)nodes_[id].p[c]);   // double -> float, rrt.h:206-209
          ++cnt;
        }
      if (cnt) {
        idx.push_back(trees_[t].idx);
        per.push_back(cnt);
      }
    }
    if (!idx.empty()) {
      check(sffg_index_add_multi(idx.data(), per.data(), (int)idx.size(), rows.data()));
      ++calls_;
    }
    pending_.clear();
  }

  // RandGen<T>::randomPointInSpace, src/randGen.h:123-146
  void random_point(double *out) {
    std::uniform_real_distribution<double> ux(cfg_.range[0], cfg_.range[1]), uy(cfg_.range[2], cfg_.range[3]),
        uz(cfg_.range[4], cfg_.range[5]), ang(-M_PI, M_PI), prob(0, 1);
    out[0] = ux(rng_);
    out[1] = uy(rng_);
    out[2] = out[3] = out[4] = out[5] = 0;
    if (cfg_.dim == 6) {
      out[2] = uz(rng_);
      out[3] = ang(rng_);
      double phi = std::acos(1 - 2 * prob(rng_)) + M_PI_2;
      if (prob(rng_) < 0.5) phi += phi < 0 ? M_PI : -M_PI;
      out[4] = phi;
      out[5] = ang(rng_);
    }
  }

  // Point<T>::getStateInDistance, src/primitives.h:237-250 (always steps exactly `dist`, also past the target)
  bool steer(const double *from, const double *to, double dist, double *out) const {
    const double real = dist6(from, to);
    if (!(real > 0)) return false;
    for (int i = 0; i < 3; ++i) out[i] = from[i] + (to[i] - from[i]) * (dist / real);
    for (int i = 3; i < 6; ++i) {
      const double a = from[i] + wrap(to[i] - from[i]) * (dist / real);
      out[i] = cfg_.dim == 6 ? a - 2 * M_PI * std::floor((a + M_PI) / (2 * M_PI)) : 0.0;   // same rotation, kept in [-pi, pi)
    }
    return true;
  }

  void knn_by_tree(const std::vector<std::vector<int>> &who, const std::vector<const double *> &qpos, int k,
                   std::vector<int32_t> &ids, std::vector<float> &d2, std::vector<size_t> &row_of) {
    // who[t] = query numbers asking tree t; rows are concatenated in tree order as sffg_knn_multi wants them
    const int dim = cfg_.dim, T = (int)trees_.size();
    std::vector<sffg_index *> idx;
    std::vector<int64_t> per;
    std::vector<float> q;
    size_t rows = 0;
    row_of.clear();
    for (int t = 0; t < T; ++t) {
      if (who[t].empty()) continue;
      idx.push_back(trees_[t].idx);
      per.push_back((int64_t)who[t].size());
      for (int qi : who[t]) {
        for (int c = 0; c < dim; ++c) q.push_back((float)qpos[qi][c]);
        row_of.push_back(rows++);
      }
    }
    ids.assign(rows * (size_t)k, -1);
    d2.assign(rows * (size_t)k, 0.f);
    if (!rows) return;
    check(sffg_knn_multi(idx.data(), per.data(), (int)idx.size(), q.data(), k, ids.data(), d2.data()));
    ++calls_;
    n_queries_ += (long)rows;
  }

  void run_round() {
    const int T = (int)trees_.size();
    clk_.start();
    // ---- 1. samples: carried ones first, then fresh draws.  Small trees get small rounds so that few samples have to be
    // carried (a sample is carried when a node created earlier in the round is nearer than its snapshot neighbour).
    // The goal tree is never expanded (rrt.h:62-64).  It has the highest id, so it is the last frontier entry until it is
    // eaten; the reference keeps a running "numTrees" that also shrinks when the goal is eaten, which stops expanding the
    // last root tree in the (untested, src/main.cpp:295) multi-root + goal case -- here every root tree stays expandable.
    const int n_exp = (int)frontier_.size() - (cfg_.has_goal && trees_[goal_tree()].eaten_by < 0 ? 1 : 0);
    size_t expandable = 0;
    for (int i = 0; i < n_exp; ++i) expandable += trees_[frontier_[i]].members.size();
    const long left = cfg_.max_iterations - iter_;
    const int B = (int)std::max<long>(1, std::min<long>({(long)batch_, (long)(expandable / 4), left}));
    std::vector<Cand> cand;
    while (!carry_.empty() && (int)cand.size() < B) {
      Cand c;
      c.s = carry_.front();
      c.s.tree = owner(c.s.tree);
      carry_.pop_front();
      cand.push_back(c);
    }
    while ((int)cand.size() < B) {
      Cand c;
      std::uniform_int_distribution<int> pick(0, std::max(0, n_exp - 1));
      c.s.tree = frontier_[pick(rng_)];
      if (cfg_.priority_bias != 0 && std::uniform_real_distribution<double>(0, 1)(rng_) <= cfg_.priority_bias)
        std::memcpy(c.s.rnd, nodes_[goal_node_].p, sizeof c.s.rnd);
      else
        random_point(c.s.rnd);
      cand.push_back(c);
    }
    // ---- 2. nearest node of the sample's tree
    std::vector<std::vector<int>> who(T);
    std::vector<const double *> qpos(cand.size());
    for (size_t i = 0; i < cand.size(); ++i) {
      who[cand[i].s.tree].push_back((int)i);
      qpos[i] = cand[i].s.rnd;
    }
    std::vector<int32_t> ids;
    std::vector<float> d2;
    std::vector<size_t> row_of;
    knn_by_tree(who, qpos, 1, ids, d2, row_of);
    {
      size_t r = 0;
      for (int t = 0; t < T; ++t)
        for (int qi : who[t]) cand[qi].nearest = trees_[t].members[ids[r++]];
    }
    clk_.lap(0);
    // ---- 3. steer, pose + first-edge verdicts
    std::vector<int> pose_of;
    EdgeBatch eb;
    for (size_t i = 0; i < cand.size(); ++i) {
      Cand &c = cand[i];
      if (!steer(nodes_[c.nearest].p, c.s.rnd, cfg_.circum, c.p)) continue;
      pose_of.push_back((int)i);
      c.e_first = eb.add(nodes_[c.nearest].p, c.p);
    }
    // one engine call validates every step: end pose free AND segment free (rrt.h:149, sffg_check_moves)
    std::vector<uint8_t> ok(pose_of.size(), 1);
    if (!pose_of.empty() && cfg_.has_map) {
      check(sffg_check_moves(env_, eb.s.data(), eb.e.data(), (int64_t)pose_of.size(), kSample, SFFG_ROT_REFERENCE, ok.data()));
      ++calls_;
    }
    n_poses_ += (long)pose_of.size();
    n_edges_ += (long)pose_of.size();
    std::vector<int> alive;
    for (size_t i = 0; i < pose_of.size(); ++i) {
      Cand &c = cand[pose_of[i]];
      c.alive = ok[i] != 0;
      if (c.alive) {
        c.step = dist6(nodes_[c.nearest].p, c.p);
        alive.push_back(pose_of[i]);
      }
    }
    clk_.lap(1);
    // ---- 4. RRT*: k nearest nodes of the own tree (k = 2e*log10(#nodes ever created), rrt.h:158)
    for (auto &w : who) w.clear();
    for (size_t i = 0; i < cand.size(); ++i) qpos[i] = cand[i].p;
    const int k = cfg_.optimize ? (int)std::min<double>(2 * M_E * std::log10((double)nodes_.size()), (double)SFFG_MAX_K) : 0;
    if (k >= 1 && !alive.empty()) {
      for (int ci : alive) who[cand[ci].s.tree].push_back(ci);
      knn_by_tree(who, qpos, k, ids, d2, row_of);
      size_t r = 0;
      for (int t = 0; t < T; ++t)
        for (int qi : who[t]) {
          for (int j = 0; j < k && ids[r * k + j] >= 0; ++j) cand[qi].knn.push_back(trees_[t].members[ids[r * k + j]]);
          ++r;
        }
    }
    clk_.lap(2);
    // ---- 5. nearest node of every other tree (rrt.h:219-229)
    if (frontier_.size() > 1 && !alive.empty()) {
      for (auto &w : who) w.clear();
      bool any = false;
      for (int t : frontier_) {
        if (trees_[t].members.size() == 1) continue;   // a tree of one node (the goal, a fresh root) needs no search
        for (int ci : alive)
          if (cand[ci].s.tree != t) {
            who[t].push_back(ci);
            any = true;
          }
      }
      if (any) knn_by_tree(who, qpos, 1, ids, d2, row_of);
      size_t r = 0;
      for (int t = 0; t < T; ++t) {
        if (trees_[t].eaten_by >= 0 || std::find(frontier_.begin(), frontier_.end(), t) == frontier_.end()) continue;
        if (trees_[t].members.size() == 1) {
          for (int ci : alive)
            if (cand[ci].s.tree != t) {
              cand[ci].link_tree.push_back(t);
              cand[ci].link_nb.push_back(trees_[t].members[0]);
            }
          continue;
        }
        for (int qi : who[t]) {
          cand[qi].link_tree.push_back(t);
          cand[qi].link_nb.push_back(trees_[t].members[ids[r++]]);
        }
      }
    }
    clk_.lap(3);
    // ---- 6. every edge the replay may ask for
    EdgeBatch eb2;
    for (int ci : alive) {
      Cand &c = cand[ci];
      const RNode &near = nodes_[c.nearest];
      const double best0 = c.step + near.d_root;
      double low = best0;
      c.e_parent.assign(c.knn.size(), -1);
      c.e_rewire.assign(c.knn.size(), -1);
      for (size_t j = 0; j < c.knn.size(); ++j) {
        const RNode &nb = nodes_[c.knn[j]];
        const double nd = dist6(c.p, nb.p) + nb.d_root;
        if (nd < best0 - kTol) {
          c.e_parent[j] = eb2.add(c.p, nb.p);
          low = std::min(low, nd);
        }
      }
      for (size_t j = 0; j < c.knn.size(); ++j) {
        const RNode &nb = nodes_[c.knn[j]];
        if (low + dist6(nb.p, c.p) < nb.d_root - kTol) c.e_rewire[j] = eb2.add(nb.p, c.p);
      }
      c.e_link.assign(c.link_nb.size(), -1);
      for (size_t j = 0; j < c.link_nb.size(); ++j)
        if (dist6(nodes_[c.link_nb[j]].p, c.p) < cfg_.dtree) c.e_link[j] = eb2.add(c.p, nodes_[c.link_nb[j]].p);
    }
    if (cfg_.has_map) {
      eb2.run(env_);
      if (!eb2.s.empty()) ++calls_;
    } else {
      eb2.free_flag.assign(eb2.s.size() / 6, 1);
    }
    n_edges_ += (long)(eb2.s.size() / 6);
    clk_.lap(4);
    // ---- 7. replay in sample order
    std::vector<int> added;
    size_t b = 0;
    bool merged = false;
    for (; b < cand.size() && iter_ < cfg_.max_iterations && !merged && !solved_; ++b) {
      Cand &c = cand[b];
      const int tree = c.s.tree;
      // sequential semantics: the nearest node includes the nodes created earlier in this round
      bool carried = false;
      const double dn = dist6(nodes_[c.nearest].p, c.s.rnd);
      for (int id : added) {
        if (nodes_[id].tree != tree) continue;
        // the translational part bounds the 6-D distance from below: most nodes are settled without wraps and sqrt
        const double dx = nodes_[id].p[0] - c.s.rnd[0], dy = nodes_[id].p[1] - c.s.rnd[1], dz = nodes_[id].p[2] - c.s.rnd[2];
        if (dx * dx + dy * dy + dz * dz >= dn * dn) continue;
        if (dist6(nodes_[id].p, c.s.rnd) < dn) {
          carried = true;
          break;
        }
      }
      if (carried) {
        carry_.push_back(c.s);
        ++carried_;
        continue;
      }
      ++iter_;
      if (!c.alive) continue;
      const int id = commit(c, eb2);
      added.push_back(id);
      // tree links and merges (rrt.h:219-317); the neighbour of every other tree was found on the round's snapshot
      int cur = tree;
      for (size_t j = 0; j < c.link_nb.size(); ++j) {
        const int other = c.link_tree[j];
        if (trees_[other].eaten_by >= 0 || other == cur || c.e_link[j] < 0 || !eb2.free_flag[c.e_link[j]]) continue;
        trees_[cur].links.emplace_back(id, c.link_nb[j]);
        cur = merge(cur, other);
        merged = true;
      }
    }
    // a merge changes tree membership: what was not replayed goes back to the queue (in order, before later draws)
    for (size_t r = cand.size(); r > b; --r) carry_.push_front(cand[r - 1].s);
    clk_.lap(5);
    flush_index_appends();
    clk_.lap(6);
  }

  // new node with RRT* parent choice and rewiring (rrt.h:153-201) or the plain RRT rule (:202-205)
  int commit(Cand &c, const EdgeBatch &eb2) {
    int parent = c.nearest;
    double best = c.step + nodes_[c.nearest].d_root;
    if (cfg_.optimize) {
      for (size_t j = 0; j < c.knn.size(); ++j) {
        const RNode &nb = nodes_[c.knn[j]];
        const double nd = dist6(c.p, nb.p) + nb.d_root;
        if (nd < best - kTol && c.e_parent[j] >= 0 && eb2.free_flag[c.e_parent[j]]) {
          best = nd;
          parent = c.knn[j];
        }
      }
    }
    RNode nd{};
    std::memcpy(nd.p, c.p, sizeof nd.p);
    nd.tree = c.s.tree;
    nd.parent = parent;
    if (cfg_.optimize) {
      nd.d_parent = dist6(nodes_[parent].p, c.p);
      nd.d_root = best;
    } else {   // plain RRT books the nominal step (rrt.h:203)
      nd.d_parent = cfg_.circum;
      nd.d_root = nodes_[parent].d_root + cfg_.circum;
    }
    nd.generation = iter_;
    const int id = add_node(nd);
    nodes_[parent].children.push_back(id);
    if (cfg_.optimize) {
      for (size_t j = 0; j < c.knn.size(); ++j) {
        RNode &nb = nodes_[c.knn[j]];
        const double d = dist6(nb.p, c.p);
        const double proposed = best + d;
        if (proposed < nb.d_root - kTol && c.e_rewire[j] >= 0 && eb2.free_flag[c.e_rewire[j]] && nb.parent >= 0 && c.knn[j] != parent) {
          std::vector<int> &ch = nodes_[nb.parent].children;
          auto it = std::find(ch.begin(), ch.end(), c.knn[j]);
          if (it == ch.end()) {
            std::cout << "Fatal error: Node not in children\n";
            std::exit(1);
          }
          ch.erase(it);
          nb.parent = id;
          nb.d_parent = d;
          nb.d_root = proposed;   // descendants keep their stored cost, as in the reference
          nodes_[id].children.push_back(c.knn[j]);
          ++rewires_;
        }
      }
    }
    return id;
  }

  // the tree with the higher id is eaten by the one with the lower id (rrt.h:237-315); returns the surviving tree
  int merge(int a, int b) {
    const int to = std::min(a, b), from = std::max(a, b);
    flush_index_appends();   // both indices must hold every member before the rows move
    const int dim = cfg_.dim;
    Tree &tt = trees_[to], &tf = trees_[from];
    std::vector<float> rows;
    for (int id : tf.members) {
      for (int c = 0; c < dim; ++c) rows.push_back((float)nodes_[id].p[c]);
      nodes_[id].tree = to;
      tt.members.push_back(id);
    }
    check(sffg_index_add(tt.idx, rows.data(), (int64_t)tf.members.size()));
    ++calls_;
    tt.links.insert(tt.links.end(), tf.links.begin(), tf.links.end());
    tt.eaten.push_back(from);
    tt.eaten.insert(tt.eaten.end(), tf.eaten.begin(), tf.eaten.end());
    tf.eaten_by = to;
    tf.members.clear();
    tf.links.clear();
    sffg_index_destroy(tf.idx);
    tf.idx = nullptr;
    auto it = std::find(frontier_.begin(), frontier_.end(), from);
    if (it == frontier_.end()) {
      std::cout << "Fatal error during tree merging (RRT)";
      std::exit(1);
    }
    frontier_.erase(it);
    solved_ = frontier_.size() == 1;
    return to;
  }

  // RapidExpTree::getConnectedTrees, rrt.h:381-393: the tree that has eaten most, plus what it has eaten
  void connected_trees() {
    size_t max_conn = 0;
    connected_.clear();
    central_ = -1;
    for (size_t t = 0; t < trees_.size(); ++t)
      if (trees_[t].eaten.size() > max_conn) {
        max_conn = trees_[t].eaten.size();
        central_ = (int)t;
        connected_ = trees_[t].eaten;
        connected_.push_back((int)t);
      }
  }

  // RapidExpTree::getPaths (rrt.h:324-351) + Solver::getAllPaths
  void build_paths() {
    if (central_ < 0) return;
    for (const auto &lk : trees_[central_].links) {
      Link l;
      for (int n = lk.first; n >= 0; n = nodes_[n].parent) l.plan.insert(l.plan.begin(), n);
      for (int n = lk.second; n >= 0; n = nodes_[n].parent) l.plan.push_back(n);
      l.n1 = l.plan.front();
      l.n2 = l.plan.back();
      l.distance = book_.plan_length(l.plan);
      const int r1 = root_tree_[l.n1], r2 = root_tree_[l.n2];
      if (r1 == r2) continue;
      Link &slot = book_.link(r1, r2);
      if (!slot.exists() || l.distance < slot.distance) slot = l;
    }
    book_.compose(connected_);
  }

  StageClock clk_;
  Config cfg_;
  SaveOptions save_;
  std::mt19937_64 rng_;
  int batch_;
  bool quiet_;
  sffg_env *env_ = nullptr;
  bool own_env_ = true;
  std::vector<RNode> nodes_;
  std::vector<int> root_tree_;          // for root nodes: their tree id, -1 otherwise
  std::vector<Tree> trees_;
  std::vector<int> frontier_;           // treeFrontier (tree ids), rrt.h:36
  int goal_node_ = -1, central_ = -1;
  std::vector<int> pending_;
  std::deque<Sample> carry_;
  PlanBook book_;
  std::vector<int> connected_;
  long iter_ = 0, rounds_ = 0, calls_ = 0, n_poses_ = 0, n_edges_ = 0, n_queries_ = 0, carried_ = 0, rewires_ = 0;
  bool solved_ = false;
  double elapsed_ = 0;
};

}  // namespace planner
