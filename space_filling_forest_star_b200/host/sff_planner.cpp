// sff_planner.cpp -- restructured host of the Space-Filling Forest / SFF* planner on top of the engine's C ABI.
//
// This is row (f)-1 of SURVEY.md section 8: the reference's expansion loop (SpaceForest<T,R>::Solve / expandNode,
// reference src/forest.h:112-376) evaluates one pose, one edge or one query at a time behind short-circuit conditions.
// Here the same decisions are taken over *rounds*: a round picks up to B frontier nodes, draws every node's
// ThresholdMisses candidate points at once and pushes all poses, edges and neighbour queries the sequential logic could
// possibly need through the GPU in five batched calls; the accept / reject / rewire rules are then replayed on the
// host over the returned flags in the reference's order.  Speculation changes the amount of work, never a rule:
//   candidate validity      limits, Environment::Collide(newPoint), isPathFree(expanded, newPoint)      (forest.h:244-248)
//   crowding / border rules radius search r = dtree + 2*circum over all trees, same-tree-closer-with-free-path rejects,
//                           other-tree-within-dtree records a border link and rejects                    (forest.h:255-303)
//   SFF* parent choice and rewiring over k = 2e*log10(#nodes) nearest nodes of the same tree               (forest.h:306-351)
//   frontier / closed list, ThresholdMisses, termination                                                   (forest.h:122-206)
//   path extraction: best border link per tree pair, composition through third trees                       (forest.h:420-463,
//                                                                                        problemStruct.h:183-253)
// Differences to the reference, all deliberate: neighbour search is exact with the intended metric (the reference's
// FLANN setup is approximate and its functor assigns instead of accumulating); expansions of one round
// do not see each other (a candidate that would interact with a node created in the same round is deferred to the
// next round); stored angles are normalised into [-pi, pi); the RNG is seedable.  This file covers solver="sff" (SFF / SFF*,
// plain or priority frontiers -- priorityBias, forest.h:79-89, :125-149 --, multi-goal or with a <Goal>, :91-109, :286-287);
// solver="rrt" (RRT / RRT* / Multi-T-RRT, with or without a goal) is in rrt_planner.h, solver="lazy" (Lazy-TSP) in lazy_planner.h.
//
//   sff_planner <config.xml> [run-id] [--seed S] [--batch B] [--paths file] [--quiet]
//
// Output: one line appended to the Params file in the reference's format (problemStruct.h:391-429):
//   id,run,iterations,solved|unsolved,[connected tree ids],[pairwise path lengths],elapsed seconds
#include <unordered_map>
#include <unordered_set>

#include "planner_common.h"
#include "writers.h"
#include "rrt_planner.h"
#include "lazy_planner.h"

using namespace planner;

namespace {

struct Node {
  double p[6];
  int tree;
  int parent;              // global id, -1 for a root
  double d_parent, d_root;
  std::vector<int> children;
  bool force_children = false;
  long generation = 0;
};

struct Border {
  int a, b;   // global node ids, a < b
};

// Heap<T,R> of the reference (src/heap.h:31-64) as the priority frontiers use it: a binary min-heap of node ids keyed by the
// distance to a fixed target (cost function Distance, src/primitives.h:530-533), with removal at an arbitrary heap
// position (pop(int), heap.h:203-227) and by node (the scan at forest.h:171-177).
class TargetHeap {
 public:
  bool empty() const { return h_.empty(); }
  int size() const { return (int)h_.size(); }
  bool contains(int id) const { return where_.count(id) != 0; }
  void append_ids(std::vector<int> &out) const {
    for (const auto &e : h_) out.push_back(e.second);
  }
  void push(int id, double key) {
    if (contains(id)) return;
    h_.push_back({key, id});
    where_[id] = size() - 1;
    up(size() - 1);
  }
  int pop_at(int pos) {
    const int id = h_[pos].second;
    where_.erase(id);
    if (pos != size() - 1) {
      h_[pos] = h_.back();
      where_[h_[pos].second] = pos;
      h_.pop_back();
      up(pos);
      down(pos);
    } else {
      h_.pop_back();
    }
    return id;
  }
  void remove(int id) {
    auto it = where_.find(id);
    if (it != where_.end()) pop_at(it->second);
  }

 private:
  void swap_at(int a, int b) {
    std::swap(h_[a], h_[b]);
    where_[h_[a].second] = a;
    where_[h_[b].second] = b;
  }
  void up(int i) {
    while (i > 0 && h_[i] < h_[(i - 1) / 2]) {
      swap_at(i, (i - 1) / 2);
      i = (i - 1) / 2;
    }
  }
  void down(int i) {
    for (;;) {
      int m = i;
      const int l = 2 * i + 1, r = 2 * i + 2;
      if (l < size() && h_[l] < h_[m]) m = l;
      if (r < size() && h_[r] < h_[m]) m = r;
      if (m == i) return;
      swap_at(i, m);
      i = m;
    }
  }
  std::vector<std::pair<double, int>> h_;
  std::unordered_map<int, int> where_;
};

struct Cand {
  int exp = -1;             // expanded node (global id)
  double p[6];
  bool in_limits = false, alive = false, rejected = false;
  double parent_dist = 0;
  int e_first = -1;         // edge expanded -> candidate
  // crowding / border stage
  std::vector<int> nb;                      // radius neighbours in (d2, id) order
  std::vector<int> nb_edge;                 // edge index per neighbour or -1
  int border_nb = -1;                       // neighbour that recorded a border link when this attempt is replayed
  bool reaches_goal = false;                // goal mode: the point sees the goal from within dtree (forest.h:286-287)
  // SFF* stage
  std::vector<int> knn;                     // same-tree neighbours, ascending distance
  std::vector<int> e_parent, e_rewire;      // edge index per knn entry or -1
};


class Planner {
 public:
  Planner(const Config &cfg, uint64_t seed, int batch, bool quiet) : cfg_(cfg), rng_(seed), batch_(batch), quiet_(quiet) {
    book_.pos = [this](int id) { return nodes_[id].p; };
    book_.origin = [this](int id) { return nodes_[id].tree; };
    n_trees_ = (int)cfg_.roots.size() + (cfg_.has_goal ? 1 : 0);   // Problem::GetNumRoots, src/problemStruct.h:74-80
    use_priority_ = cfg_.priority_bias != 0;                       // Solver::usePriority, src/problemStruct.h:150
  }

  void load() {
    env_ = load_environment(cfg_);
    const int T = n_trees_, R = (int)cfg_.roots.size();
    members_.resize(T);
    tree_idx_.resize(T, nullptr);
    heaps_.resize(T);
    check(sffg_index_create(cfg_.dim, &global_idx_));
    for (int t = 0; t < T; ++t) {   // one tree per root (forest.h:60-76); the goal is a tree of its own that is never expanded (:91-104)
      check(sffg_index_create(cfg_.dim, &tree_idx_[t]));
      Node nd{};
      for (int k = 0; k < 3; ++k) nd.p[k] = t < R ? cfg_.roots[t][k] : cfg_.goal[k];
      nd.tree = t;
      nd.parent = -1;
      nd.d_parent = nd.d_root = 0;
      const int id = add_node(nd);
      if (t < R) {
        if (!use_priority_) frontier_.push_back(id);
      } else {
        goal_node_ = id;
      }
    }
    if (use_priority_) {
      // priority frontiers (forest.h:79-89, :105-109): without a goal every tree keeps one heap per OTHER root, ordered by
      // the distance to that root; with a goal every tree keeps one heap ordered by the distance to the goal
      for (int t = 0; t < R; ++t) {
        for (int j = 0; j < R && !cfg_.has_goal; ++j)
          if (j != t) heap_targets_[t].push_back(members_[j][0]);
        if (cfg_.has_goal) heap_targets_[t].push_back(goal_node_);
        heaps_[t].resize(heap_targets_[t].size());
        push_to_heaps(members_[t][0]);
      }
    }
    flush_index_appends();
  }

  void solve() {
    {
      std::vector<int> roots;
      for (int t = 0; t < n_trees_; ++t) roots.push_back(members_[t][0]);
      save_goals(save_.goals, view(), roots);   // forest.h:113-116
    }
    const auto t0 = std::chrono::steady_clock::now();
    bool solved = false;
    const int T = n_trees_;
    while (!solved && iter_ < cfg_.max_iterations) {
      const bool from_closed = use_priority_ ? heaps_empty() : frontier_.empty();   // emptyFrontier, forest.h:183-193
      std::vector<int> chosen_nodes;
      std::vector<int> chosen;             // plain frontier / closed list: pool positions
      std::vector<TargetHeap *> prior;     // priority frontiers: the heap every node was popped from
      if (use_priority_ && !from_closed) {
        pick_by_priority(chosen_nodes, prior);
      } else {
        std::vector<int> &pool = from_closed ? closed_ : frontier_;
        if (pool.empty()) break;
        const int B = (int)std::min<size_t>((size_t)batch_, pool.size());
        // B distinct pool positions (partial Fisher-Yates over a position array)
        std::vector<int> pos(pool.size());
        for (size_t i = 0; i < pos.size(); ++i) pos[i] = (int)i;
        for (int i = 0; i < B; ++i) {
          std::uniform_int_distribution<int> pick(i, (int)pos.size() - 1);
          std::swap(pos[i], pos[pick(rng_)]);
        }
        chosen.assign(pos.begin(), pos.begin() + B);
        for (int i = 0; i < B; ++i) chosen_nodes.push_back(pool[chosen[i]]);
      }
      const int B = (int)chosen_nodes.size();
      if (B == 0) break;
      std::vector<char> exhausted(B, 0);
      const long iter_before = iter_;
      run_round(chosen_nodes, exhausted);
      if (!from_closed) {
        // nodes whose every attempt failed leave the frontier for the closed list (forest.h:160-180); in priority mode the
        // node was popped from one heap: it goes back there unless it is exhausted, in which case it leaves every heap
        std::vector<int> drop;
        for (int i = 0; i < B; ++i) {
          const int id = chosen_nodes[i];
          if (exhausted[i]) {
            nodes_[id].force_children = true;
            closed_.push_back(id);
            if (use_priority_) {
              for (TargetHeap &h : heaps_[nodes_[id].tree]) h.remove(id);
            } else {
              drop.push_back(chosen[i]);
            }
          } else if (use_priority_) {
            prior[i]->push(id, heap_key(id, prior[i]));
          }
        }
        std::sort(drop.begin(), drop.end(), std::greater<int>());
        for (int d : drop) {
          frontier_[d] = frontier_.back();
          frontier_.pop_back();
        }
      }
      dump_periodic(iter_before);
      if (goal_reached_) {
        solved = true;   // forest.h:287: the solve loop ends as soon as a new node sees the goal
      } else if (!cfg_.has_goal) {
        const bool empty = use_priority_ ? heaps_empty() : frontier_.empty();
        solved = empty && max_connected() == T;   // (!hasGoal && emptyFrontier && connected), forest.h:195-198
      }
      ++rounds_;
    }
    elapsed_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!solved && !cfg_.has_goal) solved = max_connected() == T;   // forest.h:204-206
    solved_ = solved;
    max_connected();
    save_trees(save_.tree, view());
    build_paths();
    save_paths(save_.raw_path, view(), book_);
    if (cfg_.smoothing) {
      book_.smooth(env_, cfg_.has_map, calls_, n_edges_);
      save_paths(save_.smooth_path, view(), book_);
    }
    book_.verify(env_, cfg_.has_map, calls_);
    save_tsp(save_.tsp, cfg_, book_, connected_);
    save_frontiers(save_.frontiers, view(), open_nodes(), use_priority_);   // forest.h:233-235
  }

  // nodes still open: the plain frontier, or (priority mode) the first heap of every tree -- every open node of a tree is
  // in all of its heaps (forest.h:527-538)
  std::vector<int> open_nodes() const {
    if (!use_priority_) return frontier_;
    std::vector<int> out;
    for (const auto &hs : heaps_)
      if (!hs.empty()) hs.front().append_ids(out);
    return out;
  }

  // Solver::saveIterCheck / SpaceForest::saveIterCheck (src/problemStruct.h:255-261, src/forest.h:570-578): a round advances
  // the iteration counter by many steps, so a dump is written at the end of the round that crossed a multiple of N,
  // labelled with that multiple
  void dump_periodic(long iter_before) {
    for (int which = 0; which < 2; ++which) {
      const long every = which ? save_.frontiers_every : save_.tree_every;
      if (every <= 0) continue;
      const long last = iter_ / every * every;
      if (last <= iter_before || last == 0) continue;
      const std::string prefix = "iter_" + std::to_string(last) + "_";
      if (which) save_frontiers(prefixed(save_.frontiers, prefix), view(), open_nodes(), use_priority_);
      else save_trees(prefixed(save_.tree, prefix), view());
    }
  }

  void set_save(const SaveOptions &so) { save_ = so; }
  NodeView view() const {
    NodeView v;
    v.n_nodes = (int)nodes_.size();
    v.n_trees = n_trees_;
    v.scale = cfg_.scale;
    v.pos = [this](int i) { return nodes_[i].p; };
    v.parent = [this](int i) { return nodes_[i].parent; };
    v.tree = [this](int i) { return nodes_[i].tree; };
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
              << "\nstage seconds: sample " << clk_.t[0] << ", pose+edge " << clk_.t[1] << ", radius " << clk_.t[2] << ", crowd edges "
              << clk_.t[3] << ", knn+rewire edges " << clk_.t[4] << ", replay " << clk_.t[5] << ", index append " << clk_.t[6]
              << "\nsmoothed plans " << book_.smoothed << ", verified segments " << book_.verified_segments << "\n";
  }

  ~Planner() {
    for (sffg_index *i : tree_idx_) sffg_index_destroy(i);
    sffg_index_destroy(global_idx_);
    sffg_env_destroy(env_);
  }

 private:
  // ---- priority frontiers (forest.h:125-135, :142-149)
  double heap_key(int id, const TargetHeap *h) const {
    const int t = nodes_[id].tree;
    const size_t j = (size_t)(h - heaps_[t].data());
    return dist6(nodes_[id].p, nodes_[heap_targets_.at(t)[j]].p);
  }
  void push_to_heaps(int id) {   // forest.h:360-363
    const int t = nodes_[id].tree;
    for (size_t j = 0; j < heaps_[t].size(); ++j) heaps_[t][j].push(id, dist6(nodes_[id].p, nodes_[heap_targets_[t][j]].p));
  }
  bool heaps_empty() const {
    for (const auto &hs : heaps_)
      for (const TargetHeap &h : hs)
        if (!h.empty()) return false;
    return true;
  }
  // up to B distinct nodes: random tree with a non-empty heap, random non-empty heap of it, then the best node of that heap
  // with probability priorityBias, else a random one
  void pick_by_priority(std::vector<int> &out, std::vector<TargetHeap *> &prior) {
    std::uniform_real_distribution<double> prob(0, 1);
    std::unordered_set<int> taken;
    // best-first search gains little from breadth: a round takes at most 4 nodes per live heap (a narrow beam), so that the
    // iteration budget is not spent on nodes the sequential search would never have reached
    int live_heaps = 0;
    for (const auto &hs : heaps_)
      for (const TargetHeap &h : hs) live_heaps += !h.empty();
    const int beam = std::min(batch_, 4 * std::max(1, live_heaps));
    for (int attempt = 0; attempt < 4 * beam && (int)out.size() < beam; ++attempt) {
      std::vector<int> live;
      for (int t = 0; t < n_trees_; ++t)
        for (const TargetHeap &h : heaps_[t])
          if (!h.empty()) {
            live.push_back(t);
            break;
          }
      if (live.empty()) break;
      const int t = live[std::uniform_int_distribution<int>(0, (int)live.size() - 1)(rng_)];
      std::vector<TargetHeap *> hs;
      for (TargetHeap &h : heaps_[t])
        if (!h.empty()) hs.push_back(&h);
      TargetHeap *h = hs[std::uniform_int_distribution<int>(0, (int)hs.size() - 1)(rng_)];
      const int pos = prob(rng_) <= cfg_.priority_bias ? 0 : std::uniform_int_distribution<int>(0, h->size() - 1)(rng_);
      const int id = h->pop_at(pos);
      if (taken.count(id)) {   // already picked through another heap of its tree in this round
        deferred_push_.push_back({h, id});
        continue;
      }
      taken.insert(id);
      out.push_back(id);
      prior.push_back(h);
    }
    for (auto &hp : deferred_push_) hp.first->push(hp.second, heap_key(hp.second, hp.first));
    deferred_push_.clear();
  }

  int add_node(const Node &nd) {
    const int id = (int)nodes_.size();
    nodes_.push_back(nd);
    members_[nd.tree].push_back(id);
    pending_.push_back(id);
    return id;
  }

  // new nodes become searchable at the end of a round: one append per tree index + one for the global index
  void flush_index_appends() {
    if (pending_.empty()) return;
    const int dim = cfg_.dim;
    // one engine call: the global index first, then every tree index that got nodes (sffg_index_add_multi)
    std::vector<sffg_index *> idx{global_idx_};
    std::vector<int64_t> per{(int64_t)pending_.size()};
    std::vector<float> base(pending_.size() * (size_t)dim);
    for (size_t i = 0; i < pending_.size(); ++i)
      for (int c = 0; c < dim; ++c) base[i * dim + c] = (float)nodes_[pending_[i]].p[c];   // double -> float, forest.h:258-260
    std::vector<float> rows(base);
    rows.reserve(2 * base.size());
    const int T = (int)tree_idx_.size();
    for (int t = 0; t < T; ++t) {
      int64_t cnt = 0;
      for (size_t i = 0; i < pending_.size(); ++i)
        if (nodes_[pending_[i]].tree == t) {
          rows.insert(rows.end(), base.begin() + i * dim, base.begin() + (i + 1) * dim);
          ++cnt;
        }
      if (cnt) {
        idx.push_back(tree_idx_[t]);
        per.push_back(cnt);
      }
    }
    // enqueued, not awaited: the appends are ordered before every later search on these indices by the engine, and the
    // host goes on to pick and sample the next round; the wait (free by then) sits in front of the next radius search
    check(sffg_index_add_multi_begin(idx.data(), per.data(), (int)idx.size(), rows.data()));
    ++calls_;
    pending_.clear();
  }

  // RandGen<T>::randomPointInDistance, src/randGen.h:69-109
  bool sample_around(const double *center, double *out) {
    std::uniform_real_distribution<double> ang(-M_PI, M_PI), prob(0, 1);
    const double distance = cfg_.circum;
    double tmp[6] = {0, 0, 0, 0, 0, 0};
    double phi = ang(rng_);
    if (cfg_.dim == 2) {
      out[0] = center[0] + std::cos(phi) * distance;
      out[1] = center[1] + std::sin(phi) * distance;
      out[2] = out[3] = out[4] = out[5] = 0;
    } else {
      const double theta = ang(rng_);
      tmp[0] = center[0] + std::cos(theta) * std::sin(phi) * distance;
      tmp[1] = center[1] + std::sin(theta) * std::sin(phi) * distance;
      tmp[2] = center[2] + std::cos(phi) * distance;
      tmp[3] = ang(rng_);
      phi = std::acos(1 - 2 * prob(rng_)) + M_PI_2;
      if (prob(rng_) < 0.5) phi += phi < 0 ? M_PI : -M_PI;
      tmp[4] = phi;
      tmp[5] = ang(rng_);
      // Point<T>::getStateInDistance, src/primitives.h:237-250 (angles are not re-normalised)
      const double real = dist6(center, tmp);
      for (int i = 0; i < 3; ++i) out[i] = center[i] + (tmp[i] - center[i]) * (distance / real);
      for (int i = 3; i < 6; ++i) {
        // the reference leaves the stepped angles unnormalised, so they random-walk away from [-pi, pi) and its
        // single-wrap metric degrades; here every stored angle is brought back into [-pi, pi) (same rotation)
        const double a = center[i] + wrap(tmp[i] - center[i]) * (distance / real);
        out[i] = a - 2 * M_PI * std::floor((a + M_PI) / (2 * M_PI));
      }
    }
    return out[0] >= cfg_.range[0] && out[0] <= cfg_.range[1] && out[1] >= cfg_.range[2] && out[1] <= cfg_.range[3] &&
           out[2] >= cfg_.range[4] && out[2] <= cfg_.range[5];
  }

  static constexpr size_t kEarlyKnnRows = 192;   // rounds with at most this many valid candidates search before the crowding rule
  // one sffg_knn_multi over all trees, in flight between start_knn and finish_knn: queries concatenated in tree order
  struct KnnAsync {
    int k = 0;
    std::vector<int> who;
    std::vector<float> q;
    std::vector<int32_t> ids;
    std::vector<float> d2;
    bool pending = false;
  };
  void start_knn(const std::vector<Cand> &cand, const std::vector<int> &alive, KnnAsync &ka) {
    const int dim = cfg_.dim, T = (int)tree_idx_.size();
    ka.k = (int)std::min<double>(2 * M_E * std::log10((double)nodes_.size()), (double)SFFG_MAX_K);   // forest.h:309
    if (ka.k < 1 || alive.empty()) return;
    std::vector<int64_t> per(T, 0);
    for (int t = 0; t < T; ++t)
      for (int ci : alive)
        if (nodes_[cand[ci].exp].tree == t) {
          ka.who.push_back(ci);
          ++per[t];
        }
    ka.q.resize(ka.who.size() * (size_t)dim);
    for (size_t i = 0; i < ka.who.size(); ++i)
      for (int c = 0; c < dim; ++c) ka.q[i * dim + c] = (float)cand[ka.who[i]].p[c];
    ka.ids.resize(ka.who.size() * (size_t)ka.k);
    ka.d2.resize(ka.ids.size());
    check(sffg_knn_multi_begin(tree_idx_.data(), per.data(), T, ka.q.data(), ka.k, ka.ids.data(), ka.d2.data()));
    ka.pending = true;
    ++calls_;
    n_queries_ += (long)ka.who.size();
  }
  void finish_knn(std::vector<Cand> &cand, KnnAsync &ka) {
    if (!ka.pending) return;
    check(sffg_index_end(tree_idx_[0]));
    ka.pending = false;
    const int k = ka.k;
    for (size_t i = 0; i < ka.who.size(); ++i) {
      Cand &c = cand[ka.who[i]];
      const int t = nodes_[c.exp].tree;
      for (int j = 0; j < k && ka.ids[i * k + j] >= 0; ++j) c.knn.push_back(members_[t][ka.ids[i * k + j]]);
    }
  }

  void run_round(const std::vector<int> &chosen, std::vector<char> &exhausted) {
    const int B = (int)chosen.size(), A = cfg_.threshold_misses, dim = cfg_.dim;
    std::vector<Cand> cand((size_t)B * A);
    clk_.start();
    // ---- stage 1: candidate points, pose + first-edge verdicts
    std::vector<int> pose_of;
    EdgeBatch eb;
    for (int b = 0; b < B; ++b)
      for (int a = 0; a < A; ++a) {
        Cand &c = cand[(size_t)b * A + a];
        c.exp = chosen[b];
        c.in_limits = sample_around(nodes_[c.exp].p, c.p);
        if (!c.in_limits) continue;
        pose_of.push_back(b * A + a);
        c.e_first = eb.add(nodes_[c.exp].p, c.p);
      }
    clk_.lap(0);
    // one engine call validates every candidate: end pose free AND segment free (forest.h:246, sffg_check_moves)
    std::vector<uint8_t> ok(pose_of.size(), 1);
    if (!pose_of.empty() && cfg_.has_map) {
      check(sffg_check_moves(env_, eb.s.data(), eb.e.data(), (int64_t)pose_of.size(), kSample, SFFG_ROT_REFERENCE, ok.data()));
      ++calls_;
      n_poses_ += (long)pose_of.size();
      n_edges_ += (long)pose_of.size();
    }
    std::vector<int> alive;
    for (size_t i = 0; i < pose_of.size(); ++i) {
      Cand &c = cand[pose_of[i]];
      c.alive = ok[i] != 0;
      if (c.alive) {
        c.parent_dist = dist6(nodes_[c.exp].p, c.p);
        alive.push_back(pose_of[i]);
      }
    }
    clk_.lap(1);
    // ---- stage 4 (SFF*), started early in small rounds: k nearest nodes of the same tree for EVERY valid candidate.  Only the first
    // surviving attempt of a node will use its row (forest.h:306-351), but which attempt survives is known only after
    // the crowding rule; the searches do not depend on it, so they are enqueued now (sffg_knn_multi_begin) and run on the
    // tree indices' streams while the radius search and the crowding edges of this round are answered.
    // (Latency-bound rounds only: with many candidates per round the extra rows cost more than the hidden call saves --
    // measured on the 6-tree scenes --, and the search is started for the winners alone once they are known.)
    KnnAsync ka;
    const bool early_knn = cfg_.optimize && alive.size() <= kEarlyKnnRows;
    if (early_knn) start_knn(cand, alive, ka);
    // ---- stage 2: radius search over every tree (one global index == union of the per-tree searches).
    // The reference asks for everything within dtree + 2*circum (forest.h:261-267) but its rules only ever fire for
    // neighbours closer than max(parentDistance, dtree) (forest.h:276, :283); with an exact search the smaller radius
    // returns exactly the neighbours that can matter (plus a float-rounding margin; the rules re-test in double).
    if (!alive.empty()) {
      std::vector<float> q(alive.size() * (size_t)dim);
      double reach = cfg_.dtree;
      for (size_t i = 0; i < alive.size(); ++i) {
        for (int k = 0; k < dim; ++k) q[i * dim + k] = (float)cand[alive[i]].p[k];
        reach = std::max(reach, cand[alive[i]].parent_dist);
      }
      reach = reach * (1.0 + 1e-5) + 1e-4 * (1.0 + std::fabs(cfg_.range[1]) + std::fabs(cfg_.range[3]) + std::fabs(cfg_.range[5]));
      const float r2 = (float)(reach * reach);
      std::vector<int32_t> counts(alive.size());
      int64_t total = 0;
      radius_ids_.resize(std::max<size_t>(radius_ids_.size(), alive.size() * 32));
      radius_d2_.resize(radius_ids_.size());
      check(sffg_index_end(global_idx_));   // the appends of the previous round (enqueued by flush_index_appends)
      int rc = sffg_radius(global_idx_, q.data(), (int64_t)alive.size(), r2, counts.data(), radius_ids_.data(), radius_d2_.data(),
                           (int64_t)radius_ids_.size(), &total);
      ++calls_;
      if (rc == SFFG_ERR_CAPACITY) {   // optimistic buffer too small: grow to the reported size and repeat once
        radius_ids_.resize((size_t)total * 2);
        radius_d2_.resize(radius_ids_.size());
        rc = sffg_radius(global_idx_, q.data(), (int64_t)alive.size(), r2, counts.data(), radius_ids_.data(), radius_d2_.data(),
                         (int64_t)radius_ids_.size(), &total);
        ++calls_;
      }
      check(rc);
      n_queries_ += (long)alive.size();
      size_t off = 0;
      for (size_t i = 0; i < alive.size(); ++i) {
        Cand &c = cand[alive[i]];
        c.nb.assign(radius_ids_.begin() + off, radius_ids_.begin() + off + counts[i]);
        off += (size_t)counts[i];
        // the reference searches tree after tree (forest.h:262): tree-major order, (d2, id) order inside a tree
        std::stable_sort(c.nb.begin(), c.nb.end(), [&](int x, int y) { return nodes_[x].tree < nodes_[y].tree; });
      }
    }
    clk_.lap(2);
    // ---- stage 3: edges the crowding / border rules may ask for (forest.h:270-302), in neighbour order
    eb.clear();
    for (int ci : alive) {
      Cand &c = cand[ci];
      const Node &ex = nodes_[c.exp];
      c.nb_edge.assign(c.nb.size(), -1);
      for (size_t j = 0; j < c.nb.size(); ++j) {
        const Node &nb = nodes_[c.nb[j]];
        const double real = dist6(nb.p, c.p);
        if (!ex.force_children && real < c.parent_dist - kTol && nb.tree == ex.tree) c.nb_edge[j] = eb.add(nb.p, c.p);
        if (nb.tree != ex.tree && real < cfg_.dtree - kTol) {
          if (!cfg_.has_goal) c.nb_edge[j] = eb.add(ex.p, nb.p);                       // forest.h:288
          else if (c.nb[j] == goal_node_) c.nb_edge[j] = eb.add(c.p, nb.p);             // forest.h:286-287
          else c.nb_edge[j] = -2;                                                       // goal mode: rejected untested
          c.nb.resize(j + 1);       // this neighbour always ends the scan
          c.nb_edge.resize(j + 1);
          break;
        }
      }
    }
    if (cfg_.has_map) {
      eb.run(env_);
      if (!eb.s.empty()) ++calls_;
    } else {
      eb.free_flag.assign(eb.s.size() / 6, 1);
    }
    n_edges_ += (long)(eb.s.size() / 6);
    for (int ci : alive) {
      Cand &c = cand[ci];
      const Node &ex = nodes_[c.exp];
      for (size_t j = 0; j < c.nb.size() && !c.rejected && !c.reaches_goal; ++j) {
        if (c.nb_edge[j] == -1) continue;
        const Node &nb = nodes_[c.nb[j]];
        if (nb.tree == ex.tree) {
          if (eb.free_flag[c.nb_edge[j]]) c.rejected = true;             // a closer node of the same tree sees the point
        } else if (cfg_.has_goal) {
          // near another tree: only a free view of the goal keeps the point (it becomes the node that ends the search)
          if (c.nb_edge[j] >= 0 && eb.free_flag[c.nb_edge[j]]) c.reaches_goal = true;
          else c.rejected = true;
        } else {
          if (eb.free_flag[c.nb_edge[j]]) c.border_nb = c.nb[j];          // trees meet: remember the link
          c.rejected = true;
        }
      }
    }
    clk_.lap(3);
    // ---- stage 4 + 5 (SFF*): k nearest nodes of the same tree for the first surviving attempt of every node
    std::vector<int> winners(B, -1);
    for (int b = 0; b < B; ++b)
      for (int a = 0; a < A; ++a) {
        const Cand &c = cand[(size_t)b * A + a];
        if (c.alive && !c.rejected) {
          winners[b] = b * A + a;
          break;
        }
      }
    EdgeBatch eb2;
    if (cfg_.optimize) {
      if (!early_knn) {
        std::vector<int> won;
        for (int b = 0; b < B; ++b)
          if (winners[b] >= 0) won.push_back(winners[b]);
        start_knn(cand, won, ka);
      }
      finish_knn(cand, ka);
      for (int b = 0; b < B; ++b) {
        if (winners[b] < 0) continue;
        Cand &c = cand[winners[b]];
        const Node &ex = nodes_[c.exp];
        const double best0 = c.parent_dist + ex.d_root;
        double low = best0;
        c.e_parent.assign(c.knn.size(), -1);
        c.e_rewire.assign(c.knn.size(), -1);
        for (size_t j = 0; j < c.knn.size(); ++j) {
          const Node &nb = nodes_[c.knn[j]];
          const double nd = dist6(c.p, nb.p) + nb.d_root;
          if (nd < best0 - kTol) {
            c.e_parent[j] = eb2.add(c.p, nb.p);
            low = std::min(low, nd);
          }
        }
        for (size_t j = 0; j < c.knn.size(); ++j) {
          const Node &nb = nodes_[c.knn[j]];
          if (low + dist6(nb.p, c.p) < nb.d_root - kTol) c.e_rewire[j] = eb2.add(nb.p, c.p);
        }
      }
      if (cfg_.has_map) {
        eb2.run(env_);
        if (!eb2.s.empty()) ++calls_;
      } else {
        eb2.free_flag.assign(eb2.s.size() / 6, 1);
      }
      n_edges_ += (long)(eb2.s.size() / 6);
    }
    clk_.lap(4);
    // ---- stage 6: replay in the reference's order
    std::vector<int> added;
    for (int b = 0; b < B && iter_ < cfg_.max_iterations; ++b) {
      bool success = false, deferred = false;
      for (int a = 0; a < A && !success && !deferred && iter_ < cfg_.max_iterations; ++a) {
        Cand &c = cand[(size_t)b * A + a];
        if (c.alive && !c.rejected) {
          // interaction with a node created earlier in this round was not evaluated: postpone this frontier node
          const Node &ex = nodes_[c.exp];
          for (int id : added) {
            const Node &o = nodes_[id];
            const double thr = o.tree == ex.tree ? c.parent_dist - kTol : cfg_.dtree - kTol;
            // the translational part is a lower bound of the 6-D distance: most pairs are settled without wraps and sqrt
            const double dx = o.p[0] - c.p[0], dy = o.p[1] - c.p[1], dz = o.p[2] - c.p[2];
            if (dx * dx + dy * dy + dz * dz >= thr * thr) continue;
            if (dist6(o.p, c.p) < thr) {
              deferred = true;
              break;
            }
          }
          if (deferred) break;
        }
        ++iter_;
        if (!c.alive) continue;
        if (c.rejected) {
          if (c.border_nb >= 0) add_border(c.border_nb, c.exp);
          continue;
        }
        added.push_back(commit(c, eb2));
        success = true;
        if (c.reaches_goal) {   // forest.h:369-372: link the new node to the goal; the solve loop ends
          add_border(added.back(), goal_node_);
          goal_reached_ = true;
        }
      }
      if (goal_reached_) {
        for (int r = b; r < B; ++r) exhausted[r] = 0;
        break;
      }
      exhausted[b] = !success && !deferred;
    }
    clk_.lap(5);
    flush_index_appends();
    clk_.lap(6);
  }

  // creates the node of an accepted candidate: SFF* parent choice + rewiring (forest.h:306-351) or plain SFF (:352-356)
  int commit(Cand &c, const EdgeBatch &eb2) {
    int parent = c.exp;
    double best = c.parent_dist + nodes_[c.exp].d_root;
    if (cfg_.optimize) {
      for (size_t j = 0; j < c.knn.size(); ++j) {
        const Node &nb = nodes_[c.knn[j]];
        const double nd = dist6(c.p, nb.p) + nb.d_root;
        if (nd < best - kTol && c.e_parent[j] >= 0 && eb2.free_flag[c.e_parent[j]]) {
          best = nd;
          parent = c.knn[j];
        }
      }
    }
    Node nd{};
    std::memcpy(nd.p, c.p, sizeof nd.p);
    nd.tree = nodes_[c.exp].tree;
    nd.parent = parent;
    nd.d_parent = dist6(c.p, nodes_[parent].p);
    nd.d_root = best;
    nd.generation = iter_;
    const int id = add_node(nd);
    nodes_[parent].children.push_back(id);
    if (cfg_.optimize) {
      for (size_t j = 0; j < c.knn.size(); ++j) {
        Node &nb = nodes_[c.knn[j]];
        const double d = dist6(nb.p, c.p);
        const double proposed = best + d;
        if (proposed < nb.d_root - kTol && c.e_rewire[j] >= 0 && eb2.free_flag[c.e_rewire[j]] && nb.parent >= 0) {
          std::vector<int> &ch = nodes_[nb.parent].children;
          auto it = std::find(ch.begin(), ch.end(), c.knn[j]);
          if (it != ch.end()) ch.erase(it);
          nb.parent = id;
          nb.d_parent = d;
          nb.d_root = proposed;                 // descendants keep their stored cost, as in the reference
          nodes_[id].children.push_back(c.knn[j]);
        }
      }
    }
    if (use_priority_) push_to_heaps(id);
    else frontier_.push_back(id);
    return id;
  }

  std::vector<Border> &borders(int t1, int t2) { return borders_[{std::min(t1, t2), std::max(t1, t2)}]; }

  void add_border(int n1, int n2) {
    Border bd{std::min(n1, n2), std::max(n1, n2)};
    std::vector<Border> &v = borders(nodes_[n1].tree, nodes_[n2].tree);
    for (const Border &o : v)
      if (o.a == bd.a && o.b == bd.b) return;
    v.push_back(bd);
  }

  // SpaceForest::maxConnected, forest.h:378-418
  int max_connected() {
    const int T = n_trees_;
    std::vector<char> seen(T, 0);
    int max_conn = 0, remaining = T, start = 0;
    while (max_conn < remaining) {
      connected_.clear();
      std::vector<int> stack{start};
      seen[start] = 1;
      while (!stack.empty()) {
        const int r = stack.front();
        stack.erase(stack.begin());
        connected_.push_back(r);
        for (int i = 0; i < T; ++i)
          if (i != r && !seen[i] && !borders(r, i).empty()) {
            seen[i] = 1;
            stack.insert(stack.begin(), i);
          }
      }
      max_conn = (int)connected_.size();
      for (int i = 0; i < T; ++i)
        if (!seen[i]) {
          start = i;
          break;
        }
      remaining -= max_conn;
    }
    return max_conn;
  }

  // SpaceForest::getPaths (forest.h:420-463) + Solver::getAllPaths (problemStruct.h:183-253)
  void build_paths() {
    const int T = n_trees_;
    for (int i = 0; i < T; ++i)
      for (int j = i + 1; j < T; ++j) {
        const std::vector<Border> &bs = borders(i, j);
        if (bs.empty()) continue;
        Link best;
        double best_d = -1;
        for (const Border &b : bs) {
          const double d = nodes_[b.a].d_root + nodes_[b.b].d_root + dist6(nodes_[b.a].p, nodes_[b.b].p);
          if (best_d == -1 || d < best_d - kTol) {
            best_d = d;
            best.n1 = b.a;
            best.n2 = b.b;
            best.distance = d;
          }
        }
        // plan: root of n1's tree ... n1, n2 ... root of n2's tree
        std::vector<int> left, right;
        for (int n = best.n1; n >= 0; n = nodes_[n].parent) left.insert(left.begin(), n);
        for (int n = best.n2; n >= 0; n = nodes_[n].parent) right.push_back(n);
        best.plan = left;
        best.plan.insert(best.plan.end(), right.begin(), right.end());
        // the border is chosen on the stored costs as in the reference; the reported length is the true length of the plan
        // (stored costs of a rewired node's descendants are stale -- never updated, forest.h:346-347 -- hence >= the truth)
        best.distance = book_.plan_length(best.plan);
        book_.link(i, j) = best;
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
  sffg_index *global_idx_ = nullptr;
  std::vector<sffg_index *> tree_idx_;
  std::vector<Node> nodes_;
  std::vector<std::vector<int>> members_;   // per tree: local index id -> global node id
  std::vector<int> pending_;
  std::vector<int32_t> radius_ids_;
  std::vector<float> radius_d2_;
  std::vector<int> frontier_, closed_;
  int n_trees_ = 0, goal_node_ = -1;
  bool use_priority_ = false, goal_reached_ = false;
  std::vector<std::vector<TargetHeap>> heaps_;          // Tree::frontiers, src/primitives.h:509
  std::map<int, std::vector<int>> heap_targets_;        // per tree: target node of every heap
  std::vector<std::pair<TargetHeap *, int>> deferred_push_;
  std::map<std::pair<int, int>, std::vector<Border>> borders_;
  PlanBook book_;
  std::vector<int> connected_;
  long iter_ = 0, rounds_ = 0, calls_ = 0, n_poses_ = 0, n_edges_ = 0, n_queries_ = 0;
  bool solved_ = false;
  double elapsed_ = 0;
};

}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) {
    std::cout << "Missing problem configuration file!\n";
    return 1;
  }
  std::string run_id = "0";
  uint64_t seed = (uint64_t)std::chrono::high_resolution_clock::now().time_since_epoch().count();
  int batch = 256;   // nodes per round: 256 measured faster than 128 at equal or better trees / path lengths, 512 loses trees
  bool quiet = false;
  std::string paths_file;
  int positional = 0;
  for (int i = 2; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--seed" && i + 1 < argc) seed = std::strtoull(argv[++i], nullptr, 10);
    else if (a == "--batch" && i + 1 < argc) batch = std::max(1, std::atoi(argv[++i]));
    else if (a == "--quiet") quiet = true;
    else if (a == "--paths" && i + 1 < argc) paths_file = argv[++i];
    else if (positional++ == 0) run_id = a;
  }
  Config cfg = load_config(argv[1]);
  const auto t0 = std::chrono::steady_clock::now();
  check(sffg_init(-1));
  const auto t1 = std::chrono::steady_clock::now();
  const SaveOptions save = load_save_options(argv[1], run_id, cfg.smoothing, cfg.solver == "sff");
  auto run = [&](auto &planner) {
    planner.set_save(save);
    planner.load();
    const auto t2 = std::chrono::steady_clock::now();
    planner.solve();
    planner.save_params(run_id);
    if (!paths_file.empty()) planner.dump_plans(paths_file);
    planner.report();
    if (!quiet)
      std::cout << "start-up seconds: device init " << std::chrono::duration<double>(t1 - t0).count() << ", meshes + BVH + indices "
                << std::chrono::duration<double>(t2 - t1).count() << "\n";
  };
  if (cfg.solver == "rrt") {
    RrtPlanner planner(cfg, seed, batch, quiet);
    run(planner);
  } else if (cfg.solver == "lazy") {
    LazyPlanner planner(cfg, seed, batch, quiet);
    run(planner);
  } else {
    Planner planner(cfg, seed, batch, quiet);
    run(planner);
  }
  return 0;
}
