// lazy_planner.h -- Lazy-TSP host (reference LazyTSP<T,R>, src/lazy.h:71-147) on top of the engine's C ABI.
//
// The reference alternates between a TSP over the roots (lower-bound matrix = straight-line 6-D distances) and RRT / RRT*
// searches for the tour's edges that have no path yet, until the tour length stops changing (src/lazy.h:84-131).  Two
// things keep the reference's version from running here: its TSP is solved by an external, non-public binary (`obst_tsp`,
// reference README.md:14, src/lazy.h:93-98) and every edge search is the sequential one-sample-at-a-time loop.
//   * the tour is computed in-process: exact dynamic programming (Held-Karp) up to 13 roots, nearest neighbour + 2-opt above;
//   * every edge search (runRRT, src/lazy.h:159-284) is a batched RrtPlanner run -- one root, the other end as the goal, no
//     goal bias (":179  NO PRIORITY BIAS"), RRT* when optimize="true" -- sharing one engine environment.
// Deliberate differences: the reference accepts a search as soon as a new node is within dtree of the goal WITHOUT checking
// the last segment (src/lazy.h:262-275); here that segment must pass the local planner like any other (the always-on plan
// verifier would reject the plan otherwise).  An unreachable edge costs 1e18 instead of numeric_limits::max (whose sums
// overflow to +inf and make the reference report "solved").
//
// Output: the reference's params row for this solver (LazyTSP::saveParams, src/lazy.h:388-425):
//   id,run,iterations,solved|unsolved,[tour: first node of every edge],[edge lengths],elapsed seconds
#pragma once
#include <memory>

#include "rrt_planner.h"

namespace planner {

class LazyPlanner {
 public:
  LazyPlanner(const Config &cfg, uint64_t seed, int batch, bool quiet) : cfg_(cfg), seed_(seed), batch_(batch), quiet_(quiet) {
    book_.pos = [this](int id) { return poses_[id].data(); };
    book_.origin = [this](int id) { return id < n_ ? id : -1; };
  }
  ~LazyPlanner() {
    if (env_) sffg_env_destroy(env_);
  }

  void set_save(const SaveOptions &so) { save_ = so; }

  void load() {
    env_ = load_environment(cfg_);
    n_ = (int)cfg_.roots.size();
    dist_.assign((size_t)n_ * n_, 0.0);
    for (int i = 0; i < n_; ++i) {
      std::array<double, 6> p{cfg_.roots[i][0], cfg_.roots[i][1], cfg_.roots[i][2], 0, 0, 0};
      poses_.push_back(p);
    }
    for (int i = 0; i < n_; ++i)
      for (int j = i + 1; j < n_; ++j) d(i, j) = d(j, i) = dist6(poses_[i].data(), poses_[j].data());   // lazy.h:52-58
  }

  void solve() {
    const auto t0 = std::chrono::steady_clock::now();
    double prev = -1, now = 0;
    const long budget = (long)n_ * cfg_.max_iterations;   // lazy.h:84
    while (!solved_ && iter_ < budget) {
      prev = now;
      tour_ = n_ >= 2 ? solve_tsp() : std::vector<int>{0};
      now = 0;
      bool all_reachable = true;
      const long searches_before = searches_;
      for (size_t e = 0; e < tour_.size() && n_ >= 2; ++e) {
        const int a = tour_[e], b = tour_[(e + 1) % tour_.size()];
        if (!book_.link(a, b).exists() && d(a, b) < kUnreachable) search_edge(a, b);
        all_reachable &= d(a, b) < kUnreachable;
        now += d(a, b);
      }
      ++passes_;
      solved_ = all_reachable && now >= prev - kTol && now <= prev + kTol;   // lazy.h:130
      if (n_ < 2) solved_ = true;
      // a pass that searched nothing leaves the matrix, hence the next tour and `now`, unchanged: with an unreachable tour
      // edge that is the reference's `newDist == prevDist` exit (lazy.h:128), reported as unsolved instead of spinning
      if (!all_reachable && searches_ == searches_before) break;
    }
    elapsed_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    tour_length_ = now;
    book_.verify(env_, cfg_.has_map, calls_);
    write_outputs();
  }

  void save_params(const std::string &run_id) const {
    if (cfg_.params_file.empty()) return;
    std::ofstream out(cfg_.params_file, std::ios_base::app);
    if (!out.good()) {
      std::cout << "Cannot create file at: " << cfg_.params_file << "\n";
      return;
    }
    out << cfg_.id << "," << run_id << "," << iter_ << "," << (solved_ ? "solved" : "unsolved") << ",[";
    for (size_t e = 0; e < tour_.size(); ++e) out << tour_[e] << (e + 1 != tour_.size() ? ";" : "");
    out << "],[";
    for (size_t e = 0; e < tour_.size(); ++e)
      out << d(tour_[e], tour_[(e + 1) % tour_.size()]) / cfg_.scale << (e + 1 != tour_.size() ? ";" : "");
    out << "]," << elapsed_ << "\n";
  }

  void dump_plans(const std::string &file) const { book_.save_paths(file); }

  void report() const {
    if (quiet_) return;
    std::cout << "roots " << n_ << ", TSP passes " << passes_ << ", edge searches " << searches_ << ", iterations " << iter_ << ", "
              << (solved_ ? "solved" : "unsolved") << ", tour length " << tour_length_ << ", elapsed " << elapsed_ << " s, engine calls "
              << calls_ << "\nverified segments " << book_.verified_segments << "\n";
  }

 private:
  static constexpr double kUnreachable = 1e18;
  double &d(int i, int j) { return dist_[(size_t)i * n_ + j]; }
  double d(int i, int j) const { return dist_[(size_t)i * n_ + j]; }

  // one RRT / RRT* search from root a to root b (runRRT, src/lazy.h:159-284) as a batched RrtPlanner run
  void search_edge(int a, int b) {
    Config c = cfg_;
    c.solver = "rrt";
    c.roots = {cfg_.roots[a]};
    c.has_goal = true;
    for (int k = 0; k < 3; ++k) c.goal[k] = cfg_.roots[b][k];
    c.priority_bias = 0;
    c.smoothing = false;
    c.params_file.clear();
    RrtPlanner rrt(c, seed_ + 7919ull * (uint64_t)(++searches_), batch_, true);
    rrt.use_env(env_);
    rrt.load();
    rrt.solve();
    iter_ += rrt.iterations();
    calls_ += rrt.calls();
    const Link &found = rrt.book().link(0, 1);
    if (!rrt.solved() || !found.exists()) {
      d(a, b) = d(b, a) = kUnreachable;
      return;
    }
    // the plan runs root a ... goal b in the search's node ids: store its interior poses here, ends are the root poses
    Link l;
    l.plan.push_back(a);
    for (size_t k = 1; k + 1 < found.plan.size(); ++k) {
      std::array<double, 6> p;
      std::memcpy(p.data(), rrt.node_pos(found.plan[k]), sizeof p);
      poses_.push_back(p);
      l.plan.push_back((int)poses_.size() - 1);
    }
    l.plan.push_back(b);
    l.n1 = a;
    l.n2 = b;
    l.distance = book_.plan_length(l.plan);
    if (a > b) std::reverse(l.plan.begin(), l.plan.end());   // links are stored from the lower to the higher root
    book_.link(a, b) = l;
    d(a, b) = d(b, a) = l.distance;
  }

  // closed tour over all roots starting at root 0: exact (Held-Karp) for small instances, nearest neighbour + 2-opt above
  std::vector<int> solve_tsp() const {
    if (n_ <= 3) {
      std::vector<int> t(n_);
      for (int i = 0; i < n_; ++i) t[i] = i;
      return t;
    }
    if (n_ <= 13) {
      const int full = 1 << (n_ - 1);   // subsets of roots 1..n-1
      std::vector<double> dp((size_t)full * (n_ - 1), std::numeric_limits<double>::infinity());
      std::vector<int> from((size_t)full * (n_ - 1), -1);
      for (int j = 0; j < n_ - 1; ++j) dp[(size_t)(1 << j) * (n_ - 1) + j] = d(0, j + 1);
      for (int mask = 1; mask < full; ++mask)
        for (int j = 0; j < n_ - 1; ++j) {
          if (!(mask & (1 << j))) continue;
          const double cur = dp[(size_t)mask * (n_ - 1) + j];
          if (!(cur < std::numeric_limits<double>::infinity())) continue;
          for (int k = 0; k < n_ - 1; ++k) {
            if (mask & (1 << k)) continue;
            const int nm = mask | (1 << k);
            const double v = cur + d(j + 1, k + 1);
            if (v < dp[(size_t)nm * (n_ - 1) + k]) {
              dp[(size_t)nm * (n_ - 1) + k] = v;
              from[(size_t)nm * (n_ - 1) + k] = j;
            }
          }
        }
      int best = 0;
      double best_len = std::numeric_limits<double>::infinity();
      for (int j = 0; j < n_ - 1; ++j) {
        const double v = dp[(size_t)(full - 1) * (n_ - 1) + j] + d(j + 1, 0);
        if (v < best_len) {
          best_len = v;
          best = j;
        }
      }
      std::vector<int> rev;
      int mask = full - 1, j = best;
      while (j >= 0) {
        rev.push_back(j + 1);
        const int pj = from[(size_t)mask * (n_ - 1) + j];
        mask &= ~(1 << j);
        j = pj;
      }
      std::vector<int> t{0};
      t.insert(t.end(), rev.rbegin(), rev.rend());
      return t;
    }
    std::vector<int> t{0};
    std::vector<char> used(n_, 0);
    used[0] = 1;
    while ((int)t.size() < n_) {
      int best = -1;
      for (int k = 0; k < n_; ++k)
        if (!used[k] && (best < 0 || d(t.back(), k) < d(t.back(), best))) best = k;
      used[best] = 1;
      t.push_back(best);
    }
    for (bool improved = true; improved;) {
      improved = false;
      for (int i = 1; i + 1 < n_; ++i)
        for (int k = i + 1; k < n_; ++k) {
          const int a = t[i - 1], b = t[i], c = t[k], e = t[(k + 1) % n_];
          if (d(a, c) + d(b, e) < d(a, b) + d(c, e) - kTol) {
            std::reverse(t.begin() + i, t.begin() + k + 1);
            improved = true;
          }
        }
    }
    return t;
  }

  void write_outputs() const {
    // LazyTSP::savePaths (src/lazy.h:335-385): the plans of the selected edges, in tour order
    if (save_.raw_path.set()) {
      std::ofstream out;
      if (open_out(out, save_.raw_path)) {
        if (save_.raw_path.is_obj) {
          out << "o Paths\n";
          for (const auto &p : poses_) {
            out << "v ";
            put_pos(out, p.data(), cfg_.scale);
            out << "\n";
          }
        }
        for (size_t e = 0; e < tour_.size() && n_ >= 2; ++e) {
          const Link &l = book_.link(tour_[e], tour_[(e + 1) % tour_.size()]);
          for (size_t k = 0; k + 1 < l.plan.size(); ++k) {
            if (save_.raw_path.is_obj) {
              out << "l " << l.plan[k] + 1 << " " << l.plan[k + 1] + 1 << "\n";
            } else {
              put_point(out, poses_[l.plan[k]].data(), cfg_.scale);
              out << " ";
              put_point(out, poses_[l.plan[k + 1]].data(), cfg_.scale);
              out << "\n";
            }
          }
          if (!save_.raw_path.is_obj) out << "\n";
        }
      }
    }
    // LazyTSP::saveTsp (src/lazy.h:300-332): the full lower-diagonal matrix over all roots
    if (save_.tsp.set()) {
      std::ofstream out;
      if (open_out(out, save_.tsp)) {
        out << "NAME: " << cfg_.id << "\nCOMMENT:\nTYPE: TSP\nDIMENSION: " << n_
            << "\nEDGE_WEIGHT_TYPE : EXPLICIT\nEDGE_WEIGHT_FORMAT : LOWER_DIAG_ROW\nEDGE_WEIGHT_SECTION\n";
        for (int i = 0; i < n_; ++i) {
          for (int j = 0; j < i; ++j) out << d(i, j) / cfg_.scale << " ";
          out << "0\n";
        }
      }
    }
  }

  Config cfg_;
  SaveOptions save_;
  uint64_t seed_;
  int batch_;
  bool quiet_;
  sffg_env *env_ = nullptr;
  int n_ = 0;
  std::vector<double> dist_;
  std::vector<std::array<double, 6>> poses_;   // roots first, then the interior nodes of every stored plan
  PlanBook book_;
  std::vector<int> tour_;
  long iter_ = 0, calls_ = 0, passes_ = 0, searches_ = 0;
  bool solved_ = false;
  double elapsed_ = 0, tour_length_ = 0;
};

}  // namespace planner
