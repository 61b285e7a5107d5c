// bvh_build.cpp -- host-side construction of the obstacle acceleration structure.
//
// The reference hands its triangles to RAPID, which builds an OBB tree at EndModel() (call site
// reference src/environment.h:114).  Any conservative hierarchy yields the same verdicts as long as every
// surviving triangle pair goes through the exact test (last bullet), so the GPU structure is
// chosen for the hardware instead: a world-space AABB BVH, binned-SAH built, collapsed to 8 children per node so
// that 8 lanes test one node with two coalesced 128-byte loads and a warp tests 4 nodes per step.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "common.h"

namespace sffg {

float round_up_f32(double v) {
  float f = (float)v;
  if ((double)f < v) f = std::nextafter(f, std::numeric_limits<float>::infinity());
  return f;
}
float round_down_f32(double v) {
  float f = (float)v;
  if ((double)f > v) f = std::nextafter(f, -std::numeric_limits<float>::infinity());
  return f;
}

namespace {

struct Box {
  double lo[3], hi[3];
  void reset() {
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::numeric_limits<double>::max();
      hi[k] = -std::numeric_limits<double>::max();
    }
  }
  void grow(const Box &o) {
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::min(lo[k], o.lo[k]);
      hi[k] = std::max(hi[k], o.hi[k]);
    }
  }
  void grow_pt(const double *p) {
    for (int k = 0; k < 3; ++k) {
      lo[k] = std::min(lo[k], p[k]);
      hi[k] = std::max(hi[k], p[k]);
    }
  }
  double area() const {
    double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return 2.0 * (dx * dy + dy * dz + dz * dx);
  }
};

struct BinNode {
  Box box;
  int left = -1, right = -1;   // binary children, -1 for a leaf
  int first = 0, count = 0;    // range in the index array
};

struct Builder {
  const double *tris;
  int64_t n;
  std::vector<Box> tbox;
  std::vector<double> cent;    // 3 per triangle
  std::vector<int32_t> idx;
  std::vector<BinNode> nodes;

  static constexpr int kBins = 16;

  int build(int first, int count) {
    int me = (int)nodes.size();
    nodes.emplace_back();
    Box b, cb;
    b.reset();
    cb.reset();
    for (int i = first; i < first + count; ++i) {
      b.grow(tbox[idx[i]]);
      cb.grow_pt(&cent[3 * (size_t)idx[i]]);
    }
    nodes[me].box = b;
    nodes[me].first = first;
    nodes[me].count = count;
    if (count == 1) return me;

    // binned SAH over the three axes
    double best_cost = std::numeric_limits<double>::max();
    int best_axis = -1, best_bin = -1;
    for (int ax = 0; ax < 3; ++ax) {
      double ext = cb.hi[ax] - cb.lo[ax];
      if (!(ext > 0)) continue;
      Box bb[kBins];
      int bc[kBins];
      for (int k = 0; k < kBins; ++k) {
        bb[k].reset();
        bc[k] = 0;
      }
      double inv = kBins / ext;
      for (int i = first; i < first + count; ++i) {
        int k = (int)((cent[3 * (size_t)idx[i] + ax] - cb.lo[ax]) * inv);
        k = std::min(std::max(k, 0), kBins - 1);
        bb[k].grow(tbox[idx[i]]);
        bc[k]++;
      }
      double ra[kBins];
      int rc[kBins];
      Box acc;
      acc.reset();
      int c = 0;
      for (int k = kBins - 1; k > 0; --k) {
        if (bc[k]) acc.grow(bb[k]);
        c += bc[k];
        ra[k] = c ? acc.area() : 0.0;
        rc[k] = c;
      }
      acc.reset();
      c = 0;
      for (int k = 0; k < kBins - 1; ++k) {
        if (bc[k]) acc.grow(bb[k]);
        c += bc[k];
        if (c == 0 || rc[k + 1] == 0) continue;
        double cost = acc.area() * c + ra[k + 1] * rc[k + 1];
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = ax;
          best_bin = k;
        }
      }
    }
    int mid;
    if (best_axis >= 0) {
      double ext = cb.hi[best_axis] - cb.lo[best_axis];
      double inv = kBins / ext;
      auto it = std::partition(idx.begin() + first, idx.begin() + first + count, [&](int32_t t) {
        int k = (int)((cent[3 * (size_t)t + best_axis] - cb.lo[best_axis]) * inv);
        k = std::min(std::max(k, 0), kBins - 1);
        return k <= best_bin;
      });
      mid = (int)(it - idx.begin());
    } else {
      mid = first + count / 2;   // all centroids coincide
    }
    if (mid == first || mid == first + count) mid = first + count / 2;
    int l = build(first, mid - first);
    int r = build(mid, first + count - mid);
    nodes[me].left = l;
    nodes[me].right = r;
    return me;
  }
};

struct Collapser {
  const Builder &b;
  HostBvh *out;
  std::vector<int32_t> tri_pos;   // original triangle -> position in leaf order

  // returns wide node index; depth tracked through `level`
  int emit(int bin, int level) {
    out->depth = std::max(out->depth, level + 1);
    int me = (int)(out->slots.size() / kWide);
    out->slots.resize(out->slots.size() + kWide);
    std::vector<int> kids;
    const BinNode &root = b.nodes[bin];
    if (root.left < 0) {
      kids.push_back(bin);   // single-triangle mesh: the root holds one leaf slot
    } else {
      kids.push_back(root.left);
      kids.push_back(root.right);
      while ((int)kids.size() < kWide) {
        int pick = -1;
        double best = -1;
        for (int i = 0; i < (int)kids.size(); ++i) {
          const BinNode &c = b.nodes[kids[i]];
          if (c.left < 0) continue;
          double a = c.box.area();
          if (a > best) {
            best = a;
            pick = i;
          }
        }
        if (pick < 0) break;
        int c = kids[pick];
        kids[pick] = b.nodes[c].left;
        kids.push_back(b.nodes[c].right);
      }
    }
    // keep spatially adjacent children adjacent: order by first triangle position
    std::sort(kids.begin(), kids.end(), [&](int x, int y) { return b.nodes[x].first < b.nodes[y].first; });
    for (int i = 0; i < kWide; ++i) {
      ChildSlot s;
      if (i < (int)kids.size()) {
        const BinNode &c = b.nodes[kids[i]];
        float h[3], ctr[3];
        for (int k = 0; k < 3; ++k) {
          double cd = 0.5 * (c.box.lo[k] + c.box.hi[k]);
          ctr[k] = (float)cd;
          double need = std::max(c.box.hi[k] - (double)ctr[k], (double)ctr[k] - c.box.lo[k]);
          h[k] = round_up_f32(need);
        }
        s.cx = ctr[0]; s.cy = ctr[1]; s.cz = ctr[2];
        s.hx = h[0]; s.hy = h[1]; s.hz = h[2];
        s.pad = 0.f;
        if (c.left < 0) {
          s.child = ~tri_pos[b.idx[c.first]];
        } else {
          int child = emit(kids[i], level + 1);
          s.child = child;
        }
      } else {
        s.cx = s.cy = s.cz = 3.0e38f;
        s.hx = s.hy = s.hz = -1.0f;
        s.pad = 0.f;
        s.child = kEmptyChild;
      }
      out->slots[(size_t)me * kWide + i] = s;
    }
    return me;
  }
};

}  // namespace

void top_cut(const HostBvh &bvh, std::vector<ChildSlot> *out) {
  out->clear();
  if (bvh.slots.empty()) return;
  auto push_node = [&](int node) {
    for (int i = 0; i < kWide; ++i) {
      const ChildSlot &s = bvh.slots[(size_t)node * kWide + i];
      if (s.child != kEmptyChild) out->push_back(s);
    }
  };
  push_node(0);
  while (true) {
    int pick = -1;
    float best = -1.f;
    for (int i = 0; i < (int)out->size(); ++i) {
      const ChildSlot &s = (*out)[i];
      if (s.child < 0) continue;
      if (((size_t)s.child + 1) * kWide > bvh.slots.size()) continue;   // only the head of a device-built hierarchy is on the host
      int kids = 0;
      for (int c = 0; c < kWide; ++c) kids += bvh.slots[(size_t)s.child * kWide + c].child != kEmptyChild;
      if ((int)out->size() - 1 + kids > 32) continue;
      const float area = s.hx * s.hy + s.hy * s.hz + s.hz * s.hx;
      if (area > best) {
        best = area;
        pick = i;
      }
    }
    if (pick < 0) break;
    const int node = (*out)[pick].child;
    out->erase(out->begin() + pick);
    push_node(node);
  }
}

void build_wide_bvh(const double *tris, int64_t n, HostBvh *out) {
  out->slots.clear();
  out->tri_order.clear();
  out->depth = 0;
  if (n <= 0) return;
  Builder b;
  b.tris = tris;
  b.n = n;
  b.tbox.resize((size_t)n);
  b.cent.resize(3 * (size_t)n);
  b.idx.resize((size_t)n);
  Box all;
  all.reset();
  for (int64_t t = 0; t < n; ++t) {
    Box bx;
    bx.reset();
    for (int v = 0; v < 3; ++v) bx.grow_pt(tris + 9 * t + 3 * v);
    b.tbox[(size_t)t] = bx;
    for (int k = 0; k < 3; ++k) b.cent[3 * (size_t)t + k] = 0.5 * (bx.lo[k] + bx.hi[k]);
    b.idx[(size_t)t] = (int32_t)t;
    all.grow(bx);
  }
  b.nodes.reserve(2 * (size_t)n);
  b.build(0, (int)n);
  // leaf order == final idx order (build() only permutes inside its own range)
  out->tri_order.assign(b.idx.begin(), b.idx.end());
  Collapser c{b, out, {}};
  c.tri_pos.resize((size_t)n);
  for (int64_t i = 0; i < n; ++i) c.tri_pos[(size_t)b.idx[(size_t)i]] = (int32_t)i;
  c.emit(0, 0);
  // emit() numbers the nodes depth-first; the kernels stage the first few hundred nodes in shared memory, so renumber
  // breadth-first: the top levels of the hierarchy become one contiguous prefix (what the device builder produces anyway)
  {
    const size_t nn = out->slots.size() / kWide;
    std::vector<int32_t> order;   // new index -> old index
    order.reserve(nn);
    order.push_back(0);
    out->level_base.clear();
    out->level_count.clear();
    for (size_t begin = 0, end = 1; begin < end; begin = end, end = order.size()) {
      out->level_base.push_back((int)begin);
      out->level_count.push_back((int)(end - begin));
      for (size_t head = begin; head < end; ++head)
        for (int i = 0; i < kWide; ++i) {
          const int32_t ch = out->slots[(size_t)order[head] * kWide + i].child;
          if (ch >= 0 && ch != kEmptyChild) order.push_back(ch);
        }
    }
    std::vector<int32_t> new_of(nn, -1);
    for (size_t i = 0; i < order.size(); ++i) new_of[(size_t)order[i]] = (int32_t)i;
    std::vector<ChildSlot> re(out->slots.size());
    for (size_t i = 0; i < order.size(); ++i)
      for (int k2 = 0; k2 < kWide; ++k2) {
        ChildSlot s2 = out->slots[(size_t)order[i] * kWide + k2];
        if (s2.child >= 0 && s2.child != kEmptyChild) s2.child = new_of[(size_t)s2.child];
        re[i * kWide + k2] = s2;
      }
    out->slots.swap(re);
  }
  for (int k = 0; k < 3; ++k) {
    out->root_lo[k] = all.lo[k];
    out->root_hi[k] = all.hi[k];
  }
}

}  // namespace sffg
