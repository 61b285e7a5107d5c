// knn_pruned.cu -- spatially sorted view of an index + block-pruned exact k-NN scan, sm_100a.
//
// The brute-force scan (knn_kernels.cu) touches every node for every query.  For large node sets the index also keeps a
// copy of its nodes sorted along a Morton curve of the translational coordinates, cut into blocks of 32 nodes with the
// exact bounding box of each block.  A query then only has to *visit* the blocks whose box could still hold one of its k
// nearest nodes: the box distance is computed with the same float operations, in the same order, as the translational part
// of the metric, and every one of those operations is monotone, so box_lb <= metric_lin(node) <= metric(node) holds
// bit-for-bit for every node of the block -- skipping a block with box_lb > (current k-th distance) can never drop a
// neighbour, and results stay identical (ids AND float distances) to the exhaustive scan / FLANN's LinearIndex.
// Because blocks are visited in spatial, not id, order, insertion uses the (d2, id)-keyed variant.
//
// One warp = 8 queries.  Per step it tests 32 block boxes (lane-per-box) against its 8 queries, then visits the blocks
// whose bit is set in the ballot (lane-per-node, as in the exhaustive scan).  Nodes appended after the last rebuild (the
// "tail") are covered by the exhaustive kernel; the partial lists are merged by knn_merge_kernel.
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>

#include "knn_common.cuh"
#include "knn_kernels.cuh"
#include "knn_pruned.cuh"

namespace sffg {
namespace {

using namespace knn;

// ---- build ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_order(float f) {   // order-preserving float -> int
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float order_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// bounds[0..2] = min, bounds[3..5] = max of the translational coordinates (as ordered ints); LIN = 2 or 3
__global__ void bounds_kernel(const float *__restrict__ coords, long long cap, int n, int lin, int *bounds) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int mn[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, mx[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  if (i < n)
    for (int c = 0; c < lin; ++c) mn[c] = mx[c] = float_order(coords[(long long)c * cap + i]);
  for (int c = 0; c < lin; ++c) {
    for (int s = 16; s > 0; s >>= 1) {
      mn[c] = min(mn[c], __shfl_xor_sync(kFull, mn[c], s));
      mx[c] = max(mx[c], __shfl_xor_sync(kFull, mx[c], s));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(bounds + c, mn[c]);
      atomicMax(bounds + 3 + c, mx[c]);
    }
  }
}

__device__ __forceinline__ unsigned spread3(unsigned v) {   // 10 bits -> every third bit
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__device__ __forceinline__ unsigned spread2(unsigned v) {   // 15 bits -> every second bit
  v = (v | (v << 8)) & 0x00FF00FFu;
  v = (v | (v << 4)) & 0x0F0F0F0Fu;
  v = (v | (v << 2)) & 0x33333333u;
  v = (v | (v << 1)) & 0x55555555u;
  return v;
}

__global__ void morton_kernel(const float *__restrict__ coords, long long cap, int n, int lin, const int *bounds, unsigned *keys,
                              unsigned *vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned q[3] = {0, 0, 0};
  const float levels = lin == 3 ? 1023.f : 32767.f;
  for (int c = 0; c < lin; ++c) {
    const float lo = order_float(bounds[c]), hi = order_float(bounds[3 + c]);
    const float span = hi - lo;
    const float t = span > 0.f ? (coords[(long long)c * cap + i] - lo) / span : 0.f;
    q[c] = (unsigned)fminf(fmaxf(t * levels, 0.f), levels);
  }
  keys[i] = lin == 3 ? (spread3(q[0]) | (spread3(q[1]) << 1) | (spread3(q[2]) << 2)) : (spread2(q[0]) | (spread2(q[1]) << 1));
  vals[i] = (unsigned)i;
}

// Morton key of every query row (AoS [nq][dim]) on the grid the node set was sorted with; rows outside it clamp
__global__ void query_morton_kernel(const float *__restrict__ q, int nq, int dim, int lin, const int *bounds, unsigned *keys, unsigned *vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  unsigned c3[3] = {0, 0, 0};
  const float levels = lin == 3 ? 1023.f : 32767.f;
  for (int c = 0; c < lin; ++c) {
    const float lo = order_float(bounds[c]), hi = order_float(bounds[3 + c]);
    const float span = hi - lo;
    const float t = span > 0.f ? (q[(long long)i * dim + c] - lo) / span : 0.f;
    c3[c] = (unsigned)fminf(fmaxf(t * levels, 0.f), levels);   // (NaN -> 0)
  }
  keys[i] = lin == 3 ? (spread3(c3[0]) | (spread3(c3[1]) << 1) | (spread3(c3[2]) << 2)) : (spread2(c3[0]) | (spread2(c3[1]) << 1));
  vals[i] = (unsigned)i;
}

__global__ void gather_kernel(const float *__restrict__ coords, long long cap, int dim, int n, const unsigned *__restrict__ order,
                              float *s_coords, long long cap_s, int *s_ids) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const unsigned src = order[p];
  s_ids[p] = (int)src;
  for (int c = 0; c < dim; ++c) s_coords[(long long)c * cap_s + p] = coords[(long long)c * cap + src];
}

// one warp per block of 32 sorted nodes: exact min / max of each translational coordinate
__global__ void block_box_kernel(const float *__restrict__ s_coords, long long cap_s, int n, int lin, float *bb, long long nblk_cap) {
  const int lane = threadIdx.x & 31;
  const int blk = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nblk = (n + 31) / 32;
  if (blk >= nblk) return;
  const int p = blk * 32 + lane;
  for (int c = 0; c < lin; ++c) {
    float lo = INFINITY, hi = -INFINITY;
    if (p < n) lo = hi = s_coords[(long long)c * cap_s + p];
    for (int s = 16; s > 0; s >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(kFull, lo, s));
      hi = fmaxf(hi, __shfl_xor_sync(kFull, hi, s));
    }
    if (lane == 0) {
      bb[(long long)c * nblk_cap + blk] = lo;
      bb[(long long)(lin + c) * nblk_cap + blk] = hi;
    }
  }
}

// one warp per superblock of 32 blocks: union of the block boxes
__global__ void superblock_box_kernel(const float *__restrict__ bb, long long nblk_cap, int nblk, int lin, float *sbb, long long nsb_cap) {
  const int lane = threadIdx.x & 31;
  const int sb = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int nsb = (nblk + 31) / 32;
  if (sb >= nsb) return;
  const int b = sb * 32 + lane;
  for (int c = 0; c < lin; ++c) {
    float lo = INFINITY, hi = -INFINITY;
    if (b < nblk) {
      lo = bb[(long long)c * nblk_cap + b];
      hi = bb[(long long)(lin + c) * nblk_cap + b];
    }
    for (int s = 16; s > 0; s >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(kFull, lo, s));
      hi = fmaxf(hi, __shfl_xor_sync(kFull, hi, s));
    }
    if (lane == 0) {
      sbb[(long long)c * nsb_cap + sb] = lo;
      sbb[(long long)(lin + c) * nsb_cap + sb] = hi;
    }
  }
}

// ---- pruned scan --------------------------------------------------------------------------------------------------
// lower bound of metric_lin over a box, same operations and order as metric_lin (all monotone): for every node n of the
// box, box_lb(q) <= metric_lin(n, q) holds exactly in float arithmetic
template <int LIN>
__device__ __forceinline__ float box_lb(const float *lo, const float *hi, const float *q) {
  float r = 0.f;
#pragma unroll
  for (int c = 0; c < LIN; ++c) {
    // |n_c - q_c| >= max(lo_c - q_c, q_c - hi_c, 0) and rounding of the subtraction is monotone
    const float t = fmaxf(fmaxf(__fsub_rn(lo[c], q[c]), __fsub_rn(q[c], hi[c])), 0.f);
    const float sq = __fmul_rn(t, t);
    r = c == 0 ? sq : __fadd_rn(r, sq);
  }
  return r;
}

// upper bound of metric_lin over a box (same monotone operations): every node n of the box has metric_lin(n, q) <= box_ub(q)
template <int LIN>
__device__ __forceinline__ float box_ub(const float *lo, const float *hi, const float *q) {
  float r = 0.f;
#pragma unroll
  for (int c = 0; c < LIN; ++c) {
    const float t = fmaxf(fabsf(__fsub_rn(lo[c], q[c])), fabsf(__fsub_rn(hi[c], q[c])));
    const float sq = __fmul_rn(t, t);
    r = c == 0 ? sq : __fadd_rn(r, sq);
  }
  return r;
}

// resident CTAs per SM the pruned kernel is compiled for: 4 (64 registers, 32 warps per SM) measured 65.9M queries/s at
// N = 1e6, k = 16 against 51.8M at 2 CTAs / 128 registers and 61.9M at 3 / 80 -- the search is latency-bound (shuffle
// chains of the top-k insertion), more warps hide it better than more registers do
#ifndef SFFG_KNN_MIN_BLOCKS
#define SFFG_KNN_MIN_BLOCKS 4
#endif
template <int DIM, int QW, int KPL>
__global__ void __launch_bounds__(kThreads, SFFG_KNN_MIN_BLOCKS) knn_pruned_kernel(SortedDev sv, const float *__restrict__ queries, long long nq, int k,
                                                              int slices, int sb_per_slice, float *out_d, int *out_i,
                                                              int slot_base, int slots_total, const unsigned *__restrict__ perm,
                                                              RowDests rows) {
  constexpr int LIN = DIM == 6 ? 3 : 2;
  const int lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long group = item / slices;
  const int slice = (int)(item - group * slices);
  if (group * QW >= nq) return;
  float q[QW][DIM];
  TopK<KPL> top[QW];
  float worst[QW];
  int worst_id[QW];
  long long row[QW];   // caller's row of the w-th query of this warp (perm: visiting order -> row)
#pragma unroll
  for (int w = 0; w < QW; ++w) {
    long long qi = group * QW + w;
    if (qi >= nq) qi = nq - 1;
    row[w] = perm ? (long long)__ldg(perm + qi) : qi;
#pragma unroll
    for (int c = 0; c < DIM; ++c) q[w][c] = __ldg(queries + row[w] * DIM + c);
    top[w].init();
    worst[w] = INFINITY;
    worst_id[w] = -1;
  }
  // un-normalised angles: warp-uniform choice of the exact wide wrap, and the seed's bound on the angular part
  bool wide = false;
  float ang_ub[QW];
  if (DIM == 6) {
    const float node_amax = __uint_as_float(__ldg(sv.amax));
#pragma unroll
    for (int w = 0; w < QW; ++w) {
      wide |= wide_needed(q[w], node_amax);
      ang_ub[w] = angular_part_ub(q[w], node_amax);
    }
  }
  const int nsb = (sv.nblk + 31) / 32;   // superblocks of 32 blocks
  const int sb_begin = slice * sb_per_slice;
  const int sb_end = min(nsb, sb_begin + sb_per_slice);
  int near_sb = -1;   // a superblock close to the warp's queries (they are neighbours in space): pass 2 starts there
  // ---- pass 1: a bound on the k-th distance without touching a node.
  // Slices of up to 128 superblocks scan every block box of the slice: G = ceil(k/32) consecutive FULL blocks hold >= k nodes,
  // all within max_g box_ub(g) (+ the largest possible angular part), so the k-th nearest node cannot be farther -- the
  // tightest bound, at one step per superblock.  Larger slices take the superblock route below.
  // The bound seeds worst[] (inclusive: worst_id = -1 compares as the largest id), which lets pass 2 prune from the start.
  if (sb_end - sb_begin <= 128) {
    const int G = (k + 31) / 32;
    const int full_blocks = sv.n_sorted / 32;
    float best[QW];
#pragma unroll
    for (int w = 0; w < QW; ++w) best[w] = INFINITY;
    for (int sb = sb_begin; sb < sb_end; ++sb) {
      const int blk = sb * 32 + lane;
      float ub[QW];
      if (blk < full_blocks) {
        float lo[LIN], hi[LIN];
#pragma unroll
        for (int c = 0; c < LIN; ++c) {
          lo[c] = __ldg(sv.bb + (long long)c * sv.nblk_cap + blk);
          hi[c] = __ldg(sv.bb + (long long)(LIN + c) * sv.nblk_cap + blk);
        }
#pragma unroll
        for (int w = 0; w < QW; ++w) ub[w] = box_ub<LIN>(lo, hi, q[w]);
      } else {
#pragma unroll
        for (int w = 0; w < QW; ++w) ub[w] = INFINITY;
      }
#pragma unroll
      for (int w = 0; w < QW; ++w) {
        float g = ub[w];
        for (int j = 1; j < G; ++j) {
          const float o = __shfl_down_sync(kFull, ub[w], j);
          g = fmaxf(g, lane + j < 32 ? o : INFINITY);
        }
        best[w] = fminf(best[w], g);
      }
    }
#pragma unroll
    for (int w = 0; w < QW; ++w) {
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) best[w] = fminf(best[w], __shfl_xor_sync(kFull, best[w], sft));
      // + the largest possible angular part; the relative margin covers the three roundings of the accumulation
      worst[w] = DIM == 6 ? __fmul_ru(__fadd_ru(best[w], ang_ub[w]), 1.000001f) : best[w];
    }
  } else
  // Superblock route (lane-per-superblock first): a FULL
  // superblock holds 1024 >= k nodes, all within box_ub of its box.  Inside the best superblock of every query the bound is
  // refined at block level: G = ceil(k/32) consecutive blocks (all full) hold >= k nodes, all within max_g box_ub(g).  Plus
  // the largest possible angular part.  The bound seeds worst[] (inclusive: worst_id = -1 compares as the largest id),
  // which lets pass 2 prune from the start.
  {
    const int G = (k + 31) / 32;
    const int full_sb = sv.n_sorted / 1024;
    float best[QW];
    int arg[QW];
#pragma unroll
    for (int w = 0; w < QW; ++w) {
      best[w] = INFINITY;
      arg[w] = -1;
    }
    for (int g = sb_begin; g < sb_end; g += 32) {
      const int s = g + lane;
      if (s < sb_end && s < full_sb) {
        float lo[LIN], hi[LIN];
#pragma unroll
        for (int c = 0; c < LIN; ++c) {
          lo[c] = __ldg(sv.sbb + (long long)c * sv.nsb_cap + s);
          hi[c] = __ldg(sv.sbb + (long long)(LIN + c) * sv.nsb_cap + s);
        }
#pragma unroll
        for (int w = 0; w < QW; ++w) {
          const float u = box_ub<LIN>(lo, hi, q[w]);
          if (u < best[w]) {
            best[w] = u;
            arg[w] = s;
          }
        }
      }
    }
#pragma unroll
    for (int w = 0; w < QW; ++w) {
      float bv = best[w];
      int bs = arg[w];
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        const float ov = __shfl_xor_sync(kFull, bv, sft);
        const int os = __shfl_xor_sync(kFull, bs, sft);
        if (ov < bv || (ov == bv && (unsigned)os < (unsigned)bs)) {
          bv = ov;
          bs = os;
        }
      }
      float bound = bv;
      if (w == 0) near_sb = bs;
      if (bs >= 0) {   // warp-uniform; every block of a full superblock is full
        const int blk = bs * 32 + lane;
        float lo[LIN], hi[LIN];
#pragma unroll
        for (int c = 0; c < LIN; ++c) {
          lo[c] = __ldg(sv.bb + (long long)c * sv.nblk_cap + blk);
          hi[c] = __ldg(sv.bb + (long long)(LIN + c) * sv.nblk_cap + blk);
        }
        const float ub = box_ub<LIN>(lo, hi, q[w]);
        float gmax = ub;
        for (int j = 1; j < G; ++j) {
          const float o = __shfl_down_sync(kFull, ub, j);
          gmax = fmaxf(gmax, lane + j < 32 ? o : INFINITY);
        }
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) gmax = fminf(gmax, __shfl_xor_sync(kFull, gmax, sft));
        bound = fminf(bound, gmax);
      }
      worst[w] = DIM == 6 ? __fmul_ru(__fadd_ru(bound, ang_ub[w]), 1.000001f) : bound;
    }
  }
  float seed[QW];
#pragma unroll
  for (int w = 0; w < QW; ++w) seed[w] = worst[w];
  // ---- pass 2: superblock boxes first (lane-per-superblock, 32 768 nodes per step), then the block boxes of the superblocks
  // that can still hold a candidate (lane-per-block), then the nodes of the blocks that can
  // Visiting order: nearest first at every level (the group of 32 superblocks around `near_sb` first, then inside a group
  // the superblock and inside a superblock the block whose box is closest to any of the warp's queries): the k-th
  // distances tighten after a few blocks, and almost everything visited later fails the `d < worst` vote without a
  // single insertion.  The order changes nothing in the result -- insertion is (d2, id)-keyed.
  const int first_sg = near_sb >= 0 ? sb_begin + ((near_sb - sb_begin) & ~31) : -1;
  for (int gi = first_sg >= 0 ? -1 : 0;; ++gi) {
    const int sg = gi < 0 ? first_sg : sb_begin + gi * 32;
    if (gi >= 0 && sg >= sb_end) break;
    if (gi >= 0 && sg == first_sg) continue;
    float lbs[QW];
    bool need_s = false;
    float key_s = INFINITY;
    if (sg + lane < sb_end) {
      float lo[LIN], hi[LIN];
#pragma unroll
      for (int c = 0; c < LIN; ++c) {
        lo[c] = __ldg(sv.sbb + (long long)c * sv.nsb_cap + sg + lane);
        hi[c] = __ldg(sv.sbb + (long long)(LIN + c) * sv.nsb_cap + sg + lane);
      }
#pragma unroll
      for (int w = 0; w < QW; ++w) {
        lbs[w] = box_lb<LIN>(lo, hi, q[w]);
        need_s |= !(lbs[w] > worst[w]);
        key_s = fminf(key_s, lbs[w]);
      }
    } else {
#pragma unroll
      for (int w = 0; w < QW; ++w) lbs[w] = INFINITY;
    }
    unsigned todo_s = __ballot_sync(kFull, need_s);
    const unsigned pk_s = (__float_as_uint(key_s) & 0xffffffe0u) | (unsigned)lane;   // (key >= 0: bit order = value order)
    while (todo_s) {
      const int sl = (int)(__reduce_min_sync(kFull, ((todo_s >> lane) & 1u) ? pk_s : 0xffffffffu) & 31u);
      todo_s &= ~(1u << sl);
      const int sb = sg + sl;
      const int blk = sb * 32 + lane;
      float lb[QW];
      bool need = false;
      float key_b = INFINITY;
      if (blk < sv.nblk) {
        float lo[LIN], hi[LIN];
#pragma unroll
        for (int c = 0; c < LIN; ++c) {
          lo[c] = __ldg(sv.bb + (long long)c * sv.nblk_cap + blk);
          hi[c] = __ldg(sv.bb + (long long)(LIN + c) * sv.nblk_cap + blk);
        }
#pragma unroll
        for (int w = 0; w < QW; ++w) {
          lb[w] = box_lb<LIN>(lo, hi, q[w]);
          need |= !(lb[w] > worst[w]);
          key_b = fminf(key_b, lb[w]);
        }
      } else {
#pragma unroll
        for (int w = 0; w < QW; ++w) lb[w] = INFINITY;
      }
      unsigned todo = __ballot_sync(kFull, need);
      const unsigned pk_b = (__float_as_uint(key_b) & 0xffffffe0u) | (unsigned)lane;
      while (todo) {
        const int bl = (int)(__reduce_min_sync(kFull, ((todo >> lane) & 1u) ? pk_b : 0xffffffffu) & 31u);
        todo &= ~(1u << bl);
        // ---- visit block sb*32 + bl: lane-per-node
        const int base = (sb * 32 + bl) * 32;
        const int pos = base + lane;
        const bool valid = pos < sv.n_sorted;
        float nd[LIN];
#pragma unroll
        for (int c = 0; c < LIN; ++c) nd[c] = valid ? __ldg(sv.coords + (long long)c * sv.cap_s + pos) : 0.f;
        float d[QW];
        bool any = false;
#pragma unroll
        for (int w = 0; w < QW; ++w) {
          d[w] = valid ? metric_lin<DIM>(nd, q[w]) : INFINITY;
          any |= !(d[w] > worst[w]);
        }
        if (__any_sync(kFull, any)) {
          const int id = valid ? __ldg(sv.ids + pos) : -1;
          if (DIM == 6) {
            float ang[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) ang[c] = valid ? __ldg(sv.coords + (long long)(3 + c) * sv.cap_s + pos) : 0.f;
            if (!wide) {
#pragma unroll
              for (int w = 0; w < QW; ++w) d[w] = valid ? metric_ang<false>(d[w], ang, q[w]) : INFINITY;
            } else {
#pragma unroll
              for (int w = 0; w < QW; ++w) d[w] = valid ? metric_ang<true>(d[w], ang, q[w]) : INFINITY;
            }
          }
#pragma unroll
          for (int w = 0; w < QW; ++w) {
            unsigned mask = __ballot_sync(kFull, valid && (d[w] < worst[w] || (d[w] == worst[w] && (unsigned)id < (unsigned)worst_id[w])));
            while (mask) {
              const int src = __ffs(mask) - 1;
              mask &= mask - 1;
              const float cd = __shfl_sync(kFull, d[w], src);
              const int ci = __shfl_sync(kFull, id, src);
              if (cd < worst[w] || (cd == worst[w] && (unsigned)ci < (unsigned)worst_id[w])) {
                top[w].insert_keyed(cd, ci, lane);
                const float kd = top[w].kth(k);
                if (kd <= seed[w]) {          // list full and at least as tight as the seed bound
                  worst[w] = kd;
                  worst_id[w] = top[w].kth_id(k);
                }
              }
            }
          }
        }
        // the k-th distances may have shrunk: drop the remaining blocks of this superblock that no longer matter
        need = false;
#pragma unroll
        for (int w = 0; w < QW; ++w) need |= !(lb[w] > worst[w]);
        todo &= __ballot_sync(kFull, need);
      }
      // ... and the remaining superblocks of this group
      need_s = false;
#pragma unroll
      for (int w = 0; w < QW; ++w) need_s |= !(lbs[w] > worst[w]);
      todo_s &= __ballot_sync(kFull, need_s);
    }
  }
#pragma unroll
  for (int w = 0; w < QW; ++w) {
    if (group * QW + w >= nq) break;
    const long long qi = row[w];
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
      const int pos = lane * KPL + s;
      if (pos < k) {
        if (slots_total == 1) {
          store_row_entry(rows, qi, k, pos, top[w].d[s], top[w].id[s]);
        } else {
          const long long o = (qi * slots_total + slot_base + slice) * k + pos;
          out_d[o] = top[w].d[s];
          out_i[o] = top[w].id[s];
        }
      }
    }
  }
}

// radius search over the sorted view: a block whose box lower bound already reaches r2 cannot hold a hit (d2 < r2 is
// strict and box_lb <= d2).  FILL == false counts, FILL == true writes (d2 bits << 32 | original id) keys.
#ifndef SFFG_RADIUS_MIN_BLOCKS
#define SFFG_RADIUS_MIN_BLOCKS 4   // as the k-NN kernel: 6.0 ms per 16 384-query call at N = 1e6 against 6.6 (3) and 7.5 (1)
#endif
template <int DIM, int QW, bool FILL>
__global__ void __launch_bounds__(kThreads, SFFG_RADIUS_MIN_BLOCKS) radius_pruned_kernel(SortedDev sv, const float *__restrict__ queries, long long nq,
                                                                 float r2, int slices, int sb_per_slice, int *counts,
                                                                 const long long *offsets, int *cursor, unsigned long long *keys) {
  constexpr int LIN = DIM == 6 ? 3 : 2;
  const int lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long group = item / slices;
  const int slice = (int)(item - group * slices);
  if (group * QW >= nq) return;
  float q[QW][DIM];
  int cnt[QW];
#pragma unroll
  for (int w = 0; w < QW; ++w) {
    long long qi = group * QW + w;
    if (qi >= nq) qi = nq - 1;
#pragma unroll
    for (int c = 0; c < DIM; ++c) q[w][c] = __ldg(queries + qi * DIM + c);
    cnt[w] = 0;
  }
  bool wide = false;
  if (DIM == 6) {
    const float node_amax = __uint_as_float(__ldg(sv.amax));
#pragma unroll
    for (int w = 0; w < QW; ++w) wide |= wide_needed(q[w], node_amax);
  }
  const unsigned lt = (1u << lane) - 1u;
  const int nsb = (sv.nblk + 31) / 32;
  const int sb_begin = slice * sb_per_slice;
  const int sb_end = min(nsb, sb_begin + sb_per_slice);
  // two levels, as in the k-NN kernel: 32 superblock boxes per step (lane-per-superblock, 32 768 nodes), then the block
  // boxes of the superblocks whose box is within reach, then the nodes of the blocks that are
  for (int sg = sb_begin; sg < sb_end; sg += 32) {
    bool need_s = false;
    if (sg + lane < sb_end) {
      float lo[LIN], hi[LIN];
#pragma unroll
      for (int c = 0; c < LIN; ++c) {
        lo[c] = __ldg(sv.sbb + (long long)c * sv.nsb_cap + sg + lane);
        hi[c] = __ldg(sv.sbb + (long long)(LIN + c) * sv.nsb_cap + sg + lane);
      }
#pragma unroll
      for (int w = 0; w < QW; ++w) need_s |= box_lb<LIN>(lo, hi, q[w]) < r2;
    }
    unsigned todo_s = __ballot_sync(kFull, need_s);
    while (todo_s) {
      const int sb = sg + __ffs(todo_s) - 1;
      todo_s &= todo_s - 1;
      const int blk = sb * 32 + lane;
      bool need = false;
      if (blk < sv.nblk) {
        float lo[LIN], hi[LIN];
  #pragma unroll
        for (int c = 0; c < LIN; ++c) {
          lo[c] = __ldg(sv.bb + (long long)c * sv.nblk_cap + blk);
          hi[c] = __ldg(sv.bb + (long long)(LIN + c) * sv.nblk_cap + blk);
        }
  #pragma unroll
        for (int w = 0; w < QW; ++w) need |= box_lb<LIN>(lo, hi, q[w]) < r2;
      }
      unsigned todo = __ballot_sync(kFull, need);
      while (todo) {
        const int bl = __ffs(todo) - 1;
        todo &= todo - 1;
        const int pos = (sb * 32 + bl) * 32 + lane;
        const bool valid = pos < sv.n_sorted;
        float nd[DIM];
  #pragma unroll
        for (int c = 0; c < DIM; ++c) nd[c] = valid ? __ldg(sv.coords + (long long)c * sv.cap_s + pos) : 0.f;
        const int id = (FILL && valid) ? __ldg(sv.ids + pos) : 0;
  #pragma unroll
        for (int w = 0; w < QW; ++w) {
          const float d = !valid ? INFINITY : (wide ? metric<DIM, true>(nd, q[w]) : metric<DIM, false>(nd, q[w]));
          const bool in = d < r2;
          const unsigned mask = __ballot_sync(kFull, in);
          if (FILL) {
            const long long qi = group * QW + w;
            if (mask && qi < nq) {
              int at = 0;
              if (lane == 0) at = atomicAdd(cursor + qi, __popc(mask));
              at = __shfl_sync(kFull, at, 0);
              const long long o = offsets[qi];   // negative: the device-side scan found the result buffers too small
              if (in && o >= 0) keys[o + at + __popc(mask & lt)] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(unsigned)id;
            }
          } else {
            cnt[w] += __popc(mask);
          }
        }
      }
  }
  }
  if (!FILL && lane == 0) {
#pragma unroll
    for (int w = 0; w < QW; ++w) {
      const long long qi = group * QW + w;
      if (qi < nq && cnt[w]) atomicAdd(counts + qi, cnt[w]);
    }
  }
}

template <int DIM, int QW>
cudaError_t launch_pruned_kpl(const SortedDev &sv, const float *q, int64_t nq, int k, int slices, int sb_per_slice, float *od, int *oi,
                              int slot_base, int slots_total, unsigned grid, const unsigned *perm, const RowDests &rows, cudaStream_t st) {
  if (k <= 32) knn_pruned_kernel<DIM, QW, 1><<<grid, kThreads, 0, st>>>(sv, q, nq, k, slices, sb_per_slice, od, oi, slot_base, slots_total, perm, rows);
  else if (k <= 64) knn_pruned_kernel<DIM, QW, 2><<<grid, kThreads, 0, st>>>(sv, q, nq, k, slices, sb_per_slice, od, oi, slot_base, slots_total, perm, rows);
  else knn_pruned_kernel<DIM, QW, 4><<<grid, kThreads, 0, st>>>(sv, q, nq, k, slices, sb_per_slice, od, oi, slot_base, slots_total, perm, rows);
  return cudaGetLastError();
}

}  // namespace

size_t sorted_build_temp_bytes(int n) {
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned *)nullptr, (unsigned *)nullptr, (const unsigned *)nullptr,
                                  (unsigned *)nullptr, n, 0, 30);
  return tmp + 256;
}

cudaError_t launch_sorted_build(const IndexDev &idx, int n, const SortedBuildBuffers &b, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int lin = idx.dim == 6 ? 3 : 2;
  const int threads = 256, blocks = (n + threads - 1) / threads;
  const int init[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, (int)0x80000000, (int)0x80000000, (int)0x80000000};
  cudaError_t e = cudaMemcpyAsync(b.bounds, init, sizeof init, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) return e;
  bounds_kernel<<<blocks, threads, 0, st>>>(idx.coords, idx.capacity, n, lin, b.bounds);
  morton_kernel<<<blocks, threads, 0, st>>>(idx.coords, idx.capacity, n, lin, b.bounds, b.keys_in, b.vals_in);
  size_t tmp = b.temp_bytes;
  e = cub::DeviceRadixSort::SortPairs(b.temp, tmp, b.keys_in, b.keys_out, b.vals_in, b.vals_out, n, 0, 30, st);
  if (e != cudaSuccess) return e;
  gather_kernel<<<blocks, threads, 0, st>>>(idx.coords, idx.capacity, idx.dim, n, b.vals_out, b.s_coords, b.cap_s, b.s_ids);
  const int nblk = (n + 31) / 32;
  block_box_kernel<<<(nblk * 32 + threads - 1) / threads, threads, 0, st>>>(b.s_coords, b.cap_s, n, lin, b.bb, b.nblk_cap);
  const int nsb = (nblk + 31) / 32;
  superblock_box_kernel<<<(nsb * 32 + threads - 1) / threads, threads, 0, st>>>(b.bb, b.nblk_cap, nblk, lin, b.sbb, b.nsb_cap);
  return cudaGetLastError();
}

PrunedPlan plan_pruned(int64_t nq, const SortedDev &sv, int64_t tail, int sm_count) {
  PrunedPlan p;
  p.qw = nq >= 8192 ? 8 : (nq >= 1024 ? 4 : 1);
  const int64_t groups = (nq + p.qw - 1) / p.qw;
  const int64_t want_warps = (int64_t)sm_count * kWarps * 4;
  const int nsb = (sv.nblk + 31) / 32;
  int slices = 1;
  if (groups < want_warps) {
    slices = (int)std::min<int64_t>((want_warps + groups - 1) / groups, std::max(1, nsb / 4));   // >= 4 superblocks per slice
    if (slices < 1) slices = 1;
    if (slices > 1024) slices = 1024;
  }
  p.sb_per_slice = (nsb + slices - 1) / slices;
  p.slices = (nsb + p.sb_per_slice - 1) / std::max(1, p.sb_per_slice);
  if (p.slices < 1) p.slices = 1;
  // the unsorted tail goes through the exhaustive kernel in pieces of >= 2048 nodes
  p.tail_slices = 0;
  p.tail_len = 0;
  if (tail > 0) {
    int64_t ts = 1;
    if (groups < want_warps) ts = std::min<int64_t>((want_warps + groups - 1) / groups, (tail + 2047) / 2048);
    if (ts < 1) ts = 1;
    int64_t len = (tail + ts - 1) / ts;
    len = (len + 31) / 32 * 32;
    p.tail_len = len;
    p.tail_slices = (int)((tail + len - 1) / len);
  }
  const char *qs = std::getenv("SFFG_KNN_QUERY_SORT");
  p.sort_queries = nq >= 4096 && nq < (int64_t(1) << 31) && !(qs && qs[0] == '0');
  p.sort_temp_bytes = 0;
  if (p.sort_queries) {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned *)nullptr, (unsigned *)nullptr, (const unsigned *)nullptr,
                                    (unsigned *)nullptr, (int)nq, 0, 30);
    p.sort_temp_bytes = tmp + 256;
  }
  return p;
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

size_t pruned_scratch_bytes(const PrunedPlan &p, int64_t nq, int k) {
  const int slots = p.slices + p.tail_slices;
  size_t bytes = slots <= 1 ? 0 : align256((size_t)nq * slots * k * 8);
  if (p.sort_queries) bytes += 4 * align256((size_t)nq * 4) + p.sort_temp_bytes;
  return bytes;
}

cudaError_t launch_knn_pruned(const IndexDev &idx, const SortedDev &sv, const float *d_queries, int64_t nq, int k, const RowDests &out,
                              void *d_scratch, const PrunedPlan &p, cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  const int slots = p.slices + p.tail_slices;
  float *od = nullptr;
  int *oi = nullptr;
  if (slots > 1) {
    od = reinterpret_cast<float *>(d_scratch);
    oi = reinterpret_cast<int *>(od + (size_t)nq * slots * k);
  }
  const int64_t groups = (nq + p.qw - 1) / p.qw;
  const unsigned grid = (unsigned)((groups * p.slices + kWarps - 1) / kWarps);
  cudaError_t e;
  const unsigned *perm = nullptr;
  if (p.sort_queries) {
    // visiting order of the query rows: Morton keys on the node grid, one radix sort of (key, row)
    unsigned char *base = reinterpret_cast<unsigned char *>(d_scratch) + (slots > 1 ? align256((size_t)nq * slots * k * 8) : 0);
    const size_t arr = align256((size_t)nq * 4);
    unsigned *keys_in = reinterpret_cast<unsigned *>(base), *keys_out = reinterpret_cast<unsigned *>(base + arr);
    unsigned *vals_in = reinterpret_cast<unsigned *>(base + 2 * arr), *vals_out = reinterpret_cast<unsigned *>(base + 3 * arr);
    void *tmp = base + 4 * arr;
    size_t tmp_bytes = p.sort_temp_bytes;
    const int lin = idx.dim == 6 ? 3 : 2;
    query_morton_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(d_queries, (int)nq, idx.dim, lin, sv.bounds, keys_in, vals_in);
    e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int)nq, 0, 30, st);
    if (e != cudaSuccess) return e;
    perm = vals_out;
  }
  if (idx.dim == 6) {
    if (p.qw == 8) e = launch_pruned_kpl<6, 8>(sv, d_queries, nq, k, p.slices, p.sb_per_slice, od, oi, 0, slots, grid, perm, out, st);
    else if (p.qw == 4) e = launch_pruned_kpl<6, 4>(sv, d_queries, nq, k, p.slices, p.sb_per_slice, od, oi, 0, slots, grid, perm, out, st);
    else e = launch_pruned_kpl<6, 1>(sv, d_queries, nq, k, p.slices, p.sb_per_slice, od, oi, 0, slots, grid, perm, out, st);
  } else {
    if (p.qw == 8) e = launch_pruned_kpl<2, 8>(sv, d_queries, nq, k, p.slices, p.sb_per_slice, od, oi, 0, slots, grid, perm, out, st);
    else if (p.qw == 4) e = launch_pruned_kpl<2, 4>(sv, d_queries, nq, k, p.slices, p.sb_per_slice, od, oi, 0, slots, grid, perm, out, st);
    else e = launch_pruned_kpl<2, 1>(sv, d_queries, nq, k, p.slices, p.sb_per_slice, od, oi, 0, slots, grid, perm, out, st);
  }
  if (e != cudaSuccess) return e;
  if (p.tail_slices > 0) {
    e = launch_knn_scan_range(idx, d_queries, nq, k, p.qw, p.tail_slices, p.tail_len, sv.n_sorted, od, oi, p.slices, slots, out, st);
    if (e != cudaSuccess) return e;
  }
  if (slots > 1) e = launch_knn_merge(od, oi, nq, k, slots, out, st);
  return e;
}

template <bool FILL>
static cudaError_t launch_radius_pruned_any(const IndexDev &idx, const SortedDev &sv, const float *q, int64_t nq, float r2, int *counts,
                                            const long long *offsets, int *cursor, unsigned long long *keys, int sm_count,
                                            cudaStream_t st) {
  if (nq <= 0) return cudaSuccess;
  // sorted part
  const int qw = nq >= 1024 ? 4 : 1;
  const int64_t groups = (nq + qw - 1) / qw;
  const int64_t want_warps = (int64_t)sm_count * kWarps * 4;
  const int nsb = (sv.nblk + 31) / 32;
  int slices = 1;
  if (groups < want_warps) slices = (int)std::max<int64_t>(1, std::min<int64_t>((want_warps + groups - 1) / groups, std::max(1, nsb / 4)));
  const int sb_per_slice = (nsb + slices - 1) / slices;
  slices = (nsb + sb_per_slice - 1) / std::max(1, sb_per_slice);
  const unsigned grid = (unsigned)((groups * slices + kWarps - 1) / kWarps);
  if (idx.dim == 6) {
    if (qw == 4) radius_pruned_kernel<6, 4, FILL><<<grid, kThreads, 0, st>>>(sv, q, nq, r2, slices, sb_per_slice, counts, offsets, cursor, keys);
    else radius_pruned_kernel<6, 1, FILL><<<grid, kThreads, 0, st>>>(sv, q, nq, r2, slices, sb_per_slice, counts, offsets, cursor, keys);
  } else {
    if (qw == 4) radius_pruned_kernel<2, 4, FILL><<<grid, kThreads, 0, st>>>(sv, q, nq, r2, slices, sb_per_slice, counts, offsets, cursor, keys);
    else radius_pruned_kernel<2, 1, FILL><<<grid, kThreads, 0, st>>>(sv, q, nq, r2, slices, sb_per_slice, counts, offsets, cursor, keys);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // unsorted tail through the exhaustive kernels
  const int64_t tail = idx.n - sv.n_sorted;
  if (tail > 0) {
    KnnPlan tp = plan_knn(nq, tail, sm_count);
    if (FILL) e = launch_radius_fill(idx, q, nq, r2, reinterpret_cast<const int64_t *>(offsets), cursor, keys, tp, st, sv.n_sorted);
    else e = launch_radius_count(idx, q, nq, r2, counts, tp, st, sv.n_sorted);
  }
  return e;
}

cudaError_t launch_radius_count_pruned(const IndexDev &idx, const SortedDev &sv, const float *d_queries, int64_t nq, float r2,
                                       int32_t *d_counts, int sm_count, cudaStream_t st) {
  return launch_radius_pruned_any<false>(idx, sv, d_queries, nq, r2, d_counts, nullptr, nullptr, nullptr, sm_count, st);
}

cudaError_t launch_radius_fill_pruned(const IndexDev &idx, const SortedDev &sv, const float *d_queries, int64_t nq, float r2,
                                      const int64_t *d_offsets, int32_t *d_cursor, unsigned long long *d_keys, int sm_count,
                                      cudaStream_t st) {
  return launch_radius_pruned_any<true>(idx, sv, d_queries, nq, r2, nullptr, reinterpret_cast<const long long *>(d_offsets), d_cursor,
                                        d_keys, sm_count, st);
}

}  // namespace sffg
