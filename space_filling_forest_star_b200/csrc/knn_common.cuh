// knn_common.cuh -- device helpers shared by the neighbour-search kernels (knn_kernels.cu, knn_pruned.cu)
#pragma once
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>

namespace sffg {
namespace knn {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr float kTwoPiHi = 6.28318548202514648f;      // float(2*pi)
constexpr float kTwoPiLoNeg = 1.74845553146951715e-07f;  // float(2*pi) - 2*pi, nearest float

constexpr double kTwoPiD = 6.283185307179586476925286766559;   // 2 * M_PI as the reference's double arithmetic sees it
constexpr float kWideFrom = 12.5f;   // the FP32 form below is exact for |b - a| < 14 (needs |b - a| <= 2 * float(2*pi) for Sterbenz)

// |wrap(b - a)| of NormalizeAngle<float> (reference src/primitives.h:277-292: ONE +-2*pi, done in double and narrowed;
// stored angles are never re-normalised, src/primitives.h:237-250, so |b - a| is unbounded).
//   |d| <  kWideFrom: FP32 only, verified equal for every float in [pi, 14) by tests/test_metric_wrap.py:
//     |d| >= float(pi)  ->  |(|d| - hi) + lo'|   (first subtraction exact by Sterbenz, second correctly rounded)
//     otherwise         ->  |d|, and min() selects between the two without a branch
//   WIDE (a warp-uniform decision of the caller, see wide_needed()): differences of kWideFrom and more take the
//     reference's own float -> double -> float route, so the result is exact for every float.
template <bool WIDE>
__device__ __forceinline__ float wrapped_abs(float qa, float na) {
  const float a = fabsf(__fsub_rn(qa, na));
  const float t = __fadd_rn(__fsub_rn(a, kTwoPiHi), kTwoPiLoNeg);
  float r = fminf(a, fabsf(t));
  if (WIDE) {
    if (!(a < kWideFrom)) r = fabsf(__double2float_rn(__dsub_rn((double)a, kTwoPiD)));
  }
  return r;
}

// Can |query angle - node angle| reach kWideFrom?  node_amax = largest |angle| stored in the index (kept by the append
// kernel), rounding of the sum and of the difference are both monotone, so fl(|qa| + node_amax) < kWideFrom proves
// fl(|qa - na|) < kWideFrom for every node.  NaN anywhere selects the wide path (which propagates it like the reference).
__device__ __forceinline__ bool wide_needed(const float *q, float node_amax) {
  const float qm = fmaxf(fmaxf(fabsf(q[3]), fabsf(q[4])), fabsf(q[5]));
  return !(__fadd_rn(qm, node_amax) < kWideFrom) || q[3] != q[3] || q[4] != q[4] || q[5] != q[5];
}

// upper bound of the angular part of the metric over every node of the index (seeds of the pruned search): each wrapped
// difference is at most max(pi, A - 2*pi) with A >= |qa| + |na|.  The caller adds a relative margin for the roundings.
__device__ __forceinline__ float angular_part_ub(const float *q, float node_amax) {
  const float qm = fmaxf(fmaxf(fabsf(q[3]), fabsf(q[4])), fabsf(q[5]));
  const float A = __fadd_ru(qm, node_amax);
  const float m = fmaxf(3.1415928f, __fsub_ru(A, 6.28f));
  return __fmul_ru(__fmul_ru(m, m), 3.0001f);
}

// translational part (all of the metric for DIM == 2): ((dx^2 + dy^2) + dz^2), float, unfused
template <int DIM>
__device__ __forceinline__ float metric_lin(const float *nd, const float *q) {
  float d = __fsub_rn(nd[0], q[0]);
  float r = __fmul_rn(d, d);
  d = __fsub_rn(nd[1], q[1]);
  r = __fadd_rn(r, __fmul_rn(d, d));
  if (DIM == 6) {
    d = __fsub_rn(nd[2], q[2]);
    r = __fadd_rn(r, __fmul_rn(d, d));
  }
  return r;
}
// angular part continues the same accumulator: (((r + wy^2) + wp^2) + wr^2).  Every term is >= 0 and round-to-nearest
// addition is monotone, so metric_lin() is a lower bound of the full distance: a 32-node block whose translational
// parts all reach the current k-th distance cannot contain a candidate and its angles are never loaded.
template <bool WIDE = false>
__device__ __forceinline__ float metric_ang(float r, const float *na, const float *q) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float w = wrapped_abs<WIDE>(q[3 + c], na[c]);
    r = __fadd_rn(r, __fmul_rn(w, w));
  }
  return r;
}
template <int DIM, bool WIDE = false>
__device__ __forceinline__ float metric(const float *nd, const float *q) {
  float r = metric_lin<DIM>(nd, q);
  if (DIM == 6) r = metric_ang<WIDE>(r, nd + 3, q);
  return r;
}

// final row of query qi (position pos < k) -> every destination
template <class Dests>
__device__ __forceinline__ void store_row_entry(const Dests &out, long long qi, int k, int pos, float d, int id) {
  const long long o = qi * k + pos;
#pragma unroll 1
  for (int r = 0; r < out.n; ++r) {
    out.d2[r][o] = d;
    out.ids[r][o] = id;
  }
}

// sorted (ascending) list of 32*KPL entries spread over the warp: position j lives in lane j / KPL, slot j % KPL
template <int KPL>
struct TopK {
  float d[KPL];
  int id[KPL];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int s = 0; s < KPL; ++s) { d[s] = INFINITY; id[s] = -1; }
  }
  // value at list position k-1, broadcast
  __device__ __forceinline__ float kth(int k) const {
    const int pos = k - 1, lane = pos / KPL, slot = pos % KPL;
    float v = d[0];
#pragma unroll
    for (int s = 1; s < KPL; ++s) if (slot == s) v = d[s];
    return __shfl_sync(kFull, v, lane);
  }
  // insert (cd, ci) AFTER all entries with distance <= cd (candidates arrive in ascending id order, so this is the
  // (d2, id) order of FLANN's KNNSimpleResultSet, result_set.h:151-171).  All lanes call with identical arguments.
  __device__ __forceinline__ void insert(float cd, int ci, int lane) {
    float upd = __shfl_up_sync(kFull, d[KPL - 1], 1);
    int upi = __shfl_up_sync(kFull, id[KPL - 1], 1);
    if (lane == 0) upd = -INFINITY;
#pragma unroll
    for (int s = KPL - 1; s >= 0; --s) {
      const float pd = s > 0 ? d[s - 1] : upd;
      const int pi = s > 0 ? id[s - 1] : upi;
      if (d[s] > cd) {
        const bool shift = pd > cd;
        d[s] = shift ? pd : cd;
        id[s] = shift ? pi : ci;
      }
    }
  }
  // id at list position k-1, broadcast (needed when candidates do not arrive in id order)
  __device__ __forceinline__ int kth_id(int k) const {
    const int pos = k - 1, lane = pos / KPL, slot = pos % KPL;
    int v = id[0];
#pragma unroll
    for (int s = 1; s < KPL; ++s) if (slot == s) v = id[s];
    return __shfl_sync(kFull, v, lane);
  }
  // (d2, id)-ordered insert for candidates that arrive in ARBITRARY id order (spatially sorted scan, merges): the entry
  // goes before every entry with a larger distance or the same distance and a larger id.  Empty slots hold (+inf, -1)
  // and compare as larger than anything through the unsigned id.
  __device__ __forceinline__ void insert_keyed(float cd, int ci, int lane) {
    float upd = __shfl_up_sync(kFull, d[KPL - 1], 1);
    int upi = __shfl_up_sync(kFull, id[KPL - 1], 1);
    if (lane == 0) { upd = -INFINITY; upi = 0; }
#pragma unroll
    for (int s = KPL - 1; s >= 0; --s) {
      const float pd = s > 0 ? d[s - 1] : upd;
      const int pi = s > 0 ? id[s - 1] : upi;
      const bool mine_after = d[s] > cd || (d[s] == cd && (unsigned)id[s] > (unsigned)ci);
      if (mine_after) {
        const bool prev_after = pd > cd || (pd == cd && (unsigned)pi > (unsigned)ci);
        d[s] = prev_after ? pd : cd;
        id[s] = prev_after ? pi : ci;
      }
    }
  }
};


}  // namespace knn
}  // namespace sffg
