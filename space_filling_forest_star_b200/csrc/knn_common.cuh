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

// |wrap(b - a)| of NormalizeAngle<float> (reference src/primitives.h:277-292: the +-2*pi is done in double and
// narrowed).  Exhaustively verified equal for every float |b - a| < 14 (tests/test_metric_wrap.py):
//   |d| >= float(pi)  ->  |(|d| - hi) + lo'|   (first subtraction exact by Sterbenz, second correctly rounded)
//   otherwise         ->  |d|, and min() selects between the two without a branch
__device__ __forceinline__ float wrapped_abs(float qa, float na) {
  const float a = fabsf(__fsub_rn(qa, na));
  const float t = __fadd_rn(__fsub_rn(a, kTwoPiHi), kTwoPiLoNeg);
  return fminf(a, fabsf(t));
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
__device__ __forceinline__ float metric_ang(float r, const float *na, const float *q) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float w = wrapped_abs(q[3 + c], na[c]);
    r = __fadd_rn(r, __fmul_rn(w, w));
  }
  return r;
}
template <int DIM>
__device__ __forceinline__ float metric(const float *nd, const float *q) {
  float r = metric_lin<DIM>(nd, q);
  if (DIM == 6) r = metric_ang(r, nd + 3, q);
  return r;
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
