// knn_kernels.cu -- exact k-NN / radius search over the planner's node sets, sm_100a.
//
// Replaces flann::Index<D6Distance<float>>::knnSearch / radiusSearch as the planner calls them
// (reference src/forest.h:266-267, :317; src/rrt.h:143,:166,:228) with an EXACT brute-force scan of the intended
// metric (src/primitives.h:404-438 with `+=`): d2 = dx^2+dy^2+dz^2 + wrap(dyaw)^2 + wrap(dpitch)^2 + wrap(droll)^2
// evaluated in float in exactly that order, no FMA contraction, so distances are bit-identical to FLANN's
// LinearIndex with the fixed functor and neighbour ids are identical including ties (lower id first).
//
// A 2..6-D metric is not a dense contraction, so this is CUDA-core FP32 work: one warp owns QW queries (held in
// registers, warp-uniform) and streams the node set 32 nodes per step (lane-per-node, coalesced SoA loads).  The
// running top-k of every query is a sorted list spread over the lanes (KPL entries per lane); a candidate that beats
// the current k-th distance is inserted with warp shuffles.  Small query batches are split into node slices across
// warps and merged by a second kernel.
#include <cfloat>

#include "../../include/sffg.h"
#include "knn_common.cuh"
#include "knn_kernels.cuh"

namespace sffg {
namespace {

using namespace knn;

// One warp = QW queries x one node slice.
//   item = blockIdx.x * kWarps + warp;  group = item / slices;  slice = item % slices
//   out_d / out_i: [nq][slices][k] when slices > 1 (partials), else the final [nq][k]
// The node stream is software-pipelined: the next 32-node block is loaded into registers while the current one is
// evaluated against the QW queries (QW independent dependency chains give the ILP), and one warp vote decides whether
// any of the QW queries has a candidate in this block before the per-query ballots are taken.
template <int DIM, int QW, int KPL>
__global__ void __launch_bounds__(kThreads) knn_scan_kernel(IndexDev idx, const float *__restrict__ queries, long long nq,
                                                            int k, int slices, long long slice_len, float *out_d,
                                                            int *out_i, long long first, int slot_base, int slots_total,
                                                            RowDests rows) {
  const int lane = threadIdx.x & 31;
  const long long item = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  const long long group = item / slices;
  const int slice = (int)(item - group * slices);
  if (group * QW >= nq) return;
  float q[QW][DIM];
  TopK<KPL> top[QW];
  float worst[QW];
#pragma unroll
  for (int w = 0; w < QW; ++w) {
    long long qi = group * QW + w;
    if (qi >= nq) qi = nq - 1;   // duplicate work for the ragged tail, never written
#pragma unroll
    for (int c = 0; c < DIM; ++c) q[w][c] = __ldg(queries + qi * DIM + c);
    top[w].init();
    worst[w] = INFINITY;
  }
  // un-normalised angles (the reference lets them drift): one warp-uniform decision selects the exact wide wrap
  bool wide = false;
  if (DIM == 6) {
    const float node_amax = __uint_as_float(__ldg(idx.amax));
#pragma unroll
    for (int w = 0; w < QW; ++w) wide |= wide_needed(q[w], node_amax);
  }
  // the scan covers nodes [first, idx.n): the whole index, or the not-yet-sorted tail behind a sorted view
  const long long begin = first + (long long)slice * slice_len;
  long long end = begin + slice_len;
  if (end > idx.n) end = idx.n;
  constexpr int LIN = DIM == 6 ? 3 : 2;   // coordinates streamed for every block; the 3 angles only on demand
  // 32-bit bookkeeping in the hot loop (node ids fit in 31 bits).  Two 32-node blocks per iteration, both prefetched
  // one iteration (2 blocks) ahead; the index capacity is padded by 96 nodes so the prefetch never needs a bounds branch.
  const int nvalid = end > begin ? (int)(end - begin) : 0;
  const int nsteps = (nvalid + 63) / 64;
  const float *ptr[LIN];
#pragma unroll
  for (int c = 0; c < LIN; ++c) ptr[c] = idx.coords + (long long)c * idx.capacity + begin + lane;
  const float *ang_base = idx.coords + 3LL * idx.capacity + lane;
  float cur[2][LIN], nxt[2][LIN];
#pragma unroll
  for (int c = 0; c < LIN; ++c) {
    cur[0][c] = __ldg(ptr[c]);
    cur[1][c] = __ldg(ptr[c] + 32);
  }
  int b = (int)begin;
  int left = nvalid;   // valid nodes from block `b` on
  for (int it = 0; it < nsteps; ++it) {
#pragma unroll
    for (int c = 0; c < LIN; ++c) {
      ptr[c] += 64;
      nxt[0][c] = __ldg(ptr[c]);
      nxt[1][c] = __ldg(ptr[c] + 32);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h, b += 32, left -= 32) {
      float d[QW];
      bool any = false;
      if (left >= 32) {
#pragma unroll
        for (int w = 0; w < QW; ++w) {
          d[w] = metric_lin<DIM>(cur[h], q[w]);
          any |= d[w] < worst[w];
        }
      } else {
        const bool valid = lane < left;
#pragma unroll
        for (int w = 0; w < QW; ++w) {
          d[w] = valid ? metric_lin<DIM>(cur[h], q[w]) : INFINITY;
          any |= d[w] < worst[w];
        }
      }
      if (__any_sync(kFull, any)) {
        if (DIM == 6) {
          const bool valid = lane < left;
          float ang[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) ang[c] = __ldg(ang_base + (long long)c * idx.capacity + b);
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
          unsigned mask = __ballot_sync(kFull, d[w] < worst[w]);
          while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const float cd = __shfl_sync(kFull, d[w], src);
            if (cd < worst[w]) {
              top[w].insert(cd, b + src, lane);
              worst[w] = top[w].kth(k);
            }
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < LIN; ++c) {
      cur[0][c] = nxt[0][c];
      cur[1][c] = nxt[1][c];
    }
  }
#pragma unroll
  for (int w = 0; w < QW; ++w) {
    const long long qi = group * QW + w;
    if (qi >= nq) break;
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

// merges the partial lists of one query (one warp per query) with the (d2, id)-keyed insert: lists may come from node
// slices, from the spatially sorted view and from the unsorted tail, in any id order.
template <int KPL>
__global__ void __launch_bounds__(kThreads) knn_merge_kernel(const float *__restrict__ part_d, const int *__restrict__ part_i,
                                                             long long nq, int k, int slices, RowDests rows) {
  const int lane = threadIdx.x & 31;
  const long long qi = (long long)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (qi >= nq) return;
  TopK<KPL> top;
  top.init();
  float worst = INFINITY;
  int worst_id = -1;
  const long long total = (long long)slices * k;
  const float *pd = part_d + qi * total;
  const int *pi = part_i + qi * total;
  for (long long b = 0; b < total; b += 32) {
    const long long i = b + lane;
    const float d = i < total ? pd[i] : INFINITY;
    const int id = i < total ? pi[i] : -1;
    unsigned mask = __ballot_sync(kFull, id >= 0 && (d < worst || (d == worst && (unsigned)id < (unsigned)worst_id)));
    while (mask) {
      const int src = __ffs(mask) - 1;
      mask &= mask - 1;
      const float cd = __shfl_sync(kFull, d, src);
      const int ci = __shfl_sync(kFull, id, src);
      if (cd < worst || (cd == worst && (unsigned)ci < (unsigned)worst_id)) {
        top.insert_keyed(cd, ci, lane);
        worst = top.kth(k);
        worst_id = top.kth_id(k);
      }
    }
  }
#pragma unroll
  for (int s = 0; s < KPL; ++s) {
    const int pos = lane * KPL + s;
    if (pos < k) store_row_entry(rows, qi, k, pos, top.d[s], top.id[s]);
  }
}

// ---- radius --------------------------------------------------------------------------------------------
// FILL == false: counts[q] += #nodes with d2 < r2.   FILL == true: keys[offsets[q] + cursor[q]++] = (d2,id)
template <int DIM, int QW, bool FILL>
__global__ void __launch_bounds__(kThreads) radius_scan_kernel(IndexDev idx, const float *__restrict__ queries, long long nq,
                                                               float r2, int slices, long long slice_len, int *counts,
                                                               const long long *offsets, int *cursor,
                                                               unsigned long long *keys, long long first) {
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
    const float node_amax = __uint_as_float(__ldg(idx.amax));
#pragma unroll
    for (int w = 0; w < QW; ++w) wide |= wide_needed(q[w], node_amax);
  }
  const long long begin = first + (long long)slice * slice_len;
  long long end = begin + slice_len;
  if (end > idx.n) end = idx.n;
  const unsigned lt = (1u << lane) - 1u;
  for (long long b = begin; b < end; b += 32) {
    const long long i = b + lane;
    float nd[DIM];
    const bool valid = i < end;
#pragma unroll
    for (int c = 0; c < DIM; ++c) nd[c] = valid ? __ldg(idx.coords + (long long)c * idx.capacity + i) : 0.f;
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
          if (in && o >= 0)
            keys[o + at + __popc(mask & lt)] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(unsigned)i;
        }
      } else {
        cnt[w] += __popc(mask);
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

// Ascending sort of one row per block with a comparator network whose exchanges all point the same way
// (bitonic "flip + disperse" form), so positions >= len behave as +inf padding without being stored.
constexpr int kSortSmem = 4096;
__device__ __forceinline__ void cmpswap(unsigned long long *a, long long i, long long l) {
  const unsigned long long x = a[i], y = a[l];
  if (x > y) { a[i] = y; a[l] = x; }
}
__global__ void __launch_bounds__(kThreads) radius_sort_kernel(unsigned long long *keys, const long long *offsets,
                                                               const int *counts, int *ids, float *d2) {
  __shared__ unsigned long long sk[kSortSmem];
  const long long row = blockIdx.x;
  const long long len = counts[row];
  if (len == 0 || offsets[row] < 0) return;
  unsigned long long *g = keys + offsets[row];
  unsigned long long *a = g;
  const bool in_smem = len <= kSortSmem;
  if (in_smem) {
    for (long long i = threadIdx.x; i < len; i += blockDim.x) sk[i] = g[i];
    a = sk;
  }
  __syncthreads();
  long long pow2 = 1;
  while (pow2 < len) pow2 <<= 1;
  for (long long k = 2; k <= pow2; k <<= 1) {
    for (long long i = threadIdx.x; i < len; i += blockDim.x) {
      const long long l = i ^ (k - 1);
      if (l > i && l < len) cmpswap(a, i, l);
    }
    __syncthreads();
    for (long long j = k >> 2; j > 0; j >>= 1) {
      for (long long i = threadIdx.x; i < len; i += blockDim.x) {
        const long long l = i ^ j;
        if (l > i && l < len) cmpswap(a, i, l);
      }
      __syncthreads();
    }
  }
  for (long long i = threadIdx.x; i < len; i += blockDim.x) {
    const unsigned long long key = a[i];
    ids[offsets[row] + i] = (int)(unsigned)(key & 0xffffffffull);
    d2[offsets[row] + i] = __uint_as_float((unsigned)(key >> 32));
  }
}

// Planner-sized radius search in ONE kernel (small index, few queries): a block per query scans every node, gathers the
// hits in shared memory, sorts them by (d2, id) and stores the row at a position it reserves in the packed output with one
// atomic; the host re-orders the rows by query while it copies them out of the pinned staging area.  counts_out is exact
// even when a row is too long for the shared buffer or the output is full -- those cases raise `overflow` and the caller
// repeats the search on the count / scan / fill / sort path.
constexpr int kFusedRowCap = 2048;
template <int DIM>
__global__ void __launch_bounds__(kThreads) radius_fused_kernel(IndexDev idx, const float *__restrict__ queries, float r2, long long out_cap,
                                                                int *counts_out, long long *rowoff_out, int *ids_out, float *d2_out,
                                                                unsigned long long *total, int *overflow) {
  __shared__ unsigned long long sk[kFusedRowCap];
  __shared__ int s_cnt;
  __shared__ long long s_off;
  const long long qi = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_cnt = 0;
  float q[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) q[c] = queries[qi * DIM + c];
  bool wide = false;
  if (DIM == 6) wide = wide_needed(q, __uint_as_float(__ldg(idx.amax)));
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  for (long long b = (long long)warp * 32; b < idx.n; b += kThreads) {
    const long long i = b + lane;
    const bool valid = i < idx.n;
    float nd[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) nd[c] = valid ? __ldg(idx.coords + (long long)c * idx.capacity + i) : 0.f;
    const float d = !valid ? INFINITY : (wide ? metric<DIM, true>(nd, q) : metric<DIM, false>(nd, q));
    const bool in = d < r2;
    const unsigned mask = __ballot_sync(kFull, in);
    if (mask) {
      int at = 0;
      if (lane == 0) at = atomicAdd(&s_cnt, __popc(mask));
      at = __shfl_sync(kFull, at, 0);
      const int pos = at + __popc(mask & lt);
      if (in && pos < kFusedRowCap) sk[pos] = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(unsigned)i;
    }
  }
  __syncthreads();
  const int len = s_cnt;
  if (len > kFusedRowCap) {
    if (threadIdx.x == 0) {
      counts_out[qi] = len;
      rowoff_out[qi] = -1;
      *reinterpret_cast<volatile int *>(overflow) = 1;
    }
    return;
  }
  int pow2 = 1;
  while (pow2 < len) pow2 <<= 1;
  for (int k = 2; k <= pow2; k <<= 1) {
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
      const int l = i ^ (k - 1);
      if (l > i && l < len) cmpswap(sk, i, l);
    }
    __syncthreads();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < len; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i && l < len) cmpswap(sk, i, l);
      }
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) s_off = len ? (long long)atomicAdd(total, (unsigned long long)len) : 0;
  __syncthreads();
  const long long off = s_off;
  if (off + len > out_cap) {
    if (threadIdx.x == 0) {
      counts_out[qi] = len;
      rowoff_out[qi] = -1;
      *reinterpret_cast<volatile int *>(overflow) = 1;
    }
    return;
  }
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    const unsigned long long key = sk[i];
    ids_out[off + i] = (int)(unsigned)(key & 0xffffffffull);
    d2_out[off + i] = __uint_as_float((unsigned)(key >> 32));
  }
  if (threadIdx.x == 0) {
    counts_out[qi] = len;
    rowoff_out[qi] = off;
  }
}

// Exclusive scan of the per-query counts on the device (one block; planner-sized calls), so that count, fill and sort run
// back to back without a host round trip.  The total goes to `total_out` (pinned mapped memory the host reads after its one
// synchronisation); when it exceeds `capacity` every offset is set to -1, which turns the fill and sort kernels into no-ops.
constexpr int kScanThreads = 1024;
__global__ void __launch_bounds__(kScanThreads) radius_offsets_kernel(const int *__restrict__ counts, long long nq, long long capacity,
                                                                      long long *offsets, long long *total_out) {
  __shared__ long long part[kScanThreads];
  const long long per = (nq + kScanThreads - 1) / kScanThreads;
  const long long b = (long long)threadIdx.x * per, e = b + per < nq ? b + per : nq;
  long long s = 0;
  for (long long i = b; i < e; ++i) s += counts[i];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int d = 1; d < kScanThreads; d <<= 1) {
    const long long v = (int)threadIdx.x >= d ? part[threadIdx.x - d] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  const long long total = part[kScanThreads - 1];
  const bool fits = total <= capacity;
  long long run = part[threadIdx.x] - s;
  for (long long i = b; i < e; ++i) {
    offsets[i] = fits ? run : -1;
    run += counts[i];
  }
  if (threadIdx.x == 0) *total_out = total;
}

__global__ void index_append_kernel(float *coords, long long capacity, int dim, long long at, const float *__restrict__ pts,
                                    long long n, unsigned *amax) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned mag = 0;   // float bits of |angle|: for non-negative floats the unsigned order is the float order, NaN sorts above inf
  if (t < n * dim) {
    const long long i = t / dim;
    const int c = (int)(t - i * dim);
    const float v = pts[t];
    coords[(long long)c * capacity + at + i] = v;
    if (c >= 3) mag = __float_as_uint(fabsf(v));
  }
  if (dim == 6) {
    mag = __reduce_max_sync(kFull, mag);
    if ((threadIdx.x & 31) == 0 && mag > *(volatile unsigned *)amax) atomicMax(amax, mag);
  }
}

template <int DIM, int QW>
cudaError_t launch_knn_kpl(const IndexDev &idx, const float *q, int64_t nq, int k, int slices, int64_t slice_len,
                           float *od, int *oi, unsigned grid, cudaStream_t st, long long first, int slot_base, int slots_total,
                           const RowDests &rows) {
  if (k <= 32) knn_scan_kernel<DIM, QW, 1><<<grid, kThreads, 0, st>>>(idx, q, nq, k, slices, slice_len, od, oi, first, slot_base, slots_total, rows);
  else if (k <= 64) knn_scan_kernel<DIM, QW, 2><<<grid, kThreads, 0, st>>>(idx, q, nq, k, slices, slice_len, od, oi, first, slot_base, slots_total, rows);
  else knn_scan_kernel<DIM, QW, 4><<<grid, kThreads, 0, st>>>(idx, q, nq, k, slices, slice_len, od, oi, first, slot_base, slots_total, rows);
  return cudaGetLastError();
}

}  // namespace

KnnPlan plan_knn(int64_t nq, int64_t n, int sm_count) {
  KnnPlan p;
  p.qw = nq >= 8192 ? 8 : (nq >= 1024 ? 4 : 1);
  p.groups = (nq + p.qw - 1) / p.qw;
  const int64_t want_warps = (int64_t)sm_count * kWarps * 4;
  int64_t slices = 1;
  if (p.groups < want_warps) {
    slices = (want_warps + p.groups - 1) / p.groups;
    const int64_t max_slices = (n + 2047) / 2048;   // at least 2048 nodes per slice
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    if (slices > 4096) slices = 4096;
  }
  int64_t len = (n + slices - 1) / slices;
  len = (len + 31) / 32 * 32;
  if (len < 32) len = 32;
  p.slices = (int)((n + len - 1) / len);
  if (p.slices < 1) p.slices = 1;
  p.slice_len = len;
  return p;
}

size_t knn_scratch_bytes(const KnnPlan &p, int64_t nq, int k) {
  if (p.slices <= 1) return 0;
  return (size_t)nq * p.slices * k * 8;
}

// brute-force scan of nodes [first, idx.n) in `slices` pieces; partial list of slice s goes to slot slot_base + s of
// slots_total (slots_total == 1: od / oi are the final [nq][k] outputs)
cudaError_t launch_knn_scan_range(const IndexDev &idx, const float *d_queries, int64_t nq, int k, int qw, int slices,
                                  int64_t slice_len, int64_t first, float *od, int *oi, int slot_base, int slots_total,
                                  const RowDests &rows, cudaStream_t stream) {
  if (nq <= 0 || slices <= 0) return cudaSuccess;
  const int64_t groups = (nq + qw - 1) / qw;
  const int64_t items = groups * slices;
  const unsigned grid = (unsigned)((items + kWarps - 1) / kWarps);
  if (idx.dim == 6) {
    if (qw == 8) return launch_knn_kpl<6, 8>(idx, d_queries, nq, k, slices, slice_len, od, oi, grid, stream, first, slot_base, slots_total, rows);
    if (qw == 4) return launch_knn_kpl<6, 4>(idx, d_queries, nq, k, slices, slice_len, od, oi, grid, stream, first, slot_base, slots_total, rows);
    return launch_knn_kpl<6, 1>(idx, d_queries, nq, k, slices, slice_len, od, oi, grid, stream, first, slot_base, slots_total, rows);
  }
  if (qw == 8) return launch_knn_kpl<2, 8>(idx, d_queries, nq, k, slices, slice_len, od, oi, grid, stream, first, slot_base, slots_total, rows);
  if (qw == 4) return launch_knn_kpl<2, 4>(idx, d_queries, nq, k, slices, slice_len, od, oi, grid, stream, first, slot_base, slots_total, rows);
  return launch_knn_kpl<2, 1>(idx, d_queries, nq, k, slices, slice_len, od, oi, grid, stream, first, slot_base, slots_total, rows);
}

cudaError_t launch_knn_merge(const float *part_d, const int *part_i, int64_t nq, int k, int slots, const RowDests &rows,
                             cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  const unsigned mg = (unsigned)((nq + kWarps - 1) / kWarps);
  if (k <= 32) knn_merge_kernel<1><<<mg, kThreads, 0, stream>>>(part_d, part_i, nq, k, slots, rows);
  else if (k <= 64) knn_merge_kernel<2><<<mg, kThreads, 0, stream>>>(part_d, part_i, nq, k, slots, rows);
  else knn_merge_kernel<4><<<mg, kThreads, 0, stream>>>(part_d, part_i, nq, k, slots, rows);
  return cudaGetLastError();
}

cudaError_t launch_knn(const IndexDev &idx, const float *d_queries, int64_t nq, int k, const RowDests &out,
                       void *d_scratch, const KnnPlan &plan, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  float *od = nullptr;
  int *oi = nullptr;
  if (plan.slices > 1) {
    od = reinterpret_cast<float *>(d_scratch);
    oi = reinterpret_cast<int *>(od + (size_t)nq * plan.slices * k);
  }
  cudaError_t e = launch_knn_scan_range(idx, d_queries, nq, k, plan.qw, plan.slices, plan.slice_len, 0, od, oi, 0, plan.slices, out, stream);
  if (e != cudaSuccess) return e;
  if (plan.slices > 1) e = launch_knn_merge(od, oi, nq, k, plan.slices, out, stream);
  return e;
}

template <bool FILL>
static cudaError_t launch_radius_any(const IndexDev &idx, const float *q, int64_t nq, float r2, int *counts,
                                     const long long *offsets, int *cursor, unsigned long long *keys, const KnnPlan &plan_in,
                                     cudaStream_t st, long long first) {
  if (nq <= 0) return cudaSuccess;
  KnnPlan plan = plan_in;
  if (plan.qw == 8) {   // the radius kernels are instantiated for 1 and 4 queries per warp
    plan.qw = 4;
    plan.groups = (nq + 3) / 4;
  }
  const int64_t items = plan.groups * plan.slices;
  const unsigned grid = (unsigned)((items + kWarps - 1) / kWarps);
  if (idx.dim == 6) {
    if (plan.qw == 4) radius_scan_kernel<6, 4, FILL><<<grid, kThreads, 0, st>>>(idx, q, nq, r2, plan.slices, plan.slice_len, counts, offsets, cursor, keys, first);
    else radius_scan_kernel<6, 1, FILL><<<grid, kThreads, 0, st>>>(idx, q, nq, r2, plan.slices, plan.slice_len, counts, offsets, cursor, keys, first);
  } else {
    if (plan.qw == 4) radius_scan_kernel<2, 4, FILL><<<grid, kThreads, 0, st>>>(idx, q, nq, r2, plan.slices, plan.slice_len, counts, offsets, cursor, keys, first);
    else radius_scan_kernel<2, 1, FILL><<<grid, kThreads, 0, st>>>(idx, q, nq, r2, plan.slices, plan.slice_len, counts, offsets, cursor, keys, first);
  }
  return cudaGetLastError();
}

cudaError_t launch_radius_count(const IndexDev &idx, const float *d_queries, int64_t nq, float r2, int32_t *d_counts,
                                const KnnPlan &plan, cudaStream_t stream, int64_t first) {
  return launch_radius_any<false>(idx, d_queries, nq, r2, d_counts, nullptr, nullptr, nullptr, plan, stream, first);
}

cudaError_t launch_radius_fill(const IndexDev &idx, const float *d_queries, int64_t nq, float r2, const int64_t *d_offsets,
                               int32_t *d_cursor, unsigned long long *d_keys, const KnnPlan &plan, cudaStream_t stream,
                               int64_t first) {
  return launch_radius_any<true>(idx, d_queries, nq, r2, nullptr, reinterpret_cast<const long long *>(d_offsets), d_cursor,
                                 d_keys, plan, stream, first);
}

cudaError_t launch_radius_sort(unsigned long long *d_keys, const int64_t *d_offsets, const int32_t *d_counts, int64_t nq,
                               int32_t *d_ids, float *d_d2, cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  radius_sort_kernel<<<(unsigned)nq, kThreads, 0, stream>>>(d_keys, reinterpret_cast<const long long *>(d_offsets), d_counts,
                                                            d_ids, d_d2);
  return cudaGetLastError();
}

cudaError_t launch_radius_fused(const IndexDev &idx, const float *queries, int64_t nq, float r2, int64_t out_cap, int32_t *counts_out,
                                int64_t *rowoff_out, int32_t *ids_out, float *d2_out, unsigned long long *d_total, int *overflow,
                                cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  if (idx.dim == 6)
    radius_fused_kernel<6><<<(unsigned)nq, kThreads, 0, stream>>>(idx, queries, r2, (long long)out_cap, counts_out,
                                                                  reinterpret_cast<long long *>(rowoff_out), ids_out, d2_out, d_total, overflow);
  else
    radius_fused_kernel<2><<<(unsigned)nq, kThreads, 0, stream>>>(idx, queries, r2, (long long)out_cap, counts_out,
                                                                  reinterpret_cast<long long *>(rowoff_out), ids_out, d2_out, d_total, overflow);
  return cudaGetLastError();
}

cudaError_t launch_radius_offsets(const int32_t *d_counts, int64_t nq, int64_t capacity, int64_t *d_offsets, int64_t *total_out,
                                  cudaStream_t stream) {
  if (nq <= 0) return cudaSuccess;
  radius_offsets_kernel<<<1, kScanThreads, 0, stream>>>(d_counts, (long long)nq, (long long)capacity,
                                                        reinterpret_cast<long long *>(d_offsets), reinterpret_cast<long long *>(total_out));
  return cudaGetLastError();
}

cudaError_t launch_index_append(float *d_coords, int64_t capacity, int dim, int64_t at, const float *d_pts, int64_t n,
                                unsigned *d_amax, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const long long total = (long long)n * dim;
  index_append_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(d_coords, capacity, dim, at, d_pts, n, d_amax);
  return cudaGetLastError();
}

}  // namespace sffg
