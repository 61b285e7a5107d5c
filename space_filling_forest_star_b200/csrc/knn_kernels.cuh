// knn_kernels.cuh -- launch wrappers of the exact neighbour-search kernels (implemented in knn_kernels.cu)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace sffg {

// Device view of an index: struct-of-arrays, coordinate c of node i at coords[c * capacity + i]
struct IndexDev {
  const float *coords;
  int64_t capacity;
  int64_t n;
  int dim;   // 2 or 6
  const unsigned *amax;   // one word: float bits of the largest |angle| ever appended (dim 6), kept by the append kernel
};

// Destinations of the final result rows: the caller's buffers, or (multi-GPU) the same row range in every rank's gathered
// buffer, mapped through CUDA IPC -- the kernel that finishes a row stores it to all of them, so the row exchange rides on
// the search's own stores over NVLink instead of a collective after it.
constexpr int kMaxRowDests = 8;
struct RowDests {
  float *d2[kMaxRowDests];
  int32_t *ids[kMaxRowDests];
  int n;
};
inline RowDests single_dest(int32_t *ids, float *d2) {
  RowDests r{};
  r.n = 1;
  r.ids[0] = ids;
  r.d2[0] = d2;
  return r;
}

struct KnnPlan {
  int qw;        // queries per warp
  int slices;    // node slices per query group
  int64_t slice_len;
  int64_t groups;
};
KnnPlan plan_knn(int64_t nq, int64_t n, int sm_count);

// partial scratch: slices > 1 needs nq * slices * k (float,int) pairs = 8 bytes each
size_t knn_scratch_bytes(const KnnPlan &p, int64_t nq, int k);

cudaError_t launch_knn(const IndexDev &idx, const float *d_queries, int64_t nq, int k, const RowDests &out,
                       void *d_scratch, const KnnPlan &plan, cudaStream_t stream);

// od / oi: partial lists [nq][slots_total][k] when slots_total > 1; with slots_total == 1 the rows go to `out`
cudaError_t launch_knn_scan_range(const IndexDev &idx, const float *d_queries, int64_t nq, int k, int qw, int slices,
                                  int64_t slice_len, int64_t first, float *od, int *oi, int slot_base, int slots_total,
                                  const RowDests &out, cudaStream_t stream);
cudaError_t launch_knn_merge(const float *part_d, const int *part_i, int64_t nq, int k, int slots, const RowDests &out,
                             cudaStream_t stream);

// the radius scans cover nodes [first, idx.n); counts are accumulated with atomicAdd (zero them first)
cudaError_t launch_radius_count(const IndexDev &idx, const float *d_queries, int64_t nq, float r2, int32_t *d_counts,
                                const KnnPlan &plan, cudaStream_t stream, int64_t first = 0);

// d_cursor must be zeroed [nq]; d_offsets = exclusive scan of counts [nq]; d_keys receives (d2 bits << 32 | id)
cudaError_t launch_radius_fill(const IndexDev &idx, const float *d_queries, int64_t nq, float r2, const int64_t *d_offsets,
                               int32_t *d_cursor, unsigned long long *d_keys, const KnnPlan &plan, cudaStream_t stream,
                               int64_t first = 0);

// planner-sized radius search in one kernel (block per query, exhaustive over a small index): rows sorted by (d2, id), packed in
// the order the blocks finish (rowoff_out[q] = start of row q, -1 + *overflow = 1 when a row or the output did not fit);
// counts_out is exact in every case; *d_total must be zero
cudaError_t launch_radius_fused(const IndexDev &idx, const float *queries, int64_t nq, float r2, int64_t out_cap, int32_t *counts_out,
                                int64_t *rowoff_out, int32_t *ids_out, float *d2_out, unsigned long long *d_total, int *overflow,
                                cudaStream_t stream);

// d_offsets = exclusive scan of d_counts (one block, device side), *total_out = sum (device-accessible host word); when the sum
// exceeds `capacity` all offsets become -1 and the fill / sort launches that follow do nothing
cudaError_t launch_radius_offsets(const int32_t *d_counts, int64_t nq, int64_t capacity, int64_t *d_offsets, int64_t *total_out,
                                  cudaStream_t stream);

// sorts every row of d_keys ascending and unpacks it into ids / d2
cudaError_t launch_radius_sort(unsigned long long *d_keys, const int64_t *d_offsets, const int32_t *d_counts, int64_t nq,
                               int32_t *d_ids, float *d_d2, cudaStream_t stream);

// AoS [n][dim] -> SoA append at position `at`
cudaError_t launch_index_append(float *d_coords, int64_t capacity, int dim, int64_t at, const float *d_pts, int64_t n,
                                unsigned *d_amax, cudaStream_t stream);

}  // namespace sffg
