// knn_pruned.cuh -- spatially sorted view of an index and the block-pruned exact k-NN scan (knn_pruned.cu)
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "knn_kernels.cuh"

namespace sffg {

// Morton-sorted copy of the first n_sorted nodes of an index, in blocks of 32 with exact per-block bounding boxes
struct SortedDev {
  const float *coords;   // SoA, coordinate c of sorted position p at coords[c * cap_s + p]
  const int *ids;        // original (insertion-order) id of sorted position p
  const float *bb;       // block boxes over the translational coordinates: lo_c at bb[c * nblk_cap + b], hi_c at bb[(LIN + c) * nblk_cap + b]
  const float *sbb;      // superblock boxes (32 blocks = 1024 nodes), same layout with stride nsb_cap
  long long cap_s;
  long long nblk_cap;
  long long nsb_cap;
  int n_sorted;
  int nblk;
  const unsigned *amax;  // as IndexDev::amax
  const int *bounds;     // 6 ordered ints: min / max of the translational coordinates at build time (Morton quantisation)
};

struct SortedBuildBuffers {
  int *bounds;                     // 6 ints
  unsigned *keys_in, *keys_out, *vals_in, *vals_out;
  void *temp;
  size_t temp_bytes;
  float *s_coords;
  long long cap_s;
  int *s_ids;
  float *bb;
  long long nblk_cap;
  float *sbb;
  long long nsb_cap;
};

size_t sorted_build_temp_bytes(int n);
cudaError_t launch_sorted_build(const IndexDev &idx, int n, const SortedBuildBuffers &b, cudaStream_t st);

struct PrunedPlan {
  int qw;
  int slices;         // slices over superblocks (32 blocks = 1024 nodes) of the sorted view
  int sb_per_slice;
  int tail_slices;    // slices of the exhaustive scan over the unsorted tail
  int64_t tail_len;
  // large batches: the query rows are visited in Morton order of their translational part (same curve as the nodes), so the
  // queries that share a warp are neighbours in space and need the same blocks; results still land in the caller's row order
  bool sort_queries;
  size_t sort_temp_bytes;
};
PrunedPlan plan_pruned(int64_t nq, const SortedDev &sv, int64_t tail, int sm_count);
size_t pruned_scratch_bytes(const PrunedPlan &p, int64_t nq, int k);
cudaError_t launch_knn_pruned(const IndexDev &idx, const SortedDev &sv, const float *d_queries, int64_t nq, int k, const RowDests &out,
                              void *d_scratch, const PrunedPlan &p, cudaStream_t st);

// radius search over sorted view + tail (counts must be zeroed / cursor zeroed by the caller)
cudaError_t launch_radius_count_pruned(const IndexDev &idx, const SortedDev &sv, const float *d_queries, int64_t nq, float r2,
                                       int32_t *d_counts, int sm_count, cudaStream_t st);
cudaError_t launch_radius_fill_pruned(const IndexDev &idx, const SortedDev &sv, const float *d_queries, int64_t nq, float r2,
                                      const int64_t *d_offsets, int32_t *d_cursor, unsigned long long *d_keys, int sm_count,
                                      cudaStream_t st);

}  // namespace sffg
