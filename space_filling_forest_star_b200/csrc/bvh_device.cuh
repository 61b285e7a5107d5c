// bvh_device.cuh -- on-device builder of the 8-wide obstacle hierarchy (implemented in bvh_device.cu)
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "common.h"

namespace sffg {

// everything is device memory owned by the caller afterwards (cudaFree); leaf order = Morton order of the triangle boxes
struct DeviceBvh {
  ChildSlot *d_slots = nullptr;   // n_nodes * kWide, level by level (root = node 0)
  double *d_tris64 = nullptr;     // 9 doubles per triangle, leaf order
  float4 *d_tris32 = nullptr;     // 3 float4 per triangle, leaf order; p[0].w = representation error bound
  int *d_order = nullptr;         // leaf position -> original triangle index
  int n_nodes = 0, depth = 0;
  double root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
  std::vector<int> level_base, level_count;   // level l = nodes [level_base[l], +level_count[l])
};

// d_soup: n triangles, 9 doubles each, original order, on the device.  Synchronises `st` before returning.
cudaError_t build_bvh_device(const double *d_soup, int n, cudaStream_t st, DeviceBvh *out);

// Refit: the topology of an existing hierarchy (child links, leaf order, level table -- from either builder) is kept and
// only what depends on the vertex positions is recomputed from a new soup of the SAME n triangles in the SAME order: the
// leaf-order triangle arrays (FP64, FP32 + representation error bound) and, bottom-up, every slot box with outward
// rounding.  What an obstacle that moves or deforms between frames needs (the reference re-creates its RAPID model,
// src/environment.h:101-115).  Boxes stay conservative whatever the motion; only their tightness depends on how far the
// soup moved since the last full build.  Synchronises `st` before returning; root_lo / root_hi receive the new bounds.
cudaError_t refit_bvh_device(const double *d_soup, int n, ChildSlot *d_slots, double *d_tris64, float4 *d_tris32, const int *d_order,
                             const std::vector<int> &level_base, const std::vector<int> &level_count, cudaStream_t st,
                             double root_lo[3], double root_hi[3]);

}  // namespace sffg
