// bvh_device.cuh -- on-device builder of the 8-wide obstacle hierarchy (implemented in bvh_device.cu)
#pragma once
#include <cuda_runtime.h>

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
};

// d_soup: n triangles, 9 doubles each, original order, on the device.  Synchronises `st` before returning.
cudaError_t build_bvh_device(const double *d_soup, int n, cudaStream_t st, DeviceBvh *out);

}  // namespace sffg
