// ag_common.cuh — shared device-side definitions for the sm_100a grasp-hypothesis kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ag_b200.h"

#define AG_CUDA_CHECK(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ag::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                 \
      return AG_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

namespace ag {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Uniform hash grid over the voxelised cloud.  Points are sorted by linear cell id with z
// fastest, so all cells (cx, cy, z0..z1) of a column are one contiguous run of float4 records:
// every neighbourhood query is a handful of coalesced 16-byte-per-lane streams out of L2/HBM.
struct GridDesc {
  double gmin[3];   // lower corner (double so that binning is the same monotone function everywhere)
  double inv_cell;  // 1 / cell size
  int dim[3];       // cells per axis
  int n_points;
};

// point record in the cell-sorted array: xyz + (original voxel index | cam << 30 | has_normal << 31)
struct __align__(16) GPoint {
  float x, y, z;
  uint32_t tag;
};
constexpr uint32_t kTagIndexMask = 0x3FFFFFFFu;
constexpr uint32_t kTagCamBit = 0x40000000u;
constexpr uint32_t kTagNormalBit = 0x80000000u;

__host__ __device__ inline int cell_of(double v, double gmin, double inv_cell, int dim) {
  // monotone non-decreasing in v: subtraction and multiplication by a positive constant round
  // monotonically, floor is monotone.  Used for both binning and query ranges.
  double c = floor((v - gmin) * inv_cell);
  if (c < 0) c = 0;
  if (c > double(dim - 1)) c = double(dim - 1);
  return int(c);
}

// FLANN L2_Simple<float> distance, exactly as the reference's kd-tree evaluates it:
// ((dx*dx) + dy*dy) + dz*dz in binary32, no FMA contraction.
__device__ __forceinline__ float dist2_flann(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  float r = __fmul_rn(dx, dx);
  r = __fadd_rn(r, __fmul_rn(dy, dy));
  r = __fadd_rn(r, __fmul_rn(dz, dz));
  return r;
}

// Query footprint in the grid: cell ranges per axis and the list of (cx,cy) columns.
struct QueryBox {
  int lo[3], hi[3];
};
__device__ __forceinline__ QueryBox query_box(const GridDesc& g, float qx, float qy, float qz, double rpad) {
  QueryBox b;
  const double q[3] = {double(qx), double(qy), double(qz)};
#pragma unroll
  for (int a = 0; a < 3; a++) {
    b.lo[a] = cell_of(q[a] - rpad, g.gmin[a], g.inv_cell, g.dim[a]);
    b.hi[a] = cell_of(q[a] + rpad, g.gmin[a], g.inv_cell, g.dim[a]);
  }
  return b;
}
__device__ __forceinline__ int cell_linear(const GridDesc& g, int cx, int cy, int cz) {
  return (cx * g.dim[1] + cy) * g.dim[2] + cz;
}

// ordered-int encoding of floats for atomicMin/atomicMax
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__host__ __device__ inline float ordered_to_float(int i) {
  int j = i >= 0 ? i : i ^ 0x7FFFFFFF;
#ifdef __CUDA_ARCH__
  return __int_as_float(j);
#else
  float f;
  memcpy(&f, &j, 4);
  return f;
#endif
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ag
