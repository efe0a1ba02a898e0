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

// ---- neighbour index ---------------------------------------------------------------------------
// The voxelised cloud is emitted in the reference's order: per camera, voxels sorted
// lexicographically by integer key (kx, ky, kz) (localization.h:281-292), i.e. all voxels with the
// same kx are contiguous and sorted by (y, z).  The only index built on top is, per camera, a row
// table row_ptr[kx] = first voxel with key_x >= kx.  A radius query is then, for each of the
// <= 2r/voxel + 1 x-rows it can touch, ONE contiguous run found by a binary search on y: coalesced
// 16-byte-per-lane streams, no second sort, no permutation, no hash-grid cell table.
struct RowIndex {         // lives in device memory, written by the voxelisation kernels
  double mn[2][3];        // per-camera lattice origin (the per-camera coordinate minima)
  double inv_cell;        // 1 / voxel size
  int nx[2];              // number of x-rows per camera
  int row_base[2];        // offset of camera c's rows inside the row table
  int first[2];           // index of the camera's first voxel
  int count[2];           // voxels per camera
  int n_points;           // total voxels
  int n_samples;          // samples actually used by the current call (<= requested)
  int error;              // sticky device-side error flags (kErr*)
  int use_cols;           // 1 if the dense (kx, ky) column table below is populated
  int cols_bad;           // set when the table was abandoned (a gap too long to fill inline: very sparse cloud)
  int ny[2];              // y-cells per x-row of the column table
  int col_base[2];        // offset of camera c inside the column table
  int pad;
};
constexpr int kErrKeyOverflow = 1;   // voxel key outside the workspace-derived bit budget
constexpr int kErrBadIndex = 2;      // a caller-supplied sample index is outside the voxelised cloud
constexpr int kErrBallOverflow = 4;  // a radius ball held more points than a neighbour-pool slot
constexpr int kErrBitmapRetry = 8;   // the voxel lattice does not fit the occupancy bitmap: re-run on the key-sort path
constexpr int AG_RETRY_KEYSORT = 1;  // internal status (> 0): the caller switches the context to the key-sort path and re-runs

// voxel record: xyz + tag.  tag bit 0 = camera source, bit 1 = "cloud_normals_ holds a non-zero normal"
struct __align__(16) GPoint {
  float x, y, z;
  uint32_t tag;
};
constexpr uint32_t kTagCamBit = 1u;
constexpr uint32_t kTagNormalBit = 2u;

// FLANN L2_Simple<float> distance, exactly as the reference's kd-tree evaluates it:
// ((dx*dx) + dy*dy) + dz*dz in binary32, no FMA contraction.
__device__ __forceinline__ float dist2_flann(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  float r = __fmul_rn(dx, dx);
  r = __fadd_rn(r, __fmul_rn(dy, dy));
  r = __fadd_rn(r, __fmul_rn(dz, dz));
  return r;
}

// x-row range of camera c that can hold points within rpad of qx (conservative: the stored x of row k
// is float(k*cell + min), within 1e-5 cells of the binary64 value used here)
__device__ __forceinline__ void row_range(const RowIndex& ri, int c, float qx, double rpad, int& k_lo, int& k_hi) {
  const double a = (double(qx) - rpad - ri.mn[c][0]) * ri.inv_cell - 1e-3;
  const double b = (double(qx) + rpad - ri.mn[c][0]) * ri.inv_cell + 1e-3;
  k_lo = a <= 0.0 ? 0 : (a >= 2147483000.0 ? 2147483000 : int(a));  // floor for a >= 0
  k_hi = b < 0.0 ? -1 : (b >= double(ri.nx[c] - 1) ? ri.nx[c] - 1 : int(b));
}

// the run of row `k` of camera c with |y - qy| <= chord(rpad, dx): [j0, j1).
// Fast path: the dense column table col_ptr[kx*ny + ky] = first voxel with key >= (kx, ky) gives both
// ends with two independent loads.  Fallback (table too large for the cloud's extent, or a cloud loaded
// through ag_set_cloud): two lock-stepped binary searches on the row's sorted y.
__device__ __forceinline__ void row_run(const RowIndex& ri, const int* __restrict__ row_ptr,
                                        const int* __restrict__ col_ptr, const GPoint* __restrict__ pts, int c, int k,
                                        float qx, float qy, double rpad, int& j0, int& j1) {
  // row x in binary64 (the stored float is within 1e-7 of it); shrink |dx| by a margin so the chord is conservative
  const double dx = fmax(0.0, fabs(double(k) / ri.inv_cell + ri.mn[c][0] - double(qx)) - 1e-6);
  const double half = sqrt(fmax(0.0, rpad * rpad - dx * dx)) + 1e-7;
  const double ylo = double(qy) - half, yhi = double(qy) + half;
  if (ri.use_cols && !ri.cols_bad) {
    const int ny = ri.ny[c];
    const double fa = (ylo - ri.mn[c][1]) * ri.inv_cell - 1e-3, fb = (yhi - ri.mn[c][1]) * ri.inv_cell + 1e-3;
    const int ka = fa <= 0.0 ? 0 : (fa >= double(ny) ? ny : int(fa));
    const int kb = fb < 0.0 ? 0 : (fb >= double(ny - 1) ? ny : int(fb) + 1);
    const int* col = col_ptr + ri.col_base[c] + k * ny;
    j0 = __ldg(col + ka);
    j1 = __ldg(col + kb);
    return;
  }
  int a = __ldg(row_ptr + ri.row_base[c] + k), b = __ldg(row_ptr + ri.row_base[c] + k + 1);
  // two lower-bound searches advanced in lock step (independent load chains -> half the latency)
  int lo0 = a, hi0 = b, lo1 = a, hi1 = b;
  while (lo0 < hi0 || lo1 < hi1) {
    const int m0 = (lo0 + hi0) >> 1, m1 = (lo1 + hi1) >> 1;
    const float y0 = lo0 < hi0 ? __ldg(&pts[m0].y) : 0.f;
    const float y1 = lo1 < hi1 ? __ldg(&pts[m1].y) : 0.f;
    if (lo0 < hi0) {  // first j with y >= ylo
      if (double(y0) < ylo) lo0 = m0 + 1;
      else hi0 = m0;
    }
    if (lo1 < hi1) {  // first j with y > yhi
      if (double(y1) <= yhi) lo1 = m1 + 1;
      else hi1 = m1;
    }
  }
  j0 = lo0;
  j1 = lo1 < lo0 ? lo0 : lo1;
}

// ---- seeded stratified sample draw (replaces the time-seeded pcl::RandomSample, hand_search.cpp:36-39): sample k of
// S lies in [floor(k n / S), floor((k + 1) n / S)), one index from each stratum -> sorted, distinct.  Runs on the
// device because the voxel count n never visits the host mid-pipeline.
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
struct DrawArgs {   // a shard handles the samples k = k_first + j k_step, j < count (0, 1, s_req without sharding)
  int* out;         // null: no draw
  uint64_t seed;
  int s_req, k_first, k_step, count;
};
__device__ __forceinline__ int draw_count(const DrawArgs& d, int n) {  // valid entries of d.out
  const int S = d.s_req < n ? d.s_req : n;
  return S > d.k_first ? min((S - d.k_first + d.k_step - 1) / d.k_step, d.count) : 0;
}
__device__ __forceinline__ int draw_sample(const DrawArgs& d, int n, int j) {
  const int S = d.s_req < n ? d.s_req : n;
  const int k = d.k_first + j * d.k_step;
  if (k >= S) return -1;
  const long long lo = (static_cast<long long>(k) * n) / S, hi = (static_cast<long long>(k + 1) * n) / S;
  const uint64_t h = splitmix64(d.seed ^ splitmix64(uint64_t(k)));
  return int(lo + static_cast<long long>(h % uint64_t(hi - lo)));
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

// Sum-reduction of 32 per-lane arrays across the warp by recursive halving: after 5 exchange rounds
// (16+8+4+2+1 = 31 value exchanges instead of 32 x 5 for 32 independent butterflies) lane i holds the
// warp-wide sum of element i.  v[] must be indexed with compile-time constants only (registers).
template <int N>
__device__ __forceinline__ double warp_reduce_transpose32(double (&v)[N]) {
  static_assert(N >= 32, "needs at least 32 elements");
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; i++) {
      // keep elements whose index bit `half` equals this lane's bit; send the other half to the partner
      const double send = upper ? v[i] : v[i + half];
      const double keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

// Builds the run list of a radius query for one warp: runs (start, length) of candidate points, as an
// exclusive prefix over lengths in pre[0..n_runs].  rs / pre are this warp's shared arrays of capacity
// cap / cap+1.  Returns the number of runs; `more` is set when the query touches more rows than fit
// (the caller then continues with `row_off` advanced) — never the case for the shipped radii.
__device__ __forceinline__ int build_runs_warp(const RowIndex& ri, const int* __restrict__ row_ptr,
                                               const int* __restrict__ col_ptr, const GPoint* __restrict__ pts,
                                               float qx, float qy, double rpad,
                                               int* rs, int* pre, int cap, int row_off, bool& more) {
  const int lane = threadIdx.x & 31;
  int n_runs = 0;
  int skip = row_off;
  more = false;
  for (int c = 0; c < 2; c++) {
    if (ri.count[c] == 0) continue;
    int k_lo, k_hi;
    row_range(ri, c, qx, rpad, k_lo, k_hi);
    int rows = k_hi - k_lo + 1;
    if (rows <= 0) continue;
    if (skip >= rows) {
      skip -= rows;
      continue;
    }
    k_lo += skip;
    rows -= skip;
    skip = 0;
    if (n_runs + rows > cap) {
      rows = cap - n_runs;
      more = true;
    }
    for (int t = lane; t < rows; t += 32) {
      int j0, j1;
      row_run(ri, row_ptr, col_ptr, pts, c, k_lo + t, qx, qy, rpad, j0, j1);
      rs[n_runs + t] = j0;
      pre[n_runs + t + 1] = j1 - j0;
    }
    n_runs += rows;
    if (more) break;
  }
  __syncwarp();
  // inclusive scan of the run lengths -> exclusive prefix in pre[]
  int carry = 0;
  for (int base = 0; base < n_runs; base += 32) {
    const int i = base + lane;
    int v = i < n_runs ? pre[i + 1] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (i < n_runs) pre[i + 1] = v + carry;
    carry += __shfl_sync(0xffffffffu, v, 31);
  }
  if (lane == 0) pre[0] = 0;
  __syncwarp();
  return n_runs;
}

}  // namespace ag
