// handles.cu — the step that follows the hot path in every caller: clustering of the positive grasps
// into "handles" (HandleSearch::findHandles, src/agile_grasp/handle_search.cpp:4-118; class Handle,
// src/agile_grasp/handle.cpp:3-73; callers grasp_localizer.cpp:103, src/nodes/test.cpp:97).
//
// The reference tests every ordered pair (i, j) with three acos calls inside a serial double loop.
// Here the pair predicate — it does not depend on the clustering state — is evaluated for all n^2
// pairs on the GPU into an n x n bit matrix (one warp per row, ballot per 32 columns); the greedy,
// order-dependent part (skip eliminated grasps, sort the inliers along the axis, cut at the first gap,
// accept, eliminate) walks the set bits on the host, where it costs O(n * inliers).
// Compiled with -fmad=false: the binary64 expressions round like the reference's SSE2 build.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ag_internal.h"

namespace ag {
namespace {

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
__device__ __forceinline__ double safe_acos(double x) {  // handle_search.cpp:121-128
  if (x < -1.0) x = -1.0;
  else if (x > 1.0) x = 1.0;
  return acos(x);
}

// bits[i * words + w] bit b  <=>  grasp j = 32 w + b is an inlier of the line through grasp i
// (handle_search.cpp:31-37: dist_from_line < 0.01, dist_angle_axis < 0.34, dist_from_normal < 0.34)
__global__ void __launch_bounds__(256)
k_handle_pairs(const ag_grasp* __restrict__ g, int n, int words, uint32_t* __restrict__ bits) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  double ia[3], ip[3], in_[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    ia[k] = g[i].axis[k];
    ip[k] = g[i].bottom[k];
    in_[k] = g[i].approach[k];
  }
  for (int w = 0; w < words; w++) {
    const int j = w * 32 + lane;
    bool ok = false;
    if (j < n) {
      double d[3], v[3];
#pragma unroll
      for (int k = 0; k < 3; k++) d[k] = g[j].bottom[k] - ip[k];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const double m0 = (r == 0 ? 1.0 : 0.0) - ia[r] * ia[0];
        const double m1 = (r == 1 ? 1.0 : 0.0) - ia[r] * ia[1];
        const double m2 = (r == 2 ? 1.0 : 0.0) - ia[r] * ia[2];
        v[r] = m0 * d[0] + (m1 * d[1] + m2 * d[2]);
      }
      const double dist_from_line = sqrt(dot3(v, v));
      const double aa = safe_acos(dot3(ia, g[j].axis));
      const double dist_angle_axis = fmin(aa, M_PI - aa);
      const double dist_from_normal = safe_acos(dot3(in_, g[j].approach));
      ok = dist_from_line < 0.01 && dist_angle_axis < 0.34 && dist_from_normal < 0.34;
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) bits[size_t(i) * words + w] = m;
  }
}

inline double hdot3(const double* a, const double* b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }

// unit eigenvector of the largest eigenvalue of the symmetric 3x3 matrix S (closed form: trigonometric
// eigenvalue, eigenvector from the cross products of the rows of S - lambda I, refined by two inverse-free
// power steps)
void principal_axis(const double S[3][3], double v[3]) {
  const double p1 = S[0][1] * S[0][1] + S[0][2] * S[0][2] + S[1][2] * S[1][2];
  const double q = (S[0][0] + S[1][1] + S[2][2]) / 3.0;
  double lam;
  if (p1 == 0.0) {
    lam = std::max(S[0][0], std::max(S[1][1], S[2][2]));
  } else {
    const double p2 = (S[0][0] - q) * (S[0][0] - q) + (S[1][1] - q) * (S[1][1] - q) + (S[2][2] - q) * (S[2][2] - q) + 2.0 * p1;
    const double p = std::sqrt(p2 / 6.0);
    double B[3][3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) B[r][c] = (S[r][c] - (r == c ? q : 0.0)) / p;
    const double detB = B[0][0] * (B[1][1] * B[2][2] - B[1][2] * B[2][1]) - B[0][1] * (B[1][0] * B[2][2] - B[1][2] * B[2][0]) +
                        B[0][2] * (B[1][0] * B[2][1] - B[1][1] * B[2][0]);
    double r = detB / 2.0;
    r = std::min(1.0, std::max(-1.0, r));
    lam = q + 2.0 * p * std::cos(std::acos(r) / 3.0);
  }
  double M[3][3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) M[r][c] = S[r][c] - (r == c ? lam : 0.0);
  double best[3] = {0, 0, 0}, bn = -1.0;
  for (int a = 0; a < 3; a++) {
    const int b = (a + 1) % 3;
    const double c[3] = {M[a][1] * M[b][2] - M[a][2] * M[b][1], M[a][2] * M[b][0] - M[a][0] * M[b][2],
                         M[a][0] * M[b][1] - M[a][1] * M[b][0]};
    const double nn = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    if (nn > bn) {
      bn = nn;
      std::memcpy(best, c, sizeof(c));
    }
  }
  if (!(bn > 0.0)) {  // (near) multiple largest eigenvalue: any unit vector of the eigenspace; take the largest diagonal
    int m = 0;
    if (S[1][1] > S[m][m]) m = 1;
    if (S[2][2] > S[m][m]) m = 2;
    best[0] = best[1] = best[2] = 0.0;
    best[m] = 1.0;
  }
  for (int it = 0; it < 3; it++) {  // power steps polish the closed-form vector to working precision
    const double nrm = std::sqrt(best[0] * best[0] + best[1] * best[1] + best[2] * best[2]);
    double u[3] = {best[0] / nrm, best[1] / nrm, best[2] / nrm};
    if (it == 2) {
      std::memcpy(v, u, sizeof(u));
      return;
    }
    for (int r = 0; r < 3; r++) best[r] = S[r][0] * u[0] + S[r][1] * u[1] + S[r][2] * u[2];
  }
}

}  // namespace
}  // namespace ag

using namespace ag;

int ag_find_handles(ag_ctx* h, const ag_grasp* hands, int n, int min_inliers, double min_length,
                    ag_handle** handles_out, int* n_handles, int32_t** inliers_out, int* n_inliers_total) {
  if (!h || !handles_out || !n_handles || !inliers_out || !n_inliers_total || (n > 0 && !hands)) return AG_ERR_INVALID;
  *handles_out = nullptr;
  *inliers_out = nullptr;
  *n_handles = 0;
  *n_inliers_total = 0;
  Ctx& c = ctx_of(h);
  cudaSetDevice(c.device);
  std::vector<ag_handle> out;
  std::vector<int32_t> all_in;
  if (n > 0) {
    const int words = (n + 31) / 32;
    if (c.handle_in.reserve(size_t(n) * sizeof(ag_grasp)) || c.handle_bits.reserve(size_t(n) * words * 4)) return AG_ERR_CUDA;
    AG_CUDA_CHECK(cudaMemcpyAsync(c.handle_in.p, hands, size_t(n) * sizeof(ag_grasp), cudaMemcpyHostToDevice, c.stream));
    k_handle_pairs<<<(n + 7) / 8, 256, 0, c.stream>>>(c.handle_in.as<ag_grasp>(), n, words, c.handle_bits.as<uint32_t>());
    c.launches += 1;
    std::vector<uint32_t> bits(size_t(n) * words);
    AG_CUDA_CHECK(cudaMemcpyAsync(bits.data(), c.handle_bits.p, bits.size() * 4, cudaMemcpyDeviceToHost, c.stream));
    AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    AG_CUDA_CHECK(cudaGetLastError());
    // greedy clustering on the host (handle_search.cpp:11-84)
    std::vector<uint8_t> gone(n, 0);
    for (int i = 0; i < n; i++)
      if (hands[i].width == -1) gone[i] = 1;  // the reference's sentinel (:13,:23) also hides such inputs
    struct In {
      int j;
      double d;
    };
    std::vector<In> inl;
    for (int i = 0; i < n; i++) {
      if (gone[i]) continue;
      inl.clear();
      const uint32_t* row = &bits[size_t(i) * words];
      for (int w = 0; w < words; w++) {
        uint32_t m = row[w];
        while (m) {
          const int b = __builtin_ctz(m);
          m &= m - 1;
          const int j = w * 32 + b;
          if (gone[j]) continue;
          const double d[3] = {hands[j].bottom[0] - hands[i].bottom[0], hands[j].bottom[1] - hands[i].bottom[1],
                               hands[j].bottom[2] - hands[i].bottom[2]};
          inl.push_back({j, hdot3(hands[i].axis, d)});  // dist_along_line (:32)
        }
      }
      if (int(inl.size()) < min_inliers) continue;
      // shortenHandle (:92-118) as it executes: `inliers[i](2)` reads the next element's index, never
      // negative, so the list is cut to the elements strictly before the first gap position
      std::sort(inl.begin(), inl.end(), [](const In& a, const In& b) { return a.d != b.d ? a.d < b.d : a.j < b.j; });
      for (size_t k = 0; k + 1 < inl.size(); k++)
        if (inl[k + 1].d - inl[k].d > 0.02) {
          inl.resize(k);
          break;
        }
      if (int(inl.size()) < min_inliers) continue;
      // an empty list (cut at position 0, or a NaN axis that is not its own inlier) yields max - min < 0 in the
      // reference and no handle; it must not reach back() / front() when min_inliers <= 0
      if (inl.empty()) continue;
      if (!(inl.back().d - inl.front().d > min_length)) continue;  // sorted: max - min (:58-72)
      // Handle (handle.cpp:3-73)
      ag_handle H;
      std::memset(&H, 0, sizeof(H));
      double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (const In& e : inl)
        for (int r = 0; r < 3; r++)
          for (int cc = 0; cc < 3; cc++) S[r][cc] += hands[e.j].axis[r] * hands[e.j].axis[cc];
      principal_axis(S, H.axis);
      if (hdot3(H.axis, hands[inl[0].j].axis) < 0)
        for (int k = 0; k < 3; k++) H.axis[k] = -H.axis[k];
      double lo = 0, hi = 0;
      std::vector<double> along(inl.size());
      for (size_t k = 0; k < inl.size(); k++) {
        along[k] = hdot3(H.axis, hands[inl[k].j].bottom);
        lo = k == 0 ? along[k] : std::min(lo, along[k]);
        hi = k == 0 ? along[k] : std::max(hi, along[k]);
      }
      const double center_dist = (hi + lo) / 2.0;
      size_t bi = 0;
      double best = 10000000;
      for (size_t k = 0; k < inl.size(); k++) {
        const double dd = std::fabs(along[k] - center_dist);
        if (dd < best) {
          best = dd;
          bi = k;
        }
      }
      const ag_grasp& g = hands[inl[bi].j];
      for (int k = 0; k < 3; k++) {
        H.center[k] = g.bottom[k];
        H.approach[k] = g.approach[k];
        H.hands_center[k] = g.surface[k];
      }
      H.binormal[0] = H.approach[1] * H.axis[2] - H.approach[2] * H.axis[1];
      H.binormal[1] = H.approach[2] * H.axis[0] - H.approach[0] * H.axis[2];
      H.binormal[2] = H.approach[0] * H.axis[1] - H.approach[1] * H.axis[0];
      double wsum = 0.0;
      for (const In& e : inl) wsum += hands[e.j].width;
      H.width = wsum / double(inl.size());
      H.n_inliers = int(inl.size());
      H.inlier_offset = int(all_in.size());
      for (const In& e : inl) {
        all_in.push_back(e.j);
        gone[e.j] = 1;
      }
      out.push_back(H);
    }
  }
  *n_handles = int(out.size());
  *n_inliers_total = int(all_in.size());
  *handles_out = static_cast<ag_handle*>(std::malloc(std::max<size_t>(1, out.size()) * sizeof(ag_handle)));
  *inliers_out = static_cast<int32_t*>(std::malloc(std::max<size_t>(1, all_in.size()) * sizeof(int32_t)));
  if (!out.empty()) std::memcpy(*handles_out, out.data(), out.size() * sizeof(ag_handle));
  if (!all_in.empty()) std::memcpy(*inliers_out, all_in.data(), all_in.size() * sizeof(int32_t));
  return AG_OK;
}
