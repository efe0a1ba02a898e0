// oracle_handles.cpp — CPU oracle of the step right after the hot path: HandleSearch::findHandles
// (src/agile_grasp/handle_search.cpp:4-89), shortenHandle (:92-118), safeAcos (:121-128) and the Handle
// constructor (src/agile_grasp/handle.cpp:3-73), restated on ag_grasp records.
//
// TEST INFRASTRUCTURE ONLY (see ag_oracle.h).  Parity unpinned: the reference has no test or data for
// this step and cannot be built here.  Reference behaviour kept as it actually executes:
//  * shortenHandle reads `inliers[i](2)` on a Vector2d (handle_search.cpp:103): with the contiguous
//    std::vector<Vector2d> storage that is the first component of the NEXT element, i.e. an inlier index,
//    never negative — so the `< 0` branch is dead, the list is always cut to the elements strictly before
//    position i (element i itself is dropped too) and the while loop ends after one call.
//  * std::sort with LastElementComparator is not stable; ties are broken here by inlier index.
//  * Handle::setAxis takes the eigenvector of Eigen::EigenSolver whose sign is an artefact of its QR
//    iteration; here the sign is fixed so that the axis has a non-negative dot product with the axis of
//    the first inlier.
// Eigen evaluation order assumed for fixed-size 3-vectors: a0*b0 + (a1*b1 + a2*b2).  Build with
// -ffp-contract=off.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ag_oracle.h"

namespace {

inline double dot3(const double* a, const double* b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
inline double safe_acos(double x) {  // handle_search.cpp:121-128
  if (x < -1.0) x = -1.0;
  else if (x > 1.0) x = 1.0;
  return acos(x);
}

// eigenvector of the largest eigenvalue of a symmetric 3x3 matrix: cyclic Jacobi in binary64
void max_eigvec3(const double S[3][3], double v[3]) {
  double a[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  std::memcpy(a, S, sizeof(a));
  for (int sweep = 0; sweep < 60; sweep++) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    if (off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int m = 0;
  if (a[1][1] > a[m][m]) m = 1;
  if (a[2][2] > a[m][m]) m = 2;
  const double n = sqrt(V[0][m] * V[0][m] + V[1][m] * V[1][m] + V[2][m] * V[2][m]);
  for (int k = 0; k < 3; k++) v[k] = V[k][m] / n;
}

struct Inlier {
  int j;
  double d;
};

}  // namespace

extern "C" int ago_find_handles(const ag_grasp* hands, int n, int min_inliers, double min_length,
                                ag_handle** handles_out, int* n_handles, int32_t** inliers_out, int* n_inliers_total) {
  std::vector<double> width(n);
  for (int i = 0; i < n; i++) width[i] = hands[i].width;  // reduced_hand_list (:9)
  std::vector<ag_handle> out;
  std::vector<int32_t> all_in;
  for (int i = 0; i < n; i++) {
    if (width[i] == -1) continue;  // :13
    const double* iaxis = hands[i].axis;
    const double* ipt = hands[i].bottom;
    const double* inormal = hands[i].approach;
    std::vector<Inlier> inl;
    for (int j = 0; j < n; j++) {
      if (width[j] == -1) continue;  // :23
      const double* jaxis = hands[j].axis;
      const double* jpt = hands[j].bottom;
      const double* jnormal = hands[j].approach;
      const double d[3] = {jpt[0] - ipt[0], jpt[1] - ipt[1], jpt[2] - ipt[2]};
      // (I - a a^T) * d, then norm   (:31)
      double v[3];
      for (int r = 0; r < 3; r++) {
        const double m0 = (r == 0 ? 1.0 : 0.0) - iaxis[r] * iaxis[0];
        const double m1 = (r == 1 ? 1.0 : 0.0) - iaxis[r] * iaxis[1];
        const double m2 = (r == 2 ? 1.0 : 0.0) - iaxis[r] * iaxis[2];
        v[r] = m0 * d[0] + (m1 * d[1] + m2 * d[2]);
      }
      const double dist_from_line = sqrt(dot3(v, v));
      const double dist_along_line = dot3(iaxis, d);  // :32
      const double aa = safe_acos(dot3(iaxis, jaxis));
      const double dist_angle_axis = std::min(aa, M_PI - aa);  // :33-35
      const double dist_from_normal = safe_acos(dot3(inormal, jnormal));  // :36
      if (dist_from_line < 0.01 && dist_angle_axis < 0.34 && dist_from_normal < 0.34) inl.push_back({j, dist_along_line});
    }
    if (int(inl.size()) < min_inliers) continue;  // :45
    // shortenHandle (:92-118), as it executes (see the header of this file)
    std::sort(inl.begin(), inl.end(), [](const Inlier& a, const Inlier& b) { return a.d != b.d ? a.d < b.d : a.j < b.j; });
    for (size_t k = 0; k + 1 < inl.size(); k++) {
      if (inl[k + 1].d - inl[k].d > 0.02) {
        inl.resize(k);
        break;
      }
    }
    if (int(inl.size()) < min_inliers) continue;  // :54
    double min_dist = 10000000, max_dist = -10000000;
    for (const Inlier& e : inl) {
      if (e.d < min_dist) min_dist = e.d;
      if (e.d > max_dist) max_dist = e.d;
    }
    if (!(max_dist - min_dist > min_length)) continue;  // :72
    // Handle::Handle (handle.cpp:3-73)
    ag_handle H;
    std::memset(&H, 0, sizeof(H));
    double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (const Inlier& e : inl)
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) S[r][c] += hands[e.j].axis[r] * hands[e.j].axis[c];
    max_eigvec3(S, H.axis);
    if (dot3(H.axis, hands[inl[0].j].axis) < 0)
      for (int k = 0; k < 3; k++) H.axis[k] = -H.axis[k];
    std::vector<double> along(inl.size());
    double lo = 0, hi = 0;
    for (size_t k = 0; k < inl.size(); k++) {
      along[k] = dot3(H.axis, hands[inl[k].j].bottom);
      if (k == 0 || along[k] < lo) lo = along[k];
      if (k == 0 || along[k] > hi) hi = along[k];
    }
    const double center_dist = (hi + lo) / 2.0;
    double best = 10000000;
    int bi = -1;
    for (size_t k = 0; k < inl.size(); k++) {
      const double dd = fabs(along[k] - center_dist);
      if (dd < best) {
        best = dd;
        bi = int(k);
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
    double w = 0.0;
    for (const Inlier& e : inl) w += hands[e.j].width;
    H.width = w / double(inl.size());
    H.n_inliers = int(inl.size());
    H.inlier_offset = int(all_in.size());
    for (const Inlier& e : inl) {
      all_in.push_back(e.j);
      width[e.j] = -1;  // :79-82
    }
    out.push_back(H);
  }
  *n_handles = int(out.size());
  *n_inliers_total = int(all_in.size());
  *handles_out = static_cast<ag_handle*>(std::malloc(std::max<size_t>(1, out.size()) * sizeof(ag_handle)));
  *inliers_out = static_cast<int32_t*>(std::malloc(std::max<size_t>(1, all_in.size()) * sizeof(int32_t)));
  if (!out.empty()) std::memcpy(*handles_out, out.data(), out.size() * sizeof(ag_handle));
  if (!all_in.empty()) std::memcpy(*inliers_out, all_in.data(), all_in.size() * sizeof(int32_t));
  return 0;
}
