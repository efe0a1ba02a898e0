// oracle_core.cpp — CPU oracle, geometric half of the hot path.
//
// TEST INFRASTRUCTURE ONLY (see ag_oracle.h).  A literal restatement, in plain C++17 without
// Eigen/PCL/OpenCV, of what the reference computes.  Every block cites the reference lines it
// follows (paths relative to the reference repo).  It is deliberately *not* optimised: it walks
// the same loops in the same order as the reference so that the CUDA path — which uses different,
// GPU-friendly formulations — is checked against the reference's semantics, not against itself.
//
// Build with -ffp-contract=off (no FMA contraction) so products and sums round exactly like the
// reference's SSE2 build (CMakeLists.txt:18 has no -march flag).
//
// PARITY STATUS: parity unpinned by reference tests (none exist); third-party arithmetic pinned
// against this image's LAPACK (dggev_) and cv2 — see ag_oracle.h.

#include "oracle_internal.h"

#include <dlfcn.h>
#include <omp.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <set>
#include <string>
#include <vector>

namespace ago {

thread_local std::string g_err;
int fail(const std::string& msg) {
  g_err = msg;
  return -1;
}

// ---------------------------------------------------------------------------------------------
// LAPACK dggev_ provider (quadric.h:43-45 declares it; quadric.cpp:353,359 call it)
// ---------------------------------------------------------------------------------------------
using dggev_fn = void (*)(const char*, const char*, const int*, double*, const int*, double*, const int*,
                          double*, double*, double*, double*, const int*, double*, const int*, double*,
                          const int*, int*, size_t, size_t);
static dggev_fn g_dggev = nullptr;
static void* g_lapack_handle = nullptr;

static int set_lapack(const char* path, const char* symbol) {
  void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return fail(std::string("dlopen failed: ") + dlerror());
  void* f = dlsym(h, symbol);
  if (!f) {
    dlclose(h);
    return fail(std::string("symbol not found: ") + symbol);
  }
  g_lapack_handle = h;
  g_dggev = reinterpret_cast<dggev_fn>(f);
  // keep the BLAS single-threaded: the reference parallelises over samples, not inside LAPACK
  for (const char* name : {"openblas_set_num_threads", "scipy_openblas_set_num_threads"}) {
    if (void* s = dlsym(h, name)) {
      reinterpret_cast<void (*)(int)>(s)(1);
      break;
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// A.1 preprocess — localization.cpp:17-45 (cam source, NaN removal), :216-245 (filterWorkspace),
// :247-355 (voxelizeCloud), :357-362 (floorVector), localization.h:273-293 (comparator)
// ---------------------------------------------------------------------------------------------
struct Key3 {
  int k[3];
  bool operator<(const Key3& o) const {  // lexicographic (x,y,z): localization.h:281-292
    for (int i = 0; i < 3; i++)
      if (k[i] != o.k[i]) return k[i] < o.k[i];
    return false;
  }
  bool operator==(const Key3& o) const { return k[0] == o.k[0] && k[1] == o.k[1] && k[2] == o.k[2]; }
};

int preprocess(const void* points, int stride, int n_in, int size_left, const ag_params& P, bool use_std_set,
               std::vector<float>& xyz_out, std::vector<int32_t>& cam_out) {
  xyz_out.clear();
  cam_out.clear();
  if (n_in <= 0 || size_left == 0) return fail("Input cloud is empty!");  // localization.cpp:9-15
  const char* base = static_cast<const char*>(points);
  // localization.cpp:19-23: camera source by ORIGINAL position; :27 removeNaN compacts the cloud but
  // not the labels, so compacted point i reads label i (App. B#5) unless fix_cam_source.
  std::vector<float> p;  // finite points
  std::vector<int32_t> c;
  p.reserve(size_t(n_in) * 3);
  c.reserve(n_in);
  int compact = 0;
  for (int i = 0; i < n_in; i++) {
    float v[3];
    std::memcpy(v, base + size_t(i) * stride, 12);
    if (!std::isfinite(v[0]) || !std::isfinite(v[1]) || !std::isfinite(v[2])) continue;
    int label = P.fix_cam_source ? (i < size_left ? 0 : 1) : (compact < size_left ? 0 : 1);
    // localization.cpp:228-229 inclusive box test, float promoted to double
    const double* w = P.workspace;
    if (double(v[0]) >= w[0] && double(v[0]) <= w[1] && double(v[1]) >= w[2] && double(v[1]) <= w[3] &&
        double(v[2]) >= w[4] && double(v[2]) <= w[5]) {
      p.insert(p.end(), v, v + 3);
      c.push_back(label);
    }
    compact++;
  }
  const int n = int(c.size());
  // localization.cpp:250-279 per-camera minimum, initialised to 10000
  double mn[2][3] = {{10000, 10000, 10000}, {10000, 10000, 10000}};
  for (int i = 0; i < n; i++) {
    int s = c[i];
    if (s != 0 && s != 1) continue;
    for (int d = 0; d < 3; d++)
      if (double(p[3 * i + d]) < mn[s][d]) mn[s][d] = double(p[3 * i + d]);
  }
  const double cell = P.voxel_size;
  std::vector<Key3> keys[2];
  if (use_std_set) {
    // localization.cpp:282-298: std::set of integer cells, literally
    std::set<Key3> bins[2];
    for (int i = 0; i < n; i++) {
      int s = c[i];
      if (s != 0 && s != 1) continue;
      Key3 k;
      for (int d = 0; d < 3; d++) k.k[d] = int(std::floor((double(p[3 * i + d]) - mn[s][d]) / cell));
      bins[s].insert(k);
    }
    for (int s = 0; s < 2; s++) keys[s].assign(bins[s].begin(), bins[s].end());
  } else {
    for (int i = 0; i < n; i++) {
      int s = c[i];
      if (s != 0 && s != 1) continue;
      Key3 k;
      for (int d = 0; d < 3; d++) k.k[d] = int(std::floor((double(p[3 * i + d]) - mn[s][d]) / cell));
      keys[s].push_back(k);
    }
    for (int s = 0; s < 2; s++) {
      std::sort(keys[s].begin(), keys[s].end());
      keys[s].erase(std::unique(keys[s].begin(), keys[s].end()), keys[s].end());
    }
  }
  // localization.cpp:318-351: voxel corner = key*cell + min, stored as float; left block then right
  for (int s = 0; s < 2; s++)
    for (const Key3& k : keys[s]) {
      for (int d = 0; d < 3; d++) xyz_out.push_back(float(double(k.k[d]) * cell + mn[s][d]));
      cam_out.push_back(s);
    }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// C.1 neighbour search — call sites hand_search.cpp:10-11,85,147 (pcl::KdTreeFLANN radiusSearch).
// Semantics of FLANN L2_Simple<float>: dist = ((dx*dx) + dy*dy) + dz*dz in float, accepted iff
// dist < (float)(r*r); results sorted by (dist, index).
// ---------------------------------------------------------------------------------------------
static inline float dist2f(const float* a, const float* b) {
  float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
  float r = 0.0f;
  r += dx * dx;
  r += dy * dy;
  r += dz * dz;
  return r;
}

struct Tree {
  struct Node {
    int lo, hi;        // point range in perm (leaf)
    int left, right;   // children (-1 for leaf)
    int dim;
    float split_lo, split_hi;  // max of left side, min of right side along dim
  };
  const float* xyz = nullptr;
  int n = 0;
  std::vector<int> perm;
  std::vector<Node> nodes;
  static constexpr int kLeaf = 15;  // PCL builds KDTreeSingleIndexParams(15)

  int build(int lo, int hi) {
    Node nd{lo, hi, -1, -1, 0, 0.f, 0.f};
    int id = int(nodes.size());
    nodes.push_back(nd);
    if (hi - lo <= kLeaf) return id;
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = lo; i < hi; i++)
      for (int d = 0; d < 3; d++) {
        float v = xyz[3 * perm[i] + d];
        mn[d] = std::min(mn[d], v);
        mx[d] = std::max(mx[d], v);
      }
    int dim = 0;
    for (int d = 1; d < 3; d++)
      if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
    int mid = (lo + hi) / 2;
    std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi,
                     [&](int a, int b) { return xyz[3 * a + dim] < xyz[3 * b + dim]; });
    float slo = -1e30f, shi = 1e30f;
    for (int i = lo; i < mid; i++) slo = std::max(slo, xyz[3 * perm[i] + dim]);
    for (int i = mid; i < hi; i++) shi = std::min(shi, xyz[3 * perm[i] + dim]);
    int l = build(lo, mid);
    int r = build(mid, hi);
    nodes[id].left = l;
    nodes[id].right = r;
    nodes[id].dim = dim;
    nodes[id].split_lo = slo;
    nodes[id].split_hi = shi;
    return id;
  }
  void search(int id, const float* q, float r2, double rpad, std::vector<std::pair<float, int>>& out) const {
    const Node& nd = nodes[id];
    if (nd.left < 0) {
      for (int i = nd.lo; i < nd.hi; i++) {
        float d = dist2f(q, xyz + 3 * perm[i]);
        if (d < r2) out.emplace_back(d, perm[i]);
      }
      return;
    }
    // conservative pruning in double with padding; the exact test is the float one above
    double qd = q[nd.dim];
    if (qd - rpad <= double(nd.split_lo)) search(nd.left, q, r2, rpad, out);
    if (qd + rpad >= double(nd.split_hi)) search(nd.right, q, r2, rpad, out);
  }
};

void radius_search(const Tree* t, const float* xyz, int n, const float* q, double radius, int method,
                   std::vector<std::pair<float, int>>& out) {
  out.clear();
  const float r2 = float(radius * radius);  // PCL passes radius*radius to FLANN as float
  if (method == 0 || !t) {
    for (int i = 0; i < n; i++) {
      float d = dist2f(q, xyz + 3 * i);
      if (d < r2) out.emplace_back(d, i);
    }
  } else {
    double rpad = std::sqrt(double(r2)) * (1.0 + 1e-5) + 1e-7;
    if (!t->nodes.empty()) t->search(0, q, r2, rpad, out);
  }
  std::sort(out.begin(), out.end());  // (dist, index) ascending
}

// ---------------------------------------------------------------------------------------------
// small dense helpers
// ---------------------------------------------------------------------------------------------
// Fixed-size 3-vector reductions follow Eigen's unrolled redux order a0 + (a1 + a2)
// (Eigen/src/Core/Redux.h redux_novec_unroller, used by dot()/norm() of Vector3d).
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
static inline void cross3(const double* a, const double* b, double* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// symmetric 3x3 eigen-decomposition (cyclic Jacobi).  The reference uses Eigen::EigenSolver on
// this symmetric matrix (quadric.cpp:268-270); eigenvectors are unit length, sign is irrelevant
// downstream (quadric.cpp:285-304 re-orients everything).
static void eig3_sym(const double A[3][3], double w[3], double V[3][3]) {
  double a[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      a[i][j] = A[i][j];
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-40 * diag || off == 0.0) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; i++) w[i] = a[i][i];
}

// ---------------------------------------------------------------------------------------------
// A.3 Quadric::fitQuadric — quadric.cpp:14-157, solveGeneralizedEigenProblem :330-363
// ---------------------------------------------------------------------------------------------
struct QuadricFit {
  double params[10];   // as seen by findTaubinNormalAxis: a,b,c,d,e,f,g,h,i,j (d,e,f = 2*0.5*v3..5)
  double M[100], N[100];
  double lambda[10];
  int min_index;
};

static int fit_quadric(const float* xyz, const std::vector<int>& nn, QuadricFit& Q) {
  if (!g_dggev) return fail("no LAPACK dggev_ provider set (ago_set_lapack)");
  double* M = Q.M;
  double* N = Q.N;
  std::fill(M, M + 100, 0.0);
  std::fill(N, N + 100, 0.0);
  auto m = [&](int r, int c) -> double& { return M[r * 10 + c]; };
  auto nn_ = [&](int r, int c) -> double& { return N[r * 10 + c]; };
  const int n = int(nn.size());
  for (int t = 0; t < n; t++) {
    const float* pf = xyz + 3 * nn[t];
    if (std::isnan(pf[0])) continue;  // quadric.cpp:26
    const double x = pf[0], y = pf[1], z = pf[2];
    const double sq[3] = {x * x, y * y, z * z};          // x2 y2 z2
    const double b[9] = {sq[0], sq[1], sq[2], x * y, y * z, x * z, x, y, z};
    // quadric.cpp:40-66: rows 0..2 of M: M(r,c) += sq[r]*b[c] for c>=r, M(r,9) += sq[r]
    for (int r = 0; r < 3; r++) {
      for (int c = r; c < 9; c++) m(r, c) += sq[r] * b[c];
      m(r, 9) += sq[r];
    }
    m(3, 8) += x * b[4];  // :67  x*yz
    m(3, 9) += b[3];      // :68
    m(4, 9) += b[4];
    m(5, 9) += b[5];
    m(6, 9) += x;
    m(7, 9) += y;
    m(8, 9) += z;
    // quadric.cpp:103-131
    nn_(0, 0) += 4.0 * sq[0];
    nn_(0, 3) += 2.0 * b[3];
    nn_(0, 5) += 2.0 * b[5];
    nn_(0, 6) += 2.0 * x;
    nn_(1, 1) += 4.0 * sq[1];
    nn_(1, 3) += 2.0 * b[3];
    nn_(1, 4) += 2.0 * b[4];
    nn_(1, 7) += 2.0 * y;
    nn_(2, 2) += 4.0 * sq[2];
    nn_(2, 4) += 2.0 * b[4];
    nn_(2, 5) += 2.0 * b[5];
    nn_(2, 8) += 2.0 * z;
    nn_(3, 3) += sq[0] + sq[1];
    nn_(3, 4) += b[5];
    nn_(3, 5) += b[4];
    nn_(3, 6) += y;
    nn_(3, 7) += x;
    nn_(4, 4) += sq[1] + sq[2];
    nn_(4, 5) += b[3];
    nn_(4, 7) += z;
    nn_(4, 8) += y;
    nn_(5, 5) += sq[0] + sq[2];
    nn_(5, 6) += z;
    nn_(5, 8) += x;
  }
  // quadric.cpp:76-100 repeated entries
  m(3, 3) = m(0, 1); m(5, 5) = m(0, 2); m(3, 5) = m(0, 4); m(3, 6) = m(0, 7); m(5, 6) = m(0, 8);
  m(6, 6) = m(0, 9); m(4, 4) = m(1, 2); m(3, 4) = m(1, 5); m(3, 7) = m(1, 6); m(4, 7) = m(1, 8);
  m(7, 7) = m(1, 9); m(4, 5) = m(2, 3); m(5, 8) = m(2, 6); m(4, 8) = m(2, 7); m(8, 8) = m(2, 9);
  m(4, 6) = m(3, 8); m(5, 7) = m(3, 8); m(6, 7) = m(3, 9); m(7, 8) = m(4, 9); m(6, 8) = m(5, 9);
  m(9, 9) = n;  // :134
  nn_(6, 6) = n; nn_(7, 7) = n; nn_(8, 8) = n;  // :137-139
  for (int r = 0; r < 10; r++)
    for (int c = r + 1; c < 10; c++) {  // :136,141 mirror upper into lower
      m(c, r) = m(r, c);
      nn_(c, r) = nn_(r, c);
    }
  // quadric.cpp:330-363 — dggev_("N","V") with a workspace query first.  A,B are symmetric so
  // row-/column-major does not matter; dggev overwrites its inputs, so work on copies.
  double A[100], B[100], alphar[10], alphai[10], beta[10], VR[100], wq;
  std::memcpy(A, M, sizeof(A));
  std::memcpy(B, N, sizeof(B));
  int n10 = 10, lwork = -1, info = 0;
  g_dggev("N", "V", &n10, A, &n10, B, &n10, alphar, alphai, beta, nullptr, &n10, VR, &n10, &wq, &lwork, &info, 1, 1);
  lwork = int(wq) + 32;
  std::vector<double> work(lwork);
  g_dggev("N", "V", &n10, A, &n10, B, &n10, alphar, alphai, beta, nullptr, &n10, VR, &n10, work.data(), &lwork, &info,
          1, 1);
  // quadric.cpp:149-152: lambda = alphar ./ beta ; argmin over the first 9 (Eigen visitor: strict <)
  for (int i = 0; i < 10; i++) Q.lambda[i] = alphar[i] / beta[i];
  int mi = 0;
  double best = Q.lambda[0];
  for (int i = 1; i < 9; i++)
    if (Q.lambda[i] < best) {
      best = Q.lambda[i];
      mi = i;
    }
  Q.min_index = mi;
  for (int k = 0; k < 10; k++) Q.params[k] = VR[mi * 10 + k];  // column mi (column-major)
  // :153 halves 3..5, :165-167 doubles them again: both exact, so d,e,f = v3,v4,v5
  for (int k = 3; k < 6; k++) Q.params[k] = 2.0 * (Q.params[k] * 0.5);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// A.4 Quadric::findTaubinNormalAxis + findAverageNormalAxis — quadric.cpp:159-305
// (deterministic evaluation set = all neighbours, :204-212)
// ---------------------------------------------------------------------------------------------
// `coords` (3 per neighbour) are the coordinates in which `par` is expressed: the raw metric
// coordinates for the reference path, sample-centred 1/r-scaled ones for the extended-precision
// check (gradient DIRECTIONS are the same in both because the map is a translation + uniform scale).
static void local_axes(const std::vector<double>& coords, const int32_t* cam, const std::vector<int>& nn,
                       const double* par, const double sample[3], const double cam_origin[2][3], ag_frame& F) {
  const double a = par[0], b = par[1], c = par[2], d = par[3], e = par[4], f = par[5], g = par[6], h = par[7],
               i = par[8];
  const int m = int(nn.size());
  // :217-226 majority camera (maxCoeff: first max wins -> tie = 0)
  double cnt[2] = {0, 0};
  for (int t = 0; t < m; t++) {
    if (cam[nn[t]] == 0) cnt[0]++;
    else if (cam[nn[t]] == 1) cnt[1]++;
  }
  const int major = cnt[1] > cnt[0] ? 1 : 0;
  // :238-247 gradient normals
  std::vector<double> G(size_t(3) * m);
  for (int t = 0; t < m; t++) {
    const double x = coords[3 * t], y = coords[3 * t + 1], z = coords[3 * t + 2];
    double fx = (((2.0 * a) * x + d * y) + f * z) + g;
    double fy = (((2.0 * b) * y + d * x) + e * z) + h;
    double fz = (((2.0 * c) * z + e * y) + f * x) + i;
    double mag = std::sqrt((fx * fx + fy * fy) + fz * fz);
    G[3 * t] = fx / mag;
    G[3 * t + 1] = fy / mag;
    G[3 * t + 2] = fz / mag;
  }
  // :266 M = normals * normals^T (depth m, accumulated in column order)
  double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int t = 0; t < m; t++)
    for (int r = 0; r < 3; r++)
      for (int s = 0; s < 3; s++) C[r][s] += G[3 * t + r] * G[3 * t + s];
  double w[3], V[3][3];
  eig3_sym(C, w, V);
  int mi = 0;
  for (int k = 1; k < 3; k++)
    if (w[k] < w[mi]) mi = k;  // :278-280 minCoeff
  double ax[3] = {V[0][mi], V[1][mi], V[2][mi]};
  // :283-284 argmax_j sum_i ((n_i . n_j)^6); pow via libm as Eigen's array().pow(6) does
  int best = 0;
  double bestv = -1.0;
  for (int j = 0; j < m; j++) {
    double s = 0.0;
    const double* gj = &G[3 * j];
    for (int t = 0; t < m; t++) {
      const double* gt = &G[3 * t];
      double dt = (gt[0] * gj[0] + gt[1] * gj[1]) + gt[2] * gj[2];
      s += std::pow(dt, 6.0);
    }
    if (j == 0 || s > bestv) {
      bestv = s;
      best = j;
    }
  }
  // :285-288 normal = normalize((I - c c^T) n_best), gemv accumulates columns left to right
  double Pm[3][3];
  for (int r = 0; r < 3; r++)
    for (int s = 0; s < 3; s++) Pm[r][s] = (r == s ? 1.0 : 0.0) - ax[r] * ax[s];
  double np[3];
  for (int r = 0; r < 3; r++) np[r] = (Pm[r][0] * G[3 * best] + Pm[r][1] * G[3 * best + 1]) + Pm[r][2] * G[3 * best + 2];
  double nrm = std::sqrt(dot3(np, np));
  double nor[3] = {np[0] / nrm, np[1] / nrm, np[2] / nrm};
  double bin[3];
  cross3(ax, nor, bin);  // :291
  // :294-301 orient towards the majority camera
  double s2s[3] = {sample[0] - cam_origin[major][0], sample[1] - cam_origin[major][1],
                   sample[2] - cam_origin[major][2]};
  if (dot3(nor, s2s) > 0)
    for (double& v : nor) v *= -1.0;
  if (dot3(bin, s2s) > 0)
    for (double& v : bin) v *= -1.0;
  cross3(nor, bin, ax);  // :304
  for (int k = 0; k < 3; k++) {
    F.normal[k] = nor[k];
    F.axis[k] = ax[k];
    F.binormal[k] = bin[k];
  }
  F.num_neighbors = m;
  F.majority_cam = major;
}

// ---------------------------------------------------------------------------------------------
// Extended-precision check solve (NOT a reference path): the same Taubin fit, but posed in
// sample-centred, 1/r-scaled coordinates (exact in long double because the inputs are floats) and
// solved as the reduced 9x9 symmetric-definite pencil (A - m m^T/n) u = lambda B u by Cholesky +
// cyclic Jacobi in 80-bit long double.  It is accurate to ~1e-16 and is used to measure (a) how far
// LAPACK's dggev_ on the reference's uncentred, ill-conditioned 10x10 pencil lands from the exact
// answer and (b) how close the CUDA path is to it.
// ---------------------------------------------------------------------------------------------
typedef long double ld;
static void fit_quadric_exact(const float* xyz, const std::vector<int>& nn, const float* q, double radius,
                              double par_out[10], std::vector<double>& coords) {
  static const int E[10][3] = {{2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {1, 1, 0}, {0, 1, 1},
                               {1, 0, 1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
  const int n = int(nn.size());
  coords.resize(size_t(3) * n);
  ld mom[5][5][5];
  for (auto& a : mom) for (auto& b : a) for (auto& c : b) c = 0;
  const ld inv_r = ld(1) / ld(radius);
  for (int t = 0; t < n; t++) {
    ld v[3];
    for (int d = 0; d < 3; d++) {
      v[d] = (ld(xyz[3 * nn[t] + d]) - ld(q[d])) * inv_r;
      coords[3 * t + d] = double(v[d]);
    }
    ld px[5] = {1, 0, 0, 0, 0}, py[5] = {1, 0, 0, 0, 0}, pz[5] = {1, 0, 0, 0, 0};
    for (int k = 1; k < 5; k++) { px[k] = px[k - 1] * v[0]; py[k] = py[k - 1] * v[1]; pz[k] = pz[k - 1] * v[2]; }
    for (int a = 0; a < 5; a++)
      for (int b = 0; a + b < 5; b++)
        for (int c = 0; a + b + c < 5; c++) mom[a][b][c] += px[a] * py[b] * pz[c];
  }
  ld A[9][9], B[9][9], mv[9];
  const ld nn_ = mom[0][0][0];
  for (int i = 0; i < 9; i++) mv[i] = mom[E[i][0]][E[i][1]][E[i][2]];
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) {
      A[i][j] = mom[E[i][0] + E[j][0]][E[i][1] + E[j][1]][E[i][2] + E[j][2]] - mv[i] * mv[j] / nn_;
      ld bs = 0;
      for (int a = 0; a < 3; a++)
        if (E[i][a] >= 1 && E[j][a] >= 1) {
          int ex[3] = {E[i][0] + E[j][0], E[i][1] + E[j][1], E[i][2] + E[j][2]};
          ex[a] -= 2;
          bs += ld(E[i][a] * E[j][a]) * mom[ex[0]][ex[1]][ex[2]];
        }
      B[i][j] = bs;
    }
  // B = Q D Q^T (Jacobi); W = Q_k D_k^(-1/2) over the directions with D_i > 1e-11 D_max (B is singular
  // for degenerate neighbourhoods, e.g. voxel corners exactly in one lattice plane: those directions
  // have infinite or 0/0 eigenvalues and are excluded); C = W^T A W; smallest eigenpair; u = W y.
  auto jacobi = [](ld Mx[9][9], ld Vx[9][9]) {
    for (int i = 0; i < 9; i++)
      for (int j = 0; j < 9; j++) Vx[i][j] = i == j ? 1 : 0;
    for (int sweep = 0; sweep < 80; sweep++) {
      bool rotated = false;
      for (int p = 0; p < 8; p++)
        for (int q2 = p + 1; q2 < 9; q2++) {
          const ld apq = Mx[p][q2], app = Mx[p][p], aqq = Mx[q2][q2];
          if (!(fabsl(apq) > 1e-19L * sqrtl(fabsl(app * aqq)))) continue;
          rotated = true;
          const ld theta = (aqq - app) / (2 * apq);
          const ld tt = (theta >= 0 ? 1 : -1) / (fabsl(theta) + sqrtl(theta * theta + 1));
          const ld c = 1 / sqrtl(tt * tt + 1), s2 = tt * c;
          for (int k = 0; k < 9; k++) {
            const ld akp = Mx[k][p], akq = Mx[k][q2];
            Mx[k][p] = c * akp - s2 * akq;
            Mx[k][q2] = s2 * akp + c * akq;
            const ld vkp = Vx[k][p], vkq = Vx[k][q2];
            Vx[k][p] = c * vkp - s2 * vkq;
            Vx[k][q2] = s2 * vkp + c * vkq;
          }
          for (int k = 0; k < 9; k++) {
            const ld apk = Mx[p][k], aqk = Mx[q2][k];
            Mx[p][k] = c * apk - s2 * aqk;
            Mx[q2][k] = s2 * apk + c * aqk;
          }
          Mx[p][q2] = Mx[q2][p] = 0;
        }
      if (!rotated) break;
    }
  };
  ld Bq[9][9], Q[9][9], W[9][9], Cm[9][9], Y[9][9];
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) Bq[i][j] = B[i][j];
  jacobi(Bq, Q);
  ld dmax = 0;
  for (int i = 0; i < 9; i++) dmax = std::max(dmax, Bq[i][i]);
  bool dropped[9];
  for (int j = 0; j < 9; j++) {
    dropped[j] = !(Bq[j][j] > 1e-11L * dmax);
    const ld sc = dropped[j] ? 0 : 1 / sqrtl(Bq[j][j]);
    for (int i = 0; i < 9; i++) W[i][j] = Q[i][j] * sc;
  }
  for (int i = 0; i < 9; i++)
    for (int j = 0; j < 9; j++) {
      ld v = 0;
      for (int k = 0; k < 9; k++)
        for (int l = 0; l < 9; l++) v += W[k][i] * A[k][l] * W[l][j];
      Cm[i][j] = v;
    }
  for (int i = 0; i < 9; i++) {
    for (int j = i + 1; j < 9; j++) Cm[i][j] = Cm[j][i] = (Cm[i][j] + Cm[j][i]) / 2;
    if (dropped[i]) Cm[i][i] = 1e300L;
  }
  jacobi(Cm, Y);
  int mi = 0;
  for (int k = 1; k < 9; k++)
    if (Cm[k][k] < Cm[mi][mi]) mi = k;
  ld u[9];
  for (int i = 0; i < 9; i++) {
    ld v = 0;
    for (int k = 0; k < 9; k++) v += W[i][k] * Y[k][mi];
    u[i] = v;
  }
  ld mu = 0;
  for (int i = 0; i < 9; i++) { par_out[i] = double(u[i]); mu += mv[i] * u[i]; }
  par_out[9] = double(-mu / nn_);
}

// deterministic permutation used only for the summation-sensitivity probe
static void permute(std::vector<int>& v, int k) {
  uint64_t s = 0x9E3779B97F4A7C15ull * uint64_t(k + 1);
  for (size_t i = v.size(); i > 1; i--) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    std::swap(v[i - 1], v[s % i]);
  }
}

// glibc rand(): random_r TYPE_3 (additive feedback, degree 31, separation 3), the generator behind the
// reference's unseeded `rand() % indices.size()` (quadric.cpp:184).  Seed 1 = a process that never called
// srand.  Pinned against the C library itself in tests/test_oracle_quadric.py.
struct GlibcRand {
  uint32_t r[34];
  int pos = 0;  // next output is built at ring position pos
  explicit GlibcRand(uint32_t seed) {
    int32_t w[34];
    w[0] = int32_t(seed ? seed : 1u);
    for (int i = 1; i < 31; i++) {
      int64_t v = (16807LL * int64_t(w[i - 1])) % 2147483647LL;
      if (v < 0) v += 2147483647LL;
      w[i] = int32_t(v);
    }
    for (int i = 31; i < 34; i++) w[i] = w[i - 31];
    std::vector<uint32_t> u(w, w + 34);
    for (int i = 34; i < 344; i++) u.push_back(u[i - 31] + u[i - 3]);
    for (int i = 0; i < 34; i++) r[i] = u[344 - 34 + i];  // the last 34 words, r[33] = word 343
    pos = 0;
  }
  uint32_t next() {  // word k = word(k - 31) + word(k - 3); the ring holds the last 34 words, oldest at pos
    const uint32_t v = r[(pos + 3) % 34] + r[(pos + 31) % 34];
    r[pos] = v;
    pos = (pos + 1) % 34;
    return v >> 1;
  }
};

int fit_quadrics(const float* xyz, const int32_t* cam, int n, const Tree* tree, const int* indices, int S,
                 double radius, const ag_params& P, int sum_perm, ag_frame* frames, double* params_out,
                 double* MN_out, double* eig_out, size_t* rand_consumed) {
  const double cam_origin[2][3] = {{P.cam_tf_left[3], P.cam_tf_left[7], P.cam_tf_left[11]},
                                   {P.cam_tf_right[3], P.cam_tf_right[7], P.cam_tf_right[11]}};
  const int threads = std::max(1, P.num_threads);
  // production mode of the reference (is_deterministic = false, hand_search.h:84): normals from 50 neighbours
  // drawn with rand() % n (quadric.cpp:177-192).  The pinned stream is the one a single-threaded, never-seeded
  // process sees: 50 draws per sample with more than 50 neighbours, consumed in sample order.  The draws do not
  // depend on the fit, so the stream is laid out up front — a first parallel pass finds the
  // neighbour lists, a prefix count gives every sample its slice of the stream — and the samples are then
  // fitted on all threads with results identical to the serial walk.
  const bool rand_mode = P.deterministic_normals == 0;
  std::vector<std::vector<std::pair<float, int>>> res_all;
  std::vector<uint32_t> stream;
  std::vector<size_t> stream_off;
  if (rand_mode) {
    res_all.resize(S);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 8)
    for (int s = 0; s < S; s++) radius_search(tree, xyz, n, xyz + 3 * indices[s], radius, tree ? 1 : 0, res_all[s]);
    stream_off.resize(S);
    size_t used = 0;
    for (int s = 0; s < S; s++) {
      stream_off[s] = used;
      if (res_all[s].size() > 50) used += 50;
    }
    // rand() is process state: a second fit of the same localizeHands call (all-points pass, then the samples:
    // hand_search.cpp:17-26 then :77) continues where the first one stopped
    GlibcRand rng(1);
    const size_t skip = rand_consumed ? *rand_consumed : 0;
    for (size_t k = 0; k < skip; k++) rng.next();
    stream.resize(used);
    for (size_t k = 0; k < used; k++) stream[k] = rng.next();
    if (rand_consumed) *rand_consumed = skip + used;
  }
  int err = 0;
  std::string errmsg;
  // hand_search.cpp:77-80 omp parallel for over samples
#pragma omp parallel for num_threads(threads) schedule(dynamic, 8)
  for (int s = 0; s < S; s++) {
    std::vector<std::pair<float, int>> res;
    const float* q = xyz + 3 * indices[s];
    if (rand_mode) res.swap(res_all[s]);
    else radius_search(tree, xyz, n, q, radius, tree ? 1 : 0, res);  // hand_search.cpp:85
    std::vector<int> nn(res.size());
    for (size_t t = 0; t < res.size(); t++) nn[t] = res[t].second;
    ag_frame F;
    std::memset(&F, 0, sizeof(F));
    if (!nn.empty()) {
      QuadricFit Q = {};
      const double sample[3] = {double(q[0]), double(q[1]), double(q[2])};  // hand_search.cpp:95
      std::vector<double> coords;
      if (sum_perm < 0) {  // extended-precision check solve
        fit_quadric_exact(xyz, nn, q, radius, Q.params, coords);
      } else {
        std::vector<int> nn_sum = nn;
        if (sum_perm > 0) permute(nn_sum, sum_perm);
        if (fit_quadric(xyz, nn_sum, Q) != 0) {
#pragma omp critical
          {
            err = -1;
            errmsg = g_err;
          }
          continue;
        }
        coords.resize(size_t(3) * nn.size());
        for (size_t t = 0; t < nn.size(); t++)
          for (int d = 0; d < 3; d++) coords[3 * t + d] = double(xyz[3 * nn[t] + d]);
      }
      if (rand_mode && nn.size() > 50) {  // quadric.cpp:177-192: 50 picks (repeats allowed) in FLANN order
        std::vector<int> pick_nn(50);
        std::vector<double> pick_xyz(150);
        for (int t = 0; t < 50; t++) {
          const int r = int(stream[stream_off[s] + t] % uint32_t(nn.size()));
          pick_nn[t] = nn[r];
          for (int d = 0; d < 3; d++) pick_xyz[3 * t + d] = coords[3 * r + d];
        }
        local_axes(pick_xyz, cam, pick_nn, Q.params, sample, cam_origin, F);
        F.num_neighbors = int(nn.size());
      } else {
        local_axes(coords, cam, nn, Q.params, sample, cam_origin, F);
      }
      if (params_out) std::memcpy(params_out + size_t(10) * s, Q.params, sizeof(Q.params));
      if (MN_out) {
        std::memcpy(MN_out + size_t(200) * s, Q.M, sizeof(Q.M));
        std::memcpy(MN_out + size_t(200) * s + 100, Q.N, sizeof(Q.N));
      }
      if (eig_out) std::memcpy(eig_out + size_t(10) * s, Q.lambda, sizeof(Q.lambda));
    }
    frames[s] = F;
  }
  if (err) return fail(errmsg);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// A.6 FingerHand — finger_hand.cpp:3-233 (restated over plain arrays)
// ---------------------------------------------------------------------------------------------
struct FingerHand {
  double finger_width, outer_diameter, depth;
  double spacing[20];
  bool fingers[20];
  bool hand[10];
  double back_of_hand = 0;
  const double* px = nullptr;  // rotated slab points (x,y per point)
  const double* py = nullptr;
  int np = 0;
  double surface[2], bottom[2], width;

  FingerHand(double fw, double od, double dp) : finger_width(fw), outer_diameter(od), depth(dp) {
    // finger_hand.cpp:8-15: fs_half = LinSpaced(10, 0, od-fw) (Eigen 3.2: low + i*step)
    const double hi = od - fw, step = (hi - 0.0) / double(10 - 1);
    for (int i = 0; i < 10; i++) {
      double v = 0.0 + double(i) * step;
      spacing[i] = (v - od) + fw;
      spacing[10 + i] = v;
    }
    std::fill(fingers, fingers + 20, false);
    std::fill(hand, hand + 10, false);
  }
  void evaluateFingers(double bite) {  // finger_hand.cpp:20-98
    back_of_hand = -1.0 * (depth - bite);
    std::fill(fingers, fingers + 20, false);
    std::vector<int> crop;
    for (int i = 0; i < np; i++)
      if (py[i] < bite) {
        crop.push_back(i);
        if (py[i] < back_of_hand) return;  // :37-40
      }
    const int m = 20;
    for (int i = 0; i < m; i++) {
      int in_gap = 0;
      for (int j : crop)
        if (px[j] > spacing[i] && px[j] < spacing[i] + finger_width) in_gap++;
      if (in_gap == 0) {
        int sum = 0;
        if (i <= m / 2) {  // :72 (index 10 included)
          for (int j : crop)
            if (px[j] > spacing[i] + finger_width) sum++;
        } else {
          for (int j : crop)
            if (px[j] < spacing[i]) sum++;
        }
        if (sum > 0) fingers[i] = true;
      }
    }
  }
  void evaluateHand() {  // :100-115
    for (int i = 0; i < 10; i++) hand[i] = fingers[i] && fingers[10 + i];
  }
  bool anyHand() const {
    for (bool h : hand)
      if (h) return true;
    return false;
  }
  int deepenHand(double init, double maxd, int* steps_kept) {  // :173-233
    std::vector<int> idx;
    for (int i = 0; i < 10; i++)
      if (hand[i]) idx.push_back(i);
    *steps_kept = 0;
    if (idx.empty()) return -1;
    int e = idx[int(std::ceil(idx.size() / 2.0)) - 1];  // :190
    FingerHand cur = *this, last = *this;
    for (double d = init + 0.005; d <= maxd; d += 0.005) {  // :204
      cur.evaluateFingers(d);
      cur.evaluateHand();
      if (!cur.hand[e]) break;
      last = cur;
      (*steps_kept)++;
    }
    *this = last;  // :229
    for (int i = 0; i < 10; i++) hand[i] = (i == e);
    return e;
  }
  void evaluateGraspParameters(double bite) {  // :117-171
    double fs_sum = 0.0;
    int hsum = 0;
    for (int i = 0; i < 10; i++) {
      fs_sum += spacing[i] * double(hand[i] ? 1 : 0);
      hsum += hand[i] ? 1 : 0;
    }
    double hor = (outer_diameter / 2.0) + (fs_sum / double(hsum));
    double ymax = py[0], ymin = py[0];
    for (int i = 1; i < np; i++) {
      ymax = std::max(ymax, py[i]);
      ymin = std::min(ymin, py[i]);
    }
    bottom[0] = hor;
    bottom[1] = ymax;
    surface[0] = hor;
    surface[1] = ymin;
    std::vector<int> idx;
    for (int i = 0; i < 10; i++)
      if (hand[i]) idx.push_back(i);
    int e = idx[idx.size() / 2];
    double left = spacing[e], right = spacing[10 + e];
    double mx = -100000.0, mn = 100000.0;
    for (int i = 0; i < np; i++)
      if (py[i] < bite && px[i] > left && px[i] < right) {
        if (px[i] < mn) mn = px[i];
        if (px[i] > mx) mx = px[i];
      }
    width = mx - mn;
  }
};

// ---------------------------------------------------------------------------------------------
// A.5, A.7, A.8 — hand_search.cpp:116-206, rotating_hand.cpp:4-177, antipodal.cpp:12-86
// ---------------------------------------------------------------------------------------------
static int antipodal_eval(const std::vector<double>& nrm /*3 x m*/, double th_half, double th_full) {
  // antipodal.cpp:12-86, loops kept literal (the early breaks do not change the outcome)
  const int num_thresh = 6;
  const int m = int(nrm.size() / 3);
  double cos_thresh = std::cos(th_half * M_PI / 180.0);
  int numl = 0, numr = 0;
  bool half = false, full = false;
  for (int i = 0; i < m; i++) {
    const double* v = &nrm[3 * i];
    double ld = (-1.0 * v[0] + 0.0 * v[1]) + 0.0 * v[2];  // l.dot(n), l = (-1,0,0)
    double rd = (1.0 * v[0] + 0.0 * v[1]) + 0.0 * v[2];
    if (ld > cos_thresh) {
      numl++;
      if (numl > num_thresh) { half = true; break; }
    }
    if (rd > cos_thresh) {
      numr++;
      if (numr > num_thresh) { half = true; break; }
    }
  }
  cos_thresh = std::cos(th_full * M_PI / 180.0);
  numl = numr = 0;
  for (int i = 0; i < m; i++) {
    const double* v = &nrm[3 * i];
    double ld = (-1.0 * v[0] + 0.0 * v[1]) + 0.0 * v[2];
    double rd = (1.0 * v[0] + 0.0 * v[1]) + 0.0 * v[2];
    if (ld > cos_thresh) {
      numl++;
      if (numl > num_thresh && numr > num_thresh) { full = true; break; }
    }
    if (rd > cos_thresh) {
      numr++;
      if (numl > num_thresh && numr > num_thresh) { full = true; break; }
    }
  }
  return full ? 2 : (half ? 1 : 0);
}

Hands* find_hands(const float* xyz, const int32_t* cam, int n, const Tree* tree, const int* indices, int S,
                  const ag_frame* frames, const int32_t* sample_cam, const double* cloud_normals,
                  const ag_params& P) {
  Hands* H = new Hands;
  H->status.assign(size_t(S) * 8, 0);
  H->hand_idx.assign(size_t(S) * 8, -1);
  H->depth_steps.assign(size_t(S) * 8, 0);
  H->finger_mask.assign(size_t(S) * 8, 0);
  H->num_slab.assign(size_t(S), 0);
  struct PerSample {
    std::vector<ag_grasp> g;
    std::vector<std::vector<double>> pts;
    std::vector<std::vector<int32_t>> pcam;
  };
  std::vector<PerSample> per(S);
  // rotating_hand.cpp:12-15: first 8 of LinSpaced(9, -pi, pi)  (low + i*step)
  double angles[8];
  {
    const double lo = -1.0 * M_PI, hi = M_PI, step = (hi - lo) / double(9 - 1);
    for (int i = 0; i < 8; i++) angles[i] = lo + double(i) * step;
  }
  const double camL[3] = {P.cam_tf_left[3], P.cam_tf_left[7], P.cam_tf_left[11]};
  const double camR[3] = {P.cam_tf_right[3], P.cam_tf_right[7], P.cam_tf_right[11]};
  int threads = std::max(1, P.num_threads);
  // hand_search.cpp:135-138
#pragma omp parallel for num_threads(threads) schedule(dynamic, 4)
  for (int s = 0; s < S; s++) {
    const float* sf = xyz + 3 * indices[s];  // :141-144 (float sample, exact round trip)
    std::vector<std::pair<float, int>> res;
    radius_search(tree, xyz, n, sf, P.nn_radius_hands, tree ? 1 : 0, res);  // :147
    if (res.empty()) continue;
    const int k0 = int(res.size());
    // :154-160 centred neighbourhood (float subtraction, then cast), normals, cam
    std::vector<double> pc(size_t(3) * k0), nc(size_t(3) * k0, 0.0);
    std::vector<int32_t> cc(k0);
    for (int j = 0; j < k0; j++) {
      int id = res[j].second;
      cc[j] = cam[id];
      for (int d = 0; d < 3; d++) pc[3 * j + d] = double(float(xyz[3 * id + d] - sf[d]));
      if (cloud_normals)
        for (int d = 0; d < 3; d++) nc[3 * j + d] = cloud_normals[size_t(3) * id + d];
    }
    const double se[3] = {double(sf[0]), double(sf[1]), double(sf[2])};  // :164
    double cams[2][3];
    for (int d = 0; d < 3; d++) {
      cams[0][d] = camL[d] - se[d];  // :165-166
      cams[1][d] = camR[d] - se[d];
    }
    // rotating_hand.cpp:19-75 transformPoints
    const double* nor = frames[s].normal;
    const double* axis = frames[s].axis;
    double nxa[3];
    cross3(nor, axis, nxa);
    double Fm[3][3];  // frame = [normal | normal x axis | axis] (columns)
    for (int r = 0; r < 3; r++) {
      Fm[r][0] = nor[r];
      Fm[r][1] = nxa[r];
      Fm[r][2] = axis[r];
    }
    std::vector<double> hx, hy, hz, hn;  // cropped hand-frame points / normals
    std::vector<int32_t> hc;
    for (int j = 0; j < k0; j++) {
      double q[3], qn[3];
      for (int r = 0; r < 3; r++) {  // frame^T * p : sum over k of F(k,r) p_k, left to right
        q[r] = (Fm[0][r] * pc[3 * j] + Fm[1][r] * pc[3 * j + 1]) + Fm[2][r] * pc[3 * j + 2];
        qn[r] = (Fm[0][r] * nc[3 * j] + Fm[1][r] * nc[3 * j + 1]) + Fm[2][r] * nc[3 * j + 2];
      }
      if (q[2] > -1.0 * P.hand_height && q[2] < P.hand_height) {  // :44
        hx.push_back(q[0]);
        hy.push_back(q[1]);
        hz.push_back(q[2]);
        hn.insert(hn.end(), qn, qn + 3);
        hc.push_back(cc[j]);
      }
    }
    const int k = int(hx.size());
    H->num_slab[s] = k;
    // rotating_hand.cpp:78-177 evaluateHand
    std::vector<double> rx(k), ry(k);
    for (int o = 0; o < 8; o++) {
      const double cs = std::cos(angles[o]), sn = std::sin(angles[o]);
      const double rot[3][3] = {{cs, -1.0 * sn, 0.0}, {sn, cs, 0.0}, {0.0, 0.0, 1.0}};  // :90
      for (int j = 0; j < k; j++) {  // :91 rot * points
        rx[j] = (rot[0][0] * hx[j] + rot[0][1] * hy[j]) + rot[0][2] * hz[j];
        ry[j] = (rot[1][0] * hx[j] + rot[1][1] * hy[j]) + rot[1][2] * hz[j];
      }
      // T = frame * rot^T
      double T[3][3];
      for (int r = 0; r < 3; r++)
        for (int c2 = 0; c2 < 3; c2++) T[r][c2] = (Fm[r][0] * rot[c2][0] + Fm[r][1] * rot[c2][1]) + Fm[r][2] * rot[c2][2];
      double approach[3], binormal[3];
      for (int r = 0; r < 3; r++) {
        approach[r] = (T[r][0] * 0.0 + T[r][1] * 1.0) + T[r][2] * 0.0;  // :96
        binormal[r] = (T[r][0] * 1.0 + T[r][1] * 0.0) + T[r][2] * 0.0;  // :104
      }
      if (dot3(approach, cams[0]) > 0 && dot3(approach, cams[1]) > 0) {  // :99
        H->status[s * 8 + o] = 0;
        continue;
      }
      FingerHand fh(P.finger_width, P.hand_outer_diameter, P.hand_depth);
      fh.px = rx.data();
      fh.py = ry.data();
      fh.np = k;
      fh.evaluateFingers(P.init_bite);  // :108
      fh.evaluateHand();
      if (!fh.anyHand()) {  // :111
        H->status[s * 8 + o] = 1;
        continue;
      }
      int steps = 0;
      int e = fh.deepenHand(P.init_bite, fh.depth, &steps);  // :114
      fh.evaluateGraspParameters(P.init_bite);               // :115
      H->status[s * 8 + o] = 2;
      H->hand_idx[s * 8 + o] = e;
      H->depth_steps[s * 8 + o] = steps;
      int mask = 0;
      for (int i = 0; i < 20; i++) mask |= fh.fingers[i] ? (1 << i) : 0;
      H->finger_mask[s * 8 + o] = mask;
      double surface[3], bottom[3];
      for (int r = 0; r < 3; r++) {  // :118-122
        surface[r] = (T[r][0] * fh.surface[0] + T[r][1] * fh.surface[1]) + T[r][2] * 0.0;
        bottom[r] = (T[r][0] * fh.bottom[0] + T[r][1] * fh.bottom[1]) + T[r][2] * 0.0;
      }
      // :125-151 points in box, shifted by the (world-frame) surface vector — quirk kept
      const double lim = fh.back_of_hand + fh.depth;
      std::vector<double> pbox, nbox;
      std::vector<int32_t> cbox;
      for (int j = 0; j < k; j++)
        if (ry[j] < lim) {
          double rz = (rot[2][0] * hx[j] + rot[2][1] * hy[j]) + rot[2][2] * hz[j];
          pbox.push_back(rx[j] - surface[0]);
          pbox.push_back(ry[j] - surface[1]);
          pbox.push_back(rz - surface[2]);
          for (int r = 0; r < 3; r++)  // :92 rot * normals
            nbox.push_back((rot[r][0] * hn[3 * j] + rot[r][1] * hn[3 * j + 1]) + rot[r][2] * hn[3 * j + 2]);
          cbox.push_back(hc[j]);
        }
      ag_grasp g;
      std::memset(&g, 0, sizeof(g));
      for (int r = 0; r < 3; r++) {
        g.axis[r] = axis[r];
        g.approach[r] = approach[r];
        g.binormal[r] = binormal[r];
        g.surface[r] = surface[r] + se[r];  // :153-154
        g.bottom[r] = bottom[r] + se[r];
      }
      g.width = fh.width;
      g.score = std::numeric_limits<float>::quiet_NaN();
      g.sample_index = indices[s];
      g.sample_slot = s;
      g.orientation = o;
      g.cam_source = sample_cam ? sample_cam[s] : 0;
      g.num_points = int(cbox.size());
      int at = antipodal_eval(nbox, 20, 20);  // :159-170
      g.half_antipodal = at >= 1;
      g.full_antipodal = at == 2;
      per[s].g.push_back(g);
      per[s].pts.push_back(std::move(pbox));
      per[s].pcam.push_back(std::move(cbox));
    }
  }
  // hand_search.cpp:194-200 stable concat in sample order
  for (int s = 0; s < S; s++)
    for (size_t t = 0; t < per[s].g.size(); t++) {
      per[s].g[t].image_id = int(H->grasps.size());
      H->grasps.push_back(per[s].g[t]);
      H->pts.push_back(std::move(per[s].pts[t]));
      H->pcam.push_back(std::move(per[s].pcam[t]));
    }
  return H;
}

// A.9 — localization.cpp:364-388
void filter_hands(const ag_grasp* g, int n, const ag_params& P, uint8_t* keep) {
  for (int i = 0; i < n; i++) {
    int k;
    for (k = 0; k < 6; k++)
      if (std::fabs(g[i].surface[int(std::floor(k / 2.0))] - P.workspace[k]) < 0.02) break;
    keep[i] = k == 6;
  }
}

// stratified sorted distinct sample draw (replaces the time-seeded pcl::RandomSample,
// hand_search.cpp:36-39; App. C.2) — shared definition with the product
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
int draw_samples(int n, int S, uint64_t seed, int32_t* out) {
  if (S > n) S = n;  // App. B#4
  for (int k = 0; k < S; k++) {
    int64_t lo = (int64_t(k) * n) / S, hi = (int64_t(k + 1) * n) / S;
    uint64_t h = splitmix64(seed ^ splitmix64(uint64_t(k)));
    out[k] = int32_t(lo + int64_t(h % uint64_t(hi - lo)));
  }
  return S;
}

}  // namespace ago

// =============================================================================================
// C API
// =============================================================================================
using namespace ago;

struct ago_tree {
  Tree t;
};
namespace ago {
const Tree* tree_of(const ago_tree* t) { return t ? &t->t : nullptr; }
}
struct ago_hands {
  Hands* h;
};

extern "C" {

const char* ago_last_error(void) { return g_err.c_str(); }
void ago_free(void* p) { std::free(p); }
int ago_set_lapack(const char* so_path, const char* symbol) { return set_lapack(so_path, symbol); }
int ago_have_lapack(void) { return g_dggev != nullptr; }

int ago_preprocess(const void* points, int stride, int n_in, int size_left, const ag_params* P, int use_std_set,
                   float** xyz_out, int32_t** cam_out, int* n_out) {
  std::vector<float> xyz;
  std::vector<int32_t> cam;
  if (preprocess(points, stride, n_in, size_left, *P, use_std_set != 0, xyz, cam) != 0) return -1;
  *n_out = int(cam.size());
  *xyz_out = static_cast<float*>(std::malloc(std::max<size_t>(1, xyz.size() * sizeof(float))));
  *cam_out = static_cast<int32_t*>(std::malloc(std::max<size_t>(1, cam.size() * sizeof(int32_t))));
  std::memcpy(*xyz_out, xyz.data(), xyz.size() * sizeof(float));
  std::memcpy(*cam_out, cam.data(), cam.size() * sizeof(int32_t));
  return 0;
}

ago_tree* ago_tree_build(const float* xyz, int n) {
  ago_tree* t = new ago_tree;
  t->t.xyz = xyz;
  t->t.n = n;
  t->t.perm.resize(n);
  for (int i = 0; i < n; i++) t->t.perm[i] = i;
  if (n > 0) t->t.build(0, n);
  return t;
}
void ago_tree_free(ago_tree* t) { delete t; }

int ago_radius_search(const ago_tree* t, const float* xyz, int n, const float q[3], double radius, int method,
                      int32_t** idx_out, float** dist_out, int* n_out) {
  std::vector<std::pair<float, int>> res;
  radius_search(t ? &t->t : nullptr, xyz, n, q, radius, method, res);
  *n_out = int(res.size());
  *idx_out = static_cast<int32_t*>(std::malloc(std::max<size_t>(1, res.size() * 4)));
  *dist_out = static_cast<float*>(std::malloc(std::max<size_t>(1, res.size() * 4)));
  for (size_t i = 0; i < res.size(); i++) {
    (*idx_out)[i] = res[i].second;
    (*dist_out)[i] = res[i].first;
  }
  return 0;
}

int ago_fit_quadrics(const float* xyz, const int32_t* cam, int n, const ago_tree* tree, const int* indices,
                     int n_indices, double radius, const ag_params* P, int sum_perm, ag_frame* frames_out,
                     double* params_out, double* MN_out, double* eigvals_out) {
  return fit_quadrics(xyz, cam, n, tree ? &tree->t : nullptr, indices, n_indices, radius, *P, sum_perm, frames_out,
                      params_out, MN_out, eigvals_out);
}

ago_hands* ago_find_hands(const float* xyz, const int32_t* cam, int n, const ago_tree* tree, const int* indices,
                          int n_indices, const ag_frame* frames, const int32_t* sample_cam,
                          const double* cloud_normals, const ag_params* P) {
  ago_hands* h = new ago_hands;
  h->h = find_hands(xyz, cam, n, tree ? &tree->t : nullptr, indices, n_indices, frames, sample_cam, cloud_normals, *P);
  h->h->n_voxels = n;
  return h;
}
void ago_hands_free(ago_hands* h) {
  if (!h) return;
  delete h->h;
  delete h;
}
int ago_hands_count(const ago_hands* h) { return int(h->h->grasps.size()); }
const ag_grasp* ago_hands_grasps(const ago_hands* h) { return h->h->grasps.data(); }
int ago_hands_points(const ago_hands* h, int k, const double** pts, const int32_t** cam, int* m) {
  if (k < 0 || k >= int(h->h->grasps.size())) return fail("hypothesis index out of range");
  *pts = h->h->pts[k].data();
  *cam = h->h->pcam[k].data();
  *m = int(h->h->pcam[k].size());
  return 0;
}
int ago_hands_debug(const ago_hands* h, const int32_t** status, const int32_t** hand_idx,
                    const int32_t** depth_steps, const int32_t** finger_mask, const int32_t** num_slab) {
  *status = h->h->status.data();
  *hand_idx = h->h->hand_idx.data();
  *depth_steps = h->h->depth_steps.data();
  *finger_mask = h->h->finger_mask.data();
  *num_slab = h->h->num_slab.data();
  return 0;
}
int ago_filter_hands(const ag_grasp* grasps, int n, const ag_params* P, uint8_t* keep) {
  filter_hands(grasps, n, *P, keep);
  return 0;
}
int ago_glibc_rand(uint32_t seed, int n, int32_t* out) {
  GlibcRand g(seed);
  for (int i = 0; i < n; i++) out[i] = int32_t(g.next());
  return 0;
}
int ago_draw_samples(int n, int num_samples, uint64_t seed, int32_t* out) {
  return draw_samples(n, num_samples, seed, out);
}

}  // extern "C"
