// quadric.cu — Taubin quadric fit and local frame, one warp per sample.
//
// Replaces (reference paths): HandSearch::findQuadrics (src/agile_grasp/hand_search.cpp:65-113),
// pcl::KdTreeFLANN::radiusSearch (:85), Quadric::fitQuadric (src/agile_grasp/quadric.cpp:14-157, LAPACK
// dggev_ :330-363), Quadric::findTaubinNormalAxis (:159-251) and findAverageNormalAxis (:263-305).
//
// GPU formulation:
//  * k_taubin_moments: each warp walks the <=3x3 z-contiguous cell columns that cover its sample's
//    ball, tests membership with FLANN's exact binary32 arithmetic, compacts accepted points through a
//    64-entry shared-memory ring and accumulates the 35 monomial moments of degree <=4 in binary64
//    registers — in coordinates centred on the sample and scaled by 1/r (Taubin's fit is invariant
//    under translation and uniform scale, and the centred problem is well conditioned whereas the
//    reference's uncentred 10x10 pencil is not).  Warp-shuffle reduction, 36 doubles out per sample.
//  * k_taubin_axes: builds M and N from the moments, eliminates the constant term
//    (A - m m^T/n) u = lambda B u, Cholesky + cyclic Jacobi on the 9x9 symmetric-definite pencil in
//    shared memory, then two more ball walks for the gradient normals: one accumulates sum g g^T and
//    the symmetric order-6 moment tensor T = sum g^(x6) (28 numbers), the other evaluates
//    sum_i (g_i . g_j)^6 = <T, g_j^(x6)> for every j — O(m) instead of the reference's O(m^2) — and
//    takes the arg max with the reference's first-max tie-break in (distance, index) order.

#include <cstdlib>

#include "ag_internal.h"

namespace ag {

namespace {

constexpr int kWarps = 4;  // warps per CTA

// ---- monomial bookkeeping ---------------------------------------------------------------------
// index of the moment sum x^a y^b z^c, a+b+c <= 4
__host__ __device__ constexpr int midx(int a, int b, int c) {
  const int k = a * 25 + b * 5 + c;
  return k == 0 ? 0 : k == 25 ? 1 : k == 5 ? 2 : k == 1 ? 3 : k == 50 ? 4 : k == 10 ? 5 : k == 2 ? 6 : k == 30 ? 7
       : k == 6 ? 8 : k == 26 ? 9 : k == 75 ? 10 : k == 15 ? 11 : k == 3 ? 12 : k == 55 ? 13 : k == 51 ? 14
       : k == 35 ? 15 : k == 11 ? 16 : k == 27 ? 17 : k == 7 ? 18 : k == 31 ? 19 : k == 100 ? 20 : k == 20 ? 21
       : k == 4 ? 22 : k == 80 ? 23 : k == 76 ? 24 : k == 40 ? 25 : k == 16 ? 26 : k == 28 ? 27 : k == 8 ? 28
       : k == 60 ? 29 : k == 12 ? 30 : k == 52 ? 31 : k == 56 ? 32 : k == 36 ? 33 : k == 32 ? 34 : -1;
}
// device-side lookup of the same map (index a*25 + b*5 + c), filled from midx() at compile time
struct MidxTable {
  signed char v[125];
  constexpr MidxTable() : v() {
    for (int a = 0; a < 5; a++)
      for (int b = 0; b < 5; b++)
        for (int c = 0; c < 5; c++) v[a * 25 + b * 5 + c] = (a + b + c <= 4) ? static_cast<signed char>(midx(a, b, c)) : -1;
  }
};
__constant__ MidxTable c_midx = MidxTable();
__device__ __forceinline__ int midx_d(int a, int b, int c) { return c_midx.v[a * 25 + b * 5 + c]; }
constexpr int kNumMoments = 35;
constexpr int kMomentStride = 36;  // + count of camera-1 neighbours

// exponents of the quadric basis [x2 y2 z2 xy yz xz x y z 1] (quadric.cpp:40-73)
__constant__ int c_basis[10][3] = {{2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {1, 1, 0}, {0, 1, 1},
                                   {1, 0, 1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};

// ---- warp-level ball walk with compaction ----------------------------------------------------
// The query's candidate runs (one per x-row, see ag_common.cuh) are flattened into one index space
// and streamed 64 candidates at a time (two independent 16-byte loads in flight per lane); accepted
// points are compacted through a 64-entry shared ring so that f(point, active) always sees full
// batches of 32 (except the last).  The ring entry's tag is (point index << 2) | (tag bits).
constexpr int kRunCap = 128;
struct RunList {
  int rs[kRunCap];
  int pre[kRunCap + 1];
};

template <typename F>
__device__ __forceinline__ void walk_runs(const GPoint* __restrict__ pts, const RunList& rl, int n_runs, float qx,
                                          float qy, float qz, float r2, GPoint* ring, F&& f) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const int total = rl.pre[n_runs];
  int head = 0, qn = 0, cur = 0;
  auto push = [&](const GPoint& p, bool ok) {
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (m == 0) return;
    if (ok) ring[(head + qn + __popc(m & lt)) & 63] = p;
    qn += __popc(m);
    __syncwarp();
    if (qn >= 32) {
      const GPoint v = ring[(head + lane) & 63];
      __syncwarp();
      f(v, true);
      head = (head + 32) & 63;
      qn -= 32;
    }
  };
  for (int base = 0; base < total; base += 64) {
    GPoint p[2];
    bool valid[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int pos = base + u * 32 + lane;
      valid[u] = pos < total;
      p[u].x = p[u].y = p[u].z = 0.f;
      p[u].tag = 0;
      if (valid[u]) {
        while (pos >= rl.pre[cur + 1]) cur++;
        const int j = rl.rs[cur] + (pos - rl.pre[cur]);
        p[u] = pts[j];
        p[u].tag = (uint32_t(j) << 2) | (p[u].tag & 3u);
      }
    }
    push(p[0], valid[0] && dist2_flann(qx, qy, qz, p[0].x, p[0].y, p[0].z) < r2);
    if (base + 32 < total) push(p[1], valid[1] && dist2_flann(qx, qy, qz, p[1].x, p[1].y, p[1].z) < r2);
  }
  if (qn > 0) {
    const GPoint v = ring[(head + lane) & 63];
    __syncwarp();
    f(v, lane < qn);
  }
  __syncwarp();
}

// ---- kernel 1: moments ------------------------------------------------------------------------
// Roofline-graded kernel.  One warp per sample, four independent warps per CTA.  Per sample:
//   1. lane r finds the candidate run of x-row r of the ball (two independent loads from the column table);
//      a run is contiguous in the voxel list, so
//   2. each row lane issues ONE bulk async copy (TMA, cp.async.bulk) of its run into the warp's shared
//      staging buffer at the run's prefix offset; one mbarrier transaction count tracks all of them — every
//      load of the ball is in flight at once, no registers are held for them;
//   3. the staged candidates are tested 32 per step with FLANN's exact binary32 expression and compacted
//      IN PLACE (the write cursor never passes the read cursor);
//   4. full 32-point batches of accepted points are accumulated into the 35 monomial moments in binary64
//      registers, in coordinates centred on the sample (exact: differences of binary32 values); a remainder
//      of < 32 points is carried to the front of the buffer for the next pass (balls with more candidates
//      than the buffer holds, or more than 32 rows, take several passes);
//   5. the 32 x 35 partial sums are transposed through the same buffer (lane i sums moment i), the
//      power-of-two coordinate scale is applied exactly (moment of degree d times scale^d), 36 doubles out.
constexpr int kCandCap = 560;               // staging buffer entries per warp (8960 B = 35 x 32 doubles)
constexpr int kCandNew = kCandCap - 32;     // new candidates per pass (the rest holds the carried remainder)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16), completion counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// explicit shared-space 16-byte accesses (one 32-bit address register, immediate offsets)
__device__ __forceinline__ GPoint lds_point(uint32_t addr) {
  GPoint p;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=r"(p.tag)
               : "r"(addr)
               : "memory");
  return p;
}
__device__ __forceinline__ void sts_point(uint32_t addr, const GPoint& p) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(p.x), "f"(p.y), "f"(p.z), "r"(p.tag)
               : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef AG_MOM_MINB
#define AG_MOM_MINB 4
#endif
__global__ void __launch_bounds__(kWarps * 32, AG_MOM_MINB)
k_taubin_moments(const GPoint* __restrict__ pts, const int* __restrict__ row_ptr, const int* __restrict__ col_ptr,
                 const RowIndex* __restrict__ rip,
                 const int* __restrict__ indices, int n_samples_max, const int* __restrict__ d_count, float r2,
                 double rpad, double scale, double* __restrict__ moments, int2* __restrict__ nn_counts) {
  __shared__ __align__(16) GPoint s_cand[kWarps][kCandCap];
  static_assert(kCandCap * sizeof(GPoint) >= 32 * 33 * sizeof(double), "the reduction reuses the staging buffer");
  __shared__ __align__(8) unsigned long long s_bar[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * kWarps + warp;
  // the index descriptor is read in place (L1 hits) instead of being held in ~30 registers per thread
  const RowIndex& ri = *rip;
  if (s >= n_samples_max || s >= *d_count) return;
  const int idx = indices[s];
  if (idx < 0 || idx >= ri.n_points) return;
  const GPoint q = pts[idx];
  const double qx = double(q.x), qy = double(q.y), qz = double(q.z);
  uint32_t cand_s, bar;  // opaque to the compiler so the addresses stay in registers instead of being rebuilt
  asm volatile("mov.u32 %0, %1;" : "=r"(cand_s) : "r"(smem_u32(s_cand[warp])));
  asm volatile("mov.u32 %0, %1;" : "=r"(bar) : "r"(smem_u32(&s_bar[warp])));
  const uint32_t lane_s = cand_s + uint32_t(lane) * 16u;  // this lane's slot of a 32-point batch
  if (lane == 0) mbar_init(bar, 1);
  fence_proxy_async_smem();  // the initialised barrier must be visible to the async proxy
  __syncwarp();
  uint32_t parity = 0;
  double acc[kNumMoments];  // acc[0] (the count) is filled from the integer counter at the end
#pragma unroll
  for (int i = 0; i < kNumMoments; i++) acc[i] = 0.0;
  int cam1 = 0, n_cand = 0, n_acc = 0, carry = 0;
  const unsigned lt = (1u << lane) - 1u;
  auto process = [&](const GPoint& p) {
    const double x = double(p.x) - qx, y = double(p.y) - qy, z = double(p.z) - qz;
    const double x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, yz = y * z, xz = x * z;
    acc[1] += x; acc[2] += y; acc[3] += z;
    acc[4] += x2; acc[5] += y2; acc[6] += z2; acc[7] += xy; acc[8] += yz; acc[9] += xz;
    acc[10] += x2 * x; acc[11] += y2 * y; acc[12] += z2 * z; acc[13] += x2 * y; acc[14] += x2 * z;
    acc[15] += x * y2; acc[16] += y2 * z; acc[17] += x * z2; acc[18] += y * z2; acc[19] += xy * z;
    acc[20] += x2 * x2; acc[21] += y2 * y2; acc[22] += z2 * z2; acc[23] += x2 * xy; acc[24] += x2 * xz;
    acc[25] += xy * y2; acc[26] += y2 * yz; acc[27] += xz * z2; acc[28] += yz * z2; acc[29] += x2 * y2;
    acc[30] += y2 * z2; acc[31] += x2 * z2; acc[32] += x2 * yz; acc[33] += xy * yz; acc[34] += xz * yz;
    cam1 += int(p.tag & kTagCamBit);
  };
  // a point that fails the radius test and contributes zero to every sum
  GPoint far_pt;
  far_pt.x = 1e30f;
  far_pt.y = far_pt.z = 0.f;
  far_pt.tag = 0;
  // one loop over (camera, batch of 32 x-rows): one batch per camera for the shipped radii
  int k_lo0 = 0, k_hi0 = -1, k_lo1 = 0, k_hi1 = -1;
  if (ri.count[0] > 0) row_range(ri, 0, q.x, rpad, k_lo0, k_hi0);
  if (ri.count[1] > 0) row_range(ri, 1, q.x, rpad, k_lo1, k_hi1);
  const int nb0 = k_hi0 >= k_lo0 ? (k_hi0 - k_lo0) / 32 + 1 : 0;
  const int nb1 = k_hi1 >= k_lo1 ? (k_hi1 - k_lo1) / 32 + 1 : 0;
  for (int t = 0; t < nb0 + nb1; t++) {
    {
      const int c = t < nb0 ? 0 : 1;
      const int kb = c ? k_lo1 + (t - nb0) * 32 : k_lo0 + t * 32;
      const int k_hi = c ? k_hi1 : k_hi0;
      const int nrows = min(32, k_hi - kb + 1);
      int j_lo = 0, j_end = 0;
      if (lane < nrows) row_run(ri, row_ptr, col_ptr, pts, c, kb + lane, q.x, q.y, rpad, j_lo, j_end);
      int rem = max(0, j_end - j_lo);
      n_cand += rem;
      while (true) {
        int incl = rem;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        const int total_rem = __shfl_sync(0xffffffffu, incl, 31);
        if (total_rem == 0) break;
        const int excl = incl - rem;
        const int take = min(rem, max(0, kCandNew - excl));
        const int total = min(total_rem, kCandNew);
        fence_proxy_async_smem();  // earlier generic-proxy accesses of the buffer precede the async writes
        __syncwarp();
        if (lane == 0) mbar_expect_tx(bar, uint32_t(total) * 16u);
        __syncwarp();
        if (take > 0) bulk_g2s(cand_s + uint32_t(carry + excl) * 16u, pts + j_lo, uint32_t(take) * 16u, bar);
        j_lo += take;
        rem -= take;
        mbar_wait(bar, parity);
        parity ^= 1u;
        // membership test + in-place compaction, 64 candidates per step (two independent chains)
        const int end = carry + total;
        int w = carry;
        for (int base = carry; base < end; base += 64) {
          const int i0 = base + lane, i1 = i0 + 32;
          GPoint p0 = far_pt, p1 = far_pt;
          if (i0 < end) p0 = lds_point(lane_s + uint32_t(base) * 16u);
          if (i1 < end) p1 = lds_point(lane_s + uint32_t(base) * 16u + 512u);
          const bool ok0 = dist2_flann(q.x, q.y, q.z, p0.x, p0.y, p0.z) < r2;
          const bool ok1 = dist2_flann(q.x, q.y, q.z, p1.x, p1.y, p1.z) < r2;
          const unsigned m0 = __ballot_sync(0xffffffffu, ok0), m1 = __ballot_sync(0xffffffffu, ok1);
          const int w1 = w + __popc(m0);
          if (ok0) sts_point(cand_s + uint32_t(w + __popc(m0 & lt)) * 16u, p0);
          if (ok1) sts_point(cand_s + uint32_t(w1 + __popc(m1 & lt)) * 16u, p1);
          w = w1 + __popc(m1);
        }
        __syncwarp();
        n_acc += w - carry;
        const int nfull = w & ~31;
        for (int b = 0; b < nfull; b += 32) process(lds_point(lane_s + uint32_t(b) * 16u));
        carry = w - nfull;
        if (nfull > 0 && carry > 0) {  // move the remainder to the front
          GPoint v = far_pt;
          if (lane < carry) v = lds_point(lane_s + uint32_t(nfull) * 16u);
          __syncwarp();
          if (lane < carry) sts_point(lane_s, v);
        }
        __syncwarp();
      }
    }
  }
  if (carry > 0) {  // inactive lanes process the sample itself: every term is exactly zero
    GPoint v = q;
    v.tag = 0;
    if (lane < carry) v = lds_point(lane_s);
    process(v);
  }
  // warp reduction: moments 1..32 transposed through shared memory with a 33-double row pitch (lane i
  // sums moment i + 1 with immediate offsets, bank-conflict free), 33..34 by butterflies
  __syncwarp();
#pragma unroll
  for (int m = 0; m < 32; m++) sts_f64(cand_s + uint32_t(m * 33 + 0) * 8u + uint32_t(lane) * 8u, acc[m + 1]);
  const double m33 = warp_sum(acc[33]), m34 = warp_sum(acc[34]);
  cam1 = __reduce_add_sync(0xffffffffu, cam1);
  n_cand = __reduce_add_sync(0xffffffffu, n_cand);
  __syncwarp();
  double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
  const uint32_t row_s = cand_s + uint32_t(lane) * (33u * 8u);
#pragma unroll
  for (int k = 0; k < 32; k += 4) {
    t0 += lds_f64(row_s + uint32_t(k) * 8u);
    t1 += lds_f64(row_s + uint32_t(k + 1) * 8u);
    t2 += lds_f64(row_s + uint32_t(k + 2) * 8u);
    t3 += lds_f64(row_s + uint32_t(k + 3) * 8u);
  }
  const double rsum = (t0 + t1) + (t2 + t3);                       // moment lane + 1
  double r32 = __shfl_up_sync(0xffffffffu, rsum, 1);               // lane i (i >= 1): moment i
  const double m32 = __shfl_sync(0xffffffffu, rsum, 31);           // moment 32
  if (lane == 0) r32 = double(n_acc);                              // moment 0: the neighbour count
  // coordinate scale (a power of two, so this equals accumulating scaled coordinates bit for bit)
  const double s2 = scale * scale, s4 = s2 * s2;
  const double scl = lane == 0 ? 1.0 : lane <= 3 ? scale : lane <= 9 ? s2 : lane <= 19 ? s2 * scale : s4;
  double* out = moments + size_t(s) * kMomentStride;
  out[lane] = r32 * scl;
  if (lane == 0) {
    out[32] = m32 * s4;
    out[33] = m33 * s4;
    out[34] = m34 * s4;
    out[35] = double(cam1);
    nn_counts[s] = make_int2(int(r32), n_cand);
  }
}

// ---- small dense helpers (per warp, shared memory, lane-parallel where it is cheap) -----------
struct AxesSmem {
  double A[81];   // Schur complement, later C = L^-1 A L^-T, diagonalised in place
  double L[81];   // B, then its Cholesky factor
  double V[81];   // Jacobi eigenvectors
  double m[10];   // last column of M (9 entries) and n
  double par[10]; // quadric parameters in centred/scaled coordinates
  double T[28];   // weighted order-6 normal tensor
  GPoint ring[64];
  RunList runs;
};

// Jacobi rotation (c, s) annihilating a_pq, the small-angle root (|angle| <= pi/4), computed without
// divisions: with h = (a_qq - a_pp)/2, b = a_pq, r = hypot(h, b):  cos 2phi = |h|/r, sin 2phi = b/r,
// c = sqrt((1 + cos 2phi)/2), s = sign(h) * sin 2phi / (2c)   — two rsqrt, no div/sqrt sequences.
__device__ __forceinline__ void jacobi_rotate_params(double app, double aqq, double apq, double& c, double& s) {
  if (apq == 0.0) {
    c = 1.0;
    s = 0.0;
    return;
  }
  const double h = 0.5 * (aqq - app);
  const double rr = h * h + apq * apq;
  if (!(rr > 1e-300) || !(rr < 1e300)) {  // squares under/overflow: entries this small/large are not rotated
    c = 1.0;
    s = 0.0;
    return;
  }
  const double inv_r = rsqrt(rr);
  const double c2 = fabs(h) * inv_r;        // cos 2phi in [0, 1]
  const double s2 = apq * inv_r;            // sin 2phi (sign of a_pq)
  const double y = 0.5 * (1.0 + c2);        // cos^2 phi in [1/2, 1]
  const double ry = rsqrt(y);
  c = y * ry;
  s = (h >= 0.0 ? 0.5 : -0.5) * s2 * ry;
}

// Parallel-ordered cyclic Jacobi on the 9x9 symmetric matrix A (row-major) with eigenvectors in V,
// executed by one warp.  Round-robin ("circle") ordering over 10 players — index 9 is a bye — gives 9
// rounds of 4 DISJOINT rotations per sweep; disjoint rotations commute, so each round applies
// J = J_0 J_1 J_2 J_3 as A <- A J (36 column work items) then A <- J^T A (36 row work items).
// Rotations are skipped when |a_pq| <= eps*sqrt(a_pp*a_qq) — the scaled criterion that gives
// positive-definite matrices high RELATIVE accuracy in their small eigenpairs (the one we need is the
// smallest); a sweep without rotations ends the iteration.
__device__ void warp_jacobi9(double* A, double* V, int lane) {
  constexpr int N = 9;
  for (int i = lane; i < N * N; i += 32) V[i] = (i / N == i % N) ? 1.0 : 0.0;
  __syncwarp();
  const int t_item0 = lane / N, k_item0 = lane % N;            // work item `lane`      (t in 0..3)
  const int t_item1 = (lane + 32) / N, k_item1 = (lane + 32) % N;  // work item `lane+32` (lanes 0..3 only)
  for (int sweep = 0; sweep < 30; sweep++) {
    bool rotated = false;
    for (int r = 0; r < N; r++) {
      // lanes 0..3 own pair t = lane of this round
      int p = (r + (lane & 3) + 1) % N, q = (r - (lane & 3) - 1 + 2 * N) % N;
      if (p > q) { const int tmp = p; p = q; q = tmp; }
      double c = 1.0, s = 0.0;
      bool rot = false;
      if (lane < 4) {
        const double app = A[p * N + p], aqq = A[q * N + q], apq = A[p * N + q];
        rot = apq * apq > 1.0e-32 * fabs(app * aqq);
        if (rot) jacobi_rotate_params(app, aqq, apq, c, s);
      }
      const unsigned any = __ballot_sync(0xffffffffu, rot);
      if (any == 0) continue;  // uniform
      rotated = true;
      // broadcast the four rotations
      const int p0 = __shfl_sync(0xffffffffu, p, t_item0), q0 = __shfl_sync(0xffffffffu, q, t_item0);
      const double c0 = __shfl_sync(0xffffffffu, c, t_item0), s0 = __shfl_sync(0xffffffffu, s, t_item0);
      const int p1 = __shfl_sync(0xffffffffu, p, t_item1 & 3), q1 = __shfl_sync(0xffffffffu, q, t_item1 & 3);
      const double c1 = __shfl_sync(0xffffffffu, c, t_item1 & 3), s1 = __shfl_sync(0xffffffffu, s, t_item1 & 3);
      // columns: A <- A J, V <- V J
      {
        const int k = k_item0;
        const double akp = A[k * N + p0], akq = A[k * N + q0], vkp = V[k * N + p0], vkq = V[k * N + q0];
        A[k * N + p0] = c0 * akp - s0 * akq;
        A[k * N + q0] = s0 * akp + c0 * akq;
        V[k * N + p0] = c0 * vkp - s0 * vkq;
        V[k * N + q0] = s0 * vkp + c0 * vkq;
      }
      if (lane < 4) {
        const int k = k_item1;
        const double akp = A[k * N + p1], akq = A[k * N + q1], vkp = V[k * N + p1], vkq = V[k * N + q1];
        A[k * N + p1] = c1 * akp - s1 * akq;
        A[k * N + q1] = s1 * akp + c1 * akq;
        V[k * N + p1] = c1 * vkp - s1 * vkq;
        V[k * N + q1] = s1 * vkp + c1 * vkq;
      }
      __syncwarp();
      // rows: A <- J^T A
      {
        const int k = k_item0;
        const double apk = A[p0 * N + k], aqk = A[q0 * N + k];
        A[p0 * N + k] = c0 * apk - s0 * aqk;
        A[q0 * N + k] = s0 * apk + c0 * aqk;
      }
      if (lane < 4) {
        const int k = k_item1;
        const double apk = A[p1 * N + k], aqk = A[q1 * N + k];
        A[p1 * N + k] = c1 * apk - s1 * aqk;
        A[q1 * N + k] = s1 * apk + c1 * aqk;
      }
      __syncwarp();
      if (lane < 4 && rot) A[p * N + q] = A[q * N + p] = 0.0;  // annihilated by construction
      __syncwarp();
    }
    if (!rotated) break;
  }
}

// symmetric 3x3 eigen-decomposition in registers (every lane redundantly)
__device__ void eig3(const double Cm[6] /*xx yy zz xy yz xz*/, double w[3], double V[3][3]) {
  double a[3][3] = {{Cm[0], Cm[3], Cm[5]}, {Cm[3], Cm[1], Cm[4]}, {Cm[5], Cm[4], Cm[2]}};
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 40; sweep++) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    const double diag = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
    if (off <= 1e-34 * diag || off == 0.0) break;
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
      for (int q = p + 1; q < 3; q++) {
        double c, s;
        jacobi_rotate_params(a[p][p], a[q][q], a[p][q], c, s);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  w[0] = a[0][0];
  w[1] = a[1][1];
  w[2] = a[2][2];
}

// gradient direction of the implicit quadric at (x,y,z)   (quadric.cpp:238-247)
__device__ __forceinline__ void quad_normal(const double* par, double x, double y, double z, double g[3]) {
  const double fx = 2.0 * par[0] * x + par[3] * y + par[5] * z + par[6];
  const double fy = 2.0 * par[1] * y + par[3] * x + par[4] * z + par[7];
  const double fz = 2.0 * par[2] * z + par[4] * y + par[5] * x + par[8];
  const double inv = 1.0 / sqrt(fx * fx + fy * fy + fz * fz);
  g[0] = fx * inv;
  g[1] = fy * inv;
  g[2] = fz * inv;
}

// the 28 monomials g^alpha, |alpha| = 6, in a fixed order (a descending, then b descending)
__device__ __forceinline__ void monomials6(const double g[3], double out[28]) {
  double px[7], py[7], pz[7];
  px[0] = py[0] = pz[0] = 1.0;
#pragma unroll
  for (int k = 1; k <= 6; k++) {
    px[k] = px[k - 1] * g[0];
    py[k] = py[k - 1] * g[1];
    pz[k] = pz[k - 1] * g[2];
  }
  int t = 0;
#pragma unroll
  for (int a = 6; a >= 0; a--)
#pragma unroll
    for (int b = 6 - a; b >= 0; b--) out[t++] = px[a] * py[b] * pz[6 - a - b];
}
__constant__ double c_multinomial6[28] = {
    1,                          // a=6
    6, 6,                       // a=5: b=1,0
    15, 30, 15,                 // a=4: b=2,1,0
    20, 60, 60, 20,             // a=3
    15, 60, 90, 60, 15,         // a=2
    6, 30, 60, 60, 30, 6,       // a=1
    1, 6, 15, 20, 15, 6, 1};    // a=0

// ---- kernel 2: eigen-solve + local axes -------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32, 4)
k_taubin_axes(const GPoint* __restrict__ pts_c, const int* __restrict__ row_ptr, const int* __restrict__ col_ptr,
              const RowIndex* __restrict__ rip,
              const int* __restrict__ indices, int n_samples_max, const int* __restrict__ d_count,
              float r2, double rpad, double inv_r, const double* __restrict__ moments,
              double cam0x, double cam0y, double cam0z, double cam1x, double cam1y, double cam1z,
              ag_frame* __restrict__ frames, double* normals_out /* may be null */) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  AxesSmem* sm_all = reinterpret_cast<AxesSmem*>(s_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * kWarps + warp;
  const RowIndex ri = *rip;
  if (s >= n_samples_max || s >= *d_count) return;
  AxesSmem& sm = sm_all[warp];
  const int idx = indices[s];
  if (idx < 0 || idx >= ri.n_points) return;
  const GPoint q = pts_c[idx];
  const double qx = double(q.x), qy = double(q.y), qz = double(q.z);
  const double* mom = moments + size_t(s) * kMomentStride;
  const double n = mom[0];
  ag_frame F;
  F.num_neighbors = int(n);
  const int major = (mom[35] > n - mom[35]) ? 1 : 0;  // quadric.cpp:217-226 (tie -> camera 0)
  F.majority_cam = major;

  // --- build the reduced pencil: A = M[0:9,0:9] - m m^T / n, B = N[0:9,0:9]
  for (int e = lane; e < 81; e += 32) {
    const int i = e / 9, j = e % 9;
    const int* bi = c_basis[i];
    const int* bj = c_basis[j];
    const double mij = mom[midx_d(bi[0] + bj[0], bi[1] + bj[1], bi[2] + bj[2])];
    const double mi = mom[midx_d(bi[0], bi[1], bi[2])], mj = mom[midx_d(bj[0], bj[1], bj[2])];
    sm.A[e] = mij - mi * mj / n;
    double bsum = 0.0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      if (bi[a] >= 1 && bj[a] >= 1) {
        int ex[3] = {bi[0] + bj[0], bi[1] + bj[1], bi[2] + bj[2]};
        ex[a] -= 2;
        bsum += double(bi[a] * bj[a]) * mom[midx_d(ex[0], ex[1], ex[2])];
      }
    }
    sm.L[e] = bsum;
  }
  if (lane < 9) sm.m[lane] = mom[midx_d(c_basis[lane][0], c_basis[lane][1], c_basis[lane][2])];
  __syncwarp();

  // --- Cholesky B = L L^T (lane-parallel over rows below the pivot).  B is singular when the
  // neighbourhood is degenerate for the gradient form (e.g. voxel corners lying exactly in one lattice
  // plane); those directions have lambda = infinity (or 0/0) and must be excluded, which the robust
  // branch below does by working in range(B).
  bool ok = n >= 1.0;
  bool singular = false;
  double dmax = 0.0;
  for (int i = 0; i < 9; i++) dmax = fmax(dmax, sm.L[i * 9 + i]);
  if (!(dmax > 0.0)) ok = false;
  {
    const double tol = 1e-10 * dmax;
    for (int k = 0; k < 9 && ok; k++) {
      const double piv = sm.L[k * 9 + k];
      if (!(piv > tol)) {  // uniform across the warp (shared memory value)
        singular = true;
        break;
      }
      const double d = sqrt(piv);
      __syncwarp();
      if (lane == 0) sm.L[k * 9 + k] = d;
      if (lane > k && lane < 9) sm.L[lane * 9 + k] /= d;
      __syncwarp();
      // trailing update: entries (i,j), k < j <= i < 9
      for (int e = lane; e < 81; e += 32) {
        const int i = e / 9, j = e % 9;
        if (j > k && i >= j) sm.L[i * 9 + j] -= sm.L[i * 9 + k] * sm.L[j * 9 + k];
      }
      __syncwarp();
    }
  }
  if (!singular) {
    // --- C = L^-1 A L^-T : X = L^-1 A (columns in parallel), then C^T = L^-1 X^T
    if (lane < 9) {
      const int col = lane;
      for (int i = 0; i < 9; i++) {
        double v = sm.A[i * 9 + col];
        for (int k = 0; k < i; k++) v -= sm.L[i * 9 + k] * sm.A[k * 9 + col];
        sm.A[i * 9 + col] = v / sm.L[i * 9 + i];
      }
    }
    __syncwarp();
    if (lane < 9) {
      const int row = lane;  // solve L y = (row of X)^T
      for (int i = 0; i < 9; i++) {
        double v = sm.A[row * 9 + i];
        for (int k = 0; k < i; k++) v -= sm.L[i * 9 + k] * sm.A[row * 9 + k];
        sm.A[row * 9 + i] = v / sm.L[i * 9 + i];
      }
    }
    __syncwarp();
    // symmetrise (round-off) and diagonalise
    for (int e = lane; e < 81; e += 32) {
      const int i = e / 9, j = e % 9;
      if (i < j) {
        const double v = 0.5 * (sm.A[i * 9 + j] + sm.A[j * 9 + i]);
        sm.A[i * 9 + j] = v;
        sm.A[j * 9 + i] = v;
      }
    }
    __syncwarp();
    warp_jacobi9(sm.A, sm.V, lane);
    // smallest eigenvalue (quadric.cpp:149-152), u = L^-T y, j = -m.u/n
    int mi = 0;
    for (int k = 1; k < 9; k++)
      if (sm.A[k * 9 + k] < sm.A[mi * 9 + mi]) mi = k;
    __syncwarp();
    if (lane == 0) {
      double u[9];
      for (int i = 8; i >= 0; i--) {
        double v = sm.V[i * 9 + mi];
        for (int k = i + 1; k < 9; k++) v -= sm.L[k * 9 + i] * u[k];
        u[i] = v / sm.L[i * 9 + i];
      }
      for (int i = 0; i < 9; i++) sm.par[i] = u[i];
    }
  } else {
    // --- robust branch: B = Q D Q^T, W = Q_k D_k^(-1/2) over the directions with D_i > 1e-11 D_max,
    //     C = W^T A W on range(B), smallest eigenpair y, u = W y
    for (int e = lane; e < 81; e += 32) {  // rebuild B (the partial Cholesky overwrote it)
      const int i = e / 9, j = e % 9;
      const int* bi = c_basis[i];
      const int* bj = c_basis[j];
      double bsum = 0.0;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        if (bi[a] >= 1 && bj[a] >= 1) {
          int ex[3] = {bi[0] + bj[0], bi[1] + bj[1], bi[2] + bj[2]};
          ex[a] -= 2;
          bsum += double(bi[a] * bj[a]) * mom[midx_d(ex[0], ex[1], ex[2])];
        }
      }
      sm.L[e] = bsum;
    }
    __syncwarp();
    warp_jacobi9(sm.L, sm.V, lane);  // eigenvalues on diag(sm.L), eigenvectors in the columns of sm.V
    double d_max = 0.0;
    for (int i = 0; i < 9; i++) d_max = fmax(d_max, sm.L[i * 9 + i]);
    __syncwarp();
    if (lane < 9) {  // scale column `lane` of Q
      const double d = sm.L[lane * 9 + lane];
      const double sc = d > 1e-11 * d_max ? 1.0 / sqrt(d) : 0.0;
      for (int i = 0; i < 9; i++) sm.V[i * 9 + lane] *= sc;
    }
    __syncwarp();
    for (int e = lane; e < 81; e += 32) {  // T = A W  -> sm.L
      const int i = e / 9, j = e % 9;
      double v = 0.0;
      for (int k = 0; k < 9; k++) v += sm.A[i * 9 + k] * sm.V[k * 9 + j];
      sm.L[e] = v;
    }
    __syncwarp();
    double cnew[3];
    for (int e = lane, t = 0; e < 81; e += 32, t++) {  // C = W^T T
      const int i = e / 9, j = e % 9;
      double v = 0.0;
      for (int k = 0; k < 9; k++) v += sm.V[k * 9 + i] * sm.L[k * 9 + j];
      cnew[t] = v;
    }
    __syncwarp();
    for (int e = lane, t = 0; e < 81; e += 32, t++) sm.A[e] = cnew[t];
    __syncwarp();
    for (int e = lane; e < 81; e += 32) {  // symmetrise; excluded directions get a huge diagonal
      const int i = e / 9, j = e % 9;
      if (i < j) {
        const double v = 0.5 * (sm.A[i * 9 + j] + sm.A[j * 9 + i]);
        sm.A[i * 9 + j] = v;
        sm.A[j * 9 + i] = v;
      }
    }
    __syncwarp();
    if (lane < 9) {
      bool zero_col = true;
      for (int i = 0; i < 9; i++) zero_col = zero_col && sm.V[i * 9 + lane] == 0.0;
      if (zero_col) sm.A[lane * 9 + lane] = 1e300;
    }
    __syncwarp();
    warp_jacobi9(sm.A, sm.L, lane);  // eigenvectors Y in sm.L
    int mi = 0;
    for (int k = 1; k < 9; k++)
      if (sm.A[k * 9 + k] < sm.A[mi * 9 + mi]) mi = k;
    __syncwarp();
    if (lane < 9) {
      double v = 0.0;
      for (int k = 0; k < 9; k++) v += sm.V[lane * 9 + k] * sm.L[k * 9 + mi];
      sm.par[lane] = v;
    }
  }
  __syncwarp();
  if (lane == 0) {
    double mu = 0.0;
    for (int i = 0; i < 9; i++) mu += sm.m[i] * sm.par[i];
    sm.par[9] = -mu / n;
  }
  __syncwarp();
  double par[9];
#pragma unroll
  for (int i = 0; i < 9; i++) par[i] = sm.par[i];

  // --- walk 2: C = sum g g^T and T = sum g^(x6)
  double acc[34];
#pragma unroll
  for (int i = 0; i < 34; i++) acc[i] = 0.0;
  // the run list of this ball is built once and reused by both walks (queries touching more than
  // kRunCap rows are walked in batches)
  int nr = 0, row_off = 0;
  bool more = true;
  while (more) {
    nr = build_runs_warp(ri, row_ptr, col_ptr, pts_c, q.x, q.y, rpad, sm.runs.rs, sm.runs.pre, kRunCap, row_off, more);
    row_off += nr;
    walk_runs(pts_c, sm.runs, nr, q.x, q.y, q.z, r2, sm.ring, [&](const GPoint& p, bool active) {
      if (!active) return;
      const double x = (double(p.x) - qx) * inv_r, y = (double(p.y) - qy) * inv_r, z = (double(p.z) - qz) * inv_r;
      double gn[3], m6[28];
      quad_normal(par, x, y, z, gn);
      acc[0] += gn[0] * gn[0]; acc[1] += gn[1] * gn[1]; acc[2] += gn[2] * gn[2];
      acc[3] += gn[0] * gn[1]; acc[4] += gn[1] * gn[2]; acc[5] += gn[0] * gn[2];
      monomials6(gn, m6);
#pragma unroll
      for (int t = 0; t < 28; t++) acc[6 + t] += m6[t];
    });
  }
  const bool single_batch = row_off == nr;  // the common case: the list in shared memory is complete
  // warp reduction by recursive halving: lane i holds sum i (i < 32); sums 32, 33 by butterflies
  {
    const double r32 = warp_reduce_transpose32(acc);
    const double a32 = warp_sum(acc[32]), a33 = warp_sum(acc[33]);
#pragma unroll
    for (int i = 0; i < 6; i++) acc[i] = __shfl_sync(0xffffffffu, r32, i);  // sum g g^T, needed by every lane
    if (lane >= 6) sm.T[lane - 6] = r32 * c_multinomial6[lane - 6];          // T_0 .. T_25
    if (lane == 0) {
      sm.T[26] = a32 * c_multinomial6[26];
      sm.T[27] = a33 * c_multinomial6[27];
    }
  }
  double w3[3], V3[3][3];
  eig3(acc, w3, V3);
  int m3 = 0;
  if (w3[1] < w3[m3]) m3 = 1;
  if (w3[2] < w3[m3]) m3 = 2;  // quadric.cpp:278-280
  double ax[3] = {V3[0][m3], V3[1][m3], V3[2][m3]};
  __syncwarp();

  // --- walk 3: j* = argmax_j sum_i (g_i.g_j)^6 = argmax_j <T, g_j^(x6)>, first max in (dist, index) order
  double bestS = -1.0;
  float bestD = 3.0e38f;
  unsigned bestI = 0xFFFFFFFFu;
  double bestG[3] = {0, 0, 0};
  row_off = 0;
  more = true;
  while (more) {
    if (single_batch) more = false;
    else {
      nr = build_runs_warp(ri, row_ptr, col_ptr, pts_c, q.x, q.y, rpad, sm.runs.rs, sm.runs.pre, kRunCap, row_off, more);
      row_off += nr;
    }
    walk_runs(pts_c, sm.runs, nr, q.x, q.y, q.z, r2, sm.ring, [&](const GPoint& p, bool active) {
      if (!active) return;
      const double x = (double(p.x) - qx) * inv_r, y = (double(p.y) - qy) * inv_r, z = (double(p.z) - qz) * inv_r;
      double gn[3], m6[28];
      quad_normal(par, x, y, z, gn);
      monomials6(gn, m6);
      double S = 0.0;
#pragma unroll
      for (int t = 0; t < 28; t++) S += sm.T[t] * m6[t];
      const float d = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
      const unsigned id = p.tag >> 2;
      const bool better = S > bestS || (S == bestS && (d < bestD || (d == bestD && id < bestI)));
      if (better) {
        bestS = S; bestD = d; bestI = id;
        bestG[0] = gn[0]; bestG[1] = gn[1]; bestG[2] = gn[2];
      }
    });
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oS = __shfl_xor_sync(0xffffffffu, bestS, o);
    const float oD = __shfl_xor_sync(0xffffffffu, bestD, o);
    const unsigned oI = __shfl_xor_sync(0xffffffffu, bestI, o);
    const double g0 = __shfl_xor_sync(0xffffffffu, bestG[0], o);
    const double g1 = __shfl_xor_sync(0xffffffffu, bestG[1], o);
    const double g2 = __shfl_xor_sync(0xffffffffu, bestG[2], o);
    const bool better = oS > bestS || (oS == bestS && (oD < bestD || (oD == bestD && oI < bestI)));
    if (better) {
      bestS = oS; bestD = oD; bestI = oI;
      bestG[0] = g0; bestG[1] = g1; bestG[2] = g2;
    }
  }
  // --- frame (quadric.cpp:285-304)
  double np[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const double d0 = (r == 0 ? 1.0 : 0.0) - ax[r] * ax[0];
    const double d1 = (r == 1 ? 1.0 : 0.0) - ax[r] * ax[1];
    const double d2 = (r == 2 ? 1.0 : 0.0) - ax[r] * ax[2];
    np[r] = (d0 * bestG[0] + d1 * bestG[1]) + d2 * bestG[2];
  }
  const double nrm = sqrt(np[0] * np[0] + (np[1] * np[1] + np[2] * np[2]));
  double nor[3] = {np[0] / nrm, np[1] / nrm, np[2] / nrm};
  double bin[3] = {ax[1] * nor[2] - ax[2] * nor[1], ax[2] * nor[0] - ax[0] * nor[2], ax[0] * nor[1] - ax[1] * nor[0]};
  const double cx = major ? cam1x : cam0x, cy = major ? cam1y : cam0y, cz = major ? cam1z : cam0z;
  const double t0 = qx - cx, t1 = qy - cy, t2 = qz - cz;
  if (nor[0] * t0 + (nor[1] * t1 + nor[2] * t2) > 0) {
    nor[0] = -nor[0]; nor[1] = -nor[1]; nor[2] = -nor[2];
  }
  if (bin[0] * t0 + (bin[1] * t1 + bin[2] * t2) > 0) {
    bin[0] = -bin[0]; bin[1] = -bin[1]; bin[2] = -bin[2];
  }
  ax[0] = nor[1] * bin[2] - nor[2] * bin[1];
  ax[1] = nor[2] * bin[0] - nor[0] * bin[2];
  ax[2] = nor[0] * bin[1] - nor[1] * bin[0];
  if (lane == 0) {
    const bool good = ok && n >= 1.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      F.normal[k] = good ? nor[k] : 0.0;
      F.axis[k] = good ? ax[k] : 0.0;
      F.binormal[k] = good ? bin[k] : 0.0;
    }
    frames[s] = F;
    if (normals_out && n >= 1.0) {  // hand_search.cpp:102: cloud_normals_.col(idx) = normal
      normals_out[size_t(3) * idx + 0] = F.normal[0];
      normals_out[size_t(3) * idx + 1] = F.normal[1];
      normals_out[size_t(3) * idx + 2] = F.normal[2];
    }
  }
}

// one thread per sample after the fit: totals of the per-sample neighbour / candidate counts (one atomic
// pair per CTA, reported through ag_timings) and, when the normals were written to cloud_normals_, the
// tag bit that lets the sweep fetch only normals that exist
__global__ void k_quadric_finish(GPoint* pts, const RowIndex* __restrict__ rip, const int* __restrict__ indices, int n,
                                 const int* __restrict__ d_count, const int2* __restrict__ nn_counts,
                                 unsigned long long* __restrict__ counters, int mark) {
  __shared__ unsigned long long s_sum[2];
  if (threadIdx.x < 2) s_sum[threadIdx.x] = 0ull;
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int nn = 0, nc = 0;
  if (i < n && i < *d_count) {
    const int idx = indices[i];
    if (idx >= 0 && idx < rip->n_points) {
      if (mark) atomicOr(&pts[idx].tag, kTagNormalBit);
      const int2 v = nn_counts[i];
      nn = v.x;
      nc = v.y;
    }
  }
  nn = __reduce_add_sync(0xffffffffu, nn);
  nc = __reduce_add_sync(0xffffffffu, nc);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_sum[0], (unsigned long long)nn);
    atomicAdd(&s_sum[1], (unsigned long long)nc);
  }
  __syncthreads();
  if (threadIdx.x < 2 && s_sum[threadIdx.x]) atomicAdd(&counters[threadIdx.x], s_sum[threadIdx.x]);
}

}  // namespace

int fit_quadrics_device(Ctx* c, const int* d_indices, int n, const int* d_count, double radius, ag_frame* d_frames,
                        bool write_normals) {
  if (n <= 0) return AG_OK;
  if (c->moments.reserve(size_t(n) * kMomentStride * sizeof(double)) || c->nn_counts.reserve(size_t(n) * 8) ||
      c->counters.reserve(64))
    return AG_ERR_CUDA;
  const float r2 = float(radius * radius);  // PCL hands radius*radius to FLANN as float
  const double rpad = sqrt(double(r2)) * (1.0 + 1e-5) + 1e-7;
  // coordinates are centred on the sample and scaled by the power of two nearest to 1/r (the fit is
  // invariant under translation and uniform scale; a power of two makes the scaling exact)
  const double inv_r = ldexp(1.0, int(lrint(log2(1.0 / radius))));
  const int blocks = (n + kWarps - 1) / kWarps;
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  const RowIndex* ri = c->row_index.as<RowIndex>();
  cudaEventRecord(c->ev_k[0], c->stream);
  k_taubin_moments<<<blocks, kWarps * 32, 0, c->stream>>>(c->vox.as<GPoint>(), c->row_ptr.as<int>(),
                                                          c->col_ptr.as<int>(), ri, d_indices, n, d_count, r2, rpad,
                                                          inv_r, c->moments.as<double>(), c->nn_counts.as<int2>());
  cudaEventRecord(c->ev_k[1], c->stream);
  const size_t smem = sizeof(AxesSmem) * kWarps;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_taubin_axes, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    cudaFuncSetAttribute(k_taubin_axes, cudaFuncAttributePreferredSharedMemoryCarveout, 50);
    attr_set = true;
  }
  const HandConst& h = c->hand;
  k_taubin_axes<<<blocks, kWarps * 32, smem, c->stream>>>(
      c->vox.as<GPoint>(), c->row_ptr.as<int>(), c->col_ptr.as<int>(), ri, d_indices, n, d_count, r2, rpad, inv_r,
      c->moments.as<double>(),
      h.cam[0][0], h.cam[0][1], h.cam[0][2], h.cam[1][0], h.cam[1][1], h.cam[1][2], d_frames,
      write_normals ? c->normals.as<double>() : nullptr);
  cudaEventRecord(c->ev_k[2], c->stream);
  c->launches += 3;
  k_quadric_finish<<<(n + 255) / 256, 256, 0, c->stream>>>(c->vox.as<GPoint>(), ri, d_indices, n, d_count,
                                                          c->nn_counts.as<int2>(), ctr, write_normals ? 1 : 0);
  AG_CUDA_CHECK(cudaGetLastError());
  return AG_OK;
}

}  // namespace ag
