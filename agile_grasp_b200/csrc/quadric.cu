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
constexpr int kNumMoments = 35;
constexpr int kMomentStride = 36;  // + count of camera-1 neighbours

// exponents of the quadric basis [x2 y2 z2 xy yz xz x y z] (quadric.cpp:40-73)
constexpr int kBasis[9][3] = {{2, 0, 0}, {0, 2, 0}, {0, 0, 2}, {1, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
// where entry (i, j) of the reduced pencil comes from: A_ij = mom[a_ij] - mom[a_i] mom[a_j] / n and
// B_ij = sum_t b_coef[t] mom[b_idx[t]] (N = sum grad phi^T grad phi, quadric.cpp:103-131), as moment indices —
// a per-entry table read with one coalesced load instead of chains of lane-divergent constant-memory lookups
struct PencilEntry {
  unsigned char a_ij, a_i, a_j, nb, b_idx[3], b_coef[3], pad[2];
};
struct PencilTable {
  PencilEntry e[81];
  constexpr PencilTable() : e() {
    for (int i = 0; i < 9; i++)
      for (int j = 0; j < 9; j++) {
        PencilEntry& t = e[i * 9 + j];
        const int* bi = kBasis[i];
        const int* bj = kBasis[j];
        t.a_ij = static_cast<unsigned char>(midx(bi[0] + bj[0], bi[1] + bj[1], bi[2] + bj[2]));
        t.a_i = static_cast<unsigned char>(midx(bi[0], bi[1], bi[2]));
        t.a_j = static_cast<unsigned char>(midx(bj[0], bj[1], bj[2]));
        t.nb = 0;
        for (int a = 0; a < 3; a++)
          if (bi[a] >= 1 && bj[a] >= 1) {
            int ex[3] = {bi[0] + bj[0], bi[1] + bj[1], bi[2] + bj[2]};
            ex[a] -= 2;
            t.b_idx[t.nb] = static_cast<unsigned char>(midx(ex[0], ex[1], ex[2]));
            t.b_coef[t.nb] = static_cast<unsigned char>(bi[a] * bj[a]);
            t.nb++;
          }
        for (int k = t.nb; k < 3; k++) {
          t.b_idx[k] = 0;
          t.b_coef[k] = 0;
        }
        t.pad[0] = t.pad[1] = 0;
      }
  }
};
__device__ const PencilTable g_pencil = PencilTable();

// ---- PTX helpers: shared-space accesses, mbarrier, TMA bulk copy ---------------------------------
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
// explicit shared-space accesses (one 32-bit address register, immediate offsets)
__device__ __forceinline__ GPoint lds_point(uint32_t addr) {
  GPoint p;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=f"(p.x), "=f"(p.y), "=f"(p.z), "=r"(p.tag)
               : "r"(addr)
               : "memory");
  return p;
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
// the result is opaque to the compiler, so the address stays in a register instead of being rebuilt
__device__ __forceinline__ uint32_t pin_u32(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

// ---- kernel 0: radius search -> neighbour lists -------------------------------------------------
// Replaces pcl::KdTreeFLANN::radiusSearch (hand_search.cpp:85).  One warp per sample, four independent
// warps per CTA.  Per sample:
//   1. lane r finds the candidate run of x-row r of the ball (two independent loads from the column table);
//      a run is contiguous in the voxel list, so
//   2. each row lane issues ONE bulk async copy (TMA, cp.async.bulk) of its run into the warp's shared
//      staging buffer at the run's prefix offset; one mbarrier transaction count tracks all of them — every
//      load of the ball is in flight at once and no registers are held for them;
//   3. the staged candidates are tested 64 per step with FLANN's exact binary32 expression and the
//      accepted records are written, compacted (ballot + popc), to this sample's slot of the neighbour
//      pool: `stride` 16-byte records per sample, consumed as one contiguous stream by the moments kernel
//      and by the two normal walks of the axes kernel.
// List order = (camera, x-row, y, z) = ascending point index; no consumer depends on FLANN's
// (distance, index) order except through explicit tie-breaks.
constexpr int kStageCap = 384;  // candidates staged per pass and warp (6 KB); bigger balls take more passes

template <bool WRITE>  // WRITE = false: neighbour counts only (no lists)
__global__ void __launch_bounds__(kWarps * 32, 8)
k_ball_search(const GPoint* __restrict__ pts, const int* __restrict__ row_ptr, const int* __restrict__ col_ptr,
              RowIndex* __restrict__ rip, const int* __restrict__ indices, int s0, int n_samples_max,
              const int* __restrict__ d_count, float r2, double rpad, GPoint* __restrict__ pool, int stride,
              int2* __restrict__ nn_counts, float4* __restrict__ heads) {
  __shared__ __align__(16) GPoint s_cand[kWarps][kStageCap];
  __shared__ __align__(8) unsigned long long s_bar[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sl = blockIdx.x * kWarps + warp;  // slot of this launch
  const int s = s0 + sl;
  const RowIndex& ri = *rip;  // read in place (L1 hits), not held in registers
  if (s >= n_samples_max) return;
  const int idx = s < *d_count ? indices[s] : -1;
  if (idx < 0 || idx >= ri.n_points) {  // not a sample: an empty list
    if (lane == 0) {
      nn_counts[s] = make_int2(0, 0);
      if (WRITE) heads[sl] = make_float4(0.f, 0.f, 0.f, __int_as_float(0));
    }
    return;
  }
  const GPoint q = pts[idx];
  const uint32_t cand_s = pin_u32(smem_u32(s_cand[warp])), bar = pin_u32(smem_u32(&s_bar[warp]));
  const uint32_t lane_s = cand_s + uint32_t(lane) * 16u;
  if (lane == 0) mbar_init(bar, 1);
  fence_proxy_async_smem();  // the initialised barrier must be visible to the async proxy
  __syncwarp();
  uint32_t parity = 0;
  GPoint* out = pool + size_t(sl) * size_t(stride);
  int n_cand = 0, n_out = 0;
  const unsigned lt = (1u << lane) - 1u;
  GPoint far_pt;  // fails the radius test
  far_pt.x = 1e30f;
  far_pt.y = far_pt.z = 0.f;
  far_pt.tag = 0;
  // one loop over (camera, batch of 32 x-rows): one batch per camera for the shipped Taubin radii
  int k_lo0 = 0, k_hi0 = -1, k_lo1 = 0, k_hi1 = -1;
  if (ri.count[0] > 0) row_range(ri, 0, q.x, rpad, k_lo0, k_hi0);
  if (ri.count[1] > 0) row_range(ri, 1, q.x, rpad, k_lo1, k_hi1);
  const int nb0 = k_hi0 >= k_lo0 ? (k_hi0 - k_lo0) / 32 + 1 : 0;
  const int nb1 = k_hi1 >= k_lo1 ? (k_hi1 - k_lo1) / 32 + 1 : 0;
  for (int t = 0; t < nb0 + nb1; t++) {
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
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int total_rem = __shfl_sync(0xffffffffu, incl, 31);
      if (total_rem == 0) break;
      const int excl = incl - rem;
      const int take = min(rem, max(0, kStageCap - excl));
      const int total = min(total_rem, kStageCap);
      fence_proxy_async_smem();  // earlier generic-proxy reads of the buffer precede the async writes
      __syncwarp();
      if (lane == 0) mbar_expect_tx(bar, uint32_t(total) * 16u);
      __syncwarp();
      if (take > 0) bulk_g2s(cand_s + uint32_t(excl) * 16u, pts + j_lo, uint32_t(take) * 16u, bar);
      j_lo += take;
      rem -= take;
      mbar_wait(bar, parity);
      parity ^= 1u;
      // membership test, 64 candidates per step (two independent chains), compacted store to the pool
      for (int base = 0; base < total; base += 64) {
        GPoint p0 = far_pt, p1 = far_pt;
        if (base + lane < total) p0 = lds_point(lane_s + uint32_t(base) * 16u);
        if (base + lane + 32 < total) p1 = lds_point(lane_s + uint32_t(base) * 16u + 512u);
        const bool ok0 = dist2_flann(q.x, q.y, q.z, p0.x, p0.y, p0.z) < r2;
        const bool ok1 = dist2_flann(q.x, q.y, q.z, p1.x, p1.y, p1.z) < r2;
        const unsigned m0 = __ballot_sync(0xffffffffu, ok0), m1 = __ballot_sync(0xffffffffu, ok1);
        const int w0 = n_out + __popc(m0 & lt), w1 = n_out + __popc(m0) + __popc(m1 & lt);
        if (WRITE && ok0 && w0 < stride) out[w0] = p0;
        if (WRITE && ok1 && w1 < stride) out[w1] = p1;
        n_out += __popc(m0) + __popc(m1);
      }
      __syncwarp();
    }
  }
  n_cand = __reduce_add_sync(0xffffffffu, n_cand);
  if (lane == 0) {
    nn_counts[s] = make_int2(min(n_out, stride), n_cand);
    if (WRITE) heads[sl] = make_float4(q.x, q.y, q.z, __int_as_float(min(n_out, stride)));  // what the moments kernel needs
    if (WRITE && n_out > stride) atomicOr(&rip->error, kErrBallOverflow);
  }
}

// ---- "more than 50 neighbours?" for ALL samples of a sharded call -------------------------------------
// In the production normal mode a sample consumes 50 rand() draws iff its ball holds more than 50 points
// (quadric.cpp:177-192), so a shard needs that one bit of every sample of the call to find its slice of the stream.
// One warp per sample, one lane per x-row of the ball; every lane walks its candidate run from the middle outwards
// (the points nearest in y first), two candidates per step straight from L2, and the warp stops as soon as the count
// passes 50 — a few steps for a typical ball of ~260 points.  nn_counts[s].x = min(count, 51).
__global__ void __launch_bounds__(kWarps * 32, 8)
k_ball_over50(const GPoint* __restrict__ pts, const int* __restrict__ row_ptr, const int* __restrict__ col_ptr,
              const RowIndex* __restrict__ rip, const int* __restrict__ indices, int n_samples_max,
              const int* __restrict__ d_count, float r2, double rpad, int2* __restrict__ nn_counts) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = blockIdx.x * kWarps + warp;
  const RowIndex& ri = *rip;
  if (s >= n_samples_max) return;
  const int idx = s < *d_count ? indices[s] : -1;
  if (idx < 0 || idx >= ri.n_points) {
    if (lane == 0) nn_counts[s] = make_int2(0, 0);
    return;
  }
  const GPoint q = pts[idx];
  int n_out = 0;
  int k_lo0 = 0, k_hi0 = -1, k_lo1 = 0, k_hi1 = -1;
  if (ri.count[0] > 0) row_range(ri, 0, q.x, rpad, k_lo0, k_hi0);
  if (ri.count[1] > 0) row_range(ri, 1, q.x, rpad, k_lo1, k_hi1);
  const int nb0 = k_hi0 >= k_lo0 ? (k_hi0 - k_lo0) / 32 + 1 : 0;
  const int nb1 = k_hi1 >= k_lo1 ? (k_hi1 - k_lo1) / 32 + 1 : 0;
  for (int t = 0; t < nb0 + nb1 && n_out <= 50; t++) {
    const int c = t < nb0 ? 0 : 1;
    const int kb = c ? k_lo1 + (t - nb0) * 32 : k_lo0 + t * 32;
    const int k_hi = c ? k_hi1 : k_hi0;
    const int nrows = min(32, k_hi - kb + 1);
    int j_lo = 0, j_end = 0;
    if (lane < nrows) row_run(ri, row_ptr, col_ptr, pts, c, kb + lane, q.x, q.y, rpad, j_lo, j_end);
    const int len = max(0, j_end - j_lo), mid = j_lo + (len >> 1);
    const int steps = __reduce_max_sync(0xffffffffu, (len + 1) >> 1);
    for (int i = 0; i < steps; i++) {
      const int a = mid + i, b = mid - 1 - i;
      bool oka = false, okb = false;
      if (a < j_end) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(pts + a));
        oka = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2;
      }
      if (b >= j_lo) {
        const float4 p = __ldg(reinterpret_cast<const float4*>(pts + b));
        okb = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < r2;
      }
      n_out += __popc(__ballot_sync(0xffffffffu, oka)) + __popc(__ballot_sync(0xffffffffu, okb));
      if (n_out > 50) break;
    }
  }
  if (lane == 0) nn_counts[s] = make_int2(min(n_out, 51), 0);
}

// ---- kernel 0+1 fused: radius search AND Taubin moments ------------------------------------------
// The accepted points of the search are already in registers, so the 34 monomial moments are accumulated right
// there instead of being streamed back from the neighbour lists by a second kernel (which made the lists' 16 B
// per neighbour travel through HBM twice and cost a launch).  Same staging as k_ball_search — one TMA bulk copy
// per x-row run, one mbarrier transaction count — then per step of 64 staged candidates: FLANN's binary32
// membership test, ballot-compacted store of the accepted records to the neighbour list (the normal walks and
// the pick ranking read it later), and 34 predicated binary64 fused multiply-adds per lane on coordinates centred
// on the sample (a rejected candidate contributes exact zeros).  The 32 x 34 partial sums are transposed through
// the staging buffer (two rounds of 16 moments, row pitch 33) and scaled by the power-of-two coordinate scale.
constexpr int kFusedWarps = 4;
// non-deterministic normal mode: sample s consumes rand() draws [50 k_s, 50 k_s + 50), k_s = number of earlier
// samples (in sample order, continuing across the launches of one call) with more than 50 neighbours
// (one CTA, any multiple of 32 threads up to 1024)
__device__ __forceinline__ void rand_offsets_block(const int2* __restrict__ nn_counts, int s0, int m, int cnt_valid,
                                                   int* __restrict__ rand_off, int* __restrict__ carry, int* s_warp) {
  // every warp owns a contiguous range of samples and walks it 32 at a time (coalesced, independent loads): a ballot
  // of "more than 50 neighbours" per step; pass 1 counts, one block scan over the warps, pass 2 writes the offsets
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int per = ((m + nw - 1) / nw + 31) & ~31, lo = min(m, warp * per), hi = min(m, lo + per);
  auto flag = [&](int i) { return i < hi && s0 + i < cnt_valid && __ldcg(&nn_counts[s0 + i].x) > 50; };
  // the ballots of pass 1 are kept (shared memory, up to kKeep steps per warp) so that pass 2 issues no loads
  constexpr int kKeep = 64;
  __shared__ unsigned s_mask[32][kKeep];
  int tot = 0;
#pragma unroll 8
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const unsigned mask = __ballot_sync(0xffffffffu, flag(i0 + lane));
    const int step = (i0 - lo) >> 5;
    if (lane == 0 && step < kKeep) s_mask[warp][step] = mask;
    tot += __popc(mask);
  }
  if (lane == 0) s_warp[warp] = tot;
  __syncthreads();
  if (warp == 0) {
    int w = lane < nw ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;  // inclusive; entry 31 = the total
  }
  __syncthreads();
  const int base = *carry;
  int run = base + (warp > 0 ? s_warp[warp - 1] : 0);
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int step = (i0 - lo) >> 5;
    const unsigned mask = step < kKeep ? s_mask[warp][step] : __ballot_sync(0xffffffffu, flag(i0 + lane));
    if (i0 + lane < hi) rand_off[s0 + i0 + lane] = run + __popc(mask & lt);
    run += __popc(mask);
  }
  __syncthreads();
  if (tid == 0) *carry = base + s_warp[31];
}
__device__ __forceinline__ void ball_moments_warp(const GPoint* __restrict__ pts, const int* __restrict__ row_ptr,
                                                  const int* __restrict__ col_ptr, RowIndex* __restrict__ rip,
                                                  const int* __restrict__ indices, int s0, int n_samples_max,
                                                  const int* __restrict__ d_count, float r2, double rpad,
                                                  GPoint* __restrict__ pool, int stride, int2* __restrict__ nn_counts,
                                                  double scale, double* __restrict__ moments) {
  __shared__ __align__(16) GPoint s_cand[kFusedWarps][kStageCap];
  __shared__ __align__(8) unsigned long long s_bar[kFusedWarps];
  static_assert(kStageCap * sizeof(GPoint) >= 16 * 33 * sizeof(double), "the reduction reuses the staging buffer");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sl = blockIdx.x * kFusedWarps + warp;  // slot of this launch
  const int s = s0 + sl;
  const RowIndex& ri = *rip;
  if (s >= n_samples_max) return;
  const int idx = s < *d_count ? indices[s] : -1;
  if (idx < 0 || idx >= ri.n_points) {  // not a sample: an empty list
    if (lane == 0) nn_counts[s] = make_int2(0, 0);
    return;
  }
  const GPoint q = pts[idx];
  const double qx = double(q.x), qy = double(q.y), qz = double(q.z);
  const uint32_t cand_s = pin_u32(smem_u32(s_cand[warp])), bar = pin_u32(smem_u32(&s_bar[warp]));
  const uint32_t lane_s = cand_s + uint32_t(lane) * 16u;
  if (lane == 0) mbar_init(bar, 1);
  fence_proxy_async_smem();
  __syncwarp();
  uint32_t parity = 0;
  GPoint* out = pool + size_t(sl) * size_t(stride);
  int n_cand = 0, n_out = 0, cam1 = 0;
  const unsigned lt = (1u << lane) - 1u;
  double acc[kNumMoments];  // acc[0] unused: the count is known
#pragma unroll
  for (int i = 0; i < kNumMoments; i++) acc[i] = 0.0;
  auto accumulate = [&](const GPoint& p, bool ok) {
    const double x = ok ? double(p.x) - qx : 0.0, y = ok ? double(p.y) - qy : 0.0, z = ok ? double(p.z) - qz : 0.0;
    const double x2 = x * x, y2 = y * y, z2 = z * z, xy = x * y, yz = y * z, xz = x * z;
    acc[1] += x; acc[2] += y; acc[3] += z;
    acc[4] += x2; acc[5] += y2; acc[6] += z2; acc[7] += xy; acc[8] += yz; acc[9] += xz;
    acc[10] += x2 * x; acc[11] += y2 * y; acc[12] += z2 * z; acc[13] += x2 * y; acc[14] += x2 * z;
    acc[15] += x * y2; acc[16] += y2 * z; acc[17] += x * z2; acc[18] += y * z2; acc[19] += xy * z;
    acc[20] += x2 * x2; acc[21] += y2 * y2; acc[22] += z2 * z2; acc[23] += x2 * xy; acc[24] += x2 * xz;
    acc[25] += xy * y2; acc[26] += y2 * yz; acc[27] += xz * z2; acc[28] += yz * z2; acc[29] += x2 * y2;
    acc[30] += y2 * z2; acc[31] += x2 * z2; acc[32] += x2 * yz; acc[33] += xy * yz; acc[34] += xz * yz;
  };
  int k_lo0 = 0, k_hi0 = -1, k_lo1 = 0, k_hi1 = -1;
  if (ri.count[0] > 0) row_range(ri, 0, q.x, rpad, k_lo0, k_hi0);
  if (ri.count[1] > 0) row_range(ri, 1, q.x, rpad, k_lo1, k_hi1);
  const int nb0 = k_hi0 >= k_lo0 ? (k_hi0 - k_lo0) / 32 + 1 : 0;
  const int nb1 = k_hi1 >= k_lo1 ? (k_hi1 - k_lo1) / 32 + 1 : 0;
  for (int t = 0; t < nb0 + nb1; t++) {
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
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int total_rem = __shfl_sync(0xffffffffu, incl, 31);
      if (total_rem == 0) break;
      const int excl = incl - rem;
      const int take = min(rem, max(0, kStageCap - excl));
      const int total = min(total_rem, kStageCap);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_expect_tx(bar, uint32_t(total) * 16u);
      __syncwarp();
      if (take > 0) bulk_g2s(cand_s + uint32_t(excl) * 16u, pts + j_lo, uint32_t(take) * 16u, bar);
      j_lo += take;
      rem -= take;
      mbar_wait(bar, parity);
      parity ^= 1u;
      for (int base = 0; base < total; base += 64) {
        GPoint p0 = q, p1 = q;  // padding = the sample itself: distance 0 but masked out below
        const bool in0 = base + lane < total, in1 = base + lane + 32 < total;
        if (in0) p0 = lds_point(lane_s + uint32_t(base) * 16u);
        if (in1) p1 = lds_point(lane_s + uint32_t(base) * 16u + 512u);
        const bool ok0 = in0 && dist2_flann(q.x, q.y, q.z, p0.x, p0.y, p0.z) < r2;
        const bool ok1 = in1 && dist2_flann(q.x, q.y, q.z, p1.x, p1.y, p1.z) < r2;
        const unsigned m0 = __ballot_sync(0xffffffffu, ok0), m1 = __ballot_sync(0xffffffffu, ok1);
        const int w0 = n_out + __popc(m0 & lt), w1 = n_out + __popc(m0) + __popc(m1 & lt);
        if (ok0 && w0 < stride) out[w0] = p0;
        if (ok1 && w1 < stride) out[w1] = p1;
        n_out += __popc(m0) + __popc(m1);
        accumulate(p0, ok0);
        accumulate(p1, ok1);
        cam1 += (ok0 ? int(p0.tag & kTagCamBit) : 0) + (ok1 ? int(p1.tag & kTagCamBit) : 0);
      }
      __syncwarp();
    }
  }
  n_cand = __reduce_add_sync(0xffffffffu, n_cand);
  cam1 = __reduce_add_sync(0xffffffffu, cam1);
  // warp reduction of the 34 moments through the (now idle) staging buffer: two rounds of 16 with a 33-double row
  // pitch; lane i sums half (i >> 4) of row (i & 15), the halves meet by one exchange
  const double s2 = scale * scale, s4 = s2 * s2;
  const int mj = lane + 1;  // the moment this lane writes
  const double scl = mj <= 3 ? scale : mj <= 9 ? s2 : mj <= 19 ? s2 * scale : s4;
  const uint32_t row_s = cand_s + uint32_t(lane & 15) * (33u * 8u) + uint32_t(lane >> 4) * (16u * 8u);
  double r01[2];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; j++) sts_f64(cand_s + uint32_t(j * 33) * 8u + uint32_t(lane) * 8u, acc[1 + 16 * r + j]);
    __syncwarp();
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      t0 += lds_f64(row_s + uint32_t(k) * 8u);
      t1 += lds_f64(row_s + uint32_t(k + 1) * 8u);
    }
    const double tt = t0 + t1;
    r01[r] = tt + __shfl_xor_sync(0xffffffffu, tt, 16);
  }
  const double mom = lane < 16 ? r01[0] : r01[1];
  const double m33 = warp_sum(acc[33]), m34 = warp_sum(acc[34]);
  const int n_list = min(n_out, stride);
  if (n_list > 0) {
    double* mo = moments + size_t(s) * kMomentStride;
    mo[mj] = mom * scl;  // moments 1..32
    if (lane == 0) {
      mo[0] = double(n_list);
      mo[33] = m33 * s4;
      mo[34] = m34 * s4;
      mo[35] = double(cam1);
    }
  }
  if (lane == 0) {
    nn_counts[s] = make_int2(n_list, n_cand);
    if (n_out > stride) atomicOr(&rip->error, kErrBallOverflow);
  }
}

// The kernel: one warp per sample.  With rand_off != null the CTA that finishes last also lays out the rand() stream
// (rand_offsets_block); measured on B200 that epilogue lengthens this kernel — which is on the critical path — by
// 7 us, while the separate k_rand_offsets launch runs on the side stream next to the eigen-solve, so ag_localize
// passes null.
__global__ void __launch_bounds__(kFusedWarps * 32, 4)
k_ball_moments(const GPoint* __restrict__ pts, const int* __restrict__ row_ptr, const int* __restrict__ col_ptr,
               RowIndex* __restrict__ rip, const int* __restrict__ indices, int s0, int n_samples_max,
               const int* __restrict__ d_count, float r2, double rpad, GPoint* __restrict__ pool, int stride,
               int2* __restrict__ nn_counts, double scale, double* __restrict__ moments, int* __restrict__ rand_off,
               int* __restrict__ rand_carry) {
  __shared__ int s_scan[32];
  __shared__ int s_last;
  ball_moments_warp(pts, row_ptr, col_ptr, rip, indices, s0, n_samples_max, d_count, r2, rpad, pool, stride, nn_counts,
                    scale, moments);
  if (!rand_off) return;
  __threadfence();  // this CTA's neighbour counts are visible before it signs off
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&rand_carry[1], 1) == int(gridDim.x) - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  rand_offsets_block(nn_counts, s0, n_samples_max - s0, *d_count, rand_off, rand_carry, s_scan);
  if (threadIdx.x == 0) rand_carry[1] = 0;  // ready for the next launch
}

// ---- kernel 1: moments ------------------------------------------------------------------------
// Roofline-graded kernel: streams each sample's neighbour list (16 B per neighbour, contiguous) and
// accumulates the 34 monomial moments of degree 1..4 in binary64.  Persistent warps (one CTA slot per SM
// and register budget), each walking samples gw, gw + G, ...; the stream is cut into segments of kSeg
// records that ping-pong between two shared buffers per warp:
//   1. one elected lane issues the bulk async copy (TMA) of the NEXT segment — the rest of this list or,
//      speculatively at full size, the head of the next sample's list — before the warp touches the
//      current one, so a copy is always in flight behind the arithmetic and no registers are held for it;
//   2. 32 points per step: coordinates centred on the sample (exact: differences of binary32 values),
//      6 products, 34 fused multiply-adds per lane;
//   3. the 32 x 34 partial sums are transposed through the just-consumed buffer in two rounds of 16
//      moments (both half-warps sum half a row each, immediate offsets, conflict free), the power-of-two
//      coordinate scale is applied exactly (moment of degree d times scale^d), 36 doubles out.
constexpr int kSeg = 272;  // records per segment (4352 B >= the 16 x 33 doubles of a reduction round)
static_assert(kSeg * sizeof(GPoint) >= 16 * 33 * sizeof(double), "the reduction reuses a segment buffer");

// WPS = warps per sample: 1 when the launch fills the machine; 2 / 4 for small launches, where a sample's
// list is shared by a team of warps (the team leader issues the copies, warp `sub` takes every WPS-th batch of
// 32 records, the partial sums meet in shared memory in a fixed order) so that a 2000-sample launch is not
// bounded by the latency of one warp walking a whole list.
template <int WPS>
__global__ void __launch_bounds__(kWarps * 32, 4)
k_taubin_moments(int s0, int m, const float4* __restrict__ heads, const GPoint* __restrict__ pool, int stride,
                 double scale, double* __restrict__ moments) {
  __shared__ __align__(16) GPoint s_seg[kWarps][2][kSeg];
  __shared__ __align__(8) unsigned long long s_bar[kWarps][2];
  __shared__ double s_part[WPS > 1 ? kWarps : 1][kMomentStride];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kTeams = kWarps / WPS;           // teams per CTA
  const int team = warp / WPS, sub = warp % WPS;  // sub == 0: the team leader
  const int lead = team * WPS;                    // leader's warp index: its buffers stage the team's lists
  const int G = gridDim.x * kTeams;
  int sl = blockIdx.x * kTeams + team;
  if (sl >= m) return;  // (the whole team leaves together)
  auto team_sync = [&]() {
    if (WPS == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(WPS * 32) : "memory");
  };
  const uint32_t buf0 = pin_u32(smem_u32(s_seg[lead][0])), bar0 = pin_u32(smem_u32(&s_bar[lead][0]));
  const uint32_t own0 = pin_u32(smem_u32(s_seg[warp][0]));  // non-leaders: private reduction scratch
  constexpr uint32_t kBufBytes = kSeg * sizeof(GPoint);
  const uint32_t seg_bytes = uint32_t(min(kSeg, stride)) * 16u;  // speculative first segment of a list
  if (sub == 0 && lane == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8u, 1);
    fence_proxy_async_smem();
    mbar_expect_tx(bar0, seg_bytes);
    bulk_g2s(buf0, pool + size_t(sl) * size_t(stride), seg_bytes, bar0);
  }
  team_sync();  // the barriers are initialised before anyone waits on them
  float4 head = heads[sl];
  const double s2 = scale * scale, s4 = s2 * s2;
  const int mj = lane + 1;  // the moment this lane writes
  const double scl = mj <= 3 ? scale : mj <= 9 ? s2 : mj <= 19 ? s2 * scale : s4;
  int g = 0;  // segments consumed so far: buffer g & 1, barrier phase (g >> 1) & 1
  while (sl < m) {
    const int n = __float_as_int(head.w);
    const int sl_next = sl + G;
    float4 head_next = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sl_next < m) head_next = heads[sl_next];
    const double qx = double(head.x), qy = double(head.y), qz = double(head.z);
    GPoint self;  // padding of the last batch: every term of the sample itself is exactly zero
    self.x = head.x;
    self.y = head.y;
    self.z = head.z;
    self.tag = 0;
    const GPoint* src = pool + size_t(sl) * size_t(stride);
    double acc[kNumMoments];  // acc[0] is unused: the count is known
#pragma unroll
    for (int i = 0; i < kNumMoments; i++) acc[i] = 0.0;
    int cam1 = 0;
    const int nseg = max(1, (n + kSeg - 1) / kSeg);
    for (int k = 0; k < nseg; k++, g++) {
      const uint32_t cur = buf0 + uint32_t(g & 1) * kBufBytes, cur_bar = bar0 + uint32_t(g & 1) * 8u;
      const uint32_t oth = buf0 + uint32_t((g + 1) & 1) * kBufBytes, oth_bar = bar0 + uint32_t((g + 1) & 1) * 8u;
      team_sync();  // everyone is done with the other buffer (segment g - 1 or the leader's reduction scratch)
      if (sub == 0 && lane == 0) {
        fence_proxy_async_smem();
        if (k + 1 < nseg) {
          const uint32_t bytes = uint32_t(min(kSeg, n - (k + 1) * kSeg)) * 16u;
          mbar_expect_tx(oth_bar, bytes);
          bulk_g2s(oth, src + (k + 1) * kSeg, bytes, oth_bar);
        } else if (sl_next < m) {
          mbar_expect_tx(oth_bar, seg_bytes);
          bulk_g2s(oth, pool + size_t(sl_next) * size_t(stride), seg_bytes, oth_bar);
        }
      }
      mbar_wait(cur_bar, uint32_t(g >> 1) & 1u);
      const int cnt = min(kSeg, n - k * kSeg);
      const uint32_t lane_s = cur + uint32_t(lane) * 16u;
      for (int b = sub * 32; b < cnt; b += WPS * 32) {
        GPoint p = self;
        if (b + lane < cnt) p = lds_point(lane_s + uint32_t(b) * 16u);
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
      }
    }
    // warp reduction through shared memory: two rounds of 16 moments with a 33-double row pitch; lane i sums
    // half (i >> 4) of row (i & 15), the halves meet by one exchange.  Scratch = the buffer consumed last
    // (leader, once the team is done reading it) or the warp's own, otherwise unused, buffer.
    if (WPS > 1) team_sync();
    const uint32_t scr = sub == 0 ? buf0 + uint32_t((g - 1) & 1) * kBufBytes : own0;
    const uint32_t row_s = scr + uint32_t(lane & 15) * (33u * 8u) + uint32_t(lane >> 4) * (16u * 8u);
    double r01[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 16; j++) sts_f64(scr + uint32_t(j * 33) * 8u + uint32_t(lane) * 8u, acc[1 + 16 * r + j]);
      __syncwarp();
      double t0 = 0.0, t1 = 0.0;
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        t0 += lds_f64(row_s + uint32_t(k) * 8u);
        t1 += lds_f64(row_s + uint32_t(k + 1) * 8u);
      }
      const double t = t0 + t1;
      r01[r] = t + __shfl_xor_sync(0xffffffffu, t, 16);
    }
    double mom = lane < 16 ? r01[0] : r01[1];  // moment lane + 1 (this warp's share)
    double m33 = warp_sum(acc[33]), m34 = warp_sum(acc[34]);
    cam1 = __reduce_add_sync(0xffffffffu, cam1);
    if (WPS > 1) {  // the team's shares meet in shared memory, summed by the leader in warp order
      s_part[warp][mj] = mom;
      if (lane == 0) {
        s_part[warp][33] = m33;
        s_part[warp][34] = m34;
        s_part[warp][35] = double(cam1);
      }
      team_sync();
      if (sub == 0) {
        mom = 0.0;
        double e = 0.0;
#pragma unroll
        for (int w = 0; w < WPS; w++) {
          mom += s_part[lead + w][mj];
          if (lane < 3) e += s_part[lead + w][33 + lane];
        }
        m33 = __shfl_sync(0xffffffffu, e, 0);
        m34 = __shfl_sync(0xffffffffu, e, 1);
        cam1 = int(__shfl_sync(0xffffffffu, e, 2));
      }
    }
    if (n > 0 && sub == 0) {
      double* out = moments + size_t(s0 + sl) * kMomentStride;
      out[mj] = mom * scl;  // moments 1..32
      if (lane == 0) {
        out[0] = double(n);
        out[33] = m33 * s4;
        out[34] = m34 * s4;
        out[35] = double(cam1);
      }
    }
    sl = sl_next;
    head = head_next;
  }
}

// ---- small dense helpers (per warp, shared memory, lane-parallel where it is cheap) -----------
constexpr int kRankCap = 2048;  // neighbours per sample the rand() % n mode can rank (larger balls fall back)
__device__ int g_force_jacobi = 0;  // diagnostics (AG_FORCE_JACOBI): always diagonalise with the cyclic Jacobi
struct AxesSmem {
  double A[81];   // Schur complement, later C = L^-1 A L^-T (diagonalised in place only on the Jacobi fallback)
  double L[81];   // B, then its Cholesky factor
  double V[96];   // scratch of the eigen-solve: L D L^T factors (81 + 9), or the Jacobi eigenvectors
  double m[20];   // last column of M (9 entries); from [10]: the scaling S = D^(-1/2) of the reduction
  double par[10]; // quadric parameters in centred/scaled coordinates
  double T[28];   // weighted order-6 normal tensor
};
// non-deterministic normal mode: neighbours per sample the rank kernel can order, distance buckets of its counting sort
constexpr int kRankBuckets = 256;

// streams a sample's neighbour list (written by k_ball_search) 32 records per step, the next step's load
// already in flight; f(point, active)
template <typename F>
__device__ __forceinline__ void walk_list(const GPoint* __restrict__ list, int n, F&& f) {
  const int lane = threadIdx.x & 31;
  GPoint nxt;
  nxt.x = nxt.y = nxt.z = 0.f;
  nxt.tag = 0;
  if (lane < n) nxt = list[lane];
  for (int b = 0; b < n; b += 32) {
    const GPoint p = nxt;
    if (b + 32 + lane < n) nxt = list[b + 32 + lane];
    f(p, b + lane < n);
  }
}

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


// ---- smallest eigenpair of the reduced 9x9 pencil without diagonalising it ------------------------
// C (symmetric, positive semi-definite up to round-off) is only needed for its SMALLEST eigenpair
// (quadric.cpp:149-152).  lambda_1 is bracketed by multisection on the predicate "C - sigma I is positive
// definite": every lane factorises C - sigma_lane I = L D L^T in registers (fully unrolled, no communication)
// and reports whether all nine pivots stayed positive — a Cholesky test, stable exactly where it says yes — so
// one ballot narrows the bracket 33 times.  Eight rounds pin lambda_1 to ~1e-12 of the smallest diagonal
// entry from below; inverse iteration with that (positive definite) shift then converges in 3-4 solves.
// ~2.5k warp instructions instead of ~10k for the cyclic Jacobi, and almost no dependent shared-memory traffic.
__device__ __forceinline__ double fast_rcp(double x) {  // reciprocal to ~1 ulp: approximation + 2 Newton steps
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}
// L D L^T of C - sigma I from the row-major 9x9 in shared memory (lower triangle read).  Returns whether all
// pivots were positive.  STORE: lane 0 also leaves L (strict lower part, row-major 9x9) and 1/D in `fac`.
template <bool STORE>
__device__ __forceinline__ bool ldl9_shifted(const double* __restrict__ C, double sigma, double* fac, int lane,
                                             double tol = 0.0) {
  double a[9][9];  // lower triangle, compile-time indices only (registers)
#pragma unroll
  for (int i = 0; i < 9; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) a[i][j] = C[i * 9 + j] - (i == j ? sigma : 0.0);
  bool pd = true;
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const double d = a[k][k];
    pd = pd && (d > tol);
    const double inv = fast_rcp(d);
    if (STORE && lane == 0) fac[81 + k] = inv;
    double l[9];
#pragma unroll
    for (int i = k + 1; i < 9; i++) {
      l[i] = a[i][k] * inv;
      if (STORE && lane == 0) fac[i * 9 + k] = l[i];
    }
#pragma unroll
    for (int i = k + 1; i < 9; i++)
#pragma unroll
      for (int j = k + 1; j <= i; j++) a[i][j] = fma(-l[i], a[j][k], a[i][j]);
  }
  return pd;
}
// Returns true and the unit eigenvector y[9] of the smallest eigenvalue of C (row-major 9x9 in shared memory,
// symmetric); false if the bracket or the iteration did not behave (the caller then diagonalises C with the
// cyclic Jacobi).  `fac` = 90 doubles of shared scratch.  Every lane ends up with the same y.
__device__ bool smallest_eigvec9(const double* __restrict__ C, double* fac, int lane, double y[9]) {
  double dmin = C[0], dmax = C[0];
#pragma unroll
  for (int k = 1; k < 9; k++) {
    dmin = fmin(dmin, C[k * 9 + k]);
    dmax = fmax(dmax, C[k * 9 + k]);
  }
  if (!(dmax > 0.0) || !(dmax < 1e300)) return false;
  // lambda_1 <= every diagonal entry (Rayleigh quotient of a unit vector); >= -round-off
  double lo = -1e-9 * dmax, hi = dmin + 1e-9 * dmax;
  if (!__all_sync(0xffffffffu, ldl9_shifted<false>(C, lo, fac, lane))) return false;  // not PSD: leave it to Jacobi
#pragma unroll 1
  for (int round = 0; round < 8; round++) {
    const double step = (hi - lo) * (1.0 / 33.0);
    const double sigma = lo + step * double(lane + 1);
    const unsigned ok = __ballot_sync(0xffffffffu, ldl9_shifted<false>(C, sigma, fac, lane));
    const int j = __ffs(~ok) - 1;  // first shift that is not below lambda_1 (-1: all 32 are)
    const double lo_new = j == 0 ? lo : lo + step * double(j < 0 ? 32 : j);
    if (j >= 0) hi = lo + step * double(j + 1);
    lo = lo_new;
    if (!(hi - lo > 1e-15 * dmax)) break;
  }
  // inverse iteration with the shift at the lower end of the bracket (C - lo I is positive definite)
  __syncwarp();
  if (!__all_sync(0xffffffffu, ldl9_shifted<true>(C, lo, fac, lane))) return false;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 9; i++) y[i] = 1.0 + 0.0625 * double((i * 5) % 9);  // a fixed vector with no symmetry
  bool converged = false;
#pragma unroll 1
  for (int it = 0; it < 10 && !converged; it++) {
    double z[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {  // L z = y
      double v = y[i];
#pragma unroll
      for (int j = 0; j < i; j++) v = fma(-fac[i * 9 + j], z[j], v);
      z[i] = v;
    }
#pragma unroll
    for (int i = 0; i < 9; i++) z[i] *= fac[81 + i];  // D
#pragma unroll
    for (int i = 8; i >= 0; i--) {  // L^T w = z
      double v = z[i];
#pragma unroll
      for (int j = i + 1; j < 9; j++) v = fma(-fac[j * 9 + i], z[j], v);
      z[i] = v;
    }
    double nn = 0.0;
#pragma unroll
    for (int i = 0; i < 9; i++) nn = fma(z[i], z[i], nn);
    if (!(nn > 0.0) || !(nn < 1e300)) return false;
    const double inv = rsqrt(nn);
    // fix the sign by the largest component so successive iterates are comparable
    double big = 0.0;
#pragma unroll
    for (int i = 0; i < 9; i++)
      if (fabs(z[i]) > fabs(big)) big = z[i];
    const double sc = big < 0.0 ? -inv : inv;
    double diff = 0.0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
      const double v = z[i] * sc;
      diff = fmax(diff, fabs(v - y[i]));
      y[i] = v;
    }
    converged = it > 0 && diff < 4e-15;
  }
  return converged;
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

// ---- kernel 2: the 9x9 pencil and its smallest eigenpair -> quadric parameters -----------------------
// One warp per sample; needs only the 36 moments.  Output: 10 quadric parameters in centred / scaled
// coordinates + a validity flag, 12 doubles per sample.
__global__ void __launch_bounds__(kWarps * 32, 4)
k_taubin_solve(const RowIndex* __restrict__ rip, const int* __restrict__ indices, int s0, int n_samples_max,
               const int* __restrict__ d_count, const double* __restrict__ moments, double* __restrict__ par_out) {
  __shared__ AxesSmem sm_all[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int s = s0 + blockIdx.x * kWarps + warp;
  if (s >= n_samples_max || s >= *d_count) return;
  AxesSmem& sm = sm_all[warp];
  const int idx = indices[s];
  if (idx < 0 || idx >= rip->n_points) return;
  const double* mom = moments + size_t(s) * kMomentStride;
  const double n = mom[0];

  // --- build the reduced pencil: A = M[0:9,0:9] - m m^T / n, B = N[0:9,0:9]
  const double inv_n = 1.0 / n;
  for (int e = lane; e < 81; e += 32) {
    const PencilEntry t = g_pencil.e[e];
    sm.A[e] = mom[t.a_ij] - mom[t.a_i] * mom[t.a_j] * inv_n;
    double bsum = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++)
      if (k < t.nb) bsum += double(t.b_coef[k]) * mom[t.b_idx[k]];
    sm.L[e] = bsum;
  }
  if (lane < 9) sm.m[lane] = mom[g_pencil.e[lane * 9 + lane].a_i];
  __syncwarp();

  // --- B = G G^T with G = L sqrt(D): every lane factorises B = L D L^T redundantly in registers (no
  // communication, no divisions beyond nine reciprocals); lane 0 leaves L and 1/D in sm.V.  B is singular when
  // the neighbourhood is degenerate for the gradient form (e.g. voxel corners lying exactly in one lattice
  // plane); those directions have lambda = infinity (or 0/0) and must be excluded, which the robust branch
  // below does by working in range(B).
  bool ok = n >= 1.0;
  double dmax = 0.0;
  for (int i = 0; i < 9; i++) dmax = fmax(dmax, sm.L[i * 9 + i]);
  if (!(dmax > 0.0)) ok = false;
  const bool singular = ok && !__all_sync(0xffffffffu, ldl9_shifted<true>(sm.L, 0.0, sm.V, lane, 1e-10 * dmax));
  __syncwarp();
  if (!singular) {
    // --- C = G^-1 A G^-T = S (L^-1 A L^-T) S, S = D^(-1/2): unit-triangular solves (columns, then rows, in
    // parallel over nine lanes), then the diagonal scaling
    double rs = 0.0;  // lane i < 9: 1 / sqrt(D_i)
    if (lane < 9) {
      rs = sqrt(sm.V[81 + lane]);
      const int col = lane;
      double x[9];
#pragma unroll
      for (int i = 0; i < 9; i++) {
        double v = sm.A[i * 9 + col];
#pragma unroll
        for (int k = 0; k < i; k++) v = fma(-sm.V[i * 9 + k], x[k], v);
        x[i] = v;
      }
#pragma unroll
      for (int i = 0; i < 9; i++) sm.A[i * 9 + col] = x[i];
    }
    __syncwarp();
    if (lane < 9) {
      const int row = lane;  // solve L y = (row of X)^T
      double x[9];
#pragma unroll
      for (int i = 0; i < 9; i++) {
        double v = sm.A[row * 9 + i];
#pragma unroll
        for (int k = 0; k < i; k++) v = fma(-sm.V[i * 9 + k], x[k], v);
        x[i] = v;
      }
#pragma unroll
      for (int i = 0; i < 9; i++) sm.A[row * 9 + i] = x[i];
    }
    __syncwarp();
    if (lane < 9) sm.m[10 + lane] = rs;  // (sm.m has 10 + 10 entries: last column of M, then S)
    __syncwarp();
    // scale, symmetrise (round-off)
    double cnew[3];
    for (int e = lane, t = 0; e < 81; e += 32, t++) {
      const int i = e / 9, j = e % 9;
      cnew[t] = 0.5 * (sm.A[i * 9 + j] + sm.A[j * 9 + i]) * (sm.m[10 + i] * sm.m[10 + j]);
    }
    __syncwarp();
    for (int e = lane, t = 0; e < 81; e += 32, t++) sm.A[e] = cnew[t];
    __syncwarp();
    // keep L and S for the back-transformation: the eigen-solve below overwrites sm.V, so move L into sm.L
    for (int e = lane; e < 81; e += 32) sm.L[e] = sm.V[e];
    __syncwarp();
    // smallest eigenvalue (quadric.cpp:149-152): bracket + inverse iteration; the cyclic Jacobi is the fallback
    double yv[9];
    const bool fast = g_force_jacobi ? false : smallest_eigvec9(sm.A, sm.V, lane, yv);
    if (!fast) {
      __syncwarp();
      warp_jacobi9(sm.A, sm.V, lane);
      int mi = 0;
      for (int k = 1; k < 9; k++)
        if (sm.A[k * 9 + k] < sm.A[mi * 9 + mi]) mi = k;
#pragma unroll
      for (int i = 0; i < 9; i++) yv[i] = sm.V[i * 9 + mi];
    }
    __syncwarp();
    if (lane == 0) {  // u = G^-T y = L^-T (S y), j = -m.u/n
      double u[9];
#pragma unroll
      for (int i = 8; i >= 0; i--) {
        double v = yv[i] * sm.m[10 + i];
#pragma unroll
        for (int k = i + 1; k < 9; k++) v = fma(-sm.L[k * 9 + i], u[k], v);
        u[i] = v;
      }
#pragma unroll
      for (int i = 0; i < 9; i++) sm.par[i] = u[i];
    }
  } else {
    // --- robust branch: B = Q D Q^T, W = Q_k D_k^(-1/2) over the directions with D_i > 1e-11 D_max,
    //     C = W^T A W on range(B), smallest eigenpair y, u = W y
    // (sm.L still holds B: the factorisation above works in registers)
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
  if (lane < 10) par_out[size_t(s) * 12 + lane] = sm.par[lane];
  if (lane == 0) par_out[size_t(s) * 12 + 10] = ok ? 1.0 : 0.0;

}

// ---- kernel 2b (production normal mode only): which 50 neighbours the reference's rand() % n picks -------
// The picks index the kd-tree's result order = ascending (distance, index).  Independent of the fit, so this
// runs on a second stream next to the moments / solve kernels.  One warp per sample, everything in shared
// memory: binary32 distances of the neighbours, a STABLE counting sort into 256 distance buckets (a monotone
// function of the distance), then every lane finishes its buckets by insertion on (distance, position) — the list is
// in index order, so the position breaks distance ties.  Voxel corners sit on a lattice, so whole groups of
// neighbours share a distance exactly; the stable scatter leaves them in order and the insertion pass only moves
// elements of buckets that mix different distances.  Output: 50 list positions per sample (samples with at most 50 neighbours evaluate all of
// them and get no picks).
template <int CAP>
__global__ void __launch_bounds__(kWarps * 32, 8)
k_rank_picks(const GPoint* __restrict__ pts_c, const RowIndex* __restrict__ rip, const int* __restrict__ indices,
             int s0, int n_samples_max, const int* __restrict__ d_count, const GPoint* __restrict__ pool, int stride,
             const int2* __restrict__ nn_counts, float r2_f, const uint32_t* __restrict__ rand_raw,
             const int* __restrict__ rand_off, int off_first, int off_step, unsigned short* __restrict__ picks_out,
             const int* __restrict__ carry_in, int* __restrict__ carry_out) {
  extern __shared__ __align__(16) unsigned char s_rank[];
  __shared__ int s_cnt[kWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sl = blockIdx.x * kWarps + warp;
  const int s = s0 + sl;
  // carry_in != null (launches of a few thousand samples): no separate layout of the rand() stream — the CTA counts
  // the earlier samples of this launch with more than 50 neighbours itself (coalesced reads of the counts, at most
  // a few per thread), so sample s reads draws [50 k_s, 50 k_s + 50) with k_s = carry + that count; the last CTA
  // leaves the carry of the next launch in carry_out (a different word: the other CTAs still read carry_in)
  int self_off = 0;
  if (carry_in) {
    const int cnt_valid = min(*d_count, n_samples_max);
    const int first = s0 + int(blockIdx.x) * kWarps;
    int c = 0;
    for (int i = s0 + int(threadIdx.x); i < first; i += kWarps * 32) c += (i < cnt_valid && nn_counts[i].x > 50) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) s_cnt[warp] = c;
    __syncthreads();
    int before = *carry_in;
    for (int w = 0; w < kWarps; w++) before += s_cnt[w];
    int mine = 0;  // the samples of this CTA in front of this warp's
    for (int w = 0; w < kWarps; w++) {
      const int f = (first + w < cnt_valid && nn_counts[first + w].x > 50) ? 1 : 0;
      if (w < warp) mine += f;
      if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && w == kWarps - 1) {
        int tot = before;
        for (int w2 = 0; w2 < kWarps; w2++) tot += (first + w2 < cnt_valid && nn_counts[first + w2].x > 50) ? 1 : 0;
        *carry_out = tot;
      }
    }
    self_off = before + mine;
  }
  if (s >= n_samples_max || s >= *d_count) return;
  const int idx = indices[s];
  if (idx < 0 || idx >= rip->n_points) return;
  const GPoint* list = pool + size_t(sl) * size_t(stride);
  const int n_list = nn_counts[s].x;
  if (!(n_list > 50 && n_list <= CAP)) return;
  const GPoint q = pts_c[idx];
  float* d2 = reinterpret_cast<float*>(s_rank) + size_t(warp) * CAP;
  int* cur = reinterpret_cast<int*>(s_rank + size_t(kWarps) * CAP * 4) + warp * kRankBuckets;
  unsigned short* order = reinterpret_cast<unsigned short*>(s_rank + size_t(kWarps) * (CAP * 4 + kRankBuckets * 4)) + size_t(warp) * CAP;
  constexpr int nb = kRankBuckets;
  const float bscale = float(nb) / r2_f;
  for (int i = lane; i < nb; i += 32) cur[i] = 0;
  __syncwarp();
  for (int i0 = lane; i0 < n_list; i0 += 128) {  // four independent loads in flight per lane
    GPoint p[4];
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i0 + 32 * u < n_list) p[u] = list[i0 + 32 * u];
#pragma unroll
    for (int u = 0; u < 4; u++)
      if (i0 + 32 * u < n_list) {
        const float d = dist2_flann(q.x, q.y, q.z, p[u].x, p[u].y, p[u].z);
        d2[i0 + 32 * u] = d;
        atomicAdd(&cur[min(nb - 1, int(d * bscale))], 1);
      }
  }
  __syncwarp();
  int first[8];  // start of this lane's 8 consecutive buckets, kept for the insertion pass
  {
    int c[8], tot = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      c[k] = cur[8 * lane + k];
      tot += c[k];
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    int run = incl - tot;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      first[k] = run;
      cur[8 * lane + k] = run;
      run += c[k];
    }
  }
  __syncwarp();
  // STABLE scatter: 32 consecutive list positions per step, lanes of the same bucket take consecutive places in lane
  // order (match.any), so neighbours with EQUAL distances — whole groups of them on the voxel lattice — land in
  // list order, which is the kd-tree's tie-break; only buckets that mix different distances are left to sort
  for (int i0 = 0; i0 < n_list; i0 += 32) {
    const int i = i0 + lane;
    const bool in = i < n_list;
    const int b = in ? min(nb - 1, int(d2[i] * bscale)) : nb + lane;  // (idle lanes: a bucket of their own)
    const unsigned grp = __match_any_sync(0xffffffffu, b);
    const int rank = __popc(grp & ((1u << lane) - 1u));
    const int base = in ? cur[b] : 0;
    __syncwarp();
    if (in) {
      order[base + rank] = (unsigned short)i;
      if (rank == __popc(grp) - 1) cur[b] = base + rank + 1;
    }
    __syncwarp();
  }
  // cur[b] is now the end of bucket b; this lane owns buckets 8 lane .. 8 lane + 7 = one contiguous range
  {
    const int lo = first[0], hi = cur[8 * lane + 7];
    // insertion sort of the whole range: elements only ever move inside their bucket (bucket = monotone in d)
    for (int a = lo + 1; a < hi; a++) {
      const unsigned short e = order[a];
      const float de = d2[e];
      int t = a - 1;
      while (t >= lo) {
        const unsigned short f = order[t];
        const float df = d2[f];
        if (df < de || (df == de && f < e)) break;
        order[t + 1] = f;
        t--;
      }
      order[t + 1] = e;
    }
  }
  __syncwarp();
  const uint32_t* raw = rand_raw + size_t(50) * size_t(carry_in ? self_off : rand_off[off_first + s * off_step]);
#pragma unroll
  for (int u = 0; u < 2; u++) {
    const int t = lane + 32 * u;
    if (t < 50) picks_out[size_t(s) * 50 + t] = order[raw[t] % uint32_t(n_list)];
  }
}

// ---- kernel 3: gradient normals at the evaluation points, curvature axis, frame ---------------------------
__global__ void __launch_bounds__(kWarps * 32, 4)
k_axes_finish(GPoint* pts_c, const RowIndex* __restrict__ rip,
              const int* __restrict__ indices, int s0, int n_samples_max, const int* __restrict__ d_count,
              const GPoint* __restrict__ pool, int stride, const int2* __restrict__ nn_counts, double inv_r,
              const double* __restrict__ moments, const double* __restrict__ par_in,
              double cam0x, double cam0y, double cam0z, double cam1x, double cam1y, double cam1z,
              ag_frame* __restrict__ frames, double* normals_out /* may be null */,
              const unsigned short* __restrict__ picks_in /* null = deterministic normals */,
              unsigned long long* __restrict__ counters, float4* __restrict__ sample_q) {
  __shared__ double s_T[kWarps][28];  // weighted order-6 normal tensor
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sl = blockIdx.x * kWarps + warp;
  const int s = s0 + sl;
  if (s >= n_samples_max) return;
  double* sT = s_T[warp];
  const int idx = s < *d_count ? indices[s] : -1;
  const bool is_sample = idx >= 0 && idx < rip->n_points;
  // what the hand sweep needs to know about sample s in ONE record (it would otherwise chase indices -> cloud):
  // x, y, z and (index << 1 | camera); -1 for a slot that holds no sample
  if (sample_q && lane == 0) {
    float4 v = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    if (is_sample) {
      const GPoint qs = pts_c[idx];
      v = make_float4(qs.x, qs.y, qs.z, __int_as_float((idx << 1) | int(qs.tag & kTagCamBit)));
    }
    sample_q[s] = v;
  }
  if (!is_sample) return;
  const GPoint* list = pool + size_t(sl) * size_t(stride);
  const int2 nnc = nn_counts[s];
  const int n_list = nnc.x;
  const GPoint q = pts_c[idx];
  if (lane == 0) {
    // totals of the neighbour / candidate counts (reported through ag_timings) and, when the normal goes to
    // cloud_normals_, the tag bit that lets the sweep fetch only normals that exist
    atomicAdd(&counters[0], (unsigned long long)nnc.x);
    atomicAdd(&counters[1], (unsigned long long)nnc.y);
    if (normals_out) atomicOr(&pts_c[idx].tag, kTagNormalBit);
  }
  const double qx = double(q.x), qy = double(q.y), qz = double(q.z);
  const double* mom = moments + size_t(s) * kMomentStride;
  const double n = mom[0];
  ag_frame F;
  F.num_neighbors = int(n);
  int major = (mom[35] > n - mom[35]) ? 1 : 0;  // quadric.cpp:217-226 (tie -> camera 0)
  // the reference's production mode (is_deterministic = false, quadric.cpp:177-192): with more than 50
  // neighbours the normals are evaluated at 50 picks rand() % n of the (distance, index)-sorted neighbour list
  const bool picks = picks_in != nullptr && n_list > 50 && n_list <= kRankCap;
  double par[9];
#pragma unroll
  for (int i = 0; i < 9; i++) par[i] = par_in[size_t(s) * 12 + i];
  const bool ok = par_in[size_t(s) * 12 + 10] != 0.0;

  // --- walk 2: C = sum g g^T and T = sum g^(x6)
  double acc[34];
#pragma unroll
  for (int i = 0; i < 34; i++) acc[i] = 0.0;
  auto accumulate = [&](const GPoint& p) {
    const double x = (double(p.x) - qx) * inv_r, y = (double(p.y) - qy) * inv_r, z = (double(p.z) - qz) * inv_r;
    double gn[3], m6[28];
    quad_normal(par, x, y, z, gn);
    acc[0] += gn[0] * gn[0]; acc[1] += gn[1] * gn[1]; acc[2] += gn[2] * gn[2];
    acc[3] += gn[0] * gn[1]; acc[4] += gn[1] * gn[2]; acc[5] += gn[0] * gn[2];
    monomials6(gn, m6);
#pragma unroll
    for (int t = 0; t < 28; t++) acc[6 + t] += m6[t];
  };
  GPoint pick[2];  // picks t = lane and t = lane + 32 (t < 50)
  pick[0] = pick[1] = q;
  if (picks) {
    int cam1 = 0;
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int t = lane + 32 * u;
      if (t < 50) {
        pick[u] = list[picks_in[size_t(s) * 50 + t]];
        cam1 += int(pick[u].tag & kTagCamBit);
        accumulate(pick[u]);
      }
    }
    cam1 = __reduce_add_sync(0xffffffffu, cam1);
    major = cam1 > 50 - cam1 ? 1 : 0;  // majority over the picks (quadric.cpp:217-226)
  } else {
    walk_list(list, n_list, [&](const GPoint& p, bool active) {
      if (active) accumulate(p);
    });
  }
  F.majority_cam = major;
  // warp reduction by recursive halving: lane i holds sum i (i < 32); sums 32, 33 by butterflies
  {
    const double r32 = warp_reduce_transpose32(acc);
    const double a32 = warp_sum(acc[32]), a33 = warp_sum(acc[33]);
#pragma unroll
    for (int i = 0; i < 6; i++) acc[i] = __shfl_sync(0xffffffffu, r32, i);  // sum g g^T, needed by every lane
    if (lane >= 6) sT[lane - 6] = r32 * c_multinomial6[lane - 6];          // T_0 .. T_25
    if (lane == 0) {
      sT[26] = a32 * c_multinomial6[26];
      sT[27] = a33 * c_multinomial6[27];
    }
  }
  double w3[3], V3[3][3];
  eig3(acc, w3, V3);
  int m3 = 0;
  if (w3[1] < w3[m3]) m3 = 1;
  if (w3[2] < w3[m3]) m3 = 2;  // quadric.cpp:278-280
  double ax[3] = {V3[0][m3], V3[1][m3], V3[2][m3]};
  __syncwarp();

  // --- walk 3: j* = argmax_j sum_i (g_i.g_j)^6 = argmax_j <T, g_j^(x6)>, first max in the reference's
  // (distance, index) order; index order == (camera, x, y, z) order of the voxel list.  In the picks mode the
  // first max is in pick order t.
  double bestS = -1.0;
  float bestD = 3.0e38f;
  GPoint bestP;
  bestP.x = bestP.y = bestP.z = 3.0e38f;
  bestP.tag = 1u;
  double bestG[3] = {0, 0, 0};
  auto before = [](const GPoint& a, const GPoint& b) {  // a precedes b in the voxel list
    const unsigned ca = a.tag & kTagCamBit, cb = b.tag & kTagCamBit;
    if (ca != cb) return ca < cb;
    if (a.x != b.x) return a.x < b.x;
    if (a.y != b.y) return a.y < b.y;
    return a.z < b.z;
  };
  auto score = [&](const GPoint& p, float d, const GPoint& key) {
    const double x = (double(p.x) - qx) * inv_r, y = (double(p.y) - qy) * inv_r, z = (double(p.z) - qz) * inv_r;
    double gn[3], m6[28];
    quad_normal(par, x, y, z, gn);
    monomials6(gn, m6);
    double S = 0.0;
#pragma unroll
    for (int t = 0; t < 28; t++) S += sT[t] * m6[t];
    const bool better = S > bestS || (S == bestS && (d < bestD || (d == bestD && before(key, bestP))));
    if (better) {
      bestS = S; bestD = d; bestP = key;
      bestG[0] = gn[0]; bestG[1] = gn[1]; bestG[2] = gn[2];
    }
  };
  if (picks) {
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int t = lane + 32 * u;
      if (t < 50) {  // order key = pick number t (encoded so that `before` compares t)
        GPoint key;
        key.x = float(t);
        key.y = key.z = 0.f;
        key.tag = 0u;
        score(pick[u], 0.f, key);
      }
    }
  } else {
    walk_list(list, n_list, [&](const GPoint& p, bool active) {
      if (active) score(p, dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z), p);
    });
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double oS = __shfl_xor_sync(0xffffffffu, bestS, o);
    const float oD = __shfl_xor_sync(0xffffffffu, bestD, o);
    GPoint oP;
    oP.x = __shfl_xor_sync(0xffffffffu, bestP.x, o);
    oP.y = __shfl_xor_sync(0xffffffffu, bestP.y, o);
    oP.z = __shfl_xor_sync(0xffffffffu, bestP.z, o);
    oP.tag = __shfl_xor_sync(0xffffffffu, bestP.tag, o);
    const double g0 = __shfl_xor_sync(0xffffffffu, bestG[0], o);
    const double g1 = __shfl_xor_sync(0xffffffffu, bestG[1], o);
    const double g2 = __shfl_xor_sync(0xffffffffu, bestG[2], o);
    const bool better = oS > bestS || (oS == bestS && (oD < bestD || (oD == bestD && before(oP, bestP))));
    if (better) {
      bestS = oS; bestD = oD; bestP = oP;
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

}  // namespace

// non-deterministic normal mode: sample s consumes rand() draws [50 k_s, 50 k_s + 50), k_s = number of earlier
// samples (in sample order, continuing across the launches of one call) with more than 50 neighbours
__global__ void __launch_bounds__(1024)
k_rand_offsets(const int2* __restrict__ nn_counts, int s0, int m, const int* __restrict__ d_count, int* __restrict__ rand_off,
               int* __restrict__ carry) {
  __shared__ int s_warp[32];
  rand_offsets_block(nn_counts, s0, m, *d_count, rand_off, carry, s_warp);
}

// glibc rand() outputs after srand(1) — what a process that never seeds gets from the reference's
// `rand() % indices.size()` (quadric.cpp:184): random_r TYPE_3, word_k = word_(k-31) + word_(k-3), output >> 1
static void glibc_rand_stream(size_t count, std::vector<uint32_t>& out) {
  std::vector<uint32_t> w(34);
  int64_t x = 1;
  w[0] = 1u;
  for (int i = 1; i < 31; i++) {
    x = (16807 * x) % 2147483647;
    if (x < 0) x += 2147483647;
    w[i] = uint32_t(x);
  }
  for (int i = 31; i < 34; i++) w[i] = w[i - 31];
  w.resize(344 + count);
  for (size_t i = 34; i < w.size(); i++) w[i] = w[i - 31] + w[i - 3];
  out.resize(count);
  for (size_t k = 0; k < count; k++) out[k] = w[344 + k] >> 1;
}

// neighbour-pool records per sample for a radius: the lattice-ball bound, capped at four times what a
// surface sampled by the voxel lattice puts into the ball (a ball holding more sets kErrBallOverflow)
static int ball_stride(double radius, double voxel, bool two_cams) {
  const double cells = radius / (voxel > 0 ? voxel : 0.003);
  const double ball = 4.18879 * (cells + 0.87) * (cells + 0.87) * (cells + 0.87) + 32.0;
  const double surf = std::max(288.0, 1024.0 * (cells / 10.0) * (cells / 10.0));
  const double v = std::min(ball, surf) * (two_cams ? 2.0 : 1.0);
  return int((std::min(v, 1.0e6) + 31.0) / 32.0) * 32;
}

int quadric_rand_reset(Ctx* c) {
  c->rand_consumed_bound = 0;
  c->rand_slot = 0;
  if (c->params.deterministic_normals != 0) return AG_OK;
  // the carry lives in its own buffer: rand_off is re-sized per launch (DevBuf::reserve does not keep contents)
  if (c->rand_carry.reserve(16)) return AG_ERR_CUDA;
  if (c->fold_resets & 4u) return AG_OK;  // (ag_localize: zeroed by k_init_state, which runs next)
  AG_CUDA_CHECK(cudaMemsetAsync(c->rand_carry.p, 0, 16, c->stream));
  return AG_OK;
}

int fit_quadrics_device(Ctx* c, const int* d_indices, int n, const int* d_count, double radius, ag_frame* d_frames,
                        bool write_normals, const RandShare* share) {
  if (n <= 0) return AG_OK;
  const int stride = ball_stride(radius, c->params.voxel_size, c->two_cams);
  // samples are processed in chunks whose neighbour pool stays below kPoolBytes (one chunk for every
  // BASELINE config; only all-points passes over multi-million-point inputs take more)
  constexpr size_t kPoolBytes = size_t(4) << 30;
  const int chunk = int(std::min<size_t>(size_t(n), std::max<size_t>(kWarps, kPoolBytes / (size_t(stride) * sizeof(GPoint)))));
  if (c->moments.reserve(size_t(n) * kMomentStride * sizeof(double)) || c->nn_counts.reserve(size_t(n) * 8) ||
      c->quad_par.reserve(size_t(n) * 12 * sizeof(double)) || c->sample_q.reserve(size_t(n) * sizeof(float4)) ||
      c->counters.reserve(64) || c->nbr_pool.reserve(size_t(chunk) * size_t(stride) * sizeof(GPoint)) ||
      c->nbr_heads.reserve(size_t(chunk) * sizeof(float4)))
    return AG_ERR_CUDA;
  const float r2 = float(radius * radius);  // PCL hands radius*radius to FLANN as float
  const double rpad = sqrt(double(r2)) * (1.0 + 1e-5) + 1e-7;
  // coordinates are centred on the sample and scaled by the power of two nearest to 1/r (the fit is
  // invariant under translation and uniform scale; a power of two makes the scaling exact)
  const double inv_r = ldexp(1.0, int(lrint(log2(1.0 / radius))));
  unsigned long long* ctr = c->counters.as<unsigned long long>();
  RowIndex* ri = c->row_index.as<RowIndex>();
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_ball_search<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(k_ball_moments, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(k_rank_picks<kRankCap>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         int(kWarps * (kRankCap * 6 + kRankBuckets * 4)));
    cudaFuncSetAttribute(k_rank_picks<1024>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(k_rank_picks<kRankCap>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (getenv("AG_FORCE_JACOBI")) {
      const int one = 1;
      cudaMemcpyToSymbol(g_force_jacobi, &one, sizeof(one));
    }
    attr_set = true;
  }
  const HandConst& h = c->hand;
  // non-deterministic normal mode (ag_params.deterministic_normals = 0): the rand() stream of an unseeded process,
  // 50 draws per sample with more than 50 neighbours; the stream restarts with every ag_localize / ag_fit_quadrics
  const bool rand_mode = c->params.deterministic_normals == 0;
  const uint32_t* d_rand = nullptr;
  int* d_rand_off = nullptr;
  if (rand_mode) {
    const size_t need = size_t(50) * (size_t(c->rand_consumed_bound) + size_t(share ? std::max(share->n_all, n) : n));
    if (c->rand_count < need) {
      std::vector<uint32_t> hs;
      glibc_rand_stream(need + need / 2, hs);
      if (c->rand_raw.reserve(hs.size() * 4)) return AG_ERR_CUDA;
      AG_CUDA_CHECK(cudaMemcpyAsync(c->rand_raw.p, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice, c->stream));
      AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
      c->rand_count = hs.size();
    }
    const int n_off = share ? std::max(share->n_all, n) : n;
    if (c->rand_off.reserve(size_t(n_off) * 4 + 16) || c->rand_carry.reserve(16) || c->picks.reserve(size_t(n) * 100) ||
        (share && c->nn_counts_all.reserve(size_t(share->n_all) * 8)))
      return AG_ERR_CUDA;
    d_rand = c->rand_raw.as<uint32_t>();
    d_rand_off = c->rand_off.as<int>();
    c->rand_consumed_bound += share ? std::max(share->n_all, n) : n;
  }
  if (rand_mode && share) {
    // This context fits a SHARE of the call's samples (ag_params.shard_*), but the reference's rand() stream is
    // consumed by every sample with more than 50 neighbours in sample order: that one bit of ALL samples is computed
    // here (k_ball_over50), so that each of this share's samples finds the slice of the stream the unsharded call gives it
    // (on the side stream, next to this share's own search: only the pick ranking needs the result)
    AG_CUDA_CHECK(cudaEventRecord(c->ev_fork0, c->stream));
    AG_CUDA_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_fork0, 0));
    const int blocks_all = (share->n_all + kWarps - 1) / kWarps;
    k_ball_over50<<<blocks_all, kWarps * 32, 0, c->stream2>>>(c->vox.as<GPoint>(), c->row_ptr.as<int>(), c->col_ptr.as<int>(),
                                                             c->row_index.as<RowIndex>(), share->d_all, share->n_all,
                                                             share->d_count_all, r2, rpad, c->nn_counts_all.as<int2>());
    k_rand_offsets<<<1, 1024, 0, c->stream2>>>(c->nn_counts_all.as<int2>(), 0, share->n_all, share->d_count_all, d_rand_off,
                                               c->rand_carry.as<int>() + c->rand_slot);
    c->launches += 2;
  }
  // search + moments: ONE kernel for launches of a few thousand samples (latency bound: 23.7 vs 26 us at 2000
  // samples), search -> neighbour lists -> streaming moments kernel for launches that fill the machine (the fused
  // kernel needs 128 registers = 16 warps/SM and hides its load latencies worse: 0.27 vs 0.24 ms at 82k samples).
  // AG_MOMENTS=split|fused forces one variant (measurements).
  static const int forced_variant = [] {
    const char* e = getenv("AG_MOMENTS");
    return !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 'f' ? 2 : 0));
  }();
  static const bool fold_offsets = [] {  // AG_RAND_FOLD=0|1 (measurements): rand() stream layout inside k_ball_moments
    const char* e = getenv("AG_RAND_FOLD");
    return e ? atoi(e) != 0 : true;
  }();
  for (int s0 = 0; s0 < n; s0 += chunk) {
    const int m = std::min(chunk, n - s0);
    const bool split = forced_variant ? forced_variant == 1 : m > 16384;
    const int blocks = (m + kWarps - 1) / kWarps;
    const bool timed = s0 == 0;  // ag_timings reports the kernels of the first chunk
    if (timed) record_event(c, c->ev_k[0]);
    if (split)
      k_ball_search<true><<<blocks, kWarps * 32, 0, c->stream>>>(c->vox.as<GPoint>(), c->row_ptr.as<int>(),
                                                           c->col_ptr.as<int>(), ri, d_indices, s0, s0 + m, d_count, r2,
                                                           rpad, c->nbr_pool.as<GPoint>(), stride,
                                                           c->nn_counts.as<int2>(), c->nbr_heads.as<float4>());
    else
      k_ball_moments<<<blocks, kFusedWarps * 32, 0, c->stream>>>(c->vox.as<GPoint>(), c->row_ptr.as<int>(),
                                                                 c->col_ptr.as<int>(), ri, d_indices, s0, s0 + m, d_count,
                                                                 r2, rpad, c->nbr_pool.as<GPoint>(), stride,
                                                                 c->nn_counts.as<int2>(), inv_r, c->moments.as<double>(),
                                                                 (fold_offsets && rand_mode && !share && m > 4096) ? d_rand_off : nullptr,
                                                                 c->rand_carry.as<int>() + c->rand_slot);
    if (timed) record_event(c, c->ev_k[1]);
    if (rand_mode) {
      // the reference's rand() % n picks depend on the neighbour lists only: ranked on a second stream while
      // this one accumulates the moments and solves the eigenproblem (fork / join by events, also under capture)
      AG_CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
      AG_CUDA_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
      // layout of the rand() stream: small launches let k_rank_picks count for itself (no extra node, no serial tail),
      // large ones fold the scan into k_ball_moments' last CTA (or run k_rand_offsets after the two-kernel search)
      const bool self_off = !share && !split && m <= 4096 && fold_offsets;
      int* carry = c->rand_carry.as<int>();
      const int* carry_in = self_off ? carry + c->rand_slot : nullptr;
      int* carry_out = self_off ? carry + (c->rand_slot ^ 2) : nullptr;
      if (!share && !self_off && (split || !fold_offsets))
        k_rand_offsets<<<1, 1024, 0, c->stream2>>>(c->nn_counts.as<int2>(), s0, m, d_count, d_rand_off, carry + c->rand_slot);
      const int off_first = share ? share->first : 0, off_step = share ? share->step : 1;
      const size_t rank_smem = size_t(kWarps) * (size_t(stride <= 1024 ? 1024 : kRankCap) * 6 + kRankBuckets * 4);
      if (stride <= 1024)
        k_rank_picks<1024><<<blocks, kWarps * 32, rank_smem, c->stream2>>>(
            c->vox.as<GPoint>(), ri, d_indices, s0, s0 + m, d_count, c->nbr_pool.as<GPoint>(), stride,
            c->nn_counts.as<int2>(), r2, d_rand, d_rand_off, off_first, off_step, c->picks.as<unsigned short>(), carry_in,
            carry_out);
      else
        k_rank_picks<kRankCap><<<blocks, kWarps * 32, rank_smem, c->stream2>>>(
            c->vox.as<GPoint>(), ri, d_indices, s0, s0 + m, d_count, c->nbr_pool.as<GPoint>(), stride,
            c->nn_counts.as<int2>(), r2, d_rand, d_rand_off, off_first, off_step, c->picks.as<unsigned short>(), carry_in,
            carry_out);
      if (self_off) c->rand_slot ^= 2;
      AG_CUDA_CHECK(cudaEventRecord(c->ev_join, c->stream2));
      c->launches += (!share && !self_off && (split || !fold_offsets)) ? 2 : 1;
    }
    if (split) {
      // warps per sample.  Measured on B200: a 2000-sample launch takes 10.2 / 11.3 / 14.3 us with 1 / 2 / 4 warps
      // per sample — it is bounded by fixed costs (launch, first-touch latencies, reduction), not by one warp
      // walking a whole list — so teams are only used when asked for (AG_MOM_WPS = 2 | 4, a tuning knob).
      static const int forced = [] {
        const char* e = getenv("AG_MOM_WPS");
        return e ? atoi(e) : 0;
      }();
      const int wps = (forced == 2 || forced == 4) ? forced : 1;
      const float4* hd = c->nbr_heads.as<float4>();
      const GPoint* pl = c->nbr_pool.as<GPoint>();
      double* mo = c->moments.as<double>();
      const int teams_per_cta = kWarps / wps;
      const int grid = std::min((m + teams_per_cta - 1) / teams_per_cta, kNumSMs * 4);
      if (wps == 4) k_taubin_moments<4><<<grid, kWarps * 32, 0, c->stream>>>(s0, m, hd, pl, stride, inv_r, mo);
      else if (wps == 2) k_taubin_moments<2><<<grid, kWarps * 32, 0, c->stream>>>(s0, m, hd, pl, stride, inv_r, mo);
      else k_taubin_moments<1><<<grid, kWarps * 32, 0, c->stream>>>(s0, m, hd, pl, stride, inv_r, mo);
    }
    if (timed) record_event(c, c->ev_k[2]);
    k_taubin_solve<<<blocks, kWarps * 32, 0, c->stream>>>(ri, d_indices, s0, s0 + m, d_count, c->moments.as<double>(),
                                                           c->quad_par.as<double>());
    if (rand_mode) AG_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));  // the picks are ready
    k_axes_finish<<<blocks, kWarps * 32, 0, c->stream>>>(
        c->vox.as<GPoint>(), ri, d_indices, s0, s0 + m, d_count, c->nbr_pool.as<GPoint>(), stride,
        c->nn_counts.as<int2>(), inv_r, c->moments.as<double>(), c->quad_par.as<double>(), h.cam[0][0], h.cam[0][1],
        h.cam[0][2], h.cam[1][0], h.cam[1][1], h.cam[1][2], d_frames, write_normals ? c->normals.as<double>() : nullptr,
        rand_mode ? c->picks.as<unsigned short>() : nullptr, ctr, c->sample_q.as<float4>());
    if (timed) record_event(c, c->ev_k[3]);
    c->launches += split ? 4 : 3;
  }
  AG_CUDA_CHECK(cudaGetLastError());
  return AG_OK;
}

}  // namespace ag
