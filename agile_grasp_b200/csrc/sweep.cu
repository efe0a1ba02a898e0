// sweep.cu — rotating-hand / finger-placement sweep, antipodal test and grasp-image rasterisation.
// One CTA per sample, one warp per hand orientation.
//
// Replaces (reference paths): HandSearch::findHands private (src/agile_grasp/hand_search.cpp:116-206,
// radius search :147, float centring :154-160), RotatingHand::transformPoints / evaluateHand
// (src/agile_grasp/rotating_hand.cpp:19-177), FingerHand (src/agile_grasp/finger_hand.cpp:3-233),
// Antipodal::evaluateGrasp (src/agile_grasp/antipodal.cpp:12-86), Localization::filterHands
// (src/agile_grasp/localization.cpp:364-388) and the image half of Learning::createInstance /
// convertToImage (src/agile_grasp/learning.cpp:320-400).
//
// GPU formulation.  The reference re-scans the slab points for every finger slot at
// every deepening step.  Every boolean it derives is an existence test "is there a point with
// y < d_t and x in some interval", so one pass per orientation that ORs 20-bit slot masks into the
// 11 depth levels d_t reproduces all of them exactly: the same binary64 comparisons on the same
// binary64 values, evaluated without FMA contraction (this file is compiled with -fmad=false and
// the products/sums are written in the reference's left-to-right order).

#include <algorithm>
#include <cstdlib>
#include <vector>


#include "ag_internal.h"

namespace ag {

namespace {

constexpr int kThreads = 256;       // 8 warps = 8 orientations (rotating_hand.cpp:13)
constexpr int kSlabCapSmall = 1920; // slab points kept in shared memory by the common kernel (37.5 KB; 55 KB per CTA, 4 CTAs / SM)
constexpr int kSlabCapBig = 9600;   // fallback instantiation for dense neighbourhoods (187.5 KB, 1 CTA / SM)
constexpr int kHeavyImage = 160;    // occupied pixels from which a grasp image goes to the front of the scorer's list
constexpr int kStage = 1024;        // candidates staged per pass of the ball gather (16 KB, reused by phase B)

struct SweepArgs {
  const GPoint* pts;
  const int* row_ptr;
  const int* col_ptr;
  const RowIndex* ri;
  const int* indices;
  const ag_frame* frames;
  const double* normals;  // 3 per voxel point, indexed by original index
  ag_grasp* grasps;       // [n_samples * 8]
  uint8_t* valid;         // [n_samples * 8]
  uint32_t* images;       // [n_samples * 8 * 250]
  int* debug;             // [n_samples * 8] or null
  int* slab_counts;       // [n_samples] or null
  unsigned long long* counters;
  int* hyp_list;          // unordered list of the (sample, orientation) slots that hold a hypothesis (for the scorer)
  int* hyp_count;         // [0] entries filled from the front, [1] from the back
  int heavy_image;        // occupied pixels from which a grasp image goes to the front
  int* overflow;          // [0] = number of samples whose slab exceeded the capacity, [1..] their slots
  const int* sample_list; // if non-null: blockIdx.x indexes this list of sample slots (large-slab pass)
  const int* sample_count;// if non-null: device-side length of sample_list (inline large-slab pass: CTAs beyond it leave)
  const float4* sample_q; // if non-null: x, y, z, (index << 1 | camera) of every sample slot (left by the fit; -1: no sample)
  int n_samples;
  float r2;
  double rpad;
  int filter_boundaries;
  double workspace[6];
};

__device__ __forceinline__ double dot3e(const double a[3], const double b[3]) {
  return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]);  // Eigen's unrolled 3-vector reduction order
}

// ---- PTX helpers: mbarrier + TMA bulk copy (global -> shared) -------------------------------------
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- the finger-slot predicates of one slab point, literally (finger_hand.cpp:54-95) ---------------
// in bit i  : the point lies strictly inside slot i,            sp[i] < x < sp[i] + finger_width
// side bit i: the point lies beyond slot i on the object side,  x > sp[i] + w (i <= 10)  |  x < sp[i] (i > 10)
__host__ __device__ inline unsigned long long slot_masks_exact(const double* spacing, double fw, double x) {
  unsigned in_mask = 0u, sd_mask = 0u;
  for (int i = 0; i < 20; i++) {
    const double lo = spacing[i], hi = spacing[i] + fw;  // finger_hand.cpp:56-57
    if (x > lo && x < hi) in_mask |= 1u << i;
    const bool side = (i <= 10) ? (x > hi) : (x < lo);  // finger_hand.cpp:72-82
    if (side) sd_mask |= 1u << i;
  }
  return (static_cast<unsigned long long>(sd_mask) << 32) | in_mask;
}

constexpr float kLutMargin = 1e-4f;  // distance (in slot steps / depth steps) from a threshold below which the
                                     // exact comparisons decide instead of the table

// Shared memory of one CTA.  The 16 KB `u` block is the TMA staging buffer of the ball gather (phase A) and
// then, per orientation, the depth-level slot masks (two lanes share a word: reductions) and — once those are
// consumed — the grasp image in the same bytes (phase B).
struct SweepShared {
  union {
    GPoint stage[kStage];
    struct {
      unsigned long long lvl[8][12][16];  // per warp, depth level and lane pair: (side << 32 | in) slot masks
    } b;
  } u;
  unsigned long long lut[AG_SWEEP_LUT];  // slot masks per zone between two slot edges
  double bite[12];
  int run_start[128], run_pre[129];      // candidate run of every x-row of the ball, exclusive prefix of the lengths
  int warp_tot[8];
  unsigned long long bar;
  unsigned long long cand;
  int count;
};

template <int CAP>
__global__ void __launch_bounds__(kThreads, CAP <= 2048 ? 4 : 1)
k_hand_sweep(const SweepArgs A, const __grid_constant__ HandConst hc) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  SweepShared& sh = *reinterpret_cast<SweepShared*>(s_raw);
  double2* slab_xy = reinterpret_cast<double2*>(s_raw + sizeof(SweepShared));  // hand-frame x, y of every slab point
  uint32_t* slab_tag = reinterpret_cast<uint32_t*>(slab_xy + CAP);             // point index << 2 | tag bits

  if (A.sample_count) {
    // inline large-slab pass: the list was written by the common kernel just before; what the grid cannot take is
    // handed on to the host-side pass (it joins the unresolved list)
    const int cnt = *A.sample_count;
    if (blockIdx.x == 0)
      for (int i = int(gridDim.x) + int(threadIdx.x); i < cnt; i += kThreads) A.overflow[1 + atomicAdd(A.overflow, 1)] = A.sample_list[i];
    if (int(blockIdx.x) >= cnt) return;
  }
  const int s = A.sample_list ? A.sample_list[blockIdx.x] : blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const RowIndex& ri = *A.ri;
  int idx;
  GPoint q;
  if (A.sample_q) {  // one load instead of the dependent pair indices -> cloud
    const float4 v = A.sample_q[s];
    const int w = __float_as_int(v.w);
    idx = (s < ri.n_samples && w >= 0) ? (w >> 1) : -1;
    q.x = v.x;
    q.y = v.y;
    q.z = v.z;
    q.tag = uint32_t(w) & kTagCamBit;
  } else {
    idx = (s < ri.n_samples) ? A.indices[s] : -1;
    if (idx >= 0 && idx < ri.n_points) q = A.pts[idx];
  }
  if (idx < 0 || idx >= ri.n_points) {  // unused sample slot (fewer voxels than requested samples)
    if (lane == 0) A.valid[size_t(s) * 8 + warp] = 0;
    if (threadIdx.x == 0 && A.slab_counts) A.slab_counts[s] = 0;
    return;
  }
  const int sample_cam = (q.tag & kTagCamBit) ? 1 : 0;  // hands_cam_source (hand_search.cpp:40-42, App. B#3)
  const uint32_t bar = smem_u32(&sh.bar);
  if (threadIdx.x == 0) {
    sh.count = 0;
    sh.cand = 0;
    mbar_init(bar, 1);
    fence_proxy_async_smem();
  }
  if (threadIdx.x < AG_SWEEP_LUT) sh.lut[threadIdx.x] = hc.lut[threadIdx.x];
  if (threadIdx.x < 12) sh.bite[threadIdx.x] = hc.bite[threadIdx.x];

  // ---- frame = [normal | normal x axis | axis]   (rotating_hand.cpp:25)
  const ag_frame fr = A.frames[s];
  double F[3][3];
  {
    const double* a = fr.normal;
    const double* b = fr.axis;
    const double nxa[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
#pragma unroll
    for (int r = 0; r < 3; r++) {
      F[r][0] = a[r];
      F[r][1] = nxa[r];
      F[r][2] = b[r];
    }
  }

  // ---- phase A: gather the r = 0.08 ball, keep the |z_hand| < hand_height slab -----------------
  // One candidate run per x-row the ball can touch (two loads from the column table, ag_common.cuh), one thread
  // per row.  A run is contiguous in the voxel list, so every row thread moves its run with ONE TMA bulk copy
  // (cp.async.bulk) into the CTA's staging buffer at the run's prefix offset; one mbarrier transaction count
  // tracks them all: the whole ball is in flight at once and no registers are held for it.  The 256 threads then
  // test the staged candidates (FLANN's binary32 distance, then the slab) and append the survivors — already
  // rotated into the hand frame — to the slab.
  int klo[2] = {0, 0}, rows[2] = {0, 0};
  for (int c = 0; c < 2; c++) {
    if (ri.count[c] == 0) continue;
    int k_hi;
    row_range(ri, c, q.x, A.rpad, klo[c], k_hi);
    rows[c] = max(0, k_hi - klo[c] + 1);
  }
  const int ncol = rows[0] + rows[1];
  const float fzx = float(F[0][2]), fzy = float(F[1][2]), fzz = float(F[2][2]), hh = float(hc.hand_height);
  unsigned n_ball = 0;
  unsigned long long n_cand = 0;
  uint32_t parity = 0;
  for (int cbase = 0; cbase < ncol; cbase += 128) {  // one batch for the shipped radii (55 rows per camera)
    const int nc = min(128, ncol - cbase);
    int j0 = 0, len = 0;
    if (threadIdx.x < nc) {
      const int col = cbase + threadIdx.x;
      const int c = col < rows[0] ? 0 : 1;
      const int k = klo[c] + (c == 0 ? col : col - rows[0]);
      int j1;
      row_run(ri, A.row_ptr, A.col_ptr, A.pts, c, k, q.x, q.y, A.rpad, j0, j1);
      len = max(0, j1 - j0);
    }
    // exclusive prefix of the run lengths over the (<= 128) row threads
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) sh.warp_tot[warp] = incl;
    __syncthreads();
    int pre = incl - len;
    for (int w = 0; w < warp && w < 4; w++) pre += sh.warp_tot[w];
    const int total = sh.warp_tot[0] + sh.warp_tot[1] + sh.warp_tot[2] + sh.warp_tot[3];
    if (threadIdx.x < nc) {
      sh.run_start[threadIdx.x] = j0;
      sh.run_pre[threadIdx.x] = pre;
    }
    if (threadIdx.x == 0) sh.run_pre[nc] = total;
    n_cand += threadIdx.x == 0 ? (unsigned long long)total : 0ull;
    for (int base = 0; base < total; base += kStage) {
      const int cnt = min(kStage, total - base);
      fence_proxy_async_smem();  // earlier generic-proxy accesses of the staging buffer precede the async writes
      __syncthreads();
      if (threadIdx.x == 0) mbar_expect_tx(bar, uint32_t(cnt) * 16u);
      __syncthreads();
      {  // this row's share of the pass: [pre, pre + len) clipped to [base, base + cnt)
        const int a = max(pre, base), b = min(pre + len, base + cnt);
        if (threadIdx.x < nc && b > a)
          bulk_g2s(smem_u32(&sh.u.stage[a - base]), A.pts + j0 + (a - pre), uint32_t(b - a) * 16u, bar);
      }
      if (warp == 0) mbar_wait(bar, parity);  // one warp polls the barrier, the others sleep at the CTA barrier
      __syncthreads();
      parity ^= 1u;
      // two candidates per thread and step (independent loads and tests), one reservation of slab places per warp
      for (int i0 = 0; i0 < cnt; i0 += 2 * kThreads) {
        bool keep[2] = {false, false};
        GPoint p[2];
        float cx[2], cy[2], cz[2];
        unsigned in_any = 0;
#pragma unroll
        for (int u = 0; u < 2; u++) {
          const int i = i0 + u * kThreads + threadIdx.x;
          bool inball = false;
          cx[u] = cy[u] = cz[u] = 0.f;
          p[u] = sh.u.stage[i < cnt ? i : 0];
          if (i < cnt && dist2_flann(q.x, q.y, q.z, p[u].x, p[u].y, p[u].z) < A.r2) {
            inball = true;
            // hand_search.cpp:157-158: subtraction in binary32, then cast
            cx[u] = __fsub_rn(p[u].x, q.x);
            cy[u] = __fsub_rn(p[u].y, q.y);
            cz[u] = __fsub_rn(p[u].z, q.z);
            // slab test -h < z_hand < h (rotating_hand.cpp:44): decided in binary32 when the point is clearly
            // inside or outside (the binary32 value is within 1e-7 of the binary64 one), exactly otherwise
            const float hzf = fmaf(fzx, cx[u], fmaf(fzy, cy[u], fzz * cz[u]));
            const float az = fabsf(hzf);
            if (az < hh - 1e-5f) keep[u] = true;
            else if (az <= hh + 1e-5f) {
              const double hz = (F[0][2] * double(cx[u]) + F[1][2] * double(cy[u])) + F[2][2] * double(cz[u]);
              keep[u] = hz > -1.0 * hc.hand_height && hz < hc.hand_height;
            }
          }
          in_any += __popc(__ballot_sync(0xffffffffu, inball));
        }
        n_ball += in_any;
        const unsigned m0 = __ballot_sync(0xffffffffu, keep[0]), m1 = __ballot_sync(0xffffffffu, keep[1]);
        if (m0 | m1) {
          int wbase = 0;
          if (lane == 0) wbase = atomicAdd(&sh.count, __popc(m0) + __popc(m1));
          wbase = __shfl_sync(0xffffffffu, wbase, 0);
#pragma unroll
          for (int u = 0; u < 2; u++) {
            const int pos = wbase + (u ? __popc(m0) + __popc(m1 & lt) : __popc(m0 & lt));
            if (keep[u] && pos < CAP) {
              const double px = double(cx[u]), py = double(cy[u]), pz = double(cz[u]);
              double2 h;  // frame^T * p (rotating_hand.cpp:26)
              h.x = (F[0][0] * px + F[1][0] * py) + F[2][0] * pz;
              h.y = (F[0][1] * px + F[1][1] * py) + F[2][1] * pz;
              slab_xy[pos] = h;
              uint32_t tag = p[u].tag & 3u;
              if (tag & kTagNormalBit) {  // the point index is only needed to fetch its normal: row of position base + i
                const int g = base + i0 + u * kThreads + int(threadIdx.x);
                int lo_r = 0, hi_r = nc - 1;
                while (lo_r < hi_r) {
                  const int mid = (lo_r + hi_r + 1) >> 1;
                  if (sh.run_pre[mid] <= g) lo_r = mid;
                  else hi_r = mid - 1;
                }
                tag |= uint32_t(sh.run_start[lo_r] + (g - sh.run_pre[lo_r])) << 2;
              }
              slab_tag[pos] = tag;
            }
          }
        }
      }
    }
    __syncthreads();
  }
  if (lane == 0) atomicAdd(&sh.cand, (unsigned long long)n_ball << 32);
  __syncthreads();
  const int k = sh.count;
  if (threadIdx.x == 0) {
    if (A.slab_counts) A.slab_counts[s] = k;
    atomicAdd(&A.counters[2], sh.cand >> 32);
    atomicAdd(&A.counters[3], n_cand);
    if (k > CAP) {  // does not fit: queue the sample for the large-capacity instantiation
      const int w = atomicAdd(A.overflow, 1);
      A.overflow[1 + w] = s;
    }
  }
  const int o = warp;
  const size_t slot = size_t(s) * 8 + o;
  if (k > CAP) {
    if (lane == 0) A.valid[slot] = 0;
    return;
  }

  // ---- phase B: orientation o -------------------------------------------------------------------
  const double cs = hc.cosv[o], sn = hc.sinv[o];
  const double msn = -1.0 * sn;  // rot = [cs -sn 0; sn cs 0; 0 0 1]   (rotating_hand.cpp:90)
  // T = frame * rot^T ; approach = T*y, binormal = T*x   (rotating_hand.cpp:96,104)
  double T[3][3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    T[r][0] = (F[r][0] * cs + F[r][1] * msn) + F[r][2] * 0.0;
    T[r][1] = (F[r][0] * sn + F[r][1] * cs) + F[r][2] * 0.0;
    T[r][2] = (F[r][0] * 0.0 + F[r][1] * 0.0) + F[r][2] * 1.0;
  }
  double approach[3], binormal[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    approach[r] = (T[r][0] * 0.0 + T[r][1] * 1.0) + T[r][2] * 0.0;
    binormal[r] = (T[r][0] * 1.0 + T[r][1] * 0.0) + T[r][2] * 0.0;
  }
  const double se[3] = {double(q.x), double(q.y), double(q.z)};
  const double camv0[3] = {hc.cam[0][0] - se[0], hc.cam[0][1] - se[1], hc.cam[0][2] - se[2]};
  const double camv1[3] = {hc.cam[1][0] - se[0], hc.cam[1][1] - se[1], hc.cam[1][2] - se[2]};
  int status = 0, e_idx = -1, last = 0;
  unsigned fingers_last = 0;
  bool have = false;
  double minY = 1e300, maxY = -1e300;
  uint32_t* img = reinterpret_cast<uint32_t*>(&sh.u.b.lvl[warp][0][0]);  // (after the level masks have been read)
  static_assert(sizeof(unsigned long long) * 12 * 16 >= sizeof(uint32_t) * AG_IMAGE_WORDS, "image fits the level masks");
  const bool cam_ok = !(dot3e(approach, camv0) > 0 && dot3e(approach, camv1) > 0);  // rotating_hand.cpp:99
  if (cam_ok) {
    // pass 1: slot masks per depth level.  Every boolean of FingerHand is "is there a point with y < d_t whose x
    // lies in some interval", so each point ORs its 20 + 20 slot bits into the first depth level that crops it
    // in.  Between two consecutive slot edges the bits are constant: they come from a 44-entry table indexed
    // by the zone of x (slot edges are two interleaved uniform grids, so the zone is a floor and a fraction
    // test); a point closer than kLutMargin steps to an edge — or a hand geometry without the table — takes the
    // reference's own comparisons slot by slot.  Same for the depth level of y.
    unsigned long long(*lvl)[16] = sh.u.b.lvl[warp];
    uint32_t* lvl32 = reinterpret_cast<uint32_t*>(&lvl[0][lane & 15]);  // level t: words [32 t] (in), [32 t + 1] (side)
    for (int i = lane; i < 12 * 16; i += 32) (&lvl[0][0])[i] = 0ull;
    __syncwarp();
    const double deepest = hc.bite[hc.n_depths - 1];
    const float phi = hc.lut_phi;
#pragma unroll 2
    for (int j = lane; j < k; j += 32) {
      const double2 h = slab_xy[j];
      const double rx = cs * h.x + msn * h.y;  // rot * p (rotating_hand.cpp:91); the 0*z term is exact
      const double ry = sn * h.x + cs * h.y;
      minY = fmin(minY, ry);
      maxY = fmax(maxY, ry);
      if (!(ry < deepest)) continue;  // above the deepest bite: never cropped in
      const double tx = (rx - hc.spacing[0]) * hc.inv_slot_step;
      const double ty = (ry - hc.bite[0]) * hc.inv_bite_step;
      const int kx = __double2int_rd(tx), ky = __double2int_rd(ty);
      const float fx = float(tx - double(kx)), fy = float(ty - double(ky));
      const bool zone_a = fx > kLutMargin && fx < phi - kLutMargin;
      const bool zone_b = fx > phi + kLutMargin && fx < 1.0f - kLutMargin;
      const bool far_x = kx < -1 || kx > AG_SWEEP_LUT / 2 - 2;
      const bool y_ok = fy > kLutMargin && fy < 1.0f - kLutMargin;
      unsigned long long m;
      int L;
      if (hc.lut_ok && (zone_a || zone_b || far_x) && y_ok) {
        const int kc = min(max(kx, -1), AG_SWEEP_LUT / 2 - 2) + 1;
        m = sh.lut[2 * kc + (zone_a && !far_x ? 0 : 1)];
        L = min(max(ky + 1, 0), hc.n_depths - 1);
      } else {
        m = slot_masks_exact(hc.spacing, hc.finger_width, rx);
        L = 0;
        for (int t = 0; t < hc.n_depths; t++) L += (sh.bite[t] <= ry) ? 1 : 0;
      }
      // the point is cropped in at every depth level d_t > ry, i.e. at levels t >= L = #{t : d_t <= ry}
      // (L < n_depths here); record it at level L only, the prefix-OR below spreads it upwards
      if (unsigned(m)) atomicOr(lvl32 + 32 * L, unsigned(m));
      if (unsigned(m >> 32)) atomicOr(lvl32 + 32 * L + 1, unsigned(m >> 32));
    }
    __syncwarp();
    unsigned IN[12], SD[12];
    {
      unsigned long long acc = 0ull;
#pragma unroll
      for (int t = 0; t < 12; t++) {
        acc |= lvl[t][lane & 15];
        IN[t] = unsigned(acc);
        SD[t] = unsigned(acc >> 32);
      }
    }
#pragma unroll
    for (int t = 0; t < 12; t++) {
      IN[t] = __reduce_or_sync(0xffffffffu, IN[t]);
      SD[t] = __reduce_or_sync(0xffffffffu, SD[t]);
    }
    minY = warp_min(minY);
    maxY = warp_max(maxY);
    // finger / hand logic at depth level t (finger_hand.cpp:20-115)
    auto hand_at = [&](int t, unsigned& fingers) -> unsigned {
      const bool abort = minY < hc.back[t];  // a cropped point behind the back of the hand (:37-40)
      fingers = abort ? 0u : ((~IN[t]) & SD[t] & 0xFFFFFu);
      return fingers & (fingers >> 10) & 0x3FFu;
    };
    status = 1;
    unsigned f0;
    const unsigned hand0 = k > 0 ? hand_at(0, f0) : 0u;
    if (hand0) {  // rotating_hand.cpp:111
      have = true;
      status = 2;
      const int len = __popc(hand0);
      e_idx = __fns(hand0, 0, (len + 1) / 2);  // idx[ceil(len/2) - 1]   (finger_hand.cpp:190)
      fingers_last = f0;
      for (int t = 1; t < hc.n_depths; t++) {  // deepenHand (finger_hand.cpp:204-228)
        unsigned ft;
        const unsigned ht = hand_at(t, ft);
        if (!((ht >> e_idx) & 1u)) break;
        last = t;
        fingers_last = ft;
      }
    }
  }
  if (A.debug && lane == 0)
    A.debug[slot] = status | ((e_idx & 0xF) << 4) | (last << 8) | int(fingers_last << 12);
  if (!have) {
    if (lane == 0) A.valid[slot] = 0;
    return;
  }
  __syncwarp();  // the level masks have been consumed: their bytes become this orientation's grasp image
  for (int i = lane; i < AG_IMAGE_WORDS; i += 32) img[i] = 0u;
  __syncwarp();

  // ---- grasp parameters (finger_hand.cpp:117-171, rotating_hand.cpp:118-154) --------------------
  const double hor = hc.half_od + hc.spacing[e_idx];
  double surface3[3], bottom3[3], surf_w[3], bot_w[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    surface3[r] = (T[r][0] * hor + T[r][1] * minY) + T[r][2] * 0.0;
    bottom3[r] = (T[r][0] * hor + T[r][1] * maxY) + T[r][2] * 0.0;
    surf_w[r] = surface3[r] + se[r];
    bot_w[r] = bottom3[r] + se[r];
  }
  const double lim = hc.lim[last];  // back_of_hand + hand_depth   (rotating_hand.cpp:128)
  // image orientation (learning.cpp:330-333, 382-383)
  const double s2c[3] = {surf_w[0] - hc.cam[sample_cam][0], surf_w[1] - hc.cam[sample_cam][1],
                         surf_w[2] - hc.cam[sample_cam][2]};
  const bool keep_sign = dot3e(binormal, s2c) > 0;
  const double left = hc.spacing[e_idx], right = hc.spacing[10 + e_idx];
  double wmin = 100000.0, wmax = -100000.0;
  int m_box = 0, numl = 0, numr = 0;
  // floor(v / img_cell) (learning.cpp:330-335): a product with the reciprocal, unless that lands within 1e-6 of
  // an integer — then the reference's own division decides
  auto cell_of = [&](double v) {
    const double qf = v * hc.inv_img_cell;
    const double fl = floor(qf);
    const double fr2 = qf - fl;
    return (fr2 > 1e-6 && fr2 < 1.0 - 1e-6) ? fl : floor(v / hc.img_cell);
  };
  for (int j = lane; j < k; j += 32) {
    const double2 h = slab_xy[j];
    const double rx = cs * h.x + msn * h.y;
    const double ry = sn * h.x + cs * h.y;
    if (ry < hc.bite[0] && rx > left && rx < right) {  // finger_hand.cpp:159-167
      wmin = fmin(wmin, rx);
      wmax = fmax(wmax, rx);
    }
    if (ry < lim) {  // points in the box (rotating_hand.cpp:126-130)
      m_box++;
      // learning.cpp:320-365 on points_for_learning = rotated point - surface (frame mix-up kept)
      const double bx = rx - surface3[0], by = ry - surface3[1];
      const double hcell = cell_of((keep_sign ? bx : -bx) - (-0.05));
      const double vcell = cell_of(by - 0.0);
      const int hpx = int(fmin(99.0, fmax(0.0, hcell)));
      const int vpx = int(fmin(79.0, fmax(0.0, vcell)));
      const int bit = (AG_IMAGE_ROWS - 1 - vpx) * AG_IMAGE_COLS + hpx;
      atomicOr(&img[bit >> 5], 1u << (bit & 31));
      const uint32_t tag = slab_tag[j];
      if (tag & kTagNormalBit) {  // antipodal.cpp:12-86 on rot * frame^T * normal
        const double* nv = A.normals + size_t(3) * (tag >> 2);
        const double n0 = nv[0], n1 = nv[1], n2 = nv[2];
        const double hn0 = (F[0][0] * n0 + F[1][0] * n1) + F[2][0] * n2;
        const double hn1 = (F[0][1] * n0 + F[1][1] * n1) + F[2][1] * n2;
        const double nrx = cs * hn0 + msn * hn1;
        numl += (-1.0 * nrx > hc.cos_thresh) ? 1 : 0;
        numr += (nrx > hc.cos_thresh) ? 1 : 0;
      }
    }
  }
  wmin = warp_min(wmin);
  wmax = warp_max(wmax);
  m_box = __reduce_add_sync(0xffffffffu, m_box);
  numl = __reduce_add_sync(0xffffffffu, numl);
  numr = __reduce_add_sync(0xffffffffu, numr);
  __syncwarp();
  bool keep_hyp = true;
  if (A.filter_boundaries) {  // localization.cpp:364-388
#pragma unroll
    for (int kk = 0; kk < 6; kk++)
      if (fabs(surf_w[kk / 2] - A.workspace[kk]) < 0.02) keep_hyp = false;
  }
  if (lane == 0) {
    ag_grasp gr;
#pragma unroll
    for (int r = 0; r < 3; r++) {
      gr.axis[r] = fr.axis[r];
      gr.approach[r] = approach[r];
      gr.binormal[r] = binormal[r];
      gr.bottom[r] = bot_w[r];
      gr.surface[r] = surf_w[r];
    }
    gr.width = wmax - wmin;
    gr.score = __int_as_float(0x7FC00000);
    gr.sample_index = idx;
    gr.sample_slot = s;
    gr.orientation = o;
    gr.cam_source = sample_cam;
    gr.num_points = m_box;
    gr.image_id = -1;
    gr.half_antipodal = (numl > 6 || numr > 6) ? 1 : 0;
    gr.full_antipodal = (numl > 6 && numr > 6) ? 1 : 0;
    gr.label = 0;
    gr.reserved = 0;
    A.grasps[slot] = gr;
    A.valid[slot] = keep_hyp ? 1 : 0;
  }
  uint32_t* gimg = A.images + slot * AG_IMAGE_WORDS;
  int occupied = 0;
  for (int i = lane; i < AG_IMAGE_WORDS; i += 32) {
    const uint32_t w = img[i];
    gimg[i] = w;
    occupied += __popc(w);
  }
  // The scorer's list (its order is free: scoring does not need the sample-major order) is filled from both ends:
  // images with many occupied pixels — the expensive ones for the HOG kernel, which lasts as long as its heaviest
  // image — from the front, so that their CTAs are dispatched first; the others from the back.
  occupied = __reduce_add_sync(0xffffffffu, occupied);
  if (lane == 0 && keep_hyp) {
    if (occupied >= A.heavy_image) A.hyp_list[atomicAdd(&A.hyp_count[0], 1)] = int(slot);
    else A.hyp_list[A.n_samples * 8 - 1 - atomicAdd(&A.hyp_count[1], 1)] = int(slot);
  }
}

// Stable compaction of the valid (sample, orientation) slots: slots[h] = raw slot of hypothesis h in
// (sample, orientation) order, *n_sel = number of hypotheses.  One CTA: the flag array is tiny
// (8 bytes per sample) and a single-block scan needs neither a second launch nor temporary storage.
__global__ void __launch_bounds__(1024)
k_compact_slots(const uint8_t* __restrict__ valid, int n_slots, int* __restrict__ slots, int* __restrict__ n_sel) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n_slots; base += 1024 * 8) {
    // each thread owns 8 consecutive flags (= one sample)
    const int first = base + tid * 8;
    unsigned bits = 0;
    if (first + 8 <= n_slots) {
      const uint2 v = *reinterpret_cast<const uint2*>(valid + first);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        bits |= ((v.x >> (8 * k)) & 0xFFu ? 1u : 0u) << k;
        bits |= ((v.y >> (8 * k)) & 0xFFu ? 1u : 0u) << (4 + k);
      }
    } else {
      for (int k = 0; k < 8 && first + k < n_slots; k++) bits |= (valid[first + k] ? 1u : 0u) << k;
    }
    const int cnt = __popc(bits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    int pos = s_carry + (warp > 0 ? s_warp[warp - 1] : 0) + incl - cnt;
    for (unsigned b2 = bits; b2; b2 &= b2 - 1) slots[pos++] = first + (__ffs(b2) - 1);
    __syncthreads();
    if (tid == 1023) s_carry = pos;  // thread 1023's end position = total so far
    __syncthreads();
  }
  if (tid == 0) *n_sel = s_carry;
}

// On-demand materialisation of GraspHypothesis::points_for_learning_ for ONE hypothesis (sample s,
// orientation o): the rotated slab points inside the hand box minus the surface vector
// (rotating_hand.cpp:125-151), with their camera source, FLANN distance and cloud index (the host sorts
// them into the reference's (distance, index) neighbour order).  Same arithmetic as k_hand_sweep.
__global__ void __launch_bounds__(256)
k_box_points(const SweepArgs A, const __grid_constant__ HandConst hc, int s, int o, double* out_pts, int* out_cam,
             float* out_d2, int* out_idx, int* out_count, int cap) {
  __shared__ double s_min[8];
  __shared__ int s_j0, s_j1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RowIndex ri = *A.ri;
  const int idx = A.indices[s];
  const GPoint q = A.pts[idx];
  const ag_frame fr = A.frames[s];
  double F[3][3];
  {
    const double* a = fr.normal;
    const double* b = fr.axis;
    const double nxa[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    for (int r = 0; r < 3; r++) {
      F[r][0] = a[r];
      F[r][1] = nxa[r];
      F[r][2] = b[r];
    }
  }
  const double cs = hc.cosv[o], sn = hc.sinv[o], msn = -1.0 * sn;
  const unsigned dbg = unsigned(A.debug[size_t(s) * 8 + o]);
  const int e_idx = int((dbg >> 4) & 0xFu), last = int((dbg >> 8) & 0xFu);
  double surface3[3] = {0, 0, 0};
  for (int pass = 0; pass < 2; pass++) {
    double minY = 1e300;
    for (int c = 0; c < 2; c++) {
      if (ri.count[c] == 0) continue;
      int k_lo, k_hi;
      row_range(ri, c, q.x, A.rpad, k_lo, k_hi);
      for (int k = k_lo; k <= k_hi; k++) {
        __syncthreads();
        if (threadIdx.x == 0) {
          int j0, j1;
          row_run(ri, A.row_ptr, A.col_ptr, A.pts, c, k, q.x, q.y, A.rpad, j0, j1);
          s_j0 = j0;
          s_j1 = j1;
        }
        __syncthreads();
        for (int j = s_j0 + int(threadIdx.x); j < s_j1; j += blockDim.x) {
          const GPoint p = A.pts[j];
          const float d2 = dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z);
          if (!(d2 < A.r2)) continue;
          const double px = double(__fsub_rn(p.x, q.x)), py = double(__fsub_rn(p.y, q.y)), pz = double(__fsub_rn(p.z, q.z));
          const double hz = (F[0][2] * px + F[1][2] * py) + F[2][2] * pz;
          if (!(hz > -1.0 * hc.hand_height && hz < hc.hand_height)) continue;
          const double hx = (F[0][0] * px + F[1][0] * py) + F[2][0] * pz;
          const double hy = (F[0][1] * px + F[1][1] * py) + F[2][1] * pz;
          const double rx = cs * hx + msn * hy, ry = sn * hx + cs * hy;
          if (pass == 0) {
            minY = fmin(minY, ry);
          } else if (ry < hc.lim[last]) {
            const int w = atomicAdd(out_count, 1);
            if (w < cap) {
              const double rz = (0.0 * hx + 0.0 * hy) + 1.0 * hz;
              out_pts[3 * size_t(w)] = rx - surface3[0];
              out_pts[3 * size_t(w) + 1] = ry - surface3[1];
              out_pts[3 * size_t(w) + 2] = rz - surface3[2];
              out_cam[w] = (p.tag & kTagCamBit) ? 1 : 0;
              out_d2[w] = d2;
              out_idx[w] = j;
            }
          }
        }
      }
    }
    if (pass == 0) {
      minY = warp_min(minY);
      if (lane == 0) s_min[warp] = minY;
      __syncthreads();
      double m = s_min[0];
      for (int w2 = 1; w2 < 8; w2++) m = fmin(m, s_min[w2]);
      const double hor = hc.half_od + hc.spacing[e_idx];
      for (int r = 0; r < 3; r++) {
        const double T0 = (F[r][0] * cs + F[r][1] * msn) + F[r][2] * 0.0;
        const double T1 = (F[r][0] * sn + F[r][1] * cs) + F[r][2] * 0.0;
        const double T2 = (F[r][0] * 0.0 + F[r][1] * 0.0) + F[r][2] * 1.0;
        surface3[r] = (T0 * hor + T1 * m) + T2 * 0.0;
      }
    }
  }
}

// Training instances of one hypothesis from the points of ONE camera (Learning::createInstance(h, cam_pos, cam),
// learning.cpp:375-400: the "simulated camera" instances of Learning::train, :76-141): the grasp image rasterised
// from the box points whose camera source is 0, and from those whose source is 1.  One CTA per hypothesis, the
// r = 0.08 ball gathered again with the arithmetic of k_hand_sweep; images[2 h], images[2 h + 1] must be zeroed.
__global__ void __launch_bounds__(256)
k_camera_images(const SweepArgs A, const __grid_constant__ HandConst hc, const int* __restrict__ slots, int n,
                uint32_t* __restrict__ images) {
  __shared__ double s_min[8];
  __shared__ int s_j0, s_j1;
  const int h = blockIdx.x;
  if (h >= n) return;
  const int slot = slots[h], s = slot >> 3, o = slot & 7;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const RowIndex& ri = *A.ri;
  const int idx = A.indices[s];
  const GPoint q = A.pts[idx];
  const int sample_cam = (q.tag & kTagCamBit) ? 1 : 0;
  const ag_frame fr = A.frames[s];
  double F[3][3];
  {
    const double* a = fr.normal;
    const double* b = fr.axis;
    const double nxa[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
    for (int r = 0; r < 3; r++) {
      F[r][0] = a[r];
      F[r][1] = nxa[r];
      F[r][2] = b[r];
    }
  }
  const double cs = hc.cosv[o], sn = hc.sinv[o], msn = -1.0 * sn;
  const unsigned dbg = unsigned(A.debug[size_t(s) * 8 + o]);
  const int e_idx = int((dbg >> 4) & 0xFu), last = int((dbg >> 8) & 0xFu);
  double T[3][3];
  for (int r = 0; r < 3; r++) {
    T[r][0] = (F[r][0] * cs + F[r][1] * msn) + F[r][2] * 0.0;
    T[r][1] = (F[r][0] * sn + F[r][1] * cs) + F[r][2] * 0.0;
    T[r][2] = (F[r][0] * 0.0 + F[r][1] * 0.0) + F[r][2] * 1.0;
  }
  double surface3[3] = {0, 0, 0};
  bool keep_sign = false;
  for (int pass = 0; pass < 2; pass++) {
    double minY = 1e300;
    for (int c = 0; c < 2; c++) {
      if (ri.count[c] == 0) continue;
      int k_lo, k_hi;
      row_range(ri, c, q.x, A.rpad, k_lo, k_hi);
      for (int k = k_lo; k <= k_hi; k++) {
        __syncthreads();
        if (threadIdx.x == 0) {
          int j0, j1;
          row_run(ri, A.row_ptr, A.col_ptr, A.pts, c, k, q.x, q.y, A.rpad, j0, j1);
          s_j0 = j0;
          s_j1 = j1;
        }
        __syncthreads();
        for (int j = s_j0 + int(threadIdx.x); j < s_j1; j += blockDim.x) {
          const GPoint p = A.pts[j];
          if (!(dist2_flann(q.x, q.y, q.z, p.x, p.y, p.z) < A.r2)) continue;
          const double px = double(__fsub_rn(p.x, q.x)), py = double(__fsub_rn(p.y, q.y)), pz = double(__fsub_rn(p.z, q.z));
          const double hz = (F[0][2] * px + F[1][2] * py) + F[2][2] * pz;
          if (!(hz > -1.0 * hc.hand_height && hz < hc.hand_height)) continue;
          const double hx = (F[0][0] * px + F[1][0] * py) + F[2][0] * pz;
          const double hy = (F[0][1] * px + F[1][1] * py) + F[2][1] * pz;
          const double rx = cs * hx + msn * hy, ry = sn * hx + cs * hy;
          if (pass == 0) {
            minY = fmin(minY, ry);
          } else if (ry < hc.lim[last]) {  // learning.cpp:320-365 on the box points of one camera
            const double bx = rx - surface3[0], by = ry - surface3[1];
            const double hcell = floor(((keep_sign ? bx : -bx) - (-0.05)) / hc.img_cell);
            const double vcell = floor((by - 0.0) / hc.img_cell);
            const int hpx = int(fmin(99.0, fmax(0.0, hcell)));
            const int vpx = int(fmin(79.0, fmax(0.0, vcell)));
            const int bit = (AG_IMAGE_ROWS - 1 - vpx) * AG_IMAGE_COLS + hpx;
            uint32_t* img = images + (size_t(2) * h + ((p.tag & kTagCamBit) ? 1 : 0)) * AG_IMAGE_WORDS;
            atomicOr(&img[bit >> 5], 1u << (bit & 31));
          }
        }
      }
    }
    if (pass == 0) {
      minY = warp_min(minY);
      if (lane == 0) s_min[warp] = minY;
      __syncthreads();
      double m = s_min[0];
      for (int w2 = 1; w2 < 8; w2++) m = fmin(m, s_min[w2]);
      const double hor = hc.half_od + hc.spacing[e_idx];
      double surf_w[3], binormal[3];
      for (int r = 0; r < 3; r++) {
        surface3[r] = (T[r][0] * hor + T[r][1] * m) + T[r][2] * 0.0;
        surf_w[r] = surface3[r] + double(r == 0 ? q.x : r == 1 ? q.y : q.z);
        binormal[r] = (T[r][0] * 1.0 + T[r][1] * 0.0) + T[r][2] * 0.0;
      }
      const double s2c[3] = {surf_w[0] - hc.cam[sample_cam][0], surf_w[1] - hc.cam[sample_cam][1],
                             surf_w[2] - hc.cam[sample_cam][2]};
      keep_sign = dot3e(binormal, s2c) > 0;  // learning.cpp:330-333,382-383
    }
  }
}

// caller-supplied cloud_normals_: flag the points whose normal is non-zero
__global__ void k_flag_normals(GPoint* pts, const double* __restrict__ normals, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool nz = normals && (normals[3 * size_t(i)] != 0.0 || normals[3 * size_t(i) + 1] != 0.0 ||
                              normals[3 * size_t(i) + 2] != 0.0);
  if (nz) atomicOr(&pts[i].tag, kTagNormalBit);
  else atomicAnd(&pts[i].tag, ~kTagNormalBit);
}

}  // namespace

int set_normals_device(Ctx* c, const double* h_normals) {
  const int n = c->n_vox;
  if (n <= 0) return AG_OK;
  if (c->normals.reserve(size_t(n) * 24)) return AG_ERR_CUDA;
  if (h_normals) AG_CUDA_CHECK(cudaMemcpyAsync(c->normals.p, h_normals, size_t(n) * 24, cudaMemcpyHostToDevice, c->stream));
  else AG_CUDA_CHECK(cudaMemsetAsync(c->normals.p, 0, size_t(n) * 24, c->stream));
  k_flag_normals<<<(n + 255) / 256, 256, 0, c->stream>>>(c->vox.as<GPoint>(),
                                                        h_normals ? c->normals.as<double>() : nullptr, n);
  AG_CUDA_CHECK(cudaGetLastError());
  AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return AG_OK;
}

void compute_hand_const(const ag_params& p, HandConst& h) {
  // finger_hand.cpp:8-15 — fs_half = LinSpaced(10, 0, od - fw) evaluated as low + i*step (Eigen 3.2)
  const double hi = p.hand_outer_diameter - p.finger_width;
  const double step = (hi - 0.0) / double(10 - 1);
  for (int i = 0; i < 10; i++) {
    const double v = 0.0 + double(i) * step;
    h.spacing[i] = (v - p.hand_outer_diameter) + p.finger_width;
    h.spacing[10 + i] = v;
  }
  h.finger_width = p.finger_width;
  h.outer_diameter = p.hand_outer_diameter;
  h.depth = p.hand_depth;
  h.hand_height = p.hand_height;
  h.init_bite = p.init_bite;
  // rotating_hand.cpp:12-15 — first 8 of LinSpaced(9, -pi, pi); :90 cos/sin through libm
  const double lo = -1.0 * M_PI, step_a = (M_PI - lo) / double(9 - 1);
  for (int i = 0; i < 8; i++) {
    const double a = lo + double(i) * step_a;
    h.cosv[i] = cos(a);
    h.sinv[i] = sin(a);
  }
  // finger_hand.cpp:199-204 — d = init + 0.005; d <= depth; d += 0.005 (binary64 accumulation)
  h.n_depths = 0;
  h.bite[h.n_depths++] = p.init_bite;
  for (double d = p.init_bite + 0.005; d <= p.hand_depth && h.n_depths < 12; d += 0.005) h.bite[h.n_depths++] = d;
  for (int t = 0; t < 12; t++) {
    if (t >= h.n_depths) h.bite[t] = h.bite[h.n_depths - 1];
    h.back[t] = -1.0 * (p.hand_depth - h.bite[t]);  // finger_hand.cpp:22
    h.lim[t] = h.back[t] + p.hand_depth;            // rotating_hand.cpp:128
  }
  h.cos_thresh = cos(20.0 * M_PI / 180.0);  // antipodal.cpp:15 with thresh 20 (rotating_hand.cpp:162)
  for (int a = 0; a < 3; a++) {
    h.cam[0][a] = p.cam_tf_left[4 * a + 3];
    h.cam[1][a] = p.cam_tf_right[4 * a + 3];
  }
  // slot-mask table: valid when the 40 slot edges are where the uniform model puts them (lower edges at integer
  // multiples of the step from spacing[0], upper edges a fraction phi further) — checked by evaluating the exact
  // predicates at three probes of every zone; any disagreement (odd geometry) disables the table
  h.inv_slot_step = step > 0 ? 1.0 / step : 0.0;
  h.inv_bite_step = 1.0 / 0.005;
  h.lut_ok = 0;
  h.lut_phi = 0.5f;
  for (int z = 0; z < AG_SWEEP_LUT; z++) h.lut[z] = 0ull;
  if (step > 0 && p.finger_width > 0) {
    const double w = p.finger_width / step;
    const double phi = w - floor(w);
    const int k_last = 18 + int(floor(w)) + 1;  // zone index of the last upper edge, plus one
    bool ok = phi > 0.01 && phi < 0.99 && k_last <= AG_SWEEP_LUT / 2 - 2;
    for (int k = -1; k <= AG_SWEEP_LUT / 2 - 2 && ok; k++)
      for (int zb = 0; zb < 2 && ok; zb++) {
        const double f_lo = zb ? phi : 0.0, f_hi = zb ? 1.0 : phi;
        const double probes[3] = {f_lo + 0.5e-4, 0.5 * (f_lo + f_hi), f_hi - 0.5e-4};
        unsigned long long m[3];
        for (int t = 0; t < 3; t++) m[t] = slot_masks_exact(h.spacing, p.finger_width, h.spacing[0] + (double(k) + probes[t]) * step);
        ok = m[0] == m[1] && m[1] == m[2];
        h.lut[2 * (k + 1) + zb] = m[1];
      }
    // beyond the table on either side nothing changes any more
    ok = ok && slot_masks_exact(h.spacing, p.finger_width, h.spacing[0] - 1e3) == h.lut[0] &&
         slot_masks_exact(h.spacing, p.finger_width, h.spacing[0] + 1e3) == h.lut[AG_SWEEP_LUT - 1];
    // the depth levels must sit at integer multiples of 0.005 from the first one (to 1e-9)
    for (int t = 0; t < 12 && ok; t++) {
      const double d = p.init_bite + 0.005 * double(t);
      // (h.bite is filled below; recompute the accumulation here)
      double acc = p.init_bite;
      for (int u = 0; u < t; u++) acc += 0.005;
      ok = fabs(acc - d) < 1e-12;
    }
    if (const char* e = getenv("AG_SWEEP_FAST")) ok = ok && atoi(e) != 0;  // diagnostics: force the exact comparisons
    h.lut_ok = ok ? 1 : 0;
    h.lut_phi = float(phi);
  }
  h.img_cell = (0.05 - (-0.05)) / double(AG_IMAGE_COLS);  // learning.cpp:322-324
  h.inv_img_cell = 1.0 / h.img_cell;
  h.half_od = p.hand_outer_diameter / 2.0;
}

static SweepArgs make_args(Ctx* c, const int* d_indices, int n, const ag_frame* d_frames, unsigned flags) {
  SweepArgs A;
  A.pts = c->vox.as<GPoint>();
  A.row_ptr = c->row_ptr.as<int>();
  A.col_ptr = c->col_ptr.as<int>();
  A.ri = c->row_index.as<RowIndex>();
  A.indices = d_indices;
  A.frames = d_frames;
  A.normals = c->normals.as<double>();
  A.grasps = c->grasps_raw.as<ag_grasp>();
  A.valid = c->valid.as<uint8_t>();
  A.images = c->images_raw.as<uint32_t>();
  A.slab_counts = c->sweep_dbg.as<int>();
  A.debug = c->sweep_dbg.as<int>() + n;
  A.counters = c->counters.as<unsigned long long>();
  A.overflow = c->overflow.as<int>();
  A.hyp_list = c->hyp_list.as<int>();
  A.hyp_count = reinterpret_cast<int*>(c->counters.as<unsigned long long>() + 4);
  static const int heavy = [] {  // AG_HEAVY_IMAGE (measurements)
    const char* e = getenv("AG_HEAVY_IMAGE");
    return e ? atoi(e) : kHeavyImage;
  }();
  A.heavy_image = heavy;
  A.sample_list = nullptr;
  A.sample_count = nullptr;
  A.sample_q = nullptr;
  A.n_samples = n;
  const double radius = c->params.nn_radius_hands;
  A.r2 = float(radius * radius);
  A.rpad = sqrt(double(A.r2)) * (1.0 + 1e-5) + 1e-7;
  A.filter_boundaries = (flags & 0x100u) ? 1 : 0;
  for (int i = 0; i < 6; i++) A.workspace[i] = c->params.workspace[i];
  return A;
}

static int compact(Ctx* c, const SweepArgs& A, size_t slots, cudaStream_t st) {
  int* d_slots = c->hyp_slots.as<int>();
  int* d_nsel = d_slots + slots;
  k_compact_slots<<<1, 1024, 0, st>>>(A.valid, int(slots), d_slots, d_nsel);
  AG_CUDA_CHECK(cudaGetLastError());
  return AG_OK;
}

// points_for_learning of hypothesis `slot` = sample*8 + orientation of the last sweep
int box_points_device(Ctx* c, int n_samples, int slot, std::vector<double>& pts, std::vector<int>& cam) {
  pts.clear();
  cam.clear();
  const int cap = std::max(c->n_vox, 1);
  DevBuf buf;
  if (buf.reserve(size_t(cap) * (24 + 4 + 4 + 4) + 16)) return AG_ERR_CUDA;
  double* d_pts = buf.as<double>();
  int* d_cam = reinterpret_cast<int*>(d_pts + size_t(cap) * 3);
  float* d_d2 = reinterpret_cast<float*>(d_cam + cap);
  int* d_idx = reinterpret_cast<int*>(d_d2 + cap);
  int* d_cnt = d_idx + cap;
  AG_CUDA_CHECK(cudaMemsetAsync(d_cnt, 0, 4, c->stream));
  SweepArgs A = make_args(c, c->sweep_indices, n_samples, c->sweep_frames, c->sweep_flags);
  k_box_points<<<1, 256, 0, c->stream>>>(A, c->hand, slot / 8, slot % 8, d_pts, d_cam, d_d2, d_idx, d_cnt, cap);
  int m = 0;
  AG_CUDA_CHECK(cudaMemcpyAsync(&m, d_cnt, 4, cudaMemcpyDeviceToHost, c->stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  m = std::min(m, cap);
  std::vector<double> hp(size_t(m) * 3);
  std::vector<int> hc2(m), hi(m);
  std::vector<float> hd(m);
  if (m > 0) {
    AG_CUDA_CHECK(cudaMemcpy(hp.data(), d_pts, size_t(m) * 24, cudaMemcpyDeviceToHost));
    AG_CUDA_CHECK(cudaMemcpy(hc2.data(), d_cam, size_t(m) * 4, cudaMemcpyDeviceToHost));
    AG_CUDA_CHECK(cudaMemcpy(hd.data(), d_d2, size_t(m) * 4, cudaMemcpyDeviceToHost));
    AG_CUDA_CHECK(cudaMemcpy(hi.data(), d_idx, size_t(m) * 4, cudaMemcpyDeviceToHost));
  }
  buf.release();
  // the reference's column order is the kd-tree's neighbour order: ascending (distance, index)
  std::vector<int> order(m);
  for (int i = 0; i < m; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return hd[a] != hd[b] ? hd[a] < hd[b] : hi[a] < hi[b]; });
  pts.resize(size_t(m) * 3);
  cam.resize(m);
  for (int i = 0; i < m; i++) {
    for (int d = 0; d < 3; d++) pts[size_t(i) * 3 + d] = hp[size_t(order[i]) * 3 + d];
    cam[i] = hc2[order[i]];
  }
  return AG_OK;
}

// per-camera grasp images of n hypotheses (raw slots in d_slots) of the last sweep: [n][2][AG_IMAGE_WORDS], zeroed here
int camera_images_device(Ctx* c, int n_samples, const int* d_slots, int n, uint32_t* d_images) {
  if (n <= 0) return AG_OK;
  AG_CUDA_CHECK(cudaMemsetAsync(d_images, 0, size_t(n) * 2 * AG_IMAGE_WORDS * 4, c->stream));
  SweepArgs A = make_args(c, c->sweep_indices, n_samples, c->sweep_frames, c->sweep_flags);
  k_camera_images<<<n, 256, 0, c->stream>>>(A, c->hand, d_slots, n, d_images);
  c->launches += 1;
  AG_CUDA_CHECK(cudaGetLastError());
  return AG_OK;
}

int* hand_sweep_count_ptr(Ctx* c, int n) { return c->hyp_slots.as<int>() + size_t(n) * 8; }
int* hand_sweep_overflow_ptr(Ctx* c) { return c->overflow.as<int>(); }

int* hand_sweep_list_ptr(Ctx* c) { return c->hyp_list.as<int>(); }
int* hand_sweep_list_count_ptr(Ctx* c) { return reinterpret_cast<int*>(c->counters.as<unsigned long long>() + 4); }

int hand_sweep_enqueue(Ctx* c, const int* d_indices, int n, const ag_frame* d_frames, unsigned flags, bool fork_compact,
                       bool frames_from_fit, bool inline_big) {
  c->n_hyp = 0;
  c->images_valid = false;
  if (n <= 0) return AG_OK;
  const size_t slots = size_t(n) * 8;
  if (c->grasps_raw.reserve(slots * sizeof(ag_grasp)) || c->valid.reserve(slots + 64) ||
      c->images_raw.reserve(slots * AG_IMAGE_WORDS * 4 + 64) || c->hyp_slots.reserve(slots * 4 + 16) ||
      c->grasps.reserve(slots * sizeof(ag_grasp)) || c->counters.reserve(64) ||
      c->sweep_dbg.reserve(size_t(n) * 4 + slots * 4) || c->overflow.reserve(size_t(n + 1) * 4) ||
      c->hyp_list.reserve(slots * 4 + 16))
    return AG_ERR_CUDA;
  SweepArgs A = make_args(c, d_indices, n, d_frames, flags);
  c->sweep_from_fit = frames_from_fit && c->sample_q.p != nullptr;
  if (c->sweep_from_fit) A.sample_q = c->sample_q.as<float4>();
  const size_t smem_small = sizeof(SweepShared) + size_t(kSlabCapSmall) * 20;
  const size_t smem_big = sizeof(SweepShared) + size_t(kSlabCapBig) * 20;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_hand_sweep<kSlabCapSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_small));
    cudaFuncSetAttribute(k_hand_sweep<kSlabCapSmall>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(k_hand_sweep<kSlabCapBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_big));
    attr_set = true;
  }
  if (c->fold_resets & 2u) c->fold_resets &= ~2u;  // (ag_localize: zeroed by k_init_state)
  else AG_CUDA_CHECK(cudaMemsetAsync(A.overflow, 0, 4, c->stream));
  if (inline_big) {
    // Dense clouds (the fused 7-view cloud of config 5: ~0.5 % of the samples) overflow the common kernel's slab on
    // every call: the large-slab instantiation then runs right behind it on the samples the common kernel listed —
    // count and list stay on the device, no host round trip, one compaction / scoring / export for both.  What is
    // left in c->overflow afterwards is unresolved (more samples than this grid, or a slab beyond 9600 points).
    if (c->overflow_list.reserve(size_t(n + 1) * 4)) return AG_ERR_CUDA;
    if (c->fold_resets & 8u) c->fold_resets &= ~8u;
    else AG_CUDA_CHECK(cudaMemsetAsync(c->overflow_list.p, 0, 4, c->stream));
    SweepArgs B = A;
    A.overflow = c->overflow_list.as<int>();
    k_hand_sweep<kSlabCapSmall><<<n, kThreads, smem_small, c->stream>>>(A, c->hand);
    B.sample_list = c->overflow_list.as<int>() + 1;
    B.sample_count = c->overflow_list.as<int>();
    k_hand_sweep<kSlabCapBig><<<std::min(n, 2 * kNumSMs), kThreads, smem_big, c->stream>>>(B, c->hand);
    c->launches += 1;
  } else {
    k_hand_sweep<kSlabCapSmall><<<n, kThreads, smem_small, c->stream>>>(A, c->hand);
  }
  c->launches += 2;  // + k_compact_grasps
  c->sweep_flags = flags;
  c->sweep_indices = d_indices;
  c->sweep_frames = d_frames;
  if (fork_compact) {
    // the stable compaction (sample-major order for the export) runs on the side stream while the main stream
    // scores the hypotheses from the sweep's unordered list; the caller joins on ev_join before the export
    AG_CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
    AG_CUDA_CHECK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    int rc = compact(c, A, slots, c->stream2);
    AG_CUDA_CHECK(cudaEventRecord(c->ev_join, c->stream2));
    return rc;
  }
  return compact(c, A, slots, c->stream);
}

// Samples whose slab did not fit the 1920-point instantiation (dense neighbourhoods) are redone with the 9600-point
// one: enqueue only.  n_over = the overflow count the host has read; fresh_list: the scorer's list restarts at zero, so
// that it holds exactly the hypotheses of the redone samples (the others have been scored already).  The overflow
// counter is zeroed first: after this pass it counts the samples that do not fit the large instantiation either.
int hand_sweep_rerun_enqueue(Ctx* c, int n, int n_over, bool fresh_list) {
  const size_t slots = size_t(n) * 8;
  SweepArgs A = make_args(c, c->sweep_indices, n, c->sweep_frames, c->sweep_flags);
  if (c->overflow_list.reserve(size_t(n_over) * 4)) return AG_ERR_CUDA;
  AG_CUDA_CHECK(cudaMemcpyAsync(c->overflow_list.p, A.overflow + 1, size_t(n_over) * 4, cudaMemcpyDeviceToDevice, c->stream));
  AG_CUDA_CHECK(cudaMemsetAsync(A.overflow, 0, 4, c->stream));
  if (fresh_list) AG_CUDA_CHECK(cudaMemsetAsync(A.hyp_count, 0, 8, c->stream));
  A.sample_list = c->overflow_list.as<int>();
  const size_t smem_big = sizeof(SweepShared) + size_t(kSlabCapBig) * 20;
  k_hand_sweep<kSlabCapBig><<<n_over, kThreads, smem_big, c->stream>>>(A, c->hand);
  c->launches += 2;
  return compact(c, A, slots, c->stream);
}

// Called after the stream has been synchronised and the overflow counter read (stage call ag_hand_sweep): redo the
// overflow samples, wait, and read the hypothesis count.
int hand_sweep_finish(Ctx* c, int n, int n_over, int* n_hyp) {
  const size_t slots = size_t(n) * 8;
  int* d_nsel = c->hyp_slots.as<int>() + slots;
  if (n_over > 0) {
    int rc = hand_sweep_rerun_enqueue(c, n, n_over, false);
    if (rc) return rc;
    int over2 = 0;
    AG_CUDA_CHECK(cudaMemcpyAsync(&over2, hand_sweep_overflow_ptr(c), 4, cudaMemcpyDeviceToHost, c->stream));
    AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (over2) {
      set_error("hand sweep: a sample's slab exceeded the shared-memory capacity (9600 points)");
      return AG_ERR_CAPACITY;
    }
  }
  int h = 0;
  AG_CUDA_CHECK(cudaMemcpyAsync(&h, d_nsel, 4, cudaMemcpyDeviceToHost, c->stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->n_hyp = h;
  c->images_valid = true;
  if (n_hyp) *n_hyp = h;
  return AG_OK;
}

}  // namespace ag
