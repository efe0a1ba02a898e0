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
constexpr int kSlabCapSmall = 2048; // slab points kept in shared memory by the common kernel (32 KB)
constexpr int kSlabCapBig = 12032;  // fallback instantiation for dense neighbourhoods (188 KB, 1 CTA/SM)

struct SlabPoint {  // centred neighbour, binary32 exactly as hand_search.cpp:157-158 produces it
  float x, y, z;
  uint32_t tag;
};

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
  int* overflow;          // [0] = number of samples whose slab exceeded the capacity, [1..] their slots
  const int* sample_list; // if non-null: blockIdx.x indexes this list of sample slots (fallback pass)
  int n_samples;
  float r2;
  double rpad;
  int filter_boundaries;
  double workspace[6];
};

__device__ __forceinline__ double dot3e(const double a[3], const double b[3]) {
  return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]);  // Eigen's unrolled 3-vector reduction order
}

// Exact rank of x in an ascending, (nearly) uniformly spaced threshold table, branch-free.
// tab is stored with sentinels: tab[0] = -inf, tab[1..n] = thresholds, tab[n+1] = +inf.  The estimate
// c0 = clamp(floor((x-base)/step) + 1, 0, n) can be off by one only when x is within rounding distance
// of a threshold, so the true count is (c0-1) + [tab[c0] <cmp> x] + [tab[c0+1] <cmp> x]: two exact
// binary64 comparisons — the very comparisons the reference evaluates slot by slot.
__device__ __forceinline__ int rank_est(double x, double base, double inv_step, int n) {
  return min(max(__double2int_rd((x - base) * inv_step) + 1, 0), n);
}
struct SlotTables {    // shared-memory threshold tables with -inf / +inf sentinels (lane-varying indices)
  double sp[2][12];    // slot lower edges per hand (ascending)
  double spw[2][12];   // slot upper edges sp[i] + finger_width
  double bite[14];     // deepening levels d_t (ascending), n_depths <= 12
};

template <int CAP>
__global__ void __launch_bounds__(kThreads, CAP <= 2048 ? 3 : 1)
k_hand_sweep(const SweepArgs A, const __grid_constant__ HandConst hc) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  SlabPoint* slab = reinterpret_cast<SlabPoint*>(s_raw);
  __shared__ uint32_t s_img[8][AG_IMAGE_WORDS];
  __shared__ int s_count;
  __shared__ unsigned long long s_cand;
  __shared__ SlotTables s_tab;
  __shared__ int s_rs[kThreads], s_pre[kThreads + 1], s_next;
  __shared__ unsigned long long s_lvl[8][12][32];  // per warp, depth level and lane: (side<<32 | in) slot masks

  const int s = A.sample_list ? A.sample_list[blockIdx.x] : blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  const RowIndex ri = *A.ri;
  const int idx = (s < ri.n_samples) ? A.indices[s] : -1;
  if (idx < 0 || idx >= ri.n_points) {  // unused sample slot (fewer voxels than requested samples)
    if (lane == 0) A.valid[size_t(s) * 8 + warp] = 0;
    if (threadIdx.x == 0 && A.slab_counts) A.slab_counts[s] = 0;
    return;
  }
  const GPoint q = A.pts[idx];
  const int sample_cam = (q.tag & kTagCamBit) ? 1 : 0;  // hands_cam_source (hand_search.cpp:40-42, App. B#3)
  if (threadIdx.x == 0) {
    s_count = 0;
    s_cand = 0;
  }
  for (int i = lane; i < AG_IMAGE_WORDS; i += 32) s_img[warp][i] = 0u;
  if (threadIdx.x < 24) {
    const int hnd = threadIdx.x / 12, j = threadIdx.x % 12;
    const double inf = __longlong_as_double(0x7FF0000000000000ll);
    const double lo = j == 0 ? -inf : (j == 11 ? inf : hc.spacing[hnd * 10 + j - 1]);
    s_tab.sp[hnd][j] = lo;
    s_tab.spw[hnd][j] = (j == 0 || j == 11) ? lo : lo + hc.finger_width;  // finger_hand.cpp:56-57
  }
  if (threadIdx.x < 14) {
    const double inf = __longlong_as_double(0x7FF0000000000000ll);
    const int j = threadIdx.x;
    s_tab.bite[j] = j == 0 ? -inf : (j <= hc.n_depths ? hc.bite[j - 1] : inf);
  }

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
  __syncthreads();

  // ---- phase A: gather the r = 0.08 ball, keep the |z_hand| < hand_height slab -----------------
  // One candidate run per x-row the ball can touch (binary search on y, ag_common.cuh), one thread per
  // row; the warps then pull runs from a shared counter and stream them, two independent 16-byte loads
  // in flight per lane.
  {
    int klo[2] = {0, 0}, rows[2] = {0, 0};
    for (int c = 0; c < 2; c++) {
      if (ri.count[c] == 0) continue;
      int k_hi;
      row_range(ri, c, q.x, A.rpad, klo[c], k_hi);
      rows[c] = max(0, k_hi - klo[c] + 1);
    }
    const int ncol = rows[0] + rows[1];
    unsigned long long nball = 0, ncand = 0;
    for (int cbase = 0; cbase < ncol; cbase += kThreads) {  // one batch for the shipped radii
      const int nc = min(kThreads, ncol - cbase);
      if (threadIdx.x < nc) {
        const int col = cbase + threadIdx.x;
        const int c = col < rows[0] ? 0 : 1;
        const int k = klo[c] + (c == 0 ? col : col - rows[0]);
        int j0, j1;
        row_run(ri, A.row_ptr, A.col_ptr, A.pts, c, k, q.x, q.y, A.rpad, j0, j1);
        s_rs[threadIdx.x] = j0;
        s_pre[threadIdx.x] = j1;
      }
      if (threadIdx.x == 0) s_next = 0;
      __syncthreads();
      for (;;) {
        int run = 0;
        if (lane == 0) run = atomicAdd(&s_next, 1);
        run = __shfl_sync(0xffffffffu, run, 0);
        if (run >= nc) break;
        const int j0 = s_rs[run], j1 = s_pre[run];
        ncand += (unsigned long long)(j1 - j0);
        for (int jb = j0; jb < j1; jb += 64) {
          GPoint p[2];
          bool valid[2];
#pragma unroll
          for (int u = 0; u < 2; u++) {
            const int j = jb + u * 32 + lane;
            valid[u] = j < j1;
            p[u].x = p[u].y = p[u].z = 0.f;
            p[u].tag = 0;
            if (valid[u]) {
              p[u] = A.pts[j];
              p[u].tag = (uint32_t(j) << 2) | (p[u].tag & 3u);  // keep the point index for the normal fetch
            }
          }
#pragma unroll
          for (int u = 0; u < 2; u++) {
            if (u == 1 && jb + 32 >= j1) break;  // uniform
            bool keep = false, inball = false;
            SlabPoint sp;
            sp.x = sp.y = sp.z = 0.f;
            sp.tag = 0;
            if (valid[u] && dist2_flann(q.x, q.y, q.z, p[u].x, p[u].y, p[u].z) < A.r2) {
              inball = true;
              // hand_search.cpp:157-158: subtraction in binary32, then cast
              sp.x = __fsub_rn(p[u].x, q.x);
              sp.y = __fsub_rn(p[u].y, q.y);
              sp.z = __fsub_rn(p[u].z, q.z);
              sp.tag = p[u].tag;
              const double hz = (F[0][2] * double(sp.x) + F[1][2] * double(sp.y)) + F[2][2] * double(sp.z);
              keep = hz > -1.0 * hc.hand_height && hz < hc.hand_height;  // rotating_hand.cpp:44
            }
            nball += __popc(__ballot_sync(0xffffffffu, inball));
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (m) {
              int wbase = 0;
              if (lane == 0) wbase = atomicAdd(&s_count, __popc(m));
              wbase = __shfl_sync(0xffffffffu, wbase, 0);
              const int pos = wbase + __popc(m & lt);
              if (keep && pos < CAP) slab[pos] = sp;
            }
          }
        }
      }
      __syncthreads();
    }
    if (lane == 0) atomicAdd(&s_cand, ncand | (nball << 32));
  }
  __syncthreads();
  const int k = s_count;
  if (threadIdx.x == 0) {
    if (A.slab_counts) A.slab_counts[s] = k;
    atomicAdd(&A.counters[2], s_cand >> 32);
    atomicAdd(&A.counters[3], s_cand & 0xFFFFFFFFull);
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
  const bool cam_ok = !(dot3e(approach, camv0) > 0 && dot3e(approach, camv1) > 0);  // rotating_hand.cpp:99
  if (cam_ok) {
    // pass 1: slot masks per depth level
    unsigned IN[12], SD[12];
#pragma unroll
    for (int t = 0; t < 12; t++) s_lvl[warp][t][lane] = 0ull;
#pragma unroll 2
    for (int j = lane; j < k; j += 32) {
      const SlabPoint p = slab[j];
      const double px = double(p.x), py = double(p.y), pz = double(p.z);
      const double hx = (F[0][0] * px + F[1][0] * py) + F[2][0] * pz;  // frame^T * p (rotating_hand.cpp:26)
      const double hy = (F[0][1] * px + F[1][1] * py) + F[2][1] * pz;
      const double rx = cs * hx + msn * hy;  // rot * p (rotating_hand.cpp:91); the 0*z term is exact
      const double ry = sn * hx + cs * hy;
      minY = fmin(minY, ry);
      maxY = fmax(maxY, ry);
      if (!(ry < hc.bite[hc.n_depths - 1])) continue;  // above the deepest bite: never cropped in
      unsigned in_mask, sd_mask;
      if (hc.uniform_slots) {
        // slot i is entered iff sp[i] < rx < spw[i]; both tables ascend within a hand, so the sets
        // {i: rx > sp[i]}, {i: rx >= spw[i]}, {i: rx > spw[i]} are prefixes given by three exact ranks
        unsigned gt[2], ge[2], gw[2];
        int a_rank[2];
#pragma unroll
        for (int hnd = 0; hnd < 2; hnd++) {
          const double* sp = s_tab.sp[hnd];
          const double* sw = s_tab.spw[hnd];
          const int e0 = rank_est(rx, hc.spacing[hnd * 10], hc.inv_slot_step, 10);
          const int e1 = rank_est(rx, hc.spacing[hnd * 10] + hc.finger_width, hc.inv_slot_step, 10);
          const double s0 = sp[e0], s1 = sp[e0 + 1], w0 = sw[e1], w1 = sw[e1 + 1];
          const int a = e0 - 1 + (s0 < rx ? 1 : 0) + (s1 < rx ? 1 : 0);    // #{sp  <  rx}
          const int bq = e1 - 1 + (w0 <= rx ? 1 : 0) + (w1 <= rx ? 1 : 0); // #{spw <= rx}
          const int cq = e1 - 1 + (w0 < rx ? 1 : 0) + (w1 < rx ? 1 : 0);   // #{spw <  rx}
          a_rank[hnd] = a;
          gt[hnd] = (1u << a) - 1u;
          ge[hnd] = (1u << bq) - 1u;
          gw[hnd] = (1u << cq) - 1u;
        }
        in_mask = (gt[0] & ~ge[0]) | ((gt[1] & ~ge[1]) << 10);
        // finger_hand.cpp:72-82: slots 0..10 need a point beyond their upper edge, 11..19 one below their lower edge
        sd_mask = gw[0] | ((gw[1] & 1u) << 10) | (((~gt[1]) & 0x3FEu) << 10);
        // rx < sp[i] is NOT(rx > sp[i]) AND NOT(rx == sp[i]); correct the equality case exactly
        const int aR = a_rank[1];
        if (aR >= 1 && aR < 10 && rx == s_tab.sp[1][aR + 1]) sd_mask &= ~(1u << (10 + aR));
      } else {
        in_mask = 0u;
        sd_mask = 0u;
#pragma unroll
        for (int i = 0; i < 20; i++) {
          const double lo = hc.spacing[i], hi = hc.spacing[i] + hc.finger_width;  // finger_hand.cpp:56-57
          if (rx > lo && rx < hi) in_mask |= 1u << i;
          const bool side = (i <= 10) ? (rx > hi) : (rx < lo);  // finger_hand.cpp:72-82
          if (side) sd_mask |= 1u << i;
        }
      }
      // the point is cropped in at every depth level d_t > ry, i.e. at levels t >= L = #{t : d_t <= ry}
      // (L < n_depths here); record it at level L only, the prefix-OR below spreads it upwards
      const int e = rank_est(ry, hc.bite[0], 200.0, hc.n_depths);
      const int L = e - 1 + (s_tab.bite[e] <= ry ? 1 : 0) + (s_tab.bite[e + 1] <= ry ? 1 : 0);
      s_lvl[warp][L][lane] |= (static_cast<unsigned long long>(sd_mask) << 32) | in_mask;
    }
    {
      unsigned long long acc = 0ull;
#pragma unroll
      for (int t = 0; t < 12; t++) {
        acc |= s_lvl[warp][t][lane];
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
  uint32_t* img = s_img[warp];
  for (int j = lane; j < k; j += 32) {
    const SlabPoint p = slab[j];
    const double px = double(p.x), py = double(p.y), pz = double(p.z);
    const double hx = (F[0][0] * px + F[1][0] * py) + F[2][0] * pz;
    const double hy = (F[0][1] * px + F[1][1] * py) + F[2][1] * pz;
    const double rx = cs * hx + msn * hy;
    const double ry = sn * hx + cs * hy;
    if (ry < hc.bite[0] && rx > left && rx < right) {  // finger_hand.cpp:159-167
      wmin = fmin(wmin, rx);
      wmax = fmax(wmax, rx);
    }
    if (ry < lim) {  // points in the box (rotating_hand.cpp:126-130)
      m_box++;
      // learning.cpp:320-365 on points_for_learning = rotated point - surface (frame mix-up kept)
      const double bx = rx - surface3[0], by = ry - surface3[1];
      const double hcell = floor(((keep_sign ? bx : -bx) - (-0.05)) / hc.img_cell);
      const double vcell = floor((by - 0.0) / hc.img_cell);
      const int h = int(fmin(99.0, fmax(0.0, hcell)));
      const int v = int(fmin(79.0, fmax(0.0, vcell)));
      const int bit = (AG_IMAGE_ROWS - 1 - v) * AG_IMAGE_COLS + h;
      atomicOr(&img[bit >> 5], 1u << (bit & 31));
      if (p.tag & kTagNormalBit) {  // antipodal.cpp:12-86 on rot * frame^T * normal
        const double* nv = A.normals + size_t(3) * (p.tag >> 2);
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
  for (int i = lane; i < AG_IMAGE_WORDS; i += 32) gimg[i] = img[i];
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
  // fast slot lookup needs both hands' edge tables strictly ascending with a common step
  h.inv_slot_step = step > 0 ? 1.0 / step : 0.0;
  h.uniform_slots = step > 0 && p.finger_width > 0 ? 1 : 0;
  if (const char* e = getenv("AG_SWEEP_FAST")) h.uniform_slots = h.uniform_slots && atoi(e) != 0;  // diagnostics
  for (int i = 1; i < 10 && h.uniform_slots; i++)
    if (!(h.spacing[i] > h.spacing[i - 1]) || !(h.spacing[10 + i] > h.spacing[9 + i]) ||
        !(h.spacing[i] + p.finger_width > h.spacing[i - 1] + p.finger_width))
      h.uniform_slots = 0;
  h.img_cell = (0.05 - (-0.05)) / double(AG_IMAGE_COLS);  // learning.cpp:322-324
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
  A.sample_list = nullptr;
  A.n_samples = n;
  const double radius = c->params.nn_radius_hands;
  A.r2 = float(radius * radius);
  A.rpad = sqrt(double(A.r2)) * (1.0 + 1e-5) + 1e-7;
  A.filter_boundaries = (flags & 0x100u) ? 1 : 0;
  for (int i = 0; i < 6; i++) A.workspace[i] = c->params.workspace[i];
  return A;
}

static int compact(Ctx* c, const SweepArgs& A, size_t slots) {
  int* d_slots = c->hyp_slots.as<int>();
  int* d_nsel = d_slots + slots;
  k_compact_slots<<<1, 1024, 0, c->stream>>>(A.valid, int(slots), d_slots, d_nsel);
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

int* hand_sweep_count_ptr(Ctx* c, int n) { return c->hyp_slots.as<int>() + size_t(n) * 8; }
int* hand_sweep_overflow_ptr(Ctx* c) { return c->overflow.as<int>(); }

int hand_sweep_enqueue(Ctx* c, const int* d_indices, int n, const ag_frame* d_frames, unsigned flags) {
  c->n_hyp = 0;
  c->images_valid = false;
  if (n <= 0) return AG_OK;
  const size_t slots = size_t(n) * 8;
  if (c->grasps_raw.reserve(slots * sizeof(ag_grasp)) || c->valid.reserve(slots + 64) ||
      c->images_raw.reserve(slots * AG_IMAGE_WORDS * 4) || c->hyp_slots.reserve(slots * 4 + 16) ||
      c->grasps.reserve(slots * sizeof(ag_grasp)) || c->counters.reserve(64) ||
      c->sweep_dbg.reserve(size_t(n) * 4 + slots * 4) || c->overflow.reserve(size_t(n + 1) * 4))
    return AG_ERR_CUDA;
  SweepArgs A = make_args(c, d_indices, n, d_frames, flags);
  const size_t smem_small = size_t(kSlabCapSmall) * sizeof(SlabPoint);
  const size_t smem_big = size_t(kSlabCapBig) * sizeof(SlabPoint);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_hand_sweep<kSlabCapSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_small));
    cudaFuncSetAttribute(k_hand_sweep<kSlabCapSmall>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    cudaFuncSetAttribute(k_hand_sweep<kSlabCapBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_big));
    attr_set = true;
  }
  AG_CUDA_CHECK(cudaMemsetAsync(A.overflow, 0, 4, c->stream));
  k_hand_sweep<kSlabCapSmall><<<n, kThreads, smem_small, c->stream>>>(A, c->hand);
  c->launches += 2;  // + k_compact_grasps
  c->sweep_flags = flags;
  c->sweep_indices = d_indices;
  c->sweep_frames = d_frames;
  return compact(c, A, slots);
}

// Called after the stream has been synchronised and the overflow counter read.  Samples whose slab
// did not fit the 2048-point instantiation (dense neighbourhoods) are redone with the 12k-point one.
int hand_sweep_finish(Ctx* c, int n, int n_over, int* n_hyp) {
  const size_t slots = size_t(n) * 8;
  int* d_nsel = c->hyp_slots.as<int>() + slots;
  if (n_over > 0) {
    SweepArgs A = make_args(c, c->sweep_indices, n, c->sweep_frames, c->sweep_flags);
    DevBuf list;
    if (list.reserve(size_t(n_over) * 4)) return AG_ERR_CUDA;
    AG_CUDA_CHECK(cudaMemcpyAsync(list.p, A.overflow + 1, size_t(n_over) * 4, cudaMemcpyDeviceToDevice, c->stream));
    AG_CUDA_CHECK(cudaMemsetAsync(A.overflow, 0, 4, c->stream));
    A.sample_list = list.as<int>();
    const size_t smem_big = size_t(kSlabCapBig) * sizeof(SlabPoint);
    k_hand_sweep<kSlabCapBig><<<n_over, kThreads, smem_big, c->stream>>>(A, c->hand);
    c->launches += 2;
    int rc = compact(c, A, slots);
    if (rc) return rc;
    int over2 = 0;
    AG_CUDA_CHECK(cudaMemcpyAsync(&over2, A.overflow, 4, cudaMemcpyDeviceToHost, c->stream));
    AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    list.release();
    if (over2) {
      set_error("hand sweep: a sample's slab exceeded the shared-memory capacity (12032 points)");
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
