// hog_svm.cu — grasp image -> HOG(3528) -> SVM decision value, one CTA per hypothesis.
//
// Replaces (reference paths): cv::HOGDescriptor::compute as configured at
// src/agile_grasp/learning.cpp:194-195,220 (default-constructed descriptor: block 16, stride 8,
// cell 8, 9 bins, gamma correction, sigma 4, L2-Hys 0.2; winSize 64x64, winStride 32, padding 0
// => 2 windows x 7x7 blocks x 36 = 3528 floats) and CvSVM::predict at learning.cpp:225-226.
//
// The image is binary, so after OpenCV's sqrt gamma LUT every pixel gradient is one of 9 cases
// (dx,dy in {-s,0,+s}, s = sqrt(255.f)); magnitude/angle for the 9 cases are tabulated from
// cv::cartToPolar (which uses a polynomial atan: the diagonals are NOT k*pi/4).  Only pixels on the
// outline of the rasterised points have a gradient at all (config 2: ~160 of 8000 pixels, 93 % of the
// 16x16 blocks hold none), so the kernel is driven by those pixels: the gradient planes scatter their
// set bits into per-column masks and 8x8 tile flags, a block without a flagged tile is never visited
// (its histogram, its normalised descriptor entries and its terms of the SVM dot product are exact zeros),
// and a visited (block, cell) walks its contributing pixels in exactly OpenCV's order (count1 | count2 |
// count4 lists, column-major inside the block) with separate binary32 multiply and add, so the
// descriptor is bit-identical to OpenCV's; the two windows share 4 of their 7 block columns, so there
// are only 11x7 = 77 distinct blocks.  Compiled with -fmad=false.

#include <algorithm>
#include <cmath>
#include <cstring>

#include "ag_internal.h"

namespace ag {

namespace {

constexpr int W = AG_IMAGE_COLS, H = AG_IMAGE_ROWS;
constexpr int NB = 9, BS = 16, CS = 8;
constexpr int UBX = 11, UBY = 7, NUB = UBX * UBY;  // distinct blocks
constexpr int CELL_LIST = 144;                      // pixels contributing to one cell of a block
constexpr int kThreads = 256;                       // 8 warps per hypothesis (one per (block, cell) item at a time), 6 CTAs per SM

struct CellRun {     // consecutive rows of one block column that feed one cell, in OpenCV's order
  short j, i0, len, first;  // block column, first block row, run length, index of the first weight
};
constexpr int kMaxRuns = 24;
struct HogTables {
  float w[4][CELL_LIST];         // gradWeight * histWeight per contribution
  CellRun runs[4][kMaxRuns];
  int n_runs[4];
  float4 cases[9];               // per gradient case: magnitude share of the two nearest bins, and the bins (as int bits)
};
// global memory, read through the read-only path: the lookups are indexed per lane (constant memory would
// serialise the warp) and only the few (block, cell) items that see an outline pixel ever touch them
__device__ HogTables g_hog;

// cv::cartToPolar(dx,dy) for (sign dx, sign dy): index (sy+1)*3 + (sx+1); values measured from
// cv2 4.13 (tests/test_oracle_hog_svm.py checks them against the live library), stored as bit patterns
const uint32_t kMagBits[9] = {0x41b4aa5a, 0x417f7fe0, 0x41b4aa5a, 0x417f7fe0, 0x0, 0x417f7fe0, 0x41b4aa5a, 0x417f7fe0,
                              0x41b4aa5a};
const uint32_t kAngBits[9] = {0x407b5116, 0x4096cbe4, 0x40afef3d, 0x40490fdb, 0x0, 0x0, 0x4016ce9f, 0x3fc90fdb,
                              0x3f4904f0};

void build_tables(HogTables& T) {
  auto u2f = [](uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
  };
  // HOGDescriptor::computeGradient: bin = angle*(nbins/pi) - 0.5, linear split between 2 bins
  const float angleScale = float(NB / M_PI);
  for (int k = 0; k < 9; k++) {
    const float mag = u2f(kMagBits[k]);
    float angle = u2f(kAngBits[k]) * angleScale - 0.5f;
    int hidx = int(std::floor(angle));
    angle -= float(hidx);
    T.cases[k].x = mag * (1.f - angle);
    T.cases[k].y = mag * angle;
    if (hidx < 0) hidx += NB;
    else if (hidx >= NB) hidx -= NB;
    const int h0 = hidx;
    hidx++;
    if (hidx >= NB) hidx = 0;
    const int h1 = hidx;
    std::memcpy(&T.cases[k].z, &h0, 4);
    std::memcpy(&T.cases[k].w, &h1, 4);
  }
  // HOGCache::init: gaussian weights exp(-(di^2+dj^2)/(2 sigma^2)), di = i - 8, sigma = 4, and the
  // bilinear cell interpolation classes; OpenCV's per-pixel lists (pixels touching 1 | 2 | 4 cells, each
  // in column-major block order) are regrouped into per-cell lists that preserve the order in which
  // each histogram bin receives its contributions, then cut into runs of consecutive rows.
  float weights[BS][BS];
  const float sigma = 4.0f, scale = 1.f / (sigma * sigma * 2);
  float d2[BS];
  for (int i = 0; i < BS; i++) {
    d2[i] = float(i) - BS * 0.5f;
    d2[i] *= d2[i];
  }
  for (int i = 0; i < BS; i++)
    for (int j = 0; j < BS; j++) weights[i][j] = std::exp(-(d2[i] + d2[j]) * scale);
  struct Contribution {
    int cell, i, j;
    float w;
  };
  std::vector<Contribution> lists[3];  // pixels touching 1, 2, 4 cells
  for (int j = 0; j < BS; j++)
    for (int i = 0; i < BS; i++) {
      float cellX = (j + 0.5f) / CS - 0.5f, cellY = (i + 0.5f) / CS - 0.5f;
      const int ix0 = int(std::floor(cellX)), iy0 = int(std::floor(cellY));
      const int ix1 = ix0 + 1, iy1 = iy0 + 1;
      cellX -= ix0;
      cellY -= iy0;
      auto in = [](int v) { return v >= 0 && v < 2; };
      const float gw = weights[i][j];
      const float wx[2] = {1.f - cellX, cellX}, wy[2] = {1.f - cellY, cellY};
      const bool bx = in(ix0) && in(ix1), by = in(iy0) && in(iy1);
      std::vector<Contribution>& dst = lists[(bx ? 1 : 0) + (by ? 1 : 0)];
      if (bx && by) {  // order: (x0,y0) (x1,y0) (x0,y1) (x1,y1)
        dst.push_back({ix0 * 2 + iy0, i, j, gw * (wx[0] * wy[0])});
        dst.push_back({ix1 * 2 + iy0, i, j, gw * (wx[1] * wy[0])});
        dst.push_back({ix0 * 2 + iy1, i, j, gw * (wx[0] * wy[1])});
        dst.push_back({ix1 * 2 + iy1, i, j, gw * (wx[1] * wy[1])});
      } else if (bx) {  // two cells along x; y clamps to the one valid cell (weight cellY or 1-cellY)
        const int cy = in(iy0) ? iy0 : iy1;
        const float wyv = in(iy0) ? 1.f - cellY : cellY;
        dst.push_back({ix0 * 2 + cy, i, j, gw * (wx[0] * wyv)});
        dst.push_back({ix1 * 2 + cy, i, j, gw * (wx[1] * wyv)});
      } else if (by) {
        const int cx = in(ix0) ? ix0 : ix1;
        const float wxv = in(ix0) ? 1.f - cellX : cellX;
        dst.push_back({cx * 2 + iy0, i, j, gw * (wxv * wy[0])});
        dst.push_back({cx * 2 + iy1, i, j, gw * (wxv * wy[1])});
      } else {
        const int cx = in(ix0) ? ix0 : ix1, cy = in(iy0) ? iy0 : iy1;
        const float wxv = in(ix0) ? 1.f - cellX : cellX, wyv = in(iy0) ? 1.f - cellY : cellY;
        dst.push_back({cx * 2 + cy, i, j, gw * (wxv * wyv)});
      }
    }
  int fill[4] = {0, 0, 0, 0};
  int last_i[4], last_j[4];
  for (int c = 0; c < 4; c++) T.n_runs[c] = 0;
  for (int cls = 0; cls < 3; cls++) {
    for (int c = 0; c < 4; c++) last_i[c] = last_j[c] = -9;  // a run never spans two classes
    for (const Contribution& ct : lists[cls]) {
      const int c = ct.cell;
      T.w[c][fill[c]] = ct.w;
      if (ct.j == last_j[c] && ct.i == last_i[c] + 1) {
        T.runs[c][T.n_runs[c] - 1].len++;
      } else {
        CellRun& r = T.runs[c][T.n_runs[c]++];
        r.j = short(ct.j);
        r.i0 = short(ct.i);
        r.len = 1;
        r.first = short(fill[c]);
      }
      last_i[c] = ct.i;
      last_j[c] = ct.j;
      fill[c]++;
    }
  }
}

struct SvmDev {
  const float* sv;
  const double* alpha;
  const int* index;
  int sv_total, sv_count, kernel, degree;
  double gamma, coef0, rho;
};

constexpr int RW = 4;  // 32-bit words per image row (100 px -> 128-bit rows)

// ---- PTX helpers: mbarrier + TMA bulk copy (global -> shared), as in quadric.cu / sweep.cu ----------------
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

__device__ __forceinline__ uint32_t bits_at(const uint32_t* img, int bit0) {  // 32 bits starting at bit0
  const int w = bit0 >> 5, sh = bit0 & 31;
  return __funnelshift_r(img[w], img[w + 1], sh);
}

// Per-hypothesis flags of the distinct blocks that hold an outline pixel, written next to the descriptors for the
// batched scorer of many-vector models (3 words: bit ub of word ub / 32).
constexpr int kFlagWords = 3;

__global__ void __launch_bounds__(kThreads, 6)
k_hog_svm(const uint32_t* __restrict__ images, const int* __restrict__ image_slots, int n_bound,
          const int* __restrict__ n_dev, SvmDev svm, float* __restrict__ descriptors, float* __restrict__ scores,
          ag_grasp* __restrict__ grasps_out, int score_by_slot, uint32_t* __restrict__ block_flags, int two_ended) {
  // per warp: the ordered contributions of one cell (step 3); steps 1-2 use the same bytes for the packed and the
  // row-aligned image
  __shared__ __align__(16) float4 s_rec[kThreads / 32][CELL_LIST];
  static_assert(sizeof(float4) * (kThreads / 32) * CELL_LIST >= 4 * (256 + H * RW), "overlay");
  uint32_t* const s_stage = reinterpret_cast<uint32_t*>(&s_rec[0][0]);  // 1 KB: the 16-byte blocks holding the packed image
  uint32_t(*const s_row)[RW] = reinterpret_cast<uint32_t(*)[RW]>(s_stage + 256);
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_xp[H][RW], s_xn[H][RW], s_yp[H][RW], s_yn[H][RW];  // gradient sign bit-planes
  __shared__ uint32_t s_col[W][3];                  // per column: 80-bit mask of non-zero-gradient pixels
  __shared__ uint32_t s_tile[H / 8];                // per row of 8x8 tiles: bit tx = the tile holds such a pixel
  __shared__ __align__(16) float s_hist[NUB * 36];  // only the entries of flagged blocks are ever written or read
  __shared__ float4 s_case[9];
  __shared__ uint8_t s_list[NUB + 3];               // flagged blocks, ascending
  __shared__ int s_nflag;
  __shared__ double s_part[kThreads / 32];
  // (independent loads first: the count, this CTA's first image slot and the model's coefficient are in flight together)
  // two_ended (the sweep's list): n_dev[0] entries from the front of image_slots — the heavy images, dispatched first —
  // and n_dev[1] from its back (index n_bound - 1 downwards)
  const int first_slot = (image_slots && !two_ended && int(blockIdx.x) < n_bound) ? image_slots[blockIdx.x] : int(blockIdx.x);
  const double alpha0 = (svm.sv_total == 1 && svm.sv_count > 0) ? svm.alpha[0] : 0.0;
  const int n_front = n_dev ? min(n_dev[0], n_bound) : n_bound;
  const int n = two_ended ? min(n_front + n_dev[1], n_bound) : n_front;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  if (tid < 9) s_case[tid] = g_hog.cases[tid];
  const uint32_t bar = smem_u32(&s_bar);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_proxy_async_smem();
  }
  uint32_t parity = 0;
  // block ub = (bx, by) covers the tiles (bx..bx+1, by..by+1)
  auto flagged = [&](int ub) -> bool {
    const int bx = ub / UBY, by = ub - bx * UBY;
    return (((s_tile[by] | s_tile[by + 1]) >> bx) & 3u) != 0u;
  };
  for (int hyp = blockIdx.x; hyp < n; hyp += gridDim.x) {
    __syncthreads();  // the previous hypothesis is finished with the shared arrays
    const int slot = two_ended ? image_slots[hyp < n_front ? hyp : n_bound - 1 - (hyp - n_front)]
                               : (hyp == int(blockIdx.x) ? first_slot : (image_slots ? image_slots[hyp] : hyp));
    // the packed image (1000 B, contiguous) arrives by ONE TMA bulk copy: the 16-byte blocks that contain it (an image
    // starts on an 8-byte boundary), tracked by an mbarrier transaction count
    const char* src = reinterpret_cast<const char*>(images + size_t(slot) * AG_IMAGE_WORDS);
    const unsigned off = unsigned(reinterpret_cast<uintptr_t>(src) & 15u);
    const uint32_t bytes = (off + uint32_t(AG_IMAGE_WORDS) * 4u + 15u) & ~15u;
    if (tid == 0) {
      fence_proxy_async_smem();  // the generic-proxy accesses of the previous hypothesis precede the async writes
      mbar_expect_tx(bar, bytes);
      bulk_g2s(smem_u32(s_stage), src - off, bytes, bar);
    }
    const uint32_t* s_bits = s_stage + (off >> 2);
    for (int i = tid; i < W * 3; i += kThreads) (&s_col[0][0])[i] = 0u;
    if (tid < H / 8) s_tile[tid] = 0u;
    if (warp == 0) mbar_wait(bar, parity);
    parity ^= 1u;
    __syncthreads();
    // 1. row-aligned copy: row r = bits [100 r, 100 r + 100)
    for (int it = tid; it < H * RW; it += kThreads) {
      const int r = it >> 2, w = it & 3;
      uint32_t v = bits_at(s_bits, r * W + 32 * w);
      if (w == 3) v &= 0xFu;  // columns 96..99
      s_row[r][w] = v;
    }
    __syncthreads();
    // 2. gradient sign planes ([-1,0,1] derivative with BORDER_REFLECT_101 on the binary image); every pixel with
    //    a gradient is scattered into its column mask and flags its tile
    for (int it = tid; it < H * RW; it += kThreads) {
      const int r = it >> 2, w = it & 3;
      const uint32_t cur = s_row[r][w];
      const uint32_t prev = w > 0 ? s_row[r][w - 1] : 0u, next = w < 3 ? s_row[r][w + 1] : 0u;
      uint32_t R = (cur >> 1) | (next << 31);   // pixel x+1
      uint32_t L = (cur << 1) | (prev >> 31);   // pixel x-1
      if (w == 0) L = (L & ~1u) | ((cur >> 1) & 1u);                 // x = 0 : left neighbour is pixel 1
      if (w == 3) R = (R & ~(1u << 3)) | (((cur >> 2) & 1u) << 3);   // x = 99: right neighbour is pixel 98
      const uint32_t U = s_row[r == 0 ? 1 : r - 1][w], D = s_row[r == H - 1 ? H - 2 : r + 1][w];
      const uint32_t valid = w == 3 ? 0xFu : 0xFFFFFFFFu;
      const uint32_t xp = R & ~L & valid, xn = L & ~R & valid, yp = D & ~U & valid, yn = U & ~D & valid;
      s_xp[r][w] = xp;
      s_xn[r][w] = xn;
      s_yp[r][w] = yp;
      s_yn[r][w] = yn;
      uint32_t nz = xp | xn | yp | yn;
      if (nz) {
        const uint32_t tiles = ((nz & 0xFFu) ? 1u : 0u) | ((nz & 0xFF00u) ? 2u : 0u) | ((nz & 0xFF0000u) ? 4u : 0u) |
                               ((nz & 0xFF000000u) ? 8u : 0u);
        atomicOr(&s_tile[r >> 3], tiles << (4 * w));
        const uint32_t rbit = 1u << (r & 31);
        while (nz) {
          const int x = 32 * w + __ffs(nz) - 1;
          nz &= nz - 1;
          atomicOr(&s_col[x][r >> 5], rbit);
        }
      }
    }
    __syncthreads();
    // compact list of the flagged blocks (warp 0: three ballots over the 77 blocks)
    if (warp == 0) {
      int base = 0;
      for (int u0 = 0; u0 < NUB; u0 += 32) {
        const int ub = u0 + lane;
        const bool f = ub < NUB && flagged(ub);
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (f) s_list[base + __popc(m & lt)] = uint8_t(ub);
        base += __popc(m);
      }
      if (lane == 0) s_nflag = base;
    }
    __syncthreads();
    const int nflag = s_nflag;
    // 3. block histograms.  One WARP per (flagged block, cell).  The binary32 additions into a bin are a sequential
    //    chain in OpenCV's pixel order, but finding the pixels and forming their contributions is not: lane L decodes
    //    the pixels of run L of the cell (<= 24 runs of consecutive rows, in OpenCV's order) and stores their two
    //    contributions at the run's prefix offset; then lane b < 9 walks the ordered records and adds what belongs to
    //    bin b — the same values in the same order as HOGCache::getBlock.
    for (int item = warp; item < nflag * 4; item += kThreads / 32) {
      const int ub = s_list[item >> 2], cell = item & 3;
      const int ox = (ub / UBY) * 8, oy = (ub % UBY) * 8;
      float* hist = s_hist + ub * 36 + cell * 9;
      CellRun run = {0, 0, 0, 0};
      uint32_t m = 0;
      int x = 0, y0 = 0;
      if (lane < g_hog.n_runs[cell]) {
        run = g_hog.runs[cell][lane];
        x = ox + run.j;
        y0 = oy + run.i0;
        // bits y0 .. y0+len-1 of the 80-bit column mask
        const int wq = y0 >> 5, sh = y0 & 31;
        const uint32_t lo = s_col[x][wq], hi = wq < 2 ? s_col[x][wq + 1] : 0u;
        m = __funnelshift_r(lo, hi, sh) & ((1u << run.len) - 1u);
      }
      // exclusive prefix of the per-lane pixel counts (< 16: a run has at most 12 rows) from four ballots
      const int cnt = __popc(m);
      int excl = 0, total = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const unsigned bal = __ballot_sync(0xffffffffu, (cnt >> k) & 1);
        excl += __popc(bal & lt) << k;
        total += __popc(bal) << k;
      }
      if (total == 0) {  // no outline pixel feeds this cell
        if (lane < 9) hist[lane] = 0.f;
        continue;
      }
      float4* rec = s_rec[warp];
      int pos = excl;
      const int wx = x >> 5, bx = x & 31;
      while (m) {
        const int i = __ffs(m) - 1;
        m &= m - 1;
        const int y = y0 + i;
        const int sx = int((s_xp[y][wx] >> bx) & 1u) - int((s_xn[y][wx] >> bx) & 1u);
        const int sy = int((s_yp[y][wx] >> bx) & 1u) - int((s_yn[y][wx] >> bx) & 1u);
        const float4 cs = s_case[(sy + 1) * 3 + (sx + 1)];
        const float wgt = __ldg(&g_hog.w[cell][run.first + i]);
        rec[pos++] = make_float4(__fmul_rn(cs.x, wgt), __fmul_rn(cs.y, wgt), cs.z, cs.w);
      }
      __syncwarp();
      if (lane < 9) {
        float acc = 0.f;
        // (the two bins of a gradient are adjacent, never equal)
        auto take = [&](const float4& c4) {
          const bool b0 = __float_as_int(c4.z) == lane, b1 = __float_as_int(c4.w) == lane;
          if (b0 || b1) acc = __fadd_rn(acc, b0 ? c4.x : c4.y);
        };
        int r = 0;
        for (; r + 4 <= total; r += 4) {  // four records in flight: the chain is only the additions
          const float4 c0 = rec[r], c1 = rec[r + 1], c2 = rec[r + 2], c3 = rec[r + 3];
          take(c0);
          take(c1);
          take(c2);
          take(c3);
        }
        for (; r < total; r++) take(rec[r]);
        hist[lane] = acc;
      }
      __syncwarp();  // the records are consumed before the next item overwrites them
    }
    __syncthreads();
    // 4. L2-Hys normalisation per flagged block (HOGCache::normalizeBlockHistogram).  OpenCV keeps 4 interleaved
    //    partial sums (element i goes to partial i mod 4) and combines them as (p0+p1)+(p2+p3): thread l of
    //    a 4-thread group owns partial l, so the binary32 rounding sequence is identical.
    for (int base = 0; base < nflag * 4; base += kThreads) {  // (CTA-uniform trip count: the shuffles need every lane)
      const int it = base + tid;
      const bool act = it < nflag * 4;
      float* hist = s_hist + (act ? int(s_list[it >> 2]) : 0) * 36;
      const int l = tid & 3, gbase = lane & ~3;
      float v[9];
      float part = 0.f;
#pragma unroll
      for (int k = 0; k < 9; k++) {
        v[k] = act ? hist[l + 4 * k] : 0.f;
        part = __fadd_rn(part, __fmul_rn(v[k], v[k]));
      }
      float p0 = __shfl_sync(0xffffffffu, part, gbase), p1 = __shfl_sync(0xffffffffu, part, gbase + 1);
      float p2 = __shfl_sync(0xffffffffu, part, gbase + 2), p3 = __shfl_sync(0xffffffffu, part, gbase + 3);
      float sum = __fadd_rn(__fadd_rn(p0, p1), __fadd_rn(p2, p3));
      float scale = __fdiv_rn(1.f, __fadd_rn(__fsqrt_rn(sum), __fmul_rn(36.f, 0.1f)));
      part = 0.f;
#pragma unroll
      for (int k = 0; k < 9; k++) {
        v[k] = fminf(__fmul_rn(v[k], scale), 0.2f);
        part = __fadd_rn(part, __fmul_rn(v[k], v[k]));
      }
      p0 = __shfl_sync(0xffffffffu, part, gbase);
      p1 = __shfl_sync(0xffffffffu, part, gbase + 1);
      p2 = __shfl_sync(0xffffffffu, part, gbase + 2);
      p3 = __shfl_sync(0xffffffffu, part, gbase + 3);
      sum = __fadd_rn(__fadd_rn(p0, p1), __fadd_rn(p2, p3));
      scale = __fdiv_rn(1.f, __fadd_rn(__fsqrt_rn(sum), 1e-3f));
      if (act) {  // (every entry is read and rewritten by the same thread)
#pragma unroll
        for (int k = 0; k < 9; k++) hist[l + 4 * k] = __fmul_rn(v[k], scale);
      }
    }
    __syncthreads();
    // descriptor group k4 (4 consecutive floats) -> distinct block and histogram entry:
    // k = ((w*7 + bx)*7 + by)*36 + e
    auto group_block = [&](int k4, int& e) -> int {
      const int blk = k4 / 9;
      e = (k4 - blk * 9) * 4;
      const int by = blk % 7, bxw = blk / 7;          // bxw = w*7 + bx
      const int ubx = (bxw / 7) * 4 + (bxw % 7);
      return ubx * UBY + by;
    };
    if (descriptors)
      for (int k4 = tid; k4 < AG_HOG_DIM / 4; k4 += kThreads) {
        int e;
        const int ub = group_block(k4, e);
        float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);  // a block without an outline pixel: exact zeros
        if (flagged(ub)) d4 = *reinterpret_cast<const float4*>(s_hist + ub * 36 + e);
        reinterpret_cast<float4*>(descriptors + size_t(hyp) * AG_HOG_DIM)[k4] = d4;
      }
    if (block_flags && tid < kFlagWords) {
      uint32_t f = 0;
      for (int ub = tid * 32; ub < min(NUB, tid * 32 + 32); ub++) f |= (flagged(ub) ? 1u : 0u) << (ub & 31);
      block_flags[size_t(hyp) * kFlagWords + tid] = f;
    }
    if (svm.sv_total != 1) continue;  // descriptors only: the batched kernels below score them (CTA uniform)
    // 5. SVM kernel value (CvSVMKernel::calc_non_rbf_base): binary32 products, 4-term binary32 sums, binary64
    //    accumulation; the groups of blocks without an outline pixel are exact zeros and are skipped
    //    (a distinct block column bx is column bx of window 0 when bx <= 6 and column bx - 4 of window 1 when bx >= 4)
    double acc = 0.0;
    for (int e = tid; e < nflag * 18; e += kThreads) {
      const int f = e / 18, rem = e - f * 18, w = rem >= 9 ? 1 : 0, g = rem - 9 * w;
      const int ub = s_list[f], bx = ub / UBY, by = ub - bx * UBY;
      if (w == 0 ? bx > 6 : bx < 4) continue;
      const int k4 = ((w * 7 + (bx - 4 * w)) * 7 + by) * 9 + g;
      const float4 s4 = __ldg(reinterpret_cast<const float4*>(svm.sv) + k4);
      const float4 d4 = *reinterpret_cast<const float4*>(s_hist + ub * 36 + g * 4);
      float t = __fmul_rn(s4.x, d4.x);
      t = __fadd_rn(t, __fmul_rn(s4.y, d4.y));
      t = __fadd_rn(t, __fmul_rn(s4.z, d4.z));
      t = __fadd_rn(t, __fmul_rn(s4.w, d4.w));
      acc += double(t);
    }
    acc = warp_sum(acc);
    if (lane == 0) s_part[warp] = acc;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w2 = 0; w2 < kThreads / 32; w2++) t += s_part[w2];
      float kv;
      if (svm.kernel == 0) {
        kv = float(t * 1.0 + 0.0);
      } else {
        kv = float(t * svm.gamma + svm.coef0);
        float b = kv, a = 1.f;  // cv::pow with an integer exponent: repeated binary32 multiplication
        int p = svm.degree;
        while (p > 1) {
          if (p & 1) a = __fmul_rn(a, b);
          b = __fmul_rn(b, b);
          p >>= 1;
        }
        kv = __fmul_rn(a, b);
      }
      // decision value: sum = -rho + alpha_0 * K   (CvSVM::predict)
      double sum = -svm.rho;
      if (svm.sv_count > 0) sum += alpha0 * double(kv);
      const float sc = float(sum);
      scores[score_by_slot ? slot : hyp] = sc;
      if (grasps_out) {  // fused classify: write score and label into the compacted record
        grasps_out[hyp].score = sc;
        grasps_out[hyp].label = sc > 0.f ? 0 : 1;  // CvSVM::predict: label +1 <=> sum <= 0
      }
    }
  }  // hypothesis loop
}

// ---- batched scoring for models with many support vectors (POLY: 588 / 1190 x 3528) -------------------
// CvSVM::predict evaluates, per hypothesis, sv_total dot products of length 3528.  Batched over the H
// hypotheses of a cloud that is a [H x 3528] . [3528 x sv_total] product; it is NOT routed to tensor cores:
// OpenCV's calc_non_rbf_base rounds every product to binary32, sums groups of four in binary32 and
// accumulates the groups in binary64, and the scores must match that bit for bit.  So this is a classic
// shared-memory tiled SIMT kernel in which every thread keeps the reference's order for its 4 x 4 outputs:
// 32 hypotheses x 64 support vectors per CTA (128 threads), K tiles of 24 (3528 = 147 x 24) double-buffered with
// cp.async, row pitch 28 floats so the 16-byte reads of a quarter warp hit distinct banks.  The support
// vector matrix is read H/64 times instead of H times.
constexpr int GM = 32, GN = 64, GK = 24, GP = 28, kGemmThreads = 128;
static_assert(AG_HOG_DIM % GK == 0, "K tiles must divide the descriptor length");

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool valid) {
  const uint32_t d = uint32_t(__cvta_generic_to_shared(dst_smem));
  const int sz = valid ? 16 : 0;  // 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

__device__ __forceinline__ float svm_kernel_value(double acc, int kernel, double gamma, double coef0, int degree) {
  if (kernel == 0) return float(acc * 1.0 + 0.0);
  float kv = float(acc * gamma + coef0);
  float b = kv, a = 1.f;  // cv::pow with an integer exponent: repeated binary32 multiplication
  int p = degree;
  while (p > 1) {
    if (p & 1) a = __fmul_rn(a, b);
    b = __fmul_rn(b, b);
    p >>= 1;
  }
  return __fmul_rn(a, b);
}

__global__ void __launch_bounds__(kGemmThreads)
k_svm_gemm(const float* __restrict__ desc, const float* __restrict__ sv, int n_bound, const int* __restrict__ n_dev,
           int nsv, int kernel, double gamma, double coef0, int degree, float* __restrict__ kvals) {
  __shared__ __align__(16) float sA[2][GM][GP];
  __shared__ __align__(16) float sB[2][GN][GP];
  const int n = n_dev ? min(*n_dev, n_bound) : n_bound;
  const int h0 = blockIdx.x * GM, k0 = blockIdx.y * GN;
  if (h0 >= n) return;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  auto load_tiles = [&](int kt, int st) {
    for (int e = t; e < (GM + GN) * (GK / 4); e += kGemmThreads) {
      const bool isB = e >= GM * (GK / 4);
      const int f = isB ? e - GM * (GK / 4) : e;
      const int r = f / (GK / 4), c4 = f % (GK / 4);
      const int row = (isB ? k0 : h0) + r;
      const bool valid = row < (isB ? nsv : n);
      const float* src = (isB ? sv : desc) + size_t(valid ? row : 0) * AG_HOG_DIM + kt * GK + c4 * 4;
      cp_async16(isB ? &sB[st][r][c4 * 4] : &sA[st][r][c4 * 4], src, valid);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
  load_tiles(0, 0);
  constexpr int NT = AG_HOG_DIM / GK;
  for (int kt = 0; kt < NT; kt++) {
    const int st = kt & 1;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // tile kt has landed; every thread is done with tile kt - 1
    if (kt + 1 < NT) load_tiles(kt + 1, st ^ 1);
#pragma unroll
    for (int g = 0; g < GK / 4; g++) {
      float4 a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = *reinterpret_cast<const float4*>(&sA[st][ty * 4 + i][g * 4]);
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = *reinterpret_cast<const float4*>(&sB[st][tx + 16 * j][g * 4]);
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float p = __fmul_rn(b[j].x, a[i].x);
          p = __fadd_rn(p, __fmul_rn(b[j].y, a[i].y));
          p = __fadd_rn(p, __fmul_rn(b[j].z, a[i].z));
          p = __fadd_rn(p, __fmul_rn(b[j].w, a[i].w));
          acc[i][j] += double(p);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int h = h0 + ty * 4 + i;
    if (h >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int k = k0 + tx + 16 * j;
      if (k < nsv) kvals[size_t(h) * nsv + k] = svm_kernel_value(acc[i][j], kernel, gamma, coef0, degree);
    }
  }
}

// ---- kernel values of many-vector models from the NON-ZERO part of the descriptors ---------------------
// Only the blocks that hold an outline pixel have non-zero descriptor entries (k_hog_svm), ~8 % of the 3528 at
// config 2, and a zero group contributes an exact +-0 to calc_non_rbf_base's binary64 accumulation — so the kernel
// value of (hypothesis, support vector) is the sum over the flagged groups only, in ascending feature order: the
// same operations on the same values as the dense product above, bit for bit.  One CTA = SB hypotheses x 256
// support vectors (blockIdx.y): the window blocks flagged in ANY of the SB hypotheses are staged in shared memory
// (unflagged ones as zeros), thread = support vector: one coalesced 16-byte load of the transposed model per group
// serves all SB hypotheses.
constexpr int SB = 4, kSparseThreads = 256, kSparseChunk = 24;  // window blocks staged per pass
__global__ void __launch_bounds__(kSparseThreads)
k_svm_sparse(const float* __restrict__ desc, const uint32_t* __restrict__ flags, int n_bound, const int* __restrict__ n_dev,
             const float4* __restrict__ svT, int nsv, int kernel, double gamma, double coef0, int degree,
             float* __restrict__ kvals) {
  __shared__ uint8_t s_wb[2 * 7 * 7];
  __shared__ int s_nwb;
  __shared__ __align__(16) float4 s_d[kSparseChunk * 9][SB];
  const int n = n_dev ? min(*n_dev, n_bound) : n_bound;
  const int tid = threadIdx.x, lane = tid & 31;
  const int k = blockIdx.y * kSparseThreads + tid;  // this thread's support vector
  for (int h0 = blockIdx.x * SB; h0 < n; h0 += gridDim.x * SB) {
    const int nh = min(SB, n - h0);
    __syncthreads();
    if (tid < 32) {  // window blocks flagged in any hypothesis of the batch, ascending = ascending feature index
      uint32_t f[kFlagWords] = {0u, 0u, 0u};
      for (int i = 0; i < nh; i++)
        for (int w = 0; w < kFlagWords; w++) f[w] |= flags[size_t(h0 + i) * kFlagWords + w];
      int base = 0;
      for (int w0 = 0; w0 < 98; w0 += 32) {
        const int wb = w0 + lane;
        bool on = false;
        if (wb < 98) {
          const int by = wb % 7, bxw = wb / 7, ub = ((bxw / 7) * 4 + bxw % 7) * UBY + by;
          on = (f[ub >> 5] >> (ub & 31)) & 1u;
        }
        const unsigned m = __ballot_sync(0xffffffffu, on);
        if (on) s_wb[base + __popc(m & ((1u << lane) - 1u))] = uint8_t(wb);
        base += __popc(m);
      }
      if (lane == 0) s_nwb = base;
    }
    __syncthreads();
    const int nwb = s_nwb;
    double acc[SB];
#pragma unroll
    for (int i = 0; i < SB; i++) acc[i] = 0.0;
    for (int c0 = 0; c0 < nwb; c0 += kSparseChunk) {
      const int cn = min(kSparseChunk, nwb - c0);
      __syncthreads();
      for (int e = tid; e < cn * 9 * SB; e += kSparseThreads) {  // stage: [window block][group][hypothesis]
        const int i = e % SB, g = (e / SB) % 9, u = e / (SB * 9);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < nh) v = *reinterpret_cast<const float4*>(desc + size_t(h0 + i) * AG_HOG_DIM + int(s_wb[c0 + u]) * 36 + g * 4);
        s_d[u * 9 + g][i] = v;
      }
      __syncthreads();
      if (k < nsv) {
        for (int u = 0; u < cn; u++) {
          const float4* sv_g = svT + size_t(int(s_wb[c0 + u]) * 9) * nsv + k;
#pragma unroll
          for (int g = 0; g < 9; g++) {
            const float4 b = __ldg(sv_g + size_t(g) * nsv);
#pragma unroll
            for (int i = 0; i < SB; i++) {
              const float4 a = s_d[u * 9 + g][i];
              float p = __fmul_rn(b.x, a.x);
              p = __fadd_rn(p, __fmul_rn(b.y, a.y));
              p = __fadd_rn(p, __fmul_rn(b.z, a.z));
              p = __fadd_rn(p, __fmul_rn(b.w, a.w));
              acc[i] += double(p);
            }
          }
        }
      }
    }
    if (k < nsv)
      for (int i = 0; i < nh; i++) kvals[size_t(h0 + i) * nsv + k] = svm_kernel_value(acc[i], kernel, gamma, coef0, degree);
  }
}

// decision value per hypothesis, in CvSVM::predict's order: sum = -rho; sum += alpha_k * K[index_k], k ascending.
// One warp per 32 hypotheses: tiles of 32 x 32 kernel values are loaded coalesced into shared memory, then
// lane r walks row r sequentially (binary64 multiply, then add — no contraction).  `identity` = index[k] == k
// for all k (what CvSVM writes for two-class models); otherwise the row is read directly.
__global__ void __launch_bounds__(32)
k_svm_decide(const float* __restrict__ kvals, int n_bound, const int* __restrict__ n_dev, SvmDev svm, int identity,
             float* __restrict__ scores, ag_grasp* __restrict__ grasps_out) {
  __shared__ float s_tile[32][33];
  const int n = n_dev ? min(*n_dev, n_bound) : n_bound;
  const int h0 = blockIdx.x * 32, lane = threadIdx.x;
  if (h0 >= n) return;
  const int h = h0 + lane;
  const int nsv = svm.sv_total;
  double sum = -svm.rho;
  if (identity) {
    for (int k0 = 0; k0 < svm.sv_count; k0 += 32) {
      __syncwarp();
      for (int r = 0; r < 32; r++)
        s_tile[r][lane] = (h0 + r < n && k0 + lane < svm.sv_count) ? kvals[size_t(h0 + r) * nsv + k0 + lane] : 0.f;
      __syncwarp();
      const int kn = min(32, svm.sv_count - k0);
      for (int kk = 0; kk < kn; kk++) sum += svm.alpha[k0 + kk] * double(s_tile[lane][kk]);
    }
  } else if (h < n) {
    const float* K = kvals + size_t(h) * nsv;
    for (int k = 0; k < svm.sv_count; k++) sum += svm.alpha[k] * double(K[svm.index[k]]);
  }
  if (h >= n) return;
  const float sc = float(sum);
  scores[h] = sc;
  if (grasps_out) {
    grasps_out[h].score = sc;
    grasps_out[h].label = sc > 0.f ? 0 : 1;  // CvSVM::predict: label +1 <=> sum <= 0
  }
}

bool g_tables_ready[64] = {false};

}  // namespace

int svm_to_device(SvmModel* svm, int device) {
  if (svm->device == device && svm->d_sv) return AG_OK;
  if (svm->d_sv) {
    cudaFree(svm->d_sv);
    if (svm->d_svT) cudaFree(svm->d_svT);
    cudaFree(svm->d_alpha);
    cudaFree(svm->d_index);
    svm->d_sv = nullptr;
    svm->d_svT = nullptr;
  }
  AG_CUDA_CHECK(cudaMalloc(&svm->d_sv, svm->sv.size() * sizeof(float)));
  AG_CUDA_CHECK(cudaMalloc(&svm->d_alpha, svm->alpha.size() * sizeof(double)));
  AG_CUDA_CHECK(cudaMalloc(&svm->d_index, svm->index.size() * sizeof(int)));
  AG_CUDA_CHECK(cudaMemcpy(svm->d_sv, svm->sv.data(), svm->sv.size() * sizeof(float), cudaMemcpyHostToDevice));
  AG_CUDA_CHECK(cudaMemcpy(svm->d_alpha, svm->alpha.data(), svm->alpha.size() * sizeof(double), cudaMemcpyHostToDevice));
  AG_CUDA_CHECK(cudaMemcpy(svm->d_index, svm->index.data(), svm->index.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (svm->sv_total > 1 && svm->var_count == AG_HOG_DIM) {
    // k_svm_sparse reads group g (4 consecutive features) of consecutive support vectors with consecutive lanes
    const int G = AG_HOG_DIM / 4, nsv = svm->sv_total;
    std::vector<float> t(size_t(G) * nsv * 4);
    for (int k = 0; k < nsv; k++)
      for (int g = 0; g < G; g++) std::memcpy(&t[(size_t(g) * nsv + k) * 4], &svm->sv[size_t(k) * AG_HOG_DIM + g * 4], 16);
    AG_CUDA_CHECK(cudaMalloc(&svm->d_svT, t.size() * sizeof(float)));
    AG_CUDA_CHECK(cudaMemcpy(svm->d_svT, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
  }
  svm->device = device;
  return AG_OK;
}

static int ensure_hog_tables(Ctx* c) {
  if (!g_tables_ready[c->device & 63]) {
    HogTables T;
    std::memset(&T, 0, sizeof(T));
    build_tables(T);
    AG_CUDA_CHECK(cudaMemcpyToSymbol(g_hog, &T, sizeof(T)));
    cudaFuncSetAttribute(k_hog_svm, cudaFuncAttributePreferredSharedMemoryCarveout, 100);  // 6 CTAs x 36 KB per SM
    g_tables_ready[c->device & 63] = true;
  }
  return AG_OK;
}

// HOG descriptors only (training features, learning.cpp:249-290): n packed images -> n x 3528 floats
int hog_descriptors_device(Ctx* c, const uint32_t* d_images, const int* d_image_slots, int n, float* d_descriptors) {
  if (n <= 0) return AG_OK;
  int rc = ensure_hog_tables(c);
  if (rc) return rc;
  SvmDev none;
  std::memset(&none, 0, sizeof(none));
  k_hog_svm<<<std::min(n, kNumSMs * 16), kThreads, 0, c->stream>>>(d_images, d_image_slots, n, nullptr, none, d_descriptors,
                                                                  nullptr, nullptr, 0, nullptr, 0);
  c->launches += 1;
  AG_CUDA_CHECK(cudaGetLastError());
  return AG_OK;
}

int hog_svm_device(Ctx* c, SvmModel* svm, const uint32_t* d_images, const int* d_image_slots, int n, const int* n_dev,
                   float* d_descriptors, float* d_scores, ag_grasp* d_grasps_out, bool score_by_slot) {
  if (n <= 0) return AG_OK;
  if (score_by_slot && (svm->sv_total > 1 || !d_image_slots)) {
    set_error("hog_svm_device: scores by slot need a single-vector model and a slot list");
    return AG_ERR_INVALID;
  }
  if (svm->var_count != AG_HOG_DIM) {
    set_error("SVM var_count != 3528");
    return AG_ERR_INVALID;
  }
  {
    int rc0 = ensure_hog_tables(c);
    if (rc0) return rc0;
  }
  int rc = svm_to_device(svm, c->device);
  if (rc) return rc;
  SvmDev sd;
  sd.sv = svm->d_sv;
  sd.alpha = svm->d_alpha;
  sd.index = svm->d_index;
  sd.sv_total = svm->sv_total;
  sd.sv_count = svm->sv_count;
  sd.kernel = svm->kernel;
  sd.degree = svm->degree;
  sd.gamma = svm->gamma;
  sd.coef0 = svm->coef0;
  sd.rho = svm->rho;
  // one CTA per hypothesis slot; the count lives on the device, CTAs beyond it leave at once and the block scheduler
  // balances light and heavy images (the outline of a dense image costs 10x the average)
  const int grid = std::min(n, kNumSMs * 16);
  if (svm->sv_total > 1) {
    // many support vectors: descriptors -> [H x 3528] . [3528 x sv_total] tiled product -> decision values
    float* desc = d_descriptors;
    if (!desc) {
      if (c->descriptors.reserve(size_t(n) * AG_HOG_DIM * sizeof(float))) return AG_ERR_CUDA;
      desc = c->descriptors.as<float>();
    }
    if (c->kvals.reserve(size_t(n) * svm->sv_total * sizeof(float))) return AG_ERR_CUDA;
    SvmDev none = sd;
    none.sv_total = 0;
    if (c->block_flags.reserve(size_t(n) * kFlagWords * 4)) return AG_ERR_CUDA;
    k_hog_svm<<<grid, kThreads, 0, c->stream>>>(d_images, d_image_slots, n, n_dev, none, desc, d_scores, nullptr, 0,
                                                c->block_flags.as<uint32_t>(), 0);
    static const bool dense = getenv("AG_SVM_DENSE") != nullptr;  // diagnostics: the dense tiled product
    if (dense || !svm->d_svT) {
      const dim3 gg((n + GM - 1) / GM, (svm->sv_total + GN - 1) / GN);
      k_svm_gemm<<<gg, kGemmThreads, 0, c->stream>>>(desc, svm->d_sv, n, n_dev, svm->sv_total, svm->kernel, svm->gamma,
                                                    svm->coef0, svm->degree, c->kvals.as<float>());
    } else {
      const dim3 gs(std::min((n + SB - 1) / SB, kNumSMs * 8), (svm->sv_total + kSparseThreads - 1) / kSparseThreads);
      k_svm_sparse<<<gs, kSparseThreads, 0, c->stream>>>(desc, c->block_flags.as<uint32_t>(), n, n_dev,
                                                         reinterpret_cast<const float4*>(svm->d_svT), svm->sv_total,
                                                         svm->kernel, svm->gamma, svm->coef0, svm->degree,
                                                         c->kvals.as<float>());
    }
    bool identity = svm->sv_count <= svm->sv_total;
    for (int k = 0; k < svm->sv_count && identity; k++) identity = svm->index[k] == k;
    k_svm_decide<<<(n + 31) / 32, 32, 0, c->stream>>>(c->kvals.as<float>(), n, n_dev, sd, identity ? 1 : 0, d_scores,
                                                     d_grasps_out);
    c->launches += 3;
  } else {
    k_hog_svm<<<grid, kThreads, 0, c->stream>>>(d_images, d_image_slots, n, n_dev, sd, d_descriptors, d_scores,
                                                 d_grasps_out, score_by_slot ? 1 : 0, nullptr, score_by_slot ? 1 : 0);
    c->launches += 1;
  }
  AG_CUDA_CHECK(cudaGetLastError());
  return AG_OK;
}

}  // namespace ag
