// preprocess.cu — NaN removal, workspace filter, voxelisation and the x-row / column index build on the GPU.
//
// Replaces (reference paths): pcl::removeNaNFromPointCloud + camera labelling
// (src/agile_grasp/localization.cpp:17-27), Localization::filterWorkspace (:216-245),
// Localization::voxelizeCloud (:247-355; std::set<Vector3i> -> 64-bit key radix sort + unique) and
// pcl::KdTreeFLANN::setInputCloud (src/agile_grasp/hand_search.cpp:10-11; kd-tree -> per-camera x-row
// table over the already sorted voxel list, see ag_common.cuh).  All of it is HBM-streaming
// integer/byte work.  Nothing here synchronises with the host: counts stay in device memory.
//
// Two voxelisation paths produce the same bits:
//  * occupancy bitmap (the common case: the voxel lattice of the cropped scene fits kBitmapBytes): one bit per
//    lattice cell, laid out (camera, x, y, z) so that the set bits in address order ARE the reference's
//    std::set order (localization.h:281-292).  k_classify (flags, per-camera extents, lattice dimensions by its
//    last CTA) -> k_mark (one atomic OR per kept point: duplicates collapse for free) -> k_emit_bitmap (one pass:
//    popcount, single-pass decoupled look-back scan across tiles, voxel records, the dense (kx, ky) column table
//    and the x-row table as by-products of the scan, cloud_normals_ zeroed per voxel; every word is cleared
//    again as it is read, so the bitmap is clean for the next call).  Three launches, no sort, no library call.
//  * key sort (any extent): 64-bit keys, CUB radix sort + unique.  Taken when the lattice does not fit the
//    bitmap; the device reports that (kErrBitmapRetry) and the host re-runs the call on this path.

#include <cstring>
#include <algorithm>

#include <cub/cub.cuh>

#include "ag_internal.h"

namespace ag {

namespace {

constexpr int kBlock = 256;
constexpr uint64_t kInvalidKey = ~0ull;
constexpr int kColCap = 4 << 20;  // entries of the dense (kx, ky) column table (16 MB); larger extents fall back
constexpr size_t kBitmapBytes = size_t(64) << 20;  // occupancy bitmap: 512 M lattice cells (e.g. 2.4 m x 2.4 m x 0.8 m at 3 mm)

// voxel key = cam | kx | ky | kz packed with per-axis bit widths derived from the workspace extent, so
// the radix sort only touches the bits that can be set
struct KeyBits {
  int bx, by, bz;  // bits per axis (<= 21)
  int total;       // 1 + bx + by + bz
};

struct PreState {
  int cam_min[2][3];  // ordered-int encoded float minima per camera
  int cam_max[2][3];  // maxima per camera (extent of the column table / of the occupancy bitmap)
  int n_unique;       // output of DeviceSelect::Unique
  int n_vox;
  int key_overflow;
  // occupancy-bitmap path: lattice dimensions per camera (published by the last CTA of k_classify)
  int bnx[2], bny[2], bnzw[2];       // x cells, y cells, 32-bit words per (x, y) column
  unsigned long long bword0[2];      // first bitmap word of each camera
  unsigned long long bwords;         // words in use (both cameras)
  int bcol0[2];                      // offset of each camera inside the column table
  int bitmap_fail;                   // the lattice does not fit: the host re-runs the call on the key-sort path
  unsigned classify_done;            // CTA completion counter of k_classify
  unsigned tile_ticket;              // next tile of k_emit_bitmap
  unsigned emit_done;                // CTA completion counter of k_emit_bitmap
  unsigned epoch;                    // call counter tagging the tile states of the scan (never reset)
};
constexpr int kTileWords = 2048;     // bitmap words per scan tile (256 threads x 8 words)

__device__ __forceinline__ bool load_xyz(const char* base, int stride, int i, float& x, float& y, float& z) {
  const char* p = base + size_t(i) * stride;
  if ((stride & 15) == 0) {  // PointXYZRGBA: 16-byte aligned records -> one 128-bit load
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    x = v.x; y = v.y; z = v.z;
  } else {
    const float* f = reinterpret_cast<const float*>(p);
    x = __ldg(f); y = __ldg(f + 1); z = __ldg(f + 2);
  }
  // pcl::removeNaNFromPointCloud keeps points whose x, y and z are all finite
  return isfinite(x) && isfinite(y) && isfinite(z);
}

struct ResetList {  // small device buffers zeroed by k_init_state (word counts)
  unsigned* p[5];
  int words[5];
};
__global__ void k_init_state(PreState* st, ResetList rl) {
  if (threadIdx.x == 0) {
    for (int c = 0; c < 2; c++)
      for (int a = 0; a < 3; a++) st->cam_min[c][a] = float_to_ordered(10000.0f);  // localization.cpp:251-252
    for (int c = 0; c < 2; c++)
      for (int a = 0; a < 3; a++) st->cam_max[c][a] = float_to_ordered(-3.0e38f);
    st->n_unique = 0;
    st->n_vox = 0;
    st->key_overflow = 0;
    st->bitmap_fail = 0;
    st->classify_done = 0;
    st->tile_ticket = 0;
    st->emit_done = 0;
    st->epoch++;
  }
  for (int k = 0; k < 5; k++)
    if (rl.p[k])
      for (int i = threadIdx.x; i < rl.words[k]; i += blockDim.x) rl.p[k][i] = 0u;
}

// number of finite points per block (needed only for the reference's label-after-compaction quirk)
__global__ void k_count_finite(const char* pts, int stride, int n, int* block_counts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float x, y, z;
  bool fin = i < n && load_xyz(pts, stride, i, x, y, z);
  int cnt = __syncthreads_count(fin);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = cnt;
}

__global__ void k_scan_blocks(int* block_counts, int nb) {  // single block, exclusive scan in place
  __shared__ int carry;
  __shared__ int wsum[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < nb ? block_counts[i] : 0;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane == 31) wsum[w] = s;
    __syncthreads();
    if (w == 0) {
      int t = lane < (blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      wsum[lane] = t;
    }
    __syncthreads();
    int excl = s - v + (w > 0 ? wsum[w - 1] : 0) + carry;
    if (i < nb) block_counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
}

// flag[i] = 0 dropped, 1 camera 0, 2 camera 1; per-camera coordinate minima.  Persistent CTAs walk the
// 256-point chunks with a grid stride and keep the minima in registers, so each CTA issues one atomic
// per camera and axis at the very end (same-address atomics serialise in L2: one per chunk was the
// dominant cost of this kernel).
__global__ void __launch_bounds__(kBlock)
k_classify(const char* pts, int stride, int n, int size_left, const int* block_offsets, double w0, double w1, double w2,
           double w3, double w4, double w5, uint8_t* flag, PreState* st, double cell, unsigned long long bitmap_words,
           int col_cap, KeyBits kb) {
  __shared__ int wcount[kBlock / 32];
  __shared__ int s_min[12][kBlock / 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int mn[6] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF};
  int mxv[6] = {int(0x80000000), int(0x80000000), int(0x80000000), int(0x80000000), int(0x80000000), int(0x80000000)};
  const int n_chunks = (n + kBlock - 1) / kBlock;
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int i = chunk * kBlock + threadIdx.x;
    float x = 0, y = 0, z = 0;
    const bool fin = i < n && load_xyz(pts, stride, i, x, y, z);
    int label;
    if (block_offsets) {
      // localization.cpp:19-27: labels are assigned by position BEFORE NaN removal but read AFTER
      // compaction, i.e. finite point #k gets label (k < size_left ? 0 : 1)   
      const unsigned m = __ballot_sync(0xffffffffu, fin);
      __syncthreads();
      if (lane == 0) wcount[w] = __popc(m);
      __syncthreads();
      int before = block_offsets[chunk];
      for (int k = 0; k < w; k++) before += wcount[k];
      const int rank = before + __popc(m & ((1u << lane) - 1));
      label = rank < size_left ? 0 : 1;
    } else {
      label = i < size_left ? 0 : 1;
    }
    // localization.cpp:228-229 inclusive workspace box, float promoted to double
    const bool keep = fin && double(x) >= w0 && double(x) <= w1 && double(y) >= w2 && double(y) <= w3 &&
                      double(z) >= w4 && double(z) <= w5;
    if (i < n) flag[i] = keep ? uint8_t(1 + label) : uint8_t(0);
    if (keep) {  // localization.cpp:256-277
      const int o = label * 3;
      mn[o] = min(mn[o], float_to_ordered(x));
      mn[o + 1] = min(mn[o + 1], float_to_ordered(y));
      mn[o + 2] = min(mn[o + 2], float_to_ordered(z));
      mxv[o] = max(mxv[o], float_to_ordered(x));
      mxv[o + 1] = max(mxv[o + 1], float_to_ordered(y));
      mxv[o + 2] = max(mxv[o + 2], float_to_ordered(z));
    }
  }
#pragma unroll
  for (int a = 0; a < 6; a++) {
    const int r = __reduce_min_sync(0xffffffffu, mn[a]);
    if (lane == 0) s_min[a][w] = r;
  }
#pragma unroll
  for (int a = 0; a < 6; a++) {
    const int r = __reduce_max_sync(0xffffffffu, mxv[a]);
    if (lane == 0) s_min[6 + a][w] = r;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    int r = s_min[threadIdx.x][0];
    for (int k = 1; k < kBlock / 32; k++) r = min(r, s_min[threadIdx.x][k]);
    if (r != 0x7FFFFFFF) atomicMin(&st->cam_min[threadIdx.x / 3][threadIdx.x % 3], r);
  } else if (threadIdx.x < 12) {
    const int a = threadIdx.x - 6;
    int r = s_min[threadIdx.x][0];
    for (int k = 1; k < kBlock / 32; k++) r = max(r, s_min[threadIdx.x][k]);
    if (r != int(0x80000000)) atomicMax(&st->cam_max[a / 3][a % 3], r);
  }
  if (bitmap_words == 0) return;  // key-sort path: no lattice to lay out
  // the last CTA to finish sees every minimum / maximum and lays out the occupancy bitmap
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&st->classify_done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last || threadIdx.x != 0) return;
  __threadfence();
  unsigned long long words = 0;
  long long cols = 0;
  bool fail = false;
  for (int c = 0; c < 2; c++) {
    int dim[3] = {0, 0, 0};
    const volatile int* vmin = st->cam_min[c];
    const volatile int* vmax = st->cam_max[c];
    for (int a = 0; a < 3; a++) {
      const double lo = double(ordered_to_float(vmin[a])), hi = double(ordered_to_float(vmax[a]));
      if (hi >= lo) {  // largest key of this axis, the very expression k_mark evaluates (localization.cpp:289,357-362)
        const double kmax = floor(__ddiv_rn(__dsub_rn(hi, lo), cell));
        const int bits = a == 0 ? kb.bx : a == 1 ? kb.by : kb.bz;
        if (kmax >= double((1u << bits) - 1u)) fail = true;  // outside the key range: the key-sort path reports it
        dim[a] = kmax < 2.0e9 ? int(kmax) + 1 : 0x7FFFFFF0;
      }
    }
    const int nzw = (dim[2] + 31) / 32;
    st->bnx[c] = dim[0];
    st->bny[c] = dim[1];
    st->bnzw[c] = nzw;
    st->bword0[c] = words;
    st->bcol0[c] = int(cols);
    const double wc = double(dim[0]) * double(dim[1]) * double(nzw);
    if (wc > 4.0e18) fail = true;
    else words += (unsigned long long)(dim[0]) * (unsigned long long)(dim[1]) * (unsigned long long)(nzw);
    cols += static_cast<long long>(dim[0]) * dim[1] + 1;
    if (words > bitmap_words || cols > static_cast<long long>(col_cap)) fail = true;
  }
  st->bwords = fail ? 0ull : words;
  st->bitmap_fail = fail ? 1 : 0;
}

// one atomic OR per kept point: bit (camera, kx, ky, kz) of the occupancy bitmap, k = floor((p - min)/cell) in
// binary64 exactly as the reference computes its voxel keys (localization.cpp:289,357-362)
__global__ void __launch_bounds__(kBlock)
k_mark(const char* pts, int stride, int n, const uint8_t* flag, double cell, const PreState* st, uint32_t* bitmap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || st->bitmap_fail) return;
  const uint8_t f = flag[i];
  if (f == 0) return;
  float x, y, z;
  load_xyz(pts, stride, i, x, y, z);
  const int c = f - 1;
  const double mx = double(ordered_to_float(st->cam_min[c][0]));
  const double my = double(ordered_to_float(st->cam_min[c][1]));
  const double mz = double(ordered_to_float(st->cam_min[c][2]));
  const unsigned long long kx = (unsigned long long)(floor(__ddiv_rn(__dsub_rn(double(x), mx), cell)));
  const unsigned long long ky = (unsigned long long)(floor(__ddiv_rn(__dsub_rn(double(y), my), cell)));
  const unsigned kz = unsigned(floor(__ddiv_rn(__dsub_rn(double(z), mz), cell)));
  const unsigned long long word = st->bword0[c] + (kx * (unsigned long long)(st->bny[c]) + ky) * (unsigned long long)(st->bnzw[c]) + (kz >> 5);
  atomicOr(bitmap + word, 1u << (kz & 31u));
}

// tile states of the single-pass scan: epoch << 34 | status << 32 | value.  The epoch (one per call) makes
// entries of earlier calls read as "not yet published", so the array is never cleared.
constexpr unsigned long long kTileAggregate = 1ull, kTilePrefix = 2ull;
__device__ __forceinline__ unsigned long long tile_pack(unsigned epoch, unsigned long long status, unsigned value) {
  return ((unsigned long long)(epoch & 0x3FFFFFFFu) << 34) | (status << 32) | value;
}

// Set bits in address order -> voxel list.  Persistent CTAs take tiles of kTileWords words by ticket; per tile:
// 8 words per thread (read, then cleared for the next call), popcount, block scan, the tile's exclusive prefix by
// decoupled look-back over the published tile states, then every set bit becomes the voxel record at its rank.
// The prefix at the first word of a lattice column / x-row is that column's / row's first voxel: the (kx, ky)
// column table and the x-row table fall out of the same scan.  voxel corner = (float)(k*cell + min)
// (localization.cpp:318-351); cloud_normals_ of the voxel is zeroed (hand_search.cpp:13-14).
__global__ void __launch_bounds__(kBlock)
k_emit_bitmap(uint32_t* bitmap, unsigned long long* tile_state, double cell, PreState* st, GPoint* vox,
              double* normals, RowIndex* ri, int* row_ptr, int row_stride, int* col_ptr, DrawArgs draw) {
  __shared__ int s_warp[kBlock / 32];
  __shared__ unsigned s_tile, s_last;
  __shared__ unsigned s_base;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned epoch = st->epoch;  // advanced on the device by k_init_state: a replayed CUDA graph gets a new one
  if (st->bitmap_fail) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      atomicOr(&ri->error, kErrBitmapRetry);
      ri->n_points = 0;
    }
    return;
  }
  const unsigned long long words = st->bwords;
  const unsigned n_tiles = unsigned((words + kTileWords - 1) / kTileWords);
  const unsigned long long w0c[2] = {st->bword0[0], st->bword0[1]};
  const int nx[2] = {st->bnx[0], st->bnx[1]}, ny[2] = {st->bny[0], st->bny[1]}, nzw[2] = {st->bnzw[0], st->bnzw[1]};
  const int col0[2] = {st->bcol0[0], st->bcol0[1]};
  double mn[2][3];
  for (int c = 0; c < 2; c++)
    for (int a = 0; a < 3; a++) mn[c][a] = double(ordered_to_float(st->cam_min[c][a]));
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // the descriptor fields that do not depend on the scan
    ri->inv_cell = 1.0 / cell;
    ri->row_base[0] = 0;
    ri->row_base[1] = row_stride;
    ri->use_cols = 1;
    ri->cols_bad = 0;
    for (int c = 0; c < 2; c++) {
      for (int a = 0; a < 3; a++) ri->mn[c][a] = mn[c][a];
      ri->nx[c] = nx[c];
      ri->ny[c] = ny[c];
      ri->col_base[c] = col0[c];
    }
    if (st->key_overflow) atomicOr(&ri->error, kErrKeyOverflow);
    if (words == 0) {  // nothing survived the filters
      ri->n_points = 0;
      st->n_vox = 0;
      for (int c = 0; c < 2; c++) ri->first[c] = ri->count[c] = 0;
    }
  }
  // tiles are taken by ticket, so a tile only exists once a RUNNING CTA owns it: every tile a look-back waits for is
  // being worked on, whatever else shares the GPU (other lanes of a batch, the side stream) — a static
  // assignment would need all CTAs of the launch to be co-resident
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_tile = atomicAdd(&st->tile_ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    if (tile >= n_tiles) {
      // the last CTA to leave closes the descriptor from the sentinels of the column table and, when asked to, draws
      // the samples (the voxel count is known right here; used to be a launch of its own)
      __threadfence();
      if (threadIdx.x == 0) s_last = atomicAdd(&st->emit_done, 1u) == gridDim.x - 1 ? 1u : 0u;  // (not s_tile: slower
                                                                                              // warps still read it)
      __syncthreads();
      if (!s_last) return;
      if (threadIdx.x == 0) {
        int n_points = 0;
        if (words > 0) {
          __threadfence();
          const volatile int* cp = col_ptr;
          const bool has0 = nx[0] > 0 && ny[0] > 0 && nzw[0] > 0, has1 = nx[1] > 0 && ny[1] > 0 && nzw[1] > 0;
          const int end0 = has0 ? cp[col0[0] + nx[0] * ny[0]] : 0;
          const int end1 = has1 ? cp[col0[1] + nx[1] * ny[1]] : end0;
          ri->first[0] = 0;
          ri->count[0] = end0;
          ri->first[1] = end0;
          ri->count[1] = end1 - end0;
          ri->n_points = end1;
          st->n_vox = end1;
          n_points = end1;
        }
        s_base = unsigned(n_points);
        if (draw.out) ri->n_samples = draw_count(draw, n_points);
      }
      __syncthreads();
      if (draw.out) {
        const int n_points = int(s_base);
        for (int j = threadIdx.x; j < draw.count; j += blockDim.x) draw.out[j] = draw_sample(draw, n_points, j);
      }
      return;
    }
    const unsigned long long wbase = (unsigned long long)(tile) * kTileWords + (unsigned long long)(threadIdx.x) * 8ull;
    uint32_t wd[8];
    if (wbase + 8 <= words) {
      const uint4 a = *reinterpret_cast<const uint4*>(bitmap + wbase), b = *reinterpret_cast<const uint4*>(bitmap + wbase + 4);
      wd[0] = a.x; wd[1] = a.y; wd[2] = a.z; wd[3] = a.w; wd[4] = b.x; wd[5] = b.y; wd[6] = b.z; wd[7] = b.w;
      *reinterpret_cast<uint4*>(bitmap + wbase) = make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(bitmap + wbase + 4) = make_uint4(0u, 0u, 0u, 0u);
    } else {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        wd[k] = wbase + k < words ? bitmap[wbase + k] : 0u;
        if (wbase + k < words) bitmap[wbase + k] = 0u;
      }
    }
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) cnt += __popc(wd[k]);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    int before = 0, tile_total = 0;
#pragma unroll
    for (int k = 0; k < kBlock / 32; k++) {
      before += k < w ? s_warp[k] : 0;
      tile_total += s_warp[k];
    }
    // decoupled look-back (one warp): publish the aggregate, walk the predecessors until an inclusive prefix
    if (w == 0) {
      unsigned base = 0;
      if (tile == 0) {
        if (lane == 0) atomicExch(&tile_state[0], tile_pack(epoch, kTilePrefix, unsigned(tile_total)));
      } else {
        if (lane == 0) atomicExch(&tile_state[tile], tile_pack(epoch, kTileAggregate, unsigned(tile_total)));
        int look = int(tile) - 1;
        for (;;) {  // lanes inspect tiles look, look - 1, ..., look - 31 (nearest first)
          const int t = look - lane;
          unsigned long long sv = 0ull;
          bool ready;
          do {
            ready = true;
            if (t >= 0) {
              sv = *reinterpret_cast<volatile unsigned long long*>(&tile_state[t]);
              ready = (sv >> 34) == (unsigned long long)(epoch & 0x3FFFFFFFu) && ((sv >> 32) & 3ull) != 0ull;
            }
          } while (!__all_sync(0xffffffffu, ready));
          const bool is_prefix = t >= 0 && ((sv >> 32) & 3ull) == kTilePrefix;
          const unsigned pm = __ballot_sync(0xffffffffu, is_prefix);
          // aggregates of the tiles in front of the nearest inclusive prefix, plus that prefix
          const int stop = pm ? __ffs(pm) - 1 : 31;
          const unsigned v = (t >= 0 && lane <= stop) ? unsigned(sv & 0xFFFFFFFFull) : 0u;
          base += __reduce_add_sync(0xffffffffu, v);
          if (pm) break;  // (tile 0 always publishes a prefix, so the walk ends)
          look -= 32;
        }
        if (lane == 0) atomicExch(&tile_state[tile], tile_pack(epoch, kTilePrefix, base + unsigned(tile_total)));
      }
      if (lane == 0) s_base = base;
    }
    __syncthreads();
    int run = int(s_base) + before + incl - cnt;  // rank of this thread's first set bit
    // lattice position of this thread's first word, then stepped word by word (no divisions in the loop)
    int c = 0;
    unsigned column = 0, kx = 0, ky = 0;
    int zw = 0;
    bool have = false;
#pragma unroll 1
    for (int k = 0; k < 8; k++) {
      const unsigned long long wi = wbase + k;
      if (wi >= words) break;
      const int c_now = (wi >= w0c[1] && nx[1] > 0 && ny[1] > 0 && nzw[1] > 0) ? 1 : 0;
      if (!have || c_now != c) {
        c = c_now;
        const unsigned rel0 = unsigned(wi - w0c[c]);  // (the bitmap has < 2^32 words)
        column = rel0 / unsigned(nzw[c]);
        zw = int(rel0 - column * unsigned(nzw[c]));
        kx = column / unsigned(ny[c]);
        ky = column - kx * unsigned(ny[c]);
        have = true;
      }
      if (zw == 0) {  // first word of a lattice column: the scan value is the column's first voxel
        col_ptr[col0[c] + int(column)] = run;
        if (ky == 0) row_ptr[c * row_stride + int(kx)] = run;
      }
      uint32_t bits = wd[k];
      if (bits) {
        const double px = __dadd_rn(__dmul_rn(double(kx), cell), mn[c][0]);  // two roundings, like the reference's SSE2 build
        const double py = __dadd_rn(__dmul_rn(double(ky), cell), mn[c][1]);
        while (bits) {
          const int b = __ffs(bits) - 1;
          bits &= bits - 1;
          GPoint p;
          p.x = float(px);
          p.y = float(py);
          p.z = float(__dadd_rn(__dmul_rn(double(zw * 32 + b), cell), mn[c][2]));
          p.tag = c ? kTagCamBit : 0u;
          vox[run] = p;
          normals[3 * size_t(run)] = 0.0;
          normals[3 * size_t(run) + 1] = 0.0;
          normals[3 * size_t(run) + 2] = 0.0;
          run++;
        }
      }
      // step to the next word of the lattice; the last word of a camera closes its tables
      if (++zw == nzw[c]) {
        zw = 0;
        column++;
        if (++ky == unsigned(ny[c])) {
          ky = 0;
          kx++;
          if (kx == unsigned(nx[c])) {
            col_ptr[col0[c] + nx[c] * ny[c]] = run;
            row_ptr[c * row_stride + nx[c]] = run;
            have = false;
          }
        }
      }
    }
  }
}

// key = cam<<63 | kx<<42 | ky<<21 | kz with k = floor((p - min)/cell) in binary64 (localization.cpp:289,357-362)
__global__ void k_keys(const char* pts, int stride, int n, const uint8_t* flag, double cell, PreState* st,
                       uint64_t* keys, KeyBits kb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t f = flag[i];
  if (f == 0) {
    keys[i] = kInvalidKey;
    return;
  }
  float x, y, z;
  load_xyz(pts, stride, i, x, y, z);
  int c = f - 1;
  double mx = double(ordered_to_float(st->cam_min[c][0]));
  double my = double(ordered_to_float(st->cam_min[c][1]));
  double mz = double(ordered_to_float(st->cam_min[c][2]));
  double kx = floor(__ddiv_rn(__dsub_rn(double(x), mx), cell));
  double ky = floor(__ddiv_rn(__dsub_rn(double(y), my), cell));
  double kz = floor(__ddiv_rn(__dsub_rn(double(z), mz), cell));
  if (kx >= double((1u << kb.bx) - 1u) || ky >= double((1u << kb.by) - 1u) || kz >= double((1u << kb.bz) - 1u)) {
    st->key_overflow = 1;
    keys[i] = kInvalidKey;
    return;
  }
  keys[i] = (uint64_t(c) << (kb.bx + kb.by + kb.bz)) | (uint64_t(kx) << (kb.by + kb.bz)) | (uint64_t(ky) << kb.bz) |
            uint64_t(kz);
}

// voxel corner = (float)(k*cell + min) (localization.cpp:318-351), camera-0 voxels first (key order);
// the same pass fills the per-camera x-row table and the RowIndex descriptor.
__global__ void k_emit(const uint64_t* keys_unique, int n_cap, double cell, PreState* st, GPoint* vox, KeyBits kb,
                       RowIndex* ri, int* row_ptr, int row_stride, int* col_ptr, int col_cap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int nu = min(st->n_unique, n_cap);
  if (i >= nu) return;
  const uint64_t key = keys_unique[i];
  const int n_vox = (keys_unique[nu - 1] == kInvalidKey) ? nu - 1 : nu;  // the invalid key sorts last
  if (i == 0) {
    st->n_vox = n_vox;
    ri->n_points = n_vox;
    ri->inv_cell = 1.0 / cell;
    ri->row_base[0] = 0;
    ri->row_base[1] = row_stride;
    for (int c = 0; c < 2; c++)
      for (int a = 0; a < 3; a++) ri->mn[c][a] = double(ordered_to_float(st->cam_min[c][a]));
    if (st->key_overflow) atomicOr(&ri->error, kErrKeyOverflow);
    if (n_vox == 0) {
      for (int c = 0; c < 2; c++) ri->nx[c] = ri->first[c] = ri->count[c] = 0;
    }
  }
  // extent of the dense (kx, ky) column table, from the per-camera coordinate ranges (every thread
  // derives the same numbers; thread 0 publishes them)
  int nyc[2], nxc[2], cbase[2];
  long long total = 0;
  for (int cc = 0; cc < 2; cc++) {
    const double x0 = double(ordered_to_float(st->cam_min[cc][0])), y0 = double(ordered_to_float(st->cam_min[cc][1]));
    const double x1 = double(ordered_to_float(st->cam_max[cc][0])), y1 = double(ordered_to_float(st->cam_max[cc][1]));
    const bool any = x1 >= x0;
    nxc[cc] = any ? int(floor((x1 - x0) / cell)) + 1 : 0;
    nyc[cc] = any ? int(floor((y1 - y0) / cell)) + 1 : 0;
    cbase[cc] = int(total);
    total += static_cast<long long>(nxc[cc]) * nyc[cc] + 1;
  }
  const bool use_cols = total <= static_cast<long long>(col_cap);
  if (i == 0) {
    ri->use_cols = use_cols ? 1 : 0;
    for (int cc = 0; cc < 2; cc++) {
      ri->ny[cc] = nyc[cc];
      ri->col_base[cc] = cbase[cc];
    }
  }
  if (key == kInvalidKey) return;
  const int sh_c = kb.bx + kb.by + kb.bz, sh_x = kb.by + kb.bz;
  const int c = int((key >> sh_c) & 1u);
  const int kxi = int((key >> sh_x) & ((1ull << kb.bx) - 1));
  const double kx = double(kxi), ky = double((key >> kb.bz) & ((1ull << kb.by) - 1)),
               kz = double(key & ((1ull << kb.bz) - 1));
  const double mx = double(ordered_to_float(st->cam_min[c][0]));
  const double my = double(ordered_to_float(st->cam_min[c][1]));
  const double mz = double(ordered_to_float(st->cam_min[c][2]));
  GPoint p;
  p.x = float(__dadd_rn(__dmul_rn(kx, cell), mx));  // two roundings, like the reference's SSE2 build
  p.y = float(__dadd_rn(__dmul_rn(ky, cell), my));
  p.z = float(__dadd_rn(__dmul_rn(kz, cell), mz));
  p.tag = c ? kTagCamBit : 0u;
  vox[i] = p;
  // x-row table: row_ptr[kx'] = i for every row kx' in (previous row, this row]
  int prev_c = -1, prev_kx = -1, prev_ky = -1;
  if (i > 0) {
    const uint64_t pk = keys_unique[i - 1];
    prev_c = int((pk >> sh_c) & 1u);
    prev_kx = int((pk >> sh_x) & ((1ull << kb.bx) - 1));
    prev_ky = int((pk >> kb.bz) & ((1ull << kb.by) - 1));
  }
  if (use_cols) {
    // column table: col[kx*ny + ky'] = i for every cell in (previous voxel's cell, this voxel's cell]
    const int kyi = int(ky);
    int* col = col_ptr + cbase[c];
    const long long cur = static_cast<long long>(kxi) * nyc[c] + kyi;
    const long long prv = (prev_c == c) ? static_cast<long long>(prev_kx) * nyc[c] + prev_ky : -1;
    constexpr long long kMaxGap = 8192;  // a single thread fills a gap inline; beyond this the table is abandoned
    if (cur - prv > kMaxGap) atomicOr(&ri->cols_bad, 1);
    else
      for (long long L = prv + 1; L <= cur; L++) col[L] = i;
    if (i == n_vox - 1 || int((keys_unique[i + 1] >> sh_c) & 1u) != c) {  // last voxel of this camera
      const long long endL = static_cast<long long>(nxc[c]) * nyc[c];
      if (endL - cur > kMaxGap) atomicOr(&ri->cols_bad, 1);
      else
        for (long long L = cur + 1; L <= endL; L++) col[L] = i + 1;
    }
  }
  int* rows = row_ptr + c * row_stride;
  if (prev_c != c) {
    for (int k = 0; k <= kxi; k++) rows[k] = i;
    ri->first[c] = i;
    if (prev_c == 0) {  // camera 0 ended at i-1
      ri->nx[0] = prev_kx + 1;
      ri->count[0] = i;
      row_ptr[prev_kx + 1] = i;
    } else if (c == 1) {  // no camera-0 voxels at all
      ri->nx[0] = 0;
      ri->count[0] = 0;
      ri->first[0] = 0;
    }
  } else if (prev_kx != kxi) {
    for (int k = prev_kx + 1; k <= kxi; k++) rows[k] = i;
  }
  if (i == n_vox - 1) {  // last voxel closes its camera's table
    rows[kxi + 1] = n_vox;
    ri->nx[c] = kxi + 1;
    int first_c = 0;
    if (c == 1) {  // first camera-1 key (the boundary thread writes ri->first[1]; do not race on it)
      int lo = 0, hi = n_vox;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((keys_unique[mid] >> sh_c) & 1u) hi = mid;
        else lo = mid + 1;
      }
      first_c = lo;
    }
    ri->count[c] = n_vox - first_c;
    if (c == 0) {
      ri->nx[1] = 0;
      ri->count[1] = 0;
      ri->first[1] = n_vox;
    }
  }
}

// cloud supplied directly (ag_set_cloud): the caller's order is kept as the API index space, and the
// x-row table is built over an x-sorted COPY of it?  No — the walkers need rows contiguous in the point
// array itself, so the supplied cloud must already be in voxel order (it is when it comes from
// ag_preprocess or from the oracle's preprocess, which is how every caller obtains it).  The kernel
// below derives the lattice rows from the coordinates and flags a cloud that is not row-sorted.
__global__ void k_index_cloud(const GPoint* vox, int n, double cell, RowIndex* ri, int* row_ptr, int row_stride,
                              int* sorted_ok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const GPoint p = vox[i];
  const int c = (p.tag & kTagCamBit) ? 1 : 0;
  const double mx = ri->mn[c][0];
  const int kxi = int(floor((double(p.x) - mx) / cell + 0.5));
  int prev_c = -1, prev_kx = -1;
  float prev_y = 0.f;
  if (i > 0) {
    const GPoint q = vox[i - 1];
    prev_c = (q.tag & kTagCamBit) ? 1 : 0;
    prev_kx = int(floor((double(q.x) - ri->mn[prev_c][0]) / cell + 0.5));
    prev_y = q.y;
  }
  if (prev_c > c || (prev_c == c && (prev_kx > kxi || (prev_kx == kxi && prev_y > p.y)))) *sorted_ok = 0;
  if (kxi < 0 || kxi >= row_stride - 1) {
    *sorted_ok = 0;
    return;
  }
  int* rows = row_ptr + c * row_stride;
  if (prev_c != c) {
    for (int k = 0; k <= kxi; k++) rows[k] = i;
    ri->first[c] = i;
    if (prev_c == 0) {
      ri->nx[0] = prev_kx + 1;
      ri->count[0] = i;
      row_ptr[prev_kx + 1] = i;
    } else if (c == 1) {
      ri->nx[0] = 0;
      ri->count[0] = 0;
      ri->first[0] = 0;
    }
  } else if (prev_kx != kxi) {
    for (int k = prev_kx + 1; k <= kxi; k++) rows[k] = i;
  }
  if (i == n - 1) {
    rows[kxi + 1] = n;
    ri->nx[c] = kxi + 1;
    ri->count[c] = n - ri->first[c];
    if (c == 0) {
      ri->nx[1] = 0;
      ri->count[1] = 0;
      ri->first[1] = n;
    }
  }
}

// per-camera minima of a supplied cloud (lattice origin of each camera)
__global__ void k_cloud_min(const GPoint* vox, int n, PreState* st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = i < n;
  GPoint p;
  p.x = p.y = p.z = 0.f;
  p.tag = 0;
  if (ok) p = vox[i];
  const int ov[3] = {float_to_ordered(p.x), float_to_ordered(p.y), float_to_ordered(p.z)};
  for (int c = 0; c < 2; c++) {
    const bool mine = ok && ((p.tag & kTagCamBit) ? 1 : 0) == c;
    for (int a = 0; a < 3; a++) {
      const int r = __reduce_min_sync(0xffffffffu, mine ? ov[a] : 0x7FFFFFFF);
      if ((threadIdx.x & 31) == 0 && r != 0x7FFFFFFF) atomicMin(&st->cam_min[c][a], r);
    }
  }
}
__global__ void k_cloud_desc(PreState* st, RowIndex* ri, int n, double cell, int row_stride) {
  ri->n_points = n;
  ri->inv_cell = 1.0 / cell;
  ri->row_base[0] = 0;
  ri->row_base[1] = row_stride;
  for (int c = 0; c < 2; c++) {
    for (int a = 0; a < 3; a++) ri->mn[c][a] = double(ordered_to_float(st->cam_min[c][a]));
    ri->nx[c] = ri->first[c] = ri->count[c] = 0;
  }
  st->n_vox = n;
}

__global__ void k_radius_search(const GPoint* __restrict__ pts, const int* __restrict__ row_ptr,
                                const int* __restrict__ col_ptr, const RowIndex* rip,
                                float qx, float qy, float qz, float r2, double rpad, int* out, int* out_count, int cap) {
  __shared__ int s_rs[256], s_pre[257];
  const RowIndex ri = *rip;
  int row_off = 0;
  bool more = true;
  while (more) {
    const int nr = build_runs_warp(ri, row_ptr, col_ptr, pts, qx, qy, rpad, s_rs, s_pre, 256, row_off, more);
    row_off += nr;
    for (int r = 0; r < nr; r++)
      for (int j = s_rs[r] + int(threadIdx.x); j < s_rs[r] + (s_pre[r + 1] - s_pre[r]); j += 32) {
        const GPoint p = pts[j];
        if (dist2_flann(qx, qy, qz, p.x, p.y, p.z) < r2) {
          const int k = atomicAdd(out_count, 1);
          if (k < cap) out[k] = j;
        }
      }
    __syncwarp();
  }
}

// uses_clustering (localization.cpp:51-98): RANSAC score of T candidate planes in one pass over the voxel cloud —
// counts[t] = number of points with |n_t . p + d_t| < thresh (pcl::SampleConsensusModelPlane::countWithinDistance)
__global__ void __launch_bounds__(256)
k_plane_score(const GPoint* __restrict__ vox, int n, const double* __restrict__ planes, int T, double thresh,
              int* __restrict__ counts) {
  __shared__ double s_pl[128][4];
  __shared__ int s_cnt[128];
  for (int i = threadIdx.x; i < T * 4; i += blockDim.x) (&s_pl[0][0])[i] = planes[i];
  for (int i = threadIdx.x; i < T; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {
    const int i = base + threadIdx.x;
    GPoint p;
    p.x = p.y = p.z = 0.f;
    p.tag = 0;
    if (i < n) p = vox[i];
    const double x = double(p.x), y = double(p.y), z = double(p.z);
    for (int t = 0; t < T; t++) {
      const double d = (s_pl[t][0] * x + s_pl[t][1] * y) + (s_pl[t][2] * z + s_pl[t][3]);
      const unsigned m = __ballot_sync(0xffffffffu, i < n && fabs(d) < thresh);
      if (lane == 0 && m) atomicAdd(&s_cnt[t], __popc(m));
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x)
    if (s_cnt[t]) atomicAdd(&counts[t], s_cnt[t]);
}

}  // namespace

static PreState* state_ptr(Ctx* c) { return c->misc.as<PreState>(); }

// ---- uses_clustering: dominant plane removal -------------------------------------------------------------------
// PCL (not vendored) draws its RANSAC samples from its own generator, so which triples are drawn is this library's
// choice (splitmix64 of the sample seed); everything else follows pcl::SACSegmentation as configured at
// localization.cpp:56-68: 100 iterations, inlier = |distance| < 0.01, first best count wins, coefficients refitted
// to the winner's inliers (centroid + smallest eigenvector of their covariance), the refitted plane's inliers
// removed.  The 100 x N scoring runs on the GPU; the 3x3 refit and the compaction are host work on the 1-2 MB
// voxel cloud (training-time path).
namespace {
uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
bool inlier_of(const double pl[4], const GPoint& p, double thresh) {
  const double d = (pl[0] * double(p.x) + pl[1] * double(p.y)) + (pl[2] * double(p.z) + pl[3]);
  return std::fabs(d) < thresh;
}
}  // namespace

int remove_plane_device(Ctx* c) {
  constexpr int kIter = 100;
  constexpr double kThresh = 0.01;
  RowIndex hri;
  AG_CUDA_CHECK(cudaMemcpyAsync(&hri, c->row_index.p, sizeof(hri), cudaMemcpyDeviceToHost, c->stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (hri.error & kErrBitmapRetry) return AG_RETRY_KEYSORT;
  const int n = hri.n_points;
  if (n < 3) {  // no plane can be estimated: the reference returns no hands (localization.cpp:69-74)
    set_error(" Could not estimate a planar model for the given dataset.");
    return AG_ERR_EMPTY;
  }
  std::vector<GPoint> v(n);
  AG_CUDA_CHECK(cudaMemcpyAsync(v.data(), c->vox.p, size_t(n) * 16, cudaMemcpyDeviceToHost, c->stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  // candidate planes: three distinct points per iteration
  std::vector<double> planes(size_t(kIter) * 4, 0.0);
  std::vector<char> valid(kIter, 0);
  for (int t = 0; t < kIter; t++) {
    int idx[3] = {0, 0, 0};
    uint64_t st = mix64(c->params.seed ^ mix64(uint64_t(t) + 0x1234567ull));
    for (int k = 0; k < 3; k++)
      for (;;) {
        st = mix64(st);
        const int cand = int(st % uint64_t(n));
        bool dup = false;
        for (int q = 0; q < k; q++) dup = dup || idx[q] == cand;
        if (!dup) {
          idx[k] = cand;
          break;
        }
      }
    const GPoint &a = v[idx[0]], &b = v[idx[1]], &d = v[idx[2]];
    const double e1[3] = {double(b.x) - double(a.x), double(b.y) - double(a.y), double(b.z) - double(a.z)};
    const double e2[3] = {double(d.x) - double(a.x), double(d.y) - double(a.y), double(d.z) - double(a.z)};
    const double nx = e1[1] * e2[2] - e1[2] * e2[1], ny = e1[2] * e2[0] - e1[0] * e2[2], nz = e1[0] * e2[1] - e1[1] * e2[0];
    const double len = std::sqrt(nx * nx + ny * ny + nz * nz);
    double* pl = &planes[size_t(t) * 4];
    if (len > 1e-12) {
      valid[t] = 1;
      pl[0] = nx / len;
      pl[1] = ny / len;
      pl[2] = nz / len;
      pl[3] = -1.0 * (pl[0] * double(a.x) + pl[1] * double(a.y) + pl[2] * double(a.z));
    } else {
      pl[3] = 1e30;  // (scores zero)
    }
  }
  DevBuf dpl;
  if (dpl.reserve(planes.size() * 8 + kIter * 4 + 64)) return AG_ERR_CUDA;
  int* d_counts = reinterpret_cast<int*>(dpl.as<double>() + planes.size());
  AG_CUDA_CHECK(cudaMemcpyAsync(dpl.p, planes.data(), planes.size() * 8, cudaMemcpyHostToDevice, c->stream));
  AG_CUDA_CHECK(cudaMemsetAsync(d_counts, 0, kIter * 4, c->stream));
  k_plane_score<<<kNumSMs * 2, 256, 0, c->stream>>>(c->vox.as<GPoint>(), n, dpl.as<double>(), kIter, kThresh, d_counts);
  c->launches += 1;
  int counts[kIter];
  AG_CUDA_CHECK(cudaMemcpyAsync(counts, d_counts, kIter * 4, cudaMemcpyDeviceToHost, c->stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  dpl.release();
  int best = -1, best_count = 0;
  for (int t = 0; t < kIter; t++)
    if (valid[t] && counts[t] > best_count) {
      best_count = counts[t];
      best = t;
    }
  if (best < 0) {  // localization.cpp:69-74
    set_error(" Could not estimate a planar model for the given dataset.");
    return AG_ERR_EMPTY;
  }
  // refit to the winner's inliers (index order), then the refitted plane's inliers go
  double pl[4] = {planes[size_t(best) * 4], planes[size_t(best) * 4 + 1], planes[size_t(best) * 4 + 2], planes[size_t(best) * 4 + 3]};
  {
    double ctr[3] = {0, 0, 0};
    long long m = 0;
    for (int i = 0; i < n; i++)
      if (inlier_of(pl, v[i], kThresh)) {
        ctr[0] += double(v[i].x);
        ctr[1] += double(v[i].y);
        ctr[2] += double(v[i].z);
        m++;
      }
    if (m >= 3) {
      for (int d = 0; d < 3; d++) ctr[d] /= double(m);
      double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
      for (int i = 0; i < n; i++)
        if (inlier_of(pl, v[i], kThresh)) {
          const double w[3] = {double(v[i].x) - ctr[0], double(v[i].y) - ctr[1], double(v[i].z) - ctr[2]};
          for (int r = 0; r < 3; r++)
            for (int q = 0; q < 3; q++) C[r][q] += w[r] * w[q];
        }
      for (int sweep = 0; sweep < 60; sweep++) {  // cyclic Jacobi
        if (C[0][1] * C[0][1] + C[0][2] * C[0][2] + C[1][2] * C[1][2] <= 1e-40) break;
        for (int p = 0; p < 2; p++)
          for (int q = p + 1; q < 3; q++) {
            if (C[p][q] == 0.0) continue;
            const double theta = (C[q][q] - C[p][p]) / (2.0 * C[p][q]);
            const double tt = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
            const double cs = 1.0 / std::sqrt(tt * tt + 1.0), sn = tt * cs;
            for (int k = 0; k < 3; k++) {
              const double akp = C[k][p], akq = C[k][q];
              C[k][p] = cs * akp - sn * akq;
              C[k][q] = sn * akp + cs * akq;
            }
            for (int k = 0; k < 3; k++) {
              const double apk = C[p][k], aqk = C[q][k];
              C[p][k] = cs * apk - sn * aqk;
              C[q][k] = sn * apk + cs * aqk;
            }
            for (int k = 0; k < 3; k++) {
              const double vkp = V[k][p], vkq = V[k][q];
              V[k][p] = cs * vkp - sn * vkq;
              V[k][q] = sn * vkp + cs * vkq;
            }
          }
      }
      int mi = 0;
      if (C[1][1] < C[mi][mi]) mi = 1;
      if (C[2][2] < C[mi][mi]) mi = 2;
      const double len = std::sqrt(V[0][mi] * V[0][mi] + V[1][mi] * V[1][mi] + V[2][mi] * V[2][mi]);
      for (int d = 0; d < 3; d++) pl[d] = V[d][mi] / len;
      pl[3] = -1.0 * (pl[0] * ctr[0] + pl[1] * ctr[1] + pl[2] * ctr[2]);
    }
  }
  size_t w = 0;
  bool cam1 = false;
  for (int i = 0; i < n; i++)
    if (!inlier_of(pl, v[i], kThresh)) {
      v[w] = v[i];
      v[w].tag &= kTagCamBit;
      cam1 = cam1 || (v[w].tag & kTagCamBit);
      w++;
    }
  c->two_cams = c->two_cams && cam1;
  if (w > 0) AG_CUDA_CHECK(cudaMemcpyAsync(c->vox.p, v.data(), w * 16, cudaMemcpyHostToDevice, c->stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  return set_cloud_device(c, int(w));  // re-index the remaining (still voxel-ordered) cloud
}

static int row_stride_for(const ag_params& P, int* bx_out) {
  // rows per camera: workspace extent / voxel (+2), rounded up to the key bit budget
  const double cells = floor((P.workspace[1] - P.workspace[0]) / P.voxel_size) + 2.0;
  int bits = 1;
  while (bits < 21 && double(1u << bits) - 1.0 <= cells) bits++;
  if (bx_out) *bx_out = bits;
  return (1 << bits) + 2;
}

int preprocess_device(Ctx* c, const void* d_points, int stride, int n_in, int size_left) {
  const ag_params& P = c->params;
  const int nb = (n_in + kBlock - 1) / kBlock;
  KeyBits kb;
  {
    int* b[3] = {&kb.bx, &kb.by, &kb.bz};
    for (int a = 0; a < 3; a++) {
      const double cells = floor((P.workspace[2 * a + 1] - P.workspace[2 * a]) / P.voxel_size) + 2.0;
      int bits = 1;
      while (bits < 21 && double(1u << bits) - 1.0 <= cells) bits++;
      *b[a] = bits;
    }
    kb.total = 1 + kb.bx + kb.by + kb.bz;
  }
  const int row_stride = (1 << kb.bx) + 2;
  const bool fresh_state = c->misc.cap < sizeof(PreState);
  if (c->misc.reserve(sizeof(PreState)) || c->block_counts.reserve(size_t(nb) * 4 + size_t(n_in)) ||
      c->vox.reserve(size_t(n_in) * 16) || c->row_ptr.reserve(size_t(row_stride) * 2 * 4) ||
      c->col_ptr.reserve(size_t(kColCap) * 4) ||
      c->row_index.reserve(sizeof(RowIndex)) || c->normals.reserve(size_t(n_in) * 24))
    return AG_ERR_CUDA;
  if (fresh_state) AG_CUDA_CHECK(cudaMemsetAsync(c->misc.p, 0, sizeof(PreState), c->stream));  // (the scan epoch starts at 0)
  const bool bitmap = c->bitmap_ok;
  if (bitmap && !c->bitmap.p) {  // once per context: the occupancy bitmap (kept clean by its reader) and the tile states
    if (c->bitmap.reserve(kBitmapBytes) || c->tile_state.reserve((kBitmapBytes / 4 / kTileWords + 1) * 8)) return AG_ERR_CUDA;
    AG_CUDA_CHECK(cudaMemsetAsync(c->bitmap.p, 0, c->bitmap.cap, c->stream));
    AG_CUDA_CHECK(cudaMemsetAsync(c->tile_state.p, 0, c->tile_state.cap, c->stream));
  }
  if (!bitmap && (c->keys.reserve(size_t(n_in) * 8) || c->keys_sorted.reserve(size_t(n_in) * 8) ||
                  c->keys_unique.reserve(size_t(n_in) * 8)))
    return AG_ERR_CUDA;
  c->n_cap = n_in;
  PreState* st = state_ptr(c);
  int* d_block = c->block_counts.as<int>();
  uint8_t* d_flag = reinterpret_cast<uint8_t*>(d_block + nb);
  const char* pts = static_cast<const char*>(d_points);
  ResetList rl;
  std::memset(&rl, 0, sizeof(rl));
  rl.p[0] = c->row_index.as<unsigned>();
  rl.words[0] = int(sizeof(RowIndex) / 4);
  if ((c->fold_resets & 1u) && c->counters.p) {
    rl.p[1] = c->counters.as<unsigned>();
    rl.words[1] = 16;
  } else {
    c->fold_resets &= ~1u;
  }
  if ((c->fold_resets & 2u) && c->overflow.p) {
    rl.p[2] = c->overflow.as<unsigned>();
    rl.words[2] = 1;
  } else {
    c->fold_resets &= ~2u;
  }
  if ((c->fold_resets & 4u) && c->rand_carry.p) {
    rl.p[3] = c->rand_carry.as<unsigned>();
    rl.words[3] = 4;
  }
  c->fold_resets &= ~4u;  // (quadric_rand_reset ran before this call)
  if ((c->fold_resets & 8u) && c->overflow_list.p) {
    rl.p[4] = c->overflow_list.as<unsigned>();
    rl.words[4] = 1;
  } else {
    c->fold_resets &= ~8u;
  }
  k_init_state<<<1, 64, 0, c->stream>>>(st, rl);
  const bool quirk = !P.fix_cam_source && size_left < n_in;
  if (quirk) {
    k_count_finite<<<nb, kBlock, 0, c->stream>>>(pts, stride, n_in, d_block);
    k_scan_blocks<<<1, 1024, 0, c->stream>>>(d_block, nb);
  }
  k_classify<<<std::min(nb, kNumSMs * 4), kBlock, 0, c->stream>>>(
      pts, stride, n_in, size_left, quirk ? d_block : nullptr, P.workspace[0], P.workspace[1], P.workspace[2],
      P.workspace[3], P.workspace[4], P.workspace[5], d_flag, st, P.voxel_size,
      bitmap ? (unsigned long long)(kBitmapBytes / 4) : 0ull, kColCap, kb);
  DrawArgs draw;
  std::memset(&draw, 0, sizeof(draw));
  if (c->fold_draw) draw = *static_cast<const DrawArgs*>(c->fold_draw);
  c->draw_folded = false;
  if (bitmap) {
    k_mark<<<nb, kBlock, 0, c->stream>>>(pts, stride, n_in, d_flag, P.voxel_size, st, c->bitmap.as<uint32_t>());
    k_emit_bitmap<<<kNumSMs * 4, kBlock, 0, c->stream>>>(c->bitmap.as<uint32_t>(), c->tile_state.as<unsigned long long>(),
                                                         P.voxel_size, st, c->vox.as<GPoint>(), c->normals.as<double>(),
                                                         c->row_index.as<RowIndex>(), c->row_ptr.as<int>(), row_stride,
                                                         c->col_ptr.as<int>(), draw);
    c->draw_folded = draw.out != nullptr;
    c->launches += quirk ? 6 : 4;  // init, [count, scan], classify, mark, emit
    AG_CUDA_CHECK(cudaGetLastError());
    return AG_OK;
  }
  k_keys<<<nb, kBlock, 0, c->stream>>>(pts, stride, n_in, d_flag, P.voxel_size, st, c->keys.as<uint64_t>(), kb);
  size_t tmp1 = 0, tmp2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp1, c->keys.as<uint64_t>(), c->keys_sorted.as<uint64_t>(), n_in, 0,
                                 kb.total, c->stream);
  cub::DeviceSelect::Unique(nullptr, tmp2, c->keys_sorted.as<uint64_t>(), c->keys_unique.as<uint64_t>(),
                            &st->n_unique, n_in, c->stream);
  if (c->cub_tmp.reserve(tmp1 > tmp2 ? tmp1 : tmp2)) return AG_ERR_CUDA;
  AG_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(c->cub_tmp.p, tmp1, c->keys.as<uint64_t>(),
                                               c->keys_sorted.as<uint64_t>(), n_in, 0, kb.total, c->stream));
  AG_CUDA_CHECK(cub::DeviceSelect::Unique(c->cub_tmp.p, tmp2, c->keys_sorted.as<uint64_t>(),
                                          c->keys_unique.as<uint64_t>(), &st->n_unique, n_in, c->stream));
  c->launches += quirk ? 6 : 4;  // init, [count, scan], classify, keys, emit (CUB kernels not counted)
  k_emit<<<nb, kBlock, 0, c->stream>>>(c->keys_unique.as<uint64_t>(), n_in, P.voxel_size, st, c->vox.as<GPoint>(), kb,
                                       c->row_index.as<RowIndex>(), c->row_ptr.as<int>(), row_stride,
                                       c->col_ptr.as<int>(), kColCap);
  // cloud_normals_ is zeroed on every call (hand_search.cpp:13-14)
  AG_CUDA_CHECK(cudaMemsetAsync(c->normals.p, 0, size_t(n_in) * 24, c->stream));
  AG_CUDA_CHECK(cudaGetLastError());
  return AG_OK;
}

// reads the voxel count (and device-side error flags) back: the one place that waits for the GPU
int fetch_cloud_size(Ctx* c) {
  RowIndex hri;
  AG_CUDA_CHECK(cudaMemcpyAsync(&hri, c->row_index.p, sizeof(hri), cudaMemcpyDeviceToHost, c->stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  c->n_vox = hri.n_points;
  if (hri.error & kErrBitmapRetry) return AG_RETRY_KEYSORT;
  if (hri.error & kErrKeyOverflow) {
    set_error("voxel index exceeds the key range (workspace extent / voxel_size > 2^21 cells)");
    return AG_ERR_CAPACITY;
  }
  return AG_OK;
}

// cloud supplied directly (already voxelised, in voxel order): c->vox holds n records
int set_cloud_device(Ctx* c, int n) {
  const ag_params& P = c->params;
  int bx = 0;
  int row_stride = row_stride_for(P, &bx);
  if (c->misc.reserve(sizeof(PreState) + 16) || c->row_ptr.reserve(size_t(row_stride) * 2 * 4) ||
      c->col_ptr.reserve(64) ||
      c->row_index.reserve(sizeof(RowIndex)) || c->normals.reserve(std::max<size_t>(24, size_t(n) * 24)))
    return AG_ERR_CUDA;
  c->n_cap = n;
  c->n_vox = n;
  PreState* st = state_ptr(c);
  int* d_ok = reinterpret_cast<int*>(st + 1);
  const int one = 1;
  AG_CUDA_CHECK(cudaMemsetAsync(c->row_index.p, 0, sizeof(RowIndex), c->stream));
  AG_CUDA_CHECK(cudaMemcpyAsync(d_ok, &one, 4, cudaMemcpyHostToDevice, c->stream));
  {
    ResetList none;
    std::memset(&none, 0, sizeof(none));
    k_init_state<<<1, 32, 0, c->stream>>>(st, none);
  }
  if (n > 0) {
    const int nb = (n + kBlock - 1) / kBlock;
    k_cloud_min<<<nb, kBlock, 0, c->stream>>>(c->vox.as<GPoint>(), n, st);
    k_cloud_desc<<<1, 1, 0, c->stream>>>(st, c->row_index.as<RowIndex>(), n, P.voxel_size, row_stride);
    k_index_cloud<<<nb, kBlock, 0, c->stream>>>(c->vox.as<GPoint>(), n, P.voxel_size, c->row_index.as<RowIndex>(),
                                                c->row_ptr.as<int>(), row_stride, d_ok);
  } else {
    k_cloud_desc<<<1, 1, 0, c->stream>>>(st, c->row_index.as<RowIndex>(), 0, P.voxel_size, row_stride);
  }
  AG_CUDA_CHECK(cudaMemsetAsync(c->normals.p, 0, std::max<size_t>(24, size_t(n) * 24), c->stream));
  int ok = 1;
  AG_CUDA_CHECK(cudaMemcpyAsync(&ok, d_ok, 4, cudaMemcpyDeviceToHost, c->stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  AG_CUDA_CHECK(cudaGetLastError());
  if (!ok) {
    set_error("ag_set_cloud: the cloud is not in voxel order (per camera sorted by x, then y, on the voxel lattice, "
              "inside the workspace): pass a cloud produced by ag_preprocess");
    return AG_ERR_INVALID;
  }
  return AG_OK;
}

int radius_search_device(Ctx* c, const float q[3], double radius, std::vector<int>& out) {
  out.clear();
  if (c->n_vox <= 0) return AG_OK;
  const int cap = c->n_vox;
  DevBuf buf;
  if (buf.reserve(size_t(cap + 1) * 4)) return AG_ERR_CUDA;
  int* d_out = buf.as<int>();
  int* d_cnt = d_out + cap;
  cudaMemsetAsync(d_cnt, 0, 4, c->stream);
  float r2 = float(radius * radius);
  double rpad = sqrt(double(r2)) * (1.0 + 1e-5) + 1e-7;
  k_radius_search<<<1, 32, 0, c->stream>>>(c->vox.as<GPoint>(), c->row_ptr.as<int>(), c->col_ptr.as<int>(),
                                           c->row_index.as<RowIndex>(), q[0],
                                           q[1], q[2], r2, rpad, d_out, d_cnt, cap);
  int cnt = 0;
  cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (e == cudaSuccess) {
    out.resize(cnt);
    e = cudaMemcpy(out.data(), d_out, size_t(cnt) * 4, cudaMemcpyDeviceToHost);
  }
  buf.release();
  if (e != cudaSuccess) {
    set_error(std::string("radius search: ") + cudaGetErrorString(e));
    return AG_ERR_CUDA;
  }
  return AG_OK;
}

}  // namespace ag
