// preprocess.cu — NaN removal, workspace filter, voxelisation and hash-grid build on the GPU.
//
// Replaces (reference paths): pcl::removeNaNFromPointCloud + camera labelling
// (src/agile_grasp/localization.cpp:17-27), Localization::filterWorkspace (:216-245),
// Localization::voxelizeCloud (:247-355; std::set<Vector3i> -> 64-bit key radix sort + unique) and
// pcl::KdTreeFLANN::setInputCloud (src/agile_grasp/hand_search.cpp:10-11; kd-tree -> uniform hash
// grid with z-contiguous cell runs).  All of it is HBM-streaming integer/byte work: one coalesced
// pass per kernel, CUB only for the device-wide radix sort / scan / select primitives.

#include <cub/cub.cuh>

#include "ag_internal.h"

namespace ag {

namespace {

constexpr int kBlock = 256;
constexpr uint64_t kInvalidKey = ~0ull;

// voxel key = cam | kx | ky | kz packed with per-axis bit widths derived from the workspace extent, so
// the radix sort only touches the bits that can be set
struct KeyBits {
  int bx, by, bz;  // bits per axis (<= 21)
  int total;       // 1 + bx + by + bz
};

struct PreState {
  int cam_min[2][3];  // ordered-int encoded float minima per camera
  int bb_min[3];      // bounding box of the voxelised cloud (ordered ints)
  int bb_max[3];
  int n_unique;       // output of DeviceSelect::Unique
  int n_vox;
  int key_overflow;
};

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

__global__ void k_init_state(PreState* st) {
  if (threadIdx.x == 0) {
    for (int c = 0; c < 2; c++)
      for (int a = 0; a < 3; a++) st->cam_min[c][a] = float_to_ordered(10000.0f);  // localization.cpp:251-252
    for (int a = 0; a < 3; a++) {
      st->bb_min[a] = 0x7FFFFFFF;
      st->bb_max[a] = int(0x80000000);
    }
    st->n_unique = 0;
    st->n_vox = 0;
    st->key_overflow = 0;
  }
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

// flag[i] = 0 dropped, 1 camera 0, 2 camera 1; per-camera coordinate minima
__global__ void k_classify(const char* pts, int stride, int n, int size_left, const int* block_offsets,
                           double w0, double w1, double w2, double w3, double w4, double w5, uint8_t* flag,
                           PreState* st) {
  __shared__ int wcount[kBlock / 32];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float x = 0, y = 0, z = 0;
  bool fin = i < n && load_xyz(pts, stride, i, x, y, z);
  int label;
  if (block_offsets) {
    // localization.cpp:19-27: labels are assigned by position BEFORE NaN removal but read AFTER
    // compaction, i.e. finite point #k gets label (k < size_left ? 0 : 1)   
    unsigned m = __ballot_sync(0xffffffffu, fin);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wcount[w] = __popc(m);
    __syncthreads();
    int before = block_offsets[blockIdx.x];
    for (int k = 0; k < w; k++) before += wcount[k];
    int rank = before + __popc(m & ((1u << lane) - 1));
    label = rank < size_left ? 0 : 1;
  } else {
    label = i < size_left ? 0 : 1;
  }
  // localization.cpp:228-229 inclusive workspace box, float promoted to double
  bool keep = fin && double(x) >= w0 && double(x) <= w1 && double(y) >= w2 && double(y) <= w3 && double(z) >= w4 &&
              double(z) <= w5;
  if (i < n) flag[i] = keep ? uint8_t(1 + label) : uint8_t(0);
  // minima (localization.cpp:256-277): warp reduce -> shared -> one atomic per block, camera and axis
  __shared__ int s_min[6][kBlock / 32];
  {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ov[3] = {float_to_ordered(x), float_to_ordered(y), float_to_ordered(z)};
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const bool mine = keep && label == c;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const int r = __reduce_min_sync(0xffffffffu, mine ? ov[a] : 0x7FFFFFFF);
        if (lane == 0) s_min[c * 3 + a][w] = r;
      }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      int r = s_min[threadIdx.x][0];
      for (int k = 1; k < kBlock / 32; k++) r = min(r, s_min[threadIdx.x][k]);
      if (r != 0x7FFFFFFF) atomicMin(&st->cam_min[threadIdx.x / 3][threadIdx.x % 3], r);
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

// voxel corner = (float)(k*cell + min) (localization.cpp:318-351), camera-0 voxels first (key order)
__global__ void k_emit(const uint64_t* keys_unique, int n_cap, double cell, PreState* st, float4* vox, KeyBits kb) {
  __shared__ int s_red[6][kBlock / 32];
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int nu = st->n_unique;
  bool ok = i < nu && i < n_cap;
  uint64_t key = ok ? keys_unique[i] : kInvalidKey;
  ok = ok && key != kInvalidKey;
  if (i == nu - 1) st->n_vox = (key == kInvalidKey) ? nu - 1 : nu;
  float x = 0, y = 0, z = 0;
  if (ok) {
    int c = int((key >> (kb.bx + kb.by + kb.bz)) & 1u);
    double kx = double((key >> (kb.by + kb.bz)) & ((1ull << kb.bx) - 1)), ky = double((key >> kb.bz) & ((1ull << kb.by) - 1)),
           kz = double(key & ((1ull << kb.bz) - 1));
    double mx = double(ordered_to_float(st->cam_min[c][0]));
    double my = double(ordered_to_float(st->cam_min[c][1]));
    double mz = double(ordered_to_float(st->cam_min[c][2]));
    x = float(__dadd_rn(__dmul_rn(kx, cell), mx));  // two roundings, like the reference's SSE2 build
    y = float(__dadd_rn(__dmul_rn(ky, cell), my));
    z = float(__dadd_rn(__dmul_rn(kz, cell), mz));
    vox[i] = make_float4(x, y, z, __int_as_float(c));
  }
  // bounding box for the hash grid: warp reduce -> shared -> one atomic per block and bound
  int v[3] = {float_to_ordered(x), float_to_ordered(y), float_to_ordered(z)};
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const int lo = __reduce_min_sync(0xffffffffu, ok ? v[a] : 0x7FFFFFFF);
    const int hi = __reduce_max_sync(0xffffffffu, ok ? v[a] : int(0x80000000));
    if (lane == 0) {
      s_red[a][w] = lo;
      s_red[3 + a][w] = hi;
    }
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    const int a = threadIdx.x;
    int r = s_red[a][0];
    for (int k = 1; k < kBlock / 32; k++) r = a < 3 ? min(r, s_red[a][k]) : max(r, s_red[a][k]);
    if (a < 3) {
      if (r != 0x7FFFFFFF) atomicMin(&st->bb_min[a], r);
    } else {
      if (r != int(0x80000000)) atomicMax(&st->bb_max[a - 3], r);
    }
  }
}

__global__ void k_bbox(const float4* vox, int n, PreState* st) {  // for ag_set_cloud (cloud given directly)
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = i < n;
  float4 p = ok ? vox[i] : make_float4(0, 0, 0, 0);
  int v[3] = {float_to_ordered(p.x), float_to_ordered(p.y), float_to_ordered(p.z)};
#pragma unroll
  for (int a = 0; a < 3; a++) {
    int lo = __reduce_min_sync(0xffffffffu, ok ? v[a] : 0x7FFFFFFF);
    int hi = __reduce_max_sync(0xffffffffu, ok ? v[a] : int(0x80000000));
    if ((threadIdx.x & 31) == 0) {
      atomicMin(&st->bb_min[a], lo);
      atomicMax(&st->bb_max[a], hi);
    }
  }
}

// ---- hash grid ------------------------------------------------------------------------------
__global__ void k_cell_ids(const float4* vox, int n, GridDesc g, uint32_t* cell_ids, int* perm, int* cell_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = vox[i];
  int cx = cell_of(double(p.x), g.gmin[0], g.inv_cell, g.dim[0]);
  int cy = cell_of(double(p.y), g.gmin[1], g.inv_cell, g.dim[1]);
  int cz = cell_of(double(p.z), g.gmin[2], g.inv_cell, g.dim[2]);
  int lin = (cx * g.dim[1] + cy) * g.dim[2] + cz;
  cell_ids[i] = uint32_t(lin);
  perm[i] = i;
  atomicAdd(&cell_count[lin], 1);
}

__global__ void k_gather(const float4* vox, const int* perm_sorted, int n, GPoint* pts, int* inv) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  int src = perm_sorted[j];
  float4 p = vox[src];
  GPoint q;
  q.x = p.x; q.y = p.y; q.z = p.z;
  q.tag = uint32_t(src) | (__float_as_int(p.w) ? kTagCamBit : 0u);
  pts[j] = q;
  inv[src] = j;
}

}  // namespace

static PreState* state_ptr(Ctx* c) { return c->misc.as<PreState>(); }

static int finish_cloud(Ctx* c, bool have_bbox_on_device) {
  // read N and the bounding box back (the only host sync of the preprocessing stage)
  PreState hs;
  AG_CUDA_CHECK(cudaMemcpyAsync(&hs, state_ptr(c), sizeof(hs), cudaMemcpyDeviceToHost, c->stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  if (hs.key_overflow) {
    set_error("voxel index exceeds the key range (workspace extent / voxel_size > 2^21 cells, or a point lies "
              "outside the workspace box it passed)");
    return AG_ERR_CAPACITY;
  }
  (void)have_bbox_on_device;
  c->n_vox = hs.n_vox;
  if (c->n_vox <= 0) {
    c->n_vox = 0;
    return AG_OK;
  }
  GridDesc& g = c->grid;
  double lo[3], hi[3];
  for (int a = 0; a < 3; a++) {
    lo[a] = double(ordered_to_float(hs.bb_min[a]));
    hi[a] = double(ordered_to_float(hs.bb_max[a]));
  }
  // cell edge ~ the Taubin radius: a radius-r query touches <=3 cells per axis, the r=0.08 hand
  // query <=7.  Grow the cell until the dense cell table stays below 16M entries.
  double cell = c->params.nn_radius_taubin * 1.001;
  if (!(cell > 1e-6)) cell = 0.03;
  for (;;) {
    double total = 1;
    for (int a = 0; a < 3; a++) {
      g.dim[a] = int(floor((hi[a] - lo[a]) / cell)) + 1;
      total *= double(g.dim[a]);
    }
    if (total <= double(1 << 24)) break;
    cell *= 1.26;
  }
  for (int a = 0; a < 3; a++) g.gmin[a] = lo[a];
  g.inv_cell = 1.0 / cell;
  g.n_points = c->n_vox;
  return AG_OK;
}

int preprocess_device(Ctx* c, const void* d_points, int stride, int n_in, int size_left) {
  const ag_params& P = c->params;
  const int nb = (n_in + kBlock - 1) / kBlock;
  if (c->misc.reserve(sizeof(PreState)) || c->keys.reserve(size_t(n_in) * 8) || c->keys_sorted.reserve(size_t(n_in) * 8) ||
      c->keys_unique.reserve(size_t(n_in) * 8) || c->block_counts.reserve(size_t(nb) * 4 + size_t(n_in)) ||
      c->vox.reserve(size_t(n_in) * 16))
    return AG_ERR_CUDA;
  PreState* st = state_ptr(c);
  int* d_block = c->block_counts.as<int>();
  uint8_t* d_flag = reinterpret_cast<uint8_t*>(d_block + nb);
  const char* pts = static_cast<const char*>(d_points);
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
  k_init_state<<<1, 32, 0, c->stream>>>(st);
  const bool quirk = !P.fix_cam_source && size_left < n_in;
  if (quirk) {
    k_count_finite<<<nb, kBlock, 0, c->stream>>>(pts, stride, n_in, d_block);
    k_scan_blocks<<<1, 1024, 0, c->stream>>>(d_block, nb);
  }
  k_classify<<<nb, kBlock, 0, c->stream>>>(pts, stride, n_in, size_left, quirk ? d_block : nullptr, P.workspace[0],
                                           P.workspace[1], P.workspace[2], P.workspace[3], P.workspace[4],
                                           P.workspace[5], d_flag, st);
  k_keys<<<nb, kBlock, 0, c->stream>>>(pts, stride, n_in, d_flag, P.voxel_size, st, c->keys.as<uint64_t>(), kb);
  size_t tmp1 = 0, tmp2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, tmp1, c->keys.as<uint64_t>(), c->keys_sorted.as<uint64_t>(), n_in, 0,
                                 kb.total, c->stream);
  cub::DeviceSelect::Unique(nullptr, tmp2, c->keys_sorted.as<uint64_t>(), c->keys_unique.as<uint64_t>(),
                            &st->n_unique, n_in, c->stream);
  size_t tmp = tmp1 > tmp2 ? tmp1 : tmp2;
  if (c->cub_tmp.reserve(tmp)) return AG_ERR_CUDA;
  AG_CUDA_CHECK(cub::DeviceRadixSort::SortKeys(c->cub_tmp.p, tmp1, c->keys.as<uint64_t>(),
                                               c->keys_sorted.as<uint64_t>(), n_in, 0, kb.total, c->stream));
  AG_CUDA_CHECK(cub::DeviceSelect::Unique(c->cub_tmp.p, tmp2, c->keys_sorted.as<uint64_t>(),
                                          c->keys_unique.as<uint64_t>(), &st->n_unique, n_in, c->stream));
  c->launches += quirk ? 6 : 4;  // init, [count, scan], classify, keys, emit (CUB kernels not counted)
  k_emit<<<nb, kBlock, 0, c->stream>>>(c->keys_unique.as<uint64_t>(), n_in, P.voxel_size, st, c->vox.as<float4>(), kb);
  AG_CUDA_CHECK(cudaGetLastError());
  return finish_cloud(c, true);
}

// cloud supplied directly (already voxelised): c->vox holds n float4 records
int set_cloud_device(Ctx* c, int n) {
  if (c->misc.reserve(sizeof(PreState))) return AG_ERR_CUDA;
  PreState* st = state_ptr(c);
  k_init_state<<<1, 32, 0, c->stream>>>(st);
  if (n > 0) k_bbox<<<(n + kBlock - 1) / kBlock, kBlock, 0, c->stream>>>(c->vox.as<float4>(), n, st);
  PreState tmp;
  (void)tmp;
  // n_vox is known on the host here
  int* d_nvox = &st->n_vox;
  AG_CUDA_CHECK(cudaMemcpyAsync(d_nvox, &n, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  AG_CUDA_CHECK(cudaGetLastError());
  return finish_cloud(c, true);
}

int build_grid(Ctx* c) {
  const int n = c->n_vox;
  if (n <= 0) return AG_OK;
  GridDesc& g = c->grid;
  const size_t ncells = size_t(g.dim[0]) * g.dim[1] * g.dim[2];
  if (c->cell_ids.reserve(size_t(n) * 4) || c->cell_ids_sorted.reserve(size_t(n) * 4) || c->perm.reserve(size_t(n) * 4) ||
      c->perm_sorted.reserve(size_t(n) * 4) || c->cell_start.reserve((ncells + 1) * 2 * sizeof(int)) ||
      c->pts.reserve(size_t(n) * sizeof(GPoint)) || c->inv.reserve(size_t(n) * 4))
    return AG_ERR_CUDA;
  int* cell_count = c->cell_start.as<int>() + (ncells + 1);
  int* cell_start = c->cell_start.as<int>();
  AG_CUDA_CHECK(cudaMemsetAsync(cell_count, 0, (ncells + 1) * sizeof(int), c->stream));
  const int nb = (n + kBlock - 1) / kBlock;
  k_cell_ids<<<nb, kBlock, 0, c->stream>>>(c->vox.as<float4>(), n, g, c->cell_ids.as<uint32_t>(), c->perm.as<int>(),
                                           cell_count);
  int bits = 1;
  while ((size_t(1) << bits) < ncells) bits++;
  size_t tmp1 = 0, tmp2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp1, cell_count, cell_start, int(ncells + 1), c->stream);
  cub::DeviceRadixSort::SortPairs(nullptr, tmp2, c->cell_ids.as<uint32_t>(), c->cell_ids_sorted.as<uint32_t>(),
                                  c->perm.as<int>(), c->perm_sorted.as<int>(), n, 0, bits, c->stream);
  if (c->cub_tmp.reserve(tmp1 > tmp2 ? tmp1 : tmp2)) return AG_ERR_CUDA;
  AG_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp1, cell_count, cell_start, int(ncells + 1), c->stream));
  AG_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp2, c->cell_ids.as<uint32_t>(),
                                                c->cell_ids_sorted.as<uint32_t>(), c->perm.as<int>(),
                                                c->perm_sorted.as<int>(), n, 0, bits, c->stream));
  c->launches += 2;  // cell_ids, gather
  k_gather<<<nb, kBlock, 0, c->stream>>>(c->vox.as<float4>(), c->perm_sorted.as<int>(), n, c->pts.as<GPoint>(),
                                         c->inv.as<int>());
  // cloud_normals_ is zeroed on every call (hand_search.cpp:13-14)
  if (c->normals.reserve(size_t(n) * 3 * sizeof(double))) return AG_ERR_CUDA;
  AG_CUDA_CHECK(cudaMemsetAsync(c->normals.p, 0, size_t(n) * 3 * sizeof(double), c->stream));
  AG_CUDA_CHECK(cudaGetLastError());
  return AG_OK;
}

// ---- brute radius search through the grid (stage-level API / tests) ---------------------------
namespace {
__global__ void k_radius_search(const GPoint* __restrict__ pts, const int* __restrict__ cell_start, GridDesc g,
                                float qx, float qy, float qz, float r2, double rpad, int* out, int* out_count, int cap) {
  QueryBox b = query_box(g, qx, qy, qz, rpad);
  int ncx = b.hi[0] - b.lo[0] + 1, ncy = b.hi[1] - b.lo[1] + 1;
  for (int col = blockIdx.x; col < ncx * ncy; col += gridDim.x) {
    int cx = b.lo[0] + col / ncy, cy = b.lo[1] + col % ncy;
    int s = cell_start[cell_linear(g, cx, cy, b.lo[2])];
    int e = cell_start[cell_linear(g, cx, cy, b.hi[2]) + 1];
    for (int j = s + threadIdx.x; j < e; j += blockDim.x) {
      GPoint p = pts[j];
      if (dist2_flann(qx, qy, qz, p.x, p.y, p.z) < r2) {
        int k = atomicAdd(out_count, 1);
        if (k < cap) out[k] = int(p.tag & kTagIndexMask);
      }
    }
  }
}
}  // namespace

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
  k_radius_search<<<64, 128, 0, c->stream>>>(c->pts.as<GPoint>(), c->cell_start.as<int>(), c->grid, q[0], q[1], q[2], r2,
                                             rpad, d_out, d_cnt, cap);
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
