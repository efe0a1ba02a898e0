// api.cu — the extern "C" ABI declared in include/ag_b200.h: context, parameters, SVM model file
// loader, the full localize/classify path and the stage-level entry points.

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "ag_internal.h"

namespace ag {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

// bumped whenever any buffer a captured graph may point to is (re)allocated: a cached graph is only replayed
// while the generation it was captured under is current
std::atomic<unsigned long long> g_alloc_gen{0};

// Every list of hypotheses is stamped (ag_grasp.reserved) with a non-zero byte derived from the context and its
// call counter, so that ag_classify can tell records of the resident grasp images from records of an earlier
// call, another context or a batch lane (their image_id would silently address unrelated images).
static std::atomic<unsigned> g_ctx_serial{0};
uint8_t next_stamp(Ctx* c) {
  c->call_gen++;
  c->stamp = uint8_t(1u + (c->serial * 151u + c->call_gen) % 255u);
  return c->stamp;
}

int DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return 0;
  g_alloc_gen++;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e));
    p = nullptr;
    return AG_ERR_CUDA;
  }
  cap = want;
  return 0;
}
void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

int ctx_pinned(Ctx* c, size_t bytes) {
  if (bytes <= c->h_pinned_cap) return 0;
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  c->h_pinned = nullptr;
  c->h_pinned_cap = 0;
  size_t want = bytes + bytes / 4 + 4096;
  cudaError_t e = cudaMallocHost(&c->h_pinned, want);
  if (e != cudaSuccess) {
    set_error(std::string("cudaMallocHost failed: ") + cudaGetErrorString(e));
    return AG_ERR_CUDA;
  }
  c->h_pinned_cap = want;
  return 0;
}


// ---- device-side sample handling (the draw itself: ag_common.cuh; ag_localize folds it into the voxel scan) -----
__global__ void k_draw_samples(RowIndex* ri, DrawArgs d, int set_count = 1) {
  const int n = ri->n_points;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0 && set_count) ri->n_samples = draw_count(d, n);
  if (j < d.count) d.out[j] = draw_sample(d, n, j);
}
__global__ void k_check_samples(RowIndex* ri, int s_req, const int* idx) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k == 0) ri->n_samples = s_req;
  if (k < s_req && (idx[k] < 0 || idx[k] >= ri->n_points)) atomicOr(&ri->error, kErrBadIndex);
}
__global__ void k_iota(int* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

__global__ void k_gather_grasps(const ag_grasp* __restrict__ raw, const int* __restrict__ slots, int n, ag_grasp* out,
                                uint8_t stamp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ag_grasp gr = raw[slots[i]];
  gr.image_id = i;
  gr.reserved = stamp;
  out[i] = gr;
}

// results header + records, written straight into pinned device-mapped host memory
struct HostOut {
  int n_hyp, n_vox, n_samples, error, n_over, pad[3];
  unsigned long long counters[4];
};
// Gathers the surviving hypotheses in (sample, orientation) order (= the reference's stable concat,
// hand_search.cpp:194-200) straight from the per-slot records, merges the SVM results and writes the
// list to mapped host memory, to the device copy used by ag_classify / ag_get_*, and to the optional
// caller-registered device buffer.
// Peer gather (ag_gather_*): slot `rank` of every rank's gather buffer, reached over NVLink through CUDA IPC
// mappings.  Slot layout: [uint32 epoch flag, int32 first, step, pad][int32 n_hyp, n_vox, n_samples, error][records][masks].
struct PeerOut {
  char* slot[AG_MAX_GATHER_RANKS];
  const unsigned* ack;  // this rank's ack words (one per consumer, kAckStride bytes apart): last epoch it has consumed
  int world;         // 0 = peer gather not configured
  int cap;           // records per slot
  unsigned epoch;
  unsigned* done;    // CTA completion counter of this launch (device, zeroed by the last CTA)
  int final_pass;    // 0: first export of a call (publishes only if no sample needs the large-slab re-run)
  // per-sample masks of the valid orientations (one byte per local sample, behind the records of the slot): with
  // them a consumer places every record of a sharded call by a prefix sum instead of searching the other lists
  size_t mask_off;   // byte offset of the mask area inside a slot
  int mask_cap;      // bytes of the mask area
  const uint8_t* valid;  // [local samples x 8] valid flags of this call
  int n_local;       // local samples (launch bound)
  int first, step;   // position of local sample j in the full sample list: first + j * step
};
constexpr int kAckStride = 64;                                    // bytes between the ack words of two consumers
constexpr int kAckBytes = AG_MAX_GATHER_RANKS * kAckStride;       // ack area in front of the slots of a gather buffer
constexpr int kSlotHeaderBytes = 32;

__global__ void k_export(const ag_grasp* __restrict__ raw, const int* __restrict__ slots, const int* __restrict__ n_sel,
                         const float* __restrict__ scores, const RowIndex* ri, const int* overflow,
                         const unsigned long long* counters, HostOut* hdr, ag_grasp* out_host, ag_grasp* out_dev,
                         int* exp_hdr, ag_grasp* out_exp, int cap, int cap_exp, PeerOut peer, unsigned stamp,
                         int slot_first, int slot_step, int scores_by_slot) {
  // a record is 10 x 16 bytes: three records per warp pass, every lane moves one uint4 (coalesced
  // 480-byte stores — the mapped host destination is written over PCIe and needs full-width writes)
  static_assert(sizeof(ag_grasp) == 160, "record layout");
  const int n = min(*n_sel, cap);
  const int lane = threadIdx.x & 31;
  if (peer.world > 0) {
    // back-pressure: the slot about to be written held epoch - 2; every consumer must have acknowledged it
    // (k_gather_merge stores the epoch it has finished reading into this rank's ack words)
    if (threadIdx.x < peer.world && peer.epoch > 2u) {
      const volatile unsigned* a = reinterpret_cast<const volatile unsigned*>(
          reinterpret_cast<const char*>(peer.ack) + size_t(threadIdx.x) * kAckStride);
      const long long t0 = clock64();
      while (int(*a - (peer.epoch - 2u)) < 0 && clock64() - t0 < 4000000000ll) __nanosleep(100);
    }
    __syncthreads();
  }
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  const int sub = lane / 10, part = lane % 10;
  for (int base = warp_global * 3; base < n; base += n_warps * 3) {
    const int i = base + sub;
    if (lane >= 30 || i >= n) continue;
    uint4 v = reinterpret_cast<const uint4*>(raw + slots[i])[part];
    const float sc = scores ? scores[scores_by_slot ? slots[i] : i] : 0.f;
    if (part == 8) {
      if (scores) v.x = __float_as_uint(sc);                               // byte 128: score
      v.z = uint32_t(slot_first + int(v.z) * slot_step);                   // byte 136: sample_slot, position in the FULL sample list
    }
    if (part == 9) {
      v.z = uint32_t(i);                                                   // byte 152: image_id
      v.w = (v.w & 0x00FFFFFFu) | (stamp << 24);                           // byte 159: call stamp
      if (scores) {                                                        // byte 158: label (+1 <=> sum <= 0)
        const uint32_t label = sc > 0.f ? 0u : 1u;
        v.w = (v.w & 0xFF00FFFFu) | (label << 16);
      }
    }
    reinterpret_cast<uint4*>(out_host + i)[part] = v;
    reinterpret_cast<uint4*>(out_dev + i)[part] = v;
    if (out_exp && i < cap_exp) reinterpret_cast<uint4*>(out_exp + i)[part] = v;
    if (i < peer.cap)  // fused all-gather: the record goes straight into every rank's buffer over NVLink
      for (int r = 0; r < peer.world; r++)
        reinterpret_cast<uint4*>(peer.slot[r] + kSlotHeaderBytes + size_t(i) * sizeof(ag_grasp))[part] = v;
  }
  if (peer.world > 0 && peer.valid) {
    // valid-orientation mask of every local sample -> every rank's slot (16 masks per 16-byte store)
    const int n_s = min(ri->n_samples, min(peer.n_local, peer.mask_cap));
    for (int j0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16; j0 < n_s; j0 += gridDim.x * blockDim.x * 16) {
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      for (int t = 0; t < 16 && j0 + t < n_s; t++) {
        const uint2 v = *reinterpret_cast<const uint2*>(peer.valid + size_t(j0 + t) * 8);
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
          m |= ((v.x >> (8 * k)) & 0xFFu ? 1u : 0u) << k;
          m |= ((v.y >> (8 * k)) & 0xFFu ? 1u : 0u) << (4 + k);
        }
        w[t >> 2] |= m << (8 * (t & 3));
      }
      for (int r = 0; r < peer.world; r++)
        *reinterpret_cast<uint4*>(peer.slot[r] + peer.mask_off + j0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
  if (peer.world > 0) {
    // publish: every CTA fences its peer stores; the last one to finish writes the headers, fences again and
    // raises the epoch flags the consumers (k_gather_merge on each rank) spin on
    __shared__ bool s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(peer.done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
      if (threadIdx.x < peer.world && (peer.final_pass || overflow[0] == 0)) {
        int* h0 = reinterpret_cast<int*>(peer.slot[threadIdx.x]);
        h0[1] = peer.first;
        h0[2] = peer.step;
        int* h4 = reinterpret_cast<int*>(peer.slot[threadIdx.x] + 16);
        h4[0] = min(n, peer.cap);
        h4[1] = ri->n_points;
        h4[2] = ri->n_samples;
        h4[3] = ri->error | (n > peer.cap ? 0x100 : 0);
        __threadfence_system();
        *reinterpret_cast<volatile unsigned*>(peer.slot[threadIdx.x]) = peer.epoch;
      }
      if (threadIdx.x == 0) *peer.done = 0u;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    hdr->n_hyp = n;
    hdr->n_vox = ri->n_points;
    hdr->n_samples = ri->n_samples;
    hdr->error = ri->error;
    hdr->n_over = overflow[0];
    for (int k = 0; k < 4; k++) hdr->counters[k] = counters[k];
    if (exp_hdr) {
      exp_hdr[0] = min(n, cap_exp);
      exp_hdr[1] = ri->n_points;
      exp_hdr[2] = ri->n_samples;
      exp_hdr[3] = ri->error;
    }
  }
}

// Consumer side of the peer gather, enqueued right behind the export kernel of the same call: waits until every
// rank's list of this epoch has landed in this rank's buffer, merges the lists into ONE list in the reference's
// order (sample-major, orientation-minor: hand_search.cpp:194-200) — whatever the sample assignment was,
// contiguous or interleaved: the position of a record is the number of records with a smaller
// (sample position, orientation) key over all lists, found by one binary search per list — writes it to device
// and mapped host memory, and acknowledges the epoch to every producer (its slot may then be overwritten).
struct GatherHost {  // mapped host header of the merged list
  int n_total, status, world, pad;
  int n_per_rank[AG_MAX_GATHER_RANKS];
  int err_per_rank[AG_MAX_GATHER_RANKS];
};
struct MergeArgs {
  const char* buf;         // this epoch's slots of this rank's gather buffer
  size_t slot_bytes, mask_off;
  unsigned* ack_peer[AG_MAX_GATHER_RANKS];  // producer p's ack word for this consumer
  int world;
  unsigned epoch;
  const int* overflow;     // [0] != 0: this rank's own list is re-exported after a host-side re-run -> skip this pass
  int final_pass;
  ag_grasp* merged_dev;
  GatherHost* host;
  int cap_total;
  unsigned* done;
  int* plan;               // device: [0] mode (1 = shards of one call), [1] total, [2] skip, [4 + r] n_r, [12 + r] first_r,
                           // [20 + r] offset of list r in rank-major order, [28 + r] err_r; [64 ...] base[k] of every sample
  int plan_cap;            // entries available for base[]
};
constexpr int kPlanBase = 64;

// Consumer side of the peer gather, kernel 1 (one CTA): waits until every rank's list of this epoch has landed,
// reads the headers and lays out the merged list.  Shards of one call (every rank reports the same stride = world
// and its own offset): the merged list is in the reference's order (sample-major, orientation-minor:
// hand_search.cpp:194-200), so base[k] = number of hypotheses of the samples in front of sample k — a prefix sum
// over the valid-orientation masks that came with the lists.  Independent calls (weak scaling): rank-major.
__global__ void __launch_bounds__(1024)
k_gather_plan(const MergeArgs A) {
  __shared__ int s_n[AG_MAX_GATHER_RANKS], s_err[AG_MAX_GATHER_RANKS], s_first[AG_MAX_GATHER_RANKS], s_step[AG_MAX_GATHER_RANKS],
      s_ns[AG_MAX_GATHER_RANKS];
  __shared__ int s_bad, s_warp[32], s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (!A.final_pass && A.overflow[0] != 0) {
    if (tid == 0) A.plan[2] = 1;  // skip: the re-export after the host-side re-run brings the final list
    return;
  }
  if (tid == 0) {
    s_bad = 0;
    s_carry = 0;
    A.plan[2] = 0;
  }
  __syncthreads();
  if (tid < A.world) {
    const int r = tid;
    const volatile unsigned* flag = reinterpret_cast<const volatile unsigned*>(A.buf + size_t(r) * A.slot_bytes);
    const long long t0 = clock64();
    bool ok = true;
    while (*flag != A.epoch) {  // (a producer cannot run ahead: it waits for this consumer's acknowledgement)
      if (clock64() - t0 > 4000000000ll) {
        ok = false;
        break;
      }
      __nanosleep(64);
    }
    __threadfence_system();
    const volatile int* h0 = reinterpret_cast<const volatile int*>(A.buf + size_t(r) * A.slot_bytes);
    s_first[r] = ok ? h0[1] : 0;
    s_step[r] = ok ? h0[2] : 0;
    s_n[r] = ok ? h0[4] : 0;
    s_ns[r] = ok ? h0[6] : 0;
    s_err[r] = ok ? h0[7] : 0x200;
    if (!ok) atomicOr(&s_bad, 1);
  }
  __syncthreads();
  // shards of one call?  every rank strides by `world` and the offsets are a permutation of 0 .. world - 1
  bool sharded = A.world > 1;
  unsigned seen = 0;
  int total = 0, n_samples_all = 0;
  for (int r = 0; r < A.world; r++) {
    sharded = sharded && s_step[r] == A.world && s_first[r] >= 0 && s_first[r] < A.world;
    if (s_first[r] >= 0 && s_first[r] < 32) seen |= 1u << s_first[r];
    total += s_n[r];
    if (s_ns[r] > 0) n_samples_all = max(n_samples_all, s_first[r] + (s_ns[r] - 1) * A.world + 1);
  }
  sharded = sharded && seen == (A.world >= 32 ? 0xFFFFFFFFu : (1u << A.world) - 1u) && n_samples_all <= A.plan_cap;
  if (tid == 0) {
    A.plan[0] = sharded ? 1 : 0;
    A.plan[1] = total;
    int off = 0;
    for (int r = 0; r < A.world; r++) {
      A.plan[4 + r] = s_n[r];
      A.plan[12 + r] = s_first[r];
      A.plan[20 + r] = off;
      A.plan[28 + r] = s_err[r];
      off += s_n[r];
    }
    A.host->n_total = min(total, A.cap_total);
    A.host->world = A.world;
    A.host->status = s_bad | (total > A.cap_total ? 2 : 0);
    for (int r = 0; r < A.world; r++) {
      A.host->n_per_rank[r] = s_n[r];
      A.host->err_per_rank[r] = s_err[r];
    }
  }
  if (!sharded) return;
  // rank owning sample k: the one whose offset is k mod world
  __shared__ int s_owner[AG_MAX_GATHER_RANKS];
  if (tid < A.world) s_owner[s_first[tid]] = tid;
  __syncthreads();
  for (int base = 0; base < n_samples_all; base += 1024) {
    const int k = base + tid;
    int cnt = 0;
    if (k < n_samples_all) {
      const int r = s_owner[k % A.world], j = k / A.world;
      cnt = j < s_ns[r] ? __popc(unsigned(*reinterpret_cast<const volatile uint8_t*>(A.buf + size_t(r) * A.slot_bytes + A.mask_off + j))) : 0;
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    if (k < n_samples_all) A.plan[kPlanBase + k] = s_carry + (warp > 0 ? s_warp[warp - 1] : 0) + incl - cnt;
    __syncthreads();
    if (tid == 0) s_carry += s_warp[31];
    __syncthreads();
  }
}

// kernel 2: every record goes to its place in the merged list (device memory), then the epoch is acknowledged to
// every producer (its slot may be overwritten from now on)
__global__ void __launch_bounds__(256)
k_gather_place(const MergeArgs A) {
  __shared__ bool s_last;
  if (A.plan[2]) return;
  const int sharded = A.plan[0], total = A.plan[1];
  const int lane = threadIdx.x & 31, sub = lane / 10, part = lane % 10;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int base = warp_global * 3; base < total; base += n_warps * 3) {
    const int item = base + sub;
    if (lane >= 30 || item >= total) continue;
    int r = 0;
    while (r + 1 < A.world && item >= A.plan[20 + r + 1]) r++;
    const int i = item - A.plan[20 + r];
    const char* slot = A.buf + size_t(r) * A.slot_bytes;
    const ag_grasp* mine = reinterpret_cast<const ag_grasp*>(slot + kSlotHeaderBytes) + i;
    int pos = item;  // rank-major
    if (sharded) {   // sample-major: hypotheses of earlier samples, then the valid orientations below this one
      const int k = mine->sample_slot, j = k / A.world, o = mine->orientation & 7;
      const unsigned mask = *reinterpret_cast<const uint8_t*>(slot + A.mask_off + j);
      pos = A.plan[kPlanBase + k] + __popc(mask & ((1u << o) - 1u));
    }
    if (pos < A.cap_total) {
      uint4 v = reinterpret_cast<const uint4*>(mine)[part];
      if (part == 9) v.z = uint32_t(-1);  // image_id: the grasp images stay on the producing rank
      reinterpret_cast<uint4*>(A.merged_dev + pos)[part] = v;  // (the host copy is made when ag_gather_result asks for it)
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(A.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x < A.world)  // every list of this epoch has been read: the producers may reuse the slots
    *reinterpret_cast<volatile unsigned*>(A.ack_peer[threadIdx.x]) = A.epoch;
  if (threadIdx.x == 0) *A.done = 0u;
}
// publishes an empty list (a call without samples still takes part in the exchange)
__global__ void k_publish_empty(PeerOut peer, const RowIndex* ri) {
  if (threadIdx.x < peer.world) {
    if (peer.epoch > 2u) {
      const volatile unsigned* a = reinterpret_cast<const volatile unsigned*>(
          reinterpret_cast<const char*>(peer.ack) + size_t(threadIdx.x) * kAckStride);
      const long long t0 = clock64();
      while (int(*a - (peer.epoch - 2u)) < 0 && clock64() - t0 < 4000000000ll) __nanosleep(100);
    }
    int* h0 = reinterpret_cast<int*>(peer.slot[threadIdx.x]);
    h0[1] = peer.first;
    h0[2] = peer.step;
    int* h4 = reinterpret_cast<int*>(peer.slot[threadIdx.x] + 16);
    h4[0] = 0;
    h4[1] = ri->n_points;
    h4[2] = 0;
    h4[3] = ri->error;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned*>(peer.slot[threadIdx.x]) = peer.epoch;
  }
}

// ---- SVM model file (OpenCV 2.4 YAML, svm_032015_linear_20_20_same:1-16,780-789) -------------
static bool read_list(const std::string& text, size_t& pos, std::vector<double>& out) {
  const size_t lb = text.find('[', pos);
  if (lb == std::string::npos) return false;
  const size_t rb = text.find(']', lb);
  if (rb == std::string::npos) return false;
  const char* p = text.c_str() + lb + 1;
  const char* end = text.c_str() + rb;
  while (p < end) {
    while (p < end && (*p == ' ' || *p == ',' || *p == '\n' || *p == '\r' || *p == '\t')) p++;
    if (p >= end) break;
    char* q;
    const double v = std::strtod(p, &q);
    if (q == p) return false;
    out.push_back(v);
    p = q;
  }
  pos = rb + 1;
  return true;
}
static bool read_scalar(const std::string& text, const char* key, size_t from, double& v) {
  const std::string k = std::string(key) + ":";
  const size_t p = text.find(k, from);
  if (p == std::string::npos) return false;
  const char* s = text.c_str() + p + k.size();
  char* q;
  v = std::strtod(s, &q);
  return q != s;
}

static SvmModel* load_svm(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f.good()) {
    set_error(std::string("Error: File ") + path + " does not exist!");  // learning.cpp:172-178
    return nullptr;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string text = ss.str();
  // OpenCV 2.4 (the reference's files): "my_svm: !!opencv-ml-svm"; OpenCV 3 / 4 write "opencv_ml_svm:" with the same keys
  if (text.find("opencv-ml-svm") == std::string::npos && text.find("opencv_ml_svm") == std::string::npos) {
    set_error("not an !!opencv-ml-svm file");
    return nullptr;
  }
  SvmModel* m = new SvmModel;
  auto bad = [&](const char* why) {
    set_error(std::string("SVM file: ") + why);
    delete m;
    return static_cast<SvmModel*>(nullptr);
  };
  const size_t kp = text.find("kernel:");
  if (kp == std::string::npos) return bad("kernel missing");
  const std::string kline = text.substr(kp, text.find('}', kp) - kp);
  if (kline.find("LINEAR") != std::string::npos) m->kernel = 0;
  else if (kline.find("POLY") != std::string::npos) m->kernel = 1;
  else return bad("only LINEAR and POLY kernels are supported");
  double v;
  if (m->kernel == 1) {
    if (read_scalar(kline, "degree", 0, v)) m->degree = int(v);
    if (read_scalar(kline, "gamma", 0, v)) m->gamma = v;
    if (read_scalar(kline, "coef0", 0, v)) m->coef0 = v;
    if (m->degree < 1) return bad("POLY degree must be a positive integer");
  }
  if (!read_scalar(text, "var_count", 0, v)) return bad("var_count missing");
  m->var_count = int(v);
  if (!read_scalar(text, "sv_total", 0, v)) return bad("sv_total missing");
  m->sv_total = int(v);
  size_t pos = text.find("support_vectors:");
  if (pos == std::string::npos) return bad("support_vectors missing");
  m->sv.reserve(size_t(m->sv_total) * m->var_count);
  for (int k = 0; k < m->sv_total; k++) {
    std::vector<double> row;
    if (!read_list(text, pos, row) || int(row.size()) != m->var_count) return bad("bad support vector row");
    for (double d : row) m->sv.push_back(float(d));
  }
  const size_t df = text.find("decision_functions:", pos);
  if (df == std::string::npos) return bad("decision_functions missing");
  if (!read_scalar(text, "sv_count", df, v)) return bad("sv_count missing");
  m->sv_count = int(v);
  if (!read_scalar(text, "rho", df, v)) return bad("rho missing");
  m->rho = v;
  size_t ap = text.find("alpha:", df);
  if (ap == std::string::npos || !read_list(text, ap, m->alpha) || int(m->alpha.size()) != m->sv_count)
    return bad("bad alpha");
  m->index.resize(m->sv_count);
  for (int k = 0; k < m->sv_count; k++) m->index[k] = k;
  size_t ip = text.find("index:", ap);
  if (ip != std::string::npos) {
    std::vector<double> idx;
    if (read_list(text, ip, idx) && int(idx.size()) == m->sv_count)
      for (int k = 0; k < m->sv_count; k++) m->index[k] = int(idx[k]);
  }
  for (int k : m->index)
    if (k < 0 || k >= m->sv_total) return bad("index out of range");
  return m;
}

static float elapsed(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {  // an event that was not recorded in this call
    cudaGetLastError();
    return 0.f;
  }
  return ms;
}

static int ensure_out(Ctx* c, size_t bytes) {
  if (bytes <= c->h_out_cap) return 0;
  g_alloc_gen++;
  if (c->h_out) cudaFreeHost(c->h_out);
  c->h_out = nullptr;
  c->h_out_cap = 0;
  const size_t want = bytes + bytes / 4 + 4096;
  if (cudaHostAlloc(&c->h_out, want, cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer(&c->d_out_mapped, c->h_out, 0) != cudaSuccess) {
    set_error("cudaHostAlloc(mapped) failed");
    return AG_ERR_CUDA;
  }
  c->h_out_cap = want;
  return 0;
}

// The full device-side path, cloud already in device memory.  Everything is enqueued without a single
// host round trip: the voxel count, the sample list, the hypothesis count and the grasp records stay
// on the device until the export kernel writes them into mapped host memory; the host waits once.
static void launch_export(Ctx* c, const PeerOut& peer) {
  HostOut* hdr = static_cast<HostOut*>(c->d_out_mapped);
  ag_grasp* recs = reinterpret_cast<ag_grasp*>(hdr + 1);
  int* exp_hdr = static_cast<int*>(c->d_export);
  ag_grasp* exp_recs = c->d_export ? reinterpret_cast<ag_grasp*>(static_cast<char*>(c->d_export) + 16) : nullptr;
  const int cap_exp = c->d_export ? int((c->d_export_cap - 16) / sizeof(ag_grasp)) : 0;
  k_export<<<32, 256, 0, c->stream>>>(c->grasps_raw.as<ag_grasp>(), c->hyp_slots.as<int>(), c->pend_nsel,
                                      c->attached_svm ? c->scores.as<float>() : nullptr, c->row_index.as<RowIndex>(),
                                      hand_sweep_overflow_ptr(c), c->counters.as<unsigned long long>(), hdr, recs,
                                      c->grasps.as<ag_grasp>(), exp_hdr, exp_recs, int(size_t(c->pend_S) * 8), cap_exp, peer,
                                      c->stamp, c->pend_slot_first, c->pend_slot_step, c->scores_by_slot ? 1 : 0);
}

// consumer side of the peer gather for the current epoch (right behind the export of the same call)
static void launch_merge(Ctx* c, const PeerOut& peer) {
  if (peer.world <= 0) return;
  MergeArgs A;
  std::memset(&A, 0, sizeof(A));
  const size_t half = size_t(c->gather_world) * c->gather_slot_bytes;
  A.buf = static_cast<const char*>(c->gather_buf) + kAckBytes + (peer.epoch & 1u) * half;
  A.slot_bytes = c->gather_slot_bytes;
  for (int r = 0; r < c->gather_world; r++)
    A.ack_peer[r] = reinterpret_cast<unsigned*>(static_cast<char*>(c->gather_peer[r]) + size_t(c->gather_rank) * kAckStride);
  A.world = c->gather_world;
  A.epoch = peer.epoch;
  A.overflow = c->overflow.p ? hand_sweep_overflow_ptr(c) : static_cast<const int*>(c->gather_zero);
  A.final_pass = peer.final_pass;
  A.merged_dev = static_cast<ag_grasp*>(c->gather_merged);
  A.host = static_cast<GatherHost*>(c->gather_host_dev);
  A.cap_total = c->gather_cap_total;
  A.done = static_cast<unsigned*>(c->gather_done) + 4;
  A.mask_off = peer.mask_off;
  A.plan = static_cast<int*>(c->gather_plan);
  A.plan_cap = c->gather_plan_cap;
  k_gather_plan<<<1, 1024, 0, c->stream>>>(A);
  k_gather_place<<<64, 256, 0, c->stream>>>(A);
  c->launches += 2;
}

// First half of a localize call: everything is enqueued on the context's stream, nothing is waited for.
static int localize_begin(Ctx* c, const void* d_points, int stride, int n_in, int size_left, const int* indices,
                          int n_indices, unsigned flags) {
  c->pend_active = false;
  next_stamp(c);
  std::memset(&c->timings, 0, sizeof(c->timings));
  c->timings.n_in = n_in;
  c->images_valid = false;
  c->scores_valid = false;
  c->n_hyp = 0;
  c->launches = 0;
  cudaStream_t st = c->stream;
  c->two_cams = size_left < n_in;
  const bool given = indices && n_indices > 0;
  const int* const indices_all = indices;
  const int S_total = given ? n_indices : std::max(0, c->params.num_samples);
  // sample sharding (ag_params.shard_index / shard_count): this context's share of the samples — a contiguous
  // range, or (shard_interleave) every shard_count-th sample, which balances scenes whose hypotheses cluster
  const int sh_n = std::max(1, c->params.shard_count), sh_i = std::min(std::max(0, c->params.shard_index), sh_n - 1);
  const bool interleave = sh_n > 1 && c->params.shard_interleave != 0;
  const int k_lo = int((long long)S_total * sh_i / sh_n), k_hi = int((long long)S_total * (sh_i + 1) / sh_n);
  const int k_first = interleave ? sh_i : k_lo, k_step = interleave ? sh_n : 1;
  const int S = interleave ? (S_total > sh_i ? (S_total - sh_i + sh_n - 1) / sh_n : 0) : k_hi - k_lo;
  c->n_samples = S;
  c->pend_slot_first = k_first;
  c->pend_slot_step = k_step;
  const size_t slots = size_t(S) * 8;
  RowIndex* ri = c->row_index.as<RowIndex>();
  if (given) {  // caller's indices go to a staging buffer first: the (graph-captured) body only copies device to device
    if (c->sample_stage.reserve(size_t(std::max(S, 1)) * 4)) return AG_ERR_CUDA;
    if (interleave) {
      c->sample_host.resize(S);
      for (int j = 0; j < S; j++) c->sample_host[j] = indices[k_first + j * k_step];
      indices = c->sample_host.data();
    } else {
      indices += k_lo;
    }
    if (S > 0) AG_CUDA_CHECK(cudaMemcpyAsync(c->sample_stage.p, indices, size_t(S) * 4, cudaMemcpyHostToDevice, st));
  }
  // a share of the samples in the production normal mode: the whole sample list is needed once more, to slice the
  // reference's rand() stream exactly as the unsharded call does (fit_quadrics_device, RandShare)
  const bool rand_share = sh_n > 1 && c->params.deterministic_normals == 0 && S_total > 0;
  if (rand_share) {
    if (c->samples_all.reserve(size_t(S_total) * 4) || c->count_all.reserve(16)) return AG_ERR_CUDA;
    AG_CUDA_CHECK(cudaMemcpyAsync(c->count_all.p, &S_total, 4, cudaMemcpyHostToDevice, st));
    if (given)
      AG_CUDA_CHECK(cudaMemcpyAsync(c->samples_all.p, indices_all, size_t(S_total) * 4, cudaMemcpyHostToDevice, st));
  }
  int* d_nsel = nullptr;
  // Everything from the voxelisation to the scoring: ~25 launches without a host dependency.  The second call
  // with the same shapes captures it into a CUDA graph, later calls replay the graph (one launch call, no
  // per-kernel launch gaps); any change of shape, parameters or buffers falls back to the eager path.
  auto body = [&]() -> int {
    record_event(c, c->ev[1]);
    // the buffers the voxelisation kernels reset / fill on the way (fold_resets, fold_draw) exist before it is enqueued
    struct FoldGuard {  // nothing of this call's folding survives it (error paths included)
      Ctx* c;
      ~FoldGuard() {
        c->fold_resets = 0;
        c->fold_draw = nullptr;
      }
    } fold_guard{c};
    DrawArgs draw;
    std::memset(&draw, 0, sizeof(draw));
    c->fold_resets = 0;
    c->fold_draw = nullptr;
    if (S > 0) {
      if (c->samples.reserve(size_t(std::max(S, n_in)) * 4) || c->counters.reserve(64) ||
          c->overflow.reserve(size_t(S + 1) * 4))
        return AG_ERR_CUDA;
      if (c->inline_big && c->overflow_list.reserve(size_t(S + 1) * 4)) return AG_ERR_CUDA;
      c->fold_resets = 1u | 2u | (c->params.deterministic_normals == 0 ? 4u : 0u) | (c->inline_big ? 8u : 0u);
      if (!given && !(flags & (AG_FLAG_USE_CLUSTERING | AG_FLAG_CALC_ANTIPODAL))) {
        draw.out = c->samples.as<int>();
        draw.seed = c->params.seed;
        draw.s_req = S_total;
        draw.k_first = k_first;
        draw.k_step = k_step;
        draw.count = S;
        c->fold_draw = &draw;
      }
    }
    int rc = quadric_rand_reset(c);
    if (rc) return rc;
    rc = preprocess_device(c, d_points, stride, n_in, size_left);
    c->fold_draw = nullptr;
    if (rc) return rc;
    if (flags & AG_FLAG_USE_CLUSTERING) {  // localization.cpp:51-98 (training-time path: waits for the stream)
      rc = remove_plane_device(c);
      if (rc) return rc;
    }
    record_event(c, c->ev[2]);
    record_event(c, c->ev[3]);
    ri = c->row_index.as<RowIndex>();
    if (S == 0) return AG_OK;
    if (c->samples.reserve(size_t(std::max(S, n_in)) * 4) || c->frames.reserve(size_t(S) * sizeof(ag_frame)) ||
        c->counters.reserve(64) || ensure_out(c, sizeof(HostOut) + slots * sizeof(ag_grasp)))
      return AG_ERR_CUDA;
    if (c->fold_resets & 1u) c->fold_resets &= ~1u;  // (zeroed by k_init_state)
    else AG_CUDA_CHECK(cudaMemsetAsync(c->counters.p, 0, 64, st));
    if (flags & AG_FLAG_CALC_ANTIPODAL) {
      // hand_search.cpp:17-26: normals for ALL points with radius 0.01 (launch bound = number of inputs)
      DevBuf& all_frames = c->all_frames;
      if (all_frames.reserve(size_t(n_in) * sizeof(ag_frame))) return AG_ERR_CUDA;
      k_iota<<<(n_in + 255) / 256, 256, 0, st>>>(c->samples.as<int>(), n_in);
      rc = fit_quadrics_device(c, c->samples.as<int>(), n_in, &ri->n_points, c->params.nn_radius_normals,
                               all_frames.as<ag_frame>(), true);
      if (rc) return rc;
      AG_CUDA_CHECK(cudaMemsetAsync(c->counters.p, 0, 64, st));
    }
    record_event(c, c->ev[4]);
    if (given) {
      AG_CUDA_CHECK(cudaMemcpyAsync(c->samples.p, c->sample_stage.p, size_t(S) * 4, cudaMemcpyDeviceToDevice, st));
      k_check_samples<<<(S + 255) / 256, 256, 0, st>>>(ri, S, c->samples.as<int>());
    } else {
      if (!c->draw_folded) {
        DrawArgs d2 = {c->samples.as<int>(), c->params.seed, S_total, k_first, k_step, S};
        k_draw_samples<<<(S + 255) / 256, 256, 0, st>>>(ri, d2);
        c->launches += 1;
      }
    }
    if (given) c->launches += 1;
    RandShare share;
    if (rand_share) {
      if (!given)  // (the full draw first: the share's own draw below leaves its count in ri->n_samples)
      {
        DrawArgs d3 = {c->samples_all.as<int>(), c->params.seed, S_total, 0, 1, S_total};
        k_draw_samples<<<(S_total + 255) / 256, 256, 0, st>>>(ri, d3, 0);
      }
      share.d_all = c->samples_all.as<int>();
      share.d_count_all = c->count_all.as<int>();
      share.n_all = S_total;
      share.first = k_first;
      share.step = k_step;
    }
    rc = fit_quadrics_device(c, c->samples.as<int>(), S, &ri->n_samples, c->params.nn_radius_taubin,
                             c->frames.as<ag_frame>(), true, rand_share ? &share : nullptr);
    if (rc) return rc;
    record_event(c, c->ev[5]);
    // a single-vector (linear) model is scored straight from the sweep's unordered hypothesis list while the stable
    // compaction for the export runs beside it; models with many support vectors go through the ordered list
    const bool fork = c->attached_svm && c->attached_svm->sv_total == 1;
    rc = hand_sweep_enqueue(c, c->samples.as<int>(), S, c->frames.as<ag_frame>(),
                            c->params.filters_boundaries ? 0x100u : 0u, fork, true, c->inline_big);
    if (rc) return rc;
    d_nsel = hand_sweep_count_ptr(c, S);
    record_event(c, c->ev[8]);
    c->scores_by_slot = false;
    if (c->attached_svm) {  // fused scoring, hypothesis count read on the device
      if (c->scores.reserve(slots * 8 + 64)) return AG_ERR_CUDA;
      if (fork) {
        c->scores_by_slot = true;
        rc = hog_svm_device(c, c->attached_svm, c->images_raw.as<uint32_t>(), hand_sweep_list_ptr(c), int(slots),
                            hand_sweep_list_count_ptr(c), nullptr, c->scores.as<float>(), nullptr, true);
        if (rc) return rc;
        AG_CUDA_CHECK(cudaStreamWaitEvent(st, c->ev_join, 0));  // the ordered list is ready for the export
      } else {
        rc = hog_svm_device(c, c->attached_svm, c->images_raw.as<uint32_t>(), c->hyp_slots.as<int>(), int(slots), d_nsel,
                            nullptr, c->scores.as<float>(), nullptr);
        if (rc) return rc;
      }
    }
    record_event(c, c->ev[9]);
    record_event(c, c->ev[6]);
    return AG_OK;
  };
  GraphKey key;
  std::memset(&key, 0, sizeof(key));
  key.d_points = d_points;
  key.stride = stride;
  key.n_in = n_in;
  key.size_left = size_left;
  key.S = S;
  key.given = given ? 1 : 0;
  key.flags = flags | (c->stage_timing ? 0x80000000u : 0u) |  // (a graph with and one without the stage events)
              (c->inline_big ? 0x40000000u : 0u);              // (... and with the inline large-slab pass)
  key.state_gen = c->state_gen;
  key.svm = c->attached_svm;
  static const bool graphs_off = getenv("AG_NO_GRAPH") != nullptr;
  const bool can_graph = !graphs_off && S > 0 && !(flags & (AG_FLAG_CALC_ANTIPODAL | AG_FLAG_USE_CLUSTERING));
  // small cache of captured pipelines (callers typically alternate between a few input buffers)
  GraphSlot* slot = nullptr;
  for (GraphSlot& g : c->gslots)
    if (g.valid && std::memcmp(&key, &g.key, sizeof(key)) == 0 && g.allocgen == g_alloc_gen) slot = &g;
  c->g_tick++;
  int rc = AG_OK;
  if (can_graph && slot && slot->exec) {  // replay
    AG_CUDA_CHECK(cudaGraphLaunch(slot->exec, st));
    c->launches = slot->launches;
    slot->last_use = c->g_tick;
    d_nsel = hand_sweep_count_ptr(c, S);
    c->n_graph_replays++;
  } else if (can_graph && slot) {  // second call with these shapes: capture, instantiate, launch
    const int launches0 = c->launches;
    cudaGraph_t graph = nullptr;
    AG_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    c->capturing = true;
    rc = body();
    c->capturing = false;
    const cudaError_t ce = cudaStreamEndCapture(st, &graph);
    const bool moved = slot->allocgen != g_alloc_gen;
    if (rc == AG_OK && ce == cudaSuccess && !moved && cudaGraphInstantiate(&slot->exec, graph, 0) == cudaSuccess) {
      slot->launches = c->launches;
      slot->last_use = c->g_tick;
      AG_CUDA_CHECK(cudaGraphLaunch(slot->exec, st));
    } else {  // could not capture (a buffer moved, an error): run it eagerly
      cudaGetLastError();
      slot->exec = nullptr;
      slot->valid = false;
      c->launches = launches0;
      rc = body();
    }
    if (graph) cudaGraphDestroy(graph);
    if (rc) return rc;
  } else {  // first call with these shapes (or graphs not applicable): eager, remember the shapes
    rc = body();
    if (rc) return rc;
    if (can_graph) {
      GraphSlot* victim = &c->gslots[0];
      for (GraphSlot& g : c->gslots) {
        if (!g.valid || g.allocgen != g_alloc_gen) {
          victim = &g;
          break;
        }
        if (g.last_use < victim->last_use) victim = &g;
      }
      if (victim->exec) cudaGraphExecDestroy(victim->exec);
      victim->exec = nullptr;
      victim->key = key;
      victim->valid = true;
      victim->allocgen = g_alloc_gen;
      victim->last_use = c->g_tick;
    }
  }
  PeerOut peer;
  std::memset(&peer, 0, sizeof(peer));
  if (c->gather_world > 0 && c->gather_connected) {
    c->gather_epoch++;
    const size_t half = size_t(c->gather_world) * c->gather_slot_bytes;
    for (int r = 0; r < c->gather_world; r++)
      peer.slot[r] = static_cast<char*>(c->gather_peer[r]) + kAckBytes + (c->gather_epoch & 1u) * half +
                     size_t(c->gather_rank) * c->gather_slot_bytes;
    peer.ack = static_cast<const unsigned*>(c->gather_buf);
    peer.world = c->gather_world;
    peer.cap = c->gather_rec_cap;
    peer.epoch = c->gather_epoch;
    peer.done = static_cast<unsigned*>(c->gather_done);
    peer.mask_off = size_t(kSlotHeaderBytes) + size_t(c->gather_rec_cap) * sizeof(ag_grasp);
    peer.mask_cap = c->gather_mask_cap;
    peer.valid = S > 0 ? c->valid.as<uint8_t>() : nullptr;
    peer.n_local = S;
    peer.first = k_first;
    peer.step = k_step;
  }
  static_assert(sizeof(PeerOut) <= sizeof(c->pend_peer), "pending export arguments");
  std::memcpy(c->pend_peer, &peer, sizeof(peer));
  if (S == 0) {  // (localize_end reads the cloud size back); a connected gather still gets this rank's (empty) list
    if (peer.world > 0) {
      k_publish_empty<<<1, 32, 0, st>>>(peer, ri);
      peer.final_pass = 1;
      launch_merge(c, peer);
      c->gather_pending = true;
    }
    return AG_OK;
  }
  c->pend_S = S;
  c->pend_nsel = d_nsel;
  launch_export(c, peer);
  launch_merge(c, peer);
  c->gather_pending = peer.world > 0;
  c->launches += 1;
  cudaEventRecord(c->ev[7], st);
  c->pend_active = true;
  return AG_OK;
}

// after the stream has been waited for: status of the merged list of the peer gather
static int gather_collect(Ctx* c) {
  c->gather_pending = false;
  const GatherHost* gh = static_cast<const GatherHost*>(c->gather_host);
  c->gather_valid = true;
  if (gh->status & 1) {
    set_error("peer gather: a rank did not publish its grasp list (timeout)");
    return AG_ERR_CUDA;
  }
  if (gh->status & 2) {
    set_error("peer gather: the merged list exceeds the gather capacity (ag_gather_create num_samples too small)");
    return AG_ERR_CAPACITY;
  }
  for (int r = 0; r < c->gather_world; r++)
    if (gh->err_per_rank[r] & 0x100) {
      set_error("peer gather: a rank produced more hypotheses than a gather slot holds");
      return AG_ERR_CAPACITY;
    }
  return AG_OK;
}

// Second half: the one wait of the call, error flags, the rare large-slab re-run, the host copy of the list.
static int localize_end(Ctx* c, ag_grasp** out, int* n_out) {
  *out = nullptr;
  *n_out = 0;
  cudaStream_t st = c->stream;
  if (!c->pend_active) {  // no samples requested: only the voxelised cloud exists
    int rc = fetch_cloud_size(c);
    c->timings.n_voxels = c->n_vox;
    if (rc == AG_OK && c->gather_pending) rc = gather_collect(c);
    return rc;
  }
  c->pend_active = false;
  const int S = c->pend_S;
  PeerOut peer;
  std::memcpy(&peer, c->pend_peer, sizeof(peer));
  int rc = AG_OK;
  AG_CUDA_CHECK(cudaStreamSynchronize(st));
  AG_CUDA_CHECK(cudaGetLastError());
  HostOut* h = static_cast<HostOut*>(c->h_out);
  c->n_vox = h->n_vox;
  c->timings.n_voxels = h->n_vox;
  c->timings.n_samples = h->n_samples;
  if (h->error & kErrBitmapRetry) return AG_RETRY_KEYSORT;  // (the caller re-runs the call on the key-sort path)
  if (h->error & kErrKeyOverflow) {
    set_error("voxel index exceeds the key range (workspace extent / voxel_size > 2^21 cells)");
    return AG_ERR_CAPACITY;
  }
  if (h->error & kErrBadIndex) {
    set_error("sample index out of range of the voxelised cloud");
    return AG_ERR_INVALID;
  }
  if (h->error & kErrBallOverflow) {
    set_error("a radius ball holds more points than a neighbour-pool slot (cloud far denser than a voxelised surface)");
    return AG_ERR_CAPACITY;
  }
  int Hn = h->n_hyp;
  c->n_hyp = Hn;
  c->images_valid = true;
  if (h->n_over > 0) {
    // some samples need the large-capacity sweep (dense neighbourhoods: ~0.5 % of the samples of the fused 7-view
    // cloud of config 5): redo them, score THEIR hypotheses, export again — one more enqueue and one more wait.  A
    // single-vector model scored the first pass by raw slot, so only the redone samples' hypotheses (a fresh scorer's
    // list) need the HOG kernel; models with many support vectors rescore the ordered list.
    c->inline_big = true;  // from the next call on the large-slab pass is part of the pipeline (no host round trip)
    const size_t slots = size_t(S) * 8;
    const bool by_slot = c->attached_svm && c->scores_by_slot;
    rc = hand_sweep_rerun_enqueue(c, S, h->n_over, by_slot);
    if (rc) return rc;
    if (c->attached_svm) {
      if (by_slot)
        rc = hog_svm_device(c, c->attached_svm, c->images_raw.as<uint32_t>(), hand_sweep_list_ptr(c), int(slots),
                            hand_sweep_list_count_ptr(c), nullptr, c->scores.as<float>(), nullptr, true);
      else
        rc = hog_svm_device(c, c->attached_svm, c->images_raw.as<uint32_t>(), c->hyp_slots.as<int>(), int(slots),
                            hand_sweep_count_ptr(c, S), nullptr, c->scores.as<float>(), nullptr);
      if (rc) return rc;
    }
    c->pend_nsel = hand_sweep_count_ptr(c, S);
    peer.final_pass = 1;
    launch_export(c, peer);
    launch_merge(c, peer);
    cudaEventRecord(c->ev[7], st);  // (total_ms covers this pass and the host round trip in front of it)
    AG_CUDA_CHECK(cudaStreamSynchronize(st));
    AG_CUDA_CHECK(cudaGetLastError());
    if (h->n_over > 0) {  // (the overflow counter was zeroed before the large-slab pass)
      set_error("hand sweep: a sample's slab exceeded the shared-memory capacity (9600 points)");
      return AG_ERR_CAPACITY;
    }
    Hn = h->n_hyp;
    c->n_hyp = Hn;
  }
  ag_grasp* res = static_cast<ag_grasp*>(std::malloc(std::max<size_t>(1, size_t(Hn)) * sizeof(ag_grasp)));
  if (Hn > 0) std::memcpy(res, reinterpret_cast<const ag_grasp*>(h + 1), size_t(Hn) * sizeof(ag_grasp));
  c->timings_pending = true;  // the event differences are read when ag_get_timings asks for them
  c->timings.n_hyp = Hn;
  c->timings.kernel_launches = c->launches;
  c->timings.taubin_neighbor_points = int64_t(h->counters[0]);
  c->timings.taubin_candidates = int64_t(h->counters[1]);
  c->timings.hand_neighbor_points = int64_t(h->counters[2]);
  c->timings.hand_candidates = int64_t(h->counters[3]);
  if (c->attached_svm) {  // what ag_classify hands back for this model: decision value + label per hypothesis
    c->scores_valid = true;
    c->last_scores.resize(Hn);
    c->last_labels.resize(Hn);
    for (int i = 0; i < Hn; i++) {
      c->last_scores[i] = res[i].score;
      c->last_labels[i] = res[i].label;
    }
  }
  *out = res;
  *n_out = Hn;
  if (c->gather_pending) return gather_collect(c);
  return AG_OK;
}

// localize_begin + localize_end; a cloud whose voxel lattice does not fit the occupancy bitmap is re-run once on the
// key-sort path, and the context stays on that path (the scene extents of a camera setup do not change per frame)
static int localize_run(Ctx* c, const void* d_points, int stride, int n_in, int size_left, const int* indices,
                        int n_indices, unsigned flags, ag_grasp** out, int* n_out) {
  int rc = localize_begin(c, d_points, stride, n_in, size_left, indices, n_indices, flags);
  if (rc == AG_OK) rc = localize_end(c, out, n_out);
  if (rc == AG_RETRY_KEYSORT) {
    c->bitmap_ok = false;
    c->state_gen++;
    rc = localize_begin(c, d_points, stride, n_in, size_left, indices, n_indices, flags);
    if (rc == AG_OK) rc = localize_end(c, out, n_out);
  }
  return rc;
}

}  // namespace ag

using namespace ag;

struct ag_svm {
  SvmModel* m;
};

extern "C" {

const char* ag_last_error(void) { return g_error.c_str(); }

void ag_default_params(ag_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->finger_width = 0.01;  // find_grasps.cpp:13-17
  p->hand_outer_diameter = 0.09;
  p->hand_depth = 0.06;
  p->hand_height = 0.02;
  p->init_bite = 0.01;
  const double ws[6] = {-10, 10, -10, 10, -10, 10};
  std::memcpy(p->workspace, ws, sizeof(ws));
  const double base_tf[16] = {0, 0.445417, 0.895323, 0.215, 1, 0, 0, -0.015, 0, 0.895323, -0.445417, 0.23, 0, 0, 0, 1};
  std::memcpy(p->cam_tf_left, base_tf, sizeof(base_tf));
  std::memcpy(p->cam_tf_right, base_tf, sizeof(base_tf));
  p->nn_radius_taubin = 0.03;  // hand_search.h:85
  p->nn_radius_hands = 0.08;
  p->nn_radius_normals = 0.01;  // hand_search.cpp:20
  p->voxel_size = 0.003;        // localization.cpp:43
  p->num_samples = 2000;        // find_grasps.cpp:11
  p->num_threads = 1;
  p->deterministic_normals = 1;
  p->shard_index = 0;
  p->shard_count = 1;
  p->filters_boundaries = 0;
  p->fix_cam_source = 0;
  p->seed = 20150320;
}

ag_ctx* ag_create(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error(std::string("no CUDA device available (there is no CPU fallback): ") +
              (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    return nullptr;
  }
  if (device < 0 || device >= count) {
    set_error("device index out of range");
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    set_error("cudaSetDevice failed");
    return nullptr;
  }
  ag_ctx* h = new ag_ctx;
  Ctx& c = h->c;
  c.device = device;
  c.serial = ++g_ctx_serial;
  if (cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_error("cudaStreamCreate failed");
    delete h;
    return nullptr;
  }
  cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c.ev_fork0, cudaEventDisableTiming);
  for (auto& ev : c.ev) cudaEventCreate(&ev);
  for (auto& ev : c.ev_k) cudaEventCreate(&ev);
  ag_default_params(&c.params);
  compute_hand_const(c.params, c.hand);
  std::memset(&c.timings, 0, sizeof(c.timings));
  return h;
}

void ag_destroy(ag_ctx* h) {
  if (!h) return;
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  cudaStreamSynchronize(c.stream);
  ag_gather_destroy(h);
  for (ag_ctx* ch : h->children) ag_destroy(ch);
  h->children.clear();
  for (GraphSlot& g : c.gslots)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  if (c.h_out) cudaFreeHost(c.h_out);
  for (DevBuf* b : {&c.raw, &c.keys, &c.keys_sorted, &c.keys_unique, &c.cub_tmp, &c.block_counts, &c.misc, &c.bitmap, &c.tile_state, &c.vox,
                    &c.row_ptr, &c.col_ptr, &c.row_index, &c.all_frames,
                    &c.normals, &c.samples, &c.sample_stage, &c.samples_all, &c.nn_counts_all, &c.count_all, &c.moments, &c.frames, &c.nn_counts, &c.nbr_pool, &c.nbr_heads, &c.rand_raw, &c.rand_off, &c.rand_carry, &c.picks, &c.quad_par, &c.grasps_raw, &c.valid,
                    &c.images_raw, &c.hyp_slots, &c.hyp_list, &c.overflow_list, &c.sample_q, &c.block_flags, &c.grasps, &c.counters, &c.scores, &c.descriptors, &c.kvals, &c.handle_in, &c.handle_bits, &c.sweep_dbg, &c.overflow})
    b->release();
  if (c.h_pinned) cudaFreeHost(c.h_pinned);
  for (auto& ev : c.ev) cudaEventDestroy(ev);
  for (auto& ev : c.ev_k) cudaEventDestroy(ev);
  cudaEventDestroy(c.ev_fork);
  cudaEventDestroy(c.ev_join);
  cudaEventDestroy(c.ev_fork0);
  cudaStreamDestroy(c.stream2);
  cudaStreamDestroy(c.stream);
  delete h;
}

int ag_set_params(ag_ctx* h, const ag_params* p) {
  if (!h || !p) return AG_ERR_INVALID;
  if (!(p->voxel_size > 0) || !(p->nn_radius_taubin > 0) || !(p->nn_radius_hands > 0) || !(p->hand_depth > 0) ||
      p->shard_count < 0 || p->shard_index < 0 || (p->shard_count > 0 && p->shard_index >= p->shard_count)) {
    set_error("invalid parameters");
    return AG_ERR_INVALID;
  }
  h->c.params = *p;
  compute_hand_const(h->c.params, h->c.hand);
  h->c.state_gen++;
  return AG_OK;
}
int ag_get_params(ag_ctx* h, ag_params* p) {
  *p = h->c.params;
  return AG_OK;
}
int ag_set_stage_timing(ag_ctx* h, int on) {
  h->c.stage_timing = on != 0;
  return AG_OK;
}
int ag_get_timings(ag_ctx* h, ag_timings* t) {
  Ctx& c = h->c;
  if (c.timings_pending) {  // stage times of the last ag_localize, from the events recorded in its stream
    c.timings_pending = false;
    c.timings.total_ms = elapsed(c.ev[0], c.ev[7]);
    if (!c.stage_timing) {
      *t = c.timings;
      return AG_OK;
    }
    c.timings.preprocess_ms = elapsed(c.ev[1], c.ev[2]);
    c.timings.grid_ms = 0.f;  // the x-row index is built inside the voxelisation pass
    c.timings.normals_all_ms = elapsed(c.ev[3], c.ev[4]);
    c.timings.quadric_ms = elapsed(c.ev[4], c.ev[5]);
    c.timings.sweep_ms = elapsed(c.ev[5], c.ev[8]);
    c.timings.hog_svm_ms = elapsed(c.ev[8], c.ev[9]);
    c.timings.d2h_ms = elapsed(c.ev[6], c.ev[7]);
    c.timings.total_ms = elapsed(c.ev[0], c.ev[7]);
    c.timings.search_ms = elapsed(c.ev_k[0], c.ev_k[1]);
    c.timings.moments_ms = elapsed(c.ev_k[1], c.ev_k[2]);
    c.timings.axes_ms = elapsed(c.ev_k[2], c.ev_k[3]);
    if (c.timings_h2d) c.timings.h2d_ms = elapsed(c.ev[0], c.ev[1]);
  }
  *t = c.timings;
  return AG_OK;
}
void ag_free(void* p) { std::free(p); }

ag_svm* ag_svm_load(const char* path) {
  SvmModel* m = load_svm(path);
  if (!m) return nullptr;
  ag_svm* s = new ag_svm;
  s->m = m;
  return s;
}
void ag_svm_free(ag_svm* s) {
  if (!s) return;
  if (s->m->d_sv) {
    cudaFree(s->m->d_sv);
    if (s->m->d_svT) cudaFree(s->m->d_svT);
    cudaFree(s->m->d_alpha);
    cudaFree(s->m->d_index);
  }
  delete s->m;
  delete s;
}
int ag_svm_info(const ag_svm* s, int* kernel_type, int* var_count, int* sv_total, double* rho) {
  if (kernel_type) *kernel_type = s->m->kernel;
  if (var_count) *var_count = s->m->var_count;
  if (sv_total) *sv_total = s->m->sv_total;
  if (rho) *rho = s->m->rho;
  return AG_OK;
}

int ag_localize(ag_ctx* h, const void* points, int stride, int n_in, int size_left, const int* indices, int n_indices,
                unsigned flags, ag_grasp** out, int* n_out) {
  if (!h || !out || !n_out) return AG_ERR_INVALID;
  *out = nullptr;
  *n_out = 0;
  if (n_in <= 0 || size_left == 0 || !points) {  // localization.cpp:9-15
    set_error("Input cloud is empty!");
    return AG_ERR_EMPTY;
  }
  if (stride < 12) {
    set_error("stride must be >= 12 bytes");
    return AG_ERR_INVALID;
  }
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  const size_t bytes = size_t(n_in) * stride;
  if (c.raw.reserve(bytes)) return AG_ERR_CUDA;
  cudaEventRecord(c.ev[0], c.stream);
  AG_CUDA_CHECK(cudaMemcpyAsync(c.raw.p, points, bytes, cudaMemcpyHostToDevice, c.stream));
  c.timings_h2d = true;
  return localize_run(&c, c.raw.p, stride, n_in, size_left, indices, n_indices, flags, out, n_out);
}

int ag_localize_device(ag_ctx* h, const void* d_points, int stride, int n_in, int size_left, const int* indices,
                       int n_indices, unsigned flags, ag_grasp** out, int* n_out) {
  if (!h || !out || !n_out) return AG_ERR_INVALID;
  if (n_in <= 0 || size_left == 0 || !d_points) {
    set_error("Input cloud is empty!");
    return AG_ERR_EMPTY;
  }
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  cudaEventRecord(c.ev[0], c.stream);
  c.timings_h2d = false;
  return localize_run(&c, d_points, stride, n_in, size_left, indices, n_indices, flags, out, n_out);
}

int ag_classify(ag_ctx* h, const ag_svm* svm, ag_grasp* grasps, int n, uint8_t* keep) {
  if (!h || !svm) return AG_ERR_INVALID;
  if (n <= 0) return AG_OK;
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  if (!c.images_valid) {
    set_error("ag_classify: no grasp images resident (call ag_localize / ag_hand_sweep first)");
    return AG_ERR_INVALID;
  }
  bool identity = n == c.n_hyp;
  for (int i = 0; i < n; i++) {
    const int id = grasps[i].image_id;
    if (id < 0 || id >= c.n_hyp || grasps[i].reserved != c.stamp) {
      set_error("ag_classify: hypothesis does not belong to the last ag_localize / ag_hand_sweep call on this context "
                "(records of an earlier call, another context or a batch lane cannot be scored: their grasp images "
                "are no longer resident)");
      return AG_ERR_INVALID;
    }
    identity = identity && id == i;
  }
  if (c.scores_valid && c.attached_svm == svm->m) {
    // already scored inside ag_localize (ag_set_svm): hand the results back
    for (int i = 0; i < n; i++) {
      const int id = grasps[i].image_id;
      grasps[i].score = c.last_scores[id];
      grasps[i].label = c.last_labels[id];
      if (keep) keep[i] = c.last_labels[id];
    }
    return AG_OK;
  }
  cudaEventRecord(c.ev[8], c.stream);
  DevBuf& sc = c.scores;
  if (sc.reserve(size_t(n) * 8 + 64)) return AG_ERR_CUDA;
  float* d_scores = sc.as<float>();
  const int* d_slots = c.hyp_slots.as<int>();  // image_id h -> raw (sample, orientation) slot
  if (!identity) {
    // arbitrary subset / order: translate ids to raw slots through a host copy of the table
    std::vector<int> raw_slots(c.n_hyp), slots(n);
    AG_CUDA_CHECK(cudaMemcpyAsync(raw_slots.data(), c.hyp_slots.p, size_t(c.n_hyp) * 4, cudaMemcpyDeviceToHost, c.stream));
    AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < n; i++) slots[i] = raw_slots[grasps[i].image_id];
    int* d_ids = reinterpret_cast<int*>(d_scores + n);
    AG_CUDA_CHECK(cudaMemcpyAsync(d_ids, slots.data(), size_t(n) * 4, cudaMemcpyHostToDevice, c.stream));
    AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    d_slots = d_ids;
  }
  int rc = hog_svm_device(&c, svm->m, c.images_raw.as<uint32_t>(), d_slots, n, nullptr, nullptr, d_scores);
  if (rc) return rc;
  std::vector<float> sc_h(n);
  AG_CUDA_CHECK(cudaMemcpyAsync(sc_h.data(), d_scores, size_t(n) * 4, cudaMemcpyDeviceToHost, c.stream));
  cudaEventRecord(c.ev[9], c.stream);
  AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  c.timings.hog_svm_ms = elapsed(c.ev[8], c.ev[9]);
  c.timings.kernel_launches = c.launches;
  for (int i = 0; i < n; i++) {
    grasps[i].score = sc_h[i];
    // CvSVM::predict: vote[sum > 0 ? 0 : 1], class_labels = [-1, 1] => label +1 <=> sum <= 0
    grasps[i].label = sc_h[i] > 0 ? 0 : 1;
    if (keep) keep[i] = grasps[i].label;
  }
  return AG_OK;
}

// ---- batches of clouds (BASELINE config 4): up to kBatchLanes clouds in flight on one GPU -------------------
// Every stage of one cloud is a few hundred warps of latency-bound work, so one cloud leaves most of the GPU
// idle; clouds are independent, so a batch is spread over lanes — the context itself plus lazily created
// children with the same parameters and SVM, each with its own stream, buffers and cached CUDA graph — and the
// lanes' pipelines overlap on the device.  Results are identical to calling ag_localize cloud by cloud.
int ag_localize_batch(ag_ctx* h, int n_clouds, const void* const* points, const int* strides, const int* n_in,
                      const int* size_left, unsigned flags, ag_grasp** out, int* n_out) {
  if (!h || n_clouds < 0 || (n_clouds > 0 && (!points || !strides || !n_in || !size_left || !out || !n_out))) return AG_ERR_INVALID;
  Ctx& c0 = h->c;
  cudaSetDevice(c0.device);
  for (int i = 0; i < n_clouds; i++) {
    out[i] = nullptr;
    n_out[i] = 0;
  }
  static const int max_lanes = [] {  // AG_BATCH_LANES overrides the default (tuning), 1..16
    const char* e = getenv("AG_BATCH_LANES");
    const int v = e ? atoi(e) : AG_BATCH_LANES;
    return std::min(16, std::max(1, v));
  }();
  const int lanes = std::min(n_clouds, max_lanes);
  while (int(h->children.size()) < lanes - 1) {
    ag_ctx* ch = ag_create(c0.device);
    if (!ch) return AG_ERR_CUDA;
    h->children.push_back(ch);
  }
  std::vector<Ctx*> lane(lanes);
  for (int l = 0; l < lanes; l++) {
    lane[l] = l == 0 ? &c0 : &h->children[l - 1]->c;
    if (l > 0 && (lane[l]->batch_parent_gen != c0.state_gen)) {  // mirror parameters / SVM of the parent
      int rc = ag_set_params(h->children[l - 1], &c0.params);
      if (rc) return rc;
      lane[l]->attached_svm = c0.attached_svm;
      lane[l]->state_gen++;
      lane[l]->batch_parent_gen = c0.state_gen;
    }
  }
  // rolling pipeline: cloud i goes to lane i % lanes; before a lane is reused its previous cloud is collected,
  // so every lane always has a cloud in flight while the host collects / enqueues on the others
  int first_err = AG_OK;
  std::vector<int> pending(lanes, -1), rcs(n_clouds, AG_OK);
  auto collect = [&](int l) {
    const int i = pending[l];
    if (i < 0) return;
    pending[l] = -1;
    if (rcs[i] == AG_OK) rcs[i] = localize_end(lane[l], &out[i], &n_out[i]);
    if (rcs[i] == AG_RETRY_KEYSORT) {  // lattice too large for the occupancy bitmap: this lane moves to the key-sort path
      lane[l]->bitmap_ok = false;
      lane[l]->state_gen++;
      rcs[i] = localize_run(lane[l], lane[l]->raw.p, strides[i], n_in[i], size_left[i], nullptr, 0, flags, &out[i], &n_out[i]);
    }
    if (rcs[i] == AG_ERR_EMPTY) rcs[i] = AG_OK;  // an empty cloud yields an empty list (localization.cpp:9-15)
    if (rcs[i] != AG_OK && first_err == AG_OK) first_err = rcs[i];
  };
  for (int i = 0; i < n_clouds; i++) {
    const int l = i % lanes;
    collect(l);
    Ctx& c = *lane[l];
    pending[l] = i;
    if (n_in[i] <= 0 || size_left[i] == 0 || !points[i] || strides[i] < 12) {
      rcs[i] = AG_ERR_EMPTY;
      continue;
    }
    const size_t bytes = size_t(n_in[i]) * strides[i];
    if (c.raw.reserve(bytes)) {
      rcs[i] = AG_ERR_CUDA;
      continue;
    }
    cudaEventRecord(c.ev[0], c.stream);
    if (cudaMemcpyAsync(c.raw.p, points[i], bytes, cudaMemcpyHostToDevice, c.stream) != cudaSuccess) {
      rcs[i] = AG_ERR_CUDA;
      continue;
    }
    rcs[i] = localize_begin(&c, c.raw.p, strides[i], n_in[i], size_left[i], nullptr, 0, flags);
  }
  for (int l = 0; l < lanes; l++) collect(l);
  return first_err;
}

int ag_set_export_buffer(ag_ctx* h, void* d_buffer, size_t bytes) {
  if (!h) return AG_ERR_INVALID;
  if (d_buffer && bytes < 16 + sizeof(ag_grasp)) {
    set_error("export buffer too small");
    return AG_ERR_INVALID;
  }
  h->c.d_export = d_buffer;
  h->c.d_export_cap = d_buffer ? bytes : 0;
  h->c.state_gen++;
  return AG_OK;
}

// ---- peer gather: the grasp-list all-gather fused into the export kernel (NVLink peer stores) ----------
// [header][8 records per sample][one valid-orientation mask per sample, padded to 16 bytes]
size_t ag_gather_slot_bytes(int num_samples) {
  return size_t(kSlotHeaderBytes) + size_t(8) * size_t(num_samples) * sizeof(ag_grasp) + ((size_t(num_samples) + 31) / 16) * 16;
}

int ag_gather_create(ag_ctx* h, int num_samples, int world, int rank, unsigned char* ipc_handle_out) {
  if (!h || world < 1 || world > AG_MAX_GATHER_RANKS || rank < 0 || rank >= world || num_samples < 1 || !ipc_handle_out) {
    set_error("ag_gather_create: bad arguments");
    return AG_ERR_INVALID;
  }
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  if (c.gather_buf) {
    set_error("ag_gather_create: already created for this context");
    return AG_ERR_INVALID;
  }
  c.gather_slot_bytes = ag_gather_slot_bytes(num_samples);
  c.gather_rec_cap = 8 * num_samples;
  c.gather_mask_cap = int(((size_t(num_samples) + 31) / 16) * 16);
  c.gather_plan_cap = num_samples * world + 64;
  AG_CUDA_CHECK(cudaMalloc(&c.gather_plan, (size_t(kPlanBase) + size_t(c.gather_plan_cap)) * sizeof(int)));
  AG_CUDA_CHECK(cudaMemset(c.gather_plan, 0, (size_t(kPlanBase) + size_t(c.gather_plan_cap)) * sizeof(int)));
  // [ack words of the consumers][two epochs (parity) x world slots]
  const size_t bytes = kAckBytes + 2 * size_t(world) * c.gather_slot_bytes;
  AG_CUDA_CHECK(cudaMalloc(&c.gather_buf, bytes));
  AG_CUDA_CHECK(cudaMemset(c.gather_buf, 0, bytes));
  AG_CUDA_CHECK(cudaMalloc(&c.gather_done, 64));
  AG_CUDA_CHECK(cudaMemset(c.gather_done, 0, 64));
  c.gather_zero = static_cast<char*>(c.gather_done) + 32;  // a device int that stays 0
  // the merged list: every rank's share of up to 8 hypotheses per sample of the whole call
  c.gather_cap_total = int(std::min<size_t>(size_t(8) * size_t(num_samples) * size_t(world), size_t(1) << 24));
  const size_t mbytes = size_t(c.gather_cap_total) * sizeof(ag_grasp);
  AG_CUDA_CHECK(cudaMalloc(&c.gather_merged, mbytes));
  AG_CUDA_CHECK(cudaHostAlloc(&c.gather_host, sizeof(GatherHost) + mbytes, cudaHostAllocMapped));
  std::memset(c.gather_host, 0, sizeof(GatherHost));
  AG_CUDA_CHECK(cudaHostGetDevicePointer(&c.gather_host_dev, c.gather_host, 0));
  cudaIpcMemHandle_t hd;
  AG_CUDA_CHECK(cudaIpcGetMemHandle(&hd, c.gather_buf));
  static_assert(sizeof(hd) == AG_IPC_HANDLE_BYTES, "IPC handle size");
  std::memcpy(ipc_handle_out, &hd, sizeof(hd));
  c.gather_world = world;
  c.gather_rank = rank;
  c.gather_epoch = 0;
  c.gather_connected = false;
  c.gather_valid = false;
  AG_CUDA_CHECK(cudaDeviceSynchronize());
  return AG_OK;
}

int ag_gather_connect(ag_ctx* h, const unsigned char* handles) {
  if (!h || !handles || h->c.gather_world < 1 || !h->c.gather_buf) {
    set_error("ag_gather_connect: ag_gather_create first");
    return AG_ERR_INVALID;
  }
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  for (int r = 0; r < c.gather_world; r++) {
    if (r == c.gather_rank) {
      c.gather_peer[r] = c.gather_buf;
      continue;
    }
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handles + size_t(r) * AG_IPC_HANDLE_BYTES, sizeof(hd));
    AG_CUDA_CHECK(cudaIpcOpenMemHandle(&c.gather_peer[r], hd, cudaIpcMemLazyEnablePeerAccess));
  }
  c.gather_connected = true;
  c.state_gen++;
  return AG_OK;
}

int ag_gather_result(ag_ctx* h, int32_t* n_hyp_per_rank, int* n_total, const ag_grasp** d_merged, const ag_grasp** h_merged) {
  if (!h || !h->c.gather_connected || !h->c.gather_valid) {
    set_error("ag_gather_result: no connected peer gather / no completed ag_localize yet");
    return AG_ERR_INVALID;
  }
  Ctx& c = h->c;
  const GatherHost* gh = static_cast<const GatherHost*>(c.gather_host);
  for (int r = 0; r < c.gather_world && n_hyp_per_rank; r++) n_hyp_per_rank[r] = gh->n_per_rank[r];
  if (n_total) *n_total = gh->n_total;
  if (d_merged) *d_merged = static_cast<const ag_grasp*>(c.gather_merged);
  if (h_merged) {  // host copy on demand (the merged list lives in device memory: its consumers are kernels)
    ag_grasp* dst = reinterpret_cast<ag_grasp*>(static_cast<char*>(c.gather_host) + sizeof(GatherHost));
    cudaSetDevice(c.device);
    if (gh->n_total > 0) {
      AG_CUDA_CHECK(cudaMemcpyAsync(dst, c.gather_merged, size_t(gh->n_total) * sizeof(ag_grasp), cudaMemcpyDeviceToHost, c.stream));
      AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
    *h_merged = dst;
  }
  return AG_OK;
}

int ag_gather_wait(ag_ctx* h, int32_t* n_hyp_per_rank, const void** d_slots, size_t* slot_bytes) {
  if (!h || !h->c.gather_connected || !h->c.gather_valid) {
    set_error("ag_gather_wait: no connected peer gather / no completed ag_localize yet");
    return AG_ERR_INVALID;
  }
  Ctx& c = h->c;
  // the exchange completed inside ag_localize (export -> merge in the same stream): this only reports it
  const GatherHost* gh = static_cast<const GatherHost*>(c.gather_host);
  for (int r = 0; r < c.gather_world && n_hyp_per_rank; r++) n_hyp_per_rank[r] = gh->n_per_rank[r];
  if (d_slots)
    *d_slots = static_cast<const char*>(c.gather_buf) + kAckBytes + (c.gather_epoch & 1u) * size_t(c.gather_world) * c.gather_slot_bytes;
  if (slot_bytes) *slot_bytes = c.gather_slot_bytes;
  return AG_OK;
}

int ag_gather_destroy(ag_ctx* h) {
  if (!h) return AG_ERR_INVALID;
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  cudaStreamSynchronize(c.stream);
  for (int r = 0; r < c.gather_world; r++)
    if (c.gather_connected && r != c.gather_rank && c.gather_peer[r]) cudaIpcCloseMemHandle(c.gather_peer[r]);
  if (c.gather_buf) cudaFree(c.gather_buf);
  if (c.gather_done) cudaFree(c.gather_done);
  if (c.gather_merged) cudaFree(c.gather_merged);
  if (c.gather_plan) cudaFree(c.gather_plan);
  c.gather_plan = nullptr;
  if (c.gather_host) cudaFreeHost(c.gather_host);
  c.gather_buf = c.gather_done = c.gather_merged = c.gather_host = c.gather_host_dev = c.gather_zero = nullptr;
  c.gather_world = 0;
  c.gather_connected = false;
  c.gather_valid = false;
  c.state_gen++;
  return AG_OK;
}

int ag_set_svm(ag_ctx* h, const ag_svm* svm) {
  if (!h) return AG_ERR_INVALID;
  h->c.attached_svm = svm ? svm->m : nullptr;
  h->c.scores_valid = false;
  h->c.state_gen++;
  return AG_OK;
}

int ag_get_points(ag_ctx* h, int image_id, double** pts3xm, int32_t** cam, int* m) {
  if (!h || !pts3xm || !cam || !m) return AG_ERR_INVALID;
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  *pts3xm = nullptr;
  *cam = nullptr;
  *m = 0;
  if (!c.images_valid || image_id < 0 || image_id >= c.n_hyp) {
    set_error("ag_get_points: image_id does not belong to the last localize / sweep call");
    return AG_ERR_INVALID;
  }
  int slot = 0;
  AG_CUDA_CHECK(cudaMemcpy(&slot, c.hyp_slots.as<int>() + image_id, 4, cudaMemcpyDeviceToHost));
  std::vector<double> p;
  std::vector<int> cm;
  int rc = box_points_device(&c, c.n_samples, slot, p, cm);
  if (rc) return rc;
  const int n = int(cm.size());
  double* op = static_cast<double*>(std::malloc(std::max<size_t>(1, p.size()) * sizeof(double)));
  int32_t* oc = static_cast<int32_t*>(std::malloc(std::max<size_t>(1, cm.size()) * sizeof(int32_t)));
  std::memcpy(op, p.data(), p.size() * sizeof(double));
  for (int i = 0; i < n; i++) oc[i] = cm[i];
  *pts3xm = op;
  *cam = oc;
  *m = n;
  return AG_OK;
}

// Training features (Learning::train / convertData, learning.cpp:76-163,249-290): for every hypothesis the HOG
// descriptor of its grasp image and of the images made from the points of camera 1 only and camera 2 only
// (createInstance(h, cam_pos), createInstance(h, cam_pos, 0), createInstance(h, cam_pos, 1)).
int ag_train_features(ag_ctx* h, const ag_grasp* grasps, int n, float* features) {
  if (!h || (n > 0 && (!grasps || !features))) return AG_ERR_INVALID;
  if (n <= 0) return AG_OK;
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  if (!c.images_valid) {
    set_error("ag_train_features: no grasp images resident (call ag_localize / ag_hand_sweep first)");
    return AG_ERR_INVALID;
  }
  for (int i = 0; i < n; i++)
    if (grasps[i].image_id < 0 || grasps[i].image_id >= c.n_hyp || grasps[i].reserved != c.stamp) {
      set_error("ag_train_features: hypothesis does not belong to the last ag_localize / ag_hand_sweep call on this context");
      return AG_ERR_INVALID;
    }
  std::vector<int> raw_slots(c.n_hyp), slots(n);
  AG_CUDA_CHECK(cudaMemcpyAsync(raw_slots.data(), c.hyp_slots.p, size_t(c.n_hyp) * 4, cudaMemcpyDeviceToHost, c.stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  for (int i = 0; i < n; i++) slots[i] = raw_slots[grasps[i].image_id];
  DevBuf d_slots, d_img3, d_desc;
  // images: [n own][n x 2 per camera] -> descriptor rows reordered to (own, camera 1, camera 2) per hypothesis
  if (d_slots.reserve(size_t(n) * 4) || d_img3.reserve(size_t(n) * 3 * AG_IMAGE_WORDS * 4 + 64) ||
      d_desc.reserve(size_t(n) * 3 * AG_HOG_DIM * 4))
    return AG_ERR_CUDA;
  AG_CUDA_CHECK(cudaMemcpyAsync(d_slots.p, slots.data(), size_t(n) * 4, cudaMemcpyHostToDevice, c.stream));
  uint32_t* img_cam = d_img3.as<uint32_t>() + size_t(n) * AG_IMAGE_WORDS;
  int rc = camera_images_device(&c, c.n_samples, d_slots.as<int>(), n, img_cam);
  if (rc == AG_OK) rc = hog_descriptors_device(&c, c.images_raw.as<uint32_t>(), d_slots.as<int>(), n, d_desc.as<float>());
  if (rc == AG_OK) rc = hog_descriptors_device(&c, img_cam, nullptr, 2 * n, d_desc.as<float>() + size_t(n) * AG_HOG_DIM);
  if (rc == AG_OK) {
    const size_t row = size_t(AG_HOG_DIM) * 4;
    // row 3 i = own image, 3 i + 1 / 3 i + 2 = camera 1 / 2 only
    cudaMemcpy2DAsync(features, 3 * row, d_desc.p, row, row, n, cudaMemcpyDeviceToHost, c.stream);
    cudaMemcpy2DAsync(reinterpret_cast<char*>(features) + row, 3 * row, d_desc.as<char>() + size_t(n) * row, 2 * row, row, n,
                      cudaMemcpyDeviceToHost, c.stream);
    cudaMemcpy2DAsync(reinterpret_cast<char*>(features) + 2 * row, 3 * row, d_desc.as<char>() + size_t(n) * row + row, 2 * row,
                      row, n, cudaMemcpyDeviceToHost, c.stream);
    if (cudaStreamSynchronize(c.stream) != cudaSuccess) {
      set_error("ag_train_features: copy failed");
      rc = AG_ERR_CUDA;
    }
  }
  d_slots.release();
  d_img3.release();
  d_desc.release();
  return rc;
}

int ag_get_normals(ag_ctx* h, double* normals3n, int n) {
  if (!h || !normals3n || n < 0) return AG_ERR_INVALID;
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  if (n > c.n_vox || c.normals.cap < size_t(n) * 24) {
    set_error("ag_get_normals: n exceeds the voxelised cloud of the last call");
    return AG_ERR_INVALID;
  }
  if (n > 0) AG_CUDA_CHECK(cudaMemcpyAsync(normals3n, c.normals.p, size_t(n) * 24, cudaMemcpyDeviceToHost, c.stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  return AG_OK;
}

int ag_get_images(ag_ctx* h, uint32_t** bits, int* n_images) {
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  *bits = nullptr;
  *n_images = 0;
  if (!c.images_valid) {
    set_error("no grasp images resident");
    return AG_ERR_INVALID;
  }
  const int n = c.n_hyp;
  std::vector<int> raw_slots(n);
  uint32_t* out = static_cast<uint32_t*>(std::malloc(std::max<size_t>(1, size_t(n)) * AG_IMAGE_WORDS * 4));
  if (n > 0) {
    AG_CUDA_CHECK(cudaMemcpy(raw_slots.data(), c.hyp_slots.p, size_t(n) * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++)
      AG_CUDA_CHECK(cudaMemcpy(out + size_t(i) * AG_IMAGE_WORDS,
                               c.images_raw.as<uint32_t>() + size_t(raw_slots[i]) * AG_IMAGE_WORDS, AG_IMAGE_WORDS * 4,
                               cudaMemcpyDeviceToHost));
  }
  *bits = out;
  *n_images = n;
  return AG_OK;
}

// ---- stage-level entry points ----------------------------------------------------------------
int ag_preprocess(ag_ctx* h, const void* points, int stride, int n_in, int size_left, float** xyz_out,
                  int32_t** cam_out, int* n_out) {
  if (!h || !xyz_out || !cam_out || !n_out) return AG_ERR_INVALID;
  *xyz_out = nullptr;
  *cam_out = nullptr;
  *n_out = 0;
  if (n_in <= 0 || size_left == 0 || !points) {
    set_error("Input cloud is empty!");
    return AG_ERR_EMPTY;
  }
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  const size_t bytes = size_t(n_in) * stride;
  if (c.raw.reserve(bytes)) return AG_ERR_CUDA;
  AG_CUDA_CHECK(cudaMemcpyAsync(c.raw.p, points, bytes, cudaMemcpyHostToDevice, c.stream));
  c.launches = 0;
  c.two_cams = size_left < n_in;
  int rc = preprocess_device(&c, c.raw.p, stride, n_in, size_left);
  if (rc) return rc;
  rc = fetch_cloud_size(&c);
  if (rc == AG_RETRY_KEYSORT) {  // lattice too large for the occupancy bitmap: key-sort path from now on
    c.bitmap_ok = false;
    c.state_gen++;
    rc = preprocess_device(&c, c.raw.p, stride, n_in, size_left);
    if (rc) return rc;
    rc = fetch_cloud_size(&c);
  }
  if (rc) return rc;
  const int n = c.n_vox;
  std::vector<GPoint> v(n);
  if (n > 0) AG_CUDA_CHECK(cudaMemcpyAsync(v.data(), c.vox.p, size_t(n) * 16, cudaMemcpyDeviceToHost, c.stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  float* xyz = static_cast<float*>(std::malloc(std::max<size_t>(1, size_t(n)) * 12));
  int32_t* cam = static_cast<int32_t*>(std::malloc(std::max<size_t>(1, size_t(n)) * 4));
  for (int i = 0; i < n; i++) {
    xyz[3 * i] = v[i].x;
    xyz[3 * i + 1] = v[i].y;
    xyz[3 * i + 2] = v[i].z;
    cam[i] = (v[i].tag & kTagCamBit) ? 1 : 0;
  }
  *xyz_out = xyz;
  *cam_out = cam;
  *n_out = n;
  return AG_OK;
}

static int copy_cloud_out(Ctx& c, float** xyz_out, int32_t** cam_out, int* n_out) {
  const int n = c.n_vox;
  std::vector<GPoint> v(n);
  if (n > 0) AG_CUDA_CHECK(cudaMemcpyAsync(v.data(), c.vox.p, size_t(n) * 16, cudaMemcpyDeviceToHost, c.stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  float* xyz = static_cast<float*>(std::malloc(std::max<size_t>(1, size_t(n)) * 12));
  int32_t* cam = static_cast<int32_t*>(std::malloc(std::max<size_t>(1, size_t(n)) * 4));
  for (int i = 0; i < n; i++) {
    xyz[3 * i] = v[i].x;
    xyz[3 * i + 1] = v[i].y;
    xyz[3 * i + 2] = v[i].z;
    cam[i] = (v[i].tag & kTagCamBit) ? 1 : 0;
  }
  *xyz_out = xyz;
  *cam_out = cam;
  *n_out = n;
  return AG_OK;
}

int ag_remove_plane(ag_ctx* h, float** xyz_out, int32_t** cam_out, int* n_out) {
  if (!h || !xyz_out || !cam_out || !n_out) return AG_ERR_INVALID;
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  *xyz_out = nullptr;
  *cam_out = nullptr;
  *n_out = 0;
  c.images_valid = false;
  int rc = remove_plane_device(&c);
  if (rc == AG_RETRY_KEYSORT) {
    set_error("ag_remove_plane: no voxelised cloud (call ag_preprocess / ag_set_cloud first)");
    return AG_ERR_INVALID;
  }
  if (rc) return rc;
  return copy_cloud_out(c, xyz_out, cam_out, n_out);
}

int ag_set_cloud(ag_ctx* h, const float* xyz, const int32_t* cam, int n) {
  if (!h || (n > 0 && !xyz)) return AG_ERR_INVALID;
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  c.images_valid = false;
  c.two_cams = false;
  std::vector<GPoint> v(n);
  for (int i = 0; i < n; i++) {
    v[i].x = xyz[3 * i];
    v[i].y = xyz[3 * i + 1];
    v[i].z = xyz[3 * i + 2];
    v[i].tag = (cam && cam[i]) ? kTagCamBit : 0u;
    if (cam && cam[i]) c.two_cams = true;
  }
  if (c.vox.reserve(std::max<size_t>(16, size_t(n) * 16))) return AG_ERR_CUDA;
  if (n > 0) AG_CUDA_CHECK(cudaMemcpyAsync(c.vox.p, v.data(), size_t(n) * 16, cudaMemcpyHostToDevice, c.stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  return set_cloud_device(&c, n);
}

int ag_radius_search(ag_ctx* h, const float q[3], double radius, int32_t** idx_out, int* n_out) {
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  std::vector<int> res;
  int rc = radius_search_device(&c, q, radius, res);
  if (rc) return rc;
  std::sort(res.begin(), res.end());
  int32_t* o = static_cast<int32_t*>(std::malloc(std::max<size_t>(1, res.size()) * 4));
  std::memcpy(o, res.data(), res.size() * 4);
  *idx_out = o;
  *n_out = int(res.size());
  return AG_OK;
}

int ag_fit_quadrics(ag_ctx* h, const int* indices, int n_indices, double radius, ag_frame* frames_out) {
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  if (n_indices <= 0) return AG_OK;
  for (int i = 0; i < n_indices; i++)
    if (indices[i] < 0 || indices[i] >= c.n_vox) {
      set_error("sample index out of range");
      return AG_ERR_INVALID;
    }
  if (c.samples.reserve(size_t(n_indices) * 4) || c.frames.reserve(size_t(n_indices) * sizeof(ag_frame)) ||
      c.counters.reserve(64))
    return AG_ERR_CUDA;
  AG_CUDA_CHECK(cudaMemsetAsync(c.counters.p, 0, 64, c.stream));
  AG_CUDA_CHECK(cudaMemcpyAsync(c.samples.p, indices, size_t(n_indices) * 4, cudaMemcpyHostToDevice, c.stream));
  RowIndex* ri = c.row_index.as<RowIndex>();
  k_check_samples<<<(n_indices + 255) / 256, 256, 0, c.stream>>>(ri, n_indices, c.samples.as<int>());
  AG_CUDA_CHECK(cudaMemsetAsync(c.frames.p, 0, size_t(n_indices) * sizeof(ag_frame), c.stream));
  int rc = quadric_rand_reset(&c);
  if (rc) return rc;
  rc = fit_quadrics_device(&c, c.samples.as<int>(), n_indices, &ri->n_samples, radius, c.frames.as<ag_frame>(),
                           false);
  if (rc) return rc;
  AG_CUDA_CHECK(cudaMemcpyAsync(frames_out, c.frames.p, size_t(n_indices) * sizeof(ag_frame), cudaMemcpyDeviceToHost,
                                c.stream));
  unsigned long long ctr[8];
  int dev_err = 0;
  AG_CUDA_CHECK(cudaMemcpyAsync(ctr, c.counters.p, 64, cudaMemcpyDeviceToHost, c.stream));
  AG_CUDA_CHECK(cudaMemcpyAsync(&dev_err, &ri->error, 4, cudaMemcpyDeviceToHost, c.stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  if (dev_err & kErrBallOverflow) {
    dev_err &= ~kErrBallOverflow;
    AG_CUDA_CHECK(cudaMemcpy(&ri->error, &dev_err, 4, cudaMemcpyHostToDevice));
    set_error("a radius ball holds more points than a neighbour-pool slot (cloud far denser than a voxelised surface)");
    return AG_ERR_CAPACITY;
  }
  c.timings.n_samples = n_indices;
  c.timings.n_voxels = c.n_vox;
  c.timings.search_ms = elapsed(c.ev_k[0], c.ev_k[1]);
  c.timings.moments_ms = elapsed(c.ev_k[1], c.ev_k[2]);
  c.timings.axes_ms = elapsed(c.ev_k[2], c.ev_k[3]);
  c.timings.taubin_neighbor_points = int64_t(ctr[0]);
  c.timings.taubin_candidates = int64_t(ctr[1]);
  return AG_OK;
}

int ag_hand_sweep(ag_ctx* h, const int* indices, int n_indices, const ag_frame* frames, const double* cloud_normals,
                  unsigned flags, ag_grasp** out, int* n_out) {
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  *out = nullptr;
  *n_out = 0;
  if (n_indices <= 0) return AG_OK;
  for (int i = 0; i < n_indices; i++)
    if (indices[i] < 0 || indices[i] >= c.n_vox) {
      set_error("sample index out of range");
      return AG_ERR_INVALID;
    }
  next_stamp(&c);
  if (c.samples.reserve(size_t(n_indices) * 4) || c.frames.reserve(size_t(n_indices) * sizeof(ag_frame)) ||
      c.counters.reserve(64))
    return AG_ERR_CUDA;
  AG_CUDA_CHECK(cudaMemsetAsync(c.counters.p, 0, 64, c.stream));
  AG_CUDA_CHECK(cudaMemcpyAsync(c.samples.p, indices, size_t(n_indices) * 4, cudaMemcpyHostToDevice, c.stream));
  AG_CUDA_CHECK(cudaMemcpyAsync(c.frames.p, frames, size_t(n_indices) * sizeof(ag_frame), cudaMemcpyHostToDevice,
                                c.stream));
  int rc = set_normals_device(&c, cloud_normals);
  if (rc) return rc;
  RowIndex* ri = c.row_index.as<RowIndex>();
  k_check_samples<<<(n_indices + 255) / 256, 256, 0, c.stream>>>(ri, n_indices, c.samples.as<int>());
  c.n_samples = n_indices;
  rc = hand_sweep_enqueue(&c, c.samples.as<int>(), n_indices, c.frames.as<ag_frame>(), flags & 0x100u);
  if (rc) return rc;
  int n_over = 0;
  AG_CUDA_CHECK(cudaMemcpyAsync(&n_over, hand_sweep_overflow_ptr(&c), 4, cudaMemcpyDeviceToHost, c.stream));
  AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  int Hn = 0;
  rc = hand_sweep_finish(&c, n_indices, n_over, &Hn);
  if (rc) return rc;
  ag_grasp* res = static_cast<ag_grasp*>(std::malloc(std::max<size_t>(1, size_t(Hn)) * sizeof(ag_grasp)));
  if (Hn > 0) {
    k_gather_grasps<<<(Hn + 127) / 128, 128, 0, c.stream>>>(c.grasps_raw.as<ag_grasp>(), c.hyp_slots.as<int>(), Hn,
                                                           c.grasps.as<ag_grasp>(), c.stamp);
    AG_CUDA_CHECK(cudaMemcpyAsync(res, c.grasps.p, size_t(Hn) * sizeof(ag_grasp), cudaMemcpyDeviceToHost, c.stream));
    AG_CUDA_CHECK(cudaStreamSynchronize(c.stream));
  }
  *out = res;
  *n_out = Hn;
  return AG_OK;
}

int ag_sweep_debug(ag_ctx* h, int n_samples, int32_t* slab_counts, int32_t* debug8) {
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  if (slab_counts) AG_CUDA_CHECK(cudaMemcpy(slab_counts, c.sweep_dbg.p, size_t(n_samples) * 4, cudaMemcpyDeviceToHost));
  if (debug8)
    AG_CUDA_CHECK(cudaMemcpy(debug8, c.sweep_dbg.as<int>() + n_samples, size_t(n_samples) * 32, cudaMemcpyDeviceToHost));
  return AG_OK;
}

int ag_hog_svm(ag_ctx* h, const ag_svm* svm, const uint32_t* images, int n, float* descriptors, float* scores) {
  Ctx& c = h->c;
  cudaSetDevice(c.device);
  if (n <= 0) return AG_OK;
  DevBuf img, sc;
  if (img.reserve(size_t(n) * AG_IMAGE_WORDS * 4 + 64) || sc.reserve(size_t(n) * 4)) return AG_ERR_CUDA;
  if (descriptors && c.descriptors.reserve(size_t(n) * AG_HOG_DIM * 4)) return AG_ERR_CUDA;
  AG_CUDA_CHECK(cudaMemcpyAsync(img.p, images, size_t(n) * AG_IMAGE_WORDS * 4, cudaMemcpyHostToDevice, c.stream));
  int rc = hog_svm_device(&c, svm->m, img.as<uint32_t>(), nullptr, n, nullptr, descriptors ? c.descriptors.as<float>() : nullptr,
                          sc.as<float>());
  if (rc == AG_OK) {
    cudaMemcpyAsync(scores, sc.p, size_t(n) * 4, cudaMemcpyDeviceToHost, c.stream);
    if (descriptors)
      cudaMemcpyAsync(descriptors, c.descriptors.p, size_t(n) * AG_HOG_DIM * 4, cudaMemcpyDeviceToHost, c.stream);
    cudaError_t e = cudaStreamSynchronize(c.stream);
    if (e != cudaSuccess) {
      set_error(std::string("hog_svm: ") + cudaGetErrorString(e));
      rc = AG_ERR_CUDA;
    }
  }
  img.release();
  sc.release();
  return rc;
}

}  // extern "C"
