// ag_internal.h — host-side context and the launcher prototypes of each stage.
#pragma once

#include <cstdlib>
#include <cuda_runtime.h>

#include <atomic>
#include <string>
#include <vector>

#include "ag_common.cuh"

namespace ag {

void set_error(const std::string& msg);

// growable device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes);
  void release();
  template <typename T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

#define AG_SWEEP_LUT 44  // zones k = -1 .. 20, two per slot step

// constants of the hand model, evaluated on the host with the same libm as the reference
// (finger_hand.cpp:8-15, rotating_hand.cpp:12-15,90, finger_hand.cpp:199-204, antipodal.cpp:14)
struct HandConst {
  double spacing[20];      // finger slot positions
  double finger_width, outer_diameter, depth, hand_height, init_bite;
  double cosv[8], sinv[8]; // cos/sin of the 8 hand orientations
  double bite[12];         // deepening sequence d_t (d_0 = init_bite, d_t = d_{t-1} + 0.005)
  double back[12];         // back_of_hand at d_t = -1.0 * (depth - d_t)
  double lim[12];          // back[t] + depth (points-in-box limit)
  int n_depths;            // number of valid entries (d_t <= depth)
  double cos_thresh;       // cos(20 deg)
  double cam[2][3];        // camera origins
  double img_cell;         // (0.05 - -0.05) / 100
  double half_od;          // outer_diameter / 2.0
  double inv_slot_step;    // 1 / spacing step (zone estimate for the slot-mask table)
  double inv_bite_step;    // 1 / 0.005 (depth-level estimate)
  double inv_img_cell;     // 1 / img_cell (cell estimate, exact division near cell borders)
  // slot masks (side << 32 | in) of a point per zone between two consecutive slot edges: the 40 edges are two
  // interleaved uniform grids (lower edges at integer slot steps from spacing[0], upper edges lut_phi further),
  // so zone (k, a|b) = floor and fraction of (x - spacing[0]) / step; entry 2 (k + 1) + (fraction > lut_phi)
  unsigned long long lut[AG_SWEEP_LUT];
  float lut_phi;
  int lut_ok;              // 0: hand geometry without the table (every point takes the exact comparisons)
};

struct SvmModel {
  int kernel = 0;  // 0 linear, 1 poly
  int degree = 0;
  double gamma = 1, coef0 = 0, rho = 0;
  int var_count = 0, sv_total = 0, sv_count = 0;
  std::vector<float> sv;      // sv_total x var_count
  std::vector<double> alpha;  // sv_count
  std::vector<int> index;
  // device copies (created lazily per device)
  int device = -1;
  float* d_sv = nullptr;
  float* d_svT = nullptr;  // many-vector models: support vectors regrouped [882 groups of 4][sv_total] (float4)
  double* d_alpha = nullptr;
  int* d_index = nullptr;
};


// what a captured localize graph depends on besides buffer addresses (tracked by g_alloc_gen)
struct GraphKey {
  const void* d_points;
  const void* svm;
  unsigned long long state_gen;
  int stride, n_in, size_left, S, given;
  unsigned flags;
};
extern std::atomic<unsigned long long> g_alloc_gen;  // process-wide: any (re)allocation retires every cached graph
struct GraphSlot {
  GraphKey key;
  bool valid = false;
  unsigned long long allocgen = 0, last_use = 0;
  cudaGraphExec_t exec = nullptr;
  int launches = 0;
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  ag_params params;
  HandConst hand;
  // raw input + preprocessing
  DevBuf raw, keys, keys_sorted, keys_unique, cub_tmp, block_counts, misc;
  DevBuf bitmap, tile_state;  // occupancy-bitmap voxelisation: one bit per lattice cell, tile states of its scan
  bool bitmap_ok = true;      // false once a cloud's lattice did not fit: the context stays on the key-sort path
  void* h_pinned = nullptr;  // pinned staging for inputs / outputs
  size_t h_pinned_cap = 0;
  // voxelised cloud (API index space = the reference's voxel order) and its x-row index
  DevBuf vox;        // GPoint: xyz + tag
  DevBuf row_ptr;    // per camera: first voxel of every x-row
  DevBuf col_ptr;    // dense (kx, ky) column table (first voxel of every lattice column), when it fits
  DevBuf row_index;  // RowIndex descriptor (device resident; counts never round-trip through the host)
  DevBuf normals;    // double x 3 per voxel point (cloud_normals_)
  int n_vox = 0;     // host copy, valid after fetch_cloud_size / the end of ag_localize
  int n_cap = 0;     // upper bound on the voxel count known to the host (number of input points)
  // results land in pinned, device-mapped host memory (written by the export kernel)
  void* h_out = nullptr;
  void* d_out_mapped = nullptr;
  size_t h_out_cap = 0;
  void* d_export = nullptr;   // caller-registered device buffer (ag_set_export_buffer)
  size_t d_export_cap = 0;
  // peer gather (ag_gather_*): this rank's buffer, every rank's buffer as mapped through CUDA IPC
  void* gather_buf = nullptr;
  void* gather_peer[AG_MAX_GATHER_RANKS] = {nullptr};
  void* gather_done = nullptr;
  void* gather_zero = nullptr;        // device int that stays 0
  void* gather_merged = nullptr;      // merged list of all ranks (device), in the reference's sample-major order
  void* gather_host = nullptr;        // mapped host copy: GatherHost header + merged records
  void* gather_host_dev = nullptr;
  int gather_cap_total = 0, gather_rec_cap = 0, gather_mask_cap = 0, gather_plan_cap = 0;
  void* gather_plan = nullptr;        // layout of the merged list (k_gather_plan -> k_gather_place)
  bool gather_pending = false;        // a merge of this call is in the stream
  bool gather_valid = false;          // gather_host describes the last completed call
  int pend_slot_first = 0, pend_slot_step = 1;  // position of this context's samples in the full sample list
  std::vector<int> sample_host;       // interleaved share of caller-supplied indices
  size_t gather_slot_bytes = 0;
  int gather_world = 0, gather_rank = 0;
  unsigned gather_epoch = 0;
  bool gather_connected = false;
  // samples
  DevBuf samples, sample_stage, moments, frames, nn_counts, all_frames;
  DevBuf samples_all, nn_counts_all, count_all;  // sharded call in the production normal mode: the full sample list
  // CUDA graph of the localize pipeline (second call with the same shapes captures, later calls replay)
  // state between localize_begin and localize_end
  bool pend_active = false;
  int pend_S = 0;
  int* pend_nsel = nullptr;
  alignas(8) unsigned char pend_peer[192];
  unsigned long long batch_parent_gen = ~0ull;  // child lane of ag_localize_batch: parent state it mirrors
  GraphSlot gslots[4];
  unsigned long long g_tick = 0, state_gen = 0;
  int n_graph_replays = 0;
  bool capturing = false;  // the stream is being captured: timing events must be recorded as external events
  // non-deterministic normal mode: glibc rand() stream, per-sample stream offsets, the carry across launches, bound on the
  // number of samples that may already have consumed draws in the current call
  DevBuf rand_raw, rand_off, rand_carry, picks;
  DevBuf quad_par;             // quadric parameters per sample (k_taubin_solve -> k_axes_finish)
  cudaStream_t stream2 = nullptr;  // side branch of the pipeline (rank selection next to the eigen-solve)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_fork0 = nullptr;
  size_t rand_count = 0;
  int rand_consumed_bound = 0;
  int rand_slot = 0;  // word of rand_carry holding the current carry (0 or 2: k_rank_picks reads one, writes the other)
  DevBuf nbr_heads;  // per sample of the chunk: sample xyz + neighbour count (float4)
  DevBuf nbr_pool;   // neighbour lists of the current chunk of samples: stride x 16-byte records per sample
  bool two_cams = false;  // the last cloud had points of both cameras (sizes the neighbour pool)
  int n_samples = 0;
  // sweep outputs
  DevBuf grasps_raw, valid, images_raw, hyp_slots, grasps, counters, scores, descriptors, kvals, sweep_dbg, overflow;
  DevBuf hyp_list;             // unordered hypothesis slots (sweep -> scorer)
  DevBuf overflow_list;        // [0] count, [1..] samples whose slab exceeded the common kernel's capacity (inline large-slab pass)
  bool inline_big = false;     // the large-slab pass is part of the pipeline (switched on by the first call that needs it)
  DevBuf sample_q;             // per sample of the last fit: x, y, z, (index << 1 | camera) — read by the sweep instead of
                               // the indices -> cloud chain of dependent loads
  DevBuf block_flags;          // per hypothesis: which HOG blocks hold an outline pixel (k_hog_svm -> k_svm_sparse)
  bool scores_by_slot = false; // c->scores is indexed by raw slot (fused linear scoring) instead of hypothesis number
  DevBuf handle_in, handle_bits;  // ag_find_handles: grasp records and the n x n inlier bit matrix
  int n_hyp = 0;
  bool images_valid = false;
  unsigned serial = 0, call_gen = 0;  // context number and call counter behind the record stamp
  uint8_t stamp = 0;                  // ag_grasp.reserved of every record of the last localize / sweep call
  bool sweep_from_fit = false;       // the last sweep's samples came with sample_q records
  unsigned sweep_flags = 0;          // arguments of the last hand_sweep_enqueue (for the overflow re-run)
  const int* sweep_indices = nullptr;
  const ag_frame* sweep_frames = nullptr;
  SvmModel* attached_svm = nullptr;  // ag_set_svm: score inside ag_localize
  bool scores_valid = false;         // last_grasps carry scores of attached_svm
  bool keep_points = false;
  // results of the fused scoring of the last ag_localize (what ag_classify returns for the attached model)
  std::vector<float> last_scores;
  std::vector<uint8_t> last_labels;
  ag_timings timings;
  bool timings_pending = false, timings_h2d = false;  // stage times are read from the events on demand
  cudaEvent_t ev[10];
  cudaEvent_t ev_k[4];   // around k_ball_search / k_taubin_moments / k_taubin_axes
  int launches = 0;      // own-kernel launch counter (reset per localize call)
  bool stage_timing = false;  // record the per-stage events (ag_set_stage_timing)
  // ag_localize's pipeline: small resets and the sample draw ride on kernels of the voxelisation instead of being graph
  // nodes of their own.  fold_resets: bit 0 counters, bit 1 sweep overflow counter, bit 2 rand() carry, bit 3 counter of the inline large-slab pass — set by the
  // caller of preprocess_device, each bit cleared by the stage that would otherwise issue the memset.
  unsigned fold_resets = 0;
  void* fold_draw = nullptr;   // DrawArgs* (host) for the draw to fold into the voxel scan; draw_folded = it was
  bool draw_folded = false;
};

int ctx_pinned(Ctx* c, size_t bytes);

// timing events: inside a stream capture a plain cudaEventRecord only marks a dependency; the external flavour
// becomes an event-record node that is executed (and can be timed) at every graph launch
inline void record_event(Ctx* c, cudaEvent_t ev) {
  if (!c->stage_timing) return;  // only the call's begin / end events (ag_set_stage_timing)
  if (c->capturing) cudaEventRecordWithFlags(ev, c->stream, cudaEventRecordExternal);
  else cudaEventRecord(ev, c->stream);
}

// ---- stage launchers (each returns AG_OK or an error code); all work is enqueued on c->stream
int preprocess_device(Ctx* c, const void* d_points, int stride, int n_in, int size_left);  // -> c->vox (no sync)
int fetch_cloud_size(Ctx* c);  // waits for the stream and reads the voxel count / error flags
int set_cloud_device(Ctx* c, int n);  // cloud already in c->vox (ag_set_cloud)
int set_normals_device(Ctx* c, const double* h_normals);  // cloud_normals_ supplied by the caller
// d_count: device int holding the number of valid entries of d_indices (<= n, the launch bound)
// a context that fits only a share of a call's samples (sample sharding) in the production normal mode: the full
// sample list, so that the rand() stream is sliced as in the unsharded call; this share = first + j * step
struct RandShare {
  const int* d_all;          // device: all sample indices of the call
  const int* d_count_all;    // device int: how many of them are valid
  int n_all, first, step;
};
int fit_quadrics_device(Ctx* c, const int* d_indices, int n, const int* d_count, double radius, ag_frame* d_frames,
                        bool write_normals, const RandShare* share = nullptr);
// restarts the rand() stream of the non-deterministic normal mode (start of every ag_localize / ag_fit_quadrics)
int quadric_rand_reset(Ctx* c);
// enqueue only (no sync): sweep + stable compaction; the hypothesis count stays in device memory
int hand_sweep_enqueue(Ctx* c, const int* d_indices, int n, const ag_frame* d_frames, unsigned flags,
                       bool fork_compact = false, bool frames_from_fit = false, bool inline_big = false);
int* hand_sweep_list_ptr(Ctx* c);        // unordered list of hypothesis slots of the last sweep
int* hand_sweep_list_count_ptr(Ctx* c);
// after a sync: handles samples whose slab overflowed the small instantiation; returns the hypothesis count
int hand_sweep_finish(Ctx* c, int n, int n_over, int* n_hyp);
int hand_sweep_rerun_enqueue(Ctx* c, int n, int n_over, bool fresh_list);  // enqueue only (ag_localize's overflow pass)
int box_points_device(Ctx* c, int n_samples, int slot, std::vector<double>& pts, std::vector<int>& cam);
int* hand_sweep_count_ptr(Ctx* c, int n);     // device address of the hypothesis count of the last enqueue
int* hand_sweep_overflow_ptr(Ctx* c);         // device address of the overflow counter
// n_dev (may be null): device int with the number of hypotheses; n is then only the launch bound
// score_by_slot: d_scores is indexed by the image slot instead of the position in d_image_ids (single-vector models)
int hog_svm_device(Ctx* c, SvmModel* svm, const uint32_t* d_images, const int* d_image_ids, int n, const int* n_dev,
                   float* d_descriptors, float* d_scores, ag_grasp* d_grasps_out = nullptr, bool score_by_slot = false);
int radius_search_device(Ctx* c, const float q[3], double radius, std::vector<int>& out);
// training-data path
int hog_descriptors_device(Ctx* c, const uint32_t* d_images, const int* d_image_slots, int n, float* d_descriptors);
int camera_images_device(Ctx* c, int n_samples, const int* d_slots, int n, uint32_t* d_images);
// uses_clustering: removes the dominant RANSAC plane from the voxelised cloud (waits for the stream; the cloud is
// re-indexed).  Returns AG_RETRY_KEYSORT if the voxelisation asked for the key-sort path.
int remove_plane_device(Ctx* c);

void compute_hand_const(const ag_params& p, HandConst& h);
int svm_to_device(SvmModel* svm, int device);

}  // namespace ag

// the opaque handle of the C ABI
struct ag_ctx {
  ag::Ctx c;
  std::vector<ag_ctx*> children;  // extra lanes of ag_localize_batch (same device, own stream and buffers)
};
namespace ag {
inline Ctx& ctx_of(ag_ctx* h) { return h->c; }
}  // namespace ag
