/*
 * ag_b200.h — C ABI of the B200-native grasp-hypothesis hot path.
 *
 * This is the drop-in boundary for agile_grasp's `Localization::localizeHands` +
 * `Localization::predictAntipodalHands` path.  Everything is POD: plain pointers, sizes,
 * int status codes; no C++/torch/Eigen/PCL types cross it.  The C++ shim in
 * include/agile_grasp/localization.h (same class/method names as the reference) sits on top
 * of exactly these entry points.
 *
 * Reference interfaces replaced (paths relative to the reference repo):
 *   ag_localize            <- Localization::localizeHands(cloud,size_left,indices,...)
 *                             include/agile_grasp/localization.h:115-116, src/agile_grasp/localization.cpp:3-140
 *   ag_classify            <- Localization::predictAntipodalHands -> Learning::classify
 *                             include/agile_grasp/localization.h:104-105, src/agile_grasp/learning.cpp:165-247
 *   ag_svm_load            <- CvSVM::load call in src/agile_grasp/learning.cpp:181-185 (same on-disk YAML)
 *   ag_set_params          <- Localization setters include/agile_grasp/localization.h:148-259
 *   ag_preprocess          <- NaN removal / filterWorkspace / voxelizeCloud, localization.cpp:17-45,216-355
 *   ag_fit_quadrics        <- HandSearch::findQuadrics + Quadric, hand_search.cpp:65-113, quadric.cpp:14-305
 *   ag_hand_sweep          <- HandSearch::findHands(private) + RotatingHand/FingerHand/Antipodal,
 *                             hand_search.cpp:116-206, rotating_hand.cpp:19-177, finger_hand.cpp, antipodal.cpp
 *   ag_hog_svm             <- Learning::convertToImage + cv::HOGDescriptor::compute + CvSVM::predict,
 *                             learning.cpp:194-226,320-365
 *   ag_find_handles        <- Localization::findHandles -> HandleSearch::findHandles + Handle,
 *                             localization.cpp:390-408, handle_search.cpp:4-118, handle.cpp:3-73
 *   ag_load_pcd            <- pcl::io::loadPCDFile<pcl::PointXYZRGBA> in the file overloads, localization.cpp:169-214
 *   ag_train_features      <- Learning::train* / convertData feature extraction, learning.cpp:76-163,249-290;
 *                             AG_FLAG_USE_CLUSTERING <- uses_clustering plane removal, localization.cpp:51-98
 *   ag_localize_batch, ag_gather_*, ag_set_export_buffer, ag_params.shard_*  (no reference counterpart:
 *                             batches of clouds, multi-GPU exchange of the grasp list, sample sharding)
 *
 * Error model: every function returns 0 on success, <0 on error; ag_last_error() gives the
 * message (thread-local).  The reference's "print and return an empty vector" behaviour
 * (localization.cpp:9-15) is reproduced by the C++ shim on top of these codes.
 *
 * There is NO CPU fallback behind this ABI: if no CUDA device is usable, ag_create fails.
 */
#ifndef AG_B200_H_
#define AG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AG_OK 0
#define AG_ERR_INVALID (-1)
#define AG_ERR_CUDA (-2)
#define AG_ERR_IO (-3)
#define AG_ERR_CAPACITY (-4)
#define AG_ERR_EMPTY (-5)

#define AG_IMAGE_COLS 100        /* Learning::num_horizontal_cells_, learning.h:64 */
#define AG_IMAGE_ROWS 80         /* Learning::num_vertical_cells_ */
#define AG_IMAGE_WORDS 250       /* 100*80 bits packed in uint32, bit (row*100+col) */
#define AG_HOG_DIM 3528          /* 2 windows x 1764, learning.cpp:222 */
#define AG_NUM_ORIENTATIONS 8    /* rotating_hand.cpp:13 */

/* flags for ag_localize */
#define AG_FLAG_CALC_ANTIPODAL 1u   /* calculates_antipodal: all-points r=0.01 normals pass */
#define AG_FLAG_KEEP_POINTS 2u      /* also materialise points_for_learning on the host */
#define AG_FLAG_USE_CLUSTERING 4u   /* uses_clustering: remove the dominant (RANSAC) plane before sampling, localization.cpp:51-98 */

/* Parameters = the union of Localization's setters (localization.h:148-259), HandSearch's
 * hard-coded radii (hand_search.h:85, hand_search.cpp:20) and the Localization ctor args. */
typedef struct ag_params {
  double finger_width;         /* 0.01  find_grasps.cpp:13 */
  double hand_outer_diameter;  /* 0.09 */
  double hand_depth;           /* 0.06 */
  double hand_height;          /* 0.02 */
  double init_bite;            /* 0.01 */
  double workspace[6];         /* xmin xmax ymin ymax zmin zmax */
  double cam_tf_left[16];      /* row-major 4x4; only the translation column is used */
  double cam_tf_right[16];
  double nn_radius_taubin;     /* 0.03 */
  double nn_radius_hands;      /* 0.08 */
  double nn_radius_normals;    /* 0.01, all-points pass when calculates_antipodal */
  double voxel_size;           /* 0.003 localization.cpp:43 */
  int32_t num_samples;         /* used when no explicit indices are given */
  int32_t num_threads;         /* CPU oracle only; ignored by the GPU path */
  int32_t deterministic_normals; /* 1 = Quadric(is_deterministic=true) (parity mode) */
  int32_t filters_boundaries;  /* Localization ctor arg, localization.h:84 */
  int32_t fix_cam_source;      /* 0 reproduces the reference's pre-NaN camera labelling quirk */
  int32_t shard_interleave;    /* sample sharding: 0 = contiguous ranges, 1 = shard i takes samples i, i + n, i + 2n ... */
  uint64_t seed;               /* sample-index RNG seed when indices are not supplied */
  /* Sample sharding of ONE cloud over several GPUs: this context handles share shard_index of
   * shard_count of the samples (drawn or explicit) — a contiguous range, or with shard_interleave every
   * shard_count-th sample (hypotheses cluster on objects, so interleaving balances the shards); the cloud
   * itself is voxelised on every rank.  ag_grasp.sample_slot is the position in the FULL sample list, so the
   * shards' lists merge into the unsharded list by (sample_slot, orientation) — which is what the peer gather
   * (ag_gather_*) does; for contiguous ranges that is plain concatenation in shard order.  0 / 1 = no sharding. */
  int32_t shard_index;
  int32_t shard_count;
} ag_params;

/* One grasp hypothesis = GraspHypothesis (grasp_hypothesis.h:46-231) minus the variable-length
 * members, plus bookkeeping.  `center`/`surface_center`/`width` of Grasp.msg map to
 * bottom / surface / width (grasp_localizer.cpp:137-146). */
typedef struct ag_grasp {
  double axis[3];
  double approach[3];
  double binormal[3];
  double bottom[3];
  double surface[3];
  double width;
  float score;            /* SVM decision value sum (label +1 <=> score <= 0); NaN until classified */
  int32_t sample_index;   /* index of the sample in the voxelised cloud */
  int32_t sample_slot;    /* position of the sample in the sample list */
  int32_t orientation;    /* 0..7, angle = -pi + orientation*pi/4 */
  int32_t cam_source;
  int32_t num_points;     /* number of points_for_learning (box points) */
  int32_t image_id;       /* handle of the device/host grasp image of this hypothesis */
  uint8_t half_antipodal;
  uint8_t full_antipodal;
  uint8_t label;          /* 1 if the SVM says antipodal (kept by classify) */
  uint8_t reserved;       /* call stamp (non-zero): ties the record to the ag_localize / ag_hand_sweep call whose
                             grasp images image_id addresses; ag_classify rejects records of any other call */
} ag_grasp;

/* Per-sample local frame = the outputs of Quadric that are used downstream. */
typedef struct ag_frame {
  double normal[3];
  double axis[3];      /* curvature axis */
  double binormal[3];
  int32_t num_neighbors;
  int32_t majority_cam;
} ag_frame;

/* "Average grasp" of a handle = collinear cluster of grasp hypotheses: class Handle, handle.h / handle.cpp:3-73 */
typedef struct ag_handle {
  double axis[3];          /* Handle::getAxis: principal direction of the inliers' axes */
  double center[3];        /* getCenter: grasp bottom of the inlier nearest the middle of the handle */
  double approach[3];      /* getApproach of that inlier */
  double binormal[3];      /* approach x axis */
  double hands_center[3];  /* getHandsCenter: grasp surface of that inlier */
  double width;            /* mean grasp width of the inliers */
  int32_t n_inliers;       /* inliers of this handle: inlier list [inlier_offset, inlier_offset + n_inliers) */
  int32_t inlier_offset;
} ag_handle;

typedef struct ag_timings {
  float h2d_ms, preprocess_ms, grid_ms, normals_all_ms, quadric_ms, sweep_ms, hog_svm_ms, d2h_ms, total_ms;
  int32_t n_in, n_voxels, n_samples, n_hyp;
  int64_t taubin_neighbor_points;   /* sum over samples of n_T(s): algorithmic bytes = 16 * this */
  int64_t hand_neighbor_points;     /* sum over samples of n_H(s) */
  int64_t taubin_candidates;        /* candidate points staged by the radius search (rows x chord) */
  int64_t hand_candidates;
  float moments_ms;                 /* k_taubin_moments alone (the roofline-graded kernel), last call */
  float axes_ms;                    /* k_taubin_axes alone */
  int32_t kernel_launches;          /* launches of this library's own kernels in the last localize(+classify) */
  float search_ms;                  /* k_ball_search alone (radius search -> neighbour lists of the Taubin fit) */
} ag_timings;

typedef struct ag_ctx ag_ctx;
typedef struct ag_svm ag_svm;

const char* ag_last_error(void);
void ag_default_params(ag_params* p);

ag_ctx* ag_create(int device);
void ag_destroy(ag_ctx* ctx);
int ag_set_params(ag_ctx* ctx, const ag_params* p);
int ag_get_params(ag_ctx* ctx, ag_params* p);
int ag_get_timings(ag_ctx* ctx, ag_timings* t);
/* Per-stage times in ag_timings (preprocess_ms ... axes_ms) come from a dozen extra event records inside the
 * pipeline (about 16 us per call as CUDA-graph nodes); off by default: total_ms and the counters are always
 * filled, the stage fields read 0.  (The reference has no counterpart: it prints wall-clock times per stage,
 * localization.cpp:100-135.) */
int ag_set_stage_timing(ag_ctx* ctx, int on);
void ag_free(void* p);

/* SVM model in OpenCV-2.4 YAML ("!!opencv-ml-svm"), LINEAR or POLY kernels. */
ag_svm* ag_svm_load(const char* path);
void ag_svm_free(ag_svm* svm);
int ag_svm_info(const ag_svm* svm, int* kernel_type, int* var_count, int* sv_total, double* rho);

/* Full path, host buffers in -> host grasp list out.
 * points: n_in records, `stride` bytes apart, x,y,z float32 at byte offsets 0,4,8
 * (pcl::PointXYZRGBA has stride 32).  indices: sample indices into the VOXELISED cloud
 * (may be NULL/0 -> num_samples are drawn).  The grasp images stay resident on the device
 * for a following ag_classify.  *out is malloc'ed; release with ag_free. */
int ag_localize(ag_ctx* ctx, const void* points, int stride, int n_in, int size_left,
                const int* indices, int n_indices, unsigned flags, ag_grasp** out, int* n_out);

/* Same, but the cloud is already resident in device memory (device pointer, same layout). */
int ag_localize_device(ag_ctx* ctx, const void* d_points, int stride, int n_in, int size_left,
                       const int* indices, int n_indices, unsigned flags, ag_grasp** out, int* n_out);

/* Score hypotheses returned by the last ag_localize on this context (image_id refers to the
 * device-resident grasp images).  Fills score/label in place; `keep` (may be NULL) gets 1 for
 * hypotheses the reference would return from predictAntipodalHands. */
int ag_classify(ag_ctx* ctx, const ag_svm* svm, ag_grasp* grasps, int n, uint8_t* keep);

/* A batch of clouds on one GPU (BASELINE config 4; the reference has no batch call — its node handles one
 * cloud per second, grasp_localizer.cpp:82).  Up to AG_BATCH_LANES clouds are in flight at a time on separate
 * streams (the context plus internal child contexts mirroring its parameters and attached SVM); out[i] / n_out[i]
 * are exactly what ag_localize(points[i], ...) with sampled indices returns.  Returns the first error. */
#define AG_BATCH_LANES 4
int ag_localize_batch(ag_ctx* ctx, int n_clouds, const void* const* points, const int* strides, const int* n_in,
                      const int* size_left, unsigned flags, ag_grasp** out, int* n_out);

/* Attach an SVM to the context (NULL detaches): every following ag_localize also scores its hypotheses
 * in the same stream (score/label of the returned records are filled) and a following ag_classify with
 * the same model just returns those results.  The model must outlive the attachment. */
int ag_set_svm(ag_ctx* ctx, const ag_svm* svm);

/* Multi-GPU plumbing: when a device buffer is registered, every ag_localize also leaves
 * [int32 n_hyp, int32 n_vox, int32 n_samples, int32 error][n_hyp x ag_grasp] in it (same stream, complete
 * when ag_localize returns), so the caller can all-gather the grasp list over NCCL without a host copy.
 * bytes must be >= 16 + 8 * samples * sizeof(ag_grasp); NULL unregisters. */
int ag_set_export_buffer(ag_ctx* ctx, void* d_buffer, size_t bytes);

/* Multi-GPU grasp-list exchange WITHOUT a collective call (one process per GPU, all GPUs of one NVLink/
 * NVSwitch box).  Every rank creates a gather buffer (consumer acknowledgements + two epochs x world slots),
 * publishes its CUDA IPC handle, and connects to the handles of all ranks (rank order, its own included).  From
 * then on every ag_localize is a collective step: its export kernel stores this rank's [header][records]
 * directly into slot `rank` of EVERY rank's buffer over NVLink and raises an epoch flag; a merge kernel in the
 * same stream waits (on the device) for all `world` lists of the epoch, merges them by (sample_slot,
 * orientation) into ONE list in the reference's sample-major order (hand_search.cpp:194-200) in device and
 * mapped host memory, and acknowledges the epoch to every producer — a producer never overwrites a slot whose
 * epoch a consumer has not acknowledged (back-pressure instead of "hope the consumer was fast enough").  The one
 * host wait of ag_localize covers the whole exchange.  ag_gather_result returns the merged list of the last
 * call (valid until the next ag_localize); ag_gather_wait (kept for callers that want the per-rank slots)
 * returns the per-rank counts and slot addresses: slot r = d_slots + r * slot_bytes, records start
 * AG_GATHER_SLOT_HEADER bytes into a slot.  All ranks must call ag_localize the same number of times.  This
 * replaces the NCCL all-gather of ag_set_export_buffer (kept for other backends). */
#define AG_MAX_GATHER_RANKS 8
#define AG_IPC_HANDLE_BYTES 64
#define AG_GATHER_SLOT_HEADER 32
size_t ag_gather_slot_bytes(int num_samples);
int ag_gather_create(ag_ctx* ctx, int num_samples, int world, int rank, unsigned char* ipc_handle_out /* 64 B */);
int ag_gather_connect(ag_ctx* ctx, const unsigned char* handles /* world x 64 B, rank order */);
int ag_gather_wait(ag_ctx* ctx, int32_t* n_hyp_per_rank /* world, may be NULL */, const void** d_slots,
                   size_t* slot_bytes);
int ag_gather_result(ag_ctx* ctx, int32_t* n_hyp_per_rank /* world, may be NULL */, int* n_total,
                     const ag_grasp** d_merged, const ag_grasp** h_merged);
int ag_gather_destroy(ag_ctx* ctx);

/* pcl::io::loadPCDFile<pcl::PointXYZRGBA> (localization.cpp:184,198; file overloads :169-214): reads a PCD
 * v0.7 file (DATA ascii | binary | binary_compressed) into malloc'ed pcl::PointXYZRGBA records (32 bytes
 * each, x y z float32 at bytes 0/4/8, rgba uint32 at byte 16) — the layout ag_localize takes with
 * stride 32.  Host only (no device needed).  width/height may be NULL. */
int ag_load_pcd(const char* path, void** points_out, int* n_out, int* width_out, int* height_out);

/* HandleSearch::findHandles (handle_search.cpp:4-89) + Handle (handle.cpp:3-73): the step that follows
 * predictAntipodalHands in every caller (grasp_localizer.cpp:103, src/nodes/test.cpp:97).  The O(n^2) pair
 * predicate (distance from the axis line < 0.01, axis angle and approach angle < 0.34 rad: three acos per
 * pair) is evaluated on the GPU into an n x n bit matrix; the greedy, order-dependent clustering walks the
 * set bits on the host.  hands: n records (host).  Outputs malloc'ed (ag_free): handles and the flat inlier
 * index list.  Deviations from the reference are listed in INTEGRATION.md (eigenvector sign, sort ties). */
int ag_find_handles(ag_ctx* ctx, const ag_grasp* hands, int n, int min_inliers, double min_length,
                    ag_handle** handles_out, int* n_handles, int32_t** inliers_out, int* n_inliers_total);

/* Variable-length members of GraspHypothesis for hypothesis `image_id` of the last ag_localize
 * (requires AG_FLAG_KEEP_POINTS): points_for_learning (3 x m, column-major doubles) and the
 * camera source of each column. */
int ag_get_points(ag_ctx* ctx, int image_id, double** pts3xm, int32_t** cam, int* m);
int ag_get_images(ag_ctx* ctx, uint32_t** bits, int* n_images);   /* AG_IMAGE_WORDS per image */
/* Training-data path (Learning::train -> convertData, learning.cpp:76-163,249-290; caller src/nodes/train.cpp:105-129):
 * HOG descriptors of the three training instances of each hypothesis of the last ag_localize — its grasp image, and
 * the images made from the points of camera 1 only / camera 2 only (createInstance(h, cam_pos [, 0 | 1])).
 * features: n x 3 x AG_HOG_DIM floats, row 3 i + k = instance k of hypothesis i.  Labels are the hypotheses'
 * full_antipodal flags (learning.cpp:379); which hypotheses become instances and the SMO solve (CvSVM::train) stay
 * with the caller — the model file it saves loads through ag_svm_load. */
int ag_train_features(ag_ctx* ctx, const ag_grasp* grasps, int n, float* features);

/* cloud_normals_ (hand_search.cpp:13-26,102) as the last ag_localize / ag_hand_sweep left it: 3 doubles for each
 * of the first n voxels (n <= voxel count); zero where no normal was computed. */
int ag_get_normals(ag_ctx* ctx, double* normals3n, int n);

/* ---- stage-level entry points (used by the parity tests and by callers that want one stage) */

/* NaN removal + workspace filter + voxelisation. Outputs malloc'ed: xyz (3 floats per voxel), cam. */
int ag_preprocess(ag_ctx* ctx, const void* points, int stride, int n_in, int size_left,
                  float** xyz_out, int32_t** cam_out, int* n_out);

/* uses_clustering as a stage (localization.cpp:51-98): removes the dominant RANSAC plane (100 iterations, inlier
 * distance 0.01, refitted coefficients) from the context's current voxelised cloud and re-indexes it; returns the
 * remaining cloud.  AG_ERR_EMPTY if no plane was found (the reference then returns no hands). */
int ag_remove_plane(ag_ctx* ctx, float** xyz_out, int32_t** cam_out, int* n_out);

/* Load an already-voxelised cloud (skips preprocessing; must be in voxel order) and build the row index. */
int ag_set_cloud(ag_ctx* ctx, const float* xyz, const int32_t* cam, int n);

/* Radius search on the current cloud; returns neighbour indices in ascending index order. */
int ag_radius_search(ag_ctx* ctx, const float q[3], double radius, int32_t** idx_out, int* n_out);

/* Taubin quadric fit + local frame for the given sample indices of the current cloud. */
int ag_fit_quadrics(ag_ctx* ctx, const int* indices, int n_indices, double radius, ag_frame* frames_out);

/* Hand sweep for the given samples with caller-supplied frames (normal + curvature axis).
 * cloud_normals: 3 doubles per cloud point (may be NULL = all zero). */
int ag_hand_sweep(ag_ctx* ctx, const int* indices, int n_indices, const ag_frame* frames,
                  const double* cloud_normals, unsigned flags, ag_grasp** out, int* n_out);

/* Diagnostics of the last sweep: slab point count per sample and, per (sample, orientation), a word
 * status | hand_index<<4 | deepening_steps<<8 | finger_mask<<12 (status 0 camera-rejected, 1 no hand,
 * 2 hypothesis). */
int ag_sweep_debug(ag_ctx* ctx, int n_samples, int32_t* slab_counts, int32_t* debug8);

/* HOG descriptor + SVM decision value for n packed grasp images (AG_IMAGE_WORDS uint32 each).
 * descriptors (may be NULL): n x AG_HOG_DIM floats. */
int ag_hog_svm(ag_ctx* ctx, const ag_svm* svm, const uint32_t* images, int n,
               float* descriptors, float* scores);

#ifdef __cplusplus
}
#endif
#endif /* AG_B200_H_ */
