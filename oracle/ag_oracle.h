/*
 * ag_oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is a plain-C++ restatement of the reference's
 * (atenpas/agile_grasp) algorithm for the Localization::localizeHands +
 * predictAntipodalHands hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product (libag_b200.so) never does.
 *
 * PARITY STATUS: "parity unpinned" by the reference's own tests — the reference ships no
 * assertions, golden vectors or input data.  The third-party arithmetic is
 * pinned instead against the executables present in this image: LAPACK dggev_ is the real one
 * (dlopen'ed from the OpenBLAS bundled with cv2 / scipy), HOG and the SVM decision value are
 * checked against cv2 4.13 in tests/test_oracle_hog_svm.py, the neighbour search against a
 * brute-force definition.
 */
#ifndef AG_ORACLE_H_
#define AG_ORACLE_H_

#include "../include/ag_b200.h" /* POD types only (ag_params, ag_grasp, ag_frame) */

#ifdef __cplusplus
extern "C" {
#endif

const char* ago_last_error(void);
void ago_free(void* p);

/* dggev_ provider: path of a shared library exporting `symbol` (LP64 dggev). Returns 0 if found. */
int ago_set_lapack(const char* so_path, const char* symbol);
int ago_have_lapack(void);

/* A.1 preprocess: localization.cpp:17-45,216-355. use_std_set=1 follows the reference's
 * std::set voxelisation literally (timed baseline); 0 uses sort+unique (same result). */
int ago_preprocess(const void* points, int stride, int n_in, int size_left, const ag_params* P,
                   int use_std_set, float** xyz_out, int32_t** cam_out, int* n_out);

/* C.1 radius search (FLANN L2_Simple<float> semantics). method 0 = brute force, 1 = kd-tree.
 * Results sorted ascending by (dist, index). */
typedef struct ago_tree ago_tree;
ago_tree* ago_tree_build(const float* xyz, int n);
void ago_tree_free(ago_tree* t);
int ago_radius_search(const ago_tree* t, const float* xyz, int n, const float q[3], double radius,
                      int method, int32_t** idx_out, float** dist_out, int* n_out);

/* A.3 + A.4: Quadric::fitQuadric + findTaubinNormalAxis for each sample index.
 * Optional outputs (may be NULL): params (10 per sample, the dggev eigenvector as used by
 * findTaubinNormalAxis: a,b,c,d,e,f,g,h,i,j of the implicit quadric), MN (200 per sample: M then
 * N, row-major 10x10), eigvals (10 per sample, alphar/beta).
 * sum_perm: 0 = reference summation order; k>0 = deterministic permutation #k of the neighbour
 * order for the M/N accumulation only (used to measure the reference's own rounding sensitivity);
 * -1 = extended-precision check solve (long double, sample-centred coordinates, no LAPACK): NOT the
 * reference's arithmetic, used to measure how far dggev_ and the CUDA path each are from exact. */
int ago_fit_quadrics(const float* xyz, const int32_t* cam, int n, const ago_tree* tree,
                     const int* indices, int n_indices, double radius, const ag_params* P, int sum_perm,
                     ag_frame* frames_out, double* params_out, double* MN_out, double* eigvals_out);

/* A.5-A.8: HandSearch::findHands (private) for the given samples and frames.
 * cloud_normals: 3 doubles per point (column j = normal of point j), may be NULL (zeros).
 * sample_cam: cam source per sample (hands_cam_source).  The result handle owns grasps, their
 * points_for_learning and per-(sample,rotation) debug info. */
typedef struct ago_hands ago_hands;
ago_hands* ago_find_hands(const float* xyz, const int32_t* cam, int n, const ago_tree* tree,
                          const int* indices, int n_indices, const ag_frame* frames,
                          const int32_t* sample_cam, const double* cloud_normals, const ag_params* P);
void ago_hands_free(ago_hands* h);
int ago_hands_count(const ago_hands* h);
const ag_grasp* ago_hands_grasps(const ago_hands* h);
/* points_for_learning of hypothesis k: 3 x m column-major; cam source per column */
int ago_hands_points(const ago_hands* h, int k, const double** pts, const int32_t** cam, int* m);
/* debug: per (sample_slot*8 + orientation): status (0 cam-rejected, 1 no hand, 2 hypothesis),
 * chosen hand index, number of deepening steps kept, finger mask at final depth */
int ago_hands_debug(const ago_hands* h, const int32_t** status, const int32_t** hand_idx,
                    const int32_t** depth_steps, const int32_t** finger_mask, const int32_t** num_slab);
/* A.9 boundary filter (localization.cpp:364-388): keep[k]=1 if hypothesis survives */
int ago_filter_hands(const ag_grasp* grasps, int n, const ag_params* P, uint8_t* keep);

/* A.10 image: Learning::createInstance + convertToImage for hypothesis k -> 80x100 u8 (0/255) */
int ago_grasp_image(const ago_hands* h, int k, const ag_params* P, uint8_t* image80x100);
/* generic: image from explicit points */
int ago_points_image(const double* pts3xm, int m, const double binormal[3], const double surface[3],
                     const double cam_pos[3], uint8_t* image80x100);

/* training instances (learning.cpp:375-400): image of hypothesis k from all box points (cam = -1) or only from the
 * points of camera 1 / 2 (cam = 0 / 1: the "simulated camera" instances of Learning::train, learning.cpp:76-141) */
int ago_grasp_image_cam(const ago_hands* h, int k, int cam, const ag_params* P, uint8_t* image80x100);

/* uses_clustering (localization.cpp:51-98): RANSAC plane (100 iterations, threshold 0.01, refitted coefficients)
 * removed from a voxelised cloud; keep[i] = 0 for plane inliers.  counts_out (may be NULL): inlier count of every
 * iteration; plane_out (may be NULL): the refitted plane n.p + d = 0.  Returns 1 if no plane was found. */
int ago_remove_plane(const float* xyz, int n, uint64_t seed, int max_iterations, double thresh, uint8_t* keep,
                     int32_t* counts_out, double* plane_out);

/* C.3 HOG as configured at learning.cpp:194-195,220 -> 3528 floats */
int ago_hog(const uint8_t* image80x100, float* desc3528);

/* C.4 SVM */
typedef struct ago_svm ago_svm;
ago_svm* ago_svm_load(const char* path);
void ago_svm_free(ago_svm* s);
int ago_svm_info(const ago_svm* s, int* kernel_type, int* var_count, int* sv_total, double* rho,
                 int* degree, double* gamma, double* coef0);
const float* ago_svm_sv(const ago_svm* s);      /* sv_total x var_count */
const double* ago_svm_alpha(const ago_svm* s);  /* sv_count */
/* decision value exactly as CvSVM::predict(returnDFVal=true): (float)(-rho + sum alpha_k K_k) */
float ago_svm_decision(const ago_svm* s, const float* x);

/* classify all hypotheses of `h`: fills score/label of the grasps inside h; keep[k]=1 iff label==1 */
int ago_classify(ago_hands* h, const ago_svm* s, const ag_params* P, uint8_t* keep);

/* Full path (the timed CPU baseline): preprocess -> tree -> [all-points normals] -> quadrics ->
 * hands -> [boundary filter] -> classify.  indices may be NULL (then drawn from P->seed).
 * times_ms (may be NULL): [preprocess, tree, normals_all, quadrics, hands, classify, total]. */
ago_hands* ago_localize(const void* points, int stride, int n_in, int size_left, const ag_params* P,
                        const int* indices, int n_indices, unsigned flags, const ago_svm* svm,
                        int use_std_set, double* times_ms, int* n_voxels_out);

/* HandleSearch::findHandles + Handle (handle_search.cpp:4-118, handle.cpp:3-73) on grasp records; outputs
 * malloc'ed: handles, flat inlier index list (handle k owns [inlier_offset, inlier_offset + n_inliers)). */
int ago_find_handles(const ag_grasp* hands, int n, int min_inliers, double min_length, ag_handle** handles_out,
                     int* n_handles, int32_t** inliers_out, int* n_inliers_total);

/* the first n outputs of glibc rand() after srand(seed) (restatement used by the non-deterministic normal mode) */
int ago_glibc_rand(uint32_t seed, int n, int32_t* out);

/* deterministic sample draw shared with the product: sorted distinct indices in [0,n) */
int ago_draw_samples(int n, int num_samples, uint64_t seed, int32_t* out);

#ifdef __cplusplus
}
#endif
#endif
