// oracle_internal.h — declarations shared by the oracle's translation units.
// TEST INFRASTRUCTURE ONLY (see ag_oracle.h).
#ifndef AG_ORACLE_INTERNAL_H_
#define AG_ORACLE_INTERNAL_H_
#include <string>
#include <vector>

#include "ag_oracle.h"

namespace ago {
extern thread_local std::string g_err;
int fail(const std::string& msg);

struct Tree;

struct Hands {
  std::vector<ag_grasp> grasps;
  std::vector<std::vector<double>> pts;    // 3 x m column-major per grasp (points_for_learning)
  std::vector<std::vector<int32_t>> pcam;  // cam source per column
  std::vector<int32_t> status, hand_idx, depth_steps, finger_mask, num_slab;
  int n_voxels = 0;
};

int preprocess(const void* points, int stride, int n_in, int size_left, const ag_params& P, bool use_std_set,
               std::vector<float>& xyz_out, std::vector<int32_t>& cam_out);
Hands* find_hands(const float* xyz, const int32_t* cam, int n, const Tree* tree, const int* indices, int S,
                  const ag_frame* frames, const int32_t* sample_cam, const double* cloud_normals, const ag_params& P);
int fit_quadrics(const float* xyz, const int32_t* cam, int n, const Tree* tree, const int* indices, int S,
                 double radius, const ag_params& P, int sum_perm, ag_frame* frames, double* params_out,
                 double* MN_out, double* eig_out, size_t* rand_consumed = nullptr);
void filter_hands(const ag_grasp* g, int n, const ag_params& P, uint8_t* keep);
int draw_samples(int n, int S, uint64_t seed, int32_t* out);
const Tree* tree_of(const ago_tree* t);
}  // namespace ago
#endif
