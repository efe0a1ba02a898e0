// handle.h — same public interface as the reference's Handle (include/agile_grasp/handle.h, class Handle;
// src/agile_grasp/handle.cpp:3-73): the "average grasp" of a collinear cluster of grasp hypotheses.  The
// values are computed by ag_find_handles (include/ag_b200.h); this class only carries them.
#ifndef AGILE_GRASP_HANDLE_H_
#define AGILE_GRASP_HANDLE_H_

#include <vector>

#include "../ag_b200.h"
#include "compat_types.h"
#include "grasp_hypothesis.h"

class Handle {
 public:
  Handle(const std::vector<GraspHypothesis>& hand_list, const std::vector<int>& inliers, const ag_handle& h)
      : inliers_(inliers), hand_list_(hand_list), width_(h.width) {
    for (int k = 0; k < 3; k++) {
      center_(k) = h.center[k];
      axis_(k) = h.axis[k];
      approach_(k) = h.approach[k];
      binormal_(k) = h.binormal[k];
      hands_center_(k) = h.hands_center[k];
    }
  }
  const Eigen::Vector3d& getApproach() const { return approach_; }
  const Eigen::Vector3d& getAxis() const { return axis_; }
  const Eigen::Vector3d& getCenter() const { return center_; }
  const Eigen::Vector3d& getHandsCenter() const { return hands_center_; }
  double getWidth() const { return width_; }
  const std::vector<GraspHypothesis>& getHandList() const { return hand_list_; }
  const std::vector<int>& getInliers() const { return inliers_; }
  const Eigen::Vector3d& getBinormal() const { return binormal_; }

 private:
  std::vector<int> inliers_;
  std::vector<GraspHypothesis> hand_list_;
  Eigen::Vector3d center_, axis_, approach_, binormal_;
  double width_;
  Eigen::Vector3d hands_center_;
};

#endif
