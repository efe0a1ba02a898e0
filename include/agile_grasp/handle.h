// handle.h — the reference's Handle interface (include/agile_grasp/handle.h; src/agile_grasp/handle.cpp:3-73) as a
// thin value carrier: the "average grasp" of a collinear cluster of grasp hypotheses is computed by
// ag_find_handles (include/ag_b200.h, GPU pair predicate + host clustering); this class only exposes the numbers
// through the getters the reference's callers use (grasp_localizer.cpp:149-188, plot.cpp).
#ifndef AGILE_GRASP_HANDLE_H_
#define AGILE_GRASP_HANDLE_H_

#include <vector>

#include "../ag_b200.h"
#include "compat_types.h"
#include "grasp_hypothesis.h"

class Handle {
  enum Field { kCenter = 0, kAxis, kApproach, kBinormal, kHandsCenter, kNumFields };

 public:
  /** hand_list / inliers as in the reference constructor (handle.h, Handle(hand_list, inliers)); `rec` is the
   *  record ag_find_handles produced for this cluster. */
  Handle(const std::vector<GraspHypothesis>& hand_list, const std::vector<int>& inliers, const ag_handle& rec)
      : members_(inliers), hands_(hand_list), mean_width_(rec.width) {
    const double* src[kNumFields] = {rec.center, rec.axis, rec.approach, rec.binormal, rec.hands_center};
    for (int f = 0; f < kNumFields; f++)
      for (int k = 0; k < 3; k++) vec_[f](k) = src[f][k];
  }

  // the reference's accessors
  const Eigen::Vector3d& getCenter() const { return vec_[kCenter]; }            // grasp bottom of the middle inlier
  const Eigen::Vector3d& getAxis() const { return vec_[kAxis]; }                // principal direction of the inlier axes
  const Eigen::Vector3d& getApproach() const { return vec_[kApproach]; }        // approach of the middle inlier
  const Eigen::Vector3d& getBinormal() const { return vec_[kBinormal]; }        // approach x axis
  const Eigen::Vector3d& getHandsCenter() const { return vec_[kHandsCenter]; }  // grasp surface of the middle inlier
  double getWidth() const { return mean_width_; }                               // mean grasp width of the inliers
  const std::vector<int>& getInliers() const { return members_; }
  const std::vector<GraspHypothesis>& getHandList() const { return hands_; }

 private:
  Eigen::Vector3d vec_[kNumFields];
  std::vector<int> members_;
  std::vector<GraspHypothesis> hands_;
  double mean_width_;
};

#endif
