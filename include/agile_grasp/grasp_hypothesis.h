// grasp_hypothesis.h — same public interface as the reference's GraspHypothesis
// (include/agile_grasp/grasp_hypothesis.h:46-231), filled from the POD ag_grasp records of the C ABI.
#ifndef AGILE_GRASP_GRASP_HYPOTHESIS_H_
#define AGILE_GRASP_GRASP_HYPOTHESIS_H_

#include <iostream>
#include <vector>

#include "../ag_b200.h"
#include "compat_types.h"

class GraspHypothesis {
 public:
  GraspHypothesis() : cam_source_(-1), grasp_width_(0), full_antipodal_(false), half_antipodal_(false) { rec_.image_id = -1; }

  GraspHypothesis(const Eigen::Vector3d& axis, const Eigen::Vector3d& approach, const Eigen::Vector3d& binormal,
                  const Eigen::Vector3d& bottom, const Eigen::Vector3d& surface, double width,
                  const Eigen::Matrix3Xd& points_for_learning, const std::vector<int>& indices_cam1,
                  const std::vector<int>& indices_cam2, int cam_source)
      : cam_source_(cam_source), points_for_learning_(points_for_learning),
        indices_points_for_learning_cam1_(indices_cam1), indices_points_for_learning_cam2_(indices_cam2), axis_(axis),
        approach_(approach), binormal_(binormal), grasp_bottom_(bottom), grasp_surface_(surface), grasp_width_(width),
        full_antipodal_(false), half_antipodal_(false) {
    rec_.image_id = -1;
  }

  // construction from a record of the C ABI (the path the B200 Localization uses)
  explicit GraspHypothesis(const ag_grasp& g) : cam_source_(g.cam_source), grasp_width_(g.width),
      full_antipodal_(g.full_antipodal != 0), half_antipodal_(g.half_antipodal != 0), rec_(g) {
    for (int k = 0; k < 3; k++) {
      axis_(k) = g.axis[k]; approach_(k) = g.approach[k]; binormal_(k) = g.binormal[k];
      grasp_bottom_(k) = g.bottom[k]; grasp_surface_(k) = g.surface[k];
    }
  }

  void print() {
    auto row = [](const char* n, const Eigen::Vector3d& v) { std::cout << n << v(0) << " " << v(1) << " " << v(2) << std::endl; };
    row("axis: ", axis_); row("approach: ", approach_); row("binormal: ", binormal_);
    std::cout << "grasp width: " << grasp_width_ << std::endl;
    row("grasp surface: ", grasp_surface_); row("grasp bottom: ", grasp_bottom_);
  }

  const Eigen::Vector3d& getApproach() const { return approach_; }
  const Eigen::Vector3d& getAxis() const { return axis_; }
  const Eigen::Vector3d& getBinormal() const { return binormal_; }
  bool isFullAntipodal() const { return full_antipodal_; }
  const Eigen::Vector3d& getGraspBottom() const { return grasp_bottom_; }
  const Eigen::Vector3d& getGraspSurface() const { return grasp_surface_; }
  double getGraspWidth() const { return grasp_width_; }
  bool isHalfAntipodal() const { return half_antipodal_; }
  const std::vector<int>& getIndicesPointsForLearningCam1() const { return indices_points_for_learning_cam1_; }
  const std::vector<int>& getIndicesPointsForLearningCam2() const { return indices_points_for_learning_cam2_; }
  const Eigen::Matrix3Xd& getPointsForLearning() const { return points_for_learning_; }
  int getCamSource() const { return cam_source_; }
  void setFullAntipodal(bool b) { full_antipodal_ = b; rec_.full_antipodal = b; }
  void setHalfAntipodal(bool b) { half_antipodal_ = b; rec_.half_antipodal = b; }
  void setGraspWidth(double w) { grasp_width_ = w; rec_.width = w; }

  // B200 addition: fill the variable-length members from ag_get_points (3 x m column-major + camera per column)
  void setPointsForLearning(const double* pts3xm, const int32_t* cam, int m) {
    points_for_learning_.resize(3, m);
    indices_points_for_learning_cam1_.clear();
    indices_points_for_learning_cam2_.clear();
    for (int j = 0; j < m; j++) {
      for (int r = 0; r < 3; r++) points_for_learning_(r, j) = pts3xm[3 * j + r];
      if (cam[j] == 0) indices_points_for_learning_cam1_.push_back(j);       // rotating_hand.cpp:143-151
      else if (cam[j] == 1) indices_points_for_learning_cam2_.push_back(j);
    }
  }

  // B200 addition (training path): HOG descriptors of the hypothesis' three training instances (own image, camera 1
  // only, camera 2 only; learning.cpp:76-141), 3 x 3528 floats, filled by Localization when asked to keep them
  void setTrainFeatures(const float* f, size_t n) { train_features_.assign(f, f + n); }
  const std::vector<float>& getTrainFeatures() const { return train_features_; }

  // B200 additions: the SVM decision value and the underlying record
  float getScore() const { return rec_.score; }
  const ag_grasp& record() const { return rec_; }
  ag_grasp& record() { return rec_; }

 private:
  int cam_source_;
  Eigen::Matrix3Xd points_for_learning_;
  std::vector<int> indices_points_for_learning_cam1_;
  std::vector<int> indices_points_for_learning_cam2_;
  Eigen::Vector3d axis_, approach_, binormal_, grasp_bottom_, grasp_surface_;
  double grasp_width_;
  bool full_antipodal_, half_antipodal_;
  std::vector<float> train_features_;
  ag_grasp rec_ = ag_grasp();
};

#endif
