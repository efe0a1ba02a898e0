// localization.h — the reference's `Localization` facade (include/agile_grasp/localization.h:68-349,
// src/agile_grasp/localization.cpp) re-implemented header-only on top of the B200 C ABI (ag_b200.h).
// Same method names, argument meaning and error behaviour: errors print to stdout and return an empty
// vector (localization.cpp:9-15,184-189; learning.cpp:172-191).  Differences, all documented in
// num_threads is accepted and ignored (the GPU path has no thread knob); only
// NO_PLOTTING is honoured; createVisualsPub needs ROS and is not provided; uses_clustering draws its RANSAC
// triples from the library's own generator (PCL's is not reproducible outside PCL);
// points_for_learning stay on the device unless requested.
#ifndef AGILE_GRASP_LOCALIZATION_H_
#define AGILE_GRASP_LOCALIZATION_H_

#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "../ag_b200.h"
#include "compat_types.h"
#include "grasp_hypothesis.h"
#include "handle.h"

typedef pcl::PointCloud<pcl::PointXYZRGBA> PointCloud;

class Localization {
 public:
  Localization() : num_threads_(1), filters_boundaries_(false), plotting_mode_(1), ctx_(nullptr), svm_(nullptr) { init(); }
  Localization(int num_threads, bool filters_boundaries, int plotting_mode)
      : num_threads_(num_threads), filters_boundaries_(filters_boundaries), plotting_mode_(plotting_mode),
        ctx_(nullptr), svm_(nullptr) { init(); }
  ~Localization() {
    if (ctx_) ag_set_svm(ctx_, nullptr);
    if (svm_) ag_svm_free(svm_);
    if (ctx_) ag_destroy(ctx_);
  }
  Localization(const Localization&) = delete;
  Localization& operator=(const Localization&) = delete;

  /** Localize hands in a point cloud (reference: localization.cpp:3-140). `indices` index the voxelised
   *  cloud exactly as in the reference; empty = draw num_samples. */
  std::vector<GraspHypothesis> localizeHands(const PointCloud::Ptr& cloud_in, int size_left,
                                             const std::vector<int>& indices, bool calculates_antipodal,
                                             bool uses_clustering) {
    std::vector<GraspHypothesis> hand_list;
    if (size_left == 0 || !cloud_in || cloud_in->size() == 0) {
      std::cout << "Input cloud is empty!\n" << size_left << std::endl;
      return hand_list;
    }
    if (uses_clustering) std::cout << "Finding point cloud clusters ... \n";  // localization.cpp:53
    if (!ensure_ctx()) return hand_list;
    // the reference removes NaN points from the CALLER's cloud in place (localization.cpp:27)
    std::vector<pcl::PointXYZRGBA>& pts = cloud_in->points;
    ag_grasp* out = nullptr;
    int n = 0;
    int rc = ag_localize(ctx_, pts.data(), int(sizeof(pcl::PointXYZRGBA)), int(pts.size()), size_left,
                         indices.empty() ? nullptr : indices.data(), int(indices.size()),
                         (calculates_antipodal ? AG_FLAG_CALC_ANTIPODAL : 0u) | (uses_clustering ? AG_FLAG_USE_CLUSTERING : 0u),
                         &out, &n);
    size_t w = 0;
    for (size_t i = 0; i < pts.size(); i++)
      if (pts[i].x == pts[i].x && pts[i].y == pts[i].y && pts[i].z == pts[i].z &&
          pts[i].x - pts[i].x == 0 && pts[i].y - pts[i].y == 0 && pts[i].z - pts[i].z == 0) pts[w++] = pts[i];
    pts.resize(w);
    if (rc != AG_OK) {
      std::cout << ag_last_error() << "\n";
      return hand_list;
    }
    hand_list.reserve(n);
    std::vector<float> feats;
    if (keep_train_features_ && n > 0) {  // training path: the three instances' HOG descriptors of every hypothesis
      feats.resize(size_t(n) * 3 * AG_HOG_DIM);
      if (ag_train_features(ctx_, out, n, feats.data()) != AG_OK) {
        std::cout << ag_last_error() << "\n";
        feats.clear();
      }
    }
    for (int i = 0; i < n; i++) {
      GraspHypothesis g(out[i]);
      if (!feats.empty()) g.setTrainFeatures(feats.data() + size_t(i) * 3 * AG_HOG_DIM, size_t(3) * AG_HOG_DIM);
      if (keep_points_) {  // materialise points_for_learning + per-camera index lists (grasp_hypothesis.h:217-220)
        double* p = nullptr;
        int32_t* cm = nullptr;
        int m = 0;
        if (ag_get_points(ctx_, out[i].image_id, &p, &cm, &m) == AG_OK) {
          g.setPointsForLearning(p, cm, m);
          ag_free(p);
          ag_free(cm);
        }
      }
      hand_list.push_back(g);
    }
    ag_free(out);
    std::cout << " # hands: " << hand_list.size() << "\n";
    return hand_list;
  }

  /** Localize hands given two point cloud files (reference: localization.cpp:169-214; the right file name may be
   *  empty).  PCD v0.7 files, DATA ascii / binary / binary_compressed. */
  std::vector<GraspHypothesis> localizeHands(const std::string& pcd_filename_left, const std::string& pcd_filename_right,
                                             bool calculates_antipodal = false, bool uses_clustering = false) {
    std::vector<int> indices(0);
    return localizeHands(pcd_filename_left, pcd_filename_right, indices, calculates_antipodal, uses_clustering);
  }
  std::vector<GraspHypothesis> localizeHands(const std::string& pcd_filename_left, const std::string& pcd_filename_right,
                                             const std::vector<int>& indices, bool calculates_antipodal = false,
                                             bool uses_clustering = false) {
    std::vector<GraspHypothesis> none;
    PointCloud::Ptr cloud(new PointCloud);
    size_t size_left = 0;
    for (int side = 0; side < 2; side++) {
      const std::string& name = side == 0 ? pcd_filename_left : pcd_filename_right;
      if (side == 1 && name.empty()) break;
      void* pts = nullptr;
      int n = 0, w = 0, h = 0;
      if (ag_load_pcd(name.c_str(), &pts, &n, &w, &h) != AG_OK) {
        std::cout << "Couldn't read pcd_filename_left file: " << name << " \n";  // (sic, localization.cpp:186,200)
        return none;
      }
      if (side == 0 && !pcd_filename_right.empty()) std::cout << "Loaded left point cloud with " << w * h << " data points.\n";
      else if (side == 0) std::cout << "Loaded point cloud with " << w * h << " data points.\n";
      else std::cout << "Loaded right point cloud with " << w * h << " data points.\n";
      const pcl::PointXYZRGBA* p = static_cast<const pcl::PointXYZRGBA*>(pts);
      cloud->points.insert(cloud->points.end(), p, p + n);
      ag_free(pts);
      if (side == 0) size_left = cloud->points.size();
    }
    std::cout << "Concatenating point clouds ...\n";
    return localizeHands(cloud, int(size_left), indices, calculates_antipodal, uses_clustering);
  }

  /** Find handles = collinear clusters of grasps (reference: localization.cpp:390-408 -> HandleSearch::findHandles,
   *  handle_search.cpp:4-89).  Plotting modes are not honoured. */
  std::vector<Handle> findHandles(const std::vector<GraspHypothesis>& hand_list, int min_inliers, double min_length) {
    std::vector<Handle> handles;
    if (!ensure_ctx()) return handles;
    std::vector<ag_grasp> recs(hand_list.size());
    for (size_t i = 0; i < recs.size(); i++) {
      recs[i] = hand_list[i].record();
      for (int k = 0; k < 3; k++) {  // (hypotheses built through the reference's own constructor carry no record)
        recs[i].axis[k] = hand_list[i].getAxis()(k);
        recs[i].approach[k] = hand_list[i].getApproach()(k);
        recs[i].bottom[k] = hand_list[i].getGraspBottom()(k);
        recs[i].surface[k] = hand_list[i].getGraspSurface()(k);
      }
      recs[i].width = hand_list[i].getGraspWidth();
    }
    ag_handle* hs = nullptr;
    int32_t* inl = nullptr;
    int nh = 0, ni = 0;
    if (ag_find_handles(ctx_, recs.data(), int(recs.size()), min_inliers, min_length, &hs, &nh, &inl, &ni) != AG_OK) {
      std::cout << ag_last_error() << "\n";
      return handles;
    }
    for (int k = 0; k < nh; k++) {
      std::vector<int> in(inl + hs[k].inlier_offset, inl + hs[k].inlier_offset + hs[k].n_inliers);
      handles.push_back(Handle(hand_list, in, hs[k]));
      std::cout << "handle found with " << in.size() << " inliers\n";
    }
    ag_free(hs);
    ag_free(inl);
    std::cout << "Handle Search\n " << handles.size() << " handles found\n";
    return handles;
  }

  /** Predict antipodal hands with the SVM stored in svm_filename (reference: localization.cpp:142-167,
   *  learning.cpp:165-247).  The model file is parsed once per path and cached. */
  std::vector<GraspHypothesis> predictAntipodalHands(const std::vector<GraspHypothesis>& hand_list,
                                                     const std::string& svm_filename) {
    std::vector<GraspHypothesis> antipodal_hands;
    if (!ensure_ctx()) return antipodal_hands;
    if (!svm_ || svm_path_ != svm_filename) {
      ag_set_svm(ctx_, nullptr);
      if (svm_) ag_svm_free(svm_);
      svm_ = ag_svm_load(svm_filename.c_str());
      svm_path_ = svm_filename;
      if (!svm_) {
        std::cout << " " << ag_last_error() << "\n";
        return antipodal_hands;
      }
      ag_set_svm(ctx_, svm_);  // later localizeHands calls score their hypotheses in the same pass
    }
    std::vector<ag_grasp> recs(hand_list.size());
    for (size_t i = 0; i < recs.size(); i++) recs[i] = hand_list[i].record();
    std::vector<uint8_t> keep(recs.size());
    if (!recs.empty() && ag_classify(ctx_, svm_, recs.data(), int(recs.size()), keep.data()) != AG_OK) {
      std::cout << " " << ag_last_error() << "\n";
      return antipodal_hands;
    }
    for (size_t i = 0; i < recs.size(); i++)
      if (keep[i]) {
        GraspHypothesis g(hand_list[i]);  // the caller's hypothesis (points_for_learning included) ...
        g.record().score = recs[i].score;
        g.record().label = recs[i].label;
        g.setFullAntipodal(true);  // ... marked as learning.cpp:236-243 does
        antipodal_hands.push_back(g);
      }
    std::cout << " " << antipodal_hands.size() << " antipodal grasps found.\n";
    return antipodal_hands;
  }

  void setCameraTransforms(const Eigen::Matrix4d& cam_tf_left, const Eigen::Matrix4d& cam_tf_right) {
    cam_tf_left_ = cam_tf_left;
    cam_tf_right_ = cam_tf_right;
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) {
        params_.cam_tf_left[4 * r + c] = cam_tf_left(r, c);
        params_.cam_tf_right[4 * r + c] = cam_tf_right(r, c);
      }
    dirty_ = true;
  }
  const Eigen::Matrix4d& getCameraTransform(bool is_left) { return is_left ? cam_tf_left_ : cam_tf_right_; }
  void setWorkspace(const Eigen::VectorXd& workspace) {
    for (int i = 0; i < 6 && i < workspace.size(); i++) params_.workspace[i] = workspace(i);
    dirty_ = true;
  }
  void setNumSamples(int num_samples) { params_.num_samples = num_samples; dirty_ = true; }
  // the reference stores but never forwards these two (hand_search.h:85 hard-codes 0.03 / 0.08); kept inert
  void setNeighborhoodRadiusHands(double r) { nn_radius_hands_ = r; }
  void setNeighborhoodRadiusTaubin(double r) { nn_radius_taubin_ = r; }
  void setFingerWidth(double v) { params_.finger_width = v; dirty_ = true; }
  void setHandDepth(double v) { params_.hand_depth = v; dirty_ = true; }
  void setHandOuterDiameter(double v) { params_.hand_outer_diameter = v; dirty_ = true; }
  void setInitBite(double v) { params_.init_bite = v; dirty_ = true; }
  void setHandHeight(double v) { params_.hand_height = v; dirty_ = true; }
  void setSampleSeed(uint64_t seed) { params_.seed = seed; dirty_ = true; }  // B200 addition (App. C.2)
  // B200 addition: the grasp image is rasterised on the device, so the variable-length members of
  // GraspHypothesis are only copied to the host when asked for
  void setKeepPointsForLearning(bool keep) { keep_points_ = keep; }
  // B200 addition (training path, src/nodes/train.cpp): keep the HOG descriptors of the three training instances of
  // every hypothesis (the grasp images live on the device and only for the last call; Learning::train needs them
  // for the hypotheses of ALL training clouds)
  void setKeepTrainingFeatures(bool keep) { keep_train_features_ = keep; }

  static const int NO_PLOTTING = 0;
  static const int PCL_PLOTTING = 1;
  static const int PCL_PLOTTING_FINGERS = 2;
  static const int RVIZ_PLOTTING = 3;

 private:
  void init() {
    ag_default_params(&params_);
    params_.num_threads = num_threads_;
    params_.filters_boundaries = filters_boundaries_ ? 1 : 0;
    cam_tf_left_ = Eigen::Matrix4d::Identity();
    cam_tf_right_ = Eigen::Matrix4d::Identity();
    nn_radius_taubin_ = 0.03;
    nn_radius_hands_ = 0.08;
    dirty_ = true;
  }
  bool ensure_ctx() {
    if (!ctx_) {
      ctx_ = ag_create(0);
      if (!ctx_) {
        std::cout << ag_last_error() << "\n";
        return false;
      }
      dirty_ = true;
    }
    if (dirty_) {
      if (ag_set_params(ctx_, &params_) != AG_OK) {
        std::cout << ag_last_error() << "\n";
        return false;
      }
      dirty_ = false;
    }
    return true;
  }

  Eigen::Matrix4d cam_tf_left_, cam_tf_right_;
  int num_threads_;
  bool filters_boundaries_;
  int plotting_mode_;
  double nn_radius_taubin_, nn_radius_hands_;
  ag_params params_;
  bool dirty_;
  bool keep_points_ = false;
  bool keep_train_features_ = false;
  ag_ctx* ctx_;
  ag_svm* svm_;
  std::string svm_path_;
};

#endif
