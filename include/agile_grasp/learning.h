// learning.h — the reference's `Learning` facade (include/agile_grasp/learning.h:56-161, src/agile_grasp/learning.cpp)
// on top of the B200 path.  What the reference computes per hypothesis — grasp image, HOG descriptor, SVM decision —
// runs on the device (ag_classify / ag_train_features); this class keeps the reference's host-side logic around it:
// which hypotheses become training instances (train / trainBalanced, learning.cpp:3-163), three instances per
// hypothesis (own image, simulated camera 1, simulated camera 2), labels = isFullAntipodal (learning.cpp:379).
//
// The SMO solve is OpenCV's CvSVM::train (learning.cpp:296-315).  With OpenCV 2.4 headers available compile with
// -DAG_HAVE_OPENCV2 and the model is trained and saved exactly like the reference does; without them (this image)
// the training matrix is written next to the requested model file (`<file>.train`: int32 rows, int32 cols,
// rows x cols float32 features, rows float32 labels) for any CvSVM-compatible trainer — tests/ train it with cv2
// and load the result through ag_svm_load.
//
// Hypotheses must carry their training descriptors: Localization::setKeepTrainingFeatures(true) before
// localizeHands (the grasp images live on the device for one call only).
#ifndef AGILE_GRASP_LEARNING_H_
#define AGILE_GRASP_LEARNING_H_

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <set>
#include <string>
#include <vector>

#include "grasp_hypothesis.h"
#include "localization.h"

#ifdef AG_HAVE_OPENCV2
#include <opencv2/ml/ml.hpp>
#endif

class Learning {
 public:
  Learning() : num_horizontal_cells_(100), num_vertical_cells_(80), num_threads_(1), loc_(nullptr) {}
  explicit Learning(int num_threads) : num_horizontal_cells_(100), num_vertical_cells_(80), num_threads_(num_threads), loc_(nullptr) {}

  /** learning.cpp:3-74: per training cloud at most max_positive positives, then as many random negatives. */
  void trainBalanced(const std::vector<GraspHypothesis>& hands_list, const std::vector<int>& sizes,
                     const std::string& file_name, const Eigen::Matrix3Xd& cam_pos, int max_positive = 1000000000,
                     bool is_plotting = false) {
    (void)cam_pos;
    (void)is_plotting;
    std::vector<int> positives, negatives, positives_sub, selected;
    size_t k = 0;
    for (int i = 0; i < int(hands_list.size()); i++) {
      if (hands_list[i].isFullAntipodal()) positives_sub.push_back(i);
      else if (!hands_list[i].isHalfAntipodal()) negatives.push_back(i);
      if (k < sizes.size() && i == sizes[k]) {
        take_positives(positives_sub, max_positive, positives);
        positives_sub.resize(0);
        k++;
      }
    }
    selected = positives;
    std::set<int> pick;
    while (pick.size() < positives.size() && pick.size() < negatives.size()) pick.insert(std::rand() % int(negatives.size()));
    for (int v : pick) selected.push_back(negatives[v]);
    std::cout << "size(positives): " << positives.size() << std::endl;
    std::cout << "indices_selected.size: " << selected.size() << std::endl;
    convert(hands_list, selected, file_name);
  }

  /** learning.cpp:76-141: every hand that is not half antipodal is a negative; positives per cloud limited. */
  void train(const std::vector<GraspHypothesis>& hands_list, const std::vector<int>& sizes, const std::string& file_name,
             const Eigen::Matrix3Xd& cam_pos, int max_positive = 1000000000, bool is_plotting = false) {
    (void)cam_pos;
    (void)is_plotting;
    std::vector<int> positives, instances;
    size_t k = 0;
    for (int i = 0; i < int(hands_list.size()); i++) {
      if (hands_list[i].isFullAntipodal()) positives.push_back(i);
      else if (!hands_list[i].isHalfAntipodal()) instances.push_back(i);
      if (k < sizes.size() && i == sizes[k]) {
        take_positives(positives, max_positive, instances);
        positives.resize(0);
        k++;
      }
    }
    convert(hands_list, instances, file_name);
  }

  /** learning.cpp:143-163: every hand that is not merely half antipodal. */
  void train(const std::vector<GraspHypothesis>& hands_list, const std::string& file_name, const Eigen::Matrix3Xd& cam_pos,
             bool is_plotting = false) {
    (void)cam_pos;
    (void)is_plotting;
    std::vector<int> instances;
    for (int i = 0; i < int(hands_list.size()); i++)
      if (!hands_list[i].isHalfAntipodal() || hands_list[i].isFullAntipodal()) instances.push_back(i);
    convert(hands_list, instances, file_name);
  }

  /** learning.cpp:165-247.  The grasp images live in the Localization that produced hands_list: attach it first. */
  void attach(Localization& loc) { loc_ = &loc; }
  std::vector<GraspHypothesis> classify(const std::vector<GraspHypothesis>& hands_list, const std::string& svm_filename,
                                        const Eigen::Matrix3Xd& cam_pos, bool is_plotting = false) {
    (void)cam_pos;
    (void)is_plotting;
    if (!loc_) {
      std::cout << " Learning::classify: attach(Localization&) the object that localized these hands first\n";
      return std::vector<GraspHypothesis>();
    }
    return loc_->predictAntipodalHands(hands_list, svm_filename);
  }

  /** the training matrix of the last train* call (rows = 3 per selected hypothesis) */
  const std::vector<float>& features() const { return features_; }
  const std::vector<float>& labels() const { return labels_; }

 private:
  // learning.cpp:20-45,105-137: all positives of a cloud, or max_positive of them drawn with std::rand()
  static void take_positives(const std::vector<int>& positives, int max_positive, std::vector<int>& out) {
    if (int(positives.size()) <= max_positive) {
      out.insert(out.end(), positives.begin(), positives.end());
      return;
    }
    std::set<int> idx;
    while (int(idx.size()) < max_positive) idx.insert(std::rand() % int(positives.size()));
    std::cout << positives.size() << " positive examples found\n randomly selected indices:";
    for (int v : idx) {
      std::cout << " " << v;
      out.push_back(positives[v]);
    }
    std::cout << std::endl;
  }

  // Learning::convertData (learning.cpp:249-318): features (3 rows per hypothesis) + labels -> trainer
  void convert(const std::vector<GraspHypothesis>& hands, const std::vector<int>& sel, const std::string& file_name) {
    const int cols = AG_HOG_DIM;
    features_.clear();
    labels_.clear();
    int num_positives = 0;
    for (int idx : sel) {
      const std::vector<float>& f = hands[idx].getTrainFeatures();
      if (f.size() != size_t(3) * cols) {
        std::cout << " hypothesis " << idx << " carries no training descriptors (Localization::setKeepTrainingFeatures)\n";
        continue;
      }
      features_.insert(features_.end(), f.begin(), f.end());
      for (int k = 0; k < 3; k++) labels_.push_back(hands[idx].isFullAntipodal() ? 1.0f : -1.0f);
      if (hands[idx].isFullAntipodal()) num_positives += 3;
    }
    const int rows = int(labels_.size());
    std::cout << "Converting " << rows << " training examples (grasps) to images\n";
#ifdef AG_HAVE_OPENCV2
    cv::Mat F(rows, cols, CV_32FC1, features_.data()), L(rows, 1, CV_32FC1, labels_.data());
    CvSVMParams params;  // learning.cpp:296-311 (uses_linear_kernel = false: POLY degree 2)
    params.svm_type = CvSVM::C_SVC;
    params.kernel_type = CvSVM::POLY;
    params.degree = 2;
    CvSVM svm;
    svm.train(F, L, cv::Mat(), cv::Mat(), params);
    svm.save(file_name.c_str());
    std::cout << "Saved trained SVM as " << file_name << "\n";
#else
    const std::string out = file_name + ".train";
    if (FILE* fp = std::fopen(out.c_str(), "wb")) {
      const int hdr[2] = {rows, cols};
      std::fwrite(hdr, sizeof(int), 2, fp);
      std::fwrite(features_.data(), sizeof(float), features_.size(), fp);
      std::fwrite(labels_.data(), sizeof(float), labels_.size(), fp);
      std::fclose(fp);
      std::cout << "Saved the training matrix as " << out << " (CvSVM::train needs OpenCV 2.4: -DAG_HAVE_OPENCV2)\n";
    }
#endif
    std::cout << "# training examples: " << rows << " (# positives: " << num_positives << ", # negatives: " << rows - num_positives
              << ")\n";
  }

  int num_horizontal_cells_, num_vertical_cells_, num_threads_;
  Localization* loc_;
  std::vector<float> features_, labels_;
};

#endif
