// examples/train_svm.cpp — the reference's offline CLI `train_svm` (src/nodes/train.cpp:8-139) on the B200 path:
// for every training cloud localizeHands(left, right, calculates_antipodal = true, uses_clustering = true), then
// Learning::train over all hypotheses (at most 20 positives per cloud).
//   train_svm svm_file num_samples left.pcd [right.pcd] [more pairs ...]      ("" for a missing right cloud)
#include <agile_grasp/learning.h>
#include <agile_grasp/localization.h>

#include <cstdlib>

// 4x4 product and general inverse written out (the stand-in Eigen types of compat_types.h carry no algebra)
static Eigen::Matrix4d mul4(const Eigen::Matrix4d& a, const Eigen::Matrix4d& b) {
  Eigen::Matrix4d c;
  for (int r = 0; r < 4; r++)
    for (int q = 0; q < 4; q++) {
      double v = 0;
      for (int k = 0; k < 4; k++) v += a(r, k) * b(k, q);
      c(r, q) = v;
    }
  return c;
}
static Eigen::Matrix4d inv4(const Eigen::Matrix4d& m) {  // Gauss-Jordan with partial pivoting
  double a[4][8];
  for (int r = 0; r < 4; r++)
    for (int q = 0; q < 4; q++) {
      a[r][q] = m(r, q);
      a[r][4 + q] = r == q ? 1.0 : 0.0;
    }
  for (int col = 0; col < 4; col++) {
    int piv = col;
    for (int r = col + 1; r < 4; r++)
      if ((a[r][col] < 0 ? -a[r][col] : a[r][col]) > (a[piv][col] < 0 ? -a[piv][col] : a[piv][col])) piv = r;
    for (int q = 0; q < 8; q++) {
      const double t = a[col][q];
      a[col][q] = a[piv][q];
      a[piv][q] = t;
    }
    const double d = a[col][col];
    for (int q = 0; q < 8; q++) a[col][q] /= d;
    for (int r = 0; r < 4; r++)
      if (r != col) {
        const double f = a[r][col];
        for (int q = 0; q < 8; q++) a[r][q] -= f * a[col][q];
      }
  }
  Eigen::Matrix4d out;
  for (int r = 0; r < 4; r++)
    for (int q = 0; q < 4; q++) out(r, q) = a[r][4 + q];
  return out;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    std::cout << "Usage: train_svm svm_file num_samples left.pcd [right.pcd] ...\n";
    return -1;
  }
  const std::string svm_file_name = argv[1];
  const int num_samples = atoi(argv[2]);
  Eigen::Matrix4d base_tf, sqrt_tf;
  base_tf << 0, 0.445417, 0.895323, 0.21, 1, 0, 0, -0.02, 0, 0.895323, -0.445417, 0.24, 0, 0, 0, 1;  // train.cpp:83-86
  sqrt_tf << 0.9366, -0.0162, 0.3500, -0.2863, 0.0151, 0.9999, 0.0058, 0.0058, -0.3501, -0.0002, 0.9367, 0.0554, 0, 0, 0, 1;
  Localization loc(4, false, 0);  // train.cpp:94
  loc.setCameraTransforms(mul4(base_tf, inv4(sqrt_tf)), mul4(base_tf, sqrt_tf));  // train.cpp:95
  loc.setNumSamples(num_samples);
  loc.setNeighborhoodRadiusTaubin(0.03);
  loc.setNeighborhoodRadiusHands(0.08);
  loc.setFingerWidth(0.01);
  loc.setHandOuterDiameter(0.09);
  loc.setHandDepth(0.06);
  loc.setInitBite(0.015);  // train.cpp:102
  loc.setHandHeight(0.02);
  loc.setKeepTrainingFeatures(true);
  Eigen::VectorXd workspace(6);
  workspace << -10, 10, -10, 10, -10, 10;
  loc.setWorkspace(workspace);
  std::cout << "Acquiring training data ...\n";
  std::vector<GraspHypothesis> hand_list;
  std::vector<int> hand_list_sizes;
  for (int a = 3; a < argc; a += 2) {
    const std::string left = argv[a], right = a + 1 < argc ? argv[a + 1] : "";
    std::cout << " Creating training data from file " << left << " ...\n";
    std::vector<GraspHypothesis> hands = loc.localizeHands(left, right, true, true);  // train.cpp:115
    hand_list.insert(hand_list.end(), hands.begin(), hands.end());
    hand_list_sizes.push_back(int(hand_list.size()));
    std::cout << hand_list_sizes.size() - 1 << ") # hands: " << hands.size() << std::endl;
  }
  std::cout << "Training the SVM ...\n";
  Learning learn;
  Eigen::Matrix3Xd cam_pos(3, 2);
  for (int r = 0; r < 3; r++) {
    cam_pos(r, 0) = loc.getCameraTransform(true)(r, 3);
    cam_pos(r, 1) = loc.getCameraTransform(false)(r, 3);
  }
  const int max_positives = 20;  // train.cpp:125
  learn.train(hand_list, hand_list_sizes, svm_file_name, cam_pos, max_positives);
  return 0;
}
