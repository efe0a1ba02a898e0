// examples/test_svm.cpp — the reference's offline CLI `test_svm` (src/nodes/test.cpp:4-114) on the B200
// Localization.  The reference reads .pcd files through PCL; this example reads a raw dump of
// PointXYZRGBA records (n x 32 bytes, e.g. numpy (n,8) float32 .tofile()).
//   test_svm cloud.pcd|cloud.bin svm_file [num_samples] [num_threads]
#include <agile_grasp/Grasp.h>
#include <agile_grasp/localization.h>

#include <cstdio>
#include <cstdlib>

int main(int argc, char** argv) {
  if (argc < 3) {
    std::cout << "Usage: test_svm cloud.bin svm_filepath [num_samples] [num_threads]\n";
    return -1;
  }
  const int num_samples = argc > 3 ? atoi(argv[3]) : 400;  // test.cpp:29
  const int num_threads = argc > 4 ? atoi(argv[4]) : 1;    // test.cpp:33
  // a .pcd file goes through the reference's file overload (localization.cpp:169-214); anything else is read as
  // a raw dump of PointXYZRGBA records
  const std::string cloud_path = argv[1];
  const bool is_pcd = cloud_path.size() > 4 && cloud_path.substr(cloud_path.size() - 4) == ".pcd";
  PointCloud::Ptr cloud(new PointCloud);
  if (!is_pcd) {
    FILE* f = fopen(argv[1], "rb");
    if (!f) {
      std::cout << "Couldn't read file: " << argv[1] << " \n";
      return -1;
    }
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    cloud->points.resize(size_t(bytes) / sizeof(pcl::PointXYZRGBA));
    if (fread(cloud->points.data(), sizeof(pcl::PointXYZRGBA), cloud->points.size(), f) != cloud->points.size()) return -1;
    fclose(f);
    std::cout << "Loaded point cloud with " << cloud->size() << " data points.\n";
  }

  Eigen::Matrix4d base_tf;
  base_tf << 0, 0.445417, 0.895323, 0.215, 1, 0, 0, -0.015, 0, 0.895323, -0.445417, 0.23, 0, 0, 0, 1;  // test.cpp:47-50
  Eigen::VectorXd workspace(6);
  workspace << -10, 10, -10, 10, -10, 1;  // test.cpp:68

  Localization loc(num_threads, true, 0);  // test.cpp:72
  loc.setCameraTransforms(base_tf, base_tf);
  loc.setWorkspace(workspace);
  loc.setNumSamples(num_samples);
  loc.setNeighborhoodRadiusTaubin(0.03);
  loc.setNeighborhoodRadiusHands(0.08);
  loc.setFingerWidth(0.01);
  loc.setHandOuterDiameter(0.09);
  loc.setHandDepth(0.06);
  loc.setInitBite(0.01);
  loc.setHandHeight(0.02);
  std::cout << "Localizing hands ...\n";
  std::vector<int> indices;
  std::vector<GraspHypothesis> hands = is_pcd ? loc.localizeHands(cloud_path, "", indices, false, false)  // test.cpp:95
                                              : loc.localizeHands(cloud, int(cloud->size()), indices, false, false);
  std::vector<GraspHypothesis> antipodal_hands = loc.predictAntipodalHands(hands, argv[2]);  // test.cpp:96
  std::vector<Handle> handles = loc.findHandles(antipodal_hands, 3, 0.005);                  // test.cpp:97
  agile_grasp::Grasps msg = agile_grasp::createGraspsMsg(antipodal_hands);
  agile_grasp::Grasps handle_msg = agile_grasp::createGraspsMsg(handles);
  std::cout << handles.size() << " handles; serialized Grasps message of the handles: " << agile_grasp::serialize(handle_msg).size()
            << " bytes\n";
  std::cout << hands.size() << " hands, " << msg.grasps.size() << " antipodal grasps in the Grasps message\n";
  for (size_t i = 0; i < msg.grasps.size() && i < 3; i++)
    std::cout << "  center " << msg.grasps[i].center.x << " " << msg.grasps[i].center.y << " " << msg.grasps[i].center.z
              << " width " << msg.grasps[i].width.data << " score " << antipodal_hands[i].getScore() << "\n";
  return 0;
}
