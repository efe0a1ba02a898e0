// shim_test.cpp — host-side checks of the drop-in C++ interface that need no GPU: the headers compile as the
// reference's would be used, the Grasps message helpers follow grasp_localizer.cpp:123-188, the ROS 1 wire
// format round-trips, and the file overloads report unreadable files the way localization.cpp:184-189 does.
#include <agile_grasp/Grasp.h>
#include <agile_grasp/localization.h>

#include <cstdio>
#include <cstdlib>

#define CHECK(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
  // hypotheses through the reference's own constructor (grasp_hypothesis.h:67-75)
  std::vector<GraspHypothesis> hands;
  for (int i = 0; i < 4; i++) {
    Eigen::Vector3d axis, approach, binormal, bottom, surface;
    axis << 1, 0, 0;
    approach << 0, 0, -1;
    binormal << 0, 1, 0;
    bottom << 0.01 * i, 0.2, 0.5;
    surface << 0.01 * i, 0.2, 0.52;
    Eigen::Matrix3Xd pts(3, 0);
    hands.push_back(GraspHypothesis(axis, approach, binormal, bottom, surface, 0.03 + 0.001 * i, pts, {}, {}, 0));
  }
  agile_grasp::Grasps msg = agile_grasp::createGraspsMsg(hands);
  CHECK(msg.grasps.size() == 4);
  CHECK(msg.grasps[2].center.x == 0.02 && msg.grasps[2].surface_center.z == 0.52 && msg.grasps[2].axis.x == 1.0);
  CHECK(msg.grasps[3].width.data == float(0.033));
  msg.header.seq = 7;
  msg.header.stamp_sec = 1426809600u;
  msg.header.stamp_nsec = 250;
  msg.header.frame_id = "camera_rgb_optical_frame";
  std::vector<uint8_t> wire = agile_grasp::serialize(msg);
  CHECK(wire.size() == 16 + msg.header.frame_id.size() + 4 + 4 * 100);
  agile_grasp::Grasps back;
  CHECK(agile_grasp::deserialize(wire, back));
  CHECK(back.header.seq == 7 && back.header.stamp_sec == 1426809600u && back.header.frame_id == msg.header.frame_id);
  CHECK(back.grasps.size() == 4 && back.grasps[1].center.x == msg.grasps[1].center.x &&
        back.grasps[3].width.data == msg.grasps[3].width.data && back.grasps[0].approach.z == -1.0);
  wire.pop_back();
  CHECK(!agile_grasp::deserialize(wire, back));
  // a Handle carries what ag_find_handles computed; createGraspMsg(handle) maps it as grasp_localizer.cpp:179-188
  ag_handle h;
  std::memset(&h, 0, sizeof(h));
  h.axis[0] = 1;
  h.center[1] = 0.2;
  h.hands_center[2] = 0.52;
  h.approach[2] = -1;
  h.width = 0.031;
  Handle handle(hands, {0, 1, 2}, h);
  agile_grasp::Grasp hm = agile_grasp::createGraspMsg(handle);
  CHECK(hm.center.y == 0.2 && hm.surface_center.z == 0.52 && hm.axis.x == 1.0 && hm.width.data == float(0.031));
  CHECK(agile_grasp::createGraspsMsgFromHands({handle}).grasps.size() == 3);
  CHECK(agile_grasp::createGraspsMsg(std::vector<Handle>{handle, handle}).grasps.size() == 2);
  // file overload: unreadable file -> message + empty list (localization.cpp:184-189), no device needed
  Localization loc(1, true, Localization::NO_PLOTTING);
  CHECK(loc.localizeHands("/nonexistent/left.pcd", "", false, false).empty());
  if (argc > 1) {  // a PCD file written by the Python test: the loader is reached through the C ABI
    void* pts = nullptr;
    int n = 0, w = 0, hh = 0;
    CHECK(ag_load_pcd(argv[1], &pts, &n, &w, &hh) == AG_OK);
    CHECK(n == std::atoi(argv[2]) && w * hh == n);
    ag_free(pts);
  }
  std::printf("shim ok\n");
  return 0;
}
