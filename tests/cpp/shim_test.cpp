// shim_test.cpp — host-side checks of the drop-in C++ interface that need no GPU: the headers compile as the
// reference's would be used, the Grasps message helpers follow grasp_localizer.cpp:123-188, the ROS 1 wire
// format round-trips, and the file overloads report unreadable files the way localization.cpp:184-189 does.
#include <agile_grasp/Grasp.h>
#include <agile_grasp/cloud_msgs.h>
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
  {  // CloudSized / PointCloud2 wire format -> PointXYZRGBA cloud (grasp_localizer.cpp:40-77), written here byte by byte
    std::vector<uint8_t> w;
    auto put = [&w](const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); w.insert(w.end(), b, b + n); };
    auto put32 = [&](uint32_t v) { put(&v, 4); };
    auto puts = [&](const std::string& s) { put32(uint32_t(s.size())); put(s.data(), s.size()); };
    put32(3); put32(100); put32(5); puts("camera_rgb_optical_frame");   // header
    const uint32_t H = 2, W = 3, step = 32;
    put32(H); put32(W);
    put32(4);                                                            // fields
    const char* names[4] = {"x", "y", "z", "rgb"};
    const uint32_t offs[4] = {0, 4, 8, 16};
    for (int k = 0; k < 4; k++) { puts(names[k]); put32(offs[k]); uint8_t dt = 7; put(&dt, 1); put32(1); }
    uint8_t z8 = 0; put(&z8, 1);                                         // is_bigendian
    put32(step); put32(step * W);
    put32(step * W * H);                                                 // data length
    for (uint32_t i = 0; i < H * W; i++) {
      float rec[8] = {float(i), float(i) + 0.5f, 1.0f + i, 0, 0, 0, 0, 0};
      uint32_t rgb = 0x00102030u + i;
      std::memcpy(&rec[4], &rgb, 4);
      put(rec, 32);
    }
    uint8_t dense = 1; put(&dense, 1);
    int64_t size_left = 4; put(&size_left, 8);
    agile_grasp::CloudSized cs;
    CHECK(agile_grasp::deserialize(w, cs));
    CHECK(cs.size_left == 4 && cs.cloud.frame_id == "camera_rgb_optical_frame" && cs.cloud.fields.size() == 4);
    PointCloud cloud;
    CHECK(agile_grasp::fromROSMsg(cs.cloud, cloud));
    CHECK(cloud.size() == 6 && cloud.points[5].x == 5.0f && cloud.points[5].y == 5.5f && cloud.points[2].z == 3.0f);
    CHECK(cloud.points[4].rgba == 0x00102034u && cloud.width == 3 && cloud.height == 2);
    {  // malformed / hostile messages are rejected instead of read out of bounds
      agile_grasp::PointCloud2 bad = cs.cloud;
      bad.fields[0].offset = 30;  // x would straddle the end of the 32-byte record
      CHECK(!agile_grasp::fromROSMsg(bad, cloud));
      bad = cs.cloud;
      bad.fields[1].offset = 0xFFFFFFF0u;
      CHECK(!agile_grasp::fromROSMsg(bad, cloud));
      bad = cs.cloud;
      bad.fields[2].datatype = 9;
      CHECK(!agile_grasp::fromROSMsg(bad, cloud));
      bad = cs.cloud;
      bad.fields[3].offset = 29;  // rgb word past the record
      CHECK(!agile_grasp::fromROSMsg(bad, cloud));
      bad = cs.cloud;
      bad.row_step = 16;  // shorter than width * point_step
      CHECK(!agile_grasp::fromROSMsg(bad, cloud));
      bad = cs.cloud;
      bad.height = 3;  // more rows than data
      CHECK(!agile_grasp::fromROSMsg(bad, cloud));
      bad = cs.cloud;
      bad.row_step = 0xFFFFFFFFu;
      CHECK(!agile_grasp::fromROSMsg(bad, cloud));
      CHECK(agile_grasp::fromROSMsg(cs.cloud, cloud));
    }
    w.resize(w.size() - 9);  // a bare PointCloud2 ends after is_dense... and a truncated one is rejected
    agile_grasp::PointCloud2 pc2;
    w.push_back(1);
    CHECK(agile_grasp::deserialize(w, pc2) && pc2.width == 3);
    w.pop_back();
    CHECK(!agile_grasp::deserialize(w, pc2));
  }
  std::printf("shim ok\n");
  return 0;
}
