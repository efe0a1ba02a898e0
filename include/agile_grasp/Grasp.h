// Grasp.h — plain-struct equivalents of the ROS messages msg/Grasp.msg:1-5 and msg/Grasps.msg:1-2 with
// the field mapping of GraspLocalizer::createGraspMsg (src/agile_grasp/grasp_localizer.cpp:137-146).
#ifndef AGILE_GRASP_GRASP_MSG_H_
#define AGILE_GRASP_GRASP_MSG_H_
#include <cstdint>
#include <string>
#include <vector>

#include <cstring>

#include "grasp_hypothesis.h"
#include "handle.h"

namespace geometry_msgs_lite { struct Vector3 { double x, y, z; }; }
namespace agile_grasp {
struct Float32 { float data; };
struct Grasp {
  geometry_msgs_lite::Vector3 center, axis, approach, surface_center;
  Float32 width;
};
struct Header { uint32_t seq = 0; uint32_t stamp_sec = 0, stamp_nsec = 0; std::string frame_id; };
struct Grasps { Header header; std::vector<Grasp> grasps; };

inline Grasp createGraspMsg(const GraspHypothesis& h) {
  auto v = [](const Eigen::Vector3d& e) { return geometry_msgs_lite::Vector3{e(0), e(1), e(2)}; };
  Grasp m;
  m.center = v(h.getGraspBottom());
  m.axis = v(h.getAxis());
  m.approach = v(h.getApproach());
  m.surface_center = v(h.getGraspSurface());
  m.width.data = float(h.getGraspWidth());
  return m;
}
inline Grasps createGraspsMsg(const std::vector<GraspHypothesis>& hands) {
  Grasps msg;
  for (const GraspHypothesis& h : hands) msg.grasps.push_back(createGraspMsg(h));
  return msg;
}
// grasp_localizer.cpp:179-188: the "average grasp" of a handle
inline Grasp createGraspMsg(const Handle& h) {
  auto v = [](const Eigen::Vector3d& e) { return geometry_msgs_lite::Vector3{e(0), e(1), e(2)}; };
  Grasp m;
  m.center = v(h.getCenter());
  m.axis = v(h.getAxis());
  m.approach = v(h.getApproach());
  m.surface_center = v(h.getHandsCenter());
  m.width.data = float(h.getWidth());
  return m;
}
// grasp_localizer.cpp:168-176
inline Grasps createGraspsMsg(const std::vector<Handle>& handles) {
  Grasps msg;
  for (const Handle& h : handles) msg.grasps.push_back(createGraspMsg(h));
  return msg;
}
// grasp_localizer.cpp:149-165: every inlier hand of every handle
inline Grasps createGraspsMsgFromHands(const std::vector<Handle>& handles) {
  Grasps msg;
  for (const Handle& h : handles)
    for (int j : h.getInliers()) msg.grasps.push_back(createGraspMsg(h.getHandList()[j]));
  return msg;
}

// ROS 1 wire format of agile_grasp/Grasps (msg/Grasps.msg:1-2, msg/Grasp.msg:1-5), little endian:
// Header = uint32 seq, uint32 stamp.sec, uint32 stamp.nsec, uint32 len + frame_id bytes; then uint32 count and,
// per grasp, 4 x geometry_msgs/Vector3 (3 x float64) + std_msgs/Float32 = 100 bytes.  What a roscpp
// subscriber of the "grasps" topic (grasp_localizer.cpp:18) receives after the 4-byte message length.
inline std::vector<uint8_t> serialize(const Grasps& msg) {
  std::vector<uint8_t> out;
  auto put = [&out](const void* p, size_t n) { const uint8_t* b = static_cast<const uint8_t*>(p); out.insert(out.end(), b, b + n); };
  const uint32_t flen = uint32_t(msg.header.frame_id.size()), cnt = uint32_t(msg.grasps.size());
  put(&msg.header.seq, 4);
  put(&msg.header.stamp_sec, 4);
  put(&msg.header.stamp_nsec, 4);
  put(&flen, 4);
  put(msg.header.frame_id.data(), flen);
  put(&cnt, 4);
  for (const Grasp& g : msg.grasps) {
    for (const geometry_msgs_lite::Vector3* v : {&g.center, &g.axis, &g.approach, &g.surface_center}) {
      put(&v->x, 8);
      put(&v->y, 8);
      put(&v->z, 8);
    }
    put(&g.width.data, 4);
  }
  return out;
}
inline bool deserialize(const std::vector<uint8_t>& in, Grasps& msg) {
  size_t pos = 0;
  auto get = [&](void* p, size_t n) { if (pos + n > in.size()) return false; std::memcpy(p, in.data() + pos, n); pos += n; return true; };
  uint32_t flen = 0, cnt = 0;
  if (!get(&msg.header.seq, 4) || !get(&msg.header.stamp_sec, 4) || !get(&msg.header.stamp_nsec, 4) || !get(&flen, 4)) return false;
  if (pos + flen > in.size()) return false;
  msg.header.frame_id.assign(reinterpret_cast<const char*>(in.data() + pos), flen);
  pos += flen;
  if (!get(&cnt, 4) || (in.size() - pos) / 100 < cnt) return false;
  msg.grasps.resize(cnt);
  for (Grasp& g : msg.grasps) {
    for (geometry_msgs_lite::Vector3* v : {&g.center, &g.axis, &g.approach, &g.surface_center})
      if (!get(&v->x, 8) || !get(&v->y, 8) || !get(&v->z, 8)) return false;
    if (!get(&g.width.data, 4)) return false;
  }
  return pos == in.size();
}
}  // namespace agile_grasp
#endif
