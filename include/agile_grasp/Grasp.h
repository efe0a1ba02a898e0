// Grasp.h — plain-struct equivalents of the ROS messages msg/Grasp.msg:1-5 and msg/Grasps.msg:1-2 with
// the field mapping of GraspLocalizer::createGraspMsg (src/agile_grasp/grasp_localizer.cpp:137-146).
#ifndef AGILE_GRASP_GRASP_MSG_H_
#define AGILE_GRASP_GRASP_MSG_H_
#include <cstdint>
#include <string>
#include <vector>

#include "grasp_hypothesis.h"

namespace geometry_msgs_lite { struct Vector3 { double x, y, z; }; }
namespace agile_grasp {
struct Float32 { float data; };
struct Grasp {
  geometry_msgs_lite::Vector3 center, axis, approach, surface_center;
  Float32 width;
};
struct Header { uint32_t seq = 0; double stamp = 0; std::string frame_id; };
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
}  // namespace agile_grasp
#endif
