// cloud_msgs.h — the input side of the node boundary without ROS: sensor_msgs/PointCloud2 and
// agile_grasp/CloudSized (msg/CloudSized.msg:1-2) in their ROS 1 wire format, converted to the
// pcl::PointCloud<pcl::PointXYZRGBA> that Localization::localizeHands takes — what pcl::fromROSMsg does in
// GraspLocalizer::cloud_callback / cloud_sized_callback (src/agile_grasp/grasp_localizer.cpp:40-77).
//
// Wire layout (little endian): Header {uint32 seq, uint32 sec, uint32 nsec, string frame_id}, uint32 height,
// uint32 width, PointField[] {string name, uint32 offset, uint8 datatype, uint32 count}, uint8 is_bigendian,
// uint32 point_step, uint32 row_step, uint8[] data, uint8 is_dense.  strings / arrays carry a uint32 length.
// CloudSized = PointCloud2 cloud, std_msgs/Int64 size_left.
#ifndef AGILE_GRASP_CLOUD_MSGS_H_
#define AGILE_GRASP_CLOUD_MSGS_H_

#include <cstdint>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "compat_types.h"

namespace agile_grasp {

struct PointField {
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 7;  // 1 INT8 2 UINT8 3 INT16 4 UINT16 5 INT32 6 UINT32 7 FLOAT32 8 FLOAT64
  uint32_t count = 1;
};
struct PointCloud2 {
  uint32_t seq = 0, stamp_sec = 0, stamp_nsec = 0;
  std::string frame_id;
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  uint8_t is_bigendian = 0;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  uint8_t is_dense = 0;
};
struct CloudSized {
  PointCloud2 cloud;
  int64_t size_left = 0;
};

namespace detail {
struct Reader {
  const uint8_t* p;
  size_t n, pos;
  bool get(void* out, size_t k) {
    if (pos + k > n) return false;
    std::memcpy(out, p + pos, k);
    pos += k;
    return true;
  }
  bool str(std::string& s) {
    uint32_t len;
    if (!get(&len, 4) || pos + len > n) return false;
    s.assign(reinterpret_cast<const char*>(p + pos), len);
    pos += len;
    return true;
  }
};
inline bool read_cloud(Reader& r, PointCloud2& m) {
  uint32_t nf = 0, nd = 0;
  if (!r.get(&m.seq, 4) || !r.get(&m.stamp_sec, 4) || !r.get(&m.stamp_nsec, 4) || !r.str(m.frame_id)) return false;
  if (!r.get(&m.height, 4) || !r.get(&m.width, 4) || !r.get(&nf, 4) || nf > 1024 || size_t(nf) * 13 > r.n - r.pos) return false;
  m.fields.resize(nf);
  for (PointField& f : m.fields)
    if (!r.str(f.name) || !r.get(&f.offset, 4) || !r.get(&f.datatype, 1) || !r.get(&f.count, 4)) return false;
  if (!r.get(&m.is_bigendian, 1) || !r.get(&m.point_step, 4) || !r.get(&m.row_step, 4) || !r.get(&nd, 4)) return false;
  if (nd > r.n - r.pos) return false;
  m.data.assign(r.p + r.pos, r.p + r.pos + nd);
  r.pos += nd;
  return r.get(&m.is_dense, 1);
}
inline double scalar(const uint8_t* p, uint8_t datatype) {
  switch (datatype) {
    case 1: return *reinterpret_cast<const int8_t*>(p);
    case 2: return *p;
    case 3: { int16_t v; std::memcpy(&v, p, 2); return v; }
    case 4: { uint16_t v; std::memcpy(&v, p, 2); return v; }
    case 5: { int32_t v; std::memcpy(&v, p, 4); return v; }
    case 6: { uint32_t v; std::memcpy(&v, p, 4); return v; }
    case 7: { float v; std::memcpy(&v, p, 4); return v; }
    case 8: { double v; std::memcpy(&v, p, 8); return v; }
  }
  return std::numeric_limits<double>::quiet_NaN();
}
}  // namespace detail

inline bool deserialize(const std::vector<uint8_t>& wire, PointCloud2& msg) {
  detail::Reader r{wire.data(), wire.size(), 0};
  return detail::read_cloud(r, msg) && r.pos == wire.size();
}
inline bool deserialize(const std::vector<uint8_t>& wire, CloudSized& msg) {
  detail::Reader r{wire.data(), wire.size(), 0};
  return detail::read_cloud(r, msg.cloud) && r.get(&msg.size_left, 8) && r.pos == wire.size();
}

/** pcl::fromROSMsg for PointXYZRGBA (grasp_localizer.cpp:55-58,73): x, y, z by field name (any numeric
 *  datatype), rgb / rgba as the packed 32-bit word; missing colour stays 0.  Returns false if x/y/z are
 *  missing, the message is big endian, a used field does not fit inside point_step (offset / datatype), row_step is
 *  shorter than width * point_step, or data is shorter than height * row_step. */
inline bool fromROSMsg(const PointCloud2& msg, pcl::PointCloud<pcl::PointXYZRGBA>& cloud) {
  const PointField *fx = nullptr, *fy = nullptr, *fz = nullptr, *fc = nullptr;
  for (const PointField& f : msg.fields) {
    if (f.name == "x") fx = &f;
    else if (f.name == "y") fy = &f;
    else if (f.name == "z") fz = &f;
    else if (f.name == "rgba" || f.name == "rgb") fc = &f;
  }
  const size_t n = size_t(msg.width) * msg.height;
  if (!fx || !fy || !fz || msg.is_bigendian || msg.point_step == 0) return false;
  // the message is external input: every field that is read must lie inside a point record, every record inside
  // its row, every row inside data (a hostile offset / row_step must not turn into an out-of-bounds read)
  static const uint32_t kTypeSize[9] = {0, 1, 1, 2, 2, 4, 4, 4, 8};
  for (const PointField* f : {fx, fy, fz}) {
    if (f->datatype < 1 || f->datatype > 8) return false;
    if (uint64_t(f->offset) + kTypeSize[f->datatype] > msg.point_step) return false;
  }
  if (fc && uint64_t(fc->offset) + 4 > msg.point_step) return false;
  const size_t row_step = msg.row_step ? msg.row_step : size_t(msg.point_step) * msg.width;
  if (row_step < size_t(msg.point_step) * msg.width) return false;
  if (msg.height != 0 && row_step > msg.data.size() / msg.height) return false;  // height * row_step <= size, no overflow
  cloud.points.resize(n);
  cloud.width = msg.width;
  cloud.height = msg.height;
  for (uint32_t r = 0; r < msg.height; r++)
    for (uint32_t c = 0; c < msg.width; c++) {
      const uint8_t* p = msg.data.data() + size_t(r) * row_step + size_t(c) * msg.point_step;
      pcl::PointXYZRGBA& o = cloud.points[size_t(r) * msg.width + c];
      o.x = float(detail::scalar(p + fx->offset, fx->datatype));
      o.y = float(detail::scalar(p + fy->offset, fy->datatype));
      o.z = float(detail::scalar(p + fz->offset, fz->datatype));
      o.data3 = 1.0f;
      o.rgba = 0;
      if (fc) std::memcpy(&o.rgba, p + fc->offset, 4);
    }
  return true;
}

}  // namespace agile_grasp
#endif
