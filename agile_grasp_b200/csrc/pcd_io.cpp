// pcd_io.cpp — PCD (Point Cloud Data v0.7) reader for the file overloads of Localization::localizeHands
// (src/agile_grasp/localization.cpp:169-214, which call pcl::io::loadPCDFile<pcl::PointXYZRGBA> :184,:198).
// PCL is not vendored in the reference; the format is restated from its published specification: a text
// header (VERSION, FIELDS, SIZE, TYPE, COUNT, WIDTH, HEIGHT, VIEWPOINT, POINTS, DATA) followed by
// `ascii` rows, `binary` records in header field order, or `binary_compressed`: uint32 compressed size,
// uint32 uncompressed size, LZF stream holding the fields one after the other (structure of arrays).
// Output = pcl::PointXYZRGBA records (32 bytes: x y z f32 at 0/4/8, rgba u32 at 16), the layout ag_localize
// takes.  Missing x/y/z -> error; a missing rgb/rgba field leaves colour 0 (the hot path ignores colour).
// Host-only code (no CUDA): loading a file does not need a device.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/ag_b200.h"

namespace ag {
void set_error(const std::string& msg);
}

namespace {

struct Field {
  std::string name;
  int size = 4;
  char type = 'F';
  int count = 1;
  size_t offset = 0;  // byte offset inside one binary record
};

// LZF decompression (Marc Lehmann's liblzf format, the one PCL uses): control byte c;
// c < 32: literal run of c + 1 bytes; else back reference: len = c >> 5 (7 -> + next byte), + 2,
// distance = ((c & 0x1f) << 8 | next byte) + 1.
bool lzf_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  size_t ip = 0, op = 0;
  while (ip < in_len) {
    unsigned ctrl = in[ip++];
    if (ctrl < 32) {
      const size_t run = ctrl + 1;
      if (op + run > out_len || ip + run > in_len) return false;
      std::memcpy(out + op, in + ip, run);
      op += run;
      ip += run;
    } else {
      size_t len = ctrl >> 5;
      if (len == 7) {
        if (ip >= in_len) return false;
        len += in[ip++];
      }
      if (ip >= in_len) return false;
      const size_t dist = ((size_t(ctrl) & 0x1f) << 8 | in[ip++]) + 1;
      len += 2;
      if (dist > op || op + len > out_len) return false;
      for (size_t k = 0; k < len; k++, op++) out[op] = out[op - dist];  // may overlap: byte by byte
    }
  }
  return op == out_len;
}

double read_scalar(const uint8_t* p, const Field& f) {
  switch (f.type) {
    case 'F':
      if (f.size == 4) { float v; std::memcpy(&v, p, 4); return v; }
      if (f.size == 8) { double v; std::memcpy(&v, p, 8); return v; }
      break;
    case 'U':
      if (f.size == 1) return *p;
      if (f.size == 2) { uint16_t v; std::memcpy(&v, p, 2); return v; }
      if (f.size == 4) { uint32_t v; std::memcpy(&v, p, 4); return v; }
      break;
    case 'I':
      if (f.size == 1) return *reinterpret_cast<const int8_t*>(p);
      if (f.size == 2) { int16_t v; std::memcpy(&v, p, 2); return v; }
      if (f.size == 4) { int32_t v; std::memcpy(&v, p, 4); return v; }
      break;
  }
  return std::numeric_limits<double>::quiet_NaN();
}

struct Rec {
  float x, y, z, data3;
  uint32_t rgba;
  uint32_t pad[3];
};
static_assert(sizeof(Rec) == 32, "pcl::PointXYZRGBA layout");

}  // namespace

extern "C" int ag_load_pcd(const char* path, void** points_out, int* n_out, int* width_out, int* height_out) {
  if (!path || !points_out || !n_out) return AG_ERR_INVALID;
  *points_out = nullptr;
  *n_out = 0;
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    ag::set_error(std::string("Couldn't read file: ") + path);
    return AG_ERR_INVALID;
  }
  std::fseek(f, 0, SEEK_END);
  const long fsz = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf(fsz > 0 ? size_t(fsz) : 0);
  const size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
  std::fclose(f);
  if (got != buf.size()) {
    ag::set_error(std::string("short read: ") + path);
    return AG_ERR_INVALID;
  }
  // ---- header
  std::vector<Field> fields;
  long width = -1, height = 1, points = -1;
  std::string data_kind;
  size_t pos = 0;
  bool have_size = false, have_type = false;
  while (pos < buf.size()) {
    size_t eol = pos;
    while (eol < buf.size() && buf[eol] != '\n') eol++;
    std::string line(reinterpret_cast<const char*>(buf.data() + pos), eol - pos);
    pos = eol < buf.size() ? eol + 1 : eol;
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ss(line);
    std::string key;
    ss >> key;
    if (key == "VERSION" || key == "VIEWPOINT") continue;
    if (key == "FIELDS" || key == "COLUMNS") {
      std::string nm;
      while (ss >> nm) {
        Field fd;
        fd.name = nm;
        fields.push_back(fd);
      }
    } else if (key == "SIZE") {
      for (Field& fd : fields) ss >> fd.size;
      have_size = true;
    } else if (key == "TYPE") {
      for (Field& fd : fields) ss >> fd.type;
      have_type = true;
    } else if (key == "COUNT") {
      for (Field& fd : fields) ss >> fd.count;
    } else if (key == "WIDTH") {
      ss >> width;
    } else if (key == "HEIGHT") {
      ss >> height;
    } else if (key == "POINTS") {
      ss >> points;
    } else if (key == "DATA") {
      ss >> data_kind;
      break;
    } else {
      ag::set_error("PCD header: unknown entry '" + key + "' in " + path);
      return AG_ERR_INVALID;
    }
  }
  if (fields.empty() || !have_size || !have_type || data_kind.empty()) {
    ag::set_error(std::string("PCD header incomplete (FIELDS/SIZE/TYPE/DATA): ") + path);
    return AG_ERR_INVALID;
  }
  if (points < 0) points = width > 0 ? width * height : 0;
  if (width < 0) width = points;
  int ix = -1, iy = -1, iz = -1, ic = -1;
  size_t rec_bytes = 0;
  for (size_t k = 0; k < fields.size(); k++) {
    Field& fd = fields[k];
    if (fd.size <= 0 || fd.count <= 0 || fd.size > 8) {
      ag::set_error(std::string("PCD header: bad SIZE/COUNT in ") + path);
      return AG_ERR_INVALID;
    }
    fd.offset = rec_bytes;
    rec_bytes += size_t(fd.size) * fd.count;
    if (fd.name == "x") ix = int(k);
    else if (fd.name == "y") iy = int(k);
    else if (fd.name == "z") iz = int(k);
    else if (fd.name == "rgba" || fd.name == "rgb") ic = int(k);
  }
  if (ix < 0 || iy < 0 || iz < 0) {
    ag::set_error(std::string("PCD file has no x/y/z fields: ") + path);
    return AG_ERR_INVALID;
  }
  const size_t n = size_t(points);
  Rec* out = static_cast<Rec*>(std::calloc(n ? n : 1, sizeof(Rec)));
  if (!out) return AG_ERR_CUDA;
  auto fail = [&](const std::string& m) {
    std::free(out);
    ag::set_error(m + ": " + path);
    return AG_ERR_INVALID;
  };
  for (size_t i = 0; i < n; i++) out[i].data3 = 1.0f;  // PCL initialises the padding float of XYZ to 1
  auto colour_bits = [&](const uint8_t* p, const Field& fd) -> uint32_t {
    // rgb / rgba are stored as a 4-byte word (TYPE U or, historically, the same bits in TYPE F)
    uint32_t v = 0;
    if (fd.size == 4) std::memcpy(&v, p, 4);
    return v;
  };
  if (data_kind == "ascii") {
    const char* p = reinterpret_cast<const char*>(buf.data()) + pos;
    const char* end = reinterpret_cast<const char*>(buf.data()) + buf.size();
    std::string text(p, end - p);
    std::istringstream ss(text);
    for (size_t i = 0; i < n; i++) {
      for (size_t k = 0; k < fields.size(); k++) {
        for (int c = 0; c < fields[k].count; c++) {
          std::string tok;
          if (!(ss >> tok)) return fail("PCD ascii data ends early");
          if (c > 0) continue;
          if (int(k) == ic) {
            if (fields[k].type == 'F') {  // the packed colour printed as a float
              float fv = std::strtof(tok.c_str(), nullptr);
              std::memcpy(&out[i].rgba, &fv, 4);
            } else {
              out[i].rgba = uint32_t(std::strtoul(tok.c_str(), nullptr, 10));
            }
          } else if (int(k) == ix || int(k) == iy || int(k) == iz) {
            const float v = tok == "nan" || tok == "NaN" || tok == "-nan" ? std::numeric_limits<float>::quiet_NaN()
                                                                          : std::strtof(tok.c_str(), nullptr);
            (int(k) == ix ? out[i].x : int(k) == iy ? out[i].y : out[i].z) = v;
          }
        }
      }
    }
  } else if (data_kind == "binary") {
    if (buf.size() - pos < n * rec_bytes) return fail("PCD binary data ends early");
    const uint8_t* base = buf.data() + pos;
    for (size_t i = 0; i < n; i++) {
      const uint8_t* r = base + i * rec_bytes;
      out[i].x = float(read_scalar(r + fields[ix].offset, fields[ix]));
      out[i].y = float(read_scalar(r + fields[iy].offset, fields[iy]));
      out[i].z = float(read_scalar(r + fields[iz].offset, fields[iz]));
      if (ic >= 0) out[i].rgba = colour_bits(r + fields[ic].offset, fields[ic]);
    }
  } else if (data_kind == "binary_compressed") {
    if (buf.size() - pos < 8) return fail("PCD compressed data ends early");
    uint32_t csz, usz;
    std::memcpy(&csz, buf.data() + pos, 4);
    std::memcpy(&usz, buf.data() + pos + 4, 4);
    if (buf.size() - pos - 8 < csz || size_t(usz) != n * rec_bytes) return fail("PCD compressed sizes inconsistent");
    std::vector<uint8_t> raw(usz);
    if (!lzf_decompress(buf.data() + pos + 8, csz, raw.data(), usz)) return fail("PCD LZF stream corrupt");
    // structure of arrays: all values of field 0, then field 1, ...
    std::vector<size_t> start(fields.size());
    size_t acc = 0;
    for (size_t k = 0; k < fields.size(); k++) {
      start[k] = acc;
      acc += size_t(fields[k].size) * fields[k].count * n;
    }
    for (size_t i = 0; i < n; i++) {
      auto at = [&](int k) { return raw.data() + start[k] + i * size_t(fields[k].size) * fields[k].count; };
      out[i].x = float(read_scalar(at(ix), fields[ix]));
      out[i].y = float(read_scalar(at(iy), fields[iy]));
      out[i].z = float(read_scalar(at(iz), fields[iz]));
      if (ic >= 0) out[i].rgba = colour_bits(at(ic), fields[ic]);
    }
  } else {
    return fail("PCD DATA kind '" + data_kind + "' unknown");
  }
  *points_out = out;
  *n_out = int(n);
  if (width_out) *width_out = int(width);
  if (height_out) *height_out = int(height);
  return AG_OK;
}
