// oracle_learn.cpp — CPU oracle, scoring half of the hot path: grasp image, HOG, SVM, and the
// full-path driver used as the timed CPU baseline.
//
// TEST INFRASTRUCTURE ONLY (see ag_oracle.h).  HOG and the SVM decision function belong to
// OpenCV (not vendored in the reference; reference call sites learning.cpp:194-195,220,225).
// They are restated here from OpenCV's published algorithm (objdetect/hog.cpp
// HOGDescriptor::computeGradient / HOGCache::init / getBlock / normalizeBlockHistogram and
// ml/svm.cpp calc_non_rbf_base / predict) and pinned against the cv2 4.13 present in this image
// by tests/test_oracle_hog_svm.py.
//
// Build with -ffp-contract=off.

#include "oracle_internal.h"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <set>
#include <sstream>
#include <string>
#include <vector>

namespace ago {

// ---------------------------------------------------------------------------------------------
// A.10 grasp image — learning.cpp:375-400 (createInstance), :320-365 (convertToImage),
// :367-373 (floorVector)
// ---------------------------------------------------------------------------------------------
void points_image(const double* pts, int m, const double binormal[3], const double surface[3],
                  const double cam_pos[3], uint8_t* img /*80x100*/) {
  const int W = 100, Hh = 80;
  const double HL[2] = {-0.05, 0.05};
  const double VL0 = 0.0;
  const double cell = (HL[1] - HL[0]) / double(W);  // :324
  double s2c[3] = {surface[0] - cam_pos[0], surface[1] - cam_pos[1], surface[2] - cam_pos[2]};  // :382-383
  // Vector3d dot: Eigen unrolled order a0 + (a1 + a2)
  double dt = binormal[0] * s2c[0] + (binormal[1] * s2c[1] + binormal[2] * s2c[2]);
  const bool keep_sign = dt > 0;  // :330
  std::memset(img, 0, size_t(W) * Hh);
  // the std::set of cells (:339-345) only de-duplicates; writing 255 is idempotent
  for (int i = 0; i < m; i++) {
    double x = pts[3 * i], y = pts[3 * i + 1];
    int h = int(std::floor(((keep_sign ? x : -x) - HL[0]) / cell));
    int v = int(std::floor((y - VL0) / cell));
    h = std::min(W - 1, std::max(0, h));   // :354-359
    v = std::min(Hh - 1, std::max(0, v));
    img[(Hh - 1 - v) * W + h] = 255;       // :360
  }
}

// ---------------------------------------------------------------------------------------------
// C.3 HOG — cv::HOGDescriptor default ctor + winSize 64x64, compute(img, winStride 32, padding 0)
// ---------------------------------------------------------------------------------------------
namespace hog {
constexpr int W = 100, H = 80, NB = 9, BS = 16, CS = 8, STRIDE = 8, WIN = 64;
constexpr int NBLK = (WIN - BS) / STRIDE + 1;  // 7
constexpr int BH = 4 * NB;                      // 36 floats per block

// cv::cartToPolar on (dx,dy) in {-s,0,+s}^2, s = sqrt(255.f): values produced by cv2 4.13
// (tools/gen_hog_tables.py; OpenCV uses a polynomial atan, so diagonals are not k*pi/4)
static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
// index = (sy+1)*3 + (sx+1), sx = sign(dx), sy = sign(dy)
static const uint32_t kMagBits[9] = {0x41b4aa5a, 0x417f7fe0, 0x41b4aa5a, 0x417f7fe0, 0x0,
                                     0x417f7fe0, 0x41b4aa5a, 0x417f7fe0, 0x41b4aa5a};
static const uint32_t kAngBits[9] = {0x407b5116, 0x4096cbe4, 0x40afef3d, 0x40490fdb, 0x0,
                                     0x0,        0x4016ce9f, 0x3fc90fdb, 0x3f4904f0};

struct PixData {
  int gradOfs, histOfs[4];
  float histWeights[4], gradWeight;
};
struct Tables {
  float g0[9], g1[9];
  int h0[9], h1[9];
  std::vector<PixData> pix;  // count1 | count2 | count4 in OpenCV's order
  int c1, c2, c4;
  Tables() {
    // computeGradient: angle*angleScale - 0.5f, split between bins
    const float angleScale = float(NB / M_PI);
    for (int k = 0; k < 9; k++) {
      float mag = u2f(kMagBits[k]);
      float angle = u2f(kAngBits[k]) * angleScale - 0.5f;
      int hidx = int(std::floor(angle));
      angle -= float(hidx);
      g0[k] = mag * (1.f - angle);
      g1[k] = mag * angle;
      if (hidx < 0) hidx += NB;
      else if (hidx >= NB) hidx -= NB;
      h0[k] = hidx;
      hidx++;
      if (hidx >= NB) hidx = 0;
      h1[k] = hidx;
    }
    // HOGCache::init: gaussian block weights and bilinear cell weights
    float weights[BS][BS];
    const float sigma = 4.0f;  // winSigma=-1 -> (blockSize.w+blockSize.h)/8
    const float scale = 1.f / (sigma * sigma * 2);
    float di[BS], dj[BS];
    for (int i = 0; i < BS; i++) {
      di[i] = float(i) - BS * 0.5f;
      di[i] *= di[i];
      dj[i] = di[i];
    }
    for (int i = 0; i < BS; i++)
      for (int j = 0; j < BS; j++) weights[i][j] = std::exp(-(di[i] + dj[j]) * scale);
    std::vector<PixData> p1, p2, p4;
    const int nc = BS / CS;  // 2 cells per side
    for (int j = 0; j < BS; j++)
      for (int i = 0; i < BS; i++) {
        PixData d;
        std::memset(&d, 0, sizeof(d));
        float cellX = (j + 0.5f) / CS - 0.5f;
        float cellY = (i + 0.5f) / CS - 0.5f;
        int ix0 = int(std::floor(cellX)), iy0 = int(std::floor(cellY));
        int ix1 = ix0 + 1, iy1 = iy0 + 1;
        cellX -= ix0;
        cellY -= iy0;
        std::vector<PixData>* dst;
        auto okx = [&](int v) { return unsigned(v) < unsigned(nc); };
        if (okx(ix0) && okx(ix1)) {
          if (okx(iy0) && okx(iy1)) {
            dst = &p4;
            d.histOfs[0] = (ix0 * nc + iy0) * NB;
            d.histWeights[0] = (1.f - cellX) * (1.f - cellY);
            d.histOfs[1] = (ix1 * nc + iy0) * NB;
            d.histWeights[1] = cellX * (1.f - cellY);
            d.histOfs[2] = (ix0 * nc + iy1) * NB;
            d.histWeights[2] = (1.f - cellX) * cellY;
            d.histOfs[3] = (ix1 * nc + iy1) * NB;
            d.histWeights[3] = cellX * cellY;
          } else {
            dst = &p2;
            if (okx(iy0)) {
              iy1 = iy0;
              cellY = 1.f - cellY;
            }
            d.histOfs[0] = (ix0 * nc + iy1) * NB;
            d.histWeights[0] = (1.f - cellX) * cellY;
            d.histOfs[1] = (ix1 * nc + iy1) * NB;
            d.histWeights[1] = cellX * cellY;
          }
        } else {
          if (okx(ix0)) {
            ix1 = ix0;
            cellX = 1.f - cellX;
          }
          if (okx(iy0) && okx(iy1)) {
            dst = &p2;
            d.histOfs[0] = (ix1 * nc + iy0) * NB;
            d.histWeights[0] = cellX * (1.f - cellY);
            d.histOfs[1] = (ix1 * nc + iy1) * NB;
            d.histWeights[1] = cellX * cellY;
          } else {
            dst = &p1;
            if (okx(iy0)) {
              iy1 = iy0;
              cellY = 1.f - cellY;
            }
            d.histOfs[0] = (ix1 * nc + iy1) * NB;
            d.histWeights[0] = cellX * cellY;
          }
        }
        d.gradOfs = i * W + j;  // pixel offset inside the gradient image, relative to block origin
        d.gradWeight = weights[i][j];
        dst->push_back(d);
      }
    c1 = int(p1.size());
    c2 = c1 + int(p2.size());
    c4 = c2 + int(p4.size());
    pix = p1;
    pix.insert(pix.end(), p2.begin(), p2.end());
    pix.insert(pix.end(), p4.begin(), p4.end());
  }
};
static const Tables& tables() {
  static const Tables t;
  return t;
}

static inline int reflect101(int p, int len) {  // cv::borderInterpolate(BORDER_REFLECT_101)
  if (p < 0) return -p;
  if (p >= len) return 2 * len - 2 - p;
  return p;
}

static void normalize_block(float* hist) {  // HOGCache::normalizeBlockHistogram (L2-Hys, 0.2)
  const int sz = BH;
  float part[4] = {0, 0, 0, 0};
  for (int i = 0; i < sz; i += 4)
    for (int l = 0; l < 4; l++) part[l] += hist[i + l] * hist[i + l];
  float t0 = part[0] + part[1], t1 = part[2] + part[3];
  float sum = t0 + t1;
  float scale = 1.f / (std::sqrt(sum) + sz * 0.1f);
  const float thresh = 0.2f;
  part[0] = part[1] = part[2] = part[3] = 0;
  for (int i = 0; i < sz; i += 4)
    for (int l = 0; l < 4; l++) {
      float p = std::min(hist[i + l] * scale, thresh);
      hist[i + l] = p;
      part[l] += p * p;
    }
  t0 = part[0] + part[1];
  t1 = part[2] + part[3];
  sum = t0 + t1;
  scale = 1.f / (std::sqrt(sum) + 1e-3f);
  for (int i = 0; i < sz; i++) hist[i] *= scale;
}

void compute(const uint8_t* img, float* desc) {
  const Tables& T = tables();
  // gradient "case" per pixel: gamma LUT sqrt(v) -> values 0 or sqrt(255); [-1,0,1] derivative
  // with BORDER_REFLECT_101 on the whole 100x80 image
  std::vector<uint8_t> gcase(size_t(W) * H);
  for (int y = 0; y < H; y++) {
    int yp = reflect101(y - 1, H), yn = reflect101(y + 1, H);
    for (int x = 0; x < W; x++) {
      int xp = reflect101(x - 1, W), xn = reflect101(x + 1, W);
      int sx = (img[y * W + xn] != 0) - (img[y * W + xp] != 0);
      int sy = (img[yn * W + x] != 0) - (img[yp * W + x] != 0);
      gcase[y * W + x] = uint8_t((sy + 1) * 3 + (sx + 1));
    }
  }
  const int nwin = (W - WIN) / 32 + 1;  // 2 windows across, 1 down
  for (int w = 0; w < nwin; w++)
    for (int bx = 0; bx < NBLK; bx++)
      for (int by = 0; by < NBLK; by++) {
        float* hist = desc + size_t(w) * NBLK * NBLK * BH + (bx * NBLK + by) * BH;
        std::fill(hist, hist + BH, 0.f);
        const int ox = w * 32 + bx * STRIDE, oy = by * STRIDE;
        const uint8_t* gc = &gcase[oy * W + ox];
        int k = 0;
        for (; k < T.c1; k++) {
          const PixData& pk = T.pix[k];
          int c = gc[pk.gradOfs];
          float wgt = pk.gradWeight * pk.histWeights[0];
          float* h = hist + pk.histOfs[0];
          float t0 = h[T.h0[c]] + T.g0[c] * wgt;
          float t1 = h[T.h1[c]] + T.g1[c] * wgt;
          h[T.h0[c]] = t0;
          h[T.h1[c]] = t1;
        }
        for (; k < T.c2; k++) {
          const PixData& pk = T.pix[k];
          int c = gc[pk.gradOfs];
          for (int q = 0; q < 2; q++) {
            float wgt = pk.gradWeight * pk.histWeights[q];
            float* h = hist + pk.histOfs[q];
            float t0 = h[T.h0[c]] + T.g0[c] * wgt;
            float t1 = h[T.h1[c]] + T.g1[c] * wgt;
            h[T.h0[c]] = t0;
            h[T.h1[c]] = t1;
          }
        }
        for (; k < T.c4; k++) {
          const PixData& pk = T.pix[k];
          int c = gc[pk.gradOfs];
          for (int q = 0; q < 4; q++) {
            float wgt = pk.gradWeight * pk.histWeights[q];
            float* h = hist + pk.histOfs[q];
            float t0 = h[T.h0[c]] + T.g0[c] * wgt;
            float t1 = h[T.h1[c]] + T.g1[c] * wgt;
            h[T.h0[c]] = t0;
            h[T.h1[c]] = t1;
          }
        }
        normalize_block(hist);
      }
}
}  // namespace hog

// ---------------------------------------------------------------------------------------------
// C.4 SVM — on-disk format svm_032015_linear_20_20_same:1-16,780-789; CvSVM::predict semantics
// ---------------------------------------------------------------------------------------------
struct Svm {
  int kernel = 0;  // 0 LINEAR, 1 POLY (OpenCV enum values)
  double degree = 0, gamma = 1, coef0 = 0, rho = 0;
  int var_count = 0, sv_total = 0, sv_count = 0;
  std::vector<float> sv;
  std::vector<double> alpha;
  std::vector<int> index;
  int class_labels[2] = {-1, 1};
};

static bool read_number_list(const std::string& text, size_t& pos, std::vector<double>& out) {
  // reads "[ a, b, c ]" starting at the first '[' at/after pos
  size_t lb = text.find('[', pos);
  if (lb == std::string::npos) return false;
  size_t rb = text.find(']', lb);
  if (rb == std::string::npos) return false;
  const char* p = text.c_str() + lb + 1;
  const char* end = text.c_str() + rb;
  while (p < end) {
    while (p < end && (*p == ' ' || *p == ',' || *p == '\n' || *p == '\r' || *p == '\t')) p++;
    if (p >= end) break;
    char* q;
    double v = std::strtod(p, &q);
    if (q == p) return false;
    out.push_back(v);
    p = q;
  }
  pos = rb + 1;
  return true;
}
static bool find_scalar(const std::string& text, const std::string& key, size_t from, double& v, size_t* at = nullptr) {
  size_t p = text.find(key + ":", from);
  if (p == std::string::npos) return false;
  const char* s = text.c_str() + p + key.size() + 1;
  char* q;
  v = std::strtod(s, &q);
  if (q == s) return false;
  if (at) *at = p;
  return true;
}

Svm* svm_load(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f.good()) {
    fail(std::string("File ") + path + " does not exist!");  // learning.cpp:172-178
    return nullptr;
  }
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string text = ss.str();
  if (text.find("opencv-ml-svm") == std::string::npos) {
    fail("not an opencv-ml-svm file");
    return nullptr;
  }
  Svm* s = new Svm;
  size_t kp = text.find("kernel:");
  std::string kline = text.substr(kp, text.find('}', kp) - kp);
  if (kline.find("LINEAR") != std::string::npos) s->kernel = 0;
  else if (kline.find("POLY") != std::string::npos) s->kernel = 1;
  else {
    fail("unsupported SVM kernel (only LINEAR and POLY)");
    delete s;
    return nullptr;
  }
  double v;
  if (s->kernel == 1) {
    size_t kb = kp;
    if (find_scalar(text, "degree", kb, v)) s->degree = v;
    if (find_scalar(text, "gamma", kb, v)) s->gamma = v;
    if (find_scalar(text, "coef0", kb, v)) s->coef0 = v;
  }
  if (!find_scalar(text, "var_count", 0, v)) { fail("var_count missing"); delete s; return nullptr; }
  s->var_count = int(v);
  if (!find_scalar(text, "sv_total", 0, v)) { fail("sv_total missing"); delete s; return nullptr; }
  s->sv_total = int(v);
  size_t pos = text.find("support_vectors:");
  s->sv.reserve(size_t(s->sv_total) * s->var_count);
  for (int k = 0; k < s->sv_total; k++) {
    std::vector<double> row;
    if (!read_number_list(text, pos, row) || int(row.size()) != s->var_count) {
      fail("bad support vector row");
      delete s;
      return nullptr;
    }
    for (double d : row) s->sv.push_back(float(d));  // stored as CV_32F
  }
  size_t df = text.find("decision_functions:", pos);
  if (!find_scalar(text, "sv_count", df, v)) { fail("sv_count missing"); delete s; return nullptr; }
  s->sv_count = int(v);
  if (!find_scalar(text, "rho", df, v)) { fail("rho missing"); delete s; return nullptr; }
  s->rho = v;
  size_t ap = text.find("alpha:", df);
  std::vector<double> a;
  if (!read_number_list(text, ap, a) || int(a.size()) != s->sv_count) { fail("bad alpha"); delete s; return nullptr; }
  s->alpha = a;
  size_t ip = text.find("index:", ap);
  s->index.resize(s->sv_count);
  for (int k = 0; k < s->sv_count; k++) s->index[k] = k;
  if (ip != std::string::npos) {
    std::vector<double> idx;
    if (read_number_list(text, ip, idx) && int(idx.size()) == s->sv_count)
      for (int k = 0; k < s->sv_count; k++) s->index[k] = int(idx[k]);
  }
  return s;
}

float svm_decision(const Svm& s, const float* x) {
  // calc_non_rbf_base: float products, 4-term float sums, double accumulation; Qfloat = float
  std::vector<float> K(s.sv_total);
  const int vc = s.var_count;
  const double alpha = s.kernel == 0 ? 1.0 : s.gamma, beta = s.kernel == 0 ? 0.0 : s.coef0;
  for (int j = 0; j < s.sv_total; j++) {
    const float* sv = &s.sv[size_t(j) * vc];
    double acc = 0;
    int k = 0;
    for (; k <= vc - 4; k += 4)
      acc += sv[k] * x[k] + sv[k + 1] * x[k + 1] + sv[k + 2] * x[k + 2] + sv[k + 3] * x[k + 3];
    for (; k < vc; k++) acc += sv[k] * x[k];
    K[j] = float(acc * alpha + beta);
  }
  if (s.kernel == 1) {
    // cv::pow(R, degree, R) on CV_32F; integer degree -> repeated multiplication
    int ip = int(s.degree);
    for (float& r : K) {
      float b = r, a = 1.f;
      int p = ip;
      while (p > 1) {
        if (p & 1) a *= b;
        b *= b;
        p >>= 1;
      }
      r = a * b;
    }
  }
  double sum = -s.rho;
  for (int k = 0; k < s.sv_count; k++) sum += s.alpha[k] * double(K[s.index[k]]);
  return float(sum);
}

}  // namespace ago

using namespace ago;

struct ago_tree;
struct ago_hands {
  Hands* h;
};
struct ago_svm {
  Svm* s;
};

extern "C" {

int ago_points_image(const double* pts3xm, int m, const double binormal[3], const double surface[3],
                     const double cam_pos[3], uint8_t* image80x100) {
  points_image(pts3xm, m, binormal, surface, cam_pos, image80x100);
  return 0;
}

int ago_grasp_image(const ago_hands* h, int k, const ag_params* P, uint8_t* image) {
  if (k < 0 || k >= int(h->h->grasps.size())) return fail("hypothesis index out of range");
  const ag_grasp& g = h->h->grasps[k];
  const double* tf = g.cam_source == 1 ? P->cam_tf_right : P->cam_tf_left;  // learning.cpp:203,382
  const double cam_pos[3] = {tf[3], tf[7], tf[11]};
  points_image(h->h->pts[k].data(), int(h->h->pcam[k].size()), g.binormal, g.surface, cam_pos, image);
  return 0;
}

// Learning::createInstance(h, cam_pos, cam) + convertToImage (learning.cpp:375-400,320-365): cam = -1 all box
// points, cam = 0 / 1 only the points seen by camera 1 / 2 (the "simulated camera" instances of Learning::train,
// learning.cpp:76-141)
int ago_grasp_image_cam(const ago_hands* h, int k, int cam, const ag_params* P, uint8_t* image) {
  if (k < 0 || k >= int(h->h->grasps.size())) return fail("hypothesis index out of range");
  const ag_grasp& g = h->h->grasps[k];
  const double* tf = g.cam_source == 1 ? P->cam_tf_right : P->cam_tf_left;  // learning.cpp:382-383
  const double cam_pos[3] = {tf[3], tf[7], tf[11]};
  const std::vector<double>& pts = h->h->pts[k];
  const std::vector<int32_t>& pc = h->h->pcam[k];
  if (cam < 0) {
    points_image(pts.data(), int(pc.size()), g.binormal, g.surface, cam_pos, image);
    return 0;
  }
  std::vector<double> sub;  // columns listed by getIndicesPointsForLearningCam1/2 (rotating_hand.cpp:143-151)
  for (size_t j = 0; j < pc.size(); j++)
    if (pc[j] == cam) sub.insert(sub.end(), pts.begin() + 3 * j, pts.begin() + 3 * j + 3);
  points_image(sub.data(), int(sub.size() / 3), g.binormal, g.surface, cam_pos, image);
  return 0;
}

int ago_hog(const uint8_t* image80x100, float* desc3528) {
  hog::compute(image80x100, desc3528);
  return 0;
}

ago_svm* ago_svm_load(const char* path) {
  Svm* s = svm_load(path);
  if (!s) return nullptr;
  ago_svm* r = new ago_svm;
  r->s = s;
  return r;
}
void ago_svm_free(ago_svm* s) {
  if (!s) return;
  delete s->s;
  delete s;
}
int ago_svm_info(const ago_svm* s, int* kernel_type, int* var_count, int* sv_total, double* rho, int* degree,
                 double* gamma, double* coef0) {
  if (kernel_type) *kernel_type = s->s->kernel;
  if (var_count) *var_count = s->s->var_count;
  if (sv_total) *sv_total = s->s->sv_total;
  if (rho) *rho = s->s->rho;
  if (degree) *degree = int(s->s->degree);
  if (gamma) *gamma = s->s->gamma;
  if (coef0) *coef0 = s->s->coef0;
  return 0;
}
const float* ago_svm_sv(const ago_svm* s) { return s->s->sv.data(); }
const double* ago_svm_alpha(const ago_svm* s) { return s->s->alpha.data(); }
float ago_svm_decision(const ago_svm* s, const float* x) { return svm_decision(*s->s, x); }

int ago_classify(ago_hands* h, const ago_svm* s, const ag_params* P, uint8_t* keep) {
  // learning.cpp:198-243
  const int n = int(h->h->grasps.size());
  if (s->s->var_count != AG_HOG_DIM) return fail("SVM var_count != 3528");
  int threads = std::max(1, P->num_threads);
#pragma omp parallel for num_threads(threads) schedule(dynamic, 16)
  for (int i = 0; i < n; i++) {
    uint8_t img[AG_IMAGE_ROWS * AG_IMAGE_COLS];
    ago_grasp_image(h, i, P, img);
    std::vector<float> desc(AG_HOG_DIM);
    hog::compute(img, desc.data());
    float sum = svm_decision(*s->s, desc.data());
    ag_grasp& g = h->h->grasps[i];
    g.score = sum;
    // CvSVM::predict: vote[sum > 0 ? 0 : 1], class_labels = [-1, 1]; kept iff prediction == 1
    g.label = sum > 0 ? 0 : 1;
    if (keep) keep[i] = g.label;
  }
  return 0;
}

// uses_clustering (localization.cpp:51-98): remove the dominant plane found by RANSAC (pcl::SACSegmentation,
// SACMODEL_PLANE / SAC_RANSAC, 100 iterations, distance threshold 0.01, optimised coefficients) from the
// voxelised cloud.  PCL is not vendored and draws its samples from its own generator, so WHICH triples are drawn
// is restated, not reproduced: iteration t draws three distinct points from splitmix64(seed, t); all 100
// iterations run (PCL may stop earlier once its inlier ratio makes more draws pointless — more draws can only
// find an equal or better plane).  The rest follows PCL: score = number of points with |n.p + d| < threshold,
// first best wins; coefficients refitted to the winner's inliers (centroid + smallest eigenvector of the
// covariance); final inliers of the refitted plane removed.
static uint64_t sm64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
void ransac_triple(int n, uint64_t seed, int t, int idx[3]) {
  uint64_t s = sm64(seed ^ sm64(uint64_t(t) + 0x1234567ull));
  for (int k = 0; k < 3; k++) {
    for (;;) {
      s = sm64(s);
      const int v = int(s % uint64_t(n));
      bool dup = false;
      for (int q = 0; q < k; q++) dup = dup || idx[q] == v;
      if (!dup) {
        idx[k] = v;
        break;
      }
    }
  }
}
// plane through three points: unit normal and offset; false if (nearly) collinear
bool plane_from_triple(const float* xyz, const int idx[3], double pl[4]) {
  const double p0[3] = {xyz[3 * idx[0]], xyz[3 * idx[0] + 1], xyz[3 * idx[0] + 2]};
  double a[3], b[3];
  for (int d = 0; d < 3; d++) {
    a[d] = double(xyz[3 * idx[1] + d]) - p0[d];
    b[d] = double(xyz[3 * idx[2] + d]) - p0[d];
  }
  const double nx = a[1] * b[2] - a[2] * b[1], ny = a[2] * b[0] - a[0] * b[2], nz = a[0] * b[1] - a[1] * b[0];
  const double len = std::sqrt(nx * nx + ny * ny + nz * nz);
  if (!(len > 1e-12)) return false;
  pl[0] = nx / len;
  pl[1] = ny / len;
  pl[2] = nz / len;
  pl[3] = -1.0 * (pl[0] * p0[0] + pl[1] * p0[1] + pl[2] * p0[2]);
  return true;
}
static inline bool plane_inlier(const double pl[4], const float* p, double thresh) {
  const double d = (pl[0] * double(p[0]) + pl[1] * double(p[1])) + (pl[2] * double(p[2]) + pl[3]);
  return std::fabs(d) < thresh;
}
// refit to the inliers: centroid and covariance accumulated in index order, smallest eigenvector (cyclic Jacobi)
bool refit_plane(const float* xyz, int n, const double pl[4], double thresh, double out[4]) {
  double c[3] = {0, 0, 0};
  long long m = 0;
  for (int i = 0; i < n; i++)
    if (plane_inlier(pl, xyz + 3 * i, thresh)) {
      for (int d = 0; d < 3; d++) c[d] += double(xyz[3 * i + d]);
      m++;
    }
  if (m < 3) return false;
  for (int d = 0; d < 3; d++) c[d] /= double(m);
  double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < n; i++)
    if (plane_inlier(pl, xyz + 3 * i, thresh)) {
      const double v[3] = {double(xyz[3 * i]) - c[0], double(xyz[3 * i + 1]) - c[1], double(xyz[3 * i + 2]) - c[2]};
      for (int r = 0; r < 3; r++)
        for (int q = 0; q < 3; q++) C[r][q] += v[r] * v[q];
    }
  // cyclic Jacobi on the symmetric 3x3
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 60; sweep++) {
    const double off = C[0][1] * C[0][1] + C[0][2] * C[0][2] + C[1][2] * C[1][2];
    if (off <= 1e-40) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        if (C[p][q] == 0.0) continue;
        const double theta = (C[q][q] - C[p][p]) / (2.0 * C[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double cs = 1.0 / std::sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 3; k++) {
          const double akp = C[k][p], akq = C[k][q];
          C[k][p] = cs * akp - sn * akq;
          C[k][q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < 3; k++) {
          const double apk = C[p][k], aqk = C[q][k];
          C[p][k] = cs * apk - sn * aqk;
          C[q][k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = cs * vkp - sn * vkq;
          V[k][q] = sn * vkp + cs * vkq;
        }
      }
  }
  int mi = 0;
  if (C[1][1] < C[mi][mi]) mi = 1;
  if (C[2][2] < C[mi][mi]) mi = 2;
  const double len = std::sqrt(V[0][mi] * V[0][mi] + V[1][mi] * V[1][mi] + V[2][mi] * V[2][mi]);
  for (int d = 0; d < 3; d++) out[d] = V[d][mi] / len;
  out[3] = -1.0 * (out[0] * c[0] + out[1] * c[1] + out[2] * c[2]);
  return true;
}

int ago_remove_plane(const float* xyz, int n, uint64_t seed, int max_iterations, double thresh,
                                uint8_t* keep, int32_t* counts_out, double* plane_out) {
  for (int i = 0; i < n; i++) keep[i] = 1;
  if (n < 3) return 1;
  int best_t = -1, best_count = 0;
  double best_pl[4] = {0, 0, 0, 0};
  for (int t = 0; t < max_iterations; t++) {
    int idx[3];
    ransac_triple(n, seed, t, idx);
    double pl[4];
    int cnt = 0;
    if (plane_from_triple(xyz, idx, pl))
      for (int i = 0; i < n; i++) cnt += plane_inlier(pl, xyz + 3 * i, thresh) ? 1 : 0;
    if (counts_out) counts_out[t] = cnt;
    if (cnt > best_count) {
      best_count = cnt;
      best_t = t;
      std::memcpy(best_pl, pl, sizeof(pl));
    }
  }
  if (best_t < 0) return 1;  // "Could not estimate a planar model for the given dataset."
  double ref[4];
  if (!refit_plane(xyz, n, best_pl, thresh, ref)) std::memcpy(ref, best_pl, sizeof(ref));
  if (plane_out) std::memcpy(plane_out, ref, sizeof(ref));
  for (int i = 0; i < n; i++) keep[i] = plane_inlier(ref, xyz + 3 * i, thresh) ? 0 : 1;
  return 0;
}

ago_hands* ago_localize(const void* points, int stride, int n_in, int size_left, const ag_params* P,
                        const int* indices, int n_indices, unsigned flags, const ago_svm* svm, int use_std_set,
                        double* times_ms, int* n_voxels_out) {
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  double tm[7] = {0, 0, 0, 0, 0, 0, 0};
  auto t0 = now();
  std::vector<float> xyz;
  std::vector<int32_t> cam;
  if (preprocess(points, stride, n_in, size_left, *P, use_std_set != 0, xyz, cam) != 0) return nullptr;
  if (flags & AG_FLAG_USE_CLUSTERING) {  // localization.cpp:51-98
    std::vector<uint8_t> keep(cam.size());
    if (ago_remove_plane(xyz.data(), int(cam.size()), P->seed, 100, 0.01, keep.data(), nullptr, nullptr) != 0) {
      ago_hands* none = new ago_hands;
      none->h = new Hands;
      if (n_voxels_out) *n_voxels_out = 0;
      return none;  // "Could not estimate a planar model": empty hand list (localization.cpp:69-74)
    }
    size_t w = 0;
    for (size_t i = 0; i < keep.size(); i++)
      if (keep[i]) {
        for (int d = 0; d < 3; d++) xyz[3 * w + d] = xyz[3 * i + d];
        cam[w++] = cam[i];
      }
    xyz.resize(3 * w);
    cam.resize(w);
  }
  const int n = int(cam.size());
  if (n_voxels_out) *n_voxels_out = n;
  auto t1 = now();
  tm[0] = ms(t0, t1);
  ago_tree* tree = ago_tree_build(xyz.data(), n);  // hand_search.cpp:10-11
  auto t2 = now();
  tm[1] = ms(t1, t2);
  const Tree* T = tree_of(tree);
  std::vector<double> normals(size_t(3) * n, 0.0);       // hand_search.cpp:13-14
  size_t rand_consumed = 0;  // rand() draws of the production normal mode so far in this call
  if (flags & AG_FLAG_CALC_ANTIPODAL) {                  // hand_search.cpp:17-26
    std::vector<int> all(n);
    for (int i = 0; i < n; i++) all[i] = i;
    std::vector<ag_frame> fr(n);
    if (fit_quadrics(xyz.data(), cam.data(), n, T, all.data(), n, P->nn_radius_normals, *P, 0, fr.data(), nullptr,
                     nullptr, nullptr, &rand_consumed) != 0) {
      ago_tree_free(tree);
      return nullptr;
    }
    for (int i = 0; i < n; i++)
      if (fr[i].num_neighbors > 0)
        for (int d = 0; d < 3; d++) normals[size_t(3) * i + d] = fr[i].normal[d];
  }
  auto t3 = now();
  tm[2] = ms(t2, t3);
  std::vector<int32_t> idx;
  if (indices && n_indices > 0) idx.assign(indices, indices + n_indices);
  else {
    idx.resize(std::min(n, P->num_samples));
    draw_samples(n, P->num_samples, P->seed, idx.data());
  }
  const int S = int(idx.size());
  for (int i : idx)
    if (i < 0 || i >= n) {
      fail("sample index out of range");
      ago_tree_free(tree);
      return nullptr;
    }
  std::vector<int32_t> scam(S);
  for (int i = 0; i < S; i++) scam[i] = cam[idx[i]];  // hand_search.cpp:40-42 (+ App. B#3)
  std::vector<ag_frame> frames(S);
  if (fit_quadrics(xyz.data(), cam.data(), n, T, idx.data(), S, P->nn_radius_taubin, *P, 0, frames.data(), nullptr,
                   nullptr, nullptr, &rand_consumed) != 0) {
    ago_tree_free(tree);
    return nullptr;
  }
  for (int i = 0; i < S; i++)  // hand_search.cpp:102 sample normals leak into cloud_normals_ (App. B#11)
    if (frames[i].num_neighbors > 0)
      for (int d = 0; d < 3; d++) normals[size_t(3) * idx[i] + d] = frames[i].normal[d];
  auto t4 = now();
  tm[3] = ms(t3, t4);
  Hands* H = find_hands(xyz.data(), cam.data(), n, T, idx.data(), S, frames.data(), scam.data(), normals.data(), *P);
  H->n_voxels = n;
  if (P->filters_boundaries) {  // localization.cpp:116-121
    std::vector<uint8_t> keep(H->grasps.size());
    filter_hands(H->grasps.data(), int(H->grasps.size()), *P, keep.data());
    Hands* F = new Hands;
    F->status = H->status; F->hand_idx = H->hand_idx; F->depth_steps = H->depth_steps;
    F->finger_mask = H->finger_mask; F->num_slab = H->num_slab; F->n_voxels = n;
    for (size_t k = 0; k < keep.size(); k++)
      if (keep[k]) {
        F->grasps.push_back(H->grasps[k]);
        F->grasps.back().image_id = int(F->grasps.size()) - 1;
        F->pts.push_back(std::move(H->pts[k]));
        F->pcam.push_back(std::move(H->pcam[k]));
      }
    delete H;
    H = F;
  }
  auto t5 = now();
  tm[4] = ms(t4, t5);
  ago_tree_free(tree);
  ago_hands* out = new ago_hands;
  out->h = H;
  if (svm) ago_classify(out, svm, P, nullptr);
  auto t6 = now();
  tm[5] = ms(t5, t6);
  tm[6] = ms(t0, t6);
  if (times_ms) std::memcpy(times_ms, tm, sizeof(tm));
  return out;
}

}  // extern "C"
