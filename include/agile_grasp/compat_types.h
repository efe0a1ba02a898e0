// compat_types.h — layout-compatible stand-ins for the third-party types that cross the reference's
// Localization API (include/agile_grasp/localization.h:57,75-265 of the reference): a subset of Eigen's
// fixed/dynamic vectors and matrices and pcl::PointCloud<pcl::PointXYZRGBA>.  A catkin workspace that has
// the real libraries defines AG_HAVE_EIGEN / AG_HAVE_PCL and gets the real headers instead.
#ifndef AGILE_GRASP_COMPAT_TYPES_H_
#define AGILE_GRASP_COMPAT_TYPES_H_

#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <memory>
#include <vector>

#ifdef AG_HAVE_EIGEN
#include <Eigen/Dense>
#else
namespace Eigen {
template <typename T, int R, int C>
class Matrix {  // column-major, like Eigen's default
 public:
  Matrix() : rows_(R > 0 ? R : 0), cols_(C > 0 ? C : 0), d_(size_t(rows_) * cols_, T(0)) {}
  Matrix(int r, int c) : rows_(r), cols_(c), d_(size_t(r) * c, T(0)) {}
  explicit Matrix(int n) : rows_(C == 1 ? n : R), cols_(C == 1 ? 1 : n), d_(size_t(rows_) * cols_, T(0)) {}
  int rows() const { return rows_; }
  int cols() const { return cols_; }
  int size() const { return rows_ * cols_; }
  void resize(int r, int c) { rows_ = r; cols_ = c; d_.assign(size_t(r) * c, T(0)); }
  T& operator()(int r, int c) { return d_[size_t(c) * rows_ + r]; }
  const T& operator()(int r, int c) const { return d_[size_t(c) * rows_ + r]; }
  T& operator()(int i) { return d_[i]; }
  const T& operator()(int i) const { return d_[i]; }
  T& operator[](int i) { return d_[i]; }
  const T& operator[](int i) const { return d_[i]; }
  T* data() { return d_.data(); }
  const T* data() const { return d_.data(); }
  // Eigen's comma initialiser: `m << a, b, c, ...;` fills row by row
  struct CommaInit {
    Matrix& m;
    int i;
    CommaInit& operator,(T x) { m.set_rowmajor(i++, x); return *this; }
  };
  CommaInit operator<<(T x) { set_rowmajor(0, x); return CommaInit{*this, 1}; }
  void set_rowmajor(int i, T x) { (*this)(i / cols_, i % cols_) = x; }
  static Matrix Identity() { Matrix m; for (int i = 0; i < m.rows_ && i < m.cols_; i++) m(i, i) = T(1); return m; }
 private:
  int rows_, cols_;
  std::vector<T> d_;
};
enum { Dynamic = -1 };
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<double, 3, Dynamic> Matrix3Xd;
}  // namespace Eigen
#endif

#ifdef AG_HAVE_PCL
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#else
namespace pcl {
struct alignas(16) PointXYZRGBA {  // 32 bytes: x y z pad | rgba pad pad pad  (PCL_ADD_POINT4D + PCL_ADD_RGB)
  float x, y, z, data3;
  uint32_t rgba;
  uint32_t pad_[3];
};
static_assert(sizeof(PointXYZRGBA) == 32, "pcl::PointXYZRGBA layout");
template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT> > Ptr;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  size_t size() const { return points.size(); }
  void resize(size_t n) { points.resize(n); }
};
}  // namespace pcl
#endif

#endif
