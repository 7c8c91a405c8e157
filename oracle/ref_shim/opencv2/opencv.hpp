// ORACLE BUILD SHIM (test infrastructure, not product).
// The reference's vendored GMS matcher (/root/reference/src/utils/GMSMatcher/gms_matcher.{h,cpp}) only needs a handful
// of OpenCV value types; OpenCV's C++ headers are not installed here, so this header supplies exactly those -- with
// OpenCV's semantics -- so that the UNMODIFIED reference sources compile where they lie (oracle/Makefile) into
// oracle/_ref/libgms_ref.so.  Nothing of the reference is copied: this file only declares stand-ins for cv::Mat
// (CV_32SC1, zeros / ptr / at / row / setTo), cv::sum, cv::Size, cv::Point2f, cv::KeyPoint, cv::DMatch.
#pragma once
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#define CV_32SC1 4

namespace cv {

struct Size {
  int width = 0, height = 0;
  Size() {}
  Size(int w, int h) : width(w), height(h) {}
};

struct Point2f {
  float x = 0.f, y = 0.f;
  Point2f() {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct KeyPoint {
  Point2f pt;
};

struct DMatch {
  int queryIdx = -1, trainIdx = -1, imgIdx = -1;
  float distance = 0.f;
};

struct Scalar {
  double v[4] = {0, 0, 0, 0};
  double operator[](int i) const { return v[i]; }
};

// int32 single-channel matrix with shared storage (rows() views alias the parent, like cv::Mat headers)
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  static Mat zeros(int r, int c, int /*type*/) {
    Mat m;
    m.rows = r;
    m.cols = c;
    m.store_ = std::make_shared<std::vector<int>>((size_t)r * c, 0);
    m.data_ = m.store_->data();
    return m;
  }
  template <typename T>
  T* ptr(int r) {
    return reinterpret_cast<T*>(data_ + (size_t)r * cols);
  }
  template <typename T>
  const T* ptr(int r) const {
    return reinterpret_cast<const T*>(data_ + (size_t)r * cols);
  }
  template <typename T>
  T& at(int r, int c) {
    return reinterpret_cast<T*>(data_)[(size_t)r * cols + c];
  }
  Mat row(int r) const {
    Mat m;
    m.rows = 1;
    m.cols = cols;
    m.store_ = store_;
    m.data_ = data_ + (size_t)r * cols;
    return m;
  }
  Mat& setTo(int v) {
    for (size_t i = 0; i < (size_t)rows * cols; ++i) data_[i] = v;
    return *this;
  }
  const int* raw() const { return data_; }

 private:
  std::shared_ptr<std::vector<int>> store_;
  int* data_ = nullptr;
};

inline Scalar sum(const Mat& m) {
  Scalar s;
  for (size_t i = 0; i < (size_t)m.rows * m.cols; ++i) s.v[0] += m.raw()[i];
  return s;
}
inline Scalar sum(const std::vector<bool>& v) {  // cv::sum over a vector<bool> InputArray counts the true entries
  Scalar s;
  for (bool b : v) s.v[0] += b ? 1 : 0;
  return s;
}

}  // namespace cv
