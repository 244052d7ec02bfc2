// opencv2/opencv.hpp stand-in (OpenCV is not installed in this image): the handful of cv:: types the reference's wrapper
// and offline drivers mention -- an owning 2-D byte matrix, DMatch, KeyPoint and no-op window calls.  Enough to compile
// and to run the image-free (VRO-log + IMU-log) path; the front end that would fill these images is out of scope.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
typedef unsigned char uchar;
typedef unsigned short ushort;
namespace cv {
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
template <class T> struct Point_ { T x = 0, y = 0; Point_() {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
typedef Point_<int> Point;
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
struct DMatch { int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0; };
struct Scalar { double v[4] = {0, 0, 0, 0}; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
class Mat {
  std::shared_ptr<std::vector<unsigned char>> buf_;
  int type_ = 0;
  static int elem(int type) { static const int sz[] = {1, 1, 2, 2, 4, 4, 8}; return sz[type & 7] * ((type >> 3) + 1); }
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size s, int type) { create(s.height, s.width, type); }
  Mat(int r, int c, int type, const Scalar& s) { create(r, c, type); setTo(s); }
  void create(int r, int c, int type) { rows = r; cols = c; type_ = type; buf_.reset(new std::vector<unsigned char>((size_t)r * c * elem(type), 0)); data = buf_->data(); }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  static Mat zeros(Size s, int type) { return Mat(s, type); }
  Mat clone() const { Mat m; m.rows = rows; m.cols = cols; m.type_ = type_; if (buf_) { m.buf_.reset(new std::vector<unsigned char>(*buf_)); m.data = m.buf_->data(); } return m; }
  void copyTo(Mat& o) const { o = clone(); }
  void setTo(const Scalar& s) { if (!buf_) return; const int ch = channels(); for (size_t i = 0; i < buf_->size() / elem(type_); ++i) for (int c = 0; c < ch; ++c) if ((type_ & 7) == CV_8U) (*buf_)[i * ch + c] = (unsigned char)s.v[c]; }
  Size size() const { return Size(cols, rows); }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  bool empty() const { return !buf_ || buf_->empty(); }
  size_t total() const { return (size_t)rows * cols; }
  size_t elemSize() const { return elem(type_); }
  template <class T> T& at(int i) { return reinterpret_cast<T*>(data)[i]; }
  template <class T> const T& at(int i) const { return reinterpret_cast<const T*>(data)[i]; }
  template <class T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + ((size_t)r * cols + c) * elem(type_)); }
  template <class T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + ((size_t)r * cols + c) * elem(type_)); }
  template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * cols * elem(type_)); }
  template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * cols * elem(type_)); }
};
inline void namedWindow(const std::string&, int = 0) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }
inline void destroyWindow(const std::string&) {}
inline void destroyAllWindows() {}
inline bool imwrite(const std::string&, const Mat&) { return false; }
inline Mat imread(const std::string&, int = 1) { return Mat(); }
}  // namespace cv
