// oracle/shim/opencv2/opencv.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A private stand-in for the slice of the OpenCV C++ API that the reference's
// Transform360/Library/VideoFrameTransform.{h,cpp} touches, so that the
// UNMODIFIED reference sources compile here (no C++ OpenCV headers exist in
// this image).  Geometry and the low-pass plan are
// computed entirely by the reference code with this header; the three pixel
// routines it calls (cv::remap, cv::sepFilter2D, cv::resize) are forwarded to
// hook function pointers, which the Python harness (oracle/ref_harness.py)
// points at the real OpenCV (cv2 4.13.0 wheel).  If no hook is installed the
// calls throw, which the reference turns into "return false".
//
// Nothing under transform360_b200/ includes or links this file.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>  // the reference calls memcpy without including it (cpp:207)
#include <memory>
#include <stdexcept>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32F 5
#define CV_32FC2 13

extern "C" {
// Hook signatures (set through t360ref_set_hooks in ref_driver.cpp).
// remap: src plane, dst plane, float2 map (rows x cols == dst size), interpolation, border.
typedef int (*t360ref_remap_hook)(
    const uchar* src, int srcRows, int srcCols, size_t srcStep,
    uchar* dst, int dstRows, int dstCols, size_t dstStep,
    const float* mapxy, size_t mapStep, int interpolation, int borderMode);
// sepFilter2D on a NON-ISOLATED roi: parent plane pointer + dims, roi rect, dst roi pointer.
typedef int (*t360ref_sep_hook)(
    const uchar* parent, int parentRows, int parentCols, size_t parentStep,
    int roiX, int roiY, int roiW, int roiH,
    uchar* dstRoi, size_t dstStep,
    const float* kx, int nkx, const float* ky, int nky, int borderMode);
typedef int (*t360ref_resize_hook)(
    const uchar* src, int srcRows, int srcCols, size_t srcStep,
    uchar* dst, int dstRows, int dstCols, size_t dstStep, int interpolation);
extern t360ref_remap_hook g_t360ref_remap;
extern t360ref_sep_hook g_t360ref_sep;
extern t360ref_resize_hook g_t360ref_resize;
}

namespace cv {

enum {
  BORDER_CONSTANT = 0,
  BORDER_REPLICATE = 1,
  BORDER_REFLECT = 2,
  BORDER_WRAP = 3,
  BORDER_REFLECT_101 = 4,
  BORDER_TRANSPARENT = 5
};
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3, INTER_LANCZOS4 = 4 };

struct Point {
  int x, y;
  Point() : x(0), y(0) {}
  Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Point2f {
  float x, y;
  Point2f() : x(0), y(0) {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};
struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Scalar {
  double val[4];
  Scalar(double v0 = 0, double v1 = 0, double v2 = 0, double v3 = 0) {
    val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3;
  }
};

class Mat {
 public:
  int rows, cols;
  uchar* data;
  size_t step;
  // whole-matrix bookkeeping so that a ROI stays "non-isolated" like cv::Mat
  uchar* datastart;
  int wholeRows, wholeCols;
  int ofsX, ofsY;

  Mat() : rows(0), cols(0), data(nullptr), step(0), datastart(nullptr),
          wholeRows(0), wholeCols(0), ofsX(0), ofsY(0), type_(CV_8U) {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* ext, size_t extStep)
      : rows(r), cols(c), data(static_cast<uchar*>(ext)), step(extStep),
        datastart(static_cast<uchar*>(ext)), wholeRows(r), wholeCols(c),
        ofsX(0), ofsY(0), type_(type) {}
  Mat(Size sz, int type, const Scalar& s) {
    create(sz.height, sz.width, type);
    setTo(s);
  }

  static Mat zeros(int r, int c, int type) {
    Mat m(r, c, type);
    if (m.data) std::memset(m.data, 0, m.step * static_cast<size_t>(r));
    return m;
  }
  static Mat zeros(Size sz, int type) { return zeros(sz.height, sz.width, type); }

  int type() const { return type_; }
  Size size() const { return Size(cols, rows); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t elemSize() const { return esz(type_); }

  template <typename T> T& at(int i, int j) {
    return *reinterpret_cast<T*>(data + static_cast<size_t>(i) * step + sizeof(T) * static_cast<size_t>(j));
  }
  template <typename T> const T& at(int i, int j) const {
    return *reinterpret_cast<const T*>(data + static_cast<size_t>(i) * step + sizeof(T) * static_cast<size_t>(j));
  }

  Mat operator()(const Rect& r) const {
    if (r.x < 0 || r.y < 0 || r.width < 0 || r.height < 0 || r.x + r.width > cols || r.y + r.height > rows)
      throw std::runtime_error("shim cv::Mat: roi out of range");
    Mat m(*this);
    m.rows = r.height;
    m.cols = r.width;
    m.data = data + static_cast<size_t>(r.y) * step + static_cast<size_t>(r.x) * esz(type_);
    m.ofsX = ofsX + r.x;
    m.ofsY = ofsY + r.y;
    return m;
  }

  Mat& setTo(const Scalar& s) {
    for (int i = 0; i < rows; ++i) {
      uchar* p = data + static_cast<size_t>(i) * step;
      if (type_ == CV_8U) {
        double v = s.val[0];
        int iv = static_cast<int>(std::lrint(v));
        std::memset(p, iv < 0 ? 0 : (iv > 255 ? 255 : iv), static_cast<size_t>(cols));
      } else {
        int cn = (type_ == CV_32FC2) ? 2 : 1;
        float* f = reinterpret_cast<float*>(p);
        for (int j = 0; j < cols; ++j)
          for (int c = 0; c < cn; ++c) f[j * cn + c] = static_cast<float>(s.val[c]);
      }
    }
    return *this;
  }

 private:
  static size_t esz(int type) { return type == CV_8U ? 1 : (type == CV_32F ? 4 : 8); }
  void create(int r, int c, int type) {
    // OpenCV refuses negative sizes (cv::Mat::create -> CV_Assert / error -211); the reference then returns false, e.g.
    // when a band's horizontal sigma comes out negative because cosf(angle) is -4e-8 at the pole (cpp:219, 78-81)
    if (r < 0 || c < 0) throw std::runtime_error("shim cv::Mat: negative size");
    rows = r; cols = c; type_ = type;
    step = esz(type) * static_cast<size_t>(c);
    size_t bytes = step * static_cast<size_t>(r);
    store_ = std::shared_ptr<uchar>(static_cast<uchar*>(std::malloc(bytes ? bytes : 1)), std::free);
    data = datastart = store_.get();
    wholeRows = r; wholeCols = c; ofsX = ofsY = 0;
  }
  int type_;
  std::shared_ptr<uchar> store_;
};

// cv::Mat::operator/=(double) is convertTo(self, -1, 1./s); for CV_32F data the
// scale is applied in single precision.
inline Mat& operator/=(Mat& m, double s) {
  const float a = static_cast<float>(1.0 / s);
  for (int i = 0; i < m.rows; ++i)
    for (int j = 0; j < m.cols; ++j) m.at<float>(i, j) = m.at<float>(i, j) * a;
  return m;
}

inline void sepFilter2D(const Mat& src, Mat dst, int /*ddepth*/, const Mat& kx, const Mat& ky,
                        Point /*anchor*/, double /*delta*/, int borderType) {
  if (!g_t360ref_sep) throw std::runtime_error("shim: no sepFilter2D hook installed");
  int rc = g_t360ref_sep(src.datastart, src.wholeRows, src.wholeCols, src.step, src.ofsX, src.ofsY,
                         src.cols, src.rows, dst.data, dst.step,
                         reinterpret_cast<const float*>(kx.data), kx.rows * kx.cols,
                         reinterpret_cast<const float*>(ky.data), ky.rows * ky.cols, borderType);
  if (rc != 0) throw std::runtime_error("shim: sepFilter2D hook failed");
}

inline void remap(const Mat& src, Mat dst, const Mat& map1, const Mat& /*map2*/, int interpolation,
                  int borderMode) {
  if (map1.empty()) throw std::runtime_error("shim: remap with empty map");
  if (dst.rows != map1.rows || dst.cols != map1.cols)
    throw std::runtime_error("shim: remap dst size differs from map size");
  if (!g_t360ref_remap) throw std::runtime_error("shim: no remap hook installed");
  int rc = g_t360ref_remap(src.data, src.rows, src.cols, src.step, dst.data, dst.rows, dst.cols, dst.step,
                           reinterpret_cast<const float*>(map1.data), map1.step, interpolation, borderMode);
  if (rc != 0) throw std::runtime_error("shim: remap hook failed");
}

inline void resize(const Mat& src, Mat dst, Size /*dsize*/, double, double, int interpolation) {
  if (!g_t360ref_resize) throw std::runtime_error("shim: no resize hook installed");
  int rc = g_t360ref_resize(src.data, src.rows, src.cols, src.step, dst.data, dst.rows, dst.cols, dst.step,
                            interpolation);
  if (rc != 0) throw std::runtime_error("shim: resize hook failed");
}

}  // namespace cv
