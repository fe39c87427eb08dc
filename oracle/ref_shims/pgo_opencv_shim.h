// ORACLE -- TEST INFRASTRUCTURE ONLY.  Stand-in for the part of OpenCV that thirdparty/orb-slam2/src/ORBextractor.cc uses
// (OpenCV's C++ headers are not in this image), so that the reference's extractor can be compiled where it lies
// (oracle/Makefile, target _ref) and run as a real-reference pin of the oracle's restatement of that FILE: the pyramid
// orchestration, the cell grid with its threshold retry, ExtractorNode / DistributeOctTree, IC_Angle, computeOrbDescriptor,
// the scale tables and quotas, the level-major output order and the keypoint rescale all execute from the reference's
// own source.  What lives inside OpenCV itself -- cv::resize(INTER_LINEAR), cv::FAST(TYPE_9_16), cv::GaussianBlur(7x7, 2),
// cv::fastAtan2, cvRound -- is forwarded to the oracle's restatements of those primitives (oracle/pgo_orb.cc), which are
// pinned bit-exact against cv2 4.13 by tests/test_oracle_orb.py.  cv::Mat here is single-channel 8-bit (images,
// descriptors) or 32-bit float (the small pose / point matrices ORBmatcher.cc multiplies), with
// OpenCV's view semantics (rowRange / colRange / operator()(Rect) share the buffer; create() keeps a buffer of the right
// size, which is what lets resize() and copyMakeBorder() write through the pyramid's views).
// Not part of the product; nothing under pilotguru_b200/ includes it.
#pragma once
#include <algorithm>   // the real headers pull these in; ORBextractor.cc relies on it (sort, back_inserter, list, pair)
#include <iterator>
#include <list>
#include <utility>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

extern "C" {
void pgo_pca3(const double* rows, int64_t n, double* eigvec9, double* eigval3, double* mean3);
void pgo_resize_linear(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
int pgo_fast(const uint8_t* img, int w, int h, int th, int nms, int32_t* xys, int cap);
void pgo_gaussian_blur7(const uint8_t* src, int w, int h, uint8_t* dst);
float pgo_fast_atan2(float y, float x);
}

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
#define CV_64FC1 6
#define CV_PCA_DATA_AS_ROW 0
#define CV_PI 3.1415926535897932384626433832795

typedef unsigned char uchar;

inline int cvRound(double v) { return (int)lrint(v); }   // round half to even, like OpenCV's SSE2 / lrint path
inline int cvFloor(double v) { const int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { const int i = (int)v; return i + (i < v); }

namespace cv {

using ::uchar;
enum { BORDER_REPLICATE = 1, BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_LINEAR = 1 };

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

struct Vec3d;

class Mat {
 public:
  int rows, cols;
  size_t step;   // bytes per row
  uchar* data;
  int type_;
  std::shared_ptr<std::vector<uchar> > buf;

  Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8UC1) {}
  Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr), type_(type) { create(r, c, type); }
  Mat(Size sz, int type) : rows(0), cols(0), step(0), data(nullptr), type_(type) { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, void* ext, size_t st) : rows(r), cols(c), step(st), data((uchar*)ext), type_(type) {}
  explicit Mat(const Vec3d& v);   // 3 x 1 CV_64F, like cv::Mat(const Vec<T, n>&)
  // Mat::zeros returns a MatExpr in OpenCV, and assigning a MatExpr evaluates it INTO the destination: create() keeps a
  // destination of the right size, so `descriptors = Mat::zeros(n, 32, CV_8UC1)` inside computeDescriptors() clears the
  // rows of the caller's output matrix that `descriptors` views (ORBextractor.cc:1036, :1088) instead of rebinding it.
  struct ZerosExpr { int r, c, type; };
  static ZerosExpr zeros(int r, int c, int type) { return ZerosExpr{r, c, type}; }
  Mat(const ZerosExpr& e) : rows(0), cols(0), step(0), data(nullptr), type_(e.type) { *this = e; }
  Mat& operator=(const ZerosExpr& e) {
    create(e.r, e.c, e.type);
    for (int y = 0; y < rows; y++) memset(data + (size_t)y * step, 0, (size_t)cols * elemSize());
    return *this;
  }

  size_t elemSize() const { return type_ == CV_32F ? 4 : type_ == CV_64F ? 8 : 1; }
  void create(int r, int c, int type) {
    if (data && rows == r && cols == c && type_ == type) return;   // cv::Mat::create keeps a matrix of the right size (views included)
    type_ = type;
    buf = std::make_shared<std::vector<uchar> >((size_t)r * c * elemSize() + 16);
    rows = r; cols = c; step = (size_t)c * elemSize(); data = buf->data();
  }
  void release() { buf.reset(); rows = cols = 0; step = 0; data = nullptr; }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  Size size() const { return Size(cols, rows); }
  size_t step1() const { return step / elemSize(); }
  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
  template <typename T> const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
  Mat view(int x, int y, int w, int h) const {
    Mat m;
    m.rows = h; m.cols = w; m.step = step; m.type_ = type_; m.data = data + (size_t)y * step + (size_t)x * elemSize(); m.buf = buf;
    return m;
  }
  Mat rowRange(int a, int b) const { return view(0, a, cols, b - a); }
  Mat colRange(int a, int b) const { return view(a, 0, b - a, rows); }
  Mat row(int y) const { return view(0, y, cols, 1); }
  Mat col(int x) const { return view(x, 0, 1, rows); }
  Mat operator()(const Rect& r) const { return view(r.x, r.y, r.width, r.height); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * elemSize());
    return m;
  }
  std::vector<uchar> tight() const {  // helper of the stand-in: an 8-bit matrix as a dense rows x cols buffer
    std::vector<uchar> t((size_t)rows * cols + 1);
    for (int y = 0; y < rows; y++) memcpy(t.data() + (size_t)y * cols, data + (size_t)y * step, (size_t)cols);
    return t;
  }
  // ---- the float algebra ORBmatcher.cc writes on poses and points.  OpenCV evaluates these MatExpr through gemm();
  // for the 3x3 / 3x1 operands used here its small-matrix path forms each dot product in float, left to right, then
  // adds the addend (alpha = beta = 1).  Only the monocular branch consumes the values (see ref_wrap_match.cc).
  double getd(int y, int x) const { return type_ == CV_64F ? at<double>(y, x) : (double)at<float>(y, x); }
  void setd(int y, int x, double v) { if (type_ == CV_64F) at<double>(y, x) = v; else at<float>(y, x) = (float)v; }
  Mat t() const {
    Mat m(cols, rows, type_);
    for (int y = 0; y < rows; y++)
      for (int x = 0; x < cols; x++) m.setd(x, y, getd(y, x));
    return m;
  }
  Mat operator-() const {
    Mat m(rows, cols, type_);
    for (int y = 0; y < rows; y++)
      for (int x = 0; x < cols; x++) m.setd(y, x, -getd(y, x));
    return m;
  }
};
inline Mat operator*(const Mat& a, const Mat& b) {   // dot products left to right, in the matrices' own precision
  Mat m(a.rows, b.cols, a.type());
  for (int y = 0; y < a.rows; y++)
    for (int x = 0; x < b.cols; x++) {
      if (a.type() == CV_64F) {
        double s = 0;
        for (int k = 0; k < a.cols; k++) s = k == 0 ? a.at<double>(y, 0) * b.at<double>(0, x) : s + a.at<double>(y, k) * b.at<double>(k, x);
        m.at<double>(y, x) = s;
      } else {
        float s = 0.f;
        for (int k = 0; k < a.cols; k++) s = k == 0 ? a.at<float>(y, 0) * b.at<float>(0, x) : s + a.at<float>(y, k) * b.at<float>(k, x);
        m.at<float>(y, x) = s;
      }
    }
  return m;
}
inline Mat operator+(const Mat& a, const Mat& b) {
  Mat m(a.rows, a.cols, a.type());
  for (int y = 0; y < a.rows; y++)
    for (int x = 0; x < a.cols; x++) {
      if (a.type() == CV_64F) m.at<double>(y, x) = a.at<double>(y, x) + b.at<double>(y, x);
      else m.at<float>(y, x) = a.at<float>(y, x) + b.at<float>(y, x);
    }
  return m;
}

class _InputArray {
 public:
  const Mat* m;
  _InputArray() : m(nullptr) {}
  _InputArray(const Mat& mm) : m(&mm) {}
  bool empty() const { return !m || m->empty(); }
  Mat getMat() const { return m ? *m : Mat(); }
};
typedef const _InputArray& InputArray;
class _OutputArray {
 public:
  Mat* m;
  _OutputArray(Mat& mm) : m(&mm) {}
  void create(int r, int c, int type) const { m->create(r, c, type); }
  Mat getMat() const { return *m; }
  void release() const { m->release(); }
};
typedef const _OutputArray& OutputArray;
inline _InputArray noArray() { return _InputArray(); }

inline float fastAtan2(float y, float x) { return pgo_fast_atan2(y, x); }

inline void resize(const Mat& src, Mat& dst, Size dsize, double /*fx*/, double /*fy*/, int /*interpolation*/) {
  dst.create(dsize.height, dsize.width, CV_8UC1);
  const std::vector<uchar> s = src.tight();
  std::vector<uchar> d((size_t)dsize.width * dsize.height + 1);
  pgo_resize_linear(s.data(), src.cols, src.rows, d.data(), dsize.width, dsize.height);
  for (int y = 0; y < dst.rows; y++) memcpy(dst.data + (size_t)y * dst.step, d.data() + (size_t)y * dst.cols, (size_t)dst.cols);
}

inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int /*borderType*/) {
  const Mat s = src.clone();   // the ISOLATED call passes a view of dst itself
  dst.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
  for (int y = 0; y < dst.rows; y++) {
    const uchar* srow = s.data + (size_t)reflect101(y - top, s.rows) * s.step;
    uchar* drow = dst.data + (size_t)y * dst.step;
    for (int x = 0; x < dst.cols; x++) drow[x] = srow[reflect101(x - left, s.cols)];
  }
}

inline void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
  keypoints.clear();
  if (image.rows < 7 || image.cols < 7) return;
  // scratch is kept between calls: the library's allocator never frees (ref_bump_alloc.cc) and this runs once per cell
  static thread_local std::vector<uchar> s;
  static thread_local std::vector<int32_t> xys;
  const size_t px = (size_t)image.rows * image.cols;
  if (s.size() < px + 1) s.resize(px + 1);
  if (xys.size() < px * 3 + 3) xys.resize(px * 3 + 3);
  for (int y = 0; y < image.rows; y++) memcpy(s.data() + (size_t)y * image.cols, image.data + (size_t)y * image.step, (size_t)image.cols);
  const int n = pgo_fast(s.data(), image.cols, image.rows, threshold, nonmaxSuppression ? 1 : 0, xys.data(), (int)px);
  assert(n >= 0);
  for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1, (float)xys[3 * i + 2]));
}

inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sx, double sy, int /*borderType*/) {
  assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2);
  (void)ksize; (void)sx; (void)sy;
  const std::vector<uchar> s = src.tight();
  std::vector<uchar> d((size_t)src.rows * src.cols + 1);
  pgo_gaussian_blur7(s.data(), src.cols, src.rows, d.data());
  dst.create(src.rows, src.cols, CV_8UC1);
  for (int y = 0; y < dst.rows; y++) memcpy(dst.data + (size_t)y * dst.step, d.data() + (size_t)y * dst.cols, (size_t)dst.cols);
}

// ---- what src/calibration/rotation.cc uses: Vec3d, norm, and cv::PCA over an n x 3 CV_64F matrix with the samples as
// rows.  The PCA itself (covariance + OpenCV's cyclic Jacobi sweep, eigenvectors as rows, descending eigenvalues) is the
// oracle's restatement, pinned against cv2.PCACompute2 including the eigenvector signs (tests/golden/cv2_pca.npz).
struct Vec3d {
  double val[3];
  Vec3d() : val{0, 0, 0} {}
  Vec3d(double a, double b, double c) : val{a, b, c} {}
  double& operator[](int i) { return val[i]; }
  const double& operator[](int i) const { return val[i]; }
  double dot(const Vec3d& o) const { return val[0] * o.val[0] + val[1] * o.val[1] + val[2] * o.val[2]; }
  double operator()(int i) const { return val[i]; }
  Vec3d cross(const Vec3d& o) const {
    return Vec3d(val[1] * o.val[2] - val[2] * o.val[1], val[2] * o.val[0] - val[0] * o.val[2], val[0] * o.val[1] - val[1] * o.val[0]);
  }
};
inline Mat::Mat(const Vec3d& v) : rows(0), cols(0), step(0), data(nullptr), type_(CV_64F) {
  create(3, 1, CV_64F);
  for (int i = 0; i < 3; i++) at<double>(i, 0) = v.val[i];
}
inline Vec3d operator*(const Vec3d& a, double s) { return Vec3d(a.val[0] * s, a.val[1] * s, a.val[2] * s); }
inline Vec3d operator/(const Vec3d& a, double s) { return Vec3d(a.val[0] / s, a.val[1] / s, a.val[2] / s); }
enum { NORM_L2 = 4 };
inline double norm(const Vec3d& v, int /*normType*/) { return std::sqrt(v.val[0] * v.val[0] + v.val[1] * v.val[1] + v.val[2] * v.val[2]); }

inline double norm(const Mat& m, int /*normType*/) {   // NORM_L2 of a CV_64F matrix
  double s = 0;
  for (int y = 0; y < m.rows; y++)
    for (int x = 0; x < m.cols; x++) s += m.at<double>(y, x) * m.at<double>(y, x);
  return std::sqrt(s);
}
// cv::getGaussianKernel (imgproc/smooth.cpp, OpenCV 2.4): the sigma > 0 branch; CV_64F, n x 1.
inline Mat getGaussianKernel(int n, double sigma, int ktype = CV_64F) {
  assert(sigma > 0 && ktype == CV_64F);
  Mat kernel(n, 1, CV_64F);
  const double scale2X = -0.5 / (sigma * sigma);
  double sum = 0;
  for (int i = 0; i < n; i++) {
    const double x = i - (n - 1) * 0.5;
    const double t = std::exp(scale2X * x * x);
    kernel.at<double>(i, 0) = t;
    sum += t;
  }
  sum = 1. / sum;
  for (int i = 0; i < n; i++) kernel.at<double>(i, 0) *= sum;
  return kernel;
}
// cv::sepFilter2D for CV_64F with a 1-tap kernelY (what SmoothHeadingDirections asks for): the generic RowFilter sums
// kx[k] * src[x + k - anchor] for k = 0 .. ksize-1 with replicated borders; the column pass multiplies by kernelY[0].
inline void sepFilter2D(const Mat& src, Mat& dst, int /*ddepth*/, const Mat& kernelX, const Mat& kernelY, Point /*anchor*/, double delta,
                        int borderType) {
  assert(src.type() == CV_64F && kernelY.rows * kernelY.cols == 1 && borderType == BORDER_REPLICATE);
  (void)borderType;
  const int ks = kernelX.rows * kernelX.cols, anchor = ks / 2;
  Mat out(src.rows, src.cols, CV_64F);
  for (int y = 0; y < src.rows; y++)
    for (int x = 0; x < src.cols; x++) {
      double s = 0;
      for (int k = 0; k < ks; k++) {
        int j = x + k - anchor;
        j = j < 0 ? 0 : (j >= src.cols ? src.cols - 1 : j);
        const double term = kernelX.at<double>(k) * src.at<double>(y, j);
        s = k == 0 ? term : s + term;
      }
      out.at<double>(y, x) = kernelY.at<double>(0) * s + delta;
    }
  dst.create(src.rows, src.cols, CV_64F);
  for (int y = 0; y < src.rows; y++) memcpy(dst.data + (size_t)y * dst.step, out.data + (size_t)y * out.step, (size_t)src.cols * 8);
}

class PCA {
 public:
  Mat eigenvectors, eigenvalues, mean;
  PCA(const Mat& data, const _InputArray& /*mean*/, int /*flags*/) {
    assert(data.type() == CV_64F && data.cols == 3);
    std::vector<double> rows((size_t)data.rows * 3);
    for (int y = 0; y < data.rows; y++)
      for (int x = 0; x < 3; x++) rows[(size_t)y * 3 + x] = data.at<double>(y, x);
    double ev[9], ew[3], mu[3];
    pgo_pca3(rows.data(), data.rows, ev, ew, mu);
    eigenvectors = Mat(3, 3, CV_64F); eigenvalues = Mat(3, 1, CV_64F); mean = Mat(1, 3, CV_64F);
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) eigenvectors.at<double>(i, j) = ev[3 * i + j];
      eigenvalues.at<double>(i, 0) = ew[i];
      mean.at<double>(0, i) = mu[i];
    }
  }
};

struct KeyPointsFilter {  // only named by ComputeKeyPointsOld, which operator() does not call
  static void retainBest(std::vector<KeyPoint>& keypoints, int npoints) {
    if (npoints >= 0 && (size_t)npoints < keypoints.size()) keypoints.resize((size_t)npoints);
  }
};

}  // namespace cv
