// ORACLE -- TEST INFRASTRUCTURE ONLY.  Stand-in for the few Eigen types the reference's calibration objective uses
// (Eigen is not in this image), so that src/geometry/geometry.cc and the AccelerometerCalibrator part of
// src/calibration/velocity.cc can be compiled where they lie (oracle/Makefile, target _ref) and run as a
// real-reference pin of the oracle's restatement.  Arithmetic follows Eigen's published formulas coefficient by
// coefficient (quaternion product, _transformVector, toRotationMatrix, fixed-size products evaluated lazily); where
// Eigen's own evaluation order depends on its version (the association of 3-term reductions), results can differ from a
// real Eigen build in the last bits, which is why the pin compares with a 1e-12 relative tolerance rather than bitwise.
// Not part of the product; nothing under pilotguru_b200/ includes it.
#pragma once
#include <cmath>
#include <cstddef>
#include <vector>

namespace Eigen {

enum { AutoAlign = 0, DontAlign = 0x2 };

template <typename S, int R, int C, int O = 0> class Matrix;

template <int O> class Matrix<double, 3, 1, O> {
 public:
  double v[3];
  Matrix() : v{0, 0, 0} {}
  Matrix(double x, double y, double z) : v{x, y, z} {}
  template <int O2> Matrix(const Matrix<double, 3, 1, O2>& o) : v{o.v[0], o.v[1], o.v[2]} {}
  static Matrix Zero() { return Matrix(0, 0, 0); }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double squaredNorm() const { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }
  double norm() const { return std::sqrt(squaredNorm()); }
  template <int O2> Matrix& operator+=(const Matrix<double, 3, 1, O2>& o) {
    for (int i = 0; i < 3; i++) v[i] += o.v[i];
    return *this;
  }
  template <int O2> Matrix& operator-=(const Matrix<double, 3, 1, O2>& o) {
    for (int i = 0; i < 3; i++) v[i] -= o.v[i];
    return *this;
  }
  Matrix& operator/=(double s) {
    for (int i = 0; i < 3; i++) v[i] /= s;
    return *this;
  }
  template <int O2> double dot(const Matrix<double, 3, 1, O2>& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
  double operator()(int i) const { return v[i]; }
  double& operator()(int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double& operator[](int i) { return v[i]; }
  void setZero() { v[0] = v[1] = v[2] = 0; }
};
typedef Matrix<double, 3, 1, 0> Vector3d;

template <int A, int B> Vector3d operator+(const Matrix<double, 3, 1, A>& a, const Matrix<double, 3, 1, B>& b) {
  return Vector3d(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]);
}
template <int A, int B> Vector3d operator-(const Matrix<double, 3, 1, A>& a, const Matrix<double, 3, 1, B>& b) {
  return Vector3d(a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]);
}
template <int A> Vector3d operator*(const Matrix<double, 3, 1, A>& a, double s) { return Vector3d(a.v[0] * s, a.v[1] * s, a.v[2] * s); }
template <int A> Vector3d operator*(double s, const Matrix<double, 3, 1, A>& a) { return Vector3d(s * a.v[0], s * a.v[1], s * a.v[2]); }
template <int A> Vector3d operator/(const Matrix<double, 3, 1, A>& a, double s) { return Vector3d(a.v[0] / s, a.v[1] / s, a.v[2] / s); }

template <int O> class Matrix<double, 3, 3, O> {
 public:
  double m[3][3];
  Matrix() : m{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}} {}
  static Matrix Zero() { return Matrix(); }
  static Matrix Identity() { Matrix r; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1; return r; }
  void fill(double x) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m[i][j] = x; }
  double& operator()(int i, int j) { return m[i][j]; }
  double operator()(int i, int j) const { return m[i][j]; }
  Matrix transpose() const {
    Matrix t;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) t.m[i][j] = m[j][i];
    return t;
  }
  Matrix& operator+=(const Matrix& o) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) m[i][j] += o.m[i][j];
    return *this;
  }
};
typedef Matrix<double, 3, 3, 0> Matrix3d;

inline Matrix3d operator*(const Matrix3d& a, double s) {
  Matrix3d r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] * s;
  return r;
}
inline Matrix3d operator*(double s, const Matrix3d& a) {
  Matrix3d r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = s * a.m[i][j];
  return r;
}
inline Matrix3d operator+(const Matrix3d& a, const Matrix3d& b) {
  Matrix3d r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j];
  return r;
}
inline Matrix3d operator*(const Matrix3d& a, const Matrix3d& b) {
  Matrix3d r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
template <int A> Vector3d operator*(const Matrix3d& a, const Matrix<double, 3, 1, A>& x) {  // lazy coefficient-based product
  Vector3d r;
  for (int i = 0; i < 3; i++) r.v[i] = a.m[i][0] * x.v[0] + a.m[i][1] * x.v[1] + a.m[i][2] * x.v[2];
  return r;
}

template <int O> class Matrix<double, 2, 1, O> {
 public:
  double v[2];
  Matrix() : v{0, 0} {}
  Matrix(double x, double y) : v{x, y} {}
  double operator()(int i) const { return v[i]; }
  double& operator()(int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double& operator[](int i) { return v[i]; }
};
typedef Matrix<double, 2, 1, 0> Vector2d;
inline Vector2d operator-(const Vector2d& a, const Vector2d& b) { return Vector2d(a.v[0] - b.v[0], a.v[1] - b.v[1]); }
template <int O> class Matrix<double, 6, 1, O> {
 public:
  double v[6];
  Matrix() : v{0, 0, 0, 0, 0, 0} {}
  double operator[](int i) const { return v[i]; }
  double& operator[](int i) { return v[i]; }
};
template <int O> class Matrix<double, 2, 6, O> {
 public:
  double m[2][6];
  Matrix() : m{{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}} {}
  double& operator()(int i, int j) { return m[i][j]; }
  double operator()(int i, int j) const { return m[i][j]; }
};

// `v << a, b, c;` (Eigen's comma initialiser) for the fixed vectors Optimizer.cc fills that way
template <typename V> struct CommaInit {
  V* v;
  int i;
  CommaInit& operator,(double x) { (*v)[i++] = x; return *this; }
};
template <int O> CommaInit<Matrix<double, 2, 1, O> > operator<<(Matrix<double, 2, 1, O>& v, double x) { v[0] = x; return CommaInit<Matrix<double, 2, 1, O> >{&v, 1}; }
template <int O> CommaInit<Matrix<double, 3, 1, O> > operator<<(Matrix<double, 3, 1, O>& v, double x) { v[0] = x; return CommaInit<Matrix<double, 3, 1, O> >{&v, 1}; }
template <int O> class Matrix<double, 2, 2, O> {
 public:
  double m[2][2];
  Matrix() : m{{0, 0}, {0, 0}} {}
  static Matrix Identity() { Matrix r; r.m[0][0] = r.m[1][1] = 1; return r; }
  double operator()(int i, int j) const { return m[i][j]; }
};
typedef Matrix<double, 2, 2, 0> Matrix2d;
inline Matrix2d operator*(const Matrix2d& a, double s) { Matrix2d r; for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) r.m[i][j] = a.m[i][j] * s; return r; }

template <typename S, int O = 0> class Quaternion;
template <int O> class Quaternion<double, O> {
 public:
  double qx, qy, qz, qw;
  Quaternion() : qx(0), qy(0), qz(0), qw(1) {}
  Quaternion(double w, double x, double y, double z) : qx(x), qy(y), qz(z), qw(w) {}
  template <int O2> Quaternion(const Quaternion<double, O2>& o) : qx(o.qx), qy(o.qy), qz(o.qz), qw(o.qw) {}
  explicit Quaternion(const Matrix<double, 3, 3, 0>& mat) {   // Eigen quaternionbase_assign_impl<Matrix3>: trace-based conversion
    const double (*m)[3] = mat.m;
    double t = m[0][0] + m[1][1] + m[2][2];
    if (t > 0) {
      t = std::sqrt(t + 1.0);
      qw = 0.5 * t;
      t = 0.5 / t;
      qx = (m[2][1] - m[1][2]) * t;
      qy = (m[0][2] - m[2][0]) * t;
      qz = (m[1][0] - m[0][1]) * t;
    } else {
      int i = 0;
      if (m[1][1] > m[0][0]) i = 1;
      if (m[2][2] > m[i][i]) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0);
      double q[3];
      q[i] = 0.5 * t;
      t = 0.5 / t;
      qw = (m[k][j] - m[j][k]) * t;
      q[j] = (m[j][i] + m[i][j]) * t;
      q[k] = (m[k][i] + m[i][k]) * t;
      qx = q[0]; qy = q[1]; qz = q[2];
    }
  }
  void setIdentity() { qx = qy = qz = 0; qw = 1; }
  double squaredNorm() const { return qx * qx + qy * qy + qz * qz + qw * qw; }
  double norm() const { return std::sqrt(squaredNorm()); }
  void normalize() { const double n = norm(); qx /= n; qy /= n; qz /= n; qw /= n; }
  struct Coeffs {   // coeffs() in Eigen's storage order x, y, z, w
    Quaternion* q;
    Coeffs& operator*=(double s) { q->qx *= s; q->qy *= s; q->qz *= s; q->qw *= s; return *this; }
  };
  Coeffs coeffs() { return Coeffs{this}; }
  template <int O2> Quaternion& operator*=(const Quaternion<double, O2>& b) { *this = Quaternion(*this * b); return *this; }
  template <int A> Matrix<double, 3, 1, 0> operator*(const Matrix<double, 3, 1, A>& v) const { return _transformVector(v); }
  Quaternion<double, 0> conjugate() const { return Quaternion<double, 0>(qw, -qx, -qy, -qz); }
  double w() const { return qw; }
  double x() const { return qx; }
  double y() const { return qy; }
  double z() const { return qz; }
  double& w() { return qw; }
  double& x() { return qx; }
  double& y() { return qy; }
  double& z() { return qz; }
  template <int O2> Quaternion<double, 0> operator*(const Quaternion<double, O2>& b) const {  // Eigen quat_product
    return Quaternion<double, 0>(qw * b.qw - qx * b.qx - qy * b.qy - qz * b.qz, qw * b.qx + qx * b.qw + qy * b.qz - qz * b.qy,
                                 qw * b.qy + qy * b.qw + qz * b.qx - qx * b.qz, qw * b.qz + qz * b.qw + qx * b.qy - qy * b.qx);
  }
  template <int A> Vector3d _transformVector(const Matrix<double, 3, 1, A>& v) const {  // Eigen QuaternionBase::_transformVector
    Vector3d uv(qy * v.v[2] - qz * v.v[1], qz * v.v[0] - qx * v.v[2], qx * v.v[1] - qy * v.v[0]);
    uv += uv;
    const Vector3d c(qy * uv.v[2] - qz * uv.v[1], qz * uv.v[0] - qx * uv.v[2], qx * uv.v[1] - qy * uv.v[0]);
    return Vector3d(v.v[0] + qw * uv.v[0] + c.v[0], v.v[1] + qw * uv.v[1] + c.v[1], v.v[2] + qw * uv.v[2] + c.v[2]);
  }
  Matrix3d toRotationMatrix() const {  // Eigen QuaternionBase::toRotationMatrix
    Matrix3d r;
    const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
    const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    r.m[0][0] = 1 - (tyy + tzz); r.m[0][1] = txy - twz; r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz; r.m[1][1] = 1 - (txx + tzz); r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy; r.m[2][1] = tyz + twx; r.m[2][2] = 1 - (txx + tyy);
    return r;
  }
};
typedef Quaternion<double, 0> Quaterniond;

// ---- dynamic vectors / matrices / maps: what thirdparty/LBFGS (LBFGS.h, LineSearch.h) and the functor signature use.
// Every operation is evaluated eagerly, coefficient by coefficient.  dot() / squaredNorm() sum LEFT TO RIGHT; the real
// Eigen's vectorised reductions associate differently (packet width and version dependent), so a real build can differ
// from this in the last bits -- the pin therefore checks the driver's logic, not Eigen's rounding.
enum { Dynamic = -1 };

template <typename V> class Map;

template <int O> class Matrix<double, Dynamic, 1, O> {
 public:
  std::vector<double> d;
  Matrix() {}
  explicit Matrix(long n) : d((size_t)n, 0.0) {}
  static Matrix Zero(long n) { return Matrix(n); }
  long size() const { return (long)d.size(); }
  void resize(long n) { d.resize((size_t)n); }
  double& operator[](long i) { return d[(size_t)i]; }
  const double& operator[](long i) const { return d[(size_t)i]; }
  double* data() { return d.data(); }
  const double* data() const { return d.data(); }
  double squaredNorm() const { double s = 0; for (double v : d) s += v * v; return s; }
  double norm() const { return std::sqrt(squaredNorm()); }
  double dot(const Matrix& o) const { double s = 0; for (size_t i = 0; i < d.size(); i++) s += d[i] * o.d[i]; return s; }
  Matrix& noalias() { return *this; }
  Matrix operator-() const { Matrix r(size()); for (size_t i = 0; i < d.size(); i++) r.d[i] = -d[i]; return r; }
  Matrix& operator+=(const Matrix& o) { for (size_t i = 0; i < d.size(); i++) d[i] += o.d[i]; return *this; }
  Matrix& operator-=(const Matrix& o) { for (size_t i = 0; i < d.size(); i++) d[i] -= o.d[i]; return *this; }
  Matrix& operator*=(double s) { for (double& v : d) v *= s; return *this; }
};
typedef Matrix<double, Dynamic, 1, 0> VectorXd;

inline VectorXd operator+(const VectorXd& a, const VectorXd& b) { VectorXd r(a.size()); for (long i = 0; i < a.size(); i++) r[i] = a[i] + b[i]; return r; }
inline VectorXd operator-(const VectorXd& a, const VectorXd& b) { VectorXd r(a.size()); for (long i = 0; i < a.size(); i++) r[i] = a[i] - b[i]; return r; }
inline VectorXd operator*(double s, const VectorXd& a) { VectorXd r(a.size()); for (long i = 0; i < a.size(); i++) r[i] = s * a[i]; return r; }
inline VectorXd operator*(const VectorXd& a, double s) { VectorXd r(a.size()); for (long i = 0; i < a.size(); i++) r[i] = a[i] * s; return r; }

template <int O> class Matrix<double, Dynamic, Dynamic, O> {  // column-major, like Eigen's default
 public:
  std::vector<double> d;
  long rows_ = 0, cols_ = 0;
  void resize(long r, long c) { rows_ = r; cols_ = c; d.assign((size_t)(r * c), 0.0); }
  double& operator()(long i, long j) { return d[(size_t)(j * rows_ + i)]; }
  const double& operator()(long i, long j) const { return d[(size_t)(j * rows_ + i)]; }
};
typedef Matrix<double, Dynamic, Dynamic, 0> MatrixXd;

template <> class Map<VectorXd> {
 public:
  double* p;
  long n;
  Map(double* ptr, long size) : p(ptr), n(size) {}
  long size() const { return n; }
  double squaredNorm() const { double s = 0; for (long i = 0; i < n; i++) s += p[i] * p[i]; return s; }
  double dot(const Map& o) const { double s = 0; for (long i = 0; i < n; i++) s += p[i] * o.p[i]; return s; }
  double dot(const VectorXd& o) const { double s = 0; for (long i = 0; i < n; i++) s += p[i] * o[i]; return s; }
  Map& noalias() { return *this; }
  Map& operator=(const VectorXd& v) { for (long i = 0; i < n; i++) p[i] = v[i]; return *this; }
};
inline VectorXd operator*(double s, const Map<VectorXd>& a) { VectorXd r(a.size()); for (long i = 0; i < a.size(); i++) r[i] = s * a.p[i]; return r; }

}  // namespace Eigen
