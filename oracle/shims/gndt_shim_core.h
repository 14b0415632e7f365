// gndt_shim_core.h — inert stand-ins for the third-party headers the reference includes
// (ROS, PCL, Eigen, Boost, OpenCV), none of which exist in this image.
//
// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  These let the reference's OWN sources
// (src/receiver.cpp + include/*.h, compiled from where they lie under /root/reference)
// build as a ROS-free shared library, so that its control flow, container ordering and
// key strings are authoritative for the oracle.  Only the third-party ARITHMETIC is
// restated here (PCL centroid/scatter, Eigen eigen-solver) — see each function.
#ifndef GNDT_SHIM_CORE_H
#define GNDT_SHIM_CORE_H
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

// ------------------------------------------------------------------ Eigen (subset)
namespace Eigen {
template <int R, int C>
struct Mat {
  float d[R * C];
  Mat() { for (int i = 0; i < R * C; ++i) d[i] = 0.f; }
  static Mat Zero() { return Mat(); }
  static Mat Zero(int, int) { return Mat(); }
  float &operator()(int i, int j) { return d[i * C + j]; }
  const float &operator()(int i, int j) const { return d[i * C + j]; }
  float &operator()(int i) { return d[i]; }
  const float &operator()(int i) const { return d[i]; }
  float &operator[](int i) { return d[i]; }
  const float &operator[](int i) const { return d[i]; }
  float x() const { return d[0]; }
  float y() const { return d[1]; }
  float z() const { return d[2]; }
  bool operator==(const Mat &o) const {
    for (int i = 0; i < R * C; ++i) if (d[i] != o.d[i]) return false;
    return true;
  }
  struct CommaInit {
    Mat *m; int k;
    CommaInit &operator,(float v) { m->d[k++] = v; return *this; }
  };
  CommaInit operator<<(float v) { d[0] = v; return CommaInit{this, 1}; }
  Mat operator+(const Mat &o) const { Mat r; for (int i = 0; i < R * C; ++i) r.d[i] = d[i] + o.d[i]; return r; }
  Mat operator-(const Mat &o) const { Mat r; for (int i = 0; i < R * C; ++i) r.d[i] = d[i] - o.d[i]; return r; }
  Mat operator*(float s) const { Mat r; for (int i = 0; i < R * C; ++i) r.d[i] = d[i] * s; return r; }
  Mat operator/(float s) const { Mat r; for (int i = 0; i < R * C; ++i) r.d[i] = d[i] / s; return r; }
  Mat<C, R> transpose() const { Mat<C, R> r; for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) r.d[j * R + i] = d[i * C + j]; return r; }
  Mat<R, 1> col(int j) const { Mat<R, 1> r; for (int i = 0; i < R; ++i) r.d[i] = d[i * C + j]; return r; }
  float dot(const Mat &o) const { float s = d[0] * o.d[0]; for (int i = 1; i < R * C; ++i) s = s + d[i] * o.d[i]; return s; }
};
template <int R, int C> Mat<R, C> operator*(float s, const Mat<R, C> &m) { return m * s; }
template <int R, int K, int C>
Mat<R, C> operator*(const Mat<R, K> &a, const Mat<K, C> &b) {
  Mat<R, C> r;
  for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s = s + a.d[i * K + k] * b.d[k * C + j];
    r.d[i * C + j] = s;
  }
  return r;
}
typedef Mat<3, 3> Matrix3f;
typedef Mat<3, 1> Vector3f;
typedef Mat<4, 1> Vector4f;

// Stand-in for Eigen::EigenSolver<Matrix3f> as used at map2D.h:111-113 on a symmetric
// scatter matrix: cyclic Jacobi in binary64 (identical to oracle/gndt_oracle.c:jacobi3 so
// the two oracles agree bit-for-bit), results rounded to float.  [3P restated]
template <class M> struct EigenSolver {
  Matrix3f vals, vecs;
  explicit EigenSolver(const Matrix3f &m) {
    double A[3][3], V[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { A[i][j] = m(i, j); V[i][j] = (i == j); }
    A[1][0] = A[0][1]; A[2][0] = A[0][2]; A[2][1] = A[1][2];
    for (int sweep = 0; sweep < 60; ++sweep) {
      double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
      if (off == 0.0) break;
      for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        double app = A[p][p], aqq = A[q][q], apq = A[p][q];
        A[p][p] = app - t * apq; A[q][q] = aqq + t * apq; A[p][q] = A[q][p] = 0.0;
        int r = 3 - p - q;
        double arp = A[r][p], arq = A[r][q];
        A[r][p] = A[p][r] = c * arp - s * arq;
        A[r][q] = A[q][r] = s * arp + c * arq;
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
    }
    for (int i = 0; i < 3; ++i) { vals(i, i) = (float)A[i][i]; for (int j = 0; j < 3; ++j) vecs(i, j) = (float)V[i][j]; }
    for (int i = 0; i < 3; ++i) evd[i] = A[i][i];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) vvd[i][j] = V[i][j];
  }
  double evd[3], vvd[3][3];  // unrounded copies, read by the _ref driver only
  Matrix3f pseudoEigenvalueMatrix() const { return vals; }
  Matrix3f pseudoEigenvectors() const { return vecs; }
};
}  // namespace Eigen

// ------------------------------------------------------------------ boost (subset)
namespace boost { template <class T> using shared_ptr = std::shared_ptr<T>; }

// ------------------------------------------------------------------ PCL (subset)
namespace pcl {
struct PointXYZ {  // 16 bytes like pcl::PointXYZ
  float x, y, z, pad;
  PointXYZ() : x(0), y(0), z(0), pad(1.f) {}
  PointXYZ(float a, float b, float c) : x(a), y(b), z(c), pad(1.f) {}
};
template <class P> struct PointCloud {
  std::vector<P> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  typedef boost::shared_ptr<PointCloud<P>> Ptr;
  typedef boost::shared_ptr<const PointCloud<P>> ConstPtr;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  const P &operator[](size_t i) const { return points[i]; }
};
// pcl::compute3DCentroid, dense branch (PCL common/impl/centroid.hpp) [3P restated]:
// binary32 running sums in point order, divided by the point count.
template <class P> unsigned compute3DCentroid(const PointCloud<P> &cloud, Eigen::Vector4f &c) {
  if (cloud.empty()) return 0;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (size_t i = 0; i < cloud.size(); ++i) { sx = sx + cloud[i].x; sy = sy + cloud[i].y; sz = sz + cloud[i].z; }
  float n = static_cast<float>(cloud.size());
  c[0] = sx / n; c[1] = sy / n; c[2] = sz / n; c[3] = 1.f;
  return static_cast<unsigned>(cloud.size());
}
// pcl::computeCovarianceMatrix(cloud, centroid, Matrix3f&), dense branch [3P restated]:
// the UN-normalised scatter, binary32, point order; upper triangle mirrored.
template <class P> unsigned computeCovarianceMatrix(const PointCloud<P> &cloud, const Eigen::Vector4f &c, Eigen::Matrix3f &m) {
  if (cloud.empty()) return 0;
  float xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0;
  for (size_t i = 0; i < cloud.size(); ++i) {
    float px = cloud[i].x - c[0], py = cloud[i].y - c[1], pz = cloud[i].z - c[2];
    yy = yy + py * py; yz = yz + py * pz; zz = zz + pz * pz;
    float ax = px * px, ay = py * px, az = pz * px;
    xx = xx + ax; xy = xy + ay; xz = xz + az;
  }
  m(0, 0) = xx; m(0, 1) = xy; m(0, 2) = xz; m(1, 1) = yy; m(1, 2) = yz; m(2, 2) = zz;
  m(1, 0) = xy; m(2, 0) = xz; m(2, 1) = yz;
  return static_cast<unsigned>(cloud.size());
}
// pcl::io::savePCDFileASCII stand-in: instead of writing a file, park the cloud where the
// _ref driver can read it (used to capture src/test/genePcd.cpp's output).
namespace io {
inline std::vector<float> &captured_cloud() { static std::vector<float> v; return v; }
template <class P> int savePCDFileASCII(const std::string &, const PointCloud<P> &cloud) {
  std::vector<float> &v = captured_cloud();
  v.resize(cloud.points.size() * 4);
  for (size_t i = 0; i < cloud.points.size(); ++i) {
    v[4 * i] = cloud.points[i].x; v[4 * i + 1] = cloud.points[i].y; v[4 * i + 2] = cloud.points[i].z; v[4 * i + 3] = 0.f;
  }
  return 0;
}
}  // namespace io
struct PCLPointCloud2 { const PointXYZ *data = nullptr; size_t n = 0; };
template <class P> void fromPCLPointCloud2(const PCLPointCloud2 &in, PointCloud<P> &out) {
  out.points.assign(in.data, in.data + in.n);
  out.width = (uint32_t)in.n; out.height = 1;
}
}  // namespace pcl

// ------------------------------------------------------------------ ROS (inert)
namespace sensor_msgs {
struct PointCloud2 {
  const pcl::PointXYZ *data = nullptr; size_t n = 0;
  typedef boost::shared_ptr<const PointCloud2> ConstPtr;
};
}
namespace pcl_conversions {
inline void toPCL(const sensor_msgs::PointCloud2 &m, pcl::PCLPointCloud2 &o) { o.data = m.data; o.n = m.n; }
}
namespace ros {
struct Time { static Time now() { return Time(); } };
struct Duration { Duration() {} explicit Duration(double) {} };
struct Rate { explicit Rate(double) {} void sleep() {} };
struct Publisher {
  int getNumSubscribers() const { return 0; }
  template <class M> void publish(const M &) const {}
};
struct Subscriber {};
struct NodeHandle {
  template <class M> Publisher advertise(const std::string &, int) { return Publisher(); }
  template <class F> Subscriber subscribe(const std::string &, int, F) { return Subscriber(); }
};
namespace param { template <class T> bool get(const std::string &, T &) { return false; } }
inline void init(int &, char **, const std::string &) {}
inline void start() {}
inline void spin() {}
inline void spinOnce() {}
inline void shutdown() {}
inline bool ok() { return false; }
}  // namespace ros
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4, LINE_LIST = 5, CUBE_LIST = 6, SPHERE_LIST = 7, POINTS = 8, ADD = 0, DELETE = 2 };
  struct { std::string frame_id; ros::Time stamp; } header;
  std::string ns; int id = 0; int type = 0; int action = 0;
  struct { struct { double x = 0, y = 0, z = 0; } position; struct { double x = 0, y = 0, z = 0, w = 1; } orientation; } pose;
  struct { double x = 0, y = 0, z = 0; } scale;
  struct { float r = 0, g = 0, b = 0, a = 0; } color;
  ros::Duration lifetime;
  struct Pt { double x = 0, y = 0, z = 0; };
  std::vector<Pt> points;
};
struct MarkerArray { std::vector<Marker> markers; };
}
namespace geometry_msgs { typedef visualization_msgs::Marker::Pt Point; }
#endif
