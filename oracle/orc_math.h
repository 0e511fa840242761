// orc_math.h — CPU ORACLE (test infrastructure, never shipped, never on the product path).
//
// Restates the Eigen 3.2 / PCL 1.7.x float arithmetic the reference reaches through
// seg.segment(), ne.compute(). PARITY UNPINNED: PCL/Eigen are not vendored
// in /root/reference and cannot be built here, so the operation orders below are *pinned choices*
// (documented per function); the CUDA path is tested bit-for-bit against them.
//
// Build: g++ -O2 -ffp-contract=off (no FMA contraction, no fast-math): every float op rounds once.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace orc {

// Eigen::Vector4f as used by PCL (w = 0 for directions, 1 for PointXYZ::data[3]).
struct V4 {
  float v[4];
  float& operator[](int i) { return v[i]; }
  float operator[](int i) const { return v[i]; }
};
inline V4 mk(float x, float y, float z, float w) { return V4{{x, y, z, w}}; }
inline V4 operator+(const V4& a, const V4& b) { return mk(a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3]); }
inline V4 operator-(const V4& a, const V4& b) { return mk(a[0] - b[0], a[1] - b[1], a[2] - b[2], a[3] - b[3]); }
inline V4 operator*(float s, const V4& a) { return mk(s * a[0], s * a[1], s * a[2], s * a[3]); }
inline V4 operator/(const V4& a, float s) { return mk(a[0] / s, a[1] / s, a[2] / s, a[3] / s); }

// Vector4f::dot — Eigen 3.2 SSE2 path: packet multiply, then predux =
// _mm_add_ps(a, movehl(a,a)) followed by add_ss with lane 1  ⇒  (p0+p2) + (p1+p3).
// (Eigen/src/Core/arch/SSE/PacketMath.h predux<Packet4f>, non-SSE3 branch.) PINNED CHOICE.
inline float dot4(const V4& a, const V4& b) {
  float p0 = a[0] * b[0], p1 = a[1] * b[1], p2 = a[2] * b[2], p3 = a[3] * b[3];
  return (p0 + p2) + (p1 + p3);
}
inline float sqnorm4(const V4& a) { return dot4(a, a); }
inline float norm4(const V4& a) { return sqrtf(sqnorm4(a)); }
// MatrixBase::normalize(): *this /= norm()  (true division per component, Eigen 3.2)
inline V4 normalized4(const V4& a) { return a / norm4(a); }
// MatrixBase::cross3 (Eigen/src/Geometry/arch/Geometry_SSE.h): two packet products, one subtract.
inline V4 cross3(const V4& a, const V4& b) {
  return mk(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0], 0.0f);
}

// pcl::sqrPointToLineDistance (common/distances.h): float arithmetic, returned as double.
inline double sqrPointToLineDistance(const V4& pt, const V4& line_pt, const V4& line_dir) {
  return (double)(sqnorm4(cross3(line_dir, line_pt - pt)) / sqnorm4(line_dir));
}

// pcl::getAngle3D (common/impl/common.hpp, 1.7): float dot / sqrtf(float product) widened to
// double, clamped, double acos. PINNED CHOICE for the float/double split.
inline double getAngle3D(const V4& v1, const V4& v2) {
  double rad = (double)(dot4(v1, v2) / sqrtf(sqnorm4(v1) * sqnorm4(v2)));
  if (rad < -1.0) rad = -1.0;
  else if (rad > 1.0) rad = 1.0;
  return acos(rad);
}

// Transcendentals that PCL evaluates in float (acosf, sinf, cosf, tanf, atan2f): libm and CUDA
// differ in the last ulp, so the oracle DEFINES them as the double function rounded once to float
// (agrees with glibc's float functions except for ~1e-7 of arguments). PINNED CHOICE.
inline float acosf_d(float x) { return (float)acos((double)x); }
inline float sinf_d(float x) { return (float)sin((double)x); }
inline float cosf_d(float x) { return (float)cos((double)x); }
inline float tanf_d(float x) { return (float)tan((double)x); }
inline float atan2f_d(float y, float x) { return (float)atan2((double)y, (double)x); }

// Error-free double-double accumulator. Used wherever PCL/Eigen sums O(n) float terms: the oracle
// DEFINES those sums as exact, rounded once to float (order independent), because a parallel
// reduction cannot reproduce a sequential float order. PINNED CHOICE (SURVEY.md hard part 4).
struct DD {
  double hi = 0.0, lo = 0.0;
  inline void add(double x) {
    double s = hi + x;
    double bb = s - hi;
    double err = (hi - (s - bb)) + (x - bb);
    hi = s;
    lo += err;
  }
  inline double value() const { return hi + lo; }
  inline float f() const { return (float)(hi + lo); }
};

// pcl::computeRoots2 (common/impl/eigen.hpp)
inline void computeRoots2(float b, float c, float roots[3]) {
  roots[0] = 0.0f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0) d = 0.0f;
  float sd = sqrtf(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

// pcl::computeRoots for a symmetric 3x3 float matrix m (row-major, 9 entries).
inline void computeRoots(const float m[9], float roots[3]) {
  const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
  float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
  float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
  float c2 = m00 + m11 + m22;
  if (fabsf(c0) < std::numeric_limits<float>::epsilon()) {
    computeRoots2(c2, c1, roots);
  } else {
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = sqrtf(3.0f);
    float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    float rho = sqrtf(-a_over_3);
    float theta = atan2f_d(sqrtf(-q), half_b) * s_inv3;
    float cos_theta = cosf_d(theta);
    float sin_theta = sinf_d(theta);
    roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
    roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
    roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
      std::swap(roots[1], roots[2]);
      if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
    }
    if (roots[0] <= 0.0f) computeRoots2(c2, c1, roots);
  }
}

// pcl::eigen33(mat, eigenvalue, eigenvector): smallest eigenpair.
inline void eigen33(const float mat[9], float& eigenvalue, float evec[3]) {
  float scale = 0.0f;
  for (int i = 0; i < 9; ++i) scale = std::max(scale, fabsf(mat[i]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float s[9];
  for (int i = 0; i < 9; ++i) s[i] = mat[i] / scale;
  float roots[3];
  computeRoots(s, roots);
  eigenvalue = roots[0] * scale;
  s[0] -= roots[0];
  s[4] -= roots[0];
  s[8] -= roots[0];
  const float* r0 = s;
  const float* r1 = s + 3;
  const float* r2 = s + 6;
  auto cross = [](const float* a, const float* b, float o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
  };
  float v1[3], v2[3], v3[3];
  cross(r0, r1, v1);
  cross(r0, r2, v2);
  cross(r1, r2, v3);
  // Vector3f::squaredNorm (no packet path for 3 floats): left to right.
  float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float* v;
  float l;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
  else { v = v3; l = l3; }
  float sl = sqrtf(l);
  evec[0] = v[0] / sl;
  evec[1] = v[1] / sl;
  evec[2] = v[2] / sl;
}

// Tail of pcl::computeMeanAndCovarianceMatrix: accu (9 sums, already as float) /= n, then
// cov = E[pp^T] - mu mu^T in float.
inline void covFromAccu(float accu[9], float n, float cov[9], float centroid[4]) {
  for (int i = 0; i < 9; ++i) accu[i] /= n;
  centroid[0] = accu[6];
  centroid[1] = accu[7];
  centroid[2] = accu[8];
  centroid[3] = 0.0f;
  cov[0] = accu[0] - accu[6] * accu[6];
  cov[1] = accu[1] - accu[6] * accu[7];
  cov[2] = accu[2] - accu[6] * accu[8];
  cov[4] = accu[3] - accu[7] * accu[7];
  cov[5] = accu[4] - accu[7] * accu[8];
  cov[8] = accu[5] - accu[8] * accu[8];
  cov[3] = cov[1];
  cov[6] = cov[2];
  cov[7] = cov[5];
}

}  // namespace orc
