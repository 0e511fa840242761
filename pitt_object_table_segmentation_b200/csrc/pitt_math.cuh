// pitt_math.cuh — device float math of the hot path (sm_100a).
//
// The whole library is compiled with -fmad=false: a*b+c is NEVER contracted into an FMA unless the
// code says __fmaf_rn explicitly, division and sqrt are IEEE (nvcc defaults -prec-div/-prec-sqrt),
// denormals are kept. That makes every float expression below round exactly like the x86 SSE
// scalar code the reference's PCL runs, so inlier predicates are bit-exact.
//
// Operation orders follow Eigen 3.2 / PCL 1.7.x as stated in SURVEY.md Appendix B:
//   Vector4f::dot  = (p0 + p2) + (p1 + p3)   (SSE2 predux)
//   cross3         = two products, one subtract per component
//   normalize()    = component-wise true division by sqrtf(squaredNorm)
// Transcendentals that PCL evaluates in float are evaluated in double and rounded once
// (acosf/sinf/cosf/tanf/atan2f differ by an ulp between libm and CUDA otherwise).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace pitt {

struct f3 {
  float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
// Vector4f (w = 0) dot in Eigen's SSE2 reduction order
__device__ __forceinline__ float dot0(f3 a, f3 b) { return (a.x * b.x + a.z * b.z) + a.y * b.y; }
__device__ __forceinline__ float sqn0(f3 a) { return dot0(a, a); }
__device__ __forceinline__ float nrm0(f3 a) { return sqrtf(sqn0(a)); }
__device__ __forceinline__ f3 unit0(f3 a) { return a / nrm0(a); }
__device__ __forceinline__ f3 cross0(f3 a, f3 b) {
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// pcl::sqrPointToLineDistance in float (the caller widens)
__device__ __forceinline__ float sqr_pt_line(f3 pt, f3 lp, f3 ld) { return sqn0(cross0(ld, lp - pt)) / sqn0(ld); }
// pcl::getAngle3D: float quotient, clamped, double acos
__device__ __forceinline__ double angle3d(f3 a, f3 b) {
  double rad = (double)(dot0(a, b) / sqrtf(sqn0(a) * sqn0(b)));
  if (rad < -1.0) rad = -1.0;
  else if (rad > 1.0) rad = 1.0;
  return acos(rad);
}
__device__ __forceinline__ float acosf_d(float x) { return (float)acos((double)x); }
__device__ __forceinline__ float sinf_d(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cosf_d(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float tanf_d(float x) { return (float)tan((double)x); }
__device__ __forceinline__ float atan2f_d(float y, float x) { return (float)atan2((double)y, (double)x); }

// ---- pcl::computeRoots2 / computeRoots / eigen33 (common/impl/eigen.hpp), float
__device__ inline void compute_roots2(float b, float c, float* roots) {
  roots[0] = 0.0f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  float sd = sqrtf(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}
__device__ inline void swapf(float& a, float& b) {
  float t = a;
  a = b;
  b = t;
}
__device__ inline void compute_roots(const float* m, float* roots) {
  const float m00 = m[0], m01 = m[1], m02 = m[2], m11 = m[4], m12 = m[5], m22 = m[8];
  float c0 = m00 * m11 * m22 + 2.0f * m01 * m02 * m12 - m00 * m12 * m12 - m11 * m02 * m02 - m22 * m01 * m01;
  float c1 = m00 * m11 - m01 * m01 + m00 * m22 - m02 * m02 + m11 * m22 - m12 * m12;
  float c2 = m00 + m11 + m22;
  if (fabsf(c0) < 1.1920928955078125e-07f) {
    compute_roots2(c2, c1, roots);
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
    if (roots[0] >= roots[1]) swapf(roots[0], roots[1]);
    if (roots[1] >= roots[2]) {
      swapf(roots[1], roots[2]);
      if (roots[0] >= roots[1]) swapf(roots[0], roots[1]);
    }
    if (roots[0] <= 0.0f) compute_roots2(c2, c1, roots);
  }
}
__device__ inline void eigen33(const float* mat, float& eigenvalue, float* evec) {
  float scale = 0.0f;
  for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(mat[i]));
  if (scale <= 1.17549435e-38f) scale = 1.0f;
  float s[9];
  for (int i = 0; i < 9; ++i) s[i] = mat[i] / scale;
  float roots[3];
  compute_roots(s, roots);
  eigenvalue = roots[0] * scale;
  s[0] -= roots[0];
  s[4] -= roots[0];
  s[8] -= roots[0];
  float v1[3], v2[3], v3[3];
  v1[0] = s[1] * s[5] - s[2] * s[4]; v1[1] = s[2] * s[3] - s[0] * s[5]; v1[2] = s[0] * s[4] - s[1] * s[3];
  v2[0] = s[1] * s[8] - s[2] * s[7]; v2[1] = s[2] * s[6] - s[0] * s[8]; v2[2] = s[0] * s[7] - s[1] * s[6];
  v3[0] = s[4] * s[8] - s[5] * s[7]; v3[1] = s[5] * s[6] - s[3] * s[8]; v3[2] = s[3] * s[7] - s[4] * s[6];
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
// tail of pcl::computeMeanAndCovarianceMatrix
__device__ inline void cov_from_accu(float* accu, float n, float* cov, float* centroid) {
  for (int i = 0; i < 9; ++i) accu[i] /= n;
  centroid[0] = accu[6]; centroid[1] = accu[7]; centroid[2] = accu[8];
  cov[0] = accu[0] - accu[6] * accu[6];
  cov[1] = accu[1] - accu[6] * accu[7];
  cov[2] = accu[2] - accu[6] * accu[8];
  cov[4] = accu[3] - accu[7] * accu[7];
  cov[5] = accu[4] - accu[7] * accu[8];
  cov[8] = accu[5] - accu[8] * accu[8];
  cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
}

}  // namespace pitt
