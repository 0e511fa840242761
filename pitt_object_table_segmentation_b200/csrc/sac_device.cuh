// sac_device.cuh — device side of the sample-consensus models (plane, sphere, cylinder, cone):
// closed-form model estimation (K2) and the per-point inlier predicate (K3/K5).
//
// Replaces pcl::SampleConsensusModel{Plane,Sphere,Cylinder,Cone}::{computeModelCoefficients,
// isModelValid,countWithinDistance,selectWithinDistance} reached from seg.segment() at
// supports_segmentation_srv.cpp:110, plane…:67, sphere…:73, cylinder…:126, cone…:127
// (SURVEY.md B.3-B.6). Operation order is the oracle's (oracle/orc_sac.h), written independently.
#pragma once
#include <float.h>

#include "pitt_math.cuh"

namespace pitt {

// One scored hypothesis, 64 B, laid out for LDS.128 broadcasts.
//  plane    v[0..3]  = a,b,c,d
//  sphere   v[0..3]  = cx,cy,cz,r, v[4..5] = [D_lo, D_hi]: the squared distances that are inliers (exact)
//  cylinder v[0..2]  = point on axis, v[4..6] = axis dir, v[3] = r, v[7] = pt.dir, v[8] = 1/dir.dir,
//           v[9..10] = (lo2, hi2): squared axis distances outside (lo2, hi2) are certain outliers,
//           v[11..14] = the same decisions as bounds on N = |dir x (p0 - pt)|^2, i.e. before the division by dir.dir
//           (exact: fl(N / D) is monotone in N): N < v[11] or N >= v[12] certain outlier, v[13] <= N < v[14] certain inlier
//  cone     v[0..2]  = apex, v[4..6] = axis dir, v[3] = opening angle, v[7] = apex.dir,
//           v[8] = 1/dir.dir, v[9] = sin(angle), v[10] = cos(angle), v[11] = float tan(angle),
//           v[12..13] = tan(angle) as double, v[14] = T: |axis distance - cone radius| >= T is a certain outlier,
//           v[15] = D_in: |axis distance - cone radius| <= D_in is a certain inlier (-1: never)
// An invalid hypothesis is all NaN: every comparison is false, so it scores 0.
struct __align__(16) HypRec {
  float v[16];
};

// model limits after SACSegmentation(FromNormals)::initSACModel forwarding
struct Limits {
  double radius_min, radius_max;
  double min_angle, max_angle;
  double eps_angle;
  double w;  // normal_distance_weight of the model
  float ax, ay, az;
};

// One problem of a batched fit (the primitive fits of a frame: problem = cluster, one batch per model). The kernels of sac.cu /
// lm.cu / services.cu take a nullable descriptor array; with it, block z (or y, or x: the first free grid dimension) works on
// problem blockIdx.{z,y,x} and reads its pointers and sizes here instead of from the kernel arguments.
struct FitDesc {
  const float4* xyz;
  const float4* nrm;
  const int* samples;  // [H][S]
  HypRec* recs;        // [H]; recs[0] is reused as the winner's record after the scan
  float* coeffs8;      // [H][8]
  unsigned char* flags;  // [H]
  int* counts;         // [H]
  int* ints;           // 16 ints, see sac_segment_async
  float* flt;          // model[8], refined[8], axis extent[16]
  int* inl;            // [n]
  double* partial;     // plane refinement partial sums
  float4* proj;        // axis extent: projections
  void* bb;            // axis extent: per-block best pairs
  int n, H;
};

struct ScoreParams {
  double thr;    // distance threshold (double, as PCL compares)
  float thr_up;  // smallest float >= thr:  (double)f < thr  <=>  f < thr_up
  double w;      // normal distance weight
  float band;    // FP32 fast path: |score - thr| <= band is re-evaluated exactly
};

__device__ __forceinline__ f3 ld3(const float4* p, int i) {
  float4 v = __ldg(p + i);
  return mk3(v.x, v.y, v.z);
}

// ------------------------------------------------------------------ computeModelCoefficients
__device__ inline bool estimate_plane(const float4* xyz, const int* s, float* mc) {
  f3 p0 = ld3(xyz, s[0]), p1 = ld3(xyz, s[1]), p2 = ld3(xyz, s[2]);
  f3 d1 = p1 - p0, d2 = p2 - p0;
  float qx = d1.x / d2.x, qy = d1.y / d2.y, qz = d1.z / d2.z;
  if ((qx == qy) && (qz == qy)) return false;  // collinear
  f3 v = cross0(d1, d2);
  v = unit0(v);
  mc[0] = v.x; mc[1] = v.y; mc[2] = v.z;
  mc[3] = -dot0(v, p0);
  return true;
}

__device__ __forceinline__ float det4h(const float (*m)[4], int j, int k, int mm, int nn) {
  return (m[j][0] * m[k][1] - m[k][0] * m[j][1]) * (m[mm][2] * m[nn][3] - m[nn][2] * m[mm][3]);
}
__device__ __forceinline__ float det4(const float (*m)[4]) {
  return det4h(m, 0, 1, 2, 3) - det4h(m, 0, 2, 1, 3) + det4h(m, 0, 3, 1, 2) + det4h(m, 1, 2, 0, 3) -
         det4h(m, 1, 3, 0, 2) + det4h(m, 2, 3, 0, 1);
}
__device__ inline bool estimate_sphere(const float4* xyz, const int* s, float* mc) {
  float t[4][4], x[4], y[4], z[4], sq[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f3 p = ld3(xyz, s[i]);
    x[i] = p.x; y[i] = p.y; z[i] = p.z;
    sq[i] = p.x * p.x + p.y * p.y + p.z * p.z;
    t[i][0] = x[i]; t[i][1] = y[i]; t[i][2] = z[i]; t[i][3] = 1.0f;
  }
  float m11 = det4(t);
  if (m11 == 0.0f) return false;
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i][0] = sq[i];
  float m12 = det4(t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { t[i][1] = sq[i]; t[i][0] = x[i]; }
  float m13 = det4(t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { t[i][2] = sq[i]; t[i][1] = y[i]; }
  float m14 = det4(t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { t[i][0] = sq[i]; t[i][1] = x[i]; t[i][2] = y[i]; t[i][3] = z[i]; }
  float m15 = det4(t);
  mc[0] = 0.5f * m12 / m11;
  mc[1] = 0.5f * m13 / m11;
  mc[2] = 0.5f * m14 / m11;
  mc[3] = sqrtf(mc[0] * mc[0] + mc[1] * mc[1] + mc[2] * mc[2] - m15 / m11);
  return true;
}

__device__ inline bool estimate_cylinder(const float4* xyz, const float4* nrm, const int* s, const Limits& L, float* mc) {
  const float eps = 1.1920928955078125e-07f;
  f3 p1 = ld3(xyz, s[0]), p2 = ld3(xyz, s[1]);
  if (fabsf(p1.x - p2.x) <= eps && fabsf(p1.y - p2.y) <= eps && fabsf(p1.z - p2.z) <= eps) return false;
  f3 n1 = ld3(nrm, s[0]), n2 = ld3(nrm, s[1]);
  f3 w = (n1 + p1) - p2;
  float a = dot0(n1, n1), b = dot0(n1, n2), c = dot0(n2, n2), d = dot0(n1, w), e = dot0(n2, w);
  float den = a * c - b * b;
  float sc, tc;
  if ((double)den < 1e-8) {
    sc = 0.0f;
    tc = (b > c ? d / b : e / c);
  } else {
    sc = (b * e - c * d) / den;
    tc = (a * e - b * d) / den;
  }
  f3 line_pt = (p1 + n1) + sc * n1;
  f3 line_dir = unit0((p2 + tc * n2) - line_pt);
  mc[0] = line_pt.x; mc[1] = line_pt.y; mc[2] = line_pt.z;
  mc[3] = line_dir.x; mc[4] = line_dir.y; mc[5] = line_dir.z;
  mc[6] = (float)sqrt((double)sqr_pt_line(p1, line_pt, line_dir));
  if ((double)mc[6] > L.radius_max || (double)mc[6] < L.radius_min) return false;
  return true;
}

__device__ inline bool estimate_cone(const float4* xyz, const float4* nrm, const int* s, const Limits& L, float* mc) {
  f3 p1 = ld3(xyz, s[0]), p2 = ld3(xyz, s[1]), p3 = ld3(xyz, s[2]);
  f3 n1 = ld3(nrm, s[0]), n2 = ld3(nrm, s[1]), n3 = ld3(nrm, s[2]);
  f3 o12 = cross0(n1, n2), o23 = cross0(n2, n3), o31 = cross0(n3, n1);
  float den = dot0(n1, o23);
  float d1 = dot0(p1, n1), d2 = dot0(p2, n2), d3 = dot0(p3, n3);
  f3 apex = ((d1 * o23 + d2 * o31) + d3 * o12) / den;
  f3 ap1 = p1 - apex, ap2 = p2 - apex, ap3 = p3 - apex;
  f3 np1 = apex + ap1 / nrm0(ap1), np2 = apex + ap2 / nrm0(ap2), np3 = apex + ap3 / nrm0(ap3);
  f3 axis = unit0(cross0(np2 - np1, np3 - np1));
  ap1 = unit0(ap1); ap2 = unit0(ap2); ap3 = unit0(ap3);
  float ang = ((acosf_d(dot0(ap1, axis)) + acosf_d(dot0(ap2, axis))) + acosf_d(dot0(ap3, axis))) / 3.0f;
  mc[0] = apex.x; mc[1] = apex.y; mc[2] = apex.z;
  mc[3] = axis.x; mc[4] = axis.y; mc[5] = axis.z;
  mc[6] = ang;
  if ((double)ang < L.min_angle) return false;
  if ((double)ang > L.max_angle) return false;
  return true;
}

// ------------------------------------------------------------------ isModelValid
__device__ inline bool axis_angle_ok(const Limits& L, const float* mc) {
  if (L.eps_angle > 0.0) {
    double ad = fabs(angle3d(mk3(L.ax, L.ay, L.az), mk3(mc[3], mc[4], mc[5])));
    double other = 3.14159265358979323846 - ad;
    ad = (other < ad) ? other : ad;  // std::min semantics: NaN stays NaN and never rejects
    if (ad > L.eps_angle) return false;
  }
  return true;
}
template <int MODEL>
__device__ inline bool model_valid(const Limits& L, const float* mc) {
  if (MODEL == PITT_MODEL_PLANE) return true;
  if (MODEL == PITT_MODEL_SPHERE) {
    if (L.radius_min != -DBL_MAX && (double)mc[3] < L.radius_min) return false;
    if (L.radius_max != DBL_MAX && (double)mc[3] > L.radius_max) return false;
    return true;
  }
  if (!axis_angle_ok(L, mc)) return false;
  if (MODEL == PITT_MODEL_CYLINDER) {
    if (L.radius_min != -DBL_MAX && (double)mc[6] < L.radius_min) return false;
    if (L.radius_max != DBL_MAX && (double)mc[6] > L.radius_max) return false;
    return true;
  }
  if ((double)mc[6] < L.min_angle) return false;
  if ((double)mc[6] > L.max_angle) return false;
  return true;
}

// ---- exact reformulation of the sphere predicate |fl(fl(sqrt(d2)) - r)| < thr_up as D_lo <= d2 <= D_hi.
// fl(s - r) is non-decreasing in s and fl(sqrt(d2)) is non-decreasing in d2 (both correctly rounded), so the
// inlier set is an interval of floats in d2; its end points are found by bisection on the float ordering
// (non-negative floats order like their bit patterns). No square root is left in the scoring loop.
__device__ __forceinline__ float ord_f(unsigned u) { return __uint_as_float(u); }
template <typename Pred>  // smallest non-negative finite-or-inf float (as bits, <= 0x7f800000) with pred true, pred monotone false->true
__device__ inline unsigned first_true_bits(Pred pred) {
  unsigned lo = 0u, hi = 0x7f800001u;  // hi = "none"
  while (lo < hi) {
    unsigned mid = lo + ((hi - lo) >> 1);
    if (pred(ord_f(mid))) hi = mid;
    else lo = mid + 1;
  }
  return lo;
}
__device__ inline void sphere_interval(float r, float thr_up, float& d_lo, float& d_hi) {
  d_lo = CUDART_INF_F;
  d_hi = -CUDART_INF_F;  // empty
  if (!(r == r) || !(thr_up > 0.0f)) return;
  // s range: -thr_up < fl(s - r) < thr_up
  const unsigned s_min_b = first_true_bits([&](float sv) { return (sv - r) > -thr_up; });
  const unsigned s_end_b = first_true_bits([&](float sv) { return (sv - r) >= thr_up; });  // first s that is NOT an inlier any more
  if (s_min_b > 0x7f800000u || s_end_b == 0u || s_min_b >= s_end_b) return;
  const float s_min = ord_f(s_min_b), s_max = ord_f(s_end_b - 1u);
  // d2 range: s_min <= fl(sqrt(d2)) <= s_max
  const unsigned d_lo_b = first_true_bits([&](float dv) { return sqrtf(dv) >= s_min; });
  const unsigned d_end_b = first_true_bits([&](float dv) { return sqrtf(dv) > s_max; });
  if (d_lo_b > 0x7f800000u || d_end_b == 0u || d_lo_b >= d_end_b) return;
  d_lo = ord_f(d_lo_b);
  d_hi = ord_f(d_end_b - 1u);
}

// coefficients (PCL order) -> scoring record with the per-hypothesis terms PCL hoists out of the loop
template <int MODEL>
__device__ inline void make_rec(const float* mc, bool ok, const ScoreParams& sp, HypRec& r) {
  if (!ok) {
#pragma unroll
    for (int i = 0; i < 16; ++i) r.v[i] = CUDART_NAN_F;
    return;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) r.v[i] = 0.0f;
  if (MODEL == PITT_MODEL_PLANE || MODEL == PITT_MODEL_SPHERE) {
    r.v[0] = mc[0]; r.v[1] = mc[1]; r.v[2] = mc[2]; r.v[3] = mc[3];
    if (MODEL == PITT_MODEL_SPHERE) sphere_interval(mc[3], sp.thr_up, r.v[4], r.v[5]);
  } else {
    f3 p0 = mk3(mc[0], mc[1], mc[2]), dir = mk3(mc[3], mc[4], mc[5]);
    r.v[0] = p0.x; r.v[1] = p0.y; r.v[2] = p0.z; r.v[3] = mc[6];
    r.v[4] = dir.x; r.v[5] = dir.y; r.v[6] = dir.z;
    r.v[7] = dot0(p0, dir);
    r.v[8] = 1.0f / dot0(dir, dir);
    // Certain-outlier bound shared by cylinder and cone: the weighted score is
    // w*d_normal + (1-w)*d_euclid >= (1-w)*d_euclid (d_normal in [0, pi/2], 0 <= w <= 1), so
    // d_euclid >= T = (thr + band)/(1-w) can never be an inlier, whatever the normal says.
    float T = CUDART_INF_F;  // no pre-filter unless 0 <= w < 1
    if (sp.w >= 0.0 && sp.w < 0.999) T = (float)((sp.thr + (double)sp.band) / (1.0 - sp.w) * (1.0 + 1e-5)) + 1e-30f;
    // Certain-inlier bound: the FP32 score of the fast path is w*dn + (1-w)*de with dn in [0, pi/2] whenever it is not
    // NaN, so de <= D_in = ((thr - band) - w*pi/2) / (1-w) (margins of 2e-6 relative for the float roundings of the
    // score) puts it below thr - band: the fast path returns true without looking at the normal. The NaN cases (normal
    // not finite / zero, point on the axis, coordinates so large that the float difference pt - proj could vanish) are
    // excluded by the callers: `nice` per point, axis distance >= 1 mm, |coordinates| <= 1000.
    float D_in = -1.0f;
    if (sp.w >= 0.0 && sp.w < 0.999 && fabsf(p0.x) <= 1000.0f && fabsf(p0.y) <= 1000.0f && fabsf(p0.z) <= 1000.0f) {
      const double d = (((sp.thr - (double)sp.band) * (1.0 - 2e-6) - sp.w * 1.57079651) / (1.0 - sp.w)) * (1.0 - 2e-6);
      if (d > 0.0) D_in = (float)(d * (1.0 - 3e-5));  // 3e-5: the squared-interval margins of the callers (1e-5 on sq) + rsqrt
    }
    if (MODEL == PITT_MODEL_CYLINDER) {
      // d_euclid = |sqrt(sq) - r| >= T  <=>  sq >= (r+T)^2  or  (r-T > 0 and sq <= (r-T)^2); margins of 1e-5 relative
      const float rr = mc[6];
      float lo2 = -1.0f, hi2 = CUDART_INF_F;  // (lo2, hi2) = the interval that still needs the full evaluation
      if (T < CUDART_INF_F && rr == rr) {
        const float up = rr + T;
        hi2 = up > 0.0f ? up * up * (1.0f + 1e-5f) : -1.0f;  // up <= 0: every point is a certain outlier
        const float dn = rr - T;
        if (dn > 0.0f) lo2 = dn * dn * (1.0f - 1e-5f);
      }
      r.v[9] = lo2;
      r.v[10] = hi2;
      // certain inliers: in2_lo < sq < in2_hi
      float in2_lo = CUDART_INF_F, in2_hi = -1.0f;
      if (D_in > 0.0f && rr == rr) {
        const float a = fmaxf(rr - D_in, 0.0f), b = rr + D_in;
        in2_lo = fmaxf(a * a * (1.0f + 1e-5f), 1e-6f);
        in2_hi = b * b * (1.0f - 1e-5f);
      }
      // the same four decisions on N = sqn0(cross0(dir, p0 - pt)): sq = fl(N / D) is non-decreasing in N
      const float D = dot0(dir, dir);
      float n_out_lo = 0.0f, n_out_hi = CUDART_INF_F, n_in_lo = CUDART_INF_F, n_in_end = 0.0f;
      if (D > 0.0f && D < CUDART_INF_F) {
        n_out_lo = ord_f(first_true_bits([&](float nv) { return (nv / D) > lo2; }));    // N < n_out_lo   <=> sq <= lo2
        n_out_hi = ord_f(first_true_bits([&](float nv) { return (nv / D) >= hi2; }));   // N >= n_out_hi  <=> sq >= hi2
        n_in_lo = ord_f(first_true_bits([&](float nv) { return (nv / D) > in2_lo; }));  // N >= n_in_lo   <=> sq > in2_lo
        n_in_end = ord_f(first_true_bits([&](float nv) { return (nv / D) >= in2_hi; }));  // N < n_in_end <=> sq < in2_hi
      }
      r.v[11] = n_out_lo;
      r.v[12] = n_out_hi;
      r.v[13] = n_in_lo;
      r.v[14] = n_in_end;
    }
    if (MODEL == PITT_MODEL_CONE) {
      r.v[9] = sinf_d(mc[6]);
      r.v[10] = cosf_d(mc[6]);
      double t = tan((double)mc[6]);
      r.v[11] = (float)t;
      r.v[12] = __int_as_float(__double2loint(t));
      r.v[13] = __int_as_float(__double2hiint(t));
      r.v[14] = T;
      r.v[15] = (t > 0.0) ? D_in : -1.0f;  // a negative cone radius (opening angle > 90 deg) never takes the shortcut
    }
  }
}

// ------------------------------------------------------------------ inlier predicates
// Registers-only view of a record for the scoring loops.
template <int MODEL>
struct RecRegs;

template <>
struct RecRegs<PITT_MODEL_PLANE> {
  float a, b, c, d;
  __device__ __forceinline__ void load(const HypRec* r) {
    float4 q = *reinterpret_cast<const float4*>(r->v);
    a = q.x; b = q.y; c = q.z; d = q.w;
  }
  // fabs(coeff.dot(Vector4f(x,y,z,1))) < thr   with Eigen's (p0+p2)+(p1+p3) order
  __device__ __forceinline__ bool inlier(f3 p, f3, const ScoreParams& sp) const {
    float s = (a * p.x + c * p.z) + (b * p.y + d);
    return fabsf(s) < sp.thr_up;
  }
};

template <>
struct RecRegs<PITT_MODEL_SPHERE> {
  float cx, cy, cz, r, d_lo, d_hi;
  __device__ __forceinline__ void load(const HypRec* rec) {
    float4 q = *reinterpret_cast<const float4*>(rec->v);
    float2 iv = *reinterpret_cast<const float2*>(rec->v + 4);
    cx = q.x; cy = q.y; cz = q.z; r = q.w;
    d_lo = iv.x; d_hi = iv.y;
  }
  // PCL: fabs(sqrtf(dx*dx + dy*dy + dz*dz) - r) < thr, evaluated as the equivalent interval test on the
  // squared distance (sphere_interval above): same float d2, no sqrt, bit-identical predicate
  __device__ __forceinline__ bool inlier(f3 p, f3, const ScoreParams&) const {
    float dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
    float d2 = dx * dx + dy * dy + dz * dz;
    return d2 >= d_lo && d2 <= d_hi;
  }
  // the literal PCL sequence (kept for the parity test of the interval reformulation)
  __device__ __forceinline__ bool inlier_literal(f3 p, const ScoreParams& sp) const {
    float dx = p.x - cx, dy = p.y - cy, dz = p.z - cz;
    float d = sqrtf(dx * dx + dy * dy + dz * dz) - r;
    return fabsf(d) < sp.thr_up;
  }
};

__device__ __forceinline__ bool weighted_inlier(double d_normal, double d_euclid, const ScoreParams& sp) {
  double other = 3.14159265358979323846 - d_normal;
  d_normal = (other < d_normal) ? other : d_normal;
  return fabs(sp.w * d_normal + (1.0 - sp.w) * d_euclid) < sp.thr;
}

template <>
struct RecRegs<PITT_MODEL_CYLINDER> {
  f3 p0, dir;
  float r, ptdotdir, dirdotdir, lo2, hi2;
  __device__ __forceinline__ void load(const HypRec* rec) {
    float4 q0 = *reinterpret_cast<const float4*>(rec->v);
    float4 q1 = *reinterpret_cast<const float4*>(rec->v + 4);
    float4 q2 = *reinterpret_cast<const float4*>(rec->v + 8);
    p0 = mk3(q0.x, q0.y, q0.z); r = q0.w;
    dir = mk3(q1.x, q1.y, q1.z); ptdotdir = q1.w;
    dirdotdir = q2.x; lo2 = q2.y; hi2 = q2.z;
  }
  // exact PCL sequence (float geometry, double sqrt/acos/weighting)
  __device__ __forceinline__ bool inlier_exact(f3 pt, f3 n, const ScoreParams& sp) const {
    double d_euclid = fabs(sqrt((double)sqr_pt_line(pt, p0, dir)) - (double)r);
    float k = (dot0(pt, dir) - ptdotdir) * dirdotdir;
    f3 proj = p0 + k * dir;
    f3 d = unit0(pt - proj);
    double d_normal = fabs(angle3d(n, d));
    return weighted_inlier(d_normal, d_euclid, sp);
  }
  // FP32 filter: same geometry with float sqrt/acos; only scores within sp.band of the threshold
  // are re-evaluated with the exact sequence, so the predicate is identical to inlier_exact.
  __device__ __forceinline__ bool inlier(f3 pt, f3 n, const ScoreParams& sp) const {
    float sq = sqr_pt_line(pt, p0, dir);
    // certain outlier by the euclidean term alone (make_rec): the large majority of the evaluations of
    // a random hypothesis stop here, before any sqrt / rsqrt / acos. NaN falls through and ends false.
    if (sq >= hi2 || sq <= lo2) return false;
    float de = fabsf(sqrtf(sq) - r);
    float k = (dot0(pt, dir) - ptdotdir) * dirdotdir;
    f3 d = pt - (p0 + k * dir);
    const float nd2 = sqn0(n) * sqn0(d);
    // the approximate rsqrt flushes denormals and overflows where the exact quotient does not: normals or offsets of
    // absurd length (and NaN) take the exact sequence
    if (!(nd2 > 1e-30f && nd2 < 1e30f)) return inlier_exact(pt, n, sp);
    float cosang = dot0(n, d) * rsqrtf(nd2);
    cosang = fminf(1.0f, fmaxf(-1.0f, cosang));
    float dn = acosf(cosang);
    dn = fminf(dn, 3.14159265f - dn);
    float wf = (float)sp.w;
    float score = wf * dn + (1.0f - wf) * de;
    float thr = (float)sp.thr;
    if (fabsf(score - thr) <= sp.band) return inlier_exact(pt, n, sp);
    return score < thr;  // NaN (point on the axis, zero normal, padding) is false on both paths
  }
};

template <>
struct RecRegs<PITT_MODEL_CONE> {
  f3 apex, dir;
  float angle, apexdotdir, dirdotdir, sin_a, cos_a, tan_f, T;
  double tan_a;
  __device__ __forceinline__ void load(const HypRec* rec) {
    float4 q0 = *reinterpret_cast<const float4*>(rec->v);
    float4 q1 = *reinterpret_cast<const float4*>(rec->v + 4);
    float4 q2 = *reinterpret_cast<const float4*>(rec->v + 8);
    float4 q3 = *reinterpret_cast<const float4*>(rec->v + 12);
    apex = mk3(q0.x, q0.y, q0.z); angle = q0.w;
    dir = mk3(q1.x, q1.y, q1.z); apexdotdir = q1.w;
    dirdotdir = q2.x; sin_a = q2.y; cos_a = q2.z; tan_f = q2.w;
    tan_a = __hiloint2double(__float_as_int(q3.y), __float_as_int(q3.x));
    T = q3.z;
  }
  __device__ __forceinline__ bool inlier_exact(f3 pt, f3 n, const ScoreParams& sp) const {
    float k = (dot0(pt, dir) - apexdotdir) * dirdotdir;
    f3 proj = apex + k * dir;
    f3 pp = unit0(pt - proj);
    f3 height = apex - proj;
    double actual_r = tan_a * (double)nrm0(height);
    height = unit0(height);
    f3 cone_normal = sin_a * height + cos_a * pp;
    double d_euclid = fabs(sqrt((double)sqr_pt_line(pt, apex, dir)) - actual_r);
    double d_normal = fabs(angle3d(n, cone_normal));
    return weighted_inlier(d_normal, d_euclid, sp);
  }
  __device__ __forceinline__ bool inlier(f3 pt, f3 n, const ScoreParams& sp) const {
    float k = (dot0(pt, dir) - apexdotdir) * dirdotdir;
    f3 proj = apex + k * dir;
    f3 pp = pt - proj;
    f3 height = apex - proj;
    const float hn2 = sqn0(height);
    const float sq = sqr_pt_line(pt, apex, dir);
    {
      // certain outlier by the euclidean term alone: |sqrt(sq) - tan*|height|| >= T (make_rec). Squares are
      // compared so that only one approximate rsqrt is needed; 1e-5 relative margins cover the float roundings.
      const float ar = tan_f * (hn2 * rsqrtf(hn2));  // ~ tan * |height|; hn2 == 0 gives NaN and falls through
      const float up = fabsf(ar) + T, dn = fabsf(ar) - T;
      if (sq >= up * up * (1.0f + 1e-5f)) return false;
      if (dn > 0.0f && sq <= dn * dn * (1.0f - 1e-5f)) return false;
    }
    float hn = sqrtf(hn2);
    float ppn = nrm0(pp);
    float actual_r = (float)tan_a * hn;
    // cone normal = sin * unit(height) + cos * unit(pp)
    float ih = 1.0f / hn, ip = 1.0f / ppn;
    f3 cn = (sin_a * ih) * height + (cos_a * ip) * pp;
    float de = fabsf(sqrtf(sq) - actual_r);
    const float nc2 = sqn0(n) * sqn0(cn);
    if (!(nc2 > 1e-30f && nc2 < 1e30f) || !(hn2 > 1e-30f)) return inlier_exact(pt, n, sp);  // see the cylinder
    float cosang = dot0(n, cn) * rsqrtf(nc2);
    cosang = fminf(1.0f, fmaxf(-1.0f, cosang));
    float dn = acosf(cosang);
    dn = fminf(dn, 3.14159265f - dn);
    float wf = (float)sp.w;
    float score = wf * dn + (1.0f - wf) * de;
    float thr = (float)sp.thr;
    if (fabsf(score - thr) <= sp.band) return inlier_exact(pt, n, sp);
    return score < thr;  // NaN (point on the axis, zero normal, padding) is false on both paths
  }
};

// ------------------------------------------------------------------ tier-1 classification (cylinder, cone)
// The scoring kernel decides most evaluations from the axis distance alone: 0 = certain outlier, 1 = certain inlier
// (both proven in make_rec to be what RecRegs::inlier would return), 2 = undecided -> RecRegs::inlier on a compacted
// queue. `nice` = the point may take the certain-inlier shortcut (finite normal of sane length, |coordinates| <= 1000).
__device__ __forceinline__ bool nice_point(f3 pt, f3 nv) {
  const float n2 = sqn0(nv);
  return fabsf(pt.x) <= 1000.0f && fabsf(pt.y) <= 1000.0f && fabsf(pt.z) <= 1000.0f && fabsf(nv.x) < 1e9f && fabsf(nv.y) < 1e9f &&
         fabsf(nv.z) < 1e9f && n2 > 1e-18f && n2 < 1e18f;
}

template <int MODEL>
struct Tier1;

template <>
struct Tier1<PITT_MODEL_CYLINDER> {
  f3 p0, dir;
  float n_out_lo, n_out_hi, n_in_lo, n_in_end;
  __device__ __forceinline__ void load(const HypRec* rec) {
    float4 q0 = *reinterpret_cast<const float4*>(rec->v);
    float4 q1 = *reinterpret_cast<const float4*>(rec->v + 4);
    float4 q2 = *reinterpret_cast<const float4*>(rec->v + 8);
    float4 q3 = *reinterpret_cast<const float4*>(rec->v + 12);
    p0 = mk3(q0.x, q0.y, q0.z);
    dir = mk3(q1.x, q1.y, q1.z);
    n_out_lo = q2.w; n_out_hi = q3.x; n_in_lo = q3.y; n_in_end = q3.z;
  }
  // in = certain inlier, und = undecided (needs RecRegs::inlier); neither = certain outlier. NaN ends undecided.
  __device__ __forceinline__ void classify(f3 pt, bool nice, bool& in, bool& und) const {
    const float N = sqn0(cross0(dir, p0 - pt));  // numerator of pcl::sqrPointToLineDistance, same operations as sqr_pt_line
    const bool out = (N >= n_out_hi) | (N < n_out_lo);
    in = nice & (N >= n_in_lo) & (N < n_in_end);
    und = !(out | in);
  }
};

template <>
struct Tier1<PITT_MODEL_CONE> {
  f3 apex, dir;
  float apexdotdir, dirdotdir, tan_f, T, D_in, Dp, Dm;
  __device__ __forceinline__ void load(const HypRec* rec) {
    float4 q0 = *reinterpret_cast<const float4*>(rec->v);
    float4 q1 = *reinterpret_cast<const float4*>(rec->v + 4);
    float4 q2 = *reinterpret_cast<const float4*>(rec->v + 8);
    float4 q3 = *reinterpret_cast<const float4*>(rec->v + 12);
    apex = mk3(q0.x, q0.y, q0.z);
    dir = mk3(q1.x, q1.y, q1.z); apexdotdir = q1.w;
    dirdotdir = q2.x; tan_f = q2.w;
    T = q3.z; D_in = q3.w;
    // sq = fl(N / D) with N = |dir x (apex - pt)|^2, D = dir.dir: the tests on sq (margins of 1e-5, see make_rec) are made
    // on N against bounds scaled by D (1 +- 2e-5), which implies them whatever the rounding of the division: no division here
    const float D = dot0(dir, dir);
    Dp = D * (1.0f + 2e-5f);
    Dm = D * (1.0f - 2e-5f);
  }
  __device__ __forceinline__ void classify(f3 pt, bool nice, bool& in, bool& und) const {
    // same float quantities as RecRegs<CONE>::inlier
    const float k = (dot0(pt, dir) - apexdotdir) * dirdotdir;
    const f3 proj = apex + k * dir;
    const f3 height = apex - proj;
    const float hn2 = sqn0(height);
    const float N = sqn0(cross0(dir, apex - pt));  // numerator of sqr_pt_line(pt, apex, dir)
    const float ar = tan_f * (hn2 * rsqrtf(hn2));  // ~ tan * |height| (within 1e-6 relative of the fast path's actual_r)
    const float up = fabsf(ar) + T, dn = fabsf(ar) - T;
    const bool out = (N >= up * up * Dp) | ((dn > 0.0f) & (N <= dn * dn * Dm));
    // D_in > 0 implies tan > 0: ar is the cone radius itself
    const float b = ar + D_in, a = fmaxf(ar - D_in, 0.0f);
    in = nice & (D_in > 0.0f) & (hn2 >= 1e-12f) & (N >= 1e-6f * Dp) & (N < b * b * Dm) & (N > a * a * Dp) & !out;
    und = !(out | in);
  }
};

}  // namespace pitt
