// orc_sac.h — CPU ORACLE (test infrastructure only). Sample consensus models of PCL 1.7.x as
// reached from the reference's seg.segment() call sites; see orc_math.h for the parity status.
#pragma once
#include <cfloat>
#include <climits>
#include <unordered_map>

#include "../include/pitt_b200.h"
#include "orc_math.h"

namespace orc {

// ---------------------------------------------------------------- boost::mt19937 (seed 12345u)
struct MT19937 {
  uint32_t s[624];
  int idx;
  explicit MT19937(uint32_t seed = 12345u) {
    s[0] = seed;
    for (int i = 1; i < 624; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int i = 0; i < 624; ++i) {
        uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
        s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = s[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};

struct Cloud {
  const float* xyz;  // n x {x,y,z,pad}
  const float* nrm;  // n x {nx,ny,nz,curv} or nullptr
  int n;
  V4 p(int i) const { return mk(xyz[4 * i], xyz[4 * i + 1], xyz[4 * i + 2], 0.0f); }
  V4 nv(int i) const { return mk(nrm[4 * i], nrm[4 * i + 1], nrm[4 * i + 2], 0.0f); }
};

inline int sampleSize(int model) {
  switch (model) {
    case PITT_MODEL_PLANE: return 3;
    case PITT_MODEL_SPHERE: return 4;
    case PITT_MODEL_CYLINDER: return 2;
    default: return 3;
  }
}
inline int coeffCount(int model) { return (model == PITT_MODEL_PLANE || model == PITT_MODEL_SPHERE) ? 4 : 7; }

// SampleConsensusModelPlane::isSampleGood (sac_model_plane.hpp): element-wise Array4f quotient.
inline bool planeSampleGood(const Cloud& c, const int* s) {
  float q[3];
  for (int k = 0; k < 3; ++k) {
    float a = c.xyz[4 * s[1] + k] - c.xyz[4 * s[0] + k];
    float b = c.xyz[4 * s[2] + k] - c.xyz[4 * s[0] + k];
    q[k] = a / b;
  }
  return (q[0] != q[1]) || (q[2] != q[1]);
}

// pcl::SampleConsensusModel::getSamples / drawIndexSample (sac_model.h): rnd() = mt() >> 1
// (boost::uniform_int<>(0,INT_MAX) over a 32-bit engine), persistent shuffled_indices_.
// The shuffle is kept sparse (only displaced entries) so that a 50 M-point cloud costs O(H).
struct PclSampler {
  MT19937 mt;
  std::unordered_map<int, int> moved;
  int n;
  explicit PclSampler(int n_) : mt(12345u), n(n_) {}
  int get(int i) const {
    auto it = moved.find(i);
    return it == moved.end() ? i : it->second;
  }
  void draw(int S, int* out) {
    for (int i = 0; i < S; ++i) {
      uint32_t r = mt.next() >> 1;
      int j = i + (int)(r % (uint32_t)(n - i));
      int a = get(i), b = get(j);
      moved[i] = b;
      moved[j] = a;
    }
    for (int i = 0; i < S; ++i) out[i] = get(i);
  }
  // returns false when no good sample could be drawn (samples.clear() in PCL)
  bool getSamples(const Cloud& c, int model, int* out) {
    int S = sampleSize(model);
    if (n < S) return false;
    for (int iter = 0; iter < 1000; ++iter) {
      draw(S, out);
      if (model != PITT_MODEL_PLANE || planeSampleGood(c, out)) return true;
    }
    return false;
  }
};

// ---------------------------------------------------------------- model limits (initSACModel)
struct Limits {
  double radius_min = -DBL_MAX, radius_max = DBL_MAX;
  double min_angle = -DBL_MAX, max_angle = DBL_MAX;
  double eps_angle = 0.0;
  double w = 0.0;  // normal_distance_weight_ of the model
  V4 axis = mk(0, 0, 0, 0);
};
inline Limits limitsFor(const pitt_sac_params& p) {
  Limits L;
  bool fwd_radius = (p.radius_min != -DBL_MAX) && (p.radius_max != DBL_MAX);
  if (p.model == PITT_MODEL_SPHERE) {
    if (fwd_radius) { L.radius_min = p.radius_min; L.radius_max = p.radius_max; }
  } else if (p.model == PITT_MODEL_CYLINDER) {
    if (fwd_radius) { L.radius_min = p.radius_min; L.radius_max = p.radius_max; }
    L.w = p.normal_distance_weight;
    if (p.axis[0] != 0.0f || p.axis[1] != 0.0f || p.axis[2] != 0.0f) L.axis = mk(p.axis[0], p.axis[1], p.axis[2], 0);
    if (p.eps_angle != 0.0) L.eps_angle = p.eps_angle;
  } else if (p.model == PITT_MODEL_CONE) {
    L.w = p.normal_distance_weight;
    if (p.axis[0] != 0.0f || p.axis[1] != 0.0f || p.axis[2] != 0.0f) L.axis = mk(p.axis[0], p.axis[1], p.axis[2], 0);
    if (p.eps_angle != 0.0) L.eps_angle = p.eps_angle;
    // setMinMaxOpeningAngle is forwarded when both differ from the model defaults
    if (p.min_angle != -DBL_MAX && p.max_angle != DBL_MAX) { L.min_angle = p.min_angle; L.max_angle = p.max_angle; }
  }
  return L;
}

// ---------------------------------------------------------------- computeModelCoefficients
inline float det4(const float m[4][4]) {
  // Eigen 3.2 determinant_impl<Derived,4> (bruteforce_det4_helper, "trick by Martin Costabel")
  auto h = [&](int j, int k, int mm, int nn) {
    return (m[j][0] * m[k][1] - m[k][0] * m[j][1]) * (m[mm][2] * m[nn][3] - m[nn][2] * m[mm][3]);
  };
  return h(0, 1, 2, 3) - h(0, 2, 1, 3) + h(0, 3, 1, 2) + h(1, 2, 0, 3) - h(1, 3, 0, 2) + h(2, 3, 0, 1);
}

inline bool planeCoeffs(const Cloud& c, const int* s, float* mc) {
  float p0[3], d1[3], d2[3], q[3];
  for (int k = 0; k < 3; ++k) {
    p0[k] = c.xyz[4 * s[0] + k];
    d1[k] = c.xyz[4 * s[1] + k] - p0[k];
    d2[k] = c.xyz[4 * s[2] + k] - p0[k];
    q[k] = d1[k] / d2[k];
  }
  if ((q[0] == q[1]) && (q[2] == q[1])) return false;
  V4 v = mk(d1[1] * d2[2] - d1[2] * d2[1], d1[2] * d2[0] - d1[0] * d2[2], d1[0] * d2[1] - d1[1] * d2[0], 0.0f);
  v = normalized4(v);
  // PointXYZ::data[3] == 1 but the coefficient is still 0 here
  float d = -1.0f * dot4(v, mk(p0[0], p0[1], p0[2], 1.0f));
  mc[0] = v[0]; mc[1] = v[1]; mc[2] = v[2]; mc[3] = d;
  return true;
}

inline bool sphereCoeffs(const Cloud& c, const int* s, float* mc) {
  float t[4][4], x[4], y[4], z[4], sq[4];
  for (int i = 0; i < 4; ++i) {
    x[i] = c.xyz[4 * s[i]]; y[i] = c.xyz[4 * s[i] + 1]; z[i] = c.xyz[4 * s[i] + 2];
    sq[i] = x[i] * x[i] + y[i] * y[i] + z[i] * z[i];
  }
  for (int i = 0; i < 4; ++i) { t[i][0] = x[i]; t[i][1] = y[i]; t[i][2] = z[i]; t[i][3] = 1.0f; }
  float m11 = det4(t);
  if (m11 == 0.0f) return false;
  for (int i = 0; i < 4; ++i) t[i][0] = sq[i];
  float m12 = det4(t);
  for (int i = 0; i < 4; ++i) { t[i][1] = t[i][0]; t[i][0] = x[i]; }
  float m13 = det4(t);
  for (int i = 0; i < 4; ++i) { t[i][2] = t[i][1]; t[i][1] = y[i]; }
  float m14 = det4(t);
  for (int i = 0; i < 4; ++i) { t[i][0] = t[i][2]; t[i][1] = x[i]; t[i][2] = y[i]; t[i][3] = z[i]; }
  float m15 = det4(t);
  mc[0] = 0.5f * m12 / m11;
  mc[1] = 0.5f * m13 / m11;
  mc[2] = 0.5f * m14 / m11;
  mc[3] = sqrtf(mc[0] * mc[0] + mc[1] * mc[1] + mc[2] * mc[2] - m15 / m11);
  return true;
}

inline bool cylinderCoeffs(const Cloud& c, const int* s, const Limits& L, float* mc) {
  const float eps = std::numeric_limits<float>::epsilon();
  if (fabsf(c.xyz[4 * s[0]] - c.xyz[4 * s[1]]) <= eps && fabsf(c.xyz[4 * s[0] + 1] - c.xyz[4 * s[1] + 1]) <= eps &&
      fabsf(c.xyz[4 * s[0] + 2] - c.xyz[4 * s[1] + 2]) <= eps)
    return false;
  V4 p1 = c.p(s[0]), p2 = c.p(s[1]), n1 = c.nv(s[0]), n2 = c.nv(s[1]);
  V4 w = (n1 + p1) - p2;
  float a = dot4(n1, n1), b = dot4(n1, n2), cc = dot4(n2, n2), d = dot4(n1, w), e = dot4(n2, w);
  float den = a * cc - b * b;
  float sc, tc;
  if ((double)den < 1e-8) {
    sc = 0.0f;
    tc = (b > cc ? d / b : e / cc);
  } else {
    sc = (b * e - cc * d) / den;
    tc = (a * e - b * d) / den;
  }
  V4 line_pt = (p1 + n1) + sc * n1;
  V4 line_dir = normalized4((p2 + tc * n2) - line_pt);
  mc[0] = line_pt[0]; mc[1] = line_pt[1]; mc[2] = line_pt[2];
  mc[3] = line_dir[0]; mc[4] = line_dir[1]; mc[5] = line_dir[2];
  mc[6] = (float)sqrt(sqrPointToLineDistance(p1, line_pt, line_dir));
  if ((double)mc[6] > L.radius_max || (double)mc[6] < L.radius_min) return false;
  return true;
}

inline bool coneCoeffs(const Cloud& c, const int* s, const Limits& L, float* mc) {
  V4 p1 = c.p(s[0]), p2 = c.p(s[1]), p3 = c.p(s[2]);
  V4 n1 = c.nv(s[0]), n2 = c.nv(s[1]), n3 = c.nv(s[2]);
  V4 o12 = cross3(n1, n2), o23 = cross3(n2, n3), o31 = cross3(n3, n1);
  float den = dot4(n1, o23);
  float d1 = dot4(p1, n1), d2 = dot4(p2, n2), d3 = dot4(p3, n3);
  V4 apex = ((d1 * o23 + d2 * o31) + d3 * o12) / den;
  V4 ap1 = p1 - apex, ap2 = p2 - apex, ap3 = p3 - apex;
  V4 np1 = apex + ap1 / norm4(ap1), np2 = apex + ap2 / norm4(ap2), np3 = apex + ap3 / norm4(ap3);
  V4 axis = normalized4(cross3(np2 - np1, np3 - np1));
  ap1 = normalized4(ap1); ap2 = normalized4(ap2); ap3 = normalized4(ap3);
  float ang = ((acosf_d(dot4(ap1, axis)) + acosf_d(dot4(ap2, axis))) + acosf_d(dot4(ap3, axis))) / 3.0f;
  mc[0] = apex[0]; mc[1] = apex[1]; mc[2] = apex[2];
  mc[3] = axis[0]; mc[4] = axis[1]; mc[5] = axis[2];
  mc[6] = ang;
  if ((double)ang != -DBL_MAX && (double)ang < L.min_angle) return false;
  if ((double)ang != DBL_MAX && (double)ang > L.max_angle) return false;
  return true;
}

inline bool computeModelCoefficients(const Cloud& c, int model, const int* s, const Limits& L, float* mc) {
  switch (model) {
    case PITT_MODEL_PLANE: return planeCoeffs(c, s, mc);
    case PITT_MODEL_SPHERE: return sphereCoeffs(c, s, mc);
    case PITT_MODEL_CYLINDER: return cylinderCoeffs(c, s, L, mc);
    default: return coneCoeffs(c, s, L, mc);
  }
}

// ---------------------------------------------------------------- isModelValid
inline bool axisAngleOk(const Limits& L, const float* mc) {
  if (L.eps_angle > 0.0) {
    V4 coeff = mk(mc[3], mc[4], mc[5], 0);
    double ad = fabs(getAngle3D(L.axis, coeff));
    double other = M_PI - ad;
    ad = (other < ad) ? other : ad;  // std::min(a,b): b<a ? b : a — NaN stays NaN
    if (ad > L.eps_angle) return false;
  }
  return true;
}
inline bool isModelValid(int model, const Limits& L, const float* mc) {
  if (model == PITT_MODEL_PLANE) return true;
  if (model == PITT_MODEL_SPHERE) {
    if (L.radius_min != -DBL_MAX && (double)mc[3] < L.radius_min) return false;
    if (L.radius_max != DBL_MAX && (double)mc[3] > L.radius_max) return false;
    return true;
  }
  if (!axisAngleOk(L, mc)) return false;
  if (model == PITT_MODEL_CYLINDER) {
    if (L.radius_min != -DBL_MAX && (double)mc[6] < L.radius_min) return false;
    if (L.radius_max != DBL_MAX && (double)mc[6] > L.radius_max) return false;
    return true;
  }
  if ((double)mc[6] != -DBL_MAX && (double)mc[6] < L.min_angle) return false;
  if ((double)mc[6] != DBL_MAX && (double)mc[6] > L.max_angle) return false;
  return true;
}

// ---------------------------------------------------------------- per-point inlier predicate
// One object per (model, coefficients): hoists what PCL hoists out of the point loop.
struct Scorer {
  int model;
  double thr, w;
  float mc[7];
  // cylinder / cone
  V4 pt0, dir;
  float ptdotdir, dirdotdir;
  double tan_a;
  float sin_a, cos_a;
  Scorer(int model_, const Limits& L, double thr_, const float* c) : model(model_), thr(thr_), w(L.w) {
    for (int i = 0; i < coeffCount(model); ++i) mc[i] = c[i];
    if (model == PITT_MODEL_CYLINDER || model == PITT_MODEL_CONE) {
      pt0 = mk(mc[0], mc[1], mc[2], 0);
      dir = mk(mc[3], mc[4], mc[5], 0);
      ptdotdir = dot4(pt0, dir);
      dirdotdir = 1.0f / dot4(dir, dir);
      if (model == PITT_MODEL_CONE) {
        tan_a = tan((double)mc[6]);  // `tan(opening_angle)` → ::tan(double). PINNED CHOICE.
        sin_a = sinf_d(mc[6]);
        cos_a = cosf_d(mc[6]);
      }
    }
  }
  inline bool inlier(const Cloud& c, int i) const {
    const float x = c.xyz[4 * i], y = c.xyz[4 * i + 1], z = c.xyz[4 * i + 2];
    switch (model) {
      case PITT_MODEL_PLANE: {
        // fabs(model_coefficients.dot(Vector4f(x,y,z,1))) < threshold
        float d = dot4(mk(mc[0], mc[1], mc[2], mc[3]), mk(x, y, z, 1.0f));
        return (double)fabsf(d) < thr;
      }
      case PITT_MODEL_SPHERE: {
        float dx = x - mc[0], dy = y - mc[1], dz = z - mc[2];
        float d = sqrtf(dx * dx + dy * dy + dz * dz) - mc[3];
        return (double)fabsf(d) < thr;
      }
      case PITT_MODEL_CYLINDER: {
        V4 pt = mk(x, y, z, 0), n = c.nv(i);
        double d_euclid = fabs(sqrt(sqrPointToLineDistance(pt, pt0, dir)) - (double)mc[6]);
        float k = (dot4(pt, dir) - ptdotdir) * dirdotdir;
        V4 pt_proj = pt0 + k * dir;
        V4 d = normalized4(pt - pt_proj);
        double d_normal = fabs(getAngle3D(n, d));
        double other = M_PI - d_normal;
        d_normal = (other < d_normal) ? other : d_normal;
        return fabs(w * d_normal + (1.0 - w) * d_euclid) < thr;
      }
      default: {
        V4 pt = mk(x, y, z, 0), n = c.nv(i);
        float k = (dot4(pt, dir) - ptdotdir) * dirdotdir;
        V4 pt_proj = pt0 + k * dir;
        V4 pp = normalized4(pt - pt_proj);
        V4 height = pt0 - pt_proj;
        double actual_r = tan_a * (double)norm4(height);
        height = normalized4(height);
        V4 cone_normal = sin_a * height + cos_a * pp;
        double d_euclid = fabs(sqrt(sqrPointToLineDistance(pt, pt0, dir)) - actual_r);
        double d_normal = fabs(getAngle3D(n, cone_normal));
        double other = M_PI - d_normal;
        d_normal = (other < d_normal) ? other : d_normal;
        return fabs(w * d_normal + (1.0 - w) * d_euclid) < thr;
      }
    }
  }
};

int countWithinDistance(const Cloud& c, int model, const Limits& L, double thr, const float* mc);
void selectWithinDistance(const Cloud& c, int model, const Limits& L, double thr, const float* mc, std::vector<int>& out);
// optimizeModelCoefficients; returns LM info (0 for plane), nfev through *nfev
int optimizeModelCoefficients(const Cloud& c, int model, const std::vector<int>& inliers, const float* mc, float* refined, int* nfev);

}  // namespace orc
