// orc_prefilter.cpp — CPU ORACLE (test infrastructure) of the pre-path of depthAcquisition:
// PCManager::downSampling (pc_manager.cpp:55-67 -> pcl::VoxelGrid<PointXYZ>::applyFilter, PCL 1.7.x
// filters/impl/voxel_grid.hpp), the deep filter (deep_filter_srv.cpp:27-58) and
// pcl::transformPointCloud with a Matrix4f (obj_segmentation.cpp:248; PCL 1.7.2 common/impl/transforms.hpp).
//
// Pinned choices (PCL leaves them to the build):
//  * VoxelGrid sorts (voxel index, point index) pairs with std::sort on the voxel index only, which is not
//    stable: the order in which the points of one voxel are accumulated is unspecified upstream. Here they
//    are accumulated in ascending point index (what a stable sort gives), in float, like PCL.
//  * `centroid /= float(count)` is a true division (the same convention as the covariance code of this
//    oracle; Eigen releases differ between division and multiplication by the reciprocal).
//  * transformPointCloud: x' = m00*x + m01*y + m02*z + m03 evaluated left to right in float, no FMA
//    (the scalar form of PCL 1.7.2).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>
#include <algorithm>

#include "oracle.h"

namespace {

struct P3 {
  float x, y, z;
};

// returns false when PCL would warn "Leaf size is too small" and pass the input through
bool voxel_grid(const std::vector<P3>& in, const float leaf[3], std::vector<P3>* out) {
  out->clear();
  float inv[3];
  for (int a = 0; a < 3; ++a) inv[a] = 1.0f / leaf[a];  // Eigen::Array4f::Ones () / leaf_size_.array ()
  // getMinMax3D over the finite points
  float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
  float mx[3] = {-std::numeric_limits<float>::max(), -std::numeric_limits<float>::max(), -std::numeric_limits<float>::max()};
  size_t n_finite = 0;
  for (const P3& p : in) {
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    const float v[3] = {p.x, p.y, p.z};
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], v[a]);
      mx[a] = std::max(mx[a], v[a]);
    }
    ++n_finite;
  }
  if (n_finite == 0) return true;
  const int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv[0]) + 1;
  const int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv[1]) + 1;
  const int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv[2]) + 1;
  if ((dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
    *out = in;
    return false;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = static_cast<int>(std::floor(mn[a] * inv[a]));
    max_b[a] = static_cast<int>(std::floor(mx[a] * inv[a]));
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  struct Entry {
    unsigned idx;
    unsigned pt;
  };
  std::vector<Entry> index_vector;
  index_vector.reserve(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    const P3& p = in[i];
    if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
    const int ijk0 = static_cast<int>(std::floor(p.x * inv[0]) - static_cast<float>(min_b[0]));
    const int ijk1 = static_cast<int>(std::floor(p.y * inv[1]) - static_cast<float>(min_b[1]));
    const int ijk2 = static_cast<int>(std::floor(p.z * inv[2]) - static_cast<float>(min_b[2]));
    const int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
    index_vector.push_back({static_cast<unsigned>(idx), static_cast<unsigned>(i)});
  }
  std::stable_sort(index_vector.begin(), index_vector.end(), [](const Entry& a, const Entry& b) { return a.idx < b.idx; });
  size_t index = 0;
  while (index < index_vector.size()) {
    size_t i = index + 1;
    while (i < index_vector.size() && index_vector[i].idx == index_vector[index].idx) ++i;
    float c[3] = {in[index_vector[index].pt].x, in[index_vector[index].pt].y, in[index_vector[index].pt].z};
    for (size_t j = index + 1; j < i; ++j) {
      c[0] += in[index_vector[j].pt].x;
      c[1] += in[index_vector[j].pt].y;
      c[2] += in[index_vector[j].pt].z;
    }
    const float cnt = static_cast<float>(i - index);
    out->push_back({c[0] / cnt, c[1] / cnt, c[2] / cnt});
    index = i;
  }
  return true;
}

}  // namespace

extern "C" int orc_prefilter(const void* data, int point_step, int n_points, const pitt_prefilter_params* p, float* out4,
                             int cap, int* n_out, pitt_prefilter_info* info) {
  std::vector<P3> cloud((size_t)std::max(n_points, 0));
  for (int i = 0; i < n_points; ++i) {
    const float* f = reinterpret_cast<const float*>(static_cast<const unsigned char*>(data) + (size_t)i * point_step);
    cloud[i] = {f[0], f[1], f[2]};
  }
  pitt_prefilter_info I;
  memset(&I, 0, sizeof(I));
  I.n_input = n_points;
  std::vector<P3> cur = cloud;
  if (p->leaf[0] > 0.0f && p->leaf[1] > 0.0f && p->leaf[2] > 0.0f) {
    std::vector<P3> ds;
    if (!voxel_grid(cur, p->leaf, &ds)) I.voxel_overflow = 1;
    cur.swap(ds);
  }
  I.n_voxel = (int)cur.size();
  const float thDeep = p->deep_threshold >= 0.0f ? p->deep_threshold : 3.000f;  // srvm::getServiceFloatParameter
  I.used_deep_threshold = thDeep;
  if (p->apply_deep_filter) {
    std::vector<P3> closer;
    int further = 0;
    for (const P3& q : cur) {
      if (q.z == q.z) {
        if (q.z > thDeep) ++further;
        else closer.push_back(q);
      }
    }
    I.n_further = further;
    cur.swap(closer);
  }
  I.n_closer = (int)cur.size();
  if (p->apply_transform) {
    const float* m = p->transform;
    for (P3& q : cur) {
      if (!std::isfinite(q.x) || !std::isfinite(q.y) || !std::isfinite(q.z)) continue;  // non-dense clouds: left as is
      const float x = q.x, y = q.y, z = q.z;
      q.x = m[0] * x + m[1] * y + m[2] * z + m[3];
      q.y = m[4] * x + m[5] * y + m[6] * z + m[7];
      q.z = m[8] * x + m[9] * y + m[10] * z + m[11];
    }
  }
  *n_out = (int)cur.size();
  if (info) *info = I;
  if ((int)cur.size() > cap) return PITT_ERR_CAPACITY;
  for (size_t i = 0; i < cur.size(); ++i) {
    out4[4 * i] = cur[i].x; out4[4 * i + 1] = cur[i].y; out4[4 * i + 2] = cur[i].z; out4[4 * i + 3] = 1.0f;
  }
  return PITT_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Arm filter: segmentation_services/arm_filter_srv.cpp. armFiltering (:66-103) = one pcl::CropBox with
// setMin/setMax (corners in the link frame), setTranslation (tf origin :78-80), setRotation (roll, pitch, yaw of the tf
// rotation :73, :83-85), setTransform(identity), setNegative(true); filter() chains four of them, each on the previous
// output (:134-141). PCL 1.7.2 filters/impl/crop_box.hpp, restated (PARITY UNPINNED, like the rest of the oracle):
//   transform = getTransformation(0, 0, 0, roll, pitch, yaw) and inverse_transform = transform.inverse() when the
//   rotation is not zero; per point: skipped when the cloud is not dense and the point is not finite; local = point;
//   local -= translation when the translation is not zero; local = inverse_transform * local when it is not the identity;
//   "outside" = any coordinate below min or above max; negative => the outside points are kept, in order. The output of
//   a CropBox is marked dense, so only the first box of the chain looks at finiteness.
// Pinned choices: cosf/sinf = the double function rounded once; Eigen's Affine inverse of a rotation = the 3x3 cofactor
// inverse (compute_inverse_size3: det from the first column, (a + b) + c); matrix * vector = (m0 x + m1 y) + m2 z.
namespace {
struct Box {
  bool rot, tr;
  float inv[3][3];
  float t[3], mn[3], mx[3];
};
void rotation_inverse(const float rpy[3], float inv[3][3]) {
  const float roll = rpy[0], pitch = rpy[1], yaw = rpy[2];
  const float A = (float)std::cos((double)yaw), B = (float)std::sin((double)yaw);
  const float C = (float)std::cos((double)pitch), D = (float)std::sin((double)pitch);
  const float E = (float)std::cos((double)roll), F = (float)std::sin((double)roll);
  const float DE = D * E, DF = D * F;
  const float t[3][3] = {{A * C, A * DF - B * E, B * F + A * DE}, {B * C, A * E + B * DF, B * DE - A * F}, {-D, C * F, C * E}};
  float cofactor[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const int ia = (i + 1) % 3, ib = (i + 2) % 3, ja = (j + 1) % 3, jb = (j + 2) % 3;
      cofactor[i][j] = t[ia][ja] * t[ib][jb] - t[ia][jb] * t[ib][ja];
    }
  const float det = (cofactor[0][0] * t[0][0] + cofactor[1][0] * t[1][0]) + cofactor[2][0] * t[2][0];
  const float invdet = 1.0f / det;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) inv[r][c] = cofactor[c][r] * invdet;
}
}  // namespace

extern "C" int orc_arm_filter(const float* xyz4, int n, const pitt_arm_filter_params* p, float* out4, int cap, int* n_out,
                              int* removed) {
  if (!p || p->n_boxes < 0 || p->n_boxes > 4 || !n_out) return PITT_ERR_INVALID;
  std::vector<P3> cur((size_t)std::max(n, 0));
  for (int i = 0; i < n; ++i) cur[i] = {xyz4[4 * i], xyz4[4 * i + 1], xyz4[4 * i + 2]};
  bool dense = p->input_is_dense != 0;
  for (int k = 0; k < p->n_boxes; ++k) {
    const pitt_crop_box& b = p->box[k];
    Box B;
    B.rot = !(b.rotation_rpy[0] == 0.0f && b.rotation_rpy[1] == 0.0f && b.rotation_rpy[2] == 0.0f);
    B.tr = !(b.translation[0] == 0.0f && b.translation[1] == 0.0f && b.translation[2] == 0.0f);
    if (B.rot) rotation_inverse(b.rotation_rpy, B.inv);
    std::vector<P3> kept;
    for (const P3& q : cur) {
      if (!dense && !(std::isfinite(q.x) && std::isfinite(q.y) && std::isfinite(q.z))) continue;
      P3 l = q;
      if (B.tr) { l.x -= b.translation[0]; l.y -= b.translation[1]; l.z -= b.translation[2]; }
      if (B.rot) {
        const float x = l.x, y = l.y, z = l.z;
        l.x = (B.inv[0][0] * x + B.inv[0][1] * y) + B.inv[0][2] * z;
        l.y = (B.inv[1][0] * x + B.inv[1][1] * y) + B.inv[1][2] * z;
        l.z = (B.inv[2][0] * x + B.inv[2][1] * y) + B.inv[2][2] * z;
      }
      const bool outside = (l.x < b.min_pt[0] || l.y < b.min_pt[1] || l.z < b.min_pt[2]) ||
                           (l.x > b.max_pt[0] || l.y > b.max_pt[1] || l.z > b.max_pt[2]);
      if (outside) kept.push_back(q);
    }
    if (removed) removed[k] = (int)(cur.size() - kept.size());
    cur.swap(kept);
    dense = true;  // CropBox marks its output dense
  }
  if (removed) for (int k = p->n_boxes; k < 4; ++k) removed[k] = 0;
  *n_out = (int)cur.size();
  if ((int)cur.size() > cap) return PITT_ERR_CAPACITY;
  for (size_t i = 0; i < cur.size(); ++i) {
    out4[4 * i] = cur[i].x; out4[4 * i + 1] = cur[i].y; out4[4 * i + 2] = cur[i].z; out4[4 * i + 3] = 1.0f;
  }
  return PITT_OK;
}
