// orc_services.cpp — CPU ORACLE (test infrastructure only): the reference's own service glue,
// restated around the oracle's PCL restatements. Paths under /root/reference/src:
//   findSupports            segmentation_services/supports_segmentation_srv.cpp:241-361
//   clusterize              segmentation_services/cluster_segmentation_srv.cpp:38-108
//   ransac*Detection        segmentation_services/{plane,sphere,cylinder,cone}_segmentation_srv.cpp
//   inlierToVectorMsg       point_cloud_library/pc_manager.cpp:105-111
//   clustersAcquisition     ransac_segmentation.cpp:223-343 (selection rule :265-302)
//   depthAcquisition        obj_segmentation.cpp:251-316
// Observable quirks are reproduced literally, not fixed.
// PINNED CHOICE: the cluster centroid sums (cluster…:87-90) are exact sums rounded once to float
// (orc_math.h DD), like every other O(n) reduction of the oracle.
#include <cfloat>

#include "oracle.h"
#include "orc_grid.h"
#include "orc_sac.h"

namespace orc {
void radiusComponents(const float* xyz4, int n, double tol, std::vector<int>& comp, std::vector<std::vector<int>>& comps);
}
using namespace orc;

// ------------------------------------------------------------------ supports
namespace {

struct SupportCfg {
  float minCloudPct, minPlanePct, maxVar, minVar, thr, w;
  int maxIter;
  float axis[3], offset[3];
};

SupportCfg resolve(const pitt_support_params& p) {
  SupportCfg c;
  // srvm::getServiceFloatParameter: input >= 0 ? input : default (srv_manager.h:163-172)
  c.minCloudPct = p.min_iterative_cloud_percentual_size >= 0.0f ? p.min_iterative_cloud_percentual_size : 0.030f;
  c.minPlanePct = p.min_iterative_plane_percentual_size >= 0.0f ? p.min_iterative_plane_percentual_size : 0.030f;
  c.maxVar = p.variance_threshold_for_horizontal >= 0.0f ? p.variance_threshold_for_horizontal : 0.09f;
  c.minVar = -1 * c.maxVar;
  c.thr = p.ransac_distance_point_in_shape_threshold >= 0.0f ? p.ransac_distance_point_in_shape_threshold : 0.02f;
  c.w = p.ransac_model_normal_distance_weigth >= 0.0f ? p.ransac_model_normal_distance_weigth : 0.9f;
  c.maxIter = p.ransac_max_iteration_threshold >= 0 ? p.ransac_max_iteration_threshold : 10;
  const float defAxis[3] = {0.0f, 0.0f, -1.0f};
  const float defOff[3] = {0.02f, 0.02f, 0.005f};
  for (int i = 0; i < 3; ++i) {
    c.axis[i] = p.horizontal_axis_len == 3 ? p.horizontal_axis[i] : defAxis[i];
    c.offset[i] = p.support_edge_remove_offset_len == 3 ? p.support_edge_remove_offset[i] : defOff[i];
  }
  return c;
}

// isHorizontalPlane (supports…:161-179); the normals argument is unused in the reference
bool isHorizontalPlane(const float* co, const SupportCfg& c) {
  float div = sqrtf(co[0] * co[0] + co[1] * co[1] + co[2] * co[2]);
  float nx = co[0] / div, ny = co[1] / div, nz = co[2] / div;
  float crossX = ny * c.axis[2] - nz * c.axis[1];
  float crossY = nz * c.axis[0] - nx * c.axis[2];
  float crossZ = nx * c.axis[1] - ny * c.axis[0];
  return ((crossX > c.minVar) && (crossX < c.maxVar)) && ((crossY > c.minVar) && (crossY < c.maxVar)) &&
         ((crossZ > c.minVar) && (crossZ < c.maxVar));
}

// createNewIdxMap (supports…:139-157) with valueBelongsToArray as a flag lookup
std::vector<int> createNewIdxMap(const std::vector<int>& prev, const std::vector<int>& inliers, int level) {
  int cnt = 0;
  int mx = -1;
  for (int v : inliers) mx = std::max(mx, v);
  std::vector<char> flag((size_t)mx + 1, 0);
  for (int v : inliers) flag[v] = 1;
  std::vector<int> out(prev.size());
  for (size_t p = 0; p < prev.size(); ++p) {
    int v = prev[p];
    if (v > level && v < 0) out[p] = v;
    else if (v >= 0 && v <= mx && flag[v]) out[p] = level;
    else out[p] = cnt++;
  }
  return out;
}

// getPointOnPlane (supports…:187-238): order dependent if / else-if bounding box (C.7), double mean z
std::vector<int> pointsOnPlane(const float* orig, int n0, const std::vector<float>& plane, const std::vector<int>& map,
                               int level, const SupportCfg& c) {
  const double inf = std::numeric_limits<double>::infinity();
  double xMax = -inf, yMax = -inf, zMed = 0, xMin = inf, yMin = inf;
  const size_t m = plane.size() / 4;
  for (size_t i = 0; i < m; ++i) {
    float x = plane[4 * i], y = plane[4 * i + 1], z = plane[4 * i + 2];
    if (x > xMax) xMax = x;
    else if (x < xMin) xMin = x;
    if (y > yMax) yMax = y;
    else if (y < yMin) yMin = y;
    zMed += z;
  }
  xMax -= c.offset[0];
  xMin += c.offset[0];
  yMax -= c.offset[1];
  yMin += c.offset[1];
  zMed = zMed / m + c.offset[2];
  std::vector<int> out;
  for (int i = 0; i < n0; ++i) {
    if (map[i] == level) continue;  // valueBelongsToArray(i, removingIdx)
    float x = orig[4 * i], y = orig[4 * i + 1], z = orig[4 * i + 2];
    if (x > xMin && x < xMax && z > zMed && y > yMin && y < yMax) out.push_back(i);
  }
  return out;
}

}  // namespace

extern "C" int orc_find_supports(const float* xyz4, const float* nrm4, int n0, const pitt_support_params* params,
                                 pitt_support_result* res) {
  const SupportCfg c = resolve(*params);
  res->n_supports = 0;
  res->loop_trips = 0;
  res->maps_used = 0;
  res->points_used = 0;
  res->used_min_iterative_cloud_percentual_size = c.minCloudPct;
  res->used_min_iterative_plane_percentual_size = c.minPlanePct;
  res->used_max_variance_threshold_for_horizontal = c.maxVar;
  res->used_min_variance_threshold_for_horizontal = c.minVar;
  res->used_ransac_max_iteration_threshold = c.maxIter;
  res->used_ransac_distance_point_in_shape_threshold = c.thr;
  res->used_ransac_model_normal_distance_weigth = c.w;
  for (int i = 0; i < 3; ++i) {
    res->used_horizontal_axis[i] = c.axis[i];
    res->used_support_edge_remove_offset[i] = c.offset[i];
  }
  pitt_sac_params sp;
  orc_default_support_sac_params(&sp);
  sp.distance_threshold = (double)c.thr;
  sp.normal_distance_weight = (double)c.w;
  sp.max_iterations = c.maxIter;

  std::vector<float> iter(xyz4, xyz4 + (size_t)n0 * 4);
  std::vector<int> newMap;
  std::vector<int> inl((size_t)std::max(n0, 1));
  int idxMapLayer = -2, cnt = 0, status = PITT_OK;
  const int k = params->normals_k > 0 ? params->normals_k : 50;
  std::vector<float> scratchN;
  (void)nrm4;
  while (true) {
    const int ni = (int)(iter.size() / 4);
    int n_inl = 0, n_co = 0;
    float co[8];
    orc_sac_segment(iter.data(), nullptr, ni, &sp, inl.data(), (int)inl.size(), &n_inl, co, &n_co, nullptr);
    res->loop_trips++;
    if (n_inl == 0) break;
    else if ((float)ni < (float)n0 * c.minCloudPct) break;
    else if ((float)n_inl < (float)n0 * c.minPlanePct) break;
    std::vector<int> inliersIdx;
    if (!cnt) {
      inliersIdx.resize(n0);
      for (int i = 0; i < n0; ++i) inliersIdx[i] = i;
    } else {
      inliersIdx = newMap;
    }
    // removePlaneInliner: ExtractIndices positive -> support cloud, negative in place
    std::vector<float> support((size_t)n_inl * 4), rest((size_t)(ni - n_inl) * 4);
    {
      std::vector<char> flag(ni, 0);
      for (int i = 0; i < n_inl; ++i) flag[inl[i]] = 1;
      size_t a = 0, b = 0;
      for (int i = 0; i < ni; ++i) {
        float* dst = flag[i] ? &support[4 * a++] : &rest[4 * b++];
        memcpy(dst, &iter[4 * (size_t)i], 16);
      }
    }
    iter.swap(rest);
    if (params->compute_discarded_normals) {
      // supports…:297,300: computed by the reference, never used
      const float vp[3] = {0, 0, 0};
      scratchN.resize(std::max(iter.size(), support.size()));
      orc_estimate_normals(iter.data(), (int)(iter.size() / 4), k, vp, scratchN.data());
      orc_estimate_normals(support.data(), n_inl, k, vp, scratchN.data());
    }
    std::vector<int> inlVec(inl.begin(), inl.begin() + n_inl);
    if (isHorizontalPlane(co, c)) {
      newMap = createNewIdxMap(inliersIdx, inlVec, idxMapLayer);
      std::vector<int> on = pointsOnPlane(xyz4, n0, support, newMap, idxMapLayer, c);
      const int s = res->n_supports;
      const int64_t need_pts = (int64_t)n_inl + (int64_t)on.size();
      if (s < res->supports_cap && res->maps_used + n0 <= res->maps_cap && res->points_used + need_pts <= res->points_cap) {
        pitt_support& S = res->supports[s];
        S.n_map = n0;
        S.n_support = n_inl;
        S.n_on_support = (int)on.size();
        S.a = co[0]; S.b = co[1]; S.c = co[2]; S.d = co[3];
        S.map_offset = res->maps_used;
        S.support_offset = res->points_used;
        S.on_support_offset = res->points_used + n_inl;
        memcpy(res->maps + res->maps_used, newMap.data(), (size_t)n0 * 4);
        memcpy(res->points + 4 * res->points_used, support.data(), (size_t)n_inl * 16);
        for (size_t i = 0; i < on.size(); ++i)
          memcpy(res->points + 4 * (res->points_used + n_inl + (int64_t)i), xyz4 + 4 * (size_t)on[i], 16);
      } else {
        status = PITT_ERR_CAPACITY;
      }
      res->maps_used += n0;
      res->points_used += need_pts;
      res->n_supports++;
    } else {
      newMap = createNewIdxMap(inliersIdx, inlVec, -1);
    }
    cnt++;
    idxMapLayer--;
  }
  return status;
}

// ------------------------------------------------------------------ clusters
extern "C" int orc_cluster_service(const float* xyz4, int n, const pitt_cluster_params* p, pitt_clusters_result* res) {
  res->n_clusters = 0;
  res->indices_used = 0;
  if (!(n >= p->min_input_size)) return PITT_OK;
  const int min_sz = (int)round((double)n * p->min_rate);
  const int max_sz = (int)round((double)n * p->max_rate);
  std::vector<int> labels(n);
  int nc = 0;
  orc_euclidean_clusters(xyz4, n, p->tolerance, min_sz, max_sz, labels.data(), &nc);
  std::vector<std::vector<int>> lists(nc);
  for (int i = 0; i < n; ++i)
    if (labels[i] >= 0) lists[labels[i]].push_back(i);
  int status = PITT_OK;
  for (int cidx = 0; cidx < nc; ++cidx) {
    const std::vector<int>& L = lists[cidx];
    DD sx, sy, sz;
    for (int i : L) { sx.add(xyz4[4 * (size_t)i]); sy.add(xyz4[4 * (size_t)i + 1]); sz.add(xyz4[4 * (size_t)i + 2]); }
    int cnt = 1 + (int)L.size();  // `int cnt = 1; ... cnt++` (cluster…:78,91): divides by n+1
    if (cidx < res->clusters_cap && res->indices_used + (int)L.size() <= res->indices_cap) {
      pitt_cluster& C = res->clusters[cidx];
      C.n = (int)L.size();
      C.offset = res->indices_used;
      C.x_centroid = sx.f() / cnt;
      C.y_centroid = sy.f() / cnt;
      C.z_centroid = sz.f() / cnt;
      memcpy(res->indices + res->indices_used, L.data(), L.size() * 4);
    } else {
      status = PITT_ERR_CAPACITY;
    }
    res->indices_used += (int)L.size();
    res->n_clusters++;
  }
  return status;
}

// ------------------------------------------------------------------ primitives
// axis extent of cylinder/cone services: project ALL cloud points on the axis, literal O(n^2)
// farthest pair with strict '>' (cylinder…:143-171, cone…:143-171)
static void axisExtent(const float* xyz4, int n, const float* co, float* height, int* idx1, int* idx2, float* dirn,
                       std::vector<float>& proj) {
  float norm = sqrtf(co[3] * co[3] + co[4] * co[4] + co[5] * co[5]);
  dirn[0] = co[3] / norm; dirn[1] = co[4] / norm; dirn[2] = co[5] / norm;
  const float t1 = -1.0f, t2 = +1.0f;
  float A1[3] = {co[0] + dirn[0] * t1, co[1] + dirn[1] * t1, co[2] + dirn[2] * t1};
  float A2[3] = {co[0] + dirn[0] * t2, co[1] + dirn[1] * t2, co[2] + dirn[2] * t2};
  float A1A2[3] = {A2[0] - A1[0], A2[1] - A1[1], A2[2] - A1[2]};
  float gDivis = A1A2[0] * A1A2[0] + A1A2[1] * A1A2[1] + A1A2[2] * A1A2[2];
  proj.resize((size_t)n * 3);
  for (int i = 0; i < n; ++i) {
    float A1P[3] = {xyz4[4 * (size_t)i] - A1[0], xyz4[4 * (size_t)i + 1] - A1[1], xyz4[4 * (size_t)i + 2] - A1[2]};
    float G = (A1P[0] * A1A2[0] + A1P[1] * A1A2[1] + A1P[2] * A1A2[2]) / gDivis;
    proj[3 * (size_t)i] = A1[0] + G * A1A2[0];
    proj[3 * (size_t)i + 1] = A1[1] + G * A1A2[1];
    proj[3 * (size_t)i + 2] = A1[2] + G * A1A2[2];
  }
  *height = -1.0f;
  *idx1 = -1;
  *idx2 = -1;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) {
      float dx = proj[3 * (size_t)i] - proj[3 * (size_t)j], dy = proj[3 * (size_t)i + 1] - proj[3 * (size_t)j + 1],
            dz = proj[3 * (size_t)i + 2] - proj[3 * (size_t)j + 2];
      float d = sqrtf(dx * dx + dy * dy + dz * dz);
      if (d > *height) { *height = d; *idx1 = i; *idx2 = j; }
    }
}

extern "C" int orc_primitive_service(const float* xyz4, const float* nrm4, int n, const pitt_sac_params* p,
                                     pitt_primitive_result* res) {
  std::vector<int> inl((size_t)std::max(n, 1));
  int n_inl = 0, n_co = 0;
  float co[8] = {0};
  int st = orc_sac_segment(xyz4, nrm4, n, p, inl.data(), (int)inl.size(), &n_inl, co, &n_co, &res->info);
  if (st != PITT_OK) return st;
  res->n_coefficients = 0;
  res->x_centroid = res->y_centroid = res->z_centroid = 0.0f;
  res->centroid_valid = 0;
  for (int i = 0; i < 8; ++i) res->coefficients[i] = 0.0f;
  for (int i = 0; i < n_co; ++i) res->coefficients[i] = co[i];
  res->n_coefficients = n_co;
  if (p->model == PITT_MODEL_SPHERE) {
    if (n_co > 0) { res->x_centroid = co[0]; res->y_centroid = co[1]; res->z_centroid = co[2]; res->centroid_valid = 1; }
  } else if (p->model == PITT_MODEL_CYLINDER || p->model == PITT_MODEL_CONE) {
    float height = -1.0f;
    if (n_inl > 0) {
      int i1, i2;
      float dirn[3];
      std::vector<float> proj;
      axisExtent(xyz4, n, co, &height, &i1, &i2, dirn, proj);
      if (p->model == PITT_MODEL_CYLINDER) {
        if (i1 >= 0) {
          res->x_centroid = (proj[3 * (size_t)i1] + proj[3 * (size_t)i2]) / 2;
          res->y_centroid = (proj[3 * (size_t)i1 + 1] + proj[3 * (size_t)i2 + 1]) / 2;
          res->z_centroid = (proj[3 * (size_t)i1 + 2] + proj[3 * (size_t)i2 + 2]) / 2;
          res->centroid_valid = 1;
        }
      } else {
        res->x_centroid = co[0] + 3.0f / 4.0f * height * dirn[0];
        res->y_centroid = co[1] + 3.0f / 4.0f * height * dirn[1];
        res->z_centroid = co[2] + 3.0f / 4.0f * height * dirn[2];
        res->centroid_valid = 1;
      }
    }
    res->coefficients[n_co] = height;  // coefficientVector.push_back(height)
    res->n_coefficients = n_co + 1;
  }
  // PCManager::inlierToVectorMsg drops every inlier whose index VALUE is 0 (pc_manager.cpp:108)
  int m = 0;
  int status = PITT_OK;
  for (int i = 0; i < n_inl; ++i) {
    if (inl[i] == 0) continue;
    if (res->inliers && m < res->inliers_cap) res->inliers[m] = inl[i];
    else if (res->inliers) status = PITT_ERR_CAPACITY;
    ++m;
  }
  res->n_inliers = m;
  return status;
}

// selection rule of clustersAcquisition (ransac_segmentation.cpp:265-302)
extern "C" int orc_select_primitive(int64_t planeInl, int64_t sphereInl, int64_t cylinderInl, int64_t coneInl, float prio) {
  if ((!planeInl) && (!sphereInl) && (!cylinderInl) && (!coneInl)) return PITT_TAG_UNKNOWN;
  // size_t * float -> float comparison
  if ((coneInl >= planeInl) && (coneInl >= sphereInl) && ((float)(uint64_t)coneInl >= (float)(uint64_t)cylinderInl * prio))
    return PITT_TAG_CONE;
  if ((cylinderInl >= planeInl) && (cylinderInl >= coneInl) && (cylinderInl >= sphereInl)) return PITT_TAG_CYLINDER;
  if ((planeInl >= coneInl) && (planeInl >= sphereInl) && (planeInl >= cylinderInl)) return PITT_TAG_PLANE;
  if ((sphereInl >= planeInl) && (sphereInl >= coneInl) && (sphereInl >= cylinderInl)) return PITT_TAG_SPHERE;
  return PITT_TAG_UNKNOWN;
}

// ------------------------------------------------------------------ frame
extern "C" int orc_segment_frame(const float* xyz4, int n, const pitt_frame_params* fp, pitt_frame_result* res) {
  res->n_supports = res->n_clusters = res->n_shapes = 0;
  memset(res->support_coefficients, 0, sizeof(res->support_coefficients));
  memset(res->support_sizes, 0, sizeof(res->support_sizes));
  memset(res->on_support_sizes, 0, sizeof(res->on_support_sizes));
  if (!(n > fp->min_points)) return PITT_OK;  // obj_segmentation.cpp:251
  std::vector<float> nrm((size_t)n * 4);
  orc_estimate_normals(xyz4, n, fp->normals_k, fp->viewpoint, nrm.data());
  // supports
  std::vector<pitt_support> sup(8);
  std::vector<int> maps((size_t)8 * n);
  std::vector<float> pts((size_t)8 * 2 * n * 4);
  pitt_support_result sr;
  memset(&sr, 0, sizeof(sr));
  sr.supports = sup.data(); sr.supports_cap = 8;
  sr.maps = maps.data(); sr.maps_cap = (int64_t)maps.size();
  sr.points = pts.data(); sr.points_cap = (int64_t)pts.size() / 4;
  int st = orc_find_supports(xyz4, nrm.data(), n, &fp->support, &sr);
  if (st != PITT_OK) return st;
  res->n_supports = sr.n_supports;
  int status = PITT_OK;
  for (int s = 0; s < sr.n_supports; ++s) {
    const pitt_support& S = sup[s];
    if (s < 8) {
      res->support_coefficients[4 * s] = S.a; res->support_coefficients[4 * s + 1] = S.b;
      res->support_coefficients[4 * s + 2] = S.c; res->support_coefficients[4 * s + 3] = S.d;
      res->support_sizes[s] = S.n_support;
      res->on_support_sizes[s] = S.n_on_support;
    }
    const float* on = pts.data() + 4 * S.on_support_offset;
    const int non = S.n_on_support;
    std::vector<pitt_cluster> cl(256);
    std::vector<int> cidx((size_t)std::max(non, 1));
    pitt_clusters_result cr;
    memset(&cr, 0, sizeof(cr));
    cr.clusters = cl.data(); cr.clusters_cap = 256;
    cr.indices = cidx.data(); cr.indices_cap = (int)cidx.size();
    st = orc_cluster_service(on, non, &fp->cluster, &cr);
    if (st != PITT_OK) return st;
    for (int c = 0; c < cr.n_clusters; ++c) {
      const pitt_cluster& C = cl[c];
      std::vector<float> cx((size_t)C.n * 4), cn((size_t)C.n * 4);
      for (int i = 0; i < C.n; ++i) memcpy(&cx[4 * (size_t)i], on + 4 * (size_t)cidx[C.offset + i], 16);
      orc_estimate_normals(cx.data(), C.n, fp->normals_k, fp->viewpoint, cn.data());
      pitt_primitive_result pr[4];
      const pitt_sac_params* sp[4] = {&fp->sphere, &fp->cylinder, &fp->cone, &fp->plane};
      for (int m = 0; m < 4; ++m) {
        memset(&pr[m], 0, sizeof(pr[m]));
        st = orc_primitive_service(cx.data(), cn.data(), C.n, sp[m], &pr[m]);
        if (st != PITT_OK) return st;
      }
      const int64_t sphereInl = pr[0].n_inliers, cylinderInl = pr[1].n_inliers, coneInl = pr[2].n_inliers, planeInl = pr[3].n_inliers;
      int tag = orc_select_primitive(planeInl, sphereInl, cylinderInl, coneInl, fp->cone_over_cylinder_priority);
      if (res->n_shapes < res->shapes_cap) {
        pitt_tracked_shape& T = res->shapes[res->n_shapes];
        memset(&T, 0, sizeof(T));
        T.object_id = res->n_clusters;
        T.shape_tag = tag;
        T.x_pc_centroid = C.x_centroid; T.y_pc_centroid = C.y_centroid; T.z_pc_centroid = C.z_centroid;
        const pitt_primitive_result* sel = tag == PITT_TAG_CONE ? &pr[2] : tag == PITT_TAG_CYLINDER ? &pr[1]
                                           : tag == PITT_TAG_PLANE ? &pr[3] : tag == PITT_TAG_SPHERE ? &pr[0] : nullptr;
        if (sel) {
          T.x_est_centroid = sel->x_centroid; T.y_est_centroid = sel->y_centroid; T.z_est_centroid = sel->z_centroid;
          T.n_coefficients = sel->n_coefficients;
          for (int i = 0; i < 8; ++i) T.coefficients[i] = sel->coefficients[i];
        }
        T.n_points = C.n;
        T.inl_plane = (int)planeInl; T.inl_sphere = (int)sphereInl; T.inl_cylinder = (int)cylinderInl; T.inl_cone = (int)coneInl;
      } else {
        status = PITT_ERR_CAPACITY;
      }
      res->n_shapes++;
      res->n_clusters++;
    }
  }
  return status;
}
