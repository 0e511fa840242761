/* oracle.h — C entry points of the CPU ORACLE (liborc.so).
 *
 * TEST INFRASTRUCTURE ONLY. The oracle is a dependency-free CPU restatement of the PCL 1.7.x /
 * Eigen 3.2 / Boost 1.54 algorithms that the reference reaches at its seg.segment(), ne.compute(),
 * ec.extract() call sites, plus the reference's own service glue. Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product library never does.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors and its arithmetic lives
 * in un-vendored PCL that cannot be built here. What pins the oracle instead:
 * known-answer vectors of mt19937, analytic scenes with known ground truth, and property tests
 * (tests/test_oracle_*.py).
 *
 * Clouds are n x float4 {x,y,z,pad}; normals n x float4 {nx,ny,nz,curvature}. All pointers are host.
 * The parameter/result structs are the ones of include/pitt_b200.h.
 */
#ifndef PITT_ORACLE_H_
#define PITT_ORACLE_H_
#include "../include/pitt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

int orc_sac_segment(const float* xyz4, const float* nrm4, int n, const pitt_sac_params* p, int32_t* inliers, int cap,
                    int* n_inliers, float* coeffs, int* n_coeffs, pitt_sac_info* info);
int orc_sac_score(const float* xyz4, const float* nrm4, int n, const pitt_sac_params* p, const int32_t* samples, int H,
                  int32_t* counts, float* coeffs8, uint8_t* valid);
int orc_sac_select(const float* xyz4, const float* nrm4, int n, const pitt_sac_params* p, const float* coeffs,
                   int32_t* inliers, int cap, int* n_inliers);
int orc_sac_refine(const float* xyz4, const float* nrm4, int n, const pitt_sac_params* p, const float* coeffs,
                   const int32_t* inliers, int n_inliers, float* refined, pitt_sac_info* info);
int orc_pcl_sample_stream(const float* xyz4, int n, int model, int count, int32_t* out);
uint32_t orc_mt19937_nth(uint32_t seed, int nth);

int orc_knn(const float* xyz4, int n, int k, int32_t* out_idx, float* out_sqdist);
int orc_estimate_normals(const float* xyz4, int n, int k, const float* viewpoint, float* out4);
int orc_euclidean_clusters(const float* xyz4, int n, double tolerance, int min_size, int max_size, int32_t* labels,
                           int* n_clusters);

int orc_find_supports(const float* xyz4, const float* nrm4, int n, const pitt_support_params* params,
                      pitt_support_result* result);
int orc_cluster_service(const float* xyz4, int n, const pitt_cluster_params* params, pitt_clusters_result* result);
int orc_primitive_service(const float* xyz4, const float* nrm4, int n, const pitt_sac_params* params,
                          pitt_primitive_result* result);
int orc_select_primitive(int64_t plane_inl, int64_t sphere_inl, int64_t cylinder_inl, int64_t cone_inl, float prio);
int orc_segment_frame(const float* xyz4, int n, const pitt_frame_params* params, pitt_frame_result* result);

/* pre-path: VoxelGrid + deep filter + transformPointCloud; data = PointCloud2 payload (x,y,z @ 0,4,8) */
int orc_prefilter(const void* data, int point_step, int n_points, const pitt_prefilter_params* p, float* out4, int cap,
                  int* n_out, pitt_prefilter_info* info);

/* arm filter: chained negative CropBoxes (arm_filter_srv.cpp:66-103, 134-141); removed = 4 ints, nullable */
int orc_arm_filter(const float* xyz4, int n, const pitt_arm_filter_params* p, float* out4, int cap, int* n_out, int* removed);

/* defaults shared with the product header semantics (reference launch parameters) */
void orc_default_sac_params(int model, pitt_sac_params* out);
void orc_default_support_sac_params(pitt_sac_params* out);

#ifdef __cplusplus
}
#endif
#endif
