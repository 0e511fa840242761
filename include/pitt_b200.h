/*
 * pitt_b200.h — C ABI of the B200-native tabletop segmentation hot path.
 *
 * Drop-in boundary for the PCL calls made by the reference package
 * TheEngineRoom-UniGe/pitt_object_table_segmentation (paths relative to /root/reference/src):
 *
 *   seg.segment(inliers, coeffs)       segmentation_services/supports_segmentation_srv.cpp:110
 *                                      segmentation_services/plane_segmentation_srv.cpp:67
 *                                      segmentation_services/sphere_segmentation_srv.cpp:73
 *                                      segmentation_services/cylinder_segmentation_srv.cpp:126
 *                                      segmentation_services/cone_segmentation_srv.cpp:127
 *   ne.compute(normals)                point_cloud_library/pc_manager.cpp:76
 *   ec.extract(cluster_indices)        segmentation_services/cluster_segmentation_srv.cpp:69
 *   extract.filter(...)                segmentation_services/supports_segmentation_srv.cpp:120,126
 *
 * and for the service callbacks that wrap them (findSupports, clusterize, ransac*Detection).
 *
 * Conventions: extern "C", plain pointers and sizes, no exceptions, no ROS/PCL/torch types.
 * Every function returns an int status (PITT_OK == 0). "No model found" is PITT_OK with a zero
 * length output, exactly like the reference callbacks that always `return true`.
 * Outputs are caller allocated (capacity + returned length) unless stated otherwise.
 * A pitt_ctx owns one CUDA stream; contexts are independent (one per host thread / per GPU).
 * There is NO CPU fallback: every entry point fails with PITT_ERR_CUDA when no device is usable.
 *
 * The same POD structs are used by the CPU oracle (oracle/, test infrastructure only).
 */
#ifndef PITT_B200_H_
#define PITT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status codes */
#define PITT_OK 0
#define PITT_ERR_INVALID 1  /* bad argument */
#define PITT_ERR_CUDA 2     /* CUDA runtime/driver error, or no device: the call did no work */
#define PITT_ERR_CAPACITY 3 /* caller buffer too small; required length is reported */
#define PITT_ERR_STATE 4    /* e.g. normals required but not present on the cloud */

/* ------------------------------------------------------------------ enums */
/* pcl::SacModel values used by the reference (supports…:93, plane…:56, sphere…:62,
 * cylinder…:115, cone…:116). Numeric values are ours. */
#define PITT_MODEL_PLANE 0
#define PITT_MODEL_SPHERE 1
#define PITT_MODEL_CYLINDER 2
#define PITT_MODEL_CONE 3

/* Hypothesis sampler. PCL_MT19937 replays pcl::SampleConsensusModel::getSamples bit for bit
 * (boost::mt19937 seeded 12345, persistent partial Fisher-Yates) on the host and is
 * the validation mode. PHILOX draws independent minimal sets on the device from a counter-based
 * generator (ctx seed, hypothesis id) and is the production mode. REPLAY takes the caller's table. */
#define PITT_SAMPLER_PCL_MT19937 0
#define PITT_SAMPLER_PHILOX 1
#define PITT_SAMPLER_REPLAY 2

/* Stop rule. PCL_ADAPTIVE reproduces RandomSampleConsensus::computeModel (best-so-far with strict >,
 * adaptive k from probability, iterations_ > max_iterations_ break, skipped <= 10*max_iterations).
 * ALL_H scores exactly max_iterations hypotheses and takes the earliest arg-max. */
#define PITT_STOP_PCL_ADAPTIVE 0
#define PITT_STOP_ALL_H 1

/* shape tags of ransac_segmentation.cpp:42-46 */
#define PITT_TAG_UNKNOWN 0
#define PITT_TAG_PLANE 1
#define PITT_TAG_SPHERE 2
#define PITT_TAG_CONE 3
#define PITT_TAG_CYLINDER 4

/* ------------------------------------------------------------------ parameter structs */

/* Mirrors the setters called on pcl::SACSegmentationFromNormals at the reference call sites. */
typedef struct pitt_sac_params {
  int32_t model;                 /* PITT_MODEL_* (setModelType) */
  int32_t max_iterations;        /* setMaxIterations */
  double distance_threshold;     /* setDistanceThreshold */
  double probability;            /* PCL default 0.99 */
  double normal_distance_weight; /* setNormalDistanceWeight (inert for plane/sphere) */
  double radius_min, radius_max; /* setRadiusLimits (sphere, cylinder; not forwarded for cones) */
  double min_angle, max_angle;   /* setMinMaxOpeningAngle, radians (cone only) */
  double eps_angle;              /* setEpsAngle (cylinder/cone; with axis 0 it never rejects) */
  float axis[3];                 /* setAxis — the reference never calls it: (0,0,0) */
  int32_t optimize;              /* setOptimizeCoefficients(true) at every call site */
  int32_t sampler;               /* PITT_SAMPLER_* */
  int32_t stop;                  /* PITT_STOP_* */
  const int32_t* replay_samples; /* PITT_SAMPLER_REPLAY: [replay_count][sample_size] host indices */
  int32_t replay_count;
  int32_t reserved;
} pitt_sac_params;

/* What the RANSAC loop did; all fields are outputs. */
typedef struct pitt_sac_info {
  int32_t iterations;       /* PCL iterations_ at exit */
  int32_t skipped;          /* degenerate / invalid samples skipped */
  int32_t hypotheses;       /* hypotheses actually scored on the device */
  int32_t best_hypothesis;  /* position in the sample stream of the winner (-1: none) */
  int32_t best_count;       /* its inlier count (before refinement) */
  int32_t n_inliers_model;  /* inliers of the un-refined winner */
  int32_t lm_info;          /* Eigen::LevenbergMarquardt status of the refinement (0 for plane) */
  int32_t lm_nfev;
  float model_coeffs[8];    /* un-refined winner */
  double device_ms;         /* device time of the whole call measured with CUDA events */
} pitt_sac_info;

/* supports_segmentation_srv.cpp:30-39,70-86. Negative scalars / non-3-vectors select defaults,
 * as srvm::getService*Parameter does (srv_manager.h:163-188). */
typedef struct pitt_support_params {
  float min_iterative_cloud_percentual_size;
  float min_iterative_plane_percentual_size;
  float variance_threshold_for_horizontal;
  float ransac_distance_point_in_shape_threshold;
  float ransac_model_normal_distance_weigth;
  int32_t ransac_max_iteration_threshold;
  int32_t horizontal_axis_len;            /* != 3 → default (0,0,-1) */
  float horizontal_axis[3];
  int32_t support_edge_remove_offset_len; /* != 3 → default (0.02,0.02,0.005) */
  float support_edge_remove_offset[3];
  int32_t normals_k;                      /* k of the per-iteration normal re-estimation (50) */
  int32_t compute_discarded_normals;      /* 1: also run the two estimateNormal calls whose result
                                             the reference discards (supports…:297,300) */
} pitt_support_params;

/* One element of SupportSegmentation::Response::supports_description (pitt_msgs/Support). */
typedef struct pitt_support {
  int32_t n_map;            /* == N0, length of the label map */
  int32_t n_support;        /* points of support_cloud */
  int32_t n_on_support;     /* points of on_support_cloud */
  float a, b, c, d;         /* support_coefficient_a..d */
  /* offsets into the flat result buffers of pitt_support_result */
  int64_t map_offset;       /* int32 labels */
  int64_t support_offset;   /* float4 points */
  int64_t on_support_offset;/* float4 points */
} pitt_support;

typedef struct pitt_support_result {
  int32_t n_supports;
  int32_t loop_trips;       /* RANSAC calls made by findSupports */
  /* used_* of the response */
  float used_min_iterative_cloud_percentual_size;
  float used_min_iterative_plane_percentual_size;
  float used_max_variance_threshold_for_horizontal;
  float used_min_variance_threshold_for_horizontal;
  int32_t used_ransac_max_iteration_threshold;
  float used_ransac_distance_point_in_shape_threshold;
  float used_ransac_model_normal_distance_weigth;
  float used_horizontal_axis[3];
  float used_support_edge_remove_offset[3];
  /* caller-allocated flat buffers + capacities (elements); needed sizes reported in *_used */
  pitt_support* supports; int32_t supports_cap;
  int32_t* maps;   int64_t maps_cap;   int64_t maps_used;
  float* points;   int64_t points_cap; int64_t points_used; /* float4 per point */
} pitt_support_result;

/* cluster_segmentation_srv.cpp:32-35 */
typedef struct pitt_cluster_params {
  double tolerance;     /* 0.03  */
  double min_rate;      /* 0.01  */
  double max_rate;      /* 0.99  */
  int32_t min_input_size; /* 30 (the reference reads it from the tolerance key) */
  int32_t reserved;
} pitt_cluster_params;

/* One pitt_msgs/InliersCluster. */
typedef struct pitt_cluster {
  int32_t n;                 /* points in the cluster */
  int32_t offset;            /* into pitt_clusters_result::indices (ascending point indices) */
  float x_centroid, y_centroid, z_centroid; /* sum/(n+1), SURVEY C.2 */
} pitt_cluster;

typedef struct pitt_clusters_result {
  int32_t n_clusters;
  pitt_cluster* clusters; int32_t clusters_cap;
  int32_t* indices; int32_t indices_cap; int32_t indices_used;
} pitt_clusters_result;

/* PrimitiveSegmentation::Response (pitt_msgs) as filled by the four ransac*Detection callbacks. */
typedef struct pitt_primitive_result {
  int32_t n_inliers;         /* after inlierToVectorMsg (index value 0 dropped) */
  int32_t n_coefficients;    /* plane 4, sphere 4, cylinder/cone 8 (height appended) or 1 ([-1]) */
  float coefficients[8];
  float x_centroid, y_centroid, z_centroid;
  int32_t centroid_valid;    /* 0 where the reference leaves the centroid uninitialised */
  int32_t* inliers; int32_t inliers_cap;
  pitt_sac_info info;
} pitt_primitive_result;

/* One pitt_msgs/TrackedShape produced by clustersAcquisition (ransac_segmentation.cpp:223-343). */
typedef struct pitt_tracked_shape {
  int32_t object_id;
  int32_t shape_tag;         /* PITT_TAG_* */
  float x_pc_centroid, y_pc_centroid, z_pc_centroid;
  float x_est_centroid, y_est_centroid, z_est_centroid;
  int32_t n_coefficients;
  float coefficients[8];
  int32_t n_points;
  int32_t inl_plane, inl_sphere, inl_cylinder, inl_cone; /* counts used by the selection rule */
} pitt_tracked_shape;

/* Per-primitive launch parameters (ROS params pitt/srv/<name>_segmentation/..., srv_manager.h:35-95). */
typedef struct pitt_frame_params {
  pitt_support_params support;
  pitt_cluster_params cluster;
  pitt_sac_params plane, sphere, cylinder, cone;
  int32_t normals_k;          /* 50, pc_manager.cpp:18 */
  int32_t min_points;         /* 30, obj_segmentation.cpp:55 (strict >) */
  float viewpoint[3];         /* (0,0,0) */
  float cone_over_cylinder_priority; /* 0.9, ransac_segmentation.cpp:37 */
} pitt_frame_params;

typedef struct pitt_frame_result {
  int32_t n_supports;
  int32_t n_clusters;
  int32_t n_shapes;
  pitt_tracked_shape* shapes; int32_t shapes_cap;
  float support_coefficients[4 * 8]; /* first 8 supports */
  int32_t support_sizes[8];
  int32_t on_support_sizes[8];
  double device_ms;
} pitt_frame_result;

/* Pre-path of depthAcquisition (obj_segmentation.cpp:233-248): fromROSMsg -> PCManager::downSampling
 * (VoxelGrid, pc_manager.cpp:55-67, leaf 0.01 :19) -> deep filter (deep_filter_srv.cpp:27-58, threshold
 * 3.0 :21) -> pcl::transformPointCloud with the camera->world matrix (:248). The arm filter (:245,
 * robot specific CropBox on live tf) is not part of it. */
typedef struct pitt_prefilter_params {
  float leaf[3];              /* VoxelGrid leaf size; any value <= 0 skips the down-sampling */
  int32_t apply_deep_filter;  /* 1: drop z != z and z > deep_threshold (the "closer" cloud of the service) */
  float deep_threshold;       /* < 0 -> 3.0 (srvm::getServiceFloatParameter) */
  int32_t apply_transform;    /* 1: x' = m00 x + m01 y + m02 z + m03, ... (float, unfused, left to right) */
  float transform[16];        /* row major 4x4 */
} pitt_prefilter_params;

typedef struct pitt_prefilter_info {
  int32_t n_input;        /* points in the message */
  int32_t n_voxel;        /* after the VoxelGrid (= n_input when skipped) */
  int32_t n_closer;       /* deep filter: kept */
  int32_t n_further;      /* deep filter: z > threshold */
  int32_t voxel_overflow; /* 1: leaf too small for the extent (PCL warns and returns the input unchanged) */
  float used_deep_threshold;
} pitt_prefilter_info;

/* the arm filter service (segmentation_services/arm_filter_srv.cpp). armFiltering (:66-103) runs a
 * pcl::CropBox with setNegative(true) per arm link: min/max corner in the link frame (generateBoxVector :105-109, defaults
 * :32-35), translation = origin of the tf frame (:78-80), rotation = its roll/pitch/yaw (:73, :83-85); the service
 * chains four of them (left/right forearm, left/right elbow :134-141). The tf look-ups stay with the caller. */
typedef struct pitt_crop_box {
  float min_pt[3], max_pt[3];
  float translation[3];
  float rotation_rpy[3];
} pitt_crop_box;
typedef struct pitt_arm_filter_params {
  int32_t n_boxes;         /* 0..4; 0 = the service's tfError branch: output = input */
  int32_t input_is_dense;  /* PointCloud::is_dense of the request: 0 drops non-finite points before the box test */
  pitt_crop_box box[4];
} pitt_arm_filter_params;

/* ------------------------------------------------------------------ opaque handles */
typedef struct pitt_ctx pitt_ctx;
typedef struct pitt_cloud pitt_cloud; /* cloud staged in HBM as float4 {x,y,z,1} (+ normals float4) */

/* ------------------------------------------------------------------ lifecycle */
pitt_ctx* pitt_create(int device, uint64_t seed);
/* Same, but every kernel of this ctx is launched on the caller's cudaStream_t (e.g. torch's
 * current stream, so that torch.cuda.Event brackets the work). */
pitt_ctx* pitt_create_on_stream(int device, uint64_t seed, void* cuda_stream);
void pitt_destroy(pitt_ctx* ctx);
/* pitt_segment_frame runs the independent primitive fits of a frame (clusters x {sphere, cylinder,
 * cone, plane}; the reference calls the four services one after the other, ransac_segmentation.cpp:
 * 235-262) concurrently on n_workers helper streams, one host thread each. 0 = all on the ctx stream.
 * Default 4. Results do not depend on it. */
int pitt_set_workers(pitt_ctx* ctx, int n_workers);
const char* pitt_last_error(const pitt_ctx* ctx);
const char* pitt_version(void);
int pitt_device_count(void);
/* Host waits of this context sleep on a cudaEventBlockingSync event instead of spinning (PITT_BLOCKING_SYNC=1 makes it
 * the default of new contexts). For frame streams with more host threads than cores: 8 ranks x 16 contexts on a 32-core box
 * went from 1936 to 2901 frames/s; with a core per thread spinning is faster (597 vs 553 frames/s on one GPU). */
int pitt_set_blocking_sync(pitt_ctx* ctx, int enable);
int pitt_synchronize(pitt_ctx* ctx);

/* ------------------------------------------------------------------ defaults (reference launch parameters) */
void pitt_default_sac_params(int model, pitt_sac_params* out);      /* plane…:19-24 sphere…:19-26 cylinder…:23-30 cone…:24-31 */
void pitt_default_support_sac_params(pitt_sac_params* out);        /* supports…:35-37 */
void pitt_default_support_params(pitt_support_params* out);        /* all negative → defaults */
void pitt_default_cluster_params(pitt_cluster_params* out);
void pitt_default_frame_params(pitt_frame_params* out);
void pitt_default_prefilter_params(pitt_prefilter_params* out);    /* leaf 0.01, deep 3.0, identity transform */
void pitt_default_arm_filter_params(pitt_arm_filter_params* out);  /* 4 boxes with the default corners, identity frames */

/* ------------------------------------------------------------------ staging (K0; replaces fromROSMsg, pc_manager.cpp:85-94) */
/* xyz: host pointer to n points, x,y,z float32 at byte offsets 0,4,8 of each stride_bytes record
 * (12 = packed xyz, 16 = pcl::PointXYZ / PointCloud2 point_step 16). Copied once into HBM. */
int pitt_stage_cloud(pitt_ctx* ctx, const void* xyz, int stride_bytes, int n, pitt_cloud** out);
/* Wrap n float4 {x,y,z,*} already resident in device memory (copied device-to-device). */
int pitt_stage_cloud_device(pitt_ctx* ctx, const void* d_xyz4, int n, pitt_cloud** out);
/* Attach request normals (host, nx,ny,nz at 0,4,8 and curvature at byte 16 when stride is 32;
 * stride 16 = {nx,ny,nz,curvature}). */
int pitt_set_normals(pitt_ctx* ctx, pitt_cloud* cloud, const void* normals, int stride_bytes);
int pitt_cloud_size(const pitt_cloud* cloud);
int pitt_cloud_has_normals(const pitt_cloud* cloud);
/* device pointers for zero-copy consumers (float4 per point); normals may be NULL */
const void* pitt_cloud_device_points(const pitt_cloud* cloud);
const void* pitt_cloud_device_normals(const pitt_cloud* cloud);
void pitt_release_cloud(pitt_ctx* ctx, pitt_cloud* cloud);

/* ------------------------------------------------------------------ pre-path */
/* data: the PointCloud2 payload (host): n_points records of point_step bytes with x,y,z float32 at byte
 * offsets 0,4,8 (PointXYZ, PointXYZRGB, ...). One H2D copy, then everything on the device. The result
 * is a staged cloud in the world frame, ready for pitt_segment_frame. info is nullable. */
int pitt_prefilter_cloud(pitt_ctx* ctx, const void* data, int point_step, int n_points,
                         const pitt_prefilter_params* params, pitt_cloud** out, pitt_prefilter_info* info);
/* the same on a cloud that is already staged (device to device) */
int pitt_prefilter_staged(pitt_ctx* ctx, const pitt_cloud* in, const pitt_prefilter_params* params,
                          pitt_cloud** out, pitt_prefilter_info* info);
/* arm filter (CropBox x n_boxes, negative): the points outside every box, in input order. removed[k] (nullable, 4 ints) =
 * points removed by box k, as the service logs them (arm_filter_srv.cpp:135-141). */
int pitt_arm_filter(pitt_ctx* ctx, const pitt_cloud* in, const pitt_arm_filter_params* params, pitt_cloud** out,
                    int32_t* removed);
/* copy a staged cloud back as n x float4 (tests, debugging) */
int pitt_get_points(pitt_ctx* ctx, const pitt_cloud* cloud, float* out4);

/* ------------------------------------------------------------------ a1: PCManager::estimateNormal (pc_manager.cpp:68-78) */
/* k nearest neighbours (query included, ties by lower index) → 3x3 covariance → eigen33 →
 * flip towards viewpoint. Result stays on the device attached to the cloud. */
int pitt_estimate_normals(pitt_ctx* ctx, pitt_cloud* cloud, int k, const float viewpoint[3]);
/* copy out as n x {nx,ny,nz,curvature} */
int pitt_get_normals(pitt_ctx* ctx, const pitt_cloud* cloud, float* out4);
/* debugging/parity: the k neighbour indices of every point, sorted by (distance, index) */
int pitt_knn(pitt_ctx* ctx, const pitt_cloud* cloud, int k, int32_t* out_idx, float* out_sqdist);

/* ------------------------------------------------------------------ a2,a9-a12: seg.segment() */
int pitt_sac_segment(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_sac_params* params,
                     int32_t* inliers /* NULL: count only */, int inliers_cap, int* n_inliers,
                     float coeffs[8], int* n_coeffs, pitt_sac_info* info /* nullable */);
/* The same on a cloud that is still in host memory, i.e. what a service callback has in hand: fromROSMsg
 * (pc_manager.cpp:85-94) + seg.segment() (supports_segmentation_srv.cpp:110, plane_segmentation_srv.cpp:67) in one call,
 * without the intermediate synchronisations; clouds of 16 M points and more travel in chunks on a second CUDA stream and
 * are scored chunk by chunk while the rest is still being copied (plane models on point_step-16 clouds of >= 2^18 points;
 * any other input is staged, segmented and released in sequence).
 * xyz is borrowed for the call. Results are identical to pitt_stage_cloud + pitt_sac_segment. */
int pitt_sac_segment_host(pitt_ctx* ctx, const void* xyz, int stride_bytes, int n, const pitt_sac_params* params,
                          int32_t* inliers /* NULL: count only */, int inliers_cap, int* n_inliers,
                          float coeffs[8], int* n_coeffs, pitt_sac_info* info /* nullable */);

/* Parity/bench hook below segment(): score a caller supplied sample table. For hypothesis h the
 * library estimates the model from samples[h*S .. h*S+S) (computeModelCoefficients), then counts
 * inliers (countWithinDistance incl. isModelValid). valid[h]=0 where computeModelCoefficients
 * returned false. All outputs are host pointers and nullable. */
int pitt_sac_score(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_sac_params* params,
                   const int32_t* samples, int n_hypotheses,
                   int32_t* counts, float* coeffs8, uint8_t* valid);
/* Same, but samples (int32) and counts (int32) are device pointers and nothing is copied back:
 * the resident-input arm of bench.py and the multi-GPU hypothesis split (config 5). */
int pitt_sac_score_device(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_sac_params* params,
                          const void* d_samples, int n_hypotheses, void* d_counts);
/* Earliest arg-max over H device-resident int32 counts (e.g. all-gathered over NVLink from the
 * ranks of a hypothesis split); d_best (device) receives {index, count}. */
int pitt_argmax_counts_device(pitt_ctx* ctx, const void* d_counts, int n_hypotheses, void* d_best);
/* Second half of seg.segment() for a hypothesis split: d_best (device, {index, count}) names the winner
 * in the full device-resident sample table; its model is re-estimated from its sample, refined
 * (optimizeModelCoefficients) and the final inliers selected. inliers may be NULL (count only). */
int pitt_sac_finish_device(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_sac_params* params,
                           const void* d_samples_all, int n_hypotheses_all, const void* d_best,
                           int32_t* inliers, int inliers_cap, int* n_inliers, float* coeffs, int* n_coeffs,
                           pitt_sac_info* info /* nullable */);
/* SURVEY 8e / BASELINE configs[4]: seg.segment() of one cloud with the hypothesis set split over `world` ranks (one process
 * per GPU; every rank holds the cloud). params->stop must be PITT_STOP_ALL_H; params->max_iterations = hypotheses of the WHOLE
 * job (rank r scores stream positions [r*Hl, (r+1)*Hl), Hl = ceil(H / world)). `allgather` is the caller's collective (NCCL over
 * NVLink in a ROS node or under torchrun): it must gather count_per_rank int32 from d_send of every rank
 * into d_recv in rank order, enqueued on cuda_stream (or ordered after it), and return 0. With world == 1 it may be NULL.
 * Every rank returns the result pitt_sac_segment gives on one GPU. */
typedef int (*pitt_allgather_fn)(void* user, const void* d_send, void* d_recv, int count_per_rank, void* cuda_stream);
int pitt_sac_segment_split(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_sac_params* params, int rank, int world,
                           pitt_allgather_fn allgather, void* user, int32_t* inliers /* NULL: count only */, int inliers_cap,
                           int* n_inliers, float coeffs[8], int* n_coeffs, pitt_sac_info* info /* nullable */);
/* selectWithinDistance for given coefficients (ascending indices). */
int pitt_sac_select(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_sac_params* params,
                    const float* coeffs, int32_t* inliers, int inliers_cap, int* n_inliers);
/* optimizeModelCoefficients on a given inlier set. */
int pitt_sac_refine(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_sac_params* params,
                    const float* coeffs, const int32_t* inliers, int n_inliers, float* refined,
                    pitt_sac_info* info /* nullable */);
/* the sample stream getSamples would produce (host only; needs the cloud only for the plane
 * collinearity redraw). count*sample_size indices are written. */
int pitt_pcl_sample_stream(pitt_ctx* ctx, const pitt_cloud* cloud, int model, int count, int32_t* out);

/* ------------------------------------------------------------------ a8: ec.extract() */
/* labels[i] = cluster id (0 = largest, PCL order) or -1 when the component was filtered out. */
int pitt_euclidean_clusters(pitt_ctx* ctx, const pitt_cloud* cloud, double tolerance,
                            int min_size, int max_size, int32_t* labels, int* n_clusters);

/* ------------------------------------------------------------------ service-shaped entry points */
/* findSupports (supports_segmentation_srv.cpp:241-361). cloud must carry normals (the request does). */
int pitt_find_supports(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_support_params* params,
                       pitt_support_result* result);
/* clusterize (cluster_segmentation_srv.cpp:38-108) */
int pitt_cluster_service(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_cluster_params* params,
                         pitt_clusters_result* result);
/* ransac{Plane,Sphere,Cylinder,Cone}Detection (…_segmentation_srv.cpp) incl. index-0 drop,
 * axis height and centroid. params->model selects the service. */
int pitt_primitive_service(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_sac_params* params,
                           pitt_primitive_result* result);
/* selection rule of clustersAcquisition (ransac_segmentation.cpp:265-302) */
int pitt_select_primitive(int64_t plane_inl, int64_t sphere_inl, int64_t cylinder_inl,
                          int64_t cone_inl, float cone_over_cylinder_priority);
/* depthAcquisition + clustersAcquisition from the world-frame cloud on
 * (obj_segmentation.cpp:251-316, ransac_segmentation.cpp:223-343). */
int pitt_segment_frame(pitt_ctx* ctx, const pitt_cloud* cloud, const pitt_frame_params* params,
                       pitt_frame_result* result);

/* C4: a stream of frames. Thread t of n_ctx host threads drives ctxs[t] (one stream each; the
 * contexts may sit on one GPU or on several) over frames t, t+n_ctx, ...: stage from the host
 * pointer (frames[i], n_points[i], stride_bytes), pitt_segment_frame, release. results[i] must
 * have its shapes buffer set. Returns the first non-OK status. The frame buffers are borrowed until the
 * call returns: with stride_bytes == 16 (pinned memory recommended) the copy of a context's next frame is
 * queued on a second stream while its current frame is segmented. */
int pitt_segment_frames_batched(pitt_ctx* const* ctxs, int n_ctx, const void* const* frames, const int* n_points,
                                int stride_bytes, int n_frames, const pitt_frame_params* params,
                                pitt_frame_result* results);
/* The same stream from raw sensor messages: every frame first goes through pitt_prefilter_cloud
 * (prefilter may be NULL = frames are already world-frame clouds). */
int pitt_segment_raw_frames_batched(pitt_ctx* const* ctxs, int n_ctx, const void* const* frames, const int* n_points,
                                    int stride_bytes, int n_frames, const pitt_prefilter_params* prefilter,
                                    const pitt_frame_params* params, pitt_frame_result* results);

/* The same stream over clouds that are already staged in HBM (any context of the same device may have staged them):
 * the resident-input arm of bench.py, and callers that keep the sensor's frames on the device. */
int pitt_segment_clouds_batched(pitt_ctx* const* ctxs, int n_ctx, const pitt_cloud* const* clouds, int n_frames,
                                const pitt_frame_params* params, pitt_frame_result* results);

/* ------------------------------------------------------------------ measurement helpers */
/* FP32 pipe micro-benchmark used as the roofline denominator of the scoring kernels:
 * kind 0 = FFMA, 1 = FMUL+FADD (unfused, what bit-exact scoring needs), 2 = packed f32x2 FMUL2+FADD2.
 * Returns achieved TFLOP/s (FMA = 2 flop). */
int pitt_fp32_peak(pitt_ctx* ctx, int kind, double* tflops);
/* device time in ms of the most recent call on this ctx (CUDA events on the ctx stream) */
double pitt_last_device_ms(const pitt_ctx* ctx);
/* number of kernels this ctx has launched so far */
int64_t pitt_kernel_launches(const pitt_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PITT_B200_H_ */
