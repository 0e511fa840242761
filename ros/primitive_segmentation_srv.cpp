// primitive_segmentation_srv.cpp — the four RANSAC primitive services on libpitt_b200.so.
// One source, four nodes: build with -DPITT_PRIMITIVE=PLANE | SPHERE | CYLINDER | CONE. Replaces the bodies of
//   ransacPlaneDetaction     segmentation_services/plane_segmentation_srv.cpp:27-74
//   ransacSphereDetection    segmentation_services/sphere_segmentation_srv.cpp:29-96
//   ransacCylinderDetaction  segmentation_services/cylinder_segmentation_srv.cpp:82-216
//   ransacConeDetaction      segmentation_services/cone_segmentation_srv.cpp:83-216
// Request/response types, service names, ROS parameter names and the "always return true" convention are the
// reference's; the PCL objects (SACSegmentation(FromNormals), the axis-extent loops, inlierToVectorMsg) are one
// call to pitt_primitive_service.
#include <cmath>

#include "pitt_ros_glue.h"
#include "pitt_msgs/PrimitiveSegmentation.h"
#include "point_cloud_library/srv_manager.h"

#define PITT_CAT_(a, b) a##b
#define PITT_CAT(a, b) PITT_CAT_(a, b)
#ifndef PITT_PRIMITIVE
#define PITT_PRIMITIVE CYLINDER
#endif
#define PITT_MODEL_ID PITT_CAT(PITT_MODEL_, PITT_PRIMITIVE)
#define PITT_PARAM(suffix) PITT_CAT(PITT_CAT(srvm::PARAM_NAME_, PITT_PRIMITIVE), suffix)
#define PITT_SRV_NAME PITT_CAT(PITT_CAT(srvm::SRV_NAME_RANSAC_, PITT_PRIMITIVE), _FILTER)

using namespace pitt_msgs;
static ros::NodeHandle* nh_ptr = NULL;

static bool ransacPrimitiveDetection(PrimitiveSegmentation::Request& req, PrimitiveSegmentation::Response& res) {
  // launch parameters: the reference's constants are the defaults, the parameter server overrides them
  pitt_sac_params p;
  pitt_default_sac_params(PITT_MODEL_ID, &p);
  double minDeg = p.min_angle * 180.0 / M_PI, maxDeg = p.max_angle * 180.0 / M_PI;
  nh_ptr->param(PITT_PARAM(_NORMAL_DISTANCE_WEIGHT), p.normal_distance_weight, p.normal_distance_weight);
  nh_ptr->param(PITT_PARAM(_DISTANCE_TH), p.distance_threshold, p.distance_threshold);
  nh_ptr->param(PITT_PARAM(_MAX_ITERATION_LIMIT), p.max_iterations, p.max_iterations);
  nh_ptr->param(PITT_PARAM(_EPS_ANGLE_TH), p.eps_angle, p.eps_angle);
  nh_ptr->param(PITT_PARAM(_MIN_OPENING_ANGLE_DEGREE), minDeg, minDeg);
  nh_ptr->param(PITT_PARAM(_MAX_OPENING_ANGLE_DEGREE), maxDeg, maxDeg);
  p.min_angle = minDeg / 180.0 * M_PI;
  p.max_angle = maxDeg / 180.0 * M_PI;
#if PITT_MODEL_ID != PITT_MODEL_PLANE
  nh_ptr->param(PITT_PARAM(_MIN_RADIUS_LIMIT), p.radius_min, p.radius_min);
  nh_ptr->param(PITT_PARAM(_MAX_RADIUS_LIMIT), p.radius_max, p.radius_max);
#endif

  pitt_cloud* cloud = pitt_ros::stage(req.cloud, &req.normals);
  if (!cloud) return true;  // empty response, like a failed segmentation
  std::vector<int32_t> inliers((size_t)pitt_cloud_size(cloud) + 1);
  pitt_primitive_result r;
  memset(&r, 0, sizeof(r));
  r.inliers = &inliers[0];
  r.inliers_cap = (int)inliers.size();
  if (pitt_primitive_service(pitt_ros::ctx(), cloud, &p, &r) != PITT_OK) ROS_ERROR("pitt_b200: %s\n", pitt_last_error(pitt_ros::ctx()));
  pitt_release_cloud(pitt_ros::ctx(), cloud);

  res.inliers.assign(inliers.begin(), inliers.begin() + r.n_inliers);  // index value 0 already dropped (pc_manager.cpp:108)
  res.coefficients.assign(r.coefficients, r.coefficients + r.n_coefficients);
  if (r.centroid_valid) {
    res.x_centroid = r.x_centroid;
    res.y_centroid = r.y_centroid;
    res.z_centroid = r.z_centroid;
  }
  return true;
}

int main(int argc, char** argv) {
  ros::init(argc, argv, PITT_SRV_NAME);
  ros::NodeHandle nh;
  nh_ptr = &nh;
  if (!pitt_ros::start()) return 1;
  ros::ServiceServer service = nh.advertiseService(PITT_SRV_NAME, ransacPrimitiveDetection);
  (void)service;
  while (nh.ok()) ros::spinOnce();
  pitt_destroy(pitt_ros::ctx());
  return 0;
}
