// supports_segmentation_srv.cpp — findSupports (segmentation_services/supports_segmentation_srv.cpp:241-361)
// on libpitt_b200.so: the iterative horizontal-plane removal, label maps, support and on-support clouds.
#include "pitt_ros_glue.h"
#include "pitt_msgs/SupportSegmentation.h"
#include "point_cloud_library/srv_manager.h"

using namespace pitt_msgs;

static bool findSupports(SupportSegmentation::Request& req, SupportSegmentation::Response& res) {
  pitt_support_params p;
  pitt_default_support_params(&p);  // every field "< 0 => default" exactly like srvm::getService*Parameter
  p.min_iterative_cloud_percentual_size = req.min_iterative_cloud_percentual_size;
  p.min_iterative_plane_percentual_size = req.min_iterative_plane_percentual_size;
  p.variance_threshold_for_horizontal = req.variance_threshold_for_horizontal;
  p.ransac_distance_point_in_shape_threshold = req.ransac_distance_point_in_shape_threshold;
  p.ransac_model_normal_distance_weigth = req.ransac_model_normal_distance_weigth;
  p.ransac_max_iteration_threshold = req.ransac_max_iteration_threshold;
  p.horizontal_axis_len = (int)req.horizontal_axis.size();
  for (int i = 0; i < 3 && i < p.horizontal_axis_len; ++i) p.horizontal_axis[i] = req.horizontal_axis[i];
  p.support_edge_remove_offset_len = (int)req.support_edge_remove_offset.size();
  for (int i = 0; i < 3 && i < p.support_edge_remove_offset_len; ++i) p.support_edge_remove_offset[i] = req.support_edge_remove_offset[i];

  pitt_cloud* cloud = pitt_ros::stage(req.input_cloud, &req.input_norm);
  if (!cloud) return true;
  const size_t n0 = (size_t)pitt_cloud_size(cloud);
  const int cap = 8;  // supports found per frame: 1-2 in practice
  std::vector<pitt_support> sup(cap);
  std::vector<int32_t> maps(cap * n0 + 1);
  std::vector<float> pts(cap * 2 * n0 * 4 + 4);
  pitt_support_result r;
  memset(&r, 0, sizeof(r));
  r.supports = &sup[0]; r.supports_cap = cap;
  r.maps = &maps[0]; r.maps_cap = (int64_t)maps.size();
  r.points = &pts[0]; r.points_cap = (int64_t)(pts.size() / 4);
  if (pitt_find_supports(pitt_ros::ctx(), cloud, &p, &r) != PITT_OK) ROS_ERROR("pitt_b200: %s\n", pitt_last_error(pitt_ros::ctx()));
  pitt_release_cloud(pitt_ros::ctx(), cloud);

  for (int s = 0; s < r.n_supports && s < cap; ++s) {  // Support.msg, supports…:309-328
    Support out;
    out.inliers.assign(maps.begin() + sup[s].map_offset, maps.begin() + sup[s].map_offset + sup[s].n_map);
    out.support_cloud = pitt_ros::to_msg(&pts[4 * (size_t)sup[s].support_offset], sup[s].n_support);
    out.on_support_cloud = pitt_ros::to_msg(&pts[4 * (size_t)sup[s].on_support_offset], sup[s].n_on_support);
    out.support_coefficient_a = sup[s].a; out.support_coefficient_b = sup[s].b;
    out.support_coefficient_c = sup[s].c; out.support_coefficient_d = sup[s].d;
    res.supports_description.push_back(out);
  }
  res.used_min_iterative_cloud_percentual_size = r.used_min_iterative_cloud_percentual_size;
  res.used_min_iterative_plane_percentual_size = r.used_min_iterative_plane_percentual_size;
  res.used_max_variance_threshold_for_horizontal = r.used_max_variance_threshold_for_horizontal;
  res.used_min_variance_threshold_for_horizontal = r.used_min_variance_threshold_for_horizontal;
  res.used_ransac_distance_point_in_shape_threshold = r.used_ransac_distance_point_in_shape_threshold;
  res.used_ransac_model_normal_distance_weigth = r.used_ransac_model_normal_distance_weigth;
  res.used_ransac_max_iteration_threshold = r.used_ransac_max_iteration_threshold;
  res.used_horizontal_axis.assign(r.used_horizontal_axis, r.used_horizontal_axis + 3);
  res.used_support_edge_remove_offset.assign(r.used_support_edge_remove_offset, r.used_support_edge_remove_offset + 3);
  return true;
}

int main(int argc, char** argv) {
  ros::init(argc, argv, srvm::SRV_NAME_SUPPORT_FILTER);
  ros::NodeHandle n;
  if (!pitt_ros::start()) return 1;
  ros::ServiceServer service = n.advertiseService(srvm::SRV_NAME_SUPPORT_FILTER, findSupports);
  (void)service;
  ros::spin();
  pitt_destroy(pitt_ros::ctx());
  return 0;
}
