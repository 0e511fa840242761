// Stand-in for the pitt_msgs package: the message / service fields the reference accesses
// (SURVEY.md Appendix A lists the evidence per field). Not part of the product.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "sensor_msgs/PointCloud2.h"
namespace pitt_msgs {
struct PrimitiveSegmentation {
  struct Request { sensor_msgs::PointCloud2 cloud, normals; };
  struct Response {
    std::vector<int32_t> inliers;
    std::vector<float> coefficients;
    float x_centroid = 0, y_centroid = 0, z_centroid = 0;
  };
};
struct InliersCluster {
  std::vector<int32_t> inliers;
  sensor_msgs::PointCloud2 cloud;
  float x_centroid = 0, y_centroid = 0, z_centroid = 0;
  int32_t shape_id = 0;
};
struct ClustersOutput { std::vector<InliersCluster> cluster_objs; };
struct ClusterSegmentation {
  struct Request { sensor_msgs::PointCloud2 cloud; };
  struct Response { std::vector<InliersCluster> cluster_objs; };
};
struct Support {
  std::vector<int32_t> inliers;
  sensor_msgs::PointCloud2 support_cloud, on_support_cloud;
  float support_coefficient_a = 0, support_coefficient_b = 0, support_coefficient_c = 0, support_coefficient_d = 0;
};
struct SupportSegmentation {
  struct Request {
    sensor_msgs::PointCloud2 input_cloud, input_norm;
    float min_iterative_cloud_percentual_size = -1, min_iterative_plane_percentual_size = -1,
          variance_threshold_for_horizontal = -1, ransac_distance_point_in_shape_threshold = -1,
          ransac_model_normal_distance_weigth = -1;
    int32_t ransac_max_iteration_threshold = -1;
    std::vector<float> horizontal_axis, support_edge_remove_offset;
  };
  struct Response {
    std::vector<Support> supports_description;
    float used_min_iterative_cloud_percentual_size = 0, used_min_iterative_plane_percentual_size = 0,
          used_max_variance_threshold_for_horizontal = 0, used_min_variance_threshold_for_horizontal = 0,
          used_ransac_distance_point_in_shape_threshold = 0, used_ransac_model_normal_distance_weigth = 0;
    int32_t used_ransac_max_iteration_threshold = 0;
    std::vector<float> used_horizontal_axis, used_support_edge_remove_offset;
  };
};
struct DeepFilter {
  struct Request { sensor_msgs::PointCloud2 input_cloud; float deep_threshold = -1; };
  struct Response { sensor_msgs::PointCloud2 cloud_closer, cloud_further; float used_deep_threshold = 0; };
};
struct TrackedShape {
  int32_t object_id = 0;
  float x_pc_centroid = 0, y_pc_centroid = 0, z_pc_centroid = 0;
  std::string shape_tag;
  float x_est_centroid = 0, y_est_centroid = 0, z_est_centroid = 0;
  std::vector<float> coefficients;
};
struct TrackedShapes { std::vector<TrackedShape> tracked_shapes; };
}  // namespace pitt_msgs
