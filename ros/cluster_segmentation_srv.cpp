// cluster_segmentation_srv.cpp — clusterize (segmentation_services/cluster_segmentation_srv.cpp:38-108) on
// libpitt_b200.so: Euclidean cluster extraction, PCL cluster order, per-cluster cloud and sum/(n+1) centroid.
#include "pitt_ros_glue.h"
#include "pitt_msgs/ClusterSegmentation.h"
#include "point_cloud_library/srv_manager.h"

using namespace pitt_msgs;
static ros::NodeHandle* nh_ptr = NULL;

static bool clusterize(ClusterSegmentation::Request& req, ClusterSegmentation::Response& res) {
  pitt_cluster_params p;
  pitt_default_cluster_params(&p);  // 0.03 / 0.01 / 0.99 / 30, cluster…:32-35
  nh_ptr->param(srvm::PARAM_NAME_CLUSTER_TOLERANCE, p.tolerance, p.tolerance);
  nh_ptr->param(srvm::PARAM_NAME_CLUSTER_MIN_RATE, p.min_rate, p.min_rate);
  nh_ptr->param(srvm::PARAM_NAME_CLUSTER_MAX_RATE, p.max_rate, p.max_rate);
  // the reference reads the minimum input size through the *tolerance* parameter name (cluster…:51-52), which
  // yields the default whenever that parameter is a double; kept as is
  nh_ptr->param(srvm::PARAM_NAME_CLUSTER_TOLERANCE, p.min_input_size, p.min_input_size);

  pitt_cloud* cloud = pitt_ros::stage(req.cloud);
  if (!cloud) return true;
  const int n = pitt_cloud_size(cloud);
  std::vector<pitt_cluster> cl(256);
  std::vector<int32_t> idx((size_t)n + 1);
  pitt_clusters_result r;
  memset(&r, 0, sizeof(r));
  r.clusters = &cl[0]; r.clusters_cap = (int)cl.size();
  r.indices = &idx[0]; r.indices_cap = (int)idx.size();
  if (pitt_cluster_service(pitt_ros::ctx(), cloud, &p, &r) != PITT_OK) ROS_ERROR("pitt_b200: %s\n", pitt_last_error(pitt_ros::ctx()));
  pitt_release_cloud(pitt_ros::ctx(), cloud);

  const float* xyz = reinterpret_cast<const float*>(req.cloud.data.empty() ? NULL : &req.cloud.data[0]);
  const size_t stride = req.cloud.point_step / 4;
  for (int c = 0; c < r.n_clusters && c < (int)cl.size(); ++c) {
    InliersCluster out;
    out.inliers.assign(idx.begin() + cl[c].offset, idx.begin() + cl[c].offset + cl[c].n);
    std::vector<float> pts((size_t)cl[c].n * 4);
    for (int i = 0; i < cl[c].n; ++i) {
      const float* q = xyz + (size_t)out.inliers[i] * stride;
      pts[4 * i] = q[0]; pts[4 * i + 1] = q[1]; pts[4 * i + 2] = q[2]; pts[4 * i + 3] = 1.0f;
    }
    out.cloud = pitt_ros::to_msg(pts.empty() ? NULL : &pts[0], cl[c].n);
    out.x_centroid = cl[c].x_centroid; out.y_centroid = cl[c].y_centroid; out.z_centroid = cl[c].z_centroid;
    res.cluster_objs.push_back(out);
  }
  return true;
}

int main(int argc, char** argv) {
  ros::init(argc, argv, srvm::SRV_NAME_CUSTER_FILTER);
  ros::NodeHandle n;
  nh_ptr = &n;
  if (!pitt_ros::start()) return 1;
  ros::ServiceServer service = n.advertiseService(srvm::SRV_NAME_CUSTER_FILTER, clusterize);
  (void)service;
  ros::spin();
  pitt_destroy(pitt_ros::ctx());
  return 0;
}
