// obj_segmentation_b200.cpp — depthAcquisition (obj_segmentation.cpp:230-323) and clustersAcquisition
// (ransac_segmentation.cpp:223-343) as ONE node on libpitt_b200.so: a frame stays in HBM from the raw message
// to the TrackedShapes instead of crossing ten ROS service hops. The arm filter (obj_segmentation.cpp:245) and
// the external geometric tracker between the two reference nodes are not part of this path.
#include "pitt_ros_glue.h"
#include "pitt_msgs/TrackedShapes.h"
#include "point_cloud_library/srv_manager.h"

using namespace pitt_msgs;
static ros::Publisher shapesPub;
static float cameraToWorld[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // filled from tf (obj_segmentation.cpp:408-434)
static const char* TAGS[] = {"unknown", "plane", "sphere", "cone", "cylinder"};

static void depthAcquisition(const sensor_msgs::PointCloud2& input) {
  pitt_prefilter_params pf;
  pitt_default_prefilter_params(&pf);  // VoxelGrid 0.01 (pc_manager.cpp:19), deep threshold 3.0 (deep_filter_srv.cpp:21)
  memcpy(pf.transform, cameraToWorld, sizeof(cameraToWorld));
  pitt_cloud* world = NULL;
  if (pitt_prefilter_cloud(pitt_ros::ctx(), input.data.empty() ? NULL : &input.data[0], (int)input.point_step,
                           (int)(input.width * input.height), &pf, &world, NULL) != PITT_OK) {
    ROS_ERROR("pitt_b200: %s\n", pitt_last_error(pitt_ros::ctx()));
    return;
  }
  pitt_frame_params fp;
  pitt_default_frame_params(&fp);  // the launch parameters of the six services; override from the parameter server here
  pitt_tracked_shape shapes[64];
  pitt_frame_result fr;
  memset(&fr, 0, sizeof(fr));
  fr.shapes = shapes;
  fr.shapes_cap = 64;
  if (pitt_segment_frame(pitt_ros::ctx(), world, &fp, &fr) != PITT_OK) ROS_ERROR("pitt_b200: %s\n", pitt_last_error(pitt_ros::ctx()));
  pitt_release_cloud(pitt_ros::ctx(), world);

  TrackedShapes out;  // ransac_segmentation.cpp:304-330
  for (int i = 0; i < fr.n_shapes && i < 64; ++i) {
    const pitt_tracked_shape& s = shapes[i];
    TrackedShape t;
    t.object_id = s.object_id;
    t.x_pc_centroid = s.x_pc_centroid; t.y_pc_centroid = s.y_pc_centroid; t.z_pc_centroid = s.z_pc_centroid;
    t.shape_tag = TAGS[(s.shape_tag >= 0 && s.shape_tag <= 4) ? s.shape_tag : 0];
    t.x_est_centroid = s.x_est_centroid; t.y_est_centroid = s.y_est_centroid; t.z_est_centroid = s.z_est_centroid;
    t.coefficients.assign(s.coefficients, s.coefficients + s.n_coefficients);
    out.tracked_shapes.push_back(t);
  }
  shapesPub.publish(out);
}

int main(int argc, char** argv) {
  ros::init(argc, argv, "obj_segmentation_b200");
  ros::NodeHandle n;
  if (!pitt_ros::start()) return 1;
  ros::Subscriber sub = n.subscribe<sensor_msgs::PointCloud2>(srvm::DEFAULT_INPUT_PARAM_RAW_CLOUD_TOPIC, 1, depthAcquisition);
  (void)sub;
  shapesPub = n.advertise<TrackedShapes>("ransac_segmentation/trackedShapes", 10);
  ros::spin();
  pitt_destroy(pitt_ros::ctx());
  return 0;
}
