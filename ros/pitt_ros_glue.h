// pitt_ros_glue.h — the few lines of glue every patched node shares: one pitt_ctx per process, PointCloud2 <->
// staged cloud. Replaces PCManager::cloudForRosMsg / normForRosMsg / cloudToRosMsg (pc_manager.cpp:80-104)
// for the nodes that call into libpitt_b200.so. Host code stays C++/ROS; CUDA is only behind pitt_b200.h.
#pragma once
#include <cstring>
#include <vector>

#include "pitt_b200.h"
#include "ros/ros.h"
#include "sensor_msgs/PointCloud2.h"

namespace pitt_ros {

inline pitt_ctx*& ctx() {
  static pitt_ctx* c = NULL;
  return c;
}
// call first in main(): no GPU => refuse to start (there is no CPU fallback)
inline bool start(int device = 0) {
  ctx() = pitt_create(device, 12345u /* PCL's fixed RANSAC seed */);
  if (!ctx()) ROS_FATAL("pitt_b200: no usable CUDA device\n");
  return ctx() != NULL;
}
// PointXYZ clouds have point_step 16 (x,y,z at 0,4,8); Normal clouds point_step 32 (curvature at 16)
inline pitt_cloud* stage(const sensor_msgs::PointCloud2& c, const sensor_msgs::PointCloud2* normals = NULL) {
  pitt_cloud* pc = NULL;
  if (pitt_stage_cloud(ctx(), c.data.empty() ? NULL : &c.data[0], (int)c.point_step, (int)(c.width * c.height), &pc) != PITT_OK) {
    ROS_ERROR("pitt_b200: %s\n", pitt_last_error(ctx()));
    return NULL;
  }
  if (normals && normals->width * normals->height == c.width * c.height && !normals->data.empty())
    pitt_set_normals(ctx(), pc, &normals->data[0], (int)normals->point_step);
  return pc;
}
// n x float4 {x,y,z,1} -> unorganised PointXYZ message (what toROSMsg produces for PointCloud<PointXYZ>)
inline sensor_msgs::PointCloud2 to_msg(const float* xyz4, int n) {
  sensor_msgs::PointCloud2 m;
  m.height = 1;
  m.width = (uint32_t)n;
  m.point_step = 16;
  m.row_step = 16u * (uint32_t)n;
  m.is_dense = true;
  const char* names[3] = {"x", "y", "z"};
  for (int a = 0; a < 3; ++a) {
    sensor_msgs::PointField f;
    f.name = names[a];
    f.offset = 4u * a;
    f.datatype = sensor_msgs::PointField::FLOAT32;
    f.count = 1;
    m.fields.push_back(f);
  }
  m.data.resize((size_t)n * 16);
  if (n > 0) memcpy(&m.data[0], xyz4, (size_t)n * 16);
  return m;
}

}  // namespace pitt_ros
