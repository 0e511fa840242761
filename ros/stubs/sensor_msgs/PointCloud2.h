// Stand-in for sensor_msgs/PointCloud2 (fields used by the shims only).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
namespace sensor_msgs {
struct PointField {
  std::string name;
  uint32_t offset;
  uint8_t datatype;
  uint32_t count;
  enum { FLOAT32 = 7 };
};
struct PointCloud2 {
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = false;
};
}  // namespace sensor_msgs
