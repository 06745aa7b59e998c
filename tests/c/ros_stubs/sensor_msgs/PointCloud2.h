#pragma once
#include <memory>
#include <vector>
#include <std_msgs/Header.h>
namespace sensor_msgs {
struct PointCloud2 { std_msgs::Header header; std::vector<unsigned char> data; typedef std::shared_ptr<const PointCloud2> ConstPtr; };
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
}
