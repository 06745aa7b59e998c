#pragma once
#include <pcl/point_cloud.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
template <typename P> void fromROSMsg(const sensor_msgs::PointCloud2&, PointCloud<P>&) {}
template <typename P> void toROSMsg(const PointCloud<P>&, sensor_msgs::PointCloud2&) {}
}
