#pragma once
#include <memory>
#include <string>
#include <geometry_msgs/PoseStamped.h>
namespace nav_msgs {
struct Odometry { std_msgs::Header header; std::string child_frame_id; geometry_msgs::PoseWithCovariance pose; typedef std::shared_ptr<const Odometry> ConstPtr; };
}
