#pragma once
#include <string>
#include <ros/time.h>
namespace tf {
struct Vector3 { Vector3(double, double, double) {} };
struct Quaternion { Quaternion(double, double, double, double) {} };
struct Transform { void setOrigin(const Vector3&) {} void setRotation(const Quaternion&) {} };
struct StampedTransform { StampedTransform(const Transform&, const ros::Time&, const std::string&, const std::string&) {} };
struct TransformBroadcaster { void sendTransform(const StampedTransform&) {} };
}
