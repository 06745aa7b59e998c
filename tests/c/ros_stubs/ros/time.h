#pragma once
namespace ros {
struct Time { double t = 0; double toSec() const { return t; } };
struct Rate { explicit Rate(double) {} bool sleep() { return true; } };
}
