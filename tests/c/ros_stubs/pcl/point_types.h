#pragma once
namespace pcl { struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; }; }
