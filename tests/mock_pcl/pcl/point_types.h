#pragma once
namespace pcl { struct PointXYZ { float x, y, z, pad; }; }
