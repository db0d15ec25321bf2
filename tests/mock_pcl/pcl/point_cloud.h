#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <class P> struct PointCloud {
  typedef std::shared_ptr<PointCloud<P> > Ptr;
  typedef std::shared_ptr<const PointCloud<P> > ConstPtr;
  std::vector<P> points;
  uint32_t width = 0, height = 1;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
};
}
