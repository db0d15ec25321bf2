// Compiles include/mvr_pcl_adapter.hpp against tests/mock_pcl and exercises the parts that need no GPU.
#include <cstdio>
#include "mvr_pcl_adapter.hpp"
#ifndef MVR_HAVE_PCL
#error "the adapter did not see the (stand-in) PCL headers"
#endif
int main() {
  try {
    mvr::GpuICP icp(0);
    icp.setUseReciprocalCorrespondences(true);
    icp.setMaxCorrespondenceDistance(4.0);
    icp.setMaximumIterations(3);
    icp.setTransformationEpsilon(1e-6);
    icp.setEuclideanFitnessEpsilon(64);
    mvr::GpuICP::Cloud::Ptr a(new mvr::GpuICP::Cloud), b(new mvr::GpuICP::Cloud);
    for (int k = 0; k < 2000; ++k) {
      pcl::PointXYZ p = {(float)(k % 40), (float)(k / 40), 900.f + 0.01f * (float)(k % 7), 1.f};
      a->points.push_back(p);
      p.x += 0.3f; b->points.push_back(p);
    }
    icp.setInputSource(a); icp.setInputTarget(b);
    mvr::GpuICP::Cloud out;
    icp.align(out);
    Eigen::Matrix4f T = icp.getFinalTransformation();
    mvr::GpuCorrespondenceEstimation ce(0);
    ce.setInputSource(a); ce.setInputTarget(b);
    pcl::Correspondences corr;
    ce.determineReciprocalCorrespondences(corr, 4.0);
    std::printf("gpu ok: iterations %d tx %.3f fitness %.4f corr %zu converged %d\n", icp.report().iterations, T.m[12], icp.getFitnessScore(), corr.size(), (int)icp.hasConverged());
  } catch (const std::exception& e) {
    std::printf("no gpu: %s\n", e.what());
  }
  return 0;
}
