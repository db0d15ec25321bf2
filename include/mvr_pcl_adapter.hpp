// mvr_pcl_adapter.hpp -- header-only adapters with the surface of the PCL classes the reference calls on its ICP path, on top of
// the C ABI of mvr_b200.h.  A maintainer swaps the types at three places of the reference and links libmvr_b200.so:
//   mvr/include/registrator.h:91   pcl::IterativeClosestPoint<PCLPoint, PCLPoint> icp_;            -> mvr::GpuICP icp_;
//   mvr/src/registrator.cpp:551    pcl::IterativeClosestPoint<PCLPoint, PCLPoint> icp;             -> mvr::GpuICP icp;
//   mvr/src/registrator.cpp:496, 644  pcl::registration::CorrespondenceEstimation<PCLPoint, PCLPoint, float>
//                                                                                                   -> mvr::GpuCorrespondenceEstimation
// Only compiled where PCL and Eigen are installed (this repository's image has neither: tests/test_abi_and_host.py compiles
// the header against the stand-in declarations of tests/mock_pcl/).  pcl::PointXYZ already is the 16-byte {x, y, z, pad} record
// and Eigen::Matrix4f the column-major float[16] of the ABI (mvr/include/types.h:14-50): nothing is marshalled.
#ifndef MVR_PCL_ADAPTER_HPP
#define MVR_PCL_ADAPTER_HPP

#if defined(__has_include)
#if __has_include(<pcl/point_cloud.h>) && __has_include(<pcl/point_types.h>) && __has_include(<pcl/correspondence.h>) && __has_include(<Eigen/Core>)
#define MVR_HAVE_PCL 1
#endif
#endif

#ifdef MVR_HAVE_PCL
#include <cfloat>
#include <cstdint>
#include <stdexcept>
#include <vector>

#include <Eigen/Core>
#include <pcl/correspondence.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "mvr_b200.h"

namespace mvr {

// The members of pcl::IterativeClosestPoint<pcl::PointXYZ, pcl::PointXYZ> the reference uses (mvr/src/registrator.cpp:551-576,
// 768-777, 901-923, 1012-1015, 1024-1025), same names, argument meaning and error behaviour: align() never throws; a failed
// align leaves hasConverged() false and getFinalTransformation() at what had been accumulated (PCL logs and carries on).
class GpuICP {
 public:
  typedef pcl::PointCloud<pcl::PointXYZ> Cloud;
  explicit GpuICP(int device = 0) : ctx_(NULL), status_(MVR_OK) {
    if (mvr_ctx_create(device, &ctx_) != MVR_OK) throw std::runtime_error("mvr_ctx_create: no usable CUDA device (there is no CPU fallback)");
    mvr_icp_params_default(&prm_);   // PCL's own defaults: 10 iterations, no gate, epsilons off, one-way correspondences
    rep_ = mvr_icp_report();
    final_.setIdentity();
  }
  ~GpuICP() { mvr_ctx_destroy(ctx_); }
  void setUseReciprocalCorrespondences(bool b) { prm_.use_reciprocal_correspondences = b ? 1 : 0; }   // registrator.cpp:552, 768, 901
  void setMaxCorrespondenceDistance(double d) { prm_.max_correspondence_distance = d; }               // :554, 769, 902
  void setMaximumIterations(int n) { prm_.max_iterations = n; }                                       // :556, 770, 903
  void setTransformationEpsilon(double e) { prm_.transformation_epsilon = e; }                        // :558
  void setEuclideanFitnessEpsilon(double e) { prm_.euclidean_fitness_epsilon = e; }                   // :560, 771, 904
  void setInputSource(const Cloud::ConstPtr& c) {                                                     // :566, 776, 913
    src_ = c;
    status_ = mvr_set_source(ctx_, c->empty() ? NULL : &c->points[0].x, c->size());
  }
  void setInputTarget(const Cloud::ConstPtr& c) {                                                     // :567, 777, 914
    status_ = mvr_set_target(ctx_, c->empty() ? NULL : &c->points[0].x, c->size());
  }
  // icp.align(output): output = transform(source, final); output may be the source cloud itself (the reference aligns in place,
  // icp_.align(*source_), :920, 1012, 1024)
  void align(Cloud& output) { align(output, Eigen::Matrix4f::Identity()); }
  void align(Cloud& output, const Eigen::Matrix4f& guess) {
    const size_t n = src_ ? src_->size() : 0;
    if (&output != src_.get()) { output.points.resize(n); output.width = (uint32_t)n; output.height = 1; output.is_dense = src_ ? src_->is_dense : true; }
    status_ = mvr_icp_align(ctx_, &prm_, guess.data(), final_.data(), n ? &output.points[0].x : NULL, &rep_);
  }
  Eigen::Matrix4f getFinalTransformation() const { return final_; }                                   // :573, 921, 1013, 1025
  double getFitnessScore(double max_range = DBL_MAX) {                                                // :572, 923, 1015
    double s = DBL_MAX;
    mvr_fitness_score(ctx_, max_range, &s);
    return s;
  }
  bool hasConverged() const { return rep_.converged != 0; }
  int status() const { return status_; }   // the mvr_status of the last call (PCL has no counterpart: it logs)
  const mvr_icp_report& report() const { return rep_; }
  mvr_ctx* context() const { return ctx_; }

 private:
  GpuICP(const GpuICP&);
  GpuICP& operator=(const GpuICP&);
  mvr_ctx* ctx_;
  mvr_icp_params prm_;
  mvr_icp_report rep_;
  Cloud::ConstPtr src_;
  Eigen::Matrix4f final_;
  int status_;
};

// pcl::registration::CorrespondenceEstimation<pcl::PointXYZ, pcl::PointXYZ, float> as used at mvr/src/registrator.cpp:496-502
// and 644-650: correspondences in ascending source index, distance = squared float distance.
class GpuCorrespondenceEstimation {
 public:
  typedef pcl::PointCloud<pcl::PointXYZ> Cloud;
  explicit GpuCorrespondenceEstimation(int device = 0) : ctx_(NULL), n_src_(0) {
    if (mvr_ctx_create(device, &ctx_) != MVR_OK) throw std::runtime_error("mvr_ctx_create: no usable CUDA device (there is no CPU fallback)");
  }
  ~GpuCorrespondenceEstimation() { mvr_ctx_destroy(ctx_); }
  void setInputSource(const Cloud::ConstPtr& c) { n_src_ = c->size(); mvr_set_source(ctx_, c->empty() ? NULL : &c->points[0].x, c->size()); }
  void setInputTarget(const Cloud::ConstPtr& c) { mvr_set_target(ctx_, c->empty() ? NULL : &c->points[0].x, c->size()); }
  void determineCorrespondences(pcl::Correspondences& out, double max_distance = DBL_MAX) { run(out, max_distance, 0); }
  void determineReciprocalCorrespondences(pcl::Correspondences& out, double max_distance = DBL_MAX) { run(out, max_distance, 1); }

 private:
  GpuCorrespondenceEstimation(const GpuCorrespondenceEstimation&);
  GpuCorrespondenceEstimation& operator=(const GpuCorrespondenceEstimation&);
  void run(pcl::Correspondences& out, double max_distance, int reciprocal) {
    std::vector<int32_t> q(n_src_ ? n_src_ : 1), m(n_src_ ? n_src_ : 1);
    std::vector<float> d(n_src_ ? n_src_ : 1);
    size_t n = 0;
    // PCL's default "no gate" is DBL_MAX; the ABI takes any distance whose square is representable
    if (mvr_correspondences(ctx_, max_distance > 1e150 ? 1e150 : max_distance, reciprocal, &q[0], &m[0], &d[0], &n) != MVR_OK) n = 0;
    out.resize(n);
    for (size_t k = 0; k < n; ++k) out[k] = pcl::Correspondence(q[k], m[k], d[k]);
  }
  mvr_ctx* ctx_;
  size_t n_src_;
};

}  // namespace mvr
#endif  // MVR_HAVE_PCL
#endif  // MVR_PCL_ADAPTER_HPP
