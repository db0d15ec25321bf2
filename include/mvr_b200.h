/* mvr_b200.h -- C ABI of the B200-native ICP alignment path (libmvr_b200.so).
 *
 * The reference (fanxiaochen/Multi-View-Registration) has no FFI of its own: its registration
 * driver (mvr/src/registrator.cpp) talks to PCL classes directly.  Each entry point below names the
 * reference / PCL call it replaces (paths relative to /root/reference); INTEGRATION.md shows the
 * shim a maintainer would put under those call sites.
 *
 * Conventions (identical to the reference's PCL types):
 *   points  : array of 16-byte records {float x, y, z, pad}  == pcl::PointXYZ (mvr/include/types.h:14)
 *   poses   : float[16] column-major, column-vector convention (p' = M p) == Eigen::Matrix4f,
 *             i.e. what icp.getFinalTransformation() returns (mvr/src/registrator.cpp:573)
 *   indices : int32, -1 = "no neighbour"
 *   distances are SQUARED float32 distances computed as ((dx*dx + dy*dy) + dz*dz) without FMA,
 *   nearest-neighbour ties resolve to the LOWEST original index.
 *
 * Every function returns an mvr_status (0 = ok) and never throws.  A context is bound to one CUDA
 * device and one stream and may be used by one thread at a time.  There is NO CPU fallback: without a
 * usable CUDA device mvr_ctx_create fails with MVR_ERR_CUDA.
 */
#ifndef MVR_B200_H
#define MVR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mvr_ctx mvr_ctx;

typedef enum {
  MVR_OK = 0,
  MVR_ERR_BAD_ARG = 1,
  MVR_ERR_TOO_FEW_CORRESPONDENCES = 2, /* < min_correspondences (PCL: "Not enough correspondences") */
  MVR_ERR_CUDA = 3,
  MVR_ERR_NO_INPUT = 4,                /* target/source not set */
  MVR_ERR_NOT_SPD = 5,                 /* point-to-plane normal equations not positive definite */
  MVR_ERR_ALLOC = 6
} mvr_status;

typedef enum { MVR_CLOUD_TARGET = 0, MVR_CLOUD_SOURCE = 1 } mvr_cloud;
typedef enum { MVR_POINT_TO_POINT = 0, MVR_POINT_TO_PLANE = 1 } mvr_estimator;
typedef enum {
  MVR_REASON_NONE = 0,
  MVR_REASON_ITERATIONS = 1,
  MVR_REASON_TRANSFORM = 2,
  MVR_REASON_ABS_MSE = 3,
  MVR_REASON_REL_MSE = 4,
  MVR_REASON_NO_CORRESPONDENCES = 5
} mvr_reason;

/* The setters the reference calls on pcl::IterativeClosestPoint (mvr/src/registrator.cpp:551-560,
 * 768-771, 901-904).  mvr_icp_params_default() fills in PCL's own defaults. */
typedef struct {
  int max_iterations;                  /* setMaximumIterations          (PCL default 10)            */
  double max_correspondence_distance;  /* setMaxCorrespondenceDistance  (PCL default sqrt(DBL_MAX)) */
  double transformation_epsilon;       /* setTransformationEpsilon      (PCL default 0)             */
  double euclidean_fitness_epsilon;    /* setEuclideanFitnessEpsilon    (PCL default -DBL_MAX)      */
  int use_reciprocal_correspondences;  /* setUseReciprocalCorrespondences (reference: true)         */
  int estimator;                       /* mvr_estimator; the reference only uses point-to-point     */
  int fixed_iterations;                /* 1: run exactly max_iterations (bench mode, criteria 2-4 off) */
  int min_correspondences;             /* PCL min_number_correspondences_ = 3                       */
} mvr_icp_params;

typedef struct {
  int iterations;
  int converged;          /* hasConverged() */
  int reason;             /* mvr_reason */
  int n_correspondences;  /* size of the last correspondence set */
  double mse;             /* mean squared distance of the last correspondence set */
  double gpu_ms;          /* device time of the align (CUDA events on the context stream) */
  uint64_t nn_queries;    /* nearest-neighbour queries answered (forward + reciprocal) */
} mvr_icp_report;

typedef struct {
  int iteration;          /* 1-based */
  int n_correspondences;
  double mse;
  float delta[16];        /* the float32 increment applied to the source in this iteration */
} mvr_icp_iteration;

/* Geometry of a uniform-grid index: cell c_a = clamp(floor((p_a - origin_a) * inv_cell), 0, 2^bits-1),
 * key = Morton interleave (x bit 0, y bit 1, z bit 2) of the three cell coordinates. */
typedef struct {
  float origin[3];
  float inv_cell;
  float cell;
  int bits;
} mvr_grid;

/* Per-kernel device timing accumulated while mvr_ctx_set_profiling(ctx, 1) is on. */
typedef struct {
  uint64_t launches;
  double ms;       /* sum of CUDA-event durations */
  double bytes;    /* algorithmic bytes summed over those launches */
  double units;    /* queries / points processed */
} mvr_kernel_stat;

enum { MVR_K_MORTON = 0, MVR_K_SORT = 1, MVR_K_TABLE = 2, MVR_K_NN = 3, MVR_K_CORR = 4, MVR_K_REDUCE = 5,
       MVR_K_TRANSFORM = 6, MVR_K_NORMALS = 7, MVR_K_COUNT = 8 };

const char* mvr_version(void);
/* Number of CUDA kernels this library has launched in the calling process (all contexts). */
uint64_t mvr_kernel_launch_count(void);
const char* mvr_status_string(int status);
void mvr_icp_params_default(mvr_icp_params* p);

/* -- context ---------------------------------------------------------------------------------- */
int mvr_ctx_create(int device, mvr_ctx** out);
int mvr_ctx_destroy(mvr_ctx* ctx);
/* Run on a caller-owned stream (e.g. torch's current stream) instead of the context's own. */
int mvr_ctx_set_stream(mvr_ctx* ctx, void* cuda_stream);
int mvr_ctx_synchronize(mvr_ctx* ctx);
const char* mvr_last_error(mvr_ctx* ctx);
int mvr_ctx_set_profiling(mvr_ctx* ctx, int on);
int mvr_ctx_get_kernel_stats(mvr_ctx* ctx, mvr_kernel_stat* out /* [MVR_K_COUNT] */, int reset);
/* Diagnostics of the last align (development aid): k = 0 clock cycles in the serial per-iteration solve, 1 solves. */
double mvr_debug_value(mvr_ctx* ctx, int k);
/* Tuning knobs of the spatial index: cell edge (<= 0: automatic) and maximum bits per axis (1..10). */
int mvr_ctx_set_index_options(mvr_ctx* ctx, float cell_edge, int max_bits);

/* -- inputs: icp.setInputTarget / icp.setInputSource (mvr/src/registrator.cpp:566-567, 776-777,
 *    913-914) and CorrespondenceEstimation::setInputSource/Target (:497-498, 645-646).
 *    Host variants copy n*16 bytes to the device; *_device variants adopt a device pointer that the
 *    caller keeps alive (no copy).  The target's grid index is built lazily by the first consumer. */
int mvr_set_target(mvr_ctx* ctx, const float* xyzw, size_t n);
int mvr_set_source(mvr_ctx* ctx, const float* xyzw, size_t n);
int mvr_set_target_device(mvr_ctx* ctx, const float* d_xyzw, size_t n);
int mvr_set_source_device(mvr_ctx* ctx, const float* d_xyzw, size_t n);
/* Optional per-target-point normals (n x {nx,ny,nz,curvature}) for point-to-plane ICP. */
int mvr_set_target_normals(mvr_ctx* ctx, const float* nxyzc, size_t n);

/* -- spatial index (replaces pcl::KdTreeFLANN::setInputCloud inside icp.align) ------------------- */
/* Build the Morton-sorted uniform grid over a cloud with an explicit grid; exported for parity tests. */
int mvr_index_build(mvr_ctx* ctx, int which /* mvr_cloud */, const mvr_grid* grid /* NULL: automatic */);
/* Copy out the index: sorted keys (n), permutation perm[sorted position] = original index (n),
 * cell_start ((1 << 3*bits) + 1 entries).  Any pointer may be NULL. */
int mvr_index_export(mvr_ctx* ctx, int which, mvr_grid* grid, uint32_t* sorted_keys, int32_t* perm,
                     uint32_t* cell_start);

/* -- queries ---------------------------------------------------------------------------------- */
/* Exact un-gated 1-NN of n query points in the target: pcl::KdTreeFLANN::nearestKSearch(k=1). */
int mvr_nn_query(mvr_ctx* ctx, const float* q_xyzw, size_t n, int32_t* idx, float* d2);
int mvr_nn_query_device(mvr_ctx* ctx, const float* d_q_xyzw, size_t n, int32_t* d_idx, float* d_d2);

/* CorrespondenceEstimation::determineCorrespondences / determineReciprocalCorrespondences
 * (mvr/src/registrator.cpp:502, 649): source -> target, gate d2 <= max_dist^2, output compacted in
 * ascending source index.  Arrays must hold source-size entries. */
int mvr_correspondences(mvr_ctx* ctx, double max_dist, int reciprocal, int32_t* index_query, int32_t* index_match,
                        float* distance, size_t* count);

/* -- ICP: icp.align(output, guess) + getFinalTransformation (mvr/src/registrator.cpp:569-573, 920-921,
 *    1012-1013).  guess may be NULL (identity).  out_xyzw (nullable, may alias the source host buffer)
 *    receives transform(source, final).  Returns MVR_ERR_TOO_FEW_CORRESPONDENCES when PCL would have
 *    logged "Not enough correspondences"; out_pose then holds the transform accumulated so far. */
int mvr_icp_align(mvr_ctx* ctx, const mvr_icp_params* params, const float* guess, float* out_pose, float* out_xyzw,
                  mvr_icp_report* report);
/* Per-iteration records of the last align (n_correspondences, mse, delta). */
int mvr_icp_get_iterations(mvr_ctx* ctx, mvr_icp_iteration* out, int max_records, int* count);
/* Registration::getFitnessScore(max_range) (mvr/src/registrator.cpp:572, 923, 1015): mean squared
 * un-gated NN distance of the last align's output cloud (or of the source if no align ran). */
int mvr_fitness_score(mvr_ctx* ctx, double max_range, double* score);

/* -- extensions named by the north star (no reference call site) -------------------------------- */
/* pcl::NormalEstimation semantics: kNN(k) PCA normals of a cloud, flipped towards viewpoint.
 * out: n x {nx, ny, nz, curvature}.  neighbours (nullable): n x k int32, ascending (d2, index). */
int mvr_estimate_normals(mvr_ctx* ctx, int which, int k, const float viewpoint[3], float* out_nxyzc,
                         int32_t* neighbours);

#ifdef __cplusplus
}
#endif
#endif /* MVR_B200_H */
