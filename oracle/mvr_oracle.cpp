// mvr_oracle.cpp -- CPU ORACLE for the ICP alignment path.  TEST INFRASTRUCTURE ONLY.
//
// This file is a from-scratch CPU restatement of the arithmetic that the reference
// (fanxiaochen/Multi-View-Registration) delegates to PCL on its ICP hot path.  It is the
// CHECKER for the CUDA product in multi-view-registration_b200/: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// The product library never links, imports or falls back to anything in oracle/.
//
// PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors, and the library
// that holds the arithmetic (PCL, un-vendored, no version pin: mvr/CMakeLists.txt:10) is not
// installable here, so this oracle could not be checked against the reference binary.  It is
// pinned instead against independent implementations (brute force, scipy cKDTree, numpy SVD)
// in tests/ and against the committed fixtures in tests/golden/.
//
// What each function follows (reference call sites, /root/reference relative):
//   orc_transform            PointCloud::getTransformedPoints  mvr/src/point_cloud.cpp:290-303 (pose apply)
//                            and ICP's in-place transformCloud  [PCL, called by icp.align at
//                            mvr/src/registrator.cpp:569, 920, 1012, 1024]
//   orc_nn_* / kd-tree       pcl::KdTreeFLANN exact 1-NN (L2_Simple float, sequential sum) used by
//                            icp.align and CorrespondenceEstimation  mvr/src/registrator.cpp:496-502, 566-569
//   orc_correspondences      determineCorrespondences / determineReciprocalCorrespondences
//                            mvr/src/registrator.cpp:502, 552, 649, 768, 901
//   orc_estimate_rigid_svd   TransformationEstimationSVD (Eigen::umeyama)   [inside icp.align]
//   orc_icp_align            IterativeClosestPoint::computeTransformation + DefaultConvergenceCriteria
//                            configured at mvr/src/registrator.cpp:551-560, 768-771, 901-904
//   orc_fitness_score        Registration::getFitnessScore  mvr/src/registrator.cpp:572, 923, 1015
//   orc_morton_* / cell tbl  no reference counterpart (index is an implementation choice); restated
//                            here so the GPU index can be compared bit for bit
//   orc_estimate_normals     pcl::NormalEstimation semantics (north-star extension, no call site)
//   orc_estimate_point_to_plane  TransformationEstimationPointToPlaneLLS (north-star extension)
//
// Arithmetic rules, identical on CPU and GPU:
//   d2 = ((dx*dx + dy*dy) + dz*dz) in IEEE float32, round-to-nearest, NO fused multiply-add
//   (compile with -ffp-contract=off); argmin is lexicographic on (d2, original index);
//   gate keeps a pair iff !((double)d2 > max_dist*max_dist).
//
// Build: see oracle/Makefile  (g++ -O2 -ffp-contract=off -fopenmp -shared -fPIC).

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

struct orc_icp_params {
  int max_iterations;                  // PCL default 10
  double max_correspondence_distance;  // PCL default sqrt(DBL_MAX)
  double transformation_epsilon;       // PCL default 0
  double euclidean_fitness_epsilon;    // PCL default -DBL_MAX
  int use_reciprocal;                  // reference: true
  int estimator;                       // 0 point-to-point SVD, 1 point-to-plane LLS
  int fixed_iterations;                // 1: only the iteration cap stops the loop
  int min_correspondences;             // PCL: 3
};

struct orc_icp_report {
  int iterations;
  int converged;        // 1 if a criterion fired
  int reason;           // 0 none, 1 iterations, 2 transform, 3 abs mse, 4 rel mse, 5 no correspondences
  int n_correspondences;
  double mse;           // mean squared distance of the last correspondence set
  unsigned long long nn_queries;  // tree queries answered: one per source point per iteration + one per
                                  // gate-passing point when the reciprocal test is on
};

struct orc_iter_record {
  int iteration;
  int n_corr;
  double mse;
  double delta[16];     // column-major 4x4, value of the float delta applied this iteration
};

}  // extern "C"

namespace {

struct P4 { float x, y, z, w; };

inline bool finite3(const P4& p) { return std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z); }

// Pinned squared distance: float32, no FMA (file is built with -ffp-contract=off).
inline float dist2(const P4& a, const P4& b) {
  float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  float xx = dx * dx, yy = dy * dy, zz = dz * dz;
  return (xx + yy) + zz;
}

inline bool lex_less(float d2, int idx, float bd2, int bidx) {
  return d2 < bd2 || (d2 == bd2 && idx < bidx);
}

// ------------------------------------------------------------------------------------------
// Exact 1-NN kd-tree (median split on the widest axis, leaves of <= 16 points).
// Pruning is done on the float value the distance formula itself would produce for the
// splitting-plane gap, so the traversal is exact in float arithmetic including ties.
// ------------------------------------------------------------------------------------------
struct KdNode { float split; int axis; int left, right; int begin, end; };

struct KdTree {
  std::vector<P4> pts;       // copy of the finite points, permuted into tree order
  std::vector<int> idx;      // original index of pts[k]
  std::vector<KdNode> nodes;

  int build(int begin, int end) {
    int id = (int)nodes.size();
    nodes.push_back(KdNode{0.f, -1, -1, -1, begin, end});
    if (end - begin <= 16) return id;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = begin; k < end; ++k) {
      const float c[3] = {pts[k].x, pts[k].y, pts[k].z};
      for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], c[a]); hi[a] = std::max(hi[a], c[a]); }
    }
    int axis = 0;
    if (hi[1] - lo[1] > hi[axis] - lo[axis]) axis = 1;
    if (hi[2] - lo[2] > hi[axis] - lo[axis]) axis = 2;
    if (!(hi[axis] > lo[axis])) return id;  // all identical: keep as a (large) leaf
    int mid = (begin + end) / 2;
    std::vector<int> order(end - begin);
    std::iota(order.begin(), order.end(), begin);
    auto coord = [&](int k) { return axis == 0 ? pts[k].x : axis == 1 ? pts[k].y : pts[k].z; };
    std::nth_element(order.begin(), order.begin() + (mid - begin), order.end(),
                     [&](int a, int b) { return coord(a) < coord(b); });
    std::vector<P4> tp(end - begin);
    std::vector<int> ti(end - begin);
    for (int k = 0; k < end - begin; ++k) { tp[k] = pts[order[k]]; ti[k] = idx[order[k]]; }
    std::copy(tp.begin(), tp.end(), pts.begin() + begin);
    std::copy(ti.begin(), ti.end(), idx.begin() + begin);
    float split = axis == 0 ? pts[mid].x : axis == 1 ? pts[mid].y : pts[mid].z;
    // left: coord <= split (indices < mid hold values <= split), right: coord >= split
    nodes[id].axis = axis;
    nodes[id].split = split;
    int l = build(begin, mid);
    int r = build(mid, end);
    nodes[id].left = l;
    nodes[id].right = r;
    return id;
  }

  void init(const P4* p, int n) {
    pts.clear(); idx.clear(); nodes.clear();
    pts.reserve(n); idx.reserve(n);
    for (int i = 0; i < n; ++i)
      if (finite3(p[i])) { pts.push_back(p[i]); idx.push_back(i); }
    if (!pts.empty()) build(0, (int)pts.size());
  }

  void search(int node, const P4& q, float& bd2, int& bidx) const {
    const KdNode& nd = nodes[node];
    if (nd.axis < 0) {
      for (int k = nd.begin; k < nd.end; ++k) {
        float d2 = dist2(q, pts[k]);
        if (lex_less(d2, idx[k], bd2, bidx)) { bd2 = d2; bidx = idx[k]; }
      }
      return;
    }
    float qc = nd.axis == 0 ? q.x : nd.axis == 1 ? q.y : q.z;
    float gap = qc - nd.split;
    float gap2 = gap * gap;   // the value dx*dx would have for a point lying on the plane
    int near = (qc <= nd.split) ? nd.left : nd.right;
    int far = (qc <= nd.split) ? nd.right : nd.left;
    search(near, q, bd2, bidx);
    // any point on the far side has d2 >= gap2 in float arithmetic (rounding is monotone);
    // visit on equality too so that a lower index at the same distance is still found.
    if (!(gap2 > bd2)) search(far, q, bd2, bidx);
  }

  void query(const P4& q, int& out_idx, float& out_d2) const {
    float bd2 = INFINITY; int bidx = INT_MAX;
    if (!pts.empty() && finite3(q)) search(0, q, bd2, bidx);
    if (bidx == INT_MAX) { out_idx = -1; out_d2 = INFINITY; }
    else { out_idx = bidx; out_d2 = bd2; }
  }
};

// ------------------------------------------------------------------------------------------
// 3x3 SVD through the symmetric Jacobi eigen-decomposition of A^T A (double).
// ------------------------------------------------------------------------------------------
void jacobi_eig3(double S[3][3], double V[3][3], double w[3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = S[0][1] * S[0][1] + S[0][2] * S[0][2] + S[1][2] * S[1][2];
    double diag = S[0][0] * S[0][0] + S[1][1] * S[1][1] + S[2][2] * S[2][2];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (S[p][q] == 0.0) continue;
        double theta = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) {  // S <- S J
          double skp = S[k][p], skq = S[k][q];
          S[k][p] = c * skp - s * skq; S[k][q] = s * skp + c * skq;
        }
        for (int k = 0; k < 3; ++k) {  // S <- J^T S
          double spk = S[p][k], sqk = S[q][k];
          S[p][k] = c * spk - s * sqk; S[q][k] = s * spk + c * sqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < 3; ++i) w[i] = S[i][i];
}

double det3(const double M[3][3]) {
  return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
         M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

// A = U diag(s) V^T, s descending, U and V orthogonal (not forced to be proper).
void svd3(const double A[3][3], double U[3][3], double s[3], double V[3][3]) {
  double AtA[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { AtA[i][j] = 0; for (int k = 0; k < 3; ++k) AtA[i][j] += A[k][i] * A[k][j]; }
  double Vt[3][3], w[3];
  jacobi_eig3(AtA, Vt, w);
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int a, int b) { return w[a] > w[b]; });
  for (int c = 0; c < 3; ++c) {
    s[c] = std::sqrt(std::max(w[ord[c]], 0.0));
    for (int r = 0; r < 3; ++r) V[r][c] = Vt[r][ord[c]];
  }
  // U columns: A v / s, re-orthonormalised (Gram-Schmidt); degenerate columns completed by cross product
  double Ucol[3][3];
  for (int c = 0; c < 3; ++c)
    for (int r = 0; r < 3; ++r) { Ucol[c][r] = 0; for (int k = 0; k < 3; ++k) Ucol[c][r] += A[r][k] * V[k][c]; }
  auto norm = [](double* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
  double n0 = norm(Ucol[0]);
  if (n0 > 0) for (int r = 0; r < 3; ++r) Ucol[0][r] /= n0; else { Ucol[0][0] = 1; Ucol[0][1] = Ucol[0][2] = 0; }
  double d01 = Ucol[1][0] * Ucol[0][0] + Ucol[1][1] * Ucol[0][1] + Ucol[1][2] * Ucol[0][2];
  for (int r = 0; r < 3; ++r) Ucol[1][r] -= d01 * Ucol[0][r];
  double n1 = norm(Ucol[1]);
  if (n1 > 1e-12 * (s[0] > 0 ? s[0] : 1.0)) for (int r = 0; r < 3; ++r) Ucol[1][r] /= n1;
  else {  // pick any unit vector orthogonal to column 0
    int m = std::fabs(Ucol[0][0]) < std::fabs(Ucol[0][1]) ? (std::fabs(Ucol[0][0]) < std::fabs(Ucol[0][2]) ? 0 : 2)
                                                            : (std::fabs(Ucol[0][1]) < std::fabs(Ucol[0][2]) ? 1 : 2);
    double e[3] = {0, 0, 0}; e[m] = 1;
    double d = e[0] * Ucol[0][0] + e[1] * Ucol[0][1] + e[2] * Ucol[0][2];
    for (int r = 0; r < 3; ++r) Ucol[1][r] = e[r] - d * Ucol[0][r];
    n1 = norm(Ucol[1]);
    for (int r = 0; r < 3; ++r) Ucol[1][r] /= n1;
  }
  // third column: keep the sign of A v2 when it is meaningful, else the cross product
  double cr[3] = {Ucol[0][1] * Ucol[1][2] - Ucol[0][2] * Ucol[1][1], Ucol[0][2] * Ucol[1][0] - Ucol[0][0] * Ucol[1][2],
                  Ucol[0][0] * Ucol[1][1] - Ucol[0][1] * Ucol[1][0]};
  double sgn = Ucol[2][0] * cr[0] + Ucol[2][1] * cr[1] + Ucol[2][2] * cr[2];
  for (int r = 0; r < 3; ++r) Ucol[2][r] = (sgn < 0 ? -cr[r] : cr[r]);
  for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) U[r][c] = Ucol[c][r];
}

// Umeyama without scaling (Eigen::umeyama as used by TransformationEstimationSVD):
// returns column-major 4x4 double T such that dst ~ T * src.
void umeyama_from_sums(double n, const double mu_s[3], const double mu_d[3], const double Sigma[3][3], double T[16]) {
  (void)n;
  double U[3][3], V[3][3], sv[3];
  svd3(Sigma, U, sv, V);
  double S[3] = {1, 1, 1};
  if (det3(U) * det3(V) < 0) S[2] = -1;
  double R[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) { R[i][j] = 0; for (int k = 0; k < 3; ++k) R[i][j] += U[i][k] * S[k] * V[j][k]; }
  double t[3];
  for (int i = 0; i < 3; ++i) t[i] = mu_d[i] - (R[i][0] * mu_s[0] + R[i][1] * mu_s[1] + R[i][2] * mu_s[2]);
  for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) T[c * 4 + r] = (r == c);
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[j * 4 + i] = R[i][j]; T[12 + i] = t[i]; }
}

// Pinned point transform (float32, no FMA): x' = ((m00*x + m01*y) + m02*z) + m03.
inline P4 xform(const float M[16], const P4& p) {
  P4 r;
  r.x = ((M[0] * p.x + M[4] * p.y) + M[8] * p.z) + M[12];
  r.y = ((M[1] * p.x + M[5] * p.y) + M[9] * p.z) + M[13];
  r.z = ((M[2] * p.x + M[6] * p.y) + M[10] * p.z) + M[14];
  r.w = 1.0f;
  return r;
}

void matmul4d(const double A[16], const double B[16], double C[16]) {  // column-major C = A*B
  double t[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) { double s = 0; for (int k = 0; k < 4; ++k) s += A[k * 4 + r] * B[c * 4 + k]; t[c * 4 + r] = s; }
  std::memcpy(C, t, sizeof(t));
}

struct Corr { int q, m; float d2; };

void correspondences_impl(const P4* src, int n, const P4* tgt, const KdTree& ttree, double max_dist, int reciprocal,
                          std::vector<Corr>& out, unsigned long long* n_queries = nullptr) {
  const double max2 = max_dist * max_dist;
  KdTree stree;
  if (reciprocal) stree.init(src, n);
  std::vector<int> fj(n); std::vector<float> fd(n); std::vector<char> keep(n, 0);
  unsigned long long nq = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : nq)
  for (int i = 0; i < n; ++i) {
    int j; float d2;
    ttree.query(src[i], j, d2);
    fj[i] = j; fd[i] = d2;
    ++nq;
    if (j < 0 || (double)d2 > max2) continue;
    if (reciprocal) {
      int ib; float db;
      ++nq;
      stree.query(tgt[j], ib, db);
      if (ib != i || (double)db > max2) continue;
    }
    keep[i] = 1;
  }
  out.clear();
  for (int i = 0; i < n; ++i) if (keep[i]) out.push_back(Corr{i, fj[i], fd[i]});
  if (n_queries) *n_queries += nq;
}

void estimate_svd_impl(const P4* src, const P4* tgt, const std::vector<Corr>& corr, double T[16]) {
  double n = (double)corr.size();
  double ms[3] = {0, 0, 0}, md[3] = {0, 0, 0};
  for (const Corr& c : corr) {
    ms[0] += src[c.q].x; ms[1] += src[c.q].y; ms[2] += src[c.q].z;
    md[0] += tgt[c.m].x; md[1] += tgt[c.m].y; md[2] += tgt[c.m].z;
  }
  for (int a = 0; a < 3; ++a) { ms[a] /= n; md[a] /= n; }
  double Sg[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (const Corr& c : corr) {
    double s[3] = {src[c.q].x - ms[0], src[c.q].y - ms[1], src[c.q].z - ms[2]};
    double d[3] = {tgt[c.m].x - md[0], tgt[c.m].y - md[1], tgt[c.m].z - md[2]};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Sg[i][j] += d[i] * s[j];
  }
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Sg[i][j] /= n;
  umeyama_from_sums(n, ms, md, Sg, T);
}

// 6x6 SPD solve by Cholesky (double).  Returns false if not positive definite.
bool cholesky_solve6(double A[6][6], double b[6], double x[6]) {
  double L[6][6] = {};
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = A[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
      if (i == j) { if (!(s > 0)) return false; L[i][i] = std::sqrt(s); }
      else L[i][j] = s / L[j][j];
    }
  double y[6];
  for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
  for (int i = 5; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < 6; ++k) s -= L[k][i] * x[k]; x[i] = s / L[i][i]; }
  return true;
}

// TransformationEstimationPointToPlaneLLS: normals live in tgt_n (xyz of float4).
bool estimate_p2plane_impl(const P4* src, const P4* tgt, const P4* tgt_n, const std::vector<Corr>& corr, double T[16]) {
  double A[6][6] = {}, b[6] = {};
  for (const Corr& c : corr) {
    double sx = src[c.q].x, sy = src[c.q].y, sz = src[c.q].z;
    double dx = tgt[c.m].x, dy = tgt[c.m].y, dz = tgt[c.m].z;
    double nx = tgt_n[c.m].x, ny = tgt_n[c.m].y, nz = tgt_n[c.m].z;
    double J[6] = {nz * sy - ny * sz, nx * sz - nz * sx, ny * sx - nx * sy, nx, ny, nz};
    double r = nx * (dx - sx) + ny * (dy - sy) + nz * (dz - sz);
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) A[i][j] += J[i] * J[j]; b[i] += J[i] * r; }
  }
  double x[6];
  if (!cholesky_solve6(A, b, x)) return false;
  double ca = std::cos(x[0]), sa = std::sin(x[0]), cb = std::cos(x[1]), sb = std::sin(x[1]), cg = std::cos(x[2]), sg = std::sin(x[2]);
  double R[3][3] = {{cg * cb, -sg * ca + cg * sb * sa, sg * sa + cg * sb * ca},
                    {sg * cb, cg * ca + sg * sb * sa, -cg * sa + sg * sb * ca},
                    {-sb, cb * sa, cb * ca}};
  for (int c = 0; c < 4; ++c) for (int r = 0; r < 4; ++r) T[c * 4 + r] = (r == c);
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[j * 4 + i] = R[i][j]; T[12 + i] = x[3 + i]; }
  return true;
}

inline uint32_t part1by2(uint32_t v) {  // spread the low 10 bits to every third bit
  v &= 0x3ffu;
  v = (v | (v << 16)) & 0x030000ffu;
  v = (v | (v << 8)) & 0x0300f00fu;
  v = (v | (v << 4)) & 0x030c30c3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

inline int cell_coord(float p, float o, float inv, int G) {
  float t = (p - o) * inv;
  float f = std::floor(t);
  f = std::min(std::max(f, 0.0f), (float)(G - 1));
  return (int)f;
}

// symmetric 3x3 eigen-decomposition, ascending eigenvalues
void eig_sym3(const double C[3][3], double w[3], double V[3][3]) {
  double S[3][3];
  std::memcpy(S, C, sizeof(S));
  double Vt[3][3], ww[3];
  jacobi_eig3(S, Vt, ww);
  int ord[3] = {0, 1, 2};
  std::sort(ord, ord + 3, [&](int a, int b) { return ww[a] < ww[b]; });
  for (int c = 0; c < 3; ++c) { w[c] = ww[ord[c]]; for (int r = 0; r < 3; ++r) V[r][c] = Vt[r][ord[c]]; }
}

struct KnnHeap {  // k smallest (d2, idx) lexicographic; simple insertion into a sorted array
  int k, n; float* d; int* id;
  void push(float d2, int i) {
    if (n == k && !lex_less(d2, i, d[n - 1], id[n - 1])) return;
    int p = (n < k) ? n++ : k - 1;
    while (p > 0 && lex_less(d2, i, d[p - 1], id[p - 1])) { d[p] = d[p - 1]; id[p] = id[p - 1]; --p; }
    d[p] = d2; id[p] = i;
  }
};

void knn_search(const KdTree& t, int node, const P4& q, KnnHeap& h) {
  const KdNode& nd = t.nodes[node];
  if (nd.axis < 0) {
    for (int k = nd.begin; k < nd.end; ++k) h.push(dist2(q, t.pts[k]), t.idx[k]);
    return;
  }
  float qc = nd.axis == 0 ? q.x : nd.axis == 1 ? q.y : q.z;
  float gap = qc - nd.split, gap2 = gap * gap;
  int near = (qc <= nd.split) ? nd.left : nd.right, far = (qc <= nd.split) ? nd.right : nd.left;
  knn_search(t, near, q, h);
  if (h.n < h.k || !(gap2 > h.d[h.n - 1])) knn_search(t, far, q, h);
}

}  // namespace

extern "C" {

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n > 0 ? n : 1);
#else
  (void)n;
#endif
}

// p' = M * [x y z 1]^T, float32, pinned order; w lane written as 1 (PCL data[3] convention).
// Non-finite points are copied through unchanged (PCL's transform skips them).
void orc_transform(const float* pts, int n, const float* M, float* out) {
  const P4* p = (const P4*)pts; P4* o = (P4*)out;
  for (int i = 0; i < n; ++i) o[i] = finite3(p[i]) ? xform(M, p[i]) : p[i];
}

// PointCloud::getTransformedPoints (mvr/src/point_cloud.cpp:290-303): the pose is a double 4x4
// (given here column-major, column-vector convention), applied in double and narrowed to float.
void orc_apply_pose_double(const float* pts, int n, int stride_floats, const double* M, float* out_xyzw) {
  for (int i = 0; i < n; ++i) {
    const float* p = pts + (size_t)i * stride_floats;
    double x = p[0], y = p[1], z = p[2];
    // osg::Matrix::preMult(Vec3f): d = 1/(m[0][3]x+m[1][3]y+m[2][3]z+m[3][3]); result*(d) ; d == 1 for rigid poses
    double rx = M[0] * x + M[4] * y + M[8] * z + M[12];
    double ry = M[1] * x + M[5] * y + M[9] * z + M[13];
    double rz = M[2] * x + M[6] * y + M[10] * z + M[14];
    out_xyzw[4 * (size_t)i + 0] = (float)rx; out_xyzw[4 * (size_t)i + 1] = (float)ry;
    out_xyzw[4 * (size_t)i + 2] = (float)rz; out_xyzw[4 * (size_t)i + 3] = 1.0f;
  }
}

void orc_nn_brute(const float* tgt, int m, const float* q, int n, int* idx, float* d2) {
  const P4* t = (const P4*)tgt; const P4* qq = (const P4*)q;
#pragma omp parallel for
  for (int i = 0; i < n; ++i) {
    float bd = INFINITY; int bi = INT_MAX;
    if (finite3(qq[i]))
      for (int j = 0; j < m; ++j) {
        if (!finite3(t[j])) continue;
        float d = dist2(qq[i], t[j]);
        if (lex_less(d, j, bd, bi)) { bd = d; bi = j; }
      }
    if (bi == INT_MAX) { idx[i] = -1; d2[i] = INFINITY; } else { idx[i] = bi; d2[i] = bd; }
  }
}

void* orc_kdtree_build(const float* tgt, int m) {
  KdTree* t = new KdTree();
  t->init((const P4*)tgt, m);
  return t;
}
void orc_kdtree_free(void* h) { delete (KdTree*)h; }
void orc_kdtree_query(void* h, const float* q, int n, int* idx, float* d2) {
  const KdTree* t = (const KdTree*)h; const P4* qq = (const P4*)q;
#pragma omp parallel for schedule(dynamic, 1024)
  for (int i = 0; i < n; ++i) t->query(qq[i], idx[i], d2[i]);
}

void orc_nn_kdtree(const float* tgt, int m, const float* q, int n, int* idx, float* d2) {
  KdTree t; t.init((const P4*)tgt, m);
  orc_kdtree_query(&t, q, n, idx, d2);
}

// out arrays must hold n entries; returns the count (ascending source index, compacted).
int orc_correspondences(const float* src, int n, const float* tgt, int m, double max_dist, int reciprocal, int* q_out,
                        int* m_out, float* d2_out) {
  KdTree tt; tt.init((const P4*)tgt, m);
  std::vector<Corr> c;
  correspondences_impl((const P4*)src, n, (const P4*)tgt, tt, max_dist, reciprocal, c);
  for (size_t k = 0; k < c.size(); ++k) { q_out[k] = c[k].q; m_out[k] = c[k].m; d2_out[k] = c[k].d2; }
  return (int)c.size();
}

// T (column-major double 4x4) with tgt ~ T*src over the given correspondences.
int orc_estimate_rigid_svd(const float* src, const float* tgt, const int* q, const int* m, int count, double* T) {
  if (count < 1) return 1;
  std::vector<Corr> c(count);
  for (int k = 0; k < count; ++k) c[k] = Corr{q[k], m[k], 0.f};
  estimate_svd_impl((const P4*)src, (const P4*)tgt, c, T);
  return 0;
}

int orc_estimate_point_to_plane(const float* src, const float* tgt, const float* tgt_normals, const int* q, const int* m,
                                int count, double* T) {
  std::vector<Corr> c(count);
  for (int k = 0; k < count; ++k) c[k] = Corr{q[k], m[k], 0.f};
  return estimate_p2plane_impl((const P4*)src, (const P4*)tgt, (const P4*)tgt_normals, c, T) ? 0 : 1;
}

void orc_svd3(const double* A_rowmajor, double* U_rowmajor, double* s, double* V_rowmajor) {
  double A[3][3], U[3][3], V[3][3];
  std::memcpy(A, A_rowmajor, sizeof(A));
  svd3(A, U, s, V);
  std::memcpy(U_rowmajor, U, sizeof(U)); std::memcpy(V_rowmajor, V, sizeof(V));
}

// IterativeClosestPoint::computeTransformation with DefaultConvergenceCriteria (A8).
// guess / final are column-major float[16] (Eigen::Matrix4f layout); out_xyzw (nullable) receives
// transform(input, final).  log (nullable) receives up to max_log per-iteration records.
int orc_icp_align(const float* src_in, int n, const float* tgt_in, int m, const float* tgt_normals,
                  const orc_icp_params* prm, const float* guess, float* final_out, float* out_xyzw, orc_icp_report* rep,
                  orc_iter_record* log, int max_log, int* n_log) {
  const P4* src0 = (const P4*)src_in; const P4* tgt = (const P4*)tgt_in;
  std::vector<P4> cur(n);
  float g[16];
  static const float I16[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  std::memcpy(g, guess ? guess : I16, sizeof(g));
  bool guess_is_identity = std::memcmp(g, I16, sizeof(g)) == 0;
  for (int i = 0; i < n; ++i) cur[i] = (guess_is_identity || !finite3(src0[i])) ? src0[i] : xform(g, src0[i]);
  double fin[16];
  for (int k = 0; k < 16; ++k) fin[k] = g[k];

  KdTree ttree; ttree.init(tgt, m);
  std::vector<Corr> corr;
  int iter = 0, reason = 0, converged = 0;
  double prev_mse = DBL_MAX, cur_mse = 0;
  int nlog = 0;
  unsigned long long queries = 0;
  const double rot_thr = 1.0 - prm->transformation_epsilon, trans_thr = prm->transformation_epsilon;
  const int min_corr = prm->min_correspondences > 0 ? prm->min_correspondences : 3;
  if (prm->max_iterations <= 0) { reason = 1; converged = 1; }
  while (!converged) {
    correspondences_impl(cur.data(), n, tgt, ttree, prm->max_correspondence_distance, prm->use_reciprocal, corr, &queries);
    if ((int)corr.size() < min_corr) { reason = 5; converged = 0; break; }
    double T[16];
    if (prm->estimator == 1) {
      if (!estimate_p2plane_impl(cur.data(), tgt, (const P4*)tgt_normals, corr, T)) { reason = 5; break; }
    } else {
      estimate_svd_impl(cur.data(), tgt, corr, T);
    }
    float Tf[16];
    for (int k = 0; k < 16; ++k) Tf[k] = (float)T[k];
    Tf[3] = Tf[7] = Tf[11] = 0.f; Tf[15] = 1.f;
    for (int i = 0; i < n; ++i) if (finite3(cur[i])) cur[i] = xform(Tf, cur[i]);
    double Td[16];
    for (int k = 0; k < 16; ++k) Td[k] = Tf[k];
    matmul4d(Td, fin, fin);
    ++iter;
    double s = 0; for (const Corr& c : corr) s += c.d2;
    cur_mse = s / (double)corr.size();
    if (log && nlog < max_log) {
      log[nlog].iteration = iter; log[nlog].n_corr = (int)corr.size(); log[nlog].mse = cur_mse;
      for (int k = 0; k < 16; ++k) log[nlog].delta[k] = Td[k];
      ++nlog;
    }
    // DefaultConvergenceCriteria::hasConverged
    if (iter >= prm->max_iterations) { converged = 1; reason = 1; break; }
    if (!prm->fixed_iterations) {
      double cos_angle = 0.5 * (Td[0] + Td[5] + Td[10] - 1.0);
      double t2 = Td[12] * Td[12] + Td[13] * Td[13] + Td[14] * Td[14];
      if (cos_angle >= rot_thr && t2 <= trans_thr) { converged = 1; reason = 2; break; }
      if (std::fabs(cur_mse - prev_mse) < 1e-12) { converged = 1; reason = 3; break; }
      if (std::fabs(cur_mse - prev_mse) / prev_mse < prm->euclidean_fitness_epsilon) { converged = 1; reason = 4; break; }
      prev_mse = cur_mse;
    }
  }
  float finf[16];
  for (int k = 0; k < 16; ++k) finf[k] = (float)fin[k];
  if (final_out) std::memcpy(final_out, finf, sizeof(finf));
  if (out_xyzw) {
    P4* o = (P4*)out_xyzw;
    for (int i = 0; i < n; ++i) o[i] = finite3(src0[i]) ? xform(finf, src0[i]) : src0[i];
  }
  if (rep) { rep->iterations = iter; rep->converged = converged; rep->reason = reason; rep->n_correspondences = (int)corr.size(); rep->mse = cur_mse; rep->nn_queries = queries; }
  if (n_log) *n_log = nlog;
  return reason == 5 ? 2 : 0;
}

// Registration::getFitnessScore(max_range): mean un-gated squared NN distance of cloud -> target.
double orc_fitness_score(const float* cloud, int n, const float* tgt, int m, double max_range) {
  KdTree t; t.init((const P4*)tgt, m);
  std::vector<int> idx(n); std::vector<float> d2(n);
  orc_kdtree_query(&t, cloud, n, idx.data(), d2.data());
  double s = 0; long cnt = 0;
  for (int i = 0; i < n; ++i) if (idx[i] >= 0 && (double)d2[i] <= max_range) { s += d2[i]; ++cnt; }
  return cnt ? s / (double)cnt : DBL_MAX;
}

// ---- uniform-grid index restatement ------------------------------------
// key = 3*bits-bit Morton code of the clamped cell coordinates; non-finite points get the
// sentinel key 1 << (3*bits) so that a stable sort moves them behind every real cell.
void orc_morton_keys(const float* pts, int n, const float* origin, float inv_cell, int bits, uint32_t* keys) {
  const P4* p = (const P4*)pts; const int G = 1 << bits;
  for (int i = 0; i < n; ++i) {
    if (!finite3(p[i])) { keys[i] = 1u << (3 * bits); continue; }
    uint32_t cx = cell_coord(p[i].x, origin[0], inv_cell, G), cy = cell_coord(p[i].y, origin[1], inv_cell, G),
             cz = cell_coord(p[i].z, origin[2], inv_cell, G);
    keys[i] = part1by2(cx) | (part1by2(cy) << 1) | (part1by2(cz) << 2);
  }
}

void orc_stable_sort_perm(const uint32_t* keys, int n, int32_t* perm) {
  std::iota(perm, perm + n, 0);
  std::stable_sort(perm, perm + n, [&](int a, int b) { return keys[a] < keys[b]; });
}

// start has (1<<3*bits)+1 entries: start[c] = first sorted position whose key >= c.
void orc_cell_table(const uint32_t* sorted_keys, int n, int bits, uint32_t* start) {
  const uint32_t C = 1u << (3 * bits);
  int pos = 0;
  for (uint32_t c = 0; c <= C; ++c) {
    while (pos < n && sorted_keys[pos] < c) ++pos;
    start[c] = (uint32_t)pos;
  }
}

// pcl::NormalEstimation semantics: kNN (self included, ties -> lowest index),
// covariance from the k neighbours, eigenvector of the smallest eigenvalue, curvature =
// l0/(l0+l1+l2), flipped towards the viewpoint.  out = n x (nx, ny, nz, curvature).
// nbr_out (nullable) receives the k neighbour indices per point, ascending (d2, idx).
void orc_estimate_normals(const float* pts, int n, int k, const float* viewpoint, float* out, int* nbr_out) {
  const P4* p = (const P4*)pts;
  KdTree t; t.init(p, n);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < n; ++i) {
    std::vector<float> hd(k); std::vector<int> hi(k);
    KnnHeap h{k, 0, hd.data(), hi.data()};
    if (finite3(p[i]) && !t.pts.empty()) knn_search(t, 0, p[i], h);
    if (nbr_out) for (int a = 0; a < k; ++a) nbr_out[(size_t)i * k + a] = a < h.n ? hi[a] : -1;
    if (h.n < 3) { out[4 * i] = out[4 * i + 1] = out[4 * i + 2] = out[4 * i + 3] = NAN; continue; }
    double c[3] = {0, 0, 0};
    for (int a = 0; a < h.n; ++a) { c[0] += p[hi[a]].x; c[1] += p[hi[a]].y; c[2] += p[hi[a]].z; }
    for (int a = 0; a < 3; ++a) c[a] /= h.n;
    double C[3][3] = {};
    for (int a = 0; a < h.n; ++a) {
      double d[3] = {p[hi[a]].x - c[0], p[hi[a]].y - c[1], p[hi[a]].z - c[2]};
      for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) C[r][s] += d[r] * d[s];
    }
    for (int r = 0; r < 3; ++r) for (int s = 0; s < 3; ++s) C[r][s] /= h.n;
    double w[3], V[3][3];
    eig_sym3(C, w, V);
    double nx = V[0][0], ny = V[1][0], nz = V[2][0];
    double vx = viewpoint[0] - p[i].x, vy = viewpoint[1] - p[i].y, vz = viewpoint[2] - p[i].z;
    if (nx * vx + ny * vy + nz * vz < 0) { nx = -nx; ny = -ny; nz = -nz; }
    double tr = w[0] + w[1] + w[2];
    out[4 * i] = (float)nx; out[4 * i + 1] = (float)ny; out[4 * i + 2] = (float)nz;
    out[4 * i + 3] = (float)(tr > 0 ? std::fabs(w[0] / tr) : 0.0);
  }
}

}  // extern "C"
