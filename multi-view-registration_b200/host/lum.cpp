// lum.cpp -- host-side global adjustment of a turntable ring and the turntable axis fit.
//
// ringClose: the loop-closure step after the (possibly multi-GPU) pairwise aligns.  The reference's precedent is
// pcl::registration::LUM over the ring graph i -> i+1, 11 -> 0 (mvr/src/registrator.cpp:627-663): vertices are
// 6-DoF view poses with vertex 0 fixed, every edge carries a relative-pose constraint with an information
// weight, and the poses are relaxed by solving the dense 6(V-1) normal equations a fixed number of times.
// Here an edge constraint is the relative pose measured by the pair's ICP, weighted by its correspondence count
// (isotropic information), and the relaxation is Gauss-Newton on SE(3) -- the same graph, unknowns and solve
// size as LUM; the per-correspondence 6x6 information of PCL's computeEdge is a listed next step.
//
// refineAxisFromPoses: Registrator::refineAxis (mvr/src/registrator.cpp:402-455) with math_solvers::least_squares
// (mvr/src/math_solvers.cpp:24-39; LAPACK dgels there, Householder QR here).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "registrator.h"

namespace mvr {

namespace {

struct Vec6 { double v[6]; };   // (omega, upsilon): rotation vector, translation part

void mat3_mul(const double* A, const double* B, double* C) {   // row-major 3x3
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
  std::memcpy(C, t, sizeof(t));
}

void skew(const double* w, double* S) {
  S[0] = 0; S[1] = -w[2]; S[2] = w[1];
  S[3] = w[2]; S[4] = 0; S[5] = -w[0];
  S[6] = -w[1]; S[7] = w[0]; S[8] = 0;
}

void get_Rt(const Matrix4d& T, double* R, double* t) {   // R row-major
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[i * 3 + j] = T.m[j * 4 + i];
    t[i] = T.m[12 + i];
  }
}

Matrix4d from_Rt(const double* R, const double* t) {
  Matrix4d T = identity4d();
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T.m[j * 4 + i] = R[i * 3 + j];
    T.m[12 + i] = t[i];
  }
  return T;
}

// SO(3) logarithm, robust near 0 and near pi.
void so3_log(const double* R, double* w) {
  const double tr = R[0] + R[4] + R[8];
  double c = 0.5 * (tr - 1.0);
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  const double ax[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};   // 2 sin(theta) * axis
  const double s2 = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);   // 2 sin(theta)
  const double theta = std::atan2(0.5 * s2, c);
  if (s2 < 1e-12) {
    if (c > 0) { w[0] = 0.5 * ax[0]; w[1] = 0.5 * ax[1]; w[2] = 0.5 * ax[2]; return; }
    // theta ~ pi: axis from the diagonal of (R + I) / 2
    double d[3] = {std::sqrt(std::fmax(0.0, 0.5 * (R[0] + 1))), std::sqrt(std::fmax(0.0, 0.5 * (R[4] + 1))), std::sqrt(std::fmax(0.0, 0.5 * (R[8] + 1)))};
    int k = d[0] >= d[1] ? (d[0] >= d[2] ? 0 : 2) : (d[1] >= d[2] ? 1 : 2);
    double a[3];
    a[k] = d[k];
    for (int j = 0; j < 3; ++j) if (j != k) a[j] = 0.25 * (R[k * 3 + j] + R[j * 3 + k]) / d[k];
    for (int j = 0; j < 3; ++j) w[j] = theta * a[j];
    return;
  }
  const double f = theta / s2;
  w[0] = f * ax[0]; w[1] = f * ax[1]; w[2] = f * ax[2];
}

void so3_exp(const double* w, double* R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(th2);
  double A, B;
  if (th < 1e-6) { A = 1.0 - th2 / 6.0; B = 0.5 - th2 / 24.0; }
  else { A = std::sin(th) / th; B = (1.0 - std::cos(th)) / th2; }
  double S[9], S2[9];
  skew(w, S);
  mat3_mul(S, S, S2);
  for (int k = 0; k < 9; ++k) R[k] = ((k % 4 == 0) ? 1.0 : 0.0) + A * S[k] + B * S2[k];
}

// SE(3) log with the exact left-Jacobian inverse on the translation.
Vec6 se3_log(const Matrix4d& T) {
  double R[9], t[3], w[3];
  get_Rt(T, R, t);
  so3_log(R, w);
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(th2);
  double S[9], S2[9];
  skew(w, S);
  mat3_mul(S, S, S2);
  double cf;
  if (th < 1e-6) cf = 1.0 / 12.0 + th2 / 720.0;
  else cf = (1.0 - 0.5 * th * std::sin(th) / (1.0 - std::cos(th))) / th2;
  Vec6 r;
  for (int i = 0; i < 3; ++i) {
    r.v[i] = w[i];
    double u = 0;
    for (int j = 0; j < 3; ++j) u += (((i == j) ? 1.0 : 0.0) - 0.5 * S[i * 3 + j] + cf * S2[i * 3 + j]) * t[j];
    r.v[3 + i] = u;
  }
  return r;
}

Matrix4d se3_exp(const Vec6& x) {
  const double* w = x.v;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(th2);
  double R[9], S[9], S2[9];
  so3_exp(w, R);
  skew(w, S);
  mat3_mul(S, S, S2);
  double B, C;
  if (th < 1e-6) { B = 0.5 - th2 / 24.0; C = 1.0 / 6.0 - th2 / 120.0; }
  else { B = (1.0 - std::cos(th)) / th2; C = (th - std::sin(th)) / (th2 * th); }
  double t[3];
  for (int i = 0; i < 3; ++i) {
    double u = 0;
    for (int j = 0; j < 3; ++j) u += (((i == j) ? 1.0 : 0.0) + B * S[i * 3 + j] + C * S2[i * 3 + j]) * x.v[3 + j];
    t[i] = u;
  }
  return from_Rt(R, t);
}

// 6x6 adjoint of T for twists ordered (omega, upsilon), row-major.
void se3_adjoint(const Matrix4d& T, double* Ad) {
  double R[9], t[3], tx[9], tR[9];
  get_Rt(T, R, t);
  skew(t, tx);
  mat3_mul(tx, R, tR);
  for (int k = 0; k < 36; ++k) Ad[k] = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Ad[i * 6 + j] = R[i * 3 + j];
      Ad[(3 + i) * 6 + 3 + j] = R[i * 3 + j];
      Ad[(3 + i) * 6 + j] = tR[i * 3 + j];
    }
}

// SPD solve (Cholesky), A row-major n x n, overwritten.  band: A(i, j) = 0 for |i - j| > band (n - 1 for a dense matrix);
// the factor of a banded matrix keeps the band, so a chain-structured system costs O(n band^2) instead of O(n^3).
bool chol_solve(std::vector<double>& A, std::vector<double>& b, int n, int band) {
  if (band < 0 || band > n - 1) band = n - 1;
  for (int i = 0; i < n; ++i)
    for (int j = std::max(0, i - band); j <= i; ++j) {
      double s = A[(size_t)i * n + j];
      for (int k = std::max(0, i - band); k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
      if (i == j) {
        if (!(s > 0)) return false;
        A[(size_t)i * n + i] = std::sqrt(s);
      } else {
        A[(size_t)i * n + j] = s / A[(size_t)j * n + j];
      }
    }
  for (int i = 0; i < n; ++i) {
    double s = b[(size_t)i];
    for (int k = std::max(0, i - band); k < i; ++k) s -= A[(size_t)i * n + k] * b[(size_t)k];
    b[(size_t)i] = s / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[(size_t)i];
    for (int k = i + 1; k <= std::min(n - 1, i + band); ++k) s -= A[(size_t)k * n + i] * b[(size_t)k];
    b[(size_t)i] = s / A[(size_t)i * n + i];
  }
  return true;
}

}  // namespace

// Absolute poses X_v (X_0 = I) of a ring from relative measurements rel[p] ~ X_p^-1 X_{(p+1) % V}.
// relax = false: plain chaining of pairs 0 .. V-2 (the closing pair is ignored).
// relax = true : Gauss-Newton on  sum_p w_p | Log(rel[p]^-1 X_p^-1 X_{p+1}) |^2, X_v <- X_v Exp(d_v),
//                 `iterations` sweeps of the dense 6(V-1) solve (LUM runs 16 of them).
//                 The residual twist is taken about `centre` (the turntable pivot: the object sits there in every
//                 view's frame, so the poses are nearly pure rotations about it) and its rotation part is scaled
//                 by `rot_scale` (the object's radius), which makes the cost the mean squared displacement of the
//                 object's points -- what LUM's per-correspondence information encodes -- instead of mixing
//                 radians and millimetres measured at the sensor origin.
int ringClose(const std::vector<Matrix4d>& rel_in, const std::vector<double>& weight, bool relax, int iterations,
              const double* centre, double rot_scale, std::vector<Matrix4d>& X) {
  const int V = (int)rel_in.size();
  X.assign((size_t)std::max(V, 1), identity4d());
  if (V < 2) return MVR_OK;
  if (!relax) {
    for (int p = 0; p + 1 < V; ++p) X[(size_t)p + 1] = multiply(X[(size_t)p], rel_in[(size_t)p]);
    return MVR_OK;
  }
  // work in coordinates centred on the pivot: T' = S^-1 T S with S = translate(centre)
  Matrix4d S = identity4d(), Si = identity4d();
  if (centre) for (int k = 0; k < 3; ++k) { S.m[12 + k] = centre[k]; Si.m[12 + k] = -centre[k]; }
  std::vector<Matrix4d> rel((size_t)V);
  for (int p = 0; p < V; ++p) rel[(size_t)p] = multiply(multiply(Si, rel_in[(size_t)p]), S);
  const double rho2 = (rot_scale > 0) ? rot_scale * rot_scale : 1.0;
  const double Wd[6] = {rho2, rho2, rho2, 1.0, 1.0, 1.0};
  for (int p = 0; p + 1 < V; ++p) X[(size_t)p + 1] = multiply(X[(size_t)p], rel[(size_t)p]);
  double wmax = 0;
  for (double w : weight) wmax = std::fmax(wmax, w);
  if (!(wmax > 0)) return MVR_OK;
  const int n = 6 * (V - 1);
  // With vertex 0 fixed the ring's unknowns 1 .. V-1 form a CHAIN (the edges 0 -> 1 and V-1 -> 0 touch one unknown
  // each): the normal equations are block tridiagonal with 6 x 6 blocks, half bandwidth 11.
  const int band = 11;
  // (dense row-major storage, but only the band is ever cleared, copied or read: 23 of 138 columns per row for 24 views)
  std::vector<double> H((size_t)n * n, 0.0), g((size_t)n, 0.0), Hd((size_t)n * n, 0.0), gd;
  for (int it = 0; it < iterations; ++it) {
    for (int i = 0; i < n; ++i) std::fill(H.begin() + (size_t)i * n + std::max(0, i - band), H.begin() + (size_t)i * n + std::min(n - 1, i + band) + 1, 0.0);
    std::fill(g.begin(), g.end(), 0.0);
    double cost = 0;
    for (int p = 0; p < V; ++p) {
      const double w = weight[(size_t)p] / wmax;
      if (!(w > 0)) continue;
      const int a = p, b = (p + 1) % V;
      // E = rel^-1 X_a^-1 X_b ; residual r = Log(E).  With X <- X Exp(d):
      //   dr/dd_b ~ I,  dr/dd_a ~ -Ad(E^-1 ... ) ~ -Ad(X_b^-1 X_a)   (first order, exact at r = 0)
      const Matrix4d XaiXb = multiply(inverseRigid(X[(size_t)a]), X[(size_t)b]);
      const Matrix4d E = multiply(inverseRigid(rel[(size_t)p]), XaiXb);
      const Vec6 r = se3_log(E);
      double Ja[36];
      se3_adjoint(inverseRigid(XaiXb), Ja);
      for (int k = 0; k < 36; ++k) Ja[k] = -Ja[k];
      for (int k = 0; k < 6; ++k) cost += w * Wd[k] * r.v[k] * r.v[k];
      const int ia = 6 * (a - 1), ib = 6 * (b - 1);   // vertex 0 is fixed (no unknowns)
      // H += J^T w J ; g += J^T w r   with J = [Ja at a, I at b]
      if (a > 0) {
        for (int i = 0; i < 6; ++i) {
          double gi = 0;
          for (int k = 0; k < 6; ++k) gi += Ja[k * 6 + i] * Wd[k] * r.v[k];
          g[(size_t)ia + i] += w * gi;
          for (int j = 0; j < 6; ++j) {
            double h = 0;
            for (int k = 0; k < 6; ++k) h += Ja[k * 6 + i] * Wd[k] * Ja[k * 6 + j];
            H[(size_t)(ia + i) * n + ia + j] += w * h;
          }
        }
      }
      if (b > 0) {
        for (int i = 0; i < 6; ++i) {
          g[(size_t)ib + i] += w * Wd[i] * r.v[i];
          H[(size_t)(ib + i) * n + ib + i] += w * Wd[i];
        }
      }
      if (a > 0 && b > 0) {
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) {
            H[(size_t)(ia + i) * n + ib + j] += w * Wd[j] * Ja[j * 6 + i];   // Ja^T W I
            H[(size_t)(ib + j) * n + ia + i] += w * Wd[j] * Ja[j * 6 + i];
          }
      }
    }
    // Levenberg damping, raised only if the plain system is not numerically positive definite (edges dropped for
    // lack of correspondences can leave a vertex without information: it then simply keeps its pose).
    double dmax = 0;
    for (int i = 0; i < n; ++i) { dmax = std::fmax(dmax, H[(size_t)i * n + i]); g[(size_t)i] = -g[(size_t)i]; }
    if (!(dmax > 0) || !std::isfinite(dmax)) break;
    bool solved = false;
    for (double lambda = 1e-12 * dmax; lambda <= 1e3 * dmax; lambda *= 1e3) {
      gd = g;
      for (int i = 0; i < n; ++i) {   // the factorisation reads the lower band only
        const size_t lo = (size_t)i * n + std::max(0, i - band), hi = (size_t)i * n + i + 1;
        std::copy(H.begin() + lo, H.begin() + hi, Hd.begin() + lo);
        Hd[(size_t)i * n + i] += lambda;
      }
      if (chol_solve(Hd, gd, n, band)) { g.swap(gd); solved = true; break; }
    }
    if (!solved) return MVR_ERR_NOT_SPD;
    double step = 0;
    for (int v = 1; v < V; ++v) {
      Vec6 d;
      for (int k = 0; k < 6; ++k) { d.v[k] = g[(size_t)6 * (v - 1) + k]; step = std::fmax(step, std::fabs(d.v[k])); }
      X[(size_t)v] = multiply(X[(size_t)v], se3_exp(d));
    }
    if (step < 1e-12 || cost < 1e-30) break;   // converged (rotations in rad, translations in mm): further sweeps change nothing
  }
  for (int v = 0; v < V; ++v) X[(size_t)v] = multiply(multiply(S, X[(size_t)v]), Si);
  return MVR_OK;
}

// ---------------------------------------------------------------------------------------------------------
// LUM relaxation on correspondence moments.
//
// pcl::registration::LUM (the reference's global adjustment, mvr/src/registrator.cpp:627-663) gives every edge
// of the pose graph a 6x6 information matrix and a 6-vector built from the edge's point correspondences
// (computeEdge: midpoints and differences of the pairs), then solves the dense 6 (V - 1) system `max_iterations`
// times, re-deriving the edges from the SAME correspondences after every pose update.  The cost that procedure
// descends is  sum_edges sum_k |X_s a_k - X_t b_k|^2.  For rigid X that cost -- and its Gauss-Newton model at any
// X -- is a function of the pairs' first and second moments only (mvr_pair_moments, one GPU reduction per edge),
// so the sweeps below run on 30 doubles per edge and never touch a point.
// ---------------------------------------------------------------------------------------------------------
namespace {

struct Sym3 { double xx, xy, xz, yy, yz, zz; };

void sym_to_full(const double* s6, double* F) {   // xx xy xz yy yz zz -> row-major 3x3
  F[0] = s6[0]; F[1] = s6[1]; F[2] = s6[2];
  F[3] = s6[1]; F[4] = s6[3]; F[5] = s6[4];
  F[6] = s6[2]; F[7] = s6[4]; F[8] = s6[5];
}

void full_to_sym(const double* F, double* s6) {
  s6[0] = F[0]; s6[1] = 0.5 * (F[1] + F[3]); s6[2] = 0.5 * (F[2] + F[6]);
  s6[3] = F[4]; s6[4] = 0.5 * (F[5] + F[7]); s6[5] = F[8];
}

void mat3_T(const double* A, double* At) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) At[i * 3 + j] = A[j * 3 + i];
}

// out = Ra S Rb^T + (Ra u) cb^T + ca (Rb v)^T + n ca cb^T   -- the second moment sum (Ra x + ca)(Rb y + cb)^T
// given S = sum x y^T, u = sum x, v = sum y.
void moment2(const double* Ra, const double* ca, const double* Rb, const double* cb, const double* S, const double* u,
             const double* v, double n, double* out) {
  double RbT[9], T[9];
  mat3_T(Rb, RbT);
  mat3_mul(Ra, S, T);
  mat3_mul(T, RbT, out);
  double Rau[3], Rbv[3];
  for (int i = 0; i < 3; ++i) {
    Rau[i] = Ra[i * 3] * u[0] + Ra[i * 3 + 1] * u[1] + Ra[i * 3 + 2] * u[2];
    Rbv[i] = Rb[i * 3] * v[0] + Rb[i * 3 + 1] * v[1] + Rb[i * 3 + 2] * v[2];
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[i * 3 + j] += Rau[i] * cb[j] + ca[i] * Rbv[j] + n * ca[i] * cb[j];
}

// Moments of the pairs (Xa a, Xb b) about new_origin, from those of (a, b) about in.origin.
void moments_apply(const mvr_pair_moments& in, const Matrix4d& Xa, const Matrix4d& Xb, const double* new_origin, mvr_pair_moments& out) {
  double Ra[9], ta[3], Rb[9], tb[3], ca[3], cb[3];
  get_Rt(Xa, Ra, ta);
  get_Rt(Xb, Rb, tb);
  for (int i = 0; i < 3; ++i) {
    ca[i] = Ra[i * 3] * in.origin[0] + Ra[i * 3 + 1] * in.origin[1] + Ra[i * 3 + 2] * in.origin[2] + ta[i] - new_origin[i];
    cb[i] = Rb[i * 3] * in.origin[0] + Rb[i * 3 + 1] * in.origin[1] + Rb[i * 3 + 2] * in.origin[2] + tb[i] - new_origin[i];
  }
  mvr_pair_moments o;
  o.n = in.n;
  for (int i = 0; i < 3; ++i) {
    o.origin[i] = new_origin[i];
    o.sa[i] = Ra[i * 3] * in.sa[0] + Ra[i * 3 + 1] * in.sa[1] + Ra[i * 3 + 2] * in.sa[2] + in.n * ca[i];
    o.sb[i] = Rb[i * 3] * in.sb[0] + Rb[i * 3 + 1] * in.sb[1] + Rb[i * 3 + 2] * in.sb[2] + in.n * cb[i];
  }
  double Saa[9], Sbb[9], F[9];
  sym_to_full(in.saa, Saa);
  sym_to_full(in.sbb, Sbb);
  moment2(Ra, ca, Ra, ca, Saa, in.sa, in.sa, in.n, F);
  full_to_sym(F, o.saa);
  moment2(Rb, cb, Rb, cb, Sbb, in.sb, in.sb, in.n, F);
  full_to_sym(F, o.sbb);
  moment2(Rb, cb, Ra, ca, in.sba, in.sb, in.sa, in.n, o.sba);
  o.d2 = (o.saa[0] + o.saa[3] + o.saa[5]) + (o.sbb[0] + o.sbb[3] + o.sbb[5]) - 2.0 * (o.sba[0] + o.sba[4] + o.sba[8]);
  out = o;
}

// Gauss-Newton model of one edge at the identity:  sum_k |delta_k + omega x m_k + upsilon|^2 = d^T M d + 2 v^T d + c
// with m = (a + b) / 2, delta = a - b, d = (omega, upsilon) = d_source - d_target, twists about the moments' origin.
void edge_model(const mvr_pair_moments& e, double* M /* 36 */, double* v /* 6 */) {
  double Saa[9], Sbb[9], Smm[9];
  sym_to_full(e.saa, Saa);
  sym_to_full(e.sbb, Sbb);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Smm[i * 3 + j] = 0.25 * (Saa[i * 3 + j] + Sbb[i * 3 + j] + e.sba[i * 3 + j] + e.sba[j * 3 + i]);
  const double tr = Smm[0] + Smm[4] + Smm[8];
  const double sm[3] = {0.5 * (e.sa[0] + e.sb[0]), 0.5 * (e.sa[1] + e.sb[1]), 0.5 * (e.sa[2] + e.sb[2])};
  double K[9];
  skew(sm, K);   // sum [m]x
  for (int k = 0; k < 36; ++k) M[k] = 0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      M[i * 6 + j] = ((i == j) ? tr : 0.0) - Smm[i * 3 + j];   // sum |m|^2 I - m m^T
      M[i * 6 + 3 + j] = K[i * 3 + j];                        // sum [m]x
      M[(3 + i) * 6 + j] = -K[i * 3 + j];                     // its transpose
      M[(3 + i) * 6 + 3 + j] = (i == j) ? e.n : 0.0;
    }
  // sum m x delta = sum b x a ; sum delta = sa - sb
  v[0] = e.sba[1 * 3 + 2] - e.sba[2 * 3 + 1];
  v[1] = e.sba[2 * 3 + 0] - e.sba[0 * 3 + 2];
  v[2] = e.sba[0 * 3 + 1] - e.sba[1 * 3 + 0];
  for (int i = 0; i < 3; ++i) v[3 + i] = e.sa[i] - e.sb[i];
}

double graph_cost(const std::vector<mvr_pair_moments>& base, const int* src, const int* tgt, const std::vector<Matrix4d>& X,
                  const double* origin, std::vector<mvr_pair_moments>* moved) {
  double c = 0;
  if (moved) moved->resize(base.size());
  for (size_t e = 0; e < base.size(); ++e) {
    mvr_pair_moments m;
    moments_apply(base[e], X[(size_t)src[e]], X[(size_t)tgt[e]], origin, m);
    c += m.d2;
    if (moved) (*moved)[e] = m;
  }
  return c;
}

}  // namespace

void momentsTransform(const mvr_pair_moments& in, const Matrix4d& pose, const double* new_origin, mvr_pair_moments& out) {
  double o[3];
  if (new_origin) {
    for (int i = 0; i < 3; ++i) o[i] = new_origin[i];
  } else {
    for (int i = 0; i < 3; ++i) o[i] = pose.m[i] * in.origin[0] + pose.m[4 + i] * in.origin[1] + pose.m[8 + i] * in.origin[2] + pose.m[12 + i];
  }
  moments_apply(in, pose, pose, o, out);
  out.d2 = in.d2;   // a common rigid motion keeps every distance: carry the measured sum
}

// ---------------------------------------------------------------------------------------------------------
// pcl::registration::LUM::compute() itself, on moments.
//
// What the reference runs (mvr/src/registrator.cpp:627-663; PCL's registration/impl/lum.hpp): vertices carry a
// 6-vector pose (x, y, z, roll, pitch, yaw), vertex 0 fixed; per sweep every edge is re-linearised (computeEdge) with both
// clouds compounded onto their current poses -- per correspondence the averaged point m = (s + t) / 2 and the difference
// d = s - t enter M'M = sum M_k^T M_k and M'Z = sum M_k^T d_k, M_k = [ I | e_x x m, e_z x m, e_y x m ], then
// D = (M'M)^-1 M'Z, s^2 = sum |d_k - M_k D|^2, cinv = M'M / s^2, cinvd = M'Z / s^2 --, the dense system G X = B over the
// 6 (V - 1) unknowns is solved and every pose moves by -incidenceCorrection(pose)^-1 X_v.
// Every sum computeEdge takes over the correspondences is a linear function of the pairs' first and second moments:
//   sum m = (Sa + Sb) / 2,  sum m m^T = (Saa + Sab + Sba + Sbb) / 4,  sum d = Sa - Sb,  sum m x d = sum t x s (the
//   antisymmetric part of Sba),  sum |d|^2 = tr Saa - 2 tr Sba + tr Sbb,  s^2 = sum |d|^2 - D . M'Z,
// and the moments of the COMPOUNDED pairs follow from those of the raw pairs (moments_apply): the sweeps below are PCL's,
// on 30 doubles per edge (GPU-reduced, mvr_pair_moments) instead of the point lists.  PCL runs them in float32; here
// double.  oracle/lum_oracle.py restates the same procedure on the points themselves.
// ---------------------------------------------------------------------------------------------------------
namespace {

// pcl::getTransformation(x, y, z, roll, pitch, yaw) = Translation * Rz(yaw) * Ry(pitch) * Rx(roll)
Matrix4d pcl_transformation(const double* p) {
  const double cr = std::cos(p[3]), sr = std::sin(p[3]), cp = std::cos(p[4]), sp = std::sin(p[4]), cy = std::cos(p[5]), sy = std::sin(p[5]);
  const double R[9] = {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
                       sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
                       -sp,     cp * sr,                cp * cr};
  return from_Rt(R, p);
}

// LUM::incidenceCorrection (row-major 6 x 6)
void incidence_correction(const double* p, double* H) {
  const double cx = std::cos(p[3]), sx = std::sin(p[3]), cy = std::cos(p[4]), sy = std::sin(p[4]);
  for (int k = 0; k < 36; ++k) H[k] = (k % 7 == 0) ? 1.0 : 0.0;
  H[0 * 6 + 4] = p[1] * sx - p[2] * cx;
  H[0 * 6 + 5] = p[1] * cx * cy + p[2] * sx * cy;
  H[1 * 6 + 3] = p[2];
  H[1 * 6 + 4] = -p[0] * sx;
  H[1 * 6 + 5] = -p[0] * cx * cy + p[2] * sy;
  H[2 * 6 + 3] = -p[1];
  H[2 * 6 + 4] = p[0] * cx;
  H[2 * 6 + 5] = -p[0] * sx * cy - p[1] * sy;
  H[3 * 6 + 5] = sy;
  H[4 * 6 + 4] = sx;
  H[4 * 6 + 5] = cx * cy;
  H[5 * 6 + 4] = cx;
  H[5 * 6 + 5] = -sx * cy;
}

// Gaussian elimination with partial pivoting, A (n x n row-major) and b overwritten; false if singular.
bool lu_solve(double* A, double* b, int n) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    for (int r = c + 1; r < n; ++r) if (std::fabs(A[r * n + c]) > std::fabs(A[piv * n + c])) piv = r;
    if (!(std::fabs(A[piv * n + c]) > 0)) return false;
    if (piv != c) { for (int k = 0; k < n; ++k) std::swap(A[c * n + k], A[piv * n + k]); std::swap(b[c], b[piv]); }
    for (int r = c + 1; r < n; ++r) {
      const double f = A[r * n + c] / A[c * n + c];
      if (f == 0) continue;
      for (int k = c; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double x = b[r];
    for (int k = r + 1; k < n; ++k) x -= A[r * n + k] * b[k];
    b[r] = x / A[r * n + r];
  }
  return true;
}

// computeEdge from the moments of the compounded pairs (absolute coordinates: origin 0).  false: the edge carries no information.
bool pcl_edge(const mvr_pair_moments& m, double* cinv /* 36 */, double* cinvd /* 6 */) {
  for (int k = 0; k < 36; ++k) cinv[k] = 0;
  for (int k = 0; k < 6; ++k) cinvd[k] = 0;
  if (!(m.n >= 3)) return false;
  double Saa[9], Sbb[9], Smm[9];
  sym_to_full(m.saa, Saa);
  sym_to_full(m.sbb, Sbb);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Smm[i * 3 + j] = 0.25 * (Saa[i * 3 + j] + Sbb[i * 3 + j] + m.sba[i * 3 + j] + m.sba[j * 3 + i]);
  const double mx = 0.5 * (m.sa[0] + m.sb[0]), my = 0.5 * (m.sa[1] + m.sb[1]), mz = 0.5 * (m.sa[2] + m.sb[2]);
  double MM[36];
  for (int k = 0; k < 36; ++k) MM[k] = 0;
  MM[0 * 6 + 4] = -my; MM[0 * 6 + 5] = mz;
  MM[1 * 6 + 3] = -mz; MM[1 * 6 + 4] = mx;
  MM[2 * 6 + 3] = my;  MM[2 * 6 + 5] = -mx;
  MM[3 * 6 + 4] = -Smm[0 * 3 + 2]; MM[3 * 6 + 5] = -Smm[0 * 3 + 1]; MM[4 * 6 + 5] = -Smm[1 * 3 + 2];
  MM[3 * 6 + 3] = Smm[4] + Smm[8]; MM[4 * 6 + 4] = Smm[0] + Smm[4]; MM[5 * 6 + 5] = Smm[0] + Smm[8];
  MM[0] = MM[7] = MM[14] = m.n;
  for (int i = 0; i < 6; ++i)
    for (int j = i + 1; j < 6; ++j) MM[j * 6 + i] = MM[i * 6 + j];
  // sum m x d = sum t x s: component x = sum (t_y s_z - t_z s_y), Sba[r][c] = sum t_r s_c
  const double MZ[6] = {m.sa[0] - m.sb[0], m.sa[1] - m.sb[1], m.sa[2] - m.sb[2],
                        m.sba[1 * 3 + 2] - m.sba[2 * 3 + 1],    // sum (m_y d_z - m_z d_y)
                        m.sba[0 * 3 + 1] - m.sba[1 * 3 + 0],    // sum (m_x d_y - m_y d_x)
                        m.sba[2 * 3 + 0] - m.sba[0 * 3 + 2]};   // sum (m_z d_x - m_x d_z)
  double A[36], D[6];
  std::memcpy(A, MM, sizeof(A));
  std::memcpy(D, MZ, sizeof(D));
  if (!lu_solve(A, D, 6)) return false;
  const double dd = (m.saa[0] + m.saa[3] + m.saa[5]) + (m.sbb[0] + m.sbb[3] + m.sbb[5]) - 2.0 * (m.sba[0] + m.sba[4] + m.sba[8]);
  double ss = dd;
  for (int k = 0; k < 6; ++k) ss -= D[k] * MZ[k];
  if (!(ss >= 0.0000000000001) || !std::isfinite(ss)) return false;   // PCL: "limitations of computation due to linearization"
  for (int k = 0; k < 36; ++k) cinv[k] = MM[k] / ss;
  for (int k = 0; k < 6; ++k) cinvd[k] = MZ[k] / ss;
  return true;
}

}  // namespace

int lumComputePcl(const std::vector<mvr_pair_moments>& edges, const int* src, const int* tgt, int V, int iterations, double convergence_threshold,
                  std::vector<double>& poses6, std::vector<Matrix4d>& X) {
  poses6.assign((size_t)std::max(V, 1) * 6, 0.0);
  X.assign((size_t)std::max(V, 1), identity4d());
  const int E = (int)edges.size();
  if (V < 2) return MVR_OK;
  for (int e = 0; e < E; ++e)
    if (src[e] < 0 || src[e] >= V || tgt[e] < 0 || tgt[e] >= V || src[e] == tgt[e]) return MVR_ERR_BAD_ARG;
  const int n = 6 * (V - 1);
  const double zero[3] = {0, 0, 0};
  std::vector<double> G((size_t)n * n), B((size_t)n), cinv((size_t)E * 36), cinvd((size_t)E * 6);
  for (int it = 0; it < iterations; ++it) {
    for (int e = 0; e < E; ++e) {
      mvr_pair_moments m;
      moments_apply(edges[(size_t)e], pcl_transformation(&poses6[(size_t)src[e] * 6]), pcl_transformation(&poses6[(size_t)tgt[e] * 6]), zero, m);
      pcl_edge(m, &cinv[(size_t)e * 36], &cinvd[(size_t)e * 6]);
    }
    std::fill(G.begin(), G.end(), 0.0);
    std::fill(B.begin(), B.end(), 0.0);
    // PCL walks vertex pairs (vi, vj) and takes the forward edge if there is one, else the backward edge
    for (int vi = 1; vi < V; ++vi)
      for (int vj = 0; vj < V; ++vj) {
        int e_use = -1;
        double sign = 1.0;
        for (int e = 0; e < E && e_use < 0; ++e) if (src[e] == vi && tgt[e] == vj) e_use = e;
        if (e_use < 0) { for (int e = 0; e < E && e_use < 0; ++e) if (src[e] == vj && tgt[e] == vi) e_use = e; sign = -1.0; }
        if (e_use < 0) continue;
        const double* C = &cinv[(size_t)e_use * 36];
        for (int i = 0; i < 6; ++i) {
          for (int j = 0; j < 6; ++j) {
            if (vj > 0) G[(size_t)(6 * (vi - 1) + i) * n + 6 * (vj - 1) + j] = -C[i * 6 + j];
            G[(size_t)(6 * (vi - 1) + i) * n + 6 * (vi - 1) + j] += C[i * 6 + j];
          }
          B[(size_t)6 * (vi - 1) + i] += sign * cinvd[(size_t)e_use * 6 + i];
        }
      }
    // G X = B (PCL: colPivHouseholderQr).  A vertex without information has a zero block: a vanishing ridge keeps it where it is.
    double dmax = 0;
    for (int i = 0; i < n; ++i) dmax = std::fmax(dmax, std::fabs(G[(size_t)i * n + i]));
    if (!(dmax > 0) || !std::isfinite(dmax)) break;
    std::vector<double> A(G), x(B);
    for (int i = 0; i < n; ++i) if (!(A[(size_t)i * n + i] > 1e-14 * dmax)) A[(size_t)i * n + i] += 1e-14 * dmax;
    if (!lu_solve(A.data(), x.data(), n)) return MVR_ERR_NOT_SPD;
    double sum = 0;
    for (int v = 1; v < V; ++v) {
      double H[36], d[6];
      incidence_correction(&poses6[(size_t)v * 6], H);
      for (int k = 0; k < 6; ++k) d[k] = x[(size_t)6 * (v - 1) + k];
      if (!lu_solve(H, d, 6)) return MVR_ERR_NOT_SPD;
      double nrm = 0;
      for (int k = 0; k < 6; ++k) { poses6[(size_t)v * 6 + k] -= d[k]; nrm += d[k] * d[k]; }
      sum += std::sqrt(nrm);
    }
    if (sum <= convergence_threshold * (double)(V - 1)) break;
  }
  for (int v = 0; v < V; ++v) X[(size_t)v] = pcl_transformation(&poses6[(size_t)v * 6]);
  return MVR_OK;
}

int lumRelax(const std::vector<mvr_pair_moments>& edges_in, const int* src, const int* tgt, int V, int iterations, std::vector<Matrix4d>& X) {
  X.assign((size_t)std::max(V, 1), identity4d());
  const int E = (int)edges_in.size();
  if (V < 2 || E == 0) return MVR_OK;
  for (int e = 0; e < E; ++e)
    if (src[e] < 0 || src[e] >= V || tgt[e] < 0 || tgt[e] >= V || src[e] == tgt[e]) return MVR_ERR_BAD_ARG;
  // one common origin for every twist: the correspondence centroid of the whole graph
  double origin[3] = {0, 0, 0}, ntot = 0;
  for (const mvr_pair_moments& e : edges_in) {
    if (!(e.n > 0)) continue;
    for (int i = 0; i < 3; ++i) origin[i] += e.n * e.origin[i] + 0.5 * (e.sa[i] + e.sb[i]);
    ntot += e.n;
  }
  if (!(ntot > 0)) return MVR_OK;
  for (int i = 0; i < 3; ++i) origin[i] /= ntot;
  std::vector<mvr_pair_moments> base;
  std::vector<int> es, et;
  for (int e = 0; e < E; ++e)
    if (edges_in[(size_t)e].n > 0) { base.push_back(edges_in[(size_t)e]); es.push_back(src[e]); et.push_back(tgt[e]); }
  Matrix4d S = identity4d(), Si = identity4d();
  for (int k = 0; k < 3; ++k) { S.m[12 + k] = origin[k]; Si.m[12 + k] = -origin[k]; }
  const int n = 6 * (V - 1);
  // half bandwidth of the normal equations: vertex 0 carries no unknowns, so a ring is a chain (block tridiagonal)
  int band = 5;
  for (size_t e = 0; e < es.size(); ++e)
    if (es[e] > 0 && et[e] > 0) band = std::max(band, 6 * std::abs(es[e] - et[e]) + 5);
  std::vector<mvr_pair_moments> cur;
  double cost = graph_cost(base, es.data(), et.data(), X, origin, &cur);
  double lambda = 0.0;
  for (int it = 0; it < iterations; ++it) {
    std::vector<double> H((size_t)n * n, 0.0), g((size_t)n, 0.0);
    for (size_t e = 0; e < cur.size(); ++e) {
      double M[36], v[6];
      edge_model(cur[e], M, v);
      const int a = es[e], b = et[e];
      const int ia = 6 * (a - 1), ib = 6 * (b - 1);   // vertex 0 is fixed (LUM: the first cloud is the reference)
      for (int i = 0; i < 6; ++i) {
        if (a > 0) g[(size_t)ia + i] += v[i];
        if (b > 0) g[(size_t)ib + i] -= v[i];
        for (int j = 0; j < 6; ++j) {
          const double m = M[i * 6 + j];
          if (a > 0) H[(size_t)(ia + i) * n + ia + j] += m;
          if (b > 0) H[(size_t)(ib + i) * n + ib + j] += m;
          if (a > 0 && b > 0) { H[(size_t)(ia + i) * n + ib + j] -= m; H[(size_t)(ib + i) * n + ia + j] -= m; }
        }
      }
    }
    double dmax = 0;
    for (int i = 0; i < n; ++i) dmax = std::fmax(dmax, H[(size_t)i * n + i]);
    if (!(dmax > 0) || !std::isfinite(dmax)) break;
    // Levenberg-Marquardt safeguard: the plain Gauss-Newton step (lambda = 0) is LUM's; damping is raised only when
    // the system is not positive definite (a view without edges keeps its pose) or the true cost would grow.
    bool accepted = false;
    double step = 0;
    for (int attempt = 0; attempt < 12 && !accepted; ++attempt) {
      std::vector<double> Hd(H), x(g);
      for (int i = 0; i < n; ++i) { Hd[(size_t)i * n + i] += lambda * std::fmax(H[(size_t)i * n + i], 1e-12 * dmax) + 1e-14 * dmax; x[(size_t)i] = -x[(size_t)i]; }
      if (!chol_solve(Hd, x, n, band)) { lambda = lambda > 0 ? lambda * 10.0 : 1e-6; continue; }
      std::vector<Matrix4d> Xn(X);
      step = 0;
      for (int v = 1; v < V; ++v) {
        Vec6 d;
        for (int k = 0; k < 6; ++k) { d.v[k] = x[(size_t)6 * (v - 1) + k]; step = std::fmax(step, std::fabs(d.v[k])); }
        Xn[(size_t)v] = multiply(multiply(multiply(S, se3_exp(d)), Si), X[(size_t)v]);   // twist about the common origin
      }
      std::vector<mvr_pair_moments> moved;
      const double c2 = graph_cost(base, es.data(), et.data(), Xn, origin, &moved);
      if (c2 <= cost * (1.0 + 1e-12) || !(step > 1e-14)) {
        X.swap(Xn); cur.swap(moved); cost = c2; accepted = true;
        lambda *= 0.1;
        if (lambda < 1e-9) lambda = 0.0;
      } else {
        lambda = lambda > 0 ? lambda * 10.0 : 1e-6;
      }
    }
    if (!accepted || step < 1e-13) break;
  }
  return MVR_OK;
}

// min |A x - b|_2 by Householder QR (A rows x cols row-major, rows >= cols, full column rank).
bool leastSquares(const std::vector<double>& A_in, const std::vector<double>& b_in, int rows, int cols, std::vector<double>& x) {
  if (rows < cols || cols <= 0) return false;
  std::vector<double> A(A_in), b(b_in);
  for (int k = 0; k < cols; ++k) {
    double nrm = 0;
    for (int i = k; i < rows; ++i) nrm += A[(size_t)i * cols + k] * A[(size_t)i * cols + k];
    nrm = std::sqrt(nrm);
    if (!(nrm > 0)) return false;
    const double alpha = A[(size_t)k * cols + k] > 0 ? -nrm : nrm;
    std::vector<double> v((size_t)rows, 0.0);
    for (int i = k; i < rows; ++i) v[(size_t)i] = A[(size_t)i * cols + k];
    v[(size_t)k] -= alpha;
    double vv = 0;
    for (int i = k; i < rows; ++i) vv += v[(size_t)i] * v[(size_t)i];
    if (vv > 0) {
      for (int j = k; j < cols; ++j) {
        double s = 0;
        for (int i = k; i < rows; ++i) s += v[(size_t)i] * A[(size_t)i * cols + j];
        s = 2.0 * s / vv;
        for (int i = k; i < rows; ++i) A[(size_t)i * cols + j] -= s * v[(size_t)i];
      }
      double s = 0;
      for (int i = k; i < rows; ++i) s += v[(size_t)i] * b[(size_t)i];
      s = 2.0 * s / vv;
      for (int i = k; i < rows; ++i) b[(size_t)i] -= s * v[(size_t)i];
    }
  }
  x.assign((size_t)cols, 0.0);
  for (int i = cols - 1; i >= 0; --i) {
    double s = b[(size_t)i];
    for (int j = i + 1; j < cols; ++j) s -= A[(size_t)i * cols + j] * x[(size_t)j];
    const double d = A[(size_t)i * cols + i];
    if (d == 0.0) return false;
    x[(size_t)i] = s / d;
  }
  return true;
}

int refineAxisFromPoses(const std::vector<Matrix4d>& poses, double pivot[3], double axis[3]) {
  if (poses.empty()) return MVR_OK;   // reference: nothing registered, nothing to do (:420-421)
  const int m = (int)poses.size(), rows = 3 * m + 1;
  std::vector<double> A((size_t)rows * 3, 0.0), b((size_t)rows, 0.0), x;
  // axis: (R_i - I) n = 0 for every pose, regularised by n_x + n_y + n_z = 1 (:426-439).  The reference fills
  // A(i*3+j, k) = M(k, j) - delta_jk from the row-vector osg matrix M = R^T, i.e. exactly R(j, k) - delta_jk.
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < 3; ++j)
      for (int k = 0; k < 3; ++k) A[(size_t)(i * 3 + j) * 3 + k] = poses[(size_t)i].m[k * 4 + j] - ((j == k) ? 1.0 : 0.0);
  A[(size_t)(3 * m) * 3 + 0] = 1; A[(size_t)(3 * m) * 3 + 1] = 1; A[(size_t)(3 * m) * 3 + 2] = 1; b[(size_t)3 * m] = 1;
  if (!leastSquares(A, b, rows, 3, x)) return MVR_ERR_BAD_ARG;
  // the reference narrows to osg::Vec3 (float) before normalising (:437-438)
  float nf[3] = {(float)x[0], (float)x[1], (float)x[2]};
  const float len = std::sqrt(nf[0] * nf[0] + nf[1] * nf[1] + nf[2] * nf[2]);
  if (len > 0) for (int k = 0; k < 3; ++k) nf[k] /= len;
  for (int k = 0; k < 3; ++k) axis[k] = nf[k];
  // pivot: (R_i - I) c = -t_i, regularised by c_y = old pivot y (:442-450)
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < 3; ++j) b[(size_t)i * 3 + j] = -poses[(size_t)i].m[12 + j];
  A[(size_t)(3 * m) * 3 + 0] = 0; A[(size_t)(3 * m) * 3 + 1] = 1; A[(size_t)(3 * m) * 3 + 2] = 0; b[(size_t)3 * m] = (float)pivot[1];
  if (!leastSquares(A, b, rows, 3, x)) return MVR_ERR_BAD_ARG;
  for (int k = 0; k < 3; ++k) pivot[k] = (float)x[k];
  return MVR_OK;
}

}  // namespace mvr
