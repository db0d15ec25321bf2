// normals.cu -- kNN-PCA normal estimation (K10), a north-star extension with pcl::NormalEstimation
// semantics; the reference itself never estimates normals (they arrive inside its
// PointXYZRGBNormal .pcd files, mvr/include/types.h:14-18).
//
// One thread per indexed point: exact k nearest neighbours (self included, ties to the lowest index)
// by ring expansion over the uniform grid, two-pass covariance in double, Jacobi eigen-solve,
// normal = eigenvector of the smallest eigenvalue flipped towards the viewpoint,
// curvature = l0 / (l0 + l1 + l2).
#include "launch.h"
#include "small_solve.h"

namespace mvr {

constexpr int KNN_MAX = 32;

struct KnnList {
  float d[KNN_MAX];
  int id[KNN_MAX];
  int n, k;
  __device__ __forceinline__ float worst() const { return n < k ? MVR_INF : d[n - 1]; }
  __device__ __forceinline__ void push(float d2, int i) {
    if (n == k && !lex_less(d2, i, d[n - 1], id[n - 1])) return;
    int p = (n < k) ? n++ : k - 1;
    while (p > 0 && lex_less(d2, i, d[p - 1], id[p - 1])) { d[p] = d[p - 1]; id[p] = id[p - 1]; --p; }
    d[p] = d2; id[p] = i;
  }
};

__global__ void __launch_bounds__(128) k_normals(IndexDev ix, const float4* __restrict__ orig, int n, int k, float3 vp,
                                                 float4* __restrict__ out, int32_t* __restrict__ nbr) {
  int sp = blockIdx.x * blockDim.x + threadIdx.x;
  if (sp >= n) return;
  const float4 q = __ldg(ix.pts + sp);
  const int self = __float_as_int(q.w);
  if (sp >= ix.n_valid) {  // non-finite point
    out[self] = make_float4(nanf(""), nanf(""), nanf(""), nanf(""));
    if (nbr) for (int a = 0; a < k; ++a) nbr[(size_t)self * k + a] = -1;
    return;
  }
  const GridDev& g = ix.g;
  const int G = g.G;
  const float tx = grid_t(q.x, g.ox, g.inv_cell), ty = grid_t(q.y, g.oy, g.inv_cell), tz = grid_t(q.z, g.oz, g.inv_cell);
  const int cx = grid_cell(tx, G), cy = grid_cell(ty, G), cz = grid_cell(tz, G);
  const float margin = MVR_CELL_MARGIN + 1.0e-6f * fmaxf(fabsf(tx), fmaxf(fabsf(ty), fabsf(tz)));
  const float cell2 = g.cell_lo * g.cell_lo * MVR_REL_SHRINK;
  KnnList L;
  L.n = 0; L.k = k;
  for (int r = 0;; ++r) {
    const int x0 = max(cx - r, 0), x1 = min(cx + r, G - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, G - 1);
    const int z0 = max(cz - r, 0), z1 = min(cz + r, G - 1);
    for (int z = z0; z <= z1; ++z) {
      const float ez = fmaxf(fmaxf(fmaxf((float)z - tz, tz - (float)(z + 1)), 0.0f) - margin, 0.0f);
      const bool zshell = (z - cz == r) || (cz - z == r);
      for (int y = y0; y <= y1; ++y) {
        const float ey = fmaxf(fmaxf(fmaxf((float)y - ty, ty - (float)(y + 1)), 0.0f) - margin, 0.0f);
        const bool shell = zshell || (y - cy == r) || (cy - y == r);
        const int xstep = (shell || r == 0) ? 1 : 2 * r;
        for (int x = shell ? x0 : cx - r; x <= x1; x += xstep) {
          if (x < x0) continue;
          const float ex = fmaxf(fmaxf(fmaxf((float)x - tx, tx - (float)(x + 1)), 0.0f) - margin, 0.0f);
          if ((ez * ez + ey * ey + ex * ex) * cell2 > L.worst()) continue;
          const uint32_t m = morton3((uint32_t)x, (uint32_t)y, (uint32_t)z);
          const uint32_t s = __ldg(ix.start + m), e = __ldg(ix.start + m + 1);
          for (uint32_t a = s; a < e; ++a) {
            float4 p = __ldg(ix.pts + a);
            L.push(d2_pinned(q.x, q.y, q.z, p.x, p.y, p.z), __float_as_int(p.w));
          }
        }
      }
    }
    float u = MVR_INF;
    if (cx - r > 0) u = fminf(u, tx - (float)(cx - r));
    if (cx + r < G - 1) u = fminf(u, (float)(cx + r + 1) - tx);
    if (cy - r > 0) u = fminf(u, ty - (float)(cy - r));
    if (cy + r < G - 1) u = fminf(u, (float)(cy + r + 1) - ty);
    if (cz - r > 0) u = fminf(u, tz - (float)(cz - r));
    if (cz + r < G - 1) u = fminf(u, (float)(cz + r + 1) - tz);
    if (u == MVR_INF) break;
    const float bu = fmaxf(u - margin, 0.0f);
    if (L.worst() < bu * bu * cell2) break;
  }
  if (nbr) for (int a = 0; a < k; ++a) nbr[(size_t)self * k + a] = a < L.n ? L.id[a] : -1;
  if (L.n < 3) { out[self] = make_float4(nanf(""), nanf(""), nanf(""), nanf("")); return; }
  // two-pass covariance in double over the neighbours (coordinates re-read by original index)
  double c[3] = {0, 0, 0};
  for (int a = 0; a < L.n; ++a) { float4 p = __ldg(orig + L.id[a]); c[0] += p.x; c[1] += p.y; c[2] += p.z; }
  const double inv = 1.0 / (double)L.n;
  c[0] *= inv; c[1] *= inv; c[2] *= inv;
  double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int a = 0; a < L.n; ++a) {
    float4 p = __ldg(orig + L.id[a]);
    double d[3] = {(double)p.x - c[0], (double)p.y - c[1], (double)p.z - c[2]};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) C[r * 3 + s] += d[r] * d[s];
  }
#pragma unroll
  for (int a = 0; a < 9; ++a) C[a] *= inv;
  double w[3], V[9];
  eig_sym3(C, w, V);
  double nx = V[0], ny = V[3], nz = V[6];
  double vx = (double)vp.x - q.x, vy = (double)vp.y - q.y, vz = (double)vp.z - q.z;
  if (nx * vx + ny * vy + nz * vz < 0) { nx = -nx; ny = -ny; nz = -nz; }
  double tr = w[0] + w[1] + w[2];
  out[self] = make_float4((float)nx, (float)ny, (float)nz, (float)(tr > 0 ? fabs(w[0] / tr) : 0.0));
}

cudaError_t launch_normals(IndexDev ix, const float4* pts_orig, int n, int k, float3 viewpoint, float4* out_nxyzc,
                           int32_t* out_nbr, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  if (k < 1 || k > KNN_MAX) return cudaErrorInvalidValue;
  k_normals<<<(n + 127) / 128, 128, 0, s>>>(ix, pts_orig, n, k, viewpoint, out_nxyzc, out_nbr); count_launch();
  return cudaGetLastError();
}

}  // namespace mvr
