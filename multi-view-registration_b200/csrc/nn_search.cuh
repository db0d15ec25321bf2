// nn_search.cuh -- exact nearest neighbour of one query in a Morton-sorted uniform grid (device side).
//
// This is the search pcl::KdTreeFLANN::nearestKSearch(k = 1) performs for the reference inside
// icp.align / determineReciprocalCorrespondences / getFitnessScore (mvr/src/registrator.cpp:502, 569,
// 572, 649), restated for a grid: Chebyshev rings of cells around the query's cell are scanned until
// the best squared distance is provably smaller than anything outside the scanned cube.
//
// Exactness argument: a point filed outside the cube of radius r has scaled
// coordinate difference >= u on some axis, where u is the distance in cells from the query to the
// nearest cube face that still has grid behind it.  grid_t() rounds with |error| <= 2^-23 * |t| per
// point, so after shrinking u by `margin` cells and the edge to cell_lo, b = (u - margin) * cell_lo
// is a true lower bound on |p_a - q_a|, and b*b*(1 - 1e-6) a lower bound on the float d2 the pinned
// formula yields.  We stop only when best < that bound (strict), so ties are never cut off.
#pragma once
#include "common.cuh"

namespace mvr {

struct NnBest {
  float d2;
  int idx;   // original index of the best point, INT_MAX if none
};

// Scan one cell's points; strict lexicographic improvement.
__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, uint32_t s, uint32_t e, float qx, float qy,
                                           float qz, NnBest& b) {
  for (uint32_t k = s; k < e; ++k) {
    float4 p = __ldg(pts + k);
    float d2 = d2_pinned(qx, qy, qz, p.x, p.y, p.z);
    int id = __float_as_int(p.w);
    if (lex_less(d2, id, b.d2, b.idx)) { b.d2 = d2; b.idx = id; }
  }
}

// Search the index for the nearest point to (qx,qy,qz).  `b` may be pre-seeded with a known
// candidate (reciprocal test).  max_d2f: nothing farther than this (float, rounded up) is of
// interest; +inf for an un-gated search.  The query must be finite.
__device__ __forceinline__ void nn_search(const IndexDev& ix, float qx, float qy, float qz, float max_d2f, NnBest& b) {
  const GridDev& g = ix.g;
  if (ix.n_valid <= 0) return;
  const int G = g.G;
  const float tx = grid_t(qx, g.ox, g.inv_cell), ty = grid_t(qy, g.oy, g.inv_cell), tz = grid_t(qz, g.oz, g.inv_cell);
  const int cx = grid_cell(tx, G), cy = grid_cell(ty, G), cz = grid_cell(tz, G);
  const float margin = MVR_CELL_MARGIN + 1.0e-6f * fmaxf(fabsf(tx), fmaxf(fabsf(ty), fabsf(tz)));
  const float cell2 = g.cell_lo * g.cell_lo * MVR_REL_SHRINK;

  for (int r = 0;; ++r) {
    const int x0 = max(cx - r, 0), x1 = min(cx + r, G - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, G - 1);
    const int z0 = max(cz - r, 0), z1 = min(cz + r, G - 1);
    for (int z = z0; z <= z1; ++z) {
      const float gz = fmaxf(fmaxf((float)z - tz, tz - (float)(z + 1)), 0.0f);
      const float ez = fmaxf(gz - margin, 0.0f);
      const bool zshell = (z - cz == r) || (cz - z == r);
      for (int y = y0; y <= y1; ++y) {
        const float gy = fmaxf(fmaxf((float)y - ty, ty - (float)(y + 1)), 0.0f);
        const float ey = fmaxf(gy - margin, 0.0f);
        const float eyz = ez * ez + ey * ey;
        const bool shell = zshell || (y - cy == r) || (cy - y == r);
        // on a shell face every x of the ring is new; otherwise only the two x end caps are
        const int xstep = (shell || r == 0) ? 1 : 2 * r;
        for (int x = shell ? x0 : cx - r; x <= x1; x += xstep) {
          if (x < x0) continue;
          const float gx = fmaxf(fmaxf((float)x - tx, tx - (float)(x + 1)), 0.0f);
          const float ex = fmaxf(gx - margin, 0.0f);
          const float lb = (eyz + ex * ex) * cell2;
          if (lb > b.d2 || lb > max_d2f) continue;
          const uint32_t m = morton3((uint32_t)x, (uint32_t)y, (uint32_t)z);
          const uint32_t s = __ldg(ix.start + m), e = __ldg(ix.start + m + 1);
          scan_range(ix.pts, s, e, qx, qy, qz, b);
        }
      }
    }
    // distance (cells) to the nearest face of the scanned cube that still has grid behind it
    float u = MVR_INF;
    if (cx - r > 0) u = fminf(u, tx - (float)(cx - r));
    if (cx + r < G - 1) u = fminf(u, (float)(cx + r + 1) - tx);
    if (cy - r > 0) u = fminf(u, ty - (float)(cy - r));
    if (cy + r < G - 1) u = fminf(u, (float)(cy + r + 1) - ty);
    if (cz - r > 0) u = fminf(u, tz - (float)(cz - r));
    if (cz + r < G - 1) u = fminf(u, (float)(cz + r + 1) - tz);
    if (u == MVR_INF) break;  // whole grid scanned
    const float bu = fmaxf(u - margin, 0.0f);
    const float B = bu * bu * cell2;
    if (b.d2 < B || B > max_d2f) break;
  }
}

}  // namespace mvr
