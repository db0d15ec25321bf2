// nn_search.cuh -- exact nearest neighbour of one query in a Morton-sorted uniform grid (device side).
//
// This is the search pcl::KdTreeFLANN::nearestKSearch(k = 1) performs for the reference inside
// icp.align / determineReciprocalCorrespondences / getFitnessScore (mvr/src/registrator.cpp:502, 569,
// 572, 649), restated for a grid:
//   1. scan the query's own cell, which for an aligned scan almost always holds the answer;
//   2. scan the (few) other cells that intersect the ball of radius sqrt(min(best, gate)) around the
//      query, pruning each by its lower-bound distance;
//   3. only when there is neither a candidate nor a gate, expand Chebyshev rings until the best
//      distance is provably smaller than anything outside the scanned cube.
//
// Exactness argument.  grid_t() rounds with |error| <= 2^-23 * |t| per point;
// `margin` (cells) absorbs that on both the query and the indexed point.  A point whose float d2 is
// <= lim has |p_a - q_a| <= sqrt(lim) * (1 + 2^-22) on every axis, hence its cell coordinate lies in
// [floor(t_a - rc), floor(t_a + rc)] with rc = sqrt(lim) * inv_cell * (1 + 1e-6) + margin: the box of
// step 2 cannot miss it.  Cell lower bounds shrink the geometric gap by `margin`, use cell_lo (a float
// strictly below the true edge) and a relative 1e-6, so lb is a true lower bound on the float d2 of
// every point filed in that cell; a cell is skipped only when lb > best (strict), so equal-distance
// ties with a lower index are never cut off.  Points (or queries) outside the grid are filed in the
// clamped boundary cell, so boundary cells are treated as unbounded outwards.
#pragma once
#include "common.cuh"

namespace mvr {

struct NnBest {
  float d2;
  int idx;   // original index of the best point, INT_MAX if none
  int pos;   // its position in the sorted array (-1 for a seed)
};

// Scan one cell's points; strict lexicographic improvement.
__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, uint32_t s, uint32_t e, float qx, float qy,
                                           float qz, NnBest& b) {
  for (uint32_t k = s; k < e; ++k) {
    float4 p = __ldg(pts + k);
    float d2 = d2_pinned(qx, qy, qz, p.x, p.y, p.z);
    int id = __float_as_int(p.w);
    if (lex_less(d2, id, b.d2, b.idx)) { b.d2 = d2; b.idx = id; b.pos = (int)k; }
  }
}

// Gap (cells, >= 0) between scaled coordinate t and the slab of cell c, boundary cells unbounded outwards.
__device__ __forceinline__ float cell_gap(float t, int c, int G) {
  float lo = (c > 0) ? (float)c - t : 0.0f;
  float hi = (c < G - 1) ? t - (float)(c + 1) : 0.0f;
  return fmaxf(fmaxf(lo, hi), 0.0f);
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

  // 1. own cell
  {
    const uint32_t m = morton3((uint32_t)cx, (uint32_t)cy, (uint32_t)cz);
    scan_range(ix.pts, __ldg(ix.start + m), __ldg(ix.start + m + 1), qx, qy, qz, b);
  }
  const float lim = fminf(b.d2, max_d2f);
  if (lim < MVR_INF) {
    // 2. the box of cells that the ball of radius sqrt(lim) can reach
    const float rc = sqrtf(lim) * g.inv_cell * 1.000001f + margin;
    const float Gm = (float)(G - 1);
    const int x0 = (int)fminf(fmaxf(floorf(tx - rc), 0.0f), Gm), x1 = (int)fminf(fmaxf(floorf(tx + rc), 0.0f), Gm);
    const int y0 = (int)fminf(fmaxf(floorf(ty - rc), 0.0f), Gm), y1 = (int)fminf(fmaxf(floorf(ty + rc), 0.0f), Gm);
    const int z0 = (int)fminf(fmaxf(floorf(tz - rc), 0.0f), Gm), z1 = (int)fminf(fmaxf(floorf(tz + rc), 0.0f), Gm);
    for (int z = z0; z <= z1; ++z) {
      const float ez = fmaxf(cell_gap(tz, z, G) - margin, 0.0f);
      const uint32_t mz = part1by2((uint32_t)z) << 2;
      for (int y = y0; y <= y1; ++y) {
        const float ey = fmaxf(cell_gap(ty, y, G) - margin, 0.0f);
        const float eyz = ez * ez + ey * ey;
        const uint32_t myz = mz | (part1by2((uint32_t)y) << 1);
        const bool own_row = (z == cz) && (y == cy);
        for (int x = x0; x <= x1; ++x) {
          if (own_row && x == cx) continue;
          const float ex = fmaxf(cell_gap(tx, x, G) - margin, 0.0f);
          const float lb = (eyz + ex * ex) * cell2;
          if (lb > b.d2 || lb > max_d2f) continue;
          const uint32_t m = myz | part1by2((uint32_t)x);
          const uint32_t s = __ldg(ix.start + m), e = __ldg(ix.start + m + 1);
          scan_range(ix.pts, s, e, qx, qy, qz, b);
        }
      }
    }
    return;
  }

  // 3. no candidate and no gate: ring expansion (ring 0 is done)
  for (int r = 1;; ++r) {
    const int x0 = max(cx - r, 0), x1 = min(cx + r, G - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, G - 1);
    const int z0 = max(cz - r, 0), z1 = min(cz + r, G - 1);
    for (int z = z0; z <= z1; ++z) {
      const float ez = fmaxf(cell_gap(tz, z, G) - margin, 0.0f);
      const bool zshell = (z - cz == r) || (cz - z == r);
      for (int y = y0; y <= y1; ++y) {
        const float ey = fmaxf(cell_gap(ty, y, G) - margin, 0.0f);
        const float eyz = ez * ez + ey * ey;
        const bool shell = zshell || (y - cy == r) || (cy - y == r);
        // on a shell face every x of the ring is new; otherwise only the two x end caps are
        const int xstep = shell ? 1 : 2 * r;
        for (int x = shell ? x0 : cx - r; x <= x1; x += xstep) {
          if (x < x0) continue;
          const float ex = fmaxf(cell_gap(tx, x, G) - margin, 0.0f);
          const float lb = (eyz + ex * ex) * cell2;
          if (lb > b.d2) continue;
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
    if (b.d2 < bu * bu * cell2) break;
  }
}

}  // namespace mvr
