// tile_search.cuh -- block-cooperative exact nearest neighbour for a block of SPATIALLY COHERENT queries: the
// candidates the whole block can need are staged ONCE in shared memory and bucketed along one axis; every thread
// then answers its own query from a short contiguous window of that staged list.
//
// This is the search pcl::KdTreeFLANN::nearestKSearch(k = 1) answers for the reference inside icp.align
// (mvr/src/registrator.cpp:569, 920, 1012, 1024), organised around what bounds the per-thread row walk
// of pair_search.cuh on B200 (round-1 ncu: 1 280 thread-instructions per query at 19 of 32 lanes, the L1 load pipe 72 %
// busy with 32 scattered 16-byte gathers per warp-level load):
//   * queries arrive sorted by the cells of a row-major grid (the per-align index of pair_index.cu), so the few hundred
//     queries of a block are a RIBBON of surface a few cells wide; the cells their balls can touch are a few dozen
//     short x runs of the candidate index.  The block fetches those runs with coalesced loads -- each candidate is
//     read from global memory once per block instead of once per query that looks at it;
//   * the staged candidates are counting-sorted in shared memory into 256 buckets along the axis the ribbon is
//     longest in.  A query then only needs the buckets its ball [q_a - rho, q_a + rho] overlaps: ONE contiguous
//     window of shared memory, scanned in a single loop -- no rows, no cell tables, no divergent walk, and roughly
//     (ribbon width x 2 rho) candidates instead of every point of every cell the ball's bounding box touches.
//
// Exactness.  A query is `near` when it holds a bound lim such that the answer is the lexicographic minimum
// (d2, index) over the candidates with pinned float d2 <= lim (a real candidate at distance lim exists, or nothing
// farther than lim is of interest).  Such a candidate k satisfies, as in pair_search.cuh,
//   (a) its BINNING cell lies in the box [floor(t - rc), floor(t + rc)] per axis around the query's scaled binning
//       coordinate t, rc = sqrt(lim) stretch inv_cell (1 + 1e-6) + margin: the query marks every row of that box
//       with its x range, the block stages the union of the marked ranges, so k is staged;
//   (b) |q_a - cur_k,a| <= sqrt(lim) (1 + 2^-22) on every axis of the CURRENT frame.  Buckets are
//       beta(v) = clamp(floor((v - a0) * inv_w), 0, 255), a monotone non-decreasing function of v evaluated with the
//       same pinned float operations for candidates and window ends; the window is [beta(q_a - rho'), beta(q_a + rho')]
//       with rho' = sqrt(lim) (1 + 1e-6) + 1e-6 |q_a| (the rounding of the subtraction), so beta(cur_k,a) is inside it.
// Every candidate of the window is compared: strictly smaller d2 wins; on EQUAL d2 (rare) the original indices of the
// two points are fetched and the lower one wins -- the lexicographic order of nn_search.cuh.  Candidates outside the
// ball are harmless.  A block whose ribbon does not fit (rows > TS_ROWS or candidates > TS_CAP) reports failure and its
// queries fall back to the per-thread search: the result never depends on which path ran.
#pragma once
#include "pair_search.cuh"

namespace mvr {

constexpr int TS_THREADS = FUSED_THREADS;   // 256
constexpr int TS_ROWS = 512;                // (y, z) rows of a tile: two per thread
constexpr int TS_NB = 256;                  // buckets along the sort axis: one per thread
#ifndef MVR_TS_CAP
#define MVR_TS_CAP 2048
#endif
constexpr int TS_CAP = MVR_TS_CAP;          // staged candidates
#ifndef MVR_TS_NEAR
#define MVR_TS_NEAR 2.1f
#endif
constexpr float TS_RC_NEAR = MVR_TS_NEAR;   // a query whose ball reaches farther than this many cells is not staged for

static_assert(TS_THREADS == 256, "tile search is laid out for 256-thread blocks");
static_assert(TS_CAP * sizeof(float4) >= PG_SEGS * PG_STRIDE * sizeof(uint2), "the fallback's segment lists live in the staging area");

struct TileSmem {
  float4 pts[TS_CAP];            // staged candidates in bucket order, .w = bits(sorted position in the candidate index);
                                 // the per-thread fallback's segment lists reuse the array
  uint32_t rowa[TS_ROWS + 2];    // per row: min x cell (atomicMin)  -> first sorted position of the row's run
  uint32_t rowb[TS_ROWS + 2];    // per row: max x cell + 1 (atomicMax) -> exclusive prefix of the run lengths
  uint32_t bucket[TS_NB + 1];    // counts -> exclusive prefix (bucket starts)
  uint32_t cursor[TS_NB];        // scatter cursors
  int red[12];                   // block reductions: y/z cell range, current-frame extent per axis
  uint32_t wsum[TS_THREADS / 32];
};

// One query of a tile, as its owner thread sees it.
struct TileQuery {
  bool near;              // participates (false: the slot is empty or the query's ball is too wide for a tile)
  float qx, qy, qz;       // the query in the current frame (distances are evaluated against it)
  float tx, ty, tz;       // its scaled coordinate in the binning frame of the candidate index
  float rc;               // radius in cells of the box that holds every candidate within lim (margins included, as in pg_search)
  float rho;              // sqrt(lim) (1 + 1e-6): ball radius in the current frame
  NnBest b;               // running best, pre-seeded like for pg_search (idx may be stale after the tile: pos and d2 are final)
  bool skipped;           // out: the tile left this query unanswered (its window is too long for one thread)
};

__device__ __forceinline__ int ts_f2o(float x) {   // order-preserving float -> signed int
  int o = __float_as_int(x);
  return o ^ ((o >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ts_o2f(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7fffffff)); }

// Exclusive scan of one value per thread over the block; *total gets the sum.  Two barriers.
__device__ __forceinline__ uint32_t ts_block_excl_scan(uint32_t v, uint32_t* wsum, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < TS_THREADS / 32; ++w) { const uint32_t c = wsum[w]; if (w < warp) base += c; tot += c; }
  *total = tot;
  __syncthreads();
  return base + incl - v;
}

__device__ __forceinline__ int ts_bucket(float v, float a0, float inv_w) {
  return (int)fminf(fmaxf(floorf(__fmul_rn(__fsub_rn(v, a0), inv_w)), 0.0f), (float)(TS_NB - 1));
}

// Block-collective.  nq queries, query k belongs to thread k % 256; load(k, q, stage) fills its TileQuery (called three
// times per query, stage 0 / 1 / 2 = extent, rows, window; reloading is cheap next to keeping several queries in
// registers across the barriers), store(k, q)
// receives the answered query (q.b.pos / q.b.d2 final) -- only for near queries and only when the tile was built.
//   g, start, pts : the candidate index (row-major grid over the BINNING frame, points sorted by binning cell holding
//                   their CURRENT coordinates, .w = original index)
// Returns true when the tile was built (every near query was stored), false when nothing was done.
template <class Load, class Store>
__device__ __forceinline__ bool tile_search(TileSmem& S, const PairGrid& g, const uint32_t* __restrict__ start, const float4* __restrict__ pts, int nq,
                                            Load load, Store store) {
  const int tid = threadIdx.x, lane = tid & 31;
  // ---- S0: reset
  if (tid < 12) S.red[tid] = (tid & 1) ? (int)0x80000000 : 0x7fffffff;   // even slots: minima, odd slots: maxima
  S.bucket[tid] = 0u;
  if (tid == 0) S.bucket[TS_NB] = 0u;
  __syncthreads();

  // ---- S1: the block's cell range in y and z (binning frame) and its extent per axis (current frame)
  const float fx = (float)(g.nx - 1), fy = (float)(g.ny - 1), fz = (float)(g.nz - 1);
  {
    int y0 = 0x7fffffff, y1 = (int)0x80000000, z0 = 0x7fffffff, z1 = (int)0x80000000;
    int e[6] = {0x7fffffff, (int)0x80000000, 0x7fffffff, (int)0x80000000, 0x7fffffff, (int)0x80000000};
    for (int k = tid; k < nq; k += TS_THREADS) {
      TileQuery q;
      load(k, q, 0);
      if (!q.near) continue;
      y0 = min(y0, (int)fminf(fmaxf(floorf(q.ty - q.rc), 0.0f), fy)); y1 = max(y1, (int)fminf(fmaxf(floorf(q.ty + q.rc), 0.0f), fy));
      z0 = min(z0, (int)fminf(fmaxf(floorf(q.tz - q.rc), 0.0f), fz)); z1 = max(z1, (int)fminf(fmaxf(floorf(q.tz + q.rc), 0.0f), fz));
      const float rx = q.rho + 1.0e-6f * fabsf(q.qx), ry = q.rho + 1.0e-6f * fabsf(q.qy), rz = q.rho + 1.0e-6f * fabsf(q.qz);
      e[0] = min(e[0], ts_f2o(__fsub_rn(q.qx, rx))); e[1] = max(e[1], ts_f2o(__fadd_rn(q.qx, rx)));
      e[2] = min(e[2], ts_f2o(__fsub_rn(q.qy, ry))); e[3] = max(e[3], ts_f2o(__fadd_rn(q.qy, ry)));
      e[4] = min(e[4], ts_f2o(__fsub_rn(q.qz, rz))); e[5] = max(e[5], ts_f2o(__fadd_rn(q.qz, rz)));
    }
    const int wy0 = __reduce_min_sync(0xffffffffu, y0), wy1 = __reduce_max_sync(0xffffffffu, y1);
    const int wz0 = __reduce_min_sync(0xffffffffu, z0), wz1 = __reduce_max_sync(0xffffffffu, z1);
    int we[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) we[k] = (k & 1) ? __reduce_max_sync(0xffffffffu, e[k]) : __reduce_min_sync(0xffffffffu, e[k]);
    if (lane == 0 && wy0 <= wy1) {
      atomicMin(&S.red[0], wy0); atomicMax(&S.red[1], wy1); atomicMin(&S.red[2], wz0); atomicMax(&S.red[3], wz1);
#pragma unroll
      for (int k = 0; k < 6; ++k) { if (k & 1) atomicMax(&S.red[4 + k], we[k]); else atomicMin(&S.red[4 + k], we[k]); }
    }
  }
  __syncthreads();
  const int Y0 = S.red[0], Y1 = S.red[1], Z0 = S.red[2], Z1 = S.red[3];
  if (Y0 > Y1) return true;                      // no near query at all
  const int NY = Y1 - Y0 + 1, NR = NY * (Z1 - Z0 + 1);
  if (NR > TS_ROWS) return false;
  // sort axis: the one the block's queries extend farthest along
  float a0, inv_w;
  int axis;
  {
    const float lo0 = ts_o2f(S.red[4]), hi0 = ts_o2f(S.red[5]), lo1 = ts_o2f(S.red[6]), hi1 = ts_o2f(S.red[7]), lo2 = ts_o2f(S.red[8]), hi2 = ts_o2f(S.red[9]);
    const float ex = hi0 - lo0, ey = hi1 - lo1, ez = hi2 - lo2;
    axis = (ex >= ey && ex >= ez) ? 0 : (ey >= ez ? 1 : 2);
    a0 = axis == 0 ? lo0 : (axis == 1 ? lo1 : lo2);
    const float ext = fmaxf(axis == 0 ? ex : (axis == 1 ? ey : ez), 1.0e-20f);
    inv_w = (float)TS_NB / ext;
  }

  // ---- S2: rows of the block's box, each with the union of the x ranges of the queries that touch it
  for (int r = tid; r < NR; r += TS_THREADS) { S.rowa[r] = 0x7fffffffu; S.rowb[r] = 0u; }
  __syncthreads();
  for (int k = tid; k < nq; k += TS_THREADS) {
    TileQuery q;
    load(k, q, 1);
    if (!q.near) continue;
    const uint32_t x0 = (uint32_t)fminf(fmaxf(floorf(q.tx - q.rc), 0.0f), fx), x1 = (uint32_t)fminf(fmaxf(floorf(q.tx + q.rc), 0.0f), fx);
    const int y0 = (int)fminf(fmaxf(floorf(q.ty - q.rc), 0.0f), fy), y1 = (int)fminf(fmaxf(floorf(q.ty + q.rc), 0.0f), fy);
    const int z0 = (int)fminf(fmaxf(floorf(q.tz - q.rc), 0.0f), fz), z1 = (int)fminf(fmaxf(floorf(q.tz + q.rc), 0.0f), fz);
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        const int r = (z - Z0) * NY + (y - Y0);
        atomicMin(&S.rowa[r], x0);
        atomicMax(&S.rowb[r], x1 + 1u);
      }
  }
  __syncthreads();

  // ---- S3: the runs (two rows per thread), their exclusive prefix = staging offsets
  uint32_t s_run[2] = {0u, 0u}, n_run[2] = {0u, 0u};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r = 2 * tid + h;
    if (r < NR) {
      const uint32_t xa = S.rowa[r], xb1 = S.rowb[r];
      if (xb1 > xa) {
        const int z = Z0 + r / NY, y = Y0 + r % NY;
        const uint32_t* row = start + ((size_t)z * g.ny + (size_t)y) * g.nx;
        s_run[h] = __ldg(row + xa);
        n_run[h] = __ldg(row + xb1) - s_run[h];
      }
    }
  }
  uint32_t T;
  const uint32_t ofs = ts_block_excl_scan(n_run[0] + n_run[1], S.wsum, &T);
  if (T > (uint32_t)TS_CAP) return false;
  if (2 * tid < NR) { S.rowa[2 * tid] = s_run[0]; S.rowb[2 * tid] = ofs; }
  if (2 * tid + 1 < NR) { S.rowa[2 * tid + 1] = s_run[1]; S.rowb[2 * tid + 1] = ofs + n_run[0]; }
  if (tid == 0) S.rowb[NR] = T;
  __syncthreads();

  if (T != 0u) {
    // ---- S4: count the staged candidates per bucket (groups of 8 lanes take one row at a time; a row holds ~10 points)
    const int grp = tid >> 3, l8 = tid & 7;
    for (int r = grp; r < NR; r += TS_THREADS / 8) {
      const uint32_t s = S.rowa[r], o = S.rowb[r], cnt = S.rowb[r + 1] - o;
      for (uint32_t k = l8; k < cnt; k += 8u) {
        const float4 p = __ldg(pts + s + k);
        const float pa = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
        atomicAdd(&S.bucket[ts_bucket(pa, a0, inv_w)], 1u);
      }
    }
    __syncthreads();
    // ---- S5: bucket starts
    {
      const uint32_t c = S.bucket[tid];
      uint32_t tot;
      const uint32_t ex = ts_block_excl_scan(c, S.wsum, &tot);
      S.bucket[tid] = ex;
      S.cursor[tid] = ex;
      if (tid == 0) S.bucket[TS_NB] = tot;
    }
    __syncthreads();
    // ---- S6: scatter into bucket order (the order inside a bucket follows the atomics; no result depends on it)
    for (int r = grp; r < NR; r += TS_THREADS / 8) {
      const uint32_t s = S.rowa[r], o = S.rowb[r], cnt = S.rowb[r + 1] - o;
      for (uint32_t k = l8; k < cnt; k += 8u) {
        float4 p = __ldg(pts + s + k);
        const float pa = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
        const uint32_t dst = atomicAdd(&S.cursor[ts_bucket(pa, a0, inv_w)], 1u);
        p.w = __uint_as_float(s + k);
        S.pts[dst] = p;
      }
    }
    __syncthreads();
  }

  // ---- S7: every near query scans the window of its ball
  for (int k = tid; k < nq; k += TS_THREADS) {
    TileQuery q;
    load(k, q, 2);
    if (!q.near) continue;
    q.skipped = false;
    if (T != 0u) {
      const float qa = axis == 0 ? q.qx : (axis == 1 ? q.qy : q.qz);
      const float ra = q.rho + 1.0e-6f * fabsf(qa);
      const int blo = ts_bucket(__fsub_rn(qa, ra), a0, inv_w), bhi = ts_bucket(__fadd_rn(qa, ra), a0, inv_w);
      const uint32_t k0 = S.bucket[blo], k1 = S.bucket[bhi + 1];
#ifdef MVR_TS_WIDE
      if (k1 - k0 > (uint32_t)MVR_TS_WIDE) { q.skipped = true; store(k, q); continue; }
#endif
      float bd = q.b.d2;
      int bpos = q.b.pos, bidx = q.b.idx;
      bool idx_known = true;   // bidx is the original index of the point at bpos (INT_MAX for "none yet")
      for (uint32_t kk = k0; kk < k1; kk += 4u) {
        float4 c[4];
        float d[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) c[u] = S.pts[min(kk + (uint32_t)u, k1 - 1u)];
#pragma unroll
        for (int u = 0; u < 4; ++u) d[u] = d2_pinned(q.qx, q.qy, q.qz, c[u].x, c[u].y, c[u].z);
        // most candidates are farther than the running best: one test settles four of them
        if (fminf(fminf(d[0], d[1]), fminf(d[2], d[3])) <= bd) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int cpos = __float_as_int(c[u].w);
            if (d[u] < bd) { bd = d[u]; bpos = cpos; idx_known = false; }
            else if (d[u] == bd && cpos != bpos) {   // equal distances: the lower original index wins (fetched only now)
              if (!idx_known) { bidx = __float_as_int(__ldg(&pts[bpos].w)); idx_known = true; }
              const int cidx = __float_as_int(__ldg(&pts[cpos].w));
              if (cidx < bidx) { bpos = cpos; bidx = cidx; }
            }
          }
        }
      }
      q.b.d2 = bd; q.b.pos = bpos; q.b.idx = bidx;
    }
    store(k, q);
  }
  return true;
}

}  // namespace mvr
