// search.cu -- brick-tiled exact nearest-neighbour search with TMA-staged candidates (K4/K5).
//
// Replaces the per-point pcl::KdTreeFLANN::nearestKSearch(k = 1) calls that icp.align(),
// determineReciprocalCorrespondences() and getFitnessScore() make for every source point
// (reference call sites mvr/src/registrator.cpp:502, 569, 572, 649, 920, 1012).
//
// Both clouds of a search are binned in ONE grid and sorted by Morton cell key, so a level-2 Morton
// cell (a "brick" of 4 x 4 x 4 grid cells) is one contiguous run of the sorted points and of the
// cell-start table.  One CTA takes one occupied brick of QUERY points and stages, with one 1-D TMA
// bulk copy (cp.async.bulk -> mbarrier) per 2 x 2 x 2 sub-brick, every CANDIDATE point of the
// surrounding 8 x 8 x 8 cells (the brick and a two-cell halo) in shared memory together with a local
// cell table.  Each thread then resolves one query against shared memory only: its own cell, ring 1
// if that was empty, then the cells that the ball of radius sqrt(min(best, gate)) reaches, each
// pruned by its lower-bound distance.  A query whose ball leaves the staged region (no gate, a gate
// wider than the halo, or a brick whose candidates do not fit the staging buffer) falls through to
// the global-memory search of nn_search.cuh, seeded with what it already found, so every answer is
// exact.  With the pair grid's cell edge >= gate / 2 (api.cu) the halo covers the gate and a gated
// ICP iteration never leaves shared memory.
//
// The arithmetic per candidate is the pinned float32 expression of common.cuh, ties resolve to the
// lowest original index, so results are bit-identical to the CPU oracle whatever the grid.
//
// Algorithmic bytes: 16 B per query + 8 B per result + 16 B per candidate
// point once per launch.  Staging re-reads a candidate once per neighbouring brick (from L2).
#include <algorithm>

#include "launch.h"
#include "nn_search.cuh"

namespace mvr {

#ifndef MVR_BS_THREADS
#define MVR_BS_THREADS 128
#endif
#ifndef MVR_BS_CAP
#define MVR_BS_CAP 1536
#endif
#ifndef MVR_BS_CTAS_PER_SM
#define MVR_BS_CTAS_PER_SM 7
#endif
constexpr int BS_THREADS = MVR_BS_THREADS;   // a brick holds ~100 queries: 4 warps keep the lanes busy, more CTAs per SM hide its barriers
constexpr int BS_CAP = MVR_BS_CAP;           // staged candidate points per brick (16 B each); bigger regions fall back to global memory
constexpr int BS_RUNS = 64;    // 4 x 4 x 4 sub-bricks of 2 x 2 x 2 cells
constexpr int BS_REGION = 8;   // staged region edge, cells
constexpr int BS_HALO = 2;     // cells around the brick

enum { BS_MODE_NN = 0, BS_MODE_FWD = 1, BS_MODE_REV = 2 };

__device__ __forceinline__ uint32_t compact1by2(uint32_t v) {
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030c30c3u;
  v = (v | (v >> 4)) & 0x0300f00fu;
  v = (v | (v >> 8)) & 0x030000ffu;
  v = (v | (v >> 16)) & 0x3ffu;
  return v;
}

// ---- mbarrier / TMA bulk copy (PTX) ----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completes on `bar`.
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- search over the staged region -----------------------------------------------------------
// Threads of a warp hold different queries, so "walk my cells, scan the points of each" would run the
// point loop with a handful of lanes at a time.  Instead every phase first COLLECTS the candidate
// cell ranges of a fixed set of cells (same trip count on every lane: own cell; the 3 slabs of ring 1,
// fully unrolled; rows of a larger box) into a small per-thread list, then scans the list with ONE
// flat loop whose body is one point, so lanes diverge only in how many points they have left.
constexpr int BS_LIST = 9;   // list entries per thread: 9 cells of a ring-1 slab / 8 cells of a box row

struct LocalIx {
  const float4* pts;      // shared: staged candidates, sub-brick after sub-brick, cell order inside
  const uint32_t* cell;   // shared: [run * 8 + inner] = first slot | end slot << 12 | run << 24
  const int* delta;       // shared: global sorted position = staged slot + delta[run]
  uint32_t* list;         // shared: this thread's range list, stride BS_THREADS
  int rx, ry, rz;         // grid cell of the region's corner (may be negative: cells outside the grid are empty)
  bool empty;             // nothing staged
};

__device__ __forceinline__ bool range_nonempty(uint32_t w) { return ((w ^ (w >> 12)) & 0xfffu) != 0u; }

// Per-axis parts of a cell's index into LocalIx::cell (they add up to run * 8 + inner).
__device__ __forceinline__ int cell_part(int l, int axis) { return ((l >> 1) << (3 + 2 * axis)) + ((l & 1) << axis); }

__device__ __forceinline__ void fold(const float4 p, int pos, float qx, float qy, float qz, NnBest& b) {
  const float d2 = d2_pinned(qx, qy, qz, p.x, p.y, p.z);
  const int id = __float_as_int(p.w);
  if (lex_less(d2, id, b.d2, b.idx)) { b.d2 = d2; b.idx = id; b.pos = pos; }
}

// One loop over all points of the listed ranges, four points per trip: the four shared-memory loads and
// distance evaluations are independent, which is what hides their latency (a warp has few peers to
// overlap with).  Slots past the end of a range are clamped to its last point -- folding a point twice
// changes nothing, the comparison is strict.
__device__ __forceinline__ void flat_scan(const LocalIx& L, int cnt, float qx, float qy, float qz, NnBest& b) {
  int li = 0, dk = 0;
  uint32_t k = 0, e = 0;
  for (;;) {
    if (k >= e) {
      if (li >= cnt) break;
      const uint32_t w = L.list[li * BS_THREADS];
      ++li;
      k = w & 0xfffu; e = (w >> 12) & 0xfffu; dk = L.delta[w >> 24];   // BS_CAP <= 2048 keeps slots in 12 bits
    }
    const uint32_t last = e - 1u;
    const uint32_t k1 = min(k + 1u, last), k2 = min(k + 2u, last), k3 = min(k + 3u, last);
    const float4 p0 = L.pts[k], p1 = L.pts[k1], p2 = L.pts[k2], p3 = L.pts[k3];
    fold(p0, (int)k + dk, qx, qy, qz, b);
    fold(p1, (int)k1 + dk, qx, qy, qz, b);
    fold(p2, (int)k2 + dk, qx, qy, qz, b);
    fold(p3, (int)k3 + dk, qx, qy, qz, b);
    k += 4u;
  }
}

// Exact NN of a query of the brick among the staged candidates.  Returns false when the staged
// region cannot prove the answer (b then holds the best found so far, a valid seed for nn_search).
// Exactness: as in nn_search.cuh -- a cell is left out only when its lower bound exceeds min(best, gate),
// and the box of cells finally required is either covered by what was scanned or the search gives up.
__device__ __forceinline__ bool local_search(const LocalIx& L, const GridDev& g, float qx, float qy, float qz, float max_d2f, NnBest& b) {
  const int G = g.G;
  const float tx = grid_t(qx, g.ox, g.inv_cell), ty = grid_t(qy, g.oy, g.inv_cell), tz = grid_t(qz, g.oz, g.inv_cell);
  const int cx = grid_cell(tx, G), cy = grid_cell(ty, G), cz = grid_cell(tz, G);
  const float margin = MVR_CELL_MARGIN + 1.0e-6f * fmaxf(fabsf(tx), fmaxf(fabsf(ty), fabsf(tz)));
  const float cell2 = g.cell_lo * g.cell_lo * MVR_REL_SHRINK;
  const float Gm = (float)(G - 1);
  const int lx = cx - L.rx, ly = cy - L.ry, lz = cz - L.rz;
  int done = 0;   // cells [c - done, c + done]^3 are dealt with
  if (!L.empty) {
    // ---- own cell -------------------------------------------------------------------------------
    {
      const uint32_t w = L.cell[cell_part(lx, 0) + cell_part(ly, 1) + cell_part(lz, 2)];
      L.list[0] = w;
      flat_scan(L, range_nonempty(w) ? 1 : 0, qx, qy, qz, b);
    }
    // ---- ring 1, unless the ball already fits the own cell ---------------------------------------
    float gx[3], gy[3], gz[3];   // squared gaps (cells^2) to the slabs c-1, c, c+1; +inf outside the grid
    int fx[3], fy[3], fz[3];
#pragma unroll
    for (int o = 0; o < 3; ++o) {
      const int X = cx + o - 1, Y = cy + o - 1, Z = cz + o - 1;
      float a = fmaxf(cell_gap(tx, X, G) - margin, 0.0f); gx[o] = (X >= 0 && X < G) ? a * a : MVR_INF;
      a = fmaxf(cell_gap(ty, Y, G) - margin, 0.0f); gy[o] = (Y >= 0 && Y < G) ? a * a : MVR_INF;
      a = fmaxf(cell_gap(tz, Z, G) - margin, 0.0f); gz[o] = (Z >= 0 && Z < G) ? a * a : MVR_INF;
      fx[o] = cell_part(lx + o - 1, 0); fy[o] = cell_part(ly + o - 1, 1); fz[o] = cell_part(lz + o - 1, 2);
    }
    const float near2 = fminf(fminf(fminf(gx[0], gx[2]), fminf(gy[0], gy[2])), fminf(gz[0], gz[2])) * cell2;
    if (near2 <= fminf(b.d2, max_d2f)) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int oz = (s == 0) ? 1 : (s == 1 ? 0 : 2);   // own slab first: it tightens the bound for the other two
        int cnt = 0;
        if (gz[oz] * cell2 <= fminf(b.d2, max_d2f)) {
          const float lim = fminf(b.d2, max_d2f);
#pragma unroll
          for (int oy = 0; oy < 3; ++oy) {
#pragma unroll
            for (int ox = 0; ox < 3; ++ox) {
              if (oz == 1 && oy == 1 && ox == 1) continue;
              const float lb = (gz[oz] + gy[oy] + gx[ox]) * cell2;
              if (lb <= lim) {
                const uint32_t w = L.cell[fz[oz] + fy[oy] + fx[ox]];
                if (range_nonempty(w)) { L.list[cnt * BS_THREADS] = w; ++cnt; }
              }
            }
          }
        }
        flat_scan(L, cnt, qx, qy, qz, b);
      }
    }
    done = 1;
  }
  const float lim = fminf(b.d2, max_d2f);
  if (!(lim < MVR_INF)) return false;
  const float rc = sqrtf(lim) * g.inv_cell * 1.000001f + margin;
  const int x0 = (int)fminf(fmaxf(floorf(tx - rc), 0.0f), Gm), x1 = (int)fminf(fmaxf(floorf(tx + rc), 0.0f), Gm);
  const int y0 = (int)fminf(fmaxf(floorf(ty - rc), 0.0f), Gm), y1 = (int)fminf(fmaxf(floorf(ty + rc), 0.0f), Gm);
  const int z0 = (int)fminf(fmaxf(floorf(tz - rc), 0.0f), Gm), z1 = (int)fminf(fmaxf(floorf(tz + rc), 0.0f), Gm);
  if (done && x0 >= cx - 1 && x1 <= cx + 1 && y0 >= cy - 1 && y1 <= cy + 1 && z0 >= cz - 1 && z1 <= cz + 1) return true;
  if (x0 < L.rx || y0 < L.ry || z0 < L.rz || x1 >= L.rx + BS_REGION || y1 >= L.ry + BS_REGION || z1 >= L.rz + BS_REGION) return false;
  if (L.empty) return true;
  // ---- the rest of the box, row by row (early iterations, wide gates) -----------------------------
  for (int z = z0; z <= z1; ++z) {
    const float ez = fmaxf(cell_gap(tz, z, G) - margin, 0.0f);
    const bool zin = (z >= cz - 1) && (z <= cz + 1);
    const int pz = cell_part(z - L.rz, 2);
    for (int y = y0; y <= y1; ++y) {
      const float ey = fmaxf(cell_gap(ty, y, G) - margin, 0.0f);
      const float eyz = ez * ez + ey * ey;
      if (eyz * cell2 > fminf(b.d2, max_d2f)) continue;
      const bool yzin = zin && (y >= cy - 1) && (y <= cy + 1);
      const int pyz = pz + cell_part(y - L.ry, 1);
      const float lim2 = fminf(b.d2, max_d2f);
      int cnt = 0;
      for (int x = x0; x <= x1; ++x) {   // at most BS_REGION = 8 <= BS_LIST cells
        if (yzin && x >= cx - 1 && x <= cx + 1) continue;   // ring 1: scanned above
        const float ex = fmaxf(cell_gap(tx, x, G) - margin, 0.0f);
        if ((eyz + ex * ex) * cell2 > lim2) continue;
        const uint32_t w = L.cell[pyz + cell_part(x - L.rx, 0)];
        if (range_nonempty(w)) { L.list[cnt * BS_THREADS] = w; ++cnt; }
      }
      flat_scan(L, cnt, qx, qy, qz, b);
    }
  }
  return true;
}

struct BrickArgs {
  QueryDev q;              // queries, sorted by (at least) brick
  IndexDev c;              // candidates
  const uint32_t* bricks;  // occupied query bricks (level-2 Morton codes)
  const uint32_t* n_bricks;
  const int* done;         // nullable: non-zero turns the launch into a no-op
  // NN
  int32_t* out_idx;
  float* out_d2;
  // FWD
  double max2;             // gate, exact (PCL compares in double)
  float max_d2f;           // gate rounded up to float: bounds the search
  int32_t* corr_p;         // [source original index] sorted position of the matched target point, -1 none
  float* corr_d2;
  uint32_t* rmin;          // [target sorted position] min over choosers of d2 bits (nullable: not reciprocal)
  // REV
  int32_t* rnn;            // [target sorted position] original index of the nearest source point
};

template <int MODE>
__global__ void __launch_bounds__(BS_THREADS) k_brick_search(BrickArgs a) {
  __shared__ __align__(16) float4 s_pts[BS_CAP];
  __shared__ uint32_t s_cell[BS_RUNS * 8];
  __shared__ uint32_t s_list[BS_LIST * BS_THREADS];
  __shared__ int s_delta[BS_RUNS];
  __shared__ uint32_t s_warp[2];
  __shared__ __align__(8) uint64_t s_bar;

  if (a.done && *a.done) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  uint32_t phase = 0;
  const GridDev g = a.c.g;
  const int G1 = g.G >> 1;   // level-1 cells per axis
  const uint32_t n_work = __ldg(a.n_bricks);

  for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
    const uint32_t brick = __ldg(a.bricks + w);
    const int bx = (int)compact1by2(brick), by = (int)compact1by2(brick >> 1), bz = (int)compact1by2(brick >> 2);
    const uint32_t q_s = __ldg(a.q.start + ((size_t)brick << a.q.shift)), q_e = __ldg(a.q.start + ((size_t)(brick + 1) << a.q.shift));

    // ---- stage the candidates of the 4 x 4 x 4 sub-bricks around the brick -------------------
    uint32_t gs[9];
    uint32_t len = 0;
    if (tid < BS_RUNS) {
      const int l1x = 2 * bx - 1 + (tid & 3), l1y = 2 * by - 1 + ((tid >> 2) & 3), l1z = 2 * bz - 1 + (tid >> 4);
      if (l1x >= 0 && l1y >= 0 && l1z >= 0 && l1x < G1 && l1y < G1 && l1z < G1) {
        const uint32_t* t = a.c.start + ((size_t)morton3((uint32_t)l1x, (uint32_t)l1y, (uint32_t)l1z) << 3);
#pragma unroll
        for (int k = 0; k < 9; ++k) gs[k] = __ldg(t + k);
        len = gs[8] - gs[0];
      } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) gs[k] = 0;
      }
    }
    uint32_t incl = len;
    if (warp < 2) {
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) s_warp[warp] = incl;
    }
    __syncthreads();   // also: every thread is past the previous brick's reads of s_pts / s_cell
    const uint32_t total = s_warp[0] + s_warp[1];
    const bool staged = total <= (uint32_t)BS_CAP;
    if (staged && tid < BS_RUNS) {
      const uint32_t base = incl - len + (warp == 1 ? s_warp[0] : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        s_cell[tid * 8 + k] = (base + (gs[k] - gs[0])) | ((base + (gs[k + 1] - gs[0])) << 12) | ((uint32_t)tid << 24);
      s_delta[tid] = (int)gs[0] - (int)base;
      if (tid == 0 && total > 0) mbar_arrive_expect_tx(&s_bar, total * (uint32_t)sizeof(float4));
      if (len > 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of s_pts vs the async write
      if (len > 0) tma_load_1d(s_pts + base, a.c.pts + gs[0], len * (uint32_t)sizeof(float4), &s_bar);
    }
    __syncthreads();   // s_cell / s_delta visible
    if (staged && total > 0) {
      uint32_t spins = 0;
      while (!mbar_try_wait(&s_bar, phase)) {
        if (++spins > (1u << 24)) __trap();   // a lost bulk copy must fail loudly, not hang the device
      }
      phase ^= 1u;
    }
    LocalIx L;
    L.pts = s_pts; L.cell = s_cell; L.delta = s_delta; L.list = s_list + tid;
    L.rx = 4 * bx - BS_HALO; L.ry = 4 * by - BS_HALO; L.rz = 4 * bz - BS_HALO;
    L.empty = (total == 0);

    // ---- one query per thread -----------------------------------------------------------------
    for (uint32_t k = q_s + tid; k < q_e; k += BS_THREADS) {
      const float4 p = __ldg(a.q.pts + k);
      float gate = (MODE == BS_MODE_NN) ? MVR_INF : a.max_d2f;
      NnBest b{MVR_INF, 0x7fffffff, -1};
      if (MODE == BS_MODE_REV) {
        const uint32_t r = a.rmin[k];
        if (r == 0x7f800000u) continue;   // no source point chose this target point
        a.rmin[k] = 0x7f800000u;          // re-armed for the next iteration
        gate = __uint_as_float(r);
        b.d2 = gate;                      // a chooser sits at exactly this distance: (gate, lowest chooser) is found
      }
      bool ok = staged && local_search(L, g, p.x, p.y, p.z, gate, b);
      if (!ok) nn_search(a.c, p.x, p.y, p.z, gate, b);
      const bool found = b.idx != 0x7fffffff;
      if (MODE == BS_MODE_NN) {
        const int i = __float_as_int(p.w);
        a.out_idx[i] = found ? b.idx : -1;
        a.out_d2[i] = found ? b.d2 : MVR_INF;
      } else if (MODE == BS_MODE_FWD) {
        const int i = __float_as_int(p.w);
        const bool keep = found && !((double)b.d2 > a.max2);   // PCL: if (distance > max_dist_sqr) continue;
        a.corr_p[i] = keep ? b.pos : -1;
        a.corr_d2[i] = keep ? b.d2 : MVR_INF;
        if (keep && a.rmin) atomicMin(a.rmin + b.pos, __float_as_uint(b.d2));
      } else {
        a.rnn[k] = found ? b.idx : -1;
      }
    }
  }

  // non-finite query points are filed after the last cell: they have no neighbour
  if (MODE != BS_MODE_REV && blockIdx.x == 0) {
    const size_t cells = (size_t)1 << (3 * g.bits);
    const uint32_t q_s = __ldg(a.q.start + ((cells >> 6) << a.q.shift)), q_e = __ldg(a.q.start + ((cells >> 6) << a.q.shift) + 1);
    for (uint32_t k = q_s + tid; k < q_e; k += BS_THREADS) {
      const int i = __float_as_int(__ldg(a.q.pts + k).w);
      if (MODE == BS_MODE_NN) { a.out_idx[i] = -1; a.out_d2[i] = MVR_INF; }
      else { a.corr_p[i] = -1; a.corr_d2[i] = MVR_INF; }
    }
  }
}

// ---- occupied-brick list ---------------------------------------------------------------------
// list[0 .. *count) = level-2 Morton codes of the bricks that hold points (any order).  `next` is the
// counter of the following launch, zeroed here so that no memset is needed between iterations.
__global__ void __launch_bounds__(256) k_list_bricks(const uint32_t* __restrict__ start, uint32_t n_bricks, int shift,
                                                     uint32_t* __restrict__ list, uint32_t* __restrict__ count, uint32_t* __restrict__ next,
                                                     const int* __restrict__ done) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0 && next) *next = 0;
  if (done && *done) return;
  if (b >= n_bricks) return;
  const bool occ = __ldg(start + ((size_t)b << shift)) != __ldg(start + ((size_t)(b + 1) << shift));
  const uint32_t m = __ballot_sync(__activemask(), occ);
  if (!occ) return;
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(count, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  list[base + __popc(m & ((1u << lane) - 1u))] = b;
}

__global__ void k_fill_u32(uint32_t* __restrict__ p, size_t n, uint32_t v) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

cudaError_t launch_fill_u32(uint32_t* p, size_t n, uint32_t v, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  k_fill_u32<<<blocks, 256, 0, s>>>(p, n, v); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_list_bricks(const uint32_t* start, int bits, int shift, uint32_t* list, uint32_t* count, uint32_t* next,
                               const int* d_done, cudaStream_t s) {
  const uint32_t nb = 1u << (3 * (bits - 2));
  k_list_bricks<<<(nb + 255) / 256, 256, 0, s>>>(start, nb, shift, list, count, next, d_done); count_launch();
  return cudaGetLastError();
}

static int brick_grid(int n_queries) {
  // persistent CTAs: enough to fill the machine, no more than there can be bricks with work
  const int want = 148 * MVR_BS_CTAS_PER_SM;
  return std::max(1, std::min(want, (n_queries + 15) / 16));
}

cudaError_t launch_brick_nn(QueryDev q, int nq, IndexDev c, const uint32_t* bricks, const uint32_t* n_bricks, int32_t* out_idx,
                            float* out_d2, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  BrickArgs a{};
  a.q = q; a.c = c; a.bricks = bricks; a.n_bricks = n_bricks; a.out_idx = out_idx; a.out_d2 = out_d2;
  a.max_d2f = INFINITY;
  k_brick_search<BS_MODE_NN><<<brick_grid(nq), BS_THREADS, 0, s>>>(a); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_brick_forward(QueryDev q, int nq, IndexDev c, const uint32_t* bricks, const uint32_t* n_bricks, double max2,
                                 float max_d2f, int32_t* corr_p, float* corr_d2, uint32_t* rmin, const int* d_done, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  BrickArgs a{};
  a.q = q; a.c = c; a.bricks = bricks; a.n_bricks = n_bricks; a.done = d_done;
  a.max2 = max2; a.max_d2f = max_d2f; a.corr_p = corr_p; a.corr_d2 = corr_d2; a.rmin = rmin;
  k_brick_search<BS_MODE_FWD><<<brick_grid(nq), BS_THREADS, 0, s>>>(a); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_brick_reverse(QueryDev q, int nq, IndexDev c, const uint32_t* bricks, const uint32_t* n_bricks, uint32_t* rmin,
                                 int32_t* rnn, const int* d_done, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  BrickArgs a{};
  a.q = q; a.c = c; a.bricks = bricks; a.n_bricks = n_bricks; a.done = d_done;
  a.rmin = rmin; a.rnn = rnn; a.max_d2f = INFINITY;
  k_brick_search<BS_MODE_REV><<<brick_grid(nq), BS_THREADS, 0, s>>>(a); count_launch();
  return cudaGetLastError();
}

}  // namespace mvr
