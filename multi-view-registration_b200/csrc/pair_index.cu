// pair_index.cu -- the per-align spatial index of the fused ICP loop: row-major uniform grid, built ONCE.
//
// pcl::IterativeClosestPoint builds the target kd-tree once per align and, with reciprocal
// correspondences, a source kd-tree in EVERY iteration (enabled by the reference at
// mvr/src/registrator.cpp:552, 768, 901).  Here both clouds are binned once per align: the target in a
// grid over its own box, the source -- after the initial guess has been applied -- in a grid over the
// guessed source box.  The source then moves rigidly every iteration but keeps its cells: the search
// (pair_search.cuh) carries the query into the binning frame instead of re-sorting 200k points 30 times.
//
// The build is a counting sort made deterministic (four launches, no multi-pass radix sort):
//   k_pair_count   : p <- guess * p (pinned float transform, ICP's transformCloud) when a guess is
//                    given; key = row-major cell (z, y, x; x fastest), non-finite points get the sentinel key
//                    `cells`; rank = the point's arrival number in its cell (one atomicAdd on the cell counter)
//   k_scan_cells   : exclusive scan of the counters -> start[c] = first sorted position of cell c (bin.cu)
//   k_pair_scatter : tmp[start[key] + rank] = {moved point, bits(original index)} -- arrival order inside a cell
//   k_pair_rerank  : the final position of a point = start[key] + number of points of its cell with a SMALLER
//                    original index (its cell is ~10 contiguous records of tmp), so points of a cell end up in
//                    ascending original index whatever order the atomics ran in: every sorted position, and with
//                    it the order of every later reduction, is reproducible run to run.
//
// Algorithmic bytes per point: 16 read + 16 written, + 8 B per cell.
#include <algorithm>

#include "launch.h"
#include "pair_search.cuh"

namespace mvr {

// One launch per phase serves EVERY build of a batch (blockIdx.y = job; the jobs travel by value in the kernel's
// parameter space): a 24-pair registration builds 48 indices with 4 launches instead of 192.
__device__ __forceinline__ const BuildJob& load_job(const BuildBatch& b, BuildJob& s_job) {
  static_assert(sizeof(BuildJob) % 4 == 0 && sizeof(BuildJob) / 4 <= 256, "job block");
  if (threadIdx.x < sizeof(BuildJob) / 4) ((uint32_t*)&s_job)[threadIdx.x] = ((const uint32_t*)&b.j[blockIdx.y])[threadIdx.x];
  __syncthreads();
  return s_job;
}

__global__ void __launch_bounds__(256) k_pair_count(const __grid_constant__ BuildBatch b) {
  __shared__ BuildJob s_job;
  const BuildJob& j = load_job(b, s_job);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.n) return;
  float4 p = __ldg(j.in + i);
  const bool ok = finite3(p);
  if (j.apply && ok) p = xform_pinned(j.M, p);
  const uint32_t key = ok ? pg_key(p, j.g) : j.cells;
  j.keys[i] = key;
  j.rank[i] = atomicAdd(j.counters + key, 1u);
}

__global__ void __launch_bounds__(256) k_pair_scatter(const __grid_constant__ BuildBatch b) {
  __shared__ BuildJob s_job;
  const BuildJob& j = load_job(b, s_job);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.n) return;
  float4 p = __ldg(j.in + i);
  if (j.apply && finite3(p)) p = xform_pinned(j.M, p);   // the same pinned arithmetic as k_pair_count: the same point
  p.w = __uint_as_float((uint32_t)i);
  (j.ordered ? j.tmp : j.sorted)[__ldg(j.start + __ldg(j.keys + i)) + __ldg(j.rank + i)] = p;
}

// A cell that holds more than RERANK_LIMIT points (duplicated or invalid returns piled on one spot, outliers clamped into a
// boundary cell) would make the rank count below quadratic: its points keep their arrival order instead and the population is
// reported through *crowded (searches stay exact; only the run-to-run order of sums over that cell's points is lost).
constexpr uint32_t RERANK_LIMIT = 1024;
__global__ void __launch_bounds__(256) k_pair_rerank(const __grid_constant__ BuildBatch b) {
  __shared__ BuildJob s_job;
  const BuildJob& j = load_job(b, s_job);
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= j.n || !j.ordered) return;
  const float4* __restrict__ tmp = j.tmp;
  const float4 p = __ldg(tmp + k);
  const uint32_t idx = __float_as_uint(p.w);
  const uint32_t key = __ldg(j.keys + idx);
  const uint32_t s = __ldg(j.start + key), e = __ldg(j.start + key + 1);
  if (e - s > RERANK_LIMIT) {
    j.sorted[k] = p;
    if (j.crowded && (uint32_t)k == s) atomicMax(j.crowded, e - s);
    return;
  }
  uint32_t r = 0;
  for (uint32_t q = s; q < e; ++q) r += (__float_as_uint(__ldg(&tmp[q].w)) < idx) ? 1u : 0u;
  j.sorted[s + r] = p;
}

// ---------------------------------------------------------------------------------------------
// Gate mask of a target index: bit x of word [(z * ny + y) * wstride + x / 32] is set when some target point lies in a
// cell within D cells in y and z and Dx cells in x of cell (x, y, z).  Cells that differ by more than D in some axis are at least D cells
// apart along it, so with D cell >= gate a CLEAR bit proves that a query filed in that cell has no target point within the
// gate: pcl's "if (distance > max_dist_sqr) continue"  without a search.  The part of a turntable view the
// neighbouring view does not cover is settled by one load per point and iteration.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gate_occ(const uint32_t* __restrict__ start, PairGrid g, int wstride, size_t words, uint32_t* __restrict__ occ) {
  const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words) return;
  const size_t row = w / (size_t)wstride;
  const int xw = (int)(w % (size_t)wstride);
  const uint32_t* r = start + row * (size_t)g.nx;
  uint32_t bits = 0;
  uint32_t prev = __ldg(r + xw * 32);
  for (int b = 0; b < 32; ++b) {
    const int x = xw * 32 + b;
    if (x >= g.nx) break;
    const uint32_t next = __ldg(r + x + 1);
    if (next > prev) bits |= 1u << b;
    prev = next;
  }
  occ[w] = bits;
}

__global__ void __launch_bounds__(256) k_gate_dilate(const uint32_t* __restrict__ occ, PairGrid g, int wstride, int D, int Dx, size_t words, uint32_t* __restrict__ out) {
  const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words) return;
  const int xw = (int)(w % (size_t)wstride);
  const size_t row = w / (size_t)wstride;
  const int y = (int)(row % (size_t)g.ny), z = (int)(row / (size_t)g.ny);
  uint32_t acc = 0;
  for (int zz = max(z - D, 0); zz <= min(z + D, g.nz - 1); ++zz)
    for (int yy = max(y - D, 0); yy <= min(y + D, g.ny - 1); ++yy) {
      const uint32_t* r = occ + ((size_t)zz * g.ny + yy) * wstride;
      const uint32_t cur = __ldg(r + xw), prev = xw > 0 ? __ldg(r + xw - 1) : 0u, next = xw + 1 < wstride ? __ldg(r + xw + 1) : 0u;
      uint32_t v = cur;
      for (int k = 1; k <= Dx; ++k) v |= (cur << k) | (cur >> k) | (prev >> (32 - k)) | (next << (32 - k));
      acc |= v;
    }
  out[w] = acc;
}

cudaError_t launch_gate_mask(const uint32_t* start, PairGrid g, int wstride, int D, int Dx, uint32_t* occ, uint32_t* mask, cudaStream_t s) {
  const size_t words = (size_t)g.ny * g.nz * wstride;
  if (words == 0 || D < 1 || Dx < 1 || Dx > 31) return cudaErrorInvalidValue;
  const unsigned blocks = (unsigned)((words + 255) / 256);
  k_gate_occ<<<blocks, 256, 0, s>>>(start, g, wstride, words, occ); count_launch();
  k_gate_dilate<<<blocks, 256, 0, s>>>(occ, g, wstride, D, Dx, words, mask); count_launch();
  return cudaGetLastError();
}

// The four phases of `count` index builds (<= BUILD_MAX_JOBS) on stream s.
cudaError_t launch_pair_builds(const BuildBatch& batch, int count, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  if (count > BUILD_MAX_JOBS) return cudaErrorInvalidValue;
  int max_n = 0, max_tiles = 0;
  bool any_ordered = false;
  for (int k = 0; k < count; ++k) {
    max_n = std::max(max_n, batch.j[k].n);
    max_tiles = std::max(max_tiles, batch.j[k].ntiles);
    any_ordered = any_ordered || batch.j[k].ordered;
  }
  const dim3 gp((unsigned)std::max((max_n + 255) / 256, 1), (unsigned)count);
  if (max_n > 0) { k_pair_count<<<gp, 256, 0, s>>>(batch); count_launch(); }
  launch_scan_cells_batch(batch, count, max_tiles, s);
  if (max_n > 0) { k_pair_scatter<<<gp, 256, 0, s>>>(batch); count_launch(); }
  if (max_n > 0 && any_ordered) { k_pair_rerank<<<gp, 256, 0, s>>>(batch); count_launch(); }
  return cudaGetLastError();
}

}  // namespace mvr
