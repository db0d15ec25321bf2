// pair_index.cu -- the per-align spatial index of the fused ICP loop: row-major uniform grid, built ONCE.
//
// pcl::IterativeClosestPoint builds the target kd-tree once per align and, with reciprocal
// correspondences, a source kd-tree in EVERY iteration (enabled by the reference at
// mvr/src/registrator.cpp:552, 768, 901).  Here both clouds are binned once per align: the target in a
// grid over its own box, the source -- after the initial guess has been applied -- in a grid over the
// guessed source box.  The source then moves rigidly every iteration but keeps its cells: the search
// (pair_search.cuh) carries the query into the binning frame instead of re-sorting 200k points 30 times.
//
//   k_pair_keys    : p <- guess * p (pinned float transform, ICP's transformCloud) when a
//                    guess is given; key = row-major cell (z, y, x; x fastest), non-finite points get the
//                    sentinel key `cells`; value = original index
//   radix sort     : the stable LSD sort of index.cu on ceil(log2(cells + 1)) bits -> points of a cell
//                    stay in ascending original index, so every sorted position, and with it the order
//                    of every later reduction, is reproducible run to run
//   k_pair_gather  : sorted[k] = {moved point, bits(original index)} (+ a second copy: the source keeps
//                    its binning-time coordinates s0 next to the coordinates that move)
//   k_cell_table   : start[c] = first sorted position of cell c (index.cu)
//
// Algorithmic bytes per point: 16 read + 16 (32 with the s0 copy) written, + 4 B per cell.
#include "launch.h"
#include "pair_search.cuh"

namespace mvr {

__global__ void __launch_bounds__(256) k_pair_keys(const float4* __restrict__ in, int n, Mat4f M, int apply, PairGrid g, uint32_t cells,
                                                   float4* __restrict__ moved, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(in + i);
  const bool ok = finite3(p);
  if (apply && ok) p = xform_pinned(M, p);
  if (moved) moved[i] = p;
  keys[i] = ok ? pg_key(p, g) : cells;
  vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_pair_gather(const float4* __restrict__ pts, const uint32_t* __restrict__ perm, int n,
                                                     float4* __restrict__ sorted, float4* __restrict__ copy) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t j = __ldg(perm + k);
  float4 p = __ldg(pts + j);
  p.w = __uint_as_float(j);
  sorted[k] = p;
  if (copy) copy[k] = p;
}

cudaError_t launch_pair_keys(const float4* in, int n, const Mat4f* guess, PairGrid g, uint32_t cells, float4* moved, uint32_t* keys,
                             uint32_t* vals, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  Mat4f M{};
  if (guess) M = *guess;
  k_pair_keys<<<(n + 255) / 256, 256, 0, s>>>(in, n, M, guess ? 1 : 0, g, cells, moved, keys, vals); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_pair_gather(const float4* pts, const uint32_t* perm, int n, float4* sorted, float4* copy, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_pair_gather<<<(n + 255) / 256, 256, 0, s>>>(pts, perm, n, sorted, copy); count_launch();
  return cudaGetLastError();
}

}  // namespace mvr
