// denoise.cu -- PointCloud::denoise (mvr/src/point_cloud.cpp:423-514) on the GPU: drop the connected components with fewer than
// segment_threshold points of the graph whose edges join points at most triangle_length apart.
//
// The reference builds that graph from a CGAL Delaunay triangulation (initPointGraph, :469-497: every finite Delaunay edge no
// longer than the threshold) and labels it with boost::connected_components (:434-435).  The Euclidean minimum spanning tree is
// a subgraph of the Delaunay triangulation, and the components of "all pairs within L" are exactly those of the spanning-tree
// edges within L (single linkage) -- so the short Delaunay edges and the plain radius graph have the SAME connected components,
// and no triangulation is needed: every point unions itself with the points within triangle_length found in the 27 cells
// around it (row-major grid of pair_index.cu, cell edge = triangle_length), lock-free union-find hooking the larger root under
// the smaller, so a component's root is its smallest point index.
//
// Output order is the reference's: components in the order boost discovers them (ascending smallest point index), points of
// a component in ascending index (:441-447) -- a stable sort of the kept points by root.
// Distances follow the reference: sqrt of the double squared distance of the float coordinates, an edge iff it is <= threshold.
#include "launch.h"
#include "pair_search.cuh"

namespace mvr {

__device__ __forceinline__ uint32_t dn_find(uint32_t* __restrict__ parent, uint32_t i) {
  // (L2 loads: a root another block has just hooked must not be served stale from this SM's L1 for ever)
  uint32_t p = __ldcg(parent + i);
  while (p != i) {
    const uint32_t g = __ldcg(parent + p);
    if (g != p) parent[i] = g;   // path halving (benign race: any ancestor is a valid parent)
    i = p; p = g;
  }
  return i;
}

__device__ __forceinline__ void dn_union(uint32_t* __restrict__ parent, uint32_t a, uint32_t b) {
  for (;;) {
    a = dn_find(parent, a);
    b = dn_find(parent, b);
    if (a == b) return;
    if (a < b) { const uint32_t t = a; a = b; b = t; }       // a = larger root: hook it under the smaller one
    const uint32_t old = atomicCAS(parent + a, a, b);
    if (old == a) return;
  }
}

__global__ void __launch_bounds__(256) k_dn_init(uint32_t* __restrict__ parent, uint32_t* __restrict__ count, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { parent[i] = (uint32_t)i; count[i] = 0u; }
}

// One thread per sorted point: union with every point within `thr` that sits at a LOWER sorted position (each edge once).
__global__ void __launch_bounds__(128) k_dn_edges(const float4* __restrict__ sorted, const uint32_t* __restrict__ start, PairGrid g, int n_valid,
                                                  double thr, uint32_t* __restrict__ parent) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_valid) return;
  const float4 p = __ldg(sorted + k);
  const uint32_t me = __float_as_uint(p.w);
  const int cx = pg_cell(grid_t(p.x, g.ox, g.inv_cell), g.nx), cy = pg_cell(grid_t(p.y, g.oy, g.inv_cell), g.ny), cz = pg_cell(grid_t(p.z, g.oz, g.inv_cell), g.nz);
  for (int z = max(cz - 1, 0); z <= min(cz + 1, g.nz - 1); ++z)
    for (int y = max(cy - 1, 0); y <= min(cy + 1, g.ny - 1); ++y) {
      const uint32_t* row = start + ((size_t)z * g.ny + y) * g.nx;
      const uint32_t s = __ldg(row + max(cx - 1, 0)), e = min(__ldg(row + min(cx + 1, g.nx - 1) + 1), (uint32_t)k);
      for (uint32_t j = s; j < e; ++j) {
        const float4 q = __ldg(sorted + j);
        const double dx = (double)p.x - (double)q.x, dy = (double)p.y - (double)q.y, dz = (double)p.z - (double)q.z;
        if (sqrt(dx * dx + dy * dy + dz * dz) > thr) continue;     // reference: if (distance > threshold) continue
        dn_union(parent, me, __float_as_uint(q.w));
      }
    }
}

// root[i] goes to its own array: parent[i] may still be rewritten (to an ancestor that is not the root) by the path halving
// of another thread's find that passes through i.
__global__ void __launch_bounds__(256) k_dn_count(uint32_t* __restrict__ parent, uint32_t* __restrict__ root, uint32_t* __restrict__ count, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = dn_find(parent, (uint32_t)i);
  root[i] = r;
  atomicAdd(count + r, 1u);
}

// in: keys[i] = root of point i; out: key = root for a kept point, 0xffffffff for a dropped one; val = point index
__global__ void __launch_bounds__(256) k_dn_keys(const uint32_t* __restrict__ count, int n, uint32_t threshold,
                                                 uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t* __restrict__ n_noise) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = keys[i];
  const bool keep = count[r] >= threshold;
  keys[i] = keep ? r : 0xffffffffu;
  vals[i] = (uint32_t)i;
  const unsigned int bal = __ballot_sync(__activemask(), !keep);
  if ((threadIdx.x & 31) == (__ffs(__activemask()) - 1) && bal) atomicAdd(n_noise, (uint32_t)__popc(bal));
}

cudaError_t launch_denoise_components(const float4* sorted, const uint32_t* start, PairGrid g, int n, int n_valid, double threshold,
                                      uint32_t* parent, uint32_t* root, uint32_t* count, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_dn_init<<<(n + 255) / 256, 256, 0, s>>>(parent, count, n); count_launch();
  if (n_valid > 0) { k_dn_edges<<<(n_valid + 127) / 128, 128, 0, s>>>(sorted, start, g, n_valid, threshold, parent); count_launch(); }
  k_dn_count<<<(n + 255) / 256, 256, 0, s>>>(parent, root, count, n); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_denoise_keys(const uint32_t* count, int n, uint32_t segment_threshold, uint32_t* keys, uint32_t* vals,
                                uint32_t* n_noise, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_dn_keys<<<(n + 255) / 256, 256, 0, s>>>(count, n, segment_threshold, keys, vals, n_noise); count_launch();
  return cudaGetLastError();
}

}  // namespace mvr
