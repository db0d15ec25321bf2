// icp.cu -- per-iteration ICP kernels: correspondence search with gate + reciprocal test (K4+K5),
// estimator sums (K6 point-to-point 3x3 cross-covariance, K9 point-to-plane 6x6 normal equations),
// fitness reduction (K8) and order-preserving correspondence compaction.
//
// Restates, for the GPU, what pcl::IterativeClosestPoint does per iteration inside icp.align()
// (reference call sites mvr/src/registrator.cpp:569, 920, 1012, 1024; semantics SURVEY.md A3-A8):
//   determine(Reciprocal)Correspondences -> TransformationEstimationSVD sums -> (host) SVD.
#include "launch.h"
#include "nn_search.cuh"

namespace mvr {

// ---------------------------------------------------------------------------------------------
// correspondences: source point i -> nearest target j, gate, optional reciprocal test
// ---------------------------------------------------------------------------------------------
template <bool RECIP, bool QIDX>
__global__ void __launch_bounds__(128) k_correspond(const float4* __restrict__ q, int nq, IndexDev tgt,
                                                    const float4* __restrict__ tgt_orig, IndexDev src, double max2,
                                                    float max_d2f, int32_t* __restrict__ corr_j, float* __restrict__ corr_d2) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nq) return;
  float4 p = __ldg(q + k);
  const int i = QIDX ? __float_as_int(p.w) : k;
  int j = -1;
  float d2 = MVR_INF;
  if (finite3(p)) {
    NnBest b{MVR_INF, 0x7fffffff};
    nn_search(tgt, p.x, p.y, p.z, max_d2f, b);
    if (b.idx != 0x7fffffff && !((double)b.d2 > max2)) {   // PCL: if (distance > max_dist_sqr) continue;
      bool keep = true;
      if (RECIP) {
        // nearest source point of the matched target point; seeded with (d2, i), which is exactly
        // what the search would compute for source point i (fsub(a,b) == -fsub(b,a)), so the result
        // is i iff no other source point is lexicographically closer.
        float4 t = __ldg(tgt_orig + b.idx);
        NnBest rb{b.d2, i};
        nn_search(src, t.x, t.y, t.z, b.d2, rb);
        keep = (rb.idx == i);
      }
      // -2-j marks "passed the gate, failed the reciprocal test" (counted as an answered query)
      j = keep ? b.idx : -2 - b.idx;
      d2 = b.d2;
    }
  }
  corr_j[i] = j;
  corr_d2[i] = d2;
}

cudaError_t launch_correspond(const float4* q, int nq, bool q_has_index, IndexDev tgt, const float4* tgt_orig,
                              IndexDev src, bool reciprocal, double max_dist2, float max_d2f, int32_t* corr_j,
                              float* corr_d2, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  dim3 grid((nq + 127) / 128), block(128);
  if (reciprocal) {
    if (q_has_index) k_correspond<true, true><<<grid, block, 0, s>>>(q, nq, tgt, tgt_orig, src, max_dist2, max_d2f, corr_j, corr_d2);
    else k_correspond<true, false><<<grid, block, 0, s>>>(q, nq, tgt, tgt_orig, src, max_dist2, max_d2f, corr_j, corr_d2);
  } else {
    if (q_has_index) k_correspond<false, true><<<grid, block, 0, s>>>(q, nq, tgt, tgt_orig, src, max_dist2, max_d2f, corr_j, corr_d2);
    else k_correspond<false, false><<<grid, block, 0, s>>>(q, nq, tgt, tgt_orig, src, max_dist2, max_d2f, corr_j, corr_d2);
  }
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// estimator sums: fixed topology (thread grid-stride -> warp shuffle tree -> 8 warps in order ->
// REDUCE_BLOCKS partials summed in block order), so results are identical run to run.
// ---------------------------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double* __restrict__ partials) {
  __shared__ double sm[8][NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < NV; ++a) {
    double x = v[a];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) sm[warp][a] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double x = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) x += sm[w][threadIdx.x];
    partials[(size_t)blockIdx.x * REDUCE_MAX_VALS + threadIdx.x] = x;
  }
}

__global__ void k_reduce_final(const double* __restrict__ partials, int nblk, int nv, double* __restrict__ out) {
  int t = threadIdx.x;
  if (t >= nv) return;
  double x = 0;
  for (int b = 0; b < nblk; ++b) x += partials[(size_t)b * REDUCE_MAX_VALS + t];
  out[t] = x;
}

// Point-to-point sums about `origin` o (a = s - o, b = t - o, exact in double):
// [0] n, [1..3] sum a, [4..6] sum b, [7..15] sum b_r * a_c (row r, col c), [16] sum d2,
// [17] number of source points that passed the distance gate (= reciprocal queries issued).
__global__ void __launch_bounds__(256) k_reduce_p2p(const float4* __restrict__ src, int n, const int32_t* __restrict__ corr_j,
                                                    const float* __restrict__ corr_d2, const float4* __restrict__ tgt,
                                                    double3 o, double* __restrict__ partials) {
  double v[REDUCE_P2P_VALS];
#pragma unroll
  for (int a = 0; a < REDUCE_P2P_VALS; ++a) v[a] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int j = __ldg(corr_j + i);
    if (j < -1) v[17] += 1.0;
    if (j < 0) continue;
    v[17] += 1.0;
    float4 s = __ldg(src + i), t = __ldg(tgt + j);
    double ax = (double)s.x - o.x, ay = (double)s.y - o.y, az = (double)s.z - o.z;
    double bx = (double)t.x - o.x, by = (double)t.y - o.y, bz = (double)t.z - o.z;
    v[0] += 1.0;
    v[1] += ax; v[2] += ay; v[3] += az;
    v[4] += bx; v[5] += by; v[6] += bz;
    v[7] += bx * ax; v[8] += bx * ay; v[9] += bx * az;
    v[10] += by * ax; v[11] += by * ay; v[12] += by * az;
    v[13] += bz * ax; v[14] += bz * ay; v[15] += bz * az;
    v[16] += (double)__ldg(corr_d2 + i);
  }
  block_reduce_store<REDUCE_P2P_VALS>(v, partials);
}

// Point-to-plane normal equations: J = [cross(s, n), n], r = n.(d - s).
// [0..20] upper triangle of J^T J row-major, [21..26] J^T r, [27] n, [28] sum d2, [29] gate-passing count.
__global__ void __launch_bounds__(256) k_reduce_p2l(const float4* __restrict__ src, int n, const int32_t* __restrict__ corr_j,
                                                    const float* __restrict__ corr_d2, const float4* __restrict__ tgt,
                                                    const float4* __restrict__ nrm, double3 o, double* __restrict__ partials) {
  (void)o;
  double v[REDUCE_P2L_VALS];
#pragma unroll
  for (int a = 0; a < REDUCE_P2L_VALS; ++a) v[a] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int j = __ldg(corr_j + i);
    if (j < -1) v[29] += 1.0;
    if (j < 0) continue;
    v[29] += 1.0;
    float4 s = __ldg(src + i), t = __ldg(tgt + j), nn = __ldg(nrm + j);
    double sx = s.x, sy = s.y, sz = s.z, nx = nn.x, ny = nn.y, nz = nn.z;
    double J[6] = {nz * sy - ny * sz, nx * sz - nz * sx, ny * sx - nx * sy, nx, ny, nz};
    double r = nx * ((double)t.x - sx) + ny * ((double)t.y - sy) + nz * ((double)t.z - sz);
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
      for (int b = a; b < 6; ++b) v[k++] += J[a] * J[b];
#pragma unroll
    for (int a = 0; a < 6; ++a) v[21 + a] += J[a] * r;
    v[27] += 1.0;
    v[28] += (double)__ldg(corr_d2 + i);
  }
  block_reduce_store<REDUCE_P2L_VALS>(v, partials);
}

__global__ void __launch_bounds__(256) k_reduce_fitness(const int32_t* __restrict__ idx, const float* __restrict__ d2, int n,
                                                        double max_range, double* __restrict__ partials) {
  double v[2] = {0.0, 0.0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float d = __ldg(d2 + i);
    if (__ldg(idx + i) >= 0 && (double)d <= max_range) { v[0] += (double)d; v[1] += 1.0; }
  }
  block_reduce_store<2>(v, partials);
}

cudaError_t launch_reduce_p2p(const float4* src_cur, int n, const int32_t* corr_j, const float* corr_d2,
                              const float4* tgt_orig, double3 origin, double* partials, double* out, cudaStream_t s) {
  k_reduce_p2p<<<REDUCE_BLOCKS, 256, 0, s>>>(src_cur, n, corr_j, corr_d2, tgt_orig, origin, partials); count_launch();
  k_reduce_final<<<1, 32, 0, s>>>(partials, REDUCE_BLOCKS, REDUCE_P2P_VALS, out); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_reduce_p2l(const float4* src_cur, int n, const int32_t* corr_j, const float* corr_d2,
                              const float4* tgt_orig, const float4* tgt_normals, double3 origin, double* partials,
                              double* out, cudaStream_t s) {
  k_reduce_p2l<<<REDUCE_BLOCKS, 256, 0, s>>>(src_cur, n, corr_j, corr_d2, tgt_orig, tgt_normals, origin, partials); count_launch();
  k_reduce_final<<<1, 32, 0, s>>>(partials, REDUCE_BLOCKS, REDUCE_P2L_VALS, out); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_reduce_fitness(const int32_t* idx, const float* d2, int n, double max_range, double* partials,
                                  double* out, cudaStream_t s) {
  k_reduce_fitness<<<REDUCE_BLOCKS, 256, 0, s>>>(idx, d2, n, max_range, partials); count_launch();
  k_reduce_final<<<1, 32, 0, s>>>(partials, REDUCE_BLOCKS, 2, out); count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// order-preserving compaction (ascending source index, like PCL's correspondence vector)
// ---------------------------------------------------------------------------------------------
constexpr int CP_THREADS = 256;
constexpr int CP_ITEMS = 4;
constexpr int CP_TILE = CP_THREADS * CP_ITEMS;

size_t compact_scratch_elems(int n) { return (size_t)(n > 0 ? (n + CP_TILE - 1) / CP_TILE : 1) + 1; }

__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* total) {
  __shared__ uint32_t ws[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { uint32_t c = ws[w]; if (w < warp) base += c; tot += c; }
  *total = tot;
  __syncthreads();
  return base + incl - v;
}

__global__ void __launch_bounds__(CP_THREADS) k_compact_count(const int32_t* __restrict__ corr_j, int n, uint32_t* __restrict__ counts) {
  int base = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
  uint32_t c = 0;
#pragma unroll
  for (int a = 0; a < CP_ITEMS; ++a) if (base + a < n && __ldg(corr_j + base + a) >= 0) ++c;
  uint32_t tot;
  block_excl_scan_256(c, &tot);
  if (threadIdx.x == 0) counts[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) k_compact_scan(uint32_t* __restrict__ counts, int nblk, uint32_t* __restrict__ total_out) {
  __shared__ uint32_t ws[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int b0 = 0; b0 < nblk; b0 += 1024) {
    int i = b0 + threadIdx.x;
    uint32_t v = i < nblk ? counts[i] : 0, incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) ws[warp] = incl;
    __syncthreads();
    uint32_t base = carry;
    for (int w = 0; w < warp; ++w) base += ws[w];
    if (i < nblk) counts[i] = base + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = base + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) { counts[nblk] = carry; *total_out = carry; }
}

__global__ void __launch_bounds__(CP_THREADS) k_compact_write(const int32_t* __restrict__ corr_j, const float* __restrict__ corr_d2,
                                                              int n, const uint32_t* __restrict__ offsets, int32_t* __restrict__ out_q,
                                                              int32_t* __restrict__ out_m, float* __restrict__ out_d2) {
  int base = blockIdx.x * CP_TILE + threadIdx.x * CP_ITEMS;
  int32_t j[CP_ITEMS];
  uint32_t c = 0;
#pragma unroll
  for (int a = 0; a < CP_ITEMS; ++a) { j[a] = (base + a < n) ? __ldg(corr_j + base + a) : -1; if (j[a] >= 0) ++c; }
  uint32_t tot;
  uint32_t pos = offsets[blockIdx.x] + block_excl_scan_256(c, &tot);
#pragma unroll
  for (int a = 0; a < CP_ITEMS; ++a)
    if (j[a] >= 0) { out_q[pos] = base + a; out_m[pos] = j[a]; out_d2[pos] = __ldg(corr_d2 + base + a); ++pos; }
}

cudaError_t launch_compact_corr(const int32_t* corr_j, const float* corr_d2, int n, uint32_t* scratch, int32_t* out_q,
                                int32_t* out_m, float* out_d2, uint32_t* count_out, cudaStream_t s) {
  int nblk = n > 0 ? (n + CP_TILE - 1) / CP_TILE : 0;
  if (nblk > 0) { k_compact_count<<<nblk, CP_THREADS, 0, s>>>(corr_j, n, scratch); count_launch(); }
  k_compact_scan<<<1, 1024, 0, s>>>(scratch, nblk, count_out); count_launch();
  if (nblk > 0) { k_compact_write<<<nblk, CP_THREADS, 0, s>>>(corr_j, corr_d2, n, scratch, out_q, out_m, out_d2); count_launch(); }
  return cudaGetLastError();
}

}  // namespace mvr
