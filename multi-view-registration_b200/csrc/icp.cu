// icp.cu -- per-iteration ICP kernels (the search itself is in search.cu): estimator sums (K6 point-to-point 3x3 cross-covariance, K9 point-to-plane 6x6 normal equations)
// with the solve and the convergence test in the reduction's last block, fitness reduction (K8) and
// order-preserving correspondence compaction.
//
// Restates, for the GPU, what pcl::IterativeClosestPoint does per iteration inside icp.align()
// (reference call sites mvr/src/registrator.cpp:569, 920, 1012, 1024; semantics SURVEY.md A3-A8):
//   determine(Reciprocal)Correspondences -> TransformationEstimationSVD -> transformCloud ->
//   DefaultConvergenceCriteria.
// The whole loop state lives in an IcpState on the device, so an align needs no host round trip per
// iteration: the host only enqueues iterations and reads the state back when a batch has run.
#include "launch.h"
#include "small_solve.h"

namespace mvr {

// ---------------------------------------------------------------------------------------------
// estimator sums: fixed topology (thread grid-stride over ORIGINAL indices -> warp shuffle tree ->
// 8 warps in order -> REDUCE_BLOCKS partials summed in block order by the last block to finish), so
// results are identical run to run and do not depend on the order of points inside grid cells.
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

// True in exactly one block per launch: the one that finished last.  On return in that block every
// block's partials are visible.
__device__ __forceinline__ bool last_block_done(unsigned int* ticket) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
    if (is_last) *ticket = 0;
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// Ordered sum of the blocks' partials: 8 chunks of consecutive blocks summed in parallel (each in
// block order), then the 8 chunk sums in chunk order.  Called by all 256 threads of the last block.
template <int NV>
__device__ __forceinline__ void ordered_total(const double* __restrict__ partials, int nblk, double* __restrict__ out) {
  __shared__ double ch[8][REDUCE_MAX_VALS];
  const int a = threadIdx.x & 31, c = threadIdx.x >> 5;
  const int per = (nblk + 7) / 8;
  if (a < NV) {
    const int b0 = c * per, b1 = min(b0 + per, nblk);
    double x = 0;
#pragma unroll 8
    for (int b = b0; b < b1; ++b) x += __ldcg(partials + (size_t)b * REDUCE_MAX_VALS + a);
    ch[c][a] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double x = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) x += ch[k][threadIdx.x];
    out[threadIdx.x] = x;
  }
  __syncthreads();
}

// The tail of one ICP iteration, run by one thread: estimate the increment from the sums, compose it
// into the accumulated transform, log, and evaluate DefaultConvergenceCriteria.
__device__ __noinline__ void icp_finish_iteration(IcpState* st, IterRec* log) {
  // everything the step needs, fetched up front so the loads overlap (st lives in global memory)
  double S[REDUCE_MAX_VALS], F0[16];
#pragma unroll
  for (int k = 0; k < REDUCE_MAX_VALS; ++k) S[k] = st->sums[k];
#pragma unroll
  for (int k = 0; k < 16; ++k) F0[k] = st->fin[k];
  const bool p2l = st->p2l != 0;
  const double ox = st->ox, oy = st->oy, oz = st->oz;
  const int min_corr = st->min_corr, max_iter = st->max_iter, fixed = st->fixed, iter0 = st->iter;
  const double prev_mse = st->prev_mse, rot_thr = st->rot_thr, trans_thr = st->trans_thr, fit_eps = st->fit_eps;
  const double cnt = p2l ? S[27] : S[0];
  const double d2sum = p2l ? S[28] : S[16];
  st->queries += (unsigned long long)st->n_src + (st->recip ? (unsigned long long)(p2l ? S[29] : S[17]) : 0ull);
  const int n_corr = (int)cnt;
  st->n_corr = n_corr;
  if (n_corr < min_corr) { st->reason = 5; st->status = 2; st->done = 1; return; }  // "Not enough correspondences"
  double T[16];
  if (p2l) {
    double A[36], b[6], x[6];
    int k = 0;
    for (int a = 0; a < 6; ++a)
      for (int c = a; c < 6; ++c) { A[a * 6 + c] = S[k]; A[c * 6 + a] = S[k]; ++k; }
    for (int a = 0; a < 6; ++a) b[a] = S[21 + a];
    if (!cholesky_solve6(A, b, x)) { st->reason = 5; st->status = 5; st->done = 1; return; }
    pose_from_6(x, T);
  } else {
    const double inv = 1.0 / cnt;
    const double mu_a[3] = {S[1] * inv, S[2] * inv, S[3] * inv}, mu_b[3] = {S[4] * inv, S[5] * inv, S[6] * inv};
    double Sg[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) Sg[r * 3 + c] = S[7 + r * 3 + c] * inv - mu_b[r] * mu_a[c];
    const double mu_s[3] = {mu_a[0] + ox, mu_a[1] + oy, mu_a[2] + oz};
    const double mu_d[3] = {mu_b[0] + ox, mu_b[1] + oy, mu_b[2] + oz};
    umeyama_rigid(mu_s, mu_d, Sg, T);
  }
  float Tf[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) Tf[k] = (float)T[k];
  Tf[3] = Tf[7] = Tf[11] = 0.f; Tf[15] = 1.f;
  double Td[16], F[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) Td[k] = Tf[k];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double x = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) x += Td[k * 4 + r] * F0[c * 4 + k];
      F[c * 4 + r] = x;
    }
#pragma unroll
  for (int k = 0; k < 16; ++k) { st->fin[k] = F[k]; st->delta[k] = Tf[k]; }
  const int iter = iter0 + 1;
  st->iter = iter;
  const double cur_mse = d2sum / cnt;
  st->cur_mse = cur_mse;
  if (iter <= ICP_MAX_LOG) {
    IterRec& r = log[iter - 1];
    r.iteration = iter; r.n_corr = n_corr; r.mse = cur_mse;
#pragma unroll
    for (int k = 0; k < 16; ++k) r.delta[k] = Tf[k];
  }
  if (iter >= max_iter) { st->done = 1; st->reason = 1; return; }
  if (!fixed) {
    const double cos_angle = 0.5 * (Td[0] + Td[5] + Td[10] - 1.0);
    const double t2 = Td[12] * Td[12] + Td[13] * Td[13] + Td[14] * Td[14];
    if (cos_angle >= rot_thr && t2 <= trans_thr) { st->done = 1; st->reason = 2; return; }
    if (fabs(cur_mse - prev_mse) < 1e-12) { st->done = 1; st->reason = 3; return; }
    if (fabs(cur_mse - prev_mse) / prev_mse < fit_eps) { st->done = 1; st->reason = 4; return; }
    st->prev_mse = cur_mse;
  }
}

// Point-to-point sums about the origin o (a = s - o, b = t - o, exact in double):
// [0] n, [1..3] sum a, [4..6] sum b, [7..15] sum b_r * a_c (row r, col c), [16] sum d2,
// [17] number of source points that passed the distance gate (= reciprocal queries when enabled).
__global__ void __launch_bounds__(256) k_reduce_p2p(const float4* __restrict__ src, int n, const int32_t* __restrict__ corr_p,
                                                    const float* __restrict__ corr_d2, const int32_t* __restrict__ rnn,
                                                    const float4* __restrict__ tgt, double* __restrict__ partials,
                                                    IcpState* __restrict__ st, IterRec* __restrict__ log) {
  if (st->done) return;
  const double ox = st->ox, oy = st->oy, oz = st->oz;
  double v[REDUCE_P2P_VALS];
#pragma unroll
  for (int a = 0; a < REDUCE_P2P_VALS; ++a) v[a] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int p = __ldg(corr_p + i);
    if (p < 0) continue;
    v[17] += 1.0;
    if (rnn && __ldg(rnn + p) != i) continue;   // reciprocal test: the matched point's nearest source point is not i
    const float4 s = __ldg(src + i);
    const float4 t = __ldg(tgt + p);
    double ax = (double)s.x - ox, ay = (double)s.y - oy, az = (double)s.z - oz;
    double bx = (double)t.x - ox, by = (double)t.y - oy, bz = (double)t.z - oz;
    v[0] += 1.0;
    v[1] += ax; v[2] += ay; v[3] += az;
    v[4] += bx; v[5] += by; v[6] += bz;
    v[7] += bx * ax; v[8] += bx * ay; v[9] += bx * az;
    v[10] += by * ax; v[11] += by * ay; v[12] += by * az;
    v[13] += bz * ax; v[14] += bz * ay; v[15] += bz * az;
    v[16] += (double)__ldg(corr_d2 + i);
  }
  block_reduce_store<REDUCE_P2P_VALS>(v, partials);
  if (!last_block_done(&st->ticket)) return;
  ordered_total<REDUCE_P2P_VALS>(partials, gridDim.x, st->sums);
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    icp_finish_iteration(st, log);
    st->dbg[0] += clock64() - t0; st->dbg[1] += 1;
  }
}

// Point-to-plane normal equations: J = [cross(s, n), n], r = n.(d - s).
// [0..20] upper triangle of J^T J row-major, [21..26] J^T r, [27] n, [28] sum d2, [29] gate-passing count.
__global__ void __launch_bounds__(256) k_reduce_p2l(const float4* __restrict__ src, int n, const int32_t* __restrict__ corr_p,
                                                    const float* __restrict__ corr_d2, const int32_t* __restrict__ rnn,
                                                    const float4* __restrict__ tgt, const float4* __restrict__ nrm,
                                                    double* __restrict__ partials, IcpState* __restrict__ st, IterRec* __restrict__ log) {
  if (st->done) return;
  double v[REDUCE_P2L_VALS];
#pragma unroll
  for (int a = 0; a < REDUCE_P2L_VALS; ++a) v[a] = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int p = __ldg(corr_p + i);
    if (p < 0) continue;
    v[29] += 1.0;
    if (rnn && __ldg(rnn + p) != i) continue;
    const float4 s = __ldg(src + i);
    const float4 t = __ldg(tgt + p), nn = __ldg(nrm + __float_as_int(t.w));
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
  if (!last_block_done(&st->ticket)) return;
  ordered_total<REDUCE_P2L_VALS>(partials, gridDim.x, st->sums);
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    icp_finish_iteration(st, log);
    st->dbg[0] += clock64() - t0; st->dbg[1] += 1;
  }
}

__global__ void k_reduce_final(const double* __restrict__ partials, int nblk, int nv, double* __restrict__ out) {
  int t = threadIdx.x;
  if (t >= nv) return;
  double x = 0;
#pragma unroll 8
  for (int b = 0; b < nblk; ++b) x += partials[(size_t)b * REDUCE_MAX_VALS + t];
  out[t] = x;
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

__global__ void __launch_bounds__(256) k_transform_final(const float4* __restrict__ in, float4* __restrict__ out, int n,
                                                         const IcpState* __restrict__ st) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Mat4f M;
#pragma unroll
  for (int k = 0; k < 16; ++k) M.m[k] = (float)st->fin[k];
  float4 p = __ldg(in + i);
  out[i] = finite3(p) ? xform_pinned(M, p) : p;
}

cudaError_t launch_reduce_solve(const float4* src_cur, int n, const int32_t* corr_p, const float* corr_d2, const int32_t* rnn,
                                const float4* tgt_sorted, const float4* tgt_normals, double* partials, IcpState* st,
                                IterRec* log, bool p2l, cudaStream_t s) {
  if (p2l) k_reduce_p2l<<<REDUCE_BLOCKS, 256, 0, s>>>(src_cur, n, corr_p, corr_d2, rnn, tgt_sorted, tgt_normals, partials, st, log);
  else k_reduce_p2p<<<REDUCE_BLOCKS, 256, 0, s>>>(src_cur, n, corr_p, corr_d2, rnn, tgt_sorted, partials, st, log);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_transform_final(const float4* in, float4* out, int n, const IcpState* st, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_transform_final<<<(n + 255) / 256, 256, 0, s>>>(in, out, n, st); count_launch();
  return cudaGetLastError();
}

cudaError_t launch_reduce_fitness(const int32_t* idx, const float* d2, int n, double max_range, double* partials,
                                  double* out, cudaStream_t s) {
  k_reduce_fitness<<<REDUCE_BLOCKS, 256, 0, s>>>(idx, d2, n, max_range, partials); count_launch();
  k_reduce_final<<<1, 32, 0, s>>>(partials, REDUCE_BLOCKS, 2, out); count_launch();
  return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_resolve_corr(const int32_t* __restrict__ corr_p, const int32_t* __restrict__ rnn,
                                                      const float4* __restrict__ tgt, int n, int32_t* __restrict__ corr_j) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = __ldg(corr_p + i);
  int j = -1;
  if (p >= 0) {
    j = __float_as_int(__ldg(tgt + p).w);
    if (rnn && __ldg(rnn + p) != i) j = -2 - j;
  }
  corr_j[i] = j;
}

cudaError_t launch_resolve_corr(const int32_t* corr_p, const int32_t* rnn, const float4* tgt_sorted, int n, int32_t* corr_j,
                                cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  k_resolve_corr<<<(n + 255) / 256, 256, 0, s>>>(corr_p, rnn, tgt_sorted, n, corr_j); count_launch();
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
