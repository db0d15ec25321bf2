// nn.cu -- exact 1-NN query kernel (K4) over the Morton-sorted uniform grid.
// Replaces pcl::KdTreeFLANN::nearestKSearch(k=1) as called for every source point by icp.align,
// determineReciprocalCorrespondences and getFitnessScore (mvr/src/registrator.cpp:502, 569, 572, 649).
//
// Algorithmic bytes per query: 16 B query + 8 B result, plus each target point
// (16 B) and each cell-table entry (4 B) once per launch.
#include "launch.h"
#include "nn_search.cuh"

namespace mvr {

template <bool QIDX>
__global__ void __launch_bounds__(128) k_nn_query(const float4* __restrict__ q, int nq, IndexDev ix, float max_d2f,
                                                  int32_t* __restrict__ out_idx, float* __restrict__ out_d2) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nq) return;
  float4 p = __ldg(q + k);
  const int i = QIDX ? __float_as_int(p.w) : k;
  NnBest b{MVR_INF, 0x7fffffff, -1};
  if (finite3(p)) nn_search(ix, p.x, p.y, p.z, max_d2f, b);
  bool found = b.idx != 0x7fffffff;
  out_idx[i] = found ? b.idx : -1;
  out_d2[i] = found ? b.d2 : MVR_INF;
}

cudaError_t launch_nn_query(const float4* q, int nq, bool q_has_index, IndexDev ix, float max_d2, int32_t* out_idx,
                            float* out_d2, cudaStream_t s) {
  if (nq <= 0) return cudaSuccess;
  if (q_has_index) k_nn_query<true><<<(nq + 127) / 128, 128, 0, s>>>(q, nq, ix, max_d2, out_idx, out_d2);
  else k_nn_query<false><<<(nq + 127) / 128, 128, 0, s>>>(q, nq, ix, max_d2, out_idx, out_d2);
  count_launch();
  return cudaGetLastError();
}

}  // namespace mvr
