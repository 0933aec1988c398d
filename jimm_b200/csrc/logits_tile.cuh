// fp32 64x64 logits tile shared by the single-GPU logits kernel (elementwise.cu) and the fused
// normalise + peer-scatter + logits kernel (comm.cu).  256 threads, 4x4 micro-tile per thread, K step 16.
#pragma once
#include <cuda_runtime.h>

namespace jimm {

// out[i, j] = sc * <A[i, :E], B[j, :E]> + bs   for the 64x64 tile at (i0, j0); row strides lda / ldb / ldl (elements).
// VOLATILE_B: read B with ld.global.cg (data written by peer GPUs into local memory; bypass L1).
template <bool CG_LOADS>
__device__ __forceinline__ void logits_tile(const float* __restrict__ A, size_t lda, const float* __restrict__ B, size_t ldb,
                                            float* __restrict__ out, size_t ldl, int Bi, int Bt, int E, int i0, int j0, float sc,
                                            float bs, float (*As)[65], float (*Bs)[65]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < E; k0 += 16) {
    for (int l = threadIdx.x; l < 64 * 16; l += 256) {
      const int r = l >> 4, k = l & 15;
      float a = 0.f, b = 0.f;
      if (i0 + r < Bi && k0 + k < E) a = CG_LOADS ? __ldcg(A + static_cast<size_t>(i0 + r) * lda + k0 + k) : A[static_cast<size_t>(i0 + r) * lda + k0 + k];
      if (j0 + r < Bt && k0 + k < E) b = CG_LOADS ? __ldcg(B + static_cast<size_t>(j0 + r) * ldb + k0 + k) : B[static_cast<size_t>(j0 + r) * ldb + k0 + k];
      As[k][r] = a;
      Bs[k][r] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = As[k][ty * 4 + u]; b[u] = Bs[k][tx * 4 + u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int w = 0; w < 4; ++w) acc[u][w] = fmaf(a[u], b[w], acc[u][w]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int i = i0 + ty * 4 + u, j = j0 + tx * 4 + w;
      if (i < Bi && j < Bt) out[static_cast<size_t>(i) * ldl + j] = sc * acc[u][w] + bs;
    }
}

}  // namespace jimm
