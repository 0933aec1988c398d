// Zero-shot / classification epilogue: what the reference's examples do in JAX after the forward path.
//   examples/clip_inference.py:46-51   scores = logits[0]; softmax = exp(scores) / sum(exp(scores)); order = argsort(scores)[::-1]
//   examples/vit_inference.py:58       predicted = argmax(logits, -1)
//   SigLIP (sigmoid loss, models/siglip.py:169-174 logits + bias): per-pair probability = sigmoid(logit)
// One CTA per row of logits: probabilities (fp32, the example's un-shifted exp / sum), the full descending order and the argmax.
// Integer outputs are exact, including ties: argsort is stable ascending and then reversed, so equal scores come out with the
// LARGER index first; argmax returns the first maximum.
#include "../../include/jimm_b200.h"
#include "common.cuh"

namespace jimm {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxCols = 4096;

// Monotonic map float -> uint32 (larger float = larger key); -0 is folded into +0 and every NaN into the canonical quiet NaN,
// which sorts above +inf (numpy / jnp sort NaN last in ascending order, i.e. first once reversed).
__device__ __forceinline__ uint32_t order_key(float v) {
  uint32_t u = __float_as_uint(v);
  if (v != v) u = 0x7fc00000u;
  if (u == 0x80000000u) u = 0u;
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(kThreads) postprocess_kernel(const float* __restrict__ logits, int cols, int ld, int mode, float* __restrict__ probs,
                                                             int ldp, int32_t* __restrict__ order, int32_t* __restrict__ argmax, int npow2) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);  // [npow2] (key << 32 | index), sorted descending
  __shared__ float red[kThreads / 32];
  __shared__ float total;
  const int row = blockIdx.x, tid = threadIdx.x;
  const float* x = logits + static_cast<size_t>(row) * ld;

  // ---- probabilities ----
  if (probs) {
    float* p = probs + static_cast<size_t>(row) * ldp;
    if (mode == 1) {
      for (int i = tid; i < cols; i += kThreads) p[i] = 1.0f / (1.0f + expf(-x[i]));
    } else {
      float s = 0.f;
      for (int i = tid; i < cols; i += kThreads) s += expf(x[i]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((tid & 31) == 0) red[tid >> 5] = s;
      __syncthreads();
      if (tid == 0) {
        float t = 0.f;
        for (int w = 0; w < kThreads / 32; ++w) t += red[w];
        total = t;
      }
      __syncthreads();
      const float t = total;
      for (int i = tid; i < cols; i += kThreads) p[i] = expf(x[i]) / t;
    }
  }
  if (!order && !argmax) return;

  // ---- descending order by (score, index): bitonic sort of 64-bit keys in shared memory ----
  for (int i = tid; i < npow2; i += kThreads)
    keys[i] = i < cols ? (static_cast<unsigned long long>(order_key(x[i])) << 32) | static_cast<unsigned>(i) : 0ull;  // padding sorts last
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < npow2; i += kThreads) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], b = keys[l];
          const bool desc = (i & k) == 0;
          if (desc ? a < b : a > b) {
            keys[i] = b;
            keys[l] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  if (order)
    for (int i = tid; i < cols; i += kThreads) order[static_cast<size_t>(row) * cols + i] = static_cast<int32_t>(keys[i] & 0xffffffffu);
  if (argmax && tid == 0) {
    // first occurrence of the maximum: the sorted head has the largest index among equal maxima, walk to the smallest
    const uint32_t top = static_cast<uint32_t>(keys[0] >> 32);
    int best = static_cast<int>(keys[0] & 0xffffffffu);
    for (int i = 1; i < cols && static_cast<uint32_t>(keys[i] >> 32) == top; ++i) best = static_cast<int>(keys[i] & 0xffffffffu);
    argmax[row] = best;
  }
}

}  // namespace
}  // namespace jimm

using namespace jimm;

extern "C" int jimm_postprocess(const float* logits, int rows, int cols, int ld, int mode, float* probs, int ldp, int32_t* order, int32_t* argmax,
                                void* stream) {
  if (rows < 0 || cols <= 0 || ld < cols || (probs && ldp < cols)) { set_last_error("bad shape rows=%d cols=%d ld=%d ldp=%d", rows, cols, ld, ldp); return JIMM_EINVAL; }
  if (mode != 0 && mode != 1) { set_last_error("mode must be 0 (softmax) or 1 (sigmoid), got %d", mode); return JIMM_EINVAL; }
  if (!logits) { set_last_error("null logits"); return JIMM_EINVAL; }
  if (rows == 0) return 0;
  int npow2 = 1;
  if (order || argmax) {
    if (cols > kMaxCols) { set_last_error("ordering supports up to %d columns, got %d", kMaxCols, cols); return JIMM_EINVAL; }
    while (npow2 < cols) npow2 <<= 1;
  }
  const size_t smem = static_cast<size_t>(npow2) * sizeof(unsigned long long);
  JIMM_CUDA_CHECK(launch_k(postprocess_kernel, dim3(rows), dim3(kThreads), smem, static_cast<cudaStream_t>(stream), 1, false, logits, cols, ld, mode,
                           probs, ldp, order, argmax, npow2));
  note_launch();
  return 0;
}
