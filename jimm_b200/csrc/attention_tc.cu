// tcgen05 softmax attention for sequences of up to 256 tokens, head_dim 64.
//
//   o[b*S+s, h*64+d] = softmax_k((q/8) k^T  masked) v            nnx.MultiHeadAttention core, common/transformer.py:130
//
// Every jimm tower with S <= 256 lands here (ViT-B/16@224: 197, SigLIP-B/16@256: 256, CLIP-B/32: 50 / 77 causal); longer
// sequences use the flash kernel in attention.cu.  With S <= 256 a whole score row fits one TMEM accumulator, so there is
// no online-softmax rescaling: one persistent CTA per SM loops over (sample, head) items:
//
//   warp 0   TMA: Q, K, V of the item (three 256 x 128 B boxes of the fused qkv buffer, SWIZZLE_128B), 2-deep ring
//   warp 1   MMA: S_t = Q_t K^T  (tcgen05.mma SS, M=128, N=ceil16(S), K=64) for the one or two 128-row query tiles t,
//                 then O_t = P_t V (tcgen05.mma with A = P_t read from TENSOR MEMORY, B = V as an MN-major smem operand)
//   warp 2   TMEM allocator (512 columns: tile t owns columns [256t, 256t+256): S, overwritten in place by fp16/bf16 P in
//            the first N/2 columns, O in columns 128..191)
//   warps 4-11  softmax + output, one group of 4 warps per query tile, thread = query row: pass 1 row max from TMEM,
//            pass 2 exp2 / row sum / pack / tcgen05.st P, then O * (1/l) -> global after the P V MMA.
// Measured bound: TMEM read bandwidth (~64 B/clk/SM), hence the single pass over S with a lazily raised reference maximum.
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"

namespace jimm {

static constexpr int ATC_THREADS = 384;
static constexpr int ATC_TILE_BYTES = 256 * 128;                 // one Q / K / V box
static constexpr int ATC_BUF_BYTES = 3 * ATC_TILE_BYTES;          // 96 KB per item
static constexpr int ATC_OBUF_BYTES = 8 * 32 * 128;                // one 32 x 128 B store box per softmax warp
static constexpr int ATC_SMEM = 2 * ATC_BUF_BYTES + ATC_OBUF_BYTES + 256 + 1024;

// multiply a packed pair of 16-bit values by f (rare lazy-rescale path)
template <typename T>
__device__ __forceinline__ uint32_t scale_pair(uint32_t v, float f) {
  if constexpr (std::is_same<T, __half>::value) {
    const __half2 h = __hmul2(*reinterpret_cast<const __half2*>(&v), __float2half2_rn(f));
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const __nv_bfloat162 h = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&v), __float2bfloat162_rn(f));
    return *reinterpret_cast<const uint32_t*>(&h);
  }
}

struct AtcParams {
  int B, S, H, D;
  int nq;       // query tiles per item (1 or 2)
  int Nk;       // keys rounded up to 16 (MMA N of S = Q K^T, MMA K of O = P V)
  float scale_log2;
  void* out;
  int reverse;  // walk the (sample, head) items from the end (see kernels.cuh)
};

template <typename T, typename OutT, bool CAUSAL>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_out, const AtcParams p) {
  constexpr uint32_t FMT = std::is_same<T, __half>::value ? 0u : 1u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* obuf_base = smem + 2 * ATC_BUF_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * ATC_BUF_BYTES + ATC_OBUF_BYTES);
  uint64_t* kv_full = bars;        // [2]
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;     // [2] per query tile
  uint64_t* p_ready = bars + 6;    // [2]
  uint64_t* o_full = bars + 8;     // [2]
  uint64_t* slot_free = bars + 10; // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.B * p.H;

  pdl_launch_dependents();
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_out);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_full[i], 1);
      mbar_init(&slot_free[i], 4);
    }
    fence_barrier_init();
  }
  if (warp_idx == 2) tmem_alloc(tmem_ptr_smem, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp_idx == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int ie = p.reverse ? num_items - 1 - item : item;
        const int b = ie / p.H, h = ie - b * p.H;
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1;
        uint8_t* base = smem + buf * ATC_BUF_BYTES;
        mbar_wait(&kv_empty[buf], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[buf], ATC_BUF_BYTES);
        const int row0 = b * p.S;
        tma_load_2d(base, &map_qkv, &kv_full[buf], h * 64, row0);
        tma_load_2d(base + ATC_TILE_BYTES, &map_qkv, &kv_full[buf], p.D + h * 64, row0);
        tma_load_2d(base + 2 * ATC_TILE_BYTES, &map_qkv, &kv_full[buf], 2 * p.D + h * 64, row0);
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    {
      // The whole warp runs this control flow (waits, counters and descriptors stay warp-uniform, i.e. in uniform registers); only the
      // elected lane issues the tcgen05 instructions.  With 16 small MMAs per item the issue cost matters (long kernel: -4 %).
      const bool leader = lane == 0;
      const uint32_t idesc_qk = make_idesc(FMT, 128, static_cast<uint32_t>(p.Nk), 0);
      const uint32_t idesc_pv = make_idesc(FMT, 128, 64, 1);  // B = V is MN-major (keys are the strided dimension)
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t ph = (it >> 1) & 1, sp = it & 1;
        const uint32_t q_addr = smem_u32(smem + buf * ATC_BUF_BYTES);
        // descriptors advance by (bytes >> 4) in their address field: 32 B per 16-element K step, 2048 B per 16 keys of V
        const uint64_t qdesc = make_umma_desc_sw128(q_addr), kdesc = make_umma_desc_sw128(q_addr + ATC_TILE_BYTES),
                       vdesc = make_umma_desc_sw128(q_addr + 2 * ATC_TILE_BYTES);
        mbar_wait(&kv_full[buf], ph);
        tcgen05_fence_after();
        for (int t = 0; t < p.nq; ++t) {
          mbar_wait(&slot_free[t], sp ^ 1);  // the previous item's O of this tile has been read out
          tcgen05_fence_after();
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss<0>(tmem_base + t * 256, qdesc + static_cast<uint64_t>(t * (16384 >> 4) + k * 2), kdesc + static_cast<uint64_t>(k * 2), idesc_qk,
                         k > 0 ? 1u : 0u);
            tcgen05_commit(&s_full[t]);
          }
        }
        for (int t = 0; t < p.nq; ++t) {
          mbar_wait(&p_ready[t], sp);
          tcgen05_fence_after();
          if (leader) {
            for (int kk = 0; kk < p.Nk / 16; ++kk)
              umma_ts_f16(tmem_base + t * 256 + 128, tmem_base + t * 256 + kk * 8, vdesc + static_cast<uint64_t>(kk * (2048 >> 4)), idesc_pv,
                          kk > 0 ? 1u : 0u);
            tcgen05_commit(&o_full[t]);
          }
        }
        if (leader) tcgen05_commit(&kv_empty[buf]);  // every MMA reading this item's smem has retired
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== softmax + output =====================
    const int q = warp_idx & 3;        // TMEM lane quarter
    const int t = (warp_idx - 4) >> 2; // query tile
    if (t < p.nq) {
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 256;
      uint8_t* obuf = obuf_base + (warp_idx - 4) * (32 * 128);
      const int row = t * 128 + q * 32 + lane;  // query index inside the sample
      const int S = p.S;
      int kmax = S;  // number of keys this row attends to
      if (CAUSAL) kmax = row + 1 < S ? row + 1 : S;
      // warp-uniform upper bound of keys any row of this warp needs (rows >= S are clamped: finite garbage, never stored)
      const int kmax_warp = CAUSAL ? min(S, t * 128 + q * 32 + 32) : S;
      const int kmin_warp = CAUSAL ? min(S, t * 128 + q * 32 + 1) : S;  // keys valid for EVERY lane of this warp
      const int n_chunks = (p.Nk + 31) / 32, n_live = (kmax_warp + 31) / 32, n_full = kmin_warp / 32;
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int ie = p.reverse ? num_items - 1 - item : item;
        const int b = ie / p.H, h = ie - b * p.H;
        const uint32_t sp = it & 1;
        mbar_wait(&s_full[t], sp);
        tcgen05_fence_after();
        // ---- single pass over S (the kernel is bound by TMEM read bandwidth, ~64 B/clk/SM: reading the scores twice for an
        //      exact row max first cost 45 % more; profiles/r1_d).  Lazy-rescale softmax: p = exp2((s - m_ref) * scale) against a
        //      reference maximum that is only raised when a chunk's maximum exceeds it by more than 2^8 in the exp2 domain (so
        //      p <= 256, far inside fp16/bf16 range); the rare raise rescales the row sum and the P chunks already written. ----
        float m_ref = -INFINITY;
        float2 l2 = make_float2(0.f, 0.f);
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
        const float th_raw = 8.0f / p.scale_log2;
        uint32_t r[32], rn[32];
        auto softmax_chunk = [&](const uint32_t (&sv)[32], int c) {
          // chunk maximum over this row's valid keys
          float cm = -INFINITY;
          if (c < n_full) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) cm = fmax3(cm, __uint_as_float(sv[j]), __uint_as_float(sv[j + 1]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) cm = (c * 32 + j < kmax) ? fmaxf(cm, __uint_as_float(sv[j])) : cm;
          }
          const bool raise = cm > m_ref + th_raw;  // always true for the first valid chunk (m_ref = -inf)
          if (__any_sync(0xffffffffu, raise)) {
            const float new_ref = raise ? cm : m_ref;
            const float f = raise ? ex2_approx((m_ref - new_ref) * p.scale_log2) : 1.0f;  // exp2(-inf) = 0 on the first chunk
            l2.x *= f;
            l2.y *= f;
            m_ref = new_ref;
            if (c > 0) {  // rescale the P chunks already stored (rare)
              tmem_st_wait();
              for (int j = 0; j < c; ++j) {
                uint32_t pp[16];
                tmem_ld_32x32b_x16(taddr + j * 16, pp);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) pp[e] = scale_pair<T>(pp[e], f);
                tmem_st_32x32b_x16(taddr + j * 16, pp);
              }
            }
          }
          const float2 mo2 = make_float2(-m_ref * p.scale_log2, -m_ref * p.scale_log2);
          uint32_t pk[16];
          if (c < n_full) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 a = ffma2(make_float2(__uint_as_float(sv[j]), __uint_as_float(sv[j + 1])), sc2, mo2);
              const float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
              l2 = fadd2(l2, e);
              pk[j >> 1] = pack2(e.x, e.y, FMT == 0 ? 1 : 2);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float2 a = ffma2(make_float2(__uint_as_float(sv[j]), __uint_as_float(sv[j + 1])), sc2, mo2);
              float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
              e.x = (c * 32 + j < kmax) ? e.x : 0.f;
              e.y = (c * 32 + j + 1 < kmax) ? e.y : 0.f;
              l2 = fadd2(l2, e);
              pk[j >> 1] = pack2(e.x, e.y, FMT == 0 ? 1 : 2);
            }
          }
          tmem_st_32x32b_x16(taddr + c * 16, pk);
        };
        // P chunk c (16 columns) lands on S columns [16c, 16c+16), inside S chunk c/2 <= c, which is already in registers; S
        // chunks c+1 .. c+3 that may be in flight start at column 32(c+1) >= 16c+16.
        // two x32 loads per tcgen05.wait::ld (the wait is a MEMBAR-class instruction: halve their number); a buffer is refilled
        // as soon as its chunk has been consumed, so the loads of chunks c+2 / c+3 fly during the math of chunks c / c+1
        tmem_ld_32x32b_x32(taddr, r);
        if (n_live > 1) tmem_ld_32x32b_x32(taddr + 32, rn);
        for (int c = 0; c < n_live; c += 2) {
          tmem_ld_wait();
          softmax_chunk(r, c);
          if (c + 2 < n_live) tmem_ld_32x32b_x32(taddr + (c + 2) * 32, r);
          if (c + 1 < n_live) {
            softmax_chunk(rn, c + 1);
            if (c + 3 < n_live) tmem_ld_32x32b_x32(taddr + (c + 3) * 32, rn);
          }
        }
        const float l = l2.x + l2.y;
        for (int c = n_live; c < n_chunks; ++c) {  // keys masked for the whole warp (causal): P = 0
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
          tmem_st_32x32b_x16(taddr + c * 16, pk);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
        // ---- output: O / l ----
        const float inv = 1.0f / l;
        mbar_wait(&o_full[t], sp);
        tcgen05_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32b_x32(taddr + 128, o0);
        tmem_ld_32x32b_x32(taddr + 160, o1);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_free[t]);  // TMEM of this tile may be overwritten by the next item's S
        // O tile rows -> swizzled 32 x 128 B box in smem -> 3-D TMA store (rows >= S are clipped by the [B, S, D] tensor map;
        // the row-per-thread 16-byte global stores this replaces cost 32 LSU wavefronts per instruction)
        {
          constexpr int NBOX = sizeof(OutT) == 2 ? 1 : 2;
#pragma unroll
          for (int bx = 0; bx < NBOX; ++bx) {
            uint32_t pk[32];
            if constexpr (sizeof(OutT) == 2) {
              constexpr int ot = std::is_same<OutT, __half>::value ? 1 : 2;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                pk[j] = pack2(__uint_as_float(o0[2 * j]) * inv, __uint_as_float(o0[2 * j + 1]) * inv, ot);
                pk[16 + j] = pack2(__uint_as_float(o1[2 * j]) * inv, __uint_as_float(o1[2 * j + 1]) * inv, ot);
              }
            } else {
              constexpr bool RT = std::is_same<OutT, tf32_t>::value;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(bx == 0 ? o0[j] : o1[j]) * inv;
                pk[j] = __float_as_uint(RT ? round_tf32(v) : v);
              }
            }
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(obuf + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&map_out, smem_u32(obuf), h * 64 + bx * 32, t * 128 + q * 32, b);
              tma_store_commit();
            }
          }
        }
      }
    }
  }

  if (warp_idx >= 4 && lane == 0) tma_store_wait_all();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp_idx == 2) tmem_dealloc(tmem_base, 512);
}

// ---- host ----------------------------------------------------------------------------------------------------------
int make_tensor_map_2d(CUtensorMap* map, int dtype, const void* ptr, int rows, int cols, int ld, int box_rows);  // gemm.cu
int make_tensor_map_3d(CUtensorMap* map, int dtype, const void* ptr, int B, int S, int N, int ld);              // gemm.cu

template <typename T, typename OutT>
static int atc_launch(const void* qkv, int io_type, void* out, int out_type, int B, int S, int H, int causal, cudaStream_t stream, int reverse) {
  const int D = H * 64;
  CUtensorMap map;
  if (int rc = make_tensor_map_2d(&map, io_type, qkv, B * S, 3 * D, 3 * D, 256)) return rc;
  CUtensorMap map_out;
  if (int rc = make_tensor_map_3d(&map_out, out_type, out, B, S, D, D)) return rc;
  AtcParams p;
  p.B = B; p.S = S; p.H = H; p.D = D;
  p.nq = (S + 127) / 128;
  p.Nk = ((S + 15) / 16) * 16;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.out = out;
  p.reverse = reverse;
  const int items = B * H;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  static bool attr_set = false;
  if (!attr_set) {
    JIMM_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel<T, OutT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
    JIMM_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel<T, OutT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM));
    attr_set = true;
  }
  if (causal) JIMM_CUDA_CHECK(launch_k(attention_tc_kernel<T, OutT, true>, dim3(grid), dim3(ATC_THREADS), ATC_SMEM, stream, 1, true, map, map_out, p));
  else JIMM_CUDA_CHECK(launch_k(attention_tc_kernel<T, OutT, false>, dim3(grid), dim3(ATC_THREADS), ATC_SMEM, stream, 1, true, map, map_out, p));
  note_launch();
  return 0;
}

// Returns 1 when this configuration is not handled here (caller falls back to the flash kernel).
int attention_tc_run(const void* qkv, int io_type, void* out, int out_type, int B, int S, int H, int causal, cudaStream_t stream, int reverse) {
  // One query tile (S <= 128) leaves half of the softmax warps idle: the flash kernel is as fast or faster there (measured:
  // S=50 39 us vs 25 us, S=77 causal 35 vs 37 us; S=197 128 vs 218 us, S=256 140 vs 215 us at B=256).
  if (S > 256 || S <= 128) return 1;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return 1;
  if (io_type == DT_F16 && out_type == DT_F16) return atc_launch<__half, __half>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_F16 && out_type == DT_F32) return atc_launch<__half, float>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_F16 && out_type == DT_TF32) return atc_launch<__half, tf32_t>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_BF16 && out_type == DT_BF16) return atc_launch<__nv_bfloat16, __nv_bfloat16>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_BF16 && out_type == DT_F32) return atc_launch<__nv_bfloat16, float>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  return 1;
}

}  // namespace jimm
