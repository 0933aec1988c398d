// tcgen05 softmax attention for sequences of up to 256 tokens, head_dim 64.
//
//   o[b*S+s, h*64+d] = softmax_k((q/8) k^T  masked) v            nnx.MultiHeadAttention core, common/transformer.py:130
//
// Every jimm tower with S <= 256 lands here (ViT-B/16@224: 197, SigLIP-B/16@256: 256, CLIP-B/32: 50 / 77 causal, SigLIP text: 64);
// longer sequences use attention_tc_long.cu.  With S <= 256 a whole score row fits one TMEM accumulator, so there is no
// online-softmax rescaling.  One persistent CTA per SM loops over (sample, head) ITEMS; the work UNIT is one 128-row query tile of an
// item (one or two per item).  Units alternate between two TMEM slots / softmax warp groups and run as two DECOUPLED streams: while
// one group is in its softmax (MUFU / issue bound), the other slot is in its tensor phase (P V, O read-out, next Q K^T), so the
// stage latencies of a unit overlap with the other stream instead of adding up (round 1 ran both tiles of an item in lock-step:
// 110 us per ViT-B/16 layer; the stages of an item were serial, profiles/r1_d).
//
//   warp 0   TMA: Q, K, V of the item (three `rows` x 128 B boxes of the fused qkv buffer, SWIZZLE_128B), ring of 2-4 items
//   warp 1   MMA issuer, software pipelined over units u:  Q K^T(u) -> P V(u-1) -> Q K^T(u+1) -> P V(u) ...
//            S_u = Q_t K^T (tcgen05.mma SS, M=128, N=ceil16(S), K=64); O_u = P_u V (A = P_u read from TENSOR MEMORY, B = V MN-major)
//   warp 2   TMEM allocator (512 columns: slot g = u & 1 owns columns [256g, 256g+256): S, overwritten in place by fp16/bf16 P in
//            the first N/2 columns, O in columns 128..191)
//   warps 4-11  softmax + output, group g = 4 warps per slot, thread = query row: one pass exp2 / row sum / pack / tcgen05.st P with a
//            lazily raised reference maximum, then O * (1/l) -> smem -> 3-D TMA store after the P V MMA.
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"

namespace jimm {

static constexpr int ATC_THREADS = 384;
static constexpr int ATC_OBUF_BYTES = 8 * 32 * 128;                // one 32 x 128 B store box per softmax warp
static constexpr int ATC_MAX_BUFS = 4;                             // item ring depth (Q, K, V boxes of `rows` rows each)
static constexpr int ATC_SMEM_BUDGET = 232448 - ATC_OBUF_BYTES - 256 - 1024;  // bytes left for the item ring

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
  int rows;     // rows of one Q / K / V box (>= S, multiple of 8): the item buffer is 3 * rows * 128 bytes
  int nbuf;     // item buffers in the smem ring (2 .. ATC_MAX_BUFS)
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
  const int tile_bytes = p.rows * 128;   // one Q / K / V box
  const int item_bytes = 3 * tile_bytes;  // multiple of 1024 (rows % 8 == 0)
  uint8_t* obuf_base = smem + p.nbuf * item_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(obuf_base + ATC_OBUF_BYTES);
  uint64_t* kv_full = bars;                      // [ATC_MAX_BUFS]
  uint64_t* kv_empty = bars + ATC_MAX_BUFS;      // [ATC_MAX_BUFS]
  uint64_t* s_full = bars + 2 * ATC_MAX_BUFS;    // [2] per TMEM slot
  uint64_t* p_ready = s_full + 2;                // [2]
  uint64_t* o_full = s_full + 4;                 // [2]
  uint64_t* slot_free = s_full + 6;              // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_full + 8);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.B * p.H;
  const int my_items = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int my_units = my_items * p.nq;

  pdl_launch_dependents();
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_out);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < ATC_MAX_BUFS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
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
      int buf = 0;
      uint32_t ph = 0;
      for (int it = 0; it < my_items; ++it) {
        const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
        const int ie = p.reverse ? num_items - 1 - item : item;
        const int b = ie / p.H, h = ie - b * p.H;
        uint8_t* base = smem + buf * item_bytes;
        mbar_wait(&kv_empty[buf], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[buf], item_bytes);
        const int row0 = b * p.S;
        tma_load_2d(base, &map_qkv, &kv_full[buf], h * 64, row0);
        tma_load_2d(base + tile_bytes, &map_qkv, &kv_full[buf], p.D + h * 64, row0);
        tma_load_2d(base + 2 * tile_bytes, &map_qkv, &kv_full[buf], 2 * p.D + h * 64, row0);
        if (++buf == p.nbuf) { buf = 0; ph ^= 1; }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    {
      // The whole warp runs this control flow (waits, counters and descriptors stay warp-uniform, i.e. in uniform registers); only the
      // elected lane issues the tcgen05 instructions.
      const bool leader = lane == 0;
      const uint32_t idesc_qk = make_idesc(FMT, 128, static_cast<uint32_t>(p.Nk), 0);
      const uint32_t idesc_pv = make_idesc(FMT, 128, 64, 1);  // B = V is MN-major (keys are the strided dimension)
      const int nkk = p.Nk / 16;
      // P V of unit v (slot v & 1, k-th use of that slot k = v >> 1), reading V from item buffer vbuf; `last` = last unit of its item
      auto issue_pv = [&](int v, int vbuf, bool last) {
        const int g = v & 1;
        mbar_wait(&p_ready[g], static_cast<uint32_t>(v >> 1) & 1u);
        tcgen05_fence_after();
        if (leader) {
          const uint64_t vdesc = make_umma_desc_sw128(smem_u32(smem + vbuf * item_bytes + 2 * tile_bytes));
          for (int kk = 0; kk < nkk; ++kk)
            umma_ts_f16(tmem_base + g * 256 + 128, tmem_base + g * 256 + kk * 8, vdesc + static_cast<uint64_t>(kk * (2048 >> 4)), idesc_pv,
                        kk > 0 ? 1u : 0u);
          tcgen05_commit(&o_full[g]);
          if (last) tcgen05_commit(&kv_empty[vbuf]);  // every MMA reading this item's smem has retired
        }
      };
      int buf = 0, un = 0, prev_buf = 0;
      uint32_t ph = 0;
      bool prev_last = false;
      for (int it = 0; it < my_items; ++it) {
        const uint32_t q_addr = smem_u32(smem + buf * item_bytes);
        // descriptors advance by (bytes >> 4) in their address field: 32 B per 16-element K step, 2048 B per 16 keys of V
        const uint64_t qdesc = make_umma_desc_sw128(q_addr), kdesc = make_umma_desc_sw128(q_addr + tile_bytes);
        mbar_wait(&kv_full[buf], ph);
        tcgen05_fence_after();
        for (int t = 0; t < p.nq; ++t, ++un) {
          const int g = un & 1;
          mbar_wait(&slot_free[g], (static_cast<uint32_t>(un >> 1) & 1u) ^ 1u);  // the previous unit of this slot has been read out
          tcgen05_fence_after();
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss<0>(tmem_base + g * 256, qdesc + static_cast<uint64_t>(t * (16384 >> 4) + k * 2), kdesc + static_cast<uint64_t>(k * 2), idesc_qk,
                         k > 0 ? 1u : 0u);
            tcgen05_commit(&s_full[g]);
          }
          if (un > 0) issue_pv(un - 1, prev_buf, prev_last);  // the other stream's tensor phase, under this unit's softmax
          prev_buf = buf;
          prev_last = t == p.nq - 1;
        }
        if (++buf == p.nbuf) { buf = 0; ph ^= 1; }
      }
      if (un > 0) issue_pv(un - 1, prev_buf, prev_last);
    }
  } else if (warp_idx >= 4) {
    // ===================== softmax + output =====================
    const int q = warp_idx & 3;        // TMEM lane quarter
    const int g = (warp_idx - 4) >> 2; // softmax group == TMEM slot: units g, g+2, g+4, ...
    {
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 256;
      uint8_t* obuf = obuf_base + (warp_idx - 4) * (32 * 128);
      const int S = p.S;
      const int n_chunks = (p.Nk + 31) / 32;
      for (int un = g; un < my_units; un += 2) {
        const int it = p.nq == 2 ? un >> 1 : un, t = p.nq == 2 ? un & 1 : 0;
        const int item = static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x);
        const int ie = p.reverse ? num_items - 1 - item : item;
        const int b = ie / p.H, h = ie - b * p.H;
        const uint32_t sp = static_cast<uint32_t>(un >> 1) & 1u;  // k-th use of this slot
        const int row = t * 128 + q * 32 + lane;  // query index inside the sample
        int kmax = S;  // number of keys this row attends to
        if (CAUSAL) kmax = row + 1 < S ? row + 1 : S;
        // warp-uniform upper bound of keys any row of this warp needs (rows >= S are clamped: finite garbage, never stored)
        const int kmax_warp = CAUSAL ? min(S, t * 128 + q * 32 + 32) : S;
        const int kmin_warp = CAUSAL ? min(S, t * 128 + q * 32 + 1) : S;  // keys valid for EVERY lane of this warp
        const int n_live = (kmax_warp + 31) / 32, n_full = kmin_warp / 32;
        mbar_wait(&s_full[g], sp);
        tcgen05_fence_after();
        if (t * 128 + q * 32 >= S) {
          // No query row of this warp exists (S = 197: rows 224..255 of the second tile; S = 50: the upper two warps): keep the barrier
          // protocol, skip the work.  tcgen05.ld moves 64 B/clk/SM and the score tile is the bulk of it -- a dead warp's share (1/8 of
          // an item at S = 197) is pure loss; its P / O rows are garbage that the 3-D output map clips.
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_ready[g]);
          mbar_wait(&o_full[g], sp);
          tcgen05_fence_after();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&slot_free[g]);
          continue;
        }
        // ---- single pass over S (the kernel is bound by TMEM read bandwidth, ~64 B/clk/SM: reading the scores twice for an
        //      exact row max first cost 45 % more; profiles/r1_d).  Lazy-rescale softmax: p = exp2((s - m_ref) * scale) against a
        //      reference maximum that is only raised when a chunk's maximum exceeds it by more than 2^8 in the exp2 domain (so
        //      p <= 256, far inside fp16/bf16 range); the rare raise rescales the row sum and the P chunks already written. ----
        float m_ref = -INFINITY;
        float2 l2 = make_float2(0.f, 0.f);
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
        const float th_raw = 8.0f / p.scale_log2;
        uint32_t r[32], rn[32];
        // One chunk = 32 scores of the row.  The dependent chains are kept short (four interleaved max / sum chains instead of one of
        // 16) and, after the first chunk, the exponentials are issued SPECULATIVELY against the current reference maximum while the
        // chunk maximum is still being reduced: with one warp per SMSP per stream the MUFU was only ~48 % busy because every chunk
        // serialised max chain -> raise vote -> 32 ex2 -> sum chain (profiles/r2_b_attention.md).  A raise (rare) discards the
        // speculative values and redoes the chunk exactly.
        auto softmax_chunk = [&](const uint32_t (&sv)[32], int c) {
          const bool full = c < n_full;
          uint32_t pk[16];
          float cm;
          if (full) {
            float q0 = -INFINITY, q1 = -INFINITY, q2 = -INFINITY, q3 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              q0 = fmax3(q0, __uint_as_float(sv[j]), __uint_as_float(sv[j + 1]));
              q1 = fmax3(q1, __uint_as_float(sv[j + 2]), __uint_as_float(sv[j + 3]));
              q2 = fmax3(q2, __uint_as_float(sv[j + 4]), __uint_as_float(sv[j + 5]));
              q3 = fmax3(q3, __uint_as_float(sv[j + 6]), __uint_as_float(sv[j + 7]));
            }
            cm = fmaxf(fmaxf(q0, q1), fmaxf(q2, q3));
          } else {
            cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) cm = (c * 32 + j < kmax) ? fmaxf(cm, __uint_as_float(sv[j])) : cm;
          }
          // exponentials of a full chunk against reference maximum `mref`: packed P values + four partial row sums
          auto exp_full = [&](float mref, float2& sum) {
            const float2 mo2 = make_float2(-mref * p.scale_log2, -mref * p.scale_log2);
            float2 s0 = make_float2(0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const float2 a0 = ffma2(make_float2(__uint_as_float(sv[j]), __uint_as_float(sv[j + 1])), sc2, mo2);
              const float2 a1 = ffma2(make_float2(__uint_as_float(sv[j + 2]), __uint_as_float(sv[j + 3])), sc2, mo2);
              const float2 a2 = ffma2(make_float2(__uint_as_float(sv[j + 4]), __uint_as_float(sv[j + 5])), sc2, mo2);
              const float2 a3 = ffma2(make_float2(__uint_as_float(sv[j + 6]), __uint_as_float(sv[j + 7])), sc2, mo2);
              const float2 e0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y)), e1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
              const float2 e2 = make_float2(ex2_approx(a2.x), ex2_approx(a2.y)), e3 = make_float2(ex2_approx(a3.x), ex2_approx(a3.y));
              s0 = fadd2(s0, e0); s1 = fadd2(s1, e1); s2 = fadd2(s2, e2); s3 = fadd2(s3, e3);
              pk[(j >> 1)] = pack2(e0.x, e0.y, FMT == 0 ? 1 : 2);
              pk[(j >> 1) + 1] = pack2(e1.x, e1.y, FMT == 0 ? 1 : 2);
              pk[(j >> 1) + 2] = pack2(e2.x, e2.y, FMT == 0 ? 1 : 2);
              pk[(j >> 1) + 3] = pack2(e3.x, e3.y, FMT == 0 ? 1 : 2);
            }
            sum = fadd2(fadd2(s0, s1), fadd2(s2, s3));
          };
          if (full && c > 0) {  // speculative: m_ref is finite after the first chunk (key 0 is valid for every row)
            float2 sum;
            exp_full(m_ref, sum);
            if (!__any_sync(0xffffffffu, cm > m_ref + th_raw)) {
              l2 = fadd2(l2, sum);
              tmem_st_32x32b_x16(taddr + c * 16, pk);
              return;
            }
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
          if (full) {
            float2 sum;
            exp_full(m_ref, sum);
            l2 = fadd2(l2, sum);
          } else {
            const float2 mo2 = make_float2(-m_ref * p.scale_log2, -m_ref * p.scale_log2);
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
        // the MMA wrote Nk = ceil16(S) columns: when the last chunk holds only 16 of them, load 16 (its other 16 registers keep stale
        // values, all beyond kmax and masked) -- 7 % of the score bytes at S = 197
        auto ld_chunk = [&](int c, uint32_t (&buf)[32]) {
          if (c * 32 + 16 >= p.Nk) tmem_ld_32x32b_x16(taddr + c * 32, reinterpret_cast<uint32_t (&)[16]>(buf));
          else tmem_ld_32x32b_x32(taddr + c * 32, buf);
        };
        ld_chunk(0, r);
        if (n_live > 1) ld_chunk(1, rn);
        for (int c = 0; c < n_live; c += 2) {
          tmem_ld_wait();
          softmax_chunk(r, c);
          if (c + 2 < n_live) ld_chunk(c + 2, r);
          if (c + 1 < n_live) {
            softmax_chunk(rn, c + 1);
            if (c + 3 < n_live) ld_chunk(c + 3, rn);
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
        if (lane == 0) mbar_arrive(&p_ready[g]);
        // ---- output: O / l ----
        const float inv = 1.0f / l;
        mbar_wait(&o_full[g], sp);
        tcgen05_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32b_x32(taddr + 128, o0);
        tmem_ld_32x32b_x32(taddr + 160, o1);
        tmem_ld_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_free[g]);  // TMEM of this slot may be overwritten by its next unit's S
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
  AtcParams p;
  p.B = B; p.S = S; p.H = H; p.D = D;
  p.nq = (S + 127) / 128;
  p.Nk = ((S + 15) / 16) * 16;
  // Box rows: the keys the MMAs read (Nk) -- for two query tiles the second Q tile reads rows 128..255 of the Q box, rows past the box
  // land in the K box of the same item buffer (finite or not, they only feed query rows >= S, which the output map clips).
  p.rows = p.Nk;
  const int item_bytes = 3 * p.rows * 128;
  p.nbuf = ATC_SMEM_BUDGET / item_bytes;
  if (p.nbuf > ATC_MAX_BUFS) p.nbuf = ATC_MAX_BUFS;
  if (p.nbuf < 2) { set_last_error("attention_tc: item of %d bytes does not fit a 2-deep ring", item_bytes); return -1; }
  const int smem_bytes = p.nbuf * item_bytes + ATC_OBUF_BYTES + 256 + 1024;
  CUtensorMap map;
  if (int rc = make_tensor_map_2d(&map, io_type, qkv, B * S, 3 * D, 3 * D, p.rows)) return rc;
  CUtensorMap map_out;
  if (int rc = make_tensor_map_3d(&map_out, out_type, out, B, S, D, D)) return rc;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.out = out;
  p.reverse = reverse;
  const int items = B * H;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    JIMM_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel<T, OutT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    JIMM_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_kernel<T, OutT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  }
  if (causal) JIMM_CUDA_CHECK(launch_k(attention_tc_kernel<T, OutT, true>, dim3(grid), dim3(ATC_THREADS), smem_bytes, stream, 1, true, map, map_out, p));
  else JIMM_CUDA_CHECK(launch_k(attention_tc_kernel<T, OutT, false>, dim3(grid), dim3(ATC_THREADS), smem_bytes, stream, 1, true, map, map_out, p));
  note_launch();
  return 0;
}

// Returns 1 when this configuration is not handled here (caller falls back to attention.cu / attention_tc_long.cu).
int attention_tc_run(const void* qkv, int io_type, void* out, int out_type, int B, int S, int H, int causal, cudaStream_t stream, int reverse) {
  if (S > 256 || S < 1) return 1;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return 1;
  if (io_type == DT_F16 && out_type == DT_F16) return atc_launch<__half, __half>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_F16 && out_type == DT_F32) return atc_launch<__half, float>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_F16 && out_type == DT_TF32) return atc_launch<__half, tf32_t>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_BF16 && out_type == DT_BF16) return atc_launch<__nv_bfloat16, __nv_bfloat16>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_BF16 && out_type == DT_F32) return atc_launch<__nv_bfloat16, float>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  return 1;
}

}  // namespace jimm
