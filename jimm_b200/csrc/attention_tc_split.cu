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
//   warp 0   TMEM allocator, then TMA: Q, K, V of the item (three `rows` x 128 B boxes of the fused qkv buffer, SWIZZLE_128B), ring of 2-4 items
//   warp 1   MMA issuer, software pipelined over units u:  Q K^T(u) -> P V(u-1) -> Q K^T(u+1) -> P V(u) ...
//            S_u = Q_t K^T (tcgen05.mma SS, M=128, N=ceil16(S), K=64); O_u = P_u V (A = P_u read from TENSOR MEMORY, B = V MN-major)
//   warps 2-17  softmax + output: 8 warps per TMEM slot, TWO threads per query row (column halves A / B).  Round 2 ran one thread per row
//            (2 softmax warps per SM sub-partition): neither the MUFU (37 % busy) nor the issue slots (31 %) were the limit, the per-unit
//            dependency chain was (profiles/r2_b_attention.md).  Splitting a row's score chunks between two warps halves the softmax leg
//            of that chain and doubles the warps the schedulers can pick from.
//
// Tensor memory of slot g (columns [256g, 256g+256)), n = ceil(Nk/32) score chunks, nA = ceil(n/2) of them owned by half A:
//   S  [0, Nk)                          written by Q K^T
//   P  chunk c (16 columns, 16-bit pairs) over the owner's OWN consumed score columns: A: [16c, 16c+16), B: [32nA + 16(c-nA), +16)
//      -- no cross-warp hazard; the P V MMAs take one tensor-memory address per 16-key step, so P need not be contiguous
//   O  [192, 256) when n <= 6, else [64, 128) (A's score chunks 2-3, consumed long before the P V MMAs start)
//   (m_ref, l) of each half: two columns right behind its P region (exchange at the end of the softmax)
// The two halves of a row need ONE reference maximum: they swap the maxima of their first chunks through shared memory (bf16 -- the
// reference only has to be common and within 2^8 of the true maximum), raise it lazily as before, and reconcile the rare divergence at the
// end when they swap (m_ref, l) through tensor memory.
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"

namespace jimm {

static constexpr int ATCS_THREADS = 576;                            // 18 warps: producer, issuer, 16 softmax
static constexpr int ATCS_SM_WARPS = 16;
static constexpr int ATCS_OBUF_BYTES = ATCS_SM_WARPS * 32 * 64;      // one 32 x 64 B store box per softmax warp
static constexpr int ATCS_XCH_BYTES = 2 * 4 * 2 * 32 * 2;           // first-chunk maxima: [slot][lane quarter][half][lane] bf16
static constexpr int ATCS_MAX_BUFS = 4;                             // item ring depth (Q, K, V boxes of `rows` rows each)
static constexpr int ATCS_SMEM_BUDGET = 232448 - ATCS_OBUF_BYTES - ATCS_XCH_BYTES - 256 - 1024;  // bytes left for the item ring

// multiply a packed pair of 16-bit values by f (rare lazy-rescale path)
template <typename T>
__device__ __forceinline__ uint32_t scale_pair_s(uint32_t v, float f) {
  if constexpr (std::is_same<T, __half>::value) {
    const __half2 h = __hmul2(*reinterpret_cast<const __half2*>(&v), __float2half2_rn(f));
    return *reinterpret_cast<const uint32_t*>(&h);
  } else {
    const __nv_bfloat162 h = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&v), __float2bfloat162_rn(f));
    return *reinterpret_cast<const uint32_t*>(&h);
  }
}

struct AtcsParams {
  int B, S, H, D;
  int nq;       // query tiles per item (1 or 2)
  int Nk;       // keys rounded up to 16 (MMA N of S = Q K^T, MMA K of O = P V)
  int rows;     // rows of one Q / K / V box (>= S, multiple of 8): the item buffer is 3 * rows * 128 bytes
  int nbuf;     // item buffers in the smem ring (2 .. ATCS_MAX_BUFS)
  float scale_log2;
  void* out;
  int reverse;  // walk the (sample, head) items from the end (see kernels.cuh)
  int debug;    // JIMM_ATC_DEBUG bring-up probes (wrong results): 1 = no ex2, 2 = only the first score chunk is loaded, 4 = no O read-out / store
};

// column of P chunk c inside a slot (see the layout above)
__device__ __forceinline__ int atc_pcol(int c, int nA) { return c < nA ? 16 * c : 32 * nA + 16 * (c - nA); }

template <typename T, typename OutT, bool CAUSAL>
__global__ void __launch_bounds__(ATCS_THREADS, 1)
attention_tc_split_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_out, const AtcsParams p) {
  constexpr uint32_t FMT = std::is_same<T, __half>::value ? 0u : 1u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int tile_bytes = p.rows * 128;   // one Q / K / V box
  const int item_bytes = 3 * tile_bytes;  // multiple of 1024 (rows % 8 == 0)
  uint8_t* obuf_base = smem + p.nbuf * item_bytes;
  __nv_bfloat16* xch = reinterpret_cast<__nv_bfloat16*>(obuf_base + ATCS_OBUF_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(obuf_base + ATCS_OBUF_BYTES + ATCS_XCH_BYTES);
  uint64_t* kv_full = bars;                      // [ATCS_MAX_BUFS]
  uint64_t* kv_empty = bars + ATCS_MAX_BUFS;      // [ATCS_MAX_BUFS]
  uint64_t* s_full = bars + 2 * ATCS_MAX_BUFS;    // [2] per TMEM slot
  uint64_t* p_ready = s_full + 2;                // [2]
  uint64_t* o_full = s_full + 4;                 // [2]
  uint64_t* slot_free = s_full + 6;              // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(s_full + 8);

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_items = p.B * p.H;
  const int my_items = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int my_units = my_items * p.nq;
  const int n_chunks = (p.Nk + 31) / 32;
  const int nA = (n_chunks + 1) / 2;              // score chunks of half A; half B owns [nA, n_chunks)
  const int o_col = n_chunks <= 6 ? 192 : 64;

  pdl_launch_dependents();
  if (warp_idx == 1 && lane == 0) {
    tma_prefetch_desc(&map_qkv);
    tma_prefetch_desc(&map_out);
    for (int i = 0; i < ATCS_MAX_BUFS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 8);
      mbar_init(&o_full[i], 1);
      mbar_init(&slot_free[i], 8);
    }
    fence_barrier_init();
  }
  if (warp_idx == 0) tmem_alloc(tmem_ptr_smem, 512);
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
    __syncwarp();
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
          for (int kk = 0; kk < nkk; ++kk)  // 16 keys per step = half a P chunk = 8 columns
            umma_ts_f16(tmem_base + g * 256 + o_col, tmem_base + g * 256 + atc_pcol(kk >> 1, nA) + (kk & 1) * 8,
                        vdesc + static_cast<uint64_t>(kk * (2048 >> 4)), idesc_pv, kk > 0 ? 1u : 0u);
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
  } else {
    // ===================== softmax + output =====================
    const int sidx = warp_idx - 2;   // 0..15; four consecutive warps cover the four TMEM lane quarters
    const int q = warp_idx & 3;      // TMEM lane quarter (hardware rule: warp % 4)
    const int g = (sidx >> 2) & 1;   // softmax group == TMEM slot: units g, g+2, g+4, ...
    const int ch = sidx >> 3;        // column half: 0 = A (chunks [0, nA)), 1 = B (chunks [nA, n_chunks))
    {
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 256;
      uint8_t* obuf = obuf_base + sidx * (32 * 64);
      const int pair_bar = 1 + g * 4 + q;  // named barrier of the two warps that share these 32 rows
      __nv_bfloat16* xmine = xch + ((g * 4 + q) * 2 + ch) * 32 + lane;
      const __nv_bfloat16* xpeer = xch + ((g * 4 + q) * 2 + (ch ^ 1)) * 32 + lane;
      const int S = p.S;
      const int c0 = ch == 0 ? 0 : nA, c1 = ch == 0 ? nA : n_chunks;  // this half's chunks
      // exchange columns: right behind each half's P region (its own consumed score columns)
      const int xa_col = 16 * nA, xb_col = 32 * nA + 16 * (n_chunks - nA);
      const bool peer_has = ch == 0 ? (n_chunks > nA) : true;  // half B is empty when the keys fit one chunk
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
        const int n_live = min(c1, (kmax_warp + 31) / 32), n_full = kmin_warp / 32;  // this half: chunks [c0, n_live) carry keys
        mbar_wait(&s_full[g], sp);
        tcgen05_fence_after();
        if (t * 128 + q * 32 >= S) {
          // No query row of this warp pair exists (S = 197: rows 224..255 of the second tile; S = 50: the upper two quarters): keep the
          // mbarrier protocol, skip the work (both halves take this branch together, so the pair barriers stay matched).
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
        // ---- single pass over this half's scores.  Lazy-rescale softmax: p = exp2((s - m_ref) * scale) against a reference maximum that
        //      is only raised when a chunk's maximum exceeds it by more than 2^8 in the exp2 domain (so p <= 256, far inside fp16/bf16
        //      range); the rare raise rescales the row sum and the P chunks already written. ----
        float2 l2 = make_float2(0.f, 0.f);
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
        const float th_raw = 8.0f / p.scale_log2;
        uint32_t r[32];  // ONE score buffer: with 4-5 warps per sub-partition the other warps cover the tcgen05.ld latency, and the
                         // kernel has to live in 96 registers (5 warps on a sub-partition share its 16 K registers)
        auto ld_chunk = [&](int c, uint32_t (&buf)[32]) {
          // the MMA wrote Nk = ceil16(S) columns: when the last chunk holds only 16 of them, load 16 (the other 16 registers keep stale
          // values, all beyond kmax and masked)
          if ((p.debug & 2) && c > c0) return;
          if (c * 32 + 16 >= p.Nk) tmem_ld_32x32b_x16(taddr + c * 32, reinterpret_cast<uint32_t (&)[16]>(buf));
          else tmem_ld_32x32b_x32(taddr + c * 32, buf);
        };
        auto chunk_max = [&](const uint32_t (&sv)[32], int c) {
          float cm;
          if (c < n_full) {
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
          return cm;
        };
        // ---- common reference maximum: the two halves swap the maxima of their first chunks (bf16, so both compute the same value) ----
        float cm_first = -INFINITY;
        if (c0 < n_live) {
          ld_chunk(c0, r);
          tmem_ld_wait();
          cm_first = chunk_max(r, c0);
        }
        const __nv_bfloat16 mine16 = __float2bfloat16_rn(cm_first);
        *xmine = mine16;
        named_bar_sync(pair_bar, 64);
        float m_ref = fmaxf(__bfloat162float(mine16), __bfloat162float(*xpeer));  // finite: key 0 is valid for every row
        const float m_ref0 = m_ref;
        // One chunk = 32 scores of the row.  Short dependent chains (four interleaved max / sum chains) and the exponentials issued
        // SPECULATIVELY against the current reference maximum while the chunk maximum is still being reduced; a raise (rare) discards
        // the speculative values and redoes the chunk exactly.
        auto softmax_chunk = [&](const uint32_t (&sv)[32], int c, float cm) {
          const bool full = c < n_full;
          uint32_t pk[16];
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
              float2 e0, e1, e2, e3;
              if (p.debug & 1) {
                e0 = a0; e1 = a1; e2 = a2; e3 = a3;
              } else {
                e0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y)); e1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
                e2 = make_float2(ex2_approx(a2.x), ex2_approx(a2.y)); e3 = make_float2(ex2_approx(a3.x), ex2_approx(a3.y));
              }
              s0 = fadd2(s0, e0); s1 = fadd2(s1, e1); s2 = fadd2(s2, e2); s3 = fadd2(s3, e3);
              pk[(j >> 1)] = pack2(e0.x, e0.y, FMT == 0 ? 1 : 2);
              pk[(j >> 1) + 1] = pack2(e1.x, e1.y, FMT == 0 ? 1 : 2);
              pk[(j >> 1) + 2] = pack2(e2.x, e2.y, FMT == 0 ? 1 : 2);
              pk[(j >> 1) + 3] = pack2(e3.x, e3.y, FMT == 0 ? 1 : 2);
            }
            sum = fadd2(fadd2(s0, s1), fadd2(s2, s3));
          };
          const uint32_t pdst = taddr + atc_pcol(c, nA);
          if (full) {
            float2 sum;
            exp_full(m_ref, sum);
            if (!__any_sync(0xffffffffu, cm > m_ref + th_raw)) {
              l2 = fadd2(l2, sum);
              tmem_st_32x32b_x16(pdst, pk);
              return;
            }
          }
          const bool raise = cm > m_ref + th_raw;
          if (__any_sync(0xffffffffu, raise)) {
            const float new_ref = raise ? cm : m_ref;
            const float f = raise ? ex2_approx((m_ref - new_ref) * p.scale_log2) : 1.0f;
            l2.x *= f;
            l2.y *= f;
            m_ref = new_ref;
            if (c > c0) {  // rescale the P chunks already stored (rare)
              tmem_st_wait();
              for (int j = c0; j < c; ++j) {
                uint32_t pp[16];
                tmem_ld_32x32b_x16(taddr + atc_pcol(j, nA), pp);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) pp[e] = scale_pair_s<T>(pp[e], f);
                tmem_st_32x32b_x16(taddr + atc_pcol(j, nA), pp);
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
          tmem_st_32x32b_x16(pdst, pk);
        };
        // P chunk c lands inside this half's score chunk c0 + (c - c0) / 2 <= c, which is already in registers; the chunk in flight
        // (c+1) starts at column 32(c+1) >= the end of P chunk c.
        for (int c = c0; c < n_live; ++c) {
          if (c > c0) tmem_ld_wait();
          softmax_chunk(r, c, c == c0 ? cm_first : chunk_max(r, c));
          if (c + 1 < n_live) ld_chunk(c + 1, r);
        }
        for (int c = max(c0, n_live); c < c1; ++c) {  // keys masked for the whole warp (causal): P = 0
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
          tmem_st_32x32b_x16(taddr + atc_pcol(c, nA), pk);
        }
        float l = l2.x + l2.y;
        // ---- the halves swap (m_ref, l) through tensor memory and agree on the row's reference and sum ----
        if (c1 > c0) {
          tmem_st_32x32b_x2(taddr + (ch == 0 ? xa_col : xb_col), __float_as_uint(m_ref), __float_as_uint(l));
        }
        tmem_st_wait();
        named_bar_sync(pair_bar, 64);
        float pm = -INFINITY, pl = 0.f;
        if (peer_has) {
          uint32_t u0, u1;
          tmem_ld_32x32b_x2(taddr + (ch == 0 ? xb_col : xa_col), u0, u1);
          tmem_ld_wait();
          pm = __uint_as_float(u0);
          pl = __uint_as_float(u1);
        }
        {
          const float m = fmaxf(m_ref, pm);
          const bool lower = m_ref < m;  // the peer raised its reference further than this half did (rare): bring this half's P along
          if (__any_sync(0xffffffffu, m_ref != m_ref0 || pm != m_ref0)) {  // (usually both halves are still on the common starting reference)
            const float f = lower ? ex2_approx((m_ref - m) * p.scale_log2) : 1.0f;
            if (__any_sync(0xffffffffu, lower)) {
              for (int j = c0; j < c1; ++j) {
                uint32_t pp[16];
                tmem_ld_32x32b_x16(taddr + atc_pcol(j, nA), pp);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) pp[e] = scale_pair_s<T>(pp[e], f);
                tmem_st_32x32b_x16(taddr + atc_pcol(j, nA), pp);
              }
              tmem_st_wait();
            }
            l = l * f + pl * ex2_approx((pm - m) * p.scale_log2);
          } else {
            l += pl;
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[g]);
        // ---- output: this half's 32 columns of O / l ----
        const float inv = 1.0f / l;
        mbar_wait(&o_full[g], sp);
        tcgen05_fence_after();
        if (!(p.debug & 4)) {
          tmem_ld_32x32b_x32(taddr + o_col + ch * 32, r);
          tmem_ld_wait();
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_free[g]);  // TMEM of this slot may be overwritten by its next unit's S
        // O rows -> 64B-swizzled 32 x 64 B box in smem -> 3-D TMA store (rows >= S are clipped by the [B, S, D] tensor map)
        if (!(p.debug & 4)) {
          constexpr int NBOX = sizeof(OutT) == 2 ? 1 : 2;
#pragma unroll
          for (int bx = 0; bx < NBOX; ++bx) {
            uint32_t pk[16];
            if constexpr (sizeof(OutT) == 2) {
              constexpr int ot = std::is_same<OutT, __half>::value ? 1 : 2;
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = pack2(__uint_as_float(r[2 * j]) * inv, __uint_as_float(r[2 * j + 1]) * inv, ot);
            } else {
              constexpr bool RT = std::is_same<OutT, tf32_t>::value;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float v = __uint_as_float(r[bx * 16 + j]) * inv;
                pk[j] = __float_as_uint(RT ? round_tf32(v) : v);
              }
            }
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)  // SWIZZLE_64B: 16-byte chunk j of row `lane` sits at chunk j ^ ((lane >> 1) & 3)
              *reinterpret_cast<uint4*>(obuf + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_3d(&map_out, smem_u32(obuf), h * 64 + ch * 32 + bx * 16, t * 128 + q * 32, b);
              tma_store_commit();
            }
          }
        }
      }
    }
  }

  if (warp_idx >= 2 && lane == 0) tma_store_wait_all();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp_idx == 0) tmem_dealloc(tmem_base, 512);
}

// ---- host ----------------------------------------------------------------------------------------------------------
int make_tensor_map_2d(CUtensorMap* map, int dtype, const void* ptr, int rows, int cols, int ld, int box_rows);  // gemm.cu
int make_tensor_map_3d_box64(CUtensorMap* map, int dtype, const void* ptr, int B, int S, int N, int ld);        // gemm.cu

template <typename T, typename OutT>
static int atcs_launch(const void* qkv, int io_type, void* out, int out_type, int B, int S, int H, int causal, cudaStream_t stream, int reverse) {
  const int D = H * 64;
  AtcsParams p;
  p.B = B; p.S = S; p.H = H; p.D = D;
  p.nq = (S + 127) / 128;
  p.Nk = ((S + 15) / 16) * 16;
  // Box rows: the keys the MMAs read (Nk) -- for two query tiles the second Q tile reads rows 128..255 of the Q box, rows past the box
  // land in the K box of the same item buffer (finite or not, they only feed query rows >= S, which the output map clips).
  p.rows = p.Nk;
  const int item_bytes = 3 * p.rows * 128;
  p.nbuf = ATCS_SMEM_BUDGET / item_bytes;
  if (p.nbuf > ATCS_MAX_BUFS) p.nbuf = ATCS_MAX_BUFS;
  if (p.nbuf < 2) { set_last_error("attention_tc: item of %d bytes does not fit a 2-deep ring", item_bytes); return -1; }
  const int smem_bytes = p.nbuf * item_bytes + ATCS_OBUF_BYTES + ATCS_XCH_BYTES + 256 + 1024;
  CUtensorMap map;
  if (int rc = make_tensor_map_2d(&map, io_type, qkv, B * S, 3 * D, 3 * D, p.rows)) return rc;
  CUtensorMap map_out;
  if (int rc = make_tensor_map_3d_box64(&map_out, out_type, out, B, S, D, D)) return rc;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.out = out;
  p.reverse = reverse;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* env = getenv("JIMM_ATC_DEBUG"); dbg = env ? atoi(env) : 0; }
    p.debug = dbg;
  }
  const int items = B * H;
  const int grid = items < device_sm_count() ? items : device_sm_count();
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    JIMM_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_split_kernel<T, OutT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    JIMM_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_split_kernel<T, OutT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
  }
  if (causal) JIMM_CUDA_CHECK(launch_k(attention_tc_split_kernel<T, OutT, true>, dim3(grid), dim3(ATCS_THREADS), smem_bytes, stream, 1, true, map, map_out, p));
  else JIMM_CUDA_CHECK(launch_k(attention_tc_split_kernel<T, OutT, false>, dim3(grid), dim3(ATCS_THREADS), smem_bytes, stream, 1, true, map, map_out, p));
  note_launch();
  return 0;
}

// Returns 1 when this configuration is not handled here (caller falls back to attention.cu / attention_tc_long.cu).
int attention_tc_split_run(const void* qkv, int io_type, void* out, int out_type, int B, int S, int H, int causal, cudaStream_t stream, int reverse) {
  if (S > 256 || S < 1) return 1;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return 1;
  if (io_type == DT_F16 && out_type == DT_F16) return atcs_launch<__half, __half>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_F16 && out_type == DT_F32) return atcs_launch<__half, float>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_F16 && out_type == DT_TF32) return atcs_launch<__half, tf32_t>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_BF16 && out_type == DT_BF16) return atcs_launch<__nv_bfloat16, __nv_bfloat16>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  if (io_type == DT_BF16 && out_type == DT_F32) return atcs_launch<__nv_bfloat16, float>(qkv, io_type, out, out_type, B, S, H, causal, stream, reverse);
  return 1;
}

}  // namespace jimm
