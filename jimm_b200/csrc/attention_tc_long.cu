// tcgen05 softmax attention for sequences longer than 256 tokens (non-causal), head_dim 64:
// ViT-L/16@384 (S = 576) and SigLIP2-L/16@512 (S = 1024) -- BASELINE configs 3 and 5.
//
// One pass over the keys with a lazily raised reference maximum: P = exp2((S - m_ref) * scale) is written over S in place
// (16-bit, tensor memory) and O += P V accumulates with the tensor core's own accumulate flag.  m_ref starts at the maximum of
// the row's first 32 scores and is only replaced when a later 32-score chunk exceeds it by more than 8 in the log2 domain
// (P <= 2^8 stays comfortably inside fp16), which almost never happens; when it does, the row's running sum, the P chunks already
// written for the current block and the O accumulator (a 64-column TMEM round trip, after the previous P V has completed) are
// multiplied by 2^(old - new).  O / l at the end is the exact softmax whatever the reference was.  (The earlier two-pass version
// recomputed Q K^T for the row maxima first: 1.5x the tensor work and twice the TMEM reads; 527 us at B=128, S=576, H=16.)
//
// One persistent CTA per SM; work unit = (sample, head, pair of 128-row query tiles); keys in blocks of 192:
//   warp 0   TMA: Q pair (256 x 128 B box, 2-deep), K / V blocks (192 x 128 B boxes) through a 4-stage ring: K_0, V_0, K_1, V_1, ...
//   warp 1   MMA issuer (tcgen05.mma: S_t = Q_t K_j^T, SS, N = 192; O_t += P_t V_j, A from TMEM, B MN-major, N = 64)
//   warp 2   TMEM allocator: query tile t owns columns [256t, 256t+256): S [0,192) / P [0,48) + [96,144) / O [192,256)
//   warps 4-19  softmax + output: 8 warps per query tile = 4 lane quarters x 2 halves of the 192 score columns, so every SM
//            sub-partition holds four softmax warps (the phase is MUFU-bound; two warps per sub-partition left it latency-bound:
//            332 us at B=32 S=1024 H=16).  The two warps of a row agree on the reference maximum once per key block through shared
//            memory (each first takes the maximum of its 96 columns -- a second, cheap read of tensor memory), so the raise
//            decision is taken before any P of the block is written and only the running sums and the O accumulator need rescaling.
#include <type_traits>

#include "common.cuh"
#include "kernels.cuh"

namespace jimm {

static constexpr int ATL_THREADS = 640;
static constexpr int ATL_KB = 192;                          // keys per block
static constexpr int ATL_Q_BYTES = 256 * 128;               // Q pair box
static constexpr int ATL_KV_BYTES = ATL_KB * 128;           // K or V block box
static constexpr int ATL_NST = 4;                           // K/V ring stages
static constexpr int ATL_XCH_BYTES = 2 * 128 * 2 * 4;  // [tile][row][column half] floats exchanged between the two warps of a row
static constexpr int ATL_SMEM = 2 * ATL_Q_BYTES + ATL_NST * ATL_KV_BYTES + 512 + ATL_XCH_BYTES + 1024;

struct AtlParams {
  int B, S, H, D;
  int n_qp;    // query pairs per (sample, head)
  int n_blk;   // key blocks
  float scale_log2;
  void* out;
  int reverse;
};

template <typename T, typename OutT>
__global__ void __launch_bounds__(ATL_THREADS, 1)
attention_tc_long_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, const AtlParams p) {
  constexpr uint32_t FMT = std::is_same<T, __half>::value ? 0u : 1u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_kv = smem + 2 * ATL_Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * ATL_Q_BYTES + ATL_NST * ATL_KV_BYTES);
  uint64_t* kv_full = bars;              // [NST]
  uint64_t* kv_empty = bars + ATL_NST;   // [NST]
  uint64_t* q_full = bars + 2 * ATL_NST; // [2]
  uint64_t* q_empty = q_full + 2;        // [2]
  uint64_t* s_full = q_empty + 2;        // [2] per query tile
  uint64_t* pv_done = s_full + 2;        // [2] every P V block (only waited on by the rare O rescale)
  uint64_t* p_ready = pv_done + 2;       // [2]
  uint64_t* o_full = p_ready + 2;        // [2]
  uint64_t* o_free = o_full + 2;         // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_free + 2);
  float* xch = reinterpret_cast<float*>(smem + 2 * ATL_Q_BYTES + ATL_NST * ATL_KV_BYTES + 512);  // [2][128][2]

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_units = p.B * p.H * p.n_qp;
  const int S = p.S;

  pdl_launch_dependents();
  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_kv);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int i = 0; i < ATL_NST; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&pv_done[i], 1);
      mbar_init(&p_ready[i], 8);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_free[i], 8);
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
      int st = 0;
      uint32_t st_ph = 0;
      int ui = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++ui) {
        const int ue = p.reverse ? num_units - 1 - unit : unit;
        const int qp = ue % p.n_qp, bh = ue / p.n_qp;
        const int b = bh / p.H, h = bh - b * p.H;
        const int row0 = b * S;
        const int qb = ui & 1;
        mbar_wait(&q_empty[qb], ((ui >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[qb], ATL_Q_BYTES);
        tma_load_2d(smem_q + qb * ATL_Q_BYTES, &map_q, &q_full[qb], h * 64, row0 + qp * 256);
        for (int j = 0; j < p.n_blk; ++j) {
          for (int kv = 0; kv < 2; ++kv) {  // K_j then V_j
            mbar_wait(&kv_empty[st], st_ph ^ 1);
            mbar_arrive_expect_tx(&kv_full[st], ATL_KV_BYTES);
            tma_load_2d(smem_kv + st * ATL_KV_BYTES, &map_kv, &kv_full[st], (kv + 1) * p.D + h * 64, row0 + j * ATL_KB);
            if (++st == ATL_NST) { st = 0; st_ph ^= 1; }
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    {
      // The whole warp runs this control flow (waits, counters, addresses stay warp-uniform: the compiler can keep the MMA operands in
      // uniform registers); only the elected lane issues the tcgen05 instructions.
      const bool leader = lane == 0;
      const uint32_t idesc_qk = make_idesc(FMT, 128, ATL_KB, 0);
      const uint32_t idesc_pv = make_idesc(FMT, 128, 64, 1);
      int st = 0;
      uint32_t st_ph = 0;
      uint32_t n_pready[2] = {0, 0}, n_used[2] = {0, 0};
      int ui = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++ui) {
        const int ue = p.reverse ? num_units - 1 - unit : unit;
        const int qp = ue % p.n_qp;
        const int nq = (S - qp * 256 > 128) ? 2 : 1;
        const int qb = ui & 1;
        const uint32_t q_addr = smem_u32(smem_q + qb * ATL_Q_BYTES);
        mbar_wait(&q_full[qb], (ui >> 1) & 1);
        tcgen05_fence_after();
        const uint64_t qdesc = make_umma_desc_sw128(q_addr);  // descriptors advance by (bytes >> 4) in their address field
        auto issue_qk = [&](int t, uint32_t k_addr) {
          // S_t = Q_t K^T.  The S / P columns of tile t are free: its previous P V was issued after the softmax warps had finished with
          // them, and the tensor pipe executes in issue order.
          if (leader) {
            const uint64_t kdesc = make_umma_desc_sw128(k_addr);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss<0>(tmem_base + t * 256, qdesc + static_cast<uint64_t>(t * (16384 >> 4) + k * 2), kdesc + static_cast<uint64_t>(k * 2), idesc_qk,
                         k > 0 ? 1u : 0u);
            tcgen05_commit(&s_full[t]);
          }
        };
        auto next_stage = [&]() { if (++st == ATL_NST) { st = 0; st_ph ^= 1; } };
        // prologue: scores of block 0 for both tiles
        mbar_wait(&kv_full[st], st_ph);
        tcgen05_fence_after();
        for (int t = 0; t < nq; ++t) issue_qk(t, smem_u32(smem_kv + st * ATL_KV_BYTES));
        if (leader) tcgen05_commit(&kv_empty[st]);
        next_stage();
        for (int j = 0; j < p.n_blk; ++j) {
          // Per tile: O_t += P_t V_j as soon as ITS softmax is done, immediately followed by ITS next scores S_t = Q_t K_{j+1}^T -- the
          // other tile's softmax overlaps these MMAs (issuing both tiles' scores together, as before, made the tiles run in lock-step:
          // every softmax warp then waited for the whole MMA phase, 30 % of all stall samples).
          const int sv = st;
          const uint32_t phv = st_ph;  // V_j
          next_stage();
          const bool more = j + 1 < p.n_blk;
          const int sk = st;
          const uint32_t phk = st_ph;  // K_{j+1} (only when `more`)
          if (more) next_stage();
          const uint32_t v_addr = smem_u32(smem_kv + sv * ATL_KV_BYTES), k_addr = smem_u32(smem_kv + sk * ATL_KV_BYTES);
          for (int t = 0; t < nq; ++t) {
            mbar_wait(&p_ready[t], n_pready[t] & 1);
            ++n_pready[t];
            if (j == 0 && n_used[t] > 0) mbar_wait(&o_free[t], (n_used[t] - 1) & 1);  // previous unit's O of this tile was read out
            if (t == 0) mbar_wait(&kv_full[sv], phv);
            tcgen05_fence_after();
            if (leader) {
              const uint64_t vdesc = make_umma_desc_sw128(v_addr);
#pragma unroll
              for (int kk = 0; kk < ATL_KB / 16; ++kk)
                umma_ts_f16(tmem_base + t * 256 + 192, tmem_base + t * 256 + (kk < ATL_KB / 32 ? kk * 8 : 96 + (kk - ATL_KB / 32) * 8),
                            vdesc + static_cast<uint64_t>(kk * (2048 >> 4)), idesc_pv, (j | kk) != 0 ? 1u : 0u);  // P: keys 0-95 at columns 0-47, 96-191 at 96-143
            }
            if (leader) tcgen05_commit(&pv_done[t]);
            if (!more && leader) tcgen05_commit(&o_full[t]);
            if (more) {
              if (t == 0) { mbar_wait(&kv_full[sk], phk); tcgen05_fence_after(); }
              issue_qk(t, k_addr);
            }
          }
          if (leader) tcgen05_commit(&kv_empty[sv]);
          if (more && leader) tcgen05_commit(&kv_empty[sk]);
        }
        if (leader) tcgen05_commit(&q_empty[qb]);
        for (int t = 0; t < nq; ++t) ++n_used[t];
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== softmax + output =====================
    const int q = warp_idx & 3;              // TMEM lane quarter (hardware rule: warp % 4)
    const int t = ((warp_idx - 4) >> 2) & 1; // query tile
    const int hf = (warp_idx - 4) >> 3;      // column half: score chunks [3 hf, 3 hf + 3), O columns [32 hf, 32 hf + 32)
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + t * 256;
    float* my_x = xch + ((t * 128 + q * 32 + lane) * 2 + hf);
    const float* peer_x = xch + ((t * 128 + q * 32 + lane) * 2 + (hf ^ 1));
    const int pair_bar = 1 + t * 4 + q;      // named barrier of the two warps that share these rows
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory"); };
    uint32_t n_sfull = 0, n_ofull = 0, n_pv = 0;  // n_pv: P V blocks issued for this tile so far (phases of pv_done)
    constexpr int PT = FMT == 0 ? 1 : 2;
    constexpr int CH = ATL_KB / 64;          // 32-column chunks per half block
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int ue = p.reverse ? num_units - 1 - unit : unit;
      const int qp = ue % p.n_qp, bh = ue / p.n_qp;
      const int b = bh / p.H, h = bh - b * p.H;
      const int nq = (S - qp * 256 > 128) ? 2 : 1;
      if (t >= nq) continue;
      const int row = qp * 256 + t * 128 + q * 32 + lane;
      float ms = -INFINITY;  // reference maximum (scaled by scale_log2), identical in both warps of the row
      const float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
      float2 l2 = make_float2(0.f, 0.f);   // this half's share of the row sum
      uint32_t r[32];
      for (int j = 0; j < p.n_blk; ++j) {
        const int kvalid = min(ATL_KB, S - j * ATL_KB);   // valid keys in this block (only the last block is partial)
        const int n_live = (kvalid + 31) / 32, n_full = kvalid / 32;
        mbar_wait(&s_full[t], n_sfull & 1);
        ++n_sfull;
        tcgen05_fence_after();
        // ---- pass 1: maximum of this half's columns, agreed with the other half through shared memory ----
        float cm = -INFINITY;
        for (int cc = 0; cc < CH; ++cc) {
          const int c = hf * CH + cc;
          if (c >= n_live) break;
          tmem_ld_32x32b_x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c < n_full) {
#pragma unroll
            for (int jj = 0; jj < 32; jj += 2) cm = fmax3(cm, __uint_as_float(r[jj]), __uint_as_float(r[jj + 1]));
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) cm = (c * 32 + jj < kvalid) ? fmaxf(cm, __uint_as_float(r[jj])) : cm;
          }
        }
        *my_x = cm;
        pair_sync();
        const float bs = fmaxf(cm, *peer_x) * p.scale_log2;  // block maximum of the whole row (chunk 0 is always live: finite)
        pair_sync();                                          // both have read before either overwrites its slot again
        const bool raise = bs > ms + 8.0f;                    // the same decision in both warps of the row
        if (__any_sync(0xffffffffu, raise)) {
          // rare after the first block: rescale what the row has accumulated under the old reference (factor 1 for rows that keep it)
          const float f = raise ? ex2_approx(ms - bs) : 1.0f;
          l2.x *= f;
          l2.y *= f;
          if (j > 0) {  // this half's 32 columns of O: wait for the last P V, scale in place
            mbar_wait(&pv_done[t], (n_pv - 1) & 1);
            tcgen05_fence_after();
#pragma unroll
            for (int oc = 0; oc < 2; ++oc) {
              uint32_t ov[16];
              tmem_ld_32x32b_x16(taddr + 192 + hf * 32 + oc * 16, ov);
              tmem_ld_wait();
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) ov[jj] = __float_as_uint(__uint_as_float(ov[jj]) * f);
              tmem_st_32x32b_x16(taddr + 192 + hf * 32 + oc * 16, ov);
            }
            tmem_st_wait();
          }
          if (raise) ms = bs;
        }
        const float2 mo2 = make_float2(-ms, -ms);
        // ---- pass 2: P = exp2(S * scale - ms) over this half's chunks ----
        for (int cc = 0; cc < CH; ++cc) {
          const int c = hf * CH + cc;
          uint32_t pk[16];
          if (c < n_live) {
            tmem_ld_32x32b_x32(taddr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int jj = 0; jj < 32; jj += 2) {
              const float2 a = ffma2(make_float2(__uint_as_float(r[jj]), __uint_as_float(r[jj + 1])), sc2, mo2);
              float2 e = make_float2(ex2_approx(a.x), ex2_approx(a.y));
              if (c >= n_full) {
                e.x = (c * 32 + jj < kvalid) ? e.x : 0.f;
                e.y = (c * 32 + jj + 1 < kvalid) ? e.y : 0.f;
              }
              l2 = fadd2(l2, e);
              pk[jj >> 1] = pack2(e.x, e.y, PT);
            }
          } else {  // keys beyond S: P = 0
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) pk[jj] = 0u;
          }
          // P of half hf lives at columns [96 hf + 16 cc, +16): inside this half's OWN score columns and always behind its read position
          // (chunk 3 hf + cc has just been read), so the two halves never touch each other's unread scores.
          tmem_st_32x32b_x16(taddr + hf * 96 + cc * 16, pk);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
        ++n_pv;
      }
      // ---- output: this half's 32 columns of O / l ----
      *my_x = l2.x + l2.y;
      pair_sync();
      const float inv = 1.0f / (l2.x + l2.y + *peer_x);
      pair_sync();
      mbar_wait(&o_full[t], n_ofull & 1);
      ++n_ofull;
      tcgen05_fence_after();
      uint32_t o0[32];
      tmem_ld_32x32b_x32(taddr + 192 + hf * 32, o0);
      tmem_ld_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[t]);
      if (row < S) {
        const size_t off = (static_cast<size_t>(b) * S + row) * p.D + h * 64 + hf * 32;
        if constexpr (sizeof(OutT) == 2) {
          uint4* dst = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.out) + off);
          constexpr int ot = std::is_same<OutT, __half>::value ? 1 : 2;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            dst[jj] = make_uint4(pack2(__uint_as_float(o0[8 * jj]) * inv, __uint_as_float(o0[8 * jj + 1]) * inv, ot),
                                 pack2(__uint_as_float(o0[8 * jj + 2]) * inv, __uint_as_float(o0[8 * jj + 3]) * inv, ot),
                                 pack2(__uint_as_float(o0[8 * jj + 4]) * inv, __uint_as_float(o0[8 * jj + 5]) * inv, ot),
                                 pack2(__uint_as_float(o0[8 * jj + 6]) * inv, __uint_as_float(o0[8 * jj + 7]) * inv, ot));
        } else {
          float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + off);
          constexpr bool RT = std::is_same<OutT, tf32_t>::value;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            float4 v = make_float4(__uint_as_float(o0[4 * jj]) * inv, __uint_as_float(o0[4 * jj + 1]) * inv, __uint_as_float(o0[4 * jj + 2]) * inv,
                                   __uint_as_float(o0[4 * jj + 3]) * inv);
            if (RT) v = make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w));
            dst[jj] = v;
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp_idx == 2) tmem_dealloc(tmem_base, 512);
}

int make_tensor_map_2d(CUtensorMap* map, int dtype, const void* ptr, int rows, int cols, int ld, int box_rows);  // gemm.cu

template <typename T, typename OutT>
static int atl_launch(const void* qkv, int io_type, void* out, int B, int S, int H, cudaStream_t stream, int reverse) {
  const int D = H * 64;
  CUtensorMap map_q, map_kv;
  if (int rc = make_tensor_map_2d(&map_q, io_type, qkv, B * S, 3 * D, 3 * D, 256)) return rc;
  if (int rc = make_tensor_map_2d(&map_kv, io_type, qkv, B * S, 3 * D, 3 * D, ATL_KB)) return rc;
  AtlParams p;
  p.B = B; p.S = S; p.H = H; p.D = D;
  p.n_qp = (S + 255) / 256;
  p.n_blk = (S + ATL_KB - 1) / ATL_KB;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.out = out;
  p.reverse = reverse;
  const long long units = static_cast<long long>(B) * H * p.n_qp;
  const int grid = units < device_sm_count() ? static_cast<int>(units) : device_sm_count();
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    JIMM_CUDA_CHECK(cudaFuncSetAttribute(attention_tc_long_kernel<T, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATL_SMEM));
  }
  JIMM_CUDA_CHECK(launch_k(attention_tc_long_kernel<T, OutT>, dim3(grid), dim3(ATL_THREADS), ATL_SMEM, stream, 1, true, map_q, map_kv, p));
  note_launch();
  return 0;
}

// Returns 1 when this configuration is not handled here (caller falls back to the flash kernel).
int attention_tc_long_run(const void* qkv, int io_type, void* out, int out_type, int B, int S, int H, int causal, cudaStream_t stream, int reverse) {
  if (S <= 256 || causal) return 1;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return 1;
  if (io_type == DT_F16 && out_type == DT_F16) return atl_launch<__half, __half>(qkv, io_type, out, B, S, H, stream, reverse);
  if (io_type == DT_F16 && out_type == DT_F32) return atl_launch<__half, float>(qkv, io_type, out, B, S, H, stream, reverse);
  if (io_type == DT_F16 && out_type == DT_TF32) return atl_launch<__half, tf32_t>(qkv, io_type, out, B, S, H, stream, reverse);
  if (io_type == DT_BF16 && out_type == DT_BF16) return atl_launch<__nv_bfloat16, __nv_bfloat16>(qkv, io_type, out, B, S, H, stream, reverse);
  if (io_type == DT_BF16 && out_type == DT_F32) return atl_launch<__nv_bfloat16, float>(qkv, io_type, out, B, S, H, stream, reverse);
  return 1;
}

}  // namespace jimm
