// Fused L2-normalise -> all-gather over NVLink peer memory -> contrastive logits, in ONE kernel.
//
// Reference op: under batch sharding the only place data from different shards meets is
//   logits = exp(logit_scale) * image_features @ text_features.T (+ logit_bias)      models/clip.py:183-187, models/siglip.py:169-173
// which XLA's GSPMD resolves with an all-gather of the features.  Here every rank
//   phase 1  L2-normalises its [B_local, E] image and text rows and stores them straight into EVERY peer's gather buffer
//            (st.global on CUDA-IPC-mapped peer pointers -> NVLink 5 / NVSwitch), row (rank*B_local + r), cols [0,E) | [E,2E);
//   phase 2  fences (system scope); the last CTA to finish publishes flag[rank] = epoch into every peer (st.release.sys);
//   phase 3  every CTA spins (ld.acquire.sys) on its LOCAL flags until all ranks have published this epoch -- BOUNDED: a peer that
//            does not publish within the timeout (default 10 s, JIMM_COMM_TIMEOUT_MS) or that published a different B_local makes the
//            kernel write an error word to a host-mapped status flag and NaN logits instead of hanging; the next comm call (or
//            jimm_comm_status) reports it;
//   phase 4  computes its own logits row block [B_local, world*B_local] from local memory only (fp32 FMA tiles).
// The grid is persistent (<= #SMs CTAs, 1 CTA/SM) so the phase-3 spin cannot starve phase 1/2 of a co-resident CTA.
// Two parity buffers make back-to-back calls safe without a host barrier: a peer can only be one epoch ahead.
#include <stdlib.h>
#include <string.h>

#include "comm.cuh"
#include "common.cuh"
#include "gemm.cuh"
#include "logits_tile.cuh"

namespace jimm {

static constexpr size_t kFlagsBytes = 4096;

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

struct CommPtrs {
  float* buf[kMaxWorld];          // this epoch's parity buffer on each rank
  unsigned int* flags[kMaxWorld]; // flags array on each rank: [0,kMaxWorld) epoch published by rank r, [kMaxWorld, 2 kMaxWorld) its B_local
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(256, 1)
comm_logits_kernel(CommPtrs ptrs, const float* __restrict__ img_e, const float* __restrict__ txt_e, int B_local, int E, int rank,
                   int world, unsigned int epoch, unsigned int* counter, const float* __restrict__ logit_scale,
                   const float* __restrict__ logit_bias, float* __restrict__ logits_local, unsigned long long timeout_ns,
                   unsigned int* status /* host-mapped */) {
  __shared__ __align__(16) float As[16][LOGITS_LDS], Bs[16][LOGITS_LDS];
  __shared__ int s_last;
  __shared__ unsigned int s_err;
  if (threadIdx.x == 0) s_err = 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t ld = static_cast<size_t>(2) * E;

  // ---- phase 1: normalise + scatter to all peers ----
  const int total_rows = 2 * B_local;  // image rows then text rows
  for (int r = blockIdx.x * 8 + warp; r < total_rows; r += gridDim.x * 8) {
    const bool is_txt = r >= B_local;
    const int row = is_txt ? r - B_local : r;
    const float* src = (is_txt ? txt_e : img_e) + static_cast<size_t>(row) * E;
    float ss = 0.f;
    for (int i = lane; i < E; i += 32) { const float v = src[i]; ss += v * v; }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const size_t dst_off = (static_cast<size_t>(rank) * B_local + row) * ld + (is_txt ? E : 0);
    for (int i = lane; i < E; i += 32) {
      const float v = src[i] / nrm;
      for (int p = 0; p < world; ++p) ptrs.buf[p][dst_off + i] = v;
    }
  }
  // ---- phase 2: publish ----
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(counter, 1u);
    s_last = (ticket == epoch * gridDim.x - 1u) ? 1 : 0;
  }
  __syncthreads();
  if (s_last && threadIdx.x < world) {
    __threadfence_system();
    ptrs.flags[threadIdx.x][kMaxWorld + rank] = static_cast<unsigned int>(B_local);  // ordered before the flag by the release below
    st_release_sys(ptrs.flags[threadIdx.x] + rank, epoch);
  }
  // ---- phase 3: wait (bounded) for every rank's rows of this epoch; every rank must have sent the same number of rows ----
  if (threadIdx.x < world) {
    const unsigned int* f = ptrs.flags[rank] + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    unsigned int err = 0u;
    while (static_cast<int>(ld_acquire_sys(f) - epoch) < 0) {
      __nanosleep(64);
      if (globaltimer_ns() - t0 > timeout_ns) { err = 1u | (static_cast<unsigned int>(threadIdx.x) << 8); break; }
    }
    if (err == 0u && ptrs.flags[rank][kMaxWorld + threadIdx.x] != static_cast<unsigned int>(B_local)) err = 2u | (static_cast<unsigned int>(threadIdx.x) << 8);
    if (err != 0u) atomicMax(&s_err, err);
  }
  __syncthreads();
  const unsigned int err = s_err;
  if (err != 0u && blockIdx.x == 0 && threadIdx.x == 0) {
    *status = err | (epoch << 16);  // 1: peer (bits 8..15) never published this epoch; 2: peer sent a different B_local
    __threadfence_system();
  }
  // ---- phase 4: local logits row block ----
  const float sc = expf(*logit_scale);
  const float bs = logit_bias ? *logit_bias : 0.f;
  const int Bt = world * B_local;
  const float* gathered = ptrs.buf[rank];
  const float* A = gathered + static_cast<size_t>(rank) * B_local * ld;  // my normalised image rows
  const float* Bm = gathered + E;                                         // all normalised text rows
  const int ti = (B_local + 63) / 64, tj = (Bt + 63) / 64;
  if (err != 0u) {  // no valid gathered batch: make the result visibly invalid instead of silently stale
    const float qnan = __int_as_float(0x7fc00000);
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < static_cast<size_t>(B_local) * Bt; i += static_cast<size_t>(gridDim.x) * blockDim.x)
      logits_local[i] = qnan;
    return;
  }
  for (int t = blockIdx.x; t < ti * tj; t += gridDim.x) {
    const int i0 = (t / tj) * 64, j0 = (t % tj) * 64;
    logits_tile<true>(A, ld, Bm, ld, logits_local, Bt, B_local, Bt, E, i0, j0, sc, bs, As, Bs);
  }
}

int comm_init(CommState* c, int rank, int world, int max_rows, int E, unsigned char* handle_out) {
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || max_rows <= 0 || E <= 0 || !handle_out) {
    set_last_error("comm_init: bad arguments (rank %d world %d rows %d E %d)", rank, world, max_rows, E);
    return -1;
  }
  if (c->ready) { set_last_error("comm already initialised"); return -4; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  c->rank = rank; c->world = world; c->max_rows = max_rows; c->E = E;
  c->buf_floats = static_cast<size_t>(world) * max_rows * 2 * E;
  const size_t bytes = 2 * c->buf_floats * sizeof(float) + kFlagsBytes;
  JIMM_CUDA_CHECK(cudaMalloc(&c->base, bytes));
  JIMM_CUDA_CHECK(cudaMemset(c->base, 0, bytes));
  JIMM_CUDA_CHECK(cudaMalloc(&c->counter, sizeof(unsigned int)));
  JIMM_CUDA_CHECK(cudaMemset(c->counter, 0, sizeof(unsigned int)));
  JIMM_CUDA_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  JIMM_CUDA_CHECK(cudaIpcGetMemHandle(&h, c->base));
  memcpy(handle_out, &h, sizeof(h));
  for (int i = 0; i < kMaxWorld; ++i) c->peer_base[i] = nullptr;
  c->peer_base[rank] = c->base;
  c->local_buf = static_cast<float*>(c->base);
  // One CTA per SM: BASELINE configs[4] is 128 logits tiles per rank (64 CTAs took two rounds and lost to the NCCL path, r2_bench_n8).
  // Every CTA must become resident for the rank to publish (the last-CTA ticket), and resident CTAs spin on the peers' flags: the grid has
  // to fit the SMs that are free when the head runs.  At 95 registers x 256 threads and 8.7 KB of shared memory two CTAs fit one SM, so
  // #SM CTAs need half the machine; the head is launched after the towers have joined, with nothing else in flight.  A launch that cannot
  // become resident ends in the bounded wait's timeout (NaN logits + jimm_comm_status), not in a hang.
  c->grid = device_sm_count();
  JIMM_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&c->status_host), sizeof(unsigned int), cudaHostAllocMapped));
  *c->status_host = 0u;
  JIMM_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->status_dev), c->status_host, 0));
  c->timeout_ns = 10ull * 1000 * 1000 * 1000;
  if (const char* env = getenv("JIMM_COMM_TIMEOUT_MS")) { if (atoll(env) > 0) c->timeout_ns = static_cast<unsigned long long>(atoll(env)) * 1000000ull; }
  c->epoch = 0;
  c->ready = true;
  c->connected = (world == 1);
  return 0;
}

int comm_connect(CommState* c, const unsigned char* handles) {
  if (!c->ready) { set_last_error("comm_connect before comm_init"); return -4; }
  if (!handles) { set_last_error("comm_connect: null handles"); return -1; }
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + static_cast<size_t>(r) * 64, 64);
    void* p = nullptr;
    JIMM_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_base[r] = p;
  }
  c->connected = true;
  return 0;
}

int comm_contrastive_logits(CommState* c, const float* img_e, const float* txt_e, int B_local, const float* logit_scale,
                            const float* logit_bias, float* logits_local, cudaStream_t stream) {
  if (!c->ready || !c->connected) { set_last_error("comm not initialised / connected"); return -4; }
  if (B_local <= 0 || B_local > c->max_rows) { set_last_error("comm: B_local %d outside (0, %d]", B_local, c->max_rows); return -1; }
  if (int rc = comm_status(c)) return rc;  // an earlier call timed out / met a mismatched peer: its logits are NaN, say why
  c->epoch += 1;
  const int parity = static_cast<int>(c->epoch & 1);
  CommPtrs ptrs;
  for (int r = 0; r < kMaxWorld; ++r) { ptrs.buf[r] = nullptr; ptrs.flags[r] = nullptr; }
  for (int r = 0; r < c->world; ++r) {
    uint8_t* b = static_cast<uint8_t*>(c->peer_base[r]);
    ptrs.buf[r] = reinterpret_cast<float*>(b) + static_cast<size_t>(parity) * c->buf_floats;
    ptrs.flags[r] = reinterpret_cast<unsigned int*>(b + 2 * c->buf_floats * sizeof(float));
  }
  c->local_buf = ptrs.buf[c->rank];
  comm_logits_kernel<<<c->grid, 256, 0, stream>>>(ptrs, img_e, txt_e, B_local, c->E, c->rank, c->world,
                                                  static_cast<unsigned int>(c->epoch), c->counter, logit_scale, logit_bias, logits_local,
                                                  c->timeout_ns, c->status_dev);
  JIMM_LAUNCH_CHECK();
  return 0;
}

// 0, or an error describing the first failed exchange (sticky until comm_destroy)
int comm_status(CommState* c) {
  if (!c->ready || !c->status_host) return 0;
  const unsigned int st = *reinterpret_cast<volatile unsigned int*>(c->status_host);
  if (st == 0u) return 0;
  const unsigned int code = st & 0xffu, peer = (st >> 8) & 0xffu, ep = st >> 16;
  if (code == 1u) set_last_error("comm: rank %u did not publish its embeddings for exchange %u within %llu ms (peer died, or skipped the call)", peer, ep,
                                 c->timeout_ns / 1000000ull);
  else set_last_error("comm: rank %u sent a different number of rows than this rank in exchange %u (every rank must pass the same B_local)", peer, ep);
  return -2;
}

void comm_destroy(CommState* c) {
  if (!c->ready) return;
  if (c->status_host) { cudaFreeHost(c->status_host); c->status_host = nullptr; c->status_dev = nullptr; }
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
  if (c->base) cudaFree(c->base);
  if (c->counter) cudaFree(c->counter);
  c->ready = false;
}

}  // namespace jimm
