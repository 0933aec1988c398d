// Micro-benchmark, not on the product path: how many bytes per second the SMs can pull from L2 through TMA, and whether
// cluster multicast raises that number.  Motivation: every tcgen05 GEMM of the encoder runs at
// 10.0-10.4 TB/s of L2->SM traffic whatever its shape (QKV 1.40 GB in 139 us, FC2 1.86 GB in 179 us), i.e. the 256x256 pair
// tile is bound by the fill bandwidth, not by the tensor pipe.
//
// Every CTA keeps a ring of 16 KB stages (128 rows of 128 bytes) filled by TMA from an L2-resident buffer and frees a stage
// as soon as it has landed.  mode 0: every CTA loads its own tiles.  mode 1: the CS CTAs of a cluster need the SAME tile and
// each loads all of it (what CTA pairs sharing a B tile do today).  mode 2: same tile, each CTA loads 1/CS of it and
// multicasts its part to the whole cluster.
#include "../../include/jimm_b200.h"
#include "common.cuh"
#include "gemm.cuh"

namespace jimm {

int make_tensor_map_2d(CUtensorMap* map, int dtype, const void* ptr, int rows, int cols, int ld, int box_rows);  // gemm.cu

namespace {

constexpr int PB_STAGES = 12;
constexpr int PB_STAGE_BYTES = 128 * 128;

__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

template <int CS>
__global__ void __launch_bounds__(64, 1) l2_probe_kernel(const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_part,
                                                         int mode, int iters, int total_rows) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + PB_STAGES * PB_STAGE_BYTES);
  uint64_t* empty = full + PB_STAGES;
  const uint32_t rank = CS > 1 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CS;
  const bool shared_writes = CS > 1 && mode == 2;  // other CTAs write into my stages: they must know when I have freed them
  if (threadIdx.x == 0) {
    for (int i = 0; i < PB_STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], shared_writes ? CS : 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (CS > 1) cluster_sync_all();
  const int tiles = total_rows / 128;
  if (threadIdx.x == 0) {
    // producer
    int st = 0;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      const long long tile = mode == 0 ? (static_cast<long long>(blockIdx.x) * iters + it) : (static_cast<long long>(cluster_id) * iters + it);
      const int row = static_cast<int>((tile * 7919) % tiles) * 128;
      mbar_wait(&empty[st], ph ^ 1);
      mbar_arrive_expect_tx(&full[st], PB_STAGE_BYTES);
      if (mode == 2 && CS > 1) {
        constexpr int part = 128 / CS;
        tma_load_2d_multicast(smem + st * PB_STAGE_BYTES + rank * part * 128, &map_part, &full[st], 0, row + rank * part,
                              static_cast<uint16_t>((1u << CS) - 1));
      } else {
        tma_load_2d(smem + st * PB_STAGE_BYTES, &map_full, &full[st], 0, row);
      }
      if (++st == PB_STAGES) { st = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    // consumer: a stage is "used" the moment it has landed
    int st = 0;
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&full[st], ph);
      if (shared_writes) {
        for (uint32_t r = 0; r < CS; ++r)
          asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_shared(smem_u32(&empty[st]), r)) : "memory");
      } else {
        mbar_arrive(&empty[st]);
      }
      if (++st == PB_STAGES) { st = 0; ph ^= 1; }
    }
  }
  __syncthreads();
  if (CS > 1) cluster_sync_all();  // nobody leaves while a peer may still write into its shared memory
}

template <int CS>
int probe_launch(const CUtensorMap& mf, const CUtensorMap& mp, int mode, int iters, int rows, int grid, cudaStream_t s) {
  const size_t smem = PB_STAGES * PB_STAGE_BYTES + 2 * PB_STAGES * 8 + 1024;
  JIMM_CUDA_CHECK(cudaFuncSetAttribute(l2_probe_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  JIMM_CUDA_CHECK(launch_k(l2_probe_kernel<CS>, dim3(grid), dim3(64), smem, s, CS, false, mf, mp, mode, iters, rows));
  return 0;
}

}  // namespace
}  // namespace jimm

using namespace jimm;

// buf: device buffer of rows x 128 bytes (fp16 [rows, 64]).  Returns through *ms the kernel time of `iters` stage fills per CTA.
extern "C" int jimm_k_l2_probe(const void* buf, int rows, int mode, int cluster, int iters, float* ms, void* stream) {
  if (!buf || rows < 128 || rows % 128 || (cluster != 1 && cluster != 2 && cluster != 4 && cluster != 8) || mode < 0 || mode > 2 || iters <= 0 || !ms) {
    set_last_error("l2_probe: bad arguments");
    return JIMM_EINVAL;
  }
  CUtensorMap mf, mp;
  if (int rc = make_tensor_map_2d(&mf, DT_F16, buf, rows, 64, 64, 128)) return rc;
  if (int rc = make_tensor_map_2d(&mp, DT_F16, buf, rows, 64, 64, 128 / cluster)) return rc;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = sms / cluster * cluster;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  JIMM_CUDA_CHECK(cudaEventCreate(&e0));
  JIMM_CUDA_CHECK(cudaEventCreate(&e1));
  int rc = 0;
  for (int rep = 0; rep < 2 && rc == 0; ++rep) {  // first launch warms L2 and the instruction cache
    JIMM_CUDA_CHECK(cudaEventRecord(e0, s));
    if (cluster == 1) rc = probe_launch<1>(mf, mp, mode, iters, rows, grid, s);
    else if (cluster == 2) rc = probe_launch<2>(mf, mp, mode, iters, rows, grid, s);
    else if (cluster == 4) rc = probe_launch<4>(mf, mp, mode, iters, rows, grid, s);
    else rc = probe_launch<8>(mf, mp, mode, iters, rows, grid, s);
    JIMM_CUDA_CHECK(cudaEventRecord(e1, s));
    JIMM_CUDA_CHECK(cudaEventSynchronize(e1));
  }
  if (rc == 0) JIMM_CUDA_CHECK(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}
