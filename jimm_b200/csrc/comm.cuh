// Multi-GPU contrastive head over NVLink peer memory: one process per GPU; every rank owns a symmetric
// gather buffer that all peers map through CUDA IPC.  See comm.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace jimm {

static constexpr int kMaxWorld = 16;

struct CommState {
  bool ready = false;
  bool connected = false;
  int rank = 0, world = 1, max_rows = 0, E = 0;
  size_t buf_floats = 0;            // floats per parity buffer: world * max_rows * 2E
  void* base = nullptr;             // local allocation: [2 parity buffers][flags]
  float* local_buf = nullptr;       // parity buffer used by the most recent call
  void* peer_base[kMaxWorld] = {};  // mapped base pointers (peer_base[rank] == base)
  unsigned int* counter = nullptr;  // local CTA-completion ticket counter
  unsigned long long epoch = 0;
  int grid = 0;
  unsigned int* status_host = nullptr;  // host-mapped error word written by the kernel (0 = fine)
  unsigned int* status_dev = nullptr;
  unsigned long long timeout_ns = 0;    // bound of the device-side wait for the peers
};

int comm_init(CommState* c, int rank, int world, int max_rows, int E, unsigned char* handle_out);
int comm_connect(CommState* c, const unsigned char* handles);
int comm_contrastive_logits(CommState* c, const float* img_e, const float* txt_e, int B_local, const float* logit_scale,
                            const float* logit_bias, float* logits_local, cudaStream_t stream);
int comm_status(CommState* c);
void comm_destroy(CommState* c);

}  // namespace jimm
