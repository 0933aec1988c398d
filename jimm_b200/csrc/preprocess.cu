// Image front-end on the GPU: what the reference's examples run on the host before the forward path
// (examples/vit_inference.py:27-37, examples/clip_inference.py:35-38): HuggingFace image processor = Pillow 8-bit resize
// (bilinear / bicubic with antialiasing) -> optional centre crop -> rescale by 1/255 -> per-channel normalise -> NHWC.
//
// One fused kernel, input read once and output written once.  A CTA owns TY output rows of one image:
//   1. the input rows its vertical taps need go, one warp per row (coalesced 16-byte loads into a per-warp row buffer, no
//      CTA-wide synchronisation), through the horizontal pass, for the cropped columns only, into an 8-bit tile in shared
//      memory -- Pillow's temporary image, never written to HBM;
//   2. the vertical pass reads that tile four bytes per thread, and each resulting 8-bit sample goes through a 768-entry
//      table (channel, value) -> normalised float, then to the output dtype.
// All arithmetic is Pillow's: int32 fixed point with 22 fractional bits, taps and weights from the same double-precision
// recipe, so the result is bit-exact (oracle/preprocess_oracle.py restates it and is pinned against Pillow itself).
// HBM-bound byte work: algorithmic bytes = H*W*3 read + oh*ow*3*sizeof(out) written per image.
#include <cmath>
#include <map>
#include <utility>
#include <vector>

#include "../../include/jimm_b200.h"
#include "common.cuh"

#define JIMM_TRY(expr)          \
  do {                          \
    const int _rc = (expr);     \
    if (_rc != 0) return _rc;   \
  } while (0)

namespace jimm {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;
constexpr int kThreads = 256;

// ---- host: Pillow's coefficient recipe (Resample.c precompute_coeffs + normalize_coeffs_8bpc) ----
double filter_bilinear(double x) {
  if (x < 0.0) x = -x;
  return x < 1.0 ? 1.0 - x : 0.0;
}
double filter_bicubic(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

struct ResampleTable {
  int ksize = 0, kpad = 0;            // Pillow's window capacity; the same rounded up to a multiple of 4
  std::vector<int> first, count, kk;  // kk[out][ksize]

  // Device layout: rows padded with zero weights to kpad taps so the kernels run groups of four taps without a tail (a zero
  // weight makes whatever byte it meets irrelevant).
  std::vector<int> padded(int stride) const {
    const size_t n = first.size();
    std::vector<int> p(n * stride, 0);
    for (size_t i = 0; i < n; ++i)
      for (int k = 0; k < ksize; ++k) p[i * stride + k] = kk[i * ksize + k];
    return p;
  }
  // Row stride for shared memory: an odd number of 16-byte units, so a warp's int4 reads of consecutive rows are conflict-free.
  int smem_stride() const { return (kpad / 4) % 2 ? kpad : kpad + 4; }
};

ResampleTable make_table(int in_size, int out_size, int resample) {
  double (*f)(double) = resample == 3 ? filter_bicubic : filter_bilinear;
  const double fsupport = resample == 3 ? 2.0 : 1.0;
  const double scale = static_cast<double>(in_size) / out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = fsupport * filterscale;
  ResampleTable t;
  t.ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  t.kpad = (t.ksize + 3) / 4 * 4;
  t.first.assign(out_size, 0);
  t.count.assign(out_size, 0);
  t.kk.assign(static_cast<size_t>(out_size) * t.ksize, 0);
  std::vector<double> w(t.ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = f((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      const double k = ww != 0.0 ? w[x] / ww : w[x];
      t.kk[static_cast<size_t>(xx) * t.ksize + x] =
          k < 0 ? static_cast<int>(-0.5 + k * (1 << kPrecisionBits)) : static_cast<int>(0.5 + k * (1 << kPrecisionBits));
    }
    t.first[xx] = xmin;
    t.count[xx] = xmax;
  }
  return t;
}

struct KernelArgs {
  const uint8_t* img;
  void* out;
  const float* lut;                    // [3][256] normalised value of an 8-bit sample
  const int *hfirst, *hcount, *hk;     // horizontal tables, already offset to the first cropped column
  const int *vfirst, *vcount, *vk;     // vertical tables, already offset to the first cropped row
  int H, W, oh, ow, hks, hstride, vks;  // hks / vks: taps rounded up to a multiple of 4; hstride: row stride of hk
  int TY;                              // output rows per CTA
  int rowb;                            // bytes per tile row (ow*3 rounded up to 4)
  int stage_bytes;                     // one warp's row buffer (W*3 + 32, rounded to 16)
  int vec_ok;                          // img is 16-byte aligned
};

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

template <typename OUT>
__device__ __forceinline__ OUT to_out(float v);
template <> __device__ __forceinline__ float to_out<float>(float v) { return v; }
template <> __device__ __forceinline__ __half to_out<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 to_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename OUT>
__global__ void __launch_bounds__(kThreads) preprocess_kernel(const KernelArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  int* hk_s = reinterpret_cast<int*>(smem);                 // [ow][hstride]
  int* hfirst_s = hk_s + a.ow * a.hstride;  // byte offset of the window in a row | number of four-tap groups << 24
  uint8_t* stage = reinterpret_cast<uint8_t*>(hfirst_s + a.ow);
  stage += (16 - (reinterpret_cast<uintptr_t>(stage) & 15)) & 15;
  uint8_t* tile = stage + static_cast<size_t>(kThreads / 32) * a.stage_bytes;  // stage = one row buffer per warp

  const int tid = threadIdx.x;
  const int b = blockIdx.y;
  const int yo0 = blockIdx.x * a.TY;
  const int yo1 = min(yo0 + a.TY, a.oh);
  const int in_y0 = a.vfirst[yo0];
  const int in_y1 = a.vfirst[yo1 - 1] + a.vcount[yo1 - 1];

  {
    const int4* src = reinterpret_cast<const int4*>(a.hk);
    int4* dst = reinterpret_cast<int4*>(hk_s);
    for (int i = tid; i < a.ow * a.hstride / 4; i += kThreads) dst[i] = __ldg(src + i);
    for (int i = tid; i < a.ow; i += kThreads) hfirst_s[i] = (a.hfirst[i] * 3) | (((a.hcount[i] + 3) >> 2) << 24);
  }

  const size_t row_bytes = static_cast<size_t>(a.W) * 3;
  const uint8_t* img_b = a.img + static_cast<size_t>(b) * a.H * row_bytes;
  __syncthreads();
  // ---- horizontal pass: one warp per input row, no CTA-wide synchronisation (warps hide each other's load latency) ----
  const int warp = tid >> 5, lane = tid & 31;
  uint8_t* wb = stage + static_cast<size_t>(warp) * a.stage_bytes;
  const uint32_t* wb32 = reinterpret_cast<const uint32_t*>(wb);
  for (int row = warp; row < in_y1 - in_y0; row += kThreads / 32) {
    const uint8_t* src = img_b + static_cast<size_t>(in_y0 + row) * row_bytes;
    const int bytes = static_cast<int>(row_bytes);
    const int mis = a.vec_ok ? static_cast<int>(reinterpret_cast<uintptr_t>(src) & 15) : 0;  // data starts at wb + mis
    if (a.vec_ok) {
      const int head = (16 - mis) & 15;  // bytes before the first aligned vector
      const int nvec = bytes > head ? (bytes - head) / 16 : 0;
      const uint4* vsrc = reinterpret_cast<const uint4*>(src + head);
      uint4* vdst = reinterpret_cast<uint4*>(wb + mis + head);
      for (int i = lane; i < nvec; i += 32) vdst[i] = __ldg(vsrc + i);
      const int tail0 = head + nvec * 16;
      for (int i = lane; i < head && i < bytes; i += 32) wb[mis + i] = __ldg(src + i);
      for (int i = tail0 + lane; i < bytes; i += 32) wb[mis + i] = __ldg(src + i);
    } else {
      for (int i = lane; i < bytes; i += 32) wb[i] = __ldg(src + i);
    }
    __syncwarp();
    uint8_t* trow = tile + static_cast<size_t>(row) * a.rowb;
    for (int xo = lane; xo < a.ow; xo += 32) {
      // The window is a run of RGB bytes from an arbitrary byte offset: read aligned words, realign them with funnel shifts;
      // four taps = twelve bytes = three realigned words (R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3) and one int4 of weights.
      const int fg = hfirst_s[xo];
      const int start = (fg & 0xffffff) + mis, hgroups = fg >> 24;
      const uint32_t* wp = wb32 + (start >> 2);
      const uint32_t sh = (start & 3) * 8;
      const int4* kp = reinterpret_cast<const int4*>(hk_s + xo * a.hstride);
      int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
      uint32_t prev = wp[0];
      for (int g = 0; g < hgroups; ++g) {
        const uint32_t w1 = wp[3 * g + 1], w2 = wp[3 * g + 2], w3 = wp[3 * g + 3];
        const int4 k = kp[g];
        const uint32_t s0 = __funnelshift_r(prev, w1, sh), s1 = __funnelshift_r(w1, w2, sh), s2 = __funnelshift_r(w2, w3, sh);
        prev = w3;
        a0 += static_cast<int>(s0 & 255u) * k.x;
        a1 += static_cast<int>((s0 >> 8) & 255u) * k.x;
        a2 += static_cast<int>((s0 >> 16) & 255u) * k.x;
        a0 += static_cast<int>(s0 >> 24) * k.y;
        a1 += static_cast<int>(s1 & 255u) * k.y;
        a2 += static_cast<int>((s1 >> 8) & 255u) * k.y;
        a0 += static_cast<int>((s1 >> 16) & 255u) * k.z;
        a1 += static_cast<int>(s1 >> 24) * k.z;
        a2 += static_cast<int>(s2 & 255u) * k.z;
        a0 += static_cast<int>((s2 >> 8) & 255u) * k.w;
        a1 += static_cast<int>((s2 >> 16) & 255u) * k.w;
        a2 += static_cast<int>(s2 >> 24) * k.w;
      }
      trow[xo * 3] = static_cast<uint8_t>(clip8(a0));
      trow[xo * 3 + 1] = static_cast<uint8_t>(clip8(a1));
      trow[xo * 3 + 2] = static_cast<uint8_t>(clip8(a2));
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- vertical pass + normalise + store: four consecutive samples per thread, four taps per step ----
  const int words = a.rowb / 4;
  const int n_el = a.ow * 3;
  const uint32_t* tile32 = reinterpret_cast<const uint32_t*>(tile);
  OUT* out = static_cast<OUT*>(a.out);
  for (int it = tid; it < (yo1 - yo0) * words; it += kThreads) {
    const int r = it / words, wd = it - r * words;
    const int yo = yo0 + r;
    const uint32_t* tp = tile32 + (a.vfirst[yo] - in_y0) * words + wd;
    const int vgroups = (a.vcount[yo] + 3) >> 2;
    const int4* kp = reinterpret_cast<const int4*>(a.vk + static_cast<size_t>(yo) * a.vks);
    int acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = 1 << (kPrecisionBits - 1);
    for (int g = 0; g < vgroups; ++g) {
      const int4 k = __ldg(kp + g);
      const uint32_t t0 = tp[(4 * g) * words], t1 = tp[(4 * g + 1) * words], t2 = tp[(4 * g + 2) * words], t3 = tp[(4 * g + 3) * words];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j] += static_cast<int>((t0 >> (8 * j)) & 255u) * k.x;
        acc[j] += static_cast<int>((t1 >> (8 * j)) & 255u) * k.y;
        acc[j] += static_cast<int>((t2 >> (8 * j)) & 255u) * k.z;
        acc[j] += static_cast<int>((t3 >> (8 * j)) & 255u) * k.w;
      }
    }
    const int e0 = wd * 4;
    const size_t base = (static_cast<size_t>(b) * a.oh + yo) * n_el + e0;
    OUT v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (e0 + j) % 3;
      v[j] = to_out<OUT>(__ldg(a.lut + c * 256 + clip8(acc[j])));
    }
    if (e0 + 3 < n_el && (base * sizeof(OUT)) % (4 * sizeof(OUT)) == 0) {
      if constexpr (sizeof(OUT) == 4) {
        *reinterpret_cast<float4*>(out + base) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        uint2 pk;
        pk.x = static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&v[0])) | (static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&v[1])) << 16);
        pk.y = static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&v[2])) | (static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&v[3])) << 16);
        *reinterpret_cast<uint2*>(out + base) = pk;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (e0 + j < n_el) out[base + j] = v[j];
    }
  }
}

struct DevTable {
  int ksize = 0, stride = 0;
  int *first = nullptr, *count = nullptr, *kk = nullptr;
};

struct SizePlan {
  int rh = 0, rw = 0, top = 0, left = 0, oh = 0, ow = 0;
  DevTable h, v;
  int TY = 0, tile_rows = 0, rowb = 0, stage_bytes = 0;
  size_t smem = 0;
};

}  // namespace
}  // namespace jimm

using namespace jimm;

struct jimm_preproc {
  jimm_preproc_config_t cfg;
  int device = 0;
  float* lut = nullptr;
  std::map<std::pair<int, int>, SizePlan> plans;
  std::vector<void*> allocs;
};

namespace {

int upload(jimm_preproc* p, const std::vector<int>& v, int** out) {
  void* d = nullptr;
  JIMM_CUDA_CHECK(cudaMalloc(&d, v.size() * sizeof(int)));
  p->allocs.push_back(d);
  JIMM_CUDA_CHECK(cudaMemcpy(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice));
  *out = static_cast<int*>(d);
  return 0;
}

void resized_size(const jimm_preproc_config_t& c, int H, int W, int* rh, int* rw) {
  if (!c.shortest_edge) { *rh = c.height; *rw = c.width; return; }
  // transformers get_resize_output_image_size(size=shortest_edge, default_to_square=False)
  const int s = W <= H ? W : H, l = W <= H ? H : W;
  const int new_long = static_cast<int>(static_cast<double>(c.shortest_edge) * l / s);
  if (W <= H) { *rw = c.shortest_edge; *rh = new_long; } else { *rh = c.shortest_edge; *rw = new_long; }
}

int check_cfg(const jimm_preproc_config_t* c) {
  if (!c) { set_last_error("null preprocessing config"); return JIMM_EINVAL; }
  if (c->resample != 2 && c->resample != 3) { set_last_error("resample must be 2 (bilinear) or 3 (bicubic), got %d", c->resample); return JIMM_EINVAL; }
  if (!c->shortest_edge && (c->height <= 0 || c->width <= 0)) { set_last_error("size needs height and width, or shortest_edge"); return JIMM_EINVAL; }
  if ((c->crop_h > 0) != (c->crop_w > 0)) { set_last_error("crop needs both height and width"); return JIMM_EINVAL; }
  for (int i = 0; i < 3; ++i)
    if (c->std[i] == 0.f) { set_last_error("std evaluated to zero, leading to division by zero."); return JIMM_EINVAL; }
  return 0;
}

int get_plan(jimm_preproc* p, int H, int W, SizePlan** out) {
  auto it = p->plans.find({H, W});
  if (it != p->plans.end()) { *out = &it->second; return 0; }
  if (H <= 0 || W <= 0) { set_last_error("bad image size %dx%d", H, W); return JIMM_EINVAL; }
  SizePlan s;
  resized_size(p->cfg, H, W, &s.rh, &s.rw);
  s.oh = p->cfg.crop_h ? p->cfg.crop_h : s.rh;
  s.ow = p->cfg.crop_w ? p->cfg.crop_w : s.rw;
  if (s.oh > s.rh || s.ow > s.rw) {
    set_last_error("centre crop %dx%d larger than the resized image %dx%d", s.oh, s.ow, s.rh, s.rw);
    return JIMM_EINVAL;
  }
  s.top = (s.rh - s.oh) / 2;
  s.left = (s.rw - s.ow) / 2;
  ResampleTable th = make_table(W, s.rw, p->cfg.resample), tv = make_table(H, s.rh, p->cfg.resample);
  s.h.ksize = th.kpad;
  s.h.stride = th.smem_stride();
  s.v.ksize = s.v.stride = tv.kpad;
  JIMM_TRY(upload(p, th.first, &s.h.first));
  JIMM_TRY(upload(p, th.count, &s.h.count));
  JIMM_TRY(upload(p, th.padded(s.h.stride), &s.h.kk));
  JIMM_TRY(upload(p, tv.first, &s.v.first));
  JIMM_TRY(upload(p, tv.count, &s.v.count));
  JIMM_TRY(upload(p, tv.padded(s.v.stride), &s.v.kk));
  // shared-memory budget: horizontal tables for the cropped columns + staging group + 8-bit tile
  const size_t row_bytes = static_cast<size_t>(W) * 3;
  if (row_bytes + 3 * th.kpad + 48 > 64 * 1024) { set_last_error("image width %d too large for the staging buffer", W); return JIMM_EINVAL; }
  s.rowb = (s.ow * 3 + 3) / 4 * 4;
  const size_t tables = (static_cast<size_t>(s.ow) * s.h.stride + s.ow) * sizeof(int) + 16;
  s.stage_bytes = static_cast<int>((row_bytes + 15 + 3 * th.kpad + 8 + 15) / 16 * 16);  // alignment shift + zero-weight taps past the row + word read-ahead
  // Largest tile of output rows (<= 32) that fits three CTAs per SM; when that leaves fewer than 16 rows (wide inputs, large
  // outputs) the halo rows recomputed per tile dominate, so trade occupancy for a taller tile: two CTAs, then one.
  const size_t budgets[3] = {72 * 1024, 110 * 1024, 200 * 1024};
  for (int bi = 0; bi < 3; ++bi) {
    for (s.TY = 32; s.TY >= 1; s.TY /= 2) {
      int rows = 0;  // worst-case number of input rows one tile of TY output rows touches
      for (int y0 = 0; y0 < s.oh; y0 += s.TY) {
        const int y1 = (y0 + s.TY < s.oh ? y0 + s.TY : s.oh) - 1;
        const int r = tv.first[s.top + y1] + tv.count[s.top + y1] - tv.first[s.top + y0];
        rows = r > rows ? r : rows;
      }
      s.tile_rows = rows;
      s.smem = tables + static_cast<size_t>(kThreads / 32) * s.stage_bytes + static_cast<size_t>(rows + tv.kpad) * s.rowb + 16;  // zero-weight taps may run past the last row
      if (s.smem <= budgets[bi] || s.TY == 1) break;
    }
    if (s.smem <= budgets[bi] && (s.TY >= 16 || s.TY >= s.oh)) break;
  }
  if (s.smem > 200 * 1024) { set_last_error("resize %dx%d -> %dx%d needs %zu bytes of shared memory", H, W, s.rh, s.rw, s.smem); return JIMM_EINVAL; }
  auto ins = p->plans.emplace(std::make_pair(H, W), s);
  *out = &ins.first->second;
  return 0;
}

template <typename OUT>
int launch(const KernelArgs& a, const SizePlan& s, int B, cudaStream_t stream) {
  static DeviceOnce attr_set;
  if (attr_set.first()) {
    JIMM_CUDA_CHECK(cudaFuncSetAttribute(preprocess_kernel<OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const dim3 grid((s.oh + s.TY - 1) / s.TY, B);
  JIMM_CUDA_CHECK(launch_k(preprocess_kernel<OUT>, grid, dim3(kThreads), s.smem, stream, 1, false, a));
  note_launch();
  return 0;
}

}  // namespace

extern "C" {

int jimm_preproc_create(const jimm_preproc_config_t* cfg, int device, jimm_preproc_t** out) {
  JIMM_TRY(check_cfg(cfg));
  if (!out) { set_last_error("null output handle"); return JIMM_EINVAL; }
  JIMM_CUDA_CHECK(cudaSetDevice(device));
  jimm_preproc* p = new jimm_preproc();
  p->cfg = *cfg;
  p->device = device;
  // transformers rescale + normalize of one 8-bit sample: float32(float64(v) * factor), then (x - float32(mean)) / float32(std)
  std::vector<float> lut(3 * 256);
  for (int c = 0; c < 3; ++c)
    for (int v = 0; v < 256; ++v) {
      const float x = static_cast<float>(static_cast<double>(v) * cfg->rescale_factor);
      volatile float d = x - cfg->mean[c];  // volatile: keep the two roundings separate
      lut[c * 256 + v] = d / cfg->std[c];
    }
  void* d = nullptr;
  if (cudaMalloc(&d, lut.size() * sizeof(float)) != cudaSuccess) { delete p; set_last_error("cudaMalloc failed"); return JIMM_ENOMEM; }
  p->lut = static_cast<float*>(d);
  p->allocs.push_back(d);
  if (cudaMemcpy(d, lut.data(), lut.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d); delete p; set_last_error("cudaMemcpy failed"); return JIMM_ECUDA;
  }
  *out = p;
  return 0;
}

int jimm_preproc_output_size(const jimm_preproc_t* p, int H, int W, int* out_h, int* out_w) {
  if (!p || H <= 0 || W <= 0) { set_last_error("bad arguments"); return JIMM_EINVAL; }
  int rh, rw;
  resized_size(p->cfg, H, W, &rh, &rw);
  const int oh = p->cfg.crop_h ? p->cfg.crop_h : rh, ow = p->cfg.crop_w ? p->cfg.crop_w : rw;
  if (oh > rh || ow > rw) { set_last_error("centre crop %dx%d larger than the resized image %dx%d", oh, ow, rh, rw); return JIMM_EINVAL; }
  if (out_h) *out_h = oh;
  if (out_w) *out_w = ow;
  return 0;
}

int jimm_preproc_run(jimm_preproc_t* p, const uint8_t* img, int B, int H, int W, void* out, int out_dtype, void* stream) {
  if (!p || !img || !out) { set_last_error("null argument"); return JIMM_EINVAL; }
  if (B <= 0) return 0;
  if (out_dtype < JIMM_F32 || out_dtype > JIMM_BF16) { set_last_error("bad output dtype %d", out_dtype); return JIMM_EINVAL; }
  JIMM_CUDA_CHECK(cudaSetDevice(p->device));
  SizePlan* s = nullptr;
  JIMM_TRY(get_plan(p, H, W, &s));
  KernelArgs a;
  a.img = img;
  a.out = out;
  a.lut = p->lut;
  a.hfirst = s->h.first + s->left;
  a.hcount = s->h.count + s->left;
  a.hk = s->h.kk + static_cast<size_t>(s->left) * s->h.stride;
  a.vfirst = s->v.first + s->top;
  a.vcount = s->v.count + s->top;
  a.vk = s->v.kk + static_cast<size_t>(s->top) * s->v.ksize;
  a.H = H; a.W = W; a.oh = s->oh; a.ow = s->ow; a.hks = s->h.ksize; a.hstride = s->h.stride; a.vks = s->v.ksize;
  a.TY = s->TY; a.rowb = s->rowb; a.stage_bytes = s->stage_bytes;
  a.vec_ok = (reinterpret_cast<uintptr_t>(img) & 15) == 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the grid's y dimension is limited to 65535 images per launch
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = B - b0 < 65535 ? B - b0 : 65535;
    KernelArgs c = a;
    c.img = img + static_cast<size_t>(b0) * H * W * 3;
    const size_t out_off = static_cast<size_t>(b0) * s->oh * s->ow * 3;
    if (out_dtype == JIMM_F32) { c.out = static_cast<float*>(out) + out_off; JIMM_TRY(launch<float>(c, *s, nb, st)); }
    else if (out_dtype == JIMM_F16) { c.out = static_cast<__half*>(out) + out_off; JIMM_TRY(launch<__half>(c, *s, nb, st)); }
    else { c.out = static_cast<__nv_bfloat16*>(out) + out_off; JIMM_TRY(launch<__nv_bfloat16>(c, *s, nb, st)); }
  }
  return 0;
}

int jimm_preproc_destroy(jimm_preproc_t* p) {
  if (!p) return 0;
  cudaSetDevice(p->device);
  for (void* d : p->allocs) cudaFree(d);
  delete p;
  return 0;
}

// Host-only: the resampling tables, for the CPU test that pins them to the oracle's.
int jimm_k_resample_coeffs(int in_size, int out_size, int resample, int* ksize, int* first, int* count, int* kk, int kk_capacity) {
  if (in_size <= 0 || out_size <= 0 || (resample != 2 && resample != 3)) { set_last_error("bad arguments"); return JIMM_EINVAL; }
  ResampleTable t = make_table(in_size, out_size, resample);
  if (ksize) *ksize = t.ksize;
  if (first) for (int i = 0; i < out_size; ++i) first[i] = t.first[i];
  if (count) for (int i = 0; i < out_size; ++i) count[i] = t.count[i];
  if (kk) {
    if (kk_capacity < out_size * t.ksize) { set_last_error("kk buffer too small: %d < %d", kk_capacity, out_size * t.ksize); return JIMM_EINVAL; }
    for (int i = 0; i < out_size * t.ksize; ++i) kk[i] = t.kk[i];
  }
  return 0;
}

}  // extern "C"
