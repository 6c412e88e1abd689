// layout.cu -- module-boundary layout changes and weight preparation.
//   NCHW fp32 <-> NHWC operand-format views; OIHW fp32 -> [Cout][9][Cin] / [Cin][9][Cout] planes.
#include <cstdarg>

#include "common.cuh"
#include <atomic>
#include <vector>

namespace aide {

// ------------------------------------------------------------------ error plumbing (one TU owns it)
static thread_local char g_err[1024] = "";
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ------------------------------------------------------------------ fp16 saturation flag (see common.cuh::f16_split)
static std::vector<unsigned (*)(bool)>& f16_sat_readers() {
  static std::vector<unsigned (*)(bool)> v;
  return v;
}
void register_f16_sat_reader(unsigned (*fn)(bool)) { f16_sat_readers().push_back(fn); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return 2;
}

// ------------------------------------------------------------------ NCHW -> NHWC
// Tile transpose through shared memory: block handles 32 pixels x 32 channels of one image.
template <int FMT>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, void* p0, void* p1, int ctot, int coff,
                                    int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p_base = blockIdx.x * 32, c_base = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c_base + i, p = p_base + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? src[((size_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p_base + i, c = c_base + threadIdx.x;
    if (c < C && p < HW) st1<FMT>(p0, p1, ((size_t)n * HW + p) * ctot + coff + c, tile[threadIdx.x][i]);
  }
}

// Network inputs: [N,3,H,W] fp32 -> [N,H,W,3] fp32 (the only layout change on the step's input side).  The generic tile
// transpose above runs at 3/32 channel occupancy (155 GB/s measured); here a thread owns 4 consecutive pixels: three
// 128-bit plane reads, three 128-bit interleaved writes, both fully coalesced.
__global__ void nchw3_to_nhwc_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t HW4, size_t total4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / HW4, q = i - n * HW4;                 // 4-pixel group q of image n
    const float4* s = reinterpret_cast<const float4*>(src) + n * 3 * HW4 + q;
    const float4 r = s[0], g = s[HW4], b = s[2 * HW4];
    float4* d = reinterpret_cast<float4*>(dst) + (n * HW4 + q) * 3;
    d[0] = make_float4(r.x, g.x, b.x, r.y);
    d[1] = make_float4(g.y, b.y, r.z, g.z);
    d[2] = make_float4(b.z, r.w, g.w, b.w);
  }
}

template <int FMT>
__global__ void nhwc_to_nchw_kernel(const void* p0, const void* p1, int ctot, int coff, float* __restrict__ dst,
                                    int C, int HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p_base = blockIdx.x * 32, c_base = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p_base + i, c = c_base + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? ld1<FMT>(p0, p1, ((size_t)n * HW + p) * ctot + coff + c) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c_base + i, p = p_base + threadIdx.x;
    if (c < C && p < HW) dst[((size_t)n * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------ weight prep
// one thread per (co, ci, tap) element; writes both layouts.
template <int FMT>
__global__ void weight_prep_kernel(const float* __restrict__ w, int cout, int cin, void* f0, void* f1, void* d0,
                                   void* d1) {
  size_t total = (size_t)cout * cin * 9;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    // iterate in the fwd layout order so the fwd store is coalesced: i = (co*9 + tap)*cin + ci
    int ci = (int)(i % cin);
    int tap = (int)((i / cin) % 9);
    int co = (int)(i / ((size_t)cin * 9));
    float v = w[((size_t)co * cin + ci) * 9 + tap];
    if constexpr (FMT == AIDE_FMT_F16X2) v *= kF16WScale / kF16ActScale;   // st1 applies 2^8; weights carry 2^12
    if (f0) st1<FMT>(f0, f1, i, v);
    if (d0) st1<FMT>(d0, d1, ((size_t)ci * 9 + (8 - tap)) * cout + co, v);
  }
}

// Tiled version for cout % 32 == 0 and cin % 32 == 0 (every tensor-core layer): a block stages a [32 co][32 ci][9]
// tile through shared memory so that the OIHW read (288 contiguous floats per output channel) and both operand-plane
// writes (32 consecutive ci per (co, tap) row of the forward layout, 32 consecutive co per (ci, tap) row of the dgrad
// layout) are coalesced; the one-thread-per-element kernel above strides every access.
template <int FMT>
__global__ void __launch_bounds__(256) weight_prep_tiled_kernel(const float* __restrict__ w, int cout, int cin, void* f0,
                                                                void* f1, void* d0, void* d1) {
  constexpr int CO_STRIDE = 32 * 9 + 1;                 // +1: lanes across co hit different banks
  __shared__ float tile[32 * CO_STRIDE];
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  for (int i = threadIdx.x; i < 32 * 288; i += 256) {
    const int co = i / 288, r = i - co * 288;           // r = ci_local * 9 + tap: contiguous in OIHW
    tile[co * CO_STRIDE + r] = w[((size_t)(co0 + co) * cin + ci0) * 9 + r];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr float ws = FMT == AIDE_FMT_F16X2 ? kF16WScale / kF16ActScale : 1.0f;   // st1 applies 2^8; weights carry 2^12
  if (f0) {
    for (int row = wid; row < 32 * 9; row += 8) {       // row = co_local * 9 + tap, lanes = ci
      const int co = row / 9, tap = row - co * 9;
      const float v = tile[co * CO_STRIDE + lane * 9 + tap] * ws;
      st1<FMT>(f0, f1, ((size_t)(co0 + co) * 9 + tap) * cin + ci0 + lane, v);
    }
  }
  if (d0) {
    for (int row = wid; row < 32 * 9; row += 8) {       // row = ci_local * 9 + tap, lanes = co
      const int ci = row / 9, tap = row - ci * 9;
      const float v = tile[lane * CO_STRIDE + ci * 9 + tap] * ws;
      st1<FMT>(d0, d1, ((size_t)(ci0 + ci) * 9 + (8 - tap)) * cout + co0 + lane, v);
    }
  }
}

// All tensor-core layers of a network in ONE launch (a step used to spend 60 launches here): the per-layer arguments
// travel as a kernel parameter table, blockIdx.x is mapped to (layer, co tile, ci tile) through the tile prefix sums.
constexpr int kPrepMaxLayers = 40;
struct PrepTable {
  const float* w[kPrepMaxLayers];
  void* f0[kPrepMaxLayers];
  void* f1[kPrepMaxLayers];
  void* d0[kPrepMaxLayers];
  void* d1[kPrepMaxLayers];
  int cout[kPrepMaxLayers], cin[kPrepMaxLayers], tile_end[kPrepMaxLayers];
  int n;
};
template <int FMT>
__global__ void __launch_bounds__(256) weight_prep_batch_kernel(const __grid_constant__ PrepTable t) {
  constexpr int CO_STRIDE = 32 * 9 + 1;
  __shared__ float tile[32 * CO_STRIDE];
  int l = 0;
  while (l + 1 < t.n && (int)blockIdx.x >= t.tile_end[l]) ++l;
  const int local = (int)blockIdx.x - (l ? t.tile_end[l - 1] : 0);
  const int cout = t.cout[l], cin = t.cin[l];
  const int co0 = (local % (cout / 32)) * 32, ci0 = (local / (cout / 32)) * 32;
  const float* __restrict__ w = t.w[l];
  void *f0 = t.f0[l], *f1 = t.f1[l], *d0 = t.d0[l], *d1 = t.d1[l];
  for (int i = threadIdx.x; i < 32 * 288; i += 256) {
    const int co = i / 288, r = i - co * 288;
    tile[co * CO_STRIDE + r] = w[((size_t)(co0 + co) * cin + ci0) * 9 + r];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr float ws = FMT == AIDE_FMT_F16X2 ? kF16WScale / kF16ActScale : 1.0f;
  if (f0) {
    for (int row = wid; row < 32 * 9; row += 8) {
      const int co = row / 9, tap = row - co * 9;
      st1<FMT>(f0, f1, ((size_t)(co0 + co) * 9 + tap) * cin + ci0 + lane, tile[co * CO_STRIDE + lane * 9 + tap] * ws);
    }
  }
  if (d0) {
    for (int row = wid; row < 32 * 9; row += 8) {
      const int ci = row / 9, tap = row - ci * 9;
      st1<FMT>(d0, d1, ((size_t)(ci0 + ci) * 9 + (8 - tap)) * cout + co0 + lane, tile[lane * CO_STRIDE + ci * 9 + tap] * ws);
    }
  }
}

}  // namespace aide

using namespace aide;

extern "C" const char* aide_last_error(void) { return g_err; }
extern "C" int aide_version(void) { return 200; }
extern "C" int aide_f16_saturated(int reset) {
  unsigned any = 0;
  for (auto fn : f16_sat_readers()) any |= fn(reset != 0);
  return any ? 1 : 0;
}
extern "C" unsigned long long aide_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int aide_nchw_to_nhwc(int fmt, const float* src, void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff,
                                 int N, int C, int H, int W, void* stream) {
  AIDE_REQUIRE(src && dst_p0 && N > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad arguments");
  AIDE_REQUIRE(fmt_planes(fmt) == 1 || dst_p1, "nchw_to_nhwc: TF32X2 / F16X2 need two planes");
  if (fmt == AIDE_FMT_F32 && C == 3 && dst_ctot == 3 && dst_coff == 0 && ((size_t)H * W) % 4 == 0 &&
      ((uintptr_t)src | (uintptr_t)dst_p0) % 16 == 0) {
    const size_t HW4 = (size_t)H * W / 4, total4 = HW4 * N;
    long long blocks = (long long)((total4 + 255) / 256);
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    nchw3_to_nhwc_f32_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(src, reinterpret_cast<float*>(dst_p0), HW4, total4);
    AIDE_CHECK_LAUNCH();
    return 0;
  }
  dim3 grid(ceil_div((long long)H * W, 32), ceil_div(C, 32), N), block(32, 8);
  AIDE_DISPATCH_FMT(fmt, (nchw_to_nhwc_kernel<FMT><<<grid, block, 0, as_stream(stream)>>>(
                             src, dst_p0, dst_p1, dst_ctot, dst_coff, C, H * W)));
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_nhwc_to_nchw(int fmt, const void* src_p0, const void* src_p1, int src_ctot, int src_coff,
                                 float* dst, int N, int C, int H, int W, void* stream) {
  AIDE_REQUIRE(src_p0 && dst && N > 0 && C > 0 && H > 0 && W > 0, "nhwc_to_nchw: bad arguments");
  dim3 grid(ceil_div((long long)H * W, 32), ceil_div(C, 32), N), block(32, 8);
  AIDE_DISPATCH_FMT(fmt, (nhwc_to_nchw_kernel<FMT><<<grid, block, 0, as_stream(stream)>>>(
                             src_p0, src_p1, src_ctot, src_coff, dst, C, H * W)));
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_weight_prep(int fmt, const float* w_oihw, int cout, int cin, void* fwd_p0, void* fwd_p1,
                                void* dgrad_p0, void* dgrad_p1, void* stream) {
  AIDE_REQUIRE(w_oihw && cout > 0 && cin > 0 && (fwd_p0 || dgrad_p0), "weight_prep: bad arguments");
  AIDE_REQUIRE(fmt_planes(fmt) == 1 || ((!fwd_p0 || fwd_p1) && (!dgrad_p0 || dgrad_p1)),
               "weight_prep: TF32X2 / F16X2 need two planes");
  if (cout % 32 == 0 && cin % 32 == 0) {
    AIDE_DISPATCH_FMT(fmt, (weight_prep_tiled_kernel<FMT><<<dim3(cout / 32, cin / 32), 256, 0, as_stream(stream)>>>(
                               w_oihw, cout, cin, fwd_p0, fwd_p1, dgrad_p0, dgrad_p1)));
    AIDE_CHECK_LAUNCH();
    return 0;
  }
  size_t total = (size_t)cout * cin * 9;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  AIDE_DISPATCH_FMT(fmt, (weight_prep_kernel<FMT><<<blocks, 256, 0, as_stream(stream)>>>(
                             w_oihw, cout, cin, fwd_p0, fwd_p1, dgrad_p0, dgrad_p1)));
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_weight_prep_batch(int fmt, int n_layers, const float* const* w_oihw, const int* cout, const int* cin,
                                      void* const* fwd_p0, void* const* fwd_p1, void* const* dgrad_p0,
                                      void* const* dgrad_p1, void* stream) {
  AIDE_REQUIRE(n_layers >= 1 && w_oihw && cout && cin && fwd_p0 && fwd_p1 && dgrad_p0 && dgrad_p1,
               "weight_prep_batch: bad arguments");
  AIDE_REQUIRE(fmt == AIDE_FMT_TF32X2 || fmt == AIDE_FMT_BF16 || fmt == AIDE_FMT_F16X2,
               "weight_prep_batch: tensor-core operand formats only");
  for (int base = 0; base < n_layers; base += kPrepMaxLayers) {
    PrepTable t{};
    t.n = n_layers - base < kPrepMaxLayers ? n_layers - base : kPrepMaxLayers;
    int tiles = 0;
    for (int i = 0; i < t.n; ++i) {
      const int l = base + i;
      AIDE_REQUIRE(w_oihw[l] && cout[l] > 0 && cin[l] > 0 && cout[l] % 32 == 0 && cin[l] % 32 == 0,
                   "weight_prep_batch: layer %d needs cout %% 32 == 0 and cin %% 32 == 0", l);
      AIDE_REQUIRE((fwd_p0[l] || dgrad_p0[l]) && (fmt_planes(fmt) == 1 || ((!fwd_p0[l] || fwd_p1[l]) && (!dgrad_p0[l] || dgrad_p1[l]))),
                   "weight_prep_batch: layer %d: missing destination plane", l);
      t.w[i] = w_oihw[l]; t.f0[i] = fwd_p0[l]; t.f1[i] = fwd_p1[l]; t.d0[i] = dgrad_p0[l]; t.d1[i] = dgrad_p1[l];
      t.cout[i] = cout[l]; t.cin[i] = cin[l];
      tiles += (cout[l] / 32) * (cin[l] / 32);
      t.tile_end[i] = tiles;
    }
    AIDE_DISPATCH_FMT(fmt, (weight_prep_batch_kernel<FMT><<<tiles, 256, 0, as_stream(stream)>>>(t)));
    AIDE_CHECK_LAUNCH();
  }
  return 0;
}
