// head.cu -- last_conv1: 1x1 convolution C -> K (+bias), forward and backward
// (reference: models_twomodalinputs/fuseunet.py:41,89; models_singlemodalinput/UNet.py:150,164).
// K is tiny (num_classes = 2), so this is an HBM-bound per-pixel dot product on CUDA cores:
// reads the NHWC activation once (128-bit loads), writes NCHW fp32 logits.
#include "common.cuh"

namespace aide {

constexpr int kMaxK = 8;

// LP lanes cooperate on one pixel (LP = power of two <= 32); 32/LP pixels per warp per iteration.
template <int FMT>
__global__ void conv1x1_fwd_kernel(CView x, int C, const float* __restrict__ w, const float* __restrict__ bias,
                                   float* __restrict__ out, int K, size_t npix, int HW, int LP) {
  const int lane = threadIdx.x & 31;
  const int sub = lane / LP, l = lane % LP, PW = 32 / LP;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int C4 = C >> 2;
  for (size_t base = warp * PW; base < npix; base += nwarps * PW) {
    size_t pix = base + sub;
    bool valid = pix < npix;
    float acc[kMaxK];
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) acc[k] = 0.f;
    if (valid) {
      for (int c4 = l; c4 < C4; c4 += LP) {
        float4 v = ld4<FMT>(x.p0, x.p1, pix * x.ctot + x.coff + c4 * 4);
#pragma unroll
        for (int k = 0; k < kMaxK; ++k) {
          if (k < K) {
            float4 wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)k * C + c4 * 4));
            acc[k] += v.x * wv.x + v.y * wv.y + v.z * wv.z + v.w * wv.w;
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxK; ++k) {
      if (k < K) {
        for (int o = LP >> 1; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
      }
    }
    if (valid && l == 0) {
      size_t n = pix / HW, hw = pix % HW;
#pragma unroll
      for (int k = 0; k < kMaxK; ++k)
        if (k < K) out[(n * K + k) * HW + hw] = acc[k] + (bias ? bias[k] : 0.f);
    }
  }
}

// 16-bit operand planes (F16X2 / BF16), K == 2, C == 8 * LP with LP a power of two <= 32: each lane owns 8 channels
// (one 128-bit load per plane), keeps its 2 x 8 weights in registers, LP lanes finish a pixel with log2(LP) shuffles.
template <int FMT, int LP>
__global__ void conv1x1_fwd16_kernel(CView x, const float* __restrict__ w, const float* __restrict__ bias,
                                     float* __restrict__ out, size_t npix, int HW) {
  constexpr int C = 8 * LP, PW = 32 / LP;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LP, l = lane % LP;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  float w0[8], w1[8];
  constexpr float inv = FMT == AIDE_FMT_F16X2 ? 1.0f / kF16ActScale : 1.0f;      // fold the activation prescale into w
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    w0[k] = __ldg(w + l * 8 + k) * inv;
    w1[k] = __ldg(w + C + l * 8 + k) * inv;
  }
  const float b0 = bias ? bias[0] : 0.f, b1 = bias ? bias[1] : 0.f;
  for (size_t base = warp * PW; base < npix; base += nwarps * PW) {
    const size_t pix = base + sub;
    const bool valid = pix < npix;
    float a0 = 0.f, a1 = 0.f;
    if (valid) {
      const size_t e = pix * x.ctot + x.coff + l * 8;
      float v[8];
      if constexpr (FMT == AIDE_FMT_F16X2) {
        const uint4 rh = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(x.p0) + e);
        const uint4 rl = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(x.p1) + e);
        const __half2* h = reinterpret_cast<const __half2*>(&rh);
        const __half2* lo = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 p = __half22float2(h[k]), q = __half22float2(lo[k]);
          v[2 * k] = p.x + q.x;
          v[2 * k + 1] = p.y + q.y;
        }
      } else {
        const uint4 r = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(x.p0) + e);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 p = __bfloat1622float2(h[k]);
          v[2 * k] = p.x;
          v[2 * k + 1] = p.y;
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        a0 += v[k] * w0[k];
        a1 += v[k] * w1[k];
      }
    }
#pragma unroll
    for (int o = LP >> 1; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o);
      a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    }
    if (valid && l == 0) {
      const size_t n = pix / HW, hw = pix % HW;
      out[(n * 2 + 0) * HW + hw] = a0 + b0;
      out[(n * 2 + 1) * HW + hw] = a1 + b1;
    }
  }
}

// blockDim = (cx, ty), grid = (rows, cgroups); each thread owns 4 channels.
template <int FMT, int KB>            // KB: compile-time bound of the class count (2 for every AIDE head, else KB)
__global__ void __launch_bounds__(256, 2) conv1x1_bwd_kernel(CView x, int C, const float* __restrict__ w, const float* __restrict__ dl, int K,
                                   size_t npix, int HW, float* __restrict__ dx, float* __restrict__ partial) {
  extern __shared__ float smem[];  // [ty][cx][K*4 + K]
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  const bool cvalid = c < C;
  const int stride = K * 4 + K;
  float aw[KB][4], ab[KB];
  float4 wv[KB];
#pragma unroll
  for (int k = 0; k < KB; ++k) {
    aw[k][0] = aw[k][1] = aw[k][2] = aw[k][3] = 0.f;
    ab[k] = 0.f;
    wv[k] = (cvalid && k < K) ? *reinterpret_cast<const float4*>(w + (size_t)k * C + c) : make_float4(0, 0, 0, 0);
  }
  if (cvalid) {
    // two pixels per iteration (independent loads in flight; 32-bit index arithmetic -- the 64-bit division per pixel
    // and the single pixel in flight held this kernel at 0.19 of the copy peak); same accumulation order
    const unsigned int np = (unsigned int)npix, step = gridDim.x * blockDim.y;
    for (unsigned int p = blockIdx.x * blockDim.y + threadIdx.y; p < np; p += 2 * step) {
      const unsigned int q = p + step;
      const bool two = q < np;
      const unsigned int n0 = p / (unsigned int)HW, hw0 = p - n0 * (unsigned int)HW;
      const unsigned int n1 = two ? q / (unsigned int)HW : n0, hw1 = two ? q - n1 * (unsigned int)HW : hw0;
      const float4 xv0 = ld4<FMT>(x.p0, x.p1, (size_t)p * x.ctot + x.coff + c);
      const float4 xv1 = two ? ld4<FMT>(x.p0, x.p1, (size_t)q * x.ctot + x.coff + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      float g0[KB], g1[KB];
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        g0[k] = k < K ? __ldg(dl + ((size_t)n0 * K + k) * HW + hw0) : 0.f;
        g1[k] = (k < K && two) ? __ldg(dl + ((size_t)n1 * K + k) * HW + hw1) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        const float4 xv = u ? xv1 : xv0;
        float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < KB; ++k) {
          if (k < K) {
            const float g = u ? g1[k] : g0[k];
            d.x += g * wv[k].x; d.y += g * wv[k].y; d.z += g * wv[k].z; d.w += g * wv[k].w;
            aw[k][0] += g * xv.x; aw[k][1] += g * xv.y; aw[k][2] += g * xv.z; aw[k][3] += g * xv.w;
            ab[k] += g;
          }
        }
        *reinterpret_cast<float4*>(dx + (size_t)(u ? q : p) * C + c) = d;
      }
    }
  }
  float* row = smem + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * stride;
#pragma unroll
  for (int k = 0; k < KB; ++k) {
    if (k < K) {
      row[k * 4 + 0] = aw[k][0]; row[k * 4 + 1] = aw[k][1]; row[k * 4 + 2] = aw[k][2]; row[k * 4 + 3] = aw[k][3];
      row[K * 4 + k] = ab[k];
    }
  }
  __syncthreads();
  if (threadIdx.y == 0 && cvalid) {
    float* out = partial + (size_t)blockIdx.x * ((size_t)K * C + K);
    for (int j = 0; j < stride; ++j) {
      float a = 0.f;
      for (int t = 0; t < (int)blockDim.y; ++t) a += smem[((size_t)t * blockDim.x + threadIdx.x) * stride + j];
      if (j < K * 4) out[(size_t)(j >> 2) * C + c + (j & 3)] = a;
      else if (c == 0) out[(size_t)K * C + (j - K * 4)] = a;
    }
  }
}

struct HeadGeom { int cx, ty, cgroups, rows; };
static HeadGeom head_geom(int N, int H, int W, int C) {
  HeadGeom g;
  int c4 = C / 4;
  g.cx = c4 < 64 ? c4 : 64;
  g.ty = 256 / g.cx;
  g.cgroups = ceil_div(c4, g.cx);
  long long npix = (long long)N * H * W;
  long long want = (npix + g.ty * 8 - 1) / (g.ty * 8);
  long long cap = (long long)kNumSMs * 8 / g.cgroups;
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  g.rows = (int)want;
  return g;
}

}  // namespace aide

using namespace aide;

extern "C" int aide_conv1x1_fwd(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int C,
                                const float* w, const float* bias, float* logits_nchw, int K, int N, int H, int W,
                                void* stream) {
  AIDE_REQUIRE(x_p0 && w && logits_nchw && C % 4 == 0 && x_coff % 4 == 0 && K >= 1 && K <= kMaxK,
               "conv1x1_fwd: bad arguments (C%%4==0, 1<=K<=%d)", kMaxK);
  CView x{x_p0, x_p1, x_ctot, x_coff};
  size_t npix = (size_t)N * H * W;
  if ((fmt == AIDE_FMT_F16X2 || fmt == AIDE_FMT_BF16) && K == 2 && C == 64 && x_ctot % 8 == 0 && x_coff % 8 == 0) {
    size_t warps16 = (npix + 3) / 4;
    long long blocks16 = (long long)((warps16 + 7) / 8);
    if (blocks16 > kNumSMs * 16) blocks16 = kNumSMs * 16;
    if (fmt == AIDE_FMT_F16X2)
      conv1x1_fwd16_kernel<AIDE_FMT_F16X2, 8><<<(int)blocks16, 256, 0, as_stream(stream)>>>(x, w, bias, logits_nchw, npix, H * W);
    else
      conv1x1_fwd16_kernel<AIDE_FMT_BF16, 8><<<(int)blocks16, 256, 0, as_stream(stream)>>>(x, w, bias, logits_nchw, npix, H * W);
    AIDE_CHECK_LAUNCH();
    return 0;
  }
  int LP = 1;
  while (LP < C / 4 && LP < 32) LP <<= 1;
  size_t warps = (npix + (32 / LP) - 1) / (32 / LP);
  int blocks = (int)((warps + 7) / 8);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (blocks < 1) blocks = 1;
  AIDE_DISPATCH_FMT(fmt, (conv1x1_fwd_kernel<FMT><<<blocks, 256, 0, as_stream(stream)>>>(x, C, w, bias, logits_nchw, K,
                                                                                        npix, H * W, LP)));
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_conv1x1_bwd_rows(int N, int H, int W, int C) { return head_geom(N, H, W, C).rows; }

extern "C" int aide_conv1x1_bwd(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int C,
                                const float* w, const float* dlogits_nchw, int K, int N, int H, int W, float* dx,
                                float* dw_db, float* partial, void* stream) {
  AIDE_REQUIRE(x_p0 && w && dlogits_nchw && dx && dw_db && partial && C % 4 == 0 && x_coff % 4 == 0 && K >= 1 &&
                   K <= kMaxK,
               "conv1x1_bwd: bad arguments");
  CView x{x_p0, x_p1, x_ctot, x_coff};
  HeadGeom g = head_geom(N, H, W, C);
  dim3 block(g.cx, g.ty), grid(g.rows, g.cgroups);
  size_t smem = (size_t)g.cx * g.ty * (K * 5) * sizeof(float);
  size_t npix = (size_t)N * H * W;
  AIDE_REQUIRE(npix < (1ull << 31), "conv1x1_bwd: more than 2^31 pixels");
  if (K <= 2) {
    AIDE_DISPATCH_FMT(fmt, (conv1x1_bwd_kernel<FMT, 2><<<grid, block, smem, as_stream(stream)>>>(
                               x, C, w, dlogits_nchw, K, npix, H * W, dx, partial)));
  } else {
    AIDE_DISPATCH_FMT(fmt, (conv1x1_bwd_kernel<FMT, kMaxK><<<grid, block, smem, as_stream(stream)>>>(
                               x, C, w, dlogits_nchw, K, npix, H * W, dx, partial)));
  }
  AIDE_CHECK_LAUNCH();
  int ld = K * C + K;
  return launch_reduce_rows(partial, g.rows, ld, ld, dw_db, as_stream(stream));
}
