// resample.cu -- nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) forward / backward
// on NHWC data (reference: models_twomodalinputs/netblocks.py:16).  HBM-bound gather kernels,
// 4 channels (128-bit) per thread, channel fastest.
//
// Index math follows ATen (UpSample.h): scale = (in-1)/(out-1) in fp32, src = scale*dst,
// i0 = (int)src, i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1;  value interpolated along W
// first, then along H.
#include "common.cuh"

namespace aide {

__device__ __forceinline__ void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * (float)dst;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = fminf(fmaxf(src - (float)i0, 0.f), 1.f);
  l0 = 1.f - l1;
}

__device__ __forceinline__ float4 lerp4(float a, float4 x, float b, float4 y) {
  return make_float4(a * x.x + b * y.x, a * x.y + b * y.y, a * x.z + b * y.z, a * x.w + b * y.w);
}

template <int FMT>
__global__ void upsample2x_fwd_kernel(CView src, View dst, int N, int h, int w, int C, float sh, float sw) {
  const int C4 = C >> 2, H = 2 * h, W = 2 * w;
  size_t total = (size_t)N * H * W * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C4) * 4;
    size_t pix = i / C4;
    int ox = (int)(pix % W);
    int oy = (int)((pix / W) % H);
    int n = (int)(pix / ((size_t)W * H));
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    src_index(sh, oy, h, y0, y1, ly0, ly1);
    src_index(sw, ox, w, x0, x1, lx0, lx1);
    size_t r0 = ((size_t)n * h + y0) * w, r1 = ((size_t)n * h + y1) * w;
    float4 a = ld4<FMT>(src.p0, src.p1, (r0 + x0) * src.ctot + src.coff + c);
    float4 b = ld4<FMT>(src.p0, src.p1, (r0 + x1) * src.ctot + src.coff + c);
    float4 d = ld4<FMT>(src.p0, src.p1, (r1 + x0) * src.ctot + src.coff + c);
    float4 e = ld4<FMT>(src.p0, src.p1, (r1 + x1) * src.ctot + src.coff + c);
    float4 t0 = lerp4(lx0, a, lx1, b), t1 = lerp4(lx0, d, lx1, e);
    st4<FMT>(dst.p0, dst.p1, pix * dst.ctot + dst.coff + c, lerp4(ly0, t0, ly1, t1));
  }
}

// Gather form of the transpose: every low-res pixel sums the (<= ~5x5) high-res pixels that read it.
__global__ void upsample2x_bwd_kernel(const float* __restrict__ dhi, int ctot, int coff, float* __restrict__ dlo,
                                      int N, int h, int w, int C, float sh, float sw) {
  const int C4 = C >> 2, H = 2 * h, W = 2 * w;
  size_t total = (size_t)N * h * w * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C4) * 4;
    size_t pix = i / C4;
    int ix = (int)(pix % w);
    int iy = (int)((pix / w) % h);
    int n = (int)(pix / ((size_t)w * h));
    // candidate output rows: src in (iy-1, iy+1)  ->  dst in ((iy-1)/s, (iy+1)/s)
    int oy_lo = 0, oy_hi = H - 1, ox_lo = 0, ox_hi = W - 1;
    if (sh > 0.f) {
      oy_lo = max(0, (int)floorf((float)(iy - 1) / sh) - 1);
      oy_hi = min(H - 1, (int)ceilf((float)(iy + 1) / sh) + 1);
    }
    if (sw > 0.f) {
      ox_lo = max(0, (int)floorf((float)(ix - 1) / sw) - 1);
      ox_hi = min(W - 1, (int)ceilf((float)(ix + 1) / sw) + 1);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1;
      float ly0, ly1;
      src_index(sh, oy, h, y0, y1, ly0, ly1);
      float wy = (y0 == iy ? ly0 : 0.f) + (y1 == iy ? ly1 : 0.f);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1;
        float lx0, lx1;
        src_index(sw, ox, w, x0, x1, lx0, lx1);
        float wx = (x0 == ix ? lx0 : 0.f) + (x1 == ix ? lx1 : 0.f);
        if (wx == 0.f) continue;
        float4 v = *reinterpret_cast<const float4*>(dhi + (((size_t)n * H + oy) * W + ox) * ctot + coff + c);
        float ww = wy * wx;
        acc.x += ww * v.x;
        acc.y += ww * v.y;
        acc.z += ww * v.z;
        acc.w += ww * v.w;
      }
    }
    *reinterpret_cast<float4*>(dlo + pix * C + c) = acc;
  }
}

}  // namespace aide

using namespace aide;

static inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

extern "C" int aide_upsample2x_fwd(int fmt, const void* src_p0, const void* src_p1, int src_ctot, int src_coff,
                                   void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff, int N, int h, int w, int C,
                                   void* stream) {
  AIDE_REQUIRE(src_p0 && dst_p0 && C % 4 == 0 && src_coff % 4 == 0 && dst_coff % 4 == 0 && N > 0 && h > 0 && w > 0,
               "upsample2x_fwd: bad arguments");
  CView src{src_p0, src_p1, src_ctot, src_coff};
  View dst{dst_p0, dst_p1, dst_ctot, dst_coff};
  size_t total = (size_t)N * 4 * h * w * (C / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  AIDE_DISPATCH_FMT(fmt, (upsample2x_fwd_kernel<FMT><<<blocks, 256, 0, as_stream(stream)>>>(
                             src, dst, N, h, w, C, ac_scale(h, 2 * h), ac_scale(w, 2 * w))));
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_upsample2x_bwd(const float* dhi, int dhi_ctot, int dhi_coff, float* dlo, int N, int h, int w,
                                   int C, void* stream) {
  AIDE_REQUIRE(dhi && dlo && C % 4 == 0 && dhi_coff % 4 == 0 && dhi_ctot % 4 == 0, "upsample2x_bwd: bad arguments");
  size_t total = (size_t)N * h * w * (C / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  upsample2x_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(dhi, dhi_ctot, dhi_coff, dlo, N, h, w, C,
                                                               ac_scale(h, 2 * h), ac_scale(w, 2 * w));
  AIDE_CHECK_LAUNCH();
  return 0;
}
