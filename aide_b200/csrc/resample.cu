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

// 16-bit operand planes (F16X2 / BF16): 8 channels = one 128-bit access per plane per thread
template <int FMT>
__device__ __forceinline__ void ld8(const void* p0, const void* p1, size_t e, float (&v)[8]) {
  if constexpr (FMT == AIDE_FMT_F16X2) {
    const uint4 rh = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p0) + e);
    const uint4 rl = *reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p1) + e);
    const __half2* h = reinterpret_cast<const __half2*>(&rh);
    const __half2* l = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = __half22float2(h[k]), b = __half22float2(l[k]);
      v[2 * k] = a.x + b.x;            // kept in the 2^8-prescaled domain: the interpolation is linear
      v[2 * k + 1] = a.y + b.y;
    }
  } else {
    const uint4 r = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p0) + e);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 a = __bfloat1622float2(h[k]);
      v[2 * k] = a.x;
      v[2 * k + 1] = a.y;
    }
  }
}
template <int FMT>
__device__ __forceinline__ void st8(void* p0, void* p1, size_t e, const float (&v)[8]) {
  if constexpr (FMT == AIDE_FMT_F16X2) {
    __align__(16) __half h[8];
    __align__(16) __half l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) f16_split(v[k], h[k], l[k]);      // already prescaled
    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p0) + e) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p1) + e) = *reinterpret_cast<const uint4*>(l);
  } else {
    __align__(16) __nv_bfloat16 h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = __float2bfloat16_rn(v[k]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p0) + e) = *reinterpret_cast<const uint4*>(h);
  }
}

// blockDim = (cx, ty): threadIdx.x walks the channels of a pixel in 8-channel (128-bit) vectors, whole OUTPUT rows are
// dealt to blocks and threadIdx.y walks along the row -- the vertical source rows / weights are block-uniform and the
// inner loop has no integer division.
template <int FMT>
__global__ void upsample2x_fwd16_kernel(CView src, View dst, int N, int h, int w, int C, float sh, float sw) {
  const int H = 2 * h, W = 2 * w;
  const int total_rows = N * H;
  for (int row = blockIdx.x; row < total_rows; row += gridDim.x) {
    const int n = row / H, oy = row - n * H;
    int y0, y1;
    float ly0, ly1;
    src_index(sh, oy, h, y0, y1, ly0, ly1);
    const size_t r0 = ((size_t)n * h + y0) * w, r1 = ((size_t)n * h + y1) * w;
    const size_t drow = (size_t)row * W;
    // two output pixels per iteration (16 independent 128-bit plane loads in flight per thread: the one-pixel loop ran
    // at ~0.5 of the copy peak, latency bound)
    for (int ox = threadIdx.y; ox < W; ox += 2 * blockDim.y) {
      const int oxb = ox + blockDim.y;
      const bool two = oxb < W;
      int x0, x1, u0, u1;
      float lx0, lx1, mx0, mx1;
      src_index(sw, ox, w, x0, x1, lx0, lx1);
      src_index(sw, two ? oxb : ox, w, u0, u1, mx0, mx1);
      const size_t s00 = (r0 + x0) * src.ctot + src.coff, s01 = (r0 + x1) * src.ctot + src.coff;
      const size_t s10 = (r1 + x0) * src.ctot + src.coff, s11 = (r1 + x1) * src.ctot + src.coff;
      const size_t t00 = (r0 + u0) * src.ctot + src.coff, t01 = (r0 + u1) * src.ctot + src.coff;
      const size_t t10 = (r1 + u0) * src.ctot + src.coff, t11 = (r1 + u1) * src.ctot + src.coff;
      const size_t d0 = (drow + ox) * dst.ctot + dst.coff, d1 = (drow + oxb) * dst.ctot + dst.coff;
      for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8) {
        float a[8], b[8], d[8], e[8], o[8], a2[8], b2[8], d2[8], e2[8];
        ld8<FMT>(src.p0, src.p1, s00 + c, a);
        ld8<FMT>(src.p0, src.p1, s01 + c, b);
        ld8<FMT>(src.p0, src.p1, s10 + c, d);
        ld8<FMT>(src.p0, src.p1, s11 + c, e);
        ld8<FMT>(src.p0, src.p1, t00 + c, a2);
        ld8<FMT>(src.p0, src.p1, t01 + c, b2);
        ld8<FMT>(src.p0, src.p1, t10 + c, d2);
        ld8<FMT>(src.p0, src.p1, t11 + c, e2);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // same operation order as the 4-channel kernel: along W first, then along H
          const float t0 = lx0 * a[k] + lx1 * b[k], t1 = lx0 * d[k] + lx1 * e[k];
          o[k] = ly0 * t0 + ly1 * t1;
        }
        st8<FMT>(dst.p0, dst.p1, d0 + c, o);
        if (two) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float t0 = mx0 * a2[k] + mx1 * b2[k], t1 = mx0 * d2[k] + mx1 * e2[k];
            o[k] = ly0 * t0 + ly1 * t1;
          }
          st8<FMT>(dst.p0, dst.p1, d1 + c, o);
        }
      }
    }
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
    // horizontal weights of the candidate columns once per pixel (not once per candidate row)
    constexpr int kMaxCand = 10;    // (ix +- 1) / scale spans <= ~4 output columns; the floor / ceil / +-1 margins widen the window to <= 9
    float wxs[kMaxCand];
    const int ncx = min(ox_hi - ox_lo + 1, kMaxCand);
#pragma unroll
    for (int j = 0; j < kMaxCand; ++j) {
      wxs[j] = 0.f;
      if (j < ncx) {
        int x0, x1;
        float lx0, lx1;
        src_index(sw, ox_lo + j, w, x0, x1, lx0, lx1);
        wxs[j] = (x0 == ix ? lx0 : 0.f) + (x1 == ix ? lx1 : 0.f);
      }
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1;
      float ly0, ly1;
      src_index(sh, oy, h, y0, y1, ly0, ly1);
      float wy = (y0 == iy ? ly0 : 0.f) + (y1 == iy ? ly1 : 0.f);
      if (wy == 0.f) continue;
      const float* rowp = dhi + (((size_t)n * H + oy) * W + ox_lo) * ctot + coff + c;
#pragma unroll
      for (int j = 0; j < kMaxCand; ++j) {
        const float wx = wxs[j];
        if (j >= ncx || wx == 0.f) continue;
        float4 v = *reinterpret_cast<const float4*>(rowp + (size_t)j * ctot);
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

// ---------------------------------------------------------------- zero-insert x2 (ConvTranspose2d(k=2, s=2) front end)
// nn.ConvTranspose2d(Cin, Cout, kernel_size=2, stride=2) (netblocks.py:12, learned_bilinear=True) equals a 3x3 / pad 1
// convolution of the zero-inserted input X'[2h,2w] = x[h,w] (0 elsewhere) with K[co,ci,1-a,1-b] = W[ci,co,a,b]: the
// engine reuses its conv3x3 unit (forward, dgrad, wgrad) and only needs this copy.  Raw 16-byte vectors per plane.
__global__ void zero_insert2x_kernel(const uint4* __restrict__ s0, const uint4* __restrict__ s1, int s_ctot_v, int s_coff_v,
                                     uint4* __restrict__ d0, uint4* __restrict__ d1, int d_ctot_v, int d_coff_v, int N,
                                     int h, int w, int Cv) {
  const int H = 2 * h, W = 2 * w;
  const size_t total = (size_t)N * H * W * Cv;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cv);
    const size_t pix = i / Cv;
    const int ox = (int)(pix % W), oy = (int)((pix / W) % H);
    const size_t n = pix / ((size_t)W * H);
    uint4 a = make_uint4(0u, 0u, 0u, 0u), b = a;
    if (((ox | oy) & 1) == 0) {
      const size_t sp = ((n * h + (oy >> 1)) * w + (ox >> 1)) * s_ctot_v + s_coff_v + c;
      a = s0[sp];
      if (s1) b = s1[sp];
    }
    const size_t dp = pix * d_ctot_v + d_coff_v + c;
    d0[dp] = a;
    if (d1) d1[dp] = b;
  }
}

// transpose: d_lo[n,h,w,c] = d_hi[n,2h,2w,c]
__global__ void zero_insert2x_bwd_kernel(const float* __restrict__ dhi, int ctot, int coff, float* __restrict__ dlo, int N,
                                         int h, int w, int C) {
  const int C4 = C >> 2;
  const size_t total = (size_t)N * h * w * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const size_t pix = i / C4;
    const int ix = (int)(pix % w), iy = (int)((pix / w) % h);
    const size_t n = pix / ((size_t)w * h);
    *reinterpret_cast<float4*>(dlo + pix * C + c) =
        *reinterpret_cast<const float4*>(dhi + ((n * 2 * h + 2 * iy) * (size_t)(2 * w) + 2 * ix) * ctot + coff + c);
  }
}

}  // namespace aide

using namespace aide;

extern "C" int aide_zero_insert2x_fwd(int fmt, const void* src_p0, const void* src_p1, int src_ctot, int src_coff,
                                      void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff, int N, int h, int w, int C,
                                      void* stream) {
  AIDE_REQUIRE(fmt_valid(fmt) && src_p0 && dst_p0 && N > 0 && h > 0 && w > 0, "zero_insert2x_fwd: bad arguments");
  const int vec = 16 / fmt_elem_bytes(fmt);            // channels per 16-byte vector
  AIDE_REQUIRE(C % vec == 0 && src_ctot % vec == 0 && src_coff % vec == 0 && dst_ctot % vec == 0 && dst_coff % vec == 0,
               "zero_insert2x_fwd: channel counts / offsets must be multiples of %d", vec);
  AIDE_REQUIRE(fmt_planes(fmt) == 1 || (src_p1 && dst_p1), "zero_insert2x_fwd: two-plane format needs both planes");
  const size_t total = (size_t)N * 4 * h * w * (C / vec);
  long long blocks = (long long)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  zero_insert2x_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const uint4*>(src_p0), reinterpret_cast<const uint4*>(fmt_planes(fmt) == 2 ? src_p1 : nullptr),
      src_ctot / vec, src_coff / vec, reinterpret_cast<uint4*>(dst_p0),
      reinterpret_cast<uint4*>(fmt_planes(fmt) == 2 ? dst_p1 : nullptr), dst_ctot / vec, dst_coff / vec, N, h, w, C / vec);
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_zero_insert2x_bwd(const float* dhi, int dhi_ctot, int dhi_coff, float* dlo, int N, int h, int w,
                                      int C, void* stream) {
  AIDE_REQUIRE(dhi && dlo && C % 4 == 0 && dhi_coff % 4 == 0 && dhi_ctot % 4 == 0, "zero_insert2x_bwd: bad arguments");
  const size_t total = (size_t)N * h * w * (C / 4);
  long long blocks = (long long)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  zero_insert2x_bwd_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(dhi, dhi_ctot, dhi_coff, dlo, N, h, w, C);
  AIDE_CHECK_LAUNCH();
  return 0;
}

static inline float ac_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f; }

extern "C" int aide_upsample2x_fwd(int fmt, const void* src_p0, const void* src_p1, int src_ctot, int src_coff,
                                   void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff, int N, int h, int w, int C,
                                   void* stream) {
  AIDE_REQUIRE(src_p0 && dst_p0 && C % 4 == 0 && src_coff % 4 == 0 && dst_coff % 4 == 0 && N > 0 && h > 0 && w > 0,
               "upsample2x_fwd: bad arguments");
  CView src{src_p0, src_p1, src_ctot, src_coff};
  View dst{dst_p0, dst_p1, dst_ctot, dst_coff};
  if ((fmt == AIDE_FMT_F16X2 || fmt == AIDE_FMT_BF16) && C % 8 == 0 && src_coff % 8 == 0 && dst_coff % 8 == 0 &&
      src_ctot % 8 == 0 && dst_ctot % 8 == 0) {
    int cx = C / 8 < 32 ? C / 8 : 32;                   // threads across the channels of one pixel
    while (cx & (cx - 1)) cx &= cx - 1;                 // power of two
    const dim3 block(cx, 256 / cx);
    long long blocks = (long long)N * 2 * h;            // one output row per block iteration
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    const float sh = ac_scale(h, 2 * h), sw = ac_scale(w, 2 * w);
    if (fmt == AIDE_FMT_F16X2)
      upsample2x_fwd16_kernel<AIDE_FMT_F16X2><<<(int)blocks, block, 0, as_stream(stream)>>>(src, dst, N, h, w, C, sh, sw);
    else
      upsample2x_fwd16_kernel<AIDE_FMT_BF16><<<(int)blocks, block, 0, as_stream(stream)>>>(src, dst, N, h, w, C, sh, sw);
    AIDE_CHECK_LAUNCH();
    return 0;
  }
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
