// resample.cu -- nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) forward / backward
// on NHWC data (reference: models_twomodalinputs/netblocks.py:16).  HBM-bound gather kernels,
// 4 channels (128-bit) per thread, channel fastest.
//
// Index math follows ATen (UpSample.h): scale = (in-1)/(out-1) in fp32, src = scale*dst,
// i0 = (int)src, i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1;  value interpolated along W
// first, then along H.
#include "common.cuh"

#include <cstdlib>

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

// blockDim = (cx, ty): threadIdx.x walks the channels of a pixel in 8-channel (128-bit) vectors; a block owns a PAIR of
// output rows (2i, 2i+1) and threadIdx.y walks the output column pairs (2j, 2j+1).  With align_corners the four outputs
// of such a 2x2 block read only 3 x 3 distinct source pixels (rows {y0, y1} of the upper output and y1 of the lower one;
// the lower output's first row is one of the upper one's two) -- 18 plane loads per 4 outputs instead of 32, which is
// what bounds this kernel (L1 / L2 request rate, not HBM: 0.50 of the copy peak before).  Per output the operations and
// their order are those of the 4-channel kernel: along W first, then along H.
template <int FMT>
__global__ void __launch_bounds__(256, 2) upsample2x_fwd16_kernel(CView src, View dst, int N, int h, int w, int C, float sh, float sw) {
  const int H = 2 * h, W = 2 * w;
  const int total_pairs = N * h;
  for (int rp = blockIdx.x; rp < total_pairs; rp += gridDim.x) {
    const int n = rp / h, oya = 2 * (rp - n * h);
    int ya0, ya1, yb0, yb1;
    float lya0, lya1, lyb0, lyb1;
    src_index(sh, oya, h, ya0, ya1, lya0, lya1);
    src_index(sh, oya + 1, h, yb0, yb1, lyb0, lyb1);
    const bool ib0 = yb0 != ya0;                               // the lower output's first source row is ya1 (else ya0)
    const size_t r0 = ((size_t)n * h + ya0) * w, r1 = ((size_t)n * h + ya1) * w, r2 = ((size_t)n * h + yb1) * w;
    const size_t drow_a = ((size_t)n * H + oya) * W, drow_b = drow_a + W;
    for (int j = threadIdx.y; j < w; j += blockDim.y) {
      const int oxa = 2 * j;
      int xa0, xa1, xb0, xb1;
      float lxa0, lxa1, lxb0, lxb1;
      src_index(sw, oxa, w, xa0, xa1, lxa0, lxa1);
      src_index(sw, oxa + 1, w, xb0, xb1, lxb0, lxb1);
      const bool jb0 = xb0 != xa0;
      const size_t c0 = (size_t)xa0 * src.ctot + src.coff, c1 = (size_t)xa1 * src.ctot + src.coff,
                   c2 = (size_t)xb1 * src.ctot + src.coff;
      const size_t da = (drow_a + oxa) * dst.ctot + dst.coff, db = (drow_b + oxa) * dst.ctot + dst.coff;
      for (int c = threadIdx.x * 8; c < C; c += blockDim.x * 8) {
        float v[3][3][8];
        ld8<FMT>(src.p0, src.p1, r0 * src.ctot + c0 + c, v[0][0]);
        ld8<FMT>(src.p0, src.p1, r0 * src.ctot + c1 + c, v[0][1]);
        ld8<FMT>(src.p0, src.p1, r0 * src.ctot + c2 + c, v[0][2]);
        ld8<FMT>(src.p0, src.p1, r1 * src.ctot + c0 + c, v[1][0]);
        ld8<FMT>(src.p0, src.p1, r1 * src.ctot + c1 + c, v[1][1]);
        ld8<FMT>(src.p0, src.p1, r1 * src.ctot + c2 + c, v[1][2]);
        ld8<FMT>(src.p0, src.p1, r2 * src.ctot + c0 + c, v[2][0]);
        ld8<FMT>(src.p0, src.p1, r2 * src.ctot + c1 + c, v[2][1]);
        ld8<FMT>(src.p0, src.p1, r2 * src.ctot + c2 + c, v[2][2]);
        float oaa[8], oab[8], oba[8], obb[8];                  // [output row a/b][output column a/b]
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          float ha[3], hb[3];                                  // the three source rows, interpolated along W
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            ha[r] = lxa0 * v[r][0][k] + lxa1 * v[r][1][k];
            hb[r] = lxb0 * (jb0 ? v[r][1][k] : v[r][0][k]) + lxb1 * v[r][2][k];
          }
          oaa[k] = lya0 * ha[0] + lya1 * ha[1];
          oab[k] = lya0 * hb[0] + lya1 * hb[1];
          oba[k] = lyb0 * (ib0 ? ha[1] : ha[0]) + lyb1 * ha[2];
          obb[k] = lyb0 * (ib0 ? hb[1] : hb[0]) + lyb1 * hb[2];
        }
        st8<FMT>(dst.p0, dst.p1, da + c, oaa);
        st8<FMT>(dst.p0, dst.p1, da + dst.ctot + c, oab);
        st8<FMT>(dst.p0, dst.p1, db + c, oba);
        st8<FMT>(dst.p0, dst.p1, db + dst.ctot + c, obb);
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

// Row form of the same gather (the fast path): a block owns one low-res row, so the <= 4 contributing output rows and
// their vertical weights are block-uniform; threadIdx.y walks the pixels of the row, threadIdx.x the channels in 128-bit
// vectors; the <= 4 x 4 taps are independent loads.  Same candidates, same weights and the same summation order (output
// rows ascending, columns ascending inside a row) as upsample2x_bwd_kernel: bit-identical results, ~3x its bandwidth.
__global__ void __launch_bounds__(256, 2) upsample2x_bwd_rows_kernel(const float* __restrict__ dhi, int ctot, int coff, float* __restrict__ dlo,
                                           int N, int h, int w, int C, float sh, float sw) {
  const int H = 2 * h, W = 2 * w;
  const int total_rows = N * h;
  for (int row = blockIdx.x; row < total_rows; row += gridDim.x) {
    const int n = row / h, iy = row - n * h;
    int oy_lo = 0, oy_hi = H - 1;
    if (sh > 0.f) {
      oy_lo = max(0, (int)floorf((float)(iy - 1) / sh) - 1);
      oy_hi = min(H - 1, (int)ceilf((float)(iy + 1) / sh) + 1);
    }
    int oys[4] = {0, 0, 0, 0};
    float wys[4] = {0.f, 0.f, 0.f, 0.f};
    int nr = 0;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1;
      float ly0, ly1;
      src_index(sh, oy, h, y0, y1, ly0, ly1);
      const float wy = (y0 == iy ? ly0 : 0.f) + (y1 == iy ? ly1 : 0.f);
      if (wy != 0.f && nr < 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k == nr) { oys[k] = oy; wys[k] = wy; }
        ++nr;
      }
    }
    for (int ix = threadIdx.y; ix < w; ix += blockDim.y) {
      int ox_lo = 0, ox_hi = W - 1;
      if (sw > 0.f) {
        ox_lo = max(0, (int)floorf((float)(ix - 1) / sw) - 1);
        ox_hi = min(W - 1, (int)ceilf((float)(ix + 1) / sw) + 1);
      }
      int oxs[4] = {0, 0, 0, 0};
      float wxs[4] = {0.f, 0.f, 0.f, 0.f};
      int nc = 0;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1;
        float lx0, lx1;
        src_index(sw, ox, w, x0, x1, lx0, lx1);
        const float wx = (x0 == ix ? lx0 : 0.f) + (x1 == ix ? lx1 : 0.f);
        if (wx != 0.f && nc < 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k == nc) { oxs[k] = ox; wxs[k] = wx; }
          ++nc;
        }
      }
      const size_t opix = ((size_t)n * h + iy) * w + ix;
      for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
        float4 v[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            v[r][j] = (r < nr && j < nc)
                          ? __ldg(reinterpret_cast<const float4*>(dhi + (((size_t)n * H + oys[r]) * W + oxs[j]) * ctot + coff + c))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (r < nr && j < nc) {
              const float ww = wys[r] * wxs[j];
              acc.x += ww * v[r][j].x;
              acc.y += ww * v[r][j].y;
              acc.z += ww * v[r][j].z;
              acc.w += ww * v[r][j].w;
            }
        *reinterpret_cast<float4*>(dlo + opix * C + c) = acc;
      }
    }
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
    long long blocks = (long long)N * h;                // one pair of output rows per block iteration
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
  static const bool rows_form = !(std::getenv("AIDE_UPSAMPLE_BWD_ROWS") && std::atoi(std::getenv("AIDE_UPSAMPLE_BWD_ROWS")) == 0);
  if (rows_form && h >= 2 && w >= 2) {        // (a 1-pixel map has > 4 taps per direction: every output reads it)
    int cx = C / 4 < 32 ? C / 4 : 32;
    while (cx & (cx - 1)) cx &= cx - 1;
    long long blocks = (long long)N * h;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    upsample2x_bwd_rows_kernel<<<(int)blocks, dim3(cx, 256 / cx), 0, as_stream(stream)>>>(
        dhi, dhi_ctot, dhi_coff, dlo, N, h, w, C, ac_scale(h, 2 * h), ac_scale(w, 2 * w));
    AIDE_CHECK_LAUNCH();
    return 0;
  }
  size_t total = (size_t)N * h * w * (C / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  upsample2x_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(dhi, dhi_ctot, dhi_coff, dlo, N, h, w, C,
                                                               ac_scale(h, 2 * h), ac_scale(w, 2 * w));
  AIDE_CHECK_LAUNCH();
  return 0;
}
