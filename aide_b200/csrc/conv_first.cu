// conv_first.cu -- the networks' first convolution: conv3x3, Cin = 3 (grey slice replicated to RGB,
// reference: datasetchaos_proposed/dataset.py:25-32; netblocks.py:24 with input_channel = 3), Cout = 32 (fuseunet)
// or 64 (UNet).  K = 27 does not fill an MMA K-slice and the layer is 0.2 % of the FLOPs, so it
// stays on the CUDA cores in exact fp32 in every mode; what matters is HBM traffic: the forward reads 12 B and
// writes 4*Cout B per pixel, the weight gradient reads 12 + 4*Cout B per pixel.
#include "common.cuh"

namespace aide {

// ------------------------------------------------------------------ forward
// CTA = 16x16 output pixels, 128 threads, TWO pixels per thread (rows ty and ty + 8): every weight vector read from
// shared memory feeds 8 FMAs instead of 4 -- with one pixel per thread the kernel was bound by the broadcast LDS.128
// stream (216 per 864 FMAs), 3.5x off its FMA time.  27 inputs per pixel in registers, 32 output channels at a time
// staged through shared memory for 128-byte coalesced stores and the BatchNorm partial statistics (sum z, sum z^2 per
// channel, one row per CTA; same per-channel summation order as before: 8 strided values per lane, then the warp tree).
constexpr int FT = 16;
template <int COUT>
__global__ void __launch_bounds__(128)
conv3x3_c3_fwd_kernel(const float* __restrict__ x, int x_ctot, int x_coff, const float* __restrict__ w /*[COUT][9][3]*/,
                      const float* __restrict__ bias, float* __restrict__ z, int z_ctot, int z_coff, int H, int W,
                      int tiles_w, int tiles_h, float* __restrict__ stat_partial) {
  __shared__ float in_s[FT + 2][FT + 2][3];
  __shared__ __align__(16) float w_s[27][COUT];
  __shared__ float out_s[256][33];
  const int tile = blockIdx.x;
  const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
  const int h0 = th * FT, w0 = tw * FT;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;                 // ty in [0, 8)
  for (int i = t; i < (FT + 2) * (FT + 2) * 3; i += 128) {
    const int c = i % 3, pp = i / 3, xx = pp % (FT + 2), yy = pp / (FT + 2);
    const int hh = h0 + yy - 1, ww = w0 + xx - 1;
    float v = 0.f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[(((size_t)n * H + hh) * W + ww) * x_ctot + x_coff + c];
    in_s[yy][xx][c] = v;
  }
  for (int i = t; i < 27 * COUT; i += 128) {
    const int co = i / 27, j = i % 27;
    w_s[j][co] = w[(size_t)co * 27 + j];
  }
  __syncthreads();
  float a0[27], a1[27];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        a0[(ky * 3 + kx) * 3 + c] = in_s[ty + ky][tx + kx][c];
        a1[(ky * 3 + kx) * 3 + c] = in_s[ty + 8 + ky][tx + kx][c];
      }
  const int ww = w0 + tx;
  const bool valid0 = h0 + ty < H && ww < W, valid1 = h0 + ty + 8 < H && ww < W;
  const int p0 = ty * 16 + tx, p1 = p0 + 128;
  const int warp = t >> 5, lane = t & 31;
#pragma unroll 1
  for (int co0 = 0; co0 < COUT; co0 += 32) {
    float acc0[32], acc1[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) acc0[k] = acc1[k] = bias ? __ldg(bias + co0 + k) : 0.f;
#pragma unroll
    for (int j = 0; j < 27; ++j) {
#pragma unroll
      for (int k = 0; k < 32; k += 4) {
        const float4 wv = *reinterpret_cast<const float4*>(&w_s[j][co0 + k]);
        acc0[k] = fmaf(a0[j], wv.x, acc0[k]);
        acc0[k + 1] = fmaf(a0[j], wv.y, acc0[k + 1]);
        acc0[k + 2] = fmaf(a0[j], wv.z, acc0[k + 2]);
        acc0[k + 3] = fmaf(a0[j], wv.w, acc0[k + 3]);
        acc1[k] = fmaf(a1[j], wv.x, acc1[k]);
        acc1[k + 1] = fmaf(a1[j], wv.y, acc1[k + 1]);
        acc1[k + 2] = fmaf(a1[j], wv.z, acc1[k + 2]);
        acc1[k + 3] = fmaf(a1[j], wv.w, acc1[k + 3]);
      }
    }
    if (co0) __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      out_s[p0][k] = valid0 ? acc0[k] : 0.f;
      out_s[p1][k] = valid1 ? acc1[k] : 0.f;
    }
    __syncthreads();
    // coalesced stores: 8 lanes cover the 32 channels (128 B) of one pixel
    for (int i = t; i < 256 * 8; i += 128) {
      const int p = i >> 3, c4 = (i & 7) * 4;
      const int h2 = h0 + (p >> 4), w2 = w0 + (p & 15);
      if (h2 < H && w2 < W) {
        const float4 v = make_float4(out_s[p][c4], out_s[p][c4 + 1], out_s[p][c4 + 2], out_s[p][c4 + 3]);
        *reinterpret_cast<float4*>(z + (((size_t)n * H + h2) * W + w2) * z_ctot + z_coff + co0 + c4) = v;
      }
    }
    if (stat_partial) {
      // warp w reduces channels 8w .. 8w+7 over the 256 pixels (invalid pixels hold zeros)
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        const int c = warp * 8 + cc;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float v = out_s[lane + 32 * k][c];
          s1 += v;
          s2 = fmaf(v, v, s2);
        }
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (lane == 0) {
          stat_partial[((size_t)tile * 2 + 0) * COUT + co0 + c] = s1;
          stat_partial[((size_t)tile * 2 + 1) * COUT + co0 + c] = s2;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ weight gradient
// dW[co][tap][ci] = sum_pixels dz[p][co] * x[p + tap][ci].  lane <-> output channel(s); a warp walks one row of 32
// pixels of an 8x32 tile: dz is read with one coalesced 128-byte load per pixel, the 27 inputs come from a
// shared-memory halo tile as broadcasts.  CTAs are persistent over tiles (accumulators stay in registers), then
// the 8 warps are folded in a fixed order and one row per CTA goes to the workspace [grid][COUT][27].
constexpr int GT_H = 8, GT_W = 32;
template <int COUT>
__global__ void __launch_bounds__(256)
wgrad_c3_kernel(const float* __restrict__ x, int x_ctot, int x_coff, const float* __restrict__ dz /*[N,H,W,COUT]*/,
                int H, int W, int tiles_w, int tiles_h, int n_tiles, float* __restrict__ ws) {
  constexpr int R = COUT / 32;
  __shared__ float in_s[GT_H + 2][GT_W + 2][3];
  __shared__ float red[4][27 * COUT];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  float acc[R][27];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int j = 0; j < 27; ++j) acc[r][j] = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
    const int h0 = th * GT_H, w0 = tw * GT_W;
    __syncthreads();
    for (int i = t; i < (GT_H + 2) * (GT_W + 2) * 3; i += 256) {
      const int c = i % 3, pp = i / 3, xx = pp % (GT_W + 2), yy = pp / (GT_W + 2);
      const int hh = h0 + yy - 1, ww = w0 + xx - 1;
      float v = 0.f;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = x[(((size_t)n * H + hh) * W + ww) * x_ctot + x_coff + c];
      in_s[yy][xx][c] = v;
    }
    __syncthreads();
    const int hh = h0 + warp;
    if (hh < H) {
      const float* drow = dz + (((size_t)n * H + hh) * W + w0) * COUT;
      const int npx = min(GT_W, W - w0);
#pragma unroll 4
      for (int px = 0; px < npx; ++px) {
        float d[R];
#pragma unroll
        for (int r = 0; r < R; ++r) d[r] = __ldg(drow + (size_t)px * COUT + r * 32 + lane);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const float xv = in_s[warp + ky][px + kx][c];
#pragma unroll
              for (int r = 0; r < R; ++r) acc[r][(ky * 3 + kx) * 3 + c] = fmaf(d[r], xv, acc[r][(ky * 3 + kx) * 3 + c]);
            }
      }
    }
  }
  // fold the 8 warps: (4..7 -> 0..3), (2,3 -> 0,1), (1 -> 0); fixed order
  for (int half = 4; half >= 1; half >>= 1) {
    __syncthreads();
    if (warp >= half && warp < 2 * half) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 27; ++j) red[warp - half][(r * 32 + lane) * 27 + j] = acc[r][j];
    }
    __syncthreads();
    if (warp < half) {
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 27; ++j) acc[r][j] += red[warp][(r * 32 + lane) * 27 + j];
    }
  }
  if (warp == 0) {
    float* out = ws + (size_t)blockIdx.x * COUT * 27;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < 27; ++j) out[(r * 32 + lane) * 27 + j] = acc[r][j];
  }
}

// dw_oihw[co][ci][tap] = sum_rows ws[row][co][tap*3+ci]   (fp64, fixed order; block = 32 columns x 8 row lanes)
__global__ void wgrad_c3_reduce_kernel(const float* __restrict__ ws, int rows, int cout, float* __restrict__ dw) {
  __shared__ double sm[8][33];
  const int cols = cout * 27;
  const int j = blockIdx.x * 32 + threadIdx.x;
  double a = 0.0;
  if (j < cols)
    for (int r = threadIdx.y; r < rows; r += 8) a += (double)ws[(size_t)r * cols + j];
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && j < cols) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k][threadIdx.x];
    const int co = j / 27, tap = (j % 27) / 3, ci = j % 3;
    dw[((size_t)co * 3 + ci) * 9 + tap] = (float)s;
  }
}

// ------------------------------------------------------------------ host
bool c3_shape_ok(int cin, int cout) { return cin == 3 && (cout == 32 || cout == 64); }

int c3_stat_rows(int N, int H, int W) { return N * ceil_div(H, FT) * ceil_div(W, FT); }

int c3_conv3x3(const float* x, int x_ctot, int x_coff, const float* w, const float* bias, float* z, int z_ctot,
               int z_coff, int cout, int N, int H, int W, float* stat_partial, cudaStream_t st) {
  const int tiles_w = ceil_div(W, FT), tiles_h = ceil_div(H, FT);
  const int grid = N * tiles_w * tiles_h;
  if (cout == 32)
    conv3x3_c3_fwd_kernel<32><<<grid, 128, 0, st>>>(x, x_ctot, x_coff, w, bias, z, z_ctot, z_coff, H, W, tiles_w,
                                                     tiles_h, stat_partial);
  else
    conv3x3_c3_fwd_kernel<64><<<grid, 128, 0, st>>>(x, x_ctot, x_coff, w, bias, z, z_ctot, z_coff, H, W, tiles_w,
                                                     tiles_h, stat_partial);
  AIDE_CHECK_LAUNCH();
  return 0;
}

static int c3_wgrad_grid(int N, int H, int W) {
  const int tiles = N * ceil_div(H, GT_H) * ceil_div(W, GT_W);
  return tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs;
}

size_t c3_wgrad_workspace_bytes(int cout, int N, int H, int W) {
  return (size_t)c3_wgrad_grid(N, H, W) * cout * 27 * sizeof(float);
}

int c3_wgrad(const float* x, int x_ctot, int x_coff, const float* dz, int cout, int N, int H, int W, float* ws,
             size_t ws_bytes, float* dw, cudaStream_t st) {
  const int tiles_w = ceil_div(W, GT_W), tiles_h = ceil_div(H, GT_H);
  const int n_tiles = N * tiles_w * tiles_h;
  const int grid = c3_wgrad_grid(N, H, W);
  AIDE_REQUIRE(ws && ws_bytes >= (size_t)grid * cout * 27 * sizeof(float), "conv3x3_wgrad(c3): workspace too small");
  if (cout == 32)
    wgrad_c3_kernel<32><<<grid, 256, 0, st>>>(x, x_ctot, x_coff, dz, H, W, tiles_w, tiles_h, n_tiles, ws);
  else
    wgrad_c3_kernel<64><<<grid, 256, 0, st>>>(x, x_ctot, x_coff, dz, H, W, tiles_w, tiles_h, n_tiles, ws);
  AIDE_CHECK_LAUNCH();
  wgrad_c3_reduce_kernel<<<ceil_div(cout * 27, 32), dim3(32, 8), 0, st>>>(ws, grid, cout, dw);
  AIDE_CHECK_LAUNCH();
  return 0;
}

}  // namespace aide
