// attention.cu -- Spatial_Attention gate of the attention variants (fuseunetsa, fuseunetsaseparate, UNetsa):
//
//   reference: models_twomodalinputs/netblocks.py:68-89 (twin: models_singlemodalinput/UNet.py:85-108)
//     y -> conv1 (1x1, C -> r = C/16) -> conv2 (3x3, dilation d, padding d, r -> r) -> conv3 (same) -> conv4 (1x1, r -> 1)
//       -> BatchNorm2d(1) -> sigmoid = gate [N,1,H,W];      the networks then use  gate * y
//   (fuseunet.py:145-147, UNet.py:191-192).  No non-linearity sits between the four convolutions.
//
// The branch carries C/16 channels: < 0.5 % of a block's FLOPs, HBM / latency bound -> plain fp32 CUDA-core kernels over
// NHWC data, one thread per output element, fixed-order two-stage reductions (bit-reproducible).  Forward:
//   sa_conv1 -> sa_dilated x2 -> sa_conv4 (+ partial BatchNorm statistics) -> [aide_bn_finalize_grouped, C = 1]
//   -> sa_gate_apply (gate = sigmoid(scale*a+shift); writes gate*y to the consumer's channel slice and its 2x2 max-pool).
// Backward (aide_sa_bwd_gate, aide_sa_bwd_chain): gradient routing of the gated tensor (same-resolution and max-pool
// routed sources), d gate, BatchNorm(1) backward, then the four convolutions in reverse with their weight gradients.
#include "common.cuh"

namespace aide {

int launch_reduce_rows(const float* partial, int rows, int ld, int cols, float* out, cudaStream_t st);

namespace {

constexpr int kSaThreads = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------- forward
// a1[p][j] = b1[j] + sum_c y[p][c] * w1[j][c]          one thread per (pixel, j)
template <int FMT>
__global__ void sa_conv1_kernel(CView y, int C, const float* __restrict__ w, const float* __restrict__ b, int r,
                                float* __restrict__ a1, size_t P) {
  const size_t total = P * r;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i / r;
    const int j = (int)(i - p * r);
    const float* wj = w + (size_t)j * C;
    float acc = 0.f;
    for (int c = 0; c < C; c += 4) {
      const float4 v = ld4<FMT>(y.p0, y.p1, p * y.ctot + y.coff + c);
      const float4 ww = __ldg(reinterpret_cast<const float4*>(wj + c));
      acc += v.x * ww.x + v.y * ww.y + v.z * ww.z + v.w * ww.w;
    }
    a1[i] = acc + b[j];
  }
}

// out[p][co] = b[co] + sum_{ky,kx,ci} in[(h + (ky-1) d, w + (kx-1) d)][ci] * w[co][ci][ky][kx]      (zero padding)
__global__ void sa_dilated_fwd_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ b,
                                      int r, int dil, int N, int H, int W, float* __restrict__ out) {
  const size_t total = (size_t)N * H * W * r;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % r);
    const size_t p = i / r;
    const int x = (int)(p % W), yy = (int)((p / W) % H);
    const size_t n = p / ((size_t)W * H);
    float acc = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int sy = yy + (ky - 1) * dil;
      if (sy < 0 || sy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int sx = x + (kx - 1) * dil;
        if (sx < 0 || sx >= W) continue;
        const float* src = in + ((n * H + sy) * W + sx) * r;
        const float* wk = w + (size_t)co * r * 9 + ky * 3 + kx;
        for (int ci = 0; ci < r; ++ci) acc += src[ci] * __ldg(wk + ci * 9);
      }
    }
    out[i] = acc + b[co];
  }
}

// a[p] = b4 + sum_j a3[p][j] * w4[j];  per-block partial (sum a, sum a^2) -> stat_partial[(n * bpi + blockIdx.x)][2]
__global__ void sa_conv4_kernel(const float* __restrict__ a3, const float* __restrict__ w4, const float* __restrict__ b4,
                                int r, int HW, float* __restrict__ a, float* __restrict__ stat_partial) {
  __shared__ float red[2][kSaThreads / 32];
  const int n = blockIdx.y;
  float s1 = 0.f, s2 = 0.f;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < HW; q += gridDim.x * blockDim.x) {
    const size_t p = (size_t)n * HW + q;
    float acc = 0.f;
    for (int j = 0; j < r; ++j) acc += a3[p * r + j] * __ldg(w4 + j);
    acc += b4[0];
    a[p] = acc;
    s1 += acc;
    s2 += acc * acc;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = s1;
    red[1][threadIdx.x >> 5] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < kSaThreads / 32; ++k) {
      t1 += red[0][k];
      t2 += red[1][k];
    }
    float* out = stat_partial + ((size_t)n * gridDim.x + blockIdx.x) * 2;
    out[0] = t1;
    out[1] = t2;
  }
}

__device__ __forceinline__ float4 scale4(float4 v, float g) { return make_float4(v.x * g, v.y * g, v.z * g, v.w * g); }
__device__ __forceinline__ float4 vmax4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// gate[p] = sigmoid(scale * a[p] + shift) (scale/shift of image n's statistics group); t = gate * y -> dst, max-pool(t) -> pools
// blockDim = (cx, ty): threadIdx.x owns 4 channels, rows of pixels / 2x2 windows are dealt to blocks (cf. bn_relu_apply_kernel)
template <int FMT, bool POOL>
__global__ void sa_gate_apply_kernel(CView y, const float* __restrict__ a, const float* __restrict__ ss, int N, int H,
                                     int W, int C, int imgs_per_group, float* __restrict__ gate, View dst, View pa, View pb) {
  const int rows_per_img = POOL ? (H >> 1) : H;
  const int total_rows = N * rows_per_img;
  for (int row = blockIdx.x; row < total_rows; row += gridDim.x) {
    const int n = row / rows_per_img, yy = row - n * rows_per_img;
    const float sc = __ldg(ss + (size_t)(n / imgs_per_group) * 2), sh = __ldg(ss + (size_t)(n / imgs_per_group) * 2 + 1);
    if constexpr (!POOL) {
      const size_t p0 = ((size_t)n * H + yy) * W;
      for (int x = threadIdx.y; x < W; x += blockDim.y) {
        const size_t pix = p0 + x;
        const float g = sigmoidf_(sc * a[pix] + sh);
        if (threadIdx.x == 0 && blockIdx.y == 0) gate[pix] = g;
        for (int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4; c < C; c += gridDim.y * blockDim.x * 4)
          st4<FMT>(dst.p0, dst.p1, pix * dst.ctot + dst.coff + c, scale4(ld4<FMT>(y.p0, y.p1, pix * y.ctot + y.coff + c), g));
      }
    } else {
      const int Wh = W >> 1;
      const size_t prow = ((size_t)n * H + 2 * yy) * W;
      const size_t wrow = ((size_t)n * rows_per_img + yy) * Wh;
      for (int wx = threadIdx.y; wx < Wh; wx += blockDim.y) {
        const size_t p[4] = {prow + 2 * wx, prow + 2 * wx + 1, prow + 2 * wx + W, prow + 2 * wx + W + 1};
        float g[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          g[k] = sigmoidf_(sc * a[p[k]] + sh);
          if (threadIdx.x == 0 && blockIdx.y == 0) gate[p[k]] = g[k];
        }
        for (int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4; c < C; c += gridDim.y * blockDim.x * 4) {
          float4 t[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            t[k] = scale4(ld4<FMT>(y.p0, y.p1, p[k] * y.ctot + y.coff + c), g[k]);
            if (dst.p0) st4<FMT>(dst.p0, dst.p1, p[k] * dst.ctot + dst.coff + c, t[k]);
          }
          const float4 m = vmax4(vmax4(t[0], t[1]), vmax4(t[2], t[3]));
          if (pa.p0) st4<FMT>(pa.p0, pa.p1, (wrow + wx) * pa.ctot + pa.coff + c, m);
          if (pb.p0) st4<FMT>(pb.p0, pb.p1, (wrow + wx) * pb.ctot + pb.coff + c, m);
        }
      }
    }
  }
}

// ---------------------------------------------------------------- backward
struct SaSrcs {
  const float* dptr[3];
  int dctot[3], dcoff[3], nd;
  const float* pptr[3];
  int pctot[3], pcoff[3], np;
};

__device__ __forceinline__ int argmax4f(float a, float b, float c, float d) {
  int k = 0;
  float m = a;
  if (b > m) { m = b; k = 1; }
  if (c > m) { m = c; k = 2; }
  if (d > m) { m = d; k = 3; }
  return k;
}

// One warp per pixel (WINDOW = false) or per 2x2 window (WINDOW = true, needed when max-pool routed sources exist).
//   dt[c]   = sum of the upstream gradients w.r.t. t = gate * y (pooled ones go to the window's first maximum of t)
//   dy[p][c] = dt[c] * gate[p]                         (the conv1 path is added later by sa_conv1_bwd_data_kernel)
//   dahat[p] = (sum_c dt[c] * y[p][c]) * gate (1 - gate)
// and per-block partial sums of (dahat, dahat * ahat), ahat = (a - mean) * rstd.
template <int FMT, bool WINDOW>
__global__ void sa_bwd_gate_kernel(CView y, int C, const float* __restrict__ gate, const float* __restrict__ a,
                                   const float* __restrict__ mr, int N, int H, int W, SaSrcs s, float* __restrict__ dy,
                                   float* __restrict__ dahat, float* __restrict__ partial) {
  __shared__ float red[2][kSaThreads / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float mean = mr[0], rstd = mr[1];
  constexpr int KP = WINDOW ? 4 : 1;
  const int Hh = WINDOW ? (H >> 1) : H, Wh = WINDOW ? (W >> 1) : W;
  const size_t nwin = (size_t)N * Hh * Wh;
  float s1 = 0.f, s2 = 0.f;
  for (size_t win = (size_t)blockIdx.x * (kSaThreads / 32) + wid; win < nwin; win += (size_t)gridDim.x * (kSaThreads / 32)) {
    size_t p[KP];
    if constexpr (WINDOW) {
      const int wx = (int)(win % Wh), hy = (int)((win / Wh) % Hh);
      const size_t n = win / ((size_t)Wh * Hh);
      p[0] = (n * H + 2 * hy) * W + 2 * wx;
      p[1] = p[0] + 1;
      p[2] = p[0] + W;
      p[3] = p[2] + 1;
    } else {
      p[0] = win;
    }
    float g[KP], dg[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      g[k] = gate[p[k]];
      dg[k] = 0.f;
    }
    for (int c = lane * 4; c < C; c += 128) {
      float4 yy[KP], dt[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        yy[k] = ld4<FMT>(y.p0, y.p1, p[k] * y.ctot + y.coff + c);
        dt[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int d = 0; d < s.nd; ++d) {
          const float4 v = *reinterpret_cast<const float4*>(s.dptr[d] + p[k] * s.dctot[d] + s.dcoff[d] + c);
          dt[k].x += v.x; dt[k].y += v.y; dt[k].z += v.z; dt[k].w += v.w;
        }
      }
      if constexpr (WINDOW) {
        float4 pg = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int d = 0; d < s.np; ++d) {
          const float4 v = *reinterpret_cast<const float4*>(s.pptr[d] + win * s.pctot[d] + s.pcoff[d] + c);
          pg.x += v.x; pg.y += v.y; pg.z += v.z; pg.w += v.w;
        }
        // the forward pooled t as it was STORED (operand format) -- recompute the products exactly as sa_gate_apply did
        const int kx = argmax4f(yy[0].x * g[0], yy[1].x * g[1], yy[2].x * g[2], yy[3].x * g[3]);
        const int ky = argmax4f(yy[0].y * g[0], yy[1].y * g[1], yy[2].y * g[2], yy[3].y * g[3]);
        const int kz = argmax4f(yy[0].z * g[0], yy[1].z * g[1], yy[2].z * g[2], yy[3].z * g[3]);
        const int kw = argmax4f(yy[0].w * g[0], yy[1].w * g[1], yy[2].w * g[2], yy[3].w * g[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (kx == k) dt[k].x += pg.x;
          if (ky == k) dt[k].y += pg.y;
          if (kz == k) dt[k].z += pg.z;
          if (kw == k) dt[k].w += pg.w;
        }
      }
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        *reinterpret_cast<float4*>(dy + p[k] * C + c) = scale4(dt[k], g[k]);
        dg[k] += dt[k].x * yy[k].x + dt[k].y * yy[k].y + dt[k].z * yy[k].z + dt[k].w * yy[k].w;
      }
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const float t = warp_sum(dg[k]);
      const float dh = t * g[k] * (1.0f - g[k]);
      if (lane == 0) {
        dahat[p[k]] = dh;
        s1 += dh;
        s2 += dh * ((a[p[k]] - mean) * rstd);
      }
    }
  }
  if (lane == 0) {
    red[0][wid] = s1;
    red[1][wid] = s2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t1 = 0.f, t2 = 0.f;
    for (int k = 0; k < kSaThreads / 32; ++k) {
      t1 += red[0][k];
      t2 += red[1][k];
    }
    partial[(size_t)blockIdx.x * 2] = t1;
    partial[(size_t)blockIdx.x * 2 + 1] = t2;
  }
}

// BatchNorm2d(1) backward + conv4 backward:
//   da[p]     = gamma * rstd * (dahat - mean(dahat) - ahat * mean(dahat * ahat))
//   da3[p][j] = w4[j] * da[p];  partial rows [r + 1]: dW4[j] += a3[p][j] * da[p], db4 += da[p]
__global__ void sa_bn_conv4_bwd_kernel(const float* __restrict__ dahat, const float* __restrict__ a,
                                       const float* __restrict__ mr, const float* __restrict__ gamma,
                                       const float* __restrict__ sums /*[2]*/, float inv_count, const float* __restrict__ a3,
                                       const float* __restrict__ w4, int r, size_t P, float* __restrict__ da3,
                                       float* __restrict__ partial /*[rows][r+1]*/) {
  extern __shared__ float sm[];                 // [warps][r + 1]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float mean = mr[0], rstd = mr[1];
  const float k0 = gamma[0] * rstd, m1 = sums[0] * inv_count, m2 = sums[1] * inv_count;
  float db = 0.f;
  float* mine = sm + (size_t)wid * (r + 1);
  for (int j = lane; j <= r; j += 32) mine[j] = 0.f;
  __syncwarp();
  // a warp walks pixels; lanes hold the r channel products (r <= 64: lanes j and j + 32)
  float acc0 = 0.f, acc1 = 0.f;
  for (size_t p = (size_t)blockIdx.x * nw + wid; p < P; p += (size_t)gridDim.x * nw) {
    const float ahat = (a[p] - mean) * rstd;
    const float d = k0 * (dahat[p] - m1 - ahat * m2);
    db += d;
    if (lane < r) {
      da3[p * r + lane] = __ldg(w4 + lane) * d;
      acc0 += a3[p * r + lane] * d;
    }
    if (lane + 32 < r) {
      da3[p * r + lane + 32] = __ldg(w4 + lane + 32) * d;
      acc1 += a3[p * r + lane + 32] * d;
    }
  }
  if (lane < r) mine[lane] = acc0;
  if (lane + 32 < r) mine[lane + 32] = acc1;
  if (lane == 0) mine[r] = db;
  __syncthreads();
  for (int j = threadIdx.x; j <= r; j += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += sm[(size_t)w * (r + 1) + j];
    partial[(size_t)blockIdx.x * (r + 1) + j] = t;
  }
}

// data gradient of the dilated conv: din[q][ci] = sum_{ky,kx,co} dout[(h - (ky-1) d, w - (kx-1) d)][co] * w[co][ci][ky][kx]
__global__ void sa_dilated_bwd_data_kernel(const float* __restrict__ dout, const float* __restrict__ w, int r, int dil,
                                           int N, int H, int W, float* __restrict__ din) {
  const size_t total = (size_t)N * H * W * r;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % r);
    const size_t p = i / r;
    const int x = (int)(p % W), yy = (int)((p / W) % H);
    const size_t n = p / ((size_t)W * H);
    float acc = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int sy = yy - (ky - 1) * dil;
      if (sy < 0 || sy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int sx = x - (kx - 1) * dil;
        if (sx < 0 || sx >= W) continue;
        const float* src = dout + ((n * H + sy) * W + sx) * r;
        const float* wk = w + (size_t)ci * 9 + ky * 3 + kx;
        for (int co = 0; co < r; ++co) acc += src[co] * __ldg(wk + (size_t)co * r * 9);
      }
    }
    din[i] = acc;
  }
}

// weight / bias gradient of the dilated conv, partial over a chunk of pixels per block row:
//   output o < r*r*9: (co, ci, ky, kx) in OIHW order: sum_p dout[p][co] * in[p + off][ci];  o >= r*r*9: bias co: sum_p dout[p][co]
__global__ void sa_dilated_bwd_w_kernel(const float* __restrict__ dout, const float* __restrict__ in, int r, int dil, int N,
                                        int H, int W, int chunk, float* __restrict__ partial /*[rows][r*r*9 + r]*/) {
  const int n_out = r * r * 9 + r;
  const int o = blockIdx.y * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const size_t P = (size_t)N * H * W;
  const size_t p0 = (size_t)blockIdx.x * chunk, p1 = p0 + chunk < P ? p0 + chunk : P;
  float acc = 0.f;
  if (o < r * r * 9) {
    const int kx = o % 3, ky = (o / 3) % 3, ci = (o / 9) % r, co = o / (9 * r);
    const int oy = (ky - 1) * dil, ox = (kx - 1) * dil;
    for (size_t p = p0; p < p1; ++p) {
      const int x = (int)(p % W), yy = (int)((p / W) % H);
      const int sy = yy + oy, sx = x + ox;
      if (sy < 0 || sy >= H || sx < 0 || sx >= W) continue;
      const size_t q = p + (ptrdiff_t)oy * W + ox;
      acc += dout[p * r + co] * in[q * r + ci];
    }
  } else {
    const int co = o - r * r * 9;
    for (size_t p = p0; p < p1; ++p) acc += dout[p * r + co];
  }
  partial[(size_t)blockIdx.x * n_out + o] = acc;
}

// conv1 backward, data: dy[p][c] += sum_j w1[j][c] * da1[p][j]      (thread per pixel x 4 channels)
__global__ void sa_conv1_bwd_data_kernel(const float* __restrict__ da1, const float* __restrict__ w1, int r, int C, size_t P,
                                         float* __restrict__ dy) {
  const int C4 = C >> 2;
  const size_t total = P * C4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i / C4;
    const int c = (int)(i - p * C4) * 4;
    float4 acc = *reinterpret_cast<const float4*>(dy + p * C + c);
    for (int j = 0; j < r; ++j) {
      const float d = da1[p * r + j];
      const float4 ww = __ldg(reinterpret_cast<const float4*>(w1 + (size_t)j * C + c));
      acc.x += ww.x * d; acc.y += ww.y * d; acc.z += ww.z * d; acc.w += ww.w * d;
    }
    *reinterpret_cast<float4*>(dy + p * C + c) = acc;
  }
}

// conv1 backward, weights: output o < r*C: (j, c): sum_p da1[p][j] * y[p][c];  o >= r*C: bias j
template <int FMT>
__global__ void sa_conv1_bwd_w_kernel(const float* __restrict__ da1, CView y, int r, int C, size_t P, int chunk,
                                      float* __restrict__ partial /*[rows][r*C + r]*/) {
  const int n_out = r * C + r;
  const int o = blockIdx.y * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const size_t p0 = (size_t)blockIdx.x * chunk, p1 = p0 + chunk < P ? p0 + chunk : P;
  float acc = 0.f;
  if (o < r * C) {
    const int c = o % C, j = o / C;
    for (size_t p = p0; p < p1; ++p) acc += da1[p * r + j] * ld1<FMT>(y.p0, y.p1, p * y.ctot + y.coff + c);
  } else {
    const int j = o - r * C;
    for (size_t p = p0; p < p1; ++p) acc += da1[p * r + j];
  }
  partial[(size_t)blockIdx.x * n_out + o] = acc;
}

int grid_for(size_t total, int threads = kSaThreads) {
  long long b = (long long)((total + threads - 1) / threads);
  if (b > (long long)kNumSMs * 16) b = (long long)kNumSMs * 16;
  return b < 1 ? 1 : (int)b;
}

int stat_blocks_per_image(int H, int W) {
  int b = ceil_div((long long)H * W, kSaThreads * 4);
  return b < 1 ? 1 : (b > 64 ? 64 : b);
}

// rows of the pixel-chunked weight-gradient partials
int wgrad_rows(size_t P) {
  long long rws = (long long)((P + 63) / 64);
  if (rws > kNumSMs * 4) rws = kNumSMs * 4;
  return rws < 1 ? 1 : (int)rws;
}
int gate_rows(int N, int H, int W, bool window) {
  const size_t nwin = (size_t)N * (window ? H / 2 : H) * (window ? W / 2 : W);
  long long b = (long long)((nwin + (kSaThreads / 32) * 4 - 1) / ((kSaThreads / 32) * 4));
  if (b > kNumSMs * 8) b = kNumSMs * 8;
  return b < 1 ? 1 : (int)b;
}

}  // namespace
}  // namespace aide

using namespace aide;

extern "C" int aide_sa_stat_rows(int N, int H, int W) { return N * stat_blocks_per_image(H, W); }

extern "C" int aide_sa_fwd(int fmt, const void* y_p0, const void* y_p1, int y_ctot, int y_coff, int C, int r, int dilation,
                           const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                           const float* b3, const float* w4, const float* b4, float* a1, float* a2, float* a3, float* a,
                           float* stat_partial, int N, int H, int W, void* stream) {
  AIDE_REQUIRE(y_p0 && w1 && b1 && w2 && b2 && w3 && b3 && w4 && b4 && a1 && a2 && a3 && a && stat_partial,
               "sa_fwd: null argument");
  AIDE_REQUIRE(C % 4 == 0 && r >= 1 && r <= 64 && dilation >= 1 && y_coff % 4 == 0 && y_ctot % 4 == 0,
               "sa_fwd: need C %% 4 == 0, 1 <= r <= 64 (got C=%d r=%d)", C, r);
  cudaStream_t st = as_stream(stream);
  const size_t P = (size_t)N * H * W;
  CView y{y_p0, y_p1, y_ctot, y_coff};
  AIDE_DISPATCH_FMT(fmt, (sa_conv1_kernel<FMT><<<grid_for(P * r), kSaThreads, 0, st>>>(y, C, w1, b1, r, a1, P)));
  AIDE_CHECK_LAUNCH();
  sa_dilated_fwd_kernel<<<grid_for(P * r), kSaThreads, 0, st>>>(a1, w2, b2, r, dilation, N, H, W, a2);
  AIDE_CHECK_LAUNCH();
  sa_dilated_fwd_kernel<<<grid_for(P * r), kSaThreads, 0, st>>>(a2, w3, b3, r, dilation, N, H, W, a3);
  AIDE_CHECK_LAUNCH();
  sa_conv4_kernel<<<dim3(stat_blocks_per_image(H, W), N), kSaThreads, 0, st>>>(a3, w4, b4, r, H * W, a, stat_partial);
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_sa_gate_apply(int fmt, const void* y_p0, const void* y_p1, int y_ctot, int y_coff, const float* a,
                                  const float* scale_shift, int N, int imgs_per_group, int H, int W, int C, float* gate,
                                  void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff, void* poolA_p0, void* poolA_p1,
                                  int poolA_ctot, int poolA_coff, void* poolB_p0, void* poolB_p1, int poolB_ctot,
                                  int poolB_coff, void* stream) {
  AIDE_REQUIRE(y_p0 && a && scale_shift && gate && C % 4 == 0 && N > 0 && imgs_per_group > 0 && N % imgs_per_group == 0,
               "sa_gate_apply: bad arguments");
  const bool pool = poolA_p0 || poolB_p0;
  AIDE_REQUIRE(pool || dst_p0, "sa_gate_apply: no destination");
  AIDE_REQUIRE(!pool || (H % 2 == 0 && W % 2 == 0), "sa_gate_apply: pooling needs even H, W");
  AIDE_REQUIRE(dst_coff % 4 == 0 && dst_ctot % 4 == 0 && poolA_coff % 4 == 0 && poolB_coff % 4 == 0 && y_coff % 4 == 0,
               "sa_gate_apply: channel offsets must be multiples of 4");
  CView y{y_p0, y_p1, y_ctot, y_coff};
  View dst{dst_p0, dst_p1, dst_ctot, dst_coff}, pa{poolA_p0, poolA_p1, poolA_ctot, poolA_coff},
      pb{poolB_p0, poolB_p1, poolB_ctot, poolB_coff};
  const int c4 = C / 4;
  int cx = c4 < 32 ? c4 : 32;
  while (cx & (cx - 1)) cx &= cx - 1;
  const dim3 block(cx, 256 / cx);
  const int total_rows = N * (pool ? H / 2 : H);
  const int rb = total_rows < kNumSMs * 16 ? total_rows : kNumSMs * 16;
  const dim3 grid(rb, 1);
  if (pool) {
    AIDE_DISPATCH_FMT(fmt, (sa_gate_apply_kernel<FMT, true><<<grid, block, 0, as_stream(stream)>>>(
                               y, a, scale_shift, N, H, W, C, imgs_per_group, gate, dst, pa, pb)));
  } else {
    AIDE_DISPATCH_FMT(fmt, (sa_gate_apply_kernel<FMT, false><<<grid, block, 0, as_stream(stream)>>>(
                               y, a, scale_shift, N, H, W, C, imgs_per_group, gate, dst, pa, pb)));
  }
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_sa_bwd_rows(int N, int H, int W, int pooled) { return gate_rows(N, H, W, pooled != 0); }

extern "C" int aide_sa_bwd_gate(int fmt, const void* y_p0, const void* y_p1, int y_ctot, int y_coff, int C,
                                const float* gate, const float* a, const float* mean_rstd, int N, int H, int W,
                                const float* const* direct_ptr, const int* direct_ctot, const int* direct_coff,
                                int n_direct, const float* const* pool_ptr, const int* pool_ctot, const int* pool_coff,
                                int n_pool, float* dy, float* dahat, float* partial, void* stream) {
  AIDE_REQUIRE(y_p0 && gate && a && mean_rstd && dy && dahat && partial, "sa_bwd_gate: null argument");
  AIDE_REQUIRE(C % 4 == 0 && n_direct >= 0 && n_direct <= 3 && n_pool >= 0 && n_pool <= 3 && n_direct + n_pool > 0,
               "sa_bwd_gate: up to 3 direct and 3 pooled gradient sources (at least one)");
  AIDE_REQUIRE(n_pool == 0 || (H % 2 == 0 && W % 2 == 0), "sa_bwd_gate: pooled sources need even H, W");
  SaSrcs s{};
  s.nd = n_direct;
  s.np = n_pool;
  for (int i = 0; i < n_direct; ++i) {
    s.dptr[i] = direct_ptr[i]; s.dctot[i] = direct_ctot[i]; s.dcoff[i] = direct_coff[i];
    AIDE_REQUIRE(s.dptr[i] && s.dctot[i] % 4 == 0 && s.dcoff[i] % 4 == 0, "sa_bwd_gate: bad direct source");
  }
  for (int i = 0; i < n_pool; ++i) {
    s.pptr[i] = pool_ptr[i]; s.pctot[i] = pool_ctot[i]; s.pcoff[i] = pool_coff[i];
    AIDE_REQUIRE(s.pptr[i] && s.pctot[i] % 4 == 0 && s.pcoff[i] % 4 == 0, "sa_bwd_gate: bad pooled source");
  }
  CView y{y_p0, y_p1, y_ctot, y_coff};
  const bool window = n_pool > 0;
  const int rows = gate_rows(N, H, W, window);
  if (window) {
    AIDE_DISPATCH_FMT(fmt, (sa_bwd_gate_kernel<FMT, true><<<rows, kSaThreads, 0, as_stream(stream)>>>(
                               y, C, gate, a, mean_rstd, N, H, W, s, dy, dahat, partial)));
  } else {
    AIDE_DISPATCH_FMT(fmt, (sa_bwd_gate_kernel<FMT, false><<<rows, kSaThreads, 0, as_stream(stream)>>>(
                               y, C, gate, a, mean_rstd, N, H, W, s, dy, dahat, partial)));
  }
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" size_t aide_sa_bwd_workspace_floats(int C, int r, int N, int H, int W) {
  const size_t P = (size_t)N * H * W;
  size_t n_out = (size_t)r * r * 9 + r;
  if ((size_t)r * C + r > n_out) n_out = (size_t)r * C + r;
  size_t part = (size_t)wgrad_rows(P) * n_out;                       // pixel-chunked weight-gradient partial rows ...
  const size_t part4 = (size_t)kNumSMs * 8 * (r + 1);                // ... or the conv4 / BatchNorm stage's block rows
  if (part4 > part) part = part4;
  return 2 * P * r + part + 64;
}

// gradients: dw1 [r][C] db1 [r] dw2 [r][r][3][3] db2 [r] dw3 db3 dw4 [r] db4 [1] dbn [2] = {dbeta, dgamma}
extern "C" int aide_sa_bwd_chain(int fmt, const void* y_p0, const void* y_p1, int y_ctot, int y_coff, int C, int r,
                                 int dilation, const float* w1, const float* w2, const float* w3, const float* w4,
                                 const float* gamma, const float* a1, const float* a2, const float* a3, const float* a,
                                 const float* mean_rstd, const float* dahat, const float* partial, int partial_rows, int N,
                                 int H, int W, float* workspace, size_t workspace_floats, float* dy, float* dw1,
                                 float* db1, float* dw2, float* db2, float* dw3, float* db3, float* dw4, float* db4,
                                 float* dbn, void* stream) {
  AIDE_REQUIRE(y_p0 && w1 && w2 && w3 && w4 && gamma && a1 && a2 && a3 && a && mean_rstd && dahat && partial && workspace &&
                   dy && dw1 && db1 && dw2 && db2 && dw3 && db3 && dw4 && db4 && dbn,
               "sa_bwd_chain: null argument");
  AIDE_REQUIRE(dw1 + (size_t)r * C == db1 && dw2 + (size_t)r * r * 9 == db2 && dw3 + (size_t)r * r * 9 == db3 && dw4 + r == db4,
               "sa_bwd_chain: every bias gradient must directly follow its weight gradient");
  AIDE_REQUIRE(workspace_floats >= aide_sa_bwd_workspace_floats(C, r, N, H, W), "sa_bwd_chain: workspace too small");
  cudaStream_t st = as_stream(stream);
  const size_t P = (size_t)N * H * W;
  float* buf0 = workspace;                  // da3, then da1
  float* buf1 = workspace + P * r;          // da2
  float* part = workspace + 2 * P * r;
  CView y{y_p0, y_p1, y_ctot, y_coff};
  // dbeta = sum dahat, dgamma = sum dahat * ahat
  if (launch_reduce_rows(partial, partial_rows, 2, 2, dbn, st)) return 1;
  {
    int blocks = (int)((P + 7) / 8);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    sa_bn_conv4_bwd_kernel<<<blocks, kSaThreads, (size_t)(kSaThreads / 32) * (r + 1) * sizeof(float), st>>>(
        dahat, a, mean_rstd, gamma, dbn, (float)(1.0 / (double)P), a3, w4, r, P, buf0, part);
    AIDE_CHECK_LAUNCH();
    if (launch_reduce_rows(part, blocks, r + 1, r + 1, dw4, st)) return 1;      // dw4 [r] then db4
  }
  const int rows = wgrad_rows(P);
  const int chunk = (int)((P + rows - 1) / rows);
  const int n_out_d = r * r * 9 + r;
  const dim3 grid_d(rows, ceil_div(n_out_d, kSaThreads));
  // conv3: dout = da3 (buf0), in = a2
  sa_dilated_bwd_w_kernel<<<grid_d, kSaThreads, 0, st>>>(buf0, a2, r, dilation, N, H, W, chunk, part);
  AIDE_CHECK_LAUNCH();
  if (launch_reduce_rows(part, rows, n_out_d, n_out_d, dw3, st)) return 1;
  sa_dilated_bwd_data_kernel<<<grid_for(P * r), kSaThreads, 0, st>>>(buf0, w3, r, dilation, N, H, W, buf1);
  AIDE_CHECK_LAUNCH();
  // conv2: dout = da2 (buf1), in = a1
  sa_dilated_bwd_w_kernel<<<grid_d, kSaThreads, 0, st>>>(buf1, a1, r, dilation, N, H, W, chunk, part);
  AIDE_CHECK_LAUNCH();
  if (launch_reduce_rows(part, rows, n_out_d, n_out_d, dw2, st)) return 1;
  sa_dilated_bwd_data_kernel<<<grid_for(P * r), kSaThreads, 0, st>>>(buf1, w2, r, dilation, N, H, W, buf0);
  AIDE_CHECK_LAUNCH();
  // conv1: dout = da1 (buf0), in = y
  const int n_out_1 = r * C + r;
  AIDE_DISPATCH_FMT(fmt, (sa_conv1_bwd_w_kernel<FMT><<<dim3(rows, ceil_div(n_out_1, kSaThreads)), kSaThreads, 0, st>>>(
                             buf0, y, r, C, P, chunk, part)));
  AIDE_CHECK_LAUNCH();
  if (launch_reduce_rows(part, rows, n_out_1, n_out_1, dw1, st)) return 1;
  sa_conv1_bwd_data_kernel<<<grid_for(P * (C / 4)), kSaThreads, 0, st>>>(buf0, w1, r, C, P, dy);
  AIDE_CHECK_LAUNCH();
  return 0;
}
