// conv_simt.cu -- exact-fp32 CUDA-core conv3x3 (stride 1, pad 1) forward / dgrad and wgrad on NHWC.
//
// Role: (1) the F32 operand format ("exact" mode, used as on-device cross-check of the tcgen05 path),
// (2) the layers that do not fit an MMA shape (first conv, Cin = 3) in every mode.
// Shared-memory tiled direct convolution: 8x8 output pixels x 64 output channels per CTA,
// 4 pixels x 4 channels per thread, input halo tile and weight slice staged per 16-channel chunk.
#include "common.cuh"

namespace aide {

constexpr int TS = 8;        // spatial tile edge
constexpr int TCO = 64;      // output channels per CTA
constexpr int CK = 16;       // input-channel chunk

__global__ void __launch_bounds__(256)
conv3x3_simt_kernel(const float* __restrict__ x, int x_ctot, int x_coff, int cin, const float* __restrict__ w,
                    const float* __restrict__ bias, float* __restrict__ z, int z_ctot, int z_coff, int cout, int H,
                    int W, int tiles_w, int tiles_h, float* __restrict__ stat_partial) {
  __shared__ float in_s[TS + 2][TS + 2][CK + 1];
  __shared__ __align__(16) float w_s[9][CK][TCO];
  const int tile = blockIdx.x;
  const int tw = tile % tiles_w, th = (tile / tiles_w) % tiles_h, n = tile / (tiles_w * tiles_h);
  const int h0 = th * TS, w0 = tw * TS, co0 = blockIdx.y * TCO;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int py = ty >> 1, px0 = (ty & 1) * 4;
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;

  for (int c0 = 0; c0 < cin; c0 += CK) {
    // stage input halo tile (zero padding outside the image / beyond cin)
    for (int i = threadIdx.x; i < (TS + 2) * (TS + 2) * CK; i += 256) {
      int c = i % CK, pp = i / CK;
      int xx = pp % (TS + 2), yy = pp / (TS + 2);
      int hh = h0 + yy - 1, ww = w0 + xx - 1;
      float v = 0.f;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W && c0 + c < cin)
        v = x[(((size_t)n * H + hh) * W + ww) * x_ctot + x_coff + c0 + c];
      in_s[yy][xx][c] = v;
    }
    // stage weight slice  w[co][tap][ci] -> w_s[tap][c][co]
    for (int i = threadIdx.x; i < 9 * CK * TCO; i += 256) {
      int c = i % CK, tap = (i / CK) % 9, co = i / (CK * 9);
      float v = 0.f;
      if (co0 + co < cout && c0 + c < cin) v = w[((size_t)(co0 + co) * 9 + tap) * cin + c0 + c];
      w_s[tap][c][co] = v;
    }
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int ky = tap / 3, kx = tap % 3;
#pragma unroll 4
      for (int c = 0; c < CK; ++c) {
        float4 b = *reinterpret_cast<const float4*>(&w_s[tap][c][tx * 4]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float a = in_s[py + ky][px0 + j + kx][c];
          acc[j][0] = fmaf(a, b.x, acc[j][0]);
          acc[j][1] = fmaf(a, b.y, acc[j][1]);
          acc[j][2] = fmaf(a, b.z, acc[j][2]);
          acc[j][3] = fmaf(a, b.w, acc[j][3]);
        }
      }
    }
    __syncthreads();
  }
  // epilogue: bias, store, statistics
  const int co = co0 + tx * 4;
  float bs[4] = {0, 0, 0, 0};
  if (bias) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (co + k < cout) bs[k] = bias[co + k];
  }
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  const int hh = h0 + py;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int ww = w0 + px0 + j;
    if (hh < H && ww < W) {
      size_t o = (((size_t)n * H + hh) * W + ww) * z_ctot + z_coff + co;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (co + k < cout) {
          float v = acc[j][k] + bs[k];
          z[o + k] = v;
          s1[k] += v;
          s2[k] += v * v;
        }
      }
    }
  }
  if (stat_partial) {
    float* red = &w_s[0][0][0];  // reuse: [16 ty][64 co][2]
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      red[(ty * TCO + tx * 4 + k) * 2 + 0] = s1[k];
      red[(ty * TCO + tx * 4 + k) * 2 + 1] = s2[k];
    }
    __syncthreads();
    if (threadIdx.x < TCO && co0 + threadIdx.x < cout) {
      float a = 0.f, b = 0.f;
      for (int t = 0; t < 16; ++t) {
        a += red[(t * TCO + threadIdx.x) * 2 + 0];
        b += red[(t * TCO + threadIdx.x) * 2 + 1];
      }
      stat_partial[((size_t)tile * 2 + 0) * cout + co0 + threadIdx.x] = a;
      stat_partial[((size_t)tile * 2 + 1) * cout + co0 + threadIdx.x] = b;
    }
  }
}

// ------------------------------------------------------------------ wgrad
// CTA: 64 co x 64 ci of one tap over one contiguous pixel range; partial -> ws[split][co][tap][ci].
constexpr int WP = 16;  // pixels per smem stage
__global__ void __launch_bounds__(256)
wgrad_simt_kernel(const float* __restrict__ x, int x_ctot, int x_coff, int cin, const float* __restrict__ dz,
                  int cout, int N, int H, int W, int ci_tiles, size_t pix_per_split, float* __restrict__ ws) {
  __shared__ __align__(16) float dz_s[WP][64];
  __shared__ __align__(16) float x_s[WP][64];
  const int co0 = (blockIdx.x / ci_tiles) * 64, ci0 = (blockIdx.x % ci_tiles) * 64;
  const int tap = blockIdx.y, dy = tap / 3 - 1, dx = tap % 3 - 1;
  const size_t npix = (size_t)N * H * W;
  const size_t p_begin = (size_t)blockIdx.z * pix_per_split;
  const size_t p_end = min(npix, p_begin + pix_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // tx -> ci group, ty -> co group
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[j][k] = 0.f;
  for (size_t p0 = p_begin; p0 < p_end; p0 += WP) {
    for (int i = threadIdx.x; i < WP * 64; i += 256) {
      int c = i & 63, pp = i >> 6;
      size_t p = p0 + pp;
      float a = 0.f, b = 0.f;
      if (p < p_end) {
        if (co0 + c < cout) a = dz[p * cout + co0 + c];
        int ww = (int)(p % W), hh = (int)((p / W) % H);
        int h2 = hh + dy, w2 = ww + dx;
        if (ci0 + c < cin && h2 >= 0 && h2 < H && w2 >= 0 && w2 < W)
          b = x[(p + (long long)dy * W + dx) * x_ctot + x_coff + ci0 + c];
      }
      dz_s[pp][c] = a;
      x_s[pp][c] = b;
    }
    __syncthreads();
#pragma unroll
    for (int pp = 0; pp < WP; ++pp) {
      float4 a = *reinterpret_cast<const float4*>(&dz_s[pp][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&x_s[pp][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[j][0] = fmaf(av[j], b.x, acc[j][0]);
        acc[j][1] = fmaf(av[j], b.y, acc[j][1]);
        acc[j][2] = fmaf(av[j], b.z, acc[j][2]);
        acc[j][3] = fmaf(av[j], b.w, acc[j][3]);
      }
    }
    __syncthreads();
  }
  float* out = ws + (size_t)blockIdx.z * cout * 9 * cin;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int co = co0 + ty * 4 + j;
    if (co >= cout) continue;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int ci = ci0 + tx * 4 + k;
      if (ci < cin) out[((size_t)co * 9 + tap) * cin + ci] = acc[j][k];
    }
  }
}

// dw_oihw[co][ci][tap] = sum_split ws[split][...]; layout 0: [co][tap][ci], layout 1: [tap][ci][co].
// The last `job_blocks` blocks of the grid do a posted column sum instead (32 columns x 8 row lanes each, fp64, fixed order).
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int cout, int cin, int layout,
                                    float* __restrict__ dw, float scale, const float* __restrict__ scale_ptr,
                                    ColSumJob job, int job_blocks) {
  const int main_blocks = gridDim.x - job_blocks;
  if ((int)blockIdx.x >= main_blocks) {
    __shared__ double sm[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = (blockIdx.x - main_blocks) * 32 + tx;
    double acc = 0.0;
    if (j < job.cols) {
#pragma unroll 4
      for (int r = ty; r < job.rows; r += 8) acc += (double)job.src[(size_t)r * job.ld + j];
    }
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && j < job.cols) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sm[k][tx];
      job.dst[j] = (float)t;
    }
    return;
  }
  size_t total = (size_t)cout * cin * 9;
  if (scale_ptr) scale *= __ldg(scale_ptr);      // descale of the operand formats (static) x gradient scale (dynamic)
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)main_blocks * blockDim.x) {
    // i indexes the workspace layout (coalesced reads)
    int co, ci, tap;
    if (layout == 0) {
      ci = (int)(i % cin);
      tap = (int)((i / cin) % 9);
      co = (int)(i / ((size_t)cin * 9));
    } else {
      co = (int)(i % cout);
      ci = (int)((i / cout) % cin);
      tap = (int)(i / ((size_t)cout * cin));
    }
    float a = 0.f;
    for (int s = 0; s < splits; ++s) a += ws[(size_t)s * total + i];
    dw[((size_t)co * cin + ci) * 9 + tap] = a * scale;
  }
}

namespace {
thread_local ColSumJob t_job{nullptr, 0, 0, 0, nullptr};
}
void wgrad_reduce_post_job(const ColSumJob& job) { t_job = job; }
bool wgrad_reduce_take_job(ColSumJob* job) {
  if (!t_job.src) return false;
  *job = t_job;
  t_job.src = nullptr;
  return true;
}

int launch_wgrad_reduce(const float* ws, int splits, int cout, int cin, int layout, float* dw, float scale,
                        const float* scale_ptr, cudaStream_t st) {
  size_t total = (size_t)cout * cin * 9;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  ColSumJob job{nullptr, 0, 0, 0, nullptr};
  const int job_blocks = wgrad_reduce_take_job(&job) ? ceil_div(job.cols, 32) : 0;
  wgrad_reduce_kernel<<<blocks + job_blocks, 256, 0, st>>>(ws, splits, cout, cin, layout, dw, scale, scale_ptr, job,
                                                           job_blocks);
  AIDE_CHECK_LAUNCH();
  return 0;
}

int simt_stat_rows(int N, int H, int W) { return N * ceil_div(H, TS) * ceil_div(W, TS); }

int simt_conv3x3(const float* x, int x_ctot, int x_coff, int cin, const float* w, const float* bias, float* z,
                 int z_ctot, int z_coff, int cout, int N, int H, int W, float* stat_partial, cudaStream_t st) {
  int tiles_w = ceil_div(W, TS), tiles_h = ceil_div(H, TS);
  dim3 grid(N * tiles_w * tiles_h, ceil_div(cout, TCO));
  conv3x3_simt_kernel<<<grid, 256, 0, st>>>(x, x_ctot, x_coff, cin, w, bias, z, z_ctot, z_coff, cout, H, W, tiles_w,
                                            tiles_h, stat_partial);
  AIDE_CHECK_LAUNCH();
  return 0;
}

static int simt_wgrad_splits(int cin, int cout, size_t npix) {
  long long tiles = (long long)ceil_div(cout, 64) * ceil_div(cin, 64) * 9;
  long long s = (kNumSMs * 4 + tiles - 1) / tiles;
  long long maxs = (long long)((npix + WP * 8 - 1) / (WP * 8));
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return (int)s;
}

size_t simt_wgrad_workspace_bytes(int cin, int cout, int N, int H, int W) {
  return (size_t)simt_wgrad_splits(cin, cout, (size_t)N * H * W) * cout * 9 * cin * sizeof(float);
}

int simt_wgrad(const float* x, int x_ctot, int x_coff, int cin, const float* dz, int cout, int N, int H, int W,
               float* ws, size_t ws_bytes, float* dw, cudaStream_t st) {
  size_t npix = (size_t)N * H * W;
  int splits = simt_wgrad_splits(cin, cout, npix);
  AIDE_REQUIRE(ws && ws_bytes >= (size_t)splits * cout * 9 * cin * sizeof(float), "conv3x3_wgrad: workspace too small");
  size_t per = (npix + splits - 1) / splits;
  per = (per + WP - 1) / WP * WP;
  int ci_tiles = ceil_div(cin, 64);
  dim3 grid(ceil_div(cout, 64) * ci_tiles, 9, splits);
  wgrad_simt_kernel<<<grid, 256, 0, st>>>(x, x_ctot, x_coff, cin, dz, cout, N, H, W, ci_tiles, per, ws);
  AIDE_CHECK_LAUNCH();
  return launch_wgrad_reduce(ws, splits, cout, cin, 0, dw, 1.0f, nullptr, st);
}

}  // namespace aide
