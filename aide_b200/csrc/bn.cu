// bn.cu -- BatchNorm2d (+ReLU, +MaxPool2d(2,2), +concat-slot write) forward and backward.
//
// All kernels are HBM-bound streaming kernels over NHWC fp32 data: 128-bit loads/stores, 4 channels
// per thread, channel index fastest across threadIdx.x so that a warp reads contiguous 512 B.
// Reductions are two-stage and fixed-order (per-block partial rows -> fp64 column reduction), so
// results are bit-reproducible run to run (no float atomics).
#include "common.cuh"

#include <cstdlib>

namespace aide {

// ------------------------------------------------------------------ column reduction of partial rows
// out[j] = sum_r partial[r*ld + j]  for j < cols, accumulated in fp64 in fixed order.
// block = (32 columns, kRedLanes row lanes).  These launches walk ~1200 partial rows of a few hundred columns: with 8 row
// lanes every thread chained ~150 dependent loads (19-26 us per launch, 130 launches per step); 32 lanes and four loads
// in flight cut the chain to ~10 round trips.
constexpr int kRedLanes = 32;
__global__ void reduce_rows_kernel(const float* __restrict__ partial, int rows, int ld, int cols,
                                   float* __restrict__ out) {
  __shared__ double sm[kRedLanes][33];
  int j = blockIdx.x * 32 + threadIdx.x;
  double acc = 0.0;
  if (j < cols) {
#pragma unroll 4
    for (int r = threadIdx.y; r < rows; r += kRedLanes) acc += (double)partial[(size_t)r * ld + j];
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && j < cols) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kRedLanes; ++k) t += sm[k][threadIdx.x];
    out[j] = (float)t;
  }
}

int launch_reduce_rows(const float* partial, int rows, int ld, int cols, float* out, cudaStream_t st) {
  reduce_rows_kernel<<<ceil_div(cols, 32), dim3(32, kRedLanes), 0, st>>>(partial, rows, ld, cols, out);
  AIDE_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------ forward statistics finalize
// The conv kernels emit one partial row per output tile (thousands of rows at 256x256).  Stage 1 folds every
// chunk of kStatChunk rows into an fp64 sum with many blocks and stores it IN PLACE as a (hi, lo) float pair in
// the first two rows of its own chunk (no other block touches those cells); stage 2 -- the finalize kernel --
// then walks only the folded rows.  Fixed order throughout: bit-reproducible run to run.
// The chunk length depends on the rows of ONE statistics group only (a stacked-batch forward folds every view exactly
// like a separate forward would): 64 rows up to 4096, then rows/64 -- at 16384 rows (8 slices of 256x256) 64 chunks of
// 256 rows per channel group instead of 256 chunks of 64, whose 5120 blocks of 2 loads per thread were pure
// block-scheduling latency (92 us for 42 MB).
constexpr int kStatChunkMin = 64;
__host__ __device__ inline int stat_chunk(int rows) {
  const int c = (rows / 64) & ~31;
  return c < kStatChunkMin ? kStatChunkMin : c > 1024 ? 1024 : c;
}

constexpr int kStatLanes = 8;    // row lanes per block: 256-thread blocks, 8 independent loads per thread on a 64-row chunk
                                 // (1024-thread blocks with 2 loads per thread were bound by block scheduling)

__global__ void bn_stat_fold_kernel(float* __restrict__ partial, int rows, int cols /* = 2*C */, int chunk) {
  partial += (size_t)blockIdx.z * rows * cols;          // statistics group (stacked-batch forward): its own row block
  __shared__ double sm[kStatLanes][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * chunk;
  const int r1 = min(rows, r0 + chunk);
  double acc = 0.0;
  if (j < cols) {
#pragma unroll 8
    for (int r = r0 + threadIdx.y; r < r1; r += kStatLanes) acc += (double)partial[(size_t)r * cols + j];
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && j < cols) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kStatLanes; ++k) t += sm[k][threadIdx.x];
    const float hi = (float)t;
    partial[(size_t)r0 * cols + j] = hi;
    if (r0 + 1 < rows) partial[(size_t)(r0 + 1) * cols + j] = (float)(t - (double)hi);
  }
}

// rows are visited as r = g*row_stride + {0 .. sub-1} for g < groups  (plain: row_stride = 1, sub = 1, groups = rows).
// `sgroups` statistics groups (the views of a stacked-batch forward) are finalised one after the other by the same
// block, so the running-statistics updates chain in group order exactly like `sgroups` separate forward calls.
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int groups, int row_stride, int sub, int rows,
                                   int C, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean,
                                   float* __restrict__ rvar, float momentum, float eps, int training,
                                   float* __restrict__ scale_shift, float* __restrict__ mean_rstd, int sgroups) {
  __shared__ double s1[kStatLanes][33], s2[kStatLanes][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float rm = 0.f, rv = 0.f;
  const bool owner = threadIdx.y == 0 && c < C;
  if (owner && rmean) {
    rm = rmean[c];
    rv = rvar[c];
  }
  for (int sg = 0; sg < sgroups; ++sg) {
    const float* part = partial + (size_t)sg * rows * 2 * C;
    double a = 0.0, b = 0.0;
    if (training && c < C) {
#pragma unroll 4
      for (int g = threadIdx.y; g < groups; g += kStatLanes) {
        for (int k = 0; k < sub; ++k) {
          const int r = g * row_stride + k;
          if (r < rows) {
            a += (double)part[((size_t)r * 2 + 0) * C + c];
            b += (double)part[((size_t)r * 2 + 1) * C + c];
          }
        }
      }
    }
    __syncthreads();                      // previous group's owner is done reading s1/s2
    s1[threadIdx.y][threadIdx.x] = a;
    s2[threadIdx.y][threadIdx.x] = b;
    __syncthreads();
    if (!owner) continue;
    float mean, var;
    if (training) {
      double sa = 0.0, sb = 0.0;
#pragma unroll
      for (int k = 0; k < kStatLanes; ++k) {
        sa += s1[k][threadIdx.x];
        sb += s2[k][threadIdx.x];
      }
      double m = sa / count;
      double v = sb / count - m * m;
      if (v < 0.0) v = 0.0;
      mean = (float)m;
      var = (float)v;
      if (rmean) {
        double unbiased = count > 1.0 ? v * (count / (count - 1.0)) : v;
        rm = (1.f - momentum) * rm + momentum * mean;
        rv = (1.f - momentum) * rv + momentum * (float)unbiased;
      }
    } else {
      mean = rm;
      var = rv;
    }
    float rstd = 1.0f / sqrtf(var + eps);
    float sc = gamma[c] * rstd;
    float* ss = scale_shift + (size_t)sg * 2 * C;
    ss[c] = sc;
    ss[C + c] = beta[c] - mean * sc;
    if (mean_rstd) {
      float* mr = mean_rstd + (size_t)sg * 2 * C;
      mr[c] = mean;
      mr[C + c] = rstd;
    }
  }
  if (owner && training && rmean) {
    rmean[c] = rm;
    rvar[c] = rv;
  }
}

// Fold + finalize in ONE launch (the 112 separate fold launches of a step were pure launch latency): blocks fold their
// chunk exactly like bn_stat_fold_kernel -- a block owns 16 channels x {sum, sum of squares} -- then take a ticket; the
// LAST block of a channel group to arrive finalises those 16 channels for all statistics groups, reading the folded
// (hi, lo) pairs of every chunk in fixed order (bit-reproducible whichever block happens to be last).  tickets: one
// zero-initialised counter per channel group, left at zero again.
__global__ void bn_stat_fold_finalize_kernel(float* __restrict__ partial, int rows, int C, int chunk, int n_chunks, int sgroups,
                                             double count, const float* __restrict__ gamma, const float* __restrict__ beta,
                                             float* __restrict__ rmean, float* __restrict__ rvar, float momentum, float eps,
                                             float* __restrict__ scale_shift, float* __restrict__ mean_rstd,
                                             unsigned int* __restrict__ tickets) {
  __shared__ double sm[kStatLanes][33];
  __shared__ unsigned int s_last;
  const int cols = 2 * C;
  const int cl = threadIdx.x & 15, stat = threadIdx.x >> 4;
  const int c = blockIdx.x * 16 + cl;
  const bool cvalid = c < C;
  const int j = stat * C + c;                                 // column of this thread in a partial row
  {
    float* part = partial + (size_t)blockIdx.z * rows * cols;
    const int r0 = blockIdx.y * chunk;
    const int r1 = min(rows, r0 + chunk);
    double acc = 0.0;
    if (cvalid) {
#pragma unroll 8
      for (int r = r0 + threadIdx.y; r < r1; r += kStatLanes) acc += (double)part[(size_t)r * cols + j];
    }
    sm[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && cvalid) {
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < kStatLanes; ++k) t += sm[k][threadIdx.x];
      const float hi = (float)t;
      part[(size_t)r0 * cols + j] = hi;
      if (r0 + 1 < rows) part[(size_t)(r0 + 1) * cols + j] = (float)(t - (double)hi);
      __threadfence();                       // only the 32 writing threads fence (a block-wide fence costs ~5x the fold)
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    const unsigned int total = (unsigned int)(n_chunks * sgroups);
    s_last = atomicAdd(&tickets[blockIdx.x], 1u) == total - 1u ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // ---- finalize the 16 channels of this block (thread (x, 0), x < 16, owns channel c)
  float rm = 0.f, rv = 0.f;
  const bool owner = threadIdx.y == 0 && stat == 0 && cvalid;
  if (owner && rmean) {
    rm = rmean[c];
    rv = rvar[c];
  }
  for (int sg = 0; sg < sgroups; ++sg) {
    const float* part = partial + (size_t)sg * rows * cols;
    double a = 0.0;
    if (cvalid) {
#pragma unroll 4
      for (int g = threadIdx.y; g < n_chunks; g += kStatLanes) {
        const int r = g * chunk;
        a += (double)__ldcg(part + (size_t)r * cols + j);
        if (r + 1 < rows) a += (double)__ldcg(part + (size_t)(r + 1) * cols + j);
      }
    }
    __syncthreads();
    sm[threadIdx.y][threadIdx.x] = a;
    __syncthreads();
    if (!owner) continue;
    double sa = 0.0, sb = 0.0;
#pragma unroll
    for (int k = 0; k < kStatLanes; ++k) {
      sa += sm[k][cl];
      sb += sm[k][cl + 16];
    }
    const double m = sa / count;
    double v = sb / count - m * m;
    if (v < 0.0) v = 0.0;
    const float mean = (float)m, var = (float)v;
    if (rmean) {
      const double unbiased = count > 1.0 ? v * (count / (count - 1.0)) : v;
      rm = (1.f - momentum) * rm + momentum * mean;
      rv = (1.f - momentum) * rv + momentum * (float)unbiased;
    }
    const float rstd = 1.0f / sqrtf(var + eps);
    const float sc = gamma[c] * rstd;
    float* ss = scale_shift + (size_t)sg * 2 * C;
    ss[c] = sc;
    ss[C + c] = beta[c] - mean * sc;
    if (mean_rstd) {
      float* mr = mean_rstd + (size_t)sg * 2 * C;
      mr[c] = mean;
      mr[C + c] = rstd;
    }
  }
  if (owner && rmean) {
    rmean[c] = rm;
    rvar[c] = rv;
  }
  if (threadIdx.x == 0 && threadIdx.y == 0) tickets[blockIdx.x] = 0u;     // ready for the next launch
}

// ------------------------------------------------------------------ eval mode: scale / shift of MANY BatchNorm layers at once
// y = scale * z + shift with scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale (the same fp32
// operations as bn_finalize_kernel's eval branch).  One launch for every unit of a network (an eval forward used to spend
// one finalize launch per unit); blockIdx.x -> (layer, 256-channel chunk) through the prefix sums.
constexpr int kEvalMaxLayers = 48;
struct EvalSsTable {
  const float* gamma[kEvalMaxLayers];
  const float* beta[kEvalMaxLayers];
  const float* rmean[kEvalMaxLayers];
  const float* rvar[kEvalMaxLayers];
  float* ss[kEvalMaxLayers];
  int C[kEvalMaxLayers], block_end[kEvalMaxLayers];
  int n;
};
__global__ void bn_eval_scale_shift_batch_kernel(const __grid_constant__ EvalSsTable t, float eps) {
  int l = 0;
  while (l + 1 < t.n && (int)blockIdx.x >= t.block_end[l]) ++l;
  const int c = ((int)blockIdx.x - (l ? t.block_end[l - 1] : 0)) * 256 + threadIdx.x;
  const int C = t.C[l];
  if (c >= C) return;
  const float rstd = 1.0f / sqrtf(t.rvar[l][c] + eps);
  const float sc = t.gamma[l][c] * rstd;
  t.ss[l][c] = sc;
  t.ss[l][C + c] = t.beta[l][c] - t.rmean[l][c] * sc;
}

// ------------------------------------------------------------------ forward apply
__device__ __forceinline__ float4 bn_relu4(float4 z, float4 sc, float4 sh) {
  return make_float4(fmaxf(fmaf(z.x, sc.x, sh.x), 0.f), fmaxf(fmaf(z.y, sc.y, sh.y), 0.f),
                     fmaxf(fmaf(z.z, sc.z, sh.z), 0.f), fmaxf(fmaf(z.w, sc.w, sh.w), 0.f));
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}

// blockDim = (cx, ty), grid = (row blocks, channel groups).  threadIdx.x owns 4 channels; whole rows of pixels (or of
// 2x2 windows) are dealt to blocks and threadIdx.y walks along the row, so the inner loop has no integer division
// (the 1-D version spent most of its instruction budget on 64-bit index arithmetic: 0.63 of the HBM roofline).
template <int FMT, bool POOL>
__global__ void bn_relu_apply_kernel(const float* __restrict__ z, int N, int H, int W, int C,
                                     const float* __restrict__ scale_shift, View dst, View pa, View pb,
                                     int imgs_per_group) {
  // scale_shift: [groups][2][C]; image n uses group n / imgs_per_group (one group = plain forward)
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (c >= C) return;
  const int rows_per_img = POOL ? (H >> 1) : H;
  const int total_rows = N * rows_per_img;
  for (int row = blockIdx.x; row < total_rows; row += gridDim.x) {
    const int n = row / rows_per_img, y = row - n * rows_per_img;          // block-uniform
    const float* ss = scale_shift + (size_t)(n / imgs_per_group) * 2 * C;
    const float4 sc = __ldg(reinterpret_cast<const float4*>(ss + c));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(ss + C + c));
    if constexpr (!POOL) {
      const size_t p0 = ((size_t)n * H + y) * W;
      for (int x = threadIdx.y; x < W; x += blockDim.y) {
        const size_t pix = p0 + x;
        const float4 v = *reinterpret_cast<const float4*>(z + pix * C + c);
        st4<FMT>(dst.p0, dst.p1, pix * dst.ctot + dst.coff + c, bn_relu4(v, sc, sh));
      }
    } else {
      const int Wh = W >> 1;
      const size_t prow = ((size_t)n * H + 2 * y) * W;
      const size_t wrow = ((size_t)n * rows_per_img + y) * Wh;
      for (int wx = threadIdx.y; wx < Wh; wx += blockDim.y) {
        const size_t p00 = prow + 2 * wx, p01 = p00 + 1, p10 = p00 + W, p11 = p10 + 1;
        const size_t win = wrow + wx;
        float4 y00 = bn_relu4(*reinterpret_cast<const float4*>(z + p00 * C + c), sc, sh);
        float4 y01 = bn_relu4(*reinterpret_cast<const float4*>(z + p01 * C + c), sc, sh);
        float4 y10 = bn_relu4(*reinterpret_cast<const float4*>(z + p10 * C + c), sc, sh);
        float4 y11 = bn_relu4(*reinterpret_cast<const float4*>(z + p11 * C + c), sc, sh);
        if (dst.p0) {
          st4<FMT>(dst.p0, dst.p1, p00 * dst.ctot + dst.coff + c, y00);
          st4<FMT>(dst.p0, dst.p1, p01 * dst.ctot + dst.coff + c, y01);
          st4<FMT>(dst.p0, dst.p1, p10 * dst.ctot + dst.coff + c, y10);
          st4<FMT>(dst.p0, dst.p1, p11 * dst.ctot + dst.coff + c, y11);
        }
        float4 m = max4(max4(y00, y01), max4(y10, y11));
        if (pa.p0) st4<FMT>(pa.p0, pa.p1, win * pa.ctot + pa.coff + c, m);
        if (pb.p0) st4<FMT>(pb.p0, pb.p1, win * pb.ctot + pb.coff + c, m);
      }
    }
  }
}

// ------------------------------------------------------------------ backward stage 1
struct BwdSrcs {
  const float* dptr[3];
  int dctot[3], dcoff[3], nd;
  const float* pptr[3];
  int pctot[3], pcoff[3], np;
};

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// first-max-wins selection mask for one channel of a 2x2 window (scan order 00,01,10,11; strict >)
__device__ __forceinline__ int argmax4(float a, float b, float c, float d) {
  int k = 0;
  float m = a;
  if (b > m) { m = b; k = 1; }
  if (c > m) { m = c; k = 2; }
  if (d > m) { m = d; k = 3; }
  return k;
}

// blockDim = (cx, ty); grid = (rows, cgroups).  Each thread owns 4 channels and walks 2x2 windows
// (WINDOW: needed when pooled sources exist) or single pixels (any H, W).
template <bool WINDOW>
__global__ void bn_relu_bwd_reduce_kernel(const float* __restrict__ z, const float* __restrict__ scale_shift,
                                          const float* __restrict__ mean_rstd, int N, int H, int W, int C,
                                          BwdSrcs s, float* __restrict__ g, float* __restrict__ partial,
                                          unsigned int* __restrict__ gmax_bits) {
  extern __shared__ __align__(16) float smem[];  // [ty][cx*8]
  float gmax = 0.f;
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  const bool cvalid = c < C;
  // WINDOW: the four pixels of a 2x2 window per iteration; otherwise TWO pixels one grid stride apart (independent
  // loads in flight -- the one-pixel loop was latency bound; the per-thread accumulation order is unchanged)
  constexpr int KP = WINDOW ? 4 : 2;
  const int Hh = WINDOW ? (H >> 1) : H, Wh = WINDOW ? (W >> 1) : W;
  const size_t nwin = (size_t)N * Hh * Wh;
  float sg[4] = {0, 0, 0, 0}, sgx[4] = {0, 0, 0, 0};
  if (cvalid) {
    const float4 sc = *reinterpret_cast<const float4*>(scale_shift + c);
    const float4 sh = *reinterpret_cast<const float4*>(scale_shift + C + c);
    const float4 mu = *reinterpret_cast<const float4*>(mean_rstd + c);
    const float4 rs = *reinterpret_cast<const float4*>(mean_rstd + C + c);
    const size_t stride = (size_t)gridDim.x * blockDim.y;
    for (size_t win = (size_t)blockIdx.x * blockDim.y + threadIdx.y; win < nwin; win += (WINDOW ? 1 : 2) * stride) {
      size_t p[KP];
      bool ok[KP];
      if constexpr (WINDOW) {
        int wx = (int)(win % Wh);
        int hy = (int)((win / Wh) % Hh);
        int n = (int)(win / ((size_t)Wh * Hh));
        p[0] = ((size_t)n * H + 2 * hy) * W + 2 * wx;
        p[1] = p[0] + 1;
        p[2] = p[0] + W;
        p[3] = p[2] + 1;
        ok[0] = ok[1] = ok[2] = ok[3] = true;
      } else {
        p[0] = win;
        p[1] = win + stride;
        ok[0] = true;
        ok[1] = p[1] < nwin;
        if (!ok[1]) p[1] = win;                  // harmless duplicate address, masked below
      }
      float4 zz[KP], yy[KP], gg[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        zz[k] = *reinterpret_cast<const float4*>(z + p[k] * C + c);
        yy[k] = bn_relu4(zz[k], sc, sh);
        gg[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int d = 0; d < s.nd; ++d) {
#pragma unroll
        for (int k = 0; k < KP; ++k)
          gg[k] = add4(gg[k], *reinterpret_cast<const float4*>(s.dptr[d] + p[k] * s.dctot[d] + s.dcoff[d] + c));
      }
      if constexpr (WINDOW) {
        float4 pg = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int d = 0; d < s.np; ++d)
          pg = add4(pg, *reinterpret_cast<const float4*>(s.pptr[d] + win * s.pctot[d] + s.pcoff[d] + c));
        int kx = argmax4(yy[0].x, yy[1].x, yy[2].x, yy[3].x);
        int ky = argmax4(yy[0].y, yy[1].y, yy[2].y, yy[3].y);
        int kz = argmax4(yy[0].z, yy[1].z, yy[2].z, yy[3].z);
        int kw = argmax4(yy[0].w, yy[1].w, yy[2].w, yy[3].w);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (kx == k) gg[k].x += pg.x;
          if (ky == k) gg[k].y += pg.y;
          if (kz == k) gg[k].z += pg.z;
          if (kw == k) gg[k].w += pg.w;
        }
      }
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        if (!ok[k]) continue;
        float4 o;
        o.x = yy[k].x > 0.f ? gg[k].x : 0.f;
        o.y = yy[k].y > 0.f ? gg[k].y : 0.f;
        o.z = yy[k].z > 0.f ? gg[k].z : 0.f;
        o.w = yy[k].w > 0.f ? gg[k].w : 0.f;
        *reinterpret_cast<float4*>(g + p[k] * C + c) = o;
        gmax = fmaxf(gmax, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
        sg[0] += o.x; sg[1] += o.y; sg[2] += o.z; sg[3] += o.w;
        sgx[0] += o.x * ((zz[k].x - mu.x) * rs.x);
        sgx[1] += o.y * ((zz[k].y - mu.y) * rs.y);
        sgx[2] += o.z * ((zz[k].z - mu.z) * rs.z);
        sgx[3] += o.w * ((zz[k].w - mu.w) * rs.w);
      }
    }
  }
  // block reduction over threadIdx.y (fixed order)
  float* row = smem + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * 8;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    row[k] = sg[k];
    row[4 + k] = sgx[k];
  }
  __syncthreads();
  if (threadIdx.y == 0 && cvalid) {
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int t = 0; t < (int)blockDim.y; ++t) {
      const float* r = smem + ((size_t)t * blockDim.x + threadIdx.x) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) a[k] += r[k];
    }
    float* out = partial + (size_t)blockIdx.x * 2 * C;
    *reinterpret_cast<float4*>(out + c) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(out + C + c) = make_float4(a[4], a[5], a[6], a[7]);
  }
  if (gmax_bits) {   // max |g| of the whole tensor: order-independent, so the atomic keeps the result deterministic
    gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, 16));
    gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, 8));
    gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, 4));
    gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, 2));
    gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, 1));
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && gmax > 0.f) atomicMax(gmax_bits, __float_as_uint(gmax));
  }
}

// F16X2 gradients: choose the power-of-two scale s of dZ from a bound on max|dz|,
//   |dz| <= |gamma*rstd| * (|g| + |mean g| + |xhat| * |mean g*xhat|),  |xhat| taken as <= 16,
// so that the bound lands at 2^10 (64x headroom to the fp16 maximum; conversions saturate beyond it, and the two
// fp16 planes keep 22 significant bits down to 2^-24 of the bound).  scale_out = {s, 1/s}.  Resets *gmax_bits to 0
// (the reduce stage of the next unit accumulates into it with atomicMax).
__device__ __forceinline__ void dz_scale_block(unsigned int* gmax_bits, const float* mean_rstd, const float* gamma,
                                               const float* sums, float inv_count, int C, float* scale_out, float* red) {
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
  const float gmax = __uint_as_float(__ldcg(gmax_bits));
  float b = 0.f;
  for (int c = tid; c < C; c += nthr) {
    const float k0 = fabsf(gamma[c] * mean_rstd[C + c]);
    b = fmaxf(b, k0 * (gmax + fabsf(__ldcg(sums + c) * inv_count) + 16.f * fabsf(__ldcg(sums + C + c) * inv_count)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
  if ((tid & 31) == 0) red[tid >> 5] = b;
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < (nthr >> 5); ++i) b = fmaxf(b, red[i]);
    int e = 0;
    if (b > 0.f && isfinite(b)) {
      frexpf(b, &e);                 // b = m * 2^e, m in [0.5, 1)
      e = 10 - e;
      e = e < -100 ? -100 : e > 100 ? 100 : e;
    }
    scale_out[0] = ldexpf(1.f, e);
    scale_out[1] = ldexpf(1.f, -e);
    *gmax_bits = 0u;
  }
}

__global__ void dz_scale_kernel(unsigned int* __restrict__ gmax_bits, const float* __restrict__ mean_rstd,
                                const float* __restrict__ gamma, const float* __restrict__ sums, float inv_count, int C,
                                float* __restrict__ scale_out) {
  __shared__ float red[32];
  dz_scale_block(gmax_bits, mean_rstd, gamma, sums, inv_count, C, scale_out, red);
}

// Column reduction of the backward partial rows (reduce_rows_kernel) whose LAST block (ticket) also derives the dZ scale:
// one launch instead of two per unit.  block = (32, kRedLanes); ticket: zero on entry, left zero.
__global__ void bn_bwd_sums_scale_kernel(const float* __restrict__ partial, int rows, int C, float* __restrict__ sums,
                                         unsigned int* __restrict__ gmax_bits, const float* __restrict__ mean_rstd,
                                         const float* __restrict__ gamma, float inv_count, float* __restrict__ scale_out,
                                         unsigned int* __restrict__ ticket) {
  __shared__ double sm[kRedLanes][33];
  __shared__ float red[32];
  __shared__ unsigned int s_last;
  const int cols = 2 * C;
  const int j = blockIdx.x * 32 + threadIdx.x;
  double acc = 0.0;
  if (j < cols) {
#pragma unroll 4
    for (int r = threadIdx.y; r < rows; r += kRedLanes) acc += (double)partial[(size_t)r * cols + j];
  }
  sm[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && j < cols) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kRedLanes; ++k) t += sm[k][threadIdx.x];
    sums[j] = (float)t;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1u ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  dz_scale_block(gmax_bits, mean_rstd, gamma, sums, inv_count, C, scale_out, red);
  if (threadIdx.x == 0 && threadIdx.y == 0) *ticket = 0u;
}

// ------------------------------------------------------------------ backward stage 2
template <int FMT>
__global__ void bn_relu_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ z,
                                         const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                                         const float* __restrict__ sums /*[2][C]: sum g, sum g*xhat*/, float inv_count,
                                         size_t npix, int C, void* dz0, void* dz1, float* __restrict__ partial2,
                                         const float* __restrict__ dz_scale, unsigned int* __restrict__ tickets,
                                         float* __restrict__ dbias) {
  extern __shared__ __align__(16) float smem[];  // [ty][cx*4]
  __shared__ unsigned int s_last;
  // F16X2 planes store dz * s / 2^8 through st4 (which multiplies by the activation scale 2^8): pass dz * s * 2^-8
  float fs = 1.f;
  if constexpr (FMT == AIDE_FMT_F16X2) fs = __ldg(dz_scale) * (1.0f / kF16ActScale);
  const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  const bool cvalid = c < C;
  float sd[4] = {0, 0, 0, 0};
  if (cvalid) {
    const float4 mu = *reinterpret_cast<const float4*>(mean_rstd + c);
    const float4 rs = *reinterpret_cast<const float4*>(mean_rstd + C + c);
    const float4 ga = make_float4(gamma[c], gamma[c + 1], gamma[c + 2], gamma[c + 3]);
    const float4 s1 = *reinterpret_cast<const float4*>(sums + c);
    const float4 s2 = *reinterpret_cast<const float4*>(sums + C + c);
    const float4 k0 = make_float4(ga.x * rs.x, ga.y * rs.y, ga.z * rs.z, ga.w * rs.w);
    const float4 m1 = make_float4(s1.x * inv_count, s1.y * inv_count, s1.z * inv_count, s1.w * inv_count);
    const float4 m2 = make_float4(s2.x * inv_count, s2.y * inv_count, s2.z * inv_count, s2.w * inv_count);
    // two pixels per iteration: four independent 128-bit loads in flight per thread (the one-pixel loop ran at 0.56 of
    // the copy peak, latency bound); same accumulation order as the sequential loop
    const size_t stride = (size_t)gridDim.x * blockDim.y;
    for (size_t p = (size_t)blockIdx.x * blockDim.y + threadIdx.y; p < npix; p += 2 * stride) {
      const size_t q = p + stride;
      const bool two = q < npix;
      const float4 gv0 = *reinterpret_cast<const float4*>(g + p * C + c);
      const float4 zv0 = *reinterpret_cast<const float4*>(z + p * C + c);
      const float4 gv1 = two ? *reinterpret_cast<const float4*>(g + q * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 zv1 = two ? *reinterpret_cast<const float4*>(z + q * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
        const float4 gv = u ? gv1 : gv0, zv = u ? zv1 : zv0;
        const size_t pp = u ? q : p;
        float4 d;
        d.x = k0.x * (gv.x - m1.x - ((zv.x - mu.x) * rs.x) * m2.x);
        d.y = k0.y * (gv.y - m1.y - ((zv.y - mu.y) * rs.y) * m2.y);
        d.z = k0.z * (gv.z - m1.z - ((zv.z - mu.z) * rs.z) * m2.z);
        d.w = k0.w * (gv.w - m1.w - ((zv.w - mu.w) * rs.w) * m2.w);
        if constexpr (FMT == AIDE_FMT_F16X2)
          st4<FMT>(dz0, dz1, pp * C + c, make_float4(d.x * fs, d.y * fs, d.z * fs, d.w * fs));
        else
          st4<FMT>(dz0, dz1, pp * C + c, d);
        sd[0] += d.x; sd[1] += d.y; sd[2] += d.z; sd[3] += d.w;
      }
    }
  }
  float* row = smem + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) row[k] = sd[k];
  __syncthreads();
  if (threadIdx.y == 0 && cvalid) {
    float a[4] = {0, 0, 0, 0};
    for (int t = 0; t < (int)blockDim.y; ++t) {
      const float* r = smem + ((size_t)t * blockDim.x + threadIdx.x) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] += r[k];
    }
    *reinterpret_cast<float4*>(partial2 + (size_t)blockIdx.x * C + c) = make_float4(a[0], a[1], a[2], a[3]);
    if (tickets) __threadfence();            // writers only
  }
  if (!tickets) return;
  // dbias_conv = column sums of partial2, folded by the LAST block of this channel group (fixed order, fp64): saves
  // the separate reduce launch.  tickets[blockIdx.y]: zero on entry, left zero.
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) s_last = atomicAdd(&tickets[blockIdx.y], 1u) == gridDim.x - 1u ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double* dsm = reinterpret_cast<double*>(smem);       // [ty][cx*4] doubles: the launch reserves 8 bytes per slot
  double t4[4] = {0, 0, 0, 0};
  if (cvalid) {
    for (int r = threadIdx.y; r < (int)gridDim.x; r += blockDim.y) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(partial2 + (size_t)r * C + c));
      t4[0] += (double)v.x; t4[1] += (double)v.y; t4[2] += (double)v.z; t4[3] += (double)v.w;
    }
  }
  __syncthreads();
  double* drow = dsm + ((size_t)threadIdx.y * blockDim.x + threadIdx.x) * 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) drow[k] = t4[k];
  __syncthreads();
  if (threadIdx.y == 0 && cvalid) {
    double a[4] = {0, 0, 0, 0};
    for (int t = 0; t < (int)blockDim.y; ++t) {
      const double* r = dsm + ((size_t)t * blockDim.x + threadIdx.x) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] += r[k];
    }
    *reinterpret_cast<float4*>(dbias + c) = make_float4((float)a[0], (float)a[1], (float)a[2], (float)a[3]);
  }
  if (threadIdx.x == 0 && threadIdx.y == 0) tickets[blockIdx.y] = 0u;
}

// shared geometry for the two backward kernels
struct BwdGeom {
  int cx, ty, cgroups, rows;
};
static BwdGeom bwd_geom(int N, int H, int W, int C) {
  BwdGeom gm;
  int c4 = C / 4;
  gm.cx = c4 < 64 ? c4 : 64;
  gm.ty = 256 / gm.cx;
  if (gm.ty < 1) gm.ty = 1;
  gm.cgroups = ceil_div(c4, gm.cx);
  long long nwin = ((long long)N * H * W + 3) / 4;
  long long want = (nwin + gm.ty * 4 - 1) / (gm.ty * 4);  // >= 4 windows per thread
  long long cap = (long long)kNumSMs * 8 / gm.cgroups;
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  gm.rows = (int)want;
  return gm;
}

}  // namespace aide

using namespace aide;

extern "C" int aide_bn_finalize(float* stat_partial, int rows, int C, double count, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, float momentum,
                                float eps, int training, float* scale_shift, float* mean_rstd, void* stream) {
  return aide_bn_finalize_grouped(stat_partial, rows, 1, C, count, gamma, beta, running_mean, running_var, momentum, eps,
                                  training, scale_shift, mean_rstd, nullptr, stream);
}

extern "C" int aide_bn_ticket_slots(int C) { return ceil_div(C, 16); }

extern "C" int aide_bn_finalize_grouped(float* stat_partial, int rows, int sgroups, int C, double count,
                                        const float* gamma, const float* beta, float* running_mean, float* running_var,
                                        float momentum, float eps, int training, float* scale_shift, float* mean_rstd,
                                        unsigned int* tickets, void* stream) {
  AIDE_REQUIRE(C > 0 && gamma && beta && scale_shift && sgroups >= 1, "bn_finalize: bad arguments");
  AIDE_REQUIRE(training ? (stat_partial && rows > 0 && count > 0) : (running_mean && running_var),
               "bn_finalize: missing statistics input");
  dim3 block(32, kStatLanes), grid(ceil_div(C, 32));
  int groups = rows, row_stride = 1, sub = 1;
  const int chunk = stat_chunk(rows);
  if (training && rows > 2 * kStatChunkMin) {   // many tiles: fold chunks in parallel first (in place)
    groups = ceil_div(rows, chunk);
    if (tickets) {                           // ... and let the last block of each channel group finalize: one launch
      bn_stat_fold_finalize_kernel<<<dim3(ceil_div(C, 16), groups, sgroups), block, 0, as_stream(stream)>>>(
          stat_partial, rows, C, chunk, groups, sgroups, count, gamma, beta, running_mean, running_var, momentum, eps,
          scale_shift, mean_rstd, tickets);
      AIDE_CHECK_LAUNCH();
      return 0;
    }
    bn_stat_fold_kernel<<<dim3(ceil_div(2 * C, 32), groups, sgroups), block, 0, as_stream(stream)>>>(stat_partial, rows,
                                                                                                      2 * C, chunk);
    AIDE_CHECK_LAUNCH();
    row_stride = chunk;
    sub = 2;
  }
  bn_finalize_kernel<<<grid, block, 0, as_stream(stream)>>>(stat_partial, groups, row_stride, sub, rows, C, count,
                                                            gamma, beta, running_mean, running_var, momentum, eps,
                                                            training, scale_shift, mean_rstd, sgroups);
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_bn_relu_apply(int fmt, const float* z, int N, int H, int W, int C, const float* scale_shift,
                                  void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff, void* poolA_p0,
                                  void* poolA_p1, int poolA_ctot, int poolA_coff, void* poolB_p0, void* poolB_p1,
                                  int poolB_ctot, int poolB_coff, void* stream) {
  return aide_bn_relu_apply_grouped(fmt, z, N, N, H, W, C, scale_shift, dst_p0, dst_p1, dst_ctot, dst_coff, poolA_p0,
                                    poolA_p1, poolA_ctot, poolA_coff, poolB_p0, poolB_p1, poolB_ctot, poolB_coff, stream);
}

extern "C" int aide_bn_relu_apply_grouped(int fmt, const float* z, int N, int imgs_per_group, int H, int W, int C,
                                          const float* scale_shift, void* dst_p0, void* dst_p1, int dst_ctot,
                                          int dst_coff, void* poolA_p0, void* poolA_p1, int poolA_ctot, int poolA_coff,
                                          void* poolB_p0, void* poolB_p1, int poolB_ctot, int poolB_coff, void* stream) {
  AIDE_REQUIRE(z && scale_shift && C % 4 == 0 && N > 0 && imgs_per_group > 0 && N % imgs_per_group == 0,
               "bn_relu_apply: bad arguments (C %% 4 must be 0, N a multiple of the group size)");
  const bool pool = poolA_p0 || poolB_p0;
  AIDE_REQUIRE(pool || dst_p0, "bn_relu_apply: no destination");
  AIDE_REQUIRE(!pool || (H % 2 == 0 && W % 2 == 0), "bn_relu_apply: pooling needs even H, W");
  AIDE_REQUIRE((dst_coff % 4 == 0) && (dst_ctot % 4 == 0) && (poolA_coff % 4 == 0) && (poolB_coff % 4 == 0),
               "bn_relu_apply: channel offsets must be multiples of 4");
  View dst{dst_p0, dst_p1, dst_ctot, dst_coff}, pa{poolA_p0, poolA_p1, poolA_ctot, poolA_coff},
      pb{poolB_p0, poolB_p1, poolB_ctot, poolB_coff};
  const int c4 = C / 4;
  int cx = c4 < 32 ? c4 : 32;                     // threads across channels (x 4 channels each)
  while (cx & (cx - 1)) cx &= cx - 1;             // power of two
  const dim3 block(cx, 256 / cx);
  const int total_rows = N * (pool ? H / 2 : H);
  int rb = total_rows;
  const int cgroups = ceil_div(c4, cx);
  const int cap = kNumSMs * 16 / cgroups > 0 ? kNumSMs * 16 / cgroups : 1;
  if (rb > cap) rb = cap;
  const dim3 grid(rb, cgroups);
  if (pool) {
    AIDE_DISPATCH_FMT(fmt, (bn_relu_apply_kernel<FMT, true><<<grid, block, 0, as_stream(stream)>>>(
                               z, N, H, W, C, scale_shift, dst, pa, pb, imgs_per_group)));
  } else {
    AIDE_DISPATCH_FMT(fmt, (bn_relu_apply_kernel<FMT, false><<<grid, block, 0, as_stream(stream)>>>(
                               z, N, H, W, C, scale_shift, dst, pa, pb, imgs_per_group)));
  }
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_bn_bwd_rows(int N, int H, int W, int C) { return bwd_geom(N, H, W, C).rows; }

extern "C" int aide_bn_relu_bwd_reduce(const float* z, const float* scale_shift, const float* mean_rstd, int N,
                                       int H, int W, int C, const float* const* direct_ptr, const int* direct_ctot,
                                       const int* direct_coff, int n_direct, const float* const* pool_ptr,
                                       const int* pool_ctot, const int* pool_coff, int n_pool, float* g,
                                       float* partial, float* gmax, void* stream) {
  AIDE_REQUIRE(z && scale_shift && mean_rstd && g && partial, "bn_relu_bwd_reduce: null argument");
  AIDE_REQUIRE(C % 4 == 0, "bn_relu_bwd_reduce: need C%%4==0");
  AIDE_REQUIRE(n_pool == 0 || (H % 2 == 0 && W % 2 == 0), "bn_relu_bwd_reduce: pooled sources need even H,W");
  AIDE_REQUIRE(n_direct >= 0 && n_direct <= 3 && n_pool >= 0 && n_pool <= 3 && n_direct + n_pool > 0,
               "bn_relu_bwd_reduce: up to 3 direct and 3 pooled gradient sources supported (at least one)");
  BwdSrcs s{};
  s.nd = n_direct;
  s.np = n_pool;
  for (int i = 0; i < n_direct; ++i) {
    s.dptr[i] = direct_ptr[i];
    s.dctot[i] = direct_ctot[i];
    s.dcoff[i] = direct_coff[i];
    AIDE_REQUIRE(s.dptr[i] && s.dctot[i] % 4 == 0 && s.dcoff[i] % 4 == 0, "bn_relu_bwd_reduce: bad direct source");
  }
  for (int i = 0; i < n_pool; ++i) {
    s.pptr[i] = pool_ptr[i];
    s.pctot[i] = pool_ctot[i];
    s.pcoff[i] = pool_coff[i];
    AIDE_REQUIRE(s.pptr[i] && s.pctot[i] % 4 == 0 && s.pcoff[i] % 4 == 0, "bn_relu_bwd_reduce: bad pooled source");
  }
  BwdGeom gm = bwd_geom(N, H, W, C);
  dim3 block(gm.cx, gm.ty), grid(gm.rows, gm.cgroups);
  size_t smem = (size_t)gm.cx * gm.ty * 8 * sizeof(float);
  unsigned int* gbits = reinterpret_cast<unsigned int*>(gmax);     // zero on entry (aide_bn_relu_bwd_apply re-zeroes it)
  if (n_pool > 0)
    bn_relu_bwd_reduce_kernel<true><<<grid, block, smem, as_stream(stream)>>>(z, scale_shift, mean_rstd, N, H, W, C,
                                                                              s, g, partial, gbits);
  else
    bn_relu_bwd_reduce_kernel<false><<<grid, block, smem, as_stream(stream)>>>(z, scale_shift, mean_rstd, N, H, W, C,
                                                                               s, g, partial, gbits);
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_bn_relu_bwd_apply(int fmt, const float* g, const float* z, const float* mean_rstd,
                                      const float* gamma, const float* partial, int rows, int N, int H, int W, int C,
                                      void* dz_p0, void* dz_p1, float* dgamma, float* dbeta, float* dbias_conv,
                                      float* partial2, float* gmax, float* dz_scale, unsigned int* tickets,
                                      void* stream) {
  AIDE_REQUIRE(g && z && mean_rstd && gamma && partial && dz_p0 && dgamma && dbeta && partial2,
               "bn_relu_bwd_apply: null argument");
  AIDE_REQUIRE(dbeta + C == dgamma, "bn_relu_bwd_apply: dbeta/dgamma must be adjacent ([2][C] buffer: dbeta, dgamma)");
  BwdGeom gm = bwd_geom(N, H, W, C);
  AIDE_REQUIRE(rows == gm.rows, "bn_relu_bwd_apply: rows mismatch (%d vs %d)", rows, gm.rows);
  AIDE_REQUIRE(!tickets || gm.cgroups < 15, "bn_relu_bwd_apply: too many channel groups for the ticket block");
  cudaStream_t st = as_stream(stream);
  size_t npix = (size_t)N * H * W;
  float inv = (float)(1.0 / (double)npix);
  unsigned int* gbits = reinterpret_cast<unsigned int*>(gmax);
  // sums[0..C) = sum g (= dbeta), sums[C..2C) = sum g*xhat (= dgamma)
  if (fmt == AIDE_FMT_F16X2) {
    AIDE_REQUIRE(gmax && dz_scale, "bn_relu_bwd_apply: F16X2 needs gmax (from bn_relu_bwd_reduce) and dz_scale[2]");
    if (tickets) {
      bn_bwd_sums_scale_kernel<<<ceil_div(2 * C, 32), dim3(32, kRedLanes), 0, st>>>(partial, rows, C, dbeta, gbits, mean_rstd, gamma,
                                                                            inv, dz_scale, tickets);
      AIDE_CHECK_LAUNCH();
    } else {
      reduce_rows_kernel<<<ceil_div(2 * C, 32), dim3(32, kRedLanes), 0, st>>>(partial, rows, 2 * C, 2 * C, dbeta);
      AIDE_CHECK_LAUNCH();
      dz_scale_kernel<<<1, 256, 0, st>>>(gbits, mean_rstd, gamma, dbeta, inv, C, dz_scale);
      AIDE_CHECK_LAUNCH();
    }
  } else {
    reduce_rows_kernel<<<ceil_div(2 * C, 32), dim3(32, kRedLanes), 0, st>>>(partial, rows, 2 * C, 2 * C, dbeta);
    AIDE_CHECK_LAUNCH();
  }
  dim3 block(gm.cx, gm.ty), grid(gm.rows, gm.cgroups);
  // Folding dbias_conv in the apply kernel's last block is OFF by default: one block walking ~1 200 partial rows at the
  // end of the launch cost more (0.46 of the copy peak for the whole kernel, tools/hbm_probe.py) than the separate
  // column-reduce launch it saved (0.56).  AIDE_BN_FUSE_DBIAS=1 turns it back on.
  static const bool fuse_dbias_env = []() { const char* e = std::getenv("AIDE_BN_FUSE_DBIAS"); return e && *e == '1'; }();
  const bool fuse_dbias = tickets && dbias_conv && fuse_dbias_env;
  size_t smem = (size_t)gm.cx * gm.ty * 4 * (fuse_dbias ? sizeof(double) : sizeof(float));
  AIDE_DISPATCH_FMT(fmt, (bn_relu_bwd_apply_kernel<FMT><<<grid, block, smem, st>>>(
                             g, z, mean_rstd, gamma, dbeta, inv, npix, C, dz_p0, dz_p1, partial2, dz_scale,
                             fuse_dbias ? tickets + 1 : nullptr, dbias_conv)));
  AIDE_CHECK_LAUNCH();
  if (dbias_conv && !fuse_dbias) {
    reduce_rows_kernel<<<ceil_div(C, 32), dim3(32, kRedLanes), 0, st>>>(partial2, rows, C, C, dbias_conv);
    AIDE_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int aide_bn_eval_scale_shift_batch(int n_layers, const float* const* gamma, const float* const* beta,
                                              const float* const* running_mean, const float* const* running_var,
                                              const int* C, float* const* scale_shift, float eps, void* stream) {
  AIDE_REQUIRE(n_layers >= 1 && gamma && beta && running_mean && running_var && C && scale_shift,
               "bn_eval_scale_shift_batch: bad arguments");
  for (int base = 0; base < n_layers; base += kEvalMaxLayers) {
    EvalSsTable t{};
    t.n = n_layers - base < kEvalMaxLayers ? n_layers - base : kEvalMaxLayers;
    int blocks = 0;
    for (int i = 0; i < t.n; ++i) {
      const int l = base + i;
      AIDE_REQUIRE(gamma[l] && beta[l] && running_mean[l] && running_var[l] && scale_shift[l] && C[l] > 0,
                   "bn_eval_scale_shift_batch: layer %d: null argument", l);
      t.gamma[i] = gamma[l]; t.beta[i] = beta[l]; t.rmean[i] = running_mean[l]; t.rvar[i] = running_var[l];
      t.ss[i] = scale_shift[l]; t.C[i] = C[l];
      blocks += ceil_div(C[l], 256);
      t.block_end[i] = blocks;
    }
    bn_eval_scale_shift_batch_kernel<<<blocks, 256, 0, as_stream(stream)>>>(t, eps);
    AIDE_CHECK_LAUNCH();
  }
  return 0;
}
