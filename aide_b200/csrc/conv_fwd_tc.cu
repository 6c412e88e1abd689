// conv_fwd_tc.cu -- conv3x3 (stride 1, pad 1) forward / dgrad: persistent, im2col-free implicit GEMM on tcgen05.
//
//   D[128 pixels, BN cout] = sum_{tap, c-slice} A_tap[128 pixels, kc] * W_tap[BN, kc]^T       (both operands K-major)
//
//   * one CTA per SM walks (pixel tile, cout tile) pairs in a fixed round-robin order (persistent grid)
//   * warp 0: TMA producer.  A_tap is ONE 4-D box (kc channels x TW x TH pixels x 1 image) of the NHWC activation
//     view loaded at spatial offset (dy, dx); the convolution's zero padding is TMA's out-of-bounds zero fill.
//   * warp 1: TMEM owner + MMA issuer (one elected lane).  Accumulators live in TMEM and are DOUBLE-BUFFERED:
//     the MMAs of tile i+1 run while the epilogue warps drain tile i.
//   * warps 2..5: epilogue.  tcgen05.ld -> fp32 registers (one output pixel per thread) -> scale, +bias ->
//     128-bit global stores, and the BatchNorm partial statistics of the tile (sum z, sum z^2 per channel) reduced
//     across the 32 pixel rows of each warp with a halving shuffle butterfly (31 shuffles per 32 columns).
//
//   operand kinds
//     BF16   ("fast")   one kind::f16 MMA per 32-byte K step.
//     TF32X2 ("parity") hi = rn_tf32(x), lo = x - hi; kind::tf32 MMAs lo*hi + hi*lo + hi*hi.
//     F16X2  ("parity") hi/lo fp16 planes of the power-of-two pre-scaled tensor (common.cuh); kind::f16 MMAs
//                       lo*hi + hi*lo + hi*hi: same 22-bit products at twice the tensor rate, half the bytes.
//   tcgen05.mma truncates when it adds into the accumulator, so one chain's error grows linearly with its length
//   (tools/accum_probe.py); long reductions are spread over `nacc` accumulators that the epilogue adds in fp32 RN.
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace aide {

using namespace ptx;

int act_tmap(CUtensorMap* m, int dtype, const void* plane, int ctot, int coff, int C, int N, int H, int W, int box_c,
             int box_w, int box_h, int swizzle_bytes);
int mat_tmap(CUtensorMap* m, int dtype, const void* plane, int rows, int kdim, int box_k, int box_rows,
             int swizzle_bytes);
void pick_tile(int H, int W, int npix, int* TW, int* TH);
int ilog2(int v);

enum { K_TF32X2 = 0, K_BF16 = 1, K_F16X2 = 2 };
constexpr int kFwdThreads = 192;

struct FwdParams {
  CUtensorMap tmA0, tmA1, tmB0, tmB1;
  float* z;
  const float* bias;
  float* stat_partial;
  const float* out_scale_ptr;   // optional device scalar multiplied into the result (dynamic gradient scale)
  float out_scale;              // static power-of-two descale of the operand formats
  int z_ctot, z_coff, cout, cin, H, W;
  int TW, TH, tw_log2, tiles_w, tiles_h, n_tiles, total_tiles;
  int BN, kc, n_cchunks, row_bytes, stages, nacc, nbuf, tmem_cols;
  int a_plane_bytes, b_plane_bytes, stage_bytes, data_bytes;
};

// sum over the 32 lanes of v[j] for each of the 32 columns j; lane l ends up with the total of column l in v[0]
__device__ __forceinline__ float column_sums_32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; ++j) {
      const float send = up ? v[j] : v[j + s];
      const float keep = up ? v[j + s] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

template <int KIND>
__global__ void __launch_bounds__(kFwdThreads, 1) conv3x3_fwd_tc_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr bool TF32 = KIND == K_TF32X2;
  constexpr int NPL = KIND == K_BF16 ? 1 : 2;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int S = p.stages;
  const uint32_t bar_base = base + p.data_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * S + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * S + 2 + b); };
  const uint32_t slot_addr = bar_base + 8u * (2 * S + 4);
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + p.data_bytes + 8 * (2 * S + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = 9 * p.n_cchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmB0);
    if (NPL == 2) {
      tma_prefetch_desc(&p.tmA1);
      tma_prefetch_desc(&p.tmB1);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);   // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(slot_addr, (uint32_t)p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const uint32_t tx_bytes = NPL * (p.a_plane_bytes + p.b_plane_bytes);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
        const int tw_i = m_tile % p.tiles_w, th_i = (m_tile / p.tiles_w) % p.tiles_h;
        const int n_img = m_tile / (p.tiles_w * p.tiles_h);
        const int h0 = th_i * p.TH, w0 = tw_i * p.TW, n0 = n_tile * p.BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % S, ph = (it / S) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_arrive_expect_tx(full_bar(s), tx_bytes);
          const int tap = kb / p.n_cchunks, cc = kb - tap * p.n_cchunks;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          const uint32_t a_dst = base + s * p.stage_bytes;
          const uint32_t b_dst = a_dst + NPL * p.a_plane_bytes;
          tma_load_4d(a_dst, &p.tmA0, full_bar(s), cc * p.kc, w0 + dx, h0 + dy, n_img);
          tma_load_2d(b_dst, &p.tmB0, full_bar(s), tap * p.cin + cc * p.kc, n0);
          if (NPL == 2) {
            tma_load_4d(a_dst + p.a_plane_bytes, &p.tmA1, full_bar(s), cc * p.kc, w0 + dx, h0 + dy, n_img);
            tma_load_2d(b_dst + p.b_plane_bytes, &p.tmB1, full_bar(s), tap * p.cin + cc * p.kc, n0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TF32 ? 2u : (KIND == K_BF16 ? 1u : 0u), 0u, 0u, 128u, (uint32_t)p.BN);
      const uint32_t layout = p.row_bytes == 128 ? 2u : 4u;
      const uint32_t sbo = 8u * p.row_bytes;
      const int nks = p.row_bytes / 32;
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount % p.nbuf, bph = (tcount / p.nbuf) & 1;
        mbar_wait(tempty_bar(buf), bph ^ 1);     // epilogue has drained this accumulator buffer
        tc_fence_after_sync();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % S, ph = (it / S) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after_sync();
          const uint32_t a0 = base + s * p.stage_bytes;
          const uint32_t b0 = a0 + NPL * p.a_plane_bytes;
          const uint32_t acc = tmem_base + (uint32_t)((buf * p.nacc + (kb % p.nacc)) * p.BN);
          uint32_t accum = kb >= p.nacc ? 1u : 0u;
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t ko = ks * 32;
            if (NPL == 2) {
              umma<TF32>(acc, make_smem_desc(a0 + p.a_plane_bytes + ko, 16, sbo, layout),
                         make_smem_desc(b0 + ko, 16, sbo, layout), idesc, accum);
              umma<TF32>(acc, make_smem_desc(a0 + ko, 16, sbo, layout),
                         make_smem_desc(b0 + p.b_plane_bytes + ko, 16, sbo, layout), idesc, 1u);
              umma<TF32>(acc, make_smem_desc(a0 + ko, 16, sbo, layout), make_smem_desc(b0 + ko, 16, sbo, layout),
                         idesc, 1u);
            } else {
              umma<TF32>(acc, make_smem_desc(a0 + ko, 16, sbo, layout), make_smem_desc(b0 + ko, 16, sbo, layout),
                         idesc, accum);
            }
            accum = 1u;
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(buf));
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (TMEM lane quadrant = warp % 4)
    const int q = warp & 3;
    const int row = q * 32 + lane;
    float scale = p.out_scale;
    if (p.out_scale_ptr) scale *= __ldg(p.out_scale_ptr);
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tcount) {
      const int n_tile = tile % p.n_tiles, m_tile = tile / p.n_tiles;
      const int tw_i = m_tile % p.tiles_w, th_i = (m_tile / p.tiles_w) % p.tiles_h;
      const int n_img = m_tile / (p.tiles_w * p.tiles_h);
      const int h0 = th_i * p.TH, w0 = tw_i * p.TW, n0 = n_tile * p.BN;
      const uint32_t buf = tcount % p.nbuf, bph = (tcount / p.nbuf) & 1;
      const int hh = h0 + (row >> p.tw_log2), ww = w0 + (row & (p.TW - 1));
      const bool valid = hh < p.H && ww < p.W;
      float* zrow = p.z + (((size_t)n_img * p.H + hh) * p.W + ww) * p.z_ctot + p.z_coff + n0;
      mbar_wait(tfull_bar(buf), bph);
      tc_fence_after_sync();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * p.nacc * p.BN);
      for (int ch = 0; ch < p.BN / 32; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32(tbase + (uint32_t)(ch * 32), r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        for (int a = 1; a < p.nacc; ++a) {
          tmem_ld_32x32(tbase + (uint32_t)(a * p.BN + ch * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(r[j]);
        }
        const float* bp = p.bias ? p.bias + n0 + ch * 32 : nullptr;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float t = v[j] * scale;
          if (bp) t += __ldg(bp + j);
          v[j] = valid ? t : 0.f;
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(zrow + ch * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (p.stat_partial) {
          float sq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
          const float s1 = column_sums_32(v, lane);
          const float s2 = column_sums_32(sq, lane);
          float* out = p.stat_partial + ((size_t)(m_tile * 4 + q) * 2) * p.cout + n0 + ch * 32 + lane;
          out[0] = s1;
          out[p.cout] = s2;
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ================================================================================================ host side
static inline int kind_of(int fmt) { return fmt == AIDE_FMT_BF16 ? K_BF16 : fmt == AIDE_FMT_F16X2 ? K_F16X2 : K_TF32X2; }

int tc_stat_rows(int N, int H, int W) {
  int TW = 128, TH = 1;
  pick_tile(H, W, 128, &TW, &TH);
  return 4 * N * ceil_div(W, TW) * ceil_div(H, TH);     // one row per (pixel tile, TMEM lane quadrant)
}

bool tc_shape_ok(int fmt, int cin, int cout) {
  if (fmt != AIDE_FMT_TF32X2 && fmt != AIDE_FMT_BF16 && fmt != AIDE_FMT_F16X2) return false;
  return cin % 32 == 0 && cout % 32 == 0 && cin >= 32 && cout >= 32;
}

template <int KIND>
static int launch_fwd(const FwdParams& p, int grid, int smem, cudaStream_t st) {
  static thread_local bool done[16] = {false};   // per kernel instantiation and device; not a stream operation
  int dev = 0;
  AIDE_CUDA(cudaGetDevice(&dev));
  if (dev >= 16 || !done[dev]) {
    AIDE_CUDA(cudaFuncSetAttribute(conv3x3_fwd_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (dev < 16) done[dev] = true;
  }
  conv3x3_fwd_tc_kernel<KIND><<<grid, kFwdThreads, smem, st>>>(p);
  AIDE_CHECK_LAUNCH();
  return 0;
}

int tc_conv3x3(int fmt, const void* x0, const void* x1, int x_ctot, int x_coff, int cin, const void* w0, const void* w1,
               const float* bias, float* z, int z_ctot, int z_coff, int cout, int N, int H, int W, float* stat_partial,
               float out_scale, const float* out_scale_ptr, cudaStream_t st) {
  const int kind = kind_of(fmt);
  const int es = fmt_elem_bytes(fmt), npl = fmt_planes(fmt);
  const int dtype = fmt == AIDE_FMT_BF16 ? 1 : fmt == AIDE_FMT_F16X2 ? 2 : 0;
  AIDE_REQUIRE(npl == 1 || (x1 && w1), "conv3x3(tc): two-plane operand formats need hi and lo planes");
  FwdParams p{};
  pick_tile(H, W, 128, &p.TW, &p.TH);
  p.tw_log2 = ilog2(p.TW);
  p.tiles_w = ceil_div(W, p.TW);
  p.tiles_h = ceil_div(H, p.TH);
  const int m_tiles = N * p.tiles_w * p.tiles_h;
  // cout tile: the widest that still gives every SM a tile (wide tiles amortise the A operand: shared-memory and
  // L2 traffic per MMA cycle fall with BN), narrower when the pixel-tile count alone cannot fill the GPU
  p.BN = 32;
  for (int bn = 256; bn >= 32; bn >>= 1) {
    if (cout % bn) continue;
    if ((long long)m_tiles * (cout / bn) >= (kNumSMs * 4) / 5 || bn == 32) {
      p.BN = bn;
      break;
    }
  }
  p.n_tiles = cout / p.BN;
  p.total_tiles = m_tiles * p.n_tiles;
  // K slice: 128-byte rows when that leaves >= 4 pipeline stages, else 64-byte rows (finer-grained pipeline)
  const int smem_budget = 227 * 1024 - 2048;
  int rb = (cin * es >= 128 && cin % (128 / es) == 0) ? 128 : 64;
  if (rb == 128 && smem_budget / (npl * (128 + p.BN) * 128) < 4) rb = 64;
  AIDE_REQUIRE(cin % (rb / es) == 0, "conv3x3(tc): cin=%d not a multiple of the K slice %d", cin, rb / es);
  p.row_bytes = rb;
  p.kc = rb / es;
  p.n_cchunks = cin / p.kc;
  p.a_plane_bytes = 128 * rb;
  p.b_plane_bytes = p.BN * rb;
  p.stage_bytes = npl * (p.a_plane_bytes + p.b_plane_bytes);
  p.stages = smem_budget / p.stage_bytes;
  if (p.stages > 8) p.stages = 8;
  if (p.stages > 9 * p.n_cchunks) p.stages = 9 * p.n_cchunks;
  AIDE_REQUIRE(p.stages >= 2, "conv3x3(tc): stage too large (%d bytes)", p.stage_bytes);
  p.data_bytes = (p.stages * p.stage_bytes + 1023) / 1024 * 1024;
  const int smem = p.data_bytes + 8 * (2 * p.stages + 5) + 1024;
  AIDE_REQUIRE(smem <= 227 * 1024, "conv3x3(tc): shared memory %d too large", smem);
  // accumulators: long split-precision chains are spread over 2 (4) accumulators; double-buffer when TMEM allows
  const long long chain = (long long)9 * p.n_cchunks * (rb / 32) * (npl == 2 ? 3 : 1);
  p.nacc = 1;
  if (npl == 2) {
    if (chain > 1200 && 4 * p.BN <= 512) p.nacc = 4;
    else if (chain > 400 && 2 * p.BN <= 512) p.nacc = 2;
  }
  p.nbuf = (2 * p.nacc * p.BN <= 512) ? 2 : 1;
  int cols = p.nbuf * p.nacc * p.BN;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  p.z = z; p.bias = bias; p.stat_partial = stat_partial;
  p.out_scale = out_scale; p.out_scale_ptr = out_scale_ptr;
  p.z_ctot = z_ctot; p.z_coff = z_coff; p.cout = cout; p.cin = cin; p.H = H; p.W = W;

  if (act_tmap(&p.tmA0, dtype, x0, x_ctot, x_coff, cin, N, H, W, p.kc, p.TW, p.TH, rb)) return 1;
  if (mat_tmap(&p.tmB0, dtype, w0, cout, 9 * cin, p.kc, p.BN, rb)) return 1;
  if (npl == 2) {
    if (act_tmap(&p.tmA1, dtype, x1, x_ctot, x_coff, cin, N, H, W, p.kc, p.TW, p.TH, rb)) return 1;
    if (mat_tmap(&p.tmB1, dtype, w1, cout, 9 * cin, p.kc, p.BN, rb)) return 1;
  }
  const int grid = p.total_tiles < kNumSMs ? p.total_tiles : kNumSMs;
  if (kind == K_BF16) return launch_fwd<K_BF16>(p, grid, smem, st);
  if (kind == K_F16X2) return launch_fwd<K_F16X2>(p, grid, smem, st);
  return launch_fwd<K_TF32X2>(p, grid, smem, st);
}

}  // namespace aide
