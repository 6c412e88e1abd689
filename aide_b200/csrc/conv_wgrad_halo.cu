// conv_wgrad_halo.cu -- conv3x3 backward-filter (wgrad), second-generation tcgen05 kernel for 16-bit operand planes
// (AIDE_FMT_F16X2: 3 kind::f16 MMAs per product; AIDE_FMT_BF16: 1).
//
//   dW[tap=(dy,dx)][ci][co] = sum_{n,h,w} x[n, h+dy-1, w+dx-1, ci] * dz[n, h, w, co]
//
// The first kernel (conv_tc.cu) loads a fresh activation box per tap and a fresh dZ box per 128 (tap,ci) rows: at
// ~96 FLOP per TMA byte it is bound by L2->SM ingest (~32 B/clk/SM), not by the tensor pipe.  Here one CTA owns
// (a block of 128 input channels) x (one filter row dy) x (a tile of BN output channels) and keeps THREE accumulators
// in TMEM, one per dx: a pixel tile is 16 (w) x 4 (h); its activation box is loaded once WITH a one-pixel halo in w
// (18 x 4 pixels) and the three dx taps are UMMA descriptors into that box (start address moved by whole pixel rows --
// the 128B swizzle is a function of the shared-memory address bits, see conv_halo_tc.cu), against one dZ box.  Both
// operands are MN-major (channels contiguous, pixels = K): each MMA consumes one image row segment of 16 pixels.
// Bytes per MMA fall ~3x (277 FLOP per TMA byte at BN = 128).  Split-K over pixel tiles into a workspace, fixed-order
// reduction by wgrad_reduce_kernel (deterministic).
//
// STACKM (F16X2, cin % 128 == 64 -- the 64-channel layers, which used to fall back to the first kernel at 9 % tensor
// pipe): the 128 MMA rows are [hi plane, 64 channels | lo plane, 64 channels] -- the two 64-channel blocks of the
// MN-major operand are simply the two planes (LBO = plane distance) -- so x_hi*dz and x_lo*dz come out of ONE MMA in
// TMEM lanes [0,64) and [64,128); two MMAs per tap (against dz_hi and dz_lo) give the full 4-term product and the
// epilogue adds the two lane halves through shared memory.  Layers with 32 channels on either side (cin or cout % 64 ==
// 32: the first encoder level) run the same kernel on 64-channel boxes whose upper half lies outside the tensor -- TMA
// zero-fills it, the MMAs multiply zeros, the epilogue stores the valid rows / columns only (31 % of the time of the
// first-generation kernel these layers used to fall back to).
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer (warp-uniform loop, one elected lane), warps 2..5
// epilogue (TMEM lane = input channel).
#include "common.cuh"
#include "sm100_ptx.cuh"

#include <cstdlib>

namespace aide {

using namespace ptx;

int act_tmap(CUtensorMap* m, int dtype, const void* plane, int ctot, int coff, int C, int N, int H, int W, int box_c,
             int box_w, int box_h, int swizzle_bytes);
int launch_wgrad_reduce(const float* ws, int splits, int cout, int cin, int layout, float* dw, float scale,
                        const float* scale_ptr, cudaStream_t st);

namespace {

constexpr int kThreads = 192;
constexpr int kTW = 16, kTH = 4;                 // pixel tile: one MMA (K = 16) per image row segment
constexpr int kHW = kTW + 2;                     // activation box width (halo of one pixel left and right)
constexpr int kABlk = kHW * kTH * 128;           // bytes of one 64-channel activation block (72 rows x 128 B)
constexpr int kBBlk = kTW * kTH * 128;           // bytes of one 64-channel dZ block (64 rows x 128 B)
constexpr int kStagesMax = 4;
constexpr int kSmemMax = 227 * 1024;

struct WParams {
  CUtensorMap tmX0, tmX1, tmD0, tmD1;
  float* ws;
  int cin, cout, H, W;
  int tiles_w, tiles_h, tiles_total, tiles_per_split;
  int BN, nb_b;                                  // cout tile, number of 64-channel dZ blocks in it
  int a_plane, b_plane, stage_bytes, stages, bar_off, tmem_cols;
  int scratch_off;                               // STACKM: 8 KB exchange buffer of the epilogue
};

template <int NPL, bool F16, bool STACKM>
__global__ void __launch_bounds__(kThreads, 1) wgrad_halo_kernel(const __grid_constant__ WParams p) {
  static_assert(!STACKM || NPL == 2, "stacking the planes along M needs a two-plane format");
  constexpr int CB = STACKM ? 64 : 128;          // input channels per CTA
  constexpr int NA = STACKM ? 1 : 2;             // 64-channel activation blocks per plane
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int S = p.stages;
  const uint32_t bar_base = base + p.bar_off;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStagesMax + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kStagesMax);
  const uint32_t slot_addr = bar_base + 8u * (2 * kStagesMax + 1);
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + p.bar_off + 8 * (2 * kStagesMax + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ci_blk = blockIdx.x / 3, dy = blockIdx.x - ci_blk * 3;
  const int n0 = blockIdx.y * p.BN;
  const int t_begin = blockIdx.z * p.tiles_per_split;
  const int t_end = min(p.tiles_total, t_begin + p.tiles_per_split);
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX0);
    tma_prefetch_desc(&p.tmD0);
    if (NPL == 2) {
      tma_prefetch_desc(&p.tmX1);
      tma_prefetch_desc(&p.tmD1);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
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
    const bool leader = elect_one();
    const uint32_t tx = NPL * (NA * kABlk + p.nb_b * kBBlk);
    int s = 0;
    uint32_t ph = 0;
    int n_img = t_begin / tiles_per_img;
    int r = t_begin - n_img * tiles_per_img;
    int th_i = r / p.tiles_w, tw_i = r - th_i * p.tiles_w;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(empty_bar(s), ph ^ 1);
      if (leader) {
        mbar_arrive_expect_tx(full_bar(s), tx);
        const int h0 = th_i * kTH, w0 = tw_i * kTW;
        const uint32_t a_dst = base + s * p.stage_bytes;
        const uint32_t b_dst = a_dst + NPL * p.a_plane;
#pragma unroll
        for (int j = 0; j < NA; ++j) {
          tma_load_4d(a_dst + j * kABlk, &p.tmX0, full_bar(s), ci_blk * CB + j * 64, w0 - 1, h0 + dy - 1, n_img);
          if (NPL == 2)
            tma_load_4d(a_dst + p.a_plane + j * kABlk, &p.tmX1, full_bar(s), ci_blk * CB + j * 64, w0 - 1, h0 + dy - 1,
                        n_img);
        }
        for (int j = 0; j < p.nb_b; ++j) {
          tma_load_4d(b_dst + j * kBBlk, &p.tmD0, full_bar(s), n0 + j * 64, w0, h0, n_img);
          if (NPL == 2) tma_load_4d(b_dst + p.b_plane + j * kBBlk, &p.tmD1, full_bar(s), n0 + j * 64, w0, h0, n_img);
        }
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
      if (++tw_i == p.tiles_w) {
        tw_i = 0;
        if (++th_i == p.tiles_h) { th_i = 0; ++n_img; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // MN-major operands, SWIZZLE_128B: LBO = distance between 64-channel blocks, SBO = 8 pixel rows (1024 B)
    const uint32_t idesc = make_idesc(F16 ? 0u : 1u, 1u, 1u, 128u, (uint32_t)p.BN);
    const uint64_t a_desc0 = make_smem_desc(base, kABlk, 1024, 2u);
    const uint64_t b_desc0 = make_smem_desc(base + NPL * p.a_plane, kBBlk, 1024, 2u);
    const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4;
    const uint32_t a_plane16 = (uint32_t)p.a_plane >> 4, b_plane16 = (uint32_t)p.b_plane >> 4;
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0, accum = 0;
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after_sync();
      if (leader) {
        uint64_t a_row = a_desc0 + (uint64_t)(s * stage16);          // pixel row y of the halo box, dx = 0
        uint64_t b_row = b_desc0 + (uint64_t)(s * stage16);
#pragma unroll
        for (int y = 0; y < kTH; ++y, a_row += (uint64_t)(kHW * 8), b_row += (uint64_t)(kTW * 8)) {   // 128 B = 8 units
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const uint32_t acc = tmem_base + (uint32_t)(dx * p.BN);
            const uint64_t ah = a_row + (uint64_t)(dx * 8);
            if (STACKM) {          // rows [x_hi | x_lo] (LBO = plane distance = kABlk) against dz_hi, then dz_lo
              umma<false>(acc, ah, b_row, idesc, accum);
              umma<false>(acc, ah, b_row + (uint64_t)b_plane16, idesc, 1u);
            } else if (NPL == 2) {
              umma<false>(acc, ah + (uint64_t)a_plane16, b_row, idesc, accum);
              umma<false>(acc, ah, b_row + (uint64_t)b_plane16, idesc, 1u);
              umma<false>(acc, ah, b_row, idesc, 1u);
            } else {
              umma<false>(acc, ah, b_row, idesc, accum);
            }
          }
          accum = 1u;
        }
        umma_commit(empty_bar(s));
      }
      __syncwarp();
      accum = 1u;
      if (++s == S) { s = 0; ph ^= 1; }
    }
    if (leader) umma_commit(done_bar);
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: TMEM lane = input channel
    const int q = warp & 3;
    mbar_wait(done_bar, 0);
    tc_fence_after_sync();
    const bool any = t_end > t_begin;
    if constexpr (STACKM) {
      // lanes [0,64) hold x_hi * dz, lanes [64,128) x_lo * dz of the same 64 channels: warps with q >= 2 hand their
      // values over through shared memory, warps with q < 2 add and store
      float* scratch = reinterpret_cast<float*>(sm + p.scratch_off);            // [64 channels][32 columns]
      const int cl = (q & 1) * 32 + lane;
      const int ci = ci_blk * 64 + cl;
      const int nch = min(p.BN, p.cout - n0) / 32;                              // 32-column chunks inside the tensor
      for (int dx = 0; dx < 3; ++dx) {
        const int tap = dy * 3 + dx;
        float* out = p.ws + (size_t)blockIdx.z * 9 * p.cin * p.cout + ((size_t)tap * p.cin + ci) * p.cout + n0;
        for (int ch = 0; ch < nch; ++ch) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(dx * p.BN + ch * 32), r);
          tmem_ld_wait();
          if (q >= 2) {
#pragma unroll
            for (int k = 0; k < 32; k += 4)
              *reinterpret_cast<float4*>(scratch + cl * 32 + k) =
                  make_float4(__uint_as_float(r[k]), __uint_as_float(r[k + 1]), __uint_as_float(r[k + 2]), __uint_as_float(r[k + 3]));
          }
          named_bar_sync(1, 128);
          if (q < 2 && ci < p.cin) {
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              const float4 lo = *reinterpret_cast<const float4*>(scratch + cl * 32 + k);
              *reinterpret_cast<float4*>(out + ch * 32 + k) =
                  any ? make_float4(__uint_as_float(r[k]) + lo.x, __uint_as_float(r[k + 1]) + lo.y,
                                    __uint_as_float(r[k + 2]) + lo.z, __uint_as_float(r[k + 3]) + lo.w)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          named_bar_sync(2, 128);                 // the exchange buffer is free again
        }
      }
    } else {
      const int ci = ci_blk * 128 + q * 32 + lane;
      const int nch = min(p.BN, p.cout - n0) / 32;
      for (int dx = 0; dx < 3; ++dx) {
        const int tap = dy * 3 + dx;
        float* out = p.ws + (size_t)blockIdx.z * 9 * p.cin * p.cout + ((size_t)tap * p.cin + ci) * p.cout + n0;
        for (int ch = 0; ch < nch; ++ch) {
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(dx * p.BN + ch * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 4)
            *reinterpret_cast<float4*>(out + ch * 32 + k) =
                any ? make_float4(__uint_as_float(r[k]), __uint_as_float(r[k + 1]), __uint_as_float(r[k + 2]),
                                  __uint_as_float(r[k + 3]))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

struct WPlan {
  WParams p;
  int splits, smem;
  dim3 grid;
};

int make_wplan(int fmt, int cin, int cout, int N, int H, int W, WPlan* o) {
  const int npl = fmt_planes(fmt);
  WParams& p = o->p;
  p = WParams{};
  p.cin = cin; p.cout = cout; p.H = H; p.W = W;
  p.BN = cout % 128 == 0 ? 128 : 64;
  p.nb_b = p.BN / 64;
  p.tiles_w = ceil_div(W, kTW);
  p.tiles_h = ceil_div(H, kTH);
  p.tiles_total = N * p.tiles_w * p.tiles_h;
  const bool stackm = cin % 128 != 0;            // 64-channel blocks: hi / lo planes stacked along the MMA's M
  const int cb = stackm ? 64 : 128;
  p.a_plane = (stackm ? 1 : 2) * kABlk;
  p.b_plane = p.nb_b * kBBlk;
  p.stage_bytes = npl * (p.a_plane + p.b_plane);
  p.stages = (kSmemMax - 2048 - (stackm ? 8192 + 128 : 0)) / p.stage_bytes;
  if (p.stages > kStagesMax) p.stages = kStagesMax;
  if (p.stages < 2) return 1;
  p.bar_off = p.stages * p.stage_bytes;
  p.scratch_off = p.bar_off + 128;
  o->smem = 1024 + p.bar_off + 128 + (stackm ? 8192 : 0);
  p.tmem_cols = 3 * p.BN <= 256 ? 256 : 512;
  // split-K over pixel tiles: minimise (waves over 148 SMs) x (tiles per CTA); ties -> fewer splits
  const long long base_ctas = (long long)ceil_div(cin, cb) * 3 * ceil_div(cout, p.BN);
  long long best_cost = -1;
  int best_s = 1;
  const int max_s = p.tiles_total / 8 > 0 ? p.tiles_total / 8 : 1;
  for (int s = 1; s <= max_s && s <= 4096; ++s) {
    const long long per = ceil_div(p.tiles_total, s);
    const long long waves = (base_ctas * s + kNumSMs - 1) / kNumSMs;
    const long long cost = waves * (per + 6);            // + pipeline fill / epilogue per CTA
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_s = s;
    }
    if (base_ctas * s > 8LL * kNumSMs) break;
  }
  p.tiles_per_split = ceil_div(p.tiles_total, best_s);
  o->splits = ceil_div(p.tiles_total, p.tiles_per_split);
  o->grid = dim3(ceil_div(cin, cb) * 3, ceil_div(cout, p.BN), o->splits);
  return 0;
}

template <int NPL, bool F16, bool STACKM>
int launch(const WPlan& pl, cudaStream_t st) {
  static thread_local bool done[16] = {false};
  int dev = 0;
  AIDE_CUDA(cudaGetDevice(&dev));
  if (dev >= 16 || !done[dev]) {
    AIDE_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel<NPL, F16, STACKM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    if (dev < 16) done[dev] = true;
  }
  wgrad_halo_kernel<NPL, F16, STACKM><<<pl.grid, kThreads, pl.smem, st>>>(pl.p);
  AIDE_CHECK_LAUNCH();
  return 0;
}

int env_flag(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s && *s ? std::atoi(s) : dflt;
}

}  // namespace

bool wgrad_halo_ok(int fmt, int cin, int cout, int N, int H, int W) {
  if (fmt != AIDE_FMT_F16X2 && fmt != AIDE_FMT_BF16) return false;
  if (env_flag("AIDE_WGRAD_HALO", 1) == 0) return false;
  if (cin % 32 || cout % 32 || W < 8 || H < 2) return false;
  if ((cin % 64 || cout % 64) && env_flag("AIDE_WGRAD_PAD32", 1) == 0) return false;
  if (cin % 128 && (fmt != AIDE_FMT_F16X2 || env_flag("AIDE_WGRAD_STACKM", 1) == 0)) return false;   // 64-channel blocks: two-plane format only
  WPlan pl;
  return make_wplan(fmt, cin, cout, N, H, W, &pl) == 0;
}

size_t wgrad_halo_workspace_bytes(int fmt, int cin, int cout, int N, int H, int W) {
  WPlan pl;
  if (make_wplan(fmt, cin, cout, N, H, W, &pl)) return 0;
  return (size_t)pl.splits * 9 * cin * cout * sizeof(float);
}

int wgrad_halo(int fmt, const void* x0, const void* x1, int x_ctot, int x_coff, int cin, const void* dz0, const void* dz1,
               int cout, int N, int H, int W, void* ws, size_t ws_bytes, float* dw, float out_scale,
               const float* out_scale_ptr, cudaStream_t st) {
  WPlan pl;
  AIDE_REQUIRE(make_wplan(fmt, cin, cout, N, H, W, &pl) == 0, "conv3x3_wgrad(halo): no plan");
  WParams& p = pl.p;
  AIDE_REQUIRE(ws && ws_bytes >= (size_t)pl.splits * 9 * cin * cout * sizeof(float), "conv3x3_wgrad(halo): workspace too small");
  p.ws = reinterpret_cast<float*>(ws);
  const int npl = fmt_planes(fmt);
  const int dtype = fmt == AIDE_FMT_BF16 ? 1 : 2;
  AIDE_REQUIRE(npl == 1 || (x1 && dz1), "conv3x3_wgrad(halo): two-plane operand formats need hi and lo planes");
  if (act_tmap(&p.tmX0, dtype, x0, x_ctot, x_coff, cin, N, H, W, 64, kHW, kTH, 128)) return 1;
  if (act_tmap(&p.tmD0, dtype, dz0, cout, 0, cout, N, H, W, 64, kTW, kTH, 128)) return 1;
  if (npl == 2) {
    if (act_tmap(&p.tmX1, dtype, x1, x_ctot, x_coff, cin, N, H, W, 64, kHW, kTH, 128)) return 1;
    if (act_tmap(&p.tmD1, dtype, dz1, cout, 0, cout, N, H, W, 64, kTW, kTH, 128)) return 1;
  }
  int rc = fmt == AIDE_FMT_BF16 ? launch<1, false, false>(pl, st)
           : cin % 128         ? launch<2, true, true>(pl, st)
                               : launch<2, true, false>(pl, st);
  if (rc) return rc;
  return launch_wgrad_reduce(p.ws, pl.splits, cout, cin, 1, dw, out_scale, out_scale_ptr, st);
}

}  // namespace aide
