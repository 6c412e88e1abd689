// conv_halo_tc.cu -- conv3x3 (stride 1, pad 1) forward / dgrad, second-generation tcgen05 kernel.
//
// Why it exists (profiles/r1a_*): the first kernel (conv_fwd_tc.cu) fetches a fresh 128-pixel activation box for each
// of the 9 filter taps and a fresh weight slice per 128-pixel tile.  ncu shows it pinned at ~32 B/clk/SM of TMA
// ingest (l1tex__m_xbar2l1tex_read_bytes 8.1-9.0 TB/s) with the tensor pipe 25-50 % busy: it is L2->SM bandwidth
// bound, not tensor bound.  This kernel cuts the bytes per MMA:
//
//   * HALO REUSE.  An output tile is 8 (w) x 16 (h) pixels.  Its (8+2) x (16+2) input halo is loaded ONCE per channel
//     slice as one dense 4-D TMA box (channels x 10 x 18 x 1, out-of-bounds zero fill = the convolution's padding) and
//     all nine taps read it in place: tap (dy,dx) is the same shared-memory tile addressed from pixel
//     (dy+1)*10 + (dx+1), i.e. the UMMA descriptor's start address moves by whole rows and the eight-row core
//     matrices (one output row of 8 pixels each) sit 10 pixel rows apart (SBO = 10 * row bytes).  The 128B/64B
//     swizzle is a function of the shared-memory address bits, so TMA's write pattern and the MMA's read pattern agree
//     for any row offset.  Activation traffic drops from 9 x 128 to 180 pixel rows per slice (6.4x).
//   * PIXEL-TILE BLOCKING.  One CTA iteration computes MB (1, 2 or 4) pixel tiles against the same weight stage, so
//     each weight slice fetched from L2 feeds MB x 128 output pixels.
//
//   D_m[128 pixels, BN cout] += A_m,tap[128 pixels, kc] * W_tap[BN, kc]^T     m < MB, tap < 9, channel slices of kc
//
// Warp roles (192 threads): warp 0 = TMA producer (two rings: activation halos per channel slice, weights per
// (slice, tap)); warp 1 = TMEM owner + MMA issuer; warps 2..5 = epilogue (tcgen05.ld -> scale, +bias -> 128-bit
// stores, BatchNorm partial statistics by shuffle butterfly).  Accumulators are double-buffered in TMEM when
// MB * nacc * BN * 2 <= 512 columns, so the epilogue of one item overlaps the MMAs of the next.
// Operand kinds as in conv_fwd_tc.cu: BF16 (1 MMA), TF32X2 / F16X2 (3 MMAs lo*hi + hi*lo + hi*hi).
#include "common.cuh"
#include "sm100_ptx.cuh"

#include <cstdlib>

namespace aide {

using namespace ptx;

struct HaloFused {                // fused inference epilogue: eval-mode BN scale/shift + ReLU (+ 2x2 max-pool) destinations
  const float* ss;                // [2][cout]
  View dst, pa, pb;
};

int act_tmap(CUtensorMap* m, int dtype, const void* plane, int ctot, int coff, int C, int N, int H, int W, int box_c,
             int box_w, int box_h, int swizzle_bytes);
int mat_tmap(CUtensorMap* m, int dtype, const void* plane, int rows, int kdim, int box_k, int box_rows,
             int swizzle_bytes);

namespace {

enum { K_TF32X2 = 0, K_BF16 = 1, K_F16X2 = 2 };
constexpr int kThreads = 192;
constexpr int kTW = 8, kTH = 16;                 // output tile (pixels); 8 = one UMMA core-matrix group per output row
constexpr int kHW = kTW + 2, kHH = kTH + 2;      // halo box
constexpr int kHaloPix = kHW * kHH;              // 180 pixel rows per halo
constexpr int kMaxAStages = 3, kMaxBStages = 8;

struct HaloParams {
  CUtensorMap tmA0, tmA1, tmB0, tmB1;
  float* z;
  const float* bias;
  float* stat_partial;
  // fused inference epilogue (eval-mode BatchNorm folded in): y = relu(scale[c] * (acc + bias) + shift[c]) written as
  // operand planes into the consumer's channel slice (+ its 2x2 max-pool into up to two half-resolution slices)
  const float* ep_ss;                            // [2][cout] scale, shift (nullptr: plain fp32 z output)
  void *y0, *y1, *pa0, *pa1, *pb0, *pb1;
  int y_ctot, y_coff, pa_ctot, pa_coff, pb_ctot, pb_coff;
  const float* out_scale_ptr;
  float out_scale;
  int z_ctot, z_coff, cout, cin, H, W;
  int tiles_w, tiles_h, m_tiles, n_tiles, total_items;
  int BN, kc, n_cchunks, row_bytes;
  int MB, nacc, nbuf, tmem_cols;
  int stack, acc_w;                              // hi/lo weight planes stacked along N (see the MMA issuer); columns per accumulator
  int b_resident;                                // all 9 * n_cchunks weight stages stay in shared memory (loaded once per CTA)
  int lo_col;                                    // stacked mode: column offset of the A_lo x W_hi product (BN = with hi*lo, 0 = with hi*hi)
  int a_slot_bytes, a_stage_bytes, a_stages;
  int b_plane_bytes, b_stage_bytes, b_stages;
  int b_off, bar_off;
};

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

// 32 consecutive channels of one pixel -> operand planes (the conversions of common.cuh::st4, 128-bit stores)
template <int KIND>
__device__ __forceinline__ void store_planes_32(void* p0, void* p1, size_t e, const float (&v)[32]) {
  if constexpr (KIND == K_F16X2) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      __align__(16) __half h[8];
      __align__(16) __half l[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) f16_split(v[j + k] * kF16ActScale, h[k], l[k]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p0) + e + j) = *reinterpret_cast<const uint4*>(h);
      *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p1) + e + j) = *reinterpret_cast<const uint4*>(l);
    }
  } else if constexpr (KIND == K_BF16) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      __align__(16) __nv_bfloat16 h[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) h[k] = __float2bfloat16_rn(v[j + k]);
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p0) + e + j) = *reinterpret_cast<const uint4*>(h);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 hi = make_float4(tf32_rn(v[j]), tf32_rn(v[j + 1]), tf32_rn(v[j + 2]), tf32_rn(v[j + 3]));
      float4 lo = make_float4(v[j] - hi.x, v[j + 1] - hi.y, v[j + 2] - hi.z, v[j + 3] - hi.w);
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p0) + e + j) = hi;
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p1) + e + j) = lo;
    }
  }
}

// Epilogue of one 8x16-pixel tile for one warp (TMEM lane quadrant q): add up the accumulator column blocks, scale + bias,
// then either fp32 z + BatchNorm partial sums (training) or BN(eval) + ReLU (+ 2x2 max-pool) into operand planes.
template <int KIND>
__device__ __forceinline__ void epilogue_tile(const HaloParams& p, uint32_t tmem_base, uint32_t buf, int m, int mt, int n0,
                                              int q, int lane, float scale, int tiles_per_img) {
  const int row = q * 32 + lane;
  const int ty = row >> 3, tx = row & 7;
  const int n_img = mt / tiles_per_img, r = mt - n_img * tiles_per_img;
  const int hh = (r / p.tiles_w) * kTH + ty, ww = (r % p.tiles_w) * kTW + tx;
  const bool valid = hh < p.H && ww < p.W;
  float* zrow = p.z + (((size_t)n_img * p.H + hh) * p.W + ww) * p.z_ctot + p.z_coff + n0;
  const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((buf * p.MB + m) * p.nacc * p.acc_w);
  const int nparts = p.nacc * (p.stack ? 2 : 1);           // BN-wide column blocks to add up per accumulator set
  for (int ch = 0; ch < p.BN / 32; ++ch) {
    uint32_t rr[32];
    tmem_ld_32x32(tbase + (uint32_t)(ch * 32), rr);
    tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rr[j]);
    for (int a = 1; a < nparts; ++a) {
      tmem_ld_32x32(tbase + (uint32_t)(a * p.BN + ch * 32), rr);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(rr[j]);
    }
    const float* bp = p.bias ? p.bias + n0 + ch * 32 : nullptr;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float t = v[j] * scale;
      if (bp) t += __ldg(bp + j);
      v[j] = valid ? t : 0.f;
    }
    if (p.ep_ss) {
      // ---- inference: BatchNorm (running statistics) + ReLU (+ max-pool) here, same operations in the same
      // order as bn_relu_apply_kernel applies to the fp32 z of the unfused path -> bit-identical planes
      const float* sc = p.ep_ss + n0 + ch * 32;
      const float* sh = sc + p.cout;
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(fmaf(v[j], __ldg(sc + j), __ldg(sh + j)), 0.f);
      const size_t pix = ((size_t)n_img * p.H + hh) * p.W + ww;
      if (valid && p.y0) store_planes_32<KIND>(p.y0, p.y1, pix * p.y_ctot + p.y_coff + n0 + ch * 32, v);
      if (p.pa0 || p.pb0) {
        // 2x2 max-pool: the window partners of pixel (ty, tx) are lanes ^1 (tx) and ^8 (ty) of this warp
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float m = fmaxf(v[j], __shfl_xor_sync(0xffffffffu, v[j], 1));
          v[j] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
        }
        if (valid && !(tx & 1) && !(ty & 1)) {
          const size_t win = ((size_t)n_img * (p.H >> 1) + (hh >> 1)) * (p.W >> 1) + (ww >> 1);
          if (p.pa0) store_planes_32<KIND>(p.pa0, p.pa1, win * p.pa_ctot + p.pa_coff + n0 + ch * 32, v);
          if (p.pb0) store_planes_32<KIND>(p.pb0, p.pb1, win * p.pb_ctot + p.pb_coff + n0 + ch * 32, v);
        }
      }
      continue;
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
      float* out = p.stat_partial + ((size_t)(mt * 4 + q) * 2) * p.cout + n0 + ch * 32 + lane;
      out[0] = s1;
      out[p.cout] = s2;
    }
  }
}

template <int KIND, int NKS>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_halo_tc_kernel(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  constexpr bool TF32 = KIND == K_TF32X2;
  constexpr int NPL = KIND == K_BF16 ? 1 : 2;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int SA = p.a_stages, SB = p.b_stages;
  const uint32_t bar_base = base + p.bar_off;
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (kMaxAStages + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (2 * kMaxAStages + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (2 * kMaxAStages + kMaxBStages + s); };
  auto tfull = [&](int b) { return bar_base + 8u * (2 * kMaxAStages + 2 * kMaxBStages + b); };
  auto tempty = [&](int b) { return bar_base + 8u * (2 * kMaxAStages + 2 * kMaxBStages + 2 + b); };
  constexpr int kSlotIdx = 2 * kMaxAStages + 2 * kMaxBStages + 4;
  const uint32_t slot_addr = bar_base + 8u * kSlotIdx;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + p.bar_off + 8 * kSlotIdx);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmB0);
    if (NPL == 2) {
      tma_prefetch_desc(&p.tmA1);
      tma_prefetch_desc(&p.tmB1);
    }
    for (int s = 0; s < SA; ++s) {
      mbar_init(afull(s), 1);
      mbar_init(aempty(s), 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(bfull(s), 1);
      mbar_init(bempty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), 4);
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
    // The whole warp walks the loops (warp-uniform addresses and ring counters, no divisions in the inner loops);
    // one elected lane arms the barriers and issues the bulk-tensor loads.
    const bool leader = elect_one();
    const uint32_t halo_bytes = kHaloPix * p.row_bytes;
    const uint32_t b_tx = NPL * p.b_plane_bytes;
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    if (p.b_resident) {
      // RESIDENT WEIGHTS (small layers: the whole [cout][9*cin] matrix fits next to the halo ring): every
      // (channel slice, tap) stage is fetched ONCE per CTA instead of once per item -- the per-item weight re-fetch
      // was what pinned the Cout <= 64 layers at the L2 throughput cap (profiles/r1f_conv_halo_f16x2_b32_full.csv).
      if (leader) {
        mbar_arrive_expect_tx(bfull(0), (uint32_t)(9 * p.n_cchunks) * b_tx);
        int s = 0;
        for (int c = 0, c0 = 0; c < p.n_cchunks; ++c, c0 += p.kc) {
          for (int tap = 0; tap < 9; ++tap, ++s) {
            const uint32_t b_dst = base + p.b_off + s * p.b_stage_bytes;
            tma_load_2d(b_dst, &p.tmB0, bfull(0), tap * p.cin + c0, 0);
            if (NPL == 2) tma_load_2d(b_dst + p.b_plane_bytes, &p.tmB1, bfull(0), tap * p.cin + c0, 0);
          }
        }
      }
      __syncwarp();
    }
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
      const int n_tile = item % p.n_tiles, m0 = (item / p.n_tiles) * p.MB;
      const int mbv = min(p.MB, p.m_tiles - m0);
      const int n0 = n_tile * p.BN;
      int th0[4], tw0[4], tn[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int mt = m0 + (m < mbv ? m : 0);
        tn[m] = mt / tiles_per_img;
        const int r = mt - tn[m] * tiles_per_img;
        th0[m] = (r / p.tiles_w) * kTH - 1;
        tw0[m] = (r % p.tiles_w) * kTW - 1;
      }
      int c0 = 0;
      for (int c = 0; c < p.n_cchunks; ++c, c0 += p.kc) {
        mbar_wait(aempty(sa), pha ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(afull(sa), NPL * mbv * halo_bytes);
          const uint32_t a_dst = base + sa * p.a_stage_bytes;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            if (m < mbv) {
              tma_load_4d(a_dst + m * p.a_slot_bytes, &p.tmA0, afull(sa), c0, tw0[m], th0[m], tn[m]);
              if (NPL == 2)
                tma_load_4d(a_dst + (p.MB + m) * p.a_slot_bytes, &p.tmA1, afull(sa), c0, tw0[m], th0[m], tn[m]);
            }
          }
        }
        __syncwarp();
        if (++sa == SA) { sa = 0; pha ^= 1; }
        if (p.b_resident) continue;
        int kcol = c0;                                  // column of (tap, channel slice) in the [cout][9*cin] matrix
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap, kcol += p.cin) {
          mbar_wait(bempty(sb), phb ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(bfull(sb), b_tx);
            const uint32_t b_dst = base + p.b_off + sb * p.b_stage_bytes;
            tma_load_2d(b_dst, &p.tmB0, bfull(sb), kcol, n0);
            if (NPL == 2) tma_load_2d(b_dst + p.b_plane_bytes, &p.tmB1, bfull(sb), kcol, n0);
          }
          __syncwarp();
          if (++sb == SB) { sb = 0; phb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The WHOLE warp walks the loop so that every address below is warp-uniform (uniform registers); one elected
    // lane issues tcgen05.mma / tcgen05.commit.  Descriptors are advanced with 64-bit adds on the start-address
    // field (16-byte units): +2 per 32-byte K step, +row offset per tap, +slot per pixel tile.
    const uint32_t idesc = make_idesc(TF32 ? 2u : (KIND == K_BF16 ? 1u : 0u), 0u, 0u, 128u, (uint32_t)p.BN);
    // Stacked split-precision step: the weight stage holds the hi plane followed by the lo plane, i.e. ONE K-major
    // operand of 2*BN rows, so  A_hi x [W_hi; W_lo]  is a single MMA of width 2*BN writing hi*hi to columns [0, BN) and
    // hi*lo to [BN, 2*BN); A_lo x W_hi follows into [BN, 2*BN) as well (lo_col) and the epilogue adds the two halves.
    // Two MMAs instead of three, the first of them wide -- narrow cout tiles are bound by operand fetch per MMA, not by
    // tensor work.  Keeping BOTH small cross terms out of the hi*hi columns halves the number of truncating
    // accumulations into the large partial sum (tcgen05 truncates when it adds into TMEM): the cross terms are 2^-11
    // smaller, so is the truncation error of their own chain.
    const uint32_t idesc2 = make_idesc(TF32 ? 2u : (KIND == K_BF16 ? 1u : 0u), 0u, 0u, 128u, (uint32_t)(2 * p.BN));
    const bool stack = NPL == 2 && p.stack != 0;
    const uint32_t layout = NKS == 4 ? 2u : 4u;
    const uint64_t a_desc0 = make_smem_desc(base, 16, kHW * p.row_bytes, layout);
    const uint64_t b_desc0 = make_smem_desc(base + p.b_off, 16, 8 * p.row_bytes, layout);
    const uint32_t a_plane16 = (uint32_t)(p.MB * p.a_slot_bytes) >> 4;    // hi plane -> lo plane, in 16-byte units
    const uint32_t b_plane16 = (uint32_t)p.b_plane_bytes >> 4;
    const uint32_t a_slot16 = (uint32_t)p.a_slot_bytes >> 4;
    const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4, b_stage16 = (uint32_t)p.b_stage_bytes >> 4;
    const uint32_t row16 = (uint32_t)p.row_bytes >> 4;
    const uint32_t acc_tile = (uint32_t)(p.nacc * p.acc_w);
    const bool leader = elect_one();
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0, tcount = 0;
    const bool resident = p.b_resident != 0;
    if (resident && (int)blockIdx.x < p.total_items) {
      mbar_wait(bfull(0), 0);
      tc_fence_after_sync();
    }
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++tcount) {
      const int m0 = (item / p.n_tiles) * p.MB;
      const int mbv = min(p.MB, p.m_tiles - m0);
      const uint32_t buf = tcount % p.nbuf, bph = (tcount / p.nbuf) & 1;
      mbar_wait(tempty(buf), bph ^ 1);
      tc_fence_after_sync();
      const uint32_t acc_item = tmem_base + buf * p.MB * acc_tile;
      int ai = 0;
      uint32_t first = 0;                      // bit a set once accumulator a has been written in this item
      for (int c = 0; c < p.n_cchunks; ++c) {
        mbar_wait(afull(sa), pha);
        tc_fence_after_sync();
        uint64_t a_row = a_desc0 + (uint64_t)(sa * a_stage16);
        if (resident) sb = c * 9;                       // stage index = (channel slice, tap)
#pragma unroll 1
        for (int dy = 0; dy < 3; ++dy, a_row += (uint64_t)(kHW * row16)) {
          uint64_t a_tap = a_row;
#pragma unroll 1
          for (int dx = 0; dx < 3; ++dx, a_tap += (uint64_t)row16) {
            if (!resident) {
              mbar_wait(bfull(sb), phb);
              tc_fence_after_sync();
            }
            if (leader) {
              const uint64_t b_hi = b_desc0 + (uint64_t)(sb * b_stage16);
              const uint32_t flag0 = (first >> ai) & 1u;
              uint64_t a_hi = a_tap;
              uint32_t acc = acc_item + (uint32_t)ai * p.acc_w;
              for (int m = 0; m < mbv; ++m, a_hi += (uint64_t)a_slot16, acc += acc_tile) {
#pragma unroll
                for (int ks = 0; ks < NKS; ++ks) {
                  const uint32_t flag = ks == 0 ? flag0 : 1u;
                  if (NPL == 2 && stack) {
                    umma<TF32>(acc, a_hi + (uint64_t)(2 * ks), b_hi + (uint64_t)(2 * ks), idesc2, flag);
                    umma<TF32>(acc + (uint32_t)p.lo_col, a_hi + (uint64_t)(a_plane16 + 2 * ks), b_hi + (uint64_t)(2 * ks), idesc, 1u);
                  } else if (NPL == 2) {
                    umma<TF32>(acc, a_hi + (uint64_t)(a_plane16 + 2 * ks), b_hi + (uint64_t)(2 * ks), idesc, flag);
                    umma<TF32>(acc, a_hi + (uint64_t)(2 * ks), b_hi + (uint64_t)(b_plane16 + 2 * ks), idesc, 1u);
                    umma<TF32>(acc, a_hi + (uint64_t)(2 * ks), b_hi + (uint64_t)(2 * ks), idesc, 1u);
                  } else {
                    umma<TF32>(acc, a_hi + (uint64_t)(2 * ks), b_hi + (uint64_t)(2 * ks), idesc, flag);
                  }
                }
              }
              if (!resident) umma_commit(bempty(sb));
            }
            __syncwarp();
            first |= 1u << ai;
            ai = ai + 1 == p.nacc ? 0 : ai + 1;
            if (resident) ++sb;
            else if (++sb == SB) { sb = 0; phb ^= 1; }
          }
        }
        if (leader) umma_commit(aempty(sa));
        __syncwarp();
        if (++sa == SA) { sa = 0; pha ^= 1; }
      }
      if (leader) umma_commit(tfull(buf));
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue (TMEM lane quadrant = warp % 4)
    const int q = warp & 3;
    float scale = p.out_scale;
    if (p.out_scale_ptr) scale *= __ldg(p.out_scale_ptr);
    uint32_t tcount = 0;
    for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++tcount) {
      const int n_tile = item % p.n_tiles, m0 = (item / p.n_tiles) * p.MB;
      const int mbv = min(p.MB, p.m_tiles - m0);
      const int n0 = n_tile * p.BN;
      const uint32_t buf = tcount % p.nbuf, bph = (tcount / p.nbuf) & 1;
      mbar_wait(tfull(buf), bph);
      tc_fence_after_sync();
      for (int m = 0; m < mbv; ++m) {
        epilogue_tile<KIND>(p, tmem_base, buf, m, m0 + m, n0, q, lane, scale, tiles_per_img);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(buf));
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ================================================================================================ CTA-pair kernel
// conv3x3_halo2_tc_kernel: the same algorithm on cta_group::2 -- two CTAs on the two SMs of a TPC (cluster of 2) issue ONE
// MMA of M = 256: each CTA owns MB pixel tiles (its 128 rows of A, its own halo ring, its own TMEM accumulators) and
// holds only HALF of the weight operand's N rows, which the tensor cores of the two SMs exchange.  Per SM the weight
// stage shrinks from 2*BN to 1.5*BN rows (stacked split precision) and the shared-memory operand reads per MMA from
// 128 + N to 128 + N/2 rows -- the bound of the BN <= 128 layers.
//
//   stacked F16X2 stage of a CTA:   Y (BN rows): rank 0 = W_hi[n0, n0+BN), rank 1 = W_lo[n0, n0+BN)
//                                    X (BN/2 rows): rank r = W_hi[n0 + r*BN/2, +BN/2)
//   MMA 1: A_hi x Y  (N = 2*BN: columns [0,BN) = hi*hi, [BN,2BN) = hi*lo)     MMA 2: A_lo x X (N = BN) -> columns lo_col
//   BF16 (one plane): Y only, BN/2 rows per CTA, one MMA of N = BN.
//
// Barriers: "full" barriers live in the LEADER's shared memory and count the TMA bytes of both CTAs; "empty" / "TMEM full"
// barriers exist in both CTAs and are signalled by one multicast tcgen05.commit; "TMEM empty" is the leader's, with the
// epilogue warps of both CTAs arriving on it.  Only the leader's warp 1 issues MMAs.
struct Halo2Params {
  HaloParams h;
  CUtensorMap tmBh;                // hi-plane weights with a box of BN/2 rows (the X region)
  int pair_items, yx_bytes;        // work items of a pair; bytes of region Y (= offset of X inside a weight stage)
};

template <int KIND, int NKS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv3x3_halo2_tc_kernel(const __grid_constant__ Halo2Params pp) {
  static_assert(KIND == K_F16X2 || KIND == K_BF16, "the pair kernel issues kind::f16 MMAs");
  const HaloParams& p = pp.h;
  extern __shared__ uint8_t smem_raw[];
  constexpr int NPL = KIND == K_BF16 ? 1 : 2;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const int SA = p.a_stages, SB = p.b_stages;
  const uint32_t bar_base = base + p.bar_off;
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (kMaxAStages + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (2 * kMaxAStages + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (2 * kMaxAStages + kMaxBStages + s); };
  auto tfull = [&](int b) { return bar_base + 8u * (2 * kMaxAStages + 2 * kMaxBStages + b); };
  auto tempty = [&](int b) { return bar_base + 8u * (2 * kMaxAStages + 2 * kMaxBStages + 2 + b); };
  constexpr int kSlotIdx = 2 * kMaxAStages + 2 * kMaxBStages + 4;
  const uint32_t slot_addr = bar_base + 8u * kSlotIdx;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + p.bar_off + 8 * kSlotIdx);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmB0);
    if (NPL == 2) {
      tma_prefetch_desc(&p.tmA1);
      tma_prefetch_desc(&p.tmB1);
      tma_prefetch_desc(&pp.tmBh);
    }
    for (int s = 0; s < SA; ++s) {
      mbar_init(afull(s), 1);
      mbar_init(aempty(s), 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(bfull(s), 1);
      mbar_init(bempty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull(b), 1);
      mbar_init(tempty(b), 8);                   // 4 epilogue warps of each CTA of the pair
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(slot_addr, (uint32_t)p.tmem_cols);
    tmem_relinquish2();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                            // the peer's barriers are initialised before anyone signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *slot_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    const bool leader = elect_one();
    const uint32_t halo_bytes = kHaloPix * p.row_bytes;
    const uint32_t a_tx = 2u * NPL * p.MB * halo_bytes;                         // both CTAs
    const uint32_t b_tx = 2u * (uint32_t)p.b_stage_bytes;
    const int half = p.BN / 2;
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    for (int item = pair; item < pp.pair_items; item += n_pairs) {
      const int n_tile = item % p.n_tiles, m0 = ((item / p.n_tiles) * 2 + (int)rank) * p.MB;
      const int n0 = n_tile * p.BN;
      int th0[4], tw0[4], tn[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int mt = m0 + (m < p.MB ? m : 0);
        tn[m] = mt / tiles_per_img;              // >= N for tiles past the end: the whole box is out of bounds -> zeros
        const int r = mt - tn[m] * tiles_per_img;
        th0[m] = (r / p.tiles_w) * kTH - 1;
        tw0[m] = (r % p.tiles_w) * kTW - 1;
      }
      int c0 = 0;
      for (int c = 0; c < p.n_cchunks; ++c, c0 += p.kc) {
        mbar_wait(aempty(sa), pha ^ 1);
        if (leader) {
          if (rank == 0) mbar_arrive_expect_tx(afull(sa), a_tx);
          const uint32_t a_dst = base + sa * p.a_stage_bytes;
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            if (m < p.MB) {
              tma2_load_4d(a_dst + m * p.a_slot_bytes, &p.tmA0, afull(sa), c0, tw0[m], th0[m], tn[m]);
              if (NPL == 2)
                tma2_load_4d(a_dst + (p.MB + m) * p.a_slot_bytes, &p.tmA1, afull(sa), c0, tw0[m], th0[m], tn[m]);
            }
          }
        }
        __syncwarp();
        if (++sa == SA) { sa = 0; pha ^= 1; }
        int kcol = c0;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap, kcol += p.cin) {
          mbar_wait(bempty(sb), phb ^ 1);
          if (leader) {
            if (rank == 0) mbar_arrive_expect_tx(bfull(sb), b_tx);
            const uint32_t b_dst = base + p.b_off + sb * p.b_stage_bytes;
            if (NPL == 2) {
              // Y: the full-height plane of this rank (hi on the leader, lo on the peer); X: this rank's half of hi
              tma2_load_2d(b_dst, rank == 0 ? &p.tmB0 : &p.tmB1, bfull(sb), kcol, n0);
              tma2_load_2d(b_dst + pp.yx_bytes, &pp.tmBh, bfull(sb), kcol, n0 + (int)rank * half);
            } else {
              tma2_load_2d(b_dst, &pp.tmBh, bfull(sb), kcol, n0 + (int)rank * half);
            }
          }
          __syncwarp();
          if (++sb == SB) { sb = 0; phb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t fmtc = KIND == K_BF16 ? 1u : 0u;
      const uint32_t idesc1 = make_idesc(fmtc, 0u, 0u, 256u, (uint32_t)(NPL == 2 ? 2 * p.BN : p.BN));
      const uint32_t idesc2 = make_idesc(fmtc, 0u, 0u, 256u, (uint32_t)p.BN);
      const uint32_t layout = NKS == 4 ? 2u : 4u;
      const uint64_t a_desc0 = make_smem_desc(base, 16, kHW * p.row_bytes, layout);
      const uint64_t b_desc0 = make_smem_desc(base + p.b_off, 16, 8 * p.row_bytes, layout);
      const uint32_t a_plane16 = (uint32_t)(p.MB * p.a_slot_bytes) >> 4;
      const uint32_t x16 = (uint32_t)pp.yx_bytes >> 4;
      const uint32_t a_slot16 = (uint32_t)p.a_slot_bytes >> 4;
      const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4, b_stage16 = (uint32_t)p.b_stage_bytes >> 4;
      const uint32_t row16 = (uint32_t)p.row_bytes >> 4;
      const uint32_t acc_tile = (uint32_t)(p.nacc * p.acc_w);
      const bool leader = elect_one();
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0, tcount = 0;
      for (int item = pair; item < pp.pair_items; item += n_pairs, ++tcount) {
        const uint32_t buf = tcount % p.nbuf, bph = (tcount / p.nbuf) & 1;
        mbar_wait(tempty(buf), bph ^ 1);
        tc_fence_after_sync();
        const uint32_t acc_item = tmem_base + buf * p.MB * acc_tile;
        int ai = 0;
        uint32_t first = 0;
        for (int c = 0; c < p.n_cchunks; ++c) {
          mbar_wait(afull(sa), pha);
          tc_fence_after_sync();
          uint64_t a_row = a_desc0 + (uint64_t)(sa * a_stage16);
#pragma unroll 1
          for (int dy = 0; dy < 3; ++dy, a_row += (uint64_t)(kHW * row16)) {
            uint64_t a_tap = a_row;
#pragma unroll 1
            for (int dx = 0; dx < 3; ++dx, a_tap += (uint64_t)row16) {
              mbar_wait(bfull(sb), phb);
              tc_fence_after_sync();
              if (leader) {
                const uint64_t b_y = b_desc0 + (uint64_t)(sb * b_stage16);
                const uint32_t flag0 = (first >> ai) & 1u;
                uint64_t a_hi = a_tap;
                uint32_t acc = acc_item + (uint32_t)ai * p.acc_w;
                for (int m = 0; m < p.MB; ++m, a_hi += (uint64_t)a_slot16, acc += acc_tile) {
#pragma unroll
                  for (int ks = 0; ks < NKS; ++ks) {
                    const uint32_t flag = ks == 0 ? flag0 : 1u;
                    umma_pair_f16(acc, a_hi + (uint64_t)(2 * ks), b_y + (uint64_t)(2 * ks), idesc1, flag);
                    if (NPL == 2)
                      umma_pair_f16(acc + (uint32_t)p.lo_col, a_hi + (uint64_t)(a_plane16 + 2 * ks),
                                    b_y + (uint64_t)(x16 + 2 * ks), idesc2, 1u);
                  }
                }
                umma_commit_pair(bempty(sb));
              }
              __syncwarp();
              first |= 1u << ai;
              ai = ai + 1 == p.nacc ? 0 : ai + 1;
              if (++sb == SB) { sb = 0; phb ^= 1; }
            }
          }
          if (leader) umma_commit_pair(aempty(sa));
          __syncwarp();
          if (++sa == SA) { sa = 0; pha ^= 1; }
        }
        if (leader) umma_commit_pair(tfull(buf));
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, each its own tiles)
    const int q = warp & 3;
    float scale = p.out_scale;
    if (p.out_scale_ptr) scale *= __ldg(p.out_scale_ptr);
    uint32_t tcount = 0;
    for (int item = pair; item < pp.pair_items; item += n_pairs, ++tcount) {
      const int n_tile = item % p.n_tiles, m0 = ((item / p.n_tiles) * 2 + (int)rank) * p.MB;
      const int n0 = n_tile * p.BN;
      const uint32_t buf = tcount % p.nbuf, bph = (tcount / p.nbuf) & 1;
      mbar_wait(tfull(buf), bph);
      tc_fence_after_sync();
      for (int m = 0; m < p.MB; ++m) {
        if (m0 + m >= p.m_tiles) break;                          // padding tile of the last pair
        epilogue_tile<KIND>(p, tmem_base, buf, m, m0 + m, n0, q, lane, scale, tiles_per_img);
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty(buf));
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();                            // nobody leaves while the peer may still read this CTA's memory
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc2(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ================================================================================================ host side
inline int kind_of(int fmt) { return fmt == AIDE_FMT_BF16 ? K_BF16 : fmt == AIDE_FMT_F16X2 ? K_F16X2 : K_TF32X2; }

int env_int(const char* name, int dflt) {
  const char* s = std::getenv(name);
  return s && *s ? std::atoi(s) : dflt;
}

struct HaloPlan {
  int BN, MB, nacc, nbuf, row_bytes, a_stages, b_stages, stack, resident, occ;   // occ: CTAs per SM (1 or 2); 4 = CTA pair
  int yx;                                                                        // pair: bytes of weight region Y
  int a_slot, a_stage, b_plane, b_stage, smem;
  double cost;
};

constexpr int kSmemMax = 227 * 1024;
constexpr int kBarBytes = 8 * (2 * kMaxAStages + 2 * kMaxBStages + 4 + 1);

// Accumulation chains are split over `nacc` TMEM accumulators in the split-precision formats: tcgen05.mma truncates
// when it adds into the accumulator, so one chain's error grows linearly with its length (tools/accum_probe.py).
int wanted_nacc(int fmt, long long chain) {
  if (fmt == AIDE_FMT_BF16) return 1;
  const int cap = env_int("AIDE_CONV_NACC_MAX", 4);
  int n = chain > 1200 ? 4 : chain > 400 ? 2 : 1;
  return n < cap ? n : cap;
}

// Pick (BN, MB, row bytes, stages) from a time model of one CTA iteration: max(MMA clocks, TMA bytes / ingest rate) scaled
// by a ring-depth penalty, plus the epilogue when the accumulators cannot be double-buffered, times the number of waves
// over 148 SMs.  Two parameter sets, both fitted to the per-layer sweeps of tools/halo_probe.py --sweep on B200
// (profiles/r1b_halo_sweep.json at batch 8, profiles/r1d_halo_sweep_b32.json at batch 32):
//   kRow  chooses the shared-memory row width (64 / 128 B) for a given (BN, MB) -- the choice the sweeps were run with;
//   kSel  ranks the (BN, MB) candidates; its pick is within 15 % of the sweep's best on all 68 measured cases.
struct CostModel {
  double sm_a, sm_d, issue, stage_ovh, bw, pb, pa, epi;
};
constexpr CostModel kRow{256.0, 3.0, 20.0, 100.0, 32.0, 0.2, 0.0, 150.0};
constexpr CostModel kSel{192.0, 6.0, 10.0, 50.0, 64.0, 0.2, 0.0, 150.0};

double model_cost(const CostModel& m, const HaloPlan& c, int npl, int n_cchunks, int nks, long long m_tiles, int cout) {
  if (c.resident) {
    // weights live in shared memory: an item ingests its halos only, against the chip-wide L2 throughput cap
    // (~36 B/clk/SM measured); the MMA side is as in the streaming plan
    const long long items = (m_tiles + c.MB - 1) / c.MB;
    const long long waves = (items + kNumSMs - 1) / kNumSMs;
    auto mma_clk = [&](double n) {
      const double fetch = (m.sm_a + n) / m.sm_d;
      return (n / 2.0 > fetch ? n / 2.0 : fetch) + m.issue;
    };
    const double per_step = npl == 1 ? mma_clk(c.BN) : c.stack ? mma_clk(2.0 * c.BN) + mma_clk(c.BN) : 3.0 * mma_clk(c.BN);
    const double mma = (double)c.MB * 9 * n_cchunks * nks * per_step;
    const double bytes = (double)n_cchunks * npl * c.row_bytes * (c.MB * kHaloPix);
    const double epi = (double)c.MB * (c.BN / 32) * m.epi * c.nacc * (c.stack ? 1.5 : 1.0);
    double t = mma > bytes / 32.0 ? mma : bytes / 32.0;
    t *= 1.0 + 0.3 / c.a_stages;
    if (c.nbuf == 1) t += 3.0 * epi;
    else if (epi > t) t = epi;
    return (double)waves * t * 0.85;       // streaming plans of these layers measure ~15 % above their model (L2 cap)
  }
  const long long items = (m_tiles + c.MB - 1) / c.MB * (cout / c.BN);
  const long long waves = (items + kNumSMs - 1) / kNumSMs;
  // an MMA of 128 x BN x 32 B: tensor rate (BN/2 clk) vs operand fetch from shared memory, plus issue overhead
  auto mma_clk = [&](double n) {
    const double fetch = (m.sm_a + n) / m.sm_d;
    return (n / 2.0 > fetch ? n / 2.0 : fetch) + m.issue;
  };
  const double per_step = npl == 1 ? mma_clk(c.BN) : c.stack ? mma_clk(2.0 * c.BN) + mma_clk(c.BN) : 3.0 * mma_clk(c.BN);
  const double mma = (double)c.MB * 9 * n_cchunks * nks * per_step + 9.0 * n_cchunks * m.stage_ovh;
  const double bytes = (double)n_cchunks * npl * c.row_bytes * (c.MB * kHaloPix + 9.0 * c.BN);
  const double epi = (double)c.MB * (c.BN / 32) * m.epi * c.nacc * (c.stack ? 1.5 : 1.0);
  double t = mma > bytes / m.bw ? mma : bytes / m.bw;
  t *= 1.0 + m.pb / c.b_stages + m.pa / c.a_stages;            // shallow rings expose L2 latency
  if (c.nbuf == 1) t += 3.0 * epi;      // single TMEM buffer: the tensor pipe drains while the epilogue runs (measured)
  else if (epi > t) t = epi;
  return (double)waves * t;
}

// Measured tilings (tools/halo_probe.py --sweep-full on B200 -> tools/make_plan_table.py): the best (cout tile, pixel-tile
// blocking, stacked planes, row width) for the layer shapes of fuseunet / UNet at the batches the AIDE step launches.
struct PlanEntry {
  int fmt, cin, cout, m_tiles, BN, MB, stack, rb, occ;
};
constexpr PlanEntry kPlanTable[] = {
#include "conv_plan_table.inc"
};

const PlanEntry* lookup_plan(int fmt, int cin, int cout, long long m_tiles, bool allow_pair) {
  const PlanEntry* hit = nullptr;
  double best_ratio = 1.26;                       // same layer, tile count within 25 %: the same tiling regime
  for (const PlanEntry& e : kPlanTable) {
    if (e.fmt != fmt || e.cin != cin || e.cout != cout) continue;
    if (e.occ == 4 && !allow_pair) continue;
    const double r = (double)m_tiles > e.m_tiles ? (double)m_tiles / e.m_tiles : (double)e.m_tiles / (double)m_tiles;
    // same distance: a pair entry (listed next to the best single-CTA tiling of the shape) wins
    if (r < best_ratio || (hit && r == best_ratio && e.occ == 4)) {
      best_ratio = r;
      hit = &e;
    }
  }
  return hit;
}

bool make_plan(int fmt, int cin, int cout, long long m_tiles, HaloPlan* best, bool use_table = true, bool allow_pair = true) {
  const int es = fmt_elem_bytes(fmt), npl = fmt_planes(fmt);
  int force_bn = env_int("AIDE_CONV_BN", 0), force_mb = env_int("AIDE_CONV_MB", 0);
  int force_stack = env_int("AIDE_CONV_STACK", -1), force_rb = env_int("AIDE_CONV_RB", 0);
  // resident-weight plans (-1 planner's choice, 0 never, 1 only those).  Default 0: measured slower than streaming on
  // every layer they fit (profiles/r2c_*: 64->64 @256 235 vs 311 TFLOP/s) -- the Cout <= 64 layers are bound by
  // shared-memory operand reads per MMA, not by the weight re-fetch.
  const int force_res = env_int("AIDE_CONV_WRES", 0);
  // Two CTAs per SM (half the shared memory, <= 256 TMEM columns each): a second resident CTA issues MMAs while the
  // first waits on its barriers -- pays on the narrow layers whose tensor pipe idles between short MMAs.  1 / 2 force it.
  // 4: the CTA-pair kernel (cta_group::2, conv3x3_halo2_tc_kernel) -- table / env only, like 2.
  int force_occ = env_int("AIDE_CONV_OCC", 0);
  if (force_occ == 4 && (!allow_pair || fmt == AIDE_FMT_TF32X2)) force_occ = 0;
  if (!env_int("AIDE_CONV_PAIR", 1)) allow_pair = false;
  if (use_table && !force_bn && !force_mb && force_stack < 0 && !force_rb && force_res == 0 && !force_occ &&
      env_int("AIDE_CONV_TABLE", 1)) {
    if (const PlanEntry* e = lookup_plan(fmt, cin, cout, m_tiles, allow_pair)) {
      // build exactly the measured tiling through the same enumeration (forced parameters); fall back to the model
      // if it does not fit (cannot happen for the shapes it was measured on)
      force_bn = e->BN; force_mb = e->MB; force_stack = e->stack; force_rb = e->rb; force_occ = e->occ;
    }
  }
  best->cost = -1;
  for (int bn = 256; bn >= 32; bn >>= 1) {
    if (cout % bn) continue;
    if (force_bn && bn != force_bn) continue;
    for (int mb = 4; mb >= 1; mb >>= 1) {
      if (force_mb && mb != force_mb) continue;
      HaloPlan pick{};
      double pick_row_cost = -1;
      for (int rb = 128; rb >= 64; rb >>= 1) {
        const int kc = rb / es;
        if (cin % kc) continue;
        if (force_rb && rb != force_rb) continue;
        const int n_cchunks = cin / kc;
        const int nks = rb / 32;
       for (int occ = 1; occ <= 4; occ <<= 1) {
       if ((force_occ ? force_occ : 1) != occ) continue;          // the cost model never picks 2 by itself: table / env only
       for (int stack = 0; stack <= 1; ++stack) {
        if (stack && (npl != 2 || 2 * bn > 256)) continue;
        if (occ == 4 && stack != (npl == 2)) continue;            // the pair kernel: stacked planes only
        if (force_stack >= 0 && stack != force_stack && !(stack == 0 && (npl != 2 || 2 * bn > 256))) continue;
        // accumulation chain per TMEM column of the LARGE partial sum: 3 MMAs per 32-byte K step, 2 when the weight
        // planes are stacked, 1 when both cross terms go to the small accumulator (lo_col; measured on the 256x256
        // known-answer forward: logits vs fp64 4.1e-5 -> 2.3e-5, same value with a cap of 2 accumulators as with 4)
        const int losep = env_int("AIDE_CONV_LOSEP", 1);
        const long long chain = (long long)9 * n_cchunks * nks * (npl == 2 ? (stack ? (losep ? 1 : 2) : 3) : 1);
        const int nacc = wanted_nacc(fmt, chain);
        const int acc_w = bn * (1 + stack);
        const int tmem_max = occ == 2 ? 256 : 512;
        if (mb * nacc * acc_w > tmem_max) continue;
        HaloPlan c{};
        c.BN = bn; c.MB = mb; c.nacc = nacc; c.row_bytes = rb; c.stack = stack; c.occ = occ;
        c.nbuf = (2 * mb * nacc * acc_w <= tmem_max) ? 2 : 1;
        c.a_slot = (kHaloPix * rb + 1023) / 1024 * 1024;
        c.a_stage = npl * mb * c.a_slot;
        c.b_plane = bn * rb;
        c.b_stage = npl * c.b_plane;
        if (occ == 4) {                                            // per CTA: Y = one full plane (or half of the only one), X = half
          c.yx = npl == 2 ? c.b_plane : 0;
          c.b_stage = c.yx + c.b_plane / 2;
        }
        // 1 KB per CTA is reserved by the system; single-CTA plans stop at 226 KB so that one more CTA without shared
        // memory of its own -- the peer-memory all-reduce (comm.cu) -- can share the SM with a persistent conv CTA
        const int avail = (occ == 2 ? (kSmemMax - 2048) / 2 : kSmemMax - 1024) - 1024 - kBarBytes;
        c.a_stages = 2;
        int rest = avail - c.a_stages * c.a_stage;
        if (rest < 2 * c.b_stage) {                                // try a single halo stage before giving up
          c.a_stages = 1;
          rest = avail - c.a_stage;
        }
        if (rest < 2 * c.b_stage) continue;      // (a resident plan needs even more room)
        c.b_stages = rest / c.b_stage;
        if (c.b_stages > kMaxBStages) c.b_stages = kMaxBStages;
        if (c.a_stages == 2 && n_cchunks > 2 && c.b_stages > 4 && rest - 4 * c.b_stage >= c.a_stage) {
          c.a_stages = 3;                                          // spare room: a third halo stage instead of > 4 weight stages
          c.b_stages = (avail - 3 * c.a_stage) / c.b_stage;
          if (c.b_stages > kMaxBStages) c.b_stages = kMaxBStages;
        }
        c.smem = 1024 + c.a_stages * c.a_stage + c.b_stages * c.b_stage + kBarBytes;
        if (force_res != 1) {
          const double row_cost = model_cost(kRow, c, npl, n_cchunks, nks, m_tiles, cout);
          if (pick_row_cost < 0 || row_cost < pick_row_cost * 0.999) {
            pick_row_cost = row_cost;
            pick = c;
            pick.cost = model_cost(kSel, c, npl, n_cchunks, nks, m_tiles, cout);
          }
        }
        // resident-weight variant: the whole weight matrix of the layer next to >= 2 halo stages
        const int wbytes = 9 * n_cchunks * c.b_stage;
        if (force_res != 0 && occ != 4 && bn == cout && wbytes + 2 * c.a_stage <= avail) {
          HaloPlan r = c;
          r.resident = 1;
          r.a_stages = (avail - wbytes) / c.a_stage;
          if (r.a_stages > kMaxAStages) r.a_stages = kMaxAStages;
          r.b_stages = 9 * n_cchunks;
          r.smem = 1024 + r.a_stages * r.a_stage + wbytes + kBarBytes;
          const double row_cost = model_cost(kRow, r, npl, n_cchunks, nks, m_tiles, cout);
          if (pick_row_cost < 0 || row_cost < pick_row_cost * 0.999) {
            pick_row_cost = row_cost;
            pick = r;
            pick.cost = model_cost(kSel, r, npl, n_cchunks, nks, m_tiles, cout);
          }
        }
       }
       }
      }
      if (pick_row_cost < 0) continue;
      if (best->cost < 0 || pick.cost < best->cost * 0.999) *best = pick;
    }
  }
  if (best->cost < 0 && use_table) return make_plan(fmt, cin, cout, m_tiles, best, false, allow_pair);
  return best->cost >= 0;
}

template <int KIND, int NKS>
int launch2(const HaloParams& p, int grid, int smem, cudaStream_t st) {
  static thread_local bool done[16] = {false};
  int dev = 0;
  AIDE_CUDA(cudaGetDevice(&dev));
  if (dev >= 16 || !done[dev]) {
    AIDE_CUDA(cudaFuncSetAttribute(conv3x3_halo_tc_kernel<KIND, NKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    if (dev < 16) done[dev] = true;
  }
  conv3x3_halo_tc_kernel<KIND, NKS><<<grid, kThreads, smem, st>>>(p);
  AIDE_CHECK_LAUNCH();
  return 0;
}
template <int KIND>
int launch(const HaloParams& p, int grid, int smem, cudaStream_t st) {
  return p.row_bytes == 128 ? launch2<KIND, 4>(p, grid, smem, st) : launch2<KIND, 2>(p, grid, smem, st);
}

template <int KIND, int NKS>
int launch_pair2(const Halo2Params& p, int grid, int smem, cudaStream_t st) {
  static thread_local bool done[16] = {false};
  int dev = 0;
  AIDE_CUDA(cudaGetDevice(&dev));
  if (dev >= 16 || !done[dev]) {
    AIDE_CUDA(cudaFuncSetAttribute(conv3x3_halo2_tc_kernel<KIND, NKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    if (dev < 16) done[dev] = true;
  }
  conv3x3_halo2_tc_kernel<KIND, NKS><<<grid, kThreads, smem, st>>>(p);      // __cluster_dims__(2,1,1): grid is even
  AIDE_CHECK_LAUNCH();
  return 0;
}
template <int KIND>
int launch_pair(const Halo2Params& p, int grid, int smem, cudaStream_t st) {
  return p.h.row_bytes == 128 ? launch_pair2<KIND, 4>(p, grid, smem, st) : launch_pair2<KIND, 2>(p, grid, smem, st);
}

}  // namespace

// The halo kernel needs whole 8-pixel output rows per UMMA core-matrix group; tiny feature maps (tests at 32x32
// inputs reach 2x2) stay on the first-generation kernel.
bool halo_shape_ok(int fmt, int cin, int cout, int N, int H, int W) {
  if (fmt != AIDE_FMT_TF32X2 && fmt != AIDE_FMT_BF16 && fmt != AIDE_FMT_F16X2) return false;
  if (env_int("AIDE_CONV_HALO", 1) == 0) return false;
  if (cin % 32 || cout % 32 || cin < 32 || cout < 32) return false;
  if (W < kTW || H < kTH / 2) return false;
  HaloPlan pl;
  const long long m_tiles = (long long)N * ceil_div(W, kTW) * ceil_div(H, kTH);
  return make_plan(fmt, cin, cout, m_tiles, &pl);
}

int halo_stat_rows(int N, int H, int W) { return 4 * N * ceil_div(W, kTW) * ceil_div(H, kTH); }

int halo_conv3x3(int fmt, const void* x0, const void* x1, int x_ctot, int x_coff, int cin, const void* w0, const void* w1,
                 const float* bias, float* z, int z_ctot, int z_coff, int cout, int N, int H, int W, float* stat_partial,
                 float out_scale, const float* out_scale_ptr, cudaStream_t st, const HaloFused* fused) {
  const int kind = kind_of(fmt);
  const int es = fmt_elem_bytes(fmt), npl = fmt_planes(fmt);
  const int dtype = fmt == AIDE_FMT_BF16 ? 1 : fmt == AIDE_FMT_F16X2 ? 2 : 0;
  AIDE_REQUIRE(npl == 1 || (x1 && w1), "conv3x3(halo): two-plane operand formats need hi and lo planes");
  HaloParams p{};
  p.tiles_w = ceil_div(W, kTW);
  p.tiles_h = ceil_div(H, kTH);
  p.m_tiles = N * p.tiles_w * p.tiles_h;
  HaloPlan pl;
  AIDE_REQUIRE(make_plan(fmt, cin, cout, p.m_tiles, &pl), "conv3x3(halo): no tiling fits (cin=%d cout=%d)", cin, cout);
  p.BN = pl.BN; p.MB = pl.MB; p.nacc = pl.nacc; p.nbuf = pl.nbuf;
  p.stack = pl.stack; p.acc_w = pl.BN * (1 + pl.stack);
  p.lo_col = (pl.stack && env_int("AIDE_CONV_LOSEP", 1)) ? pl.BN : 0;
  p.row_bytes = pl.row_bytes;
  p.kc = pl.row_bytes / es;
  p.n_cchunks = cin / p.kc;
  p.n_tiles = cout / p.BN;
  p.total_items = ceil_div(p.m_tiles, p.MB) * p.n_tiles;
  p.a_slot_bytes = pl.a_slot; p.a_stage_bytes = pl.a_stage; p.a_stages = pl.a_stages;
  p.b_plane_bytes = pl.b_plane; p.b_stage_bytes = pl.b_stage;
  p.b_resident = pl.resident;
  p.b_stages = pl.resident ? 1 : pl.b_stages;          // ring depth (barriers); resident: one "all weights landed" barrier
  p.b_off = pl.a_stages * pl.a_stage;
  p.bar_off = p.b_off + pl.b_stages * pl.b_stage;
  int cols = p.nbuf * p.MB * p.nacc * p.acc_w;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  p.z = z; p.bias = bias; p.stat_partial = stat_partial;
  if (fused) {
    AIDE_REQUIRE(fused->ss && (fused->dst.p0 || fused->pa.p0 || fused->pb.p0), "conv3x3(fused): no destination");
    AIDE_REQUIRE(!(fused->pa.p0 || fused->pb.p0) || (H % 2 == 0 && W % 2 == 0), "conv3x3(fused): pooling needs even H, W");
    p.ep_ss = fused->ss;
    p.y0 = fused->dst.p0; p.y1 = fused->dst.p1; p.y_ctot = fused->dst.ctot; p.y_coff = fused->dst.coff;
    p.pa0 = fused->pa.p0; p.pa1 = fused->pa.p1; p.pa_ctot = fused->pa.ctot; p.pa_coff = fused->pa.coff;
    p.pb0 = fused->pb.p0; p.pb1 = fused->pb.p1; p.pb_ctot = fused->pb.ctot; p.pb_coff = fused->pb.coff;
  }
  p.out_scale = out_scale; p.out_scale_ptr = out_scale_ptr;
  p.z_ctot = z_ctot; p.z_coff = z_coff; p.cout = cout; p.cin = cin; p.H = H; p.W = W;
  AIDE_REQUIRE(pl.smem <= kSmemMax, "conv3x3(halo): shared memory %d too large", pl.smem);

  if (act_tmap(&p.tmA0, dtype, x0, x_ctot, x_coff, cin, N, H, W, p.kc, kHW, kHH, p.row_bytes)) return 1;
  if (mat_tmap(&p.tmB0, dtype, w0, cout, 9 * cin, p.kc, p.BN, p.row_bytes)) return 1;
  if (npl == 2) {
    if (act_tmap(&p.tmA1, dtype, x1, x_ctot, x_coff, cin, N, H, W, p.kc, kHW, kHH, p.row_bytes)) return 1;
    if (mat_tmap(&p.tmB1, dtype, w1, cout, 9 * cin, p.kc, p.BN, p.row_bytes)) return 1;
  }
  if (pl.occ == 4) {
    Halo2Params pp{};
    pp.h = p;
    pp.yx_bytes = pl.yx;
    pp.pair_items = ceil_div(p.m_tiles, 2 * p.MB) * p.n_tiles;
    if (mat_tmap(&pp.tmBh, dtype, w0, cout, 9 * cin, p.kc, p.BN / 2, p.row_bytes)) return 1;
    const int pairs = pp.pair_items < kNumSMs / 2 ? pp.pair_items : kNumSMs / 2;
    if (kind == K_BF16) return launch_pair<K_BF16>(pp, 2 * pairs, pl.smem, st);
    return launch_pair<K_F16X2>(pp, 2 * pairs, pl.smem, st);
  }
  const int max_ctas = kNumSMs * (pl.occ == 2 ? 2 : 1);
  const int grid = p.total_items < max_ctas ? p.total_items : max_ctas;
  if (kind == K_BF16) return launch<K_BF16>(p, grid, pl.smem, st);
  if (kind == K_F16X2) return launch<K_F16X2>(p, grid, pl.smem, st);
  return launch<K_TF32X2>(p, grid, pl.smem, st);
}

// for tools/ and bench.py: the tiling chosen for a layer
extern "C" int aide_conv3x3_plan_info(int fmt, int cin, int cout, int N, int H, int W, int* out /*[9]*/) {
  if (!halo_shape_ok(fmt, cin, cout, N, H, W)) return 1;
  HaloPlan pl;
  make_plan(fmt, cin, cout, (long long)N * ceil_div(W, kTW) * ceil_div(H, kTH), &pl);
  out[0] = pl.BN; out[1] = pl.MB; out[2] = pl.nacc; out[3] = pl.nbuf; out[4] = pl.row_bytes; out[5] = pl.a_stages;
  out[6] = pl.b_stages; out[7] = pl.smem; out[8] = pl.stack + 2 * pl.resident + 4 * (pl.occ == 2) + 8 * (pl.occ == 4);
  return 0;
}

}  // namespace aide
