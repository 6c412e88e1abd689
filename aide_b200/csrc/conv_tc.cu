// conv_tc.cu -- conv3x3 (stride 1, pad 1) forward / dgrad / wgrad as im2col-free implicit GEMM on the
// Blackwell tensor cores: TMA (cp.async.bulk.tensor) stages NHWC activation boxes and weight slices
// into 128B/64B-swizzled shared memory, one elected thread issues tcgen05.mma into a TMEM fp32
// accumulator, four epilogue warps drain it with tcgen05.ld.
//
//   forward / dgrad :  D[128 pixels, BN cout] = sum_{tap, c-slice} A_tap[128 pixels, kc] * W_tap[BN, kc]^T
//       A_tap is ONE 4-D TMA box (kc channels x TW x TH pixels x 1 image) loaded at spatial offset
//       (dy,dx); the conv's zero padding is TMA's out-of-bounds zero fill.  Both operands K-major.
//   wgrad           :  D[128 rows of (tap,ci), BN cout] = sum_{pixel blocks} X_tap[KP pixels, 128]^T * dZ[KP pixels, BN]
//       both operands MN-major (channels contiguous), split-K over pixel blocks, fixed-order reduce.
//
//   BF16   ("fast")   : one kind::f16 MMA per K-step.
//   TF32X2 ("parity") : operands split hi = rn_tf32(x), lo = x - hi; three kind::tf32 MMAs
//                       lo*hi + hi*lo + hi*hi accumulate in fp32 -> fp32-equivalent products
//                       (2.0e-5 rel, 0 argmax flips vs 2e-2 / 1428 for single TF32).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4).
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace aide {

using namespace ptx;

// ------------------------------------------------------------------ driver entry point for tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

constexpr int kSwizzle128Atom32 = 1128;

static int encode_tmap(CUtensorMap* m, bool bf16, int rank, const void* base, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  AIDE_REQUIRE(fn, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  // swizzle_bytes: 128 / 64 / 32 = classic 16B-chunk swizzles; kSwizzle128Atom32 = 128B span with 32B atoms
  // (the only shared-memory layout tcgen05 accepts for MN-major 32-bit (tf32) operands)
  CUtensorMapSwizzle sw = swizzle_bytes == kSwizzle128Atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                          : swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                  const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AIDE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d box %u,%u swizzle %d)", (int)r,
               rank, box[0], box[1], swizzle_bytes);
  return 0;
}

// NHWC activation view as a 4-D tensor (C_view, W, H, N); box = (box_c, box_w, box_h, 1)
static int act_tmap(CUtensorMap* m, bool bf16, const void* plane, int ctot, int coff, int C, int N, int H, int W,
                    int box_c, int box_w, int box_h, int swizzle_bytes) {
  const size_t es = bf16 ? 2 : 4;
  const char* base = reinterpret_cast<const char*>(plane) + (size_t)coff * es;
  AIDE_REQUIRE(((uintptr_t)base % 16) == 0 && ((size_t)ctot * es) % 16 == 0,
               "activation view is not 16-byte aligned (ctot=%d coff=%d)", ctot, coff);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t str[3] = {(cuuint64_t)ctot * es, (cuuint64_t)W * ctot * es, (cuuint64_t)H * W * ctot * es};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  return encode_tmap(m, bf16, 4, base, dims, str, box, swizzle_bytes);
}

// weights [rows][kdim] row-major as a 2-D tensor (kdim, rows); box = (box_k, box_rows)
static int mat_tmap(CUtensorMap* m, bool bf16, const void* plane, int rows, int kdim, int box_k, int box_rows,
                    int swizzle_bytes) {
  const size_t es = bf16 ? 2 : 4;
  AIDE_REQUIRE(((uintptr_t)plane % 16) == 0 && ((size_t)kdim * es) % 16 == 0, "weight plane is not 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)kdim * es};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  return encode_tmap(m, bf16, 2, plane, dims, str, box, swizzle_bytes);
}

// pixel tile TW x TH (powers of two, TW*TH = npix) minimising the number of tiles covering H x W
static void pick_tile(int H, int W, int npix, int* TW, int* TH) {
  long long best = -1;
  for (int tw = npix; tw >= 1; tw >>= 1) {
    int th = npix / tw;
    if (tw > 256 || th > 256) continue;
    long long tiles = (long long)ceil_div(W, tw) * ceil_div(H, th);
    if (best < 0 || tiles < best) {
      best = tiles;
      *TW = tw;
      *TH = th;
    }
  }
}
static int ilog2(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
}
static uint32_t tmem_cols(int n) {
  uint32_t c = 32;
  while ((int)c < n) c <<= 1;
  return c;
}

constexpr int kThreads = 192;

// ================================================================================================
// forward / dgrad
// ================================================================================================
struct FwdParams {
  CUtensorMap tmA0, tmA1, tmB0, tmB1;
  float* z;
  const float* bias;
  float* stat_partial;
  int z_ctot, z_coff, cout, cin, H, W;
  int TW, TH, tw_log2, tiles_w, tiles_h;
  int BN, kc, n_cchunks, row_bytes, stages;
  int a_plane_bytes, b_plane_bytes, stage_bytes, data_bytes;
};

template <bool TF32, bool PARITY>
__global__ void __launch_bounds__(kThreads) conv3x3_tc_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  constexpr int NPL = PARITY ? 2 : 1;
  // tcgen05.mma truncates (rounds toward zero) when it adds into the TMEM accumulator, so the error of one
  // accumulation chain grows linearly with its length (tools/accum_probe.py: 1.7e-5 rms at K = 9216 vs 1e-6 for
  // fp32 FMA chains).  Parity mode therefore rotates the K blocks over NACC independent accumulators and adds
  // them with round-to-nearest fp32 in the epilogue.
  constexpr int NACC = PARITY ? 4 : 1;
  const int S = p.stages;
  const uint32_t bar_base = base + p.data_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * S);
  const uint32_t slot_addr = bar_base + 8u * (2 * S + 1);
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + p.data_bytes + 8 * (2 * S + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x;
  const int tw_i = m_tile % p.tiles_w, th_i = (m_tile / p.tiles_w) % p.tiles_h, n_img = m_tile / (p.tiles_w * p.tiles_h);
  const int h0 = th_i * p.TH, w0 = tw_i * p.TW, n0 = blockIdx.y * p.BN;
  const int num_kb = 9 * p.n_cchunks;
  const uint32_t ncols_need = (uint32_t)(NACC * p.BN);
  const uint32_t ncols = ncols_need <= 32 ? 32u : ncols_need <= 64 ? 64u : ncols_need <= 128 ? 128u : ncols_need <= 256 ? 256u : 512u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA0);
    tma_prefetch_desc(&p.tmB0);
    if (PARITY) {
      tma_prefetch_desc(&p.tmA1);
      tma_prefetch_desc(&p.tmB1);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(slot_addr, ncols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = NPL * (p.a_plane_bytes + p.b_plane_bytes);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S, ph = (kb / S) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(s), tx_bytes);
        const int tap = kb / p.n_cchunks, cc = kb - tap * p.n_cchunks;
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const uint32_t a_dst = base + s * p.stage_bytes;
        const uint32_t b_dst = a_dst + NPL * p.a_plane_bytes;
        tma_load_4d(a_dst, &p.tmA0, full_bar(s), cc * p.kc, w0 + dx, h0 + dy, n_img);
        tma_load_2d(b_dst, &p.tmB0, full_bar(s), tap * p.cin + cc * p.kc, n0);
        if (PARITY) {
          tma_load_4d(a_dst + p.a_plane_bytes, &p.tmA1, full_bar(s), cc * p.kc, w0 + dx, h0 + dy, n_img);
          tma_load_2d(b_dst + p.b_plane_bytes, &p.tmB1, full_bar(s), tap * p.cin + cc * p.kc, n0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TF32 ? 2u : 1u, 0u, 0u, 128u, (uint32_t)p.BN);
      const uint32_t layout = p.row_bytes == 128 ? 2u : 4u;
      const uint32_t sbo = 8u * p.row_bytes;
      const int nks = p.row_bytes / 32;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S, ph = (kb / S) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after_sync();
        const uint32_t a0 = base + s * p.stage_bytes;
        const uint32_t b0 = a0 + NPL * p.a_plane_bytes;
        const uint32_t acc = tmem_base + (uint32_t)((kb % NACC) * p.BN);
        uint32_t accum = kb >= NACC ? 1u : 0u;
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t ko = ks * 32;
          if (PARITY) {
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
      umma_commit(tmem_full_bar);
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: TMEM -> registers -> smem staging -> coalesced global + BN partial stats
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after_sync();
    const int hh = h0 + (row >> p.tw_log2), ww = w0 + (row & (p.TW - 1));
    const bool valid = hh < p.H && ww < p.W;
    float* stg = reinterpret_cast<float*>(sm);
    const int pitch = p.BN + 4;
    for (int ch = 0; ch < p.BN / 32; ++ch) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), r);
      tmem_ld_wait();
      if constexpr (NACC > 1) {
        // pairwise (a0 + a1) + (a2 + a3), round-to-nearest fp32
        uint32_t r1[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.BN + ch * 32), r1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r1[j]));
        uint32_t r2[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(2 * p.BN + ch * 32), r1);
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(3 * p.BN + ch * 32), r2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = __float_as_uint(__uint_as_float(r[j]) + (__uint_as_float(r1[j]) + __uint_as_float(r2[j])));
      }
      float* dst = stg + (size_t)row * pitch + ch * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                          __uint_as_float(r[j + 3]));
          if (p.bias) {
            const float* bp = p.bias + n0 + ch * 32 + j;  // scalar loads: parameter storage may be only 4B-aligned
            v.x += __ldg(bp); v.y += __ldg(bp + 1); v.z += __ldg(bp + 2); v.w += __ldg(bp + 3);
          }
        }
        *reinterpret_cast<float4*>(dst + j) = v;
      }
    }
    tc_fence_before_sync();
    named_bar_sync(1, 128);
    const int nq = p.BN >> 2;
    for (int i = et; i < 128 * nq; i += 128) {
      const int r = i / nq, c4 = i - r * nq;
      const int h2 = h0 + (r >> p.tw_log2), w2 = w0 + (r & (p.TW - 1));
      if (h2 < p.H && w2 < p.W) {
        const float4 v = *reinterpret_cast<const float4*>(stg + (size_t)r * pitch + c4 * 4);
        const size_t pix = ((size_t)n_img * p.H + h2) * p.W + w2;
        *reinterpret_cast<float4*>(p.z + pix * p.z_ctot + p.z_coff + n0 + c4 * 4) = v;
      }
    }
    if (p.stat_partial) {
      for (int col = et; col < p.BN; col += 128) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
        for (int r = 0; r < 128; ++r) {
          const float v = stg[(size_t)r * pitch + col];
          s1 += v;
          s2 = fmaf(v, v, s2);
        }
        p.stat_partial[((size_t)m_tile * 2 + 0) * p.cout + n0 + col] = s1;
        p.stat_partial[((size_t)m_tile * 2 + 1) * p.cout + n0 + col] = s2;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, ncols);
  }
}

// ================================================================================================
// wgrad
// ================================================================================================
struct WgradParams {
  CUtensorMap tmX0, tmX1, tmD0, tmD1;
  float* ws;
  int cin, cout, H, W;
  int TW, TH, tiles_w, tiles_h, KP;
  int kcA, rbA, boxesA, n_cchunksA, a_box_bytes;
  int kcB, rbB, boxesB, BN, b_box_bytes;
  int a_plane_bytes, b_plane_bytes, stage_bytes, stages, data_bytes;
  int kb_total, kb_per_split;
};

template <bool TF32, bool PARITY>
__global__ void __launch_bounds__(kThreads) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  constexpr int NPL = PARITY ? 2 : 1;
  const int S = p.stages;
  const uint32_t bar_base = base + p.data_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * S);
  const uint32_t slot_addr = bar_base + 8u * (2 * S + 1);
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(sm + p.data_bytes + 8 * (2 * S + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, n0 = blockIdx.y * p.BN, split = blockIdx.z;
  const int kb_begin = split * p.kb_per_split;
  const int kb_end = min(p.kb_total, kb_begin + p.kb_per_split);
  const uint32_t ncols = p.BN <= 32 ? 32u : p.BN <= 64 ? 64u : p.BN <= 128 ? 128u : 256u;
  int nvalidA = 0;
  for (int j = 0; j < p.boxesA; ++j) nvalidA += ((mt * p.boxesA + j) / p.n_cchunksA) < 9 ? 1 : 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX0);
    tma_prefetch_desc(&p.tmD0);
    if (PARITY) {
      tma_prefetch_desc(&p.tmX1);
      tma_prefetch_desc(&p.tmD1);
    }
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(slot_addr, ncols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = NPL * (nvalidA * p.a_box_bytes + p.boxesB * p.b_box_bytes);
      int it = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const int s = it % S, ph = (it / S) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(s), tx_bytes);
        const int tw_i = kb % p.tiles_w, th_i = (kb / p.tiles_w) % p.tiles_h, n_img = kb / (p.tiles_w * p.tiles_h);
        const int h0 = th_i * p.TH, w0 = tw_i * p.TW;
        const uint32_t a_dst = base + s * p.stage_bytes;
        const uint32_t b_dst = a_dst + NPL * p.a_plane_bytes;
        for (int j = 0; j < p.boxesA; ++j) {
          const int gb = mt * p.boxesA + j;
          const int tap = gb / p.n_cchunksA, cc = gb - tap * p.n_cchunksA;
          if (tap >= 9) continue;
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          tma_load_4d(a_dst + j * p.a_box_bytes, &p.tmX0, full_bar(s), cc * p.kcA, w0 + dx, h0 + dy, n_img);
          if (PARITY)
            tma_load_4d(a_dst + p.a_plane_bytes + j * p.a_box_bytes, &p.tmX1, full_bar(s), cc * p.kcA, w0 + dx,
                        h0 + dy, n_img);
        }
        for (int j = 0; j < p.boxesB; ++j) {
          tma_load_4d(b_dst + j * p.b_box_bytes, &p.tmD0, full_bar(s), n0 + j * p.kcB, w0, h0, n_img);
          if (PARITY)
            tma_load_4d(b_dst + p.b_plane_bytes + j * p.b_box_bytes, &p.tmD1, full_bar(s), n0 + j * p.kcB, w0, h0,
                        n_img);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TF32 ? 2u : 1u, 1u, 1u, 128u, (uint32_t)p.BN);
      // MN-major operands.  16-bit: SWIZZLE_128B / SWIZZLE_64B atoms of 8 K-rows.  32-bit (tf32): the only legal
      // layout is SWIZZLE_128B_BASE32B (type 1): 128-byte rows, 32-byte swizzle atoms, 4 K-rows per atom.
      const uint32_t layA = TF32 ? 1u : (p.rbA == 128 ? 2u : 4u), layB = TF32 ? 1u : (p.rbB == 128 ? 2u : 4u);
      const uint32_t krows = TF32 ? 4u : 8u;
      const int kel = TF32 ? 8 : 16;  // pixels (K elements) per MMA
      const int nks = p.KP / kel;
      uint32_t accum = 0;
      int it = 0;
      for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
        const int s = it % S, ph = (it / S) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after_sync();
        const uint32_t a0 = base + s * p.stage_bytes;
        const uint32_t b0 = a0 + NPL * p.a_plane_bytes;
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t ao = ks * kel * p.rbA, bo = ks * kel * p.rbB;
          const uint64_t ah = make_smem_desc(a0 + ao, p.a_box_bytes, krows * p.rbA, layA);
          const uint64_t bh = make_smem_desc(b0 + bo, p.b_box_bytes, krows * p.rbB, layB);
          if (PARITY) {
            const uint64_t al = make_smem_desc(a0 + p.a_plane_bytes + ao, p.a_box_bytes, krows * p.rbA, layA);
            const uint64_t bl = make_smem_desc(b0 + p.b_plane_bytes + bo, p.b_box_bytes, krows * p.rbB, layB);
            umma<TF32>(tmem_base, al, bh, idesc, accum);
            umma<TF32>(tmem_base, ah, bl, idesc, 1u);
            umma<TF32>(tmem_base, ah, bh, idesc, 1u);
          } else {
            umma<TF32>(tmem_base, ah, bh, idesc, accum);
          }
          accum = 1u;
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(tmem_full_bar);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after_sync();
    const int j = row / p.kcA, ci_l = row - j * p.kcA;
    const int gb = mt * p.boxesA + j;
    const int tap = gb / p.n_cchunksA, cc = gb - tap * p.n_cchunksA;
    const bool valid = tap < 9;
    const int ci = cc * p.kcA + ci_l;
    float* out = p.ws + (size_t)split * 9 * p.cin * p.cout + ((size_t)tap * p.cin + ci) * p.cout + n0;
    for (int ch = 0; ch < p.BN / 32; ++ch) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          *reinterpret_cast<float4*>(out + ch * 32 + k) = make_float4(
              __uint_as_float(r[k]), __uint_as_float(r[k + 1]), __uint_as_float(r[k + 2]), __uint_as_float(r[k + 3]));
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, ncols);
  }
}

// ================================================================================================
// host side
// ================================================================================================
constexpr int kSmemBudget2 = 100 * 1024;  // two CTAs per SM
constexpr int kSmemBudget1 = 200 * 1024;  // one CTA per SM

// opt in to > 48 KB dynamic shared memory: once per (kernel, device); not a stream operation
template <typename K>
static int set_smem(K kernel, int bytes) {
  struct Seen { const void* fn; int dev; };
  static thread_local Seen seen[64];
  static thread_local int n_seen = 0;
  int dev = 0;
  AIDE_CUDA(cudaGetDevice(&dev));
  const void* fn = reinterpret_cast<const void*>(kernel);
  for (int i = 0; i < n_seen; ++i)
    if (seen[i].fn == fn && seen[i].dev == dev) return 0;
  AIDE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (n_seen < 64) seen[n_seen++] = Seen{fn, dev};
  return 0;
}

int tc_stat_rows(int N, int H, int W) {
  int TW = 128, TH = 1;
  pick_tile(H, W, 128, &TW, &TH);
  return N * ceil_div(W, TW) * ceil_div(H, TH);
}

bool tc_shape_ok(int fmt, int cin, int cout) {
  if (fmt != AIDE_FMT_TF32X2 && fmt != AIDE_FMT_BF16) return false;
  return cin % 32 == 0 && cout % 32 == 0 && cin >= 32 && cout >= 32;
}

int tc_conv3x3(int fmt, const void* x0, const void* x1, int x_ctot, int x_coff, int cin, const void* w0, const void* w1,
               const float* bias, float* z, int z_ctot, int z_coff, int cout, int N, int H, int W, float* stat_partial,
               cudaStream_t st) {
  const bool bf16 = fmt == AIDE_FMT_BF16;
  const int es = bf16 ? 2 : 4, npl = bf16 ? 1 : 2;
  FwdParams p{};
  p.row_bytes = (cin * es >= 128 && cin % (128 / es) == 0) ? 128 : 64;
  p.kc = p.row_bytes / es;
  AIDE_REQUIRE(cin % p.kc == 0, "conv3x3(tc): cin=%d not a multiple of the K slice %d", cin, p.kc);
  p.n_cchunks = cin / p.kc;
  p.BN = cout % 128 == 0 ? 128 : cout % 64 == 0 ? 64 : 32;
  pick_tile(H, W, 128, &p.TW, &p.TH);
  p.tw_log2 = ilog2(p.TW);
  p.tiles_w = ceil_div(W, p.TW);
  p.tiles_h = ceil_div(H, p.TH);
  p.z = z; p.bias = bias; p.stat_partial = stat_partial;
  p.z_ctot = z_ctot; p.z_coff = z_coff; p.cout = cout; p.cin = cin; p.H = H; p.W = W;
  p.a_plane_bytes = 128 * p.row_bytes;
  p.b_plane_bytes = p.BN * p.row_bytes;
  p.stage_bytes = npl * (p.a_plane_bytes + p.b_plane_bytes);
  const int staging = 128 * (p.BN + 4) * 4;
  int budget = (3 * p.stage_bytes <= kSmemBudget2 - 2048 && staging <= kSmemBudget2 - 2048) ? kSmemBudget2 : kSmemBudget1;
  p.stages = (budget - 2048) / p.stage_bytes;
  if (p.stages > 6) p.stages = 6;
  if (p.stages > 9 * p.n_cchunks) p.stages = 9 * p.n_cchunks;
  AIDE_REQUIRE(p.stages >= 2, "conv3x3(tc): stage too large (%d bytes)", p.stage_bytes);
  int data = p.stages * p.stage_bytes;
  if (data < staging) data = staging;
  p.data_bytes = (data + 1023) / 1024 * 1024;
  const int smem = p.data_bytes + 8 * (2 * p.stages + 2) + 1024;
  AIDE_REQUIRE(smem <= 227 * 1024, "conv3x3(tc): shared memory %d too large", smem);

  if (act_tmap(&p.tmA0, bf16, x0, x_ctot, x_coff, cin, N, H, W, p.kc, p.TW, p.TH, p.row_bytes)) return 1;
  if (mat_tmap(&p.tmB0, bf16, w0, cout, 9 * cin, p.kc, p.BN, p.row_bytes)) return 1;
  if (!bf16) {
    AIDE_REQUIRE(x1 && w1, "conv3x3(tc): TF32X2 needs hi and lo planes");
    if (act_tmap(&p.tmA1, false, x1, x_ctot, x_coff, cin, N, H, W, p.kc, p.TW, p.TH, p.row_bytes)) return 1;
    if (mat_tmap(&p.tmB1, false, w1, cout, 9 * cin, p.kc, p.BN, p.row_bytes)) return 1;
  }
  dim3 grid(N * p.tiles_w * p.tiles_h, cout / p.BN);
  if (bf16) {
    if (set_smem(conv3x3_tc_kernel<false, false>, 227 * 1024)) return 2;
    conv3x3_tc_kernel<false, false><<<grid, kThreads, smem, st>>>(p);
  } else {
    if (set_smem(conv3x3_tc_kernel<true, true>, 227 * 1024)) return 2;
    conv3x3_tc_kernel<true, true><<<grid, kThreads, smem, st>>>(p);
  }
  AIDE_CHECK_LAUNCH();
  return 0;
}

struct WgradPlan {
  WgradParams p;
  int mt, nt, splits, smem;
};

static int wgrad_plan(int fmt, int cin, int cout, int N, int H, int W, WgradPlan* o) {
  const bool bf16 = fmt == AIDE_FMT_BF16;
  const int es = bf16 ? 2 : 4, npl = bf16 ? 1 : 2;
  WgradParams& p = o->p;
  p = WgradParams{};
  p.cin = cin; p.cout = cout; p.H = H; p.W = W;
  p.KP = bf16 ? 64 : 32;
  pick_tile(H, W, p.KP, &p.TW, &p.TH);
  p.tiles_w = ceil_div(W, p.TW);
  p.tiles_h = ceil_div(H, p.TH);
  p.rbA = (cin * es >= 128 && cin % (128 / es) == 0) ? 128 : 64;
  p.kcA = p.rbA / es;
  p.boxesA = 128 / p.kcA;
  p.n_cchunksA = cin / p.kcA;
  p.a_box_bytes = p.KP * p.rbA;
  p.rbB = (cout * es >= 128 && cout % (128 / es) == 0) ? 128 : 64;
  p.kcB = p.rbB / es;
  p.BN = cout % 128 == 0 ? 128 : cout % 64 == 0 ? 64 : 32;
  p.boxesB = p.BN / p.kcB;
  p.b_box_bytes = p.KP * p.rbB;
  p.a_plane_bytes = p.boxesA * p.a_box_bytes;
  p.b_plane_bytes = p.boxesB * p.b_box_bytes;
  p.stage_bytes = npl * (p.a_plane_bytes + p.b_plane_bytes);
  int budget = (3 * p.stage_bytes <= kSmemBudget2 - 2048) ? kSmemBudget2 : kSmemBudget1;
  p.stages = (budget - 2048) / p.stage_bytes;
  if (p.stages > 6) p.stages = 6;
  AIDE_REQUIRE(p.stages >= 2, "conv3x3_wgrad(tc): stage too large (%d bytes)", p.stage_bytes);
  p.data_bytes = (p.stages * p.stage_bytes + 1023) / 1024 * 1024;
  o->smem = p.data_bytes + 8 * (2 * p.stages + 2) + 1024;
  p.kb_total = N * p.tiles_w * p.tiles_h;
  o->mt = ceil_div(9 * p.n_cchunksA, p.boxesA);
  o->nt = cout / p.BN;
  long long base_ctas = (long long)o->mt * o->nt;
  long long want = (2LL * kNumSMs + base_ctas - 1) / base_ctas;  // ~2 CTAs per SM in flight
  long long maxs = (p.kb_total + 3) / 4;                         // >= 4 pixel blocks per split
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  p.kb_per_split = ceil_div(p.kb_total, want);
  o->splits = ceil_div(p.kb_total, p.kb_per_split);
  return 0;
}

size_t tc_wgrad_workspace_bytes(int fmt, int cin, int cout, int N, int H, int W) {
  WgradPlan pl;
  if (wgrad_plan(fmt, cin, cout, N, H, W, &pl)) return 0;
  return (size_t)pl.splits * 9 * cin * cout * sizeof(float);
}

int launch_wgrad_reduce(const float* ws, int splits, int cout, int cin, int layout, float* dw, cudaStream_t st);

int tc_wgrad(int fmt, const void* x0, const void* x1, int x_ctot, int x_coff, int cin, const void* dz0, const void* dz1,
             int cout, int N, int H, int W, void* ws, size_t ws_bytes, float* dw, cudaStream_t st) {
  const bool bf16 = fmt == AIDE_FMT_BF16;
  WgradPlan pl;
  if (wgrad_plan(fmt, cin, cout, N, H, W, &pl)) return 1;
  WgradParams& p = pl.p;
  AIDE_REQUIRE(ws && ws_bytes >= (size_t)pl.splits * 9 * cin * cout * sizeof(float), "conv3x3_wgrad(tc): workspace too small");
  p.ws = reinterpret_cast<float*>(ws);
  const int swA = bf16 ? p.rbA : kSwizzle128Atom32, swB = bf16 ? p.rbB : kSwizzle128Atom32;
  if (!bf16) AIDE_REQUIRE(p.rbA == 128 && p.rbB == 128, "conv3x3_wgrad(tc): tf32 operands need 32-channel (128 B) rows");
  if (act_tmap(&p.tmX0, bf16, x0, x_ctot, x_coff, cin, N, H, W, p.kcA, p.TW, p.TH, swA)) return 1;
  if (act_tmap(&p.tmD0, bf16, dz0, cout, 0, cout, N, H, W, p.kcB, p.TW, p.TH, swB)) return 1;
  if (!bf16) {
    AIDE_REQUIRE(x1 && dz1, "conv3x3_wgrad(tc): TF32X2 needs hi and lo planes");
    if (act_tmap(&p.tmX1, false, x1, x_ctot, x_coff, cin, N, H, W, p.kcA, p.TW, p.TH, swA)) return 1;
    if (act_tmap(&p.tmD1, false, dz1, cout, 0, cout, N, H, W, p.kcB, p.TW, p.TH, swB)) return 1;
  }
  dim3 grid(pl.mt, pl.nt, pl.splits);
  if (bf16) {
    if (set_smem(wgrad_tc_kernel<false, false>, 227 * 1024)) return 2;
    wgrad_tc_kernel<false, false><<<grid, kThreads, pl.smem, st>>>(p);
  } else {
    if (set_smem(wgrad_tc_kernel<true, true>, 227 * 1024)) return 2;
    wgrad_tc_kernel<true, true><<<grid, kThreads, pl.smem, st>>>(p);
  }
  AIDE_CHECK_LAUNCH();
  return launch_wgrad_reduce(p.ws, pl.splits, cout, cin, 1, dw, st);
}

bool tma_available() { return get_encode_fn() != nullptr; }

}  // namespace aide
