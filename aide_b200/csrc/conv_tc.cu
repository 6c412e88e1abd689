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

#include <unordered_map>

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

// dtype: 0 = fp32 (also tf32), 1 = bf16, 2 = fp16
static int encode_tmap(CUtensorMap* m, int dtype, int rank, const void* base, const cuuint64_t* dims,
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
  const CUtensorMapDataType dt = dtype == 1   ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                 : dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                              : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  // Encoded maps are cached per host thread, keyed on everything that defines them (the module path of
  // an unmodified training script re-issues the same ~100 conv calls per forward on the same arena addresses).
  struct Key {
    const void* base;
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4];
    int dtype, rank, swizzle;
    bool operator==(const Key& o) const { return std::memcmp(this, &o, sizeof(Key)) == 0; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
      uint64_t h = 1469598103934665603ull;
      for (size_t i = 0; i < sizeof(Key) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
      return (size_t)h;
    }
  };
  static_assert(sizeof(Key) % 8 == 0, "Key is hashed as 64-bit words");
  static thread_local std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  Key key;
  std::memset(&key, 0, sizeof(key));
  key.base = base;
  for (int i = 0; i < rank; ++i) {
    key.dims[i] = dims[i];
    key.box[i] = box[i];
    if (i + 1 < rank) key.strides[i] = strides_bytes[i];
  }
  key.dtype = dtype; key.rank = rank; key.swizzle = swizzle_bytes;
  auto it = cache.find(key);
  if (it != cache.end()) {
    *m = it->second;
    return 0;
  }
  CUresult r = fn(m, dt, (cuuint32_t)rank,
                  const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AIDE_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d box %u,%u swizzle %d)", (int)r,
               rank, box[0], box[1], swizzle_bytes);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *m);
  return 0;
}

// NHWC activation view as a 4-D tensor (C_view, W, H, N); box = (box_c, box_w, box_h, 1)
int act_tmap(CUtensorMap* m, int dtype, const void* plane, int ctot, int coff, int C, int N, int H, int W,
             int box_c, int box_w, int box_h, int swizzle_bytes) {
  const size_t es = dtype ? 2 : 4;
  const char* base = reinterpret_cast<const char*>(plane) + (size_t)coff * es;
  AIDE_REQUIRE(((uintptr_t)base % 16) == 0 && ((size_t)ctot * es) % 16 == 0,
               "activation view is not 16-byte aligned (ctot=%d coff=%d)", ctot, coff);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t str[3] = {(cuuint64_t)ctot * es, (cuuint64_t)W * ctot * es, (cuuint64_t)H * W * ctot * es};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  return encode_tmap(m, dtype, 4, base, dims, str, box, swizzle_bytes);
}

// weights [rows][kdim] row-major as a 2-D tensor (kdim, rows); box = (box_k, box_rows)
int mat_tmap(CUtensorMap* m, int dtype, const void* plane, int rows, int kdim, int box_k, int box_rows,
             int swizzle_bytes) {
  const size_t es = dtype ? 2 : 4;
  AIDE_REQUIRE(((uintptr_t)plane % 16) == 0 && ((size_t)kdim * es) % 16 == 0, "weight plane is not 16-byte aligned");
  cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)kdim * es};
  cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  return encode_tmap(m, dtype, 2, plane, dims, str, box, swizzle_bytes);
}

// pixel tile TW x TH (powers of two, TW*TH = npix) minimising the number of tiles covering H x W
void pick_tile(int H, int W, int npix, int* TW, int* TH) {
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
int ilog2(int v) {
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
      const uint32_t idesc = make_idesc(TF32 ? 2u : (PARITY ? 0u : 1u), 1u, 1u, 128u, (uint32_t)p.BN);   // tf32 | fp16 (F16X2) | bf16
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
int set_smem(K kernel, int bytes) {
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

struct WgradPlan {
  WgradParams p;
  int mt, nt, splits, smem;
};

static int wgrad_plan(int fmt, int cin, int cout, int N, int H, int W, WgradPlan* o) {
  const bool bf16 = fmt != AIDE_FMT_TF32X2;            // 16-bit operands (BF16 one plane, F16X2 two planes)
  const int es = fmt_elem_bytes(fmt), npl = fmt_planes(fmt);
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

int launch_wgrad_reduce(const float* ws, int splits, int cout, int cin, int layout, float* dw, float scale,
                        const float* scale_ptr, cudaStream_t st);

int tc_wgrad(int fmt, const void* x0, const void* x1, int x_ctot, int x_coff, int cin, const void* dz0, const void* dz1,
             int cout, int N, int H, int W, void* ws, size_t ws_bytes, float* dw, float out_scale,
             const float* out_scale_ptr, cudaStream_t st) {
  const bool bf16 = fmt != AIDE_FMT_TF32X2;
  const int npl = fmt_planes(fmt);
  const int dtype = fmt == AIDE_FMT_BF16 ? 1 : fmt == AIDE_FMT_F16X2 ? 2 : 0;
  WgradPlan pl;
  if (wgrad_plan(fmt, cin, cout, N, H, W, &pl)) return 1;
  WgradParams& p = pl.p;
  AIDE_REQUIRE(ws && ws_bytes >= (size_t)pl.splits * 9 * cin * cout * sizeof(float), "conv3x3_wgrad(tc): workspace too small");
  p.ws = reinterpret_cast<float*>(ws);
  const int swA = bf16 ? p.rbA : kSwizzle128Atom32, swB = bf16 ? p.rbB : kSwizzle128Atom32;
  if (!bf16) AIDE_REQUIRE(p.rbA == 128 && p.rbB == 128, "conv3x3_wgrad(tc): tf32 operands need 32-channel (128 B) rows");
  if (act_tmap(&p.tmX0, dtype, x0, x_ctot, x_coff, cin, N, H, W, p.kcA, p.TW, p.TH, swA)) return 1;
  if (act_tmap(&p.tmD0, dtype, dz0, cout, 0, cout, N, H, W, p.kcB, p.TW, p.TH, swB)) return 1;
  if (npl == 2) {
    AIDE_REQUIRE(x1 && dz1, "conv3x3_wgrad(tc): two-plane operand formats need hi and lo planes");
    if (act_tmap(&p.tmX1, dtype, x1, x_ctot, x_coff, cin, N, H, W, p.kcA, p.TW, p.TH, swA)) return 1;
    if (act_tmap(&p.tmD1, dtype, dz1, cout, 0, cout, N, H, W, p.kcB, p.TW, p.TH, swB)) return 1;
  }
  dim3 grid(pl.mt, pl.nt, pl.splits);
  if (fmt == AIDE_FMT_BF16) {
    if (set_smem(wgrad_tc_kernel<false, false>, 227 * 1024)) return 2;
    wgrad_tc_kernel<false, false><<<grid, kThreads, pl.smem, st>>>(p);
  } else if (fmt == AIDE_FMT_F16X2) {
    if (set_smem(wgrad_tc_kernel<false, true>, 227 * 1024)) return 2;
    wgrad_tc_kernel<false, true><<<grid, kThreads, pl.smem, st>>>(p);
  } else {
    if (set_smem(wgrad_tc_kernel<true, true>, 227 * 1024)) return 2;
    wgrad_tc_kernel<true, true><<<grid, kThreads, pl.smem, st>>>(p);
  }
  AIDE_CHECK_LAUNCH();
  return launch_wgrad_reduce(p.ws, pl.splits, cout, cin, 1, dw, out_scale, out_scale_ptr, st);
}

bool tma_available() { return get_encode_fn() != nullptr; }

}  // namespace aide
