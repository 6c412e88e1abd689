// loss.cu -- fused per-image CE + Dice (+ weighted-MSE consistency) reductions, their closed-form
// backward, pseudo-label generation and the small-loss (co-teaching) selection.
//
// Reference: utils/loss2d.py:5-13,35-61,87-154 (CrossEntropyLoss2d, DiceLoss, MulticlassDiceLoss,
// MulticlassMSELoss, CEMDiceLoss, CEMDiceLossImage), utils/metrics2d.py:8-29 (Dice_fn),
// train_files/trainchaos_proposed_30cases1labeled.py:274-292 (pseudo label), :303-321 (selection).
//
// With two classes everything is a function of d = z1 - z0 and s = softmax(z)[1]:
//   ce  = softplus(z_other - z_t) * wc[t]              dce/dd   = wc[t] * (s - t)
//   dice_n = 1 - (2 I + sm)/(S + T + sm)               ddice/ds = -(2 t D - (2 I + sm)) / D^2,  D = S+T+sm
//   mse = wm * ((1-s-q0)^2 + (s-q1)^2)                 dmse/ds  = 2 wm ((s-q1) - (1-s-q0))
//   ds/dd = s (1 - s);   dL/dz1 = +dL/dd,  dL/dz0 = -dL/dd.
// The sums are accumulated in fp64 (one pass, fixed-order two-stage reduction) so per-image losses
// match the fp32 reference to ~1e-7 and the small-loss ordering is reproducible.
#include "common.cuh"

namespace aide {

constexpr int kLossThreads = 256;
constexpr int kPixPerBlock = 4096;

__device__ __forceinline__ void softmax2(float z0, float z1, float& s0, float& s1) {
  float m = fmaxf(z0, z1);
  float e0 = expf(z0 - m), e1 = expf(z1 - m);
  float inv = 1.0f / (e0 + e1);
  s0 = e0 * inv;
  s1 = e1 * inv;
}
__device__ __forceinline__ float softplusf(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

__global__ void loss_sums_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets,
                                 const float* __restrict__ q, const float* __restrict__ wm, int HW, float wc0,
                                 float wc1, int ignore_index, float thr, double* __restrict__ scratch) {
  const int n = blockIdx.y;
  const float* z0p = logits + (size_t)n * 2 * HW;
  const float* z1p = z0p + HW;
  const int64_t* tp = targets + (size_t)n * HW;
  double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int p_end = min(HW, (int)(blockIdx.x + 1) * kPixPerBlock);
  for (int p = blockIdx.x * kPixPerBlock + threadIdx.x; p < p_end; p += kLossThreads) {
    float z0 = z0p[p], z1 = z1p[p];
    long long t = tp[p];
    float s0, s1;
    softmax2(z0, z1, s0, s1);
    float tf = (float)t;
    if (t != ignore_index) {
      float w = t == 1 ? wc1 : wc0;
      float ce = softplusf(t == 1 ? z0 - z1 : z1 - z0);
      a[0] += (double)(ce * w);
      a[1] += (double)w;
    }
    a[2] += (double)(s1 * tf);
    a[3] += (double)s1;
    a[4] += (double)tf;
    if (s1 >= thr) {
      a[5] += (double)tf;
      a[6] += 1.0;
    }
    if (q) {
      float q0 = q[(size_t)n * 2 * HW + p], q1 = q[(size_t)n * 2 * HW + HW + p];
      float w = wm ? wm[(size_t)n * HW + p] : 1.f;
      float e0 = s0 - q0, e1 = s1 - q1;
      a[7] += (double)(w * (e0 * e0)) + (double)(w * (e1 * e1));
    }
  }
  __shared__ double sm[kLossThreads / 32][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double v = warp_sum(a[k]);
    if (lane == 0) sm[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double t = 0.0;
    for (int w = 0; w < kLossThreads / 32; ++w) t += sm[w][threadIdx.x];
    scratch[((size_t)n * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = t;
  }
}

__global__ void loss_sums_finalize_kernel(const double* __restrict__ scratch, int blocks, double* __restrict__ sums) {
  int n = blockIdx.x, k = threadIdx.x;
  if (k >= 8) return;
  double t = 0.0;
  for (int b = 0; b < blocks; ++b) t += scratch[((size_t)n * blocks + b) * 8 + k];
  sums[(size_t)n * 8 + k] = t;
}

__global__ void loss_image_finalize_kernel(const double* __restrict__ sums, int N, double hw, float w_ce, float w_dice,
                                           float smooth, float* __restrict__ loss_img, float* __restrict__ dice_img,
                                           float* __restrict__ dice_fn_out) {
  // single block; N <= 1024
  __shared__ float dsum[1024];
  int n = threadIdx.x;
  float dfn = 0.f;
  if (n < N) {
    const double* s = sums + (size_t)n * 8;
    float ce = (float)(s[0] / hw);
    float I = (float)s[2], S = (float)s[3], T = (float)s[4];
    float dice = 1.0f - (2.0f * I + smooth) / (S + T + smooth);
    if (loss_img) loss_img[n] = ce * w_ce + dice * w_dice;
    if (dice_img) dice_img[n] = dice;
    // Dice_fn (metrics2d.py:14-28): thresholded prediction, empty-target rule
    float pi = (float)s[5], ps = (float)s[6];
    if (T == 0.f) dfn = (ps == 0.f) ? 1.f : 0.f;
    else dfn = (2.f * pi) / (ps + T);
  }
  dsum[threadIdx.x] = dfn;
  __syncthreads();
  if (threadIdx.x == 0 && dice_fn_out) {
    float t = 0.f;
    for (int i = 0; i < N; ++i) t += dsum[i];
    *dice_fn_out = t;
  }
}

__global__ void loss_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets,
                                const float* __restrict__ q, const float* __restrict__ wm,
                                const double* __restrict__ sums, const float* __restrict__ a_ce,
                                const float* __restrict__ a_dice, const float* __restrict__ a_mse, int HW, float wc0,
                                float wc1, int ignore_index, float smooth, float* __restrict__ dlogits) {
  const int n = blockIdx.y;
  const float ace = a_ce ? a_ce[n] : 0.f, adi = a_dice ? a_dice[n] : 0.f, ams = (a_mse && q) ? a_mse[n] : 0.f;
  const double* s = sums + (size_t)n * 8;
  const float I = (float)s[2], D = (float)s[3] + (float)s[4] + smooth;
  const float invD2 = 1.0f / (D * D), num = 2.0f * I + smooth;
  const size_t base = (size_t)n * 2 * HW;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    float z0 = logits[base + p], z1 = logits[base + HW + p];
    long long t = targets[(size_t)n * HW + p];
    float s0, s1;
    softmax2(z0, z1, s0, s1);
    float tf = (float)t;
    float gd = 0.f;  // dL/dd
    if (ace != 0.f && t != ignore_index) gd += ace * (t == 1 ? wc1 : wc0) * (s1 - tf);
    float gs = 0.f;  // dL/ds1 (through the softmax)
    if (adi != 0.f) gs += adi * (-(2.0f * tf * D - num) * invD2);
    if (ams != 0.f) {
      float q0 = q[base + p], q1 = q[base + HW + p];
      float w = wm ? wm[(size_t)n * HW + p] : 1.f;
      gs += ams * 2.0f * w * ((s1 - q1) - (s0 - q0));
    }
    gd += gs * (s1 * s0);
    dlogits[base + p] = -gd;
    dlogits[base + HW + p] = gd;
  }
}

struct AugPtrs {
  const float* p[8];
};
__global__ void pseudo_label_kernel(AugPtrs aug, int n_aug, int HW, size_t total, float expo, float* __restrict__ q,
                                    float* __restrict__ wm) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t n = i / HW, p = i % HW;
    size_t base = n * 2 * HW + p;
    float a0 = 0.f, a1 = 0.f;
    for (int k = 0; k < n_aug; ++k) {
      float s0, s1;
      softmax2(aug.p[k][base], aug.p[k][base + HW], s0, s1);
      a0 += s0;
      a1 += s1;
    }
    float inv = 1.0f / (float)n_aug;
    a0 = a0 / (float)n_aug;
    a1 = a1 / (float)n_aug;
    (void)inv;
    float m0 = expo == 1.0f ? a0 : powf(a0, expo), m1 = expo == 1.0f ? a1 : powf(a1, expo);
    float sum = m0 + m1;
    float q0 = m0 / sum, q1 = m1 / sum;
    q[base] = q0;
    q[base + HW] = q1;
    wm[n * HW + p] = 1.0f - 4.0f * q0 * q1;
  }
}

// single block, n_total <= 1024 threads.  `pre_other` holds the OTHER net's per-image losses of the whole (global) batch of
// n_total images; this rank's N images are entries [first, first + N).  NaN losses sort last (like torch.sort) and ties
// break on the global index, so the ranks are always a permutation.
__device__ __forceinline__ bool sorts_before(float o, int j, float mine, int i) {
  const bool on = o != o, mn = mine != mine;
  if (on || mn) return on == mn ? j < i : mn;
  return o < mine || (o == mine && j < i);
}
__global__ void coteach_select_kernel(const float* __restrict__ pre_other, int n_total, int first,
                                      const float* __restrict__ loss_img, const double* __restrict__ sums, int N,
                                      double hw, int n_clean, float rate, const float* __restrict__ rate_dev, float seg_w,
                                      float cor_w, float w_ce, float w_dice, int64_t* __restrict__ idx,
                                      float* __restrict__ a_ce, float* __restrict__ a_dice, float* __restrict__ a_mse,
                                      float* __restrict__ loss_out) {
  __shared__ float v[1024];
  __shared__ int order[1024];
  const int i = threadIdx.x;
  if (rate_dev) rate = *rate_dev;            // read from device memory: the warm-up schedule moves under a captured graph
  if (i < n_total) v[i] = pre_other[i];
  __syncthreads();
  int rank = 0;
  if (i < n_total) {
    const float mine = v[i];
    for (int j = 0; j < n_total; ++j) rank += sorts_before(v[j], j, mine, i) ? 1 : 0;   // stable ascending
    order[rank] = i;
  }
  __syncthreads();
  const int n_rest = n_total - n_clean;
  if (i < n_total) {
    idx[i] = order[i];
    const int l = i - first;                 // local image index
    if (l >= 0 && l < N) {
      const bool clean = rank < n_clean;
      // loss = seg_w * (mean_clean L + (1-rate) mean_rest L) + cor_w * rate * sum_rest(mse) / (n_rest*2*HW)
      const float cseg = clean ? seg_w / (float)n_clean : (n_rest > 0 ? seg_w * (1.0f - rate) / (float)n_rest : 0.f);
      a_ce[l] = cseg * w_ce / (float)hw;
      a_dice[l] = cseg * w_dice;
      a_mse[l] = (!clean && n_rest > 0) ? cor_w * rate / (float)((double)n_rest * 2.0 * hw) : 0.f;
    }
  }
  __syncthreads();
  if (i == 0 && loss_out) {
    // fixed-order scalar assembly following the reference expression order; with n_total > N this is the rank's
    // share of the global loss (the shares of all ranks add up to it)
    float seg1 = 0.f, seg2 = 0.f;
    double mse = 0.0;
    for (int r = 0; r < n_total; ++r) {
      const int k = order[r] - first;
      if (k < 0 || k >= N) continue;
      if (r < n_clean) seg1 += loss_img[k];
      else {
        seg2 += loss_img[k];
        mse += sums[(size_t)k * 8 + 7];
      }
    }
    seg1 = seg1 / (float)n_clean;
    float cor = 0.f;
    if (n_rest > 0) {
      seg2 = seg2 / (float)n_rest;
      cor = (float)(mse / ((double)n_rest * 2.0 * hw));
    }
    *loss_out = seg_w * (seg1 + (1.0f - rate) * seg2) + cor_w * rate * cor;
  }
}

}  // namespace aide

using namespace aide;

extern "C" int aide_loss_blocks(int H, int W) { return ceil_div((long long)H * W, kPixPerBlock); }

extern "C" int aide_loss_sums(const float* logits, const int64_t* targets, const float* q, const float* wm, int N,
                              int H, int W, float wc0, float wc1, int ignore_index, float threshold, double* sums,
                              double* scratch, void* stream) {
  AIDE_REQUIRE(logits && targets && sums && scratch && N > 0 && H > 0 && W > 0, "loss_sums: bad arguments");
  int blocks = aide_loss_blocks(H, W);
  loss_sums_kernel<<<dim3(blocks, N), kLossThreads, 0, as_stream(stream)>>>(logits, targets, q, wm, H * W, wc0, wc1,
                                                                            ignore_index, threshold, scratch);
  AIDE_CHECK_LAUNCH();
  loss_sums_finalize_kernel<<<N, 32, 0, as_stream(stream)>>>(scratch, blocks, sums);
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_loss_image_finalize(const double* sums, int N, int H, int W, float w_ce, float w_dice,
                                        float smooth, float* loss_img, float* dice_img, float* dice_fn_out,
                                        void* stream) {
  AIDE_REQUIRE(sums && N > 0 && N <= 1024, "loss_image_finalize: bad arguments (N <= 1024)");
  int threads = ((N + 31) / 32) * 32;
  loss_image_finalize_kernel<<<1, threads, 0, as_stream(stream)>>>(sums, N, (double)H * W, w_ce, w_dice, smooth,
                                                                   loss_img, dice_img, dice_fn_out);
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_loss_bwd(const float* logits, const int64_t* targets, const float* q, const float* wm,
                             const double* sums, const float* a_ce, const float* a_dice, const float* a_mse, int N,
                             int H, int W, float wc0, float wc1, int ignore_index, float smooth, float* dlogits,
                             void* stream) {
  AIDE_REQUIRE(logits && targets && sums && dlogits && N > 0, "loss_bwd: bad arguments");
  int HW = H * W;
  int bx = ceil_div(HW, 256 * 4);
  if (bx < 1) bx = 1;
  loss_bwd_kernel<<<dim3(bx, N), 256, 0, as_stream(stream)>>>(logits, targets, q, wm, sums, a_ce, a_dice, a_mse, HW,
                                                              wc0, wc1, ignore_index, smooth, dlogits);
  AIDE_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------- hard mask: argmax(softmax(logits, 1), 1) as uint8
// (trainchaos_proposed_30cases1labeled.py:407-409, evalchaos_comparison_1cases.py:208-214).  The softmax is evaluated
// like the reference does (exp(z - max) / sum, fp32) and the FIRST maximum wins, as in torch.argmax.
namespace aide {
__global__ void argmax_mask_kernel(const float* __restrict__ logits, uint8_t* __restrict__ mask, int N, int K, size_t HW) {
  const size_t total = (size_t)N * HW;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / HW, hw = i - n * HW;
    const float* z = logits + n * K * HW + hw;
    float m = z[0];
    for (int k = 1; k < K; ++k) m = fmaxf(m, z[k * HW]);
    float sum = 0.f;
    for (int k = 0; k < K; ++k) sum += expf(z[k * HW] - m);
    int best = 0;
    float pb = expf(z[0] - m) / sum;
    for (int k = 1; k < K; ++k) {
      const float p = expf(z[k * HW] - m) / sum;
      if (p > pb) {
        pb = p;
        best = k;
      }
    }
    mask[i] = (uint8_t)best;
  }
}
}  // namespace aide

extern "C" int aide_argmax_mask(const float* logits, uint8_t* mask, int N, int K, int H, int W, void* stream) {
  AIDE_REQUIRE(logits && mask && N > 0 && K >= 1 && K <= 255 && H > 0 && W > 0, "argmax_mask: bad arguments");
  const size_t total = (size_t)N * H * W;
  long long blocks = (long long)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  aide::argmax_mask_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(logits, mask, N, K, (size_t)H * W);
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_pseudo_label(const float* const* aug_logits, int n_aug, int N, int H, int W, float expo, float* q,
                                 float* wm, void* stream) {
  AIDE_REQUIRE(aug_logits && n_aug >= 1 && n_aug <= 8 && q && wm, "pseudo_label: 1..8 augmented logit tensors");
  AugPtrs a{};
  for (int i = 0; i < n_aug; ++i) {
    AIDE_REQUIRE(aug_logits[i], "pseudo_label: null logits pointer");
    a.p[i] = aug_logits[i];
  }
  size_t total = (size_t)N * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  pseudo_label_kernel<<<blocks, 256, 0, as_stream(stream)>>>(a, n_aug, H * W, total, expo, q, wm);
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_coteach_select_ex(const float* pre_other_all, int n_total, int first, const float* loss_img,
                                      const double* sums, int N, int H, int W, int n_clean, float rate,
                                      const float* rate_dev, float seg_w, float cor_w, float w_ce, float w_dice,
                                      int64_t* idx, float* a_ce, float* a_dice, float* a_mse, float* loss_out,
                                      void* stream) {
  AIDE_REQUIRE(pre_other_all && loss_img && sums && idx && a_ce && a_dice && a_mse, "coteach_select: null argument");
  AIDE_REQUIRE(N >= 1 && n_total >= N && n_total <= 1024 && first >= 0 && first + N <= n_total && n_clean >= 1 &&
                   n_clean <= n_total,
               "coteach_select: need 1 <= n_clean <= n_total <= 1024 and [first, first+N) inside the global batch");
  int threads = ((n_total + 31) / 32) * 32;
  coteach_select_kernel<<<1, threads, 0, as_stream(stream)>>>(pre_other_all, n_total, first, loss_img, sums, N,
                                                              (double)H * W, n_clean, rate, rate_dev, seg_w, cor_w, w_ce,
                                                              w_dice, idx, a_ce, a_dice, a_mse, loss_out);
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_coteach_select(const float* pre_other, const float* loss_img, const double* sums, int N, int H,
                                   int W, int n_clean, float rate, float seg_w, float cor_w, float w_ce, float w_dice,
                                   int64_t* idx, float* a_ce, float* a_dice, float* a_mse, float* loss_out,
                                   void* stream) {
  return aide_coteach_select_ex(pre_other, N, 0, loss_img, sums, N, H, W, n_clean, rate, nullptr, seg_w, cor_w, w_ce,
                                w_dice, idx, a_ce, a_dice, a_mse, loss_out, stream);
}

// ---------------------------------------------------------------- per-pixel maps of the drop-in loss classes
// CrossEntropyLoss2d(reduction='none') (utils/loss2d.py:5-13), the pixel branch of Coteachingloss_dropimagedroppixel
// (utils/coteach_loss.py:223-254: KLbidirection :85-92 + CE), Coteachingloss_dropregionce (:163-196) and
// MulticlassMSELoss(reduction='none') (loss2d.py:109-117) return MAPS the calling script reduces itself; these kernels
// produce the maps and their gradients in one pass each (two classes, as everywhere in the reference).
namespace aide {

// out[p] = [mode & 1] wc[t] * CE(z, t) (0 where t == ignore_index) + [mode & 2] KL(p1||p2) + KL(p2||p1)
// two classes: KL(p||q) + KL(q||p) = (s1 - s2) * (d1 - d2) with d = z1 - z0, s = sigmoid(d)
__global__ void pixel_loss_fwd_kernel(const float* __restrict__ l1, const float* __restrict__ l2,
                                      const int64_t* __restrict__ tg, int HW, size_t total, float wc0, float wc1,
                                      int ignore_index, int mode, float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / HW, p = i - n * HW, b = n * 2 * HW + p;
    const float z0 = l1[b], z1 = l1[b + HW];
    const long long t = tg[i];
    float v = 0.f;
    if ((mode & 1) && t != ignore_index) v += (t == 1 ? wc1 : wc0) * softplusf(t == 1 ? z0 - z1 : z1 - z0);
    if (mode & 2) {
      const float y0 = l2[b], y1 = l2[b + HW];
      float a0, a1, b0, b1;
      softmax2(z0, z1, a0, a1);
      softmax2(y0, y1, b0, b1);
      v += (a1 - b1) * ((z1 - z0) - (y1 - y0));
    }
    out[i] = v;
  }
}
__global__ void pixel_loss_bwd_kernel(const float* __restrict__ l1, const float* __restrict__ l2,
                                      const int64_t* __restrict__ tg, const float* __restrict__ gout, int HW, size_t total,
                                      float wc0, float wc1, int ignore_index, int mode, float* __restrict__ d1,
                                      float* __restrict__ d2) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / HW, p = i - n * HW, b = n * 2 * HW + p;
    const float z0 = l1[b], z1 = l1[b + HW], g = gout[i];
    const long long t = tg[i];
    float a0, a1;
    softmax2(z0, z1, a0, a1);
    float gd1 = 0.f, gd2 = 0.f;                 // d/d(d1), d/d(d2)
    if ((mode & 1) && t != ignore_index) gd1 += (t == 1 ? wc1 : wc0) * (a1 - (float)t);
    if (mode & 2) {
      const float y0 = l2[b], y1 = l2[b + HW];
      float b0, b1;
      softmax2(y0, y1, b0, b1);
      const float dd = (z1 - z0) - (y1 - y0), ds = a1 - b1;
      gd1 += a1 * a0 * dd + ds;
      gd2 -= b1 * b0 * dd + ds;
    }
    d1[b] = -gd1 * g;
    d1[b + HW] = gd1 * g;
    if (d2) {
      d2[b] = -gd2 * g;
      d2[b + HW] = gd2 * g;
    }
  }
}

// out[n][k][p] = (softmax(z)[k] - target[n][k][p])^2   (two classes)
__global__ void softmax_mse_fwd_kernel(const float* __restrict__ z, const float* __restrict__ tg, int HW, size_t npix,
                                       float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / HW, p = i - n * HW, b = n * 2 * HW + p;
    float s0, s1;
    softmax2(z[b], z[b + HW], s0, s1);
    const float e0 = s0 - tg[b], e1 = s1 - tg[b + HW];
    out[b] = e0 * e0;
    out[b + HW] = e1 * e1;
  }
}
__global__ void softmax_mse_bwd_kernel(const float* __restrict__ z, const float* __restrict__ tg,
                                       const float* __restrict__ gout, int HW, size_t npix, float* __restrict__ dz) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / HW, p = i - n * HW, b = n * 2 * HW + p;
    float s0, s1;
    softmax2(z[b], z[b + HW], s0, s1);
    const float g0 = 2.f * (s0 - tg[b]) * gout[b], g1 = 2.f * (s1 - tg[b + HW]) * gout[b + HW];
    const float gd = s1 * s0 * (g1 - g0);       // through softmax: d s1/dd = s1 s0, d s0/dd = -s1 s0
    dz[b] = -gd;
    dz[b + HW] = gd;
  }
}

// max_pool2d(kernel = stride = (kh, kw), padding 0, ceil_mode) on NCHW planes, forward (+ argmax) and backward
__global__ void maxpool_nchw_fwd_kernel(const float* __restrict__ x, int H, int W, int kh, int kw, int OH, int OW,
                                        size_t total, float* __restrict__ y, int* __restrict__ arg) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ow = (int)(i % OW), oh = (int)((i / OW) % OH);
    const size_t plane = i / ((size_t)OW * OH);
    const float* xp = x + plane * (size_t)H * W;
    float m = -INFINITY;
    int best = -1;
    for (int yy = oh * kh; yy < min(H, oh * kh + kh); ++yy)
      for (int xx = ow * kw; xx < min(W, ow * kw + kw); ++xx) {
        const float v = xp[(size_t)yy * W + xx];
        if (best < 0 || v > m || v != v) {            // first maximum wins, NaN propagates (ATen max_pool2d)
          m = v;
          best = yy * W + xx;
        }
      }
    y[i] = m;
    if (arg) arg[i] = best;
  }
}
__global__ void maxpool_nchw_bwd_kernel(const float* __restrict__ gy, const int* __restrict__ arg, int HWin, int OHW,
                                        size_t total, float* __restrict__ gx /* zero-initialised */) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t plane = i / OHW;
    gx[plane * (size_t)HWin + arg[i]] = gy[i];        // windows do not overlap (stride = kernel): no atomics needed
  }
}

static int grid1d(size_t total) {
  long long b = (long long)((total + 255) / 256);
  if (b > kNumSMs * 16) b = kNumSMs * 16;
  return b < 1 ? 1 : (int)b;
}

}  // namespace aide

extern "C" int aide_pixel_loss_fwd(const float* logits1, const float* logits2, const int64_t* targets, int N, int H, int W,
                                   float wc0, float wc1, int ignore_index, int mode, float* out, void* stream) {
  AIDE_REQUIRE(logits1 && targets && out && N > 0 && H > 0 && W > 0 && (mode & 3) && (!(mode & 2) || logits2),
               "pixel_loss_fwd: bad arguments (mode bit 0 = CE, bit 1 = bidirectional KL, which needs logits2)");
  const size_t total = (size_t)N * H * W;
  aide::pixel_loss_fwd_kernel<<<aide::grid1d(total), 256, 0, as_stream(stream)>>>(logits1, logits2, targets, H * W, total, wc0,
                                                                                  wc1, ignore_index, mode, out);
  AIDE_CHECK_LAUNCH();
  return 0;
}
extern "C" int aide_pixel_loss_bwd(const float* logits1, const float* logits2, const int64_t* targets, const float* gout,
                                   int N, int H, int W, float wc0, float wc1, int ignore_index, int mode, float* dlogits1,
                                   float* dlogits2, void* stream) {
  AIDE_REQUIRE(logits1 && targets && gout && dlogits1 && (mode & 3) && (!(mode & 2) || logits2), "pixel_loss_bwd: bad arguments");
  const size_t total = (size_t)N * H * W;
  aide::pixel_loss_bwd_kernel<<<aide::grid1d(total), 256, 0, as_stream(stream)>>>(logits1, logits2, targets, gout, H * W, total,
                                                                                  wc0, wc1, ignore_index, mode, dlogits1,
                                                                                  dlogits2);
  AIDE_CHECK_LAUNCH();
  return 0;
}
extern "C" int aide_softmax_mse_fwd(const float* logits, const float* target, int N, int H, int W, float* out, void* stream) {
  AIDE_REQUIRE(logits && target && out && N > 0, "softmax_mse_fwd: bad arguments");
  const size_t npix = (size_t)N * H * W;
  aide::softmax_mse_fwd_kernel<<<aide::grid1d(npix), 256, 0, as_stream(stream)>>>(logits, target, H * W, npix, out);
  AIDE_CHECK_LAUNCH();
  return 0;
}
extern "C" int aide_softmax_mse_bwd(const float* logits, const float* target, const float* gout, int N, int H, int W,
                                    float* dlogits, void* stream) {
  AIDE_REQUIRE(logits && target && gout && dlogits && N > 0, "softmax_mse_bwd: bad arguments");
  const size_t npix = (size_t)N * H * W;
  aide::softmax_mse_bwd_kernel<<<aide::grid1d(npix), 256, 0, as_stream(stream)>>>(logits, target, gout, H * W, npix, dlogits);
  AIDE_CHECK_LAUNCH();
  return 0;
}
extern "C" int aide_maxpool_nchw_fwd(const float* x, int planes, int H, int W, int kh, int kw, float* y, int* argmax,
                                     void* stream) {
  AIDE_REQUIRE(x && y && planes > 0 && H > 0 && W > 0 && kh > 0 && kw > 0, "maxpool_nchw_fwd: bad arguments");
  const int OH = (H + kh - 1) / kh, OW = (W + kw - 1) / kw;
  const size_t total = (size_t)planes * OH * OW;
  aide::maxpool_nchw_fwd_kernel<<<aide::grid1d(total), 256, 0, as_stream(stream)>>>(x, H, W, kh, kw, OH, OW, total, y, argmax);
  AIDE_CHECK_LAUNCH();
  return 0;
}
extern "C" int aide_maxpool_nchw_bwd(const float* gy, const int* argmax, int planes, int H, int W, int kh, int kw, float* gx,
                                     void* stream) {
  AIDE_REQUIRE(gy && argmax && gx && planes > 0, "maxpool_nchw_bwd: bad arguments");
  const int OH = (H + kh - 1) / kh, OW = (W + kw - 1) / kw;
  const size_t total = (size_t)planes * OH * OW;
  AIDE_CUDA(cudaMemsetAsync(gx, 0, (size_t)planes * H * W * sizeof(float), as_stream(stream)));
  aide::maxpool_nchw_bwd_kernel<<<aide::grid1d(total), 256, 0, as_stream(stream)>>>(gy, argmax, H * W, OH * OW, total, gx);
  AIDE_CHECK_LAUNCH();
  return 0;
}
