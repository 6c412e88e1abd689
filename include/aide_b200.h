/* aide_b200.h -- C ABI of libaide_b200.so: the B200 (sm_100a) engine for the AIDE hot path.
 *
 * The reference (lich0031/AIDE) is pure Python/PyTorch and has no FFI of its own; each entry point
 * below cites the reference code (file:line, relative to the reference tree) whose work it replaces.
 * The Python host (aide_b200/) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; aide_last_error() returns a
 *     thread-local description.  Nothing here allocates device memory: all buffers (including
 *     workspaces, whose sizes are queried with the *_workspace_bytes / *_rows helpers) are owned
 *     by the caller and passed as raw device pointers.  `stream` is a cudaStream_t.
 *   - activations are NHWC ("pixel-major"): element (n,h,w,c) of a buffer with `ctot` channels
 *     lives at ((n*H+h)*W+w)*ctot + c.  A *view* (p0,p1,ctot,coff) addresses channels
 *     [coff, coff+C) of such a buffer -- this is how torch.cat (fuseunet.py:49-81,
 *     netblocks.py:145) becomes zero-copy: producers write straight into their slice.
 *   - `fmt` is the operand format of activation/weight planes:
 *       AIDE_FMT_F32    one fp32 plane (CUDA-core exact path)
 *       AIDE_FMT_TF32X2 two fp32 planes hi = rn_tf32(x), lo = x - hi ("parity mode": 3 tcgen05
 *                       kind::tf32 MMAs hi*hi + hi*lo + lo*hi reproduce fp32 products)
 *       AIDE_FMT_BF16   one bf16 plane ("fast mode": single kind::f16 MMA)
 *       AIDE_FMT_F16X2  two fp16 planes hi = fp16(x*2^s), lo = fp16(x*2^s - hi), s a power of two (activations 2^8,
 *                       weights 2^12, gradients: chosen per tensor ON THE DEVICE by aide_bn_relu_bwd_apply): the
 *                       products hi*hi + hi*lo + lo*hi give the same 22-bit mantissas as TF32X2 at twice the
 *                       tensor-core rate and half the operand bytes.  This is the default "parity" format for
 *                       forward, dgrad and wgrad.
 *     Raw convolution outputs (z), gradients w.r.t. activations and all reductions are fp32.
 */
#ifndef AIDE_B200_H_
#define AIDE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { AIDE_FMT_F32 = 0, AIDE_FMT_TF32X2 = 1, AIDE_FMT_BF16 = 2, AIDE_FMT_F16X2 = 3 };

const char* aide_last_error(void);
int aide_version(void);
/* number of CUDA kernels this library has launched in this process (monotonic; bench.py reports the
 * per-step difference as "gpu_launches"). */
unsigned long long aide_launch_count(void);
/* AIDE_FMT_F16X2 stores tensors pre-scaled by a power of two in two fp16 planes; a value outside the fp16 range after
 * scaling (|activation| >= 255.87, |weight| >= 16, a gradient beyond its dynamic scale's head-room) is CLIPPED.  Every
 * conversion kernel sets a sticky device flag when that happens: returns 1 if any tensor saturated since the last
 * reset (synchronising read of the current device; reset != 0 clears the flag). */
int aide_f16_saturated(int reset);
/* 1 if the tcgen05/TMA path can run (driver entry point for cuTensorMapEncodeTiled found). */
int aide_has_tma(void);

/* ---- layout / weight preparation -------------------------------------------------------------- */
/* NCHW fp32 [N,C,H,W] (module boundary, fuseunet.py:43) -> NHWC view in operand format. */
int aide_nchw_to_nhwc(int fmt, const float* src, void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff,
                      int N, int C, int H, int W, void* stream);
/* NHWC view (operand format) -> NCHW fp32. */
int aide_nhwc_to_nchw(int fmt, const void* src_p0, const void* src_p1, int src_ctot, int src_coff,
                      float* dst, int N, int C, int H, int W, void* stream);
/* nn.Conv2d weight OIHW fp32 [Cout,Cin,3,3] (netblocks.py:24,26,17) ->
 *   fwd   planes [Cout][9][Cin]   (tap = ky*3+kx)            used by aide_conv3x3_fwd
 *   dgrad planes [Cin][9][Cout]   (tap flipped: 8 - tap)     used by aide_conv3x3_fwd as dgrad
 * either destination may be NULL. */
int aide_weight_prep(int fmt, const float* w_oihw, int cout, int cin,
                     void* fwd_p0, void* fwd_p1, void* dgrad_p0, void* dgrad_p1, void* stream);
/* The same for n_layers tensor-core layers (cout % 32 == 0, cin % 32 == 0) in ONE launch: host arrays of per-layer
 * pointers / sizes (every conv of fuseunet.py:12-39 after an optimiser step). */
int aide_weight_prep_batch(int fmt, int n_layers, const float* const* w_oihw, const int* cout, const int* cin,
                           void* const* fwd_p0, void* const* fwd_p1, void* const* dgrad_p0, void* const* dgrad_p1,
                           void* stream);

/* ---- conv3x3, stride 1, pad 1 (netblocks.py:24,26,17 -> ATen conv2d / cuDNN) -------------------- */
/* z[n,h,w,co] = bias[co] + sum_{tap,ci} x[n,h+dy,w+dx,ci] * w[co][tap][ci].
 * Also emits per-tile BatchNorm partial statistics (sum z, sum z^2 per channel) when stat_partial
 * != NULL: layout [rows][2][cout] fp32 with rows = aide_conv3x3_stat_rows(...).
 * F32 -> CUDA-core kernels (a dedicated HBM-bound kernel for the networks' first layer, cin == 3 and cout in
 * {32, 64}; a generic tiled kernel otherwise); TF32X2 / BF16 / F16X2 -> tcgen05 implicit GEMM fed by TMA (needs
 * cin % 32 == 0 and cout % 32 == 0; other shapes -> error): the halo-reuse kernel (conv_halo_tc.cu) for maps of at
 * least 8x8 pixels, the first-generation kernel (conv_fwd_tc.cu) below that. */
int aide_conv3x3_stat_rows(int fmt, int cin, int cout, int N, int H, int W);
int aide_conv3x3_fwd(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int cin,
                     const void* w_p0, const void* w_p1, const float* bias,
                     float* z, int z_ctot, int z_coff, int cout, int N, int H, int W,
                     float* stat_partial, void* stream);
/* Inference form of a whole conv3x3 -> BatchNorm2d(eval) -> ReLU (-> MaxPool2d(2,2)) unit (netblocks.py:30-33 under
 * module.eval(); evalchaos_comparison_1cases.py:203-214) as ONE tcgen05 launch: the epilogue applies
 * y = relu(scale*(acc+bias)+shift) with scale_shift [2][cout] from aide_bn_eval_scale_shift_batch and writes the operand
 * planes straight into the consumer's channel slice (dst) and the 2x2 max-pool into up to two half-resolution slices --
 * no fp32 z round trip, no finalize / apply launches; bit-identical to aide_conv3x3_fwd + aide_bn_relu_apply.
 * aide_conv3x3_bn_relu_ok: 1 when the (format, shape) is covered (tensor-core formats, maps of at least 8x8). */
int aide_conv3x3_bn_relu_ok(int fmt, int cin, int cout, int N, int H, int W);
int aide_conv3x3_bn_relu_fwd(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int cin,
                             const void* w_p0, const void* w_p1, const float* bias, const float* scale_shift, int cout,
                             int N, int H, int W, void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff,
                             void* poolA_p0, void* poolA_p1, int poolA_ctot, int poolA_coff,
                             void* poolB_p0, void* poolB_p1, int poolB_ctot, int poolB_coff, void* stream);
/* Tiling the tcgen05 forward/dgrad kernel picks for a layer (diagnostics for bench.py / tools): out[9] =
 * {cout tile, pixel tiles per CTA iteration, accumulators per tile, TMEM buffers, smem row bytes, halo stages,
  *  weight stages, dynamic smem bytes, bit 0: hi/lo weight planes stacked along N, bit 1: weights resident in shared memory}.  Returns non-zero when the layer runs on the first-generation kernel. */
int aide_conv3x3_plan_info(int fmt, int cin, int cout, int N, int H, int W, int* out);
/* dX[n,h,w,ci] = sum_{tap,co} dz[n,h+dy,w+dx,co] * w[co][ci][flipped tap]  (ATen conv backward-data): the forward
 * kernel on dgrad-prepared weights (aide_weight_prep).  dz is a plain [N,H,W,cout] operand-format buffer, dx an fp32
 * NHWC view.  dz_inv_scale: device pointer to 1/s where the dz planes hold dz*s (AIDE_FMT_F16X2 gradients are stored
 * with a power-of-two scale chosen on the device by aide_bn_relu_bwd_apply); NULL for the other formats. */
int aide_conv3x3_dgrad(int fmt, const void* dz_p0, const void* dz_p1, int cout, const void* w_p0, const void* w_p1,
                       const float* dz_inv_scale, float* dx, int dx_ctot, int dx_coff, int cin, int N, int H, int W,
                       void* stream);
/* dW[co,ci,ky,kx] = sum_{n,h,w} dz[n,h,w,co] * x[n,h+ky-1,w+kx-1,ci]  (ATen conv backward-filter).
 * dz is a plain [N,H,W,cout] operand-format buffer (dz_inv_scale as above).  Result written (not accumulated) as
 * OIHW fp32. */
size_t aide_conv3x3_wgrad_workspace_bytes(int fmt, int cin, int cout, int N, int H, int W);
int aide_conv3x3_wgrad(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int cin,
                       const void* dz_p0, const void* dz_p1, const float* dz_inv_scale, int cout, int N, int H, int W,
                       void* workspace, size_t workspace_bytes, float* dw_oihw, void* stream);
/* aide_conv3x3_wgrad plus a column sum colsum_dst[j] = sum_r colsum_src[r * colsum_cols + j] folded into the weight gradient's
 * split-K reduction launch -- the conv-bias gradient of the same unit from the partial rows aide_bn_relu_bwd_apply leaves in
 * partial2 (pass dbias_conv = NULL there): one launch less per unit (nn.Conv2d bias gradient, netblocks.py:24,26). */
int aide_conv3x3_wgrad_ex(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int cin,
                          const void* dz_p0, const void* dz_p1, const float* dz_inv_scale, int cout, int N, int H, int W,
                          void* workspace, size_t workspace_bytes, float* dw_oihw, const float* colsum_src,
                          int colsum_rows, int colsum_cols, float* colsum_dst, void* stream);

/* ---- BatchNorm2d (+ReLU, +MaxPool2d(2,2), +concat slot) (netblocks.py:25,27,28; fuseunet.py:13-31) */
/* Reduce the conv's partial statistics in fixed order (fp64), then
 *   training: mean/biased var -> scale,shift ; running_mean/var updated with momentum, unbiased var
 *   eval    : scale,shift from running stats (stat_partial ignored)
 * scale_shift: [2][C] fp32 (y = scale*z + shift); mean_rstd: [2][C] fp32 (saved for backward). */
int aide_bn_finalize(float* stat_partial /* folded in place: consumed */, int rows, int C, double count,
                     const float* gamma, const float* beta, float* running_mean, float* running_var,
                     float momentum, float eps, int training,
                     float* scale_shift, float* mean_rstd, void* stream);
/* The same for `sgroups` statistics groups at once (a stacked-batch forward whose views keep their own BatchNorm
 * statistics -- the 4 augmented forwards of trainchaos_proposed_30cases1labeled.py:265-269 run as one batch):
 * stat_partial holds sgroups consecutive blocks of `rows` rows, scale_shift / mean_rstd are [sgroups][2][C], and the
 * running statistics are updated once per group, in group order, exactly like sgroups separate forward calls. */
int aide_bn_finalize_grouped(float* stat_partial, int rows, int sgroups, int C, double count, const float* gamma,
                             const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                             int training, float* scale_shift, float* mean_rstd, unsigned int* tickets, void* stream);
/* tickets (nullable): aide_bn_ticket_slots(C) zero-initialised counters.  When given, the chunk fold and the finalize
 * run as ONE launch: the last fold block of a channel group to arrive (ticket) finalises it, reading the folded
 * partials in fixed order -- bit-identical to the two-launch path.  The counters are left at zero. */
int aide_bn_ticket_slots(int C);
/* Eval mode (module.eval(): running statistics): scale / shift [2][C] of n_layers BatchNorm layers in ONE launch (host
 * arrays of per-layer device pointers) -- the per-epoch evaluation / pseudo-label rewrite loops
 * (trainchaos_proposed_30cases1labeled.py:373-496) run thousands of single-slice eval forwards. */
int aide_bn_eval_scale_shift_batch(int n_layers, const float* const* gamma, const float* const* beta,
                                   const float* const* running_mean, const float* const* running_var, const int* C,
                                   float* const* scale_shift, float eps, void* stream);
/* y = relu(scale*z+shift) written to `dst` (full resolution) and, when given, the 2x2 max-pooled y to
 * up to two half-resolution views (the fused-encoder concat and the modal-2 branch, fuseunet.py:51-56). */
int aide_bn_relu_apply(int fmt, const float* z, int N, int H, int W, int C, const float* scale_shift,
                       void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff,
                       void* poolA_p0, void* poolA_p1, int poolA_ctot, int poolA_coff,
                       void* poolB_p0, void* poolB_p1, int poolB_ctot, int poolB_coff, void* stream);
/* Stacked-batch form: image n is normalised with scale_shift[n / imgs_per_group] ([groups][2][C]). */
int aide_bn_relu_apply_grouped(int fmt, const float* z, int N, int imgs_per_group, int H, int W, int C,
                               const float* scale_shift, void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff,
                               void* poolA_p0, void* poolA_p1, int poolA_ctot, int poolA_coff,
                               void* poolB_p0, void* poolB_p1, int poolB_ctot, int poolB_coff, void* stream);
/* backward, stage 1: g = relu'(y) * (sum of upstream gradients).  Upstream sources (all fp32 NHWC):
 * up to 3 same-resolution slices and up to 3 half-resolution slices routed through the max-pool
 * arg-max (first maximum in window scan order wins, like ATen max_pool2d_with_indices).  y is
 * recomputed from z.  Writes g [N,H,W,C] and per-block partial sums of (g, g*xhat): [rows][2][C],
 * rows = aide_bn_bwd_rows(N,H,W,C).  gmax (nullable): device word that receives max |g| by atomicMax -- it must be
 * ZERO on entry; aide_bn_relu_bwd_apply consumes it and leaves it zero for the next unit. */
int aide_bn_bwd_rows(int N, int H, int W, int C);
int aide_bn_relu_bwd_reduce(const float* z, const float* scale_shift, const float* mean_rstd,
                            int N, int H, int W, int C,
                            const float* const* direct_ptr, const int* direct_ctot, const int* direct_coff, int n_direct,
                            const float* const* pool_ptr, const int* pool_ctot, const int* pool_coff, int n_pool,
                            float* g, float* partial, float* gmax /* nullable: device scalar <- max |g| */, void* stream);
/* backward, stage 2: sums = fixed-order reduction of `partial`; dgamma = sum g*xhat, dbeta = sum g,
 * dz = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)) stored in operand format [N,H,W,C];
 * dbias_conv = sum dz (mathematically 0 under train-mode BN; kept for fidelity).
 * partial2: scratch [rows][C] fp32 (rows as above).
 * AIDE_FMT_F16X2: the dz planes hold dz*s, s a power of two derived on the device from `gmax` (the reduce stage's
 * max |g|) and the channel statistics; dz_scale[2] receives {s, 1/s} for aide_conv3x3_dgrad / _wgrad.  Both may be
 * NULL for the other formats.  tickets (nullable): 16 zero-initialised words, left zero; when given, the column
 * reduction + scale derivation run as one launch and dbias_conv is folded by the last block of the apply kernel
 * (fixed order, same values) instead of two extra launches. */
int aide_bn_relu_bwd_apply(int fmt, const float* g, const float* z, const float* mean_rstd, const float* gamma,
                           const float* partial, int rows, int N, int H, int W, int C,
                           void* dz_p0, void* dz_p1, float* dgamma, float* dbeta, float* dbias_conv,
                           float* partial2, float* gmax, float* dz_scale, unsigned int* tickets, void* stream);

/* ---- nn.Upsample(scale_factor=2, bilinear, align_corners=True) (netblocks.py:16) ---------------- */
int aide_upsample2x_fwd(int fmt, const void* src_p0, const void* src_p1, int src_ctot, int src_coff,
                        void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff,
                        int N, int h, int w, int C, void* stream);
/* d_lo[N,h,w,C] (fp32) = transpose of the interpolation applied to d_hi view [N,2h,2w,*] (fp32). */
int aide_upsample2x_bwd(const float* dhi, int dhi_ctot, int dhi_coff, float* dlo, int N, int h, int w, int C,
                        void* stream);

/* ---- nn.ConvTranspose2d(Cin, Cout, kernel_size=2, stride=2) (netblocks.py:12, learned_bilinear=True) ------------- */
/* The transposed convolution equals conv3x3(pad 1) of the zero-inserted input X'[2h,2w] = x[h,w] (0 elsewhere) with
 * K[co,ci,1-a,1-b] = W[ci,co,a,b]; these two entry points are the zero insertion and its transpose, the GEMM work runs
 * on aide_conv3x3_fwd / _dgrad / _wgrad.  src/dst are operand-format views, dhi/dlo fp32. */
int aide_zero_insert2x_fwd(int fmt, const void* src_p0, const void* src_p1, int src_ctot, int src_coff,
                           void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff, int N, int h, int w, int C,
                           void* stream);
int aide_zero_insert2x_bwd(const float* dhi, int dhi_ctot, int dhi_coff, float* dlo, int N, int h, int w, int C,
                           void* stream);

/* ---- last_conv1: 1x1 conv C -> K + bias (fuseunet.py:41,89; UNet.py:150,164) --------------------- */
int aide_conv1x1_fwd(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int C,
                     const float* w /*[K][C]*/, const float* bias, float* logits_nchw, int K,
                     int N, int H, int W, void* stream);
int aide_conv1x1_bwd_rows(int N, int H, int W, int C);
int aide_conv1x1_bwd(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int C,
                     const float* w, const float* dlogits_nchw, int K, int N, int H, int W,
                     float* dx /*[N,H,W,C] fp32*/, float* dw_db /*[K*C] dW then [K] dbias*/,
                     float* partial /*[rows][K*C+K]*/, void* stream);

/* ---- Spatial_Attention gate of the attention variants (netblocks.py:68-89, UNet.py:85-108) ------------------------ */
/* fuseunetsa / fuseunetsaseparate (fuseunet.py:93-208, :210-325) and UNetsa (UNet.py:168-208) multiply every encoder
 * block's output y [N,C,H,W] by gate = sigmoid(BatchNorm2d(1)(conv4(conv3(conv2(conv1(y)))))) with conv1 1x1 C -> r = C/16,
 * conv2 / conv3 3x3 r -> r (dilation d, padding d), conv4 1x1 r -> 1 (weights in nn.Conv2d OIHW layout, fp32).
 * aide_sa_fwd: a1, a2, a3 [N,H,W,r] and a [N,H,W] (conv4 output, before BatchNorm) in fp32 + per-block partial
 *   statistics (sum a, sum a^2): stat_partial [aide_sa_stat_rows(N,H,W)][2] (image-major rows).  Finalise them with
 *   aide_bn_finalize_grouped(C = 1) -> scale_shift [groups][2], mean_rstd [groups][2].
 * aide_sa_gate_apply: gate [N,H,W] = sigmoid(scale * a + shift); writes gate * y into the consumer's channel slice
 *   (dst view) and its 2x2 max-pool into up to two half-resolution views (the fused-encoder concat / next level). */
int aide_sa_stat_rows(int N, int H, int W);
int aide_sa_fwd(int fmt, const void* y_p0, const void* y_p1, int y_ctot, int y_coff, int C, int r, int dilation,
                const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                const float* w4, const float* b4, float* a1, float* a2, float* a3, float* a, float* stat_partial,
                int N, int H, int W, void* stream);
int aide_sa_gate_apply(int fmt, const void* y_p0, const void* y_p1, int y_ctot, int y_coff, const float* a,
                       const float* scale_shift, int N, int imgs_per_group, int H, int W, int C, float* gate,
                       void* dst_p0, void* dst_p1, int dst_ctot, int dst_coff,
                       void* poolA_p0, void* poolA_p1, int poolA_ctot, int poolA_coff,
                       void* poolB_p0, void* poolB_p1, int poolB_ctot, int poolB_coff, void* stream);
/* Backward of t = gate * y.  Stage 1 (aide_sa_bwd_gate): dt = sum of the upstream gradients w.r.t. t (same-resolution
 * slices and slices routed through the 2x2 max-pool of t, first maximum wins); dy [N,H,W,C] fp32 = dt * gate;
 * dahat [N,H,W] = (sum_c dt*y) * gate * (1 - gate); partial [aide_sa_bwd_rows()][2] = block sums of (dahat, dahat*ahat).
 * Stage 2 (aide_sa_bwd_chain): BatchNorm2d(1) backward (train mode), conv4..conv1 backward: adds W1^T da1 into dy and
 * writes the parameter gradients (each bias gradient directly after its weight gradient; dbn = {dbeta, dgamma}). */
int aide_sa_bwd_rows(int N, int H, int W, int pooled);
int aide_sa_bwd_gate(int fmt, const void* y_p0, const void* y_p1, int y_ctot, int y_coff, int C, const float* gate,
                     const float* a, const float* mean_rstd, int N, int H, int W,
                     const float* const* direct_ptr, const int* direct_ctot, const int* direct_coff, int n_direct,
                     const float* const* pool_ptr, const int* pool_ctot, const int* pool_coff, int n_pool,
                     float* dy, float* dahat, float* partial, void* stream);
size_t aide_sa_bwd_workspace_floats(int C, int r, int N, int H, int W);
int aide_sa_bwd_chain(int fmt, const void* y_p0, const void* y_p1, int y_ctot, int y_coff, int C, int r, int dilation,
                      const float* w1, const float* w2, const float* w3, const float* w4, const float* gamma,
                      const float* a1, const float* a2, const float* a3, const float* a, const float* mean_rstd,
                      const float* dahat, const float* partial, int partial_rows, int N, int H, int W,
                      float* workspace, size_t workspace_floats, float* dy,
                      float* dw1, float* db1, float* dw2, float* db2, float* dw3, float* db3, float* dw4, float* db4,
                      float* dbn, void* stream);

/* ---- losses (utils/loss2d.py, utils/metrics2d.py, train_files/trainchaos_proposed_*.py) --------- */
/* One pass over logits [N,2,H,W] (NCHW fp32) + targets [N,H,W] int64 producing per image, in fp64:
 *   sums[n][0]=sum ce*wc[t]  [1]=sum wc[t] (non-ignored)  [2]=sum s*t  [3]=sum s  [4]=sum t
 *   [5]=#(s>=thr & t)  [6]=#(s>=thr)  [7]=sum wm*((1-s-q0)^2+(s-q1)^2)   (s = softmax prob of class 1)
 * q (pseudo label [N,2,H,W]) / wm (weight map [N,1,H,W]) may be NULL (-> [7] = 0).
 * Replaces CrossEntropyLoss2d (loss2d.py:5-13), DiceLoss (:35-61), MulticlassDiceLoss (:87-107),
 * MulticlassMSELoss*weightmap (:109-117 + trainchaos_proposed_30cases1labeled.py:311-313) and the
 * counting part of Dice_fn (metrics2d.py:8-29). */
int aide_loss_sums(const float* logits, const int64_t* targets, const float* q, const float* wm,
                   int N, int H, int W, float wc0, float wc1, int ignore_index, float threshold,
                   double* sums /*[N][8]*/, double* scratch /*[N*aide_loss_blocks()][8]*/, void* stream);
int aide_loss_blocks(int H, int W);
/* per-image loss [N] fp32: w_ce * sums0/(H*W) + w_dice * (1 - (2*sums2+smooth)/(sums3+sums4+smooth))
 * (CEMDiceLossImage.forward, loss2d.py:146-154); dice_fn_out (nullable): batch SUM of per-image
 * Dice with the empty-target rule (metrics2d.py:19-24). */
int aide_loss_image_finalize(const double* sums, int N, int H, int W, float w_ce, float w_dice, float smooth,
                             float* loss_img, float* dice_img, float* dice_fn_out, void* stream);
/* dL/dlogits for L = sum_n [ a_ce[n]*sum_hw ce*wc[t] + a_dice[n]*dice_n + a_mse[n]*sums7_n ]
 * (closed form).  Any coefficient array may be NULL (= 0).  dlogits [N,2,H,W]. */
int aide_loss_bwd(const float* logits, const int64_t* targets, const float* q, const float* wm,
                  const double* sums, const float* a_ce, const float* a_dice, const float* a_mse,
                  int N, int H, int W, float wc0, float wc1, int ignore_index, float smooth,
                  float* dlogits, void* stream);
/* Per-pixel MAPS of the drop-in loss classes (two classes; NCHW fp32 logits, int64 targets) and their gradients:
 *   aide_pixel_loss_*: mode bit 0 -> wc[t] * cross-entropy (CrossEntropyLoss2d(reduction='none'), loss2d.py:5-13; 0 at
 *     ignore_index); bit 1 -> + KL(p1||p2) + KL(p2||p1) of the two softmaxes (KLbidirection, coteach_loss.py:85-92) --
 *     together the "drop" map of Coteachingloss_dropimagedroppixel (coteach_loss.py:231-233).  out / gout: [N,H,W].
 *   aide_softmax_mse_*: (softmax(z) - target)^2, MulticlassMSELoss(reduction='none') (loss2d.py:109-117); [N,2,H,W].
 *   aide_maxpool_nchw_*: max_pool2d(kernel = stride = (kh, kw), ceil_mode=True) on NCHW planes with arg-max for the
 *     backward routing (Coteachingloss_dropregionce pools logits and targets, coteach_loss.py:170-178). */
int aide_pixel_loss_fwd(const float* logits1, const float* logits2, const int64_t* targets, int N, int H, int W,
                        float wc0, float wc1, int ignore_index, int mode, float* out, void* stream);
int aide_pixel_loss_bwd(const float* logits1, const float* logits2, const int64_t* targets, const float* gout,
                        int N, int H, int W, float wc0, float wc1, int ignore_index, int mode,
                        float* dlogits1, float* dlogits2, void* stream);
int aide_softmax_mse_fwd(const float* logits, const float* target, int N, int H, int W, float* out, void* stream);
int aide_softmax_mse_bwd(const float* logits, const float* target, const float* gout, int N, int H, int W,
                         float* dlogits, void* stream);
int aide_maxpool_nchw_fwd(const float* x, int planes, int H, int W, int kh, int kw, float* y, int* argmax, void* stream);
int aide_maxpool_nchw_bwd(const float* gy, const int* argmax, int planes, int H, int W, int kh, int kw, float* gx,
                          void* stream);
/* softmax of n_aug logit tensors, mean, sharpen p^expo / sum, weight map 1-4*q0*q1
 * (trainchaos_proposed_30cases1labeled.py:274-292,97-101). */
int aide_pseudo_label(const float* const* aug_logits, int n_aug, int N, int H, int W, float expo,
                      float* q /*[N,2,H,W]*/, float* wm /*[N,1,H,W]*/, void* stream);
/* Hard segmentation mask argmax(softmax(logits, 1), 1) as uint8 [N,H,W] (pseudo-label / evaluation passes,
 * trainchaos_proposed_30cases1labeled.py:407-409, evalchaos_comparison_1cases.py:208-214); first maximum wins. */
int aide_argmax_mask(const float* logits, uint8_t* mask, int N, int K, int H, int W, void* stream);
/* Reverse augmentation of augmented-forward outputs (trainchaos_proposed_30cases1labeled.py:81-95, which round-trips
 * every plane through PIL on the CPU): optional horizontal flip, then PIL's Image.rotate(-degree, BILINEAR) about the
 * centre with fill 0, reproduced bit for bit.  src/dst: [n_img,K,H,W] fp32 (different buffers); per image a row-major
 * inverse affine matrix (6 doubles, as PIL builds it), a mode (0 affine, 1 copy, 2 rotate-180, 3 / 4 PIL's ROTATE_90 /
 * ROTATE_270 fast paths on square images) and a flip flag -- all device arrays. */
int aide_reverse_aug(const float* src, float* dst, const double* matrices, const int* modes, const int* hflips,
                     int n_img, int K, int H, int W, void* stream);
/* Forward augmentation of the INPUT images (the 4 views per sample of the proposed loop): what the data loader does on
 * the CPU through PIL for every view -- datasetchaos_proposed/transform.py:81-106 Image.rotate(degree, BILINEAR) on the
 * uint8 RGB image, :16-34 FLIP_LEFT_RIGHT (after the rotation), :108-131 ToTensor (/255), :134-170 Normalize with the
 * un-augmented image's per-channel mean / std -- reproduced bit for bit.  src: [n_img,H,W,3] uint8 (PIL "RGB" memory
 * order), dst: [n_img,3,H,W] fp32; matrices / modes as for aide_reverse_aug (built for +degree); mean, std: [n_img][3]. */
int aide_forward_aug(const uint8_t* src_hwc, float* dst_chw, const double* matrices, const int* modes, const int* hflips,
                     const float* mean, const float* stdv, int n_img, int H, int W, void* stream);
/* Small-loss selection (trainchaos_proposed_30cases1labeled.py:305-321): ascending argsort of
 * `pre_other` (the OTHER net's per-image loss) -> idx[N]; the first n_clean images are "clean".
 * Emits this net's per-image coefficients and its scalar loss
 *   loss = seg_w*(mean_clean L + (1-rate)*mean_rest L) + cor_w*rate*sum_rest(sums7)/(n_rest*2*H*W)
 * with L = own per-image CE+Dice loss (loss_img).  a_* are the d loss / d(per-image term) weights
 * consumed by aide_loss_bwd. */
int aide_coteach_select(const float* pre_other, const float* loss_img, const double* sums, int N, int H, int W,
                        int n_clean, float rate, float seg_w, float cor_w, float w_ce, float w_dice,
                        int64_t* idx, float* a_ce, float* a_dice, float* a_mse, float* loss_out, void* stream);
/* The same for a batch that is sharded over data-parallel ranks (the reference's nn.DataParallel sorts the GATHERED
 * per-image losses on GPU 0, trainchaos_proposed_30cases1labeled.py:183-186,303-310): pre_other_all holds the other net's
 * per-image losses of all n_total images (all-gathered, rank-major), this rank owns images [first, first+N).  idx
 * [n_total] is the global argsort; the coefficients cover the N local images with the GLOBAL counts as denominators,
 * and loss_out is this rank's share of the global loss (summing the shares / the gradients over ranks gives the
 * reference's loss / gradient).  NaN losses sort last.  rate_dev (nullable): device scalar that overrides `rate`, so the
 * warm-up schedule (:248) can move under a captured CUDA graph. */
int aide_coteach_select_ex(const float* pre_other_all, int n_total, int first, const float* loss_img,
                           const double* sums, int N, int H, int W, int n_clean, float rate, const float* rate_dev,
                           float seg_w, float cor_w, float w_ce, float w_dice, int64_t* idx, float* a_ce,
                           float* a_dice, float* a_mse, float* loss_out, void* stream);

/* ---- optimiser ("next" row f1): torch.optim.Adam(amsgrad=True) on a flat fp32 buffer ------------- */
int aide_adam_amsgrad(float* p, const float* g, float* m, float* v, float* vmax, size_t n,
                      float lr, float beta1, float beta2, float eps, int step, float grad_scale, void* stream);
/* Same update with the step count kept ON THE DEVICE: *step_counter is incremented and the two bias corrections are
 * written to bc_scratch[2] by a one-thread prelude kernel -- no host scalar changes between steps, so a captured
 * CUDA graph of the whole training step can be replayed.  lr_dev (nullable): device scalar that overrides `lr`, so
 * StepLR / PolyLR (trainchaos_proposed_30cases1labeled.py:236-241) keep working under a captured graph. */
int aide_adam_amsgrad_dev(float* p, const float* g, float* m, float* v, float* vmax, size_t n,
                          float lr, float beta1, float beta2, float eps, int* step_counter, float* bc_scratch,
                          float grad_scale, const float* lr_dev, void* stream);

/* ---- gradient all-reduce over NVLink / NVSwitch peer memory (csrc/comm.cu) ------------------------------------------------
 * Replaces the gradient reduction of nn.DataParallel (train_files/trainchaos_proposed_30cases1labeled.py:188-189; SURVEY.md
 * 8e) for the one-process-per-GPU layout.  Buffers that take part are allocated with aide_comm_alloc (plain cudaMalloc +
 * a 64-byte CUDA IPC handle), the handles travel through the host-side process group once, every rank maps its peers'
 * allocations with aide_comm_open.  aide_allreduce_p2p sums the float range [lo, lo + count) of the `world` buffers in place
 * (two-shot: reduce-scatter by peer loads in rank order, then all-gather; bit-identical on every rank).  pads: one flag pad
 * of aide_comm_pad_words(world, channels) zeroed 32-bit words per rank.  All ranks must issue the calls of a channel in the
 * same order with the same (lo, count, blocks); calls that can be in flight together use different channels. */
int aide_comm_alloc(size_t bytes, void** ptr, unsigned char* handle /*[64]*/);
int aide_comm_open(const unsigned char* handle /*[64]*/, void** ptr);
int aide_comm_close(void* ptr);
int aide_comm_free(void* ptr);
int aide_comm_pad_words(int world, int channels);
int aide_allreduce_p2p(float* const* bufs, unsigned int* const* pads, int rank, int world, int channel, int channels,
                       size_t lo, size_t count, int blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AIDE_B200_H_ */
