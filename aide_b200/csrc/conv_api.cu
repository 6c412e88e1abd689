// conv_api.cu -- C-ABI entry points for conv3x3 forward/dgrad/wgrad: dispatch between the CUDA-core
// exact path (AIDE_FMT_F32, and layers that do not fit an MMA shape) and the tcgen05 path.
#include "common.cuh"

namespace aide {
int simt_stat_rows(int N, int H, int W);
int simt_conv3x3(const float* x, int x_ctot, int x_coff, int cin, const float* w, const float* bias, float* z,
                 int z_ctot, int z_coff, int cout, int N, int H, int W, float* stat_partial, cudaStream_t st);
size_t simt_wgrad_workspace_bytes(int cin, int cout, int N, int H, int W);
int simt_wgrad(const float* x, int x_ctot, int x_coff, int cin, const float* dz, int cout, int N, int H, int W,
               float* ws, size_t ws_bytes, float* dw, cudaStream_t st);
int tc_stat_rows(int N, int H, int W);
bool tc_shape_ok(int fmt, int cin, int cout);
int tc_conv3x3(int fmt, const void* x0, const void* x1, int x_ctot, int x_coff, int cin, const void* w0, const void* w1,
               const float* bias, float* z, int z_ctot, int z_coff, int cout, int N, int H, int W, float* stat_partial,
               float out_scale, const float* out_scale_ptr, cudaStream_t st);
size_t tc_wgrad_workspace_bytes(int fmt, int cin, int cout, int N, int H, int W);
int tc_wgrad(int fmt, const void* x0, const void* x1, int x_ctot, int x_coff, int cin, const void* dz0, const void* dz1,
             int cout, int N, int H, int W, void* ws, size_t ws_bytes, float* dw, float out_scale,
             const float* out_scale_ptr, cudaStream_t st);
bool halo_shape_ok(int fmt, int cin, int cout, int N, int H, int W);
int halo_stat_rows(int N, int H, int W);
struct HaloFused {
  const float* ss;
  View dst, pa, pb;
};
int halo_conv3x3(int fmt, const void* x0, const void* x1, int x_ctot, int x_coff, int cin, const void* w0, const void* w1,
                 const float* bias, float* z, int z_ctot, int z_coff, int cout, int N, int H, int W, float* stat_partial,
                 float out_scale, const float* out_scale_ptr, cudaStream_t st, const HaloFused* fused = nullptr);
bool wgrad_halo_ok(int fmt, int cin, int cout, int N, int H, int W);
size_t wgrad_halo_workspace_bytes(int fmt, int cin, int cout, int N, int H, int W);
int wgrad_halo(int fmt, const void* x0, const void* x1, int x_ctot, int x_coff, int cin, const void* dz0, const void* dz1,
               int cout, int N, int H, int W, void* ws, size_t ws_bytes, float* dw, float out_scale,
               const float* out_scale_ptr, cudaStream_t st);
bool tma_available();
bool c3_shape_ok(int cin, int cout);
int c3_stat_rows(int N, int H, int W);
int c3_conv3x3(const float* x, int x_ctot, int x_coff, const float* w, const float* bias, float* z, int z_ctot,
               int z_coff, int cout, int N, int H, int W, float* stat_partial, cudaStream_t st);
size_t c3_wgrad_workspace_bytes(int cout, int N, int H, int W);
int launch_reduce_rows(const float* partial, int rows, int ld, int cols, float* out, cudaStream_t st);
int c3_wgrad(const float* x, int x_ctot, int x_coff, const float* dz, int cout, int N, int H, int W, float* ws,
             size_t ws_bytes, float* dw, cudaStream_t st);
}  // namespace aide

using namespace aide;

extern "C" int aide_has_tma(void) { return tma_available() ? 1 : 0; }

extern "C" int aide_conv3x3_stat_rows(int fmt, int cin, int cout, int N, int H, int W) {
  if (fmt != AIDE_FMT_F32) return halo_shape_ok(fmt, cin, cout, N, H, W) ? halo_stat_rows(N, H, W) : tc_stat_rows(N, H, W);
  return c3_shape_ok(cin, cout) ? c3_stat_rows(N, H, W) : simt_stat_rows(N, H, W);
}

static int conv3x3_any(int fmt, float op_scale, const float* out_scale_ptr, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int cin,
                                const void* w_p0, const void* w_p1, const float* bias, float* z, int z_ctot,
                                int z_coff, int cout, int N, int H, int W, float* stat_partial, void* stream) {
  AIDE_REQUIRE(x_p0 && w_p0 && z && cin > 0 && cout > 0 && N > 0 && H > 0 && W > 0, "conv3x3_fwd: bad arguments");
  AIDE_REQUIRE(x_coff >= 0 && x_coff + cin <= x_ctot && z_coff >= 0 && z_coff + cout <= z_ctot,
               "conv3x3_fwd: channel view out of range");
  if (fmt == AIDE_FMT_F32 && c3_shape_ok(cin, cout) && z_ctot % 4 == 0 && z_coff % 4 == 0)
    return c3_conv3x3(reinterpret_cast<const float*>(x_p0), x_ctot, x_coff, reinterpret_cast<const float*>(w_p0), bias, z,
                      z_ctot, z_coff, cout, N, H, W, stat_partial, as_stream(stream));
  if (fmt == AIDE_FMT_F32)
    return simt_conv3x3(reinterpret_cast<const float*>(x_p0), x_ctot, x_coff, cin,
                        reinterpret_cast<const float*>(w_p0), bias, z, z_ctot, z_coff, cout, N, H, W, stat_partial,
                        as_stream(stream));
  AIDE_REQUIRE(fmt == AIDE_FMT_TF32X2 || fmt == AIDE_FMT_BF16 || fmt == AIDE_FMT_F16X2,
               "conv3x3_fwd: bad operand format %d", fmt);
  AIDE_REQUIRE(tc_shape_ok(fmt, cin, cout),
               "conv3x3_fwd: tcgen05 path needs cin %% 32 == 0 and cout %% 32 == 0 (got %d -> %d); use AIDE_FMT_F32",
               cin, cout);
  AIDE_REQUIRE(z_ctot % 4 == 0 && z_coff % 4 == 0, "conv3x3_fwd: output view must be 16-byte aligned");
  const float out_scale = op_scale;
  // second-generation kernel (halo reuse across the 9 taps, pixel-tile blocking) wherever the feature map is large
  // enough for whole 8-pixel output rows; the first-generation kernel covers the tiny maps of the deepest levels
  if (halo_shape_ok(fmt, cin, cout, N, H, W))
    return halo_conv3x3(fmt, x_p0, x_p1, x_ctot, x_coff, cin, w_p0, w_p1, bias, z, z_ctot, z_coff, cout, N, H, W,
                        stat_partial, out_scale, out_scale_ptr, as_stream(stream));
  return tc_conv3x3(fmt, x_p0, x_p1, x_ctot, x_coff, cin, w_p0, w_p1, bias, z, z_ctot, z_coff, cout, N, H, W,
                    stat_partial, out_scale, out_scale_ptr, as_stream(stream));
}

extern "C" int aide_conv3x3_fwd(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int cin,
                                const void* w_p0, const void* w_p1, const float* bias, float* z, int z_ctot,
                                int z_coff, int cout, int N, int H, int W, float* stat_partial, void* stream) {
  // F16X2 operands are stored pre-scaled: activations by 2^8, weights by 2^12 (common.cuh)
  const float op_scale = fmt == AIDE_FMT_F16X2 ? 1.0f / (kF16ActScale * kF16WScale) : 1.0f;
  return conv3x3_any(fmt, op_scale, nullptr, x_p0, x_p1, x_ctot, x_coff, cin, w_p0, w_p1, bias, z, z_ctot, z_coff, cout, N, H,
                     W, stat_partial, stream);
}

extern "C" int aide_conv3x3_dgrad(int fmt, const void* dz_p0, const void* dz_p1, int cout, const void* w_p0,
                                  const void* w_p1, const float* dz_inv_scale, float* dx, int dx_ctot, int dx_coff,
                                  int cin, int N, int H, int W, void* stream) {
  // F16X2: dZ is stored as dz * s with s a power of two chosen on the device (aide_bn_relu_bwd_apply writes
  // {s, 1/s}); the epilogue multiplies by 1/s (dz_inv_scale) and by the static 2^-12 of the weight planes
  AIDE_REQUIRE(fmt != AIDE_FMT_F16X2 || dz_inv_scale, "conv3x3_dgrad: F16X2 needs the gradient's inverse scale");
  const float op_scale = fmt == AIDE_FMT_F16X2 ? 1.0f / kF16WScale : 1.0f;
  return conv3x3_any(fmt, op_scale, fmt == AIDE_FMT_F16X2 ? dz_inv_scale : nullptr, dz_p0, dz_p1, cout, 0, cout, w_p0, w_p1,
                     nullptr, dx, dx_ctot, dx_coff, cin, N, H, W, nullptr, stream);
}

extern "C" size_t aide_conv3x3_wgrad_workspace_bytes(int fmt, int cin, int cout, int N, int H, int W) {
  if (fmt == AIDE_FMT_F32 && c3_shape_ok(cin, cout)) return c3_wgrad_workspace_bytes(cout, N, H, W);
  if (fmt == AIDE_FMT_F32) return simt_wgrad_workspace_bytes(cin, cout, N, H, W);
  if (!tc_shape_ok(fmt, cin, cout)) return 0;
  if (wgrad_halo_ok(fmt, cin, cout, N, H, W)) return wgrad_halo_workspace_bytes(fmt, cin, cout, N, H, W);
  return tc_wgrad_workspace_bytes(fmt, cin, cout, N, H, W);
}

extern "C" int aide_conv3x3_wgrad(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int cin,
                                  const void* dz_p0, const void* dz_p1, const float* dz_inv_scale, int cout, int N, int H,
                                  int W, void* workspace, size_t workspace_bytes, float* dw_oihw, void* stream) {
  AIDE_REQUIRE(x_p0 && dz_p0 && dw_oihw && workspace, "conv3x3_wgrad: null argument");
  AIDE_REQUIRE(x_coff >= 0 && x_coff + cin <= x_ctot, "conv3x3_wgrad: channel view out of range");
  if (fmt == AIDE_FMT_F32 && c3_shape_ok(cin, cout))
    return c3_wgrad(reinterpret_cast<const float*>(x_p0), x_ctot, x_coff, reinterpret_cast<const float*>(dz_p0), cout, N,
                    H, W, reinterpret_cast<float*>(workspace), workspace_bytes, dw_oihw, as_stream(stream));
  if (fmt == AIDE_FMT_F32)
    return simt_wgrad(reinterpret_cast<const float*>(x_p0), x_ctot, x_coff, cin, reinterpret_cast<const float*>(dz_p0),
                      cout, N, H, W, reinterpret_cast<float*>(workspace), workspace_bytes, dw_oihw, as_stream(stream));
  AIDE_REQUIRE(fmt == AIDE_FMT_TF32X2 || fmt == AIDE_FMT_BF16 || fmt == AIDE_FMT_F16X2, "conv3x3_wgrad: bad operand format %d",
               fmt);
  AIDE_REQUIRE(fmt != AIDE_FMT_F16X2 || dz_inv_scale, "conv3x3_wgrad: F16X2 needs the gradient's inverse scale");
  AIDE_REQUIRE(tc_shape_ok(fmt, cin, cout), "conv3x3_wgrad: tcgen05 path needs cin %% 32 == 0 and cout %% 32 == 0");
  // F16X2: x planes carry 2^8, dZ planes carry the dynamic scale s -> dW = acc * 2^-8 * (1/s)
  if (wgrad_halo_ok(fmt, cin, cout, N, H, W))
    return wgrad_halo(fmt, x_p0, x_p1, x_ctot, x_coff, cin, dz_p0, dz_p1, cout, N, H, W, workspace, workspace_bytes, dw_oihw,
                      fmt == AIDE_FMT_F16X2 ? 1.0f / kF16ActScale : 1.0f, fmt == AIDE_FMT_F16X2 ? dz_inv_scale : nullptr,
                      as_stream(stream));
  return tc_wgrad(fmt, x_p0, x_p1, x_ctot, x_coff, cin, dz_p0, dz_p1, cout, N, H, W, workspace, workspace_bytes, dw_oihw,
                  fmt == AIDE_FMT_F16X2 ? 1.0f / kF16ActScale : 1.0f, fmt == AIDE_FMT_F16X2 ? dz_inv_scale : nullptr,
                  as_stream(stream));
}

// Same, plus a column sum out[j] = sum_r src[r * cols + j] (rows x cols fp32, e.g. the conv-bias gradient from the partial
// rows of aide_bn_relu_bwd_apply) folded into the split-K reduction launch of the weight gradient.
extern "C" int aide_conv3x3_wgrad_ex(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int cin,
                                     const void* dz_p0, const void* dz_p1, const float* dz_inv_scale, int cout, int N, int H,
                                     int W, void* workspace, size_t workspace_bytes, float* dw_oihw, const float* colsum_src,
                                     int colsum_rows, int colsum_cols, float* colsum_dst, void* stream) {
  const bool want = colsum_src && colsum_dst && colsum_rows > 0 && colsum_cols > 0;
  if (want) wgrad_reduce_post_job(ColSumJob{colsum_src, colsum_rows, colsum_cols, colsum_cols, colsum_dst});
  const int rc = aide_conv3x3_wgrad(fmt, x_p0, x_p1, x_ctot, x_coff, cin, dz_p0, dz_p1, dz_inv_scale, cout, N, H, W, workspace,
                                    workspace_bytes, dw_oihw, stream);
  ColSumJob left;
  if (want && wgrad_reduce_take_job(&left)) {       // a path without the shared reduction kernel (first layer), or an error
    if (rc) return rc;
    return launch_reduce_rows(left.src, left.rows, left.ld, left.cols, left.dst, as_stream(stream));
  }
  return rc;
}

// ---- inference: conv3x3 + eval-mode BatchNorm + ReLU (+ MaxPool2d) in ONE launch ---------------------------------------
extern "C" int aide_conv3x3_bn_relu_ok(int fmt, int cin, int cout, int N, int H, int W) {
  return (fmt == AIDE_FMT_F16X2 || fmt == AIDE_FMT_BF16 || fmt == AIDE_FMT_TF32X2) && tc_shape_ok(fmt, cin, cout) &&
                 halo_shape_ok(fmt, cin, cout, N, H, W)
             ? 1
             : 0;
}

extern "C" int aide_conv3x3_bn_relu_fwd(int fmt, const void* x_p0, const void* x_p1, int x_ctot, int x_coff, int cin,
                                        const void* w_p0, const void* w_p1, const float* bias, const float* scale_shift,
                                        int cout, int N, int H, int W, void* dst_p0, void* dst_p1, int dst_ctot,
                                        int dst_coff, void* poolA_p0, void* poolA_p1, int poolA_ctot, int poolA_coff,
                                        void* poolB_p0, void* poolB_p1, int poolB_ctot, int poolB_coff, void* stream) {
  AIDE_REQUIRE(x_p0 && w_p0 && scale_shift && cin > 0 && cout > 0 && N > 0 && H > 0 && W > 0, "conv3x3_bn_relu_fwd: bad arguments");
  AIDE_REQUIRE(aide_conv3x3_bn_relu_ok(fmt, cin, cout, N, H, W),
               "conv3x3_bn_relu_fwd: shape / format not covered by the fused kernel (use aide_conv3x3_fwd + aide_bn_relu_apply)");
  AIDE_REQUIRE(x_coff >= 0 && x_coff + cin <= x_ctot, "conv3x3_bn_relu_fwd: channel view out of range");
  AIDE_REQUIRE(dst_coff % 8 == 0 && dst_ctot % 8 == 0 && poolA_coff % 8 == 0 && poolA_ctot % 8 == 0 && poolB_coff % 8 == 0 &&
                   poolB_ctot % 8 == 0,
               "conv3x3_bn_relu_fwd: destination channel offsets must be multiples of 8 (128-bit plane stores)");
  HaloFused f{scale_shift, View{dst_p0, dst_p1, dst_ctot, dst_coff}, View{poolA_p0, poolA_p1, poolA_ctot, poolA_coff},
              View{poolB_p0, poolB_p1, poolB_ctot, poolB_coff}};
  const float op_scale = fmt == AIDE_FMT_F16X2 ? 1.0f / (kF16ActScale * kF16WScale) : 1.0f;
  return halo_conv3x3(fmt, x_p0, x_p1, x_ctot, x_coff, cin, w_p0, w_p1, bias, nullptr, cout, 0, cout, N, H, W, nullptr,
                      op_scale, nullptr, as_stream(stream), &f);
}
