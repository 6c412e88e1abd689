// augment.cu -- reverse augmentation of the augmented forwards' logits on the GPU ("next" row f2).
//
// Reference: train_files/trainchaos_proposed_30cases1labeled.py:81-95 sends every (sample, view, class) plane through
// PIL on the CPU: optional FLIP_LEFT_RIGHT, then Image.rotate(-degree, BILINEAR) about the image centre, fill 0.  This
// kernel reproduces Pillow's arithmetic bit for bit (src/libImaging/Geometry.c: affine_transform + bilinear_filter32F):
// source position in double from the inverse affine matrix at the pixel centre, neighbours clamped to the image,
// horizontal differences rounded to float32 first, everything else in double without fused multiply-adds, result
// rounded to float32, source positions outside the image -> 0.  PIL's exact fast paths (angle % 360 == 0, 180, and
// 90/270 on square images) arrive as `mode` 1..4 and are pure index permutations.  HBM-bound: 4 B in, 4 B out per pixel.
#include "common.cuh"

namespace aide {

__global__ void reverse_aug_kernel(const float* __restrict__ src, float* __restrict__ dst, const double* __restrict__ mats,
                                   const int* __restrict__ modes, const int* __restrict__ flips, int n, int K, int H,
                                   int W) {
  const size_t total = (size_t)n * K * H * W;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const size_t plane = idx / ((size_t)W * H);
    const int img = (int)(plane / K);
    const float* s = src + plane * (size_t)H * W;
    const int mode = modes[img];
    const bool flip = flips[img] != 0;
    auto at = [&](int yy, int xx) { return s[(size_t)yy * W + (flip ? W - 1 - xx : xx)]; };
    float out = 0.f;
    if (mode == 1) {
      out = at(y, x);
    } else if (mode == 2) {
      out = at(H - 1 - y, W - 1 - x);
    } else if (mode == 3) {                       // Transpose.ROTATE_90 (counter-clockwise), square images only
      out = at(x, W - 1 - y);
    } else if (mode == 4) {                       // Transpose.ROTATE_270
      out = at(H - 1 - x, y);
    } else {
      const double* m = mats + (size_t)img * 6;
      const double px = (double)x + 0.5, py = (double)y + 0.5;
      double xin = __dadd_rn(__dadd_rn(__dmul_rn(m[0], px), __dmul_rn(m[1], py)), m[2]);
      double yin = __dadd_rn(__dadd_rn(__dmul_rn(m[3], px), __dmul_rn(m[4], py)), m[5]);
      if (xin >= 0.0 && xin < (double)W && yin >= 0.0 && yin < (double)H) {
        xin -= 0.5;
        yin -= 0.5;
        const int xi = (int)floor(xin), yi = (int)floor(yin);
        const double dx = xin - (double)xi, dy = yin - (double)yi;
        const int x0 = min(max(xi, 0), W - 1), x1 = min(max(xi + 1, 0), W - 1);
        const int yc = min(max(yi, 0), H - 1);
        const float a = at(yc, x0), b = at(yc, x1);
        double v1 = __dadd_rn((double)a, __dmul_rn((double)__fsub_rn(b, a), dx));
        double v2 = v1;
        if (yi + 1 >= 0 && yi + 1 < H) {
          const float c = at(yi + 1, x0), d = at(yi + 1, x1);
          v2 = __dadd_rn((double)c, __dmul_rn((double)__fsub_rn(d, c), dx));
        }
        out = (float)__dadd_rn(v1, __dmul_rn(__dsub_rn(v2, v1), dy));
      }
    }
    dst[idx] = out;
  }
}

// ---------------------------------------------------------------- forward augmentation of the input images
// Reference: datasetchaos_proposed/transform.py:81-106 (RandomRotate: Image.rotate(degree, BILINEAR) on the uint8 RGB
// image), :16-34 (RandomHorizontallyFlip: FLIP_LEFT_RIGHT, applied AFTER the rotation -- Compose order of
// train_files/trainchaos_proposed_30cases1labeled.py:191-197), :108-131 (ToTensor: HWC uint8 -> CHW float / 255),
// :134-170 (Normalize with the mean / std of the un-augmented image).  Pillow's 8-bit bilinear filter
// (Geometry.c: bilinear_filter32RGB): neighbours clamped, interpolation in double, result TRUNCATED to uint8, source
// positions outside the image -> 0.  One thread per output element; bit-exact against the PIL + torch chain.
__global__ void forward_aug_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, const double* __restrict__ mats,
                                   const int* __restrict__ modes, const int* __restrict__ flips,
                                   const float* __restrict__ mean, const float* __restrict__ stdv, int n, int H, int W) {
  const size_t total = (size_t)n * 3 * H * W;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int xo = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const int c = (int)((idx / ((size_t)W * H)) % 3);
    const int img = (int)(idx / ((size_t)3 * W * H));
    const uint8_t* s = src + (size_t)img * H * W * 3;
    const int x = flips[img] ? W - 1 - xo : xo;            // the flip acts on the rotated image
    auto at = [&](int yy, int xx) { return (int)s[((size_t)yy * W + xx) * 3 + c]; };
    const int mode = modes[img];
    int u = 0;
    if (mode == 1) {
      u = at(y, x);
    } else if (mode == 2) {
      u = at(H - 1 - y, W - 1 - x);
    } else if (mode == 3) {
      u = at(x, W - 1 - y);
    } else if (mode == 4) {
      u = at(H - 1 - x, y);
    } else {
      const double* m = mats + (size_t)img * 6;
      const double px = (double)x + 0.5, py = (double)y + 0.5;
      double xin = __dadd_rn(__dadd_rn(__dmul_rn(m[0], px), __dmul_rn(m[1], py)), m[2]);
      double yin = __dadd_rn(__dadd_rn(__dmul_rn(m[3], px), __dmul_rn(m[4], py)), m[5]);
      if (xin >= 0.0 && xin < (double)W && yin >= 0.0 && yin < (double)H) {
        xin -= 0.5;
        yin -= 0.5;
        const int xi = (int)floor(xin), yi = (int)floor(yin);
        const double dx = xin - (double)xi, dy = yin - (double)yi;
        const int x0 = min(max(xi, 0), W - 1), x1 = min(max(xi + 1, 0), W - 1);
        const int yc = min(max(yi, 0), H - 1);
        const int a = at(yc, x0), b = at(yc, x1);
        double v1 = __dadd_rn((double)a, __dmul_rn((double)(b - a), dx));
        double v2 = v1;
        if (yi + 1 >= 0 && yi + 1 < H) {
          const int cc = at(yi + 1, x0), d = at(yi + 1, x1);
          v2 = __dadd_rn((double)cc, __dmul_rn((double)(d - cc), dx));
        }
        u = (int)(uint8_t)__dadd_rn(v1, __dmul_rn(__dsub_rn(v2, v1), dy));
      }
    }
    const float v = __fdiv_rn((float)u, 255.0f);
    dst[idx] = __fdiv_rn(__fsub_rn(v, mean[img * 3 + c]), stdv[img * 3 + c]);
  }
}

}  // namespace aide

using namespace aide;

extern "C" int aide_forward_aug(const uint8_t* src_hwc, float* dst_chw, const double* matrices, const int* modes,
                                const int* hflips, const float* mean, const float* stdv, int n_img, int H, int W,
                                void* stream) {
  AIDE_REQUIRE(src_hwc && dst_chw && matrices && modes && hflips && mean && stdv && n_img > 0 && H > 0 && W > 0,
               "forward_aug: bad arguments");
  const size_t total = (size_t)n_img * 3 * H * W;
  long long blocks = (long long)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  forward_aug_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(src_hwc, dst_chw, matrices, modes, hflips, mean, stdv,
                                                                  n_img, H, W);
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_reverse_aug(const float* src, float* dst, const double* matrices, const int* modes, const int* hflips,
                                int n_img, int K, int H, int W, void* stream) {
  AIDE_REQUIRE(src && dst && src != dst && matrices && modes && hflips && n_img > 0 && K > 0 && H > 0 && W > 0,
               "reverse_aug: bad arguments (src and dst must be different buffers)");
  const size_t total = (size_t)n_img * K * H * W;
  long long blocks = (long long)((total + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  reverse_aug_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(src, dst, matrices, modes, hflips, n_img, K, H, W);
  AIDE_CHECK_LAUNCH();
  return 0;
}
