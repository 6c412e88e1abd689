// adam.cu -- torch.optim.Adam(lr, betas, eps, amsgrad=True) over one flat fp32 buffer
// (reference: train_files/trainchaos_proposed_30cases1labeled.py:231-232,323,325 -> torch.optim.Adam;
// formula of torch/optim/adam.py::_single_tensor_adam, weight_decay = 0, maximize = False).
// HBM-bound: 5 reads + 4 writes of fp32 per parameter, 128-bit accesses.
#include "common.cuh"

namespace aide {

__global__ void adam_amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                    float* __restrict__ v, float* __restrict__ vmax, size_t n, float lr, float b1,
                                    float b2, float eps, float bc1, float bc2_sqrt, float gscale,
                                    const float* __restrict__ bc_dev, const float* __restrict__ lr_dev) {
  if (bc_dev) {               // bias corrections computed on the device (CUDA-graph friendly step counter)
    bc1 = bc_dev[0];
    bc2_sqrt = bc_dev[1];
  }
  if (lr_dev) lr = *lr_dev;   // learning rate read from device memory: a scheduler can change it under a captured graph
  const float step_size = lr / bc1;
  size_t n4 = n >> 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float4 xv = reinterpret_cast<float4*>(vmax)[i];
    float* pp = reinterpret_cast<float*>(&pv);
    float* gp = reinterpret_cast<float*>(&gv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
    float* xp = reinterpret_cast<float*>(&xv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = gp[k] * gscale;
      mp[k] = mp[k] + (gg - mp[k]) * (1.0f - b1);          // exp_avg.lerp_(grad, 1-beta1)
      vp[k] = vp[k] * b2 + (1.0f - b2) * gg * gg;          // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
      xp[k] = fmaxf(xp[k], vp[k]);                         // max_exp_avg_sqs
      float denom = sqrtf(xp[k]) / bc2_sqrt + eps;
      pp[k] = pp[k] - step_size * (mp[k] / denom);         // param.addcdiv_(exp_avg, denom, -step_size)
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(vmax)[i] = xv;
  }
  // tail
  size_t t = (n4 << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t < n) {
    float gg = g[t] * gscale;
    float mm = m[t] + (gg - m[t]) * (1.0f - b1);
    float vv = v[t] * b2 + (1.0f - b2) * gg * gg;
    float xx = fmaxf(vmax[t], vv);
    m[t] = mm; v[t] = vv; vmax[t] = xx;
    p[t] = p[t] - step_size * (mm / (sqrtf(xx) / bc2_sqrt + eps));
  }
}

// step = ++(*counter); bc[0] = 1 - beta1^step; bc[1] = sqrt(1 - beta2^step)
__global__ void adam_prelude_kernel(int* __restrict__ counter, float* __restrict__ bc, double b1, double b2) {
  const int step = *counter + 1;
  *counter = step;
  bc[0] = (float)(1.0 - pow(b1, (double)step));
  bc[1] = (float)sqrt(1.0 - pow(b2, (double)step));
}

}  // namespace aide

using namespace aide;

static int adam_launch(float* p, const float* g, float* m, float* v, float* vmax, size_t n, float lr, float beta1,
                       float beta2, float eps, float bc1, float bc2_sqrt, float grad_scale, const float* bc_dev,
                       const float* lr_dev, cudaStream_t st) {
  int blocks = (int)(((n >> 2) + 255) / 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  if (blocks < 1) blocks = 1;
  adam_amsgrad_kernel<<<blocks, 256, 0, st>>>(p, g, m, v, vmax, n, lr, beta1, beta2, eps, bc1, bc2_sqrt, grad_scale,
                                              bc_dev, lr_dev);
  AIDE_CHECK_LAUNCH();
  return 0;
}

extern "C" int aide_adam_amsgrad(float* p, const float* g, float* m, float* v, float* vmax, size_t n, float lr,
                                 float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  AIDE_REQUIRE(p && g && m && v && vmax && step >= 1, "adam_amsgrad: bad arguments");
  AIDE_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)vmax) % 16 == 0,
               "adam_amsgrad: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  return adam_launch(p, g, m, v, vmax, n, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale, nullptr,
                     nullptr, as_stream(stream));
}

extern "C" int aide_adam_amsgrad_dev(float* p, const float* g, float* m, float* v, float* vmax, size_t n, float lr,
                                     float beta1, float beta2, float eps, int* step_counter, float* bc_scratch,
                                     float grad_scale, const float* lr_dev, void* stream) {
  AIDE_REQUIRE(p && g && m && v && vmax && step_counter && bc_scratch, "adam_amsgrad_dev: bad arguments");
  AIDE_REQUIRE(((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v | (uintptr_t)vmax) % 16 == 0,
               "adam_amsgrad_dev: buffers must be 16-byte aligned");
  adam_prelude_kernel<<<1, 1, 0, as_stream(stream)>>>(step_counter, bc_scratch, (double)beta1, (double)beta2);
  AIDE_CHECK_LAUNCH();
  if (n == 0) return 0;
  return adam_launch(p, g, m, v, vmax, n, lr, beta1, beta2, eps, 1.f, 1.f, grad_scale, bc_scratch, lr_dev,
                     as_stream(stream));
}
