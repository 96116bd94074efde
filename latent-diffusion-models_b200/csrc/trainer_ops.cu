// Optimizer step and image output stage: the two elementwise passes that follow the hot path in the reference's
// trainer (src/Trainer.py:68-71 -> torch.optim.Adam defaults) and sample writers (src/transforms.py:22-35,
// src/utils.py:121-130 -> torchvision save_image).  Both are one streaming pass, HBM bound.
#include "common.cuh"
#include "kernels.h"

namespace {

struct AdamArgs {
  float step_size;      // lr / (1 - beta1^step)
  float beta2, eps;
  float one_minus_beta1, one_minus_beta2;   // rounded from the double differences, as torch does with its Python scalars
  float bc2_sqrt;       // sqrt(1 - beta2^step)
  float grad_scale;     // gradients are multiplied by this first (1/world_size after an all-reduce(sum), 1/loss_scale)
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
  g = __fmul_rn(g, a.grad_scale);
  m = __fadd_rn(m, __fmul_rn(a.one_minus_beta1, __fsub_rn(g, m)));                      // exp_avg.lerp_(grad, 1 - beta1)
  v = __fadd_rn(__fmul_rn(v, a.beta2), __fmul_rn(__fmul_rn(a.one_minus_beta2, g), g));  // mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), a.bc2_sqrt), a.eps);
  p = __fadd_rn(p, __fmul_rn(-a.step_size, __fdiv_rn(m, denom)));                   // addcdiv_(exp_avg, denom, -step_size)
}

// 28 bytes of traffic per parameter (read p, g, m, v; write p, m, v); 4 parameters per thread
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            const AdamArgs a) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (i + 4 <= n) {
    float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g + i));
    adam_one(pp.x, gg.x, mm.x, vv.x, a);
    adam_one(pp.y, gg.y, mm.y, vv.y, a);
    adam_one(pp.z, gg.z, mm.z, vv.z, a);
    adam_one(pp.w, gg.w, mm.w, vv.w, a);
    *reinterpret_cast<float4*>(p + i) = pp;
    *reinterpret_cast<float4*>(m + i) = mm;
    *reinterpret_cast<float4*>(v + i) = vv;
  } else {
    for (int64_t j = i; j < n; ++j) adam_one(p[j], g[j], m[j], v[j], a);
  }
}

// fp32 NCHW images -> uint8 NHWC (the byte layout PIL / PNG writers take).
//   convention 0: torchvision.utils.save_image on the raw tensor: x*255 + 0.5, clamp to [0,255], truncate
//   convention 1: the reference's reverse transform: ((x+1)/2)*255, numpy astype(uint8) = truncate toward zero and keep the
//                 low byte (values outside [0,256) wrap, as numpy does on x86-64 for |v| < 2^31)
__global__ void __launch_bounds__(256)
images_to_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int C, int HW, int64_t total, int convention) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // output index (n, p, c)
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t np_ = i / C;
  const int p = (int)(np_ % HW);
  const int64_t n = np_ / HW;
  const float v = __ldg(x + (n * C + c) * HW + p);
  int r;
  if (convention == 0) {
    float f = __fadd_rn(__fmul_rn(v, 255.f), 0.5f);
    f = fminf(fmaxf(f, 0.f), 255.f);            // NaN -> 0 through fmaxf, torch's clamp would keep NaN; then cast gives 0
    r = __float2int_rz(f);
  } else {
    const float f = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.f), 0.5f), 255.f);
    r = __float2int_rz(f) & 0xff;
  }
  out[i] = (uint8_t)r;
}

// sum((a-b)^2) over n elements, fp32 -> one float (atomically accumulated; caller zeroes and divides)
__global__ void __launch_bounds__(256)
sq_diff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t n, float scale) {
  __shared__ float red[8];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    s = fmaf(d, d, s);
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(out, t * scale);
  }
}

// d/dpred mean((pred - target)^2) = 2 (pred - target) / n, times the incoming gradient of the (scalar) loss
__global__ void __launch_bounds__(256)
mse_backward_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ gout,
                    float* __restrict__ dpred, int64_t n, float two_over_n) {
  const float g = gout ? __ldg(gout) : 1.0f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dpred[i] = __fmul_rn(__fmul_rn(two_over_n, __fsub_rn(pred[i], target[i])), g);
}

}  // namespace

int k_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
                int step, double grad_scale, cudaStream_t st) {
  if (n == 0) return 0;
  LDM_REQUIRE(step >= 1, "adam_step: step counts from 1");
  LDM_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam_step: buffers must be 16-byte aligned");
  AdamArgs a;
  const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  a.step_size = (float)(lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.beta2 = (float)beta2; a.eps = (float)eps; a.grad_scale = (float)grad_scale;
  a.one_minus_beta1 = (float)(1.0 - beta1); a.one_minus_beta2 = (float)(1.0 - beta2);
  adam_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, st>>>(p, g, m, v, n, a);
  LDM_LAUNCHED("adam_step");
  return 0;
}

int k_mse_backward(const float* pred, const float* target, const float* gout, float* dpred, int64_t n, cudaStream_t st) {
  LDM_REQUIRE(n > 0, "mse_backward: empty input");
  const int grid = (int)((n + 256 * 4 - 1) / (256 * 4));
  mse_backward_kernel<<<grid < 1184 ? grid : 1184, 256, 0, st>>>(pred, target, gout, dpred, n, 2.0f / (float)n);
  LDM_LAUNCHED("mse_backward");
  return 0;
}

int k_images_to_u8(const float* x, uint8_t* out, int batch, int C, int hw, int convention, cudaStream_t st) {
  const int64_t total = (int64_t)batch * C * hw;
  if (total == 0) return 0;
  LDM_REQUIRE(convention == 0 || convention == 1, "images_to_uint8: unknown convention %d", convention);
  images_to_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(x, out, C, hw, total, convention);
  LDM_LAUNCHED("images_to_uint8");
  return 0;
}

int k_mse(const float* a, const float* b, float* out, int64_t n, cudaStream_t st) {
  LDM_REQUIRE(n > 0, "mse: empty input");
  LDM_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  const int grid = (int)((n + 256 * 8 - 1) / (256 * 8));
  sq_diff_sum_kernel<<<grid < 1184 ? grid : 1184, 256, 0, st>>>(a, b, out, n, 1.f / (float)n);
  LDM_LAUNCHED("mse");
  return 0;
}
