// First-stage autoencoder pieces that the UNet path does not already provide (src/Autoencoder.py):
//   * GroupNorm(32 groups, eps 1e-6) + swish when a group is narrower than one vector chunk (2 or 4 channels per group
//     at 64 / 128 channels, :9-18) -- the UNet's GroupNorm kernels need >= 8 (bf16) / 4 (fp32) channels per group
//   * nearest-neighbour 2x upsample (UpSample, :142-157) and the stride-2 pick that turns a full-resolution pad-1 3x3
//     conv into DownSample's pad-(0,1,0,1) stride-2 conv (:160-180)
//   * single-head full attention with head dim = channels (AttnBlock, :87-139)
//   * GaussianDistribution (:21-43): mu, log_var, sigma, z = mu + sigma * eps
// The 3x3 / 1x1 convolutions are the UNet's implicit-GEMM kernels.  Encode / decode run once per batch, next to 1000 UNet
// steps, so these are written for exactness first; all of them are one or three streaming passes over NHWC data.
#include "common.cuh"
#include "kernels.h"

namespace {

template <typename T> __device__ __forceinline__ float ldv(const T* p);
template <> __device__ __forceinline__ float ldv<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldv<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stv(T* p, float v);
template <> __device__ __forceinline__ void stv<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stv<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// ------------------------------------------------------------------ GroupNorm for any group width
// grid (batch); blockDim is a multiple of C, so a thread's channel (and group) is fixed and it walks pixels.
// Exact two-pass statistics (mean, then centred sum of squares) in fp32, then the apply pass; x of one sample is at most
// a few hundred KB, so passes two and three hit L2.
template <typename T>
__global__ void __launch_bounds__(1024)
gn_any_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, const float* __restrict__ gamma,
              const float* __restrict__ beta, int HW, int C, int G, float eps, int silu) {
  __shared__ float s_part[1024];
  __shared__ float s_mean[256], s_rstd[256];
  const int n = blockIdx.x;
  const int c = threadIdx.x % C, pl = threadIdx.x / C, lanes = blockDim.x / C;
  const int cpg = C / G, g = c / cpg;
  const T* xs = x + (int64_t)n * HW * ldx + c;
  const float inv_cnt = 1.f / ((float)HW * (float)cpg);
  // fixed-order reduction of the per-thread partials of one group (bit-reproducible, unlike shared-memory atomics)
  auto group_total = [&](int gg) {
    float t = 0.f;
    for (int l = 0; l < lanes; ++l)
      for (int k = 0; k < cpg; ++k) t += s_part[l * C + gg * cpg + k];
    return t;
  };
  float s = 0.f;
  for (int p = pl; p < HW; p += lanes) s += ldv(xs + (int64_t)p * ldx);
  s_part[threadIdx.x] = s;
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) s_mean[i] = group_total(i) * inv_cnt;
  __syncthreads();
  const float mean = s_mean[g];
  float q = 0.f;
  for (int p = pl; p < HW; p += lanes) { const float d = ldv(xs + (int64_t)p * ldx) - mean; q = fmaf(d, d, q); }
  s_part[threadIdx.x] = q;
  __syncthreads();
  for (int i = threadIdx.x; i < G; i += blockDim.x) s_rstd[i] = rsqrtf(group_total(i) * inv_cnt + eps);
  __syncthreads();
  const float a = s_rstd[g] * gamma[c], b = beta[c] - mean * a;
  T* ys = y + (int64_t)n * HW * ldy + c;
  for (int p = pl; p < HW; p += lanes) {
    float o = fmaf(ldv(xs + (int64_t)p * ldx), a, b);
    if (silu) o = o / (1.f + __expf(-o));
    stv(ys + (int64_t)p * ldy, o);
  }
}

// ------------------------------------------------------------------ nearest 2x upsample / stride-2 pick (NHWC)
template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, int H, int W, int C, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // output element (n, oh, ow, c)
  if (i >= total) return;
  const int c = (int)(i % C);
  int64_t r = i / C;
  const int ow = (int)(r % (2 * W)); r /= 2 * W;
  const int oh = (int)(r % (2 * H));
  const int64_t n = r / (2 * H);
  y[((n * 2 * H + oh) * 2 * W + ow) * ldy + c] = x[((n * H + oh / 2) * W + ow / 2) * ldx + c];
}
// y[n][i][j][c] = x[n][2i+1][2j+1][c]
template <typename T>
__global__ void pick_odd_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, int H, int W, int C, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // output element (n, oh, ow, c), output is H/2 x W/2
  if (i >= total) return;
  const int c = (int)(i % C);
  int64_t r = i / C;
  const int ow = (int)(r % (W / 2)); r /= W / 2;
  const int oh = (int)(r % (H / 2));
  const int64_t n = r / (H / 2);
  y[((n * (H / 2) + oh) * (W / 2) + ow) * ldy + c] = x[((n * H + 2 * oh + 1) * W + 2 * ow + 1) * ldx + c];
}

// ------------------------------------------------------------------ single-head attention, head dim = C
// qkv [B][N][3C] (q | k | v per token).  out[b][i][:] = sum_j softmax_j(scale * q_i . k_j) v_j.
// grid (ceil(N/QT), B), 256 threads; dynamic smem: q tile [QT][C] + scores [QT][N], fp32.
constexpr int A1_QT = 16;
template <typename T>
__global__ void __launch_bounds__(256)
attn1_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N, int C, float scale) {
  extern __shared__ float a1_smem[];
  float* sq = a1_smem;                 // [QT][C]
  float* sp = a1_smem + A1_QT * C;     // [QT][N]
  const int b = blockIdx.y, q0 = blockIdx.x * A1_QT;
  const int nq = min(A1_QT, N - q0);
  const T* base = qkv + (int64_t)b * N * 3 * C;
  for (int i = threadIdx.x; i < A1_QT * C; i += blockDim.x) {
    const int qi = i / C, c = i % C;
    sq[i] = qi < nq ? ldv(base + (int64_t)(q0 + qi) * 3 * C + c) * scale : 0.f;
  }
  __syncthreads();
  // scores: one key per thread at a time, all QT queries against it
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const T* kr = base + (int64_t)j * 3 * C + C;
    float acc[A1_QT];
#pragma unroll
    for (int qi = 0; qi < A1_QT; ++qi) acc[qi] = 0.f;
    for (int c = 0; c < C; ++c) {
      const float kv = ldv(kr + c);
#pragma unroll
      for (int qi = 0; qi < A1_QT; ++qi) acc[qi] = fmaf(sq[qi * C + c], kv, acc[qi]);
    }
#pragma unroll
    for (int qi = 0; qi < A1_QT; ++qi) sp[qi * N + j] = acc[qi];
  }
  __syncthreads();
  // softmax over keys, one warp per query row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int qi = warp; qi < nq; qi += 8) {
    float m = -INFINITY;
    for (int j = lane; j < N; j += 32) m = fmaxf(m, sp[qi * N + j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int j = lane; j < N; j += 32) { const float e = expf(sp[qi * N + j] - m); sp[qi * N + j] = e; s += e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float inv = 1.f / s;
    for (int j = lane; j < N; j += 32) sp[qi * N + j] *= inv;
  }
  __syncthreads();
  // out = P V: one channel per thread, all QT queries
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc[A1_QT];
#pragma unroll
    for (int qi = 0; qi < A1_QT; ++qi) acc[qi] = 0.f;
    const T* vc = base + 2 * C + c;
    for (int j = 0; j < N; ++j) {
      const float vv = ldv(vc + (int64_t)j * 3 * C);
#pragma unroll
      for (int qi = 0; qi < A1_QT; ++qi) acc[qi] = fmaf(sp[qi * N + j], vv, acc[qi]);
    }
    for (int qi = 0; qi < nq; ++qi) stv(out + ((int64_t)b * N + q0 + qi) * C + c, acc[qi]);
  }
}

// ------------------------------------------------------------------ GaussianDistribution
// moments NHWC [B][HW][ld] (channels 0..Z-1 = mu, Z..2Z-1 = log variance) -> fp32 NCHW mu, log_var, sigma, z
template <typename T>
__global__ void gaussian_kernel(const T* __restrict__ moments, int ld, const float* __restrict__ eps, float* __restrict__ mu,
                                float* __restrict__ log_var, float* __restrict__ sigma, float* __restrict__ z, int Z, int HW,
                                int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // NCHW index (n, c, p)
  if (i >= total) return;
  const int p = (int)(i % HW);
  const int64_t r = i / HW;
  const int c = (int)(r % Z);
  const int64_t n = r / Z;
  const T* m = moments + (n * HW + p) * ld;
  const float mv = ldv(m + c), lv = ldv(m + Z + c);
  const float sg = expf(lv * 0.5f);
  if (mu) mu[i] = mv;
  if (log_var) log_var[i] = lv;
  if (sigma) sigma[i] = sg;
  if (z) z[i] = fmaf(sg, eps[i], mv);
}

}  // namespace

#define AE_DISPATCH(dtype, ...)                          \
  do {                                                   \
    if ((dtype) == LDM_DT_BF16) { using T = bf16; __VA_ARGS__; } \
    else { using T = float; __VA_ARGS__; }               \
  } while (0)

int k_group_norm_any(const void* x, int ldx, void* y, int ldy, const float* gamma, const float* beta, int batch, int hw,
                     int channels, int groups, float eps, int silu, int dtype, cudaStream_t st) {
  LDM_REQUIRE(groups >= 1 && groups <= 256 && channels % groups == 0, "group_norm: %d channels / %d groups unsupported", channels, groups);
  LDM_REQUIRE(channels <= 1024, "group_norm: too many channels (%d)", channels);
  if (batch == 0 || hw == 0) return 0;
  int threads = 1024 / channels * channels;
  const int64_t want = (int64_t)channels * hw;
  if (threads > want) threads = (int)((want + channels - 1) / channels) * channels;
  AE_DISPATCH(dtype, gn_any_kernel<T><<<batch, threads, 0, st>>>((const T*)x, ldx, (T*)y, ldy, gamma, beta, hw, channels, groups, eps, silu));
  LDM_LAUNCHED("group_norm_any");
  return 0;
}

int k_upsample_nearest2x(const void* x, int ldx, void* y, int ldy, int batch, int H, int W, int C, int dtype, cudaStream_t st) {
  const int64_t total = (int64_t)batch * 4 * H * W * C;
  if (total == 0) return 0;
  AE_DISPATCH(dtype, upsample2x_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const T*)x, ldx, (T*)y, ldy, H, W, C, total));
  LDM_LAUNCHED("upsample_nearest2x");
  return 0;
}

int k_pick_odd(const void* x, int ldx, void* y, int ldy, int batch, int H, int W, int C, int dtype, cudaStream_t st) {
  LDM_REQUIRE(H % 2 == 0 && W % 2 == 0, "downsample: odd image size %dx%d", H, W);
  const int64_t total = (int64_t)batch * (H / 2) * (W / 2) * C;
  if (total == 0) return 0;
  AE_DISPATCH(dtype, pick_odd_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const T*)x, ldx, (T*)y, ldy, H, W, C, total));
  LDM_LAUNCHED("downsample_pick");
  return 0;
}

int k_attention_single_head(const void* qkv, void* out, int batch, int n_tokens, int channels, int dtype, cudaStream_t st) {
  if (batch == 0 || n_tokens == 0) return 0;
  const size_t smem = (size_t)A1_QT * (channels + n_tokens) * sizeof(float);
  LDM_REQUIRE(smem <= 200 * 1024, "attention_single_head: %d tokens x %d channels exceed shared memory", n_tokens, channels);
  const float scale = 1.f / sqrtf((float)channels);
  const dim3 grid((n_tokens + A1_QT - 1) / A1_QT, batch);
  if (dtype == LDM_DT_BF16) {
    LDM_CUDA(cudaFuncSetAttribute(attn1_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn1_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)qkv, (bf16*)out, n_tokens, channels, scale);
  } else {
    LDM_CUDA(cudaFuncSetAttribute(attn1_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn1_kernel<float><<<grid, 256, smem, st>>>((const float*)qkv, (float*)out, n_tokens, channels, scale);
  }
  LDM_LAUNCHED("attention_single_head");
  return 0;
}

int k_gaussian(const void* moments, int ld, const float* eps, float* mu, float* log_var, float* sigma, float* z, int batch,
               int zc, int hw, int dtype, cudaStream_t st) {
  const int64_t total = (int64_t)batch * zc * hw;
  if (total == 0) return 0;
  LDM_REQUIRE(z == nullptr || eps != nullptr, "gaussian_distribution: sampling needs eps");
  AE_DISPATCH(dtype, gaussian_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const T*)moments, ld, eps, mu, log_var, sigma, z, zc, hw, total));
  LDM_LAUNCHED("gaussian_distribution");
  return 0;
}
