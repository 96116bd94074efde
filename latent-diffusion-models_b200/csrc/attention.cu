// Attention cores on NHWC qkv tensors (the 1x1 to_qkv / to_out convolutions run on the conv kernels).
//   LinearAttention  src/UNet.py:149-163   q <- softmax_d(q) * 32^-1/2 ; k <- softmax_n(k) ;
//                                          ctx = k v^T (32x32 per head) ; out = ctx^T q
//   Attention        src/UNet.py:122-135   sim = (q 32^-1/2)^T k ; softmax_j ; out = attn v^T
// qkv is [B, N, 384]: channel = part*128 + head*32 + c (part 0/1/2 = q/k/v; "b (h c) x y" split, :124-127);
// out is [B, N, 128] with channel = head*32 + c.
// One CTA per (sample, head).  fp32 math throughout; T only selects the storage type.
#include "kernels.h"

#define LA_HEADS 4
#define LA_D 32
#define LA_TILE 64
#define LA_LD 36  // padded row stride (floats): 16-byte aligned rows, conflict-free column access

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) { load_chunk(p, v); }
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float a[4], b[4];
  load_chunk(p, a);
  load_chunk(p + 4, b);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[i] = a[i]; v[4 + i] = b[i]; }
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<bf16>(bf16* p, const float (&v)[8]) { store_chunk(p, v); }
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  float a[4] = {v[0], v[1], v[2], v[3]}, b[4] = {v[4], v[5], v[6], v[7]};
  store_chunk(p, a);
  store_chunk(p + 4, b);
}
template <typename T>
__device__ __forceinline__ float exp_t(float x) { return sizeof(T) == 4 ? expf(x) : __expf(x); }

template <typename T>
__global__ void __launch_bounds__(256) linattn_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N) {
  __shared__ __align__(16) float tA[LA_TILE * LA_LD];
  __shared__ __align__(16) float tB[LA_TILE * LA_LD];
  __shared__ __align__(16) float ctx[LA_D * LA_LD];
  __shared__ float red[8][LA_D];
  __shared__ float kmax[LA_D], zinv[LA_D];
  const int b = blockIdx.x / LA_HEADS, h = blockIdx.x % LA_HEADS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* base = qkv + (int64_t)b * N * 384 + h * LA_D;
  // ---- phase A: per-channel max of k over the N tokens
  float m = -INFINITY;
  for (int n = warp; n < N; n += 8) m = fmaxf(m, to_float(base[(int64_t)n * 384 + 128 + lane]));
  red[warp][lane] = m;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w][lane]);
    kmax[lane] = m;
  }
  __syncthreads();
  // ---- phase B: ctx[d][e] = sum_n exp(k[n][d]-kmax[d]) v[n][e],  Z[d] = sum_n exp(...)
  const int lp = tid >> 2, lc = (tid & 3) * 8;  // loader mapping: pixel-in-tile, 8-channel part
  const int d = tid >> 3, e0 = (tid & 7) * 4;   // accumulator mapping
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, zacc = 0.f;
  for (int n0 = 0; n0 < N; n0 += LA_TILE) {
    {
      float kv[8], vv[8];
      int n = n0 + lp;
      if (n < N) {
        load8<T>(base + (int64_t)n * 384 + 128 + lc, kv);
        load8<T>(base + (int64_t)n * 384 + 256 + lc, vv);
#pragma unroll
        for (int i = 0; i < 8; ++i) kv[i] = exp_t<T>(kv[i] - kmax[lc + i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) { kv[i] = 0.f; vv[i] = 0.f; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) { tA[lp * LA_LD + lc + i] = kv[i]; tB[lp * LA_LD + lc + i] = vv[i]; }
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < LA_TILE; ++n) {
      float p = tA[n * LA_LD + d];
      float4 v4 = *reinterpret_cast<const float4*>(&tB[n * LA_LD + e0]);
      acc[0] = fmaf(p, v4.x, acc[0]); acc[1] = fmaf(p, v4.y, acc[1]);
      acc[2] = fmaf(p, v4.z, acc[2]); acc[3] = fmaf(p, v4.w, acc[3]);
      zacc += p;
    }
    __syncthreads();
  }
  if ((tid & 7) == 0) zinv[d] = 0.17677669529663687f / zacc;  // 32^-1/2 (q scale, :158) / softmax denominator
  __syncthreads();
  {
    float zi = zinv[d];
#pragma unroll
    for (int i = 0; i < 4; ++i) ctx[d * LA_LD + e0 + i] = acc[i] * zi;
  }
  __syncthreads();
  // ---- phase C: out[n][e] = sum_d ctx[d][e] * softmax_d(q[n][:])[d]
  const int oe = (tid & 3) * 8;
  for (int n0 = 0; n0 < N; n0 += LA_TILE) {
    int n = n0 + lp;
    {
      float qv[8];
      if (n < N) load8<T>(base + (int64_t)n * 384 + lc, qv);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) qv[i] = 0.f;
      }
      float mx = qv[0];
#pragma unroll
      for (int i = 1; i < 8; ++i) mx = fmaxf(mx, qv[i]);
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) { qv[i] = exp_t<T>(qv[i] - mx); s += qv[i]; }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      float inv = 1.0f / s;
#pragma unroll
      for (int i = 0; i < 8; ++i) tA[lp * LA_LD + lc + i] = qv[i] * inv;
    }
    __syncthreads();
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int dd = 0; dd < LA_D; ++dd) {
      float qd = tA[lp * LA_LD + dd];
      float4 c0 = *reinterpret_cast<const float4*>(&ctx[dd * LA_LD + oe]);
      float4 c1 = *reinterpret_cast<const float4*>(&ctx[dd * LA_LD + oe + 4]);
      o[0] = fmaf(qd, c0.x, o[0]); o[1] = fmaf(qd, c0.y, o[1]); o[2] = fmaf(qd, c0.z, o[2]); o[3] = fmaf(qd, c0.w, o[3]);
      o[4] = fmaf(qd, c1.x, o[4]); o[5] = fmaf(qd, c1.y, o[5]); o[6] = fmaf(qd, c1.z, o[6]); o[7] = fmaf(qd, c1.w, o[7]);
    }
    if (n < N) store8<T>(out + ((int64_t)b * N + n) * 128 + h * LA_D + oe, o);
    __syncthreads();
  }
}

int k_linear_attention(const void* qkv, void* out, int batch, int n_tokens, int dtype, cudaStream_t st) {
  if (batch == 0 || n_tokens == 0) return 0;
  int grid = batch * LA_HEADS;
  if (dtype == LDM_DT_BF16) linattn_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)qkv, (bf16*)out, n_tokens);
  else linattn_kernel<float><<<grid, 256, 0, st>>>((const float*)qkv, (float*)out, n_tokens);
  LDM_LAUNCHED("linear_attention");
  return 0;
}

// ---- full softmax attention, one thread per query token (N <= 256; N = 4 in the reference config)
template <typename T>
__global__ void attn_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N) {
  extern __shared__ __align__(16) float sm[];  // k[N][LA_LD], v[N][LA_LD]
  float* sk = sm;
  float* sv = sm + (size_t)N * LA_LD;
  const int b = blockIdx.x / LA_HEADS, h = blockIdx.x % LA_HEADS;
  const T* base = qkv + (int64_t)b * N * 384 + h * LA_D;
  for (int idx = threadIdx.x; idx < N * 4; idx += blockDim.x) {
    int n = idx >> 2, c = (idx & 3) * 8;
    float kv[8], vv[8];
    load8<T>(base + (int64_t)n * 384 + 128 + c, kv);
    load8<T>(base + (int64_t)n * 384 + 256 + c, vv);
#pragma unroll
    for (int i = 0; i < 8; ++i) { sk[n * LA_LD + c + i] = kv[i]; sv[n * LA_LD + c + i] = vv[i]; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float q[LA_D];
#pragma unroll
    for (int c = 0; c < LA_D; c += 8) {
      float t8[8];
      load8<T>(base + (int64_t)i * 384 + c, t8);
#pragma unroll
      for (int u = 0; u < 8; ++u) q[c + u] = t8[u] * 0.17677669529663687f;  // q * scale before QK^T (:128)
    }
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < LA_D; ++c) s = fmaf(q[c], sk[j * LA_LD + c], s);
      mx = fmaxf(mx, s);
    }
    float o[LA_D];
#pragma unroll
    for (int c = 0; c < LA_D; ++c) o[c] = 0.f;
    float den = 0.f;
    for (int j = 0; j < N; ++j) {
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < LA_D; ++c) s = fmaf(q[c], sk[j * LA_LD + c], s);
      float p = exp_t<T>(s - mx);
      den += p;
#pragma unroll
      for (int c = 0; c < LA_D; ++c) o[c] = fmaf(p, sv[j * LA_LD + c], o[c]);
    }
    float inv = 1.0f / den;
    T* op = out + ((int64_t)b * N + i) * 128 + h * LA_D;
#pragma unroll
    for (int c = 0; c < LA_D; c += 8) {
      float t8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) t8[u] = o[c + u] * inv;
      store8<T>(op + c, t8);
    }
  }
}

int k_attention(const void* qkv, void* out, int batch, int n_tokens, int dtype, cudaStream_t st) {
  if (batch == 0 || n_tokens == 0) return 0;
  LDM_REQUIRE(n_tokens <= 256, "attention: %d tokens > 256 not supported by the bottleneck kernel", n_tokens);
  size_t smem = (size_t)2 * n_tokens * LA_LD * sizeof(float);
  int threads = n_tokens < 32 ? 32 : (n_tokens + 31) / 32 * 32;
  if (threads > 256) threads = 256;
  int grid = batch * LA_HEADS;
  if (dtype == LDM_DT_BF16) {
    if (smem > 48 * 1024)
      LDM_CUDA(cudaFuncSetAttribute(attn_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_kernel<bf16><<<grid, threads, smem, st>>>((const bf16*)qkv, (bf16*)out, n_tokens);
  } else {
    if (smem > 48 * 1024)
      LDM_CUDA(cudaFuncSetAttribute(attn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_kernel<float><<<grid, threads, smem, st>>>((const float*)qkv, (float*)out, n_tokens);
  }
  LDM_LAUNCHED("attention");
  return 0;
}
