// Shared host/device helpers for the ldm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- host side
int ldm_set_error(const char* fmt, ...);
extern std::atomic<long long> g_ldm_launches;

#define LDM_REQUIRE(cond, ...)                      \
  do {                                              \
    if (!(cond)) return ldm_set_error(__VA_ARGS__); \
  } while (0)

#define LDM_CUDA(expr)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return ldm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                           __LINE__);                                                         \
  } while (0)

// Call right after every <<<>>> launch: counts it and surfaces launch-configuration errors.
#define LDM_LAUNCHED(name)                                                                 \
  do {                                                                                     \
    g_ldm_launches.fetch_add(1, std::memory_order_relaxed);                                \
    cudaError_t e__ = cudaPeekAtLastError();                                               \
    if (e__ != cudaSuccess) {                                                              \
      cudaGetLastError();                                                                  \
      return ldm_set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));      \
    }                                                                                      \
  } while (0)

// Programmatic dependent launch: kernels of the per-timestep chain are launched with the programmatic-stream-
// serialization attribute, call pdl_wait() before touching anything their predecessor wrote and pdl_trigger() right
// after it, so that the successor's launch latency, block scheduling and (for the conv kernels) barrier / TMEM /
// descriptor set-up overlap the predecessor's tail.  Opt-in with LDM_PDL=1 (without the attribute the device calls are
// no-ops); see api.cu for the measurement that keeps it off by default.
extern int g_ldm_pdl;
template <typename... KArgs, typename... Args>
static inline cudaError_t ldm_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_ldm_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t align_up64(int64_t a, int64_t b) { return ceil_div64(a, b) * b; }

enum { LDM_DT_F32 = 0, LDM_DT_BF16 = 1 };
static inline int dtype_size(int dtype) { return dtype == LDM_DT_BF16 ? 2 : 4; }

// ---------------------------------------------------------------- device side
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// The explicit early trigger let successors' CTAs take SM slots that the remaining waves of a multi-wave predecessor
// still needed (measured slower); without it the successor is released when every predecessor CTA has exited, which
// still hides its launch latency and prologue behind the predecessor's drain.
#ifdef LDM_PDL_EARLY_TRIGGER
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_trigger() {}
#endif

template <typename T>
struct VecTraits;
template <>
struct VecTraits<float> {
  static constexpr int N = 4;  // elements per 16-byte chunk
};
template <>
struct VecTraits<bf16> {
  static constexpr int N = 8;
};

// 16-byte chunk load/store with conversion to/from fp32 registers.
__device__ __forceinline__ void load_chunk(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load_chunk(const bf16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store_chunk(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store_chunk(bf16* p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}

// the same store with an L2 evict_last policy: the tensor is read back by the very next kernels (GroupNorm statistics and
// apply) and, at 67 MB, fits the 126 MB L2
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void store_chunk_keep(bf16* p, const float (&v)[8], uint64_t pol) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(t.x), "r"(t.y), "r"(t.z), "r"(t.w),
               "l"(pol) : "memory");
}

__device__ __forceinline__ float to_float(float x) { return x; }
__device__ __forceinline__ float to_float(bf16 x) { return __bfloat162float(x); }
template <typename T>
__device__ __forceinline__ T from_float(float x);
template <>
__device__ __forceinline__ float from_float<float>(float x) { return x; }
template <>
__device__ __forceinline__ bf16 from_float<bf16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// accurate variant for the fp32 parity path
__device__ __forceinline__ float silu_acc(float x) { return x / (1.0f + expf(-x)); }

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based so results do not depend on launch shape
struct Philox4 {
  uint32_t x, y, z, w;
};
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}
// four N(0,1) draws from one Philox block (Box-Muller on two uniform pairs)
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                               uint32_t c3, float (&z)[4]) {
  Philox4 r = philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float S = 5.9604644775390625e-08f;  // 2^-24
  float u0 = ((r.x >> 8) + 0.5f) * S, u1 = ((r.y >> 8) + 0.5f) * S;
  float u2 = ((r.z >> 8) + 0.5f) * S, u3 = ((r.w >> 8) + 0.5f) * S;
  float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u2));
  float s0, c0f, s1, c1f;
  sincospif(2.0f * u1, &s0, &c0f);
  sincospif(2.0f * u3, &s1, &c1f);
  z[0] = r0 * c0f; z[1] = r0 * s0; z[2] = r1 * c1f; z[3] = r1 * s1;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif  // __CUDACC__
