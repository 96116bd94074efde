// GroupNorm (+SiLU) (+residual) on NHWC tensors -- replaces F.group_norm / nn.SiLU / Residual add of
// src/UNet.py:52-58 (Block: GN(8,C) -> SiLU), :106 (PreNorm GN(1,C)), :147 (to_out GN(1,C)), :20 (x + fn(x)).
//
// Two HBM-bound kernels (bytes: stats reads x once; apply reads x once (mostly L2 hits) and writes y once):
//   gn_stats : per (sample, pixel-slab) CTA, fully coalesced 16-byte loads; every thread owns a fixed
//              channel chunk (so a fixed group) and keeps a running (count, mean, M2) merged with
//              Chan's parallel update -- exact two-pass-quality variance from a single read, no atomics,
//              deterministic reduction order.
//   gn_apply : merges the slab partials per group, then y = [silu](x*a + b) [+ res] element-wise.
#include "kernels.h"

#define GN_MAX_SPLITS 32
#define GN_MAX_GROUPS 32

struct Moments {
  float n, mean, m2;
};
__device__ __forceinline__ void chan_merge(Moments& a, const Moments& b) {
  if (b.n == 0.f) return;
  if (a.n == 0.f) { a = b; return; }
  float n = a.n + b.n;
  float d = b.mean - a.mean;
  float rb = b.n / n;
  a.mean = fmaf(d, rb, a.mean);
  a.m2 = a.m2 + b.m2 + d * d * a.n * rb;
  a.n = n;
}

template <typename T>
__global__ void gn_stats_kernel(const T* __restrict__ x, int ldx, float* __restrict__ part, int HW, int C, int G,
                                int splits, int k) {
  constexpr int V = VecTraits<T>::N;
  extern __shared__ float sm[];  // [blockDim][3]
  const int cpp = C / V;
  const int ci = threadIdx.x % cpp, pl = threadIdx.x / cpp;  // threads with pl >= k are padding lanes
  const int n = blockIdx.y, s = blockIdx.x;
  const int p_begin = (int)((int64_t)HW * s / splits), p_end = (int)((int64_t)HW * (s + 1) / splits);
  const T* base = x + (int64_t)n * HW * ldx + ci * V;
  Moments run{0.f, 0.f, 0.f};
  for (int p = p_begin + pl; p < p_end && pl < k; p += k) {
    float v[V];
    load_chunk(base + (int64_t)p * ldx, v);
    float cs = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) cs += v[i];
    Moments c;
    c.n = (float)V;
    c.mean = cs * (1.0f / V);
    c.m2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) { float d = v[i] - c.mean; c.m2 = fmaf(d, d, c.m2); }
    chan_merge(run, c);
  }
  sm[threadIdx.x * 3 + 0] = run.n;
  sm[threadIdx.x * 3 + 1] = run.mean;
  sm[threadIdx.x * 3 + 2] = run.m2;
  __syncthreads();
  const int cpg = cpp / G;  // chunks per group per pixel
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5;
  const int members = k * cpg;
  for (int g = warp; g < G; g += nwarps) {
    Moments acc{0.f, 0.f, 0.f};
    for (int idx = lane; idx < members; idx += 32) {
      int pi = idx / cpg, j = idx % cpg;
      int tt = pi * cpp + g * cpg + j;
      Moments m{sm[tt * 3], sm[tt * 3 + 1], sm[tt * 3 + 2]};
      chan_merge(acc, m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Moments m;
      m.n = __shfl_xor_sync(0xffffffffu, acc.n, o);
      m.mean = __shfl_xor_sync(0xffffffffu, acc.mean, o);
      m.m2 = __shfl_xor_sync(0xffffffffu, acc.m2, o);
      // merge in a lane-symmetric order so both partners compute the same value
      Moments lo = (lane & o) ? m : acc, hi = (lane & o) ? acc : m;
      chan_merge(lo, hi);
      acc = lo;
    }
    if (lane == 0) {
      float* o = part + (((int64_t)n * splits + s) * G + g) * 3;
      o[0] = acc.n; o[1] = acc.mean; o[2] = acc.m2;
    }
  }
}

template <typename T>
__global__ void gn_apply_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy,
                                const T* __restrict__ res, int ldres, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ part, int HW, int C,
                                int G, int splits, float eps, int silu, int pix_per_block, int k) {
  constexpr int V = VecTraits<T>::N;
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
  const int n = blockIdx.y;
  if (threadIdx.x < G) {
    Moments acc{0.f, 0.f, 0.f};
    for (int s = 0; s < splits; ++s) {
      const float* p = part + (((int64_t)n * splits + s) * G + threadIdx.x) * 3;
      Moments m{p[0], p[1], p[2]};
      chan_merge(acc, m);
    }
    s_mean[threadIdx.x] = acc.mean;
    s_rstd[threadIdx.x] = 1.0f / sqrtf(acc.m2 / acc.n + eps);  // biased variance, as F.group_norm
  }
  __syncthreads();
  const int cpp = C / V;
  const int ci = threadIdx.x % cpp, pl = threadIdx.x / cpp;
  if (pl >= k) return;  // padding lanes (blockDim is rounded up to a warp multiple)
  const int g = ci / (cpp / G);
  float a[V], b[V];
  {
    const float mean = s_mean[g], rstd = s_rstd[g];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float ga = gamma[ci * V + i], be = beta[ci * V + i];
      a[i] = ga * rstd;
      b[i] = be - mean * a[i];
    }
  }
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(p0 + pix_per_block, HW);
  const T* xb = x + (int64_t)n * HW * ldx + ci * V;
  T* yb = y + (int64_t)n * HW * ldy + ci * V;
  const T* rb = res ? res + (int64_t)n * HW * ldres + ci * V : nullptr;
  for (int p = p0 + pl; p < p1; p += k) {
    float v[V];
    load_chunk(xb + (int64_t)p * ldx, v);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float o = fmaf(v[i], a[i], b[i]);
      if (silu) o = sizeof(T) == 4 ? silu_acc(o) : silu_f(o);
      v[i] = o;
    }
    if (rb) {
      float r[V];
      load_chunk(rb + (int64_t)p * ldres, r);
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] += r[i];
    }
    store_chunk(yb + (int64_t)p * ldy, v);
  }
}

int64_t k_group_norm_ws_bytes(int batch, int groups) {
  return (int64_t)batch * GN_MAX_SPLITS * groups * 3 * sizeof(float);
}

template <typename T>
static int gn_launch(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                     const float* beta, int batch, int hw, int channels, int groups, float eps, int silu,
                     void* workspace, cudaStream_t st) {
  constexpr int V = VecTraits<T>::N;
  const int cpp = channels / V;
  int k = (256 + cpp - 1) / cpp;
  if (k < 1) k = 1;
  if (k > hw) k = hw;
  const int threads = (cpp * k + 31) / 32 * 32;
  // Slab count depends only on the per-sample geometry (never on the batch), so a sample's reduction
  // order -- and therefore its bits -- is the same however the batch is sharded across GPUs.
  int splits = hw / (k * 4);
  if (splits > 8) splits = 8;
  if (splits > GN_MAX_SPLITS) splits = GN_MAX_SPLITS;
  if (splits < 1) splits = 1;
  (void)batch;
  float* part = (float*)workspace;
  gn_stats_kernel<T><<<dim3(splits, batch), threads, threads * 3 * sizeof(float), st>>>(
      (const T*)x, ldx, part, hw, channels, groups, splits, k);
  LDM_LAUNCHED("gn_stats");
  int ppb = k * 8;
  gn_apply_kernel<T><<<dim3((hw + ppb - 1) / ppb, batch), threads, 0, st>>>(
      (const T*)x, ldx, (T*)y, ldy, (const T*)res, ldres, gamma, beta, part, hw, channels, groups, splits, eps,
      silu, ppb, k);
  LDM_LAUNCHED("gn_apply");
  return 0;
}

int k_group_norm(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                 const float* beta, int batch, int hw, int channels, int groups, float eps, int silu, int dtype,
                 void* workspace, cudaStream_t st) {
  const int V = dtype == LDM_DT_BF16 ? 8 : 4;
  LDM_REQUIRE(groups >= 1 && groups <= GN_MAX_GROUPS, "group_norm: groups=%d unsupported", groups);
  LDM_REQUIRE(channels % groups == 0 && (channels / groups) % V == 0,
              "group_norm: channels/groups (%d/%d) must be a multiple of %d", channels, groups, V);
  LDM_REQUIRE(ldx % V == 0 && ldy % V == 0 && (res == nullptr || ldres % V == 0), "group_norm: unaligned stride");
  LDM_REQUIRE(channels / V <= 1024, "group_norm: too many channels (%d)", channels);
  LDM_REQUIRE(workspace != nullptr, "group_norm: workspace required");
  if (batch == 0 || hw == 0) return 0;
  LDM_REQUIRE(batch <= 65535, "group_norm: batch %d exceeds grid limit", batch);
  if (dtype == LDM_DT_BF16)
    return gn_launch<bf16>(x, ldx, y, ldy, res, ldres, gamma, beta, batch, hw, channels, groups, eps, silu, workspace, st);
  return gn_launch<float>(x, ldx, y, ldy, res, ldres, gamma, beta, batch, hw, channels, groups, eps, silu, workspace, st);
}
