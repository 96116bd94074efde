// GroupNorm (+SiLU) (+residual) on NHWC tensors -- replaces F.group_norm / nn.SiLU / Residual add of
// src/UNet.py:52-58 (Block: GN(8,C) -> SiLU), :106 (PreNorm GN(1,C)), :147 (to_out GN(1,C)), :20 (x + fn(x)).
//
// Two implementations:
//  (A) bf16 product path: two streaming kernels with full occupancy.  gn_stats reads x once from HBM and leaves
//      per-(sample, slab, group) shifted sums; gn_apply re-reads x -- an L2 hit for every tensor of the reference
//      UNet but one, the producer->consumer distance being a single 10-20 us kernel -- and writes y.  HBM traffic is
//      the ideal one read + one write, and unlike a one-CTA-per-sample kernel nothing ever waits on a reduction.
//      Sums are taken about a per-group pivot K = x[n][first pixel][first channel of the group] so that
//      var = E[(x-K)^2] - (E[x-K])^2 does not cancel; partials are combined in a fixed order (deterministic,
//      independent of batch and GPU count).
//  (B) fp32 parity path: ONE kernel, one CTA per sample, exact two-pass statistics:
//   * every thread owns a fixed 16-byte channel chunk (so a fixed group and fixed affine coefficients) and walks
//     the pixels with stride blockDim/chunks_per_pixel; all of a thread's loads are issued up front and the
//     values stay in registers (NJ <= 8 chunks per thread) between the statistics and the apply step.
//     Samples too large for that (NJ > 8) use the same code with a second read, which hits L2.
//   * statistics are exact two-pass (mean, then sum of squared deviations) in fp32, reduced per group through
//     shared memory in a fixed order: no atomics, so a sample's bits do not depend on the batch, the launch
//     shape or the GPU count.
//   * y = [silu](x * a + b) [+ res];  SiLU via one MUFU (tanh.approx) on the bf16 path.
#include <stdlib.h>

#include "kernels.h"

#define GN_MAX_GROUPS 32
#define GN_MAX_THREADS 1024
#define GN_HOLD 8  // chunks a thread keeps in registers

namespace {

// raw 16-byte chunks stay packed in registers (4 regs each) and are unpacked to fp32 on use
__device__ __forceinline__ uint4 load_raw(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void unpack(const uint4& t, float (&v)[4]) {
  v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
}
__device__ __forceinline__ void unpack(const uint4& t, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

__device__ __forceinline__ float silu_fast(float x) {
  // x * sigmoid(x) = h * tanh(h) + h with h = x/2: one MUFU op
  float h = 0.5f * x, t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// Sum `v` over the threads of each group; result for this thread's group is returned to every thread.
// part: smem [blockDim], gsum: smem [G].  Deterministic order.
__device__ __forceinline__ float group_reduce(float v, float* part, float* gsum, int G, int cpp, int cpg, int my_group) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  if (G == 1) {
    v = warp_sum(v);
    if (lane == 0) part[warp] = v;
    __syncthreads();
    if (warp == 0) {
      float s = 0.f;
      for (int i = lane; i < nwarps; i += 32) s += part[i];
      s = warp_sum(s);
      if (lane == 0) gsum[0] = s;
    }
    __syncthreads();
    return gsum[0];
  }
  part[tid] = v;
  __syncthreads();
  const int members = (blockDim.x / cpp) * cpg;
  for (int g = warp; g < G; g += nwarps) {
    float s = 0.f;
    for (int idx = lane; idx < members; idx += 32) s += part[(idx / cpg) * cpp + g * cpg + idx % cpg];
    s = warp_sum(s);
    if (lane == 0) gsum[g] = s;
  }
  __syncthreads();
  return gsum[my_group];
}

template <typename T, int NJ, bool RV>  // NJ > 0: hold NJ chunks per thread in registers; NJ == 0: re-read mode
__global__ void __launch_bounds__(GN_MAX_THREADS)
gn_fused_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, const T* __restrict__ res, int ldres,
                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ rowvec,
                int ld_rowvec, int HW, int C, int G, float eps, int silu) {
  constexpr int V = VecTraits<T>::N;
  constexpr int NH = NJ > 0 ? NJ : 1;
  __shared__ float part[GN_MAX_THREADS];
  __shared__ float gsum[GN_MAX_GROUPS];
  const int n = blockIdx.x;
  const int cpp = C / V;               // chunks per pixel
  const int cpg = cpp / G;             // chunks per group per pixel
  const int ci = threadIdx.x % cpp;    // this thread's channel chunk
  const int pl = threadIdx.x / cpp;    // first pixel
  const int ppi = blockDim.x / cpp;    // pixels per iteration
  const int g = ci / cpg;
  const T* xb = x + (int64_t)n * HW * ldx + ci * V;
  uint4 raw[NH];
  float w[V];
  // optional per-sample, per-channel vector added to x before the statistics (the ResNetBlock time embedding,
  // src/UNet.py:88-93: h = h + mlp_t(t)[:, :, None, None], then block2's GroupNorm)
  float rv[V];
#pragma unroll
  for (int i = 0; i < V; ++i) rv[i] = RV ? rowvec[(int64_t)n * ld_rowvec + ci * V + i] : 0.f;
  // ---- pass 1: load + sum
  float s = 0.f;
  if (NJ > 0) {
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      const int p = pl + j * ppi;
      raw[j] = p < HW ? load_raw(xb + (int64_t)p * ldx) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      if (pl + j * ppi < HW) {
        unpack(raw[j], w);
#pragma unroll
        for (int i = 0; i < V; ++i) s += w[i] + rv[i];
      }
    }
  } else {
    for (int p = pl; p < HW; p += ppi) {
      unpack(load_raw(xb + (int64_t)p * ldx), w);
#pragma unroll
      for (int i = 0; i < V; ++i) s += w[i] + rv[i];
    }
  }
  const float inv_n = 1.0f / ((float)HW * (float)(cpg * V));
  const float mean = group_reduce(s, part, gsum, G, cpp, cpg, g) * inv_n;
  // ---- pass 2: sum of squared deviations
  float q = 0.f;
  if (NJ > 0) {
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      if (pl + j * ppi < HW) {
        unpack(raw[j], w);
#pragma unroll
        for (int i = 0; i < V; ++i) { float d = w[i] + rv[i] - mean; q = fmaf(d, d, q); }
      }
    }
  } else {
    for (int p = pl; p < HW; p += ppi) {
      unpack(load_raw(xb + (int64_t)p * ldx), w);
#pragma unroll
      for (int i = 0; i < V; ++i) { float d = w[i] + rv[i] - mean; q = fmaf(d, d, q); }
    }
  }
  __syncthreads();  // part/gsum reuse
  const float var = group_reduce(q, part, gsum, G, cpp, cpg, g) * inv_n;  // biased, as F.group_norm
  const float rstd = 1.0f / sqrtf(var + eps);
  // ---- apply
  float a[V], b[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float ga = gamma[ci * V + i], be = beta[ci * V + i];
    a[i] = ga * rstd;
    b[i] = be + (rv[i] - mean) * a[i];
  }
  T* yb = y + (int64_t)n * HW * ldy + ci * V;
  const T* rb = res ? res + (int64_t)n * HW * ldres + ci * V : nullptr;
  auto emit = [&](const uint4& rw, int p) {
    unpack(rw, w);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float o = fmaf(w[i], a[i], b[i]);
      if (silu) o = sizeof(T) == 4 ? silu_acc(o) : silu_fast(o);
      w[i] = o;
    }
    if (rb) {
      float r[V];
      load_chunk(rb + (int64_t)p * ldres, r);
#pragma unroll
      for (int i = 0; i < V; ++i) w[i] += r[i];
    }
    store_chunk(yb + (int64_t)p * ldy, w);
  };
  if (NJ > 0) {
#pragma unroll
    for (int j = 0; j < NH; ++j) {
      const int p = pl + j * ppi;
      if (p < HW) emit(raw[j], p);
    }
  } else {
    for (int p = pl; p < HW; p += ppi) emit(load_raw(xb + (int64_t)p * ldx), p);
  }
}

// ------------------------------------------------------------------ max-pool 2x2 + GroupNorm (+SiLU), bf16
// Encoder level transition (src/UNet.py:183,193 MaxPool2d(2,2), then the next ResNetBlock's block1.norm + SiLU, :52-58):
// one CTA per sample reads the 2x2 windows ONCE, keeps the pooled sample in registers (NJ chunks per thread), writes the
// pooled tensor (the block's residual / shortcut input) and, after exact two-pass statistics, the normalised one.
// Replaces three launches (maxpool2, gn_stats, gn_apply) and two re-reads of the pooled tensor.
template <int NJ>
__global__ void __launch_bounds__(256)
pool_gn_kernel(const bf16* __restrict__ x, int ldx, int W, bf16* __restrict__ pool, int ldp, bf16* __restrict__ y, int ldy,
               const float* __restrict__ gamma, const float* __restrict__ beta, int HW2, int C, int G, float eps, int silu) {
  constexpr int V = 8;
  __shared__ float part[256];
  __shared__ float gsum[GN_MAX_GROUPS];
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.x;
  const int cpp = C / V, cpg = cpp / G;
  const int ci = threadIdx.x % cpp, pl = threadIdx.x / cpp, ppi = blockDim.x / cpp;
  const int g = ci / cpg;
  const int W2 = W / 2;
  const bf16* xs = x + (int64_t)n * (4 * HW2) * ldx + ci * V;     // source image: (2 H2) x W pixels
  uint4 raw[NJ];
  float w[V];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int p = pl + j * ppi;
    raw[j] = make_uint4(0u, 0u, 0u, 0u);
    if (p < HW2) {
      const int py = p / W2, px = p - py * W2;
      const bf16* s0 = xs + ((int64_t)(2 * py) * W + 2 * px) * ldx;
      const uint4 a = load_raw(s0), b = load_raw(s0 + ldx), c = load_raw(s0 + (int64_t)W * ldx), d = load_raw(s0 + (int64_t)(W + 1) * ldx);
      const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&b);
      const __nv_bfloat162* hc = reinterpret_cast<const __nv_bfloat162*>(&c);
      const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&d);
      __nv_bfloat162* hr = reinterpret_cast<__nv_bfloat162*>(&raw[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i) hr[i] = __hmax2(__hmax2(ha[i], hb[i]), __hmax2(hc[i], hd[i]));
      *reinterpret_cast<uint4*>(pool + ((int64_t)n * HW2 + p) * ldp + ci * V) = raw[j];
      unpack(raw[j], w);
#pragma unroll
      for (int i = 0; i < V; ++i) s += w[i];
    }
  }
  const float inv_n = 1.0f / ((float)HW2 * (float)(cpg * V));
  const float mean = group_reduce(s, part, gsum, G, cpp, cpg, g) * inv_n;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    if (pl + j * ppi < HW2) {
      unpack(raw[j], w);
#pragma unroll
      for (int i = 0; i < V; ++i) { const float d = w[i] - mean; q = fmaf(d, d, q); }
    }
  }
  __syncthreads();
  const float var = group_reduce(q, part, gsum, G, cpp, cpg, g) * inv_n;
  const float rstd = 1.0f / sqrtf(var + eps);
  float a[V], b[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    a[i] = gamma[ci * V + i] * rstd;
    b[i] = beta[ci * V + i] - mean * a[i];
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int p = pl + j * ppi;
    if (p < HW2) {
      unpack(raw[j], w);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float o = fmaf(w[i], a[i], b[i]);
        if (silu) o = silu_fast(o);
        w[i] = o;
      }
      store_chunk(y + ((int64_t)n * HW2 + p) * ldy + ci * V, w);
    }
  }
}

// ------------------------------------------------------------------ (A) streaming stats + apply, bf16
constexpr int GS_THREADS = 256;
constexpr int GS_MAX_SPLITS = 16;

// grid (splits, batch); thread -> fixed channel chunk; writes part[n][split][g] = {sum(x-K), sum((x-K)^2)}
template <bool RV>
__global__ void __launch_bounds__(GS_THREADS)
gn_stats_kernel(const bf16* __restrict__ x, int ldx, const float* __restrict__ rowvec, int ld_rowvec,
                float2* __restrict__ part, int HW, int C, int G, int pix_per_split, int x_mod) {
  constexpr int V = 8;
  __shared__ float part_s[GS_THREADS], part_q[GS_THREADS];
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.y, split = blockIdx.x;
  const int cpp = C / V, cpg = cpp / G;
  const int ci = threadIdx.x % cpp, pl = threadIdx.x / cpp, ppi = blockDim.x / cpp;
  const int g = ci / cpg;
  // x_mod > 0: sample n reads image n % x_mod (the sampler's cond / uncond halves share everything up to the first
  // time-embedding add, so that prefix is computed once)
  const bf16* xs = x + (int64_t)(x_mod > 0 ? n % x_mod : n) * HW * ldx;
  float rv[V];
#pragma unroll
  for (int i = 0; i < V; ++i) rv[i] = RV ? rowvec[(int64_t)n * ld_rowvec + ci * V + i] : 0.f;
  // pivot of this thread's group (same for every CTA of the sample)
  const int c_first = g * cpg * V;
  const float K = __bfloat162float(xs[c_first]) + (RV ? rowvec[(int64_t)n * ld_rowvec + c_first] : 0.f);
  const int p0 = split * pix_per_split, p1 = min(p0 + pix_per_split, HW);
  float s = 0.f, q = 0.f;
  const bf16* xb = xs + ci * V;
  int p = p0 + pl;
  // 4 independent 16-byte loads in flight per thread
  for (; p + 3 * ppi < p1; p += 4 * ppi) {
    uint4 r0 = load_raw(xb + (int64_t)p * ldx), r1 = load_raw(xb + (int64_t)(p + ppi) * ldx);
    uint4 r2 = load_raw(xb + (int64_t)(p + 2 * ppi) * ldx), r3 = load_raw(xb + (int64_t)(p + 3 * ppi) * ldx);
    float w[V];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      unpack(u == 0 ? r0 : (u == 1 ? r1 : (u == 2 ? r2 : r3)), w);
#pragma unroll
      for (int i = 0; i < V; ++i) { const float d = w[i] + rv[i] - K; s += d; q = fmaf(d, d, q); }
    }
  }
  for (; p < p1; p += ppi) {
    float w[V];
    unpack(load_raw(xb + (int64_t)p * ldx), w);
#pragma unroll
    for (int i = 0; i < V; ++i) { const float d = w[i] + rv[i] - K; s += d; q = fmaf(d, d, q); }
  }
  part_s[threadIdx.x] = s;
  part_q[threadIdx.x] = q;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int members = (blockDim.x / cpp) * cpg;
  for (int gg = warp; gg < G; gg += nwarps) {
    float a = 0.f, b = 0.f;
    for (int idx = lane; idx < members; idx += 32) {
      const int t = (idx / cpg) * cpp + gg * cpg + idx % cpg;
      a += part_s[t]; b += part_q[t];
    }
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) part[((int64_t)n * gridDim.x + split) * G + gg] = make_float2(a, b);
  }
}

// grid (ceil(HW / pix_per_block), batch)
template <bool RV>
__global__ void __launch_bounds__(GS_THREADS)
gn_apply_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ y, int ldy, const bf16* __restrict__ res,
                int ldres, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ rowvec, int ld_rowvec, const float2* __restrict__ part, int splits, int HW,
                int C, int G, float eps, int silu, int pix_per_block, int x_mod) {
  constexpr int V = 8;
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.y;
  const int cpp = C / V, cpg = cpp / G;
  const bf16* xs = x + (int64_t)(x_mod > 0 ? n % x_mod : n) * HW * ldx;
  if (threadIdx.x < G && splits < 0) {
    // un-pivoted sums {S, Q} left by the producing convolution's epilogue (conv_epilogue.cuh, mode 1):
    // part[(n * G + g) * nslots + slot], nslots = -splits
    // (x_mod > 0: sample n is variant n / x_mod of image n % x_mod -- part[((image * G + g) * nvar + variant) * nslots + slot])
    const int gg = threadIdx.x, ns = -splits;
    const int nv = x_mod > 0 ? gridDim.y / x_mod : 1, img = x_mod > 0 ? n % x_mod : n, kk = x_mod > 0 ? n / x_mod : 0;
    double a = 0.0, b = 0.0;
    for (int sp = 0; sp < ns; ++sp) {
      const float2 v = part[(((int64_t)img * G + gg) * nv + kk) * ns + sp];
      a += (double)v.x; b += (double)v.y;
    }
    const double cnt = (double)HW * (double)(cpg * V);
    const double m1 = a / cnt;
    double var = b / cnt - m1 * m1;
    if (var < 0.0) var = 0.0;
    s_mean[gg] = (float)m1;
    s_rstd[gg] = (float)(1.0 / sqrt(var + (double)eps));
  } else if (threadIdx.x < G) {
    const int gg = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int sp = 0; sp < splits; ++sp) {  // fixed order
      const float2 v = part[((int64_t)n * splits + sp) * G + gg];
      a += v.x; b += v.y;
    }
    const int c_first = gg * cpg * V;
    const float K = __bfloat162float(xs[c_first]) + (RV ? rowvec[(int64_t)n * ld_rowvec + c_first] : 0.f);
    const float inv_n = 1.0f / ((float)HW * (float)(cpg * V));
    const float m1 = a * inv_n;                       // E[x - K]
    const float var = fmaxf(b * inv_n - m1 * m1, 0.f);  // biased, as F.group_norm
    s_mean[gg] = K + m1;
    s_rstd[gg] = 1.0f / sqrtf(var + eps);
  }
  __syncthreads();
  const int ci = threadIdx.x % cpp, pl = threadIdx.x / cpp, ppi = blockDim.x / cpp;
  if (pl >= ppi) return;
  const int g = ci / cpg;
  float a[V], b[V];
  {
    const float mean = s_mean[g], rstd = s_rstd[g];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float ga = gamma[ci * V + i], be = beta[ci * V + i];
      const float rvi = RV ? rowvec[(int64_t)n * ld_rowvec + ci * V + i] : 0.f;
      a[i] = ga * rstd;
      b[i] = be + (rvi - mean) * a[i];
    }
  }
  const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, HW);
  const bf16* xb = xs + ci * V;
  bf16* yb = y + (int64_t)n * HW * ldy + ci * V;
  const bf16* rb = res ? res + (int64_t)n * HW * ldres + ci * V : nullptr;
  auto emit = [&](const uint4& rw, const uint4& rr, int p) {
    float w[V];
    unpack(rw, w);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float o = fmaf(w[i], a[i], b[i]);
      if (silu) o = silu_fast(o);
      w[i] = o;
    }
    if (rb) {
      float r[V];
      unpack(rr, r);
#pragma unroll
      for (int i = 0; i < V; ++i) w[i] += r[i];
    }
    store_chunk(yb + (int64_t)p * ldy, w);
  };
  int p = p0 + pl;
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  for (; p + 3 * ppi < p1; p += 4 * ppi) {
    uint4 r0 = load_raw(xb + (int64_t)p * ldx), r1 = load_raw(xb + (int64_t)(p + ppi) * ldx);
    uint4 r2 = load_raw(xb + (int64_t)(p + 2 * ppi) * ldx), r3 = load_raw(xb + (int64_t)(p + 3 * ppi) * ldx);
    uint4 q0 = z4, q1 = z4, q2 = z4, q3 = z4;
    if (rb) {
      q0 = load_raw(rb + (int64_t)p * ldres); q1 = load_raw(rb + (int64_t)(p + ppi) * ldres);
      q2 = load_raw(rb + (int64_t)(p + 2 * ppi) * ldres); q3 = load_raw(rb + (int64_t)(p + 3 * ppi) * ldres);
    }
    emit(r0, q0, p); emit(r1, q1, p + ppi); emit(r2, q2, p + 2 * ppi); emit(r3, q3, p + 3 * ppi);
  }
  for (; p < p1; p += ppi) emit(load_raw(xb + (int64_t)p * ldx), rb ? load_raw(rb + (int64_t)p * ldres) : z4, p);
}

// ------------------------------------------------------------------ (B) streaming backward, bf16
// y = [silu](gamma * xh + beta), xh = (x + rowvec - mean) * rstd.  Three multi-CTA kernels instead of one CTA per sample:
//   gn_stats_kernel          part  [n][split][g] = pivoted sums of x            (or the forward's, if the caller kept them)
//   gn_bwd_sums_kernel       part2 [n][split][g] = {sum dyh, sum dyh*xh};  dgamma, dbeta accumulated (atomics)
//   gn_bwd_apply_kernel      dx = rstd * (dyh - mean(dyh) - xh * mean(dyh*xh));  drowvec[n][c] = sum_p dx  (atomics)
__device__ __forceinline__ float silu_grad(float z) {
  const float sg = 1.0f / (1.0f + __expf(-z));
  return sg * (1.0f + z * (1.0f - sg));
}

// mean / rstd of every group of sample n from the pivoted partial sums (same arithmetic as gn_apply_kernel)
template <bool RV>
__device__ __forceinline__ void gn_group_stats(const bf16* xs, const float* rowvec, int ld_rowvec, int n, const float2* part,
                                               int splits, int HW, int cpg, int G, float eps, float* s_mean, float* s_rstd) {
  constexpr int V = 8;
  if (threadIdx.x < G) {
    const int gg = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int sp = 0; sp < splits; ++sp) {
      const float2 v = part[((int64_t)n * splits + sp) * G + gg];
      a += v.x; b += v.y;
    }
    const int c_first = gg * cpg * V;
    const float K = __bfloat162float(xs[c_first]) + (RV ? rowvec[(int64_t)n * ld_rowvec + c_first] : 0.f);
    const float inv_n = 1.0f / ((float)HW * (float)(cpg * V));
    const float m1 = a * inv_n;
    const float var = fmaxf(b * inv_n - m1 * m1, 0.f);
    s_mean[gg] = K + m1;
    s_rstd[gg] = 1.0f / sqrtf(var + eps);
  }
}

// grid (splits, batch)
template <bool RV>
__global__ void __launch_bounds__(GS_THREADS)
gn_bwd_sums_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ dy, int lddy, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ rowvec, int ld_rowvec,
                   const float2* __restrict__ part, float2* __restrict__ part2, float* __restrict__ chan_part,
                   float* __restrict__ drowvec, int ld_drowvec, int HW, int C, int G, float eps, int silu, int pix_per_split) {
  constexpr int V = 8;
  __shared__ float red_a[GS_THREADS], red_b[GS_THREADS];
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
  const int n = blockIdx.y, split = blockIdx.x, splits = gridDim.x;
  const int cpp = C / V, cpg = cpp / G;
  const bf16* xs = x + (int64_t)n * HW * ldx;
  gn_group_stats<RV>(xs, rowvec, ld_rowvec, n, part, splits, HW, cpg, G, eps, s_mean, s_rstd);
  if (drowvec && split == 0)   // the apply kernel accumulates into it
    for (int c = threadIdx.x; c < C; c += blockDim.x) drowvec[(int64_t)n * ld_drowvec + c] = 0.f;
  __syncthreads();
  const int ci = threadIdx.x % cpp, pl = threadIdx.x / cpp, ppi = blockDim.x / cpp;
  const int g = ci / cpg;
  const float mean = s_mean[g], rstd = s_rstd[g];
  float ga[V], be[V], sh[V];   // xh = x * rstd + sh
#pragma unroll
  for (int i = 0; i < V; ++i) {
    ga[i] = gamma[ci * V + i];
    be[i] = beta[ci * V + i];
    sh[i] = ((RV ? rowvec[(int64_t)n * ld_rowvec + ci * V + i] : 0.f) - mean) * rstd;
  }
  float s1 = 0.f, s2 = 0.f, dg[V], db[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { dg[i] = 0.f; db[i] = 0.f; }
  const int p0 = split * pix_per_split, p1 = min(p0 + pix_per_split, HW);
  const bf16* xb = xs + ci * V;
  const bf16* dyb = dy + (int64_t)n * HW * lddy + ci * V;
  auto acc = [&](const uint4& rx, const uint4& rd) {
    float w[V], d[V];
    unpack(rx, w);
    unpack(rd, d);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xh = fmaf(w[i], rstd, sh[i]);
      float dz = d[i];
      if (silu) dz *= silu_grad(fmaf(xh, ga[i], be[i]));
      dg[i] = fmaf(dz, xh, dg[i]);
      db[i] += dz;
      const float dyh = dz * ga[i];
      s1 += dyh;
      s2 = fmaf(dyh, xh, s2);
    }
  };
  int p = p0 + pl;
  for (; p + ppi < p1; p += 2 * ppi) {
    const uint4 x0 = load_raw(xb + (int64_t)p * ldx), x1 = load_raw(xb + (int64_t)(p + ppi) * ldx);
    const uint4 d0 = load_raw(dyb + (int64_t)p * lddy), d1 = load_raw(dyb + (int64_t)(p + ppi) * lddy);
    acc(x0, d0); acc(x1, d1);
  }
  for (; p < p1; p += ppi) acc(load_raw(xb + (int64_t)p * ldx), load_raw(dyb + (int64_t)p * lddy));
  // group sums of this CTA's slab
  red_a[threadIdx.x] = s1;
  red_b[threadIdx.x] = s2;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int members = ppi * cpg;
  for (int gg = warp; gg < G; gg += nwarps) {
    float a = 0.f, b = 0.f;
    for (int idx = lane; idx < members; idx += 32) {
      const int t = (idx / cpg) * cpp + gg * cpg + idx % cpg;
      a += red_a[t]; b += red_b[t];
    }
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) part2[((int64_t)n * splits + split) * G + gg] = make_float2(a, b);
  }
  // per-channel parameter gradients: sum over the pixel lanes of this CTA and write the CTA's row of partials
  // chan_part[(n*splits + split)][2C] = {dgamma | dbeta}; gn_bwd_param_kernel adds the rows up (thousands of CTAs doing
  // atomics on the same 2C floats serialise in L2: measured 3x the kernel's streaming time)
  float* row = chan_part + ((int64_t)n * splits + split) * 2 * C;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    __syncthreads();
    red_a[threadIdx.x] = dg[i];
    red_b[threadIdx.x] = db[i];
    __syncthreads();
    if (pl == 0) {
      float a = 0.f, b = 0.f;
      for (int k = 0; k < ppi; ++k) { a += red_a[k * cpp + ci]; b += red_b[k * cpp + ci]; }
      row[ci * V + i] = a;
      row[C + ci * V + i] = b;
    }
  }
}

// dgamma[c] += sum_rows chan_part[r][c], dbeta[c] += sum_rows chan_part[r][C + c].  grid (ceil(2C / 32)), 256 threads =
// 8 row lanes x 32 columns; fixed summation order (bit-reproducible)
__global__ void __launch_bounds__(256)
gn_bwd_param_kernel(const float* __restrict__ chan_part, int rows, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  float a = 0.f;
  if (col < 2 * C)
    for (int r = rl; r < rows; r += 8) a += chan_part[(int64_t)r * 2 * C + col];
  red[rl][threadIdx.x & 31] = a;
  __syncthreads();
  if (rl == 0 && col < 2 * C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    if (col < C) dgamma[col] += t; else dbeta[col - C] += t;
  }
}

// grid (ceil(HW / pix_per_block), batch)
template <bool RV>
__global__ void __launch_bounds__(GS_THREADS)
gn_bwd_apply_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ dy, int lddy, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ rowvec, int ld_rowvec,
                    const float2* __restrict__ part, const float2* __restrict__ part2, int splits, bf16* __restrict__ dx,
                    int lddx, float* __restrict__ drowvec, int ld_drowvec, int HW, int C, int G, float eps, int silu,
                    int pix_per_block) {
  constexpr int V = 8;
  __shared__ float red[GS_THREADS];
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS], s_m1[GN_MAX_GROUPS], s_m2[GN_MAX_GROUPS];
  const int n = blockIdx.y;
  const int cpp = C / V, cpg = cpp / G;
  const bf16* xs = x + (int64_t)n * HW * ldx;
  gn_group_stats<RV>(xs, rowvec, ld_rowvec, n, part, splits, HW, cpg, G, eps, s_mean, s_rstd);
  if (threadIdx.x < G) {
    float a = 0.f, b = 0.f;
    for (int sp = 0; sp < splits; ++sp) {
      const float2 v = part2[((int64_t)n * splits + sp) * G + threadIdx.x];
      a += v.x; b += v.y;
    }
    const float inv_n = 1.0f / ((float)HW * (float)(cpg * V));
    s_m1[threadIdx.x] = a * inv_n;
    s_m2[threadIdx.x] = b * inv_n;
  }
  __syncthreads();
  const int ci = threadIdx.x % cpp, pl = threadIdx.x / cpp, ppi = blockDim.x / cpp;
  const int g = ci / cpg;
  const float mean = s_mean[g], rstd = s_rstd[g], m1 = s_m1[g], m2 = s_m2[g];
  float ga[V], be[V], sh[V], dr[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    ga[i] = gamma[ci * V + i];
    be[i] = beta[ci * V + i];
    sh[i] = ((RV ? rowvec[(int64_t)n * ld_rowvec + ci * V + i] : 0.f) - mean) * rstd;
    dr[i] = 0.f;
  }
  const int p0 = blockIdx.x * pix_per_block, p1 = min(p0 + pix_per_block, HW);
  const bf16* xb = xs + ci * V;
  const bf16* dyb = dy + (int64_t)n * HW * lddy + ci * V;
  bf16* dxb = dx + (int64_t)n * HW * lddx + ci * V;
  auto emit = [&](const uint4& rx, const uint4& rd, int p) {
    float w[V], d[V];
    unpack(rx, w);
    unpack(rd, d);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xh = fmaf(w[i], rstd, sh[i]);
      float dz = d[i];
      if (silu) dz *= silu_grad(fmaf(xh, ga[i], be[i]));
      const float o = rstd * (dz * ga[i] - m1 - xh * m2);
      dr[i] += o;
      w[i] = o;
    }
    store_chunk(dxb + (int64_t)p * lddx, w);
  };
  int p = p0 + pl;
  for (; p + ppi < p1; p += 2 * ppi) {
    const uint4 x0 = load_raw(xb + (int64_t)p * ldx), x1 = load_raw(xb + (int64_t)(p + ppi) * ldx);
    const uint4 d0 = load_raw(dyb + (int64_t)p * lddy), d1 = load_raw(dyb + (int64_t)(p + ppi) * lddy);
    emit(x0, d0, p); emit(x1, d1, p + ppi);
  }
  for (; p < p1; p += ppi) emit(load_raw(xb + (int64_t)p * ldx), load_raw(dyb + (int64_t)p * lddy), p);
  if (drowvec) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      __syncthreads();
      red[threadIdx.x] = dr[i];
      __syncthreads();
      if (pl == 0) {
        float a = 0.f;
        for (int k = 0; k < ppi; ++k) a += red[k * cpp + ci];
        atomicAdd(drowvec + (int64_t)n * ld_drowvec + ci * V + i, a);
      }
    }
  }
}

int gcd_i2(int a, int b) { return b ? gcd_i2(b, a % b) : a; }

// geometry shared by the stats producer and its consumers (gn_apply, the fused attention kernel)
void gn_stream_geometry(int hw, int channels, int& threads, int& ppi, int& splits, int& pps) {
  const int cpp = channels / 8;
  const int unit = cpp / gcd_i2(cpp, 32) * 32;
  threads = GS_THREADS / unit * unit;
  if (threads < unit) threads = unit;
  ppi = threads / cpp;
  splits = hw / (ppi * 8);
  if (splits > GS_MAX_SPLITS) splits = GS_MAX_SPLITS;
  if (splits < 1) splits = 1;
  pps = (hw + splits - 1) / splits;
}

int gn_stream_launch(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                     const float* beta, const float* rowvec, int ld_rowvec, int batch, int hw, int channels, int groups,
                     float eps, int silu, void* workspace, int x_mod, cudaStream_t st) {
  const int cpp = channels / 8;
  const int unit = cpp / gcd_i2(cpp, 32) * 32;
  int threads = GS_THREADS / unit * unit;
  if (threads < unit) threads = unit;  // unit <= 1024 guaranteed by the caller; <= 256 for the reference widths
  LDM_REQUIRE(threads <= GS_THREADS, "group_norm: %d channels need the one-CTA-per-sample kernel", channels);
  const int ppi = threads / cpp;
  // slab sizes depend only on the per-sample geometry (never on the batch): a sample's bits are the same however
  // the batch is sharded
  int splits = hw / (ppi * 8);
  if (splits > GS_MAX_SPLITS) splits = GS_MAX_SPLITS;
  if (splits < 1) splits = 1;
  const int pps = (hw + splits - 1) / splits;
  float2* part = (float2*)workspace;
  if (rowvec)
    LDM_CUDA(ldm_launch_pdl(gn_stats_kernel<true>, dim3(splits, batch), dim3(threads), 0, st, (const bf16*)x, ldx, rowvec, ld_rowvec, part, hw, channels, groups, pps, x_mod));
  else
    LDM_CUDA(ldm_launch_pdl(gn_stats_kernel<false>, dim3(splits, batch), dim3(threads), 0, st, (const bf16*)x, ldx, rowvec, ld_rowvec, part, hw, channels, groups, pps, x_mod));
  LDM_LAUNCHED("gn_stats");
  int ppb = ppi * 8;
  if (ppb > hw) ppb = hw;
  const dim3 grid((hw + ppb - 1) / ppb, batch);
  if (rowvec)
    LDM_CUDA(ldm_launch_pdl(gn_apply_kernel<true>, grid, dim3(threads), 0, st, (const bf16*)x, ldx, (bf16*)y, ldy, (const bf16*)res,
                            ldres, gamma, beta, rowvec, ld_rowvec, (const float2*)part, splits, hw, channels, groups, eps, silu, ppb, x_mod));
  else
    LDM_CUDA(ldm_launch_pdl(gn_apply_kernel<false>, grid, dim3(threads), 0, st, (const bf16*)x, ldx, (bf16*)y, ldy, (const bf16*)res,
                            ldres, gamma, beta, rowvec, ld_rowvec, (const float2*)part, splits, hw, channels, groups, eps, silu, ppb, x_mod));
  LDM_LAUNCHED("gn_apply");
  return 0;
}

int gcd_i(int a, int b) { return b ? gcd_i(b, a % b) : a; }

template <typename T>
int gn_launch(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
              const float* beta, const float* rowvec, int ld_rowvec, int batch, int hw, int channels, int groups,
              float eps, int silu, cudaStream_t st) {
  constexpr int V = VecTraits<T>::N;
  const int cpp = channels / V;
  const int unit = cpp / gcd_i(cpp, 32) * 32;  // lcm(cpp, 32): whole warps and whole pixels
  LDM_REQUIRE(unit <= GN_MAX_THREADS, "group_norm: %d channels do not fit one CTA row", channels);
  // as many threads as useful: at most one chunk-slot per (pixel, chunk), at most GN_MAX_THREADS
  int64_t want = (int64_t)cpp * hw;
  int threads = (int)((want + unit - 1) / unit) * unit;
  if (threads > GN_MAX_THREADS / unit * unit) threads = GN_MAX_THREADS / unit * unit;
  const int ppi = threads / cpp;
  const int nj = (hw + ppi - 1) / ppi;
  const T* xp = (const T*)x; T* yp = (T*)y; const T* rp = (const T*)res;
#define GN_GO(NJV)                                                                                                  \
  if (rowvec) gn_fused_kernel<T, NJV, true><<<batch, threads, 0, st>>>(xp, ldx, yp, ldy, rp, ldres, gamma, beta, rowvec, ld_rowvec, hw, channels, groups, eps, silu); \
  else gn_fused_kernel<T, NJV, false><<<batch, threads, 0, st>>>(xp, ldx, yp, ldy, rp, ldres, gamma, beta, rowvec, ld_rowvec, hw, channels, groups, eps, silu)
  if (nj <= 1) GN_GO(1);
  else if (nj <= 2) GN_GO(2);
  else if (nj <= 4) GN_GO(4);
  else if (nj <= GN_HOLD) GN_GO(8);
  else GN_GO(0);
#undef GN_GO
  LDM_LAUNCHED("group_norm");
  return 0;
}

}  // namespace

// Apply only (bf16, streaming kernel): the statistics are the un-pivoted {S, Q} slots a convolution epilogue left
// (ConvGn mode 1): part[(n * groups + g) * nslots + slot].
int k_group_norm_apply_raw(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                           const float* beta, const float* rowvec, int ld_rowvec, int batch, int hw, int channels, int groups,
                           float eps, int silu, const void* part, int nslots, int x_mod, cudaStream_t st) {
  LDM_REQUIRE(channels % groups == 0 && (channels / groups) % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && (!res || ldres % 8 == 0) &&
                  channels / 8 <= GS_THREADS && part && nslots >= 1 && groups <= GN_MAX_GROUPS && batch <= 65535 &&
                  (x_mod == 0 || batch % x_mod == 0),
              "group_norm_apply_raw: unsupported shape");
  if (batch == 0 || hw == 0) return 0;
  int threads, ppi, splits, pps;
  gn_stream_geometry(hw, channels, threads, ppi, splits, pps);
  int ppb = ppi * 8;
  if (ppb > hw) ppb = hw;
  const dim3 grid((hw + ppb - 1) / ppb, batch);
  if (rowvec)
    LDM_CUDA(ldm_launch_pdl(gn_apply_kernel<true>, grid, dim3(threads), 0, st, (const bf16*)x, ldx, (bf16*)y, ldy, (const bf16*)res,
                            ldres, gamma, beta, rowvec, ld_rowvec, (const float2*)part, -nslots, hw, channels, groups, eps, silu, ppb, x_mod));
  else
    LDM_CUDA(ldm_launch_pdl(gn_apply_kernel<false>, grid, dim3(threads), 0, st, (const bf16*)x, ldx, (bf16*)y, ldy, (const bf16*)res,
                            ldres, gamma, beta, (const float*)nullptr, 0, (const float2*)part, -nslots, hw, channels, groups, eps, silu, ppb, x_mod));
  LDM_LAUNCHED("gn_apply");
  return 0;
}

// Per-(sample, channel) affine form of GroupNorm for a consumer that normalises on the fly (conv_halo.cu's transform warps):
// the statistics are combined exactly as gn_apply_kernel does (same order, same precision).
__global__ void __launch_bounds__(256)
gn_coef_kernel(const float2* __restrict__ part, int splits, const float* __restrict__ gamma, const float* __restrict__ beta,
               const float* __restrict__ rowvec, int ld_rowvec, int HW, int C, int G, float eps, int x_mod, int rows,
               const bf16* __restrict__ x, int ldx, int silu, float2* __restrict__ ab) {
  __shared__ float s_mean[GN_MAX_GROUPS], s_rstd[GN_MAX_GROUPS];
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.x;
  const int cpg = C / G;
  if ((int)threadIdx.x < G && splits < 0) {
    const int gg = threadIdx.x, ns = -splits;
    const int nv = x_mod > 0 ? rows / x_mod : 1, img = x_mod > 0 ? n % x_mod : n, kk = x_mod > 0 ? n / x_mod : 0;
    double a = 0.0, b = 0.0;
    for (int sp = 0; sp < ns; ++sp) {
      const float2 v = part[(((int64_t)img * G + gg) * nv + kk) * ns + sp];
      a += (double)v.x; b += (double)v.y;
    }
    const double cnt = (double)HW * (double)cpg;
    const double m1 = a / cnt;
    double var = b / cnt - m1 * m1;
    if (var < 0.0) var = 0.0;
    s_mean[gg] = (float)m1;
    s_rstd[gg] = (float)(1.0 / sqrt(var + (double)eps));
  } else if ((int)threadIdx.x < G) {
    const int gg = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int sp = 0; sp < splits; ++sp) {  // fixed order
      const float2 v = part[((int64_t)n * splits + sp) * G + gg];
      a += v.x; b += v.y;
    }
    const int c_first = gg * cpg;
    const bf16* xs = x + (int64_t)(x_mod > 0 ? n % x_mod : n) * HW * ldx;
    const float K = __bfloat162float(xs[c_first]) + (rowvec ? rowvec[(int64_t)n * ld_rowvec + c_first] : 0.f);
    const float inv_n = 1.0f / ((float)HW * (float)cpg);
    const float m1 = a * inv_n;
    const float var = fmaxf(b * inv_n - m1 * m1, 0.f);
    s_mean[gg] = K + m1;
    s_rstd[gg] = 1.0f / sqrtf(var + eps);
  }
  __syncthreads();
  const float h = silu ? 0.5f : 1.0f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float a = gamma[c] * s_rstd[g];
    const float rvi = rowvec ? rowvec[(int64_t)n * ld_rowvec + c] : 0.f;
    const float b = beta[c] + (rvi - s_mean[g]) * a;
    ab[(int64_t)n * C + c] = make_float2(h * a, h * b);
  }
}
int k_group_norm_coef(const void* part, int nslots, const float* gamma, const float* beta, const float* rowvec, int ld_rowvec,
                      int rows, int hw, int channels, int groups, float eps, int x_mod, const void* x, int ldx, int silu,
                      void* ab_out, cudaStream_t st) {
  LDM_REQUIRE(part && nslots != 0 && gamma && beta && ab_out && channels % groups == 0 && groups <= GN_MAX_GROUPS &&
                  (nslots < 0 || x) && (x_mod == 0 || rows % x_mod == 0), "group_norm_coef: unsupported arguments");
  if (rows == 0) return 0;
  LDM_CUDA(ldm_launch_pdl(gn_coef_kernel, dim3(rows), dim3(256), 0, st, (const float2*)part, nslots, gamma, beta, rowvec, ld_rowvec, hw,
                          channels, groups, eps, x_mod, rows, (const bf16*)x, ldx, silu, (float2*)ab_out));
  LDM_LAUNCHED("gn_coef");
  return 0;
}

// max-pool 2x2 of x [B, H, W, ldx] (C channels) -> pool [B, H/2, W/2, ldp] and y = [silu](GroupNorm(pool)) [B, H/2, W/2, ldy]
bool k_pool_group_norm_applicable(int H, int W, int C, int groups, int dtype) {
  if (dtype != LDM_DT_BF16 || H % 2 || W % 2 || C % 8 || groups < 1 || groups > GN_MAX_GROUPS) return false;
  const int cpp = C / 8;
  if (cpp > 256 || 256 % cpp != 0 || C % groups != 0 || (C / groups) % 8 != 0) return false;
  const int hw2 = (H / 2) * (W / 2), ppi = 256 / cpp;
  return (hw2 + ppi - 1) / ppi <= 8 && getenv("LDM_NO_POOL_GN") == nullptr;
}
int k_pool_group_norm(const void* x, int ldx, int H, int W, int C, void* pool, int ldp, void* y, int ldy, const float* gamma,
                      const float* beta, int groups, float eps, int silu, int batch, cudaStream_t st) {
  LDM_REQUIRE(k_pool_group_norm_applicable(H, W, C, groups, LDM_DT_BF16) && ldx % 8 == 0 && ldp % 8 == 0 && ldy % 8 == 0,
              "pool_group_norm: unsupported shape %dx%dx%d", H, W, C);
  if (batch == 0) return 0;
  const int hw2 = (H / 2) * (W / 2), ppi = 256 / (C / 8), nj = (hw2 + ppi - 1) / ppi;
#define PG_GO(NJV) LDM_CUDA(ldm_launch_pdl(pool_gn_kernel<NJV>, dim3(batch), dim3(256), 0, st, (const bf16*)x, ldx, W, (bf16*)pool, ldp, (bf16*)y, ldy, gamma, beta, hw2, C, groups, eps, silu))
  if (nj <= 1) PG_GO(1); else if (nj <= 2) PG_GO(2); else if (nj <= 4) PG_GO(4); else PG_GO(8);
#undef PG_GO
  LDM_LAUNCHED("pool_group_norm");
  return 0;
}

// Statistics only (bf16, no rowvec): part[(n*splits + s)*groups + g] = {sum(x-K), sum((x-K)^2)} with the pivot
// K = x[n][pixel 0][first channel of group g].  Consumers rebuild mean / rstd from it (see gn_apply_kernel).
bool k_group_norm_streams(int hw, int channels, int dtype) {
  return dtype == LDM_DT_BF16 && channels / 8 <= GS_THREADS && (int64_t)hw * channels > 2048 && getenv("LDM_GN_ONE_CTA") == nullptr;
}

int k_group_norm_stats(const void* x, int ldx, int batch, int hw, int channels, int groups, void* workspace, int* splits_out,
                       cudaStream_t st) {
  LDM_REQUIRE(channels % groups == 0 && (channels / groups) % 8 == 0 && ldx % 8 == 0 && channels / 8 <= GS_THREADS && workspace,
              "group_norm_stats: unsupported shape");
  int threads, ppi, splits, pps;
  gn_stream_geometry(hw, channels, threads, ppi, splits, pps);
  *splits_out = splits;
  if (batch == 0) return 0;
  LDM_CUDA(ldm_launch_pdl(gn_stats_kernel<false>, dim3(splits, batch), dim3(threads), 0, st, (const bf16*)x, ldx, (const float*)nullptr,
                          0, (float2*)workspace, hw, channels, groups, pps, 0));
  LDM_LAUNCHED("gn_stats");
  return 0;
}

// Streaming GroupNorm backward (bf16).  workspace: k_group_norm_backward_ws_bytes.  fwd_part: the partial sums the forward's
// workspace holds (k_group_norm_rv with the same arguments) or nullptr to recompute them.
int64_t k_group_norm_backward_ws_bytes(int batch, int hw, int channels, int groups) {
  int threads, ppi, splits, pps;
  gn_stream_geometry(hw, channels > 0 ? channels : 8, threads, ppi, splits, pps);
  return 2 * k_group_norm_ws_bytes(batch, groups) + (int64_t)batch * splits * 2 * channels * sizeof(float) + 256;
}
bool k_group_norm_backward_streams(int batch, int hw, int channels, int groups, int dtype) {
  return k_group_norm_streams(hw, channels, dtype) && channels % groups == 0 && (channels / groups) % 8 == 0 && batch <= 65535 &&
         getenv("LDM_GN_BWD_ONE_CTA") == nullptr;
}
int k_group_norm_backward_stream(const void* x, int ldx, const void* dy, int lddy, const float* gamma, const float* beta,
                                 const float* rowvec, int ld_rowvec, void* dx, int lddx, float* dgamma, float* dbeta,
                                 float* drowvec, int ld_drowvec, int batch, int hw, int channels, int groups, float eps, int silu,
                                 const void* fwd_part, void* workspace, cudaStream_t st) {
  LDM_REQUIRE(ldx % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 && workspace, "group_norm_backward: unaligned stride / no workspace");
  if (batch == 0 || hw == 0) return 0;
  int threads, ppi, splits, pps;
  gn_stream_geometry(hw, channels, threads, ppi, splits, pps);
  float2* part = (float2*)workspace;
  float2* part2 = (float2*)((char*)workspace + k_group_norm_ws_bytes(batch, groups));
  float* chan_part = (float*)((char*)workspace + 2 * k_group_norm_ws_bytes(batch, groups));
  const bf16* xp = (const bf16*)x;
  const bf16* dyp = (const bf16*)dy;
  if (fwd_part) {
    part = (float2*)fwd_part;
  } else {
    if (rowvec) gn_stats_kernel<true><<<dim3(splits, batch), threads, 0, st>>>(xp, ldx, rowvec, ld_rowvec, part, hw, channels, groups, pps, 0);
    else gn_stats_kernel<false><<<dim3(splits, batch), threads, 0, st>>>(xp, ldx, rowvec, ld_rowvec, part, hw, channels, groups, pps, 0);
    LDM_LAUNCHED("gn_stats");
  }
  if (rowvec) gn_bwd_sums_kernel<true><<<dim3(splits, batch), threads, 0, st>>>(xp, ldx, dyp, lddy, gamma, beta, rowvec, ld_rowvec, part, part2, chan_part, drowvec, ld_drowvec, hw, channels, groups, eps, silu, pps);
  else gn_bwd_sums_kernel<false><<<dim3(splits, batch), threads, 0, st>>>(xp, ldx, dyp, lddy, gamma, beta, rowvec, ld_rowvec, part, part2, chan_part, drowvec, ld_drowvec, hw, channels, groups, eps, silu, pps);
  LDM_LAUNCHED("gn_bwd_sums");
  gn_bwd_param_kernel<<<(2 * channels + 31) / 32, 256, 0, st>>>(chan_part, batch * splits, channels, dgamma, dbeta);
  LDM_LAUNCHED("gn_bwd_param");
  int ppb = ppi * 8;
  if (ppb > hw) ppb = hw;
  const dim3 grid((hw + ppb - 1) / ppb, batch);
  if (rowvec) gn_bwd_apply_kernel<true><<<grid, threads, 0, st>>>(xp, ldx, dyp, lddy, gamma, beta, rowvec, ld_rowvec, part, part2, splits, (bf16*)dx, lddx, drowvec, ld_drowvec, hw, channels, groups, eps, silu, ppb);
  else gn_bwd_apply_kernel<false><<<grid, threads, 0, st>>>(xp, ldx, dyp, lddy, gamma, beta, rowvec, ld_rowvec, part, part2, splits, (bf16*)dx, lddx, drowvec, ld_drowvec, hw, channels, groups, eps, silu, ppb);
  LDM_LAUNCHED("gn_bwd_apply");
  return 0;
}

int64_t k_group_norm_ws_bytes(int batch, int groups) {
  return (int64_t)batch * GS_MAX_SPLITS * groups * sizeof(float2) + 1024;
}

int k_group_norm(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                 const float* beta, int batch, int hw, int channels, int groups, float eps, int silu, int dtype,
                 void* workspace, cudaStream_t st) {
  return k_group_norm_rv(x, ldx, y, ldy, res, ldres, gamma, beta, nullptr, 0, batch, hw, channels, groups, eps, silu, dtype,
                         workspace, st);
}

int k_group_norm_rv(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                    const float* beta, const float* rowvec, int ld_rowvec, int batch, int hw, int channels, int groups,
                    float eps, int silu, int dtype, void* workspace, cudaStream_t st) {
  return k_group_norm_mod(x, ldx, y, ldy, res, ldres, gamma, beta, rowvec, ld_rowvec, batch, hw, channels, groups, eps, silu, dtype,
                          workspace, 0, st);
}

int k_group_norm_mod(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                     const float* beta, const float* rowvec, int ld_rowvec, int batch, int hw, int channels, int groups,
                     float eps, int silu, int dtype, void* workspace, int x_mod, cudaStream_t st) {
  const int V = dtype == LDM_DT_BF16 ? 8 : 4;
  if (groups >= 1 && channels % groups == 0 && (channels / groups) % V != 0 && res == nullptr && rowvec == nullptr && x_mod == 0)
    // groups narrower than a vector chunk (the autoencoder's GroupNorm(32, C) at 64 / 128 channels, src/Autoencoder.py:9-11)
    return k_group_norm_any(x, ldx, y, ldy, gamma, beta, batch, hw, channels, groups, eps, silu, dtype, st);
  LDM_REQUIRE(groups >= 1 && groups <= GN_MAX_GROUPS, "group_norm: groups=%d unsupported", groups);
  LDM_REQUIRE(channels % groups == 0 && (channels / groups) % V == 0,
              "group_norm: channels/groups (%d/%d) must be a multiple of %d", channels, groups, V);
  LDM_REQUIRE(ldx % V == 0 && ldy % V == 0 && (res == nullptr || ldres % V == 0), "group_norm: unaligned stride");
  LDM_REQUIRE(channels / V <= 1024, "group_norm: too many channels (%d)", channels);
  if (batch == 0 || hw == 0) return 0;
  // tiny samples (the 2x2 bottleneck): one CTA per sample holds everything in registers -- one launch and one round trip
  // instead of two of each
  const bool tiny = (int64_t)hw * channels <= 2048;
  if (dtype == LDM_DT_BF16 && workspace != nullptr && channels / 8 <= GS_THREADS && batch <= 65535 && !tiny &&
      getenv("LDM_GN_ONE_CTA") == nullptr)
    return gn_stream_launch(x, ldx, y, ldy, res, ldres, gamma, beta, rowvec, ld_rowvec, batch, hw, channels, groups, eps,
                            silu, workspace, x_mod, st);
  LDM_REQUIRE(x_mod == 0, "group_norm: row aliasing is only implemented in the streaming bf16 kernels");
  if (dtype == LDM_DT_BF16)
    return gn_launch<bf16>(x, ldx, y, ldy, res, ldres, gamma, beta, rowvec, ld_rowvec, batch, hw, channels, groups, eps, silu, st);
  return gn_launch<float>(x, ldx, y, ldy, res, ldres, gamma, beta, rowvec, ld_rowvec, batch, hw, channels, groups, eps, silu, st);
}
