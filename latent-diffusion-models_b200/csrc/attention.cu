// Attention cores on NHWC qkv tensors (the 1x1 to_qkv / to_out convolutions run on the conv kernels).
//   LinearAttention  src/UNet.py:149-163   q <- softmax_d(q) * 32^-1/2 ; k <- softmax_n(k) ;
//                                          ctx = k v^T (32x32 per head) ; out = ctx^T q
//   Attention        src/UNet.py:122-135   sim = (q 32^-1/2)^T k ; softmax_j ; out = attn v^T
// qkv is [B, N, 384]: channel = part*128 + head*32 + c (part 0/1/2 = q/k/v; "b (h c) x y" split, :124-127);
// out is [B, N, 128] with channel = head*32 + c.
// One CTA per (sample, head).  fp32 math throughout; T only selects the storage type.
#include <stdlib.h>

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

// ------------------------------------------------------------------------------------------------------------
// bf16 LinearAttention on the legacy tensor path (mma.sync m16n8k16, bf16 in / fp32 accumulate).
// The two GEMMs of a head are 32x32xN (ctx = softmax_N(k) v^T) and Nx32x32 (out = softmax_d(q) ctx): far too small
// for a tcgen05 tile (M >= 64 per CTA), and the kernel is HBM-bound by an order of magnitude (0.5 KB moved per
// token for 8 kFLOP), so the point of the tensor pipe here is only to get the FFMA work out of the way.
//   one CTA per sample, one warp per head; every warp runs its own cp.async double buffer (no __syncthreads)
//   phase 1: stream k,v once, online softmax over the tokens (running per-channel max, accumulators rescaled in
//            registers), ctx accumulated in the C fragments
//   phase 2: stream q once, softmax over the 32 head channels inside a quad, out tile staged through shared memory
//            so that global stores are 16-byte / full-sector
// Every qkv element is read exactly once and every out element written once.
namespace {

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void hmma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float2 unpack_bf2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ uint32_t pack_bf2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

constexpr int LM_TILE = 32;                 // tokens per pipeline stage
constexpr int LM_ROW = 64;                  // bytes per token row of one head (32 bf16)
constexpr int LM_BUF = LM_TILE * LM_ROW;    // 2 KB
constexpr float LOG2E = 1.4426950408889634f;

// byte offset of 16-byte chunk `c` (0..3) of row `r` inside a [rows][64 B] tile; the XOR keeps the 8 row addresses
// of every ldmatrix 8x8 block (and the quad-strided fragment stores) on distinct bank groups
__device__ __forceinline__ uint32_t lm_off(int r, int c) { return (uint32_t)(r * LM_ROW + ((c ^ ((r >> 1) & 3)) << 4)); }

// AFF: qkv holds the RAW projection (W diag(gamma)) x of the un-normalised block input; the PreNorm GroupNorm(1, C) is applied
// here as the per-sample affine map it is (q = r q_raw + cq, k = r k_raw + const, v = r v_raw + cv with r the sample's rstd and
// cq / cv = fold constants - r mu rowsums: k_fold_prenorm_qkv), so the normalised tensor is never written (src/UNet.py:106-110).
struct LaAffine {
  const float* uv;             // [2][384] fold constants
  const float2* gn_part; int gn_splits; float gn_eps;   // statistics: > 0 pivoted slabs, < 0 raw {S, Q} slots
  const bf16* x; int ldx;      // raw input (pivot of the slab statistics)
  int cin;
};
template <bool AFF>
__global__ void __launch_bounds__(128) linattn_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int N, const LaAffine af) {
  // per warp: [stage][k|v] tiles of 2 KB
  __shared__ __align__(128) uint8_t smem[4][2][2][LM_BUF];
  const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  float gr = 1.f, gmu = 0.f;
  if (AFF) {
    const double cnt = (double)N * (double)af.cin;
    if (af.gn_splits < 0) {
      double a = 0.0, q2 = 0.0;
      for (int sp = lane; sp < -af.gn_splits; sp += 32) { const float2 v = __ldg(af.gn_part + (int64_t)b * (-af.gn_splits) + sp); a += (double)v.x; q2 += (double)v.y; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q2 += __shfl_xor_sync(0xffffffffu, q2, o); }
      const double m1 = a / cnt;
      double var = q2 / cnt - m1 * m1;
      if (var < 0.0) var = 0.0;
      gmu = (float)m1;
      gr = (float)(1.0 / sqrt(var + (double)af.gn_eps));
    } else {
      float a = 0.f, q2 = 0.f;
      for (int sp = 0; sp < af.gn_splits; ++sp) { const float2 v = __ldg(af.gn_part + (int64_t)b * af.gn_splits + sp); a += v.x; q2 += v.y; }
      const float K = __bfloat162float(af.x[(int64_t)b * N * af.ldx]);
      const float inv_n = 1.0f / (float)cnt;
      const float m1 = a * inv_n;
      const float var = fmaxf(q2 * inv_n - m1 * m1, 0.f);
      gmu = K + m1;
      gr = 1.0f / sqrtf(var + af.gn_eps);
    }
  }
  const float KL = gr * 1.4426950408889634f;     // log2(e) times the sample's rstd (1 when not AFF)
  const bf16* base = qkv + (int64_t)b * N * 384 + h * LA_D;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&smem[h][0][0][0]);
  auto buf = [&](int stage, int which) { return sbase + (uint32_t)((stage * 2 + which) * LM_BUF); };

  const int steps = N >> 4;              // 16-token k-steps
  const int tiles = (steps + 1) >> 1;    // 32-token pipeline tiles (the last one may hold a single step)
  pdl_wait();
  pdl_trigger();

  // ---------------- phase 1: ctx[d][e] = sum_n exp(k[n][d] - m[d]) v[n][e]
  auto load_kv = [&](int tile, int stage) {
    const int tok0 = tile * LM_TILE;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;              // 0..127: k chunks, 128..255: v chunks
      const int which = c >> 7, r = (c & 127) >> 2, ch = c & 3;
      if (tok0 + r < N)
        cp_async16(buf(stage, which) + lm_off(r, ch), base + (int64_t)(tok0 + r) * 384 + 128 * (1 + which) + ch * 8);
    }
  };
  float acc[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  float m_run[2][2], z[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) { m_run[mt][hf] = -INFINITY; z[mt][hf] = 0.f; }

  load_kv(0, 0);
  cp_async_commit();
  if (tiles > 1) load_kv(1, 1);
  cp_async_commit();
  for (int t = 0; t < tiles; ++t) {
    cp_async_wait<1>();
    __syncwarp();
    const int stage = t & 1;
    const int nst = min(2, steps - 2 * t);
    uint32_t a[2][2][4];
    // A fragments of P^T: rows = head channel d, cols = token.  Matrix i of the x4 load: a0 (tok 0-7, d 0-7),
    // a1 (tok 0-7, d 8-15), a2 (tok 8-15, d 0-7), a3 (tok 8-15, d 8-15), transposed on the way in.
#pragma unroll
    for (int st = 0; st < 2; ++st)
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        if (st < nst) {
          const int r = st * 16 + ((lane >> 4) << 3) + (lane & 7);
          const int ch = mt * 2 + ((lane >> 3) & 1);
          ldsm_x4_t(a[st][mt], buf(stage, 0) + lm_off(r, ch));
        } else {
          a[st][mt][0] = a[st][mt][1] = a[st][mt][2] = a[st][mt][3] = 0u;
        }
      }
    // running per-channel max over the tokens
    float sc[2][2], ml2[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float mx = -INFINITY;
#pragma unroll
        for (int st = 0; st < 2; ++st)
          if (st < nst) {
            float2 x0 = unpack_bf2(a[st][mt][hf]), x1 = unpack_bf2(a[st][mt][hf + 2]);
            mx = fmaxf(mx, fmaxf(fmaxf(x0.x, x0.y), fmaxf(x1.x, x1.y)));
          }
        mx = quad_max(mx);
        const float nm = fmaxf(m_run[mt][hf], mx);
        sc[mt][hf] = ex2f((m_run[mt][hf] - nm) * KL);
        m_run[mt][hf] = nm;
        ml2[mt][hf] = nm * KL;
        z[mt][hf] *= sc[mt][hf];
      }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        acc[mt][nt][0] *= sc[mt][0]; acc[mt][nt][1] *= sc[mt][0];
        acc[mt][nt][2] *= sc[mt][1]; acc[mt][nt][3] *= sc[mt][1];
      }
    // p = exp(k - m), in place in the A fragments
#pragma unroll
    for (int st = 0; st < 2; ++st)
      if (st < nst) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int hf = i & 1;
            float2 x = unpack_bf2(a[st][mt][i]);
            const float p0 = ex2f(fmaf(x.x, KL, -ml2[mt][hf])), p1 = ex2f(fmaf(x.y, KL, -ml2[mt][hf]));
            z[mt][hf] += p0 + p1;
            a[st][mt][i] = pack_bf2(p0, p1);
          }
      }
    // B fragments of V (rows = token, cols = e) and the MMAs
#pragma unroll
    for (int st = 0; st < 2; ++st)
      if (st < nst) {
        uint32_t bv[2][4];
#pragma unroll
        for (int jp = 0; jp < 2; ++jp) {
          // matrices: (tok 0-7, chunk 2jp), (tok 8-15, chunk 2jp), (tok 0-7, chunk 2jp+1), (tok 8-15, chunk 2jp+1)
          const int r = st * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
          const int ch = jp * 2 + (lane >> 4);
          ldsm_x4_t(bv[jp], buf(stage, 1) + lm_off(r, ch));
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) hmma_bf16(acc[mt][nt], a[st][mt], bv[nt >> 1][(nt & 1) * 2], bv[nt >> 1][(nt & 1) * 2 + 1]);
      }
    __syncwarp();
    if (t + 2 < tiles) load_kv(t + 2, stage);
    cp_async_commit();
  }
  cp_async_wait<0>();
  __syncwarp();

  // ---------------- ctx -> bf16 in shared memory, rows scaled by 32^-1/2 / Z[d]  (src/UNet.py:156-161)
  const uint32_t ctx_s = buf(0, 0);
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float f = 0.17677669529663687f * gr / quad_sum(z[mt][hf]);
      const int d = mt * 16 + hf * 8 + g;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float c0 = 0.f, c1 = 0.f;                 // v constants of columns e = 8 nt + 2 tq, + 1 (times 32^-1/2)
        if (AFF) {
          const int e = 256 + 32 * h + 8 * nt + 2 * tq;
          c0 = 0.17677669529663687f * (af.uv[384 + e] - gr * gmu * af.uv[e]);
          c1 = 0.17677669529663687f * (af.uv[384 + e + 1] - gr * gmu * af.uv[e + 1]);
        }
        const uint32_t v = pack_bf2(fmaf(acc[mt][nt][hf * 2], f, c0), fmaf(acc[mt][nt][hf * 2 + 1], f, c1));
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ctx_s + lm_off(d, nt) + tq * 4), "r"(v) : "memory");
      }
    }
  __syncwarp();
  uint32_t bc[2][2][4];  // [k-step over d][pair of n-tiles]
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int jp = 0; jp < 2; ++jp) {
      const int r = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
      const int ch = jp * 2 + (lane >> 4);
      ldsm_x4_t(bc[ks][jp], ctx_s + lm_off(r, ch));
    }
  __syncwarp();

  // ---------------- phase 2: out[n][e] = sum_d softmax_d(q[n][:])[d] * ctx[d][e]
  auto load_q = [&](int tile, int stage) {
    const int tok0 = tile * LM_TILE;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = lane + 32 * i, r = c >> 2, ch = c & 3;
      if (tok0 + r < N) cp_async16(buf(stage, 1) + lm_off(r, ch), base + (int64_t)(tok0 + r) * 384 + ch * 8);
    }
  };
  float2 cq[4];                    // q constants of this lane's channels d = 16 ks + 8 j + 2 tq, + 1 (index 2 ks + j)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    cq[i] = make_float2(0.f, 0.f);
    if (AFF) {
      const int d = 32 * h + 16 * (i >> 1) + 8 * (i & 1) + 2 * tq;
      cq[i] = make_float2(af.uv[384 + d] - gr * gmu * af.uv[d], af.uv[384 + d + 1] - gr * gmu * af.uv[d + 1]);
    }
  }
  const uint32_t stg = buf(1, 0);  // 16-token output staging tile
  bf16* obase = out + (int64_t)b * N * 128 + h * LA_D;
  load_q(0, 0);
  cp_async_commit();
  if (tiles > 1) load_q(1, 1);
  cp_async_commit();
  for (int t = 0; t < tiles; ++t) {
    cp_async_wait<1>();
    __syncwarp();
    const int stage = t & 1;
    const int nst = min(2, steps - 2 * t);
    for (int mi = 0; mi < nst; ++mi) {
      uint32_t aq[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        // matrices: a0 (tok 0-7, d-chunk 2ks), a1 (tok 8-15, 2ks), a2 (tok 0-7, 2ks+1), a3 (tok 8-15, 2ks+1)
        const int r = mi * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int ch = ks * 2 + (lane >> 4);
        ldsm_x4(aq[ks], buf(stage, 1) + lm_off(r, ch));
      }
      // softmax over the 32 channels of a token: rows g (regs 0,2) and g+8 (regs 1,3), spread over the quad
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float2 x[4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) { x[2 * ks] = unpack_bf2(aq[ks][hf]); x[2 * ks + 1] = unpack_bf2(aq[ks][hf + 2]); }
        if (AFF) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { x[i].x = fmaf(x[i].x, gr, cq[i].x); x[i].y = fmaf(x[i].y, gr, cq[i].y); }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 4; ++i) mx = fmaxf(mx, fmaxf(x[i].x, x[i].y));
        mx = quad_max(mx) * LOG2E;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          x[i].x = ex2f(fmaf(x[i].x, LOG2E, -mx)); x[i].y = ex2f(fmaf(x[i].y, LOG2E, -mx));
          s += x[i].x + x[i].y;
        }
        const float inv = 1.0f / quad_sum(s);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          aq[ks][hf] = pack_bf2(x[2 * ks].x * inv, x[2 * ks].y * inv);
          aq[ks][hf + 2] = pack_bf2(x[2 * ks + 1].x * inv, x[2 * ks + 1].y * inv);
        }
      }
      float o[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) hmma_bf16(o[nt], aq[ks], bc[ks][nt >> 1][(nt & 1) * 2], bc[ks][nt >> 1][(nt & 1) * 2 + 1]);
      }
      // stage the 16x32 tile, then 16-byte coalesced stores
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg + lm_off(g, nt) + tq * 4), "r"(pack_bf2(o[nt][0], o[nt][1])) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg + lm_off(g + 8, nt) + tq * 4), "r"(pack_bf2(o[nt][2], o[nt][3])) : "memory");
      }
      __syncwarp();
      const int tok0 = t * LM_TILE + mi * 16;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = lane + 32 * i, r = c >> 2, ch = c & 3;
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stg + lm_off(r, ch)));
        *reinterpret_cast<uint4*>(obase + (int64_t)(tok0 + r) * 128 + ch * 8) = v;
      }
      __syncwarp();
    }
    if (t + 2 < tiles) load_q(t + 2, stage);
    cp_async_commit();
  }
  cp_async_wait<0>();
}


// ------------------------------------------------------------------------------------------------------------
// LinearAttention backward on mma.sync (bf16 operands, fp32 accumulate), two kernels:
//   linattn_bwd_ctx_kernel    grid (batch), one warp per head: a pass over (k, v) like the forward's phase 1 gives
//                             ctx[d][e] = sum_t ks[t][d] v[t][e] (ks = softmax over tokens) with its row max and 1/Z; a pass
//                             over (q, dout) gives dctx[d][e] = sum_t qs[t][d] dout[t][e] (qs = 32^-1/2 softmax_d(q)).
//   linattn_bwd_apply_kernel  grid (batch, token chunks): per 16-token tile  dqs = dout ctx^T, dks = v dctx^T, dv = ks dctx,
//                             then dq = qs (dqs - <qs,dqs>/s), dk = ks (dks - cs[d]) with cs[d] = sum_e dctx[d][e] ctx[d][e]
//                             (= sum_t ks dks, so no third pass over the tokens is needed).
// ws per (sample, head): ctx[32][32], dctx[32][32], kmax[32], kzinv[32]  (fp32)
constexpr int LB_WS = 2 * 32 * 32 + 64;

// grid (batch, S): split s takes tokens [s*chunk, (s+1)*chunk) and writes UNNORMALISED partial results (ctx accumulator with
// its running row max and row sum, dctx); the apply kernel merges the S partials.  Small batches get S > 1 so that the pass
// fills the machine.
__global__ void __launch_bounds__(128) linattn_bwd_ctx_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                              float* __restrict__ ws, int N, int chunk) {
  __shared__ __align__(128) uint8_t smem[4][2][2][LM_BUF];
  const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const bf16* base = qkv + (int64_t)b * N * 384 + h * LA_D;
  const bf16* dob = dout + (int64_t)b * N * 128 + h * LA_D;
  const int S = gridDim.y, sp = blockIdx.y;
  const int tok_b = sp * chunk, tok_e = min(N, tok_b + chunk);
  float* wsp = ws + (((int64_t)b * 4 + h) * S + sp) * LB_WS;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&smem[h][0][0][0]);
  auto buf = [&](int stage, int which) { return sbase + (uint32_t)((stage * 2 + which) * LM_BUF); };
  const int steps = (tok_e - tok_b) >> 4;
  const int tiles = (steps + 1) >> 1;

  // pass 0: (k, v) -> ctx;  pass 1: (q, dout) -> dctx
  for (int pass = 0; pass < 2; ++pass) {
    auto load = [&](int tile, int stage) {
      const int tok0 = tok_b + tile * LM_TILE;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = lane + 32 * i;
        const int which = c >> 7, r = (c & 127) >> 2, ch = c & 3;
        if (tok0 + r < tok_e) {
          const bf16* src = pass == 0 ? base + (int64_t)(tok0 + r) * 384 + 128 * (1 + which) + ch * 8
                                      : (which == 0 ? base + (int64_t)(tok0 + r) * 384 + ch * 8
                                                    : dob + (int64_t)(tok0 + r) * 128 + ch * 8);
          cp_async16(buf(stage, which) + lm_off(r, ch), src);
        }
      }
    };
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    float m_run[2][2], z[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) { m_run[mt][hf] = -INFINITY; z[mt][hf] = 0.f; }

    load(0, 0);
    cp_async_commit();
    if (tiles > 1) load(1, 1);
    cp_async_commit();
    for (int t = 0; t < tiles; ++t) {
      cp_async_wait<1>();
      __syncwarp();
      const int stage = t & 1;
      const int nst = min(2, steps - 2 * t);
      uint32_t a[2][2][4];   // A fragments of the transposed left operand: rows = head channel d, cols = token
#pragma unroll
      for (int st = 0; st < 2; ++st)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (st < nst) {
            const int r = st * 16 + ((lane >> 4) << 3) + (lane & 7);
            const int ch = mt * 2 + ((lane >> 3) & 1);
            ldsm_x4_t(a[st][mt], buf(stage, 0) + lm_off(r, ch));
          } else {
            a[st][mt][0] = a[st][mt][1] = a[st][mt][2] = a[st][mt][3] = 0u;
          }
        }
      if (pass == 0) {
        // online softmax over the tokens (rows d are fixed per thread), exactly the forward's phase 1
        float sc[2][2], ml2[2][2];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            float mx = -INFINITY;
#pragma unroll
            for (int st = 0; st < 2; ++st)
              if (st < nst) {
                float2 x0 = unpack_bf2(a[st][mt][hf]), x1 = unpack_bf2(a[st][mt][hf + 2]);
                mx = fmaxf(mx, fmaxf(fmaxf(x0.x, x0.y), fmaxf(x1.x, x1.y)));
              }
            mx = quad_max(mx);
            const float nm = fmaxf(m_run[mt][hf], mx);
            sc[mt][hf] = ex2f((m_run[mt][hf] - nm) * LOG2E);
            m_run[mt][hf] = nm;
            ml2[mt][hf] = nm * LOG2E;
            z[mt][hf] *= sc[mt][hf];
          }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            acc[mt][nt][0] *= sc[mt][0]; acc[mt][nt][1] *= sc[mt][0];
            acc[mt][nt][2] *= sc[mt][1]; acc[mt][nt][3] *= sc[mt][1];
          }
#pragma unroll
        for (int st = 0; st < 2; ++st)
          if (st < nst) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int hf = i & 1;
                float2 x = unpack_bf2(a[st][mt][i]);
                const float p0 = ex2f(fmaf(x.x, LOG2E, -ml2[mt][hf])), p1 = ex2f(fmaf(x.y, LOG2E, -ml2[mt][hf]));
                z[mt][hf] += p0 + p1;
                a[st][mt][i] = pack_bf2(p0, p1);
              }
          }
      } else {
        // softmax over d for every token column: a token's 32 channels are rows g, g+8 of both m-tiles on the 8 lanes
        // that share tq -> reduce over lane bits 2..4
#pragma unroll
        for (int st = 0; st < 2; ++st)
          if (st < nst) {
#pragma unroll
            for (int cg = 0; cg < 2; ++cg) {   // registers (0,1): tokens 2tq, 2tq+1;  (2,3): tokens 2tq+8, 2tq+9
              float2 x[2][2];
#pragma unroll
              for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) x[mt][hf] = unpack_bf2(a[st][mt][cg * 2 + hf]);
              float m0 = fmaxf(fmaxf(x[0][0].x, x[0][1].x), fmaxf(x[1][0].x, x[1][1].x));
              float m1 = fmaxf(fmaxf(x[0][0].y, x[0][1].y), fmaxf(x[1][0].y, x[1][1].y));
#pragma unroll
              for (int o = 4; o <= 16; o <<= 1) {
                m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
                m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
              }
              m0 *= LOG2E; m1 *= LOG2E;
              float s0 = 0.f, s1 = 0.f;
#pragma unroll
              for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                  x[mt][hf].x = ex2f(fmaf(x[mt][hf].x, LOG2E, -m0));
                  x[mt][hf].y = ex2f(fmaf(x[mt][hf].y, LOG2E, -m1));
                  s0 += x[mt][hf].x; s1 += x[mt][hf].y;
                }
#pragma unroll
              for (int o = 4; o <= 16; o <<= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
              }
              const float i0 = 0.17677669529663687f / s0, i1 = 0.17677669529663687f / s1;
#pragma unroll
              for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) a[st][mt][cg * 2 + hf] = pack_bf2(x[mt][hf].x * i0, x[mt][hf].y * i1);
            }
          }
      }
      // B fragments of the right operand (rows = token, cols = e) and the MMAs
#pragma unroll
      for (int st = 0; st < 2; ++st)
        if (st < nst) {
          uint32_t bv[2][4];
#pragma unroll
          for (int jp = 0; jp < 2; ++jp) {
            const int r = st * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
            const int ch = jp * 2 + (lane >> 4);
            ldsm_x4_t(bv[jp], buf(stage, 1) + lm_off(r, ch));
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) hmma_bf16(acc[mt][nt], a[st][mt], bv[nt >> 1][(nt & 1) * 2], bv[nt >> 1][(nt & 1) * 2 + 1]);
        }
      __syncwarp();
      if (t + 2 < tiles) load(t + 2, stage);
      cp_async_commit();
    }
    cp_async_wait<0>();
    __syncwarp();
    // results: row d = mt*16 + hf*8 + g, columns e = nt*8 + 2tq, +1
    float* dst = wsp + pass * 1024;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int d = mt * 16 + hf * 8 + g;
        const float f = 1.f;   // unnormalised: the apply kernel divides by the merged row sum
        if (pass == 0) {
          const float zs = quad_sum(z[mt][hf]);
          if (tq == 0) { wsp[2048 + d] = m_run[mt][hf]; wsp[2048 + 32 + d] = zs; }
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
          *reinterpret_cast<float2*>(dst + d * 32 + nt * 8 + 2 * tq) = make_float2(acc[mt][nt][hf * 2] * f, acc[mt][nt][hf * 2 + 1] * f);
      }
  }
}

constexpr int LB_CHUNK = 128;   // tokens per CTA of the apply kernel

__global__ void __launch_bounds__(128) linattn_bwd_apply_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                const float* __restrict__ ws, bf16* __restrict__ dqkv, int N, int S) {
  // per warp: [stage][q|k|v|dout] tiles of 2 KB, then ctx and dctx as bf16 [d][e] tiles, then kmax | kzinv | cs
  extern __shared__ __align__(128) uint8_t lb_dyn[];
  constexpr int WARP_BYTES = 2 * 4 * LM_BUF + 2 * LM_BUF + 3 * 32 * 4;
  const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const bf16* base = qkv + (int64_t)b * N * 384 + h * LA_D;
  const bf16* dob = dout + (int64_t)b * N * 128 + h * LA_D;
  bf16* dbase = dqkv + (int64_t)b * N * 384 + h * LA_D;
  const float* wsp = ws + ((int64_t)b * 4 + h) * S * LB_WS;   // S partial results of the ctx pass
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(lb_dyn + (size_t)h * WARP_BYTES);
  auto buf = [&](int stage, int which) { return sbase + (uint32_t)((stage * 4 + which) * LM_BUF); };
  const uint32_t ctx_s = sbase + 8 * LM_BUF, dctx_s = ctx_s + LM_BUF;
  float* fl = reinterpret_cast<float*>(lb_dyn + (size_t)h * WARP_BYTES + 10 * LM_BUF);   // kmax[32] kzinv[32] cs[32]
  const float scale = 0.17677669529663687f;

  const int tok_begin = blockIdx.y * LB_CHUNK, tok_end = min(N, tok_begin + LB_CHUNK);
  const int steps = (tok_end - tok_begin) >> 4;
  const int tiles = (steps + 1) >> 1;
  auto load = [&](int tile, int stage) {
    const int tok0 = tok_begin + tile * LM_TILE;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = lane + 32 * i;                 // 0..511: which = c / 128 (q, k, v, dout)
      const int which = c >> 7, r = (c & 127) >> 2, ch = c & 3;
      if (tok0 + r < tok_end) {
        const bf16* src = which < 3 ? base + (int64_t)(tok0 + r) * 384 + 128 * which + ch * 8 : dob + (int64_t)(tok0 + r) * 128 + ch * 8;
        cp_async16(buf(stage, which) + lm_off(r, ch), src);
      }
    }
  };
  load(0, 0);
  cp_async_commit();
  if (tiles > 1) load(1, 1);
  cp_async_commit();

  // merge the S partials of row d = lane (softmax-over-tokens state: max m_s, sum z_s, accumulator acc_s):
  //   ctx[d][e] = sum_s acc_s[d][e] exp(m_s - m) / sum_s z_s exp(m_s - m),  dctx = sum_s dctx_s
  // then ctx, dctx -> bf16 tiles and cs[d] = sum_e dctx[d][e] ctx[d][e] in fp32
  {
    float mrow = -INFINITY;
    for (int sp = 0; sp < S; ++sp) mrow = fmaxf(mrow, wsp[(int64_t)sp * LB_WS + 2048 + lane]);
    float zt = 0.f;
    for (int sp = 0; sp < S; ++sp)
      zt += wsp[(int64_t)sp * LB_WS + 2048 + 32 + lane] * ex2f((wsp[(int64_t)sp * LB_WS + 2048 + lane] - mrow) * LOG2E);
    const float zinv = 1.0f / zt;
    float csd = 0.f;
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      float cv[8], dv8[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { cv[j] = 0.f; dv8[j] = 0.f; }
      for (int sp = 0; sp < S; ++sp) {
        const float* pw = wsp + (int64_t)sp * LB_WS;
        const float wgt = ex2f((pw[2048 + lane] - mrow) * LOG2E) * zinv;
#pragma unroll
        for (int j = 0; j < 8; j += 4) {
          const float4 c4 = *reinterpret_cast<const float4*>(pw + lane * 32 + c8 * 8 + j);
          const float4 d4 = *reinterpret_cast<const float4*>(pw + 1024 + lane * 32 + c8 * 8 + j);
          cv[j] = fmaf(c4.x, wgt, cv[j]); cv[j + 1] = fmaf(c4.y, wgt, cv[j + 1]);
          cv[j + 2] = fmaf(c4.z, wgt, cv[j + 2]); cv[j + 3] = fmaf(c4.w, wgt, cv[j + 3]);
          dv8[j] += d4.x; dv8[j + 1] += d4.y; dv8[j + 2] += d4.z; dv8[j + 3] += d4.w;
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) csd = fmaf(cv[j], dv8[j], csd);
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ctx_s + lm_off(lane, c8)), "r"(pack_bf2(cv[0], cv[1])),
                   "r"(pack_bf2(cv[2], cv[3])), "r"(pack_bf2(cv[4], cv[5])), "r"(pack_bf2(cv[6], cv[7])) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dctx_s + lm_off(lane, c8)), "r"(pack_bf2(dv8[0], dv8[1])),
                   "r"(pack_bf2(dv8[2], dv8[3])), "r"(pack_bf2(dv8[4], dv8[5])), "r"(pack_bf2(dv8[6], dv8[7])) : "memory");
    }
    fl[lane] = mrow;
    fl[32 + lane] = zinv;
    fl[64 + lane] = csd;
  }
  __syncwarp();
  // B fragments.  dqs / dks contract over e with B[k = e][n = d] = M[d][e]: plain ldmatrix on the [d][e] tile (rows = n);
  // dv contracts over d with B[k = d][n = e] = dctx[d][e]: transposing ldmatrix (rows = k), as the forward does for ctx.
  uint32_t bctx[2][2][4], bdc[2][2][4], bdct[2][2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int jp = 0; jp < 2; ++jp) {
      // matrices: (n rows jp*16 + 0-7, k chunk 2ks), (same rows, chunk 2ks+1), (n rows +8, chunk 2ks), (n rows +8, chunk 2ks+1)
      const int rn = jp * 16 + ((lane >> 4) << 3) + (lane & 7);
      const int chn = ks * 2 + ((lane >> 3) & 1);
      ldsm_x4(bctx[ks][jp], ctx_s + lm_off(rn, chn));
      ldsm_x4(bdc[ks][jp], dctx_s + lm_off(rn, chn));
      const int rk = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
      const int chk = jp * 2 + (lane >> 4);
      ldsm_x4_t(bdct[ks][jp], dctx_s + lm_off(rk, chk));
    }
  // per-column constants in C layout (cols nt*8 + 2tq, +1) and in A layout (k index ks*16 + 2tq (+1), +8 (+9))
  float2 kmC[4], kzC[4], csC[4], kmA[2][2], kzA[2][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int c = nt * 8 + 2 * tq;
    kmC[nt] = make_float2(fl[c] * LOG2E, fl[c + 1] * LOG2E);
    kzC[nt] = make_float2(fl[32 + c], fl[32 + c + 1]);
    csC[nt] = make_float2(fl[64 + c], fl[64 + c + 1]);
  }
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int c = ks * 16 + hi * 8 + 2 * tq;
      kmA[ks][hi] = make_float2(fl[c] * LOG2E, fl[c + 1] * LOG2E);
      kzA[ks][hi] = make_float2(fl[32 + c], fl[32 + c + 1]);
    }

  for (int t = 0; t < tiles; ++t) {
    cp_async_wait<1>();
    __syncwarp();
    const int stage = t & 1;
    const int nst = min(2, steps - 2 * t);
    for (int mi = 0; mi < nst; ++mi) {
      // A fragments (rows = tokens, k = channel): dout, v, and ks = exp(k - kmax) / Z
      uint32_t ado[2][4], av[2][4], ak[2][4];
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int r = mi * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int ch = ks * 2 + (lane >> 4);
        ldsm_x4(ado[ks], buf(stage, 3) + lm_off(r, ch));
        ldsm_x4(av[ks], buf(stage, 2) + lm_off(r, ch));
        ldsm_x4(ak[ks], buf(stage, 1) + lm_off(r, ch));
#pragma unroll
        for (int i = 0; i < 4; ++i) {             // regs 0,1: k cols 2tq,+1 (rows g, g+8); regs 2,3: k cols +8
          const int hi = i >> 1;
          float2 x = unpack_bf2(ak[ks][i]);
          x.x = ex2f(fmaf(x.x, LOG2E, -kmA[ks][hi].x)) * kzA[ks][hi].x;
          x.y = ex2f(fmaf(x.y, LOG2E, -kmA[ks][hi].y)) * kzA[ks][hi].y;
          ak[ks][i] = pack_bf2(x.x, x.y);
        }
      }
      float dqs[4][4], dks[4][4], dvv[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { dqs[nt][i] = 0.f; dks[nt][i] = 0.f; dvv[nt][i] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          // plain-ldmatrix B: n-tile nt = rows (nt>>1)*16 + (nt&1)*8: registers (nt&1)*2 (k lo), (nt&1)*2+1 (k hi)
          hmma_bf16(dqs[nt], ado[ks], bctx[ks][nt >> 1][(nt & 1) * 2], bctx[ks][nt >> 1][(nt & 1) * 2 + 1]);
          hmma_bf16(dks[nt], av[ks], bdc[ks][nt >> 1][(nt & 1) * 2], bdc[ks][nt >> 1][(nt & 1) * 2 + 1]);
          hmma_bf16(dvv[nt], ak[ks], bdct[ks][nt >> 1][(nt & 1) * 2], bdct[ks][nt >> 1][(nt & 1) * 2 + 1]);
        }
      }
      // elementwise part in C layout: rows g (regs 0,1) and g+8 (regs 2,3), cols nt*8 + 2tq, +1
      const int tok0 = tok_begin + t * LM_TILE + mi * 16;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int row = mi * 16 + hf * 8 + g;
        float2 qv[4], kv[4];
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          uint32_t uq, uk;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(uq) : "r"(buf(stage, 0) + lm_off(row, nt) + tq * 4));
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(uk) : "r"(buf(stage, 1) + lm_off(row, nt) + tq * 4));
          qv[nt] = unpack_bf2(uq);
          kv[nt] = unpack_bf2(uk);
          mx = fmaxf(mx, fmaxf(qv[nt].x, qv[nt].y));
        }
        mx = quad_max(mx) * LOG2E;
        float s = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          qv[nt].x = ex2f(fmaf(qv[nt].x, LOG2E, -mx)); qv[nt].y = ex2f(fmaf(qv[nt].y, LOG2E, -mx));
          s += qv[nt].x + qv[nt].y;
        }
        const float inv = scale / quad_sum(s);
        float inner = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          qv[nt].x *= inv; qv[nt].y *= inv;                       // qs
          inner = fmaf(qv[nt].x, dqs[nt][hf * 2], inner);
          inner = fmaf(qv[nt].y, dqs[nt][hf * 2 + 1], inner);
        }
        inner = quad_sum(inner) * (1.0f / scale);
        bf16* orow = dbase + (int64_t)(tok0 + hf * 8 + g) * 384 + 2 * tq;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float dq0 = qv[nt].x * (dqs[nt][hf * 2] - inner), dq1 = qv[nt].y * (dqs[nt][hf * 2 + 1] - inner);
          const float k0 = ex2f(fmaf(kv[nt].x, LOG2E, -kmC[nt].x)) * kzC[nt].x;
          const float k1 = ex2f(fmaf(kv[nt].y, LOG2E, -kmC[nt].y)) * kzC[nt].y;
          const float dk0 = k0 * (dks[nt][hf * 2] - csC[nt].x), dk1 = k1 * (dks[nt][hf * 2 + 1] - csC[nt].y);
          *reinterpret_cast<uint32_t*>(orow + nt * 8) = pack_bf2(dq0, dq1);
          *reinterpret_cast<uint32_t*>(orow + 128 + nt * 8) = pack_bf2(dk0, dk1);
          *reinterpret_cast<uint32_t*>(orow + 256 + nt * 8) = pack_bf2(dvv[nt][hf * 2], dvv[nt][hf * 2 + 1]);
        }
      }
    }
    __syncwarp();
    if (t + 2 < tiles) load(t + 2, stage);
    cp_async_commit();
  }
  cp_async_wait<0>();
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// LinearAttention with the to_qkv 1x1 convolution fused in (C = 64 input channels: the two full-resolution sites,
// where the 384-channel qkv tensor is 0.4 GB per pass -- 150 us to write at HBM write speed and 80 us to read back).
//   xn [B][N][64] (the PreNorm output) -> out [B][N][128];  wqkv [384][64] bf16 (row = output channel, no bias)
// One CTA per sample, one warp per head, private cp.async double buffer of x tiles per warp (x is tiny: 128 KB per
// sample, the 4x re-read comes from L2).  All GEMMs on mma.sync m16n8k16 with operand roles chosen so that every
// accumulator fragment is directly the next GEMM's operand fragment (no shared-memory round trips in the loop):
//   K^T[d x tok] = Wk[d x c] X^T[c x tok]   (A = Wk rows, in registers; B = x tile via ldmatrix)
//   V^T[e x tok] = Wv[e x c] X^T
//   P^T = exp(K^T - rowmax) (online, fp32)   -> C fragment == A fragment of  ctx[d x e] += P^T[d x tok] V[tok x e]
//                                               and the V^T C fragment     == B fragment of the same MMA
//   Q[tok x d] = X[tok x c] Wq^T[c x d]      -> softmax over d inside a quad -> C fragment == A fragment of
//   out[tok x e] = q~[tok x d] ctx[d x e]    (ctx through shared memory once per sample: it needs a transpose)
namespace {

// FOLD: x is the raw PreNorm input; GroupNorm(1,C) is applied algebraically:  W (r (x - mu) gamma + beta)
//   = r (W diag(gamma)) x  +  (W beta - r mu rowsum(W diag(gamma)))  -- one scale r and one per-channel constant.
//   K: the constant is uniform along the softmax (token) axis and cancels; V: it passes through the normalised
//   ctx unchanged (sum_n softmax = 1); Q: added before the softmax over d.
template <bool FOLD>
#ifndef LINATTN_FUSED_MIN_CTAS
#define LINATTN_FUSED_MIN_CTAS 3   // 4 (128 registers, spills) measured 4 % slower: the kernel is issue-bound, not tail-bound
#endif
__global__ void __launch_bounds__(128, LINATTN_FUSED_MIN_CTAS) linattn_qkv_fused_kernel(const bf16* __restrict__ xn, int ldx,
                                                               const bf16* __restrict__ wqkv,
                                                               const float* __restrict__ uv,
                                                               const float2* __restrict__ gn_part, int gn_splits,
                                                               float gn_eps, bf16* __restrict__ out, int N) {
  constexpr int C = 64;
  constexpr int XROW = 128;              // bytes per token row of the x tile (64 bf16)
  constexpr int XBUF = LM_TILE * XROW;   // 4 KB
  __shared__ __align__(128) uint8_t smem[4][2][XBUF];   // per warp: two x-tile stages
  __shared__ __align__(128) uint8_t smem_aux[4][3072];  // per warp: ctx (2 KB) + 16-token output staging tile (1 KB)
  const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const bf16* xb = xn + (int64_t)b * N * ldx;
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&smem[h][0][0]);
  const uint32_t ctx_s = (uint32_t)__cvta_generic_to_shared(&smem_aux[h][0]);
  const uint32_t stg = ctx_s + 2048;
  // 16-byte chunk c (0..7) of row r in a [rows][128 B] tile, XOR-swizzled (conflict-free ldmatrix)
  auto xoff = [](int r, int c) { return (uint32_t)(r * XROW + ((c ^ (r & 7)) << 4)); };
  auto load_x = [&](int tile, int stage) {
    const int tok0 = tile * LM_TILE;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i, r = c >> 3, ch = c & 7;
      if (tok0 + r < N) cp_async16(sbase + stage * XBUF + xoff(r, ch), xb + (int64_t)(tok0 + r) * ldx + ch * 8);
    }
  };
  const int steps = N >> 4, tiles = (steps + 1) >> 1;
  pdl_wait();
  pdl_trigger();
  load_x(0, 0);
  cp_async_commit();
  if (tiles > 1) load_x(1, 1);
  cp_async_commit();

  // per-sample GroupNorm(1, C) statistics from the slab partials (fixed order), see gn_apply_kernel
  float gr = 1.f, gmu = 0.f;
  if (FOLD) {
    if (gn_splits < 0) {
      // un-pivoted sums {S, Q} left by the producing convolution's epilogue (conv_epilogue.cuh, mode 1): -gn_splits slots
      double a = 0.0, bsum = 0.0;
      for (int sp = 0; sp < -gn_splits; ++sp) {
        const float2 v = gn_part[(int64_t)b * (-gn_splits) + sp];
        a += (double)v.x; bsum += (double)v.y;
      }
      const double cnt = (double)N * (double)C;
      const double m1 = a / cnt;
      double var = bsum / cnt - m1 * m1;
      if (var < 0.0) var = 0.0;
      gmu = (float)m1;
      gr = (float)(1.0 / sqrt(var + (double)gn_eps));
    } else {
      float a = 0.f, bsum = 0.f;
      for (int sp = 0; sp < gn_splits; ++sp) {
        const float2 v = gn_part[(int64_t)b * gn_splits + sp];
        a += v.x; bsum += v.y;
      }
      const float K = __bfloat162float(xb[0]);
      const float inv_n = 1.0f / ((float)N * (float)C);
      const float m1 = a * inv_n;
      const float var = fmaxf(bsum * inv_n - m1 * m1, 0.f);
      gmu = K + m1;
      gr = 1.0f / sqrtf(var + gn_eps);
    }
  }
  const float kscale = gr * LOG2E;   // softmax over tokens of r*k: the scale folds into the exp2 argument

  // A fragments of a 32 x 64 weight block (rows = d or e): frag[mt][ks] = {(g, 2tq), (g+8, 2tq), (g, 2tq+8), (g+8, 2tq+8)}
  auto load_w_a = [&](const bf16* w, uint32_t (&f)[2][4][4]) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const bf16* p0 = w + (mt * 16 + g) * C + ks * 16 + 2 * tq;
        f[mt][ks][0] = *reinterpret_cast<const uint32_t*>(p0);
        f[mt][ks][1] = *reinterpret_cast<const uint32_t*>(p0 + 8 * C);
        f[mt][ks][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
        f[mt][ks][3] = *reinterpret_cast<const uint32_t*>(p0 + 8 * C + 8);
      }
  };
  uint32_t wk[2][4][4], wv[2][4][4];
  load_w_a(wqkv + (int64_t)(128 + h * LA_D) * C, wk);
  load_w_a(wqkv + (int64_t)(256 + h * LA_D) * C, wv);

  float acc[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  float m_run[2][2], z[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) { m_run[mt][hf] = -INFINITY; z[mt][hf] = 0.f; }

  // ---------------- phase 1
  for (int t = 0; t < tiles; ++t) {
    cp_async_wait<1>();
    __syncwarp();
    const uint32_t xs = sbase + (t & 1) * XBUF;
    const int nst = min(2, steps - 2 * t);
    for (int st = 0; st < nst; ++st) {
      // B fragments of X^T for the 16 tokens of this step: xf[ks] = {b0,b1 of tokens 0-7, b0,b1 of tokens 8-15}
      uint32_t xf[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = st * 16 + ((lane >> 4) << 3) + (lane & 7);
        const int ch = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(xf[ks], xs + xoff(r, ch));
      }
      float kc[2][2][4], vc[2][2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { kc[mt][nt][i] = 0.f; vc[mt][nt][i] = 0.f; }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            hmma_bf16(kc[mt][nt], wk[mt][ks], xf[ks][2 * nt], xf[ks][2 * nt + 1]);
            hmma_bf16(vc[mt][nt], wv[mt][ks], xf[ks][2 * nt], xf[ks][2 * nt + 1]);
          }
        }
      // online softmax over tokens, per channel row d = mt*16 + hf*8 + g
      uint32_t pa[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          float mx = fmaxf(fmaxf(kc[mt][0][2 * hf], kc[mt][0][2 * hf + 1]), fmaxf(kc[mt][1][2 * hf], kc[mt][1][2 * hf + 1]));
          mx = quad_max(mx);
          const float nm = fmaxf(m_run[mt][hf], mx);
          const float sc = ex2f((m_run[mt][hf] - nm) * kscale);
          m_run[mt][hf] = nm;
          const float ml2 = nm * kscale;
          const float p00 = ex2f(fmaf(kc[mt][0][2 * hf], kscale, -ml2)), p01 = ex2f(fmaf(kc[mt][0][2 * hf + 1], kscale, -ml2));
          const float p10 = ex2f(fmaf(kc[mt][1][2 * hf], kscale, -ml2)), p11 = ex2f(fmaf(kc[mt][1][2 * hf + 1], kscale, -ml2));
          z[mt][hf] = z[mt][hf] * sc + (p00 + p01 + p10 + p11);
          pa[mt][hf] = pack_bf2(p00, p01);        // a0 / a1: tokens 2tq,2tq+1
          pa[mt][hf + 2] = pack_bf2(p10, p11);    // a2 / a3: tokens 8+2tq, 8+2tq+1
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) { acc[mt][nt][2 * hf] *= sc; acc[mt][nt][2 * hf + 1] *= sc; }
        }
      // ctx[d x e] += P^T[d x tok] V[tok x e]:  B fragment of n-tile j (e = 8j + g) comes from V^T rows 8j + g
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int mv = j >> 1, hf = j & 1;
        const uint32_t b0 = pack_bf2(vc[mv][0][2 * hf], vc[mv][0][2 * hf + 1]);
        const uint32_t b1 = pack_bf2(vc[mv][1][2 * hf], vc[mv][1][2 * hf + 1]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) hmma_bf16(acc[mt][j], pa[mt], b0, b1);
      }
    }
    __syncwarp();
    if (t + 2 < tiles) load_x(t + 2, t & 1);
    cp_async_commit();
  }
  cp_async_wait<0>();
  __syncwarp();

  // ---------------- ctx -> bf16 in shared memory (64-byte rows, lm_off swizzle), scaled by 32^-1/2 / Z[d]
  // (phase 2's first x tiles are requested now, so their latency hides behind the ctx hand-over)
  load_x(0, 0);
  cp_async_commit();
  if (tiles > 1) load_x(1, 1);
  cp_async_commit();
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float zinv = 1.0f / quad_sum(z[mt][hf]);
      const float f = 0.17677669529663687f * zinv * gr;     // 32^-1/2 (q scale) / Z[d] * GroupNorm scale of v
      const int d = mt * 16 + hf * 8 + g;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float c0v = acc[mt][nt][hf * 2] * f, c1v = acc[mt][nt][hf * 2 + 1] * f;
        if (FOLD) {  // v's per-channel constant passes through the normalised ctx (sum_n softmax_n(k) = 1)
          const int e = 256 + h * LA_D + nt * 8 + 2 * tq;
          c0v += 0.17677669529663687f * (uv[384 + e] - gr * gmu * uv[e]);
          c1v += 0.17677669529663687f * (uv[384 + e + 1] - gr * gmu * uv[e + 1]);
        }
        const uint32_t v = pack_bf2(c0v, c1v);
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ctx_s + lm_off(d, nt) + tq * 4), "r"(v) : "memory");
      }
    }
  __syncwarp();
  uint32_t bc[2][2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int jp = 0; jp < 2; ++jp) {
      const int r = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
      const int ch = jp * 2 + (lane >> 4);
      ldsm_x4_t(bc[ks][jp], ctx_s + lm_off(r, ch));
    }
  __syncwarp();
  // q's per-channel constants (columns d = 8nt + 2tq, +1 of this thread's C fragment)
  float qoff[4][2];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int o = h * LA_D + nt * 8 + 2 * tq + u;
      qoff[nt][u] = FOLD ? uv[384 + o] - gr * gmu * uv[o] : 0.f;
    }
  // B fragments of Wq^T (k = c, n = d): wq[ks][nt] = {Wq[d = 8nt + g][c = 16ks + 2tq..], [.. + 8]}
  uint32_t wq[4][4][2];
  {
    const bf16* w = wqkv + (int64_t)(h * LA_D) * C;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const bf16* p0 = w + (nt * 8 + g) * C + ks * 16 + 2 * tq;
        wq[ks][nt][0] = *reinterpret_cast<const uint32_t*>(p0);
        wq[ks][nt][1] = *reinterpret_cast<const uint32_t*>(p0 + 8);
      }
  }

  // ---------------- phase 2: the same double-buffered x stream, now as the A operand of Q = X Wq^T
  bf16* obase = out + (int64_t)b * N * 128 + h * LA_D;
  for (int t = 0; t < tiles; ++t) {
    cp_async_wait<1>();
    __syncwarp();
    const uint32_t xs = sbase + (t & 1) * XBUF;
    const int nst = min(2, steps - 2 * t);
    for (int mi = 0; mi < nst; ++mi) {
      uint32_t xa[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        // A fragments of X: a0 (tok 0-7, chunk 2ks), a1 (tok 8-15, 2ks), a2 (tok 0-7, 2ks+1), a3 (tok 8-15, 2ks+1)
        const int r = mi * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int ch = ks * 2 + (lane >> 4);
        ldsm_x4(xa[ks], xs + xoff(r, ch));
      }
      float qc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        qc[nt][0] = qc[nt][1] = qc[nt][2] = qc[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) hmma_bf16(qc[nt], xa[ks], wq[ks][nt][0], wq[ks][nt][1]);
      }
      // softmax over d (32 columns of a token row): rows g (regs 0,1) and g+8 (regs 2,3)
      uint32_t qa[2][4];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (FOLD) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            qc[nt][2 * hf] = fmaf(qc[nt][2 * hf], gr, qoff[nt][0]);
            qc[nt][2 * hf + 1] = fmaf(qc[nt][2 * hf + 1], gr, qoff[nt][1]);
          }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mx = fmaxf(mx, fmaxf(qc[nt][2 * hf], qc[nt][2 * hf + 1]));
        mx = quad_max(mx) * LOG2E;
        float s = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          qc[nt][2 * hf] = ex2f(fmaf(qc[nt][2 * hf], LOG2E, -mx));
          qc[nt][2 * hf + 1] = ex2f(fmaf(qc[nt][2 * hf + 1], LOG2E, -mx));
          s += qc[nt][2 * hf] + qc[nt][2 * hf + 1];
        }
        const float inv = 1.0f / quad_sum(s);
        // A fragment of q~ for k-step ks (d in [16ks, 16ks+16)): n-tiles 2ks (cols 2tq..) and 2ks+1 (cols 8+2tq..)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          qa[ks][hf] = pack_bf2(qc[2 * ks][2 * hf] * inv, qc[2 * ks][2 * hf + 1] * inv);
          qa[ks][hf + 2] = pack_bf2(qc[2 * ks + 1][2 * hf] * inv, qc[2 * ks + 1][2 * hf + 1] * inv);
        }
      }
      float o[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) hmma_bf16(o[nt], qa[ks], bc[ks][nt >> 1][(nt & 1) * 2], bc[ks][nt >> 1][(nt & 1) * 2 + 1]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg + lm_off(g, nt) + tq * 4), "r"(pack_bf2(o[nt][0], o[nt][1])) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(stg + lm_off(g + 8, nt) + tq * 4), "r"(pack_bf2(o[nt][2], o[nt][3])) : "memory");
      }
      __syncwarp();
      const int tok0 = t * LM_TILE + mi * 16;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = lane + 32 * i, r = c >> 2, ch = c & 3;
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(stg + lm_off(r, ch)));
        *reinterpret_cast<uint4*>(obase + (int64_t)(tok0 + r) * 128 + ch * 8) = v;
      }
      __syncwarp();
    }
    if (t + 2 < tiles) load_x(t + 2, t & 1);
    cp_async_commit();
  }
  cp_async_wait<0>();
}

}  // namespace

int k_linear_attention_qkv(const void* xn, int ldx, int cin, const void* wqkv, void* out, int batch, int n_tokens,
                           int dtype, cudaStream_t st) {
  LDM_REQUIRE(dtype == LDM_DT_BF16 && cin == 64 && n_tokens % 16 == 0 && ldx % 8 == 0,
              "linear_attention_qkv: needs bf16, 64 input channels and a multiple of 16 tokens");
  if (batch == 0 || n_tokens == 0) return 0;
  LDM_CUDA(ldm_launch_pdl(linattn_qkv_fused_kernel<false>, dim3(batch), dim3(128), 0, st, (const bf16*)xn, ldx, (const bf16*)wqkv,
                          (const float*)nullptr, (const float2*)nullptr, 0, 0.f, (bf16*)out, n_tokens));
  LDM_LAUNCHED("linear_attention_qkv");
  return 0;
}
int k_linear_attention_qkv_prenorm(const void* x, int ldx, int cin, const void* wfold, const float* uv, const void* gn_part,
                                   int gn_splits, float eps, void* out, int batch, int n_tokens, int dtype, cudaStream_t st) {
  LDM_REQUIRE(dtype == LDM_DT_BF16 && cin == 64 && n_tokens % 16 == 0 && ldx % 8 == 0 && uv && gn_part && gn_splits != 0,
              "linear_attention_qkv_prenorm: needs bf16, 64 input channels, a multiple of 16 tokens and GroupNorm statistics");
  if (batch == 0 || n_tokens == 0) return 0;
  LDM_CUDA(ldm_launch_pdl(linattn_qkv_fused_kernel<true>, dim3(batch), dim3(128), 0, st, (const bf16*)x, ldx, (const bf16*)wfold, uv,
                          (const float2*)gn_part, gn_splits, eps, (bf16*)out, n_tokens));
  LDM_LAUNCHED("linear_attention_qkv_prenorm");
  return 0;
}

// wfold[o][c] = bf16(w[o][c] gamma[c]);  uv[o] = sum_c float(wfold[o][c]);  uv[384 + o] = sum_c w[o][c] beta[c]
__global__ void fold_prenorm_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    int cin, bf16* __restrict__ wfold, float* __restrict__ uv) {
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;   // one warp per output row
  if (o >= 384) return;
  float u = 0.f, v = 0.f;
  for (int c = lane; c < cin; c += 32) {
    const float wv = w[(int64_t)o * cin + c];
    const bf16 f = __float2bfloat16_rn(wv * gamma[c]);
    wfold[(int64_t)o * cin + c] = f;
    u += __bfloat162float(f);
    v = fmaf(wv, beta[c], v);
  }
  u = warp_sum(u);
  v = warp_sum(v);
  if (lane == 0) { uv[o] = u; uv[384 + o] = v; }
}
int k_fold_prenorm_qkv(const float* wqkv, const float* gamma, const float* beta, int cin, void* wfold, float* uv,
                       cudaStream_t st) {
  fold_prenorm_kernel<<<96, 128, 0, st>>>(wqkv, gamma, beta, cin, (bf16*)wfold, uv);
  LDM_LAUNCHED("fold_prenorm_qkv");
  return 0;
}
bool k_linear_attention_qkv_applicable(int cin, int n_tokens, int dtype) {
  return dtype == LDM_DT_BF16 && cin == 64 && n_tokens % 16 == 0 && getenv("LDM_LINATTN_UNFUSED") == nullptr;
}

int k_linear_attention(const void* qkv, void* out, int batch, int n_tokens, int dtype, cudaStream_t st) {
  if (batch == 0 || n_tokens == 0) return 0;
  int grid = batch * LA_HEADS;
  if (dtype == LDM_DT_BF16 && n_tokens % 16 == 0 && getenv("LDM_LINATTN_SIMT") == nullptr)
    LDM_CUDA(ldm_launch_pdl(linattn_mma_kernel<false>, dim3(batch), dim3(128), 0, st, (const bf16*)qkv, (bf16*)out, n_tokens, LaAffine{}));
  else if (dtype == LDM_DT_BF16) linattn_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)qkv, (bf16*)out, n_tokens);
  else linattn_kernel<float><<<grid, 256, 0, st>>>((const float*)qkv, (float*)out, n_tokens);
  LDM_LAUNCHED("linear_attention");
  return 0;
}

// LinearAttention core on the RAW to_qkv projection of the un-normalised block input (see LaAffine): bf16, N % 16 == 0
int k_linear_attention_prenorm_core(const void* qkv_raw, void* out, const float* uv, const void* gn_part, int gn_splits, float eps,
                                    const void* x, int ldx, int cin, int batch, int n_tokens, cudaStream_t st) {
  LDM_REQUIRE(n_tokens % 16 == 0 && uv && gn_part && gn_splits != 0, "linear_attention_prenorm_core: needs N %% 16 == 0 and statistics");
  if (batch == 0 || n_tokens == 0) return 0;
  LaAffine af;
  af.uv = uv; af.gn_part = (const float2*)gn_part; af.gn_splits = gn_splits; af.gn_eps = eps; af.x = (const bf16*)x; af.ldx = ldx; af.cin = cin;
  LDM_CUDA(ldm_launch_pdl(linattn_mma_kernel<true>, dim3(batch), dim3(128), 0, st, (const bf16*)qkv_raw, (bf16*)out, n_tokens, af));
  LDM_LAUNCHED("linear_attention_prenorm_core");
  return 0;
}

// ---- full softmax attention, one thread per query token (N <= 256; N = 4 in the reference config)
template <typename T>
__global__ void attn_kernel(const T* __restrict__ qkv, T* __restrict__ out, int N) {
  extern __shared__ __align__(16) float sm[];  // k[N][LA_LD], v[N][LA_LD]
  pdl_wait();
  pdl_trigger();
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
    LDM_CUDA(ldm_launch_pdl(attn_kernel<bf16>, dim3(grid), dim3(threads), smem, st, (const bf16*)qkv, (bf16*)out, n_tokens));
  } else {
    if (smem > 48 * 1024)
      LDM_CUDA(cudaFuncSetAttribute(attn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_kernel<float><<<grid, threads, smem, st>>>((const float*)qkv, (float*)out, n_tokens);
  }
  LDM_LAUNCHED("attention");
  return 0;
}


// ---- LinearAttention backward on mma.sync (bf16, N % 16 == 0); workspace: k_linear_attention_backward_ws_bytes
// token splits of the ctx pass: enough CTAs for two per SM at small batch, never more than 8
static int linattn_bwd_splits(int batch) {
  int s = batch > 0 ? 296 / batch : 1;
  return s < 1 ? 1 : (s > 8 ? 8 : s);
}
int64_t k_linear_attention_backward_ws_bytes(int batch) {
  return (int64_t)batch * 4 * linattn_bwd_splits(batch) * LB_WS * sizeof(float) + 256;
}
bool k_linear_attention_backward_mma_applicable(int n_tokens, int dtype) {
  return dtype == LDM_DT_BF16 && n_tokens % 16 == 0 && getenv("LDM_LINATTN_BWD_SIMT") == nullptr;
}
int k_linear_attention_backward_mma(const void* qkv, const void* dout, void* dqkv, int batch, int N, void* workspace,
                                    cudaStream_t st) {
  LDM_REQUIRE(workspace && ((uintptr_t)workspace & 15) == 0, "linear_attention_backward: workspace missing or unaligned");
  if (batch == 0 || N == 0) return 0;
  constexpr int smem = 4 * (2 * 4 * LM_BUF + 2 * LM_BUF + 3 * 32 * 4);
  static bool attr_set[64] = {};   // the dynamic shared-memory opt-in is per device
  int dev = 0;
  LDM_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    LDM_CUDA(cudaFuncSetAttribute(linattn_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set[dev & 63] = true;
  }
  int S = linattn_bwd_splits(batch);
  if (S > N / 64) S = N / 64 > 0 ? N / 64 : 1;              // at least 64 tokens per split
  const int chunk = ((N + S - 1) / S + 31) / 32 * 32;       // whole 32-token tiles
  S = (N + chunk - 1) / chunk;
  linattn_bwd_ctx_kernel<<<dim3(batch, S), 128, 0, st>>>((const bf16*)qkv, (const bf16*)dout, (float*)workspace, N, chunk);
  LDM_LAUNCHED("linattn_bwd_ctx");
  linattn_bwd_apply_kernel<<<dim3(batch, (N + LB_CHUNK - 1) / LB_CHUNK), 128, smem, st>>>((const bf16*)qkv, (const bf16*)dout,
                                                                                      (const float*)workspace, (bf16*)dqkv, N, S);
  LDM_LAUNCHED("linattn_bwd_apply");
  return 0;
}
