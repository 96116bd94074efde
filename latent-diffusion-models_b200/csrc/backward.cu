// Backward kernels of the training step (src/DiffusionModelTrainer.py:36-67: loss.backward() through UNet.forward).
// The reference gets these from autograd over ATen/cuDNN; here each forward kernel has a
// hand-written counterpart, launched from torch.autograd.Function wrappers (ldm_b200/train.py).  First version:
// FFMA kernels templated on the storage type (fp32 parity path and bf16), fp32 accumulation, fp32 parameter grads.
//   conv wgrad          dW[co][ci][kh][kw] = sum_m dY[m][co] X[m @ tap][ci]        (split over M, atomicAdd)
//   conv dgrad          = the forward implicit-GEMM kernel on dY with the transposed, tap-flipped filter (pack below)
//   GroupNorm(+SiLU)    dx, dgamma, dbeta, d(rowvec)   one CTA per sample, statistics recomputed
//   MaxPool / ConvTranspose(k2 s2) / LinearAttention / Attention / initial & final conv / time-embedding MLPs
#include "kernels.h"

namespace {

template <typename T>
__device__ __forceinline__ float ldf(const T* p) { return to_float(*p); }
template <typename T>
__device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) { load_chunk(p, v); }
template <>
__device__ __forceinline__ void ld4<bf16>(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T>
__device__ __forceinline__ void st4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) { store_chunk(p, v); }
template <>
__device__ __forceinline__ void st4<bf16>(bf16* p, const float (&v)[4]) {
  uint2 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// ------------------------------------------------------------------ conv wgrad
// grid (cout/64, taps*cin/64, msplit); 256 threads, 4x4 register tiles; dW is OIHW fp32, accumulated with atomics.
template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const T* __restrict__ x, int ldx, int cin, const T* __restrict__ dy, int lddy, int cout,
                  float* __restrict__ dw, float* __restrict__ dbias, int B, int H, int W, int ksize, int m_per_split) {
  __shared__ __align__(16) float As[16][68];  // dY tile  [m][co]
  __shared__ __align__(16) float Bs[16][68];  // X  tile  [m][ci]
  const int tid = threadIdx.x;
  const int co0 = blockIdx.x * 64;
  const int cblocks = cin / 64;
  const int tap = blockIdx.y / cblocks, ci0 = (blockIdx.y % cblocks) * 64;
  const int taps = ksize * ksize, pad = ksize / 2;
  const int dyy = taps == 9 ? tap / 3 - pad : 0, dxx = taps == 9 ? tap % 3 - pad : 0;
  const int M = B * H * W;
  const int m_begin = blockIdx.z * m_per_split, m_end = min(M, m_begin + m_per_split);
  const int lrow = tid >> 4, lc = (tid & 15) * 4;  // loader: 16 rows x 16 four-element chunks
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // the bias gradient (column sums of dY) rides along in the CTAs of tap 0 / channel block 0: they stream every dY row
  // of their (co tile, M split) exactly once
  const bool do_bias = dbias != nullptr && blockIdx.y == 0;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int m0 = m_begin; m0 < m_end; m0 += 16) {
    const int m = m0 + lrow;
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < m_end) {
      ld4<T>(dy + (int64_t)m * lddy + co0 + lc, av);
      if (do_bias) { bsum[0] += av[0]; bsum[1] += av[1]; bsum[2] += av[2]; bsum[3] += av[3]; }
      const int w_ = m % W, r_ = m / W, h_ = r_ % H, n_ = r_ / H;
      const int hh = h_ + dyy, ww = w_ + dxx;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) ld4<T>(x + ((int64_t)(n_ * H + hh) * W + ww) * ldx + ci0 + lc, bv);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) { As[lrow][lc + i] = av[i]; Bs[lrow][lc + i] = bv[i]; }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + ty * 4 + i, ci = ci0 + tx * 4 + j;
      atomicAdd(dw + ((int64_t)co * cin + ci) * taps + tap, acc[i][j]);
    }
  if (do_bias) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) As[lrow][lc + i] = bsum[i];
    __syncthreads();
    if (tid < 64) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 16; ++r) t += As[r][tid];
      atomicAdd(dbias + co0 + tid, t);
    }
  }
}

// column sums: out[c] += sum_m a[m][c]   (bias gradients)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ a, int lda, float* __restrict__ out, int M, int C, int rows_per_block) {
  const int m0 = blockIdx.x * rows_per_block, m1 = min(M, m0 + rows_per_block);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int m = m0; m < m1; ++m) s += ldf(a + (int64_t)m * lda + c);
    atomicAdd(out + c, s);
  }
}

// bf16 variant for channel counts that are multiples of 8: 16-byte loads, a thread owns one 8-channel chunk and walks the
// rows of its CTA's slab; shared-memory reduction over the row lanes, then one atomic per channel and CTA
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const bf16* __restrict__ a, int lda, float* __restrict__ out, int M, int C, int rows_per_block) {
  __shared__ float red[256][9];
  const int cpp = C / 8, ppi = 256 / cpp;
  const int ci = threadIdx.x % cpp, pl = threadIdx.x / cpp;
  const int m0 = blockIdx.x * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (pl < ppi) {
    const bf16* base = a + ci * 8;
    for (int m = m0 + pl; m < m1; m += ppi) {
      const uint4 r = *reinterpret_cast<const uint4*>(base + (int64_t)m * lda);
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h2[i]);
        acc[2 * i] += f.x; acc[2 * i + 1] += f.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {   // one thread per channel sums the row lanes
    float t = 0.f;
    for (int k = 0; k < ppi; ++k) t += red[k * cpp + (c >> 3)][c & 7];
    atomicAdd(out + c, t);
  }
}

// dgrad filter: out[ci][tap'][co] = w[co][ci][2-kh][2-kw] (3x3) / w[co][ci] (1x1), in the conv kernels' packed layout
template <typename T>
__global__ void pack_dgrad_kernel(const float* __restrict__ w, int cout, int cin, int ksize, T* __restrict__ out) {
  const int taps = ksize * ksize;
  const int64_t total = (int64_t)cin * taps * cout;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = (int)(i % cout);
  const int tp = (int)((i / cout) % taps);
  const int ci = (int)(i / ((int64_t)cout * taps));
  const int src_tap = taps - 1 - tp;  // (2-kh)*3 + (2-kw)
  out[i] = from_float<T>(w[((int64_t)co * cin + ci) * taps + src_tap]);
}

// both packed forms in one launch: the first half of the grid writes the forward filter [cout][tap*cin + ci], the second
// half the dgrad filter above (one thread per output element; a shared-memory tiled variant measured slower: the filters
// are small and thousands of independent CTAs hide the strided reads better than 256 looping ones)
template <typename T>
__global__ void pack_pair_kernel(const float* __restrict__ w, int cout, int cin, int ksize, T* __restrict__ fwd, T* __restrict__ dgrad,
                                 int half_blocks) {
  const int taps = ksize * ksize;
  const int64_t total = (int64_t)cin * taps * cout;
  const bool second = (int)blockIdx.x >= half_blocks;
  const int64_t i = (int64_t)(blockIdx.x - (second ? half_blocks : 0)) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (!second) {
    const int k = (int)(i % (taps * cin)), o = (int)(i / (taps * cin));
    const int tap = k / cin, c = k % cin;
    fwd[i] = from_float<T>(w[((int64_t)o * cin + c) * taps + tap]);
  } else {
    const int co = (int)(i % cout);
    const int tp = (int)((i / cout) % taps);
    const int ci = (int)(i / ((int64_t)cout * taps));
    dgrad[i] = from_float<T>(w[((int64_t)co * cin + ci) * taps + (taps - 1 - tp)]);
  }
}

// ------------------------------------------------------------------ GroupNorm (+SiLU) backward, one CTA per sample
template <typename T>
__global__ void __launch_bounds__(1024)
gn_backward_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ dy, int lddy, const float* __restrict__ gamma,
                   const float* __restrict__ beta, const float* __restrict__ rowvec, int ld_rowvec, T* __restrict__ dx,
                   int lddx, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ drowvec,
                   int ld_drowvec, int HW, int C, int G, float eps, int silu) {
  constexpr int V = 4;
  __shared__ float part[1024];
  __shared__ float gsum[32];
  __shared__ float s_mean[32], s_rstd[32], s_s1[32], s_s2[32];
  const int n = blockIdx.x;
  const int cpp = C / V, cpg = cpp / G;
  const int ci = threadIdx.x % cpp, pl = threadIdx.x / cpp, ppi = blockDim.x / cpp;
  const int g = ci / cpg;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const T* xb = x + (int64_t)n * HW * ldx + ci * V;
  const T* dyb = dy + (int64_t)n * HW * lddy + ci * V;
  float rv[V], ga[V], be[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    rv[i] = rowvec ? rowvec[(int64_t)n * ld_rowvec + ci * V + i] : 0.f;
    ga[i] = gamma[ci * V + i];
    be[i] = beta[ci * V + i];
  }
  auto greduce = [&](float v, float* dst) {  // per-group block sum -> dst[g]
    part[threadIdx.x] = v;
    __syncthreads();
    const int members = ppi * cpg;
    for (int gg = warp; gg < G; gg += nwarps) {
      float s = 0.f;
      for (int idx = lane; idx < members; idx += 32) s += part[(idx / cpg) * cpp + gg * cpg + idx % cpg];
      s = warp_sum(s);
      if (lane == 0) dst[gg] = s;
    }
    __syncthreads();
  };
  const float inv_n = 1.0f / ((float)HW * (float)(cpg * V));
  // statistics (exact two-pass)
  float s = 0.f;
  for (int p = pl; p < HW; p += ppi) {
    float w[V];
    ld4<T>(xb + (int64_t)p * ldx, w);
#pragma unroll
    for (int i = 0; i < V; ++i) s += w[i] + rv[i];
  }
  greduce(s, gsum);
  if (threadIdx.x < G) s_mean[threadIdx.x] = gsum[threadIdx.x] * inv_n;
  __syncthreads();
  const float mean = s_mean[g];
  float q = 0.f;
  for (int p = pl; p < HW; p += ppi) {
    float w[V];
    ld4<T>(xb + (int64_t)p * ldx, w);
#pragma unroll
    for (int i = 0; i < V; ++i) { const float d = w[i] + rv[i] - mean; q = fmaf(d, d, q); }
  }
  greduce(q, gsum);
  if (threadIdx.x < G) s_rstd[threadIdx.x] = 1.0f / sqrtf(gsum[threadIdx.x] * inv_n + eps);
  __syncthreads();
  const float rstd = s_rstd[g];
  // group sums of dyhat and dyhat*xhat; per-channel dgamma / dbeta
  float s1 = 0.f, s2 = 0.f, dg[V], db[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { dg[i] = 0.f; db[i] = 0.f; }
  for (int p = pl; p < HW; p += ppi) {
    float w[V], d[V];
    ld4<T>(xb + (int64_t)p * ldx, w);
    ld4<T>(dyb + (int64_t)p * lddy, d);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xh = (w[i] + rv[i] - mean) * rstd;
      float dz = d[i];
      if (silu) {
        const float z = fmaf(xh, ga[i], be[i]);
        const float sg = sigmoid_acc(z);
        dz *= sg * (1.0f + z * (1.0f - sg));
      }
      dg[i] = fmaf(dz, xh, dg[i]);
      db[i] += dz;
      const float dyh = dz * ga[i];
      s1 += dyh;
      s2 = fmaf(dyh, xh, s2);
    }
  }
  greduce(s1, s_s1);
  greduce(s2, s_s2);
  // per-channel parameter gradients: reduce over the threads that share this channel chunk (pl = 0..ppi-1)
#pragma unroll
  for (int i = 0; i < V; ++i) {
    part[threadIdx.x] = dg[i];
    __syncthreads();
    if (pl == 0) {
      float t = 0.f;
      for (int k = 0; k < ppi; ++k) t += part[k * cpp + ci];
      atomicAdd(dgamma + ci * V + i, t);
    }
    __syncthreads();
    part[threadIdx.x] = db[i];
    __syncthreads();
    if (pl == 0) {
      float t = 0.f;
      for (int k = 0; k < ppi; ++k) t += part[k * cpp + ci];
      atomicAdd(dbeta + ci * V + i, t);
    }
    __syncthreads();
  }
  const float m1 = s_s1[g] * inv_n, m2 = s_s2[g] * inv_n;
  float drv[V];
#pragma unroll
  for (int i = 0; i < V; ++i) drv[i] = 0.f;
  T* dxb = dx + (int64_t)n * HW * lddx + ci * V;
  for (int p = pl; p < HW; p += ppi) {
    float w[V], d[V], o[V];
    ld4<T>(xb + (int64_t)p * ldx, w);
    ld4<T>(dyb + (int64_t)p * lddy, d);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xh = (w[i] + rv[i] - mean) * rstd;
      float dz = d[i];
      if (silu) {
        const float z = fmaf(xh, ga[i], be[i]);
        const float sg = sigmoid_acc(z);
        dz *= sg * (1.0f + z * (1.0f - sg));
      }
      o[i] = rstd * (dz * ga[i] - m1 - xh * m2);
      drv[i] += o[i];
    }
    st4<T>(dxb + (int64_t)p * lddx, o);
  }
  if (drowvec) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      part[threadIdx.x] = drv[i];
      __syncthreads();
      if (pl == 0) {
        float t = 0.f;
        for (int k = 0; k < ppi; ++k) t += part[k * cpp + ci];
        drowvec[(int64_t)n * ld_drowvec + ci * V + i] = t;
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------ MaxPool2d(2,2) backward (first maximum wins)
template <typename T>
__global__ void maxpool2_bwd_kernel(const T* __restrict__ x, int ldx, const T* __restrict__ dy, int lddy, T* __restrict__ dx,
                                    int lddx, int H, int W, int C, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  int64_t r = i / C;
  const int wo = (int)(r % (W / 2)); r /= (W / 2);
  const int ho = (int)(r % (H / 2));
  const int64_t n = r / (H / 2);
  const T* xp = x + ((n * H + 2 * ho) * W + 2 * wo) * (int64_t)ldx + c;
  float v[4] = {ldf(xp), ldf(xp + ldx), ldf(xp + (int64_t)W * ldx), ldf(xp + (int64_t)(W + 1) * ldx)};
  int best = 0;
#pragma unroll
  for (int k = 1; k < 4; ++k)
    if (v[k] > v[best]) best = k;
  const float g = ldf(dy + ((n * (H / 2) + ho) * (W / 2) + wo) * (int64_t)lddy + c);
  T* dp = dx + ((n * H + 2 * ho) * W + 2 * wo) * (int64_t)lddx + c;
  dp[0] = from_float<T>(best == 0 ? g : 0.f);
  dp[lddx] = from_float<T>(best == 1 ? g : 0.f);
  dp[(int64_t)W * lddx] = from_float<T>(best == 2 ? g : 0.f);
  dp[(int64_t)(W + 1) * lddx] = from_float<T>(best == 3 ? g : 0.f);
}

// dyq[n,h,w,q*C + c] = dy[n, 2h + q/2, 2w + q%2, c]   (ConvTranspose k2 s2 backward gather)
template <typename T>
__global__ void unshuffle2_kernel(const T* __restrict__ dy, int lddy, T* __restrict__ out, int H, int W, int C, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  int64_t r = i / C;
  const int q = (int)(r % 4); r /= 4;
  const int w_ = (int)(r % W); r /= W;
  const int h_ = (int)(r % H);
  const int64_t n = r / H;
  out[i] = dy[((n * 2 * H + 2 * h_ + (q >> 1)) * (2 * W) + 2 * w_ + (q & 1)) * (int64_t)lddy + c];
}

// ------------------------------------------------------------------ LinearAttention core backward, CTA per (b, head)
// forward: qs = softmax_d(q) * s ; ks = softmax_n(k) ; ctx = ks^T v ; out = qs ctx      (src/UNet.py:149-163)
// backward: dctx = qs^T dout ; dqs = dout ctx^T ; dks = v dctx^T ; dv = ks dctx ; then the two softmax backwards.
// Every product is a 32x32x32 GEMM per 32-token tile; each is computed by 64 threads with 4x4 register blocks from
// shared-memory operands stored in the orientation that makes both fragments one 16-byte load.
#define LB_D 32
#define LB_TILE 32
#define LB_LD 36
struct LbSmem {
  float q_r[LB_TILE][LB_LD], k_r[LB_TILE][LB_LD], v_r[LB_TILE][LB_LD], do_r[LB_TILE][LB_LD];   // [token][channel]
  float kT[LB_D][LB_LD], vT[LB_D][LB_LD], doT[LB_D][LB_LD];                                     // [channel][token]
  float ctx[LB_D][LB_LD], ctxT[LB_D][LB_LD], dctx[LB_D][LB_LD], dctxT[LB_D][LB_LD];
  float o_dqs[LB_TILE][LB_LD], o_dks[LB_TILE][LB_LD], o_dv[LB_TILE][LB_LD];
  float kmax[LB_D], kzinv[LB_D], colsum[LB_D], red[8][LB_D];
};
// out[r0..r0+3][c0..c0+3] += sum_k AT[k][r0..] * B[k][c0..]   (AT, B: [32][LB_LD])
__device__ __forceinline__ void lb_gemm4x4(const float (*AT)[LB_LD], const float (*B)[LB_LD], int r0, int c0, float (&acc)[4][4]) {
#pragma unroll 8
  for (int k = 0; k < 32; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(&AT[k][r0]);
    const float4 b = *reinterpret_cast<const float4*>(&B[k][c0]);
    const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
  }
}
template <typename T>
__global__ void __launch_bounds__(256)
linattn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, T* __restrict__ dqkv, int N) {
  extern __shared__ __align__(16) uint8_t lb_raw[];
  LbSmem& S = *reinterpret_cast<LbSmem*>(lb_raw);
  const float scale = 0.17677669529663687f;
  const int b = blockIdx.x / 4, h = blockIdx.x % 4;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* base = qkv + (int64_t)b * N * 384 + h * LB_D;
  const T* dob = dout + (int64_t)b * N * 128 + h * LB_D;
  T* dbase = dqkv + (int64_t)b * N * 384 + h * LB_D;
  // ---- softmax-over-tokens statistics of k
  float m = -INFINITY;
  for (int n = warp; n < N; n += 8) m = fmaxf(m, ldf(base + (int64_t)n * 384 + 128 + lane));
  S.red[warp][lane] = m;
  __syncthreads();
  if (warp == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, S.red[w][lane]);
    S.kmax[lane] = m;
  }
  __syncthreads();
  float zs = 0.f;
  for (int n = warp; n < N; n += 8) zs += expf(ldf(base + (int64_t)n * 384 + 128 + lane) - S.kmax[lane]);
  __syncthreads();
  S.red[warp][lane] = zs;
  __syncthreads();
  if (warp == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += S.red[w][lane];
    S.kzinv[lane] = 1.0f / t;
    S.colsum[lane] = 0.f;
  }
  __syncthreads();
  // tile loader: q_r <- softmax_d(q)*s, k_r/kT <- softmax_n(k), v_r/vT <- v, do_r/doT <- dout   (rows >= N are zero)
  auto load_tile = [&](int n0) {
    for (int idx = tid; idx < LB_TILE * LB_D; idx += 256) {
      const int r = idx / LB_D, c = idx % LB_D, n = n0 + r;
      float qv = 0.f, kv = 0.f, vv = 0.f, dv = 0.f;
      if (n < N) {
        qv = ldf(base + (int64_t)n * 384 + c);
        kv = expf(ldf(base + (int64_t)n * 384 + 128 + c) - S.kmax[c]) * S.kzinv[c];
        vv = ldf(base + (int64_t)n * 384 + 256 + c);
        dv = ldf(dob + (int64_t)n * 128 + c);
      }
      S.q_r[r][c] = qv;
      S.k_r[r][c] = kv; S.kT[c][r] = kv;
      S.v_r[r][c] = vv; S.vT[c][r] = vv;
      S.do_r[r][c] = dv; S.doT[c][r] = dv;
    }
    __syncthreads();
    for (int r = warp; r < LB_TILE; r += 8) {   // softmax over d for each row of q_r
      const float v = S.q_r[r][lane];
      const float mx = warp_max(v);
      const float e = expf(v - mx);
      const float sum = warp_sum(e);
      S.q_r[r][lane] = (n0 + r < N) ? e / sum * scale : 0.f;
    }
    __syncthreads();
  };
  const int grp = tid >> 6, gt = tid & 63;           // four groups of 64 threads
  const int r0 = (gt >> 3) * 4, c0 = (gt & 7) * 4;   // 4x4 block of a 32x32 output
  // ---- sweep 1: ctx = ks^T v (group 0), dctx = qs^T dout (group 1)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int n0 = 0; n0 < N; n0 += LB_TILE) {
    load_tile(n0);
    if (grp == 0) lb_gemm4x4(S.k_r, S.v_r, r0, c0, acc);         // out[d][e] += sum_t ks[t][d] v[t][e]
    else if (grp == 1) lb_gemm4x4(S.q_r, S.do_r, r0, c0, acc);   // out[d][e] += sum_t qs[t][d] dout[t][e]
    __syncthreads();
  }
  if (grp < 2) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (grp == 0) { S.ctx[r0 + i][c0 + j] = acc[i][j]; S.ctxT[c0 + j][r0 + i] = acc[i][j]; }
        else { S.dctx[r0 + i][c0 + j] = acc[i][j]; S.dctxT[c0 + j][r0 + i] = acc[i][j]; }
      }
  }
  __syncthreads();
  // ---- sweep 2: per token tile  dqs = dout ctx^T (group 0), dks = v dctx^T (group 1), dv = ks dctx (group 2)
  const int tr = tid >> 3, part = tid & 7;  // epilogue mapping: 32 rows x 8 parts of 4 channels
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  for (int n0 = 0; n0 < N; n0 += LB_TILE) {
    load_tile(n0);
    if (grp < 3) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      float (*O)[LB_LD] = grp == 0 ? S.o_dqs : (grp == 1 ? S.o_dks : S.o_dv);
      if (grp == 0) lb_gemm4x4(S.doT, S.ctxT, r0, c0, acc);        // out[t][d] = sum_e dout[t][e] ctx[d][e]
      else if (grp == 1) lb_gemm4x4(S.vT, S.dctxT, r0, c0, acc);   // out[t][d] = sum_e v[t][e] dctx[d][e]
      else lb_gemm4x4(S.kT, S.dctx, r0, c0, acc);                  // out[t][e] = sum_d ks[t][d] dctx[d][e]
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(&O[r0 + i][c0]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    __syncthreads();
    const int n = n0 + tr;
    float dqs[4], dks[4], dvv[4], qv[4], kv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dqs[j] = S.o_dqs[tr][part * 4 + j]; dks[j] = S.o_dks[tr][part * 4 + j]; dvv[j] = S.o_dv[tr][part * 4 + j];
      qv[j] = S.q_r[tr][part * 4 + j]; kv[j] = S.k_r[tr][part * 4 + j];
    }
    float inner = 0.f;   // <qs, dqs> over the 32 channels of the row: 8 consecutive threads
#pragma unroll
    for (int j = 0; j < 4; ++j) inner = fmaf(qv[j], dqs[j], inner);
    inner += __shfl_xor_sync(0xffffffffu, inner, 1);
    inner += __shfl_xor_sync(0xffffffffu, inner, 2);
    inner += __shfl_xor_sync(0xffffffffu, inner, 4);
    if (n < N) {
      float dq[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dq[j] = qv[j] * (dqs[j] - inner / scale);   // qs = s*qhat: dq = qs * (dqs - <qs,dqs>/s)
        cs[j] = fmaf(kv[j], dks[j], cs[j]);
      }
      st4<T>(dbase + (int64_t)n * 384 + part * 4, dq);
      st4<T>(dbase + (int64_t)n * 384 + 128 + part * 4, dks);   // dks parked in the dk slot until colsum is known
      st4<T>(dbase + (int64_t)n * 384 + 256 + part * 4, dvv);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) atomicAdd(&S.colsum[part * 4 + j], cs[j]);   // colsum[d] = sum_n ks[n][d] dks[n][d]
  __syncthreads();
  // ---- dk = ks * (dks - colsum)
  for (int idx = tid; idx < N * LB_D; idx += 256) {
    const int n = idx / LB_D, c = idx % LB_D;
    const float ks = expf(ldf(base + (int64_t)n * 384 + 128 + c) - S.kmax[c]) * S.kzinv[c];
    T* p = dbase + (int64_t)n * 384 + 128 + c;
    *p = from_float<T>(ks * (to_float(*p) - S.colsum[c]));
  }
}

// ------------------------------------------------------------------ full attention backward (N <= 64), CTA per (b, head)
template <typename T>
__global__ void __launch_bounds__(64)
attn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, T* __restrict__ dqkv, int N) {
  extern __shared__ float sm[];
  float* sq = sm;                  // [N][33]  q * scale
  float* sk = sq + N * 33;
  float* sv = sk + N * 33;
  float* sd = sv + N * 33;         // dout
  float* sp = sd + N * 33;         // P  [N][N+1]
  float* sds = sp + N * (N + 1);   // dS [N][N+1]
  const float scale = 0.17677669529663687f;
  const int b = blockIdx.x / 4, h = blockIdx.x % 4;
  const T* base = qkv + (int64_t)b * N * 384 + h * 32;
  const T* dob = dout + (int64_t)b * N * 128 + h * 32;
  T* dbase = dqkv + (int64_t)b * N * 384 + h * 32;
  for (int idx = threadIdx.x; idx < N * 32; idx += blockDim.x) {
    const int n = idx >> 5, c = idx & 31;
    sq[n * 33 + c] = ldf(base + (int64_t)n * 384 + c) * scale;
    sk[n * 33 + c] = ldf(base + (int64_t)n * 384 + 128 + c);
    sv[n * 33 + c] = ldf(base + (int64_t)n * 384 + 256 + c);
    sd[n * 33 + c] = ldf(dob + (int64_t)n * 128 + c);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float mx = -INFINITY;
    for (int j = 0; j < N; ++j) {
      float s = 0.f;
      for (int c = 0; c < 32; ++c) s = fmaf(sq[i * 33 + c], sk[j * 33 + c], s);
      sp[i * (N + 1) + j] = s;
      mx = fmaxf(mx, s);
    }
    float den = 0.f;
    for (int j = 0; j < N; ++j) { const float e = expf(sp[i * (N + 1) + j] - mx); sp[i * (N + 1) + j] = e; den += e; }
    const float inv = 1.0f / den;
    float inner = 0.f;
    for (int j = 0; j < N; ++j) {
      const float pj = sp[i * (N + 1) + j] * inv;
      sp[i * (N + 1) + j] = pj;
      float dp = 0.f;
      for (int c = 0; c < 32; ++c) dp = fmaf(sd[i * 33 + c], sv[j * 33 + c], dp);
      sds[i * (N + 1) + j] = dp;
      inner = fmaf(pj, dp, inner);
    }
    for (int j = 0; j < N; ++j) sds[i * (N + 1) + j] = sp[i * (N + 1) + j] * (sds[i * (N + 1) + j] - inner);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    for (int c0 = 0; c0 < 32; c0 += 4) {
      float dq[4] = {0.f, 0.f, 0.f, 0.f}, dk[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
      for (int j = 0; j < N; ++j) {
        const float dsij = sds[i * (N + 1) + j], dsji = sds[j * (N + 1) + i], pji = sp[j * (N + 1) + i];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          dq[u] = fmaf(dsij, sk[j * 33 + c0 + u], dq[u]);   // dq_i = scale * sum_j dS_ij k_j
          dk[u] = fmaf(dsji, sq[j * 33 + c0 + u], dk[u]);   // dk_i = sum_j dS_ji (q_j*scale)
          dv[u] = fmaf(pji, sd[j * 33 + c0 + u], dv[u]);    // dv_i = sum_j P_ji dout_j
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) dq[u] *= scale;
      st4<T>(dbase + (int64_t)i * 384 + c0, dq);
      st4<T>(dbase + (int64_t)i * 384 + 128 + c0, dk);
      st4<T>(dbase + (int64_t)i * 384 + 256 + c0, dv);
    }
  }
}

// ------------------------------------------------------------------ initial conv wgrad (fp32 NCHW input, tiny Cin)
// grid (msplit), 256 threads = 4 pixel lanes x 64 output channels; every thread keeps the 9*CIN partial sums of its
// channel for its pixels (the input patch values are warp-uniform broadcast loads), then one shared-memory reduction
// over the pixel lanes and 9*CIN*64 atomics per CTA.  Also accumulates the bias gradient.
template <typename T, int CIN>
__global__ void __launch_bounds__(256)
initial_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, float* __restrict__ dbias,
                     float* __restrict__ part, int B, int cout, int H, int W, int m_per_split) {
  __shared__ float red[4][64];
  const int M = B * H * W;
  const int m0 = blockIdx.x * m_per_split, m1 = min(M, m0 + m_per_split);
  const int pl = threadIdx.x >> 6, cl = threadIdx.x & 63;
  for (int c0 = 0; c0 < cout; c0 += 64) {
    float acc[9 * CIN], bs = 0.f;
#pragma unroll
    for (int j = 0; j < 9 * CIN; ++j) acc[j] = 0.f;
    for (int m = m0 + pl; m < m1; m += 4) {
      const int w_ = m % W, r_ = m / W, h_ = r_ % H, n_ = r_ / H;
      const float g = ldf(dy + (int64_t)m * cout + c0 + cl);
      bs += g;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int hh = h_ + tap / 3 - 1, ww = w_ + tap % 3 - 1;
        const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float xv = ok ? __ldg(x + ((int64_t)(n_ * CIN + ci) * H + hh) * W + ww) : 0.f;
          acc[tap * CIN + ci] = fmaf(xv, g, acc[tap * CIN + ci]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j <= 9 * CIN; ++j) {
      red[pl][cl] = j < 9 * CIN ? acc[j < 9 * CIN ? j : 0] : bs;
      __syncthreads();
      if (pl == 0) {
        const float t = red[0][cl] + red[1][cl] + red[2][cl] + red[3][cl];
        const int64_t widx = ((int64_t)(c0 + cl) * CIN + (j % CIN)) * 9 + (j / CIN);
        if (part) {   // this CTA's row of partials [dW (OIHW order) | db]; summed by rows_sum_f32_kernel: a thousand CTAs
          //            doing atomics on the same 1.8 k floats serialise in L2
          float* row = part + (int64_t)blockIdx.x * ((int64_t)cout * CIN * 9 + cout);
          if (j < 9 * CIN) row[widx] = t; else row[(int64_t)cout * CIN * 9 + c0 + cl] = t;
        } else if (j < 9 * CIN) atomicAdd(dw + widx, t);
        else if (dbias) atomicAdd(dbias + c0 + cl, t);
      }
      __syncthreads();
    }
  }
}

// out0[c] += sum_r part[r][c] (c < n0), out1[c - n0] += ... (c >= n0).  grid (ceil(cols / 32)), 256 threads = 8 row lanes x
// 32 columns, fixed summation order
__global__ void __launch_bounds__(256)
rows_sum_f32_kernel(const float* __restrict__ part, int rows, int cols, float* __restrict__ out0, int n0, float* __restrict__ out1) {
  __shared__ float red[8][33];
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  float a = 0.f;
  if (col < cols)
    for (int r = rl; r < rows; r += 8) a += part[(int64_t)r * cols + col];
  red[rl][threadIdx.x & 31] = a;
  __syncthreads();
  if (rl == 0 && col < cols) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    if (col < n0) out0[col] += t; else if (out1) out1[col - n0] += t;
  }
}

// ------------------------------------------------------------------ final 1x1 conv backward (NCHW fp32 dout -> NHWC dx)
template <typename T>
__global__ void final_conv_dx_kernel(const float* __restrict__ dout, const float* __restrict__ w, T* __restrict__ dx, int cin,
                                     int cout, int HW, int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // pixel
  if (i >= total) return;
  const int64_t b = i / HW;
  const int p = (int)(i % HW);
  float g[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) g[o] = o < cout ? dout[(b * cout + o) * HW + p] : 0.f;
  for (int c0 = 0; c0 < cin; c0 += 4) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int o = 0; o < 8; ++o)
      if (o < cout) {
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = fmaf(g[o], w[o * cin + c0 + u], v[u]);
      }
    st4<T>(dx + i * cin + c0, v);
  }
}
// dw[o][c] += sum_pix dout[b][o][p] x[pix][c];  db[o] += sum dout.   grid (msplit), 256 threads = pixel lanes x channel
// lanes (cpad = channels rounded up to a power of two): x rows are read coalesced, dout is a warp-uniform broadcast.
template <typename T>
__global__ void __launch_bounds__(256)
final_conv_dw_kernel(const float* __restrict__ dout, const T* __restrict__ x, int ldx, float* __restrict__ dw,
                     float* __restrict__ db, int cin, int cout, int HW, int total, int m_per_split, int cpad) {
  __shared__ float red[256];
  const int m0 = blockIdx.x * m_per_split, m1 = min(total, m0 + m_per_split);
  const int cl = threadIdx.x % cpad, pl = threadIdx.x / cpad, lanes = 256 / cpad;
  for (int c = cl; c < ((cin + cpad - 1) / cpad) * cpad; c += cpad) {
    const bool live = c < cin;
    for (int o0 = 0; o0 < cout; o0 += 4) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f}, ds[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int m = m0 + pl; m < m1; m += lanes) {
        const float xv = live ? ldf(x + (int64_t)m * ldx + c) : 0.f;
        const int b = m / HW, p = m - b * HW;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (o0 + j < cout) {
            const float d = __ldg(dout + ((int64_t)b * cout + o0 + j) * HW + p);
            acc[j] = fmaf(d, xv, acc[j]);
            ds[j] += d;
          }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // 4 weight-gradient columns, then the 4 bias sums (channel lane 0 only)
        red[threadIdx.x] = j < 4 ? acc[j & 3] : ds[j & 3];
        __syncthreads();
        if (pl == 0 && o0 + (j & 3) < cout) {
          float t = 0.f;
          for (int l = 0; l < lanes; ++l) t += red[l * cpad + cl];
          if (j < 4) { if (live) atomicAdd(dw + (o0 + j) * cin + c, t); }
          else if (c == 0) atomicAdd(db + o0 + (j & 3), t);
        }
        __syncthreads();
      }
    }
  }
}

// ------------------------------------------------------------------ small fp32 GEMM with arbitrary strides (time path)
// C[i][j] (+)= sum_k A(i,k) B(k,j);  A(i,k) = a[i*a_rs + k*a_cs], B(k,j) = b[k*b_rs + j*b_cs]
__global__ void gemm_f32_kernel(const float* __restrict__ a, int64_t a_rs, int64_t a_cs, const float* __restrict__ b,
                                int64_t b_rs, int64_t b_cs, float* __restrict__ c, int64_t c_rs, int M, int N, int K,
                                int accumulate) {
  __shared__ float As[16][17], Bs[16][17];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int i = blockIdx.y * 16 + ty, j = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    As[ty][tx] = (i < M && k0 + tx < K) ? a[i * a_rs + (k0 + tx) * a_cs] : 0.f;
    Bs[ty][tx] = (k0 + ty < K && j < N) ? b[(k0 + ty) * b_rs + j * b_cs] : 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) acc = fmaf(As[ty][k], Bs[k][tx], acc);
    __syncthreads();
  }
  if (i < M && j < N) c[i * c_rs + j] = (accumulate ? c[i * c_rs + j] : 0.f) + acc;
}

// elementwise helpers of the time path
__global__ void sinusoid_kernel(const int64_t* __restrict__ t, float* __restrict__ emb, int batch, int Din) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * Din) return;
  const int b = i / Din, k = i % Din, half = Din / 2;
  const float neg = -(float)(9.210340371976184 / (double)(half - 1));
  const int fi = k < half ? k : k - half;
  const float arg = (float)t[b] * expf((float)fi * neg);
  emb[i] = k < half ? sinf(arg) : cosf(arg);
}
// mode 0: y = x + bias[col]; 1: y = gelu(x + bias); 2: y = silu(x); 3: y = dy * gelu'(x); 4: y = dy * silu'(x); 5: y = a + b
__global__ void ew_kernel(const float* __restrict__ x, const float* __restrict__ other, float* __restrict__ y, int rows,
                          int cols, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int c = i % cols;
  const float v = x[i];
  float o = v;
  switch (mode) {
    case 0: o = v + other[c]; break;
    case 1: { const float u = v + other[c]; o = 0.5f * u * (1.0f + erff(u * 0.70710678118654752440f)); break; }
    case 2: o = v * sigmoid_acc(v); break;
    case 3: {  // other = dy, x = pre-activation
      const float cdf = 0.5f * (1.0f + erff(v * 0.70710678118654752440f));
      const float pdf = 0.3989422804014327f * expf(-0.5f * v * v);
      o = other[i] * (cdf + v * pdf);
      break;
    }
    case 4: { const float sg = sigmoid_acc(v); o = other[i] * sg * (1.0f + v * (1.0f - sg)); break; }
    case 5: o = v + other[i]; break;
  }
  y[i] = o;
}
// out[idx[b] or idx[0]][:] += src[b][:]   (label embedding gradient) ; temb[b][:] += table[idx][:] (forward, sign = +1)
__global__ void rows_scatter_add_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int idx_len,
                                        float* __restrict__ table, int batch, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * D) return;
  const int b = i / D, k = i % D;
  const int64_t r = idx_len == 1 ? idx[0] : idx[b];
  atomicAdd(table + r * D + k, src[i]);
}
__global__ void rows_gather_add_kernel(float* __restrict__ dst, const int64_t* __restrict__ idx, int idx_len,
                                       const float* __restrict__ table, int batch, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= batch * D) return;
  const int b = i / D, k = i % D;
  const int64_t r = idx_len == 1 ? idx[0] : idx[b];
  dst[i] += table[r * D + k];
}

template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = from_float<T>(to_float(a[i]) + to_float(b[i]));
}
// dst[r][0..C) = src[r][0..C) with independent row strides (channel concat / split)
template <typename T>
__global__ void copy_channels_kernel(const T* __restrict__ src, int lds, T* __restrict__ dst, int ldd, int C, int64_t rows) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const int64_t r = i / C;
  const int c = (int)(i % C);
  dst[r * ldd + c] = src[r * lds + c];
}

int split_for(int64_t M, int per_block_min, int target_blocks) {
  int64_t per = (M + target_blocks - 1) / target_blocks;
  if (per < per_block_min) per = per_block_min;
  per = (per + 15) / 16 * 16;
  return (int)per;
}

}  // namespace

#define DISPATCH_T(dtype, ...)                    \
  do {                                            \
    if ((dtype) == LDM_DT_BF16) { using T = bf16; __VA_ARGS__; } \
    else { using T = float; __VA_ARGS__; }        \
  } while (0)

int k_conv_wgrad(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw, float* dbias, int batch,
                 int H, int W, int ksize, int dtype, cudaStream_t st) {
  LDM_REQUIRE(cin % 64 == 0 && cout % 64 == 0, "conv_wgrad: channel counts (%d -> %d) must be multiples of 64", cin, cout);
  LDM_REQUIRE(ksize == 1 || ksize == 3, "conv_wgrad: kernel size %d unsupported", ksize);
  const int M = batch * H * W;
  if (M == 0) return 0;
  const int taps = ksize * ksize;
  const int base_blocks = (cout / 64) * (cin / 64) * taps;
  int target = (4 * 148 + base_blocks - 1) / base_blocks;
  if (target < 1) target = 1;
  const int mps = split_for(M, 256, target);
  const dim3 grid(cout / 64, taps * (cin / 64), (M + mps - 1) / mps);
  DISPATCH_T(dtype, conv_wgrad_kernel<T><<<grid, 256, 0, st>>>((const T*)x, ldx, cin, (const T*)dy, lddy, cout, dw, dbias, batch, H, W, ksize, mps));
  LDM_LAUNCHED("conv_wgrad");
  return 0;
}

int k_colsum(const void* a, int lda, float* out, int M, int C, int dtype, cudaStream_t st) {
  if (M == 0 || C == 0) return 0;
  const int rpb = split_for(M, 64, 4 * 148);
  if (dtype == LDM_DT_BF16 && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0 && lda % 8 == 0 && ((uintptr_t)a & 15) == 0) {
    const int rpb2 = split_for(M, 256, 2 * 148);
    colsum_bf16_kernel<<<(M + rpb2 - 1) / rpb2, 256, 0, st>>>((const bf16*)a, lda, out, M, C, rpb2);
    LDM_LAUNCHED("colsum");
    return 0;
  }
  DISPATCH_T(dtype, colsum_kernel<T><<<(M + rpb - 1) / rpb, 256, 0, st>>>((const T*)a, lda, out, M, C, rpb));
  LDM_LAUNCHED("colsum");
  return 0;
}

int k_pack_dgrad_weight(const float* w_oihw, int cout, int cin, int ksize, void* out, int dtype, cudaStream_t st) {
  const int64_t total = (int64_t)cin * ksize * ksize * cout;
  if (total == 0) return 0;
  DISPATCH_T(dtype, pack_dgrad_kernel<T><<<(int)ceil_div64(total, 256), 256, 0, st>>>(w_oihw, cout, cin, ksize, (T*)out));
  LDM_LAUNCHED("pack_dgrad_weight");
  return 0;
}

int k_pack_conv_weight_pair(const float* w_oihw, int cout, int cin, int ksize, void* fwd, void* dgrad, int dtype, cudaStream_t st) {
  const int64_t total = (int64_t)cin * ksize * ksize * cout;
  if (total == 0) return 0;
  const int half = (int)ceil_div64(total, 256);
  DISPATCH_T(dtype, pack_pair_kernel<T><<<2 * half, 256, 0, st>>>(w_oihw, cout, cin, ksize, (T*)fwd, (T*)dgrad, half));
  LDM_LAUNCHED("pack_conv_weight_pair");
  return 0;
}

int k_group_norm_backward(const void* x, int ldx, const void* dy, int lddy, const float* gamma, const float* beta,
                          const float* rowvec, int ld_rowvec, void* dx, int lddx, float* dgamma, float* dbeta,
                          float* drowvec, int ld_drowvec, int batch, int hw, int channels, int groups, float eps, int silu,
                          int dtype, cudaStream_t st) {
  LDM_REQUIRE(groups >= 1 && groups <= 32 && channels % groups == 0 && (channels / groups) % 4 == 0,
              "group_norm_backward: channels/groups (%d/%d) must be a multiple of 4", channels, groups);
  LDM_REQUIRE(ldx % 4 == 0 && lddy % 4 == 0 && lddx % 4 == 0, "group_norm_backward: unaligned stride");
  if (batch == 0 || hw == 0) return 0;
  const int cpp = channels / 4;
  LDM_REQUIRE(cpp <= 1024, "group_norm_backward: too many channels");
  int ppi = 1024 / cpp;
  if (ppi > hw) ppi = hw;
  int threads = cpp * ppi;
  // whole warps: pad the block with idle pixel lanes is not possible with this mapping, so shrink ppi until cpp*ppi % 32 == 0
  while (ppi > 1 && (cpp * ppi) % 32 != 0) --ppi;
  threads = cpp * ppi;
  LDM_REQUIRE(threads % 32 == 0, "group_norm_backward: %d channels do not map to whole warps", channels);
  DISPATCH_T(dtype, gn_backward_kernel<T><<<batch, threads, 0, st>>>((const T*)x, ldx, (const T*)dy, lddy, gamma, beta, rowvec,
                                                                     ld_rowvec, (T*)dx, lddx, dgamma, dbeta, drowvec, ld_drowvec,
                                                                     hw, channels, groups, eps, silu));
  LDM_LAUNCHED("group_norm_backward");
  return 0;
}

int k_maxpool2_backward(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int batch, int H, int W, int C,
                        int dtype, cudaStream_t st) {
  const int64_t total = (int64_t)batch * (H / 2) * (W / 2) * C;
  if (total == 0) return 0;
  DISPATCH_T(dtype, maxpool2_bwd_kernel<T><<<(int)ceil_div64(total, 256), 256, 0, st>>>((const T*)x, ldx, (const T*)dy, lddy, (T*)dx, lddx, H, W, C, total));
  LDM_LAUNCHED("maxpool2_backward");
  return 0;
}

int k_unshuffle2(const void* dy, int lddy, void* out, int batch, int H, int W, int C, int dtype, cudaStream_t st) {
  const int64_t total = (int64_t)batch * H * W * 4 * C;
  if (total == 0) return 0;
  DISPATCH_T(dtype, unshuffle2_kernel<T><<<(int)ceil_div64(total, 256), 256, 0, st>>>((const T*)dy, lddy, (T*)out, H, W, C, total));
  LDM_LAUNCHED("unshuffle2");
  return 0;
}

int k_linear_attention_backward(const void* qkv, const void* dout, void* dqkv, int batch, int N, int dtype, cudaStream_t st) {
  if (batch == 0 || N == 0) return 0;
  static bool attr_set[64] = {};   // the dynamic shared-memory opt-in is per device
  int dev = 0;
  LDM_CUDA(cudaGetDevice(&dev));
  if (!attr_set[dev & 63]) {
    LDM_CUDA(cudaFuncSetAttribute(linattn_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LbSmem)));
    LDM_CUDA(cudaFuncSetAttribute(linattn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LbSmem)));
    attr_set[dev & 63] = true;
  }
  DISPATCH_T(dtype, linattn_bwd_kernel<T><<<batch * 4, 256, sizeof(LbSmem), st>>>((const T*)qkv, (const T*)dout, (T*)dqkv, N));
  LDM_LAUNCHED("linear_attention_backward");
  return 0;
}

int k_attention_backward(const void* qkv, const void* dout, void* dqkv, int batch, int N, int dtype, cudaStream_t st) {
  LDM_REQUIRE(N <= 64, "attention_backward: %d tokens > 64 not supported", N);
  if (batch == 0 || N == 0) return 0;
  const size_t smem = (size_t)(4 * N * 33 + 2 * N * (N + 1)) * sizeof(float);
  if (dtype == LDM_DT_BF16) {
    if (smem > 48 * 1024) LDM_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_kernel<bf16><<<batch * 4, 64, smem, st>>>((const bf16*)qkv, (const bf16*)dout, (bf16*)dqkv, N);
  } else {
    if (smem > 48 * 1024) LDM_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attn_bwd_kernel<float><<<batch * 4, 64, smem, st>>>((const float*)qkv, (const float*)dout, (float*)dqkv, N);
  }
  LDM_LAUNCHED("attention_backward");
  return 0;
}

// ---- initial-conv weight gradient on the tensor cores: 3x3 patches of the fp32 NCHW image as a 64-channel bf16 NHWC
// tensor (channel j = ci*9 + tap, the OIHW order; zero beyond 9*CIN), then the 1x1 weight-gradient GEMM
template <int CIN>
__global__ void __launch_bounds__(256)
im2col3x3_small_kernel(const float* __restrict__ x, bf16* __restrict__ patches, int H, int W, int64_t total_chunks) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (pixel, 8-channel chunk)
  if (i >= total_chunks) return;
  const int chunk = (int)(i & 7);
  const int64_t m = i >> 3;
  const int w_ = (int)(m % W);
  const int64_t r_ = m / W;
  const int h_ = (int)(r_ % H);
  const int64_t n_ = r_ / H;
  float v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int j = chunk * 8 + u;
    float val = 0.f;
    if (j < 9 * CIN) {
      const int ci = j / 9, tap = j - ci * 9;
      const int hh = h_ + tap / 3 - 1, ww = w_ + tap % 3 - 1;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) val = __ldg(x + ((n_ * CIN + ci) * H + hh) * W + ww);
    }
    v[u] = val;
  }
  store_chunk(patches + m * 64 + chunk * 8, v);
}
__global__ void add_cols_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int rows, int cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  dst[i] += src[(i / cols) * ld_src + i % cols];
}

static int initial_wgrad_grid(int batch, int H, int W, int* mps_out) {
  const int M = batch * H * W;
  const int mps = split_for(M, 64, 8 * 148);
  if (mps_out) *mps_out = mps;
  return (M + mps - 1) / mps;
}
int64_t k_initial_conv_wgrad_scratch_bytes(int batch, int cin, int cout, int H, int W) {
  if (batch * H * W == 0) return 0;
  const int64_t rows = (int64_t)initial_wgrad_grid(batch, H, W, nullptr) * ((int64_t)cout * cin * 9 + cout) * sizeof(float);
  const int64_t tc = (int64_t)batch * H * W * 64 * 2 + (int64_t)cout * 64 * 4 + 512;   // patches + [cout][64] accumulator
  return rows > tc ? rows : tc;
}
int k_initial_conv_wgrad(const float* x, const void* dy, float* dw, float* dbias, int batch, int cin, int cout, int H, int W,
                         int dtype, void* scratch, cudaStream_t st) {
  LDM_REQUIRE(cout % 64 == 0, "initial_conv_wgrad: channels must be a multiple of 64");
  const int M = batch * H * W;
  if (M == 0) return 0;
  LDM_REQUIRE(cin >= 1 && cin <= 4, "initial_conv_wgrad: in_channels %d not in [1,4]", cin);
  if (scratch && dtype == LDM_DT_BF16 && k_conv_wgrad_mn_applicable(64, cout, H, W, 1, dtype) && ((uintptr_t)scratch & 255) == 0 &&
      getenv("LDM_INITIAL_WGRAD_SIMT") == nullptr) {
    bf16* patches = (bf16*)scratch;
    float* acc = (float*)((char*)scratch + ((int64_t)M * 64 * 2 + 255) / 256 * 256);
    const int64_t chunks = (int64_t)M * 8;
    const unsigned g = (unsigned)((chunks + 255) / 256);
    switch (cin) {
      case 1: im2col3x3_small_kernel<1><<<g, 256, 0, st>>>(x, patches, H, W, chunks); break;
      case 2: im2col3x3_small_kernel<2><<<g, 256, 0, st>>>(x, patches, H, W, chunks); break;
      case 3: im2col3x3_small_kernel<3><<<g, 256, 0, st>>>(x, patches, H, W, chunks); break;
      default: im2col3x3_small_kernel<4><<<g, 256, 0, st>>>(x, patches, H, W, chunks); break;
    }
    LDM_LAUNCHED("im2col3x3_small");
    LDM_CUDA(cudaMemsetAsync(acc, 0, (size_t)cout * 64 * sizeof(float), st));
    if (int rc = k_conv_wgrad_mn(patches, 64, 64, dy, cout, cout, acc, dbias, nullptr, batch, H, W, 1, st)) return rc;
    add_cols_kernel<<<(cout * 9 * cin + 255) / 256, 256, 0, st>>>(acc, 64, dw, cout, 9 * cin);
    LDM_LAUNCHED("add_cols");
    return 0;
  }
  int mps = 0;
  const int grid = initial_wgrad_grid(batch, H, W, &mps);
  float* part = (float*)scratch;
#define IW_GO(C) DISPATCH_T(dtype, initial_wgrad_kernel<T, C><<<grid, 256, 0, st>>>(x, (const T*)dy, dw, dbias, part, batch, cout, H, W, mps))
  switch (cin) {
    case 1: IW_GO(1); break;
    case 2: IW_GO(2); break;
    case 3: IW_GO(3); break;
    default: IW_GO(4); break;
  }
#undef IW_GO
  LDM_LAUNCHED("initial_conv_wgrad");
  if (part) {
    const int cols = cout * cin * 9 + cout;
    rows_sum_f32_kernel<<<(cols + 31) / 32, 256, 0, st>>>(part, grid, cols, dw, cout * cin * 9, dbias);
    LDM_LAUNCHED("rows_sum_f32");
  }
  return 0;
}

int k_final_conv_backward(const float* dout, const void* x, int ldx, const float* w, void* dx, float* dw, float* db, int batch,
                          int cin, int cout, int hw, int dtype, cudaStream_t st) {
  LDM_REQUIRE(cout <= 8 && cin % 4 == 0, "final_conv_backward: unsupported channel counts");
  const int64_t total = (int64_t)batch * hw;
  if (total == 0) return 0;
  DISPATCH_T(dtype, final_conv_dx_kernel<T><<<(int)ceil_div64(total, 128), 128, 0, st>>>(dout, w, (T*)dx, cin, cout, hw, total));
  LDM_LAUNCHED("final_conv_dx");
  LDM_REQUIRE(total < (int64_t)1 << 31, "final_conv_backward: batch too large");
  const int mps = split_for(total, 64, 8 * 148);
  int cpad = 1;
  while (cpad < cin && cpad < 256) cpad *= 2;
  DISPATCH_T(dtype, final_conv_dw_kernel<T><<<(int)((total + mps - 1) / mps), 256, 0, st>>>(dout, (const T*)x, ldx, dw, db, cin, cout, hw,
                                                                                       (int)total, mps, cpad));
  LDM_LAUNCHED("final_conv_dw");
  return 0;
}

int k_gemm_f32(const float* a, int64_t a_rs, int64_t a_cs, const float* b, int64_t b_rs, int64_t b_cs, float* c, int64_t c_rs,
               int M, int N, int K, int accumulate, cudaStream_t st) {
  if (M == 0 || N == 0) return 0;
  gemm_f32_kernel<<<dim3((N + 15) / 16, (M + 15) / 16), 256, 0, st>>>(a, a_rs, a_cs, b, b_rs, b_cs, c, c_rs, M, N, K, accumulate);
  LDM_LAUNCHED("gemm_f32");
  return 0;
}
int k_sinusoid(const int64_t* t, float* emb, int batch, int Din, cudaStream_t st) {
  if (batch == 0) return 0;
  sinusoid_kernel<<<(batch * Din + 255) / 256, 256, 0, st>>>(t, emb, batch, Din);
  LDM_LAUNCHED("sinusoid");
  return 0;
}
int k_ew(const float* x, const float* other, float* y, int rows, int cols, int mode, cudaStream_t st) {
  if (rows * cols == 0) return 0;
  ew_kernel<<<(rows * cols + 255) / 256, 256, 0, st>>>(x, other, y, rows, cols, mode);
  LDM_LAUNCHED("ew");
  return 0;
}
int k_rows_scatter_add(const float* src, const int64_t* idx, int idx_len, float* table, int batch, int D, cudaStream_t st) {
  if (batch * D == 0) return 0;
  rows_scatter_add_kernel<<<(batch * D + 255) / 256, 256, 0, st>>>(src, idx, idx_len, table, batch, D);
  LDM_LAUNCHED("rows_scatter_add");
  return 0;
}
int k_rows_gather_add(float* dst, const int64_t* idx, int idx_len, const float* table, int batch, int D, cudaStream_t st) {
  if (batch * D == 0) return 0;
  rows_gather_add_kernel<<<(batch * D + 255) / 256, 256, 0, st>>>(dst, idx, idx_len, table, batch, D);
  LDM_LAUNCHED("rows_gather_add");
  return 0;
}
int k_add(const void* a, const void* b, void* out, int64_t n, int dtype, cudaStream_t st) {
  if (n == 0) return 0;
  DISPATCH_T(dtype, add_kernel<T><<<(int)ceil_div64(n, 256), 256, 0, st>>>((const T*)a, (const T*)b, (T*)out, n));
  LDM_LAUNCHED("add");
  return 0;
}
int k_copy_channels(const void* src, int lds, void* dst, int ldd, int C, int64_t rows, int dtype, cudaStream_t st) {
  if (rows * C == 0) return 0;
  DISPATCH_T(dtype, copy_channels_kernel<T><<<(int)ceil_div64(rows * C, 256), 256, 0, st>>>((const T*)src, lds, (T*)dst, ldd, C, rows));
  LDM_LAUNCHED("copy_channels");
  return 0;
}
