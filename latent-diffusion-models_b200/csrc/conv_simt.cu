// FFMA implicit-GEMM convolution (3x3 pad 1, 1x1, and ConvTranspose2d k2 s2 as a 1x1 GEMM with a
// scatter epilogue) on NHWC tensors.  This is the fp32 parity path (north-star: fp32 <= 1e-4 relative;
// single-pass TF32 tensor-core math measures 2.1e-4) and the A/B baseline for the tcgen05
// kernel in conv_tc.cu.  Replaces F.conv2d / F.conv_transpose2d behind src/UNet.py:54,82,119-120,145-147,231.
//
// GEMM view: C[M = B*H*W, N = Cout] = A_im2col[M, K] * W[N, K]^T, K = taps*Cin (+ Cin2 for the fused
// 1x1 shortcut source).  64x64x16 tiles, 256 threads, 4x4 register micro-tiles, fp32 accumulation.
#include "kernels.h"

#define CS_BM 64
#define CS_BN 64
#define CS_BK 16
#define CS_LD 68

struct ConvSimtParams {
  ConvArgs a;
  int M;          // batch*H*W
  int ktot;       // taps*cin + cin2
  int cout_real;  // output channels (cout/4 for up2)
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) { load_chunk(p, v); }
template <>
__device__ __forceinline__ void load4<bf16>(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) { store_chunk(p, v); }
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, const float (&v)[4]) {
  uint2 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}

template <typename T>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvSimtParams p) {
  __shared__ __align__(16) float As[CS_BK][CS_LD];
  __shared__ __align__(16) float Bs[CS_BK][CS_LD];
  const ConvArgs& a = p.a;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * CS_BM, n0 = blockIdx.y * CS_BN;
  const int H = a.height, W = a.width;
  // loader mapping
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int lm = m0 + lrow;
  const bool lvalid = lm < p.M;
  int ln = 0, lh = 0, lw = 0;
  if (lvalid) { lw = lm % W; int r = lm / W; lh = r % H; ln = r / H; }
  const T* wrow = (const T*)a.w + (int64_t)(n0 + lrow) * p.ktot + lk;
  const bool wvalid = (n0 + lrow) < a.cout;
  // compute mapping
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int taps = a.ksize * a.ksize, pad = a.ksize / 2;
  const int nsrc = a.x2 ? 2 : 1;
  int koff = 0;
  for (int src = 0; src < nsrc; ++src) {
    const T* xs = src == 0 ? (const T*)a.x : (const T*)a.x2;
    const int ld = src == 0 ? a.ldx : a.ldx2;
    const int cin = src == 0 ? a.cin : a.cin2;
    const int ntap = src == 0 ? taps : 1;
    for (int tap = 0; tap < ntap; ++tap) {
      int dy = 0, dx = 0;
      if (src == 0 && taps == 9) { dy = tap / 3 - pad; dx = tap % 3 - pad; }
      const int hh = lh + dy, ww = lw + dx;
      const bool inb = lvalid && hh >= 0 && hh < H && ww >= 0 && ww < W;
      const T* xrow = xs + ((int64_t)(ln * H + hh) * W + ww) * ld + lk;
      for (int c0 = 0; c0 < cin; c0 += CS_BK) {
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        if (inb) load4<T>(xrow + c0, av);
        if (wvalid) load4<T>(wrow + koff + c0, bv);
        __syncthreads();  // previous tile fully consumed
#pragma unroll
        for (int i = 0; i < 4; ++i) { As[lk + i][lrow] = av[i]; Bs[lk + i][lrow] = bv[i]; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CS_BK; ++k) {
          float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
          float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
          float ar[4] = {a4.x, a4.y, a4.z, a4.w}, br[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
        }
      }
      koff += cin;
    }
  }
  // ---- epilogue: + bias + per-sample row vector + residual, store (optionally scattered for up2)
  const int c = n0 + tx * 4;
  if (c >= a.cout) return;
  const int q = a.up2 ? c / p.cout_real : 0;
  const int cc = a.up2 ? c % p.cout_real : c;
  float bsv[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) bsv[j] = a.bias[cc + j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bsv[j];
    const int img = m / (H * W);
    if (a.rowvec) {
      const float* rv = a.rowvec + (int64_t)img * a.ld_rowvec + cc;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += rv[j];
    }
    if (a.res) {
      float r[4];
      load4<T>((const T*)a.res + (int64_t)m * a.ldres + cc, r);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += r[j];
    }
    int64_t orow = m;
    if (a.up2) {
      const int w_ = m % W, r_ = m / W, h_ = r_ % H;
      orow = ((int64_t)img * 2 * H + 2 * h_ + (q >> 1)) * (2 * W) + 2 * w_ + (q & 1);
    }
    store4<T>((T*)a.y + orow * a.ldy + cc, v);
  }
}

static int conv_check(const ConvArgs& a) {
  LDM_REQUIRE(a.ksize == 1 || a.ksize == 3, "conv2d: kernel size %d unsupported (1 or 3)", a.ksize);
  LDM_REQUIRE(a.res_mod == 0, "conv2d: residual row aliasing is only implemented in the halo kernel");
  LDM_REQUIRE(a.cin % 16 == 0 && (a.x2 == nullptr || a.cin2 % 16 == 0), "conv2d: Cin must be a multiple of 16");
  LDM_REQUIRE(a.cout % 4 == 0, "conv2d: Cout must be a multiple of 4");
  LDM_REQUIRE(!a.up2 || (a.ksize == 1 && a.x2 == nullptr && a.res == nullptr && a.rowvec == nullptr),
              "conv2d: up2 epilogue only for plain 1x1 GEMMs");
  LDM_REQUIRE(a.ldx % 4 == 0 && a.ldy % 4 == 0, "conv2d: strides must be multiples of 4 elements");
  return 0;
}

int k_conv_simt(const ConvArgs& a, cudaStream_t st) {
  if (int rc = conv_check(a)) return rc;
  ConvSimtParams p;
  p.a = a;
  p.M = a.batch * a.height * a.width;
  p.ktot = a.ksize * a.ksize * a.cin + (a.x2 ? a.cin2 : 0);
  p.cout_real = a.up2 ? a.cout / 4 : a.cout;
  if (p.M == 0) return 0;
  dim3 grid((p.M + CS_BM - 1) / CS_BM, (a.cout + CS_BN - 1) / CS_BN);
  if (a.dtype == LDM_DT_BF16) conv_simt_kernel<bf16><<<grid, 256, 0, st>>>(p);
  else conv_simt_kernel<float><<<grid, 256, 0, st>>>(p);
  LDM_LAUNCHED("conv_simt");
  return 0;
}

int k_conv(const ConvArgs& a, int impl, cudaStream_t st) {
  if (a.dtype == LDM_DT_BF16 && impl == 0) {
    if (k_conv_halo_applicable(a)) return k_conv_halo(a, st);
    LDM_REQUIRE(!a.xf_ab && a.x_mod == 0, "conv: the on-the-fly source GroupNorm / x_mod exist only in the halo kernel");
    return k_conv_tc(a, st);
  }
  LDM_REQUIRE(!a.xf_ab && a.x_mod == 0, "conv: the on-the-fly source GroupNorm / x_mod exist only in the halo kernel");
  LDM_REQUIRE(a.gn.mode == 0, "conv: the fused GroupNorm epilogue exists only in the tcgen05 kernels (bf16, impl 0)");
  return k_conv_simt(a, st);
}
