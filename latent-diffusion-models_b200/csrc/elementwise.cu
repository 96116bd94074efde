// Memory-bound kernels of the DDPM hot path: layout converters, max-pool, the fused
// p_sample / q_sample updates, the time-embedding MLP and the tiny first/last convolutions.
// All are HBM-bound: 16-byte vector accesses, one read + one write of every live tensor.
#include "kernels.h"

// ------------------------------------------------------------------ layout converters (tests, taps)
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int C, int HW,
                                    int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  int64_t r = i / C;
  int p = (int)(r % HW);
  int64_t b = r / HW;
  y[i] = from_float<T>(x[(b * C + c) * HW + p]);
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, int ldx, float* __restrict__ y, int C, int HW,
                                    int64_t total) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int p = (int)(i % HW);
  int64_t r = i / HW;
  int c = (int)(r % C);
  int64_t b = r / C;
  y[i] = to_float(x[(b * HW + p) * (int64_t)ldx + c]);
}
int k_nchw_to_nhwc(const float* x, void* y, int batch, int channels, int hw, int dtype, cudaStream_t st) {
  int64_t total = (int64_t)batch * channels * hw;
  if (total == 0) return 0;
  int grid = (int)ceil_div64(total, 256);
  if (dtype == LDM_DT_BF16) nchw_to_nhwc_kernel<bf16><<<grid, 256, 0, st>>>(x, (bf16*)y, channels, hw, total);
  else nchw_to_nhwc_kernel<float><<<grid, 256, 0, st>>>(x, (float*)y, channels, hw, total);
  LDM_LAUNCHED("nchw_to_nhwc");
  return 0;
}
int k_nhwc_to_nchw(const void* x, int ldx, float* y, int batch, int channels, int hw, int dtype,
                   cudaStream_t st) {
  int64_t total = (int64_t)batch * channels * hw;
  if (total == 0) return 0;
  int grid = (int)ceil_div64(total, 256);
  if (dtype == LDM_DT_BF16) nhwc_to_nchw_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, ldx, y, channels, hw, total);
  else nhwc_to_nchw_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ldx, y, channels, hw, total);
  LDM_LAUNCHED("nhwc_to_nchw");
  return 0;
}

// ------------------------------------------------------------------ MaxPool2d(2,2)  src/UNet.py:183,207
template <typename T>
__global__ void maxpool2_kernel(const T* __restrict__ x, int ldx, T* __restrict__ y, int ldy, int H, int W,
                                int C, int64_t total_chunks) {
  pdl_wait();
  pdl_trigger();
  constexpr int V = VecTraits<T>::N;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total_chunks) return;
  int cpp = C / V;
  int ci = (int)(i % cpp);
  int64_t r = i / cpp;
  int Wo = W / 2, Ho = H / 2;
  int wo = (int)(r % Wo); r /= Wo;
  int ho = (int)(r % Ho);
  int64_t b = r / Ho;
  const T* p00 = x + ((b * H + 2 * ho) * W + 2 * wo) * (int64_t)ldx + ci * V;
  float a[V], c[V], d[V], e[V], o[V];
  load_chunk(p00, a);
  load_chunk(p00 + ldx, c);
  load_chunk(p00 + (int64_t)W * ldx, d);
  load_chunk(p00 + (int64_t)W * ldx + ldx, e);
#pragma unroll
  for (int k = 0; k < V; ++k) o[k] = fmaxf(fmaxf(a[k], c[k]), fmaxf(d[k], e[k]));
  store_chunk(y + ((b * Ho + ho) * Wo + wo) * (int64_t)ldy + ci * V, o);
}
int k_maxpool2(const void* x, int ldx, void* y, int ldy, int batch, int height, int width, int channels,
               int dtype, cudaStream_t st) {
  int V = dtype == LDM_DT_BF16 ? 8 : 4;
  LDM_REQUIRE(channels % V == 0 && ldx % V == 0 && ldy % V == 0, "maxpool2: channels/ld must be a multiple of %d", V);
  LDM_REQUIRE(height % 2 == 0 && width % 2 == 0, "maxpool2: odd spatial size %dx%d", height, width);
  int64_t total = (int64_t)batch * (height / 2) * (width / 2) * (channels / V);
  if (total == 0) return 0;
  int grid = (int)ceil_div64(total, 256);
  if (dtype == LDM_DT_BF16)
    LDM_CUDA(ldm_launch_pdl(maxpool2_kernel<bf16>, dim3(grid), dim3(256), 0, st, (const bf16*)x, ldx, (bf16*)y, ldy, height, width, channels, total));
  else
    maxpool2_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ldx, (float*)y, ldy, height, width, channels, total);
  LDM_LAUNCHED("maxpool2");
  return 0;
}

// ------------------------------------------------------------------ schedule coefficient table
// coef[t] = {alpha^-1/2, (1-alpha)/sqrt(1-abar), sqrt(beta), 0}: the per-step scalars of src/DDPM.py:74-96
__global__ void build_coef_kernel(const float* beta, const float* alpha, const float* abar, int T, float* coef) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float a = alpha[t], ab = abar[t], b = beta[t];
  float4 c;
  c.x = 1.0f / sqrtf(a);
  c.y = (1.0f - a) / sqrtf(1.0f - ab);
  c.z = sqrtf(b);
  c.w = 0.0f;
  reinterpret_cast<float4*>(coef)[t] = c;
}
int k_build_coef(const float* beta, const float* alpha, const float* alpha_bar, int n_steps, float* coef,
                 cudaStream_t st) {
  LDM_REQUIRE(n_steps > 0, "build_coef: n_steps must be positive");
  build_coef_kernel<<<(n_steps + 127) / 128, 128, 0, st>>>(beta, alpha, alpha_bar, n_steps, coef);
  LDM_LAUNCHED("build_coef");
  return 0;
}

// ------------------------------------------------------------------ randn (x_T ~ N(0,I), src/DDPM.py:108)
__global__ void randn_kernel(float* __restrict__ out, int64_t n4, int64_t total4, uint64_t seed,
                             uint64_t sample_offset, uint32_t stream_id, uint32_t step) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  int64_t b = i / n4, e4 = i % n4;
  uint64_t s = sample_offset + (uint64_t)b;
  float z[4];
  philox_normal4(seed, (uint32_t)e4, (uint32_t)s, step, stream_id ^ ((uint32_t)(s >> 32) << 8), z);
  reinterpret_cast<float4*>(out)[i] = make_float4(z[0], z[1], z[2], z[3]);
}
int k_randn(float* out, int batch, int64_t n_per_sample, uint64_t seed, uint64_t sample_offset,
            uint64_t stream_id, cudaStream_t st) {
  LDM_REQUIRE(n_per_sample % 4 == 0, "randn: n_per_sample must be a multiple of 4");
  int64_t n4 = n_per_sample / 4, total4 = n4 * batch;
  if (total4 == 0) return 0;
  randn_kernel<<<(int)ceil_div64(total4, 256), 256, 0, st>>>(out, n4, total4, seed, sample_offset,
                                                            (uint32_t)stream_id, 0xFFFFFFFFu);
  LDM_LAUNCHED("randn");
  return 0;
}

// ------------------------------------------------------------------ q_sample  src/DDPM.py:46-68,133-149
__global__ void q_sample_kernel(const float4* __restrict__ x0, const int64_t* __restrict__ t,
                                const float* __restrict__ abar, int T, const float4* __restrict__ eps,
                                float4* __restrict__ eps_out, float4* __restrict__ xt, int64_t n4,
                                int64_t total4, uint64_t seed, uint64_t sample_offset) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  int64_t b = i / n4, e4 = i % n4;
  int64_t tb = t[b];
  tb = tb < 0 ? 0 : (tb >= T ? T - 1 : tb);
  float ab = abar[tb];
  float sa = sqrtf(ab), sb = sqrtf(1.0f - ab);
  float4 x = x0[i], e;
  if (eps) {
    e = eps[i];
  } else {
    float z[4];
    uint64_t s = sample_offset + (uint64_t)b;
    philox_normal4(seed, (uint32_t)e4, (uint32_t)s, (uint32_t)tb, 0x71u ^ ((uint32_t)(s >> 32) << 8), z);
    e = make_float4(z[0], z[1], z[2], z[3]);
  }
  if (eps_out) eps_out[i] = e;
  xt[i] = make_float4(sa * x.x + sb * e.x, sa * x.y + sb * e.y, sa * x.z + sb * e.z, sa * x.w + sb * e.w);
}
int k_q_sample(const float* x0, const int64_t* t, const float* alpha_bar, int n_steps, const float* eps,
               float* eps_out, float* xt, int batch, int64_t n_per_sample, uint64_t seed,
               uint64_t sample_offset, cudaStream_t st) {
  LDM_REQUIRE(n_per_sample % 4 == 0, "q_sample: n_per_sample must be a multiple of 4");
  LDM_REQUIRE(eps != nullptr || eps_out != nullptr, "q_sample: eps_out is required when eps is drawn in-kernel");
  int64_t n4 = n_per_sample / 4, total4 = n4 * batch;
  if (total4 == 0) return 0;
  q_sample_kernel<<<(int)ceil_div64(total4, 256), 256, 0, st>>>(
      (const float4*)x0, t, alpha_bar, n_steps, (const float4*)eps, (float4*)eps_out, (float4*)xt, n4, total4,
      seed, sample_offset);
  LDM_LAUNCHED("q_sample");
  return 0;
}

// ------------------------------------------------------------------ p_sample + CFG  src/DDPM.py:71-96,120-124
__device__ __forceinline__ float lerp_torch(float start, float end, float w) {
  // at::lerp: w < 0.5 ? start + w (end-start) : end - (end-start)(1-w)
  float d = end - start;
  return w < 0.5f ? start + w * d : end - d * (1.0f - w);
}
__global__ void p_sample_kernel(const float4* __restrict__ xt, const float4* __restrict__ ec,
                                const float4* __restrict__ eu, float cfg, const int64_t* __restrict__ t_dev,
                                int t_stride, const float4* __restrict__ coef, int T, const float* __restrict__ noise,
                                int64_t noise_t_stride, uint64_t seed, uint64_t sample_offset,
                                const uint64_t* __restrict__ seed_dev, float4* __restrict__ out, int64_t n4, int64_t total4) {
  pdl_wait();
  pdl_trigger();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  // device-resident {seed, sample_offset}: a captured sampler step is replayed with fresh seeds without re-capturing
  if (seed_dev) { seed = seed_dev[0]; sample_offset = seed_dev[1]; }
  // the reference gathers alpha/alpha_bar per sample but takes the noise branch on t[0] (src/DDPM.py:74-85)
  int64_t t = t_dev[0];
  t = t < 0 ? 0 : (t >= T ? T - 1 : t);
  int64_t tb = t_stride ? t_dev[(i / n4) * t_stride] : t;
  tb = tb < 0 ? 0 : (tb >= T ? T - 1 : tb);
  float4 c = coef[tb];
  float4 x = xt[i], e = ec[i];
  if (eu) {
    float4 u = eu[i];
    e.x = lerp_torch(u.x, e.x, cfg); e.y = lerp_torch(u.y, e.y, cfg);
    e.z = lerp_torch(u.z, e.z, cfg); e.w = lerp_torch(u.w, e.w, cfg);
  }
  float4 m;
  m.x = c.x * (x.x - c.y * e.x); m.y = c.x * (x.y - c.y * e.y);
  m.z = c.x * (x.z - c.y * e.z); m.w = c.x * (x.w - c.y * e.w);
  if (t > 0) {
    float z[4];
    if (noise) {
      float4 zz = reinterpret_cast<const float4*>(noise + t * noise_t_stride)[i];
      z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
    } else {
      int64_t b = i / n4, e4 = i % n4;
      uint64_t s = sample_offset + (uint64_t)b;
      philox_normal4(seed, (uint32_t)e4, (uint32_t)s, (uint32_t)t, 0x9Eu ^ ((uint32_t)(s >> 32) << 8), z);
    }
    m.x += c.z * z[0]; m.y += c.z * z[1]; m.z += c.z * z[2]; m.w += c.z * z[3];
  }
  out[i] = m;
}
int k_p_sample(const float* xt, const float* eps_c, const float* eps_u, float cfg_scale, const int64_t* t_dev,
               int t_stride, const float* coef, int n_steps, const float* noise, int64_t noise_t_stride, uint64_t seed,
               uint64_t sample_offset, float* out, int batch, int64_t n_per_sample, cudaStream_t st, const uint64_t* seed_dev) {
  LDM_REQUIRE(n_per_sample % 4 == 0, "p_sample: n_per_sample must be a multiple of 4");
  int64_t n4 = n_per_sample / 4, total4 = n4 * batch;
  if (total4 == 0) return 0;
  LDM_CUDA(ldm_launch_pdl(p_sample_kernel, dim3((unsigned)ceil_div64(total4, 256)), dim3(256), 0, st, (const float4*)xt,
                          (const float4*)eps_c, (const float4*)eps_u, cfg_scale, t_dev, t_stride, (const float4*)coef, n_steps, noise,
                          noise_t_stride, seed, sample_offset, seed_dev, (float4*)out, n4, total4));
  LDM_LAUNCHED("p_sample");
  return 0;
}

__global__ void set_i64_kernel(int64_t* p, int64_t v) { *p = v; }
__global__ void add_i64_kernel(int64_t* p, int64_t v) {
  pdl_wait();
  pdl_trigger();
  *p += v;
}
int k_set_i64(int64_t* p, int64_t v, cudaStream_t st) {
  set_i64_kernel<<<1, 1, 0, st>>>(p, v);
  LDM_LAUNCHED("set_i64");
  return 0;
}
int k_add_i64(int64_t* p, int64_t v, cudaStream_t st) {
  LDM_CUDA(ldm_launch_pdl(add_i64_kernel, dim3(1), dim3(1), 0, st, p, v));
  LDM_LAUNCHED("add_i64");
  return 0;
}

// ------------------------------------------------------------------ time embedding  src/UNet.py:23-44,263-268,373-376
// One CTA handles TE_ROWS batch rows so each weight is read once per TE_ROWS rows. blockDim = D.
#define TE_ROWS 8
__global__ void time_embed_kernel(const int64_t* __restrict__ t, const int64_t* __restrict__ t_scalar,
                                  const int64_t* __restrict__ y, int y_len, int y_rows,
                                  const float* __restrict__ w1t, const float* __restrict__ b1,
                                  const float* __restrict__ w3t, const float* __restrict__ b3,
                                  const float* __restrict__ label_emb, float* __restrict__ temb, int batch,
                                  int D, int table_classes) {
  extern __shared__ float sm[];
  const int Din = D / 4, half = D / 8;
  float* emb = sm;                 // [TE_ROWS][Din]
  float* h1 = sm + TE_ROWS * Din;  // [TE_ROWS][D]
  const int j = threadIdx.x;
  const int b0 = blockIdx.x * TE_ROWS;
  // sinusoid: f_i = exp(i * -(ln 1e4 / (half-1))) in fp32, arg = float(t) * f_i
  const float neg = -(float)(9.210340371976184 / (double)(half - 1));
  for (int idx = j; idx < TE_ROWS * Din; idx += blockDim.x) {
    int r = idx / Din, i = idx % Din;
    int b = b0 + r;
    float v = 0.f;
    if (b < batch) {
      float tv = (float)(t ? t[b] : *t_scalar);
      int fi = i < half ? i : i - half;
      float arg = tv * expf((float)fi * neg);
      v = i < half ? sinf(arg) : cosf(arg);
    }
    emb[idx] = v;
  }
  __syncthreads();
  float acc[TE_ROWS];
#pragma unroll
  for (int r = 0; r < TE_ROWS; ++r) acc[r] = 0.f;
#pragma unroll 8
  for (int i = 0; i < Din; ++i) {
    float w = w1t[i * D + j];
#pragma unroll
    for (int r = 0; r < TE_ROWS; ++r) acc[r] = fmaf(emb[r * Din + i], w, acc[r]);
  }
  float bb = b1[j];
#pragma unroll
  for (int r = 0; r < TE_ROWS; ++r) {
    float v = acc[r] + bb;
    h1[r * D + j] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));  // exact-erf GELU
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < TE_ROWS; ++r) acc[r] = 0.f;
#pragma unroll 8
  for (int k = 0; k < D; ++k) {
    float w = w3t[k * D + j];
#pragma unroll
    for (int r = 0; r < TE_ROWS; ++r) acc[r] = fmaf(h1[r * D + k], w, acc[r]);
  }
  bb = b3[j];
#pragma unroll
  for (int r = 0; r < TE_ROWS; ++r) {
    int b = b0 + r;
    if (b >= batch) break;
    float v = acc[r] + bb;
    if (table_classes != 0) {  // table mode: row b = class b; rows >= table_classes (or all, if < 0) carry no label
      if (b < table_classes) v += label_emb[(int64_t)b * D + j];
    } else if (y && y_len > 0 && b < y_rows) {
      int64_t cls = y_len == 1 ? y[0] : y[b];
      v += label_emb[cls * D + j];
    }
    temb[(int64_t)b * D + j] = v;
  }
}
int k_time_embed(const int64_t* t, const int64_t* t_scalar, const int64_t* y, int y_len, int y_rows,
                 const float* w1t, const float* b1, const float* w3t, const float* b3, const float* label_emb,
                 float* temb, int batch, int D, int table_classes, cudaStream_t st) {
  LDM_REQUIRE(D % 32 == 0 && D <= 1024 && D >= 32, "time_embed: unsupported embedding width %d", D);
  LDM_REQUIRE(t != nullptr || t_scalar != nullptr, "time_embed: need t or t_dev_scalar");
  if (batch == 0) return 0;
  size_t smem = (size_t)TE_ROWS * (D / 4 + D) * sizeof(float);
  time_embed_kernel<<<(batch + TE_ROWS - 1) / TE_ROWS, D, smem, st>>>(t, t_scalar, y, y_len, y_rows, w1t, b1,
                                                                      w3t, b3, label_emb, temb, batch, D,
                                                                      table_classes);
  LDM_LAUNCHED("time_embed");
  return 0;
}

// tproj = Linear(SiLU(temb)) for all ResNetBlocks at once  src/UNet.py:70-73,90-93
#define TP_ROWS 16
__global__ void time_proj_kernel(const float* __restrict__ temb, const float* __restrict__ wt,
                                 const float* __restrict__ bias, float* __restrict__ tproj, int batch, int D,
                                 int total) {
  extern __shared__ float s[];  // [TP_ROWS][D] silu(temb)
  const int b0 = blockIdx.y * TP_ROWS;
  for (int idx = threadIdx.x; idx < TP_ROWS * D; idx += blockDim.x) {
    int r = idx / D, k = idx % D;
    float v = 0.f;
    if (b0 + r < batch) v = silu_acc(temb[(int64_t)(b0 + r) * D + k]);
    s[idx] = v;
  }
  __syncthreads();
  int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= total) return;
  float acc[TP_ROWS];
#pragma unroll
  for (int r = 0; r < TP_ROWS; ++r) acc[r] = 0.f;
#pragma unroll 8
  for (int k = 0; k < D; ++k) {
    float w = wt[(int64_t)k * total + o];
#pragma unroll
    for (int r = 0; r < TP_ROWS; ++r) acc[r] = fmaf(s[r * D + k], w, acc[r]);
  }
  float bb = bias[o];
#pragma unroll
  for (int r = 0; r < TP_ROWS; ++r)
    if (b0 + r < batch) tproj[(int64_t)(b0 + r) * total + o] = acc[r] + bb;
}
int k_time_proj(const float* temb, const float* wt, const float* bias, float* tproj, int batch, int D,
                int total, cudaStream_t st) {
  if (batch == 0 || total == 0) return 0;
  dim3 grid((total + 127) / 128, (batch + TP_ROWS - 1) / TP_ROWS);
  size_t smem = (size_t)TP_ROWS * D * sizeof(float);
  LDM_REQUIRE(smem <= 48 * 1024, "time_proj: embedding width %d too large", D);
  time_proj_kernel<<<grid, 128, smem, st>>>(temb, wt, bias, tproj, batch, D, total);
  LDM_LAUNCHED("time_proj");
  return 0;
}

// ------------------------------------------------------------------ initial 3x3 conv (Cin <= 8)  src/UNet.py:331,378
// fp32 NCHW in -> NHWC out.  One thread: one pixel x 8 output channels.  w smem [9][Cin][Cout].
// One thread per vertical pixel PAIR of one distinct input image: the 4x3xCin input patch sits in registers, weights
// are warp-uniform shared-memory broadcasts (each feeds two pixels), input reads are coalesced along W.  Rows
// b, b + x_batch, ... of the output alias the same image (the sampler's cond/uncond halves): computed once, stored
// `reps` times.
template <typename T, int CIN>
__global__ void __launch_bounds__(128)
initial_conv_kernel(const float* __restrict__ x, int x_batch, int reps, const float* __restrict__ w,
                    const float* __restrict__ bias, T* __restrict__ y, int Cout, int H, int W, int64_t total) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(16) float sw[];  // [9*CIN][Cout] weights, then Cout bias
  for (int i = threadIdx.x; i < 9 * CIN * Cout; i += blockDim.x) sw[i] = w[i];
  float* sb = sw + 9 * CIN * Cout;
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sb[i] = bias[i];
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int H2 = H >> 1;
  const int wq = (int)(i % W);
  const int h0 = (int)((i / W) % H2) * 2;
  const int64_t bs = i / ((int64_t)W * H2);
  float xv[12 * CIN];  // rows h0-1 .. h0+2
#pragma unroll
  for (int dy = 0; dy < 4; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int hh = h0 + dy - 1, ww = wq + dx - 1;
      const bool ok = hh >= 0 && hh < H && ww >= 0 && ww < W;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
        xv[(dy * 3 + dx) * CIN + ci] = ok ? __ldg(x + ((bs * CIN + ci) * H + hh) * (int64_t)W + ww) : 0.f;
    }
  for (int c0 = 0; c0 < Cout; c0 += 8) {
    float acc[2][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[0][k] = acc[1][k] = sb[c0 + k];
#pragma unroll
    for (int j = 0; j < 9 * CIN; ++j) {
      const float4* wp = reinterpret_cast<const float4*>(sw + j * Cout + c0);
      const float4 wa = wp[0], wb = wp[1];
      const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[0][k] = fmaf(xv[j], wv[k], acc[0][k]);
        acc[1][k] = fmaf(xv[j + 3 * CIN], wv[k], acc[1][k]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      for (int rep = 0; rep < reps; ++rep) {
        T* yp = y + ((((int64_t)rep * x_batch + bs) * H + h0 + r) * W + wq) * (int64_t)Cout + c0;
        if constexpr (sizeof(T) == 2) {
          store_chunk(yp, acc[r]);
        } else {
          float lo[4] = {acc[r][0], acc[r][1], acc[r][2], acc[r][3]}, hi[4] = {acc[r][4], acc[r][5], acc[r][6], acc[r][7]};
          store_chunk(yp, lo);
          store_chunk(yp + 4, hi);
        }
      }
    }
  }
}
template <typename T>
static int initial_conv_launch(const float* x, int x_batch, const float* w, const float* bias, T* y, int batch, int cin,
                               int cout, int height, int width, size_t smem, cudaStream_t st) {
  const int64_t total = (int64_t)x_batch * (height / 2) * width;
  const int reps = batch / x_batch;
  const int grid = (int)ceil_div64(total, 128);
#define IC_GO(C) LDM_CUDA(ldm_launch_pdl(initial_conv_kernel<T, C>, dim3(grid), dim3(128), smem, st, x, x_batch, reps, w, bias, y, cout, height, width, total))
  switch (cin) {
    case 1: IC_GO(1); break;
    case 2: IC_GO(2); break;
    case 3: IC_GO(3); break;
    case 4: IC_GO(4); break;
    case 5: IC_GO(5); break;
    case 6: IC_GO(6); break;
    case 7: IC_GO(7); break;
    default: IC_GO(8); break;
  }
#undef IC_GO
  LDM_LAUNCHED("initial_conv");
  return 0;
}
int k_initial_conv(const float* x, int x_batch, const float* w, const float* bias, void* y, int batch, int cin,
                   int cout, int height, int width, int dtype, cudaStream_t st) {
  LDM_REQUIRE(cin >= 1 && cin <= 8, "initial_conv: in_channels %d not in [1,8]", cin);
  LDM_REQUIRE(cout % 8 == 0, "initial_conv: channels must be a multiple of 8");
  LDM_REQUIRE(height % 2 == 0, "initial_conv: height must be even");
  LDM_REQUIRE(x_batch > 0 && batch % x_batch == 0, "initial_conv: x_batch must divide batch");
  if ((int64_t)batch * height * width == 0) return 0;
  size_t smem = (size_t)(9 * cin * cout + cout) * sizeof(float);
  LDM_REQUIRE(smem <= 48 * 1024, "initial_conv: weights do not fit shared memory");
  if (dtype == LDM_DT_BF16)
    return initial_conv_launch<bf16>(x, x_batch, w, bias, (bf16*)y, batch, cin, cout, height, width, smem, st);
  return initial_conv_launch<float>(x, x_batch, w, bias, (float*)y, batch, cin, cout, height, width, smem, st);
}

// ---- batch-constant timestep (sampler): the time MLP runs once, the label rows and the 8 mlp_t projections on a
// (num_classes + 1)-row table.  Warp-per-output dot products over the ORIGINAL [out][in] weight rows: every weight
// is requested by exactly one coalesced load issued up front, so the kernels cost one memory latency, not D of them.
__global__ void __launch_bounds__(1024)
time_table_kernel(const int64_t* __restrict__ t_scalar, const float* __restrict__ w1, const float* __restrict__ b1,
                  const float* __restrict__ w3, const float* __restrict__ b3, const float* __restrict__ label_emb,
                  float* __restrict__ s_tab, int R, int n_classes, int D) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float sm[];
  const int Din = D / 4, half = D / 8;
  float* emb = sm;             // [Din]
  float* h1 = sm + Din;        // [D]
  float* tt = h1 + D;          // [D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float neg = -(float)(9.210340371976184 / (double)(half - 1));
  const float tv = (float)(*t_scalar);
  for (int i = threadIdx.x; i < Din; i += blockDim.x) {
    const int fi = i < half ? i : i - half;
    const float arg = tv * expf((float)fi * neg);
    emb[i] = i < half ? sinf(arg) : cosf(arg);
  }
  __syncthreads();
  // Each warp owns D/nwarps outputs; ALL of its weight loads are issued before the first reduction, so the layer
  // costs one memory round trip instead of one per output.
  constexpr int TT_MAX = 8;   // outputs per warp (D <= 8 * nwarps)
  {
    float a[TT_MAX];
#pragma unroll
    for (int u = 0; u < TT_MAX; ++u) {
      const int j = warp + u * nwarps;
      a[u] = 0.f;
      if (j < D)
        for (int i = lane; i < Din; i += 32) a[u] = fmaf(emb[i], __ldg(w1 + (int64_t)j * Din + i), a[u]);
    }
#pragma unroll
    for (int u = 0; u < TT_MAX; ++u) {
      const int j = warp + u * nwarps;
      const float s = warp_sum(a[u]);
      if (lane == 0 && j < D) {
        const float v = s + b1[j];
        h1[j] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
      }
    }
  }
  __syncthreads();
  {
    float4 wv[TT_MAX][2];
#pragma unroll
    for (int u = 0; u < TT_MAX; ++u) {
      const int j = warp + u * nwarps;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int k = lane * 4 + q * 128;
        wv[u][q] = (j < D && k < D) ? __ldg(reinterpret_cast<const float4*>(w3 + (int64_t)j * D + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < TT_MAX; ++u) {
      const int j = warp + u * nwarps;
      float a = 0.f;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int k = lane * 4 + q * 128;
        if (k < D) {
          a = fmaf(h1[k], wv[u][q].x, a); a = fmaf(h1[k + 1], wv[u][q].y, a);
          a = fmaf(h1[k + 2], wv[u][q].z, a); a = fmaf(h1[k + 3], wv[u][q].w, a);
        }
      }
      for (int k = lane * 4 + 256; k < D; k += 128) {   // D > 256: remaining columns the slow way
        if (j < D) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(w3 + (int64_t)j * D + k));
          a = fmaf(h1[k], w4.x, a); a = fmaf(h1[k + 1], w4.y, a); a = fmaf(h1[k + 2], w4.z, a); a = fmaf(h1[k + 3], w4.w, a);
        }
      }
      a = warp_sum(a);
      if (lane == 0 && j < D) tt[j] = a + b3[j];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < R * D; idx += blockDim.x) {
    const int r = idx / D, j = idx % D;
    float v = tt[j];
    if (r < n_classes) v += label_emb[(int64_t)r * D + j];
    s_tab[idx] = silu_acc(v);   // SiLU of mlp_t (src/UNet.py:72) applied here, once per table entry
  }
}
__global__ void __launch_bounds__(256)
time_proj_table_kernel(const float* __restrict__ s_tab, const float* __restrict__ w, const float* __restrict__ bias,
                       float* __restrict__ tproj_tab, int R, int D, int total) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float s[];  // [R][D]
  for (int i = threadIdx.x; i < R * D; i += blockDim.x) s[i] = s_tab[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + warp;
  if (o >= total) return;
  float4 wv[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int k = lane * 4 + q * 128;
    wv[q] = k < D ? __ldg(reinterpret_cast<const float4*>(w + (int64_t)o * D + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float bo = bias[o];
  for (int r = 0; r < R; ++r) {
    float a = 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int k = lane * 4 + q * 128;
      if (k < D) {
        const float4 s4 = *reinterpret_cast<const float4*>(s + r * D + k);
        a = fmaf(s4.x, wv[q].x, a); a = fmaf(s4.y, wv[q].y, a); a = fmaf(s4.z, wv[q].z, a); a = fmaf(s4.w, wv[q].w, a);
      }
    }
    a = warp_sum(a);
    if (lane == 0) tproj_tab[(int64_t)r * total + o] = a + bo;
  }
}
int k_time_table(const int64_t* t_scalar, const float* w1, const float* b1, const float* w3, const float* b3,
                 const float* label_emb, const float* tproj_w, const float* tproj_b, float* s_tab, float* tproj_tab,
                 int R, int n_classes, int D, int total, cudaStream_t st) {
  LDM_REQUIRE(D % 128 == 0 && D <= 256, "time_table: unsupported embedding width %d", D);
  LDM_REQUIRE((size_t)R * D * 4 <= 48 * 1024, "time_table: %d table rows do not fit shared memory", R);
  LDM_CUDA(ldm_launch_pdl(time_table_kernel, dim3(1), dim3(1024), (size_t)(D / 4 + 2 * D) * sizeof(float), st, t_scalar, w1, b1, w3, b3,
                          label_emb, s_tab, R, n_classes, D));
  LDM_LAUNCHED("time_table");
  LDM_CUDA(ldm_launch_pdl(time_proj_table_kernel, dim3((total + 7) / 8), dim3(256), (size_t)R * D * sizeof(float), st, (const float*)s_tab,
                          tproj_w, tproj_b, tproj_tab, R, D, total));
  LDM_LAUNCHED("time_proj_table");
  return 0;
}

// tproj[b][:] = tab[class(b)][:]: expands the per-class table of a batch-constant timestep (sampler) to batch rows.
__global__ void tproj_gather_kernel(const float4* __restrict__ tab, const int64_t* __restrict__ y, int y_len,
                                    int y_rows, int n_classes, float4* __restrict__ tproj, int batch, int total4) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y;
  int row = n_classes;  // unlabeled
  if (y && y_len > 0 && b < y_rows) row = (int)(y_len == 1 ? y[0] : y[b]);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x)
    tproj[(int64_t)b * total4 + i] = tab[(int64_t)row * total4 + i];
}
int k_tproj_gather(const float* tab, const int64_t* y, int y_len, int y_rows, int n_classes, float* tproj, int batch,
                   int total, cudaStream_t st) {
  LDM_REQUIRE(total % 4 == 0, "tproj_gather: width %d not a multiple of 4", total);
  if (batch == 0 || total == 0) return 0;
  LDM_CUDA(ldm_launch_pdl(tproj_gather_kernel, dim3((total / 4 + 127) / 128, batch), dim3(128), 0, st, (const float4*)tab, y, y_len,
                          y_rows, n_classes, (float4*)tproj, batch, total / 4));
  LDM_LAUNCHED("tproj_gather");
  return 0;
}

// ------------------------------------------------------------------ final 1x1 conv (Cout <= 8)  src/UNet.py:347
template <typename T>
__global__ void final_conv_kernel(const T* __restrict__ x, int ldx, const float* __restrict__ w,
                                  const float* __restrict__ bias, float* __restrict__ y, int Cin, int Cout,
                                  int HW, int64_t total) {
  constexpr int V = VecTraits<T>::N;
  extern __shared__ float sw[];  // [Cout][Cin] then bias
  for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) sw[i] = w[i];
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) sw[Cout * Cin + i] = bias[i];
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t b = i / HW;
  int p = (int)(i % HW);
  const T* xp = x + i * (int64_t)ldx;
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = o < Cout ? sw[Cout * Cin + o] : 0.f;
  for (int c0 = 0; c0 < Cin; c0 += V) {
    float v[V];
    load_chunk(xp + c0, v);
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      if (o < Cout) {
#pragma unroll
        for (int k = 0; k < V; ++k) acc[o] = fmaf(v[k], sw[o * Cin + c0 + k], acc[o]);
      }
    }
  }
  for (int o = 0; o < Cout; ++o) y[(b * Cout + o) * (int64_t)HW + p] = acc[o];
}
int k_final_conv(const void* x, int ldx, const float* w, const float* bias, float* y, int batch, int cin,
                 int cout, int hw, int dtype, cudaStream_t st) {
  LDM_REQUIRE(cout >= 1 && cout <= 8, "final_conv: out_channels %d not in [1,8]", cout);
  int V = dtype == LDM_DT_BF16 ? 8 : 4;
  LDM_REQUIRE(cin % V == 0 && ldx % V == 0, "final_conv: channels must be a multiple of %d", V);
  int64_t total = (int64_t)batch * hw;
  if (total == 0) return 0;
  size_t smem = (size_t)(cout * cin + cout) * sizeof(float);
  int grid = (int)ceil_div64(total, 128);
  if (dtype == LDM_DT_BF16)
    final_conv_kernel<bf16><<<grid, 128, smem, st>>>((const bf16*)x, ldx, w, bias, y, cin, cout, hw, total);
  else
    final_conv_kernel<float><<<grid, 128, smem, st>>>((const float*)x, ldx, w, bias, y, cin, cout, hw, total);
  LDM_LAUNCHED("final_conv");
  return 0;
}
