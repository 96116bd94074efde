// Weight repacking: PyTorch state_dict layouts (fp32) -> the K-major GEMM operands the conv kernels read.
//   Conv2d           OIHW [Cout][Cin][kh][kw]  -> [Cout][kh][kw][Cin] (+ [Cout][Cin2] 1x1 shortcut appended on K)
//   ConvTranspose2d  IOHW [Cin][Cout][2][2]    -> [(dy,dx,co)][Cin]     (src/UNet.py:231-233; GEMM N = 4*Cout)
//   Linear           [out][in]                 -> [in][out] (transposed, coalesced for the time-MLP kernels)
#include "kernels.h"

template <typename T>
__global__ void pack_conv_kernel(const float* __restrict__ w, int cout, int cin, int taps,
                                 const float* __restrict__ w2, int cin2, T* __restrict__ out) {
  const int ktot = taps * cin + cin2;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)cout * ktot) return;
  int k = (int)(i % ktot);
  int o = (int)(i / ktot);
  float v;
  if (k < taps * cin) {
    int tap = k / cin, c = k % cin;
    v = w[((int64_t)o * cin + c) * taps + tap];
  } else {
    v = w2[(int64_t)o * cin2 + (k - taps * cin)];
  }
  out[i] = from_float<T>(v);
}
int k_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, const float* w2_oi11, int cin2,
                       void* w_packed, int dtype, cudaStream_t st) {
  int taps = ksize * ksize;
  if (!w2_oi11) cin2 = 0;
  int64_t total = (int64_t)cout * (taps * cin + cin2);
  if (total == 0) return 0;
  int grid = (int)ceil_div64(total, 256);
  if (dtype == LDM_DT_BF16) pack_conv_kernel<bf16><<<grid, 256, 0, st>>>(w_oihw, cout, cin, taps, w2_oi11, cin2, (bf16*)w_packed);
  else pack_conv_kernel<float><<<grid, 256, 0, st>>>(w_oihw, cout, cin, taps, w2_oi11, cin2, (float*)w_packed);
  LDM_LAUNCHED("pack_conv_weight");
  return 0;
}

template <typename T>
__global__ void pack_convT_kernel(const float* __restrict__ w, int cin, int cout, T* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)4 * cout * cin) return;
  int ci = (int)(i % cin);
  int r = (int)(i / cin);
  int co = r % cout, q = r / cout;  // q = dy*2 + dx
  out[i] = from_float<T>(w[((int64_t)ci * cout + co) * 4 + q]);
}
int k_pack_convT_weight(const float* w_iohw, int cin, int cout, void* w_packed, int dtype, cudaStream_t st) {
  int64_t total = (int64_t)4 * cout * cin;
  if (total == 0) return 0;
  int grid = (int)ceil_div64(total, 256);
  if (dtype == LDM_DT_BF16) pack_convT_kernel<bf16><<<grid, 256, 0, st>>>(w_iohw, cin, cout, (bf16*)w_packed);
  else pack_convT_kernel<float><<<grid, 256, 0, st>>>(w_iohw, cin, cout, (float*)w_packed);
  LDM_LAUNCHED("pack_convT_weight");
  return 0;
}

// wt[c][col_off + r] = w[r][c]   (w is [rows][cols]; wt has row stride ld_out)
__global__ void transpose_kernel(const float* __restrict__ w, int rows, int cols, float* __restrict__ wt,
                                 int ld_out, int col_off) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * cols) return;
  int r = (int)(i % rows), c = (int)(i / rows);
  wt[(int64_t)c * ld_out + col_off + r] = w[(int64_t)r * cols + c];
}
int k_transpose_f32(const float* w, int rows, int cols, float* wt, int ld_out, int col_off, cudaStream_t st) {
  int64_t total = (int64_t)rows * cols;
  if (total == 0) return 0;
  transpose_kernel<<<(int)ceil_div64(total, 256), 256, 0, st>>>(w, rows, cols, wt, ld_out, col_off);
  LDM_LAUNCHED("transpose_f32");
  return 0;
}

__global__ void copy_f32_kernel(const float* __restrict__ s, float* __restrict__ d, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = s[i];
}
int k_copy_f32(const float* src, float* dst, int64_t n, cudaStream_t st) {
  if (n == 0) return 0;
  copy_f32_kernel<<<(int)ceil_div64(n, 256), 256, 0, st>>>(src, dst, n);
  LDM_LAUNCHED("copy_f32");
  return 0;
}

// initial conv OIHW [Cout][Cin][3][3] -> [3][3][Cin][Cout]
__global__ void pack_initial_kernel(const float* __restrict__ w, int cout, int cin, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 9 * cin * cout) return;
  int o = i % cout;
  int r = i / cout;
  int c = r % cin, tap = r / cin;
  out[i] = w[((int64_t)o * cin + c) * 9 + tap];
}
int k_pack_initial_weight(const float* w_oihw, int cout, int cin, float* out, cudaStream_t st) {
  int total = 9 * cin * cout;
  pack_initial_kernel<<<(total + 255) / 256, 256, 0, st>>>(w_oihw, cout, cin, out);
  LDM_LAUNCHED("pack_initial_weight");
  return 0;
}

__global__ void add2_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ d, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = a[i] + (b ? b[i] : 0.f);
}
int k_add2_f32(const float* a, const float* b, float* dst, int n, cudaStream_t st) {
  if (n == 0) return 0;
  add2_f32_kernel<<<(n + 255) / 256, 256, 0, st>>>(a, b, dst, n);
  LDM_LAUNCHED("add2_f32");
  return 0;
}

// Dense form of a pad-1 3x3 filter on 2x2 images: out[(po*cout + co)][(pi*cin + ci)] = w[co][ci][ky][kx] with
// (ky, kx) = (pi_y - po_y + 1, pi_x - po_x + 1) -- always a valid tap, because two pixels of a 2x2 image are at most one
// step apart.  Row-major [4*cout][4*cin] = the packed layout of a 1x1 conv with 4*cin inputs.
template <typename T>
__global__ void pack_dense2x2_kernel(const float* __restrict__ w, int cout, int cin, T* __restrict__ out, int ld_out) {
  const int64_t total = (int64_t)16 * cout * cin;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % (4 * cin)), r = (int)(i / (4 * cin));
  const int pi = k / cin, ci = k % cin, po = r / cout, co = r % cout;
  const int ky = (pi >> 1) - (po >> 1) + 1, kx = (pi & 1) - (po & 1) + 1;
  out[(int64_t)r * ld_out + k] = from_float<T>(w[((int64_t)co * cin + ci) * 9 + ky * 3 + kx]);
}
// the 1x1 shortcut in the same dense form: out[(po*cout + co)][col_off + pi*cin + ci] = (pi == po) ? wsc[co][ci] : 0
template <typename T>
__global__ void pack_blockdiag4_kernel(const float* __restrict__ wsc, int cout, int cin, T* __restrict__ out, int ld_out,
                                       int col_off) {
  const int64_t total = (int64_t)16 * cout * cin;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = (int)(i % (4 * cin)), r = (int)(i / (4 * cin));
  const int pi = k / cin, ci = k % cin, po = r / cout, co = r % cout;
  out[(int64_t)r * ld_out + col_off + k] = from_float<T>(pi == po ? wsc[(int64_t)co * cin + ci] : 0.f);
}
// w2_oi11 (optional): a K-concatenated 1x1 source of cin2 channels behind the 4*cin dense columns (row stride 4*cin + 4*cin2)
int k_pack_dense2x2_weight(const float* w_oihw, int cout, int cin, const float* w2_oi11, int cin2, void* out, int dtype,
                           cudaStream_t st) {
  const int64_t total = (int64_t)16 * cout * cin;
  if (total == 0) return 0;
  if (!w2_oi11) cin2 = 0;
  const int ld = 4 * cin + 4 * cin2;
  const int grid = (int)ceil_div64(total, 256);
  if (dtype == LDM_DT_BF16) pack_dense2x2_kernel<bf16><<<grid, 256, 0, st>>>(w_oihw, cout, cin, (bf16*)out, ld);
  else pack_dense2x2_kernel<float><<<grid, 256, 0, st>>>(w_oihw, cout, cin, (float*)out, ld);
  LDM_LAUNCHED("pack_dense2x2_weight");
  if (cin2 > 0) {
    const int grid2 = (int)ceil_div64((int64_t)16 * cout * cin2, 256);
    if (dtype == LDM_DT_BF16) pack_blockdiag4_kernel<bf16><<<grid2, 256, 0, st>>>(w2_oi11, cout, cin2, (bf16*)out, ld, 4 * cin);
    else pack_blockdiag4_kernel<float><<<grid2, 256, 0, st>>>(w2_oi11, cout, cin2, (float*)out, ld, 4 * cin);
    LDM_LAUNCHED("pack_blockdiag4_weight");
  }
  return 0;
}

// centre tap of a 3x3 filter as a 1x1 filter: out[co][ci] = w[co][ci][1][1]
template <typename T>
__global__ void pack_center_tap_kernel(const float* __restrict__ w, int64_t total, T* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) out[i] = from_float<T>(w[i * 9 + 4]);
}
int k_pack_center_tap_weight(const float* w_oihw, int cout, int cin, void* out, int dtype, cudaStream_t st) {
  const int64_t total = (int64_t)cout * cin;
  if (total == 0) return 0;
  const int grid = (int)ceil_div64(total, 256);
  if (dtype == LDM_DT_BF16) pack_center_tap_kernel<bf16><<<grid, 256, 0, st>>>(w_oihw, total, (bf16*)out);
  else pack_center_tap_kernel<float><<<grid, 256, 0, st>>>(w_oihw, total, (float*)out);
  LDM_LAUNCHED("pack_center_tap_weight");
  return 0;
}
