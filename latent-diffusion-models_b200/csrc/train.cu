// C ABI of the training-step kernels (declared in include/ldm_b200.h, "training step" section).
#include "../../include/ldm_b200.h"
#include "kernels.h"

#define RC(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

extern "C" {

int ldm_conv2d_wgrad(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw_oihw, float* dbias,
                     int batch, int height, int width, int ksize, int dtype, void* stream) {
  LDM_REQUIRE(x && dy && dw_oihw, "ldm_conv2d_wgrad: null argument");
  return k_conv_wgrad(x, ldx, cin, dy, lddy, cout, dw_oihw, dbias, batch, height, width, ksize, dtype, (cudaStream_t)stream);
}
static int64_t wgrad_nat_bytes(int cin, int cout, int ksize) {   // fp32 [tap][cout][cin] accumulator of 3x3 filters
  return ksize == 1 ? 0 : ((int64_t)ksize * ksize * cout * cin * 4 + 255) / 256 * 256;
}
int64_t ldm_conv2d_wgrad_scratch_bytes(int cin, int cout, int batch, int height, int width, int ksize, int dtype) {
  if (k_conv_wgrad_mn_applicable(cin, cout, height, width, ksize, dtype)) return wgrad_nat_bytes(cin, cout, ksize) + 256;
  if (k_conv_wgrad_tc_flat_applicable(cin, cout, batch, height, width, ksize, dtype))
    return ((int64_t)batch * height * width * (cin + (int64_t)ksize * ksize * cout) * 2 + 255) / 256 * 256 + 512 +
           wgrad_nat_bytes(cin, cout, ksize);
  if (!k_conv_wgrad_tc_applicable(cin, cout, height, width, ksize, dtype)) return 0;
  return (int64_t)batch * height * width * (cin + (ksize == 3 ? 3 : 1) * (int64_t)cout) * 2 + 4 * 256 + wgrad_nat_bytes(cin, cout, ksize);
}
// Tensor-core weight gradient: scratch (ldm_conv2d_wgrad_scratch_bytes, 256-byte aligned) receives the channel-major
// copies of x and dy and, for 3x3 filters, the fp32 accumulator in GEMM-natural layout.
int ldm_conv2d_wgrad_tc(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw_oihw, float* dbias,
                        int batch, int height, int width, int ksize, void* scratch, void* stream) {
  LDM_REQUIRE(x && dy && dw_oihw && scratch, "ldm_conv2d_wgrad_tc: null argument");
  LDM_REQUIRE(((uintptr_t)scratch & 255) == 0, "ldm_conv2d_wgrad_tc: scratch must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int hw = height * width;
  auto up = [](int64_t v) { return (v + 255) / 256 * 256; };
  float* nat = (float*)scratch;                      // first: keeps its alignment whatever follows
  char* rest = (char*)scratch + wgrad_nat_bytes(cin, cout, ksize);
  if (ksize == 1) nat = nullptr;
  if (k_conv_wgrad_mn_applicable(cin, cout, height, width, ksize, LDM_DT_BF16)) {
    // MN-major operands straight from the NHWC tensors; the bias gradient (column sums of dy) rides along as one more
    // narrow MMA against a block of ones (LDM_WGRAD_COLSUM=1: the separate column-sum kernel instead)
    static const bool sep = getenv("LDM_WGRAD_COLSUM") != nullptr;
    if (dbias && sep) RC(k_colsum(dy, lddy, dbias, batch * hw, cout, LDM_DT_BF16, st));
    return k_conv_wgrad_mn(x, ldx, cin, dy, lddy, cout, dw_oihw, sep ? nullptr : dbias, nat, batch, height, width, ksize, st);
  }
  if (k_conv_wgrad_tc_flat_applicable(cin, cout, batch, height, width, ksize, LDM_DT_BF16)) {
    char* xF = rest;
    char* dyF = xF + up((int64_t)batch * hw * cin * 2);
    RC(k_nhwc_to_flat_taps_bf16(x, ldx, xF, nullptr, batch, cin, height, width, 1, st));
    RC(k_nhwc_to_flat_taps_bf16(dy, lddy, dyF, dbias, batch, cout, height, width, ksize * ksize, st));
    return k_conv_wgrad_tc_flat(xF, cin, dyF, cout, dw_oihw, nat, batch, height, width, ksize, st);
  }
  char* xT = rest;
  char* dyT = xT + up((int64_t)batch * hw * cin * 2);
  char* dyL = ksize == 3 ? dyT + up((int64_t)batch * hw * cout * 2) : nullptr;
  char* dyR = ksize == 3 ? dyL + up((int64_t)batch * hw * cout * 2) : nullptr;
  RC(k_nhwc_to_chw_bf16(x, ldx, xT, nullptr, nullptr, nullptr, batch, cin, hw, width, st));
  RC(k_nhwc_to_chw_bf16(dy, lddy, dyT, dyL, dyR, dbias, batch, cout, hw, width, st));
  return k_conv_wgrad_tc(xT, cin, dyT, dyL, dyR, cout, dw_oihw, nat, batch, height, width, ksize, st);
}
int ldm_pack_conv_weight_pair(const float* w_oihw, int cout, int cin, int ksize, void* w_packed, void* w_packed_dgrad, int dtype,
                              void* stream) {
  LDM_REQUIRE(w_oihw && w_packed && w_packed_dgrad, "ldm_pack_conv_weight_pair: null argument");
  return k_pack_conv_weight_pair(w_oihw, cout, cin, ksize, w_packed, w_packed_dgrad, dtype, (cudaStream_t)stream);
}
int ldm_pack_conv_weight_dgrad(const float* w_oihw, int cout, int cin, int ksize, void* w_packed, int dtype, void* stream) {
  LDM_REQUIRE(w_oihw && w_packed, "ldm_pack_conv_weight_dgrad: null argument");
  return k_pack_dgrad_weight(w_oihw, cout, cin, ksize, w_packed, dtype, (cudaStream_t)stream);
}
int ldm_column_sum(const void* a, int lda, float* out, int rows, int cols, int dtype, void* stream) {
  LDM_REQUIRE(a && out, "ldm_column_sum: null argument");
  return k_colsum(a, lda, out, rows, cols, dtype, (cudaStream_t)stream);
}
int ldm_group_norm_rowvec(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                          const float* beta, const float* rowvec, int ld_rowvec, int batch, int hw, int channels, int groups,
                          float eps, int silu, int dtype, void* workspace, void* stream) {
  LDM_REQUIRE(x && y && gamma && beta, "ldm_group_norm_rowvec: null argument");
  return k_group_norm_rv(x, ldx, y, ldy, res, ldres, gamma, beta, rowvec, ld_rowvec, batch, hw, channels, groups, eps, silu,
                         dtype, workspace, (cudaStream_t)stream);
}
int64_t ldm_group_norm_backward_workspace_bytes(int batch, int hw, int channels, int groups) {
  return k_group_norm_backward_ws_bytes(batch, hw, channels, groups);
}
int ldm_group_norm_backward(const void* x, int ldx, const void* dy, int lddy, const float* gamma, const float* beta,
                            const float* rowvec, int ld_rowvec, void* dx, int lddx, float* dgamma, float* dbeta,
                            float* drowvec, int ld_drowvec, int batch, int hw, int channels, int groups, float eps, int silu,
                            int dtype, const void* forward_workspace, void* workspace, void* stream) {
  LDM_REQUIRE(x && dy && gamma && beta && dx && dgamma && dbeta, "ldm_group_norm_backward: null argument");
  if (workspace && k_group_norm_backward_streams(batch, hw, channels, groups, dtype))
    return k_group_norm_backward_stream(x, ldx, dy, lddy, gamma, beta, rowvec, ld_rowvec, dx, lddx, dgamma, dbeta, drowvec,
                                        ld_drowvec, batch, hw, channels, groups, eps, silu, forward_workspace, workspace,
                                        (cudaStream_t)stream);
  return k_group_norm_backward(x, ldx, dy, lddy, gamma, beta, rowvec, ld_rowvec, dx, lddx, dgamma, dbeta, drowvec, ld_drowvec,
                               batch, hw, channels, groups, eps, silu, dtype, (cudaStream_t)stream);
}
int ldm_max_pool2x2_backward(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int batch, int height,
                             int width, int channels, int dtype, void* stream) {
  LDM_REQUIRE(x && dy && dx, "ldm_max_pool2x2_backward: null argument");
  return k_maxpool2_backward(x, ldx, dy, lddy, dx, lddx, batch, height, width, channels, dtype, (cudaStream_t)stream);
}
int ldm_pixel_unshuffle2x2(const void* dy, int lddy, void* out, int batch, int height, int width, int channels, int dtype,
                           void* stream) {
  LDM_REQUIRE(dy && out, "ldm_pixel_unshuffle2x2: null argument");
  return k_unshuffle2(dy, lddy, out, batch, height, width, channels, dtype, (cudaStream_t)stream);
}
int64_t ldm_linear_attention_backward_workspace_bytes(int batch) { return k_linear_attention_backward_ws_bytes(batch); }
int ldm_linear_attention_backward(const void* qkv, const void* dout, void* dqkv, int batch, int n_tokens, int dtype,
                                  void* workspace, void* stream) {
  LDM_REQUIRE(qkv && dout && dqkv, "ldm_linear_attention_backward: null argument");
  if (workspace && k_linear_attention_backward_mma_applicable(n_tokens, dtype))
    return k_linear_attention_backward_mma(qkv, dout, dqkv, batch, n_tokens, workspace, (cudaStream_t)stream);
  return k_linear_attention_backward(qkv, dout, dqkv, batch, n_tokens, dtype, (cudaStream_t)stream);
}
int ldm_attention_backward(const void* qkv, const void* dout, void* dqkv, int batch, int n_tokens, int dtype, void* stream) {
  LDM_REQUIRE(qkv && dout && dqkv, "ldm_attention_backward: null argument");
  return k_attention_backward(qkv, dout, dqkv, batch, n_tokens, dtype, (cudaStream_t)stream);
}
int ldm_initial_conv(const float* x_nchw, const float* w_oihw, const float* bias, void* y, int batch, int cin, int cout, int height,
                     int width, int dtype, float* w_scratch, void* stream) {
  LDM_REQUIRE(x_nchw && w_oihw && bias && y && w_scratch, "ldm_initial_conv: null argument");
  RC(k_pack_initial_weight(w_oihw, cout, cin, w_scratch, (cudaStream_t)stream));
  return k_initial_conv(x_nchw, batch, w_scratch, bias, y, batch, cin, cout, height, width, dtype, (cudaStream_t)stream);
}
int64_t ldm_initial_conv_wgrad_scratch_bytes(int batch, int cin, int cout, int height, int width) {
  return k_initial_conv_wgrad_scratch_bytes(batch, cin, cout, height, width);
}
int ldm_initial_conv_wgrad(const float* x_nchw, const void* dy, float* dw_oihw, float* dbias, int batch, int cin, int cout, int height,
                           int width, int dtype, void* scratch, void* stream) {
  LDM_REQUIRE(x_nchw && dy && dw_oihw, "ldm_initial_conv_wgrad: null argument");
  return k_initial_conv_wgrad(x_nchw, dy, dw_oihw, dbias, batch, cin, cout, height, width, dtype, scratch, (cudaStream_t)stream);
}
int ldm_final_conv(const void* x, int ldx, const float* w, const float* bias, float* y_nchw, int batch, int cin, int cout, int hw,
                   int dtype, void* stream) {
  LDM_REQUIRE(x && w && bias && y_nchw, "ldm_final_conv: null argument");
  return k_final_conv(x, ldx, w, bias, y_nchw, batch, cin, cout, hw, dtype, (cudaStream_t)stream);
}
int ldm_final_conv_backward(const float* dout_nchw, const void* x, int ldx, const float* w, void* dx, float* dw, float* db, int batch,
                            int cin, int cout, int hw, int dtype, void* stream) {
  LDM_REQUIRE(dout_nchw && x && w && dx && dw && db, "ldm_final_conv_backward: null argument");
  return k_final_conv_backward(dout_nchw, x, ldx, w, dx, dw, db, batch, cin, cout, hw, dtype, (cudaStream_t)stream);
}

// ---- time embedding (src/UNet.py:23-44,263-268,373-376) and the mlp_t projections (:70-73,90-93), fp32
int64_t ldm_time_workspace_bytes(int batch, int D, int total) {
  return (int64_t)batch * (D / 4 + 4 * (int64_t)D + total) * 4 + 1024;
}
static void time_ws(float* ws, int batch, int D, float*& emb, float*& pre1, float*& h1, float*& t0, float*& t1) {
  emb = ws; pre1 = emb + (int64_t)batch * (D / 4); h1 = pre1 + (int64_t)batch * D; t0 = h1 + (int64_t)batch * D;
  t1 = t0 + (int64_t)batch * D;
}
int ldm_time_embed(const int64_t* t, const int64_t* y, int y_len, const float* w1, const float* b1, const float* w3,
                   const float* b3, const float* label_emb, float* temb, int batch, int D, void* workspace, void* stream) {
  LDM_REQUIRE(t && w1 && b1 && w3 && b3 && temb && workspace, "ldm_time_embed: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  float *emb, *pre1, *h1, *t0, *t1;
  time_ws((float*)workspace, batch, D, emb, pre1, h1, t0, t1);
  const int Din = D / 4;
  RC(k_sinusoid(t, emb, batch, Din, st));
  RC(k_gemm_f32(emb, Din, 1, w1, 1, Din, pre1, D, batch, D, Din, 0, st));   // pre1 = emb w1^T
  RC(k_ew(pre1, b1, h1, batch, D, 1, st));                                  // h1 = gelu(pre1 + b1)
  RC(k_gemm_f32(h1, D, 1, w3, 1, D, temb, D, batch, D, D, 0, st));          // temb = h1 w3^T
  RC(k_ew(temb, b3, temb, batch, D, 0, st));
  if (y && y_len > 0) RC(k_rows_gather_add(temb, y, y_len, label_emb, batch, D, st));
  return 0;
}
// gradients are ACCUMULATED into dw1/db1/dw3/db3/dlabel (zero them first)
int ldm_time_embed_backward(const int64_t* t, const int64_t* y, int y_len, const float* w1, const float* b1, const float* w3,
                            const float* dtemb, float* dw1, float* db1, float* dw3, float* db3, float* dlabel, int batch, int D,
                            void* workspace, void* stream) {
  LDM_REQUIRE(t && w1 && b1 && w3 && dtemb && dw1 && db1 && dw3 && db3 && workspace, "ldm_time_embed_backward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  float *emb, *pre1, *h1, *t0, *t1;
  time_ws((float*)workspace, batch, D, emb, pre1, h1, t0, t1);
  const int Din = D / 4;
  RC(k_sinusoid(t, emb, batch, Din, st));
  RC(k_gemm_f32(emb, Din, 1, w1, 1, Din, pre1, D, batch, D, Din, 0, st));
  RC(k_ew(pre1, b1, h1, batch, D, 1, st));                                   // h1 = gelu(pre1 + b1)
  RC(k_ew(pre1, b1, pre1, batch, D, 0, st));                                 // pre1 += b1 (the GELU's input)
  // label embedding rows, output bias, second Linear
  if (y && y_len > 0 && dlabel) RC(k_rows_scatter_add(dtemb, y, y_len, dlabel, batch, D, st));
  RC(k_colsum(dtemb, D, db3, batch, D, LDM_DT_F32, st));
  RC(k_gemm_f32(dtemb, 1, D, h1, D, 1, dw3, D, D, D, batch, 1, st));         // dw3[o][k] += sum_b dtemb[b][o] h1[b][k]
  RC(k_gemm_f32(dtemb, D, 1, w3, D, 1, t0, D, batch, D, D, 0, st));          // dh1 = dtemb w3
  RC(k_ew(pre1, t0, t1, batch, D, 3, st));                                   // dpre1 = dh1 * gelu'(pre1)
  RC(k_colsum(t1, D, db1, batch, D, LDM_DT_F32, st));
  RC(k_gemm_f32(t1, 1, D, emb, Din, 1, dw1, Din, D, Din, batch, 1, st));     // dw1[o][i] += sum_b dpre1[b][o] emb[b][i]
  return 0;
}

// tproj = SiLU(temb) w^T + bias   with w = the mlp_t weights of all blocks stacked: [total][D]
int ldm_time_proj(const float* temb, const float* w, const float* bias, float* tproj, int batch, int D, int total,
                  void* workspace, void* stream) {
  LDM_REQUIRE(temb && w && bias && tproj && workspace, "ldm_time_proj: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  float* s = (float*)workspace;
  RC(k_ew(temb, nullptr, s, batch, D, 2, st));
  RC(k_gemm_f32(s, D, 1, w, 1, D, tproj, total, batch, total, D, 0, st));
  RC(k_ew(tproj, bias, tproj, batch, total, 0, st));
  return 0;
}
// dw / dbias are ACCUMULATED; dtemb is overwritten
int ldm_time_proj_backward(const float* temb, const float* w, const float* dtproj, float* dw, float* dbias, float* dtemb,
                           int batch, int D, int total, void* workspace, void* stream) {
  LDM_REQUIRE(temb && w && dtproj && dw && dbias && dtemb && workspace, "ldm_time_proj_backward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  float* s = (float*)workspace;
  float* ds = s + (int64_t)batch * D;
  RC(k_ew(temb, nullptr, s, batch, D, 2, st));
  RC(k_colsum(dtproj, total, dbias, batch, total, LDM_DT_F32, st));
  RC(k_gemm_f32(dtproj, 1, total, s, D, 1, dw, D, total, D, batch, 1, st));  // dw[o][k] += sum_b dtproj[b][o] s[b][k]
  RC(k_gemm_f32(dtproj, total, 1, w, D, 1, ds, D, batch, D, total, 0, st));  // ds = dtproj w
  RC(k_ew(temb, ds, dtemb, batch, D, 4, st));                                // dtemb = ds * silu'(temb)
  return 0;
}

int64_t ldm_linear_attention_prenorm_scratch_bytes(int batch) {
  return 384 * 64 * 2 + 768 * 4 + 1024 + 64 * 128 * 2 + 256 * 64 * 2 + 1024 + ((int64_t)(batch > 0 ? batch : 1) * 4 + 255) / 256 * 256 +
         k_group_norm_ws_bytes(batch > 0 ? batch : 1, 1);
}
int ldm_linear_attention_prenorm(const void* x, int ldx, int cin, const float* w_qkv, const float* gamma, const float* beta,
                                 float eps, void* out, int batch, int n_tokens, int impl, void* scratch, int64_t scratch_bytes,
                                 void* stream) {
  LDM_REQUIRE(x && w_qkv && gamma && beta && out && scratch, "ldm_linear_attention_prenorm: null argument");
  LDM_REQUIRE(cin == 64, "ldm_linear_attention_prenorm: the fused kernels take 64 input channels");
  LDM_REQUIRE(scratch_bytes >= ldm_linear_attention_prenorm_scratch_bytes(batch) && ((uintptr_t)scratch & 255) == 0,
              "ldm_linear_attention_prenorm: scratch too small or unaligned");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* s = (uint8_t*)scratch;
  void* wfold = s;
  float* uv = (float*)(s + 384 * 64 * 2);
  void* gnws = s + 384 * 64 * 2 + 768 * 4 + 1024 - ((384 * 64 * 2 + 768 * 4) % 1024);
  if (int rc = k_fold_prenorm_qkv(w_qkv, gamma, beta, cin, wfold, uv, st)) return rc;
  int splits = 1;
  if (int rc = k_group_norm_stats(x, ldx, batch, n_tokens, cin, 1, gnws, &splits, st)) return rc;
  if (impl == 0) {
    LDM_REQUIRE(k_linear_attention_tc_applicable(cin, n_tokens, LDM_DT_BF16), "ldm_linear_attention_prenorm: tcgen05 kernel needs N %% 128 == 0");
    return k_linear_attention_tc(x, ldx, wfold, uv, gnws, splits, eps, out, batch, n_tokens, st);
  }
  return k_linear_attention_qkv_prenorm(x, ldx, cin, wfold, uv, gnws, splits, eps, out, batch, n_tokens, LDM_DT_BF16, st);
}
int ldm_linear_attention_prenorm_to_out(const void* x, int ldx, int cin, const float* w_qkv, const float* gamma, const float* beta,
                                        float eps, const float* w_out, const float* b_out, void* y, int ldy, float* ystats,
                                        int batch, int n_tokens, void* scratch, int64_t scratch_bytes, void* stream) {
  LDM_REQUIRE(x && w_qkv && gamma && beta && w_out && b_out && y && ystats && scratch, "ldm_linear_attention_prenorm_to_out: null argument");
  LDM_REQUIRE(cin == 64 && k_linear_attention_tc_applicable(cin, n_tokens, LDM_DT_BF16),
              "ldm_linear_attention_prenorm_to_out: needs 64 channels and a multiple of 128 tokens");
  LDM_REQUIRE(scratch_bytes >= ldm_linear_attention_prenorm_scratch_bytes(batch) && ((uintptr_t)scratch & 255) == 0,
              "ldm_linear_attention_prenorm_to_out: scratch too small or unaligned");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* s = (uint8_t*)scratch;
  void* wfold = s;
  float* uv = (float*)(s + 384 * 64 * 2);
  void* wout = s + 384 * 64 * 2 + 768 * 4 + 1024;                       // packed [64][128] bf16 (16 KB)
  uint8_t* s2 = s + 384 * 64 * 2 + 768 * 4 + 1024 + 64 * 128 * 2;
  void* ucat = s2;                                                      // to_out folded over Wv (32 KB) | c12 | flags
  float* c12 = (float*)(s2 + 256 * 64 * 2);
  int* flags = (int*)(s2 + 256 * 64 * 2 + 1024);
  void* gnws = s2 + 256 * 64 * 2 + 1024 + ((int64_t)(batch > 0 ? batch : 1) * 4 + 255) / 256 * 256;
  if (int rc = k_fold_prenorm_qkv(w_qkv, gamma, beta, cin, wfold, uv, st)) return rc;
  if (int rc = k_pack_conv_weight(w_out, 64, 128, 1, nullptr, 0, wout, LDM_DT_BF16, st)) return rc;
  if (int rc = k_fold_to_out(w_qkv, gamma, uv, w_out, ucat, c12, st)) return rc;
  int splits = 1;
  if (int rc = k_group_norm_stats(x, ldx, batch, n_tokens, cin, 1, gnws, &splits, st)) return rc;
  LinAttnOut f;
  f.wout = wout; f.bout = b_out; f.y = y; f.ldy = ldy; f.ystats = ystats; f.ystats_bytes = (int64_t)batch * (n_tokens / 16) * 8;
  f.nslots_out = nullptr;
  f.ucat = ucat; f.c12 = c12; f.flags = flags;
  return k_linear_attention_tc(x, ldx, wfold, uv, gnws, splits, eps, nullptr, batch, n_tokens, st, &f);
}
int ldm_linear_attention_qkv(const void* xn, int ldx, int cin, const void* wqkv_packed, void* out, int batch, int n_tokens,
                             int dtype, void* stream) {
  LDM_REQUIRE(xn && wqkv_packed && out, "ldm_linear_attention_qkv: null argument");
  return k_linear_attention_qkv(xn, ldx, cin, wqkv_packed, out, batch, n_tokens, dtype, (cudaStream_t)stream);
}
int ldm_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream) {
  LDM_REQUIRE(a && b && out, "ldm_add: null argument");
  return k_add(a, b, out, n, dtype, (cudaStream_t)stream);
}
int ldm_copy_channels(const void* src, int ld_src, void* dst, int ld_dst, int channels, int64_t rows, int dtype, void* stream) {
  LDM_REQUIRE(src && dst, "ldm_copy_channels: null argument");
  return k_copy_channels(src, ld_src, dst, ld_dst, channels, rows, dtype, (cudaStream_t)stream);
}


// ---- optimizer step, validation loss and output stage
int ldm_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, int step, double grad_scale, void* stream) {
  LDM_REQUIRE(param && grad && exp_avg && exp_avg_sq, "ldm_adam_step: null argument");
  return k_adam_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale, (cudaStream_t)stream);
}
int ldm_images_to_uint8(const float* x_nchw, uint8_t* out_nhwc, int batch, int channels, int hw, int convention, void* stream) {
  LDM_REQUIRE(x_nchw && out_nhwc, "ldm_images_to_uint8: null argument");
  return k_images_to_u8(x_nchw, out_nhwc, batch, channels, hw, convention, (cudaStream_t)stream);
}
int ldm_mse(const float* a, const float* b, float* out_scalar, int64_t n, void* stream) {
  LDM_REQUIRE(a && b && out_scalar, "ldm_mse: null argument");
  return k_mse(a, b, out_scalar, n, (cudaStream_t)stream);
}
int ldm_mse_backward(const float* pred, const float* target, const float* grad_loss, float* dpred, int64_t n, void* stream) {
  LDM_REQUIRE(pred && target && dpred, "ldm_mse_backward: null argument");
  return k_mse_backward(pred, target, grad_loss, dpred, n, (cudaStream_t)stream);
}


// ---- first-stage autoencoder (src/Autoencoder.py), the pieces beyond the UNet's conv / GroupNorm kernels
int ldm_upsample_nearest2x(const void* x, int ldx, void* y, int ldy, int batch, int height, int width, int channels, int dtype,
                           void* stream) {
  LDM_REQUIRE(x && y, "ldm_upsample_nearest2x: null argument");
  return k_upsample_nearest2x(x, ldx, y, ldy, batch, height, width, channels, dtype, (cudaStream_t)stream);
}
int ldm_downsample_pick(const void* x, int ldx, void* y, int ldy, int batch, int height, int width, int channels, int dtype,
                        void* stream) {
  LDM_REQUIRE(x && y, "ldm_downsample_pick: null argument");
  return k_pick_odd(x, ldx, y, ldy, batch, height, width, channels, dtype, (cudaStream_t)stream);
}
int ldm_attention_single_head(const void* qkv, void* out, int batch, int n_tokens, int channels, int dtype, void* stream) {
  LDM_REQUIRE(qkv && out, "ldm_attention_single_head: null argument");
  return k_attention_single_head(qkv, out, batch, n_tokens, channels, dtype, (cudaStream_t)stream);
}
int ldm_gaussian_distribution(const void* moments, int ld, const float* eps, float* mu, float* log_var, float* sigma, float* z,
                              int batch, int z_channels, int hw, int dtype, void* stream) {
  LDM_REQUIRE(moments, "ldm_gaussian_distribution: null argument");
  return k_gaussian(moments, ld, eps, mu, log_var, sigma, z, batch, z_channels, hw, dtype, (cudaStream_t)stream);
}

}  // extern "C"
