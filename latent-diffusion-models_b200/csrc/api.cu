// C-ABI glue: error reporting, launch accounting and the kernel-level entry points of include/ldm_b200.h.
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/ldm_b200.h"
#include "kernels.h"

static thread_local char g_err[1024] = "";
std::atomic<long long> g_ldm_launches{0};
// measured on B200 (profiles/README.md): the per-timestep chain is 3.8 % SLOWER with programmatic dependent launch
// (2988 vs 2879 us), so it is opt-in
int g_ldm_pdl = [] { const char* e = getenv("LDM_PDL"); return e ? atoi(e) : 0; }();

int ldm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}

extern "C" {

int ldm_abi_version(void) { return LDM_B200_ABI_VERSION; }
const char* ldm_last_error(void) { return g_err; }
int64_t ldm_launch_count(void) { return g_ldm_launches.load(); }
void ldm_reset_launch_count(void) { g_ldm_launches.store(0); }

int ldm_q_sample(const float* x0, const int64_t* t, const float* alpha_bar, int n_steps, const float* eps,
                 float* eps_out, float* xt, int batch, int64_t n_per_sample, uint64_t seed, uint64_t sample_offset,
                 void* stream) {
  LDM_REQUIRE(x0 && t && alpha_bar && xt, "ldm_q_sample: null argument");
  return k_q_sample(x0, t, alpha_bar, n_steps, eps, eps_out, xt, batch, n_per_sample, seed, sample_offset,
                    (cudaStream_t)stream);
}

int ldm_p_sample(const float* xt, const float* eps_cond, const float* eps_uncond, float cfg_scale,
                 const int64_t* t_dev, int t_len, const float* coef, int n_steps, const float* noise, uint64_t seed,
                 uint64_t sample_offset, float* out, int batch, int64_t n_per_sample, void* stream) {
  LDM_REQUIRE(xt && eps_cond && t_dev && coef && out, "ldm_p_sample: null argument");
  LDM_REQUIRE(t_len == 1 || t_len == batch, "ldm_p_sample: t has %d entries for a batch of %d", t_len, batch);
  return k_p_sample(xt, eps_cond, eps_uncond, cfg_scale, t_dev, t_len == 1 ? 0 : 1, coef, n_steps, noise, 0, seed, sample_offset, out,
                    batch, n_per_sample, (cudaStream_t)stream);
}

int ldm_build_coef_table(const float* beta, const float* alpha, const float* alpha_bar, int n_steps, float* coef,
                         void* stream) {
  LDM_REQUIRE(beta && alpha && alpha_bar && coef, "ldm_build_coef_table: null argument");
  return k_build_coef(beta, alpha, alpha_bar, n_steps, coef, (cudaStream_t)stream);
}

int ldm_randn(float* out, int batch, int64_t n_per_sample, uint64_t seed, uint64_t sample_offset,
              uint64_t stream_id, void* stream) {
  LDM_REQUIRE(out, "ldm_randn: null argument");
  return k_randn(out, batch, n_per_sample, seed, sample_offset, stream_id, (cudaStream_t)stream);
}

int ldm_group_norm(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                   const float* beta, int batch, int hw, int channels, int groups, float eps, int silu, int dtype,
                   void* workspace, void* stream) {
  LDM_REQUIRE(x && y && gamma && beta, "ldm_group_norm: null argument");
  return k_group_norm(x, ldx, y, ldy, res, ldres, gamma, beta, batch, hw, channels, groups, eps, silu, dtype,
                      workspace, (cudaStream_t)stream);
}
int64_t ldm_group_norm_workspace_bytes(int batch, int groups) { return k_group_norm_ws_bytes(batch, groups); }

int64_t ldm_conv2d_gn_scratch_bytes(int batch) { return (int64_t)(batch > 0 ? batch : 1) * 256 * 16; }

int ldm_conv2d_gn(const void* x, int ldx, int cin, const void* x2, int ldx2, int cin2, const void* w_packed,
                  const float* bias, const void* res, int ldres, void* y, int ldy, int cout, int batch, int height,
                  int width, int ksize, int gn_mode, int groups, float eps, int silu, const float* gamma,
                  const float* beta, const float* gn_rowvec, int ld_gn_rowvec, const void* gn_res, int ld_gn_res,
                  int nvar, int var_rows, void* scratch, int64_t scratch_bytes, unsigned tag, int* nslots_out,
                  void* stream) {
  LDM_REQUIRE(x && w_packed && y, "ldm_conv2d_gn: null argument");
  LDM_REQUIRE(gn_mode == 1 || gn_mode == 2, "ldm_conv2d_gn: gn_mode must be 1 or 2");
  ConvArgs a;
  a.x = x; a.ldx = ldx; a.cin = cin; a.x2 = x2; a.ldx2 = ldx2; a.cin2 = x2 ? cin2 : 0; a.w = w_packed; a.bias = bias;
  a.rowvec = nullptr; a.ld_rowvec = 0; a.res = res; a.ldres = ldres; a.y = y; a.ldy = ldy; a.cout = cout;
  a.batch = batch; a.height = height; a.width = width; a.ksize = ksize; a.up2 = 0; a.dtype = LDM_DT_BF16;
  a.gn.mode = gn_mode; a.gn.groups = groups; a.gn.eps = eps; a.gn.silu = silu; a.gn.gamma = gamma; a.gn.beta = beta;
  a.gn.rowvec = gn_rowvec; a.gn.ld_rowvec = ld_gn_rowvec; a.gn.res = gn_res; a.gn.ldres = ld_gn_res;
  a.gn.nvar = nvar > 0 ? nvar : 1; a.gn.var_rows = var_rows; a.gn.scratch = scratch; a.gn.scratch_bytes = scratch_bytes;
  a.gn.tag = tag; a.gn.nslots_out = nslots_out;
  if (int rc = k_conv_tc_prepare()) return rc;
  return k_conv(a, 0, (cudaStream_t)stream);
}

int ldm_conv2d(const void* x, int ldx, int cin, const void* x2, int ldx2, int cin2, const void* w_packed,
               const float* bias, const float* rowvec, int ld_rowvec, const void* res, int ldres, void* y, int ldy,
               int cout, int batch, int height, int width, int ksize, int dtype, int impl, void* stream) {
  LDM_REQUIRE(x && w_packed && y, "ldm_conv2d: null argument");
  ConvArgs a;
  a.x = x; a.ldx = ldx; a.cin = cin; a.x2 = x2; a.ldx2 = ldx2; a.cin2 = x2 ? cin2 : 0; a.w = w_packed; a.bias = bias;
  a.rowvec = rowvec; a.ld_rowvec = ld_rowvec; a.res = res; a.ldres = ldres; a.y = y; a.ldy = ldy; a.cout = cout;
  a.batch = batch; a.height = height; a.width = width; a.ksize = ksize; a.up2 = 0; a.dtype = dtype;
  if (dtype == LDM_DT_BF16 && impl == 0) {
    if (int rc = k_conv_tc_prepare()) return rc;
  }
  return k_conv(a, impl, (cudaStream_t)stream);
}

int ldm_conv_transpose2x2(const void* x, int ldx, int cin, const void* w_packed, const float* bias, void* y, int ldy,
                          int cout, int batch, int height, int width, int dtype, int impl, void* stream) {
  LDM_REQUIRE(x && w_packed && y, "ldm_conv_transpose2x2: null argument");
  ConvArgs a;
  a.x = x; a.ldx = ldx; a.cin = cin; a.x2 = nullptr; a.ldx2 = 0; a.cin2 = 0; a.w = w_packed; a.bias = bias;
  a.rowvec = nullptr; a.ld_rowvec = 0; a.res = nullptr; a.ldres = 0; a.y = y; a.ldy = ldy; a.cout = 4 * cout;
  a.batch = batch; a.height = height; a.width = width; a.ksize = 1; a.up2 = 1; a.dtype = dtype;
  if (dtype == LDM_DT_BF16 && impl == 0) {
    if (int rc = k_conv_tc_prepare()) return rc;
  }
  return k_conv(a, impl, (cudaStream_t)stream);
}

int ldm_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, const float* w2_oi11, int cin2,
                         void* w_packed, int dtype, void* stream) {
  LDM_REQUIRE(w_oihw && w_packed, "ldm_pack_conv_weight: null argument");
  return k_pack_conv_weight(w_oihw, cout, cin, ksize, w2_oi11, cin2, w_packed, dtype, (cudaStream_t)stream);
}
int ldm_pack_conv_transpose_weight(const float* w_iohw, int cin, int cout, void* w_packed, int dtype, void* stream) {
  LDM_REQUIRE(w_iohw && w_packed, "ldm_pack_conv_transpose_weight: null argument");
  return k_pack_convT_weight(w_iohw, cin, cout, w_packed, dtype, (cudaStream_t)stream);
}

int ldm_linear_attention(const void* qkv, void* out, int batch, int n_tokens, int dtype, void* stream) {
  LDM_REQUIRE(qkv && out, "ldm_linear_attention: null argument");
  return k_linear_attention(qkv, out, batch, n_tokens, dtype, (cudaStream_t)stream);
}
int ldm_attention(const void* qkv, void* out, int batch, int n_tokens, int dtype, void* stream) {
  LDM_REQUIRE(qkv && out, "ldm_attention: null argument");
  return k_attention(qkv, out, batch, n_tokens, dtype, (cudaStream_t)stream);
}
int ldm_max_pool2x2(const void* x, int ldx, void* y, int ldy, int batch, int height, int width, int channels,
                    int dtype, void* stream) {
  LDM_REQUIRE(x && y, "ldm_max_pool2x2: null argument");
  return k_maxpool2(x, ldx, y, ldy, batch, height, width, channels, dtype, (cudaStream_t)stream);
}
int ldm_nchw_to_nhwc(const float* x, void* y, int batch, int channels, int hw, int dtype, void* stream) {
  LDM_REQUIRE(x && y, "ldm_nchw_to_nhwc: null argument");
  return k_nchw_to_nhwc(x, y, batch, channels, hw, dtype, (cudaStream_t)stream);
}
int ldm_nhwc_to_nchw(const void* x, int ldx, float* y, int batch, int channels, int hw, int dtype, void* stream) {
  LDM_REQUIRE(x && y, "ldm_nhwc_to_nchw: null argument");
  return k_nhwc_to_nchw(x, ldx, y, batch, channels, hw, dtype, (cudaStream_t)stream);
}

}  // extern "C"
