// Internal host-side launchers shared between translation units (not part of the C ABI).
#pragma once
#include "common.cuh"

// ---- elementwise.cu
int k_nchw_to_nhwc(const float* x, void* y, int batch, int channels, int hw, int dtype, cudaStream_t st);
int k_nhwc_to_nchw(const void* x, int ldx, float* y, int batch, int channels, int hw, int dtype, cudaStream_t st);
int k_maxpool2(const void* x, int ldx, void* y, int ldy, int batch, int height, int width, int channels,
               int dtype, cudaStream_t st);
int k_build_coef(const float* beta, const float* alpha, const float* alpha_bar, int n_steps, float* coef,
                 cudaStream_t st);
int k_randn(float* out, int batch, int64_t n_per_sample, uint64_t seed, uint64_t sample_offset,
            uint64_t stream_id, cudaStream_t st);
int k_q_sample(const float* x0, const int64_t* t, const float* alpha_bar, int n_steps, const float* eps,
               float* eps_out, float* xt, int batch, int64_t n_per_sample, uint64_t seed,
               uint64_t sample_offset, cudaStream_t st);
// noise_t_stride: elements between consecutive timesteps in `noise` (0: `noise` is this step's tensor)
// t_stride: 0 = one batch-constant timestep at t_dev[0]; 1 = per-sample t_dev[b] (noise branch still on t_dev[0])
int k_p_sample(const float* xt, const float* eps_c, const float* eps_u, float cfg_scale, const int64_t* t_dev,
               int t_stride, const float* coef, int n_steps, const float* noise, int64_t noise_t_stride, uint64_t seed,
               uint64_t sample_offset, float* out, int batch, int64_t n_per_sample, cudaStream_t st,
               const uint64_t* seed_dev = nullptr);   // seed_dev: device {seed, sample_offset} overriding the two values
int k_set_i64(int64_t* p, int64_t v, cudaStream_t st);
int k_add_i64(int64_t* p, int64_t v, cudaStream_t st);

// time embedding (src/UNet.py:23-44,263-268,373-376): temb [batch,D] fp32
//   w1t [D/4][D], w3t [D][D] are the transposed Linear weights; label_emb [num_classes][D]
int k_time_embed(const int64_t* t, const int64_t* t_scalar, const int64_t* y, int y_len, int y_rows,
                 const float* w1t, const float* b1, const float* w3t, const float* b3, const float* label_emb,
                 float* temb, int batch, int D, int table_classes, cudaStream_t st);
// table_classes > 0: row b carries the label embedding of class b (rows >= table_classes none); y is ignored
// batch-constant t: s_tab [R][D] = SiLU(time_mlp(t) + label row r), tproj_tab [R][total] = s_tab . tproj_w^T + b
//   w1 [D][D/4], w3 [D][D], tproj_w [total][D] in the original PyTorch [out][in] layout
int k_time_table(const int64_t* t_scalar, const float* w1, const float* b1, const float* w3, const float* b3,
                 const float* label_emb, const float* tproj_w, const float* tproj_b, float* s_tab, float* tproj_tab,
                 int R, int n_classes, int D, int total, cudaStream_t st);
int k_tproj_gather(const float* tab, const int64_t* y, int y_len, int y_rows, int n_classes, float* tproj, int batch,
                   int total, cudaStream_t st);
// tproj[b][o] = sum_k silu(temb[b][k]) wt[k][o] + bias[o]   (src/UNet.py:70-73,90-93), o < total
int k_time_proj(const float* temb, const float* wt, const float* bias, float* tproj, int batch, int D,
                int total, cudaStream_t st);

// initial 3x3 conv, tiny Cin: fp32 NCHW in (row b reads image b % x_batch) -> NHWC out (src/UNet.py:331,378)
//   w [3][3][Cin][Cout] fp32
int k_initial_conv(const float* x, int x_batch, const float* w, const float* bias, void* y, int batch, int cin,
                   int cout, int height, int width, int dtype, cudaStream_t st);
// final 1x1 conv, tiny Cout: NHWC in -> fp32 NCHW out (src/UNet.py:347); w [Cout][Cin] fp32
int k_final_conv(const void* x, int ldx, const float* w, const float* bias, float* y, int batch, int cin,
                 int cout, int hw, int dtype, cudaStream_t st);

// ---- groupnorm.cu
int64_t k_group_norm_ws_bytes(int batch, int groups);
int k_group_norm(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                 const float* beta, int batch, int hw, int channels, int groups, float eps, int silu, int dtype,
                 void* workspace, cudaStream_t st);

// same, with x + rowvec[n][c] (fp32 [batch][ld_rowvec]) normalised instead of x
int k_group_norm_rv(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                    const float* beta, const float* rowvec, int ld_rowvec, int batch, int hw, int channels, int groups,
                    float eps, int silu, int dtype, void* workspace, cudaStream_t st);

int k_group_norm_stats(const void* x, int ldx, int batch, int hw, int channels, int groups, void* workspace, int* splits_out,
                       cudaStream_t st);

// same; x_mod > 0: sample n reads image n % x_mod of x (bf16 streaming kernels only)
int k_group_norm_mod(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                     const float* beta, const float* rowvec, int ld_rowvec, int batch, int hw, int channels, int groups,
                     float eps, int silu, int dtype, void* workspace, int x_mod, cudaStream_t st);
bool k_group_norm_streams(int hw, int channels, int dtype);
// max-pool 2x2 + GroupNorm (+SiLU) in one kernel: pooled tensor and its normalised form (encoder level transitions)
bool k_pool_group_norm_applicable(int H, int W, int C, int groups, int dtype);
int k_pool_group_norm(const void* x, int ldx, int H, int W, int C, void* pool, int ldp, void* y, int ldy, const float* gamma,
                      const float* beta, int groups, float eps, int silu, int batch, cudaStream_t st);
int k_group_norm_apply_raw(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                           const float* beta, const float* rowvec, int ld_rowvec, int batch, int hw, int channels, int groups,
                           float eps, int silu, const void* part, int nslots, int x_mod, cudaStream_t st);

// ---- attention.cu
int k_linear_attention(const void* qkv, void* out, int batch, int n_tokens, int dtype, cudaStream_t st);
// LinearAttention with the to_qkv 1x1 convolution fused in (src/UNet.py:145,149-163): xn [B,N,cin] -> out [B,N,128];
// wqkv = packed [384][cin] bf16 (no bias).  Only cin == 64, bf16, N % 16 == 0 (the full-resolution sites).
bool k_linear_attention_qkv_applicable(int cin, int n_tokens, int dtype);
int k_linear_attention_qkv(const void* xn, int ldx, int cin, const void* wqkv, void* out, int batch, int n_tokens,
                           int dtype, cudaStream_t st);
// Same with the PreNorm GroupNorm(1, C) folded in (src/UNet.py:106-110): x is the RAW block input, wfold = wqkv diag(gamma)
// (bf16 [384][cin]), uv = {row sums of wfold, wqkv beta} (fp32 [2][384], see k_fold_prenorm_qkv), gn_part / gn_splits the
// statistics left by k_group_norm_stats (groups = 1).
int k_linear_attention_qkv_prenorm(const void* x, int ldx, int cin, const void* wfold, const float* uv, const void* gn_part,
                                   int gn_splits, float eps, void* out, int batch, int n_tokens, int dtype, cudaStream_t st);
// The same on tcgen05 / TMEM / TMA (linattn_tc.cu): N % 128 == 0; uv == null: plain to_qkv weights on an already normalised x
bool k_linear_attention_tc_applicable(int cin, int n_tokens, int dtype);
// fuse != null: LinearAttention.to_out's 1x1 convolution (src/UNet.py:146) is folded in -- y [B, N, ldy] (64 channels, bf16) =
// out Wout^T + bout, and ystats receives GroupNorm(1, 64) partial sums {S, Q} of y ([batch][nslots], the ConvGn mode-1 layout)
struct LinAttnOut {
  const void* wout;            // packed [64][128] bf16 (k_pack_conv_weight of to_out.0.weight)
  const float* bout;           // [64]
  void* y; int ldy;
  void* ystats; int64_t ystats_bytes;
  int* nslots_out;
  // optional: to_out's GroupNorm(1, 64) and the Residual add too (src/UNet.py:147,:20): o [B, N, ldo] = x + GroupNorm(y)
  const float* og = nullptr; const float* ob = nullptr; void* o = nullptr; int ldo = 0; float o_eps = 1e-5f;
  // optional (all three, and no `o`): the optimistic kernel linattn_tc2_kernel runs first -- ucat / c12 from k_fold_to_out,
  // flags [batch] ints of scratch (samples it could not finish in range are redone by the exact kernel in the same call)
  const void* ucat = nullptr; const float* c12 = nullptr; int* flags = nullptr;
};
int k_fold_to_out(const float* wqkv /*[384][64] fp32*/, const float* gamma, const float* uv /*k_fold_prenorm_qkv's*/,
                  const float* wout /*[64][128] fp32*/, void* ucat /*bf16 [256][64]*/, float* c12 /*[64]*/, cudaStream_t st);
int k_linear_attention_tc(const void* x, int ldx, const void* wqkv, const float* uv, const void* gn_part, int gn_splits, float eps,
                          void* out, int batch, int n_tokens, cudaStream_t st, const LinAttnOut* fuse = nullptr);
int k_fold_prenorm_qkv(const float* wqkv /*[384][cin] fp32*/, const float* gamma, const float* beta, int cin, void* wfold,
                       float* uv, cudaStream_t st);
// LinearAttention core reading the RAW projection wfold x of the un-normalised block input: the PreNorm GroupNorm(1, C) is the
// per-sample affine map the kernel applies on the fly (any cin; bf16; N % 16 == 0)
int k_linear_attention_prenorm_core(const void* qkv_raw, void* out, const float* uv, const void* gn_part, int gn_splits, float eps,
                                    const void* x, int ldx, int cin, int batch, int n_tokens, cudaStream_t st);
int k_attention(const void* qkv, void* out, int batch, int n_tokens, int dtype, cudaStream_t st);

// ---- conv (conv_simt.cu / conv_tc.cu)
// GroupNorm fused into the epilogue of the tcgen05 convolutions (conv_epilogue.cuh): the conv that PRODUCES a tensor also
// takes its GroupNorm statistics (mode 1) or stores it already normalised (mode 2) -- src/UNet.py:52-58,106,147,20.
struct ConvGn {
  int mode = 0;                  // 0 off | 1 statistics only -> scratch | 2 normalise (+SiLU) (+residual) before the store
  int groups = 1;
  int silu = 0;
  int nvar = 1, var_rows = 0;    // mode 2, halo kernel: sample n is normalised nvar times with rowvec row n + k*var_rows and
                                 // stored as image n + k*var_rows (the sampler's cond / uncond halves share the convolution)
  float eps = 1e-5f;
  const float* gamma = nullptr; const float* beta = nullptr;   // [cout]
  const float* rowvec = nullptr; int ld_rowvec = 0;            // per-image channel vector added BEFORE the statistics
  const void* res = nullptr; int ldres = 0;                    // NHWC tensor added AFTER the normalisation (Residual)
  // mode 1: float2 [batch][groups][nslots] partial sums {S, Q} of the stored values (consumer adds the slots in order)
  // mode 2: 16-byte packets [batch][groups][nvar][nslots]; the caller zeroes the area before the first launch that uses
  //         it and gives every launch since then a different non-zero tag
  void* scratch = nullptr; int64_t scratch_bytes = 0;
  unsigned tag = 0;
  int* nslots_out = nullptr;     // receives nslots
};
struct ConvArgs {
  const void* x; int ldx; int cin;        // main source, NHWC [batch,H,W,*]
  const void* x2; int ldx2; int cin2;     // optional K-concatenated 1x1 source (same spatial size)
  const void* w;                          // packed [Cout_total][taps*cin + cin2], element type = dtype
  const float* bias;                      // [cout_total] or null (indexed by output channel, see `up2`)
  const float* rowvec; int ld_rowvec;     // optional [batch][ld_rowvec] per-sample, per-channel add
  const void* res; int ldres;             // optional residual NHWC (same spatial size, cout channels)
  void* y; int ldy;                       // output NHWC (may be null with fin_out)
  const float* fin_w = nullptr;           // optional fused trailing 1x1 projection [fin_cout][cout] fp32 ...
  const float* fin_b = nullptr;           // ... + bias, written as fp32 NCHW [batch][fin_cout][H*W] (tcgen05 path only)
  float* fin_out = nullptr;
  int fin_cout = 0;
  int res_mod = 0;                        // > 0: the residual of image n is read from image n % res_mod (halo kernel only)
  // GroupNorm (+SiLU) of the MAIN SOURCE applied on the fly (halo kernel only): x is the RAW tensor and xf_ab [batch][cin] holds
  // per-(output image, channel) {a, b} from k_group_norm_coef; a transform warp pair rewrites each slab in shared memory as
  // act(a x + b) between its TMA landing and the MMAs (padding stays zero).  x_mod > 0: image n is read from image n % x_mod.
  const void* xf_ab = nullptr; int xf_silu = 0; int x_mod = 0;
  ConvGn gn;                              // fused GroupNorm epilogue (tcgen05 kernels only)
  int cout;                               // GEMM N (for up2: 4 * output channels)
  int batch, height, width;
  int ksize;                              // 1 or 3 (pad = ksize/2)
  int up2;                                // 1: ConvTranspose2d k2 s2 scatter epilogue (GEMM N = 4*Cout, quadrant-major)
  int dtype;
};
int k_conv_simt(const ConvArgs& a, cudaStream_t st);
int k_conv_tc(const ConvArgs& a, cudaStream_t st);   // bf16 only, tcgen05/TMEM/TMA
int k_conv_tc_prepare();  // resolve the driver entry point + opt in to large dynamic smem (call outside graph capture)
// 3x3 convolutions at full resolution: A slab + halo loaded once per tile, taps = descriptor shifts (conv_halo.cu)
int k_conv_halo_prepare();
bool k_conv_halo_applicable(const ConvArgs& a);
int k_conv_halo(const ConvArgs& a, cudaStream_t st);
int k_conv(const ConvArgs& a, int impl, cudaStream_t st);
// {a, b} [rows][channels] (float2) with act(a x + b) == [SiLU](GroupNorm(x + rowvec)) for the statistics in `part`
// (nslots < 0: raw {S, Q} slots of a convolution epilogue, ConvGn mode 1; > 0: pivoted partial sums of k_group_norm_stats,
// pivot read from x).  silu: the pair is pre-halved for the tanh form of SiLU used by the halo kernel's transform warps.
int k_group_norm_coef(const void* part, int nslots, const float* gamma, const float* beta, const float* rowvec, int ld_rowvec,
                      int rows, int hw, int channels, int groups, float eps, int x_mod, const void* x, int ldx, int silu,
                      void* ab_out, cudaStream_t st);

// weight packing (pack.cu)
int k_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize, const float* w2_oi11, int cin2,
                       void* w_packed, int dtype, cudaStream_t st);
// ConvTranspose2d IOHW [Cin][Cout][2][2] -> [(dy,dx,co)][ci]
int k_pack_convT_weight(const float* w_iohw, int cin, int cout, void* w_packed, int dtype, cudaStream_t st);
int k_transpose_f32(const float* w, int rows, int cols, float* wt, int ld_out, int col_off, cudaStream_t st);
int k_copy_f32(const float* src, float* dst, int64_t n, cudaStream_t st);
// initial conv weight OIHW -> [3][3][Cin][Cout]
int k_pack_initial_weight(const float* w_oihw, int cout, int cin, float* out, cudaStream_t st);
// dst = a + b (b may be null): fused conv2 + shortcut bias
int k_add2_f32(const float* a, const float* b, float* dst, int n, cudaStream_t st);

// ---- backward.cu (training step)
int k_conv_wgrad(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw, float* dbias, int batch,
                 int H, int W, int ksize, int dtype, cudaStream_t st);
int k_colsum(const void* a, int lda, float* out, int M, int C, int dtype, cudaStream_t st);
int k_pack_dgrad_weight(const float* w_oihw, int cout, int cin, int ksize, void* out, int dtype, cudaStream_t st);
int k_group_norm_backward(const void* x, int ldx, const void* dy, int lddy, const float* gamma, const float* beta,
                          const float* rowvec, int ld_rowvec, void* dx, int lddx, float* dgamma, float* dbeta,
                          float* drowvec, int ld_drowvec, int batch, int hw, int channels, int groups, float eps, int silu,
                          int dtype, cudaStream_t st);
int k_maxpool2_backward(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int batch, int H, int W, int C,
                        int dtype, cudaStream_t st);
int k_unshuffle2(const void* dy, int lddy, void* out, int batch, int H, int W, int C, int dtype, cudaStream_t st);
int k_linear_attention_backward(const void* qkv, const void* dout, void* dqkv, int batch, int N, int dtype, cudaStream_t st);
int k_attention_backward(const void* qkv, const void* dout, void* dqkv, int batch, int N, int dtype, cudaStream_t st);
int64_t k_initial_conv_wgrad_scratch_bytes(int batch, int cin, int cout, int H, int W);
int k_initial_conv_wgrad(const float* x, const void* dy, float* dw, float* dbias, int batch, int cin, int cout, int H, int W,
                         int dtype, void* scratch, cudaStream_t st);
int k_final_conv_backward(const float* dout, const void* x, int ldx, const float* w, void* dx, float* dw, float* db, int batch,
                          int cin, int cout, int hw, int dtype, cudaStream_t st);
int k_gemm_f32(const float* a, int64_t a_rs, int64_t a_cs, const float* b, int64_t b_rs, int64_t b_cs, float* c, int64_t c_rs,
               int M, int N, int K, int accumulate, cudaStream_t st);
int k_sinusoid(const int64_t* t, float* emb, int batch, int Din, cudaStream_t st);
int k_ew(const float* x, const float* other, float* y, int rows, int cols, int mode, cudaStream_t st);
int k_rows_scatter_add(const float* src, const int64_t* idx, int idx_len, float* table, int batch, int D, cudaStream_t st);
int k_rows_gather_add(float* dst, const int64_t* idx, int idx_len, const float* table, int batch, int D, cudaStream_t st);
int k_add(const void* a, const void* b, void* out, int64_t n, int dtype, cudaStream_t st);
int k_copy_channels(const void* src, int lds, void* dst, int ldd, int C, int64_t rows, int dtype, cudaStream_t st);

// ---- wgrad_tc.cu: tcgen05 weight gradient on channel-major (pixel-contiguous) operand copies
int k_nhwc_to_chw_bf16(const void* x, int ld, void* y, void* y_l, void* y_r, float* colsum, int batch, int C, int hw, int W,
                       cudaStream_t st);
bool k_conv_wgrad_tc_applicable(int cin, int cout, int H, int W, int ksize, int dtype);
bool k_conv_wgrad_tc_flat_applicable(int cin, int cout, int batch, int H, int W, int ksize, int dtype);
int k_nhwc_to_flat_taps_bf16(const void* x, int ld, void* y, float* colsum, int batch, int C, int H, int W, int taps,
                             cudaStream_t st);
int k_conv_wgrad_tc_flat(const void* xF, int cin, const void* dyF, int cout, float* dw, float* nat, int batch, int H, int W,
                         int ksize, cudaStream_t st);
int k_conv_wgrad_tc(const void* xT, int cin, const void* dyT, const void* dyT_l, const void* dyT_r, int cout, float* dw,
                    float* nat, int batch, int H, int W, int ksize, cudaStream_t st);

// ---- trainer_ops.cu: Adam step over flat fp32 buffers, uint8 image output, MSE
int k_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2, double eps,
                int step, double grad_scale, cudaStream_t st);
int k_images_to_u8(const float* x, uint8_t* out, int batch, int C, int hw, int convention, cudaStream_t st);
int k_mse(const float* a, const float* b, float* out, int64_t n, cudaStream_t st);
int k_mse_backward(const float* pred, const float* target, const float* gout, float* dpred, int64_t n, cudaStream_t st);

// ---- autoencoder.cu: the first-stage autoencoder's extra pieces (src/Autoencoder.py)
int k_group_norm_any(const void* x, int ldx, void* y, int ldy, const float* gamma, const float* beta, int batch, int hw,
                     int channels, int groups, float eps, int silu, int dtype, cudaStream_t st);
int k_upsample_nearest2x(const void* x, int ldx, void* y, int ldy, int batch, int H, int W, int C, int dtype, cudaStream_t st);
int k_pick_odd(const void* x, int ldx, void* y, int ldy, int batch, int H, int W, int C, int dtype, cudaStream_t st);
int k_attention_single_head(const void* qkv, void* out, int batch, int n_tokens, int channels, int dtype, cudaStream_t st);
int k_gaussian(const void* moments, int ld, const float* eps, float* mu, float* log_var, float* sigma, float* z, int batch,
               int zc, int hw, int dtype, cudaStream_t st);

// ---- groupnorm.cu: streaming (multi-CTA) GroupNorm backward, bf16
bool k_group_norm_backward_streams(int batch, int hw, int channels, int groups, int dtype);
int64_t k_group_norm_backward_ws_bytes(int batch, int hw, int channels, int groups);
int k_group_norm_backward_stream(const void* x, int ldx, const void* dy, int lddy, const float* gamma, const float* beta,
                                 const float* rowvec, int ld_rowvec, void* dx, int lddx, float* dgamma, float* dbeta,
                                 float* drowvec, int ld_drowvec, int batch, int hw, int channels, int groups, float eps, int silu,
                                 const void* fwd_part, void* workspace, cudaStream_t st);

// ---- attention.cu: LinearAttention backward on mma.sync
int64_t k_linear_attention_backward_ws_bytes(int batch);
bool k_linear_attention_backward_mma_applicable(int n_tokens, int dtype);
int k_linear_attention_backward_mma(const void* qkv, const void* dout, void* dqkv, int batch, int N, void* workspace,
                                    cudaStream_t st);
int k_pack_conv_weight_pair(const float* w_oihw, int cout, int cin, int ksize, void* fwd, void* dgrad, int dtype, cudaStream_t st);
int k_pack_dense2x2_weight(const float* w_oihw, int cout, int cin, const float* w2_oi11, int cin2, void* out, int dtype,
                           cudaStream_t st);
bool k_conv_wgrad_mn_applicable(int cin, int cout, int H, int W, int ksize, int dtype);
int k_conv_wgrad_mn(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw, float* dbias, float* nat,
                    int batch, int H, int W, int ksize, cudaStream_t st);
int k_pack_center_tap_weight(const float* w_oihw, int cout, int cin, void* out, int dtype, cudaStream_t st);
