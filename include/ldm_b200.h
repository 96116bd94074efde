/*
 * ldm_b200.h -- C ABI of the B200-native DDPM denoising hot path.
 *
 * The reference (JohanLundberg12/latent-diffusion-models) is pure Python/PyTorch
 * and has no FFI of its own; its "operator API" for this path is the duck-typed
 * Python protocol of src/UNet.py and src/DDPM.py.  Every entry point below
 * replaces the PyTorch library calls behind one reference function and cites it
 * (file:line relative to the reference tree).  INTEGRATION.md shows the ctypes
 * binding a reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch types.
 *   - every function returns 0 on success, <0 on error; ldm_last_error() gives the
 *     thread-local message.  Nothing throws or aborts across the boundary.
 *   - all pointers are DEVICE pointers unless the name ends in _host.
 *   - every launch goes to the caller's `stream` (a cudaStream_t passed as void*);
 *     no hidden synchronisation, no host reads of device data, no allocation after
 *     *_create / *_load_params (so the calls are CUDA-graph capturable).
 *   - reference-facing tensors are fp32 NCHW contiguous, indices are int64
 *     (what src/DDPM.py and src/UNet.py exchange).  Internally activations are
 *     NHWC in the handle's compute dtype.
 */
#ifndef LDM_B200_H_
#define LDM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDM_B200_ABI_VERSION 1

/* compute precision of a UNet handle */
enum ldm_dtype {
  LDM_F32 = 0,  /* fp32 activations, FFMA implicit-GEMM convs: parity path (<=1e-4 rel.)   */
  LDM_BF16 = 1  /* bf16 NHWC activations, tcgen05/TMEM/TMA implicit-GEMM convs (<=2e-2 rel.) */
};

/* ---- library ---------------------------------------------------------------------- */
int ldm_abi_version(void);
const char* ldm_last_error(void);
/* number of kernels this library has launched since load / last reset (all handles). */
int64_t ldm_launch_count(void);
void ldm_reset_launch_count(void);

/* ---- UNet eps-model: replaces src/UNet.py:293-389 (UNet.__init__ / UNet.forward) -- */
typedef struct ldm_unet ldm_unet;

typedef struct ldm_unet_desc {
  int32_t in_channels;        /* src/UNet.py:296 */
  int32_t out_channels;       /* src/UNet.py:297 */
  int32_t channels;           /* src/UNet.py:298 (default 64)                        */
  int32_t n_levels;           /* len(channel_multipliers), src/UNet.py:299           */
  int32_t channel_multipliers[8];
  int32_t with_time_emb;      /* src/UNet.py:300                                      */
  int32_t num_classes;        /* src/UNet.py:301; 0 = None                            */
  int32_t image_size;         /* H = W of x_noisy; must be divisible by 2^n_levels    */
  int32_t dtype;              /* enum ldm_dtype                                       */
  int32_t conv_impl;          /* 0 = default for dtype (bf16: tcgen05), 1 = force FFMA (debug/A-B) */
} ldm_unet_desc;

int ldm_unet_create(const ldm_unet_desc* desc, ldm_unet** out);
void ldm_unet_destroy(ldm_unet* h);

/* Number of state_dict tensors and their canonical order / names / element counts
 * (order == reference state_dict order). */
int ldm_unet_num_params(const ldm_unet* h);
const char* ldm_unet_param_name(const ldm_unet* h, int index);
int64_t ldm_unet_param_numel(const ldm_unet* h, int index);

/* (Re)pack all parameters from the caller's fp32 state_dict tensors (device pointers,
 * canonical order, PyTorch layouts: conv OIHW, ConvTranspose IOHW, Linear [out,in]).
 * Replaces nn.Module.load_state_dict for the packed copies (src/utils.py:36-45). */
int ldm_unet_load_params(ldm_unet* h, const float* const* params, int n_params, void* stream);

/* Workspace the caller must provide to forward for a batch of `batch` rows. */
int64_t ldm_unet_workspace_bytes(const ldm_unet* h, int batch);

/* eps = UNet(x_noisy, t, y)  -- src/UNet.py:361-389.
 *   x      [batch, Cin, S, S] fp32 NCHW          out [batch, Cout, S, S] fp32 NCHW
 *   t      [batch] int64, or NULL with t_dev_scalar: one int64 on the device broadcast to all rows
 *   y      int64 labels: y_len == batch, or y_len == 1 (broadcast, src/UNet.py:375-376), or NULL/0 = None
 *   y_rows applies labels only to rows [0, y_rows) (classifier-free-guidance batching: rows
 *          [0,B) conditional, [B,2B) unconditional); pass `batch` for the plain call.        */
int ldm_unet_forward(ldm_unet* h, const float* x, const int64_t* t, const int64_t* t_dev_scalar,
                     const int64_t* y, int y_len, int y_rows, int batch, float* out,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* Debug/parity tap: subsequent forwards also copy the named internal NHWC activation to fp32 NCHW
 * `out_nchw` (device, `out_numel` elements); name NULL or out NULL clears the tap.
 * names: "initial", "enc{i}.res", "enc{i}.attn", "bottleneck", "dec{i}", "final.res", "temb". */
int ldm_unet_set_tap(ldm_unet* h, const char* name, float* out_nchw, int64_t out_numel);

/* Measurement aid (bench.py's roofline leg; no reference counterpart): one forward with a CUDA-event pair
 * around every kernel launch on `stream`, summed per kernel family.  flops / bytes are the ALGORITHMIC work of
 * the launches (conv: 2*M*N*K; memory-bound kernels: one read + one write of each live tensor).
 * `t` is ONE device int64: the batch-constant timestep, used exactly as the sampler uses it.
 * Synchronises `stream` before returning; not capturable. */
enum ldm_kernel_family {
  LDM_FAM_CONV_TC = 0,          /* tcgen05 implicit-GEMM 3x3 convolutions (tensor bound)          */
  LDM_FAM_CONV_FFMA = 1,        /* fp32 / forced-FFMA implicit-GEMM convolutions                  */
  LDM_FAM_GROUP_NORM = 2,       /* GroupNorm (+SiLU) (+residual)                                  */
  LDM_FAM_LINEAR_ATTENTION = 3,
  LDM_FAM_ATTENTION = 4,
  LDM_FAM_OTHER = 5,            /* time-embedding MLPs, initial/final conv, max-pool              */
  LDM_FAM_CONV_TC_1X1 = 6,      /* tcgen05 1x1 convolutions and conv-transpose: K <= 768, HBM bound */
  LDM_FAM_CONV_HALO = 7,        /* the 3x3 convolutions that run in conv_halo_kernel (full resolution)     */
  LDM_FAM_COUNT = 8
};
typedef struct ldm_profile_family {
  double ms;        /* sum of launch durations (CUDA events) */
  double flops;     /* algorithmic FLOPs of those launches   */
  double bytes;     /* algorithmic bytes of those launches   */
  int64_t launches;
} ldm_profile_family;
typedef struct ldm_profile {
  ldm_profile_family family[LDM_FAM_COUNT];
} ldm_profile;
int ldm_unet_profile(ldm_unet* h, const float* x, const int64_t* t, const int64_t* y, int y_len, int y_rows,
                     int batch, float* out, void* workspace, int64_t workspace_bytes, void* stream,
                     ldm_profile* result);

/* ---- diffusion process: replaces src/DDPM.py --------------------------------------- */

/* q_sample: x_t = sqrt(abar[t_b]) x0 + sqrt(1-abar[t_b]) eps      (src/DDPM.py:46-68)
 * eps==NULL: eps is drawn in-kernel (Philox4x32-10, key seed, counter = global element) and
 * written to eps_out (src/DDPM.py:63-64, torch.randn_like).  n_per_sample = C*H*W.         */
int ldm_q_sample(const float* x0, const int64_t* t, const float* alpha_bar, int n_steps,
                 const float* eps, float* eps_out, float* xt, int batch, int64_t n_per_sample,
                 uint64_t seed, uint64_t sample_offset, void* stream);

/* p_sample with fused classifier-free guidance (src/DDPM.py:71-96 and :120-124):
 *   eps  = eps_uncond ? eps_uncond + cfg_scale*(eps_cond - eps_uncond) : eps_cond
 *   mean = alpha_t^-1/2 (x_t - (1-alpha_t)/sqrt(1-abar_t) eps)
 *   out  = t==0 ? mean : mean + sqrt(beta_t) z
 * t is read from the device: t_dev holds t_len int64 (1 = batch-constant, or `batch` per-sample values:
 * alpha/alpha_bar are gathered per sample, the t==0 branch follows t_dev[0] as in the reference, :74-85).
 * z = noise if given, else in-kernel Philox keyed by (seed, sample_offset+b, t).
 * coef is the [n_steps,4] fp32 table built by ldm_build_coef_table.  In-place (out==xt) allowed. */
int ldm_p_sample(const float* xt, const float* eps_cond, const float* eps_uncond, float cfg_scale,
                 const int64_t* t_dev, int t_len, const float* coef, int n_steps, const float* noise,
                 uint64_t seed, uint64_t sample_offset, float* out, int batch, int64_t n_per_sample,
                 void* stream);

/* coef[t] = { alpha_t^-1/2, (1-alpha_t)/sqrt(1-abar_t), sqrt(beta_t), 0 } from the reference's
 * fp32 schedule tensors (src/DDPM.py:31-43), computed on the device in fp32. */
int ldm_build_coef_table(const float* beta, const float* alpha, const float* alpha_bar, int n_steps,
                         float* coef, void* stream);

/* Standard normal fill (x_T ~ N(0,I), src/DDPM.py:108) with the sampler's Philox keying:
 * element e of sample (sample_offset+b) at stream `stream_id` -- invariant to batch sharding. */
int ldm_randn(float* out, int batch, int64_t n_per_sample, uint64_t seed, uint64_t sample_offset,
              uint64_t stream_id, void* stream);

/* ---- sampler: replaces the loop of Diffusion.sample, src/DDPM.py:98-130 ------------- */
typedef struct ldm_sampler ldm_sampler;

typedef struct ldm_sampler_desc {
  int32_t batch;          /* images per call on this GPU                                  */
  int32_t n_steps;        /* T (src/DDPM.py:23)                                           */
  float cfg_scale;        /* >0: cond+uncond batched as one 2B pass (src/DDPM.py:119-124) */
  int32_t y_len;          /* 0 (None), 1 (broadcast) or batch                             */
  int32_t use_graph;      /* 1: capture one step as a CUDA graph and replay T times       */
} ldm_sampler_desc;

int ldm_sampler_create(ldm_unet* unet, const ldm_sampler_desc* desc, ldm_sampler** out);
void ldm_sampler_destroy(ldm_sampler* s);
int64_t ldm_sampler_workspace_bytes(const ldm_sampler* s);

/* Runs the whole reverse process on `stream`:
 *   x      [batch,C,S,S] fp32 NCHW: holds x_T on entry when x_is_init!=0 (fixed-noise parity),
 *          otherwise it is filled with N(0,I) from (seed, sample_offset); holds x_0 on return.
 *   y      int64[y_len] device labels (or NULL)
 *   coef   [n_steps,4] table (ldm_build_coef_table)
 *   noise  optional [n_steps, batch, C,S,S] fp32 injected per-step noise indexed by t (parity), else NULL
 *   first_step / num_steps: run timesteps t = first_step, first_step-1, ... (num_steps of them);
 *          pass n_steps-1 / n_steps for the full trajectory.
 * No host synchronisation inside; the caller syncs the stream. */
int ldm_sampler_run(ldm_sampler* s, float* x, int x_is_init, const int64_t* y, const float* coef,
                    const float* noise, uint64_t seed, uint64_t sample_offset,
                    int first_step, int num_steps, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* ---- kernel-level entry points (unit parity tests and ncu targets) ------------------
 * NHWC tensors; `dtype` selects float or bf16 element type; ld* = pixel stride in elements. */

/* GroupNorm(+SiLU)(+residual): y = [silu](gn(x)) [+ res]  -- src/UNet.py:52-58,106,147,20 */
int ldm_group_norm(const void* x, int ldx, void* y, int ldy, const void* res, int ldres,
                   const float* gamma, const float* beta, int batch, int hw, int channels, int groups,
                   float eps, int silu, int dtype, void* workspace, void* stream);
int64_t ldm_group_norm_workspace_bytes(int batch, int groups);

/* 3x3 (pad 1) / 1x1 convolution as implicit GEMM -- src/UNet.py:54,82,119,145 (F.conv2d).
 *   w_packed: [Cout][taps*Cin (+Cin2)] in `dtype` (see ldm_pack_conv_weight); bias fp32 or NULL
 *   x2/cin2 : optional second 1x1 source K-concatenated after the taps (fused ResNetBlock shortcut, :99)
 *   rowvec  : optional fp32 [batch][ld_rowvec] added per sample and output channel (time embedding, :88-93)
 *   res     : optional residual tensor added in the epilogue (identity shortcut, :99)
 *   impl    : 0 default (bf16: tcgen05, f32: FFMA), 1 force FFMA                              */
int ldm_conv2d(const void* x, int ldx, int cin, const void* x2, int ldx2, int cin2,
               const void* w_packed, const float* bias, const float* rowvec, int ld_rowvec,
               const void* res, int ldres, void* y, int ldy, int cout,
               int batch, int height, int width, int ksize, int dtype, int impl, void* stream);
/* The same convolution with the GroupNorm that follows it fused into the epilogue (bf16, impl 0 only):
 *   gn_mode 1: y = conv(x) (+res) is stored as usual and the epilogue also leaves GroupNorm(groups, Cout) partial sums of the
 *              stored values, float2 {S, Q} at scratch[(n * groups + g) * nslots + slot]; *nslots_out slots per (image,
 *              group) -- what PreNorm consumers read (src/UNet.py:106-110)
 *   gn_mode 2: y = [silu](GroupNorm(conv(x) + gn_rowvec[n])) [+ gn_res] -- Block (src/UNet.py:52-58) with the ResNetBlock
 *              time-embedding add (:88-93), or LinearAttention.to_out's GroupNorm(1, C) + the Residual add (:147,:20).
 *              nvar = 2: image n is normalised twice (row vectors n and n + var_rows) and stored as images n and
 *              n + var_rows (cond / uncond halves sharing one convolution).  scratch: zero-filled packet area of
 *              ldm_conv2d_gn_scratch_bytes(batch); tag: non-zero, different for every call since the area was zeroed. */
int64_t ldm_conv2d_gn_scratch_bytes(int batch);
int ldm_conv2d_gn(const void* x, int ldx, int cin, const void* x2, int ldx2, int cin2, const void* w_packed,
                  const float* bias, const void* res, int ldres, void* y, int ldy, int cout, int batch, int height,
                  int width, int ksize, int gn_mode, int groups, float eps, int silu, const float* gamma,
                  const float* beta, const float* gn_rowvec, int ld_gn_rowvec, const void* gn_res, int ld_gn_res,
                  int nvar, int var_rows, void* scratch, int64_t scratch_bytes, unsigned tag, int* nslots_out,
                  void* stream);
/* OIHW fp32 -> packed [Cout][kh][kw][Cin] (+ optional OI11 second source appended along K) */
int ldm_pack_conv_weight(const float* w_oihw, int cout, int cin, int ksize,
                         const float* w2_oi11, int cin2, void* w_packed, int dtype, void* stream);

/* ConvTranspose2d(kernel 2, stride 2) -- src/UNet.py:231-233 (F.conv_transpose2d): a [B*H*W, 4*Cout] GEMM
 * whose epilogue scatters quadrant (dy,dx) to pixel (2h+dy, 2w+dx); y is [B,2H,2W,*] with pixel stride ldy
 * (so it can write straight into the first Cout channels of the decoder's concat buffer, :245).
 *   w_packed: [(dy,dx,co)][Cin] from ldm_pack_conv_transpose_weight (PyTorch layout IOHW [Cin][Cout][2][2]) */
int ldm_conv_transpose2x2(const void* x, int ldx, int cin, const void* w_packed, const float* bias, void* y,
                          int ldy, int cout, int batch, int height, int width, int dtype, int impl, void* stream);
int ldm_pack_conv_transpose_weight(const float* w_iohw, int cin, int cout, void* w_packed, int dtype,
                                   void* stream);

/* MaxPool2d(2,2) -- src/UNet.py:183,207 */
int ldm_max_pool2x2(const void* x, int ldx, void* y, int ldy, int batch, int height, int width, int channels,
                    int dtype, void* stream);

/* LinearAttention core (src/UNet.py:149-163): qkv [B,N,384] -> out [B,N,128] */
int ldm_linear_attention(const void* qkv, void* out, int batch, int n_tokens, int dtype, void* stream);
/* Attention core (src/UNet.py:122-135): qkv [B,N,384] -> out [B,N,128], N <= 256 */
int ldm_attention(const void* qkv, void* out, int batch, int n_tokens, int dtype, void* stream);

/* layout converters used by the tests: fp32 NCHW <-> NHWC(dtype) */
int ldm_nchw_to_nhwc(const float* x, void* y, int batch, int channels, int hw, int dtype, void* stream);
int ldm_nhwc_to_nchw(const void* x, int ldx, float* y, int batch, int channels, int hw, int dtype, void* stream);

/* ---- training step: replaces autograd through src/UNet.py for DiffusionModelTrainer._train_epoch
 * (src/DiffusionModelTrainer.py:36-67: eps = UNet(xt, t, y); loss = mse(noise, eps); loss.backward()).
 * Each forward kernel has a backward counterpart; ldm_b200/train.py chains them with torch.autograd.Function so
 * that loss.backward(), .grad on the 196 live parameters, Adam and wandb.watch keep working.  Activation tensors are
 * NHWC in `dtype`; parameter gradients are fp32 in the PyTorch parameter layouts.  "accumulated" outputs must be
 * zero-initialised by the caller (they are written with atomics). */

/* dW[co][ci][kh][kw] (OIHW, accumulated) = sum_{n,h,w} dy[n,h,w,co] x[n,h+kh-p,w+kw-p,ci]; dbias[co] (accumulated) or NULL */
int ldm_conv2d_wgrad(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw_oihw, float* dbias,
                     int batch, int height, int width, int ksize, int dtype, void* stream);
/* the same on tcgen05 (bf16): the contraction runs over pixels with MN-major operands read straight from the NHWC tensors;
 * `scratch` (ldm_conv2d_wgrad_scratch_bytes, 256-byte aligned; 0 = this shape is not supported, use ldm_conv2d_wgrad) holds
 * the fp32 accumulator of a 3x3 filter in GEMM-natural layout (and the operand copies of the K-major fallback kernel) */
int64_t ldm_conv2d_wgrad_scratch_bytes(int cin, int cout, int batch, int height, int width, int ksize, int dtype);
int ldm_conv2d_wgrad_tc(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw_oihw, float* dbias,
                        int batch, int height, int width, int ksize, void* scratch, void* stream);
/* filter for the data gradient: dx = ldm_conv2d(dy, w_dgrad) with the roles of Cin and Cout exchanged.
 * w_packed [Cin][kh'][kw'][Cout] = w[co][ci][k-1-kh'][k-1-kw'] */
/* ldm_pack_conv_weight and ldm_pack_conv_weight_dgrad of one filter in a single launch (the training forward needs both) */
int ldm_pack_conv_weight_pair(const float* w_oihw, int cout, int cin, int ksize, void* w_packed, void* w_packed_dgrad, int dtype,
                              void* stream);
int ldm_pack_conv_weight_dgrad(const float* w_oihw, int cout, int cin, int ksize, void* w_packed, int dtype, void* stream);
/* out[c] (accumulated) = sum_r a[r][c] */
int ldm_column_sum(const void* a, int lda, float* out, int rows, int cols, int dtype, void* stream);
/* GroupNorm of (x + rowvec[n][c]) -- the ResNetBlock time-embedding add folded into block2's norm (src/UNet.py:88-96) */
int ldm_group_norm_rowvec(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* gamma,
                          const float* beta, const float* rowvec, int ld_rowvec, int batch, int hw, int channels, int groups,
                          float eps, int silu, int dtype, void* workspace, void* stream);
/* backward of y = [silu](GroupNorm(x + rowvec)): dx (overwritten), dgamma/dbeta (accumulated), drowvec [batch][ld] (overwritten) or NULL.
 * workspace (ldm_group_norm_backward_workspace_bytes, or NULL) enables the multi-CTA streaming kernels for bf16; forward_workspace
 * (or NULL) is the workspace the matching ldm_group_norm_rowvec call used: its statistics are reused instead of recomputed. */
int64_t ldm_group_norm_backward_workspace_bytes(int batch, int hw, int channels, int groups);
int ldm_group_norm_backward(const void* x, int ldx, const void* dy, int lddy, const float* gamma, const float* beta,
                            const float* rowvec, int ld_rowvec, void* dx, int lddx, float* dgamma, float* dbeta,
                            float* drowvec, int ld_drowvec, int batch, int hw, int channels, int groups, float eps, int silu,
                            int dtype, const void* forward_workspace, void* workspace, void* stream);
int ldm_max_pool2x2_backward(const void* x, int ldx, const void* dy, int lddy, void* dx, int lddx, int batch, int height,
                             int width, int channels, int dtype, void* stream);
/* ConvTranspose2d(k2,s2) backward gather: out[n,h,w,q*C+c] = dy[n,2h+q/2,2w+q%2,c]; then dx / dW are a 1x1 dgrad / wgrad */
int ldm_pixel_unshuffle2x2(const void* dy, int lddy, void* out, int batch, int height, int width, int channels, int dtype,
                           void* stream);
/* backward of ldm_linear_attention: dqkv [B][N][384] from qkv and dout [B][N][128].  workspace
 * (ldm_linear_attention_backward_workspace_bytes, 16-byte aligned, or NULL) enables the mma.sync kernels (bf16, N % 16 == 0) */
int64_t ldm_linear_attention_backward_workspace_bytes(int batch);
int ldm_linear_attention_backward(const void* qkv, const void* dout, void* dqkv, int batch, int n_tokens, int dtype,
                                  void* workspace, void* stream);
int ldm_attention_backward(const void* qkv, const void* dout, void* dqkv, int batch, int n_tokens, int dtype, void* stream);
/* initial 3x3 conv on the fp32 NCHW input (src/UNet.py:331) and its weight gradient; w_scratch: 9*cin*cout floats */
int ldm_initial_conv(const float* x_nchw, const float* w_oihw, const float* bias, void* y, int batch, int cin, int cout,
                     int height, int width, int dtype, float* w_scratch, void* stream);
/* scratch (ldm_initial_conv_wgrad_scratch_bytes, 256-byte aligned, or NULL): bf16 -> 3x3 patches of the image as a 64-channel
 * tensor + the tcgen05 weight-gradient GEMM; fp32 -> per-CTA partial rows summed by a second kernel; NULL -> atomics */
int64_t ldm_initial_conv_wgrad_scratch_bytes(int batch, int cin, int cout, int height, int width);
int ldm_initial_conv_wgrad(const float* x_nchw, const void* dy, float* dw_oihw, float* dbias, int batch, int cin, int cout,
                           int height, int width, int dtype, void* scratch, void* stream);
/* final 1x1 conv to fp32 NCHW (src/UNet.py:347) and its backward (dw/db accumulated, dx overwritten) */
int ldm_final_conv(const void* x, int ldx, const float* w, const float* bias, float* y_nchw, int batch, int cin, int cout,
                   int hw, int dtype, void* stream);
int ldm_final_conv_backward(const float* dout_nchw, const void* x, int ldx, const float* w, void* dx, float* dw, float* db,
                            int batch, int cin, int cout, int hw, int dtype, void* stream);
/* time embedding + label embedding (src/UNet.py:23-44,263-268,373-376) and the stacked mlp_t projections (:70-73), fp32;
 * parameter gradients accumulated */
int64_t ldm_time_workspace_bytes(int batch, int D, int total);
int ldm_time_embed(const int64_t* t, const int64_t* y, int y_len, const float* w1, const float* b1, const float* w3,
                   const float* b3, const float* label_emb, float* temb, int batch, int D, void* workspace, void* stream);
int ldm_time_embed_backward(const int64_t* t, const int64_t* y, int y_len, const float* w1, const float* b1, const float* w3,
                            const float* dtemb, float* dw1, float* db1, float* dw3, float* db3, float* dlabel, int batch, int D,
                            void* workspace, void* stream);
int ldm_time_proj(const float* temb, const float* w, const float* bias, float* tproj, int batch, int D, int total,
                  void* workspace, void* stream);
int ldm_time_proj_backward(const float* temb, const float* w, const float* dtproj, float* dw, float* dbias, float* dtemb,
                           int batch, int D, int total, void* workspace, void* stream);
/* residual add and channel concat / split (src/UNet.py:20,99,245) */
int ldm_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);
int ldm_copy_channels(const void* src, int ld_src, void* dst, int ld_dst, int channels, int64_t rows, int dtype, void* stream);
/* LinearAttention with to_qkv fused (bf16, 64 input channels; src/UNet.py:145,149-163) */
int ldm_linear_attention_qkv(const void* xn, int ldx, int cin, const void* wqkv_packed, void* out, int batch, int n_tokens,
                             int dtype, void* stream);

/* Residual(PreNorm(LinearAttention)) up to the attention output, as the sampler runs it on 64-channel sites: GroupNorm(1, C)
 * statistics of the raw input x [B, N, ldx], the norm folded into to_qkv (w_qkv: fp32 [384][cin], PyTorch layout; gamma / beta:
 * the PreNorm's affine), both softmaxes and both einsums -> out [B, N, 128]  (src/UNet.py:106-110,145,149-163).
 *   impl 0: tcgen05 / TMEM / TMA kernel (N % 128 == 0);  impl 1: the mma.sync kernel (N % 16 == 0).
 *   scratch: ldm_linear_attention_prenorm_scratch_bytes(batch) bytes, 256-byte aligned. */
int64_t ldm_linear_attention_prenorm_scratch_bytes(int batch);
int ldm_linear_attention_prenorm(const void* x, int ldx, int cin, const float* w_qkv, const float* gamma, const float* beta,
                                 float eps, void* out, int batch, int n_tokens, int impl, void* scratch, int64_t scratch_bytes,
                                 void* stream);
/* The same with LinearAttention.to_out's 1x1 convolution folded in (src/UNet.py:146; tcgen05 kernel only): y [B, N, ldy] =
 * to_out.0(attention) incl. bias (64 channels), and ystats (fp32 [batch][n_tokens / 16][2]) = partial sums {S, Q} of y per
 * (32-token block, 32-channel half) for the GroupNorm(1, C) that follows (:147).  w_out: fp32 [64][128]. */
int ldm_linear_attention_prenorm_to_out(const void* x, int ldx, int cin, const float* w_qkv, const float* gamma, const float* beta,
                                        float eps, const float* w_out, const float* b_out, void* y, int ldy, float* ystats,
                                        int batch, int n_tokens, void* scratch, int64_t scratch_bytes, void* stream);

/* ---- after the hot path: optimizer step, validation loss, image output ----------------- */

/* torch.optim.Adam(params, lr) with its defaults, as src/Trainer.py:68-71 constructs it (no weight decay, no amsgrad),
 * over FLAT fp32 device buffers of n elements (16-byte aligned): one launch for the whole model instead of one foreach
 * group per tensor list.  step counts from 1.  grad is multiplied by grad_scale first (1/world_size after an
 * all-reduce(sum); 1/loss_scale under AMP, src/DiffusionModelTrainer.py:55-63).  Elements whose gradient is zero and
 * whose moments are zero are left unchanged, which is what skipping a grad-less parameter does in torch. */
int ldm_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, double lr, double beta1,
                  double beta2, double eps, int step, double grad_scale, void* stream);

/* fp32 NCHW images -> uint8 NHWC bytes on the device (one quarter of the D2H traffic of the fp32 tensor).
 * convention 0 = torchvision.utils.save_image as called by save_images (src/utils.py:121-130): x*255+0.5, clamp, truncate;
 * convention 1 = get_reverse_image_transform (src/transforms.py:22-35): ((x+1)/2)*255 then numpy astype(uint8)
 *                (truncate toward zero, low byte kept: out-of-range values wrap exactly as numpy's cast does). */
int ldm_images_to_uint8(const float* x_nchw, uint8_t* out_nhwc, int batch, int channels, int hw, int convention, void* stream);

/* F.mse_loss(a, b) (mean reduction; src/Trainer.py:59, src/DiffusionModelTrainer.py:52,105) -> one device float */
int ldm_mse(const float* a, const float* b, float* out_scalar, int64_t n, void* stream);
/* ... and its backward through the prediction (loss.backward(), src/DiffusionModelTrainer.py:58-62; src/Trainer.py:60):
 * dpred = 2 (pred - target) / n * grad_loss[0]; grad_loss is a device scalar (null = 1). */
int ldm_mse_backward(const float* pred, const float* target, const float* grad_loss, float* dpred, int64_t n, void* stream);

/* ---- first-stage autoencoder (src/Autoencoder.py) ------------------------------------------
 * Its 3x3 / 1x1 convolutions are ldm_conv2d; ldm_group_norm takes GroupNorm(32, C, eps=1e-6) + swish for every group
 * width (:9-18).  The remaining pieces: */
/* f.interpolate(x, scale_factor=2, mode="nearest") on NHWC (UpSample, :142-157) */
int ldm_upsample_nearest2x(const void* x, int ldx, void* y, int ldy, int batch, int height, int width, int channels, int dtype,
                           void* stream);
/* y[n,i,j,:] = x[n,2i+1,2j+1,:]: applied to a full-resolution pad-1 3x3 conv it yields DownSample's
 * pad (0,1,0,1) + stride-2 conv exactly (:160-180) */
int ldm_downsample_pick(const void* x, int ldx, void* y, int ldy, int batch, int height, int width, int channels, int dtype,
                        void* stream);
/* AttnBlock core (:118-130): qkv [B][N][3C] (q | k | v) -> out [B][N][C] = softmax_j(C^-1/2 q_i.k_j) v_j, one head */
int ldm_attention_single_head(const void* qkv, void* out, int batch, int n_tokens, int channels, int dtype, void* stream);
/* GaussianDistribution (:21-43): moments NHWC (pixel stride ld; channels [0,Z) = mu, [Z,2Z) = log variance) -> fp32 NCHW
 * mu, log_var, sigma = exp(log_var/2) and z = mu + sigma*eps; any output pointer may be NULL (z needs eps) */
int ldm_gaussian_distribution(const void* moments, int ld, const float* eps, float* mu, float* log_var, float* sigma, float* z,
                              int batch, int z_channels, int hw, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LDM_B200_H_ */
