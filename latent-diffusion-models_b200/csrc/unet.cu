// UNet eps-model handle: parameter inventory, weight packing and the forward pass as a fixed sequence of
// kernel launches on the caller's stream (graph-capturable: no allocation, no host sync).
// Mirrors src/UNet.py:293-389 of the reference; activations are NHWC in the handle's compute dtype.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/ldm_b200.h"
#include "kernels.h"
#include "unet_internal.h"

namespace {

constexpr int HIDDEN = 128;  // heads(4) * dim_head(32), src/UNet.py:114,140
constexpr float GN_EPS = 1e-5f;

struct ParamInfo {
  std::string name;
  int64_t numel;
};

struct ResW {
  int cin = 0, cout = 0;
  bool has_t = false, has_sc = false;
  int p_mlp_w = -1, p_mlp_b = -1, p_n1w = -1, p_n1b = -1, p_c1w = -1, p_c1b = -1, p_n2w = -1, p_n2b = -1,
      p_c2w = -1, p_c2b = -1, p_scw = -1, p_scb = -1;
  void* w1 = nullptr; float* b1 = nullptr;
  void* w2 = nullptr; float* b2 = nullptr;  // conv2 (+K-concatenated shortcut), bias = conv2.bias (+ shortcut.bias)
  // 2x2 images (the bottleneck): every output pixel sees all four input pixels, so the pad-1 3x3 conv is the dense GEMM
  // [B, 4*Cin] x [4*Cin, 4*Cout] -- 4/9 of the implicit GEMM's k-blocks (the rest multiply zero padding)
  bool dense2 = false;
  void* w1d = nullptr; float* b1d = nullptr;
  void* w2d = nullptr; float* b2d = nullptr;
  // 1x1 images (the bottleneck of a 16x16 latent): only the centre tap of a pad-1 3x3 filter ever meets data
  bool dense1 = false;
  void* w1c = nullptr; void* w2c = nullptr;
  float *g1 = nullptr, *be1 = nullptr, *g2 = nullptr, *be2 = nullptr;
  int tproj_off = -1;  // column offset into the concatenated time projection (-1: no live time embedding)
};
struct AttnW {
  int dim = 0;
  bool linear = true;
  int p_qkv = -1, p_ow = -1, p_ob = -1, p_ogw = -1, p_ogb = -1, p_nw = -1, p_nb = -1;
  void* wqkv = nullptr; void* wout = nullptr; float* bout = nullptr;
  void* wfold = nullptr; float* uv = nullptr;  // to_qkv with the PreNorm GroupNorm folded in (fused attention kernel)
  void* ucat = nullptr; float* c12 = nullptr;  // to_out folded over Wv, per head (k_fold_to_out; linattn_tc2_kernel, dim 64 only)
  float *og = nullptr, *ob = nullptr, *ng = nullptr, *nb = nullptr;
};
struct UpW {
  int cin = 0, cout = 0;
  int p_w = -1, p_b = -1;
  void* w = nullptr; float* b = nullptr;
};

struct Tap {
  std::string name;
  float* out = nullptr;
  int64_t numel = 0;
};

}  // namespace

struct ldm_unet {
  ldm_unet_desc d;
  int L = 0;                 // levels
  int D = 0;                 // time embedding width (channels*4)
  std::vector<int> dims;     // [channels, channels*m0, ...]
  std::vector<ParamInfo> params;
  // parameter indices of the non-block tensors
  int p_t1w = -1, p_t1b = -1, p_t3w = -1, p_t3b = -1, p_label = -1, p_iw = -1, p_ib = -1, p_fw = -1, p_fb = -1;
  std::vector<ResW> enc_res, dec_res;
  std::vector<AttnW> enc_attn, dec_attn;
  std::vector<UpW> ups;
  ResW bott1, bott2, final_res;
  AttnW bott_attn;
  // packed time path
  float *w1t = nullptr, *b1 = nullptr, *w3t = nullptr, *b3 = nullptr, *label = nullptr;
  float *tproj_wt = nullptr, *tproj_b = nullptr;
  float *w1 = nullptr, *w3 = nullptr, *tproj_w = nullptr;  // [out][in] copies for the warp-per-output table kernels
  int tproj_total = 0;
  float *init_w = nullptr, *init_b = nullptr, *fin_w = nullptr, *fin_b = nullptr;
  // arena
  uint8_t* arena = nullptr;
  int64_t arena_bytes = 0;
  bool loaded = false;
  Tap tap;
  // The time-embedding kernels depend only on t / y, not on the image: they run on a side stream (a parallel branch
  // of the captured graph) and join the main chain where the first ResNetBlock needs the projection.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int es() const { return dtype_size(d.dtype); }
};

namespace {

int add_param(ldm_unet* h, const std::string& name, int64_t numel) {
  h->params.push_back({name, numel});
  return (int)h->params.size() - 1;
}

void add_res(ldm_unet* h, ResW& r, const std::string& p, int cin, int cout, bool temb) {
  r.cin = cin; r.cout = cout; r.has_t = temb; r.has_sc = cin != cout;
  if (temb) {
    r.p_mlp_w = add_param(h, p + ".mlp_t.1.weight", (int64_t)cout * h->D);
    r.p_mlp_b = add_param(h, p + ".mlp_t.1.bias", cout);
  }
  r.p_n1w = add_param(h, p + ".block1.norm.weight", cin);
  r.p_n1b = add_param(h, p + ".block1.norm.bias", cin);
  r.p_c1w = add_param(h, p + ".block1.conv2d.weight", (int64_t)cout * cin * 9);
  r.p_c1b = add_param(h, p + ".block1.conv2d.bias", cout);
  r.p_n2w = add_param(h, p + ".block2.norm.weight", cout);
  r.p_n2b = add_param(h, p + ".block2.norm.bias", cout);
  r.p_c2w = add_param(h, p + ".block2.conv2d.weight", (int64_t)cout * cout * 9);
  r.p_c2b = add_param(h, p + ".block2.conv2d.bias", cout);
  if (r.has_sc) {
    r.p_scw = add_param(h, p + ".shortcut.weight", (int64_t)cout * cin);
    r.p_scb = add_param(h, p + ".shortcut.bias", cout);
  }
}
void add_attn(ldm_unet* h, AttnW& a, const std::string& p, int dim, bool linear) {
  a.dim = dim; a.linear = linear;
  a.p_qkv = add_param(h, p + ".fn.fn.to_qkv.weight", (int64_t)3 * HIDDEN * dim);
  if (linear) {
    a.p_ow = add_param(h, p + ".fn.fn.to_out.0.weight", (int64_t)dim * HIDDEN);
    a.p_ob = add_param(h, p + ".fn.fn.to_out.0.bias", dim);
    a.p_ogw = add_param(h, p + ".fn.fn.to_out.1.weight", dim);
    a.p_ogb = add_param(h, p + ".fn.fn.to_out.1.bias", dim);
  } else {
    a.p_ow = add_param(h, p + ".fn.fn.to_out.weight", (int64_t)dim * HIDDEN);
    a.p_ob = add_param(h, p + ".fn.fn.to_out.bias", dim);
  }
  a.p_nw = add_param(h, p + ".fn.norm.weight", dim);
  a.p_nb = add_param(h, p + ".fn.norm.bias", dim);
}

// bump allocator over the arena (two passes: size, then assign)
struct Bump {
  uint8_t* base;
  int64_t off = 0;
  template <typename T>
  void take(T*& p, int64_t bytes) {
    off = align_up64(off, 256);
    p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += bytes;
  }
};

void layout_res(Bump& b, ResW& r, int es) {
  b.take(r.w1, (int64_t)r.cout * 9 * r.cin * es);
  b.take(r.b1, r.cout * 4);
  b.take(r.w2, (int64_t)r.cout * (9 * r.cout + (r.has_sc ? r.cin : 0)) * es);
  b.take(r.b2, r.cout * 4);
  b.take(r.g1, r.cin * 4); b.take(r.be1, r.cin * 4);
  b.take(r.g2, r.cout * 4); b.take(r.be2, r.cout * 4);
  if (r.dense2) {
    b.take(r.w1d, (int64_t)16 * r.cout * r.cin * es); b.take(r.b1d, 4 * r.cout * 4);
    b.take(r.w2d, (int64_t)4 * r.cout * (4 * r.cout + (r.has_sc ? 4 * r.cin : 0)) * es); b.take(r.b2d, 4 * r.cout * 4);
  }
  if (r.dense1) {
    b.take(r.w1c, (int64_t)r.cout * r.cin * es);
    b.take(r.w2c, (int64_t)r.cout * r.cout * es);
  }
}
void layout_attn(Bump& b, AttnW& a, int es) {
  b.take(a.wqkv, (int64_t)3 * HIDDEN * a.dim * es);
  b.take(a.wout, (int64_t)a.dim * HIDDEN * es);
  if (a.linear) { b.take(a.wfold, (int64_t)3 * HIDDEN * a.dim * 2); b.take(a.uv, (int64_t)2 * 3 * HIDDEN * 4); }
  if (a.linear && a.dim == 64) { b.take(a.ucat, (int64_t)256 * 64 * 2); b.take(a.c12, 2 * 64 * 4); }
  b.take(a.bout, a.dim * 4);
  if (a.linear) { b.take(a.og, a.dim * 4); b.take(a.ob, a.dim * 4); }
  b.take(a.ng, a.dim * 4); b.take(a.nb, a.dim * 4);
}
int64_t layout_all(ldm_unet* h, uint8_t* base) {
  Bump b{base};
  const int es = h->es(), D = h->D, C0 = h->dims[0];
  if (h->d.with_time_emb) {
    b.take(h->w1t, (int64_t)(D / 4) * D * 4); b.take(h->b1, D * 4);
    b.take(h->w3t, (int64_t)D * D * 4); b.take(h->b3, D * 4);
    if (h->d.num_classes > 0) b.take(h->label, (int64_t)h->d.num_classes * D * 4);
    b.take(h->tproj_wt, (int64_t)D * (h->tproj_total > 0 ? h->tproj_total : 1) * 4);
    b.take(h->tproj_b, (int64_t)(h->tproj_total > 0 ? h->tproj_total : 1) * 4);
    b.take(h->w1, (int64_t)(D / 4) * D * 4);
    b.take(h->w3, (int64_t)D * D * 4);
    b.take(h->tproj_w, (int64_t)D * (h->tproj_total > 0 ? h->tproj_total : 1) * 4);
  }
  b.take(h->init_w, (int64_t)9 * h->d.in_channels * C0 * 4); b.take(h->init_b, C0 * 4);
  b.take(h->fin_w, (int64_t)h->d.out_channels * C0 * 4); b.take(h->fin_b, h->d.out_channels * 4);
  for (auto& r : h->enc_res) layout_res(b, r, es);
  for (auto& a : h->enc_attn) layout_attn(b, a, es);
  layout_res(b, h->bott1, es); layout_attn(b, h->bott_attn, es); layout_res(b, h->bott2, es);
  for (auto& r : h->dec_res) layout_res(b, r, es);
  for (auto& a : h->dec_attn) layout_attn(b, a, es);
  for (auto& u : h->ups) { b.take(u.w, (int64_t)4 * u.cout * u.cin * es); b.take(u.b, u.cout * 4); }
  layout_res(b, h->final_res, es);
  return align_up64(b.off, 256);
}

// workspace plan for one forward of `batch` rows
struct Plan {
  int64_t temb, tproj, temb_tab, tproj_tab, gnws, gnpk, gnpk_bytes, gnst, gnst_bytes, gnst2, gnst2_bytes, laflag, gnab, qkv, s[4], total;
  std::vector<int64_t> hin;  // [L+1]
  std::vector<int64_t> cat;  // [L]
};
Plan make_plan(const ldm_unet* h, int batch) {
  Plan p;
  const int es = h->es(), L = h->L, S = h->d.image_size;
  int64_t off = 0;
  auto take = [&](int64_t bytes) { off = align_up64(off, 1024); int64_t o = off; off += bytes; return o; };
  p.temb = take((int64_t)batch * (h->D > 0 ? h->D : 1) * 4);
  p.tproj = take((int64_t)batch * (h->tproj_total > 0 ? h->tproj_total : 1) * 4);
  // per-class table used when the timestep is batch-constant (sampler): num_classes + 1 rows
  p.temb_tab = take((int64_t)(h->d.num_classes + 1) * (h->D > 0 ? h->D : 1) * 4);
  p.tproj_tab = take((int64_t)(h->d.num_classes + 1) * (h->tproj_total > 0 ? h->tproj_total : 1) * 4);
  p.gnws = take(k_group_norm_ws_bytes(batch, 8));
  // fused GroupNorm epilogues (conv_epilogue.cuh): 16-byte packets (<= 256 per sample: groups x variants x tiles) and
  // statistics slots (<= 640 per sample: 8 groups x 2 variants x 36 row blocks)
  p.gnpk_bytes = (int64_t)batch * 256 * 16;
  p.gnpk = take(p.gnpk_bytes);
  // (8 groups x 2 variants x (36 or S*S/32) row blocks of the full-resolution level)
  p.gnst_bytes = (int64_t)batch * std::max<int64_t>(640, ((int64_t)S * S / 32 + 4) * 16) * 8;
  p.gnst = take(p.gnst_bytes);
  // GroupNorm(1, C) partial sums the fused attention kernel leaves for to_out's GroupNorm: S*S/16 slots per sample
  p.gnst2_bytes = (int64_t)batch * std::max<int64_t>(64, (int64_t)S * S / 16) * 8;
  p.gnst2 = take(p.gnst2_bytes);
  p.laflag = take((int64_t)batch * 4);   // per-sample "redo with the exact kernel" flags of linattn_tc2_kernel
  p.gnab = take((int64_t)batch * 1024 * 8);   // {a, b} per (image, channel) of a GroupNorm applied inside the consuming conv
  int64_t max_elems = 0;
  for (int i = 0; i < L; ++i) {
    int64_t R = S >> i;
    int64_t catc = h->dims[i + 1] + h->dims[i];  // decoder level L-1-i concat: up(dims[i]) + skip(dims[i+1])
    // (decoder j = L-1-i: rd[j+1] = dims[i], rd[j] = dims[i+1])
    int64_t c = std::max<int64_t>(std::max<int64_t>(catc, h->dims[i + 1]), HIDDEN);
    max_elems = std::max(max_elems, R * R * c);
  }
  {
    int64_t R = S >> L;
    max_elems = std::max(max_elems, R * R * std::max<int64_t>(h->dims[L], HIDDEN));
  }
  for (int i = 0; i < 4; ++i) p.s[i] = take(batch * max_elems * es);
  p.qkv = take((int64_t)batch * S * S * 3 * HIDDEN * es);
  p.hin.resize(L + 1);
  for (int i = 0; i <= L; ++i) {
    int64_t R = S >> i;
    p.hin[i] = take(batch * R * R * h->dims[i] * es);
  }
  p.cat.resize(L);
  for (int j = 0; j < L; ++j) {
    int i = L - 1 - j;  // encoder level whose skip feeds decoder level j
    int64_t R = S >> i;
    p.cat[j] = take(batch * R * R * (h->dims[i] + h->dims[i + 1]) * es);
  }
  p.total = align_up64(off, 1024);
  return p;
}

}  // namespace

// ================================================================== C ABI
extern "C" int ldm_unet_create(const ldm_unet_desc* desc, ldm_unet** out) {
  LDM_REQUIRE(desc && out, "ldm_unet_create: null argument");
  LDM_REQUIRE(desc->n_levels >= 1 && desc->n_levels <= 8, "channel_multipliers: need 1..8 levels");
  LDM_REQUIRE(desc->dtype == LDM_F32 || desc->dtype == LDM_BF16, "unknown dtype %d", desc->dtype);
  LDM_REQUIRE(desc->channels > 0 && desc->channels % 64 == 0,
              "channels=%d: the implicit-GEMM kernels need a multiple of 64", desc->channels);
  LDM_REQUIRE(desc->in_channels >= 1 && desc->in_channels <= 8 && desc->out_channels >= 1 && desc->out_channels <= 8,
              "in/out_channels must be in [1,8]");
  LDM_REQUIRE(desc->image_size > 0 && desc->image_size % (1 << desc->n_levels) == 0,
              "image_size %d is not divisible by 2^%d (the reference decoder would fail at torch.cat, src/UNet.py:245)",
              desc->image_size, desc->n_levels);
  LDM_REQUIRE(desc->num_classes == 0 || desc->with_time_emb, "num_classes without time embedding is unsupported (src/UNet.py:375)");
  ldm_unet* h = new ldm_unet();
  h->d = *desc;
  h->L = desc->n_levels;
  h->D = desc->with_time_emb ? desc->channels * 4 : 0;
  h->dims.push_back(desc->channels);
  for (int i = 0; i < h->L; ++i) {
    if (desc->channel_multipliers[i] < 1) { delete h; return ldm_set_error("channel multiplier must be >= 1"); }
    h->dims.push_back(desc->channels * desc->channel_multipliers[i]);
  }
  const bool te = desc->with_time_emb != 0;
  // ---- parameter inventory in reference state_dict order
  if (te) {
    h->p_t1w = add_param(h, "time_emb.time_mlp.1.weight", (int64_t)h->D * (h->D / 4));
    h->p_t1b = add_param(h, "time_emb.time_mlp.1.bias", h->D);
    h->p_t3w = add_param(h, "time_emb.time_mlp.3.weight", (int64_t)h->D * h->D);
    h->p_t3b = add_param(h, "time_emb.time_mlp.3.bias", h->D);
  }
  if (desc->num_classes > 0) h->p_label = add_param(h, "label_emb.weight", (int64_t)desc->num_classes * h->D);
  h->p_iw = add_param(h, "initial_conv.weight", (int64_t)desc->channels * desc->in_channels * 9);
  h->p_ib = add_param(h, "initial_conv.bias", desc->channels);
  h->enc_res.resize(h->L); h->enc_attn.resize(h->L);
  for (int i = 0; i < h->L; ++i) {
    std::string p = "encoder.downs." + std::to_string(i);
    add_res(h, h->enc_res[i], p + ".0", h->dims[i], h->dims[i + 1], te);
    add_attn(h, h->enc_attn[i], p + ".1", h->dims[i + 1], true);
  }
  const int CB = h->dims[h->L];
  add_res(h, h->bott1, "bottleneck.res1", CB, CB, te);
  add_attn(h, h->bott_attn, "bottleneck.attn", CB, false);
  add_res(h, h->bott2, "bottleneck.res2", CB, CB, te);
  {
    const int rb = h->d.image_size >> h->d.n_levels;   // bottleneck resolution
    const bool dense = rb == 2 && h->d.dtype == LDM_DT_BF16 && h->d.conv_impl == 0 && getenv("LDM_NO_DENSE2X2") == nullptr;
    h->bott1.dense2 = dense && !h->bott1.has_sc;
    h->bott2.dense2 = dense && !h->bott2.has_sc;
    // encoder level i runs at image_size >> i, decoder level j at rb << (j + 1): the 16x16 latent model has both at 2x2
    const bool any2 = h->d.dtype == LDM_DT_BF16 && h->d.conv_impl == 0 && getenv("LDM_NO_DENSE2X2") == nullptr;
    for (int i = 0; i < (int)h->enc_res.size(); ++i) h->enc_res[i].dense2 = any2 && (h->d.image_size >> i) == 2;
    for (int j = 0; j < (int)h->dec_res.size(); ++j) h->dec_res[j].dense2 = any2 && (rb << (j + 1)) == 2;
    const bool centre = rb == 1 && h->d.dtype == LDM_DT_BF16 && h->d.conv_impl == 0 && getenv("LDM_NO_DENSE2X2") == nullptr;
    h->bott1.dense1 = centre && !h->bott1.has_sc;
    h->bott2.dense1 = centre && !h->bott2.has_sc;
  }
  h->dec_res.resize(h->L); h->dec_attn.resize(h->L); h->ups.resize(h->L);
  for (int j = 0; j < h->L; ++j) {
    const int cin = h->dims[h->L - j], cout = h->dims[h->L - j - 1];  // rd[j], rd[j+1]
    std::string p = "decoder.ups." + std::to_string(j);
    add_res(h, h->dec_res[j], p + ".0", cin + cout, cout, te);
    add_attn(h, h->dec_attn[j], p + ".1", cout, true);
    h->ups[j].cin = cin; h->ups[j].cout = cout;
    h->ups[j].p_w = add_param(h, p + ".2.weight", (int64_t)cin * cout * 4);
    h->ups[j].p_b = add_param(h, p + ".2.bias", cout);
  }
  add_res(h, h->final_res, "final_conv.0", desc->channels, desc->channels, false);
  h->p_fw = add_param(h, "final_conv.1.weight", (int64_t)desc->out_channels * desc->channels);
  h->p_fb = add_param(h, "final_conv.1.bias", desc->out_channels);
  // live time projections: encoder + decoder ResNetBlocks (BottleNeck never passes t: src/UNet.py:287-288)
  int off = 0;
  if (te) {
    for (auto& r : h->enc_res) { r.tproj_off = off; off += r.cout; }
    for (auto& r : h->dec_res) { r.tproj_off = off; off += r.cout; }
  }
  h->tproj_total = off;
  if (desc->dtype == LDM_BF16 && desc->conv_impl == 0) {
    if (int rc = k_conv_tc_prepare()) { delete h; return rc; }
  }
  h->arena_bytes = layout_all(h, nullptr);
  cudaError_t e = cudaMalloc(&h->arena, h->arena_bytes);
  if (e != cudaSuccess) {
    delete h;
    return ldm_set_error("cudaMalloc(%lld bytes of packed weights) failed: %s", (long long)h->arena_bytes, cudaGetErrorString(e));
  }
  layout_all(h, h->arena);
  if (cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    ldm_unet_destroy(h);
    return ldm_set_error("ldm_unet_create: could not create the side stream / events: %s", cudaGetErrorString(cudaGetLastError()));
  }
  *out = h;
  return 0;
}

extern "C" void ldm_unet_destroy(ldm_unet* h) {
  if (!h) return;
  if (h->arena) cudaFree(h->arena);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side) cudaStreamDestroy(h->side);
  delete h;
}
extern "C" int ldm_unet_num_params(const ldm_unet* h) { return h ? (int)h->params.size() : 0; }
extern "C" const char* ldm_unet_param_name(const ldm_unet* h, int i) {
  return (h && i >= 0 && i < (int)h->params.size()) ? h->params[i].name.c_str() : nullptr;
}
extern "C" int64_t ldm_unet_param_numel(const ldm_unet* h, int i) {
  return (h && i >= 0 && i < (int)h->params.size()) ? h->params[i].numel : -1;
}

namespace {
#define RC(expr) do { int rc__ = (expr); if (rc__) return rc__; } while (0)

int pack_res(ldm_unet* h, ResW& r, const float* const* P, cudaStream_t st) {
  const int dt = h->d.dtype;
  RC(k_pack_conv_weight(P[r.p_c1w], r.cout, r.cin, 3, nullptr, 0, r.w1, dt, st));
  RC(k_copy_f32(P[r.p_c1b], r.b1, r.cout, st));
  RC(k_pack_conv_weight(P[r.p_c2w], r.cout, r.cout, 3, r.has_sc ? P[r.p_scw] : nullptr, r.has_sc ? r.cin : 0, r.w2, dt, st));
  RC(k_add2_f32(P[r.p_c2b], r.has_sc ? P[r.p_scb] : nullptr, r.b2, r.cout, st));
  RC(k_copy_f32(P[r.p_n1w], r.g1, r.cin, st)); RC(k_copy_f32(P[r.p_n1b], r.be1, r.cin, st));
  RC(k_copy_f32(P[r.p_n2w], r.g2, r.cout, st)); RC(k_copy_f32(P[r.p_n2b], r.be2, r.cout, st));
  if (r.dense2) {
    RC(k_pack_dense2x2_weight(P[r.p_c1w], r.cout, r.cin, nullptr, 0, r.w1d, dt, st));
    RC(k_pack_dense2x2_weight(P[r.p_c2w], r.cout, r.cout, r.has_sc ? P[r.p_scw] : nullptr, r.has_sc ? r.cin : 0, r.w2d, dt, st));
    for (int q = 0; q < 4; ++q) {
      RC(k_copy_f32(P[r.p_c1b], r.b1d + q * r.cout, r.cout, st));
      RC(k_copy_f32(r.b2, r.b2d + q * r.cout, r.cout, st));   // conv2.bias (+ shortcut.bias), computed above
    }
  }
  if (r.dense1) {
    RC(k_pack_center_tap_weight(P[r.p_c1w], r.cout, r.cin, r.w1c, dt, st));
    RC(k_pack_center_tap_weight(P[r.p_c2w], r.cout, r.cout, r.w2c, dt, st));
  }
  if (r.tproj_off >= 0) {
    RC(k_transpose_f32(P[r.p_mlp_w], r.cout, h->D, h->tproj_wt, h->tproj_total, r.tproj_off, st));
    RC(k_copy_f32(P[r.p_mlp_w], h->tproj_w + (int64_t)r.tproj_off * h->D, (int64_t)r.cout * h->D, st));
    RC(k_copy_f32(P[r.p_mlp_b], h->tproj_b + r.tproj_off, r.cout, st));
  }
  return 0;
}
int pack_attn(ldm_unet* h, AttnW& a, const float* const* P, cudaStream_t st) {
  const int dt = h->d.dtype;
  RC(k_pack_conv_weight(P[a.p_qkv], 3 * HIDDEN, a.dim, 1, nullptr, 0, a.wqkv, dt, st));
  RC(k_pack_conv_weight(P[a.p_ow], a.dim, HIDDEN, 1, nullptr, 0, a.wout, dt, st));
  RC(k_copy_f32(P[a.p_ob], a.bout, a.dim, st));
  if (a.linear) { RC(k_copy_f32(P[a.p_ogw], a.og, a.dim, st)); RC(k_copy_f32(P[a.p_ogb], a.ob, a.dim, st)); }
  RC(k_copy_f32(P[a.p_nw], a.ng, a.dim, st)); RC(k_copy_f32(P[a.p_nb], a.nb, a.dim, st));
  if (a.linear) RC(k_fold_prenorm_qkv(P[a.p_qkv], P[a.p_nw], P[a.p_nb], a.dim, a.wfold, a.uv, st));
  if (a.linear && a.dim == 64) RC(k_fold_to_out(P[a.p_qkv], P[a.p_nw], a.uv, P[a.p_ow], a.ucat, a.c12, st));
  return 0;
}
}  // namespace

extern "C" int ldm_unet_load_params(ldm_unet* h, const float* const* P, int n_params, void* stream) {
  LDM_REQUIRE(h && P, "ldm_unet_load_params: null argument");
  LDM_REQUIRE(n_params == (int)h->params.size(), "state_dict has %d tensors, this UNet needs %d", n_params, (int)h->params.size());
  for (int i = 0; i < n_params; ++i) LDM_REQUIRE(P[i] != nullptr, "parameter %s is null", h->params[i].name.c_str());
  cudaStream_t st = (cudaStream_t)stream;
  const int D = h->D;
  if (h->d.with_time_emb) {
    RC(k_transpose_f32(P[h->p_t1w], D, D / 4, h->w1t, D, 0, st));
    RC(k_copy_f32(P[h->p_t1b], h->b1, D, st));
    RC(k_transpose_f32(P[h->p_t3w], D, D, h->w3t, D, 0, st));
    RC(k_copy_f32(P[h->p_t1w], h->w1, (int64_t)D * (D / 4), st));
    RC(k_copy_f32(P[h->p_t3w], h->w3, (int64_t)D * D, st));
    RC(k_copy_f32(P[h->p_t3b], h->b3, D, st));
    if (h->p_label >= 0) RC(k_copy_f32(P[h->p_label], h->label, (int64_t)h->d.num_classes * D, st));
  }
  RC(k_pack_initial_weight(P[h->p_iw], h->d.channels, h->d.in_channels, h->init_w, st));
  RC(k_copy_f32(P[h->p_ib], h->init_b, h->d.channels, st));
  RC(k_copy_f32(P[h->p_fw], h->fin_w, (int64_t)h->d.out_channels * h->d.channels, st));
  RC(k_copy_f32(P[h->p_fb], h->fin_b, h->d.out_channels, st));
  for (auto& r : h->enc_res) RC(pack_res(h, r, P, st));
  for (auto& a : h->enc_attn) RC(pack_attn(h, a, P, st));
  RC(pack_res(h, h->bott1, P, st)); RC(pack_attn(h, h->bott_attn, P, st)); RC(pack_res(h, h->bott2, P, st));
  for (auto& r : h->dec_res) RC(pack_res(h, r, P, st));
  for (auto& a : h->dec_attn) RC(pack_attn(h, a, P, st));
  for (auto& u : h->ups) {
    RC(k_pack_convT_weight(P[u.p_w], u.cin, u.cout, u.w, h->d.dtype, st));
    RC(k_copy_f32(P[u.p_b], u.b, u.cout, st));
  }
  RC(pack_res(h, h->final_res, P, st));
  h->loaded = true;
  return 0;
}

extern "C" int64_t ldm_unet_workspace_bytes(const ldm_unet* h, int batch) {
  if (!h || batch < 0) return -1;
  return make_plan(h, batch > 0 ? batch : 1).total;
}

extern "C" int ldm_unet_set_tap(ldm_unet* h, const char* name, float* out_nchw, int64_t out_numel) {
  LDM_REQUIRE(h, "ldm_unet_set_tap: null handle");
  h->tap.name = name ? name : "";
  h->tap.out = out_nchw;
  h->tap.numel = out_numel;
  return 0;
}

namespace {

// Optional per-launch timing (ldm_unet_profile): one CUDA-event pair around every launch on the forward's own
// stream, aggregated per kernel family after a stream sync.  Never active inside the sampler / graph capture.
struct Prof {
  struct Rec { int fam; double flops, bytes; cudaEvent_t e0, e1; };
  std::vector<Rec> recs;
  cudaStream_t st;
  int begin(int fam, double flops, double bytes) {
    Rec r{fam, flops, bytes, nullptr, nullptr};
    LDM_CUDA(cudaEventCreate(&r.e0));
    LDM_CUDA(cudaEventCreate(&r.e1));
    LDM_CUDA(cudaEventRecord(r.e0, st));
    recs.push_back(r);
    return 0;
  }
  int end() {
    LDM_CUDA(cudaEventRecord(recs.back().e1, st));
    return 0;
  }
  ~Prof() {
    for (auto& r : recs) { if (r.e0) cudaEventDestroy(r.e0); if (r.e1) cudaEventDestroy(r.e1); }
  }
};
// LDM_SKIP_FAM (bit mask over the profile families; timing experiments only, results are wrong): the masked launches are
// not issued, so the difference of two graph-replayed runs is a family's true cost inside the step
static const int g_skip_fam = getenv("LDM_SKIP_FAM") ? atoi(getenv("LDM_SKIP_FAM")) : 0;
#define PROF(fam, flops, bytes, call)                                   \
  do {                                                                  \
    if (g_skip_fam & (1 << (fam))) break;                               \
    if (prof) RC(prof->begin((fam), (double)(flops), (double)(bytes))); \
    RC(call);                                                           \
    if (prof) RC(prof->end());                                          \
  } while (0)

struct Fwd {
  ldm_unet* h;
  cudaStream_t st;
  int B, dt, es, impl;
  uint8_t* ws;
  Plan plan;
  const float* tproj;
  Prof* prof = nullptr;
  cudaEvent_t join_event = nullptr;  // side-stream time embedding: waited for right before its first consumer
  int res_mod = 0;                   // residual row aliasing for the next conv (shared CFG prefix)
  const void* xf_ab = nullptr; int xf_silu = 0; int xf_x_mod = 0;   // on-the-fly source GroupNorm of the next conv (one-shot)
  // A ResNetBlock's second GroupNorm (+SiLU, + time row) applied INSIDE the conv that consumes it (conv_halo.cu's transform
  // warps) instead of by an apply kernel: conv1's epilogue leaves the statistics, k_group_norm_coef turns them into per-image
  // {a, b}, conv2 reads conv1's RAW output.  Only where conv2 runs in the halo kernel (3x3, Cout 64 / 128 at 32x32).
  bool xform_ok(int R, int cin, int cout) const {
    // Opt-in (LDM_CONV_XFORM=1): measured slower than the apply kernel it removes -- 94.8 -> 166 us for the 64->64 conv at
    // 32x32 x 512 rows against a 35 us apply kernel (synchronisation alone +2 us, load / FMA / store pass +39 us, tanh +33 us);
    // the two transform warps are mostly WAITING in the profile, i.e. their shared-memory traffic slows the tensor pipe's
    // operand reads rather than being the critical path itself (profiles/README.md, finding 19).
    static const bool on = getenv("LDM_CONV_XFORM") != nullptr && atoi(getenv("LDM_CONV_XFORM")) != 0;
    if (!on || dt != LDM_DT_BF16 || impl != 0 || cin > 1024) return false;
    ConvArgs a{};
    a.dtype = dt; a.ksize = 3; a.up2 = 0; a.cout = cout; a.cin = cin; a.x2 = nullptr; a.cin2 = 0; a.height = R; a.width = R;
    return k_conv_halo_applicable(a);
  }
  // GroupNorm fused into the producing convolution (conv_epilogue.cuh)
  bool fuse_gn = false;
  unsigned gn_tag = 0;               // packets: one tag per launch since the per-forward memset
  int st_slots = 0;                  // > 0: gnst holds the PreNorm statistics of the last ResNetBlock output (that many slots)
  bool want_stats = false;           // the next resblock() leaves PreNorm statistics of its output
  // images of R x R pixels the fused epilogue can handle in either tile configuration (whole samples per work unit, or
  // whole work units per sample)
  bool gn_ok(int R) const {
    const int hw = R * R;
    return fuse_gn && hw >= 16 && (hw < 128 ? 128 % hw == 0 : hw % 256 == 0);
  }
  // Normalising in the producer's epilogue (mode 2) pays where a work unit holds whole samples; where a sample spans many
  // tiles (32x32: packets, deferred second pass) the epilogue becomes the kernel's bottleneck -- measured in round 2
  // (profiles/README.md): those layers keep separate GroupNorm kernels fed by epilogue statistics.
  // Small batches (generate_images.py: one image per call) are launch-latency bound, not epilogue bound: with
  // LDM_GN_NORM_SMALL_BATCH=1 every GroupNorm rides in its producing convolution there (~20 launches fewer per timestep,
  // 724 vs 765 ms per batch-1 image).  Opt-in, because the choice then depends on the batch size and the two paths round
  // differently (fp32 accumulator vs stored bf16 value): a sample's bits would no longer be independent of its batch.
  int norm_max_hw = 64;
  bool small_batch_norm = false;
  bool gn_norm_ok(int R) const {
    return gn_ok(R) && (R * R <= norm_max_hw || (small_batch_norm && (int64_t)B * R * R <= 128 * 148));
  }
  ConvGn gn_norm(const float* gamma, const float* beta, int groups, int silu, const float* rowvec, int ldrv, const void* res,
                 int ldres, int nvar = 1, int var_rows = 0) {
    ConvGn g;
    g.mode = 2; g.groups = groups; g.silu = silu; g.eps = GN_EPS; g.gamma = gamma; g.beta = beta; g.rowvec = rowvec;
    g.ld_rowvec = ldrv; g.res = res; g.ldres = ldres; g.nvar = nvar; g.var_rows = var_rows;
    g.scratch = ws + plan.gnpk; g.scratch_bytes = plan.gnpk_bytes; g.tag = ++gn_tag;
    return g;
  }
  // GroupNorm(8, C) statistics of conv1's output + time-embedding row(s), for the apply-only kernel that follows
  ConvGn gn_stats8(const float* rowvec, int ldrv, int nvar, int var_rows, int* nslots) {
    ConvGn g;
    g.mode = 1; g.groups = 8; g.eps = GN_EPS; g.rowvec = rowvec; g.ld_rowvec = ldrv; g.nvar = nvar; g.var_rows = var_rows;
    g.scratch = ws + plan.gnst; g.scratch_bytes = plan.gnst_bytes; g.nslots_out = nslots;
    return g;
  }
  ConvGn gn_stats() {
    ConvGn g;
    g.mode = 1; g.groups = 1; g.eps = GN_EPS; g.scratch = ws + plan.gnst; g.scratch_bytes = plan.gnst_bytes; g.nslots_out = &st_slots;
    return g;
  }
  const float* fin_w = nullptr; const float* fin_b = nullptr; float* fin_out = nullptr; int fin_cout = 0;
  void* s(int i) { return ws + plan.s[i]; }
  void* gnws() { return ws + plan.gnws; }

  int tap(const char* name, const void* x, int ld, int C, int R) {
    if (h->tap.out && h->tap.name == name) {
      LDM_REQUIRE(h->tap.numel == (int64_t)B * C * R * R, "tap %s holds %lld elements, buffer has %lld", name,
                  (long long)B * C * R * R, (long long)h->tap.numel);
      return k_nhwc_to_nchw(x, ld, h->tap.out, B, C, R * R, dt, st);
    }
    return 0;
  }
  int gn(const void* x, int ldx, void* y, int ldy, const void* res, int ldres, const float* g, const float* b, int R,
         int C, int groups, int silu, const float* rowvec = nullptr, int ld_rowvec = 0, int x_mod = 0) {
    const double elems = (double)B * R * R * C;
    PROF(LDM_FAM_GROUP_NORM, 0, elems * es * (res ? 3 : 2),
         k_group_norm_mod(x, ldx, y, ldy, res, ldres, g, b, rowvec, ld_rowvec, B, R * R, C, groups, GN_EPS, silu, dt, gnws(),
                          x_mod, st));
    return 0;
  }
  int conv(const void* x, int ldx, int cin, const void* x2, int ldx2, int cin2, const void* w, const float* bias,
           const float* rowvec, int ldrv, const void* res, int ldres, void* y, int ldy, int cout, int R, int ksize,
           int up2 = 0, const ConvGn* gn = nullptr) {
    ConvArgs a;
    if (gn) a.gn = *gn;
    a.x = x; a.ldx = ldx; a.cin = cin; a.x2 = x2; a.ldx2 = ldx2; a.cin2 = cin2; a.w = w; a.bias = bias;
    a.rowvec = rowvec; a.ld_rowvec = ldrv; a.res = res; a.ldres = ldres; a.y = y; a.ldy = ldy; a.cout = cout;
    a.batch = B; a.height = R; a.width = R; a.ksize = ksize; a.up2 = up2; a.dtype = dt; a.res_mod = res_mod;
    a.xf_ab = xf_ab; a.xf_silu = xf_silu; a.x_mod = xf_x_mod;
    xf_ab = nullptr; xf_silu = 0; xf_x_mod = 0;      // one-shot: set by the caller right before the conv that consumes them
    if (y == nullptr) { a.fin_w = fin_w; a.fin_b = fin_b; a.fin_out = fin_out; a.fin_cout = fin_cout; }
    return conv_args(a);
  }
  int conv_args(const ConvArgs& a) {
    const double M = (double)a.batch * a.height * a.width, K = (double)a.ksize * a.ksize * a.cin + a.cin2;
    const double bytes = (M * (a.cin + a.cin2) + (double)a.cout * K + M * a.cout + (a.res ? M * a.cout : 0)) * es;
    const bool tc = dt == LDM_DT_BF16 && impl == 0;
    PROF(tc ? (a.ksize == 3 ? (k_conv_halo_applicable(a) ? LDM_FAM_CONV_HALO : LDM_FAM_CONV_TC) : LDM_FAM_CONV_TC_1X1) : LDM_FAM_CONV_FFMA,
         2.0 * M * a.cout * K, bytes,
         k_conv(a, impl, st));
    return 0;
  }
  // ResNetBlock  src/UNet.py:85-99.  x must not alias s0/s1/out.
  // x_rows > 0: x holds only x_rows distinct images and output row n belongs to image n % x_rows (the sampler's cond and
  // uncond halves are identical up to the first time-embedding add): norm1 + conv1 run once per distinct image.
  // x_normed: s(0) already holds SiLU(GroupNorm(8)(x)) (written by the fused max-pool + GroupNorm kernel of the level above)
  int resblock(const ResW& r, const void* x, int ldx, void* out, int ldo, int R, bool use_t, int x_rows = 0, bool x_normed = false) {
    const bool stats = want_stats && gn_ok(R);   // leave GroupNorm(1, C) statistics of the block output for the PreNorm that follows
    want_stats = false;
    st_slots = 0;
    if (x_rows > 0) {
      const int full = B;
      const float* rvm = (use_t && r.tproj_off >= 0) ? tproj + r.tproj_off : nullptr;
      B = x_rows;
      RC(gn(x, ldx, s(0), r.cin, nullptr, 0, r.g1, r.be1, R, r.cin, 8, 1));
      const bool fuse2 = gn_norm_ok(R) && rvm != nullptr;
      if (fuse2) {
        // block2's GroupNorm + SiLU in conv1's epilogue, once per time-embedding variant (cond / uncond halves)
        if (join_event) { LDM_CUDA(cudaStreamWaitEvent(st, join_event, 0)); join_event = nullptr; }
        ConvGn g2 = gn_norm(r.g2, r.be2, 8, 1, rvm, h->tproj_total, nullptr, 0, full / x_rows, x_rows);
        RC(conv(s(0), r.cin, r.cin, nullptr, 0, 0, r.w1, r.b1, nullptr, 0, nullptr, 0, s(1), r.cout, r.cout, R, 3, 0, &g2));
        B = full;
      } else if (gn_ok(R) && rvm != nullptr && k_group_norm_streams(R * R, r.cout, dt)) {
        // conv1's epilogue takes block2's GroupNorm statistics (per time-embedding variant); one apply kernel follows
        if (join_event) { LDM_CUDA(cudaStreamWaitEvent(st, join_event, 0)); join_event = nullptr; }
        int ns = 0;
        ConvGn g1 = gn_stats8(rvm, h->tproj_total, full / x_rows, x_rows, &ns);
        RC(conv(s(0), r.cin, r.cin, nullptr, 0, 0, r.w1, r.b1, nullptr, 0, nullptr, 0, s(1), r.cout, r.cout, R, 3, 0, &g1));
        B = full;
        if (xform_ok(R, r.cout, r.cout)) {
          // block2's GroupNorm + SiLU ride in conv2's operand path: every output image n reads raw image n % x_rows
          PROF(LDM_FAM_GROUP_NORM, 0, (double)B * r.cout * 8,
               k_group_norm_coef(ws + plan.gnst, -ns, r.g2, r.be2, rvm, h->tproj_total, B, R * R, r.cout, 8, GN_EPS, x_rows, nullptr, 0, 1,
                                 ws + plan.gnab, st));
          res_mod = x_rows;
          xf_ab = ws + plan.gnab; xf_silu = 1; xf_x_mod = x_rows;
          ConvGn gsx = gn_stats();
          int rcx = conv(s(1), r.cout, r.cout, nullptr, 0, 0, r.w2, r.b2, nullptr, 0, x, ldx, out, ldo, r.cout, R, 3, 0, stats ? &gsx : nullptr);
          res_mod = 0;
          return rcx;
        }
        PROF(LDM_FAM_GROUP_NORM, 0, (double)B * R * R * r.cout * es * 2,
             k_group_norm_apply_raw(s(1), r.cout, s(0), r.cout, nullptr, 0, r.g2, r.be2, rvm, h->tproj_total, B, R * R, r.cout, 8, GN_EPS, 1,
                                    ws + plan.gnst, ns, x_rows, st));
      } else {
        RC(conv(s(0), r.cin, r.cin, nullptr, 0, 0, r.w1, r.b1, nullptr, 0, nullptr, 0, s(1), r.cout, r.cout, R, 3));
        B = full;
        if (rvm && join_event) {
          LDM_CUDA(cudaStreamWaitEvent(st, join_event, 0));
          join_event = nullptr;
        }
        RC(gn(s(1), r.cout, s(0), r.cout, nullptr, 0, r.g2, r.be2, R, r.cout, 8, 1, rvm, h->tproj_total, x_rows));
      }
      res_mod = x_rows;   // identity shortcut read from the shared images
      ConvGn gs = gn_stats();
      int rc = conv(fuse2 ? s(1) : s(0), r.cout, r.cout, nullptr, 0, 0, r.w2, r.b2, nullptr, 0, x, ldx, out, ldo, r.cout, R, 3, 0,
                    stats ? &gs : nullptr);
      res_mod = 0;
      return rc;
    }
    if (r.dense2 && R == 2 && ldx == r.cin && ldo == r.cout && impl == 0 && dt == LDM_DT_BF16) {
      // [B][4][C] NHWC == [B][4C]: both convs as 1x1 GEMMs over "images" of one pixel with 4C channels
      if (!x_normed) RC(gn(x, ldx, s(0), r.cin, nullptr, 0, r.g1, r.be1, R, r.cin, 8, 1));
      RC(conv(s(0), 4 * r.cin, 4 * r.cin, nullptr, 0, 0, r.w1d, r.b1d, nullptr, 0, nullptr, 0, s(1), 4 * r.cout, 4 * r.cout, 1, 1));
      const float* rvd = (use_t && r.tproj_off >= 0) ? tproj + r.tproj_off : nullptr;
      if (rvd && join_event) {
        LDM_CUDA(cudaStreamWaitEvent(st, join_event, 0));
        join_event = nullptr;
      }
      RC(gn(s(1), r.cout, s(0), r.cout, nullptr, 0, r.g2, r.be2, R, r.cout, 8, 1, rvd, h->tproj_total));
      if (r.has_sc)   // the 1x1 shortcut as a block-diagonal K-concatenated source
        return conv(s(0), 4 * r.cout, 4 * r.cout, x, 4 * ldx, 4 * r.cin, r.w2d, r.b2d, nullptr, 0, nullptr, 0, out, 4 * ldo,
                    4 * r.cout, 1, 1);
      return conv(s(0), 4 * r.cout, 4 * r.cout, nullptr, 0, 0, r.w2d, r.b2d, nullptr, 0, x, 4 * ldx, out, 4 * ldo, 4 * r.cout, 1, 1);
    }
    if (r.dense1 && R == 1 && !(use_t && r.tproj_off >= 0) && impl == 0 && dt == LDM_DT_BF16) {
      if (!x_normed) RC(gn(x, ldx, s(0), r.cin, nullptr, 0, r.g1, r.be1, R, r.cin, 8, 1));
      RC(conv(s(0), r.cin, r.cin, nullptr, 0, 0, r.w1c, r.b1, nullptr, 0, nullptr, 0, s(1), r.cout, r.cout, 1, 1));
      RC(gn(s(1), r.cout, s(0), r.cout, nullptr, 0, r.g2, r.be2, R, r.cout, 8, 1));
      return conv(s(0), r.cout, r.cout, nullptr, 0, 0, r.w2c, r.b2, nullptr, 0, x, ldx, out, ldo, r.cout, 1, 1);
    }
    if (!x_normed) RC(gn(x, ldx, s(0), r.cin, nullptr, 0, r.g1, r.be1, R, r.cin, 8, 1));
    // the time-embedding projection (h = h + mlp_t(t), :88-93) is a per-sample channel vector: it enters block2's
    // GroupNorm statistics and shift, not the convolution
    const float* rv = (use_t && r.tproj_off >= 0) ? tproj + r.tproj_off : nullptr;
    if (rv && join_event) {
      LDM_CUDA(cudaStreamWaitEvent(st, join_event, 0));
      join_event = nullptr;
    }
    const void* h2 = s(0);
    bool xf_pending = false;
    if (gn_norm_ok(R)) {
      // block2's GroupNorm + SiLU in conv1's epilogue: conv1's raw output never exists in memory
      ConvGn g2 = gn_norm(r.g2, r.be2, 8, 1, rv, h->tproj_total, nullptr, 0);
      RC(conv(s(0), r.cin, r.cin, nullptr, 0, 0, r.w1, r.b1, nullptr, 0, nullptr, 0, s(1), r.cout, r.cout, R, 3, 0, &g2));
      h2 = s(1);
    } else if (gn_ok(R) && k_group_norm_streams(R * R, r.cout, dt)) {
      // conv1's epilogue takes block2's GroupNorm statistics (of h + time-embedding row); one apply kernel follows
      int ns = 0;
      ConvGn g1 = gn_stats8(rv, h->tproj_total, 1, 0, &ns);
      RC(conv(s(0), r.cin, r.cin, nullptr, 0, 0, r.w1, r.b1, nullptr, 0, nullptr, 0, s(1), r.cout, r.cout, R, 3, 0, &g1));
      if (xform_ok(R, r.cout, r.cout)) {
        // block2's GroupNorm + SiLU ride in conv2's operand path (conv2 reads conv1's raw output)
        PROF(LDM_FAM_GROUP_NORM, 0, (double)B * r.cout * 8,
             k_group_norm_coef(ws + plan.gnst, -ns, r.g2, r.be2, rv, h->tproj_total, B, R * R, r.cout, 8, GN_EPS, 0, nullptr, 0, 1,
                               ws + plan.gnab, st));
        h2 = s(1);
        xf_pending = true;
      } else
      PROF(LDM_FAM_GROUP_NORM, 0, (double)B * R * R * r.cout * es * 2,
           k_group_norm_apply_raw(s(1), r.cout, s(0), r.cout, nullptr, 0, r.g2, r.be2, rv, h->tproj_total, B, R * R, r.cout, 8, GN_EPS, 1,
                                  ws + plan.gnst, ns, 0, st));
    } else {
      RC(conv(s(0), r.cin, r.cin, nullptr, 0, 0, r.w1, r.b1, nullptr, 0, nullptr, 0, s(1), r.cout, r.cout, R, 3));
      RC(gn(s(1), r.cout, s(0), r.cout, nullptr, 0, r.g2, r.be2, R, r.cout, 8, 1, rv, h->tproj_total));
    }
    ConvGn gs = gn_stats();
    const ConvGn* gsp = (stats && out != nullptr) ? &gs : nullptr;
    if (xf_pending) { xf_ab = ws + plan.gnab; xf_silu = 1; xf_x_mod = 0; }
    if (r.has_sc)  // 1x1 shortcut conv K-concatenated into the second 3x3 GEMM
      RC(conv(h2, r.cout, r.cout, x, ldx, r.cin, r.w2, r.b2, nullptr, 0, nullptr, 0, out, ldo, r.cout, R, 3, 0, gsp));
    else           // identity shortcut added in the epilogue
      RC(conv(h2, r.cout, r.cout, nullptr, 0, 0, r.w2, r.b2, nullptr, 0, x, ldx, out, ldo, r.cout, R, 3, 0, gsp));
    return 0;
  }
  // Residual(PreNorm(dim, LinearAttention | Attention))  src/UNet.py:14-20,102-164.  d must not alias s0/s1/out.
  int attn_block(const AttnW& a, const void* d, int ldd, void* out, int ldo, int R) {
    void* qkv = ws + plan.qkv;
    const int slots = st_slots;   // > 0: the producing ResNetBlock left GroupNorm(1, C) statistics of d in gnst
    st_slots = 0;
    // x + GroupNorm(1,C)(to_out(att)): the GroupNorm and the residual add ride in the to_out convolution's epilogue
    auto to_out = [&](const void* att) -> int {
      if (gn_norm_ok(R)) {
        ConvGn g = gn_norm(a.og, a.ob, 1, 0, nullptr, 0, d, ldd);
        return conv(att, HIDDEN, HIDDEN, nullptr, 0, 0, a.wout, a.bout, nullptr, 0, nullptr, 0, out, ldo, a.dim, R, 1, 0, &g);
      }
      RC(conv(att, HIDDEN, HIDDEN, nullptr, 0, 0, a.wout, a.bout, nullptr, 0, nullptr, 0, s(1), a.dim, a.dim, R, 1));
      return gn(s(1), a.dim, out, ldo, d, ldd, a.og, a.ob, R, a.dim, 1, 0);
    };
    if (a.linear && impl == 0 && k_linear_attention_qkv_applicable(a.dim, R * R, dt)) {
      // PreNorm statistics, then to_qkv + both softmaxes + both einsums in one kernel that reads the RAW block input:
      // neither the normalised tensor nor the 384-channel qkv tensor ever exists
      const double N = (double)R * R;
      int splits = 1;
      if (slots > 0) splits = -slots;   // statistics left by the producing convolution's epilogue
      else PROF(LDM_FAM_GROUP_NORM, 0, (double)B * N * a.dim * es, k_group_norm_stats(d, ldd, B, R * R, a.dim, 1, gnws(), &splits, st));
      const void* gpart = slots > 0 ? (const void*)(ws + plan.gnst) : gnws();
      static const bool fold_out = getenv("LDM_LINATTN_NO_TO_OUT") == nullptr;
      if (k_linear_attention_tc_applicable(a.dim, R * R, dt) && fold_out && k_group_norm_streams(R * R, a.dim, dt)) {
        // tcgen05 kernel with to_out's 1x1 convolution folded into the per-sample context matrix: it emits to_out's output
        // (64 channels instead of the 128-channel attention tensor + a conv launch) and the partial sums of to_out's
        // GroupNorm(1, C); one apply kernel adds the residual
        int ns = 0;
        LinAttnOut fo;
        fo.wout = a.wout; fo.bout = a.bout; fo.y = s(1); fo.ldy = a.dim; fo.ystats = ws + plan.gnst2; fo.ystats_bytes = plan.gnst2_bytes;
        fo.nslots_out = &ns;
        fo.ucat = a.ucat; fo.c12 = a.c12; fo.flags = reinterpret_cast<int*>(ws + plan.laflag);
        // to_out's GroupNorm(1, C) + the Residual add can ride along too (the CTA owns the whole sample), but the serial
        // tail of that pass costs more than the stand-alone apply kernel it replaces (measured: 98.3 vs 101.0 img/s): opt-in
        static const bool fold_gn = getenv("LDM_LINATTN_GN") != nullptr && atoi(getenv("LDM_LINATTN_GN")) != 0;
        if (fold_gn && R * R / 16 <= 256) {
          fo.og = a.og; fo.ob = a.ob; fo.o = out; fo.ldo = ldo; fo.o_eps = GN_EPS;
        }
        PROF(LDM_FAM_LINEAR_ATTENTION, (double)B * (2.0 * N * 3 * HIDDEN * a.dim + 4 * 2 * 2.0 * N * 32 * 32 + 2.0 * N * HIDDEN * a.dim),
             (double)B * N * (a.dim + a.dim) * es,
             k_linear_attention_tc(d, ldd, a.wfold, a.uv, gpart, splits, GN_EPS, nullptr, B, R * R, st, &fo));
        if (!fo.o)
          PROF(LDM_FAM_GROUP_NORM, 0, (double)B * N * a.dim * es * 3,
               k_group_norm_apply_raw(s(1), a.dim, out, ldo, d, ldd, a.og, a.ob, nullptr, 0, B, R * R, a.dim, 1, GN_EPS, 0, ws + plan.gnst2, ns, 0, st));
        return 0;
      }
      if (k_linear_attention_tc_applicable(a.dim, R * R, dt))   // tcgen05 / TMEM kernel (linattn_tc.cu)
        PROF(LDM_FAM_LINEAR_ATTENTION, (double)B * (2.0 * N * 3 * HIDDEN * a.dim + 4 * 2 * 2.0 * N * 32 * 32),
             (double)B * N * (a.dim + HIDDEN) * es,
             k_linear_attention_tc(d, ldd, a.wfold, a.uv, gpart, splits, GN_EPS, qkv, B, R * R, st));
      else
        PROF(LDM_FAM_LINEAR_ATTENTION, (double)B * (2.0 * N * 3 * HIDDEN * a.dim + 4 * 2 * 2.0 * N * 32 * 32),
             (double)B * N * (a.dim + HIDDEN) * es,
             k_linear_attention_qkv_prenorm(d, ldd, a.dim, a.wfold, a.uv, gpart, splits, GN_EPS, qkv, B, R * R, dt, st));
      return to_out(qkv);
    }
    static const bool affine_off = getenv("LDM_LINATTN_NO_AFFINE") != nullptr;
    if (a.linear && impl == 0 && dt == LDM_DT_BF16 && (R * R) % 16 == 0 && !affine_off) {
      // wider LinearAttention sites: to_qkv runs on the RAW block input with the gamma-folded weights and the attention core
      // applies the PreNorm GroupNorm(1, C) as the per-sample affine map it is -- the normalised tensor is never written
      int splits = 1;
      if (slots > 0) splits = -slots;   // statistics left by the producing convolution's epilogue
      else PROF(LDM_FAM_GROUP_NORM, 0, (double)B * R * R * a.dim * es, k_group_norm_stats(d, ldd, B, R * R, a.dim, 1, gnws(), &splits, st));
      const void* gpart = slots > 0 ? (const void*)(ws + plan.gnst) : gnws();
      RC(conv(d, ldd, a.dim, nullptr, 0, 0, a.wfold, nullptr, nullptr, 0, nullptr, 0, qkv, 3 * HIDDEN, 3 * HIDDEN, R, 1));
      PROF(LDM_FAM_LINEAR_ATTENTION, (double)B * 4 * 2 * 2.0 * R * R * 32 * 32, (double)B * R * R * (3 + 1) * HIDDEN * es,
           k_linear_attention_prenorm_core(qkv, s(0), a.uv, gpart, splits, GN_EPS, d, ldd, a.dim, B, R * R, st));
      return to_out(s(0));
    }
    if (slots > 0 && dt == LDM_DT_BF16 && k_group_norm_streams(R * R, a.dim, dt)) {
      // PreNorm apply only: the statistics came out of the producing convolution's epilogue
      PROF(LDM_FAM_GROUP_NORM, 0, (double)B * R * R * a.dim * es * 2,
           k_group_norm_apply_raw(d, ldd, s(0), a.dim, nullptr, 0, a.ng, a.nb, nullptr, 0, B, R * R, a.dim, 1, GN_EPS, 0, ws + plan.gnst, slots, 0, st));
    } else {
      RC(gn(d, ldd, s(0), a.dim, nullptr, 0, a.ng, a.nb, R, a.dim, 1, 0));
    }
    RC(conv(s(0), a.dim, a.dim, nullptr, 0, 0, a.wqkv, nullptr, nullptr, 0, nullptr, 0, qkv, 3 * HIDDEN, 3 * HIDDEN, R, 1));
    if (a.linear) {
      // 2 GEMMs of 32x32xN per head (ctx = k v^T, out = ctx^T q): 2 * 2*N*32*32 * 4 heads
      PROF(LDM_FAM_LINEAR_ATTENTION, (double)B * 4 * 2 * 2.0 * R * R * 32 * 32, (double)B * R * R * (3 + 1) * HIDDEN * es,
           k_linear_attention(qkv, s(0), B, R * R, dt, st));
      return to_out(s(0));
    } else {
      PROF(LDM_FAM_ATTENTION, (double)B * 4 * 2 * 2.0 * R * R * R * R * 32, (double)B * R * R * (3 + 1) * HIDDEN * es,
           k_attention(qkv, s(0), B, R * R, dt, st));
      RC(conv(s(0), HIDDEN, HIDDEN, nullptr, 0, 0, a.wout, a.bout, nullptr, 0, d, ldd, out, ldo, a.dim, R, 1));
    }
    return 0;
  }
};

}  // namespace

extern "C" int ldm_unet_forward(ldm_unet* h, const float* x, const int64_t* t, const int64_t* t_dev_scalar,
                                const int64_t* y, int y_len, int y_rows, int batch, float* out, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  return ldm_unet_forward_ex(h, x, batch, t, t_dev_scalar, y, y_len, y_rows, batch, out, workspace, workspace_bytes,
                             stream);
}

static int forward_impl(ldm_unet* h, const float* x, int x_batch, const int64_t* t, const int64_t* t_dev_scalar,
                        const int64_t* y, int y_len, int y_rows, int batch, float* out, void* workspace,
                        int64_t workspace_bytes, void* stream, Prof* prof_sink);

int ldm_unet_forward_ex(ldm_unet* h, const float* x, int x_batch, const int64_t* t, const int64_t* t_dev_scalar,
                        const int64_t* y, int y_len, int y_rows, int batch, float* out, void* workspace,
                        int64_t workspace_bytes, void* stream) {
  return forward_impl(h, x, x_batch, t, t_dev_scalar, y, y_len, y_rows, batch, out, workspace, workspace_bytes, stream,
                      nullptr);
}

// Forward with per-launch CUDA-event timing (bench.py's roofline leg): synchronises `stream` at the end.
extern "C" int ldm_unet_profile(ldm_unet* h, const float* x, const int64_t* t, const int64_t* y, int y_len, int y_rows,
                                int batch, float* out, void* workspace, int64_t workspace_bytes, void* stream,
                                ldm_profile* result) {
  LDM_REQUIRE(result, "ldm_unet_profile: null result");
  memset(result, 0, sizeof(*result));
  Prof prof;
  prof.st = (cudaStream_t)stream;
  // `t` is used as the sampler uses it: one device-resident, batch-constant timestep (t[0])
  RC(forward_impl(h, x, batch, nullptr, t, y, y_len, y_rows, batch, out, workspace, workspace_bytes, stream, &prof));
  LDM_CUDA(cudaStreamSynchronize(prof.st));
  const bool dump = getenv("LDM_PROFILE_DUMP") != nullptr;
  int idx = 0;
  for (auto& r : prof.recs) {
    float ms = 0.f;
    LDM_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
    if (dump)
      fprintf(stderr, "ldm_profile %3d fam %d  %8.1f us  %8.2f GFLOP %7.1f TFLOP/s  %8.2f MB %7.1f GB/s\n", idx++, r.fam,
              ms * 1e3, r.flops / 1e9, r.flops / (ms * 1e-3) / 1e12, r.bytes / 1e6, r.bytes / (ms * 1e-3) / 1e9);
    ldm_profile_family& f = result->family[r.fam];
    f.ms += ms; f.flops += r.flops; f.bytes += r.bytes; f.launches += 1;
  }
  return 0;
}

static int forward_impl(ldm_unet* h, const float* x, int x_batch, const int64_t* t, const int64_t* t_dev_scalar,
                        const int64_t* y, int y_len, int y_rows, int batch, float* out, void* workspace,
                        int64_t workspace_bytes, void* stream, Prof* prof_sink) {
  LDM_REQUIRE(h && x && out, "ldm_unet_forward: null argument");
  LDM_REQUIRE(x_batch > 0 && batch % x_batch == 0, "x_batch %d must divide batch %d", x_batch, batch);
  LDM_REQUIRE(h->loaded, "ldm_unet_forward: parameters were never loaded (ldm_unet_load_params)");
  LDM_REQUIRE(batch >= 0, "negative batch");
  if (batch == 0) return 0;
  LDM_REQUIRE(y_len == 0 || y_len == 1 || y_len == batch || (y_rows > 0 && y_len == y_rows),
              "labels: y has %d entries for a batch of %d (the reference broadcasts only length 1, src/UNet.py:375-376)",
              y_len, batch);
  LDM_REQUIRE(!(y && y_len > 0) || h->p_label >= 0, "labels given but the UNet has num_classes=None");
  LDM_REQUIRE(!h->d.with_time_emb || t || t_dev_scalar, "t is required");
  Fwd f;
  f.h = h; f.st = (cudaStream_t)stream; f.B = batch; f.dt = h->d.dtype; f.es = h->es(); f.impl = h->d.conv_impl;
  f.prof = prof_sink;
  f.plan = make_plan(h, batch);
  LDM_REQUIRE(workspace && workspace_bytes >= f.plan.total, "workspace too small: %lld < %lld bytes",
              (long long)workspace_bytes, (long long)f.plan.total);
  LDM_REQUIRE(((uintptr_t)workspace & 1023) == 0, "workspace must be 1024-byte aligned");
  f.ws = (uint8_t*)workspace;
  f.fuse_gn = f.dt == LDM_DT_BF16 && f.impl == 0 && getenv("LDM_NO_GN_FUSE") == nullptr;
  if (const char* e = getenv("LDM_GN_NORM_MAXHW")) f.norm_max_hw = atoi(e);
  f.small_batch_norm = getenv("LDM_GN_NORM_SMALL_BATCH") != nullptr && atoi(getenv("LDM_GN_NORM_SMALL_BATCH")) != 0;
  if (f.fuse_gn)   // packet area of the fused GroupNorm epilogues: zero = no packet; every launch below gets its own tag
    LDM_CUDA(cudaMemsetAsync(f.ws + f.plan.gnpk, 0, f.plan.gnpk_bytes, f.st));
  const int L = h->L, S = h->d.image_size;
  float* temb = (float*)(f.ws + f.plan.temb);
  float* tproj = (float*)(f.ws + f.plan.tproj);
  f.tproj = tproj;
  bool time_forked = false;
  if (h->d.with_time_emb) {
    Prof* prof = f.prof;
    const int64_t* yy = (y && y_len > 0) ? y : nullptr;
    const int yr = y_rows > 0 ? y_rows : batch;
    const bool table = t == nullptr && h->tproj_total > 0 && !(h->tap.out && h->tap.name == "temb") &&
                       h->D % 128 == 0 && h->D <= 256 && (int64_t)(h->d.num_classes + 1) * h->D * 4 <= 48 * 1024;
    if (table) {
      // Batch-constant timestep (the sampler): at most num_classes + 1 distinct embedding rows exist, so the two
      // MLPs run on that table and a gather expands the projection to the batch (src/UNet.py:373-376,90-93).
      const int ncls = yy ? h->d.num_classes : 0, R = ncls + 1;
      float* temb_tab = (float*)(f.ws + f.plan.temb_tab);
      float* tproj_tab = (float*)(f.ws + f.plan.tproj_tab);
      cudaStream_t ts = f.st;
      if (!prof) {  // fork: these three launches overlap the initial conv and the first GroupNorm + conv
        LDM_CUDA(cudaEventRecord(h->ev_fork, f.st));
        LDM_CUDA(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
        ts = h->side;
        time_forked = true;
      }
      PROF(LDM_FAM_OTHER, 2.0 * (h->D / 4 + h->D) * h->D + 2.0 * R * h->D * h->tproj_total,
           ((double)(h->D / 4 + h->D) * h->D + (double)h->D * h->tproj_total + (double)R * (h->D + h->tproj_total)) * 4,
           k_time_table(t_dev_scalar, h->w1, h->b1, h->w3, h->b3, h->label, h->tproj_w, h->tproj_b, temb_tab, tproj_tab, R,
                        ncls, h->D, h->tproj_total, ts));
      PROF(LDM_FAM_OTHER, 0, (double)batch * h->tproj_total * 4,
           k_tproj_gather(tproj_tab, yy, y_len, yr, ncls, tproj, batch, h->tproj_total, ts));
      if (time_forked) LDM_CUDA(cudaEventRecord(h->ev_join, h->side));
    } else {
      PROF(LDM_FAM_OTHER, 2.0 * batch * (h->D / 4 + h->D) * h->D, ((double)(h->D / 4 + h->D) * h->D + 2.0 * batch * h->D) * 4,
           k_time_embed(t, t_dev_scalar, yy, y_len, yr, h->w1t, h->b1, h->w3t, h->b3, h->label, temb, batch, h->D, 0, f.st));
      if (h->tap.out && h->tap.name == "temb") {
        LDM_REQUIRE(h->tap.numel == (int64_t)batch * h->D, "tap temb size mismatch");
        RC(k_copy_f32(temb, h->tap.out, (int64_t)batch * h->D, f.st));
      }
      if (h->tproj_total > 0)
        PROF(LDM_FAM_OTHER, 2.0 * batch * h->D * h->tproj_total, ((double)h->D * h->tproj_total + (double)batch * (h->D + h->tproj_total)) * 4,
             k_time_proj(temb, h->tproj_wt, h->tproj_b, tproj, batch, h->D, h->tproj_total, f.st));
    }
  }
  Prof* prof = f.prof;
  // initial conv: fp32 NCHW -> NHWC
  void* hin0 = f.ws + f.plan.hin[0];
  // Shared CFG prefix: with aliased input rows (x_batch < batch) the initial conv and the first ResNetBlock's norm1 + conv1
  // see identical data in every alias group -> computed for x_batch images only.
  bool share_prefix = false;
  if (x_batch < batch && f.dt == LDM_DT_BF16 && f.impl == 0 && !h->enc_res[0].has_sc && h->enc_res[0].tproj_off >= 0 &&
      !(h->tap.out != nullptr) && getenv("LDM_NO_SHARED_PREFIX") == nullptr) {
    ConvArgs probe;
    probe.x = hin0; probe.ldx = h->dims[0]; probe.cin = h->dims[0]; probe.x2 = nullptr; probe.ldx2 = 0; probe.cin2 = 0;
    probe.cout = h->dims[1]; probe.batch = batch; probe.height = S; probe.width = S; probe.ksize = 3; probe.up2 = 0; probe.dtype = f.dt;
    share_prefix = k_conv_halo_applicable(probe) && k_group_norm_streams(S * S, h->dims[1], f.dt);
  }
  const int init_rows = share_prefix ? x_batch : batch;
  PROF(LDM_FAM_OTHER, 2.0 * init_rows * S * S * 9 * h->d.in_channels * h->dims[0],
       (double)init_rows * S * S * (h->d.in_channels * 4 + h->dims[0] * f.es),
       k_initial_conv(x, x_batch, h->init_w, h->init_b, hin0, init_rows, h->d.in_channels, h->dims[0], S, S, f.dt, f.st));
  RC(f.tap("initial", hin0, h->dims[0], h->dims[0], S));
  f.join_event = time_forked ? h->ev_join : nullptr;
  // ---- encoder  src/UNet.py:200-209
  bool pooled_normed = false;   // s(0) holds SiLU(GroupNorm(hin)) of the level about to run
  for (int i = 0; i < L; ++i) {
    const int R = S >> i, cout = h->dims[i + 1];
    const int j = L - 1 - i;                          // decoder level consuming this skip
    const int catc = h->dims[i] + h->dims[i + 1];     // = up channels (dims[i]) + skip channels (dims[i+1])
    uint8_t* cat = f.ws + f.plan.cat[j];
    void* skip = cat + (int64_t)h->dims[i] * f.es;    // skip occupies channels [dims[i], catc)
    void* hin = f.ws + f.plan.hin[i];
    f.want_stats = true;
    RC(f.resblock(h->enc_res[i], hin, h->dims[i], f.s(2), cout, R, true, (i == 0 && share_prefix) ? x_batch : 0, pooled_normed));
    RC(f.tap(("enc" + std::to_string(i) + ".res").c_str(), f.s(2), cout, cout, R));
    RC(f.attn_block(h->enc_attn[i], f.s(2), cout, skip, catc, R));
    RC(f.tap(("enc" + std::to_string(i) + ".attn").c_str(), skip, catc, cout, R));
    // MaxPool2d(2,2) and the next ResNetBlock's block1.norm + SiLU in one kernel (one CTA per sample holds the pooled sample)
    const ResW& nxt = i + 1 < L ? h->enc_res[i + 1] : h->bott1;
    pooled_normed = f.fuse_gn && !h->tap.out && k_pool_group_norm_applicable(R, R, cout, 8, f.dt) && nxt.cin == cout;
    if (pooled_normed)
      PROF(LDM_FAM_GROUP_NORM, 0, (double)batch * R * R * cout * f.es * 1.5,
           k_pool_group_norm(skip, catc, R, R, cout, f.ws + f.plan.hin[i + 1], cout, f.s(0), cout, nxt.g1, nxt.be1, 8, GN_EPS, 1, batch, f.st));
    else
      PROF(LDM_FAM_OTHER, 0, (double)batch * R * R * cout * f.es * 1.25,
           k_maxpool2(skip, catc, f.ws + f.plan.hin[i + 1], cout, batch, R, R, cout, f.dt, f.st));
  }
  if (f.join_event) {  // no encoder block consumed the projection (cannot happen with L >= 1, but a fork must always join)
    LDM_CUDA(cudaStreamWaitEvent(f.st, f.join_event, 0));
    f.join_event = nullptr;
  }
  // ---- bottleneck (no time embedding)  src/UNet.py:287-290
  {
    const int R = S >> L, C = h->dims[L];
    RC(f.resblock(h->bott1, f.ws + f.plan.hin[L], C, f.s(2), C, R, false, 0, pooled_normed));
    RC(f.attn_block(h->bott_attn, f.s(2), C, f.s(3), C, R));
    RC(f.resblock(h->bott2, f.s(3), C, f.s(2), C, R, false));
    RC(f.tap("bottleneck", f.s(2), C, C, R));
  }
  // ---- decoder  src/UNet.py:240-248
  const void* z = f.s(2);
  for (int j = 0; j < L; ++j) {
    const int i = L - 1 - j;
    const int Rin = S >> (i + 1), R = S >> i;
    const UpW& u = h->ups[j];
    const int catc = u.cout + u.cin;  // up(rd[j+1]) + skip(rd[j])
    uint8_t* cat = f.ws + f.plan.cat[j];
    // ConvTranspose2d(k2,s2) as a [M, 4*Cout] GEMM scattered into channels [0, Cout) of the concat buffer
    {
      ConvArgs a;
      a.x = z; a.ldx = u.cin; a.cin = u.cin; a.x2 = nullptr; a.ldx2 = 0; a.cin2 = 0; a.w = u.w; a.bias = u.b;
      a.rowvec = nullptr; a.ld_rowvec = 0; a.res = nullptr; a.ldres = 0; a.y = cat; a.ldy = catc; a.cout = 4 * u.cout;
      a.batch = batch; a.height = Rin; a.width = Rin; a.ksize = 1; a.up2 = 1; a.dtype = f.dt;
      RC(f.conv_args(a));
    }
    f.want_stats = true;
    RC(f.resblock(h->dec_res[j], cat, catc, f.s(2), u.cout, R, true));
    RC(f.attn_block(h->dec_attn[j], f.s(2), u.cout, f.s(3), u.cout, R));
    RC(f.tap(("dec" + std::to_string(j)).c_str(), f.s(3), u.cout, u.cout, R));
    z = f.s(3);
  }
  // ---- final  src/UNet.py:345-348,387
  const bool fuse_final = f.dt == LDM_DT_BF16 && f.impl == 0 && h->dims[0] <= 256 &&
                          (h->d.out_channels + 1) * h->dims[0] * 4 <= 3072 &&
                          !(h->tap.out && h->tap.name == "final.res");
  if (fuse_final) {
    // the 1x1 output projection rides in the epilogue of the last 3x3 GEMM, on the fp32 accumulators
    f.fin_w = h->fin_w; f.fin_b = h->fin_b; f.fin_out = out; f.fin_cout = h->d.out_channels;
    RC(f.resblock(h->final_res, z, h->dims[0], nullptr, h->dims[0], S, false));
    return 0;
  }
  RC(f.resblock(h->final_res, z, h->dims[0], f.s(2), h->dims[0], S, false));
  RC(f.tap("final.res", f.s(2), h->dims[0], h->dims[0], S));
  PROF(LDM_FAM_OTHER, 2.0 * batch * S * S * h->dims[0] * h->d.out_channels,
       (double)batch * S * S * (h->dims[0] * f.es + h->d.out_channels * 4),
       k_final_conv(f.s(2), h->dims[0], h->fin_w, h->fin_b, out, batch, h->dims[0], h->d.out_channels, S * S, f.dt, f.st));
  return 0;
}

int ldm_unet_in_channels(const ldm_unet* h) { return h->d.in_channels; }
int ldm_unet_out_channels(const ldm_unet* h) { return h->d.out_channels; }
int ldm_unet_image_size(const ldm_unet* h) { return h->d.image_size; }
