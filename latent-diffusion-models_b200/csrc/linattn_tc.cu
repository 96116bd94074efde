// LinearAttention with the PreNorm GroupNorm and the to_qkv 1x1 convolution fused in, on tcgen05 / TMEM / TMA (sm_100a):
// replaces src/UNet.py:106-110 (PreNorm), :145 (to_qkv), :149-163 (both softmaxes, ctx = k v^T, out = ctx^T q) for the
// 64-channel sites (the two full-resolution levels and the 16x16 decoder level of the reference UNet).
//
// Round 1 ran this on mma.sync (linattn_qkv_fused_kernel: 0.107 of the tensor peak, issue bound on the legacy HMMA pipe, 16 %
// of a sampling timestep).  Here every contraction is a tcgen05.mma with its accumulator in TMEM.  One persistent CTA per SM
// walks over samples; a sample's tokens are tiled by 128 (the UMMA M):
//   A1  K_t   [128 tok x 128]  = X_t Wk^T            -> per-channel max over ALL tokens (softmax over tokens needs it first)
//   A2  KV_t  [128 tok x 256]  = X_t [Wk; Wv]^T      -> P = exp2((K - max) * r log2e), V  -> bf16 tiles in shared memory, in
//                                                       their natural [token][channel] layout = UMMA's MN-major operand
//       ctx   [128 (h,d) x 128 (h,e)] += P_t^T V_t   (K = tokens; block-diagonal part is the four 32x32 head matrices)
//       Z     [128 (h,d) x 16]        += P_t^T 1     (the softmax denominators, as one more N = 16 MMA against ones)
//   ctx epilogue: (ctx / Z * r + v-constant) * 32^-1/2 -> block-diagonal bf16 matrix in shared memory
//   B   Q_t   [128 tok x 128]  = X_t Wq^T            -> softmax over the 32 channels of each head (within a thread's row)
//       out_t [128 tok x 128]  = softmax(Q_t) ctxBD  -> bf16 [B, N, 128]
// The PreNorm GroupNorm(1, C) is folded in algebraically as in the mma.sync kernel (W (r (x - mu) gamma + beta) =
// r (W diag(gamma)) x + const): the kernel reads the RAW block input and the statistics the producing conv's epilogue left.
// The MMA issuer always queues the NEXT tile's projection before the current tile's second GEMM (ctx / out), and P, V and
// softmax(Q) are double-buffered in shared memory, so an epilogue step overlaps a GEMM; the max pass runs over a two-deep
// accumulator ring.  x tiles stream through a 2-deep TMA ring (from L2: a sample's 128 KB is re-read three times).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kQScale = 0.17677669529663687f;   // 32^-1/2 (src/UNet.py:142)

// shared-memory map (bytes from the 1024-aligned base)
constexpr uint32_t OFF_X = 0;                 // 2 x [128 tok][64 ch] K-major SW128 (TMA)
constexpr uint32_t OFF_WQ = 32768;            // [128][64]
constexpr uint32_t OFF_WK = 49152;            // [128][64]; Wv follows (N = 256 operand)
constexpr uint32_t OFF_WV = 65536;
constexpr uint32_t OFF_PV = 81920;            // 2 x { P (MN-major, 2 blocks of [128 tok][64 ch]) | V (same) }: 64 KB each; in the
                                              // output phase P0 / P1 hold softmax(Q) (K-major, 2 atoms) and V0 the block-diagonal ctx
constexpr uint32_t PV_STRIDE = 65536;
constexpr uint32_t OFF_ONES = 212992;         // 16 k rows x 128 B of bf16 1.0
constexpr uint32_t OFF_BAR = 215040;          // mbarriers + TMEM slot
constexpr uint32_t OFF_F = 216064;            // floats: s_max [4][128] | s_m [128] | s_cq [128] | s_cv [128]
constexpr uint32_t LA_SMEM = OFF_F + (4 * 128 + 3 * 128) * 4 + 1024;   // + alignment slack
static_assert(LA_SMEM <= 227 * 1024, "shared memory plan exceeds 227 KB");

// TMEM columns
constexpr uint32_t TM_ACC = 0;                // K / KV / Q accumulators [0, 256)
constexpr uint32_t TM_OUT = 128;              // output GEMM accumulator (phase B) [128, 256)
constexpr uint32_t TM_CTX = 256;              // ctx [256, 384), Z [384, 400)
constexpr uint32_t TM_Z = 384;

struct LtParams {
  int B, N, tiles;             // samples, tokens per sample, 128-token tiles per sample
  const float* uv;             // [2][384] fold constants (k_fold_prenorm_qkv) or null (x is already normalised)
  const float2* gn_part; int gn_splits;   // GroupNorm(1, C) statistics: > 0 pivoted slabs, < 0 raw {S, Q} slots
  float gn_eps;
  const bf16* x; int ldx;      // raw input (for the statistics' pivot)
  bf16* out;                   // [B, N, 128] (null when to_out is fused)
  // fused to_out 1x1 convolution (src/UNet.py:146): y = out Wout^T + bout, folded per sample into ctxW = ctxBD Wout^T, plus the
  // GroupNorm(1, C) partial sums {S, Q} of y for the apply kernel that follows (slot = 32-row block * 2 + column half)
  const float* bout; bf16* y; int ldy; float2* ystats;
  // ... and, when `o` is set, to_out's GroupNorm(1, C) + the Residual add as well (src/UNet.py:147,:20): the CTA owns the whole
  // sample, so after the last tile it normalises y and writes o = x + GroupNorm(y) itself (y is then only a scratch tensor)
  const float* og; const float* ob; bf16* o; int ldo; float o_eps;
  int debug;                   // LDM_LA_DEBUG (timing experiments only): 1 skip the max pass, 2 skip the output phase
  const int* flag;             // != null: exact-fallback mode behind linattn_tc2_kernel -- only samples with flag[b] != 0 are computed
};

// MN-major SWIZZLE_128B operand: [k rows][64 elements] per 64-wide block of M / N, blocks `lbo` bytes apart
__device__ __forceinline__ uint64_t make_mn_desc(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;             // next group of 8 k rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// order-preserving float <-> int map (for redux.sync max)
__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }
// 16-byte store into a [rows][128 B] SWIZZLE_128B tile: logical chunk c of row r lives at chunk c ^ (r & 7)
__device__ __forceinline__ void st_sw128(uint32_t tile, int r, int c, uint32_t a, uint32_t b, uint32_t cc, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile + r * 128 + ((c ^ (r & 7)) << 4)), "r"(a), "r"(b), "r"(cc), "r"(d)
               : "memory");
}
__device__ __forceinline__ void la_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(320, 1)
linattn_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_wo, const LtParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sb - smem_u32(smem_raw));
  // mbarriers: x ring | weights | K accumulator ring (max pass) | projection done | second GEMM done | epilogue steps
  const uint32_t xfull = sb + OFF_BAR, xempty = xfull + 16, wbar = xfull + 32, kfull = xfull + 40, kempty = xfull + 56,
                 qbar = xfull + 72, obar = xfull + 80, s_bar = xfull + 88, st_bar = xfull + 96, wobar = xfull + 104,
                 tmem_slot = xfull + 112;
  float* s_max = reinterpret_cast<float*>(sgen + OFF_F);   // [4 quarters][128]
  float* s_m = s_max + 512;                                 // [128] per-channel max of k over the sample's tokens
  float* s_cq = s_m + 128;                                  // [128] q constants (fold)
  float* s_cv = s_cq + 128;                                 // [128] v constants (fold), already times 32^-1/2
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sgen + OFF_BAR + 112);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.flag) {
    // fallback launch: almost always nothing is flagged -- leave before any set-up
    pdl_wait();
    int any = 0;
    for (int b = blockIdx.x + (int)threadIdx.x * (int)gridDim.x; b < p.B; b += (int)(gridDim.x * blockDim.x)) any |= p.flag[b];
    if (!__syncthreads_or(any)) return;
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_wo);
    for (int s = 0; s < 2; ++s) {
      mbar_init(xfull + 8 * s, 1); mbar_init(xempty + 8 * s, 1);
      mbar_init(kfull + 8 * s, 1); mbar_init(kempty + 8 * s, 8);
    }
    mbar_init(wbar, 1);
    mbar_init(qbar, 1);
    mbar_init(obar, 1);
    mbar_init(s_bar, 8);
    mbar_init(st_bar, 8);
    mbar_init(wobar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 512; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sb + OFF_ONES + 4 * i), "r"(0x3F803F80u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int tiles = p.tiles;
  const bool do_max = !(p.debug & 1), do_out = !(p.debug & 2);
  const bool fuse_out = p.y != nullptr;          // to_out folded in: the kernel emits y [B, N, 64] and its GroupNorm sums
  constexpr uint32_t OFF_V1 = OFF_PV + PV_STRIDE + 32768;   // free in the output phase: ctxW operand (16 KB) | Wout (16 KB)

  if (warp == 0) {
    // ===================== TMA producer: weights once, then every sample's x tiles three times =====================
    if (lane == 0) {
      mbar_expect_tx(wbar, 3 * 16384);
      tma_load_2d(sb + OFF_WQ, &tmap_w, wbar, 0, 0);
      tma_load_2d(sb + OFF_WK, &tmap_w, wbar, 0, 128);
      tma_load_2d(sb + OFF_WV, &tmap_w, wbar, 0, 256);
      int stage = 0; uint32_t phase = 0;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x)
        for (int pass = 0; pass < 3; ++pass) {
          if (p.flag && !p.flag[b]) break;
          if ((pass == 0 && !do_max) || (pass == 2 && !do_out)) continue;
          for (int t = 0; t < tiles; ++t) {
            mbar_wait(xempty + 8 * stage, phase ^ 1);
            mbar_expect_tx(xfull + 8 * stage, 16384);
            tma_load_2d(sb + OFF_X + stage * 16384, &tmap_x, xfull + 8 * stage, 0, b * p.N + t * 128);
            if (++stage == 2) { stage = 0; phase ^= 1; }
          }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t id128 = make_idesc(128), id256 = make_idesc(256);
      constexpr uint32_t id_ctx = make_idesc(128) | (1u << 15) | (1u << 16);    // P^T V: both operands MN-major
      constexpr uint32_t id_z = make_idesc(16) | (1u << 15) | (1u << 16);
      constexpr uint32_t id_out = make_idesc(128) | (1u << 16);                 // softmax(Q) K-major, ctxBD MN-major
      const uint64_t wq = make_sw128_desc(sb + OFF_WQ), wk = make_sw128_desc(sb + OFF_WK);
      const uint64_t ones = make_mn_desc(sb + OFF_ONES, 2048);
      const uint64_t ctxbd = make_mn_desc(sb + OFF_PV + 32768, 16384);          // V0
      mbar_wait(wbar, 0);
      tc_fence_after();
      int stage = 0; uint32_t xphase = 0;
      // completions consumed so far on the epilogue -> MMA barriers (every barrier is consumed strictly in order and its
      // producer can never run two phases ahead of this thread: see the comments at each wait)
      uint32_t n_s = 0, n_st = 0, n_ke[2] = {0, 0}, n_qc = 0, n_wo = 0;
      auto wait_s = [&]() { mbar_wait(s_bar, n_s & 1); ++n_s; tc_fence_after(); };
      auto wait_st = [&]() { mbar_wait(st_bar, n_st & 1); ++n_st; tc_fence_after(); };
      auto x_desc = [&]() {
        mbar_wait(xfull + 8 * stage, xphase);
        tc_fence_after();
        return make_sw128_desc(sb + OFF_X + stage * 16384);
      };
      auto x_release = [&]() {
        umma_commit(xempty + 8 * stage);
        if (++stage == 2) { stage = 0; xphase ^= 1; }
      };
      auto proj = [&](uint32_t tmem_d, uint64_t w, uint32_t idesc) {   // X_t W^T, K = 64
        const uint64_t xd = x_desc();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, xd + 2 * k, w + 2 * k, idesc, k ? 1u : 0u);
        x_release();
      };
      auto ctx_mma = [&](int buf, bool first) {     // ctx += P^T V, Z += P^T 1 over the 128 tokens of P/V buffer `buf`
        const uint64_t pmn = make_mn_desc(sb + OFF_PV + buf * PV_STRIDE, 16384);
        const uint64_t vmn = make_mn_desc(sb + OFF_PV + buf * PV_STRIDE + 32768, 16384);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          umma_bf16(tmem_base + TM_CTX, pmn + 128 * k, vmn + 128 * k, id_ctx, (first && k == 0) ? 0u : 1u);
          umma_bf16(tmem_base + TM_Z, pmn + 128 * k, ones, id_z, (first && k == 0) ? 0u : 1u);
        }
      };
      for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        if (p.flag && !p.flag[b]) continue;
        // ---- A1: K_t into a two-deep accumulator ring (columns [0,128) / [128,256)); the epilogue takes the column max
        if (do_max) {
          for (int t = 0; t < tiles; ++t) {
            const int kb = t & 1;
            if (t >= 2) { mbar_wait(kempty + 8 * kb, n_ke[kb] & 1); ++n_ke[kb]; tc_fence_after(); }   // step t-2 has read this buffer
            proj(tmem_base + TM_ACC + 128 * kb, wk, id128);
            umma_commit(kfull + 8 * kb);
          }
          // both buffers are rewritten by the first KV projection: the last two steps must be through
          for (int t = tiles > 2 ? tiles - 2 : 0; t < tiles; ++t) { const int kb = t & 1; mbar_wait(kempty + 8 * kb, n_ke[kb] & 1); ++n_ke[kb]; }
          tc_fence_after();
        }
        // ---- A2: KV_t -> (epilogue writes P, V into buffer t & 1) -> ctx; KV_{t+1} is queued BEFORE ctx_t
        proj(tmem_base + TM_ACC, wk, id256);
        umma_commit(qbar); ++n_qc;
        for (int t = 0; t < tiles; ++t) {
          wait_s();                        // P_t, V_t are in shared memory and the accumulator has been read (the epilogue's
                                           // next step needs the next commit below, so it is at most one phase ahead)
          if (t + 1 < tiles) { proj(tmem_base + TM_ACC, wk, id256); umma_commit(qbar); ++n_qc; }
          ctx_mma(t & 1, t == 0);
        }
        umma_commit(qbar);                 // ctx complete
        ++n_qc;
        if (fuse_out) {
          // Wout streams into the (now idle) second V buffer while the epilogue turns ctx into its block-diagonal bf16 form
          mbar_wait(qbar, (n_qc - 1) & 1);   // every ctx GEMM has finished reading that buffer
          mbar_expect_tx(wobar, 16384);
          tma_load_2d(sb + OFF_V1 + 16384, &tmap_wo, wobar, 0, 0);
          tma_load_2d(sb + OFF_V1 + 16384 + 8192, &tmap_wo, wobar, 64, 0);
        }
        wait_s();                          // block-diagonal ctx is in shared memory
        if (fuse_out) {
          // ctxW [128 (h,d) x 64] = ctxBD [128 x 128 (h,e)] Wout^T: to_out applied to the context once per sample
          mbar_wait(wobar, n_wo & 1); ++n_wo;
          tc_fence_after();
          const uint64_t ca = make_sw128_desc(sb + OFF_PV + 32768), wo = make_sw128_desc(sb + OFF_V1 + 16384);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem_base + TM_ACC, ca + (uint64_t)((k >> 2) * 1024 + 2 * (k & 3)), wo + (uint64_t)((k >> 2) * 512 + 2 * (k & 3)),
                      make_idesc(64), k ? 1u : 0u);
          umma_commit(qbar); ++n_qc;
          wait_s();                        // ctxW is in shared memory as the MN-major B operand of the output GEMM
        }
        // ---- B: Q_t -> (epilogue: softmax into buffer t & 1) -> out_t -> (epilogue stores); Q_{t+1} is queued BEFORE out_t
        if (do_out) {
          proj(tmem_base + TM_ACC, wq, id128);
          umma_commit(qbar); ++n_qc;
          for (int t = 0; t < tiles; ++t) {
            wait_s();                      // softmax(Q_t) written, Q accumulator read
            if (t >= 1) wait_st();         // out_{t-1} stored: the output accumulator is free
            if (t + 1 < tiles) { proj(tmem_base + TM_ACC, wq, id128); umma_commit(qbar); ++n_qc; }
            const uint64_t qk = make_sw128_desc(sb + OFF_PV + (t & 1) * PV_STRIDE);
            if (fuse_out) {
              const uint64_t cw = make_mn_desc(sb + OFF_V1, 16384);
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_bf16(tmem_base + TM_OUT, qk + (uint64_t)((k >> 2) * 1024 + 2 * (k & 3)), cw + 128 * k, make_idesc(64) | (1u << 16), k ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k)
                umma_bf16(tmem_base + TM_OUT, qk + (uint64_t)((k >> 2) * 1024 + 2 * (k & 3)), ctxbd + 128 * k, id_out, k ? 1u : 0u);
            }
            umma_commit(obar);
          }
          wait_st();                       // last output tile stored (the next sample's projections reuse these columns)
        }
      }
    }
  } else {
    // ===================== epilogue: 8 warps, (TMEM lane quarter) x (column half) =====================
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;          // token row of the tile / (head, d) row of ctx
    const int et = threadIdx.x - 64;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t n_q = 0, n_o = 0, n_kf[2] = {0, 0};   // completions consumed on the MMA -> epilogue barriers
    auto wait_bar = [&](uint32_t bar, uint32_t& n) {
      if (lane == 0) mbar_wait(bar, n & 1);
      __syncwarp();
      tc_fence_after();
      ++n;
    };
    auto arrive = [&](uint32_t bar, bool wrote_smem) {
      if (wrote_smem) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> the MMA's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
      if (p.flag && !p.flag[b]) continue;
      // per-sample GroupNorm(1, C) scale / shift (see linattn_qkv_fused_kernel) and the fold constants
      float gr = 1.f, gmu = 0.f;
      if (p.uv) {
        const double cnt = (double)p.N * 64.0;
        if (p.gn_splits < 0) {
          double a = 0.0, q2 = 0.0;
          for (int sp = 0; sp < -p.gn_splits; ++sp) { const float2 v = p.gn_part[(int64_t)b * (-p.gn_splits) + sp]; a += (double)v.x; q2 += (double)v.y; }
          const double m1 = a / cnt;
          double var = q2 / cnt - m1 * m1;
          if (var < 0.0) var = 0.0;
          gmu = (float)m1;
          gr = (float)(1.0 / sqrt(var + (double)p.gn_eps));
        } else {
          float a = 0.f, q2 = 0.f;
          for (int sp = 0; sp < p.gn_splits; ++sp) { const float2 v = p.gn_part[(int64_t)b * p.gn_splits + sp]; a += v.x; q2 += v.y; }
          const float K = __bfloat162float(p.x[(int64_t)b * p.N * p.ldx]);
          const float inv_n = 1.0f / (float)cnt;
          const float m1 = a * inv_n;
          const float var = fmaxf(q2 * inv_n - m1 * m1, 0.f);
          gmu = K + m1;
          gr = 1.0f / sqrtf(var + p.gn_eps);
        }
      }
      const float kscale = gr * kLog2e;
      if (et < 128) {
        s_cq[et] = p.uv ? p.uv[384 + et] - gr * gmu * p.uv[et] : 0.f;
        s_cv[et] = p.uv ? kQScale * (p.uv[384 + 256 + et] - gr * gmu * p.uv[256 + et]) : 0.f;
      }
      // ---------------- A1: running max of this warp's 64 columns over its 32 token rows
      int mx[2] = {f2ord(-INFINITY), f2ord(-INFINITY)};    // lane l keeps columns 64*half + l and + 32 + l
      for (int t = 0; t < (do_max ? tiles : 0); ++t) {
        const int kb = t & 1;
        wait_bar(kfull + 8 * kb, n_kf[kb]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(lane_addr + TM_ACC + 128 * kb + 64 * half + 32 * c, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int m = __reduce_max_sync(0xffffffffu, f2ord(__uint_as_float(r[j])));
            if (lane == j) mx[c] = max(mx[c], m);
          }
        }
        arrive(kempty + 8 * kb, false);
      }
      s_max[quarter * 128 + 64 * half + lane] = ord2f(mx[0]);
      s_max[quarter * 128 + 64 * half + 32 + lane] = ord2f(mx[1]);
      la_bar();
      if (et < 128) s_m[et] = do_max ? fmaxf(fmaxf(s_max[et], s_max[128 + et]), fmaxf(s_max[256 + et], s_max[384 + et])) : 0.f;
      la_bar();
      // ---------------- A2: P = exp2((k - max) r log2e) and V, as bf16 MN-major tiles ([token][64 channels] blocks)
      for (int t = 0; t < tiles; ++t) {
        wait_bar(qbar, n_q);
        const uint32_t pbuf = sb + OFF_PV + (t & 1) * PV_STRIDE + half * 16384, vbuf = pbuf + 32768;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t kr[32], vr[32];
          tmem_ld32(lane_addr + TM_ACC + 64 * half + 32 * c, kr);
          tmem_ld32(lane_addr + TM_ACC + 128 + 64 * half + 32 * c, vr);
          tmem_ld_wait();
          const float* mc = s_m + 64 * half + 32 * c;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t pk[4], vk[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float p0 = ex2((__uint_as_float(kr[j + 2 * u]) - mc[j + 2 * u]) * kscale);
              const float p1 = ex2((__uint_as_float(kr[j + 2 * u + 1]) - mc[j + 2 * u + 1]) * kscale);
              pk[u] = pack2(p0, p1);
              vk[u] = pack2(__uint_as_float(vr[j + 2 * u]), __uint_as_float(vr[j + 2 * u + 1]));
            }
            const int chunk = 4 * c + (j >> 3);          // 16-byte chunk inside this half's 64-channel block
            st_sw128(pbuf, row, chunk, pk[0], pk[1], pk[2], pk[3]);
            st_sw128(vbuf, row, chunk, vk[0], vk[1], vk[2], vk[3]);
          }
        }
        arrive(s_bar, true);
      }
      // ---------------- ctx epilogue: row (h, d) -> (ctx / Z * r + cv) * 32^-1/2, block-diagonal, bf16, MN-major B operand
      wait_bar(qbar, n_q);
      {
        const int h = quarter;                           // rows 32h .. 32h+31 are head h
        uint32_t cr[32], zr[32];
        tmem_ld32(lane_addr + TM_CTX + 32 * h, cr);
        tmem_ld32(lane_addr + TM_Z, zr);                  // 16 identical columns of Z (the rest is unused TMEM)
        tmem_ld_wait();
        const float f = kQScale * gr / __uint_as_float(zr[0]);
        // this warp writes its half of the 16 chunks of the row: block `half`, chunks 0..7; the head's 32 columns are chunks
        // 4 (h & 1) .. + 3 of block h >> 1, everything else is zero
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4] = {0u, 0u, 0u, 0u};
          if (half == (h >> 1) && (c >> 2) == (h & 1)) {
            const int e0 = (c & 3) * 8;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              w[u] = pack2(fmaf(__uint_as_float(cr[e0 + 2 * u]), f, s_cv[32 * h + e0 + 2 * u]),
                           fmaf(__uint_as_float(cr[e0 + 2 * u + 1]), f, s_cv[32 * h + e0 + 2 * u + 1]));
          }
          st_sw128(sb + OFF_PV + 32768 + half * 16384, row, c, w[0], w[1], w[2], w[3]);
        }
      }
      arrive(s_bar, true);
      if (fuse_out) {
        // ctxW row (h, d): 64 output channels of to_out -> bf16, MN-major B operand ([k = (h,d)][64 channels], one block)
        wait_bar(qbar, n_q);
        uint32_t r[32];
        tmem_ld32(lane_addr + TM_ACC + 32 * half, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          st_sw128(sb + OFF_V1, row, 4 * half + (j >> 3), pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])),
                   pack2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), pack2(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5])),
                   pack2(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7])));
        arrive(s_bar, true);
      }
      // ---------------- B: softmax over the 32 channels of each head (step S_t), output tile store (step St_t);
      // order S_0, S_1, St_0, S_2, St_1, ...: the softmax of the next tile overlaps the output GEMM of this one
      auto softmax_step = [&](int t) {
        wait_bar(qbar, n_q);
        const uint32_t qbuf = sb + OFF_PV + (t & 1) * PV_STRIDE + half * 16384;
#pragma unroll
        for (int c = 0; c < 2; ++c) {                     // head 2 * half + c: 32 columns of this token row
          uint32_t qr[32];
          tmem_ld32(lane_addr + TM_ACC + 64 * half + 32 * c, qr);
          tmem_ld_wait();
          const float* cq = s_cq + 64 * half + 32 * c;
          float qv[32];
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) { qv[j] = fmaf(__uint_as_float(qr[j]), gr, cq[j]); m = fmaxf(m, qv[j]); }
          const float ml = m * kLog2e;
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) { qv[j] = ex2(fmaf(qv[j], kLog2e, -ml)); s += qv[j]; }
          const float inv = 1.0f / s;
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            st_sw128(qbuf, row, 4 * c + (j >> 3), pack2(qv[j] * inv, qv[j + 1] * inv), pack2(qv[j + 2] * inv, qv[j + 3] * inv),
                     pack2(qv[j + 4] * inv, qv[j + 5] * inv), pack2(qv[j + 6] * inv, qv[j + 7] * inv));
        }
        arrive(s_bar, true);
      };
      auto store_step = [&](int t) {
        wait_bar(obar, n_o);
        if (fuse_out) {
          // y = q~ ctxW + bout: this warp's 32 of the 64 channels; GroupNorm(1, C) partial sums of the fp32 values
          uint32_t r[32];
          tmem_ld32(lane_addr + TM_OUT + 32 * half, r);
          tmem_ld_wait();
          bf16* yrow = p.y + ((int64_t)b * p.N + t * 128 + row) * p.ldy + 32 * half;
          float sS = 0.f, sQ = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) { v[u] = __uint_as_float(r[j + u]) + __ldg(p.bout + 32 * half + j + u); sS += v[u]; sQ = fmaf(v[u], v[u], sQ); }
            uint4 o;
            o.x = pack2(v[0], v[1]); o.y = pack2(v[2], v[3]); o.z = pack2(v[4], v[5]); o.w = pack2(v[6], v[7]);
            *reinterpret_cast<uint4*>(yrow + j) = o;
          }
          sS = warp_sum(sS); sQ = warp_sum(sQ);
          if (lane == 0) {
            p.ystats[(int64_t)b * (p.N / 16) + (t * 4 + quarter) * 2 + half] = make_float2(sS, sQ);
            if (p.o) reinterpret_cast<float2*>(s_max)[(t * 4 + quarter) * 2 + half] = make_float2(sS, sQ);
          }
          arrive(st_bar, false);
          return;
        }
        bf16* orow = p.out + ((int64_t)b * p.N + t * 128 + row) * 128 + 64 * half;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(lane_addr + TM_OUT + 64 * half + 32 * c, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 v;
            v.x = pack2(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
            v.y = pack2(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            v.z = pack2(__uint_as_float(r[j + 4]), __uint_as_float(r[j + 5]));
            v.w = pack2(__uint_as_float(r[j + 6]), __uint_as_float(r[j + 7]));
            *reinterpret_cast<uint4*>(orow + 32 * c + j) = v;
          }
        }
        arrive(st_bar, false);
      };
      if (do_out) {
        softmax_step(0);
        for (int t = 1; t < tiles; ++t) { softmax_step(t); store_step(t - 1); }
        store_step(tiles - 1);
        if (fuse_out && p.o) {
          // o = x + GroupNorm(1, C)(y): statistics from the per-warp partial sums (added in slot order, in double: the same
          // numbers the stand-alone apply kernel would use), then one pass over this sample's y (L2 / L1 resident) and x
          la_bar();                                    // every warp's y rows and partial sums are in place
          if (et < 64) {
            double S = 0.0, Q = 0.0;
            const float2* ps = reinterpret_cast<const float2*>(s_max);
            for (int sl = 0; sl < p.N / 16; ++sl) { S += (double)ps[sl].x; Q += (double)ps[sl].y; }
            const double cnt = (double)p.N * 64.0;
            const double m1 = S / cnt;
            double var = Q / cnt - m1 * m1;
            if (var < 0.0) var = 0.0;
            const float a = __ldg(p.og + et) * (1.0f / sqrtf((float)var + p.o_eps));
            s_cq[et] = a;
            s_cv[et] = fmaf(-(float)m1, a, __ldg(p.ob + et));
          }
          la_bar();
          const int64_t row0 = (int64_t)b * p.N;
          for (int idx = et; idx < p.N * 8; idx += 256) {
            const int tok = idx >> 3, ch = (idx & 7) * 8;
            const uint4 yv = *reinterpret_cast<const uint4*>(p.y + (row0 + tok) * p.ldy + ch);
            const uint4 xv = __ldg(reinterpret_cast<const uint4*>(p.x + (row0 + tok) * p.ldx + ch));
            const __nv_bfloat162* yh = reinterpret_cast<const __nv_bfloat162*>(&yv);
            const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xv);
            uint32_t w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float2 yf = __bfloat1622float2(yh[u]), xf = __bfloat1622float2(xh[u]);
              w[u] = pack2(fmaf(yf.x, s_cq[ch + 2 * u], s_cv[ch + 2 * u]) + xf.x, fmaf(yf.y, s_cq[ch + 2 * u + 1], s_cv[ch + 2 * u + 1]) + xf.y);
            }
            *reinterpret_cast<uint4*>(p.o + (row0 + tok) * p.ldo + ch) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          la_bar();                                    // s_cq / s_cv are rewritten at the start of the next sample
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =====================================================================================================================
// linattn_tc2_kernel: the same block (PreNorm + to_qkv + LinearAttention + to_out.0) re-derived so that the epilogue warps --
// the bottleneck of linattn_tc_kernel (profiles/README.md finding 14) -- touch 5/9 of the elements and never wait for a GEMM:
//   * V is never projected.  ctx = softmax_t(K)^T V and V = X Wv'^T, so ctx = (P^T X) Wv'^T: the kernel accumulates
//     G [128 (h,d) x 64 c] += P_t^T X_t  (the x tile that fed the K projection is also the MN-major B operand) and Z += P_t^T 1.
//   * to_out and Wv are folded into ONE weight-only matrix per head, U_h [c][co] = sum_e Wout[co,(h,e)] Wv'[(h,e),c]
//     (k_fold_to_out, at weight-load time): per sample M [128 (h,d) x 64 co] = (G / Z - mu) [Ucat], one N = 256 GEMM of which row
//     (h,d) reads the 64 columns of head h.  The PreNorm mean is taken off G / Z in fp32, BEFORE the bf16 rounding (a weighted
//     average of x carries the whole mean; rounding it first costs |mu| / sigma in precision).  The remaining constants of the
//     fold (Wv beta through to_out, to_out's bias) collapse into one vector yc [64] because softmax_d(q) sums to one.
//   * no max pass: softmax over tokens is shift invariant, the shift is the sample's FIRST token (P_0 = 1, so Z >= 1); the
//     softmax over the 32 channels of a head is evaluated unshifted.  Either can only fail by overflow / total underflow when
//     a logit leaves +-88 of its reference -- the sample's y then contains inf / NaN, which the GroupNorm partial sums of the
//     store step see for free: the sample is flagged and linattn_tc_kernel (exact two-pass softmax) recomputes it right after
//     (a launch that returns at once when nothing is flagged).
//   * the K projection is issued transposed (W_k X_t^T: lanes = channels, columns = tokens), so P^T goes straight back into
//     tensor memory as the A operand of the G GEMM (tcgen05.st; no shared-memory round trip, and that GEMM reads only [X | 1] from
//     shared memory: the SS form executed in ~106 cycles per MMA on operand fetch alone); the token-0 shift is a per-lane scalar.
//   * 16 epilogue warps in two groups (TMEM lane quarter x 64-column half; each group owns the tiles of one parity), K / Q
//     accumulators double buffered in TMEM, softmax(Q) double buffered in shared memory, x tiles through a 4-deep TMA ring:
//     projection t+2 is queued as soon as the accumulator of step t has been read.
// Work per 128-token tile: K proj (N=128) -> P step -> G, Z;   Q proj (N=128) -> S step -> Y (N=64) -> store step.
constexpr uint32_t L2_XS = 4;                    // x ring depth
constexpr uint32_t L2_X = 0;                     // 4 x [128 tok][64 ch] K-major SW128 (TMA)
constexpr uint32_t L2_WQ = 65536, L2_WK = 81920; // [128][64] each
constexpr uint32_t L2_U = 98304;                 // Ucat [256 (h,co)][64 c] K-major SW128 (32 KB)
constexpr uint32_t L2_PS = 131072;               // 2 x 32 KB: P (MN-major, 2 blocks [128 tok][64 ch]) / softmax(Q) (K-major, 2 atoms):
                                                 // the same bytes either way; G / Z as a K-major A tile aliases the first 16 KB
constexpr uint32_t L2_PS_STRIDE = 32768;
constexpr uint32_t L2_MT = 196608;               // M [128 (h,d)][64 co] MN-major B operand (16 KB)
constexpr uint32_t L2_ONES = 212992;             // 16 k rows x 128 B of bf16 1.0
constexpr uint32_t L2_BAR = 215040;              // mbarriers + TMEM slot (256 B)
constexpr uint32_t L2_F = 215296;                // floats: shift [128] | cq2 [2][128] | yc [2][64]
constexpr uint32_t LA2_SMEM = L2_F + (128 + 256 + 128) * 4 + 1024;
static_assert(LA2_SMEM <= 227 * 1024, "shared memory plan exceeds 227 KB");
constexpr uint32_t T2_ACC = 0;                   // K / Q accumulators: [0,128) and [128,256)
constexpr uint32_t T2_G = 256, T2_Z = 320;       // G [256,320), Z [320,336)
constexpr uint32_t T2_PT = 352;                  // P^T of the two groups, bf16 pairs: [352,416) and [416,480) (K pass only)
constexpr uint32_t T2_M = 256;                   // M accumulators [256,512) (G and Z have been read by then)
constexpr uint32_t T2_Y = 256;                   // Y accumulators [256,320) and [320,384) (M has been read by then)

struct Lt2Params {
  int B, N, tiles;
  const float* uv;             // [2][384] fold constants (k_fold_prenorm_qkv)
  const float2* gn_part; int gn_splits; float gn_eps;
  const bf16* x; int ldx;
  const float* bout;           // [64]
  const float* c12;            // [64]: Wout (Wv beta) (k_fold_to_out)
  bf16* y; int ldy; float2* ystats;
  int* flag;                   // [B]: 1 = recompute this sample with the exact kernel
  int force_flag;              // LDM_LA2_FORCE_FALLBACK (tests): flag every sample
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return r;
}
__device__ __forceinline__ void la2_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M rows on the lanes, K packed two bf16 per 32-bit column) comes from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool TRACE>
__global__ void __launch_bounds__(576, 1)
linattn_tc2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                   const __grid_constant__ CUtensorMap tmap_u, const Lt2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sb = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sb - smem_u32(smem_raw));
  const uint32_t bar0 = sb + L2_BAR;
  const uint32_t xfull = bar0, xempty = bar0 + 32, wbar = bar0 + 64, accfull = bar0 + 72, accempty = bar0 + 88, psfull = bar0 + 104,
                 psempty = bar0 + 120, yfull = bar0 + 136, yempty = bar0 + 152, gfull = bar0 + 168, gnfull = bar0 + 176,
                 mfull = bar0 + 184, mready = bar0 + 192, tmem_slot = bar0 + 200;
  float* s_shift = reinterpret_cast<float*>(sgen + L2_F);   // [128] k of the sample's first token, times r log2e
  float* s_cq2 = s_shift + 128;                              // [2][128] q constants times log2e (by sample parity)
  float* s_yc = s_cq2 + 256;                                 // [2][64]  output constants (by sample parity)
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(sgen + L2_BAR + 200);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // LDM_LA2_TRACE=1 (debug build of the same kernel): CTA 0's MMA thread and two epilogue warps log (tag, clock) pairs
  constexpr int TR_N = TRACE ? 700 : 1;
  uint32_t tr[TR_N];
  int tr_n = 0;
  const bool tr_on = TRACE && blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 2 || warp == 11);
  auto ev = [&](int tag) {
    if (TRACE) {
      if (tr_on && tr_n < TR_N) tr[tr_n++] = ((uint32_t)tag << 24) | ((uint32_t)clock64() & 0xffffffu);
    }
  };
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    prefetch_tmap(&tmap_w);
    prefetch_tmap(&tmap_u);
    for (uint32_t s = 0; s < L2_XS; ++s) { mbar_init(xfull + 8 * s, 1); mbar_init(xempty + 8 * s, 1); }
    for (uint32_t s = 0; s < 2; ++s) {
      mbar_init(accfull + 8 * s, 1); mbar_init(accempty + 8 * s, 8);
      mbar_init(psfull + 8 * s, 8);  mbar_init(psempty + 8 * s, 1);
      mbar_init(yfull + 8 * s, 1);   mbar_init(yempty + 8 * s, 8);
    }
    mbar_init(wbar, 1);
    mbar_init(gfull, 1);
    mbar_init(gnfull, 16);
    mbar_init(mfull, 1);
    mbar_init(mready, 16);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 512; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sb + L2_ONES + 4 * i), "r"(0x3F803F80u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int T = p.tiles;

  if (warp == 0) {
    // ===================== TMA producer: weights once, then every sample's x tiles twice (K pass, Q pass) =====================
    if (lane == 0) {
      mbar_expect_tx(wbar, 2 * 16384 + 32768);
      tma_load_2d(sb + L2_WQ, &tmap_w, wbar, 0, 0);
      tma_load_2d(sb + L2_WK, &tmap_w, wbar, 0, 128);
      tma_load_2d(sb + L2_U, &tmap_u, wbar, 0, 0);
      tma_load_2d(sb + L2_U + 16384, &tmap_u, wbar, 0, 128);
      uint32_t xi = 0;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x)
        for (int pass = 0; pass < 2; ++pass)
          for (int t = 0; t < T; ++t, ++xi) {
            const uint32_t stage = xi % L2_XS;
            mbar_wait(xempty + 8 * stage, ((xi / L2_XS) & 1) ^ 1);
            mbar_expect_tx(xfull + 8 * stage, 16384);
            tma_load_2d(sb + L2_X + stage * 16384, &tmap_x, xfull + 8 * stage, 0, b * p.N + t * 128);
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t id128 = make_idesc(128), id256 = make_idesc(256);
      constexpr uint32_t id_gz = make_idesc(80) | (1u << 16);                   // P^T (tensor memory) [X | 1] (MN-major)
      constexpr uint32_t id_y = make_idesc(64) | (1u << 16);                    // softmax(Q) K-major, M MN-major
      const uint64_t wq = make_sw128_desc(sb + L2_WQ), wk = make_sw128_desc(sb + L2_WK), ud = make_sw128_desc(sb + L2_U);
      const uint64_t gnd = make_sw128_desc(sb + L2_PS);
      const uint64_t mtd = make_mn_desc(sb + L2_MT, 16384);
      mbar_wait(wbar, 0);
      tc_fence_after();
      uint32_t xi = 0;                                 // x tiles consumed so far (ring position)
      uint32_t acc_m = 0, ps_m = 0, y_m = 0;           // bit `buf` = parity of that buffer's next use
      // projection of ring tile `gi` into accumulator `buf`; `release` frees the x stage with the same commit (Q pass)
      // `release`: Q pass -- rows = tokens (X W^T); else K pass -- rows = channels (W X^T: the accumulator IS K^T, so that
      // P^T can go back into tensor memory as the A operand of the G GEMM without a transpose)
      auto proj = [&](uint32_t gi, int buf, uint64_t w, bool release) {
        const uint32_t stage = gi % L2_XS;
        mbar_wait(xfull + 8 * stage, (gi / L2_XS) & 1);
        ev(1);
        mbar_wait(accempty + 8 * buf, ((acc_m >> buf) & 1) ^ 1);
        ev(2);
        tc_fence_after();
        const uint64_t xd = make_sw128_desc(sb + L2_X + stage * 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base + T2_ACC + 128 * buf, (release ? xd : w) + 2 * k, (release ? w : xd) + 2 * k, id128, k ? 1u : 0u);
        umma_commit(accfull + 8 * buf);
        if (release) umma_commit(xempty + 8 * stage);
        acc_m ^= 1u << buf;
      };
      uint32_t it = 0;
      for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
        const uint32_t sp = it & 1;
        // ---- K pass: K_t -> (P step) -> [G | Z] += P_t^T [X_t | 1].  Projection t+2 only needs the accumulator to have been
        // READ by P step t, so it is queued before the wait for P_t itself.
        const uint32_t xa = xi;
        proj(xa, 0, wk, false);
        if (T > 1) proj(xa + 1, 1, wk, false);
        for (int t = 0; t < T; ++t) {
          const int buf = t & 1;
          if (t + 2 < T) proj(xa + t + 2, buf, wk, false);
          mbar_wait(psfull + 8 * buf, (ps_m >> buf) & 1);
          ev(3);
          tc_fence_after();
          const uint32_t stage = (xa + t) % L2_XS;
          // B = [x tile (64 channels) | ones (16 columns)]: N = 80, the second 64-wide block is the 2 KB of ones whatever the
          // k step, so its distance (LBO) shrinks as the start address advances
          const uint64_t xmn = make_mn_desc(sb + L2_X + stage * 16384, (L2_ONES - L2_X) - stage * 16384);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16_ts(tmem_base + T2_G, tmem_base + T2_PT + 64 * buf + 8 * k, xmn + 128 * k - ((uint64_t)(128 * k) << 16), id_gz,
                         (t == 0 && k == 0) ? 0u : 1u);
          umma_commit(psempty + 8 * buf);
          umma_commit(xempty + 8 * stage);
          ps_m ^= 1u << buf;
          ev(4);
        }
        umma_commit(gfull);
        xi = xa + T;
        // ---- the first two Q projections fill the tensor pipe while the epilogue turns G into the A operand of M
        const uint32_t xb = xi;
        proj(xb, 0, wq, true);
        if (T > 1) proj(xb + 1, 1, wq, true);
        mbar_wait(gnfull, sp);
        ev(5);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + T2_M, gnd + 2 * k, ud + 2 * k, id256, k ? 1u : 0u);
        umma_commit(mfull);
        mbar_wait(mready, sp);
        ev(6);
        tc_fence_after();
        // ---- Q pass: Q_t -> (S step) -> Y_t = softmax(Q_t) M -> (store step)
        for (int t = 0; t < T; ++t) {
          const int buf = t & 1;
          if (t + 2 < T) proj(xb + t + 2, buf, wq, true);
          mbar_wait(psfull + 8 * buf, (ps_m >> buf) & 1);
          ev(7);
          mbar_wait(yempty + 8 * buf, ((y_m >> buf) & 1) ^ 1);
          ev(8);
          tc_fence_after();
          const uint64_t sk = make_sw128_desc(sb + L2_PS + buf * L2_PS_STRIDE);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tmem_base + T2_Y + 64 * buf, sk + (uint64_t)((k >> 2) * 1024 + 2 * (k & 3)), mtd + 128 * k, id_y, k ? 1u : 0u);
          umma_commit(yfull + 8 * buf);
          umma_commit(psempty + 8 * buf);
          ps_m ^= 1u << buf;
          y_m ^= 1u << buf;
          ev(9);
        }
        xi = xb + T;
      }
    }
  } else {
    // ===================== epilogue: 16 warps = 2 groups x (TMEM lane quarter) x (64-column half) =====================
    // Group g owns the tiles t = g (mod 2): its own accumulator, its own P / softmax(Q) buffer, its own Y accumulator.  While one
    // group waits (for a GEMM, a barrier, a TMEM load) the other one keeps the MUFU pipe busy.
    const int quarter = warp & 3, cq = (warp - 2) >> 2;
    const int grp = cq >> 1, half = cq & 1;
    const int row = quarter * 32 + lane;          // token row of the tile / (h, d) row of G and M
    const int et = threadIdx.x - 64;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const uint32_t acc_addr = lane_addr + T2_ACC + 128 * grp + 64 * half;
    const uint32_t ps_tile = sb + L2_PS + grp * L2_PS_STRIDE + half * 16384;   // P block / softmax(Q) atom of this column half
    const uint32_t b_accfull = accfull + 8 * grp, b_accempty = accempty + 8 * grp, b_psfull = psfull + 8 * grp,
                   b_psempty = psempty + 8 * grp, b_yfull = yfull + 8 * grp, b_yempty = yempty + 8 * grp;
    uint32_t acc_n = 0, ps_n = 0, y_n = 0;        // uses so far of this group's buffers
    auto wait_bar = [&](uint32_t bar, uint32_t parity) {
      mbar_wait(bar, parity);                     // every lane probes: no divergence, no warp-level hand-off
      tc_fence_after();
    };
    auto arrive = [&](uint32_t bar, bool wrote_smem) {
      if (wrote_smem) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> the MMA's async proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // GroupNorm(1, C) statistics of sample b (left by the producing convolution's epilogue or by k_group_norm_stats): the
    // slots are spread over the lanes and combined by a butterfly, in double -- one fixed order whatever the batch
    auto sample_stats = [&](int b, float& gr, float& gmu) {
      const double cnt = (double)p.N * 64.0;
      if (p.gn_splits < 0) {
        double a = 0.0, q2 = 0.0;
        for (int sp = lane; sp < -p.gn_splits; sp += 32) { const float2 v = __ldg(p.gn_part + (int64_t)b * (-p.gn_splits) + sp); a += (double)v.x; q2 += (double)v.y; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q2 += __shfl_xor_sync(0xffffffffu, q2, o); }
        const double m1 = a / cnt;
        double var = q2 / cnt - m1 * m1;
        if (var < 0.0) var = 0.0;
        gmu = (float)m1;
        gr = (float)(1.0 / sqrt(var + (double)p.gn_eps));
      } else {
        float a = 0.f, q2 = 0.f;
        for (int sp = 0; sp < p.gn_splits; ++sp) { const float2 v = __ldg(p.gn_part + (int64_t)b * p.gn_splits + sp); a += v.x; q2 += v.y; }
        const float K = __bfloat162float(p.x[(int64_t)b * p.N * p.ldx]);
        const float inv_n = 1.0f / (float)cnt;
        const float m1 = a * inv_n;
        const float var = fmaxf(q2 * inv_n - m1 * m1, 0.f);
        gmu = K + m1;
        gr = 1.0f / sqrtf(var + p.gn_eps);
      }
    };
    float gr_n = 1.f, gmu_n = 0.f;
    if ((int)blockIdx.x < p.B) sample_stats(blockIdx.x, gr_n, gmu_n);
    uint32_t it = 0;
    for (int b = blockIdx.x; b < p.B; b += gridDim.x, ++it) {
      ev(10);
      const uint32_t sp = it & 1;
      const float gr = gr_n, gmu = gmu_n;
      const float kscale = gr * kLog2e;
      float* cq2 = s_cq2 + sp * 128;
      float* yc = s_yc + sp * 64;
      if (et < 128) cq2[et] = (p.uv[384 + et] - gr * gmu * p.uv[et]) * kLog2e;
      else if (et < 192) yc[et - 128] = __ldg(p.bout + et - 128) + kQScale * p.c12[et - 128];
      if (et == 192) p.flag[b] = p.force_flag;
      // the shift of the softmax over tokens = k of the sample's first token.  The K accumulator is K^T (lane = channel (h,d),
      // column = token): the 128 lanes of group 0's first column half hold it in column 0 of tile 0
      if (grp == 0 && half == 0) {
        wait_bar(b_accfull, acc_n & 1);           // (consumed again, with the same parity, by P step 0)
        const uint32_t k0 = tmem_ld1(acc_addr);
        tmem_ld_wait();
        s_shift[row] = __uint_as_float(k0) * kscale;
      }
      la2_bar();                                  // publishes the shift, cq2 / yc and the flag reset of this sample
      const float shift = s_shift[row];
      // ---------------- P step (tiles of this group): P^T = exp2((k - k[token 0]) r log2e), bf16 pairs back into tensor memory
      for (int t = grp; t < T; t += 2) {
        wait_bar(b_accfull, acc_n & 1); ++acc_n;
        ev(11);
#pragma unroll
        for (int c = 0; c < 2; ++c) {             // 32 tokens per round
          uint32_t kr[32];
          tmem_ld32(acc_addr + 32 * c, kr);
          tmem_ld_wait();
          if (c == 1) arrive(b_accempty, false);
          uint32_t pw[16];
#pragma unroll
          for (int j = 0; j < 16; ++j)
            pw[j] = pack2(ex2(fmaf(__uint_as_float(kr[2 * j]), kscale, -shift)), ex2(fmaf(__uint_as_float(kr[2 * j + 1]), kscale, -shift)));
          if (c == 0) { ev(12); wait_bar(b_psempty, (ps_n & 1) ^ 1); ev(13); }   // the GEMM that read this buffer two tiles ago
          tmem_st16(lane_addr + T2_PT + 64 * grp + 32 * half + 16 * c, pw);
        }
        tmem_st_wait();
        ev(14);
        arrive(b_psfull, false); ++ps_n;
        ev(15);
      }
      // ---------------- G / Z - mu -> bf16 K-major A operand of the M GEMM (row (h,d): this warp's 16 of the 64 x channels)
      wait_bar(gfull, sp);
      ev(16);
      {
        uint32_t g[16];
        tmem_ld16(lane_addr + T2_G + 16 * cq, g);
        const uint32_t zr = tmem_ld1(lane_addr + T2_Z);
        tmem_ld_wait();
        const float zinv = 1.0f / __uint_as_float(zr);
        float gn[16];
        bool bad = false;
#pragma unroll
        for (int j = 0; j < 16; ++j) { gn[j] = fmaf(__uint_as_float(g[j]), zinv, -gmu); bad |= !(fabsf(gn[j]) < 3.0e38f); }
        st_sw128(sb + L2_PS, row, 2 * cq, pack2(gn[0], gn[1]), pack2(gn[2], gn[3]), pack2(gn[4], gn[5]), pack2(gn[6], gn[7]));
        st_sw128(sb + L2_PS, row, 2 * cq + 1, pack2(gn[8], gn[9]), pack2(gn[10], gn[11]), pack2(gn[12], gn[13]), pack2(gn[14], gn[15]));
        if (bad) p.flag[b] = 1;
      }
      arrive(gnfull, true);
      ev(17);
      // the next sample's statistics while the M GEMM runs
      if (b + (int)gridDim.x < p.B) sample_stats(b + gridDim.x, gr_n, gmu_n);
      ev(18);
      // ---------------- M = r 32^-1/2 (G/Z) Ucat, head block of the row -> bf16 MN-major B operand of the output GEMM
      wait_bar(mfull, sp);
      ev(19);
      {
        uint32_t m[16];
        tmem_ld16(lane_addr + T2_M + 64 * quarter + 16 * cq, m);
        tmem_ld_wait();
        const float f = gr * kQScale;
        st_sw128(sb + L2_MT, row, 2 * cq, pack2(__uint_as_float(m[0]) * f, __uint_as_float(m[1]) * f), pack2(__uint_as_float(m[2]) * f, __uint_as_float(m[3]) * f),
                 pack2(__uint_as_float(m[4]) * f, __uint_as_float(m[5]) * f), pack2(__uint_as_float(m[6]) * f, __uint_as_float(m[7]) * f));
        st_sw128(sb + L2_MT, row, 2 * cq + 1, pack2(__uint_as_float(m[8]) * f, __uint_as_float(m[9]) * f), pack2(__uint_as_float(m[10]) * f, __uint_as_float(m[11]) * f),
                 pack2(__uint_as_float(m[12]) * f, __uint_as_float(m[13]) * f), pack2(__uint_as_float(m[14]) * f, __uint_as_float(m[15]) * f));
      }
      arrive(mready, true);
      ev(20);
      // ---------------- S step: softmax over the 32 channels of heads 2 half, 2 half + 1 (one thread = one token row), unshifted
      const float grl2 = gr * kLog2e;
      auto softmax_step = [&](int t) {
        wait_bar(b_accfull, acc_n & 1); ++acc_n;
        ev(21);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t qr[32];
          tmem_ld32(acc_addr + 32 * c, qr);
          tmem_ld_wait();
          if (c == 1) arrive(b_accempty, false);
          const float4* c4 = reinterpret_cast<const float4*>(cq2 + 64 * half + 32 * c);
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 cc = c4[j >> 2];
            float e0 = ex2(fmaf(__uint_as_float(qr[j]), grl2, cc.x)), e1 = ex2(fmaf(__uint_as_float(qr[j + 1]), grl2, cc.y));
            float e2 = ex2(fmaf(__uint_as_float(qr[j + 2]), grl2, cc.z)), e3 = ex2(fmaf(__uint_as_float(qr[j + 3]), grl2, cc.w));
            s += (e0 + e1) + (e2 + e3);
            qr[j] = __float_as_uint(e0); qr[j + 1] = __float_as_uint(e1); qr[j + 2] = __float_as_uint(e2); qr[j + 3] = __float_as_uint(e3);
          }
          const float inv = 1.0f / s;
          if (c == 0) { ev(23); wait_bar(b_psempty, (ps_n & 1) ^ 1); ev(24); }
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            st_sw128(ps_tile, row, 4 * c + (j >> 3), pack2(__uint_as_float(qr[j]) * inv, __uint_as_float(qr[j + 1]) * inv),
                     pack2(__uint_as_float(qr[j + 2]) * inv, __uint_as_float(qr[j + 3]) * inv),
                     pack2(__uint_as_float(qr[j + 4]) * inv, __uint_as_float(qr[j + 5]) * inv),
                     pack2(__uint_as_float(qr[j + 6]) * inv, __uint_as_float(qr[j + 7]) * inv));
        }
        arrive(b_psfull, true); ++ps_n;
        ev(25);
      };
      // ---------------- store step: y tile t = Y + yc (32 of the 64 channels per thread) + GroupNorm(1, C) partial sums
      auto store_step = [&](int t) {
        ev(26);
        wait_bar(b_yfull, y_n & 1); ++y_n;
        ev(27);
        uint32_t r[32];
        tmem_ld32(lane_addr + T2_Y + 64 * grp + 32 * half, r);
        tmem_ld_wait();
        arrive(b_yempty, false);
        bf16* yrow = p.y + ((int64_t)b * p.N + t * 128 + row) * p.ldy + 32 * half;
        const float4* c4 = reinterpret_cast<const float4*>(yc + 32 * half);
        float sS = 0.f, sQ = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const float4 c0 = c4[j >> 2], c1 = c4[(j >> 2) + 1];
          const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { v[u] = __uint_as_float(r[j + u]) + cc[u]; sS += v[u]; sQ = fmaf(v[u], v[u], sQ); }
          uint4 o;
          o.x = pack2(v[0], v[1]); o.y = pack2(v[2], v[3]); o.z = pack2(v[4], v[5]); o.w = pack2(v[6], v[7]);
          *reinterpret_cast<uint4*>(yrow + j) = o;
        }
        sS = warp_sum(sS); sQ = warp_sum(sQ);
        if (lane == 0) {
          p.ystats[(int64_t)b * (p.N / 16) + (t * 4 + quarter) * 2 + half] = make_float2(sS, sQ);
          if (!(sQ < 3.0e38f)) p.flag[b] = 1;      // inf / NaN anywhere in these rows: the exact kernel redoes the sample
        }
        ev(28);
      };
      // S_t, S_t+2, St_t, S_t+4, St_t+2, ...: the softmax of the group's next tile overlaps the output GEMM of this one
      if (grp < T) {
        softmax_step(grp);
        for (int t = grp + 2; t < T; t += 2) { softmax_step(t); store_step(t - 2); }
        store_step(grp + ((T - 1 - grp) & ~1));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
  if (TRACE) {
    if (tr_on)
      for (int i = 0; i < tr_n; ++i) printf("LA2TRACE w%d %u %u\n", warp, tr[i] >> 24, tr[i] & 0xffffffu);
  }
}

// Ucat [(h, co)][c] = bf16(sum_e Wout[co][(h,e)] Wv[(h,e)][c] gamma[c]);  c1[co] = sum_he Wout[co][he] (Wv beta)[he]
// (Wv beta = uv[384 + 256 + he], left by fold_prenorm_kernel).  One block per output channel co, thread = (h, c) pair.
__global__ void fold_to_out_kernel(const float* __restrict__ wqkv, const float* __restrict__ gamma, const float* __restrict__ uv,
                                   const float* __restrict__ wout, bf16* __restrict__ ucat, float* __restrict__ c1) {
  const int co = blockIdx.x, h = threadIdx.x >> 6, c = threadIdx.x & 63;
  const float* wv = wqkv + (int64_t)(256 + 32 * h) * 64 + c;   // [32 e] rows, 64 floats apart
  const float* wo = wout + (int64_t)co * 128 + 32 * h;         // [32 e]
  float a = 0.f;
#pragma unroll 8
  for (int e = 0; e < 32; ++e) a = fmaf(wo[e], wv[e * 64], a);
  ucat[((int64_t)h * 64 + co) * 64 + c] = __float2bfloat16_rn(a * gamma[c]);
  __shared__ float s_part[4];
  float v = threadIdx.x < 128 ? wout[(int64_t)co * 128 + threadIdx.x] * uv[384 + 256 + threadIdx.x] : 0.f;
  v = warp_sum(v);
  if (threadIdx.x < 128 && (threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) c1[co] = (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]);
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode_la = nullptr;
int g_sms_la[64] = {};

int la_init() {
  int dev = 0;
  LDM_CUDA(cudaGetDevice(&dev));
  if (g_encode_la && g_sms_la[dev & 63]) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  LDM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  LDM_REQUIRE(qres == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  int cc_major = 0;
  LDM_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  LDM_REQUIRE(cc_major == 10, "linattn_tc: tcgen05 kernels need an sm_100-class GPU (found cc %d.x)", cc_major);
  LDM_CUDA(cudaFuncSetAttribute(linattn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA_SMEM));
  LDM_CUDA(cudaFuncSetAttribute(linattn_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA2_SMEM));
  LDM_CUDA(cudaFuncSetAttribute(linattn_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LA2_SMEM));
  int sms = 0;
  LDM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  g_encode_la = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  g_sms_la[dev & 63] = sms;
  return 0;
}

int make_rows_map(CUtensorMap* map, const void* base, int64_t rows, int ld_elems) {
  cuuint64_t gdim[2] = {64, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld_elems * 2};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_la(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(linattn) failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace

bool k_linear_attention_tc_applicable(int cin, int n_tokens, int dtype) {
  static const int off = getenv("LDM_LINATTN_MMA_SYNC") ? atoi(getenv("LDM_LINATTN_MMA_SYNC")) : 0;
  return !off && dtype == LDM_DT_BF16 && cin == 64 && n_tokens % 128 == 0 && n_tokens >= 128;
}

// x [B, N, ldx] raw block input (bf16), wfold [384][64] bf16 (k_fold_prenorm_qkv; rows q | k | v), uv fold constants (or
// null with plain to_qkv weights and an already normalised x), gn_part / gn_splits as k_linear_attention_qkv_prenorm.
int k_linear_attention_tc(const void* x, int ldx, const void* wqkv, const float* uv, const void* gn_part, int gn_splits, float eps,
                          void* out, int batch, int n_tokens, cudaStream_t st, const LinAttnOut* fuse) {
  if (int rc = la_init()) return rc;
  LDM_REQUIRE(n_tokens % 128 == 0 && ldx % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)wqkv & 15) == 0 && ((uintptr_t)out & 15) == 0,
              "linear_attention_tc: needs a multiple of 128 tokens and 16-byte aligned tensors");
  LDM_REQUIRE(!uv || (gn_part && gn_splits != 0), "linear_attention_tc: the folded PreNorm needs GroupNorm statistics");
  if (batch == 0) return 0;
  LtParams p;
  p.B = batch; p.N = n_tokens; p.tiles = n_tokens / 128;
  p.uv = uv; p.gn_part = (const float2*)gn_part; p.gn_splits = gn_splits; p.gn_eps = eps;
  p.x = (const bf16*)x; p.ldx = ldx; p.out = (bf16*)out;
  p.bout = nullptr; p.y = nullptr; p.ldy = 0; p.ystats = nullptr;
  p.og = nullptr; p.ob = nullptr; p.o = nullptr; p.ldo = 0; p.o_eps = 1e-5f;
  if (fuse) {
    LDM_REQUIRE(fuse->wout && fuse->bout && fuse->y && fuse->ystats && fuse->ldy % 8 == 0 && ((uintptr_t)fuse->wout & 15) == 0 &&
                    ((uintptr_t)fuse->y & 15) == 0 && fuse->ystats_bytes >= (int64_t)batch * (n_tokens / 16) * 8,
                "linear_attention_tc: fused to_out needs packed weights, bias, an output and a statistics buffer");
    p.bout = fuse->bout; p.y = (bf16*)fuse->y; p.ldy = fuse->ldy; p.ystats = (float2*)fuse->ystats;
    LDM_REQUIRE(!fuse->o || (fuse->og && fuse->ob && fuse->ldo % 8 == 0 && ((uintptr_t)fuse->o & 15) == 0 && n_tokens / 16 <= 256),
                "linear_attention_tc: fused GroupNorm + residual needs gamma / beta, an aligned output and <= 4096 tokens");
    p.og = fuse->og; p.ob = fuse->ob; p.o = (bf16*)fuse->o; p.ldo = fuse->ldo; p.o_eps = fuse->o_eps;
    if (fuse->nslots_out) *fuse->nslots_out = n_tokens / 16;
  } else {
    LDM_REQUIRE(out, "linear_attention_tc: no output");
  }
  { const char* d = getenv("LDM_LA_DEBUG"); p.debug = d ? atoi(d) : 0; }
  p.flag = nullptr;
  CUtensorMap mx, mw;
  if (int rc = make_rows_map(&mx, x, (int64_t)batch * n_tokens, ldx)) return rc;
  if (int rc = make_rows_map(&mw, wqkv, 384, 64)) return rc;
  CUtensorMap mwo = mw;
  if (fuse) {   // Wout packed [64 out][128 in] bf16: two K atoms of [64 rows][64 channels]
    cuuint64_t gdim[2] = {128, 64};
    cuuint64_t gstr[1] = {128 * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode_la(&mwo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(fuse->wout), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(to_out) failed with CUresult %d", (int)r);
  }
  int dev = 0;
  LDM_CUDA(cudaGetDevice(&dev));
  const int sms = g_sms_la[dev & 63];
  const int grid = batch < sms ? batch : sms;
  static const bool v1_only = getenv("LDM_LINATTN_V1") != nullptr && atoi(getenv("LDM_LINATTN_V1")) != 0;
  if (fuse && !fuse->o && fuse->ucat && fuse->c12 && fuse->flags && uv && !v1_only) {
    // optimistic kernel (no V projection, no max pass, 16 epilogue warps), then the exact kernel over whatever it flagged
    LDM_REQUIRE(((uintptr_t)fuse->ucat & 15) == 0, "linear_attention_tc: unaligned Ucat");
    Lt2Params q;
    q.B = batch; q.N = n_tokens; q.tiles = n_tokens / 128;
    q.uv = uv; q.gn_part = (const float2*)gn_part; q.gn_splits = gn_splits; q.gn_eps = eps;
    q.x = (const bf16*)x; q.ldx = ldx; q.bout = fuse->bout; q.c12 = fuse->c12;
    q.y = (bf16*)fuse->y; q.ldy = fuse->ldy; q.ystats = (float2*)fuse->ystats; q.flag = fuse->flags;
    { const char* d = getenv("LDM_LA2_FORCE_FALLBACK"); q.force_flag = d && atoi(d) ? 1 : 0; }
    CUtensorMap mu;
    if (int rc = make_rows_map(&mu, fuse->ucat, 256, 64)) return rc;
    static const bool trace = getenv("LDM_LA2_TRACE") != nullptr;
    if (trace) LDM_CUDA(ldm_launch_pdl(linattn_tc2_kernel<true>, dim3(grid), dim3(576), (size_t)LA2_SMEM, st, mx, mw, mu, q));
    else LDM_CUDA(ldm_launch_pdl(linattn_tc2_kernel<false>, dim3(grid), dim3(576), (size_t)LA2_SMEM, st, mx, mw, mu, q));
    LDM_LAUNCHED("linattn_tc2");
    p.flag = fuse->flags;
  }
  LDM_CUDA(ldm_launch_pdl(linattn_tc_kernel, dim3(grid), dim3(320), (size_t)LA_SMEM, st, mx, mw, mwo, p));
  LDM_LAUNCHED("linattn_tc");
  return 0;
}

// Weight-only part of the to_out fold of linattn_tc2_kernel (after k_fold_prenorm_qkv, whose uv it reads): ucat bf16 [256][64], c12 fp32 [64]
int k_fold_to_out(const float* wqkv, const float* gamma, const float* uv, const float* wout, void* ucat, float* c12, cudaStream_t st) {
  fold_to_out_kernel<<<64, 256, 0, st>>>(wqkv, gamma, uv, wout, (bf16*)ucat, c12);
  LDM_LAUNCHED("fold_to_out");
  return 0;
}
