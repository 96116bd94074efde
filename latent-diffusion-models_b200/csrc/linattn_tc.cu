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
  LDM_CUDA(ldm_launch_pdl(linattn_tc_kernel, dim3(grid), dim3(320), (size_t)LA_SMEM, st, mx, mw, mwo, p));
  LDM_LAUNCHED("linattn_tc");
  return 0;
}
