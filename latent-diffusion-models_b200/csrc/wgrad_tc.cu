// Convolution weight gradient on tcgen05 (bf16 in, fp32 accumulate in TMEM, fp32 atomics into the OIHW gradient).
//   dW[co][ci][ky][kx] = sum_{n,p} dy[n,p,co] * x[n, p + (ky-1,kx-1), ci]  =  sum_{n,q} dy[n, q - (ky-1,kx-1), co] * x[n,q,ci]
// The contraction runs over PIXELS, so both operands are needed pixel-contiguous ("K-major" with K = pixels): the caller
// hands in channel-major copies xT [B][Cin][H*W], dyT [B][Cout][H*W] (one transpose pass each, k_nhwc_to_nchw_bf16).
//   GEMM rows   = (tap, co) pairs, 128 per tile (for Cout = 64 a tile holds two taps): A = dyT SHIFTED by the tap.
//                 A swizzled TMA box needs a 128-byte inner extent (64 pixels contiguous in memory) and a 16-byte
//                 aligned start, so only the ROW part of the shift goes into the TMA coordinate (flattened pixel index
//                 - dy*W; rows above / below the image fall outside the pixel dimension and are zero-filled by TMA).
//                 The column part is baked into two extra transposed copies of dy: shifted by one pixel to the right
//                 (dx = +1) or to the left (dx = -1) with a zero where the shift would wrap around a row end.
//   GEMM cols   = ci, BN per tile:  B = xT unshifted
//   GEMM K      = 64 consecutive pixels of one image per k-block, split over CTAs
// grid = (row tiles * col tiles, K splits); every CTA accumulates its [128 x BN] slab in TMEM and adds it to dW.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "kernels.h"
#include "tc_common.cuh"

namespace {

using namespace tc;

struct WgParams {
  int cout, cin, taps;        // taps = 1 or 9
  int row_tiles, col_tiles;   // (taps*cout + 127)/128, cin / BN
  int kb_total, kb_per_split; // k-blocks of 64 pixels
  int kb_per_image;           // H*W / 64
  int W;
  int flat;                   // 1: operands are [C][B*HW] (x) and [tap][C][B*HW] (dy, shifted per tap) -- images below 8x8
  float* dw;                  // accumulation target: OIHW when taps == 1, else the GEMM-natural [tap][cout][cin] scratch
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_dy_l,
                const __grid_constant__ CUtensorMap tmap_dy_r, const __grid_constant__ CUtensorMap tmap_x, const WgParams p) {
  constexpr int STAGES = BN == 256 ? 4 : 6;
  constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 64, done_bar = bars + 128, tmem_slot = bars + 136;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(&tmap_dy_l);
    prefetch_tmap(&tmap_dy_r);
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int rt = blockIdx.x / p.col_tiles, ct = blockIdx.x % p.col_tiles;
  const int kb0 = blockIdx.y * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
  const int ci0 = ct * BN;
  // the two 64-row halves of the A tile: virtual rows R = 128*rt + 64*half .. +63  ->  (tap, co0)
  int tap_h[2], co_h[2];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int R = 128 * rt + 64 * hf;
    tap_h[hf] = R / p.cout;
    co_h[hf] = R - tap_h[hf] * p.cout;
  }

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int n0 = kb / p.kb_per_image, p0 = (kb - n0 * p.kb_per_image) * 64;
        mbar_wait(empty_bar + 8 * stage, phase ^ 1);
        mbar_expect_tx(full_bar + 8 * stage, STAGE_BYTES);
        const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
        if (p.flat) {   // every tap has its own pre-shifted copy; k runs over the pixels of the whole batch
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const bool live = tap_h[hf] < p.taps;
            tma_load_3d(sa + hf * 8192, &tmap_dy, full_bar + 8 * stage, kb * 64, live ? co_h[hf] : 0, live ? tap_h[hf] : 0);
          }
          tma_load_3d(sb, &tmap_x, full_bar + 8 * stage, kb * 64, ci0, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
          continue;
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          int dyy = 0, dxx = 0;
          const bool live = tap_h[hf] < p.taps;                 // rows past the last tap: any data, never stored
          if (live && p.taps == 9) { dyy = tap_h[hf] / 3 - 1; dxx = tap_h[hf] % 3 - 1; }
          // dy[q - (dyy,dxx)]: the column shift is in the copy (dx = +1: right-shifted, dx = -1: left-shifted), the row
          // shift in the coordinate
          const CUtensorMap* m = dxx > 0 ? &tmap_dy_r : (dxx < 0 ? &tmap_dy_l : &tmap_dy);
          tma_load_3d(sa + hf * 8192, m, full_bar + 8 * stage, p0 - dyy * p.W, live ? co_h[hf] : 0, n0);
        }
        tma_load_3d(sb, &tmap_x, full_bar + 8 * stage, p0, ci0, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN);
      const uint64_t adesc0 = make_sw128_desc(smem_base), bdesc0 = make_sw128_desc(smem_base + A_BYTES);
      int stage = 0; uint32_t phase = 0;
      uint32_t accum = 0u;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar + 8 * stage, phase);
        tc_fence_after();
        const uint64_t adesc = adesc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
        const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
        umma_bf16(tmem_base, adesc, bdesc, idesc, accum);
#pragma unroll
        for (int k = 1; k < BLOCK_K / UMMA_K; ++k) umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
        accum = 1u;
        umma_commit(empty_bar + 8 * stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done_bar);
    }
  } else {
    // epilogue: rows of the slab -> (tap, co), columns -> ci; fp32 atomics into the OIHW gradient
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int hf = row >> 6;
    const int tap = tap_h[hf], co = co_h[hf] + (row & 63);
    const bool valid = tap < p.taps && kb1 > kb0;
    if (lane == 0) mbar_wait(done_bar, 0);
    __syncwarp();
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // [tap][cout][cin] (== OIHW for a 1x1 filter): a thread's 32 columns are 128 contiguous bytes, added with 16-byte
    // vector reductions -- a quarter of the L2 atomic operations of scalar adds, which bound the large filters
    float* drow = p.dw + ((int64_t)tap * p.cout + co) * p.cin + ci0;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c0, r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + j), "f"(__uint_as_float(r[j])),
                       "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                       : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, BN);
}

// ---------------------------------------------------------------------------------------------------------------------
// The same GEMM with MN-MAJOR operands, read straight from the NHWC tensors -- no transposed copies at all.
// UMMA's MN-major SWIZZLE_128B canonical layout is [k rows][64 elements = 128 bytes] per 64-wide block of M (or N), which
// is exactly what a TMA box {64 channels, pixels...} of an NHWC tensor puts into shared memory: k = pixel, M/N = channel.
//   A (dy, shifted by the tap): box {64 co, W, HB, NB} at (co0, -dx, h0 - dy, n0): out-of-range rows / columns are the
//     conv's zero padding (TMA zero fill); the shift is in PIXEL coordinates, so no alignment constraint applies
//   B (x): boxes {64 ci, W, HB, NB} at (ci0 + 64 j, 0, h0, n0)
// One k-block = 64 pixels = HB rows of NB images (W * HB * NB = 64); per UMMA (K = 16) the descriptors advance 16 rows =
// 2048 bytes; LBO = 8192 (next 64-channel block), SBO = 1024 (next 8 pixel rows).
struct WgMnParams {
  int cout, cin, taps;
  int row_tiles, col_tiles;
  int kb_total, kb_per_split;
  int kb_per_image;           // > 0: H*W / 64 k-blocks per image;  0: one k-block spans NB whole images
  int HB, NB;
  float* dw;                  // [tap][cout][cin] accumulator (== OIHW for 1x1)
  float* dbias;               // bias gradient (accumulated) or null: column sums of dy = A^T * ones, an extra N = 16 MMA
};
__device__ __forceinline__ uint64_t make_sw128_mn_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(8192u >> 4) << 16;             // leading byte offset: next 64-element block along M / N
  d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset: next group of 8 k rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

template <int BN>
__global__ void __launch_bounds__(192, 1)
wgrad_mn_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgMnParams p) {
  constexpr int STAGES = BN == 256 ? 4 : 6;
  constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 64, done_bar = bars + 128, tmem_slot = bars + 136;
  const uint32_t ones = bars + 1024;   // 16 k rows x 128 bytes of bf16 1.0: the B operand of the bias-gradient MMA
  constexpr int TMEM_COLS = BN == 64 ? 128 : (BN == 128 ? 256 : 512);   // accumulator + 16 columns of row sums
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_dy);
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  for (int i = threadIdx.x; i < 512; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(ones + 4 * i), "r"(0x3F803F80u) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int rt = blockIdx.x / p.col_tiles, ct = blockIdx.x % p.col_tiles;
  const int kb0 = blockIdx.y * p.kb_per_split, kb1 = min(p.kb_total, kb0 + p.kb_per_split);
  const int ci0 = ct * BN;
  int tap_h[2], co_h[2];
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) {
    const int R = 128 * rt + 64 * hf;
    tap_h[hf] = R / p.cout;
    co_h[hf] = R - tap_h[hf] * p.cout;
  }
  // the unshifted (centre) tap's rows of dy sum to the bias gradient; one column tile per row tile adds them
  const int centre = p.taps == 9 ? 4 : 0;
  const bool bias_tile = p.dbias != nullptr && ct == 0 && (tap_h[0] == centre || tap_h[1] == centre);

  if (warp == 0) {
    if (lane == 0) {
      int dyy[2], dxx[2], coa[2];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const bool live = tap_h[hf] < p.taps;
        dyy[hf] = (live && p.taps == 9) ? tap_h[hf] / 3 - 1 : 0;
        dxx[hf] = (live && p.taps == 9) ? tap_h[hf] % 3 - 1 : 0;
        coa[hf] = live ? co_h[hf] : 0;
      }
      int stage = 0; uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        int n0, h0;
        if (p.kb_per_image > 0) { n0 = kb / p.kb_per_image; h0 = (kb - n0 * p.kb_per_image) * p.HB; }
        else { n0 = kb * p.NB; h0 = 0; }
        mbar_wait(empty_bar + 8 * stage, phase ^ 1);
        mbar_expect_tx(full_bar + 8 * stage, STAGE_BYTES);
        const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
        // A[q] = dy[q - (dyy, dxx)]
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) tma_load_4d(sa + hf * 8192, &tmap_dy, full_bar + 8 * stage, coa[hf], -dxx[hf], h0 - dyy[hf], n0);
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_4d(sb + j * 8192, &tmap_x, full_bar + 8 * stage, ci0 + 64 * j, 0, h0, n0);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN) | (1u << 15) | (1u << 16);   // A and B MN-major
      constexpr uint32_t idesc1 = make_idesc(16) | (1u << 15) | (1u << 16);
      const uint64_t adesc0 = make_sw128_mn_desc(smem_base), bdesc0 = make_sw128_mn_desc(smem_base + A_BYTES);
      const uint64_t odesc = make_sw128_mn_desc(ones);
      int stage = 0; uint32_t phase = 0;
      uint32_t accum = 0u;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar + 8 * stage, phase);
        tc_fence_after();
        const uint64_t adesc = adesc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
        const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (STAGE_BYTES >> 4));
        umma_bf16(tmem_base, adesc, bdesc, idesc, accum);
#pragma unroll
        for (int k = 1; k < BLOCK_K / UMMA_K; ++k) umma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, 1u);   // +16 rows
        if (bias_tile) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16(tmem_base + BN, adesc + 128 * k, odesc, idesc1, k == 0 ? accum : 1u);
        }
        accum = 1u;
        umma_commit(empty_bar + 8 * stage);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(done_bar);
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int hf = row >> 6;
    const int tap = tap_h[hf], co = co_h[hf] + (row & 63);
    const bool valid = tap < p.taps && kb1 > kb0;
    if (lane == 0) mbar_wait(done_bar, 0);
    __syncwarp();
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    if (bias_tile && valid && tap == centre) {
      uint32_t r[32];
      tmem_ld32(taddr + BN, r);      // 16 identical columns of row sums (the rest of the 32 is unused TMEM)
      tmem_ld_wait();
      atomicAdd(p.dbias + co, __uint_as_float(r[0]));
    }
    float* drow = p.dw + ((int64_t)tap * p.cout + co) * p.cin + ci0;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c0, r);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + j), "f"(__uint_as_float(r[j])),
                       "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                       : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// NHWC (pixel stride ld) -> channel-major bf16 [B][C][HW]; optionally two more copies shifted by one pixel along W with
// zero fill (y_r[h][w] = x[h][w-1], y_l[h][w] = x[h][w+1]) and the column sums (bias gradient)
__global__ void __launch_bounds__(256)
nhwc_to_chw_bf16_kernel(const bf16* __restrict__ x, int ld, bf16* __restrict__ y, bf16* __restrict__ y_l, bf16* __restrict__ y_r,
                        float* __restrict__ colsum, int C, int HW, int W) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows per pass
  float cs = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty + 8 * i;
    float v = 0.f;
    if (p < HW && c0 + tx < C) v = __bfloat162float(x[((int64_t)n * HW + p) * ld + c0 + tx]);
    tile[ty + 8 * i][tx] = v;
    cs += v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, p = p0 + tx;
    if (c < C && p < HW) {
      const bf16 v = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
      const int64_t o = ((int64_t)n * C + c) * HW + p;
      y[o] = v;
      if (y_l) {
        const int w = p % W;
        const bf16 z = __float2bfloat16_rn(0.f);
        if (w != 0) y_l[o - 1] = v; else y_r[o] = z;          // y_l[p-1] = x[p];  y_r has no left neighbour at w = 0
        if (w != W - 1) y_r[o + 1] = v; else y_l[o] = z;      // y_r[p+1] = x[p];  y_l has no right neighbour at w = W-1
      }
    }
  }
  if (colsum) {
    __syncthreads();
    tile[ty][tx] = cs;
    __syncthreads();
    if (ty == 0 && c0 + tx < C) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += tile[r][tx];
      atomicAdd(colsum + c0 + tx, t);
    }
  }
}

// NHWC rows [M = B*HW][C] -> per-tap shifted channel-major copies y[tap][c][m'] with m' = the pixel one filter offset away
// (inside the same image; the buffer is zeroed first so positions nobody writes are the conv's zero padding); column
// sums for the bias gradient.  taps == 1: a plain transpose.
__global__ void __launch_bounds__(256)
nhwc_to_flat_taps_kernel(const bf16* __restrict__ x, int ld, bf16* __restrict__ y, float* __restrict__ colsum, int C, int M,
                         int H, int W, int taps) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.y * 32, m0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float cs = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty + 8 * i;
    float v = 0.f;
    if (m < M && c0 + tx < C) v = __bfloat162float(x[(int64_t)m * ld + c0 + tx]);
    tile[ty + 8 * i][tx] = v;
    cs += v;
  }
  __syncthreads();
  const int m = m0 + tx;
  const int HW = H * W;
  const int n = m / HW, r = m - n * HW, h = r / W, w = r - h * W;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i;
    if (c < C && m < M) {
      const bf16 v = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
      if (taps == 1) {
        y[(int64_t)c * M + m] = v;
      } else {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int hh = h + t / 3 - 1, ww = w + t % 3 - 1;
          if (hh >= 0 && hh < H && ww >= 0 && ww < W) y[((int64_t)t * C + c) * M + n * HW + hh * W + ww] = v;
        }
      }
    }
  }
  if (colsum) {
    __syncthreads();
    tile[ty][tx] = cs;
    __syncthreads();
    if (ty == 0 && c0 + tx < C) {
      float t = 0.f;
#pragma unroll
      for (int r2 = 0; r2 < 8; ++r2) t += tile[r2][tx];
      atomicAdd(colsum + c0 + tx, t);
    }
  }
}

// dw_oihw[co][ci][tap] += nat[tap][co][ci]: grid (cin / 32, cout), 288 threads; both sides coalesced through shared memory
__global__ void __launch_bounds__(288)
wgrad_finish_kernel(const float* __restrict__ nat, float* __restrict__ dw, int cout, int cin) {
  __shared__ float t[9][33];
  const int co = blockIdx.y, ci0 = blockIdx.x * 32;
  const int tap = threadIdx.x >> 5, l = threadIdx.x & 31;
  t[tap][l] = nat[((int64_t)tap * cout + co) * cin + ci0 + l];
  __syncthreads();
  const int o = threadIdx.x;                // 0..287 = (ci local, tap)
  dw[((int64_t)co * cin + ci0) * 9 + o] += t[o % 9][o / 9];
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode_w = nullptr;
int wg_init() {
  if (g_encode_w) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  LDM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  LDM_REQUIRE(qres == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  LDM_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * (16384 + 64 * 128) + 2048));
  LDM_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * (16384 + 128 * 128) + 2048));
  LDM_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (16384 + 256 * 128) + 2048));
  LDM_CUDA(cudaFuncSetAttribute(wgrad_mn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * (16384 + 64 * 128) + 4096 + 1024));
  LDM_CUDA(cudaFuncSetAttribute(wgrad_mn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * (16384 + 128 * 128) + 4096 + 1024));
  LDM_CUDA(cudaFuncSetAttribute(wgrad_mn_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (16384 + 256 * 128) + 4096 + 1024));
  g_encode_w = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}
// NHWC tensor as dims (C, W, H, B); box = {64 channels, W, HB rows, NB images}: one k-block of 64 pixels, MN-major
int make_nhwc_map(CUtensorMap* map, const void* t, int ld, int C, int B, int H, int W, int HB, int NB) {
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)W, (cuuint32_t)HB, (cuuint32_t)NB};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode_w(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(t), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad NHWC operand) failed with CUresult %d", (int)r);
  return 0;
}
// channel-major tensor [B][C][HW] as dims (HW, C, N); box = {64 pixels, `rows` channels, 1 image}: K-major 128-byte rows
int make_chw_map(CUtensorMap* map, const void* t, int B, int C, int64_t HW, int rows) {
  cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)B};
  cuuint64_t gstr[2] = {(cuuint64_t)HW * 2, (cuuint64_t)C * HW * 2};
  cuuint32_t box[3] = {64u, (cuuint32_t)rows, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode_w(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(t), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(wgrad operand) failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace

int k_nhwc_to_chw_bf16(const void* x, int ld, void* y, void* y_l, void* y_r, float* colsum, int batch, int C, int hw, int W,
                       cudaStream_t st) {
  if (batch == 0 || hw == 0) return 0;
  nhwc_to_chw_bf16_kernel<<<dim3((hw + 31) / 32, (C + 31) / 32, batch), 256, 0, st>>>((const bf16*)x, ld, (bf16*)y, (bf16*)y_l,
                                                                                   (bf16*)y_r, colsum, C, hw, W);
  LDM_LAUNCHED("nhwc_to_chw_bf16");
  return 0;
}

int k_nhwc_to_flat_taps_bf16(const void* x, int ld, void* y, float* colsum, int batch, int C, int H, int W, int taps,
                             cudaStream_t st) {
  const int M = batch * H * W;
  if (M == 0) return 0;
  if (taps > 1) LDM_CUDA(cudaMemsetAsync(y, 0, (size_t)taps * C * M * 2, st));
  nhwc_to_flat_taps_kernel<<<dim3((M + 31) / 32, (C + 31) / 32), 256, 0, st>>>((const bf16*)x, ld, (bf16*)y, colsum, C, M, H, W, taps);
  LDM_LAUNCHED("nhwc_to_flat_taps_bf16");
  return 0;
}

// images too small for 64-pixel k-blocks inside one image: k runs over the pixels of the whole batch instead
bool k_conv_wgrad_tc_flat_applicable(int cin, int cout, int batch, int H, int W, int ksize, int dtype) {
  if (dtype != LDM_DT_BF16 || (ksize != 1 && ksize != 3)) return false;
  if (cin % 64 != 0 || cout % 64 != 0) return false;
  if (k_conv_wgrad_tc_applicable(cin, cout, H, W, ksize, dtype)) return false;
  const int64_t M = (int64_t)batch * H * W;
  if (M % 8 != 0 || M < 64 || H * W > 64) return false;   // 16-byte row pitch; 9 shifted copies only pay for small images
  return getenv("LDM_WGRAD_FFMA") == nullptr;
}

bool k_conv_wgrad_tc_applicable(int cin, int cout, int H, int W, int ksize, int dtype) {
  if (dtype != LDM_DT_BF16 || (ksize != 1 && ksize != 3)) return false;
  if (cin % 64 != 0 || cout % 64 != 0) return false;
  if ((H * W) % 64 != 0 || W % 8 != 0) return false;   // k-blocks of 64 consecutive pixels; row shifts 16-byte aligned
  return getenv("LDM_WGRAD_FFMA") == nullptr;
}

static int wg_launch(const void* xT, int cin, const void* dyT, const void* dyT_l, const void* dyT_r, int cout, float* dw,
                     float* nat, int batch, int H, int W, int ksize, int flat, cudaStream_t st);

// xT [B][cin][HW]; dyT, dyT_l, dyT_r [B][cout][HW] (bf16, channel-major; _l / _r: shifted one pixel left / right with zero
// fill, only read by 3x3 filters); dw OIHW fp32, ACCUMULATED
int k_conv_wgrad_tc(const void* xT, int cin, const void* dyT, const void* dyT_l, const void* dyT_r, int cout, float* dw,
                    float* nat, int batch, int H, int W, int ksize, cudaStream_t st) {
  if (int rc = wg_init()) return rc;
  LDM_REQUIRE(k_conv_wgrad_tc_applicable(cin, cout, H, W, ksize, LDM_DT_BF16), "conv_wgrad_tc: unsupported shape");
  if (batch == 0) return 0;
  return wg_launch(xT, cin, dyT, dyT_l, dyT_r, cout, dw, nat, batch, H, W, ksize, 0, st);
}

// xF [cin][B*HW]; dyF [taps][cout][B*HW], copy t shifted by filter offset t (k_nhwc_to_flat_taps_bf16); dw ACCUMULATED
int k_conv_wgrad_tc_flat(const void* xF, int cin, const void* dyF, int cout, float* dw, float* nat, int batch, int H, int W,
                         int ksize, cudaStream_t st) {
  if (int rc = wg_init()) return rc;
  LDM_REQUIRE(k_conv_wgrad_tc_flat_applicable(cin, cout, batch, H, W, ksize, LDM_DT_BF16), "conv_wgrad_tc_flat: unsupported shape");
  return wg_launch(xF, cin, dyF, nullptr, nullptr, cout, dw, nat, batch, H, W, ksize, 1, st);
}

static int wg_launch(const void* xT, int cin, const void* dyT, const void* dyT_l, const void* dyT_r, int cout, float* dw,
                     float* nat, int batch, int H, int W, int ksize, int flat, cudaStream_t st) {
  WgParams p;
  p.cout = cout; p.cin = cin; p.taps = ksize * ksize; p.W = W; p.flat = flat;
  // 3x3: accumulate in the GEMM-natural layout (zeroed here), transpose-add into OIHW afterwards
  LDM_REQUIRE(ksize == 1 || nat != nullptr, "conv_wgrad_tc: 3x3 filters need the natural-layout scratch");
  LDM_REQUIRE(((uintptr_t)dw & 15) == 0 && ((uintptr_t)nat & 15) == 0, "conv_wgrad_tc: gradient buffers must be 16-byte aligned");
  p.dw = ksize == 1 ? dw : nat;
  if (ksize != 1) LDM_CUDA(cudaMemsetAsync(nat, 0, (size_t)9 * cout * cin * sizeof(float), st));
  const int hw = H * W;
  const int64_t M = (int64_t)batch * hw;
  p.kb_per_image = flat ? 1 : hw / 64;
  p.kb_total = flat ? (int)((M + 63) / 64) : batch * p.kb_per_image;
  int bn = 64;
  for (int c : {256, 128}) if (cin % c == 0) { bn = c; break; }
  p.row_tiles = (p.taps * cout + 127) / 128;
  p.col_tiles = cin / bn;
  const int tiles = p.row_tiles * p.col_tiles;
  // one CTA per SM is resident (the operand ring takes most of the shared memory), so more than one wave of split-K CTAs
  // only multiplies the fp32 atomics of the epilogue (128 x BN per CTA), which is what bounds the large filters
  static const int wg_ctas = getenv("LDM_WGRAD_CTAS") ? atoi(getenv("LDM_WGRAD_CTAS")) : 100;
  int splits = (wg_ctas + tiles - 1) / tiles;
  // long contractions (the 32x32 layers at large batch: thousands of k-blocks, few tiles) are bound by the k-loop, not by
  // the epilogue: give them two waves
  if (p.kb_total / splits > 96) splits = (2 * 148 + tiles - 1) / tiles;
  if (splits > p.kb_total) splits = p.kb_total;
  if (splits < 1) splits = 1;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  CUtensorMap mdy, mdl, mdr, mx;
  if (flat) {
    if (int rc = make_chw_map(&mdy, dyT, p.taps, cout, M, 64)) return rc;
    mdl = mdy; mdr = mdy;
    if (int rc = make_chw_map(&mx, xT, 1, cin, M, bn)) return rc;
  } else {
    if (int rc = make_chw_map(&mdy, dyT, batch, cout, hw, 64)) return rc;
    if (int rc = make_chw_map(&mdl, dyT_l ? dyT_l : dyT, batch, cout, hw, 64)) return rc;
    if (int rc = make_chw_map(&mdr, dyT_r ? dyT_r : dyT, batch, cout, hw, 64)) return rc;
    if (int rc = make_chw_map(&mx, xT, batch, cin, hw, bn)) return rc;
  }
  const dim3 grid(tiles, splits);
  switch (bn) {
    case 256: wgrad_tc_kernel<256><<<grid, 192, 4 * (16384 + 256 * 128) + 2048, st>>>(mdy, mdl, mdr, mx, p); break;
    case 128: wgrad_tc_kernel<128><<<grid, 192, 6 * (16384 + 128 * 128) + 2048, st>>>(mdy, mdl, mdr, mx, p); break;
    default: wgrad_tc_kernel<64><<<grid, 192, 6 * (16384 + 64 * 128) + 2048, st>>>(mdy, mdl, mdr, mx, p); break;
  }
  LDM_LAUNCHED("conv_wgrad_tc");
  if (ksize != 1) {
    wgrad_finish_kernel<<<dim3(cin / 32, cout), 288, 0, st>>>(nat, dw, cout, cin);
    LDM_LAUNCHED("conv_wgrad_finish");
  }
  return 0;
}


// ---- MN-major variant: operands are the NHWC tensors themselves
bool k_conv_wgrad_mn_applicable(int cin, int cout, int H, int W, int ksize, int dtype) {
  if (dtype != LDM_DT_BF16 || (ksize != 1 && ksize != 3)) return false;
  if (cin % 64 != 0 || cout % 64 != 0) return false;
  if (W < 1 || W > 64 || 64 % W != 0) return false;
  const int HB = 64 / W < H ? 64 / W : H;
  if (H % HB != 0 || 64 % (W * HB) != 0) return false;
  return getenv("LDM_WGRAD_KMAJOR") == nullptr && getenv("LDM_WGRAD_FFMA") == nullptr;
}
// x NHWC [B][H][W] (pixel stride ldx), dy NHWC (pixel stride lddy); dw OIHW fp32 ACCUMULATED; nat: [9][cout][cin] fp32 scratch
// for 3x3 filters (zeroed here)
int k_conv_wgrad_mn(const void* x, int ldx, int cin, const void* dy, int lddy, int cout, float* dw, float* dbias, float* nat,
                    int batch, int H, int W, int ksize, cudaStream_t st) {
  if (int rc = wg_init()) return rc;
  LDM_REQUIRE(k_conv_wgrad_mn_applicable(cin, cout, H, W, ksize, LDM_DT_BF16), "conv_wgrad_mn: unsupported shape");
  LDM_REQUIRE(ldx % 8 == 0 && lddy % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0,
              "conv_wgrad_mn: operands must be 16-byte aligned with pixel strides that are multiples of 8");
  LDM_REQUIRE(ksize == 1 || nat != nullptr, "conv_wgrad_mn: 3x3 filters need the natural-layout scratch");
  LDM_REQUIRE(((uintptr_t)dw & 15) == 0 && ((uintptr_t)nat & 15) == 0, "conv_wgrad_mn: gradient buffers must be 16-byte aligned");
  if (batch == 0) return 0;
  WgMnParams p;
  p.cout = cout; p.cin = cin; p.taps = ksize * ksize;
  p.HB = 64 / W < H ? 64 / W : H;
  p.NB = 64 / (W * p.HB);
  p.kb_per_image = p.NB == 1 ? (H * W) / 64 : 0;
  p.kb_total = p.NB == 1 ? batch * p.kb_per_image : (batch + p.NB - 1) / p.NB;
  p.dw = ksize == 1 ? dw : nat;
  p.dbias = dbias;
  if (ksize != 1) LDM_CUDA(cudaMemsetAsync(nat, 0, (size_t)9 * cout * cin * sizeof(float), st));
  int bn = 64;
  for (int c : {256, 128}) if (cin % c == 0) { bn = c; break; }
  p.row_tiles = (p.taps * cout + 127) / 128;
  p.col_tiles = cin / bn;
  const int tiles = p.row_tiles * p.col_tiles;
  static const int wg_ctas = getenv("LDM_WGRAD_CTAS") ? atoi(getenv("LDM_WGRAD_CTAS")) : 100;
  int splits = (wg_ctas + tiles - 1) / tiles;
  if (p.kb_total / splits > 96) splits = (2 * 148 + tiles - 1) / tiles;
  if (splits > p.kb_total) splits = p.kb_total;
  if (splits < 1) splits = 1;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  CUtensorMap mdy, mx;
  if (int rc = make_nhwc_map(&mdy, dy, lddy, cout, batch, H, W, p.HB, p.NB)) return rc;
  if (int rc = make_nhwc_map(&mx, x, ldx, cin, batch, H, W, p.HB, p.NB)) return rc;
  const dim3 grid(tiles, splits);
  switch (bn) {
    case 256: wgrad_mn_kernel<256><<<grid, 192, 4 * (16384 + 256 * 128) + 4096 + 1024, st>>>(mdy, mx, p); break;
    case 128: wgrad_mn_kernel<128><<<grid, 192, 6 * (16384 + 128 * 128) + 4096 + 1024, st>>>(mdy, mx, p); break;
    default: wgrad_mn_kernel<64><<<grid, 192, 6 * (16384 + 64 * 128) + 4096 + 1024, st>>>(mdy, mx, p); break;
  }
  LDM_LAUNCHED("conv_wgrad_mn");
  if (ksize != 1) {
    wgrad_finish_kernel<<<dim3(cin / 32, cout), 288, 0, st>>>(nat, dw, cout, cin);
    LDM_LAUNCHED("conv_wgrad_finish");
  }
  return 0;
}
