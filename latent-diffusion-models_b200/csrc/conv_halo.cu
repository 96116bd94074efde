// 3x3 (pad 1) convolution as an implicit GEMM whose A operand is loaded ONCE per tile instead of once per tap
// (tcgen05 / TMEM / TMA, bf16 NHWC, sm_100a).  Replaces F.conv2d behind src/UNet.py:54 for the full-resolution
// layers, where Cout is small (64) and conv_tc_kernel is bound by L2->smem operand traffic, not by the tensor pipe.
//
// Geometry.  Think of every image as a zero-padded plane of (H+2) x P pixels, P = W+2, and number the padded pixels
// linearly: q = r*P + c.  Shifting by tap (dy,dx) is then a shift by dy*P+dx in q for EVERY pixel -- no row-end
// special cases, because the halo columns are real (zero) entries of the plane.  An M tile is 128 consecutive padded
// positions; the rows that fall on halo columns are computed and discarded (W/P = 94 % of the MMA work is useful at
// W=32).
//   * one 4-D TMA box {64 ch, P, RB rows, 1 image} at (c0, -1, h0, n) brings in the tile plus its halo; out-of-bounds
//     rows/columns are zero-filled by TMA, which is exactly the padding.  The box lands as consecutive 128-byte
//     pixel rows with the 128-byte swizzle, i.e. as a K-major SWIZZLE_128B UMMA operand whose M index is q.
//   * the A descriptor of tap (dy,dx) is the same slab with the start address advanced by (dy*P+dx) pixel rows
//     (x128 bytes).  The swizzle XOR is a function of the shared-memory address bits, so a start address that is
//     128-byte but not 1024-byte aligned reads exactly the rows TMA wrote.
//   => L2->smem A traffic per tile drops from 9 x 16 KB to one ~30 KB slab per 64 input channels.
//   * weights: [Cout][K] K-major tiles, one per k-block; when the whole filter fits next to the A ring it is loaded
//     once per CTA and stays resident (all 32x32 layers of the reference UNet), otherwise it streams through a ring.
//   * an optional second source (the ResNetBlock 1x1 shortcut, K-concatenated) uses the same slab with the centre tap.
// Roles (320 threads, persistent, 1 CTA/SM): warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue; TMEM holds
// two accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "kernels.h"
#include "tc_common.cuh"

// timeline probe (LDM_HALO_DEBUG bit 2): CTA 0 records globaltimer stamps of its first tiles
__device__ unsigned long long g_halo_dbg[1024];
extern "C" int ldm_debug_read_halo(unsigned long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_halo_dbg, sizeof(unsigned long long) * (n < 1024 ? n : 1024)) == cudaSuccess ? 0 : -1;
}

namespace {

using namespace tc;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define DBG_STAMP(slot)                                                                              \
  do {                                                                                               \
    if ((p.debug & 4) && blockIdx.x == 0 && iter < 32) {                                             \
      g_halo_dbg[iter * 16 + (slot)] = gtime();                                                      \
      g_halo_dbg[512 + iter * 16 + (slot)] = (unsigned long long)clock64();                          \
    }                                                                                                \
  } while (0)

struct HaloParams {
  int B, H, W, P;          // batch, spatial size, padded pitch W+2
  int RB;                  // slab rows
  int tiles_per_image, num_tiles;
  int slabs_main;          // cin/64
  int slabs_total;         // + cin2/64 (centre tap only)
  int n_kb;                // 9*slabs_main + slabs2
  int a_stage_bytes;       // RB*P*128 rounded up to 1024
  int a_stages, b_stages;  // ring depths; b_stages >= n_kb means the filter is resident
  int cout;                // == BLOCK_N (one N tile)
  int a_tx_bytes, w_start; // TMA box bytes, first W coordinate (-1)
  int res_mod;             // > 0: residual image index = n % res_mod
  int debug;               // profiling knob (LDM_HALO_DEBUG bit 0: no epilogue memory traffic, bit 1: no MMAs)
  int base_offset_mode;    // experiment knob: 0 = descriptor base_offset 0, 1 = (start >> 7) & 7
  const float* bias;
  const float* rowvec; int ld_rowvec;
  const bf16* res; int ldres;
  bf16* y; int ldy;
  const float* fin_w; const float* fin_b; float* fin_out; int fin_cout;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(320, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                 const __grid_constant__ CUtensorMap tmap_b, const HaloParams p) {
  constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_b = smem_base;
  const uint32_t smem_a = smem_b + p.b_stages * B_TILE_BYTES;
  const uint32_t bars = smem_a + p.a_stages * p.a_stage_bytes;
  // barrier map: afull[8] aempty[8] bfull[32] bempty[32] tfull[2] tempty[2] | tmem slot
  const uint32_t afull = bars, aempty = bars + 64, bfull = bars + 128, bempty = bars + 384, tfull = bars + 640,
                 tempty = bars + 656, tmem_slot = bars + 672;
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bars + 1024 - smem_u32(smem_raw)));  // [BLOCK_N]
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_a2);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < 8; ++s) { mbar_init(afull + 8 * s, 1); mbar_init(aempty + 8 * s, 1); }
    for (int s = 0; s < 32; ++s) { mbar_init(bfull + 8 * s, 1); mbar_init(bempty + 8 * s, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull + 8 * i, 1); mbar_init(tempty + 8 * i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();      // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const bool resident = p.b_stages >= p.n_kb;
  const int PL_rows = p.H + 2;
  const int slabs2 = p.slabs_total - p.slabs_main;
  (void)slabs2;

  // tile -> image n = tile / tiles_per_image; q0 = P + (tile % tiles_per_image) * 128 is the plane position of the
  // tile's output row 0; the slab starts at padded row r0 = floor((q0 - P - 1) / P) (the lowest row any tap touches).

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int P = p.P, a_stages = p.a_stages, b_stages = p.b_stages, a_stage_bytes = p.a_stage_bytes;
      const int slabs_main = p.slabs_main, slabs_total = p.slabs_total, num_tiles = p.num_tiles;
      const int tiles_per_image = p.tiles_per_image, w_start = p.w_start;
      const uint32_t a_tx = (uint32_t)p.a_tx_bytes;
      // packed filter K order is (tap, channel block) for the main source, then the second source's blocks;
      // B slots are filled in CONSUMPTION order (slab-major, taps inside)
      auto kblock_of = [&](int s, int t) { return s < slabs_main ? t * slabs_main + s : 9 * slabs_main + (s - slabs_main); };
      if (resident) {
        int slot = 0;
        for (int s = 0; s < slabs_total; ++s)
          for (int t = 0; t < (s < slabs_main ? 9 : 1); ++t, ++slot) {
            mbar_expect_tx(bfull + 8 * slot, B_TILE_BYTES);
            tma_load_2d(smem_b + slot * B_TILE_BYTES, &tmap_b, bfull + 8 * slot, kblock_of(s, t) * BLOCK_K, 0);
          }
      }
      int astage = 0; uint32_t aphase = 0;
      int bstage = 0; uint32_t bphase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
        const int n = tile / tiles_per_image;
        const int q0 = P + (tile - n * tiles_per_image) * TILE_M;
        const int lo = q0 - P - 1;
        const int r0 = lo >= 0 ? lo / P : -1;
        for (int s = 0; s < slabs_total; ++s) {
          DBG_STAMP(8);
          mbar_wait(aempty + 8 * astage, aphase ^ 1);
          DBG_STAMP(9);
          mbar_expect_tx(afull + 8 * astage, a_tx);
          const bool main_src = s < slabs_main;
          const int c0 = (main_src ? s : s - slabs_main) * BLOCK_K;
          tma_load_4d(smem_a + astage * a_stage_bytes, main_src ? &tmap_a : &tmap_a2, afull + 8 * astage, c0, w_start,
                      r0 - 1, n);
          if (++astage == a_stages) { astage = 0; aphase ^= 1; }
          if (!resident) {
            for (int t = 0; t < (main_src ? 9 : 1); ++t) {
              mbar_wait(bempty + 8 * bstage, bphase ^ 1);
              mbar_expect_tx(bfull + 8 * bstage, B_TILE_BYTES);
              tma_load_2d(smem_b + bstage * B_TILE_BYTES, &tmap_b, bfull + 8 * bstage, kblock_of(s, t) * BLOCK_K, 0);
              if (++bstage == b_stages) { bstage = 0; bphase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One thread feeds the tensor pipe; at BLOCK_N = 64 an MMA retires every ~50 cycles, so the loop carries no
    // parameter loads, no divisions and no debug branches: everything below lives in registers.
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_N);
      constexpr uint32_t B_DESC_STEP = B_TILE_BYTES >> 4;
      const int P = p.P, a_stages = p.a_stages, b_stages = p.b_stages, a_stage_bytes = p.a_stage_bytes;
      const int slabs_main = p.slabs_main, slabs_total = p.slabs_total, num_tiles = p.num_tiles, n_kb = p.n_kb;
      const int tiles_per_image = p.tiles_per_image;
      const bool issue = !(p.debug & 2);
      const uint32_t bo_mode = p.base_offset_mode;
      // descriptor-unit (16-byte) displacement of each tap inside the slab: (dy*P + dx) pixel rows of 128 bytes
      int tap_delta[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) tap_delta[t] = ((t / 3 - 1) * P + (t % 3 - 1)) * 8;
      const uint64_t bdesc0 = make_sw128_desc(smem_b);
      if (resident) {  // the whole filter arrives once; after that the loop never looks at a B barrier again
        for (int kb = 0; kb < n_kb; ++kb) mbar_wait(bfull + 8 * kb, 0);
        tc_fence_after();
      }
      int astage = 0; uint32_t aphase = 0;
      int bstage = 0; uint32_t bphase = 0;
      int iter = 0;
      int tin = blockIdx.x % tiles_per_image;  // tile index inside its image, advanced incrementally
      const int tstep = gridDim.x % tiles_per_image;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
        const int q0 = P + tin * TILE_M;
        const int lo = q0 - P - 1;
        const int r0 = lo >= 0 ? lo / P : -1;
        const int off0 = q0 - r0 * P;  // slab row of the tile's first output position (centre tap)
        tin += tstep; if (tin >= tiles_per_image) tin -= tiles_per_image;
        const int acc = iter & 1;
        DBG_STAMP(0);
        mbar_wait(tempty + 8 * acc, ((iter >> 1) & 1) ^ 1);
        DBG_STAMP(1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        uint32_t first_mma = 0u;  // accumulate flag of the very first MMA of the tile
        int kb = 0;
        for (int s = 0; s < slabs_total; ++s) {
          mbar_wait(afull + 8 * astage, aphase);
          DBG_STAMP(2);
          tc_fence_after();
          const uint32_t a_addr = smem_a + astage * a_stage_bytes + (uint32_t)off0 * 128u;
          uint64_t adesc0 = make_sw128_desc(a_addr);
          if (s < slabs_main) {
            if (resident) {
              uint64_t bdesc = bdesc0 + (uint64_t)kb * B_DESC_STEP;
#pragma unroll
              for (int t = 0; t < 9; ++t) {
                uint64_t adesc = adesc0 + (uint64_t)(int64_t)tap_delta[t];
                if (bo_mode == 1) adesc |= (uint64_t)(((a_addr >> 7) + (uint32_t)(tap_delta[t] >> 3)) & 7u) << 49;
                if (issue) {
#pragma unroll
                  for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                    umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (t > 0 || k > 0) ? 1u : first_mma);
                }
                bdesc += B_DESC_STEP;
              }
              kb += 9;
            } else {
              for (int t = 0; t < 9; ++t, ++kb) {
                mbar_wait(bfull + 8 * bstage, bphase);
                tc_fence_after();
                const uint64_t adesc = adesc0 + (uint64_t)(int64_t)(((t / 3 - 1) * P + (t % 3 - 1)) * 8);
                const uint64_t bdesc = bdesc0 + (uint64_t)bstage * B_DESC_STEP;
                if (issue) {
#pragma unroll
                  for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                    umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (t > 0 || k > 0) ? 1u : first_mma);
                }
                umma_commit(bempty + 8 * bstage);
                if (++bstage == b_stages) { bstage = 0; bphase ^= 1; }
              }
            }
          } else {  // second source: centre tap only
            if (!resident) { mbar_wait(bfull + 8 * bstage, bphase); tc_fence_after(); }
            const uint64_t bdesc = bdesc0 + (uint64_t)(resident ? kb : bstage) * B_DESC_STEP;
            if (issue) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_bf16(tmem_d, adesc0 + 2 * k, bdesc + 2 * k, idesc, 1u);
            }
            if (!resident) { umma_commit(bempty + 8 * bstage); if (++bstage == b_stages) { bstage = 0; bphase ^= 1; } }
            ++kb;
          }
          first_mma = 1u;
          umma_commit(aempty + 8 * astage);  // the slab is free once its taps' MMAs retire
          if (++astage == a_stages) { astage = 0; aphase ^= 1; }
        }
        umma_commit(tfull + 8 * acc);
        DBG_STAMP(3);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // 8 epilogue warps: two per TMEM lane quarter, each taking half of the tile's columns (one warp per scheduler
    // would leave every TMEM-load / convert / store chain exposed)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int COLS = BLOCK_N / 2;
    const int cw = half * COLS;
    const int row = quarter * 32 + lane;
    const bool keep_l2 = (p.debug & 8) == 0;      // L2 evict_last on the output (2688 -> 2678 us per timestep); debug bit 3 = off
    const uint64_t l2pol = l2_evict_last_policy();
    const int H = p.H, W = p.W, P = p.P, hw = p.H * p.W;
    const int tiles_per_image = p.tiles_per_image, num_tiles = p.num_tiles;
    const int ldy = p.ldy, ldres = p.ldres, ld_rowvec = p.ld_rowvec, fin_cout = p.fin_cout;
    bf16* const y = p.y;
    const bf16* const res = p.res;
    const float* const rowvec = p.rowvec;
    const float* const fin_w = p.fin_w;
    float* const fin_out = p.fin_out;
    const bool no_mem = p.debug & 1;
    const int res_mod = p.res_mod;
    float* s_fin = s_bias + 768;  // [128][8] cross-warp sums of the fused projection (offset 3 KB in the 7 KB region)
    // the bias vector is read by every row of every tile: stage it in shared memory once (the CTA runs with the
    // maximum shared-memory carve-out, so L1 is tiny and a __ldg per tile would pay L2 latency each time)
    {
      const int et = threadIdx.x - 64;
      for (int c = et; c < BLOCK_N; c += 256) s_bias[c] = p.bias ? p.bias[c] : 0.f;
      if (fin_out)  // fused output projection: [fin_cout][BLOCK_N] weights behind the bias vector
        for (int c = et; c < fin_cout * BLOCK_N; c += 256) s_bias[BLOCK_N + c] = fin_w[c];
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    int iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
      const int n = tile / tiles_per_image;
      const int q0 = P + (tile - n * tiles_per_image) * TILE_M;
      const int acc = iter & 1;
      const int q = q0 + row;
      const int r = q / P, c = q - r * P;
      const bool valid = r >= 1 && r <= H && c >= 1 && c <= W && !no_mem;  // a real pixel, not a halo column / tail row
      const int pix = (r - 1) * W + (c - 1);
      const int64_t m = (int64_t)n * hw + pix;
      bf16* yrow = y ? y + m * ldy + cw : nullptr;
      const float* rvrow = rowvec ? rowvec + (int64_t)n * ld_rowvec + cw : nullptr;
      // residual row: requested BEFORE waiting for the accumulator, so its latency hides behind the MMAs
      uint4 rr[COLS / 8];
      if (res && valid) {
        const int64_t mr = res_mod > 0 ? (int64_t)(n % res_mod) * hw + pix : m;
        const uint4* rp = reinterpret_cast<const uint4*>(res + mr * ldres + cw);
#pragma unroll
        for (int j = 0; j < COLS / 8; ++j) rr[j] = __ldg(rp + j);
      }
      if (warp == 2 && lane == 0) DBG_STAMP(4);
      if (lane == 0) mbar_wait(tfull + 8 * acc, (iter >> 1) & 1);
      __syncwarp();
      if (warp == 2 && lane == 0) DBG_STAMP(5);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + cw;
      float fo[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c0 = 0; c0 < COLS; c0 += 32) {
        uint32_t rg[32];
        tmem_ld32(taddr + c0, rg);
        tmem_ld_wait();
        if (valid) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + cw + c0 + j);
            v[j] = __uint_as_float(rg[j]) + b4.x; v[j + 1] = __uint_as_float(rg[j + 1]) + b4.y;
            v[j + 2] = __uint_as_float(rg[j + 2]) + b4.z; v[j + 3] = __uint_as_float(rg[j + 3]) + b4.w;
          }
          if (rvrow) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 b4 = __ldg(reinterpret_cast<const float4*>(rvrow + c0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
          if (res) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rr[(c0 + j) >> 3]);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float2 f = __bfloat1622float2(h2[u]);
                v[j + 2 * u] += f.x; v[j + 2 * u + 1] += f.y;
              }
            }
          }
          if (yrow) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float t8[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) t8[u] = v[j + u];
              if (keep_l2) store_chunk_keep(yrow + c0 + j, t8, l2pol); else store_chunk(yrow + c0 + j, t8);
            }
          }
          if (fin_out) {
            for (int o = 0; o < fin_cout; ++o) {
              const float* wrow = s_bias + BLOCK_N + o * BLOCK_N + cw + c0;
              float sacc = 0.f;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 w4 = *reinterpret_cast<const float4*>(wrow + j);
                sacc = fmaf(v[j], w4.x, sacc); sacc = fmaf(v[j + 1], w4.y, sacc);
                sacc = fmaf(v[j + 2], w4.z, sacc); sacc = fmaf(v[j + 3], w4.w, sacc);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (u == o) fo[u] += sacc;
            }
          }
        }
      }
      if (fin_out) {
        // the two column halves of a row live in two warps: combine through shared memory, half 0 writes
        float* fx = s_fin + row * 8;
        if (half == 1) {
#pragma unroll
          for (int u = 0; u < 8; ++u) fx[u] = fo[u];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 0 && valid) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (u < fin_cout) fin_out[((int64_t)n * fin_cout + u) * hw + pix] = fo[u] + fx[u] + __ldg(p.fin_b + u);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // fx is rewritten by the next tile
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(tempty + 8 * acc);
      if (warp == 2 && lane == 0) DBG_STAMP(6);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * BLOCK_N);
  (void)PL_rows;
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode_h = nullptr;
int g_num_sms_h = 0;
int g_halo_enabled = -1;
int g_base_offset_mode = 0;

int halo_init() {
  if (g_encode_h) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  LDM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  LDM_REQUIRE(qres == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  int dev = 0;
  LDM_CUDA(cudaGetDevice(&dev));
  LDM_CUDA(cudaDeviceGetAttribute(&g_num_sms_h, cudaDevAttrMultiProcessorCount, dev));
  LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const char* e = getenv("LDM_CONV_HALO");
  g_halo_enabled = e ? atoi(e) : 1;
  const char* bo = getenv("LDM_HALO_BASE_OFFSET");
  g_base_offset_mode = bo ? atoi(bo) : 0;
  g_encode_h = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}

int make_slab_map(CUtensorMap* map, const void* x, int ld, int cin, int B, int H, int W, int P, int RB) {
  cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)P, (cuuint32_t)RB, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const char* ex = getenv("LDM_HALO_EXP");
  const int exp_bits = ex ? atoi(ex) : 0;
  if (exp_bits & 1) box[1] = (cuuint32_t)W;
  if (exp_bits & 2) box[2] = 4;
  CUresult r = g_encode_h(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          (exp_bits & 4) ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : ((exp_bits & 8) ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B),
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(halo slab) failed with CUresult %d", (int)r);
  return 0;
}
int make_filter_map(CUtensorMap* map, const void* w, int cout, int ktot, int block_n) {
  cuuint64_t gdim[2] = {(cuuint64_t)ktot, (cuuint64_t)cout};
  cuuint64_t gstr[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode_h(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(filter) failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace

int k_conv_halo_prepare() { return halo_init(); }

// Whether the halo kernel takes this convolution (3x3, W in {32,16}... see below); otherwise conv_tc_kernel runs it.
bool k_conv_halo_applicable(const ConvArgs& a) {
  if (g_halo_enabled == 0) return false;
  if (a.dtype != LDM_DT_BF16 || a.ksize != 3 || a.up2) return false;
  if (a.cout != 64 && a.cout != 128) return false;
  if (a.cin % BLOCK_K != 0 || (a.x2 && a.cin2 % BLOCK_K != 0)) return false;
  if (a.height < 4) return false;
  if (a.width == 32) return true;                    // full-resolution layers (W/P = 94 % useful rows)
  // 16x16: 75 % of the MMA rows are useful (3 tiles per 288-position plane); LDM_HALO_W16: 1 = all, 2 = resident filters only
  static const int w16 = getenv("LDM_HALO_W16") ? atoi(getenv("LDM_HALO_W16")) : 0;
  if (a.width == 16 && w16 == 1) return true;
  if (a.width == 16 && w16 == 2) {
    const int n_kb = 9 * (a.cin / BLOCK_K) + (a.x2 ? a.cin2 / BLOCK_K : 0);
    return (int64_t)n_kb * a.cout * BLOCK_K * 2 <= 120 * 1024;
  }
  return false;
}

int k_conv_halo(const ConvArgs& a, cudaStream_t st) {
  if (int rc = halo_init()) return rc;
  LDM_REQUIRE(k_conv_halo_applicable(a), "conv_halo: unsupported convolution");
  LDM_REQUIRE(a.ldx % 8 == 0 && (!a.y || a.ldy % 8 == 0) && (!a.x2 || a.ldx2 % 8 == 0) && (!a.res || a.ldres % 8 == 0),
              "conv_halo: pixel strides must be multiples of 8 elements");
  LDM_REQUIRE(((uintptr_t)a.x & 15) == 0 && ((uintptr_t)a.w & 15) == 0 && ((uintptr_t)a.y & 15) == 0 &&
                  ((uintptr_t)a.fin_w & 15) == 0, "conv_halo: pointers must be 16-byte aligned");
  LDM_REQUIRE(a.y || a.fin_out, "conv_halo: no output requested");
  LDM_REQUIRE(!a.fin_out || (a.fin_cout >= 1 && (a.fin_cout + 1) * a.cout * 4 <= 3072 && a.fin_cout <= 8),
              "conv_halo: fused projection supports at most %d outputs", 3072 / (a.cout * 4) - 1);
  if (a.batch == 0) return 0;
  HaloParams p;
  p.B = a.batch; p.H = a.height; p.W = a.width; p.P = a.width + 2;
  const int span = TILE_M + 2 * (p.P + 1);                 // positions a tile touches over all taps
  p.RB = (span + p.P - 1) / p.P + 1;                       // rows covering any alignment of that span
  if (p.RB > p.H + 2) p.RB = p.H + 2 > 0 ? p.RB : p.RB;    // (box may exceed the image: OOB rows are zero)
  p.tiles_per_image = (a.height * p.P + TILE_M - 1) / TILE_M;
  p.num_tiles = a.batch * p.tiles_per_image;
  p.slabs_main = a.cin / BLOCK_K;
  p.slabs_total = p.slabs_main + (a.x2 ? a.cin2 / BLOCK_K : 0);
  p.n_kb = 9 * p.slabs_main + (p.slabs_total - p.slabs_main);
  p.a_stage_bytes = (p.RB * p.P * 128 + 1023) / 1024 * 1024;
  p.cout = a.cout;
  p.base_offset_mode = g_base_offset_mode;
  { const char* d = getenv("LDM_HALO_DEBUG"); p.debug = d ? atoi(d) : 0; }
  p.a_tx_bytes = p.RB * p.P * 128; p.w_start = -1;
  {
    const char* ex = getenv("LDM_HALO_EXP");
    const int eb = ex ? atoi(ex) : 0;
    const int bw = (eb & 1) ? p.W : p.P, bh = (eb & 2) ? 4 : p.RB;
    if (eb & 1) p.w_start = 0;
    p.a_tx_bytes = bw * bh * 128;
  }
  p.bias = a.bias; p.rowvec = a.rowvec; p.ld_rowvec = a.ld_rowvec;
  p.res = (const bf16*)a.res; p.ldres = a.ldres; p.res_mod = a.res_mod;
  p.y = (bf16*)a.y; p.ldy = a.ldy;
  p.fin_w = a.fin_w; p.fin_b = a.fin_b; p.fin_out = a.fin_out; p.fin_cout = a.fin_cout;
  // shared-memory plan: filter resident if it leaves room for >= 2 slabs, else a streaming ring of 8 tiles
  const int b_tile = a.cout * BLOCK_K * 2;
  // barriers (1 KB) + bias / projection vectors (3 KB) + projection cross-warp sums (4 KB) + alignment slack
  const int budget = 227 * 1024 - 9216;
  if ((int64_t)p.n_kb * b_tile + 2 * p.a_stage_bytes <= budget && p.n_kb <= 32) p.b_stages = p.n_kb;
  else p.b_stages = 8;
  p.a_stages = (budget - p.b_stages * b_tile) / p.a_stage_bytes;
  if (p.a_stages > 8) p.a_stages = 8;
  LDM_REQUIRE(p.a_stages >= 2, "conv_halo: shared memory plan failed (cout %d, %d k-blocks)", a.cout, p.n_kb);
  const int smem = p.b_stages * b_tile + p.a_stages * p.a_stage_bytes + 8192 + 1024;
  CUtensorMap ma, ma2, mb;
  if (int rc = make_slab_map(&ma, a.x, a.ldx, a.cin, a.batch, a.height, a.width, p.P, p.RB)) return rc;
  if (a.x2) {
    if (int rc = make_slab_map(&ma2, a.x2, a.ldx2, a.cin2, a.batch, a.height, a.width, p.P, p.RB)) return rc;
  } else {
    ma2 = ma;
  }
  if (int rc = make_filter_map(&mb, a.w, a.cout, p.n_kb * BLOCK_K, a.cout)) return rc;
  const int grid = p.num_tiles < g_num_sms_h ? p.num_tiles : g_num_sms_h;
  if (a.cout == 64) LDM_CUDA(ldm_launch_pdl(conv_halo_kernel<64>, dim3(grid), dim3(320), (size_t)smem, st, ma, ma2, mb, p));
  else LDM_CUDA(ldm_launch_pdl(conv_halo_kernel<128>, dim3(grid), dim3(320), (size_t)smem, st, ma, ma2, mb, p));
  LDM_LAUNCHED("conv_halo");
  return 0;
}
