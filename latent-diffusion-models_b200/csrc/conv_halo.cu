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
// Roles (384 threads, persistent, 1 CTA/SM): warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue, warps 10..11
// operand transform (idle unless the source is normalised on the fly); TMEM holds an accumulator ring so the epilogue of
// tile i overlaps the MMAs of tile i+1.
//   * on-the-fly GroupNorm (+SiLU) of the main source (src/UNet.py:52-58: Block = conv(act(norm(x)))): the conv reads the RAW
//     tensor and, between a slab's TMA landing and its MMAs, the two transform warps rewrite the slab in place as
//     act(a x + b) with per-(image, channel) {a, b} (k_group_norm_coef).  Rows and columns of the zero padding are left
//     untouched, so the padding is the normalised tensor's.  A lane always meets the same 8 channels (its 16-byte chunk
//     index XOR the swizzle phase of its pixel rows, which advance by 8), so the coefficients sit in registers.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "kernels.h"
#include "tc_common.cuh"
#include "conv_epilogue.cuh"

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
  int P;                   // padded pitch W+2
  int tiles_per_image, num_tiles;
  int slabs_main;          // cin/64
  int slabs_total;         // + cin2/64 (centre tap only)
  int n_kb;                // 9*slabs_main + slabs2
  int a_stage_bytes;       // RB*P*128 rounded up to 1024
  int a_stages, b_stages;  // ring depths; b_stages >= n_kb means the filter is resident
  int a_tx_bytes, w_start; // TMA box bytes, first W coordinate (-1)
  int debug;               // profiling knob (LDM_HALO_DEBUG bit 0: no epilogue memory traffic, bit 1: no MMAs)
  int base_offset_mode;    // experiment knob: 0 = descriptor base_offset 0, 1 = (start >> 7) & 7
  // on-the-fly GroupNorm (+SiLU) of the main source (ConvArgs::xf_ab): table [images][cin] of {a, b}
  const float2* xf_ab; int xf_silu; int xf_cin; int x_mod; int H, RB;
  EpiP e;                  // everything the epilogue warps need (conv_epilogue.cuh)
};

constexpr int HALO_NACC = 4;   // accumulator ring depth (the fused GroupNorm defers its second pass by one tile)

constexpr int HALO_THREADS = 384;

template <int BLOCK_N, int GM = 0>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                 const __grid_constant__ CUtensorMap tmap_b, const HaloParams p) {
  constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_b = smem_base;
  const uint32_t smem_a = smem_b + p.b_stages * B_TILE_BYTES;
  const uint32_t bars = smem_a + p.a_stages * p.a_stage_bytes;
  // barrier map: afull[8] aempty[8] bfull[32] bempty[32] tfull[4] tempty[4] | tmem slot | xfull[8]
  const uint32_t afull = bars, aempty = bars + 64, bfull = bars + 128, bempty = bars + 384, tfull = bars + 640,
                 tempty = bars + 672, tmem_slot = bars + 704, xfull = bars + 768;
  float* s_epi = reinterpret_cast<float*>(smem_raw + (bars + 1024 - smem_u32(smem_raw)));   // epilogue staging area
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_a2);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < 8; ++s) { mbar_init(afull + 8 * s, 1); mbar_init(aempty + 8 * s, 1); mbar_init(xfull + 8 * s, 2); }
    for (int s = 0; s < 32; ++s) { mbar_init(bfull + 8 * s, 1); mbar_init(bempty + 8 * s, 1); }
    for (int i = 0; i < HALO_NACC; ++i) { mbar_init(tfull + 8 * i, 1); mbar_init(tempty + 8 * i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, HALO_NACC * BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();      // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const bool resident = p.b_stages >= p.n_kb;
  const int slabs2 = p.slabs_total - p.slabs_main;
  (void)slabs2;

  // tile -> image n = tile / tiles_per_image; q0 = P + (tile % tiles_per_image) * 128 is the plane position of the
  // tile's output row 0; the slab starts at padded row r0 = floor((q0 - P - 1) / P) (the lowest row any tap touches).

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int P = p.P, a_stages = p.a_stages, b_stages = p.b_stages, a_stage_bytes = p.a_stage_bytes;
      const int slabs_main = p.slabs_main, slabs_total = p.slabs_total, num_tiles = p.num_tiles;
      const int tiles_per_image = p.tiles_per_image, w_start = p.w_start, x_mod = p.x_mod;
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
                      r0 - 1, (main_src && x_mod > 0) ? n % x_mod : n);
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
      const bool xform = p.xf_ab != nullptr;
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
      int astage = 0;
      int bstage = 0; uint32_t bphase = 0;
      int iter = 0;
      int tin = blockIdx.x % tiles_per_image;  // tile index inside its image, advanced incrementally
      const int tstep = gridDim.x % tiles_per_image;
      // The waits run ONE SLAB AHEAD of the MMAs: between two bursts the thread used to spend ~0.5 us (commit, index arithmetic,
      // two satisfied mbarrier waits, fences) during which the tensor pipe sat idle -- a quarter of the tile period.  Now the
      // accumulator / slab barriers of the NEXT slab (of this tile or of the next tile) are consumed in the middle of the current
      // slab's taps, while the already queued MMAs keep the pipe busy.
      int r_tile = blockIdx.x, r_s = 0, r_iter = 0, r_stage = 0;
      uint32_t r_phase = 0;
      auto ready_next = [&]() {
        if (r_tile >= num_tiles) return;
        if (r_s == 0) mbar_wait(tempty + 8 * (r_iter % HALO_NACC), ((r_iter / HALO_NACC) & 1) ^ 1);
        mbar_wait(((xform && r_s < slabs_main) ? xfull : afull) + 8 * r_stage, r_phase);
        tc_fence_after();
        if (++r_stage == a_stages) { r_stage = 0; r_phase ^= 1; }
        if (++r_s == slabs_total) { r_s = 0; r_tile += gridDim.x; ++r_iter; }
      };
      const bool ahead = !(p.debug & 64);      // LDM_HALO_DEBUG bit 6: wait at the top of each slab instead (A/B)
      if (ahead) ready_next();
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
        const int q0 = P + tin * TILE_M;
        const int lo = q0 - P - 1;
        const int r0 = lo >= 0 ? lo / P : -1;
        const int off0 = q0 - r0 * P;  // slab row of the tile's first output position (centre tap)
        tin += tstep; if (tin >= tiles_per_image) tin -= tiles_per_image;
        const int acc = iter % HALO_NACC;
        DBG_STAMP(2);
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        uint32_t first_mma = 0u;  // accumulate flag of the very first MMA of the tile
        int kb = 0;
        for (int s = 0; s < slabs_total; ++s) {
          if (!ahead) ready_next();
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
                if (t == 5 && ahead) ready_next();
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
                if (t == 5 && ahead) ready_next();
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
            if (ahead) ready_next();
          }
          first_mma = 1u;
          umma_commit(aempty + 8 * astage);  // the slab is free once its taps' MMAs retire
          if (++astage == a_stages) astage = 0;
        }
        umma_commit(tfull + 8 * acc);
        DBG_STAMP(3);
      }
    }
  } else if (warp < 10) {
    // ===================== epilogue (warps 2..9): conv_epilogue.cuh =====================
    conv_epilogue<BLOCK_N, 1, HALO_NACC, true, GM>(p.e, tmem_base, tfull, tempty, s_epi, p.num_tiles);
  } else if (p.xf_ab != nullptr) {
    // ===================== operand transform (warps 10..11): slab <- act(a x + b), padding untouched =====================
    const int P = p.P, a_stages = p.a_stages, a_stage_bytes = p.a_stage_bytes;
    const int slabs_main = p.slabs_main, slabs_total = p.slabs_total, num_tiles = p.num_tiles;
    const int tiles_per_image = p.tiles_per_image, H = p.H, W = P - 2, npix = p.RB * P, cin = p.xf_cin;
    const bool silu = p.xf_silu != 0;
    const int t2 = threadIdx.x - 320;              // 0..63
    const int chunk = t2 & 7, pl = t2 >> 3;        // physical 16-byte chunk of the pixel row; pixel rows pl, pl + 8, ...
    const int lch = (chunk ^ pl) * 8;              // logical channels of that chunk: the swizzle phase of row i is i & 7 == pl
    int astage = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n = tile / tiles_per_image;
      const int q0 = P + (tile - n * tiles_per_image) * TILE_M;
      const int lo = q0 - P - 1;
      const int r0 = lo >= 0 ? lo / P : -1;
      for (int s = 0; s < slabs_total; ++s) {
        if (s < slabs_main) {
          float2 ab[8];                            // issued before the wait: the table reads overlap the slab's TMA
          const float2* abp = p.xf_ab + (int64_t)n * cin + s * BLOCK_K + lch;
#pragma unroll
          for (int i = 0; i < 8; ++i) ab[i] = __ldg(abp + i);
          mbar_wait(afull + 8 * astage, aphase);
          const uint32_t base = smem_a + astage * a_stage_bytes + chunk * 16;
          // pixel i = row * P + col of the box = image row r0 - 1 + row, column col - 1.  Four pixel rows (i, i + 8, i + 16,
          // i + 24) per step: their loads are issued together and the four dependent chains interleave.
          int row = 0, col = pl;
          for (int i0 = pl; i0 < ((p.debug & 8) ? 0 : npix); i0 += 32) {
            uint32_t wv[4][4];
            bool ok[4];
            int rr = row, cc = col;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int hh = r0 - 1 + rr;
              ok[u] = i0 + 8 * u < npix && (unsigned)hh < (unsigned)H && cc >= 1 && cc <= W;
              if (ok[u])
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(wv[u][0]), "=r"(wv[u][1]), "=r"(wv[u][2]), "=r"(wv[u][3]) : "r"(base + (i0 + 8 * u) * 128));
              cc += 8;
              while (cc >= P) { cc -= P; ++rr; }
            }
            row = rr; col = cc;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (!ok[u]) continue;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float v0 = fmaf(__uint_as_float(wv[u][k] << 16), ab[2 * k].x, ab[2 * k].y);
                float v1 = fmaf(__uint_as_float(wv[u][k] & 0xffff0000u), ab[2 * k + 1].x, ab[2 * k + 1].y);
                if (silu && !(p.debug & 16)) {     // {a, b} are pre-halved: silu(v) = v/2 (1 + tanh(v/2))
                  float t0, t1;
                  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(v0));
                  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(v1));
                  v0 = fmaf(v0, t0, v0);
                  v1 = fmaf(v1, t1, v1);
                }
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
                wv[u][k] = *reinterpret_cast<uint32_t*>(&h2);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(base + (i0 + 8 * u) * 128), "r"(wv[u][0]), "r"(wv[u][1]),
                           "r"(wv[u][2]), "r"(wv[u][3]) : "memory");
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> the MMA's async proxy
          __syncwarp();
          if (lane == 0) mbar_arrive(xfull + 8 * astage);
        }
        if (++astage == a_stages) { astage = 0; aphase ^= 1; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, HALO_NACC * BLOCK_N);
}

PFN_cuTensorMapEncodeTiled_v12000 g_encode_h = nullptr;
int g_num_sms_h = 0;
int g_halo_enabled = -1;
int g_base_offset_mode = 0;

int halo_init() {
  static bool inited[64] = {};   // function attributes are per device
  int dev = 0;
  LDM_CUDA(cudaGetDevice(&dev));
  if (g_encode_h && inited[dev & 63]) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  LDM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  LDM_REQUIRE(qres == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  LDM_CUDA(cudaDeviceGetAttribute(&g_num_sms_h, cudaDevAttrMultiProcessorCount, dev));
  LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<64, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<128, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  LDM_CUDA(cudaFuncSetAttribute(conv_halo_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const char* e = getenv("LDM_CONV_HALO");
  g_halo_enabled = e ? atoi(e) : 1;
  const char* bo = getenv("LDM_HALO_BASE_OFFSET");
  g_base_offset_mode = bo ? atoi(bo) : 0;
  g_encode_h = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  inited[dev & 63] = true;
  return 0;
}

inline bool P_ok_for_xform(int P) { return P >= 8; }   // the transform's row/column walk advances 8 pixels per step

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
  LDM_REQUIRE(!a.fin_out || (a.fin_cout >= 1 && a.fin_cout * a.cout <= 768 && a.fin_cout <= 8),
              "conv_halo: fused projection supports at most %d outputs", 768 / a.cout);
  if (a.batch == 0) return 0;
  HaloParams p;
  EpiP& e = p.e;
  const int H = a.height, W = a.width;
  p.P = W + 2;
  const int span = TILE_M + 2 * (p.P + 1);                 // positions a tile touches over all taps
  const int RB = (span + p.P - 1) / p.P + 1;               // rows covering any alignment of that span (OOB rows are zero)
  p.tiles_per_image = (H * p.P + TILE_M - 1) / TILE_M;
  p.num_tiles = a.batch * p.tiles_per_image;
  p.slabs_main = a.cin / BLOCK_K;
  p.slabs_total = p.slabs_main + (a.x2 ? a.cin2 / BLOCK_K : 0);
  p.n_kb = 9 * p.slabs_main + (p.slabs_total - p.slabs_main);
  p.a_stage_bytes = (RB * p.P * 128 + 1023) / 1024 * 1024;
  p.base_offset_mode = g_base_offset_mode;
  p.xf_ab = (const float2*)a.xf_ab; p.xf_silu = a.xf_silu; p.xf_cin = a.cin; p.x_mod = a.x_mod; p.H = H; p.RB = RB;
  LDM_REQUIRE(!a.xf_ab || P_ok_for_xform(p.P), "conv_halo: the operand transform needs P <= 8 * 32");
  LDM_REQUIRE(a.x_mod == 0 || a.batch % a.x_mod == 0, "conv_halo: x_mod must divide the batch");
  { const char* d = getenv("LDM_HALO_DEBUG"); p.debug = d ? atoi(d) : 0; }
  p.a_tx_bytes = RB * p.P * 128; p.w_start = -1;
  {
    const char* ex = getenv("LDM_HALO_EXP");
    const int eb = ex ? atoi(ex) : 0;
    const int bw = (eb & 1) ? W : p.P, bh = (eb & 2) ? 4 : RB;
    if (eb & 1) p.w_start = 0;
    p.a_tx_bytes = bw * bh * 128;
  }
  e.M = a.batch * H * W; e.H = H; e.W = W; e.hw = H * W; e.P = p.P; e.tiles_per_image = p.tiles_per_image;
  e.num_m_tiles = p.num_tiles; e.num_n_tiles = 1; e.batch = a.batch;
  e.up2 = 0; e.cout = a.cout; e.cout_real = a.cout;
  e.bias = a.bias; e.rowvec = a.rowvec; e.ld_rowvec = a.ld_rowvec;
  e.res = (const bf16*)a.res; e.ldres = a.ldres; e.res_mod = a.res_mod;
  e.y = (bf16*)a.y; e.ldy = a.ldy;
  e.fin_w = a.fin_w; e.fin_b = a.fin_b; e.fin_out = a.fin_out; e.fin_cout = a.fin_cout;
  e.debug = p.debug | ((getenv("LDM_EPI_DEBUG") ? atoi(getenv("LDM_EPI_DEBUG")) : 0) << 16);
  e.dbg = nullptr;
  if (p.debug & 4) LDM_CUDA(cudaGetSymbolAddress((void**)&e.dbg, g_halo_dbg));
  // ---- fused GroupNorm plan (conv_epilogue.cuh): a sample always spans tiles_per_image tiles -> packet exchange
  const ConvGn& g = a.gn;
  e.gn_mode = g.mode;
  e.gn_G = 1; e.gn_cpg = a.cout; e.gn_silu = 0; e.gn_nvar = 1; e.gn_var_rows = 0; e.gn_nslots = 1; e.gn_cross = 0; e.gn_upt = 1;
  e.gn_eps = g.eps; e.gn_tag = g.tag; e.gn_gamma = g.gamma; e.gn_beta = g.beta; e.gn_rowvec = g.rowvec; e.gn_ld_rowvec = g.ld_rowvec;
  e.gn_res = (const bf16*)g.res; e.gn_ldres = g.ldres; e.gn_scratch = g.scratch;
  if (g.mode) {
    LDM_REQUIRE(g.mode == 1 || g.mode == 2, "conv_halo: GroupNorm epilogue mode %d unknown", g.mode);
    LDM_REQUIRE(!a.fin_out && a.y, "conv_halo: the GroupNorm epilogue needs a plain NHWC output");
    LDM_REQUIRE(g.groups >= 1 && a.cout % g.groups == 0 && (a.cout / g.groups) % 8 == 0, "conv_halo: GroupNorm groups %d vs Cout %d", g.groups, a.cout);
    LDM_REQUIRE(g.nvar == 1 || (g.nvar == 2 && g.var_rows >= a.batch && g.rowvec), "conv_halo: GroupNorm variants need a row vector and var_rows >= batch");
    LDM_REQUIRE(g.mode == 1 || (g.gamma && g.beta && !a.res && !a.rowvec), "conv_halo: GroupNorm mode 2 takes gamma/beta and no pre-norm residual / row vector");
    LDM_REQUIRE(!g.res || g.ldres % 8 == 0, "conv_halo: GroupNorm residual stride must be a multiple of 8");
    e.gn_G = g.groups; e.gn_cpg = a.cout / g.groups; e.gn_silu = g.silu; e.gn_nvar = g.nvar; e.gn_var_rows = g.var_rows;
    LDM_REQUIRE(g.mode == 1 || a.cout <= EPI_FULL_VEC, "conv_halo: GroupNorm mode 2 supports at most %d channels", EPI_FULL_VEC);
    e.gn_upt = p.tiles_per_image; e.gn_cross = (g.mode == 2 && p.tiles_per_image > 1) ? 1 : 0;
    e.gn_nslots = p.tiles_per_image * (g.groups == 1 ? a.cout / 64 : 1);
    if (g.mode == 1)   // warp-local partial sums: (32-position block of the padded plane) x (group, or 32-column chunk)
      e.gn_nslots = p.tiles_per_image * 4 * (g.groups == 1 ? a.cout / 32 : (a.cout / g.groups > 32 ? a.cout / g.groups / 32 : 1));
    LDM_REQUIRE(g.mode != 2 || g.groups * g.nvar * e.gn_nslots <= 256, "conv_halo: GroupNorm packet fan-in too large");
    const int64_t need = (int64_t)a.batch * g.groups * g.nvar * e.gn_nslots * (g.mode == 1 ? 8 : 16);
    LDM_REQUIRE(g.mode == 2 && !e.gn_cross ? true : (g.scratch && g.scratch_bytes >= need && ((uintptr_t)g.scratch & 15) == 0),
                "conv_halo: GroupNorm scratch missing / too small (need %lld bytes)", (long long)need);
    LDM_REQUIRE(g.mode != 2 || !e.gn_cross || g.tag != 0, "conv_halo: GroupNorm packets need a non-zero tag");
    if (g.nslots_out) *g.nslots_out = e.gn_nslots;
  }
  // shared-memory plan: filter resident if it leaves room for >= 2 slabs, else a streaming ring of 8 tiles
  const int b_tile = a.cout * BLOCK_K * 2;
  // barriers (1 KB) + the epilogue's staging area + alignment slack
  const int fixed = 1024 + epi_smem_bytes(a.cout) + 1024;
  const int budget = 227 * 1024 - fixed;
  if ((int64_t)p.n_kb * b_tile + 2 * p.a_stage_bytes <= budget && p.n_kb <= 32) p.b_stages = p.n_kb;
  else p.b_stages = 8;
  p.a_stages = (budget - p.b_stages * b_tile) / p.a_stage_bytes;
  if (p.a_stages > 8) p.a_stages = 8;
  LDM_REQUIRE(p.a_stages >= 2, "conv_halo: shared memory plan failed (cout %d, %d k-blocks)", a.cout, p.n_kb);
  const int smem = p.b_stages * b_tile + p.a_stages * p.a_stage_bytes + fixed;
  CUtensorMap ma, ma2, mb;
  if (int rc = make_slab_map(&ma, a.x, a.ldx, a.cin, a.x_mod > 0 ? a.x_mod : a.batch, a.height, a.width, p.P, RB)) return rc;
  if (a.x2) {
    if (int rc = make_slab_map(&ma2, a.x2, a.ldx2, a.cin2, a.batch, a.height, a.width, p.P, RB)) return rc;
  } else {
    ma2 = ma;
  }
  if (int rc = make_filter_map(&mb, a.w, a.cout, p.n_kb * BLOCK_K, a.cout)) return rc;
  const int grid = p.num_tiles < g_num_sms_h ? p.num_tiles : g_num_sms_h;
#define HALO_GO(BN, GMV) LDM_CUDA(ldm_launch_pdl(conv_halo_kernel<BN, GMV>, dim3(grid), dim3(HALO_THREADS), (size_t)smem, st, ma, ma2, mb, p))
  if (a.cout == 64) {
    if (e.gn_mode == 0) HALO_GO(64, 0); else if (e.gn_mode == 1) HALO_GO(64, 1); else HALO_GO(64, 2);
  } else {
    if (e.gn_mode == 0) HALO_GO(128, 0); else if (e.gn_mode == 1) HALO_GO(128, 1); else HALO_GO(128, 2);
  }
#undef HALO_GO
  LDM_LAUNCHED("conv_halo");
  return 0;
}
