// tcgen05 implicit-GEMM convolution for sm_100a: bf16 NHWC activations, bf16 K-major packed weights,
// fp32 accumulation in TMEM.  Replaces F.conv2d (3x3 pad 1 and 1x1) and F.conv_transpose2d (k2 s2, as a
// 1x1 GEMM with a pixel-shuffle scatter epilogue) behind src/UNet.py:54,82,99,119-120,145-147,231-233.
//
//   C[M = B*H*W, N = Cout] = A_im2col[M, K] * W[N, K]^T,   K = taps*Cin (+ Cin2: fused 1x1 shortcut source)
//
// Data movement
//   * an M-tile is 128 consecutive NHWC pixels = a rectangular (n, h, w) box, so the im2col operand of tap
//     (dy,dx) is ONE 4-D TMA tiled load of box {64 ch, W, HB, NB} at coordinates {c0, dx, h0+dy, n0}:
//     the halo (and the batch tail) is zero-filled by TMA's out-of-bounds handling -- no im2col buffer,
//     no index arithmetic in the SM.  The box lands as 128 rows x 128 B with the 128-byte swizzle, which is
//     exactly the canonical K-major SWIZZLE_128B UMMA operand layout.
//   * the weight tile [BLOCK_N x 64] is a 2-D TMA load from the packed [Cout][K] matrix, same layout.
// Execution (one persistent CTA per SM, 192 threads)
//   warp 0      TMA producer (one lane): STAGES-deep mbarrier ring
//   warp 1      TMEM allocator + tcgen05.mma issuer (one lane): 4 x (128 x BLOCK_N x 16) MMAs per k-block,
//               tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> +bias +time-embedding row vector
//               +residual -> bf16 -> 16-byte global stores.  Two accumulator buffers in TMEM (2*BLOCK_N
//               columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <initializer_list>

#include "kernels.h"

namespace {

constexpr int TILE_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = TILE_M * BLOCK_K * 2;

struct TcParams {
  int M;          // valid rows
  int H, W;       // spatial size
  int HB, NB;     // M-tile box: NB images x HB rows x W columns = 128 pixels
  int num_m_tiles, num_n_tiles;
  int kb_main;    // k-blocks from the main source = taps * cin/64
  int kb_total;   // + cin2/64
  int cin_blocks; // cin/64
  int taps;       // 1 or 9
  int cout;       // GEMM N
  int cout_real;  // channels of the output tensor (cout/4 for up2)
  int up2;
  const float* bias;
  const float* rowvec; int ld_rowvec;
  const bf16* res; int ldres;
  bf16* y; int ldy;   // may be null when only the fused projection output is wanted
  // fused trailing 1x1 projection (the UNet's final_conv.1, src/UNet.py:347): out[b][o][pix] = fin_b[o] +
  // sum_c fin_w[o][c] * row[c], computed from the fp32 accumulators; requires one N-tile (BLOCK_N == cout)
  const float* fin_w; const float* fin_b; float* fin_out; int fin_cout;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must surface as a trapped kernel (launch error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long start = clock64();
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - start > 4000000000LL) __trap();  // ~2 s at 2 GHz
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"): rows are 128 B apart,
// 8-row core-matrix groups are SBO = 1024 B apart; LBO is unused for a single swizzle atom along K.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BLOCK_N
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

template <int BLOCK_N>
struct TcCfg {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = BLOCK_N == 64 ? 8 : (BLOCK_N == 128 ? 6 : 4);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // double-buffered accumulator (power of two >= 32)
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + alignment slack
};

template <int BLOCK_N>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
               const __grid_constant__ CUtensorMap tmap_b, const TcParams p) {
  using Cfg = TcCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B needs 1024-B alignment
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t bars = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
  const uint32_t tfull_bar = bars + 16 * STAGES, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_slot = tempty_bar + 16;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_a2);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);
      mbar_init(tempty_bar + 8 * i, 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
        int n0, h0;
        if (p.NB == 1) {
          const int tpi = p.H / p.HB;
          n0 = mt / tpi;
          h0 = (mt % tpi) * p.HB;
        } else {
          n0 = mt * p.NB;
          h0 = 0;
        }
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          mbar_expect_tx(full_bar + 8 * stage, Cfg::STAGE_BYTES);
          if (kb < p.kb_main) {
            const int tap = kb / p.cin_blocks, cb = kb % p.cin_blocks;
            int dy = 0, dx = 0;
            if (p.taps == 9) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
            tma_load_4d(smem_a + stage * A_STAGE_BYTES, &tmap_a, full_bar + 8 * stage, cb * BLOCK_K, dx, h0 + dy, n0);
          } else {
            tma_load_4d(smem_a + stage * A_STAGE_BYTES, &tmap_a2, full_bar + 8 * stage, (kb - p.kb_main) * BLOCK_K,
                        0, h0, n0);
          }
          tma_load_2d(smem_b + stage * Cfg::B_STAGE_BYTES, &tmap_b, full_bar + 8 * stage, kb * BLOCK_K, nt * BLOCK_N);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
        const int acc = iter & 1;
        mbar_wait(tempty_bar + 8 * acc, ((iter >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
          const uint64_t adesc = make_sw128_desc(smem_a + stage * A_STAGE_BYTES);
          const uint64_t bdesc = make_sw128_desc(smem_b + stage * Cfg::B_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the (addr>>4) field
            umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar + 8 * stage);  // frees the smem stage when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar + 8 * acc);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, +32) are the ones this warp may read
    const int row = quarter * 32 + lane;
    int iter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
      const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
      const int acc = iter & 1;
      mbar_wait(tfull_bar + 8 * acc, (iter >> 1) & 1);
      tc_fence_after();
      const int m = mt * TILE_M + row;
      const bool valid = m < p.M;
      const int hw = p.H * p.W;
      const int img = valid ? m / hw : 0;
      // output placement
      const int ncol0 = nt * BLOCK_N;              // GEMM column of this tile
      const int q = p.up2 ? ncol0 / p.cout_real : 0;
      const int cc0 = p.up2 ? ncol0 % p.cout_real : ncol0;
      int64_t orow = m;
      if (p.up2) {
        const int w_ = m % p.W, r_ = m / p.W, h_ = r_ % p.H;
        orow = ((int64_t)img * 2 * p.H + 2 * h_ + (q >> 1)) * (2 * p.W) + 2 * w_ + (q & 1);
      }
      bf16* yrow = p.y + orow * p.ldy + cc0;
      const bf16* rrow = p.res ? p.res + (int64_t)m * p.ldres + cc0 : nullptr;
      const float* rvrow = p.rowvec ? p.rowvec + (int64_t)img * p.ld_rowvec + cc0 : nullptr;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N;
      float fo[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        if (valid) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + cc0 + c0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
          if (rvrow) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 b4 = __ldg(reinterpret_cast<const float4*>(rvrow + c0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
          if (rrow) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float t8[8];
              load_chunk(rrow + c0 + j, t8);
#pragma unroll
              for (int u = 0; u < 8; ++u) v[j + u] += t8[u];
            }
          }
          if (p.y) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              float t8[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) t8[u] = v[j + u];
              store_chunk(yrow + c0 + j, t8);
            }
          }
          if (p.fin_out) {
            for (int o = 0; o < p.fin_cout; ++o) {
              const float* wrow = p.fin_w + o * p.cout + c0;
              float s = 0.f;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow + j));
                s = fmaf(v[j], w4.x, s); s = fmaf(v[j + 1], w4.y, s);
                s = fmaf(v[j + 2], w4.z, s); s = fmaf(v[j + 3], w4.w, s);
              }
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (u == o) fo[u] += s;
            }
          }
        }
      }
      if (p.fin_out && valid) {
        const int pix = m - img * hw;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (u < p.fin_cout) p.fin_out[((int64_t)img * p.fin_cout + u) * hw + pix] = fo[u] + __ldg(p.fin_b + u);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ------------------------------------------------------------------ host side
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
int g_num_sms = 0;

int tc_init() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  LDM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  LDM_REQUIRE(qres == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  int dev = 0;
  LDM_CUDA(cudaGetDevice(&dev));
  LDM_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  int cc_major = 0;
  LDM_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  LDM_REQUIRE(cc_major == 10, "conv_tc: tcgen05 kernels need an sm_100-class GPU (found cc %d.x)", cc_major);
  LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<64>::SMEM_BYTES));
  LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM_BYTES));
  LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<256>::SMEM_BYTES));
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}

int make_act_map(CUtensorMap* map, const void* x, int ld, int cin, int B, int H, int W, int HB, int NB) {
  cuuint64_t gdim[4] = {(cuuint64_t)cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)W, (cuuint32_t)HB, (cuuint32_t)NB};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed with CUresult %d", (int)r);
  return 0;
}
int make_weight_map(CUtensorMap* map, const void* w, int cout, int ktot, int block_n) {
  cuuint64_t gdim[2] = {(cuuint64_t)ktot, (cuuint64_t)cout};
  cuuint64_t gstr[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)block_n};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LDM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weight) failed with CUresult %d", (int)r);
  return 0;
}

template <int BLOCK_N>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& ma2, const CUtensorMap& mb, const TcParams& p,
              cudaStream_t st) {
  using Cfg = TcCfg<BLOCK_N>;
  int tiles = p.num_m_tiles * p.num_n_tiles;
  int grid = tiles < g_num_sms ? tiles : g_num_sms;
  conv_tc_kernel<BLOCK_N><<<grid, 192, Cfg::SMEM_BYTES, st>>>(ma, ma2, mb, p);
  LDM_LAUNCHED("conv_tc");
  return 0;
}

}  // namespace

int k_conv_tc_prepare() { return tc_init(); }

int k_conv_tc(const ConvArgs& a, cudaStream_t st) {
  if (int rc = tc_init()) return rc;
  LDM_REQUIRE(a.dtype == LDM_DT_BF16, "conv_tc: bf16 only");
  LDM_REQUIRE(a.ksize == 1 || a.ksize == 3, "conv_tc: kernel size %d unsupported", a.ksize);
  LDM_REQUIRE(a.cin % BLOCK_K == 0 && (a.x2 == nullptr || a.cin2 % BLOCK_K == 0),
              "conv_tc: Cin (%d/%d) must be a multiple of 64", a.cin, a.cin2);
  LDM_REQUIRE(a.cout % 64 == 0, "conv_tc: Cout (%d) must be a multiple of 64", a.cout);
  LDM_REQUIRE(a.ldx % 8 == 0 && a.ldy % 8 == 0 && (!a.x2 || a.ldx2 % 8 == 0) && (!a.res || a.ldres % 8 == 0),
              "conv_tc: pixel strides must be multiples of 8 elements");
  LDM_REQUIRE(((uintptr_t)a.x & 15) == 0 && ((uintptr_t)a.w & 15) == 0 && ((uintptr_t)a.y & 15) == 0 &&
                  ((uintptr_t)a.fin_w & 15) == 0,
              "conv_tc: pointers must be 16-byte aligned");
  LDM_REQUIRE(!a.up2 || (a.ksize == 1 && !a.x2 && !a.res && !a.rowvec), "conv_tc: up2 epilogue only for plain 1x1 GEMMs");
  const int H = a.height, W = a.width;
  LDM_REQUIRE(W >= 1 && W <= TILE_M && TILE_M % W == 0, "conv_tc: width %d must divide 128", W);
  int HB = TILE_M / W;
  if (HB > H) HB = H;
  LDM_REQUIRE(H % HB == 0 && TILE_M % (W * HB) == 0, "conv_tc: %dx%d images do not tile into 128-pixel boxes", H, W);
  const int NB = TILE_M / (W * HB);
  LDM_REQUIRE(NB == 1 || HB == H, "conv_tc: internal tiling error");
  TcParams p;
  p.M = a.batch * H * W;
  if (p.M == 0) return 0;
  p.H = H; p.W = W; p.HB = HB; p.NB = NB;
  p.num_m_tiles = (p.M + TILE_M - 1) / TILE_M;
  p.taps = a.ksize * a.ksize;
  p.cin_blocks = a.cin / BLOCK_K;
  p.kb_main = p.taps * p.cin_blocks;
  p.kb_total = p.kb_main + (a.x2 ? a.cin2 / BLOCK_K : 0);
  p.cout = a.cout;
  p.up2 = a.up2;
  p.cout_real = a.up2 ? a.cout / 4 : a.cout;
  p.bias = a.bias;
  p.rowvec = a.rowvec; p.ld_rowvec = a.ld_rowvec;
  p.res = (const bf16*)a.res; p.ldres = a.ldres;
  p.y = (bf16*)a.y; p.ldy = a.ldy;
  p.fin_w = a.fin_w; p.fin_b = a.fin_b; p.fin_out = a.fin_out; p.fin_cout = a.fin_cout;
  LDM_REQUIRE(!a.fin_out || (a.fin_cout >= 1 && a.fin_cout <= 8 && !a.up2 && (a.cout == 64 || a.cout == 128 || a.cout == 256)),
              "conv_tc: fused projection needs Cout in {64,128,256} and <= 8 outputs");
  LDM_REQUIRE(a.y || a.fin_out, "conv_tc: no output requested");
  const int ktot = p.kb_total * BLOCK_K;

  // tile-N: the widest of 256/128/64 that divides Cout (and the up2 quadrant) and still yields >= 1 wave
  int block_n = 64;
  const int nlimit = p.cout_real;
  if (a.fin_out) block_n = a.cout;  // the whole channel row must sit in one accumulator tile
  else
  for (int bn : {256, 128}) {
    if (a.cout % bn == 0 && nlimit % bn == 0 && (int64_t)p.num_m_tiles * (a.cout / bn) >= g_num_sms) { block_n = bn; break; }
  }
  LDM_REQUIRE(nlimit % block_n == 0, "conv_tc: output channels (%d) must be a multiple of %d", nlimit, block_n);
  p.num_n_tiles = a.cout / block_n;

  CUtensorMap ma, ma2, mb;
  if (int rc = make_act_map(&ma, a.x, a.ldx, a.cin, a.batch, H, W, HB, NB)) return rc;
  if (a.x2) {
    if (int rc = make_act_map(&ma2, a.x2, a.ldx2, a.cin2, a.batch, H, W, HB, NB)) return rc;
  } else {
    ma2 = ma;
  }
  if (int rc = make_weight_map(&mb, a.w, a.cout, ktot, block_n)) return rc;
  switch (block_n) {
    case 256: return launch_tc<256>(ma, ma2, mb, p, st);
    case 128: return launch_tc<128>(ma, ma2, mb, p, st);
    default: return launch_tc<64>(ma, ma2, mb, p, st);
  }
}
