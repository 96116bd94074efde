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
// Execution (one persistent CTA per SM, 320 threads)
//   warp 0      TMA producer (one lane): STAGES-deep mbarrier ring
//   warp 1      TMEM allocator + tcgen05.mma issuer (one lane): 4 x (128 x BLOCK_N x 16) MMAs per k-block,
//               tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2..9  epilogue (two per TMEM lane quarter, half the columns each): tcgen05.ld (32 lanes x 32 columns) -> +bias +time-embedding row vector
//               +residual -> bf16 -> 16-byte global stores.  Two accumulator buffers in TMEM (2*BLOCK_N
//               columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <stdlib.h>

#include <initializer_list>

#include "kernels.h"
#include "tc_common.cuh"
#include "conv_epilogue.cuh"

// timeline probe (LDM_TC_DEBUG=1): CTA 0 records globaltimer stamps of its first tiles
__device__ unsigned long long g_tc_dbg[1024];
extern "C" int ldm_debug_read_tc(unsigned long long* host, int n) {
  return cudaMemcpyFromSymbol(host, g_tc_dbg, sizeof(unsigned long long) * (n < 1024 ? n : 1024)) == cudaSuccess ? 0 : -1;
}

namespace {

using namespace tc;
__device__ __forceinline__ unsigned long long gtime_tc() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(slot)                                                                     \
  do {                                                                                     \
    if (p.debug && blockIdx.x == 0 && iter < 60) g_tc_dbg[iter * 16 + (slot)] = gtime_tc(); \
  } while (0)
constexpr int A_STAGE_BYTES = TILE_M * BLOCK_K * 2;

struct TcParams {
  int H;          // image height
  int HB, NB;     // M-tile box: NB images x HB rows x W columns = 128 pixels
  int num_m_tiles, num_n_tiles;
  int kb_main;    // k-blocks from the main source = taps * cin/64
  int kb_total;   // + cin2/64
  int cin_blocks; // cin/64
  int taps;       // 1 or 9
  int debug;
  EpiP e;         // everything the epilogue warps need (conv_epilogue.cuh)
};

// MT: M-tiles (128 output pixels each) a CTA works on at once.  They share every weight k-block: the kernel is bound by
// L2->SM traffic (profiles/README.md finding 8), and with MT = 2 the weight half of it is fetched once per 256 pixels.
template <int BLOCK_N, int MT = 1>
struct TcCfg {
  static constexpr int B_STAGE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int A_BYTES = MT * A_STAGE_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = MT == 1 ? (BLOCK_N == 64 ? 8 : (BLOCK_N == 128 ? 6 : 4)) : (BLOCK_N == 64 ? 5 : 4);
  // accumulator ring: as deep as TMEM allows, at most 4 (the fused GroupNorm defers its second pass by one work unit)
  static constexpr int NACC = 512 / (MT * BLOCK_N) >= 4 ? 4 : 2;
  static constexpr int TMEM_COLS = NACC * MT * BLOCK_N;  // power of two >= 32
  // mbarriers + TMEM slot (1 KB) + the epilogue's staging area
  static constexpr int BAR_BYTES = 1024 + epi_smem_bytes(BLOCK_N);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // + alignment slack
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory plan exceeds 227 KB");
};

template <int BLOCK_N, int MT = 1, int GM = 0>
__global__ void __launch_bounds__(320, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
               const __grid_constant__ CUtensorMap tmap_b, const TcParams p) {
  using Cfg = TcCfg<BLOCK_N, MT>;
  static_assert(Cfg::TMEM_COLS <= 512, "accumulators do not fit TMEM");
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B needs 1024-B alignment
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * Cfg::A_BYTES;
  const uint32_t bars = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
  constexpr int NACC = Cfg::NACC;
  const uint32_t tfull_bar = bars + 16 * STAGES, tempty_bar = tfull_bar + 8 * NACC;
  const uint32_t tmem_slot = tempty_bar + 8 * NACC;
  float* s_epi = reinterpret_cast<float*>(smem_raw + (bars + 1024 - smem_u32(smem_raw)));   // epilogue staging area
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
    for (int i = 0; i < NACC; ++i) {
      mbar_init(tfull_bar + 8 * i, 1);
      mbar_init(tempty_bar + 8 * i, 8);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();      // everything above overlapped the previous kernel's tail; its outputs are visible from here on
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int num_tiles = ((p.num_m_tiles + MT - 1) / MT) * p.num_n_tiles;   // work units: MT stacked M-tiles x one N-tile

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (all loop-carried parameters live in registers: the kernel-parameter constant bank is not touched per k-block)
    if (lane == 0) {
      const int num_n_tiles = p.num_n_tiles, NB = p.NB, HB = p.HB, tpi = p.NB == 1 ? p.H / p.HB : 1;
      const int kb_main = p.kb_main, kb_total = p.kb_total, cin_blocks = p.cin_blocks;
      const bool three = p.taps == 9;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int um = tile / num_n_tiles, nt = tile - um * num_n_tiles;
        int n0[MT], h0[MT];   // tiles past the last one read past the batch: TMA zero fill, never stored
#pragma unroll
        for (int j = 0; j < MT; ++j) {
          const int mt = um * MT + j;
          if (NB == 1) {
            n0[j] = mt / tpi;
            h0[j] = (mt - n0[j] * tpi) * HB;
          } else {
            n0[j] = mt * NB;
            h0[j] = 0;
          }
        }
        int tap = 0, cb = 0;  // (tap, channel block) of the main source, advanced incrementally
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          mbar_expect_tx(full_bar + 8 * stage, Cfg::STAGE_BYTES);
          if (kb < kb_main) {
            int dy = 0, dx = 0;
            if (three) { dy = tap / 3 - 1; dx = tap - (dy + 1) * 3 - 1; }
#pragma unroll
            for (int j = 0; j < MT; ++j)
              tma_load_4d(smem_a + stage * Cfg::A_BYTES + j * A_STAGE_BYTES, &tmap_a, full_bar + 8 * stage, cb * BLOCK_K, dx,
                          h0[j] + dy, n0[j]);
            if (++cb == cin_blocks) { cb = 0; ++tap; }
          } else {
#pragma unroll
            for (int j = 0; j < MT; ++j)
              tma_load_4d(smem_a + stage * Cfg::A_BYTES + j * A_STAGE_BYTES, &tmap_a2, full_bar + 8 * stage,
                          (kb - kb_main) * BLOCK_K, 0, h0[j], n0[j]);
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
      const int kb_total = p.kb_total;
      const uint64_t adesc0 = make_sw128_desc(smem_a), bdesc0 = make_sw128_desc(smem_b);
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++iter) {
        const int acc = iter % NACC;
        TC_STAMP(0);
        mbar_wait(tempty_bar + 8 * acc, ((iter / NACC) & 1) ^ 1);
        TC_STAMP(1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * MT * BLOCK_N;
        uint32_t accum = 0u;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          if (kb == 0) TC_STAMP(2);
          tc_fence_after();
          const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (Cfg::B_STAGE_BYTES >> 4));
#pragma unroll
          for (int j = 0; j < MT; ++j) {
            const uint64_t adesc = adesc0 + (uint64_t)(stage * (Cfg::A_BYTES >> 4) + j * (A_STAGE_BYTES >> 4));
            // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the (addr>>4) field
            umma_bf16(tmem_d + j * BLOCK_N, adesc, bdesc, idesc, accum);
#pragma unroll
            for (int k = 1; k < BLOCK_K / UMMA_K; ++k) umma_bf16(tmem_d + j * BLOCK_N, adesc + 2 * k, bdesc + 2 * k, idesc, 1u);
          }
          accum = 1u;
          umma_commit(empty_bar + 8 * stage);  // frees the smem stage when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar + 8 * acc);  // accumulator complete
        TC_STAMP(3);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9): conv_epilogue.cuh =====================
    conv_epilogue<BLOCK_N, MT, NACC, false, GM>(p.e, tmem_base, tfull_bar, tempty_bar, s_epi, num_tiles);
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
  static bool inited[64] = {};   // function attributes (and the SM count) are per device
  int dev = 0;
  LDM_CUDA(cudaGetDevice(&dev));
  if (g_encode && inited[dev & 63]) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  LDM_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  LDM_REQUIRE(qres == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  LDM_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  int cc_major = 0;
  LDM_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  LDM_REQUIRE(cc_major == 10, "conv_tc: tcgen05 kernels need an sm_100-class GPU (found cc %d.x)", cc_major);
#define TC_ATTR(BN, MTV)                                                                                                         \
  LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, MTV, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN, MTV>::SMEM_BYTES)); \
  LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, MTV, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN, MTV>::SMEM_BYTES)); \
  LDM_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, MTV, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN, MTV>::SMEM_BYTES))
  TC_ATTR(64, 1); TC_ATTR(128, 1); TC_ATTR(256, 1); TC_ATTR(64, 2); TC_ATTR(128, 2);
#undef TC_ATTR
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  inited[dev & 63] = true;
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

template <int BLOCK_N, int MT>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& ma2, const CUtensorMap& mb, const TcParams& p,
              cudaStream_t st) {
  using Cfg = TcCfg<BLOCK_N, MT>;
  int tiles = ((p.num_m_tiles + MT - 1) / MT) * p.num_n_tiles;
  int grid = tiles < g_num_sms ? tiles : g_num_sms;
  switch (p.e.gn_mode) {
    case 0: LDM_CUDA(ldm_launch_pdl(conv_tc_kernel<BLOCK_N, MT, 0>, dim3(grid), dim3(320), (size_t)Cfg::SMEM_BYTES, st, ma, ma2, mb, p)); break;
    case 1: LDM_CUDA(ldm_launch_pdl(conv_tc_kernel<BLOCK_N, MT, 1>, dim3(grid), dim3(320), (size_t)Cfg::SMEM_BYTES, st, ma, ma2, mb, p)); break;
    default: LDM_CUDA(ldm_launch_pdl(conv_tc_kernel<BLOCK_N, MT, 2>, dim3(grid), dim3(320), (size_t)Cfg::SMEM_BYTES, st, ma, ma2, mb, p)); break;
  }
  LDM_LAUNCHED("conv_tc");
  return 0;
}

}  // namespace

int k_conv_tc_prepare() {
  if (int rc = tc_init()) return rc;
  return k_conv_halo_prepare();
}

int k_conv_tc(const ConvArgs& a, cudaStream_t st) {
  if (int rc = tc_init()) return rc;
  LDM_REQUIRE(a.dtype == LDM_DT_BF16, "conv_tc: bf16 only");
  LDM_REQUIRE(a.res_mod == 0, "conv_tc: residual row aliasing is only implemented in the halo kernel");
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
  EpiP& e = p.e;
  e.M = a.batch * H * W;
  if (e.M == 0) return 0;
  p.H = H; p.HB = HB; p.NB = NB;
  e.H = H; e.W = W; e.hw = H * W; e.P = 0; e.tiles_per_image = 0; e.batch = a.batch;
  p.num_m_tiles = (e.M + TILE_M - 1) / TILE_M;
  p.taps = a.ksize * a.ksize;
  p.cin_blocks = a.cin / BLOCK_K;
  p.kb_main = p.taps * p.cin_blocks;
  p.kb_total = p.kb_main + (a.x2 ? a.cin2 / BLOCK_K : 0);
  e.cout = a.cout;
  e.up2 = a.up2;
  e.cout_real = a.up2 ? a.cout / 4 : a.cout;
  e.bias = a.bias;
  e.rowvec = a.rowvec; e.ld_rowvec = a.ld_rowvec;
  e.res = (const bf16*)a.res; e.ldres = a.ldres; e.res_mod = 0;
  e.y = (bf16*)a.y; e.ldy = a.ldy;
  e.fin_w = a.fin_w; e.fin_b = a.fin_b; e.fin_out = a.fin_out; e.fin_cout = a.fin_cout;
  { const char* d = getenv("LDM_TC_DEBUG"); p.debug = d ? atoi(d) : 0; }
  e.debug = p.debug | ((getenv("LDM_EPI_DEBUG") ? atoi(getenv("LDM_EPI_DEBUG")) : 0) << 16);
  LDM_REQUIRE(!a.fin_out || (a.fin_cout >= 1 && a.fin_cout <= 8 && !a.up2 && (a.cout == 64 || a.cout == 128 || a.cout == 256) &&
                             a.fin_cout * a.cout <= 768),
              "conv_tc: fused projection needs Cout in {64,128,256} and fin_cout * Cout <= 768");
  LDM_REQUIRE(a.y || a.fin_out, "conv_tc: no output requested");
  const int ktot = p.kb_total * BLOCK_K;

  // tile-N: the widest of 256/128/64 that divides Cout (and the up2 quadrant) and still yields >= 1 wave
  int block_n = 64;
  const int nlimit = e.cout_real;
  if (a.fin_out) block_n = a.cout;  // the whole channel row must sit in one accumulator tile
  else
  for (int bn : {256, 128}) {
    if (a.cout % bn == 0 && nlimit % bn == 0 && (int64_t)p.num_m_tiles * (a.cout / bn) >= g_num_sms) { block_n = bn; break; }
  }
  LDM_REQUIRE(nlimit % block_n == 0, "conv_tc: output channels (%d) must be a multiple of %d", nlimit, block_n);
  // two M-tiles per CTA when that still leaves every SM at least two work units
  static const int mt_env = getenv("LDM_TC_MT") ? atoi(getenv("LDM_TC_MT")) : 2;
  bool pair = mt_env == 2 && block_n <= 128 && !a.fin_out && (int64_t)(p.num_m_tiles / 2) * (a.cout / block_n) >= 2 * g_num_sms;

  // ---- fused GroupNorm plan (conv_epilogue.cuh)
  const ConvGn& g = a.gn;
  e.gn_mode = g.mode;
  e.gn_G = 1; e.gn_cpg = a.cout; e.gn_silu = 0; e.gn_nvar = 1; e.gn_var_rows = 0; e.gn_nslots = 1; e.gn_cross = 0; e.gn_upt = 1;
  e.gn_eps = g.eps; e.gn_tag = g.tag; e.gn_gamma = g.gamma; e.gn_beta = g.beta; e.gn_rowvec = g.rowvec; e.gn_ld_rowvec = g.ld_rowvec;
  e.gn_res = (const bf16*)g.res; e.gn_ldres = g.ldres; e.gn_scratch = g.scratch;
  if (g.mode) {
    LDM_REQUIRE(g.mode == 1 || g.mode == 2, "conv_tc: GroupNorm epilogue mode %d unknown", g.mode);
    LDM_REQUIRE(!a.up2 && !a.fin_out && a.y, "conv_tc: the GroupNorm epilogue needs a plain NHWC output");
    LDM_REQUIRE(g.groups >= 1 && a.cout % g.groups == 0 && (a.cout / g.groups) % 8 == 0, "conv_tc: GroupNorm groups %d vs Cout %d", g.groups, a.cout);
    LDM_REQUIRE(g.nvar == 1, "conv_tc: GroupNorm variants are only implemented in the halo kernel");
    LDM_REQUIRE(g.mode == 1 || (g.gamma && g.beta && !a.res && !a.rowvec), "conv_tc: GroupNorm mode 2 takes gamma/beta and no pre-norm residual / row vector");
    LDM_REQUIRE(!g.res || g.ldres % 8 == 0, "conv_tc: GroupNorm residual stride must be a multiple of 8");
    const int cpg = a.cout / g.groups, hw = H * W;
    if (g.groups > 1) while (block_n % cpg != 0 && block_n < 256) block_n *= 2;   // a group never straddles N tiles
    LDM_REQUIRE(g.groups == 1 || (block_n % cpg == 0 && a.cout % block_n == 0), "conv_tc: GroupNorm groups of %d channels do not fit an N tile", cpg);
    int n_tiles = a.cout / block_n;
    auto geometry_ok = [&](int mt) {
      const int ru = TILE_M * mt;
      return hw >= ru ? hw % ru == 0 : (ru % hw == 0 && ru / hw <= 8 && hw >= 16 && mt == 1);
    };
    if (pair && !geometry_ok(2)) pair = false;
    LDM_REQUIRE(geometry_ok(pair ? 2 : 1), "conv_tc: GroupNorm epilogue does not support %dx%d images", H, W);
    LDM_REQUIRE(g.mode == 1 || a.cout <= EPI_FULL_VEC, "conv_tc: GroupNorm mode 2 supports at most %d channels", EPI_FULL_VEC);
    const int upt = hw >= TILE_M ? hw / TILE_M : 1;          // M tiles per sample
    const int spu = hw >= TILE_M ? 1 : TILE_M / hw;          // samples per M tile
    // {a, b} / row-vector rows the epilogue keeps per unit: one per sample in the tile
    if (spu > epi_vec_rows(block_n)) block_n = 128;
    auto is_cross = [&](int mt, int bn) { return (hw > TILE_M * mt || (g.groups == 1 && a.cout / bn > 1)) ? 1 : 0; };
    int cross = is_cross(pair ? 2 : 1, block_n);
    if (g.mode == 2 && cross) {
      // deferred second pass: needs the 4-deep accumulator ring (MT * BLOCK_N <= 128)
      if (pair && block_n > 64) pair = false;
      if (block_n > 128) block_n = 128;
      cross = is_cross(pair ? 2 : 1, block_n);
    }
    LDM_REQUIRE(a.cout % block_n == 0 && (g.groups == 1 || block_n % cpg == 0), "conv_tc: GroupNorm tile plan failed (Cout %d, tile %d)", a.cout, block_n);
    n_tiles = a.cout / block_n;
    e.gn_G = g.groups; e.gn_cpg = cpg; e.gn_silu = g.silu; e.gn_upt = upt; e.gn_cross = g.mode == 2 ? cross : 0;
    e.gn_nslots = upt * (g.groups == 1 ? a.cout / 64 : 1);
    if (g.mode == 1)   // warp-local partial sums: (32-row block, or a whole small sample) x (group, or 32-column chunk)
      e.gn_nslots = (hw >= 32 ? hw / 32 : 1) * (g.groups == 1 ? a.cout / 32 : (cpg > 32 ? cpg / 32 : 1));
    LDM_REQUIRE(g.mode != 2 || !cross || spu * (g.groups == 1 ? 1 : block_n / cpg) * e.gn_nslots <= 256, "conv_tc: GroupNorm packet fan-in too large");
    LDM_REQUIRE(!(cross && g.rowvec && spu > 1), "conv_tc: GroupNorm row vectors with several samples per tile need whole groups per tile");
    if (g.mode == 1) {
      LDM_REQUIRE(g.scratch && g.scratch_bytes >= (int64_t)a.batch * g.groups * e.gn_nslots * 8, "conv_tc: GroupNorm statistics buffer too small");
    } else if (cross) {
      LDM_REQUIRE(g.scratch && g.tag != 0 && g.scratch_bytes >= (int64_t)a.batch * g.groups * e.gn_nslots * 16 && ((uintptr_t)g.scratch & 15) == 0,
                  "conv_tc: GroupNorm packet buffer missing / too small (need %lld bytes)", (long long)a.batch * g.groups * e.gn_nslots * 16);
    }
    if (g.nslots_out) *g.nslots_out = e.gn_nslots;
  }
  p.num_n_tiles = a.cout / block_n;
  e.num_m_tiles = p.num_m_tiles; e.num_n_tiles = p.num_n_tiles;

  CUtensorMap ma, ma2, mb;
  if (int rc = make_act_map(&ma, a.x, a.ldx, a.cin, a.batch, H, W, HB, NB)) return rc;
  if (a.x2) {
    if (int rc = make_act_map(&ma2, a.x2, a.ldx2, a.cin2, a.batch, H, W, HB, NB)) return rc;
  } else {
    ma2 = ma;
  }
  if (int rc = make_weight_map(&mb, a.w, a.cout, ktot, block_n)) return rc;
  switch (block_n) {
    case 256: return launch_tc<256, 1>(ma, ma2, mb, p, st);
    case 128: return pair ? launch_tc<128, 2>(ma, ma2, mb, p, st) : launch_tc<128, 1>(ma, ma2, mb, p, st);
    default: return pair ? launch_tc<64, 2>(ma, ma2, mb, p, st) : launch_tc<64, 1>(ma, ma2, mb, p, st);
  }
}
