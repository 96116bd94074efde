// Epilogue shared by the tcgen05 convolution kernels (conv_tc.cu, conv_halo.cu): TMEM accumulator -> +bias (+row vector)
// (+residual) -> bf16 NHWC store, the fused 1x1 output projection, and -- new in round 2 -- GroupNorm fused into the
// convolution that PRODUCES the normalised tensor (src/UNet.py:52-58 Block, :106 PreNorm, :147 to_out GroupNorm, :20 Residual).
//
// Why here: GroupNorm was 31 % of a timestep at 0.35 of the HBM peak as separate statistics + apply kernels (three passes
// over every activation).  The accumulator tile is already in TMEM in fp32, so
//   mode 1 (statistics only): the epilogue leaves per-(sample, group, M tile, 64-column block) sums {S, Q} of the values it
//           stores; the consumer (fused LinearAttention / gn_apply) adds the few slots in a fixed order.  No extra pass.
//   mode 2 (normalise): pass 1 reads the accumulator and only takes statistics (nothing is stored); pass 2 reads the SAME
//           accumulator again from TMEM, applies (x - mean) * rstd * gamma + beta (+ time-embedding row) -> SiLU
//           (+ residual) and stores the normalised tensor: the raw conv output never exists in memory.
//           When a sample spans several tiles (32x32 layers: 9 halo tiles; GroupNorm(1, C) over several N tiles) the
//           tiles exchange their partial sums through tagged 16-byte packets in global memory ({S, Q, tag, ~tag}: flag
//           in data, one store to publish, one load to observe -- no fences, no atomics, no resets: the caller zeroes
//           the packet area once per forward and gives every launch its own tag).  Pass 2 is deferred by one work unit
//           (the TMEM ring is 4 deep where it has to be) and its packet loads are issued a whole pass earlier, so the
//           packets of the neighbouring CTAs -- which work on the same sample at the same time -- have long landed.
//   Partial sums have one canonical granularity -- fp32 over (one 128-row M tile or one whole small sample) x (one group,
//   or one 64-column block of GroupNorm(1, C)) -- and are combined in double in (M tile, column block) order, whatever
//   the tile shape the launcher picked: a sample's bits do not depend on the batch size, the grid or the GPU count.
//   Deadlock freedom of the polling: CTAs are persistent and co-resident (grid <= #SMs, 1 CTA/SM); pass 2 of unit i
//   only waits for pass 1 of units i-1 .. i+1 of other CTAs, whose own program-order predecessors are pass 2 of strictly
//   older units -- the chain ends at the first unit.  The poll is bounded and traps instead of hanging.
//   Latency: nothing the epilogue needs per tile is fetched from global memory on its critical path -- bias / gamma /
//   beta are staged once per CTA, row vectors arrive by cp.async one unit ahead, packets are requested one pass ahead
//   (round-2 measurement: three exposed L2 round trips per tile made the fused conv 3x slower than conv + GroupNorm).
#pragma once
#include "tc_common.cuh"

namespace tc {

struct EpiP {
  unsigned long long* dbg = nullptr;   // timeline probe (LDM_HALO_DEBUG bit 2): globaltimer stamps of CTA 0's first epilogue warp
  // geometry
  int M, H, W, hw;             // valid GEMM rows (linear geometry), image size, pixels per image
  int P, tiles_per_image;      // halo geometry (P > 0): padded pitch W+2, tiles per padded plane
  int num_m_tiles, num_n_tiles;
  int batch;
  int up2, cout, cout_real;
  // convolution-level operands
  const float* bias;
  const float* rowvec; int ld_rowvec;
  const bf16* res; int ldres; int res_mod;
  bf16* y; int ldy;
  const float* fin_w; const float* fin_b; float* fin_out; int fin_cout;
  int debug;
  // fused GroupNorm (gn_mode 0: off)
  int gn_mode, gn_G, gn_cpg, gn_silu, gn_nvar, gn_var_rows;
  int gn_nslots;               // packets / partial slots per (sample, group[, variant])
  int gn_cross;                // 1: a sample's statistics span several work units -> packet exchange, deferred pass 2
  int gn_upt;                  // M tiles per sample
  float gn_eps;
  unsigned gn_tag;
  const float* gn_gamma; const float* gn_beta;
  const float* gn_rowvec; int gn_ld_rowvec;
  const bf16* gn_res; int gn_ldres;
  void* gn_scratch;
};

// rows of per-sample / per-variant vectors the epilogue keeps in shared memory
__host__ __device__ constexpr int epi_vec_rows(int block_n) { return block_n == 256 ? 4 : 8; }
// shared-memory floats the epilogue needs for a BLOCK_N-wide tile:
//   bias [512] | {gamma, beta} [512] | union { projection weights [768] + sums [1024] ,
//   polled packets [512] + row vectors [R][BN] + {a, b} [R][BN] + partial sums [2][8][32][2] + {mean, rstd} [2][8][8] }
__host__ __device__ constexpr int epi_gn_floats(int block_n) { return 512 + 3 * epi_vec_rows(block_n) * block_n + 1024 + 256; }
__host__ __device__ constexpr int epi_smem_floats(int block_n) {
  return 512 + 1024 + (epi_gn_floats(block_n) > 1792 ? epi_gn_floats(block_n) : 1792);
}
__host__ __device__ constexpr int epi_smem_bytes(int block_n) { return epi_smem_floats(block_n) * 4; }
constexpr int EPI_FULL_VEC = 512;   // bias / gamma / beta are staged whole when the tensor has at most this many channels

__device__ __forceinline__ float silu_mufu(float x) {
  float h = 0.5f * x, t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ uint4 ld_volatile_v4(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_packet(void* p, float s, float q, unsigned tag) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(__float_as_uint(s)), "r"(__float_as_uint(q)), "r"(tag), "r"(~tag)
               : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// One epilogue thread's view of a work unit row.
struct EpiRow {
  bool valid;
  int n, pix;
  int64_t m;   // n * hw + pix
};

// GM: fused GroupNorm mode, a compile-time constant -- the plain convolution (GM = 0) carries none of the GroupNorm code
// (round-2 measurement: the larger instruction footprint alone cost the MMA-issuing warp ~10 % at 32x32)
template <int BLOCK_N, int MT, int NACC, bool HALO, int GM>
__device__ __forceinline__ void conv_epilogue(const EpiP& p, const uint32_t tmem_base, const uint32_t tfull_bar,
                                              const uint32_t tempty_bar, float* const s_epi, const int num_units) {
  static_assert(!HALO || MT == 1, "the halo kernel works on single tiles");
  constexpr int COLS = BLOCK_N / 2;          // columns per warp (two warps per TMEM lane quarter)
  constexpr int VR = epi_vec_rows(BLOCK_N);
  float* const s_bias = s_epi;                                           // [512]  whole vector, or [2][256] per-unit slices
  float2* const s_gb = reinterpret_cast<float2*>(s_bias + 512);          // [512]  {gamma, beta}
  float* const s_finw = reinterpret_cast<float*>(s_gb + 512);            // [768]  fused projection weights ...
  float* const s_fin = s_finw + 768;                                     // [128][8] ... and cross-warp sums; never together with GroupNorm:
  float2* const s_poll = reinterpret_cast<float2*>(s_finw);              // [256] polled packets
  float* const s_rv = s_finw + 512;                                      // [VR][BLOCK_N] GroupNorm row vectors (ring of units)
  float2* const s_ab = reinterpret_cast<float2*>(s_rv + VR * BLOCK_N);   // [VR][BLOCK_N] {a, b}: y = acc * a + b
  float* const s_red = reinterpret_cast<float*>(s_ab + VR * BLOCK_N);    // [2][8][32][2] per (variant, row segment, 8-column block) {S, Q}
  float2* const s_stat = reinterpret_cast<float2*>(s_red + 1024);        // [2][8][8] {mean, rstd} per (sample in unit, group in tile, variant)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quarter = warp & 3;              // TMEM lanes [32*quarter, +32) are the ones this warp may read
  const int half = (warp - 2) >> 2;          // column half of the tile
  const int cw = half * COLS;
  const int row = quarter * 32 + lane;
  const int et = threadIdx.x - 64;           // 0..255 among the epilogue threads
  const bool keep_l2 = (p.debug & 8) == 0;
  const uint64_t l2pol = l2_evict_last_policy();
  // loop-invariant parameters in registers (no constant-bank traffic per tile)
  const int num_n_tiles = p.num_n_tiles, num_m_tiles = p.num_m_tiles, M = p.M, H = p.H, W = p.W, hw = p.hw, up2 = p.up2;
  const int P = p.P, tpi = p.tiles_per_image;
  const int cout_real = p.cout_real, ldy = p.ldy, ldres = p.ldres, ld_rowvec = p.ld_rowvec;
  const int fin_cout = p.fin_cout, res_mod = p.res_mod, batch = p.batch;
  bf16* const y = p.y;
  const bf16* const res = p.res;
  const float* const bias = p.bias;
  const float* const rowvec = p.rowvec;
  float* const fin_out = p.fin_out;
  const bool no_mem = HALO ? (p.debug & 1) : (p.debug & 2);
  // LDM_EPI_DEBUG (timing experiments only; results are wrong): 1 no statistics math, 2 no packet wait, 4 no pass-2 stores,
  // 8 no pass-2 TMEM re-read / math
  const int xdbg = p.debug >> 16;
  auto stamp = [&](int it, int slot) {
    if (p.dbg && blockIdx.x == 0 && warp == 2 && lane == 0 && it < 32) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[it * 16 + slot] = t;
    }
  };
  // GroupNorm
  constexpr int gmode = GM;
  constexpr bool mode2 = GM == 2;
  const int G = p.gn_G, cpg = p.gn_cpg, nvar = GM ? p.gn_nvar : 1, nslots = p.gn_nslots;
  const bool cross = mode2 && p.gn_cross != 0;
  const int seg = hw < 32 && !HALO ? hw : 32;                    // lanes of a warp that belong to one sample
  const bool big = HALO || hw >= TILE_M;                         // a 128-row tile lies inside one sample
  const int spu = big ? 1 : TILE_M / hw;                         // samples per work unit (MT == 1 whenever > 1)
  const int GT = G == 1 ? 1 : BLOCK_N / cpg;                     // groups inside one N tile
  const int cbs = G == 1 ? BLOCK_N / 64 : GT;                    // statistics column blocks per tile (64 columns, or a group)
  const int c8_per_cb = G == 1 ? 8 : cpg / 8;
  const int rs_per_stat = big ? TILE_M / 32 : hw / seg;          // row segments per statistics unit (M tile, or small sample)
  const int su_count = big ? MT : spu;                           // statistics units per work unit
  const int n_pub = su_count * cbs * nvar;                       // partial sums this unit publishes
  const int slots_per_mtile = G == 1 ? p.cout / 64 : 1;
  // mode 1 (warp-local partial sums): 8-column blocks merged per published sum, column sub-blocks per group
  const int m1 = G == 1 ? 4 : (cpg >= 32 ? 4 : cpg / 8);
  const int csubs = G == 1 ? p.cout / 32 : (cpg > 32 ? cpg / 32 : 1);
  const int rows_ab = spu * nvar;                                // {a, b} rows per unit
  const int n_poll = mode2 && cross ? rows_ab * GT * nslots : 0;
  const float* const gn_rowvec = p.gn_rowvec;
  const bf16* const gn_res = p.gn_res;
  const int gn_ldres = p.gn_ldres, gn_ld_rowvec = p.gn_ld_rowvec, var_rows = p.gn_var_rows;
  const bool full = cout_real <= EPI_FULL_VEC;
  const bool has_rv = gmode != 0 && gn_rowvec != nullptr;
  // row-vector ring: deferred pass 2 reads the vectors of the previous unit while the next unit's arrive
  const int rv_bufs = cross ? 3 : (2 * rows_ab <= VR ? 2 : 1);

  auto unit_um = [&](int unit) { return HALO ? unit : unit / num_n_tiles; };
  auto unit_nt = [&](int unit) { return HALO ? 0 : unit - (unit / num_n_tiles) * num_n_tiles; };
  auto row_of = [&](int unit, int sub) {
    EpiRow r;
    if (HALO) {
      r.n = unit / tpi;
      const int q = P + (unit - r.n * tpi) * TILE_M + row;
      const int rr = q / P, cc = q - rr * P;
      r.valid = rr >= 1 && rr <= H && cc >= 1 && cc <= W && !no_mem;
      r.pix = (rr - 1) * W + (cc - 1);
      r.m = (int64_t)r.n * hw + r.pix;
    } else {
      const int mt = unit_um(unit) * MT + sub;
      const int m = mt * TILE_M + row;
      r.valid = m < M && mt < num_m_tiles && !no_mem;
      r.n = r.valid ? m / hw : 0;
      r.pix = m - r.n * hw;
      r.m = m;
    }
    return r;
  };
  // sample of statistics unit `su` (an M tile of a big image, or one small sample) and the M tile's index in its sample
  auto stat_sample = [&](int unit, int su) {
    if (HALO) return unit / tpi;
    const int um = unit_um(unit);
    return big ? (int)(((int64_t)(um * MT + su) * TILE_M) / hw) : um * spu + su;
  };
  auto stat_mtile = [&](int unit, int su) {
    if (HALO) return unit % tpi;
    return big ? (int)((((int64_t)(unit_um(unit) * MT + su) * TILE_M) % hw) / TILE_M) : 0;
  };
  // asynchronous copy of the unit's GroupNorm row vectors into ring buffer `it % rv_bufs`
  auto rv_fetch = [&](int unit, int it) {
    float* dst = s_rv + (it % rv_bufs) * rows_ab * BLOCK_N;
    const int ncol0 = unit_nt(unit) * BLOCK_N;
    for (int c = et; c < rows_ab * BLOCK_N; c += 256) {
      const int rI = c / BLOCK_N, col = c - rI * BLOCK_N;
      const int su = rI / nvar, k = rI - su * nvar;
      const int n = stat_sample(unit, big ? 0 : su);
      if (n < batch) cp_async4(dst + c, gn_rowvec + (int64_t)(n + k * var_rows) * gn_ld_rowvec + ncol0 + col);
    }
  };

  // ---------------------------------------------------------------- once per CTA: whole-vector staging
  if (full) {
    for (int c = et; c < cout_real; c += 256) {
      s_bias[c] = bias ? __ldg(bias + c) : 0.f;
      if (mode2) s_gb[c] = make_float2(__ldg(p.gn_gamma + c), __ldg(p.gn_beta + c));
    }
  }
  if (fin_out) {  // fused output projection weights: [fin_cout][BLOCK_N] (host guarantees fin_cout * BLOCK_N <= 768)
    for (int c = et; c < fin_cout * BLOCK_N; c += 256) s_finw[c] = p.fin_w[c];
  }
  if (has_rv && (int)blockIdx.x < num_units && rv_bufs > 1) rv_fetch(blockIdx.x, 0);
  epi_bar();

  int iter = 0;
  int prev_unit = -1;
  uint4 pk = make_uint4(0u, 0u, 0u, 0u);       // this thread's packet of the previous unit, requested a pass ahead
  const uint4* pk_addr = nullptr;
  for (int unit = blockIdx.x;; unit += gridDim.x, ++iter) {
    const bool has = unit < num_units;
    if (!has && !(mode2 && cross && prev_unit >= 0)) break;
    // request the previous unit's packets now: they are consumed after this unit's pass 1
    if (n_poll && prev_unit >= 0 && et < n_poll) {
      const int slot = et % nslots, r2 = et / nslots;
      const int gg = r2 % GT, rI = r2 / GT, su = rI / nvar, k = rI - su * nvar;
      const int n = stat_sample(prev_unit, big ? 0 : su);
      const int g = G == 1 ? 0 : unit_nt(prev_unit) * BLOCK_N / cpg + gg;
      pk_addr = n < batch ? reinterpret_cast<const uint4*>(p.gn_scratch) + (((int64_t)n * G + g) * nvar + k) * nslots + slot : nullptr;
      if (pk_addr) pk = ld_volatile_v4(pk_addr);
    }
    if (has) {
      const int acc = iter % NACC;
      const int ncol0 = unit_nt(unit) * BLOCK_N;               // GEMM column of this tile
      const int q = up2 ? ncol0 / cout_real : 0;
      const int cc0 = up2 ? ncol0 - q * cout_real : ncol0;
      const float* sb = s_bias + cc0;
      bool need_bar = mode2;                           // GroupNorm scratch of the previous unit must be free
      if (!full) {   // wide tensors (> 512 channels): the tile's bias slice, double-buffered by unit parity
        float* dst = s_bias + (iter & 1) * 256;
        for (int c = et; c < BLOCK_N; c += 256) dst[c] = bias ? __ldg(bias + cc0 + c) : 0.f;
        sb = dst;
        need_bar = true;
      }
      if (has_rv) {
        need_bar = true;
        if (rv_bufs == 1) { epi_bar(); rv_fetch(unit, 0); }   // no room for a ring: fetch in place (small low-resolution layers)
        cp_async_wait_all();
      }
      if (need_bar) epi_bar();                         // staged vectors visible
      if (has_rv && rv_bufs > 1 && unit + (int)gridDim.x < num_units) rv_fetch(unit + gridDim.x, iter + 1);

      // ------------------------------------------------------------ pass 1
      bool waited = false;
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub) {
        if (!HALO && unit_um(unit) * MT + sub >= num_m_tiles) break;   // uniform over the CTA
        const EpiRow r = row_of(unit, sub);
        int64_t orow = r.m;
        if (up2) {
          const int m32 = (int)r.m;
          const int r_ = m32 / W, w_ = m32 - r_ * W, h_ = r_ % H;
          orow = ((int64_t)r.n * 2 * H + 2 * h_ + (q >> 1)) * (2 * W) + 2 * w_ + (q & 1);
        }
        bf16* yrow = y ? y + orow * ldy + cc0 + cw : nullptr;
        const int64_t mres = res_mod > 0 ? (int64_t)(r.n % res_mod) * hw + r.pix : r.m;
        const bf16* rrow = (res && r.valid) ? res + mres * ldres + cc0 + cw : nullptr;
        const float* rvrow = rowvec ? rowvec + (int64_t)r.n * ld_rowvec + cc0 + cw : nullptr;
        uint4 rr[4];
        if (rrow) {   // residual row: requested BEFORE waiting for the accumulator, so its latency hides behind the MMAs
#pragma unroll
          for (int j = 0; j < 4; ++j) rr[j] = __ldg(reinterpret_cast<const uint4*>(rrow) + j);
        }
        if (!waited) {
          stamp(iter, 4);
          if (lane == 0) mbar_wait(tfull_bar + 8 * acc, (iter / NACC) & 1);
          __syncwarp();
          stamp(iter, 5);
          waited = true;
        }
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (acc * MT + sub) * BLOCK_N + cw;
        const int su = big ? 0 : row / hw;                       // sample of this row inside the unit
        const int rsi = big ? sub * 4 + quarter : row / seg;     // row segment of this lane
        // mode 1 bookkeeping: first GEMM row of this M tile, its sample, whether the tile exists
        bool m1_writer = false;
        float* m1_base = nullptr;
        if (GM == 1) {
          const int d1 = seg >> 1, d2 = seg >> 2, d3 = seg >> 3;
          const int64_t m0 = HALO ? 0 : (int64_t)(unit_um(unit) * MT + sub) * TILE_M + quarter * 32 + (lane & ~(seg - 1));
          const int n = HALO ? unit / tpi : (int)(m0 / hw);
          const int rowblk = HALO ? (unit % tpi) * 4 + quarter : (hw >= 32 ? (int)(m0 % hw) >> 5 : 0);
          m1_writer = (lane & (d3 - 1)) == 0 && (m1 < 2 || !(lane & d2)) && (m1 < 4 || !(lane & d1)) && n < batch && (HALO || m0 < M);
          m1_base = reinterpret_cast<float*>(p.gn_scratch) + (((int64_t)n * G * nvar * nslots + rowblk * csubs) << 1) + ((lane & d3) ? 1 : 0);
        }
        float fo[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int c0 = 0; c0 < COLS; c0 += 32) {
          uint32_t rg[32];
          tmem_ld32(taddr + c0, rg);
          uint4 rn[4];  // next chunk's residual, in flight while this chunk is processed
          if (rrow && c0 + 32 < COLS) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rn[j] = __ldg(reinterpret_cast<const uint4*>(rrow + c0 + 32) + j);
          }
          tmem_ld_wait();
          if (!mode2 && (sub == MT - 1 || (!HALO && unit_um(unit) * MT + sub + 1 >= num_m_tiles)) && c0 + 32 >= COLS) {
            // last read of this unit's accumulator: hand the TMEM buffer back NOW, before the arithmetic and the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(tempty_bar + 8 * acc);
            stamp(iter, 6);
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = *reinterpret_cast<const float4*>(sb + cw + c0 + j);
            v[j] = __uint_as_float(rg[j]) + b4.x; v[j + 1] = __uint_as_float(rg[j + 1]) + b4.y;
            v[j + 2] = __uint_as_float(rg[j + 2]) + b4.z; v[j + 3] = __uint_as_float(rg[j + 3]) + b4.w;
          }
          if (r.valid) {
            if (rvrow) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 b4 = __ldg(reinterpret_cast<const float4*>(rvrow + c0 + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            }
            if (rrow) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rr[j]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float2 f = __bfloat1622float2(h2[u]);
                  v[8 * j + 2 * u] += f.x; v[8 * j + 2 * u + 1] += f.y;
                }
              }
            }
            if (yrow && !mode2) {
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
                const float* wrow = s_finw + o * BLOCK_N + cw + c0;
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
          if (gmode && !(xdbg & 1)) {
            // statistics of the values as the consumer sees them (+ the row vector added before the normalisation):
            // eight partial sums per lane ({S, Q} of four 8-column blocks), reduced over the `seg` lanes of a sample by a
            // transposing butterfly -- 9 shuffles instead of 40; lane bits (d1, d2, d3) end up holding sum number idx
            const int d1 = seg >> 1, d2 = seg >> 2, d3 = seg >> 3;
            const bool h1 = lane & d1, h2 = lane & d2, h3 = lane & d3;
#pragma unroll 1
            for (int k = 0; k < nvar; ++k) {
              const float* rvk = s_rv + ((iter % rv_bufs) * rows_ab + su * nvar + k) * BLOCK_N + cw + c0;
              float a8[8];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float s = 0.f, qv = 0.f;
                if (r.valid) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const float d = v[8 * u + i] + (has_rv ? rvk[8 * u + i] : 0.f);
                    s += d; qv = fmaf(d, d, qv);
                  }
                }
                a8[2 * u] = s; a8[2 * u + 1] = qv;
              }
              float b4[4], c2[2];
#pragma unroll
              for (int i = 0; i < 4; ++i) b4[i] = (h1 ? a8[i + 4] : a8[i]) + __shfl_xor_sync(0xffffffffu, h1 ? a8[i] : a8[i + 4], d1);
#pragma unroll
              for (int i = 0; i < 2; ++i) c2[i] = (h2 ? b4[i + 2] : b4[i]) + __shfl_xor_sync(0xffffffffu, h2 ? b4[i] : b4[i + 2], d2);
              float dd = (h3 ? c2[1] : c2[0]) + __shfl_xor_sync(0xffffffffu, h3 ? c2[0] : c2[1], d3);
              for (int off = d3 >> 1; off > 0; off >>= 1) dd += __shfl_xor_sync(0xffffffffu, dd, off);
              const int idx = (h1 ? 4 : 0) + (h2 ? 2 : 0) + (h3 ? 1 : 0);
              if (gmode == 1) {
                // statistics only: every warp publishes its own partial sums -- (one row segment) x (group, or 32-column
                // chunk of GroupNorm(1, C)) -- straight to the consumer's slots: no shared scratch, no CTA barrier
                if (m1 >= 2) dd += __shfl_xor_sync(0xffffffffu, dd, d2);
                if (m1 == 4) dd += __shfl_xor_sync(0xffffffffu, dd, d1);
                if (m1_writer) {
                  const int col = ncol0 + cw + c0 + 8 * (idx >> 1);       // first column of this partial sum
                  const int g = G == 1 ? 0 : col / cpg;
                  const int csub = G == 1 ? col >> 5 : (cpg > 32 ? (col - g * cpg) >> 5 : 0);
                  m1_base[(((int64_t)g * nvar + k) * nslots + csub) << 1] = dd;
                }
              } else if ((lane & (d3 - 1)) == 0) {
                s_red[(((k * 8 + rsi) * 32 + ((cw + c0) >> 3) + (idx >> 1)) << 1) + (idx & 1)] = dd;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) rr[j] = rn[j];
        }
        if (fin_out) {
          // the two column halves of a row live in two warps: combine through shared memory, half 0 writes
          float* fx = s_fin + row * 8;
          if (half == 1) {
#pragma unroll
            for (int u = 0; u < 8; ++u) fx[u] = fo[u];
          }
          epi_bar();
          if (half == 0 && r.valid) {
#pragma unroll
            for (int u = 0; u < 8; ++u)
              if (u < fin_cout) fin_out[((int64_t)r.n * fin_cout + u) * hw + r.pix] = fo[u] + fx[u] + __ldg(p.fin_b + u);
          }
          epi_bar();  // fx is rewritten by the next sub-tile
        }
      }   // sub-tiles
      // (!mode2: the TMEM buffer was handed back right after its last read, above)
      stamp(iter, 7);
      if (mode2) {
        epi_bar();   // s_red complete
        // canonical partial sums: fp32 over (statistics unit) x (column block), in (row segment, 8-column block) order
        if (cross && et < n_pub) {
          const int k = et % nvar, r2 = et / nvar, cb = r2 % cbs, su = r2 / cbs;
          const int n = stat_sample(unit, su);
          if (n < batch && (HALO || !big || unit_um(unit) * MT + su < num_m_tiles)) {
            float S = 0.f, Q = 0.f;
            for (int rs = su * rs_per_stat; rs < (su + 1) * rs_per_stat; ++rs)
              for (int c8 = cb * c8_per_cb; c8 < (cb + 1) * c8_per_cb; ++c8) {
                const float2 f = *reinterpret_cast<const float2*>(s_red + (((k * 8 + rs) * 32 + c8) << 1));
                S += f.x; Q += f.y;
              }
            const int g = G == 1 ? 0 : ncol0 / cpg + cb;
            const int slot = stat_mtile(unit, su) * slots_per_mtile + (G == 1 ? ncol0 / 64 + cb : 0);
            st_packet(reinterpret_cast<uint4*>(p.gn_scratch) + (((int64_t)n * G + g) * nvar + k) * nslots + slot, S, Q, p.gn_tag);
          }
        }
        if (!cross && et < rows_ab * GT) {
          // the unit holds whole samples: mean / rstd from the same canonical partial sums, combined in double
          const int k = et % nvar, r2 = et / nvar, gg = r2 % GT, su = r2 / GT;
          double S = 0.0, Q = 0.0;
          const int nst = big ? MT : 1;                       // M tiles of this sample inside the unit
          const int ncb = G == 1 ? BLOCK_N / 64 : 1;          // 64-column blocks of the group inside the tile
          for (int t = 0; t < nst; ++t)
            for (int cb = 0; cb < ncb; ++cb) {
              float s = 0.f, qv = 0.f;
              const int rs0 = (big ? t : su) * rs_per_stat, c80 = (G == 1 ? cb : gg) * c8_per_cb;
              for (int rs = rs0; rs < rs0 + rs_per_stat; ++rs)
                for (int c8 = c80; c8 < c80 + c8_per_cb; ++c8) {
                  const float2 f = *reinterpret_cast<const float2*>(s_red + (((k * 8 + rs) * 32 + c8) << 1));
                  s += f.x; qv += f.y;
                }
              S += (double)s; Q += (double)qv;
            }
          const double cnt = (double)hw * (double)cpg;
          const double mean = S / cnt;
          double var = Q / cnt - mean * mean;
          if (var < 0.0) var = 0.0;
          s_stat[et] = make_float2((float)mean, 1.0f / sqrtf((float)var + p.gn_eps));
        }
      }
    }   // has

    // -------------------------------------------------------------- pass 2: normalise + store (GroupNorm mode 2)
    if (mode2) {
      const int u2 = cross ? prev_unit : unit;       // deferred by one unit when the statistics cross tiles
      const int it2 = cross ? iter - 1 : iter;
      if (cross) prev_unit = has ? unit : -1;
      if (u2 >= 0) {
        const int acc2 = it2 % NACC;
        const int ncol2 = unit_nt(u2) * BLOCK_N;
        if (cross && et < n_poll) {
          float2 v = make_float2(0.f, 0.f);
          if (pk_addr && !(xdbg & 2)) {
            const long long start = clock64();
            while (!(pk.z == p.gn_tag && pk.w == ~p.gn_tag)) {
              __nanosleep(20);
              pk = ld_volatile_v4(pk_addr);
              if (clock64() - start > 4000000000LL) __trap();
            }
            v = make_float2(__uint_as_float(pk.x), __uint_as_float(pk.y));
          }
          s_poll[et] = v;
        }
        epi_bar();   // polled packets (cross) / s_stat (whole samples) complete
        // y = acc * a + b with a = gamma * rstd, b = beta + (bias + rowvec - mean) * a, per (sample / variant, column)
        const float* rvb = s_rv + (it2 % rv_bufs) * rows_ab * BLOCK_N;
        for (int e = et; e < rows_ab * BLOCK_N; e += 256) {
          const int rI = e / BLOCK_N, col = e - rI * BLOCK_N;
          const int gg = G == 1 ? 0 : col / cpg;
          float mean, rstd;
          if (cross) {
            const float2* pp = s_poll + (rI * GT + gg) * nslots;   // rI = su * nvar + k
            double S = 0.0, Q = 0.0;
            for (int s = 0; s < nslots; ++s) { S += (double)pp[s].x; Q += (double)pp[s].y; }
            const double cnt = (double)hw * (double)cpg;
            const double m = S / cnt;
            double var = Q / cnt - m * m;
            if (var < 0.0) var = 0.0;
            mean = (float)m;
            rstd = 1.0f / sqrtf((float)var + p.gn_eps);
          } else {
            const int su = rI / nvar, k = rI - su * nvar;
            const float2 st = s_stat[(su * GT + gg) * nvar + k];
            mean = st.x; rstd = st.y;
          }
          const float2 gb = s_gb[ncol2 + col];
          const float a = gb.x * rstd;
          s_ab[e] = make_float2(a, fmaf(s_bias[ncol2 + col] + (has_rv ? rvb[e] : 0.f) - mean, a, gb.y));
        }
        epi_bar();
#pragma unroll 1
        for (int sub = 0; sub < MT; ++sub) {
          if (!HALO && unit_um(u2) * MT + sub >= num_m_tiles) break;
          if (xdbg & 8) break;
          EpiRow r = row_of(u2, sub);
          if (xdbg & 4) r.valid = false;
          const int su = big ? 0 : row / hw;
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (acc2 * MT + sub) * BLOCK_N + cw;
          const bf16* rrow = (gn_res && r.valid) ? gn_res + r.m * gn_ldres + ncol2 + cw : nullptr;
          uint4 rr[4];
          if (rrow) {
#pragma unroll
            for (int j = 0; j < 4; ++j) rr[j] = __ldg(reinterpret_cast<const uint4*>(rrow) + j);
          }
#pragma unroll 1
          for (int c0 = 0; c0 < COLS; c0 += 32) {
            uint32_t rg[32];
            tmem_ld32(taddr + c0, rg);
            uint4 rn[4];
            if (rrow && c0 + 32 < COLS) {
#pragma unroll
              for (int j = 0; j < 4; ++j) rn[j] = __ldg(reinterpret_cast<const uint4*>(rrow + c0 + 32) + j);
            }
            tmem_ld_wait();
            if (r.valid) {
#pragma unroll 1
              for (int k = 0; k < nvar; ++k) {
                bf16* orow = y + ((int64_t)(r.n + k * var_rows) * hw + r.pix) * ldy + ncol2 + cw + c0;
                const float4* ab = reinterpret_cast<const float4*>(s_ab + (su * nvar + k) * BLOCK_N + cw + c0);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  float o[8];
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float4 w = ab[4 * u + i];   // {a, b} of two columns
                    float t0 = fmaf(__uint_as_float(rg[8 * u + 2 * i]), w.x, w.y);
                    float t1 = fmaf(__uint_as_float(rg[8 * u + 2 * i + 1]), w.z, w.w);
                    if (p.gn_silu) { t0 = silu_mufu(t0); t1 = silu_mufu(t1); }
                    o[2 * i] = t0; o[2 * i + 1] = t1;
                  }
                  if (rrow) {
                    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&rr[u]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      const float2 f = __bfloat1622float2(h2[i]);
                      o[2 * i] += f.x; o[2 * i + 1] += f.y;
                    }
                  }
                  if (keep_l2) store_chunk_keep(orow + 8 * u, o, l2pol); else store_chunk(orow + 8 * u, o);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) rr[j] = rn[j];
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(tempty_bar + 8 * acc2);
      }
    }
  }
}

}  // namespace tc
