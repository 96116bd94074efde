// Microbenchmark: issue rate / execution time of tcgen05.mma kind::f16 128xNx16 from shared-memory operands.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../latent-diffusion-models_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace tc;
int ldm_set_error(const char*, ...) { return -1; }
std::atomic<long long> g_ldm_launches{0};

template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int iters, int shift_rows, int commit_each) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar, bar2, bar3;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // zero operands
  for (int i = threadIdx.x; i < 190 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (i * 2654435761u) ^ (blockIdx.x * 40503u);
    h ^= h >> 13; h *= 0x5bd1e995u; h ^= h >> 15;
    // two bf16 in [-2, 2): sign | exponent 0x3f/0x40 | random mantissa
    uint32_t v = shift_rows >= 100 ? ((h & 0x807f807fu) | 0x3f803f80u) : 0u;
    reinterpret_cast<uint32_t*>(smem_raw)[i] = v;
  }
  if (shift_rows >= 100) shift_rows -= 100;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); mbar_init(smem_u32(&bar3), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc(N);
    const uint32_t a0 = base, b0 = base + 48 * 1024;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      // MODE 0: one A slab, 9 tap-shifted descriptors x 4 k-steps (the conv_halo inner loop)
      // MODE 1: same descriptors every time (no address math)
      // MODE 2: 1024-aligned A starts (tap stride 16 KB)
#pragma unroll 1
      for (int t = 0; t < 9; ++t) {
        uint32_t a_addr = a0;
        if (MODE == 0) a_addr = a0 + (uint32_t)(35 + shift_rows + (t / 3 - 1) * 34 + (t % 3 - 1)) * 128u;
        if (MODE == 2) a_addr = a0 + (uint32_t)(t & 1) * 16384u;
        const uint64_t adesc = make_sw128_desc(a_addr);
        const uint64_t bdesc = make_sw128_desc(b0 + (MODE == 1 ? 0 : t) * (N * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + (it & 1) * N, adesc + 2 * k, bdesc + 2 * k, idesc, (t > 0 || k > 0) ? 1u : 0u);
      }
      if (commit_each) { umma_commit(smem_u32(&bar2)); umma_commit(smem_u32(&bar3)); }
    }
    long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int N, int MODE>
void run(const char* name, int shift, int commit_each = 0) {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 200;
  mma_rate_kernel<N, MODE><<<148, 128, 200 * 1024>>>(d, iters, shift, commit_each);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  double n_mma = iters * 36.0;
  printf("%-28s N=%3d shift=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (nominal %d)  %s\n", name, N, shift, h[0] / n_mma, h[1] / n_mma, N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<64, 1>("fixed desc", 0); run<64, 0>("halo-shifted A", 0); run<64, 0>("halo-shifted A + commits", 0, 1);
  run<64, 1>("fixed desc + commits", 0, 1);
  run<128, 0>("halo-shifted A + commits", 0, 1);
  run<64, 1>("fixed desc RANDOM data", 100); run<64, 0>("halo-shifted RANDOM data", 100, 1);
  run<128, 0>("halo-shifted RANDOM data", 100, 1); run<256, 1>("fixed desc RANDOM data", 100);
  return 0;
}
