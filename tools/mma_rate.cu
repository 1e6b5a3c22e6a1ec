// Microbenchmark: issue rate of tcgen05.mma kind::tf32 (M=128) for the operand layouts the conv engines use,
// with operands resident in shared memory (no TMA in the loop).  Diagnostic, not a product kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multi_stylegan_b200/csrc -I include \
//        -o gpurun_out/mma_rate tools/mma_rate.cu && gpurun_out/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"

using namespace msg::ptx;

template <int BN, int MN_MAJOR>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* cycles, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = 128 * 32 * 4, B_BYTES = BN * 32 * 4;
  const uint32_t sA = base, sB = base + A_BYTES;
  const uint32_t bar = sB + B_BYTES;
  const uint32_t slot = bar + 8;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));
  // deterministic small values so the accumulators stay finite
  float* f = reinterpret_cast<float*>(smem_raw + (base - raw));
  for (int i = threadIdx.x; i < (int)((A_BYTES + B_BYTES) / 4); i += blockDim.x) f[i] = 1e-3f * (float)((i * 7) & 15);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) { tmem_alloc(slot, BN < 32 ? 32 : BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  constexpr uint32_t IDESC = make_idesc_tf32(128, BN, MN_MAJOR, MN_MAJOR);
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t ad, bd;
        if (MN_MAJOR) {
          ad = make_smem_desc(sA + k * 1024, 4096, 512, SWZ_128B_BASE32B);
          bd = make_smem_desc(sB + k * 1024, 4096, 512, SWZ_128B_BASE32B);
        } else {
          ad = make_smem_desc(sA + k * 32, 0, 1024, SWZ_128B);
          bd = make_smem_desc(sB + k * 32, 0, 1024, SWZ_128B);
        }
        mma_tf32(tmem, ad, bd, IDESC, (it | k) ? 1u : 0u);
      }
    }
    mma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < 32) {
    float r[16];
    tmem_ld_32x16(tmem, r);
    tmem_ld_wait();
    if (sink) sink[blockIdx.x * 32 + threadIdx.x] = r[0];
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, BN < 32 ? 32 : BN);
}

template <int BN, int MN>
static void run(const char* name, int blocks) {
  const int iters = 4096;
  long long* cyc; float* sink;
  cudaMalloc(&cyc, blocks * sizeof(long long));
  cudaMalloc(&sink, blocks * 32 * sizeof(float));
  const size_t smem = 128 * 128 + BN * 128 + 1024 + 64;
  cudaFuncSetAttribute(mma_rate_kernel<BN, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  mma_rate_kernel<BN, MN><<<blocks, 128, smem>>>(64, cyc, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  mma_rate_kernel<BN, MN><<<blocks, 128, smem>>>(iters, cyc, sink);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  long long h = 0; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double flops = 2.0 * 128 * BN * 8 * 4.0 * iters * blocks;
  printf("%-34s blocks=%3d  %8.1f clk per 128x%dx8 MMA   %7.1f TFLOP/s aggregate   (%.3f ms, %s)\n", name, blocks,
         (double)h / (4.0 * iters), BN, flops / (ms * 1e-3) / 1e12, ms, cudaGetErrorString(e));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int blocks : {1, 148}) {
    run<256, 0>("K-major SW128   N=256", blocks);
    run<128, 0>("K-major SW128   N=128", blocks);
    run<64, 0>("K-major SW128   N=64", blocks);
    run<256, 1>("MN-major BASE32B N=256", blocks);
    run<128, 1>("MN-major BASE32B N=128", blocks);
  }
  return 0;
}
