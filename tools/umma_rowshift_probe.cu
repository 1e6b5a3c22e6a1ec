// Probe: can a K-major SWIZZLE_128B UMMA operand start at an arbitrary ROW of a 1024-byte-aligned shared-memory image
// (start address = base + r0 * 128), and does the descriptor's base-offset field have to carry (r0 & 7)?
// A is 160 rows x 32 floats laid out exactly as TMA writes a {32 ch, rows} box with the 128B swizzle; A[row][k] = row * 8 + k.
// B[n][k] = (n == k), N = 16, K = 8  ->  D[m][n] = A[r0 + m][n] for n < 8.  Prints the mismatch count for every r0 / mode.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I multi_stylegan_b200/csrc -I include \
//        -o gpurun_out/umma_probe tools/umma_rowshift_probe.cu && gpurun_out/umma_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"

using namespace msg::ptx;

__global__ void __launch_bounds__(128, 1) probe_kernel(int r0, int use_base_offset, int kk, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  constexpr int ROWS = 160;
  const uint32_t sA = base, sB = base + ROWS * 128;            // sB: 16 rows x 128 B, 1024-aligned (160 * 128 = 20480)
  float* fa = reinterpret_cast<float*>(gen);
  float* fb = reinterpret_cast<float*>(gen + ROWS * 128);
  for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) {
    const int row = i / 32, k = i % 32;
    const int chunk = k / 4, within = k % 4;
    fa[row * 32 + ((chunk ^ (row & 7)) * 4) + within] = (float)(row * 8 + (k % 8)) + (k >= 8 ? 1000.f * (k / 8) : 0.f);
  }
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) {
    const int n = i / 32, k = i % 32;
    const int chunk = k / 4, within = k % 4;
    fb[n * 32 + ((chunk ^ (n & 7)) * 4) + within] = ((k % 8) == n && n < 8) ? 1.f : 0.f;
  }
  const uint32_t bar = sB + 16 * 128;
  const uint32_t slot = bar + 8;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) { tmem_alloc(slot, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  constexpr uint32_t IDESC = make_idesc_tf32(128, 16, 0, 0);
  if (threadIdx.x == 0) {
    const uint32_t a_addr = sA + r0 * 128 + kk * 32;
    uint64_t ad = make_smem_desc(a_addr, 0, 1024, SWZ_128B);
    if (use_base_offset) ad |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t bd = make_smem_desc(sB + kk * 32, 0, 1024, SWZ_128B);
    mma_tf32(tmem, ad, bd, IDESC, 0u);
    mma_commit(bar);
    mbar_wait(bar, 0);
  }
  __syncthreads();
  tc_fence_after();
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float r[16];
    tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16), r);
    tmem_ld_wait();
    for (int n = 0; n < 16; ++n) out[(warp * 32 + lane) * 16 + n] = r[n];
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 32);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 16 * sizeof(float));
  static float h[128 * 16];
  const size_t smem = 160 * 128 + 16 * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int kk = 0; kk < 2; ++kk)
    for (int mode = 0; mode < 2; ++mode)
      for (int r0 = 0; r0 <= 12; ++r0) {
        probe_kernel<<<1, 128, smem>>>(r0, mode, kk, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 8; ++n) {
            const float want = (float)((r0 + m) * 8 + n) + (kk ? 1000.f * kk : 0.f);
            if (h[m * 16 + n] != want) { if (first < 0) first = m * 16 + n; ++bad; }
          }
        printf("kk=%d base_offset=%d r0=%2d: %4d mismatches%s (D[0][0]=%.0f D[1][0]=%.0f D[9][3]=%.0f) %s\n", kk, mode, r0, bad,
               bad ? " <-- WRONG" : "", h[0], h[16], h[9 * 16 + 3], cudaGetErrorString(e));
      }
  cudaFree(d);
  return 0;
}
