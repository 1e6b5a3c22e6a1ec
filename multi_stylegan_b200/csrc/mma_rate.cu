// Tensor-core roofline probe: the issue rate of tcgen05.mma.kind::tf32 (M = 128, N = 256, K = 8 — the instruction shape of
// the conv kernels' main loop) with both operands resident in shared memory and nothing else running: one CTA per SM,
// one thread issues `iters` x 4 back-to-back MMAs into a TMEM accumulator.  bench.py times launches of this kernel with
// CUDA events (burst and back-to-back for seconds, the way MEASURED_PEAKS.json measures bf16) and uses the result as the
// TF32 denominator of `roofline` — cuBLAS' own TF32 GEMM is slower than the product's conv kernel on this part, so it
// cannot serve as a peak.
#include "conv_common.cuh"
#include "sm100_ptx.cuh"

namespace msg {
using namespace ptx;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, float* sink) {
  constexpr int BN = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = 128 * 32 * 4, B_BYTES = BN * 32 * 4;
  const uint32_t sA = base, sB = base + A_BYTES;
  const uint32_t bar = sB + B_BYTES;
  const uint32_t slot = bar + 8;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - raw));
  float* f = reinterpret_cast<float*>(smem_raw + (base - raw));
  for (int i = threadIdx.x; i < (int)((A_BYTES + B_BYTES) / 4); i += blockDim.x) f[i] = 1e-3f * (float)((i * 7) & 15);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  if (threadIdx.x < 32) { tmem_alloc(slot, BN); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  constexpr uint32_t IDESC = make_idesc_tf32(128, BN, 0, 0);
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        mma_tf32(tmem, make_smem_desc(sA + k * 32, 0, 1024, SWZ_128B), make_smem_desc(sB + k * 32, 0, 1024, SWZ_128B), IDESC,
                 (it | k) ? 1u : 0u);
    }
    mma_commit(bar);
    mbar_wait(bar, 0);
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
  if (threadIdx.x < 32) tmem_dealloc(tmem, BN);
}

}  // namespace msg

using namespace msg;

// Launches the probe on one CTA per SM; returns the FLOPs of the launch in *flops (2 * 128 * 256 * 8 * 4 * iters * SMs).
extern "C" int msg_tf32_mma_rate_probe(int iters, float* sink, double* flops, msg_stream_t stream) {
  if (iters <= 0 || !flops) return fail(MSG_ERR_BAD_ARG, "tf32_mma_rate_probe: bad arguments");
  if (!tc_available()) return fail(MSG_ERR_UNSUPPORTED, "tf32_mma_rate_probe: needs an sm_100 device");
  const int blocks = num_sms();
  const size_t smem = 128 * 128 + 256 * 128 + 1024 + 64;
  static bool attr_done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (!attr_done[dev]) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev] = true;
  }
  mma_rate_kernel<<<blocks, 128, smem, (cudaStream_t)stream>>>(iters, sink);
  MSG_CHECK_LAUNCH("tf32_mma_rate_probe");
  *flops = 2.0 * 128 * 256 * 8 * 4.0 * (double)iters * blocks;
  return MSG_OK;
}
