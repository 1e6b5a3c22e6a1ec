// Probe: kernel-to-kernel gap inside a replayed CUDA graph with and without programmatic dependent launch
// (cudaLaunchAttributeProgrammaticStreamSerialization; every kernel starts with griddepcontrol.wait).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/pdl_probe tools/pdl_probe.cu && tools/bin/pdl_probe
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void step_kernel(float* p, int work) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = p[i];
  for (int k = 0; k < work; ++k) v = v * 1.0000001f + 1.0f;
  p[i] = v;
}

static float run(int n, int work, int pdl, int ctas, float* buf, cudaStream_t st) {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, step_kernel, buf, work);
  }
  cudaStreamEndCapture(st, &graph);
  cudaGraphInstantiate(&exec, graph, 0);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; ++w) cudaGraphLaunch(exec, st);
  cudaEventRecord(a, st);
  for (int r = 0; r < 10; ++r) cudaGraphLaunch(exec, st);
  cudaEventRecord(b, st);
  cudaStreamSynchronize(st);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
  return ms * 1000.f / (10.f * n);
}

int main() {
  cudaStream_t st;
  cudaStreamCreate(&st);
  float* buf;
  cudaMalloc(&buf, 148 * 8 * 256 * sizeof(float));
  cudaMemset(buf, 0, 148 * 8 * 256 * sizeof(float));
  const int works[] = {0, 2000, 20000};
  const int ctas[] = {148, 148 * 8};
  for (int c : ctas)
    for (int w : works) {
      const float off = run(1000, w, 0, c, buf, st), on = run(1000, w, 1, c, buf, st);
      printf("ctas %4d work %5d: %.3f us/kernel plain edges, %.3f us/kernel programmatic edges (%.3f us saved)\n", c, w, off, on,
             off - on);
    }
  float h[4];
  cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost);
  printf("err: %s, check %.1f\n", cudaGetErrorString(cudaGetLastError()), h[0]);
  return 0;
}
