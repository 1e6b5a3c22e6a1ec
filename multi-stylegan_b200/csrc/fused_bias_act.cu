// fused_bias_act for sm_100a — bandwidth-bound elementwise op.
//
// Reference semantics: multi_stylegan/op_static/fused_bias_act_kernel.cu:18-49 (arithmetic),
// :52-99 (launcher), op_static/fused_act.py:22-51 (backward = same op + a separate ATen sum).
// Written from scratch: 128-bit accesses, one (sample, channel) plane per blockIdx.x so the bias
// lookup needs no per-element div/mod, 64-bit indexing, and the bias-gradient reduction fused
// into the backward pass (deterministic two-stage, no atomics).
#include "common.cuh"

namespace msg {

enum FbaMode { FBA_LINEAR = 0, FBA_LRELU_X = 1, FBA_LRELU_REF = 2, FBA_ZERO = 3 };

template <int MODE, typename T>
__device__ __forceinline__ T fba_apply(T x, T ref, T alpha, T scale) {
  T y;
  if (MODE == FBA_LINEAR) y = x;
  else if (MODE == FBA_LRELU_X) y = (x > T(0)) ? x : x * alpha;
  else if (MODE == FBA_LRELU_REF) y = (ref > T(0)) ? x : x * alpha;
  else y = T(0);
  return y * scale;
}

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

// ---- planar kernel: blockIdx.x = plane (n*C + c), blockIdx.y = chunk of the plane ---------------
// VEC: elements per access (4 for aligned fp32, 1 otherwise). Each thread handles LOOP accesses.
template <int MODE, typename T, int VEC, int LOOP, bool WITH_SUM>
__global__ void __launch_bounds__(256)
fba_planar_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ bias,
                  const T* __restrict__ ref, T alpha, T scale, int64_t step_b, int size_b,
                  T* __restrict__ partial) {
  const int64_t plane = blockIdx.x;
  const T b = bias ? bias[plane % size_b] : T(0);
  const int64_t base = plane * step_b;
  const int64_t chunk0 = (int64_t)blockIdx.y * (256 * VEC * LOOP);
  T acc = T(0);
  if constexpr (VEC == 4) {
    using V = float4;  // only instantiated for float
    const V* x4 = reinterpret_cast<const V*>(x + base);
    const V* r4 = reinterpret_cast<const V*>(ref ? ref + base : nullptr);
    V* o4 = reinterpret_cast<V*>(out + base);
    const int64_t n4 = step_b >> 2;
    V xv[LOOP], rv[LOOP];
#pragma unroll
    for (int l = 0; l < LOOP; ++l) {
      const int64_t i = (chunk0 >> 2) + l * 256 + threadIdx.x;
      if (i < n4) {
        xv[l] = __ldg(x4 + i);
        if (MODE == FBA_LRELU_REF) rv[l] = __ldg(r4 + i);
      }
    }
#pragma unroll
    for (int l = 0; l < LOOP; ++l) {
      const int64_t i = (chunk0 >> 2) + l * 256 + threadIdx.x;
      if (i < n4) {
        V o;
        const V r = (MODE == FBA_LRELU_REF) ? rv[l] : make_float4(0, 0, 0, 0);
        o.x = fba_apply<MODE, float>(xv[l].x + b, r.x, alpha, scale);
        o.y = fba_apply<MODE, float>(xv[l].y + b, r.y, alpha, scale);
        o.z = fba_apply<MODE, float>(xv[l].z + b, r.z, alpha, scale);
        o.w = fba_apply<MODE, float>(xv[l].w + b, r.w, alpha, scale);
        if (WITH_SUM) acc += (o.x + o.y) + (o.z + o.w);
        o4[i] = o;
      }
    }
  } else {
#pragma unroll
    for (int l = 0; l < LOOP; ++l) {
      const int64_t i = chunk0 + l * 256 + threadIdx.x;
      if (i < step_b) {
        const T r = (MODE == FBA_LRELU_REF) ? ref[base + i] : T(0);
        const T o = fba_apply<MODE, T>(x[base + i] + b, r, alpha, scale);
        if (WITH_SUM) acc += o;
        out[base + i] = o;
      }
    }
  }
  if (WITH_SUM) {
    __shared__ T scratch[32];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) partial[plane * gridDim.y + blockIdx.y] = acc;
  }
}

// ---- generic kernel: any step_b (incl. 1 for [N, C] inputs), per-element bias index --------------
template <int MODE, typename T>
__global__ void __launch_bounds__(256)
fba_generic_kernel(T* __restrict__ out, const T* __restrict__ x, const T* __restrict__ bias,
                   const T* __restrict__ ref, T alpha, T scale, int64_t size_x, int64_t step_b,
                   int64_t size_b) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < size_x; i += stride) {
    T v = x[i];
    if (bias) v += bias[(i / step_b) % size_b];
    const T r = (MODE == FBA_LRELU_REF) ? ref[i] : T(0);
    out[i] = fba_apply<MODE, T>(v, r, alpha, scale);
  }
}

// dbias[c] = sum over (outer, chunks) of partial[(n*C + c) * nchunks + j]
template <typename T>
__global__ void __launch_bounds__(256)
fba_reduce_partials_kernel(T* __restrict__ dbias, const T* __restrict__ partial, int64_t outer,
                           int size_b, int nchunks) {
  const int c = blockIdx.x;
  T acc = T(0);
  const int64_t total = outer * nchunks;
  for (int64_t t = threadIdx.x; t < total; t += blockDim.x) {
    const int64_t n = t / nchunks, j = t - n * nchunks;
    acc += partial[(n * size_b + c) * nchunks + j];
  }
  __shared__ T scratch[32];
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) dbias[c] = acc;
}

// dbias[c] = sum_{n, p} dx[n, c, p]  (small-tensor path)
template <typename T>
__global__ void __launch_bounds__(256)
fba_reduce_direct_kernel(T* __restrict__ dbias, const T* __restrict__ dx, int64_t outer, int size_b,
                         int64_t step_b) {
  const int c = blockIdx.x;
  T acc = T(0);
  const int64_t total = outer * step_b;
  for (int64_t t = threadIdx.x; t < total; t += blockDim.x) {
    const int64_t n = t / step_b, p = t - n * step_b;
    acc += dx[(n * size_b + c) * step_b + p];
  }
  __shared__ T scratch[32];
  acc = block_sum(acc, scratch);
  if (threadIdx.x == 0) dbias[c] = acc;
}

static inline int fba_mode(int act, int grad) {
  switch (act * 10 + grad) {
    case 30: return FBA_LRELU_X;
    case 31: return FBA_LRELU_REF;
    case 12:
    case 32: return FBA_ZERO;
    default: return FBA_LINEAR;  // reference: `default: case 10: case 11:` -> y = x
  }
}

constexpr int kPlanarLoop = 4;
constexpr int64_t kPlanarMinStep = 1024;  // below this one block per plane wastes the SM

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
static int launch_fba(T* out, const T* x, const T* bias, const T* ref, int mode, T alpha, T scale,
                      int64_t size_x, int64_t step_b, int64_t size_b, T* partial, int* nchunks_out,
                      cudaStream_t st) {
  if (size_x == 0) return MSG_OK;
  const bool has_bias = bias != nullptr;
  const bool planar = step_b >= kPlanarMinStep && (size_x % step_b == 0) &&
                      (!has_bias || size_b > 0) && (size_x / step_b) <= 0x7fffffffLL;
  if (partial && !planar) return fail(MSG_ERR_BAD_ARG, "fused_bias_act: fused sum needs the planar path");
  if (planar) {
    const int64_t planes = size_x / step_b;
    const bool vec = sizeof(T) == 4 && (step_b % 4 == 0) && aligned16(x) && aligned16(out) &&
                     (mode != FBA_LRELU_REF || aligned16(ref));
    const int per_block = 256 * (vec ? 4 : 1) * kPlanarLoop;
    const int nchunks = (int)ceil_div(step_b, per_block);
    if (nchunks > 65535) return fail(MSG_ERR_UNSUPPORTED, "fused_bias_act: plane too large");
    if (nchunks_out) *nchunks_out = nchunks;
    dim3 grid((unsigned)planes, (unsigned)nchunks);
    const int sb = has_bias ? (int)size_b : 1;
#define FBA_LAUNCH(MODE, VEC, SUM)                                                            \
  fba_planar_kernel<MODE, T, VEC, kPlanarLoop, SUM><<<grid, 256, 0, st>>>(out, x, bias, ref,   \
                                                                          alpha, scale, step_b, sb, partial)
#define FBA_MODE_SWITCH(VEC, SUM)                          \
  switch (mode) {                                          \
    case FBA_LINEAR: FBA_LAUNCH(FBA_LINEAR, VEC, SUM); break;     \
    case FBA_LRELU_X: FBA_LAUNCH(FBA_LRELU_X, VEC, SUM); break;   \
    case FBA_LRELU_REF: FBA_LAUNCH(FBA_LRELU_REF, VEC, SUM); break; \
    default: FBA_LAUNCH(FBA_ZERO, VEC, SUM); break;        \
  }
    bool launched = false;
    if constexpr (sizeof(T) == 4) {
      if (vec) {
        if (partial) { FBA_MODE_SWITCH(4, true) } else { FBA_MODE_SWITCH(4, false) }
        launched = true;
      }
    }
    if (!launched) {
      if (partial) { FBA_MODE_SWITCH(1, true) } else { FBA_MODE_SWITCH(1, false) }
    }
#undef FBA_MODE_SWITCH
#undef FBA_LAUNCH
    MSG_CHECK_LAUNCH("fused_bias_act(planar)");
    return MSG_OK;
  }
  const int64_t want = ceil_div(size_x, 256);
  const unsigned grid = (unsigned)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
  const int64_t sb = has_bias ? size_b : 1, stb = step_b > 0 ? step_b : 1;
  switch (mode) {
    case FBA_LINEAR: fba_generic_kernel<FBA_LINEAR, T><<<grid, 256, 0, st>>>(out, x, bias, ref, alpha, scale, size_x, stb, sb); break;
    case FBA_LRELU_X: fba_generic_kernel<FBA_LRELU_X, T><<<grid, 256, 0, st>>>(out, x, bias, ref, alpha, scale, size_x, stb, sb); break;
    case FBA_LRELU_REF: fba_generic_kernel<FBA_LRELU_REF, T><<<grid, 256, 0, st>>>(out, x, bias, ref, alpha, scale, size_x, stb, sb); break;
    default: fba_generic_kernel<FBA_ZERO, T><<<grid, 256, 0, st>>>(out, x, bias, ref, alpha, scale, size_x, stb, sb); break;
  }
  MSG_CHECK_LAUNCH("fused_bias_act(generic)");
  return MSG_OK;
}

}  // namespace msg

using namespace msg;

extern "C" int msg_fused_bias_act(void* out, const void* x, const void* bias, const void* ref, int act,
                                  int grad, double alpha, double scale, int64_t size_x, int64_t step_b,
                                  int64_t size_b, int dtype, msg_stream_t stream) {
  if (size_x < 0 || step_b < 0 || size_b < 0) return fail(MSG_ERR_BAD_ARG, "fused_bias_act: negative size");
  if (size_x == 0) return MSG_OK;
  if (!out || !x) return fail(MSG_ERR_BAD_ARG, "fused_bias_act: null out/x");
  const int mode = fba_mode(act, grad);
  if (mode == FBA_LRELU_REF && !ref) return fail(MSG_ERR_BAD_ARG, "fused_bias_act: grad=1 needs ref");
  if (bias && size_b <= 0) return fail(MSG_ERR_BAD_ARG, "fused_bias_act: bias with size_b<=0");
  if (bias && step_b <= 0) return fail(MSG_ERR_BAD_ARG, "fused_bias_act: bias with step_b<=0");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == MSG_F32)
    return launch_fba<float>((float*)out, (const float*)x, (const float*)bias, (const float*)ref, mode,
                             (float)alpha, (float)scale, size_x, step_b, size_b, nullptr, nullptr, st);
  if (dtype == MSG_F64)
    return launch_fba<double>((double*)out, (const double*)x, (const double*)bias, (const double*)ref, mode,
                              alpha, scale, size_x, step_b, size_b, nullptr, nullptr, st);
  return fail(MSG_ERR_UNSUPPORTED, "fused_bias_act: dtype %d", dtype);
}

extern "C" size_t msg_fused_bias_act_bwd_workspace(int64_t size_x, int64_t step_b, int64_t size_b, int dtype) {
  if (size_x <= 0 || step_b < kPlanarMinStep || size_x % step_b) return 16;
  const size_t es = dtype == MSG_F64 ? 8 : 4;
  const int64_t planes = size_x / step_b;
  // worst case: scalar path
  const int64_t nchunks = ceil_div(step_b, 256 * kPlanarLoop);
  (void)size_b;
  return (size_t)(planes * nchunks) * es + 16;
}

template <typename T>
static int fba_bwd_impl(T* dx, T* dbias, const T* g, const T* ref, T alpha, T scale, int64_t size_x,
                        int64_t step_b, int64_t size_b, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int64_t planes = size_x / step_b;
  const int64_t outer = planes / size_b;
  const bool planar = step_b >= kPlanarMinStep;
  if (planar) {
    const size_t need = msg_fused_bias_act_bwd_workspace(size_x, step_b, size_b, sizeof(T) == 8 ? MSG_F64 : MSG_F32);
    if (!ws || ws_bytes < need) return fail(MSG_ERR_WORKSPACE, "fused_bias_act_bwd: workspace %zu < %zu", ws_bytes, need);
    int nchunks = 0;
    int rc = launch_fba<T>(dx, g, nullptr, ref, FBA_LRELU_REF, alpha, scale, size_x, step_b, size_b, (T*)ws, &nchunks, st);
    if (rc) return rc;
    fba_reduce_partials_kernel<T><<<(unsigned)size_b, 256, 0, st>>>(dbias, (const T*)ws, outer, (int)size_b, nchunks);
    MSG_CHECK_LAUNCH("fused_bias_act_bwd(reduce)");
    return MSG_OK;
  }
  int rc = launch_fba<T>(dx, g, nullptr, ref, FBA_LRELU_REF, alpha, scale, size_x, step_b, size_b, nullptr, nullptr, st);
  if (rc) return rc;
  fba_reduce_direct_kernel<T><<<(unsigned)size_b, 256, 0, st>>>(dbias, dx, outer, (int)size_b, step_b);
  MSG_CHECK_LAUNCH("fused_bias_act_bwd(reduce-direct)");
  return MSG_OK;
}

extern "C" int msg_fused_bias_act_bwd(void* dx, void* dbias, const void* g, const void* ref, double alpha,
                                      double scale, int64_t size_x, int64_t step_b, int64_t size_b,
                                      void* workspace, size_t workspace_bytes, int dtype,
                                      msg_stream_t stream) {
  if (size_x < 0 || step_b <= 0 || size_b <= 0) return fail(MSG_ERR_BAD_ARG, "fused_bias_act_bwd: bad sizes");
  if (size_x % (step_b * size_b)) return fail(MSG_ERR_BAD_ARG, "fused_bias_act_bwd: size_x not [outer,size_b,step_b]");
  if (!dbias) return fail(MSG_ERR_BAD_ARG, "fused_bias_act_bwd: null dbias");
  cudaStream_t st = (cudaStream_t)stream;
  if (size_x == 0) {
    MSG_CHECK_CUDA(cudaMemsetAsync(dbias, 0, (size_t)size_b * (dtype == MSG_F64 ? 8 : 4), st));
    return MSG_OK;
  }
  if (!dx || !g || !ref) return fail(MSG_ERR_BAD_ARG, "fused_bias_act_bwd: null pointer");
  if (dtype == MSG_F32)
    return fba_bwd_impl<float>((float*)dx, (float*)dbias, (const float*)g, (const float*)ref, (float)alpha,
                               (float)scale, size_x, step_b, size_b, workspace, workspace_bytes, st);
  if (dtype == MSG_F64)
    return fba_bwd_impl<double>((double*)dx, (double*)dbias, (const double*)g, (const double*)ref, alpha, scale,
                                size_x, step_b, size_b, workspace, workspace_bytes, st);
  return fail(MSG_ERR_UNSUPPORTED, "fused_bias_act_bwd: dtype %d", dtype);
}
