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


// ---- flat kernel: no bias, no reduction -> the layout is irrelevant ---------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
fba_flat_kernel(float4* __restrict__ out, const float4* __restrict__ x, const float4* __restrict__ ref,
                float alpha, float scale, int64_t n4) {
  constexpr int LOOP = 4;
  int64_t i0 = ((int64_t)blockIdx.x * LOOP) * 256 + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * LOOP * 256;
  for (; i0 < n4; i0 += stride) {
    float4 xv[LOOP], rv[LOOP];
#pragma unroll
    for (int l = 0; l < LOOP; ++l) {
      const int64_t i = i0 + l * 256;
      if (i < n4) {
        xv[l] = __ldg(x + i);
        if (MODE == FBA_LRELU_REF) rv[l] = __ldg(ref + i);
      }
    }
#pragma unroll
    for (int l = 0; l < LOOP; ++l) {
      const int64_t i = i0 + l * 256;
      if (i < n4) {
        const float4 r = (MODE == FBA_LRELU_REF) ? rv[l] : make_float4(0, 0, 0, 0);
        float4 o;
        o.x = fba_apply<MODE, float>(xv[l].x, r.x, alpha, scale);
        o.y = fba_apply<MODE, float>(xv[l].y, r.y, alpha, scale);
        o.z = fba_apply<MODE, float>(xv[l].z, r.z, alpha, scale);
        o.w = fba_apply<MODE, float>(xv[l].w, r.w, alpha, scale);
        out[i] = o;
      }
    }
  }
}

// ---- channel-inner kernel: x is [rows, C] (channels-last activations, or [N, C]); C % 4 == 0 -------------
// grid.x = chunks of 32 channel quads (one warp row = 512 contiguous bytes), grid.y = row blocks.
// WITH_SUM: partial[blockIdx.y][c] = sum over the block's rows of out (deterministic two-stage reduction).
// Optional per-row term (noise injection, multi_stylegan_generator.py:292): v = x + bias[c] + rowv_w[0] * rowv[row % period];
// with WITH_SUM its weight gradient partial_n[block] = sum rowv[row] * out is produced as well.
template <int MODE, bool WITH_SUM>
__global__ void __launch_bounds__(256)
fba_inner_kernel(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ bias,
                 const float* __restrict__ ref, float alpha, float scale, int64_t rows, int C,
                 int64_t rows_per_block, float* __restrict__ partial,
                 const float* __restrict__ rowv, const float* __restrict__ rowv_w, int64_t rowv_period,
                 float* __restrict__ partial_n) {
  constexpr int LOOP = 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C4 = C >> 2;
  const int q = blockIdx.x * 32 + lane;
  const bool qok = q < C4;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float4 b = make_float4(0, 0, 0, 0);
  if (bias && qok) b = __ldg(reinterpret_cast<const float4*>(bias) + q);
  float4 acc = make_float4(0, 0, 0, 0);
  float nacc = 0.f;
  const float nw = (rowv && rowv_w) ? __ldg(rowv_w) : 0.f;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const float4* r4 = reinterpret_cast<const float4*>(ref);
  float4* o4 = reinterpret_cast<float4*>(out);
  for (int64_t r = r0 + warp; r < r1; r += 8 * LOOP) {
    float4 xv[LOOP], rv[LOOP];
    float nz[LOOP];
#pragma unroll
    for (int l = 0; l < LOOP; ++l) {
      const int64_t rr = r + 8 * l;
      nz[l] = 0.f;
      if (qok && rr < r1) {
        xv[l] = __ldg(x4 + rr * C4 + q);
        if (MODE == FBA_LRELU_REF) rv[l] = __ldg(r4 + rr * C4 + q);
        if (rowv) {
          int64_t ni = rr;
          if (rr >= rowv_period)
            ni = (rows <= 0x7fffffffLL) ? (int64_t)((uint32_t)rr % (uint32_t)rowv_period) : rr % rowv_period;
          nz[l] = __ldg(rowv + ni);
        }
      }
    }
#pragma unroll
    for (int l = 0; l < LOOP; ++l) {
      const int64_t rr = r + 8 * l;
      if (qok && rr < r1) {
        const float4 rf = (MODE == FBA_LRELU_REF) ? rv[l] : make_float4(0, 0, 0, 0);
        const float add = nw * nz[l];
        float4 o;
        o.x = fba_apply<MODE, float>(xv[l].x + add + b.x, rf.x, alpha, scale);
        o.y = fba_apply<MODE, float>(xv[l].y + add + b.y, rf.y, alpha, scale);
        o.z = fba_apply<MODE, float>(xv[l].z + add + b.z, rf.z, alpha, scale);
        o.w = fba_apply<MODE, float>(xv[l].w + add + b.w, rf.w, alpha, scale);
        if (WITH_SUM) {
          acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
          nacc = fmaf(nz[l], (o.x + o.y) + (o.z + o.w), nacc);
        }
        o4[rr * C4 + q] = o;
      }
    }
  }
  if (WITH_SUM && partial_n) {
    __shared__ float nscratch[32];
    const float t = block_sum(nacc, nscratch);
    if (threadIdx.x == 0) partial_n[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
  if (WITH_SUM) {
    __shared__ float4 sacc[8][32];
    sacc[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && qok) {
      float4 t = sacc[0][lane];
#pragma unroll
      for (int w = 1; w < 8; ++w) {
        const float4 u = sacc[w][lane];
        t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
      }
      reinterpret_cast<float4*>(partial + (int64_t)blockIdx.y * C)[q] = t;
    }
  }
}

// dbias[c] = sum_j partial[j][c];  dnw[0] = sum_j partial_n[j]  (last block, fixed order: deterministic)
__global__ void __launch_bounds__(256)
fba_reduce_rows_kernel(float* __restrict__ dbias, const float* __restrict__ partial, int nblocks, int C,
                       float* __restrict__ dnw, const float* __restrict__ partial_n, int n_partial_n) {
  if (blockIdx.x == gridDim.x - 1 && dnw) {
    float acc = 0.f;
    for (int j = threadIdx.x; j < n_partial_n; j += blockDim.x) acc += partial_n[j];
    __shared__ float scratch[32];
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) dnw[0] = acc;
    return;
  }
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C || !dbias) return;
  float acc = 0.f;
  for (int j = 0; j < nblocks; ++j) acc += partial[(int64_t)j * C + c];
  dbias[c] = acc;
}

static inline int fba_inner_row_blocks(int64_t rows, int C) {
  const int gx = (int)ceil_div(C >> 2, 32);
  int64_t gy = ((int64_t)num_sms() * 8) / gx;
  const int64_t max_by_rows = ceil_div(rows, 32);
  if (gy > max_by_rows) gy = max_by_rows;
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  return (int)gy;
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
  if constexpr (sizeof(T) == 4) {
    const bool al = aligned16(x) && aligned16(out) && (mode != FBA_LRELU_REF || aligned16(ref));
    // (a) no bias, no reduction: flat 128-bit pass whatever the layout
    if (!has_bias && !partial && al && (size_x % 4 == 0)) {
      const int64_t n4 = size_x >> 2;
      const int64_t want = ceil_div(n4, 256 * 4);
      const unsigned grid = (unsigned)(want < (int64_t)num_sms() * 16 ? want : (int64_t)num_sms() * 16);
      const float4* r4 = reinterpret_cast<const float4*>(ref);
      switch (mode) {
        case FBA_LINEAR: fba_flat_kernel<FBA_LINEAR><<<grid, 256, 0, st>>>((float4*)out, (const float4*)x, r4, alpha, scale, n4); break;
        case FBA_LRELU_X: fba_flat_kernel<FBA_LRELU_X><<<grid, 256, 0, st>>>((float4*)out, (const float4*)x, r4, alpha, scale, n4); break;
        case FBA_LRELU_REF: fba_flat_kernel<FBA_LRELU_REF><<<grid, 256, 0, st>>>((float4*)out, (const float4*)x, r4, alpha, scale, n4); break;
        default: fba_flat_kernel<FBA_ZERO><<<grid, 256, 0, st>>>((float4*)out, (const float4*)x, r4, alpha, scale, n4); break;
      }
      MSG_CHECK_LAUNCH("fused_bias_act(flat)");
      return MSG_OK;
    }
    // (b) channel-inner layout ([rows, C]: channels-last activations or [N, C] matrices)
    if (step_b == 1 && size_b > 0 && size_b % 4 == 0 && size_b <= (1 << 20) && (size_x % size_b == 0) && al &&
        (!has_bias || aligned16(bias))) {
      const int C = (int)size_b;
      const int64_t rows = size_x / C;
      const int gy = fba_inner_row_blocks(rows, C);
      const int64_t rpb = ceil_div(rows, gy);
      if (nchunks_out) *nchunks_out = gy;
      dim3 grid((unsigned)ceil_div(C >> 2, 32), (unsigned)gy);
#define FBA_INNER(MODE)                                                                                         \
  if (partial) fba_inner_kernel<MODE, true><<<grid, 256, 0, st>>>(out, x, bias, ref, alpha, scale, rows, C, rpb, partial, nullptr, nullptr, 1, nullptr); \
  else fba_inner_kernel<MODE, false><<<grid, 256, 0, st>>>(out, x, bias, ref, alpha, scale, rows, C, rpb, partial, nullptr, nullptr, 1, nullptr)
      switch (mode) {
        case FBA_LINEAR: FBA_INNER(FBA_LINEAR); break;
        case FBA_LRELU_X: FBA_INNER(FBA_LRELU_X); break;
        case FBA_LRELU_REF: FBA_INNER(FBA_LRELU_REF); break;
        default: FBA_INNER(FBA_ZERO); break;
      }
#undef FBA_INNER
      MSG_CHECK_LAUNCH("fused_bias_act(channel-inner)");
      return MSG_OK;
    }
  }
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
  if (size_x > 0 && step_b == 1 && size_b > 0 && size_b % 4 == 0 && dtype == MSG_F32 && size_x % size_b == 0)
    return (size_t)fba_inner_row_blocks(size_x / size_b, (int)size_b) * (size_t)size_b * 4 + 16;
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
  if constexpr (sizeof(T) == 4) {
    if (step_b == 1 && size_b % 4 == 0 && aligned16(dx) && aligned16(g) && aligned16(ref) && ws &&
        (reinterpret_cast<uintptr_t>(ws) & 15u) == 0) {
      const size_t need = msg_fused_bias_act_bwd_workspace(size_x, step_b, size_b, MSG_F32);
      if (ws_bytes < need) return fail(MSG_ERR_WORKSPACE, "fused_bias_act_bwd: workspace %zu < %zu", ws_bytes, need);
      int nblocks = 0;
      int rc = launch_fba<T>(dx, g, nullptr, ref, FBA_LRELU_REF, alpha, scale, size_x, step_b, size_b, (T*)ws, &nblocks, st);
      if (rc) return rc;
      fba_reduce_rows_kernel<<<(unsigned)ceil_div(size_b, 256), 256, 0, st>>>(dbias, (const float*)ws, nblocks, (int)size_b,
                                                                              nullptr, nullptr, 0);
      MSG_CHECK_LAUNCH("fused_bias_act_bwd(reduce-rows)");
      return MSG_OK;
    }
  }
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


// ---- StyledConv2d epilogue on channels-last activations ------------------------------------------------------
static inline bool nba_ok(const void* a, const void* b, const void* c) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15u) == 0;
}

extern "C" int msg_noise_bias_act_nhwc(float* out, const float* x, const float* ref, const float* noise,
                                       const float* noise_w, const float* bias, int64_t rows, int C,
                                       int64_t noise_period, float alpha, float scale, msg_stream_t stream) {
  if (rows < 0 || C <= 0) return fail(MSG_ERR_BAD_ARG, "noise_bias_act_nhwc: bad sizes");
  if (rows == 0) return MSG_OK;
  if (!out || !x) return fail(MSG_ERR_BAD_ARG, "noise_bias_act_nhwc: null pointer");
  if (C % 4) return fail(MSG_ERR_UNSUPPORTED, "noise_bias_act_nhwc: C must be a multiple of 4");
  if (noise && (!noise_w || noise_period <= 0)) return fail(MSG_ERR_BAD_ARG, "noise_bias_act_nhwc: noise needs noise_w and a period");
  if (!nba_ok(out, x, ref) || (bias && (reinterpret_cast<uintptr_t>(bias) & 15u)))
    return fail(MSG_ERR_BAD_ARG, "noise_bias_act_nhwc: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int gy = fba_inner_row_blocks(rows, C);
  const int64_t rpb = ceil_div(rows, gy);
  dim3 grid((unsigned)ceil_div(C >> 2, 32), (unsigned)gy);
  if (ref)
    fba_inner_kernel<FBA_LRELU_REF, false><<<grid, 256, 0, st>>>(out, x, bias, ref, alpha, scale, rows, C, rpb, nullptr,
                                                                 noise, noise_w, noise ? noise_period : 1, nullptr);
  else
    fba_inner_kernel<FBA_LRELU_X, false><<<grid, 256, 0, st>>>(out, x, bias, nullptr, alpha, scale, rows, C, rpb, nullptr,
                                                               noise, noise_w, noise ? noise_period : 1, nullptr);
  MSG_CHECK_LAUNCH("noise_bias_act_nhwc");
  return MSG_OK;
}

extern "C" size_t msg_noise_bias_act_nhwc_bwd_workspace(int64_t rows, int C) {
  if (rows <= 0 || C <= 0 || C % 4) return 16;
  const int gy = fba_inner_row_blocks(rows, C);
  const size_t gx = (size_t)ceil_div(C >> 2, 32);
  return ((size_t)gy * C + (size_t)gy * gx) * sizeof(float) + 32;
}

extern "C" int msg_noise_bias_act_nhwc_bwd(float* dx, float* dbias, float* dnoise_w, const float* g, const float* ref,
                                           const float* noise, int64_t rows, int C, int64_t noise_period, float alpha,
                                           float scale, void* workspace, size_t workspace_bytes, msg_stream_t stream) {
  if (rows < 0 || C <= 0) return fail(MSG_ERR_BAD_ARG, "noise_bias_act_nhwc_bwd: bad sizes");
  if (C % 4) return fail(MSG_ERR_UNSUPPORTED, "noise_bias_act_nhwc_bwd: C must be a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    if (dbias) MSG_CHECK_CUDA(cudaMemsetAsync(dbias, 0, (size_t)C * 4, st));
    if (dnoise_w) MSG_CHECK_CUDA(cudaMemsetAsync(dnoise_w, 0, 4, st));
    return MSG_OK;
  }
  if (!dx || !g || !ref) return fail(MSG_ERR_BAD_ARG, "noise_bias_act_nhwc_bwd: null pointer");
  if (dnoise_w && (!noise || noise_period <= 0)) return fail(MSG_ERR_BAD_ARG, "noise_bias_act_nhwc_bwd: dnoise_w needs noise");
  if (!nba_ok(dx, g, ref)) return fail(MSG_ERR_BAD_ARG, "noise_bias_act_nhwc_bwd: pointers must be 16-byte aligned");
  const size_t need = msg_noise_bias_act_nhwc_bwd_workspace(rows, C);
  if (!workspace || workspace_bytes < need)
    return fail(MSG_ERR_WORKSPACE, "noise_bias_act_nhwc_bwd: workspace %zu < %zu", workspace_bytes, need);
  float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 15) & ~(uintptr_t)15);
  const int gy = fba_inner_row_blocks(rows, C);
  const int gx = (int)ceil_div(C >> 2, 32);
  float* partial_n = partial + (size_t)gy * C;
  const int64_t rpb = ceil_div(rows, gy);
  dim3 grid((unsigned)gx, (unsigned)gy);
  if (!dbias && !dnoise_w) {
    // no parameter gradient wanted (the discriminator inside the generator step): the plain masked pass, no reduction
    fba_inner_kernel<FBA_LRELU_REF, false><<<grid, 256, 0, st>>>(dx, g, nullptr, ref, alpha, scale, rows, C, rpb, nullptr,
                                                                 nullptr, nullptr, 1, nullptr);
    MSG_CHECK_LAUNCH("noise_bias_act_nhwc_bwd(dx only)");
    return MSG_OK;
  }
  // dx = mask(ref) * g * scale; the per-row term only enters through its gradient (rowv_w == nullptr -> adds 0)
  fba_inner_kernel<FBA_LRELU_REF, true><<<grid, 256, 0, st>>>(dx, g, nullptr, ref, alpha, scale, rows, C, rpb, partial,
                                                              dnoise_w ? noise : nullptr, nullptr,
                                                              dnoise_w ? noise_period : 1, dnoise_w ? partial_n : nullptr);
  MSG_CHECK_LAUNCH("noise_bias_act_nhwc_bwd");
  fba_reduce_rows_kernel<<<(unsigned)ceil_div(C, 256) + 1, 256, 0, st>>>(dbias, partial, gy, C, dnoise_w, partial_n, gy * gx);
  MSG_CHECK_LAUNCH("noise_bias_act_nhwc_bwd(reduce)");
  return MSG_OK;
}
