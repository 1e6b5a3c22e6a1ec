// Small-M linears of the generator: the style-mapping network and the style (modulation) linears.
//
// Reference: EqualizedLinear (equalized_layer.py:210-254)  y = x (W * sqrt(2)/sqrt(K))^T + b * sqrt(2)/sqrt(N),
// PixelwiseNormalization (:257-277) and StyleMapping (multi_stylegan_generator.py:208-235: pixel norm, then `depth` x
// [EqualizedLinear(bias=False) -> FusedLeakyReLU]); the 26 style linears are the `modulation_mapping` of every
// ModulatedConv2d (:355-361).  All of them multiply M = batch (8 ... 32) rows with 512 x 512 weights: GEMV-shaped,
// bound by streaming the weights once, far below anything a 128-row tensor-core tile could use.  So:
//
//   * msg_style_mapping_forward / _backward: the whole mapping network as ONE launch.  A cluster of 8 CTAs keeps the
//     activations of its (up to 16) rows in shared memory; every CTA owns 1/8 of the output columns of each layer, reads
//     that slice of the weights with warp-wide 512-byte loads, and broadcasts its outputs into the next-layer buffer of
//     all 8 CTAs through distributed shared memory; one cluster barrier per layer.  The backward walks the layers in
//     reverse in the same way (dW rows and the bias gradient of the owned columns, the input gradient of the owned
//     input columns), no atomics: bit-reproducible.
//   * msg_linear_group_forward / _wgrad / _dgrad: any number of independent linears that read slices of ONE [M, R] input
//     (the per-layer latents, or the styles of the first branch) in one launch each, driven by an item table that travels in the kernel parameters.
//
// fp32 FMA arithmetic (the reference's cuBLAS runs these in TF32 under PyTorch 1.8.1 defaults).
#include <cstring>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace msg {
using namespace ptx;

constexpr int kMapRows = 16;      // rows per cluster pass
constexpr int kMapCluster = 8;
constexpr int kMaxDepth = 16;

struct MapParams {
  const float* z;                 // [M, K]
  const float* W[kMaxDepth];      // [K, K] each (out x in)
  const float* bias[kMaxDepth];   // [K] or null (FusedLeakyReLU bias, unscaled)
  int L, M, K;
  float alpha, slope, gain, eps;
  float* x0;                      // [M, K] normalised input (saved for the backward)
  float* acts;                    // [L, M, K] layer outputs (the last one is the result)
  // backward
  const float* gy;                // [M, K]
  float* dW;                      // [L, K, K]
  float* db;                      // [L, K]
};

__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

__global__ void __cluster_dims__(kMapCluster, 1, 1) __launch_bounds__(256, 1)
style_mapping_forward_kernel(const MapParams p) {
  extern __shared__ float4 smem_f4[];
  float* buf = reinterpret_cast<float*>(smem_f4);                 // [2][kMapRows][K]
  const int K = p.K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const int row0 = (int)(blockIdx.x / kMapCluster) * kMapRows;
  const int rows = min(kMapRows, p.M - row0);
  const int nc = (K + kMapCluster - 1) / kMapCluster;
  const int n_begin = (int)rank * nc, n_end = min(K, n_begin + nc);

  // pixel norm of the input rows (every CTA builds its own full copy)
  for (int r = warp; r < rows; r += 8) {
    const float* zr = p.z + (int64_t)(row0 + r) * K;
    float ss = 0.f;
    for (int k = lane; k < K; k += 32) { const float v = __ldg(zr + k); ss = fmaf(v, v, ss); }
    ss = warp_sum(ss);
    const float inv = rsqrtf(ss / (float)K + p.eps);
    for (int k = lane; k < K; k += 32) {
      const float v = __ldg(zr + k) * inv;
      buf[r * K + k] = v;
      if (rank == 0) p.x0[(int64_t)(row0 + r) * K + k] = v;
    }
  }
  __syncthreads();
  cluster_sync_all();          // nobody writes into a peer's buffer before that peer has finished its prologue

  const uint32_t buf_s = smem_u32(buf);
  for (int l = 0; l < p.L; ++l) {
    const float* cur = buf + (l & 1) * kMapRows * K;
    const uint32_t nxt_s = buf_s + (uint32_t)(((l + 1) & 1) * kMapRows * K) * 4u;
    const float* W = p.W[l];
    const float* bias = p.bias[l];
    for (int n = n_begin + warp; n < n_end; n += 8) {
      float acc[kMapRows];
#pragma unroll
      for (int r = 0; r < kMapRows; ++r) acc[r] = 0.f;
      const float* wr = W + (int64_t)n * K;
      for (int k = lane * 4; k < K; k += 128) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
        for (int r = 0; r < kMapRows; ++r) {
          if (r < rows) {
            const float4 x4 = *reinterpret_cast<const float4*>(cur + r * K + k);
            acc[r] = fmaf(w4.x, x4.x, fmaf(w4.y, x4.y, fmaf(w4.z, x4.z, fmaf(w4.w, x4.w, acc[r]))));
          }
        }
      }
      float mine = 0.f;
#pragma unroll
      for (int r = 0; r < kMapRows; ++r) {
        const float s = warp_sum(acc[r]);
        if (lane == r) mine = s;
      }
      if (lane < rows) {
        float v = p.alpha * mine + (bias ? __ldg(bias + n) : 0.f);
        v = (v > 0.f ? v : v * p.slope) * p.gain;
        p.acts[((int64_t)l * p.M + row0 + lane) * K + n] = v;
        const uint32_t a = nxt_s + (uint32_t)(lane * K + n) * 4u;
#pragma unroll
        for (uint32_t c = 0; c < kMapCluster; ++c) st_cluster_f32(mapa_cluster(a, c), v);
      }
    }
    cluster_sync_all();
  }
}

__global__ void __cluster_dims__(kMapCluster, 1, 1) __launch_bounds__(256, 1)
style_mapping_backward_kernel(const MapParams p) {
  extern __shared__ float4 smem_f4[];
  const int K = p.K;
  float* gbuf = reinterpret_cast<float*>(smem_f4);          // [2][kMapRows][K]  gradient w.r.t. a layer output
  float* gp = gbuf + 2 * kMapRows * K;                      // [kMapRows][K]     gradient w.r.t. the pre-activation
  float* xin = gp + kMapRows * K;                           // [kMapRows][K]     the layer's input
  float* red = xin + kMapRows * K;                          // [4][kMapRows][64]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const int nc = (K + kMapCluster - 1) / kMapCluster;       // <= 64 (host check)
  const int n_begin = (int)rank * nc, n_end = min(K, n_begin + nc);
  const uint32_t gbuf_s = smem_u32(gbuf);

  for (int row0 = 0, chunk = 0; row0 < p.M; row0 += kMapRows, ++chunk) {
    const int rows = min(kMapRows, p.M - row0);
    for (int i = threadIdx.x; i < rows * K; i += 256) gbuf[i] = __ldg(p.gy + (int64_t)row0 * K + i);
    __syncthreads();
    cluster_sync_all();
    for (int l = p.L - 1, step = 0; l >= 0; --l, ++step) {
      const float* g = gbuf + (step & 1) * kMapRows * K;
      const uint32_t nxt_s = gbuf_s + (uint32_t)(((step + 1) & 1) * kMapRows * K) * 4u;
      const float* act = p.acts + ((int64_t)l * p.M + row0) * K;
      const float* inp = (l > 0 ? p.acts + ((int64_t)(l - 1) * p.M + row0) * K : p.x0 + (int64_t)row0 * K);
      for (int i = threadIdx.x; i < rows * K; i += 256) {
        const float o = __ldg(act + i);
        gp[i] = g[i] * (o > 0.f ? 1.f : p.slope) * p.gain;
        xin[i] = __ldg(inp + i);
      }
      __syncthreads();
      // weight and bias gradients of the owned output columns
      float* dW = p.dW + (int64_t)l * K * K;
      for (int n = n_begin + warp; n < n_end; n += 8) {
        float gn[kMapRows];
        float bsum = 0.f;
#pragma unroll
        for (int r = 0; r < kMapRows; ++r) { gn[r] = r < rows ? gp[r * K + n] : 0.f; bsum += gn[r]; }
        for (int k = lane * 4; k < K; k += 128) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int r = 0; r < kMapRows; ++r) {
            if (r < rows) {
              const float4 x4 = *reinterpret_cast<const float4*>(xin + r * K + k);
              a.x = fmaf(gn[r], x4.x, a.x); a.y = fmaf(gn[r], x4.y, a.y); a.z = fmaf(gn[r], x4.z, a.z); a.w = fmaf(gn[r], x4.w, a.w);
            }
          }
          float4* dst = reinterpret_cast<float4*>(dW + (int64_t)n * K + k);
          a.x *= p.alpha; a.y *= p.alpha; a.z *= p.alpha; a.w *= p.alpha;
          if (chunk > 0) { const float4 o = *dst; a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w; }
          *dst = a;
        }
        if (lane == 0 && p.bias[l] != nullptr) {
          float* d = p.db + (int64_t)l * K + n;
          *d = chunk > 0 ? *d + bsum : bsum;
        }
      }
      // gradient w.r.t. the layer input, owned input columns: g_prev[r][k] = alpha * sum_n gp[r][n] * W[n][k]
      if (l > 0) {
        const int kl = threadIdx.x & 63, ng = threadIdx.x >> 6;
        const int k = n_begin + kl;
        float acc[kMapRows];
#pragma unroll
        for (int r = 0; r < kMapRows; ++r) acc[r] = 0.f;
        if (kl < nc && k < K) {
          const float* W = p.W[l];
          const int per = (K + 3) / 4;
          const int n1 = min(K, (ng + 1) * per);
          for (int n = ng * per; n < n1; ++n) {
            const float w = __ldg(W + (int64_t)n * K + k);
#pragma unroll
            for (int r = 0; r < kMapRows; ++r) acc[r] = fmaf(gp[r * K + n], w, acc[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < kMapRows; ++r) red[(ng * kMapRows + r) * 64 + kl] = acc[r];
        __syncthreads();
        for (int i = threadIdx.x; i < rows * 64; i += 256) {
          const int r = i >> 6, kk = i & 63;
          if (kk < nc && n_begin + kk < K) {
            const float v = p.alpha * (red[(0 * kMapRows + r) * 64 + kk] + red[(1 * kMapRows + r) * 64 + kk] +
                                       red[(2 * kMapRows + r) * 64 + kk] + red[(3 * kMapRows + r) * 64 + kk]);
            const uint32_t a = nxt_s + (uint32_t)(r * K + n_begin + kk) * 4u;
#pragma unroll
            for (uint32_t c = 0; c < kMapCluster; ++c) st_cluster_f32(mapa_cluster(a, c), v);
          }
        }
      }
      __syncthreads();
      cluster_sync_all();
    }
  }
}

// ---- grouped linears --------------------------------------------------------------------------------------------------
struct LinItem {          // one linear of a group: out_i [M, N] = alpha * in[:, in_off : in_off + K] W^T + beta * b, stored as
                          // the contiguous block [out_off * M, (out_off + N) * M) of the flat output (out_off = sum of the
                          // preceding items' N)
  const float* W;         // [N, K]
  const float* bias;      // [N] or null
  int N, K;
  int in_off, out_off;    // float offsets inside a row of the input / output
  int w_off, b_off;       // float offsets of this item's dW / db inside the flat gradient buffers
  float alpha, beta;
};
struct LinSlot {          // items that read the same input slice (their input gradients add up)
  int in_off, K, first, count;
};
constexpr int kMaxItems = 48;       // the tables travel in the kernel parameters (48 x 48 + 48 x 16 bytes < 4 KB)
struct LinGroupParams {
  LinItem items[kMaxItems];
  LinSlot slots[kMaxItems];
  int n_items, n_slots, M;
  const float* in;        // [M, R]
  int64_t R;
  float* out;             // flat, item-major (forward)
  const float* gout;      // the same layout (backward)
  float* dW;              // flat
  float* db;              // flat
  float* din;             // [M, R]
};

constexpr int kGroupRows = 16;

__global__ void __launch_bounds__(256)
linear_group_forward_kernel(const LinGroupParams p) {
  extern __shared__ float4 smem_f4[];
  float* xs = reinterpret_cast<float*>(smem_f4);                 // [kGroupRows][K]
  const LinItem it = p.items[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 8 + warp;
  if (blockIdx.x * 8 >= it.N) return;
  for (int row0 = 0; row0 < p.M; row0 += kGroupRows) {
    const int rows = min(kGroupRows, p.M - row0);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * it.K; i += 256) {
      const int r = i / it.K, k = i - r * it.K;
      xs[i] = __ldg(p.in + (int64_t)(row0 + r) * p.R + it.in_off + k);
    }
    __syncthreads();
    if (n < it.N) {
      float acc[kGroupRows];
#pragma unroll
      for (int r = 0; r < kGroupRows; ++r) acc[r] = 0.f;
      const float* wr = it.W + (int64_t)n * it.K;
      for (int k = lane * 4; k < it.K; k += 128) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
        for (int r = 0; r < kGroupRows; ++r) {
          if (r < rows) {
            const float4 x4 = *reinterpret_cast<const float4*>(xs + r * it.K + k);
            acc[r] = fmaf(w4.x, x4.x, fmaf(w4.y, x4.y, fmaf(w4.z, x4.z, fmaf(w4.w, x4.w, acc[r]))));
          }
        }
      }
      float mine = 0.f;
#pragma unroll
      for (int r = 0; r < kGroupRows; ++r) {
        const float s = warp_sum(acc[r]);
        if (lane == r) mine = s;
      }
      if (lane < rows)
        p.out[(int64_t)it.out_off * p.M + (int64_t)(row0 + lane) * it.N + n] = it.alpha * mine + (it.bias ? it.beta * __ldg(it.bias + n) : 0.f);
    }
  }
}

// warp = one output row n of one item: dW[n, :] = alpha * sum_m g[m, n] * x[m, :],  db[n] = beta * sum_m g[m, n]
__global__ void __launch_bounds__(256)
linear_group_wgrad_kernel(const LinGroupParams p) {
  const LinItem it = p.items[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 8 + warp;
  if (n >= it.N) return;
  float bsum = 0.f;
  for (int k0 = 0; k0 < it.K; k0 += 512) {             // 4 float4 accumulators per lane cover 512 input columns
    float4 a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = 0; m < p.M; ++m) {
      const float g = __ldg(p.gout + (int64_t)it.out_off * p.M + (int64_t)m * it.N + n);
      if (k0 == 0) bsum += g;
      const float* xr = p.in + (int64_t)m * p.R + it.in_off + k0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = lane * 4 + j * 128;
        if (k0 + k < it.K) {
          const float4 x4 = __ldg(reinterpret_cast<const float4*>(xr + k));
          a[j].x = fmaf(g, x4.x, a[j].x); a[j].y = fmaf(g, x4.y, a[j].y); a[j].z = fmaf(g, x4.z, a[j].z); a[j].w = fmaf(g, x4.w, a[j].w);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lane * 4 + j * 128;
      if (k < it.K)
        *reinterpret_cast<float4*>(p.dW + it.w_off + (int64_t)n * it.K + k) =
            make_float4(a[j].x * it.alpha, a[j].y * it.alpha, a[j].z * it.alpha, a[j].w * it.alpha);
    }
  }
  if (lane == 0 && it.bias != nullptr) p.db[it.b_off + n] = it.beta * bsum;
}

// block = 64 input columns of one slot: din[m, in_off + k] = sum over the slot's items: alpha * sum_n g[m, out_off + n] W[n, k]
__global__ void __launch_bounds__(256)
linear_group_dgrad_kernel(const LinGroupParams p) {
  __shared__ float red[4][kGroupRows][64];
  extern __shared__ float4 smem_f4[];
  float* gs = reinterpret_cast<float*>(smem_f4);                 // [kGroupRows][Nmax]
  const LinSlot sl = p.slots[blockIdx.y];
  const int k0 = blockIdx.x * 64;
  if (k0 >= sl.K) return;
  const int kl = threadIdx.x & 63, ng = threadIdx.x >> 6;
  const int k = k0 + kl;
  for (int row0 = 0; row0 < p.M; row0 += kGroupRows) {
    const int rows = min(kGroupRows, p.M - row0);
    float acc[kGroupRows];
#pragma unroll
    for (int r = 0; r < kGroupRows; ++r) acc[r] = 0.f;
    for (int ii = 0; ii < sl.count; ++ii) {
      const LinItem it = p.items[sl.first + ii];
      __syncthreads();
      for (int i = threadIdx.x; i < kGroupRows * it.N; i += 256) {
        const int r = i / it.N, n = i - r * it.N;
        gs[i] = r < rows ? it.alpha * __ldg(p.gout + (int64_t)it.out_off * p.M + (int64_t)(row0 + r) * it.N + n) : 0.f;
      }
      __syncthreads();
      if (k < sl.K) {
        const int per = (it.N + 3) / 4;
        const int n1 = min(it.N, (ng + 1) * per);
        for (int n = ng * per; n < n1; ++n) {
          const float w = __ldg(it.W + (int64_t)n * it.K + k);
#pragma unroll
          for (int r = 0; r < kGroupRows; ++r) acc[r] = fmaf(gs[r * it.N + n], w, acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kGroupRows; ++r) red[ng][r][kl] = acc[r];
    __syncthreads();
    for (int i = threadIdx.x; i < rows * 64; i += 256) {
      const int r = i >> 6, kk = i & 63;
      if (k0 + kk < sl.K)
        p.din[(int64_t)(row0 + r) * p.R + sl.in_off + k0 + kk] = red[0][r][kk] + red[1][r][kk] + red[2][r][kk] + red[3][r][kk];
    }
    __syncthreads();
  }
}

static int check_map(const char* what, int L, int M, int K) {
  if (L < 1 || L > kMaxDepth) return fail(MSG_ERR_UNSUPPORTED, "%s: depth %d (1..%d)", what, L, kMaxDepth);
  if (M < 1) return fail(MSG_ERR_BAD_ARG, "%s: M = %d", what, M);
  if (K < 4 || K % 4 != 0 || K > 64 * kMapCluster)
    return fail(MSG_ERR_UNSUPPORTED, "%s: latent dimension %d (multiple of 4, at most %d)", what, K, 64 * kMapCluster);
  return MSG_OK;
}

}  // namespace msg

using namespace msg;

extern "C" int msg_style_mapping_supported(int depth, int K) {
  return depth >= 1 && depth <= kMaxDepth && K >= 4 && K % 4 == 0 && K <= 64 * kMapCluster;
}

extern "C" int msg_style_mapping_forward(float* acts, float* x0, const float* z, const float* const* weights,
                                         const float* const* biases, int depth, int M, int K, float alpha, float slope,
                                         float gain, float eps, msg_stream_t stream) {
  int rc = check_map("style_mapping_forward", depth, M, K);
  if (rc) return rc;
  if (!acts || !x0 || !z || !weights) return fail(MSG_ERR_BAD_ARG, "style_mapping_forward: null pointer");
  MapParams p{};
  p.z = z; p.L = depth; p.M = M; p.K = K; p.alpha = alpha; p.slope = slope; p.gain = gain; p.eps = eps;
  p.x0 = x0; p.acts = acts;
  for (int l = 0; l < depth; ++l) { p.W[l] = weights[l]; p.bias[l] = biases ? biases[l] : nullptr; }
  const size_t smem = (size_t)2 * kMapRows * K * sizeof(float);
  static bool attr_done[64] = {};
  const int slot = current_device_slot();
  if (!attr_done[slot]) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(style_mapping_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kMapRows * 64 * kMapCluster * 4));
    attr_done[slot] = true;
  }
  const unsigned clusters = (unsigned)((M + kMapRows - 1) / kMapRows);
  style_mapping_forward_kernel<<<clusters * kMapCluster, 256, smem, (cudaStream_t)stream>>>(p);
  MSG_CHECK_LAUNCH("style_mapping_forward");
  return MSG_OK;
}

extern "C" int msg_style_mapping_backward(float* dW, float* db, const float* gy, const float* acts, const float* x0,
                                          const float* const* weights, const float* const* biases, int depth, int M, int K,
                                          float alpha, float slope, float gain, msg_stream_t stream) {
  int rc = check_map("style_mapping_backward", depth, M, K);
  if (rc) return rc;
  if (!dW || !gy || !acts || !x0 || !weights) return fail(MSG_ERR_BAD_ARG, "style_mapping_backward: null pointer");
  MapParams p{};
  p.L = depth; p.M = M; p.K = K; p.alpha = alpha; p.slope = slope; p.gain = gain;
  p.x0 = const_cast<float*>(x0); p.acts = const_cast<float*>(acts); p.gy = gy; p.dW = dW; p.db = db;
  for (int l = 0; l < depth; ++l) { p.W[l] = weights[l]; p.bias[l] = (biases && db) ? biases[l] : nullptr; }
  const size_t smem = (size_t)(4 * kMapRows * K + 4 * kMapRows * 64) * sizeof(float);
  static bool attr_done[64] = {};
  const int slot = current_device_slot();
  if (!attr_done[slot]) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(style_mapping_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (4 * kMapRows * 64 * kMapCluster + 4 * kMapRows * 64) * 4));
    attr_done[slot] = true;
  }
  style_mapping_backward_kernel<<<kMapCluster, 256, smem, (cudaStream_t)stream>>>(p);
  MSG_CHECK_LAUNCH("style_mapping_backward");
  return MSG_OK;
}

static int group_params(LinGroupParams& p, const char* what, const msg_linear_item* items, int n_items,
                        const msg_linear_slot* slots, int n_slots, int M) {
  static_assert(sizeof(msg_linear_item) == sizeof(LinItem) && sizeof(msg_linear_slot) == sizeof(LinSlot), "ABI structs");
  static_assert(sizeof(LinGroupParams) <= 4000, "kernel parameter space");
  if (!items || n_items < 1 || n_items > kMaxItems || n_slots < 0 || n_slots > kMaxItems || M < 1)
    return fail(MSG_ERR_BAD_ARG, "%s: bad item table (1..%d items)", what, kMaxItems);
  memcpy(p.items, items, sizeof(LinItem) * n_items);
  if (slots && n_slots) memcpy(p.slots, slots, sizeof(LinSlot) * n_slots);
  p.n_items = n_items; p.n_slots = n_slots; p.M = M;
  for (int i = 0; i < n_items; ++i)
    if (!items[i].W || items[i].N < 1 || items[i].K < 4 || items[i].K % 4 != 0)
      return fail(MSG_ERR_BAD_ARG, "%s: item %d: K must be a positive multiple of 4", what, i);
  return MSG_OK;
}

// items / slots are HOST arrays (they travel in the kernel parameters); max_n / max_k bound the grid and shared memory
extern "C" int msg_linear_group_forward(float* out, const float* in, int64_t in_row,
                                        const msg_linear_item* items, int n_items, int M, int max_n, int max_k,
                                        msg_stream_t stream) {
  LinGroupParams p{};
  int rc = group_params(p, "linear_group_forward", items, n_items, nullptr, 0, M);
  if (rc) return rc;
  if (!out || !in || max_n < 1 || max_k < 4 || max_k % 4 != 0 || max_k > 2048)
    return fail(MSG_ERR_BAD_ARG, "linear_group_forward: bad argument (K must be a multiple of 4, at most 2048)");
  p.in = in; p.R = in_row; p.out = out;
  const size_t smem = (size_t)kGroupRows * max_k * sizeof(float);
  static bool attr_done[64] = {};
  const int slot = current_device_slot();
  if (!attr_done[slot]) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(linear_group_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGroupRows * 2048 * 4));
    attr_done[slot] = true;
  }
  linear_group_forward_kernel<<<dim3((unsigned)((max_n + 7) / 8), (unsigned)n_items), 256, smem, (cudaStream_t)stream>>>(p);
  MSG_CHECK_LAUNCH("linear_group_forward");
  return MSG_OK;
}

extern "C" int msg_linear_group_backward(float* dW, float* db, float* din, const float* gout,
                                         const float* in, int64_t in_row, const msg_linear_item* items, int n_items,
                                         const msg_linear_slot* slots, int n_slots, int M, int max_n, int max_k,
                                         msg_stream_t stream) {
  LinGroupParams p{};
  int rc = group_params(p, "linear_group_backward", items, n_items, slots, n_slots, M);
  if (rc) return rc;
  if (!gout || !in || max_n < 1 || max_n > 2048 || max_k < 4 || max_k % 4 != 0 || max_k > 2048)
    return fail(MSG_ERR_BAD_ARG, "linear_group_backward: bad argument");
  p.in = in; p.R = in_row; p.gout = gout; p.dW = dW; p.db = db; p.din = din;
  cudaStream_t st = (cudaStream_t)stream;
  if (dW) {
    linear_group_wgrad_kernel<<<dim3((unsigned)((max_n + 7) / 8), (unsigned)n_items), 256, 0, st>>>(p);
    MSG_CHECK_LAUNCH("linear_group_wgrad");
  }
  if (din) {
    if (!slots || n_slots < 1) return fail(MSG_ERR_BAD_ARG, "linear_group_backward: slot table");
    const size_t smem = (size_t)kGroupRows * max_n * sizeof(float);
    static bool attr_done[64] = {};
    const int slot = current_device_slot();
    if (!attr_done[slot]) {
      MSG_CHECK_CUDA(cudaFuncSetAttribute(linear_group_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGroupRows * 2048 * 4));
      attr_done[slot] = true;
    }
    linear_group_dgrad_kernel<<<dim3((unsigned)((max_k + 63) / 64), (unsigned)n_slots), 256, smem, st>>>(p);
    MSG_CHECK_LAUNCH("linear_group_dgrad");
  }
  return MSG_OK;
}
