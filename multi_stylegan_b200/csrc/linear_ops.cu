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

constexpr int kMapRows = 16;      // rows per cluster pass (backward chain)
constexpr int kFwdRows = 8;       // rows per cluster (forward: independent row groups run on separate clusters)
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
  float* gp_all;                  // [L, M, K] gradient w.r.t. every layer's pre-activation (workspace)
};

__device__ __forceinline__ void st_cluster_f32x4(uint32_t addr, const float4& v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// Arithmetic layout (fp32 FMA): a thread owns ONE output column and a quarter of the reduction range and keeps the row
// accumulators in registers; its four weights of a reduction step are one LDS.128 from the CTA's weight slice (rows padded
// by 4 floats: the 32 lanes' 16-byte reads tile the banks exactly), the activations of the step are warp-wide broadcast
// LDS.128.  The next layer's slice is prefetched into registers while the current layer is computed.  (First version:
// a warp per column with the reduction split over lanes re-read the activations for every column and walked its columns
// with dependent loads — 360 us for the 8 layers; the per-layer cuBLAS + activation launches it replaces take 90 us.)
constexpr int kPre = 32;           // float4 per thread that hold a prefetched slice (64 columns x 512 / 256 threads / 4)

__device__ __forceinline__ void prefetch_slice(float4 (&pre)[kPre], const float* __restrict__ src, int count4) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
  for (int i = 0; i < kPre; ++i) {
    const int f = threadIdx.x + 256 * i;
    if (f < count4) pre[i] = __ldg(s4 + f);
  }
}
// slice rows [n][K] -> shared memory rows of K + 4 floats
__device__ __forceinline__ void store_slice(float* wsm, const float4 (&pre)[kPre], int count4, int K) {
#pragma unroll
  for (int i = 0; i < kPre; ++i) {
    const int f = threadIdx.x + 256 * i;
    if (f < count4) {
      const int e = 4 * f, n = e / K, k = e - n * K;
      *reinterpret_cast<float4*>(wsm + n * (K + 4) + k) = pre[i];
    }
  }
}

__global__ void __cluster_dims__(kMapCluster, 1, 1) __launch_bounds__(256, 1)
style_mapping_forward_kernel(const MapParams p) {
  extern __shared__ float4 smem_f4[];
  const int K = p.K;
  float* buf = reinterpret_cast<float*>(smem_f4);                 // [2][kFwdRows][K]   layer input / output (all columns)
  float* wsm = buf + 2 * kFwdRows * K;                            // [64][K + 4]        this CTA's rows of the layer's weights
  float* red = wsm + 64 * (K + 4);                                // [4][kFwdRows][64]  partial sums of the K quarters
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();
  const int row0 = (int)(blockIdx.x / kMapCluster) * kFwdRows;
  const int rows = min(kFwdRows, p.M - row0);
  const int nc = (K + kMapCluster - 1) / kMapCluster;             // <= 64 (host check)
  const int n_begin = min(K, (int)rank * nc), n_end = min(K, n_begin + nc);
  const int ncols = n_end - n_begin;
  const int count4 = ncols * K / 4;

  float4 pre[kPre];
  prefetch_slice(pre, p.W[0] + (int64_t)n_begin * K, count4);
  // pixel norm of the input rows (every CTA builds its own full copy); unused rows are zero
  for (int r = warp; r < kFwdRows; r += 8) {
    if (r < rows) {
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
    } else {
      for (int k = lane; k < K; k += 32) { buf[r * K + k] = 0.f; buf[(kFwdRows + r) * K + k] = 0.f; }
    }
  }
  store_slice(wsm, pre, count4, K);
  __syncthreads();
  cluster_sync_all();          // nobody writes into a peer's buffer before that peer has finished its prologue

  const uint32_t buf_s = smem_u32(buf);
  const int nl = threadIdx.x & 63, kq = threadIdx.x >> 6;
  const int per = ((K + 15) / 16) * 4;                            // reduction indices per quarter (multiple of 4)
  const int k0 = min(K, kq * per), k1 = min(K, k0 + per);
  for (int l = 0; l < p.L; ++l) {
    const float* cur = buf + (l & 1) * kFwdRows * K;
    const uint32_t nxt_s = buf_s + (uint32_t)(((l + 1) & 1) * kFwdRows * K) * 4u;
    if (l + 1 < p.L) prefetch_slice(pre, p.W[l + 1] + (int64_t)n_begin * K, count4);
    float acc[kFwdRows];
#pragma unroll
    for (int r = 0; r < kFwdRows; ++r) acc[r] = 0.f;
    if (nl < ncols) {
      const float* wr = wsm + nl * (K + 4);
#pragma unroll 2
      for (int k = k0; k < k1; k += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(wr + k);
#pragma unroll
        for (int r = 0; r < kFwdRows; ++r) {
          const float4 x4 = *reinterpret_cast<const float4*>(cur + r * K + k);
          acc[r] = fmaf(w4.x, x4.x, fmaf(w4.y, x4.y, fmaf(w4.z, x4.z, fmaf(w4.w, x4.w, acc[r]))));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kFwdRows; ++r) red[(kq * kFwdRows + r) * 64 + nl] = acc[r];
    __syncthreads();
    const float* bias = p.bias[l];
    if ((ncols & 3) == 0 && (n_begin & 3) == 0) {
      // 4 columns per thread: one 16-byte remote store per peer instead of four 4-byte ones
      for (int i = threadIdx.x; i < rows * 16; i += 256) {
        const int r = i >> 4, c = (i & 15) * 4;
        if (c < ncols) {
          const int n = n_begin + c;
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float t = p.alpha * ((red[(0 * kFwdRows + r) * 64 + c + e] + red[(1 * kFwdRows + r) * 64 + c + e]) +
                                       (red[(2 * kFwdRows + r) * 64 + c + e] + red[(3 * kFwdRows + r) * 64 + c + e])) +
                            (bias ? __ldg(bias + n + e) : 0.f);
            v[e] = (t > 0.f ? t : t * p.slope) * p.gain;
          }
          const float4 v4 = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(p.acts + ((int64_t)l * p.M + row0 + r) * K + n) = v4;
          const uint32_t a = nxt_s + (uint32_t)(r * K + n) * 4u;
#pragma unroll
          for (uint32_t cta = 0; cta < kMapCluster; ++cta) st_cluster_f32x4(mapa_cluster(a, cta), v4);
        }
      }
    } else {
      for (int i = threadIdx.x; i < rows * 64; i += 256) {
        const int r = i >> 6, c = i & 63;
        if (c < ncols) {
          const int n = n_begin + c;
          float v = p.alpha * ((red[(0 * kFwdRows + r) * 64 + c] + red[(1 * kFwdRows + r) * 64 + c]) +
                               (red[(2 * kFwdRows + r) * 64 + c] + red[(3 * kFwdRows + r) * 64 + c])) + (bias ? __ldg(bias + n) : 0.f);
          v = (v > 0.f ? v : v * p.slope) * p.gain;
          p.acts[((int64_t)l * p.M + row0 + r) * K + n] = v;
          const uint32_t a = nxt_s + (uint32_t)(r * K + n) * 4u;
#pragma unroll
          for (uint32_t cta = 0; cta < kMapCluster; ++cta) st_cluster_f32(mapa_cluster(a, cta), v);
        }
      }
    }
    if (l + 1 < p.L) store_slice(wsm, pre, count4, K);   // (every thread is past its reads of the slice: the barrier above)
    __syncthreads();
    cluster_sync_all();
  }
}

// Backward, part 1 (the sequential chain): gp_l = g_l * lrelu'(act_l) * gain for every layer, written to gp_all, and
// g_{l-1} = alpha * gp_l W_l propagated through distributed shared memory.  The weight / bias gradients do not sit on the
// chain: part 2 computes them for all layers at once on the whole chip.
__global__ void __cluster_dims__(kMapCluster, 1, 1) __launch_bounds__(256, 1)
style_mapping_backward_kernel(const MapParams p) {
  extern __shared__ float4 smem_f4[];
  const int K = p.K;
  const int nc = (K + kMapCluster - 1) / kMapCluster;       // <= 64 (host check)
  float* gbuf = reinterpret_cast<float*>(smem_f4);          // [2][kMapRows][K]  gradient w.r.t. a layer output, then (in
                                                            //                   place) w.r.t. its pre-activation
  const int ncp = (nc + 3) & ~3;                            // staged row length: 16-byte rows for the asynchronous copies
  float* wsm = gbuf + 2 * kMapRows * K;                     // [K][ncp]          W_l[n][n_begin + c]: the owned INPUT columns
  float* red = wsm + K * ncp;                               // [4][kMapRows][64]
  const uint32_t rank = cluster_ctarank();
  const int n_begin = min(K, (int)rank * nc), n_end = min(K, n_begin + nc);
  const int ncols = n_end - n_begin;
  const bool vec_cols = (ncols & 3) == 0 && (n_begin & 3) == 0;
  const uint32_t gbuf_s = smem_u32(gbuf), wsm_s = smem_u32(wsm);
  const int nl = threadIdx.x & 63, nq = threadIdx.x >> 6;
  const int per = ((K + 15) / 16) * 4;
  const int q0 = min(K, nq * per), q1 = min(K, q0 + per);

  for (int row0 = 0; row0 < p.M; row0 += kMapRows) {
    const int rows = min(kMapRows, p.M - row0);
    for (int i = threadIdx.x; i < 2 * kMapRows * K; i += 256) gbuf[i] = i < rows * K ? __ldg(p.gy + (int64_t)row0 * K + i) : 0.f;
    __syncthreads();
    cluster_sync_all();
    for (int l = p.L - 1, step = 0; l >= 0; --l, ++step) {
      float* gp = gbuf + (step & 1) * kMapRows * K;
      const uint32_t nxt_s = gbuf_s + (uint32_t)(((step + 1) & 1) * kMapRows * K) * 4u;
      const float* act = p.acts + ((int64_t)l * p.M + row0) * K;
      // the owned input columns of W_l: asynchronous 16-byte copies that land while the activation pass below runs
      if (l > 0) {
        const float* W = p.W[l];
        if (vec_cols) {
          const int q = ncols >> 2;
          for (int i = threadIdx.x; i < K * q; i += 256) {
            const int n = i / q, c = i - n * q;
            cp_async16(wsm_s + (uint32_t)(n * ncp + 4 * c) * 4u, W + (int64_t)n * K + n_begin + 4 * c);
          }
        } else {
          for (int i = threadIdx.x; i < K * ncols; i += 256) {
            const int n = i / ncols, c = i - n * ncols;
            wsm[n * ncp + c] = __ldg(W + (int64_t)n * K + n_begin + c);
          }
        }
      }
      float* gp_out = p.gp_all + ((int64_t)l * p.M + row0) * K;
      for (int i = threadIdx.x; i < rows * K; i += 256) {
        const float v = gp[i] * (__ldg(act + i) > 0.f ? 1.f : p.slope) * p.gain;
        gp[i] = v;
        const int c = i % K;
        if (c >= n_begin && c < n_end) gp_out[i] = v;        // every CTA holds all of gp; each writes its own columns
      }
      if (l > 0) {
        cp_async_wait_all();
        __syncthreads();
        // gradient w.r.t. the layer input, owned input columns: g_prev[r][k] = alpha * sum_n gp[r][n] * W[n][k]
        float acc[kMapRows];
#pragma unroll
        for (int r = 0; r < kMapRows; ++r) acc[r] = 0.f;
        if (nl < ncols) {
#pragma unroll 2
          for (int n = q0; n < q1; n += 4) {
            const float w0 = wsm[(n + 0) * ncp + nl], w1 = wsm[(n + 1) * ncp + nl];
            const float w2 = wsm[(n + 2) * ncp + nl], w3 = wsm[(n + 3) * ncp + nl];
#pragma unroll
            for (int r = 0; r < kMapRows; ++r) {
              const float4 g4 = *reinterpret_cast<const float4*>(gp + r * K + n);
              acc[r] = fmaf(w0, g4.x, fmaf(w1, g4.y, fmaf(w2, g4.z, fmaf(w3, g4.w, acc[r]))));
            }
          }
        }
#pragma unroll
        for (int r = 0; r < kMapRows; ++r) red[(nq * kMapRows + r) * 64 + nl] = acc[r];
        __syncthreads();
        if (vec_cols) {
          for (int i = threadIdx.x; i < rows * 16; i += 256) {
            const int r = i >> 4, c = (i & 15) * 4;
            if (c < ncols) {
              float v[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                v[e] = p.alpha * ((red[(0 * kMapRows + r) * 64 + c + e] + red[(1 * kMapRows + r) * 64 + c + e]) +
                                  (red[(2 * kMapRows + r) * 64 + c + e] + red[(3 * kMapRows + r) * 64 + c + e]));
              const uint32_t a = nxt_s + (uint32_t)(r * K + n_begin + c) * 4u;
#pragma unroll
              for (uint32_t cta = 0; cta < kMapCluster; ++cta) st_cluster_f32x4(mapa_cluster(a, cta), make_float4(v[0], v[1], v[2], v[3]));
            }
          }
        } else {
          for (int i = threadIdx.x; i < rows * 64; i += 256) {
            const int r = i >> 6, c = i & 63;
            if (c < ncols) {
              const float v = p.alpha * ((red[(0 * kMapRows + r) * 64 + c] + red[(1 * kMapRows + r) * 64 + c]) +
                                         (red[(2 * kMapRows + r) * 64 + c] + red[(3 * kMapRows + r) * 64 + c]));
              const uint32_t a = nxt_s + (uint32_t)(r * K + n_begin + c) * 4u;
#pragma unroll
              for (uint32_t cta = 0; cta < kMapCluster; ++cta) st_cluster_f32(mapa_cluster(a, cta), v);
            }
          }
        }
      }
      __syncthreads();
      cluster_sync_all();
    }
  }
}

// Backward, part 2: dW_l[n, :] = alpha * sum_m gp_l[m, n] * x_{l-1}[m, :],  db_l[n] = sum_m gp_l[m, n]; warp = one row n of
// one layer (grid.y), lanes run over the input index.
__global__ void __launch_bounds__(256)
style_mapping_wgrad_kernel(const MapParams p) {
  const int K = p.K, l = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 8 + warp;
  if (n >= K) return;
  const float* gp = p.gp_all + (int64_t)l * p.M * K;
  const float* xin = l > 0 ? p.acts + (int64_t)(l - 1) * p.M * K : p.x0;
  float bsum = 0.f;
  for (int k0 = 0; k0 < K; k0 += 512) {
    float4 a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int m = 0; m < p.M; ++m) {
      const float g = __ldg(gp + (int64_t)m * K + n);
      if (k0 == 0) bsum += g;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + lane * 4 + j * 128;
        if (k < K) {
          const float4 x4 = __ldg(reinterpret_cast<const float4*>(xin + (int64_t)m * K + k));
          a[j].x = fmaf(g, x4.x, a[j].x); a[j].y = fmaf(g, x4.y, a[j].y); a[j].z = fmaf(g, x4.z, a[j].z); a[j].w = fmaf(g, x4.w, a[j].w);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lane * 4 + j * 128;
      if (k < K)
        *reinterpret_cast<float4*>(p.dW + ((int64_t)l * K + n) * K + k) =
            make_float4(a[j].x * p.alpha, a[j].y * p.alpha, a[j].z * p.alpha, a[j].w * p.alpha);
    }
  }
  if (lane == 0 && p.db != nullptr && p.bias[l] != nullptr) p.db[(int64_t)l * K + n] = bsum;
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
  const size_t smem = ((size_t)2 * kFwdRows * K + (size_t)64 * (K + 4) + 4 * kFwdRows * 64) * sizeof(float);
  static bool attr_done[64] = {};
  const int slot = current_device_slot();
  if (!attr_done[slot]) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(style_mapping_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (2 * kFwdRows * 512 + 64 * 516 + 4 * kFwdRows * 64) * 4));
    attr_done[slot] = true;
  }
  const unsigned clusters = (unsigned)((M + kFwdRows - 1) / kFwdRows);
  style_mapping_forward_kernel<<<clusters * kMapCluster, 256, smem, (cudaStream_t)stream>>>(p);
  MSG_CHECK_LAUNCH("style_mapping_forward");
  return MSG_OK;
}

extern "C" int msg_style_mapping_backward(float* dW, float* db, const float* gy, const float* acts, const float* x0,
                                          const float* const* weights, const float* const* biases, int depth, int M, int K,
                                          float alpha, float slope, float gain, float* workspace, msg_stream_t stream) {
  int rc = check_map("style_mapping_backward", depth, M, K);
  if (rc) return rc;
  if (!dW || !gy || !acts || !x0 || !weights || !workspace) return fail(MSG_ERR_BAD_ARG, "style_mapping_backward: null pointer");
  MapParams p{};
  p.L = depth; p.M = M; p.K = K; p.alpha = alpha; p.slope = slope; p.gain = gain;
  p.x0 = const_cast<float*>(x0); p.acts = const_cast<float*>(acts); p.gy = gy; p.dW = dW; p.db = db; p.gp_all = workspace;
  for (int l = 0; l < depth; ++l) { p.W[l] = weights[l]; p.bias[l] = (biases && db) ? biases[l] : nullptr; }
  const int ncp = (((K + kMapCluster - 1) / kMapCluster) + 3) & ~3;
  const size_t smem = ((size_t)2 * kMapRows * K + (size_t)K * ncp + 4 * kMapRows * 64) * sizeof(float);
  static bool attr_done[64] = {};
  const int slot = current_device_slot();
  if (!attr_done[slot]) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(style_mapping_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (2 * kMapRows * 512 + 512 * 64 + 4 * kMapRows * 64) * 4));
    attr_done[slot] = true;
  }
  cudaStream_t st = (cudaStream_t)stream;
  style_mapping_backward_kernel<<<kMapCluster, 256, smem, st>>>(p);
  MSG_CHECK_LAUNCH("style_mapping_backward(chain)");
  style_mapping_wgrad_kernel<<<dim3((unsigned)((K + 7) / 8), (unsigned)depth), 256, 0, st>>>(p);
  MSG_CHECK_LAUNCH("style_mapping_backward(weight gradients)");
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
