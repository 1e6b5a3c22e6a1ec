// Bandwidth kernels of the discriminator's non-local (self-attention) block — u_net_2d_discriminator.py:332-381:
//     theta = conv1x1(x), phi = maxpool2x2(conv1x1(x)), g = maxpool2x2(conv1x1(x))
//     beta = softmax(theta^T phi) over the pooled positions,  o = conv1x1(g beta^T),  out = (gamma o + res(x)) / sqrt(2)
// The three input convolutions run as ONE tcgen05 GEMM with the filters stacked ([theta | phi | g] output channels); the
// two attention products run on the same conv engine as 1x1 convolutions with one "filter bank" per sample (phi / g of
// that sample).  What is left are four memory-bound passes, written here for channels-last fp32:
//   nl_split_pool     stacked conv output -> dense theta + 2x2 max-pooled phi, g + the argmax of every window
//   nl_merge_unpool   its adjoint: (d theta, d phi_pooled, d g_pooled) -> gradient of the stacked conv output
//   softmax_rows      in-place row softmax of the [B*HW, HW/4] score matrix (one warp per row, the row in registers)
//   softmax_rows_bwd  dS = P * (dP - sum_j P_j dP_j), in place on dP
// F.max_pool2d semantics: floor mode (an odd last row / column is dropped), the first maximum in window scan order wins.
#include "common.cuh"

namespace msg {

struct NlPoolParams {
  const float4* qkv;       // [B, H, W, CT/4]
  float4* theta;           // [B, H, W, cq/4]
  float4* phi;             // [B, PH, PW, cq/4]
  float4* g;               // [B, PH, PW, cv/4]
  uint32_t* idx;           // [B, PH, PW, (cq+cv)/4]: four 8-bit window positions (0..3) per float4
  int B, H, W, PH, PW, cq4, cv4;
};

__device__ __forceinline__ void max_sel(float& m, uint32_t& sel, float v, uint32_t pos) {
  if (v > m) { m = v; sel = pos; }
}

__global__ void __launch_bounds__(256)
nl_split_pool_kernel(const NlPoolParams p) {
  const int CT4 = 2 * p.cq4 + p.cv4;
  const int WH = (p.H + 1) / 2, WW = (p.W + 1) / 2;       // windows including the partial ones of an odd last row / column
  const int64_t total = (int64_t)p.B * WH * WW * CT4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int c4 = (int)(r % CT4); r /= CT4;
    const int wx = (int)(r % WW); r /= WW;
    const int wy = (int)(r % WH);
    const int b = (int)(r / WH);
    const int y0 = 2 * wy, x0 = 2 * wx;
    if (c4 < p.cq4) {
      // theta: plain copy of the window's pixels into the dense tensor
      for (int dy = 0; dy < 2; ++dy)
        for (int dx = 0; dx < 2; ++dx) {
          const int y = y0 + dy, x = x0 + dx;
          if (y < p.H && x < p.W) {
            const int64_t pix = ((int64_t)b * p.H + y) * p.W + x;
            p.theta[pix * p.cq4 + c4] = __ldg(p.qkv + pix * CT4 + c4);
          }
        }
      continue;
    }
    if (wy >= p.PH || wx >= p.PW) continue;                // partial window: not pooled (floor mode)
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      const int64_t pix = ((int64_t)b * p.H + y0 + (pos >> 1)) * p.W + x0 + (pos & 1);
      const float4 v = __ldg(p.qkv + pix * CT4 + c4);
      max_sel(m.x, s0, v.x, pos); max_sel(m.y, s1, v.y, pos); max_sel(m.z, s2, v.z, pos); max_sel(m.w, s3, v.w, pos);
    }
    const int64_t pp = ((int64_t)b * p.PH + wy) * p.PW + wx;
    const int cp4 = c4 - p.cq4;                            // column among the pooled channels [phi | g]
    if (cp4 < p.cq4) p.phi[pp * p.cq4 + cp4] = m;
    else p.g[pp * p.cv4 + (cp4 - p.cq4)] = m;
    p.idx[pp * (p.cq4 + p.cv4) + cp4] = s0 | (s1 << 8) | (s2 << 16) | (s3 << 24);
  }
}

struct NlUnpoolParams {
  float4* dqkv;            // [B, H, W, CT/4]
  const float4* dtheta;    // [B, H, W, cq/4]
  const float4* dphi;      // [B, PH, PW, cq/4]
  const float4* dg;        // [B, PH, PW, cv/4]
  const uint32_t* idx;
  int B, H, W, PH, PW, cq4, cv4;
};

__global__ void __launch_bounds__(256)
nl_merge_unpool_kernel(const NlUnpoolParams p) {
  const int CT4 = 2 * p.cq4 + p.cv4;
  const int WH = (p.H + 1) / 2, WW = (p.W + 1) / 2;
  const int64_t total = (int64_t)p.B * WH * WW * CT4;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int c4 = (int)(r % CT4); r /= CT4;
    const int wx = (int)(r % WW); r /= WW;
    const int wy = (int)(r % WH);
    const int b = (int)(r / WH);
    const int y0 = 2 * wy, x0 = 2 * wx;
    const bool pooled = c4 >= p.cq4 && wy < p.PH && wx < p.PW;
    float4 gv = zero;
    uint32_t sel = 0;
    if (pooled) {
      const int64_t pp = ((int64_t)b * p.PH + wy) * p.PW + wx;
      const int cp4 = c4 - p.cq4;
      gv = cp4 < p.cq4 ? __ldg(p.dphi + pp * p.cq4 + cp4) : __ldg(p.dg + pp * p.cv4 + (cp4 - p.cq4));
      sel = __ldg(p.idx + pp * (p.cq4 + p.cv4) + cp4);
    }
#pragma unroll
    for (int pos = 0; pos < 4; ++pos) {
      const int y = y0 + (pos >> 1), x = x0 + (pos & 1);
      if (y >= p.H || x >= p.W) continue;
      const int64_t pix = ((int64_t)b * p.H + y) * p.W + x;
      float4 o;
      if (c4 < p.cq4) {
        o = __ldg(p.dtheta + pix * p.cq4 + c4);
      } else if (pooled) {
        o.x = (sel & 0xffu) == (uint32_t)pos ? gv.x : 0.f;
        o.y = ((sel >> 8) & 0xffu) == (uint32_t)pos ? gv.y : 0.f;
        o.z = ((sel >> 16) & 0xffu) == (uint32_t)pos ? gv.z : 0.f;
        o.w = ((sel >> 24) & 0xffu) == (uint32_t)pos ? gv.w : 0.f;
      } else {
        o = zero;
      }
      p.dqkv[pix * CT4 + c4] = o;
    }
  }
}

// One warp per row; the row (n <= 128 * NV floats) lives in registers as NV float4 per lane.
template <int NV>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float4* __restrict__ x, int64_t rows, int n4) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4* xr = x + row * n4;
  float4 v[NV];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = lane + 32 * j;
    if (c < n4) {
      v[j] = xr[c];
      m = fmaxf(m, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = lane + 32 * j;
    if (c < n4) {
      v[j].x = expf(v[j].x - m); v[j].y = expf(v[j].y - m); v[j].z = expf(v[j].z - m); v[j].w = expf(v[j].w - m);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  }
  s = warp_sum(s);
  const float inv = 1.f / s;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = lane + 32 * j;
    if (c < n4) xr[c] = make_float4(v[j].x * inv, v[j].y * inv, v[j].z * inv, v[j].w * inv);
  }
}

template <int NV>
__global__ void __launch_bounds__(256)
softmax_rows_bwd_kernel(float4* __restrict__ dp, const float4* __restrict__ p, int64_t rows, int n4) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float4* dr = dp + row * n4;
  const float4* pr = p + row * n4;
  float4 pv[NV], gv[NV];
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = lane + 32 * j;
    if (c < n4) {
      pv[j] = __ldg(pr + c);
      gv[j] = dr[c];
      dot += (pv[j].x * gv[j].x + pv[j].y * gv[j].y) + (pv[j].z * gv[j].z + pv[j].w * gv[j].w);
    }
  }
  dot = warp_sum(dot);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = lane + 32 * j;
    if (c < n4)
      dr[c] = make_float4(pv[j].x * (gv[j].x - dot), pv[j].y * (gv[j].y - dot), pv[j].z * (gv[j].z - dot), pv[j].w * (gv[j].w - dot));
  }
}

static inline bool al16p(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static unsigned grid_items(int64_t total) {
  const int64_t want = ceil_div(total, 256), cap = (int64_t)num_sms() * 16;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace msg

using namespace msg;

extern "C" int msg_nl_split_pool(float* theta, float* phi_p, float* g_p, uint32_t* idx, const float* qkv, int B, int H, int W,
                                 int cq, int cv, msg_stream_t stream) {
  if (B < 0 || H < 1 || W < 1 || cq < 4 || cv < 4 || cq % 4 || cv % 4)
    return fail(MSG_ERR_BAD_ARG, "nl_split_pool: channel counts must be positive multiples of 4");
  if (B == 0) return MSG_OK;
  if (!theta || !phi_p || !g_p || !idx || !qkv || !al16p(theta) || !al16p(phi_p) || !al16p(g_p) || !al16p(qkv))
    return fail(MSG_ERR_BAD_ARG, "nl_split_pool: null or unaligned pointer");
  NlPoolParams p{reinterpret_cast<const float4*>(qkv), reinterpret_cast<float4*>(theta), reinterpret_cast<float4*>(phi_p),
                 reinterpret_cast<float4*>(g_p), idx, B, H, W, H / 2, W / 2, cq / 4, cv / 4};
  const int64_t total = (int64_t)B * ((H + 1) / 2) * ((W + 1) / 2) * (2 * p.cq4 + p.cv4);
  nl_split_pool_kernel<<<grid_items(total), 256, 0, (cudaStream_t)stream>>>(p);
  MSG_CHECK_LAUNCH("nl_split_pool");
  return MSG_OK;
}

extern "C" int msg_nl_merge_unpool(float* dqkv, const float* dtheta, const float* dphi_p, const float* dg_p,
                                   const uint32_t* idx, int B, int H, int W, int cq, int cv, msg_stream_t stream) {
  if (B < 0 || H < 1 || W < 1 || cq < 4 || cv < 4 || cq % 4 || cv % 4)
    return fail(MSG_ERR_BAD_ARG, "nl_merge_unpool: channel counts must be positive multiples of 4");
  if (B == 0) return MSG_OK;
  if (!dqkv || !dtheta || !dphi_p || !dg_p || !idx || !al16p(dqkv) || !al16p(dtheta) || !al16p(dphi_p) || !al16p(dg_p))
    return fail(MSG_ERR_BAD_ARG, "nl_merge_unpool: null or unaligned pointer");
  NlUnpoolParams p{reinterpret_cast<float4*>(dqkv), reinterpret_cast<const float4*>(dtheta),
                   reinterpret_cast<const float4*>(dphi_p), reinterpret_cast<const float4*>(dg_p), idx, B, H, W, H / 2, W / 2,
                   cq / 4, cv / 4};
  const int64_t total = (int64_t)B * ((H + 1) / 2) * ((W + 1) / 2) * (2 * p.cq4 + p.cv4);
  nl_merge_unpool_kernel<<<grid_items(total), 256, 0, (cudaStream_t)stream>>>(p);
  MSG_CHECK_LAUNCH("nl_merge_unpool");
  return MSG_OK;
}

template <int NV>
static int launch_softmax(float* x, const float* p, int64_t rows, int n, bool bwd, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div(rows, 8);
  if (bwd) softmax_rows_bwd_kernel<NV><<<grid, 256, 0, st>>>(reinterpret_cast<float4*>(x), reinterpret_cast<const float4*>(p), rows, n / 4);
  else softmax_rows_kernel<NV><<<grid, 256, 0, st>>>(reinterpret_cast<float4*>(x), rows, n / 4);
  MSG_CHECK_LAUNCH(bwd ? "softmax_rows_bwd" : "softmax_rows");
  return MSG_OK;
}

static int softmax_dispatch(float* x, const float* p, int64_t rows, int n, bool bwd, msg_stream_t stream) {
  const char* what = bwd ? "softmax_rows_bwd" : "softmax_rows";
  if (rows < 0 || n < 4 || n % 4 || n > 4096) return fail(MSG_ERR_UNSUPPORTED, "%s: row length %d (multiple of 4, <= 4096)", what, n);
  if (rows == 0) return MSG_OK;
  if (rows > 0x7fffffffLL * 8) return fail(MSG_ERR_UNSUPPORTED, "%s: too many rows", what);
  if (!x || !al16p(x) || (bwd && (!p || !al16p(p)))) return fail(MSG_ERR_BAD_ARG, "%s: null or unaligned pointer", what);
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 512) return launch_softmax<4>(x, p, rows, n, bwd, st);
  if (n <= 1024) return launch_softmax<8>(x, p, rows, n, bwd, st);
  if (n <= 2048) return launch_softmax<16>(x, p, rows, n, bwd, st);
  return launch_softmax<32>(x, p, rows, n, bwd, st);
}

extern "C" int msg_softmax_rows(float* x, int64_t rows, int n, msg_stream_t stream) {
  return softmax_dispatch(x, nullptr, rows, n, false, stream);
}

extern "C" int msg_softmax_rows_bwd(float* dp_inout, const float* p, int64_t rows, int n, msg_stream_t stream) {
  return softmax_dispatch(dp_inout, p, rows, n, true, stream);
}
