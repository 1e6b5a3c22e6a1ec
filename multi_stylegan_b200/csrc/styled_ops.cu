// Backward of the StyledConv2d tail in the shared-weight form of the modulated convolution.
//
// Forward (multi_stylegan_generator.py:379-411 with the style moved onto the activations, :289-292 noise,
// op_static/fused_act.py:58 bias + leaky ReLU + gain), per sample b, channel c, pixel p:
//     v    = d[b,c] * conv(scale * W, xs)[b,c,p] + nw * noise[b,p] + bias[c]
//     out  = lrelu(v) * gain                 out2 = out * s2[b,c]      (the next layer's modulated input)
// This pass turns the two incoming gradients (w.r.t. out and out2) into everything the layer's backward needs that is
// not a GEMM, in ONE sweep over the activation (channels-last, [B, rows = H*W, C]):
//     gt    = g_out + s2[b,c] * g_out2
//     gv    = gt * (out > 0 ? 1 : slope) * gain                      (gradient w.r.t. v)
//     g_pre = gv * d[b,c]                                             (gradient w.r.t. the shared-weight conv output)
//     S1[b,c] = sum_p gv              -> dbias[c]  = sum_b S1
//     S2[b,c] = sum_p gv * v          -> dd[b,c]   = (S2 - nw * S3 - bias[c] * S1) / d[b,c]
//     S3[b,c] = sum_p gv * noise[b,p] -> dnw       = sum_{b,c} S3
//     S4[b,c] = sum_p out * g_out2    -> ds2[b,c]
// v is recovered from out (v = out / gain for out > 0, out / (gain * slope) otherwise).  The reference reaches the
// same numbers through autograd over :384-388 + fused_act.py:31-40.  Deterministic two-stage reductions, no atomics.
#include "common.cuh"

namespace msg {

struct StyledBwdParams {
  float4* g_pre;
  const float4* g_out;
  const float4* g_out2;
  const float4* out;
  const float* d;
  int64_t d_bs;
  const float* s2;
  int64_t s2_bs;
  const float* noise;
  int64_t noise_bs;
  float slope, gain;
  int64_t rows;              // pixels per sample
  int C4, B;
  int64_t rows_per_block;
  float* partial;            // [B][gridDim.y][4][C]
};

__device__ __forceinline__ float4 f4mul(const float4& a, const float4& b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ void f4fma(float4& acc, const float4& a, const float4& b) {
  acc.x = fmaf(a.x, b.x, acc.x); acc.y = fmaf(a.y, b.y, acc.y); acc.z = fmaf(a.z, b.z, acc.z); acc.w = fmaf(a.w, b.w, acc.w);
}

template <bool HAS_G1, bool HAS_G2>
__global__ void __launch_bounds__(256)
styled_act_bwd_kernel(const StyledBwdParams p) {
  constexpr int LOOP = 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * 32 + lane;
  const bool qok = q < p.C4;
  const int64_t b = blockIdx.z;
  const int64_t r0 = (int64_t)blockIdx.y * p.rows_per_block;
  int64_t r1 = r0 + p.rows_per_block;
  if (r1 > p.rows) r1 = p.rows;
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f), zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 dv = one, sv = zero;
  if (qok) {
    if (p.d) dv = __ldg(reinterpret_cast<const float4*>(p.d + b * p.d_bs) + q);
    if (HAS_G2) sv = __ldg(reinterpret_cast<const float4*>(p.s2 + b * p.s2_bs) + q);
  }
  const float m_pos = p.gain, m_neg = p.gain * p.slope;
  const float i_pos = 1.f / p.gain, i_neg = 1.f / (p.gain * p.slope);
  const float* nzb = p.noise ? p.noise + b * p.noise_bs : nullptr;
  const int64_t base = b * p.rows * p.C4 + q;
  float4 S1 = zero, S2 = zero, S3 = zero, S4 = zero;
  for (int64_t r = r0 + warp; r < r1; r += 8 * LOOP) {
    float4 g1[LOOP], g2[LOOP], o[LOOP];
    float nz[LOOP];
#pragma unroll
    for (int l = 0; l < LOOP; ++l) {
      const int64_t rr = r + 8 * l;
      nz[l] = 0.f;
      if (qok && rr < r1) {
        const int64_t i = base + rr * p.C4;
        o[l] = __ldg(p.out + i);
        if (HAS_G1) g1[l] = __ldg(p.g_out + i);
        if (HAS_G2) g2[l] = __ldg(p.g_out2 + i);
        if (nzb) nz[l] = __ldg(nzb + rr);
      }
    }
#pragma unroll
    for (int l = 0; l < LOOP; ++l) {
      const int64_t rr = r + 8 * l;
      if (qok && rr < r1) {
        float4 gt = HAS_G1 ? g1[l] : zero;
        if (HAS_G2) { f4fma(gt, sv, g2[l]); f4fma(S4, o[l], g2[l]); }
        const float4 m = make_float4(o[l].x > 0.f ? m_pos : m_neg, o[l].y > 0.f ? m_pos : m_neg,
                                     o[l].z > 0.f ? m_pos : m_neg, o[l].w > 0.f ? m_pos : m_neg);
        const float4 iv = make_float4(o[l].x > 0.f ? i_pos : i_neg, o[l].y > 0.f ? i_pos : i_neg,
                                      o[l].z > 0.f ? i_pos : i_neg, o[l].w > 0.f ? i_pos : i_neg);
        const float4 gv = f4mul(gt, m);
        const float4 v = f4mul(o[l], iv);
        S1.x += gv.x; S1.y += gv.y; S1.z += gv.z; S1.w += gv.w;
        f4fma(S2, gv, v);
        S3.x = fmaf(gv.x, nz[l], S3.x); S3.y = fmaf(gv.y, nz[l], S3.y); S3.z = fmaf(gv.z, nz[l], S3.z); S3.w = fmaf(gv.w, nz[l], S3.w);
        p.g_pre[base + rr * p.C4] = f4mul(gv, dv);
      }
    }
  }
  __shared__ float4 sacc[4][8][32];
  sacc[0][warp][lane] = S1; sacc[1][warp][lane] = S2; sacc[2][warp][lane] = S3; sacc[3][warp][lane] = S4;
  __syncthreads();
  if (warp < 4 && qok) {
    float4 t = sacc[warp][0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) {
      const float4 u = sacc[warp][w][lane];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    const int C = p.C4 * 4;
    float* dst = p.partial + ((b * gridDim.y + blockIdx.y) * 4 + warp) * (int64_t)C;
    reinterpret_cast<float4*>(dst)[q] = t;
  }
}

// sums[k][b][c] = sum_j partial[b][j][k][c]   (fixed order: deterministic)
__global__ void __launch_bounds__(256)
styled_act_bwd_reduce_kernel(float* __restrict__ sums, const float* __restrict__ partial, int B, int gy, int C) {
  const int64_t total = (int64_t)4 * B * C;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int b = (int)((i / C) % B);
  const int k = (int)(i / ((int64_t)C * B));
  float acc = 0.f;
  for (int j = 0; j < gy; ++j) acc += partial[(((int64_t)b * gy + j) * 4 + k) * C + c];
  sums[i] = acc;
}

static inline int styled_row_blocks(int64_t rows, int C, int B) {
  const int gx = (int)ceil_div(C >> 2, 32);
  int64_t gy = ((int64_t)num_sms() * 8) / ((int64_t)gx * (B > 0 ? B : 1));
  const int64_t max_by_rows = ceil_div(rows, 16);
  if (gy > max_by_rows) gy = max_by_rows;
  if (gy < 1) gy = 1;
  if (gy > 65535) gy = 65535;
  return (int)gy;
}

// Wsq[o,c] = sum_t W[o,c,t]^2   (the weight half of the demodulation factor, multi_stylegan_generator.py:386-388)
__global__ void __launch_bounds__(256)
weight_sq_kernel(float* __restrict__ wsq, const float* __restrict__ w, int64_t oc, int taps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= oc) return;
  const float* src = w + i * taps;
  float acc = 0.f;
  for (int t = 0; t < taps; ++t) { const float v = __ldg(src + t); acc = fmaf(v, v, acc); }
  wsq[i] = acc;
}

// d[b,o] = rsqrt(scale^2 * sum_c s[b,c]^2 * Wsq[o,c] + 1e-8): one warp per (b, o)
__global__ void __launch_bounds__(256)
demod_kernel(float* __restrict__ d, const float* __restrict__ s, const float* __restrict__ wsq, int B, int O, int C,
             float scale2) {
  const int warp = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= B * O) return;
  const int b = warp / O, o = warp - b * O;
  const float* sb = s + (int64_t)b * C;
  const float* wr = wsq + (int64_t)o * C;
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) { const float sv = __ldg(sb + c); acc = fmaf(sv * sv, __ldg(wr + c), acc); }
  acc = warp_sum(acc);
  if (lane == 0) d[warp] = rsqrtf(scale2 * acc + 1e-8f);
}

}  // namespace msg

using namespace msg;

extern "C" size_t msg_styled_act_bwd_workspace(int B, int64_t rows, int C) {
  if (B <= 0 || rows <= 0 || C <= 0 || C % 4) return 16;
  return (size_t)B * styled_row_blocks(rows, C, B) * 4 * (size_t)C * sizeof(float) + 32;
}

extern "C" int msg_styled_act_bwd(float* g_pre, float* sums, const float* g_out, const float* g_out2, const float* out,
                                  const float* col_scale, int64_t col_scale_batch_stride, const float* out2_scale,
                                  int64_t out2_scale_batch_stride, const float* noise, int64_t noise_batch_stride,
                                  int B, int64_t rows, int C, float slope, float gain, void* workspace,
                                  size_t workspace_bytes, msg_stream_t stream) {
  if (B < 0 || rows < 0 || C <= 0) return fail(MSG_ERR_BAD_ARG, "styled_act_bwd: bad sizes");
  if (C % 4) return fail(MSG_ERR_UNSUPPORTED, "styled_act_bwd: C must be a multiple of 4");
  if (!sums) return fail(MSG_ERR_BAD_ARG, "styled_act_bwd: null sums");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0 || rows == 0) {
    MSG_CHECK_CUDA(cudaMemsetAsync(sums, 0, (size_t)4 * (B > 0 ? B : 0) * C * sizeof(float), st));
    return MSG_OK;
  }
  if (!g_pre || !out || (!g_out && !g_out2)) return fail(MSG_ERR_BAD_ARG, "styled_act_bwd: null pointer");
  if (g_out2 && !out2_scale) return fail(MSG_ERR_BAD_ARG, "styled_act_bwd: g_out2 needs out2_scale");
  if (gain == 0.f || slope == 0.f) return fail(MSG_ERR_UNSUPPORTED, "styled_act_bwd: gain and slope must be non-zero");
  const uintptr_t al = reinterpret_cast<uintptr_t>(g_pre) | reinterpret_cast<uintptr_t>(g_out) |
                       reinterpret_cast<uintptr_t>(g_out2) | reinterpret_cast<uintptr_t>(out) |
                       reinterpret_cast<uintptr_t>(col_scale) | reinterpret_cast<uintptr_t>(out2_scale);
  if ((al & 15u) || (col_scale_batch_stride % 4) || (out2_scale_batch_stride % 4))
    return fail(MSG_ERR_BAD_ARG, "styled_act_bwd: pointers must be 16-byte aligned and batch strides multiples of 4");
  const size_t need = msg_styled_act_bwd_workspace(B, rows, C);
  if (!workspace || workspace_bytes < need)
    return fail(MSG_ERR_WORKSPACE, "styled_act_bwd: workspace %zu < %zu", workspace_bytes, need);
  if (B > 65535) return fail(MSG_ERR_UNSUPPORTED, "styled_act_bwd: batch > 65535");
  StyledBwdParams p{};
  p.g_pre = reinterpret_cast<float4*>(g_pre);
  p.g_out = reinterpret_cast<const float4*>(g_out);
  p.g_out2 = reinterpret_cast<const float4*>(g_out2);
  p.out = reinterpret_cast<const float4*>(out);
  p.d = col_scale; p.d_bs = col_scale_batch_stride;
  p.s2 = out2_scale; p.s2_bs = out2_scale_batch_stride;
  p.noise = noise; p.noise_bs = noise_batch_stride;
  p.slope = slope; p.gain = gain;
  p.rows = rows; p.C4 = C / 4; p.B = B;
  const int gy = styled_row_blocks(rows, C, B);
  p.rows_per_block = ceil_div(rows, gy);
  p.partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 15) & ~(uintptr_t)15);
  dim3 grid((unsigned)ceil_div(p.C4, 32), (unsigned)gy, (unsigned)B);
  if (g_out && g_out2) styled_act_bwd_kernel<true, true><<<grid, 256, 0, st>>>(p);
  else if (g_out) styled_act_bwd_kernel<true, false><<<grid, 256, 0, st>>>(p);
  else styled_act_bwd_kernel<false, true><<<grid, 256, 0, st>>>(p);
  MSG_CHECK_LAUNCH("styled_act_bwd");
  const int64_t total = (int64_t)4 * B * C;
  styled_act_bwd_reduce_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(sums, p.partial, B, gy, C);
  MSG_CHECK_LAUNCH("styled_act_bwd(reduce)");
  return MSG_OK;
}

// Backward of the demodulation factor d[b,o] = rsqrt(scale^2 sum_c s[b,c]^2 wsq[o,c] + 1e-8) (multi_stylegan_generator.py:386-388):
//   q[b,o]   = gd[b,o] * d[b,o]^3 * (-scale^2 / 2)
//   ds[b,c]  = 2 s[b,c] * sum_o q[b,o] wsq[o,c]
//   dW[o,c,t] = W[o,c,t] * 2 sum_b q[b,o] s[b,c]^2
// Two launches instead of autograd's chain of small [B,C] / [O,C] ATen kernels.  B <= 64.
constexpr int kDemodMaxB = 64;

// block = (32 consecutive c) x 8 o-slices: ds[b, c] for all b
__global__ void __launch_bounds__(256)
demod_bwd_ds_kernel(float* __restrict__ ds, const float* __restrict__ gd, const float* __restrict__ d,
                    const float* __restrict__ s, const float* __restrict__ wsq, int B, int O, int C, float k) {
  extern __shared__ float sm[];                        // q [B][O], then the partial sums [8][B][32]
  float* q = sm;
  float* part = sm + (size_t)B * O;
  for (int i = threadIdx.x; i < B * O; i += 256) { const float dv = __ldg(d + i); q[i] = __ldg(gd + i) * dv * dv * dv * k; }
  __syncthreads();
  const int cl = threadIdx.x & 31, os = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float acc[kDemodMaxB];
#pragma unroll
  for (int b = 0; b < kDemodMaxB; ++b) acc[b] = 0.f;
  if (c < C) {
    for (int o = os; o < O; o += 8) {
      const float w = __ldg(wsq + (int64_t)o * C + c);
#pragma unroll
      for (int b = 0; b < kDemodMaxB; ++b)
        if (b < B) acc[b] = fmaf(q[b * O + o], w, acc[b]);
    }
  }
#pragma unroll
  for (int b = 0; b < kDemodMaxB; ++b)
    if (b < B) part[(os * B + b) * 32 + cl] = acc[b];
  __syncthreads();
  for (int i = threadIdx.x; i < B * 32; i += 256) {
    const int b = i >> 5, cc = blockIdx.x * 32 + (i & 31);
    if (cc < C) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += part[(j * B + b) * 32 + (i & 31)];
      ds[(int64_t)b * C + cc] = 2.f * __ldg(s + (int64_t)b * C + cc) * t;
    }
  }
}

// thread = one (o, c): coef = 2 sum_b q[b,o] s[b,c]^2, dW[o,c,:] = W[o,c,:] * coef
__global__ void __launch_bounds__(256)
demod_bwd_dw_kernel(float* __restrict__ dW, const float* __restrict__ W, const float* __restrict__ gd,
                    const float* __restrict__ d, const float* __restrict__ s, int B, int O, int C, int taps, float k) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)O * C) return;
  const int o = (int)(i / C), c = (int)(i - (int64_t)o * C);
  float coef = 0.f;
  for (int b = 0; b < B; ++b) {
    const float dv = __ldg(d + (int64_t)b * O + o), sv = __ldg(s + (int64_t)b * C + c);
    coef = fmaf(__ldg(gd + (int64_t)b * O + o) * dv * dv * dv * k, sv * sv, coef);
  }
  coef *= 2.f;
  const float* w = W + i * taps;
  float* g = dW + i * taps;
  for (int t = 0; t < taps; ++t) g[t] = __ldg(w + t) * coef;
}

extern "C" int msg_demod_factors_bwd(float* dW, float* ds, const float* gd, const float* d, const float* s, const float* wsq,
                                     const float* W, int B, int O, int C, int taps, float scale, msg_stream_t stream) {
  if (B < 1 || B > kDemodMaxB || O <= 0 || C <= 0 || taps <= 0) return fail(MSG_ERR_UNSUPPORTED, "demod_factors_bwd: sizes (B <= %d)", kDemodMaxB);
  if (!gd || !d || !s) return fail(MSG_ERR_BAD_ARG, "demod_factors_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const float k = -0.5f * scale * scale;
  if (ds) {
    if (!wsq) return fail(MSG_ERR_BAD_ARG, "demod_factors_bwd: ds needs wsq");
    const size_t smem = ((size_t)B * O + (size_t)8 * B * 32) * sizeof(float);
    if (smem > 200 * 1024) return fail(MSG_ERR_UNSUPPORTED, "demod_factors_bwd: B * O too large");
    static bool attr_done[64] = {};
    const int slot = current_device_slot();
    if (!attr_done[slot]) {
      MSG_CHECK_CUDA(cudaFuncSetAttribute(demod_bwd_ds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_done[slot] = true;
    }
    demod_bwd_ds_kernel<<<(unsigned)ceil_div(C, 32), 256, smem, st>>>(ds, gd, d, s, wsq, B, O, C, k);
    MSG_CHECK_LAUNCH("demod_factors_bwd(style)");
  }
  if (dW) {
    if (!W) return fail(MSG_ERR_BAD_ARG, "demod_factors_bwd: dW needs W");
    demod_bwd_dw_kernel<<<(unsigned)ceil_div((int64_t)O * C, 256), 256, 0, st>>>(dW, W, gd, d, s, B, O, C, taps, k);
    MSG_CHECK_LAUNCH("demod_factors_bwd(weights)");
  }
  return MSG_OK;
}

extern "C" int msg_demod_factors(float* d, float* wsq, const float* W, const float* s, int B, int O, int C, int taps,
                                 float scale, msg_stream_t stream) {
  if (B < 0 || O <= 0 || C <= 0 || taps <= 0) return fail(MSG_ERR_BAD_ARG, "demod_factors: bad sizes");
  if (!wsq || !W) return fail(MSG_ERR_BAD_ARG, "demod_factors: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t oc = (int64_t)O * C;
  weight_sq_kernel<<<(unsigned)ceil_div(oc, 256), 256, 0, st>>>(wsq, W, oc, taps);
  MSG_CHECK_LAUNCH("demod_factors(weight squares)");
  if (B == 0) return MSG_OK;
  if (!d || !s) return fail(MSG_ERR_BAD_ARG, "demod_factors: null pointer");
  const int64_t threads = (int64_t)B * O * 32;
  demod_kernel<<<(unsigned)ceil_div(threads, 256), 256, 0, st>>>(d, s, wsq, B, O, C, scale * scale);
  MSG_CHECK_LAUNCH("demod_factors");
  return MSG_OK;
}
