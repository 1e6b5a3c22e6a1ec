// CUDA-core (fp32 FFMA) engine for the two gather GEMMs of conv_common.cuh.
//
// Used for shapes the tcgen05 engine does not tile (4x4 / 8x8 feature maps, N < 16, misaligned rows)
// and as the exact-fp32 on-device cross-check of the tensor-core kernels in the parity tests.
// Classic 64x64x16 shared-memory tiling, 4x4 register micro-tile, 256 threads.
#include "conv_common.cuh"

namespace msg {

constexpr int SM_ = 64, SN_ = 64, SK_ = 16;

__global__ void __launch_bounds__(256)
pixgemm_simt_kernel(const PixGemm g) {
  __shared__ float As[SK_][SM_];       // [k][pixel]
  __shared__ float Bs[SK_][SN_ + 4];   // [k][n]

  const int b = blockIdx.z;
  const int m0 = blockIdx.x * SM_, n0 = blockIdx.y * SN_;
  const int tid = threadIdx.x;
  const int tm = tid & 15, tn = tid >> 4;

  // A loader: fixed pixel per thread
  const int lm = tid & 63, lk = tid >> 6;     // lk in 0..3, loads k = lk + 4*i
  const int pm = m0 + lm;
  const bool pvalid = pm < g.PH * g.PW;
  const int py = pvalid ? pm / g.PW : 0, px = pvalid ? pm - py * g.PW : 0;
  // B loader
  const int ln = tid & 63, lkb = tid >> 6;

  const float* inb = g.in + (int64_t)b * g.is.sb;
  const float* wb = g.w + (int64_t)b * g.w_sb;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < g.ntaps; ++t) {
    const int iy = py * g.my + g.tap_dy[t], ix = px * g.mx + g.tap_dx[t];
    const bool inb_ok = pvalid && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
    const int64_t poff = (int64_t)iy * g.is.sy + (int64_t)ix * g.is.sx;
    const int64_t woff = (int64_t)g.tap_wi[t] * g.w_st;
    for (int c0 = 0; c0 < g.Cr; c0 += SK_) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = lk + 4 * i, c = c0 + k;
        float v = 0.f;
        if (inb_ok && c < g.Cr) v = __ldg(inb + (int64_t)c * g.is.sc + poff);
        As[k][lm] = v;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k = lkb + 4 * i, c = c0 + k, n = n0 + ln;
        float v = 0.f;
        if (c < g.Cr && n < g.N) v = __ldg(wb + (int64_t)n * g.w_sn + (int64_t)c * g.w_sc + woff);
        Bs[k][ln] = v;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SK_; ++k) {
        float a[4], bb[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = As[k][tm + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i) bb[i] = Bs[k][tn * 4 + i];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[j], bb[i], acc[i][j]);
      }
      __syncthreads();
    }
  }

  float* outb = g.out + (int64_t)b * g.os.sb;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int m = m0 + tm + 16 * j;
    if (m >= g.PH * g.PW) continue;
    const int y = m / g.PW, x = m - y * g.PW;
    const int64_t o = (int64_t)(y * g.out_my + g.out_oy) * g.os.sy + (int64_t)(x * g.out_mx + g.out_ox) * g.os.sx;
    const float nz = g.ep.noise ? __ldg(g.ep.noise_w) * __ldg(g.ep.noise + (int64_t)b * g.ep.noise_sb + m) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + tn * 4 + i;
      if (n < g.N) {
        const int64_t off = (int64_t)n * g.os.sc + o;
        const float bn = g.ep.bias ? __ldg(g.ep.bias + n) : 0.f;
        const float av = g.ep.add ? __ldg(g.ep.add + (int64_t)b * g.os.sb + off) : 0.f;
        const float cs = g.ep.cscale ? __ldg(g.ep.cscale + (int64_t)b * g.ep.cscale_sb + n) : 1.f;
        const float o1 = apply_epilogue(g.ep, g.alpha * acc[i][j] * cs, bn, nz, av);
        outb[off] = o1;
        if (g.ep.out2) g.ep.out2[(int64_t)b * g.os.sb + off] = o1 * __ldg(g.ep.out2_scale + (int64_t)b * g.ep.out2_scale_sb + n);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
redgemm_simt_kernel(const RedGemm g) {
  __shared__ float As[SK_][SM_ + 4];   // [pixel k][n]
  __shared__ float Bs[SK_][SN_ + 4];   // [pixel k][c]

  const int n0 = blockIdx.x * SM_, c0 = blockIdx.y * SN_;
  const bool per_sample = g.dw_sb != 0;
  const int t = per_sample ? blockIdx.z % g.ntaps : blockIdx.z;
  const int bfix = per_sample ? blockIdx.z / g.ntaps : 0;
  const int b_begin = per_sample ? bfix : 0, b_end = per_sample ? bfix + 1 : g.B;
  const int dy = g.tap_dy[t], dx = g.tap_dx[t];

  const int tid = threadIdx.x;
  const int tn = tid & 15, tc = tid >> 4;
  const int lk = tid & 15, lr = tid >> 4;   // loader: pixel k, row (n or c) = lr + 16*i
  const int npix = g.PH * g.PW;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int b = b_begin; b < b_end; ++b) {
    const float* gb = g.g + (int64_t)b * g.gs.sb;
    const float* ib = g.in + (int64_t)b * g.is.sb;
    for (int p0 = 0; p0 < npix; p0 += SK_) {
      const int p = p0 + lk;
      const bool pv = p < npix;
      const int y = pv ? p / g.PW : 0, x = pv ? p - y * g.PW : 0;
      const int iy = y * g.my + dy, ix = x * g.mx + dx;
      const bool iv = pv && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
      const int64_t goff = (int64_t)y * g.gs.sy + (int64_t)x * g.gs.sx;
      const int64_t ioff = (int64_t)iy * g.is.sy + (int64_t)ix * g.is.sx;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = lr + 16 * i;
        const int n = n0 + r, c = c0 + r;
        As[lk][r] = (pv && n < g.N) ? __ldg(gb + (int64_t)n * g.gs.sc + goff) : 0.f;
        Bs[lk][r] = (iv && c < g.C) ? __ldg(ib + (int64_t)c * g.is.sc + ioff) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < SK_; ++k) {
        float a[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[k][tn + 16 * i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bb[j] = Bs[k][tc * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  float* dwb = g.dw + (int64_t)bfix * g.dw_sb + (int64_t)g.tap_wi[t] * g.dw_st;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + tn + 16 * i;
    if (n >= g.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tc * 4 + j;
      if (c < g.C) dwb[(int64_t)n * g.dw_sn + (int64_t)c * g.dw_sc] = g.alpha * acc[i][j];
    }
  }
}

int simt_pixgemm(const PixGemm& g, cudaStream_t st) {
  if (g.B <= 0 || g.N <= 0 || g.PH <= 0 || g.PW <= 0) return MSG_OK;
  if (g.ep.add_is_mask || g.ep.colsum) return fail(MSG_ERR_UNSUPPORTED, "conv(simt): mask-mode epilogue is a tcgen05-engine feature");
  dim3 grid((unsigned)ceil_div((int64_t)g.PH * g.PW, SM_), (unsigned)ceil_div(g.N, SN_), (unsigned)g.B);
  if (grid.y > 65535 || grid.z > 65535) return fail(MSG_ERR_UNSUPPORTED, "conv(simt): grid too large");
  pixgemm_simt_kernel<<<grid, 256, 0, st>>>(g);
  MSG_CHECK_LAUNCH("conv pixgemm(simt)");
  return MSG_OK;
}

int simt_redgemm(const RedGemm& g, cudaStream_t st) {
  if (g.N <= 0 || g.C <= 0 || g.ntaps <= 0) return MSG_OK;
  const int z = g.ntaps * (g.dw_sb != 0 ? g.B : 1);
  dim3 grid((unsigned)ceil_div(g.N, SM_), (unsigned)ceil_div(g.C, SN_), (unsigned)z);
  if (grid.y > 65535 || grid.z > 65535) return fail(MSG_ERR_UNSUPPORTED, "conv wgrad(simt): grid too large");
  redgemm_simt_kernel<<<grid, 256, 0, st>>>(g);
  MSG_CHECK_LAUNCH("conv redgemm(simt)");
  return MSG_OK;
}

}  // namespace msg
