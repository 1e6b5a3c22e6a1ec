// C-ABI entry points for the conv primitives and their lowering to the gather GEMMs.
//
//   forward (stride 1)  -> PixGemm on x
//   forward (stride 2)  -> tcgen05: space-to-depth(x) + phase-major weights, then a stride-1 PixGemm
//                          CUDA cores: strided PixGemm directly
//   dgrad   (stride s)  -> s*s PixGemms on dy, one per output phase, scattered with stride s
//                          (this is also the forward of conv_transpose2d, multi_stylegan_generator.py:398)
//   wgrad   (stride 1)  -> RedGemm on (dy, x)
//   wgrad   (stride 2)  -> tcgen05: space-to-depth(x), one RedGemm per input phase; CUDA cores: strided
//
// The tcgen05 engine runs on channels-last (MSG_LAYOUT_NHWC) activations; channel counts that are not a
// multiple of 4 (6-channel images, the +1 minibatch-stddev channel, 3-channel tRGB gradients) are
// zero-padded into the workspace so every TMA stride is a multiple of 16 bytes.  NCHW activations are
// served by the CUDA-core engine.
#include "conv_common.cuh"

namespace msg {

thread_local int g_last_engine = 0;

static inline int fdiv(int a, int b) {  // floor division, b > 0
  int q = a / b;
  if ((a % b != 0) && ((a < 0) != (b < 0))) --q;
  return q;
}
static inline int r4(int v) { return (v + 3) & ~3; }
static inline size_t r256(size_t v) { return (v + 255) & ~(size_t)255; }

// ---- helper kernels (channels-last) ------------------------------------------------------------------
// xs[b, y2, x2, ph*C4 + c] = x[b, 2*y2+py, 2*x2+px, c]  (zero outside the image / for c >= C)
__global__ void __launch_bounds__(256)
s2d_nhwc_kernel(float* __restrict__ xs, const float* __restrict__ x, int B, int C, int C4, int H, int W, int H2,
                int W2) {
  const int64_t total = (int64_t)B * H2 * W2 * 4 * C4;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int64_t r = i;
    const int c = (int)(r % C4); r /= C4;
    const int ph = (int)(r % 4); r /= 4;
    const int x2 = (int)(r % W2); r /= W2;
    const int y2 = (int)(r % H2); r /= H2;
    const int b = (int)r;
    const int iy = 2 * y2 + (ph >> 1), ix = 2 * x2 + (ph & 1);
    float v = 0.f;
    if (c < C && iy < H && ix < W) v = __ldg(x + (((int64_t)b * H + iy) * W + ix) * C + c);
    xs[i] = v;
  }
}

// same, C % 4 == 0 (C4 == C) and 16-byte aligned: one float4 (4 channels) per thread, no per-element div/mod
__global__ void __launch_bounds__(256)
s2d_nhwc_vec_kernel(float4* __restrict__ xs, const float4* __restrict__ x, int B, int Cq, int H, int W, int H2, int W2) {
  // grid.x covers (x2, ph, cq) of one output row (b, y2) = blockIdx.y
  const int row = blockIdx.y;
  const int b = row / H2, y2 = row - b * H2;
  const int per_row = W2 * 4 * Cq;
  float4* orow = xs + (int64_t)row * per_row;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_row; i += gridDim.x * blockDim.x) {
    const int cq = i % Cq;
    const int t = i / Cq;
    const int ph = t & 3, x2 = t >> 2;
    const int iy = 2 * y2 + (ph >> 1), ix = 2 * x2 + (ph & 1);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy < H && ix < W) v = __ldg(x + (((int64_t)b * H + iy) * W + ix) * Cq + cq);
    orow[i] = v;
  }
}

static int launch_s2d(float* xs, const float* x, int B, int C, int C4, int H, int W, int H2, int W2, cudaStream_t st) {
  const bool vec = (C % 4 == 0) && (((reinterpret_cast<uintptr_t>(xs) | reinterpret_cast<uintptr_t>(x)) & 15u) == 0) &&
                   (int64_t)B * H2 <= 65535;
  if (vec) {
    const int per_row = W2 * 4 * (C / 4);
    int gx = (per_row + 255) / 256;
    if (gx > 64) gx = 64;
    dim3 grid((unsigned)gx, (unsigned)(B * H2));
    s2d_nhwc_vec_kernel<<<grid, 256, 0, st>>>((float4*)xs, (const float4*)x, B, C / 4, H, W, H2, W2);
  } else {
    const int64_t tot = (int64_t)B * H2 * W2 * 4 * C4;
    const int64_t want = ceil_div(tot, 256);
    const int64_t cap = (int64_t)num_sms() * 16;
    s2d_nhwc_kernel<<<(unsigned)(want < cap ? (want > 0 ? want : 1) : cap), 256, 0, st>>>(xs, x, B, C, C4, H, W, H2, W2);
  }
  MSG_CHECK_LAUNCH("conv s2d");
  return MSG_OK;
}

// dst[pixel, Cp] = src[pixel, C], zero padded channels
__global__ void __launch_bounds__(256)
chanpad_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t pixels, int C, int Cp) {
  const int64_t total = pixels * Cp;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int c = (int)(i % Cp);
    const int64_t px = i / Cp;
    dst[i] = c < C ? __ldg(src + px * C + c) : 0.f;
  }
}

// w2[bw, n, (ph*C4 + c), a] = w[bw, n, c, ky, kx] with ky = 2*ay + py + pad_h (zero outside the filter)
struct W2Params {
  const float* w;
  int64_t w_sb;     // 0 or O*C*kh*kw
  int64_t w_sn, w_sc;   // element strides of the output- / input-channel index (weight stored [O,C,..] or [C,O,..])
  int BW, N, C, C4, kh, kw, pad_h, pad_w;
  int ay0, ax0, nay, nax;
};
__global__ void __launch_bounds__(256)
s2d_weight_kernel(float* __restrict__ w2, const W2Params p) {
  const int nt = p.nay * p.nax;
  const int64_t total = (int64_t)p.BW * p.N * 4 * p.C4 * nt;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int64_t r = i;
    const int a = (int)(r % nt); r /= nt;
    const int c = (int)(r % p.C4); r /= p.C4;
    const int ph = (int)(r % 4); r /= 4;
    const int n = (int)(r % p.N); r /= p.N;
    const int bw = (int)r;
    const int ay = p.ay0 + a / p.nax, ax = p.ax0 + a % p.nax;
    const int ky = 2 * ay + (ph >> 1) + p.pad_h, kx = 2 * ax + (ph & 1) + p.pad_w;
    float v = 0.f;
    if (c < p.C && ky >= 0 && ky < p.kh && kx >= 0 && kx < p.kw)
      v = __ldg(p.w + bw * p.w_sb + (int64_t)n * p.w_sn + (int64_t)c * p.w_sc + ky * p.kw + kx);
    w2[i] = v;
  }
}

static unsigned grid_for(int64_t total) {
  const int64_t want = ceil_div(total, 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

// ---- validation ------------------------------------------------------------------------------------
static int check_desc(const msg_conv_desc* d, const char* who) {
  if (!d) return fail(MSG_ERR_BAD_ARG, "%s: null descriptor", who);
  if (d->B < 0 || d->C <= 0 || d->O <= 0 || d->H <= 0 || d->W <= 0 || d->kh <= 0 || d->kw <= 0)
    return fail(MSG_ERR_BAD_ARG, "%s: non-positive dimension", who);
  if (d->stride_h != d->stride_w || (d->stride_h != 1 && d->stride_h != 2))
    return fail(MSG_ERR_UNSUPPORTED, "%s: stride (%d,%d) (only 1 or 2, equal)", who, d->stride_h, d->stride_w);
  if (d->pad_h < 0 || d->pad_w < 0) return fail(MSG_ERR_BAD_ARG, "%s: negative padding", who);
  if (d->kh * d->kw > kMaxTaps) return fail(MSG_ERR_UNSUPPORTED, "%s: filter larger than %d taps", who, kMaxTaps);
  if (d->layout != MSG_LAYOUT_NCHW && d->layout != MSG_LAYOUT_NHWC)
    return fail(MSG_ERR_BAD_ARG, "%s: unknown layout %d", who, d->layout);
  if (d->w_transposed != 0 && d->w_transposed != 1) return fail(MSG_ERR_BAD_ARG, "%s: w_transposed must be 0 or 1", who);
  const int oh = (d->H + 2 * d->pad_h - d->kh) / d->stride_h + 1;
  const int ow = (d->W + 2 * d->pad_w - d->kw) / d->stride_w + 1;
  if (d->H + 2 * d->pad_h < d->kh || d->W + 2 * d->pad_w < d->kw || oh != d->OH || ow != d->OW)
    return fail(MSG_ERR_BAD_ARG, "%s: OH/OW (%d,%d) do not match the conv formula (%d,%d)", who, d->OH, d->OW, oh, ow);
  const int64_t wsz = (int64_t)d->O * d->C * d->kh * d->kw;
  if (d->w_batch_stride != 0 && d->w_batch_stride != wsz)
    return fail(MSG_ERR_BAD_ARG, "%s: w_batch_stride must be 0 or O*C*kh*kw", who);
  return MSG_OK;
}

// element strides of (output channel o, input channel c) in the filter tensor: [O,C,kh,kw] or, with
// w_transposed, [C,O,kh,kw] (the layout torch's conv_transpose2d uses and the generator's up-conv hands over)
static inline int64_t w_stride_o(const msg_conv_desc* d) { return d->w_transposed ? (int64_t)d->kh * d->kw : (int64_t)d->C * d->kh * d->kw; }
static inline int64_t w_stride_c(const msg_conv_desc* d) { return d->w_transposed ? (int64_t)d->O * d->kh * d->kw : (int64_t)d->kh * d->kw; }

static inline bool want_tc(const msg_conv_desc* d, int flags) {
  return flags != MSG_CONV_FORCE_SIMT && d->layout == MSG_LAYOUT_NHWC && tc_available();
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static View4 dense_view(int layout, int C, int H, int W) {
  View4 v;
  if (layout == MSG_LAYOUT_NHWC) { v.sb = (int64_t)H * W * C; v.sc = 1; v.sy = (int64_t)W * C; v.sx = C; }
  else { v.sb = (int64_t)C * H * W; v.sc = (int64_t)H * W; v.sy = W; v.sx = 1; }
  return v;
}
static const float* const kAligned = reinterpret_cast<const float*>(256);   // placeholder for support queries

// ---- plans -----------------------------------------------------------------------------------------
// A plan is computed identically by the workspace query and by the call.
struct FwdPlan {
  bool tc, s2d, pad_in;
  bool strided;            // stride 2 through strided phase views of x (no space-to-depth copy)
  int C4, H2, W2;
  int ay0, ax0, nay, nax;
  size_t off_x, off_w2, off_eng, total;
  PixGemm g;
};

static FwdPlan plan_forward(const msg_conv_desc* d, const float* x, const float* w, float* y, float alpha, int flags,
                            const msg_conv_epilogue* ep = nullptr) {
  FwdPlan pl{};
  PixGemm& g = pl.g;
  const int s = d->stride_h;
  const int64_t taps = (int64_t)d->kh * d->kw;
  g.B = d->B; g.N = d->O; g.PH = d->OH; g.PW = d->OW;
  g.out = y; g.os = dense_view(d->layout, d->O, d->OH, d->OW);
  g.out_my = 1; g.out_mx = 1; g.out_oy = 0; g.out_ox = 0; g.alpha = alpha;
  g.ep = Epilogue{};
  g.ep.gain = 1.f;
  if (ep) {
    g.ep.bias = ep->bias; g.ep.noise = ep->noise; g.ep.noise_w = ep->noise_w; g.ep.noise_sb = ep->noise_batch_stride;
    g.ep.add = ep->add; g.ep.act = ep->act; g.ep.slope = ep->slope; g.ep.gain = ep->gain;
    g.ep.cscale = ep->col_scale; g.ep.cscale_sb = ep->col_scale_batch_stride;
    g.ep.out2 = ep->y2; g.ep.out2_scale = ep->y2_scale; g.ep.out2_scale_sb = ep->y2_scale_batch_stride;
  }
  g.in = x; g.Cr = d->C; g.IH = d->H; g.IW = d->W; g.is = dense_view(d->layout, d->C, d->H, d->W);
  g.my = s; g.mx = s;
  g.w = w; g.w_sb = d->w_batch_stride; g.w_sn = w_stride_o(d); g.w_sc = w_stride_c(d); g.w_st = 1;
  g.ntaps = (int)taps;
  for (int ky = 0; ky < d->kh; ++ky)
    for (int kx = 0; kx < d->kw; ++kx) {
      const int t = ky * d->kw + kx;
      g.tap_dy[t] = ky - d->pad_h; g.tap_dx[t] = kx - d->pad_w; g.tap_wi[t] = t;
    }
  size_t off = 0;
  pl.C4 = r4(d->C);
  if (want_tc(d, flags)) {
    if (s == 1) {
      PixGemm t = g;
      pl.pad_in = (d->C % 4 != 0) || !al16(x);
      if (pl.pad_in) { t.in = kAligned; t.is = dense_view(MSG_LAYOUT_NHWC, pl.C4, d->H, d->W); }
      if (tc_pixgemm_supported(t)) {
        pl.tc = true;
        if (pl.pad_in) { pl.off_x = off; off += r256((size_t)d->B * d->H * d->W * pl.C4 * sizeof(float)); }
        g = t;
      } else {
        pl.pad_in = false;
      }
    } else {
      // stride 2: u = k - pad = 2a + ph
      pl.ay0 = fdiv(0 - d->pad_h, 2); pl.ax0 = fdiv(0 - d->pad_w, 2);
      pl.nay = fdiv(d->kh - 1 - d->pad_h, 2) - pl.ay0 + 1;
      pl.nax = fdiv(d->kw - 1 - d->pad_w, 2) - pl.ax0 + 1;
      pl.H2 = (d->H + 1) / 2; pl.W2 = (d->W + 1) / 2;
      const int nt = pl.nay * pl.nax;
      PixGemm t = g;
      t.in = kAligned; t.Cr = 4 * pl.C4; t.IH = pl.H2; t.IW = pl.W2;
      t.is = dense_view(MSG_LAYOUT_NHWC, 4 * pl.C4, pl.H2, pl.W2);
      // whole 32-channel chunks per input phase: read the phases x[b, 2y+py, 2x+px, :] in place through doubled strides
      const bool strided = (d->C % 32 == 0) && al16(x) && d->H >= 2 && d->W >= 2;
      if (strided) {
        t.nphase = 4;
        t.is.sb = (int64_t)d->H * d->W * d->C; t.is.sc = 1; t.is.sy = (int64_t)2 * d->W * d->C; t.is.sx = (int64_t)2 * d->C;
        for (int ph = 0; ph < 4; ++ph) {
          const int py = ph >> 1, px = ph & 1;
          t.in_ph[ph] = x + ((int64_t)py * d->W + px) * d->C;
          t.IH_ph[ph] = (d->H - py + 1) / 2; t.IW_ph[ph] = (d->W - px + 1) / 2;
        }
        t.in = x;
      }
      t.my = 1; t.mx = 1;
      t.ntaps = nt;
      for (int a = 0; a < nt && a < kMaxTaps; ++a) {
        t.tap_dy[a] = pl.ay0 + a / pl.nax; t.tap_dx[a] = pl.ax0 + a % pl.nax; t.tap_wi[a] = a;
      }
      t.w_sn = (int64_t)4 * pl.C4 * nt; t.w_sc = nt; t.w_st = 1;
      t.w_sb = d->w_batch_stride ? (int64_t)d->O * 4 * pl.C4 * nt : 0;
      if (nt <= kMaxTaps && tc_pixgemm_supported(t)) {
        pl.tc = true; pl.s2d = true; pl.strided = strided;
        pl.off_x = off;
        if (!strided) off += r256((size_t)d->B * pl.H2 * pl.W2 * 4 * pl.C4 * sizeof(float));
        pl.off_w2 = off;
        const int64_t BW = d->w_batch_stride ? d->B : 1;
        off += r256((size_t)BW * d->O * 4 * pl.C4 * nt * sizeof(float));
        g = t;
      }
    }
  }
  pl.off_eng = off;
  if (pl.tc) off += r256(tc_pixgemm_workspace(g));
  pl.total = off + 256;
  return pl;
}

struct DgradPlan {
  bool tc, pad_in;
  int O4;
  int nph;                       // s*s phase problems
  PixGemm g[4];
  bool use_tc[4];
  size_t off_x, off_eng, eng_each, off_colsum, total;
};

struct DgradMask { const float* ref; float slope, gain; };

static DgradPlan plan_dgrad(const msg_conv_desc* d, const float* dy, const float* w, float* dx, float alpha, int flags,
                            const float* add = nullptr, const DgradMask* mask = nullptr) {
  DgradPlan pl{};
  const int s = d->stride_h;
  pl.nph = s * s;
  pl.O4 = r4(d->O);
  bool any_tc = false;
  const bool tcw = want_tc(d, flags);
  const bool need_pad = tcw && ((d->O % 4 != 0) || !al16(dy));
  size_t eng = 0;
  for (int ph = 0; ph < pl.nph; ++ph) {
    PixGemm& g = pl.g[ph];
    const int py = ph / s, px = ph % s;
    g.B = d->B; g.N = d->C; g.Cr = d->O;
    g.in = dy; g.IH = d->OH; g.IW = d->OW; g.is = dense_view(d->layout, d->O, d->OH, d->OW);
    g.my = 1; g.mx = 1;
    g.w = w; g.w_sb = d->w_batch_stride; g.w_sn = w_stride_c(d); g.w_sc = w_stride_o(d); g.w_st = 1;
    g.PH = (d->H - py + s - 1) / s; g.PW = (d->W - px + s - 1) / s;
    g.out = dx; g.os = dense_view(d->layout, d->C, d->H, d->W);
    g.out_my = s; g.out_mx = s; g.out_oy = py; g.out_ox = px; g.alpha = alpha;
    g.ep = Epilogue{};
    g.ep.gain = 1.f;
    g.ep.add = add;             // dx = alpha * conv^T(dy, w) + add  (gradient accumulation in the epilogue)
    if (mask) { g.ep.add = mask->ref; g.ep.add_is_mask = 1; g.ep.slope = mask->slope; g.ep.gain = mask->gain; }
    g.ntaps = 0;
    for (int ky = 0; ky < d->kh; ++ky) {
      if ((py + d->pad_h - ky) % s != 0) continue;
      for (int kx = 0; kx < d->kw; ++kx) {
        if ((px + d->pad_w - kx) % s != 0) continue;
        const int t = g.ntaps++;
        g.tap_dy[t] = (py + d->pad_h - ky) / s;
        g.tap_dx[t] = (px + d->pad_w - kx) / s;
        g.tap_wi[t] = ky * d->kw + kx;
      }
    }
    pl.use_tc[ph] = false;
    if (tcw && g.ntaps > 0 && g.PH > 0 && g.PW > 0) {
      PixGemm t = g;
      if (need_pad) { t.in = kAligned; t.is = dense_view(MSG_LAYOUT_NHWC, pl.O4, d->OH, d->OW); }
      if (tc_pixgemm_supported(t)) {
        pl.use_tc[ph] = true; any_tc = true; g = t;
        const size_t e = tc_pixgemm_workspace(g);
        if (e > eng) eng = e;
      }
    }
  }
  pl.tc = any_tc;
  pl.pad_in = any_tc && need_pad;
  size_t off = 0;
  if (pl.pad_in) { pl.off_x = off; off += r256((size_t)d->B * d->OH * d->OW * pl.O4 * sizeof(float)); }
  pl.off_eng = off;
  pl.eng_each = r256(eng);
  // each phase keeps its own transformed weights alive until its kernel has run
  off += (size_t)pl.nph * pl.eng_each;
  pl.off_colsum = off;         // mask mode: per-(CTA, epilogue warp) column sums, [2 * #SMs * 4][C]
  off += r256((size_t)2 * num_sms() * 4 * d->C * sizeof(float));
  pl.total = off + 256;
  return pl;
}

// dbias[n] = sum over the rows of the per-(CTA, warp) partial sums; block = 32 channels x 8 row groups, fixed order
__global__ void __launch_bounds__(256)
colsum_reduce_kernel(float* __restrict__ out, const float* __restrict__ partial, int rows, int N) {
  __shared__ float sm[8][33];
  const int c = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + c;
  float acc = 0.f;
  if (n < N)
    for (int r = j; r < rows; r += 8) acc += partial[(int64_t)r * N + n];
  sm[j][c] = acc;
  __syncthreads();
  if (j == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += sm[k][c];
    out[n] = t;
  }
}

struct WgradPlan {
  bool tc, s2d, pad_g, pad_x;
  bool strided;            // stride 2 on strided phase views of x itself (no space-to-depth copy)
  int nprob, O4, C4, H2, W2;
  RedGemm g[4];
  size_t off_g, off_x, off_eng, eng_each, total;
};

static WgradPlan plan_wgrad(const msg_conv_desc* d, const float* dy, const float* x, float* dw, float alpha, int flags) {
  WgradPlan pl{};
  const int s = d->stride_h;
  const int64_t taps = (int64_t)d->kh * d->kw;
  pl.O4 = r4(d->O); pl.C4 = r4(d->C);
  RedGemm base{};
  base.g = dy; base.B = d->B; base.N = d->O; base.PH = d->OH; base.PW = d->OW;
  base.gs = dense_view(d->layout, d->O, d->OH, d->OW);
  base.in = x; base.C = d->C; base.IH = d->H; base.IW = d->W; base.is = dense_view(d->layout, d->C, d->H, d->W);
  base.my = s; base.mx = s;
  base.dw = dw; base.dw_sb = d->w_batch_stride; base.dw_sn = w_stride_o(d); base.dw_sc = w_stride_c(d); base.dw_st = 1;
  base.alpha = alpha;
  base.ntaps = (int)taps;
  for (int ky = 0; ky < d->kh; ++ky)
    for (int kx = 0; kx < d->kw; ++kx) {
      const int t = ky * d->kw + kx;
      base.tap_dy[t] = ky - d->pad_h; base.tap_dx[t] = kx - d->pad_w; base.tap_wi[t] = t;
    }
  pl.nprob = 1;
  pl.g[0] = base;
  size_t off = 0;
  if (want_tc(d, flags)) {
    RedGemm t = base;
    const bool pg = (d->O % 4 != 0) || !al16(dy);
    if (pg) { t.g = kAligned; t.gs = dense_view(MSG_LAYOUT_NHWC, pl.O4, d->OH, d->OW); }
    if (s == 1) {
      const bool px_ = (d->C % 4 != 0) || !al16(x);
      if (px_) { t.in = kAligned; t.is = dense_view(MSG_LAYOUT_NHWC, pl.C4, d->H, d->W); }
      if (tc_redgemm_supported(t)) {
        pl.tc = true; pl.pad_g = pg; pl.pad_x = px_;
        pl.g[0] = t;
      }
    } else {
      pl.H2 = (d->H + 1) / 2; pl.W2 = (d->W + 1) / 2;
      // Input phase (py, px) of a stride-2 conv is the view x[b, 2*y2 + py, 2*x2 + px, c]: TMA reads it in place through
      // doubled strides when the channel count keeps every address 16-byte aligned; otherwise it is gathered into a
      // space-to-depth copy first.
      const bool strided = (d->C % 4 == 0) && al16(x);
      bool ok = true;
      RedGemm ph_g[4];
      for (int ph = 0; ph < 4 && ok; ++ph) {
        RedGemm q = t;
        const int py = ph >> 1, px = ph & 1;
        q.in = kAligned;
        if (strided) {
          q.IH = (d->H - py + 1) / 2; q.IW = (d->W - px + 1) / 2;
          q.is.sb = (int64_t)d->H * d->W * d->C; q.is.sc = 1; q.is.sy = (int64_t)2 * d->W * d->C; q.is.sx = (int64_t)2 * d->C;
        } else {
          q.IH = pl.H2; q.IW = pl.W2; q.is = dense_view(MSG_LAYOUT_NHWC, 4 * pl.C4, pl.H2, pl.W2);
        }
        q.my = 1; q.mx = 1;
        q.ntaps = 0;
        for (int ky = 0; ky < d->kh; ++ky) {
          const int u = ky - d->pad_h;
          if (((u % 2) + 2) % 2 != py) continue;
          for (int kx = 0; kx < d->kw; ++kx) {
            const int v = kx - d->pad_w;
            if (((v % 2) + 2) % 2 != px) continue;
            const int tt = q.ntaps++;
            q.tap_dy[tt] = fdiv(u, 2); q.tap_dx[tt] = fdiv(v, 2); q.tap_wi[tt] = ky * d->kw + kx;
          }
        }
        ph_g[ph] = q;
        if (q.ntaps > 0 && !tc_redgemm_supported(q)) ok = false;
      }
      if (ok) {
        pl.tc = true; pl.s2d = !strided; pl.strided = strided; pl.pad_g = pg;
        pl.nprob = 4;
        for (int ph = 0; ph < 4; ++ph) pl.g[ph] = ph_g[ph];
      }
    }
  }
  if (pl.tc) {
    if (pl.pad_g) { pl.off_g = off; off += r256((size_t)d->B * d->OH * d->OW * pl.O4 * sizeof(float)); }
    if (pl.s2d) { pl.off_x = off; off += r256((size_t)d->B * pl.H2 * pl.W2 * 4 * pl.C4 * sizeof(float)); }
    else if (pl.pad_x) { pl.off_x = off; off += r256((size_t)d->B * d->H * d->W * pl.C4 * sizeof(float)); }
    size_t eng = 0;
    for (int i = 0; i < pl.nprob; ++i)
      if (pl.g[i].ntaps > 0) { const size_t e = tc_redgemm_workspace(pl.g[i]); if (e > eng) eng = e; }
    pl.eng_each = r256(eng);
    pl.off_eng = off;
    off += (size_t)pl.nprob * pl.eng_each;
  }
  pl.total = off + 256;
  return pl;
}

static inline uint8_t* ws_base(void* ws) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
}

}  // namespace msg

using namespace msg;

extern "C" int msg_conv2d_last_engine(void) { return g_last_engine; }

extern "C" size_t msg_conv2d_workspace(const msg_conv_desc* d, int which, int flags) {
  if (check_desc(d, "conv2d_workspace")) return 0;
  // pointers only matter for their 16-byte alignment; assume aligned allocations (re-checked at call time)
  float* al = const_cast<float*>(kAligned);
  if (which == 0) return plan_forward(d, al, al, al, 1.f, flags).total;
  if (which == 1) return plan_dgrad(d, al, al, al, 1.f, flags).total;
  if (which == 2) return plan_wgrad(d, al, al, al, 1.f, flags).total;
  return 0;
}

extern "C" int msg_conv2d_forward(float* y, const float* x, const float* w, const msg_conv_desc* d, float alpha,
                                  void* workspace, size_t workspace_bytes, int flags, msg_stream_t stream) {
  return msg_conv2d_forward_fused(y, x, w, d, alpha, nullptr, workspace, workspace_bytes, flags, stream);
}

extern "C" int msg_conv2d_forward_fused(float* y, const float* x, const float* w, const msg_conv_desc* d, float alpha,
                                        const msg_conv_epilogue* ep, void* workspace, size_t workspace_bytes, int flags,
                                        msg_stream_t stream) {
  int rc = check_desc(d, "conv2d_forward");
  if (rc) return rc;
  if (d->B == 0) return MSG_OK;
  if (!y || !x || !w) return fail(MSG_ERR_BAD_ARG, "conv2d_forward: null pointer");
  if (ep) {
    if (ep->noise && !ep->noise_w) return fail(MSG_ERR_BAD_ARG, "conv2d_forward: noise epilogue needs noise_w");
    if (ep->act != 0 && ep->act != 1) return fail(MSG_ERR_BAD_ARG, "conv2d_forward: epilogue act must be 0 or 1");
    if (ep->noise && ep->noise_batch_stride != 0 && ep->noise_batch_stride != (int64_t)d->OH * d->OW)
      return fail(MSG_ERR_BAD_ARG, "conv2d_forward: noise_batch_stride must be 0 or OH*OW");
    if (ep->y2 && !ep->y2_scale) return fail(MSG_ERR_BAD_ARG, "conv2d_forward: y2 needs y2_scale");
    if ((ep->y2 || ep->col_scale) && ep->add)
      return fail(MSG_ERR_UNSUPPORTED, "conv2d_forward: col_scale / y2 cannot be combined with a residual operand");
    if ((ep->col_scale && ep->col_scale_batch_stride != 0 && ep->col_scale_batch_stride != d->O) ||
        (ep->y2 && ep->y2_scale_batch_stride != 0 && ep->y2_scale_batch_stride != d->O))
      return fail(MSG_ERR_BAD_ARG, "conv2d_forward: channel-scale batch strides must be 0 or O");
  }
  cudaStream_t st = (cudaStream_t)stream;
  FwdPlan pl = plan_forward(d, x, w, y, alpha, flags, ep);
  if (!pl.tc) {
    if (flags == MSG_CONV_FORCE_TC) return fail(MSG_ERR_UNSUPPORTED, "conv2d_forward: not eligible for tcgen05 (needs NHWC)");
    g_last_engine = 1;
    return simt_pixgemm(pl.g, st);
  }
  if (!workspace || workspace_bytes < pl.total)
    return fail(MSG_ERR_WORKSPACE, "conv2d_forward: workspace %zu < %zu", workspace_bytes, pl.total);
  uint8_t* ws = ws_base(workspace);
  if (pl.s2d) {
    float* xs = reinterpret_cast<float*>(ws + pl.off_x);
    float* w2 = reinterpret_cast<float*>(ws + pl.off_w2);
    if (!pl.strided) {
      rc = launch_s2d(xs, x, d->B, d->C, pl.C4, d->H, d->W, pl.H2, pl.W2, st);
      if (rc) return rc;
    }
    W2Params wp{};
    wp.w = w; wp.w_sb = d->w_batch_stride; wp.BW = d->w_batch_stride ? d->B : 1; wp.N = d->O; wp.C = d->C;
    wp.w_sn = w_stride_o(d); wp.w_sc = w_stride_c(d);
    wp.C4 = pl.C4; wp.kh = d->kh; wp.kw = d->kw; wp.pad_h = d->pad_h; wp.pad_w = d->pad_w;
    wp.ay0 = pl.ay0; wp.ax0 = pl.ax0; wp.nay = pl.nay; wp.nax = pl.nax;
    const int64_t wtot = (int64_t)wp.BW * d->O * 4 * pl.C4 * pl.nay * pl.nax;
    s2d_weight_kernel<<<grid_for(wtot), 256, 0, st>>>(w2, wp);
    MSG_CHECK_LAUNCH("conv s2d weights");
    if (!pl.strided) pl.g.in = xs;
    pl.g.w = w2;
  } else if (pl.pad_in) {
    float* xp = reinterpret_cast<float*>(ws + pl.off_x);
    const int64_t pixels = (int64_t)d->B * d->H * d->W;
    chanpad_kernel<<<grid_for(pixels * pl.C4), 256, 0, st>>>(xp, x, pixels, d->C, pl.C4);
    MSG_CHECK_LAUNCH("conv channel pad");
    pl.g.in = xp;
  }
  g_last_engine = 2;
  return tc_pixgemm(pl.g, ws + pl.off_eng, r256(tc_pixgemm_workspace(pl.g)), st);
}

// Convolution of the channel concatenation [x1 | x2] without materialising it (the U-Net decoder's
// torch.cat([upsampled, skip], dim=1) followed by a ResNetBlock, u_net_2d_discriminator.py:137,174-186): the K loop of
// the implicit GEMM takes its first c1 / 32 chunks from x1 and the rest from x2.
extern "C" int msg_conv2d_forward_cat2(float* y, const float* x1, int c1, const float* x2, const float* w,
                                       const msg_conv_desc* d, float alpha, const msg_conv_epilogue* ep, void* workspace,
                                       size_t workspace_bytes, int flags, msg_stream_t stream) {
  int rc = check_desc(d, "conv2d_forward_cat2");
  if (rc) return rc;
  if (d->B == 0) return MSG_OK;
  if (!y || !x1 || !x2 || !w) return fail(MSG_ERR_BAD_ARG, "conv2d_forward_cat2: null pointer");
  const int c2 = d->C - c1;
  if (c1 <= 0 || c2 <= 0) return fail(MSG_ERR_BAD_ARG, "conv2d_forward_cat2: c1 = %d of C = %d", c1, d->C);
  if (ep) {
    if (ep->noise && !ep->noise_w) return fail(MSG_ERR_BAD_ARG, "conv2d_forward_cat2: noise epilogue needs noise_w");
    if (ep->act != 0 && ep->act != 1) return fail(MSG_ERR_BAD_ARG, "conv2d_forward_cat2: epilogue act must be 0 or 1");
  }
  if (d->layout != MSG_LAYOUT_NHWC || d->stride_h != 1 || c1 % 32 != 0 || c2 % 32 != 0 || !al16(x1) || !al16(x2) ||
      flags == MSG_CONV_FORCE_SIMT || !tc_available())
    return fail(MSG_ERR_UNSUPPORTED, "conv2d_forward_cat2: needs the tcgen05 engine, NHWC, stride 1 and channel counts "
                                     "that are multiples of 32 (got %d + %d)", c1, c2);
  cudaStream_t st = (cudaStream_t)stream;
  FwdPlan pl = plan_forward(d, x1, w, y, alpha, flags, ep);
  if (!pl.tc || pl.pad_in || pl.s2d) return fail(MSG_ERR_UNSUPPORTED, "conv2d_forward_cat2: shape not eligible");
  if (!workspace || workspace_bytes < pl.total)
    return fail(MSG_ERR_WORKSPACE, "conv2d_forward_cat2: workspace %zu < %zu", workspace_bytes, pl.total);
  pl.g.nsrc = 2;
  pl.g.in_src[0] = x1; pl.g.in_src[1] = x2;
  pl.g.C_src[0] = c1; pl.g.C_src[1] = c2;
  if (!tc_pixgemm_supported(pl.g)) return fail(MSG_ERR_UNSUPPORTED, "conv2d_forward_cat2: shape not eligible");
  uint8_t* ws = ws_base(workspace);
  g_last_engine = 2;
  return tc_pixgemm(pl.g, ws + pl.off_eng, r256(tc_pixgemm_workspace(pl.g)), st);
}

extern "C" int msg_conv2d_dgrad(float* dx, const float* dy, const float* w, const msg_conv_desc* d, float alpha,
                                void* workspace, size_t workspace_bytes, int flags, msg_stream_t stream) {
  return msg_conv2d_dgrad_acc(dx, dy, w, d, alpha, nullptr, workspace, workspace_bytes, flags, stream);
}

extern "C" int msg_conv2d_dgrad_acc(float* dx, const float* dy, const float* w, const msg_conv_desc* d, float alpha,
                                    const float* add, void* workspace, size_t workspace_bytes, int flags,
                                    msg_stream_t stream) {
  int rc = check_desc(d, "conv2d_dgrad");
  if (rc) return rc;
  if (d->B == 0) return MSG_OK;
  if (!dx || !dy || !w) return fail(MSG_ERR_BAD_ARG, "conv2d_dgrad: null pointer");
  if (add == dx) return fail(MSG_ERR_BAD_ARG, "conv2d_dgrad: `add` must not alias dx");
  cudaStream_t st = (cudaStream_t)stream;
  DgradPlan pl = plan_dgrad(d, dy, w, dx, alpha, flags, add);
  if (!pl.tc && flags == MSG_CONV_FORCE_TC)
    return fail(MSG_ERR_UNSUPPORTED, "conv2d_dgrad: not eligible for tcgen05 (needs NHWC)");
  uint8_t* ws = nullptr;
  if (pl.tc) {
    if (!workspace || workspace_bytes < pl.total)
      return fail(MSG_ERR_WORKSPACE, "conv2d_dgrad: workspace %zu < %zu", workspace_bytes, pl.total);
    ws = ws_base(workspace);
  }
  const float* dyp = dy;
  if (pl.pad_in) {
    float* p = reinterpret_cast<float*>(ws + pl.off_x);
    const int64_t pixels = (int64_t)d->B * d->OH * d->OW;
    chanpad_kernel<<<grid_for(pixels * pl.O4), 256, 0, st>>>(p, dy, pixels, d->O, pl.O4);
    MSG_CHECK_LAUNCH("conv channel pad");
    dyp = p;
  }
  g_last_engine = pl.tc ? 2 : 1;
  for (int ph = 0; ph < pl.nph; ++ph) {
    PixGemm& g = pl.g[ph];
    if (g.PH <= 0 || g.PW <= 0) continue;
    if (pl.use_tc[ph]) {
      g.in = dyp;
      rc = tc_pixgemm(g, ws + pl.off_eng + ph * pl.eng_each, pl.eng_each, st);
    } else {
      rc = simt_pixgemm(g, st);   // also zero-fills phases that no filter tap reaches
    }
    if (rc) return rc;
  }
  return MSG_OK;
}

// dx = alpha * conv^T(dy, w) * (ref > 0 ? 1 : slope) * gain  and (optionally) dbias[c] = sum over pixels of dx: the dgrad of
// a layer fused with the activation backward of the layer that PRODUCED its input (ref = that layer's output) — inside a
// residual block g1_pre = fba_bwd(dgrad(g2_pre, w2), h1) is one kernel instead of a GEMM plus a pass over the activation
// (op_static/fused_act.py:31-40 behind u_net_2d_discriminator.py:174-186).
static bool dgrad_mask_plan_ok(const msg_conv_desc* d, const DgradPlan& pl) {
  return d->stride_h == 1 && pl.nph == 1 && pl.tc && pl.use_tc[0] && !pl.pad_in && tc_pixgemm_mask_ok(pl.g[0]);
}

extern "C" int msg_conv2d_dgrad_mask_supported(const msg_conv_desc* d, int flags) {
  if (check_desc(d, "conv2d_dgrad_mask")) return 0;
  if (d->B == 0) return 0;
  float* al = const_cast<float*>(kAligned);
  DgradMask m{al, 0.2f, 1.f};
  DgradPlan pl = plan_dgrad(d, al, al, al, 1.f, flags, nullptr, &m);
  return dgrad_mask_plan_ok(d, pl) ? 1 : 0;
}

extern "C" int msg_conv2d_dgrad_mask(float* dx, float* dbias, const float* dy, const float* w, const msg_conv_desc* d,
                                     float alpha, const float* ref, float slope, float gain, void* workspace,
                                     size_t workspace_bytes, int flags, msg_stream_t stream) {
  int rc = check_desc(d, "conv2d_dgrad_mask");
  if (rc) return rc;
  if (d->B == 0) return MSG_OK;
  if (!dx || !dy || !w || !ref) return fail(MSG_ERR_BAD_ARG, "conv2d_dgrad_mask: null pointer");
  if (ref == dx) return fail(MSG_ERR_BAD_ARG, "conv2d_dgrad_mask: `ref` must not alias dx");
  cudaStream_t st = (cudaStream_t)stream;
  DgradMask m{ref, slope, gain};
  DgradPlan pl = plan_dgrad(d, dy, w, dx, alpha, flags, nullptr, &m);
  if (!dgrad_mask_plan_ok(d, pl))
    return fail(MSG_ERR_UNSUPPORTED, "conv2d_dgrad_mask: needs the tcgen05 engine, NHWC, stride 1, C % 4 == 0 "
                                     "(query msg_conv2d_dgrad_mask_supported)");
  if (!workspace || workspace_bytes < pl.total)
    return fail(MSG_ERR_WORKSPACE, "conv2d_dgrad_mask: workspace %zu < %zu", workspace_bytes, pl.total);
  uint8_t* ws = ws_base(workspace);
  PixGemm& g = pl.g[0];
  const int rows = 2 * num_sms() * 4;
  float* partial = reinterpret_cast<float*>(ws + pl.off_colsum);
  if (dbias) {
    MSG_CHECK_CUDA(cudaMemsetAsync(partial, 0, (size_t)rows * d->C * sizeof(float), st));
    g.ep.colsum = partial;
  }
  g.in = dy;
  g_last_engine = 2;
  rc = tc_pixgemm(g, ws + pl.off_eng, pl.eng_each, st);
  if (rc) return rc;
  if (dbias) {
    colsum_reduce_kernel<<<(unsigned)((d->C + 31) / 32), 256, 0, st>>>(dbias, partial, rows, d->C);
    MSG_CHECK_LAUNCH("conv2d_dgrad_mask(bias gradient)");
  }
  return MSG_OK;
}

extern "C" int msg_conv2d_wgrad(float* dw, const float* dy, const float* x, const msg_conv_desc* d, float alpha,
                                void* workspace, size_t workspace_bytes, int flags, msg_stream_t stream) {
  int rc = check_desc(d, "conv2d_wgrad");
  if (rc) return rc;
  if (!dw) return fail(MSG_ERR_BAD_ARG, "conv2d_wgrad: null dw");
  cudaStream_t st = (cudaStream_t)stream;
  if (d->B == 0) {
    if (d->w_batch_stride == 0)
      MSG_CHECK_CUDA(cudaMemsetAsync(dw, 0, (size_t)d->O * d->C * d->kh * d->kw * sizeof(float), st));
    return MSG_OK;
  }
  if (!dy || !x) return fail(MSG_ERR_BAD_ARG, "conv2d_wgrad: null pointer");
  WgradPlan pl = plan_wgrad(d, dy, x, dw, alpha, flags);
  if (!pl.tc) {
    if (flags == MSG_CONV_FORCE_TC) return fail(MSG_ERR_UNSUPPORTED, "conv2d_wgrad: not eligible for tcgen05 (needs NHWC)");
    g_last_engine = 1;
    return simt_redgemm(pl.g[0], st);
  }
  if (!workspace || workspace_bytes < pl.total)
    return fail(MSG_ERR_WORKSPACE, "conv2d_wgrad: workspace %zu < %zu", workspace_bytes, pl.total);
  uint8_t* ws = ws_base(workspace);
  const float* gp = dy;
  const float* xp = x;
  if (pl.pad_g) {
    float* p = reinterpret_cast<float*>(ws + pl.off_g);
    const int64_t pixels = (int64_t)d->B * d->OH * d->OW;
    chanpad_kernel<<<grid_for(pixels * pl.O4), 256, 0, st>>>(p, dy, pixels, d->O, pl.O4);
    MSG_CHECK_LAUNCH("conv channel pad");
    gp = p;
  }
  if (pl.s2d) {
    float* xs = reinterpret_cast<float*>(ws + pl.off_x);
    rc = launch_s2d(xs, x, d->B, d->C, pl.C4, d->H, d->W, pl.H2, pl.W2, st);
    if (rc) return rc;
    xp = xs;
  } else if (pl.pad_x) {
    float* p = reinterpret_cast<float*>(ws + pl.off_x);
    const int64_t pixels = (int64_t)d->B * d->H * d->W;
    chanpad_kernel<<<grid_for(pixels * pl.C4), 256, 0, st>>>(p, x, pixels, d->C, pl.C4);
    MSG_CHECK_LAUNCH("conv channel pad");
    xp = p;
  }
  g_last_engine = 2;
  for (int i = 0; i < pl.nprob; ++i) {
    RedGemm& g = pl.g[i];
    if (g.ntaps == 0) continue;
    g.g = gp;
    if (pl.s2d) g.in = xp + (int64_t)i * pl.C4;       // phase i = channel block i of the s2d tensor
    else if (pl.strided) g.in = x + ((int64_t)(i >> 1) * d->W + (i & 1)) * d->C;   // phase i = strided view of x
    else g.in = xp;
    rc = tc_redgemm(g, ws + pl.off_eng + i * pl.eng_each, pl.eng_each, st);
    if (rc) return rc;
  }
  return MSG_OK;
}
