// C-ABI entry points for the conv primitives and their lowering to the gather GEMMs.
//
//   forward (stride 1)  -> PixGemm on x
//   forward (stride 2)  -> tcgen05: space-to-depth(x) + phase-major weights, then a stride-1 PixGemm
//                          CUDA cores: strided PixGemm directly
//   dgrad   (stride s)  -> s*s PixGemms on dy, one per output phase, scattered with stride s
//                          (this is also the forward of conv_transpose2d, multi_stylegan_generator.py:398)
//   wgrad   (stride 1)  -> RedGemm on (dy, x)
//   wgrad   (stride 2)  -> tcgen05: space-to-depth(x), one RedGemm per input phase; CUDA cores: strided
// Rows whose pitch is not a multiple of 16 bytes (the 127x127 maps after the discriminator's stride-2
// convs, u_net_2d_discriminator.py:59-63) are re-pitched into the workspace so TMA can address them.
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

// ---- helper kernels --------------------------------------------------------------------------------
// xs[b, (py*2+px)*C + c, y2, x2 (pitch W2p)] = x[b, c, 2*y2+py, 2*x2+px]  (zero outside)
__global__ void __launch_bounds__(256)
s2d_kernel(float* __restrict__ xs, const float* __restrict__ x, int B, int C, int H, int W, int H2, int W2p) {
  const int64_t total = (int64_t)B * 4 * C * H2 * W2p;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int64_t r = i;
    const int x2 = (int)(r % W2p); r /= W2p;
    const int y2 = (int)(r % H2); r /= H2;
    const int c = (int)(r % C); r /= C;
    const int ph = (int)(r % 4); r /= 4;
    const int b = (int)r;
    const int iy = 2 * y2 + (ph >> 1), ix = 2 * x2 + (ph & 1);
    float v = 0.f;
    if (iy < H && ix < W) v = __ldg(x + (((int64_t)b * C + c) * H + iy) * W + ix);
    xs[i] = v;
  }
}

// dst[plane, y, x (pitch Wp)] = src[plane, y, x (pitch W)], zero padded columns
__global__ void __launch_bounds__(256)
repitch_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t planes, int H, int W, int Wp) {
  const int64_t total = planes * H * Wp;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int x = (int)(i % Wp);
    const int64_t row = i / Wp;
    dst[i] = x < W ? __ldg(src + row * W + x) : 0.f;
  }
}

// w2[bw, n, (ph*C + c), a] = w[bw, n, c, ky, kx] with ky = 2*ay + py + pad_h (zero if outside the filter)
struct W2Params {
  const float* w;
  int64_t w_sb;     // 0 or O*C*kh*kw
  int BW, N, C, kh, kw, pad_h, pad_w;
  int ay0, ax0, nay, nax;
};
__global__ void __launch_bounds__(256)
s2d_weight_kernel(float* __restrict__ w2, const W2Params p) {
  const int nt = p.nay * p.nax;
  const int64_t total = (int64_t)p.BW * p.N * 4 * p.C * nt;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int64_t r = i;
    const int a = (int)(r % nt); r /= nt;
    const int c = (int)(r % p.C); r /= p.C;
    const int ph = (int)(r % 4); r /= 4;
    const int n = (int)(r % p.N); r /= p.N;
    const int bw = (int)r;
    const int ay = p.ay0 + a / p.nax, ax = p.ax0 + a % p.nax;
    const int ky = 2 * ay + (ph >> 1) + p.pad_h, kx = 2 * ax + (ph & 1) + p.pad_w;
    float v = 0.f;
    if (ky >= 0 && ky < p.kh && kx >= 0 && kx < p.kw)
      v = __ldg(p.w + bw * p.w_sb + (((int64_t)n * p.C + c) * p.kh + ky) * p.kw + kx);
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
  const int oh = (d->H + 2 * d->pad_h - d->kh) / d->stride_h + 1;
  const int ow = (d->W + 2 * d->pad_w - d->kw) / d->stride_w + 1;
  if (d->H + 2 * d->pad_h < d->kh || d->W + 2 * d->pad_w < d->kw || oh != d->OH || ow != d->OW)
    return fail(MSG_ERR_BAD_ARG, "%s: OH/OW (%d,%d) do not match the conv formula (%d,%d)", who, d->OH, d->OW, oh, ow);
  const int64_t wsz = (int64_t)d->O * d->C * d->kh * d->kw;
  if (d->w_batch_stride != 0 && d->w_batch_stride != wsz)
    return fail(MSG_ERR_BAD_ARG, "%s: w_batch_stride must be 0 or O*C*kh*kw", who);
  return MSG_OK;
}

static inline bool want_tc(int flags) { return flags != MSG_CONV_FORCE_SIMT && tc_available(); }
static inline bool tma_ok_rows(const void* p, int W) { return (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 15u) == 0); }

// ---- plans -----------------------------------------------------------------------------------------
// A plan is computed identically by the workspace query and by the call.
struct FwdPlan {
  bool tc;
  bool s2d;         // stride-2 lowering
  bool repitch;     // stride-1 input rows need re-pitching
  int H2, W2p;      // s2d buffer geometry
  int ay0, ax0, nay, nax;
  size_t off_x, off_w2, off_eng, total;
  PixGemm g;
};

static void fill_taps_fwd(PixGemm& g, const msg_conv_desc* d) {
  g.ntaps = d->kh * d->kw;
  for (int ky = 0; ky < d->kh; ++ky)
    for (int kx = 0; kx < d->kw; ++kx) {
      const int t = ky * d->kw + kx;
      g.tap_dy[t] = ky - d->pad_h;
      g.tap_dx[t] = kx - d->pad_w;
      g.tap_wi[t] = t;
    }
}

static FwdPlan plan_forward(const msg_conv_desc* d, const float* x, const float* w, float* y, float alpha, int flags) {
  FwdPlan pl{};
  PixGemm& g = pl.g;
  const int s = d->stride_h;
  const int64_t taps = (int64_t)d->kh * d->kw;
  g.B = d->B; g.N = d->O; g.PH = d->OH; g.PW = d->OW;
  g.out = y; g.out_sb = (int64_t)d->O * d->OH * d->OW; g.out_sn = (int64_t)d->OH * d->OW; g.out_pitch = d->OW;
  g.out_sy = 1; g.out_sx = 1; g.out_oy = 0; g.out_ox = 0; g.alpha = alpha;
  // direct (CUDA-core capable) formulation
  g.in = x; g.Cr = d->C; g.IH = d->H; g.IW = d->W;
  g.in_sb = (int64_t)d->C * d->H * d->W; g.in_sc = (int64_t)d->H * d->W; g.in_pitch = d->W;
  g.in_sy = s; g.in_sx = s;
  g.w = w; g.w_sb = d->w_batch_stride; g.w_sn = d->C * taps; g.w_sc = taps; g.w_st = 1;
  fill_taps_fwd(g, d);
  size_t off = 0;
  if (want_tc(flags)) {
    if (s == 1) {
      PixGemm t = g;
      if (!tma_ok_rows(x, d->W)) {
        pl.repitch = true;
        const int Wp = r4(d->W);
        t.in = reinterpret_cast<const float*>(16);  // aligned placeholder for the support query
        t.in_pitch = Wp; t.in_sc = (int64_t)d->H * Wp; t.in_sb = (int64_t)d->C * d->H * Wp;
      }
      if (tc_pixgemm_supported(t)) {
        pl.tc = true;
        if (pl.repitch) {
          pl.off_x = off;
          off += r256((size_t)d->B * d->C * d->H * r4(d->W) * sizeof(float));
        }
        g = t;
      } else {
        pl.repitch = false;
      }
    } else {
      // stride 2: u = k - pad = 2a + ph
      pl.ay0 = fdiv(0 - d->pad_h, 2); pl.ax0 = fdiv(0 - d->pad_w, 2);
      pl.nay = fdiv(d->kh - 1 - d->pad_h, 2) - pl.ay0 + 1;
      pl.nax = fdiv(d->kw - 1 - d->pad_w, 2) - pl.ax0 + 1;
      pl.H2 = (d->H + 1) / 2;
      const int W2 = (d->W + 1) / 2;
      pl.W2p = r4(W2);
      PixGemm t = g;
      t.in = reinterpret_cast<const float*>(16);
      t.Cr = 4 * d->C; t.IH = pl.H2; t.IW = W2;
      t.in_pitch = pl.W2p; t.in_sc = (int64_t)pl.H2 * pl.W2p; t.in_sb = (int64_t)4 * d->C * pl.H2 * pl.W2p;
      t.in_sy = 1; t.in_sx = 1;
      const int nt = pl.nay * pl.nax;
      t.ntaps = nt;
      for (int a = 0; a < nt; ++a) {
        t.tap_dy[a] = pl.ay0 + a / pl.nax;
        t.tap_dx[a] = pl.ax0 + a % pl.nax;
        t.tap_wi[a] = a;
      }
      t.w_sn = (int64_t)4 * d->C * nt; t.w_sc = nt; t.w_st = 1;
      t.w_sb = d->w_batch_stride ? (int64_t)d->O * 4 * d->C * nt : 0;
      if (nt <= kMaxTaps && tc_pixgemm_supported(t)) {
        pl.tc = true; pl.s2d = true;
        pl.off_x = off;
        off += r256((size_t)d->B * 4 * d->C * pl.H2 * pl.W2p * sizeof(float));
        pl.off_w2 = off;
        const int64_t BW = d->w_batch_stride ? d->B : 1;
        off += r256((size_t)BW * d->O * 4 * d->C * nt * sizeof(float));
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
  bool tc, repitch;
  int nph;                       // s*s phase problems
  PixGemm g[4];
  bool use_tc[4];
  size_t off_x, off_eng, total;
};

static DgradPlan plan_dgrad(const msg_conv_desc* d, const float* dy, const float* w, float* dx, float alpha, int flags) {
  DgradPlan pl{};
  const int s = d->stride_h;
  const int64_t taps = (int64_t)d->kh * d->kw;
  pl.nph = s * s;
  bool any_tc = false;
  const bool tcw = want_tc(flags);
  const bool need_repitch = tcw && !tma_ok_rows(dy, d->OW);
  size_t eng = 0;
  for (int ph = 0; ph < pl.nph; ++ph) {
    PixGemm& g = pl.g[ph];
    const int py = ph / s, px = ph % s;
    g.B = d->B; g.N = d->C; g.Cr = d->O;
    g.in = dy; g.IH = d->OH; g.IW = d->OW;
    g.in_sb = (int64_t)d->O * d->OH * d->OW; g.in_sc = (int64_t)d->OH * d->OW; g.in_pitch = d->OW;
    g.in_sy = 1; g.in_sx = 1;
    g.w = w; g.w_sb = d->w_batch_stride; g.w_sn = taps; g.w_sc = d->C * taps; g.w_st = 1;
    g.PH = (d->H - py + s - 1) / s; g.PW = (d->W - px + s - 1) / s;
    g.out = dx; g.out_sb = (int64_t)d->C * d->H * d->W; g.out_sn = (int64_t)d->H * d->W; g.out_pitch = d->W;
    g.out_sy = s; g.out_sx = s; g.out_oy = py; g.out_ox = px; g.alpha = alpha;
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
      if (need_repitch) {
        const int Wp = r4(d->OW);
        t.in = reinterpret_cast<const float*>(16);
        t.in_pitch = Wp; t.in_sc = (int64_t)d->OH * Wp; t.in_sb = (int64_t)d->O * d->OH * Wp;
      }
      if (tc_pixgemm_supported(t)) {
        pl.use_tc[ph] = true; any_tc = true; g = t;
        const size_t e = tc_pixgemm_workspace(g);
        if (e > eng) eng = e;
      }
    }
  }
  pl.tc = any_tc;
  pl.repitch = any_tc && need_repitch;
  size_t off = 0;
  if (pl.repitch) { pl.off_x = off; off += r256((size_t)d->B * d->O * d->OH * r4(d->OW) * sizeof(float)); }
  pl.off_eng = off;
  // the phases run back to back on one stream but each needs its own transformed weights alive until
  // its kernel has run; give every phase a private slice.
  off += (size_t)pl.nph * r256(eng);
  pl.total = off + 256;
  return pl;
}

struct WgradPlan {
  bool tc, s2d, repitch_g, repitch_x;
  int nprob;
  RedGemm g[4];
  int H2, W2p;
  size_t off_g, off_x, off_eng, eng_each, total;
};

static WgradPlan plan_wgrad(const msg_conv_desc* d, const float* dy, const float* x, float* dw, float alpha, int flags) {
  WgradPlan pl{};
  const int s = d->stride_h;
  const int64_t taps = (int64_t)d->kh * d->kw;
  RedGemm base{};
  base.g = dy; base.B = d->B; base.N = d->O; base.PH = d->OH; base.PW = d->OW;
  base.g_sb = (int64_t)d->O * d->OH * d->OW; base.g_sn = (int64_t)d->OH * d->OW; base.g_pitch = d->OW;
  base.in = x; base.C = d->C; base.IH = d->H; base.IW = d->W;
  base.in_sb = (int64_t)d->C * d->H * d->W; base.in_sc = (int64_t)d->H * d->W; base.in_pitch = d->W;
  base.in_sy = s; base.in_sx = s;
  base.dw = dw; base.dw_sb = d->w_batch_stride; base.dw_sn = d->C * taps; base.dw_sc = taps; base.dw_st = 1;
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
  if (want_tc(flags)) {
    RedGemm t = base;
    const bool rg = !tma_ok_rows(dy, d->OW);
    if (rg) {
      const int Wp = r4(d->OW);
      t.g = reinterpret_cast<const float*>(16);
      t.g_pitch = Wp; t.g_sn = (int64_t)d->OH * Wp; t.g_sb = (int64_t)d->O * d->OH * Wp;
    }
    if (s == 1) {
      const bool rx = !tma_ok_rows(x, d->W);
      if (rx) {
        const int Wp = r4(d->W);
        t.in = reinterpret_cast<const float*>(16);
        t.in_pitch = Wp; t.in_sc = (int64_t)d->H * Wp; t.in_sb = (int64_t)d->C * d->H * Wp;
      }
      if (tc_redgemm_supported(t)) {
        pl.tc = true; pl.repitch_g = rg; pl.repitch_x = rx;
        pl.g[0] = t;
      }
    } else {
      pl.H2 = (d->H + 1) / 2;
      const int W2 = (d->W + 1) / 2;
      pl.W2p = r4(W2);
      bool ok = true;
      RedGemm ph_g[4];
      for (int ph = 0; ph < 4 && ok; ++ph) {
        RedGemm q = t;
        const int py = ph >> 1, px = ph & 1;
        q.in = reinterpret_cast<const float*>(16);
        q.IH = pl.H2; q.IW = W2; q.in_pitch = pl.W2p; q.in_sc = (int64_t)pl.H2 * pl.W2p;
        q.in_sb = (int64_t)4 * d->C * pl.H2 * pl.W2p;
        q.in_sy = 1; q.in_sx = 1;
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
        pl.tc = true; pl.s2d = true; pl.repitch_g = rg;
        pl.nprob = 4;
        for (int ph = 0; ph < 4; ++ph) pl.g[ph] = ph_g[ph];
      }
    }
  }
  if (pl.tc) {
    if (pl.repitch_g) { pl.off_g = off; off += r256((size_t)d->B * d->O * d->OH * r4(d->OW) * sizeof(float)); }
    if (pl.s2d) { pl.off_x = off; off += r256((size_t)d->B * 4 * d->C * pl.H2 * pl.W2p * sizeof(float)); }
    else if (pl.repitch_x) { pl.off_x = off; off += r256((size_t)d->B * d->C * d->H * r4(d->W) * sizeof(float)); }
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
  // pointers only matter for their 16-byte alignment; assume torch's (always >= 256-byte aligned)
  // allocations, and re-check at call time.
  const float* al = reinterpret_cast<const float*>(256);
  if (which == 0) return plan_forward(d, al, al, const_cast<float*>(al), 1.f, flags).total;
  if (which == 1) return plan_dgrad(d, al, al, const_cast<float*>(al), 1.f, flags).total;
  if (which == 2) return plan_wgrad(d, al, al, const_cast<float*>(al), 1.f, flags).total;
  return 0;
}

extern "C" int msg_conv2d_forward(float* y, const float* x, const float* w, const msg_conv_desc* d, float alpha,
                                  void* workspace, size_t workspace_bytes, int flags, msg_stream_t stream) {
  int rc = check_desc(d, "conv2d_forward");
  if (rc) return rc;
  if (d->B == 0) return MSG_OK;
  if (!y || !x || !w) return fail(MSG_ERR_BAD_ARG, "conv2d_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  FwdPlan pl = plan_forward(d, x, w, y, alpha, flags);
  if (!pl.tc) {
    if (flags == MSG_CONV_FORCE_TC) return fail(MSG_ERR_UNSUPPORTED, "conv2d_forward: shape does not tile on tcgen05");
    g_last_engine = 1;
    return simt_pixgemm(pl.g, st);
  }
  if (!workspace || workspace_bytes < pl.total)
    return fail(MSG_ERR_WORKSPACE, "conv2d_forward: workspace %zu < %zu", workspace_bytes, pl.total);
  uint8_t* ws = ws_base(workspace);
  if (pl.s2d) {
    float* xs = reinterpret_cast<float*>(ws + pl.off_x);
    float* w2 = reinterpret_cast<float*>(ws + pl.off_w2);
    const int64_t tot = (int64_t)d->B * 4 * d->C * pl.H2 * pl.W2p;
    s2d_kernel<<<grid_for(tot), 256, 0, st>>>(xs, x, d->B, d->C, d->H, d->W, pl.H2, pl.W2p);
    MSG_CHECK_LAUNCH("conv s2d");
    W2Params wp{};
    wp.w = w; wp.w_sb = d->w_batch_stride; wp.BW = d->w_batch_stride ? d->B : 1; wp.N = d->O; wp.C = d->C;
    wp.kh = d->kh; wp.kw = d->kw; wp.pad_h = d->pad_h; wp.pad_w = d->pad_w;
    wp.ay0 = pl.ay0; wp.ax0 = pl.ax0; wp.nay = pl.nay; wp.nax = pl.nax;
    const int64_t wtot = (int64_t)wp.BW * d->O * 4 * d->C * pl.nay * pl.nax;
    s2d_weight_kernel<<<grid_for(wtot), 256, 0, st>>>(w2, wp);
    MSG_CHECK_LAUNCH("conv s2d weights");
    pl.g.in = xs;
    pl.g.w = w2;
  } else if (pl.repitch) {
    float* xp = reinterpret_cast<float*>(ws + pl.off_x);
    const int64_t planes = (int64_t)d->B * d->C;
    repitch_kernel<<<grid_for(planes * d->H * r4(d->W)), 256, 0, st>>>(xp, x, planes, d->H, d->W, r4(d->W));
    MSG_CHECK_LAUNCH("conv repitch");
    pl.g.in = xp;
  }
  g_last_engine = 2;
  return tc_pixgemm(pl.g, ws + pl.off_eng, r256(tc_pixgemm_workspace(pl.g)), st);
}

extern "C" int msg_conv2d_dgrad(float* dx, const float* dy, const float* w, const msg_conv_desc* d, float alpha,
                                void* workspace, size_t workspace_bytes, int flags, msg_stream_t stream) {
  int rc = check_desc(d, "conv2d_dgrad");
  if (rc) return rc;
  if (d->B == 0) return MSG_OK;
  if (!dx || !dy || !w) return fail(MSG_ERR_BAD_ARG, "conv2d_dgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  DgradPlan pl = plan_dgrad(d, dy, w, dx, alpha, flags);
  if (!pl.tc && flags == MSG_CONV_FORCE_TC)
    return fail(MSG_ERR_UNSUPPORTED, "conv2d_dgrad: shape does not tile on tcgen05");
  uint8_t* ws = nullptr;
  if (pl.tc) {
    if (!workspace || workspace_bytes < pl.total)
      return fail(MSG_ERR_WORKSPACE, "conv2d_dgrad: workspace %zu < %zu", workspace_bytes, pl.total);
    ws = ws_base(workspace);
  }
  const float* dyp = dy;
  if (pl.repitch) {
    float* p = reinterpret_cast<float*>(ws + pl.off_x);
    const int64_t planes = (int64_t)d->B * d->O;
    repitch_kernel<<<grid_for(planes * d->OH * r4(d->OW)), 256, 0, st>>>(p, dy, planes, d->OH, d->OW, r4(d->OW));
    MSG_CHECK_LAUNCH("conv repitch");
    dyp = p;
  }
  const size_t eng_each = pl.tc ? (pl.total - 256 - pl.off_eng) / pl.nph : 0;
  g_last_engine = pl.tc ? 2 : 1;
  for (int ph = 0; ph < pl.nph; ++ph) {
    PixGemm& g = pl.g[ph];
    if (g.PH <= 0 || g.PW <= 0) continue;
    if (pl.use_tc[ph]) {
      g.in = dyp;
      rc = tc_pixgemm(g, ws + pl.off_eng + ph * eng_each, eng_each, st);
    } else {
      rc = simt_pixgemm(g, st);   // also zero-fills phases that no filter tap reaches
    }
    if (rc) return rc;
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
    if (flags == MSG_CONV_FORCE_TC) return fail(MSG_ERR_UNSUPPORTED, "conv2d_wgrad: shape does not tile on tcgen05");
    g_last_engine = 1;
    return simt_redgemm(pl.g[0], st);
  }
  if (!workspace || workspace_bytes < pl.total)
    return fail(MSG_ERR_WORKSPACE, "conv2d_wgrad: workspace %zu < %zu", workspace_bytes, pl.total);
  uint8_t* ws = ws_base(workspace);
  const float* gp = dy;
  const float* xp = x;
  if (pl.repitch_g) {
    float* p = reinterpret_cast<float*>(ws + pl.off_g);
    const int64_t planes = (int64_t)d->B * d->O;
    repitch_kernel<<<grid_for(planes * d->OH * r4(d->OW)), 256, 0, st>>>(p, dy, planes, d->OH, d->OW, r4(d->OW));
    MSG_CHECK_LAUNCH("conv repitch");
    gp = p;
  }
  if (pl.s2d) {
    float* xs = reinterpret_cast<float*>(ws + pl.off_x);
    const int64_t tot = (int64_t)d->B * 4 * d->C * pl.H2 * pl.W2p;
    s2d_kernel<<<grid_for(tot), 256, 0, st>>>(xs, x, d->B, d->C, d->H, d->W, pl.H2, pl.W2p);
    MSG_CHECK_LAUNCH("conv s2d");
    xp = xs;
  } else if (pl.repitch_x) {
    float* p = reinterpret_cast<float*>(ws + pl.off_x);
    const int64_t planes = (int64_t)d->B * d->C;
    repitch_kernel<<<grid_for(planes * d->H * r4(d->W)), 256, 0, st>>>(p, x, planes, d->H, d->W, r4(d->W));
    MSG_CHECK_LAUNCH("conv repitch");
    xp = p;
  }
  g_last_engine = 2;
  for (int i = 0; i < pl.nprob; ++i) {
    RedGemm& g = pl.g[i];
    if (g.ntaps == 0) continue;
    g.g = gp;
    if (pl.s2d) g.in = xp + (int64_t)i * d->C * pl.H2 * pl.W2p;  // phase plane block
    else g.in = xp;
    rc = tc_redgemm(g, ws + pl.off_eng + i * pl.eng_each, pl.eng_each, st);
    if (rc) return rc;
  }
  return MSG_OK;
}
