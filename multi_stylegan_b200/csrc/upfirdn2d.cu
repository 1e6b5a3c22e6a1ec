// upfirdn2d for sm_100a — zero-insert upsample, pad/crop, 2-D FIR (true convolution), decimate.
//
// Reference semantics: multi_stylegan/op_static/upfirdn2d_kernel.cu:52-137 (index math),
// :140-272 (launcher, out size :167-168, flipped taps :77) and op_static/upfirdn2d.py:156-190
// (`upfirdn2d_native`, the authors' own restatement).  Written from scratch:
//   * tiled kernel (minor == 1, fp32): shared-memory input tile with zero-filled halo, every thread
//     produces 4 horizontally adjacent outputs from registers (taps in registers for up == 1) and
//     writes them with one 128-bit store; 1-D grid over (plane, tile) so B*C is unbounded;
//   * generic kernel: one output per thread, any up/down/pad/kernel size/minor, fp32 or fp64.
//     The reference launches nothing outside its six hard-coded modes; this never does that.
#include "common.cuh"

namespace msg {

__host__ __device__ __forceinline__ int floor_div_i(int a, int b) {
  int q = a / b;
  if (q * b > a) --q;
  return q;
}

struct FirParams {
  int in_h, in_w, out_h, out_w;
  int pad_x0, pad_y0;
  int kernel_h, kernel_w;
  int tile_oh, tile_ow;      // outputs per tile
  int tile_ih, tile_iw;      // input rows/cols staged per tile
  int tile_iw_pad;           // smem row pitch
  int tiles_x, tiles_y;
};

// ---- tiled fast path ------------------------------------------------------------------------------
template <int UP, int DOWN, int KH, int KW>
__global__ void __launch_bounds__(256)
fir_tiled_kernel(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ kernel,
                 const FirParams p) {
  extern __shared__ float smem[];
  float* sx = smem;                         // [tile_ih][tile_iw_pad]
  __shared__ float sk[KH][KW];              // flipped taps, zero padded to KH x KW

  const int tiles = p.tiles_x * p.tiles_y;
  const int64_t plane = blockIdx.x / tiles;
  const int tile = blockIdx.x - (int)(plane * tiles);
  const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
  const int tile_out_y = ty * p.tile_oh, tile_out_x = tx * p.tile_ow;

  for (int t = threadIdx.x; t < KH * KW; t += 256) {
    const int ky = t / KW, kx = t - ky * KW;
    float v = 0.f;
    if (ky < p.kernel_h && kx < p.kernel_w)
      v = kernel[(p.kernel_h - 1 - ky) * p.kernel_w + (p.kernel_w - 1 - kx)];
    sk[ky][kx] = v;
  }

  const int tile_mid_x = tile_out_x * DOWN + UP - 1 - p.pad_x0;
  const int tile_mid_y = tile_out_y * DOWN + UP - 1 - p.pad_y0;
  const int tile_in_x = floor_div_i(tile_mid_x, UP);
  const int tile_in_y = floor_div_i(tile_mid_y, UP);
  const int res_x = tile_mid_x - tile_in_x * UP;   // in [0, UP)
  const int res_y = tile_mid_y - tile_in_y * UP;

  const float* inp = in + plane * ((int64_t)p.in_h * p.in_w);
  for (int t = threadIdx.x; t < p.tile_ih * p.tile_iw; t += 256) {
    const int ry = t / p.tile_iw, rx = t - ry * p.tile_iw;
    const int iy = ry + tile_in_y, ix = rx + tile_in_x;
    float v = 0.f;
    if (ix >= 0 && iy >= 0 && ix < p.in_w && iy < p.in_h) v = __ldg(inp + (int64_t)iy * p.in_w + ix);
    sx[ry * p.tile_iw_pad + rx] = v;
  }
  __syncthreads();

  float kreg[KH][KW];
  if (UP == 1) {
#pragma unroll
    for (int y = 0; y < KH; ++y)
#pragma unroll
      for (int x = 0; x < KW; ++x) kreg[y][x] = sk[y][x];
  }

  float* outp = out + plane * ((int64_t)p.out_h * p.out_w);
  const bool vec_ok = (p.out_w & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
  const int quads_x = p.tile_ow >> 2;
  const int items = quads_x * p.tile_oh;
  for (int it = threadIdx.x; it < items; it += 256) {
    const int ry = it / quads_x;
    const int rx = (it - ry * quads_x) << 2;
    const int oy = tile_out_y + ry, ox = tile_out_x + rx;
    if (oy >= p.out_h || ox >= p.out_w) continue;

    const int mid_y = res_y + ry * DOWN;
    const int rin_y = mid_y / UP;
    const int ph_y = (rin_y + 1) * UP - mid_y - 1;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (UP == 1) {
      // contiguous window: rows rin_y..rin_y+KH-1, cols rx*DOWN + res_x .. + 3*DOWN + KW - 1
      constexpr int WIN = 3 * DOWN + KW;
      const int cx = res_x + rx * DOWN;
#pragma unroll
      for (int y = 0; y < KH; ++y) {
        const float* row = sx + (rin_y + y) * p.tile_iw_pad + cx;
        float win[WIN];
#pragma unroll
        for (int i = 0; i < WIN; ++i) win[i] = row[i];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int x = 0; x < KW; ++x) acc[j] = fmaf(win[j * DOWN + x], kreg[y][x], acc[j]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int mid_x = res_x + (rx + j) * DOWN;
        const int rin_x = mid_x / UP;
        const int ph_x = (rin_x + 1) * UP - mid_x - 1;
#pragma unroll
        for (int y = 0; y < (KH + UP - 1) / UP; ++y) {
          const int ky = ph_y + y * UP;
          if (ky < KH) {
#pragma unroll
            for (int x = 0; x < (KW + UP - 1) / UP; ++x) {
              const int kx = ph_x + x * UP;
              if (kx < KW)
                acc[j] = fmaf(sx[(rin_y + y) * p.tile_iw_pad + rin_x + x], sk[ky][kx], acc[j]);
            }
          }
        }
      }
    }
    float* o = outp + (int64_t)oy * p.out_w + ox;
    if (vec_ok && ox + 3 < p.out_w) {
      *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (ox + j < p.out_w) o[j] = acc[j];
    }
  }
}


// ---- channels-last paths (minor = C, C % 4 == 0, fp32): one float4 = 4 channels of one pixel -------------
struct FirClParams {
  int in_h, in_w, out_h, out_w, C4;
  int pad_x0, pad_y0;
  int kernel_h, kernel_w;
  int tiles_x, tiles_y, chunks;   // blur kernel: CTA grid decomposition
  int tile_rows;
  int64_t total4;                 // gather kernel: number of output float4
  // optional fused tail of the blur kernel (StyledConv2d after an up-conv): + noise_w * noise[b, oy, ox] + bias[c],
  // leaky ReLU, gain  (multi_stylegan_generator.py:289-292, op_static/fused_act.py:58)
  const float* ep_noise;
  const float* ep_noise_w;
  const float* ep_bias;
  int64_t ep_noise_bs;
  int ep_act;
  float ep_slope, ep_gain;
  // shared-weight modulated up-convolution: v = cscale[b, c] * fir(x) + noise + bias (demodulation commutes with the
  // per-channel FIR), second output out2 = out * out2_scale[b, c] (the next layer's modulated input)
  const float* ep_cscale;
  int64_t ep_cscale_bs;
  float4* ep_out2;
  const float* ep_out2_scale;
  int64_t ep_out2_scale_bs;
};

__device__ __forceinline__ void fma4(float4& a, const float4& v, float k) {
  a.x = fmaf(v.x, k, a.x); a.y = fmaf(v.y, k, a.y); a.z = fmaf(v.z, k, a.z); a.w = fmaf(v.w, k, a.w);
}

// up == down == 1, taps <= 4x4.  A warp spans 32 channel quads (512 contiguous bytes) of one output column pair;
// each thread slides down the rows of its tile keeping the four partially accumulated output rows in registers,
// so every input row is loaded once per thread (COLS + 3 float4 loads per COLS outputs; neighbours hit in L1).
// EP: 0 = plain FIR (the upfirdn2d op), 1 = fused noise / bias / leaky ReLU tail, 2 = the tail of the shared-weight
// modulated up-convolution (per-sample channel scale before the tail, second output).  Separate instantiations: the
// plain blur runs at 64 registers (4 CTAs per SM) only without the tail's operands live across the row loop.
template <int COLS, int EP>
__global__ void __launch_bounds__(256, EP == 0 ? 4 : 3)
fir_cl_blur_kernel(float4* __restrict__ out, const float4* __restrict__ in, const float* __restrict__ kernel,
                   const FirClParams p) {
  __shared__ float sk[4][4];
  if (threadIdx.x < 16) {
    const int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    float v = 0.f;
    if (ky < p.kernel_h && kx < p.kernel_w) v = kernel[(p.kernel_h - 1 - ky) * p.kernel_w + (p.kernel_w - 1 - kx)];
    sk[ky][kx] = v;
  }
  __syncthreads();
  float kf[4][4];
#pragma unroll
  for (int y = 0; y < 4; ++y)
#pragma unroll
    for (int x = 0; x < 4; ++x) kf[y][x] = sk[y][x];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t bid = blockIdx.x;
  const int chunk = (int)(bid % p.chunks); bid /= p.chunks;
  const int tx = (int)(bid % p.tiles_x); bid /= p.tiles_x;
  const int ty = (int)(bid % p.tiles_y); bid /= p.tiles_y;
  const int64_t b = bid;
  const int q = chunk * 32 + lane;
  if (q >= p.C4) return;
  const int ox0 = (tx * 8 + warp) * COLS;
  if (ox0 >= p.out_w) return;
  const int oy0 = ty * p.tile_rows;
  const int ix0 = ox0 - p.pad_x0;
  const float4* inb = in + (b * p.in_h * (int64_t)p.in_w) * p.C4 + q;
  float4* outb = out + (b * p.out_h * (int64_t)p.out_w) * p.C4 + q;

  float4 acc[4][COLS];
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int c = 0; c < COLS; ++c) acc[s][c] = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr bool fused = EP > 0;
  float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 cs = make_float4(1.f, 1.f, 1.f, 1.f);
  float4 s2 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4* out2b = nullptr;
  float nw = 0.f;
  const float* nzb = nullptr;
  if (EP > 0) {
    if (p.ep_bias) bz = __ldg(reinterpret_cast<const float4*>(p.ep_bias) + q);
    nw = p.ep_noise ? __ldg(p.ep_noise_w) : 0.f;
    nzb = p.ep_noise ? p.ep_noise + b * p.ep_noise_bs : nullptr;
  }
  if (EP > 1) {
    if (p.ep_cscale) cs = __ldg(reinterpret_cast<const float4*>(p.ep_cscale + b * p.ep_cscale_bs) + q);
    if (p.ep_out2) {
      s2 = __ldg(reinterpret_cast<const float4*>(p.ep_out2_scale + b * p.ep_out2_scale_bs) + q);
      out2b = p.ep_out2 + (b * p.out_h * (int64_t)p.out_w) * p.C4 + q;
    }
  }

  const int n_in = p.tile_rows + 3;               // input rows i = 0 .. tile_rows + 2  (iy = oy0 - pad_y0 + i)
  for (int i0 = 0; i0 < n_in; i0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u;
      const int iy = oy0 - p.pad_y0 + i;
      if (iy >= 0 && iy < p.in_h && i < n_in) {
        float4 v[COLS + 3];
        const float4* row = inb + ((int64_t)iy * p.in_w) * p.C4;
#pragma unroll
        for (int j = 0; j < COLS + 3; ++j) {
          const int ix = ix0 + j;
          v[j] = (ix >= 0 && ix < p.in_w) ? __ldg(row + (int64_t)ix * p.C4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int ky = 0; ky < 4; ++ky) {
          const int s = (u - ky) & 3;               // slot of output row t = i - ky
#pragma unroll
          for (int c = 0; c < COLS; ++c)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) fma4(acc[s][c], v[c + kx], kf[ky][kx]);
        }
      }
      // output row t = i - 3 is complete
      {
        const int s = (u - 3) & 3;
        const int t = i - 3;
        const int oy = oy0 + t;
        if (t >= 0 && t < p.tile_rows && oy < p.out_h) {
          float4* orow = outb + ((int64_t)oy * p.out_w) * p.C4;
#pragma unroll
          for (int c = 0; c < COLS; ++c)
            if (ox0 + c < p.out_w) {
              float4 v = acc[s][c];
              if (fused) {
                const float add = nzb ? nw * __ldg(nzb + (int64_t)oy * p.out_w + ox0 + c) : 0.f;
                if (EP > 1) {
                  v.x = fmaf(v.x, cs.x, add + bz.x); v.y = fmaf(v.y, cs.y, add + bz.y);
                  v.z = fmaf(v.z, cs.z, add + bz.z); v.w = fmaf(v.w, cs.w, add + bz.w);
                } else {
                  v.x += add + bz.x; v.y += add + bz.y; v.z += add + bz.z; v.w += add + bz.w;
                }
                if (p.ep_act) {
                  v.x = v.x > 0.f ? v.x : v.x * p.ep_slope; v.y = v.y > 0.f ? v.y : v.y * p.ep_slope;
                  v.z = v.z > 0.f ? v.z : v.z * p.ep_slope; v.w = v.w > 0.f ? v.w : v.w * p.ep_slope;
                }
                v.x *= p.ep_gain; v.y *= p.ep_gain; v.z *= p.ep_gain; v.w *= p.ep_gain;
              }
              orow[(int64_t)(ox0 + c) * p.C4] = v;
              if (EP > 1) {
                if (out2b)
                  out2b[((int64_t)oy * p.out_w + ox0 + c) * p.C4] = make_float4(v.x * s2.x, v.y * s2.y, v.z * s2.z, v.w * s2.w);
              }
            }
        }
#pragma unroll
        for (int c = 0; c < COLS; ++c) acc[s][c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
}

// up == 2, down == 1, taps <= 4x4 (polyphase: every output reads 2x2 inputs).  The two outputs per axis that share an
// input pair are produced together: a thread loads one 2x2 input patch (4 float4) and writes a 2x2 output block.
__global__ void __launch_bounds__(256)
fir_cl_up2_kernel(float4* __restrict__ out, const float4* __restrict__ in, const float* __restrict__ kernel,
                  const FirClParams p, int mx_min, int my_min, int nmx, int nmy) {
  __shared__ float sk[4][4];
  if (threadIdx.x < 16) {
    const int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    float v = 0.f;
    if (ky < p.kernel_h && kx < p.kernel_w) v = kernel[(p.kernel_h - 1 - ky) * p.kernel_w + (p.kernel_w - 1 - kx)];
    sk[ky][kx] = v;
  }
  __syncthreads();
  float kf[4][4];
#pragma unroll
  for (int y = 0; y < 4; ++y)
#pragma unroll
    for (int x = 0; x < 4; ++x) kf[y][x] = sk[y][x];
  const int64_t total = (int64_t)p.total4;            // here: number of (b, jy, jx, q) work items
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; idx < total; idx += stride) {
    int64_t t = idx;
    const int q = (int)(t % p.C4); t /= p.C4;
    const int jx = (int)(t % nmx); t /= nmx;
    const int jy = (int)(t % nmy); t /= nmy;
    const int64_t b = t;
    const int mx = mx_min + jx, my = my_min + jy;     // input coordinates of the patch's top-left element
    const float4* inb = in + (b * p.in_h * (int64_t)p.in_w) * p.C4 + q;
    float4 I[2][2];
#pragma unroll
    for (int y = 0; y < 2; ++y)
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int iy = my + y, ix = mx + x;
        I[y][x] = (iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w) ? __ldg(inb + ((int64_t)iy * p.in_w + ix) * p.C4)
                                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    float4* outb = out + (b * p.out_h * (int64_t)p.out_w) * p.C4 + q;
    // outputs o = 2m - 1 + pad0 (tap phase 1) and o = 2m + pad0 (tap phase 0)
#pragma unroll
    for (int ay = 0; ay < 2; ++ay) {
      const int oy = 2 * my - 1 + p.pad_y0 + ay;
      if (oy < 0 || oy >= p.out_h) continue;
      const int phy = 1 - ay;
#pragma unroll
      for (int ax = 0; ax < 2; ++ax) {
        const int ox = 2 * mx - 1 + p.pad_x0 + ax;
        if (ox < 0 || ox >= p.out_w) continue;
        const int phx = 1 - ax;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int y = 0; y < 2; ++y)
#pragma unroll
          for (int x = 0; x < 2; ++x) fma4(acc, I[y][x], kf[phy + 2 * y][phx + 2 * x]);
        outb[((int64_t)oy * p.out_w + ox) * p.C4] = acc;
      }
    }
  }
}

// up == 1, down == 2, taps <= 4x4.  Same thread layout as the blur kernel; a thread walks down its strip of output
// rows, loads two input rows per output row and adds them to the two output rows they touch (taps 0,1 of the current
// row, taps 2,3 of the previous one): 3 * COLS + ... float4 loads per output instead of 16.
template <int COLS>
__global__ void __launch_bounds__(256)
fir_cl_down2_kernel(float4* __restrict__ out, const float4* __restrict__ in, const float* __restrict__ kernel,
                    const FirClParams p) {
  __shared__ float sk[4][4];
  if (threadIdx.x < 16) {
    const int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    float v = 0.f;
    if (ky < p.kernel_h && kx < p.kernel_w) v = kernel[(p.kernel_h - 1 - ky) * p.kernel_w + (p.kernel_w - 1 - kx)];
    sk[ky][kx] = v;
  }
  __syncthreads();
  float kf[4][4];
#pragma unroll
  for (int y = 0; y < 4; ++y)
#pragma unroll
    for (int x = 0; x < 4; ++x) kf[y][x] = sk[y][x];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t bid = blockIdx.x;
  const int chunk = (int)(bid % p.chunks); bid /= p.chunks;
  const int tx = (int)(bid % p.tiles_x); bid /= p.tiles_x;
  const int ty = (int)(bid % p.tiles_y); bid /= p.tiles_y;
  const int64_t b = bid;
  const int q = chunk * 32 + lane;
  if (q >= p.C4) return;
  const int ox0 = (tx * 8 + warp) * COLS;
  if (ox0 >= p.out_w) return;
  const int oy0 = ty * p.tile_rows;
  const int ix0 = 2 * ox0 - p.pad_x0;
  constexpr int NIN = 2 * COLS + 2;
  const float4* inb = in + (b * p.in_h * (int64_t)p.in_w) * p.C4 + q;
  float4* outb = out + (b * p.out_h * (int64_t)p.out_w) * p.C4 + q;

  float4 prev[COLS], cur[COLS];
#pragma unroll
  for (int c = 0; c < COLS; ++c) prev[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = oy0; i <= oy0 + p.tile_rows && i <= p.out_h; ++i) {
#pragma unroll
    for (int c = 0; c < COLS; ++c) cur[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int iy = 2 * i - p.pad_y0 + rr;
      if (iy < 0 || iy >= p.in_h) continue;
      const float4* row = inb + ((int64_t)iy * p.in_w) * p.C4;
      float4 v[NIN];
#pragma unroll
      for (int j = 0; j < NIN; ++j) {
        const int ix = ix0 + j;
        v[j] = (ix >= 0 && ix < p.in_w) ? __ldg(row + (int64_t)ix * p.C4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int c = 0; c < COLS; ++c)
#pragma unroll
        for (int kx = 0; kx < 4; ++kx) {
          fma4(cur[c], v[2 * c + kx], kf[rr][kx]);          // taps 0,1 of output row i
          fma4(prev[c], v[2 * c + kx], kf[rr + 2][kx]);     // taps 2,3 of output row i - 1
        }
    }
    const int oy = i - 1;
    if (oy >= oy0 && oy < p.out_h) {
      float4* orow = outb + ((int64_t)oy * p.out_w) * p.C4;
#pragma unroll
      for (int c = 0; c < COLS; ++c)
        if (ox0 + c < p.out_w) orow[(int64_t)(ox0 + c) * p.C4] = prev[c];
    }
#pragma unroll
    for (int c = 0; c < COLS; ++c) prev[c] = cur[c];
  }
}

// any (UP, DOWN) in {1,2}, taps <= 4x4: one output float4 per thread, taps gathered through L1.
template <int UP, int DOWN>
__global__ void __launch_bounds__(256)
fir_cl_gather_kernel(float4* __restrict__ out, const float4* __restrict__ in, const float* __restrict__ kernel,
                     const FirClParams p) {
  __shared__ float sk[4][4];
  if (threadIdx.x < 16) {
    const int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
    float v = 0.f;
    if (ky < p.kernel_h && kx < p.kernel_w) v = kernel[(p.kernel_h - 1 - ky) * p.kernel_w + (p.kernel_w - 1 - kx)];
    sk[ky][kx] = v;
  }
  __syncthreads();
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; idx < p.total4; idx += stride) {
    int64_t t = idx;
    const int q = (int)(t % p.C4); t /= p.C4;
    const int ox = (int)(t % p.out_w); t /= p.out_w;
    const int oy = (int)(t % p.out_h); t /= p.out_h;
    const int64_t b = t;
    const int mid_x = ox * DOWN + UP - 1 - p.pad_x0;
    const int mid_y = oy * DOWN + UP - 1 - p.pad_y0;
    const int in_x0 = floor_div_i(mid_x, UP), in_y0 = floor_div_i(mid_y, UP);
    const int ph_x = (in_x0 + 1) * UP - mid_x - 1, ph_y = (in_y0 + 1) * UP - mid_y - 1;
    const float4* inb = in + (b * p.in_h * (int64_t)p.in_w) * p.C4 + q;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int y = 0; y < (4 + UP - 1) / UP; ++y) {
      const int ky = ph_y + y * UP, iy = in_y0 + y;
      if (ky < 4 && iy >= 0 && iy < p.in_h) {
#pragma unroll
        for (int x = 0; x < (4 + UP - 1) / UP; ++x) {
          const int kx = ph_x + x * UP, ix = in_x0 + x;
          if (kx < 4 && ix >= 0 && ix < p.in_w)
            fma4(acc, __ldg(inb + ((int64_t)iy * p.in_w + ix) * p.C4), sk[ky][kx]);
        }
      }
    }
    out[idx] = acc;
  }
}

// ---- generic path ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
fir_generic_kernel(T* __restrict__ out, const T* __restrict__ in, const T* __restrict__ kernel,
                   int64_t total, int in_h, int in_w, int minor, int out_h, int out_w, int kernel_h,
                   int kernel_w, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_y0) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; idx < total; idx += stride) {
    int64_t t = idx;
    const int mi = (int)(t % minor); t /= minor;
    const int ox = (int)(t % out_w); t /= out_w;
    const int oy = (int)(t % out_h); t /= out_h;
    const int64_t major = t;
    const int mid_x = ox * down_x + up_x - 1 - pad_x0;
    const int mid_y = oy * down_y + up_y - 1 - pad_y0;
    const int in_x0 = floor_div_i(mid_x, up_x), in_y0 = floor_div_i(mid_y, up_y);
    const int ph_x = (in_x0 + 1) * up_x - mid_x - 1, ph_y = (in_y0 + 1) * up_y - mid_y - 1;
    T acc = T(0);
    for (int ky = ph_y, iy = in_y0; ky < kernel_h; ky += up_y, ++iy) {
      if (iy < 0 || iy >= in_h) continue;
      for (int kx = ph_x, ix = in_x0; kx < kernel_w; kx += up_x, ++ix) {
        if (ix < 0 || ix >= in_w) continue;
        const T kv = kernel[(kernel_h - 1 - ky) * kernel_w + (kernel_w - 1 - kx)];
        acc += in[((major * in_h + iy) * in_w + ix) * minor + mi] * kv;
      }
    }
    out[idx] = acc;
  }
}

template <int UP, int DOWN, int KH, int KW>
static int launch_tiled(float* out, const float* in, const float* kernel, int64_t major, FirParams p,
                        cudaStream_t st) {
  // tile: 4..128 outputs wide (multiple of 4), about 4096 outputs
  int tow = ((p.out_w + 3) / 4) * 4;
  if (tow > 128) tow = 128;
  int toh = 4096 / tow;
  if (toh > p.out_h) toh = p.out_h;
  if (toh < 1) toh = 1;
  p.tile_ow = tow;
  p.tile_oh = toh;
  p.tile_ih = ((toh - 1) * DOWN + KH - 1) / UP + 1;
  p.tile_iw = ((tow - 1) * DOWN + KW - 1) / UP + 1;
  p.tile_iw_pad = p.tile_iw | 1;  // odd pitch: no systematic bank conflicts between rows
  p.tiles_x = (p.out_w + tow - 1) / tow;
  p.tiles_y = (p.out_h + toh - 1) / toh;
  const size_t smem = (size_t)p.tile_ih * p.tile_iw_pad * sizeof(float);
  const int64_t blocks = major * p.tiles_x * p.tiles_y;
  if (blocks > 0x7fffffffLL) return fail(MSG_ERR_UNSUPPORTED, "upfirdn2d: grid too large");
  auto kfn = fir_tiled_kernel<UP, DOWN, KH, KW>;
  if (smem > 48 * 1024) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  kfn<<<(unsigned)blocks, 256, smem, st>>>(out, in, kernel, p);
  MSG_CHECK_LAUNCH("upfirdn2d(tiled)");
  return MSG_OK;
}

}  // namespace msg

using namespace msg;

extern "C" int msg_upfirdn2d_out_size(int in_size, int up, int down, int pad0, int pad1, int ksize) {
  if (up <= 0 || down <= 0) return -1;
  // upfirdn2d_kernel.cu:167-168
  return (in_size * up + pad0 + pad1 - ksize + down) / down;
}

struct FirEpilogue {
  const float* noise; const float* noise_w; const float* bias; int64_t noise_bs; int act; float slope, gain;
  const float* cscale; int64_t cscale_bs; float* out2; const float* out2_scale; int64_t out2_scale_bs;
};

static int upfirdn2d_impl(void* out, const void* in, const void* kernel, int64_t major, int in_h,
                          int in_w, int minor, int kernel_h, int kernel_w, int up_x, int up_y,
                          int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                          int dtype, msg_stream_t stream, const FirEpilogue* ep);

extern "C" int msg_upfirdn2d(void* out, const void* in, const void* kernel, int64_t major, int in_h,
                             int in_w, int minor, int kernel_h, int kernel_w, int up_x, int up_y,
                             int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                             int dtype, msg_stream_t stream) {
  return upfirdn2d_impl(out, in, kernel, major, in_h, in_w, minor, kernel_h, kernel_w, up_x, up_y, down_x, down_y, pad_x0,
                        pad_x1, pad_y0, pad_y1, dtype, stream, nullptr);
}

extern "C" int msg_upfirdn2d_bias_act(float* out, const float* in, const float* kernel, int64_t major, int in_h, int in_w,
                                      int minor, int kernel_h, int kernel_w, int pad_x0, int pad_x1, int pad_y0,
                                      int pad_y1, const float* noise, const float* noise_w, int64_t noise_batch_stride,
                                      const float* bias, int act, float slope, float gain, msg_stream_t stream) {
  if (noise && !noise_w) return fail(MSG_ERR_BAD_ARG, "upfirdn2d_bias_act: noise needs noise_w");
  if (act != 0 && act != 1) return fail(MSG_ERR_BAD_ARG, "upfirdn2d_bias_act: act must be 0 or 1");
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15u)) return fail(MSG_ERR_BAD_ARG, "upfirdn2d_bias_act: bias must be 16-byte aligned");
  FirEpilogue ep{noise, noise_w, bias, noise_batch_stride, act, slope, gain, nullptr, 0, nullptr, nullptr, 0};
  return upfirdn2d_impl(out, in, kernel, major, in_h, in_w, minor, kernel_h, kernel_w, 1, 1, 1, 1, pad_x0, pad_x1, pad_y0,
                        pad_y1, MSG_F32, stream, &ep);
}

extern "C" int msg_upfirdn2d_bias_act_mod(float* out, float* out2, const float* in, const float* kernel, int64_t major,
                                          int in_h, int in_w, int minor, int kernel_h, int kernel_w, int pad_x0, int pad_x1,
                                          int pad_y0, int pad_y1, const float* col_scale, int64_t col_scale_batch_stride,
                                          const float* noise, const float* noise_w, int64_t noise_batch_stride,
                                          const float* bias, int act, float slope, float gain, const float* out2_scale,
                                          int64_t out2_scale_batch_stride, msg_stream_t stream) {
  if (noise && !noise_w) return fail(MSG_ERR_BAD_ARG, "upfirdn2d_bias_act_mod: noise needs noise_w");
  if (act != 0 && act != 1) return fail(MSG_ERR_BAD_ARG, "upfirdn2d_bias_act_mod: act must be 0 or 1");
  if (out2 && !out2_scale) return fail(MSG_ERR_BAD_ARG, "upfirdn2d_bias_act_mod: out2 needs out2_scale");
  const uintptr_t al = reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(col_scale) |
                       reinterpret_cast<uintptr_t>(out2) | reinterpret_cast<uintptr_t>(out2_scale);
  if ((al & 15u) || (col_scale_batch_stride % 4) || (out2_scale_batch_stride % 4))
    return fail(MSG_ERR_BAD_ARG, "upfirdn2d_bias_act_mod: bias / scales / out2 must be 16-byte aligned");
  FirEpilogue ep{noise, noise_w, bias, noise_batch_stride, act, slope, gain, col_scale, col_scale_batch_stride,
                 out2, out2_scale, out2_scale_batch_stride};
  return upfirdn2d_impl(out, in, kernel, major, in_h, in_w, minor, kernel_h, kernel_w, 1, 1, 1, 1, pad_x0, pad_x1, pad_y0,
                        pad_y1, MSG_F32, stream, &ep);
}

static int upfirdn2d_impl(void* out, const void* in, const void* kernel, int64_t major, int in_h,
                          int in_w, int minor, int kernel_h, int kernel_w, int up_x, int up_y,
                          int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                          int dtype, msg_stream_t stream, const FirEpilogue* ep) {
  if (major < 0 || in_h < 0 || in_w < 0 || minor < 0) return fail(MSG_ERR_BAD_ARG, "upfirdn2d: negative size");
  if (up_x < 1 || up_y < 1 || down_x < 1 || down_y < 1) return fail(MSG_ERR_BAD_ARG, "upfirdn2d: up/down must be >= 1");
  if (kernel_h < 1 || kernel_w < 1) return fail(MSG_ERR_BAD_ARG, "upfirdn2d: empty FIR kernel");
  if (kernel_h > 32 || kernel_w > 32) return fail(MSG_ERR_UNSUPPORTED, "upfirdn2d: FIR kernel larger than 32x32");
  const int out_h = msg_upfirdn2d_out_size(in_h, up_y, down_y, pad_y0, pad_y1, kernel_h);
  const int out_w = msg_upfirdn2d_out_size(in_w, up_x, down_x, pad_x0, pad_x1, kernel_w);
  if (out_h < 0 || out_w < 0) return fail(MSG_ERR_BAD_ARG, "upfirdn2d: negative output size (%d x %d)", out_h, out_w);
  const int64_t total = major * out_h * (int64_t)out_w * minor;
  if (total == 0) return MSG_OK;
  if (!out || !in || !kernel) return fail(MSG_ERR_BAD_ARG, "upfirdn2d: null pointer");
  cudaStream_t st = (cudaStream_t)stream;

  if (ep) {
    // the fused tail lives in the channels-last blur kernel only
    const bool ok = dtype == MSG_F32 && minor >= 4 && (minor % 4 == 0) && up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 &&
                    kernel_h <= 4 && kernel_w <= 4 && in_h > 0 && in_w > 0 &&
                    ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    if (!ok) return fail(MSG_ERR_UNSUPPORTED, "upfirdn2d_bias_act: needs fp32 channels-last data (minor %% 4 == 0), up = down = 1, taps <= 4x4");
  }
  if (dtype == MSG_F32 && minor == 1 && up_x == up_y && down_x == down_y && kernel_h <= 4 && kernel_w <= 4 &&
      in_h > 0 && in_w > 0) {
    FirParams p{};
    p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
    p.pad_x0 = pad_x0; p.pad_y0 = pad_y0; p.kernel_h = kernel_h; p.kernel_w = kernel_w;
    if (up_x == 1 && down_x == 1) return launch_tiled<1, 1, 4, 4>((float*)out, (const float*)in, (const float*)kernel, major, p, st);
    if (up_x == 2 && down_x == 1) return launch_tiled<2, 1, 4, 4>((float*)out, (const float*)in, (const float*)kernel, major, p, st);
    if (up_x == 1 && down_x == 2) return launch_tiled<1, 2, 4, 4>((float*)out, (const float*)in, (const float*)kernel, major, p, st);
  }
  if (dtype == MSG_F32 && minor >= 4 && (minor % 4 == 0) && up_x == up_y && down_x == down_y && up_x <= 2 && down_x <= 2 &&
      kernel_h <= 4 && kernel_w <= 4 && in_h > 0 && in_w > 0 &&
      ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
    FirClParams p{};
    p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w; p.C4 = minor / 4;
    p.pad_x0 = pad_x0; p.pad_y0 = pad_y0; p.kernel_h = kernel_h; p.kernel_w = kernel_w;
    p.total4 = total / 4;
    p.ep_gain = 1.f;
    if (ep) {
      p.ep_noise = ep->noise; p.ep_noise_w = ep->noise_w; p.ep_bias = ep->bias; p.ep_noise_bs = ep->noise_bs;
      p.ep_act = ep->act; p.ep_slope = ep->slope; p.ep_gain = ep->gain;
      p.ep_cscale = ep->cscale; p.ep_cscale_bs = ep->cscale_bs;
      p.ep_out2 = reinterpret_cast<float4*>(ep->out2); p.ep_out2_scale = ep->out2_scale; p.ep_out2_scale_bs = ep->out2_scale_bs;
    }
    if (up_x == 1 && down_x == 1) {
      constexpr int COLS = 2;
      p.chunks = (int)ceil_div(p.C4, 32);
      p.tiles_x = (int)ceil_div(out_w, 8 * COLS);
      // strips of 32 output rows (3 halo rows each); shorter strips when that leaves too few CTAs to fill the machine
      p.tile_rows = out_h < 32 ? out_h : 32;
      while (p.tile_rows > 8 && major * ceil_div(out_h, p.tile_rows) * p.tiles_x * p.chunks < 8 * (int64_t)num_sms())
        p.tile_rows >>= 1;
      p.tiles_y = (int)ceil_div(out_h, p.tile_rows);
      const int64_t blocks = major * p.tiles_y * p.tiles_x * p.chunks;
      if (blocks > 0x7fffffffLL) return fail(MSG_ERR_UNSUPPORTED, "upfirdn2d: grid too large");
      const bool tail = p.ep_act || p.ep_bias || p.ep_noise, mod = p.ep_cscale || p.ep_out2;
      if (mod) fir_cl_blur_kernel<COLS, 2><<<(unsigned)blocks, 256, 0, st>>>((float4*)out, (const float4*)in, (const float*)kernel, p);
      else if (tail || p.ep_gain != 1.f) fir_cl_blur_kernel<COLS, 1><<<(unsigned)blocks, 256, 0, st>>>((float4*)out, (const float4*)in, (const float*)kernel, p);
      else fir_cl_blur_kernel<COLS, 0><<<(unsigned)blocks, 256, 0, st>>>((float4*)out, (const float4*)in, (const float*)kernel, p);
      MSG_CHECK_LAUNCH("upfirdn2d(channels-last blur)");
      return MSG_OK;
    }
    const int64_t want4 = ceil_div(p.total4, 256);
    const int64_t cap4 = (int64_t)num_sms() * 64;
    const unsigned g4 = (unsigned)(want4 < cap4 ? want4 : cap4);
    if (up_x == 2 && down_x == 1) {
      // input patches m with outputs {2m - 1 + pad0, 2m + pad0} intersecting [0, out)
      const int mx_min = floor_div_i(1 - pad_x0, 2), mx_max = floor_div_i(out_w - pad_x0, 2);
      const int my_min = floor_div_i(1 - pad_y0, 2), my_max = floor_div_i(out_h - pad_y0, 2);
      const int nmx = mx_max - mx_min + 1, nmy = my_max - my_min + 1;
      FirClParams pu = p;
      pu.total4 = major * nmy * (int64_t)nmx * p.C4;
      const int64_t wantu = ceil_div(pu.total4, 256);
      const unsigned gu = (unsigned)(wantu < cap4 ? wantu : cap4);
      fir_cl_up2_kernel<<<gu, 256, 0, st>>>((float4*)out, (const float4*)in, (const float*)kernel, pu, mx_min, my_min, nmx, nmy);
      MSG_CHECK_LAUNCH("upfirdn2d(channels-last up2)");
      return MSG_OK;
    }
    if (up_x == 1 && down_x == 2) {
      constexpr int COLS = 2;
      FirClParams pd = p;
      pd.chunks = (int)ceil_div(p.C4, 32);
      pd.tiles_x = (int)ceil_div(out_w, 8 * COLS);
      pd.tile_rows = out_h < 32 ? out_h : 32;
      while (pd.tile_rows > 8 && major * ceil_div(out_h, pd.tile_rows) * pd.tiles_x * pd.chunks < 8 * (int64_t)num_sms())
        pd.tile_rows >>= 1;
      pd.tiles_y = (int)ceil_div(out_h, pd.tile_rows);
      const int64_t blocks = major * pd.tiles_y * pd.tiles_x * pd.chunks;
      if (blocks > 0x7fffffffLL) return fail(MSG_ERR_UNSUPPORTED, "upfirdn2d: grid too large");
      fir_cl_down2_kernel<COLS><<<(unsigned)blocks, 256, 0, st>>>((float4*)out, (const float4*)in, (const float*)kernel, pd);
      MSG_CHECK_LAUNCH("upfirdn2d(channels-last down2)");
      return MSG_OK;
    }
    if (up_x == 2 && down_x == 1)
      fir_cl_gather_kernel<2, 1><<<g4, 256, 0, st>>>((float4*)out, (const float4*)in, (const float*)kernel, p);
    else if (up_x == 1 && down_x == 2)
      fir_cl_gather_kernel<1, 2><<<g4, 256, 0, st>>>((float4*)out, (const float4*)in, (const float*)kernel, p);
    else
      fir_cl_gather_kernel<2, 2><<<g4, 256, 0, st>>>((float4*)out, (const float4*)in, (const float*)kernel, p);
    MSG_CHECK_LAUNCH("upfirdn2d(channels-last gather)");
    return MSG_OK;
  }
  const int64_t want = ceil_div(total, 256);
  const int64_t cap = (int64_t)num_sms() * 32;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
  if (dtype == MSG_F32) {
    fir_generic_kernel<float><<<grid, 256, 0, st>>>((float*)out, (const float*)in, (const float*)kernel, total,
                                                    in_h, in_w, minor, out_h, out_w, kernel_h, kernel_w, up_x, up_y,
                                                    down_x, down_y, pad_x0, pad_y0);
  } else if (dtype == MSG_F64) {
    fir_generic_kernel<double><<<grid, 256, 0, st>>>((double*)out, (const double*)in, (const double*)kernel, total,
                                                     in_h, in_w, minor, out_h, out_w, kernel_h, kernel_w, up_x, up_y,
                                                     down_x, down_y, pad_x0, pad_y0);
  } else {
    return fail(MSG_ERR_UNSUPPORTED, "upfirdn2d: dtype %d", dtype);
  }
  MSG_CHECK_LAUNCH("upfirdn2d(generic)");
  return MSG_OK;
}
