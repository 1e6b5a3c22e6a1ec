// Library bookkeeping + the small bandwidth-bound kernels around the modulated convolution and ADA.
#include "common.cuh"
#include "conv_common.cuh"

namespace msg {
thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};

// ---- weight modulation + demodulation (multi_stylegan_generator.py:384-388) ------------------------
// one block per (b, o): w_mod[b,o,:,:] = scale*W[o,:,:]*s[b,:] * rsqrt(sum(.)^2 + 1e-8)
__global__ void __launch_bounds__(256)
modulate_weights_kernel(float* __restrict__ w_mod, float* __restrict__ demod_out, const float* __restrict__ W,
                        const float* __restrict__ s, int O, int C, int taps, float scale, int demodulate) {
  const int o = blockIdx.x, b = blockIdx.y;
  const int n = C * taps;
  const float* Wo = W + (int64_t)o * n;
  const float* sb = s + (int64_t)b * C;
  float* out = w_mod + ((int64_t)b * O + o) * n;
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = scale * __ldg(Wo + i) * __ldg(sb + i / taps);
    out[i] = v;
    ss = fmaf(v, v, ss);
  }
  if (!demodulate) return;
  __shared__ float scratch[32];
  __shared__ float dsh;
  ss = block_sum(ss, scratch);
  if (threadIdx.x == 0) {
    dsh = rsqrtf(ss + 1e-8f);
    if (demod_out) demod_out[(int64_t)b * O + o] = dsh;
  }
  __syncthreads();
  const float dm = dsh;
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] *= dm;
}

// ---- backward of the modulation (first order; the double backward is composed on the host) --------------------
// u = scale*W*s, d = rsqrt(sum_{c,t} u^2 + 1e-8), w_mod = u*d.  Given g = dL/dw_mod:
//   dd[b,o] = sum_{c,t} g*u ;  du = d*(g - d^2*dd*u)   (demodulated)      du = g   (not demodulated)
//   dW[o,c,t] = scale * sum_b s[b,c]*du[b,o,c,t] ;  ds[b,c] = scale * sum_{o,t} W[o,c,t]*du[b,o,c,t]
// K1: one block per (b,o): du and part[b,o,c] = sum_t W*du.   K2: dW over the batch.   K3: ds over o.
__global__ void __launch_bounds__(256)
modulate_bwd_du_kernel(float* __restrict__ du, float* __restrict__ part, const float* __restrict__ g,
                       const float* __restrict__ W, const float* __restrict__ s, const float* __restrict__ demod, int O,
                       int C, int taps, float scale, int demodulate) {
  const int o = blockIdx.x, b = blockIdx.y;
  const int n = C * taps;
  const float* Wo = W + (int64_t)o * n;
  const float* sb = s + (int64_t)b * C;
  const float* gb = g + ((int64_t)b * O + o) * n;
  float* dub = du + ((int64_t)b * O + o) * n;
  float dm = 1.f, coef = 0.f;
  if (demodulate) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
      acc = fmaf(__ldg(gb + i), scale * __ldg(Wo + i) * __ldg(sb + i / taps), acc);
    __shared__ float scratch[32];
    __shared__ float ddsh;
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) ddsh = acc;
    __syncthreads();
    dm = __ldg(demod + (int64_t)b * O + o);
    coef = dm * dm * ddsh;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float gv = __ldg(gb + i);
    float v = gv;
    if (demodulate) v = dm * (gv - coef * (scale * __ldg(Wo + i) * __ldg(sb + i / taps)));
    dub[i] = v;
  }
  __syncthreads();
  // part[b,o,c] = sum_t W[o,c,t] * du[b,o,c,t]   (du re-read from L1/L2: this block just wrote it)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < taps; ++t) acc = fmaf(__ldg(Wo + c * taps + t), dub[c * taps + t], acc);
    part[((int64_t)b * O + o) * C + c] = acc;
  }
}

__global__ void __launch_bounds__(256)
modulate_bwd_dw_kernel(float* __restrict__ dW, const float* __restrict__ du, const float* __restrict__ s, int B, int O,
                       int C, int taps, float scale) {
  const int64_t n = (int64_t)O * C * taps;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const int c = (int)((i / taps) % C);
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc = fmaf(__ldg(s + (int64_t)b * C + c), __ldg(du + (int64_t)b * n + i), acc);
    dW[i] = scale * acc;
  }
}

__global__ void __launch_bounds__(256)
modulate_bwd_ds_kernel(float* __restrict__ ds, const float* __restrict__ part, int B, int O, int C, float scale) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int o = 0; o < O; ++o) acc += __ldg(part + ((int64_t)b * O + o) * C + c);
  ds[(int64_t)b * C + c] = scale * acc;
}

// ---- noise + bias + leaky ReLU ------------------------------------------------------------------------
// grid: (chunks of HW/4, B*C)
__global__ void __launch_bounds__(256)
noise_bias_act_kernel(float* __restrict__ out, const float* __restrict__ x, const float* __restrict__ noise,
                      const float* __restrict__ noise_w, const float* __restrict__ bias, int C, int64_t HW,
                      int64_t noise_bs, float alpha, float scale) {
  const int64_t plane = blockIdx.y;
  const int b = (int)(plane / C), c = (int)(plane - (int64_t)b * C);
  const float bv = bias ? __ldg(bias + c) : 0.f;
  const float nw = noise ? __ldg(noise_w) : 0.f;
  const float* xp = x + plane * HW;
  const float* np = noise ? noise + (int64_t)b * noise_bs : nullptr;
  float* op = out + plane * HW;
  const bool vec = (HW & 3) == 0 && (noise_bs & 3) == 0 &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                     reinterpret_cast<uintptr_t>(noise)) & 15u) == 0;
  if (vec) {
    const int64_t n4 = HW >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      float4 v = __ldg(reinterpret_cast<const float4*>(xp) + i);
      float4 nz = np ? __ldg(reinterpret_cast<const float4*>(np) + i) : make_float4(0, 0, 0, 0);
      float4 o;
      float t;
      t = v.x + nw * nz.x + bv; o.x = (t > 0.f ? t : t * alpha) * scale;
      t = v.y + nw * nz.y + bv; o.y = (t > 0.f ? t : t * alpha) * scale;
      t = v.z + nw * nz.z + bv; o.z = (t > 0.f ? t : t * alpha) * scale;
      t = v.w + nw * nz.w + bv; o.w = (t > 0.f ? t : t * alpha) * scale;
      reinterpret_cast<float4*>(op)[i] = o;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (int64_t)gridDim.x * blockDim.x) {
      const float t = xp[i] + (np ? nw * np[i] : 0.f) + bv;
      op[i] = (t > 0.f ? t : t * alpha) * scale;
    }
  }
}

// ---- affine warp (grid_sample bilinear, align_corners=True) ---------------------------------------------
__device__ __forceinline__ float reflect_coord(float x, int size) {
  // torch grid_sample reflection with align_corners=True: reflect over [0, size-1]
  if (size <= 1) return 0.f;
  const float span = (float)(size - 1);
  x = fabsf(x);
  const float flips = floorf(x / span);
  const float extra = x - flips * span;
  const float r = (fmodf(flips, 2.f) == 0.f) ? extra : span - extra;
  return fminf(fmaxf(r, 0.f), span);
}

__global__ void __launch_bounds__(256)
affine_warp_kernel(float* __restrict__ out, const float* __restrict__ in, const float* __restrict__ theta, int C,
                   int H, int W, int mode) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const float* th = theta + b * 6;
  float sx = th[0] * x + th[1] * y + th[2];
  float sy = th[3] * x + th[4] * y + th[5];
  if (mode == 0) { sx = reflect_coord(sx, W); sy = reflect_coord(sy, H); }
  const float fx = floorf(sx), fy = floorf(sy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = sx - fx, wy1 = sy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
  const int64_t plane = (int64_t)H * W;
  const float* ip = in + (int64_t)b * C * plane;
  float* op = out + (int64_t)b * C * plane + (int64_t)y * W + x;
  for (int c = 0; c < C; ++c) {
    const float* p = ip + c * plane;
    float v = 0.f;
    if (vy0 && vx0) v += p[(int64_t)y0 * W + x0] * (wy0 * wx0);
    if (vy0 && vx1) v += p[(int64_t)y0 * W + x1] * (wy0 * wx1);
    if (vy1 && vx0) v += p[(int64_t)y1 * W + x0] * (wy1 * wx0);
    if (vy1 && vx1) v += p[(int64_t)y1 * W + x1] * (wy1 * wx1);
    op[c * plane] = v;
  }
}

// Adjoint of the warp (the gradient w.r.t. the warped image; theta is a constant): every output pixel scatters its
// gradient to the four source pixels with the same bilinear weights.  fp32 atomics (12.6 MB tensors: not a hot spot).
__global__ void __launch_bounds__(128)
affine_warp_bwd_kernel(float* __restrict__ dx, const float* __restrict__ g, const float* __restrict__ theta, int C,
                       int H, int W, int mode) {
  const int b = blockIdx.z;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const float* th = theta + b * 6;
  float sx = th[0] * x + th[1] * y + th[2];
  float sy = th[3] * x + th[4] * y + th[5];
  if (mode == 0) { sx = reflect_coord(sx, W); sy = reflect_coord(sy, H); }
  const float fx = floorf(sx), fy = floorf(sy);
  const int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
  const float wx1 = sx - fx, wy1 = sy - fy, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
  const bool vx0 = x0 >= 0 && x0 < W, vx1 = x1 >= 0 && x1 < W, vy0 = y0 >= 0 && y0 < H, vy1 = y1 >= 0 && y1 < H;
  const int64_t plane = (int64_t)H * W;
  float* dp = dx + (int64_t)b * C * plane;
  const float* gp = g + (int64_t)b * C * plane + (int64_t)y * W + x;
  for (int c = 0; c < C; ++c) {
    const float gv = gp[c * plane];
    float* p = dp + c * plane;
    if (vy0 && vx0) atomicAdd(p + (int64_t)y0 * W + x0, gv * (wy0 * wx0));
    if (vy0 && vx1) atomicAdd(p + (int64_t)y0 * W + x1, gv * (wy0 * wx1));
    if (vy1 && vx0) atomicAdd(p + (int64_t)y1 * W + x0, gv * (wy1 * wx0));
    if (vy1 && vx1) atomicAdd(p + (int64_t)y1 * W + x1, gv * (wy1 * wx1));
  }
}

}  // namespace msg

using namespace msg;

extern "C" int msg_abi_version(void) { return MSG_B200_ABI_VERSION; }
extern "C" const char* msg_last_error(void) { return g_last_error; }
extern "C" uint64_t msg_launch_count(void) { return g_launch_count.load(); }
extern "C" int msg_tensor_core_path_available(void) { return tc_available() ? 1 : 0; }
extern "C" const uint32_t* msg_debug_buffer(size_t* words) { return tc_debug_host(words); }
extern "C" void msg_profile_enable(int on) { tc_profile_enable(on); }
extern "C" int msg_profile_summary(msg_profile_entry* out, int max_entries) {
  if (!out || max_entries <= 0) return 0;
  return tc_profile_summary(out, max_entries);
}

extern "C" int msg_modulate_weights(float* w_mod, float* demod_out, const float* W, const float* s, int B, int O,
                                    int C, int taps, float scale, int demodulate, msg_stream_t stream) {
  if (B < 0 || O <= 0 || C <= 0 || taps <= 0) return fail(MSG_ERR_BAD_ARG, "modulate_weights: bad sizes");
  if (B == 0) return MSG_OK;
  if (!w_mod || !W || !s) return fail(MSG_ERR_BAD_ARG, "modulate_weights: null pointer");
  if (B > 65535) return fail(MSG_ERR_UNSUPPORTED, "modulate_weights: batch > 65535");
  dim3 grid((unsigned)O, (unsigned)B);
  modulate_weights_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w_mod, demod_out, W, s, O, C, taps, scale, demodulate);
  MSG_CHECK_LAUNCH("modulate_weights");
  return MSG_OK;
}

extern "C" size_t msg_modulate_weights_bwd_workspace(int B, int O, int C, int taps) {
  if (B <= 0 || O <= 0 || C <= 0 || taps <= 0) return 16;
  return ((size_t)B * O * C * taps + (size_t)B * O * C) * sizeof(float) + 256;
}

extern "C" int msg_modulate_weights_bwd(float* dW, float* ds, const float* g, const float* W, const float* s,
                                        const float* demod, int B, int O, int C, int taps, float scale, int demodulate,
                                        void* workspace, size_t workspace_bytes, msg_stream_t stream) {
  if (B < 0 || O <= 0 || C <= 0 || taps <= 0) return fail(MSG_ERR_BAD_ARG, "modulate_weights_bwd: bad sizes");
  if (!dW || !ds) return fail(MSG_ERR_BAD_ARG, "modulate_weights_bwd: null output");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    MSG_CHECK_CUDA(cudaMemsetAsync(dW, 0, (size_t)O * C * taps * sizeof(float), st));
    return MSG_OK;
  }
  if (!g || !W || !s || (demodulate && !demod)) return fail(MSG_ERR_BAD_ARG, "modulate_weights_bwd: null pointer");
  if (B > 65535) return fail(MSG_ERR_UNSUPPORTED, "modulate_weights_bwd: batch > 65535");
  const size_t need = msg_modulate_weights_bwd_workspace(B, O, C, taps);
  if (!workspace || workspace_bytes < need)
    return fail(MSG_ERR_WORKSPACE, "modulate_weights_bwd: workspace %zu < %zu", workspace_bytes, need);
  float* du = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  float* part = du + (size_t)B * O * C * taps;
  modulate_bwd_du_kernel<<<dim3((unsigned)O, (unsigned)B), 256, 0, st>>>(du, part, g, W, s, demod, O, C, taps, scale, demodulate);
  MSG_CHECK_LAUNCH("modulate_weights_bwd(du)");
  const int64_t n = (int64_t)O * C * taps;
  const int64_t want = ceil_div(n, 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  modulate_bwd_dw_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(dW, du, s, B, O, C, taps, scale);
  MSG_CHECK_LAUNCH("modulate_weights_bwd(dW)");
  modulate_bwd_ds_kernel<<<dim3((unsigned)ceil_div(C, 256), (unsigned)B), 256, 0, st>>>(ds, part, B, O, C, scale);
  MSG_CHECK_LAUNCH("modulate_weights_bwd(ds)");
  return MSG_OK;
}

extern "C" int msg_noise_bias_act(float* out, const float* x, const float* noise, const float* noise_w,
                                  const float* bias, int B, int C, int64_t HW, int64_t noise_batch_stride,
                                  float alpha, float scale, msg_stream_t stream) {
  if (B < 0 || C <= 0 || HW < 0) return fail(MSG_ERR_BAD_ARG, "noise_bias_act: bad sizes");
  if (B == 0 || HW == 0) return MSG_OK;
  if (!out || !x) return fail(MSG_ERR_BAD_ARG, "noise_bias_act: null pointer");
  if (noise && !noise_w) return fail(MSG_ERR_BAD_ARG, "noise_bias_act: noise without noise_w");
  const int64_t planes = (int64_t)B * C;
  if (planes > 65535) {
    return fail(MSG_ERR_UNSUPPORTED, "noise_bias_act: B*C > 65535");
  }
  int64_t chunks = ceil_div(HW, 256 * 4 * 4);
  if (chunks < 1) chunks = 1;
  dim3 grid((unsigned)chunks, (unsigned)planes);
  noise_bias_act_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, x, noise, noise_w, bias, C, HW,
                                                               noise_batch_stride, alpha, scale);
  MSG_CHECK_LAUNCH("noise_bias_act");
  return MSG_OK;
}

extern "C" int msg_affine_warp(float* out, const float* in, const float* theta, int B, int C, int H, int W, int mode,
                               msg_stream_t stream) {
  if (B < 0 || C <= 0 || H <= 0 || W <= 0) return fail(MSG_ERR_BAD_ARG, "affine_warp: bad sizes");
  if (mode != 0 && mode != 1) return fail(MSG_ERR_BAD_ARG, "affine_warp: mode must be 0 (reflection) or 1 (zeros)");
  if (B == 0) return MSG_OK;
  if (!out || !in || !theta) return fail(MSG_ERR_BAD_ARG, "affine_warp: null pointer");
  if (out == in) return fail(MSG_ERR_BAD_ARG, "affine_warp: in-place not supported");
  if (B > 65535 || H > 65535) return fail(MSG_ERR_UNSUPPORTED, "affine_warp: B or H > 65535");
  dim3 grid((unsigned)ceil_div(W, 128), (unsigned)H, (unsigned)B);
  affine_warp_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(out, in, theta, C, H, W, mode);
  MSG_CHECK_LAUNCH("affine_warp");
  return MSG_OK;
}

extern "C" int msg_affine_warp_bwd(float* dx, const float* g, const float* theta, int B, int C, int H, int W, int mode,
                                   msg_stream_t stream) {
  if (B < 0 || C <= 0 || H <= 0 || W <= 0) return fail(MSG_ERR_BAD_ARG, "affine_warp_bwd: bad sizes");
  if (mode != 0 && mode != 1) return fail(MSG_ERR_BAD_ARG, "affine_warp_bwd: mode must be 0 (reflection) or 1 (zeros)");
  if (B == 0) return MSG_OK;
  if (!dx || !g || !theta) return fail(MSG_ERR_BAD_ARG, "affine_warp_bwd: null pointer");
  if (dx == g) return fail(MSG_ERR_BAD_ARG, "affine_warp_bwd: in-place not supported");
  if (B > 65535 || H > 65535) return fail(MSG_ERR_UNSUPPORTED, "affine_warp_bwd: B or H > 65535");
  cudaStream_t st = (cudaStream_t)stream;
  MSG_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)B * C * H * W * sizeof(float), st));
  dim3 grid((unsigned)ceil_div(W, 128), (unsigned)H, (unsigned)B);
  affine_warp_bwd_kernel<<<grid, 128, 0, st>>>(dx, g, theta, C, H, W, mode);
  MSG_CHECK_LAUNCH("affine_warp_bwd");
  return MSG_OK;
}

// ---- deterministic reductions over channels-last activations --------------------------------------------------------------
namespace msg {

// partial[blockIdx.y][c] = sum of x[r][c] over this block's rows; block = 32 float4 columns x 8 row lanes
__global__ void __launch_bounds__(256)
colsum_partial_kernel(float4* __restrict__ partial, const float4* __restrict__ x, int64_t rows, int C4, int64_t rows_per_block) {
  __shared__ float4 sm[8][32];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C4)
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const float4 v = __ldg(x + r * C4 + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  sm[rl][cl] = acc;
  __syncthreads();
  if (rl == 0 && c < C4) {
    float4 t = sm[0][cl];
#pragma unroll
    for (int j = 1; j < 8; ++j) { t.x += sm[j][cl].x; t.y += sm[j][cl].y; t.z += sm[j][cl].z; t.w += sm[j][cl].w; }
    partial[(int64_t)blockIdx.y * C4 + c] = t;
  }
}

__global__ void __launch_bounds__(256)
colsum_final_kernel(float* __restrict__ out, const float* __restrict__ partial, int nblocks, int C, float scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int j = 0; j < nblocks; ++j) t += partial[(int64_t)j * C + c];
  out[c] = t * scale;
}

__global__ void __launch_bounds__(256)
dot_partial_kernel(float* __restrict__ partial, const float4* __restrict__ a, const float4* __restrict__ b, int64_t n4) {
  __shared__ float sm[32];
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 u = __ldg(a + i), v = __ldg(b + i);
    acc += (u.x * v.x + u.y * v.y) + (u.z * v.z + u.w * v.w);
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(256)
dot_final_kernel(float* __restrict__ out, const float* __restrict__ partial, int n, float scale) {
  __shared__ float sm[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) out[0] = acc * scale;
}

}  // namespace msg

// out[c] = scale * sum_r x[r, c] for a dense [rows, C] matrix (a channels-last activation: rows = B*H*W): the bias gradient
// of a convolution without activation (equalized_layer.py:70-73).  workspace: msg_colsum_workspace(rows, C) bytes.
extern "C" size_t msg_colsum_workspace(int64_t rows, int C) {
  (void)rows;
  return (size_t)4 * msg::num_sms() * (size_t)((C + 3) / 4 * 4) * sizeof(float) + 256;
}

extern "C" int msg_colsum_nhwc(float* out, const float* x, int64_t rows, int C, float scale, void* workspace,
                               size_t workspace_bytes, msg_stream_t stream) {
  using namespace msg;
  if (rows < 0 || C < 4 || C % 4) return fail(MSG_ERR_UNSUPPORTED, "colsum_nhwc: C must be a positive multiple of 4");
  if (!out || (rows > 0 && !x) || (reinterpret_cast<uintptr_t>(x) & 15u)) return fail(MSG_ERR_BAD_ARG, "colsum_nhwc: null or unaligned pointer");
  if (!workspace || workspace_bytes < msg_colsum_workspace(rows, C)) return fail(MSG_ERR_WORKSPACE, "colsum_nhwc: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  const int C4 = C / 4;
  const int gx = (int)ceil_div(C4, 32);
  int64_t nb = ceil_div((int64_t)4 * num_sms(), gx);
  if (nb > ceil_div(rows, 8)) nb = ceil_div(rows, 8);
  if (nb < 1) nb = 1;
  const int64_t rpb = ceil_div(rows > 0 ? rows : 1, nb);
  nb = ceil_div(rows > 0 ? rows : 1, rpb);
  colsum_partial_kernel<<<dim3((unsigned)gx, (unsigned)nb), 256, 0, st>>>(reinterpret_cast<float4*>(partial),
                                                                        reinterpret_cast<const float4*>(x), rows, C4, rpb);
  MSG_CHECK_LAUNCH("colsum_nhwc(partial)");
  colsum_final_kernel<<<(unsigned)ceil_div(C, 256), 256, 0, st>>>(out, partial, (int)nb, C, scale);
  MSG_CHECK_LAUNCH("colsum_nhwc");
  return MSG_OK;
}

// out[0] = scale * <a, b> over n floats (n % 4 == 0, 16-byte aligned); deterministic two-stage sum.  workspace: 4096 floats.
extern "C" int msg_dot(float* out, const float* a, const float* b, int64_t n, float scale, float* workspace, msg_stream_t stream) {
  using namespace msg;
  if (n < 0 || n % 4) return fail(MSG_ERR_UNSUPPORTED, "dot: n must be a multiple of 4");
  if (!out || !workspace || (n > 0 && (!a || !b)) || ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u))
    return fail(MSG_ERR_BAD_ARG, "dot: null or unaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = ceil_div(n / 4 > 0 ? n / 4 : 1, 256 * 8);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  if (blocks > 4096) blocks = 4096;
  dot_partial_kernel<<<(unsigned)blocks, 256, 0, st>>>(workspace, reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), n / 4);
  MSG_CHECK_LAUNCH("dot(partial)");
  dot_final_kernel<<<1, 256, 0, st>>>(out, workspace, (int)blocks, scale);
  MSG_CHECK_LAUNCH("dot");
  return MSG_OK;
}

// ---- MinibatchStdDev (u_net_2d_discriminator.py:189-217) ---------------------------------------------------------------
// out = cat(x, plane): plane[b] = mean over (c,h,w) of sqrt(max(var over the sub-batch of b at that position, alpha)); the
// batch is G independent sub-batches of B / G consecutive samples.  Channels-last: x [B][HW][C], out [B][HW][C + 1].
namespace msg {

struct MbStdParams {
  const float* x;
  float* out;            // forward: [B][HW][C+1]
  const float* gout;     // backward: [B][HW][C+1]
  float* gx;             // backward: [B][HW][C]
  float* partial;        // [G][nblk]
  float* stats;          // [G] plane values (forward) / sums of the plane's gradient (backward)
  int B, C, G, nblk;
  int64_t HW;
  float alpha;
};

__global__ void __launch_bounds__(256)
mbstd_partial_kernel(const MbStdParams p) {
  __shared__ float sm[32];
  const int g = blockIdx.y, Bg = p.B / p.G;
  const int64_t n = p.HW * p.C;
  const float* xg = p.x + (int64_t)g * Bg * n;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float mean = 0.f;
    for (int m = 0; m < Bg; ++m) mean += __ldg(xg + (int64_t)m * n + i);
    mean /= (float)Bg;
    float var = 0.f;
    for (int m = 0; m < Bg; ++m) { const float d = __ldg(xg + (int64_t)m * n + i) - mean; var = fmaf(d, d, var); }
    acc += sqrtf(fmaxf(var / (float)Bg, p.alpha));
  }
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) p.partial[g * p.nblk + blockIdx.x] = acc;
}

// stats[g] = scale * sum of the group's partials (fixed order)
__global__ void mbstd_final_kernel(float* __restrict__ stats, const float* __restrict__ partial, int nblk, float scale) {
  const int g = blockIdx.x;
  __shared__ float sm[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) acc += partial[g * nblk + i];
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) stats[g] = acc * scale;
}

__global__ void __launch_bounds__(256)
mbstd_write_kernel(const MbStdParams p) {
  const int C1 = p.C + 1, Bg = p.B / p.G;
  const int64_t total = (int64_t)p.B * p.HW * C1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C1);
    const int64_t row = i / C1;                       // b * HW + pos
    p.out[i] = c < p.C ? __ldg(p.x + row * p.C + c) : __ldg(p.stats + (int)(row / p.HW) / Bg);
  }
}

// partial[g][blk] = sum over the group's (b, pos) of gout[b][pos][C]
__global__ void __launch_bounds__(256)
mbstd_gplane_partial_kernel(const MbStdParams p) {
  __shared__ float sm[32];
  const int g = blockIdx.y, Bg = p.B / p.G, C1 = p.C + 1;
  const int64_t n = (int64_t)Bg * p.HW;
  const float* gg = p.gout + (int64_t)g * n * C1 + p.C;
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) acc += __ldg(gg + i * C1);
  acc = block_sum(acc, sm);
  if (threadIdx.x == 0) p.partial[g * p.nblk + blockIdx.x] = acc;
}

// gx[b][pos][c] = gout[b][pos][c] + stats[g] / (HW * C) * (x - mean) / (Bg * std_pos)   (0 where the variance was clamped)
__global__ void __launch_bounds__(256)
mbstd_backward_kernel(const MbStdParams p) {
  const int g = blockIdx.y, Bg = p.B / p.G, C1 = p.C + 1;
  const int64_t n = p.HW * p.C;
  const float* xg = p.x + (int64_t)g * Bg * n;
  const float gs = __ldg(p.stats + g) / ((float)n * (float)Bg);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float mean = 0.f;
    for (int m = 0; m < Bg; ++m) mean += __ldg(xg + (int64_t)m * n + i);
    mean /= (float)Bg;
    float var = 0.f;
    for (int m = 0; m < Bg; ++m) { const float d = __ldg(xg + (int64_t)m * n + i) - mean; var = fmaf(d, d, var); }
    var /= (float)Bg;
    const float coef = var > p.alpha ? gs * rsqrtf(var) : 0.f;
    const int64_t pos = i / p.C;
    const int c = (int)(i - pos * p.C);
    for (int m = 0; m < Bg; ++m) {
      const int64_t b = (int64_t)g * Bg + m;
      p.gx[b * n + i] = __ldg(p.gout + (b * p.HW + pos) * C1 + c) + coef * (__ldg(xg + (int64_t)m * n + i) - mean);
    }
  }
}

static int mbstd_check(const char* what, int B, int C, int64_t HW, int G) {
  if (B < 1 || C < 1 || HW < 1 || G < 1 || B % G != 0) return fail(MSG_ERR_BAD_ARG, "%s: batch %d must be a positive multiple of the group count %d", what, B, G);
  if (G > 65535) return fail(MSG_ERR_UNSUPPORTED, "%s: too many groups", what);
  return MSG_OK;
}
static int mbstd_blocks(int64_t n) {
  int64_t b = ceil_div(n, 256 * 4);
  const int64_t cap = 2 * num_sms();
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace msg

// workspace (both directions): (groups * 2 * #SMs + groups) floats
extern "C" size_t msg_mbstd_workspace(int groups) { return ((size_t)groups * 2 * msg::num_sms() + groups) * sizeof(float) + 256; }

extern "C" int msg_mbstd_forward(float* out, const float* x, int B, int C, int64_t HW, int groups, float alpha, void* workspace,
                                 size_t workspace_bytes, msg_stream_t stream) {
  using namespace msg;
  int rc = mbstd_check("mbstd_forward", B, C, HW, groups);
  if (rc) return rc;
  if (!out || !x) return fail(MSG_ERR_BAD_ARG, "mbstd_forward: null pointer");
  if (!workspace || workspace_bytes < msg_mbstd_workspace(groups)) return fail(MSG_ERR_WORKSPACE, "mbstd_forward: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  MbStdParams p{};
  p.x = x; p.out = out; p.B = B; p.C = C; p.G = groups; p.HW = HW; p.alpha = alpha;
  p.partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  p.nblk = mbstd_blocks(HW * C);
  p.stats = p.partial + (size_t)groups * 2 * num_sms();
  mbstd_partial_kernel<<<dim3((unsigned)p.nblk, (unsigned)groups), 256, 0, st>>>(p);
  MSG_CHECK_LAUNCH("mbstd_forward(partial)");
  mbstd_final_kernel<<<(unsigned)groups, 256, 0, st>>>(p.stats, p.partial, p.nblk, 1.f / ((float)HW * (float)C));
  MSG_CHECK_LAUNCH("mbstd_forward(final)");
  const int64_t total = (int64_t)B * HW * (C + 1);
  const int64_t want = ceil_div(total, 256 * 4), cap = (int64_t)16 * num_sms();
  mbstd_write_kernel<<<(unsigned)(want < cap ? (want > 0 ? want : 1) : cap), 256, 0, st>>>(p);
  MSG_CHECK_LAUNCH("mbstd_forward(write)");
  return MSG_OK;
}

extern "C" int msg_mbstd_backward(float* gx, const float* gout, const float* x, int B, int C, int64_t HW, int groups, float alpha,
                                  void* workspace, size_t workspace_bytes, msg_stream_t stream) {
  using namespace msg;
  int rc = mbstd_check("mbstd_backward", B, C, HW, groups);
  if (rc) return rc;
  if (!gx || !gout || !x) return fail(MSG_ERR_BAD_ARG, "mbstd_backward: null pointer");
  if (!workspace || workspace_bytes < msg_mbstd_workspace(groups)) return fail(MSG_ERR_WORKSPACE, "mbstd_backward: workspace");
  cudaStream_t st = (cudaStream_t)stream;
  MbStdParams p{};
  p.x = x; p.gout = gout; p.gx = gx; p.B = B; p.C = C; p.G = groups; p.HW = HW; p.alpha = alpha;
  p.partial = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
  p.stats = p.partial + (size_t)groups * 2 * num_sms();
  p.nblk = mbstd_blocks((int64_t)(B / groups) * HW);
  mbstd_gplane_partial_kernel<<<dim3((unsigned)p.nblk, (unsigned)groups), 256, 0, st>>>(p);
  MSG_CHECK_LAUNCH("mbstd_backward(plane gradient)");
  mbstd_final_kernel<<<(unsigned)groups, 256, 0, st>>>(p.stats, p.partial, p.nblk, 1.f);
  MSG_CHECK_LAUNCH("mbstd_backward(final)");
  const int nb = mbstd_blocks(HW * C);
  mbstd_backward_kernel<<<dim3((unsigned)nb, (unsigned)groups), 256, 0, st>>>(p);
  MSG_CHECK_LAUNCH("mbstd_backward");
  return MSG_OK;
}
