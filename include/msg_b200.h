/*
 * msg_b200.h — C-ABI of the B200-native Multi-StyleGAN training-step hot path.
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream; none allocates,
 * none synchronises, none touches a torch type.  All functions return MSG_OK (0) or a
 * negative-free error code below; msg_last_error() gives the thread-local reason string.
 * All tensors are dense row-major ("contiguous") unless stated otherwise.
 *
 * Each declaration cites the reference interface (file:line under
 * ChristophReich1996/Multi-StyleGAN) whose arithmetic it replaces.
 */
#ifndef MSG_B200_H_
#define MSG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSG_B200_ABI_VERSION 3

/* ---- status codes ------------------------------------------------------------------------- */
enum {
  MSG_OK = 0,
  MSG_ERR_BAD_ARG = 1,      /* null pointer, negative size, inconsistent shapes            */
  MSG_ERR_UNSUPPORTED = 2,  /* configuration outside what the kernels implement            */
  MSG_ERR_CUDA = 3,         /* a CUDA runtime / driver call or a kernel launch failed      */
  MSG_ERR_WORKSPACE = 4     /* caller-provided workspace too small                         */
};

/* ---- element types accepted by the bandwidth-bound ops ------------------------------------ */
enum { MSG_F32 = 0, MSG_F64 = 1 };

/* opaque cudaStream_t */
typedef void* msg_stream_t;

int msg_abi_version(void);
/* thread-local, valid until the next failing call on the same thread */
const char* msg_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t msg_launch_count(void);
/* 1 when the tcgen05/TMA implicit-GEMM path is usable on the current device (sm_100) */
int msg_tensor_core_path_available(void);
/* Diagnostics (NULL unless the process runs with MSG_B200_TC_DEBUG=1): host-mapped words that CTA 0 of
 * every tcgen05 kernel fills with progress markers and its first shared-memory operand tiles. */
const uint32_t* msg_debug_buffer(size_t* words);

/* Per-kernel timing of the tcgen05 conv kernels with CUDA events recorded on the launching stream.
 * msg_profile_enable(1) resets and starts; after the caller has synchronised the stream(s),
 * msg_profile_summary() fills one entry per distinct (kind, taps, K channels, N channels, pixels). */
typedef struct {
  int kind;                 /* 0 = pixel-major GEMM (forward / dgrad), 1 = pixel-reduction GEMM (wgrad) */
  int taps, k_channels, n_channels;
  int64_t pixels;           /* B * PH * PW of one launch                                            */
  int64_t launches;
  double ms_total;          /* sum of event-measured durations                                      */
  double flops_per_launch;  /* algorithmic: 2 * pixels * n_channels * k_channels * taps             */
} msg_profile_entry;
void msg_profile_enable(int on);
int msg_profile_summary(msg_profile_entry* out, int max_entries);

/* ---- small-M linears of the generator (csrc/linear_ops.cu) ------------------------------------------------------------
 * StyleMapping (multi_stylegan_generator.py:208-235): pixel norm (equalized_layer.py:257-277), then `depth` x
 * [EqualizedLinear(K, K, bias=False) (:210-254) -> FusedLeakyReLU (op_static/fused_act.py:76-85)] as ONE launch
 * (a cluster of 8 CTAs per 16 rows, activations in distributed shared memory):
 *   x0 = z * rsqrt(mean_k z^2 + eps);  x_{l+1} = lrelu(alpha * x_l W_l^T + bias_l) * gain
 * acts [depth, M, K] receives every layer's output (acts[depth-1] is the result; the rest is what the backward needs),
 * x0 [M, K] the normalised input.  `weights` / `biases`: HOST arrays of `depth` device pointers (biases or its entries
 * may be NULL).  K: multiple of 4, <= 512.  Backward: dW [depth, K, K], db [depth, K] (NULL: skipped) from gy [M, K]
 * in two launches (the sequential chain on one cluster, then all weight gradients on the whole chip); workspace:
 * depth * M * K floats; the gradient w.r.t. z is not produced.  Deterministic. */
int msg_style_mapping_supported(int depth, int K);
int msg_style_mapping_forward(float* acts, float* x0, const float* z, const float* const* weights,
                              const float* const* biases, int depth, int M, int K, float alpha, float slope, float gain,
                              float eps, msg_stream_t stream);
int msg_style_mapping_backward(float* dW, float* db, const float* gy, const float* acts, const float* x0,
                               const float* const* weights, const float* const* biases, int depth, int M, int K,
                               float alpha, float slope, float gain, float* workspace, msg_stream_t stream);

/* A group of independent linears that read slices of ONE input [M, in_row] (the style linears `modulation_mapping` of all
 * ModulatedConv2d layers, multi_stylegan_generator.py:355-361, reading the per-layer latents) in one launch:
 *   out_i [M, N] = alpha * in[:, in_off : in_off + K] W^T + beta * bias
 * Item i's output is the contiguous block [out_off * M, (out_off + N) * M) of the flat `out` (out_off = sum of the
 * preceding items' N), the layout `gout` has as well.  Backward: dW (flat, item i at w_off), db (flat, b_off), din [M, in_row] (slots = the distinct input slices; input
 * columns no slot covers are NOT written).  At most 48 items; K multiple of 4, <= 2048.  Tables are host arrays. */
typedef struct msg_linear_item {
  const float* W;      /* [N, K] */
  const float* bias;   /* [N] or NULL */
  int N, K;
  int in_off, out_off;
  int w_off, b_off;
  float alpha, beta;
} msg_linear_item;
typedef struct msg_linear_slot { int in_off, K, first, count; } msg_linear_slot;   /* items [first, first+count) share the slice */
int msg_linear_group_forward(float* out, const float* in, int64_t in_row, const msg_linear_item* items,
                             int n_items, int M, int max_n, int max_k, msg_stream_t stream);
int msg_linear_group_backward(float* dW, float* db, float* din, const float* gout, const float* in,
                              int64_t in_row, const msg_linear_item* items, int n_items, const msg_linear_slot* slots,
                              int n_slots, int M, int max_n, int max_k, msg_stream_t stream);

/* ---- non-local (self-attention) block of the discriminator, u_net_2d_discriminator.py:332-381 (csrc/attention_ops.cu) ----
 * The block's three input 1x1 convolutions run as one GEMM with stacked filters and the two attention products as 1x1
 * convolutions with one filter bank per sample (msg_conv2d_forward / _wgrad, w_batch_stride != 0); these are the
 * memory-bound passes in between, channels-last fp32:
 *   msg_nl_split_pool:   qkv [B,H,W,cq+cq+cv] -> theta [B,H,W,cq] (copy), phi_p [B,H/2,W/2,cq], g_p [B,H/2,W/2,cv]
 *                        (F.max_pool2d(kernel 2, stride 2), :367-368) and idx [B,H/2,W/2,(cq+cv)/4] (four 8-bit argmax
 *                        positions per 32-bit word; first maximum in scan order)
 *   msg_nl_merge_unpool: the adjoint: dqkv from dtheta, dphi_p, dg_p, idx (every element of dqkv is written)
 *   msg_softmax_rows:    in-place softmax of `rows` rows of n floats (F.softmax(.., dim=-1), :372); n % 4 == 0, n <= 4096
 *   msg_softmax_rows_bwd: dp_inout <- p * (dp_inout - sum_j p_j dp_inout_j)
 * cq, cv: multiples of 4; pointers 16-byte aligned. */
int msg_nl_split_pool(float* theta, float* phi_p, float* g_p, uint32_t* idx, const float* qkv, int B, int H, int W, int cq,
                      int cv, msg_stream_t stream);
int msg_nl_merge_unpool(float* dqkv, const float* dtheta, const float* dphi_p, const float* dg_p, const uint32_t* idx,
                        int B, int H, int W, int cq, int cv, msg_stream_t stream);
int msg_softmax_rows(float* x, int64_t rows, int n, msg_stream_t stream);
int msg_softmax_rows_bwd(float* dp_inout, const float* p, int64_t rows, int n, msg_stream_t stream);

/* Backward of msg_demod_factors (what autograd derives from multi_stylegan_generator.py:386-388): from gd = dL/dd [B,O]
 *   ds [B,C] (NULL: skipped; needs wsq) and dW [O,C,taps] = W * 2 sum_b q[b,o] s[b,c]^2 (NULL: skipped; needs W),
 *   q = gd * d^3 * (-scale^2 / 2).  B <= 64. */
int msg_demod_factors_bwd(float* dW, float* ds, const float* gd, const float* d, const float* s, const float* wsq,
                          const float* W, int B, int O, int C, int taps, float scale, msg_stream_t stream);

/* Deterministic reductions over channels-last activations: out[c] = scale * sum_r x[r,c] for a dense [rows, C] matrix (the
 * bias gradient of a convolution without activation, equalized_layer.py:70-73; C % 4 == 0), and out[0] = scale * <a, b>
 * (n % 4 == 0; the gradient of the non-local block's gamma, u_net_2d_discriminator.py:381).  msg_dot workspace: 4096 floats. */
size_t msg_colsum_workspace(int64_t rows, int C);
int msg_colsum_nhwc(float* out, const float* x, int64_t rows, int C, float scale, void* workspace, size_t workspace_bytes,
                    msg_stream_t stream);
int msg_dot(float* out, const float* a, const float* b, int64_t n, float scale, float* workspace, msg_stream_t stream);

/* MinibatchStdDev (u_net_2d_discriminator.py:189-217), channels-last: out [B, HW, C+1] = cat(x [B, HW, C], plane) with
 * plane[b] = mean over (c, pos) of sqrt(max(var over b's sub-batch, alpha)); the batch is `groups` sub-batches of B / groups
 * consecutive samples (groups = 1: the reference; 2: the real and fake halves of one batched discriminator pass).
 * Backward: gx = gout[..., :C] + the plane's gradient through the standard deviation.  Deterministic. */
size_t msg_mbstd_workspace(int groups);
int msg_mbstd_forward(float* out, const float* x, int B, int C, int64_t HW, int groups, float alpha, void* workspace,
                      size_t workspace_bytes, msg_stream_t stream);
int msg_mbstd_backward(float* gx, const float* gout, const float* x, int B, int C, int64_t HW, int groups, float alpha,
                       void* workspace, size_t workspace_bytes, msg_stream_t stream);

/* Roofline probe (bench.py): one launch of `iters` x 4 back-to-back tcgen05.mma.kind::tf32 (128 x 256 x 8, operands in
 * shared memory, one CTA per SM); *flops receives the FLOPs of the launch.  sink: >= 32 * #SMs floats or NULL. */
int msg_tf32_mma_rate_probe(int iters, float* sink, double* flops, msg_stream_t stream);

/* -------------------------------------------------------------------------------------------
 * fused_bias_act  — replaces fused_act_cuda.fused_bias_act
 *   multi_stylegan/op_static/fused_bias_act.cpp:11-20, fused_bias_act_kernel.cu:18-99
 *   out[i] = act(x[i] + bias[(i / step_b) % size_b], ref[i]) * scale
 *     act*10+grad: 10,11 -> y=x; 12 -> 0; 30 -> x>0?x:alpha*x; 31 -> ref>0?x:alpha*x; 32 -> 0
 *   bias == NULL / ref == NULL mean "absent" (reference: numel()==0, kernel.cu:62-63).
 *   64-bit indexing (the reference's int32 index overflows at 2^31 elements, kernel.cu:21).
 * ------------------------------------------------------------------------------------------- */
int msg_fused_bias_act(void* out, const void* x, const void* bias, const void* ref,
                       int act, int grad, double alpha, double scale,
                       int64_t size_x, int64_t step_b, int64_t size_b,
                       int dtype, msg_stream_t stream);

/* Backward of the above fused with the bias-gradient reduction that the reference performs as
 * a separate ATen sum (op_static/fused_act.py:31-40):
 *   dx[i] = (ref[i] > 0 ? g[i] : alpha*g[i]) * scale ;  dbias[c] = sum_{i in channel c} dx[i]
 * x is viewed as [outer, size_b, step_b].  workspace: msg_fused_bias_act_bwd_workspace() bytes.
 * Deterministic (two-stage reduction, no atomics). */
size_t msg_fused_bias_act_bwd_workspace(int64_t size_x, int64_t step_b, int64_t size_b, int dtype);
int msg_fused_bias_act_bwd(void* dx, void* dbias, const void* g, const void* ref,
                           double alpha, double scale,
                           int64_t size_x, int64_t step_b, int64_t size_b,
                           void* workspace, size_t workspace_bytes,
                           int dtype, msg_stream_t stream);

/* -------------------------------------------------------------------------------------------
 * upfirdn2d — replaces upfirdn2d_cuda.upfirdn2d
 *   multi_stylegan/op_static/upfirdn2d.cpp:12-22, upfirdn2d_kernel.cu:52-272
 *   in  [major, in_h, in_w, minor], kernel [kernel_h, kernel_w] (applied flipped: true convolution,
 *   kernel.cu:77), out [major, out_h, out_w, minor] with
 *   out_h = (in_h*up_y + pad_y0 + pad_y1 - kernel_h + down_y) / down_y   (kernel.cu:167-168)
 *   Negative pads crop.  Every (up, down, pad, kernel<=32x32) combination is implemented; the
 *   reference launches nothing (returns uninitialised memory) outside its six modes.
 * ------------------------------------------------------------------------------------------- */
int msg_upfirdn2d_out_size(int in_size, int up, int down, int pad0, int pad1, int ksize);
int msg_upfirdn2d(void* out, const void* in, const void* kernel,
                  int64_t major, int in_h, int in_w, int minor,
                  int kernel_h, int kernel_w,
                  int up_x, int up_y, int down_x, int down_y,
                  int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                  int dtype, msg_stream_t stream);

/* FIR (up = down = 1) with the StyledConv2d tail fused into the store — the blur that follows an up-convolution
 * (multi_stylegan_generator.py:403) + noise (:289-292) + bias / leaky ReLU / gain (op_static/fused_act.py:58):
 *   out[b,y,x,c] = act(fir(in)[b,y,x,c] + noise_w[0]*noise[b*noise_batch_stride + y*out_w + x] + bias[c]) * gain
 * fp32 channels-last data only (in [major=B, h, w, minor=C], C % 4 == 0); MSG_ERR_UNSUPPORTED otherwise. */
int msg_upfirdn2d_bias_act(float* out, const float* in, const float* kernel, int64_t major, int in_h, int in_w,
                           int minor, int kernel_h, int kernel_w, int pad_x0, int pad_x1, int pad_y0, int pad_y1,
                           const float* noise, const float* noise_w, int64_t noise_batch_stride,
                           const float* bias, int act, float slope, float gain, msg_stream_t stream);

/* The same pass for the shared-weight form of the modulated up-convolution (multi_stylegan_generator.py:379-403 with the
 * modulation moved onto the activations): the demodulation factor commutes with the per-channel FIR, so
 *   out[b,y,x,c]  = act(col_scale[b*col_scale_batch_stride + c] * fir(in)[b,y,x,c] + noise term + bias[c]) * gain
 *   out2[b,y,x,c] = out[b,y,x,c] * out2_scale[b*out2_scale_batch_stride + c]      (optional: the next layer's input)
 * col_scale / out2 may be NULL; pointers 16-byte aligned, batch strides multiples of 4. */
int msg_upfirdn2d_bias_act_mod(float* out, float* out2, const float* in, const float* kernel, int64_t major,
                               int in_h, int in_w, int minor, int kernel_h, int kernel_w, int pad_x0, int pad_x1,
                               int pad_y0, int pad_y1, const float* col_scale, int64_t col_scale_batch_stride,
                               const float* noise, const float* noise_w, int64_t noise_batch_stride,
                               const float* bias, int act, float slope, float gain, const float* out2_scale,
                               int64_t out2_scale_batch_stride, msg_stream_t stream);

/* -------------------------------------------------------------------------------------------
 * Dense / per-sample ("grouped by batch") 2-D convolution primitives, fp32 storage, TF32 tensor
 * cores (tcgen05) with fp32 accumulation for the shapes the implicit-GEMM kernels tile, an fp32
 * CUDA-core kernel for the rest.  Replaces the ATen/cuDNN calls at
 *   multi_stylegan/multi_stylegan_generator.py:398 (conv_transpose2d, groups=B),
 *   multi_stylegan/multi_stylegan_generator.py:409 (conv2d, groups=B),
 *   multi_stylegan/equalized_layer.py:70-73        (conv2d)
 * and their autograd-generated dgrad / wgrad.
 *
 *   x  [B, C, H, W]            w  [O, C, kh, kw]        (w_batch_stride == 0, shared weights)
 *   y  [B, O, OH, OW]          w  [B, O, C, kh, kw]     (w_batch_stride == O*C*kh*kw)
 *   (logical shapes; activations are stored NCHW or channels-last NHWC, see msg_conv_desc.layout)
 *   OH = (H + 2*pad_h - kh) / stride_h + 1
 *
 * forward : y  = conv(x, w)
 * dgrad   : dx = conv^T(dy, w)     (also the forward of conv_transpose2d)
 * wgrad   : dw = sum_{b,p} dy (x) x   (per sample when dw_batch_stride != 0)
 * `alpha` scales the result (used to fold the equalised-lr constant); `flags` selects the engine.
 * ------------------------------------------------------------------------------------------- */
enum {
  MSG_LAYOUT_NCHW = 0,      /* dense [B,C,H,W]: CUDA-core engine                                   */
  MSG_LAYOUT_NHWC = 1       /* dense channels-last [B,H,W,C] (torch.channels_last): tcgen05 engine */
};

typedef struct {
  int B, C, H, W;          /* input activation (logical NCHW shape)              */
  int O, kh, kw;           /* filter, always stored [O, C, kh, kw]               */
  int stride_h, stride_w;
  int pad_h, pad_w;
  int OH, OW;              /* output activation extent (must match the formula)  */
  int64_t w_batch_stride;  /* 0 = shared weights, else elements between samples  */
  int layout;              /* memory layout of x, y, dx, dy: MSG_LAYOUT_*        */
  int w_transposed;        /* 0: filters [O, C, kh, kw]; 1: [C, O, kh, kw] — the layout of conv_transpose2d's weight
                              (multi_stylegan_generator.py:393-398), read / written in place by all three kernels */
} msg_conv_desc;

enum {
  MSG_CONV_AUTO = 0,        /* tensor cores when the shape tiles, else CUDA cores */
  MSG_CONV_FORCE_SIMT = 1,  /* fp32 CUDA-core kernel (exact fp32)                 */
  MSG_CONV_FORCE_TC = 2     /* fail with MSG_ERR_UNSUPPORTED if the shape does not tile */
};

/* Output transform fused into the forward kernel's epilogue (accumulators are read from TMEM once and never
 * round-trip through HBM before the activation) — the tail of StyledConv2d.forward
 * (multi_stylegan_generator.py:289-292 noise, op_static/fused_act.py:58 bias + leaky ReLU + gain) and of
 * ResNetBlock.forward (u_net_2d_discriminator.py:174-186: activation, residual add, 1/sqrt(2)):
 *   v = alpha * conv ; v += noise_w[0] * noise[b * noise_batch_stride + oy * OW + ox] ; v += bias[o]
 *   v = act ? (v > 0 ? v : slope * v) : v ; v += add[b, o, oy, ox] ; y = v * gain
 * Null pointers skip their term; `add` has the layout of y.  (col_scale / y2: see the field comments.) */
typedef struct {
  const float* bias;            /* [O]                                                  */
  const float* noise;           /* [B or 1, 1, OH, OW]                                   */
  const float* noise_w;         /* device scalar (NoiseInjection.weight)                 */
  int64_t noise_batch_stride;   /* OH*OW, or 0 when one map is shared by the batch       */
  const float* add;             /* tensor added after the activation, layout of y        */
  int act;                      /* 0 = linear, 1 = leaky ReLU                            */
  float slope, gain;
  /* Shared-weight form of the modulated convolution (multi_stylegan_generator.py:379-411 rewritten as
   * y = demod[b,o] * conv(scale * W, s[b,c] * x), algebraically identical): the accumulator is multiplied by
   * col_scale[b * col_scale_batch_stride + o] (the demodulation factor, :386-388) before the noise / bias terms, and
   * a second tensor y2 = y * y2_scale[b * y2_scale_batch_stride + o] (the next layer's style-modulated input,
   * :384) is written next to y.  Null pointers skip either; not combinable with `add`. */
  const float* col_scale;       /* [B or 1, O]                                           */
  int64_t col_scale_batch_stride;
  float* y2;                    /* layout of y                                           */
  const float* y2_scale;        /* [B or 1, O]                                           */
  int64_t y2_scale_batch_stride;
} msg_conv_epilogue;

size_t msg_conv2d_workspace(const msg_conv_desc* d, int which /*0 fwd,1 dgrad,2 wgrad*/, int flags);
int msg_conv2d_forward_fused(float* y, const float* x, const float* w, const msg_conv_desc* d,
                             float alpha, const msg_conv_epilogue* epilogue, void* workspace,
                             size_t workspace_bytes, int flags, msg_stream_t stream);
/* y = epilogue(alpha * conv([x1 | x2], w)): the channel concatenation of two NHWC tensors (c1 and C - c1 channels,
 * both multiples of 32) is read in place by the K loop — replaces torch.cat + conv at
 * u_net_2d_discriminator.py:137,174-186.  tcgen05 engine, stride 1 only; workspace = msg_conv2d_workspace(d, 0, flags).
 * MSG_ERR_UNSUPPORTED otherwise (the caller concatenates and uses msg_conv2d_forward_fused). */
int msg_conv2d_forward_cat2(float* y, const float* x1, int c1, const float* x2, const float* w,
                            const msg_conv_desc* d, float alpha, const msg_conv_epilogue* epilogue,
                            void* workspace, size_t workspace_bytes, int flags, msg_stream_t stream);
int msg_conv2d_forward(float* y, const float* x, const float* w, const msg_conv_desc* d,
                       float alpha, void* workspace, size_t workspace_bytes, int flags,
                       msg_stream_t stream);
int msg_conv2d_dgrad(float* dx, const float* dy, const float* w, const msg_conv_desc* d,
                     float alpha, void* workspace, size_t workspace_bytes, int flags,
                     msg_stream_t stream);
/* dx = alpha * conv^T(dy, w) + add: the sum of two input gradients (a block input that feeds the main path and the
 * residual path, u_net_2d_discriminator.py:174-186) formed in the dgrad epilogue instead of a separate pass.
 * `add` has the layout of dx and must not alias it; NULL = msg_conv2d_dgrad. */
int msg_conv2d_dgrad_acc(float* dx, const float* dy, const float* w, const msg_conv_desc* d,
                         float alpha, const float* add, void* workspace, size_t workspace_bytes, int flags,
                         msg_stream_t stream);
/* The dgrad of a layer fused with the activation backward of the layer that produced its input
 * (op_static/fused_act.py:31-40 behind u_net_2d_discriminator.py:174-186; `ref` = that layer's activation output, layout of dx):
 *   dx = alpha * conv^T(dy, w) * (ref > 0 ? 1 : slope) * gain ;   dbias[c] = sum_{b,y,x} dx[b,c,y,x]   (dbias may be NULL)
 * tcgen05 engine, NHWC, stride 1, C % 4 == 0 (at least 32 channels); msg_conv2d_dgrad_mask_supported()
 * answers 1 when the shape qualifies (otherwise run msg_conv2d_dgrad + msg_fused_bias_act_bwd).  Deterministic.
 * workspace: msg_conv2d_workspace(d, 1, flags). */
int msg_conv2d_dgrad_mask_supported(const msg_conv_desc* d, int flags);
int msg_conv2d_dgrad_mask(float* dx, float* dbias, const float* dy, const float* w, const msg_conv_desc* d,
                          float alpha, const float* ref, float slope, float gain, void* workspace,
                          size_t workspace_bytes, int flags, msg_stream_t stream);
int msg_conv2d_wgrad(float* dw, const float* dy, const float* x, const msg_conv_desc* d,
                     float alpha, void* workspace, size_t workspace_bytes, int flags,
                     msg_stream_t stream);
/* which engine the last conv call on this thread used: 0 none, 1 simt, 2 tcgen05 */
int msg_conv2d_last_engine(void);

/* -------------------------------------------------------------------------------------------
 * Weight modulation / demodulation — multi_stylegan/multi_stylegan_generator.py:384-388
 *   w_mod[b,o,c,t] = scale * W[o,c,t] * s[b,c] * (demod ? rsqrt(sum_{c,t}(scale*W*s)^2 + 1e-8) : 1)
 *   demod_out[b,o] (optional) receives the demodulation factor.
 * ------------------------------------------------------------------------------------------- */
int msg_modulate_weights(float* w_mod, float* demod_out, const float* W, const float* s,
                         int B, int O, int C, int taps, float scale, int demodulate,
                         msg_stream_t stream);

/* First-order backward of the above (what autograd derives from multi_stylegan_generator.py:384-388):
 * given g = dL/dw_mod [B,O,C,taps], the shared weight W, the styles s [B,C] and the demodulation factors the forward
 * returned, writes dW [O,C,taps] and ds [B,C].  workspace: msg_modulate_weights_bwd_workspace() bytes. */
size_t msg_modulate_weights_bwd_workspace(int B, int O, int C, int taps);
int msg_modulate_weights_bwd(float* dW, float* ds, const float* g, const float* W, const float* s,
                             const float* demod, int B, int O, int C, int taps, float scale, int demodulate,
                             void* workspace, size_t workspace_bytes, msg_stream_t stream);

/* -------------------------------------------------------------------------------------------
 * Shared-weight form of the modulated convolution (multi_stylegan_generator.py:379-411):
 *   conv(scale*W*s*demod, x) == demod[b,o] * conv(scale*W, s[b,c]*x)        (exact algebra; the survey checked it
 * against the reference module in fp64).  The style multiplies the activations (written by the previous layer's
 * epilogue as its second output), the demodulation factor multiplies the accumulator (msg_conv_epilogue.col_scale),
 * and all samples share ONE weight operand, so wgrad is one batch-reduced GEMM instead of B per-sample ones.
 *
 * msg_demod_factors: wsq[o,c] = sum_t W[o,c,t]^2 ; d[b,o] = rsqrt(scale^2 * sum_c s[b,c]^2 * wsq[o,c] + 1e-8)  (:386-388)
 *
 * msg_styled_act_bwd: the non-GEMM part of the layer's first-order backward in one pass over the channels-last
 * activation [B, rows = H*W, C] (C % 4 == 0), given the gradients w.r.t. both outputs of the forward epilogue
 * (out = lrelu(v) * gain, out2 = out * out2_scale; v = col_scale * conv + noise_w * noise + bias):
 *   gt = g_out + out2_scale[b,c] * g_out2 ;  gv = gt * (out > 0 ? 1 : slope) * gain ;  g_pre = gv * col_scale[b,c]
 *   sums[0][b][c] = sum_p gv               (-> dbias)
 *   sums[1][b][c] = sum_p gv * v           (-> d col_scale, with v recovered from out)
 *   sums[2][b][c] = sum_p gv * noise[b,p]  (-> dnoise_w)
 *   sums[3][b][c] = sum_p out * g_out2     (-> d out2_scale)
 * g_out or g_out2 may be NULL (not both); col_scale NULL = 1; noise NULL = 0.  Deterministic (two-stage reduction in
 * `workspace`, msg_styled_act_bwd_workspace bytes).  Replaces what autograd derives from :384-388, :289-292 and
 * op_static/fused_act.py:31-40.
 * ------------------------------------------------------------------------------------------- */
int msg_demod_factors(float* d, float* wsq, const float* W, const float* s, int B, int O, int C, int taps,
                      float scale, msg_stream_t stream);
size_t msg_styled_act_bwd_workspace(int B, int64_t rows, int C);
int msg_styled_act_bwd(float* g_pre, float* sums, const float* g_out, const float* g_out2, const float* out,
                       const float* col_scale, int64_t col_scale_batch_stride, const float* out2_scale,
                       int64_t out2_scale_batch_stride, const float* noise, int64_t noise_batch_stride,
                       int B, int64_t rows, int C, float slope, float gain, void* workspace,
                       size_t workspace_bytes, msg_stream_t stream);

/* -------------------------------------------------------------------------------------------
 * Fused StyledConv2d epilogue — multi_stylegan_generator.py:292 (noise) + op_static/fused_act.py:58
 *   out[b,c,p] = lrelu(x[b,c,p] + noise_w * noise[b or 0, p] + bias[c], alpha) * scale
 * noise may be NULL (noise_w ignored); noise_batch_stride is 0 when one map is shared by the batch.
 * ------------------------------------------------------------------------------------------- */
int msg_noise_bias_act(float* out, const float* x, const float* noise, const float* noise_w,
                       const float* bias, int B, int C, int64_t HW, int64_t noise_batch_stride,
                       float alpha, float scale, msg_stream_t stream);

/* Same epilogue on channels-last activations x [rows = B*H*W, C] (C % 4 == 0, 16-byte aligned), optionally in
 * its masked ("backward") form, plus the fused first-order backward:
 *   forward      (ref == NULL): out = lrelu(x + noise_w[0]*noise[row % noise_period] + bias[c], alpha) * scale
 *   masked       (ref != NULL): out = (ref > 0 ? v : alpha*v) * scale, v as above  — the double-backward
 *                               (op_static/fused_act.py:44-51 extended by the noise term)
 *   backward: dx = (ref > 0 ? g : alpha*g) * scale ; dbias[c] = sum_rows dx ; dnoise_w[0] = sum noise[row] * dx
 *             (dbias / dnoise_w may be NULL; deterministic two-stage reduction in `workspace`).
 * noise_period = H*W when one noise map is shared by the batch, rows when every sample has its own. */
int msg_noise_bias_act_nhwc(float* out, const float* x, const float* ref, const float* noise,
                            const float* noise_w, const float* bias, int64_t rows, int C,
                            int64_t noise_period, float alpha, float scale, msg_stream_t stream);
size_t msg_noise_bias_act_nhwc_bwd_workspace(int64_t rows, int C);
int msg_noise_bias_act_nhwc_bwd(float* dx, float* dbias, float* dnoise_w, const float* g, const float* ref,
                                const float* noise, int64_t rows, int C, int64_t noise_period, float alpha,
                                float scale, void* workspace, size_t workspace_bytes, msg_stream_t stream);

/* -------------------------------------------------------------------------------------------
 * ADA geometric warp — multi_stylegan/adaptive_discriminator_augmentation.py:116-199
 *   out[b,c,y,x] = bilinear sample of in[b,c] at (x,y,1) * theta[b]^T  in pixel coordinates,
 *   reflection padding, align_corners=True (grid_sample semantics); theta [B,2,3] is the
 *   composed inverse map of every stage applied to sample b (identity rows leave the sample
 *   untouched bit-exactly).  mode 0 = reflection padding, 1 = zeros (kornia rotate).
 * ------------------------------------------------------------------------------------------- */
int msg_affine_warp(float* out, const float* in, const float* theta, int B, int C, int H, int W,
                    int mode, msg_stream_t stream);
/* Gradient of the warp w.r.t. its input image (the generator step differentiates through the augmentation like the
 * reference's kornia warps do): dx = A^T g with the same bilinear weights, scattered with fp32 atomics. */
int msg_affine_warp_bwd(float* dx, const float* g, const float* theta, int B, int C, int H, int W,
                        int mode, msg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MSG_B200_H_ */
