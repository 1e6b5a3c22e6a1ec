// Internal problem descriptions shared by the conv engines (CUDA-core and tcgen05).
//
// Every convolution the hot path needs (forward, dgrad, wgrad; stride 1 or 2; shared or per-sample
// weights; transposed) is lowered by conv_api.cu to one of two "gather GEMMs" over strided 4-D views
// (so NCHW and channels-last NHWC are the same code):
//
//   PixGemm  :  out[b, n, y, x]   = alpha * sum_{t, c} in[b, c, y*my + dy_t, x*mx + dx_t] * w(b, n, c, t)
//               pixels are the GEMM M dimension; output may be scattered with a stride/offset
//               (phase-wise transposed conv).
//   RedGemm  :  dw(b?, n, c, t)   = alpha * sum_{b?, y, x} g[b, n, y, x] * in[b, c, y*my + dy_t, x*mx + dx_t]
//               pixels are the GEMM K dimension.
//
// The tcgen05 engine needs channels-last operands (channel stride 1): then a filter tap is a TMA
// coordinate shift in a non-innermost dimension (TMA faults on an innermost start coordinate that is
// not 16-byte aligned — measured on B200 — so W-shifts of NCHW rows cannot be expressed).
#pragma once
#include "common.cuh"

namespace msg {

constexpr int kMaxTaps = 16;

struct View4 {            // element strides of a [B, C, H, W] tensor
  int64_t sb, sc, sy, sx;
};

// Fused output transform of a PixGemm (StyledConv2d / ResNetBlock epilogues):
//   v = alpha * acc * cscale[b * cscale_sb + n]                (cscale != nullptr: per-sample demodulation factor)
//   v += noise_w[0] * noise[b * noise_sb + y * PW + x]         (noise != nullptr; unscattered outputs only)
//   v += bias[n]                                               (bias != nullptr)
//   v = v > 0 ? v : slope * v                                  (act == 1)
//   v += add[same offset as out]                               (add != nullptr)
//   out = v * gain
//   out2 = out * out2_scale[b * out2_scale_sb + n]             (out2 != nullptr: the next layer's modulated input)
struct Epilogue {
  const float* bias;
  const float* noise;
  const float* noise_w;
  int64_t noise_sb;
  const float* add;
  int act;
  float slope, gain;
  const float* cscale;      // [B or 1, N] multiplier of the accumulator (before noise / bias)
  int64_t cscale_sb;
  float* out2;              // second output with the layout of `out`
  const float* out2_scale;  // [B or 1, N]
  int64_t out2_scale_sb;
  // mask mode (the activation backward of the PRODUCING layer fused into this dgrad): `add` is that layer's activation
  // output and multiplies instead of adds:  out = alpha * acc * (add > 0 ? 1 : slope) * gain;  with `colsum` the
  // per-channel sums of `out` (the bias gradient) are accumulated per CTA and epilogue warp into rows that this warp
  // alone owns (pre-zeroed by the caller): colsum[(cta * 4 + warp) * N + n]  (tcgen05 engine, TMA-store epilogue)
  int add_is_mask;
  float* colsum;
  __host__ __device__ bool any() const { return bias || noise || add || act || gain != 1.f || cscale || out2; }
};

__device__ __forceinline__ float apply_epilogue(const Epilogue& e, float v, float bias_n, float noise_term, float addv) {
  v += noise_term + bias_n;
  if (e.act) v = v > 0.f ? v : v * e.slope;
  return (v + addv) * e.gain;
}

struct PixGemm {
  const float* in;        // [B, Cr, IH, IW]
  View4 is;
  int B, Cr, IH, IW;
  int my, mx;             // input coordinate = pixel * m + tap_d (tcgen05 engine: 1 only)
  // tcgen05 engine, stride-2 convolutions without a space-to-depth copy: the K dimension is nphase groups of
  // Cr / nphase channels and group ph is read from its own strided view in_ph[ph] [B, Cr/nphase, IH_ph, IW_ph]
  // (strides `is`) — the input phase x[b, 2*y + py, 2*x + px, c].  nphase == 0: the single view `in`.
  int nphase;
  const float* in_ph[4];
  int IH_ph[4], IW_ph[4];
  // tcgen05 engine, channel concatenation without a copy (U-Net decoder, u_net_2d_discriminator.py:137): nsrc == 2 reads
  // the K chunks [0, C_src[0] / 32) from the dense NHWC tensor in_src[0] and the remaining ones from in_src[1].
  int nsrc;
  const float* in_src[2];
  int C_src[2];
  const float* w;         // w(b,n,c,t) = w[b*w_sb + n*w_sn + c*w_sc + t*w_st]
  int64_t w_sb, w_sn, w_sc, w_st;
  int N;
  int ntaps;
  int tap_dy[kMaxTaps], tap_dx[kMaxTaps], tap_wi[kMaxTaps];
  int PH, PW;             // pixel grid iterated per sample
  float* out;             // out[b*os.sb + n*os.sc + (y*out_my+out_oy)*os.sy + (x*out_mx+out_ox)*os.sx]
  View4 os;
  int out_my, out_mx, out_oy, out_ox;
  float alpha;
  Epilogue ep;            // zero-initialised = plain alpha * acc (gain must then be set to 1)
  int ksplit;             // tcgen05 engine, internal: > 1 = this call is the tap-split pass of a small problem (conv_tc.cu)
};

struct RedGemm {
  const float* g;         // [B, N, PH, PW]  ("dy")
  View4 gs;
  int B, N, PH, PW;
  const float* in;        // [B, C, IH, IW]
  View4 is;
  int C, IH, IW;
  int my, mx;
  int ntaps;
  int tap_dy[kMaxTaps], tap_dx[kMaxTaps], tap_wi[kMaxTaps];
  float* dw;              // dw[b*dw_sb + n*dw_sn + c*dw_sc + t*dw_st]; dw_sb == 0 -> reduce over the batch
  int64_t dw_sb, dw_sn, dw_sc, dw_st;
  float alpha;
};

// engines
int simt_pixgemm(const PixGemm& g, cudaStream_t st);
int simt_redgemm(const RedGemm& g, cudaStream_t st);

bool tc_available();
uint32_t* tc_debug_host(size_t* words);
void tc_profile_enable(int on);
int tc_profile_summary(msg_profile_entry* out, int max_entries);
bool tc_pixgemm_supported(const PixGemm& g);
bool tc_pixgemm_mask_ok(const PixGemm& g);
size_t tc_pixgemm_workspace(const PixGemm& g);
int tc_pixgemm(const PixGemm& g, void* ws, size_t ws_bytes, cudaStream_t st);
bool tc_redgemm_supported(const RedGemm& g);
size_t tc_redgemm_workspace(const RedGemm& g);
int tc_redgemm(const RedGemm& g, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace msg
