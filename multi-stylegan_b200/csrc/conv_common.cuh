// Internal problem descriptions shared by the conv engines (CUDA-core and tcgen05).
//
// Every convolution the hot path needs (forward, dgrad, wgrad; stride 1 or 2; shared or per-sample
// weights; transposed) is lowered by conv_api.cu to one of two stride-1 "gather GEMMs":
//
//   PixGemm  :  out[b, n, y, x]   = alpha * sum_{t, c} in[b, c, y*sy + dy_t, x*sx + dx_t] * w(b, n, c, t)
//               pixels are the GEMM M dimension (so the A operand is read straight from NCHW),
//               output may be scattered with a stride/offset (phase-wise transposed conv).
//   RedGemm  :  dw(b?, n, c, t)   = alpha * sum_{b?, y, x} g[b, n, y, x] * in[b, c, y + dy_t, x + dx_t]
//               pixels are the GEMM K dimension (both operands K-major in NCHW).
#pragma once
#include "common.cuh"

namespace msg {

constexpr int kMaxTaps = 16;

struct PixGemm {
  const float* in;        // [B, Cr, IH, IW] with explicit strides
  int B, Cr, IH, IW;
  int64_t in_sb, in_sc;   // element strides: batch, channel
  int in_pitch;           // row pitch in elements
  int in_sy, in_sx;       // input coordinate = pixel * in_s + tap_d  (tcgen05 engine: 1 only)
  const float* w;         // w(b,n,c,t) = w[b*w_sb + n*w_sn + c*w_sc + t*w_st]
  int64_t w_sb, w_sn, w_sc, w_st;
  int N;
  int ntaps;
  int tap_dy[kMaxTaps], tap_dx[kMaxTaps], tap_wi[kMaxTaps];
  int PH, PW;             // pixel grid iterated per sample
  float* out;             // out[b*out_sb + n*out_sn + (y*out_sy+out_oy)*out_pitch + x*out_sx+out_ox]
  int64_t out_sb, out_sn;
  int out_pitch, out_sy, out_sx, out_oy, out_ox;
  float alpha;
};

struct RedGemm {
  const float* g;         // [B, N, PH, PW] with explicit strides ("dy")
  int B, N, PH, PW;
  int64_t g_sb, g_sn;
  int g_pitch;
  const float* in;        // [B, C, IH, IW]
  int C, IH, IW;
  int64_t in_sb, in_sc;
  int in_pitch;
  int in_sy, in_sx;       // input coordinate = pixel * in_s + tap_d  (tcgen05 engine: 1 only)
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
bool tc_pixgemm_supported(const PixGemm& g);
size_t tc_pixgemm_workspace(const PixGemm& g);
int tc_pixgemm(const PixGemm& g, void* ws, size_t ws_bytes, cudaStream_t st);
bool tc_redgemm_supported(const RedGemm& g);
size_t tc_redgemm_workspace(const RedGemm& g);
int tc_redgemm(const RedGemm& g, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace msg
