// tcgen05 / TMEM / TMA engine (sm_100a) for the two gather GEMMs of conv_common.cuh, on channels-last
// (NHWC) activations.
//
// PixGemm (forward / dgrad / transposed conv):  D[128 pixels, BN channels] += A * B per (tap, 32-channel
//   chunk).  A is read by TMA straight out of the activation: box (32 ch, Wt px, 128/Wt rows) of the
//   4-D view (C, W, H, B) lands in shared memory as 128 pixel rows of 128 bytes = a K-major UMMA
//   operand with the hardware 128B swizzle; a filter tap is a shifted TMA coordinate in W/H and the
//   conv zero padding is TMA out-of-bounds fill.  B is the pre-transformed, TF32-rounded weight tile
//   [tap][n][c] (K-major, 128B swizzle).  Accumulator: 128 lanes x BN columns of TMEM, read back with
//   tcgen05.ld; a lane is a pixel, whose BN output channels are contiguous in NHWC (128-bit stores).
// RedGemm (wgrad):  D[128 out-ch, BN in-ch] += dy[32 px, n]^T * x[32 px (shifted), c]; pixels are K, so
//   both operands are MN-major: the only legal TF32 layout for that is 128-byte rows swizzled in
//   32-byte chunks (UMMA SWIZZLE_128B_BASE32B == TMA SWIZZLE_128B_ATOM_32B), one TMA box per group of
//   32 channels.  Split-K over pixels into a workspace + deterministic reduce.
//
// Warp roles (192 threads, 1 CTA/SM): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one
// elected lane), warps 2..5 = epilogue (TMEM lane quarter = warp & 3).  smem full/empty mbarrier ring.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <utility>
#include <vector>

#include "conv_common.cuh"
#include "sm100_ptx.cuh"

namespace msg {

using namespace ptx;

// ------------------------------------------------------------------------------------------------
// TMA descriptor encoding through the driver entry point (no link-time libcuda dependency).
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    (void)cudaGetLastError();
  });
  return fn;
}

bool tc_available() {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    bool ok = cudaGetDevice(&dev) == cudaSuccess &&
              cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess && major == 10 &&
              get_encode() != nullptr;
    (void)cudaGetLastError();
    const char* off = getenv("MSG_B200_DISABLE_TC");
    if (off && off[0] == '1') ok = false;
    cached = ok ? 1 : 0;
  }
  return cached == 1;
}

// Diagnostics: with MSG_B200_TC_DEBUG=1 CTA 0 of every tcgen05 kernel records progress markers and dumps
// its first A/B shared-memory tiles into host-mapped memory (readable even after a device fault), and
// barrier waits time out softly instead of trapping.
constexpr size_t kDbgWords = 1 << 18;
static uint32_t* g_dbg_host = nullptr;
static uint32_t* g_dbg_dev = nullptr;
static uint32_t* tc_debug_buffer() {
  static int init = 0;
  if (!init) {
    init = 1;
    const char* e = getenv("MSG_B200_TC_DEBUG");
    if (e && e[0] == '1') {
      void* h = nullptr;
      if (cudaHostAlloc(&h, kDbgWords * 4, cudaHostAllocMapped) == cudaSuccess) {
        memset(h, 0, kDbgWords * 4);
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) { g_dbg_host = (uint32_t*)h; g_dbg_dev = (uint32_t*)d; }
      }
      (void)cudaGetLastError();
    }
  }
  return g_dbg_dev;
}
uint32_t* tc_debug_host(size_t* words) { tc_debug_buffer(); if (words) *words = kDbgWords; return g_dbg_host; }

__device__ __forceinline__ void dbg_set(uint32_t* dbg, int idx, uint32_t v) {
  if (dbg) { *reinterpret_cast<volatile uint32_t*>(dbg + idx) = v; __threadfence_system(); }
}

static uint32_t tc_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MSG_B200_TC_VARIANT");
    v = e ? atoi(e) : 0;
  }
  return (uint32_t)v;
}

// ------------------------------------------------------------------------------------------------
// Per-kernel timing with CUDA events on the launching stream (bench.py's roofline numbers).
// ------------------------------------------------------------------------------------------------
struct ProfSlot { msg_profile_entry e; };
static bool g_prof_on = false;
static std::vector<ProfSlot> g_prof_slots;
struct ProfPair { cudaEvent_t a, b; int slot; };
static std::vector<ProfPair> g_prof_pending;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;
static std::mutex g_prof_mu;

static int prof_begin(int kind, int taps, int k_channels, int n_channels, int64_t pixels, double flops,
                      cudaStream_t st, cudaEvent_t* stop_out) {
  if (!g_prof_on) return -1;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int slot = -1;
  for (size_t i = 0; i < g_prof_slots.size(); ++i) {
    const msg_profile_entry& e = g_prof_slots[i].e;
    if (e.kind == kind && e.taps == taps && e.k_channels == k_channels && e.n_channels == n_channels &&
        e.pixels == pixels) { slot = (int)i; break; }
  }
  if (slot < 0) {
    if (g_prof_slots.size() >= 256) return -1;
    ProfSlot s{}; s.e.kind = kind; s.e.taps = taps; s.e.k_channels = k_channels; s.e.n_channels = n_channels;
    s.e.pixels = pixels; s.e.flops_per_launch = flops;
    g_prof_slots.push_back(s);
    slot = (int)g_prof_slots.size() - 1;
  }
  cudaEvent_t a, b;
  if (!g_prof_pool.empty()) { a = g_prof_pool.back().first; b = g_prof_pool.back().second; g_prof_pool.pop_back(); }
  else if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return -1;
  cudaEventRecord(a, st);
  g_prof_pending.push_back({a, b, slot});
  *stop_out = b;
  return slot;
}
static void prof_end(int slot, cudaEvent_t stop, cudaStream_t st) { if (slot >= 0) cudaEventRecord(stop, st); }

void tc_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  for (auto& p : g_prof_pending) g_prof_pool.push_back({p.a, p.b});
  g_prof_pending.clear();
  g_prof_slots.clear();
}
int tc_profile_summary(msg_profile_entry* out, int max_entries) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& p : g_prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      g_prof_slots[p.slot].e.launches += 1;
      g_prof_slots[p.slot].e.ms_total += ms;
    }
    g_prof_pool.push_back({p.a, p.b});
  }
  (void)cudaGetLastError();
  g_prof_pending.clear();
  int n = 0;
  for (auto& s : g_prof_slots) { if (n < max_entries) out[n++] = s.e; }
  return n;
}

// 4-D fp32 tensor map; dims[0] is the contiguous dimension, strides_bytes for dims 1..3.
static int make_tmap(CUtensorMap* m, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                     const uint32_t box[4], CUtensorMapSwizzle swz) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(MSG_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gs[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MSG_ERR_CUDA,
                "cuTensorMapEncodeTiled failed (%d): dims %llu,%llu,%llu,%llu strides %llu,%llu,%llu box %u,%u,%u,%u",
                (int)r, (unsigned long long)gd[0], (unsigned long long)gd[1], (unsigned long long)gd[2],
                (unsigned long long)gd[3], (unsigned long long)gs[0], (unsigned long long)gs[1],
                (unsigned long long)gs[2], bx[0], bx[1], bx[2], bx[3]);
  return MSG_OK;
}

// 5-D fp32 tensor map (channel atoms as the slowest box dimension, see tc_redgemm_kernel).
static int make_tmap5(CUtensorMap* m, const void* base, const uint64_t dims[5], const uint64_t strides_bytes[4],
                      const uint32_t box[5], CUtensorMapSwizzle swz) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return fail(MSG_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gd[5] = {dims[0], dims[1], dims[2], dims[3], dims[4]};
  cuuint64_t gs[4] = {strides_bytes[0], strides_bytes[1], strides_bytes[2], strides_bytes[3]};
  cuuint32_t bx[5] = {box[0], box[1], box[2], box[3], box[4]};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), gd, gs, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MSG_ERR_CUDA, "cuTensorMapEncodeTiled(5d) failed (%d): dims %llu,%llu,%llu,%llu,%llu box %u,%u,%u,%u,%u",
                (int)r, (unsigned long long)gd[0], (unsigned long long)gd[1], (unsigned long long)gd[2],
                (unsigned long long)gd[3], (unsigned long long)gd[4], bx[0], bx[1], bx[2], bx[3], bx[4]);
  return MSG_OK;
}

// ------------------------------------------------------------------------------------------------
// Weight transform for PixGemm: wt[bw][t][n][c] = tf32_rna(w(b, n, c, tap_wi[t])), zero padded.
// ------------------------------------------------------------------------------------------------
struct WtParams {
  const float* w;
  int64_t w_sb, w_sn, w_sc, w_st;
  int N, Cr, Npad, Cpad, ntaps, BW;
  int tap_wi[kMaxTaps];
};

__device__ __forceinline__ float to_tf32_rna(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// One block = a 32 (n) x 32 (c) tile, all taps.  Global reads run along whichever of n / c is the weight tensor's
// faster dimension (forward: c, dgrad: n; the `taps` floats in between stay in L1 for the following tap passes),
// the tile is transposed through shared memory, and the writes are 128-byte runs along c.
__global__ void __launch_bounds__(256)
tc_weight_transform_kernel(float* __restrict__ wt, const WtParams p) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, n0 = blockIdx.y * 32, b = blockIdx.z;
  const int a = threadIdx.x & 31, r = threadIdx.x >> 5;
  const bool c_fast = p.w_sc <= p.w_sn;
  const float* wb = p.w + (int64_t)b * p.w_sb;
  for (int t = 0; t < p.ntaps; ++t) {
    const int64_t toff = (int64_t)p.tap_wi[t] * p.w_st;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int nl = c_fast ? r + 8 * k : a, cl = c_fast ? a : r + 8 * k;
      const int n = n0 + nl, c = c0 + cl;
      float v = 0.f;
      if (c < p.Cr && n < p.N) v = to_tf32_rna(__ldg(wb + (int64_t)n * p.w_sn + (int64_t)c * p.w_sc + toff));
      tile[nl][cl] = v;
    }
    __syncthreads();
    float* dst = wt + (((int64_t)b * p.ntaps + t) * p.Npad + n0) * p.Cpad + c0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int nl = r + 8 * k;
      if (n0 + nl < p.Npad) dst[(int64_t)nl * p.Cpad + a] = tile[nl][a];
    }
    __syncthreads();
  }
}

// The same transform for filters whose taps are contiguous (w_st == 1, T = kh * kw floats per (n, c) pair, every tap
// used): one block = a 32 x 32 (n, c) tile; each of its 32 rows along the slower of n / c is ONE contiguous run of 32 * T
// floats in the filter tensor, read fully coalesced in a single pass into shared memory (row pitch 32 * T + 1 floats keeps
// both read-out orders bank-conflict free), then written per tap as 128-byte rows along c.  The kernel above touches every
// 128-byte line T times through L1 (3 TB/s); this one streams (the per-sample 512 x 512 x 9 banks of the generator are
// 75 MB each way).
__global__ void __launch_bounds__(256)
tc_weight_transform_runs_kernel(float* __restrict__ wt, const WtParams p, int T) {
  extern __shared__ float runs[];                     // [32][32 * T + 1]
  const int pitch = 32 * T + 1;
  const int c0 = blockIdx.x * 32, n0 = blockIdx.y * 32, b = blockIdx.z;
  const int a = threadIdx.x & 31, r = threadIdx.x >> 5;
  const bool c_fast = p.w_sc <= p.w_sn;               // forward: rows = n, runs along (c, t); dgrad: rows = c, runs along (n, t)
  const float* wb = p.w + (int64_t)b * p.w_sb;
  const int slow0 = c_fast ? n0 : c0, fast0 = c_fast ? c0 : n0;
  const int slow_n = c_fast ? p.N : p.Cr, fast_n = c_fast ? p.Cr : p.N;
  const int64_t slow_stride = c_fast ? p.w_sn : p.w_sc;
  const int run = 32 * T;
  const int run_valid = (fast_n - fast0 < 32 ? (fast_n - fast0 > 0 ? fast_n - fast0 : 0) : 32) * T;
  for (int row = r; row < 32; row += 8) {
    const bool row_ok = slow0 + row < slow_n;
    const float* src = wb + (int64_t)(slow0 + row) * slow_stride + (int64_t)fast0 * T;
    for (int i = a; i < run; i += 32)
      runs[row * pitch + i] = (row_ok && i < run_valid) ? to_tf32_rna(__ldg(src + i)) : 0.f;
  }
  __syncthreads();
  for (int t = 0; t < p.ntaps; ++t) {
    const int ti = p.tap_wi[t];
    float* dst = wt + (((int64_t)b * p.ntaps + t) * p.Npad + n0) * p.Cpad + c0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int nl = r + 8 * k;
      const float v = c_fast ? runs[nl * pitch + a * T + ti] : runs[a * pitch + nl * T + ti];
      if (n0 + nl < p.Npad) dst[(int64_t)nl * p.Cpad + a] = v;
    }
  }
}

// Third form, same contract as the runs kernel: tiles of 8 (n) x 32 (c) x T instead of 32 x 32 x T.  The 32 x 32 tiles give
// a 512 x 512 layer only 256 blocks of 37 KB shared memory each — under two per SM, every thread walking its T loads one
// after the other — and the kernel ran at 20 % of the HBM bandwidth (15.6 us for 18.8 MB, 186 launches = 1.45 ms per
// iteration).  Small tiles put ~7 blocks on every SM with T independent loads in flight per thread.
template <int TMAX>
__global__ void __launch_bounds__(256)
tc_weight_transform_runs8_kernel(float* __restrict__ wt, const WtParams p, int T) {
  extern __shared__ float runs[];
  const int c0 = blockIdx.x * 32, n0 = blockIdx.y * 8, b = blockIdx.z;
  const bool c_fast = p.w_sc <= p.w_sn;               // forward: rows = n, runs along (c, t); dgrad: rows = c, runs along (n, t)
  const float* wb = p.w + (int64_t)b * p.w_sb;
  const int tid = threadIdx.x;
  if (c_fast) {
    const int pitch = 32 * T + 1;                     // runs[8][32 * T + 1]
    const int row = tid >> 5, a = tid & 31;
    const bool row_ok = n0 + row < p.N;
    const int valid = (p.Cr - c0 < 32 ? (p.Cr - c0 > 0 ? p.Cr - c0 : 0) : 32) * T;
    const float* src = wb + (int64_t)(n0 + row) * p.w_sn + (int64_t)c0 * T;
    float v[TMAX];
#pragma unroll
    for (int k = 0; k < TMAX; ++k) {
      const int i = a + 32 * k;
      v[k] = (k < T && row_ok && i < valid) ? __ldg(src + i) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < TMAX; ++k)
      if (k < T) runs[row * pitch + a + 32 * k] = to_tf32_rna(v[k]);
  } else {
    const int pitch = 8 * T + 1;                      // runs[32][8 * T + 1]
    const int row = tid >> 3, a = tid & 7;
    const bool row_ok = c0 + row < p.Cr;
    const int valid = (p.N - n0 < 8 ? (p.N - n0 > 0 ? p.N - n0 : 0) : 8) * T;
    const float* src = wb + (int64_t)(c0 + row) * p.w_sc + (int64_t)n0 * T;
    float v[TMAX];
#pragma unroll
    for (int k = 0; k < TMAX; ++k) {
      const int i = a + 8 * k;
      v[k] = (k < T && row_ok && i < valid) ? __ldg(src + i) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < TMAX; ++k)
      if (k < T) runs[row * pitch + a + 8 * k] = to_tf32_rna(v[k]);
  }
  __syncthreads();
  const int nl = tid >> 5, cl = tid & 31;             // one (n, c) pair per thread, all taps
  if (n0 + nl >= p.Npad) return;
  for (int t = 0; t < p.ntaps; ++t) {
    const int ti = p.tap_wi[t];
    const float val = c_fast ? runs[nl * (32 * T + 1) + cl * T + ti] : runs[cl * (8 * T + 1) + nl * T + ti];
    wt[(((int64_t)b * p.ntaps + t) * p.Npad + n0 + nl) * p.Cpad + c0 + cl] = val;
  }
}

// ------------------------------------------------------------------------------------------------
// PixGemm kernel (K-major A from NHWC, K-major B)
// ------------------------------------------------------------------------------------------------
struct TMapSet { CUtensorMap m[4]; };   // activation views (one per stride-2 input phase; m[0] otherwise)

struct TcPixParams {
  int ntaps, cchunks;
  int cpv[4];                   // K chunks served by activation view 0..3 in turn (stride-2 phases, concatenated sources)
  int tap_dy[kMaxTaps], tap_dx[kMaxTaps];
  int PH, PW, N;
  int wt_log2, tiles_x, tiles_y, n_tiles;   // tile = (1 << wt_log2) x (128 >> wt_log2) pixels
  int total_tiles;
  float* out;
  View4 os;
  int out_my, out_mx, out_oy, out_ox;
  float alpha;
  int w_per_sample;
  int vec_store;               // NHWC output, 16-byte aligned channel runs, N % 4 == 0
  int tma_store;               // epilogue writes 32-pixel x 32-channel boxes with TMA
  int debug;                   // MSG_B200_TC_VARIANT bits 32/64 (timing experiments only): 1 = no stores, 2 = no proxy fence (wrong results)
  int kw;                      // row-tap kernel: taps per filter row (taps are stored row-major, dx consecutive)
  int ksplit, tpg, bsz;        // tc_pixgemm_kernel: tap split (ksplit groups of tpg taps, partial sums of group g stored as
                               // samples [g * bsz, (g + 1) * bsz) of the output view); ksplit = 1: none
  Epilogue ep;
};

// One 32-channel chunk of the TMA-store epilogue: fused transform in registers, then the pixel's 128-byte row goes to
// the swizzled staging box.  The epilogue warps are one warp per scheduler, so this loop is bound by its instruction
// count: the three common transforms are compiled select-free (NB = noise/bias term, ACT = leaky ReLU, ADD = residual).
template <bool NB, bool ACT, bool ADD, bool MOD = false, bool MASK = false>
__device__ __forceinline__ void epi_chunk_to_smem(const float (&rr)[32], const float4 (&av)[8], const float* bias_nb,
                                                  int nvalid4, float nz, float alpha, float slope, float gain,
                                                  uint32_t dst, int lane, const float* cs_nb = nullptr,
                                                  const float* s2_nb = nullptr, uint32_t dst2 = 0) {
  // MOD (the generator's shared-weight modulated convolution): the accumulator is scaled per output channel by the
  // sample's demodulation factor cs[n] before the noise / bias terms, and a second box out * s2[n] (the next layer's
  // modulated input) is staged next to the first.
  const float a2 = (!NB && !ACT && !ADD && !MOD) ? alpha * gain : alpha;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v[4] = {rr[4 * j + 0] * a2, rr[4 * j + 1] * a2, rr[4 * j + 2] * a2, rr[4 * j + 3] * a2};
    if (MOD) {
      if (cs_nb != nullptr && j < nvalid4) {
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs_nb) + j);
        v[0] *= c4.x; v[1] *= c4.y; v[2] *= c4.z; v[3] *= c4.w;
      }
    }
    if (NB) {
      float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias_nb != nullptr && j < nvalid4) bz = __ldg(reinterpret_cast<const float4*>(bias_nb) + j);
      v[0] += nz + bz.x; v[1] += nz + bz.y; v[2] += nz + bz.z; v[3] += nz + bz.w;
    }
    if (ACT) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * slope;
    }
    if (ADD && !MASK) { v[0] += av[j].x; v[1] += av[j].y; v[2] += av[j].z; v[3] += av[j].w; }
    if (ADD && MASK) {          // `av` = the producing layer's activation output: leaky-ReLU mask of its backward
      v[0] *= av[j].x > 0.f ? 1.f : slope; v[1] *= av[j].y > 0.f ? 1.f : slope;
      v[2] *= av[j].z > 0.f ? 1.f : slope; v[3] *= av[j].w > 0.f ? 1.f : slope;
    }
    if (NB || ACT || ADD || MOD) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] *= gain;
    }
    const uint32_t sw = (uint32_t)(j ^ (lane & 7)) << 4;
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + sw), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    if (MOD) {
      if (dst2 != 0) {
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < nvalid4) s4 = __ldg(reinterpret_cast<const float4*>(s2_nb) + j);
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst2 + sw), "f"(v[0] * s4.x), "f"(v[1] * s4.y), "f"(v[2] * s4.z), "f"(v[3] * s4.w) : "memory");
      }
    }
  }
}

// Epilogue of one 128-pixel x BN accumulator (one TMEM lane quarter per warp), three paths: (1) TMA store — TMEM ->
// registers -> fused output transform -> swizzled staging box -> one bulk tensor store per 32-pixel x 32-channel chunk;
// (2) outputs the store map cannot describe — per-warp 32x33 shared-memory transpose -> 128-bit stores (8 lanes cover
// one pixel's 128 contiguous bytes); (3) generic scalar stores (channel counts below 32, unaligned views).
// `release`: after the last TMEM read arrive on `release_bar` (hands the accumulator back to the MMA warp).
template <int BN, bool REMOTE>
__device__ __forceinline__ void pix_epilogue(const TcPixParams& p, float* stg, uint32_t tlane, int q, int lane, int b, int y0,
                                             int x0, int n0, int sub, float nw, bool release, uint32_t release_bar,
                                             const CUtensorMap* tm_out = nullptr, uint32_t stage_smem = 0,
                                             const CUtensorMap* tm_add = nullptr, uint32_t add_bar = 0,
                                             uint32_t* add_phase = nullptr, const CUtensorMap* tm_out2 = nullptr) {
  constexpr int CW = BN < 32 ? BN : 32;
  const int Wt = 1 << p.wt_log2;
      if (tm_out != nullptr && p.tma_store && CW == 32) {
        // TMA-store path: thread = pixel; the fused transform is applied in registers, the 32-channel run of each pixel
        // goes to a 128-byte row of a swizzled staging box (conflict-free 128-bit stores) and ONE bulk tensor store per
        // warp and chunk writes the 32-pixel x 32-channel box (tile edges and channel tails are clipped by the map).
        const int m0 = sub * 128 + q * 32;
        const int m = m0 + lane;
        const int y = y0 + (m >> p.wt_log2), x = x0 + (m & (Wt - 1));
        const bool valid = (y < p.PH) && (x < p.PW);
        const float nz = (valid && p.ep.noise) ? nw * __ldg(p.ep.noise + (int64_t)b * p.ep.noise_sb + (int64_t)y * p.PW + x) : 0.f;
        const int by = y0 + (m0 >> p.wt_log2), bx = x0 + (m0 & (Wt - 1));
        int buf = 0;
        // residual operand (ResNetBlock / NonLocalBlock join): its 32 x 32 box is fetched by TMA into the second
        // staging buffer while the accumulator chunk is read, so the output uses a single buffer in that mode
        const bool with_add = tm_add != nullptr && p.ep.add != nullptr;
        const bool has_bias = p.ep.bias != nullptr, has_nb = has_bias || p.ep.noise != nullptr, has_act = p.ep.act != 0;
        const float ep_alpha = p.alpha, ep_slope = p.ep.slope, ep_gain = p.ep.gain;
        // modulated form (never combined with a residual operand, checked on the host): per-sample channel scales and
        // the second output; both staging boxes are used per chunk
        const bool has_mod = p.ep.cscale != nullptr || p.ep.out2 != nullptr;
        const bool has_out2 = p.ep.out2 != nullptr && tm_out2 != nullptr;
        const float* cs_b = p.ep.cscale ? p.ep.cscale + (int64_t)b * p.ep.cscale_sb : nullptr;
        const float* s2_b = p.ep.out2 ? p.ep.out2_scale + (int64_t)b * p.ep.out2_scale_sb : nullptr;
        if (with_add && n0 < p.N) {
          __syncwarp();                                     // every lane is done reading the previous tile's box
          if (lane == 0) {
            fence_proxy_async_smem();
            mbar_expect_tx(add_bar, 4096);
            tma_load_4d(stage_smem + 4096, tm_add, add_bar, n0, bx, by, b);
          }
        }
#pragma unroll 1
        for (int cc = 0; cc < BN; cc += 32) {
          float rr[32];
          tmem_ld_32x32(tlane + cc, rr);
          tmem_ld_wait();
          if (release && cc + 32 >= BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (REMOTE) mbar_arrive_cluster(release_bar); else mbar_arrive(release_bar); }
          }
          const int nb = n0 + cc;
          if (nb < p.N) {
            float4 av[8];
            if (with_add) {
              mbar_wait(add_bar, *add_phase & 1u);
              *add_phase += 1;
              const uint32_t src = stage_smem + 4096 + lane * 128;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t a = src + ((uint32_t)(j ^ (lane & 7)) << 4);
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(av[j].x), "=f"(av[j].y), "=f"(av[j].z), "=f"(av[j].w) : "r"(a) : "memory");
              }
              __syncwarp();
              // the landing buffer is free again: fetch the next chunk's residual box behind the math and the store
              if (lane == 0 && cc + 32 < BN && nb + 32 < p.N) {
                fence_proxy_async_smem();
                mbar_expect_tx(add_bar, 4096);
                tma_load_4d(stage_smem + 4096, tm_add, add_bar, nb + 32, bx, by, b);
              }
              if (lane == 0) tma_store_wait_read<0>();      // single output buffer in this mode
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) av[j] = make_float4(0.f, 0.f, 0.f, 0.f);   // (dead in the ADD = false variants)
              if (lane == 0) {
                // both boxes of the previous chunk have left their buffers / a single-chunk tile (BN = 32) starts every
                // call on buffer 0, which the PREVIOUS call's store may still be reading
                if (has_out2 || BN == 32) tma_store_wait_read<0>();
                else tma_store_wait_read<1>();              // the box stored two chunks ago has left this buffer
              }
            }
            __syncwarp();
            const uint32_t dst = stage_smem + buf * 4096 + lane * 128;
            const float* bias_nb = has_bias ? p.ep.bias + nb : nullptr;
            const int nvalid4 = (p.N - nb + 3) >> 2;
            if (has_mod)
              epi_chunk_to_smem<true, true, false, true>(rr, av, bias_nb, nvalid4, nz, ep_alpha, has_act ? ep_slope : 1.f, ep_gain,
                                                         dst, lane, cs_b ? cs_b + nb : nullptr, s2_b ? s2_b + nb : nullptr,
                                                         has_out2 ? stage_smem + 4096 + lane * 128 : 0u);
            else if (!has_nb && !has_act && !with_add)
              epi_chunk_to_smem<false, false, false>(rr, av, bias_nb, nvalid4, nz, ep_alpha, ep_slope, ep_gain, dst, lane);
            else if (has_nb && has_act && !with_add)
              epi_chunk_to_smem<true, true, false>(rr, av, bias_nb, nvalid4, nz, ep_alpha, ep_slope, ep_gain, dst, lane);
            else if (!has_nb && !has_act && with_add && p.ep.add_is_mask)
              epi_chunk_to_smem<false, false, true, false, true>(rr, av, bias_nb, nvalid4, nz, ep_alpha, ep_slope,
                                                                 valid ? ep_gain : 0.f, dst, lane);   // zeros outside the image
            else if (!has_nb && !has_act && with_add)
              epi_chunk_to_smem<false, false, true>(rr, av, bias_nb, nvalid4, nz, ep_alpha, ep_slope, ep_gain, dst, lane);
            else if (has_act)
              epi_chunk_to_smem<true, true, true>(rr, av, bias_nb, nvalid4, nz, ep_alpha, ep_slope, ep_gain, dst, lane);
            else
              epi_chunk_to_smem<true, false, true>(rr, av, bias_nb, nvalid4, nz, ep_alpha, ep_slope, ep_gain, dst, lane);
            if (!(p.debug & 2)) fence_proxy_async_smem();
            __syncwarp();
            if (p.ep.colsum != nullptr) {
              // bias gradient: lane c sums column c of the staged 32-pixel x 32-channel box (bank-conflict free: for a
              // fixed row the swizzled 16-byte groups of the 32 lanes are a permutation; rows of pixels outside the image
              // were staged as zeros) and adds it to the row of the partial-sum matrix that this (CTA, warp) alone owns:
              // a fire-and-forget reduction in program order — deterministic, and nothing waits for it
              const uint32_t col = stage_smem + buf * 4096 + ((uint32_t)(lane & 3) << 2);
              float acc = 0.f;
#pragma unroll
              for (int r = 0; r < 32; ++r) {
                float t;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(col + r * 128 + ((uint32_t)((lane >> 2) ^ (r & 7)) << 4)) : "memory");
                acc += t;
              }
              if (nb + lane < p.N) atomicAdd(p.ep.colsum + ((int64_t)blockIdx.x * 4 + q) * p.N + nb + lane, acc);
            }
            if (lane == 0 && !(p.debug & 1)) {
              tma_store_4d(tm_out, stage_smem + buf * 4096, nb, bx, by, b);
              if (has_out2) tma_store_4d(tm_out2, stage_smem + 4096, nb, bx, by, b);
              tma_store_commit();
            }
            if (!with_add && !has_out2) buf ^= 1;
          }
        }
      } else if (p.vec_store && CW == 32) {
        // coalesced path: after the transpose lane l owns channels (l & 7) * 4 .. + 3 of pixels i * 4 + (l >> 3)
        const int psub = lane >> 3, ch4 = (lane & 7) * 4;
        int64_t poff[8];
        float pnz[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = sub * 128 + q * 32 + i * 4 + psub;
          const int y = y0 + (m >> p.wt_log2), x = x0 + (m & (Wt - 1));
          const bool valid = (y < p.PH) && (x < p.PW);
          poff[i] = valid ? (int64_t)b * p.os.sb + (int64_t)(y * p.out_my + p.out_oy) * p.os.sy +
                                (int64_t)(x * p.out_mx + p.out_ox) * p.os.sx
                          : -1;
          pnz[i] = (valid && p.ep.noise) ? nw * __ldg(p.ep.noise + (int64_t)b * p.ep.noise_sb + (int64_t)y * p.PW + x) : 0.f;
        }
#pragma unroll 1
        for (int cc = 0; cc < BN; cc += 32) {
          // residual operand of this chunk: all eight loads in flight before the TMEM read and the transpose
          float4 av[8];
          const bool nok = (n0 + cc + ch4) < p.N;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            av[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.ep.add && nok && poff[i] >= 0) av[i] = __ldg(reinterpret_cast<const float4*>(p.ep.add + poff[i] + n0 + cc + ch4));
          }
          float rr[32];
          tmem_ld_32x32(tlane + cc, rr);
          tmem_ld_wait();
          if (release && cc + 32 >= BN) {
            // last TMEM read of this accumulator: hand it back to the MMA warp before the stores
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (REMOTE) mbar_arrive_cluster(release_bar); else mbar_arrive(release_bar); }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = rr[j];
          __syncwarp();
          const int n = n0 + cc + ch4;
          if (n < p.N) {
            float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.ep.bias) bz = __ldg(reinterpret_cast<const float4*>(p.ep.bias + n));
            float4 cs = make_float4(p.alpha, p.alpha, p.alpha, p.alpha);
            if (p.ep.cscale) {
              const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.ep.cscale + (int64_t)b * p.ep.cscale_sb + n));
              cs.x *= c4.x; cs.y *= c4.y; cs.z *= c4.z; cs.w *= c4.w;
            }
            float4 s2 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.ep.out2) s2 = __ldg(reinterpret_cast<const float4*>(p.ep.out2_scale + (int64_t)b * p.ep.out2_scale_sb + n));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (poff[i] >= 0) {
                const float* sp = stg + (i * 4 + psub) * 33 + ch4;
                const int64_t off = poff[i] + n;
                float4 o;
                o.x = apply_epilogue(p.ep, cs.x * sp[0], bz.x, pnz[i], av[i].x);
                o.y = apply_epilogue(p.ep, cs.y * sp[1], bz.y, pnz[i], av[i].y);
                o.z = apply_epilogue(p.ep, cs.z * sp[2], bz.z, pnz[i], av[i].z);
                o.w = apply_epilogue(p.ep, cs.w * sp[3], bz.w, pnz[i], av[i].w);
                *reinterpret_cast<float4*>(p.out + off) = o;
                if (p.ep.out2)
                  *reinterpret_cast<float4*>(p.ep.out2 + off) = make_float4(o.x * s2.x, o.y * s2.y, o.z * s2.z, o.w * s2.w);
              }
            }
          }
          __syncwarp();
        }
      } else {
        // generic path (N tails, N < 32, scattered / non-NHWC outputs): thread = pixel, scalar stores
        const int m = sub * 128 + q * 32 + lane;
        const int y = y0 + (m >> p.wt_log2), x = x0 + (m & (Wt - 1));
        const bool valid = (y < p.PH) && (x < p.PW);
        const int64_t obase = (int64_t)b * p.os.sb + (int64_t)(y * p.out_my + p.out_oy) * p.os.sy +
                              (int64_t)(x * p.out_mx + p.out_ox) * p.os.sx;
        const float nz = (valid && p.ep.noise) ? nw * __ldg(p.ep.noise + (int64_t)b * p.ep.noise_sb + (int64_t)y * p.PW + x) : 0.f;
#pragma unroll 1
        for (int cc = 0; cc < BN; cc += CW) {
          float rr[32];
          if (CW == 32) {
            tmem_ld_32x32(tlane + cc, rr);
          } else {
            float r16[16];
            tmem_ld_32x16(tlane + cc, r16);
#pragma unroll
            for (int j = 0; j < 16; ++j) rr[j] = r16[j];
          }
          tmem_ld_wait();
          if (release && cc + CW >= BN) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (REMOTE) mbar_arrive_cluster(release_bar); else mbar_arrive(release_bar); }
          }
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            const int n = n0 + cc + j;
            if (valid && n < p.N) {
              const int64_t off = obase + (int64_t)n * p.os.sc;
              const float bn = p.ep.bias ? __ldg(p.ep.bias + n) : 0.f;
              const float av = p.ep.add ? __ldg(p.ep.add + off) : 0.f;
              const float cs = p.ep.cscale ? __ldg(p.ep.cscale + (int64_t)b * p.ep.cscale_sb + n) : 1.f;
              const float o1 = apply_epilogue(p.ep, p.alpha * rr[j] * cs, bn, nz, av);
              p.out[off] = o1;
              if (p.ep.out2) p.ep.out2[off] = o1 * __ldg(p.ep.out2_scale + (int64_t)b * p.ep.out2_scale_sb + n);
            }
          }
        }
      }
}

// Persistent, warp-specialised: CTA c works on tiles c, c + gridDim.x, ...  (n-tile fastest, so neighbouring CTAs
// share the activation tile in L2).  Two TMEM accumulators: the epilogue warps drain tile i (TMEM -> registers ->
// fused epilogue -> swizzled staging box -> bulk tensor store; see pix_epilogue) while the MMA warp already
// accumulates tile i + 1 and the TMA warp runs ahead through the shared-memory ring.
template <int BN, int STAGES, int MT>
__global__ void __launch_bounds__(192, 1)
tc_pixgemm_kernel(const __grid_constant__ TMapSet tmAs, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAdd,
                  const __grid_constant__ CUtensorMap tmOut2, const TcPixParams p) {
  // MT = 128-pixel sub-tiles per CTA tile: narrow-N layers (BN <= 128) take two, which halves the weight bytes and
  // TMA boxes per FLOP (one A box of 256 pixels, one B box, two MMAs per k-step sharing the B descriptor).
  constexpr uint32_t A_BYTES = MT * 128 * 32 * 4;
  constexpr uint32_t B_BYTES = BN * 32 * 4;
  constexpr uint32_t ACC_COLS = MT * BN;               // one accumulator set
  constexpr uint32_t TMEM_COLS = (2 * ACC_COLS) < 32 ? 32 : 2 * ACC_COLS;
  static_assert(2 * ACC_COLS <= 512, "TMEM columns");
  constexpr uint32_t IDESC = make_idesc_tf32(128, BN, 0, 0);
  constexpr uint32_t STG_BYTES = 4 * 2 * 4096;        // per epilogue warp: two 32 px x 128 B boxes (TMA store, 1024-byte
                                                      // aligned) or one 32 x 33 float transpose buffer (direct stores)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * A_BYTES;
  const uint32_t sStg = sB + STAGES * B_BYTES;
  const uint32_t bars = sStg + STG_BYTES;
  const uint32_t acc_full = bars + 16 * STAGES;       // 2 barriers
  const uint32_t acc_empty = acc_full + 16;           // 2 barriers
  const uint32_t tmem_slot = acc_empty + 16;
  const uint32_t add_bars = tmem_slot + 16;           // one mbarrier per epilogue warp (residual boxes)
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* stg_all = reinterpret_cast<float*>(smem_raw + (sStg - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wt = 1 << p.wt_log2, Ht = (128 * MT) >> p.wt_log2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);             // full
      mbar_init(bars + 8 * (STAGES + s), 1);  // empty
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full + 8 * a, 1);
      mbar_init(acc_empty + 8 * a, 4);        // one arrival per epilogue warp
    }
    for (int w = 0; w < 4; ++w) mbar_init(add_bars + 8 * w, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmAs.m[0]);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmOut);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // producer: the whole warp runs the loop converged (tile / tap / chunk counters stay in uniform registers, no
    // division per k-iteration) and one elected lane issues the two bulk tensor loads of a stage
    uint32_t s = 0, ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int r = tile;
      const int grp = r % p.ksplit; r /= p.ksplit;
      const int nt = r % p.n_tiles; r /= p.n_tiles;
      const int tx = r % p.tiles_x; r /= p.tiles_x;
      const int ty = r % p.tiles_y;
      const int b = r / p.tiles_y;
      const int y0 = ty * Ht, x0 = tx * Wt, n0 = nt * BN;
      const int bw = p.w_per_sample ? b : 0;
      const int t0 = grp * p.tpg, t1 = min(p.ntaps, t0 + p.tpg);
      for (int t = t0; t < t1; ++t) {
        const int xx = x0 + p.tap_dx[t], yy = y0 + p.tap_dy[t];
        int view = 0, ca = 0;
        for (int cc = 0; cc < p.cchunks; ++cc) {
          mbar_wait(bars + 8 * (STAGES + s), ph ^ 1u);
          if (elect_one()) {
            const uint32_t full = bars + 8 * s;
            mbar_expect_tx(full, A_BYTES + B_BYTES);
            tma_load_4d(sA + s * A_BYTES, &tmAs.m[view], full, ca * 32, xx, yy, b);
            tma_load_4d(sB + s * B_BYTES, &tmB, full, cc * 32, n0, t, bw);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
          if (++ca == p.cpv[view]) { ca = 0; ++view; }
        }
      }
    }
  } else if (warp == 1) {
    // MMA issuer: converged warp, one elected lane issues the four K = 8 MMAs of a stage and the commits
    constexpr uint64_t DESC_HI = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(SWZ_128B & 7) << 61);
    uint32_t s = 0, ph = 0, lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const uint32_t a = lt & 1u;
      mbar_wait(acc_empty + 8 * a, ((lt >> 1) & 1u) ^ 1u);      // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * ACC_COLS;
      const int kiters = min(p.ntaps - (tile % p.ksplit) * p.tpg, p.tpg) * p.cchunks;   // this tile's tap group
      for (int k = 0; k < kiters; ++k) {
        mbar_wait(bars + 8 * s, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_lo = ((sA + s * A_BYTES) >> 4) & 0x3FFF, b_lo = ((sB + s * B_BYTES) >> 4) & 0x3FFF;
          // (sub-tile outer: consecutive MMAs accumulate into the same TMEM columns)
#pragma unroll
          for (int sub = 0; sub < MT; ++sub) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t bd = DESC_HI | (uint64_t)(b_lo + kk * 2);
              const uint64_t ad = DESC_HI | (uint64_t)(a_lo + sub * (128 * 32 * 4 / 16) + kk * 2);
              mma_tf32(d_tmem + sub * BN, ad, bd, IDESC, (k > 0 || kk > 0) ? 1u : 0u);
            }
          }
          mma_commit(bars + 8 * (STAGES + s));  // smem stage reusable once these MMAs retire
          if (k == kiters - 1) mma_commit(acc_full + 8 * a);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    const int q = warp & 3;                             // TMEM lane quarter this warp may read
    float* stg = stg_all + q * (2 * 4096 / 4);
    const uint32_t stage_smem = sStg + q * (2 * 4096);
    uint32_t add_phase = 0;
    const float nw = p.ep.noise ? __ldg(p.ep.noise_w) : 0.f;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      int r = tile;
      const int grp = r % p.ksplit; r /= p.ksplit;
      const int nt = r % p.n_tiles; r /= p.n_tiles;
      const int tx = r % p.tiles_x; r /= p.tiles_x;
      const int ty = r % p.tiles_y;
      const int b = r / p.tiles_y + grp * p.bsz;          // tap split: group g writes its partial sums as samples g * bsz + b
      const int y0 = ty * Ht, x0 = tx * Wt, n0 = nt * BN;
      const uint32_t a = lt & 1u;
      mbar_wait(acc_full + 8 * a, (lt >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub) {
      const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + a * ACC_COLS + sub * BN;
      const bool last_sub = sub == MT - 1;

      pix_epilogue<BN, false>(p, stg, tlane, q, lane, b, y0, x0, n0, sub, nw, last_sub, acc_empty + 8 * a, &tmOut, stage_smem,
                              &tmAdd, add_bars + 8 * q, &add_phase, &tmOut2);
      }  // sub
    }
    if (p.tma_store && lane == 0) tma_store_wait_read<0>();   // staging buffers must outlive their bulk stores
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// PixGemm with activation reuse across the taps of a filter row (3x3 / stride 1, image rows of >= 65 pixels, N <= 128).
//
// In the kernel above every (tap, 32-channel chunk) step loads its own 128-pixel activation box, so a 3x3 layer pulls
// each activation 9 times from L2; for N <= 128 that load path (96 B/clk per SM against 512 cycles of MMA per 48 KB
// stage) is what bounds the layer (tensor pipe 69 % active).  Here ONE box per (filter row,
// chunk) covers the 128 + kw - 1 pixels all kw taps of that row read: the taps are the same shared-memory image entered
// one pixel row (128 bytes) later.  That works because the 128-byte swizzle of both TMA and the UMMA descriptors is a
// function of the shared-memory ADDRESS (bits 7..9 XOR-ed into bits 4..6): an operand may start at any 128-byte row of a
// 1024-byte-aligned image with the descriptor's base-offset field left 0 (measured: tools/umma_rowshift_probe.cu, every
// row offset 0..12 exact; base offset = row & 7 gives wrong data).  Activation bytes per MMA drop 3x, total fill traffic
// for N = 128 from 96 to 53 B/clk.  Two rings: activation boxes (2 stages, one per filter row and chunk) and weight tiles
// (one per tap and chunk).  Tiles are MT image-row segments of 128 pixels (the box has MT rows); everything else —
// persistent tile loop, two TMEM accumulators, epilogue — is the kernel above.
// ------------------------------------------------------------------------------------------------
template <int BN, int MT, int KW>
__global__ void __launch_bounds__(192, 1)
tc_pixgemm_rows_kernel(const __grid_constant__ TMapSet tmAs, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAdd,
                       const __grid_constant__ CUtensorMap tmOut2, const TcPixParams p) {
  constexpr int WROW = 128 + KW - 1;                        // pixels per box row
  constexpr uint32_t A_BOX = (uint32_t)WROW * MT * 128;     // bytes the TMA delivers per activation box
  constexpr uint32_t A_STAGE = (A_BOX + 1023u) & ~1023u;
  constexpr uint32_t B_BYTES = BN * 32 * 4;
  constexpr int SA = 2;
  constexpr int SB = (BN >= 128) ? 6 : 8;
  constexpr uint32_t ACC_COLS = MT * BN;
  constexpr uint32_t TMEM_COLS = (2 * ACC_COLS) < 32 ? 32 : 2 * ACC_COLS;
  static_assert(2 * ACC_COLS <= 512, "TMEM columns");
  constexpr uint32_t IDESC = make_idesc_tf32(128, BN, 0, 0);
  constexpr uint32_t STG_BYTES = 4 * 2 * 4096;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + SA * A_STAGE;
  const uint32_t sStg = sB + SB * B_BYTES;
  const uint32_t barA = sStg + STG_BYTES;                 // SA full + SA empty
  const uint32_t barB = barA + 16 * SA;                   // SB full + SB empty
  const uint32_t acc_full = barB + 16 * SB;
  const uint32_t acc_empty = acc_full + 16;
  const uint32_t tmem_slot = acc_empty + 16;
  const uint32_t add_bars = tmem_slot + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* stg_all = reinterpret_cast<float*>(smem_raw + (sStg - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int Wt = 128, Ht = MT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(barA + 8 * s, 1); mbar_init(barA + 8 * (SA + s), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(barB + 8 * s, 1); mbar_init(barB + 8 * (SB + s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(acc_full + 8 * a, 1); mbar_init(acc_empty + 8 * a, 4); }
    for (int w = 0; w < 4; ++w) mbar_init(add_bars + 8 * w, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmAs.m[0]);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmOut);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int nrows = p.ntaps / KW;                         // filter rows

  if (warp == 0) {
    uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int r = tile;
      const int nt = r % p.n_tiles; r /= p.n_tiles;
      const int tx = r % p.tiles_x; r /= p.tiles_x;
      const int ty = r % p.tiles_y;
      const int b = r / p.tiles_y;
      const int y0 = ty * Ht, x0 = tx * Wt, n0 = nt * BN;
      const int bw = p.w_per_sample ? b : 0;
      for (int fr = 0; fr < nrows; ++fr) {
        const int t0 = fr * KW;
        // the box starts at the leftmost tap of the row (forward: the first one, dgrad: the last one)
        const int dx_lo = p.tap_dx[t0] < p.tap_dx[t0 + KW - 1] ? p.tap_dx[t0] : p.tap_dx[t0 + KW - 1];
        const int xx = x0 + dx_lo, yy = y0 + p.tap_dy[t0];
        int view = 0, ca = 0;
        for (int cc = 0; cc < p.cchunks; ++cc) {
          mbar_wait(barA + 8 * (SA + sa), pha ^ 1u);
          if (elect_one()) {
            const uint32_t full = barA + 8 * sa;
            mbar_expect_tx(full, A_BOX);
            tma_load_4d(sA + sa * A_STAGE, &tmAs.m[view], full, ca * 32, xx, yy, b);
          }
          __syncwarp();
          if (++sa == SA) { sa = 0; pha ^= 1u; }
#pragma unroll
          for (int j = 0; j < KW; ++j) {
            mbar_wait(barB + 8 * (SB + sb), phb ^ 1u);
            if (elect_one()) {
              const uint32_t full = barB + 8 * sb;
              mbar_expect_tx(full, B_BYTES);
              tma_load_4d(sB + sb * B_BYTES, &tmB, full, cc * 32, n0, t0 + j, bw);
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; phb ^= 1u; }
          }
          if (++ca == p.cpv[view]) { ca = 0; ++view; }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint64_t DESC_HI = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(SWZ_128B & 7) << 61);
    uint32_t sa = 0, pha = 0, sb = 0, phb = 0, lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      const uint32_t a = lt & 1u;
      mbar_wait(acc_empty + 8 * a, ((lt >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + a * ACC_COLS;
      for (int fr = 0; fr < nrows; ++fr) {
        const int t0 = fr * KW;
        const int dx_lo = p.tap_dx[t0] < p.tap_dx[t0 + KW - 1] ? p.tap_dx[t0] : p.tap_dx[t0 + KW - 1];
        for (int cc = 0; cc < p.cchunks; ++cc) {
          const bool first = fr == 0 && cc == 0, last = fr == nrows - 1 && cc == p.cchunks - 1;
          mbar_wait(barA + 8 * sa, pha);
          tc_fence_after();
          const uint32_t a_lo = ((sA + sa * A_STAGE) >> 4) & 0x3FFF;
#pragma unroll
          for (int j = 0; j < KW; ++j) {
            mbar_wait(barB + 8 * sb, phb);
            tc_fence_after();
            // tap j of this filter row = the same box entered (dx_j - dx_lo) pixel rows (128 bytes each) later
            const uint32_t row_off = (uint32_t)(p.tap_dx[t0 + j] - dx_lo);
            if (elect_one()) {
              const uint32_t b_lo = ((sB + sb * B_BYTES) >> 4) & 0x3FFF;
#pragma unroll
              for (int sub = 0; sub < MT; ++sub) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  const uint64_t bd = DESC_HI | (uint64_t)(b_lo + kk * 2);
                  const uint64_t ad = DESC_HI | (uint64_t)(a_lo + (sub * WROW + row_off) * (128 / 16) + kk * 2);
                  mma_tf32(d_tmem + sub * BN, ad, bd, IDESC, (!first || j > 0 || kk > 0) ? 1u : 0u);
                }
              }
              mma_commit(barB + 8 * (SB + sb));
              if (j == KW - 1) {
                mma_commit(barA + 8 * (SA + sa));
                if (last) mma_commit(acc_full + 8 * a);
              }
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; phb ^= 1u; }
          }
          if (++sa == SA) { sa = 0; pha ^= 1u; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    float* stg = stg_all + q * (2 * 4096 / 4);
    const uint32_t stage_smem = sStg + q * (2 * 4096);
    uint32_t add_phase = 0;
    const float nw = p.ep.noise ? __ldg(p.ep.noise_w) : 0.f;
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++lt) {
      int r = tile;
      const int nt = r % p.n_tiles; r /= p.n_tiles;
      const int tx = r % p.tiles_x; r /= p.tiles_x;
      const int ty = r % p.tiles_y;
      const int b = r / p.tiles_y;
      const int y0 = ty * Ht, x0 = tx * Wt, n0 = nt * BN;
      const uint32_t a = lt & 1u;
      mbar_wait(acc_full + 8 * a, (lt >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub) {
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + a * ACC_COLS + sub * BN;
        pix_epilogue<BN, false>(p, stg, tlane, q, lane, b, y0, x0, n0, sub, nw, sub == MT - 1, acc_empty + 8 * a, &tmOut,
                                stage_smem, &tmAdd, add_bars + 8 * q, &add_phase, &tmOut2);
      }
    }
    if (p.tma_store && lane == 0) tma_store_wait_read<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// PixGemm on CTA pairs (N > 128): tcgen05.mma.cta_group::2, M = 256 = the pair's two 128-pixel tiles, N = 256.
// Each CTA loads its own activation tile (16 KB) and HALF of the weight tile (128 of the 256 rows, 16 KB) per stage; the
// tensor core of the pair reads both halves, so the weight bytes and TMA rows per FLOP halve and the 32 KB stages leave
// room for a 6-deep ring.  Protocol: both producers signal the LEADER's full barrier (cp.async.bulk.tensor.cta_group::2);
// the leader's MMA thread issues for the pair and releases a stage with a multicast commit onto both CTAs' empty
// barriers; accumulators (2 x 256 columns in each CTA's TMEM) are published the same way, and every epilogue warp of
// both CTAs arrives on the leader's acc_empty barrier when it has drained its lane quarter.
// ------------------------------------------------------------------------------------------------
template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
tc_pixgemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAdd,
                   const __grid_constant__ CUtensorMap tmOut2, const TcPixParams p) {
  static_assert(BN == 256 || BN == 128, "pair tiles are 256 pixels x 256 or 128 channels");
  constexpr uint32_t A_BYTES = 128 * 32 * 4;
  constexpr uint32_t B_BYTES = (BN / 2) * 32 * 4;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  constexpr uint32_t IDESC = make_idesc_tf32(256, BN, 0, 0);
  constexpr uint32_t STG_BYTES = 4 * 2 * 4096;        // per epilogue warp: two 32-pixel x 128-byte staging boxes

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * A_BYTES;
  const uint32_t sStg = sB + STAGES * B_BYTES;
  const uint32_t bars = sStg + STG_BYTES;
  const uint32_t acc_full = bars + 16 * STAGES;
  const uint32_t acc_empty = acc_full + 16;
  const uint32_t tmem_slot = acc_empty + 16;
  const uint32_t add_bars = tmem_slot + 16;           // one mbarrier per epilogue warp (residual boxes)
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* stg_all = reinterpret_cast<float*>(smem_raw + (sStg - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int Wt = 1 << p.wt_log2, Ht = 128 >> p.wt_log2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);             // full  (used in the leader: its producer's arrive + both CTAs' bytes)
      mbar_init(bars + 8 * (STAGES + s), 1);  // empty (one multicast commit per phase, in each CTA)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full + 8 * a, 1);         // multicast commit, in each CTA
      mbar_init(acc_empty + 8 * a, 8);        // used in the leader: 4 epilogue warps x 2 CTAs
    }
    for (int w = 0; w < 4; ++w) mbar_init(add_bars + 8 * w, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmOut);
  }
  __syncthreads();
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int kiters = p.ntaps * p.cchunks;
  const int npairs = gridDim.x >> 1;
  const int pair = blockIdx.x >> 1;
  const int ptiles = p.tiles_x * p.tiles_y;          // 128-pixel tiles per sample
  const int ppairs = (ptiles + 1) >> 1;              // pixel-tile pairs per sample (the last one may be half empty)

  if (warp == 0) {
    // producers of both CTAs: converged warp, one elected lane issues (see tc_pixgemm_kernel)
    const uint32_t full_leader = mapa_cluster(bars, 0);
    uint32_t s = 0, ph = 0;
    for (int tile = pair; tile < p.total_tiles; tile += npairs) {
      int r = tile;
      const int nt = r % p.n_tiles; r /= p.n_tiles;
      const int pp = r % ppairs;
      const int b = r / ppairs;
      const int pt = 2 * pp + (int)rank;           // >= ptiles: out of the image, TMA fills zeros, nothing is stored
      const int ty = pt / p.tiles_x, tx = pt - ty * p.tiles_x;
      const int y0 = ty * Ht, x0 = tx * Wt, n0 = nt * BN + (int)rank * (BN / 2);
      const int bw = p.w_per_sample ? b : 0;
      for (int t = 0; t < p.ntaps; ++t) {
        const int xx = x0 + p.tap_dx[t], yy = y0 + p.tap_dy[t];
        for (int cc = 0; cc < p.cchunks; ++cc) {
          mbar_wait(bars + 8 * (STAGES + s), ph ^ 1u);
          if (elect_one()) {
            if (leader) mbar_expect_tx(bars + 8 * s, 2 * (A_BYTES + B_BYTES));
            tma_load_4d_2cta(sA + s * A_BYTES, &tmA, full_leader + 8 * s, cc * 32, xx, yy, b);
            tma_load_4d_2cta(sB + s * B_BYTES, &tmB, full_leader + 8 * s, cc * 32, n0, t, bw);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint64_t DESC_HI = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(SWZ_128B & 7) << 61);
      uint32_t s = 0, ph = 0, lt = 0;
      for (int tile = pair; tile < p.total_tiles; tile += npairs, ++lt) {
        const uint32_t a = lt & 1u;
        mbar_wait(acc_empty + 8 * a, ((lt >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * BN;
        for (int k = 0; k < kiters; ++k) {
          mbar_wait(bars + 8 * s, ph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = ((sA + s * A_BYTES) >> 4) & 0x3FFF, b_lo = ((sB + s * B_BYTES) >> 4) & 0x3FFF;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              mma_tf32_2cta(d_tmem, DESC_HI | (uint64_t)(a_lo + kk * 2), DESC_HI | (uint64_t)(b_lo + kk * 2), IDESC,
                            (k > 0 || kk > 0) ? 1u : 0u);
            mma_commit_2cta(bars + 8 * (STAGES + s), 3);
            if (k == kiters - 1) mma_commit_2cta(acc_full + 8 * a, 3);
          }
          __syncwarp();
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else {
    const int q = warp & 3;
    float* stg = stg_all + q * (2 * 4096 / 4);
    const uint32_t stage_smem = sStg + q * (2 * 4096);
    uint32_t add_phase = 0;
    const float nw = p.ep.noise ? __ldg(p.ep.noise_w) : 0.f;
    const uint32_t acc_empty_leader = mapa_cluster(acc_empty, 0);
    uint32_t lt = 0;
    for (int tile = pair; tile < p.total_tiles; tile += npairs, ++lt) {
      int r = tile;
      const int nt = r % p.n_tiles; r /= p.n_tiles;
      const int pp = r % ppairs;
      const int b = r / ppairs;
      const int pt = 2 * pp + (int)rank;
      const int ty = pt / p.tiles_x, tx = pt - ty * p.tiles_x;
      const int y0 = ty * Ht, x0 = tx * Wt, n0 = nt * BN;
      const uint32_t a = lt & 1u;
      mbar_wait(acc_full + 8 * a, (lt >> 1) & 1u);
      tc_fence_after();
      const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + a * BN;
      pix_epilogue<BN, true>(p, stg, tlane, q, lane, b, y0, x0, n0, 0, nw, true, acc_empty_leader + 8 * a, &tmOut,
                             stage_smem, &tmAdd, add_bars + 8 * q, &add_phase, &tmOut2);
    }
    if (p.tma_store && lane == 0) tma_store_wait_read<0>();   // staging buffers must outlive their bulk stores
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// Row-tap PixGemm on CTA pairs (N = 128): the two ideas combined.  A 128 x 128 x 8 TF32 MMA reads 8 KB of operands from
// shared memory per 64 cycles = 128 B/clk, the whole shared-memory bandwidth of an SM, so on a single CTA the TMA fill
// and the epilogue staging compete with the tensor core for it (ncu on tc_pixgemm_rows_kernel<128,2,3>: tensor pipe 74 %
// active with the MMA warp back-pressured and the producer idle, profiles/r2h_rows_full.txt).  With cta_group::2 each
// SM reads its own activation rows and only HALF of the weight rows (96 B/clk), and with the activation box shared by
// the three taps of a filter row the fill drops to (16.6 KB + 3 x 8 KB) per 768 cycles.  Each CTA owns MT image-row
// segments of 128 pixels per tile; the pair's M = 256 MMA covers segment `sub` of both CTAs.
// ------------------------------------------------------------------------------------------------
template <int BN, int MT, int KW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
tc_pixgemm2_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmAdd,
                        const __grid_constant__ CUtensorMap tmOut2, const TcPixParams p) {
  constexpr int WROW = 128 + KW - 1;
  constexpr uint32_t A_BOX = (uint32_t)WROW * MT * 128;
  constexpr uint32_t A_STAGE = (A_BOX + 1023u) & ~1023u;
  constexpr uint32_t B_BYTES = (BN / 2) * 32 * 4;          // this CTA's half of the weight tile
  constexpr int SA = 3;
  constexpr int SB = 8;
  constexpr uint32_t ACC_COLS = MT * BN;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
  constexpr uint32_t IDESC = make_idesc_tf32(256, BN, 0, 0);
  constexpr uint32_t STG_BYTES = 4 * 2 * 4096;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + SA * A_STAGE;
  const uint32_t sStg = sB + SB * B_BYTES;
  const uint32_t barA = sStg + STG_BYTES;
  const uint32_t barB = barA + 16 * SA;
  const uint32_t acc_full = barB + 16 * SB;
  const uint32_t acc_empty = acc_full + 16;
  const uint32_t tmem_slot = acc_empty + 16;
  const uint32_t add_bars = tmem_slot + 16;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));
  float* stg_all = reinterpret_cast<float*>(smem_raw + (sStg - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  constexpr int Wt = 128, Ht = MT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(barA + 8 * s, 1); mbar_init(barA + 8 * (SA + s), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(barB + 8 * s, 1); mbar_init(barB + 8 * (SB + s), 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full + 8 * a, 1);
      mbar_init(acc_empty + 8 * a, 8);
    }
    for (int w = 0; w < 4; ++w) mbar_init(add_bars + 8 * w, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_store) tma_prefetch_desc(&tmOut);
  }
  __syncthreads();
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int nrows = p.ntaps / KW;
  const int npairs = gridDim.x >> 1;
  const int pair = blockIdx.x >> 1;
  const int ptiles = p.tiles_x * p.tiles_y;          // (128 * MT)-pixel tiles per sample
  const int ppairs = (ptiles + 1) >> 1;

  if (warp == 0) {
    const uint32_t fullA_leader = mapa_cluster(barA, 0), fullB_leader = mapa_cluster(barB, 0);
    uint32_t sa = 0, pha = 0, sb = 0, phb = 0;
    for (int tile = pair; tile < p.total_tiles; tile += npairs) {
      int r = tile;
      const int nt = r % p.n_tiles; r /= p.n_tiles;
      const int pp = r % ppairs;
      const int b = r / ppairs;
      const int pt = 2 * pp + (int)rank;             // >= ptiles: out of the image, TMA fills zeros, nothing is stored
      const int ty = pt / p.tiles_x, tx = pt - ty * p.tiles_x;
      const int y0 = ty * Ht, x0 = tx * Wt, n0 = nt * BN + (int)rank * (BN / 2);
      const int bw = p.w_per_sample ? b : 0;
      for (int fr = 0; fr < nrows; ++fr) {
        const int t0 = fr * KW;
        const int dx_lo = p.tap_dx[t0] < p.tap_dx[t0 + KW - 1] ? p.tap_dx[t0] : p.tap_dx[t0 + KW - 1];
        const int xx = x0 + dx_lo, yy = y0 + p.tap_dy[t0];
        for (int cc = 0; cc < p.cchunks; ++cc) {
          mbar_wait(barA + 8 * (SA + sa), pha ^ 1u);
          if (elect_one()) {
            if (leader) mbar_expect_tx(barA + 8 * sa, 2 * A_BOX);
            tma_load_4d_2cta(sA + sa * A_STAGE, &tmA, fullA_leader + 8 * sa, cc * 32, xx, yy, b);
          }
          __syncwarp();
          if (++sa == SA) { sa = 0; pha ^= 1u; }
#pragma unroll
          for (int j = 0; j < KW; ++j) {
            mbar_wait(barB + 8 * (SB + sb), phb ^ 1u);
            if (elect_one()) {
              if (leader) mbar_expect_tx(barB + 8 * sb, 2 * B_BYTES);
              tma_load_4d_2cta(sB + sb * B_BYTES, &tmB, fullB_leader + 8 * sb, cc * 32, n0, t0 + j, bw);
            }
            __syncwarp();
            if (++sb == SB) { sb = 0; phb ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      constexpr uint64_t DESC_HI = ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)(SWZ_128B & 7) << 61);
      uint32_t sa = 0, pha = 0, sb = 0, phb = 0, lt = 0;
      for (int tile = pair; tile < p.total_tiles; tile += npairs, ++lt) {
        const uint32_t a = lt & 1u;
        mbar_wait(acc_empty + 8 * a, ((lt >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + a * ACC_COLS;
        for (int fr = 0; fr < nrows; ++fr) {
          const int t0 = fr * KW;
          const int dx_lo = p.tap_dx[t0] < p.tap_dx[t0 + KW - 1] ? p.tap_dx[t0] : p.tap_dx[t0 + KW - 1];
          for (int cc = 0; cc < p.cchunks; ++cc) {
            const bool first = fr == 0 && cc == 0, last = fr == nrows - 1 && cc == p.cchunks - 1;
            mbar_wait(barA + 8 * sa, pha);
            tc_fence_after();
            const uint32_t a_lo = ((sA + sa * A_STAGE) >> 4) & 0x3FFF;
#pragma unroll
            for (int j = 0; j < KW; ++j) {
              mbar_wait(barB + 8 * sb, phb);
              tc_fence_after();
              const uint32_t row_off = (uint32_t)(p.tap_dx[t0 + j] - dx_lo);
              if (elect_one()) {
                const uint32_t b_lo = ((sB + sb * B_BYTES) >> 4) & 0x3FFF;
#pragma unroll
                for (int sub = 0; sub < MT; ++sub) {
#pragma unroll
                  for (int kk = 0; kk < 4; ++kk)
                    mma_tf32_2cta(d_tmem + sub * BN, DESC_HI | (uint64_t)(a_lo + (sub * WROW + row_off) * (128 / 16) + kk * 2),
                                  DESC_HI | (uint64_t)(b_lo + kk * 2), IDESC, (!first || j > 0 || kk > 0) ? 1u : 0u);
                }
                mma_commit_2cta(barB + 8 * (SB + sb), 3);
                if (j == KW - 1) {
                  mma_commit_2cta(barA + 8 * (SA + sa), 3);
                  if (last) mma_commit_2cta(acc_full + 8 * a, 3);
                }
              }
              __syncwarp();
              if (++sb == SB) { sb = 0; phb ^= 1u; }
            }
            if (++sa == SA) { sa = 0; pha ^= 1u; }
          }
        }
      }
    }
  } else {
    const int q = warp & 3;
    float* stg = stg_all + q * (2 * 4096 / 4);
    const uint32_t stage_smem = sStg + q * (2 * 4096);
    uint32_t add_phase = 0;
    const float nw = p.ep.noise ? __ldg(p.ep.noise_w) : 0.f;
    const uint32_t acc_empty_leader = mapa_cluster(acc_empty, 0);
    uint32_t lt = 0;
    for (int tile = pair; tile < p.total_tiles; tile += npairs, ++lt) {
      int r = tile;
      const int nt = r % p.n_tiles; r /= p.n_tiles;
      const int pp = r % ppairs;
      const int b = r / ppairs;
      const int pt = 2 * pp + (int)rank;
      const int ty = pt / p.tiles_x, tx = pt - ty * p.tiles_x;
      const int y0 = ty * Ht, x0 = tx * Wt, n0 = nt * BN;
      const uint32_t a = lt & 1u;
      mbar_wait(acc_full + 8 * a, (lt >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub) {
        const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16) + a * ACC_COLS + sub * BN;
        pix_epilogue<BN, true>(p, stg, tlane, q, lane, b, y0, x0, n0, sub, nw, sub == MT - 1, acc_empty_leader + 8 * a, &tmOut,
                               stage_smem, &tmAdd, add_bars + 8 * q, &add_phase, &tmOut2);
      }
    }
    if (p.tma_store && lane == 0) tma_store_wait_read<0>();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// RedGemm kernel (wgrad): both operands MN-major (BASE32B), split-K into `part`.
// ------------------------------------------------------------------------------------------------
struct TcRedParams {
  int ntaps;
  int tap_dy[kMaxTaps], tap_dx[kMaxTaps];
  int B, per_sample;
  int wk_log2;              // K chunk = (1 << wk_log2) x (32 >> wk_log2) pixels
  int chunks_x, chunks_y;
  int ctiles;               // number of BN-wide tiles along C
  int Npad, Cpad, splits;
  float* part;              // [splits][BS][ntaps][Npad][Cpad]
  int a5d, b5d;             // operand tensor map is 5-D (32 ch, W, H, B, C/32) -> one TMA box per stage for that operand
  uint32_t variant;
  uint32_t* dbg;
};

// TG filter taps per CTA (narrow C tiles): the dy tile (A) is loaded once per stage and multiplied against the TG shifted
// x tiles, each tap accumulating into its own BN TMEM columns — one A box + TG B boxes per TG MMAs groups instead of a
// full (A, B) pair per tap.
template <int BN, int STAGES, int TG>
__global__ void __launch_bounds__(192, 1)
tc_redgemm_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmI,
                  const TcRedParams p) {
  static_assert(BN % 32 == 0, "N-major TF32 operand comes in 32-channel atoms");
  constexpr uint32_t ATOM_BYTES = 32 * 32 * 4;        // 32 pixels x 32 channels
  constexpr uint32_t A_BYTES = 4 * ATOM_BYTES;        // 128 output channels
  constexpr uint32_t B_BYTES = (BN / 32) * ATOM_BYTES;
  constexpr uint32_t TMEM_COLS = (TG * BN <= 32) ? 32 : (TG * BN <= 64) ? 64 : (TG * BN <= 128) ? 128 : (TG * BN <= 256) ? 256 : 512;
  static_assert(TG * BN <= 512, "TMEM columns");
  constexpr uint32_t IDESC = make_idesc_tf32(128, BN, 1, 1);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + STAGES * A_BYTES;
  const uint32_t bars = sB + STAGES * TG * B_BYTES;
  const uint32_t acc_full = bars + 16 * STAGES;
  const uint32_t tmem_slot = acc_full + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = blockIdx.x / p.ctiles, ctile = blockIdx.x - ntile * p.ctiles;
  const int n0 = ntile * 128, c0 = ctile * BN;
  const int tgroups = (p.ntaps + TG - 1) / TG;
  const int tg0 = (blockIdx.y % tgroups) * TG;           // first tap of this CTA's group
  const int ntg = (TG == 1) ? 1 : ((p.ntaps - tg0) < TG ? (p.ntaps - tg0) : TG);
  const int bs = blockIdx.y / tgroups;      // sample index when per_sample, else 0
  const int split = blockIdx.z;
  const int Wk = 1 << p.wk_log2, Hk = 32 >> p.wk_log2;

  const int per = p.chunks_x * p.chunks_y;
  const int64_t total = (int64_t)per * (p.per_sample ? 1 : p.B);
  const int k_begin = (int)(total * split / p.splits);
  const int k_end = (int)(total * (split + 1) / p.splits);
  const int kiters = k_end - k_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (STAGES + s), 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmI);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  uint32_t* dbg = (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? p.dbg : nullptr;
  const bool soft = p.dbg != nullptr;
  if (threadIdx.x == 0) { dbg_set(dbg, 0, 0xC0FFEE02u); dbg_set(dbg, 5, tmem_base); dbg_set(dbg, 6, (uint32_t)kiters); }

  if (warp == 0) {
    // One TMA box per 32-channel atom: 4 + BN/32 boxes per stage.  They are issued by that many lanes in parallel
    // (a single thread issuing 12 bulk-tensor copies per 512-cycle MMA group was the measured limiter).
    constexpr int NBOX = 4 + BN / 32;
    // K chunk counters (sample, chunk row, chunk column) advance incrementally: no division per k-iteration
    int bl = k_begin / per;
    int yc = (k_begin - bl * per) / p.chunks_x, xc = (k_begin - bl * per) - yc * p.chunks_x;
    uint32_t s = 0, ph = 0;
    for (int it = 0; it < kiters; ++it) {
      if (!mbar_wait(bars + 8 * (STAGES + s), ph ^ 1u, soft)) { dbg_set(dbg, 4, 0x100u | (it << 12)); break; }
      const int b = p.per_sample ? bs : bl;
      const uint32_t full = bars + 8 * s;
      const bool issuer = elect_one();                  // same lane every time; keeps the 5-D loads' operands uniform
      if (issuer) mbar_expect_tx(full, A_BYTES + ntg * B_BYTES);
      __syncwarp();
      // 5-D maps: all channel atoms of an operand in one box ([atom][32 px][32 ch] is exactly the MN-major atom
      // layout); otherwise one 4-D box per atom, issued by consecutive lanes
      if (p.a5d) {
        if (issuer) tma_load_5d(sA + s * A_BYTES, &tmG, full, 0, xc * Wk, yc * Hk, b, n0 >> 5);
      } else if (lane < 4) {
        tma_load_4d(sA + s * A_BYTES + lane * ATOM_BYTES, &tmG, full, n0 + 32 * lane, xc * Wk, yc * Hk, b);
      }
#pragma unroll
      for (int g = 0; g < TG; ++g) {
        if (g < ntg) {
          if (p.b5d) {
            if (issuer)
              tma_load_5d(sB + (s * TG + g) * B_BYTES, &tmI, full, 0, xc * Wk + p.tap_dx[tg0 + g],
                          yc * Hk + p.tap_dy[tg0 + g], b, c0 >> 5);
          } else if (lane >= 4 && lane < NBOX) {
            tma_load_4d(sB + (s * TG + g) * B_BYTES + (lane - 4) * ATOM_BYTES, &tmI, full, c0 + 32 * (lane - 4),
                        xc * Wk + p.tap_dx[tg0 + g], yc * Hk + p.tap_dy[tg0 + g], b);
          }
        }
      }
      if (lane == 0) dbg_set(dbg, 1, 2 * it + 2);
      if (++s == STAGES) { s = 0; ph ^= 1u; }
      if (++xc == p.chunks_x) { xc = 0; if (++yc == p.chunks_y) { yc = 0; ++bl; } }
    }
  } else if (warp == 1) {
    // MN-major BASE32B: M/N atoms (32 channels) are ATOM_BYTES apart (LBO), K atoms (4 pixel rows of
    // 128 bytes) 512 bytes apart (SBO); one K=8 MMA covers 2 K atoms = 1024 bytes.
    // Converged warp, one elected lane issues (descriptors stay in uniform registers, see tc_pixgemm_kernel).
    const bool swap = p.variant & 1u;
    const uint32_t lbo = swap ? 512u : ATOM_BYTES;
    const uint32_t sbo = swap ? ATOM_BYTES : 512u;
    const uint64_t desc_hi = ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
                             ((uint64_t)1 << 46) | ((uint64_t)(SWZ_128B_BASE32B & 7) << 61);
    uint32_t s = 0, ph = 0;
    bool ok = true;
    for (int it = 0; it < kiters; ++it) {
      if (!mbar_wait(bars + 8 * s, ph, soft)) { if (lane == 0) dbg_set(dbg, 4, 0x200u | (it << 12)); ok = false; break; }
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_lo = ((sA + s * A_BYTES) >> 4) & 0x3FFF;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = desc_hi | (uint64_t)(a_lo + k * 64);
#pragma unroll
          for (int g = 0; g < TG; ++g) {
            if (g < ntg) {
              const uint32_t b_lo = ((sB + (s * TG + g) * B_BYTES) >> 4) & 0x3FFF;
              mma_tf32(tmem_base + g * BN, ad, desc_hi | (uint64_t)(b_lo + k * 64), IDESC, (it > 0 || k > 0) ? 1u : 0u);
            }
          }
        }
        mma_commit(bars + 8 * (STAGES + s));
        dbg_set(dbg, 2, 2 * it + 2);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1u; }
    }
    (void)ok;
    if (elect_one()) {
      mma_commit(acc_full);
      dbg_set(dbg, 2, 0x80000000u | (uint32_t)kiters);
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int n = q * 32 + lane;
    bool acc_ok = true;
    if (kiters > 0) {
      acc_ok = mbar_wait(acc_full, 0, soft);
      if (!acc_ok && threadIdx.x == 64) dbg_set(dbg, 4, 0x300u);
      tc_fence_after();
    }
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int64_t slots_y = (int64_t)(gridDim.y / tgroups) * p.ntaps;      // (sample, tap) slabs per split
#pragma unroll 1
    for (int g = 0; g < ntg; ++g) {
      const int64_t slot = (int64_t)bs * p.ntaps + tg0 + g;
      float* prow = p.part + ((((int64_t)split * slots_y + slot) * p.Npad) + n0 + n) * p.Cpad + c0;
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 32) {
        float r[32];
        if (kiters > 0 && acc_ok) {
          tmem_ld_32x32(tlane + g * BN + cc, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(prow + cc + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
    }
    if (threadIdx.x == 64) dbg_set(dbg, 3, 2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// RedGemm on CTA pairs (wide layers: N >= 256 output channels, 256-channel C tiles): tcgen05.mma.cta_group::2 with
// M = 256 = the 128 output channels of each CTA, N = 256 input channels of which each CTA loads HALF, and TWO such MMA
// units per pair and K step (two accumulators = all 512 TMEM columns) that share one operand tile:
//   tap groups (grid.y < full groups): the units are two filter taps — one dy tile, two shifted x tiles;
//   the remainder tap of an odd filter (3x3: the ninth): the units are two output-channel blocks — two dy tiles, one x
//   tile — so these pairs advance through K at the pace of all the others and every K split of the grid walks ONE front
//   through dy and x (with fewer, longer splits for the odd tap its pairs streamed both tensors a second time from
//   DRAM: 4.7 GB instead of 2.1 GB at 512->512, 256^2).
// Operand fill per SM: 48 KB per 1024 MMA cycles = 47 B/clk against 96 B/clk of the single-CTA kernel, whose 4 x 48 KB
// ring could not cover the L2 latency at that rate (ncu: tensor pipe 78 % active; this kernel: 99 %).  A pair left with a
// single unit (odd number of 256-channel blocks) takes proportionally fewer K splits so that it carries the same number
// of MMAs.  Barrier protocol as in tc_pixgemm2_kernel.
// ------------------------------------------------------------------------------------------------
template <int STAGES, int TG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
tc_redgemm2_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmI,
                   const TcRedParams p) {
  constexpr int BN = 256;
  constexpr uint32_t ATOM_BYTES = 32 * 32 * 4;
  constexpr uint32_t T_BYTES = 4 * ATOM_BYTES;        // one operand tile of a CTA: 128 channels x 32 pixels
  constexpr int SLOTS = 1 + TG;                       // tiles per stage
  constexpr uint32_t TMEM_COLS = TG * BN <= 256 ? 256 : 512;
  static_assert(TG == 1 || TG == 2, "one or two MMA units per K step");
  constexpr uint32_t IDESC = make_idesc_tf32(256, BN, 1, 1);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t sT = base;                           // [STAGES][SLOTS] tiles
  const uint32_t bars = sT + STAGES * SLOTS * T_BYTES;
  const uint32_t acc_full = bars + 16 * STAGES;
  const uint32_t tmem_slot = acc_full + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cl = blockIdx.x >> 1;
  const int npairs = p.Npad >> 8;
  const int npair = cl / p.ctiles, ctile = cl - npair * p.ctiles;
  const int c_out0 = ctile * BN;                         // the pair's input channels (accumulator columns)
  const int c_ld0 = c_out0 + (int)rank * (BN / 2);       // the half of them this CTA loads
  const int full_groups = p.ntaps / TG;                  // groups per sample: full_groups (+ 1 for the odd tap)
  const int groups = (p.ntaps + TG - 1) / TG;
  const int grp = blockIdx.y % groups;
  const int bs = blockIdx.y / groups;
  const int split = blockIdx.z;
  const int Wk = 1 << p.wk_log2, Hk = 32 >> p.wk_log2;

  // the (up to) two MMA units of this pair: unit u accumulates tap tap_u of output-channel block nblk_u
  const bool share_b = TG == 2 && grp >= full_groups;    // the odd tap: two output-channel blocks against one x tile
  int nunit = TG, tap0 = grp * TG, tap1 = grp * TG + 1, nblk0 = npair, nblk1 = npair;
  bool active = true;
  if (share_b) {
    tap1 = tap0;
    nblk0 = 2 * npair; nblk1 = 2 * npair + 1;
    active = nblk0 < npairs;
    if (nblk1 >= npairs) nunit = 1;
  }
  const int n0_0 = (nblk0 * 2 + (int)rank) * 128, n0_1 = (nblk1 * 2 + (int)rank) * 128;   // this CTA's A rows per unit

  // K range: a pair with one unit instead of TG is cut into proportionally fewer splits (equal MMA counts per CTA)
  int nsplit = p.splits;
  if (nunit < TG) nsplit = (p.splits + TG - 1) / TG;
  const int per = p.chunks_x * p.chunks_y;
  const int64_t total = (int64_t)per * (p.per_sample ? 1 : p.B);
  int k_begin = 0, kiters = 0;
  if (active && split < nsplit) {
    k_begin = (int)(total * split / nsplit);
    kiters = (int)(total * (split + 1) / nsplit) - k_begin;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);             // full  (used in the leader: its producer's arrive + both CTAs' bytes)
      mbar_init(bars + 8 * (STAGES + s), 1);  // empty (one multicast commit per phase, in each CTA)
    }
    mbar_init(acc_full, 1);                   // multicast commit, in each CTA
    fence_barrier_init();
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmI);
  }
  __syncthreads();
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // stage slots: 0 = dy tile of unit 0, 1 = x tile of unit 0, 2 = x tile of unit 1 (tap groups) / dy tile of unit 1
    const uint32_t full_leader = mapa_cluster(bars, 0);
    int bl = k_begin / per;
    int yc = (k_begin - bl * per) / p.chunks_x, xc = (k_begin - bl * per) - yc * p.chunks_x;
    uint32_t s = 0, ph = 0;
    for (int it = 0; it < kiters; ++it) {
      mbar_wait(bars + 8 * (STAGES + s), ph ^ 1u);
      const int b = p.per_sample ? bs : bl;
      if (elect_one()) {
        const uint32_t st0 = sT + s * SLOTS * T_BYTES, fb = full_leader + 8 * s;
        if (leader) mbar_expect_tx(bars + 8 * s, 2 * (1 + nunit) * T_BYTES);
        tma_load_5d_2cta(st0, &tmG, fb, 0, xc * Wk, yc * Hk, b, n0_0 >> 5);
        tma_load_5d_2cta(st0 + T_BYTES, &tmI, fb, 0, xc * Wk + p.tap_dx[tap0], yc * Hk + p.tap_dy[tap0], b, c_ld0 >> 5);
        if (TG == 2 && nunit == 2) {
          if (share_b) tma_load_5d_2cta(st0 + 2 * T_BYTES, &tmG, fb, 0, xc * Wk, yc * Hk, b, n0_1 >> 5);
          else tma_load_5d_2cta(st0 + 2 * T_BYTES, &tmI, fb, 0, xc * Wk + p.tap_dx[tap1], yc * Hk + p.tap_dy[tap1], b, c_ld0 >> 5);
        }
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; ph ^= 1u; }
      if (++xc == p.chunks_x) { xc = 0; if (++yc == p.chunks_y) { yc = 0; ++bl; } }
    }
  } else if (warp == 1) {
    if (leader && kiters > 0) {
      // MN-major BASE32B operands as in tc_redgemm_kernel: 32-channel atoms ATOM_BYTES apart (LBO), K atoms of 4 pixel
      // rows 512 bytes apart (SBO); the peer's operands sit at the same shared-memory offsets
      const uint64_t desc_hi = ((uint64_t)((ATOM_BYTES >> 4) & 0x3FFF) << 16) | ((uint64_t)((512u >> 4) & 0x3FFF) << 32) |
                               ((uint64_t)1 << 46) | ((uint64_t)(SWZ_128B_BASE32B & 7) << 61);
      const uint32_t a1_slot = share_b ? 2u : 0u, b1_slot = share_b ? 1u : 2u;
      uint32_t s = 0, ph = 0;
      for (int it = 0; it < kiters; ++it) {
        mbar_wait(bars + 8 * s, ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t t_lo = ((sT + s * SLOTS * T_BYTES) >> 4) & 0x3FFF;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
            mma_tf32_2cta(tmem_base, desc_hi | (uint64_t)(t_lo + k * 64), desc_hi | (uint64_t)(t_lo + (T_BYTES >> 4) + k * 64),
                          IDESC, acc);
            if (TG == 2 && nunit == 2)
              mma_tf32_2cta(tmem_base + BN, desc_hi | (uint64_t)(t_lo + a1_slot * (T_BYTES >> 4) + k * 64),
                            desc_hi | (uint64_t)(t_lo + b1_slot * (T_BYTES >> 4) + k * 64), IDESC, acc);
          }
          mma_commit_2cta(bars + 8 * (STAGES + s), 3);
          if (it == kiters - 1) mma_commit_2cta(acc_full, 3);
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
    }
  } else if (active) {
    const int q = warp & 3;
    const int n = q * 32 + lane;
    if (kiters > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int64_t slots_y = (int64_t)(gridDim.y / groups) * p.ntaps;
#pragma unroll 1
    for (int u = 0; u < nunit; ++u) {
      const int64_t slot = (int64_t)bs * p.ntaps + (u == 0 ? tap0 : tap1);
      const int nrow = (u == 0 ? n0_0 : n0_1) + n;
      float* prow = p.part + ((((int64_t)split * slots_y + slot) * p.Npad) + nrow) * p.Cpad + c_out0;
#pragma unroll 1
      for (int cc = 0; cc < BN; cc += 32) {
        float r[32];
        if (kiters > 0) {
          tmem_ld_32x32(tlane + u * BN + cc, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(prow + cc + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

struct RedReduceParams {
  const float* part;
  float* dw;
  int64_t dw_sb, dw_sn, dw_sc, dw_st;
  int splits, BS, ntaps, N, C, Npad, Cpad;
  int tap_wi[kMaxTaps];
  float alpha;
};

// thread = one (bs, n, c): sums the split-K partials of every tap (reads coalesced along c) and writes the taps of
// its filter element as one contiguous run (dw is [.., n, c, t] with t fastest for all callers).
__global__ void __launch_bounds__(256)
tc_red_reduce_kernel(const RedReduceParams p) {
  const int64_t total = (int64_t)p.BS * p.N * p.C;
  const int64_t slab = (int64_t)p.BS * p.ntaps * p.Npad * p.Cpad;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int64_t r = i;
    const int c = (int)(r % p.C); r /= p.C;
    const int n = (int)(r % p.N); r /= p.N;
    const int bs = (int)r;
    float* dst = p.dw + bs * p.dw_sb + n * p.dw_sn + c * p.dw_sc;
    for (int t = 0; t < p.ntaps; ++t) {
      const int64_t off = (((int64_t)bs * p.ntaps + t) * p.Npad + n) * p.Cpad + c;
      float acc = 0.f;
      for (int s = 0; s < p.splits; ++s) acc += __ldg(p.part + s * slab + off);
      dst[p.tap_wi[t] * p.dw_st] = p.alpha * acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int round_up(int a, int b) { return (a + b - 1) / b * b; }
static inline int ilog2_ceil(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

static int pick_bn(int n, int min_bn) {
  int bn = 16;
  if (n > 128) bn = 256;
  else if (n > 64) bn = 128;
  else if (n > 32) bn = 64;
  else if (n > 16) bn = 32;
  return bn < min_bn ? min_bn : bn;
}

// Column tile for a PixGemm: the widest tile the channel count allows, narrowed (down to 64) while the whole problem
// still fits in one wave of CTAs — small-pixel layers (4x4 ... 16x16 images, per-sample weights) otherwise leave most of
// the machine idle while a few CTAs walk the whole K loop.
static int pick_bn_pix(const PixGemm& g) {
  int bn = pick_bn(g.N, 16);
  int wt_log2 = ilog2_ceil(g.PW);
  if (wt_log2 > 7) wt_log2 = 7;
  const int Wt = 1 << wt_log2, Ht = 128 >> wt_log2;
  const int64_t ptiles = ceil_div(g.PW, Wt) * ceil_div(g.PH, Ht) * (int64_t)g.B;
  while (bn > 64 && ptiles * ceil_div(g.N, bn / 2) <= num_sms()) bn /= 2;
  return bn;
}

static bool view_tma_ok(const void* base, const View4& v) {
  return al16(base) && v.sc == 1 && (v.sx % 4 == 0) && (v.sy % 4 == 0) && (v.sb % 4 == 0);
}

bool tc_pixgemm_supported(const PixGemm& g) {
  if (!tc_available()) return false;
  if (g.ntaps <= 0 || g.ntaps > kMaxTaps || g.Cr <= 0 || g.N <= 0 || g.B <= 0 || g.B > 65535) return false;
  if (g.my != 1 || g.mx != 1) return false;
  if (g.PW < 1 || g.PH < 1) return false;
  if (g.nphase != 0) {
    if (g.nphase != 4 || g.Cr % (32 * 4) != 0) return false;
    for (int v = 0; v < 4; ++v)
      if (!view_tma_ok(g.in_ph[v], g.is) || g.IH_ph[v] < 1 || g.IW_ph[v] < 1) return false;
    return true;
  }
  if (g.nsrc != 0) {
    if (g.nsrc != 2 || g.C_src[0] + g.C_src[1] != g.Cr) return false;
    for (int v = 0; v < 2; ++v)
      if (g.C_src[v] <= 0 || g.C_src[v] % 32 != 0 || !al16(g.in_src[v])) return false;
    return true;
  }
  if (!view_tma_ok(g.in, g.is)) return false;
  return true;
}

// Mask-mode epilogue (activation backward of the producing layer + bias-gradient column sums, conv_common.cuh): lives in
// the TMA-store path only.
bool tc_pixgemm_mask_ok(const PixGemm& g) {
  if (!tc_pixgemm_supported(g) || g.nphase != 0 || g.out_my != 1 || g.out_mx != 1) return false;
  if (pick_bn_pix(g) < 32) return false;
  return g.os.sc == 1 && (g.N % 4 == 0) && al16(g.out) && g.os.sx % 4 == 0 && g.os.sy % 4 == 0 && g.os.sb % 4 == 0 &&
         al16(g.ep.add) && !(tc_variant() & 16u);
}

// Tap split for small problems.  A layer at 4x4 ... 16x16 pixels (or a 1024-channel block at 16x16) has 8 ... 40 output
// tiles for 148 SMs, and every one of them walks the whole K loop (9 taps x C / 32 chunks = 144 stages at 512 channels):
// 13-190 TFLOP/s.  Such a call is run as `ksplit` groups of taps that accumulate separate partial outputs (a tile index
// carries its group; the persistent kernel is otherwise unchanged), followed by one pass that adds the groups in a fixed
// order and applies the epilogue — deterministic, and the partial sums of a few-megabyte output cost next to nothing.
// MSG_B200_TC_VARIANT bit 65536 disables it.  Returns the number of groups (1 = no split).
static int pix_tap_split(const PixGemm& g) {
  if (g.ksplit > 1) return g.ksplit;
  if (tc_variant() & 65536u) return 1;
  if (g.ntaps < 3 || g.nphase != 0 || g.nsrc != 0 || g.ep.add_is_mask || g.ep.colsum != nullptr) return 1;
  const int BN = pick_bn(g.N, 16);
  if (BN < 64) return 1;
  int wt_log2 = ilog2_ceil(g.PW);
  if (wt_log2 > 7) wt_log2 = 7;
  const int64_t tiles = ceil_div(g.PW, 1 << wt_log2) * ceil_div(g.PH, 128 >> wt_log2) * (int64_t)g.B * ceil_div(g.N, BN);
  if (tiles * 3 > num_sms()) return 1;                              // enough tiles, or too little to win
  if ((int64_t)g.ntaps * ceil_div(g.Cr, 32) < 32) return 1;         // short K loop
  int ks = (int)(num_sms() / tiles);
  if (ks > g.ntaps) ks = g.ntaps;
  const int tpg = (int)ceil_div(g.ntaps, ks);
  return (int)ceil_div(g.ntaps, tpg);
}

struct TapReduceParams {
  const float* part;        // [ksplit][B][PH][PW][N]
  int ksplit, B, PH, PW, N;
  float* out;
  View4 os;
  int out_my, out_mx, out_oy, out_ox;
  float alpha;
  Epilogue ep;
};

__global__ void __launch_bounds__(256)
tap_split_reduce_kernel(const TapReduceParams p) {
  const int64_t total = (int64_t)p.B * p.PH * p.PW * p.N;
  const float nw = p.ep.noise ? __ldg(p.ep.noise_w) : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int n = (int)(r % p.N); r /= p.N;
    const int x = (int)(r % p.PW); r /= p.PW;
    const int y = (int)(r % p.PH);
    const int b = (int)(r / p.PH);
    float acc = 0.f;
    for (int g = 0; g < p.ksplit; ++g) acc += __ldg(p.part + (int64_t)g * total + i);
    const int64_t off = (int64_t)b * p.os.sb + (int64_t)(y * p.out_my + p.out_oy) * p.os.sy +
                        (int64_t)(x * p.out_mx + p.out_ox) * p.os.sx + (int64_t)n * p.os.sc;
    const float bn = p.ep.bias ? __ldg(p.ep.bias + n) : 0.f;
    const float nz = p.ep.noise ? nw * __ldg(p.ep.noise + (int64_t)b * p.ep.noise_sb + (int64_t)y * p.PW + x) : 0.f;
    const float av = p.ep.add ? __ldg(p.ep.add + off) : 0.f;
    const float cs = p.ep.cscale ? __ldg(p.ep.cscale + (int64_t)b * p.ep.cscale_sb + n) : 1.f;
    const float o1 = apply_epilogue(p.ep, p.alpha * acc * cs, bn, nz, av);
    p.out[off] = o1;
    if (p.ep.out2) p.ep.out2[off] = o1 * __ldg(p.ep.out2_scale + (int64_t)b * p.ep.out2_scale_sb + n);
  }
}

size_t tc_pixgemm_workspace(const PixGemm& g) {
  const int ks = pix_tap_split(g);
  const int BN = ks > 1 ? pick_bn(g.N, 16) : pick_bn_pix(g);
  const int64_t Npad = round_up(g.N, BN), Cpad = round_up(g.Cr, 32);
  const int64_t BW = g.w_sb != 0 ? g.B : 1;
  size_t bytes = (size_t)(BW * g.ntaps * Npad * Cpad) * sizeof(float) + 256;
  if (ks > 1 && g.ksplit <= 1) bytes += (size_t)ks * g.B * g.PH * g.PW * g.N * sizeof(float) + 256;   // partial outputs
  return bytes;
}

template <int BN, int MT>
static int launch_pix(const TMapSet& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmAdd,
                      const CUtensorMap& tmOut2, const TcPixParams& p, cudaStream_t st) {
  constexpr int STAGES = (BN == 256 || MT == 2) ? 4 : 6;
  constexpr size_t smem = (size_t)STAGES * (MT * 16384 + BN * 128) + 4 * 2 * 4096 + 16 * STAGES + 128 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kfn = tc_pixgemm_kernel<BN, STAGES, MT>;
  static bool attr_done[64] = {};
  const int slot = current_device_slot();
  if (!attr_done[slot]) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[slot] = true;
  }
  // one persistent CTA per SM (two co-resident ones would have to share TMEM columns and the smem ring)
  const int ctas = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  kfn<<<ctas, 192, smem, st>>>(tmA, tmB, tmOut, tmAdd, tmOut2, p);
  MSG_CHECK_LAUNCH("conv pixgemm(tcgen05)");
  return MSG_OK;
}

template <int BN, int MT>
static int launch_pix_rows(const TMapSet& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmAdd,
                           const CUtensorMap& tmOut2, const TcPixParams& p, cudaStream_t st) {
  constexpr int KW = 3;
  constexpr size_t a_stage = (((size_t)(128 + KW - 1) * MT * 128) + 1023) & ~(size_t)1023;
  constexpr int SB = (BN >= 128) ? 6 : 8;
  constexpr size_t smem = 2 * a_stage + (size_t)SB * BN * 128 + 4 * 2 * 4096 + 16 * 2 + 16 * SB + 128 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kfn = tc_pixgemm_rows_kernel<BN, MT, KW>;
  static bool attr_done[64] = {};
  const int slot = current_device_slot();
  if (!attr_done[slot]) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[slot] = true;
  }
  const int ctas = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  kfn<<<ctas, 192, smem, st>>>(tmA, tmB, tmOut, tmAdd, tmOut2, p);
  MSG_CHECK_LAUNCH("conv pixgemm(tcgen05, row taps)");
  return MSG_OK;
}

static int max_active_pairs(const void* kfn, size_t smem) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (unsigned)num_sms());
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kfn, &cfg) != cudaSuccess || n <= 0) { (void)cudaGetLastError(); n = 0; }
  return n;
}

template <int BN>
static int launch_pix2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmAdd,
                       const CUtensorMap& tmOut2, const TcPixParams& p, cudaStream_t st) {
  constexpr int STAGES = BN == 256 ? 6 : 8;
  constexpr size_t smem = (size_t)STAGES * (16384 + (BN / 2) * 128) + 4 * 2 * 4096 + 16 * STAGES + 128 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kfn = tc_pixgemm2_kernel<BN, STAGES>;
  static int pairs_max_dev[64] = {};          // 0 = not queried yet, < 0 = no cluster fits
  const int slot = current_device_slot();
  if (pairs_max_dev[slot] == 0) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n = max_active_pairs(reinterpret_cast<const void*>(kfn), smem);
    pairs_max_dev[slot] = n > 0 ? n : -1;
  }
  const int pairs_max = pairs_max_dev[slot];
  if (pairs_max <= 0) return fail(MSG_ERR_UNSUPPORTED, "conv pixgemm(tcgen05, CTA pairs): no cluster can be resident");
  const int pairs = p.total_tiles < pairs_max ? p.total_tiles : pairs_max;
  kfn<<<2 * pairs, 192, smem, st>>>(tmA, tmB, tmOut, tmAdd, tmOut2, p);
  MSG_CHECK_LAUNCH("conv pixgemm(tcgen05, CTA pairs)");
  return MSG_OK;
}

template <int BN, int MT>
static int launch_pix2_rows(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmAdd,
                            const CUtensorMap& tmOut2, const TcPixParams& p, cudaStream_t st) {
  constexpr int KW = 3;
  constexpr size_t a_stage = (((size_t)(128 + KW - 1) * MT * 128) + 1023) & ~(size_t)1023;
  constexpr size_t smem = 3 * a_stage + (size_t)8 * (BN / 2) * 128 + 4 * 2 * 4096 + 16 * 3 + 16 * 8 + 128 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kfn = tc_pixgemm2_rows_kernel<BN, MT, KW>;
  static int pairs_max_dev[64] = {};
  const int slot = current_device_slot();
  if (pairs_max_dev[slot] == 0) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n = max_active_pairs(reinterpret_cast<const void*>(kfn), smem);
    pairs_max_dev[slot] = n > 0 ? n : -1;
  }
  const int pairs_max = pairs_max_dev[slot];
  if (pairs_max <= 0) return fail(MSG_ERR_UNSUPPORTED, "conv pixgemm(tcgen05, CTA pairs, row taps): no cluster can be resident");
  const int pairs = p.total_tiles < pairs_max ? p.total_tiles : pairs_max;
  kfn<<<2 * pairs, 192, smem, st>>>(tmA, tmB, tmOut, tmAdd, tmOut2, p);
  MSG_CHECK_LAUNCH("conv pixgemm(tcgen05, CTA pairs, row taps)");
  return MSG_OK;
}

// two 128-pixel sub-tiles per CTA tile pay off for narrow N once there is more than a round of such tiles
static int pick_mt(const PixGemm& g, int BN) {
  if (BN > 128) return 1;
  int wl = ilog2_ceil(g.PW);
  if (wl > 7) wl = 7;
  const int Wt = 1 << wl, Ht2 = 256 >> wl;
  const int64_t tiles2 = ceil_div(g.PW, Wt) * ceil_div(g.PH, Ht2) * (int64_t)(round_up(g.N, BN) / BN) * g.B;
  return tiles2 >= num_sms() ? 2 : 1;
}

int tc_pixgemm(const PixGemm& g, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!tc_pixgemm_supported(g)) return fail(MSG_ERR_UNSUPPORTED, "conv pixgemm(tcgen05): shape not supported");
  const size_t need = tc_pixgemm_workspace(g);
  if (!ws || ws_bytes < need) return fail(MSG_ERR_WORKSPACE, "conv pixgemm(tcgen05): workspace %zu < %zu", ws_bytes, need);
  if (g.ksplit <= 1) {
    const int ks = pix_tap_split(g);
    if (ks > 1) {
      // pass 1: the same kernel on `ks` tap groups, raw partial sums into the tail of the workspace
      PixGemm gs = g;
      gs.ksplit = ks;
      gs.os = View4{(int64_t)g.PH * g.PW * g.N, 1, (int64_t)g.PW * g.N, (int64_t)g.N};
      gs.out_my = gs.out_mx = 1; gs.out_oy = gs.out_ox = 0;
      gs.alpha = 1.f;
      gs.ep = Epilogue{};
      gs.ep.gain = 1.f;
      const size_t inner = tc_pixgemm_workspace(gs);            // transformed weights (same column tile as this call planned)
      uint8_t* wsb = reinterpret_cast<uint8_t*>(ws);
      float* part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(wsb + inner) + 255) & ~(uintptr_t)255);
      gs.out = part;
      int rc = tc_pixgemm(gs, ws, inner, st);
      if (rc) return rc;
      // pass 2: sum of the groups (fixed order) + the call's epilogue
      TapReduceParams rp{};
      rp.part = part; rp.ksplit = ks; rp.B = g.B; rp.PH = g.PH; rp.PW = g.PW; rp.N = g.N;
      rp.out = g.out; rp.os = g.os; rp.out_my = g.out_my; rp.out_mx = g.out_mx; rp.out_oy = g.out_oy; rp.out_ox = g.out_ox;
      rp.alpha = g.alpha; rp.ep = g.ep;
      const int64_t total = (int64_t)g.B * g.PH * g.PW * g.N;
      const int64_t want = ceil_div(total, 256), cap = (int64_t)num_sms() * 8;
      tap_split_reduce_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(rp);
      MSG_CHECK_LAUNCH("conv pixgemm(tap-split reduce)");
      return MSG_OK;
    }
  }
  float* wt = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  const int BN = g.ksplit > 1 ? pick_bn(g.N, 16) : pick_bn_pix(g);
  const int Npad = round_up(g.N, BN), Cpad = round_up(g.Cr, 32);
  const int BW = g.w_sb != 0 ? g.B : 1;

  WtParams wp{};
  wp.w = g.w; wp.w_sb = g.w_sb; wp.w_sn = g.w_sn; wp.w_sc = g.w_sc; wp.w_st = g.w_st;
  wp.N = g.N; wp.Cr = g.Cr; wp.Npad = Npad; wp.Cpad = Cpad; wp.ntaps = g.ntaps; wp.BW = BW;
  for (int t = 0; t < g.ntaps; ++t) wp.tap_wi[t] = g.tap_wi[t];
  {
    // Npad is a multiple of BN (>= 16) and the tiles are 32 wide: round the tile grid up and let the bounds checks
    // of the kernel (n < N) cover the remainder; Npad itself is a multiple of 32 whenever BN >= 32.
    if (Npad % 32 != 0 && Npad != 16) return fail(MSG_ERR_UNSUPPORTED, "conv weight transform: Npad %d", Npad);
    dim3 grid((unsigned)(Cpad / 32), (unsigned)((Npad + 31) / 32), (unsigned)BW);
    if (BW > 65535) return fail(MSG_ERR_UNSUPPORTED, "conv weight transform: batch > 65535");
    // contiguous taps, all of them used, two or more: the streaming variant
    const int64_t T = wp.w_sc <= wp.w_sn ? wp.w_sc : wp.w_sn;
    bool all_taps = wp.w_st == 1 && T == g.ntaps && T >= 2 && T <= 9;
    for (int t = 0; t < g.ntaps && all_taps; ++t) all_taps = wp.tap_wi[t] >= 0 && wp.tap_wi[t] < T;
    if (all_taps && !(tc_variant() & 512u) && Npad % 8 == 0 && !(tc_variant() & 16384u)) {
      const bool c_fast = wp.w_sc <= wp.w_sn;
      const size_t sm = (size_t)(c_fast ? 8 * (32 * T + 1) : 32 * (8 * T + 1)) * sizeof(float);
      dim3 grid8((unsigned)(Cpad / 32), (unsigned)(Npad / 8), (unsigned)BW);
      if (grid8.y > 65535) return fail(MSG_ERR_UNSUPPORTED, "conv weight transform: too many output channels");
      if (T <= 4) tc_weight_transform_runs8_kernel<4><<<grid8, 256, sm, st>>>(wt, wp, (int)T);
      else tc_weight_transform_runs8_kernel<9><<<grid8, 256, sm, st>>>(wt, wp, (int)T);
    } else if (all_taps && !(tc_variant() & 512u)) {
      const size_t sm = (size_t)32 * (32 * T + 1) * sizeof(float);
      tc_weight_transform_runs_kernel<<<grid, 256, sm, st>>>(wt, wp, (int)T);
    } else {
      tc_weight_transform_kernel<<<grid, 256, 0, st>>>(wt, wp);
    }
    MSG_CHECK_LAUNCH("conv weight transform");
  }

  int wt_log2 = ilog2_ceil(g.PW);
  if (wt_log2 > 7) wt_log2 = 7;
  // CTA pairs (two 128-pixel tiles, BN = 256 or 128) when there are enough pixel-tile pairs to fill the machine
  const int64_t ptiles = ceil_div(g.PW, 1 << wt_log2) * ceil_div(g.PH, 128 >> wt_log2);
  const int64_t pair_tiles = ((ptiles + 1) / 2) * (Npad / BN) * (int64_t)g.B;
  // Measured on B200 (tools/conv_bench.py, tools/epi_probe.py) with the TMA-store epilogue in both kernels: the pair kernel
  // is 3 % faster on the 3x3 512->512 layers (2.95 vs 3.05 ms at 256^2: half the weight bytes per CTA under a power-capped
  // clock) and 24 % faster on the load-path-bound 1x1 512->512 layers (125 vs 164 us back to back), slower when there are
  // fewer than two rounds of pair tiles (768->768 at 32^2).  MSG_B200_TC_VARIANT bit 4 disables it.
  const int64_t pair_min = (tc_variant() & 256u) ? 1 : num_sms();       // bit 256 (tests): pairs for any eligible shape
  // BN = 128 pairs exist (bit 1024) but measured the same as the 256-pixel single-CTA tiles (608 vs 607 and 659 vs 644
  // TFLOP/s on the discriminator's 128-channel 3x3 layers at 256^2; ncu: tensor pipe 69 % active either way), so they
  // stay opt-in.
  bool pairs = (BN == 256 || (BN == 128 && (tc_variant() & 1024u))) && !(tc_variant() & 4u) &&
               pair_tiles >= pair_min && g.nphase == 0 && g.nsrc == 0 && g.ksplit <= 1;
  // Row-tap kernels (activation box shared by the 3 taps of a filter row): 3-tap filter rows stored row-major with
  // consecutive dx (either direction), tiles that are whole 128-pixel image-row segments, N <= 128.
  // MSG_B200_TC_VARIANT bit 2048 disables them, bit 4096 only their CTA-pair form.
  bool rows_ok = wt_log2 == 7 && g.nphase == 0 && g.ntaps % 3 == 0 && g.ntaps >= 3 && !(tc_variant() & 2048u) && g.ksplit <= 1;
  for (int t = 0; t < g.ntaps && rows_ok; ++t) {
    const int t0 = t - t % 3, step = g.tap_dx[t0 + 1] - g.tap_dx[t0];       // +1 (forward) or -1 (dgrad) within a row
    rows_ok = (step == 1 || step == -1) && g.tap_dy[t] == g.tap_dy[t0] && g.tap_dx[t] == g.tap_dx[t0] + step * (t % 3);
  }
  int MT = (pairs || g.ksplit > 1) ? 1 : pick_mt(g, BN);
  // experiment knob (MSG_B200_TC_VARIANT bit 16384): single-tap GEMMs (1x1 layers, up-convolution phases) on 128-pixel tiles
  if ((tc_variant() & 16384u) && g.ntaps == 1) MT = 1;
  int64_t pair_tiles_used = pair_tiles;
  bool pairs_rows = false;
  if (pairs && rows_ok && BN == 256 && !(tc_variant() & (4096u | 8192u))) {
    // 256-channel tiles on the pair row-tap kernel too (one image-row segment per CTA; the accumulators of two would
    // need 1024 TMEM columns): 3x3 512->512 at 256^2 865 -> 924 TFLOP/s, at 128^2 742 -> 905 (tools/conv_bench.py;
    // a third of the activation fill, the power it saves goes into clock).  Bit 8192 keeps the per-tap pair kernel.
    pairs = false; pairs_rows = true; MT = 1;
  }
  if (!pairs && !pairs_rows && rows_ok && BN == 128 && g.nsrc == 0 && !(tc_variant() & (4u | 4096u))) {
    // N = 128: a single CTA's tensor core is bound by shared-memory bandwidth (operands 128 B/clk + fill); the pair form
    // reads half the weight rows per SM.  Two image-row segments per CTA when that still gives two rounds of pairs.
    for (int mt = 2; mt >= 1 && !pairs_rows; --mt) {
      const int64_t pt = ceil_div(g.PW, 128) * ceil_div(g.PH, mt);
      const int64_t prs = ((pt + 1) / 2) * (Npad / BN) * (int64_t)g.B;
      if (prs >= pair_min) { pairs_rows = true; MT = mt; pair_tiles_used = prs; }
    }
  }
  const bool rows3 = pairs_rows || (!pairs && rows_ok && (BN == 128 || BN == 64));
  const int Wt = 1 << wt_log2, Ht = (128 * MT) >> wt_log2;
  TMapSet tmA;
  CUtensorMap tmB;
  const int nviews = g.nphase > 0 ? g.nphase : (g.nsrc == 2 ? 2 : 1);
  for (int v = 0; v < 4; ++v) {
    const int vv = v < nviews ? v : 0;
    const float* basep = g.nphase > 0 ? g.in_ph[vv] : (g.nsrc == 2 ? g.in_src[vv] : g.in);
    const int ih = g.nphase > 0 ? g.IH_ph[vv] : g.IH, iw = g.nphase > 0 ? g.IW_ph[vv] : g.IW;
    const int cv = g.nsrc == 2 ? g.C_src[vv] : g.Cr / nviews;
    const View4 vs = g.nsrc == 2 ? View4{(int64_t)ih * iw * cv, 1, (int64_t)iw * cv, cv} : g.is;
    const uint64_t dims[4] = {(uint64_t)cv, (uint64_t)iw, (uint64_t)ih, (uint64_t)g.B};
    const uint64_t strides[3] = {(uint64_t)vs.sx * 4, (uint64_t)vs.sy * 4, (uint64_t)vs.sb * 4};
    const uint32_t box[4] = {32, (uint32_t)(rows3 ? Wt + 2 : Wt), (uint32_t)Ht, 1};
    if (v < nviews) {
      int rc = make_tmap(&tmA.m[v], basep, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    } else {
      tmA.m[v] = tmA.m[0];
    }
  }
  {
    const uint64_t dims[4] = {(uint64_t)Cpad, (uint64_t)Npad, (uint64_t)g.ntaps, (uint64_t)BW};
    const uint64_t strides[3] = {(uint64_t)Cpad * 4, (uint64_t)Cpad * Npad * 4, (uint64_t)Cpad * Npad * g.ntaps * 4};
    const uint32_t box[4] = {32, (uint32_t)((pairs || pairs_rows) ? BN / 2 : BN), 1, 1};     // a CTA of a pair loads half of the rows
    int rc = make_tmap(&tmB, wt, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  TcPixParams p{};
  p.ntaps = g.ntaps; p.cchunks = Cpad / 32;
  for (int v = 0; v < 4; ++v)
    p.cpv[v] = g.nphase > 0 ? (g.Cr / g.nphase) / 32 : (g.nsrc == 2 ? (v < 2 ? g.C_src[v] / 32 : 1) : Cpad / 32);
  for (int t = 0; t < g.ntaps; ++t) { p.tap_dy[t] = g.tap_dy[t]; p.tap_dx[t] = g.tap_dx[t]; }
  p.PH = g.PH; p.PW = g.PW; p.N = g.N;
  p.wt_log2 = wt_log2;
  p.tiles_x = (int)ceil_div(g.PW, Wt);
  p.tiles_y = (int)ceil_div(g.PH, Ht);
  p.n_tiles = Npad / BN;
  p.ksplit = g.ksplit > 1 ? g.ksplit : 1;
  p.tpg = (int)ceil_div(g.ntaps, p.ksplit);
  p.bsz = g.B;
  const int64_t total_tiles = (pairs || pairs_rows) ? pair_tiles_used
                                                    : (int64_t)p.tiles_x * p.tiles_y * p.n_tiles * g.B * p.ksplit;
  if (total_tiles > 0x7fffffffLL) return fail(MSG_ERR_UNSUPPORTED, "conv pixgemm(tcgen05): too many tiles");
  p.total_tiles = (int)total_tiles;
  p.out = g.out; p.os = g.os;
  p.out_my = g.out_my; p.out_mx = g.out_mx; p.out_oy = g.out_oy; p.out_ox = g.out_ox;
  p.alpha = g.alpha; p.w_per_sample = g.w_sb != 0;
  p.ep = g.ep;
  if (p.ep.noise && (g.out_my != 1 || g.out_mx != 1))
    return fail(MSG_ERR_UNSUPPORTED, "conv pixgemm(tcgen05): noise epilogue on a scattered output");
  p.vec_store = (g.os.sc == 1 && (g.N % 4 == 0) && al16(g.out) && g.os.sx % 4 == 0 && g.os.sy % 4 == 0 && g.os.sb % 4 == 0 &&
                 (!p.ep.bias || al16(p.ep.bias)) && (!p.ep.add || al16(p.ep.add)) &&
                 (!p.ep.cscale || (al16(p.ep.cscale) && p.ep.cscale_sb % 4 == 0)) &&
                 (!p.ep.out2 || (al16(p.ep.out2) && al16(p.ep.out2_scale) && p.ep.out2_scale_sb % 4 == 0))) ? 1 : 0;
  // output tensor map for the TMA-store epilogue: the (possibly phase-scattered) NHWC output as a 4-D view
  // (N, PW, PH, B) whose pixel strides carry out_mx / out_my and whose base carries the phase offset
  CUtensorMap tmOut;
  memset(&tmOut, 0, sizeof(tmOut));
  p.tma_store = 0;
  CUtensorMap tmAdd;
  memset(&tmAdd, 0, sizeof(tmAdd));
  CUtensorMap tmOut2;
  memset(&tmOut2, 0, sizeof(tmOut2));
  if ((p.ep.cscale || p.ep.out2) && p.ep.add)
    return fail(MSG_ERR_UNSUPPORTED, "conv pixgemm(tcgen05): channel scales / second output together with a residual operand");
  if ((p.ep.add_is_mask || p.ep.colsum) && !tc_pixgemm_mask_ok(g))
    return fail(MSG_ERR_UNSUPPORTED, "conv pixgemm(tcgen05): mask-mode epilogue needs the TMA-store path");
  if (p.ep.add_is_mask && (p.ep.bias || p.ep.noise || p.ep.act))
    return fail(MSG_ERR_UNSUPPORTED, "conv pixgemm(tcgen05): mask-mode epilogue cannot be combined with bias / noise / activation");
  if (p.ep.out2 && !p.ep.out2_scale) return fail(MSG_ERR_BAD_ARG, "conv pixgemm(tcgen05): out2 needs out2_scale");
  if (p.vec_store && BN >= 32 && !(tc_variant() & 16u)) {
    const int bw = Wt < 32 ? Wt : 32;
    const uint64_t dims[4] = {(uint64_t)g.N, (uint64_t)g.PW, (uint64_t)g.PH, (uint64_t)g.B * (uint64_t)p.ksplit};
    const uint64_t strides[3] = {(uint64_t)g.os.sx * g.out_mx * 4, (uint64_t)g.os.sy * g.out_my * 4, (uint64_t)g.os.sb * 4};
    const uint32_t box[4] = {32, (uint32_t)bw, (uint32_t)(32 / bw), 1};
    const int64_t view_off = (int64_t)g.out_oy * g.os.sy + (int64_t)g.out_ox * g.os.sx;
    float* obase = g.out + view_off;
    bool ok = al16(obase) && make_tmap(&tmOut, obase, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) == MSG_OK;
    if (ok && p.ep.add)       // the residual operand has the layout of the output: same view, other base
      ok = al16(p.ep.add + view_off) &&
           make_tmap(&tmAdd, p.ep.add + view_off, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) == MSG_OK;
    if (ok && p.ep.out2)      // second output: same view, other base
      ok = al16(p.ep.out2 + view_off) &&
           make_tmap(&tmOut2, p.ep.out2 + view_off, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B) == MSG_OK;
    p.tma_store = ok ? 1 : 0;
    p.debug = (int)((tc_variant() >> 5) & 3u);
  }
  if ((p.ep.add_is_mask || p.ep.colsum) && !p.tma_store)
    return fail(MSG_ERR_UNSUPPORTED, "conv pixgemm(tcgen05): mask-mode epilogue: output not expressible as a TMA store");
  cudaEvent_t pstop;
  const int pslot = prof_begin(0, g.ntaps, g.Cr, g.N, (int64_t)g.B * g.PH * g.PW,
                               2.0 * g.B * g.PH * g.PW * (double)g.N * g.Cr * g.ntaps, st, &pstop);
  p.kw = rows3 ? 3 : 0;
  int rc;
  if (pairs_rows) {
    if (BN == 256) rc = launch_pix2_rows<256, 1>(tmA.m[0], tmB, tmOut, tmAdd, tmOut2, p, st);
    else rc = MT == 2 ? launch_pix2_rows<128, 2>(tmA.m[0], tmB, tmOut, tmAdd, tmOut2, p, st)
                      : launch_pix2_rows<128, 1>(tmA.m[0], tmB, tmOut, tmAdd, tmOut2, p, st);
  } else if (rows3) {
    if (BN == 128) rc = MT == 2 ? launch_pix_rows<128, 2>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st)
                                : launch_pix_rows<128, 1>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st);
    else rc = MT == 2 ? launch_pix_rows<64, 2>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st)
                      : launch_pix_rows<64, 1>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st);
  } else if (pairs) {
    rc = BN == 256 ? launch_pix2<256>(tmA.m[0], tmB, tmOut, tmAdd, tmOut2, p, st) : launch_pix2<128>(tmA.m[0], tmB, tmOut, tmAdd, tmOut2, p, st);
  } else if (MT == 2) {
    switch (BN) {
      case 128: rc = launch_pix<128, 2>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st); break;
      case 64: rc = launch_pix<64, 2>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st); break;
      case 32: rc = launch_pix<32, 2>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st); break;
      default: rc = launch_pix<16, 2>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st); break;
    }
  } else {
    switch (BN) {
      case 256: rc = launch_pix<256, 1>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st); break;
      case 128: rc = launch_pix<128, 1>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st); break;
      case 64: rc = launch_pix<64, 1>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st); break;
      case 32: rc = launch_pix<32, 1>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st); break;
      default: rc = launch_pix<16, 1>(tmA, tmB, tmOut, tmAdd, tmOut2, p, st); break;
    }
  }
  prof_end(pslot, pstop, st);
  return rc;
}

// ---- RedGemm host -------------------------------------------------------------------------------
bool tc_redgemm_supported(const RedGemm& g) {
  if (!tc_available()) return false;
  if (g.ntaps <= 0 || g.ntaps > kMaxTaps || g.C <= 0 || g.N <= 0 || g.B <= 0) return false;
  if (g.my != 1 || g.mx != 1) return false;
  if (g.PW < 1 || g.PH < 1) return false;
  if (!view_tma_ok(g.in, g.is) || !view_tma_ok(g.g, g.gs)) return false;
  const int BS = g.dw_sb != 0 ? g.B : 1;
  if ((int64_t)g.ntaps * BS > 65535) return false;
  return true;
}

struct RedPlan {
  int BN, Npad, Cpad, BS, splits, wk_log2, chunks_x, chunks_y;
  int TG;                   // filter taps per CTA
  bool pairs;               // tc_redgemm2_kernel (CTA pairs, M = 256)
  size_t part_bytes;
};

static RedPlan red_plan(const RedGemm& g) {
  RedPlan pl{};
  pl.BN = pick_bn(g.C, 32);
  pl.Npad = round_up(g.N, 128);
  pl.Cpad = round_up(g.C, pl.BN);
  pl.BS = g.dw_sb != 0 ? g.B : 1;
  pl.wk_log2 = ilog2_ceil(g.PW);
  if (pl.wk_log2 > 5) pl.wk_log2 = 5;
  const int Wk = 1 << pl.wk_log2, Hk = 32 >> pl.wk_log2;
  pl.chunks_x = (int)ceil_div(g.PW, Wk);
  pl.chunks_y = (int)ceil_div(g.PH, Hk);
  const int64_t kiters = (int64_t)pl.chunks_x * pl.chunks_y * (g.dw_sb != 0 ? 1 : g.B);
  // narrow C tiles share the dy tile among several taps (one A box per TG taps)
  pl.TG = 1;
  if (g.ntaps > 1 && !(tc_variant() & 8u)) {
    if (pl.BN == 128) pl.TG = 3;
    else if (pl.BN == 64) pl.TG = 4;
    else if (pl.BN == 32) pl.TG = g.ntaps < 9 ? (g.ntaps >= 4 ? 4 : 1) : 9;
    if (pl.TG == 3 && g.ntaps < 3) pl.TG = 1;
    if (pl.TG == 4 && g.ntaps < 4) pl.TG = 1;
  }
  // CTA pairs for wide layers: 256 output channels x 256 input channels x 2 taps per pair (MSG_B200_TC_VARIANT bit 32768
  // keeps the single-CTA kernel)
  if (pl.BN == 256 && g.N >= 256 && g.N % 32 == 0 && g.C % 32 == 0 && !(tc_variant() & (2u | 32768u))) {
    pl.pairs = true;
    pl.Npad = round_up(g.N, 256);
    pl.TG = g.ntaps >= 2 ? 2 : 1;
    const int64_t full_groups = g.ntaps / pl.TG, rem = g.ntaps % pl.TG;
    const int64_t npairs = pl.Npad / 256, ctiles = pl.Cpad / 256;
    const int64_t max_by_k = kiters / 8 > 0 ? kiters / 8 : 1;
    const int64_t slots = num_sms() / 2;              // clusters resident at once
    int64_t best = 1;
    double best_t = 1e300;
    for (int64_t sp = 1; sp <= 2 * slots && sp <= max_by_k; ++sp) {
      // tap groups: one pair per (output block, C tile, group, split); the odd tap: one pair per two output blocks
      // (a leftover single block runs half the splits) — see tc_redgemm2_kernel
      const int64_t clusters = pl.BS * ctiles * (npairs * full_groups * sp +
                                                 (rem ? (npairs / 2) * sp + (npairs % 2) * ceil_div(sp, pl.TG) : 0));
      const int64_t rounds = ceil_div(clusters, slots);
      const double t = (double)rounds * ((double)kiters / (double)sp * pl.TG + 24.0) + 2.0 * (double)sp;
      if (t < best_t * 0.98) { best_t = t; best = sp; }
    }
    pl.splits = (int)best;
    pl.part_bytes = (size_t)pl.splits * pl.BS * g.ntaps * pl.Npad * pl.Cpad * sizeof(float);
    return pl;
  }
  const int64_t tgroups = ceil_div(g.ntaps, pl.TG);
  const int64_t tiles = (int64_t)(pl.Npad / 128) * (pl.Cpad / pl.BN) * tgroups * pl.BS;
  // split-K factor: fill whole waves of SMs (a grid of 324 CTAs on 148 SMs runs three rounds for 2.2 rounds of work),
  // keep >= 8 k-iterations per CTA, and among equally full grids prefer fewer splits (less partial-sum traffic).
  const int64_t max_by_k = kiters / 8 > 0 ? kiters / 8 : 1;
  const int64_t sms = num_sms();
  int64_t best = 1;
  double best_t = 1e300;
  for (int64_t sp = 1; sp <= 2 * sms && sp <= max_by_k; ++sp) {
    const int64_t rounds = ceil_div(tiles * sp, sms);
    // time ~ rounds x (k-iterations per CTA + ~24 iterations of fixed per-CTA cost) + the partial-sum reduction
    const double t = (double)rounds * ((double)kiters / (double)sp + 24.0) + 2.0 * (double)sp;
    if (t < best_t * 0.98) { best_t = t; best = sp; }
  }
  pl.splits = (int)best;
  pl.part_bytes = (size_t)pl.splits * pl.BS * g.ntaps * pl.Npad * pl.Cpad * sizeof(float);
  return pl;
}

size_t tc_redgemm_workspace(const RedGemm& g) { return red_plan(g).part_bytes + 256; }

template <int BN, int TG>
static int launch_red(const CUtensorMap& tmG, const CUtensorMap& tmI, const TcRedParams& p, dim3 grid,
                      cudaStream_t st) {
  constexpr size_t per_stage = 16384 + (size_t)TG * BN * 128;
  constexpr int STAGES = (per_stage * 6 + 4096 <= 227 * 1024) ? 6 : (per_stage * 4 + 4096 <= 227 * 1024) ? 4 : 3;
  constexpr size_t smem = (size_t)STAGES * per_stage + 16 * STAGES + 64 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kfn = tc_redgemm_kernel<BN, STAGES, TG>;
  static bool attr_done[64] = {};
  const int slot = current_device_slot();
  if (!attr_done[slot]) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[slot] = true;
  }
  kfn<<<grid, 192, smem, st>>>(tmG, tmI, p);
  MSG_CHECK_LAUNCH("conv redgemm(tcgen05)");
  return MSG_OK;
}

template <int TG>
static int launch_red2(const CUtensorMap& tmG, const CUtensorMap& tmI, const TcRedParams& p, dim3 grid, cudaStream_t st) {
  constexpr size_t per_stage = (size_t)(1 + TG) * 16384;
  constexpr int STAGES = (per_stage * 6 + 4096 <= 227 * 1024) ? 6 : 4;
  constexpr size_t smem = (size_t)STAGES * per_stage + 16 * STAGES + 64 + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  auto kfn = tc_redgemm2_kernel<STAGES, TG>;
  static bool attr_done[64] = {};
  const int slot = current_device_slot();
  if (!attr_done[slot]) {
    MSG_CHECK_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[slot] = true;
  }
  kfn<<<grid, 192, smem, st>>>(tmG, tmI, p);
  MSG_CHECK_LAUNCH("conv redgemm(tcgen05, CTA pairs)");
  return MSG_OK;
}

int tc_redgemm(const RedGemm& g, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (!tc_redgemm_supported(g)) return fail(MSG_ERR_UNSUPPORTED, "conv wgrad(tcgen05): shape not supported");
  const RedPlan pl = red_plan(g);
  if (!ws || ws_bytes < pl.part_bytes + 256)
    return fail(MSG_ERR_WORKSPACE, "conv wgrad(tcgen05): workspace %zu < %zu", ws_bytes, pl.part_bytes + 256);
  float* part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  const int Wk = 1 << pl.wk_log2, Hk = 32 >> pl.wk_log2;
  CUtensorMap tmG, tmI;
  // Channel counts that are whole 32-channel atoms: 5-D maps (32 ch, W, H, B, atoms) with the atom index as the slowest
  // box dimension, so ONE box per operand fills a stage (a bulk-tensor copy costs ~55 cycles + ~1.4 per 128-byte row;
  // twelve 4 KB boxes per stage made the producer the bottleneck).  Otherwise one 4-D box per atom.
  const bool a5d = (g.N % 32 == 0) && !(tc_variant() & 2u);
  const bool b5d = (g.C % 32 == 0) && !(tc_variant() & 2u);
  if (a5d) {
    const uint64_t dims[5] = {32, (uint64_t)g.PW, (uint64_t)g.PH, (uint64_t)g.B, (uint64_t)(g.N / 32)};
    const uint64_t strides[4] = {(uint64_t)g.gs.sx * 4, (uint64_t)g.gs.sy * 4, (uint64_t)g.gs.sb * 4, 128};
    const uint32_t box[5] = {32, (uint32_t)Wk, (uint32_t)Hk, 1, 4};
    int rc = make_tmap5(&tmG, g.g, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  } else {
    const uint64_t dims[4] = {(uint64_t)g.N, (uint64_t)g.PW, (uint64_t)g.PH, (uint64_t)g.B};
    const uint64_t strides[3] = {(uint64_t)g.gs.sx * 4, (uint64_t)g.gs.sy * 4, (uint64_t)g.gs.sb * 4};
    const uint32_t box[4] = {32, (uint32_t)Wk, (uint32_t)Hk, 1};
    int rc = make_tmap(&tmG, g.g, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  if (b5d) {
    const uint64_t dims[5] = {32, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.B, (uint64_t)(g.C / 32)};
    const uint64_t strides[4] = {(uint64_t)g.is.sx * 4, (uint64_t)g.is.sy * 4, (uint64_t)g.is.sb * 4, 128};
    const uint32_t box[5] = {32, (uint32_t)Wk, (uint32_t)Hk, 1, (uint32_t)(pl.pairs ? 4 : pl.BN / 32)};   // pairs: half per CTA
    int rc = make_tmap5(&tmI, g.in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  } else {
    const uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.IW, (uint64_t)g.IH, (uint64_t)g.B};
    const uint64_t strides[3] = {(uint64_t)g.is.sx * 4, (uint64_t)g.is.sy * 4, (uint64_t)g.is.sb * 4};
    const uint32_t box[4] = {32, (uint32_t)Wk, (uint32_t)Hk, 1};
    int rc = make_tmap(&tmI, g.in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    if (rc) return rc;
  }
  TcRedParams p{};
  p.ntaps = g.ntaps;
  for (int t = 0; t < g.ntaps; ++t) { p.tap_dy[t] = g.tap_dy[t]; p.tap_dx[t] = g.tap_dx[t]; }
  p.B = g.B; p.per_sample = g.dw_sb != 0;
  p.wk_log2 = pl.wk_log2; p.chunks_x = pl.chunks_x; p.chunks_y = pl.chunks_y;
  p.ctiles = pl.Cpad / pl.BN;
  p.Npad = pl.Npad; p.Cpad = pl.Cpad; p.splits = pl.splits; p.part = part;
  p.variant = tc_variant(); p.dbg = tc_debug_buffer();
  p.a5d = a5d ? 1 : 0; p.b5d = b5d ? 1 : 0;
  dim3 grid((unsigned)((pl.Npad / 128) * p.ctiles), (unsigned)(ceil_div(g.ntaps, pl.TG) * pl.BS), (unsigned)pl.splits);
  cudaEvent_t pstop;
  const int pslot = prof_begin(1, g.ntaps, g.C, g.N, (int64_t)g.B * g.PH * g.PW,
                               2.0 * g.B * g.PH * g.PW * (double)g.N * g.C * g.ntaps, st, &pstop);
  int rc;
  if (pl.pairs) {
    if (!a5d || !b5d) return fail(MSG_ERR_UNSUPPORTED, "conv wgrad(tcgen05, CTA pairs): needs whole 32-channel atoms");
    rc = pl.TG == 2 ? launch_red2<2>(tmG, tmI, p, grid, st) : launch_red2<1>(tmG, tmI, p, grid, st);
  } else
  switch (pl.BN * 16 + pl.TG) {
    case 256 * 16 + 1: rc = launch_red<256, 1>(tmG, tmI, p, grid, st); break;
    case 128 * 16 + 1: rc = launch_red<128, 1>(tmG, tmI, p, grid, st); break;
    case 128 * 16 + 3: rc = launch_red<128, 3>(tmG, tmI, p, grid, st); break;
    case 64 * 16 + 1: rc = launch_red<64, 1>(tmG, tmI, p, grid, st); break;
    case 64 * 16 + 4: rc = launch_red<64, 4>(tmG, tmI, p, grid, st); break;
    case 32 * 16 + 1: rc = launch_red<32, 1>(tmG, tmI, p, grid, st); break;
    case 32 * 16 + 4: rc = launch_red<32, 4>(tmG, tmI, p, grid, st); break;
    case 32 * 16 + 9: rc = launch_red<32, 9>(tmG, tmI, p, grid, st); break;
    default: return fail(MSG_ERR_UNSUPPORTED, "conv wgrad(tcgen05): no kernel for BN %d x %d taps per CTA", pl.BN, pl.TG);
  }
  prof_end(pslot, pstop, st);
  if (rc) return rc;

  RedReduceParams rp{};
  rp.part = part; rp.dw = g.dw; rp.dw_sb = g.dw_sb; rp.dw_sn = g.dw_sn; rp.dw_sc = g.dw_sc; rp.dw_st = g.dw_st;
  rp.splits = pl.splits; rp.BS = pl.BS; rp.ntaps = g.ntaps; rp.N = g.N; rp.C = g.C; rp.Npad = pl.Npad; rp.Cpad = pl.Cpad;
  for (int t = 0; t < g.ntaps; ++t) rp.tap_wi[t] = g.tap_wi[t];
  rp.alpha = g.alpha;
  const int64_t total = (int64_t)pl.BS * g.N * g.C;
  const int64_t want = ceil_div(total, 256);
  const int64_t cap = (int64_t)num_sms() * 16;
  tc_red_reduce_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(rp);
  MSG_CHECK_LAUNCH("conv wgrad reduce");
  return MSG_OK;
}

}  // namespace msg
