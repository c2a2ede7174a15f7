// Shared declarations of the softspoken_b200 kernels (internal; the public ABI is include/softspoken_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/softspoken_b200.h"

namespace ss {

// ---- constants of the path (mirrors softspoken_b200/spec.py; checked through ss_get_constant) ----
constexpr int kSampleRate = 22050;                 // settings.py:16
constexpr int kWindowSamples = 66150;              // NNDetector.py:74
constexpr int kWindowSamplesUsed = 65536;          // frame 255 ends at sample 65535
constexpr int kStepSamples = 13230;                // NNDetector.py:75
constexpr int kPadSamples = 66150;                 // worker.py:59
constexpr int kWin = 512;                          // settings.py:5
constexpr int kHop = 256;                          // settings.py:6
constexpr int kNfft = 2048;                        // pytorch_neural_nets.py:94
constexpr int kFrames = 256;                       // pytorch_neural_nets.py:150
constexpr int kMels = 128;                         // pytorch_neural_nets.py:87
constexpr int kFreqs = 1025;
constexpr int kMaxMelTaps = 32;    // rows of the feature kernel's transposed tap table
constexpr int kMaxMelRec = 1280;   // records of K1's two-band filterbank walk (743 bins + the overlap of the warps' ranges)
constexpr int kFeatureWarps = 16;  // warps of K1's CTA: the walk is cut into this many ranges
constexpr int kGapBins = 42;                       // 0.5 s break (worker.py:97) on the 256/3 Hz timeline
constexpr int kNumSMs = 148;

// Guard bands (ss_debug_check_guards): kGuardPattern reads as a small finite number in fp16, bf16 and fp32 alike, so
// the halo over-reads of the convolution kernel's first / last staged runs stay harmless.
constexpr unsigned char kGuardPattern = 0x3C;
constexpr size_t kCtxGuardBytes = 4096;
struct GuardBand {
  const unsigned char* ptr;
  size_t bytes;
  const void* owner;      // allocation the band belongs to
};
void register_guard(ss_ctx* ctx, const void* ptr, size_t bytes, const void* owner);
void unregister_guards(ss_ctx* ctx, const void* owner);

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int check_ctx_public(ss_ctx* ctx);   // every kernel launch of this library is counted (ss_launch_count)

#define SS_CUDA_CHECK(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ss::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SS_E_CUDA;                                                                      \
    }                                                                                        \
  } while (0)

#define SS_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ss::set_error(__VA_ARGS__);    \
      return (code);                 \
    }                                \
  } while (0)

// One folded convolution: weights [taps][C_in][C_out] float32, bias [C_out].
struct ConvW {
  const float* w = nullptr;
  const float* b = nullptr;
  int cin = 0, cout = 0, taps = 0;
};

struct ResBlockW {
  ConvW res, c1, c2;
};

// Front-end tables living on the device.
struct FrontEnd {
  const float* window = nullptr;      // [512] Hann (from the checkpoint)
  const float* tw_a_re = nullptr;     // [512] window[n] *  cos(2 pi n / 2048)
  const float* tw_a_im = nullptr;     // [512] window[n] * -sin(2 pi n / 2048)
  const float2* tw512 = nullptr;      // [512] exp(-2 pi i k / 512)
  const float2* tw1024 = nullptr;     // [512] exp(-2 pi i m / 1024), m < 512
  const int* mel_start = nullptr;     // [128]
  const int* mel_count = nullptr;     // [128]
  const int* mel_offs = nullptr;      // [128]
  const float* mel_taps = nullptr;    // [n_taps]
  int n_taps = 0;
  // "Two-band walk" of the filterbank for K1's mel phase (built at ss_ctx_create when every bin lies in at most two
  // CONSECUTIVE bands, as in any triangular bank; n_rec = 0 otherwise and K1 walks the sparse taps band by band):
  // the bins of a warp's bands in ascending order, each as {weight in its even-numbered band, weight in its
  // odd-numbered band, bin, emit}, emit = (even band that ends at this bin + 1) | (odd band that ends here + 1) << 8.
  const float4* mel_rec = nullptr;    // [n_rec]
  const int* mel_rec_begin = nullptr; // [17] record range of each of the 16 warps
  int n_rec = 0;
  int packed = 1;                     // K1 phase 1 on packed pairs (FFMA2); SS_K1_PACKED=0 at ss_ctx_create: the scalar path
};

enum { RB_CONV1 = 0, RB_CONV2, RB_CONV3, RB_CONV4, RB_BOTTLENECK, RB_ENCODER_OUT, RB_CONV6, RB_CONV7,
       RB_CONV8, RB_CONV9, RB_SPEC, RB_COUNT };

struct HeadW {
  const float* flat_w = nullptr;  // [128 mel][32][4]
  const float* flat_b = nullptr;  // [4]
  const float* res_w = nullptr;   // [1][4][4]
  const float* res_b = nullptr;
  const float* c1_w = nullptr;    // [3][4][4]
  const float* c1_b = nullptr;
  const float* c2_w = nullptr;    // [3][4][4]
  const float* c2_b = nullptr;
  const float* out_w = nullptr;   // [4]
  const float* out_b = nullptr;   // [1]
  const float* spec_w = nullptr;  // [1][32][2]
  const float* spec_b = nullptr;  // [2]
};

// fp32 activation workspace for `max_batch` windows (NHWC).
struct WorkspaceF32 {
  float* conv1 = nullptr;   // [B,128,256,32]
  float* pool1 = nullptr;   // [B,64,128,32]
  float* conv2 = nullptr;   // [B,64,128,64]
  float* pool2 = nullptr;   // [B,32,64,64]
  float* conv3 = nullptr;   // [B,32,64,96]
  float* pool3 = nullptr;   // [B,16,32,96]
  float* conv4 = nullptr;   // [B,16,32,128]
  float* pool4 = nullptr;   // [B,8,16,128]
  float* bott = nullptr;    // [B,8,16,128]
  float* enc = nullptr;     // [B,8,16,128]
  float* conv6 = nullptr;   // [B,16,32,96]
  float* conv7 = nullptr;   // [B,32,64,64]
  float* conv8 = nullptr;   // [B,64,128,32]
  float* conv9 = nullptr;   // [B,128,256,32]
  float* spec = nullptr;    // [B,128,256,32]
  float* tmp_t = nullptr;   // conv1 output of the block in flight (max size)
  float* tmp_r = nullptr;   // residual branch of the block in flight (max size)
};

}  // namespace ss

struct ss_ctx {
  int device = 0;
  int max_batch = 0;
  int f32_batch = 0;          // batch of the fp32 classifier (set with its lazily allocated workspace)
  size_t device_bytes = 0;
  float* blob_dev = nullptr;          // the whole float payload of the weight blob
  ss::FrontEnd fe;
  ss::ResBlockW rb[ss::RB_COUNT];
  ss::HeadW head;
  ss::WorkspaceF32 ws;
  void* tc[3] = {nullptr, nullptr, nullptr};   // tensor-core state per operand precision, see conv_tc.cu
  int tc_last = -1;                   // precision slot of the most recent tensor-core classify call
  // file-level scratch (ss_detect_*): grown only inside ss_ctx_reserve_file
  int64_t file_cap_samples = 0;
  float* file_mel = nullptr;
  float* file_logits = nullptr;
  double* file_avg = nullptr;
  int32_t* file_cnt = nullptr;
  int32_t* file_regions = nullptr;
  int32_t* file_nreg = nullptr;
  int32_t* slot_nreg = nullptr;       // region counters of the clips in flight in ss_detect_host_batch
  int32_t* slot_host = nullptr;       // pinned host mirror of the batch slots: per slot {count, cap * 2 ints}
  int stage_next = 0;                 // staging buffer the next streamed chunk uses
  int32_t* scan_tmp = nullptr;        // K6 per-block counts
  int64_t scan_tmp_len = 0;
  int file_region_cap = 0;
  int64_t file_cap_windows = 0;
  int64_t chunk_windows = 0;          // windows per streamed chunk of ss_detect_*
  float* intervals = nullptr;         // device interval table of ss_silence_host
  float* stage_buf[2] = {nullptr, nullptr};   // H2D staging of ss_detect_host (unpadded chunk samples)
  int64_t stage_cap = 0;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr};
  cudaEvent_t ev_consumed[2] = {nullptr, nullptr};
  cudaStream_t compute_stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  // margin-guided refinement (ss_ctx_set_refine): windows covering a bin whose average lies within refine_eps of the
  // threshold are classified again in refine_mode and K5 runs again on the patched logits
  double refine_eps = 0.0;
  int refine_mode = SS_MODE_FP32;
  unsigned char* win_flags = nullptr;   // [refine_cap_windows] marks of the clip in flight
  int64_t refine_cap_windows = 0;
  int32_t* refine_list = nullptr;       // [1 + refine_cap_windows] device: count, then ascending window indices
  int32_t* refine_host = nullptr;       // pinned mirror of refine_list
  void* refine_raw = nullptr;           // [kRefineBatch][65536] samples (float32 or int16) of the windows being refined
  int64_t* refine_starts = nullptr;     // [kRefineBatch] k * 65536
  float* refine_logits = nullptr;       // [kRefineBatch][256]
  uint64_t stat_windows = 0, stat_refined = 0, stat_clips = 0, stat_clips_refined = 0;
  std::map<const void*, void*> allocs;          // user pointer -> cudaMalloc base of the context's own allocations
  std::vector<ss::GuardBand> guards;
};

namespace ss {

// features.cu
int features_init();
int launch_features(const ss_ctx* ctx, const float* pcm, int64_t n_padded, const int64_t* starts, int n_windows,
                    float* mel, cudaStream_t st);
// Virtual padded clip: padded index idx in [valid_begin, valid_end) reads pcm[idx - offset], the rest are 0.
// starts == nullptr -> window w starts at (w_base + w) * 13230.
// sample_fmt: kSampleF32 (float32 samples) or kSampleS16 (int16 of a PCM_16 file, decoded as value / 32768 on load).
constexpr int kSampleF32 = 0, kSampleS16 = 1;
int launch_features_virtual(const ss_ctx* ctx, const void* pcm, int sample_fmt, int64_t valid_begin, int64_t valid_end,
                            int64_t offset, const int64_t* starts, int64_t w_base, int n_windows, float* mel,
                            cudaStream_t st);
int launch_pad(const float* src, int64_t n, float* dst, cudaStream_t st);
// resample.cu (K9): y[m] = sum_j x[mM div L - j] table[j + T][mM mod L], j = -T .. T; sample_fmt as for K1
int launch_resample(const void* x, int sample_fmt, int64_t n_in, float* y, int64_t n_out, int L, int M, int taps_half,
                    const float* table, cudaStream_t st);
// K8 (review-screen spectrogram): magnitudes [257][1 + n / 256]; max_bits (optional) receives the bit pattern of the maximum
int launch_spectrogram(const ss_ctx* ctx, const void* pcm, int sample_fmt, int64_t n, float* mag, unsigned int* max_bits,
                       cudaStream_t st);
int launch_spectrogram_db(float* mag, int64_t n_elems, const unsigned int* max_bits, cudaStream_t st);
int launch_window_starts(int64_t* starts, int64_t n_windows, cudaStream_t st);
// api.cu
int ensure_workspace_f32(ss_ctx* ctx);
// conv_fp32.cu
int classify_fp32(ss_ctx* ctx, const float* mel, int n_windows, float* logits, float* spec_out, cudaStream_t st);
// head.cu
int launch_mask_head_f32(const ss_ctx* ctx, const float* conv9_nhwc, int n_windows, float* logits, cudaStream_t st);
int launch_spec_out_f32(const ss_ctx* ctx, const float* spec_nhwc, int n_windows, float* spec_out_nchw, cudaStream_t st);
// postproc.cu
int launch_average(const float* logits, int n_windows, int64_t out_len, double* avg, int32_t* cnt, cudaStream_t st);
int launch_regions(const double* avg, const int32_t* cnt, int64_t out_len, double threshold, int gap_bins,
                   int32_t* regions, int32_t* n_regions, int cap, int32_t* scan_tmp, int64_t scan_tmp_len,
                   cudaStream_t st);
int64_t regions_scan_tmp_len(int64_t out_len);
int launch_average_regions(const float* logits, int n_windows, int64_t out_len, double* avg, int32_t* cnt,
                           double threshold, int gap_bins, int32_t* regions, int32_t* n_regions, int cap,
                           int32_t* scan_tmp, int64_t scan_tmp_len, cudaStream_t st);
// the two halves of launch_average_regions; win_flags (optional, one byte per window, zeroed by the caller) receives
// a mark for every window covering a bin whose average lies within eps of the threshold
int launch_average_bits(const float* logits, int n_windows, int64_t out_len, double* avg, int32_t* cnt,
                        double threshold, int32_t* scan_tmp, int64_t scan_tmp_len, double eps,
                        unsigned char* win_flags, cudaStream_t st);
int launch_regions_after_bits(int64_t out_len, int gap_bins, int32_t* regions, int32_t* n_regions, int cap,
                              int32_t* scan_tmp, int64_t scan_tmp_len, cudaStream_t st);
int launch_compact_flags(const unsigned char* flags, int n, int32_t* list, int32_t* count, cudaStream_t st);
int launch_scatter_rows(const float* rows, const int32_t* list, int n, float* logits, cudaStream_t st);
// silence.cu
// zero [begin - shift, end - shift) ∩ [0, n_elems) of pcm for every interval
int launch_silence(float* pcm, int64_t n_elems, int64_t shift, const ss_interval* iv, int n_intervals, cudaStream_t st);
int launch_silence_s16(int16_t* pcm, int64_t n_elems, int64_t shift, const ss_interval* iv, int n_intervals,
                       cudaStream_t st);
// pcm16.cu
int launch_decode_pcm16(const int16_t* interleaved, int64_t frames, int channels, float* mono, cudaStream_t st);
int launch_encode_pcm16(const float* src, int64_t n_elems, int16_t* dst, cudaStream_t st);
int launch_requant_pcm16(int16_t* pcm, int64_t n_elems, cudaStream_t st);
// conv_tc.cu
void tc_destroy(ss_ctx* ctx);
int classify_tc(ss_ctx* ctx, int mode, const float* mel, int n_windows, float* logits, float* spec_out,
                cudaStream_t st);
int tc_error_flag(ss_ctx* ctx, int* flag, int* range_flag, cudaStream_t st);
int tc_debug_profile(ss_ctx* ctx, int select_launch, long long* out_host);
int tc_debug_dump(ss_ctx* ctx, int which, int n_windows, float* out, int* C, int* H, int* W, cudaStream_t st);

// ------------------------------------------------------------------------------------------------------------
// Packed fp32 pairs.  sm_100 issues two fp32 operations per instruction on a 64-bit register pair (PTX add / sub / mul /
// fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2), each half rounded exactly as the scalar instruction would be.  The
// kernels that are bound by instruction issue rather than by a pipe (K1's transforms, conv1_direct) run their
// arithmetic on pairs of independent values through these.
#ifdef __CUDACC__
#define SS_P2_BINARY(name, op)                                                                                   \
  __device__ __forceinline__ float2 name(float2 a, float2 b) {                                                   \
    float2 d;                                                                                                    \
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; " op " rd, ra, rb; mov.b64 {%0, %1}, rd;}" \
        : "=f"(d.x), "=f"(d.y)                                                                                   \
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));                                                               \
    return d;                                                                                                    \
  }
SS_P2_BINARY(add2, "add.rn.f32x2")
SS_P2_BINARY(sub2, "sub.rn.f32x2")
SS_P2_BINARY(mul2, "mul.rn.f32x2")
#undef SS_P2_BINARY
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd;}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 bc2(float v) { return make_float2(v, v); }
#endif  // __CUDACC__

}  // namespace ss
