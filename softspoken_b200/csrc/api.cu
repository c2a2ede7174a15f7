// C ABI of softspoken_b200 (include/softspoken_b200.h): context, weight-blob parsing, entry points.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include <map>
#include <vector>

#include "ss_common.cuh"

namespace ss {

static thread_local char g_err[1024] = "";

static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {

constexpr uint32_t kBlobMagic = 0x53534232u;   // 'SSB2' (softspoken_b200/checkpoint.py)
constexpr uint32_t kBlobVersion = 2;
constexpr int64_t kChunkWindows = 1024;        // windows per streamed chunk in ss_detect_*
constexpr int kIntervalCap = 1 << 20;
constexpr int kBatchSlots = 8;                 // clips in flight inside ss_detect_host_batch

struct BlobEntry {
  uint64_t off, count;
};

struct BlobView {
  std::map<std::string, BlobEntry> entries;
  const float* payload = nullptr;
  uint64_t payload_floats = 0;
};

int parse_blob(const void* blob, size_t bytes, BlobView* out) {
  SS_REQUIRE(blob && bytes >= 16, SS_E_BLOB, "weight blob too small (%zu bytes)", bytes);
  const unsigned char* p = static_cast<const unsigned char*>(blob);
  uint32_t head[4];
  memcpy(head, p, 16);
  SS_REQUIRE(head[0] == kBlobMagic, SS_E_BLOB, "bad blob magic 0x%08x", head[0]);
  SS_REQUIRE(head[1] == kBlobVersion, SS_E_BLOB, "blob version %u, library expects %u", head[1], kBlobVersion);
  const uint32_t n = head[2];
  const size_t table = 16 + (size_t)n * 64;
  SS_REQUIRE(bytes >= table, SS_E_BLOB, "blob truncated inside its table");
  SS_REQUIRE((bytes - table) % 4 == 0, SS_E_BLOB, "blob payload is not a whole number of float32");
  out->payload = reinterpret_cast<const float*>(p + table);
  out->payload_floats = (bytes - table) / 4;
  for (uint32_t i = 0; i < n; ++i) {
    char name[49];
    memcpy(name, p + 16 + (size_t)i * 64, 48);
    name[48] = 0;
    BlobEntry e;
    memcpy(&e.off, p + 16 + (size_t)i * 64 + 48, 8);
    memcpy(&e.count, p + 16 + (size_t)i * 64 + 56, 8);
    SS_REQUIRE(e.off + e.count <= out->payload_floats, SS_E_BLOB, "blob entry '%s' out of bounds", name);
    out->entries[name] = e;
  }
  return SS_OK;
}

int find(const BlobView& v, const float* dev_payload, const std::string& name, uint64_t want_count,
         const float** dev_ptr) {
  auto it = v.entries.find(name);
  SS_REQUIRE(it != v.entries.end(), SS_E_BLOB, "blob entry '%s' missing", name.c_str());
  SS_REQUIRE(want_count == 0 || it->second.count == want_count, SS_E_BLOB,
             "blob entry '%s' has %llu elements, expected %llu", name.c_str(),
             (unsigned long long)it->second.count, (unsigned long long)want_count);
  *dev_ptr = dev_payload + it->second.off;
  return SS_OK;
}

struct RbSpec {
  const char* name;
  int cin, cout;
};
const RbSpec kResBlocks[RB_COUNT] = {
    {"conv1_1", 1, 32},          {"conv2_1", 32, 64},      {"conv3_1", 64, 96},   {"conv4_1", 96, 128},
    {"conv_bottleneck", 128, 128}, {"encoder_out", 128, 128}, {"conv6", 256, 96},    {"conv7", 192, 64},
    {"conv8", 128, 32},          {"conv9_1", 64, 32},      {"spec_output_conv.0", 32, 32}};

template <typename T>
int dev_alloc(ss_ctx* ctx, T** p, size_t count) {
  SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T)));
  ctx->device_bytes += count * sizeof(T);
  return SS_OK;
}

int upload_tables(ss_ctx* ctx, const BlobView& v) {
  // Twiddle tables are formed in double and rounded once.
  const BlobEntry& we = v.entries.at("window");
  std::vector<float> tab(512 * 2 + 1024 * 2);
  const double two_pi = 6.283185307179586476925286766559;
  for (int n = 0; n < 512; ++n) {
    const double w = (double)v.payload[we.off + n];
    tab[n] = (float)(w * cos(two_pi * n / 2048.0));
    tab[512 + n] = (float)(-w * sin(two_pi * n / 2048.0));
  }
  for (int k = 0; k < 512; ++k) {
    tab[1024 + 2 * k] = (float)cos(two_pi * k / 512.0);
    tab[1024 + 2 * k + 1] = (float)(-sin(two_pi * k / 512.0));
    tab[2048 + 2 * k] = (float)cos(two_pi * k / 1024.0);
    tab[2048 + 2 * k + 1] = (float)(-sin(two_pi * k / 1024.0));
  }
  float* d = nullptr;
  int rc = dev_alloc(ctx, &d, tab.size());
  if (rc) return rc;
  SS_CUDA_CHECK(cudaMemcpy(d, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
  ctx->fe.tw_a_re = d;
  ctx->fe.tw_a_im = d + 512;
  ctx->fe.tw512 = reinterpret_cast<const float2*>(d + 1024);
  ctx->fe.tw1024 = reinterpret_cast<const float2*>(d + 2048);
  return SS_OK;
}

}  // namespace

// The fp32 (CUDA-core) classifier's activation workspace, allocated on its first use: contexts that only run the
// tensor-core modes (the default) never pay for it.  The fp32 path is compute-bound at any batch, so its batch is
// capped at 64 windows whatever max_batch says (a 1,005-window batch would otherwise hold 28 GB here).
int ensure_workspace_f32(ss_ctx* ctx) {
  if (ctx->ws.conv1) return SS_OK;
  ctx->f32_batch = ctx->max_batch < 64 ? ctx->max_batch : 64;
  const size_t B = (size_t)ctx->f32_batch;
  WorkspaceF32& w = ctx->ws;
  int rc;
#define A(field, n) do { if ((rc = dev_alloc(ctx, &w.field, B * (size_t)(n)))) return rc; } while (0)
  A(conv1, 128 * 256 * 32);
  A(pool1, 64 * 128 * 32);
  A(conv2, 64 * 128 * 64);
  A(pool2, 32 * 64 * 64);
  A(conv3, 32 * 64 * 96);
  A(pool3, 16 * 32 * 96);
  A(conv4, 16 * 32 * 128);
  A(pool4, 8 * 16 * 128);
  A(bott, 8 * 16 * 128);
  A(enc, 8 * 16 * 128);
  A(conv6, 16 * 32 * 96);
  A(conv7, 32 * 64 * 64);
  A(conv8, 64 * 128 * 32);
  A(conv9, 128 * 256 * 32);
  A(spec, 128 * 256 * 32);
  A(tmp_t, 128 * 256 * 32);
  A(tmp_r, 128 * 256 * 32);
#undef A
  return SS_OK;
}

namespace {

bool valid_mode(int mode) { return mode >= SS_MODE_FP32 && mode <= SS_MODE_F16X3; }

int check_ctx(ss_ctx* ctx) {
  SS_REQUIRE(ctx != nullptr, SS_E_ARG, "null context");
  SS_CUDA_CHECK(cudaSetDevice(ctx->device));
  return SS_OK;
}

// Host replica of NNDetector.plan_detection_job's arithmetic on an integer sample count
// (NNDetector.py:72-77): L = n + 6 * 22050; W = ceil((L - 66150) / 13230).
int64_t plan_windows(int64_t n_samples) {
  const int64_t L = n_samples + 2 * (int64_t)kPadSamples;
  const int64_t num = L - kWindowSamples;
  if (num <= 0) return 0;
  return (num + kStepSamples - 1) / kStepSamples;
}

// NNDetector.py:168 `int(round(audio_length_seconds * 256 / 3))` with audio_length_seconds =
// n_padded / 22050 (worker.py:89): evaluated in double, left to right, Python round (half to even).
int64_t timeline_bins(int64_t n_padded) {
  const double secs = (double)n_padded / 22050.0;
  const double v = secs * 256.0 / 3.0;
  return (int64_t)nearbyint(v);   // default rounding mode = round-half-to-even, as Python's round()
}

// Windows [w0, w1) of the virtual padded clip -> logits[w0..w1) (features then classifier).
int run_windows(ss_ctx* ctx, const void* pcm, int fmt, int64_t valid_begin, int64_t valid_end, int64_t offset,
                int64_t w0, int64_t w1, int mode, float* logits_all, cudaStream_t st) {
  for (int64_t c0 = w0; c0 < w1; c0 += ctx->chunk_windows) {
    const int n = (int)((w1 - c0 < ctx->chunk_windows) ? (w1 - c0) : ctx->chunk_windows);
    int rc = launch_features_virtual(ctx, pcm, fmt, valid_begin, valid_end, offset, nullptr, c0, n, ctx->file_mel, st);
    if (rc) return rc;
    float* lg = logits_all + c0 * kFrames;
    if (mode == SS_MODE_FP32) rc = classify_fp32(ctx, ctx->file_mel, n, lg, nullptr, st);
    else rc = classify_tc(ctx, mode, ctx->file_mel, n, lg, nullptr, st);
    if (rc) return rc;
  }
  return SS_OK;
}

}  // namespace

int check_ctx_public(ss_ctx* ctx) { return check_ctx(ctx); }

}  // namespace ss

using namespace ss;

extern "C" {

int ss_abi_version(void) { return SS_ABI_VERSION; }

const char* ss_last_error(void) { return g_err; }

int ss_get_constant(const char* name, double* value) {
  SS_REQUIRE(name && value, SS_E_ARG, "null argument");
  struct KV { const char* k; double v; };
  static const KV table[] = {
      {"sample_rate", kSampleRate}, {"window_samples", kWindowSamples}, {"window_samples_used", kWindowSamplesUsed},
      {"step_samples", kStepSamples}, {"pad_samples", kPadSamples}, {"win_length", kWin}, {"hop_length", kHop},
      {"n_fft", kNfft}, {"n_frames", kFrames}, {"n_mels", kMels}, {"n_freqs", kFreqs}, {"gap_bins", kGapBins},
      {"threshold", 0.1}, {"max_mel_taps", kMaxMelTaps}, {"chunk_windows", (double)kChunkWindows}};
  for (const KV& kv : table)
    if (!strcmp(kv.k, name)) { *value = kv.v; return SS_OK; }
  set_error("unknown constant '%s'", name);
  return SS_E_ARG;
}

int ss_launch_count(uint64_t* count) {
  SS_REQUIRE(count, SS_E_ARG, "null argument");
  *count = __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
  return SS_OK;
}

int ss_device_count(int* count) {
  SS_REQUIRE(count, SS_E_ARG, "null argument");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
  *count = n;
  return SS_OK;
}

int64_t ss_plan_windows(int64_t n_samples) { return plan_windows(n_samples); }
int64_t ss_timeline_bins(int64_t n_padded) { return timeline_bins(n_padded); }

int ss_ctx_create(int device, const void* blob, size_t blob_bytes, int max_batch_windows, ss_ctx** out) {
  SS_REQUIRE(out, SS_E_ARG, "null context pointer");
  *out = nullptr;
  SS_REQUIRE(max_batch_windows >= 1 && max_batch_windows <= 4096, SS_E_ARG, "max_batch_windows %d out of [1, 4096]",
             max_batch_windows);
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    set_error("no CUDA device: softspoken_b200 has no CPU fallback");
    return SS_E_NODEVICE;
  }
  SS_REQUIRE(device >= 0 && device < n_dev, SS_E_ARG, "device %d out of range (%d devices)", device, n_dev);
  BlobView v;
  int rc = parse_blob(blob, blob_bytes, &v);
  if (rc) return rc;
  SS_CUDA_CHECK(cudaSetDevice(device));

  ss_ctx* ctx = new ss_ctx();
  ctx->device = device;
  ctx->max_batch = max_batch_windows;
  ctx->chunk_windows = kChunkWindows;
#define FAIL_IF(e) do { if ((rc = (e))) { ss_ctx_destroy(ctx); return rc; } } while (0)
  FAIL_IF(dev_alloc(ctx, &ctx->blob_dev, v.payload_floats));
  {
    cudaError_t e = cudaMemcpy(ctx->blob_dev, v.payload, v.payload_floats * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { set_error("weight upload failed: %s", cudaGetErrorString(e)); ss_ctx_destroy(ctx); return SS_E_CUDA; }
  }
  const float* d = ctx->blob_dev;
  const float* tmp = nullptr;
  FAIL_IF(find(v, d, "window", kWin, &ctx->fe.window));
  FAIL_IF(find(v, d, "mel_start", kMels, &tmp)); ctx->fe.mel_start = reinterpret_cast<const int*>(tmp);
  FAIL_IF(find(v, d, "mel_count", kMels, &tmp)); ctx->fe.mel_count = reinterpret_cast<const int*>(tmp);
  FAIL_IF(find(v, d, "mel_offs", kMels, &tmp)); ctx->fe.mel_offs = reinterpret_cast<const int*>(tmp);
  FAIL_IF(find(v, d, "mel_taps", 0, &ctx->fe.mel_taps));
  ctx->fe.n_taps = (int)v.entries.at("mel_taps").count;
  {
    // validate the sparse filterbank on the host copy: bands must stay inside bins [1, 743] that the
    // kernel forms, and inside the packed tap array
    const int* ms = reinterpret_cast<const int*>(v.payload + v.entries.at("mel_start").off);
    const int* mc = reinterpret_cast<const int*>(v.payload + v.entries.at("mel_count").off);
    const int* mo = reinterpret_cast<const int*>(v.payload + v.entries.at("mel_offs").off);
    for (int m = 0; m < kMels; ++m) {
      const bool ok = mc[m] >= 0 && mc[m] <= kMaxMelTaps && mo[m] >= 0 && mo[m] + mc[m] <= ctx->fe.n_taps &&
                      (mc[m] == 0 || (ms[m] >= 1 && ms[m] + mc[m] <= 744));
      if (!ok) {
        set_error("mel band %d (start %d, %d taps) is outside what the feature kernel computes (bins 1..743)", m,
                  ms[m], mc[m]);
        ss_ctx_destroy(ctx);
        return SS_E_BLOB;
      }
    }
  }
  FAIL_IF(upload_tables(ctx, v));
  for (int i = 0; i < RB_COUNT; ++i) {
    const RbSpec& s = kResBlocks[i];
    ResBlockW& rb = ctx->rb[i];
    const std::string p = s.name;
    rb.res = ConvW{nullptr, nullptr, s.cin, s.cout, 1};
    rb.c1 = ConvW{nullptr, nullptr, s.cin, s.cout, 9};
    rb.c2 = ConvW{nullptr, nullptr, s.cout, s.cout, 9};
    FAIL_IF(find(v, d, p + ".res.w", (uint64_t)s.cin * s.cout, &rb.res.w));
    FAIL_IF(find(v, d, p + ".res.b", s.cout, &rb.res.b));
    FAIL_IF(find(v, d, p + ".c1.w", (uint64_t)9 * s.cin * s.cout, &rb.c1.w));
    FAIL_IF(find(v, d, p + ".c1.b", s.cout, &rb.c1.b));
    FAIL_IF(find(v, d, p + ".c2.w", (uint64_t)9 * s.cout * s.cout, &rb.c2.w));
    FAIL_IF(find(v, d, p + ".c2.b", s.cout, &rb.c2.b));
  }
  HeadW& h = ctx->head;
  FAIL_IF(find(v, d, "conv_flatten.w", 128 * 32 * 4, &h.flat_w));
  FAIL_IF(find(v, d, "conv_flatten.b", 4, &h.flat_b));
  FAIL_IF(find(v, d, "mask_output_conv.0.res.w", 16, &h.res_w));
  FAIL_IF(find(v, d, "mask_output_conv.0.res.b", 4, &h.res_b));
  FAIL_IF(find(v, d, "mask_output_conv.0.c1.w", 48, &h.c1_w));
  FAIL_IF(find(v, d, "mask_output_conv.0.c1.b", 4, &h.c1_b));
  FAIL_IF(find(v, d, "mask_output_conv.0.c2.w", 48, &h.c2_w));
  FAIL_IF(find(v, d, "mask_output_conv.0.c2.b", 4, &h.c2_b));
  FAIL_IF(find(v, d, "mask_output_conv.1.w", 4, &h.out_w));
  FAIL_IF(find(v, d, "mask_output_conv.1.b", 1, &h.out_b));
  FAIL_IF(find(v, d, "spec_output_conv.1.w", 64, &h.spec_w));
  FAIL_IF(find(v, d, "spec_output_conv.1.b", 2, &h.spec_b));
  FAIL_IF(features_init());
  {
    cudaError_t e = cudaStreamCreateWithFlags(&ctx->compute_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
      e = cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_consumed[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { set_error("stream/event creation failed: %s", cudaGetErrorString(e)); ss_ctx_destroy(ctx); return SS_E_CUDA; }
  }
#undef FAIL_IF
  *out = ctx;
  return SS_OK;
}

int ss_ctx_destroy(ss_ctx* ctx) {
  if (!ctx) return SS_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  tc_destroy(ctx);
  WorkspaceF32& w = ctx->ws;
  float* bufs[] = {ctx->blob_dev, const_cast<float*>(ctx->fe.tw_a_re), w.conv1, w.pool1, w.conv2, w.pool2, w.conv3,
                   w.pool3, w.conv4, w.pool4, w.bott, w.enc, w.conv6, w.conv7, w.conv8, w.conv9, w.spec, w.tmp_t,
                   w.tmp_r, ctx->file_mel, ctx->file_logits, ctx->stage_buf[0], ctx->stage_buf[1]};
  for (float* p : bufs) if (p) cudaFree(p);
  if (ctx->file_avg) cudaFree(ctx->file_avg);
  if (ctx->file_cnt) cudaFree(ctx->file_cnt);
  if (ctx->file_regions) cudaFree(ctx->file_regions);
  if (ctx->file_nreg) cudaFree(ctx->file_nreg);
  if (ctx->slot_nreg) cudaFree(ctx->slot_nreg);
  if (ctx->slot_host) cudaFreeHost(ctx->slot_host);
  if (ctx->scan_tmp) cudaFree(ctx->scan_tmp);
  if (ctx->intervals) cudaFree(ctx->intervals);
  for (int i = 0; i < 2; ++i) {
    if (ctx->ev_copied[i]) cudaEventDestroy(ctx->ev_copied[i]);
    if (ctx->ev_consumed[i]) cudaEventDestroy(ctx->ev_consumed[i]);
  }
  if (ctx->compute_stream) cudaStreamDestroy(ctx->compute_stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return SS_OK;
}

int ss_ctx_device_bytes(ss_ctx* ctx, size_t* bytes) {
  SS_REQUIRE(ctx && bytes, SS_E_ARG, "null argument");
  *bytes = ctx->device_bytes;
  return SS_OK;
}

int ss_ctx_reserve(ss_ctx* ctx, int64_t max_samples, int region_cap) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(max_samples >= 0 && region_cap >= 1, SS_E_ARG, "bad reservation (%lld samples, %d regions)",
             (long long)max_samples, region_cap);
  if (!ctx->file_mel) {
    if ((rc = dev_alloc(ctx, &ctx->file_mel, (size_t)ctx->chunk_windows * kMels * kFrames))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->file_nreg, 1))) return rc;
    // staging: one chunk of windows spans (chunk-1)*13230 + 65536 samples
    ctx->stage_cap = (ctx->chunk_windows - 1) * kStepSamples + kWindowSamplesUsed;
    for (int i = 0; i < 2; ++i)
      if ((rc = dev_alloc(ctx, &ctx->stage_buf[i], (size_t)ctx->stage_cap))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->intervals, (size_t)kIntervalCap))) return rc;
  }
  if (max_samples > ctx->file_cap_samples || !ctx->file_logits) {
    SS_CUDA_CHECK(cudaDeviceSynchronize());
    const int64_t W = plan_windows(max_samples);
    const int64_t bins = timeline_bins(max_samples + 2 * (int64_t)kPadSamples) + 1;
    if (ctx->file_logits) { cudaFree(ctx->file_logits); cudaFree(ctx->file_avg); cudaFree(ctx->file_cnt); cudaFree(ctx->scan_tmp); }
    ctx->file_logits = nullptr; ctx->file_avg = nullptr; ctx->file_cnt = nullptr; ctx->scan_tmp = nullptr;
    if ((rc = dev_alloc(ctx, &ctx->file_logits, (size_t)(W > 0 ? W : 1) * kFrames))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->file_avg, (size_t)bins))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->file_cnt, (size_t)bins))) return rc;
    ctx->scan_tmp_len = regions_scan_tmp_len(bins);
    if ((rc = dev_alloc(ctx, &ctx->scan_tmp, (size_t)ctx->scan_tmp_len))) return rc;
    ctx->file_cap_samples = max_samples;
    ctx->file_cap_windows = W;
  }
  if (region_cap > ctx->file_region_cap) {
    SS_CUDA_CHECK(cudaDeviceSynchronize());
    if (ctx->file_regions) cudaFree(ctx->file_regions);
    if (ctx->slot_nreg) cudaFree(ctx->slot_nreg);
    if (ctx->slot_host) cudaFreeHost(ctx->slot_host);
    ctx->file_regions = nullptr; ctx->slot_nreg = nullptr; ctx->slot_host = nullptr;
    // kBatchSlots region buffers (slot 0 doubles as the single-clip buffer) + counters + a pinned host mirror
    if ((rc = dev_alloc(ctx, &ctx->file_regions, (size_t)kBatchSlots * region_cap * 2))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->slot_nreg, (size_t)kBatchSlots))) return rc;
    SS_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&ctx->slot_host),
                                 (size_t)kBatchSlots * ((size_t)region_cap * 2 + 1) * sizeof(int32_t)));
    ctx->file_region_cap = region_cap;
  }
  return SS_OK;
}

int ss_pad(ss_ctx* ctx, const float* pcm_dev, int64_t n_samples, float* padded_dev, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_samples >= 0 && padded_dev && (pcm_dev || n_samples == 0), SS_E_ARG, "bad ss_pad arguments");
  return launch_pad(pcm_dev, n_samples, padded_dev, static_cast<cudaStream_t>(stream));
}

int ss_features(ss_ctx* ctx, const float* pcm_dev, int64_t n_padded, const int64_t* win_start_dev, int n_windows,
                float* mel_out_dev, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_windows >= 0 && n_padded >= 0, SS_E_ARG, "negative size");
  if (n_windows == 0) return SS_OK;
  SS_REQUIRE(pcm_dev && win_start_dev && mel_out_dev, SS_E_ARG, "null device pointer");
  return launch_features(ctx, pcm_dev, n_padded, win_start_dev, n_windows, mel_out_dev,
                         static_cast<cudaStream_t>(stream));
}

static int resample_impl(ss_ctx* ctx, const void* pcm_dev, int fmt, int64_t n_in, float* out_dev, int64_t n_out, int up,
                         int down, int taps_half, const float* table_dev, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_in >= 0 && n_out >= 0 && up >= 1 && down >= 1, SS_E_ARG, "bad ss_resample arguments");
  SS_REQUIRE(n_out == (n_in * up + down - 1) / down, SS_E_ARG, "n_out %lld != ceil(n_in * %d / %d) = %lld",
             (long long)n_out, up, down, (long long)((n_in * up + down - 1) / down));
  if (n_out == 0) return SS_OK;
  SS_REQUIRE(pcm_dev && out_dev && table_dev, SS_E_ARG, "null device pointer");
  return launch_resample(pcm_dev, fmt, n_in, out_dev, n_out, up, down, taps_half, table_dev,
                         static_cast<cudaStream_t>(stream));
}

int ss_resample(ss_ctx* ctx, const float* pcm_dev, int64_t n_in, float* out_dev, int64_t n_out, int up, int down,
                int taps_half, const float* table_dev, void* stream) {
  return resample_impl(ctx, pcm_dev, kSampleF32, n_in, out_dev, n_out, up, down, taps_half, table_dev, stream);
}

int ss_resample_pcm16(ss_ctx* ctx, const int16_t* pcm_dev, int64_t n_in, float* out_dev, int64_t n_out, int up, int down,
                      int taps_half, const float* table_dev, void* stream) {
  return resample_impl(ctx, pcm_dev, kSampleS16, n_in, out_dev, n_out, up, down, taps_half, table_dev, stream);
}

int64_t ss_spectrogram_frames(int64_t n_samples) { return n_samples < 0 ? 0 : 1 + n_samples / kHop; }

static int spectrogram_impl(ss_ctx* ctx, const void* pcm_dev, int fmt, int64_t n_samples, float* mag_dev, float* max_dev,
                            void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_samples >= 0, SS_E_ARG, "negative size");
  SS_REQUIRE(mag_dev && (pcm_dev || n_samples == 0), SS_E_ARG, "null device pointer");
  return launch_spectrogram(ctx, pcm_dev, fmt, n_samples, mag_dev, reinterpret_cast<unsigned int*>(max_dev),
                            static_cast<cudaStream_t>(stream));
}

int ss_spectrogram(ss_ctx* ctx, const float* pcm_dev, int64_t n_samples, float* mag_dev, float* max_dev, void* stream) {
  return spectrogram_impl(ctx, pcm_dev, kSampleF32, n_samples, mag_dev, max_dev, stream);
}

int ss_spectrogram_pcm16(ss_ctx* ctx, const int16_t* pcm_dev, int64_t n_samples, float* mag_dev, float* max_dev,
                         void* stream) {
  return spectrogram_impl(ctx, pcm_dev, kSampleS16, n_samples, mag_dev, max_dev, stream);
}

int ss_spectrogram_db(ss_ctx* ctx, float* mag_dev, int64_t n_elems, const float* max_dev, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_elems >= 0, SS_E_ARG, "negative size");
  if (n_elems == 0) return SS_OK;
  SS_REQUIRE(mag_dev && max_dev, SS_E_ARG, "null device pointer");
  return launch_spectrogram_db(mag_dev, n_elems, reinterpret_cast<const unsigned int*>(max_dev),
                               static_cast<cudaStream_t>(stream));
}

int ss_classify(ss_ctx* ctx, const float* mel_dev, int n_windows, float* logits_dev, float* spec_out_dev, int mode,
                void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_windows >= 0, SS_E_ARG, "negative window count");
  if (n_windows == 0) return SS_OK;
  SS_REQUIRE(mel_dev && logits_dev, SS_E_ARG, "null device pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SS_REQUIRE(valid_mode(mode), SS_E_ARG, "unknown classifier mode %d", mode);
  if (mode == SS_MODE_FP32) return classify_fp32(ctx, mel_dev, n_windows, logits_dev, spec_out_dev, st);
  return classify_tc(ctx, mode, mel_dev, n_windows, logits_dev, spec_out_dev, st);
}

int ss_average(ss_ctx* ctx, const float* logits_dev, int n_windows, int64_t out_len, double* avg_dev,
               int32_t* count_dev, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_windows >= 0 && out_len >= 0, SS_E_ARG, "negative size");
  SS_REQUIRE(out_len == 0 || (avg_dev && count_dev), SS_E_ARG, "null device pointer");
  SS_REQUIRE(n_windows == 0 || logits_dev, SS_E_ARG, "null logits");
  if (n_windows > 0) {
    const int64_t last = (256 * (int64_t)(n_windows - 1) + 2) / 5 + 256;
    SS_REQUIRE(last <= out_len, SS_E_ARG, "window %d ends at bin %lld beyond the %lld-bin timeline", n_windows - 1,
               (long long)last, (long long)out_len);
  }
  return launch_average(logits_dev, n_windows, out_len, avg_dev, count_dev, static_cast<cudaStream_t>(stream));
}

int ss_regions(ss_ctx* ctx, const double* avg_dev, const int32_t* count_dev, int64_t out_len, double threshold,
               int gap_bins, int32_t* regions_dev, int32_t* n_regions_dev, int cap, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(out_len >= 0 && cap >= 0 && n_regions_dev, SS_E_ARG, "bad ss_regions arguments");
  SS_REQUIRE(ctx->scan_tmp && regions_scan_tmp_len(out_len) <= ctx->scan_tmp_len, SS_E_CAPACITY,
             "timeline of %lld bins exceeds the reservation: call ss_ctx_reserve first", (long long)out_len);
  return launch_regions(avg_dev, count_dev, out_len, threshold, gap_bins, regions_dev, n_regions_dev, cap,
                        ctx->scan_tmp, ctx->scan_tmp_len, static_cast<cudaStream_t>(stream));
}

int ss_silence(ss_ctx* ctx, float* pcm_dev, int64_t n_elems, const ss_interval* intervals_dev, int n_intervals,
               void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_elems >= 0 && n_intervals >= 0, SS_E_ARG, "negative size");
  if (n_intervals == 0 || n_elems == 0) return SS_OK;
  SS_REQUIRE(pcm_dev && intervals_dev, SS_E_ARG, "null device pointer");
  return launch_silence(pcm_dev, n_elems, 0, intervals_dev, n_intervals, static_cast<cudaStream_t>(stream));
}

static int detect_tail(ss_ctx* ctx, int64_t n_samples, int64_t W, int32_t* regions_dev, int32_t* n_regions_dev,
                       int cap, cudaStream_t st) {
  const int64_t bins = timeline_bins(n_samples + 2 * (int64_t)kPadSamples);
  return launch_average_regions(ctx->file_logits, (int)W, bins, ctx->file_avg, ctx->file_cnt, 0.1, kGapBins, regions_dev,
                                n_regions_dev, cap, ctx->scan_tmp, ctx->scan_tmp_len, st);
}

static int detect_device_impl(ss_ctx* ctx, const void* pcm_dev, int fmt, int64_t n_samples, int mode,
                              int32_t* regions_dev, int32_t* n_regions_dev, int cap, float* logits_out_dev,
                              void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_samples >= 0 && cap >= 0 && regions_dev && n_regions_dev, SS_E_ARG, "bad ss_detect_device arguments");
  SS_REQUIRE(pcm_dev || n_samples == 0, SS_E_ARG, "null pcm");
  SS_REQUIRE(ctx->file_logits && n_samples <= ctx->file_cap_samples, SS_E_CAPACITY,
             "clip of %lld samples exceeds the reservation of %lld: call ss_ctx_reserve", (long long)n_samples,
             (long long)ctx->file_cap_samples);
  SS_REQUIRE(valid_mode(mode), SS_E_ARG, "unknown classifier mode %d", mode);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t W = plan_windows(n_samples);
  // virtual padding (worker.py:58-62): padded index [66150, 66150 + n) -> pcm[idx - 66150], zeros elsewhere
  rc = run_windows(ctx, pcm_dev, fmt, kPadSamples, kPadSamples + n_samples, kPadSamples, 0, W, mode, ctx->file_logits,
                   st);
  if (rc) return rc;
  rc = detect_tail(ctx, n_samples, W, regions_dev, n_regions_dev, cap, st);
  if (rc) return rc;
  if (logits_out_dev && W > 0)
    SS_CUDA_CHECK(cudaMemcpyAsync(logits_out_dev, ctx->file_logits, (size_t)W * kFrames * sizeof(float),
                                  cudaMemcpyDeviceToDevice, st));
  return SS_OK;
}

int ss_detect_device(ss_ctx* ctx, const float* pcm_dev, int64_t n_samples, int mode, int32_t* regions_dev,
                     int32_t* n_regions_dev, int cap, float* logits_out_dev, void* stream) {
  return detect_device_impl(ctx, pcm_dev, kSampleF32, n_samples, mode, regions_dev, n_regions_dev, cap, logits_out_dev,
                            stream);
}

int ss_detect_device_pcm16(ss_ctx* ctx, const int16_t* pcm_dev, int64_t n_samples, int mode, int32_t* regions_dev,
                           int32_t* n_regions_dev, int cap, float* logits_out_dev, void* stream) {
  return detect_device_impl(ctx, pcm_dev, kSampleS16, n_samples, mode, regions_dev, n_regions_dev, cap, logits_out_dev,
                            stream);
}

// Enqueue one host clip: chunked H2D on the copy stream (double-buffered staging), K1-K3 per chunk and K5/K6 on the
// compute stream, regions left in `regions_dev` / `nreg_dev`.  Does not synchronise.
static int enqueue_detect_host(ss_ctx* ctx, const void* pcm_host, int fmt, int64_t n_samples, int mode,
                               int32_t* regions_dev, int32_t* nreg_dev, int cap) {
  const size_t esz = (fmt == kSampleS16) ? sizeof(int16_t) : sizeof(float);   // the staging buffers hold either type
  cudaStream_t cs = ctx->compute_stream, xs = ctx->copy_stream;
  const int64_t W = plan_windows(n_samples);
  int rc;
  for (int64_t w0 = 0; w0 < W; w0 += ctx->chunk_windows) {
    const int buf = ctx->stage_next;
    ctx->stage_next ^= 1;
    const int64_t w1 = (w0 + ctx->chunk_windows < W) ? w0 + ctx->chunk_windows : W;
    // padded sample range the chunk's kept frames touch: [w0*step - 256 (reflection stays >= w0*step), ...)
    const int64_t plo = w0 * kStepSamples, phi = (w1 - 1) * kStepSamples + kWindowSamplesUsed;
    int64_t s0 = plo - kPadSamples, s1 = phi - kPadSamples;     // unpadded coordinates
    if (s0 < 0) s0 = 0;
    if (s1 > n_samples) s1 = n_samples;
    if (s1 < s0) s1 = s0;
    SS_CUDA_CHECK(cudaStreamWaitEvent(xs, ctx->ev_consumed[buf], 0));   // staging buffer free again
    if (s1 > s0)
      SS_CUDA_CHECK(cudaMemcpyAsync(ctx->stage_buf[buf], static_cast<const char*>(pcm_host) + (size_t)s0 * esz,
                                    (size_t)(s1 - s0) * esz, cudaMemcpyHostToDevice, xs));
    SS_CUDA_CHECK(cudaEventRecord(ctx->ev_copied[buf], xs));
    SS_CUDA_CHECK(cudaStreamWaitEvent(cs, ctx->ev_copied[buf], 0));
    rc = run_windows(ctx, ctx->stage_buf[buf], fmt, kPadSamples + s0, kPadSamples + s1, kPadSamples + s0, w0, w1, mode,
                     ctx->file_logits, cs);
    if (rc) return rc;
    SS_CUDA_CHECK(cudaEventRecord(ctx->ev_consumed[buf], cs));
  }
  return detect_tail(ctx, n_samples, W, regions_dev, nreg_dev, cap, cs);
}

static int check_tc_health(ss_ctx* ctx, int mode, cudaStream_t cs) {
  if (mode == SS_MODE_FP32) return SS_OK;
  int flag = 0, range = 0;
  int rc = tc_error_flag(ctx, &flag, &range, cs);
  if (rc) return rc;
  SS_REQUIRE(flag == 0, SS_E_CUDA, "tcgen05 pipeline timed out (role code %d)", flag);
  SS_REQUIRE(range == 0, SS_E_RANGE,
             "an activation exceeded the fp16 range (65504) in an fp16-operand classifier mode: the result is not "
             "valid; use SS_MODE_BF16 or SS_MODE_FP32 for this checkpoint");
  return SS_OK;
}

static int detect_host_impl(ss_ctx* ctx, const void* pcm_host, int fmt, int64_t n_samples, int mode,
                            int32_t* regions_host, int cap, int* n_regions, float* logits_host) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_samples >= 0 && cap >= 0 && n_regions && (regions_host || cap == 0), SS_E_ARG,
             "bad ss_detect_host arguments");
  SS_REQUIRE(pcm_host || n_samples == 0, SS_E_ARG, "null pcm");
  SS_REQUIRE(ctx->file_logits && n_samples <= ctx->file_cap_samples && cap <= ctx->file_region_cap, SS_E_CAPACITY,
             "clip of %lld samples / %d regions exceeds the reservation (%lld / %d): call ss_ctx_reserve",
             (long long)n_samples, cap, (long long)ctx->file_cap_samples, ctx->file_region_cap);
  SS_REQUIRE(valid_mode(mode), SS_E_ARG, "unknown classifier mode %d", mode);
  cudaStream_t cs = ctx->compute_stream;
  const int64_t W = plan_windows(n_samples);
  rc = enqueue_detect_host(ctx, pcm_host, fmt, n_samples, mode, ctx->file_regions, ctx->file_nreg, cap);
  if (rc) return rc;
  int32_t nreg = 0;
  SS_CUDA_CHECK(cudaMemcpyAsync(&nreg, ctx->file_nreg, sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
  SS_CUDA_CHECK(cudaStreamSynchronize(cs));
  if ((rc = check_tc_health(ctx, mode, cs))) return rc;
  *n_regions = nreg;
  const int ncopy = nreg < cap ? nreg : cap;
  if (ncopy > 0)
    SS_CUDA_CHECK(cudaMemcpyAsync(regions_host, ctx->file_regions, (size_t)ncopy * 2 * sizeof(int32_t),
                                  cudaMemcpyDeviceToHost, cs));
  if (logits_host && W > 0)
    SS_CUDA_CHECK(cudaMemcpyAsync(logits_host, ctx->file_logits, (size_t)W * kFrames * sizeof(float),
                                  cudaMemcpyDeviceToHost, cs));
  SS_CUDA_CHECK(cudaStreamSynchronize(cs));
  return SS_OK;
}

int ss_detect_host(ss_ctx* ctx, const float* pcm_host, int64_t n_samples, int mode, int32_t* regions_host, int cap,
                   int* n_regions, float* logits_host) {
  return detect_host_impl(ctx, pcm_host, kSampleF32, n_samples, mode, regions_host, cap, n_regions, logits_host);
}

int ss_detect_host_pcm16(ss_ctx* ctx, const int16_t* pcm_host, int64_t n_samples, int mode, int32_t* regions_host,
                         int cap, int* n_regions, float* logits_host) {
  return detect_host_impl(ctx, pcm_host, kSampleS16, n_samples, mode, regions_host, cap, n_regions, logits_host);
}

static int detect_host_batch_impl(ss_ctx* ctx, int n_clips, const void* const* pcm_host, int fmt,
                                  const int64_t* n_samples, int mode, int32_t* regions_host, int cap, int* n_regions) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_clips >= 0 && cap >= 0 && (n_clips == 0 || (pcm_host && n_samples && n_regions)) &&
                 (regions_host || cap == 0 || n_clips == 0),
             SS_E_ARG, "bad ss_detect_host_batch arguments");
  SS_REQUIRE(valid_mode(mode), SS_E_ARG, "unknown classifier mode %d", mode);
  for (int i = 0; i < n_clips; ++i) {
    SS_REQUIRE(n_samples[i] >= 0 && (pcm_host[i] || n_samples[i] == 0), SS_E_ARG, "clip %d: bad buffer", i);
    SS_REQUIRE(ctx->file_logits && n_samples[i] <= ctx->file_cap_samples && cap <= ctx->file_region_cap, SS_E_CAPACITY,
               "clip %d of %lld samples / %d regions exceeds the reservation (%lld / %d): call ss_ctx_reserve", i,
               (long long)n_samples[i], cap, (long long)ctx->file_cap_samples, ctx->file_region_cap);
  }
  cudaStream_t cs = ctx->compute_stream;
  const size_t slot_ints = (size_t)ctx->file_region_cap * 2;
  for (int g0 = 0; g0 < n_clips; g0 += kBatchSlots) {
    const int g1 = (g0 + kBatchSlots < n_clips) ? g0 + kBatchSlots : n_clips;
    // enqueue the whole group: clip k+1's upload overlaps clip k's compute; results land in pinned host slots
    for (int i = g0; i < g1; ++i) {
      const int slot = i - g0;
      int32_t* reg_dev = ctx->file_regions + (size_t)slot * slot_ints;
      int32_t* host_slot = ctx->slot_host + (size_t)slot * (slot_ints + 1);
      rc = enqueue_detect_host(ctx, pcm_host[i], fmt, n_samples[i], mode, reg_dev, ctx->slot_nreg + slot, cap);
      if (rc) return rc;
      SS_CUDA_CHECK(cudaMemcpyAsync(host_slot, ctx->slot_nreg + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
      if (cap > 0)
        SS_CUDA_CHECK(cudaMemcpyAsync(host_slot + 1, reg_dev, (size_t)cap * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, cs));
    }
    SS_CUDA_CHECK(cudaStreamSynchronize(cs));
    if ((rc = check_tc_health(ctx, mode, cs))) return rc;
    for (int i = g0; i < g1; ++i) {
      const int32_t* host_slot = ctx->slot_host + (size_t)(i - g0) * (slot_ints + 1);
      const int nreg = host_slot[0];
      n_regions[i] = nreg;
      const int ncopy = nreg < cap ? nreg : cap;
      if (ncopy > 0) memcpy(regions_host + (size_t)i * cap * 2, host_slot + 1, (size_t)ncopy * 2 * sizeof(int32_t));
    }
  }
  return SS_OK;
}

int ss_detect_host_batch(ss_ctx* ctx, int n_clips, const float* const* pcm_host, const int64_t* n_samples, int mode,
                         int32_t* regions_host, int cap, int* n_regions) {
  return detect_host_batch_impl(ctx, n_clips, reinterpret_cast<const void* const*>(pcm_host), kSampleF32, n_samples,
                                mode, regions_host, cap, n_regions);
}

int ss_detect_host_batch_pcm16(ss_ctx* ctx, int n_clips, const int16_t* const* pcm_host, const int64_t* n_samples,
                               int mode, int32_t* regions_host, int cap, int* n_regions) {
  return detect_host_batch_impl(ctx, n_clips, reinterpret_cast<const void* const*>(pcm_host), kSampleS16, n_samples,
                                mode, regions_host, cap, n_regions);
}

int ss_decode_pcm16(ss_ctx* ctx, const int16_t* interleaved_dev, int64_t n_frames, int channels, float* mono_dev,
                    void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_frames >= 0 && channels >= 1 && channels <= 256, SS_E_ARG, "bad ss_decode_pcm16 arguments (%lld frames, %d channels)",
             (long long)n_frames, channels);
  SS_REQUIRE((interleaved_dev && mono_dev) || n_frames == 0, SS_E_ARG, "null pointer");
  return launch_decode_pcm16(interleaved_dev, n_frames, channels, mono_dev, static_cast<cudaStream_t>(stream));
}

int ss_encode_pcm16(ss_ctx* ctx, const float* src_dev, int64_t n_elems, int16_t* dst_dev, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_elems >= 0 && ((src_dev && dst_dev) || n_elems == 0), SS_E_ARG, "bad ss_encode_pcm16 arguments");
  return launch_encode_pcm16(src_dev, n_elems, dst_dev, static_cast<cudaStream_t>(stream));
}

int ss_silence_pcm16(ss_ctx* ctx, int16_t* pcm_dev, int64_t n_elems, const ss_interval* intervals_dev, int n_intervals,
                     int requantize, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_elems >= 0 && n_intervals >= 0, SS_E_ARG, "negative size");
  if (n_elems == 0) return SS_OK;
  SS_REQUIRE(pcm_dev && (intervals_dev || n_intervals == 0), SS_E_ARG, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (requantize && (rc = launch_requant_pcm16(pcm_dev, n_elems, st))) return rc;
  return launch_silence_s16(pcm_dev, n_elems, 0, intervals_dev, n_intervals, st);
}

int ss_silence_pcm16_host(ss_ctx* ctx, int16_t* pcm_host, int64_t n_elems, const ss_interval* intervals_host,
                          int n_intervals, int requantize) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_elems >= 0 && n_intervals >= 0, SS_E_ARG, "negative size");
  if (n_elems == 0 || (n_intervals == 0 && !requantize)) return SS_OK;
  SS_REQUIRE(pcm_host && (intervals_host || n_intervals == 0), SS_E_ARG, "null pointer");
  SS_REQUIRE(ctx->intervals, SS_E_CAPACITY, "call ss_ctx_reserve before ss_silence_pcm16_host");
  SS_REQUIRE((size_t)n_intervals * sizeof(ss_interval) <= (size_t)kIntervalCap * sizeof(float), SS_E_CAPACITY,
             "%d intervals exceed the table capacity", n_intervals);
  cudaStream_t cs = ctx->compute_stream;
  ss_interval* iv = reinterpret_cast<ss_interval*>(ctx->intervals);
  if (n_intervals > 0)
    SS_CUDA_CHECK(cudaMemcpyAsync(iv, intervals_host, (size_t)n_intervals * sizeof(ss_interval), cudaMemcpyHostToDevice, cs));
  // the float-sized staging buffer holds twice as many int16 samples (kept a multiple of 8 for the vector kernels)
  const int64_t cap = (ctx->stage_cap * 2) & ~(int64_t)7;
  int16_t* stage = reinterpret_cast<int16_t*>(ctx->stage_buf[0]);
  for (int64_t c0 = 0; c0 < n_elems; c0 += cap) {
    const int64_t len = (n_elems - c0 < cap) ? (n_elems - c0) : cap;
    SS_CUDA_CHECK(cudaMemcpyAsync(stage, pcm_host + c0, (size_t)len * sizeof(int16_t), cudaMemcpyHostToDevice, cs));
    if (requantize && (rc = launch_requant_pcm16(stage, len, cs))) return rc;
    if ((rc = launch_silence_s16(stage, len, c0, iv, n_intervals, cs))) return rc;
    SS_CUDA_CHECK(cudaMemcpyAsync(pcm_host + c0, stage, (size_t)len * sizeof(int16_t), cudaMemcpyDeviceToHost, cs));
  }
  SS_CUDA_CHECK(cudaStreamSynchronize(cs));
  return SS_OK;
}

int ss_check_health(ss_ctx* ctx, void* stream) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  return check_tc_health(ctx, SS_MODE_F16X3, static_cast<cudaStream_t>(stream));
}

int ss_silence_host(ss_ctx* ctx, float* pcm_host, int64_t n_elems, const ss_interval* intervals_host,
                    int n_intervals) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(n_elems >= 0 && n_intervals >= 0, SS_E_ARG, "negative size");
  if (n_elems == 0 || n_intervals == 0) return SS_OK;
  SS_REQUIRE(pcm_host && intervals_host, SS_E_ARG, "null pointer");
  SS_REQUIRE(ctx->intervals, SS_E_CAPACITY, "call ss_ctx_reserve before ss_silence_host");
  SS_REQUIRE((size_t)n_intervals * sizeof(ss_interval) <= (size_t)kIntervalCap * sizeof(float), SS_E_CAPACITY,
             "%d intervals exceed the table capacity", n_intervals);
  cudaStream_t cs = ctx->compute_stream;
  ss_interval* iv = reinterpret_cast<ss_interval*>(ctx->intervals);
  SS_CUDA_CHECK(cudaMemcpyAsync(iv, intervals_host, (size_t)n_intervals * sizeof(ss_interval), cudaMemcpyHostToDevice, cs));
  // round trip through the staging buffer, one chunk at a time; interval offsets are shifted per chunk
  for (int64_t c0 = 0; c0 < n_elems; c0 += ctx->stage_cap) {
    const int64_t len = (n_elems - c0 < ctx->stage_cap) ? (n_elems - c0) : ctx->stage_cap;
    SS_CUDA_CHECK(cudaMemcpyAsync(ctx->stage_buf[0], pcm_host + c0, (size_t)len * sizeof(float), cudaMemcpyHostToDevice, cs));
    rc = launch_silence(ctx->stage_buf[0], len, c0, iv, n_intervals, cs);
    if (rc) return rc;
    SS_CUDA_CHECK(cudaMemcpyAsync(pcm_host + c0, ctx->stage_buf[0], (size_t)len * sizeof(float), cudaMemcpyDeviceToHost, cs));
  }
  SS_CUDA_CHECK(cudaStreamSynchronize(cs));
  return SS_OK;
}

}  // extern "C"

extern "C" int ss_debug_activation(ss_ctx* ctx, int which, int n_windows, float* out_dev, int* C, int* H, int* W,
                                   void* stream) {
  int rc = ss::check_ctx_public(ctx);
  if (rc) return rc;
  SS_REQUIRE(C && H && W && n_windows >= 0, SS_E_ARG, "bad ss_debug_activation arguments");
  return ss::tc_debug_dump(ctx, which, n_windows, out_dev, C, H, W, static_cast<cudaStream_t>(stream));
}

extern "C" int ss_debug_tc_profile(ss_ctx* ctx, int select_launch, long long* out_host) {
  int rc = ss::check_ctx_public(ctx);
  if (rc) return rc;
  return ss::tc_debug_profile(ctx, select_launch, out_host);
}
