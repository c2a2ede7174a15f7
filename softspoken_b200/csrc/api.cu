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

void register_guard(ss_ctx* ctx, const void* ptr, size_t bytes, const void* owner) {
  ctx->guards.push_back(GuardBand{static_cast<const unsigned char*>(ptr), bytes, owner});
}

void unregister_guards(ss_ctx* ctx, const void* owner) {
  size_t k = 0;
  for (size_t i = 0; i < ctx->guards.size(); ++i)
    if (ctx->guards[i].owner != owner) ctx->guards[k++] = ctx->guards[i];
  ctx->guards.resize(k);
}

// one CTA per guard band: count the bytes that no longer hold the pattern
__global__ void __launch_bounds__(256)
check_guards_kernel(const GuardBand* __restrict__ bands, unsigned long long* __restrict__ bad) {
  const GuardBand g = bands[blockIdx.x];
  const uint32_t want = 0x01010101u * kGuardPattern;
  unsigned int n = 0;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(g.ptr);        // bands are 256-byte aligned multiples of 256
  for (size_t i = threadIdx.x; i < g.bytes / 4; i += blockDim.x) {
    const uint32_t x = w[i] ^ want;
    n += ((x & 0xffu) != 0) + ((x & 0xff00u) != 0) + ((x & 0xff0000u) != 0) + ((x & 0xff000000u) != 0);
  }
  if (n) atomicAdd(bad, (unsigned long long)n);
}

namespace {

constexpr uint32_t kBlobMagic = 0x53534232u;   // 'SSB2' (softspoken_b200/checkpoint.py)
constexpr uint32_t kBlobVersion = 2;
constexpr int64_t kChunkWindows = 1024;        // windows per streamed chunk in ss_detect_*
constexpr int kIntervalCap = 1 << 20;
constexpr int kBatchSlots = 8;                 // clips in flight inside ss_detect_host_batch
constexpr int kRefineBatch = 128;              // windows per pass of the margin-guided refinement
constexpr int kRefineFirst = 4096;             // list entries fetched together with the count
constexpr double kDefaultRefineEps = 0.0;      // off: the first pass is in the reference's own noise class (ss_ctx_set_refine)

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

// Every device allocation of the context sits between two guard bands filled with kGuardPattern; ss_debug_check_guards
// counts the guard bytes that no longer hold it (compute-sanitizer is not available on the B200 pool, so this is the
// out-of-bounds-write net of the test suite).
template <typename T>
int dev_alloc(ss_ctx* ctx, T** p, size_t count) {
  const size_t body = (count * sizeof(T) + 255) & ~(size_t)255;
  unsigned char* base = nullptr;
  SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&base), body + 2 * kCtxGuardBytes));
  SS_CUDA_CHECK(cudaMemset(base, kGuardPattern, kCtxGuardBytes));
  SS_CUDA_CHECK(cudaMemset(base + kCtxGuardBytes + body, kGuardPattern, kCtxGuardBytes));
  *p = reinterpret_cast<T*>(base + kCtxGuardBytes);
  ctx->allocs[*p] = base;
  register_guard(ctx, base, kCtxGuardBytes, base);
  register_guard(ctx, base + kCtxGuardBytes + body, kCtxGuardBytes, base);
  ctx->device_bytes += body + 2 * kCtxGuardBytes;
  return SS_OK;
}

void dev_free(ss_ctx* ctx, const void* p) {
  if (!p) return;
  auto it = ctx->allocs.find(p);
  if (it == ctx->allocs.end()) return;
  unsigned char* base = static_cast<unsigned char*>(it->second);
  ctx->allocs.erase(it);
  unregister_guards(ctx, base);
  cudaFree(base);
}

// K1's two-band walk of the filterbank (FrontEnd::mel_rec).  Possible when every band is non-empty, band starts and
// ends are non-decreasing and band m + 2 starts after band m ends — every bin then lies in at most two consecutive
// bands, one even- and one odd-numbered.  The bands are cut into kFeatureWarps contiguous groups whose widest bin span
// is minimal; a group's records cover the bins from its first band's start to its last band's end and carry the
// weights of its own bands only.  Returns SS_OK with n_rec = 0 when the bank does not have that shape.
int build_mel_walk(ss_ctx* ctx, const int* ms, const int* mc, const int* mo, const float* taps) {
  ctx->fe.n_rec = 0;
  std::vector<int> lo(kMels), hi(kMels);
  for (int m = 0; m < kMels; ++m) {
    if (mc[m] <= 0) return SS_OK;
    lo[m] = ms[m];
    hi[m] = ms[m] + mc[m] - 1;
    if (m >= 1 && (lo[m] < lo[m - 1] || hi[m] < hi[m - 1])) return SS_OK;
    if (m >= 2 && lo[m] <= hi[m - 2]) return SS_OK;
  }
  auto groups_for = [&](int span, std::vector<int>* first) {
    first->clear();
    for (int m = 0; m < kMels;) {
      first->push_back(m);
      int k = m;
      while (k + 1 < kMels && hi[k + 1] - lo[m] + 1 <= span) ++k;
      m = k + 1;
    }
    return (int)first->size();
  };
  int span = 0;
  for (int m = 0; m < kMels; ++m) span = mc[m] > span ? mc[m] : span;
  std::vector<int> first;
  while (groups_for(span, &first) > kFeatureWarps) ++span;
  std::vector<float4> rec;
  std::vector<int> begin(kFeatureWarps + 1, 0);
  for (int g = 0; g < kFeatureWarps; ++g) {
    begin[g] = (int)rec.size();
    if (g >= (int)first.size()) continue;
    const int ja = first[g], jb = g + 1 < (int)first.size() ? first[g + 1] : kMels;
    for (int k = lo[ja]; k <= hi[jb - 1]; ++k) {
      float w[2] = {0.f, 0.f};
      int emit = 0;
      for (int m = ja; m < jb; ++m) {
        if (k < lo[m] || k > hi[m]) continue;
        w[m & 1] = taps[mo[m] + (k - lo[m])];
        if (k == hi[m]) emit |= (m + 1) << (8 * (m & 1));
      }
      float4 r;
      r.x = w[0];
      r.y = w[1];
      memcpy(&r.z, &k, 4);
      memcpy(&r.w, &emit, 4);
      rec.push_back(r);
    }
  }
  begin[kFeatureWarps] = (int)rec.size();
  if ((int)rec.size() > kMaxMelRec) return SS_OK;
  float4* d_rec = nullptr;
  int* d_begin = nullptr;
  int rc = dev_alloc(ctx, &d_rec, rec.size());
  if (rc) return rc;
  if ((rc = dev_alloc(ctx, &d_begin, begin.size()))) return rc;
  SS_CUDA_CHECK(cudaMemcpy(d_rec, rec.data(), rec.size() * sizeof(float4), cudaMemcpyHostToDevice));
  SS_CUDA_CHECK(cudaMemcpy(d_begin, begin.data(), begin.size() * sizeof(int), cudaMemcpyHostToDevice));
  ctx->fe.mel_rec = d_rec;
  ctx->fe.mel_rec_begin = d_begin;
  ctx->fe.n_rec = (int)rec.size();
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

// Where the samples of the clip in flight can be read back from while its flagged windows are refined.
struct ClipSource {
  const void* base;        // element 0 = unpadded sample `first`
  int64_t first, last;     // unpadded samples [first, last) are present at `base`
  cudaMemcpyKind kind;     // cudaMemcpyDeviceToDevice (resident clip, staging buffer) or cudaMemcpyHostToDevice
  int fmt;                 // kSampleF32 / kSampleS16
};

int classify_any(ss_ctx* ctx, int mode, const float* mel, int n, float* logits, cudaStream_t st) {
  if (mode == SS_MODE_FP32) return classify_fp32(ctx, mel, n, logits, nullptr, st);
  return classify_tc(ctx, mode, mel, n, logits, nullptr, st);
}

bool refine_active(const ss_ctx* ctx, int mode) {
  return ctx->refine_eps > 0.0 && mode != ctx->refine_mode && ctx->win_flags != nullptr;
}

// (Re)size the refinement scratch to the file reservation.  Called from ss_ctx_reserve / ss_ctx_set_refine only.
int reserve_refine(ss_ctx* ctx) {
  if (!(ctx->refine_eps > 0.0) || !ctx->file_logits) return SS_OK;
  int rc;
  const int64_t W = ctx->file_cap_windows > 0 ? ctx->file_cap_windows : 1;
  if (ctx->refine_cap_windows < W) {
    SS_CUDA_CHECK(cudaDeviceSynchronize());
    if (ctx->win_flags) dev_free(ctx, ctx->win_flags);
    if (ctx->refine_list) dev_free(ctx, ctx->refine_list);
    if (ctx->refine_host) cudaFreeHost(ctx->refine_host);
    ctx->win_flags = nullptr; ctx->refine_list = nullptr; ctx->refine_host = nullptr;
    if ((rc = dev_alloc(ctx, &ctx->win_flags, (size_t)W))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->refine_list, (size_t)W + 1))) return rc;
    SS_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&ctx->refine_host), ((size_t)W + 1) * sizeof(int32_t)));
    ctx->refine_cap_windows = W;
  }
  if (!ctx->refine_raw) {
    float* raw = nullptr;
    if ((rc = dev_alloc(ctx, &raw, (size_t)kRefineBatch * kWindowSamplesUsed))) return rc;
    ctx->refine_raw = raw;
    if ((rc = dev_alloc(ctx, &ctx->refine_starts, (size_t)kRefineBatch))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->refine_logits, (size_t)kRefineBatch * kFrames))) return rc;
    int64_t starts[kRefineBatch];
    for (int k = 0; k < kRefineBatch; ++k) starts[k] = (int64_t)k * kWindowSamplesUsed;
    SS_CUDA_CHECK(cudaMemcpy(ctx->refine_starts, starts, sizeof(starts), cudaMemcpyHostToDevice));
  }
  if (ctx->refine_mode == SS_MODE_FP32 && (rc = ensure_workspace_f32(ctx))) return rc;
  return SS_OK;
}

// K5 (+ refinement) + K6 of one clip whose first-pass logits are in ctx->file_logits.
//
// Margin-guided refinement.  The tensor-core classifier's logits carry ~1e-5 of rounding noise (operand
// precisions), the reference's own float32 a few 1e-6, and a detection is a comparison `avg > 0.1`: a bin whose
// average lies inside the noise band can land on the other side of the threshold.  K5 therefore marks the (at most
// five) windows covering every bin with |avg - 0.1| < refine_eps; the marked windows are gathered (their 65,536
// samples each, from wherever the clip still is), run through K1 and the classifier again in refine_mode and
// scattered over their first-pass logits; K5 runs once more and K6 sees only the refined decisions.  Costs one
// stream synchronisation per clip (the host needs the count) plus the second pass over the marked windows.
int finish_clip(ss_ctx* ctx, const ClipSource& src, int64_t n_samples, int64_t W, int mode, int32_t* regions_dev,
                int32_t* nreg_dev, int cap, cudaStream_t st) {
  const int64_t bins = timeline_bins(n_samples + 2 * (int64_t)kPadSamples);
  ctx->stat_windows += (uint64_t)W;
  ctx->stat_clips += 1;
  if (!refine_active(ctx, mode) || W <= 0)
    return launch_average_regions(ctx->file_logits, (int)W, bins, ctx->file_avg, ctx->file_cnt, 0.1, kGapBins, regions_dev,
                                  nreg_dev, cap, ctx->scan_tmp, ctx->scan_tmp_len, st);
  SS_REQUIRE(W <= ctx->refine_cap_windows, SS_E_CAPACITY, "refinement scratch holds %lld windows, clip has %lld",
             (long long)ctx->refine_cap_windows, (long long)W);
  int rc;
  SS_CUDA_CHECK(cudaMemsetAsync(ctx->win_flags, 0, (size_t)W, st));
  if ((rc = launch_average_bits(ctx->file_logits, (int)W, bins, ctx->file_avg, ctx->file_cnt, 0.1, ctx->scan_tmp,
                                ctx->scan_tmp_len, ctx->refine_eps, ctx->win_flags, st))) return rc;
  if ((rc = launch_compact_flags(ctx->win_flags, (int)W, ctx->refine_list + 1, ctx->refine_list, st))) return rc;
  const int64_t first = W < kRefineFirst ? W : kRefineFirst;
  SS_CUDA_CHECK(cudaMemcpyAsync(ctx->refine_host, ctx->refine_list, (size_t)(1 + first) * sizeof(int32_t),
                                cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(cudaStreamSynchronize(st));
  const int64_t n_ref = ctx->refine_host[0];
  if (n_ref > first) {
    SS_CUDA_CHECK(cudaMemcpyAsync(ctx->refine_host + 1 + first, ctx->refine_list + 1 + first,
                                  (size_t)(n_ref - first) * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(cudaStreamSynchronize(st));
  }
  if (n_ref > 0) {
    ctx->stat_refined += (uint64_t)n_ref;
    ctx->stat_clips_refined += 1;
    const size_t esz = (src.fmt == kSampleS16) ? sizeof(int16_t) : sizeof(float);
    char* raw = static_cast<char*>(ctx->refine_raw);
    for (int64_t k0 = 0; k0 < n_ref; k0 += kRefineBatch) {
      const int nb = (int)((n_ref - k0 < kRefineBatch) ? (n_ref - k0) : kRefineBatch);
      for (int k = 0; k < nb; ++k) {
        // window w = padded samples [w * 13230, + 65536) = unpadded [a, a + 65536), zero outside the clip
        const int64_t w = ctx->refine_host[1 + k0 + k];
        const int64_t a = w * kStepSamples - kPadSamples;
        const int64_t lo = a < 0 ? 0 : a;
        const int64_t hi = (a + kWindowSamplesUsed < n_samples) ? a + kWindowSamplesUsed : n_samples;
        char* slot = raw + (size_t)k * kWindowSamplesUsed * esz;
        if (lo > a || hi < a + kWindowSamplesUsed)
          SS_CUDA_CHECK(cudaMemsetAsync(slot, 0, (size_t)kWindowSamplesUsed * esz, st));
        if (hi > lo) {
          SS_REQUIRE(lo >= src.first && hi <= src.last, SS_E_ARG,
                     "refinement: window %lld needs samples [%lld, %lld), the source holds [%lld, %lld)", (long long)w,
                     (long long)lo, (long long)hi, (long long)src.first, (long long)src.last);
          SS_CUDA_CHECK(cudaMemcpyAsync(slot + (size_t)(lo - a) * esz,
                                        static_cast<const char*>(src.base) + (size_t)(lo - src.first) * esz,
                                        (size_t)(hi - lo) * esz, src.kind, st));
        }
      }
      if ((rc = launch_features_virtual(ctx, ctx->refine_raw, src.fmt, 0, (int64_t)nb * kWindowSamplesUsed, 0,
                                        ctx->refine_starts, 0, nb, ctx->file_mel, st))) return rc;
      if ((rc = classify_any(ctx, ctx->refine_mode, ctx->file_mel, nb, ctx->refine_logits, st))) return rc;
      if ((rc = launch_scatter_rows(ctx->refine_logits, ctx->refine_list + 1 + k0, nb, ctx->file_logits, st))) return rc;
    }
    if ((rc = launch_average_bits(ctx->file_logits, (int)W, bins, ctx->file_avg, ctx->file_cnt, 0.1, ctx->scan_tmp,
                                  ctx->scan_tmp_len, 0.0, nullptr, st))) return rc;
  }
  return launch_regions_after_bits(bins, kGapBins, regions_dev, nreg_dev, cap, ctx->scan_tmp, ctx->scan_tmp_len, st);
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
  ctx->refine_eps = kDefaultRefineEps;
  ctx->refine_mode = SS_MODE_FP32;
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
  {
    // SS_MEL_WALK=0 keeps K1's band-by-band walk of the sparse taps (A/B runs, and the path of non-triangular banks)
    const char* mw = getenv("SS_MEL_WALK");
    if (mw == nullptr || atoi(mw) != 0)
      FAIL_IF(build_mel_walk(ctx, reinterpret_cast<const int*>(v.payload + v.entries.at("mel_start").off),
                             reinterpret_cast<const int*>(v.payload + v.entries.at("mel_count").off),
                             reinterpret_cast<const int*>(v.payload + v.entries.at("mel_offs").off),
                             v.payload + v.entries.at("mel_taps").off));
  }
  {
    // SS_K1_PACKED=0 keeps K1's scalar phase 1 (A/B runs and the packed-vs-scalar bit comparison of the tests)
    const char* pk = getenv("SS_K1_PACKED");
    ctx->fe.packed = (pk == nullptr || atoi(pk) != 0) ? 1 : 0;
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
  for (float* p : bufs) if (p) dev_free(ctx, p);
  if (ctx->file_avg) dev_free(ctx, ctx->file_avg);
  if (ctx->file_cnt) dev_free(ctx, ctx->file_cnt);
  if (ctx->file_regions) dev_free(ctx, ctx->file_regions);
  if (ctx->file_nreg) dev_free(ctx, ctx->file_nreg);
  if (ctx->slot_nreg) dev_free(ctx, ctx->slot_nreg);
  if (ctx->slot_host) cudaFreeHost(ctx->slot_host);
  if (ctx->scan_tmp) dev_free(ctx, ctx->scan_tmp);
  if (ctx->intervals) dev_free(ctx, ctx->intervals);
  if (ctx->win_flags) dev_free(ctx, ctx->win_flags);
  if (ctx->refine_list) dev_free(ctx, ctx->refine_list);
  if (ctx->refine_host) cudaFreeHost(ctx->refine_host);
  if (ctx->refine_raw) dev_free(ctx, ctx->refine_raw);
  if (ctx->refine_starts) dev_free(ctx, ctx->refine_starts);
  if (ctx->refine_logits) dev_free(ctx, ctx->refine_logits);
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
    if (ctx->file_logits) { dev_free(ctx, ctx->file_logits); dev_free(ctx, ctx->file_avg); dev_free(ctx, ctx->file_cnt); dev_free(ctx, ctx->scan_tmp); }
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
    if (ctx->file_regions) dev_free(ctx, ctx->file_regions);
    if (ctx->slot_nreg) dev_free(ctx, ctx->slot_nreg);
    if (ctx->slot_host) cudaFreeHost(ctx->slot_host);
    ctx->file_regions = nullptr; ctx->slot_nreg = nullptr; ctx->slot_host = nullptr;
    // kBatchSlots region buffers (slot 0 doubles as the single-clip buffer) + counters + a pinned host mirror
    if ((rc = dev_alloc(ctx, &ctx->file_regions, (size_t)kBatchSlots * region_cap * 2))) return rc;
    if ((rc = dev_alloc(ctx, &ctx->slot_nreg, (size_t)kBatchSlots))) return rc;
    SS_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&ctx->slot_host),
                                 (size_t)kBatchSlots * ((size_t)region_cap * 2 + 1) * sizeof(int32_t)));
    ctx->file_region_cap = region_cap;
  }
  return reserve_refine(ctx);
}

int ss_ctx_set_refine(ss_ctx* ctx, double eps, int refine_mode) {
  int rc = check_ctx(ctx);
  if (rc) return rc;
  SS_REQUIRE(eps >= 0.0 && eps < 0.1, SS_E_ARG, "refinement margin %g out of [0, 0.1)", eps);
  SS_REQUIRE(valid_mode(refine_mode), SS_E_ARG, "unknown classifier mode %d", refine_mode);
  ctx->refine_eps = eps;
  ctx->refine_mode = refine_mode;
  return reserve_refine(ctx);
}

int ss_ctx_refine_stats(ss_ctx* ctx, uint64_t* stats4, int reset) {
  SS_REQUIRE(ctx && stats4, SS_E_ARG, "null argument");
  stats4[0] = ctx->stat_windows; stats4[1] = ctx->stat_refined; stats4[2] = ctx->stat_clips; stats4[3] = ctx->stat_clips_refined;
  if (reset) ctx->stat_windows = ctx->stat_refined = ctx->stat_clips = ctx->stat_clips_refined = 0;
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
  const ClipSource src{pcm_dev, 0, n_samples, cudaMemcpyDeviceToDevice, fmt};
  rc = finish_clip(ctx, src, n_samples, W, mode, regions_dev, n_regions_dev, cap, st);
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

// One streamed chunk of a host clip: windows [w0, w1) need the unpadded samples [s0, s1), staged in stage_buf[buf].
struct Chunk {
  int64_t w0 = 0, w1 = 0, s0 = 0, s1 = 0;
  int buf = 0;
  bool valid = false;
};

static Chunk plan_chunk(const ss_ctx* ctx, int64_t n_samples, int64_t W, int64_t w0) {
  Chunk c;
  c.w0 = w0;
  c.w1 = (w0 + ctx->chunk_windows < W) ? w0 + ctx->chunk_windows : W;
  // padded sample range the chunk's kept frames touch: [w0*step - 256 (reflection stays >= w0*step), ...)
  const int64_t plo = c.w0 * kStepSamples, phi = (c.w1 - 1) * kStepSamples + kWindowSamplesUsed;
  c.s0 = plo - kPadSamples;       // unpadded coordinates
  c.s1 = phi - kPadSamples;
  if (c.s0 < 0) c.s0 = 0;
  if (c.s1 > n_samples) c.s1 = n_samples;
  if (c.s1 < c.s0) c.s1 = c.s0;
  return c;
}

// H2D of a chunk on the copy stream into the next staging buffer (waits until that buffer's last reader is done).
static int upload_chunk(ss_ctx* ctx, const void* pcm_host, int fmt, Chunk* c) {
  const size_t esz = (fmt == kSampleS16) ? sizeof(int16_t) : sizeof(float);   // the staging buffers hold either type
  cudaStream_t xs = ctx->copy_stream;
  c->buf = ctx->stage_next;
  ctx->stage_next ^= 1;
  SS_CUDA_CHECK(cudaStreamWaitEvent(xs, ctx->ev_consumed[c->buf], 0));   // staging buffer free again
  if (c->s1 > c->s0)
    SS_CUDA_CHECK(cudaMemcpyAsync(ctx->stage_buf[c->buf], static_cast<const char*>(pcm_host) + (size_t)c->s0 * esz,
                                  (size_t)(c->s1 - c->s0) * esz, cudaMemcpyHostToDevice, xs));
  SS_CUDA_CHECK(cudaEventRecord(ctx->ev_copied[c->buf], xs));
  c->valid = true;
  return SS_OK;
}

// One host clip: chunked H2D on the copy stream (double-buffered staging), K1-K3 per chunk, then K5 (+ refinement)
// and K6 on the compute stream, regions left in `regions_dev` / `nreg_dev`.  `pre`: the clip's first chunk if a
// previous call already uploaded it.  `next_pcm`: the clip that follows (or null) — its first chunk is uploaded
// before this clip's K5 synchronises the host, so that the copy engine keeps working; returned in `next_pre`.
static int detect_host_clip(ss_ctx* ctx, const void* pcm_host, int fmt, int64_t n_samples, int mode,
                            int32_t* regions_dev, int32_t* nreg_dev, int cap, const Chunk* pre,
                            const void* next_pcm, int64_t next_n, Chunk* next_pre) {
  cudaStream_t cs = ctx->compute_stream;
  const int64_t W = plan_windows(n_samples);
  int rc;
  Chunk cur;
  for (int64_t w0 = 0; w0 < W; w0 += ctx->chunk_windows) {
    if (w0 == 0 && pre && pre->valid) {
      cur = *pre;
    } else {
      cur = plan_chunk(ctx, n_samples, W, w0);
      if ((rc = upload_chunk(ctx, pcm_host, fmt, &cur))) return rc;
    }
    SS_CUDA_CHECK(cudaStreamWaitEvent(cs, ctx->ev_copied[cur.buf], 0));
    rc = run_windows(ctx, ctx->stage_buf[cur.buf], fmt, kPadSamples + cur.s0, kPadSamples + cur.s1, kPadSamples + cur.s0,
                     cur.w0, cur.w1, mode, ctx->file_logits, cs);
    if (rc) return rc;
    SS_CUDA_CHECK(cudaEventRecord(ctx->ev_consumed[cur.buf], cs));
  }
  if (next_pre) {
    *next_pre = Chunk{};
    const int64_t Wn = next_pcm ? plan_windows(next_n) : 0;
    if (Wn > 0) {
      *next_pre = plan_chunk(ctx, next_n, Wn, 0);
      if ((rc = upload_chunk(ctx, next_pcm, fmt, next_pre))) return rc;
    }
  }
  // the flagged windows' samples: a clip of one chunk is still whole in its staging buffer, a longer one is read
  // back from the host buffer
  ClipSource src{pcm_host, 0, n_samples, cudaMemcpyHostToDevice, fmt};
  const bool staged = W > 0 && W <= ctx->chunk_windows;
  if (staged) src = ClipSource{ctx->stage_buf[cur.buf], cur.s0, cur.s1, cudaMemcpyDeviceToDevice, fmt};
  rc = finish_clip(ctx, src, n_samples, W, mode, regions_dev, nreg_dev, cap, cs);
  if (rc) return rc;
  if (staged) SS_CUDA_CHECK(cudaEventRecord(ctx->ev_consumed[cur.buf], cs));   // the refinement read the buffer again
  return SS_OK;
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
  rc = detect_host_clip(ctx, pcm_host, fmt, n_samples, mode, ctx->file_regions, ctx->file_nreg, cap, nullptr, nullptr, 0,
                        nullptr);
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
  Chunk pre;          // first chunk of the next clip, uploaded ahead
  for (int g0 = 0; g0 < n_clips; g0 += kBatchSlots) {
    const int g1 = (g0 + kBatchSlots < n_clips) ? g0 + kBatchSlots : n_clips;
    // clip k+1's upload overlaps clip k's compute (its first chunk is enqueued before clip k's K5 synchronises the
    // host for the refinement count); results land in pinned host slots and are harvested once per group
    for (int i = g0; i < g1; ++i) {
      const int slot = i - g0;
      int32_t* reg_dev = ctx->file_regions + (size_t)slot * slot_ints;
      int32_t* host_slot = ctx->slot_host + (size_t)slot * (slot_ints + 1);
      const Chunk mine = pre;
      rc = detect_host_clip(ctx, pcm_host[i], fmt, n_samples[i], mode, reg_dev, ctx->slot_nreg + slot, cap, &mine,
                            i + 1 < n_clips ? pcm_host[i + 1] : nullptr, i + 1 < n_clips ? n_samples[i + 1] : 0, &pre);
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

extern "C" int ss_debug_check_guards(ss_ctx* ctx, uint64_t* bad_bytes, int* n_bands) {
  int rc = ss::check_ctx_public(ctx);
  if (rc) return rc;
  SS_REQUIRE(bad_bytes, SS_E_ARG, "null argument");
  *bad_bytes = 0;
  if (n_bands) *n_bands = (int)ctx->guards.size();
  if (ctx->guards.empty()) return SS_OK;
  SS_CUDA_CHECK(cudaDeviceSynchronize());
  ss::GuardBand* bands = nullptr;
  unsigned long long* bad = nullptr;
  SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&bands), ctx->guards.size() * sizeof(ss::GuardBand) + 8));
  bad = reinterpret_cast<unsigned long long*>(bands + ctx->guards.size());
  SS_CUDA_CHECK(cudaMemcpy(bands, ctx->guards.data(), ctx->guards.size() * sizeof(ss::GuardBand), cudaMemcpyHostToDevice));
  SS_CUDA_CHECK(cudaMemset(bad, 0, 8));
  ss::check_guards_kernel<<<(int)ctx->guards.size(), 256>>>(bands, bad);
  unsigned long long h = 0;
  cudaError_t e = cudaMemcpy(&h, bad, 8, cudaMemcpyDeviceToHost);
  cudaFree(bands);
  if (e != cudaSuccess) { ss::set_error("guard check failed: %s", cudaGetErrorString(e)); return SS_E_CUDA; }
  *bad_bytes = h;
  return SS_OK;
}

extern "C" int ss_debug_tc_profile(ss_ctx* ctx, int select_launch, long long* out_host) {
  int rc = ss::check_ctx_public(ctx);
  if (rc) return rc;
  return ss::tc_debug_profile(ctx, select_launch, out_host);
}
