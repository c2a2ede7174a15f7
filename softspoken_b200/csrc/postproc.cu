// K5 / K6 — overlap averaging and speech-region finding on the 256/3 Hz timeline.
//
// K5 replaces NNDetector.average_overlapping_detections (root/code/frontend/NNDetector.py:168-186).
//   The reference scatters window i's 256 logits to sum[p_i : p_i+256] in window order, in float64.
//   Here each output bin GATHERS its <= 5 covering windows and adds them in ascending i, float32 ->
//   float64, then divides by the float64 count: the same operations in the same order, so the result
//   is bit-identical, with no atomics.  p_i = int(round(i * 0.6 / (3/256))) == (256 i + 2) / 5
//   (integer form; equality proven in tests/test_oracle_postproc.py).
// K6 replaces NNDetector.find_speech_regions (NNDetector.py:109-141) in bin space.
//   hot[j] = count[j] >= 1 && avg[j] > threshold (strict, float64).  Runs of hot bins (end inclusive) are
//   merged when next_start - cur_end <= gap_bins, which is the reference's string-time rule
//   `float(next_start) - float(cur_end) <= 0.5` for gap_bins = 42.  Equivalently bin j starts a merged
//   region iff it is hot and no bin in [j-gap, j-1] is, and ends one iff no bin in [j+1, j+gap] is; the
//   k-th start pairs with the k-th end.  Two passes over the timeline: per-CTA counts, then a
//   warp-ballot scan assigns ranks and writes (start, end) pairs in order.  Uncovered bins exist only
//   past the last window (coverage is a prefix), where the reference emits no entries at all.
#include "ss_common.cuh"

namespace ss {

namespace {

constexpr int kAvgThreads = 256;
constexpr int kHotBins = 4;          // bins per thread of hot_bits_kernel

// One thread per timeline bin.  The windows covering bin j are i_lo .. i_hi with
//   i_hi = floor((5 j + 2) / 256)            (largest i with p_i <= j)
//   i_lo = ceil((5 (j - 255) - 2) / 256)     (smallest i with p_i + 255 >= j),  clamped to [0, n_windows - 1]:
// at most five (five consecutive window spacings add up to exactly 256 bins), so the gather is five predicated
// loads in flight and five float64 additions in ascending window order — no index array, no local memory.
// kBits: the thread also votes hot[j] = covered && avg > threshold and lane 0 stores the warp's 32-bit word, so
// that K6 never has to read the float64 timeline back.
// Index: uint32_t while 5 out_len + 2 fits (116 days of audio), else int64_t — the kernel is bound by instruction
// issue, not by HBM, and 64-bit divisions by 5 were most of its instructions; p_i is divided once for i_lo and
// carried forward (256 = 5 * 51 + 1: the quotient grows by 51, and by one more whenever the remainder wraps).
// flags (optional, with kBits): margin-guided refinement — a bin whose average lies within `eps` of the threshold
// marks its covering windows in flags[] (one byte per window); see ss_ctx_set_refine.
template <bool kBits, typename Index>
__global__ void __launch_bounds__(kAvgThreads)
average_kernel(const float* __restrict__ logits, int n_windows, int64_t out_len, double* __restrict__ avg,
               int32_t* __restrict__ cnt, double threshold, uint32_t* __restrict__ bits, double eps,
               unsigned char* __restrict__ flags) {
  const int64_t j64 = (int64_t)blockIdx.x * kAvgThreads + threadIdx.x;
  const bool in = j64 < out_len;
  bool hot = false;
  if (in) {
    const Index j = (Index)j64;
    const Index j5 = 5 * j;
    Index i_hi = (j5 + 2) >> 8;
    if (i_hi > (Index)(n_windows - 1)) i_hi = (Index)(n_windows - 1);
    const Index i_lo = j5 > 1277 ? (j5 - 1277 + 255) >> 8 : 0;       // ceil((5 (j - 255) - 2) / 256), clamped at 0
    const int n = n_windows > 0 && i_hi >= i_lo ? (int)(i_hi - i_lo) + 1 : 0;
    const Index t = 256 * i_lo + 2;
    Index p = t / 5;
    int r = (int)(t - 5 * p);
    const float* src = logits + (int64_t)i_lo * 256 + (int64_t)(j - p);      // window i_lo, frame j - p
    float v[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      v[k] = (k < n) ? __ldg(src) : 0.f;
      const int step = 51 + (r == 4);        // p_{i+1} - p_i
      r = (r == 4) ? 0 : r + 1;
      src += 256 - step;
    }
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < 5; ++k)
      if (k < n) sum += (double)v[k];                  // ascending window index, as the reference's `+=` sequence
    const double a = n ? sum / (double)n : __longlong_as_double(0x7ff8000000000000LL);
    cnt[j64] = n;
    avg[j64] = a;
    hot = n >= 1 && a > threshold;
    if (kBits && flags != nullptr && n >= 1 && fabs(a - threshold) < eps) {
#pragma unroll
      for (int k = 0; k < 5; ++k)
        if (k < n) flags[(int64_t)i_lo + k] = 1;
    }
  }
  if (kBits) {
    const unsigned word = __ballot_sync(0xffffffffu, hot);
    if ((threadIdx.x & 31) == 0 && in) bits[j64 >> 5] = word;      // bins past out_len vote 0
  }
}

template <bool kBits>
static void launch_average_kernel(const float* logits, int n_windows, int64_t out_len, double* avg, int32_t* cnt,
                                  double threshold, uint32_t* bits, cudaStream_t st, double eps = 0.0,
                                  unsigned char* flags = nullptr) {
  const int grid = (int)((out_len + kAvgThreads - 1) / kAvgThreads);
  if (5 * out_len + 2 < ((int64_t)1 << 32) && (int64_t)n_windows * 256 + 2 < ((int64_t)1 << 32))
    average_kernel<kBits, uint32_t><<<grid, kAvgThreads, 0, st>>>(logits, n_windows, out_len, avg, cnt, threshold, bits,
                                                                  eps, flags);
  else
    average_kernel<kBits, int64_t><<<grid, kAvgThreads, 0, st>>>(logits, n_windows, out_len, avg, cnt, threshold, bits,
                                                                 eps, flags);
}

// Refinement bookkeeping: flags[0 .. n) (one byte per window) -> ascending list of the flagged window indices and
// their number.  One CTA; ballot + block scan per 1,024 windows.
__global__ void __launch_bounds__(1024)
compact_flags_kernel(const unsigned char* __restrict__ flags, int n, int32_t* __restrict__ list,
                     int32_t* __restrict__ count) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const bool f = i < n && flags[i] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warp_tot[wid] = __popc(m);
    __syncthreads();
    int off = carry, tot = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < wid) off += warp_tot[w];
      tot += warp_tot[w];
    }
    if (f) list[off + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (tid == 0) carry += tot;
    __syncthreads();
  }
  if (tid == 0) *count = carry;
}

// logits[list[k]][0 .. 256) = rows[k][0 .. 256): the refined windows replace their first-pass logits
__global__ void __launch_bounds__(256)
scatter_rows_kernel(const float* __restrict__ rows, const int32_t* __restrict__ list, float* __restrict__ logits) {
  logits[(int64_t)list[blockIdx.x] * kFrames + threadIdx.x] = rows[(int64_t)blockIdx.x * kFrames + threadIdx.x];
}

// hot[j] = count[j] >= 1 && avg[j] > threshold as one bit per bin (the entry for callers that bring their own timeline)
__global__ void __launch_bounds__(kAvgThreads)
hot_bits_kernel(const double* __restrict__ avg, const int32_t* __restrict__ cnt, int64_t out_len, double threshold,
                uint32_t* __restrict__ bits) {
  // four bins per thread, kAvgThreads apart: eight independent loads in flight before the first vote
  const int64_t j0 = (int64_t)blockIdx.x * (kAvgThreads * kHotBins) + threadIdx.x;
  double a[kHotBins];
  int32_t c[kHotBins];
#pragma unroll
  for (int b = 0; b < kHotBins; ++b) {
    const int64_t j = j0 + b * kAvgThreads;
    const bool in = j < out_len;
    c[b] = in ? __ldg(cnt + j) : 0;
    a[b] = in ? __ldg(avg + j) : 0.0;
  }
#pragma unroll
  for (int b = 0; b < kHotBins; ++b) {
    const int64_t j = j0 + b * kAvgThreads;
    const unsigned word = __ballot_sync(0xffffffffu, c[b] >= 1 && a[b] > threshold);
    if ((threadIdx.x & 31) == 0 && j < out_len) bits[j >> 5] = word;
  }
}

constexpr int kScanThreads = 256;
constexpr int kBinsPerThread = 16;
constexpr int kTile = kScanThreads * kBinsPerThread;   // 4,096 bins = 128 words per CTA

// any bit set in positions [a, b] (inclusive, a <= b) of the bit string w?
__device__ __forceinline__ bool any_bits(const uint32_t* w, int a, int b) {
  const int wa = a >> 5, wb = b >> 5;
  const uint32_t ma = 0xffffffffu << (a & 31), mb = 0xffffffffu >> (31 - (b & 31));
  if (wa == wb) return (w[wa] & ma & mb) != 0u;
  uint32_t acc = (w[wa] & ma) | (w[wb] & mb);
  for (int k = wa + 1; k < wb; ++k) acc |= w[k];
  return acc != 0u;
}

// PASS 0: count the merged-region starts / ends of each CTA's tile.  PASS 1: rank them and write the pairs.
// The tile's hot bits plus a halo of `gap` bins each side sit in shared memory as words; a thread owns 16
// consecutive bins, so ranks inside the CTA come from ONE block scan of (starts | ends << 16) counts.
template <int PASS>
__global__ void __launch_bounds__(kScanThreads)
regions_kernel(const uint32_t* __restrict__ bits, int64_t n_words, int gap, int32_t* __restrict__ block_counts,
               int32_t* __restrict__ regions, int cap) {
  extern __shared__ uint32_t w[];                  // [kTile / 32 + 2 hw], word 0 = global word tile0 / 32 - hw
  __shared__ uint32_t warp_tot[kScanThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int hw = (gap + 31) >> 5;
  const int64_t tile0 = (int64_t)blockIdx.x * kTile;
  const int64_t wbase = (tile0 >> 5) - hw;
  for (int k = tid; k < kTile / 32 + 2 * hw; k += kScanThreads) {
    const int64_t g = wbase + k;
    w[k] = (g >= 0 && g < n_words) ? __ldg(bits + g) : 0u;
  }
  __syncthreads();
  const int li0 = hw * 32 + tid * kBinsPerThread;     // my first bin as a bit index into w
  const uint32_t mine = (w[li0 >> 5] >> (li0 & 31)) & 0xffffu;
  uint32_t smask = 0u, emask = 0u;
  if (gap >= kBinsPerThread - 1) {
    // every hot bin but my lowest has a hot bin less than 16 <= gap + 1 bins before it, every one but my highest has
    // one after it: two range tests per thread, whatever the data
    if (mine) {
      const int lo = __ffs(mine) - 1, hi = 31 - __clz(mine);
      if (!any_bits(w, li0 + lo - gap, li0 + lo - 1)) smask = 1u << lo;
      if (!any_bits(w, li0 + hi + 1, li0 + hi + gap)) emask = 1u << hi;
    }
  } else {
    for (uint32_t m = mine; m; m &= m - 1) {
      const int b = __ffs(m) - 1;
      const int li = li0 + b;
      if (!any_bits(w, li - gap, li - 1)) smask |= 1u << b;
      if (!any_bits(w, li + 1, li + gap)) emask |= 1u << b;
    }
  }
  const uint32_t c = (uint32_t)__popc(smask) | ((uint32_t)__popc(emask) << 16);
  uint32_t incl = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  uint32_t off = 0u, tot = 0u;
#pragma unroll
  for (int k = 0; k < kScanThreads / 32; ++k) {
    if (k < wid) off += warp_tot[k];
    tot += warp_tot[k];
  }
  if (PASS == 0) {
    if (tid == 0) {
      block_counts[2 * blockIdx.x] = (int32_t)(tot & 0xffffu);
      block_counts[2 * blockIdx.x + 1] = (int32_t)(tot >> 16);
    }
  } else {
    const uint32_t excl = off + incl - c;
    int ks = block_counts[2 * blockIdx.x] + (int)(excl & 0xffffu);
    int ke = block_counts[2 * blockIdx.x + 1] + (int)(excl >> 16);
    const int64_t j0 = tile0 + tid * kBinsPerThread;
    for (uint32_t m = smask; m; m &= m - 1, ++ks)
      if (ks < cap) regions[2 * ks] = (int32_t)(j0 + __ffs(m) - 1);
    for (uint32_t m = emask; m; m &= m - 1, ++ke)
      if (ke < cap) regions[2 * ke + 1] = (int32_t)(j0 + __ffs(m) - 1);
  }
}

// Exclusive scan of the per-CTA (starts, ends) pairs, in place; total number of regions -> *n_regions.
__global__ void __launch_bounds__(1024)
scan_counts_kernel(int32_t* __restrict__ block_counts, int n_blocks, int32_t* __restrict__ n_regions) {
  __shared__ int ws[32], we[32];
  __shared__ int carry_s, carry_e;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) { carry_s = 0; carry_e = 0; }
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int i = base + tid;
    int vs = 0, ve = 0;
    if (i < n_blocks) { vs = block_counts[2 * i]; ve = block_counts[2 * i + 1]; }
    int ps = vs, pe = ve;                       // inclusive warp scan
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, ps, d);
      const int b = __shfl_up_sync(0xffffffffu, pe, d);
      if (lane >= d) { ps += a; pe += b; }
    }
    if (lane == 31) { ws[wid] = ps; we[wid] = pe; }
    __syncthreads();
    int os = 0, oe = 0, ts = 0, te = 0;
    for (int w = 0; w < 32; ++w) {
      if (w < wid) { os += ws[w]; oe += we[w]; }
      ts += ws[w]; te += we[w];
    }
    if (i < n_blocks) {
      block_counts[2 * i] = carry_s + os + ps - vs;
      block_counts[2 * i + 1] = carry_e + oe + pe - ve;
    }
    __syncthreads();
    if (tid == 0) { carry_s += ts; carry_e += te; }
    __syncthreads();
  }
  if (tid == 0) *n_regions = carry_s;   // == carry_e: every merged region has one start and one end
}

}  // namespace

int launch_average(const float* logits, int n_windows, int64_t out_len, double* avg, int32_t* cnt, cudaStream_t st) {
  if (out_len <= 0) return SS_OK;
  launch_average_kernel<false>(logits, n_windows, out_len, avg, cnt, 0.0, nullptr, st);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

// K6 scratch: [2 n_blocks + 2 per-CTA counts][(out_len + 31) / 32 + 2 hot-bit words]
static int64_t count_words(int64_t out_len) { return 2 * ((out_len + kTile - 1) / kTile) + 2; }
int64_t regions_scan_tmp_len(int64_t out_len) { return count_words(out_len) + (out_len + 31) / 32 + 2; }

static int check_regions_args(int64_t out_len, int& gap_bins, int64_t scan_tmp_len) {
  SS_REQUIRE(gap_bins >= 0 && gap_bins <= 4096, SS_E_ARG, "gap_bins %d out of range [0, 4096]", gap_bins);
  SS_REQUIRE(out_len < ((int64_t)1 << 31), SS_E_ARG, "timeline of %lld bins exceeds int32 bin indices",
             (long long)out_len);
  SS_REQUIRE(scan_tmp_len >= regions_scan_tmp_len(out_len), SS_E_CAPACITY, "scan scratch too small");
  // Two distinct runs are at least 2 bins apart, so gap 0 and 1 both mean "never merge"; a look-back of at
  // least one bin is still needed to find where a run starts and ends.
  if (gap_bins < 1) gap_bins = 1;
  return SS_OK;
}

// count pass, scan of the per-CTA counts, emit pass — over the hot-bit words only (1 bit per bin)
static int launch_regions_bits(const uint32_t* bits, int64_t out_len, int gap_bins, int32_t* regions, int32_t* n_regions,
                               int cap, int32_t* counts, cudaStream_t st) {
  const int n_blocks = (int)((out_len + kTile - 1) / kTile);
  const int64_t n_words = (out_len + 31) / 32;
  const size_t smem = (size_t)(kTile / 32 + 2 * ((gap_bins + 31) / 32)) * sizeof(uint32_t);
  regions_kernel<0><<<n_blocks, kScanThreads, smem, st>>>(bits, n_words, gap_bins, counts, regions, cap);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  scan_counts_kernel<<<1, 1024, 0, st>>>(counts, n_blocks, n_regions);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  regions_kernel<1><<<n_blocks, kScanThreads, smem, st>>>(bits, n_words, gap_bins, counts, regions, cap);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int launch_regions(const double* avg, const int32_t* cnt, int64_t out_len, double threshold, int gap_bins,
                   int32_t* regions, int32_t* n_regions, int cap, int32_t* scan_tmp, int64_t scan_tmp_len,
                   cudaStream_t st) {
  int rc = check_regions_args(out_len, gap_bins, scan_tmp_len);
  if (rc) return rc;
  if (out_len <= 0) {
    SS_CUDA_CHECK(cudaMemsetAsync(n_regions, 0, sizeof(int32_t), st));
    return SS_OK;
  }
  uint32_t* bits = reinterpret_cast<uint32_t*>(scan_tmp + count_words(out_len));
  hot_bits_kernel<<<(int)((out_len + kAvgThreads * kHotBins - 1) / (kAvgThreads * kHotBins)), kAvgThreads, 0, st>>>(
      avg, cnt, out_len, threshold, bits);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return launch_regions_bits(bits, out_len, gap_bins, regions, n_regions, cap, scan_tmp, st);
}

// K5 of the detection pipeline: the averaging kernel votes the hot bits (into the K6 scratch) while the averages are
// in registers and, when `win_flags` is given, marks the windows covering a bin within `eps` of the threshold.
int launch_average_bits(const float* logits, int n_windows, int64_t out_len, double* avg, int32_t* cnt,
                        double threshold, int32_t* scan_tmp, int64_t scan_tmp_len, double eps,
                        unsigned char* win_flags, cudaStream_t st) {
  int gap = 1;
  int rc = check_regions_args(out_len, gap, scan_tmp_len);
  if (rc) return rc;
  if (out_len <= 0) return SS_OK;
  uint32_t* bits = reinterpret_cast<uint32_t*>(scan_tmp + count_words(out_len));
  launch_average_kernel<true>(logits, n_windows, out_len, avg, cnt, threshold, bits, st, eps, win_flags);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

// K6 on the hot bits launch_average_bits left in the scratch.
int launch_regions_after_bits(int64_t out_len, int gap_bins, int32_t* regions, int32_t* n_regions, int cap,
                              int32_t* scan_tmp, int64_t scan_tmp_len, cudaStream_t st) {
  int rc = check_regions_args(out_len, gap_bins, scan_tmp_len);
  if (rc) return rc;
  if (out_len <= 0) {
    SS_CUDA_CHECK(cudaMemsetAsync(n_regions, 0, sizeof(int32_t), st));
    return SS_OK;
  }
  uint32_t* bits = reinterpret_cast<uint32_t*>(scan_tmp + count_words(out_len));
  return launch_regions_bits(bits, out_len, gap_bins, regions, n_regions, cap, scan_tmp, st);
}

int launch_average_regions(const float* logits, int n_windows, int64_t out_len, double* avg, int32_t* cnt,
                           double threshold, int gap_bins, int32_t* regions, int32_t* n_regions, int cap,
                           int32_t* scan_tmp, int64_t scan_tmp_len, cudaStream_t st) {
  int rc = launch_average_bits(logits, n_windows, out_len, avg, cnt, threshold, scan_tmp, scan_tmp_len, 0.0, nullptr, st);
  if (rc) return rc;
  return launch_regions_after_bits(out_len, gap_bins, regions, n_regions, cap, scan_tmp, scan_tmp_len, st);
}

int launch_compact_flags(const unsigned char* flags, int n, int32_t* list, int32_t* count, cudaStream_t st) {
  compact_flags_kernel<<<1, 1024, 0, st>>>(flags, n, list, count);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int launch_scatter_rows(const float* rows, const int32_t* list, int n, float* logits, cudaStream_t st) {
  if (n <= 0) return SS_OK;
  scatter_rows_kernel<<<n, kFrames, 0, st>>>(rows, list, logits);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

}  // namespace ss
