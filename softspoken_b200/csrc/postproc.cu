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

__global__ void average_kernel(const float* __restrict__ logits, int n_windows, int64_t out_len,
                               double* __restrict__ avg, int32_t* __restrict__ cnt) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= out_len) return;
  int64_t i_hi = (5 * j + 2) / 256;                 // largest i with p_i <= j
  if (i_hi > n_windows - 1) i_hi = n_windows - 1;
  int64_t idx[8];
  int n = 0;
  for (int64_t i = i_hi; i >= 0 && n < 8; --i) {
    const int64_t p = (256 * i + 2) / 5;
    if (p + 255 < j) break;
    idx[n++] = i;
  }
  double sum = 0.0;
  for (int k = n - 1; k >= 0; --k) {                // ascending window index
    const int64_t i = idx[k];
    const int64_t p = (256 * i + 2) / 5;
    sum += (double)__ldg(logits + i * 256 + (j - p));
  }
  cnt[j] = n;
  avg[j] = n ? sum / (double)n : __longlong_as_double(0x7ff8000000000000LL);
}

constexpr int kScanThreads = 256;
constexpr int kRounds = 4;
constexpr int kTile = kScanThreads * kRounds;   // bins per CTA

// PASS 0: count starts/ends per CTA.  PASS 1: rank and emit.
template <int PASS>
__global__ void __launch_bounds__(kScanThreads)
regions_kernel(const double* __restrict__ avg, const int32_t* __restrict__ cnt, int64_t out_len, double threshold,
               int gap, int32_t* __restrict__ block_counts, int32_t* __restrict__ regions, int cap) {
  extern __shared__ unsigned char hot[];           // [kTile + 2 gap], bin tile0 - gap at index 0
  __shared__ int warp_s[kScanThreads / 32], warp_e[kScanThreads / 32];
  __shared__ int base_s, base_e;
  const int tid = threadIdx.x;
  const int64_t tile0 = (int64_t)blockIdx.x * kTile;
  for (int i = tid; i < kTile + 2 * gap; i += kScanThreads) {
    const int64_t j = tile0 - gap + i;
    unsigned char h = 0;
    if (j >= 0 && j < out_len) h = (cnt[j] >= 1 && avg[j] > threshold) ? 1 : 0;
    hot[i] = h;
  }
  if (tid == 0) {
    if (PASS == 1) { base_s = block_counts[2 * blockIdx.x]; base_e = block_counts[2 * blockIdx.x + 1]; }
    else { base_s = 0; base_e = 0; }
  }
  __syncthreads();

  const int lane = tid & 31, wid = tid >> 5;
  for (int r = 0; r < kRounds; ++r) {
    const int li = r * kScanThreads + tid + gap;     // index into hot[]
    const int64_t j = tile0 + r * kScanThreads + tid;
    bool is_s = false, is_e = false;
    if (hot[li]) {
      is_s = true; is_e = true;
      for (int d = 1; d <= gap; ++d) {
        if (hot[li - d]) is_s = false;
        if (hot[li + d]) is_e = false;
      }
    }
    const unsigned bs = __ballot_sync(0xffffffffu, is_s);
    const unsigned be = __ballot_sync(0xffffffffu, is_e);
    if (lane == 0) { warp_s[wid] = __popc(bs); warp_e[wid] = __popc(be); }
    __syncthreads();
    int off_s = 0, off_e = 0, tot_s = 0, tot_e = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
      if (w < wid) { off_s += warp_s[w]; off_e += warp_e[w]; }
      tot_s += warp_s[w]; tot_e += warp_e[w];
    }
    if (PASS == 1) {
      const unsigned below = (1u << lane) - 1u;
      if (is_s) {
        const int k = base_s + off_s + __popc(bs & below);
        if (k < cap) regions[2 * k] = (int32_t)j;
      }
      if (is_e) {
        const int k = base_e + off_e + __popc(be & below);
        if (k < cap) regions[2 * k + 1] = (int32_t)j;
      }
    }
    __syncthreads();
    if (tid == 0) { base_s += tot_s; base_e += tot_e; }
    __syncthreads();
  }
  if (PASS == 0 && tid == 0) {
    block_counts[2 * blockIdx.x] = base_s;
    block_counts[2 * blockIdx.x + 1] = base_e;
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
  const int threads = 256;
  average_kernel<<<(int)((out_len + threads - 1) / threads), threads, 0, st>>>(logits, n_windows, out_len, avg, cnt);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int64_t regions_scan_tmp_len(int64_t out_len) { return 2 * ((out_len + kTile - 1) / kTile) + 2; }

int launch_regions(const double* avg, const int32_t* cnt, int64_t out_len, double threshold, int gap_bins,
                   int32_t* regions, int32_t* n_regions, int cap, int32_t* scan_tmp, int64_t scan_tmp_len,
                   cudaStream_t st) {
  SS_REQUIRE(gap_bins >= 0 && gap_bins <= 4096, SS_E_ARG, "gap_bins %d out of range [0, 4096]", gap_bins);
  SS_REQUIRE(out_len < ((int64_t)1 << 31), SS_E_ARG, "timeline of %lld bins exceeds int32 bin indices",
             (long long)out_len);
  if (out_len <= 0) {
    SS_CUDA_CHECK(cudaMemsetAsync(n_regions, 0, sizeof(int32_t), st));
    return SS_OK;
  }
  const int n_blocks = (int)((out_len + kTile - 1) / kTile);
  SS_REQUIRE(scan_tmp_len >= 2 * (int64_t)n_blocks, SS_E_CAPACITY, "scan scratch too small");
  // Two distinct runs are at least 2 bins apart, so gap 0 and 1 both mean "never merge"; a look-back of at
  // least one bin is still needed to find where a run starts and ends.
  if (gap_bins < 1) gap_bins = 1;
  const size_t smem = kTile + 2 * gap_bins;
  regions_kernel<0><<<n_blocks, kScanThreads, smem, st>>>(avg, cnt, out_len, threshold, gap_bins, scan_tmp, regions, cap);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  scan_counts_kernel<<<1, 1024, 0, st>>>(scan_tmp, n_blocks, n_regions);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  regions_kernel<1><<<n_blocks, kScanThreads, smem, st>>>(avg, cnt, out_len, threshold, gap_bins, scan_tmp, regions, cap);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

}  // namespace ss
