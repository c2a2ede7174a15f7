// K9 — sample-rate conversion to 22,050 Hz as a polyphase FIR.
//
// Stands where the reference calls librosa.resample (soxr "HQ") in voice_activity.load_audio
// (root/code/backend/voice_activity.py:44-66).  Not a restatement of soxr (absent from the build image, unpinned in
// the reference): the filter is the documented Kaiser-windowed sinc of softspoken_b200/resample.py, whose table the
// host designs once per input rate; parity with the reference is unpinned for resampled files.
//
//   output m  <->  input time m M / L = n0 + p / L      (n0 = m M div L, p = m M mod L)
//   y[m] = sum_{j = -T .. T} x[n0 - j] * table[j + T][m mod L],  x = 0 outside [0, n_in); column q = m mod L of the
//   table holds phase p = (q M) mod L (visit order: the outputs of a warp read neighbouring columns)
//
// One thread per output sample, 256 consecutive outputs per CTA.  The input span of a CTA (256 M / L + 2 T + 2 samples)
// is staged once in shared memory with the zero padding applied, so the inner loop is one shared-memory load, one
// table load (a [2T+1][L] float table of a few hundred KB at most: L1 / L2 resident, read through the read-only
// path) and one FMA per tap.  HBM traffic is the algorithmic minimum, 4 n_in + 4 n_out bytes (2 n_in for int16
// input); at 140-300 taps per output the kernel is bound by FP32 / LSU issue, not by HBM.
#include "ss_common.cuh"

namespace ss {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(int16_t v) { return (float)v * (1.0f / 32768.0f); }

template <typename T>
__global__ void __launch_bounds__(kThreads)
resample_kernel(const T* __restrict__ x, int64_t n_in, float* __restrict__ y, int64_t n_out, int L, int M, int taps_half,
                const float* __restrict__ table, int span) {
  extern __shared__ float xs[];                      // xs[i] = x[base + i], zero outside the clip
  const int64_t m0 = (int64_t)blockIdx.x * kThreads;
  const int64_t base = (m0 * M) / L - taps_half;     // first input sample any output of this CTA reads is base (j = +T)
  for (int i = threadIdx.x; i < span; i += kThreads) {
    const int64_t k = base + i;
    xs[i] = (k >= 0 && k < n_in) ? to_float(x[k]) : 0.f;
  }
  __syncthreads();
  const int64_t m = m0 + threadIdx.x;
  if (m >= n_out) return;
  const int64_t pos = m * M;
  const int64_t n0 = pos / L;
  const int q = (int)(m % L);      // table column: phase (m M) mod L in visit order, so a warp reads neighbouring columns
  // x[n0 - j] for j = -T .. T  =  xs[(n0 - base) - j]: walk the taps with j ascending, the samples descending
  const float* xp = xs + (int)(n0 - base) + taps_half;          // j = -T
  const float* tp = table + q;
  float acc0 = 0.f, acc1 = 0.f;                                  // two chains: the loop is latency-bound otherwise
  const int n_taps = 2 * taps_half + 1;
  int t = 0;
  for (; t + 1 < n_taps; t += 2) {
    acc0 = fmaf(xp[-t], __ldg(tp + (int64_t)t * L), acc0);
    acc1 = fmaf(xp[-t - 1], __ldg(tp + (int64_t)(t + 1) * L), acc1);
  }
  if (t < n_taps) acc0 = fmaf(xp[-t], __ldg(tp + (int64_t)t * L), acc0);
  y[m] = acc0 + acc1;
}

}  // namespace

int launch_resample(const void* x, int sample_fmt, int64_t n_in, float* y, int64_t n_out, int L, int M, int taps_half,
                    const float* table, cudaStream_t st) {
  if (n_out <= 0) return SS_OK;
  SS_REQUIRE(L >= 1 && M >= 1 && taps_half >= 0 && taps_half <= 4096, SS_E_ARG, "bad resampling ratio %d / %d, T = %d", L,
             M, taps_half);
  // input samples one CTA touches: from n0(first) - T to n0(last) + T
  const int64_t span64 = ((int64_t)(kThreads - 1) * M) / L + 2 * (int64_t)taps_half + 3;
  SS_REQUIRE(span64 * (int64_t)sizeof(float) <= 200 * 1024, SS_E_ARG,
             "resampling %d / %d with %d taps needs %lld staged samples per CTA: unsupported rate", L, M,
             2 * taps_half + 1, (long long)span64);
  const int span = (int)span64;
  const size_t smem = (size_t)span * sizeof(float);
  const int grid = (int)((n_out + kThreads - 1) / kThreads);
  if (sample_fmt == kSampleS16) {
    if (smem > 48 * 1024)
      SS_CUDA_CHECK(cudaFuncSetAttribute(resample_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resample_kernel<int16_t><<<grid, kThreads, smem, st>>>(static_cast<const int16_t*>(x), n_in, y, n_out, L, M, taps_half,
                                                           table, span);
  } else {
    if (smem > 48 * 1024)
      SS_CUDA_CHECK(cudaFuncSetAttribute(resample_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    resample_kernel<float><<<grid, kThreads, smem, st>>>(static_cast<const float*>(x), n_in, y, n_out, L, M, taps_half,
                                                         table, span);
  }
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

}  // namespace ss
