// K7 — interval masking.
//
// Replaces the inner loop of SilenceWorker.run, `audio_data[:, start_index:end_index] = 0.0`
// (root/code/frontend/silencer_ui.py:974-985), for a whole table of intervals in one launch.  The host
// computes the sample indices exactly as the reference does (Python round-half-even of the double product,
// clamp to [0, n]; softspoken_b200/silencer.py) and expands channels, so the kernel sees half-open element
// ranges of one packed buffer of float32 samples — or of int16 samples when a PCM_16 file is silenced without
// ever leaving its storage format (pcm16.cu).  Write-only, HBM-bound: 16-byte stores on the aligned body,
// scalar stores on the ragged head and tail.  Overlapping intervals are harmless (idempotent zero stores).
#include <stdlib.h>

#include "ss_common.cuh"

namespace ss {

namespace {

constexpr int kThreads = 256;
constexpr int kVecsPerThread = 8;
constexpr int kVecsPerIter = kThreads * kVecsPerThread;   // 16-byte vectors per CTA iteration: 32 KB
// CTAs per interval.  A detection is 0.2-4 s (18-350 KB of float32 samples), i.e. 1-11 iterations: with 32 CTAs per
// interval two thirds of the grid found nothing to do and the launch was bound by CTA turnover, not by stores.
constexpr int kChunksY = 8;

template <typename T>
__global__ void __launch_bounds__(kThreads)
silence_kernel(T* __restrict__ pcm, int64_t n_elems, int64_t shift, const ss_interval* __restrict__ iv,
               int n_intervals) {
  constexpr int kPerVec = 16 / (int)sizeof(T);
  const ss_interval r = iv[blockIdx.x];
  int64_t b = r.begin - shift, e = r.end - shift;
  if (b < 0) b = 0;
  if (e > n_elems) e = n_elems;
  if (e <= b) return;
  // aligned body [b4, e4), scalar head [b, b4) and tail [e4, e)
  const uintptr_t addr = reinterpret_cast<uintptr_t>(pcm + b);
  int64_t b4 = b + (int64_t)(((16 - (addr & 15)) & 15) / sizeof(T));
  if (b4 > e) b4 = e;
  const int64_t e4 = b4 + ((e - b4) / kPerVec) * kPerVec;
  if (blockIdx.y == 0) {
    const int64_t head = b4 - b, tail = e - e4;
    if (threadIdx.x < head) pcm[b + threadIdx.x] = T(0);
    if (threadIdx.x >= 32 && threadIdx.x - 32 < tail) pcm[e4 + threadIdx.x - 32] = T(0);
  }
  float4* body = reinterpret_cast<float4*>(pcm + b4);
  const int64_t n4 = (e4 - b4) / kPerVec;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);      // all-zero bits: 0.0f x 4 or int16 0 x 8
  for (int64_t c = (int64_t)blockIdx.y * kVecsPerIter; c < n4; c += (int64_t)gridDim.y * kVecsPerIter) {
#pragma unroll
    for (int k = 0; k < kVecsPerThread; ++k) {
      const int64_t i = c + k * kThreads + threadIdx.x;
      if (i < n4) body[i] = z;
    }
  }
}

int chunks_y() {      // SS_SILENCE_Y: tuning override of the CTAs per interval
  static const int y = [] { const char* e = getenv("SS_SILENCE_Y"); const int v = e ? atoi(e) : kChunksY; return v >= 1 && v <= 64 ? v : kChunksY; }();
  return y;
}

}  // namespace

int launch_silence(float* pcm, int64_t n_elems, int64_t shift, const ss_interval* iv, int n_intervals,
                   cudaStream_t st) {
  if (n_intervals <= 0) return SS_OK;
  dim3 grid(n_intervals, chunks_y());
  silence_kernel<float><<<grid, kThreads, 0, st>>>(pcm, n_elems, shift, iv, n_intervals);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int launch_silence_s16(int16_t* pcm, int64_t n_elems, int64_t shift, const ss_interval* iv, int n_intervals,
                       cudaStream_t st) {
  if (n_intervals <= 0) return SS_OK;
  dim3 grid(n_intervals, chunks_y());
  silence_kernel<int16_t><<<grid, kThreads, 0, st>>>(pcm, n_elems, shift, iv, n_intervals);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

}  // namespace ss
