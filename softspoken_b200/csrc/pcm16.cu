// PCM_16 decode / encode on the device: the sample-format conversions libsndfile does
// on the host for the reference, as bandwidth-bound element-wise kernels, so that a PCM_16 file crosses PCIe as
// 2-byte samples in both directions.
//
//   decode  `sf.read(path, dtype='float32')` + `librosa.to_mono` (root/code/backend/voice_activity.py:37,61-62):
//           float = int16 / 32768 (libsndfile's normalised float read), mono = mean over channels — numpy's float32
//           `add.reduce` over the channel axis followed by a float32 division by the channel count.  Every partial
//           sum of PCM_16 samples is a multiple of 2^-15 below 2^9, hence exact in float32 for up to 256 channels:
//           the result does not depend on the order of the additions, and the single rounding is the division.
//   encode  `sf.write(path, audio.T, sr)` (root/code/frontend/silencer_ui.py:998; WAV default subtype PCM_16):
//           libsndfile's float -> short conversion with normalisation on and clipping off is
//           short(lrintf(x * 32767.0f)) (pcm.c:f2les_array).  libsndfile is not in the build image, so this line is
//           restated from its published source and NOT pinned by a golden vector; samples outside
//           [-1, 1] saturate here where the C conversion would wrap.
//   requant the composition encode(decode(k)) on an int16 buffer: what the reference's read -> write round trip does
//           to the samples it does not zero (|k| >= 16384 lose one LSB because 32767 != 32768).
#include "ss_common.cuh"

namespace ss {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int16_t encode1(float x) {
  const int v = __float2int_rn(x * 32767.0f);          // round half to even, as lrintf in the default rounding mode
  return (int16_t)max(-32768, min(32767, v));
}

__global__ void __launch_bounds__(kThreads)
decode_pcm16_kernel(const int16_t* __restrict__ in, int64_t frames, int channels, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const float inv_c = (float)channels;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < frames; i += stride) {
    const int16_t* f = in + i * channels;
    float acc = (float)__ldg(f) * (1.0f / 32768.0f);
    for (int c = 1; c < channels; ++c) acc += (float)__ldg(f + c) * (1.0f / 32768.0f);
    out[i] = channels == 1 ? acc : __fdiv_rn(acc, inv_c);
  }
}

// mono fast path: 8 samples (16 bytes) per thread and iteration
__global__ void __launch_bounds__(kThreads)
decode_pcm16_mono_kernel(const int16_t* __restrict__ in, int64_t n, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n8 = n >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float f[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[2 * k] = (float)(int16_t)(w[k] & 0xFFFFu) * (1.0f / 32768.0f);
      f[2 * k + 1] = (float)(int16_t)(w[k] >> 16) * (1.0f / 32768.0f);
    }
    reinterpret_cast<float4*>(out)[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4*>(out)[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  const int64_t i = (n8 << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)__ldg(in + i) * (1.0f / 32768.0f);
}

__global__ void __launch_bounds__(kThreads)
encode_pcm16_kernel(const float* __restrict__ in, int64_t n, int16_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n8 = n >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
    const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      w[k] = (uint32_t)(uint16_t)encode1(f[2 * k]) | ((uint32_t)(uint16_t)encode1(f[2 * k + 1]) << 16);
    reinterpret_cast<uint4*>(out)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  const int64_t i = (n8 << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = encode1(__ldg(in + i));
}

__global__ void __launch_bounds__(kThreads)
requant_pcm16_kernel(int16_t* __restrict__ pcm, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n8 = n >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    uint4 v = reinterpret_cast<uint4*>(pcm)[i];
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int16_t lo = encode1((float)(int16_t)(w[k] & 0xFFFFu) * (1.0f / 32768.0f));
      const int16_t hi = encode1((float)(int16_t)(w[k] >> 16) * (1.0f / 32768.0f));
      w[k] = (uint32_t)(uint16_t)lo | ((uint32_t)(uint16_t)hi << 16);
    }
    reinterpret_cast<uint4*>(pcm)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  const int64_t i = (n8 << 3) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pcm[i] = encode1((float)pcm[i] * (1.0f / 32768.0f));
}

int grid_for(int64_t items) {
  int64_t blocks = (items + kThreads - 1) / kThreads;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  return blocks < 1 ? 1 : (int)blocks;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

int launch_decode_pcm16(const int16_t* interleaved, int64_t frames, int channels, float* mono, cudaStream_t st) {
  if (frames <= 0) return SS_OK;
  if (channels == 1 && aligned16(interleaved) && aligned16(mono))
    decode_pcm16_mono_kernel<<<grid_for((frames >> 3) + 8), kThreads, 0, st>>>(interleaved, frames, mono);
  else
    decode_pcm16_kernel<<<grid_for(frames), kThreads, 0, st>>>(interleaved, frames, channels, mono);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int launch_encode_pcm16(const float* src, int64_t n_elems, int16_t* dst, cudaStream_t st) {
  if (n_elems <= 0) return SS_OK;
  SS_REQUIRE(aligned16(src) && aligned16(dst), SS_E_ARG, "ss_encode_pcm16 needs 16-byte aligned buffers");
  encode_pcm16_kernel<<<grid_for((n_elems >> 3) + 8), kThreads, 0, st>>>(src, n_elems, dst);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int launch_requant_pcm16(int16_t* pcm, int64_t n_elems, cudaStream_t st) {
  if (n_elems <= 0) return SS_OK;
  SS_REQUIRE(aligned16(pcm), SS_E_ARG, "PCM_16 buffer must be 16-byte aligned");
  requant_pcm16_kernel<<<grid_for((n_elems >> 3) + 8), kThreads, 0, st>>>(pcm, n_elems);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

}  // namespace ss
