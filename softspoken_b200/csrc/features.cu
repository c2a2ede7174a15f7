// K1 — fused log-mel front end.
//
// Replaces, for every 3 s window, the reference's window gather + torch.stack
// (root/code/frontend/NNDetector.py:90-96), torchaudio MelSpectrogram (reflect pad, 2048-point STFT of
// 512-tap Hann frames at hop 256, |.|^2, 128-band HTK mel; root/code/backend/pytorch_neural_nets.py:92-99,144),
// sqrt(log10(x + 1)) (:147) and the trim to 256 frames (:150).
//
// Data flow: PCM is read in place from the (optionally virtual) padded clip — no [W, 66150] window
// matrix is ever materialised.  One CTA owns 32 consecutive frames of one window.  Per frame the 2048-point
// spectrum of the 512 real taps is obtained from two 512-point complex FFTs (64 threads x 8 registers,
// three radix-8 passes, two shared-memory exchanges):
//   FFT_A of a[n] = x[n] w[n] e^{-2 pi i n / 2048}         -> bins 4q+1 (A[q]) and 4q+3 (conj A[511-q])
//   FFT_B of c[n] = x[2n] w[2n] + i x[2n+1] w[2n+1], n<256 -> even bins 2m by the real-FFT split
// Only bins 1..743 are formed (the filterbank is zero elsewhere).  The mel reduction walks each band's
// contiguous taps with 4 lanes per band and a warp-shuffle sum; the 128 x 32 output tile is staged in
// shared memory and stored as full 128-byte rows of the [W][128][256] feature tensor.
#include "ss_common.cuh"

namespace ss {

namespace {

constexpr int kThreads = 128;
constexpr int kFramesPerCta = 32;
constexpr int kExA = 8 * 72;   // exchange A: [k1][t] with row stride 72 (bank-conflict-free both ways)
constexpr int kExB = 64 * 9;   // exchange B: [k2a*8+k1][u] with row stride 9
constexpr int kEvenBins = 372; // even bins 2m, m <= 371  (k <= 742)
constexpr int kOddBins = 372;  // odd bins 2j+1, j <= 371 (k <= 743)
constexpr int kMaxTapsSmem = 2048;

struct Smem {
  float exA[2][2][kExA];
  float exB[2][2][kExB];
  float evenP[kEvenBins];
  float oddP[kOddBins];
  float2 tw[512];
  float taps[kMaxTapsSmem];
  int mstart[kMels], mcount[kMels], moffs[kMels];
  float tile[kMels][kFramesPerCta + 1];
};

// In-place forward 8-point DFT: y[k] = sum_j v[j] exp(-2 pi i j k / 8).
__device__ __forceinline__ void dft8(float (&re)[8], float (&im)[8]) {
  const float c = 0.70710678118654752440f;
  // even half: v0 v2 v4 v6
  float s0r = re[0] + re[4], s0i = im[0] + im[4];
  float s1r = re[0] - re[4], s1i = im[0] - im[4];
  float s2r = re[2] + re[6], s2i = im[2] + im[6];
  float s3r = re[2] - re[6], s3i = im[2] - im[6];
  float e0r = s0r + s2r, e0i = s0i + s2i;
  float e2r = s0r - s2r, e2i = s0i - s2i;
  float e1r = s1r + s3i, e1i = s1i - s3r;
  float e3r = s1r - s3i, e3i = s1i + s3r;
  // odd half: v1 v3 v5 v7
  float t0r = re[1] + re[5], t0i = im[1] + im[5];
  float t1r = re[1] - re[5], t1i = im[1] - im[5];
  float t2r = re[3] + re[7], t2i = im[3] + im[7];
  float t3r = re[3] - re[7], t3i = im[3] - im[7];
  float o0r = t0r + t2r, o0i = t0i + t2i;
  float o2r = t0r - t2r, o2i = t0i - t2i;
  float o1r = t1r + t3i, o1i = t1i - t3r;
  float o3r = t1r - t3i, o3i = t1i + t3r;
  // twiddles W8^k
  float p1r = (o1r + o1i) * c, p1i = (o1i - o1r) * c;   // (1 - i)/sqrt2
  float p2r = o2i, p2i = -o2r;                          // -i
  float p3r = (o3i - o3r) * c, p3i = -(o3r + o3i) * c;  // (-1 - i)/sqrt2
  re[0] = e0r + o0r; im[0] = e0i + o0i;
  re[4] = e0r - o0r; im[4] = e0i - o0i;
  re[1] = e1r + p1r; im[1] = e1i + p1i;
  re[5] = e1r - p1r; im[5] = e1i - p1i;
  re[2] = e2r + p2r; im[2] = e2i + p2i;
  re[6] = e2r - p2r; im[6] = e2i - p2i;
  re[3] = e3r + p3r; im[3] = e3i + p3i;
  re[7] = e3r - p3r; im[7] = e3i - p3i;
}

__device__ __forceinline__ void cmul(float& r, float& i, float2 w) {
  float nr = r * w.x - i * w.y;
  float ni = r * w.y + i * w.x;
  r = nr;
  i = ni;
}

// Sample `l` (relative to the window start, may be negative for frame 0 -> torch 'reflect') of the
// virtual padded clip: indices inside [valid_begin, valid_end) map to pcm[idx - offset], the rest are 0.
__device__ __forceinline__ float load_sample(const float* __restrict__ pcm, int64_t wstart, int l,
                                             int64_t valid_begin, int64_t valid_end, int64_t offset) {
  int64_t idx = wstart + (l < 0 ? -l : l);
  return (idx >= valid_begin && idx < valid_end) ? __ldg(pcm + (idx - offset)) : 0.0f;
}

__global__ void __launch_bounds__(kThreads)
features_kernel(const float* __restrict__ pcm, int64_t valid_begin, int64_t valid_end, int64_t offset,
                const int64_t* __restrict__ starts, int64_t w_base, FrontEnd fe, float* __restrict__ mel) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);

  const int tid = threadIdx.x;
  const int half = tid >> 6;      // 0: FFT_A, 1: FFT_B
  const int t64 = tid & 63;
  const int w = blockIdx.x;
  const int frame0 = blockIdx.y * kFramesPerCta;
  const int64_t wstart = starts ? starts[w] : (w_base + w) * (int64_t)kStepSamples;

  for (int i = tid; i < 512; i += kThreads) s.tw[i] = fe.tw512[i];
  for (int i = tid; i < fe.n_taps; i += kThreads) s.taps[i] = fe.mel_taps[i];
  if (tid < kMels) {
    s.mstart[tid] = fe.mel_start[tid];
    s.mcount[tid] = fe.mel_count[tid];
    s.moffs[tid] = fe.mel_offs[tid];
  }
  __syncthreads();

  float* exAr = s.exA[half][0];
  float* exAi = s.exA[half][1];
  float* exBr = s.exB[half][0];
  float* exBi = s.exB[half][1];

  for (int f = 0; f < kFramesPerCta; ++f) {
    const int l0 = (frame0 + f) * kHop - kHop;   // first sample of the frame relative to the window
    float re[8], im[8];

    // ---- stage 1: load, window, (pre-twiddle,) radix-8 over j, twiddle W512^(t k1)
    if (half == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = t64 + 64 * j;
        const float x = load_sample(pcm, wstart, l0 + n, valid_begin, valid_end, offset);
        re[j] = x * __ldg(fe.tw_a_re + n);
        im[j] = x * __ldg(fe.tw_a_im + n);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = t64 + 64 * j;
        const float x0 = load_sample(pcm, wstart, l0 + 2 * n, valid_begin, valid_end, offset);
        const float x1 = load_sample(pcm, wstart, l0 + 2 * n + 1, valid_begin, valid_end, offset);
        re[j] = x0 * __ldg(fe.window + 2 * n);
        im[j] = x1 * __ldg(fe.window + 2 * n + 1);
      }
#pragma unroll
      for (int j = 4; j < 8; ++j) { re[j] = 0.f; im[j] = 0.f; }
    }
    dft8(re, im);
#pragma unroll
    for (int k1 = 0; k1 < 8; ++k1) {
      if (k1) cmul(re[k1], im[k1], s.tw[t64 * k1]);
      exAr[k1 * 72 + t64] = re[k1];
      exAi[k1 * 72 + t64] = im[k1];
    }
    __syncthreads();

    // ---- stage 2: thread (k1, u) transforms over v, twiddle W64^(u k2a)
    {
      const int k1 = t64 >> 3, u = t64 & 7;
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        re[v] = exAr[k1 * 72 + u + 8 * v];
        im[v] = exAi[k1 * 72 + u + 8 * v];
      }
      dft8(re, im);
#pragma unroll
      for (int k2a = 0; k2a < 8; ++k2a) {
        if (k2a) cmul(re[k2a], im[k2a], s.tw[8 * u * k2a]);
        exBr[(k2a * 8 + k1) * 9 + u] = re[k2a];
        exBi[(k2a * 8 + k1) * 9 + u] = im[k2a];
      }
    }
    __syncthreads();

    // ---- stage 3: thread q0 = k1 + 8 k2a transforms over u -> X[q0 + 64 k2b]
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      re[u] = exBr[t64 * 9 + u];
      im[u] = exBi[t64 * 9 + u];
    }
    dft8(re, im);
    if (half == 0) {
#pragma unroll
      for (int k2b = 0; k2b < 8; ++k2b) {
        const int q = t64 + 64 * k2b;
        const float pw = re[k2b] * re[k2b] + im[k2b] * im[k2b];
        if (q <= 185) s.oddP[2 * q] = pw;                    // bin 4q+1
        if (q >= 326) s.oddP[2 * (511 - q) + 1] = pw;        // bin 4(511-q)+3 = conj symmetry
      }
    } else {
      // C[q] parked in FFT_B's (now free) exchange-A buffers, plain layout
#pragma unroll
      for (int k2b = 0; k2b < 8; ++k2b) {
        const int q = t64 + 64 * k2b;
        s.exA[1][0][q] = re[k2b];
        s.exA[1][1][q] = im[k2b];
      }
    }
    __syncthreads();

    // ---- even bins: U[m] = (C[m] + conj C[512-m])/2 + W1024^m (C[m] - conj C[512-m])/(2i)
    for (int m = tid; m < kEvenBins; m += kThreads) {
      const int mm = (512 - m) & 511;
      const float cr = s.exA[1][0][m], ci = s.exA[1][1][m];
      const float nr = s.exA[1][0][mm], ni = -s.exA[1][1][mm];
      const float er = 0.5f * (cr + nr), ei = 0.5f * (ci + ni);
      const float dr = cr - nr, di = ci - ni;
      float orr = 0.5f * di, oi = -0.5f * dr;
      cmul(orr, oi, __ldg(fe.tw1024 + m));
      const float ur = er + orr, ui = ei + oi;
      s.evenP[m] = ur * ur + ui * ui;
    }
    __syncthreads();

    // ---- mel: 4 lanes per band, 32 bands per round, warp-shuffle reduction
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int band = r * 32 + (tid >> 2);
      const int lane4 = tid & 3;
      const int k0 = s.mstart[band], cnt = s.mcount[band], off = s.moffs[band];
      float acc = 0.f;
      for (int i = lane4; i < cnt; i += 4) {
        const int k = k0 + i;
        const float p = (k & 1) ? s.oddP[k >> 1] : s.evenP[k >> 1];
        acc = fmaf(s.taps[off + i], p, acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (lane4 == 0) s.tile[band][f] = sqrtf(log10f(acc + 1.0f));
    }
    // the next frame's stage-1 writes touch exA only; its first barrier orders them after this
    // frame's evenP/oddP reads, and the barrier above ordered the C reads before them.
  }
  __syncthreads();

  float* out = mel + ((int64_t)w * kMels) * kFrames + frame0;
  for (int idx = tid; idx < kMels * kFramesPerCta; idx += kThreads) {
    const int m = idx >> 5, f = idx & 31;
    out[(int64_t)m * kFrames + f] = s.tile[m][f];
  }
}

__global__ void pad_kernel(const float* __restrict__ src, int64_t n, float* __restrict__ dst, int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t j = i - kPadSamples;
    dst[i] = (j >= 0 && j < n) ? src[j] : 0.0f;
  }
}

__global__ void window_starts_kernel(int64_t* starts, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) starts[i] = i * kStepSamples;
}

}  // namespace

int features_init() {
  SS_CUDA_CHECK(cudaFuncSetAttribute(features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(Smem)));
  return SS_OK;
}

int launch_features_virtual(const ss_ctx* ctx, const float* pcm, int64_t valid_begin, int64_t valid_end,
                            int64_t offset, const int64_t* starts, int64_t w_base, int n_windows, float* mel,
                            cudaStream_t st) {
  if (n_windows <= 0) return SS_OK;
  SS_REQUIRE(ctx->fe.n_taps <= kMaxTapsSmem, SS_E_BLOB, "mel filterbank has %d taps (> %d)", ctx->fe.n_taps,
             kMaxTapsSmem);
  dim3 grid(n_windows, kFrames / kFramesPerCta);
  features_kernel<<<grid, kThreads, sizeof(Smem), st>>>(pcm, valid_begin, valid_end, offset, starts, w_base,
                                                        ctx->fe, mel);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int launch_features(const ss_ctx* ctx, const float* pcm, int64_t n_padded, const int64_t* starts, int n_windows,
                    float* mel, cudaStream_t st) {
  return launch_features_virtual(ctx, pcm, 0, n_padded, 0, starts, 0, n_windows, mel, st);
}

int launch_pad(const float* src, int64_t n, float* dst, cudaStream_t st) {
  const int64_t total = n + 2 * (int64_t)kPadSamples;
  const int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  pad_kernel<<<(int)blocks, threads, 0, st>>>(src, n, dst, total);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int launch_window_starts(int64_t* starts, int64_t n_windows, cudaStream_t st) {
  if (n_windows <= 0) return SS_OK;
  window_starts_kernel<<<(int)((n_windows + 255) / 256), 256, 0, st>>>(starts, n_windows);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

}  // namespace ss
