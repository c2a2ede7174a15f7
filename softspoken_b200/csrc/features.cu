// K1 — fused log-mel front end.
//
// Replaces, for every 3 s window, the reference's window gather + torch.stack
// (root/code/frontend/NNDetector.py:90-96), torchaudio MelSpectrogram (reflect pad, 2048-point STFT of
// 512-tap Hann frames at hop 256, |.|^2, 128-band HTK mel; root/code/backend/pytorch_neural_nets.py:92-99,144),
// sqrt(log10(x + 1)) (:147) and the trim to 256 frames (:150).
//
// Data flow: PCM is read in place from the (optionally virtual) padded clip — no [W, 66150] window matrix is
// ever materialised.  The kernel is persistent (one 512-thread CTA per SM, tables loaded once); a work tile is
// 32 consecutive frames of one window.  Phase 1: each of the 16 warps takes 2 of the frames and works alone (no
// block barrier, __syncwarp only).  Per frame the 2048-point spectrum of the 512 real taps comes from two
// 512-point complex FFTs, 16 points per lane (two radix-8 butterflies per pass, three passes, two exchanges
// through a private 4.6 KB shared-memory buffer with conflict-free strides):
//   FFT_A of a[n] = x[n] w[n] e^{-2 pi i n / 2048}         -> bins 4q+1 (A[q]) and 4q+3 (conj A[511-q])
//   FFT_B of c[n] = x[2n] w[2n] + i x[2n+1] w[2n+1], n<256 -> even bins 2m by the real-FFT split
// Only bins 1..743 are formed (the filterbank is zero elsewhere); their powers go to a [bin][frame] tile in
// shared memory (row stride 33, rows ordered so that every store is conflict-free).  Phase 2 (after one block
// barrier): lane = frame; each warp walks the sparse triangular taps of 8 bands with one broadcast weight load
// and one conflict-free power load per tap, and stores sqrt(log10(1 + mel)) as full 128-byte rows of the
// [W][128][256] feature tensor.
#include "ss_common.cuh"

namespace ss {

namespace {

constexpr int kWarps = kFeatureWarps;
constexpr int kThreads = kWarps * 32;
constexpr int kFramesPerTile = 32;
constexpr int kFramesPerWarp = kFramesPerTile / kWarps;
constexpr int kBandsPerWarp = kMels / kWarps;
constexpr int kEx = 8 * 72;     // exchange A: [k1][t] with row stride 72; exchange B: [k2a*8+k1][u] with row stride 9
constexpr int kEvenBins = 372;  // even bins 2m, m <= 371  (k <= 742)
constexpr int kOddQ = 186;      // bins 4q+1 and 4q+3, q <= 185 (k <= 743)
constexpr int kRows = 2 * kOddQ + kEvenBins;   // power rows: [4q+1 | 4q+3 | 2m]
constexpr int kRowStride = kFramesPerTile + 1;
constexpr int kMaxTaps = 2048;
constexpr int kMaxRec = kMaxMelRec;

struct WarpSmem {
  float re[kEx];
  float im[kEx];
};

struct Smem {
  float2 tw1[7][64];            // W512^(t k1), k1 = 1..7
  float2 tw2[7][64];            // W64^(u k2a), k2a = 1..7, u = t & 7
  float tw_a_re[512], tw_a_im[512], win[512];
  float2 tw1024[kEvenBins];
  float4 rec[kMaxRec];           // two-band walk (FrontEnd::mel_rec) with .z = row offset of the bin in the power tile
  int rec_begin[kWarps + 1];
  float P[kRows * kRowStride];  // power spectrum tile [row(bin)][frame]
  WarpSmem w[kWarps];
};

// Row of bin k (1 <= k <= 743) in the power tile.
__device__ __forceinline__ int row_of_bin(int k) {
  return (k & 1) ? (((k & 2) ? kOddQ : 0) + (k >> 2)) : (2 * kOddQ + (k >> 1));
}

// In-place forward 8-point DFT: y[k] = sum_j v[j] exp(-2 pi i j k / 8).  kLow4: v[4..7] are zero (not read).
template <bool kLow4>
__device__ __forceinline__ void dft8(float (&re)[8], float (&im)[8]) {
  const float c = 0.70710678118654752440f;
  float e0r, e0i, e1r, e1i, e2r, e2i, e3r, e3i, o0r, o0i, o1r, o1i, o2r, o2i, o3r, o3i;
  if constexpr (kLow4) {
    e0r = re[0] + re[2]; e0i = im[0] + im[2];
    e2r = re[0] - re[2]; e2i = im[0] - im[2];
    e1r = re[0] + im[2]; e1i = im[0] - re[2];
    e3r = re[0] - im[2]; e3i = im[0] + re[2];
    o0r = re[1] + re[3]; o0i = im[1] + im[3];
    o2r = re[1] - re[3]; o2i = im[1] - im[3];
    o1r = re[1] + im[3]; o1i = im[1] - re[3];
    o3r = re[1] - im[3]; o3i = im[1] + re[3];
  } else {
    // even half: v0 v2 v4 v6
    const float s0r = re[0] + re[4], s0i = im[0] + im[4];
    const float s1r = re[0] - re[4], s1i = im[0] - im[4];
    const float s2r = re[2] + re[6], s2i = im[2] + im[6];
    const float s3r = re[2] - re[6], s3i = im[2] - im[6];
    e0r = s0r + s2r; e0i = s0i + s2i;
    e2r = s0r - s2r; e2i = s0i - s2i;
    e1r = s1r + s3i; e1i = s1i - s3r;
    e3r = s1r - s3i; e3i = s1i + s3r;
    // odd half: v1 v3 v5 v7
    const float t0r = re[1] + re[5], t0i = im[1] + im[5];
    const float t1r = re[1] - re[5], t1i = im[1] - im[5];
    const float t2r = re[3] + re[7], t2i = im[3] + im[7];
    const float t3r = re[3] - re[7], t3i = im[3] - im[7];
    o0r = t0r + t2r; o0i = t0i + t2i;
    o2r = t0r - t2r; o2i = t0i - t2i;
    o1r = t1r + t3i; o1i = t1i - t3r;
    o3r = t1r - t3i; o3i = t1i + t3r;
  }
  // twiddles W8^k
  const float p1r = (o1r + o1i) * c, p1i = (o1i - o1r) * c;   // (1 - i)/sqrt2
  const float p2r = o2i, p2i = -o2r;                          // -i
  const float p3r = (o3i - o3r) * c, p3i = -(o3r + o3i) * c;  // (-1 - i)/sqrt2
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
  const float nr = r * w.x - i * w.y;
  const float ni = r * w.y + i * w.x;
  r = nr;
  i = ni;
}

// Sample `l` (relative to the window start, may be negative for frame 0 -> torch 'reflect') of the
// virtual padded clip: indices inside [valid_begin, valid_end) map to pcm[idx - offset], the rest are 0.
// Sample types: float32 (what `load_audio` returns) or the int16 of a PCM_16 file, decoded on the fly exactly as
// libsndfile's float read does (value / 32768; voice_activity.py:37) — int16 -> float and the power-of-two scale are
// both exact, so the two sample types give bit-identical features.
__device__ __forceinline__ float ld1(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld1(const int16_t* p) { return (float)__ldg(p) * (1.0f / 32768.0f); }
__device__ __forceinline__ float2 ld2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 ld2(const int16_t* p) {
  const short2 v = __ldg(reinterpret_cast<const short2*>(p));
  return make_float2((float)v.x * (1.0f / 32768.0f), (float)v.y * (1.0f / 32768.0f));
}

template <typename T>
__device__ __forceinline__ float load_sample(const T* __restrict__ pcm, int64_t wstart, int l,
                                             int64_t valid_begin, int64_t valid_end, int64_t offset) {
  const int64_t idx = wstart + (l < 0 ? -l : l);
  return (idx >= valid_begin && idx < valid_end) ? ld1(pcm + (idx - offset)) : 0.0f;
}

// Passes 2 and 3 of the 512-point FFT whose pass-1 results sit in the warp's exchange buffer (A layout).
// On return (re, im)[h][k2b] = X[t + 64 k2b] with t = lane + 32 h.
// tw2r[k - 1] = W64^(u k), u = lane & 7: the pass-2 twiddles depend on t = lane + 32 h only through t & 7, so both
// halves use the same seven values — kept in registers by the caller (14 shared-memory wavefronts per transform less).
__device__ __forceinline__ void fft512_tail(WarpSmem& ws, const float2 (&tw2r)[7], int lane, float (&re)[2][8], float (&im)[2][8]) {
  __syncwarp();
  // ---- pass 2: butterfly (k1, u) = (t >> 3, t & 7) transforms over v, twiddle W64^(u k2a)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int t = lane + 32 * h, k1 = t >> 3, u = t & 7;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      re[h][v] = ws.re[k1 * 72 + u + 8 * v];
      im[h][v] = ws.im[k1 * 72 + u + 8 * v];
    }
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int t = lane + 32 * h, k1 = t >> 3, u = t & 7;
    dft8<false>(re[h], im[h]);
#pragma unroll
    for (int k2a = 0; k2a < 8; ++k2a) {
      if (k2a) cmul(re[h][k2a], im[h][k2a], tw2r[k2a - 1]);
      ws.re[(k2a * 8 + k1) * 9 + u] = re[h][k2a];
      ws.im[(k2a * 8 + k1) * 9 + u] = im[h][k2a];
    }
  }
  __syncwarp();
  // ---- pass 3: butterfly q0 = t = k1 + 8 k2a transforms over u -> X[q0 + 64 k2b]
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int t = lane + 32 * h;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      re[h][u] = ws.re[t * 9 + u];
      im[h][u] = ws.im[t * 9 + u];
    }
  }
  __syncwarp();                 // the buffer is free again once every lane has read
#pragma unroll
  for (int h = 0; h < 2; ++h) dft8<false>(re[h], im[h]);
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
features_kernel(const T* __restrict__ pcm, int64_t valid_begin, int64_t valid_end, int64_t offset,
                const int64_t* __restrict__ starts, int64_t w_base, int n_tiles, FrontEnd fe, float* __restrict__ mel) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  // ---- CTA-resident tables
  for (int i = tid; i < 7 * 64; i += kThreads) {
    const int k = i / 64 + 1, t = i % 64;
    s.tw1[k - 1][t] = fe.tw512[(t * k) & 511];
    s.tw2[k - 1][t] = fe.tw512[(8 * (t & 7) * k) & 511];
  }
  for (int i = tid; i < 512; i += kThreads) {
    s.tw_a_re[i] = fe.tw_a_re[i];
    s.tw_a_im[i] = fe.tw_a_im[i];
    s.win[i] = fe.window[i];
  }
  for (int i = tid; i < kEvenBins; i += kThreads) s.tw1024[i] = fe.tw1024[i];
  for (int i = tid; i < fe.n_rec; i += kThreads) {
    float4 r = fe.mel_rec[i];
    r.z = __int_as_float(row_of_bin(__float_as_int(r.z)) * kRowStride);
    s.rec[i] = r;
  }
  if (tid <= kWarps) s.rec_begin[tid] = fe.n_rec > 0 ? fe.mel_rec_begin[tid] : 0;
  __syncthreads();

  WarpSmem& ws = s.w[warp];
  float re[2][8], im[2][8];
  float2 tw2r[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) tw2r[k] = s.tw2[k][lane & 7];
  // ... and the pass-1 twiddles W512^(t k1) of the lane's two columns t = lane, lane + 32
  float2 tw1r[2][7];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int k = 0; k < 7; ++k) tw1r[h][k] = s.tw1[k][lane + 32 * h];

#pragma unroll 1
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int w = tile >> 3;                             // 8 tiles of 32 frames per window
    const int frame0 = (tile & 7) * kFramesPerTile;
    const int64_t wstart = starts ? starts[w] : (w_base + w) * (int64_t)kStepSamples;

    // ======================================================================= phase 1: spectra (per warp)
#pragma unroll 1
    for (int fi = 0; fi < kFramesPerWarp; ++fi) {
      const int f = warp * kFramesPerWarp + fi;
      const int l0 = (frame0 + f) * kHop - kHop;   // first sample of the frame relative to the window
      const int64_t g0 = wstart + l0;
      const bool fast = (l0 >= 0) && (g0 >= valid_begin) && (g0 + kWin <= valid_end);   // warp-uniform
      const T* __restrict__ src = pcm + (g0 - offset);
      const bool fast2 = fast && (reinterpret_cast<uintptr_t>(src) & (2 * sizeof(T) - 1)) == 0;   // pair-aligned frame
      float* __restrict__ Pf = s.P + f;
      {
        // pull the samples of the next frame this warp will transform into L1 while this frame computes
        const int ntile = (fi + 1 < kFramesPerWarp) ? tile : tile + (int)gridDim.x;
        if (ntile < n_tiles && lane < (int)(kWin * sizeof(T) / 128) + 1) {
          const int nw = ntile >> 3;
          const int nf = (ntile & 7) * kFramesPerTile + warp * kFramesPerWarp + ((fi + 1) % kFramesPerWarp);
          const int64_t nws = starts ? starts[nw] : (w_base + nw) * (int64_t)kStepSamples;
          const int64_t ng = nws + (int64_t)nf * kHop - kHop + (int64_t)(128 / sizeof(T)) * lane;   // one 128 B line per lane
          if (ng >= valid_begin && ng < valid_end) asm volatile("prefetch.global.L1 [%0];" ::"l"(pcm + (ng - offset)));
        }
      }

      // ------------------------------- FFT_A: a[n] = x[n] w[n] e^{-2 pi i n / 2048}
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = lane + 32 * h;
        float x[8];
        if (fast) {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = ld1(src + t + 64 * j);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = load_sample(pcm, wstart, l0 + t + 64 * j, valid_begin, valid_end, offset);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = t + 64 * j;
          re[h][j] = x[j] * s.tw_a_re[n];
          im[h][j] = x[j] * s.tw_a_im[n];
        }
        dft8<false>(re[h], im[h]);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
          if (k1) cmul(re[h][k1], im[h][k1], tw1r[h][k1 - 1]);
          ws.re[k1 * 72 + t] = re[h][k1];
          ws.im[k1 * 72 + t] = im[h][k1];
        }
      }
      fft512_tail(ws, tw2r, lane, re, im);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int k2b = 0; k2b < 8; ++k2b) {
          const int q = lane + 32 * h + 64 * k2b;
          const float pw = re[h][k2b] * re[h][k2b] + im[h][k2b] * im[h][k2b];
          if (q < kOddQ) Pf[q * kRowStride] = pw;                              // bin 4q+1
          if (q > 511 - kOddQ) Pf[(kOddQ + 511 - q) * kRowStride] = pw;        // bin 4(511-q)+3 (conjugate symmetry)
        }
      }

      // ------------------------------- FFT_B: c[n] = x[2n] w[2n] + i x[2n+1] w[2n+1], n < 256 (rest zero)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int t = lane + 32 * h;
        float2 x[4];
        if (fast2) {
#pragma unroll
          for (int j = 0; j < 4; ++j) x[j] = ld2(src + 2 * (t + 64 * j));
        } else if (fast) {
#pragma unroll
          for (int j = 0; j < 4; ++j) x[j] = make_float2(ld1(src + 2 * (t + 64 * j)), ld1(src + 2 * (t + 64 * j) + 1));
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            x[j] = make_float2(load_sample(pcm, wstart, l0 + 2 * (t + 64 * j), valid_begin, valid_end, offset),
                               load_sample(pcm, wstart, l0 + 2 * (t + 64 * j) + 1, valid_begin, valid_end, offset));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 wv = *reinterpret_cast<const float2*>(&s.win[2 * (t + 64 * j)]);
          re[h][j] = x[j].x * wv.x;
          im[h][j] = x[j].y * wv.y;
        }
        dft8<true>(re[h], im[h]);
#pragma unroll
        for (int k1 = 0; k1 < 8; ++k1) {
          if (k1) cmul(re[h][k1], im[h][k1], tw1r[h][k1 - 1]);
          ws.re[k1 * 72 + t] = re[h][k1];
          ws.im[k1 * 72 + t] = im[h][k1];
        }
      }
      fft512_tail(ws, tw2r, lane, re, im);
      // ---- even bins: U[m] = (C[m] + conj C[512-m])/2 + W1024^m (C[m] - conj C[512-m])/(2i), m <= 371.
      // This lane holds C[m] for m = t + 64 k2b, t = lane + 32 h; C[512 - m] sits at column 64 - t, i.e. in lane
      // 32 - lane under (h ^ 1, 7 - k2b) — one shuffle per value instead of a round trip of all 512 values through the
      // exchange buffer (lane 0 keeps its own partners: t = 0 pairs with k2b' = 8 - k2b, t = 32 with itself).
      {
        const int src_lane = (32 - lane) & 31;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int k2b = 0; k2b < 6; ++k2b) {
            const int m = lane + 32 * h + 64 * k2b;
            float nr = __shfl_sync(0xffffffffu, re[h ^ 1][7 - k2b], src_lane);
            float ni = __shfl_sync(0xffffffffu, im[h ^ 1][7 - k2b], src_lane);
            if (lane == 0) {
              nr = h ? re[1][7 - k2b] : re[0][(8 - k2b) & 7];
              ni = h ? im[1][7 - k2b] : im[0][(8 - k2b) & 7];
            }
            ni = -ni;
            if (m < kEvenBins) {
              const float cr = re[h][k2b], ci = im[h][k2b];
              const float er = 0.5f * (cr + nr), ei = 0.5f * (ci + ni);
              const float dr = cr - nr, di = ci - ni;
              float orr = 0.5f * di, oi = -0.5f * dr;
              cmul(orr, oi, s.tw1024[m]);
              const float ur = er + orr, ui = ei + oi;
              Pf[(2 * kOddQ + m) * kRowStride] = ur * ur + ui * ui;
            }
          }
        }
      }
      __syncwarp();   // the exchange buffer is reused by the next frame
    }
    __syncthreads();

    // ======================================================================= phase 2: mel (lane = frame)
    float* __restrict__ out = mel + ((int64_t)w * kMels) * kFrames + frame0 + lane;
    const float* __restrict__ Pl = s.P + lane;
    if (fe.n_rec > 0) {
      // Two-band walk: a bin of a triangular bank lies in two consecutive bands, one even- and one odd-numbered.  The
      // warp walks the bins of its (contiguous) bands once — one broadcast record + one conflict-free power load per
      // bin — feeding both bands' sums, and stores a band when its last bin has passed.  Every band still sums its bins
      // in ascending order (a zero weight adds exactly nothing), so the result has the bits of the band-by-band walk
      // at half its shared-memory loads, which bound this kernel (74 % of the pipe's wavefronts).
      float acc_e = 0.f, acc_o = 0.f;
      const int r1 = s.rec_begin[warp + 1];
      int r = s.rec_begin[warp];
      float4 rec = r < r1 ? s.rec[r] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (; r < r1; ++r) {
        const float4 cur = rec;
        if (r + 1 < r1) rec = s.rec[r + 1];
        const float pw = Pl[__float_as_int(cur.z)];
        acc_e = fmaf(cur.x, pw, acc_e);
        acc_o = fmaf(cur.y, pw, acc_o);
        const int emit = __float_as_int(cur.w);
        if (emit) {                                     // warp-uniform
          if (emit & 0xff) {
            out[(int64_t)((emit & 0xff) - 1) * kFrames] = sqrtf(log10f(acc_e + 1.0f));
            acc_e = 0.f;
          }
          if (emit >> 8) {
            out[(int64_t)((emit >> 8) - 1) * kFrames] = sqrtf(log10f(acc_o + 1.0f));
            acc_o = 0.f;
          }
        }
      }
    } else {
      // any other sparse bank: band by band, taps straight from the blob (bands interleaved over the warps)
#pragma unroll 1
      for (int j = 0; j < kBandsPerWarp; ++j) {
        const int band = warp + kWarps * j;
        const int k0 = __ldg(fe.mel_start + band), cnt = __ldg(fe.mel_count + band), off = __ldg(fe.mel_offs + band);
        float acc = 0.f;
        for (int i = 0; i < cnt; ++i) acc = fmaf(__ldg(fe.mel_taps + off + i), Pl[row_of_bin(k0 + i) * kRowStride], acc);
        out[(int64_t)band * kFrames] = sqrtf(log10f(acc + 1.0f));
      }
    }
    __syncthreads();   // the power tile is rewritten by the next work tile
  }
}

// ------------------------------------------------------------------------------------------------------------
// K8 — review-screen spectrogram: |STFT| with n_fft = win_length = 512, hop 256, centred frames and
// zero padding, i.e. np.abs(librosa.stft(x, n_fft=512, win_length=512, hop_length=256))
// (root/code/backend/voice_activity.py:148-154; settings.py:4-6).  Output [257 bins][T = 1 + n / 256 frames] float32,
// frame t covering samples [256 t - 256, 256 t + 256).
//
// A warp transforms TWO consecutive frames with one 512-point complex FFT (the passes of K1): z[n] = w[n] (a[n] + i
// b[n]) with a, b the two real frames, then X_a[k] = (Z[k] + conj Z[512-k]) / 2 and X_b[k] = (Z[k] - conj Z[512-k]) / 2i.
// Eight warps = one tile of 16 frames; magnitudes go through a [bin][frame] shared tile (row stride 17) so that the
// global stores are 64-byte row segments of the frequency-major output.  HBM-bound by definition
// (4 B read + 4.02 B written per sample); the running maximum the dB stage needs is taken on the way.
constexpr int kSpecGroups = 2;                              // independent halves of the CTA (8 warps each)
constexpr int kSpecGroupWarps = kWarps / kSpecGroups;
constexpr int kSpecTileFrames = 2 * kSpecGroupWarps;        // 16 frames per group tile: two per warp
constexpr int kSpecRowStride = kSpecTileFrames + 1;

struct SpecSmem {
  float2 tw1[7][64];
  float2 tw2[7][64];
  float win[512];
  float mag[kSpecGroups][257 * kSpecRowStride];
  WarpSmem w[kWarps];
};

__device__ __forceinline__ void group_barrier(int group) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(kSpecGroupWarps * 32) : "memory");
}

// The two halves of the CTA run their own tile loops behind their own named barriers, so one half's transforms
// overlap the other's stores (one tile per CTA and a block barrier left every pipe under 30 % busy: 107 us per
// 10-minute clip).
template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
stft512_kernel(const T* __restrict__ pcm, int64_t n, int64_t n_frames, const float2* __restrict__ tw512,
               float* __restrict__ mag, unsigned int* __restrict__ max_bits) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SpecSmem& s = *reinterpret_cast<SpecSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 7 * 64; i += kThreads) {
    const int k = i / 64 + 1, t = i % 64;
    s.tw1[k - 1][t] = tw512[(t * k) & 511];
    s.tw2[k - 1][t] = tw512[(8 * (t & 7) * k) & 511];
  }
  for (int i = tid; i < 512; i += kThreads) s.win[i] = 0.5f - 0.5f * cospif((float)i * (1.0f / 256.0f));   // periodic Hann
  __syncthreads();
  const int group = warp / kSpecGroupWarps, gw = warp % kSpecGroupWarps;
  WarpSmem& ws = s.w[warp];
  float* __restrict__ tile_s = s.mag[group];
  float re[2][8], im[2][8];
  float2 tw2r[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) tw2r[k] = s.tw2[k][lane & 7];
  float vmax = 0.f;
  const int64_t n_tiles = (n_frames + kSpecTileFrames - 1) / kSpecTileFrames;
  const int64_t tile_step = (int64_t)gridDim.x * kSpecGroups;
#pragma unroll 1
  for (int64_t tile = (int64_t)blockIdx.x * kSpecGroups + group; tile < n_tiles; tile += tile_step) {
    const int64_t fa = tile * kSpecTileFrames + 2 * gw;       // frames fa (real part) and fa + 1 (imaginary part)
    const int64_t g0 = fa * kHop - kHop;                      // first sample of frame fa; frame fa + 1 starts kHop later
    const bool fast = g0 >= 0 && g0 + kHop + kWin <= n;       // warp-uniform: both frames inside the clip
    {
      // the 768 samples this warp reads in its next tile, pulled into L1 while this tile computes
      const int64_t ng = g0 + tile_step * kSpecTileFrames * kHop + (int64_t)(128 / sizeof(T)) * lane;
      if (lane < (int)((kHop + kWin) * sizeof(T) / 128) + 1 && ng >= 0 && ng < n)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pcm + ng));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int t = lane + 32 * h;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int l = t + 64 * j;
        float a, b;
        if (fast) {
          a = ld1(pcm + g0 + l);
          b = ld1(pcm + g0 + kHop + l);
        } else {
          const int64_t ia = g0 + l, ib = ia + kHop;
          a = (ia >= 0 && ia < n) ? ld1(pcm + ia) : 0.f;
          b = (ib >= 0 && ib < n) ? ld1(pcm + ib) : 0.f;
        }
        const float wv = s.win[l];
        re[h][j] = a * wv;
        im[h][j] = b * wv;
      }
      dft8<false>(re[h], im[h]);
#pragma unroll
      for (int k1 = 0; k1 < 8; ++k1) {
        if (k1) cmul(re[h][k1], im[h][k1], s.tw1[k1 - 1][t]);
        ws.re[k1 * 72 + t] = re[h][k1];
        ws.im[k1 * 72 + t] = im[h][k1];
      }
    }
    fft512_tail(ws, tw2r, lane, re, im);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int k2b = 0; k2b < 8; ++k2b) {
        const int q = lane + 32 * h + 64 * k2b;
        ws.re[q] = re[h][k2b];
        ws.im[q] = im[h][k2b];
      }
    }
    __syncwarp();
    for (int m = lane; m <= 256; m += 32) {
      const int mm = (512 - m) & 511;
      const float cr = ws.re[m], ci = ws.im[m], nr = ws.re[mm], ni = ws.im[mm];
      const float ar = cr + nr, ai = ci - ni;          // 2 X_a[m]
      const float br = cr - nr, bi = ci + ni;          // 2 i X_b[m]
      tile_s[m * kSpecRowStride + 2 * gw] = 0.5f * sqrtf(ar * ar + ai * ai);
      tile_s[m * kSpecRowStride + 2 * gw + 1] = 0.5f * sqrtf(br * br + bi * bi);
    }
    group_barrier(group);
    // a warp stores two rows at a time: lanes 0-15 one bin, lanes 16-31 the next, 16 frames (64 bytes) each
    const int64_t f = tile * kSpecTileFrames + (lane & 15);
    if (f < n_frames) {
      for (int row = 2 * gw + (lane >> 4); row <= 256; row += 2 * kSpecGroupWarps) {
        const float v = tile_s[row * kSpecRowStride + (lane & 15)];
        mag[(int64_t)row * n_frames + f] = v;
        vmax = fmaxf(vmax, v);
      }
    }
    group_barrier(group);      // the tile is rewritten by the group's next iteration
  }
  if (max_bits) {
#pragma unroll
    for (int d = 16; d; d >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, d));
    if (lane == 0) atomicMax(max_bits, __float_as_uint(vmax));      // magnitudes are >= 0: their bit patterns order like the values
  }
}

// The review screen's display transform (review_detections.py:880-881) on the magnitudes, in place:
//   np.abs(librosa.amplitude_to_db(S ** 2, ref=np.max))  — librosa squares its argument once more, so with p = S^4 and
// r = max(S)^4 (all float32, as numpy computes them) the value is |max(10 log10(max(1e-10, p)) - 10 log10(max(1e-10, r)),
// -80)|: 0 at the loudest cell, 80 at the floor.
__device__ __forceinline__ float spec_db_value(float m, float ref_db) {
  const float p1 = m * m;
  // numpy rounds the product and the difference separately; a fused multiply-subtract would leave 1e-6 at the
  // loudest cell, where the reference has exactly 0
  const float v = __fsub_rn(__fmul_rn(10.0f, log10f(fmaxf(1e-10f, p1 * p1))), ref_db);
  return fabsf(fmaxf(v, -80.0f));
}

__global__ void __launch_bounds__(256)
spec_db_kernel(float* __restrict__ mag, int64_t n_elems, const unsigned int* __restrict__ max_bits) {
  const float smax = __uint_as_float(*max_bits);
  const float r1 = smax * smax;
  const float ref_db = __fmul_rn(10.0f, log10f(fmaxf(1e-10f, r1 * r1)));
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // 16-byte vectors on the aligned body (two in flight per thread), scalars on the ragged head and tail
  int64_t head = (int64_t)(((16 - (reinterpret_cast<uintptr_t>(mag) & 15)) & 15) / 4);
  if (head > n_elems) head = n_elems;
  const int64_t n4 = (n_elems - head) / 4;
  float4* body = reinterpret_cast<float4*>(mag + head);
  for (int64_t i = tid; i < n4; i += 2 * stride) {
    const int64_t i2 = i + stride;
    float4 a = body[i], b = i2 < n4 ? body[i2] : make_float4(0.f, 0.f, 0.f, 0.f);
    a.x = spec_db_value(a.x, ref_db); a.y = spec_db_value(a.y, ref_db);
    a.z = spec_db_value(a.z, ref_db); a.w = spec_db_value(a.w, ref_db);
    body[i] = a;
    if (i2 < n4) {
      b.x = spec_db_value(b.x, ref_db); b.y = spec_db_value(b.y, ref_db);
      b.z = spec_db_value(b.z, ref_db); b.w = spec_db_value(b.w, ref_db);
      body[i2] = b;
    }
  }
  if (tid < head) mag[tid] = spec_db_value(mag[tid], ref_db);
  const int64_t tail0 = head + 4 * n4;
  if (tid < n_elems - tail0) mag[tail0 + tid] = spec_db_value(mag[tail0 + tid], ref_db);
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
  SS_CUDA_CHECK(cudaFuncSetAttribute(features_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(Smem)));
  SS_CUDA_CHECK(cudaFuncSetAttribute(features_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(Smem)));
  SS_CUDA_CHECK(cudaFuncSetAttribute(stft512_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(SpecSmem)));
  SS_CUDA_CHECK(cudaFuncSetAttribute(stft512_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(SpecSmem)));
  return SS_OK;
}

int launch_spectrogram(const ss_ctx* ctx, const void* pcm, int sample_fmt, int64_t n, float* mag, unsigned int* max_bits,
                       cudaStream_t st) {
  const int64_t n_frames = 1 + n / kHop;
  const int64_t n_ctas = (n_frames + kSpecGroups * kSpecTileFrames - 1) / (kSpecGroups * kSpecTileFrames);
  const int grid = (int)(n_ctas < kNumSMs ? n_ctas : kNumSMs);
  if (max_bits) SS_CUDA_CHECK(cudaMemsetAsync(max_bits, 0, sizeof(unsigned int), st));
  if (sample_fmt == kSampleS16)
    stft512_kernel<int16_t><<<grid, kThreads, sizeof(SpecSmem), st>>>(static_cast<const int16_t*>(pcm), n, n_frames,
                                                                      ctx->fe.tw512, mag, max_bits);
  else
    stft512_kernel<float><<<grid, kThreads, sizeof(SpecSmem), st>>>(static_cast<const float*>(pcm), n, n_frames,
                                                                    ctx->fe.tw512, mag, max_bits);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int launch_spectrogram_db(float* mag, int64_t n_elems, const unsigned int* max_bits, cudaStream_t st) {
  if (n_elems <= 0) return SS_OK;
  int64_t blocks = (n_elems / 4 + 255) / 256 + 1;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  spec_db_kernel<<<(int)blocks, 256, 0, st>>>(mag, n_elems, max_bits);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int launch_features_virtual(const ss_ctx* ctx, const void* pcm, int sample_fmt, int64_t valid_begin, int64_t valid_end,
                            int64_t offset, const int64_t* starts, int64_t w_base, int n_windows, float* mel,
                            cudaStream_t st) {
  if (n_windows <= 0) return SS_OK;
  SS_REQUIRE(ctx->fe.n_taps <= kMaxTaps, SS_E_BLOB, "mel filterbank has %d taps (> %d)", ctx->fe.n_taps, kMaxTaps);
  SS_REQUIRE(ctx->fe.n_rec <= kMaxRec, SS_E_BLOB, "mel filterbank walk has %d records (> %d)", ctx->fe.n_rec, kMaxRec);
  const int n_tiles = n_windows * (kFrames / kFramesPerTile);
  const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  if (sample_fmt == kSampleS16)
    features_kernel<int16_t><<<grid, kThreads, sizeof(Smem), st>>>(static_cast<const int16_t*>(pcm), valid_begin,
                                                                   valid_end, offset, starts, w_base, n_tiles, ctx->fe, mel);
  else
    features_kernel<float><<<grid, kThreads, sizeof(Smem), st>>>(static_cast<const float*>(pcm), valid_begin, valid_end,
                                                                 offset, starts, w_base, n_tiles, ctx->fe, mel);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int launch_features(const ss_ctx* ctx, const float* pcm, int64_t n_padded, const int64_t* starts, int n_windows,
                    float* mel, cudaStream_t st) {
  return launch_features_virtual(ctx, pcm, kSampleF32, 0, n_padded, 0, starts, 0, n_windows, mel, st);
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
