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
constexpr int kEvenPairs = 2 * 6 * 32;   // packed path: 6 x 32 pairs of W1024^m real parts, then of imaginary parts
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
  float2 tw1024[kEvenPairs];      // scalar path: [m]; packed path: [x pairs | y pairs], see the prologue
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

// ------------------------------------------------------------------------------------------------------------
// Packed pairs (helpers in ss_common.cuh).  sm_100 issues two fp32 operations per instruction on a 64-bit register pair
// (FADD2 / FMUL2 / FFMA2), each half rounded as the scalar instruction would be.  A lane's two columns of a transform
// (t = lane and t = lane + 32) go through the same instruction stream on independent data: the packed path keeps them
// as the two halves of a pair (.x: column lane, .y: column lane + 32), which halves the floating-point instructions of
// phase 1 (1,104 -> 512 of the kernel's 2,608 -> 2,200).  The contraction of every product into the sum that consumes it
// is spelled out below as the compiler appears to choose it for the scalar path; the two paths agree to 3.8e-7 of the
// largest feature (1 ulp-level differences: some contraction differs), NOT bit for bit — the packed path is the
// production path and defines K1's bits, the scalar one is kept as the statement it is checked against
// (tests/test_gpu_features.py).  K1: 1.006 -> 0.938 ms per 10-minute clip; the kernel is latency- rather than
// issue-bound (54 % issue utilisation at one 512-thread CTA per SM), so a quarter fewer instructions bought 7 %.

// The second half of dft8: from the four even-half and four odd-half sums to the eight outputs.
__device__ __forceinline__ void dft8p_finish(float2 (&re)[8], float2 (&im)[8], float2 e0r, float2 e0i, float2 e1r, float2 e1i,
                                             float2 e2r, float2 e2i, float2 e3r, float2 e3i, float2 o0r, float2 o0i, float2 o1r,
                                             float2 o1i, float2 o2r, float2 o2i, float2 o3r, float2 o3i) {
  const float2 c = bc2(0.70710678118654752440f), nc = bc2(-0.70710678118654752440f);
  const float2 a1 = add2(o1r, o1i), b1 = sub2(o1i, o1r);      // p1 = (a1 + i b1) c
  const float2 a3 = sub2(o3i, o3r), b3 = add2(o3r, o3i);      // p3 = (a3 - i b3) c
  re[0] = add2(e0r, o0r); im[0] = add2(e0i, o0i);
  re[4] = sub2(e0r, o0r); im[4] = sub2(e0i, o0i);
  re[1] = fma2(a1, c, e1r);  im[1] = fma2(b1, c, e1i);
  re[5] = fma2(a1, nc, e1r); im[5] = fma2(b1, nc, e1i);
  re[2] = add2(e2r, o2i); im[2] = sub2(e2i, o2r);
  re[6] = sub2(e2r, o2i); im[6] = add2(e2i, o2r);
  re[3] = fma2(a3, c, e3r);  im[3] = fma2(b3, nc, e3i);
  re[7] = fma2(a3, nc, e3r); im[7] = fma2(b3, c, e3i);
}

// dft8<false> on pairs (inputs already in registers)
__device__ __forceinline__ void dft8p(float2 (&re)[8], float2 (&im)[8]) {
  const float2 s0r = add2(re[0], re[4]), s0i = add2(im[0], im[4]);
  const float2 s1r = sub2(re[0], re[4]), s1i = sub2(im[0], im[4]);
  const float2 s2r = add2(re[2], re[6]), s2i = add2(im[2], im[6]);
  const float2 s3r = sub2(re[2], re[6]), s3i = sub2(im[2], im[6]);
  const float2 t0r = add2(re[1], re[5]), t0i = add2(im[1], im[5]);
  const float2 t1r = sub2(re[1], re[5]), t1i = sub2(im[1], im[5]);
  const float2 t2r = add2(re[3], re[7]), t2i = add2(im[3], im[7]);
  const float2 t3r = sub2(re[3], re[7]), t3i = sub2(im[3], im[7]);
  dft8p_finish(re, im, add2(s0r, s2r), add2(s0i, s2i), add2(s1r, s3i), sub2(s1i, s3r), sub2(s0r, s2r), sub2(s0i, s2i),
               sub2(s1r, s3i), add2(s1i, s3r), add2(t0r, t2r), add2(t0i, t2i), add2(t1r, t3i), sub2(t1i, t3r),
               sub2(t0r, t2r), sub2(t0i, t2i), sub2(t1r, t3i), add2(t1i, t3r));
}

// dft8<false> of v[j] = x[j] (wr[j] + i wi[j]): the products of j < 4 are fused into the first sums and differences
// (fma(x0, w0, x4 w4), fma(x0, w0, -(x4 w4))), those of j >= 4 are rounded on their own — as in the scalar path.
__device__ __forceinline__ void dft8p_windowed(const float2 (&x)[8], const float2 (&wr)[8], const float2 (&wi)[8],
                                               float2 (&re)[8], float2 (&im)[8]) {
  float2 pr[4], pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    pr[j] = mul2(x[j + 4], wr[j + 4]);
    pi[j] = mul2(x[j + 4], wi[j + 4]);
  }
  const float2 s0r = fma2(x[0], wr[0], pr[0]), s0i = fma2(x[0], wi[0], pi[0]);
  const float2 s1r = fma2(x[0], wr[0], neg2(pr[0])), s1i = fma2(x[0], wi[0], neg2(pi[0]));
  const float2 s2r = fma2(x[2], wr[2], pr[2]), s2i = fma2(x[2], wi[2], pi[2]);
  const float2 s3r = fma2(x[2], wr[2], neg2(pr[2])), s3i = fma2(x[2], wi[2], neg2(pi[2]));
  const float2 t0r = fma2(x[1], wr[1], pr[1]), t0i = fma2(x[1], wi[1], pi[1]);
  const float2 t1r = fma2(x[1], wr[1], neg2(pr[1])), t1i = fma2(x[1], wi[1], neg2(pi[1]));
  const float2 t2r = fma2(x[3], wr[3], pr[3]), t2i = fma2(x[3], wi[3], pi[3]);
  const float2 t3r = fma2(x[3], wr[3], neg2(pr[3])), t3i = fma2(x[3], wi[3], neg2(pi[3]));
  dft8p_finish(re, im, add2(s0r, s2r), add2(s0i, s2i), add2(s1r, s3i), sub2(s1i, s3r), sub2(s0r, s2r), sub2(s0i, s2i),
               sub2(s1r, s3i), add2(s1i, s3r), add2(t0r, t2r), add2(t0i, t2i), add2(t1r, t3i), sub2(t1i, t3r),
               sub2(t0r, t2r), sub2(t0i, t2i), sub2(t1r, t3i), add2(t1i, t3r));
}

// dft8<true> of v[j] = xe[j] we[j] + i xo[j] wo[j], j < 4 (v[4..7] = 0): the products of j < 2 are fused into the
// sums and differences that consume them, those of j = 2, 3 are rounded on their own — as in the scalar path.
__device__ __forceinline__ void dft8p_low4(const float2 (&xe)[4], const float2 (&xo)[4], const float2 (&we)[4],
                                           const float2 (&wo)[4], float2 (&re)[8], float2 (&im)[8]) {
  const float2 r2 = mul2(xe[2], we[2]), i2 = mul2(xo[2], wo[2]);
  const float2 r3 = mul2(xe[3], we[3]), i3 = mul2(xo[3], wo[3]);
  dft8p_finish(re, im,
               fma2(xe[0], we[0], r2), fma2(xo[0], wo[0], i2),                  // e0 = v0 + v2
               fma2(xe[0], we[0], i2), fma2(xo[0], wo[0], neg2(r2)),            // e1 = v0 - i v2
               fma2(xe[0], we[0], neg2(r2)), fma2(xo[0], wo[0], neg2(i2)),      // e2 = v0 - v2
               fma2(xe[0], we[0], neg2(i2)), fma2(xo[0], wo[0], r2),            // e3 = v0 + i v2
               fma2(xe[1], we[1], r3), fma2(xo[1], wo[1], i3),
               fma2(xe[1], we[1], i3), fma2(xo[1], wo[1], neg2(r3)),
               fma2(xe[1], we[1], neg2(r3)), fma2(xo[1], wo[1], neg2(i3)),
               fma2(xe[1], we[1], neg2(i3)), fma2(xo[1], wo[1], r3));
}

// (r + i i) *= (wx + i wy), per half
__device__ __forceinline__ void cmul2(float2& r, float2& i, float2 wx, float2 wy) {
  const float2 nr = fma2(r, wx, neg2(mul2(i, wy)));
  const float2 ni = fma2(r, wy, mul2(i, wx));
  r = nr;
  i = ni;
}

// fft512_tail on pairs: (re, im)[k2b].x = X[lane + 64 k2b], .y = X[lane + 32 + 64 k2b].
__device__ __forceinline__ void fft512_tail_p(WarpSmem& ws, const float2 (&tw2r)[7], int lane, float2 (&re)[8], float2 (&im)[8]) {
  const int k1 = lane >> 3, u = lane & 7;      // column lane + 32: k1 + 4, same u
  __syncwarp();
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    re[v] = make_float2(ws.re[k1 * 72 + u + 8 * v], ws.re[(k1 + 4) * 72 + u + 8 * v]);
    im[v] = make_float2(ws.im[k1 * 72 + u + 8 * v], ws.im[(k1 + 4) * 72 + u + 8 * v]);
  }
  __syncwarp();
  dft8p(re, im);
#pragma unroll
  for (int k2a = 0; k2a < 8; ++k2a) {
    if (k2a) cmul2(re[k2a], im[k2a], bc2(tw2r[k2a - 1].x), bc2(tw2r[k2a - 1].y));
    ws.re[(k2a * 8 + k1) * 9 + u] = re[k2a].x;
    ws.re[(k2a * 8 + k1 + 4) * 9 + u] = re[k2a].y;
    ws.im[(k2a * 8 + k1) * 9 + u] = im[k2a].x;
    ws.im[(k2a * 8 + k1 + 4) * 9 + u] = im[k2a].y;
  }
  __syncwarp();
#pragma unroll
  for (int uu = 0; uu < 8; ++uu) {
    re[uu] = make_float2(ws.re[lane * 9 + uu], ws.re[(lane + 32) * 9 + uu]);
    im[uu] = make_float2(ws.im[lane * 9 + uu], ws.im[(lane + 32) * 9 + uu]);
  }
  __syncwarp();
  dft8p(re, im);
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

// Phase 1 of one frame on packed pairs: what the scalar body of features_kernel does for (h = 0, h = 1) at once.
template <typename T>
__device__ __forceinline__ void frame_spectra_packed(Smem& s, WarpSmem& ws, const T* __restrict__ pcm, const T* __restrict__ src,
                                                     int64_t wstart, int l0, int64_t valid_begin, int64_t valid_end,
                                                     int64_t offset, bool fast, bool fast2, int lane,
                                                     const float2 (&tw1r)[2][7], const float2 (&tw2r)[7], float* __restrict__ Pf) {
  float2 re[8], im[8];
  const float2* __restrict__ twa_re = reinterpret_cast<const float2*>(s.tw_a_re) + lane;
  const float2* __restrict__ twa_im = reinterpret_cast<const float2*>(s.tw_a_im) + lane;
  const float2* __restrict__ win_e = reinterpret_cast<const float2*>(s.win) + lane;
  const float2* __restrict__ win_o = win_e + 128;
  // ------------------------------- FFT_A: a[n] = x[n] w[n] e^{-2 pi i n / 2048}
  {
    float2 x[8], wr[8], wi[8];
    if (fast) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = make_float2(ld1(src + lane + 64 * j), ld1(src + lane + 32 + 64 * j));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        x[j] = make_float2(load_sample(pcm, wstart, l0 + lane + 64 * j, valid_begin, valid_end, offset),
                           load_sample(pcm, wstart, l0 + lane + 32 + 64 * j, valid_begin, valid_end, offset));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      wr[j] = twa_re[32 * j];
      wi[j] = twa_im[32 * j];
    }
    dft8p_windowed(x, wr, wi, re, im);
  }
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    if (k1)
      cmul2(re[k1], im[k1], make_float2(tw1r[0][k1 - 1].x, tw1r[1][k1 - 1].x), make_float2(tw1r[0][k1 - 1].y, tw1r[1][k1 - 1].y));
    ws.re[k1 * 72 + lane] = re[k1].x;
    ws.re[k1 * 72 + lane + 32] = re[k1].y;
    ws.im[k1 * 72 + lane] = im[k1].x;
    ws.im[k1 * 72 + lane + 32] = im[k1].y;
  }
  fft512_tail_p(ws, tw2r, lane, re, im);
#pragma unroll
  for (int k2b = 0; k2b < 8; ++k2b) {
    const int q0 = lane + 64 * k2b, q1 = q0 + 32;
    const float2 pw = fma2(re[k2b], re[k2b], mul2(im[k2b], im[k2b]));
    if (q0 < kOddQ) Pf[q0 * kRowStride] = pw.x;                              // bin 4q+1
    if (q0 > 511 - kOddQ) Pf[(kOddQ + 511 - q0) * kRowStride] = pw.x;        // bin 4(511-q)+3 (conjugate symmetry)
    if (q1 < kOddQ) Pf[q1 * kRowStride] = pw.y;
    if (q1 > 511 - kOddQ) Pf[(kOddQ + 511 - q1) * kRowStride] = pw.y;
  }
  // ------------------------------- FFT_B: c[n] = x[2n] w[2n] + i x[2n+1] w[2n+1], n < 256 (rest zero)
  {
    float2 xe[4], xo[4], we[4], wo[4];
    if (fast2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = ld2(src + 2 * (lane + 64 * j)), b = ld2(src + 2 * (lane + 32 + 64 * j));
        xe[j] = make_float2(a.x, b.x);
        xo[j] = make_float2(a.y, b.y);
      }
    } else if (fast) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xe[j] = make_float2(ld1(src + 2 * (lane + 64 * j)), ld1(src + 2 * (lane + 32 + 64 * j)));
        xo[j] = make_float2(ld1(src + 2 * (lane + 64 * j) + 1), ld1(src + 2 * (lane + 32 + 64 * j) + 1));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        xe[j] = make_float2(load_sample(pcm, wstart, l0 + 2 * (lane + 64 * j), valid_begin, valid_end, offset),
                            load_sample(pcm, wstart, l0 + 2 * (lane + 32 + 64 * j), valid_begin, valid_end, offset));
        xo[j] = make_float2(load_sample(pcm, wstart, l0 + 2 * (lane + 64 * j) + 1, valid_begin, valid_end, offset),
                            load_sample(pcm, wstart, l0 + 2 * (lane + 32 + 64 * j) + 1, valid_begin, valid_end, offset));
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      we[j] = win_e[32 * j];
      wo[j] = win_o[32 * j];
    }
    dft8p_low4(xe, xo, we, wo, re, im);
  }
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    if (k1)
      cmul2(re[k1], im[k1], make_float2(tw1r[0][k1 - 1].x, tw1r[1][k1 - 1].x), make_float2(tw1r[0][k1 - 1].y, tw1r[1][k1 - 1].y));
    ws.re[k1 * 72 + lane] = re[k1].x;
    ws.re[k1 * 72 + lane + 32] = re[k1].y;
    ws.im[k1 * 72 + lane] = im[k1].x;
    ws.im[k1 * 72 + lane + 32] = im[k1].y;
  }
  fft512_tail_p(ws, tw2r, lane, re, im);
  // ---- even bins (see the scalar path): the partner C[512 - m] of column lane sits in lane 32 - lane's column
  // lane' + 32 and vice versa, i.e. in the OTHER half of the pair (7 - k2b) there.
  {
    const int src_lane = (32 - lane) & 31;
    const float2* __restrict__ twx = s.tw1024 + lane;      // pairs of real parts, then (at + 192) of imaginary parts
#pragma unroll
    for (int k2b = 0; k2b < 6; ++k2b) {
      const int m0 = lane + 64 * k2b, m1 = m0 + 32;
      float2 nr = make_float2(__shfl_sync(0xffffffffu, re[7 - k2b].y, src_lane), __shfl_sync(0xffffffffu, re[7 - k2b].x, src_lane));
      float2 ni = make_float2(__shfl_sync(0xffffffffu, im[7 - k2b].y, src_lane), __shfl_sync(0xffffffffu, im[7 - k2b].x, src_lane));
      if (lane == 0) {
        nr = make_float2(re[(8 - k2b) & 7].x, re[7 - k2b].y);
        ni = make_float2(im[(8 - k2b) & 7].x, im[7 - k2b].y);
      }
      // with n = conj C[512 - m] = nr - i ni:  e = (c + n) / 2,  o = (c - n) / (2 i) = (di - i dr) / 2
      const float2 sr = add2(re[k2b], nr), si = sub2(im[k2b], ni);
      const float2 dr = sub2(re[k2b], nr), di = add2(im[k2b], ni);
      float2 orr = mul2(di, bc2(0.5f)), oi = mul2(dr, bc2(-0.5f));
      cmul2(orr, oi, twx[32 * k2b], twx[192 + 32 * k2b]);
      const float2 ur = fma2(sr, bc2(0.5f), orr), ui = fma2(si, bc2(0.5f), oi);
      const float2 pw = fma2(ur, ur, mul2(ui, ui));
      if (m0 < kEvenBins) Pf[(2 * kOddQ + m0) * kRowStride] = pw.x;
      if (m1 < kEvenBins) Pf[(2 * kOddQ + m1) * kRowStride] = pw.y;
    }
  }
}


// kPk: phase 1 on packed pairs (FADD2 / FMUL2 / FFMA2, see above; default); the scalar path is kept for A/B runs
// (SS_K1_PACKED=0) and as the statement the packed one is checked against.
template <typename T, bool kPk>
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
  if constexpr (kPk) {
    // tables as pairs (column lane, column lane + 32), one 8-byte load per pair:
    //   tw_a_re / tw_a_im [j * 32 + lane] = table[lane + 64 j], table[lane + 32 + 64 j]          (j < 8)
    //   win [j * 32 + lane] = window[2 n], window[2 (n + 32)], n = lane + 64 j (j < 4); [128 + ...] = the odd samples'
    //   tw1024 [k * 32 + lane] = Re W1024^m, Re W1024^(m + 32), m = lane + 64 k (k < 6); [192 + ...] = the imaginary parts
    for (int i = tid; i < 512; i += kThreads) {
      const int h = i & 1, l = (i >> 1) & 31, j = i >> 6;
      s.tw_a_re[i] = fe.tw_a_re[l + 32 * h + 64 * j];
      s.tw_a_im[i] = fe.tw_a_im[l + 32 * h + 64 * j];
      const int odd = j >> 2;
      s.win[i] = fe.window[2 * (l + 32 * h + 64 * (j & 3)) + odd];
    }
    for (int i = tid; i < 2 * kEvenPairs; i += kThreads) {       // i counts floats: [part][k][lane][h]
      const int h = i & 1, l = (i >> 1) & 31, k = (i >> 6) % 6, part = i / (6 * 64);
      const float2 w = fe.tw1024[l + 32 * h + 64 * k];       // m < 384: inside the 512-entry table
      reinterpret_cast<float*>(s.tw1024)[i] = part ? w.y : w.x;
    }
  } else {
    for (int i = tid; i < 512; i += kThreads) {
      s.tw_a_re[i] = fe.tw_a_re[i];
      s.tw_a_im[i] = fe.tw_a_im[i];
      s.win[i] = fe.window[i];
    }
    for (int i = tid; i < kEvenBins; i += kThreads) s.tw1024[i] = fe.tw1024[i];
  }
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

      if constexpr (kPk) {
        frame_spectra_packed<T>(s, ws, pcm, src, wstart, l0, valid_begin, valid_end, offset, fast, fast2, lane, tw1r, tw2r, Pf);
        __syncwarp();
        continue;
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
  SS_CUDA_CHECK(cudaFuncSetAttribute(features_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(Smem)));
  SS_CUDA_CHECK(cudaFuncSetAttribute(features_kernel<int16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(Smem)));
  SS_CUDA_CHECK(cudaFuncSetAttribute(features_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(Smem)));
  SS_CUDA_CHECK(cudaFuncSetAttribute(features_kernel<int16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
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
  auto launch = [&](auto kernel, auto* samples) {
    kernel<<<grid, kThreads, sizeof(Smem), st>>>(samples, valid_begin, valid_end, offset, starts, w_base, n_tiles, ctx->fe, mel);
  };
  if (sample_fmt == kSampleS16) {
    if (ctx->fe.packed) launch(features_kernel<int16_t, true>, static_cast<const int16_t*>(pcm));
    else launch(features_kernel<int16_t, false>, static_cast<const int16_t*>(pcm));
  } else {
    if (ctx->fe.packed) launch(features_kernel<float, true>, static_cast<const float*>(pcm));
    else launch(features_kernel<float, false>, static_cast<const float*>(pcm));
  }
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
