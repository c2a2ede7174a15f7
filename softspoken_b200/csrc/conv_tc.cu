// K2 — tcgen05 / TMEM implicit-GEMM implementation of the residual U-Net trunk (tensor-core modes).
//
// Replaces the same reference ops as conv_fp32.cu (ResBlock x11, MaxPool2d x4, Upsample x4, torch.cat x4;
// root/code/backend/pytorch_neural_nets.py:7-41,102-123,156-185) with 16-bit operands and fp32 accumulation
// on the 5th-generation tensor cores.  Three operand precisions share the kernel (conv_tc_kernel.cuh):
//   SS_MODE_BF16   bf16, one pass           — throughput mode, documented tolerance;
//   SS_MODE_F16    fp16, one pass           — same cost, 8x finer mantissa (weights pre-scaled by 2^k per layer);
//   SS_MODE_F16X3  fp16 hi/lo split, 3 MMAs — fp32-grade logits on tensor cores (the parity mode of the bench).
//
// Layout.  Every activation tensor is "planar-8, zero-padded": [B][C/8][H+2][W+2][8] 16-bit — one 16-byte
// vector of 8 channels per padded pixel, one plane per 8 channels, a one-pixel zero border around each
// image (split precision keeps a second tensor of the same shape with the fp16 residuals).  With
// q = y (W+2) + x the flattened padded position, a 3x3 convolution is a sum of nine shifted GEMMs:
//   out[q, :] = sum_tap  in[q + dy (W+2) + dx, :] . w[tap]   — no im2col, no boundary logic.
//
// Kernel.  One CTA per SM, persistent over work units of MT*128 consecutive positions x all N = C_out:
//   * producer warp: per 16-channel K-chunk, two 1-D bulk copies (cp.async.bulk, one per 8-channel plane)
//     bring the run of positions plus a (W+3)-position halo on each side into shared memory, a third brings
//     the chunk's packed weights; an mbarrier ring (full/empty) pipelines the stages;
//   * MMA warp: one thread issues tcgen05.mma (M=128, N, K=16, kind::f16 -> fp32 in TMEM).  Shared memory
//     holds K-major, un-swizzled core matrices (8 positions x 16 bytes), so the A operand of tap (dy, dx) is
//     the same buffer with the descriptor start address advanced by (dy (W+2) + dx) * 16 bytes: the halo
//     tile is loaded once and used nine times;
//   * further sources (the ResBlock's 1x1 residual branch on the block input; the lo tensors of the split
//     precision) accumulate into the same TMEM tile, so `out = relu(conv2(t) + residual(x))` is one launch;
//   * 4 epilogue warps: tcgen05.ld the fp32 accumulators, undo the weight scale, add the folded-BN bias,
//     ReLU, force the border positions to zero, pack to 16 bits and store 16-byte vectors.  With `upsample`
//     set each value is stored to the 2x2 block of the next level's tensor at a plane offset, which is how
//     nearest-Upsample and torch.cat([skip, up]) are realised without a pass of their own.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include <vector>

#include <type_traits>

#include "conv_tc_kernel.cuh"
#include "ss_common.cuh"

namespace ss {

namespace {

using namespace ss::tc;

constexpr int kGuardBytes = 1 << 17;      // slack before/after every activation allocation (halo over-reads)
constexpr size_t kSmemBudget = 224 * 1024; // dynamic shared memory of the persistent conv kernel (one CTA per SM)

// ------------------------------------------------------------------------------------ operand helpers
// 8 channels of one padded pixel as floats: hi (+ lo for the split format).
template <Prec P>
__device__ __forceinline__ void load8(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, int64_t pix,
                                      float (&f)[8]) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(hi) + pix);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float2 t = unpack2<P>(w[h]);
    f[2 * h] = t.x;
    f[2 * h + 1] = t.y;
  }
  if constexpr (PrecTraits<P>::split) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(lo) + pix);
    const uint32_t x[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float2 t = unpack2<P>(x[h]);
      f[2 * h] += t.x;            // exact: hi and lo together span <= 24 significant bits
      f[2 * h + 1] += t.y;
    }
  }
}

// Returns the packed maximum of the four hi words (range_track; meaningful for non-negative values in the fp16 modes).
template <Prec P>
__device__ __forceinline__ uint32_t store8(uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t pix,
                                           const float (&f)[8]) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    hw[h] = pack_hi<P>(f[2 * h], f[2 * h + 1]);
    if constexpr (PrecTraits<P>::split) lw[h] = pack_lo_f16(f[2 * h], f[2 * h + 1], hw[h]);
  }
  reinterpret_cast<uint4*>(hi)[pix] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
  if constexpr (PrecTraits<P>::split) reinterpret_cast<uint4*>(lo)[pix] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  if constexpr (PrecTraits<P>::fmt == 0) return range_track(range_track(hw[0], hw[1]), range_track(hw[2], hw[3]));
  else return 0u;
}

// ------------------------------------------------------------------------------------------ small kernels
// mel f32 [B][128][256] -> the first layer's operand tensor: a 16-channel planar tensor whose channel c < 9 is
// the mel image shifted by tap c = (dy+1)*3 + (dx+1) (zero outside the image), channels 9..15 zero.  The single
// input channel of conv1_1 is thereby unrolled into the GEMM K dimension (im2col in K): its 3x3 convolution is
// one K=16 MMA per tile instead of nine MMAs with 15/16 of K empty, and its 1x1 residual reads channel 4.
template <Prec P>
__global__ void mel_to_planar(const float* __restrict__ mel, uint16_t* __restrict__ out, uint16_t* __restrict__ out_lo,
                              int64_t n_pix) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pix) return;
  const int x = (int)(i % kFrames);
  const int y = (int)((i / kFrames) % kMels);
  const int64_t b = i / ((int64_t)kFrames * kMels);
  const int Wp = kFrames + 2, Hp = kMels + 2;
  const float* img = mel + b * (int64_t)kFrames * kMels;
  float f[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const int yy = y + c / 3 - 1, xx = x + c % 3 - 1;
    f[c] = (c < 9 && yy >= 0 && yy < kMels && xx >= 0 && xx < kFrames) ? __ldg(img + yy * kFrames + xx) : 0.f;
  }
  const int64_t pix = (b * 2) * Hp * Wp + (int64_t)(y + 1) * Wp + (x + 1);
  const float lo8[8] = {f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7]};
  const float hi8[8] = {f[8], f[9], f[10], f[11], f[12], f[13], f[14], f[15]};
  store8<P>(out, out_lo, pix, lo8);
  store8<P>(out, out_lo, pix + (int64_t)Hp * Wp, hi8);
}

// conv1_1's first convolution on CUDA cores: t = relu(conv3x3(mel) + b1), C_in = 1 -> 32 channels, written as the
// planar hi/lo operand tensor conv1_1's second convolution reads.  With one input channel the layer has nine MACs per
// output value — as a tensor-core launch it was bound by its epilogue (5 k cycles per 512 positions against 0.35 k
// of MMAs) and needed the im2col'd operand tensor of mel_to_planar; here it is a write-bound streaming kernel
// (4.3 MB per window out, 0.13 MB in) in exact fp32, and x0 / mel_to_planar disappear from the path.
// grid (128 / 8 row groups, 4 planes, B), 256 threads = the 256 frames of a mel row (a warp stores 512 contiguous
// bytes); a thread walks 8 rows with its 72 weights in registers and a sliding 3 x 3 window of mel values.
constexpr int kC1Rows = 8;     // mel rows per thread of conv1_direct (weights stay in registers across them)

template <Prec P>
__global__ void __launch_bounds__(kFrames)
conv1_direct(const float* __restrict__ mel, const float* __restrict__ w /* [9][32] */, const float* __restrict__ bias,
             uint16_t* __restrict__ out, uint16_t* __restrict__ out_lo, int* __restrict__ err) {
  const int x = threadIdx.x, y0 = blockIdx.x * kC1Rows, pl = blockIdx.y;
  const int64_t b = blockIdx.z;
  // channel pairs (k, k + 1) as the halves of packed registers: nine FFMA2 per two output values (the kernel sat at
  // 60 % of the issue rate and 70 % of HBM with scalar FFMAs — 16 instructions per 4 bytes written)
  float2 wr[9][4], br[4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) wr[t][k] = __ldg(reinterpret_cast<const float2*>(w + t * 32 + pl * 8) + k);
#pragma unroll
  for (int k = 0; k < 4; ++k) br[k] = __ldg(reinterpret_cast<const float2*>(bias + pl * 8) + k);
  const float* img = mel + b * (int64_t)kMels * kFrames;
  const int Wp = kFrames + 2, Hp = kMels + 2;
  auto row3 = [&](int yy, float (&r)[3]) {       // mel[yy][x-1 .. x+1], zero outside the image
    const bool in = (yy >= 0) && (yy < kMels);
    r[0] = (in && x > 0) ? __ldg(img + yy * kFrames + x - 1) : 0.f;
    r[1] = in ? __ldg(img + yy * kFrames + x) : 0.f;
    r[2] = (in && x < kFrames - 1) ? __ldg(img + yy * kFrames + x + 1) : 0.f;
  };
  float m[3][3];
  row3(y0 - 1, m[0]);
  row3(y0, m[1]);
  uint32_t hmax = 0u;
#pragma unroll
  for (int r = 0; r < kC1Rows; ++r) {
    row3(y0 + r + 1, m[(r + 2) % 3]);
    float f[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 acc = make_float2(0.f, 0.f);      // each half: fma(m, w, acc) tap by tap from zero, as the scalar loop did
#pragma unroll
      for (int t = 0; t < 9; ++t) acc = fma2(bc2(m[(r + t / 3) % 3][t % 3]), wr[t][k], acc);
      acc = add2(acc, br[k]);
      f[2 * k] = fmaxf(acc.x, 0.f);
      f[2 * k + 1] = fmaxf(acc.y, 0.f);
    }
    hmax = range_track(hmax, store8<P>(out, out_lo, ((b * 4 + pl) * Hp + (y0 + r + 1)) * (int64_t)Wp + (x + 1), f));
  }
  if (PrecTraits<P>::fmt == 0 && range_hit(hmax)) err[2] = 1;       // fp16 operand modes: the activation was saturated (SS_E_RANGE)
}

// MaxPool2d(2) on planar tensors: in planes [plane0, plane0+planes) of a tensor with in_planes_total planes at
// H x W  ->  out [B][planes][H/2+2][W/2+2][8].  The max is taken on the reconstructed fp32 values; in the split
// precision the pooled operand pair is the (hi, lo) pair OF the window's maximum, not a fresh split of hi + lo: the
// two differ when lo is exactly half an ulp of hi (one value in 4,096: hi + lo then sits midway between two fp16
// numbers and rounds to the even one), and keeping the pair makes this kernel and the pool folded into the
// convolution epilogue (TcConv::rows, which splits the float32 maximum it still holds) write the same bits.
// Tensors are addressed through (image stride, plane stride) in positions: image- or plane-major (Tensor).
template <Prec P>
__global__ void pool_planar(const uint16_t* __restrict__ in, const uint16_t* __restrict__ in_lo, int64_t in_img_stride,
                            int64_t in_plane_stride, int plane0, int planes, int H, int W, uint16_t* __restrict__ out,
                            uint16_t* __restrict__ out_lo, int64_t out_img_stride, int64_t out_plane_stride,
                            int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int W2 = W >> 1, H2 = H >> 1;
  const int x = (int)(i % W2);
  int64_t r = i / W2;
  const int y = (int)(r % H2); r /= H2;
  const int pl = (int)(r % planes);
  const int64_t b = r / planes;
  const int Wp = W + 2, Hp = H + 2, Wq = W2 + 2, Hq = H2 + 2;
  (void)Hp; (void)Hq;
  const int64_t src = b * in_img_stride + (plane0 + pl) * in_plane_stride + (int64_t)(2 * y + 1) * Wp + (2 * x + 1);
  const int64_t dst = b * out_img_stride + pl * out_plane_stride + (int64_t)(y + 1) * Wq + (x + 1);
  if constexpr (PrecTraits<P>::split) {
    const int64_t offs[4] = {src, src + 1, src + Wp, src + Wp + 1};
    uint4 hv[4], lv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      hv[k] = __ldg(reinterpret_cast<const uint4*>(in) + offs[k]);
      lv[k] = __ldg(reinterpret_cast<const uint4*>(in_lo) + offs[k]);
    }
    uint32_t bh[4] = {hv[0].x, hv[0].y, hv[0].z, hv[0].w}, bl[4] = {lv[0].x, lv[0].y, lv[0].z, lv[0].w};
    float2 best[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const float2 a = unpack2<P>(bh[h]), c = unpack2<P>(bl[h]);
      best[h] = make_float2(a.x + c.x, a.y + c.y);
    }
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      const uint32_t ch[4] = {hv[k].x, hv[k].y, hv[k].z, hv[k].w}, cl[4] = {lv[k].x, lv[k].y, lv[k].z, lv[k].w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 a = unpack2<P>(ch[h]), c = unpack2<P>(cl[h]);
        const float vx = a.x + c.x, vy = a.y + c.y;
        const uint32_t m = (vx > best[h].x ? 0x0000ffffu : 0u) | (vy > best[h].y ? 0xffff0000u : 0u);
        bh[h] = (bh[h] & ~m) | (ch[h] & m);
        bl[h] = (bl[h] & ~m) | (cl[h] & m);
        best[h].x = fmaxf(best[h].x, vx);
        best[h].y = fmaxf(best[h].y, vy);
      }
    }
    reinterpret_cast<uint4*>(out)[dst] = make_uint4(bh[0], bh[1], bh[2], bh[3]);
    reinterpret_cast<uint4*>(out_lo)[dst] = make_uint4(bl[0], bl[1], bl[2], bl[3]);
  } else {
    float a[8], c[8];
    load8<P>(in, in_lo, src, a);
    load8<P>(in, in_lo, src + 1, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaxf(a[k], c[k]);
    load8<P>(in, in_lo, src + Wp, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaxf(a[k], c[k]);
    load8<P>(in, in_lo, src + Wp + 1, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fmaxf(a[k], c[k]);
    store8<P>(out, out_lo, dst, a);
  }
}

// Mask head on planar conv9 [B][4][130][258][8] (same arithmetic as head.cu:mask_head_f32, fp32 math).
// grid (8 frame chunks, B): a CTA produces 32 logits; its 288 threads are 8 mel-row groups x 36 frames
// (32 + a 2-frame halo each side for the two k=3 1-D convolutions), so the K = 4096 reduction of
// conv_flatten is spread over 8 x more threads than one-thread-per-frame and the loads stay coalesced.
constexpr int kHeadFrames = 32, kHeadHalo = 2, kHeadCols = kHeadFrames + 2 * kHeadHalo, kHeadGroups = 8;

// Shared tail of the two mask-head kernels: group partials -> conv_flatten bias + ReLU -> ResBlock1D(4, 4) ->
// Conv1d(4, 1, 1) (pytorch_neural_nets.py:43-77,137-140,191-195), zero padding outside [0, 256).
__device__ __forceinline__ void mask_head_tail(float (&part)[kHeadGroups][4][kHeadCols], float (&xf)[4][kHeadCols],
                                               float (&c1)[4][kHeadCols], const HeadW& hw, int f, int hg, int t,
                                               bool valid, float* __restrict__ logits_b) {
  if (hg == 0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float v = 0.f;
#pragma unroll
      for (int g = 0; g < kHeadGroups; ++g) v += part[g][c][f];
      xf[c][f] = valid ? fmaxf(v + __ldg(hw.flat_b + c), 0.f) : 0.f;     // zero padding outside [0, 256)
    }
  }
  __syncthreads();
  if (hg == 0 && f >= 1 && f < kHeadCols - 1) {
#pragma unroll
    for (int co = 0; co < 4; ++co) {
      float v = __ldg(hw.c1_b + co);
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) v = fmaf(__ldg(hw.c1_w + (k * 4 + ci) * 4 + co), xf[ci][f + k - 1], v);
      c1[co][f] = valid ? fmaxf(v, 0.f) : 0.f;
    }
  }
  __syncthreads();
  if (hg == 0 && f >= kHeadHalo && f < kHeadCols - kHeadHalo) {
    float logit = __ldg(hw.out_b);
#pragma unroll
    for (int co = 0; co < 4; ++co) {
      float v = __ldg(hw.c2_b + co);
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) v = fmaf(__ldg(hw.c2_w + (k * 4 + ci) * 4 + co), c1[ci][f + k - 1], v);
      float r = __ldg(hw.res_b + co);
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) r = fmaf(__ldg(hw.res_w + ci * 4 + co), xf[ci][f], r);
      logit = fmaf(__ldg(hw.out_w + co), fmaxf(v + r, 0.f), logit);
    }
    logits_b[t] = logit;
  }
}

template <Prec P>
__global__ void __launch_bounds__(kHeadCols * kHeadGroups)
mask_head_planar(const uint16_t* __restrict__ conv9, const uint16_t* __restrict__ conv9_lo, HeadW hw,
                 float* __restrict__ logits) {
  __shared__ float part[kHeadGroups][4][kHeadCols];
  __shared__ float xf[4][kHeadCols];
  __shared__ float c1[4][kHeadCols];
  const int f = threadIdx.x % kHeadCols, hg = threadIdx.x / kHeadCols;
  const int b = blockIdx.y;
  const int t = blockIdx.x * kHeadFrames - kHeadHalo + f;          // frame of this column
  const bool valid = (t >= 0) && (t < kFrames);
  const int Wp = kFrames + 2, Hp = kMels + 2;
  const int64_t base = (int64_t)b * 4 * Hp * Wp;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (valid) {
    const int rows = kMels / kHeadGroups;
    // two mel rows per iteration, all eight 16-byte loads of both rows issued before their FMAs (latency-bound otherwise)
    for (int h = hg * rows; h < (hg + 1) * rows; h += 2) {
      float a[2][4][8];
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int pl = 0; pl < 4; ++pl)
          load8<P>(conv9, conv9_lo, base + ((int64_t)pl * Hp + (h + r + 1)) * Wp + (t + 1), a[r][pl]);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float4* wrow = reinterpret_cast<const float4*>(hw.flat_w + (int64_t)(h + r) * 32 * 4);
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 w = __ldg(wrow + pl * 8 + k);
            acc[0] = fmaf(a[r][pl][k], w.x, acc[0]);
            acc[1] = fmaf(a[r][pl][k], w.y, acc[1]);
            acc[2] = fmaf(a[r][pl][k], w.z, acc[2]);
            acc[3] = fmaf(a[r][pl][k], w.w, acc[3]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) part[hg][c][f] = acc[c];
  __syncthreads();
  mask_head_tail(part, xf, c1, hw, f, hg, t, valid, logits + (int64_t)b * kFrames);
}

// Mask head from the conv_flatten partials the conv9_1 epilogue wrote (TcConv::head_out, [B][128][256][4]):
// sum the 128 mel rows of each frame in a fixed order (8 groups of 16 rows, then the groups in order), then the same
// 1-D tail as mask_head_planar.  grid (8 frame chunks, B).
__global__ void __launch_bounds__(kHeadCols * kHeadGroups)
mask_head_partials(const float* __restrict__ part_in, HeadW hw, float* __restrict__ logits) {
  __shared__ float part[kHeadGroups][4][kHeadCols];
  __shared__ float xf[4][kHeadCols];
  __shared__ float c1[4][kHeadCols];
  const int f = threadIdx.x % kHeadCols, hg = threadIdx.x / kHeadCols;
  const int b = blockIdx.y;
  const int t = blockIdx.x * kHeadFrames - kHeadHalo + f;          // frame of this column
  const bool valid = (t >= 0) && (t < kFrames);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) {
    const int rows = kMels / kHeadGroups;
    const float4* src = reinterpret_cast<const float4*>(part_in) + ((int64_t)b * kMels + hg * rows) * kFrames + t;
    float4 v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = __ldg(src + (int64_t)r * kFrames);
#pragma unroll
    for (int r = 0; r < 16; ++r) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }
  }
  part[hg][0][f] = acc.x; part[hg][1][f] = acc.y; part[hg][2][f] = acc.z; part[hg][3][f] = acc.w;
  __syncthreads();
  mask_head_tail(part, xf, c1, hw, f, hg, t, valid, logits + (int64_t)b * kFrames);
}

// spec head tail on planar [B][4][130][258][8] -> NCHW f32 [B][2][128][256]
template <Prec P>
__global__ void spec_out_planar(const uint16_t* __restrict__ x, const uint16_t* __restrict__ x_lo, HeadW hw,
                                float* __restrict__ out, int64_t n_pixels) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pixels) return;
  const int xx = (int)(i % kFrames);
  const int yy = (int)((i / kFrames) % kMels);
  const int64_t b = i / ((int64_t)kFrames * kMels);
  const int Wp = kFrames + 2, Hp = kMels + 2;
  const int64_t base = (int64_t)b * 4 * Hp * Wp + (int64_t)(yy + 1) * Wp + (xx + 1);
  float a0 = __ldg(hw.spec_b), a1 = __ldg(hw.spec_b + 1);
#pragma unroll
  for (int pl = 0; pl < 4; ++pl) {
    float a[8];
    load8<P>(x, x_lo, base + (int64_t)pl * Hp * Wp, a);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      a0 = fmaf(a[k], __ldg(hw.spec_w + (pl * 8 + k) * 2), a0);
      a1 = fmaf(a[k], __ldg(hw.spec_w + (pl * 8 + k) * 2 + 1), a1);
    }
  }
  const int64_t plane = (int64_t)kMels * kFrames, pix = i % plane;
  out[(b * 2) * plane + pix] = fmaxf(a0, 0.f);
  out[(b * 2 + 1) * plane + pix] = fmaxf(a1, 0.f);
}

// planar -> NCHW f32 (debug / parity localisation only)
// halfrows: the tensor is a half-row tensor (TcSource::in_up): H / 2 + 2 rows, row 1 + y / 2 holds image row y
template <Prec P>
__global__ void planar_to_nchw(const uint16_t* __restrict__ in, const uint16_t* __restrict__ in_lo, int64_t img_stride,
                               int64_t plane_stride, int plane0, int C, int H, int W, float* __restrict__ out,
                               int64_t total, int halfrows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % W);
  int64_t r = i / W;
  const int y = (int)(r % H); r /= H;
  const int c = (int)(r % C);
  const int64_t b = r / C;
  const int Wp = W + 2;
  const int row = halfrows ? (y >> 1) + 1 : y + 1;
  float a[8];
  load8<P>(in, in_lo, b * img_stride + (plane0 + c / 8) * plane_stride + (int64_t)row * Wp + (x + 1), a);
  out[i] = a[c & 7];
}

// ------------------------------------------------------------------------------------------- host state
struct Tensor {
  uint16_t* alloc = nullptr;      // includes guards
  uint16_t* data = nullptr;
  uint16_t* alloc_lo = nullptr;   // split precision only
  uint16_t* lo = nullptr;
  int planes = 0, H = 0, W = 0;
  // plane-major: [C/8][cap][H+2][W+2][8] instead of [B][C/8][H+2][W+2][8] — the padded images of a batch are then
  // contiguous in every plane, which is what packed work units stage with one bulk copy (TcConv::packed)
  bool plane_major = false;
  int cap = 0;                    // images the tensor was allocated for
  int64_t hw() const { return (int64_t)(H + 2) * (W + 2); }
  int64_t img_stride() const { return plane_major ? hw() : planes * hw(); }        // in positions (16-byte vectors)
  int64_t plane_stride() const { return plane_major ? cap * hw() : hw(); }
};

struct PackedConv {
  uint16_t* w = nullptr;   // plain: [n_chunks][parts][taps][2][n][8];  dual: [n_chunks][taps][2][2n][8] (hi rows, lo rows)
  uint16_t* w_hi = nullptr;   // dual only: the hi parts alone, [n_chunks][taps][2][n][8] (for x_lo . w_hi)
  int n_chunks = 0, taps = 0, n = 0, parts = 1;
  bool dual = false;
};

struct TcBlock {
  PackedConv res, c1, c2;
  // Decoder blocks at 128 x 256 and 64 x 128 (split precision): conv1 of cat([skip, up(below)]) packed as two sources —
  // the skip channels with their nine taps, the up-sampled channels ("rowdup", conv_tc_kernel.cuh) with the twelve
  // row-merged tap matrices [top: w(-1,.), w(0,.)+w(+1,.) | bottom: w(-1,.)+w(0,.), w(+1,.)].  skip_cin = 0: not built.
  PackedConv c1_skip, c1_up;
  int skip_cin = 0;
  float* bias2 = nullptr;       // b2 + b_res (the fused second launch)
  const float* bias1 = nullptr;
  float inv_scale1 = 1.f, inv_scale2 = 1.f;
};

}  // namespace

struct TcState {
  Prec prec = Prec::Bf16;
  int max_batch = 0;
  TcBlock rb[RB_COUNT];
  Tensor x0, m4, m3, m2, m1, p1, p2, p3, p4, bott, c9, spec;
  // half-row tensors (split precision): up(conv8) for conv9_1 [B][4][64 + 2][258][8], up(conv7) for conv8
  // [B][8][32 + 2][130][8] — the up-sampled halves of the two largest decoder inputs with only their columns replicated
  Tensor u4, u3;
  bool halfrows_last = false;    // the most recent classify call used them (ss_debug_activation 7 / 8 read them then)
  Tensor t[RB_COUNT];
  int* err = nullptr;
  float* head_part = nullptr;  // [max_batch][128][256][4] conv_flatten partials written by conv9_1's epilogue
  int* flags = nullptr;        // per-unit completion counts of a fused ResBlock launch (TcJob)
  int* flags2 = nullptr;
  int flags_cap = 0;
  long long* prof = nullptr;   // [kNumSMs][8] role timers of the selected conv launch (debug)
  int prof_layer = -1;         // launch index to capture (-1: none)
  int launch_index = 0;
  size_t bytes = 0;
};

namespace {

bool is_split(Prec p) { return p == Prec::F16x3; }
// Layers whose doubled C_out still fits one MMA cheaply use the dual layout (see conv_tc_kernel.cuh).
bool is_dual(Prec p, int n) { return is_split(p) && n <= 64; }

// body: zero (the one-pixel borders are never written afterwards); guard bands before / after: kGuardPattern, a small
// finite number in either 16-bit format.  The first and last staged runs of a launch read into the bands (halo
// over-reads); those values only reach accumulators of border positions, which the epilogue forces to zero, or of
// positions past the tensor, which it does not store.  Nothing may WRITE there: ss_debug_check_guards verifies it.
// The tensor starts 16 bytes past a 32-byte boundary: the up-sampling epilogue writes every value to the position
// pair (2x - 1, 2x) of two rows, and with that offset a pair is one aligned 32-byte store (conv_tc_kernel.cuh:st32x2).
int alloc_guarded(ss_ctx* ctx, TcState* st, uint16_t** alloc, uint16_t** data, size_t body) {
  const size_t inner = body + 32;           // 16 bytes of slack either side of the tensor
  const size_t bytes = inner + 2 * (size_t)kGuardBytes;
  unsigned char* base = nullptr;
  SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&base), bytes));
  *alloc = reinterpret_cast<uint16_t*>(base);
  SS_CUDA_CHECK(cudaMemset(base, kGuardPattern, kGuardBytes));
  SS_CUDA_CHECK(cudaMemset(base + kGuardBytes, 0, inner));
  SS_CUDA_CHECK(cudaMemset(base + kGuardBytes + inner, kGuardPattern, kGuardBytes));
  *data = reinterpret_cast<uint16_t*>(base + kGuardBytes + 16);
  register_guard(ctx, base, kGuardBytes, base);
  register_guard(ctx, base + kGuardBytes + inner, kGuardBytes, base);
  st->bytes += bytes;
  return SS_OK;
}

int alloc_tensor(ss_ctx* ctx, TcState* st, Tensor* t, int B, int C, int H, int W, bool plane_major = false) {
  t->planes = C / 8;
  t->H = H;
  t->W = W;
  t->plane_major = plane_major;
  t->cap = B;
  const size_t body = (size_t)B * t->planes * (H + 2) * (W + 2) * 8 * sizeof(uint16_t);
  int rc = alloc_guarded(ctx, st, &t->alloc, &t->data, body);
  if (rc) return rc;
  if (is_split(st->prec)) rc = alloc_guarded(ctx, st, &t->alloc_lo, &t->lo, body);
  return rc;
}

uint16_t to_bits_bf16(float v) { __nv_bfloat16 h = __float2bfloat16(v); uint16_t b; memcpy(&b, &h, 2); return b; }
uint16_t to_bits_f16(float v) { __half h = __float2half_rn(v); uint16_t b; memcpy(&b, &h, 2); return b; }
float from_bits_f16(uint16_t b) { __half h; memcpy(&h, &b, 2); return __half2float(h); }

// Power-of-two scale that lifts the largest |w| of a group of fp16 weight tensors into [1024, 2048): the hi/lo
// split then keeps the residuals of all but the tiniest weights out of the fp16 subnormal range, and fp16
// single-pass weights keep their full 11-bit mantissa.  bf16 needs none (fp32 exponent range).
float weight_scale(Prec p, std::initializer_list<const std::vector<float>*> ws) {
  if (p == Prec::Bf16) return 1.f;
  float m = 0.f;
  for (const std::vector<float>* w : ws)
    for (float v : *w) m = fmaxf(m, fabsf(v));
  if (!(m > 0.f) || !isfinite(m)) return 1.f;
  int e;
  frexpf(m, &e);                         // m = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, 11 - e);            // m * scale in [1024, 2048)
}

// w: [taps][cin][cout] f32 (host) * scale -> packed 16-bit [chunk][part][tap][2][cout][8] on the device
int pack_conv(TcState* st, const float* w, int taps, int cin, int cout, float scale, PackedConv* out) {
  const int n_chunks = (cin + 15) / 16;
  const int parts = is_split(st->prec) ? 2 : 1;
  const bool dual = is_dual(st->prec, cout);
  std::vector<uint16_t> h((size_t)n_chunks * parts * taps * 2 * cout * 8);
  // element index of (chunk, part, tap, K-half, row, j) in the plain or the dual layout
  auto at = [&](int kc, int part, int t, int hh, int n, int j) -> size_t {
    if (dual) return (((((size_t)kc * taps + t) * 2 + hh) * 2 + part) * cout + n) * 8 + j;
    return (((((size_t)kc * parts + part) * taps + t) * 2 + hh) * cout + n) * 8 + j;
  };
  for (int kc = 0; kc < n_chunks; ++kc)
    for (int t = 0; t < taps; ++t)
      for (int hh = 0; hh < 2; ++hh)
        for (int n = 0; n < cout; ++n)
          for (int j = 0; j < 8; ++j) {
            const int c = kc * 16 + hh * 8 + j;
            const float v = (c < cin) ? w[((size_t)t * cin + c) * cout + n] * scale : 0.f;
            if (st->prec == Prec::Bf16) {
              h[at(kc, 0, t, hh, n, j)] = to_bits_bf16(v);
            } else {
              const uint16_t hi = to_bits_f16(v);
              h[at(kc, 0, t, hh, n, j)] = hi;
              if (parts == 2) h[at(kc, 1, t, hh, n, j)] = to_bits_f16(v - from_bits_f16(hi));
            }
          }
  SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&out->w), h.size() * sizeof(uint16_t)));
  SS_CUDA_CHECK(cudaMemcpy(out->w, h.data(), h.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  st->bytes += h.size() * sizeof(uint16_t);
  if (dual) {
    std::vector<uint16_t> hh_only((size_t)n_chunks * taps * 2 * cout * 8);
    for (int kc = 0; kc < n_chunks; ++kc)
      for (int t = 0; t < taps; ++t)
        for (int hh = 0; hh < 2; ++hh)
          for (int n = 0; n < cout; ++n)
            for (int j = 0; j < 8; ++j)
              hh_only[((((size_t)kc * taps + t) * 2 + hh) * cout + n) * 8 + j] = h[at(kc, 0, t, hh, n, j)];
    SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&out->w_hi), hh_only.size() * sizeof(uint16_t)));
    SS_CUDA_CHECK(cudaMemcpy(out->w_hi, hh_only.data(), hh_only.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    st->bytes += hh_only.size() * sizeof(uint16_t);
  }
  out->n_chunks = n_chunks;
  out->taps = taps;
  out->n = cout;
  out->parts = parts;
  out->dual = dual;
  return SS_OK;
}

constexpr size_t kSmemTail = (2 * kMaxStages + 6) * 8 + 5 * 128 * 4 + 16;   // barriers + bias and scalar-residual weights of both phases + TMEM slot
// Launches with row-merged taps stage 12 tap matrices per chunk: four ring slots of conv9_1's first convolution
// (57,600 bytes each) fit only with the device's whole opt-in shared memory and a tail sized for the launch's N
// (three slots measured +2 % on that launch, which cancels what the merged taps save).
constexpr size_t kSmemBudgetMax = 227 * 1024;
constexpr size_t smem_tail_n(int n) { return (2 * kMaxStages + 6) * 8 + 5 * (size_t)n * 4 + 16; }

constexpr int kDefaultRing = 8;    // images of the intermediate tensor kept by a fused ResBlock launch (0: whole batch)
constexpr int kDefaultLag = 160;   // units by which conv2 trails conv1 in a fused ResBlock launch (> one round of 148 CTAs)
constexpr int kMinUnitsForPairs = 4 * kNumSMs;
constexpr int kDefaultPairPolicy = 0;   // measured with the two-issuer kernel (batch 256): G = 1 everywhere is 2 % faster   // use two-group units only while >= 4 waves of them remain

template <int N, Prec P, bool Dual, int G>
int launch_conv_npg(TcJob job, int B, cudaStream_t st) {
  TcConv& p = job.c[0];
  constexpr int MT = TilesPerUnit<N, Dual>::value;
  // Row-aligned units with the MaxPool folded into the epilogue (TcConv::rows): decided by the caller, which must know
  // whether the pooled tensor was written; here only the geometry is checked.
  constexpr bool kCanRows = PrecTraits<P>::split && Dual && G == 1 && (N == 32 || N == 64);
  const bool rows = p.rows != 0;
  const int tpr = p.W / 128;                     // tiles per image row of a row-aligned unit
  SS_REQUIRE(!rows || (kCanRows && job.n_phase == 1 && (tpr == 1 || tpr == 2) && p.W == tpr * 128 &&
                       MT % (2 * tpr) == 0 && p.H % (MT / tpr) == 0),
             SS_E_ARG, "row-aligned conv launch: unsupported geometry (N %d, %d x %d)", N, p.H, p.W);
  const int unit_rows = rows ? MT / tpr : 0;
  SS_REQUIRE(!p.pool_out || (rows && unit_rows == 2 && p.pool_lo && !p.upsample && !p.head_w && p.out && p.out_lo),
             SS_E_ARG, "folded MaxPool needs a row-aligned launch on row pairs that stores its activations");
  bool any_dup = false;
  for (int i = 0; i < p.n_src; ++i) any_dup |= p.src[i].taps == 12;
  SS_REQUIRE(!any_dup || rows, SS_E_ARG, "row-merged taps need a row-aligned launch");
  const size_t rows_extra = rows ? (size_t)(2 * unit_rows - 2) * 32 : 0;   // border positions between the unit's rows (kRowsExtra)
  // Resident weights (TcConv::wres): row-aligned dual launches whose packed weights fit in shared memory next to a
  // ring of at least four activation-only slots.  SS_TC_WRES=0: weights travel with every chunk (A/B runs).
  size_t wres = 0;
  const size_t a_only = ((size_t)G * MT * 128 + 2 * (size_t)(p.W + 3)) * 32 + rows_extra;
  if (rows && kCanRows) {
    const char* wr = getenv("SS_TC_WRES");      // read per call: A/B runs compare the two in one process
    const int want_res = wr ? atoi(wr) : 1;
    size_t total = 0;
    bool paired = true;
    for (int i = 0; i < p.n_src; ++i) {
      const TcSource& src = p.src[i];
      if (src.kind == 1) total += (size_t)src.n_chunks * src.taps * N * 64;
      else paired &= src.kind == 2 && i > 0 && p.src[i - 1].kind == 1 && p.src[i - 1].n_chunks == src.n_chunks &&
                     p.src[i - 1].taps == src.taps;
    }
    // Measured launch by launch (profiles/r2_tuning_experiments.txt, section 12): -8 % on conv1_1.c2, -4 % / -9 % on conv2_1's
    // two launches; +5 % together with row-merged taps (conv9_1.c1) and +6 .. +13 % where a 1x1 source has more
    // chunks than an activation-only slot holds (conv8 / conv9_1's second launches: their residual branches need more
    // stages then, and every stage costs a fixed hand-over).  Hence: no merged taps, and every 1x1 source in one slot.
    int most1x1 = 0;
    for (int i = 0; i < p.n_src; ++i)
      if (p.src[i].taps == 1 && p.src[i].n_chunks > most1x1) most1x1 = p.src[i].n_chunks;
    const size_t chunk1_a = (size_t)G * MT * 128 * 32 + rows_extra;
    const bool suits = !any_dup && (size_t)most1x1 * chunk1_a <= a_only;
    if ((want_res == 2 || (want_res && suits)) && paired && total > 0 &&
        total + 4 * a_only + smem_tail_n(N) <= kSmemBudgetMax)
      wres = total;
  }
  p.wres = (int)wres;
  if (wres) {
    int off = 0;
    for (int i = 0; i < p.n_src; ++i) {
      TcSource& src = p.src[i];
      if (src.kind == 1) {
        src.wres_off = off;
        src.wres_stride = src.taps * N * 64;
        src.wres_bytes = src.n_chunks * src.wres_stride;
        off += src.wres_bytes;
      } else {
        src.wres_off = p.src[i - 1].wres_off;
        src.wres_stride = p.src[i - 1].wres_stride;
        src.wres_bytes = 0;
      }
      src.w_rows = 2 * N;
    }
  }
  const size_t sb = wres ? a_only : stage_bytes(N, p.W, G * MT, Dual, any_dup ? 12 : 9) + rows_extra;
  const bool big = any_dup || wres;
  const size_t budget = (big ? kSmemBudgetMax : kSmemBudget) - wres, tail = big ? smem_tail_n(N) : kSmemTail;
  // Ring slot size.  Every stage costs a fixed hand-over (full/empty barrier round trip and an MMA-issue bubble,
  // ~500-1000 cycles measured with the tuning hooks), which a 3x3 stage hides behind its 9 x MT MMAs and a 1x1 stage
  // (MT MMAs per chunk) does not: 1x1 sources therefore pack several K-chunks into one slot, and when a launch has
  // many of them (the residual branch of a decoder block reads 64-256 channels) the slot is enlarged — down to three
  // slots — so that up to `cap` chunks fit.  SS_TC_CPS caps the chunks per stage (1 = one chunk per stage everywhere).
  const char* ce = getenv("SS_TC_CPS");
  const int cap = ce ? atoi(ce) : 4;
  const size_t chunk1 = (size_t)G * MT * 128 * 32 + rows_extra + (wres ? 0 : (Dual ? 2 : 1) * (size_t)N * 32);
  int most1 = 0;
  for (int ph = 0; ph < job.n_phase; ++ph)
    for (int i = 0; i < job.c[ph].n_src; ++i)
      if (job.c[ph].src[i].taps == 1 && job.c[ph].src[i].n_chunks > most1) most1 = job.c[ph].src[i].n_chunks;
  size_t stride = sb;
  int cps = (int)(sb / chunk1);
  for (int want = (most1 < cap ? most1 : cap); want > cps; --want) {
    const size_t need = (size_t)want * chunk1;
    if ((budget - tail) / need >= 3) { stride = need > sb ? need : sb; cps = want; break; }
  }
  if (cps > cap) cps = cap;
  p.cps = cps < 1 ? 1 : cps;
  stride = (stride + 127) & ~(size_t)127;
  p.stage_stride = (int)stride;
  int stages = (int)((budget - tail) / stride);
  if (stages > kMaxStages) stages = kMaxStages;
  { const char* ms = getenv("SS_TC_MAX_STAGES"); if (ms && atoi(ms) >= 2 && stages > atoi(ms)) stages = atoi(ms); }   // tuning
  SS_REQUIRE(stages >= 2, SS_E_ARG, "conv stage of %zu bytes does not fit twice in shared memory", stride);
  p.stages = stages;
  static const int debug = [] { const char* e = getenv("SS_TC_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug = debug;
  const size_t smem = wres + (size_t)stages * stride + tail;
  constexpr bool kCanSub = PrecTraits<P>::split && G == 1;
  const int positions = p.H * (p.W + 2) - 2;          // (1,1) .. (H,W) in flattened padded coordinates
  p.units_per_image = rows ? p.H / unit_rows : (positions + G * MT * 128 - 1) / (G * MT * 128);
  p.total_units = p.units_per_image * B;
  // Packed units (TcConv::packed) for the images of 32 x 64 and below (plane-major tensors), whose last unit is partly
  // or mostly air: the batch as one stream of positions.  SS_TC_PACK=0: units per image (A/B runs).
  p.packed = 0;
  p.batch = B;
  {
    const char* pk = getenv("SS_TC_PACK");
    bool contiguous = true;      // every source plane-major: the stream of padded images is contiguous in each plane
    for (int i = 0; i < p.n_src; ++i) contiguous &= p.src[i].img_stride == (int64_t)(p.H + 2) * (p.W + 2) && !p.src[i].in_up;
    if (!rows && contiguous && job.n_phase == 1 && p.H * p.W <= 2048 && !p.head_w && !p.res_x &&
        (pk == nullptr || atoi(pk) != 0)) {
      const int64_t stream = (int64_t)(B - 1) * (p.H + 2) * (p.W + 2) + positions;
      p.packed = 1;
      p.total_units = (int)((stream + G * MT * 128 - 1) / (G * MT * 128));
      p.units_per_image = p.total_units;      // every unit is a unit of "image 0": positions run on across images
    }
  }
  for (int ph = 0; ph < job.n_phase; ++ph)
    SS_REQUIRE(job.c[ph].relu == 1, SS_E_ARG, "conv_tc_kernel applies ReLU unconditionally");
  // Split-K sub-accumulation (TcConv::n_sub): the K-chunks of the 3x3 sources are cut into groups so that an MMA
  // chain covers 9 taps x a few chunks (+ its share of the 1x1 residual chunks) instead of 9 C_in / 16.  Every group
  // costs a TMEM buffer turn (drain by the epilogue warps, barrier round trip).  Measured (tools/sub_sweep.sh,
  // profiles/r2_sub_accumulation.txt): the logit error falls from 2.6e-5 to 5.5e-6 of float64 truth — the reference's
  // own float32 sits at 4.4e-6 — with one chunk per group in the layers at 32 x 64 and below alone (their chains
  // are the long ones: up to 144 x 3 MMAs), for +8 % of classifier time; groups in the layers at 64 x 128 and above
  // (chains of 18-72) buy nothing measurable for another +12 %.  Hence: SS_TC_SUB_DEEP groups (default: every chunk)
  // in the deep layers, SS_TC_SUB (default 1: the plain kernel) in the big ones.
  static const int sub_big = [] { const char* e = getenv("SS_TC_SUB"); return e ? atoi(e) : 1; }();
  static const int sub_deep = [] { const char* e = getenv("SS_TC_SUB_DEEP"); return e ? atoi(e) : 16; }();
  for (int ph = 0; ph < job.n_phase; ++ph) {
    int most = 1;
    for (int i = 0; i < job.c[ph].n_src; ++i)
      if (job.c[ph].src[i].taps >= 9 && job.c[ph].src[i].n_chunks > most) most = job.c[ph].src[i].n_chunks;
    int cap = (p.H >= 64) ? sub_big : sub_deep;
    if (cap < 1) cap = 1;
    job.c[ph].n_sub = kCanSub ? (most < cap ? most : cap) : 1;
  }
  bool any_sub = false;
  for (int ph = 0; ph < job.n_phase; ++ph) any_sub |= job.c[ph].n_sub > 1;
  // Stage program of a unit (TcJob::prog): accumulation group by accumulation group, source by source, the chunks of
  // the group (a source's n chunks are dealt out n / n_sub per group, one more to the first n % n_sub), `cps` chunks
  // of a 1x1 source per stage.  Every group holds at least one chunk of the widest 3x3 source (n_sub <= its chunks);
  // in the dual layout sources come in (kind 1, kind 2) pairs with equal chunk counts, so a group always starts with
  // a kind-1 stage, whose first MMA initialises both column groups.
  for (int ph = 0; ph < job.n_phase; ++ph) {
    const TcConv& c = job.c[ph];
    int len = 0;
    for (int sub = 0; sub < c.n_sub; ++sub) {
      const int first_of_group = len;
      for (int si = 0; si < c.n_src; ++si) {
        const TcSource& src = c.src[si];
        const int q = src.n_chunks / c.n_sub, r = src.n_chunks % c.n_sub;
        const int lo = sub * q + (sub < r ? sub : r), hi = lo + q + (sub < r ? 1 : 0);
        const int per_stage = (src.taps == 1) ? p.cps : 1;
        for (int kc = lo; kc < hi; kc += per_stage) {
          const int n = (hi - kc < per_stage) ? (hi - kc) : per_stage;
          SS_REQUIRE(len < kMaxProg && kc < 64 && n < 8 && si < 8, SS_E_ARG, "conv stage program too long (%d stages)", len);
          const bool half = src.in_up != nullptr && kc + n > src.up_chunk0;
          const int nfull = half ? (src.up_chunk0 > kc ? src.up_chunk0 - kc : 0) : 0;
          SS_REQUIRE(!half || (rows && (src.taps == 12 || src.taps == 1)), SS_E_ARG,
                     "half-row sources belong to row-aligned launches (merged taps or 1x1)");
          job.prog[ph][len] = prog_entry(si, kc, n, len == first_of_group, false, src.taps >= 9, src.kind, src.taps == 12,
                                         half, nfull);
          if (ph == 0 && wres)
            job.prog_b[len] = (uint32_t)((src.wres_off + kc * src.wres_stride) >> 4) | ((uint32_t)(src.wres_stride >> 4) << 16);
          ++len;
        }
      }
      SS_REQUIRE(len > first_of_group, SS_E_ARG, "empty accumulation group %d of %d", sub, c.n_sub);
      SS_REQUIRE(!Dual || ((job.prog[ph][first_of_group] >> 15) & 3u) == 1u, SS_E_ARG,
                 "dual layout: accumulation group %d does not start with a dual source", sub);
      job.prog[ph][len - 1] |= 1u << 13;
    }
    job.prog_len[ph] = len;
  }
  const int items = p.total_units * job.n_phase;
  const int grid = items < kNumSMs ? items : kNumSMs;
  if (getenv("SS_TC_VERBOSE"))
    fprintf(stderr, "conv launch N=%d %dx%d rows=%d dup=%d wres=%zu stride=%zu stages=%d cps=%d prog=%d units/img=%d smem=%zu\n", N,
            p.H, p.W, (int)rows, (int)any_dup, wres, stride, stages, p.cps, job.prog_len[0], p.units_per_image, smem);
  if (job.n_phase == 2) {
    SS_REQUIRE(p.W + 3 <= G * MT * 128, SS_E_ARG, "fused ResBlock launch: halo %d exceeds the unit", p.W + 3);
    SS_REQUIRE(p.total_units <= job.flags_cap, SS_E_ARG, "fused ResBlock launch: %d units exceed the flag array",
               p.total_units);
    SS_CUDA_CHECK(cudaMemsetAsync(job.flags, 0, (size_t)p.total_units * sizeof(int), st));
    // ring mode for the intermediate tensor (TcJob::ring): only when the ring is shorter than the batch and the
    // units it makes a c[0] unit wait for precede that unit in the item order
    const int lag_eff = job.lag < p.total_units ? job.lag : p.total_units;
    int ring = job.ring_request > 0 ? job.ring_request : 0;
    if (ring > 0) {
      // the units a c[0] unit waits for must lie >= 2 rounds of the grid (296 items) before it in the item order, or
      // the wait stalls for real: 2 (ring * units_per_image - lag) - 3 >= 2 * 148  (measured: a thin margin costs 35 %)
      const int need = (lag_eff + 2 * kNumSMs + 8 + p.units_per_image - 1) / p.units_per_image;
      if (ring < need) ring = need;
      if (ring >= B) ring = 0;
    }
    job.ring = ring;
    job.c[0].out_ring = ring;
    for (int i = 0; i < job.c[1].n_src; ++i)
      if (job.c[1].src[i].ring < 0) job.c[1].src[i].ring = ring;      // marked by the host: reads the intermediate
    if (ring > 0) SS_CUDA_CHECK(cudaMemsetAsync(job.flags2, 0, (size_t)p.total_units * sizeof(int), st));
  } else {
    for (int i = 0; i < job.c[0].n_src; ++i)
      if (job.c[0].src[i].ring < 0) job.c[0].src[i].ring = 0;
  }
  // Which epilogue the launch needs picks the instantiation (conv_tc_kernel's Epi): 1 the folded MaxPool, 2 the general
  // one with the fused mask-head partials, 0 the general one alone.
  bool any_head = false, any_resx = false;
  for (int ph = 0; ph < job.n_phase; ++ph) {
    any_head |= job.c[ph].head_w != nullptr;
    any_resx |= job.c[ph].res_x != nullptr;
  }
  SS_REQUIRE(!(any_head || any_resx) || N == 32, SS_E_ARG, "fused mask head / scalar residual: 32-channel launches only");
  SS_REQUIRE(!(any_head && any_resx), SS_E_ARG, "no launch has both the fused mask head and the scalar residual");
  bool any_up = false;
  for (int ph = 0; ph < job.n_phase; ++ph) any_up |= job.c[ph].upsample != 0;
  auto launch_k = [&](auto* kernel, bool& configured_k, bool rows_k) -> int {
    if (!configured_k) {
      SS_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(rows_k ? kSmemBudgetMax : kSmemBudget)));
      configured_k = true;
    }
    kernel<<<grid, kTcThreads, smem, st>>>(job);
    return SS_OK;
  };
  auto launch = [&](auto sub_c, auto rows_c, auto epi_c) -> int {
    constexpr bool kS = decltype(sub_c)::value;
    constexpr int kR = decltype(rows_c)::value, kE = decltype(epi_c)::value;
    static bool configured_up = false, configured_noup = false;      // per instantiation of this call operator
    if constexpr (kE != 1) {      // (the folded-pool epilogue has no such stores either way)
      if (!any_up) return launch_k(conv_tc_kernel<N, P, Dual, G, kS, kR, kE, true>, configured_noup, kR != 0);
    }
    return launch_k(conv_tc_kernel<N, P, Dual, G, kS, kR, kE, false>, configured_up, kR != 0);
  };
  using std::integral_constant;
  using False = integral_constant<bool, false>;
  using True = integral_constant<bool, true>;
  int rc = SS_OK;
  if (rows) {
    SS_REQUIRE(!any_sub, SS_E_ARG, "row-aligned conv launch: sub-accumulation is not available in this geometry");
    if constexpr (kCanRows) {
      // (tpr = tiles per image row; the folded pool needs units of two rows: MT / tpr == 2)
      if (tpr == 1) {
        if constexpr (MT == 2) {
          if (p.pool_out) rc = launch(False{}, integral_constant<int, 1>{}, integral_constant<int, 1>{});
          else rc = launch(False{}, integral_constant<int, 1>{}, integral_constant<int, 0>{});
        } else if constexpr (N == 32) {
          SS_REQUIRE(!p.pool_out, SS_E_ARG, "folded pool: the unit is not a pair of rows");
          SS_REQUIRE(!any_resx, SS_E_ARG, "scalar residual: no such epilogue in this geometry");
          if (any_head) rc = launch(False{}, integral_constant<int, 1>{}, integral_constant<int, 2>{});
          else rc = launch(False{}, integral_constant<int, 1>{}, integral_constant<int, 0>{});
        } else {
          SS_REQUIRE(!p.pool_out, SS_E_ARG, "folded pool: the unit is not a pair of rows");
          rc = launch(False{}, integral_constant<int, 1>{}, integral_constant<int, 0>{});
        }
      } else if constexpr (MT % 4 == 0) {
        if constexpr (MT == 4) {
          if (p.pool_out) rc = launch(False{}, integral_constant<int, 2>{}, integral_constant<int, 1>{});
          else if (N == 32 && (any_head || any_resx)) {
            if constexpr (N == 32) {
              if (any_head) rc = launch(False{}, integral_constant<int, 2>{}, integral_constant<int, 2>{});
              else rc = launch(False{}, integral_constant<int, 2>{}, integral_constant<int, 3>{});
            }
          } else rc = launch(False{}, integral_constant<int, 2>{}, integral_constant<int, 0>{});
        } else {
          SS_REQUIRE(!p.pool_out && !any_head, SS_E_ARG, "row-aligned launch: no such epilogue in this geometry");
          rc = launch(False{}, integral_constant<int, 2>{}, integral_constant<int, 0>{});
        }
      }
    }
  } else {
    SS_REQUIRE(!p.pool_out, SS_E_ARG, "folded pool without row-aligned units");
    if constexpr (N == 32) {
      auto by_epi = [&](auto sub_c) -> int {
        if (any_head) return launch(sub_c, integral_constant<int, 0>{}, integral_constant<int, 2>{});
        if (any_resx) return launch(sub_c, integral_constant<int, 0>{}, integral_constant<int, 3>{});
        return launch(sub_c, integral_constant<int, 0>{}, integral_constant<int, 0>{});
      };
      if constexpr (kCanSub) rc = any_sub ? by_epi(True{}) : by_epi(False{});
      else rc = by_epi(False{});
    } else if constexpr (kCanSub) {
      rc = any_sub ? launch(True{}, integral_constant<int, 0>{}, integral_constant<int, 0>{})
                   : launch(False{}, integral_constant<int, 0>{}, integral_constant<int, 0>{});
    } else {
      rc = launch(False{}, integral_constant<int, 0>{}, integral_constant<int, 0>{});
    }
  }
  if (rc) return rc;
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

template <int N, Prec P, bool Dual>
int launch_conv_np(const TcJob& job, int B, cudaStream_t st) {
  const TcConv& p = job.c[0];
  constexpr int MT = TilesPerUnit<N, Dual>::value;
  const int positions = p.H * (p.W + 2) - 2;
  const int64_t pair_units = (int64_t)((positions + 2 * MT * 128 - 1) / (2 * MT * 128)) * B;
  const bool fits = stage_bytes(N, p.W, 2 * MT, Dual) * 2 + kSmemTail <= kSmemBudget;
  // Policy bits (SS_TC_PAIRS, tuning): 1 dual layout, 2 other split-precision layers, 4 single-pass C_out >= 96,
  // 8 single-pass C_out <= 64.  Default: see kDefaultPairPolicy.
  static const int policy = [] { const char* e = getenv("SS_TC_PAIRS"); return e ? atoi(e) : kDefaultPairPolicy; }();
  const int cls = Dual ? 1 : (PrecTraits<P>::split ? 2 : (N >= 96 ? 4 : 8));
  const bool want = (policy & cls) != 0;
  if (want && pair_units >= kMinUnitsForPairs && fits) return launch_conv_npg<N, P, Dual, 2>(job, B, st);
  return launch_conv_npg<N, P, Dual, 1>(job, B, st);
}

template <Prec P>
int launch_conv_p(const TcJob& p, int N, int B, cudaStream_t st) {
  constexpr bool kDualSmall = PrecTraits<P>::split;      // == is_dual(P, N) for N <= 64
  switch (N) {
    case 32: return launch_conv_np<32, P, kDualSmall>(p, B, st);
    case 64: return launch_conv_np<64, P, kDualSmall>(p, B, st);
    case 96: return launch_conv_np<96, P, false>(p, B, st);
    case 128: return launch_conv_np<128, P, false>(p, B, st);
  }
  set_error("unsupported C_out %d", N);
  return SS_E_ARG;
}

int launch_conv(Prec prec, const TcJob& p, int N, int B, cudaStream_t st) {
  switch (prec) {
    case Prec::Bf16: return launch_conv_p<Prec::Bf16>(p, N, B, st);
    case Prec::F16: return launch_conv_p<Prec::F16>(p, N, B, st);
    case Prec::F16x3: return launch_conv_p<Prec::F16x3>(p, N, B, st);
  }
  return SS_E_ARG;
}

// Sources of one convolution of tensor `x` (planes from plane0) with packed weights `w`.
//  * single precision: one source;
//  * split, dual layout (C_out <= 64): x_hi . [w_hi | w_lo] as one 2N-column source, then x_lo . w_hi into the
//    correction columns; the first source of a launch must be a dual one (its first MMA zeroes both groups);
//  * split, C_out >= 96: three sources — x_lo . w_hi, x_hi . w_lo (the corrections) and x_hi . w_hi (the main
//    term) — and a launch issues every correction before any main term: the tensor core truncates the fp32
//    accumulator after every MMA, so the small terms are summed while the accumulator (and its ulp) is still
//    small (Ootomo & Yokota 2022 observe the same for mma.sync; tools/precision_study.py measures it here).
enum class Terms { All, Corrections, Main };

// up / up_chunk0 (dual layout only): the K-chunks from up_chunk0 on are read from the half-row tensor `up`
// (TcSource::in_up) instead of from the planes of x that follow.
void add_sources(TcConv* p, const TcState* s, const Tensor& x, int plane0, const PackedConv& w, Terms terms,
                 int ring = 0, const Tensor* up = nullptr, int up_chunk0 = 0) {
  const int part_elems = w.taps * w.n * 16;
  const int chunk_elems = w.parts * part_elems;
  const int first = p->n_src;
  auto set_strides = [&]() {
    for (int i = first; i < p->n_src; ++i) {
      p->src[i].img_stride = x.img_stride();
      p->src[i].plane_stride = x.plane_stride();
    }
  };
  if (!is_split(s->prec)) {
    if (terms != Terms::Corrections)
      p->src[p->n_src++] = TcSource{x.data, x.planes, plane0, w.n_chunks, w.taps, 0, ring, chunk_elems, w.w};
    set_strides();
    return;
  }
  if (w.dual) {
    if (terms != Terms::Corrections) {
      TcSource hi{x.data, x.planes, plane0, w.n_chunks, w.taps, 1, ring, chunk_elems, w.w};
      TcSource lo{x.lo, x.planes, plane0, w.n_chunks, w.taps, 2, ring, part_elems, w.w_hi};
      if (up) {
        hi.in_up = up->data; lo.in_up = up->lo;
        hi.up_chunk0 = lo.up_chunk0 = up_chunk0;
        hi.up_planes_total = lo.up_planes_total = up->planes;
      }
      p->src[p->n_src++] = hi;
      p->src[p->n_src++] = lo;
    }
    set_strides();
    return;
  }
  if (terms != Terms::Main) {
    p->src[p->n_src++] = TcSource{x.lo, x.planes, plane0, w.n_chunks, w.taps, 0, ring, chunk_elems, w.w};
    p->src[p->n_src++] = TcSource{x.data, x.planes, plane0, w.n_chunks, w.taps, 0, ring, chunk_elems, w.w + part_elems};
  }
  if (terms != Terms::Corrections)
    p->src[p->n_src++] = TcSource{x.data, x.planes, plane0, w.n_chunks, w.taps, 0, ring, chunk_elems, w.w};
  set_strides();
}

// One ResBlock: t = relu(conv3x3(x) + b1);  out = relu(conv3x3(t) + conv1x1(x) + b2 + b_res).
int tc_res_block(TcState* s, int which, const Tensor& x, int x_plane0, Tensor& out, int out_plane0, int upsample,
                 int B, cudaStream_t st, const float* head_w = nullptr, float* head_out = nullptr,
                 bool store_out = true, const float* res_x = nullptr, const float* res_w = nullptr,
                 bool c1_done = false, Tensor* pool = nullptr, bool* pooled = nullptr, const Tensor* x_up = nullptr) {
  // x_up: the up-sampled half of the block input lives in a half-row tensor (the caller made the block below write it
  // with upsample == 2, and has checked that this block's launches are row-aligned with merged taps); x then
  // supplies the skip half only.  upsample == 2: this block's output goes to the half-row tensor `out`.
  const TcBlock& rb = s->rb[which];
  Tensor& t = s->t[which];
  const int N = rb.c1.n;
  TcConv p{};
  p.H = x.H; p.W = x.W;
  p.bias = rb.bias1;
  p.inv_scale = rb.inv_scale1;
  p.relu = 1;
  p.out = t.data; p.out_lo = t.lo; p.out_planes_total = t.planes; p.out_plane0 = 0; p.upsample = 0;
  p.o_img_stride = t.img_stride(); p.o_plane_stride = t.plane_stride();
  p.err = s->err;
  TcConv q{};
  add_sources(&q, s, t, 0, rb.c2, Terms::Corrections, -1);     // -1: "the intermediate tensor", ring resolved at launch
  if (res_x == nullptr) {
    add_sources(&q, s, x, x_plane0, rb.res, Terms::Corrections);
    add_sources(&q, s, x, x_plane0, rb.res, Terms::Main, 0, x_up, x_up ? rb.skip_cin / 16 : 0);
  } else {            // single-channel block input: the residual branch is an FMA in the epilogue (TcConv::res_x)
    q.res_x = res_x;
    q.res_w = res_w;
  }
  add_sources(&q, s, t, 0, rb.c2, Terms::Main, -1);
  q.H = x.H; q.W = x.W;
  q.bias = rb.bias2;
  q.inv_scale = rb.inv_scale2;
  q.relu = 1;
  q.out = out.data; q.out_lo = out.lo; q.out_planes_total = out.planes; q.out_plane0 = out_plane0; q.upsample = upsample;
  q.o_img_stride = out.img_stride(); q.o_plane_stride = out.plane_stride();
  const char* pe = getenv("SS_TC_PAIR_STORE");
  q.pair_store = pe ? atoi(pe) : 1;
  q.head_w = head_w; q.head_out = head_out;
  if (head_w && !store_out) { q.out = nullptr; q.out_lo = nullptr; }
  q.err = s->err;
  // SS_TC_FUSE=1: both convolutions of the block in ONE persistent launch, conv2 trailing conv1 by SS_TC_LAG units
  // behind per-unit completion flags (TcJob).  Bit-identical results, t is read back from L2 instead of HBM and the
  // launch has one tail; measured equal in time to two launches (29.77 vs 29.74 ms per 10-min clip, f16x3, batch
  // 256: the c2 launches are bound by MMA issue and stores, not by DRAM reads), so the simpler schedule is the
  // default.  Read per call so that tests can compare the two in one process.
  const char* fe = getenv("SS_TC_FUSE");
  const char* le = getenv("SS_TC_LAG");
  const int fuse = fe ? atoi(fe) : 0;
  const int lag = le ? atoi(le) : kDefaultLag;
  const char* re = getenv("SS_TC_RING");
  TcJob job{};
  job.flags = s->flags;
  job.flags2 = s->flags2;
  job.flags_cap = s->flags_cap;
  job.lag = lag;
  const char* ly = getenv("SS_TC_LAYOUT");
  job.layout = ly ? atoi(ly) : 1;
  const char* hf = getenv("SS_TC_FENCE");
  job.heavy_fence = hf ? atoi(hf) : 0;
  const char* ep = getenv("SS_TC_EPI");
  job.epi = ep ? atoi(ep) : 3;
  job.ring_request = fuse ? (re ? atoi(re) : kDefaultRing) : 0;
  // MaxPool2d(2) of the block output folded into conv2's epilogue (row-aligned units, TcConv::rows): split precision,
  // dual layout, image width = the tiles of a unit's row (conv1_1: 256 = 2 x 128 at N = 32; conv2_1: 128 at N = 64),
  // separate launches, plain accumulation chain, one group per unit.  SS_TC_POOL_FOLD=0 keeps pool_planar (A/B runs).
  if (pooled) *pooled = false;
  // Row-aligned units (TcConv::rows) for every launch whose geometry allows them: dual layout, image width = the
  // tiles of a unit's row (256 = 2 x 128 at N = 32: conv1_1, conv9_1; 128 at N = 64: conv2_1, conv8), separate
  // launches, plain accumulation chain, one group per unit.  No border position is computed (64 / 32 units per image
  // instead of 65.5 / 33.5) and the 2 x 2 windows of a MaxPool lie inside a unit.  SS_TC_ROWS=0: flat units (A/B runs).
  bool rows_ok = false;
  int unit_rows = 0;
  if (is_dual(s->prec, N) && !fuse) {
    const char* rw = getenv("SS_TC_ROWS");
    const char* sb = getenv("SS_TC_SUB");
    const char* pp = getenv("SS_TC_PAIRS");
    const int mt = (N == 32) ? 4 : 2;                     // tiles per unit of the dual layout
    const int tpr = x.W / 128;                            // tiles per image row
    const bool plain = (sb == nullptr || atoi(sb) <= 1) && (pp == nullptr || (atoi(pp) & 1) == 0);
    rows_ok = plain && (tpr == 1 || tpr == 2) && x.W == tpr * 128 && mt % (2 * tpr) == 0 && x.H % (mt / tpr) == 0;
    unit_rows = rows_ok ? mt / tpr : 0;
    if (rows_ok && (rw == nullptr || atoi(rw) != 0)) { p.rows = 1; q.rows = 1; }
  }
  // conv1 of a decoder block on a row-aligned launch: the up-sampled half of its input as a "rowdup" source (six tap
  // MMAs per tile instead of nine; SS_TC_TAPMERGE=0 keeps the nine)
  {
    const char* tm = getenv("SS_TC_TAPMERGE");
    if (x_up) {
      SS_REQUIRE(p.rows && rb.skip_cin > 0 && (tm == nullptr || atoi(tm) != 0), SS_E_ARG,
                 "half-row block input needs a row-aligned launch with merged taps");
      add_sources(&p, s, x, x_plane0, rb.c1_skip, Terms::Main);
      add_sources(&p, s, *x_up, 0, rb.c1_up, Terms::Main, 0, x_up, 0);
    } else if (p.rows && rb.skip_cin > 0 && (tm == nullptr || atoi(tm) != 0)) {
      add_sources(&p, s, x, x_plane0, rb.c1_skip, Terms::Main);
      add_sources(&p, s, x, x_plane0 + rb.skip_cin / 8, rb.c1_up, Terms::Main);
    } else {
      add_sources(&p, s, x, x_plane0, rb.c1, Terms::Corrections);
      add_sources(&p, s, x, x_plane0, rb.c1, Terms::Main);
    }
  }
  if (pool && rows_ok && unit_rows == 2 && !upsample && !head_w) {
    const char* pf = getenv("SS_TC_POOL_FOLD");
    const int fold_mask = pf ? atoi(pf) : 3;              // bit 0: conv1_1 (N = 32), bit 1: conv2_1 (N = 64)
    if ((fold_mask & (N == 32 ? 1 : 2)) != 0 && pool->planes == N / 8 && pool->H == x.H / 2 && pool->W == x.W / 2) {
      q.rows = 1;
      q.pool_out = pool->data;
      q.pool_lo = pool->lo;
      q.pool_img_stride = pool->img_stride();
      q.pool_plane_stride = pool->plane_stride();
      if (pooled) *pooled = true;
    }
  }
  if (c1_done) {            // the intermediate tensor was produced by another kernel (conv1_direct): conv2 only
    q.prof = (s->launch_index++ == s->prof_layer) ? s->prof : nullptr;
    job.c[0] = q; job.n_phase = 1;
    return launch_conv(s->prec, job, N, B, st);
  }
  if (fuse) {
    p.prof = (s->launch_index++ == s->prof_layer) ? s->prof : nullptr;
    job.c[0] = p; job.c[1] = q; job.n_phase = 2;
    return launch_conv(s->prec, job, N, B, st);
  }
  p.prof = (s->launch_index++ == s->prof_layer) ? s->prof : nullptr;
  job.c[0] = p; job.n_phase = 1;
  int rc = launch_conv(s->prec, job, N, B, st);
  if (rc) return rc;
  q.prof = (s->launch_index++ == s->prof_layer) ? s->prof : nullptr;
  job.c[0] = q;
  return launch_conv(s->prec, job, N, B, st);     // (the -1 ring marks of q's sources resolve to 0 in the launcher)
}

template <Prec P>
int tc_pool_p(const Tensor& in, int plane0, int planes, Tensor& out, int B, cudaStream_t st) {
  const int64_t total = (int64_t)B * planes * (in.H / 2) * (in.W / 2);
  pool_planar<P><<<(int)((total + 255) / 256), 256, 0, st>>>(in.data, in.lo, in.img_stride(), in.plane_stride(), plane0,
                                                             planes, in.H, in.W, out.data, out.lo, out.img_stride(),
                                                             out.plane_stride(), total);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

// Built lazily on the first call of a mode, so that nobody pays for a workspace they do not use.
void tc_free_state(ss_ctx* ctx, TcState* s);

int tc_build_state(ss_ctx* ctx, Prec prec, TcState* s);

// The slot is published only when the whole state exists: a failed build (out of memory is realistic: 18-36 GB of
// workspace at max_batch 1005) frees what it got and leaves the slot empty, so the next call retries the build
// instead of launching kernels on a partial state.
int tc_build(ss_ctx* ctx, Prec prec, TcState** out) {
  const int slot = (int)prec;
  if (ctx->tc[slot]) { *out = static_cast<TcState*>(ctx->tc[slot]); return SS_OK; }
  TcState* s = new TcState();
  const int rc = tc_build_state(ctx, prec, s);
  if (rc) {
    cudaGetLastError();
    tc_free_state(ctx, s);
    return rc;
  }
  ctx->tc[slot] = s;
  ctx->device_bytes += s->bytes;
  *out = s;
  return SS_OK;
}

int tc_build_state(ss_ctx* ctx, Prec prec, TcState* s) {
  s->prec = prec;
  s->max_batch = ctx->max_batch;
  const int B = ctx->max_batch;
  int rc;
  // weights: read the folded f32 tensors back from the device blob (already validated at ss_ctx_create)
  for (int i = 0; i < RB_COUNT; ++i) {
    const ResBlockW& rb = ctx->rb[i];
    auto fetch = [&](const ConvW& c, std::vector<float>* w, std::vector<float>* bias) -> int {
      w->resize((size_t)c.taps * c.cin * c.cout);
      bias->resize(c.cout);
      SS_CUDA_CHECK(cudaMemcpy(w->data(), c.w, w->size() * 4, cudaMemcpyDeviceToHost));
      SS_CUDA_CHECK(cudaMemcpy(bias->data(), c.b, bias->size() * 4, cudaMemcpyDeviceToHost));
      return SS_OK;
    };
    std::vector<float> w1, w2, wr, b1, b2, br;
    if ((rc = fetch(rb.c1, &w1, &b1))) return rc;
    if ((rc = fetch(rb.c2, &w2, &b2))) return rc;
    if ((rc = fetch(rb.res, &wr, &br))) return rc;
    const float s1 = weight_scale(prec, {&w1});
    const float s2 = weight_scale(prec, {&w2, &wr});       // conv2 and the residual share one accumulator
    float s1_used = s1;
    if (i == RB_CONV1) {
      // C_in = 1: the nine taps become input channels 0..8 of a 1x1 convolution over the im2col'd operand
      // tensor written by mel_to_planar; the residual 1x1 reads the centre tap (channel 4).
      const int co = rb.c1.cout;
      std::vector<float> w1k((size_t)16 * co, 0.f), wrk((size_t)16 * co, 0.f);
      for (int t = 0; t < 9; ++t)
        for (int n = 0; n < co; ++n) w1k[(size_t)t * co + n] = w1[(size_t)t * co + n];
      for (int n = 0; n < co; ++n) wrk[(size_t)4 * co + n] = wr[n];
      if ((rc = pack_conv(s, w1k.data(), 1, 16, co, s1, &s->rb[i].c1))) return rc;
      if ((rc = pack_conv(s, wrk.data(), 1, 16, co, s2, &s->rb[i].res))) return rc;
    } else {
      if ((rc = pack_conv(s, w1.data(), 9, rb.c1.cin, rb.c1.cout, s1, &s->rb[i].c1))) return rc;
      if ((rc = pack_conv(s, wr.data(), 1, rb.res.cin, rb.res.cout, s2, &s->rb[i].res))) return rc;
    }
    if ((i == RB_CONV8 || i == RB_CONV9) && is_dual(prec, rb.c1.cout)) {
      // cat([skip, up]): the first half of the input channels is the encoder's skip tensor, the second the 2x
      // up-sampled output of the block below (classify_tc_p: out_plane0 of conv7 / conv8 = half of the planes)
      const int cin = rb.c1.cin, co = rb.c1.cout, cs = cin / 2, cu = cin - cs;
      std::vector<float> ws((size_t)9 * cs * co), wu((size_t)12 * cu * co);
      auto w_at = [&](int t, int c, int n) { return w1[((size_t)t * cin + c) * co + n]; };
      for (int t = 0; t < 9; ++t)
        for (int c = 0; c < cs; ++c)
          for (int n = 0; n < co; ++n) ws[((size_t)t * cs + c) * co + n] = w_at(t, c, n);
      for (int dx = 0; dx < 3; ++dx)
        for (int c = 0; c < cu; ++c)
          for (int n = 0; n < co; ++n) {
            const float wm = w_at(0 + dx, cs + c, n), w0 = w_at(3 + dx, cs + c, n), wp = w_at(6 + dx, cs + c, n);
            wu[((size_t)(0 + dx) * cu + c) * co + n] = wm;            // top rows: input row y - 1
            wu[((size_t)(3 + dx) * cu + c) * co + n] = w0 + wp;       //           rows y, y + 1 (identical)
            wu[((size_t)(6 + dx) * cu + c) * co + n] = wm + w0;       // bottom rows: rows y - 1, y (identical)
            wu[((size_t)(9 + dx) * cu + c) * co + n] = wp;            //              row y + 1
          }
      const float s1m = weight_scale(prec, {&w1, &wu});
      SS_REQUIRE(s1m == s1 || s1m * 2.f == s1, SS_E_BLOB, "row-merged weights of block %d need an unexpected scale", i);
      // one scale per accumulator: the merged matrices may be up to twice the largest single weight
      if (s1m != s1) {
        PackedConv old = s->rb[i].c1;
        if (old.w) cudaFree(old.w);
        if (old.w_hi) cudaFree(old.w_hi);
        s->rb[i].c1 = PackedConv{};
        if ((rc = pack_conv(s, w1.data(), 9, rb.c1.cin, rb.c1.cout, s1m, &s->rb[i].c1))) return rc;
        s1_used = s1m;
      }
      if ((rc = pack_conv(s, ws.data(), 9, cs, co, s1m, &s->rb[i].c1_skip))) return rc;
      if ((rc = pack_conv(s, wu.data(), 12, cu, co, s1m, &s->rb[i].c1_up))) return rc;
      s->rb[i].skip_cin = cs;
    }
    if ((rc = pack_conv(s, w2.data(), 9, rb.c2.cin, rb.c2.cout, s2, &s->rb[i].c2))) return rc;
    s->rb[i].inv_scale1 = 1.f / s1_used;
    s->rb[i].inv_scale2 = 1.f / s2;
    for (size_t k = 0; k < b2.size(); ++k) b2[k] += br[k];
    SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&s->rb[i].bias2), b2.size() * 4));
    SS_CUDA_CHECK(cudaMemcpy(s->rb[i].bias2, b2.data(), b2.size() * 4, cudaMemcpyHostToDevice));
    s->rb[i].bias1 = rb.c1.b;
  }
  SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&s->err), 4 * sizeof(int)));
  SS_CUDA_CHECK(cudaMemset(s->err, 0, 4 * sizeof(int)));
  SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&s->head_part), (size_t)B * kMels * kFrames * 4 * sizeof(float)));
  s->bytes += (size_t)B * kMels * kFrames * 4 * sizeof(float);
  s->flags_cap = B * (((kMels + 2) * (kFrames + 2) + 255) / 256);      // smallest unit: 256 positions
  SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&s->flags), (size_t)s->flags_cap * sizeof(int)));
  SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&s->flags2), (size_t)s->flags_cap * sizeof(int)));
  SS_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&s->prof), kNumSMs * 8 * sizeof(long long)));
  SS_CUDA_CHECK(cudaMemset(s->prof, 0, kNumSMs * 8 * sizeof(long long)));
#define T(t, C, H, W) do { if ((rc = alloc_tensor(ctx, s, &s->t, B, C, H, W))) return rc; } while (0)
  T(m4, 64, 128, 256);
  T(p1, 32, 64, 128);
  T(m3, 128, 64, 128);
  // the tensors of the layers at 32 x 64 and below are plane-major (Tensor::plane_major): packed work units
#define TP(t, C, H, W) do { if ((rc = alloc_tensor(ctx, s, &s->t, B, C, H, W, true))) return rc; } while (0)
  TP(p2, 64, 32, 64);
  TP(m2, 192, 32, 64);
  TP(p3, 96, 16, 32);
  TP(m1, 256, 16, 32);
  TP(p4, 128, 8, 16);
  TP(bott, 128, 8, 16);
  if (is_split(prec)) {
    T(u4, 32, 64, 256);
    T(u3, 64, 32, 128);
  }
  T(c9, 32, 128, 256);
  T(spec, 32, 128, 256);
  T(t[RB_CONV1], 32, 128, 256);
  T(t[RB_CONV2], 64, 64, 128);
  TP(t[RB_CONV3], 96, 32, 64);
  TP(t[RB_CONV4], 128, 16, 32);
  TP(t[RB_BOTTLENECK], 128, 8, 16);
  TP(t[RB_ENCODER_OUT], 128, 8, 16);
  TP(t[RB_CONV6], 96, 16, 32);
  TP(t[RB_CONV7], 64, 32, 64);
#undef TP
  T(t[RB_CONV8], 32, 64, 128);
  T(t[RB_CONV9], 32, 128, 256);
  T(t[RB_SPEC], 32, 128, 256);
#undef T
  return SS_OK;
}

template <Prec P>
int classify_tc_p(ss_ctx* ctx, TcState* s, const float* mel, int n_windows, float* logits, float* spec_out,
                  cudaStream_t st) {
  int rc;
  s->launch_index = 0;
  for (int b0 = 0; b0 < n_windows; b0 += s->max_batch) {
    const int B = (n_windows - b0 < s->max_batch) ? (n_windows - b0) : s->max_batch;
#define SS_TRY(e) do { if ((rc = (e))) return rc; } while (0)
    const int64_t n_pix = (int64_t)B * kMels * kFrames;
    const float* mel_b = mel + (int64_t)b0 * kMels * kFrames;
    bool pooled1 = false;
    const char* dc = getenv("SS_TC_DIRECT_C1");
    if (dc == nullptr || atoi(dc) != 0) {
      // conv1_1: first convolution on CUDA cores straight from mel, second on the tensor cores with the
      // single-channel residual as an epilogue FMA
      conv1_direct<P><<<dim3(kMels / kC1Rows, 4, B), kFrames, 0, st>>>(mel_b, ctx->rb[RB_CONV1].c1.w, ctx->rb[RB_CONV1].c1.b,
                                                           s->t[RB_CONV1].data, s->t[RB_CONV1].lo, s->err);
      SS_CUDA_CHECK(cudaGetLastError());
      count_launch();
      // (the block-input argument only supplies the geometry here: conv2 reads t, the residual reads mel)
      SS_TRY(tc_res_block(s, RB_CONV1, s->t[RB_CONV1], 0, s->m4, 0, 0, B, st, nullptr, nullptr, true, mel_b,
                          ctx->rb[RB_CONV1].res.w, true, &s->p1, &pooled1));
    } else {
      // legacy form (A/B measurements, tests of the im2col'd operand tensor): both convolutions as tcgen05 launches;
      // its operand tensor is allocated on first use
      if (!s->x0.alloc) SS_TRY(alloc_tensor(ctx, s, &s->x0, s->max_batch, 16, 128, 256));
      mel_to_planar<P><<<(int)((n_pix + 255) / 256), 256, 0, st>>>(mel_b, s->x0.data, s->x0.lo, n_pix);
      SS_CUDA_CHECK(cudaGetLastError());
      count_launch();
      const char* sr = getenv("SS_TC_SCALAR_RES");
      const bool scalar = sr == nullptr || atoi(sr) != 0;
      SS_TRY(tc_res_block(s, RB_CONV1, s->x0, 0, s->m4, 0, 0, B, st, nullptr, nullptr, true,
                          scalar ? mel_b : nullptr, scalar ? ctx->rb[RB_CONV1].res.w : nullptr));
    }
    if (!pooled1) SS_TRY(tc_pool_p<P>(s->m4, 0, 4, s->p1, B, st));
    bool pooled2 = false;
    SS_TRY(tc_res_block(s, RB_CONV2, s->p1, 0, s->m3, 0, 0, B, st, nullptr, nullptr, true, nullptr, nullptr, false,
                        &s->p2, &pooled2));
    if (!pooled2) SS_TRY(tc_pool_p<P>(s->m3, 0, 8, s->p2, B, st));
    SS_TRY(tc_res_block(s, RB_CONV3, s->p2, 0, s->m2, 0, 0, B, st));
    SS_TRY(tc_pool_p<P>(s->m2, 0, 12, s->p3, B, st));
    SS_TRY(tc_res_block(s, RB_CONV4, s->p3, 0, s->m1, 0, 0, B, st));
    SS_TRY(tc_pool_p<P>(s->m1, 0, 16, s->p4, B, st));
    SS_TRY(tc_res_block(s, RB_BOTTLENECK, s->p4, 0, s->bott, 0, 0, B, st));
    SS_TRY(tc_res_block(s, RB_ENCODER_OUT, s->bott, 0, s->m1, 16, 1, B, st));   // -> up, cat after conv4
    SS_TRY(tc_res_block(s, RB_CONV6, s->m1, 0, s->m2, 12, 1, B, st));
    // Half-row tensors for the up-sampled halves of conv8's and conv9_1's inputs (TcSource::in_up): the split
    // precision with everything their consumers need — row-aligned launches, merged taps, separate launches, the
    // plain accumulation chain.  SS_TC_HALFROWS=0: 2 x 2 replicated planes of m3 / m4 as in every other mode.
    bool halfrows = false;
    if constexpr (PrecTraits<P>::split) {
      auto off = [](const char* name) { const char* e = getenv(name); return e != nullptr && atoi(e) == 0; };
      auto on = [](const char* name) { const char* e = getenv(name); return e != nullptr && atoi(e) != 0; };
      const char* sb = getenv("SS_TC_SUB");
      const char* pp = getenv("SS_TC_PAIRS");
      halfrows = !off("SS_TC_HALFROWS") && !off("SS_TC_ROWS") && !off("SS_TC_TAPMERGE") && !on("SS_TC_FUSE") &&
                 (sb == nullptr || atoi(sb) <= 1) && (pp == nullptr || (atoi(pp) & 1) == 0);
    }
    s->halfrows_last = halfrows;
    if (halfrows) {
      SS_TRY(tc_res_block(s, RB_CONV7, s->m2, 0, s->u3, 0, 2, B, st));
      SS_TRY(tc_res_block(s, RB_CONV8, s->m3, 0, s->u4, 0, 2, B, st, nullptr, nullptr, true, nullptr, nullptr, false,
                          nullptr, nullptr, &s->u3));
    } else {
      SS_TRY(tc_res_block(s, RB_CONV7, s->m2, 0, s->m3, 8, 1, B, st));
      SS_TRY(tc_res_block(s, RB_CONV8, s->m3, 0, s->m4, 4, 1, B, st));
    }
    const Tensor* up9 = halfrows ? &s->u4 : nullptr;
    // conv9_1 with the mask head's conv_flatten folded into its epilogue (TcConv::head_w); the 32-channel output
    // itself is stored only when the spec head will read it.  SS_TC_FUSE_HEAD=0 keeps the two-kernel form.
    const char* fh = getenv("SS_TC_FUSE_HEAD");
    if (fh == nullptr || atoi(fh) != 0) {
      SS_TRY(tc_res_block(s, RB_CONV9, s->m4, 0, s->c9, 0, 0, B, st, ctx->head.flat_w, s->head_part, spec_out != nullptr,
                          nullptr, nullptr, false, nullptr, nullptr, up9));
      mask_head_partials<<<dim3(kFrames / kHeadFrames, B), kHeadCols * kHeadGroups, 0, st>>>(
          s->head_part, ctx->head, logits + (int64_t)b0 * kFrames);
    } else {
      SS_TRY(tc_res_block(s, RB_CONV9, s->m4, 0, s->c9, 0, 0, B, st, nullptr, nullptr, true, nullptr, nullptr, false,
                          nullptr, nullptr, up9));
      mask_head_planar<P><<<dim3(kFrames / kHeadFrames, B), kHeadCols * kHeadGroups, 0, st>>>(
          s->c9.data, s->c9.lo, ctx->head, logits + (int64_t)b0 * kFrames);
    }
    SS_CUDA_CHECK(cudaGetLastError());
    count_launch();
    if (spec_out) {
      SS_TRY(tc_res_block(s, RB_SPEC, s->c9, 0, s->spec, 0, 0, B, st));
      spec_out_planar<P><<<(int)((n_pix + 255) / 256), 256, 0, st>>>(s->spec.data, s->spec.lo, ctx->head,
                                                                    spec_out + (int64_t)b0 * 2 * kMels * kFrames, n_pix);
      SS_CUDA_CHECK(cudaGetLastError());
      count_launch();
    }
#undef SS_TRY
  }
  return SS_OK;
}

bool prec_of_mode(int mode, Prec* p) {
  switch (mode) {
    case SS_MODE_BF16: *p = Prec::Bf16; return true;
    case SS_MODE_F16: *p = Prec::F16; return true;
    case SS_MODE_F16X3: *p = Prec::F16x3; return true;
  }
  return false;
}

}  // namespace

namespace {
void tc_free_state(ss_ctx* ctx, TcState* s) {
  auto free_guarded = [&](uint16_t* alloc) {
    if (!alloc) return;
    unregister_guards(ctx, alloc);
    cudaFree(alloc);
  };
  Tensor* ts[] = {&s->x0, &s->m4, &s->m3, &s->m2, &s->m1, &s->p1, &s->p2, &s->p3, &s->p4, &s->bott, &s->c9, &s->spec,
                  &s->u4, &s->u3};
  for (Tensor* t : ts) { free_guarded(t->alloc); free_guarded(t->alloc_lo); }
  for (int i = 0; i < RB_COUNT; ++i) {
    free_guarded(s->t[i].alloc);
    free_guarded(s->t[i].alloc_lo);
    for (PackedConv* pc : {&s->rb[i].c1, &s->rb[i].c2, &s->rb[i].res, &s->rb[i].c1_skip, &s->rb[i].c1_up}) {
      if (pc->w) cudaFree(pc->w);
      if (pc->w_hi) cudaFree(pc->w_hi);
    }
    if (s->rb[i].bias2) cudaFree(s->rb[i].bias2);
  }
  if (s->err) cudaFree(s->err);
  if (s->flags) cudaFree(s->flags);
  if (s->flags2) cudaFree(s->flags2);
  if (s->head_part) cudaFree(s->head_part);
  if (s->prof) cudaFree(s->prof);
  delete s;
}
}  // namespace

void tc_destroy(ss_ctx* ctx) {
  for (int slot = 0; slot < 3; ++slot) {
    TcState* s = static_cast<TcState*>(ctx->tc[slot]);
    if (!s) continue;
    tc_free_state(ctx, s);
    ctx->tc[slot] = nullptr;
  }
}

int classify_tc(ss_ctx* ctx, int mode, const float* mel, int n_windows, float* logits, float* spec_out,
                cudaStream_t st) {
  Prec prec;
  SS_REQUIRE(prec_of_mode(mode, &prec), SS_E_ARG, "mode %d is not a tensor-core mode", mode);
  TcState* s = nullptr;
  int rc = tc_build(ctx, prec, &s);
  if (rc) return rc;
  ctx->tc_last = (int)prec;
  switch (prec) {
    case Prec::Bf16: return classify_tc_p<Prec::Bf16>(ctx, s, mel, n_windows, logits, spec_out, st);
    case Prec::F16: return classify_tc_p<Prec::F16>(ctx, s, mel, n_windows, logits, spec_out, st);
    case Prec::F16x3: return classify_tc_p<Prec::F16x3>(ctx, s, mel, n_windows, logits, spec_out, st);
  }
  return SS_E_ARG;
}

// Debug / parity localisation: copy one internal activation of the last tensor-core classify call to NCHW f32.
// which: 0 conv1, 1 conv2, 2 conv3, 3 conv4, 4 bottleneck, 5 up(encoder_out), 6 up(conv6), 7 up(conv7),
//        8 up(conv8), 9 conv9, 10 t(conv1_1.conv1), 11 x0 (16 ch), 12 pool(conv1), 13 pool(conv2).
int tc_debug_dump(ss_ctx* ctx, int which, int n_windows, float* out, int* C, int* H, int* W, cudaStream_t st) {
  TcState* s = ctx->tc_last >= 0 ? static_cast<TcState*>(ctx->tc[ctx->tc_last]) : nullptr;
  SS_REQUIRE(s, SS_E_ARG, "tensor-core path not initialised");
  struct Sel { const Tensor* t; int plane0, c; };
  const Sel table[] = {{&s->m4, 0, 32}, {&s->m3, 0, 64}, {&s->m2, 0, 96}, {&s->m1, 0, 128}, {&s->bott, 0, 128},
                       {&s->m1, 16, 128}, {&s->m2, 12, 96}, {&s->m3, 8, 64}, {&s->m4, 4, 32}, {&s->c9, 0, 32},
                       {&s->t[RB_CONV1], 0, 32}, {&s->x0, 0, 16}, {&s->p1, 0, 32}, {&s->p2, 0, 64}};
  // + 0x100: twice the hi operands alone, + 0x200: twice the lo operands alone (split precision; parity localisation)
  const int part = which & 0x300;
  which &= 0xff;
  SS_REQUIRE(which >= 0 && which < (int)(sizeof(table) / sizeof(table[0])), SS_E_ARG, "bad activation id %d", which);
  Sel e = table[which];
  // up(conv7) / up(conv8) live in the half-row tensors when the last call used them (full image geometry for the dump)
  int halfrows = 0;
  Tensor full_geom;
  int64_t is = e.t->img_stride(), ps = e.t->plane_stride();
  if (s->halfrows_last && (which == 7 || which == 8)) {
    full_geom = which == 7 ? s->u3 : s->u4;
    is = full_geom.img_stride();
    ps = full_geom.plane_stride();
    full_geom.H *= 2;
    e = Sel{&full_geom, 0, e.c};
    halfrows = 1;
  }
  Tensor alias = *e.t;
  if (part == 0x100) alias.lo = alias.data;
  if (part == 0x200) alias.data = alias.lo;
  e.t = &alias;
  *C = e.c; *H = e.t->H; *W = e.t->W;
  if (out) {
    const int64_t total = (int64_t)n_windows * e.c * e.t->H * e.t->W;
    const int grid = (int)((total + 255) / 256);
    switch (s->prec) {
      case Prec::Bf16: planar_to_nchw<Prec::Bf16><<<grid, 256, 0, st>>>(e.t->data, e.t->lo, is, ps, e.plane0, e.c, e.t->H, e.t->W, out, total, halfrows); break;
      case Prec::F16: planar_to_nchw<Prec::F16><<<grid, 256, 0, st>>>(e.t->data, e.t->lo, is, ps, e.plane0, e.c, e.t->H, e.t->W, out, total, halfrows); break;
      case Prec::F16x3: planar_to_nchw<Prec::F16x3><<<grid, 256, 0, st>>>(e.t->data, e.t->lo, is, ps, e.plane0, e.c, e.t->H, e.t->W, out, total, halfrows); break;
    }
    SS_CUDA_CHECK(cudaGetLastError());
  }
  int herr = 0;
  SS_CUDA_CHECK(cudaMemcpyAsync(&herr, s->err, sizeof(int), cudaMemcpyDeviceToHost, st));
  SS_CUDA_CHECK(cudaStreamSynchronize(st));
  SS_REQUIRE(herr == 0, SS_E_CUDA, "tcgen05 pipeline timed out (role code %d)", herr);
  return SS_OK;
}

// Debug: select which conv launch (0-based, in issue order within one ss_classify call) of the most recently used
// tensor-core mode records role timers, and read the timers back ([148][8] int64: producer empty-wait, mma
// acc-empty-wait, mma full-wait, mma total, epilogue acc-full-wait, epilogue total, -, units).
int tc_debug_profile(ss_ctx* ctx, int select_launch, long long* out_host) {
  TcState* s = ctx->tc_last >= 0 ? static_cast<TcState*>(ctx->tc[ctx->tc_last]) : nullptr;
  SS_REQUIRE(s, SS_E_ARG, "tensor-core path not initialised");
  if (out_host) {
    SS_CUDA_CHECK(cudaDeviceSynchronize());
    SS_CUDA_CHECK(cudaMemcpy(out_host, s->prof, kNumSMs * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
  }
  s->prof_layer = select_launch;
  return SS_OK;
}

// The pipeline's bounded waits flag a time-out in device memory and the fp16-operand epilogues flag activations that
// left the fp16 range; surface both (0 = healthy) and clear the range flag so that the next call starts clean.
int tc_error_flag(ss_ctx* ctx, int* flag, int* range_flag, cudaStream_t st) {
  *flag = 0;
  *range_flag = 0;
  for (int slot = 0; slot < 3; ++slot) {
    TcState* s = static_cast<TcState*>(ctx->tc[slot]);
    if (!s) continue;
    int h[4] = {0, 0, 0, 0};
    SS_CUDA_CHECK(cudaMemcpyAsync(h, s->err, sizeof(h), cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(cudaStreamSynchronize(st));
    if (h[0]) *flag = h[0];
    if (h[2]) {
      *range_flag = 1;
      SS_CUDA_CHECK(cudaMemsetAsync(s->err + 2, 0, sizeof(int), st));
    }
  }
  return SS_OK;
}

}  // namespace ss
