// K3 / K4 tail — the two output heads of SpecUNet_2D on float32 NHWC activations.
//
// mask head (root/code/backend/pytorch_neural_nets.py:133-140,188-195):
//   conv_flatten Conv2d(32,4,(128,1)) + ReLU + squeeze -> ResBlock1D(4,4) -> Conv1d(4,1,1): one raw logit
//   per frame (no sigmoid anywhere in the reference).  One CTA per window, one thread per frame: the
//   K = 4096 reduction runs down the 128 mel rows of that frame's column (weights are warp-uniform
//   broadcast loads, activations fully coalesced 128-byte pixels), then the 4-channel 1-D residual block
//   is evaluated through shared memory with zero padding at frames -1 and 256.
// spec head tail (pytorch_neural_nets.py:128-130,184-185): Conv2d(32,2,1) + ReLU, written NCHW as the
//   reference returns it.
#include "ss_common.cuh"

namespace ss {

namespace {

__global__ void __launch_bounds__(kFrames)
mask_head_f32(const float* __restrict__ conv9, HeadW hw, float* __restrict__ logits) {
  __shared__ float xf[4][kFrames + 2];
  __shared__ float c1[4][kFrames + 2];
  const int t = threadIdx.x;
  const int b = blockIdx.x;
  const float* col = conv9 + ((int64_t)b * kMels * kFrames + t) * 32;   // pixel (h=0, t)

  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int h = 0; h < kMels; ++h) {
    const float4* px = reinterpret_cast<const float4*>(col + (int64_t)h * kFrames * 32);
    const float4* wrow = reinterpret_cast<const float4*>(hw.flat_w + (int64_t)h * 32 * 4);
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 a = __ldg(px + c4);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 w = __ldg(wrow + c4 * 4 + k);   // weights of input channel 4*c4+k -> 4 outputs
        acc[0] = fmaf(av[k], w.x, acc[0]);
        acc[1] = fmaf(av[k], w.y, acc[1]);
        acc[2] = fmaf(av[k], w.z, acc[2]);
        acc[3] = fmaf(av[k], w.w, acc[3]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) xf[c][t + 1] = fmaxf(acc[c] + __ldg(hw.flat_b + c), 0.f);
  if (t < 4) {
    xf[t][0] = 0.f; xf[t][kFrames + 1] = 0.f;
    c1[t][0] = 0.f; c1[t][kFrames + 1] = 0.f;
  }
  __syncthreads();

  // ResBlock1D.conv1 + BN + ReLU
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    float v = __ldg(hw.c1_b + co);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) v = fmaf(__ldg(hw.c1_w + (k * 4 + ci) * 4 + co), xf[ci][t + k], v);
    c1[co][t + 1] = fmaxf(v, 0.f);
  }
  __syncthreads();

  float logit = __ldg(hw.out_b);
#pragma unroll
  for (int co = 0; co < 4; ++co) {
    float v = __ldg(hw.c2_b + co);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) v = fmaf(__ldg(hw.c2_w + (k * 4 + ci) * 4 + co), c1[ci][t + k], v);
    float r = __ldg(hw.res_b + co);
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) r = fmaf(__ldg(hw.res_w + ci * 4 + co), xf[ci][t + 1], r);
    logit = fmaf(__ldg(hw.out_w + co), fmaxf(v + r, 0.f), logit);
  }
  logits[(int64_t)b * kFrames + t] = logit;
}

__global__ void spec_out_f32(const float* __restrict__ x, HeadW hw, float* __restrict__ out, int64_t n_pixels) {
  // x: [B,128,256,32] -> out: [B,2,128,256]
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pixels) return;
  const float4* px = reinterpret_cast<const float4*>(x + i * 32);
  float a0 = __ldg(hw.spec_b), a1 = __ldg(hw.spec_b + 1);
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    const float4 a = __ldg(px + c4);
    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a0 = fmaf(av[k], __ldg(hw.spec_w + (c4 * 4 + k) * 2), a0);
      a1 = fmaf(av[k], __ldg(hw.spec_w + (c4 * 4 + k) * 2 + 1), a1);
    }
  }
  const int64_t plane = (int64_t)kMels * kFrames;
  const int64_t b = i / plane, p = i % plane;
  out[(b * 2) * plane + p] = fmaxf(a0, 0.f);
  out[(b * 2 + 1) * plane + p] = fmaxf(a1, 0.f);
}

}  // namespace

int launch_mask_head_f32(const ss_ctx* ctx, const float* conv9_nhwc, int n_windows, float* logits, cudaStream_t st) {
  if (n_windows <= 0) return SS_OK;
  mask_head_f32<<<n_windows, kFrames, 0, st>>>(conv9_nhwc, ctx->head, logits);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int launch_spec_out_f32(const ss_ctx* ctx, const float* spec_nhwc, int n_windows, float* spec_out_nchw,
                        cudaStream_t st) {
  if (n_windows <= 0) return SS_OK;
  const int64_t n_pixels = (int64_t)n_windows * kMels * kFrames;
  spec_out_f32<<<(int)((n_pixels + 255) / 256), 256, 0, st>>>(spec_nhwc, ctx->head, spec_out_nchw, n_pixels);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

}  // namespace ss
