// K2 (parity mode) — float32 CUDA-core implementation of the residual U-Net trunk.
//
// Replaces ResBlock x11, MaxPool2d x4, Upsample(nearest) x4 and torch.cat x4
// (root/code/backend/pytorch_neural_nets.py:7-41,102-123,156-185) with BatchNorm folded into the
// convolution weights (softspoken_b200/checkpoint.py:fold_bn).  Activations are NHWC float32.  The
// convolution kernel is a shared-memory tiled direct convolution (8x16 pixel tile, 8 pixels x 8 output
// channels per thread); its loader applies nearest-upsampling (src = dst >> 1) and the channel
// concatenation [skip, upsampled] on the fly, so neither `Upsample` nor `cat` touches HBM; its epilogue
// fuses bias, the residual add and ReLU.  This path exists for bit-for-bit-stable float32 parity with
// the reference's CPU arithmetic (|delta logit| ~ 1e-6); the throughput path is conv_tc.cu.
#include "ss_common.cuh"

namespace ss {

namespace {

constexpr int TH = 8, TW = 16;   // output tile (rows x cols)
constexpr int CK = 8;            // input channels per shared-memory stage

template <int KS>
struct Halo {
  static constexpr int P = KS / 2;
  static constexpr int H = TH + 2 * P;
  static constexpr int W = TW + 2 * P;
  static constexpr int WP = W + 1;   // padded row (odd stride)
};

// in0: [B,H,W,C0]; in1 (optional): [B,H/2,W/2,C1], nearest-upsampled and concatenated after in0.
template <int KS, int COUT>
__global__ void __launch_bounds__((COUT / 8) * 16)
conv_nhwc_f32(const float* __restrict__ in0, int C0, const float* __restrict__ in1, int C1, int H, int W,
              const float* __restrict__ wgt, const float* __restrict__ bias, const float* __restrict__ res,
              int relu, float* __restrict__ out) {
  using HL = Halo<KS>;
  constexpr int CG = COUT / 8;
  constexpr int NT = CG * 16;
  constexpr int TAPS = KS * KS;
  __shared__ float in_s[CK][HL::H][HL::WP];
  __shared__ __align__(16) float w_s[TAPS][CK][COUT];

  const int tid = threadIdx.x;
  const int cg = tid % CG;
  const int pg = tid / CG;
  const int prow = pg >> 1;
  const int pcol = (pg & 1) * 8;
  const int b = blockIdx.z;
  const int ty0 = blockIdx.y * TH, tx0 = blockIdx.x * TW;
  const int Cin = C0 + C1;

  float acc[8][8];
#pragma unroll
  for (int p = 0; p < 8; ++p)
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[p][o] = 0.f;

  for (int ci0 = 0; ci0 < Cin; ci0 += CK) {
    // ---- stage input halo tile (channel-major planes)
    for (int i = tid; i < HL::H * HL::W; i += NT) {
      const int y = i / HL::W, x = i % HL::W;
      const int gy = ty0 + y - HL::P, gx = tx0 + x - HL::P;
      float v[CK];
#pragma unroll
      for (int c = 0; c < CK; ++c) v[c] = 0.f;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        if (ci0 < C0) {
          const float* p = in0 + (((int64_t)b * H + gy) * W + gx) * C0 + ci0;
          if ((C0 & 7) == 0) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p));
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(p) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = c4.x; v[5] = c4.y; v[6] = c4.z; v[7] = c4.w;
          } else {
#pragma unroll
            for (int c = 0; c < CK; ++c)
              if (ci0 + c < C0) v[c] = __ldg(p + c);
          }
        } else {
          const int H2 = H >> 1, W2 = W >> 1;
          const float* p = in1 + (((int64_t)b * H2 + (gy >> 1)) * W2 + (gx >> 1)) * C1 + (ci0 - C0);
          const float4 a = __ldg(reinterpret_cast<const float4*>(p));
          const float4 c4 = __ldg(reinterpret_cast<const float4*>(p) + 1);
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
          v[4] = c4.x; v[5] = c4.y; v[6] = c4.z; v[7] = c4.w;
        }
      }
#pragma unroll
      for (int c = 0; c < CK; ++c) in_s[c][y][x] = v[c];
    }
    // ---- stage weights [tap][ci0..ci0+7][COUT]
    for (int i = tid; i < TAPS * CK * (COUT / 4); i += NT) {
      const int o4 = i % (COUT / 4);
      const int c = (i / (COUT / 4)) % CK;
      const int t = i / (COUT / 4 * CK);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ci0 + c < Cin)
        v = __ldg(reinterpret_cast<const float4*>(wgt + ((int64_t)t * Cin + ci0 + c) * COUT) + o4);
      reinterpret_cast<float4*>(&w_s[t][c][0])[o4] = v;
    }
    __syncthreads();

#pragma unroll
    for (int c = 0; c < CK; ++c) {
#pragma unroll
      for (int ky = 0; ky < KS; ++ky) {
        float a[8 + KS - 1];
#pragma unroll
        for (int x = 0; x < 8 + KS - 1; ++x) a[x] = in_s[c][prow + ky][pcol + x];
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
          const float4 w0 = *reinterpret_cast<const float4*>(&w_s[ky * KS + kx][c][cg * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&w_s[ky * KS + kx][c][cg * 8 + 4]);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int p = 0; p < 8; ++p)
#pragma unroll
            for (int o = 0; o < 8; ++o) acc[p][o] = fmaf(a[p + kx], wv[o], acc[p][o]);
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue: bias (+ residual) (+ ReLU)
  float bv[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) bv[o] = __ldg(bias + cg * 8 + o);
  const int gy = ty0 + prow;
  if (gy < H) {
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int gx = tx0 + pcol + p;
      if (gx >= W) continue;
      const int64_t off = (((int64_t)b * H + gy) * W + gx) * COUT + cg * 8;
      float r[8];
#pragma unroll
      for (int o = 0; o < 8; ++o) r[o] = acc[p][o] + bv[o];
      if (res) {
        const float4 r0 = __ldg(reinterpret_cast<const float4*>(res + off));
        const float4 r1 = __ldg(reinterpret_cast<const float4*>(res + off) + 1);
        r[0] += r0.x; r[1] += r0.y; r[2] += r0.z; r[3] += r0.w;
        r[4] += r1.x; r[5] += r1.y; r[6] += r1.z; r[7] += r1.w;
      }
      if (relu) {
#pragma unroll
        for (int o = 0; o < 8; ++o) r[o] = fmaxf(r[o], 0.f);
      }
      reinterpret_cast<float4*>(out + off)[0] = make_float4(r[0], r[1], r[2], r[3]);
      reinterpret_cast<float4*>(out + off)[1] = make_float4(r[4], r[5], r[6], r[7]);
    }
  }
}

__global__ void maxpool2_nhwc_f32(const float* __restrict__ in, int H, int W, int C, float* __restrict__ out,
                                  int64_t total4) {
  // out: [B,H/2,W/2,C]; one float4 (4 channels) per thread
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int C4 = C >> 2;
  const int c4 = (int)(i % C4);
  int64_t r = i / C4;
  const int W2 = W >> 1, H2 = H >> 1;
  const int x = (int)(r % W2); r /= W2;
  const int y = (int)(r % H2);
  const int64_t b = r / H2;
  const float4* p = reinterpret_cast<const float4*>(in + (((b * H + 2 * y) * W + 2 * x) * (int64_t)C)) + c4;
  const float4 a = __ldg(p), bq = __ldg(p + C4);
  const float4 c = __ldg(p + (int64_t)W * C4), d = __ldg(p + (int64_t)W * C4 + C4);
  float4 m;
  m.x = fmaxf(fmaxf(a.x, bq.x), fmaxf(c.x, d.x));
  m.y = fmaxf(fmaxf(a.y, bq.y), fmaxf(c.y, d.y));
  m.z = fmaxf(fmaxf(a.z, bq.z), fmaxf(c.z, d.z));
  m.w = fmaxf(fmaxf(a.w, bq.w), fmaxf(c.w, d.w));
  reinterpret_cast<float4*>(out)[i] = m;
}

template <int KS>
int launch_conv(const float* in0, int C0, const float* in1, int C1, int B, int H, int W, const ConvW& cw,
                const float* res, int relu, float* out, cudaStream_t st) {
  SS_REQUIRE(cw.cin == C0 + C1 && cw.taps == KS * KS, SS_E_ARG, "conv shape mismatch: cin %d vs %d+%d, taps %d",
             cw.cin, C0, C1, cw.taps);
  dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B);
  switch (cw.cout) {
    case 32: conv_nhwc_f32<KS, 32><<<grid, 64, 0, st>>>(in0, C0, in1, C1, H, W, cw.w, cw.b, res, relu, out); break;
    case 64: conv_nhwc_f32<KS, 64><<<grid, 128, 0, st>>>(in0, C0, in1, C1, H, W, cw.w, cw.b, res, relu, out); break;
    case 96: conv_nhwc_f32<KS, 96><<<grid, 192, 0, st>>>(in0, C0, in1, C1, H, W, cw.w, cw.b, res, relu, out); break;
    case 128: conv_nhwc_f32<KS, 128><<<grid, 256, 0, st>>>(in0, C0, in1, C1, H, W, cw.w, cw.b, res, relu, out); break;
    default: SS_REQUIRE(false, SS_E_ARG, "unsupported C_out %d", cw.cout);
  }
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

int res_block(ss_ctx* ctx, int which, const float* in0, int C0, const float* in1, int C1, int B, int H, int W,
              float* out, cudaStream_t st) {
  const ResBlockW& rb = ctx->rb[which];
  int rc;
  if ((rc = launch_conv<1>(in0, C0, in1, C1, B, H, W, rb.res, nullptr, 0, ctx->ws.tmp_r, st))) return rc;
  if ((rc = launch_conv<3>(in0, C0, in1, C1, B, H, W, rb.c1, nullptr, 1, ctx->ws.tmp_t, st))) return rc;
  return launch_conv<3>(ctx->ws.tmp_t, rb.c1.cout, nullptr, 0, B, H, W, rb.c2, ctx->ws.tmp_r, 1, out, st);
}

int pool(const float* in, int B, int H, int W, int C, float* out, cudaStream_t st) {
  const int64_t total4 = (int64_t)B * (H / 2) * (W / 2) * (C / 4);
  maxpool2_nhwc_f32<<<(int)((total4 + 255) / 256), 256, 0, st>>>(in, H, W, C, out, total4);
  SS_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return SS_OK;
}

}  // namespace

int classify_fp32(ss_ctx* ctx, const float* mel, int n_windows, float* logits, float* spec_out, cudaStream_t st) {
  {
    const int rc0 = ensure_workspace_f32(ctx);
    if (rc0) return rc0;
  }
  WorkspaceF32& ws = ctx->ws;
  for (int b0 = 0; b0 < n_windows; b0 += ctx->f32_batch) {
    const int B = (n_windows - b0 < ctx->f32_batch) ? (n_windows - b0) : ctx->f32_batch;
    const float* x = mel + (int64_t)b0 * kMels * kFrames;   // [B,128,256,1]
    int rc;
#define SS_TRY(e) do { if ((rc = (e))) return rc; } while (0)
    SS_TRY(res_block(ctx, RB_CONV1, x, 1, nullptr, 0, B, 128, 256, ws.conv1, st));
    SS_TRY(pool(ws.conv1, B, 128, 256, 32, ws.pool1, st));
    SS_TRY(res_block(ctx, RB_CONV2, ws.pool1, 32, nullptr, 0, B, 64, 128, ws.conv2, st));
    SS_TRY(pool(ws.conv2, B, 64, 128, 64, ws.pool2, st));
    SS_TRY(res_block(ctx, RB_CONV3, ws.pool2, 64, nullptr, 0, B, 32, 64, ws.conv3, st));
    SS_TRY(pool(ws.conv3, B, 32, 64, 96, ws.pool3, st));
    SS_TRY(res_block(ctx, RB_CONV4, ws.pool3, 96, nullptr, 0, B, 16, 32, ws.conv4, st));
    SS_TRY(pool(ws.conv4, B, 16, 32, 128, ws.pool4, st));
    SS_TRY(res_block(ctx, RB_BOTTLENECK, ws.pool4, 128, nullptr, 0, B, 8, 16, ws.bott, st));
    SS_TRY(res_block(ctx, RB_ENCODER_OUT, ws.bott, 128, nullptr, 0, B, 8, 16, ws.enc, st));
    // decoder: cat([skip, up(x)]) fused into the loader (skip first: pytorch_neural_nets.py:171-180)
    SS_TRY(res_block(ctx, RB_CONV6, ws.conv4, 128, ws.enc, 128, B, 16, 32, ws.conv6, st));
    SS_TRY(res_block(ctx, RB_CONV7, ws.conv3, 96, ws.conv6, 96, B, 32, 64, ws.conv7, st));
    SS_TRY(res_block(ctx, RB_CONV8, ws.conv2, 64, ws.conv7, 64, B, 64, 128, ws.conv8, st));
    SS_TRY(res_block(ctx, RB_CONV9, ws.conv1, 32, ws.conv8, 32, B, 128, 256, ws.conv9, st));
    SS_TRY(launch_mask_head_f32(ctx, ws.conv9, B, logits + (int64_t)b0 * kFrames, st));
    if (spec_out) {
      SS_TRY(res_block(ctx, RB_SPEC, ws.conv9, 32, nullptr, 0, B, 128, 256, ws.spec, st));
      SS_TRY(launch_spec_out_f32(ctx, ws.spec, B, spec_out + (int64_t)b0 * 2 * kMels * kFrames, st));
    }
#undef SS_TRY
  }
  return SS_OK;
}

}  // namespace ss
