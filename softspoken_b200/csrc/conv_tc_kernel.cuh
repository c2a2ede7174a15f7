// The tcgen05 implicit-GEMM convolution kernel of conv_tc.cu (see that file's header for the layout).
//
// Persistent, warp-specialised, one CTA per SM, 11 warps (default role layout; TcJob::layout):
//   warp 8     producer  — cp.async.bulk (1-D TMA) of activation runs + packed weights into an smem ring
//   warps 9,10 MMA       — one elected thread per warp issues tcgen05.mma into the two 256-column TMEM buffers.
//              Each warp owns half of every unit's accumulator tiles (G = 2: one group each; G = 1: half the tiles)
//              and walks every stage for them.  The tensor pipe queues only a couple of MMAs, so whatever an issuer
//              does between two MMAs (barrier probes, descriptor set-up: ~500 cycles per stage) would be a pipe
//              bubble; with two independent issuers one warp's bookkeeping runs under the other's MMAs.  Every
//              accumulator tile still has a single issuing thread and a fixed order, so results are deterministic.
//   warps 0-7  epilogue — tcgen05.ld -> (sum of the accumulation groups) -> bias / ReLU / border mask -> 16-bit pack
//              -> 16- / 32-byte global stores (a warp reads the TMEM lane quadrant warp % 4; the two warps of a
//              quadrant take alternate tiles)
// Producer and MMA issuers walk a flat per-unit stage program built by the host (TcJob::prog).
// The three roles are decoupled by mbarriers (full/empty per smem stage, acc_full/acc_empty per TMEM
// buffer), so the loads of unit k+1, the MMAs of unit k and the epilogue of unit k-1 overlap.
//
// A work unit is G groups of MT 128-position tiles.  G = 1: one group per unit, the two TMEM buffers alternate
// between consecutive units.  G = 2 (large images): a unit spans both buffers — one staged run of 2*MT*128
// positions and one copy of the chunk's weights feed both groups, which cuts the L2 -> shared-memory traffic
// per output position (halo and weights amortised over twice the positions; the N=32 layers are otherwise
// bound by exactly that traffic).  The groups still complete one after the other (group 0's MMAs of the last
// chunk are issued, and committed, before group 1's), so the epilogue of a group overlaps the MMAs of the next.
//
// Row-aligned units (template parameter Rows, dual layout at image widths of 128 / 256): a unit is 2 or 4 whole image
// rows instead of consecutive positions of the padded image, which makes four things possible (all documented at the
// TcConv / TcSource fields they add): MaxPool2d(2) in the epilogue (TcConv::pool_out), six instead of nine tap MMAs on
// up-sampled inputs ("rowdup" sources, prog_entry), inputs stored with one row per row pair (TcSource::in_up,
// TcConv::upsample == 2) and the launch's weights resident in shared memory (TcConv::wres).
//
// Operand precision is a template parameter:
//   Bf16   — bf16 operands, one pass (throughput mode; ~4e-2 relative logit error over the 25-layer stack);
//   F16    — fp16 operands, one pass (8x finer mantissa at the same cost);
//   F16x3  — every fp32 value v is carried as hi = fp16(v), lo = fp16(v - hi) in two tensors and every product
//            as hi*hi + hi*lo + lo*hi, all accumulated in the same fp32 TMEM tile: 3x the MMAs, ~2^-22
//            relative operand error, which is what lets tensor-core logits meet the 1e-4 parity budget.
//            The kernel sees the split as extra "sources".  For C_out <= 64 ("dual" layout) x_hi is multiplied
//            with [w_hi | w_lo] in ONE MMA of 2 C_out columns (the A tile — the shared-memory bottleneck of
//            small-N MMAs — is fetched once instead of twice), main and correction terms land in separate
//            TMEM column groups, x_lo . w_hi accumulates into the correction group and the epilogue adds the
//            two.  For C_out >= 96 the three products are separate sources into one accumulator, corrections
//            first (the tensor core truncates the fp32 accumulator after every MMA, so small terms are summed
//            while the accumulator is still small).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ss_common.cuh"

namespace ss {
namespace tc {

constexpr int kMaxStages = 8;
constexpr int kMaxSources = 6;
constexpr int kTcThreads = 352;          // 11 warps: 8 epilogue, 1 producer, 2 MMA issuers (roles: see the kernel)
constexpr int kAccCols = 256;             // TMEM columns per accumulator buffer (two buffers = all 512)
constexpr uint32_t kSpinLimit = 1u << 22;
// Tuning experiments (SS_TC_DEBUG: 1 skip the copies, 2 skip the stores, 4 skip the stage barriers, 8 skip the MMAs)
// are compiled in only with -DSS_TC_DEBUG_HOOKS=1; the production kernel carries none of their tests.
#ifndef SS_TC_DEBUG_HOOKS
#define SS_TC_DEBUG_HOOKS 0
#endif
constexpr bool kDebugHooks = SS_TC_DEBUG_HOOKS != 0;


enum class Prec : int { Bf16 = 0, F16 = 1, F16x3 = 2 };

template <Prec P> struct PrecTraits;
template <> struct PrecTraits<Prec::Bf16> { static constexpr bool split = false; static constexpr uint32_t fmt = 1; };
template <> struct PrecTraits<Prec::F16> { static constexpr bool split = false; static constexpr uint32_t fmt = 0; };
template <> struct PrecTraits<Prec::F16x3> { static constexpr bool split = true; static constexpr uint32_t fmt = 0; };

// two floats -> one packed pair of 16-bit operands (and the packed residuals for the split format)
// fp16 saturates instead of overflowing to inf (activations beyond +-65504 are outside what this mode carries): the
// conversion itself clamps (cvt.rn.satfinite.f16x2.f32 -> F2FP.SATFINITE, no FMNMX pair per value), and a kernel
// detects that it happened from the packed result (range_track / range_hit below).
__device__ __forceinline__ uint32_t cvt_f16x2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));      // a -> low half
  return r;
}
template <Prec P>
__device__ __forceinline__ uint32_t pack_hi(float a, float b) {
  if constexpr (PrecTraits<P>::fmt == 1) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    return cvt_f16x2_sat(a, b);
  }
}
// (one name was the saturating form, the other not, while the clamp was a pair of FMNMX in the caller)
template <Prec P>
__device__ __forceinline__ uint32_t pack_rn(float a, float b) { return pack_hi<P>(a, b); }
// the plain conversion, for callers that clamp themselves (the folded-pool epilogue, see there)
template <Prec P>
__device__ __forceinline__ uint32_t pack_plain(float a, float b) {
  if constexpr (PrecTraits<P>::fmt == 1) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
// Range check of the fp16 operand modes on NON-NEGATIVE packed pairs (everything here is stored after a ReLU): a
// running packed maximum (one HMNMX2 per pair of values), tested once per unit.  An activation of 65,488 or more
// rounds or saturates to the fp16 maximum 0x7BFF and counts as out of range (SS_E_RANGE).
__device__ __forceinline__ uint32_t range_track(uint32_t running, uint32_t packed) {
  const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&running), *reinterpret_cast<const __half2*>(&packed));
  return *reinterpret_cast<const uint32_t*>(&m);
}
__device__ __forceinline__ bool range_hit(uint32_t running) {
  return (running & 0x7fffu) >= 0x7bffu || ((running >> 16) & 0x7fffu) >= 0x7bffu;
}
__device__ __forceinline__ uint32_t pack_lo_f16(float a, float b, uint32_t hi) {
  const float2 d = sub2(make_float2(a, b), __half22float2(*reinterpret_cast<const __half2*>(&hi)));   // one FADD2
  __half2 l = __floats2half2_rn(d.x, d.y);
  return *reinterpret_cast<uint32_t*>(&l);
}
template <Prec P>
__device__ __forceinline__ float2 unpack2(uint32_t w) {
  if constexpr (PrecTraits<P>::fmt == 1) return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
  else return __half22float2(*reinterpret_cast<const __half2*>(&w));
}

// ------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
#ifndef SS_TC_SPIN
#define SS_TC_SPIN 0
#endif
// Bounded wait: a protocol bug must not hang the GPU.  Returns false (and flags the error) on timeout.
// kSpin (producer / MMA warps, -DSS_TC_SPIN=1 builds): poll with the non-blocking test_wait instead of the
// suspending try_wait.
template <bool kSpin = false>
__device__ __noinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
#pragma unroll 1
  for (uint32_t i = 0; i < kSpinLimit; ++i) {
    if constexpr (kSpin && SS_TC_SPIN) { if (mbar_test_wait(bar, parity)) return true; }
    else { if (mbar_try_wait(bar, parity)) return true; }
  }
  atomicExch(err, code);
  return false;
}
template <bool kSpin = false>
__device__ __forceinline__ bool mbar_wait_t(uint32_t bar, uint32_t parity, int* err, int code, long long& acc) {
  const long long t0 = clock64();
  const bool ok = mbar_wait<kSpin>(bar, parity, err, code);
  acc += clock64() - t0;
  return ok;
}
// Hot-path wait: one inline probe (the common case on the MMA warp: the barrier completed long ago), the bounded
// loop only when it has not; cycle accounting only when a profile was requested.
__device__ __forceinline__ bool mbar_wait_fast(uint32_t bar, uint32_t parity, int* err, int code, bool timing,
                                               long long& acc) {
  if (mbar_try_wait(bar, parity)) return true;
  if (timing) return mbar_wait_t<true>(bar, parity, err, code, acc);
  return mbar_wait<true>(bar, parity, err, code);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
#ifndef SS_TC_STCS
#define SS_TC_STCS 0
#endif
// Epilogue store of one 16-byte vector.  -DSS_TC_STCS=1: st.global.cs (evict-first), an experiment.
__device__ __forceinline__ void st16(uint16_t* p, uint4 v) {
#if SS_TC_STCS
  __stcs(reinterpret_cast<uint4*>(p), v);
#else
  *reinterpret_cast<uint4*>(p) = v;
#endif
}
// The same 16-byte vector to two neighbouring positions as ONE 32-byte store (STG.256; p is 32-byte aligned, see
// conv_tc.cu:alloc_guarded): the 2 x 2 replication of the up-sampling epilogue is two of these per row pair instead of
// four 16-byte stores, each warp writing whole 32-byte sectors.
__device__ __forceinline__ void st32x2(uint16_t* p, uint4 v) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// One lane of a converged warp (the CUTLASS elect_one_sync idiom).  Keeping the role loops warp-uniform and
// predicating only the issuing instruction lets the compiler hold descriptors in uniform registers, which
// tcgen05.mma / cp.async.bulk consume directly (no per-issue R2UR waterfall).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0, lane_out = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %2;\n\t"
      "@px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, rx;\n\t}"
      : "+r"(lane_out), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}

// K-major, un-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
//   [0,14) start >> 4 | [16,30) LBO >> 4 (stride between the two 16-byte K chunks) |
//   [32,46) SBO >> 4 (stride between 8-row core matrices) | [46,48) version = 1 | [61,64) layout = 0.
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = fmt (0 F16, 1 BF16), K-major, M = 128.
__host__ __device__ constexpr uint32_t instr_desc(int n, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// --------------------------------------------------------------------------------------------- parameters
struct TcSource {
  const uint16_t* in;  // planar-8 padded tensor of 16-bit operands at the layer's resolution
  int planes_total;    // C/8 of that tensor
  int plane0;          // first plane this convolution reads
  int n_chunks;        // C_in / 16
  int taps;            // 9 (3x3), 1 (1x1, centre) or 12 (3x3 over a row-replicated image: see "rowdup" below)
  int kind;            // 0: N columns into the main columns; 1: 2N columns (main | correction), dual weights;
                       // 2: N columns into the correction columns
  int ring;            // 0: image b of the batch is image b of the tensor; R > 0: it lives in slot b % R (the
                       // intermediate tensor of a fused ResBlock launch, see TcJob)
  int w_stride;        // 16-bit elements between consecutive chunks of `w`
  const uint16_t* w;   // plain: [n_chunks][parts][taps][2][N][8] (pointing at the part to use);
                       // dual:  [n_chunks][taps][2][2N][8] (rows 0..N-1 = w_hi, N..2N-1 = w_lo)
  // Resident weights (TcConv::wres > 0, dual layout): where this source's chunk 0 sits in the resident region, the
  // bytes between consecutive chunks, and the rows of a K-half block there (2N: a kind-2 source reads the w_hi rows of
  // its kind-1 sibling's dual matrices instead of a copy of its own).  wres_bytes > 0: this source brings the block in.
  int wres_off, wres_stride, w_rows, wres_bytes;
  // Half-row tensor (row-aligned launches): K-chunks >= up_chunk0 are the nearest-neighbour 2x up-sampled output of
  // the block below, stored with its COLUMNS replicated only: [B][up_planes_total][H/2 + 2][W + 2][8], row 1 + k of
  // the tensor = image rows 2k and 2k + 1 (TcConv::upsample == 2 writes it).  The producer stages padded image row
  // rho from tensor row ((rho - 1) >> 1) + 1, one bulk copy per row: the up-sampled half of a decoder input is
  // written once and read from DRAM once per pair of rows instead of twice.  null: every chunk comes from `in`.
  const uint16_t* in_up;
  int up_chunk0, up_planes_total;
  // Where image b / plane k of `in` starts, in positions (16-byte vectors): image-major tensors [B][C/8][..] have
  // img_stride = planes * (H+2)(W+2), plane_stride = (H+2)(W+2); plane-major ones [C/8][B_cap][..] (the small images
  // of the deep layers, see TcConv::packed) img_stride = (H+2)(W+2), plane_stride = B_cap * (H+2)(W+2).
  int64_t img_stride, plane_stride;
};

struct TcConv {
  TcSource src[kMaxSources];
  int n_src;
  int H, W;            // resolution of the inputs (and of the accumulator grid)
  const float* bias;   // [N]
  float inv_scale;     // accumulators are multiplied by this (weights are pre-scaled by a power of two) before the bias
  int relu;            // must be 1: every convolution on this path is followed by ReLU (the epilogue applies it always)
  uint16_t* out;       // hi (or only) output tensor
  uint16_t* out_lo;    // residual output tensor (split precision), else null
  int out_planes_total, out_plane0;
  int64_t o_img_stride, o_plane_stride;   // the same two strides of the OUTPUT tensor (in its own geometry)
  int upsample;        // 0: store at the position; 1: 2 x 2 replicated store into a tensor at twice the resolution;
                       // 2: 2 x 1 (columns only) into a half-row tensor (TcSource::in_up)
  int out_ring;        // like TcSource::ring, for the output tensor
  // Mask-head fusion (N = 32, the last ResBlock of the mask path): when head_w is set the epilogue does not store
  // the activations but their contraction with conv_flatten's weights for the position's mel row,
  //   head_out[b][y][x][k] = sum_c act[c] * head_w[y][c][k]   (k < 4; pytorch_neural_nets.py:133-134,188-190),
  // 16 bytes per position instead of 128: with out == null the 32-channel tensor is never written (it is kept only
  // when the spec head will read it), and the head kernel just sums 128 rows per frame — in a fixed order, so
  // results stay reproducible bit for bit and do not depend on the batch split.
  // Scalar residual (conv1_1: the block input has ONE channel, so its 1x1 residual branch is res_w[c] * x per
  // position — an FMA in the epilogue instead of two MMA stages that carry 15/16 zeros): res_x is the float32 input
  // image [B][H][W] (the mel features), res_w the folded residual weights [N]; null = no scalar residual.
  const float* res_x;
  const float* res_w;
  const float* head_w;   // [128 mel][32][4] float32 or null
  float* head_out;       // [B][128][256][4] float32
  int units_per_image, total_units;
  // Packed units (flat geometry, small images in plane-major tensors): the padded images of the batch are ONE stream
  // of positions g = b * (H+2)(W+2) + q — contiguous in every plane of a plane-major tensor — and a unit is MT*128
  // consecutive positions of the stream, wherever images begin and end: an 8 x 16 image has 180 padded positions, so
  // one 256-position unit per image computes 30 % air (16 x 32: 612 positions in three units, 20 %).  Taps that leave
  // an image read its neighbour's zero border rows or feed border positions only, exactly as inside one image.  The
  // launcher sets units_per_image = total_units (every unit belongs to "image 0", positions run on), so the producer
  // and the MMA issuers need nothing new; the epilogue finds each position's image.  `batch` = images in the stream.
  // (With image-major tensors a staged run needs one bulk copy per image it touches: measured 20-86 % SLOWER on these
  // launches — small bulk copies are what a stage can least afford — hence the layout.)
  int packed, batch;
  // Split-K sub-accumulation (split precision, G = 1).  The tensor core rounds its fp32 accumulator toward zero after
  // every MMA and aligns the 16 products of an MMA to the accumulator's exponent with only ~2 guard bits: a chain of
  // n accumulating MMAs ends ~2.3 n ulp short of the exact sum, a bias that grows linearly with the chain
  // (tools/acc_chain_test.cu: 7e-7 relative after 36 MMAs, 2.8e-6 after 144) and was the whole gap between f16x3 and
  // float32 logits.  A unit's K-chunks are therefore cut into n_sub consecutive groups ("sub-items": chunks
  // [j n / n_sub, (j+1) n / n_sub) of EVERY source); each group accumulates from zero in its own TMEM buffer (the two
  // buffers alternate between sub-items exactly as they did between units) and the epilogue warps add the groups in
  // registers, in float32 round-to-nearest, before bias / ReLU.  n_sub = 1 is the plain single chain.
  int n_sub;
  int pair_store;      // up-sampling epilogue: 1 = one 32-byte store per position pair (default), 0 = two 16-byte stores
  // Row-aligned units with MaxPool2d(2) folded into the epilogue (kernel template parameter Rows; dual layout, G = 1,
  // W = 128 * MT / 2).  A unit is the image-row pair (2Y+1, 2Y+2) of the padded tensor: MT / 2 tiles of 128 interior
  // positions per row, so border positions are never computed (the tensor's zero border stays as allocated) and the
  // 2 x 2 windows of the pool lie inside the unit — vertically in the same epilogue thread (its two tiles), horizontally
  // in neighbouring lanes.  The epilogue stores the activations as usual and, from the float32 values it still holds,
  // the pooled tensor [B][N/8][H/2+2][W/2+2][8] (hi / lo): pool_planar's read of the full-resolution tensor and its
  // launch disappear.  max commutes with the monotonic hi/lo split, so the pooled operands are bit-identical to
  // pool_planar's.
  int rows;
  // Resident weights: the launch's packed weights (wres bytes, at the start of shared memory) are loaded ONCE per CTA
  // and the ring slots carry activations only.  Re-staging them with every K-chunk of every unit was 29 % (N = 32 at
  // 128 x 256) to 62 % (N = 64 at 64 x 128) of the L2 -> shared-memory traffic of these launches and kept the ring at
  // four slots; without them a slot is 16-33 KB and the ring 4-8 deep.  0: weights travel with every chunk.
  int wres;
  uint16_t* pool_out;
  uint16_t* pool_lo;
  int64_t pool_img_stride, pool_plane_stride;     // of the pooled tensor, in positions
  int stages;          // smem ring depth (<= kMaxStages)
  int cps;             // K-chunks a stage of a 1x1 source carries (>= 1; see the producer)
  int stage_stride;    // bytes per smem ring slot (>= the 3x3 stage; larger when that buys more 1x1 chunks per stage)
  int* err;            // [0] pipeline time-out code, [1] scratch of the tuning hooks, [2] fp16 range overflow seen
  long long* prof;     // optional [gridDim.x][8] cycle counters (role wait/busy times), may be null
  int debug;           // tuning experiments only: 1 = producer skips the copies, 2 = epilogue skips the stores
};

// One launch: a single convolution (n_phase = 1) or both convolutions of a ResBlock (n_phase = 2), c[1] reading the
// tensor c[0] writes.  Geometry (H, W, units, stages, err, prof, debug) is shared and taken from c[0].
//
// Fused schedule.  The work items of the launch are the units of c[0] and c[1] interleaved one to one, c[1] running
// `lag` units behind:   c0:0 .. c0:lag-1 | c0:lag c1:0 | c0:lag+1 c1:1 | ... | c1:T-lag .. c1:T-1,   handed to the
// persistent CTAs round-robin.  A c[1] unit reads positions produced by the c[0] units v-1, v, v+1 of its image
// (the halo is shorter than a unit); every epilogue warp of a c[0] unit bumps flags[unit] after its stores, and the
// producer of the c[1] unit waits for 8 arrivals on each of the (up to) three flags before its first copy.  With
// lag > 148 the flags it needs were raised about two rounds earlier, so the wait is a formality, the intermediate
// tensor is read back from L2 a few microseconds after it was written instead of from HBM a whole batch later, and
// the launch has one tail instead of two.  Every item depends only on items with a smaller index and CTAs take
// their items in increasing order, so the lowest unfinished item can always run: no deadlock while all CTAs are
// resident (grid <= SM count, one CTA per SM); all waits are bounded and flag p.err instead of hanging.
// One staged K-step of a unit ("stage program", built by the host: launch_conv_npg).  The producer and the MMA
// issuers walk the same flat list instead of nested sub-item / source / chunk loops: the MMA-issuing thread's
// bookkeeping between two bursts of MMAs is a tensor-pipe bubble, so it is one table look-up per stage.
//   bits 0-2 source | 3-8 first chunk | 9-11 chunks in the stage | 12 first stage of an accumulation group
//   | 13 last stage of the group | 14 the source is 3x3 | 15-16 source kind | 17 "rowdup" 3x3 source
//   | 18 chunks from a half-row tensor | 19-21 ordinary chunks ahead of them (1x1 stages)
//
// "rowdup" (row-aligned launches only): the source planes hold a nearest-neighbour 2x up-sampled image, i.e. image rows
// 2k and 2k+1 are identical.  For an output row at the TOP of such a pair the taps dy = 0 and dy = +1 read the same
// data, for one at the BOTTOM dy = -1 and dy = 0 do, so each tile needs SIX tap MMAs instead of nine with weights
// merged on the host: top rows w(-1,.), w(0,.)+w(+1,.) on input rows (y-1, y); bottom rows w(-1,.)+w(0,.), w(+1,.) on
// (y, y+1).  The stage carries the 12 tap matrices [top 6 | bottom 6]; a tile's row parity picks its half.
constexpr int kMaxProg = 96;
__host__ __device__ constexpr uint32_t prog_entry(int src, int kc, int n, bool first, bool last, bool taps9, int kind,
                                                  bool dup = false, bool half = false, int nfull = 0) {
  // bit 18: the stage has chunks staged from a half-row tensor (TcSource::in_up) — as the tensor stores them, i.e.
  // one staged row per PAIR of image rows; bits 19-21: how many leading chunks of a 1x1 stage are ordinary ones
  return (uint32_t)src | ((uint32_t)kc << 3) | ((uint32_t)n << 9) | ((uint32_t)first << 12) | ((uint32_t)last << 13) |
         ((uint32_t)taps9 << 14) | ((uint32_t)kind << 15) | ((uint32_t)dup << 17) | ((uint32_t)half << 18) |
         ((uint32_t)nfull << 19);
}

struct TcJob {
  TcConv c[2];
  uint32_t prog[2][kMaxProg];   // stage program of a unit of each phase
  // resident weights (phase 0 only): per stage, bits 0-15 = (offset of its first chunk's weights in the resident
  // region) >> 4, bits 16-31 = (bytes between consecutive chunks) >> 4 — fetched a stage ahead like the program word,
  // so that no dependent parameter look-up sits between two bursts of MMAs
  uint32_t prog_b[kMaxProg];
  int prog_len[2];
  int n_phase;
  int layout;          // warp-role layout (see the kernel): 1 = critical roles on the highest warp ids (default)
  int heavy_fence;     // fused launches publish a unit with __threadfence + atomicAdd instead of a release-reduction (A/B)
  int epi;             // epilogue operand fetch: bit 0 = scalar residual preloaded before the accumulator wait (A/B runs)
  int lag;             // units by which c[1] trails c[0]
  int* flags;          // [total_units], zeroed before the launch (fused launches only)
  int flags_cap;
  // Ring mode (ring > 0): the intermediate tensor holds only `ring` images, image b in slot b % ring, so that it is
  // rewritten while its lines are still dirty in L2 and never travels to HBM.  The c[0] unit (b, u) may overwrite
  // what the c[1] units (b - ring, u-1 .. u+1) still read, so those bump flags2[unit] when they are done (their
  // copies completed long before their epilogue runs) and the c[0] producer waits for them first.  The host only
  // enables the ring when ring * units_per_image > lag + 2, i.e. when those units precede the waiting one in the
  // item order (no deadlock), which also makes the wait a formality.
  int ring;
  int ring_request;    // host-side wish (images); the launcher derives `ring` from it
  int* flags2;         // [total_units], zeroed before the launch
};

// item -> (phase, unit) of the interleaved schedule above (T units per phase, D = min(lag, T))
__device__ __forceinline__ void decode_item(int i, int T, int D, int n_phase, int& phase, int& unit) {
  if (n_phase == 1 || i < D) { phase = 0; unit = i; return; }
  const int j = i - D;
  if (j < 2 * (T - D)) { phase = j & 1; unit = (j >> 1) + (phase ? 0 : D); return; }
  phase = 1;
  unit = (T - D) + (j - 2 * (T - D));
}

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Bounded wait for `want` arrivals on a unit flag (all lanes probe the same word: one transaction).
__device__ __noinline__ bool flag_wait(const int* f, int want, int* err, int code) {
#pragma unroll 1
  for (uint32_t i = 0; i < kSpinLimit; ++i) {
    if (ld_acquire(f) >= want) return true;
    __nanosleep(64);
  }
  atomicExch(err, code);
  return false;
}

// Bounded wait for `want` arrivals on each of the flags f[-1], f[0], f[+1] (those with `lo` / `hi` set): lanes 0-2 of
// the (converged) warp poll one flag each, so the three acquire loads are one L2 round trip, not three in a row.
__device__ __noinline__ bool flag_wait3(const int* f, bool lo, bool hi, int want, int* err, int code) {
  const int lane = threadIdx.x & 31;
  const bool mine = lane == 1 || (lane == 0 && lo) || (lane == 2 && hi);
#pragma unroll 1
  for (uint32_t i = 0; i < kSpinLimit; ++i) {
    const bool ok = !mine || ld_acquire(f + lane - 1) >= want;
    if (__all_sync(0xffffffffu, ok)) return true;
    __nanosleep(64);
  }
  if (lane == 0) atomicExch(err, code);
  return false;
}

// 128-position tiles per work unit: one 256-column TMEM accumulator buffer holds MT tiles of N columns.
// (TS = TMEM columns per tile: N, or 2N in the dual layout).
template <int N, bool Dual>
struct TilesPerUnit {
  static constexpr int TS = Dual ? 2 * N : N;
  static constexpr int value = (TS == 96) ? 2 : kAccCols / TS;
};

__host__ __device__ inline size_t stage_bytes(int N, int W, int tiles, bool dual, int wtaps = 9) {
  return ((size_t)tiles * 128 + 2 * (size_t)(W + 3)) * 32 + (dual ? 2 : 1) * (size_t)wtaps * (size_t)N * 32;
}

// The MMAs of one K-chunk for one group of MT tiles: straight-line, every descriptor is (loop-invariant high
// word, base + compile-time step), so the MMAs go out back to back from uniform registers (a dependent
// uniform-ALU chain per MMA costs ~90 cycles, twice the 32 + N/4 cycles the shared-memory operand fetch allows;
// tools/umma_bench.cu).  BN = weight rows per tap and K-half in the stage (N, or 2N for dual weights).
template <int MT, int TS, int BN>
__device__ __forceinline__ void issue_group(uint32_t d0, uint32_t a_lo0, uint32_t b_lo0, uint32_t idesc, int taps,
                                            const int (&tap_off)[9], uint32_t accumulate, const uint32_t tile_step = 128u) {
  const uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = a_hi;
  if (taps == 9) {
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo0 + (uint32_t)(tap * BN * 2));
      const uint32_t a_lo_tap = a_lo0 + (uint32_t)tap_off[tap];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint64_t da = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo_tap + (uint32_t)mt * tile_step);
        tc_mma(d0 + (uint32_t)(mt * TS), da, db, idesc, tap == 0 ? accumulate : 1u);
      }
    }
  } else {
    const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)b_lo0;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const uint64_t da = ((uint64_t)a_hi << 32) | (uint64_t)(a_lo0 + (uint32_t)mt * tile_step);
      tc_mma(d0 + (uint32_t)(mt * TS), da, db, idesc, accumulate);
    }
  }
}

// The six tap MMAs per tile of a "rowdup" source (see prog_entry).  A tile of row parity `par` reads input rows
// (y - 1 + par, y + par) — the first six entries of tap_off shifted by par rows — with tap matrices [6 par, 6 par + 6)
// of the stage.  kParFromTile: the warp's tiles alternate parity (one tile per image row, two tiles per warp);
// otherwise all of its tiles lie in rows of parity `par_warp`.
// kTap0 = 0: the staged rows are image rows (offsets tap_off[0..5]: rows -1 and 0 of the tile's own row + par);
// kTap0 = 3: the stage holds half rows (one per pair of image rows) and a_lo0 points at the first of the tile's two
// half rows (offsets tap_off[3..8]: rows 0 and +1), row_step = 0.
template <int MT, int TS, int BN, bool kParFromTile, int kTap0 = 0>
__device__ __forceinline__ void issue_group_dup(uint32_t d0, uint32_t a_lo0, uint32_t b_lo0, uint32_t idesc,
                                                const int (&tap_off)[9], uint32_t accumulate, const uint32_t tile_step,
                                                const uint32_t row_step, const uint32_t par_warp) {
  const uint32_t a_hi = (128u >> 4) | (1u << 14), b_hi = a_hi;
  if constexpr (!kParFromTile) {
    a_lo0 += par_warp * row_step;
    b_lo0 += par_warp * (uint32_t)(6 * BN * 2);
  }
#pragma unroll
  for (int tap = 0; tap < 6; ++tap) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const uint32_t par = kParFromTile ? (uint32_t)(mt & 1) : 0u;
      const uint64_t db = ((uint64_t)b_hi << 32) | (uint64_t)(b_lo0 + (uint32_t)((par * 6 + tap) * BN * 2));
      const uint64_t da = ((uint64_t)a_hi << 32) |
                          (uint64_t)(a_lo0 + (uint32_t)tap_off[tap + kTap0] + (uint32_t)mt * tile_step + par * row_step);
      tc_mma(d0 + (uint32_t)(mt * TS), da, db, idesc, tap == 0 ? accumulate : 1u);
    }
  }
}

// Sub: the split-K sub-accumulation path (TcConv::n_sub > 1) is compiled in — a separate instantiation, because
// keeping a unit's sums in registers across buffer turns costs the plain single-chain launches 7 % (measured).
// Epi: which epilogue the instantiation carries — 0 the general one (`finalize`), 1 the folded-MaxPool epilogue of
// row-aligned units and nothing else, 2 the general one with the fused mask-head partials (N = 32).  The three used to
// share one kernel; its code (4,300 instructions in the N = 32 row-aligned instantiation) is what the MMA-issuing warps
// compete with for the instruction cache, and every change to one epilogue moved the time of launches that never run
// it by +-10 % (profiles/r2_tuning_experiments.txt, section 17).  3 = the general one with conv1_1's scalar residual
// (TcConv::res_x; the folded-pool epilogue of N = 32 carries it too).
// NoUp: the launch stores no up-sampled output (TcConv::upsample == 0 in every phase): the replicated-store variants
// of `finalize` are not compiled (they are copied 4 N / 32 times into the unrolled epilogue).
template <int N, Prec P, bool Dual, int G, bool Sub = false, int Rows = 0, int Epi = 0, bool NoUp = false>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_kernel(const TcJob job) {
  extern __shared__ __align__(128) unsigned char smem[];
  const TcConv& p = job.c[0];                 // shared geometry; per-item parameters are job.c[phase]
  const int n_phase = job.n_phase;
  const int T = p.total_units, D = job.lag < T ? job.lag : T;
  const int n_items = n_phase * T;
  // Round r of the persistent loop: CTA j takes item r * grid + (j + r) % grid.  The rotation matters for fused
  // launches: items alternate c[0] / c[1], the grid is even, and without it a CTA would see one phase only (the two
  // phases cost differently, so half the SMs would finish early).  Items still increase with r for every CTA.
  // (single-convolution launches need no rotation, and the MMA issuers should not pay an integer division at every
  // unit boundary — whatever they do between two units is a tensor-pipe bubble; job.layout bit 1 keeps it for A/B runs)
  const bool rotate = n_phase == 2 || (job.layout & 2) != 0;
  auto item_of = [rotate](int r) {
    return r * (int)gridDim.x + (rotate ? (int)((blockIdx.x + (unsigned)r) % gridDim.x) : (int)blockIdx.x);
  };
  static_assert(!Dual || PrecTraits<P>::split, "the dual layout belongs to the split precision");
  constexpr int MT = TilesPerUnit<N, Dual>::value;
  constexpr int TS = TilesPerUnit<N, Dual>::TS;
  constexpr bool kSplit = PrecTraits<P>::split;
  constexpr bool kSubAcc = Sub;
  constexpr bool kResX = N == 32 && (Epi == 1 || Epi == 3);      // conv1_1's scalar residual (TcConv::res_x) is compiled in
  static_assert(!Sub || (kSplit && G == 1), "sub-accumulation belongs to the split precision, one group per unit");
  // Rows = tiles of 128 positions per image row (0: flat units of consecutive positions).  A row-aligned unit is
  // kUnitRows = MT / Rows consecutive image rows starting at an even one.
  static_assert(Rows == 0 || ((Rows == 1 || Rows == 2) && Dual && G == 1 && !Sub && N % 32 == 0 && MT % (2 * Rows) == 0),
                "row-aligned units: dual layout, one group, plain chain, an even number of rows");
  constexpr int kUnitRows = Rows ? MT / (Rows ? Rows : 1) : 0;
  constexpr int kRowsExtra = Rows ? 2 * kUnitRows - 2 : 0;   // the border positions between the unit's image rows
  constexpr int kWpartsMax = Dual ? 2 : 1;   // weight rows per tap and K-half staged per chunk, in units of N
  const int dbg = kDebugHooks ? p.debug : 0;
  const int Wp = p.W + 2, Hp = p.H + 2;
  const int HpWp = Hp * Wp;
  const int halo = Wp + 1;
  const int L = G * MT * 128 + 2 * halo + kRowsExtra;   // positions staged per plane
  const uint32_t a_bytes = (uint32_t)L * 32u;           // two planes
  const uint32_t stage_sz = (uint32_t)p.stage_stride;   // >= a_bytes + kWpartsMax * 9 * N * 32 (host-checked)
  const int S = p.stages;
  const int cps = p.cps;
  const uint32_t run1 = (uint32_t)(G * MT * 128 + kRowsExtra) * 16u;      // one plane of a 1x1 source's chunk (no halo)
  const uint32_t w1_off = (uint32_t)cps * 2u * run1;          // weights of a 1x1 stage follow its cps chunk slots
  const uint32_t wres = (uint32_t)p.wres;                    // resident weights ahead of the ring (0: none)
  unsigned char* stage0 = smem + wres;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + (size_t)S * stage_sz);
  // bars: full[kMaxStages] | empty[kMaxStages] | acc_full[2] | acc_empty[2] | resident weights | (pad)
  float* bias_s = reinterpret_cast<float*>(bars + 2 * kMaxStages + 6);       // [2][N] bias, then [2][N] scalar-residual weights
  float* zero_s = bias_s + 4 * N;                                            // [N] zeros: the "bias" of border positions
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(zero_s + N);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Warp roles.  The warp scheduler of an SM sub-partition (warp id % 4) prefers its highest-numbered eligible warp, so
  // the latency-critical roles sit on top of their sub-partitions: producer = warp 8 (above epilogue warps 0 and 4),
  // MMA issuers = warps 9 and 10 (above 1, 5 and 2, 6); epilogue = warps 0-7.  (layout 0, kept for A/B runs:
  // producer 4, MMA 5 and 10, epilogue 0-3 and 6-9 — the MMA warp 5 then loses its issue slots to epilogue warp 9
  // whenever that one has an instruction ready.)
  const bool top_roles = job.layout != 0;
  const int w_prod = top_roles ? 8 : 4, w_mma0 = top_roles ? 9 : 5, w_mma1 = 10;
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + kMaxStages);
  const uint32_t accf0 = smem_u32(bars + 2 * kMaxStages), acce0 = smem_u32(bars + 2 * kMaxStages + 2);
  const uint32_t wbar = smem_u32(bars + 2 * kMaxStages + 4);

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(empty0 + 8 * s, 2);      // both MMA warps commit their share of the stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(accf0 + 8 * i, G == 2 ? 1 : 2);   // G = 2: the buffer's owner; G = 1: both MMA warps
      mbar_init(acce0 + 8 * i, 8);       // one arrival per epilogue warp
    }
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < n_phase * N; i += kTcThreads) {
    bias_s[i] = job.c[i / N].bias[i % N];
    bias_s[2 * N + i] = job.c[i / N].res_w ? job.c[i / N].res_w[i % N] : 0.f;
  }
  for (int i = threadIdx.x; i < N; i += kTcThreads) zero_s[i] = 0.f;
  if (warp == w_prod) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == w_prod) {
    // ===================================================================== producer (warp-uniform)
    int it = 0;
    bool ok = true;
    long long w_empty = 0;
    if (wres && elect_one()) {
      // the launch's weights, once: one bulk copy per kind-1 source (its chunks are contiguous in the dual layout)
      mbar_expect_tx(wbar, wres);
      for (int i = 0; i < p.n_src; ++i)
        if (p.src[i].wres_bytes > 0)
          bulk_g2s(smem_u32(smem) + (uint32_t)p.src[i].wres_off, p.src[i].w, (uint32_t)p.src[i].wres_bytes, wbar);
    }
    __syncwarp();
    for (int r = 0, item; (item = item_of(r)) < n_items && ok && !(dbg & 4); ++r) {
      int phase, u;
      decode_item(item, T, D, n_phase, phase, u);
      const TcConv& c = job.c[phase];
      const int b = u / p.units_per_image;
      const int lu = u - b * p.units_per_image;
      const int lo = Rows ? lu * kUnitRows * Wp : lu * G * MT * 128;   // first staged position (= q0 - halo; Rows: start of a padded row)
      if (phase == 0 && job.ring > 0 && b >= job.ring) {
        const int v0 = u - job.ring * p.units_per_image;       // same local unit, `ring` images earlier
        ok = flag_wait3(job.flags2 + v0, lu > 0, lu < p.units_per_image - 1, 8, p.err, 6);
        if (!ok) break;
      }
      if (phase == 1) {
        // the units of c[0] whose output this unit reads must be complete (see TcJob), and their generic-proxy
        // stores visible to the async proxy that performs the bulk copies
        // (three acquire loads at once, one per lane: one L2 round trip instead of three in a row, during which this
        // warp stages nothing; reading them one item AHEAD was slower still — an acquire load holds back the bulk
        // copies issued after it)
        ok = flag_wait3(job.flags + u, lu > 0, lu < p.units_per_image - 1, 8, p.err, 5);
        if (!ok) break;
        asm volatile("fence.proxy.async.global;" ::: "memory");
      }
      const int n_prog = job.prog_len[phase];
      for (int pi = 0; pi < n_prog && ok; ++pi, ++it) {
        const uint32_t e = job.prog[phase][pi];
        const TcSource& src = c.src[e & 7u];
        const int kc = (int)((e >> 3) & 63u), n = (int)((e >> 9) & 7u);
        const bool taps9 = (e >> 14) & 1u;
        const bool dup = (e >> 17) & 1u;
        const uint32_t w_bytes =
            wres ? 0u : (uint32_t)((Dual && ((e >> 15) & 3u) == 1u ? 2 : 1) * (dup ? 12 : (taps9 ? 9 : 1))) * N * 32u;
        const int st = it % S;
        const uint32_t ph = (uint32_t)(it / S) & 1u;
        ok = mbar_wait_t<true>(empty0 + 8 * st, ph ^ 1u, p.err, 1, w_empty);
        if (!ok) break;
        const uint32_t dst = smem_u32(stage0 + (size_t)st * stage_sz);
        const int bs = src.ring ? b % src.ring : b;
        const int64_t pstride = src.plane_stride * 8;      // 16-bit elements between consecutive planes
        const uint16_t* plane = src.in + ((int64_t)bs * src.img_stride + (int64_t)(src.plane0 + 2 * kc) * src.plane_stride + lo) * 8;
        const int Hh = (p.H >> 1) + 2;       // rows of a half-row tensor
        if (dbg & 1) {
          if (elect_one()) mbar_arrive(full0 + 8 * st);
        } else if (Rows != 0 && src.in_up != nullptr && kc + n > src.up_chunk0) {
          // the stage has chunks that live in the half-row tensor: for those, one bulk copy per staged image row and
          // plane (a 3x3 stage is one chunk; the chunks of a 1x1 stage are routed one by one)
          // Staged as the tensor stores them: a 3x3 chunk as the kUnitRows / 2 + 2 half rows that cover the unit's rows
          // and its halo (ONE copy per plane, 25-33 % fewer bytes than image rows), a 1x1 chunk as the unit's own
          // kUnitRows / 2 half rows; the MMA issuers map image rows to half rows.
          if (elect_one()) {
            const int hr0 = (kUnitRows / 2) * lu;        // tensor row of the half row above the unit's first row pair
            if (taps9) {
              const uint32_t run = (uint32_t)((kUnitRows / 2 + 2) * Wp) * 16u;
              const uint16_t* base =
                  src.in_up + (((int64_t)bs * src.up_planes_total + 2 * (kc - src.up_chunk0)) * Hh + hr0) * Wp * 8;
              mbar_expect_tx(full0 + 8 * st, 2u * run + w_bytes);
              bulk_g2s(dst, base, run, full0 + 8 * st);
              bulk_g2s(dst + run, base + (int64_t)Hh * Wp * 8, run, full0 + 8 * st);
              if (w_bytes) bulk_g2s(dst + a_bytes, src.w + (int64_t)kc * src.w_stride, w_bytes, full0 + 8 * st);
            } else {
              const uint32_t run1h = (uint32_t)((kUnitRows / 2 - 1) * Wp + p.W) * 16u;
              int n_half = kc + n - src.up_chunk0;
              if (n_half > n) n_half = n;
              mbar_expect_tx(full0 + 8 * st, (uint32_t)(n - n_half) * 2u * run1 + (uint32_t)n_half * 2u * run1h + (uint32_t)n * w_bytes);
#pragma unroll 1
              for (int j = 0; j < n; ++j) {
                if (kc + j >= src.up_chunk0) {
                  const uint16_t* pj = src.in_up + ((((int64_t)bs * src.up_planes_total + 2 * (kc + j - src.up_chunk0)) * Hh +
                                                     hr0 + 1) * Wp + 1) * 8;
                  bulk_g2s(dst + (uint32_t)(2 * j) * run1, pj, run1h, full0 + 8 * st);
                  bulk_g2s(dst + (uint32_t)(2 * j + 1) * run1, pj + (int64_t)Hh * Wp * 8, run1h, full0 + 8 * st);
                } else {
                  const uint16_t* pj = plane + (int64_t)2 * j * pstride + (int64_t)halo * 8;
                  bulk_g2s(dst + (uint32_t)(2 * j) * run1, pj, run1, full0 + 8 * st);
                  bulk_g2s(dst + (uint32_t)(2 * j + 1) * run1, pj + pstride, run1, full0 + 8 * st);
                }
                if (w_bytes)
                  bulk_g2s(dst + w1_off + (uint32_t)j * w_bytes, src.w + (int64_t)(kc + j) * src.w_stride, w_bytes,
                           full0 + 8 * st);
              }
            }
          }
        } else if (!taps9) {
          // A 1x1 source has no halo and a ninth of the weights, so a stage-sized slot takes up to `cps` of its
          // K-chunks: [chunk][plane][run1] activations, then [chunk] weights at w1_off.  (One chunk per stage left
          // these sources bound by the per-stage hand-over and by load latency: 4 MMAs per 18 KB stage.)
          if (elect_one()) {
            mbar_expect_tx(full0 + 8 * st, (uint32_t)n * (2u * run1 + w_bytes));
            for (int j = 0; j < n; ++j) {
              const uint16_t* pj = plane + (int64_t)2 * j * pstride + (int64_t)halo * 8;    // centre tap only
              bulk_g2s(dst + (uint32_t)(2 * j) * run1, pj, run1, full0 + 8 * st);
              bulk_g2s(dst + (uint32_t)(2 * j + 1) * run1, pj + pstride, run1, full0 + 8 * st);
              if (w_bytes)
                bulk_g2s(dst + w1_off + (uint32_t)j * w_bytes, src.w + (int64_t)(kc + j) * src.w_stride, w_bytes,
                         full0 + 8 * st);
            }
          }
        } else if (elect_one()) {
          const uint32_t run = (uint32_t)L * 16u;
          mbar_expect_tx(full0 + 8 * st, 2u * run + w_bytes);
          bulk_g2s(dst, plane, run, full0 + 8 * st);
          bulk_g2s(dst + run, plane + pstride, run, full0 + 8 * st);
          if (w_bytes) bulk_g2s(dst + a_bytes, src.w + (int64_t)kc * src.w_stride, w_bytes, full0 + 8 * st);
        }
        __syncwarp();
      }
    }
    if (p.prof && lane == 0) p.prof[blockIdx.x * 8 + 0] = w_empty;
  } else if (warp == w_mma0 || warp == w_mma1) {
    // ===================================================================== MMA issuers (warp-uniform)
    const int me = (warp == w_mma0) ? 0 : 1;        // which half of the accumulator tiles this warp issues for
    constexpr int MTW = (G == 2) ? MT : MT / 2;     // tiles per warp and stage
    static_assert(G == 2 || MT % 2 == 0, "tiles must split evenly between the two MMA warps");
    constexpr uint32_t idesc_n = instr_desc(N, PrecTraits<P>::fmt);
    constexpr uint32_t idesc_2n = instr_desc(2 * N, PrecTraits<P>::fmt);
    int tap_off[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) tap_off[t] = (t / 3 - 1) * Wp + (t % 3 - 1);
    // descriptor low word = (LBO >> 4) << 16 | (address >> 4)
    const uint32_t a_lo_base = ((uint32_t)L & 0x3FFFu) << 16;            // LBO = L * 16 bytes
    const uint32_t a1_lo_base = ((run1 >> 4) & 0x3FFFu) << 16;           // 1x1 stages: LBO = run1
    int k = 0;
    int kk = 0;                    // TMEM buffer turns taken so far: one per unit, or one per sub-item (kSubAcc)
    bool ok = true;
    long long w_acce = 0, w_full = 0;
    const bool timing = p.prof != nullptr;
    const long long t_begin = clock64();
    // The tensor pipe accepts only a couple of MMAs ahead of execution, so every cycle this loop spends between two
    // bursts of MMAs is a pipe bubble: the ring position is carried incrementally (no division), the leader lane is
    // elected once, waits probe inline, and a stage costs one commit.
    const uint32_t leader = elect_one();
    const uint32_t stage_base = smem_u32(stage0);
    const uint32_t res_base = smem_u32(smem);
    if (wres) ok = mbar_wait(wbar, 0u, p.err, 7);
    tc_fence_after();
    int st = 0;
    uint32_t ph = 0, a0 = stage_base;
    bool ready = false;            // the full barrier of the stage about to be consumed was already seen complete
    uint32_t e_next = job.prog[0][0];
    uint32_t pb_next = (Rows != 0) ? job.prog_b[0] : 0u;
    const int n_prog0 = job.prog_len[0];
    for (int item; (item = item_of(k)) < n_items && ok; ++k) {
      int phase, u;
      decode_item(item, T, D, n_phase, phase, u);
      // first position of my tiles (Rows: warp `me` issues for the tiles of image row `me` of the unit's pair)
      const uint32_t a_tile0 = Rows == 2 ? (uint32_t)(me * Wp) : Rows == 1 ? (uint32_t)(me * MTW * Wp)
                                                                        : (uint32_t)((G == 1 ? me * MTW : me * MT) * 128);
      const uint32_t tile_step = Rows == 1 ? (uint32_t)Wp : 128u;      // between this warp's consecutive tiles
      // the same for chunks staged as half rows: (first image row of the warp's tiles in the unit) >> 1 staged rows down
      const uint32_t a_tile0_half = Rows == 1 ? (uint32_t)(((me * MTW) >> 1) * Wp) : 0u;
      const int n_prog = (n_phase == 2) ? job.prog_len[phase] : n_prog0;
      int buf = 0;
      uint32_t d_unit = 0u;
      bool accumulate_next = false;
      // (single-phase launches: the first entry of the next unit's program was fetched during this unit's last stage;
      // job.layout bit 2 reloads it here, for A/B runs)
      if (n_phase == 2 || (job.layout & 4)) {
        e_next = job.prog[phase][0];
        if constexpr (Rows != 0) pb_next = job.prog_b[0];
      }
      for (int pi = 0; pi < n_prog && ok; ++pi) {
        const uint32_t e = e_next;
        e_next = job.prog[phase][pi + 1 < n_prog ? pi + 1 : 0];      // fetched a stage ahead (constant-bank latency)
        uint32_t pb = 0u;
        if constexpr (Rows != 0) {
          pb = pb_next;
          pb_next = job.prog_b[pi + 1 < n_prog ? pi + 1 : 0];
        }
        const int n = (int)((e >> 9) & 7u);
        const bool first = (e >> 12) & 1u, last = (e >> 13) & 1u, taps9 = (e >> 14) & 1u, dup = (e >> 17) & 1u;
        const bool half = (e >> 18) & 1u;                 // chunks staged as half rows (TcSource::in_up)
        const int nfull = (int)((e >> 19) & 7u);
        const uint32_t kind = (e >> 15) & 3u;
        // per-source MMA shape: the dual product writes 2N columns, a correction-only source the upper N
        const bool dual_src = Dual && kind == 1u;
        const uint32_t idesc = dual_src ? idesc_2n : idesc_n;
        const uint32_t col0 = (Dual && kind == 2u) ? (uint32_t)N : 0u;
        // resident weights (row-aligned dual launches): every B operand is a slice of a dual matrix, whose K-half blocks
        // have 2N rows whatever the MMA's N
        const bool wide = dual_src || (Rows != 0 && wres != 0);
        const uint32_t b_lo_base = ((uint32_t)(wide ? 2 * N : N) & 0x3FFFu) << 16;   // LBO = rows * 16 bytes
        // resident: this stage's first chunk (>> 4), chunk stride (>> 4)
        const uint32_t b_res = (res_base >> 4) + (pb & 0xffffu), b_res_step = pb >> 16;
        if (!ready && !(dbg & 4)) ok = mbar_wait_fast(full0 + 8 * st, ph, p.err, 2, timing, w_full);
        if (!ok) break;
        if (first) {
          // a new accumulation group takes the next TMEM buffer turn (G = 1: buffer kk & 1, tiles split between the
          // warps; G = 2: warp `me` owns buffer `me`); the epilogue must have drained it
          buf = (G == 1) ? (kk & 1) : me;
          const uint32_t e_parity = (G == 1) ? ((((uint32_t)kk >> 1) & 1u) ^ 1u) : (((uint32_t)kk & 1u) ^ 1u);
          d_unit = tmem_base + (uint32_t)(buf * kAccCols) + (G == 1 ? (uint32_t)(me * MTW * TS) : 0u);
          ++kk;
          accumulate_next = false;
          ok = mbar_wait_fast(acce0 + 8 * buf, e_parity, p.err, 4, timing, w_acce);
          if (!ok) break;
        }
        tc_fence_after();
        // Look ahead: a non-blocking probe of the NEXT stage's full barrier goes out before this stage's MMAs and is
        // read after them, so its latency (~150 cycles) runs under the MMAs instead of between two bursts of them.
        const int st_n = (st + 1 == S) ? 0 : st + 1;
        const uint32_t ph_n = (st + 1 == S) ? (ph ^ 1u) : ph;
        const bool probe = !(dbg & 4) && mbar_test_wait(full0 + 8 * st_n, ph_n);
        if (leader) {
          if (dbg & 8) { if (a0 == 0xdeadbeefu) p.err[1] = (int)(d_unit + idesc); }   // issue nothing
          else if (!taps9) {
            // compact 1x1 stage: chunk j at a0 + 2 j run1 (plane stride run1), its weights at a0 + w1_off + j w_bytes
            const uint32_t w_bytes16 = (Rows != 0 && wres) ? b_res_step : (uint32_t)((dual_src ? 2 : 1) * N * 2);   // 1x1 chunk weights >> 4
            const uint32_t a1 = (a1_lo_base | (a0 >> 4)) + a_tile0;
            const uint32_t b1 = b_lo_base | ((Rows != 0 && wres) ? b_res : ((a0 + w1_off) >> 4));
            for (int j = 0; j < n; ++j) {
              const uint32_t accumulate = (accumulate_next || j > 0) ? 1u : 0u;
              uint32_t aj = a1 + (uint32_t)j * (2u * run1 >> 4);
              const uint32_t bj = b1 + (uint32_t)j * w_bytes16;
              uint32_t step_j = tile_step;
              if constexpr (Rows != 0) {
                if (half && j >= nfull) {
                  // half rows: the tile in image row r of the unit reads staged row r >> 1 (both tiles of a warp that
                  // issues for two consecutive rows read the same one)
                  aj = aj - a_tile0 + a_tile0_half;
                  if (Rows == 1 && MTW == 2) step_j = 0u;
                }
              }
              if (dual_src) issue_group<MTW, TS, 2 * N>(d_unit + col0, aj, bj, idesc, 1, tap_off, accumulate, step_j);
              else issue_group<MTW, TS, N>(d_unit + col0, aj, bj, idesc, 1, tap_off, accumulate, step_j);
            }
          } else {
            const uint32_t accumulate = accumulate_next ? 1u : 0u;
            const uint32_t a_lo0 = (a_lo_base | ((a0 >> 4) + (uint32_t)halo)) + a_tile0;   // centre tap, my first tile
            const uint32_t b_lo0 = b_lo_base | ((Rows != 0 && wres) ? b_res : ((a0 + a_bytes) >> 4));
            if constexpr (Rows != 0) {
              // (the tap matrices of a B operand are 2 x rows-of-its-K-half-block x 16 bytes apart: template BN)
              if (dup) {
                // a warp's tiles share a row parity (`me`) unless it issues for two consecutive rows
                constexpr bool kParFromTile = (Rows == 1 && MTW == 2);
                static_assert(Rows == 2 || MTW <= 2, "rowdup: at most two rows per MMA warp");
                if (half) {
                  // staged: kUnitRows / 2 + 2 half rows per plane; the tile in image row r (parity par) reads staged
                  // rows (r >> 1) + par and the next: row `me` for both of a warp's tiles in one row, me + tile otherwise
                  const uint32_t ah = ((((uint32_t)((kUnitRows / 2 + 2) * Wp)) & 0x3FFFu) << 16 | ((a0 >> 4) + 1u)) + (uint32_t)(me * Wp);
                  if (wide) issue_group_dup<MTW, TS, 2 * N, kParFromTile, 3>(d_unit + col0, ah, b_lo0, idesc, tap_off, accumulate, tile_step, 0u, (uint32_t)me);
                  else issue_group_dup<MTW, TS, N, kParFromTile, 3>(d_unit + col0, ah, b_lo0, idesc, tap_off, accumulate, tile_step, 0u, (uint32_t)me);
                } else if (wide) issue_group_dup<MTW, TS, 2 * N, kParFromTile>(d_unit + col0, a_lo0, b_lo0, idesc, tap_off, accumulate, tile_step, (uint32_t)Wp, (uint32_t)me);
                else issue_group_dup<MTW, TS, N, kParFromTile>(d_unit + col0, a_lo0, b_lo0, idesc, tap_off, accumulate, tile_step, (uint32_t)Wp, (uint32_t)me);
              } else if (wide) issue_group<MTW, TS, 2 * N>(d_unit + col0, a_lo0, b_lo0, idesc, 9, tap_off, accumulate, tile_step);
              else issue_group<MTW, TS, N>(d_unit + col0, a_lo0, b_lo0, idesc, 9, tap_off, accumulate, tile_step);
            } else if (dual_src) issue_group<MTW, TS, 2 * N>(d_unit + col0, a_lo0, b_lo0, idesc, 9, tap_off, accumulate);
            else issue_group<MTW, TS, N>(d_unit + col0, a_lo0, b_lo0, idesc, 9, tap_off, accumulate);
          }
          if (last) tc_commit(accf0 + 8 * buf);         // my tiles of this accumulation group are complete
          if (!(dbg & 4)) tc_commit(empty0 + 8 * st);   // my reads of the stage retire with these MMAs
        }
        __syncwarp();
        accumulate_next = true;
        a0 += stage_sz;
        st = st_n;
        ph = ph_n;
        if (st == 0) a0 = stage_base;
        // the early probe usually says yes (the producer runs a ring ahead); otherwise ask again, suspending
        ready = probe || (!(dbg & 4) && mbar_try_wait(full0 + 8 * st, ph));
      }
    }
    __syncwarp();
    if (p.prof && lane == 0 && me == 0) {
      p.prof[blockIdx.x * 8 + 1] = w_acce;
      p.prof[blockIdx.x * 8 + 2] = w_full;
      p.prof[blockIdx.x * 8 + 3] = clock64() - t_begin;
      p.prof[blockIdx.x * 8 + 7] = k;
    }
  } else {
    // ===================================================================== epilogue (warps 0-3 and 6-9)
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int tile_par = top_roles ? (warp >> 2) : (warp >= 6 ? 1 : 0);    // the two warps of a quadrant take alternate tiles
    const int Wp2 = 2 * p.W + 2;
    int k = 0;
    int kk = 0;                    // TMEM buffer turns taken so far (see the MMA warps)
    bool ok = true;
    bool out_of_range = false;     // fp16 operand modes: an activation left the fp16 range (it was saturated)
    long long w_accf = 0;
    const long long t_begin = clock64();
    constexpr int kMyTiles = kSubAcc ? MT / 2 : 1;     // tiles of a unit this warp owns (sub-accumulation keeps them all)
    for (int item; (item = item_of(k)) < n_items && ok; ++k) {
      int phase, u;
      decode_item(item, T, D, n_phase, phase, u);
      const TcConv& c = job.c[phase];
      const float inv_scale = c.inv_scale;
      const float* bias_p = bias_s + phase * N;
      const int64_t out_plane_stride = c.o_plane_stride * 8;      // 16-bit elements between planes of the output tensor
      const int b = u / p.units_per_image;
      const int bo = c.out_ring ? b % c.out_ring : b;
      const int64_t img_off = ((int64_t)bo * c.o_img_stride + (int64_t)c.out_plane0 * c.o_plane_stride) * 8;
      uint32_t hmax = 0u;          // fp16 operand modes: running packed maximum of the stored values (range_track)
      float vmax = 0.f;            // ... and the float maximum of the folded-pool epilogue

      // Everything after the accumulators of one tile's 32-column block [n0, n0 + 32) are final: undo the weight
      // scale, bias, scalar residual, ReLU, (mask-head partials,) 16-bit pack, stores.
      // The scalar residual input of a tile's position (conv1_1: mel; 0 outside the image or without one) is fetched
      // BEFORE the wait for the accumulators: the CTA's shared memory takes the whole L1 carve-out, so the load is an
      // L2 round trip, which otherwise sits between the accumulator load and the first FMA (-7 % on conv1_1.c2).
      // (row, column) of the position travel with it: one integer division per tile in flat units, none in row-aligned
      // ones, where the caller knows the row (y_known >= 0)
      // Packed units: `pos` runs over the stream of the batch's images; the tile carries its image and the position in it.
      struct TilePre { float rx; int y, x, b, q; };
      const bool packed = Rows == 0 && p.packed != 0;
      auto preload = [&](const int pos, const int y_known = -1) {
        TilePre t;
        t.b = b;
        t.q = pos;
        if (packed) {
          t.b = pos / HpWp;
          t.q = pos - t.b * HpWp;
        }
        const int y = y_known >= 0 ? y_known : t.q / Wp, x = t.q - y * Wp;
        t.y = (packed && t.b >= p.batch) ? 0 : y;      // past the last image: a border position (stores nothing)
        t.x = x;
        const bool interior = (t.y >= 1) && (y <= p.H) && (x >= 1) && (x <= p.W);
        t.rx = ((kResX && c.res_x != nullptr) && interior && (job.epi & 1))
                   ? __ldg(c.res_x + ((int64_t)b * p.H + (y - 1)) * p.W + (x - 1)) : 0.f;
        return t;
      };
      auto finalize = [&](const uint32_t (&v)[32], const int n0, const int pos, const TilePre& pre) {
        const int y = pre.y, x = pre.x;
        const bool interior = (y >= 1) && (y <= p.H) && (x >= 1) && (x <= p.W);
        const bool in_tensor = packed ? pre.b < p.batch : pos < HpWp;
        const int64_t img_off_t =
            packed ? ((int64_t)pre.b * c.o_img_stride + (int64_t)c.out_plane0 * c.o_plane_stride) * 8 : img_off;
        const int q = pre.q;                  // position inside the image
        // (half-row tensor: row y of the tensor holds image rows 2y - 1 and 2y of the up-sampled image)
        const int64_t up = (int64_t)(c.upsample == 2 ? y : 2 * y - 1) * Wp2 + (2 * x - 1);
        const float rx = (job.epi & 1) ? pre.rx
                         : (((kResX && c.res_x != nullptr) && interior) ? __ldg(c.res_x + ((int64_t)b * p.H + (y - 1)) * p.W + (x - 1)) : 0.f);
        const float* resw_p = bias_s + 2 * N + phase * N;
        const float* bias_t = interior ? bias_p : zero_s;
        const float scale_t = interior ? inv_scale : 0.f;
        if constexpr (N == 32 && Epi == 2) {
          if (c.head_w != nullptr) {
            // fused conv_flatten partials: this position's 32 activations . head_w[y - 1][:, 0..3]
            // (the weights come straight from L2, channel by channel: staging the warp's rows in shared memory or
            // broadcasting them with shuffles measured 16 % / 26 % slower on this launch, profiles/r2_tuning.txt)
            if (interior) {
              const float4* wrow = reinterpret_cast<const float4*>(c.head_w) + (y - 1) * 32;
              // (packed pairs, ss_common.cuh: the same fmas in the same order, two per instruction — this loop is
              // 192 of the launch's ~350 epilogue instructions per position)
              float2 acc_xy = make_float2(0.f, 0.f), acc_zw = make_float2(0.f, 0.f);
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                const float2 f2 = fma2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), bc2(inv_scale),
                                       *reinterpret_cast<const float2*>(bias_p + i));
                const float fa = fmaxf(f2.x, 0.f), fb = fmaxf(f2.y, 0.f);
                const float4 wa = __ldg(wrow + i), wb = __ldg(wrow + i + 1);
                acc_xy = fma2(bc2(fa), make_float2(wa.x, wa.y), acc_xy);
                acc_zw = fma2(bc2(fa), make_float2(wa.z, wa.w), acc_zw);
                acc_xy = fma2(bc2(fb), make_float2(wb.x, wb.y), acc_xy);
                acc_zw = fma2(bc2(fb), make_float2(wb.z, wb.w), acc_zw);
              }
              reinterpret_cast<float4*>(c.head_out)[((int64_t)b * p.H + (y - 1)) * p.W + (x - 1)] =
                  make_float4(acc_xy.x, acc_xy.y, acc_zw.x, acc_zw.y);
            }
            if (c.out == nullptr) return;      // nobody reads the activations themselves (no spec head requested)
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t hw[4], lw[4];
          // Border positions get scale 0 and a zero bias vector instead of a select per value: fma(acc, 0, 0) = +0.
          // Every convolution of this network is followed by ReLU, so the fp16 clamp is one-sided and the range
          // check a running maximum (tested once per unit).
          const float4 b0 = *reinterpret_cast<const float4*>(bias_t + n0 + g * 8);
          const float4 b1 = *reinterpret_cast<const float4*>(bias_t + n0 + g * 8 + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float rw[8];
          if ((kResX && c.res_x != nullptr)) {
            const float4 r0 = *reinterpret_cast<const float4*>(resw_p + n0 + g * 8);
            const float4 r1 = *reinterpret_cast<const float4*>(resw_p + n0 + g * 8 + 4);
            rw[0] = r0.x; rw[1] = r0.y; rw[2] = r0.z; rw[3] = r0.w; rw[4] = r1.x; rw[5] = r1.y; rw[6] = r1.z; rw[7] = r1.w;
          }
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            float2 f01 = fma2(make_float2(__uint_as_float(v[g * 8 + 2 * h]), __uint_as_float(v[g * 8 + 2 * h + 1])),
                              bc2(scale_t), make_float2(bb[2 * h], bb[2 * h + 1]));
            if ((kResX && c.res_x != nullptr)) f01 = fma2(bc2(rx), make_float2(rw[2 * h], rw[2 * h + 1]), f01);
            const float f0 = fmaxf(f01.x, 0.f);
            const float f1 = fmaxf(f01.y, 0.f);
            hw[h] = pack_rn<P>(f0, f1);
            if constexpr (PrecTraits<P>::fmt == 0) hmax = range_track(hmax, hw[h]);
            if constexpr (kSplit) lw[h] = pack_lo_f16(f0, f1, hw[h]);
            else lw[h] = 0u;
          }
          const uint4 ph = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          const int64_t plane_off = img_off_t + (int64_t)(n0 / 8 + g) * out_plane_stride;
          if (dbg & 2) {
            if (ph.x == 0x12345678u && lw[0] == 0x9abcdef0u) p.err[1] = 1;      // keep the values alive
          } else if (NoUp || !c.upsample) {
            if (in_tensor) {
              st16(c.out + plane_off + (int64_t)q * 8, ph);
              if constexpr (kSplit)
                st16(c.out_lo + plane_off + (int64_t)q * 8, make_uint4(lw[0], lw[1], lw[2], lw[3]));
            }
          } else if constexpr (!NoUp) if (interior) {
            uint16_t* o = c.out + plane_off + up * 8;
            if (c.upsample == 2) {
              st32x2(o, ph);
            } else if (c.pair_store) {
              st32x2(o, ph);
              st32x2(o + (int64_t)Wp2 * 8, ph);
            } else {
              st16(o, ph);
              st16(o + 8, ph);
              st16(o + (int64_t)Wp2 * 8, ph);
              st16(o + (int64_t)Wp2 * 8 + 8, ph);
            }
            if constexpr (kSplit) {
              const uint4 pl = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              uint16_t* ol = c.out_lo + plane_off + up * 8;
              if (c.upsample == 2) {
                st32x2(ol, pl);
              } else if (c.pair_store) {
                st32x2(ol, pl);
                st32x2(ol + (int64_t)Wp2 * 8, pl);
              } else {
                st16(ol, pl);
                st16(ol + 8, pl);
                st16(ol + (int64_t)Wp2 * 8, pl);
                st16(ol + (int64_t)Wp2 * 8 + 8, pl);
              }
            }
          }
        }
      };
      // one tile's 32-column block of TMEM buffer `buf` (dual layout: main + correction columns, both loads in flight
      // under one wait)
      auto load_block = [&](uint32_t (&v)[32], const int buf, const int mt, const int n0) {
        tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * kAccCols + mt * TS + n0), v);
        if constexpr (Dual) {
          uint32_t cv[32];
          tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * kAccCols + mt * TS + N + n0), cv);
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 sum = add2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                                    make_float2(__uint_as_float(cv[i]), __uint_as_float(cv[i + 1])));
            v[i] = __float_as_uint(sum.x);
            v[i + 1] = __float_as_uint(sum.y);
          }
        } else {
          tc_wait_ld();
        }
      };

      if constexpr (Epi == 1) {
        static_assert(Epi != 1 || (Rows != 0 && kUnitRows == 2), "the folded pool belongs to row-aligned units of two rows");
        // Row-aligned unit = image rows (2 lu, 2 lu + 1), 0-based.  This warp owns the vertical tile pair (A above B)
        // of its 32 columns: W = 256 (two tiles per row): column half `tile_par`, all N = 32 channels; W = 128 (one tile
        // per row): the whole row, channel block `tile_par` of the N = 64.  Every position is interior.
        constexpr int TPR = MT / 2;                      // tiles per image row
        static_assert(TPR == 1 || TPR == 2, "row-aligned units are a pair of image rows");
        static_assert(TPR == 2 ? N == 32 : N == 64, "channel blocks of the two epilogue warps of a quadrant");
        const int lu = u - b * p.units_per_image;
        const int buf = kk & 1;
        const uint32_t f_parity = ((uint32_t)kk >> 1) & 1u;
        ++kk;
        const int mtA = (TPR == 2) ? tile_par : 0, mtB = mtA + TPR;
        const int nb = (TPR == 2) ? 0 : 32 * tile_par;   // first channel of my block
        const int xc = (TPR == 2 ? tile_par * 128 : 0) + quad * 32 + lane;      // interior column, 0-based
        const int64_t posA = (int64_t)(2 * lu + 1) * Wp + (xc + 1);
        float rxA = 0.f, rxB = 0.f;
        if ((kResX && c.res_x != nullptr)) {
          const float* rp = c.res_x + ((int64_t)b * p.H + 2 * lu) * p.W + xc;
          rxA = __ldg(rp);
          rxB = __ldg(rp + p.W);
        }
        const float* resw_p = bias_s + 2 * N + phase * N;
        const int Wq = (p.W >> 1) + 2, Hq = (p.H >> 1) + 2;
        const int64_t pool_plane_stride = c.pool_plane_stride * 8;
        const int64_t pool_off = ((int64_t)b * c.pool_img_stride + (int64_t)(lu + 1) * Wq + ((xc >> 1) + 1)) * 8;
        ok = mbar_wait_t(accf0 + 8 * buf, f_parity, p.err, 3, w_accf);
        if (!ok) break;
        tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * kAccCols);
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {
          uint32_t am[16], ac[16], bm[16], bc[16];
          tc_ld16(t_row + (uint32_t)(mtA * TS + nb + c0), am);
          tc_ld16(t_row + (uint32_t)(mtA * TS + N + nb + c0), ac);
          tc_ld16(t_row + (uint32_t)(mtB * TS + nb + c0), bm);
          tc_ld16(t_row + (uint32_t)(mtB * TS + N + nb + c0), bc);
          tc_wait_ld();
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int ch = nb + c0 + g * 8;              // first of these 8 channels
            const float4 b0 = *reinterpret_cast<const float4*>(bias_p + ch);
            const float4 b1 = *reinterpret_cast<const float4*>(bias_p + ch + 4);
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
            float rw[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if ((kResX && c.res_x != nullptr)) {
              const float4 r0 = *reinterpret_cast<const float4*>(resw_p + ch);
              const float4 r1 = *reinterpret_cast<const float4*>(resw_p + ch + 4);
              rw[0] = r0.x; rw[1] = r0.y; rw[2] = r0.z; rw[3] = r0.w; rw[4] = r1.x; rw[5] = r1.y; rw[6] = r1.z; rw[7] = r1.w;
            }
            uint32_t ah[4], al[4], bh[4], bl[4], ph[4], pl[4];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int i0 = g * 8 + 2 * h, i1 = i0 + 1;
              const float2 bb2 = make_float2(bb[2 * h], bb[2 * h + 1]), rw2 = make_float2(rw[2 * h], rw[2 * h + 1]);
              float2 a01 = fma2(add2(make_float2(__uint_as_float(am[i0]), __uint_as_float(am[i1])),
                                     make_float2(__uint_as_float(ac[i0]), __uint_as_float(ac[i1]))), bc2(inv_scale), bb2);
              float2 q01 = fma2(add2(make_float2(__uint_as_float(bm[i0]), __uint_as_float(bm[i1])),
                                     make_float2(__uint_as_float(bc[i0]), __uint_as_float(bc[i1]))), bc2(inv_scale), bb2);
              if ((kResX && c.res_x != nullptr)) {
                a01 = fma2(bc2(rxA), rw2, a01);
                q01 = fma2(bc2(rxB), rw2, q01);
              }
              float a0 = fmaxf(a01.x, 0.f), a1 = fmaxf(a01.y, 0.f), q0v = fmaxf(q01.x, 0.f), q1v = fmaxf(q01.y, 0.f);
              // the 2 x 2 window: this thread's two rows, then the neighbouring column (lanes 2k, 2k + 1)
              float m0 = fmaxf(a0, q0v), m1 = fmaxf(a1, q1v);
              m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
              m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
              // (this path keeps the explicit clamp and a float running maximum: with the saturating conversion and
              // the packed maximum of `finalize` conv1_1.c2 measured 14 % SLOWER, profiles/r2_tuning_experiments.txt 17 —
              // with or without the saturating conversion, in a kernel of its own too)
              vmax = fmaxf(vmax, fmaxf(m0, m1));
              a0 = fminf(a0, 65504.f); a1 = fminf(a1, 65504.f); q0v = fminf(q0v, 65504.f); q1v = fminf(q1v, 65504.f);
              m0 = fminf(m0, 65504.f); m1 = fminf(m1, 65504.f);
              ah[h] = pack_plain<P>(a0, a1); al[h] = pack_lo_f16(a0, a1, ah[h]);
              bh[h] = pack_plain<P>(q0v, q1v); bl[h] = pack_lo_f16(q0v, q1v, bh[h]);
              ph[h] = pack_plain<P>(m0, m1); pl[h] = pack_lo_f16(m0, m1, ph[h]);
            }
            const int64_t plane_off = img_off + (int64_t)(ch / 8) * out_plane_stride;
            st16(c.out + plane_off + posA * 8, make_uint4(ah[0], ah[1], ah[2], ah[3]));
            st16(c.out_lo + plane_off + posA * 8, make_uint4(al[0], al[1], al[2], al[3]));
            st16(c.out + plane_off + (posA + Wp) * 8, make_uint4(bh[0], bh[1], bh[2], bh[3]));
            st16(c.out_lo + plane_off + (posA + Wp) * 8, make_uint4(bl[0], bl[1], bl[2], bl[3]));
            if ((lane & 1) == 0) {
              const int64_t po = pool_off + (int64_t)(ch / 8) * pool_plane_stride;
              st16(c.pool_out + po, make_uint4(ph[0], ph[1], ph[2], ph[3]));
              st16(c.pool_lo + po, make_uint4(pl[0], pl[1], pl[2], pl[3]));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acce0 + 8 * buf);
      }
      if constexpr (kSubAcc) {
        // n_sub accumulation groups of this unit arrive one buffer turn after the other; their sums live in registers
        // (this warp's MT / 2 tiles x N columns of its 32 positions) and are added in float32, round to nearest.
        const int n_sub = c.n_sub;
        const int q0 = halo + (u - b * p.units_per_image) * MT * 128;
        float acc[kMyTiles][N];
        for (int sub = 0; sub < n_sub && ok; ++sub, ++kk) {
          const int buf = kk & 1;
          const uint32_t f_parity = ((uint32_t)kk >> 1) & 1u;
          ok = mbar_wait_t(accf0 + 8 * buf, f_parity, p.err, 3, w_accf);
          if (!ok) break;
          tc_fence_after();
#pragma unroll
          for (int ti = 0; ti < kMyTiles; ++ti) {
            if constexpr (Dual) {
              // 16 columns at a time: the unit's sums hold N registers per tile, and two 32-column blocks (main and
              // correction columns) on top of them spilled — to L2, the CTA's shared memory leaves next to no L1
              // (conv7: -4 % / -2 %; the non-dual layers, one block per load, were 2-5 % faster as they are)
#pragma unroll
              for (int n0 = 0; n0 < N; n0 += 16) {
                uint32_t v[16], cv[16];
                const uint32_t t_blk = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                       (uint32_t)(buf * kAccCols + (tile_par + 2 * ti) * TS + n0);
                tc_ld16(t_blk, v);
                tc_ld16(t_blk + (uint32_t)N, cv);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                  const float2 sum = add2(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                                          make_float2(__uint_as_float(cv[i]), __uint_as_float(cv[i + 1])));
                  if (sub == 0) {
                    acc[ti][n0 + i] = sum.x;
                    acc[ti][n0 + i + 1] = sum.y;
                  } else {
                    acc[ti][n0 + i] += sum.x;
                    acc[ti][n0 + i + 1] += sum.y;
                  }
                }
              }
            } else {
#pragma unroll
              for (int n0 = 0; n0 < N; n0 += 32) {
                uint32_t v[32];
                load_block(v, buf, tile_par + 2 * ti, n0);
                if (sub == 0) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) acc[ti][n0 + i] = __uint_as_float(v[i]);
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) acc[ti][n0 + i] += __uint_as_float(v[i]);
                }
              }
            }
          }
          // all of this warp's TMEM reads of the buffer have completed (tcgen05.wait::ld after each load pair)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acce0 + 8 * buf);
        }
        if (!ok) break;
#pragma unroll
        for (int ti = 0; ti < kMyTiles; ++ti) {
          const int pos = q0 + (tile_par + 2 * ti) * 128 + quad * 32 + lane;
          const TilePre pre = preload(pos);
#pragma unroll
          for (int n0 = 0; n0 < N; n0 += 32) {
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(acc[ti][n0 + i]);
            finalize(v, n0, pos, pre);
          }
        }
      } else if constexpr (Epi != 1) {
       for (int g = 0; g < G && ok; ++g, ++kk) {
        const int buf = (G == 1) ? (kk & 1) : g;
        const uint32_t f_parity = (G == 1) ? (((uint32_t)kk >> 1) & 1u) : ((uint32_t)k & 1u);
        // first position of tile mt of the unit: consecutive runs of 128, or (row-aligned units) MT / 2 tiles in each
        // of the image rows 2 lu + 1 and 2 lu + 2 of the padded tensor
        const int q0 = Rows ? (kUnitRows * (u - b * p.units_per_image) + 1) * Wp + 1
                            : halo + ((u - b * p.units_per_image) * G + g) * MT * 128;
        auto tile_pos = [&](const int mt) {
          if constexpr (Rows != 0) return q0 + (mt / Rows) * Wp + (mt % Rows) * 128 + quad * 32 + lane;
          else return q0 + mt * 128 + quad * 32 + lane;
        };
        auto tile_row = [&](const int mt) {        // padded image row of tile mt, where the geometry knows it
          if constexpr (Rows != 0) return kUnitRows * (u - b * p.units_per_image) + 1 + mt / Rows;
          else return -1;
        };
        // this warp's tiles: tile_par, tile_par + 2, ...; the global operands of a tile are fetched one tile ahead
        // (the first before the wait for the accumulators), the loop stays rolled (code size)
        TilePre cur = preload(tile_pos(tile_par), tile_row(tile_par));
        ok = mbar_wait_t(accf0 + 8 * buf, f_parity, p.err, 3, w_accf);
        if (!ok) break;
        tc_fence_after();
#pragma unroll 1
        for (int mt = tile_par; mt < MT; mt += 2) {
          const int pos = tile_pos(mt);
          const TilePre nxt = (mt + 2 < MT) ? preload(tile_pos(mt + 2), tile_row(mt + 2)) : cur;
#pragma unroll
          for (int n0 = 0; n0 < N; n0 += 32) {
            uint32_t v[32];
            load_block(v, buf, mt, n0);
            finalize(v, n0, pos, cur);
          }
          cur = nxt;
        }
        // all of this warp's TMEM reads of the buffer have completed (tcgen05.wait::ld after each load pair)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acce0 + 8 * buf);
       }
      }
      if constexpr (PrecTraits<P>::fmt == 0) out_of_range |= range_hit(hmax) || vmax > 65504.f;
      if (n_phase == 2 && phase == 0) {
        // publish this warp's share of the unit to the c[1] producers of other CTAs: the lanes' stores are ordered
        // before lane 0's release by the warp barrier, and the release-reduction makes them visible at gpu scope
        // (a full __threadfence() + atomicAdd here stalled every epilogue warp until its stores had drained:
        // SS_TC_FENCE=1 keeps that form for A/B runs)
        asm volatile("fence.proxy.async.global;" ::: "memory");
        if (job.heavy_fence) __threadfence();
        __syncwarp();
        if (lane == 0) {
          if (job.heavy_fence) atomicAdd(job.flags + u, 1);
          else asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(job.flags + u), "r"(1) : "memory");
        }
      }
      if (n_phase == 2 && phase == 1 && job.ring > 0) {
        // this unit's copies of the intermediate tensor completed before its accumulators did: its slot may be reused
        __syncwarp();
        if (lane == 0) atomicAdd(job.flags2 + u, 1);
      }
    }
    if (out_of_range) p.err[2] = 1;
    if (p.prof && threadIdx.x == 0) {
      p.prof[blockIdx.x * 8 + 4] = w_accf;
      p.prof[blockIdx.x * 8 + 5] = clock64() - t_begin;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == w_prod) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

}  // namespace tc
}  // namespace ss
