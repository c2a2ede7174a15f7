// Microbenchmark 2: does operand-collector reuse lift the shared-memory operand-fetch bound of small-N MMAs?
//   plain : tcgen05.mma.cta_group::1.kind::f16                        (cost measured earlier: max(N/2, 32 + N/4) cycles)
//   a     : ....collector::a::fill / ::lastuse                        (groups of `g` MMAs share A, B streams)
//   ws    : tcgen05.mma.ws....collector::b0::fill / ::use / ::lastuse (groups of `g` MMAs share B, A streams)
// Each variant also writes its accumulators so that they can be compared with the plain result (same operands).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_ws_bench umma_ws_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0, lane_out = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %2;\n\t@px mov.s32 %1, 1;\n\tmov.s32 %0, rx;\n\t}"
               : "+r"(lane_out), "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred;
}
#define MMA(NAME, OPC)                                                                                              \
  __device__ __forceinline__ void NAME(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {        \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t" OPC " [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), \
                 "l"(db), "r"(idesc), "r"(acc) : "memory");                                                          \
  }
MMA(mma_plain, "tcgen05.mma.cta_group::1.kind::f16")
MMA(mma_a_fill, "tcgen05.mma.cta_group::1.kind::f16.collector::a::fill")
MMA(mma_a_use, "tcgen05.mma.cta_group::1.kind::f16.collector::a::use")
MMA(mma_a_last, "tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse")
MMA(mma_ws_plain, "tcgen05.mma.ws.cta_group::1.kind::f16")
MMA(mma_ws_fill, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill")
MMA(mma_ws_use, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use")
MMA(mma_ws_last, "tcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse")

struct Cfg { int n, variant, g, iters; };   // variant: 0 plain, 1 collector::a, 2 ws + collector::b0, 3 ws without collector

template <int V, int G>
__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out, float* acc_out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem)[i] = (h & 0x83ff83ffu) | 0x38003800u;   // two fp16 in +-[0.5, 1)
  }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const int N = c.n;
  constexpr int g = G;
  if (warp == 1) {
    // fp16 operands, f32 accumulate, K-major, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    const uint64_t hi = ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
    // A tiles: 128 rows x 16 B x 2 K-halves, K-halves 2048 B apart, tiles 4096 B apart (16 of them);  B: N rows
    const uint64_t da0 = hi | ((uint64_t)(2048 >> 4) << 16) | (uint64_t)((a0 >> 4) & 0x3FFF);
    const uint64_t db0 = hi | ((uint64_t)((N * 16) >> 4) << 16) | (uint64_t)((b0 >> 4) & 0x3FFF);
    const uint32_t b_tile = (uint32_t)(N * 32) >> 4;
    long long t0 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one()) {
#pragma unroll 1
        for (int i = 0; i < c.iters; i += 8) {
          const uint32_t acc = (i >= 8) ? 1u : 0u;     // the first pass over the accumulators overwrites
          // straight-line groups of G MMAs, every descriptor = base + compile-time step (issue cost must not hide the pipe)
#pragma unroll
          for (int j8 = 0; j8 < 8; ++j8) {
            const int grp = (j8 / G) & 3, j = j8 % G;
            if (V == 1) {
              const uint64_t da = da0 + (uint64_t)(grp * 256), db = db0 + (uint64_t)(j * b_tile);
              const uint32_t d = tmem + (uint32_t)(j * N);
              if (j == 0) mma_a_fill(d, da, db, idesc, acc);
              else if (j == G - 1) mma_a_last(d, da, db, idesc, acc);
              else mma_a_use(d, da, db, idesc, acc);
            } else {
              const uint64_t da = da0 + (uint64_t)(j * 256), db = db0 + (uint64_t)(grp * b_tile);
              const uint32_t d = tmem + (uint32_t)(j * N);
              if (V == 0) mma_plain(d, da, db, idesc, acc);
              else if (V == 3) mma_ws_plain(d, da, db, idesc, acc);
              else if (j == 0) mma_ws_fill(d, da, db, idesc, acc);
              else if (j == G - 1) mma_ws_last(d, da, db, idesc, acc);
              else mma_ws_use(d, da, db, idesc, acc);
            }
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"((uint32_t)rep) : "memory");
      }
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (blockIdx.x == 0 && acc_out) {      // dump accumulator 0..g-1 (lane quadrant = warp)
    for (int col = 0; col < g * N; col += 32) {
      uint32_t v[32];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)col));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int k = 0; k < 32; ++k) acc_out[(size_t)threadIdx.x * 512 + col + k] = __uint_as_float(v[k]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

int main() {
  long long* d;
  float* acc;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaMalloc(&acc, 128 * 512 * sizeof(float));

  static float ref[128 * 512], got[128 * 512];
  const char* names[] = {"plain (B shared by g)", "collector::a (A shared by g)", "ws + collector::b0 (B shared by g)", "ws, no collector"};
  printf("%-36s %4s %3s | cycles/MMA (ideal N/2) | accumulators vs plain\n", "variant", "N", "g");
  for (int n : {64, 128, 32})
    for (int g : {2, 4})
      for (int variant : {0, 3, 2, 1}) {
        if (g * n > 512) continue;
        Cfg c{n, variant, g, 2048};
        cudaMemset(acc, 0, sizeof(ref));
        auto run = [&](auto kern) {
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
          kern<<<148, 128, 160 * 1024>>>(c, d, acc);
        };
        if (g == 2) { if (variant == 0) run(bench<0, 2>); else if (variant == 1) run(bench<1, 2>); else if (variant == 2) run(bench<2, 2>); else run(bench<3, 2>); }
        else { if (variant == 0) run(bench<0, 4>); else if (variant == 1) run(bench<1, 4>); else if (variant == 2) run(bench<2, 4>); else run(bench<3, 4>); }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-36s %4d %3d | CUDA error: %s\n", names[variant], n, g, cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemcpy(got, acc, sizeof(got), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += (double)h[i];
        avg /= 148;
        const char* verdict = "(reference)";
        if (variant == 0) memcpy(ref, got, sizeof(ref));
        else if (variant == 1) verdict = "n/a (different operand pairing)";
        else {
          size_t bad = 0;
          for (int r = 0; r < 128; ++r)
            for (int col = 0; col < g * n; ++col) bad += (ref[r * 512 + col] != got[r * 512 + col]);
          static char buf[64];
          snprintf(buf, sizeof(buf), bad ? "%zu of %d differ" : "identical", bad, 128 * g * n);
          verdict = buf;
        }
        printf("%-36s %4d %3d | %7.1f (%d) | %s\n", names[variant], n, g, avg / c.iters, n / 2, verdict);
      }
  return 0;
}
