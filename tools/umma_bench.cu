// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, bf16, cta_group::1, SS mode) as a function of N, the
// shared-memory layout (un-swizzled "interleave" vs 128B swizzle), the number of independent accumulators and
// operand reuse.  Data content is irrelevant (zeros).  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0, lane_out = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %2;\n\t@px mov.s32 %1, 1;\n\tmov.s32 %0, rx;\n\t}"
               : "+r"(lane_out), "+r"(pred) : "r"(0xFFFFFFFFu));
  return pred;
}

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

struct Cfg {
  int n;          // N
  int swizzle;    // 0 none, 2 = 128B
  int n_acc;      // independent accumulators cycled through
  int a_step;     // bytes added to the A start address per MMA (operand reuse vs streaming), cycled over 8 steps
  int iters;      // MMAs issued
  int b_step;     // bytes added to the B start address per MMA, cycled over 8 steps
  int b_every;    // B advances every b_every MMAs (1 = every MMA)
  int fill;       // 0: operands all zero, 1: pseudo-random bf16 in (-1, 1), 2: random with zero accumulate flag
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
};

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ volatile int done;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (c.fill) {
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 97u;
      h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      // two bf16 with exponent 0x7e (0.5 .. 1) or smaller, random sign and mantissa
      v = (h & 0x807f807fu) | 0x3f003f00u;
    }
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(1));
    done = 0;
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
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024) + (uint32_t)c.b_step;
    const uint64_t hi_a = ((uint64_t)((c.a_sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)c.swizzle << 61);
    const uint64_t hi_b = ((uint64_t)((c.b_sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)c.swizzle << 61);
    const uint64_t da0 = hi_a | ((uint64_t)((c.a_lbo >> 4) & 0x3FFF) << 16) | (uint64_t)((a0 >> 4) & 0x3FFF);
    const uint64_t db0 = hi_b | ((uint64_t)((c.b_lbo >> 4) & 0x3FFF) << 16) | (uint64_t)((b0 >> 4) & 0x3FFF);
    long long t0 = 0;
    for (int rep = 0; rep < 2; ++rep) {          // rep 0 warms up
      t0 = clock64();
      if (elect_one()) {
        const uint32_t nmask = (uint32_t)c.n_acc - 1u;      // n_acc is a power of two
        const uint32_t astep = (uint32_t)c.a_step >> 4;
#pragma unroll 1
        for (int i = 0; i < c.iters; i += 8) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint64_t da = da0 + (uint64_t)(j * astep);
            tc_mma(tmem + ((uint32_t)j & nmask) * (uint32_t)c.n, da, db0, idesc, c.fill == 2 ? 0u : 1u);
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
    done = 1;
    if ((threadIdx.x & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar2)) : "memory");
  } else if (warp >= 2) {
    if (c.fill & 1) {          // spin on a barrier that completes only at the end (what idle roles of the conv kernel do)
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(0u) : "memory");
      }
    }
    if (c.fill & 2) {          // stream TMEM loads (what the epilogue does) until the MMA warp is done
      uint32_t v[32];
      uint32_t sink = 0;
      while (!done) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + 256u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        sink ^= v[0] ^ v[31];
      }
      if (sink == 0x12345u) out[0] = 0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 2048;
  printf("%-10s %4s %5s %7s %8s | %s\n", "layout", "N", "nacc", "a_step", "ctas", "cycles/MMA (ideal N/2)");
  for (int ctas : {148})
    for (int sw : {0})
      for (int n : {32, 64, 96, 128})
       for (int b_every : {1})
        for (int b_step : {20480, 4704, 24672, 4608, 4640})
        for (int nacc : {2})
          for (int a_step : {2048}) {
            if (nacc * n > 512) continue;
            Cfg c{};
            c.n = n; c.swizzle = sw; c.n_acc = nacc; c.a_step = a_step; c.iters = iters; c.b_step = 0; c.b_every = b_every; c.fill = 0;
            if (sw == 0) {        // interleaved: core matrices of 8 rows x 16 B; K chunks far apart
              c.a_sbo = 128; c.a_lbo = (uint32_t)b_step; c.b_sbo = 128; c.b_lbo = (uint32_t)n * 16;
            } else {              // 128B swizzle K-major: rows of 128 B, 8-row atoms of 1024 B
              c.a_sbo = 1024; c.a_lbo = 16; c.b_sbo = 1024; c.b_lbo = 16;
            }
            bench<<<ctas, 128, 160 * 1024>>>(c, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < ctas; ++i) avg += (double)h[i];
            avg /= ctas;
            printf("%-10s %4d %5d %7d %8d a_lbo=%d | %7.1f (%d)\n", sw ? "sw128" : "interleave", n, nacc, a_step, ctas, b_step, avg / iters, n / 2);
          }
  return 0;
}
