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
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
};

__global__ void __launch_bounds__(128, 1) bench(Cfg c, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
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
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
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
            tc_mma(tmem + ((uint32_t)j & nmask) * (uint32_t)c.n, da, db0, idesc, 1u);
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
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  const int iters = 2048;
  printf("%-10s %4s %5s %7s %8s | %s\n", "layout", "N", "nacc", "a_step", "ctas", "cycles/MMA (ideal N/2)");
  for (int ctas : {1, 148})
    for (int sw : {0, 2})
      for (int n : {32, 64, 128, 256})
        for (int nacc : {1, 2, 8})
          for (int a_step : {0, 2048}) {
            if (nacc * n > 512) continue;
            Cfg c{};
            c.n = n; c.swizzle = sw; c.n_acc = nacc; c.a_step = a_step; c.iters = iters;
            if (sw == 0) {        // interleaved: core matrices of 8 rows x 16 B; K chunks far apart
              c.a_sbo = 128; c.a_lbo = 20480; c.b_sbo = 128; c.b_lbo = (uint32_t)n * 16;
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
            printf("%-10s %4d %5d %7d %8d | %7.1f (%d)\n", sw ? "sw128" : "interleave", n, nacc, a_step, ctas, avg / iters, n / 2);
          }
  return 0;
}
