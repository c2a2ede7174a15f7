// Accuracy experiment: how does the error of a chain of tcgen05.mma (kind::f16, fp32 accumulate in TMEM) grow with
// the chain length, and how much of it goes away when the chain is split over several TMEM accumulators that are
// added in fp32 (round-to-nearest) afterwards?  This is the measurement behind the "accurate" classifier variant
// (operand precisions).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o acc_chain_test acc_chain_test.cu
//
// One CTA, M = 128, N = 64, K = 16 per MMA.  A = relu(normal) (what a ReLU network feeds its convolutions),
// B = normal * 2^10 (weights lifted like weight_scale() does), both rounded to fp16; the reference is the exact
// float64 sum of the fp16 products.  MMA i of a chain of n goes to sub-accumulator i * SUB / n.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

constexpr int N = 64, BATCH = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// a: [n][2][128][8] fp16 (K-half major, then row, then 8 k), b: [n][2][N][8]; out: [sub][128][N] fp32
__global__ void __launch_bounds__(128, 1) chain(const uint16_t* a, const uint16_t* b, int n, int sub, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  unsigned char* sa = smem;                       // BATCH * 4096
  unsigned char* sb = smem + BATCH * 4096;        // BATCH * N * 32
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F16 x F16 -> F32
  uint32_t parity = 0;
  for (int i0 = 0; i0 < n; i0 += BATCH) {
    const int nb = (n - i0 < BATCH) ? (n - i0) : BATCH;
    for (int i = threadIdx.x; i < nb * 4096 / 16; i += 128)
      reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(a + (size_t)i0 * 2048)[i];
    for (int i = threadIdx.x; i < nb * N * 32 / 16; i += 128)
      reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(b + (size_t)i0 * N * 16)[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 32) {
      for (int j = 0; j < nb; ++j) {
        const int i = i0 + j;
        const int s = (int)((long long)i * sub / n);
        const bool first = (i == 0) || ((int)((long long)(i - 1) * sub / n) != s);
        const uint32_t aa = smem_u32(sa + j * 4096), bb = smem_u32(sb + j * N * 32);
        const uint64_t hi = ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint64_t da = hi | ((uint64_t)((2048 >> 4) & 0x3FFF) << 16) | (uint64_t)((aa >> 4) & 0x3FFF);
        const uint64_t db = hi | ((uint64_t)(((N * 16) >> 4) & 0x3FFF) << 16) | (uint64_t)((bb >> 4) & 0x3FFF);
        tc_mma(tmem + (uint32_t)(s * N), da, db, idesc, first ? 0u : 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)), "r"(parity) : "memory");
    }
    parity ^= 1u;
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int s = 0; s < sub; ++s)
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
          "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * N + c0)));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 32; ++i) out[((size_t)s * 128 + warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
    }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
}

static double randn() {
  double u1 = (rand() + 1.0) / (RAND_MAX + 2.0), u2 = (rand() + 1.0) / (RAND_MAX + 2.0);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static uint16_t f2h(float f) { __half h = __float2half_rn(f); uint16_t b; memcpy(&b, &h, 2); return b; }
static float h2f(uint16_t b) { __half h; memcpy(&h, &b, 2); return __half2float(h); }

int main() {
  srand(1234);
  const int n_max = 144;
  cudaFuncSetAttribute(chain, cudaFuncAttributeMaxDynamicSharedMemorySize, BATCH * (4096 + N * 32));
  uint16_t *da, *db;
  float* dout;
  cudaMalloc(&da, (size_t)n_max * 2048 * 2);
  cudaMalloc(&db, (size_t)n_max * N * 16 * 2);
  cudaMalloc(&dout, (size_t)8 * 128 * N * 4);
  printf("%-10s %5s %4s | %11s %11s | %s\n", "data", "chain", "sub", "max rel-max", "rms rel-rms", "mean signed err / rms ref");
  for (int data = 0; data < 2; ++data) {
    // element (row r, k) of chunk i: A[i][k / 8][r][k % 8]
    std::vector<uint16_t> ha((size_t)n_max * 2048), hb((size_t)n_max * N * 16);
    for (auto& v : ha) { double x = randn(); v = f2h((float)(data == 0 ? fmax(x, 0.0) : fabs(x))); }
    for (auto& v : hb) { double x = randn() * 1024.0; v = f2h((float)(data == 0 ? x : fabs(x))); }
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    for (int n : {1, 4, 9, 18, 36, 72, 144}) {
      std::vector<double> ref((size_t)128 * N, 0.0);
      for (int i = 0; i < n; ++i)
        for (int r = 0; r < 128; ++r)
          for (int c = 0; c < N; ++c) {
            double s = 0;
            for (int k = 0; k < 16; ++k)
              s += (double)h2f(ha[(size_t)i * 2048 + (k / 8) * 1024 + r * 8 + k % 8]) *
                   (double)h2f(hb[(size_t)i * N * 16 + (k / 8) * N * 8 + c * 8 + k % 8]);
            ref[(size_t)r * N + c] += s;
          }
      double ref_max = 0, ref_rms = 0;
      for (double v : ref) { ref_max = fmax(ref_max, fabs(v)); ref_rms += v * v; }
      ref_rms = sqrt(ref_rms / ref.size());
      for (int sub : {1, 2, 3, 4, 8}) {
        if (sub > n) continue;
        chain<<<1, 128, BATCH * (4096 + N * 32)>>>(da, db, n, sub, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> h((size_t)sub * 128 * N);
        cudaMemcpy(h.data(), dout, h.size() * 4, cudaMemcpyDeviceToHost);
        double emax = 0, erms = 0, emean = 0;
        for (size_t j = 0; j < (size_t)128 * N; ++j) {
          float s = 0.f;
          for (int q = 0; q < sub; ++q) s += h[(size_t)q * 128 * N + j];      // fp32, round to nearest
          const double err = (double)s - ref[j];
          emax = fmax(emax, fabs(err)); erms += err * err; emean += err * (ref[j] >= 0 ? 1.0 : -1.0);
        }
        printf("%-10s %5d %4d | %11.3e %11.3e | %+.3e\n", data == 0 ? "relu*norm" : "all>=0", n, sub, emax / ref_max,
               sqrt(erms / (128.0 * N)) / ref_rms, emean / (128.0 * N) / ref_rms);
      }
    }
  }
  return 0;
}
