// tcgen05.mma issue / execution rate probe for the forward-backward kernel's shapes:
// how many clocks does ONE M = 128, K = 16 bf16 MMA take as a function of
//   * N (32, 64, 128, 256),
//   * where A comes from: tensor memory ("TS" form, what fb_tc_pass_kernel uses) or a shared-memory descriptor ("SS"),
//   * one accumulator or two alternating ones,
// when a single elected lane issues a run of them back to back and commits them to an mbarrier?  The forward-backward
// step measures ~45-60 clocks per 128 x 64 x 16 MMA where the pipe's floor is 128 N / 256 = 32: this probe separates the
// fixed per-instruction cost from the N-proportional part.  Operands are zeros (only timing matters); the B operand
// uses the validated no-swizzle MN-major layout of tools/microbench_umma.cu, A (SS form) the dense K-major one.
//
// NOT YET RUN (written at the end of round 1 with the GPU budget spent); wrap it in `timeout 30` on first use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_umma_rate microbench_umma_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int M = 128;
constexpr int kSmemBytes = 64 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t start, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((start >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         (1ull << 46);                                                  // version 1, no swizzle
}

// out[0] = clocks from the first issue to the completion of `reps` MMAs; out[1] = clocks spent issuing
__global__ void __launch_bounds__(128, 1) umma_rate(int N, int ss_form, int two_acc, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int x = tid; x < kSmemBytes / 16; x += 128) reinterpret_cast<uint4*>(smem)[x] = make_uint4(0, 0, 0, 0);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = s_tmem;
  // TS form: A = TMEM columns [0, 8) (zeros after a tcgen05.st); accumulators from column 256 (second one only for N <= 128)
  {
    const uint32_t tlane = tbase + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%1,%1,%1};" ::"r"(tlane), "r"(0u) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%1,%1,%1};" ::"r"(tlane + 4), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == 0 && elect_one_sync()) {
    // instruction descriptor: D = F32, A = B = BF16, A K-major, B MN-major, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t b_lbo = (uint32_t)(N / 8) * 128u;                   // MN-major B: between groups of 8 k; SBO = 128
    const uint64_t b_desc = make_desc(smem_u32(smem), b_lbo, 128u);    // 16 k x N bf16 = 32 N bytes <= 8 KB
    const uint64_t a_desc = make_desc(smem_u32(smem) + 16 * 1024, 128u, 256u);   // K-major A: 16 row groups x 2 cores
    const uint32_t d0 = tbase + 256, d1 = tbase + 256 + (uint32_t)N;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = (two_acc && (r & 1)) ? d1 : d0;
      const uint32_t acc = r >= (two_acc ? 2 : 1);
      if (ss_form)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                     ::"r"(d), "r"(tbase), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&s_bar)), "r"(0u) : "memory");
    const long long t2 = clock64();
    out[0] = t2 - t0;
    out[1] = t1 - t0;
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512) : "memory");
}

int main() {
  CK(cudaSetDevice(0));
  long long* d_out; CK(cudaMalloc(&d_out, 2 * sizeof(long long)));
  CK(cudaFuncSetAttribute(umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  const int reps = 256;
  for (int ss = 0; ss < 2; ++ss)
    for (int two = 0; two < 2; ++two)
      for (int N : {32, 64, 128, 256}) {
        if (two && N > 128) continue;                                  // two accumulators need 2 N <= 256 columns
        long long h[2] = {0, 0};
        for (int rep = 0; rep < 2; ++rep) {                            // second run: warm instruction cache
          umma_rate<<<1, 128, kSmemBytes>>>(N, ss, two, reps, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("{\"form\": \"%s\", \"N\": %d, \"error\": \"%s\"}\n", ss ? "SS" : "TS", N, cudaGetErrorString(e)); return 1; }
        }
        CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
        printf("{\"form\": \"%s\", \"N\": %d, \"accumulators\": %d, \"mmas\": %d, \"clk_per_mma\": %.1f, \"issue_clk_per_mma\": %.1f, "
               "\"floor_clk\": %.1f}\n", ss ? "SS" : "TS", N, two ? 2 : 1, reps, (double)h[0] / reps, (double)h[1] / reps, 128.0 * N / 256.0);
      }
  return 0;
}
