// tcgen05.mma bring-up probe: D[128 x 32] (fp32, TMEM) = A[128 x K] (bf16, TMEM: "TS" form) x B[K x 32] (bf16, shared
// memory descriptor), K = 64 = 4 MMAs of K = 16, cta_group::1, kind::f16.  Small-integer operands make the product exact,
// so each variant of the layout assumptions either matches the host result bit for bit or it does not:
//   bit 0: swap the descriptor's leading / stride byte offsets
//   bit 1: B stored K-major (else MN-major, i.e. the N = clip index contiguous)
//   bit 2: A packs the ODD k in the low half of a TMEM column (else the even k)
// This decides the operand layouts of a tensor-core forward-backward kernel.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_umma microbench_umma.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int M = 128, N = 32, K = 64;

__host__ __device__ inline float a_val(int m, int k) { return (float)((m * 3 + k * 5) % 7 - 3); }
__host__ __device__ inline float b_val(int k, int n) { return (float)((k * 2 + n * 3) % 5 - 2); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) umma_probe(int variant, float* out /*[M][N]*/) {
  __shared__ __align__(128) uint8_t sB[K * N * 2];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool swap = variant & 1, kmajor = variant & 2, odd_low = variant & 4;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // B -> shared memory, canonical no-swizzle layout of 8 x 16-byte core matrices
  const uint32_t LBO = kmajor ? 128u : (uint32_t)(N / 8) * 128u;     // MN-major: between groups of 8 k
  const uint32_t SBO = kmajor ? (uint32_t)(K / 8) * 128u : 128u;     // MN-major: between 16-byte chunks along N
  for (int x = tid; x < K * N; x += 128) {
    const int k = x / N, n = x % N;
    uint32_t off;
    if (kmajor) off = (n / 8) * SBO + (k / 8) * LBO + (n % 8) * 16 + (k % 8) * 2;
    else        off = (n / 8) * SBO + (k / 8) * LBO + (k % 8) * 16 + (n % 8) * 2;
    *reinterpret_cast<__nv_bfloat16*>(sB + off) = __float2bfloat16(b_val(k, n));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = s_tmem;
  const uint32_t tlane = tbase + ((uint32_t)(warp * 32) << 16);
  // A -> TMEM: row m = 32 warp + lane, column c holds k = 2c and 2c + 1
  for (int c = 0; c < K / 2; c += 4) {
    uint32_t v[4];
    for (int i = 0; i < 4; ++i) {
      const int m = warp * 32 + lane;
      const __nv_bfloat16 e0 = __float2bfloat16(a_val(m, 2 * (c + i))), e1 = __float2bfloat16(a_val(m, 2 * (c + i) + 1));
      const uint32_t lo = __bfloat16_as_ushort(odd_low ? e1 : e0), hi = __bfloat16_as_ushort(odd_low ? e0 : e1);
      v[i] = lo | (hi << 16);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tlane + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (tid == 0) {
    // instruction descriptor: D = F32, A = B = BF16, A K-major, B major per variant, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((kmajor ? 0u : 1u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t d_tmem = tbase + 64;
    for (int kb = 0; kb < K / 16; ++kb) {
      const uint32_t start = smem_u32(sB) + kb * 2 * LBO;             // 16 k = 2 groups of 8
      const uint32_t lbo = swap ? SBO : LBO, sbo = swap ? LBO : SBO;
      const uint64_t desc = (uint64_t)((start >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
                            ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);     // version 1, no swizzle
      const uint32_t a_tmem = tbase + kb * 8;                          // 16 bf16 = 8 columns
      const uint32_t acc = kb > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                   ::"r"(d_tmem), "r"(a_tmem), "l"(desc), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)) : "memory");
  }
  // everyone waits for the MMAs
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&s_bar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t d[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]), "=r"(d[9]),
                 "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15]), "=r"(d[16]), "=r"(d[17]), "=r"(d[18]),
                 "=r"(d[19]), "=r"(d[20]), "=r"(d[21]), "=r"(d[22]), "=r"(d[23]), "=r"(d[24]), "=r"(d[25]), "=r"(d[26]), "=r"(d[27]),
                 "=r"(d[28]), "=r"(d[29]), "=r"(d[30]), "=r"(d[31])
               : "r"(tlane + 64));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < 32; ++n) out[(warp * 32 + lane) * N + n] = __uint_as_float(d[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "n"(512) : "memory");
}

int main() {
  CK(cudaSetDevice(0));
  float* d_out; CK(cudaMalloc(&d_out, M * N * sizeof(float)));
  static float h[M * N], ref[M * N];
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += a_val(m, k) * b_val(k, n); ref[m * N + n] = s; }
  // variant 0 is the layout that matches (measured on B200); the swapped-offset variants read outside the operand and
  // fault, so only the validated one is run
  for (int v = 0; v < 1; ++v) {
    CK(cudaMemset(d_out, 0, M * N * sizeof(float)));
    umma_probe<<<1, 128>>>(v, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"variant\": %d, \"error\": \"%s\"}\n", v, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
    int bad = 0; for (int i = 0; i < M * N; ++i) bad += (h[i] != ref[i]);
    printf("{\"variant\": %d, \"swap_lbo_sbo\": %d, \"b_kmajor\": %d, \"a_odd_low\": %d, \"mismatches\": %d, \"d00\": %g, \"ref00\": %g, \"d_1_1\": %g, \"ref_1_1\": %g, \"d_127_31\": %g, \"ref_127_31\": %g}\n",
           v, v & 1, (v >> 1) & 1, (v >> 2) & 1, bad, h[0], ref[0], h[N + 1], ref[N + 1], h[127 * N + 31], ref[127 * N + 31]);
  }
  return 0;
}
