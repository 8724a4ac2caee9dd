// TMEM-as-operand-store microbenchmark for the max-plus inner loop (sm_100a).
//
// Question: can the resident logA^T shard live in TENSOR MEMORY (256 KB/SM, otherwise unused by a SIMT kernel) and be
// streamed into registers with tcgen05.ld while delta comes from shared memory -- at the same FADD+FMNMX3 dispatch
// rate as an all-shared-memory kernel?  If yes, a 2-CTA cluster can keep a 181 x 368 shard (266 KB) resident, and
// 2-CTA clusters pack all 148 SMs (4-CTA clusters strand 16).
//
// Layout under test (the planned forward kernel): 256 threads = 8 warps; warp w reads TMEM lane quadrant w & 3; the two
// warps of a quadrant serve different clip groups ("pipelines").  Lane l -> (jg = l >> 2 within the quadrant, q = l & 3
// K-split).  Thread tile MB x NJ = 7 clips x 6 targets; per K chunk (4 k-values per lane) a thread needs 7 float4 of
// delta (LDS.128, broadcast over jg) and 6 float4 of logA^T = 24 consecutive TMEM columns of its own lane.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_tmem microbench_tmem.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int MB = 7, NJ = 6, KS = 4;
constexpr int NCHUNK_T = 21;            // chunks served from TMEM: 21 * 24 = 504 of 512 columns
constexpr int NCHUNK_S = 2;             // chunks served from shared memory (the K tail)
constexpr int NCHUNK = NCHUNK_T + NCHUNK_S;
constexpr int KP = NCHUNK * 4 * KS;     // 368
constexpr int THREADS = 256;
constexpr int NSLOT = 32 * NJ;          // 192 target slots per CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float a_value(int slot, int k) { return -(float)((slot * 131 + k * 17) % 1000) * 0.125f; }
__device__ __forceinline__ float d_value(int pipe, int b, int k) { return -(float)((pipe * 7 + b * 53 + k * 29) % 777) * 0.25f; }

__device__ __forceinline__ void tmem_ld24(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                 "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr));
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23])
               : "r"(taddr + 16));
}
// wait for the outstanding tcgen05.ld of this thread; the registers are threaded through so that no use can be
// scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld24(float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]),
                 "+r"(u[8]), "+r"(u[9]), "+r"(u[10]), "+r"(u[11]), "+r"(u[12]), "+r"(u[13]), "+r"(u[14]), "+r"(u[15]),
                 "+r"(u[16]), "+r"(u[17]), "+r"(u[18]), "+r"(u[19]), "+r"(u[20]), "+r"(u[21]), "+r"(u[22]), "+r"(u[23])
               :: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* r) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(r);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
}

__device__ __forceinline__ void cells(float* acc, const float4* d, const float* a) {
#pragma unroll
  for (int b = 0; b < MB; ++b)
#pragma unroll
    for (int n = 0; n < NJ; ++n) {
      float m = acc[b * NJ + n];
      m = fmaxf(m, __fadd_rn(d[b].x, a[n * 4 + 0]));
      m = fmaxf(m, __fadd_rn(d[b].y, a[n * 4 + 1]));
      m = fmaxf(m, __fadd_rn(d[b].z, a[n * 4 + 2]));
      m = fmaxf(m, __fadd_rn(d[b].w, a[n * 4 + 3]));
      acc[b * NJ + n] = m;
    }
}

// MODE 0: logA^T from TMEM (+ tail from smem); MODE 1: everything from shared memory (needs 192*368*4 = 283 KB -> uses
// only the first 16 chunks, wrapped, to fit; timing comparison only); MODE 2: TMEM with a one-chunk register prefetch
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1)
tmem_kernel(int steps, long long* cycles, int* mismatches, float* sink) {
  extern __shared__ __align__(128) float smem[];
  float* sD = smem;                              // [2 pipes][MB][KP]
  float* sA = sD + 2 * MB * KP;                  // MODE 1: [NSLOT][KP_S]; else tail [NSLOT][NCHUNK_S*16]
  constexpr int KP_S = (MODE == 1) ? 16 * 16 + 16 : NCHUNK_S * 4 * KS + 16;   // +16: row stride == 16 (mod 32) banks
  __shared__ uint32_t s_tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Q = warp & 3, pipe = warp >> 2;
  const int jg = Q * 8 + (lane >> 2), q = lane & 3;

  for (int x = tid; x < 2 * MB * KP; x += THREADS) {
    const int p = x / (MB * KP), b = (x / KP) % MB, k = x % KP;
    sD[x] = d_value(p, b, k);
  }
  for (int x = tid; x < NSLOT * KP_S; x += THREADS) {
    const int slot = x / KP_S, kk = x % KP_S;
    const int k = (MODE == 1) ? kk : NCHUNK_T * 16 + kk;
    sA[x] = a_value(slot, k);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem_base)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tbase = s_tmem_base + ((uint32_t)(Q * 32) << 16);
  if (pipe == 0) {
    // fill my quadrant: lane's columns [c*24 + n*4 + kk] = logA^T[slot jg + 32 n][k = 16 c + 4 q + kk]
    for (int c = 0; c < NCHUNK_T; ++c)
      for (int g = 0; g < 3; ++g) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int col = g * 8 + i, n = col >> 2, kk = col & 3;
          v[i] = a_value(jg + 32 * n, 16 * c + 4 * q + kk);
        }
        tmem_st8(tbase + c * 24 + g * 8, v);
      }
    asm volatile("tcgen05.wait::st.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");

  float acc[MB * NJ];
#pragma unroll
  for (int i = 0; i < MB * NJ; ++i) acc[i] = -INFINITY;
  const float4* pD = reinterpret_cast<const float4*>(sD + pipe * MB * KP) + q;
  const float4* pA = reinterpret_cast<const float4*>(sA);
  float bias = 0.f;

  const long long t0 = clock64();
  for (int s = 0; s < steps; ++s) {
    if (MODE == 2) {
      float a0[24], a1[24];
      tmem_ld24(tbase, a0);
#pragma unroll 1
      for (int c = 0; c < NCHUNK_T - 1; c += 2) {
        float4 d[MB];
        tmem_wait_ld24(a0);
        tmem_ld24(tbase + (c + 1) * 24, a1);
#pragma unroll
        for (int b = 0; b < MB; ++b) d[b] = pD[b * (KP / 4) + c * KS];
        cells(acc, d, a0);
        tmem_wait_ld24(a1);
        tmem_ld24(tbase + (c + 2) * 24, a0);
#pragma unroll
        for (int b = 0; b < MB; ++b) d[b] = pD[b * (KP / 4) + (c + 1) * KS];
        cells(acc, d, a1);
      }
      {
        float4 d[MB];
        tmem_wait_ld24(a0);
#pragma unroll
        for (int b = 0; b < MB; ++b) d[b] = pD[b * (KP / 4) + (NCHUNK_T - 1) * KS];
        cells(acc, d, a0);
      }
    } else {
#pragma unroll 2
      for (int c = 0; c < NCHUNK_T; ++c) {
        float4 d[MB];
        float a[24];
        if (MODE == 0) {
          tmem_ld24(tbase + c * 24, a);
        } else {
#pragma unroll
          for (int n = 0; n < NJ; ++n) {
            const float4 v = pA[(jg + 32 * n) * (KP_S / 4) + (c & 15) * KS + q];
            a[n * 4 + 0] = v.x; a[n * 4 + 1] = v.y; a[n * 4 + 2] = v.z; a[n * 4 + 3] = v.w;
          }
        }
#pragma unroll
        for (int b = 0; b < MB; ++b) d[b] = pD[b * (KP / 4) + c * KS];
        if (MODE == 0) tmem_wait_ld24(a);
        cells(acc, d, a);
      }
    }
    if (MODE != 1) {
#pragma unroll
      for (int c = 0; c < NCHUNK_S; ++c) {
        float4 d[MB];
        float a[24];
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
          const float4 v = pA[(jg + 32 * n) * (KP_S / 4) + c * KS + q];
          a[n * 4 + 0] = v.x; a[n * 4 + 1] = v.y; a[n * 4 + 2] = v.z; a[n * 4 + 3] = v.w;
        }
#pragma unroll
        for (int b = 0; b < MB; ++b) d[b] = pD[b * (KP / 4) + (NCHUNK_T + c) * KS];
        cells(acc, d, a);
      }
    }
    // keep steps dependent so nothing is hoisted: fold a (never true) data dependence into the next step
    bias += acc[0] > 1e30f ? 1.f : 0.f;
    if (bias != 0.f) acc[0] += bias;
  }
  const long long t1 = clock64();

  // verification (MODE 0 / 2): max over the whole K range of this lane's K-split, recomputed from the generators
  int bad = 0;
  if (MODE != 1) {
#pragma unroll 1
    for (int b = 0; b < MB; ++b)
#pragma unroll 1
      for (int n = 0; n < NJ; ++n) {
        float ref = -INFINITY;
        for (int c = 0; c < NCHUNK; ++c)
          for (int kk = 0; kk < 4; ++kk) {
            const int k = 16 * c + 4 * q + kk;
            ref = fmaxf(ref, __fadd_rn(d_value(pipe, b, k), a_value(jg + 32 * n, k)));
          }
        float got = -INFINITY;
#pragma unroll
        for (int i = 0; i < MB * NJ; ++i) if (i == b * NJ + n) got = acc[i];
        if (got != ref) ++bad;
      }
    if (bad) atomicAdd(mismatches, bad);
  }
  float ssum = 0.f;
#pragma unroll
  for (int i = 0; i < MB * NJ; ++i) ssum += acc[i];
  if (ssum == 123.456f) sink[tid] = ssum;
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem_base), "n"(512));
}

template <int MODE>
static void run(const char* name, int sms, int steps) {
  constexpr int KP_S = (MODE == 1) ? 16 * 16 + 16 : NCHUNK_S * 4 * KS + 16;
  const size_t smem = (size_t)(2 * MB * KP + NSLOT * KP_S) * sizeof(float);
  auto kern = tmem_kernel<MODE>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long* d_cyc; int* d_bad; float* d_sink;
  CK(cudaMalloc(&d_cyc, sms * sizeof(long long))); CK(cudaMalloc(&d_bad, sizeof(int))); CK(cudaMalloc(&d_sink, 1024 * sizeof(float)));
  CK(cudaMemset(d_bad, 0, sizeof(int)));
  kern<<<sms, THREADS, smem>>>(steps / 4 + 1, d_cyc, d_bad, d_sink);
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(d_bad, 0, sizeof(int)));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  kern<<<sms, THREADS, smem>>>(steps, d_cyc, d_bad, d_sink);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(sms); int bad = 0;
  CK(cudaMemcpy(cyc.data(), d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
  long long cmax = 0; for (auto c : cyc) if (c > cmax) cmax = c;
  const int nch = (MODE == 1) ? NCHUNK_T : NCHUNK;
  const double cells_per_cta = (double)steps * nch * (MB * NJ * 4) * THREADS;
  printf("{\"test\": \"%s\", \"threads\": %d, \"smem\": %zu, \"steps\": %d, \"ms\": %.3f, \"cycles_per_step\": %.0f, "
         "\"cells_per_clk_per_sm\": %.2f, \"mismatches\": %d}\n", name, THREADS, smem, steps, ms, (double)cmax / steps,
         cells_per_cta / (double)cmax, bad);
  fflush(stdout);
  cudaFree(d_cyc); cudaFree(d_bad); cudaFree(d_sink);
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int steps = argc > 1 ? atoi(argv[1]) : 2000;
  run<0>("tmem_7x6_k4", sms, steps);
  run<2>("tmem_7x6_k4_prefetch", sms, steps);
  run<1>("smem_7x6_k4", sms, steps);
  return 0;
}
