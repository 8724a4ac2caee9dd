// Second tile prototype: lanes = target states only (delta loads are warp-wide broadcasts: 1 shared-memory wavefront
// per LDS.128 instead of 4), K split across WARPS (KSW) with a shared-memory exchange of the partial maxima.
// Measures the isolated per-step cost (no cluster exchange, no HBM) for several (MB, NJ, KSW, BG) shapes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_tile2 microbench_tile2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// thread (warp w, lane l): bg = w / KSW, kq = w % KSW; targets j = l + 32 n (n < NJ); clips bg*MB .. +MB;
// K range: float4 index f = kq, kq + KSW, ... < KP4
template <int MB, int NJ, int KSW, int BG, int KP4, int UNR>
__global__ void __launch_bounds__(BG * KSW * 32, 1)
tile2_kernel(const float* __restrict__ gA, const float* __restrict__ gD, float* __restrict__ gOut,
             long long* cycles, int NC, int steps) {
  constexpr int MC = MB * BG, KP = KP4 * 4, THREADS = BG * KSW * 32;
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;                    // [NC][KP]
  float* sD0 = sA + NC * KP;           // [MC][KP]
  float* sD1 = sD0;                    // (perf prototype: the double buffer is aliased to make room for sP)
  float* sP = sD0 + MC * KP;           // partial maxima [KSW-1][MC][32*NJ]  (stand-in for the own-slice scratch)
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int bg = w / KSW, kq = w % KSW;
  for (int x = tid; x < NC * KP; x += THREADS) sA[x] = gA[x];
  for (int x = tid; x < MC * KP; x += THREADS) { sD0[x] = gD[x]; sD1[x] = gD[x]; }
  __syncthreads();
  int arow[NJ];
#pragma unroll
  for (int n = 0; n < NJ; ++n) arow[n] = min(lane + 32 * n, NC - 1) * KP4;
  long long t0 = clock64();
  for (int step = 0; step < steps; ++step) {
    const float* sD = (step & 1) ? sD1 : sD0;
    float* sDn = (step & 1) ? sD0 : sD1;
    float acc[MB][NJ];
#pragma unroll
    for (int b = 0; b < MB; ++b)
#pragma unroll
      for (int n = 0; n < NJ; ++n) acc[b][n] = -INFINITY;
    const float4* pD = reinterpret_cast<const float4*>(sD) + (bg * MB) * KP4;
    const float4* pA = reinterpret_cast<const float4*>(sA);
#pragma unroll UNR
    for (int f = kq; f < KP4; f += KSW) {
      float4 d[MB], a[NJ];
#pragma unroll
      for (int b = 0; b < MB; ++b) d[b] = pD[b * KP4 + f];          // warp-uniform address: broadcast
#pragma unroll
      for (int n = 0; n < NJ; ++n) a[n] = pA[arow[n] + f];          // 32 rows, row stride KP4 odd: conflict free
#pragma unroll
      for (int b = 0; b < MB; ++b)
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
          acc[b][n] = fmaxf(acc[b][n], __fadd_rn(d[b].x, a[n].x));
          acc[b][n] = fmaxf(acc[b][n], __fadd_rn(d[b].y, a[n].y));
          acc[b][n] = fmaxf(acc[b][n], __fadd_rn(d[b].z, a[n].z));
          acc[b][n] = fmaxf(acc[b][n], __fadd_rn(d[b].w, a[n].w));
        }
    }
    // cross-warp K reduction: warps kq > 0 publish, warp kq == 0 combines and writes delta_t
    if (KSW > 1) {
      if (kq > 0) {
#pragma unroll
        for (int b = 0; b < MB; ++b)
#pragma unroll
          for (int n = 0; n < NJ; ++n)
            sP[((kq - 1) * MC + bg * MB + b) * (32 * NJ) + lane + 32 * n] = acc[b][n];
      }
      __syncthreads();
      if (kq == 0) {
#pragma unroll
        for (int k = 1; k < KSW; ++k)
#pragma unroll
          for (int b = 0; b < MB; ++b)
#pragma unroll
            for (int n = 0; n < NJ; ++n)
              acc[b][n] = fmaxf(acc[b][n], sP[((k - 1) * MC + bg * MB + b) * (32 * NJ) + lane + 32 * n]);
      }
    }
    if (kq == 0) {
#pragma unroll
      for (int b = 0; b < MB; ++b)
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
          const int j = lane + 32 * n;
          if (j < NC) sDn[(bg * MB + b) * KP + j] = acc[b][n] * 0.25f;
        }
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (gOut) for (int x = tid; x < MC * KP; x += THREADS) gOut[blockIdx.x * MC * KP + x] = sD0[x];
}

template <int MB, int NJ, int KSW, int BG, int KP4, int UNR>
static void run(const char* name, int num_sms, int NC, int S, int steps, const float* dA, const float* dD, float* dOut,
                long long* dCyc) {
  constexpr int MC = MB * BG, KP = KP4 * 4, TH = BG * KSW * 32;
  size_t smem = ((size_t)(NC + MC) * KP + (size_t)(KSW - 1) * MC * 32 * NJ) * sizeof(float);
  auto kern = tile2_kernel<MB, NJ, KSW, BG, KP4, UNR>;
  if (smem > 227 * 1024) { printf("{\"tile2\": \"%s\", \"skipped\": \"smem %zu\"}\n", name, smem); return; }
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
  kern<<<num_sms, TH, smem>>>(dA, dD, dOut, dCyc, NC, 8);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  kern<<<num_sms, TH, smem>>>(dA, dD, dOut, dCyc, NC, steps);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(num_sms);
  CK(cudaMemcpy(cyc.data(), dCyc, num_sms * sizeof(long long), cudaMemcpyDeviceToHost));
  long long cmax = 0; for (auto c : cyc) if (c > cmax) cmax = c;
  double cells = (double)steps * MC * NC * S;     // useful cells per CTA (NC real targets, S real sources)
  printf("{\"tile2\": \"%s\", \"MB\": %d, \"NJ\": %d, \"KSW\": %d, \"BG\": %d, \"KP\": %d, \"unroll\": %d, \"threads\": %d, "
         "\"regs\": %d, \"smem\": %zu, \"MC\": %d, \"NC\": %d, \"cycles_per_step\": %.0f, \"useful_cells_per_clk_per_sm\": %.2f}\n",
         name, MB, NJ, KSW, BG, KP, UNR, TH, fa.numRegs, smem, MC, NC, (double)cmax / steps, cells / cmax);
  fflush(stdout);
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int num_sms = prop.multiProcessorCount;
  const int S = 361, steps = argc > 1 ? atoi(argv[1]) : 300;
  size_t nA = 192 * 420, nD = 64 * 420;
  std::vector<float> hA(nA), hD(nD);
  srand(1);
  for (auto& v : hA) v = -(float)(rand() % 1000) / 64.f;
  for (auto& v : hD) v = -(float)(rand() % 1000) / 64.f;
  float *dA, *dD, *dOut; long long* dCyc;
  CK(cudaMalloc(&dA, nA * 4)); CK(cudaMalloc(&dD, nD * 4)); CK(cudaMalloc(&dOut, (size_t)num_sms * 64 * 420 * 4));
  CK(cudaMalloc(&dCyc, num_sms * 8));
  CK(cudaMemcpy(dA, hA.data(), nA * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, hD.data(), nD * 4, cudaMemcpyHostToDevice));
  // KP4 = 91 -> KP = 364 (row stride odd in float4: conflict-free A loads).  NC = 91 real targets of a 4-CTA cluster.
  //  MB NJ KSW BG KP4 UNR
  run<8, 3, 2, 4, 91, 2>("8x3_ksw2_256thr_u2", num_sms, 91, S, steps, dA, dD, dOut, dCyc);
  run<8, 3, 2, 4, 91, 1>("8x3_ksw2_256thr_u1", num_sms, 91, S, steps, dA, dD, dOut, dCyc);
  run<8, 3, 2, 4, 91, 4>("8x3_ksw2_256thr_u4", num_sms, 91, S, steps, dA, dD, dOut, dCyc);
  run<8, 3, 1, 4, 91, 2>("8x3_ksw1_128thr_u2", num_sms, 91, S, steps, dA, dD, dOut, dCyc);
  run<8, 3, 3, 4, 91, 2>("8x3_ksw3_384thr_u2", num_sms, 91, S, steps, dA, dD, dOut, dCyc);
  run<4, 3, 2, 8, 91, 2>("4x3_ksw2_512thr_u2", num_sms, 91, S, steps, dA, dD, dOut, dCyc);
  run<4, 3, 4, 8, 91, 2>("4x3_ksw4_1024thr_u2", num_sms, 91, S, steps, dA, dD, dOut, dCyc);
  run<16, 3, 2, 2, 91, 2>("16x3_ksw2_128thr_u2", num_sms, 91, S, steps, dA, dD, dOut, dCyc);
  run<16, 3, 4, 2, 91, 2>("16x3_ksw4_256thr_u2", num_sms, 91, S, steps, dA, dD, dOut, dCyc);
  run<8, 3, 4, 4, 91, 2>("8x3_ksw4_512thr_u2", num_sms, 91, S, steps, dA, dD, dOut, dCyc);
  // 2-CTA-cluster shape for comparison (would need logA^T streaming): 16 clips, 181 targets -> NJ = 6
  run<8, 6, 2, 2, 91, 2>("C2shape_8x6_ksw2_128thr", num_sms, 181, S, steps, dA, dD, dOut, dCyc);
  run<8, 6, 4, 2, 91, 2>("C2shape_8x6_ksw4_256thr", num_sms, 181, S, steps, dA, dD, dOut, dCyc);
  return 0;
}
