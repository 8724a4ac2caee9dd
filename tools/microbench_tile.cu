// Prototype of the SMEM-resident register-tiled max-plus step (no cluster exchange, no HBM traffic):
// how close does a realistic inner loop (LDS.128 operand feeds + FADD2/FMNMX3) get to the pipe ceiling
// measured by microbench_pipes?  One CTA per SM; logA^T shard [NC rows][R] and delta [MC][R] live in SMEM.
//
// thread tile = MB sequences x NJ target states, K (source state) split KS ways across adjacent lanes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_tile microbench_tile.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void add2(float& rx, float& ry, float ax, float ay, float bx, float by) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rc, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(rx), "=f"(ry) : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}

// S: states, R: padded row length in floats (R % 32 == 16 keeps the LDS.128 of 8 lanes x {2 rows x 4 k-splits}
// conflict-free), NC: target states owned by this CTA, MC = MB*BG sequences.
template <int MB, int NJ, int KS, int BG, int JG, bool PACKED>
__global__ void __launch_bounds__(BG * JG * KS, 1)
tile_kernel(const float* __restrict__ gA, const float* __restrict__ gD, float* __restrict__ gOut,
            long long* cycles, int S, int R, int steps) {
  constexpr int NC = NJ * JG;
  constexpr int MC = MB * BG;
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;                 // [NC][R]
  float* sD0 = sA + NC * R;         // [MC][R]
  float* sD1 = sD0 + MC * R;        // [MC][R]
  const int tid = threadIdx.x;
  for (int x = tid; x < NC * R; x += blockDim.x) sA[x] = gA[x];
  for (int x = tid; x < MC * R; x += blockDim.x) { sD0[x] = gD[x]; sD1[x] = gD[x]; }
  __syncthreads();

  const int q = tid % KS;
  const int jg = (tid / KS) % JG;
  const int bg = tid / (KS * JG);
  const int nchunks = R / (4 * KS);
  long long t0 = clock64();
  for (int step = 0; step < steps; ++step) {
    const float* sD = (step & 1) ? sD1 : sD0;
    float* sDn = (step & 1) ? sD0 : sD1;
    float acc[MB][NJ];
#pragma unroll
    for (int b = 0; b < MB; ++b)
#pragma unroll
      for (int n = 0; n < NJ; ++n) acc[b][n] = -INFINITY;
    const float4* pD = reinterpret_cast<const float4*>(sD + (bg * MB) * R) + q;
    const float4* pA = reinterpret_cast<const float4*>(sA + jg * R) + q;
#pragma unroll 2
    for (int c = 0; c < nchunks; ++c) {
      float4 d[MB], a[NJ];
#pragma unroll
      for (int b = 0; b < MB; ++b) d[b] = pD[b * (R / 4) + c * KS];
#pragma unroll
      for (int n = 0; n < NJ; ++n) a[n] = pA[n * JG * (R / 4) + c * KS];
#pragma unroll
      for (int b = 0; b < MB; ++b)
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
          if (PACKED) {
            float v0, v1, v2, v3;
            add2(v0, v1, d[b].x, d[b].y, a[n].x, a[n].y);
            add2(v2, v3, d[b].z, d[b].w, a[n].z, a[n].w);
            acc[b][n] = max3(acc[b][n], v0, v1);
            acc[b][n] = max3(acc[b][n], v2, v3);
          } else {
            acc[b][n] = fmaxf(acc[b][n], d[b].x + a[n].x);
            acc[b][n] = fmaxf(acc[b][n], d[b].y + a[n].y);
            acc[b][n] = fmaxf(acc[b][n], d[b].z + a[n].z);
            acc[b][n] = fmaxf(acc[b][n], d[b].w + a[n].w);
          }
        }
    }
    // K-split reduction across the KS adjacent lanes
#pragma unroll
    for (int b = 0; b < MB; ++b)
#pragma unroll
      for (int n = 0; n < NJ; ++n) {
#pragma unroll
        for (int off = 1; off < KS; off <<= 1)
          acc[b][n] = fmaxf(acc[b][n], __shfl_xor_sync(0xffffffffu, acc[b][n], off));
      }
    // each of the KS lanes writes a share of the MB*NJ outputs
#pragma unroll
    for (int b = 0; b < MB; ++b)
#pragma unroll
      for (int n = 0; n < NJ; ++n) {
        if (((b * NJ + n) % KS) == q) {
          int j = jg + n * JG;
          if (j < S) sDn[(bg * MB + b) * R + j] = acc[b][n] * 0.25f;   // keep values bounded
        }
      }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (gOut) for (int x = tid; x < MC * R; x += blockDim.x) gOut[blockIdx.x * MC * R + x] = sD0[x];
}

template <int MB, int NJ, int KS, int BG, int JG, bool PACKED>
static void run(const char* name, int num_sms, int S, int steps, const float* dA, const float* dD, float* dOut, long long* dCyc) {
  constexpr int NC = NJ * JG, MC = MB * BG, TH = BG * JG * KS;
  int R = ((S + 4 * KS - 1) / (4 * KS)) * (4 * KS);
  if (KS < 8 && ((R / (4 * KS)) % 2 == 0)) R += 4 * KS;       // odd multiple of 4*KS words: conflict-free LDS.128
  size_t smem = (size_t)(NC + 2 * MC) * R * sizeof(float);
  auto kern = tile_kernel<MB, NJ, KS, BG, JG, PACKED>;
  if (smem > 227 * 1024) { printf("{\"tile\": \"%s\", \"skipped\": \"smem %zu\"}\n", name, smem); return; }
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
  kern<<<num_sms, TH, smem>>>(dA, dD, dOut, dCyc, S, R, 8);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  kern<<<num_sms, TH, smem>>>(dA, dD, dOut, dCyc, S, R, steps);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(num_sms);
  CK(cudaMemcpy(cyc.data(), dCyc, num_sms * sizeof(long long), cudaMemcpyDeviceToHost));
  long long cmax = 0; for (auto c : cyc) if (c > cmax) cmax = c;
  double cells = (double)steps * MC * (double)(NC < S ? NC : S) * S;   // useful cells per CTA
  printf("{\"tile\": \"%s\", \"MB\": %d, \"NJ\": %d, \"KS\": %d, \"BG\": %d, \"JG\": %d, \"packed\": %d, \"threads\": %d, "
         "\"regs\": %d, \"smem\": %zu, \"R\": %d, \"MC\": %d, \"NC\": %d, \"ms\": %.3f, \"cycles_per_step\": %.0f, "
         "\"useful_cells_per_clk_per_sm\": %.2f, \"padded_cells_per_clk_per_sm\": %.2f, \"sm_mhz_est\": %.0f}\n",
         name, MB, NJ, KS, BG, JG, (int)PACKED, TH, fa.numRegs, smem, R, MC, NC, ms, (double)cmax / steps,
         cells / cmax, (double)steps * MC * NC * R / cmax, cmax / (ms * 1e-3) / 1e6);
  fflush(stdout);
}

int main(int argc, char** argv) {
  int only_first = argc > 1 ? atoi(argv[1]) : 0;
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int num_sms = prop.multiProcessorCount;
  const int S = 361; const int steps = only_first ? 50 : 400;
  size_t nA = 192 * 416, nD = 64 * 416;
  std::vector<float> hA(nA), hD(nD);
  srand(1);
  for (auto& v : hA) v = -(float)(rand() % 1000) / 64.f;
  for (auto& v : hD) v = -(float)(rand() % 1000) / 64.f;
  float *dA, *dD, *dOut; long long* dCyc;
  CK(cudaMalloc(&dA, nA * 4)); CK(cudaMalloc(&dD, nD * 4)); CK(cudaMalloc(&dOut, (size_t)num_sms * 64 * 400 * 4));
  CK(cudaMalloc(&dCyc, num_sms * 8));
  CK(cudaMemcpy(dA, hA.data(), nA * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, hD.data(), nD * 4, cudaMemcpyHostToDevice));
  //   MB NJ KS BG JG packed
  run<7, 4, 4, 4, 24, true >("7x4_k4_packed", num_sms, S, steps, dA, dD, dOut, dCyc);
  run<7, 4, 4, 4, 24, false>("7x4_k4_scalar", num_sms, S, steps, dA, dD, dOut, dCyc);
  if (only_first) return 0;
  run<7, 8, 8, 4, 12, true >("7x8_k8_packed", num_sms, S, steps, dA, dD, dOut, dCyc);
  run<7, 8, 8, 4, 12, false>("7x8_k8_scalar", num_sms, S, steps, dA, dD, dOut, dCyc);
  run<14, 4, 8, 2, 24, true >("14x4_k8_packed", num_sms, S, steps, dA, dD, dOut, dCyc);
  run<7, 4, 8, 4, 24, true >("7x4_k8_packed_768thr", num_sms, S, steps, dA, dD, dOut, dCyc);
  run<7, 6, 4, 4, 16, true >("7x6_k4_packed_256thr", num_sms, S, steps, dA, dD, dOut, dCyc);
  run<4, 8, 4, 7, 12, true >("4x8_k4_packed_336thr", num_sms, S, steps, dA, dD, dOut, dCyc);
  run<7, 4, 2, 4, 24, true >("7x4_k2_packed_192thr", num_sms, S, steps, dA, dD, dOut, dCyc);
  run<7, 2, 4, 4, 48, true >("7x2_k4_packed_768thr", num_sms, S, steps, dA, dD, dOut, dCyc);
  // C=2 cluster shapes (14 sequences, 181 targets): smem too large for fp32 residency -> skipped automatically
  run<7, 4, 4, 2, 46, true >("C2_7x4_k4", num_sms, S, steps, dA, dD, dOut, dCyc);
  // C=8 cluster shapes (56 sequences, 46 targets)
  run<7, 4, 4, 8, 12, true >("C8_7x4_k4", num_sms, S, steps, dA, dD, dOut, dCyc);
  run<14, 4, 8, 4, 12, true >("C8_14x4_k8", num_sms, S, steps, dA, dD, dOut, dCyc);
  return 0;
}
