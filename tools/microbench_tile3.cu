// Third tile prototype: bigger register tiles (MB x NJ up to 8 x 8) with narrower operand vectors (VEC = 2 or 4
// floats per lane per chunk) to cut shared-memory -> register traffic (the 128 B/clk/SM return path), K split over KS
// adjacent lanes, recursive-halving shuffle reduction that leaves MB*NJ/KS finished outputs per lane.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_tile3 microbench_tile3.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int VEC> struct Vec;
template <> struct Vec<2> { using T = float2; };
template <> struct Vec<4> { using T = float4; };

template <int VEC> __device__ __forceinline__ float vget(const typename Vec<VEC>::T& v, int i);
template <> __device__ __forceinline__ float vget<2>(const float2& v, int i) { return i == 0 ? v.x : v.y; }
template <> __device__ __forceinline__ float vget<4>(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// lane = tl * KS + q : tl = tile within warp (TPW = 32 / KS tiles, same clip group, interleaved targets), q = K split.
// tile tl of target-quad jq owns targets j = jq*(TPW*NJ) + tl + TPW*n.
template <int MB, int NJ, int KS, int VEC, int BG, int JQ, int UNR>
__global__ void __launch_bounds__(BG * JQ * 32, 1)
tile3_kernel(const float* __restrict__ gA, const float* __restrict__ gD, float* __restrict__ gOut,
             long long* cycles, int NC, int KP, int steps) {
  using V = typename Vec<VEC>::T;
  constexpr int TPW = 32 / KS, MC = MB * BG, THREADS = BG * JQ * 32, NOUT = MB * NJ / KS;
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;                    // [NC][KP]
  float* sD0 = sA + NC * KP;           // [MC][KP]
  float* sD1 = sD0 + MC * KP;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int q = lane % KS, tl = lane / KS;
  const int bg = w / JQ, jq = w % JQ;
  for (int x = tid; x < NC * KP; x += THREADS) sA[x] = gA[x];
  for (int x = tid; x < MC * KP; x += THREADS) { sD0[x] = gD[x]; sD1[x] = gD[x]; }
  __syncthreads();
  const int KPV = KP / VEC;
  int arow[NJ];
#pragma unroll
  for (int n = 0; n < NJ; ++n) arow[n] = min(jq * (TPW * NJ) + tl + TPW * n, NC - 1) * KPV + q;
  const int nchunks = KPV / KS;
  long long t0 = clock64();
  for (int step = 0; step < steps; ++step) {
    const float* sD = (step & 1) ? sD1 : sD0;
    float* sDn = (step & 1) ? sD0 : sD1;
    float acc[MB][NJ];
#pragma unroll
    for (int b = 0; b < MB; ++b)
#pragma unroll
      for (int n = 0; n < NJ; ++n) acc[b][n] = -INFINITY;
    const V* pD = reinterpret_cast<const V*>(sD) + (bg * MB) * KPV + q;
    const V* pA = reinterpret_cast<const V*>(sA);
#pragma unroll UNR
    for (int c = 0; c < nchunks; ++c) {
      V d[MB], a[NJ];
#pragma unroll
      for (int b = 0; b < MB; ++b) d[b] = pD[b * KPV + c * KS];
#pragma unroll
      for (int n = 0; n < NJ; ++n) a[n] = pA[arow[n] + c * KS];
#pragma unroll
      for (int b = 0; b < MB; ++b)
#pragma unroll
        for (int n = 0; n < NJ; ++n)
#pragma unroll
          for (int v = 0; v < VEC; ++v)
            acc[b][n] = fmaxf(acc[b][n], __fadd_rn(vget<VEC>(d[b], v), vget<VEC>(a[n], v)));
    }
    // recursive halving over the KS lanes: after log2(KS) rounds lane q holds NOUT finished maxima
    float* flat = &acc[0][0];
    constexpr int TOT = MB * NJ;
    int len = TOT;
#pragma unroll
    for (int off = KS / 2; off >= 1; off >>= 1) {
      const bool upper = (q & off) != 0;
      len >>= 1;
#pragma unroll
      for (int i = 0; i < TOT / 2; ++i) {
        if (i < len) {
          const float keep = upper ? flat[i + len] : flat[i];
          const float send = upper ? flat[i] : flat[i + len];
          const float got = __shfl_xor_sync(0xffffffffu, send, off);
          flat[i] = fmaxf(keep, got);
        }
      }
    }
    // lane's outputs: flat[0..NOUT): write them somewhere plausible
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      const int j = jq * (TPW * NJ) + tl + TPW * (k % NJ);
      const int b = bg * MB + (q % MB);
      if (j < NC) sDn[b * KP + j] = flat[k] * 0.25f;
    }
    __syncthreads();
  }
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (gOut) for (int x = tid; x < MC * KP; x += THREADS) gOut[blockIdx.x * MC * KP + x] = sD0[x];
}

template <int MB, int NJ, int KS, int VEC, int BG, int JQ, int UNR>
static void run(const char* name, int num_sms, int NC, int KP, int S, int steps, const float* dA, const float* dD,
                float* dOut, long long* dCyc) {
  constexpr int MC = MB * BG, TH = BG * JQ * 32;
  size_t smem = (size_t)(NC + 2 * MC) * KP * sizeof(float);
  auto kern = tile3_kernel<MB, NJ, KS, VEC, BG, JQ, UNR>;
  if (smem > 227 * 1024) { printf("{\"tile3\": \"%s\", \"skipped\": \"smem %zu\"}\n", name, smem); return; }
  if (KP % (VEC * KS)) { printf("{\"tile3\": \"%s\", \"skipped\": \"KP %% chunk\"}\n", name); return; }
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
  kern<<<num_sms, TH, smem>>>(dA, dD, dOut, dCyc, NC, KP, 8);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  kern<<<num_sms, TH, smem>>>(dA, dD, dOut, dCyc, NC, KP, steps);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(num_sms);
  CK(cudaMemcpy(cyc.data(), dCyc, num_sms * sizeof(long long), cudaMemcpyDeviceToHost));
  long long cmax = 0; for (auto c : cyc) if (c > cmax) cmax = c;
  double cells = (double)steps * MC * NC * S;
  printf("{\"tile3\": \"%s\", \"MB\": %d, \"NJ\": %d, \"KS\": %d, \"VEC\": %d, \"threads\": %d, \"regs\": %d, \"spill\": %zu, "
         "\"smem\": %zu, \"KP\": %d, \"cycles_per_step\": %.0f, \"useful_cells_per_clk_per_sm\": %.2f}\n",
         name, MB, NJ, KS, VEC, TH, fa.numRegs, (size_t)fa.localSizeBytes, smem, KP, (double)cmax / steps, cells / cmax);
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
  // NC = 91 real targets (4-CTA cluster); 96 target slots per CTA = JQ * TPW * NJ
  //  MB NJ KS VEC BG JQ UNR
  run<8, 8, 8, 2, 4, 3, 1>("8x8_ks8_v2_384thr_u1", num_sms, 91, 368, S, steps, dA, dD, dOut, dCyc);
  run<8, 8, 8, 2, 4, 3, 2>("8x8_ks8_v2_384thr_u2", num_sms, 91, 368, S, steps, dA, dD, dOut, dCyc);
  run<8, 8, 4, 4, 4, 3, 1>("8x8_ks4_v4_192thr??", num_sms, 91, 368, S, steps, dA, dD, dOut, dCyc);
  run<8, 4, 8, 2, 4, 6, 2>("8x4_ks8_v2_768thr_u2", num_sms, 91, 368, S, steps, dA, dD, dOut, dCyc);
  run<8, 4, 4, 4, 4, 3, 2>("8x4_ks4_v4_384thr_u2(v1)", num_sms, 91, 368, S, steps, dA, dD, dOut, dCyc);
  run<8, 4, 4, 2, 4, 3, 2>("8x4_ks4_v2_384thr_u2", num_sms, 91, 368, S, steps, dA, dD, dOut, dCyc);
  run<8, 6, 8, 2, 4, 4, 2>("8x6_ks8_v2_512thr_u2", num_sms, 91, 368, S, steps, dA, dD, dOut, dCyc);
  run<16, 4, 8, 2, 2, 6, 2>("16x4_ks8_v2_384thr_u2", num_sms, 91, 368, S, steps, dA, dD, dOut, dCyc);
  run<8, 8, 16, 2, 4, 6, 2>("8x8_ks16_v2_768thr_u2", num_sms, 91, 384, S, steps, dA, dD, dOut, dCyc);
  run<4, 8, 8, 2, 8, 3, 2>("4x8_ks8_v2_768thr_u2", num_sms, 91, 368, S, steps, dA, dD, dOut, dCyc);
  return 0;
}
