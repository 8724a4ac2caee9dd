// Issue-rate microbenchmarks for the max-plus inner loop on sm_100a.
//
// Purpose: measure the roofline denominator for the Viterbi recursion
// (one "cell" = fl32(delta[i] + logA[i,j]) followed by a max), i.e. how many
// cells/clk/SM the FP32 pipes sustain with
//   * scalar  FADD + FMNMX              (BASELINE.md definition: 2 issue slots / cell)
//   * packed  FADD2 (add.f32x2) + FMNMX3 (3-input max)   (1 issue slot / cell)
//   * the direct (value,index) tracking form FADD + FSETP + FSEL + SEL
// and the cluster occupancy the resident-logA design can get.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_pipes microbench_pipes.cu
// Run  : ./microbench_pipes            (prints one JSON object per test)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int NACC = 16;      // independent chains per thread
constexpr int UNROLL = 4;

enum Test { T_FADD = 0, T_FMNMX, T_CELL, T_FADD2, T_FMNMX3, T_CELL2, T_CELL_IDX, T_CELL2_MIX, T_FMNMX3_INT, T_FADD2x2_FMNMX3, T_FADD2_FMNMX3x2, T_COUNT };
static const char* kNames[T_COUNT] = {"fadd", "fmnmx", "cell_fadd_fmnmx", "fadd2", "fmnmx3", "cell2_fadd2_fmnmx3",
                                      "cell_idx_fsetp_sel", "cell2_fadd2_2xfmnmx", "vimnmx3_s32", "mix_2fadd2_1fmnmx3", "mix_1fadd2_2fmnmx3"};
// lane-ops (useful scalar results) per inner statement for each test
static const double kCellsPerStmt[T_COUNT] = {1, 1, 1, 2, 2, 2, 1, 2, 2, 3, 3};  // mixes: instructions per stmt

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
               "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
               "add.rn.f32x2 rc, ra, rb;\n\t"
               "mov.b64 {%0, %1}, rc;\n\t}"
               : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float fadd(float a, float b) {
  float r; asm("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ float fmx(float a, float b) {
  float r; asm("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ int imax3(int a, int b, int c) {
  int r; asm("max.s32 %0, %1, %2;\n\tmax.s32 %0, %0, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}

template <int TEST>
__global__ void __launch_bounds__(1024) pipe_kernel(float* out, long long* cycles, int iters, float seed) {
  float acc[NACC], d[NACC], e[NACC];
  int idx[NACC];
#pragma unroll
  for (int k = 0; k < NACC; ++k) {
    acc[k] = seed * (float)(threadIdx.x + k);
    d[k] = seed + (float)k;
    e[k] = seed - (float)k;
    idx[k] = 0;
  }
  float a = seed, b = seed * 0.5f;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (TEST == T_CELL || TEST == T_CELL2 || TEST == T_CELL_IDX || TEST == T_CELL2_MIX || TEST == T_FADD2x2_FMNMX3 || TEST == T_FADD2_FMNMX3x2) {
        a = fadd(a, seed);  // loop-variant operand (1 extra FADD per NACC statements)
        b = fadd(b, seed);
      }
#pragma unroll
      for (int k = 0; k < NACC; ++k) {
        if (TEST == T_FADD) {
          acc[k] = fadd(acc[k], d[k]);
        } else if (TEST == T_FMNMX) {
          // alternate max/min so ptxas cannot fuse two scalar FMNMX into one FMNMX3
          if (u & 1) acc[k] = fmx(acc[k], d[k]);
          else asm("min.f32 %0, %1, %2;" : "=f"(acc[k]) : "f"(acc[k]), "f"(e[k]));
        } else if (TEST == T_CELL) {
          float v = fadd(d[k], a);
          acc[k] = fmx(acc[k], v);
        } else if (TEST == T_FADD2) {
          float2 r = add2(make_float2(acc[k], e[k]), make_float2(d[k], d[k]));
          acc[k] = r.x; e[k] = r.y;
        } else if (TEST == T_FMNMX3) {
          acc[k] = max3(acc[k], d[k], e[k]);
        } else if (TEST == T_CELL2) {
          float2 v = add2(make_float2(d[k], e[k]), make_float2(a, b));
          acc[k] = max3(acc[k], v.x, v.y);
        } else if (TEST == T_CELL2_MIX) {
          float2 v = add2(make_float2(d[k], e[k]), make_float2(a, b));
          acc[k] = fmx(fmx(acc[k], v.x), v.y);
        } else if (TEST == T_CELL_IDX) {
          float v = fadd(d[k], a);
          bool p = v > acc[k];
          acc[k] = p ? v : acc[k];
          idx[k] = p ? (it * UNROLL + u) : idx[k];
        } else if (TEST == T_FADD2x2_FMNMX3) {
          float2 v = add2(make_float2(d[k], e[k]), make_float2(a, b));
          float2 w = add2(make_float2(e[k], d[k]), make_float2(b, a));
          acc[k] = max3(acc[k], v.x + w.x, v.y + w.y) ;
        } else if (TEST == T_FADD2_FMNMX3x2) {
          float2 v = add2(make_float2(d[k], e[k]), make_float2(a, b));
          acc[k] = max3(acc[k], v.x, v.y);
          e[k] = max3(e[k], v.y, acc[k]);
        } else if (TEST == T_FMNMX3_INT) {
          idx[k] = imax3(idx[k], __float_as_int(d[k]), __float_as_int(e[k]) + it);
        }
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f; int si = 0;
#pragma unroll
  for (int k = 0; k < NACC; ++k) { s += acc[k] + e[k]; si += idx[k]; }
  if (s == 123.456f || si == -12345) out[threadIdx.x] = s + a + b;  // keep results live
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int TEST>
static void run(int num_sms, int threads, int blocks_per_sm, int iters, float* d_out, long long* d_cyc) {
  int blocks = num_sms * blocks_per_sm;
  pipe_kernel<TEST><<<blocks, threads>>>(d_out, d_cyc, iters / 8, 1.0f);  // warm-up
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  pipe_kernel<TEST><<<blocks, threads>>>(d_out, d_cyc, iters, 1.0f);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(blocks);
  CK(cudaMemcpy(cyc.data(), d_cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
  double cyc_avg = 0; long long cyc_max = 0;
  for (auto c : cyc) { cyc_avg += (double)c; if (c > cyc_max) cyc_max = c; }
  cyc_avg /= blocks;
  double stmts = (double)iters * UNROLL * NACC;                 // per thread
  double lane_ops = stmts * kCellsPerStmt[TEST] * threads * blocks;
  double mhz = (double)cyc_max / (ms * 1e-3) / 1e6;             // SM clock seen by the longest block
  double per_clk_sm = stmts * kCellsPerStmt[TEST] * threads * blocks_per_sm / (double)cyc_max;
  printf("{\"test\": \"%s\", \"threads\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"cycles_max\": %lld, "
         "\"cycles_avg\": %.0f, \"sm_mhz_est\": %.0f, \"results_per_clk_per_sm\": %.2f, \"Tresults_per_s\": %.3f}\n",
         kNames[TEST], threads, blocks_per_sm, ms, cyc_max, cyc_avg, mhz, per_clk_sm, lane_ops / (ms * 1e-3) / 1e12);
  fflush(stdout);
}

__global__ void dummy_cluster_kernel(float* p) { extern __shared__ float sm[]; if (p) p[0] = sm[0]; }

static void cluster_probe(int num_sms) {
  const int smem_opts[] = {100 * 1024, 180 * 1024, 200 * 1024, 215 * 1024, 227 * 1024};
  const int csz_opts[] = {1, 2, 3, 4, 5, 6, 7, 8, 12, 16};
  CK(cudaFuncSetAttribute(dummy_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CK(cudaFuncSetAttribute(dummy_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  for (int smem : smem_opts) for (int csz : csz_opts) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csz * 64); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, dummy_cluster_kernel, &cfg);
    printf("{\"probe\": \"max_active_clusters\", \"cluster_size\": %d, \"smem\": %d, \"clusters\": %d, \"ctas\": %d, "
           "\"sms\": %d, \"err\": \"%s\"}\n", csz, smem, nclusters, nclusters * csz, num_sms,
           e == cudaSuccess ? "ok" : cudaGetErrorString(e));
    cudaGetLastError();
  }
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d, \"smem_per_sm\": %zu, \"smem_optin\": %zu, \"l2\": %d}\n",
         prop.name, prop.multiProcessorCount, prop.major, prop.minor, clk_khz, prop.sharedMemPerMultiprocessor,
         prop.sharedMemPerBlockOptin, prop.l2CacheSize);
  int num_sms = prop.multiProcessorCount;
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 1024 * sizeof(float)));
  CK(cudaMalloc(&d_cyc, num_sms * 8 * sizeof(long long)));
  int iters = 20000;
  const int thread_opts[] = {256, 512, 1024};
  if (argc > 1) iters = atoi(argv[1]);
  int only_threads = argc > 2 ? atoi(argv[2]) : 0;
  for (int th : thread_opts) {
    if (only_threads && th != only_threads) continue;
    run<T_FADD>(num_sms, th, 1, iters, d_out, d_cyc);
    run<T_FMNMX>(num_sms, th, 1, iters, d_out, d_cyc);
    run<T_CELL>(num_sms, th, 1, iters, d_out, d_cyc);
    run<T_FADD2>(num_sms, th, 1, iters, d_out, d_cyc);
    run<T_FMNMX3>(num_sms, th, 1, iters, d_out, d_cyc);
    run<T_CELL2>(num_sms, th, 1, iters, d_out, d_cyc);
    run<T_CELL2_MIX>(num_sms, th, 1, iters, d_out, d_cyc);
    run<T_CELL_IDX>(num_sms, th, 1, iters, d_out, d_cyc);
    run<T_FMNMX3_INT>(num_sms, th, 1, iters, d_out, d_cyc);
    run<T_FADD2x2_FMNMX3>(num_sms, th, 1, iters, d_out, d_cyc);
    run<T_FADD2_FMNMX3x2>(num_sms, th, 1, iters, d_out, d_cyc);
  }
  cluster_probe(num_sms);
  return 0;
}
