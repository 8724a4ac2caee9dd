// Register-only microbenchmark of instruction ORDER for the packed max-plus tile (8 clips x 4 targets x 4 sources):
// how many cycles per (FADD2, FMNMX3) pair does the SM need depending on how the two kinds are interleaved, i.e. on
// whether the FADD2 64-bit operands can be served from the operand-reuse cache.  asm volatile pins the order.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_order microbench_order.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

#define FADD2(rx, ry, ax, ay, bx, by) \
  asm volatile("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t" \
               "add.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}" \
               : "=f"(rx), "=f"(ry) : "f"(ax), "f"(ay), "f"(bx), "f"(by))
#define FMAX3(r, a, b, c) asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c))
#define FADD1(r, a, b) asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b))
#define FMAX2(r, a, b) asm volatile("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b))

constexpr int MB = 8, NJ = 4;

template <int ORDER>
__global__ void __launch_bounds__(384, 1) order_kernel(float* out, long long* cycles, int iters, float seed) {
  float4 d[MB], a[NJ];
  float acc[MB][NJ];
#pragma unroll
  for (int b = 0; b < MB; ++b) d[b] = make_float4(seed * b, seed + b, seed - b, seed * threadIdx.x);
#pragma unroll
  for (int n = 0; n < NJ; ++n) a[n] = make_float4(seed * n, seed + n, seed - n, seed * 0.5f);
#pragma unroll
  for (int b = 0; b < MB; ++b)
#pragma unroll
    for (int n = 0; n < NJ; ++n) acc[b][n] = -1e30f;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (ORDER == 0) {          // alternate: per (b,n): 2 FADD2 then 2 FMNMX3
#pragma unroll
      for (int b = 0; b < MB; ++b)
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
          float v0, v1, v2, v3;
          FADD2(v0, v1, d[b].x, d[b].y, a[n].x, a[n].y);
          FADD2(v2, v3, d[b].z, d[b].w, a[n].z, a[n].w);
          FMAX3(acc[b][n], acc[b][n], v0, v1);
          FMAX3(acc[b][n], acc[b][n], v2, v3);
        }
    } else if (ORDER == 1) {   // per b: 4 FADD2 sharing d[b].xy, 4 sharing d[b].zw, then 8 FMNMX3
#pragma unroll
      for (int b = 0; b < MB; ++b) {
        float v[NJ][4];
#pragma unroll
        for (int n = 0; n < NJ; ++n) FADD2(v[n][0], v[n][1], d[b].x, d[b].y, a[n].x, a[n].y);
#pragma unroll
        for (int n = 0; n < NJ; ++n) FADD2(v[n][2], v[n][3], d[b].z, d[b].w, a[n].z, a[n].w);
#pragma unroll
        for (int n = 0; n < NJ; ++n) FMAX3(acc[b][n], acc[b][n], v[n][0], v[n][1]);
#pragma unroll
        for (int n = 0; n < NJ; ++n) FMAX3(acc[b][n], acc[b][n], v[n][2], v[n][3]);
      }
    } else if (ORDER == 2) {   // per n: 8 FADD2 sharing a[n].xy (second operand slot), then 8 FMNMX3; then .zw
#pragma unroll
      for (int n = 0; n < NJ; ++n) {
        float v[MB][2];
#pragma unroll
        for (int b = 0; b < MB; ++b) FADD2(v[b][0], v[b][1], d[b].x, d[b].y, a[n].x, a[n].y);
#pragma unroll
        for (int b = 0; b < MB; ++b) FMAX3(acc[b][n], acc[b][n], v[b][0], v[b][1]);
#pragma unroll
        for (int b = 0; b < MB; ++b) FADD2(v[b][0], v[b][1], d[b].z, d[b].w, a[n].z, a[n].w);
#pragma unroll
        for (int b = 0; b < MB; ++b) FMAX3(acc[b][n], acc[b][n], v[b][0], v[b][1]);
      }
    } else if (ORDER == 3) {   // software pipelined: FADD2 group of step k+1 interleaved 1:1 with FMNMX3 of step k,
                               // FADD2s of a group share d[b].xy / d[b].zw
      float v[NJ][2], w[NJ][2];
#pragma unroll
      for (int n = 0; n < NJ; ++n) FADD2(v[n][0], v[n][1], d[0].x, d[0].y, a[n].x, a[n].y);
#pragma unroll
      for (int h = 0; h < 2 * MB; ++h) {          // half-rows: (b, xy) (b, zw)
        const int b = h >> 1, nb = (h + 1) >> 1;
        const bool nxt_zw = ((h + 1) & 1);
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
          if (h + 1 < 2 * MB) {
            if (nxt_zw) FADD2(w[n][0], w[n][1], d[nb].z, d[nb].w, a[n].z, a[n].w);
            else        FADD2(w[n][0], w[n][1], d[nb].x, d[nb].y, a[n].x, a[n].y);
          }
          FMAX3(acc[b][n], acc[b][n], v[n][0], v[n][1]);
        }
#pragma unroll
        for (int n = 0; n < NJ; ++n) { v[n][0] = w[n][0]; v[n][1] = w[n][1]; }
      }
    } else if (ORDER == 4) {   // scalar FADD + FMNMX, alternating
#pragma unroll
      for (int b = 0; b < MB; ++b)
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
          float v0, v1, v2, v3;
          FADD1(v0, d[b].x, a[n].x); FADD1(v1, d[b].y, a[n].y); FADD1(v2, d[b].z, a[n].z); FADD1(v3, d[b].w, a[n].w);
          FMAX2(acc[b][n], acc[b][n], v0); FMAX2(acc[b][n], acc[b][n], v1);
          FMAX2(acc[b][n], acc[b][n], v2); FMAX2(acc[b][n], acc[b][n], v3);
        }
    } else if (ORDER == 5) {   // scalar FADD + FMNMX3 (2 scalar adds feed one 3-input max)
#pragma unroll
      for (int b = 0; b < MB; ++b)
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
          float v0, v1, v2, v3;
          FADD1(v0, d[b].x, a[n].x); FADD1(v1, d[b].y, a[n].y);
          FMAX3(acc[b][n], acc[b][n], v0, v1);
          FADD1(v2, d[b].z, a[n].z); FADD1(v3, d[b].w, a[n].w);
          FMAX3(acc[b][n], acc[b][n], v2, v3);
        }
    } else if (ORDER == 6) {   // all 64 FADD2 first (grouped by d operand), then all 64 FMNMX3
      float v[MB][NJ][4];
#pragma unroll
      for (int b = 0; b < MB; ++b) {
#pragma unroll
        for (int n = 0; n < NJ; ++n) FADD2(v[b][n][0], v[b][n][1], d[b].x, d[b].y, a[n].x, a[n].y);
#pragma unroll
        for (int n = 0; n < NJ; ++n) FADD2(v[b][n][2], v[b][n][3], d[b].z, d[b].w, a[n].z, a[n].w);
      }
#pragma unroll
      for (int b = 0; b < MB; ++b)
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
          FMAX3(acc[b][n], acc[b][n], v[b][n][0], v[b][n][1]);
          FMAX3(acc[b][n], acc[b][n], v[b][n][2], v[b][n][3]);
        }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int b = 0; b < MB; ++b)
#pragma unroll
    for (int n = 0; n < NJ; ++n) s += acc[b][n];
  if (s == 123.456f) out[threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ORDER>
static void run(const char* name, int num_sms, int threads, int iters, float* d_out, long long* d_cyc) {
  order_kernel<ORDER><<<num_sms, threads>>>(d_out, d_cyc, iters / 8, 1.0f);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  order_kernel<ORDER><<<num_sms, threads>>>(d_out, d_cyc, iters, 1.0f);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> cyc(num_sms);
  CK(cudaMemcpy(cyc.data(), d_cyc, num_sms * sizeof(long long), cudaMemcpyDeviceToHost));
  long long cmax = 0; for (auto c : cyc) if (c > cmax) cmax = c;
  double cells = (double)iters * MB * NJ * 4 * threads;
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, order_kernel<ORDER>));
  printf("{\"order\": \"%s\", \"threads\": %d, \"regs\": %d, \"ms\": %.3f, \"cells_per_clk_per_sm\": %.2f, "
         "\"cycles_per_warp_pair\": %.3f}\n", name, threads, fa.numRegs, ms, cells / cmax,
         (double)cmax / ((double)iters * MB * NJ * 2) / (threads / 32 / 4.0));
  fflush(stdout);
}

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int num_sms = prop.multiProcessorCount;
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 4096)); CK(cudaMalloc(&d_cyc, num_sms * 8));
  int iters = argc > 1 ? atoi(argv[1]) : 4000;
  for (int th : {128, 256, 384}) {
    run<0>("alternate_per_cell", num_sms, th, iters, d_out, d_cyc);
    run<1>("group_by_clip_4add2_then_4max3", num_sms, th, iters, d_out, d_cyc);
    run<2>("group_by_target_8add2_then_8max3", num_sms, th, iters, d_out, d_cyc);
    run<3>("pipelined_1to1_reuse_groups", num_sms, th, iters, d_out, d_cyc);
    run<4>("scalar_fadd_fmnmx", num_sms, th, iters, d_out, d_cyc);
    run<5>("scalar_fadd_fmnmx3", num_sms, th, iters, d_out, d_cyc);
    run<6>("all_add2_then_all_max3", num_sms, th, iters, d_out, d_cyc);
  }
  return 0;
}
