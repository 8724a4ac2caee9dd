// Shared pieces of the tensor-memory kernels (vit_tmem.cu: max-plus Viterbi; vit_fb.cu: sum-product forward-backward):
// the shard/tile plan, the re-layout of the [S][S] matrix into per-lane TMEM images, and the tcgen05 ld/st helpers.
#pragma once
#include <cstdlib>

#include "vit_common.cuh"

namespace vit {

constexpr int tMB = 7;                       // clips per thread tile = clips per pipeline
constexpr int tKS = 4;                       // K split across adjacent lanes
constexpr int tPipes = 2;
constexpr int tThreads = 256;
constexpr int tPipeThreads = 128;
constexpr int tMaxNJ = 6;
constexpr int tTmemCols = 512;

struct TmemPlan {
  int C;            // CTAs per cluster
  int NJ;           // target states per thread (target slot of thread (jg, n) = jg + 32 n, jg in [0, 32))
  int NCP;          // padded target states per shard; K positions of shard c are [c*NCP, (c+1)*NCP)
  int KP;           // C * NCP, padded K extent (multiple of 16)
  int base, rem;    // shard c owns base + (c < rem) states starting at c*base + min(c, rem)
  int NCmax;
  int nchunk_t;     // K chunks (16 K positions = 4 per K-split lane) served from TMEM
  int nchunk_s;     // K chunks served from shared memory (the tail)
  int tail_stride;  // floats per target row of the shared-memory tail (== 16 mod 32: conflict-free LDS.128)
};

// min_pad: spare K positions wanted at the end of every shard's slice (the forward-backward kernel parks its partial
// normaliser sums there)
static bool try_tmem_plan(int S, int C, TmemPlan* p, int min_pad) {
  const int NCmax = (S + C - 1) / C;
  if (NCmax > 32 * tMaxNJ) return false;
  p->C = C;
  p->base = S / C;
  p->rem = S % C;
  p->NCmax = NCmax;
  p->NJ = (NCmax + 31) / 32;
  int ncp = (NCmax + min_pad + 3) / 4 * 4;
  while ((C * ncp) % 16 != 0) ncp += 4;
  p->NCP = ncp;
  p->KP = C * ncp;
  const int nchunks = p->KP / 16;
  p->nchunk_t = tTmemCols / (p->NJ * 4);
  if (p->nchunk_t > nchunks) p->nchunk_t = nchunks;
  p->nchunk_s = nchunks - p->nchunk_t;
  const int tail_k = p->nchunk_s * 16;
  p->tail_stride = tail_k == 0 ? 0 : (tail_k % 32 == 16 ? tail_k : tail_k + 16);
  const size_t smem = (size_t)(tPipes * 2 * tMB * p->KP + 32 * p->NJ * p->tail_stride) * sizeof(float) + 64;
  return smem <= 227 * 1024;
}

// Cluster size: the smallest of {1, 2, 4, 6, 8} whose shard fits (tensor memory + the shared-memory K tail) -- 2 for
// S = 321 / 361, 6 for S = 722 (measured: 44.2 % of peak with 6-CTA clusters and 7 x 4 tiles vs 39.5 % with 8-CTA clusters
// and 7 x 3 tiles).  VIT_TMEM_C overrides it for experiments (any 1..8; the kernel is generic in C).
static bool make_tmem_plan(int S, TmemPlan* p, int min_pad = 0) {
  if (const char* e = getenv("VIT_TMEM_C")) {
    const int C = atoi(e);
    if (C >= 1 && C <= 8 && try_tmem_plan(S, C, p, min_pad)) return true;
  }
  const int candidates[] = {1, 2, 4, 6, 8};
  for (int C : candidates)
    if (try_tmem_plan(S, C, p, min_pad)) return true;
  return false;
}

static size_t tmem_smem_bytes(const TmemPlan& p) {
  return (size_t)(tPipes * 2 * tMB * p.KP + 32 * p.NJ * p.tail_stride) * sizeof(float) + 64;
}
static size_t tmem_packed_floats(const TmemPlan& p) { return (size_t)p.C * 128 * tTmemCols; }
static size_t tmem_tail_floats(const TmemPlan& p) { return (size_t)p.C * 32 * p.NJ * p.tail_stride; }

// value of the re-laid-out logA^T for shard r, target slot `slot`, K position kp (0 where padded: the matching delta
// pads are -inf, so padded cells never win the max)
// transposed: the source matrix is stored [i][j] (source-major) instead of [j][i]
__device__ __forceinline__ float plan_a_value(const float* __restrict__ logA_T, int S, const TmemPlan& p, int r, int slot,
                                              int kp, bool transposed = false) {
  const int ncj = p.base + (r < p.rem ? 1 : 0);
  const int ci = kp / p.NCP, l = kp - ci * p.NCP;
  const int nci = p.base + (ci < p.rem ? 1 : 0);
  if (slot >= ncj || l >= nci) return 0.f;
  const int j = r * p.base + min(r, p.rem) + slot;
  const int i = ci * p.base + min(ci, p.rem) + l;
  return transposed ? logA_T[(size_t)i * S + j] : logA_T[(size_t)j * S + i];
}

// packedT  [C][128 TMEM lanes][512 columns]: lane (Q*32 + l) column (c*NJ*4 + n*4 + kk) =
//          logA^T[target slot (Q*8 + l/4) + 32 n][K position 16 c + 4 (l%4) + kk]      for c < nchunk_t
// packedS  [C][32*NJ target slots][tail_stride]: K positions 16*nchunk_t + kk
static __global__ void tmem_pack_kernel(const float* __restrict__ logA_T, int S, TmemPlan p, float* __restrict__ packedT,
                                        float* __restrict__ packedS, bool transposed) {
  const size_t nT = (size_t)p.C * 128 * tTmemCols;
  const size_t nS = (size_t)p.C * 32 * p.NJ * p.tail_stride;
  const int cols = p.NJ * 4;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < nT + nS; x += (size_t)gridDim.x * blockDim.x) {
    if (x < nT) {
      const int col = (int)(x % tTmemCols);
      const int tl = (int)((x / tTmemCols) % 128);
      const int r = (int)(x / ((size_t)tTmemCols * 128));
      const int c = col / cols, w = col - c * cols, n = w >> 2, kk = w & 3;
      float v = 0.f;
      if (c < p.nchunk_t) v = plan_a_value(logA_T, S, p, r, (tl >> 2) + 32 * n, 16 * c + 4 * (tl & 3) + kk, transposed);
      packedT[x] = v;
    } else {
      const size_t y = x - nT;
      const int kk = (int)(y % p.tail_stride);
      const int slot = (int)((y / p.tail_stride) % (32 * p.NJ));
      const int r = (int)(y / ((size_t)p.tail_stride * 32 * p.NJ));
      float v = 0.f;
      if (kk < p.nchunk_s * 16) v = plan_a_value(logA_T, S, p, r, slot, 16 * p.nchunk_t + kk, transposed);
      packedS[y] = v;
    }
  }
}

// ---- tensor memory helpers -----------------------------------------------------------------------------------------
// 32x32b shape: thread l of warp w touches TMEM lane 32*(w&3) + l, N consecutive 32-bit columns.
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float* r);
template <>
__device__ __forceinline__ void tmem_ld<4>(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
}
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                 "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr));
}
// NC = NJ*4 columns of one K chunk -> registers (asynchronous: valid only after tmem_wait_ld on the same registers)
template <int NC>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, float* r) {
  if constexpr (NC >= 16) {
    tmem_ld<16>(taddr, r);
    if constexpr (NC - 16 >= 8) tmem_ld<8>(taddr + 16, r + 16);
    else if constexpr (NC - 16 >= 4) tmem_ld<4>(taddr + 16, r + 16);
  } else if constexpr (NC >= 8) {
    tmem_ld<8>(taddr, r);
    if constexpr (NC - 8 >= 4) tmem_ld<4>(taddr + 8, r + 8);
  } else {
    tmem_ld<4>(taddr, r);
  }
}
template <int N>
__device__ __forceinline__ void reg_fence(float* r) {   // compiler-level: later uses of r[0..N) stay below this point
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  if constexpr (N == 4) asm volatile("" : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]));
  if constexpr (N == 8)
    asm volatile("" : "+r"(u[0]), "+r"(u[1]), "+r"(u[2]), "+r"(u[3]), "+r"(u[4]), "+r"(u[5]), "+r"(u[6]), "+r"(u[7]));
}
// wait for this thread's outstanding tcgen05.ld; the destination registers are threaded through the asm statements so
// that nothing that reads them can be scheduled above the wait
template <int NC>
__device__ __forceinline__ void tmem_wait_ld(float* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i + 8 <= NC; i += 8) reg_fence<8>(r + i);
  if constexpr (NC % 8 == 4) reg_fence<4>(r + NC - 4);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float4 v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
               :: "r"(taddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                  "r"(__float_as_uint(v.w)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tpipe_bar_sync(int pipe) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + pipe), "n"(tPipeThreads) : "memory");
}

}  // namespace vit
