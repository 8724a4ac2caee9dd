// Shared device helpers (sm_100a PTX wrappers) and host-side launch bookkeeping for the Viterbi kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/vit_b200.h"

namespace vit {

// ---- host-side bookkeeping (vit_api.cu owns the storage) -------------------------------------------------------
void note_launch(int n = 1);          // bumps the process-wide kernel launch counter
int cuda_fail(cudaError_t e);         // records the message, returns VIT_ERR_CUDA
cudaStream_t backtrace_stream_override();   // vit_decode_opts.backtrace_stream of the call in flight on this thread, or 0
#define VIT_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t vit_e_ = (expr);                             \
    if (vit_e_ != cudaSuccess) return ::vit::cuda_fail(vit_e_); \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- (value, index) max with NumPy's first-maximum-wins rule ---------------------------------------------------
// imm/tf_viterbi.py:99 `np.argmax(Bt, axis=1)`: among equal maxima the LOWEST index wins.  `==` (not max.f32
// ordering) is used for the tie test so that +0/-0 compare equal exactly as NumPy does.
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, off);
    int oi = __shfl_xor_sync(0xffffffffu, i, off);
    argmax_combine(v, i, ov, oi);
  }
}

// The same reduction in two instructions (sm_100a): the warp maximum with redux.sync.max.f32, then the lowest index
// among the lanes that hold it with redux.sync.min.u32.  `==` again decides who holds the maximum (+0 / -0 equal).
__device__ __forceinline__ void warp_argmax_redux(float& v, int& i) {
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  const unsigned cand = (v == m) ? (unsigned)i : 0xffffffffu;
  unsigned r;
  asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(cand));
  v = m;
  i = (int)r;
}

// ---- packed fp32 math (sm_100a): FADD2 and FMNMX3 --------------------------------------------------------------
// add.rn.f32x2 rounds each lane exactly like add.rn.f32 (two independent IEEE binary32 adds), so it is bit-exact
// against the reference's np.add; it halves the issue slots and, with operand reuse, the register-file reads.
__device__ __forceinline__ void fadd2(float& rx, float& ry, float ax, float ay, float bx, float by) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rc, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(rx), "=f"(ry)
      : "f"(ax), "f"(ay), "f"(bx), "f"(by));
}
// 3-input max (one ALU-pipe instruction for two of the recursion's maxes). No NaNs enter the recursion.
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// ---- thread-block cluster / distributed shared memory / mbarrier ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync() { cluster_arrive(); cluster_wait(); }

// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa(uint32_t cta_addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank)); return r;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// CTA-scope acquire (the default semantics): what bulk-async / tcgen05.commit completions need.  The .cluster-scope form
// above makes ptxas append CCTL.IVALL (a whole-L1 invalidate) to every successful wait.
__device__ __forceinline__ bool mbar_try_wait_cta(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cta(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait_cta(bar, parity)) {}
}
// make this thread's generic-proxy shared-memory writes visible to the async proxy (bulk copies)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// bulk async copy local shared memory -> a peer CTA's shared memory, completing `bytes` on the peer's mbarrier.
// dst and mbar are shared::cluster addresses (mapa); bytes % 16 == 0, all addresses 16-byte aligned.
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(mbar_cluster) : "memory");
}

__device__ __forceinline__ float ld_global_nc_f32(const float* p) {
  float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
}
__device__ __forceinline__ void st_global_cs_f32(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace vit
