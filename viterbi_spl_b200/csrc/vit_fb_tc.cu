// Forward-backward on the 5th-generation tensor cores (tcgen05.mma, accumulators and the A operand in tensor memory).
//
// Same recursion and the same deferred-normaliser scheme as vit_fb.cu (oracle/fb_oracle.py states the semantics; there
// is no reference implementation: parity unpinned).  Per step the matrix-vector products of all clips of a cluster are
// ONE small GEMM  D[128 x 64] = M_shard[128 x K] . [V_hi | V_lo][K x 64]  (32 clips, two bf16 terms of V side by side):
//   * M_shard  = this CTA's <= 127 rows of A^T (forward) / A (backward), all K source positions, resident in TENSOR
//                MEMORY for the whole kernel as the A operand of the "TS" form of tcgen05.mma.  fp32 does not fit and
//                TF32's 10-bit mantissa cannot hold 1e-4 on gamma, so every fp32 value x is split into two bf16 terms
//                x = hi + lo (16 mantissa bits); two N = 64 MMAs per K block -- A_hi . [V_hi | V_lo] and A_lo . [V_hi | V_lo],
//                i.e. the four products hi.hi, hi.lo, lo.hi, lo.lo -- accumulate in fp32; the epilogue adds the halves.
//                Row 127 of the forward operand is all ones: D[127][n] = sum_k alpha~[k][n] is the normaliser c_{t-1}
//                for free.
//   * V        = alpha~_{t-1} (forward) / w_{t+1} (backward) of 32 clips, bf16 hi and lo copies in shared memory in the
//                canonical no-swizzle MN-major layout (8 x 16-byte core matrices; validated by tools/microbench_umma.cu;
//                per group of 8 K positions: 4 cores of hi, then 4 cores of lo = one N = 64 operand),
//                double buffered.  Each CTA writes the rows of its own states and pushes them to its peers with bulk
//                async DSMEM copies that complete on the receiver's mbarrier.
//   * D        = fp32 in TMEM; 256 threads (two per row = state, 16 clips each) read their row with tcgen05.ld, scale
//                by 1/c, multiply by the emission likelihoods, store alpha~ / gamma, and write the next V.
// One thread issues the 48 UTCHMMA of a step and commits them to an mbarrier; the step is a dependency chain
// (wait V -> MMA -> epilogue -> exchange), so the tensor pipe is lightly used -- the point is the ~4x shorter chain than
// the FFMA kernel's, not tensor throughput.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>

#include "vit_common.cuh"

namespace vit {

constexpr int cN = 32;            // clips per cluster
constexpr int cNN = 2 * cN;       // N of the MMA: the bf16 hi copy of V in columns [0, 32), the lo copy in [32, 64)
constexpr int cM = 128;           // rows of the MMA = TMEM lanes = threads
constexpr int cThreads = 256;        // two threads per row: thread (m, half) handles clips [16 half, 16 half + 16)
constexpr int cH = cN / 2;           // clips per thread
constexpr int cTmemCols = 512;

struct TcPlan {
  int C;        // CTAs per cluster
  int NCP;      // K positions per shard (multiple of 16); shard r owns K positions [r*NCP, r*NCP + nc(r))
  int KP;       // C * NCP
  int base, rem;
};

static bool make_tc_plan(int S, TcPlan* p) {
  if (S < 1) return false;
  const int C = (S + 126) / 127;                          // <= 127 states per CTA: row 127 is the ones row
  if (C > 8) return false;
  p->C = C;
  p->base = S / C;
  p->rem = S % C;
  const int ncmax = p->base + (p->rem ? 1 : 0);
  p->NCP = (ncmax + 15) / 16 * 16;
  p->KP = C * p->NCP;
  if (p->KP + cNN > cTmemCols) return false;              // A hi + lo = KP columns, D = 64 columns
  const size_t smem = (size_t)2 * 2 * p->KP * cN * 2 + 256;
  return smem <= 200 * 1024;
}

static size_t tc_smem_bytes(const TcPlan& p) { return (size_t)2 * 2 * p.KP * cN * 2 + 256; }
static size_t tc_packed_words(const TcPlan& p) { return (size_t)p.C * 2 * cM * (p.KP / 2); }

// packed [C][term: hi, lo][128 rows][KP/2 columns] uint32: column c = bf16(k = 2c) | bf16(k = 2c + 1) << 16.
// transposed = false: row m of shard r is M[j][.] = A stored [j][i] ... the operand is "rows = outputs, K = inputs":
//   forward : out j, in i, value A[i][j]  -> transposed read of the source-major A
//   backward: out i, in j, value A[i][j]  -> direct read
__global__ void tc_pack_kernel(const float* __restrict__ A, int S, TcPlan p, bool transposed, bool ones_row,
                               uint32_t* __restrict__ packed) {
  const int cols = p.KP / 2;
  const size_t total = (size_t)p.C * cM * cols;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(x % cols);
    const int m = (int)((x / cols) % cM);
    const int r = (int)(x / ((size_t)cols * cM));
    const int nc = p.base + (r < p.rem ? 1 : 0);
    uint32_t hi = 0, lo = 0;
    for (int h = 0; h < 2; ++h) {
      const int kp = 2 * c + h;
      const int ci = kp / p.NCP, l = kp - ci * p.NCP;
      const int nci = p.base + (ci < p.rem ? 1 : 0);
      float v = 0.f;
      if (l < nci) {
        const int in = ci * p.base + min(ci, p.rem) + l;
        if (m < nc) {
          const int out = r * p.base + min(r, p.rem) + m;
          v = transposed ? A[(size_t)in * S + out] : A[(size_t)out * S + in];
        } else if (m == cM - 1 && ones_row) {
          v = 1.f;
        }
      }
      const __nv_bfloat16 bh = __float2bfloat16(v);
      const __nv_bfloat16 bl = __float2bfloat16(v - __bfloat162float(bh));
      hi |= (uint32_t)__bfloat16_as_ushort(bh) << (16 * h);
      lo |= (uint32_t)__bfloat16_as_ushort(bl) << (16 * h);
    }
    packed[((size_t)(r * 2 + 0) * cM + m) * cols + c] = hi;
    packed[((size_t)(r * 2 + 1) * cM + m) * cols + c] = lo;
  }
}

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* d) {
  uint32_t* u = reinterpret_cast<uint32_t*>(d);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                 "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// One leader lane of a converged warp (the same lane every time for the full mask).  Issuing tcgen05.mma under
// `if (tid == 0)` makes ptxas wrap EVERY UTCHMMA in an ELECT / BRA.U.ANY loop over the "active threads" (~28 clocks per
// MMA on top of the tensor pipe's floor); under warp-uniform control flow + elect.sync it is a straight instruction stream.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
// 8 fp32 -> 8 bf16 hi (16 bytes) and 8 bf16 lo
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16(v[2 * i]), h1 = __float2bfloat16(v[2 * i + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16(v[2 * i] - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16(v[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <bool BWD>
__global__ void __launch_bounds__(cThreads, 1)
fb_tc_pass_kernel(const uint32_t* __restrict__ packed, const float* __restrict__ pi, const float* __restrict__ lik,
                  const int32_t* __restrict__ lengths, int B, int T_max, int S, TcPlan p, float* __restrict__ gamma,
                  float* __restrict__ cnorm, int dev) {
  // dev: timing experiments only: 1 = no HBM traffic, 2 = no exchange, 4 = no MMAs (results invalid); 8 = print the
  // clock stamps of one step (results valid; needs -DVIT_FB_STAMPS)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int KP = p.KP, NCP = p.NCP;
  const uint32_t LBO = (cNN / 8) * 128;                 // bytes between groups of 8 K positions
  const uint32_t lo_off = (cN / 8) * 128;               // the lo copy's cores follow the hi copy's within a K group
  const uint32_t buf_bytes = (uint32_t)KP * cNN * 2;    // hi + lo
  uint8_t* sV = smem_raw;                               // [2 buffers][KP/8][8 cores: 4 hi, 4 lo][8 k][8 n] bf16
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + 2 * buf_bytes);     // [0..1] V ready, [2] MMA done
  __shared__ __align__(16) float s_c[2][cN];            // 1 / normalisers of the step (forward: from D row 127)
  __shared__ float s_craw[cN];
  __shared__ int s_len[cN];
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = tid & (cM - 1), half = tid >> 7, n0 = half * cH;     // my row (= TMEM lane) and my 16 clips
  const uint32_t C = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int nc_mine = p.base + ((int)rank < p.rem ? 1 : 0);
  const int j = (int)rank * p.base + min((int)rank, p.rem) + row;      // my state
  const bool row_ok = row < nc_mine;
  const float pi_j = (!BWD && row_ok) ? pi[j] : 0.f;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&s_tmem_base)), "n"(cTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int x = tid; x < (int)(2 * buf_bytes / 16); x += cThreads) reinterpret_cast<uint4*>(sV)[x] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(smem_u32(&s_bar[0]), 1);
    mbar_init(smem_u32(&s_bar[1]), 1);
    mbar_init(smem_u32(&s_bar[2]), 1);
    mbar_fence_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = s_tmem_base;
  const uint32_t tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  {
    // A operand -> TMEM: hi copy in columns [0, KP/2), lo copy in [KP/2, KP); row = lane
    const int cols = KP / 2;
    {
      const int term = half;                               // the two thread halves fill one bf16 copy each
      const uint4* src = reinterpret_cast<const uint4*>(packed + ((size_t)(rank * 2 + term) * cM + row) * cols);
      for (int x = 0; x < cols / 4; ++x) {
        const uint4 v = src[x];
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                     ::"r"(tlane + term * cols + 4 * x), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the zero-filled V buffers, for the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (C > 1) cluster_sync();

  const uint32_t d_tmem = tbase + KP;                    // accumulator columns [KP, KP + 64): M.V_hi | M.V_lo
  // instruction descriptor: D = F32, A = B = BF16, A K-major (TMEM), B MN-major, N >> 3, M >> 4
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(cNN >> 3) << 17) | ((uint32_t)(cM >> 4) << 24);
  const uint32_t slice_bytes = (uint32_t)NCP * cNN * 2;  // my rows of V (hi and lo cores): contiguous in the canonical layout
  const uint32_t tx_bytes = (C - 1) * slice_bytes;
  const long long gamma_delta = reinterpret_cast<const char*>(gamma) - reinterpret_cast<const char*>(lik);

  // tcgen05.mma over K blocks [kb0, kb1) of V buffer `vbuf`.  All four bf16 products are needed (hi.hi, hi.lo, lo.hi,
  // lo.lo; without lo.lo the gamma error reaches 1.03e-4, measured), but the hi and lo copies of V sit side by side as
  // ONE N = 64 operand, so a K block is two MMAs -- A_hi . [V_hi | V_lo] and A_lo . [V_hi | V_lo] -- instead of four
  // N = 32 ones (the MMAs of a step are issue-latency bound: ~45 clocks each whatever N), and the epilogue adds the two
  // halves of D.  Descriptor = start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46; a K block of 16 advances
  // the start field by 2 LBO >> 4 (the smem window is < 256 KB: no carry).
  auto issue_mma = [&](uint32_t vbuf, int kb0, int kb1, uint32_t acc) {
    const uint32_t vb = smem_u32(sV) + vbuf * buf_bytes;
    const uint32_t desc_hi32 = (128u >> 4) | (1u << 14);
    uint32_t lo = (((vb >> 4) & 0x3FFF) | (((LBO >> 4) & 0x3FFF) << 16)) + kb0 * ((2 * LBO) >> 4);
    uint32_t a_hi = tbase + kb0 * 8, a_lo = tbase + KP / 2 + kb0 * 8;
#pragma unroll 4
    for (int kb = kb0; kb < kb1; ++kb) {
      const uint64_t desc = ((uint64_t)desc_hi32 << 32) | lo;
      tc_mma(d_tmem, a_hi, desc, idesc, acc);
      acc = 1;
      tc_mma(d_tmem, a_lo, desc, idesc, 1);
      lo += (2 * LBO) >> 4;
      a_hi += 8;
      a_lo += 8;
    }
  };

  uint32_t g = 0;           // the MMA of iteration g reads V buffer g & 1; the epilogue writes buffer (g + 1) & 1
  uint32_t ph[2] = {0, 0};  // completed phases of the two "V rows have landed" barriers
  uint32_t n_mma = 0;       // MMA batches committed so far (parity of s_bar[2])
  for (int seq0 = (int)cluster_id_x() * cN; seq0 < B; seq0 += (int)num_clusters_x() * cN) {
    __syncthreads();
    // no CTA may start pushing rows of the next sub-batch while a peer's last MMA still reads that buffer
    if (C > 1) cluster_sync();
    if (tid < cN) {
      const int b = seq0 + tid;
      s_len[tid] = b < B ? (lengths ? lengths[b] : T_max) : 0;
    }
    __syncthreads();
    int maxlen = 0;
    for (int n = 0; n < cN; ++n) maxlen = max(maxlen, s_len[n]);
    const int n_iter = BWD ? maxlen : maxlen + 1;         // forward: one extra MMA-only step yields the last normaliser
    // lengths of my 16 clips in registers (0 for every clip of a row that holds no state): the step's loads and stores
    // are then plain predicated instructions -- with the shared-memory reads inside, ptxas built a branch per clip and
    // the epilogue was a serial chain of 16 x (LDS, branch, LDS, branch): 2800 of a step's 8100 clocks
    int len_r[cH];
#pragma unroll
    for (int n = 0; n < cH; ++n) len_r[n] = row_ok ? s_len[n0 + n] : 0;

    bool first = true;
    for (int it = 0; it < n_iter; ++it, ++g) {
      const int t = BWD ? maxlen - 1 - it : it;
      const uint32_t cur = g & 1u, nxt = cur ^ 1u;
#ifdef VIT_FB_STAMPS
      // -DVIT_FB_STAMPS + VIT_DEV_FLAGS=8: clock stamps of one step of one CTA (thread 0), printed -- where the step's
      // time goes (costs 36 registers, hence compile-time)
      const bool stamp = (dev & 8) && it == 100 && blockIdx.x == 0 && tid == 0;
      long long ck[10];
      if (stamp) ck[0] = clock64();
#define VIT_STAMP(i) do { if (stamp) ck[i] = clock64(); } while (0)
#else
#define VIT_STAMP(i) do { } while (0)
#endif
      const bool tail = !BWD && it == maxlen;             // forward's extra step
      // this step's likelihoods (backwards also the stored alpha~): issued first, used after the MMAs
      float e[cH], al[cH];
      // lik[clip n][t][j] = p_t + n * (T_max * S): one running 64-bit pointer instead of 32 address computations
      const size_t clip_stride = (size_t)T_max * S;
      const float* p_t = lik + ((size_t)seq0 * T_max + (tail ? 0 : t)) * S + j;
      {
        const float* pn = p_t + (size_t)n0 * clip_stride;
#pragma unroll
        for (int n = 0; n < cH; ++n, pn += clip_stride) {
          const bool lv = !tail && t < len_r[n] && !(dev & 1);
          e[n] = lv ? ld_global_nc_f32(pn) : 0.f;
          if (BWD) al[n] = lv ? ld_global_nc_f32(reinterpret_cast<const float*>(reinterpret_cast<const char*>(pn) + gamma_delta)) : 0.f;
        }
      }
      float d[cH];
#pragma unroll
      for (int n = 0; n < cH; ++n) d[n] = 0.f;
      if (!first) {
        if (C > 1 && !(dev & 2)) { mbar_wait_cta(smem_u32(&s_bar[cur]), ph[cur] & 1u); ++ph[cur]; }   // peers' rows of V have landed
        VIT_STAMP(1);
        if (warp == 0 && !(dev & 4) && elect_one_sync()) {
          // the K blocks of my own rows were issued at the end of the previous iteration (see below); now that the
          // peers' rows have landed, the other shards' K blocks follow and the batch is committed
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int own0 = (int)rank * NCP / 16, own1 = own0 + NCP / 16;
          issue_mma(cur, 0, own0, 1);
          issue_mma(cur, own1, KP / 16, 1);
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar[2])) : "memory");
        }
        if (BWD && tid < 2 * cN) {
          // 1 / c_t (threads 0-31) and 1 / c_{t+1} (32-63) of the 32 clips: fetched while the MMAs run
          const int n = tid & (cN - 1), tt = t + (tid >> 5);
          // (guarded like the forward pass: a normaliser of 0 -- an impossible observation sequence -- gives gamma = 0,
          // not inf * 0 = NaN; log L is then -inf)
          const float cv = (tt < s_len[n]) ? cnorm[(size_t)(seq0 + n) * T_max + tt] : 1.f;
          s_c[tid >> 5][n] = cv > 0.f ? 1.f / cv : 0.f;
        }
        VIT_STAMP(2);
        if (!(dev & 4)) { mbar_wait_cta(smem_u32(&s_bar[2]), n_mma & 1u); ++n_mma; }
        VIT_STAMP(3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
          // D[:, 0:32] = M . V_hi, D[:, 32:64] = M . V_lo: my 16 clips of both halves
          float dl[cH];
          tc_ld16(tlane + KP + n0, d);
          tc_ld16(tlane + KP + cN + n0, dl);
#pragma unroll
          for (int n = 0; n < cH; ++n) d[n] += dl[n];
        }
        if (!BWD && (warp & 3) == 3) {
          // row 127 = the ones row: c_{t-1}[n].  Lane 31 of warps 3 and 7 holds 16 sums each; the warp turns them into
          // 1 / c and one CTA per cluster records them
          if (row == cM - 1) {
#pragma unroll
            for (int n = 0; n < cH; ++n) s_craw[n0 + n] = d[n];
          }
          __syncwarp();
          const int n = n0 + (tid & 15);
          if ((tid & 31) < cH) {
            const float c = s_craw[n];
            s_c[0][n] = c > 0.f ? 1.f / c : 0.f;                      // every thread multiplies by 1 / c
            if (rank == 0 && t - 1 < s_len[n] && !(dev & 1)) cnorm[(size_t)(seq0 + n) * T_max + (t - 1)] = c;
          }
        }
      }
      if (BWD && first && tid < cN) {
        const float cv = (t < s_len[tid]) ? cnorm[(size_t)(seq0 + tid) * T_max + t] : 1.f;
        s_c[0][tid] = cv > 0.f ? 1.f / cv : 0.f;
      }
      VIT_STAMP(4);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();                                    // s_c visible; D fully read before the next MMA batch
      VIT_STAMP(5);
      if (tail) { first = false; continue; }

      float v[cH];
      {
        // 1 / c of my 16 clips: 4 (forward) or 8 (backward) independent LDS.128
        float sc0[cH], sc1[cH];
#pragma unroll
        for (int q4 = 0; q4 < cH / 4; ++q4) {
          const float4 a4 = reinterpret_cast<const float4*>(&s_c[0][n0])[q4];
          sc0[4 * q4] = a4.x; sc0[4 * q4 + 1] = a4.y; sc0[4 * q4 + 2] = a4.z; sc0[4 * q4 + 3] = a4.w;
          if (BWD) {
            const float4 b4 = reinterpret_cast<const float4*>(&s_c[1][n0])[q4];
            sc1[4 * q4] = b4.x; sc1[4 * q4 + 1] = b4.y; sc1[4 * q4 + 2] = b4.z; sc1[4 * q4 + 3] = b4.w;
          }
        }
        float* gn = reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(p_t)) + gamma_delta) + (size_t)n0 * clip_stride;
#pragma unroll
        for (int n = 0; n < cH; ++n, gn += clip_stride) {
          const bool lv = t < len_r[n];
          if (!BWD) {
            // alpha~_t = (M alpha~_{t-1} / c_{t-1}) * b_t ;  alpha~_0 = pi * b_0
            const float u = first ? pi_j : d[n] * sc0[n];
            v[n] = u * e[n];
            if (lv && !(dev & 1)) st_global_cs_f32(gn, v[n]);
          } else {
            // beta_t = M w_{t+1} / c_{t+1} (1 at the clip's last frame); gamma_t = alpha~_t / c_t * beta_t; w_t = b_t beta_t
            const float be = (t == len_r[n] - 1) ? 1.f : (first ? 0.f : d[n] * sc1[n]);
            v[n] = lv ? e[n] * be : 0.f;
            if (lv && !(dev & 1)) st_global_cs_f32(gn, al[n] * sc0[n] * be);
          }
        }
      }
      // my half of my row of the next V (K position rank*NCP + row), bf16 hi and lo: 2 core-matrix rows of 16 bytes
      if (row_ok) {
        const uint32_t k = rank * NCP + row;
        uint8_t* vrow = sV + nxt * buf_bytes + (k >> 3) * LBO + (k & 7) * 16 + half * (cH / 8) * 128;
#pragma unroll
        for (int c8 = 0; c8 < cH / 8; ++c8) {
          uint4 hi, lo;
          split8(v + 8 * c8, hi, lo);
          *reinterpret_cast<uint4*>(vrow + c8 * 128) = hi;
          *reinterpret_cast<uint4*>(vrow + lo_off + c8 * 128) = lo;
        }
      }
      first = false;
      VIT_STAMP(6);
      fence_proxy_async_smem();
      __syncthreads();
      VIT_STAMP(7);
      // my own rows of the next V are complete.  Warp 1 pushes them to the peers at once; warp 0 starts the next step's
      // GEMM on their K blocks (accumulator reset) under the exchange.  (The push used to sit behind the 16 MMA issues
      // in thread 0's program order, ~80 clocks each: the peers got their rows 1300 clocks late.)  Every thread has
      // read D (tcgen05.ld completed before the barrier above).
      if (C > 1 && !(dev & 2) && warp == 1) {
        if (lane == 0) mbar_arrive_expect_tx(smem_u32(&s_bar[nxt]), tx_bytes);
        if (lane < (int)(C - 1)) {
          const uint32_t peer = (rank + 1 + lane) % C;
          const uint32_t src = smem_u32(sV + nxt * buf_bytes + (size_t)rank * slice_bytes);
          dsmem_bulk_copy(mapa(src, peer), src, slice_bytes, mapa(smem_u32(&s_bar[nxt]), peer));
        }
      }
      if (warp == 0 && !(dev & 4) && it + 1 < n_iter && elect_one_sync()) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_mma(nxt, (int)rank * NCP / 16, (int)rank * NCP / 16 + NCP / 16, 0);
      }
#ifdef VIT_FB_STAMPS
      if (stamp) {
        ck[8] = clock64();
        printf("fb_tc step (%s): wait V %lld | issue MMA %lld | wait MMA %lld | tmem ld + norm %lld | sync1 %lld | epilogue %lld | "
               "fence + sync2 %lld | own MMA + push %lld | total %lld clk\n", BWD ? "bwd" : "fwd", ck[1] - ck[0], ck[2] - ck[1],
               ck[3] - ck[2], ck[4] - ck[3], ck[5] - ck[4], ck[6] - ck[5], ck[7] - ck[6], ck[8] - ck[7], ck[8] - ck[0]);
      }
#endif
#undef VIT_STAMP
    }
    // backward: the rows pushed in the last step are never consumed -- wait for them so that every armed phase is
    // matched by exactly one wait (forward: the extra step consumed the last push)
    if (BWD && C > 1 && !first && !(dev & 2)) { mbar_wait_cta(smem_u32(&s_bar[g & 1u]), ph[g & 1u] & 1u); ++ph[g & 1u]; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "n"(cTmemCols) : "memory");
  if (C > 1) cluster_sync();
}

// The 16-bit-mantissa products leave a relative error of ~1e-5 per step in the SCALE of beta (forward and backward
// normalisers no longer cancel exactly), which random-walks to ~5e-4 over 3000 frames -- but it is common to all states
// of a frame, so renormalising every gamma_t to sum 1 removes it.  One warp per frame, HBM-bound (read + write gamma).
__global__ void __launch_bounds__(256) fb_normalize_gamma_kernel(float* __restrict__ gamma, const int32_t* __restrict__ lengths,
                                                                 int B, int T_max, int S) {
  const int lane = threadIdx.x & 31;
  const long long n_frames = (long long)B * T_max;
  for (long long f = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < n_frames;
       f += ((long long)gridDim.x * blockDim.x) >> 5) {
    if (lengths && (int)(f % T_max) >= lengths[f / T_max]) continue;
    float* g = gamma + f * S;
    float sum = 0.f;
    for (int k = lane; k < S; k += 32) sum += g[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    for (int k = lane; k < S; k += 32) g[k] *= inv;
  }
}

bool fb_tc_supported(int S) {
  TcPlan p;
  return make_tc_plan(S, &p);
}

size_t fb_tc_workspace_bytes(int B, int T_max, int S) {
  TcPlan p;
  if (!make_tc_plan(S, &p)) return 0;
  size_t bytes = 2 * align_up(tc_packed_words(p) * sizeof(uint32_t), 256);
  bytes += align_up((size_t)(B > 0 ? B : 1) * T_max * sizeof(float), 256);
  return bytes;
}

// vit_fb.cu
__global__ void fb_loglik_kernel(const float* __restrict__ cnorm, const int32_t* __restrict__ lengths, int B, int T_max,
                                 float* __restrict__ loglik);

int fb_tc_run(const float* A, const float* pi, const float* lik, const int32_t* lengths, int B, int T_max, int S,
              void* workspace, size_t workspace_bytes, float* gamma, float* loglik, cudaStream_t stream) {
  TcPlan p;
  if (!make_tc_plan(S, &p)) return VIT_ERR_UNSUPPORTED_ALGO;
  if (workspace_bytes < fb_tc_workspace_bytes(B, T_max, S)) return VIT_ERR_WORKSPACE_TOO_SMALL;
  if (B == 0) return VIT_OK;
  char* ws = (char*)workspace;
  uint32_t* packed[2];
  for (int k = 0; k < 2; ++k) {
    packed[k] = (uint32_t*)ws;
    ws += align_up(tc_packed_words(p) * sizeof(uint32_t), 256);
  }
  float* cnorm = (float*)ws;
  if (lengths) VIT_CUDA_TRY(cudaMemsetAsync(gamma, 0, (size_t)B * T_max * S * sizeof(float), stream));
  {
    const size_t total = (size_t)p.C * cM * (p.KP / 2);
    const int grid = (int)((total + 255) / 256);
    tc_pack_kernel<<<grid, 256, 0, stream>>>(A, S, p, /*transposed=*/true, /*ones_row=*/true, packed[0]);
    tc_pack_kernel<<<grid, 256, 0, stream>>>(A, S, p, /*transposed=*/false, /*ones_row=*/false, packed[1]);
    note_launch(2);
    VIT_CUDA_TRY(cudaGetLastError());
  }
  const size_t smem = tc_smem_bytes(p) < 120 * 1024 ? 120 * 1024 : tc_smem_bytes(p);   // one CTA per SM (whole TMEM)
  auto kf = fb_tc_pass_kernel<false>;
  auto kb = fb_tc_pass_kernel<true>;
  VIT_CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VIT_CUDA_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(cThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  cfg.gridDim = dim3(p.C);
  VIT_CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, kf, &cfg));
  if (max_clusters < 1) return VIT_ERR_UNSUPPORTED_ALGO;
  int num_sms = 148, devid = 0;
  VIT_CUDA_TRY(cudaGetDevice(&devid));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, devid));
  if (max_clusters * p.C > num_sms) max_clusters = num_sms / p.C;
  const int want = (B + cN - 1) / cN;
  const int n_clusters = want < max_clusters ? want : max_clusters;
  cfg.gridDim = dim3(n_clusters * p.C);
  const char* dev_s = getenv("VIT_DEV_FLAGS");
  const int dev = dev_s ? atoi(dev_s) : 0;
  VIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kf, (const uint32_t*)packed[0], pi, lik, lengths, B, T_max, S, p, gamma, cnorm, dev));
  note_launch();
  VIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kb, (const uint32_t*)packed[1], pi, lik, lengths, B, T_max, S, p, gamma, cnorm, dev));
  note_launch();
  {
    const long long warps = (long long)B * T_max;
    long long blocks = (warps + 7) / 8;
    if (blocks > (long long)num_sms * 16) blocks = (long long)num_sms * 16;
    fb_normalize_gamma_kernel<<<(unsigned)blocks, 256, 0, stream>>>(gamma, lengths, B, T_max, S);
    note_launch();
    VIT_CUDA_TRY(cudaGetLastError());
  }
  if (loglik) {
    fb_loglik_kernel<<<(B + 3) / 4, 128, 0, stream>>>(cnorm, lengths, B, T_max, loglik);
    note_launch();
    VIT_CUDA_TRY(cudaGetLastError());
  }
  return VIT_OK;
}

}  // namespace vit
