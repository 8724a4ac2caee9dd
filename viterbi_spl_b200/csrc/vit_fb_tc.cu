// Forward-backward on the 5th-generation tensor cores (tcgen05.mma, accumulators and the A operand in tensor memory).
//
// Same recursion as vit_fb.cu (oracle/fb_oracle.py states the semantics; there is no reference implementation: parity
// unpinned).  Per step the matrix-vector products of the CN clips of a SUB-BATCH are one small GEMM
//     D[128 x 2 CN] = M_shard[128 x K] . [V_hi | V_lo][K x 2 CN]
//   * M_shard  = this CTA's <= 124 rows of A^T (forward) / A (backward), all K source positions, resident in TENSOR
//                MEMORY for the whole kernel as the A operand of the "TS" form of tcgen05.mma.  fp32 does not fit and
//                TF32's 10-bit mantissa cannot hold 1e-4 on gamma, so every fp32 value x is split into two bf16 terms
//                x = hi + lo (16 mantissa bits); two MMAs per K block -- A_hi . [V_hi | V_lo] and A_lo . [V_hi | V_lo],
//                i.e. the four products hi.hi, hi.lo, lo.hi, lo.lo -- accumulate in fp32; the epilogue adds the halves.
//   * V        = alpha~_{t-1} (forward) / w_{t+1} (backward), bf16 hi and lo copies side by side in shared memory in
//                the canonical no-swizzle MN-major layout (8 x 16-byte core matrices; validated by
//                tools/microbench_umma.cu), double buffered.  Each CTA writes the rows of its own states and pushes them
//                to its peers with bulk-async DSMEM copies that complete on the receiver's mbarrier.
//   * D        = fp32 in TMEM; the epilogue threads (two per row = state and sub-batch, 8 clips each) read their row
//                with tcgen05.ld, scale, multiply by the emission likelihoods, store alpha~ / gamma, and write the next V.
//
// What round 2 changed (the round-1 kernel was ONE dependency chain per cluster -- wait V -> 48 MMAs -> epilogue ->
// exchange, 5200 clocks per step, the tensor pipe idle 70 % of the time; profiles/r01k_*): 19.2 -> 13.1 ms at 1024 x 3000 x 361
//   * TWO independent sub-batches X and Y per cluster, each with its own accumulator, V buffers and 8 epilogue warps
//     (warps 0-3, 8-11: X; 4-7, 12-15: Y).  A dedicated MMA warp (warp 16) issues X's 48 MMAs, then Y's, then X's of the
//     next step ...; X's epilogue runs while Y's MMAs issue and vice versa; an exchange warp (warp 17) pushes a CTA's
//     rows to its peers the moment its epilogue warps have written them.  MMA issue is ~27 clocks per N = 32 MMA
//     (11 + N / 2: tools/microbench_umma_rate.cu, profiles/r02_microbench_umma_rate.jsonl), 2 x 1280 of the ~4600 clocks
//     of a step pair; the rest is hand-off latency between the roles (DESIGN.md section 3.8 lists what was tried).
//   * no block-wide barrier in the step: lane 31 of every TMEM quadrant holds no state; its operand row is all ones
//     (forward), so D[that row] = sum_k alpha~[k] = the normaliser c_{t-1} arrives in every warp's own lane 31 and is
//     broadcast with shuffles (round 1: one ones-row, a shared-memory broadcast and two __syncthreads per step).
//   * the gamma renormalisation pass is gone (round 1: a separate HBM-bound kernel, read + write of gamma, 2 ms of 19).
//     The bf16 products leave ~1e-5 of relative error per step in the SCALE of beta, common to all states of a frame,
//     so gamma_t must be renormalised to sum 1.  Backward: every warp reduces gamma~_t of its 31 states x 8 clips
//     with shuffles and parks the partial sums as "the V row of lane 31" (a spare K position); they travel to the peers
//     with the rows anyway; the backward operand's lane-31 rows are indicators of those K positions, so the NEXT step's
//     GEMM delivers s_t = sum_j gamma~_t[j] in every warp's lane 31 for free.  gamma~_t is held in registers for one
//     step and stored as gamma~_t / s_t.
//   * the step's inputs arrive by cp.async one step ahead; 1 / c is one MUFU.RCP (the IEEE division's range-check branch
//     serialised ~110 clocks per quotient); the epilogue hands its V rows over BEFORE it stores to HBM.
//   * -DVIT_FB_STAMPS prints clock stamps of one step of the MMA warp and of one epilogue warp (how all of the above was
//     found).
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>

#include "vit_common.cuh"

namespace vit {

constexpr int cM = 128;            // rows of the MMA = TMEM lanes
constexpr int cRows = 124;         // states per CTA: 31 per TMEM quadrant, lane 31 of every quadrant is the special row
constexpr int cEW = 16;            // epilogue warps: warp w -> TMEM quadrant w & 3 (its 32 lanes), sub-batch (w >> 2) & 1, clip part w >> 3
constexpr int cParts = cEW / 8;    // threads per row (= state) and sub-batch: each handles CN / cParts clips
constexpr int cEpi = 32 * cEW;
constexpr int cThreads = cEpi + 64;   // + warp cEW: MMA issue, warp cEW + 1: exchange (pushes my rows of V to the peers)
constexpr int cTmemCols = 512;

struct TcPlan {
  int C;        // CTAs per cluster
  int NCP;      // K positions per shard (multiple of 32): shard r owns K positions [r*NCP, (r+1)*NCP); position l of a shard
                // is TMEM lane l of the CTA that produces it: l % 32 == 31 -> special, else state l - l / 32 of the shard
  int KP;       // C * NCP
  int base, rem;
};

static bool make_tc_plan(int S, TcPlan* p) {
  if (S < 1) return false;
  const int C = (S + cRows - 1) / cRows;
  if (C > 8) return false;
  p->C = C;
  p->base = S / C;
  p->rem = S % C;
  const int ncmax = p->base + (p->rem ? 1 : 0);
  const int lanes = ncmax + (ncmax + 30) / 31;            // state rows + one special lane per started quadrant
  p->NCP = (lanes + 31) / 32 * 32;
  p->KP = C * p->NCP;
  return p->KP + 2 * 32 <= cTmemCols;                     // A hi + lo = KP columns, two accumulators of >= 32 columns
}

// V buffers + barriers; with 16 clips per sub-batch also the cp.async stage of the step's inputs (both passes sized for the
// backward one: likelihoods + alpha~ of 2 sub-batches x 2 buffers x CN / cParts clips x cEpi threads, and the c values)
static size_t tc_smem_bytes(const TcPlan& p, int CN) {
  return (size_t)4 * p.KP * 2 * CN * 2 + 256 + (CN == 16 ? (size_t)2 * 2 * (CN / cParts) * cEpi * 4 + 2 * cEW * (CN / cParts) * 4 : 0);
}
static size_t tc_packed_words(const TcPlan& p) { return (size_t)p.C * 2 * cM * (p.KP / 2); }

// packed [C][term: hi, lo][128 rows][KP/2 columns] uint32: column c = bf16(k = 2c) | bf16(k = 2c + 1) << 16.
//   forward  (transposed): state row m of shard r = out state j, K = in state i, value A[i][j]; special rows (m % 32 ==
//             31) are 1 at every K position that holds a state: D[special] = sum_k alpha~[k] = the normaliser
//   backward (direct)    : state row = out state i, K = in state j, value A[i][j]; special rows are 1 at the special K
//             positions (where the warps park their partial sums of gamma~): D[special] = sum_j gamma~[j]
__global__ void tc_pack_kernel(const float* __restrict__ A, int S, TcPlan p, bool forward, uint32_t* __restrict__ packed) {
  const int cols = p.KP / 2;
  const size_t total = (size_t)p.C * cM * cols;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(x % cols);
    const int m = (int)((x / cols) % cM);
    const int r = (int)(x / ((size_t)cols * cM));
    const int nc = p.base + (r < p.rem ? 1 : 0);
    const bool m_special = (m & 31) == 31;
    const int mi = m - (m >> 5);                         // state of the shard held by row m
    uint32_t hi = 0, lo = 0;
    for (int h = 0; h < 2; ++h) {
      const int kp = 2 * c + h;
      const int ci = kp / p.NCP, l = kp - ci * p.NCP;
      const int nci = p.base + (ci < p.rem ? 1 : 0);
      const bool k_special = (l & 31) == 31;
      const int li = l - (l >> 5);
      float v = 0.f;
      if (m < p.NCP) {
        if (m_special) {
          v = forward ? ((!k_special && li < nci) ? 1.f : 0.f) : (k_special ? 1.f : 0.f);
        } else if (mi < nc && !k_special && li < nci) {
          const int in = ci * p.base + min(ci, p.rem) + li;
          const int out = r * p.base + min(r, p.rem) + mi;
          v = forward ? A[(size_t)in * S + out] : A[(size_t)out * S + in];
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
__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, float* d) {
  uint32_t* u = reinterpret_cast<uint32_t*>(d);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tc_ld4_nowait(uint32_t taddr, float* d) {
  uint32_t* u = reinterpret_cast<uint32_t*>(d);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// One leader lane of a converged warp (the same lane every time for the full mask).  Issuing tcgen05.mma under
// `if (tid == 0)` makes ptxas wrap EVERY UTCHMMA in an ELECT / BRA.U.ANY loop over the "active threads" (~28 clocks per
// MMA on top of the tensor pipe's floor); under warp-uniform control flow + elect.sync it is a straight instruction stream.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// 1 / x for a normaliser (x > 0; 0 for x == 0: an impossible observation sequence).  One MUFU.RCP: the IEEE division's
// range check + slow-path call is a branch per quotient, and 8-16 of them per step serialise their ~110-clock latencies
// (measured with clock stamps: 900 of the forward epilogue's 2100 clocks, 2400 of the backward one's 4000).  1 ulp is far
// inside the 1e-4 contract: the factor is common to all states of a frame.
__device__ __forceinline__ float rcp_pos(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));            // (the non-ftz form is 8 instructions of range fix-up)
  if (x < 1.1754944e-38f) r = x > 0.f ? 1.f / x : 0.f;              // denormal or zero normaliser: never taken for sane inputs
  return r;
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// 8 fp32 -> 8 bf16 hi (16 bytes) and 8 bf16 lo, with the packed conversion (one F2FP per pair) and the hi terms widened
// back by masking / shifting the packed word
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));     // first source -> upper half
  return r;
}
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
    l[i] = pack_bf16x2(v[2 * i] - h0, v[2 * i + 1] - h1);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void split4(const float* v, uint2& hi, uint2& lo) {
  uint32_t h[2], l[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    h[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
    const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
    l[i] = pack_bf16x2(v[2 * i] - h0, v[2 * i + 1] - h1);
  }
  hi = make_uint2(h[0], h[1]);
  lo = make_uint2(l[0], l[1]);
}

// Sums over the 32 lanes of a warp of CH values per lane by recursive halving (lane pairs exchange the half they do
// not keep: CH/2 + CH/4 + ... + 1 shuffles, then the leftover lane bits).  Returns the total of value index
// column_of_lane(lane) -- every index is held by 32 / CH lanes.
template <int CH>
__device__ __forceinline__ float warp_column_sums(const float* x, int lane) {
  static_assert(CH == 4 || CH == 8, "CH");
  float a[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) a[i] = x[i];
  constexpr int ROUNDS = CH == 16 ? 4 : (CH == 8 ? 3 : 2);
#pragma unroll
  for (int round = 0; round < ROUNDS; ++round) {
    const int bit = 16 >> round;
    const int width = CH >> (round + 1);           // values kept per lane after this round
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < width; ++i) {
      const float keep = upper ? a[i + width] : a[i];
      const float send = upper ? a[i] : a[i + width];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
  float v = a[0];
#pragma unroll
  for (int bit = 16 >> ROUNDS; bit > 0; bit >>= 1) v += __shfl_xor_sync(0xffffffffu, v, bit);
  return v;
}
template <int CH>
__device__ __forceinline__ int column_of_lane(int lane) {      // lane bit 16 -> CH/2, 8 -> CH/4, ...
  return CH == 16 ? ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)
         : CH == 8 ? ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)
                   : ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
}

template <bool BWD, int CN>
__global__ void __launch_bounds__(cThreads, 1)
fb_tc_pass_kernel(const uint32_t* __restrict__ packed, const float* __restrict__ pi, const float* __restrict__ lik,
                  const int32_t* __restrict__ lengths, int B, int T_max, int S, TcPlan p, float* __restrict__ gamma,
                  float* __restrict__ cnorm) {
  constexpr int NN = 2 * CN;           // N of one MMA: hi copy of V in columns [0, CN), lo copy in [CN, 2 CN)
  constexpr int CH = CN / cParts;      // clips per epilogue thread
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int KP = p.KP, NCP = p.NCP;
  constexpr uint32_t LBO = (NN / 8) * 128;              // bytes between groups of 8 K positions
  constexpr uint32_t lo_off = (CN / 8) * 128;           // the lo copy's cores follow the hi copy's within a K group
  const uint32_t buf_bytes = (uint32_t)KP * NN * 2;     // hi + lo
  uint8_t* sV = smem_raw;                               // [2 sub-batches][2 buffers][KP/8][NN/8 cores][8 k][8 n] bf16
  // barriers: [s] own rows written (one arrival per epilogue warp of the sub-batch), [2 + s] MMA batch done, [4 + 2 s + b] peers' rows landed
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + 4 * buf_bytes);
  __shared__ int s_len[2 * CN];
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t C = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int nc_mine = p.base + ((int)rank < p.rem ? 1 : 0);
  const int first_state = (int)rank * p.base + min((int)rank, p.rem);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&s_tmem_base)), "n"(cTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int x = tid; x < (int)(4 * buf_bytes / 16); x += cThreads) reinterpret_cast<uint4*>(sV)[x] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&s_bar[s]), cEW / 2);
      mbar_init(smem_u32(&s_bar[2 + s]), 1);
      mbar_init(smem_u32(&s_bar[4 + 2 * s]), 1);
      mbar_init(smem_u32(&s_bar[5 + 2 * s]), 1);
    }
    mbar_fence_init();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = s_tmem_base;
  if (warp < 8) {
    // A operand -> TMEM: hi copy in columns [0, KP/2), lo copy in [KP/2, KP); row = lane.  Warps 0-3 fill the hi copy,
    // warps 4-7 the lo copy, each for the 32 lanes of its quadrant.
    const int cols = KP / 2;
    const int term = warp >> 2, row = (warp & 3) * 32 + lane;
    const uint32_t tlane = tbase + ((uint32_t)((warp & 3) * 32) << 16);
    const uint4* src = reinterpret_cast<const uint4*>(packed + ((size_t)(rank * 2 + term) * cM + row) * cols);
    for (int x = 0; x < cols / 4; ++x) {
      const uint4 v = src[x];
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                   ::"r"(tlane + term * cols + 4 * x), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the zero-filled V buffers, for the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (C > 1) cluster_sync();

  const uint32_t slice_bytes = (uint32_t)NCP * NN * 2;   // my rows of V (hi and lo cores): contiguous in the canonical layout
  const uint32_t tx_bytes = (C - 1) * slice_bytes;
  // phases the roles have waited for so far, per barrier
  uint32_t ph_own[2] = {0, 0}, ph_mma[2] = {0, 0}, ph_peer[2][2] = {{0, 0}, {0, 0}};

  for (int seq0 = (int)cluster_id_x() * 2 * CN; seq0 < B; seq0 += (int)num_clusters_x() * 2 * CN) {
    __syncthreads();
    // no CTA may start pushing rows of the next round while a peer's last MMA still reads that buffer
    if (C > 1) cluster_sync();
    if (tid < 2 * CN) {
      const int b = seq0 + tid;
      s_len[tid] = b < B ? (lengths ? lengths[b] : T_max) : 0;
    }
    __syncthreads();
    int maxlen = 0;
    for (int n = 0; n < 2 * CN; ++n) maxlen = max(maxlen, s_len[n]);
    if (maxlen == 0) continue;
    const int n_iter = maxlen + 1;      // epilogue steps 0 .. maxlen; MMA batches for steps 1 .. maxlen (the last one only
                                        // delivers the final normaliser (forward) / the last gamma sums (backward))
    if (warp == cEW) {
      // ---------------------------------------- MMA warp -------------------------------------------------------
      // instruction descriptor: D = F32, A = B = BF16, A K-major (TMEM), B MN-major, N >> 3, M >> 4
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(cM >> 4) << 24);
      const int own0 = (int)rank * NCP / 16, own1 = own0 + NCP / 16;
      for (int it = 1; it < n_iter; ++it) {
        const uint32_t b = (uint32_t)it & 1u;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const uint32_t vb = smem_u32(sV) + (uint32_t)(2 * s + b) * buf_bytes;
          const uint32_t d_tmem = tbase + KP + s * NN;
          // descriptor = start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46; a K block of 16 advances the
          // start field by 2 LBO >> 4 (the smem window is < 256 KB: no carry)
          const uint32_t desc_hi32 = (128u >> 4) | (1u << 14);
          const uint32_t desc_lo0 = ((vb >> 4) & 0x3FFF) | (((LBO >> 4) & 0x3FFF) << 16);
          auto issue = [&](int kb0, int kb1, uint32_t acc) {
            uint32_t lo = desc_lo0 + kb0 * ((2 * LBO) >> 4);
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
#ifdef VIT_FB_STAMPS
          const bool stamp = it == 100 && blockIdx.x == 0 && lane == 0;
          long long ck[5];
          if (stamp) ck[0] = clock64();
#define VIT_MSTAMP(i) do { if (stamp) ck[i] = clock64(); } while (0)
#else
#define VIT_MSTAMP(i) do { } while (0)
#endif
          // my own rows of V (written by my epilogue warps, who have also read the accumulator of the previous step)
          mbar_wait_cta(smem_u32(&s_bar[s]), ph_own[s] & 1u);
          VIT_MSTAMP(1);
          ++ph_own[s];
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one_sync()) {
            if (C > 1) mbar_arrive_expect_tx(smem_u32(&s_bar[4 + 2 * s + b]), tx_bytes);
            issue(own0, own1, 0);
          }
          __syncwarp();
          VIT_MSTAMP(2);
          if (C > 1) {
            mbar_wait_cta(smem_u32(&s_bar[4 + 2 * s + b]), ph_peer[s][b] & 1u);    // the peers' rows have landed
            ++ph_peer[s][b];
          }
          VIT_MSTAMP(3);
          if (elect_one_sync()) {
            issue(0, own0, 1);
            issue(own1, KP / 16, 1);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         ::"r"(smem_u32(&s_bar[2 + s])) : "memory");
          }
          __syncwarp();
          VIT_MSTAMP(4);
#ifdef VIT_FB_STAMPS
          if (stamp) printf("fb_tc MMA warp (%s, s=%d) t0 %lld: wait own %lld | issue own %lld | wait peers %lld | issue rest + commit %lld\n",
                            BWD ? "bwd" : "fwd", s, ck[0], ck[1] - ck[0], ck[2] - ck[1], ck[3] - ck[2], ck[4] - ck[3]);
#endif
#undef VIT_MSTAMP
        }
      }
    } else if (warp == cEW + 1) {
      // ---------------------------------------- exchange warp --------------------------------------------------
      // as soon as my epilogue warps have written my rows of the next V, one lane per peer copies the whole slice into
      // that peer's buffer (bulk async DSMEM copy completing on the peer's "rows landed" barrier).  A warp of its own:
      // issued by the epilogue warps the copies sat on their critical path (~600 clocks per step, clock stamps)
      if (C > 1) {
        for (int it = 1; it < n_iter; ++it) {
          const uint32_t b = (uint32_t)it & 1u;
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            mbar_wait_cta(smem_u32(&s_bar[s]), ph_own[s] & 1u);
            ++ph_own[s];
            if (lane < (int)(C - 1)) {
              const uint32_t peer = (rank + 1 + lane) % C;
              const uint32_t src = smem_u32(sV + (size_t)(2 * s + b) * buf_bytes + (size_t)rank * slice_bytes);
              dsmem_bulk_copy(mapa(src, peer), src, slice_bytes, mapa(smem_u32(&s_bar[4 + 2 * s + b]), peer));
            }
            __syncwarp();
          }
        }
      }
    } else {
      // ---------------------------------------- epilogue warps -------------------------------------------------
      // warps 0-3 and 8-11 work on sub-batch X, warps 4-7 and 12-15 on Y: the two sub-batches' epilogues run side by
      // side (all 16 warps on one sub-batch, then on the other, made the step 2 x the epilogue instead of the MMA issue)
      const int quad = warp & 3, s = (warp >> 2) & 1, part = warp >> 3;
      const int col0 = part * CH;                              // my first clip of a sub-batch = my first D / V column
      const int row = quad * 32 + lane;                        // my TMEM lane = my K position within the shard
      const bool special = lane == 31;
      const int jl = row - quad;                               // state of the shard held by my row
      const bool row_ok = !special && jl < nc_mine;
      const int j = first_state + jl;
      const bool quad_live = quad * 32 < NCP;                   // my quadrant holds K positions of the shard at all
      const float pi_j = (!BWD && row_ok) ? pi[j] : 0.f;
      const uint32_t tlane = tbase + ((uint32_t)(quad * 32) << 16);
      const size_t clip_stride = (size_t)T_max * S;
      const long long gamma_delta = reinterpret_cast<const char*>(gamma) - reinterpret_cast<const char*>(lik);
      // lengths of my CH clips of both sub-batches in registers: loads and stores are plain predicated instructions
      int len_c[CH];
#pragma unroll
      for (int n = 0; n < CH; ++n) len_c[n] = s_len[s * CN + col0 + n];
      // frame t of clip n of sub-batch s exists for my row (a row that holds no state never loads or stores)
      auto live = [&](int n, int t_) { return row_ok && t_ < len_c[n]; };
      float held[CH];                                          // backward: gamma~ of the previous step, waiting for its sum
#pragma unroll
      for (int n = 0; n < CH; ++n) held[n] = 0.f;

      // This step's likelihoods (backward: also the stored alpha~ and c_{t+1}) are fetched ONE STEP AHEAD of their use, by
      // cp.async (LDGSTS) into a per-thread slot of shared memory: a whole step (>= 2600 clocks) hides the HBM latency
      // and nothing is held in registers meanwhile.  (Loads into registers -- issued before the wait for the MMAs, or a step
      // ahead -- stall the first shuffle / store after them for the full HBM latency, ~1200 clocks per step measured with
      // clock stamps: the pending loads and everything else share the warp's six scoreboards.)  32 clips per sub-batch
      // leave no shared memory for the stage: plain loads there.
      constexpr bool PRE = CN == 16;
      constexpr int KINDS = BWD ? 2 : 1;
      float* s_in = reinterpret_cast<float*>(smem_raw + 4 * buf_bytes + 256);        // [b][kind][n][cEpi threads]
      float* s_cn = s_in + 2 * KINDS * CH * cEpi;                                     // [b][cEW warps][CH]: c_{t+1} (backward)
      auto in_slot = [&](int b2, int kind, int n) { return s_in + (((b2 * KINDS + kind) * CH + n) * cEpi + tid); };
      auto prefetch_inputs = [&](int it_f) {
        const int t_f = BWD ? maxlen - 1 - it_f : it_f;
        const bool live_step = it_f < maxlen;                  // (the tail step loads nothing)
        if (live_step) {
          const int clip0 = seq0 + s * CN + col0;
          const int b2 = it_f & 1;
          const float* pn = lik + ((size_t)clip0 * T_max + t_f) * S + j;
#pragma unroll
          for (int n = 0; n < CH; ++n, pn += clip_stride) {
            const bool lv = live(n, t_f);
            const uint32_t sz = lv ? 4u : 0u;                  // 0 source bytes: the slot is zero-filled, nothing is read
            const float* src = lv ? pn : lik;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(in_slot(b2, 0, n))), "l"(src), "r"(sz) : "memory");
            if (BWD) {
              const float* srca = lv ? reinterpret_cast<const float*>(reinterpret_cast<const char*>(pn) + gamma_delta) : lik;
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(in_slot(b2, 1, n))), "l"(srca), "r"(sz) : "memory");
            }
          }
          if (BWD && lane < CH) {                              // c_{t+1} of the warp's CH clips: one lane each
            const int len_n = s_len[s * CN + col0 + lane];
            const bool cv = t_f + 1 < len_n;
            const float* src = cv ? cnorm + (size_t)(clip0 + lane) * T_max + t_f + 1 : lik;
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;"
                         ::"r"(smem_u32(s_cn + (b2 * cEW + warp) * CH + lane)), "l"(src), "r"(cv ? 4u : 0u) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");   // (a group per step even when empty: uniform wait counts)
      };
      auto load_inputs = [&](int it_f, float* e, float* al, float* ci) {      // CN = 32: plain loads at the top of the step
        const int t_f = BWD ? maxlen - 1 - it_f : it_f;
        const bool live_step = it_f < maxlen;
        const int clip0 = seq0 + s * CN + col0;
        const float* pn = lik + ((size_t)clip0 * T_max + (live_step ? t_f : 0)) * S + j;
#pragma unroll
        for (int n = 0; n < CH; ++n, pn += clip_stride) {
          const bool lv = live_step && live(n, t_f);
          e[n] = lv ? ld_global_nc_f32(pn) : 0.f;
          if (BWD) {
            al[n] = lv ? ld_global_nc_f32(reinterpret_cast<const float*>(reinterpret_cast<const char*>(pn) + gamma_delta)) : 0.f;
            const bool cv = live_step && t_f + 1 < len_c[n];
            ci[n] = cv ? __ldg(cnorm + (size_t)(clip0 + n) * T_max + t_f + 1) : 0.f;
          }
        }
      };
      if constexpr (PRE) prefetch_inputs(0);

      for (int it = 0; it < n_iter; ++it) {
        const int t = BWD ? maxlen - 1 - it : it;              // forward tail: t = maxlen; backward tail: t = -1
        const bool tail = it == maxlen;
        const uint32_t nxt = (uint32_t)(it + 1) & 1u;          // the V buffer the MMA of step it + 1 reads
        {
          const int clip0 = seq0 + s * CN + col0;
#ifdef VIT_FB_STAMPS
          const bool stamp = it == 100 && blockIdx.x == 0 && tid == 0;
          long long ck[8];
          if (stamp) { ck[0] = clock64(); for (int i_ = 1; i_ < 8; ++i_) ck[i_] = ck[0]; }
#define VIT_ESTAMP(i) do { if (stamp) ck[i] = clock64(); } while (0)
#else
#define VIT_ESTAMP(i) do { } while (0)
#endif
          float e[CH], al[CH], ci[CH];
          if constexpr (PRE) {
            prefetch_inputs(it + 1);                           // (past the last step: an empty group)
            asm volatile("cp.async.wait_group 1;" ::: "memory");   // step it's group; only the one just committed may be pending
            if (BWD) __syncwarp();                             // the c values were fetched by lanes 0 .. CH-1
            const int b2 = it & 1;
#pragma unroll
            for (int n = 0; n < CH; ++n) {
              e[n] = tail ? 0.f : *in_slot(b2, 0, n);
              if (BWD) {
                al[n] = tail ? 0.f : *in_slot(b2, 1, n);
                ci[n] = tail ? 0.f : s_cn[(b2 * cEW + warp) * CH + n];
              }
            }
          } else {
            load_inputs(it, e, al, ci);
          }
          float d[CH];
#pragma unroll
          for (int n = 0; n < CH; ++n) d[n] = 0.f;
          if (it > 0) {
            mbar_wait_cta(smem_u32(&s_bar[2 + s]), ph_mma[0] & 1u);
            ++ph_mma[0];
            VIT_ESTAMP(1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // D[:, 0:CN] = M . V_hi, D[:, CN:2CN] = M . V_lo: my CH clips of both halves
            float dl[CH];
            if constexpr (CH == 4) {
              tc_ld4_nowait(tlane + KP + s * NN + col0, d);
              tc_ld4_nowait(tlane + KP + s * NN + CN + col0, dl);
            } else {
              tc_ld8_nowait(tlane + KP + s * NN + col0, d);
              tc_ld8_nowait(tlane + KP + s * NN + CN + col0, dl);
            }
            tc_ld_wait();
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
#pragma unroll
            for (int n = 0; n < CH; ++n) d[n] += dl[n];
          }
          VIT_ESTAMP(2);
          // ---- the critical path: the next V.  Everything that only goes to HBM (alpha~ / gamma, the normalisers) is
          // issued AFTER my rows have been handed to the MMA warp.
          float v[CH], hnew[CH];
          if (!tail) {
            if (!BWD) {
              // lane 31's row is the normaliser c_{t-1} of my CH clips: alpha~_t = (M alpha~_{t-1} / c_{t-1}) * b_t ;
              // alpha~_0 = pi * b_0
#pragma unroll
              for (int n = 0; n < CH; ++n) {
                const float u = it == 0 ? pi_j : d[n] * rcp_pos(__shfl_sync(0xffffffffu, d[n], 31));
                v[n] = u * e[n];
              }
            } else {
              // beta_t = M w_{t+1} / c_{t+1} (1 at the clip's last frame); gamma~_t = alpha~_t * beta_t; w_t = b_t beta_t
#pragma unroll
              for (int n = 0; n < CH; ++n) {
                const bool lv = live(n, t);
                const float be = (t == len_c[n] - 1) ? 1.f : (it == 0 ? 0.f : d[n] * rcp_pos(ci[n]));
                v[n] = lv ? e[n] * be : 0.f;
                hnew[n] = lv ? al[n] * be : 0.f;
              }
            }
            VIT_ESTAMP(3);
            // my part of my row of the next V (K position rank*NCP + row), bf16 hi and lo
            if (row_ok && quad_live) {
              const uint32_t k = rank * NCP + row;
              uint8_t* vrow = sV + (size_t)(2 * s + nxt) * buf_bytes + (k >> 3) * LBO + (k & 7) * 16 + (col0 >> 3) * 128 + (col0 & 7) * 2;
              if constexpr (CH == 4) {
                uint2 hi, lo;
                split4(v, hi, lo);
                *reinterpret_cast<uint2*>(vrow) = hi;
                *reinterpret_cast<uint2*>(vrow + lo_off) = lo;
              } else {
                uint4 hi, lo;
                split8(v, hi, lo);
                *reinterpret_cast<uint4*>(vrow) = hi;
                *reinterpret_cast<uint4*>(vrow + lo_off) = lo;
              }
            }
            VIT_ESTAMP(4);
            if (BWD && quad_live) {
              // sums of gamma~_t over my warp's 31 states, parked as "the V row of lane 31" (a spare K position): the lane
              // that ends up with the total of clip n writes its two bf16 terms there itself
              const float tot = warp_column_sums<CH>(hnew, lane);
              if ((lane & (32 / CH - 1)) == 0) {
                const int n = column_of_lane<CH>(lane);
                const uint32_t k = rank * NCP + quad * 32 + 31;
                uint8_t* srow = sV + (size_t)(2 * s + nxt) * buf_bytes + (k >> 3) * LBO + (k & 7) * 16 +
                                ((col0 + n) >> 3) * 128 + ((col0 + n) & 7) * 2;
                const __nv_bfloat16 th = __float2bfloat16(tot);
                *reinterpret_cast<__nv_bfloat16*>(srow) = th;
                *reinterpret_cast<__nv_bfloat16*>(srow + lo_off) = __float2bfloat16(tot - __bfloat162float(th));
              }
            }
            VIT_ESTAMP(5);
            fence_proxy_async_smem();
            // my warp has written its rows (and has read the accumulator): tell the MMA warp and the exchange warp
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_bar[s]));
          }
          VIT_ESTAMP(6);
          // ---- off the critical path: HBM stores
          const float* p_t = lik + ((size_t)clip0 * T_max + (tail ? 0 : t)) * S + j;
          float* gn = reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(p_t)) + gamma_delta);
          if (!BWD) {
            if (!tail) {
#pragma unroll
              for (int n = 0; n < CH; ++n, gn += clip_stride)
                if (live(n, t)) st_global_cs_f32(gn, v[n]);
            }
            if (it > 0 && special && quad == 0 && rank == 0) {   // one lane of the cluster records c_{t-1}
              float* pc = cnorm + (size_t)clip0 * T_max + (t - 1);
#pragma unroll
              for (int n = 0; n < CH; ++n, pc += T_max)
                if (t - 1 < len_c[n]) *pc = d[n];
            }
          } else {
            if (it > 0) {
              // gamma_{t+1} = gamma~_{t+1} / s_{t+1}: the values held since the previous step, the sums from lane 31's row
              float* gp = gn + (tail ? 0 : S);                  // (tail: p_t points at frame 0 = t + 1)
#pragma unroll
              for (int n = 0; n < CH; ++n, gp += clip_stride) {
                const float inv = rcp_pos(__shfl_sync(0xffffffffu, d[n], 31));
                if (live(n, t + 1)) st_global_cs_f32(gp, held[n] * inv);
              }
            }
            if (!tail) {
#pragma unroll
              for (int n = 0; n < CH; ++n) held[n] = hnew[n];
            }
          }
          VIT_ESTAMP(7);
#ifdef VIT_FB_STAMPS
          if (stamp) printf("fb_tc epilogue (%s, s=%d) t0 %lld: wait MMA %lld | tmem ld %lld | math %lld | V row %lld | column sums %lld | "
                            "fence + arrive %lld | HBM stores %lld\n", BWD ? "bwd" : "fwd", s, ck[0], ck[1] - ck[0],
                            ck[2] - ck[1], ck[3] - ck[2], ck[4] - ck[3], ck[5] - ck[4], ck[6] - ck[5], ck[7] - ck[6]);
#endif
#undef VIT_ESTAMP
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "n"(cTmemCols) : "memory");
  if (C > 1) cluster_sync();
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
                                 float* __restrict__ loglik, const int* __restrict__ flag, int want);

static void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int C, size_t smem, int grid, cudaStream_t stream) {
  *cfg = cudaLaunchConfig_t{};
  cfg->blockDim = dim3(cThreads);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
  cfg->gridDim = dim3(grid);
}

template <int CN>
static int launch_passes(const TcPlan& p, uint32_t* const* packed, const float* pi, const float* lik, const int32_t* lengths,
                         int B, int T_max, int S, float* gamma, float* cnorm, int max_clusters, cudaStream_t stream) {
  const size_t smem = tc_smem_bytes(p, CN) < 120 * 1024 ? 120 * 1024 : tc_smem_bytes(p, CN);   // one CTA per SM (whole TMEM)
  auto kf = fb_tc_pass_kernel<false, CN>;
  auto kb = fb_tc_pass_kernel<true, CN>;
  VIT_CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VIT_CUDA_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  cluster_config(&cfg, attr, p.C, smem, p.C, stream);
  int occ = 0;
  VIT_CUDA_TRY(cudaOccupancyMaxActiveClusters(&occ, kb, &cfg));
  if (occ < 1) return VIT_ERR_UNSUPPORTED_ALGO;
  if (occ < max_clusters) max_clusters = occ;
  const int want = (B + 2 * CN - 1) / (2 * CN);
  cfg.gridDim = dim3((want < max_clusters ? want : max_clusters) * p.C);
  VIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kf, (const uint32_t*)packed[0], pi, lik, lengths, B, T_max, S, p, gamma, cnorm));
  note_launch();
  VIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kb, (const uint32_t*)packed[1], pi, lik, lengths, B, T_max, S, p, gamma, cnorm));
  note_launch();
  return VIT_OK;
}

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
    tc_pack_kernel<<<grid, 256, 0, stream>>>(A, S, p, /*forward=*/true, packed[0]);
    tc_pack_kernel<<<grid, 256, 0, stream>>>(A, S, p, /*forward=*/false, packed[1]);
    note_launch(2);
    VIT_CUDA_TRY(cudaGetLastError());
  }
  int num_sms = 148, devid = 0;
  VIT_CUDA_TRY(cudaGetDevice(&devid));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, devid));
  const int max_clusters = num_sms / p.C;
  // 16 clips per sub-batch (N = 32 MMAs).  32 clips per sub-batch (N = 64: 43 instead of 27 clocks per MMA for twice
  // the clips) was measured in round 2 and lost at every batch size: its epilogue has no shared memory left for the
  // cp.async input stage (55 vs 45 ms at 4096 clips), so only this instance is built.
  const int rc = launch_passes<16>(p, packed, pi, lik, lengths, B, T_max, S, gamma, cnorm, max_clusters, stream);
  if (rc != VIT_OK) return rc;
  if (loglik) {
    fb_loglik_kernel<<<(B + 3) / 4, 128, 0, stream>>>(cnorm, lengths, B, T_max, loglik, nullptr, 0);
    note_launch();
    VIT_CUDA_TRY(cudaGetLastError());
  }
  return VIT_OK;
}

}  // namespace vit
