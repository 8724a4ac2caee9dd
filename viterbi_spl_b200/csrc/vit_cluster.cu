// VIT_ALGO_CLUSTER -- the throughput path for the pitch-bin state sets (S <= 384: dcnet/msnet/ftanet 321, tonet 361).
//
// The recursion of imm/tf_viterbi.py:97-100 is a max-plus ("tropical") matrix product per frame,
//     delta_t[b][j] = max_i fl32(delta_{t-1}[b][i] + logA^T[j][i]) + logE_t[b][j],
// with M = clips, N = target states j, K = source states i.  Design:
//   * a thread-block cluster of C CTAs owns a sub-batch of MC clips for ALL T steps (persistent kernel);
//   * logA^T is column-sharded over the cluster: CTA r keeps rows j in its shard, all K, resident in shared memory
//     for the whole kernel -- no L2 streaming of the 0.5 MB matrix per step;
//   * delta_{t-1} for all MC clips and all K is replicated in every CTA's shared memory (double buffered); after
//     each step every CTA pushes its [MC][shard] slice of delta_t to its peers with bulk async DSMEM copies that
//     complete on the receiver's mbarrier (no cluster-wide barrier in the step loop);
//   * the inner loop is a register-tiled MB x NJ (clips x targets) max-plus over K with the K range split over KS
//     adjacent lanes: LDS.128 operand feeds, FADD2 (add.rn.f32x2) + FMNMX3 (3-input max) math -- one issue slot
//     per cell -- and a warp-shuffle max across the KS lanes;
//   * no index is tracked in the hot loop: the fp32 delta history (the reference's T1 table) is streamed to HBM and
//     the backtrace kernel resolves argmax_i fl32(delta_{t-1}[i] + logA^T[s_t][i]) only for the ONE state per frame
//     that lies on the decoded path (first maximum wins, exactly np.argmax) -- S work per frame instead of S^2,
//     bit-identical to following the reference's T2 table (imm/tf_viterbi.py:99, :105-107).
#include <cstdlib>

#include "vit_common.cuh"

namespace vit {

struct ClusterPlan {
  int C;      // CTAs per cluster
  int NCP;    // padded target states per shard (multiple of 4); K positions of shard c are [c*NCP, (c+1)*NCP)
  int KP;     // C * NCP, padded K extent (floats per delta row and per logA^T row)
  int base;   // shard c owns base + (c < rem) states starting at c*base + min(c, rem)
  int rem;
  int NCmax;  // base + (rem > 0)
};

// tile configuration of the forward kernel
constexpr int kMB = 8;      // clips per thread tile
constexpr int kNJ = 4;      // target states per thread tile
constexpr int kKS = 4;      // K split across adjacent lanes
constexpr int kJG = 24;     // target groups per CTA -> up to 96 target states per CTA
constexpr int kBG = 2;      // clip groups per pipeline
constexpr int kPipes = 2;   // independent pipelines (warp groups) per CTA, each decoding its own kMC clips
constexpr int kMC = kMB * kBG;                    // 16 clips per pipeline (and per cluster-wide pipeline)
constexpr int kPipeThreads = kBG * kJG * kKS;     // 192 threads = 6 warps per pipeline
constexpr int kThreads = kPipes * kPipeThreads;   // 384
constexpr int kMaxShard = kNJ * kJG;              // 96
constexpr int kNOUT = kMB * kNJ / kKS;            // outputs finalised per thread per step

static bool make_plan(int S, ClusterPlan* p) {
  int C = (S <= kMaxShard) ? 1 : (S <= 2 * kMaxShard ? 2 : 4);
  if (S > 4 * kMaxShard) return false;
  p->C = C;
  p->base = S / C;
  p->rem = S % C;
  p->NCmax = p->base + (p->rem > 0 ? 1 : 0);
  int ncp = (p->NCmax + 3) / 4 * 4;
  // KP = C*NCP must be a multiple of 4*KS (whole chunks) and == 16 (mod 32) floats so that the LDS.128 of a quarter
  // warp (2 logA^T rows x 4 K-split lanes) touches 32 distinct banks
  while ((C * ncp) % (4 * kKS) != 0 || (C * ncp) % 32 != 16) ncp += 4;
  p->NCP = ncp;
  p->KP = C * ncp;
  return true;
}

static size_t forward_smem_bytes(const ClusterPlan& p) {
  return (size_t)(p.NCmax + kPipes * 2 * kMC) * p.KP * sizeof(float) + 64;
}

// Re-lays logA^T [S][S] (dst-major) as [C][NCmax][KP]: shard-major target rows, K positions grouped by shard with
// zero padding (the matching delta pads are -inf, so padded cells never win the max).
__global__ void cluster_pack_logA_kernel(const float* __restrict__ logA_T, int S, ClusterPlan p,
                                         float* __restrict__ packed) {
  const size_t total = (size_t)p.C * p.NCmax * p.KP;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
    const int kp = (int)(x % p.KP);
    const int r = (int)((x / p.KP) % p.NCmax);
    const int c = (int)(x / ((size_t)p.KP * p.NCmax));
    const int ncj = p.base + (c < p.rem ? 1 : 0);
    const int ci = kp / p.NCP, l = kp - ci * p.NCP;
    const int nci = p.base + (ci < p.rem ? 1 : 0);
    float v = 0.f;
    if (r < ncj && l < nci) {
      const int j = c * p.base + min(c, p.rem) + r;
      const int i = ci * p.base + min(ci, p.rem) + l;
      v = logA_T[(size_t)j * S + i];
    }
    packed[x] = v;
  }
}

__device__ __forceinline__ void pipe_bar_sync(int pipe) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + pipe), "n"(kPipeThreads) : "memory");
}

// KP_CT: compile-time padded K extent (0 = take it from the plan at run time); the hot shapes (S = 361 -> 368,
// S = 321 -> 336) get immediate shared-memory offsets in the K loop.
template <int KP_CT>
__global__ void __launch_bounds__(kThreads, 1)
cluster_forward_kernel(const float* __restrict__ packedA, const float* __restrict__ log_pi,
                       const float* __restrict__ log_emis, const int32_t* __restrict__ lengths,
                       int B, int T_max, int S, ClusterPlan p, float* __restrict__ hist, int dev) {
  // dev: timing experiments only (results invalid): 1 = no HBM traffic, 2 = no delta exchange, 4 = no K loop
  constexpr int MB = kMB, NJ = kNJ, KS = kKS, JG = kJG, MC = kMC, NOUT = kNOUT;
  extern __shared__ __align__(128) float smem[];
  const int KP = KP_CT ? KP_CT : p.KP;
  const int KP4 = KP / 4;
  const int NCP = p.NCP;
  float* sA = smem;                                     // [NCmax][KP]             resident logA^T shard
  float* sDall = sA + (size_t)p.NCmax * KP;             // [kPipes][2][MC][KP]     delta double buffers
  uint64_t* s_bar_all = reinterpret_cast<uint64_t*>(sDall + (size_t)kPipes * 2 * MC * KP);   // [kPipes][2]
  __shared__ int s_len_all[kPipes][MC];

  const int tid = threadIdx.x;
  const uint32_t C = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int nc_mine = p.base + ((int)rank < p.rem ? 1 : 0);
  const int j_start = (int)rank * p.base + min((int)rank, p.rem);

  // ---- one-time setup: resident logA^T shard, -inf delta pads, mbarriers -------------------------------------
  {
    const float4* src = reinterpret_cast<const float4*>(packedA + (size_t)rank * p.NCmax * KP);
    float4* dst = reinterpret_cast<float4*>(sA);
    for (int x = tid; x < p.NCmax * KP4; x += kThreads) dst[x] = src[x];
    float4* d4 = reinterpret_cast<float4*>(sDall);
    const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int x = tid; x < kPipes * 2 * MC * KP4; x += kThreads) d4[x] = ninf;
    if (tid == 0) {
      for (int i = 0; i < kPipes * 2; ++i) mbar_init(smem_u32(&s_bar_all[i]), 1);
      mbar_fence_init();
    }
  }
  __syncthreads();
  if (C > 1) cluster_sync();   // every CTA's barriers and buffers exist before any peer copy can land

  // ---- two independent pipelines per CTA: warps 0-5 and 6-11.  Each owns MC clips, its own delta buffers and
  //      mbarriers, and is only ever synchronised with the SAME pipeline of the peer CTAs, so one pipeline's
  //      reduce / store / exchange phase overlaps the other's K loop.
  const int pipe = tid / kPipeThreads;
  const int gt = tid - pipe * kPipeThreads;
  float* sD = sDall + (size_t)pipe * 2 * MC * KP;
  uint64_t* s_bar = s_bar_all + pipe * 2;
  int* s_len = s_len_all[pipe];

  const int q = gt % KS;
  const int jg = (gt / KS) % JG;
  const int bg = gt / (KS * JG);
  // logA^T row offsets (in float4) of this thread's NJ targets; out-of-shard targets alias the last row
  int a_off[NJ];
#pragma unroll
  for (int n = 0; n < NJ; ++n) a_off[n] = min(jg + n * JG, p.NCmax - 1) * KP4 + q;
  const int nchunks = KP4 / KS;
  const uint32_t row_bytes = (uint32_t)NCP * sizeof(float);
  const uint32_t tx_bytes = (C - 1) * MC * row_bytes;
  // after the recursive-halving reduction lane q holds flat outputs [8q, 8q+8): clips 2q, 2q+1 of its tile, all NJ
  // targets
  const int m0 = bg * MB + 2 * q;
  bool o_valid[NJ];
#pragma unroll
  for (int n = 0; n < NJ; ++n) o_valid[n] = (jg + n * JG) < nc_mine;

  uint32_t g = 0;   // pipeline step counter: delta of step g lives in buffer g & 1, guarded by barrier g & 1
  for (int sb = (int)cluster_id_x() * kPipes + pipe; sb * MC < B; sb += (int)num_clusters_x() * kPipes) {
    const int seq0 = sb * MC;
    if (gt < MC) {
      const int b = seq0 + gt;
      s_len[gt] = (b < B) ? (lengths ? lengths[b] : T_max) : 0;
    }
    pipe_bar_sync(pipe);
    int maxlen = 0;
    for (int m = 0; m < MC; ++m) maxlen = max(maxlen, s_len[m]);
    const int len0 = s_len[m0], len1 = s_len[m0 + 1];
    const size_t off0 = (size_t)(seq0 + m0) * T_max * S + (j_start + jg);
    const size_t off1 = off0 + (size_t)T_max * S;

    for (int t = 0; t < maxlen; ++t, ++g) {
      const uint32_t buf = g & 1u;
      // emissions of this step for my outputs: issued first, consumed after the K loop          (hides HBM latency)
      float e[NOUT];
#pragma unroll
      for (int k = 0; k < NOUT; ++k) {
        const int n = k % NJ;
        const bool live = o_valid[n] && t < (k < NJ ? len0 : len1) && !(dev & 1);
        e[k] = live ? ld_global_nc_f32(log_emis + (k < NJ ? off0 : off1) + (size_t)t * S + n * JG) : 0.f;
      }

      float acc[MB * NJ];
      if (t == 0) {
        // T1[0] = log_pi + logE[0]                                                              (imm/tf_viterbi.py:94)
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
          const int n = k % NJ;
          acc[k] = o_valid[n] ? log_pi[j_start + jg + n * JG] : -INFINITY;
        }
      } else {
        // delta_{t-1} from the peers has landed in buffer (g-1)&1 ?
        if (C > 1 && !(dev & 2)) mbar_wait_cta(smem_u32(&s_bar[buf ^ 1u]), ((g - 1) >> 1) & 1u);
#pragma unroll
        for (int i = 0; i < MB * NJ; ++i) acc[i] = -INFINITY;
        const float4* pD = reinterpret_cast<const float4*>(sD + (size_t)(buf ^ 1u) * MC * KP) + (bg * MB) * KP4 + q;
        const float4* pA = reinterpret_cast<const float4*>(sA);
#pragma unroll 2
        for (int c = 0; c < ((dev & 4) ? 1 : nchunks); ++c) {
          float4 d[MB], a[NJ];
#pragma unroll
          for (int b = 0; b < MB; ++b) d[b] = pD[b * KP4 + c * KS];
#pragma unroll
          for (int n = 0; n < NJ; ++n) a[n] = pA[a_off[n] + c * KS];
#pragma unroll
          for (int b = 0; b < MB; ++b)
#pragma unroll
            for (int n = 0; n < NJ; ++n) {
              // Bt[j, i] = T1[t-1][i] + B[j, i]; running max over i                       (:98-99, value part).
              // ptxas fuses each pair of maxes into one FMNMX3; the adds stay scalar FADD.
              float m = acc[b * NJ + n];
              m = fmaxf(m, __fadd_rn(d[b].x, a[n].x));
              m = fmaxf(m, __fadd_rn(d[b].y, a[n].y));
              m = fmaxf(m, __fadd_rn(d[b].z, a[n].z));
              m = fmaxf(m, __fadd_rn(d[b].w, a[n].w));
              acc[b * NJ + n] = m;
            }
        }
        // combine the KS partial maxima by recursive halving over the KS adjacent lanes: each round a lane keeps one
        // half of its values, sends the other half to its partner and folds in what it receives
        int len = MB * NJ;
#pragma unroll
        for (int off = KS / 2; off >= 1; off >>= 1) {
          const bool upper = (q & off) != 0;
          len >>= 1;
#pragma unroll
          for (int i = 0; i < MB * NJ / 2; ++i) {
            if (i < len) {
              const float keep = upper ? acc[i + len] : acc[i];
              const float send = upper ? acc[i] : acc[i + len];
              acc[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
            }
          }
        }
      }

      // T1[t][j] = max + logE[t][j]                                                              (:100)
      float* sDn = sD + (size_t)buf * MC * KP + (size_t)rank * NCP;
#pragma unroll
      for (int k = 0; k < NOUT; ++k) {
        const int n = k % NJ;
        if (o_valid[n]) {
          const float v = __fadd_rn(acc[k], e[k]);
          sDn[(m0 + k / NJ) * KP + jg + n * JG] = v;
          if (t < (k < NJ ? len0 : len1) && !(dev & 1))
            st_global_cs_f32(hist + (k < NJ ? off0 : off1) + (size_t)t * S + n * JG, v);
        }
      }
      if (C > 1 && !(dev & 2)) {
        fence_proxy_async_smem();
        pipe_bar_sync(pipe);
        // all-gather: my [MC][NCP] slice of delta_t -> the same place in the same pipeline's buffer of every peer
        if (gt == 0) mbar_arrive_expect_tx(smem_u32(&s_bar[buf]), tx_bytes);
        if (gt < (int)(C - 1) * MC) {
          const int m = gt % MC;
          const uint32_t peer = (rank + 1 + gt / MC) % C;
          const uint32_t src = smem_u32(sDn + m * KP);
          dsmem_bulk_copy(mapa(src, peer), src, row_bytes, mapa(smem_u32(&s_bar[buf]), peer));
        }
      } else {
        pipe_bar_sync(pipe);
      }
    }
    // drain the last step's exchange.  Once it has completed, every peer has finished the K loop of its last step, so
    // none of them still reads the buffer that step 0 of this pipeline's next sub-batch will overwrite.
    if (C > 1 && maxlen > 0 && !(dev & 2)) mbar_wait_cta(smem_u32(&s_bar[(g - 1) & 1u]), ((g - 1) >> 1) & 1u);
    pipe_bar_sync(pipe);   // s_len is rewritten next
  }
  __syncthreads();
  if (C > 1) cluster_sync();   // no CTA may exit while peers can still address its shared memory
}

// Lazy-argmax backtrace over the delta history.  s_{T-1} = argmax_j delta_{T-1}[j]; then for t = T-1 .. 1
//   s_{t-1} = argmax_i fl32(delta_{t-1}[i] + logA^T[s_t][i])   -- the entry T2[t][s_t] of the reference's table,
// recomputed with the same fp32 add and the same first-maximum rule (imm/tf_viterbi.py:98-99, 103-107).
//
// The walk is a chain of T dependent L2 reads (the logA^T row of the state just found), ~1 us each, so one warp per
// clip is latency-bound (3.5 ms at 1024 x 3000).  It is parallelised over TIME without changing the result:
//   pass 1 (speculative): every clip is cut into segments of kBtSeg frames; a warp walks ONE segment downwards from a
//           guessed state at the segment's top frame e -- argmax_j delta_e[j], the end of the best path into frame e
//           (for the last segment this is the true s_{T-1}).  s_{t-1} is a function of s_t alone, so the segment's
//           states are exact from wherever the guess is right, or from wherever the walk MERGES with the true path.
//   pass 2 (fix-up): a warp per clip goes through the segment boundaries from the top: it steps from the final state
//           above a boundary into the segment below and, while the result differs from what pass 1 wrote, overwrites
//           and keeps walking; at the first agreement the rest of the segment is already exact.  Worst case (no
//           merge) it re-walks everything -- slower, never wrong.
constexpr int kBtSeg = 128;    // frames per segment (multiple of 32: path stores are whole 256-byte lines)
constexpr int kBtWarps = 4;

// one step of the walk: given s_t and delta_{t-1} (registers, lane l holds states l, l+32, ...) -> s_{t-1}
template <int PER>
__device__ __forceinline__ int bt_step(const float* __restrict__ logA_T, int S, int s, const float (&d)[PER], int lane) {
  const float* arow = logA_T + (size_t)s * S;
  float a[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = lane + 32 * k;
    a[k] = (i < S) ? __ldg(arow + i) : 0.f;                           // logA^T rows are re-used: L1/L2-resident
  }
  float best = -INFINITY;
  int arg = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = lane + 32 * k;
    if (i < S) argmax_combine(best, arg, __fadd_rn(d[k], a[k]), i);
  }
  warp_argmax_redux(best, arg);
  return arg;
}

template <int PER>
__device__ __forceinline__ void bt_load_row(const float* __restrict__ row, int S, float (&d)[PER], int lane) {
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = lane + 32 * k;
    d[k] = (i < S) ? ld_global_nc_f32(row + i) : -INFINITY;           // streamed once: keep it out of L1
  }
}

template <int PER>
__global__ void __launch_bounds__(32 * kBtWarps)
backtrace_segments_kernel(const float* __restrict__ logA_T, const float* __restrict__ hist,
                          const int32_t* __restrict__ lengths, int B, int T_max, int S, int nseg_max,
                          int64_t* __restrict__ paths, float* __restrict__ scores) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  // segment-major warp order: the long-latency top segments of all clips are scheduled first
  const int seg = warp / B, b = warp - seg * B;
  if (seg >= nseg_max) return;
  const int len = lengths ? lengths[b] : T_max;
  int64_t* p = paths + (size_t)b * T_max;
  const int lo = seg * kBtSeg;
  // frames past the clip's length are -1; every segment clears its own slice of them
  for (int t = max(lo, len) + lane; t < min(lo + kBtSeg, T_max); t += 32) p[t] = -1;
  if (len <= 0) {
    if (seg == 0 && lane == 0 && scores) scores[b] = -INFINITY;
    return;
  }
  if (lo >= len) return;
  const int e = min(lo + kBtSeg, len) - 1;                            // top frame of this segment
  const float* h = hist + (size_t)b * T_max * S;

  // start state: argmax(T1[e]) -- for the last segment this is `s = np.argmax(T1[-1])` (imm/tf_viterbi.py:103)
  float d[PER];
  bt_load_row<PER>(h + (size_t)e * S, S, d, lane);
  float best = -INFINITY;
  int arg = 0x7fffffff;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int i = lane + 32 * k;
    if (i < S) argmax_combine(best, arg, d[k], i);
  }
  warp_argmax_redux(best, arg);
  if (e == len - 1 && lane == 0 && scores) scores[b] = best;
  int s = arg;
  int64_t mine = 0;                         // lane l keeps states[t] for t % 32 == l until a full line is ready
  if ((e & 31) == lane) mine = s;
  if ((e & 31) == 0) { if (lane == 0) p[e] = mine; }

  if (e > lo) bt_load_row<PER>(h + (size_t)(e - 1) * S, S, d, lane);
  for (int t = e; t > lo; --t) {
    // prefetch delta_{t-2} (independent of s) while the dependent logA^T row is in flight
    float dn[PER];
    if (t - 2 >= lo) bt_load_row<PER>(h + (size_t)(t - 2) * S, S, dn, lane);
    s = bt_step<PER>(logA_T, S, s, d, lane);
    const int tt = t - 1;
    if ((tt & 31) == lane) mine = s;
    if ((tt & 31) == 0) {
      // states[tt .. tt+31] are complete: one coalesced 256-byte store
      if (tt + lane <= e) p[tt + lane] = mine;
    }
#pragma unroll
    for (int k = 0; k < PER; ++k) d[k] = dn[k];
  }
}

template <int PER>
__global__ void __launch_bounds__(32 * kBtWarps)
backtrace_fixup_kernel(const float* __restrict__ logA_T, const float* __restrict__ hist,
                       const int32_t* __restrict__ lengths, int B, int T_max, int S, int64_t* __restrict__ paths) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int len = lengths ? lengths[b] : T_max;
  if (len <= kBtSeg) return;
  int64_t* p = paths + (size_t)b * T_max;
  const float* h = hist + (size_t)b * T_max * S;
  const int nseg = (len + kBtSeg - 1) / kBtSeg;
  for (int k = nseg - 2; k >= 0; --k) {
    int t = (k + 1) * kBtSeg - 1;                                     // top frame of segment k
    int s_next = (int)p[t + 1];                                       // final (segment k+1 is already exact)
    while (t >= 0) {
      float d[PER];
      bt_load_row<PER>(h + (size_t)t * S, S, d, lane);
      const int s = bt_step<PER>(logA_T, S, s_next, d, lane);
      if (s == (int)p[t]) break;                                      // merged: everything below is already exact
      __syncwarp();
      if (lane == 0) p[t] = s;
      __syncwarp();
      s_next = s;
      --t;
    }
  }
}

// ---- backtrace for STRUCTURED matrices (vit_banded.cu / vit_banded_wide.cu) ------------------------------------------
// The dense walk above reads the whole delta row of every frame (1.4 KB at S = 361): 4.4 GB per 1024 x 3000 batch, 23 %
// of a banded step.  For a matrix with band + dense state + constant background c (vit_structure) and a path state
// s != dense state, the candidates whose matrix entry differs from c are the band window |i - s| <= d and the dense
// state; every other source i contributes fl(delta_i + c) <= fl(max_i delta_i + c) (fp32 addition of a constant is
// monotone).  The forward kernels leave max_i delta_t[i] per frame behind (`rowmax`), so:
//     best over window + dense state  >  fl(rowmax + c)   ==>   the argmax (first maximum) is among those candidates,
// found from 2d + 2 loads instead of S; otherwise -- a background source wins or ties -- the step falls back to the full
// scan.  Path states on the dense row (unvoiced) always take the full scan.  Same result as the dense walk, bit for bit.
template <int PER>
__device__ __forceinline__ int bt_step_structured(const float* __restrict__ logA_T, int S, int s,
                                                  const float* __restrict__ row, float rmax, int dband, int jd, float cbg,
                                                  float a_dd, float row_max_other, int lane) {
  if (s != jd) {
    const float* arow = logA_T + (size_t)s * S;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    const int lo = s - dband, ncand = 2 * dband + 2;                    // window, then one slot for the dense state
    for (int k = lane; k < ncand; k += 32) {
      const bool is_dense_slot = k == ncand - 1;
      const int i = is_dense_slot ? jd : lo + k;
      // (the dense state inside the window is taken by its own slot only)
      const bool ok = is_dense_slot ? (jd >= 0) : (i >= 0 && i < S && i != jd);
      if (ok) argmax_combine(best, arg, __fadd_rn(ld_global_nc_f32(row + i), __ldg(arow + i)), i);
    }
    warp_argmax_redux(best, arg);
    if (best > __fadd_rn(rmax, cbg)) return arg;
  } else {
    // on the dense row (unvoiced): staying there wins outright when fl(delta_u + A[u -> u]) beats the bound
    // fl(max_i delta_i + max_{i != u} A[i -> u]) on every other source -- the usual case inside an unvoiced stretch
    const float stay = __fadd_rn(ld_global_nc_f32(row + jd), a_dd);
    if (stay > __fadd_rn(rmax, row_max_other)) return jd;
  }
  float d[PER];
  bt_load_row<PER>(row, S, d, lane);
  return bt_step<PER>(logA_T, S, s, d, lane);
}

// pull what the next step is likely to read into L2: the window around the state just found (or, on the dense row, the
// whole row) of delta_{t-2}, and its row maximum
// `reach` = how far the path can have moved by the time that row is read: (frames of lookahead + 1) * d
__device__ __forceinline__ void bt_prefetch_structured(const float* __restrict__ row, const float* __restrict__ rmax_p,
                                                       int S, int s, int reach, int jd, int lane) {
  if (s == jd) {
    if (lane * 32 < S) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + lane * 32));
  } else {
    const int i = min(max(s - reach + lane * 32, 0), S - 1);
    if (lane * 32 <= 2 * reach + 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + i));
    if (lane == 31 && jd >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + jd));
  }
  if (lane == 30) asm volatile("prefetch.global.L2 [%0];" ::"l"(rmax_p));
}

template <int PER>
__global__ void __launch_bounds__(32 * kBtWarps)
backtrace_segments_structured_kernel(const float* __restrict__ logA_T, const float* __restrict__ hist,
                                     const float* __restrict__ rowmax, const int32_t* __restrict__ lengths, int B,
                                     int T_max, int S, int nseg_max, int dband, int jd, float cbg, float row_max_other,
                                     int la, int64_t* __restrict__ paths, float* __restrict__ scores) {
  const float a_dd = jd >= 0 ? __ldg(logA_T + (size_t)jd * S + jd) : -INFINITY;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int seg = warp / B, b = warp - seg * B;
  if (seg >= nseg_max) return;
  const int len = lengths ? lengths[b] : T_max;
  int64_t* p = paths + (size_t)b * T_max;
  const int lo = seg * kBtSeg;
  for (int t = max(lo, len) + lane; t < min(lo + kBtSeg, T_max); t += 32) p[t] = -1;
  if (len <= 0) {
    if (seg == 0 && lane == 0 && scores) scores[b] = -INFINITY;
    return;
  }
  if (lo >= len) return;
  const int e = min(lo + kBtSeg, len) - 1;                            // top frame of this segment
  const float* h = hist + (size_t)b * T_max * S;
  const float* rm = rowmax + (size_t)b * T_max;

  int s;
  {
    // start state: argmax(T1[e]) (imm/tf_viterbi.py:103 for the last segment; a guess for the others)
    float d[PER];
    bt_load_row<PER>(h + (size_t)e * S, S, d, lane);
    float best = -INFINITY;
    int arg = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int i = lane + 32 * k;
      if (i < S) argmax_combine(best, arg, d[k], i);
    }
    warp_argmax_redux(best, arg);
    if (e == len - 1 && lane == 0 && scores) scores[b] = best;
    s = arg;
  }
  int64_t mine = 0;
  if ((e & 31) == lane) mine = s;
  if ((e & 31) == 0) { if (lane == 0) p[e] = mine; }
  for (int t = e; t > lo; --t) {
    // the row read `la` steps from now, wide enough for wherever the path may be by then
    if (t - 1 - la >= lo) bt_prefetch_structured(h + (size_t)(t - 1 - la) * S, rm + (t - 1 - la), S, s, (la + 1) * dband, jd, lane);
    s = bt_step_structured<PER>(logA_T, S, s, h + (size_t)(t - 1) * S, __ldg(rm + (t - 1)), dband, jd, cbg, a_dd, row_max_other, lane);
    const int tt = t - 1;
    if ((tt & 31) == lane) mine = s;
    if ((tt & 31) == 0) {
      if (tt + lane <= e) p[tt + lane] = mine;
    }
  }
}

template <int PER>
__global__ void __launch_bounds__(32 * kBtWarps)
backtrace_fixup_structured_kernel(const float* __restrict__ logA_T, const float* __restrict__ hist,
                                  const float* __restrict__ rowmax, const int32_t* __restrict__ lengths, int B, int T_max,
                                  int S, int dband, int jd, float cbg, float row_max_other,
                                  int64_t* __restrict__ paths) {
  const float a_dd = jd >= 0 ? __ldg(logA_T + (size_t)jd * S + jd) : -INFINITY;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int len = lengths ? lengths[b] : T_max;
  if (len <= kBtSeg) return;
  int64_t* p = paths + (size_t)b * T_max;
  const float* h = hist + (size_t)b * T_max * S;
  const float* rm = rowmax + (size_t)b * T_max;
  const int nseg = (len + kBtSeg - 1) / kBtSeg;
  for (int k = nseg - 2; k >= 0; --k) {
    int t = (k + 1) * kBtSeg - 1;                                     // top frame of segment k
    int s_next = (int)p[t + 1];                                       // final (segment k+1 is already exact)
    while (t >= 0) {
      const int s = bt_step_structured<PER>(logA_T, S, s_next, h + (size_t)t * S, __ldg(rm + t), dband, jd, cbg, a_dd, row_max_other, lane);
      if (s == (int)p[t]) break;                                      // merged: everything below is already exact
      __syncwarp();
      if (lane == 0) p[t] = s;
      __syncwarp();
      s_next = s;
      --t;
    }
  }
}

// rowmax [B][T_max]: max_i delta_t[i], written by the banded forward kernels for every frame but a clip's last
int launch_structured_backtrace(const float* logA_T, const float* hist, const float* rowmax, const int32_t* lengths,
                                int B, int T_max, int S, const vit_structure* st, int64_t* paths, float* scores,
                                cudaStream_t stream) {
  if (cudaStream_t bt = backtrace_stream_override()) {
    if (bt != stream) {
      cudaEvent_t ev;
      VIT_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      VIT_CUDA_TRY(cudaEventRecord(ev, stream));
      VIT_CUDA_TRY(cudaStreamWaitEvent(bt, ev, 0));
      VIT_CUDA_TRY(cudaEventDestroy(ev));
      stream = bt;
    }
  }
  const int nseg_max = (T_max + kBtSeg - 1) / kBtSeg;
  const long long warps = (long long)B * nseg_max;
  const dim3 block(kBtWarps * 32);
  const dim3 grid1((unsigned)((warps + kBtWarps - 1) / kBtWarps)), grid2((B + kBtWarps - 1) / kBtWarps);
  const char* la_s = getenv("VIT_BT_LOOKAHEAD");            // experiment knob: frames of L2 prefetch lookahead
  const int la = la_s ? atoi(la_s) : 3;
#define VIT_LAUNCH_BTS(PER)                                                                                         \
  do {                                                                                                              \
    backtrace_segments_structured_kernel<PER><<<grid1, block, 0, stream>>>(                                         \
        logA_T, hist, rowmax, lengths, B, T_max, S, nseg_max, st->halfwidth, st->dense_index, st->background,        \
        st->dense_row_max, la, paths, scores);                                                                      \
    note_launch();                                                                                                  \
    if (nseg_max > 1) {                                                                                             \
      backtrace_fixup_structured_kernel<PER><<<grid2, block, 0, stream>>>(                                          \
          logA_T, hist, rowmax, lengths, B, T_max, S, st->halfwidth, st->dense_index, st->background,               \
          st->dense_row_max, paths);                                                                                \
      note_launch();                                                                                                \
    }                                                                                                               \
  } while (0)
  if (S <= 32 * 12) VIT_LAUNCH_BTS(12);
  else if (S <= 32 * 24) VIT_LAUNCH_BTS(24);
  else return VIT_ERR_UNSUPPORTED_ALGO;
#undef VIT_LAUNCH_BTS
  VIT_CUDA_TRY(cudaGetLastError());
  return VIT_OK;
}

// shared with vit_tmem.cu: both forward kernels leave the same fp32 delta history behind
int launch_hist_backtrace(const float* logA_T, const float* hist, const int32_t* lengths, int B, int T_max, int S,
                          int64_t* paths, float* scores, cudaStream_t stream) {
  // vit_decode_opts.backtrace_stream: run the walk on a second stream, ordered after the forward kernel, so that the
  // caller's next forward launch on `stream` overlaps it (the walk is latency-bound and leaves the SMs almost idle)
  if (cudaStream_t bt = backtrace_stream_override()) {
    if (bt != stream) {
      cudaEvent_t ev;
      VIT_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      VIT_CUDA_TRY(cudaEventRecord(ev, stream));
      VIT_CUDA_TRY(cudaStreamWaitEvent(bt, ev, 0));
      VIT_CUDA_TRY(cudaEventDestroy(ev));      // released by the runtime once the wait has been satisfied
      stream = bt;
    }
  }
  const int nseg_max = (T_max + kBtSeg - 1) / kBtSeg;
  const long long warps = (long long)B * nseg_max;
  const dim3 block(kBtWarps * 32);
  const dim3 grid1((unsigned)((warps + kBtWarps - 1) / kBtWarps)), grid2((B + kBtWarps - 1) / kBtWarps);
#define VIT_LAUNCH_BT(PER)                                                                                          \
  do {                                                                                                              \
    backtrace_segments_kernel<PER><<<grid1, block, 0, stream>>>(logA_T, hist, lengths, B, T_max, S, nseg_max, paths, \
                                                                scores);                                            \
    note_launch();                                                                                                  \
    if (nseg_max > 1) {                                                                                             \
      backtrace_fixup_kernel<PER><<<grid2, block, 0, stream>>>(logA_T, hist, lengths, B, T_max, S, paths);          \
      note_launch();                                                                                                \
    }                                                                                                               \
  } while (0)
  if (S <= 32 * 12) VIT_LAUNCH_BT(12);
  else if (S <= 32 * 24) VIT_LAUNCH_BT(24);
  else if (S <= 32 * 48) VIT_LAUNCH_BT(48);
  else return VIT_ERR_UNSUPPORTED_ALGO;
#undef VIT_LAUNCH_BT
  VIT_CUDA_TRY(cudaGetLastError());
  return VIT_OK;
}

size_t cluster_workspace_bytes(int B, int T_max, int S) {
  ClusterPlan p;
  if (!make_plan(S, &p)) return 0;
  size_t bytes = align_up((size_t)p.C * p.NCmax * p.KP * sizeof(float), 256);   // packed logA^T
  bytes += align_up((size_t)B * T_max * S * sizeof(float), 256);                 // delta history (T1 table)
  return bytes;
}

bool cluster_supported(int S) {
  ClusterPlan p;
  if (!make_plan(S, &p)) return false;
  return forward_smem_bytes(p) <= 227 * 1024;
}

int cluster_decode(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths,
                   int B, int T_max, int S, void* workspace, size_t workspace_bytes,
                   int64_t* paths, float* scores, float* delta_out, cudaEvent_t ev0, cudaEvent_t ev1,
                   cudaStream_t stream) {
  ClusterPlan p;
  if (!make_plan(S, &p) || forward_smem_bytes(p) > 227 * 1024) return VIT_ERR_UNSUPPORTED_ALGO;
  if (workspace_bytes < cluster_workspace_bytes(B, T_max, S)) return VIT_ERR_WORKSPACE_TOO_SMALL;
  if (B == 0) return VIT_OK;
  char* ws = (char*)workspace;
  float* packed = (float*)ws;
  ws += align_up((size_t)p.C * p.NCmax * p.KP * sizeof(float), 256);
  float* hist = delta_out ? delta_out : (float*)ws;

  {
    const size_t total = (size_t)p.C * p.NCmax * p.KP;
    const int grid = (int)((total + 255) / 256);
    cluster_pack_logA_kernel<<<grid, 256, 0, stream>>>(logA_T, S, p, packed);
    note_launch();
    VIT_CUDA_TRY(cudaGetLastError());
  }

  const char* dev_s = getenv("VIT_DEV_FLAGS");
  const int dev = dev_s ? atoi(dev_s) : 0;
  auto kern = cluster_forward_kernel<0>;
  if (p.KP == 368) kern = cluster_forward_kernel<368>;        // S = 361 (tonet)
  else if (p.KP == 336) kern = cluster_forward_kernel<336>;   // S = 321 (dcnet / msnet / ftanet)
  const size_t smem = forward_smem_bytes(p);
  VIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent grid: as many clusters as can be co-resident, but no more than there are sub-batches
  int max_clusters = 0;
  cfg.gridDim = dim3(p.C);
  VIT_CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg));
  if (max_clusters < 1) return VIT_ERR_UNSUPPORTED_ALGO;
  const int sub_batches = (B + kMC - 1) / kMC;                       // one per pipeline
  const int want = (sub_batches + kPipes - 1) / kPipes;
  const int n_clusters = want < max_clusters ? want : max_clusters;
  cfg.gridDim = dim3(n_clusters * p.C);
  if (ev0) VIT_CUDA_TRY(cudaEventRecord(ev0, stream));
  VIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, (const float*)packed, log_pi, log_emis, lengths, B, T_max, S, p, hist, dev));
  note_launch();
  if (ev1) VIT_CUDA_TRY(cudaEventRecord(ev1, stream));

  return launch_hist_backtrace(logA_T, hist, lengths, B, T_max, S, paths, scores, stream);
}

}  // namespace vit
