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
constexpr int kMB = 8;    // clips per thread tile
constexpr int kNJ = 4;    // target states per thread tile
constexpr int kKS = 4;    // K split across adjacent lanes
constexpr int kBG = 4;    // clip groups per CTA  -> MC = 32 clips per cluster
constexpr int kJG = 24;   // target groups per CTA -> up to 96 target states per CTA
constexpr int kMC = kMB * kBG;
constexpr int kThreads = kBG * kJG * kKS;   // 384
constexpr int kMaxShard = kNJ * kJG;        // 96

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
  return (size_t)(p.NCmax + 2 * kMC) * p.KP * sizeof(float) + 64;
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

template <bool PACKED>
__global__ void __launch_bounds__(kThreads, 1)
cluster_forward_kernel(const float* __restrict__ packedA, const float* __restrict__ log_pi,
                       const float* __restrict__ log_emis, const int32_t* __restrict__ lengths,
                       int B, int T_max, int S, ClusterPlan p, float* __restrict__ hist) {
  constexpr int MB = kMB, NJ = kNJ, KS = kKS, JG = kJG, MC = kMC;
  constexpr int NOUT = MB * NJ / KS;   // outputs finalised per thread
  extern __shared__ __align__(128) float smem[];
  const int KP = p.KP, NCP = p.NCP;
  const int KP4 = KP / 4;
  float* sA = smem;                                   // [NCmax][KP]
  float* sD = sA + (size_t)p.NCmax * KP;              // [2][MC][KP]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(sD + (size_t)2 * MC * KP);   // [2]
  __shared__ int s_len[MC];

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
    float4* d4 = reinterpret_cast<float4*>(sD);
    const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int x = tid; x < 2 * MC * KP4; x += kThreads) d4[x] = ninf;
    if (tid == 0) {
      mbar_init(smem_u32(&s_bar[0]), 1);
      mbar_init(smem_u32(&s_bar[1]), 1);
      mbar_fence_init();
    }
  }
  __syncthreads();
  if (C > 1) cluster_sync();   // every CTA's barriers and buffers exist before any peer copy can land

  const int q = tid % KS;
  const int jg = (tid / KS) % JG;
  const int bg = tid / (KS * JG);
  // logA^T row offsets (in float4) of this thread's NJ targets; out-of-shard targets alias the last row
  int a_off[NJ];
#pragma unroll
  for (int n = 0; n < NJ; ++n) a_off[n] = min(jg + n * JG, p.NCmax - 1) * KP4 + q;
  const int nchunks = KP4 / KS;
  const uint32_t row_bytes = (uint32_t)NCP * sizeof(float);
  const uint32_t tx_bytes = (C - 1) * MC * row_bytes;

  uint32_t g = 0;   // global step counter: delta of step g lives in buffer g & 1, guarded by barrier g & 1
  for (int sb = cluster_id_x(); sb * MC < B; sb += num_clusters_x()) {
    const int seq0 = sb * MC;
    if (tid < MC) {
      const int b = seq0 + tid;
      s_len[tid] = (b < B) ? (lengths ? lengths[b] : T_max) : 0;
    }
    __syncthreads();
    int maxlen = 0;
    for (int m = 0; m < MC; ++m) maxlen = max(maxlen, s_len[m]);

    // this thread's NOUT outputs: flat = k*KS + q -> clip mb = flat / NJ of its group, target n = flat % NJ
    int o_m[NOUT], o_j[NOUT], o_len[NOUT];
    size_t o_off[NOUT];
    bool o_valid[NOUT];
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
      const int flat = k * KS + q;
      const int m = bg * MB + flat / NJ;
      const int jl = jg + (flat % NJ) * JG;
      o_m[k] = m;
      o_j[k] = jl;
      o_len[k] = s_len[m];
      o_valid[k] = jl < nc_mine;
      o_off[k] = (size_t)(seq0 + m) * T_max * S + (j_start + jl);
    }

    for (int t = 0; t < maxlen; ++t, ++g) {
      const uint32_t buf = g & 1u;
      // emissions of this step for my outputs: issued first, consumed after the K loop          (hides HBM latency)
      float e[NOUT];
#pragma unroll
      for (int k = 0; k < NOUT; ++k)
        e[k] = (o_valid[k] && t < o_len[k]) ? ld_global_nc_f32(log_emis + o_off[k] + (size_t)t * S) : 0.f;

      float acc[MB][NJ];
      if (t == 0) {
        // T1[0] = log_pi + logE[0]                                                              (imm/tf_viterbi.py:94)
#pragma unroll
        for (int b = 0; b < MB; ++b)
#pragma unroll
          for (int n = 0; n < NJ; ++n) {
            const int jl = jg + n * JG;
            acc[b][n] = (jl < nc_mine) ? log_pi[j_start + jl] : -INFINITY;
          }
      } else {
        // delta_{t-1} from the peers has landed in buffer (g-1)&1 ?
        if (C > 1) mbar_wait(smem_u32(&s_bar[buf ^ 1u]), ((g - 1) >> 1) & 1u);
#pragma unroll
        for (int b = 0; b < MB; ++b)
#pragma unroll
          for (int n = 0; n < NJ; ++n) acc[b][n] = -INFINITY;
        const float4* pD = reinterpret_cast<const float4*>(sD + (size_t)(buf ^ 1u) * MC * KP) + (bg * MB) * KP4 + q;
        const float4* pA = reinterpret_cast<const float4*>(sA);
#pragma unroll 2
        for (int c = 0; c < nchunks; ++c) {
          float4 d[MB], a[NJ];
#pragma unroll
          for (int b = 0; b < MB; ++b) d[b] = pD[b * KP4 + c * KS];
#pragma unroll
          for (int n = 0; n < NJ; ++n) a[n] = pA[a_off[n] + c * KS];
#pragma unroll
          for (int b = 0; b < MB; ++b)
#pragma unroll
            for (int n = 0; n < NJ; ++n) {
              // Bt[j, i] = T1[t-1][i] + B[j, i]; running max over i                              (:98-99, value part)
              if (PACKED) {
                float v0, v1, v2, v3;
                fadd2(v0, v1, d[b].x, d[b].y, a[n].x, a[n].y);
                fadd2(v2, v3, d[b].z, d[b].w, a[n].z, a[n].w);
                acc[b][n] = fmax3(acc[b][n], v0, v1);
                acc[b][n] = fmax3(acc[b][n], v2, v3);
              } else {
                acc[b][n] = fmaxf(acc[b][n], __fadd_rn(d[b].x, a[n].x));
                acc[b][n] = fmaxf(acc[b][n], __fadd_rn(d[b].y, a[n].y));
                acc[b][n] = fmaxf(acc[b][n], __fadd_rn(d[b].z, a[n].z));
                acc[b][n] = fmaxf(acc[b][n], __fadd_rn(d[b].w, a[n].w));
              }
            }
        }
        // combine the KS partial maxima (adjacent lanes)
#pragma unroll
        for (int b = 0; b < MB; ++b)
#pragma unroll
          for (int n = 0; n < NJ; ++n)
#pragma unroll
            for (int off = 1; off < KS; off <<= 1)
              acc[b][n] = fmaxf(acc[b][n], __shfl_xor_sync(0xffffffffu, acc[b][n], off));
      }

      // T1[t][j] = max + logE[t][j]                                                              (:100)
      float outv[NOUT];
#pragma unroll
      for (int b = 0; b < MB; ++b)
#pragma unroll
        for (int n = 0; n < NJ; ++n)
          if (((b * NJ + n) % KS) == q) outv[(b * NJ + n) / KS] = acc[b][n];
      float* sDn = sD + (size_t)buf * MC * KP + (size_t)rank * NCP;
#pragma unroll
      for (int k = 0; k < NOUT; ++k) {
        if (o_valid[k]) {
          const float v = __fadd_rn(outv[k], e[k]);
          sDn[o_m[k] * KP + o_j[k]] = v;
          if (t < o_len[k]) st_global_cs_f32(hist + o_off[k] + (size_t)t * S, v);
        }
      }
      if (C > 1) {
        fence_proxy_async_smem();
        __syncthreads();
        // all-gather: my [MC][NCP] slice of delta_t -> the same place in every peer's buffer
        if (tid == 0) mbar_arrive_expect_tx(smem_u32(&s_bar[buf]), tx_bytes);
        if (tid < (int)(C - 1) * MC) {
          const int m = tid % MC;
          const uint32_t peer = (rank + 1 + tid / MC) % C;
          const uint32_t src = smem_u32(sDn + m * KP);
          dsmem_bulk_copy(mapa(src, peer), src, row_bytes, mapa(smem_u32(&s_bar[buf]), peer));
        }
      } else {
        __syncthreads();
      }
    }
    if (C > 1) {
      // drain the last step's exchange, then make sure no peer still reads a buffer the next sub-batch overwrites
      if (maxlen > 0) mbar_wait(smem_u32(&s_bar[(g - 1) & 1u]), ((g - 1) >> 1) & 1u);
      cluster_sync();
    } else {
      __syncthreads();
    }
  }
  if (C > 1) cluster_sync();   // no CTA may exit while peers can still address its shared memory
}

// Lazy-argmax backtrace: one warp per clip.  s_{T-1} = argmax_j delta_{T-1}[j]; then for t = T-1 .. 1
//   s_{t-1} = argmax_i fl32(delta_{t-1}[i] + logA^T[s_t][i])   -- the entry T2[t][s_t] of the reference's table,
// recomputed with the same fp32 add and the same first-maximum rule (imm/tf_viterbi.py:98-99, 103-107).
constexpr int kBtMaxPerLane = 12;   // S <= 384

__global__ void __launch_bounds__(128)
cluster_backtrace_kernel(const float* __restrict__ logA_T, const float* __restrict__ hist,
                         const int32_t* __restrict__ lengths, int B, int T_max, int S,
                         int64_t* __restrict__ paths, float* __restrict__ scores) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const int b = warp;
  const int len = lengths ? lengths[b] : T_max;
  int64_t* p = paths + (size_t)b * T_max;
  for (int t = len + lane; t < T_max; t += 32) p[t] = -1;
  if (len <= 0) {
    if (lane == 0 && scores) scores[b] = -INFINITY;
    return;
  }
  const float* h = hist + (size_t)b * T_max * S;
  const int nper = (S + 31) / 32;

  // s = argmax(T1[-1])
  float best = -INFINITY;
  int arg = 0x7fffffff;
  {
    const float* row = h + (size_t)(len - 1) * S;
    for (int k = 0; k < nper; ++k) {
      const int i = lane + 32 * k;
      if (i < S) argmax_combine(best, arg, row[i], i);
    }
    warp_argmax(best, arg);
  }
  if (lane == 0 && scores) scores[b] = best;
  int s = arg;
  int64_t mine = 0;                         // lane l keeps states[t] for t % 32 == l until a full line is ready
  if (((len - 1) & 31) == lane) mine = s;
  if (((len - 1) & 31) == 0) { if (lane == 0) p[len - 1] = mine; }

  float d[kBtMaxPerLane];
  if (len >= 2) {
    const float* row = h + (size_t)(len - 2) * S;
#pragma unroll
    for (int k = 0; k < kBtMaxPerLane; ++k) {
      const int i = lane + 32 * k;
      d[k] = (i < S) ? row[i] : -INFINITY;
    }
  }
  for (int t = len - 1; t >= 1; --t) {
    const float* arow = logA_T + (size_t)s * S;
    float a[kBtMaxPerLane];
#pragma unroll
    for (int k = 0; k < kBtMaxPerLane; ++k) {
      const int i = lane + 32 * k;
      a[k] = (i < S) ? arow[i] : 0.f;
    }
    // prefetch delta_{t-2} (independent of s) while the dependent logA^T row is in flight
    float dn[kBtMaxPerLane];
    if (t >= 2) {
      const float* row = h + (size_t)(t - 2) * S;
#pragma unroll
      for (int k = 0; k < kBtMaxPerLane; ++k) {
        const int i = lane + 32 * k;
        dn[k] = (i < S) ? row[i] : -INFINITY;
      }
    }
    best = -INFINITY;
    arg = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < kBtMaxPerLane; ++k) {
      const int i = lane + 32 * k;
      if (i < S) argmax_combine(best, arg, __fadd_rn(d[k], a[k]), i);
    }
    warp_argmax(best, arg);
    s = arg;
    const int tt = t - 1;
    if ((tt & 31) == lane) mine = s;
    if ((tt & 31) == 0) {
      // states[tt .. tt+31] are complete: one coalesced 256-byte store
      if (tt + lane < len) p[tt + lane] = mine;
    }
#pragma unroll
    for (int k = 0; k < kBtMaxPerLane; ++k) d[k] = dn[k];
  }
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

  auto kern = cluster_forward_kernel<true>;
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
  const int sub_batches = (B + kMC - 1) / kMC;
  const int n_clusters = sub_batches < max_clusters ? sub_batches : max_clusters;
  cfg.gridDim = dim3(n_clusters * p.C);
  if (ev0) VIT_CUDA_TRY(cudaEventRecord(ev0, stream));
  VIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, (const float*)packed, log_pi, log_emis, lengths, B, T_max, S, p, hist));
  note_launch();
  if (ev1) VIT_CUDA_TRY(cudaEventRecord(ev1, stream));

  const int warps_per_block = 4;
  cluster_backtrace_kernel<<<(B + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, stream>>>(
      logA_T, hist, lengths, B, T_max, S, paths, scores);
  note_launch();
  VIT_CUDA_TRY(cudaGetLastError());
  return VIT_OK;
}

}  // namespace vit
