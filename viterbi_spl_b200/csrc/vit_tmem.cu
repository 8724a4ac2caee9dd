// VIT_ALGO_TMEM -- the throughput path: logA^T resident in TENSOR MEMORY, all 148 SMs busy.
//
// Same recursion as every other path (imm/tf_viterbi.py:97-100),
//     delta_t[b][j] = max_i fl32(delta_{t-1}[b][i] + logA^T[j][i]) + logE_t[b][j],
// and the same "no index in the hot loop, argmax resolved lazily by the backtrace" split as VIT_ALGO_CLUSTER
// (vit_cluster.cu).  What changes is WHERE the resident logA^T shard lives:
//
//   * 4-CTA clusters strand 16 of the B200's 148 SMs (GPCs hold 16/18/20 SMs; cudaOccupancyMaxActiveClusters = 33),
//     2-CTA clusters pack all 148 -- but half of a 361 x 368 fp32 matrix is 266 KB, more than the 227 KB of shared
//     memory.  The 256 KB of tensor memory (TMEM) per SM are idle in a SIMT kernel, so the shard goes THERE: every
//     thread keeps its own NJ x (K / KS) slice of logA^T in its TMEM lane and streams it into registers with
//     tcgen05.ld (SASS LDTM) one K chunk ahead of the math; the few K chunks that do not fit (512 columns per lane)
//     stay in shared memory.  delta_{t-1} is in shared memory as before (LDS.128, broadcast over the target lanes).
//   * 256 threads = 8 warps per CTA.  Warp w can only address TMEM lane quadrant w & 3, so a quadrant = 8 target
//     groups x 4 K-split lanes, and the two warps of a quadrant serve two independent PIPELINES (7 clips each) that
//     only ever synchronise with the same pipeline of the peer CTAs: one pipeline's reduce / store / exchange phase
//     hides under the other's K loop.  Thread tile = 7 clips x NJ targets (NJ = 6 for S = 361), K split over 4 lanes.
//   * measured inner-loop rate of this layout (tools/microbench_tmem.cu): 61.3 cells/clk/SM = 96 % of the 64 the
//     FADD + FMNMX3 dispatch allows, with 2 warps per scheduler.
//   * exchange of delta_t between the CTAs of a cluster: bulk async DSMEM copies completing on the receiver's
//     mbarrier, exactly as in vit_cluster.cu.
//   * the launch takes a frame range [t_begin, t_end): a later range resumes from the delta history in HBM, so the
//     host API can overlap the host->device copy of the next time slab with the recursion over the current one.
#include <cstdlib>

#include "vit_tmem.cuh"

namespace vit {

// one K chunk of the register-tiled max-plus: Bt[j, i] = T1[t-1][i] + B[j, i]; running max over i   (:98-99, value part).
// ptxas fuses each pair of maxes into one FMNMX3.
template <int NJ>
__device__ __forceinline__ void maxplus_chunk(float* acc, const float4* d, const float* a) {
#pragma unroll
  for (int b = 0; b < tMB; ++b)
#pragma unroll
    for (int n = 0; n < NJ; ++n) {
      float m = acc[b * NJ + n];
      m = fmaxf(m, __fadd_rn(d[b].x, a[n * 4 + 0]));
      m = fmaxf(m, __fadd_rn(d[b].y, a[n * 4 + 1]));
      m = fmaxf(m, __fadd_rn(d[b].z, a[n * 4 + 2]));
      m = fmaxf(m, __fadd_rn(d[b].w, a[n * 4 + 3]));
      acc[b * NJ + n] = m;
    }
}

// KP_CT: compile-time padded K extent (0 = from the plan at run time)
template <int NJ, int KP_CT>
__global__ void __launch_bounds__(tThreads, 1)
tmem_forward_kernel(const float* __restrict__ packedT, const float* __restrict__ packedS,
                    const float* __restrict__ log_pi, const float* __restrict__ log_emis,
                    const int32_t* __restrict__ lengths, int B, int T_max, int S, TmemPlan p,
                    float* __restrict__ hist, int t_begin, int t_end, int dev) {
  // dev: timing experiments only (results invalid): 1 = no HBM traffic, 2 = no delta exchange, 4 = no K loop
  constexpr int MB = tMB, KS = tKS, NC = NJ * 4;
  constexpr int NPAD = 8 * NJ;                       // accumulators per thread: the 7 clips padded to 8 so that two
                                                     // halvings leave lane q with clips 2q and 2q+1, all NJ targets
  constexpr int NOUT = 2 * NJ;                       // outputs finalised per thread per step
  extern __shared__ __align__(128) float smem[];
  const int KP = KP_CT ? KP_CT : p.KP;
  const int KP4 = KP / 4;
  const int NCP = p.NCP;
  float* sDall = smem;                                                   // [tPipes][2][MB][KP]   delta double buffers
  float* sAt = sDall + (size_t)tPipes * 2 * MB * KP;                     // [32*NJ][tail_stride]  logA^T K tail
  uint64_t* s_bar_all = reinterpret_cast<uint64_t*>(sAt + (size_t)32 * NJ * p.tail_stride);   // [tPipes][2]
  __shared__ int s_len_all[tPipes][8];
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t C = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int nc_mine = p.base + ((int)rank < p.rem ? 1 : 0);
  const int j_start = (int)rank * p.base + min((int)rank, p.rem);
  const int Q = warp & 3, pipe = warp >> 2;
  const int jg = Q * 8 + (lane >> 2), q = lane & 3;

  // ---- one-time setup: TMEM allocation, shared-memory tail, -inf delta pads, mbarriers ----------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&s_tmem_base)), "n"(tTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {
    const float4* src = reinterpret_cast<const float4*>(packedS + (size_t)rank * 32 * NJ * p.tail_stride);
    float4* dst = reinterpret_cast<float4*>(sAt);
    for (int x = tid; x < 32 * NJ * p.tail_stride / 4; x += tThreads) dst[x] = src[x];
    float4* d4 = reinterpret_cast<float4*>(sDall);
    const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int x = tid; x < tPipes * 2 * MB * KP4; x += tThreads) d4[x] = ninf;
    if (tid == 0) {
      for (int i = 0; i < tPipes * 2; ++i) mbar_init(smem_u32(&s_bar_all[i]), 1);
      mbar_fence_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // broadcast from lane 0 so that the compiler keeps the TMEM address in a uniform register (no R2UR per tcgen05.ld)
  const uint32_t tbase = __shfl_sync(0xffffffffu, s_tmem_base + ((uint32_t)(Q * 32) << 16), 0);
  if (pipe == 0) {
    // the four warps of pipeline 0 fill their TMEM lane quadrant with this CTA's logA^T shard
    const float4* src = reinterpret_cast<const float4*>(packedT + ((size_t)rank * 128 + Q * 32 + lane) * tTmemCols);
    const int ncol4 = p.nchunk_t * NC / 4;
    for (int x = 0; x < ncol4; ++x) tmem_st4(tbase + 4 * x, src[x]);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (C > 1) cluster_sync();   // every CTA's barriers and buffers exist before any peer copy can land

  float* sD = sDall + (size_t)pipe * 2 * MB * KP;
  uint64_t* s_bar = s_bar_all + pipe * 2;
  int* s_len = s_len_all[pipe];
  const int gt = tid - pipe * tPipeThreads;
  const int nct = p.nchunk_t, ncs = p.nchunk_s;
  const int ts4 = p.tail_stride / 4;
  const uint32_t row_bytes = (uint32_t)NCP * sizeof(float);
  const uint32_t tx_bytes = (C - 1) * MB * row_bytes;

  // after the two halving rounds lane q holds outputs k = c * NJ + n: clip 2q + c (c = 0, 1), target slot jg + 32 n
  const int b0 = 2 * q;
  const bool has_b1 = (b0 + 1) < MB;
  bool n_ok[NJ];
#pragma unroll
  for (int n = 0; n < NJ; ++n) n_ok[n] = (jg + 32 * n) < nc_mine;
  const long long hist_delta = reinterpret_cast<const char*>(hist) - reinterpret_cast<const char*>(log_emis);

  uint32_t g = 0;   // pipeline step counter: delta of step g lives in buffer g & 1, guarded by barrier g & 1
  for (int sb = (int)cluster_id_x() * tPipes + pipe; sb * MB < B; sb += (int)num_clusters_x() * tPipes) {
    const int seq0 = sb * MB;
    if (gt < 8) {
      const int b = seq0 + gt;
      s_len[gt] = (gt < MB && b < B) ? (lengths ? lengths[b] : T_max) : 0;
    }
    tpipe_bar_sync(pipe);
    int maxlen = 0;
#pragma unroll
    for (int m = 0; m < MB; ++m) maxlen = max(maxlen, s_len[m]);
    const int t_stop = min(maxlen, t_end);
    if (t_begin > 0 && t_begin < t_stop) {
      // resume: delta_{t_begin-1} of my clips comes from the history in HBM (all K positions, every shard's slice)
      float* dst = sD + (size_t)((g + 1) & 1u) * MB * KP;
      for (int x = gt; x < MB * S; x += tPipeThreads) {
        const int m = x / S, i = x - m * S;
        const int hi = p.rem * (p.base + 1);
        const int ci = i < hi ? i / (p.base + 1) : p.rem + (i - hi) / p.base;
        const int l = i - (ci * p.base + min(ci, p.rem));
        float v = -INFINITY;
        if (seq0 + m < B && t_begin - 1 < s_len[m]) v = hist[((size_t)(seq0 + m) * T_max + (t_begin - 1)) * S + i];
        dst[m * KP + ci * NCP + l] = v;
      }
      tpipe_bar_sync(pipe);
    }

    // per-lane epilogue state: frames of my two clips, and the address of logE[clip][t][j_start + jg] (advanced by
    // one frame per step; the n-th target is a compile-time offset of 32 n floats, the history sits at a fixed
    // distance from the emissions)
    const int len0 = s_len[b0], len1 = has_b1 ? s_len[b0 + 1] : 0;
    const float* pe0 = log_emis + ((size_t)(seq0 + b0) * T_max + t_begin) * S + j_start + jg;
    const float* pe1 = pe0 + (size_t)T_max * S;

    bool first = true;
    for (int t = t_begin; t < t_stop; ++t, ++g, pe0 += S, pe1 += S) {
      const uint32_t buf = g & 1u;
      const bool live0 = t < len0 && !(dev & 1), live1 = t < len1 && !(dev & 1);
      // emissions of this step for my outputs: issued first, consumed after the K loop          (hides HBM latency)
      float e[NOUT];
#pragma unroll
      for (int n = 0; n < NJ; ++n) {
        e[n] = (live0 && n_ok[n]) ? ld_global_nc_f32(pe0 + 32 * n) : 0.f;
        e[NJ + n] = (live1 && n_ok[n]) ? ld_global_nc_f32(pe1 + 32 * n) : 0.f;
      }

      float acc[NPAD];
      if (t == 0) {
        // T1[0] = log_pi + logE[0]                                                              (imm/tf_viterbi.py:94)
#pragma unroll
        for (int k = 0; k < NOUT; ++k) acc[k] = n_ok[k % NJ] ? log_pi[j_start + jg + 32 * (k % NJ)] : -INFINITY;
      } else {
        // delta_{t-1} from the peers has landed in buffer (g-1)&1 ?  (the first step of a launch reads what the
        // resume code above, or nothing, put there)
        if (C > 1 && !first && !(dev & 2)) mbar_wait_cta(smem_u32(&s_bar[buf ^ 1u]), ((g - 1) >> 1) & 1u);
#pragma unroll
        for (int i = 0; i < NPAD; ++i) acc[i] = -INFINITY;
        const float4* pD = reinterpret_cast<const float4*>(sD + (size_t)(buf ^ 1u) * MB * KP) + q;
        if (!(dev & 4)) {
          // K chunks whose logA^T slice lives in TMEM: tcgen05.ld one chunk ahead of the math
          float a0[NC], a1[NC];
          float4 d[MB];
          tmem_ld_chunk<NC>(tbase, a0);
          int c = 0;
#pragma unroll 1
          for (; c + 1 < nct; c += 2) {
            tmem_wait_ld<NC>(a0);
            tmem_ld_chunk<NC>(tbase + (c + 1) * NC, a1);
#pragma unroll
            for (int b = 0; b < MB; ++b) d[b] = pD[b * KP4 + c * KS];
            maxplus_chunk<NJ>(acc, d, a0);
            tmem_wait_ld<NC>(a1);
            if (c + 2 < nct) tmem_ld_chunk<NC>(tbase + (c + 2) * NC, a0);
#pragma unroll
            for (int b = 0; b < MB; ++b) d[b] = pD[b * KP4 + (c + 1) * KS];
            maxplus_chunk<NJ>(acc, d, a1);
          }
          if (c < nct) {
            tmem_wait_ld<NC>(a0);
#pragma unroll
            for (int b = 0; b < MB; ++b) d[b] = pD[b * KP4 + c * KS];
            maxplus_chunk<NJ>(acc, d, a0);
          }
          // K tail from shared memory
          const float4* pA = reinterpret_cast<const float4*>(sAt) + q;
#pragma unroll 1
          for (int cs = 0; cs < ncs; ++cs) {
#pragma unroll
            for (int n = 0; n < NJ; ++n) {
              const float4 v = pA[(jg + 32 * n) * ts4 + cs * KS];
              a0[n * 4 + 0] = v.x; a0[n * 4 + 1] = v.y; a0[n * 4 + 2] = v.z; a0[n * 4 + 3] = v.w;
            }
#pragma unroll
            for (int b = 0; b < MB; ++b) d[b] = pD[b * KP4 + (nct + cs) * KS];
            maxplus_chunk<NJ>(acc, d, a0);
          }
        }
        // combine the KS partial maxima by recursive halving over the KS adjacent lanes: each round a lane keeps one
        // half of its values, sends the other half to its partner and folds in what it receives
        int len = NPAD;
#pragma unroll
        for (int off = KS / 2; off >= 1; off >>= 1) {
          const bool upper = (q & off) != 0;
          len >>= 1;
#pragma unroll
          for (int i = 0; i < NPAD / 2; ++i) {
            if (i < len) {
              const float keep = upper ? acc[i + len] : acc[i];
              const float send = upper ? acc[i] : acc[i + len];
              acc[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, off));
            }
          }
        }
      }
      first = false;

      // T1[t][j] = max + logE[t][j]                                                              (:100)
      float* sDn = sD + (size_t)buf * MB * KP + (size_t)rank * NCP;
      float* sd0 = sDn + b0 * KP + jg;
#pragma unroll
      for (int n = 0; n < NJ; ++n) {
        if (n_ok[n]) {
          const float v0 = __fadd_rn(acc[n], e[n]);
          sd0[32 * n] = v0;
          if (live0) st_global_cs_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pe0)) + hist_delta) + 32 * n, v0);
          if (has_b1) {
            const float v1 = __fadd_rn(acc[NJ + n], e[NJ + n]);
            sd0[KP + 32 * n] = v1;
            if (live1) st_global_cs_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pe1)) + hist_delta) + 32 * n, v1);
          }
        }
      }
      if (C > 1 && !(dev & 2)) {
        fence_proxy_async_smem();
        tpipe_bar_sync(pipe);
        // all-gather: my [MB][NCP] slice of delta_t -> the same place in the same pipeline's buffer of every peer
        if (gt == 0) mbar_arrive_expect_tx(smem_u32(&s_bar[buf]), tx_bytes);
        if (gt < (int)(C - 1) * MB) {
          const int m = gt % MB;
          const uint32_t peer = (rank + 1 + gt / MB) % C;
          const uint32_t src = smem_u32(sDn + m * KP);
          dsmem_bulk_copy(mapa(src, peer), src, row_bytes, mapa(smem_u32(&s_bar[buf]), peer));
        }
      } else {
        tpipe_bar_sync(pipe);
      }
    }
    // drain the last step's exchange.  Once it has completed, every peer has finished the K loop of its last step, so
    // none of them still reads the buffer that the first step of this pipeline's next sub-batch will overwrite.
    if (C > 1 && !first && !(dev & 2)) mbar_wait_cta(smem_u32(&s_bar[(g - 1) & 1u]), ((g - 1) >> 1) & 1u);
    tpipe_bar_sync(pipe);   // s_len is rewritten next
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "n"(tTmemCols) : "memory");
  if (C > 1) cluster_sync();   // no CTA may exit while peers can still address its shared memory
}

// vit_cluster.cu: lazy-argmax backtrace over the delta history (one warp per clip)
int launch_hist_backtrace(const float* logA_T, const float* hist, const int32_t* lengths, int B, int T_max, int S,
                          int64_t* paths, float* scores, cudaStream_t stream);

size_t tmem_workspace_bytes(int B, int T_max, int S) {
  TmemPlan p;
  if (!make_tmem_plan(S, &p)) return 0;
  size_t bytes = align_up(tmem_packed_floats(p) * sizeof(float), 256);
  bytes += align_up(tmem_tail_floats(p) * sizeof(float) + 16, 256);
  bytes += align_up((size_t)B * T_max * S * sizeof(float), 256);                 // delta history (T1 table)
  return bytes;
}

bool tmem_supported(int S) {
  TmemPlan p;
  return make_tmem_plan(S, &p);
}

template <int NJ>
static auto pick_kernel(int KP) -> decltype(&tmem_forward_kernel<NJ, 0>) {
  if (NJ == 6 && KP == 368) return tmem_forward_kernel<NJ, (NJ == 6 ? 368 : 0)>;   // S = 361 (tonet)
  if (NJ == 6 && KP == 336) return tmem_forward_kernel<NJ, (NJ == 6 ? 336 : 0)>;   // S = 321 (dcnet / msnet / ftanet)
  return tmem_forward_kernel<NJ, 0>;
}

using TmemKernel = decltype(&tmem_forward_kernel<1, 0>);

// kernel instance, launch configuration (cluster attribute in attr[0]) and the number of clusters that can be
// co-resident for a plan: TMEM is allocated whole, so one CTA per SM
static int tmem_geometry(const TmemPlan& p, TmemKernel* kern_out, cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr,
                         int* max_clusters_out) {
  TmemKernel kern = nullptr;
  switch (p.NJ) {
    case 1: kern = pick_kernel<1>(p.KP); break;
    case 2: kern = pick_kernel<2>(p.KP); break;
    case 3: kern = pick_kernel<3>(p.KP); break;
    case 4: kern = pick_kernel<4>(p.KP); break;
    case 5: kern = pick_kernel<5>(p.KP); break;
    case 6: kern = pick_kernel<6>(p.KP); break;
    default: return VIT_ERR_UNSUPPORTED_ALGO;
  }
  // at least 120 KB so that two CTAs can never share an SM: each allocates all 512 TMEM columns, and a CTA waiting
  // for its peer's columns while that peer's cluster waits for ours would deadlock
  size_t smem = tmem_smem_bytes(p);
  if (smem < 120 * 1024) smem = 120 * 1024;
  VIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cfg->blockDim = dim3(tThreads);
  cfg->dynamicSmemBytes = smem;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
  // persistent grid: as many clusters as can be co-resident (TMEM is allocated whole, so 1 CTA per SM)
  int max_clusters = 0;
  cfg->gridDim = dim3(p.C);
  VIT_CUDA_TRY(cudaOccupancyMaxActiveClusters(&max_clusters, kern, cfg));
  if (max_clusters < 1) return VIT_ERR_UNSUPPORTED_ALGO;
  int num_sms = 148, devid = 0;
  VIT_CUDA_TRY(cudaGetDevice(&devid));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, devid));
  if (max_clusters * p.C > num_sms) max_clusters = num_sms / p.C;   // one CTA per SM: TMEM is not shareable here
  *kern_out = kern;
  *max_clusters_out = max_clusters;
  return VIT_OK;
}

// clips one launch keeps in flight with every cluster busy (the wave quantum of viterbi_spl_b200.waves)
int tmem_clips_in_flight(int S, int* out) {
  TmemPlan p;
  if (!make_tmem_plan(S, &p)) return VIT_ERR_UNSUPPORTED_ALGO;
  TmemKernel kern = nullptr;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  int max_clusters = 0;
  const int rc = tmem_geometry(p, &kern, &cfg, attr, &max_clusters);
  if (rc != VIT_OK) return rc;
  *out = max_clusters * tPipes * tMB;
  return VIT_OK;
}

int tmem_decode(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths,
                int B, int T_max, int S, void* workspace, size_t workspace_bytes,
                int64_t* paths, float* scores, float* delta_out, int t_begin, int t_end, bool do_backtrace,
                cudaEvent_t ev0, cudaEvent_t ev1, cudaStream_t stream) {
  TmemPlan p;
  if (!make_tmem_plan(S, &p)) return VIT_ERR_UNSUPPORTED_ALGO;
  if (workspace_bytes < tmem_workspace_bytes(B, T_max, S)) return VIT_ERR_WORKSPACE_TOO_SMALL;
  if (B == 0) return VIT_OK;
  if (t_end > T_max) t_end = T_max;
  if (t_begin < 0 || t_begin > t_end) return VIT_ERR_INVALID_ARGUMENT;
  char* ws = (char*)workspace;
  float* packedT = (float*)ws;
  ws += align_up(tmem_packed_floats(p) * sizeof(float), 256);
  float* packedS = (float*)ws;
  ws += align_up(tmem_tail_floats(p) * sizeof(float) + 16, 256);
  float* hist = delta_out ? delta_out : (float*)ws;

  if (t_begin < t_end) {
    {
      const size_t total = tmem_packed_floats(p) + tmem_tail_floats(p);
      const int grid = (int)((total + 255) / 256);
      tmem_pack_kernel<<<grid, 256, 0, stream>>>(logA_T, S, p, packedT, packedS, false);
      note_launch();
      VIT_CUDA_TRY(cudaGetLastError());
    }
    const char* dev_s = getenv("VIT_DEV_FLAGS");
    const int dev = dev_s ? atoi(dev_s) : 0;
    TmemKernel kern = nullptr;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    int max_clusters = 0;
    const int grc = tmem_geometry(p, &kern, &cfg, attr, &max_clusters);
    if (grc != VIT_OK) return grc;
    cfg.stream = stream;
    const int sub_batches = (B + tMB - 1) / tMB;                      // one per pipeline
    const int want = (sub_batches + tPipes - 1) / tPipes;
    const int n_clusters = want < max_clusters ? want : max_clusters;
    cfg.gridDim = dim3(n_clusters * p.C);
    if (ev0) VIT_CUDA_TRY(cudaEventRecord(ev0, stream));
    VIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, (const float*)packedT, (const float*)packedS, log_pi, log_emis, lengths,
                                    B, T_max, S, p, hist, t_begin, t_end, dev));
    note_launch();
    if (ev1) VIT_CUDA_TRY(cudaEventRecord(ev1, stream));
  }
  if (do_backtrace) return launch_hist_backtrace(logA_T, hist, lengths, B, T_max, S, paths, scores, stream);
  return VIT_OK;
}

}  // namespace vit
