// Scaled sum-product forward-backward (posterior marginals) on the same HMM the Viterbi kernels decode.
//
// There is NO reference implementation of this pass (correction 2: the reference's SoftMaxViterbi
// classes are max-product decoders); its semantics are the ones defined in oracle/fb_oracle.py
// on the quantities the reference's decoders hold -- A row-stochastic (dcnet/viterbi_transition_matrix.py:81-101), pi
// (dcnet/viterbi_init_probs.py:9-24), b_t = SoftMaxViterbi.observation_probs_fn likelihoods
// (dcnet/softmax_viterbi.py:2530-2579):
//     alpha_0 ~ pi * b_0,  alpha_t ~ (alpha_{t-1} A) * b_t  (c_t = normaliser);   beta_{T-1} = 1,
//     beta_t = A (b_{t+1} * beta_{t+1}) / c_{t+1};   gamma_t = alpha_t * beta_t;   log L = sum_t log c_t.
// Parity: against the float64 oracle, |gamma error| <= 1e-4, log L within 1e-5 relative (tests/test_gpu_fb.py).
//
// Design: the tensor-memory kernel of vit_tmem.cu with the semiring swapped -- one FFMA per cell instead of
// FADD + max -- so everything that made that kernel fast carries over: the transition matrix shard resident in TMEM
// (tcgen05.ld one K chunk ahead), 7 clips x NJ targets per thread, K split over 4 lanes, two pipelines per CTA,
// bulk-async DSMEM exchange between the CTAs of a cluster.  fp32 throughout (TF32 tensor-core products cannot hold
// 1e-4 absolute on gamma without a 3x split; a tcgen05.mma formulation is future work).
//   forward pass : M = A^T; per step u = M a~_{t-1} / c_{t-1};  a~_t = u * b_t (UNNORMALISED, stored in the gamma
//                  buffer);  c_t = sum_j a~_t[j].  The normaliser is one step late on purpose: each CTA parks the
//                  partial sum of its shard in a spare K position of the slice it pushes to its peers anyway, so c_t
//                  costs no extra synchronisation and is applied as a scale after the next K loop (the recursion is
//                  linear).
//   backward pass: M = A; per step beta_t = M w_{t+1} / c_{t+1} (1 at a clip's last frame);
//                  gamma_t = a~_t / c_t * beta_t overwrites a~_t in place;  w_t = b_t * beta_t.
#include <cstdlib>

#include "vit_tmem.cuh"

namespace vit {

template <int NJ>
__device__ __forceinline__ void fma_chunk(float* acc, const float4* d, const float* a) {
#pragma unroll
  for (int b = 0; b < tMB; ++b)
#pragma unroll
    for (int n = 0; n < NJ; ++n) {
      float m = acc[b * NJ + n];
      m = fmaf(d[b].x, a[n * 4 + 0], m);
      m = fmaf(d[b].y, a[n * 4 + 1], m);
      m = fmaf(d[b].z, a[n * 4 + 2], m);
      m = fmaf(d[b].w, a[n * 4 + 3], m);
      acc[b * NJ + n] = m;
    }
}

template <int NJ, bool BWD>
__global__ void __launch_bounds__(tThreads, 1)
fb_pass_kernel(const float* __restrict__ packedT, const float* __restrict__ packedS, const float* __restrict__ pi,
               const float* __restrict__ lik, const int32_t* __restrict__ lengths, int B, int T_max, int S,
               TmemPlan p, float* __restrict__ gamma, float* __restrict__ cnorm, int dev, const int* __restrict__ skip) {
  // (skip: the verdict of fb_conv_detect_kernel when this kernel is the fall-back of the convolution kernels)
  if (skip && *skip) return;
  constexpr int MB = tMB, KS = tKS, NC = NJ * 4;
  constexpr int NPAD = 8 * NJ;
  extern __shared__ __align__(128) float smem[];
  const int KP = p.KP;
  const int KP4 = KP / 4;
  const int NCP = p.NCP;
  float* sDall = smem;                                                   // [tPipes][2][MB][KP]   vector double buffers
  float* sAt = sDall + (size_t)tPipes * 2 * MB * KP;                     // [32*NJ][tail_stride]  matrix K tail
  uint64_t* s_bar_all = reinterpret_cast<uint64_t*>(sAt + (size_t)32 * NJ * p.tail_stride);   // [tPipes][2]
  __shared__ int s_len_all[tPipes][8];
  __shared__ float s_part_all[tPipes][4][8];                             // per-quadrant partial normaliser sums
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t C = cluster_nctarank();
  const uint32_t rank = cluster_ctarank();
  const int nc_mine = p.base + ((int)rank < p.rem ? 1 : 0);
  const int j_start = (int)rank * p.base + min((int)rank, p.rem);
  const int Q = warp & 3, pipe = warp >> 2;
  const int jg = Q * 8 + (lane >> 2), q = lane & 3;

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
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);              // pads contribute 0 * 0 to every sum
    for (int x = tid; x < tPipes * 2 * MB * KP4; x += tThreads) d4[x] = zero;
    if (tid == 0) {
      for (int i = 0; i < tPipes * 2; ++i) mbar_init(smem_u32(&s_bar_all[i]), 1);
      mbar_fence_init();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = __shfl_sync(0xffffffffu, s_tmem_base + ((uint32_t)(Q * 32) << 16), 0);
  if (pipe == 0) {
    const float4* src = reinterpret_cast<const float4*>(packedT + ((size_t)rank * 128 + Q * 32 + lane) * tTmemCols);
    const int ncol4 = p.nchunk_t * NC / 4;
    for (int x = 0; x < ncol4; ++x) tmem_st4(tbase + 4 * x, src[x]);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (C > 1) cluster_sync();

  float* sD = sDall + (size_t)pipe * 2 * MB * KP;
  uint64_t* s_bar = s_bar_all + pipe * 2;
  int* s_len = s_len_all[pipe];
  float (*s_part)[8] = s_part_all[pipe];
  const int gt = tid - pipe * tPipeThreads;
  const int nct = p.nchunk_t, ncs = p.nchunk_s;
  const int ts4 = p.tail_stride / 4;
  const uint32_t row_bytes = (uint32_t)NCP * sizeof(float);
  const uint32_t tx_bytes = (C - 1) * MB * row_bytes;

  const int b0 = 2 * q;
  const bool has_b1 = (b0 + 1) < MB;
  bool n_ok[NJ];
#pragma unroll
  for (int n = 0; n < NJ; ++n) n_ok[n] = (jg + 32 * n) < nc_mine;
  const long long gamma_delta = reinterpret_cast<const char*>(gamma) - reinterpret_cast<const char*>(lik);

  uint32_t g = 0;
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
    const int len0 = s_len[b0], len1 = has_b1 ? s_len[b0 + 1] : 0;
    const int t_first = BWD ? maxlen - 1 : 0;
    const long long t_step = BWD ? -(long long)S : (long long)S;
    // address of lik[clip][t][j_start + jg] for my two clips, moved one frame per step
    const float* pe0 = lik + ((size_t)(seq0 + b0) * T_max + t_first) * S + j_start + jg;
    const float* pe1 = pe0 + (size_t)T_max * S;
    const float* pc0 = cnorm + (size_t)(seq0 + b0) * T_max;             // normalisers of my two clips
    const float* pc1 = pc0 + T_max;

    bool first = true;
    for (int it = 0; it < maxlen; ++it, ++g, pe0 += t_step, pe1 += t_step) {
      const int t = BWD ? maxlen - 1 - it : it;
      const uint32_t buf = g & 1u;
      const bool live0 = t < len0, live1 = t < len1;
      // this step's likelihoods (and, backwards, the stored alpha~ and the normalisers): issued first, used last
      float e[2 * NJ], al[2 * NJ];
      float c_t0 = 1.f, c_t1 = 1.f, c_n0 = 1.f, c_n1 = 1.f;
#pragma unroll
      for (int n = 0; n < NJ; ++n) {
        e[n] = (live0 && n_ok[n]) ? ld_global_nc_f32(pe0 + 32 * n) : 0.f;
        e[NJ + n] = (live1 && n_ok[n]) ? ld_global_nc_f32(pe1 + 32 * n) : 0.f;
        if (BWD) {
          al[n] = (live0 && n_ok[n]) ? ld_global_nc_f32(reinterpret_cast<const float*>(reinterpret_cast<const char*>(pe0) + gamma_delta) + 32 * n) : 0.f;
          al[NJ + n] = (live1 && n_ok[n]) ? ld_global_nc_f32(reinterpret_cast<const float*>(reinterpret_cast<const char*>(pe1) + gamma_delta) + 32 * n) : 0.f;
        }
      }
      if (BWD) {
        if (live0) { c_t0 = pc0[t]; if (t + 1 < len0) c_n0 = pc0[t + 1]; }
        if (live1) { c_t1 = pc1[t]; if (t + 1 < len1) c_n1 = pc1[t + 1]; }
      }

      float acc[NPAD];
      float inv0 = 1.f, inv1 = 1.f;        // scale applied to the matrix-vector product of my two clips
      // (guarded like the forward pass: an impossible observation sequence -- a normaliser of 0 -- gives gamma = 0 from
      // that frame on and log L = -inf, never inf * 0 = NaN)
      const float rc0 = c_t0 > 0.f ? 1.f / c_t0 : 0.f, rc1 = c_t1 > 0.f ? 1.f / c_t1 : 0.f;
      if (first) {
#pragma unroll
        for (int k = 0; k < 2 * NJ; ++k) acc[k] = 0.f;
      } else {
        if (C > 1) mbar_wait_cta(smem_u32(&s_bar[buf ^ 1u]), ((g - 1) >> 1) & 1u);
        const float* prev = sD + (size_t)(buf ^ 1u) * MB * KP;
        if (!BWD) {
          // c_{t-1} = sum over the shards of the partial sums parked in the last K position of every shard's slice
          float c0 = 0.f, c1 = 0.f;
          for (uint32_t r = 0; r < C; ++r) {
            c0 += prev[b0 * KP + r * NCP + NCP - 1];
            if (has_b1) c1 += prev[(b0 + 1) * KP + r * NCP + NCP - 1];
          }
          inv0 = c0 > 0.f ? 1.f / c0 : 0.f;
          inv1 = c1 > 0.f ? 1.f / c1 : 0.f;
          // one thread per clip records c_{t-1}
          if (rank == 0 && jg == 0) {
            if (t - 1 < len0 && seq0 + b0 < B) cnorm[(size_t)(seq0 + b0) * T_max + t - 1] = c0;
            if (has_b1 && t - 1 < len1 && seq0 + b0 + 1 < B) cnorm[(size_t)(seq0 + b0 + 1) * T_max + t - 1] = c1;
          }
        } else {
          inv0 = c_n0 > 0.f ? 1.f / c_n0 : 0.f;
          inv1 = c_n1 > 0.f ? 1.f / c_n1 : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NPAD; ++i) acc[i] = 0.f;
        const float4* pD = reinterpret_cast<const float4*>(prev) + q;
        {
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
            fma_chunk<NJ>(acc, d, a0);
            tmem_wait_ld<NC>(a1);
            if (c + 2 < nct) tmem_ld_chunk<NC>(tbase + (c + 2) * NC, a0);
#pragma unroll
            for (int b = 0; b < MB; ++b) d[b] = pD[b * KP4 + (c + 1) * KS];
            fma_chunk<NJ>(acc, d, a1);
          }
          if (c < nct) {
            tmem_wait_ld<NC>(a0);
#pragma unroll
            for (int b = 0; b < MB; ++b) d[b] = pD[b * KP4 + c * KS];
            fma_chunk<NJ>(acc, d, a0);
          }
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
            fma_chunk<NJ>(acc, d, a0);
          }
        }
        // combine the KS partial sums by recursive halving (as the max-plus kernel does with max)
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
              acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
            }
          }
        }
      }

      float* sDn = sD + (size_t)buf * MB * KP + (size_t)rank * NCP;
      float* sd0 = sDn + b0 * KP + jg;
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int n = 0; n < NJ; ++n) {
        if (n_ok[n]) {
          float v0, v1 = 0.f;
          if (!BWD) {
            // alpha~_t = (M alpha~_{t-1} / c_{t-1}) * b_t ; alpha~_0 = pi * b_0
            const float u0 = first ? pi[j_start + jg + 32 * n] : acc[n] * inv0;
            v0 = u0 * e[n];
            if (live0) st_global_cs_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pe0)) + gamma_delta) + 32 * n, v0);
            if (has_b1) {
              const float u1 = first ? pi[j_start + jg + 32 * n] : acc[NJ + n] * inv1;
              v1 = u1 * e[NJ + n];
              if (live1) st_global_cs_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pe1)) + gamma_delta) + 32 * n, v1);
            }
            sum0 += v0;
            sum1 += v1;
          } else {
            // beta_t = M w_{t+1} / c_{t+1} (1 at the clip's last frame); gamma_t = alpha~_t / c_t * beta_t; w_t = b_t * beta_t
            const float be0 = (t == len0 - 1) ? 1.f : acc[n] * inv0;
            v0 = live0 ? e[n] * be0 : 0.f;
            if (live0) st_global_cs_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pe0)) + gamma_delta) + 32 * n, al[n] * rc0 * be0);
            if (has_b1) {
              const float be1 = (t == len1 - 1) ? 1.f : acc[NJ + n] * inv1;
              v1 = live1 ? e[NJ + n] * be1 : 0.f;
              if (live1) st_global_cs_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pe1)) + gamma_delta) + 32 * n, al[NJ + n] * rc1 * be1);
            }
          }
          sd0[32 * n] = v0;
          if (has_b1) sd0[KP + 32 * n] = v1;
        }
      }
      if (!BWD) {
        // partial normaliser of this quadrant: sum over the 8 target groups of the warp (lanes with equal q)
#pragma unroll
        for (int off = 4; off <= 16; off <<= 1) {
          sum0 += __shfl_xor_sync(0xffffffffu, sum0, off);
          sum1 += __shfl_xor_sync(0xffffffffu, sum1, off);
        }
        if ((lane >> 2) == 0) {
          s_part[Q][b0] = sum0;
          if (has_b1) s_part[Q][b0 + 1] = sum1;
        }
        tpipe_bar_sync(pipe);
        if (gt < MB) sDn[gt * KP + NCP - 1] = (s_part[0][gt] + s_part[1][gt]) + (s_part[2][gt] + s_part[3][gt]);
      }
      first = false;
      if (C > 1) {
        fence_proxy_async_smem();
        tpipe_bar_sync(pipe);
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
    if (C > 1 && !first) mbar_wait_cta(smem_u32(&s_bar[(g - 1) & 1u]), ((g - 1) >> 1) & 1u);
    if (!BWD && !first && rank == 0 && jg == 0) {
      // the last frame's normaliser of the clips that run to the end of this sub-batch
      const float* prev = sD + (size_t)((g - 1) & 1u) * MB * KP;
      float c0 = 0.f, c1 = 0.f;
      for (uint32_t r = 0; r < C; ++r) {
        c0 += prev[b0 * KP + r * NCP + NCP - 1];
        if (has_b1) c1 += prev[(b0 + 1) * KP + r * NCP + NCP - 1];
      }
      if (len0 == maxlen && seq0 + b0 < B) cnorm[(size_t)(seq0 + b0) * T_max + maxlen - 1] = c0;
      if (has_b1 && len1 == maxlen && seq0 + b0 + 1 < B) cnorm[(size_t)(seq0 + b0 + 1) * T_max + maxlen - 1] = c1;
    }
    tpipe_bar_sync(pipe);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "n"(tTmemCols) : "memory");
  if (C > 1) cluster_sync();
}

// log L = sum_t log c_t : one warp per clip, double accumulation (shared with vit_fb_tc.cu)
__global__ void fb_loglik_kernel(const float* __restrict__ cnorm, const int32_t* __restrict__ lengths, int B, int T_max,
                                 float* __restrict__ loglik, const int* __restrict__ flag, int want) {
  if (flag && *flag != want) return;               // (one of two alternative kernel sets ran: only its normalisers count)
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  const int len = lengths ? lengths[b] : T_max;
  double s = 0.0;
  for (int t = lane; t < len; t += 32) s += (double)logf(cnorm[(size_t)b * T_max + t]);
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) loglik[b] = (float)s;
}

static bool make_fb_plan(int S, TmemPlan* p) { return make_tmem_plan(S, p, /*min_pad=*/1); }

bool fb_supported(int S) {
  TmemPlan p;
  return make_fb_plan(S, &p);
}

size_t fb_workspace_bytes(int B, int T_max, int S) {
  TmemPlan p;
  if (!make_fb_plan(S, &p)) return 0;
  size_t bytes = 2 * align_up(tmem_packed_floats(p) * sizeof(float), 256);
  bytes += 2 * align_up(tmem_tail_floats(p) * sizeof(float) + 16, 256);
  bytes += align_up((size_t)(B > 0 ? B : 1) * T_max * sizeof(float), 256);        // normalisers c[b][t]
  return bytes;
}

template <int NJ>
static int launch_fb(const TmemPlan& p, const float* pT_f, const float* pS_f, const float* pT_b, const float* pS_b,
                     const float* pi, const float* lik, const int32_t* lengths, int B, int T_max, int S, float* gamma,
                     float* cnorm, cudaEvent_t ev0, cudaEvent_t ev1, cudaStream_t stream, const int* skip) {
  size_t smem = tmem_smem_bytes(p);
  if (smem < 120 * 1024) smem = 120 * 1024;      // one CTA per SM (each allocates all 512 TMEM columns), see vit_tmem.cu
  auto kf = fb_pass_kernel<NJ, false>;
  auto kb = fb_pass_kernel<NJ, true>;
  VIT_CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VIT_CUDA_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(tThreads);
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
  const int sub_batches = (B + tMB - 1) / tMB;
  const int want = (sub_batches + tPipes - 1) / tPipes;
  const int n_clusters = want < max_clusters ? want : max_clusters;
  cfg.gridDim = dim3(n_clusters * p.C);
  if (ev0) VIT_CUDA_TRY(cudaEventRecord(ev0, stream));
  VIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kf, pT_f, pS_f, pi, lik, lengths, B, T_max, S, p, gamma, cnorm, 0, skip));
  note_launch();
  VIT_CUDA_TRY(cudaLaunchKernelEx(&cfg, kb, pT_b, pS_b, pi, lik, lengths, B, T_max, S, p, gamma, cnorm, 0, skip));
  note_launch();
  if (ev1) VIT_CUDA_TRY(cudaEventRecord(ev1, stream));
  return VIT_OK;
}

int fb_run(const float* A, const float* pi, const float* lik, const int32_t* lengths, int B, int T_max, int S,
           void* workspace, size_t workspace_bytes, float* gamma, float* loglik, cudaEvent_t ev0, cudaEvent_t ev1,
           cudaStream_t stream, const int* skip) {
  TmemPlan p;
  if (!make_fb_plan(S, &p)) return VIT_ERR_UNSUPPORTED_ALGO;
  if (workspace_bytes < fb_workspace_bytes(B, T_max, S)) return VIT_ERR_WORKSPACE_TOO_SMALL;
  if (B == 0) return VIT_OK;
  char* ws = (char*)workspace;
  float* pT[2];
  float* pS[2];
  for (int k = 0; k < 2; ++k) {
    pT[k] = (float*)ws;
    ws += align_up(tmem_packed_floats(p) * sizeof(float), 256);
    pS[k] = (float*)ws;
    ws += align_up(tmem_tail_floats(p) * sizeof(float) + 16, 256);
  }
  float* cnorm = (float*)ws;
  // frames past a clip's length carry gamma = 0 (as a fall-back -- skip != NULL -- the caller has done it, and gamma may
  // already hold the other kernel set's result)
  if (lengths && !skip) VIT_CUDA_TRY(cudaMemsetAsync(gamma, 0, (size_t)B * T_max * S * sizeof(float), stream));
  {
    const size_t total = tmem_packed_floats(p) + tmem_tail_floats(p);
    const int grid = (int)((total + 255) / 256);
    // forward: u[j] = sum_i A[i][j] a[i]  -> rows of the operand are A^T (A is stored source-major: transposed read)
    tmem_pack_kernel<<<grid, 256, 0, stream>>>(A, S, p, pT[0], pS[0], true);
    // backward: u[i] = sum_j A[i][j] w[j] -> rows of the operand are A itself
    tmem_pack_kernel<<<grid, 256, 0, stream>>>(A, S, p, pT[1], pS[1], false);
    note_launch(2);
    VIT_CUDA_TRY(cudaGetLastError());
  }
  int rc;
  switch (p.NJ) {
#define VIT_FB_CASE(N)                                                                                             \
  case N:                                                                                                          \
    rc = launch_fb<N>(p, pT[0], pS[0], pT[1], pS[1], pi, lik, lengths, B, T_max, S, gamma, cnorm, ev0, ev1, stream, skip); \
    break;
    VIT_FB_CASE(1) VIT_FB_CASE(2) VIT_FB_CASE(3) VIT_FB_CASE(4) VIT_FB_CASE(5) VIT_FB_CASE(6)
#undef VIT_FB_CASE
    default: return VIT_ERR_UNSUPPORTED_ALGO;
  }
  if (rc != VIT_OK) return rc;
  if (loglik) {
    fb_loglik_kernel<<<(B + 3) / 4, 128, 0, stream>>>(cnorm, lengths, B, T_max, loglik, skip, 0);
    note_launch();
    VIT_CUDA_TRY(cudaGetLastError());
  }
  return VIT_OK;
}

}  // namespace vit
