// VIT_ALGO_STREAM -- dense max-plus recursion with logA^T STREAMED from L2 through a TMA ring (BASELINE.json config 3:
// "fine-grid state set (~722 states), 4096 clips x 10,000 frames, logA streamed from L2 via TMA").
//
// Same recursion and the same value / lazy-argmax split as vit_tmem.cu (imm/tf_viterbi.py:97-100); what changes is where
// logA^T lives.  The tensor-memory kernel keeps a shard of the matrix resident per SM and therefore needs clusters of
// ceil(S / 192) CTAs that exchange delta_t every step: at S = 722 that is 6-CTA clusters, of which only 22 fit the chip
// (132 of 148 SMs), and a 5-peer all-gather per step.  Here NOTHING is resident and nothing is exchanged:
//
//   * one CTA per SM owns 14 clips for all T steps (two pipelines of 4 warps x 7 clips, thread tile 7 clips x 6 targets,
//     K split over 4 lanes -- the tile of vit_tmem.cu) and computes ALL S targets of its clips;
//   * a ninth warp is the producer: one thread streams the matrix, pre-packed into 24 KB tiles of 192 targets x 32
//     sources, with cp.async.bulk (TMA, SASS UBLKCP) into a 4-stage shared-memory ring, full/empty mbarriers per stage.
//     The tile sequence does not depend on the step, so the producer simply runs ahead (also across step boundaries,
//     under the consumers' epilogues);
//   * both pipelines consume the same tile (LDS.128, conflict-free by construction of the packed layout), so a step
//     moves S^2 x 4 bytes from L2 per 14 clips: 2.1 MB per ~140 k clocks at S = 722 = 35 % of the measured L2 -> SM
//     rate of the chip with all 148 SMs streaming (6300 B/clk);
//   * delta_{t-1} of the CTA's clips stays in shared memory (double buffered), never leaves the SM.
//
// Needs >= 14 clips per SM to fill the machine (B >= 2072 on 148 SMs): the throughput path for big batches and big
// state sets; VIT_ALGO_AUTO takes it for dense matrices with S > 384 once the batch is that large.
#include <cstdlib>

#include "vit_tmem.cuh"

namespace vit {

constexpr int sMB = tMB;                    // 7 clips per pipeline = per thread tile
constexpr int sPipes = 2;
constexpr int sNJ = 6;                      // targets per thread
constexpr int sJB = 32 * sNJ;               // 192 targets per target block
constexpr int sKT = 32;                     // K positions per tile = 2 chunks of 16 (4 per K-split lane)
constexpr int sTileFloats = sJB * sKT;      // 6144 floats = 24 KB
constexpr int sTileBytes = sTileFloats * 4;
constexpr int sStages = 4;
constexpr int sConsumers = 256;             // 8 warps
constexpr int sThreads = sConsumers + 32;   // + the producer warp
constexpr int sClipsPerCta = sMB * sPipes;  // 14

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// TMA bulk copy global -> this CTA's shared memory, completing `bytes` on a local mbarrier (all 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void consumers_bar_sync() {
  asm volatile("bar.sync 3, %0;" ::"n"(sConsumers) : "memory");
}

// packed [n_jb][n_kt][c = 2][n = 6][jg = 32][q = 4][kk = 4]:
//   logA^T[target jb*192 + jg + 32 n][source kt*32 + 16 c + 4 q + kk], 0 where either index is past S (the matching
//   delta pads are -inf, so padded cells never win the max)
__global__ void stream_pack_kernel(const float* __restrict__ logA_T, int S, int n_jb, int n_kt, float* __restrict__ packed) {
  const size_t total = (size_t)n_jb * n_kt * sTileFloats;
  for (size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(x % sTileFloats);
    const size_t tile = x / sTileFloats;
    const int kt = (int)(tile % n_kt), jb = (int)(tile / n_kt);
    const int kk = w & 3, q = (w >> 2) & 3, jg = (w >> 4) & 31, cn = w >> 9;   // cn = c * 6 + n
    const int n = cn % sNJ, c = cn / sNJ;
    const int j = jb * sJB + jg + 32 * n, i = kt * sKT + 16 * c + 4 * q + kk;
    packed[x] = (j < S && i < S) ? logA_T[(size_t)j * S + i] : 0.f;
  }
}

__device__ __forceinline__ void stream_maxplus_chunk(float* acc, const float4* d, const float* a) {
#pragma unroll
  for (int b = 0; b < sMB; ++b)
#pragma unroll
    for (int n = 0; n < sNJ; ++n) {
      float m = acc[b * sNJ + n];
      m = fmaxf(m, __fadd_rn(d[b].x, a[n * 4 + 0]));
      m = fmaxf(m, __fadd_rn(d[b].y, a[n * 4 + 1]));
      m = fmaxf(m, __fadd_rn(d[b].z, a[n * 4 + 2]));
      m = fmaxf(m, __fadd_rn(d[b].w, a[n * 4 + 3]));
      acc[b * sNJ + n] = m;
    }
}

__global__ void __launch_bounds__(sThreads, 1)
stream_forward_kernel(const float* __restrict__ packed, const float* __restrict__ log_pi,
                      const float* __restrict__ log_emis, const int32_t* __restrict__ lengths, int B, int T_max, int S,
                      int KP, int n_jb, int n_kt, float* __restrict__ hist, int t_begin, int t_end) {
  constexpr int MB = sMB, NJ = sNJ, KS = 4;
  constexpr int NPAD = 8 * NJ;                       // accumulators per thread (7 clips padded to 8 for the halving)
  constexpr int NOUT = 2 * NJ;                       // outputs finalised per thread per target block
  extern __shared__ __align__(128) float smem[];
  float* sRing = smem;                                                    // [sStages][sTileFloats]
  float* sDall = sRing + (size_t)sStages * sTileFloats;                   // [sPipes][2][MB][KP]
  uint64_t* s_full = reinterpret_cast<uint64_t*>(sDall + (size_t)sPipes * 2 * MB * KP);
  uint64_t* s_empty = s_full + sStages;
  __shared__ int s_len[16];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KP4 = KP / 4;
  {
    float4* d4 = reinterpret_cast<float4*>(sDall);
    const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int x = tid; x < sPipes * 2 * MB * KP4; x += sThreads) d4[x] = ninf;
    if (tid == 0) {
      for (int i = 0; i < sStages; ++i) {
        mbar_init(smem_u32(&s_full[i]), 1);                 // the producer's arrive.expect_tx
        mbar_init(smem_u32(&s_empty[i]), sConsumers / 32);  // one arrive per consumer warp
      }
      mbar_fence_init();
    }
  }
  __syncthreads();
  const int tiles_per_step = n_jb * n_kt;

  if (warp == sConsumers / 32) {
    // ---- producer: the same (sub-batch, step, tile) sequence as the consumers, as far ahead as the ring allows ------
    if (lane == 0) {
      uint32_t it = 0;
      for (int sb = blockIdx.x; sb * sClipsPerCta < B; sb += gridDim.x) {
        int maxlen = 0;
        for (int m = 0; m < sClipsPerCta; ++m) {
          const int b = sb * sClipsPerCta + m;
          if (b < B) maxlen = max(maxlen, lengths ? lengths[b] : T_max);
        }
        const int t_stop = min(maxlen, t_end);
        const int n_steps = t_stop - max(t_begin, 1);       // step 0 reads no matrix
        for (int s = 0; s < n_steps; ++s)
          for (int tile = 0; tile < tiles_per_step; ++tile, ++it) {
            const uint32_t stage = it % sStages, par = ((it / sStages) & 1u) ^ 1u;
            mbar_wait_cta(smem_u32(&s_empty[stage]), par);
            mbar_arrive_expect_tx(smem_u32(&s_full[stage]), sTileBytes);
            bulk_g2s(smem_u32(sRing + (size_t)stage * sTileFloats), packed + (size_t)tile * sTileFloats, sTileBytes,
                     smem_u32(&s_full[stage]));
          }
      }
    }
    return;
  }

  // ---- consumers -----------------------------------------------------------------------------------------------------
  const int Q = warp & 3, pipe = warp >> 2;
  const int jg = Q * 8 + (lane >> 2), q = lane & 3;
  const int gt = tid - pipe * tPipeThreads;
  float* sD = sDall + (size_t)pipe * 2 * MB * KP;
  const int b0 = 2 * q;
  const bool has_b1 = (b0 + 1) < MB;
  const long long hist_delta = reinterpret_cast<const char*>(hist) - reinterpret_cast<const char*>(log_emis);

  uint32_t it = 0;      // ring position: the same count on every consumer warp and on the producer
  uint32_t g = 0;       // step counter: delta of step g lives in buffer g & 1
  for (int sb = blockIdx.x; sb * sClipsPerCta < B; sb += gridDim.x) {
    consumers_bar_sync();                               // everybody is done with the previous s_len
    if (tid < sClipsPerCta) {
      const int b = sb * sClipsPerCta + tid;
      s_len[tid] = b < B ? (lengths ? lengths[b] : T_max) : 0;
    }
    consumers_bar_sync();
    int maxlen = 0;
#pragma unroll
    for (int m = 0; m < sClipsPerCta; ++m) maxlen = max(maxlen, s_len[m]);
    const int t_stop = min(maxlen, t_end);
    const int seq0 = sb * sClipsPerCta + pipe * MB;
    if (t_begin > 0 && t_begin < t_stop) {
      // resume: delta_{t_begin-1} of my clips comes back from the history in HBM
      float* dst = sD + (size_t)((g + 1) & 1u) * MB * KP;
      for (int x = gt; x < MB * S; x += tPipeThreads) {
        const int m = x / S, i = x - m * S;
        float v = -INFINITY;
        if (seq0 + m < B && t_begin - 1 < s_len[pipe * MB + m]) v = hist[((size_t)(seq0 + m) * T_max + (t_begin - 1)) * S + i];
        dst[m * KP + i] = v;
      }
      tpipe_bar_sync(pipe);
    }
    const int len0 = s_len[pipe * MB + b0], len1 = has_b1 ? s_len[pipe * MB + b0 + 1] : 0;
    // logE[clip][t][jg] of my two clips, advanced one frame per step (clips past the batch are never live)
    const float* pe0 = log_emis + ((size_t)min(seq0 + b0, B - 1) * T_max + t_begin) * S + jg;
    const float* pe1 = log_emis + ((size_t)min(seq0 + b0 + 1, B - 1) * T_max + t_begin) * S + jg;

    for (int t = t_begin; t < t_stop; ++t, ++g, pe0 += S, pe1 += S) {
      const uint32_t buf = g & 1u;
      const bool live0 = t < len0, live1 = t < len1;
      const float4* pD = reinterpret_cast<const float4*>(sD + (size_t)(buf ^ 1u) * MB * KP) + q;
      float* sDn = sD + (size_t)buf * MB * KP;
      for (int jb = 0; jb < n_jb; ++jb) {
        const int jbase = jb * sJB;
        bool n_ok[NJ];
#pragma unroll
        for (int n = 0; n < NJ; ++n) n_ok[n] = (jbase + jg + 32 * n) < S;
        // emissions of this target block for my outputs: issued first, consumed after the K loop
        float e[NOUT];
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
          e[n] = (live0 && n_ok[n]) ? ld_global_nc_f32(pe0 + jbase + 32 * n) : 0.f;
          e[NJ + n] = (live1 && n_ok[n]) ? ld_global_nc_f32(pe1 + jbase + 32 * n) : 0.f;
        }
        float acc[NPAD];
        if (t == 0) {
          // T1[0] = log_pi + logE[0]                                                            (imm/tf_viterbi.py:94)
#pragma unroll
          for (int k = 0; k < NOUT; ++k) acc[k] = n_ok[k % NJ] ? log_pi[jbase + jg + 32 * (k % NJ)] : -INFINITY;
        } else {
#pragma unroll
          for (int i = 0; i < NPAD; ++i) acc[i] = -INFINITY;
#pragma unroll 1
          for (int kt = 0; kt < n_kt; ++kt, ++it) {
            const uint32_t stage = it % sStages, par = (it / sStages) & 1u;
            mbar_wait_cta(smem_u32(&s_full[stage]), par);
            const float4* tile4 = reinterpret_cast<const float4*>(sRing + (size_t)stage * sTileFloats) + jg * 4 + q;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float a[NJ * 4];
              float4 d[MB];
#pragma unroll
              for (int n = 0; n < NJ; ++n) {
                const float4 v = tile4[(c * NJ + n) * 128];
                a[n * 4 + 0] = v.x; a[n * 4 + 1] = v.y; a[n * 4 + 2] = v.z; a[n * 4 + 3] = v.w;
              }
#pragma unroll
              for (int b = 0; b < MB; ++b) d[b] = pD[b * KP4 + (kt * 2 + c) * KS];
              stream_maxplus_chunk(acc, d, a);
            }
            // every value read from the stage has been consumed by the math above: hand the stage back
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_empty[stage]));
          }
          // combine the 4 K-split partial maxima by recursive halving (as vit_tmem.cu)
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
        // T1[t][j] = max + logE[t][j]                                                            (:100)
        float* sd0 = sDn + b0 * KP + jbase + jg;
#pragma unroll
        for (int n = 0; n < NJ; ++n) {
          if (n_ok[n]) {
            const float v0 = __fadd_rn(acc[n], e[n]);
            sd0[32 * n] = v0;
            if (live0) st_global_cs_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pe0)) + hist_delta) + jbase + 32 * n, v0);
            if (has_b1) {
              const float v1 = __fadd_rn(acc[NJ + n], e[NJ + n]);
              sd0[KP + 32 * n] = v1;
              if (live1) st_global_cs_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pe1)) + hist_delta) + jbase + 32 * n, v1);
            }
          }
        }
      }
      tpipe_bar_sync(pipe);      // delta_t of my pipeline's clips is complete
    }
  }
}

// vit_cluster.cu
int launch_hist_backtrace(const float* logA_T, const float* hist, const int32_t* lengths, int B, int T_max, int S,
                          int64_t* paths, float* scores, cudaStream_t stream);

static size_t stream_smem_bytes(int KP) {
  return (size_t)sStages * sTileBytes + (size_t)sPipes * 2 * sMB * KP * sizeof(float) + 2 * sStages * sizeof(uint64_t) + 64;
}
static void stream_shape(int S, int* n_jb, int* n_kt, int* KP) {
  *n_jb = (S + sJB - 1) / sJB;
  *n_kt = (S + sKT - 1) / sKT;
  *KP = *n_kt * sKT;
}

bool stream_supported(int S) {
  if (S < 1 || S > 32 * 48) return false;                 // (the shared backtrace handles S <= 1536)
  int n_jb, n_kt, KP;
  stream_shape(S, &n_jb, &n_kt, &KP);
  return stream_smem_bytes(KP) <= 227 * 1024;
}

size_t stream_workspace_bytes(int B, int T_max, int S) {
  int n_jb, n_kt, KP;
  stream_shape(S, &n_jb, &n_kt, &KP);
  return align_up((size_t)n_jb * n_kt * sTileBytes, 256) + align_up((size_t)B * T_max * S * sizeof(float), 256);
}

int stream_clips_in_flight(int* out) {
  int num_sms = 148, dev = 0;
  VIT_CUDA_TRY(cudaGetDevice(&dev));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  *out = num_sms * sClipsPerCta;
  return VIT_OK;
}

int stream_decode(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths, int B,
                  int T_max, int S, void* workspace, size_t workspace_bytes, int64_t* paths, float* scores,
                  float* delta_out, int t_begin, int t_end, bool do_backtrace, cudaEvent_t ev0, cudaEvent_t ev1,
                  cudaStream_t stream) {
  if (!stream_supported(S)) return VIT_ERR_UNSUPPORTED_ALGO;
  if (workspace_bytes < stream_workspace_bytes(B, T_max, S)) return VIT_ERR_WORKSPACE_TOO_SMALL;
  if (B == 0) return VIT_OK;
  if (t_end > T_max) t_end = T_max;
  if (t_begin < 0 || t_begin > t_end) return VIT_ERR_INVALID_ARGUMENT;
  int n_jb, n_kt, KP;
  stream_shape(S, &n_jb, &n_kt, &KP);
  float* packed = (float*)workspace;
  float* hist = delta_out ? delta_out : (float*)((char*)workspace + align_up((size_t)n_jb * n_kt * sTileBytes, 256));
  if (t_begin < t_end) {
    stream_pack_kernel<<<148, 256, 0, stream>>>(logA_T, S, n_jb, n_kt, packed);
    note_launch();
    VIT_CUDA_TRY(cudaGetLastError());
    size_t smem = stream_smem_bytes(KP);
    if (smem < 120 * 1024) smem = 120 * 1024;             // one CTA per SM
    VIT_CUDA_TRY(cudaFuncSetAttribute(stream_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int num_sms = 148, dev = 0;
    VIT_CUDA_TRY(cudaGetDevice(&dev));
    VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    const int want = (B + sClipsPerCta - 1) / sClipsPerCta;
    const int grid = want < num_sms ? want : num_sms;
    if (ev0) VIT_CUDA_TRY(cudaEventRecord(ev0, stream));
    stream_forward_kernel<<<grid, sThreads, smem, stream>>>(packed, log_pi, log_emis, lengths, B, T_max, S, KP, n_jb,
                                                            n_kt, hist, t_begin, t_end);
    note_launch();
    VIT_CUDA_TRY(cudaGetLastError());
    if (ev1) VIT_CUDA_TRY(cudaEventRecord(ev1, stream));
  }
  if (do_backtrace) return launch_hist_backtrace(logA_T, hist, lengths, B, T_max, S, paths, scores, stream);
  return VIT_OK;
}

}  // namespace vit
