// VIT_ALGO_BANDED, wide variant: the structured fast path of vit_banded.cu for the 722-state sets with wide bands
// (jdc: +-40 bins of 721, jdc/viterbi_transition_post_processing.py `d_max = 40`; the imm HMM: +-56,
// imm/viterbi_transition_post_processing.py:7-18 with 240 bins per octave).
//
// Same identity, same proof (vit_banded.cu header): with c = the minimum entry of logA^T,
//     max_i fl(delta_i + a_ji) = max( max_{i in band(j)} fl(delta_i + a_ji), fl(delta_unv + a_j,unv), fl(max_i delta_i + c) ),
// bit-identical to the dense recursion at S (2d + 3) instead of S^2 cells per frame (8.7x less at S = 722, d = 40).
//
// What changes is where the band lives: 4 x 81 entries per thread do not fit the register file, so -- as in
// vit_tmem.cu -- they go to TENSOR MEMORY.  Thread (quadrant warp, lane) owns 6 consecutive targets; its TMEM lane holds
// their 6 x (2d+1) band entries as chunks of 4 band offsets x 6 targets = 24 columns, streamed with tcgen05.ld one chunk
// ahead of the math.  To stay under the 64 B/clk TMEM read rate every chunk is applied to FOUR clips (1 byte of TMEM per
// cell).  One CTA per SM, 8 clips, two pipelines of 4 warps; delta rows in shared memory, window loads are LDS.64
// (6 targets = 24 bytes per thread: 8-byte aligned, conflict-free per half warp); dense (unvoiced) row and max_i delta_i
// as per-warp partials combined after the step's single barrier, exactly as in vit_banded.cu.
#include <cstdlib>
#include <type_traits>

#include "vit_tmem.cuh"

namespace vit {

constexpr int wNJ = 6;                     // consecutive targets per thread: 128 TMEM lanes x 6 = 768 targets
constexpr int wCPT = 4;                    // clips per thread
constexpr int wCS = 2;                     // pipelines per CTA
constexpr int wMB = wCPT * wCS;            // 8 clips per CTA
constexpr int wTGW = 4;                    // warps per pipeline = TMEM lane quadrants
constexpr int wPipeThreads = 32 * wTGW;    // 128
constexpr int wThreads = wPipeThreads * wCS;
constexpr int wMaxS = 32 * wTGW * wNJ;     // 768
constexpr int wChunkCols = 4 * wNJ;        // 24 TMEM columns per chunk of 4 band offsets
constexpr int wChunksTmem = tTmemCols / wChunkCols;   // 21 chunks fit the 512 columns (d <= 40: all of them)
// chunks past that (d = 56: 8 of 29) live in shared memory as float4 [chunk][target n][thread], like the K tail of the
// dense tensor-memory kernel
constexpr int wMaxTailChunks = 8;
constexpr size_t wTailFloats = (size_t)wMaxTailChunks * wNJ * wPipeThreads * 4;

__device__ __forceinline__ void wpipe_bar_sync(int cs) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + cs), "n"(wPipeThreads) : "memory");
}
__device__ __forceinline__ float wwarp_max(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// packed [128 lanes][512 columns]: lane tg, column c*24 + n*4 + rr = logA^T[j = 6 tg + n][i = j + (4c + rr) - D], -inf where
// the offset is past the band, the source is outside the matrix, or row / column is the dense state's
template <int D>
__global__ void wide_pack_kernel(const float* __restrict__ logA_T, int S, int jd, float* __restrict__ packed) {
  constexpr int W = 2 * D + 1, NCH = (W + 3) / 4;
  constexpr int NCH_T = NCH < wChunksTmem ? NCH : wChunksTmem;
  const int total = 128 * tTmemCols + (int)wTailFloats;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < total; x += gridDim.x * blockDim.x) {
    int c, n, rr, tg;
    bool ok;
    if (x < 128 * tTmemCols) {
      const int col = x % tTmemCols;
      tg = x / tTmemCols;
      c = col / wChunkCols;
      const int w = col - c * wChunkCols;
      n = w >> 2; rr = w & 3;
      ok = c < NCH_T;
    } else {
      // tail: float index ((ct * wNJ + n) * 128 + tg) * 4 + rr
      const int y = x - 128 * tTmemCols;
      rr = y & 3;
      tg = (y >> 2) % wPipeThreads;
      n = (y >> 2) / wPipeThreads % wNJ;
      c = NCH_T + (y >> 2) / (wPipeThreads * wNJ);
      ok = c < NCH;
    }
    float v = -INFINITY;
    if (ok) {
      const int r = 4 * c + rr, j = wNJ * tg + n, i = j + r - D;
      if (r < W && j < S && j != jd && i >= 0 && i < S && i != jd) v = logA_T[(size_t)j * S + i];
    }
    packed[x] = v;
  }
}

template <int D>
__global__ void __launch_bounds__(wThreads, 1)
wide_forward_kernel(const float* __restrict__ packed, const float* __restrict__ logA_T, const float* __restrict__ log_pi,
                    const float* __restrict__ log_emis, const int32_t* __restrict__ lengths, int B, int T_max, int S,
                    int jd, float cbg, float* __restrict__ hist, float* __restrict__ rowmax, int t_begin, int t_end,
                    int q) {
  // rowmax [B][T_max]: max_i delta_t[i] per frame, for the structured backtrace (vit_cluster.cu)
  // q: clips per CTA and pass (1..8), spread evenly by the host; pipeline cs takes clips [4 cs, 4 cs + 4) of them and
  // runs the 4-, 2- or 1-clip instance of the step (a single recording costs a quarter of the cells of a full pipeline)
  constexpr int W = 2 * D + 1, NCH = (W + 3) / 4;
  constexpr int DP = (D + 1) / 2 * 2;                // delta row: state i at float index i + DP (even: 8-byte aligned windows)
  constexpr int U0 = DP - D;                         // window element of cell (r, n) = w[U0 + r + n]
  constexpr int NWIN = U0 + 4 * NCH + wNJ;           // window floats a thread may touch (rounded up to even below)
  constexpr int ROW = (wMaxS + NWIN + 8) & ~1;       // floats per delta row
  constexpr int NCH_T = NCH < wChunksTmem ? NCH : wChunksTmem;
  constexpr int NTAIL = NCH - NCH_T;                 // chunks served from shared memory
  static_assert(NTAIL <= wMaxTailChunks, "band too wide for the packed tail");
  extern __shared__ __align__(16) float smem[];      // [wCS][2][wCPT][ROW], then the band tail [NTAIL][wNJ][128] float4
  __shared__ float s_partM[wCS][2][wTGW][wCPT];
  __shared__ float s_partD[wCS][2][wTGW][wCPT];
  __shared__ int s_len[wMB];
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = warp / wTGW, Q = warp - cs * wTGW;
  const int tg = Q * 32 + lane;
  const int j0 = wNJ * tg;
  float* sD = smem + (size_t)cs * 2 * wCPT * ROW;

  // ---- one-time: TMEM allocation + fill, dense-column and dense-row entries ---------------------------------------
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&s_tmem_base)), "n"(tTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  float acol[wNJ], arow[wNJ];
  bool jn_ok[wNJ];
#pragma unroll
  for (int n = 0; n < wNJ; ++n) {
    const int j = j0 + n;
    jn_ok[n] = j < S && j != jd;
    acol[n] = (jn_ok[n] && jd >= 0) ? logA_T[(size_t)j * S + jd] : -INFINITY;
    const int i = Q * 32 * wNJ + lane + 32 * n;      // my share of the dense row's sources: 6 per lane, coalesced
    arow[n] = (jd >= 0 && i < S) ? logA_T[(size_t)jd * S + i] : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = __shfl_sync(0xffffffffu, s_tmem_base + ((uint32_t)(Q * 32) << 16), 0);
  float4* sTail = reinterpret_cast<float4*>(smem + (size_t)wCS * 2 * wCPT * ROW);
  if (cs == 0) {
    const float4* src = reinterpret_cast<const float4*>(packed + (size_t)tg * tTmemCols);
    for (int x = 0; x < NCH_T * wChunkCols / 4; ++x) tmem_st4(tbase + 4 * x, src[x]);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  if constexpr (NTAIL > 0) {
    const float4* src = reinterpret_cast<const float4*>(packed + (size_t)128 * tTmemCols);
    for (int x = tid; x < NTAIL * wNJ * wPipeThreads; x += wThreads) sTail[x] = src[x];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const long long hist_delta = reinterpret_cast<const char*>(hist) - reinterpret_cast<const char*>(log_emis);

  const int c0 = cs * wCPT;
  const int ncl = min(wCPT, q - c0);                         // clips of this pipeline (<= 0: idle)
  auto run_pass = [&](auto cpt_tag, int seq0) {
    constexpr int CPT = decltype(cpt_tag)::value;
    int len[CPT];
    int maxlen = 0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) { len[c] = s_len[c0 + c]; maxlen = max(maxlen, len[c]); }
    const float* pe[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) pe[c] = log_emis + ((size_t)(seq0 + c0 + c) * T_max + t_begin) * S + j0;
    const int jd_off = jd - j0;
    const int t_stop = min(maxlen, t_end);
    if (t_begin > 0 && t_begin < t_stop) {
      // resume a frame range from the history (see vit_banded.cu)
      const int pb = (t_begin - 1) & 1;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const float* row = hist + ((size_t)(seq0 + c0 + c) * T_max + (t_begin - 1)) * S;
        const bool lv = t_begin - 1 < len[c];
        float mloc = -INFINITY;
        float* drow = sD + (size_t)(pb * wCPT + c) * ROW + DP + j0;
#pragma unroll
        for (int n = 0; n < wNJ; ++n) {
          const float v = (lv && jn_ok[n]) ? row[j0 + n] : -INFINITY;
          drow[n] = v;
          mloc = fmaxf(mloc, v);
        }
        mloc = wwarp_max(mloc);
        if (lane == 0) {
          s_partM[cs][pb][Q][c] = mloc;
          s_partD[cs][pb][Q][c] = (lv && jd >= 0) ? row[jd] : -INFINITY;
        }
      }
      wpipe_bar_sync(cs);
    }

    float ed_prev[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) ed_prev[c] = 0.f;
    for (int t = t_begin; t < t_stop; ++t) {
      const int buf = t & 1;
#pragma unroll
      for (int c = 0; c < CPT; ++c) asm volatile("" : "+l"(pe[c]));
      bool live[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) live[c] = t < len[c];
      float e[CPT][wNJ], ed[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        if ((lane & 3) == 0 && t + 4 < len[c]) asm volatile("prefetch.global.L2 [%0];" ::"l"(pe[c] + 4 * (size_t)S));
#pragma unroll
        for (int n = 0; n < wNJ; ++n) e[c][n] = (live[c] && jn_ok[n]) ? __ldg(pe[c] + n) : 0.f;
        ed[c] = (lane == 0 && live[c] && jd >= 0) ? __ldg(pe[c] + jd_off) : 0.f;
      }

      float acc[CPT][wNJ], pd[CPT], xd[CPT];
      if (t == 0) {
        // T1[0] = log_pi + logE[0]                                                              (imm/tf_viterbi.py:94)
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
#pragma unroll
          for (int n = 0; n < wNJ; ++n) acc[c][n] = jn_ok[n] ? log_pi[j0 + n] : -INFINITY;
          pd[c] = jd >= 0 ? log_pi[jd] : -INFINITY;
          xd[c] = -INFINITY;
        }
      } else {
        const float* prev = sD + (size_t)((buf ^ 1) * wCPT) * ROW;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const float* pm = &s_partM[cs][buf ^ 1][0][c];
          const float* pdd = &s_partD[cs][buf ^ 1][0][c];
          float dm = fmaxf(fmaxf(pm[0], pm[wCPT]), fmaxf(pm[2 * wCPT], pm[3 * wCPT]));
          xd[c] = (jd >= 0) ? __fadd_rn(fmaxf(fmaxf(pdd[0], pdd[wCPT]), fmaxf(pdd[2 * wCPT], pdd[3 * wCPT])), ed_prev[c]) : -INFINITY;
          dm = fmaxf(dm, xd[c]);
          if (Q == 0 && lane == 1 && t - 1 < len[c]) rowmax[(size_t)(seq0 + c0 + c) * T_max + (t - 1)] = dm;
          if (jd >= 0 && Q == 0 && lane == 0 && t - 1 < len[c] && t - 1 >= t_begin)
            st_global_cs_f32(hist + ((size_t)(seq0 + c0 + c) * T_max + (t - 1)) * S + jd, xd[c]);
          const float bg = __fadd_rn(dm, cbg);       // background term fl(max_i delta_i + c)
#pragma unroll
          for (int n = 0; n < wNJ; ++n) acc[c][n] = fmaxf(bg, __fadd_rn(xd[c], acol[n]));   // + the dense source column
          // my share of the dense target row
          float m = -INFINITY;
#pragma unroll
          for (int n = 0; n < wNJ; ++n) {
            const int i = Q * 32 * wNJ + lane + 32 * n;
            float dv = (i < S) ? prev[(size_t)c * ROW + DP + i] : -INFINITY;
            if (i == jd) dv = xd[c];
            m = fmaxf(m, __fadd_rn(dv, arow[n]));
          }
          pd[c] = m;
        }
        // the band: chunks of 4 offsets x 6 targets from TMEM (one chunk ahead), each applied to my 4 clips; the window
        // of delta_{t-1} slides by 4 per chunk (two LDS.64 per clip per chunk)
        float win[CPT][wNJ + 4 + 2];                  // w[U0 + 4c .. U0 + 4c + 9] of each clip (U0 <= 1)
        const float* wbase[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          wbase[c] = prev + (size_t)c * ROW + j0;     // float index of window element 0 (8-byte aligned)
#pragma unroll
          for (int k = 0; k < (wNJ + 4 + 2) / 2; ++k) {
            const float2 v = reinterpret_cast<const float2*>(wbase[c])[k];
            win[c][2 * k] = v.x; win[c][2 * k + 1] = v.y;
          }
        }
        float a0[24], a1[24];
        tmem_ld_chunk<24>(tbase, a0);
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          float* a = (ch & 1) ? a1 : a0;
          float* an = (ch & 1) ? a0 : a1;
          if (ch < NCH_T) tmem_wait_ld<24>(a);
          if (ch + 1 < NCH) {
            if (ch + 1 < NCH_T) {
              tmem_ld_chunk<24>(tbase + (ch + 1) * 24, an);
            } else {
#pragma unroll
              for (int n = 0; n < wNJ; ++n) {
                const float4 v = sTail[((ch + 1 - NCH_T) * wNJ + n) * wPipeThreads + tg];
                an[n * 4 + 0] = v.x; an[n * 4 + 1] = v.y; an[n * 4 + 2] = v.z; an[n * 4 + 3] = v.w;
              }
            }
          }
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
#pragma unroll
              for (int n = 0; n < wNJ; ++n)
                acc[c][n] = fmaxf(acc[c][n], __fadd_rn(win[c][U0 + rr + n], a[n * 4 + rr]));
            if (ch + 1 < NCH) {
              // slide: drop 4, fetch 4 (elements 4(ch+1) + 8 .. + 11 of the window)
#pragma unroll
              for (int k = 0; k < wNJ + 2; ++k) win[c][k] = win[c][k + 4];
              const float2 v0 = reinterpret_cast<const float2*>(wbase[c])[(4 * (ch + 1) + wNJ + 2) / 2];
              const float2 v1 = reinterpret_cast<const float2*>(wbase[c])[(4 * (ch + 1) + wNJ + 2) / 2 + 1];
              win[c][wNJ + 2] = v0.x; win[c][wNJ + 3] = v0.y; win[c][wNJ + 4] = v1.x; win[c][wNJ + 5] = v1.y;
            }
          }
        }
      }

      // T1[t][j] = max + logE[t][j]                                                              (:100)
      float pm[wCPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        float v[wNJ];
        float mloc = -INFINITY;
#pragma unroll
        for (int n = 0; n < wNJ; ++n) {
          v[n] = jn_ok[n] ? __fadd_rn(acc[c][n], e[c][n]) : -INFINITY;
          if (live[c] && jn_ok[n])
            st_global_cs_f32(reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pe[c])) + hist_delta) + n, v[n]);
          mloc = fmaxf(mloc, v[n]);
        }
        float2* drow = reinterpret_cast<float2*>(sD + (size_t)(buf * wCPT + c) * ROW + DP + j0);
        drow[0] = make_float2(v[0], v[1]);
        drow[1] = make_float2(v[2], v[3]);
        drow[2] = make_float2(v[4], v[5]);
        pm[c] = wwarp_max(mloc);
        pd[c] = wwarp_max(pd[c]);
        ed[c] = __shfl_sync(0xffffffffu, ed[c], 0);
      }
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          s_partM[cs][buf][Q][c] = pm[c];
          s_partD[cs][buf][Q][c] = pd[c];
        }
      }
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        ed_prev[c] = ed[c];
        pe[c] += S;
      }
      wpipe_bar_sync(cs);
    }
    // the dense state's value of the last frame of this range
    if (jd >= 0 && t_stop > t_begin && Q == 0 && lane == 0) {
      const int buf = (t_stop - 1) & 1;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        if (t_stop - 1 < len[c]) {
          const float* pdd = &s_partD[cs][buf][0][c];
          const float x = __fadd_rn(fmaxf(fmaxf(pdd[0], pdd[wCPT]), fmaxf(pdd[2 * wCPT], pdd[3 * wCPT])), ed_prev[c]);
          st_global_cs_f32(hist + ((size_t)(seq0 + c0 + c) * T_max + (t_stop - 1)) * S + jd, x);
        }
      }
    }
  };

  for (int seq0 = blockIdx.x * q; seq0 < B; seq0 += gridDim.x * q) {
    __syncthreads();
    if (tid < wMB) {
      const int b = seq0 + tid;
      s_len[tid] = (tid < q && b < B) ? (lengths ? lengths[b] : T_max) : 0;
    }
    for (int x = tid; x < wCS * 2 * wCPT * ROW; x += wThreads) smem[x] = -INFINITY;     // pads stay -inf
    __syncthreads();
    if (ncl >= 3) run_pass(std::integral_constant<int, 4>{}, seq0);
    else if (ncl == 2) run_pass(std::integral_constant<int, 2>{}, seq0);
    else if (ncl == 1) run_pass(std::integral_constant<int, 1>{}, seq0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "n"(tTmemCols) : "memory");
}

static int wide_template_D(int d) {
  const int opts[] = {20, 28, 40, 56};
  for (int o : opts) if (d <= o) return o;
  return -1;
}

bool banded_wide_supported(int S, const vit_structure* st) {
  if (!st || st->halfwidth > 56 || st->halfwidth < 0) return false;
  if (S > wMaxS || S < 2 || (S & 1)) return false;          // even S: every clip row starts 8-byte aligned
  if (st->dense_index < -1 || st->dense_index >= S) return false;
  return wide_template_D(st->halfwidth) > 0;
}

size_t banded_wide_workspace_bytes(int B, int T_max, int S) {
  return align_up(((size_t)128 * tTmemCols + wTailFloats) * sizeof(float), 256) +
         align_up((size_t)B * T_max * S * sizeof(float), 256);
}

template <int D>
static int launch_wide(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths, int B,
                       int T_max, int S, int jd, float cbg, float* packed, float* hist, float* rowmax, int t_begin,
                       int t_end, int grid, int q, cudaStream_t stream) {
  constexpr int W = 2 * D + 1, NCH = (W + 3) / 4;
  constexpr int DP = (D + 1) / 2 * 2, U0 = DP - D, NWIN = U0 + 4 * NCH + wNJ, ROW = (wMaxS + NWIN + 8) & ~1;
  constexpr int NTAIL = NCH > wChunksTmem ? NCH - wChunksTmem : 0;
  size_t smem = (size_t)wCS * 2 * wCPT * ROW * sizeof(float) + (size_t)NTAIL * wNJ * wPipeThreads * 16;
  if (smem < 120 * 1024) smem = 120 * 1024;                  // one CTA per SM: each allocates all 512 TMEM columns
  wide_pack_kernel<D><<<64, 256, 0, stream>>>(logA_T, S, jd, packed);
  note_launch();
  VIT_CUDA_TRY(cudaFuncSetAttribute(wide_forward_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  wide_forward_kernel<D><<<grid, wThreads, smem, stream>>>(packed, logA_T, log_pi, log_emis, lengths, B, T_max, S, jd, cbg,
                                                          hist, rowmax, t_begin, t_end, q);
  note_launch();
  VIT_CUDA_TRY(cudaGetLastError());
  return VIT_OK;
}

int banded_wide_forward(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths, int B,
                        int T_max, int S, const vit_structure* st, void* packed_ws, float* hist, float* rowmax, int t_begin,
                        int t_end, cudaStream_t stream) {
  int num_sms = 148, dev = 0;
  VIT_CUDA_TRY(cudaGetDevice(&dev));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  // spread the clips evenly over the SMs (as vit_banded.cu): `passes` trips of q <= 8 clips per CTA
  const int passes = (B + wMB * num_sms - 1) / (wMB * num_sms);
  int q = (B + passes * num_sms - 1) / (passes * num_sms);
  q = q < 1 ? 1 : (q > wMB ? wMB : q);
  const int want = (B + q - 1) / q;
  const int grid = want < num_sms ? want : num_sms;
  float* packed = (float*)packed_ws;
  switch (wide_template_D(st->halfwidth)) {
    case 20: return launch_wide<20>(logA_T, log_pi, log_emis, lengths, B, T_max, S, st->dense_index, st->background, packed, hist, rowmax, t_begin, t_end, grid, q, stream);
    case 28: return launch_wide<28>(logA_T, log_pi, log_emis, lengths, B, T_max, S, st->dense_index, st->background, packed, hist, rowmax, t_begin, t_end, grid, q, stream);
    case 40: return launch_wide<40>(logA_T, log_pi, log_emis, lengths, B, T_max, S, st->dense_index, st->background, packed, hist, rowmax, t_begin, t_end, grid, q, stream);
    case 56: return launch_wide<56>(logA_T, log_pi, log_emis, lengths, B, T_max, S, st->dense_index, st->background, packed, hist, rowmax, t_begin, t_end, grid, q, stream);
    default: return VIT_ERR_UNSUPPORTED_ALGO;
  }
}

}  // namespace vit
