// VIT_ALGO_BANDED -- bit-exact fast path for STRUCTURED transition matrices.
//
// Every HMM the reference builds with viterbi_transition_matrix.py / viterbi_transition_post_processing.py is a Toeplitz
// band (+-d_max pitch bins: 12 dcnet, 14 tonet/ftanet) embedded in a voiced/unvoiced switch
// (dcnet/viterbi_transition_matrix.py:81-98): outside the band the probabilities are exactly 0, so in the log domain
// every such entry is the SAME constant c = log(0 + tiny) = -87.33655 (dcnet/softmax_viterbi.py:2459-2462); only the
// unvoiced state is a dense source column and a dense target row.  With c = the minimum entry of logA^T,
//
//     max_i fl(delta_i + a_ji)  =  max(  max_{i in N(j)} fl(delta_i + a_ji),   fl(max_i delta_i + c)  ),
//
// N(j) = the entries of row j that differ from c.  Proof: fl32 addition of a constant is monotone, so
// max_i fl(delta_i + c) = fl(max_i delta_i + c) covers every background entry, and for i in N(j), a_ji >= c gives
// fl(delta_i + c) <= fl(delta_i + a_ji), so the extra terms never exceed the true maximum.  The recursion value -- hence
// the delta history, hence the lazily resolved argmax of the backtrace (which still reads the DENSE matrix) -- is
// bit-identical to the dense kernels', at S (2d + 2) instead of S^2 cells per frame: 11.6x less work at S = 361.
//
// The structure is detected once on the host (vit_analyze_structure_f32) and passed in; a matrix without it (the dense
// imm matrix, imm/transition_matrix.py:4-31) simply takes the dense tensor-memory kernel.
//
// Kernel: no clusters, no exchange -- one CTA per SM owns 8 clips for all T steps.  4 independent pipelines of 3 warps
// (2 clips each, named barriers; measured faster than 2 pipelines x 4 clips: the step is a dependency chain, and more
// pipelines hide more of it).  A thread owns 4 consecutive targets and keeps their 4 x (2d+1) band entries in
// REGISTERS for the whole kernel; its window of delta_{t-1} (4 + 2d values per clip) is 8-9 aligned LDS.128 at
// compile-time offsets, its 4 results one STS.128.  The dense target row
// (unvoiced) is split over the 3 warps of a pipeline; its value and max_i delta_i travel as per-warp partials in shared
// memory and are combined by every thread after the step's single barrier.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "vit_tmem.cuh"

namespace vit {

constexpr int bNJ = 4;                     // consecutive targets per thread
constexpr int bCPT = 2;                    // clips per thread
constexpr int bCS = 4;                     // pipelines (clip groups) per CTA
constexpr int bMB = bCPT * bCS;            // 8 clips per CTA
constexpr int bTGW = 3;                    // warps per pipeline: 96 target groups x 4 = 384 targets
constexpr int bPipeThreads = 32 * bTGW;    // 96
constexpr int bThreads = bPipeThreads * bCS;
constexpr int bMaxS = 32 * bTGW * bNJ;     // 384
// a delta row in shared memory: state i lives at float index i + DP, DP = the band half-width rounded up to a multiple
// of 4, so that the 4 targets of a thread are one aligned float4 and its window is a run of aligned float4s.
// 4 * 95 (last target group) + 36 (widest window) = 416 floats cover every thread's accesses for any S <= 384.
constexpr int bRowLen = 416;

__device__ __forceinline__ void bpipe_bar_sync(int cs) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + cs), "n"(bPipeThreads) : "memory");
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}

// one-instruction warp maximum (sm_100a: CREDUX.MAX.F32; NaNs are ignored, +0 / -0 compare equal downstream)
__device__ __forceinline__ float warp_max_redux(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

template <int D>
__global__ void __launch_bounds__(bThreads, 1)
banded_forward_kernel(const float* __restrict__ logA_T, const float* __restrict__ log_pi,
                      const float* __restrict__ log_emis, const int32_t* __restrict__ lengths, int B, int T_max, int S,
                      int jd, float cbg, float* __restrict__ hist, float* __restrict__ rowmax, int t_begin, int t_end,
                      int q) {
  // rowmax [B][T_max]: max_i delta_t[i] per frame, for the structured backtrace (vit_cluster.cu)
  // q: clips per CTA and pass (1..8).  The host spreads a batch evenly (1024 clips = 147 CTAs x 7, not 128 x 8): pipeline
  // cs takes clips [2 cs, 2 cs + 2) of the CTA's q, so the last busy pipeline may hold ONE clip and then runs the
  // one-clip instance of the step (half the cells), and a pipeline past q sits the pass out.
  constexpr int W = 2 * D + 1;
  constexpr int DP = (D + 3) / 4 * 4;
  constexpr int U0 = DP - D;                         // window element of cell (r, n) = w[U0 + r + n]
  constexpr int NW4 = (U0 + W - 1 + bNJ + 3) / 4;    // float4s in a thread's window
  // 4 x 29 band entries + everything else do not fit 168 registers: ptxas spills 12 words and the reloads cost 28 % of
  // the step (measured: the D = 14 instance on a +-12 matrix, 4.80 vs 3.75 ms).  So for D > 12 the last SPL offsets of
  // every target live in TENSOR MEMORY (16 columns of this thread's own lane; idle silicon in a SIMT kernel) and come
  // back once per step with one tcgen05.ld, issued at the top of the step and awaited where the sweep first needs them.
  constexpr int SPL = D > 12 ? 12 : (D == 12 ? 4 : 0);
  constexpr int WR = W - SPL;                        // band offsets kept in registers
  constexpr int M_WAIT = (WR + U0 - 3 + 3) / 4;      // first window float4 with a cell at offset r >= WR
  constexpr int NTC = 4 * SPL;                       // TMEM columns per thread (16 or 32)
  constexpr int kTmemColsBanded = 256;               // 3 warp groups x NTC columns, rounded up to a power of two
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_delta[bCS][2][bCPT][bRowLen];
  __shared__ float s_partM[bCS][2][bTGW][bCPT];
  __shared__ float s_partD[bCS][2][bTGW][bCPT];
  __shared__ int s_len[bMB];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = warp / bTGW, tgw = warp - cs * bTGW;
  const int tg = tgw * 32 + lane;
  const int j0 = bNJ * tg;

  // ---- one-time: my band entries -> registers; dense-column and dense-row entries --------------------------------
  // a[n][r] = logA^T[j0+n][j0+n + r - D]; entries that fall outside the matrix, on the dense state's row or on the
  // dense state's column are -inf (the dense column has its own term, the dense row its own code)
  float a[bNJ][WR], acol[bNJ], arow[bNJ];
  bool jn_ok[bNJ];
  uint32_t taddr = 0;
  if constexpr (SPL > 0) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(&s_tmem_base)), "n"(kTmemColsBanded) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // my lane of quadrant warp & 3, 16 columns of warp group warp >> 2 (uniform per warp)
    taddr = __shfl_sync(0xffffffffu, s_tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * NTC), 0);
  }
#pragma unroll
  for (int n = 0; n < bNJ; ++n) {
    const int j = j0 + n;
    const bool jok = j < S && j != jd;
    jn_ok[n] = jok;
#pragma unroll
    for (int r = 0; r < WR; ++r) {
      const int i = j + r - D;
      a[n][r] = (jok && i >= 0 && i < S && i != jd) ? logA_T[(size_t)j * S + i] : -INFINITY;
    }
    if constexpr (SPL > 0) {
#pragma unroll
      for (int g4 = 0; g4 < SPL / 4; ++g4) {
        float hi[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = j + WR + 4 * g4 + k - D;
          hi[k] = (jok && i >= 0 && i < S && i != jd) ? logA_T[(size_t)j * S + i] : -INFINITY;
        }
        tmem_st4(taddr + g4 * 16 + n * 4, make_float4(hi[0], hi[1], hi[2], hi[3]));
      }
    }
    acol[n] = (jok && jd >= 0) ? logA_T[(size_t)j * S + jd] : -INFINITY;   // dense source column: A[jd -> j]
    arow[n] = (jok && jd >= 0) ? logA_T[(size_t)jd * S + j] : -INFINITY;   // dense target row:    A[j -> jd]
  }
  if constexpr (SPL > 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  const float a_dd = jd >= 0 ? logA_T[(size_t)jd * S + jd] : -INFINITY;
  // a warp whose 128 targets are all band states takes the unpredicated loads / stores
  const bool wfull = __all_sync(0xffffffffu, jn_ok[0] && jn_ok[1] && jn_ok[2] && jn_ok[3]);
  const long long hist_delta = reinterpret_cast<const char*>(hist) - reinterpret_cast<const char*>(log_emis);
  const int jd_off = jd - j0;                                // logE[..][jd] relative to my emission pointer

  const int c0 = cs * bCPT;
  const int ncl = min(bCPT, q - c0);                         // clips of this pipeline (<= 0: idle)
  // the body of one pass, instantiated for 2 and for 1 clips per thread
  auto run_pass = [&](auto cpt_tag, int seq0) {
    constexpr int CPT = decltype(cpt_tag)::value;
    int len[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) len[c] = s_len[c0 + c];
    int maxlen = 0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) maxlen = max(maxlen, len[c]);
    // address of logE[clip][t][j0], advanced one frame per step; the history sits at a fixed distance.  A clip keeps
    // stepping (on whatever its rows hold) until the longest clip of its pipeline is done -- nothing of it is stored
    // past its length -- and a slot past the batch aliases the last clip's rows, so no load needs a per-step predicate.
    const float* pe[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) pe[c] = log_emis + ((size_t)min(seq0 + c0 + c, B - 1) * T_max + t_begin) * S + j0;
    const int t_stop = min(maxlen, t_end);
    // rowmax index of frame t - 1 is rmi[c] + t (the host keeps B * T_max below 2^31 on this path)
    int rmi[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) rmi[c] = (seq0 + c0 + c) * T_max - 1;
    float xd[CPT];                                          // delta_{t-1}[jd] of my clips
#pragma unroll
    for (int c = 0; c < CPT; ++c) xd[c] = -INFINITY;
    if (t_begin > 0 && t_begin < t_stop) {
      // resume a frame range: delta_{t_begin-1} comes back from the history -- my 4 targets into the delta row, the
      // per-warp partials the first step will combine, and the dense state's own value
      const int pb = (t_begin - 1) & 1;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const float* row = reinterpret_cast<const float*>(reinterpret_cast<const char*>(pe[c]) + hist_delta) - S - j0;
        float v[bNJ], mloc = -INFINITY, dloc = -INFINITY;
#pragma unroll
        for (int n = 0; n < bNJ; ++n) {
          v[n] = jn_ok[n] ? row[j0 + n] : -INFINITY;
          mloc = fmaxf(mloc, v[n]);
          dloc = fmaxf(dloc, __fadd_rn(v[n], arow[n]));
        }
        reinterpret_cast<float4*>(s_delta[cs][pb][c] + DP)[tg] = make_float4(v[0], v[1], v[2], v[3]);
        mloc = warp_max_redux(mloc);
        dloc = warp_max_redux(dloc);
        if (lane == 0) {
          s_partM[cs][pb][tgw][c] = mloc;
          s_partD[cs][pb][tgw][c] = dloc;
        }
        xd[c] = jd >= 0 ? row[jd] : -INFINITY;
      }
      bpipe_bar_sync(cs);
    }

    for (int t = t_begin; t < t_stop; ++t) {
      const int buf = t & 1;
      // keep the two running pointers in registers (ptxas would otherwise re-derive the 64-bit address of every load)
#pragma unroll
      for (int c = 0; c < CPT; ++c) asm volatile("" : "+l"(pe[c]));
      // this step's emissions: issued first, used last
      float e[CPT][bNJ], ed[CPT];
      const bool pf = (lane & 7) == 0 && t + 4 < T_max;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        // a step is shorter than an HBM round trip: pull the lines of frame t + 4 into L2 now (one lane per 128 B)
        if (pf) asm volatile("prefetch.global.L2 [%0];" ::"l"(pe[c] + 4 * (size_t)S));
        if (wfull) {
#pragma unroll
          for (int n = 0; n < bNJ; ++n) e[c][n] = __ldg(pe[c] + n);
        } else {
#pragma unroll
          for (int n = 0; n < bNJ; ++n) e[c][n] = jn_ok[n] ? __ldg(pe[c] + n) : 0.f;
        }
        ed[c] = jd >= 0 ? __ldg(pe[c] + jd_off) : 0.f;      // one address per warp
      }

      float acc[CPT][bNJ], xdn[CPT];
      if (t == 0) {
        // T1[0] = log_pi + logE[0]                                                              (imm/tf_viterbi.py:94)
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
#pragma unroll
          for (int n = 0; n < bNJ; ++n) acc[c][n] = jn_ok[n] ? log_pi[j0 + n] : -INFINITY;
          xdn[c] = jd >= 0 ? __fadd_rn(log_pi[jd], ed[c]) : -INFINITY;
        }
      } else {
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const float* prev = s_delta[cs][buf ^ 1][c];
#pragma unroll
          for (int n = 0; n < bNJ; ++n) acc[c][n] = -INFINITY;
          // the band: my window of delta_{t-1} is NW4 aligned float4s starting at my own targets' slot.  One float4 at
          // a time, applied to every (offset r, target n) cell that reads it: only 4 window values are ever live
          const float4* row4 = reinterpret_cast<const float4*>(prev) + tg;
          float ah[NTC > 0 ? NTC : 1];                       // band offsets WR.. of my 4 targets
#pragma unroll
          for (int m = 0; m < NW4; ++m) {
            const float4 v = row4[m];
            const float wv[4] = {v.x, v.y, v.z, v.w};
            if constexpr (SPL > 0) {
              // (short-lived on purpose: fetched per clip two float4s ahead of the first cell that needs them)
              // (each group of 4 offsets is fetched one float4 ahead of the first cell that needs it)
#pragma unroll
              for (int g4 = 0; g4 < SPL / 4; ++g4) {
                if (m == M_WAIT + g4 - 1) tmem_ld<16>(taddr + 16 * g4, ah + 16 * g4);
                if (m == M_WAIT + g4) tmem_wait_ld<16>(ah + 16 * g4);
              }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int n = 0; n < bNJ; ++n) {
                const int r = 4 * m + k - U0 - n;                    // cell (r, n) reads window element U0 + r + n
                if (r >= 0 && r < WR) acc[c][n] = fmaxf(acc[c][n], __fadd_rn(wv[k], a[n][r]));
                else if (r >= WR && r < W) acc[c][n] = fmaxf(acc[c][n], __fadd_rn(wv[k], ah[((r - WR) >> 2) * 16 + n * 4 + ((r - WR) & 3)]));
              }
          }
          // combine the per-warp partials of step t-1: max_{i != jd} delta_{t-1}[i] and the dense target row's
          // max_{i != jd} fl(delta_{t-1}[i] + A[i -> jd]); then delta_t[jd], the background term fl(max_i delta_i + c)
          // and the dense source column
          const float* pm = &s_partM[cs][buf ^ 1][0][c];
          const float* pdd = &s_partD[cs][buf ^ 1][0][c];
          const float dm = fmaxf(fmaxf(fmaxf(pm[0], pm[bCPT]), pm[2 * bCPT]), xd[c]);
          if (tgw == 0 && lane == 2 && t - 1 < len[c]) rowmax[rmi[c] + t] = dm;
          const float dd = fmaxf(fmaxf(fmaxf(pdd[0], pdd[bCPT]), pdd[2 * bCPT]), __fadd_rn(xd[c], a_dd));
          xdn[c] = jd >= 0 ? __fadd_rn(dd, ed[c]) : -INFINITY;
          const float bg = __fadd_rn(dm, cbg);
#pragma unroll
          for (int n = 0; n < bNJ; ++n) acc[c][n] = fmaxf(acc[c][n], fmaxf(bg, __fadd_rn(xd[c], acol[n])));
        }
      }

      // T1[t][j] = max + logE[t][j]                                                              (:100)
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        float v[bNJ];
        float mloc = -INFINITY, dloc = -INFINITY;
        float* ph = reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pe[c])) + hist_delta);
        const bool st = t < len[c];
        if (wfull) {
#pragma unroll
          for (int n = 0; n < bNJ; ++n) {
            v[n] = __fadd_rn(acc[c][n], e[c][n]);
            if (st) st_global_cs_f32(ph + n, v[n]);
          }
        } else {
#pragma unroll
          for (int n = 0; n < bNJ; ++n) {
            v[n] = jn_ok[n] ? __fadd_rn(acc[c][n], e[c][n]) : -INFINITY;
            if (st && jn_ok[n]) st_global_cs_f32(ph + n, v[n]);
          }
        }
#pragma unroll
        for (int n = 0; n < bNJ; ++n) {
          mloc = fmaxf(mloc, v[n]);
          dloc = fmaxf(dloc, __fadd_rn(v[n], arow[n]));
        }
        // one aligned float4 (slots of non-existent / dense targets get -inf, which is what they must hold)
        reinterpret_cast<float4*>(s_delta[cs][buf][c] + DP)[tg] = make_float4(v[0], v[1], v[2], v[3]);
        mloc = warp_max_redux(mloc);
        dloc = warp_max_redux(dloc);
        if (lane == 0) {
          s_partM[cs][buf][tgw][c] = mloc;
          s_partD[cs][buf][tgw][c] = dloc;
        }
        // the dense state's own value of this frame (every thread has it; one stores it)
        if (jd >= 0 && tgw == 0 && lane == 1 && st) st_global_cs_f32(ph + jd_off, xdn[c]);
        xd[c] = xdn[c];
        pe[c] += S;
      }
      bpipe_bar_sync(cs);
    }
  };

  for (int seq0 = blockIdx.x * q; seq0 < B; seq0 += gridDim.x * q) {
    __syncthreads();
    if (tid < bMB) {
      const int b = seq0 + tid;
      s_len[tid] = (tid < q && b < B) ? (lengths ? lengths[b] : T_max) : 0;
    }
    // (re-)arm the delta rows: pads, out-of-range states and the dense state's slot stay -inf for the whole sub-batch
    for (int x = tid; x < bCS * 2 * bCPT * bRowLen; x += bThreads) (&s_delta[0][0][0][0])[x] = -INFINITY;
    __syncthreads();
    if (ncl >= 2) run_pass(std::integral_constant<int, 2>{}, seq0);
    else if (ncl == 1) run_pass(std::integral_constant<int, 1>{}, seq0);
  }
  if constexpr (SPL > 0) {
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "n"(kTmemColsBanded) : "memory");
  }
}

// vit_cluster.cu
int launch_hist_backtrace(const float* logA_T, const float* hist, const int32_t* lengths, int B, int T_max, int S,
                          int64_t* paths, float* scores, cudaStream_t stream);

static int banded_template_D(int d) {
  const int opts[] = {4, 8, 12, 14};
  if (const char* f = getenv("VIT_BANDED_FORCE_D")) {         // experiment knob: a wider instance than the band needs
    const int fd = atoi(f);
    for (int o : opts) if (o == fd && d <= o) return o;
  }
  for (int o : opts) if (d <= o) return o;
  return -1;
}

// vit_banded_wide.cu: band entries in tensor memory (S <= 768, d <= 56)
bool banded_wide_supported(int S, const vit_structure* st);
int banded_wide_forward(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths, int B,
                        int T_max, int S, const vit_structure* st, void* packed_ws, float* hist, float* rowmax, int t_begin,
                        int t_end, cudaStream_t stream);
// vit_cluster.cu: walk that reads only the band window of the state on the path (falls back to the full row when a
// background source could win)
int launch_structured_backtrace(const float* logA_T, const float* hist, const float* rowmax, const int32_t* lengths,
                                int B, int T_max, int S, const vit_structure* st, int64_t* paths, float* scores,
                                cudaStream_t stream);
constexpr size_t kWidePackedBytes = (128 * 512 + 8 * 6 * 128 * 4) * sizeof(float);   // TMEM image + shared-memory band tail

static bool banded_narrow_supported(int S, const vit_structure* st) {
  if (!st || st->kind != 1) return false;
  if (S > bMaxS || S < 2) return false;
  if (st->dense_index < -1 || st->dense_index >= S) return false;
  return banded_template_D(st->halfwidth) > 0;
}

bool banded_supported(int S, const vit_structure* st) {
  if (!st || st->kind != 1) return false;
  return banded_narrow_supported(S, st) || banded_wide_supported(S, st);
}

// both banded kernels: one CTA per SM, 8 clips per CTA
int banded_clips_in_flight(int* out) {
  static_assert(bMB == 8, "vit_banded_wide.cu keeps 8 clips per CTA as well");
  int num_sms = 148, dev = 0;
  VIT_CUDA_TRY(cudaGetDevice(&dev));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  *out = num_sms * bMB;
  return VIT_OK;
}

static size_t banded_rowmax_bytes(int B, int T_max) { return align_up((size_t)B * T_max * sizeof(float), 256); }

size_t banded_workspace_bytes(int B, int T_max, int S) {
  // TMEM image of the band, row maxima, T1 table
  return align_up(kWidePackedBytes, 256) + banded_rowmax_bytes(B, T_max) + align_up((size_t)B * T_max * S * sizeof(float), 256);
}

int banded_decode(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths, int B,
                  int T_max, int S, const vit_structure* st, void* workspace, size_t workspace_bytes, int64_t* paths,
                  float* scores, float* delta_out, int t_begin, int t_end, bool do_backtrace, cudaEvent_t ev0,
                  cudaEvent_t ev1, cudaStream_t stream) {
  if (!banded_supported(S, st)) return VIT_ERR_UNSUPPORTED_ALGO;
  if ((long long)B * T_max >= 0x7fffffffLL) return VIT_ERR_INVALID_ARGUMENT;   // (frame indices are int32 in the kernels)
  if (t_end > T_max) t_end = T_max;
  if (t_begin < 0 || t_begin > t_end) return VIT_ERR_INVALID_ARGUMENT;
  const size_t head_bytes = align_up(kWidePackedBytes, 256) + banded_rowmax_bytes(B, T_max);
  if (workspace_bytes < (delta_out ? head_bytes : banded_workspace_bytes(B, T_max, S))) return VIT_ERR_WORKSPACE_TOO_SMALL;
  if (B == 0) return VIT_OK;
  float* rowmax = (float*)((char*)workspace + align_up(kWidePackedBytes, 256));
  float* hist = delta_out ? delta_out : (float*)((char*)workspace + head_bytes);
  if (!banded_narrow_supported(S, st)) {
    // wide bands / 722-state sets: band entries in tensor memory
    if (ev0) VIT_CUDA_TRY(cudaEventRecord(ev0, stream));
    const int rc = banded_wide_forward(logA_T, log_pi, log_emis, lengths, B, T_max, S, st, workspace, hist, rowmax, t_begin,
                                       t_end, stream);
    if (rc != VIT_OK) return rc;
    if (ev1) VIT_CUDA_TRY(cudaEventRecord(ev1, stream));
    if (!do_backtrace) return VIT_OK;
    return launch_structured_backtrace(logA_T, hist, rowmax, lengths, B, T_max, S, st, paths, scores, stream);
  }
  const int D = banded_template_D(st->halfwidth);
  int num_sms = 148, dev = 0;
  VIT_CUDA_TRY(cudaGetDevice(&dev));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  // spread the clips evenly over the SMs: `passes` trips of q <= 8 clips per CTA (1024 clips: 147 CTAs x 7)
  const int passes = (B + bMB * num_sms - 1) / (bMB * num_sms);
  int q = (B + passes * num_sms - 1) / (passes * num_sms);
  q = q < 1 ? 1 : (q > bMB ? bMB : q);
  const int want = (B + q - 1) / q;
  const int grid = want < num_sms ? want : num_sms;
  if (ev0) VIT_CUDA_TRY(cudaEventRecord(ev0, stream));
#define VIT_BANDED_CASE(DD)                                                                                          \
  case DD: {                                                                                                         \
    banded_forward_kernel<DD><<<grid, bThreads, 0, stream>>>(logA_T, log_pi, log_emis, lengths, B, T_max, S,         \
                                                              st->dense_index, st->background, hist, rowmax,        \
                                                              t_begin, t_end, q);                                    \
  } break;
  switch (D) {
    VIT_BANDED_CASE(4) VIT_BANDED_CASE(8) VIT_BANDED_CASE(12) VIT_BANDED_CASE(14)
    default: return VIT_ERR_UNSUPPORTED_ALGO;
  }
#undef VIT_BANDED_CASE
  note_launch();
  VIT_CUDA_TRY(cudaGetLastError());
  if (ev1) VIT_CUDA_TRY(cudaEventRecord(ev1, stream));
  if (!do_backtrace) return VIT_OK;
  return launch_structured_backtrace(logA_T, hist, rowmax, lengths, B, T_max, S, st, paths, scores, stream);
}

// Host-side structure analysis (h_logA_T is a HOST pointer).  kind = 1 iff, apart from at most one state that is both
// a dense source column and a dense target row, every entry that differs from the minimum entry c lies within
// |i - j| <= halfwidth and one of the two banded kernels takes the shape (S <= 384 with d <= 14: band in registers;
// even S <= 768 with d <= 56: band in tensor memory, chunks past 512 columns in shared memory).
int analyze_structure(const float* A, int S, vit_structure* out) {
  out->kind = 0;
  out->halfwidth = 0;
  out->dense_index = -1;
  out->background = 0.f;
  out->dense_row_max = INFINITY;
  if (S < 2) return VIT_OK;
  float c = A[0];
  for (size_t x = 0; x < (size_t)S * S; ++x) {
    if (A[x] != A[x]) return VIT_OK;                       // NaN: no structure claimed
    c = std::min(c, A[x]);
  }
  std::vector<int> cnt_row(S, 0), cnt_col(S, 0);
  for (int j = 0; j < S; ++j)
    for (int i = 0; i < S; ++i)
      if (A[(size_t)j * S + i] != c) { ++cnt_row[j]; ++cnt_col[i]; }
  // candidate dense state: the one with the fullest row + column
  int dense = -1, best = 0;
  for (int s = 0; s < S; ++s) {
    const int v = cnt_row[s] + cnt_col[s];
    if (v > best) { best = v; dense = s; }
  }
  int d_with = 0, d_without = 0;
  for (int j = 0; j < S; ++j)
    for (int i = 0; i < S; ++i)
      if (A[(size_t)j * S + i] != c) {
        const int d = std::abs(i - j);
        d_without = std::max(d_without, d);
        if (i != dense && j != dense) d_with = std::max(d_with, d);
      }
  int d = d_without, di = -1;
  if (d_with < d_without) { d = d_with; di = dense; }
  out->background = c;
  out->halfwidth = d;
  out->dense_index = di;
  out->dense_row_max = INFINITY;
  if (di >= 0) {
    float m = -INFINITY;
    for (int i = 0; i < S; ++i)
      if (i != di) m = std::max(m, A[(size_t)di * S + i]);
    out->dense_row_max = m;
  }
  out->kind = 1;
  out->kind = (banded_narrow_supported(S, out) || banded_wide_supported(S, out)) ? 1 : 0;
  return VIT_OK;
}

}  // namespace vit
