// Forward-backward for STRUCTURED transition matrices -- the sum-product counterpart of vit_banded.cu.
//
// Every HMM the reference builds with viterbi_transition_matrix.py / viterbi_transition_post_processing.py
// (dcnet/viterbi_transition_matrix.py:81-98) is a band of +-d_max pitch bins embedded in a voiced/unvoiced switch: in
// the PROBABILITY domain -- which is what this pass works in (oracle/fb_oracle.py) -- every entry outside the band is
// exactly 0, and only the unvoiced state is a dense source column and a dense target row.  A zero entry contributes
// nothing to a sum, so
//
//     (alpha A)[j]  =  sum_{|i-j| <= d} alpha[i] A[i][j]  +  alpha[u] A[u][j]              (j a pitch state)
//     (alpha A)[u]  =  sum_i alpha[i] A[i][u]                                               (u the unvoiced state)
//
// is the SAME sum as the dense product with the zero terms left out: S (2d + 3) instead of S^2 multiply-adds per frame
// (11.6x less at S = 361, d = 14).  The dense kernels need the tensor cores to get through S^2 (vit_fb_tc.cu: 13.1 ms at
// 1024 x 3000 x 361, bound by the issue rate of small-N MMAs); with the structure the pass is plain FFMA work whose floor is
// the HBM traffic of the recursion itself (read b_t, write alpha~_t; read b_t, alpha~_t, write gamma_t: 20 S bytes per
// frame).  fp32 throughout, no bf16 splitting: |gamma - float64 oracle| ~ 4e-6.
//
// There is NO reference implementation of forward-backward (correction 2): parity unpinned, the
// float64 oracle of this repository is the checker (tests/test_gpu_fb.py), tolerance 1e-4 absolute on gamma and 1e-5
// relative on log L (the north star's).
//
// Kernel = the layout of banded_forward_kernel (vit_banded.cu) with the semiring swapped: one CTA per SM, 4 pipelines of
// 3 warps that own 2 clips each for all T steps (named barriers, no cluster, no exchange); a thread owns 4 consecutive
// states and keeps their 4 x (2d+1) band entries in REGISTERS for the whole kernel (d = 14: the last 12 offsets in
// tensor memory, fetched with short-lived tcgen05.ld); its window of the previous vector is 8-9 aligned LDS.128.
//   forward  : M = A^T.  a~_t = (M a~_{t-1} / c_{t-1}) * b_t is stored UNNORMALISED in the gamma buffer; the normaliser
//              c_{t-1} = sum_j a~_{t-1}[j] travels as per-warp partial sums next to the vector and is applied one step
//              late (the recursion is linear), so the step has ONE barrier.  c_t goes to cnorm[b][t].
//   backward : M = A.  beta_t = M w_{t+1} / c_{t+1} (1 at a clip's last frame); gamma_t = a~_t / c_t * beta_t overwrites
//              a~_t in place; w_t = b_t * beta_t.
// The dense state's row (sum_i v[i] M[u][i]) is accumulated from the values a thread has just produced, so it is known at
// the start of the next step like in the max-plus kernel.
#include <cstdlib>
#include <type_traits>

#include "vit_tmem.cuh"

namespace vit {

constexpr int fNJ = 4;                     // consecutive states per thread
constexpr int fCPT = 2;                    // clips per thread
constexpr int fCS = 4;                     // pipelines (clip groups) per CTA
constexpr int fMB = fCPT * fCS;            // 8 clips per CTA
constexpr int fTGW = 3;                    // warps per pipeline: 96 state groups x 4 = 384 states
constexpr int fPipeThreads = 32 * fTGW;
constexpr int fThreads = fPipeThreads * fCS;
constexpr int fMaxS = 32 * fTGW * fNJ;     // 384
constexpr int fRowLen = 416;               // as in vit_banded.cu: state i lives at float index i + DP

__device__ __forceinline__ void fpipe_bar_sync(int cs) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + cs), "n"(fPipeThreads) : "memory");
}
// 1 / x for a normaliser (0 for x == 0: an impossible observation sequence).  One MUFU.RCP instead of the IEEE division's
// range check + slow-path call; every thread of a pipeline evaluates it on the same value, so all get the same result.
__device__ __forceinline__ float frcp_pos(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  if (x < 1.1754944e-38f) {                               // zero or denormal: rescale by 2^64 either side (no division call)
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 18446744073709551616.f));
    r = x > 0.f ? r * 18446744073709551616.f : 0.f;
  }
  return r;
}
// Sums over the 32 lanes of a warp of NV values per lane by recursive halving; value k ends up in the lanes whose top
// log2(NV) lane bits spell k (NV = 4: lanes 0-7 hold value 0, 8-15 value 1, ...).
template <int NV>
__device__ __forceinline__ float warp_sums(const float* x, int lane) {
  static_assert(NV == 1 || NV == 2 || NV == 4, "NV");
  float a[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) a[i] = x[i];
  int bit = 16;
#pragma unroll
  for (int width = NV / 2; width >= 1; width /= 2) {
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < width; ++i) {
      const float keep = upper ? a[i + width] : a[i];
      const float send = upper ? a[i] : a[i + width];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
    bit >>= 1;
  }
  float v = a[0];
#pragma unroll
  for (int b2 = 16 / NV; b2 > 0; b2 >>= 1) v += __shfl_xor_sync(0xffffffffu, v, b2);
  return v;
}

template <int D, bool BWD>
__global__ void __launch_bounds__(fThreads, 1)
fb_banded_pass_kernel(const float* __restrict__ A, const float* __restrict__ pi, const float* __restrict__ lik,
                      const int32_t* __restrict__ lengths, int B, int T_max, int S, int jd, float* __restrict__ gamma,
                      float* __restrict__ cnorm, int q, const int* __restrict__ conv_flag) {
  // the matrix has the scaled-Toeplitz form and the convolution kernels (vit_fb_conv.cu) have done the work
  if (*conv_flag != 0) return;
  constexpr int W = 2 * D + 1;
  constexpr int DP = (D + 3) / 4 * 4;
  constexpr int U0 = DP - D;                         // window element of cell (r, n) = w[U0 + r + n]
  constexpr int NW4 = (U0 + W - 1 + fNJ + 3) / 4;    // float4s in a thread's window
  // as in banded_forward_kernel: 4 x 29 band entries do not fit the register file next to everything else, so for
  // D > 12 the last SPL offsets of every state live in tensor memory (columns of the thread's own lane)
  constexpr int SPL = D > 12 ? 12 : (D == 12 ? 4 : 0);
  constexpr int WR = W - SPL;
  constexpr int M_WAIT = (WR + U0 - 3 + 3) / 4;
  constexpr int NTC = 4 * SPL;
  constexpr int kTmemCols = 256;
  constexpr int KINDS = BWD ? 2 : 1;                 // staged inputs: b_t (and alpha~_t in the backward pass)
  extern __shared__ __align__(16) float s_dyn[];
  __shared__ uint32_t s_tmem_base;
  // the vector of the previous step: [fCS][2][fCPT][fRowLen]
  float (*s_v)[2][fCPT][fRowLen] = reinterpret_cast<float (*)[2][fCPT][fRowLen]>(s_dyn);
  // the step's input rows, fetched ONE STEP AHEAD by cp.async (LDGSTS): [fCS][2][KINDS][fCPT][fMaxS].  A step of this
  // kernel is shorter than an L2 round trip under load, so loads issued at the top of the step and used at its end
  // stalled ~1 warp per issue cycle on the long scoreboard (ncu); staged, the latency has a whole step to hide in and
  // nothing is held in registers meanwhile.  The dense state's slot of a row is never written (stays 0): its inputs go
  // to s_ind.
  float (*s_in)[2][KINDS][fCPT][fMaxS] =
      reinterpret_cast<float (*)[2][KINDS][fCPT][fMaxS]>(s_dyn + fCS * 2 * fCPT * fRowLen);
  __shared__ float s_ind[fCS][2][3][fCPT];           // per clip: b_t[jd], alpha~_t[jd], c_t
  __shared__ __align__(16) float s_g[BWD ? fCS : 1][fCPT][fMaxS];   // backward: gamma_t on its way out
  __shared__ float s_partS[fCS][2][fTGW][fCPT];      // per-warp partial sums of the vector (forward: the normaliser)
  __shared__ float s_partD[fCS][2][fTGW][fCPT];      // per-warp partial sums of the dense state's row
  __shared__ int s_len[fMB];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cs = warp / fTGW, tgw = warp - cs * fTGW;
  const int tg = tgw * 32 + lane;
  const int j0 = fNJ * tg;

  // the operand: forward M[out][in] = A[in][out] (A is stored source-major), backward M = A
  auto getM = [&](int out, int in) { return BWD ? A[(size_t)out * S + in] : A[(size_t)in * S + out]; };

  float a[fNJ][WR], acol[fNJ], arow[fNJ];
  bool jn_ok[fNJ];
  uint32_t taddr = 0;
  if constexpr (SPL > 0) {
    if (warp == 0) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(&s_tmem_base)), "n"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    taddr = __shfl_sync(0xffffffffu, s_tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * NTC), 0);
  }
#pragma unroll
  for (int n = 0; n < fNJ; ++n) {
    const int j = j0 + n;
    const bool jok = j < S && j != jd;
    jn_ok[n] = jok;
#pragma unroll
    for (int r = 0; r < WR; ++r) {
      const int i = j + r - D;
      a[n][r] = (jok && i >= 0 && i < S && i != jd) ? getM(j, i) : 0.f;
    }
    if constexpr (SPL > 0) {
#pragma unroll
      for (int g4 = 0; g4 < SPL / 4; ++g4) {
        float hi[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = j + WR + 4 * g4 + k - D;
          hi[k] = (jok && i >= 0 && i < S && i != jd) ? getM(j, i) : 0.f;
        }
        tmem_st4(taddr + g4 * 16 + n * 4, make_float4(hi[0], hi[1], hi[2], hi[3]));
      }
    }
    acol[n] = (jok && jd >= 0) ? getM(j, jd) : 0.f;     // the dense state as an input of state j
    arow[n] = (jok && jd >= 0) ? getM(jd, j) : 0.f;     // state j as an input of the dense state
  }
  if constexpr (SPL > 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  const float a_dd = jd >= 0 ? A[(size_t)jd * S + jd] : 0.f;
  const bool wfull = __all_sync(0xffffffffu, jn_ok[0] && jn_ok[1] && jn_ok[2] && jn_ok[3]);
  const long long gamma_delta = reinterpret_cast<const char*>(gamma) - reinterpret_cast<const char*>(lik);
  // Global loads and stores are NOT done in the owner layout (lane l <-> states 4 l .. 4 l + 3: a warp-wide 4-byte access
  // at a 16-byte lane stride spans 512 bytes = 4-5 LSU wavefronts and 4-way bank conflicts on the shared-memory side);
  // lane l moves elements 128 tgw + l + 32 n of its warp's 128-state chunk instead (one wavefront per instruction), and
  // shared memory does the transposition: the rows a warp consumes / produces are exactly the ones its own lanes move.
  const int jc0 = 128 * tgw + lane;
  bool cp_ok[fNJ];
#pragma unroll
  for (int n = 0; n < fNJ; ++n) cp_ok[n] = jc0 + 32 * n < S && jc0 + 32 * n != jd;
  const bool cfull = __all_sync(0xffffffffu, cp_ok[0] && cp_ok[1] && cp_ok[2] && cp_ok[3]);
  const bool own_jd = jd >= 0 && tgw == 0 && lane == 3;   // the lane that fetches the dense state's inputs
  for (int x = tid; x < fCS * 2 * KINDS * fCPT * fMaxS; x += fThreads) (&s_in[0][0][0][0][0])[x] = 0.f;

  const int c0 = cs * fCPT;
  const int ncl = min(fCPT, q - c0);

  auto run_pass = [&](auto cpt_tag, int seq0) {
    constexpr int CPT = decltype(cpt_tag)::value;
    int len[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) len[c] = s_len[c0 + c];
    int maxlen = 0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) maxlen = max(maxlen, len[c]);
    if (maxlen == 0) return;
    // running pointer to element jc0 of the lik row staged LAST (one step ahead of the step being computed); the row
    // this step writes (gamma buffer) sits at the fixed byte distance odelta from it.  A slot past the batch aliases the
    // last clip's rows (its length is 0: nothing is stored), a clip shorter than its pipeline's longest keeps stepping.
    const float* pin[CPT];
    int cni[CPT];                                           // cnorm index of frame 0 of my clips
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int b = min(seq0 + c0 + c, B - 1);
      pin[c] = lik + ((size_t)b * T_max + (BWD ? maxlen - 1 : 0)) * S + jc0;
      cni[c] = b * T_max;
    }
    const long long dS = BWD ? -(long long)S : (long long)S;
    const long long odelta = gamma_delta - dS * 4;
    float xd[CPT];                                          // the dense state's value of the previous step
#pragma unroll
    for (int c = 0; c < CPT; ++c) xd[c] = 0.f;
    float cnext[CPT];                                       // backward: c_{t+1}
#pragma unroll
    for (int c = 0; c < CPT; ++c) cnext[c] = 1.f;

    // requests the inputs of step it_f (frame t_f; pin[] points at its rows) into stage buffer it_f & 1
    auto stage = [&](int it_f) {
      if (it_f < maxlen) {
        const int t_f = BWD ? maxlen - 1 - it_f : it_f;
        const int b2 = it_f & 1;
        const bool pf = (lane & 7) == 0 && (BWD ? t_f >= 4 : t_f + 4 < T_max);
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const float* src = pin[c];
          const float* srca = reinterpret_cast<const float*>(reinterpret_cast<const char*>(src) + gamma_delta);
          if (pf) {
            // a step is shorter than an HBM round trip: pull the lines 4 frames further into L2 now
            // (lanes 0, 8, 16, 24: one per 128-byte line of the warp's 512-byte chunk)
            const long long far = 4 * dS - lane + (lane >> 3) * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src + far));
            if (BWD) asm volatile("prefetch.global.L2 [%0];" ::"l"(srca + far));
          }
          const uint32_t d0 = smem_u32(&s_in[cs][b2][0][c][jc0]);
          const uint32_t d1 = smem_u32(&s_in[cs][b2][KINDS - 1][c][jc0]);
#pragma unroll
          for (int n = 0; n < fNJ; ++n)
            if (cfull || cp_ok[n]) {
              asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 128 * n), "l"(src + 32 * n) : "memory");
              if (BWD) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d1 + 128 * n), "l"(srca + 32 * n) : "memory");
            }
          if (own_jd) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&s_ind[cs][b2][0][c])), "l"(src + (jd - jc0)) : "memory");
            if (BWD) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&s_ind[cs][b2][1][c])), "l"(srca + (jd - jc0)) : "memory");
          }
          if (BWD && tgw == 0 && lane == 4 + c && t_f < len[c])
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&s_ind[cs][b2][2][c])), "l"(cnorm + cni[c] + t_f) : "memory");
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    fpipe_bar_sync(cs);

    // forward: steps 0 .. maxlen (the last one only delivers the normaliser of frame maxlen - 1);  backward: maxlen steps
    const int n_steps = BWD ? maxlen : maxlen + 1;
    for (int it = 0; it < n_steps; ++it) {
      const int t = BWD ? maxlen - 1 - it : it;
      const int buf = it & 1;
      const bool fin = !BWD && it == maxlen;               // forward tail: no frame to load
      // this step's inputs are in stage buffer `buf` (complete and visible since the barrier that ended the previous
      // step); the next step's are requested now
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        pin[c] += dS;
        asm volatile("" : "+l"(pin[c]));     // (keep the running pointers in registers: ptxas re-derives them otherwise)
      }
      stage(it + 1);
      float v[CPT][fNJ], xdn[CPT];
      if (!BWD && it == 0) {
        // alpha~_0 = pi * b_0
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const float4 e4 = reinterpret_cast<const float4*>(s_in[cs][buf][0][c])[tg];
          const float e[fNJ] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
          for (int n = 0; n < fNJ; ++n) v[c][n] = jn_ok[n] ? pi[j0 + n] * e[n] : 0.f;
          xdn[c] = jd >= 0 ? pi[jd] * s_ind[cs][buf][0][c] : 0.f;
        }
      } else {
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const float* prev = s_v[cs][buf ^ 1][c];
          float acc[fNJ];
#pragma unroll
          for (int n = 0; n < fNJ; ++n) acc[n] = 0.f;
          const float4* row4 = reinterpret_cast<const float4*>(prev) + tg;
          float ah[NTC > 0 ? NTC : 1];
#pragma unroll
          for (int m = 0; m < NW4; ++m) {
            const float4 wv4 = row4[m];
            const float wv[4] = {wv4.x, wv4.y, wv4.z, wv4.w};
            if constexpr (SPL > 0) {
#pragma unroll
              for (int g4 = 0; g4 < SPL / 4; ++g4) {
                if (m == M_WAIT + g4 - 1) tmem_ld<16>(taddr + 16 * g4, ah + 16 * g4);
                if (m == M_WAIT + g4) tmem_wait_ld<16>(ah + 16 * g4);
              }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
              for (int n = 0; n < fNJ; ++n) {
                const int r = 4 * m + k - U0 - n;
                if (r >= 0 && r < WR) acc[n] = fmaf(wv[k], a[n][r], acc[n]);
                else if (r >= WR && r < W) acc[n] = fmaf(wv[k], ah[((r - WR) >> 2) * 16 + n * 4 + ((r - WR) & 3)], acc[n]);
              }
          }
          const float* pss = &s_partS[cs][buf ^ 1][0][c];
          const float* pdd = &s_partD[cs][buf ^ 1][0][c];
          const float dd = fmaf(xd[c], a_dd, (pdd[0] + pdd[fCPT]) + pdd[2 * fCPT]);
#pragma unroll
          for (int n = 0; n < fNJ; ++n) acc[n] = fmaf(xd[c], acol[n], acc[n]);
          if constexpr (!BWD) {
            // c_{t-1} = sum_j alpha~_{t-1}[j]: the warps' partial sums + the dense state
            const float tot = ((pss[0] + pss[fCPT]) + pss[2 * fCPT]) + xd[c];
            if (tgw == 0 && lane == 2 && t - 1 < len[c]) cnorm[cni[c] + t - 1] = tot;
            const float inv = frcp_pos(tot);
            const float4 e4 = reinterpret_cast<const float4*>(s_in[cs][buf][0][c])[tg];
            const float e[fNJ] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
            for (int n = 0; n < fNJ; ++n) v[c][n] = (acc[n] * inv) * e[n];
            xdn[c] = (dd * inv) * s_ind[cs][buf][0][c];
          } else {
            // beta_t = M w_{t+1} / c_{t+1} (1 at the clip's last frame); gamma_t = alpha~_t / c_t * beta_t; w_t = b_t beta_t
            const bool lv = t < len[c], last = t == len[c] - 1;
            const float ct = lv ? s_ind[cs][buf][2][c] : 1.f;
            const float invn = frcp_pos(cnext[c]), invc = frcp_pos(ct);
            const float4 e4 = reinterpret_cast<const float4*>(s_in[cs][buf][0][c])[tg];
            const float4 al4 = reinterpret_cast<const float4*>(s_in[cs][buf][KINDS - 1][c])[tg];
            const float e[fNJ] = {e4.x, e4.y, e4.z, e4.w}, al[fNJ] = {al4.x, al4.y, al4.z, al4.w};
            float* po = reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pin[c])) + odelta);
            float g[fNJ];
#pragma unroll
            for (int n = 0; n < fNJ; ++n) {
              const float be = last ? 1.f : acc[n] * invn;
              g[n] = (al[n] * invc) * be;
              v[c][n] = lv ? e[n] * be : 0.f;
            }
            // gamma_t leaves through the scratch row: owner layout in, lane-contiguous out
            reinterpret_cast<float4*>(s_g[cs][c])[tg] = make_float4(g[0], g[1], g[2], g[3]);
            __syncwarp();
            if (lv) {
#pragma unroll
              for (int n = 0; n < fNJ; ++n)
                if (cfull || cp_ok[n]) st_global_cs_f32(po + 32 * n, s_g[cs][c][jc0 + 32 * n]);
            }
            const float bed = last ? 1.f : dd * invn;
            if (jd >= 0 && tgw == 0 && lane == 1 && lv) st_global_cs_f32(po + (jd - jc0), (s_ind[cs][buf][1][c] * invc) * bed);
            xdn[c] = (lv && jd >= 0) ? s_ind[cs][buf][0][c] * bed : 0.f;
            cnext[c] = ct;
          }
        }
      }

      // the new vector: HBM (forward), shared memory, the warp's partial sums
      float part[2 * CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        float ss = 0.f, sd = 0.f;
#pragma unroll
        for (int n = 0; n < fNJ; ++n) {
          ss += v[c][n];
          sd = fmaf(v[c][n], arow[n], sd);
        }
        part[2 * c] = sd;
        part[2 * c + 1] = ss;
        reinterpret_cast<float4*>(s_v[cs][buf][c] + DP)[tg] = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
        if constexpr (!BWD) {
          // alpha~_t -> the gamma buffer, read back from the row just written in the lane-contiguous layout
          __syncwarp();
          if (!fin && t < len[c]) {
            float* po = reinterpret_cast<float*>(reinterpret_cast<char*>(const_cast<float*>(pin[c])) + odelta);
#pragma unroll
            for (int n = 0; n < fNJ; ++n)
              if (cfull || cp_ok[n]) st_global_cs_f32(po + 32 * n, s_v[cs][buf][c][DP + jc0 + 32 * n]);
            if (jd >= 0 && tgw == 0 && lane == 1) st_global_cs_f32(po + (jd - jc0), xdn[c]);
          }
        }
        xd[c] = xdn[c];
      }
      if constexpr (BWD) {
        // only the dense state's row is summed over the warp
        float pd[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) pd[c] = part[2 * c];
        const float tot = warp_sums<CPT>(pd, lane);
        if ((lane & (32 / CPT - 1)) == 0) s_partD[cs][buf][tgw][lane / (32 / CPT)] = tot;
      } else {
        const float tot = warp_sums<2 * CPT>(part, lane);
        if ((lane & (16 / CPT - 1)) == 0) {
          const int k = lane / (16 / CPT);                   // value index: 2 c (dense row) or 2 c + 1 (vector sum)
          if (k & 1) s_partS[cs][buf][tgw][k >> 1] = tot;
          else s_partD[cs][buf][tgw][k >> 1] = tot;
        }
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");   // the next step's inputs have landed (mine; the barrier: everyone's)
      fpipe_bar_sync(cs);
    }
  };

  for (int seq0 = blockIdx.x * q; seq0 < B; seq0 += gridDim.x * q) {
    __syncthreads();
    if (tid < fMB) {
      const int b = seq0 + tid;
      s_len[tid] = (tid < q && b < B) ? min(lengths ? lengths[b] : T_max, T_max) : 0;
    }
    // (re-)arm the vector rows and the partial sums: pads, out-of-range states and the dense state's slot hold 0
    for (int x = tid; x < fCS * 2 * fCPT * fRowLen; x += fThreads) (&s_v[0][0][0][0])[x] = 0.f;
    for (int x = tid; x < fCS * 2 * fTGW * fCPT; x += fThreads) {
      (&s_partS[0][0][0][0])[x] = 0.f;
      (&s_partD[0][0][0][0])[x] = 0.f;
    }
    __syncthreads();
    if (ncl >= 2) run_pass(std::integral_constant<int, 2>{}, seq0);
    else if (ncl == 1) run_pass(std::integral_constant<int, 1>{}, seq0);
  }
  if constexpr (SPL > 0) {
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem_base), "n"(kTmemCols) : "memory");
  }
}

// vit_fb.cu
__global__ void fb_loglik_kernel(const float* __restrict__ cnorm, const int32_t* __restrict__ lengths, int B, int T_max,
                                 float* __restrict__ loglik, const int* __restrict__ flag, int want);
bool fb_supported(int S);
size_t fb_workspace_bytes(int B, int T_max, int S);
int fb_run(const float* A, const float* pi, const float* lik, const int32_t* lengths, int B, int T_max, int S,
           void* workspace, size_t workspace_bytes, float* gamma, float* loglik, cudaEvent_t ev0, cudaEvent_t ev1,
           cudaStream_t stream, const int* skip);

static int fb_banded_template_D(int d) {
  const int opts[] = {4, 8, 12, 14};
  for (int o : opts) if (d <= o) return o;
  return -1;
}

// vit_fb_conv.cu: the scaled-Toeplitz form (what the reference's builders produce) as a convolution, one warp per clip
size_t fb_conv_params_bytes();
int fb_conv_template_D(int S, int d);
int fb_conv_detect(const float* A, int S, const vit_structure* st, int Dt, void* params, cudaStream_t stream);
int fb_conv_passes(int D, const void* params, const float* pi, const float* lik, const int32_t* lengths, int B, int T_max,
                   int S, int jd, float* gamma, float* cnorm, cudaStream_t stream);

// general band kernel of this file: S <= 384, d <= 14
static bool fb_banded_narrow(int S, const vit_structure* st) { return S <= fMaxS && fb_banded_template_D(st->halfwidth) > 0; }

// `st` describes the PROBABILITY-domain matrix (vit_analyze_structure_f32 on A; the band and the dense state are the same
// whichever way round the matrix is stored): band + one dense state, every other entry exactly 0.  S <= 384 with d <= 14:
// convolution kernel or general band kernel; up to S = 768 with d <= 56 (jdc, imm): convolution kernel or, when the band is
// not scaled-Toeplitz, the dense FFMA kernel.
bool fb_banded_supported(int S, const vit_structure* st) {
  if (!st || st->kind != 1) return false;
  if (st->background != 0.f) return false;
  if (S < 2 || st->halfwidth < 0) return false;
  if (st->dense_index < -1 || st->dense_index >= S) return false;
  if (fb_banded_narrow(S, st)) return true;
  return fb_conv_template_D(S, st->halfwidth) > 0 && fb_supported(S);
}

static size_t fb_banded_cnorm_bytes(int B, int T_max) { return align_up((size_t)(B > 0 ? B : 1) * T_max * sizeof(float), 256); }
// what the structured kernels keep at the END of the workspace (their normalisers and the parameter block): the front
// belongs to the dense kernels, which may run as the fall-back in the same call
size_t fb_banded_extra_bytes(int B, int T_max) { return fb_banded_cnorm_bytes(B, T_max) + fb_conv_params_bytes(); }

int fb_banded_run(const float* A, const float* pi, const float* lik, const int32_t* lengths, int B, int T_max, int S,
                  const vit_structure* st, void* workspace, size_t workspace_bytes, float* gamma, float* loglik,
                  cudaStream_t stream) {
  if (!fb_banded_supported(S, st)) return VIT_ERR_UNSUPPORTED_ALGO;
  if ((long long)B * T_max >= 0x7fffffffLL) return VIT_ERR_INVALID_ARGUMENT;   // (frame indices are int32 in the kernel)
  const size_t extra = fb_banded_extra_bytes(B, T_max);
  if (workspace_bytes < extra) return VIT_ERR_WORKSPACE_TOO_SMALL;
  if (B == 0) return VIT_OK;
  const size_t front = (workspace_bytes - extra) & ~(size_t)255;
  float* cnorm = (float*)((char*)workspace + front);
  void* conv_params = (char*)cnorm + fb_banded_cnorm_bytes(B, T_max);         // its first word is the verdict
  const bool narrow = fb_banded_narrow(S, st);
  if (!narrow && front < fb_workspace_bytes(B, T_max, S)) return VIT_ERR_WORKSPACE_TOO_SMALL;
  // frames past a clip's length carry gamma = 0
  if (lengths) VIT_CUDA_TRY(cudaMemsetAsync(gamma, 0, (size_t)B * T_max * S * sizeof(float), stream));
  const int D = narrow ? fb_banded_template_D(st->halfwidth) : fb_conv_template_D(S, st->halfwidth);
  // Scaled-Toeplitz check on the device, then BOTH kernel sets: the convolution kernels return at once unless the
  // check passed, the general kernels return at once if it did -- no round trip to the host.
  int rc = fb_conv_detect(A, S, st, D, conv_params, stream);
  if (rc != VIT_OK) return rc;
  rc = fb_conv_passes(D, conv_params, pi, lik, lengths, B, T_max, S, st->dense_index, gamma, cnorm, stream);
  if (rc != VIT_OK) return rc;
  if (loglik) {
    fb_loglik_kernel<<<(B + 3) / 4, 128, 0, stream>>>(cnorm, lengths, B, T_max, loglik, (const int*)conv_params, 1);
    note_launch();
    VIT_CUDA_TRY(cudaGetLastError());
  }
  if (!narrow)   // wide bands / 722 states that are not scaled-Toeplitz: the dense FFMA kernel (its own normalisers in the front)
    return fb_run(A, pi, lik, lengths, B, T_max, S, workspace, front, gamma, loglik, nullptr, nullptr, stream,
                  (const int*)conv_params);
  int num_sms = 148, dev = 0;
  VIT_CUDA_TRY(cudaGetDevice(&dev));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  // clips spread evenly over the SMs, as banded_decode does: `passes` trips of q <= 8 clips per CTA
  const int passes = (B + fMB * num_sms - 1) / (fMB * num_sms);
  int q = (B + passes * num_sms - 1) / (passes * num_sms);
  q = q < 1 ? 1 : (q > fMB ? fMB : q);
  const int want = (B + q - 1) / q;
  const int grid = want < num_sms ? want : num_sms;
  // dynamic shared memory: the vector double buffers + the cp.async stage of the inputs (1 kind forward, 2 backward)
  const size_t smem_f = (size_t)(fCS * 2 * fCPT * fRowLen + fCS * 2 * 1 * fCPT * fMaxS) * sizeof(float);
  const size_t smem_b = (size_t)(fCS * 2 * fCPT * fRowLen + fCS * 2 * 2 * fCPT * fMaxS) * sizeof(float);
#define VIT_FBB_CASE(DD)                                                                                              \
  case DD: {                                                                                                          \
    auto kf = fb_banded_pass_kernel<DD, false>;                                                                       \
    auto kb = fb_banded_pass_kernel<DD, true>;                                                                        \
    VIT_CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));                 \
    VIT_CUDA_TRY(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));                 \
    kf<<<grid, fThreads, smem_f, stream>>>(A, pi, lik, lengths, B, T_max, S, st->dense_index, gamma, cnorm, q,        \
                                           (const int*)conv_params);                                                  \
    kb<<<grid, fThreads, smem_b, stream>>>(A, pi, lik, lengths, B, T_max, S, st->dense_index, gamma, cnorm, q,        \
                                           (const int*)conv_params);                                                  \
  } break;
  switch (D) {
    VIT_FBB_CASE(4) VIT_FBB_CASE(8) VIT_FBB_CASE(12) VIT_FBB_CASE(14)
    default: return VIT_ERR_UNSUPPORTED_ALGO;
  }
#undef VIT_FBB_CASE
  note_launch(2);
  VIT_CUDA_TRY(cudaGetLastError());
  if (loglik) {
    fb_loglik_kernel<<<(B + 3) / 4, 128, 0, stream>>>(cnorm, lengths, B, T_max, loglik, (const int*)conv_params, 0);
    note_launch();
    VIT_CUDA_TRY(cudaGetLastError());
  }
  return VIT_OK;
}

}  // namespace vit
