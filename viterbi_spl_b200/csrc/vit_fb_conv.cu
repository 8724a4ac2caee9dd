// Forward-backward for SCALED-TOEPLITZ bands: the form the reference's transition matrices actually have.
//
// dcnet/viterbi_transition_matrix.py:81-98 (and the tonet / jdc / imm copies) fill the band with ONE jump histogram,
// `transition_matrix[i, j] = d_trans[j - i + d_max]`, normalise every row, and embed the result in the voiced/unvoiced
// switch with constant entries to and from the unvoiced state.  So, apart from rounding,
//
//     A[i][j] = kappa_i * b[j - i]   (|j - i| <= d),      A[i][u] = r,      A[u][j] = q,      A[u][u] = a_uu
//
// with ONE tap vector b (2d + 1 values) and a per-source scale kappa_i (1 for interior rows, > 1 for the truncated rows
// near the ends).  The band product is then a CONVOLUTION:  (alpha A)[j] = sum_i (alpha_i kappa_i) b[j - i]  (forward),
// (A w)[i] = kappa_i sum_j b[j - i] w[j]  (backward).  A thread needs 2d + 1 tap registers whatever the number of states it
// owns -- against 4 x (2d + 1) band registers for 4 states in the general kernel (vit_fb_banded.cu) -- so ONE WARP takes a
// whole clip (12 consecutive states per lane, 384 >= S) and the step needs no block barrier, no per-warp partial sums in
// shared memory and a third fewer instructions per state; the dense state's row and column are scalars
// (sum_band alpha * r, alpha_u * q).  Clips are independent warps: 16 per SM hide each other's latencies.
//
// fb_conv_detect_kernel checks the form ON THE DEVICE at every call (the API only has the device copy of A): kappa_i =
// sum of row i's band / sum of the taps it covers, every band entry within 2e-6 relative of kappa_i b[j - i] (measured on
// the reference recipes: 2.3e-7 .. 5.7e-7, i.e. the rounding of the row normalisation), constant dense row / column,
// kappa_i > 0.  The verdict is a flag in the workspace; the convolution kernels return at once when it is 0 and the
// general banded kernels when it is 1, so nothing synchronises with the host.  Replacing A by kappa_i b[j - i] perturbs
// every entry by a few ulps -- the same order as the kernel's own fp32 rounding; |gamma - float64 oracle| (which uses
// the TRUE matrix) is checked at 1e-4 like every other kernel (tests/test_gpu_fb_banded.py).  Parity unpinned: the
// reference has no forward-backward (oracle/fb_oracle.py).
#include <cstdlib>

#include "vit_common.cuh"

namespace vit {

constexpr int vNJ = 12;                    // consecutive states per lane
constexpr int vMaxS = 32 * vNJ;            // 384
constexpr int vDP = 16;                    // state i lives at float index i + vDP of a vector row
constexpr int vRow = vMaxS + 2 * vDP;      // 416
constexpr int vStages = 3;                 // input rows are fetched two steps ahead
constexpr int vMaxTaps = 128;               // 2 * 56 + 1 taps of the widest band, rounded up
constexpr int vMaxSW = 2 * vMaxS;          // the wide kernel: two blocks of 384 states per warp

struct ConvParams {
  int flag;                                // 1: the matrix has the scaled-Toeplitz form, the convolution kernels run
  float r_in;                              // A[i][u]: every band state -> the dense state
  float q_out;                             // A[u][j]: the dense state -> every band state
  float a_uu;
  float tap[vMaxTaps];                     // tap[x + Dt] = b[x], x = j - i, Dt = the kernel instance's half-width
  float kappa[vMaxSW];
};

size_t fb_conv_params_bytes() { return align_up(sizeof(ConvParams), 256); }

// one block of 384 threads; thread i checks source rows i and i + 384
__global__ void fb_conv_detect_kernel(const float* __restrict__ A, int S, int jd, int d, int Dt, int enable,
                                      ConvParams* __restrict__ prm) {
  // d: the band's half-width; Dt >= d: the half-width of the kernel instance that will run (taps are stored for it)
  __shared__ float s_b[vMaxTaps];
  __shared__ int s_istar;
  __shared__ float s_rq[2];
  const int i = threadIdx.x, W = 2 * d + 1;
  if (i == 0) {
    // a source row whose whole band exists: the taps are read from it
    int istar = -1;
    for (int k = 0; k < S && istar < 0; ++k) {
      const int c = (S / 2 + k) % S;
      if (c - d >= 0 && c + d < S && (jd < c - d || jd > c + d)) istar = c;
    }
    s_istar = (enable && S <= vMaxSW && 2 * Dt + 1 <= vMaxTaps && d <= Dt) ? istar : -1;
    // the constants every dense-row / dense-column entry must equal
    const int i0 = jd == 0 ? 1 : 0;
    s_rq[0] = (jd >= 0 && S > 1) ? A[(size_t)i0 * S + jd] : 0.f;
    s_rq[1] = (jd >= 0 && S > 1) ? A[(size_t)jd * S + i0] : 0.f;
  }
  __syncthreads();
  const int istar = s_istar;
  if (istar < 0) {
    if (i == 0) prm->flag = 0;
    return;
  }
  if (i == 0) {
    float sum = 0.f;
    for (int r = 0; r < W; ++r) sum += A[(size_t)istar * S + istar + r - d];
    for (int r = 0; r < W; ++r) s_b[r] = sum > 0.f ? A[(size_t)istar * S + istar + r - d] / sum : 0.f;
  }
  __syncthreads();
  bool ok = true;
  for (int i2 = i; i2 < vMaxSW; i2 += vMaxS) {
    float kap = 0.f;
    if (i2 < S && i2 != jd) {
      float sa = 0.f, sb = 0.f;
      for (int r = 0; r < W; ++r) {
        const int j = i2 + r - d;
        if (j >= 0 && j < S && j != jd) { sa += A[(size_t)i2 * S + j]; sb += s_b[r]; }
      }
      kap = sb > 0.f ? sa / sb : 0.f;
      ok = ok && kap > 0.f && kap < 1e30f;
      for (int r = 0; r < W; ++r) {
        const int j = i2 + r - d;
        if (j >= 0 && j < S && j != jd) {
          const float ref = kap * s_b[r];
          ok = ok && fabsf(A[(size_t)i2 * S + j] - ref) <= 2e-6f * ref + 1e-37f;
        }
      }
      if (jd >= 0) {
        ok = ok && fabsf(A[(size_t)i2 * S + jd] - s_rq[0]) <= 2e-6f * s_rq[0];
        ok = ok && fabsf(A[(size_t)jd * S + i2] - s_rq[1]) <= 2e-6f * s_rq[1];
      }
    }
    prm->kappa[i2] = kap;
  }
  const int all_ok = __syncthreads_and(ok ? 1 : 0);
  if (i < vMaxTaps) {
    const int x = i - Dt;                                  // tap[x + Dt] = b[x]
    prm->tap[i] = (x >= -d && x <= d) ? s_b[x + d] : 0.f;
  }
  if (i == 0) {
    prm->flag = all_ok;
    prm->r_in = s_rq[0];
    prm->q_out = s_rq[1];
    prm->a_uu = jd >= 0 ? A[(size_t)jd * S + jd] : 0.f;
  }
}

__device__ __forceinline__ float vrcp_pos(float x) {        // as frcp_pos in vit_fb_banded.cu
  // branch-free (a branch here would split the step into basic blocks the scheduler cannot interleave across): the
  // direct reciprocal for normal x, the one rescaled by 2^64 either side for denormal x, 0 for x == 0
  float r, r2;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r2) : "f"(x * 18446744073709551616.f));
  r2 = x > 0.f ? r2 * 18446744073709551616.f : 0.f;
  return x < 1.1754944e-38f ? r2 : r;
}
__device__ __forceinline__ float vwarp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// One warp (= one block) per clip.  NF >= 0: the host promises that the band states are exactly [0, Sv) (dense state last or
// absent), Sv in (32 NF, 32 NF + 32]: element rows n < NF of the lane-contiguous layout need no predicate, row NF one
// compare, rows past it do not exist.  NF = -1: any layout, every element predicated.
template <int D, bool BWD, int NF>
__global__ void __launch_bounds__(32, 16)
fb_conv_pass_kernel(const ConvParams* __restrict__ prm, const float* __restrict__ pi, const float* __restrict__ lik,
                    const int32_t* __restrict__ lengths, int T_max, int S, int jd, float* __restrict__ gamma,
                    float* __restrict__ cnorm) {
  if (prm->flag == 0) return;
  constexpr int W = 2 * D + 1;
  constexpr int U0 = vDP - D;                         // row offset (relative to 12 lane) of the first window element
  constexpr int M0 = U0 / 4, M1 = (U0 + vNJ - 1 + 2 * D) / 4;    // float4s of a lane's window: [M0, M1]
  constexpr int KINDS = BWD ? 2 : 1;
  constexpr int PFL = 2;                              // L2 prefetches per lane for 4 rows of <= 384 floats
  extern __shared__ __align__(16) float sm[];
  float* s_row = sm;                                  // [2][vRow]: the vector the sweep reads (forward: alpha~ kappa, backward: w)
  float* s_in = sm + 2 * vRow;                        // [vStages][KINDS][vMaxS]: b_t (and alpha~_t) rows, fetched by cp.async
  float* s_out = s_in + vStages * KINDS * vMaxS;      // [vMaxS]: this step's output row on its way to HBM
  __shared__ float s_ind[vStages][3];                 // the dense state's b_t and alpha~_t (never part of a staged row), c_t

  const int lane = threadIdx.x;
  const int b = blockIdx.x;
  const int len = min(lengths ? lengths[b] : T_max, T_max);    // (a length past T_max would walk off the clip's rows)
  if (len <= 0) return;
  const int j0 = vNJ * lane;

  float tap[W], kap[vNJ];
#pragma unroll
  for (int r = 0; r < W; ++r) tap[r] = prm->tap[BWD ? W - 1 - r : r];      // tap[(out - in) + D]
#pragma unroll
  for (int n = 0; n < vNJ; ++n) kap[n] = prm->kappa[j0 + n];               // 0 for the dense state and states >= S
  const float r_in = prm->r_in, q_out = prm->q_out, a_uu = prm->a_uu;
  const bool has_d = jd >= 0;

  for (int x = lane; x < 2 * vRow; x += 32) s_row[x] = 0.f;
  for (int x = lane; x < vStages * KINDS * vMaxS; x += 32) s_in[x] = 0.f;
  __syncwarp();

  // lane-contiguous element ownership for everything that touches HBM: lane l moves elements l + 32 n of a row
  const int Sv = has_d ? S - 1 : S;                    // (NF >= 0 only) band states are [0, Sv)
  auto elem_ok = [&](int n) {
    if constexpr (NF >= 0) return n < NF || (n == NF && lane + 32 * NF < Sv);
    else return lane + 32 * n < S && lane + 32 * n != jd;
  };
  constexpr int NROWS = NF >= 0 ? NF + 1 : vNJ;        // element rows that can hold a band state
  const long long dS = BWD ? -(long long)S : (long long)S;
  const long long gamma_delta = reinterpret_cast<const char*>(gamma) - reinterpret_cast<const char*>(lik);
  const float* pin = lik + ((size_t)b * T_max + (BWD ? len - 1 : 0)) * S + lane;     // row staged next
  float* pout = gamma + ((size_t)b * T_max + (BWD ? len - 1 : 0)) * S + lane;        // row written next
  float* pc = cnorm + (size_t)b * T_max;

  // requests the input rows of step it_f into stage buffer it_f % vStages.  The slots of a staged row that hold no band
  // state (the dense state's, those past S) are never written and stay 0, so whatever is multiplied by them is 0: no
  // masks in the step.  The dense state's inputs go to s_ind (lane 0).
  auto stage = [&](int it_f) {
    if (it_f < len) {
      const int sb = it_f % vStages;
      float* dst = s_in + sb * KINDS * vMaxS + lane;
      const uint32_t d0 = smem_u32(dst), d1 = smem_u32(dst + (KINDS - 1) * vMaxS);
      const float* srca = reinterpret_cast<const float*>(reinterpret_cast<const char*>(pin) + gamma_delta);
#pragma unroll
      for (int n = 0; n < NROWS; ++n)
        if (elem_ok(n)) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 128 * n), "l"(pin + 32 * n) : "memory");
          if (BWD) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d1 + 128 * n), "l"(srca + 32 * n) : "memory");
        }
      if (BWD && lane == 1)                                // c_t of that step (t = len - 1 - it_f)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&s_ind[sb][2])), "l"(pc + (len - 1 - it_f)) : "memory");
      if (has_d && lane == 0) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&s_ind[sb][0])), "l"(pin + jd) : "memory");
        if (BWD) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&s_ind[sb][1])), "l"(srca + jd) : "memory");
      }
    }
    // HBM locality: the rows of a clip are one sequential stream, but a row is only 1.4 - 2.9 KB and thousands of clips
    // advance in lock step, so row-sized requests reach DRAM as scattered page-sized pieces (ncu: ~64 % of the copy
    // bandwidth at every batch size).  Every 4th step the warp pulls the NEXT FOUR rows of its stream (one contiguous
    // piece of 5.8 - 11.5 KB) into L2; the row-sized cp.asyncs then hit L2.
    if ((it_f & 3) == 0 && it_f + 4 < len) {
      const long long first = BWD ? 7 * dS : 4 * dS;        // lowest address of rows it_f + 4 .. it_f + 7 of the stream
      const char* base = reinterpret_cast<const char*>(pin - lane + first);
      const int nbytes = 4 * S * 4;
#pragma unroll
      for (int k = 0; k < PFL; ++k) {
        const int off = (lane + 32 * k) * 128;
        if (off < nbytes) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
          if (BWD) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + gamma_delta + off));
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pin += dS;
  };
  stage(0);
  stage(1);
  stage(2);

  float xd = 0.f;          // the dense state's value of the previous step (forward: alpha~_{t-1}[u]; backward: w_{t+1}[u])
  float ss = 0.f;          // my lane's part of the sum over the band states of the previous step's vector: reduced over
                           // the warp at the START of the next step, so the shuffle chain hides under the sweep's FFMAs
  float cprev = 1.f;       // backward: c_{t+1}

  // the end of a step: the new vector -> the row the next sweep reads (forward: scaled by kappa), the output -> the
  // scratch row -> HBM in the lane-contiguous layout; the inputs of step it + 3 are requested between the scratch reads
  // and the stores (in-order issue: their latencies overlap)
  auto finish_step = [&](int it, const float* v, const float* o, float xdn, float out_d) {
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < vNJ; ++n) s += v[n];
    ss = s;
    float* rown = s_row + (it & 1) * vRow + vDP;
#pragma unroll
    for (int m = 0; m < vNJ / 4; ++m) {
      float4 u;
      if (BWD) u = make_float4(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]);
      else u = make_float4(v[4 * m] * kap[4 * m], v[4 * m + 1] * kap[4 * m + 1], v[4 * m + 2] * kap[4 * m + 2], v[4 * m + 3] * kap[4 * m + 3]);
      reinterpret_cast<float4*>(rown)[3 * lane + m] = u;
      reinterpret_cast<float4*>(s_out)[3 * lane + m] = make_float4(o[4 * m], o[4 * m + 1], o[4 * m + 2], o[4 * m + 3]);
    }
    xd = xdn;
    __syncwarp();
    float og[NROWS];
#pragma unroll
    for (int n = 0; n < NROWS; ++n) og[n] = s_out[lane + 32 * n];
    stage(it + 3);
#pragma unroll
    for (int n = 0; n < NROWS; ++n)
      if (elem_ok(n)) st_global_cs_f32(pout + 32 * n, og[n]);
    if (has_d && lane == 0) st_global_cs_f32(pout + jd, out_d);
    pout += dS;
  };
  auto step_inputs = [&](int it, float* e, float& ed) {
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncwarp();
    const float* in = s_in + (it % vStages) * KINDS * vMaxS;
#pragma unroll
    for (int m = 0; m < vNJ / 4; ++m) {
      const float4 x = reinterpret_cast<const float4*>(in)[3 * lane + m];
      e[4 * m] = x.x; e[4 * m + 1] = x.y; e[4 * m + 2] = x.z; e[4 * m + 3] = x.w;
    }
    ed = has_d ? s_ind[it % vStages][0] : 0.f;
    return in;
  };

  if constexpr (!BWD) {
    // step 0: alpha~_0 = pi * b_0
    float e[vNJ], ed, v[vNJ];
    step_inputs(0, e, ed);
#pragma unroll
    for (int n = 0; n < vNJ; ++n) v[n] = j0 + n < S ? pi[j0 + n] * e[n] : 0.f;
    const float xdn = has_d ? pi[jd] * ed : 0.f;
    finish_step(0, v, v, xdn, xdn);
  }
  for (int it = BWD ? 0 : 1; it < len; ++it) {
    const int t = BWD ? len - 1 - it : it;
    float e[vNJ], ed;
    const float* in = step_inputs(it, e, ed);
    // the band: acc[out] = sum_in row[in] tap[(out - in) + D]  (forward: on top of the dense state's contribution).
    // The warp sum of the previous step's vector (5 dependent shuffles: ~150 clocks of latency, a third of the kernel's
    // stall samples when it sat in front of the sweep) is issued ONE ROUND PER WINDOW FLOAT4 inside the sweep: each
    // round carries a data dependency on accumulators of that iteration, so ptxas cannot hoist the chain back to the top.
    float acc[vNJ];
    const float acc0 = BWD ? 0.f : xd * q_out;
#pragma unroll
    for (int n = 0; n < vNJ; ++n) acc[n] = acc0;
    const float4* row4 = reinterpret_cast<const float4*>(s_row + ((it & 1) ^ 1) * vRow) + 3 * lane;
    float sband = ss;
#pragma unroll
    for (int m = M0; m <= M1; ++m) {
      const float4 x = row4[m];
      const float wv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int n = 0; n < vNJ; ++n) {
          const int rr = n - (4 * m + k) + vDP + D;        // in = 12 lane + 4 m + k - vDP, out = 12 lane + n
          if (rr >= 0 && rr < W) acc[n] = fmaf(wv[k], tap[rr], acc[n]);
        }
      if (m - M0 >= 1 && m - M0 <= 5) {
        // (a REAL data dependency on this iteration's accumulators -- an empty asm leaves no trace in the PTX and ptxas
        // hoisted the whole chain back to the loop top; 0 * acc is not foldable under IEEE rules.  The result is only
        // perturbed if an accumulator is inf / NaN, i.e. if the inputs were not finite.)
        sband = fmaf(acc[(5 * (m - M0)) % vNJ], 0.f, sband);
        sband += __shfl_xor_sync(0xffffffffu, sband, 32 >> (m - M0));
      }
    }
#pragma unroll
    for (int r = M1 - M0 + 1; r <= 5; ++r) sband += __shfl_xor_sync(0xffffffffu, sband, 32 >> r);   // (narrow windows)
    float v[vNJ], o[vNJ], xdn, out_d;
    if constexpr (!BWD) {
      // alpha~_t = ((alpha~_{t-1} kappa) * b + alpha~_{t-1}[u] q) / c_{t-1} * b_t ;  c_{t-1} = band sum + dense element
      const float tot = sband + xd;
      if (lane == 0) pc[t - 1] = tot;
      const float inv = vrcp_pos(tot);
#pragma unroll
      for (int n = 0; n < vNJ; ++n) o[n] = v[n] = (acc[n] * inv) * e[n];
      xdn = has_d ? (fmaf(sband, r_in, xd * a_uu) * inv) * ed : 0.f;
      out_d = xdn;
    } else {
      // beta_t = (kappa (b * w_{t+1}) + r w_{t+1}[u]) / c_{t+1}  (1 at the clip's last frame);  gamma_t = alpha~_t / c_t beta_t
      const float ct = s_ind[it % vStages][2];
      const float invc = vrcp_pos(ct), invn = vrcp_pos(cprev);
      const bool last = it == 0;
      const float add = xd * r_in;
      const float4* al4 = reinterpret_cast<const float4*>(in + vMaxS) + 3 * lane;
      const float ald = has_d ? s_ind[it % vStages][1] : 0.f;
#pragma unroll
      for (int m = 0; m < vNJ / 4; ++m) {
        const float4 x = al4[m];
        const float al[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int n = 4 * m + k;
          const float be = last ? 1.f : fmaf(kap[n], acc[n], add) * invn;
          o[n] = (al[k] * invc) * be;
          v[n] = e[n] * be;
        }
      }
      const float bed = last ? 1.f : fmaf(sband, q_out, xd * a_uu) * invn;
      out_d = (ald * invc) * bed;
      xdn = has_d ? ed * bed : 0.f;
      cprev = ct;
    }
    finish_step(it, v, o, xdn, out_d);
  }
  if (!BWD) {
    const float sband = vwarp_sum(ss);
    if (lane == 0) pc[len - 1] = sband + xd;
  }
}

// ---- wide bands / 722-state sets (jdc: +-40 of 721 bins; the imm HMM: +-56) ------------------------------------------
// Same convolution, one warp per clip, but 2d + 1 = 81 .. 113 taps do not fit the register file and 722 states are 24 per
// lane.  So: a lane owns TWO runs of 12 consecutive states, [12 l, 12 l + 12) and [384 + 12 l, 384 + 12 l + 12) -- both at
// the 48-byte lane stride whose LDS.128 are conflict-free -- and sweeps them one after the other; the taps live in shared
// memory and pass through a sliding window of 16 registers (one broadcast LDS.128 = 4 new taps per window float4, i.e. per
// up to 48 FFMAs), the vector's window likewise slides by one float4 per iteration.  kappa sits in shared memory too.
constexpr int wDPmax = 56;
template <int D, bool BWD, int NF>
__global__ void __launch_bounds__(32, 8)
fb_convw_pass_kernel(const ConvParams* __restrict__ prm, const float* __restrict__ pi, const float* __restrict__ lik,
                     const int32_t* __restrict__ lengths, int T_max, int S, int jd, float* __restrict__ gamma,
                     float* __restrict__ cnorm) {
  if (prm->flag == 0) return;
  static_assert(D % 4 == 0 && D <= wDPmax, "wide instance");
  constexpr int W = 2 * D + 1;
  constexpr int NM = (vNJ + 2 * D) / 4;               // float4s of a run's window: inputs [run start - D, run start + 11 + D]
  constexpr int KINDS = BWD ? 2 : 1;
  constexpr int ROW = vMaxSW + 2 * D;                 // state i at float index i + D
  constexpr int NG = D / 2 + 1;                       // tap groups of 4: tap rr in group rr >> 2; groups -1 and NG .. NG + 1 are 0
  constexpr int NEL = vMaxSW / 32;                    // 24 element rows in the lane-contiguous layout
  constexpr int PFL = 3;                              // L2 prefetches per lane for 4 rows of <= 768 floats
  extern __shared__ __align__(16) float sm[];
  float* s_row = sm;                                  // [2][ROW]
  float* s_in = s_row + 2 * ROW;                      // [vStages][KINDS][vMaxSW]
  float* s_out = s_in + vStages * KINDS * vMaxSW;     // [vMaxSW]
  float* s_kap = s_out + vMaxSW;                      // [vMaxSW]
  float* s_tap = s_kap + vMaxSW;                      // [(NG + 3) * 4]: group g at float4 index g + 1
  __shared__ float s_ind[vStages][3];

  const int lane = threadIdx.x;
  const int b = blockIdx.x;
  const int len = min(lengths ? lengths[b] : T_max, T_max);    // (a length past T_max would walk off the clip's rows)
  if (len <= 0) return;
  const float r_in = prm->r_in, q_out = prm->q_out, a_uu = prm->a_uu;
  const bool has_d = jd >= 0;

  for (int x = lane; x < 2 * ROW; x += 32) s_row[x] = 0.f;
  for (int x = lane; x < vStages * KINDS * vMaxSW; x += 32) s_in[x] = 0.f;
  for (int x = lane; x < vMaxSW; x += 32) s_kap[x] = prm->kappa[x];
  for (int x = lane; x < (NG + 3) * 4; x += 32) {
    const int rr = x - 4;                              // tap index (out - in) + D of this slot
    s_tap[x] = (rr >= 0 && rr < W) ? prm->tap[BWD ? W - 1 - rr : rr] : 0.f;
  }
  __syncwarp();

  const int Sv = has_d ? S - 1 : S;
  auto elem_ok = [&](int n) {
    if constexpr (NF >= 0) return n < NF || (n == NF && lane + 32 * NF < Sv);
    else return lane + 32 * n < S && lane + 32 * n != jd;
  };
  constexpr int NROWS = NF >= 0 ? NF + 1 : NEL;
  const long long dS = BWD ? -(long long)S : (long long)S;
  const long long gamma_delta = reinterpret_cast<const char*>(gamma) - reinterpret_cast<const char*>(lik);
  const float* pin = lik + ((size_t)b * T_max + (BWD ? len - 1 : 0)) * S + lane;
  float* pout = gamma + ((size_t)b * T_max + (BWD ? len - 1 : 0)) * S + lane;
  float* pc = cnorm + (size_t)b * T_max;

  auto stage = [&](int it_f) {
    if (it_f < len) {
      const int sb = it_f % vStages;
      float* dst = s_in + sb * KINDS * vMaxSW + lane;
      const uint32_t d0 = smem_u32(dst), d1 = smem_u32(dst + (KINDS - 1) * vMaxSW);
      const float* srca = reinterpret_cast<const float*>(reinterpret_cast<const char*>(pin) + gamma_delta);
#pragma unroll
      for (int n = 0; n < NROWS; ++n)
        if (elem_ok(n)) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d0 + 128 * n), "l"(pin + 32 * n) : "memory");
          if (BWD) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d1 + 128 * n), "l"(srca + 32 * n) : "memory");
        }
      if (BWD && lane == 1)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&s_ind[sb][2])), "l"(pc + (len - 1 - it_f)) : "memory");
      if (has_d && lane == 0) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&s_ind[sb][0])), "l"(pin + jd) : "memory");
        if (BWD) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(&s_ind[sb][1])), "l"(srca + jd) : "memory");
      }
    }
    // HBM locality: the rows of a clip are one sequential stream, but a row is only 1.4 - 2.9 KB and thousands of clips
    // advance in lock step, so row-sized requests reach DRAM as scattered page-sized pieces (ncu: ~64 % of the copy
    // bandwidth at every batch size).  Every 4th step the warp pulls the NEXT FOUR rows of its stream (one contiguous
    // piece of 5.8 - 11.5 KB) into L2; the row-sized cp.asyncs then hit L2.
    if ((it_f & 3) == 0 && it_f + 4 < len) {
      const long long first = BWD ? 7 * dS : 4 * dS;        // lowest address of rows it_f + 4 .. it_f + 7 of the stream
      const char* base = reinterpret_cast<const char*>(pin - lane + first);
      const int nbytes = 4 * S * 4;
#pragma unroll
      for (int k = 0; k < PFL; ++k) {
        const int off = (lane + 32 * k) * 128;
        if (off < nbytes) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
          if (BWD) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + gamma_delta + off));
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    pin += dS;
  };
  stage(0);
  stage(1);
  stage(2);

  float xd = 0.f, ss = 0.f, cprev = 1.f;

  // acc[n] (+)= sum_in row[in] tap[(out - in) + D] for the run of 12 states that starts at state `run`
  auto sweep = [&](const float* row, int run, float* acc) {
    const float4* row4 = reinterpret_cast<const float4*>(row + run) + 3 * lane;     // window element 0 = state run + 12 l - D
    const float4* tap4 = reinterpret_cast<const float4*>(s_tap) + 1;
    float tq[4][4];                                    // taps of groups g, slot g & 3
#pragma unroll
    for (int g = D / 2; g <= D / 2 + 2; ++g) {         // iteration m needs groups D/2 - m - 1 .. D/2 - m + 2 and loads the lowest
      const float4 x = tap4[g];
      tq[g & 3][0] = x.x; tq[g & 3][1] = x.y; tq[g & 3][2] = x.z; tq[g & 3][3] = x.w;
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const float4 x = row4[m];
      const float wv[4] = {x.x, x.y, x.z, x.w};
      {
        const int g = D / 2 - m - 1;                   // the lowest tap group this iteration needs
        if (g >= -1) {
          const float4 y = tap4[g];
          tq[g & 3][0] = y.x; tq[g & 3][1] = y.y; tq[g & 3][2] = y.z; tq[g & 3][3] = y.w;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int n = 0; n < vNJ; ++n) {
          const int rr = n - (4 * m + k) + 2 * D;      // in = run + 12 l + 4 m + k - D, out = run + 12 l + n
          if (rr >= 0 && rr < W) acc[n] = fmaf(wv[k], tq[(rr >> 2) & 3][rr & 3], acc[n]);
        }
    }
  };

  for (int it = 0; it < len; ++it) {
    const int t = BWD ? len - 1 - it : it;
    asm volatile("cp.async.wait_group 2;" ::: "memory");
    __syncwarp();
    const int sb = it % vStages;
    const float* in = s_in + sb * KINDS * vMaxSW;
    const float ed = has_d ? s_ind[sb][0] : 0.f;
    const float sband = vwarp_sum(ss);     // (needed by the per-step scalars below; the two sweeps that follow are long enough
                                           // for the other clips of the scheduler to cover this chain)
    const float* rowp = s_row + ((it & 1) ^ 1) * ROW;
    float* rown = s_row + (it & 1) * ROW + D;
    // per-step scalars
    float inv = 0.f, invc = 0.f, invn = 0.f, add = 0.f, ct = 1.f;
    const bool first = !BWD && it == 0, last = BWD && it == 0;
    if constexpr (!BWD) {
      const float tot = sband + xd;
      if (lane == 0 && it > 0) pc[t - 1] = tot;
      inv = vrcp_pos(tot);
    } else {
      ct = s_ind[sb][2];
      invc = vrcp_pos(ct);
      invn = vrcp_pos(cprev);
      add = xd * r_in;
    }
    float s = 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int run = vMaxS * h;
      float acc[vNJ];
      const float acc0 = BWD ? 0.f : xd * q_out;
#pragma unroll
      for (int n = 0; n < vNJ; ++n) acc[n] = acc0;
      if (!first) sweep(rowp, run, acc);
      float v[vNJ], o[vNJ];
#pragma unroll
      for (int m = 0; m < vNJ / 4; ++m) {
        const float4 e4 = reinterpret_cast<const float4*>(in + run)[3 * lane + m];
        const float4 k4 = reinterpret_cast<const float4*>(s_kap + run)[3 * lane + m];
        const float e[4] = {e4.x, e4.y, e4.z, e4.w}, kp[4] = {k4.x, k4.y, k4.z, k4.w};
        float al[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (BWD) {
          const float4 a4 = reinterpret_cast<const float4*>(in + vMaxSW + run)[3 * lane + m];
          al[0] = a4.x; al[1] = a4.y; al[2] = a4.z; al[3] = a4.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int n = 4 * m + k;
          if constexpr (!BWD) {
            const int j = run + vNJ * lane + n;
            const float val = first ? (j < S ? pi[j] * e[k] : 0.f) : (acc[n] * inv) * e[k];
            v[n] = val;
            o[n] = val;
          } else {
            const float be = last ? 1.f : fmaf(kp[k], acc[n], add) * invn;
            o[n] = (al[k] * invc) * be;
            v[n] = e[k] * be;
          }
          s += v[n];
        }
        const float4 u = BWD ? make_float4(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3])
                             : make_float4(v[4 * m] * kp[0], v[4 * m + 1] * kp[1], v[4 * m + 2] * kp[2], v[4 * m + 3] * kp[3]);
        reinterpret_cast<float4*>(rown + run)[3 * lane + m] = u;
        reinterpret_cast<float4*>(s_out + run)[3 * lane + m] = make_float4(o[4 * m], o[4 * m + 1], o[4 * m + 2], o[4 * m + 3]);
      }
    }
    ss = s;
    float xdn, out_d;
    if constexpr (!BWD) {
      xdn = has_d ? (first ? pi[jd] * ed : (fmaf(sband, r_in, xd * a_uu) * inv) * ed) : 0.f;
      out_d = xdn;
    } else {
      const float bed = last ? 1.f : fmaf(sband, q_out, xd * a_uu) * invn;
      const float ald = has_d ? s_ind[sb][1] : 0.f;
      out_d = (ald * invc) * bed;
      xdn = has_d ? ed * bed : 0.f;
      cprev = ct;
    }
    xd = xdn;
    __syncwarp();
    stage(it + 3);
#pragma unroll
    for (int n = 0; n < NROWS; ++n)
      if (elem_ok(n)) st_global_cs_f32(pout + 32 * n, s_out[lane + 32 * n]);
    if (has_d && lane == 0) st_global_cs_f32(pout + jd, out_d);
    pout += dS;
  }
  if (!BWD) {
    const float sband = vwarp_sum(ss);
    if (lane == 0) pc[len - 1] = sband + xd;
  }
}

// vit_fb_banded.cu calls these
int fb_conv_detect(const float* A, int S, const vit_structure* st, int Dt, void* params, cudaStream_t stream) {
  const char* e = getenv("VIT_FB_CONV");
  const int enable = !(e && e[0] == '0');
  fb_conv_detect_kernel<<<1, vMaxS, 0, stream>>>(A, S, st->dense_index, st->halfwidth, Dt, enable, (ConvParams*)params);
  note_launch();
  VIT_CUDA_TRY(cudaGetLastError());
  return VIT_OK;
}

static int fb_conv_template_wide(int d) {
  const int opts[] = {20, 28, 40, 56};
  for (int o : opts) if (d <= o) return o;
  return -1;
}
// the kernel instance's half-width for this shape (taps are stored for it), or -1
int fb_conv_template_D(int S, int d) {
  if (S <= vMaxS && d <= 14) {
    const int opts[] = {4, 8, 12, 14};
    for (int o : opts) if (d <= o) return o;
  }
  if (S <= vMaxSW) return fb_conv_template_wide(d);
  return -1;
}

int fb_conv_passes(int D, const void* params, const float* pi, const float* lik, const int32_t* lengths, int B, int T_max,
                   int S, int jd, float* gamma, float* cnorm, cudaStream_t stream) {
  const ConvParams* prm = (const ConvParams*)params;
  // NF: the element-row specialisation (see the kernels).  Instances exist for the layouts the reference's state sets
  // have -- band states [0, Sv) with Sv in (288, 320] (dcnet: 320 bins), (352, 384] (tonet / ftanet: 360), (704, 736]
  // (jdc / imm: 721) -- else generic.
  const int Sv = jd >= 0 ? S - 1 : S;
  const bool tail_dense = jd < 0 || jd == S - 1;
  const int nf = (tail_dense && Sv > 0) ? (Sv - 1) / 32 : -1;
  if (D > 14) {
    // wide instance: two runs of 12 states per lane, taps and kappa in shared memory
    const size_t fl = (size_t)2 * (vMaxSW + 2 * D) + vMaxSW + vMaxSW + (size_t)(D / 2 + 4) * 4;
    const size_t smem_f = (fl + (size_t)vStages * 1 * vMaxSW) * sizeof(float);
    const size_t smem_b = (fl + (size_t)vStages * 2 * vMaxSW) * sizeof(float);
#define VIT_FBW_LAUNCH(DD, NFF)                                                                                       \
  do {                                                                                                                \
    fb_convw_pass_kernel<DD, false, NFF><<<B, 32, smem_f, stream>>>(prm, pi, lik, lengths, T_max, S, jd, gamma, cnorm); \
    fb_convw_pass_kernel<DD, true, NFF><<<B, 32, smem_b, stream>>>(prm, pi, lik, lengths, T_max, S, jd, gamma, cnorm);  \
  } while (0)
#define VIT_FBW_CASE(DD)                                                                                              \
  case DD:                                                                                                            \
    if (nf == 22) VIT_FBW_LAUNCH(DD, 22);                                                                             \
    else VIT_FBW_LAUNCH(DD, -1);                                                                                      \
    break;
    switch (D) {
      VIT_FBW_CASE(20) VIT_FBW_CASE(28) VIT_FBW_CASE(40) VIT_FBW_CASE(56)
      default: return VIT_ERR_UNSUPPORTED_ALGO;
    }
#undef VIT_FBW_CASE
#undef VIT_FBW_LAUNCH
    note_launch(2);
    VIT_CUDA_TRY(cudaGetLastError());
    return VIT_OK;
  }
  const size_t smem_f = (size_t)(2 * vRow + vStages * 1 * vMaxS + vMaxS) * sizeof(float);
  const size_t smem_b = (size_t)(2 * vRow + vStages * 2 * vMaxS + vMaxS) * sizeof(float);
#define VIT_FBC_LAUNCH(DD, NFF)                                                                                      \
  do {                                                                                                               \
    fb_conv_pass_kernel<DD, false, NFF><<<B, 32, smem_f, stream>>>(prm, pi, lik, lengths, T_max, S, jd, gamma, cnorm); \
    fb_conv_pass_kernel<DD, true, NFF><<<B, 32, smem_b, stream>>>(prm, pi, lik, lengths, T_max, S, jd, gamma, cnorm);  \
  } while (0)
#define VIT_FBC_CASE(DD)                                                                                             \
  case DD:                                                                                                           \
    if (nf == 11) VIT_FBC_LAUNCH(DD, 11);                                                                            \
    else if (nf == 9) VIT_FBC_LAUNCH(DD, 9);                                                                         \
    else VIT_FBC_LAUNCH(DD, -1);                                                                                     \
    break;
  switch (D) {
    VIT_FBC_CASE(4) VIT_FBC_CASE(8) VIT_FBC_CASE(12) VIT_FBC_CASE(14)
    default: return VIT_ERR_UNSUPPORTED_ALGO;
  }
#undef VIT_FBC_LAUNCH
#undef VIT_FBC_CASE
  note_launch(2);
  VIT_CUDA_TRY(cudaGetLastError());
  return VIT_OK;
}

}  // namespace vit
