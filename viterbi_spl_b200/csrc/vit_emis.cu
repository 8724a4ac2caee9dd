// The step BEFORE the hot path and the step AFTER it, batched on the GPU so that
// emissions never round-trip the host:
//
//   vit_emissions_f32   acoustic-model logits -> HMM emission table [B][T][S] (probability or log(p + tiny) domain)
//       VIT_EMIS_SOFTMAX  SoftMaxViterbi.observation_probs_fn   dcnet/softmax_viterbi.py:2508-2579 (7 copies)
//       VIT_EMIS_SHAUN    Viterbi.observation_probs_fn          tonet/softmax_priors.py:1722-1786 (max-subtracted
//                         form; dcnet/softmax_viterbi.py:2316-2359 is the same model)
//   vit_voiced_bins     states -> (voiced, bins)                dcnet/softmax_viterbi.py:2427-2431
//   vit_melody_stats_f32  (bins, voiced, logits, ref notes) -> refined note per frame + the nine frame counters of
//                         MetricsInference.viterbi_update_states_tf_fn, dcnet/softmax_viterbi.py:2923-2979
//                         (est_notes_fn: dcnet/main.py:1911-1934)
//
// Peak picking is exact (comparisons only): bin k is a peak iff np.argmax of the reflect-padded window
// [k - spw, k + spw] returns the centre, i.e. every left neighbour is STRICTLY smaller and every right neighbour is
// smaller or equal (first maximum wins; find_peaks_all_at_once_np_fn, dcnet/softmax_viterbi.py:2508-2528).
// The softmax / logistic / log values use expf / logf, which differ from NumPy's SIMD float32 routines in the last
// ulp: this path is tolerance-checked (1e-5 relative) and is NOT part of the bit-exact claim of the decoder -- the
// reference-facing wrappers keep taking the log on the host.
//
// HBM-bound byte work: 4 B read + 4 B written per state-frame.  One warp per frame; the frame's logits are staged in
// shared memory for the +-spw neighbourhood tests; loads and stores are coalesced 128-byte lines.
#include <cstdint>
#include <cstdlib>

#include "vit_common.cuh"

namespace vit {

constexpr int kEmisWarps = 8;
// floats per shared-memory row of the generic kernel: reflect padding of spw either side, rounded up
__host__ __device__ inline int emis_row_floats(int n_bins, int spw) { return (n_bins + 2 * spw + 8 + 3) & ~3; }
constexpr float kTinyF = 1.1754943508222875e-38f;   // np.finfo(np.float32).tiny

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// x: the frame's n_bins voiced logits in shared memory.  np.pad(mode='reflect'): index -m -> m, n-1+m -> n-1-m.
__device__ __forceinline__ bool is_peak(const float* x, int n, int k, int spw) {
  const float c = x[k];
  bool ok = true;
  if (k >= spw && k + spw < n) {                                     // interior: no reflection
    for (int m = 1; m <= spw; ++m) ok = ok && (x[k - m] < c) && (x[k + m] <= c);
  } else {
    for (int m = 1; m <= spw; ++m) {
      int l = k - m, r = k + m;
      if (l < 0) l = -l;
      if (r > n - 1) r = 2 * (n - 1) - r;
      ok = ok && (x[l] < c) && (x[r] <= c);
    }
  }
  return ok;
}

// model 0 (SOFTMAX): logits [B][T][1 + n_bins], column 0 = unvoiced; prior [1 + n_bins] in the SAME (unvoiced-first)
//                    order as np.roll(ini_probs, 1) (dcnet/softmax_viterbi.py:2534-2538), or NULL for "unscaled".
// model 1 (SHAUN)  : logits [B][T][n_bins]; threshold = logit(voicing threshold); p = 0.8, scale = 2 as in the reference.
// out [B][T][n_bins + 1], unvoiced LAST (the np.roll(-1) at :2577; the shaun model writes it there directly).
// FAST (any single_side_peak_width >= 2: 5 dcnet / msnet / ftanet / tonet, 15 tonet/for_paper.py, 16 jdc, 20 imm): the
// frame is stored with its reflect padding materialised and window maxima of width 2, 4, ... w2 (the largest power of two
// <= spw) are built by doubling, ping-ponging between two rows; a window of width spw is then two overlapping windows of
// width w2:
//   left  max = max(m[k - spw], m[k - w2]),   right max = max(m[k + 1], m[k + 1 + spw - w2]),   m[j] = max(x[j .. j + w2 - 1])
// -- 3 log2(w2) + 5 shared-memory accesses per bin instead of 2 spw + 1 (41 at spw = 20).  spw < 2: the direct scan.
template <int MODEL, bool FAST, int kPre>
__global__ void __launch_bounds__(32 * kEmisWarps)
emissions_kernel(const float* __restrict__ logits, const float* __restrict__ prior, long long n_frames, int n_bins, int spw,
                 float threshold, int out_log, int pk_cap, float* __restrict__ out) {
  // per warp: x and two window-maxima rows of np floats, then the compacted peak list (bin, exp) of pk_cap entries --
  // peaks are more than spw bins apart, so pk_cap = n_bins / (spw + 1) + 2 suffices (5 instead of 3 blocks per SM at 360 bins)
  extern __shared__ float s_x[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int pad = FAST ? spw : 0;
  const int np = emis_row_floats(n_bins, spw);                       // padded row: spw + n_bins + spw, rounded up
  const int per_warp = 3 * np + 2 * pk_cap;
  float* x = s_x + (size_t)w * per_warp + pad;                       // x[k] = logit of bin k; x[-spw .. n+spw-1] valid
  float* ma = s_x + (size_t)w * per_warp + np + pad;
  float* mb = s_x + (size_t)w * per_warp + 2 * np + pad;
  int w2 = 1, levels = 0;                                            // largest power of two <= spw
  while (2 * w2 <= spw) { w2 *= 2; ++levels; }
  const float* m = (levels & 1) ? ma : mb;                           // the row that ends up holding the width-w2 maxima
  int* pk_idx = reinterpret_cast<int*>(s_x + (size_t)w * per_warp + 3 * np);   // compacted peak bins
  float* pk_e = s_x + (size_t)w * per_warp + 3 * np + pk_cap;        // exp(peak logit - max)
  const int n_in = MODEL == 0 ? n_bins + 1 : n_bins;
  const int S = n_bins + 1;
  const float zero_out = out_log ? logf(kTinyF) : 0.f;               // log(0 + tiny) = -87.33655
  // software pipeline over frames: the next frame's logits are fetched into registers (n_bins <= 384) before this frame
  // is processed, so the HBM latency hides under ~700 instructions of work instead of stalling every frame
  // (kPre = 12: up to 384 bins; 24: up to 768 -- jdc / imm have 721)
  const bool use_pre = n_bins <= 32 * kPre;
  const long long f_stride = (long long)gridDim.x * kEmisWarps;
  float pre[kPre];
  float pre_unv = 0.f;
  {
    const long long f0 = (long long)blockIdx.x * kEmisWarps + w;
    if (use_pre && f0 < n_frames) {
      const float* in0 = logits + f0 * n_in;
#pragma unroll
      for (int i = 0; i < kPre; ++i) pre[i] = (lane + 32 * i < n_bins) ? ld_global_nc_f32(in0 + (MODEL == 0 ? 1 : 0) + lane + 32 * i) : 0.f;
      if (MODEL == 0) pre_unv = ld_global_nc_f32(in0);
    }
  }
  for (long long f = (long long)blockIdx.x * kEmisWarps + w; f < n_frames; f += f_stride) {
    const float* in = logits + f * n_in;
    float* o = out + f * S;
    const float* vin = in + (MODEL == 0 ? 1 : 0);
    float unv_now = 0.f;
    if (use_pre) {
#pragma unroll
      for (int i = 0; i < kPre; ++i) if (lane + 32 * i < n_bins) x[lane + 32 * i] = pre[i];
      unv_now = pre_unv;
      if (f + f_stride < n_frames) {
        const float* in1 = logits + (f + f_stride) * n_in;
#pragma unroll
        for (int i = 0; i < kPre; ++i) pre[i] = (lane + 32 * i < n_bins) ? ld_global_nc_f32(in1 + (MODEL == 0 ? 1 : 0) + lane + 32 * i) : 0.f;
        if (MODEL == 0) pre_unv = ld_global_nc_f32(in1);
      }
    } else {
      for (int k = lane; k < n_bins; k += 32) x[k] = vin[k];
      if (MODEL == 0) unv_now = in[0];
    }
    if (FAST) {
      __syncwarp();
      if (lane < spw) {                                              // np.pad(mode='reflect'): -m -> m, n-1+m -> n-1-m
        x[-1 - lane] = x[1 + lane];
        x[n_bins + lane] = x[n_bins - 2 - lane];
      }
      __syncwarp();
      const float* src = x;
      for (int l = 0, h = 1; l < levels; ++l, h *= 2) {              // dst[j] = max of x[j .. j + 2h - 1]
        float* dst = (l & 1) ? mb : ma;
        for (int k = lane - spw; k <= n_bins + spw - 2 * h; k += 32) dst[k] = fmaxf(src[k], src[k + h]);
        __syncwarp();
        src = dst;
      }
    }
    __syncwarp();
    // pass 1: find the peaks and COMPACT them (ballot + prefix count) into a per-warp list, so that the expensive
    // exp / log below run once per peak on consecutive lanes instead of once per bin under divergence
    int cnt = 0;                                                     // number of voiced peaks (warp-uniform)
    for (int k = lane; k < ((n_bins + 31) & ~31); k += 32) {
      bool pk = false;
      if (k < n_bins) {
        if (FAST) {
          const float c = x[k];
          pk = (c > fmaxf(m[k - spw], m[k - w2])) && (c >= fmaxf(m[k + 1], m[k + 1 + spw - w2]));
        } else {
          pk = is_peak(x, n_bins, k, spw);
        }
      }
      const uint32_t bal = __ballot_sync(0xffffffffu, pk);
      if (pk) pk_idx[cnt + __popc(bal & ((1u << lane) - 1u))] = k;
      cnt += __popc(bal);
    }
    __syncwarp();
    const int n_peaks = cnt;
    float mx = -INFINITY;
    for (int q = lane; q < cnt; q += 32) mx = fmaxf(mx, x[pk_idx[q]]);
    float unv_logit = 0.f;
    if (MODEL == 0) {
      unv_logit = unv_now;                                           // column 0 is always a peak (:2521)
      mx = fmaxf(mx, unv_logit);
    }
    mx = warp_max(mx);
    // pass 2: exp(peak - max), kept for pass 3, and their sum
    float sum = 0.f;
    for (int q = lane; q < cnt; q += 32) {
      const float e = expf(x[pk_idx[q]] - mx);
      pk_e[q] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    float unv_out;
    float scale;                                                     // value of a voiced peak = exp(x - mx) * scale / prior
    if (MODEL == 0) {
      sum += expf(unv_logit - mx);
      scale = 0.f;
      const float p0 = prior ? prior[0] : 1.f;
      unv_out = (n_peaks == 0) ? 1.f / p0 : expf(unv_logit - mx) / sum / p0;      // lone unvoiced peak: 1 / prior (:2560-2563)
    } else {
      if (n_peaks == 0) {
        scale = 0.f;
        unv_out = 1.f;                                               // no peak: E[unvoiced] = 1 (tonet :1758-1760)
      } else {
        const float offset = logf(0.8f / (1.f - 0.8f));
        const float sg = 2.f * (mx - threshold) + (mx >= threshold ? offset : -offset);   // (:1766-1769)
        // expit (:1711-1720); 1 - p_voiced is formed directly (the reference subtracts in float64)
        float pv, qv;
        if (sg > 0.f) { const float t = expf(-sg); pv = 1.f / (1.f + t); qv = t / (1.f + t); }
        else { const float t = expf(sg); pv = t / (1.f + t); qv = 1.f / (1.f + t); }
        scale = pv / sum;
        unv_out = qv;
      }
    }
    // pass 3: the output row is assembled in shared memory (x is dead now): constant fill, peaks scattered over it,
    // then one coalesced copy.  Non-peaks are exactly 0 -> log(tiny).
    __syncwarp();
    for (int k = lane; k < n_bins; k += 32) x[k] = zero_out;
    __syncwarp();
    for (int q = lane; q < cnt; q += 32) {
      const int k = pk_idx[q];
      float p = pk_e[q];
      if (MODEL == 0) {
        p = p / sum;                                                 // np.divide(peak_logits, t) then / priors (:2568-2572)
        if (prior) p = p / prior[k + 1];
      } else {
        p = p * scale;                                               // t = p_voiced / sum; peak_logits * t (:1777-1778)
      }
      x[k] = out_log ? logf(p + kTinyF) : p;
    }
    __syncwarp();
    for (int k = lane; k < n_bins; k += 32) o[k] = x[k];
    if (lane == 0) o[n_bins] = out_log ? logf(unv_out + kTinyF) : unv_out;
    __syncwarp();
  }
}

// ---- register-window instance (single_side_peak_width == 5, 7 <= n_bins <= 384, 16-byte aligned output) ------------
// The instance the reference's models take (dcnet / msnet / ftanet / tonet: spw = 5, 320-721 bins; 360 here).  The
// generic kernel above spends ~700 instructions per frame, mostly shared-memory round trips (window maxima, peak
// compaction, row assembly) -- 31 % of HBM bandwidth.  Here a lane owns 12 CONSECUTIVE bins:
//   * the frame's logits arrive by cp.async (LDGSTS, coalesced 128-byte lines) into a 3-deep ring of rows, two frames
//     ahead of the one being processed -- no registers, no stall;
//   * a lane reads its 12 bins + 5 neighbours either side as 7 conflict-free LDS.128 (lane stride 48 B) and does the
//     whole peak test in registers: pair maxima once, then left / right = one 3-input maximum each;
//   * peaks are > 5 bins apart, so bins [0,6) and [6,12) of a lane hold at most one each: no compaction, the lane that
//     finds a peak also evaluates its exp / log;
//   * the output row is a constant (log(tiny)) except at the peaks: filled in shared memory as float4s, peaks
//     overwritten, then copied with STG.128 -- the row is staged at the shift (f * S) & 3 so that shared-memory float4s
//     and 16-byte aligned global float4s coincide (rows of S = 361 floats are not aligned themselves).
// Same comparisons as emissions_kernel<., true> (exact peaks); the softmax sum is taken in a different order (values
// differ in the last ulp from either the other kernel or NumPy; tests hold 1e-5 relative).
constexpr int kRegPer = 12;                          // bins per lane
constexpr int kRegRow = 400;                         // staged input row: x[k] at index k + 8, k in [-8, 392)
constexpr int kRegOut = 392;                         // staged output row: shift (< 4) + 385, as float4s
constexpr int kRegStages = 3;
constexpr int kRegWarpFloats = kRegStages * kRegRow + kRegOut;

__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float warp_max_redux_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

template <int MODEL>
__global__ void __launch_bounds__(32 * kEmisWarps, 4)
emissions_reg_kernel(const float* __restrict__ logits, const float* __restrict__ prior, long long n_frames, int n_bins,
                     float threshold, int out_log, float* __restrict__ out) {
  extern __shared__ __align__(16) float s_x[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* base = s_x + (size_t)w * kRegWarpFloats;
  float* orow = base + kRegStages * kRegRow;
  const int n_in = MODEL == 0 ? n_bins + 1 : n_bins;
  const int S = n_bins + 1;
  const float zero_out = out_log ? logf(kTinyF) : 0.f;               // log(0 + tiny) = -87.33655
  const long long f_stride = (long long)gridDim.x * kEmisWarps;
  const long long f0 = (long long)blockIdx.x * kEmisWarps + w;
  for (int i = lane; i < kRegStages * kRegRow; i += 32) base[i] = 0.f;   // lanes past n_bins read these (results unused)
  // the output row is the constant log(tiny) except at the peaks: filled ONCE; every frame writes its peaks, copies the
  // row out and puts the constant back (3 scalar stores instead of a 1.5 KB refill)
  for (int i = lane; i < kRegOut; i += 32) orow[i] = zero_out;
  __syncwarp();
  // frame -> ring stage st: voiced logits to x[0 .. n_bins), the unvoiced logit (model 0, column 0) to slot 0 of the
  // row (k = -8: outside every window).  A group is committed even when empty so that the wait count stays uniform.
  // Running pointers (one 64-bit add per frame); rows 0 .. n_full-1 of the 12 need no bound check.
  const int n_full = n_bins >> 5;
  const size_t in_step = (size_t)f_stride * n_in;
  const float* in_fetch = logits + f0 * n_in + (MODEL == 0 ? 1 : 0) + lane;    // my first voiced logit of the next frame to fetch
  long long f_fetch = f0;
  auto fetch = [&](int st) {
    if (f_fetch < n_frames) {
      const uint32_t dst = smem_u32(base + st * kRegRow + 8 + lane);
      if (n_full >= kRegPer - 1) {                                   // 352 <= n_bins: the reference's 360-bin models
#pragma unroll
        for (int i = 0; i < kRegPer - 1; ++i) cp_async4(dst + 128 * i, in_fetch + 32 * i);
        if (lane + 32 * (kRegPer - 1) < n_bins) cp_async4(dst + 128 * (kRegPer - 1), in_fetch + 32 * (kRegPer - 1));
      } else {
#pragma unroll
        for (int i = 0; i < kRegPer; ++i)
          if (lane + 32 * i < n_bins) cp_async4(dst + 128 * i, in_fetch + 32 * i);
      }
      if (MODEL == 0 && lane == 0) cp_async4(dst - 32, in_fetch - 1);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    in_fetch += in_step;
    f_fetch += f_stride;
  };
#pragma unroll
  for (int s = 0; s < kRegStages - 1; ++s) fetch(s);
  int st = 0;
  // alignment shift of the frame's output row, (f * S) & 3, kept incrementally
  int a = (int)((f0 * S) & 3);
  const int a_step = (int)((f_stride * S) & 3);
  float* o_f = out + f0 * S;
  const size_t out_step = (size_t)f_stride * S;
  const int kb = kRegPer * lane;
  // the unvoiced state is evaluated in the second slot of lane 31 (bins 378 .. 383: never a peak when n_bins <= 378)
  const bool unv_in_slot = n_bins <= 32 * kRegPer - 6;
  const float prior0 = (MODEL == 0 && prior) ? prior[0] : 1.f;
  for (long long f = f0; f < n_frames; f += f_stride) {
    {
      int stn = st + kRegStages - 1;
      if (stn >= kRegStages) stn -= kRegStages;
      fetch(stn);
    }
    asm volatile("cp.async.wait_group %0;" ::"n"(kRegStages - 1) : "memory");
    __syncwarp();
    float* xr = base + st * kRegRow;
    float* x = xr + 8;
    if (lane < 5) {                                                  // np.pad(mode='reflect'): -m -> m, n-1+m -> n-1-m
      x[-1 - lane] = x[1 + lane];
      x[n_bins + lane] = x[n_bins - 2 - lane];
    }
    __syncwarp();
    // my window: wv[j] = x[12 lane - 8 + j]; bin i of mine is wv[8 + i]; wv[3 .. 24] are used
    float wv[28];
    {
      const float4* w4 = reinterpret_cast<const float4*>(xr) + 3 * lane;
#pragma unroll
      for (int m = 0; m < 7; ++m) {
        const float4 v = w4[m];
        wv[4 * m] = v.x; wv[4 * m + 1] = v.y; wv[4 * m + 2] = v.z; wv[4 * m + 3] = v.w;
      }
    }
    const float unv_logit = MODEL == 0 ? xr[0] : 0.f;                // column 0 is always a peak (:2521)
    float m2[24];
#pragma unroll
    for (int j = 3; j <= 22; ++j) m2[j] = fmaxf(wv[j], wv[j + 1]);
    // bin k is a peak iff x[k-5 .. k-1] < x[k] and x[k+1 .. k+5] <= x[k] (first maximum of the window is the centre)
    float c0 = -INFINITY, c1 = -INFINITY;
    int i0 = -1, i1 = -1;
#pragma unroll
    for (int i = 0; i < kRegPer; ++i) {
      const int q = 8 + i;
      const float c = wv[q];
      const float left = fmax3(m2[q - 5], m2[q - 3], wv[q - 1]);
      const float right = fmax3(m2[q + 1], m2[q + 3], wv[q + 5]);
      const bool pk = (kb + i < n_bins) && (c > left) && (c >= right);
      if (i < 6) { c0 = pk ? c : c0; i0 = pk ? i : i0; }
      else { c1 = pk ? c : c1; i1 = pk ? i : i1; }
    }
    const bool v0 = i0 >= 0, v1 = i1 >= 0;
    const int n_peaks = __reduce_add_sync(0xffffffffu, (v0 ? 1 : 0) + (v1 ? 1 : 0));
    float mx = fmaxf(c0, c1);
    if (MODEL == 0) mx = fmaxf(mx, unv_logit);
    mx = warp_max_redux_f32(mx);
    const float e0 = v0 ? expf(c0 - mx) : 0.f;
    const float e1 = v1 ? expf(c1 - mx) : 0.f;
    float sum = warp_sum(e0 + e1);
    // value of a voiced peak = e * scale (/ prior); p_unv = the unvoiced state's value
    float scale, p_unv;
    if (MODEL == 0) {
      // softmax over the peaks and column 0; a lone unvoiced peak gives e = sum = 1: 1 / prior (:2560-2563)
      const float eu = expf(unv_logit - mx);
      sum += eu;
      scale = __fdividef(1.f, sum);                                  // np.divide(peak_logits, t) (:2568)
      p_unv = __fdividef(eu * scale, prior0);
    } else {
      if (n_peaks == 0) {
        scale = 0.f;
        p_unv = 1.f;                                                 // no peak: E[unvoiced] = 1 (tonet :1758-1760)
      } else {
        const float offset = logf(0.8f / (1.f - 0.8f));
        const float sg = 2.f * (mx - threshold) + (mx >= threshold ? offset : -offset);   // (:1766-1769)
        float pv, qv;                                                // expit (:1711-1720)
        if (sg > 0.f) { const float t = expf(-sg); pv = 1.f / (1.f + t); qv = t / (1.f + t); }
        else { const float t = expf(sg); pv = t / (1.f + t); qv = 1.f / (1.f + t); }
        scale = __fdividef(pv, sum);                                 // t = p_voiced / sum; peak_logits * t (:1777-1778)
        p_unv = qv;
      }
    }
    // two value slots per lane, evaluated without divergence: my peak in bins [0,6), my peak in bins [6,12) -- or, on
    // lane 31, the unvoiced state
    const bool s1_unv = unv_in_slot && lane == 31;
    float p0v = e0 * scale, p1v = e1 * scale;
    if (MODEL == 0 && prior) {                                       // then / priors (:2571-2572)
      p0v = __fdividef(p0v, prior[v0 ? kb + i0 + 1 : 0]);
      p1v = __fdividef(p1v, prior[v1 ? kb + i1 + 1 : 0]);
    }
    if (s1_unv) p1v = p_unv;
    if (out_log) {
      p0v = logf(p0v + kTinyF);
      p1v = logf(p1v + kTinyF);
    }
    // output row at shift a: element k of the row sits at orow[a + k], so float4 j of orow is the 16-byte aligned
    // global float4 j of (out + f * S - a)
    const int total = a + S;
    const int k1 = a + (s1_unv ? n_bins : kb + i1);
    if (v0) orow[a + kb + i0] = p0v;
    if (v1 || s1_unv) orow[k1] = p1v;
    if (!unv_in_slot && lane == 0) orow[a + n_bins] = out_log ? logf(p_unv + kTinyF) : p_unv;
    __syncwarp();
    {
      float* oa = o_f - a;                                           // 16-byte aligned
      const int j0 = a ? 1 : 0, j1 = total >> 2;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = lane + 32 * i;
        if (j >= j0 && j < j1) reinterpret_cast<float4*>(oa)[j] = reinterpret_cast<const float4*>(orow)[j];
      }
      if (a && lane >= a && lane < 4) oa[lane] = orow[lane];         // head: the rest of the first float4
      const int tl = 4 * j1 + lane;
      if (lane < 4 && tl < total) oa[tl] = orow[tl];                 // tail
    }
    __syncwarp();                                                    // the row has been copied: constant back in place
    if (v0) orow[a + kb + i0] = zero_out;
    if (v1 || s1_unv) orow[k1] = zero_out;
    if (!unv_in_slot && lane == 0) orow[a + n_bins] = zero_out;
    __syncwarp();                                                    // orow and the stage are free again
    if (++st == kRegStages) st = 0;
    a = (a + a_step) & 3;
    o_f += out_step;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- wide register-window instance (single_side_peak_width 16 or 20, 24 < n_bins <= 756: jdc / imm, 721 bins) ----------
// Same scheme as emissions_reg_kernel, for the fine-grid models (jdc/viterbi_softmax.py: 721 bins, half-width 16;
// imm/main_imm.py:141-234: 721 bins, half-width 20): the frame's logits arrive by cp.async into a 3-deep ring of rows, two
// frames ahead; a lane owns 12 consecutive bins in each of TWO passes over the row (bins [0, 384) and [384, 768)) and reads
// them with their +-SPW neighbours as 11 / 13 conflict-free LDS.128 (lane stride 48 B).  The peak test stays in registers:
// window maxima of width 3 and 9 by 3-input maxima (FMNMX3), then a window of 16 = two overlapping 9s, of 20 = three; peaks
// are more than SPW >= 12 bins apart, so a lane finds at most ONE per pass -- the two value slots of the narrow kernel are
// the two passes here, and lane 31's second slot (bins 756 ..: past n_bins) carries the unvoiced state.  Exact peaks; values
// as in the other kernels (1e-5 relative to NumPy).
constexpr int kWPer = 12;                            // bins per lane and pass
constexpr int kWPass = 2;
constexpr int kWBins = 32 * kWPer * kWPass;          // 768 bin slots
constexpr int kWStages = 3;
template <int SPW>
struct WideEmis {
  static_assert(SPW % 4 == 0 && SPW >= kWPer, "window alignment / one peak per pass");
  static constexpr int XO = SPW + 4;                 // x[k] lives at row index k + XO; index 0 holds the unvoiced logit
  static constexpr int ROW = (XO + kWBins + SPW + 7) & ~3;
  static constexpr int NW = kWPer + 2 * SPW;         // window floats per lane and pass: 44 / 52
  static constexpr int OUT = (3 + kWBins + 1 + 7) & ~3;
  static constexpr int WARP_FLOATS = kWStages * ROW + OUT;
};

template <int MODEL, int SPW>
__global__ void __launch_bounds__(32 * kEmisWarps, 2)
emissions_wide_kernel(const float* __restrict__ logits, const float* __restrict__ prior, long long n_frames, int n_bins,
                      float threshold, int out_log, float* __restrict__ out) {
  using W = WideEmis<SPW>;
  extern __shared__ __align__(16) float s_x[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* base = s_x + (size_t)w * W::WARP_FLOATS;
  float* orow = base + kWStages * W::ROW;
  const int n_in = MODEL == 0 ? n_bins + 1 : n_bins;
  const int S = n_bins + 1;
  const float zero_out = out_log ? logf(kTinyF) : 0.f;               // log(0 + tiny) = -87.33655
  const long long f_stride = (long long)gridDim.x * kEmisWarps;
  const long long f0 = (long long)blockIdx.x * kEmisWarps + w;
  for (int i = lane; i < kWStages * W::ROW; i += 32) base[i] = 0.f;  // lanes past n_bins read these (results unused)
  for (int i = lane; i < W::OUT; i += 32) orow[i] = zero_out;        // the output row: constant except at the peaks
  __syncwarp();
  const size_t in_step = (size_t)f_stride * n_in;
  const float* in_fetch = logits + f0 * n_in + (MODEL == 0 ? 1 : 0) + lane;
  long long f_fetch = f0;
  auto fetch = [&](int st) {
    if (f_fetch < n_frames) {
      const uint32_t dst = smem_u32(base + st * W::ROW + W::XO + lane);
#pragma unroll
      for (int i = 0; i < kWBins / 32; ++i)
        if (lane + 32 * i < n_bins) cp_async4(dst + 128 * i, in_fetch + 32 * i);
      if (MODEL == 0 && lane == 0) cp_async4(dst - 4 * W::XO, in_fetch - 1);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    in_fetch += in_step;
    f_fetch += f_stride;
  };
#pragma unroll
  for (int s = 0; s < kWStages - 1; ++s) fetch(s);
  int st = 0;
  int a = (int)((f0 * S) & 3);                                       // alignment shift of the frame's output row
  const int a_step = (int)((f_stride * S) & 3);
  float* o_f = out + f0 * S;
  const size_t out_step = (size_t)f_stride * S;
  const bool unv_in_slot = n_bins <= kWBins - kWPer;                 // lane 31's pass-1 bins lie past n_bins
  const float prior0 = (MODEL == 0 && prior) ? prior[0] : 1.f;
  for (long long f = f0; f < n_frames; f += f_stride) {
    {
      int stn = st + kWStages - 1;
      if (stn >= kWStages) stn -= kWStages;
      fetch(stn);
    }
    asm volatile("cp.async.wait_group %0;" ::"n"(kWStages - 1) : "memory");
    __syncwarp();
    float* xr = base + st * W::ROW;
    float* x = xr + W::XO;
    if (lane < SPW) {                                                // np.pad(mode='reflect'): -m -> m, n-1+m -> n-1-m
      x[-1 - lane] = x[1 + lane];
      x[n_bins + lane] = x[n_bins - 2 - lane];
    }
    __syncwarp();
    const float unv_logit = MODEL == 0 ? xr[0] : 0.f;                // column 0 is always a peak (:2521)
    float cpk[kWPass];                                               // my peak's logit per pass (-inf: none)
    int ipk[kWPass];                                                 // ... and its bin
#pragma unroll
    for (int p = 0; p < kWPass; ++p) {
      const int kb = 32 * kWPer * p + kWPer * lane;                  // my first bin of this pass
      // window: wv[j] = x[kb - SPW + j]; bin i of mine is wv[SPW + i]
      float wv[W::NW];
      {
        const float4* w4 = reinterpret_cast<const float4*>(xr) + (kb + 4) / 4;
#pragma unroll
        for (int m = 0; m < W::NW / 4; ++m) {
          const float4 v = w4[m];
          wv[4 * m] = v.x; wv[4 * m + 1] = v.y; wv[4 * m + 2] = v.z; wv[4 * m + 3] = v.w;
        }
      }
      float m3[W::NW - 2], m9[W::NW - 8];
#pragma unroll
      for (int j = 0; j < W::NW - 2; ++j) m3[j] = fmax3(wv[j], wv[j + 1], wv[j + 2]);
#pragma unroll
      for (int j = 0; j < W::NW - 8; ++j) m9[j] = fmax3(m3[j], m3[j + 3], m3[j + 6]);
      float c_p = -INFINITY;
      int i_p = -1;
#pragma unroll
      for (int i = 0; i < kWPer; ++i) {
        const float c = wv[SPW + i];
        // bin k is a peak iff x[k-SPW .. k-1] < x[k] and x[k+1 .. k+SPW] <= x[k]; a window of SPW starting at s:
        const int sl = i, sr = SPW + i + 1;
        float left, right;
        if constexpr (SPW == 16) {
          left = fmaxf(m9[sl], m9[sl + 7]);
          right = fmaxf(m9[sr], m9[sr + 7]);
        } else {
          left = fmax3(m9[sl], m9[sl + 9], m9[sl + SPW - 9]);
          right = fmax3(m9[sr], m9[sr + 9], m9[sr + SPW - 9]);
        }
        const bool pk = (kb + i < n_bins) && (c > left) && (c >= right);
        c_p = pk ? c : c_p;
        i_p = pk ? kb + i : i_p;
      }
      cpk[p] = c_p;
      ipk[p] = i_p;
    }
    const bool v0 = ipk[0] >= 0, v1 = ipk[1] >= 0;
    const int n_peaks = __reduce_add_sync(0xffffffffu, (v0 ? 1 : 0) + (v1 ? 1 : 0));
    float mx = fmaxf(cpk[0], cpk[1]);
    if (MODEL == 0) mx = fmaxf(mx, unv_logit);
    mx = warp_max_redux_f32(mx);
    const float e0 = v0 ? expf(cpk[0] - mx) : 0.f;
    const float e1 = v1 ? expf(cpk[1] - mx) : 0.f;
    float sum = warp_sum(e0 + e1);
    float scale, p_unv;
    if (MODEL == 0) {
      const float eu = expf(unv_logit - mx);                         // softmax over the peaks and column 0 (:2565-2572)
      sum += eu;
      scale = __fdividef(1.f, sum);
      p_unv = __fdividef(eu * scale, prior0);
    } else {
      if (n_peaks == 0) {
        scale = 0.f;
        p_unv = 1.f;                                                 // no peak: E[unvoiced] = 1 (imm/main_imm.py:203-205)
      } else {
        const float offset = logf(0.8f / (1.f - 0.8f));
        const float sg = 2.f * (mx - threshold) + (mx >= threshold ? offset : -offset);   // (imm/main_imm.py:211-215)
        float pv, qv;                                                // expit (:155-164)
        if (sg > 0.f) { const float t = expf(-sg); pv = 1.f / (1.f + t); qv = t / (1.f + t); }
        else { const float t = expf(sg); pv = t / (1.f + t); qv = 1.f / (1.f + t); }
        scale = __fdividef(pv, sum);
        p_unv = qv;
      }
    }
    const bool s1_unv = unv_in_slot && lane == 31;
    float p0v = e0 * scale, p1v = e1 * scale;
    if (MODEL == 0 && prior) {                                       // then / priors (:2571-2572)
      p0v = __fdividef(p0v, prior[v0 ? ipk[0] + 1 : 0]);
      p1v = __fdividef(p1v, prior[v1 ? ipk[1] + 1 : 0]);
    }
    if (s1_unv) p1v = p_unv;
    if (out_log) {
      p0v = logf(p0v + kTinyF);
      p1v = logf(p1v + kTinyF);
    }
    const int total = a + S;
    const int k1 = a + (s1_unv ? n_bins : ipk[1]);
    if (v0) orow[a + ipk[0]] = p0v;
    if (v1 || s1_unv) orow[k1] = p1v;
    if (!unv_in_slot && lane == 0) orow[a + n_bins] = out_log ? logf(p_unv + kTinyF) : p_unv;
    __syncwarp();
    {
      float* oa = o_f - a;                                           // 16-byte aligned
      const int j0 = a ? 1 : 0, j1 = total >> 2;
#pragma unroll
      for (int i = 0; i < (W::OUT / 4 + 31) / 32; ++i) {
        const int j = lane + 32 * i;
        if (j >= j0 && j < j1) reinterpret_cast<float4*>(oa)[j] = reinterpret_cast<const float4*>(orow)[j];
      }
      if (a && lane >= a && lane < 4) oa[lane] = orow[lane];         // head: the rest of the first float4
      const int tl = 4 * j1 + lane;
      if (lane < 4 && tl < total) oa[tl] = orow[tl];                 // tail
    }
    __syncwarp();                                                    // the row has been copied: constant back in place
    if (v0) orow[a + ipk[0]] = zero_out;
    if (v1 || s1_unv) orow[k1] = zero_out;
    if (!unv_in_slot && lane == 0) orow[a + n_bins] = zero_out;
    __syncwarp();
    if (++st == kWStages) st = 0;
    a = (a + a_step) & 3;
    o_f += out_step;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// states -> (voiced, bins): voiced = s < n_bins; bins = min(s, n_bins - 1); frames past the clip's length (state -1)
// give voiced = 0, bins = -1.
__global__ void voiced_bins_kernel(const int64_t* __restrict__ states, long long n, int n_bins, uint8_t* __restrict__ voiced,
                                   int64_t* __restrict__ bins) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int64_t s = states[i];
    voiced[i] = (s >= 0 && s < n_bins) ? 1 : 0;
    bins[i] = s < 0 ? -1 : (s < n_bins ? s : n_bins - 1);
  }
}


// ---- post-decode statistics ---------------------------------------------------------------------------------------
// Per frame (est_notes_fn, dcnet/main.py:1911-1934): probs = sigmoid(logits); the decoded bin and its two neighbours
// (|k - bin| <= 1, clipped to the bin range) give
//     note = sum_k (k * note_step) * p_k / max(sum_k p_k, 1e-3) + note_min
// then (viterbi_update_states_tf_fn, dcnet/softmax_viterbi.py:2938-2977) with ref_voicing = ref_note > 0.1,
// diff = |note - ref_note|:  voiced / unvoiced reference frames, correct / incorrect voiced and correct unvoiced
// decisions, pitch hits (diff < 0.5 on voiced reference frames; "strict" = also decoded voiced) and chroma hits
// (|diff - 12 floor(diff / 12 + 0.5)| < 0.5).  Output note = +note where decoded voiced, -note elsewhere (:2975).
// One thread per frame (3 logits, one note, one bin of a 1.3-1.4 KB row: the kernel moves ~50 B per frame);
// counters are reduced per warp and added with one 64-bit atomic per warp and counter.
constexpr int kStatCounters = 9;

__global__ void melody_stats_kernel(const float* __restrict__ logits, int logit_stride, int logit_offset,
                                    const float* __restrict__ ref_notes, const int64_t* __restrict__ bins,
                                    const uint8_t* __restrict__ voiced, const int32_t* __restrict__ lengths, int T,
                                    int n_bins, float note_min, float note_step, float* __restrict__ est_notes,
                                    unsigned long long* __restrict__ counters) {
  const int b = blockIdx.y;
  const int len = lengths ? lengths[b] : T;
  const int lane = threadIdx.x & 31;
  unsigned cnt[kStatCounters];
#pragma unroll
  for (int k = 0; k < kStatCounters; ++k) cnt[k] = 0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const size_t f = (size_t)b * T + t;
    if (t >= len) {
      est_notes[f] = 0.f;
      continue;
    }
    const int64_t bin64 = bins[f];
    const int bin = (int)(bin64 < 0 ? 0 : (bin64 >= n_bins ? n_bins - 1 : bin64));
    const bool ev = voiced[f] != 0;
    const float* row = logits + f * (size_t)logit_stride + logit_offset;
    float num = 0.f, den = 0.f;
#pragma unroll
    for (int d = -1; d <= 1; ++d) {
      const int k = bin + d;
      if (k >= 0 && k < n_bins) {
        const float p = 1.f / (1.f + expf(-row[k]));
        num = __fadd_rn(num, __fmul_rn(__fmul_rn((float)k, note_step), p));
        den = __fadd_rn(den, p);
      }
    }
    const float note = __fadd_rn(__fdiv_rn(num, fmaxf(den, 1e-3f)), note_min);
    est_notes[f] = ev ? note : -note;
    const float ref = ref_notes[f];
    const bool rv = ref > 0.1f;
    const float diff = fabsf(note - ref);
    const bool pitch = rv && diff < 0.5f;
    const float oct = floorf(diff / 12.f + 0.5f) * 12.f;
    const bool chroma = rv && fabsf(diff - oct) < 0.5f;
    cnt[0] += rv;
    cnt[1] += !rv;
    cnt[2] += rv && ev;
    cnt[3] += !rv && ev;
    cnt[4] += !rv && !ev;
    cnt[5] += pitch;
    cnt[6] += pitch && ev;
    cnt[7] += chroma;
    cnt[8] += chroma && ev;
  }
#pragma unroll
  for (int k = 0; k < kStatCounters; ++k) {
    const unsigned v = __reduce_add_sync(0xffffffffu, cnt[k]);
    if (lane == 0 && v) atomicAdd(&counters[(size_t)b * kStatCounters + k], (unsigned long long)v);
  }
}

int emissions_run(const float* logits, const float* prior, int B, int T, int n_bins, int model, int spw, float threshold,
                  int out_log, float* out, cudaStream_t stream) {
  const long long n_frames = (long long)B * T;
  if (n_frames == 0) return VIT_OK;
  if (n_bins > 1024) return VIT_ERR_UNSUPPORTED_ALGO;                 // 32 peak flags per lane
  if (model != 0 && model != 1) return VIT_ERR_INVALID_ARGUMENT;
  int num_sms = 148, dev = 0;
  VIT_CUDA_TRY(cudaGetDevice(&dev));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  const bool no_reg = getenv("VIT_EMIS_GENERIC") != nullptr;         // test / experiment knob: keep the generic kernel
  if (spw == 5 && n_bins >= 7 && n_bins <= 32 * kRegPer && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && !no_reg) {
    // the reference's configuration: register-window kernel, 4 blocks of 8 warps per SM
    const size_t smem = (size_t)kEmisWarps * kRegWarpFloats * sizeof(float);
    long long blocks = (n_frames + kEmisWarps - 1) / kEmisWarps;
    const long long cap = (long long)num_sms * 4;
    if (blocks > cap) blocks = cap;
    if (model == 0) {
      VIT_CUDA_TRY(cudaFuncSetAttribute(emissions_reg_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      emissions_reg_kernel<0><<<(unsigned)blocks, 32 * kEmisWarps, smem, stream>>>(logits, prior, n_frames, n_bins, threshold,
                                                                                  out_log, out);
    } else {
      VIT_CUDA_TRY(cudaFuncSetAttribute(emissions_reg_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      emissions_reg_kernel<1><<<(unsigned)blocks, 32 * kEmisWarps, smem, stream>>>(logits, prior, n_frames, n_bins, threshold,
                                                                                  out_log, out);
    }
    note_launch();
    VIT_CUDA_TRY(cudaGetLastError());
    return VIT_OK;
  }
  if ((spw == 16 || spw == 20) && n_bins > spw + 2 && n_bins <= kWBins - kWPer && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
      !no_reg) {
    // the fine-grid models (jdc: 721 bins, half-width 16; imm: 20): wide register-window kernel, 2 blocks of 8 warps per SM
    long long blocks = (n_frames + kEmisWarps - 1) / kEmisWarps;
    const long long cap = (long long)num_sms * 2;
    if (blocks > cap) blocks = cap;
#define VIT_EMIS_WIDE(M, W_)                                                                                         \
  do {                                                                                                              \
    const size_t smem_w = (size_t)kEmisWarps * WideEmis<W_>::WARP_FLOATS * sizeof(float);                           \
    VIT_CUDA_TRY(cudaFuncSetAttribute(emissions_wide_kernel<M, W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w)); \
    emissions_wide_kernel<M, W_><<<(unsigned)blocks, 32 * kEmisWarps, smem_w, stream>>>(logits, prior, n_frames, n_bins, \
                                                                                         threshold, out_log, out); \
  } while (0)
    if (model == 0) { if (spw == 16) VIT_EMIS_WIDE(0, 16); else VIT_EMIS_WIDE(0, 20); }
    else { if (spw == 16) VIT_EMIS_WIDE(1, 16); else VIT_EMIS_WIDE(1, 20); }
#undef VIT_EMIS_WIDE
    note_launch();
    VIT_CUDA_TRY(cudaGetLastError());
    return VIT_OK;
  }
  const bool fast = spw >= 2 && n_bins >= spw + 2;                   // (the reflect padding needs spw + 1 bins)
  const int pk_cap = (n_bins / (spw + 1) + 2 + 3) & ~3;              // peaks are more than spw bins apart
  const size_t smem = (size_t)kEmisWarps * (3 * emis_row_floats(n_bins, spw) + 2 * pk_cap) * sizeof(float);
  if (smem > 200 * 1024) return VIT_ERR_UNSUPPORTED_ALGO;
  long long blocks = (n_frames + kEmisWarps - 1) / kEmisWarps;
  const long long cap = (long long)num_sms * 8;                      // grid-stride: a multiple of the SM count
  if (blocks > cap) blocks = cap;
#define VIT_EMIS_LAUNCH(M, F, P)                                                                                    \
  do {                                                                                                              \
    if (smem > 48 * 1024)                                                                                           \
      VIT_CUDA_TRY(cudaFuncSetAttribute(emissions_kernel<M, F, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    emissions_kernel<M, F, P><<<(unsigned)blocks, 32 * kEmisWarps, smem, stream>>>(logits, prior, n_frames, n_bins, spw, \
                                                                                   threshold, out_log, pk_cap, out); \
  } while (0)
#define VIT_EMIS_LAUNCH_P(M, F) do { if (n_bins <= 384) VIT_EMIS_LAUNCH(M, F, 12); else VIT_EMIS_LAUNCH(M, F, 24); } while (0)
  if (model == 0) { if (fast) VIT_EMIS_LAUNCH_P(0, true); else VIT_EMIS_LAUNCH_P(0, false); }
  else { if (fast) VIT_EMIS_LAUNCH_P(1, true); else VIT_EMIS_LAUNCH_P(1, false); }
#undef VIT_EMIS_LAUNCH_P
#undef VIT_EMIS_LAUNCH
  note_launch();
  VIT_CUDA_TRY(cudaGetLastError());
  return VIT_OK;
}

int voiced_bins_run(const int64_t* states, long long n, int n_bins, uint8_t* voiced, int64_t* bins, cudaStream_t stream) {
  if (n == 0) return VIT_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  voiced_bins_kernel<<<(unsigned)blocks, 256, 0, stream>>>(states, n, n_bins, voiced, bins);
  note_launch();
  VIT_CUDA_TRY(cudaGetLastError());
  return VIT_OK;
}

int melody_stats_run(const float* logits, int logit_stride, int logit_offset, const float* ref_notes, const int64_t* bins,
                     const uint8_t* voiced, const int32_t* lengths, int B, int T, int n_bins, float note_min,
                     float note_step, float* est_notes, int64_t* counters, cudaStream_t stream) {
  if (B == 0) return VIT_OK;
  VIT_CUDA_TRY(cudaMemsetAsync(counters, 0, (size_t)B * kStatCounters * sizeof(int64_t), stream));
  int gx = (T + 255) / 256;
  if (gx > 64) gx = 64;
  melody_stats_kernel<<<dim3(gx, B), 256, 0, stream>>>(logits, logit_stride, logit_offset, ref_notes, bins, voiced,
                                                        lengths, T, n_bins, note_min, note_step, est_notes,
                                                        reinterpret_cast<unsigned long long*>(counters));
  note_launch();
  VIT_CUDA_TRY(cudaGetLastError());
  return VIT_OK;
}

}  // namespace vit
