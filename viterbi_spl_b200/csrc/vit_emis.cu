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
#include "vit_common.cuh"

namespace vit {

constexpr int kEmisWarps = 8;
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
// SPW5: single_side_peak_width == 5 (dcnet / msnet / ftanet / tonet) takes the fast peak test: the frame is stored with
// its reflect padding materialised, window maxima of width 2 and 4 are built by doubling, and
//   left  max = max(m4[p-5], x[p-1]),  right max = max(m4[p+1], x[p+5])        (p = padded index of the bin)
// -- 11 shared-memory accesses per bin instead of 21.  Any other width uses the direct neighbour scan.
template <int MODEL, bool SPW5>
__global__ void __launch_bounds__(32 * kEmisWarps)
emissions_kernel(const float* __restrict__ logits, const float* __restrict__ prior, long long n_frames, int n_bins, int spw,
                 float threshold, int out_log, int pk_cap, float* __restrict__ out) {
  // per warp: x, m2, m4 rows of n_bins + 16 floats, then the compacted peak list (bin, exp) of pk_cap entries -- peaks
  // are more than spw bins apart, so pk_cap = n_bins / (spw + 1) + 2 suffices and 5 instead of 3 blocks fit an SM
  extern __shared__ float s_x[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int np = n_bins + 16;                                        // padded row: 5 + n_bins + 5, rounded up
  const int per_warp = 3 * np + 2 * pk_cap;
  float* x = s_x + (size_t)w * per_warp + 5;                         // x[k] = logit of bin k; x[-5 .. n+4] valid
  float* m2 = s_x + (size_t)w * per_warp + np + 5;
  float* m4 = s_x + (size_t)w * per_warp + 2 * np + 5;
  int* pk_idx = reinterpret_cast<int*>(s_x + (size_t)w * per_warp + 3 * np);   // compacted peak bins
  float* pk_e = s_x + (size_t)w * per_warp + 3 * np + pk_cap;        // exp(peak logit - max)
  const int n_in = MODEL == 0 ? n_bins + 1 : n_bins;
  const int S = n_bins + 1;
  const float zero_out = out_log ? logf(kTinyF) : 0.f;               // log(0 + tiny) = -87.33655
  // software pipeline over frames: the next frame's logits are fetched into registers (n_bins <= 384) before this frame
  // is processed, so the HBM latency hides under ~700 instructions of work instead of stalling every frame
  constexpr int kPre = 12;
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
    if (SPW5) {
      __syncwarp();
      if (lane < 5) {                                                // np.pad(mode='reflect'): -m -> m, n-1+m -> n-1-m
        x[-1 - lane] = x[1 + lane];
        x[n_bins + lane] = x[n_bins - 2 - lane];
      }
      __syncwarp();
      for (int k = lane - 5; k < n_bins + 4; k += 32) m2[k] = fmaxf(x[k], x[k + 1]);          // max of x[k .. k+1]
      __syncwarp();
      for (int k = lane - 5; k < n_bins + 2; k += 32) m4[k] = fmaxf(m2[k], m2[k + 2]);         // max of x[k .. k+3]
    }
    __syncwarp();
    // pass 1: find the peaks and COMPACT them (ballot + prefix count) into a per-warp list, so that the expensive
    // exp / log below run once per peak on consecutive lanes instead of once per bin under divergence
    int cnt = 0;                                                     // number of voiced peaks (warp-uniform)
    for (int k = lane; k < ((n_bins + 31) & ~31); k += 32) {
      bool pk = false;
      if (k < n_bins) {
        if (SPW5) {
          const float c = x[k];
          pk = (c > fmaxf(m4[k - 5], x[k - 1])) && (c >= fmaxf(m4[k + 1], x[k + 5]));
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
  const bool fast = spw == 5 && n_bins >= 7;
  const int pk_cap = (n_bins / (spw + 1) + 2 + 3) & ~3;              // peaks are more than spw bins apart
  const size_t smem = (size_t)kEmisWarps * (3 * (n_bins + 16) + 2 * pk_cap) * sizeof(float);
  int num_sms = 148, dev = 0;
  VIT_CUDA_TRY(cudaGetDevice(&dev));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  long long blocks = (n_frames + kEmisWarps - 1) / kEmisWarps;
  const long long cap = (long long)num_sms * 8;                      // grid-stride: a multiple of the SM count
  if (blocks > cap) blocks = cap;
#define VIT_EMIS_LAUNCH(M, F)                                                                                       \
  do {                                                                                                              \
    if (smem > 48 * 1024)                                                                                           \
      VIT_CUDA_TRY(cudaFuncSetAttribute(emissions_kernel<M, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    emissions_kernel<M, F><<<(unsigned)blocks, 32 * kEmisWarps, smem, stream>>>(logits, prior, n_frames, n_bins, spw, \
                                                                                threshold, out_log, pk_cap, out);  \
  } while (0)
  if (model == 0) { if (fast) VIT_EMIS_LAUNCH(0, true); else VIT_EMIS_LAUNCH(0, false); }
  else { if (fast) VIT_EMIS_LAUNCH(1, true); else VIT_EMIS_LAUNCH(1, false); }
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
