// The step BEFORE the hot path and the step AFTER it, batched on the GPU so that
// emissions never round-trip the host:
//
//   vit_emissions_f32   acoustic-model logits -> HMM emission table [B][T][S] (probability or log(p + tiny) domain)
//       VIT_EMIS_SOFTMAX  SoftMaxViterbi.observation_probs_fn   dcnet/softmax_viterbi.py:2508-2579 (7 copies)
//       VIT_EMIS_SHAUN    Viterbi.observation_probs_fn          tonet/softmax_priors.py:1722-1786 (max-subtracted
//                         form; dcnet/softmax_viterbi.py:2316-2359 is the same model)
//   vit_voiced_bins     states -> (voiced, bins)                dcnet/softmax_viterbi.py:2427-2431
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
                 float threshold, int out_log, float* __restrict__ out) {
  extern __shared__ float s_x[];                                     // [kEmisWarps][3][n_bins + 16] (SPW5) or [.][n_bins]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int np = n_bins + 16;                                        // padded row: 5 + n_bins + 5, rounded up
  float* x = SPW5 ? s_x + (size_t)w * 3 * np + 5 : s_x + (size_t)w * n_bins;   // x[k] = logit of bin k; x[-5 .. n+4] valid
  float* m2 = s_x + (size_t)w * 3 * np + np + 5;
  float* m4 = s_x + (size_t)w * 3 * np + 2 * np + 5;
  const int n_in = MODEL == 0 ? n_bins + 1 : n_bins;
  const int S = n_bins + 1;
  const float zero_out = out_log ? logf(kTinyF) : 0.f;               // log(0 + tiny) = -87.33655
  for (long long f = (long long)blockIdx.x * kEmisWarps + w; f < n_frames; f += (long long)gridDim.x * kEmisWarps) {
    const float* in = logits + f * n_in;
    float* o = out + f * S;
    const float* vin = in + (MODEL == 0 ? 1 : 0);
    for (int k = lane; k < n_bins; k += 32) x[k] = vin[k];
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
    // pass 1: peak flags of my bins (bit i <-> bin lane + 32 i), maximum peak logit
    float mx = -INFINITY;
    uint32_t mask = 0;
    for (int k = lane, i = 0; k < n_bins; k += 32, ++i) {
      bool pk;
      if (SPW5) {
        const float c = x[k];
        pk = (c > fmaxf(m4[k - 5], x[k - 1])) && (c >= fmaxf(m4[k + 1], x[k + 5]));
      } else {
        pk = is_peak(x, n_bins, k, spw);
      }
      if (pk) {
        mx = fmaxf(mx, x[k]);
        mask |= 1u << i;
      }
    }
    const int n_peaks = __reduce_add_sync(0xffffffffu, __popc(mask));
    float unv_logit = 0.f;
    if (MODEL == 0) {
      unv_logit = in[0];                                             // column 0 is always a peak (:2521)
      mx = fmaxf(mx, unv_logit);
    }
    mx = warp_max(mx);
    // pass 2: sum of exp(peak - max)
    float sum = 0.f;
    for (uint32_t mm = mask; mm; mm &= mm - 1) sum += expf(x[lane + 32 * (__ffs(mm) - 1)] - mx);
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
    // pass 3: write the row (coalesced); non-peaks are exactly 0 -> log(tiny)
    for (int k = lane, i = 0; k < n_bins; k += 32, ++i) {
      float v = zero_out;
      if ((mask >> i) & 1u) {
        float p = expf(x[k] - mx);
        if (MODEL == 0) {
          p = p / sum;                                               // np.divide(peak_logits, t) then / priors (:2568-2572)
          if (prior) p = p / prior[k + 1];
        } else {
          p = p * scale;                                             // t = p_voiced / sum; peak_logits * t (:1777-1778)
        }
        v = out_log ? logf(p + kTinyF) : p;
      }
      o[k] = v;
    }
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

int emissions_run(const float* logits, const float* prior, int B, int T, int n_bins, int model, int spw, float threshold,
                  int out_log, float* out, cudaStream_t stream) {
  const long long n_frames = (long long)B * T;
  if (n_frames == 0) return VIT_OK;
  if (n_bins > 1024) return VIT_ERR_UNSUPPORTED_ALGO;                 // 32 peak flags per lane
  if (model != 0 && model != 1) return VIT_ERR_INVALID_ARGUMENT;
  const bool fast = spw == 5 && n_bins >= 7;
  const size_t smem = fast ? (size_t)kEmisWarps * 3 * (n_bins + 16) * sizeof(float) : (size_t)kEmisWarps * n_bins * sizeof(float);
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
                                                                                threshold, out_log, out);          \
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

}  // namespace vit
