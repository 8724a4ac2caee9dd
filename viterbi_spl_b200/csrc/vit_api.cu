// extern "C" surface of libvit_b200.so (declared in include/vit_b200.h).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>

#include "vit_common.cuh"

namespace vit {

static std::atomic<uint64_t> g_launches{0};
static thread_local std::string t_last_cuda_error;

static thread_local cudaStream_t t_backtrace_stream = nullptr;
cudaStream_t backtrace_stream_override() { return t_backtrace_stream; }
struct BacktraceStreamScope {
  explicit BacktraceStreamScope(cudaStream_t s) { t_backtrace_stream = s; }
  ~BacktraceStreamScope() { t_backtrace_stream = nullptr; }
};

void note_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int cuda_fail(cudaError_t e) {
  t_last_cuda_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e);
  cudaGetLastError();   // clear the sticky-free error state
  return VIT_ERR_CUDA;
}

// vit_backpointer.cu
size_t bp_workspace_bytes(int B, int T_max, int S, bool external_bp);
bool bp_supported(int S);
int bp_decode(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths, int B,
              int T_max, int S, void* workspace, size_t workspace_bytes, int64_t* paths, float* scores,
              uint16_t* bp_out, float* delta_out, cudaEvent_t ev0, cudaEvent_t ev1, cudaStream_t stream);
// vit_cluster.cu
size_t cluster_workspace_bytes(int B, int T_max, int S);
bool cluster_supported(int S);
int cluster_decode(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths, int B,
                   int T_max, int S, void* workspace, size_t workspace_bytes, int64_t* paths, float* scores,
                   float* delta_out, cudaEvent_t ev0, cudaEvent_t ev1, cudaStream_t stream);

// vit_tmem.cu
size_t tmem_workspace_bytes(int B, int T_max, int S);
bool tmem_supported(int S);
int tmem_decode(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths, int B,
                int T_max, int S, void* workspace, size_t workspace_bytes, int64_t* paths, float* scores,
                float* delta_out, int t_begin, int t_end, bool do_backtrace, cudaEvent_t ev0, cudaEvent_t ev1,
                cudaStream_t stream);

int tmem_clips_in_flight(int S, int* out);

// vit_stream.cu
bool stream_supported(int S);
size_t stream_workspace_bytes(int B, int T_max, int S);
int stream_clips_in_flight(int* out);
int stream_decode(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths, int B,
                  int T_max, int S, void* workspace, size_t workspace_bytes, int64_t* paths, float* scores,
                  float* delta_out, int t_begin, int t_end, bool do_backtrace, cudaEvent_t ev0, cudaEvent_t ev1,
                  cudaStream_t stream);

// vit_fb.cu
size_t fb_workspace_bytes(int B, int T_max, int S);
bool fb_supported(int S);
int fb_run(const float* A, const float* pi, const float* lik, const int32_t* lengths, int B, int T_max, int S,
           void* workspace, size_t workspace_bytes, float* gamma, float* loglik, cudaEvent_t ev0, cudaEvent_t ev1,
           cudaStream_t stream, const int* skip = nullptr);

// vit_emis.cu
int emissions_run(const float* logits, const float* prior, int B, int T, int n_bins, int model, int spw, float threshold,
                  int out_log, float* out, cudaStream_t stream);
int voiced_bins_run(const int64_t* states, long long n, int n_bins, uint8_t* voiced, int64_t* bins, cudaStream_t stream);
int melody_stats_run(const float* logits, int logit_stride, int logit_offset, const float* ref_notes, const int64_t* bins,
                     const uint8_t* voiced, const int32_t* lengths, int B, int T, int n_bins, float note_min,
                     float note_step, float* est_notes, int64_t* counters, cudaStream_t stream);

// vit_banded.cu
bool banded_supported(int S, const vit_structure* st);
size_t banded_workspace_bytes(int B, int T_max, int S);
int banded_decode(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths, int B,
                  int T_max, int S, const vit_structure* st, void* workspace, size_t workspace_bytes, int64_t* paths,
                  float* scores, float* delta_out, int t_begin, int t_end, bool do_backtrace, cudaEvent_t ev0,
                  cudaEvent_t ev1, cudaStream_t stream);
int analyze_structure(const float* A, int S, vit_structure* out);
int banded_clips_in_flight(int* out);

// vit_fb_tc.cu
bool fb_tc_supported(int S);
size_t fb_tc_workspace_bytes(int B, int T_max, int S);
int fb_tc_run(const float* A, const float* pi, const float* lik, const int32_t* lengths, int B, int T_max, int S,
              void* workspace, size_t workspace_bytes, float* gamma, float* loglik, cudaStream_t stream);
// vit_fb_banded.cu
bool fb_banded_supported(int S, const vit_structure* st);
size_t fb_banded_extra_bytes(int B, int T_max);
int fb_banded_run(const float* A, const float* pi, const float* lik, const int32_t* lengths, int B, int T_max, int S,
                  const vit_structure* st, void* workspace, size_t workspace_bytes, float* gamma, float* loglik,
                  cudaStream_t stream);
// Which forward-backward kernel runs (VIT_FB_AUTO): the banded FFMA kernel when the caller passed the structure of a
// band + dense-state matrix (S (2d+3) instead of S^2 multiply-adds per frame; every matrix the reference's builders
// produce); else the tcgen05 tensor-core kernel wherever the shape fits (S <= 372: the bf16 hi/lo image of a 124-row
// shard plus the accumulators must fit the 512 TMEM columns) -- 13.1 ms vs the FFMA kernel's 46.2 ms at
// 1024 x 3000 x 361 -- else the dense FFMA kernel (S = 722).
// VIT_FB_IMPL=tc|simt|banded forces one of them where the shape allows (tests run all three).
static int fb_pick_impl(int S, const vit_fb_opts* opts) {
  int impl = opts ? opts->impl : VIT_FB_AUTO;
  const vit_structure* st = opts ? opts->structure : nullptr;
  if (impl == VIT_FB_AUTO) {
    const char* e = getenv("VIT_FB_IMPL");
    if (e && !strcmp(e, "simt") && fb_supported(S)) return VIT_FB_SIMT;
    if (e && !strcmp(e, "tc") && fb_tc_supported(S)) return VIT_FB_TC;
    if (fb_banded_supported(S, st)) return VIT_FB_BANDED;
    return fb_tc_supported(S) ? VIT_FB_TC : VIT_FB_SIMT;
  }
  return impl;
}

static int check_shape(int B, int T_max, int S) {
  if (B < 0 || T_max < 1 || S < 1) return VIT_ERR_INVALID_ARGUMENT;
  if (S > 65535) return VIT_ERR_STATES_TOO_MANY;
  return VIT_OK;
}

// Dense kernel for VIT_ALGO_AUTO.  S <= 384: the tensor-memory kernel (2-CTA clusters, every SM busy; measured 75 % of
// the FP32 max-plus peak vs 70 % for the streaming kernel at S = 361).  Larger S: the tensor-memory kernel needs
// clusters of ceil(S / 192) CTAs -- 6 at S = 722, 132 of 148 SMs, 61 % of peak on full passes -- while the streaming
// kernel runs at 75 % but only with 14 clips on every SM, so the choice depends on how well the batch fills each
// kernel's pass (B = 0: unknown batch, assume a large one).
static int auto_dense_algo(int B, int S) {
  if (tmem_supported(S) && S <= 384) return VIT_ALGO_TMEM;
  if (stream_supported(S)) {
    if (!tmem_supported(S)) return VIT_ALGO_STREAM;
    if (B > 0) {
      // clips per pass of either kernel: asked from the device where there is one, nominal (148 SMs) otherwise
      int qt = 0, qs = 0;
      if (tmem_clips_in_flight(S, &qt) != VIT_OK || qt <= 0) qt = (148 / ((S + 191) / 192)) * 14;
      if (stream_clips_in_flight(&qs) != VIT_OK || qs <= 0) qs = 148 * 14;
      const double cost_t = (double)((B + qt - 1) / qt) * qt / 0.607;
      const double cost_s = (double)((B + qs - 1) / qs) * qs / 0.747;
      return cost_s < cost_t ? VIT_ALGO_STREAM : VIT_ALGO_TMEM;
    }
    return VIT_ALGO_STREAM;
  }
  if (tmem_supported(S)) return VIT_ALGO_TMEM;
  if (cluster_supported(S)) return VIT_ALGO_CLUSTER;
  return bp_supported(S) ? (int)VIT_ALGO_BACKPOINTER : (int)VIT_ERR_UNSUPPORTED_ALGO;
}

static int resolve_algo(int algo, int S, bool want_bp, int B = 0) {
  if (want_bp || algo == VIT_ALGO_BACKPOINTER) return bp_supported(S) ? (int)VIT_ALGO_BACKPOINTER : (int)VIT_ERR_UNSUPPORTED_ALGO;
  if (algo == VIT_ALGO_AUTO) return auto_dense_algo(B, S);
  if (algo == VIT_ALGO_CLUSTER) return cluster_supported(S) ? algo : VIT_ERR_UNSUPPORTED_ALGO;
  if (algo == VIT_ALGO_TMEM) return tmem_supported(S) ? algo : VIT_ERR_UNSUPPORTED_ALGO;
  if (algo == VIT_ALGO_STREAM) return stream_supported(S) ? algo : VIT_ERR_UNSUPPORTED_ALGO;
  return VIT_ERR_INVALID_ARGUMENT;
}

}  // namespace vit

using namespace vit;

extern "C" {

int vit_version(void) { return VIT_B200_VERSION; }

const char* vit_strerror(int code) {
  switch (code) {
    case VIT_OK: return "ok";
    case VIT_ERR_INVALID_ARGUMENT: return "invalid argument (null pointer or non-positive size)";
    case VIT_ERR_STATES_TOO_MANY: return "too many states: S must be <= 65535 (uint16 backpointers)";
    case VIT_ERR_WORKSPACE_TOO_SMALL: return "workspace too small: query vit_workspace_bytes()";
    case VIT_ERR_UNSUPPORTED_ALGO: return "requested algorithm does not support this shape on this device";
    case VIT_ERR_CUDA: return "CUDA runtime error: see vit_last_cuda_error()";
    case VIT_ERR_MISALIGNED: return "workspace pointer must be 256-byte aligned";
    default: return "unknown vit_status";
  }
}

const char* vit_last_cuda_error(void) { return t_last_cuda_error.c_str(); }

uint64_t vit_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int vit_select_algo(int B, int T_max, int S) {
  int rc = check_shape(B, T_max, S);
  if (rc != VIT_OK) return rc;
  return resolve_algo(VIT_ALGO_AUTO, S, false, B);
}

int vit_clips_in_flight(int S, int algo, const vit_structure* structure, int* out_clips) {
  if (!out_clips) return VIT_ERR_INVALID_ARGUMENT;
  int rc = check_shape(1, 1, S);
  if (rc != VIT_OK) return rc;
  if ((algo == VIT_ALGO_AUTO || algo == VIT_ALGO_BANDED) && banded_supported(S, structure))
    return banded_clips_in_flight(out_clips);
  if (algo == VIT_ALGO_BANDED) return VIT_ERR_UNSUPPORTED_ALGO;
  const int a = resolve_algo(algo, S, false);
  if (a < 0) return a;
  if (a == VIT_ALGO_TMEM) return tmem_clips_in_flight(S, out_clips);
  if (a == VIT_ALGO_STREAM) return stream_clips_in_flight(out_clips);
  // cluster / backpointer kernels: no fixed quantum worth planning for; one clip per SM is the natural unit
  int num_sms = 148, dev = 0;
  VIT_CUDA_TRY(cudaGetDevice(&dev));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  *out_clips = num_sms;
  return VIT_OK;
}

int vit_workspace_bytes(int B, int T_max, int S, int algo, size_t* out_bytes) {
  if (!out_bytes) return VIT_ERR_INVALID_ARGUMENT;
  int rc = check_shape(B, T_max, S);
  if (rc != VIT_OK) return rc;
  if (algo == VIT_ALGO_BANDED) {   // whether the matrix qualifies is decided at decode time from opts->structure
    *out_bytes = banded_workspace_bytes(B, T_max, S);
    return VIT_OK;
  }
  int a = resolve_algo(algo, S, false, B);
  if (a < 0) return a;
  *out_bytes = (a == VIT_ALGO_TMEM)      ? tmem_workspace_bytes(B, T_max, S)
               : (a == VIT_ALGO_STREAM)  ? stream_workspace_bytes(B, T_max, S)
               : (a == VIT_ALGO_CLUSTER) ? cluster_workspace_bytes(B, T_max, S)
                                         : bp_workspace_bytes(B, T_max, S, false);
  // VIT_ALGO_AUTO may resolve to the banded kernels at decode time (opts->structure) or to either dense kernel
  // depending on the batch: cover them all
  if (algo == VIT_ALGO_AUTO) {
    size_t b = banded_workspace_bytes(B, T_max, S);
    if (tmem_supported(S)) b = std::max(b, tmem_workspace_bytes(B, T_max, S));
    if (stream_supported(S)) b = std::max(b, stream_workspace_bytes(B, T_max, S));
    if (b > *out_bytes) *out_bytes = b;
  }
  return VIT_OK;
}

int vit_decode_f32_ex(const float* d_logA_T, const float* d_log_pi, const float* d_log_emis,
                      const int32_t* d_lengths, int B, int T_max, int S, void* d_workspace, size_t workspace_bytes,
                      int64_t* d_paths, float* d_scores, const vit_decode_opts* opts, void* stream) {
  int rc = check_shape(B, T_max, S);
  if (rc != VIT_OK) return rc;
  if (!d_logA_T || !d_log_pi || !d_paths || (!d_log_emis && B > 0)) return VIT_ERR_INVALID_ARGUMENT;
  if (!d_workspace && workspace_bytes > 0) return VIT_ERR_INVALID_ARGUMENT;
  if (((uintptr_t)d_workspace & 255u) != 0) return VIT_ERR_MISALIGNED;
  const int algo_req = opts ? opts->algo : VIT_ALGO_AUTO;
  uint16_t* bp_out = opts ? opts->d_backpointers : nullptr;
  float* delta_out = opts ? opts->d_delta : nullptr;
  int algo;
  const vit_structure* structure = opts ? opts->structure : nullptr;
  if (algo_req == VIT_ALGO_BANDED)
    algo = (banded_supported(S, structure) && !bp_out) ? (int)VIT_ALGO_BANDED : (int)VIT_ERR_UNSUPPORTED_ALGO;
  else if (algo_req == VIT_ALGO_AUTO && !bp_out && banded_supported(S, structure))
    algo = VIT_ALGO_BANDED;
  else
    algo = resolve_algo(algo_req, S, bp_out != nullptr, B);
  if (algo < 0) return algo;
  if (!d_workspace) return VIT_ERR_WORKSPACE_TOO_SMALL;
  cudaStream_t st = (cudaStream_t)stream;
  BacktraceStreamScope bt_scope(opts ? (cudaStream_t)opts->backtrace_stream : nullptr);
  cudaEvent_t ev0 = opts ? (cudaEvent_t)opts->ev_forward_begin : nullptr;
  cudaEvent_t ev1 = opts ? (cudaEvent_t)opts->ev_forward_end : nullptr;
  const int f_begin = opts ? opts->frame_begin : 0;
  const int f_end = (opts && opts->frame_end > 0) ? opts->frame_end : T_max;
  const bool skip_bt = opts && opts->skip_backtrace != 0;
  if (algo == VIT_ALGO_BANDED)
    return banded_decode(d_logA_T, d_log_pi, d_log_emis, d_lengths, B, T_max, S, structure, d_workspace,
                         workspace_bytes, d_paths, d_scores, delta_out, f_begin, f_end, !skip_bt, ev0, ev1, st);
  if (algo == VIT_ALGO_TMEM) {
    return tmem_decode(d_logA_T, d_log_pi, d_log_emis, d_lengths, B, T_max, S, d_workspace, workspace_bytes, d_paths,
                       d_scores, delta_out, f_begin, f_end, !skip_bt, ev0, ev1, st);
  }
  if (algo == VIT_ALGO_STREAM)
    return stream_decode(d_logA_T, d_log_pi, d_log_emis, d_lengths, B, T_max, S, d_workspace, workspace_bytes, d_paths,
                         d_scores, delta_out, f_begin, f_end, !skip_bt, ev0, ev1, st);
  if (f_begin != 0 || f_end != T_max || skip_bt) return VIT_ERR_UNSUPPORTED_ALGO;   // frame ranges: tmem / stream / banded only
  if (algo == VIT_ALGO_CLUSTER)
    return cluster_decode(d_logA_T, d_log_pi, d_log_emis, d_lengths, B, T_max, S, d_workspace, workspace_bytes,
                          d_paths, d_scores, delta_out, ev0, ev1, st);
  return bp_decode(d_logA_T, d_log_pi, d_log_emis, d_lengths, B, T_max, S, d_workspace, workspace_bytes, d_paths,
                   d_scores, bp_out, delta_out, ev0, ev1, st);
}

int vit_decode_f32(const float* d_logA_T, const float* d_log_pi, const float* d_log_emis, const int32_t* d_lengths,
                   int B, int T_max, int S, void* d_workspace, size_t workspace_bytes, int64_t* d_paths,
                   float* d_scores, void* stream) {
  return vit_decode_f32_ex(d_logA_T, d_log_pi, d_log_emis, d_lengths, B, T_max, S, d_workspace, workspace_bytes,
                           d_paths, d_scores, nullptr, stream);
}

int vit_fb_workspace_bytes(int B, int T_max, int S, size_t* out_bytes) {
  if (!out_bytes) return VIT_ERR_INVALID_ARGUMENT;
  int rc = check_shape(B, T_max, S);
  if (rc != VIT_OK) return rc;
  if (!fb_supported(S) && !fb_tc_supported(S)) return VIT_ERR_UNSUPPORTED_ALGO;
  const size_t a = fb_supported(S) ? fb_workspace_bytes(B, T_max, S) : 0;
  const size_t b = fb_tc_supported(S) ? fb_tc_workspace_bytes(B, T_max, S) : 0;
  // + the structured kernels' own normalisers and parameter block (they may run next to the dense FFMA fall-back)
  *out_bytes = (a > b ? a : b) + fb_banded_extra_bytes(B, T_max);
  return VIT_OK;
}

int vit_forward_backward_f32_ex(const float* d_A, const float* d_pi, const float* d_lik, const int32_t* d_lengths, int B,
                                int T_max, int S, void* d_workspace, size_t workspace_bytes, float* d_gamma,
                                float* d_loglik, const vit_fb_opts* opts, void* stream) {
  int rc = check_shape(B, T_max, S);
  if (rc != VIT_OK) return rc;
  if (!d_A || !d_pi || !d_gamma || (!d_lik && B > 0)) return VIT_ERR_INVALID_ARGUMENT;
  if (!fb_supported(S) && !fb_tc_supported(S)) return VIT_ERR_UNSUPPORTED_ALGO;
  if (!d_workspace) return VIT_ERR_WORKSPACE_TOO_SMALL;
  if (((uintptr_t)d_workspace & 255u) != 0) return VIT_ERR_MISALIGNED;
  const int impl = fb_pick_impl(S, opts);
  if (impl == VIT_FB_BANDED)
    return fb_banded_run(d_A, d_pi, d_lik, d_lengths, B, T_max, S, opts ? opts->structure : nullptr, d_workspace,
                         workspace_bytes, d_gamma, d_loglik, (cudaStream_t)stream);
  if (impl == VIT_FB_TC) {
    if (!fb_tc_supported(S)) return VIT_ERR_UNSUPPORTED_ALGO;
    return fb_tc_run(d_A, d_pi, d_lik, d_lengths, B, T_max, S, d_workspace, workspace_bytes, d_gamma, d_loglik,
                     (cudaStream_t)stream);
  }
  if (impl != VIT_FB_SIMT || !fb_supported(S)) return VIT_ERR_UNSUPPORTED_ALGO;
  return fb_run(d_A, d_pi, d_lik, d_lengths, B, T_max, S, d_workspace, workspace_bytes, d_gamma, d_loglik, nullptr,
                nullptr, (cudaStream_t)stream);
}

int vit_forward_backward_f32(const float* d_A, const float* d_pi, const float* d_lik, const int32_t* d_lengths, int B,
                             int T_max, int S, void* d_workspace, size_t workspace_bytes, float* d_gamma,
                             float* d_loglik, void* stream) {
  return vit_forward_backward_f32_ex(d_A, d_pi, d_lik, d_lengths, B, T_max, S, d_workspace, workspace_bytes, d_gamma,
                                     d_loglik, nullptr, stream);
}

int vit_emissions_f32(const float* d_logits, const float* d_prior, int B, int T, int n_bins, int model,
                      int single_side_peak_width, float threshold, int out_log, float* d_out, void* stream) {
  if (B < 0 || T < 0 || n_bins < 1 || single_side_peak_width < 0 || single_side_peak_width >= n_bins)
    return VIT_ERR_INVALID_ARGUMENT;
  if ((!d_logits || !d_out) && (long long)B * T > 0) return VIT_ERR_INVALID_ARGUMENT;
  return emissions_run(d_logits, d_prior, B, T, n_bins, model, single_side_peak_width, threshold, out_log, d_out,
                       (cudaStream_t)stream);
}

int vit_voiced_bins(const int64_t* d_states, long long n, int n_bins, uint8_t* d_voiced, int64_t* d_bins, void* stream) {
  if (n < 0 || n_bins < 1 || ((!d_states || !d_voiced || !d_bins) && n > 0)) return VIT_ERR_INVALID_ARGUMENT;
  return voiced_bins_run(d_states, n, n_bins, d_voiced, d_bins, (cudaStream_t)stream);
}

int vit_melody_stats_f32(const float* d_logits, int logit_stride, int logit_offset, const float* d_ref_notes,
                         const int64_t* d_bins, const uint8_t* d_voiced, const int32_t* d_lengths, int B, int T,
                         int n_bins, float note_min, float note_step, float* d_est_notes, int64_t* d_counters,
                         void* stream) {
  if (B < 0 || T < 1 || n_bins < 1 || logit_offset < 0 || logit_stride < logit_offset + n_bins)
    return VIT_ERR_INVALID_ARGUMENT;
  if (B > 65535) return VIT_ERR_INVALID_ARGUMENT;   // one grid row per clip
  if (B > 0 && (!d_logits || !d_ref_notes || !d_bins || !d_voiced || !d_est_notes || !d_counters))
    return VIT_ERR_INVALID_ARGUMENT;
  return melody_stats_run(d_logits, logit_stride, logit_offset, d_ref_notes, d_bins, d_voiced, d_lengths, B, T, n_bins,
                          note_min, note_step, d_est_notes, d_counters, (cudaStream_t)stream);
}

int vit_analyze_structure_f32(const float* h_logA_T, int S, vit_structure* out) {
  if (!h_logA_T || !out || S < 1) return VIT_ERR_INVALID_ARGUMENT;
  return analyze_structure(h_logA_T, S, out);
}

int vit_upload_frames_f32(float* d_log_emis, const float* h_log_emis, int B, int T_max, int S, int frame_begin,
                          int frame_end, void* stream) {
  int rc = check_shape(B, T_max, S);
  if (rc != VIT_OK) return rc;
  if (!d_log_emis || !h_log_emis || frame_begin < 0 || frame_end > T_max || frame_begin > frame_end)
    return VIT_ERR_INVALID_ARGUMENT;
  if (B == 0 || frame_begin == frame_end) return VIT_OK;
  const size_t pitch = (size_t)T_max * S * sizeof(float);
  const size_t width = (size_t)(frame_end - frame_begin) * S * sizeof(float);
  const size_t off = (size_t)frame_begin * S;
  VIT_CUDA_TRY(cudaMemcpy2DAsync(d_log_emis + off, pitch, h_log_emis + off, pitch, width, (size_t)B,
                                 cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return VIT_OK;
}

}  // extern "C"
