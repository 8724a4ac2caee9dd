/*
 * vit_b200.h -- C ABI of the B200-native batched HMM (Viterbi) decoder.
 *
 * Drop-in boundary for the ONE hot path of drwangxian/viterbi_spl: the float32 log-domain Viterbi recursion
 *     delta_t[j] = max_i fl32(delta_{t-1}[i] + logA[i,j]) + logE_t[j]      (first maximum wins, as np.argmax)
 * followed by the backtrace.  The reference has no FFI layer; what it would bind is
 *   - the compiled numba module `viterbi_numba.core(B, prob_init, probs)`      dcnet/aot_viterbi_core.py:8-54,
 *     called from `viterbi_numba_fn`                                            dcnet/tf_viterbi_decoding.py:119-153
 *   - and the NumPy hot loops it duplicates 26 times, canonical form            imm/tf_viterbi.py:75-109.
 * Every entry point below states which of those it replaces.
 *
 * Conventions
 *   - plain C types only; every buffer is owned by the caller (device buffers come from the caller's allocator,
 *     e.g. torch); the library allocates no device memory and keeps no mutable global state except a launch
 *     counter; it is re-entrant per stream and safe to use from one process per GPU.
 *   - all pointers prefixed d_ are DEVICE pointers valid on the current CUDA device.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous on it.
 *   - return value: VIT_OK (0) or a negative vit_status; nothing aborts or throws.
 *   - inputs are LOG-domain float32 (the reference's prob-domain families take log(x + tiny) on the host with
 *     NumPy before calling in, exactly where the reference takes it: dcnet/softmax_viterbi.py:2459-2465), so
 *     paths are bit-identical to the reference's fp32 decode.  Values must be finite or -inf, never NaN.
 */
#ifndef VIT_B200_H_
#define VIT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIT_B200_VERSION 101 /* major*10000 + minor*100 + patch */

typedef enum vit_status {
  VIT_OK = 0,
  VIT_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, B < 0, T_max < 1, S < 1 */
  VIT_ERR_STATES_TOO_MANY = -2,  /* S > 65535: backpointers are uint16 */
  VIT_ERR_WORKSPACE_TOO_SMALL = -3,
  VIT_ERR_UNSUPPORTED_ALGO = -4, /* requested algorithm cannot run this shape on this device */
  VIT_ERR_CUDA = -5,             /* a CUDA runtime call failed; see vit_last_cuda_error() */
  VIT_ERR_MISALIGNED = -6        /* workspace pointer not 256-byte aligned */
} vit_status;

typedef enum vit_algo {
  VIT_ALGO_AUTO = 0,
  /* generic kernel: one CTA per clip, warp-shuffle (value,index) argmax, uint16 backpointer table in the
   * workspace, one-thread-per-clip backtrace.  Any S <= 29,056 (two delta rows per clip in shared memory; larger S gives
   * VIT_ERR_UNSUPPORTED_ALGO from vit_select_algo / vit_workspace_bytes / vit_decode_f32 alike). */
  VIT_ALGO_BACKPOINTER = 1,
  /* persistent thread-block-cluster kernel: logA^T column-sharded and resident in shared memory, delta exchanged
   * through distributed shared memory, register-tiled FADD + FMNMX3 (3-input max) max-plus, delta history (fp32) in the workspace
   * and the argmax resolved lazily by the backtrace only along the decoded path (bit-identical result). */
  VIT_ALGO_CLUSTER = 2,
  /* the throughput path: as VIT_ALGO_CLUSTER, but the resident logA^T shard lives in TENSOR MEMORY (tcgen05.ld into
   * registers) so that 2-CTA clusters -- which pack all 148 SMs, 4-CTA clusters strand 16 -- can hold a 361-state
   * shard; 7 clips x 6 targets per thread, two pipelines per CTA.  S <= 384.  Supports frame ranges. */
  VIT_ALGO_TMEM = 3,
  /* bit-exact fast path for structured matrices (Toeplitz band + one dense voiced/unvoiced state + a constant
   * background, i.e. every matrix the reference's builders produce): S (2d + 3) instead of S^2 cells per frame.  Band
   * in registers for S <= 384, d <= 14 (dcnet / msnet / ftanet / tonet); band in tensor memory for even S <= 768,
   * d <= 40 (jdc).  Needs opts->structure from vit_analyze_structure_f32; VIT_ALGO_AUTO picks it when that says kind = 1. */
  VIT_ALGO_BANDED = 4,
  /* dense recursion with logA^T STREAMED from L2: one CTA per SM owns 14 clips and all S targets, a producer warp feeds
   * 24 KB tiles of the pre-packed matrix through a 4-stage TMA (cp.async.bulk) ring; no clusters, no exchange, all 148
   * SMs busy whatever S is (S <= 1152).  The big-batch path for big state sets (BASELINE config 3: S = 722); needs
   * >= 14 clips per SM to fill the machine.  Supports frame ranges. */
  VIT_ALGO_STREAM = 5
} vit_algo;

/* Structure of a transition matrix as found by vit_analyze_structure_f32 (see csrc/vit_banded.cu for the proof that
 * the banded recursion is bit-identical to the dense one). */
typedef struct vit_structure {
  int32_t kind;        /* 0 = no exploitable structure (use a dense kernel), 1 = band + background */
  int32_t halfwidth;   /* d: apart from dense_index, entries that differ from `background` satisfy |i - j| <= d */
  int32_t dense_index; /* the one state that is a dense source column AND a dense target row (unvoiced), or -1 */
  float background;    /* c = the minimum entry of logA^T (log(tiny) = -87.33655 for the reference's matrices) */
  float dense_row_max; /* max over i != dense_index of logA^T[dense_index][i] (transitions INTO the dense state from the
                          others; +inf if there is no dense state): lets the backtrace stay on the dense row without
                          scanning it */
} vit_structure;

/* Optional extras for vit_decode_f32_ex (all may be zero/NULL). */
typedef struct vit_decode_opts {
  int32_t algo;              /* vit_algo */
  int32_t reserved;
  uint16_t* d_backpointers;  /* [B][T_max][S] out: the reference's T2 table (row t=0 zero). Forces
                                VIT_ALGO_BACKPOINTER. Replaces the int64 T2 of imm/tf_viterbi.py:92,99. */
  float* d_delta;            /* [B][T_max][S] out: the reference's T1 table (imm/tf_viterbi.py:91,94,100). */
  void* ev_forward_begin;    /* optional cudaEvent_t recorded on `stream` right before the forward (recursion) kernel */
  void* ev_forward_end;      /* optional cudaEvent_t recorded right after it (bench.py times the kernel with these) */
  /* Frame range (VIT_ALGO_TMEM, VIT_ALGO_STREAM and VIT_ALGO_BANDED): run the recursion over frames [frame_begin, frame_end) only; frame_end = 0
   * means T_max.  A range with frame_begin > 0 resumes from the delta history that an earlier call on the SAME
   * workspace left behind, so a host can upload a long batch in time slabs and overlap each copy with the recursion
   * over the previous slab.  skip_backtrace != 0 leaves d_paths / d_scores untouched (all but the last slab). */
  int32_t frame_begin;
  int32_t frame_end;
  int32_t skip_backtrace;
  int32_t reserved2;
  const vit_structure* structure; /* HOST pointer or NULL: what vit_analyze_structure_f32 found for d_logA_T */
  /* Optional second cudaStream_t for the backtrace (history-based algorithms): it is ordered after the forward kernel
   * by an event, so the caller's NEXT decode on `stream` overlaps it.  The caller must then (a) wait on this stream
   * before reading d_paths / d_scores and (b) not reuse d_workspace until the backtrace has finished -- i.e. alternate
   * two workspaces (viterbi_spl_b200.decoder.PipelinedDecoder does).  NULL: backtrace on `stream`. */
  void* backtrace_stream;
} vit_decode_opts;

/* Library version (VIT_B200_VERSION of the built library). */
int vit_version(void);

/* Static message for a vit_status. */
const char* vit_strerror(int code);

/* Text of the last CUDA error seen by this thread ("" if none). */
const char* vit_last_cuda_error(void);

/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t vit_launch_count(void);

/* Host-side analysis of a transition matrix (h_logA_T is a HOST pointer, [S][S] dst-major as for vit_decode_f32).
 * Call once per model; pass the result in vit_decode_opts.structure.  Never fails for finite input: a matrix without
 * the structure gets kind = 0. */
int vit_analyze_structure_f32(const float* h_logA_T, int S, vit_structure* out);

/* Which algorithm VIT_ALGO_AUTO resolves to for this shape on the current device (a vit_algo), or a negative
 * vit_status. */
int vit_select_algo(int B, int T_max, int S);

/* Clips ONE launch of the chosen algorithm keeps in flight with every SM busy (persistent kernels: resident CTAs or
 * clusters x clips per CTA; 1036 for the tensor-memory kernel at S = 361, 1184 for the banded kernels on a 148-SM B200).
 * Hosts that split a job larger than HBM into waves (viterbi_spl_b200.waves; SURVEY.md section 8e: "choose
 * sequences-per-CTA so every wave is full") size each wave as a multiple of it.  `structure` as in vit_decode_opts
 * (NULL: dense kernels only).  The reference has no counterpart: it decodes one recording per call
 * (dcnet/softmax_viterbi.py:3033-3040). */
int vit_clips_in_flight(int S, int algo, const vit_structure* structure, int* out_clips);

/* Bytes of device workspace vit_decode_f32 needs for this shape and algorithm (algo may be VIT_ALGO_AUTO).
 * The workspace holds what the reference keeps in its T1/T2 tables (imm/tf_viterbi.py:91-92): the uint16
 * backpointer table or the fp32 delta history, plus the padded/sharded copy of logA^T. */
int vit_workspace_bytes(int B, int T_max, int S, int algo, size_t* out_bytes);

/*
 * Batched decode.  Replaces viterbi_numba.core / viterbi_librosa_fn (imm/tf_viterbi.py:75-109) for B independent
 * clips at once (the reference decodes one recording per call, dcnet/softmax_viterbi.py:3033-3040).
 *
 *   d_logA_T   [S][S]  dst-major: d_logA_T[j*S + i] = log A[i -> j]     (the reference's `B`,  imm/tf_viterbi.py:77)
 *   d_log_pi   [S]                                                      (`log_prob_init`,       :82)
 *   d_log_emis [B][T_max][S] row-major                                  (`probs` after the transpose at :89)
 *   d_lengths  [B] int32 frames per clip, each in [0, T_max]; NULL = all T_max
 *   d_paths    [B][T_max] int64 out (`states`, :102-107); frames >= length are set to -1
 *   d_scores   [B] float32 out, max_j delta_{T-1}[j] (the value whose argmax :103 takes); may be NULL;
 *              -inf for a zero-length clip
 */
int vit_decode_f32(const float* d_logA_T, const float* d_log_pi, const float* d_log_emis,
                   const int32_t* d_lengths, int B, int T_max, int S,
                   void* d_workspace, size_t workspace_bytes,
                   int64_t* d_paths, float* d_scores, void* stream);

/* Same, with an explicit algorithm and optional T1/T2 table outputs. */
int vit_decode_f32_ex(const float* d_logA_T, const float* d_log_pi, const float* d_log_emis,
                      const int32_t* d_lengths, int B, int T_max, int S,
                      void* d_workspace, size_t workspace_bytes,
                      int64_t* d_paths, float* d_scores, const vit_decode_opts* opts, void* stream);

/*
 * Scaled sum-product forward-backward on the same model (north-star item 4).  The reference has NO such pass (its
 * SoftMaxViterbi classes are max-product decoders, dcnet/softmax_viterbi.py:2488-2674); the semantics are the textbook
 * scaled recursion on the quantities those classes hold (oracle/fb_oracle.py states them):
 *   d_A      [S][S]  row-stochastic, row = source state (the matrix viterbi_transition_matrix.dat stores)
 *   d_pi     [S]     initial distribution                (viterbi_init_probs.dat)
 *   d_lik    [B][T_max][S]  emission likelihoods b_t >= 0 (SoftMaxViterbi.observation_probs_fn output, probability
 *                    domain, NOT logged: dcnet/softmax_viterbi.py:2530-2579)
 *   d_gamma  [B][T_max][S]  out: posterior state marginals gamma_t (0 for frames >= length)
 *   d_loglik [B]     out: log L = sum_t log c_t (0 for an empty clip); may be NULL
 * float32 arithmetic; |gamma - float64 oracle| <= 1e-4, log L within 1e-5 relative.  S <= 384 x cluster size limits
 * as for VIT_ALGO_TMEM (vit_fb_workspace_bytes returns VIT_ERR_UNSUPPORTED_ALGO otherwise).
 */
int vit_fb_workspace_bytes(int B, int T_max, int S, size_t* out_bytes);
int vit_forward_backward_f32(const float* d_A, const float* d_pi, const float* d_lik, const int32_t* d_lengths,
                             int B, int T_max, int S, void* d_workspace, size_t workspace_bytes,
                             float* d_gamma, float* d_loglik, void* stream);

/* Same, with an explicit kernel and the structure of the transition matrix.
 *   VIT_FB_TC      tcgen05 tensor-core kernel (dense matrices, S <= 372; products as 2 x bf16 terms, fp32 accumulate)
 *   VIT_FB_SIMT    dense FFMA kernel (any S the tensor-memory plan takes, e.g. 722)
 *   VIT_FB_BANDED  band + one dense state, every other entry exactly 0 -- what viterbi_transition_matrix.py builds
 *                  (dcnet/viterbi_transition_matrix.py:81-98): S (2d + 3) multiply-adds per frame instead of S^2, fp32.
 *                  `structure` = vit_analyze_structure_f32 of the HOST copy of d_A (kind 1, background 0; halfwidth
 *                  <= 14 with S <= 384, or halfwidth <= 56 with S <= 768: the 722-state sets);
 *                  VIT_ERR_UNSUPPORTED_ALGO otherwise.  When, in addition, the band is one tap vector scaled per source
 *                  row (A[i][j] = kappa_i b[j-i] within 2e-6 relative -- the rows of that recipe are one jump histogram,
 *                  normalised) and the dense row / column are constant, the band product runs as a convolution, one warp
 *                  per clip (checked on the device at every call; otherwise the general band kernel runs for S <= 384
 *                  and the dense FFMA kernel above it; VIT_FB_CONV=0 in the environment switches the check off).
 *   VIT_FB_AUTO    banded when `structure` allows it, else tc, else simt (what vit_forward_backward_f32 does with
 *                  structure = NULL).
 * The workspace size does not depend on the kernel. */
#define VIT_FB_AUTO 0
#define VIT_FB_TC 1
#define VIT_FB_SIMT 2
#define VIT_FB_BANDED 3
typedef struct vit_fb_opts {
  int32_t impl;                       /* VIT_FB_* */
  int32_t reserved;
  const vit_structure* structure;     /* of d_A (probability domain); may be NULL */
} vit_fb_opts;
int vit_forward_backward_f32_ex(const float* d_A, const float* d_pi, const float* d_lik, const int32_t* d_lengths,
                                int B, int T_max, int S, void* d_workspace, size_t workspace_bytes,
                                float* d_gamma, float* d_loglik, const vit_fb_opts* opts, void* stream);

/*
 * The step before the decode: acoustic-model logits -> HMM emission table, batched on the GPU.
 *   model VIT_EMIS_SOFTMAX (0): SoftMaxViterbi.observation_probs_fn, dcnet/softmax_viterbi.py:2508-2579:
 *       d_logits [B][T][1 + n_bins], column 0 = the unvoiced logit; d_prior [1 + n_bins] = np.roll(ini_probs, 1)
 *       ("scaled", :2534-2538) or NULL ("unscaled").
 *   model VIT_EMIS_SHAUN (1): Viterbi.observation_probs_fn, tonet/softmax_priors.py:1741-1786:
 *       d_logits [B][T][n_bins]; threshold = log(th / (1 - th)) of the voicing threshold (:1705-1706).
 *   single_side_peak_width: 5 (dcnet/msnet/ftanet/tonet), 16 (jdc), 20 (imm).
 *   d_out [B][T][n_bins + 1], unvoiced state LAST; out_log != 0 writes log(p + tiny) -- what the decoders take at
 *   dcnet/softmax_viterbi.py:2650-2653 -- so the table can go straight into vit_decode_f32.
 * Peak picking is exact; the exp/log values are within 1e-5 relative of NumPy's (not bit-identical: this entry point
 * is outside the decoder's bit-exact claim).
 */
#define VIT_EMIS_SOFTMAX 0
#define VIT_EMIS_SHAUN 1
int vit_emissions_f32(const float* d_logits, const float* d_prior, int B, int T, int n_bins, int model,
                      int single_side_peak_width, float threshold, int out_log, float* d_out, void* stream);

/* The step after the decode (dcnet/softmax_viterbi.py:2427-2431): voiced = state < n_bins,
 * bins = min(state, n_bins - 1); frames past a clip's length (state -1) give voiced = 0, bins = -1. */
int vit_voiced_bins(const int64_t* d_states, long long n, int n_bins, uint8_t* d_voiced, int64_t* d_bins, void* stream);

/*
 * The statistics step after the decode, batched (MetricsInference.viterbi_update_states_tf_fn,
 * dcnet/softmax_viterbi.py:2923-2979, with est_notes_fn, dcnet/main.py:1911-1934): per frame the decoded bin and its
 * two neighbours, weighted by sigmoid(logit), give the refined note
 *     note = sum_{|k-bin|<=1} (k * note_step) p_k / max(sum p_k, 1e-3) + note_min        (note_min 23.6, step 1/5)
 * and nine int64 counters per clip, in the reference's order: reference voiced, reference unvoiced, correct voiced,
 * incorrect voiced, correct unvoiced, pitch hits wide, strict, chroma hits wide, strict.
 *   d_logits   [B][T][logit_stride] float32; bin k of a frame is element logit_offset + k (offset 0 / stride n_bins
 *              for the dcnet form, offset 1 / stride n_bins + 1 where column 0 is the unvoiced logit)
 *   d_ref_notes[B][T] float32 (<= 0.1: unvoiced, :2938);  d_bins / d_voiced: what vit_voiced_bins produced
 *   d_lengths  [B] or NULL;  frames past a clip's length count nowhere and get est_note 0
 *   d_est_notes[B][T] out: +note where decoded voiced, -note where not (:2975)
 *   d_counters [B][VIT_MELODY_COUNTERS] int64 out (zeroed by the call)
 * The reference's version is TensorFlow (absent from this image): float32 like it, checked to 1e-5 on notes against
 * goldens produced by executing the reference's own function text on a NumPy-backed stand-in for its TensorFlow ops
 * (tests/golden/melody_stats.npz; oracle/ref_loader.py dcnet_melody_stats).
 */
#define VIT_MELODY_COUNTERS 9
int vit_melody_stats_f32(const float* d_logits, int logit_stride, int logit_offset, const float* d_ref_notes,
                         const int64_t* d_bins, const uint8_t* d_voiced, const int32_t* d_lengths, int B, int T,
                         int n_bins, float note_min, float note_step, float* d_est_notes, int64_t* d_counters,
                         void* stream);

/*
 * Host -> device upload of frames [frame_begin, frame_end) of every clip of a [B][T_max][S] float32 batch: one strided
 * 2-D async copy on `stream` (h_log_emis should be page-locked for the copy to be asynchronous).  Together with the
 * frame ranges of vit_decode_f32_ex this lets a host overlap the PCIe transfer of time slab k+1 with the recursion
 * over slab k.  (The reference hands the decoder one host array per recording: dcnet/softmax_viterbi.py:3033-3040.)
 */
int vit_upload_frames_f32(float* d_log_emis, const float* h_log_emis, int B, int T_max, int S,
                          int frame_begin, int frame_end, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIT_B200_H_ */
