// VIT_ALGO_BACKPOINTER -- generic max-plus Viterbi for any state count S <= 65535.
//
// Restates the reference recursion (imm/tf_viterbi.py:91-107) with the T2 table kept as uint16 backpointers:
//   forward  : one CTA owns MSEQ clips for all T steps; delta lives in shared memory; every warp owns target states
//              j = warp, warp + nwarps, ...; its lanes stride over the source states i so that the logA^T row
//              (dst-major, contiguous in i -- the reference's `B[j, :]`) is read coalesced once and reused for the
//              MSEQ clips; (value, index) warp-shuffle argmax with the lowest index winning ties (np.argmax);
//   backtrace: one thread per clip chases the backpointers (imm/tf_viterbi.py:102-107).
// This is the shape-agnostic path (and the on-device producer of the reference's T1/T2 tables for parity tests);
// the throughput path for the pitch-bin state sets is VIT_ALGO_CLUSTER (vit_cluster.cu).
#include "vit_common.cuh"

namespace vit {

template <int MSEQ, int THREADS>
__global__ void __launch_bounds__(THREADS)
bp_forward_kernel(const float* __restrict__ logA_T, const float* __restrict__ log_pi,
                  const float* __restrict__ log_emis, const int32_t* __restrict__ lengths,
                  int B, int T_max, int S,
                  uint16_t* __restrict__ bp, float* __restrict__ delta_out,
                  float* __restrict__ scores, int32_t* __restrict__ last_state) {
  extern __shared__ float sm[];                 // delta double buffer: [2][MSEQ][S]
  __shared__ int s_len[MSEQ];
  constexpr int NWARPS = THREADS / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b0 = blockIdx.x * MSEQ;

  if (tid < MSEQ) {
    int b = b0 + tid;
    s_len[tid] = (b < B) ? (lengths ? lengths[b] : T_max) : 0;
  }
  __syncthreads();
  int maxlen = 0;
#pragma unroll
  for (int m = 0; m < MSEQ; ++m) maxlen = max(maxlen, s_len[m]);

  float* prev = sm;
  float* cur = sm + MSEQ * S;

  // t = 0: T1[0] = log_pi + logE[0]                                  (imm/tf_viterbi.py:94)
  for (int x = tid; x < MSEQ * S; x += THREADS) {
    const int m = x / S, j = x - m * S;
    float v = -INFINITY;
    if (s_len[m] > 0) {
      const size_t row = (size_t)(b0 + m) * T_max * S;
      v = __fadd_rn(log_pi[j], log_emis[row + j]);
      if (delta_out) delta_out[row + j] = v;
      if (bp) bp[row + j] = 0;
    }
    prev[x] = v;
  }
  __syncthreads();

  for (int t = 1; t < maxlen; ++t) {
    for (int j = warp; j < S; j += NWARPS) {
      const float* __restrict__ arow = logA_T + (size_t)j * S;
      float best[MSEQ];
      int arg[MSEQ];
#pragma unroll
      for (int m = 0; m < MSEQ; ++m) { best[m] = -INFINITY; arg[m] = 0x7fffffff; }
      for (int i = lane; i < S; i += 32) {
        const float a = arow[i];
#pragma unroll
        for (int m = 0; m < MSEQ; ++m) {
          const float v = __fadd_rn(prev[m * S + i], a);            // Bt[j, i] = T1[t-1][i] + B[j, i]   (:98)
          argmax_combine(best[m], arg[m], v, i);
        }
      }
#pragma unroll
      for (int m = 0; m < MSEQ; ++m) warp_argmax(best[m], arg[m]);  // argmax over i, first maximum wins   (:99)
      if (lane < MSEQ) {
        float bm = best[0];
        int am = arg[0];
#pragma unroll
        for (int m = 1; m < MSEQ; ++m) if (lane == m) { bm = best[m]; am = arg[m]; }
        const int m = lane;
        float out;
        if (t < s_len[m]) {
          const size_t off = ((size_t)(b0 + m) * T_max + t) * S + j;
          out = __fadd_rn(bm, log_emis[off]);                        // T1[t][j] = Bt[j, T2[t][j]] + logE[t][j] (:100)
          if (bp) bp[off] = (uint16_t)am;
          if (delta_out) delta_out[off] = out;
        } else {
          out = prev[m * S + j];                                     // clip already ended: carry its last delta
        }
        cur[m * S + j] = out;
      }
    }
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  }

  // s = argmax(T1[-1])                                              (imm/tf_viterbi.py:103)
  for (int m = warp; m < MSEQ; m += NWARPS) {
    const int b = b0 + m;
    if (b >= B) continue;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    if (s_len[m] > 0) {
      for (int j = lane; j < S; j += 32) argmax_combine(best, arg, prev[m * S + j], j);
      warp_argmax(best, arg);
    }
    if (lane == 0) {
      if (scores) scores[b] = best;
      last_state[b] = (s_len[m] > 0) ? arg : -1;
    }
  }
}

// One thread per clip (imm/tf_viterbi.py:102-107).  Frames past the clip's length are set to -1.
__global__ void bp_backtrace_kernel(const uint16_t* __restrict__ bp, const int32_t* __restrict__ last_state,
                                    const int32_t* __restrict__ lengths, int B, int T_max, int S,
                                    int64_t* __restrict__ paths) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int len = lengths ? lengths[b] : T_max;
  int64_t* p = paths + (size_t)b * T_max;
  for (int t = len; t < T_max; ++t) p[t] = -1;
  if (len <= 0) return;
  int s = last_state[b];
  p[len - 1] = s;
  const uint16_t* q = bp + (size_t)b * T_max * S;
  for (int t = len - 1; t >= 1; --t) {
    s = q[(size_t)t * S + s];
    p[t - 1] = s;
  }
}

// The forward kernel keeps two delta rows of every clip of its CTA in shared memory: S <= 227 KB / 8 B = 29,056 states
// with one clip per CTA (far beyond any pitch-bin state set; uint16 backpointers would allow 65,535).
bool bp_supported(int S) { return S >= 1 && (size_t)2 * S * sizeof(float) <= (size_t)227 * 1024; }

size_t bp_workspace_bytes(int B, int T_max, int S, bool external_bp) {
  size_t bytes = align_up((size_t)(B > 0 ? B : 1) * sizeof(int32_t), 256);      // last_state
  if (!external_bp) bytes += align_up((size_t)B * T_max * S * sizeof(uint16_t), 256);
  return bytes;
}

template <int MSEQ>
static int launch_forward(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths,
                          int B, int T_max, int S, uint16_t* bp, float* delta_out, float* scores,
                          int32_t* last_state, cudaStream_t stream) {
  constexpr int THREADS = 256;
  auto kern = bp_forward_kernel<MSEQ, THREADS>;
  const size_t smem = (size_t)2 * MSEQ * S * sizeof(float);
  if (smem > 48 * 1024) {
    if (smem > 227 * 1024) return VIT_ERR_UNSUPPORTED_ALGO;
    VIT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int grid = (B + MSEQ - 1) / MSEQ;
  kern<<<grid, THREADS, smem, stream>>>(logA_T, log_pi, log_emis, lengths, B, T_max, S, bp, delta_out, scores,
                                        last_state);
  note_launch();
  VIT_CUDA_TRY(cudaGetLastError());
  return VIT_OK;
}

int bp_decode(const float* logA_T, const float* log_pi, const float* log_emis, const int32_t* lengths,
              int B, int T_max, int S, void* workspace, size_t workspace_bytes,
              int64_t* paths, float* scores, uint16_t* bp_out, float* delta_out, cudaEvent_t ev0, cudaEvent_t ev1,
              cudaStream_t stream) {
  if (S > 65535) return VIT_ERR_STATES_TOO_MANY;
  if (!bp_supported(S)) return VIT_ERR_UNSUPPORTED_ALGO;
  if (workspace_bytes < bp_workspace_bytes(B, T_max, S, bp_out != nullptr)) return VIT_ERR_WORKSPACE_TOO_SMALL;
  if (B == 0) return VIT_OK;
  char* ws = (char*)workspace;
  int32_t* last_state = (int32_t*)ws;
  ws += align_up((size_t)B * sizeof(int32_t), 256);
  uint16_t* bp = bp_out ? bp_out : (uint16_t*)ws;

  int num_sms = 148, dev = 0;
  VIT_CUDA_TRY(cudaGetDevice(&dev));
  VIT_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  // share each logA^T row among MSEQ clips once there are enough clips to keep every SM busy anyway
  int rc;
  if (ev0) VIT_CUDA_TRY(cudaEventRecord(ev0, stream));
  if (B >= 8 * num_sms && (size_t)2 * 8 * S * sizeof(float) <= 200 * 1024)
    rc = launch_forward<8>(logA_T, log_pi, log_emis, lengths, B, T_max, S, bp, delta_out, scores, last_state, stream);
  else if (B >= 2 * num_sms && (size_t)2 * 4 * S * sizeof(float) <= 200 * 1024)
    rc = launch_forward<4>(logA_T, log_pi, log_emis, lengths, B, T_max, S, bp, delta_out, scores, last_state, stream);
  else
    rc = launch_forward<1>(logA_T, log_pi, log_emis, lengths, B, T_max, S, bp, delta_out, scores, last_state, stream);
  if (rc != VIT_OK) return rc;
  if (ev1) VIT_CUDA_TRY(cudaEventRecord(ev1, stream));

  bp_backtrace_kernel<<<(B + 127) / 128, 128, 0, stream>>>(bp, last_state, lengths, B, T_max, S, paths);
  note_launch();
  VIT_CUDA_TRY(cudaGetLastError());
  return VIT_OK;
}

}  // namespace vit
