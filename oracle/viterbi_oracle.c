/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the log-domain Viterbi hot path.
 *
 * This file is a plain-C restatement of the reference's NumPy decode and is used ONLY as the
 * checker (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs).
 * Nothing under viterbi_spl_b200/ may call it; the product path is the CUDA library.
 *
 * Reference followed (paths relative to the reference checkout):
 *   imm/tf_viterbi.py:75-109  viterbi_librosa_fn  (oracle of record, log-domain fp32)
 *     :94      T1[0] = log_pi + logE[0]
 *     :97-100  for t: Bt = T1[t-1] + B ; T2[t] = argmax(Bt, axis=1) ; T1[t] = Bt[j, T2[t][j]] + logE[t]
 *     :103-107 s = argmax(T1[-1]) ; backtrace through T2
 *   identical loops: dcnet/softmax_viterbi.py:2467-2485, :2655-2674; tonet/softmax_priors.py:1860-1878;
 *   imm/tf_imm.py:109-127; dcnet/tf_viterbi_decoding.py:101-114 (numba twin).
 *
 * Semantics that matter for bit-exactness:
 *   - every add is a single IEEE-754 binary32 round-to-nearest add (no FMA, no wider intermediates:
 *     compile WITHOUT -ffast-math; x86-64 SSE scalar/vector float adds are exact binary32 ops);
 *   - np.argmax returns the FIRST index attaining the maximum ("v > best" with strict >);
 *   - B is logA transposed ("dst-major"): B[j*S + i] = log A[i -> j];
 *   - no renormalisation of T1 is ever done.
 *
 * Parity status: PINNED -- tests/test_oracle.py checks this file against golden vectors produced by
 * executing the reference's own functions (tests/golden/make_golden.py) and, when /root/reference is
 * present, against the live reference functions.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <unistd.h>

#define VIT_ORACLE_OK 0
#define VIT_ORACLE_EINVAL -1
#define VIT_ORACLE_ENOMEM -2

int vit_oracle_version(void) { return 1; }

/* argmax_i (prev[i] + brow[i]), first maximum wins; writes the fp32 sum at the argmax to *best_out.
 * imm/tf_viterbi.py:98-100 (np.add into Bt, np.argmax(axis=1), take_along_axis). */
static inline int32_t first_argmax_add(const float* prev, const float* brow, int S, float* tmp, float* best_out) {
  /* pass 1: Bt row (np.add, fp32) and its maximum -- written so gcc can vectorise it */
  float m;
  for (int i = 0; i < S; ++i) tmp[i] = prev[i] + brow[i];
  m = tmp[0];
  for (int i = 1; i < S; ++i) m = (tmp[i] > m) ? tmp[i] : m;
  /* pass 2: first index attaining it (np.argmax tie rule) */
  int32_t arg = 0;
  for (int i = 0; i < S; ++i) {
    if (tmp[i] == m) { arg = i; break; }
  }
  *best_out = tmp[arg];
  return arg;
}

/*
 * One sequence. log_emis is [T][S] row-major (the reference transposes its [S,T] input to this
 * layout at imm/tf_viterbi.py:89). Outputs: states[T] (int64, as NumPy), *score = max_j T1[T-1][j]
 * (computed and discarded by the reference at :103). Optional dumps T1_out [T][S] f32 and
 * T2_out [T][S] int32 (row 0 of T2 is zero-filled; the reference leaves it uninitialised).
 */
int vit_oracle_decode_f32(const float* logA_T, const float* log_pi, const float* log_emis, int T, int S,
                          int64_t* states, float* score, float* T1_out, int32_t* T2_out) {
  if (!logA_T || !log_pi || !log_emis || !states || T < 1 || S < 1) return VIT_ORACLE_EINVAL;
  float* T1 = T1_out ? T1_out : (float*)malloc((size_t)T * S * sizeof(float));
  int32_t* T2 = T2_out ? T2_out : (int32_t*)malloc((size_t)T * S * sizeof(int32_t));
  float* tmp = (float*)malloc((size_t)S * sizeof(float));
  if (!T1 || !T2 || !tmp) {
    if (!T1_out) free(T1);
    if (!T2_out) free(T2);
    free(tmp);
    return VIT_ORACLE_ENOMEM;
  }
  for (int j = 0; j < S; ++j) {           /* :94 */
    T1[j] = log_pi[j] + log_emis[j];
    T2[j] = 0;
  }
  for (int t = 1; t < T; ++t) {           /* :97-100 */
    const float* prev = T1 + (size_t)(t - 1) * S;
    float* cur = T1 + (size_t)t * S;
    int32_t* bp = T2 + (size_t)t * S;
    const float* e = log_emis + (size_t)t * S;
    for (int j = 0; j < S; ++j) {
      float best;
      bp[j] = first_argmax_add(prev, logA_T + (size_t)j * S, S, tmp, &best);
      cur[j] = best + e[j];
    }
  }
  {                                       /* :103-107 */
    const float* last = T1 + (size_t)(T - 1) * S;
    int32_t s = 0;
    float m = last[0];
    for (int j = 1; j < S; ++j) if (last[j] > m) { m = last[j]; s = j; }
    if (score) *score = m;
    states[T - 1] = s;
    for (int t = T - 2; t >= 0; --t) {
      s = T2[(size_t)(t + 1) * S + s];
      states[t] = s;
    }
  }
  if (!T1_out) free(T1);
  if (!T2_out) free(T2);
  free(tmp);
  return VIT_ORACLE_OK;
}

/*
 * Batch of independent clips (the reference decodes one recording per call,
 * dcnet/softmax_viterbi.py:3033-3040; clips are independent). log_emis is [B][T_max][S];
 * lengths[b] in [0, T_max] (NULL = all T_max). paths is [B][T_max], frames >= length set to -1;
 * a zero-length clip gets score -inf. Clips are handed out to `nthreads` POSIX threads through an
 * atomic counter (nthreads <= 0: one per online core).
 */
typedef struct {
  const float* logA_T; const float* log_pi; const float* log_emis; const int32_t* lengths;
  int B, T_max, S; int64_t* paths; float* scores;
  int next;   /* atomic work counter */
  int rc;     /* first error */
} vit_batch_job;

static void* vit_batch_worker(void* arg) {
  vit_batch_job* J = (vit_batch_job*)arg;
  for (;;) {
    int b = __atomic_fetch_add(&J->next, 1, __ATOMIC_RELAXED);
    if (b >= J->B) break;
    int len = J->lengths ? J->lengths[b] : J->T_max;
    int64_t* p = J->paths + (size_t)b * J->T_max;
    if (len < 0 || len > J->T_max) { __atomic_store_n(&J->rc, VIT_ORACLE_EINVAL, __ATOMIC_RELAXED); continue; }
    for (int t = len; t < J->T_max; ++t) p[t] = -1;
    if (len == 0) { if (J->scores) J->scores[b] = -INFINITY; continue; }
    float sc = 0.f;
    int r = vit_oracle_decode_f32(J->logA_T, J->log_pi, J->log_emis + (size_t)b * J->T_max * J->S, len, J->S,
                                  p, &sc, NULL, NULL);
    if (r != VIT_ORACLE_OK) __atomic_store_n(&J->rc, r, __ATOMIC_RELAXED);
    if (J->scores) J->scores[b] = sc;
  }
  return NULL;
}

int vit_oracle_max_threads(void) {
  long n = sysconf(_SC_NPROCESSORS_ONLN);
  return n > 0 ? (int)n : 1;
}

int vit_oracle_decode_batch_f32(const float* logA_T, const float* log_pi, const float* log_emis,
                                const int32_t* lengths, int B, int T_max, int S,
                                int64_t* paths, float* scores, int nthreads) {
  if (!logA_T || !log_pi || !log_emis || !paths || B < 0 || T_max < 1 || S < 1) return VIT_ORACLE_EINVAL;
  vit_batch_job J = {logA_T, log_pi, log_emis, lengths, B, T_max, S, paths, scores, 0, VIT_ORACLE_OK};
  if (nthreads <= 0) nthreads = vit_oracle_max_threads();
  if (nthreads > B) nthreads = B;
  if (nthreads <= 1) { vit_batch_worker(&J); return J.rc; }
  pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
  if (!th) return VIT_ORACLE_ENOMEM;
  int started = 0;
  for (int i = 0; i < nthreads; ++i) {
    if (pthread_create(&th[i], NULL, vit_batch_worker, &J) != 0) break;
    ++started;
  }
  if (started == 0) vit_batch_worker(&J);
  for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
  free(th);
  return J.rc;
}

/*
 * ONE long sequence, the same recursion with the S targets of every step split over `nthreads` POSIX threads (a
 * spin barrier per step): the single 1,000,000-frame sequence of BASELINE.json's config 5 takes ~200 s through
 * vit_oracle_decode_f32 on one core, ~15 s this way on 16.  Each (t, j) cell is computed by exactly the same
 * first_argmax_add as above -- threads only partition j -- so the result is bit-identical to the single-threaded
 * function (tests/test_oracle.py asserts it).  Only two T1 rows are kept (the reference's full T1 table is never read
 * again after the step that consumes it, imm/tf_viterbi.py:97-107); T2 is uint16 (S <= 65535).
 */
typedef struct {
  const float* logA_T; const float* log_emis; int T, S, nthreads;
  float* rows;        /* [2][S] */
  uint16_t* T2;       /* [T][S] */
  int count; int sense;
} vit_long_job;

typedef struct { vit_long_job* J; int tid; } vit_long_arg;

static inline void vit_spin_barrier(vit_long_job* J, int* local_sense) {
  *local_sense = !*local_sense;
  if (__atomic_add_fetch(&J->count, 1, __ATOMIC_ACQ_REL) == J->nthreads) {
    __atomic_store_n(&J->count, 0, __ATOMIC_RELAXED);
    __atomic_store_n(&J->sense, *local_sense, __ATOMIC_RELEASE);
  } else {
    while (__atomic_load_n(&J->sense, __ATOMIC_ACQUIRE) != *local_sense) __builtin_ia32_pause();
  }
}

static void* vit_long_worker(void* p) {
  vit_long_arg* a = (vit_long_arg*)p;
  vit_long_job* J = a->J;
  const int S = J->S;
  const int j0 = (int)((long long)S * a->tid / J->nthreads), j1 = (int)((long long)S * (a->tid + 1) / J->nthreads);
  float* tmp = (float*)malloc((size_t)S * sizeof(float));
  int sense = 0;
  for (int t = 1; t < J->T; ++t) {
    const float* prev = J->rows + (size_t)((t - 1) & 1) * S;
    float* cur = J->rows + (size_t)(t & 1) * S;
    const float* e = J->log_emis + (size_t)t * S;
    uint16_t* bp = J->T2 + (size_t)t * S;
    for (int j = j0; j < j1; ++j) {
      float best;
      bp[j] = (uint16_t)first_argmax_add(prev, J->logA_T + (size_t)j * S, S, tmp, &best);
      cur[j] = best + e[j];
    }
    vit_spin_barrier(J, &sense);
  }
  free(tmp);
  return NULL;
}

int vit_oracle_decode_long_f32(const float* logA_T, const float* log_pi, const float* log_emis, int T, int S,
                               int64_t* states, float* score, int nthreads) {
  if (!logA_T || !log_pi || !log_emis || !states || T < 1 || S < 1 || S > 65535) return VIT_ORACLE_EINVAL;
  if (nthreads <= 0) nthreads = vit_oracle_max_threads();
  if (nthreads > S) nthreads = S;
  if (nthreads > 64) nthreads = 64;
  vit_long_job J = {logA_T, log_emis, T, S, nthreads, NULL, NULL, 0, 0};
  J.rows = (float*)malloc((size_t)2 * S * sizeof(float));
  J.T2 = (uint16_t*)malloc((size_t)T * S * sizeof(uint16_t));
  if (!J.rows || !J.T2) { free(J.rows); free(J.T2); return VIT_ORACLE_ENOMEM; }
  for (int j = 0; j < S; ++j) J.rows[j] = log_pi[j] + log_emis[j];          /* :94 */
  pthread_t th[64];
  vit_long_arg args[64];
  int started = 0;
  for (int i = 1; i < nthreads; ++i) {
    args[i].J = &J; args[i].tid = i;
    if (pthread_create(&th[i], NULL, vit_long_worker, &args[i]) != 0) break;
    ++started;
  }
  if (started != nthreads - 1) {            /* could not start them all: the barrier would never fill */
    J.nthreads = 1;                          /* (threads already started would spin; none were: creation is all-or-nothing here) */
    if (started > 0) { free(J.rows); free(J.T2); return VIT_ORACLE_ENOMEM; }
  }
  args[0].J = &J; args[0].tid = 0;
  vit_long_worker(&args[0]);
  for (int i = 1; i <= started; ++i) pthread_join(th[i], NULL);
  {
    const float* last = J.rows + (size_t)((T - 1) & 1) * S;                   /* :103-107 */
    int32_t s = 0;
    float m = last[0];
    for (int j = 1; j < S; ++j) if (last[j] > m) { m = last[j]; s = j; }
    if (score) *score = m;
    states[T - 1] = s;
    for (int t = T - 2; t >= 0; --t) {
      s = J.T2[(size_t)(t + 1) * S + s];
      states[t] = s;
    }
  }
  free(J.rows);
  free(J.T2);
  return VIT_ORACLE_OK;
}
