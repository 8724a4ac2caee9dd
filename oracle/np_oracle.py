"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference decode (the checker, never the product).

Follows, statement by statement, the reference's canonical log-domain function
``imm/tf_viterbi.py:75-109`` (``viterbi_librosa_fn``); the same loop appears in every decoder class
(``dcnet/softmax_viterbi.py:2467-2485`` and ``:2655-2674``, ``tonet/softmax_priors.py:1860-1878``,
``imm/tf_imm.py:109-127``).  Parity status: pinned by ``tests/test_oracle.py`` against golden vectors made by
executing the reference's own code (``tests/golden/make_golden.py``).
"""
import numpy as np

TINY = np.finfo(np.float32).tiny  # dcnet/softmax_viterbi.py:2459 (`tinyp`)


def viterbi_log_np(log_transition_matrix_T, log_prob_init, log_probs_ts, return_tables=False):
    """Log-domain fp32 Viterbi, emissions laid out ``[T, S]``.

    imm/tf_viterbi.py:91-107 with ``probs`` already ``[T, S]`` C-contiguous.  Returns ``(states int64[T],
    score float32)``; ``score = max_j T1[T-1][j]`` (the reference computes the argmax of that row at :103).
    """
    B = np.require(log_transition_matrix_T, np.float32, ['C'])
    probs = np.require(log_probs_ts, np.float32, ['C'])
    prob_init = np.asarray(log_prob_init, np.float32)
    S = len(B)
    assert B.shape == (S, S) and prob_init.shape == (S,) and probs.ndim == 2 and probs.shape[1] == S
    T = probs.shape[0]
    assert T >= 1

    T1 = np.empty([T, S], np.float32)
    T2 = np.zeros([T, S], np.int64)
    T1[0] = prob_init + probs[0]                                   # :94
    Bt = np.empty([S, S], np.float32)
    for t in range(1, T):                                          # :97
        np.add(T1[t - 1], B, out=Bt)                               # :98
        np.argmax(Bt, axis=1, out=T2[t])                           # :99  first maximum wins
        np.add(np.take_along_axis(Bt, indices=T2[t][:, None], axis=1)[:, 0], probs[t], out=T1[t])  # :100

    states = np.empty([T], np.int64)
    s = np.argmax(T1[-1])                                          # :103
    states[-1] = s
    for t in range(T - 2, -1, -1):                                 # :105-107
        s = T2[t + 1, s]
        states[t] = s
    score = np.float32(T1[-1].max())
    if return_tables:
        return states, score, T1, T2
    return states, score


def viterbi_log_st_np(*, log_transition_matrix_T, log_prob_init, log_probs_st):
    """Exact signature of imm/tf_viterbi.py:75 (emissions ``[S, T]``); returns states only."""
    assert log_transition_matrix_T.flags['C_CONTIGUOUS'] and log_transition_matrix_T.dtype == np.float32
    assert log_probs_st.dtype == np.float32 and log_probs_st.shape[0] == len(log_transition_matrix_T)
    return viterbi_log_np(log_transition_matrix_T, log_prob_init, np.require(log_probs_st.T, requirements=['C']))[0]


def family_a_np(*, transition_matrix, prob_init, probs_st):
    """Family A (prob-domain in, logs taken per call): dcnet/softmax_viterbi.py:2433-2485."""
    B = transition_matrix
    probs = probs_st
    S = len(B)
    T = probs.shape[1]
    assert B.shape == (S, S) and probs.shape == (S, T)
    assert np.allclose(np.sum(B, axis=1), 1.) and len(prob_init) == S and np.isclose(np.sum(prob_init), 1.)
    tinyp = np.finfo(probs.dtype).tiny                             # :2459
    B = np.require(np.log(B.T + tinyp), requirements=['C'])        # :2461-2462
    prob_init = np.log(prob_init + tinyp)                          # :2463
    probs = np.require(np.log(probs.T + tinyp), requirements=['C'])  # :2464-2465
    return viterbi_log_np(B, prob_init, probs)[0]


def decode_batch_np(logA_T, log_pi, log_emis, lengths=None):
    """Batch helper used by the parity tests: ``log_emis [B, T_max, S]`` -> (paths int64 [B, T_max] with -1 past
    each clip's length, scores float32 [B]).  Clips are independent (dcnet/softmax_viterbi.py:3033-3040)."""
    log_emis = np.asarray(log_emis, np.float32)
    Bn, T_max, S = log_emis.shape
    paths = np.full([Bn, T_max], -1, np.int64)
    scores = np.full([Bn], -np.inf, np.float32)
    for b in range(Bn):
        n = T_max if lengths is None else int(lengths[b])
        if n == 0:
            continue
        paths[b, :n], scores[b] = viterbi_log_np(logA_T, log_pi, log_emis[b, :n])
    return paths, scores
