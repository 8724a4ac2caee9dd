"""TEST INFRASTRUCTURE ONLY -- float64 scaled forward-backward (sum-product) oracle.

PARITY UNPINNED: the reference has NO forward-backward / posterior-marginal code (correction 2:
its `SoftMaxViterbi` classes are max-product decoders, dcnet/softmax_viterbi.py:2488-2674).  The north star nevertheless
asks for a scaled sum-product pass on the same model, so its semantics are DEFINED here as the
textbook scaled recursion on exactly the quantities the reference's decoders hold:

    A    [S, S]  row-stochastic transition matrix, row = source  (dcnet/viterbi_transition_matrix.py:81-101)
    pi   [S]     initial distribution                             (dcnet/viterbi_init_probs.py:9-24)
    b_t  [S]     emission likelihoods of frame t, the output of SoftMaxViterbi.observation_probs_fn
                 (dcnet/softmax_viterbi.py:2530-2579): >= 0, unnormalised, possibly divided by the prior

    alpha_0 ~ pi * b_0                       alpha_t ~ (alpha_{t-1} A) * b_t            c_t = the normaliser
    beta_{T-1} = 1                           beta_t = A (b_{t+1} * beta_{t+1}) / c_{t+1}
    gamma_t = alpha_t * beta_t               log L = sum_t log c_t

It is validated against brute-force enumeration of all state paths on tiny models and against a log-sum-exp
formulation (tests/test_fb_oracle.py); the CUDA kernel is compared with it at 1e-4 absolute on gamma and 1e-5
relative on log L (the north star's tolerances).
"""
import itertools

import numpy as np


def forward_backward_np(A, pi, lik_ts):
    """float64 scaled forward-backward of ONE clip.  lik_ts: [T, S] likelihoods.  Returns (gamma [T, S], loglik)."""
    A = np.asarray(A, np.float64)
    pi = np.asarray(pi, np.float64)
    b = np.asarray(lik_ts, np.float64)
    T, S = b.shape
    alpha = np.empty((T, S))
    c = np.empty(T)
    a = pi * b[0]
    c[0] = a.sum()
    alpha[0] = a / c[0]
    for t in range(1, T):
        a = (alpha[t - 1] @ A) * b[t]
        c[t] = a.sum()
        alpha[t] = a / c[t]
    beta = np.ones(S)
    gamma = np.empty((T, S))
    gamma[T - 1] = alpha[T - 1]
    for t in range(T - 2, -1, -1):
        beta = (A @ (b[t + 1] * beta)) / c[t + 1]
        gamma[t] = alpha[t] * beta
    return gamma, float(np.log(c).sum())


def forward_backward_batch_np(A, pi, lik_bts, lengths=None):
    """Batch of clips [B, T, S] with optional lengths; frames past a clip's length get gamma = 0; empty clips get
    loglik = 0."""
    B, T, S = lik_bts.shape
    gamma = np.zeros((B, T, S))
    loglik = np.zeros(B)
    for b in range(B):
        n = T if lengths is None else int(lengths[b])
        if n > 0:
            gamma[b, :n], loglik[b] = forward_backward_np(A, pi, lik_bts[b, :n])
    return gamma, loglik


def forward_backward_logsumexp_np(A, pi, lik_ts):
    """Same posteriors through unscaled log-domain recursions (independent formulation, for self-validation)."""
    with np.errstate(divide='ignore'):
        lA, lpi, lb = np.log(np.asarray(A, np.float64)), np.log(np.asarray(pi, np.float64)), np.log(np.asarray(lik_ts, np.float64))
    T, S = lb.shape

    def lse(x, axis):
        m = np.max(x, axis=axis, keepdims=True)
        m = np.where(np.isfinite(m), m, 0.0)
        return np.squeeze(m, axis) + np.log(np.sum(np.exp(x - m), axis=axis))

    la = np.empty((T, S))
    la[0] = lpi + lb[0]
    for t in range(1, T):
        la[t] = lse(la[t - 1][:, None] + lA, 0) + lb[t]
    lbeta = np.zeros((T, S))
    for t in range(T - 2, -1, -1):
        lbeta[t] = lse(lA + (lb[t + 1] + lbeta[t + 1])[None, :], 1)
    ll = lse(la[T - 1], 0)
    return np.exp(la + lbeta - ll), float(ll)


def forward_backward_bruteforce(A, pi, lik_ts):
    """Enumerates all S^T state paths (tiny models only)."""
    A, pi, b = np.asarray(A, np.float64), np.asarray(pi, np.float64), np.asarray(lik_ts, np.float64)
    T, S = b.shape
    gamma = np.zeros((T, S))
    total = 0.0
    for path in itertools.product(range(S), repeat=T):
        p = pi[path[0]] * b[0, path[0]]
        for t in range(1, T):
            p *= A[path[t - 1], path[t]] * b[t, path[t]]
        total += p
        for t in range(T):
            gamma[t, path[t]] += p
    return gamma / total, float(np.log(total))
