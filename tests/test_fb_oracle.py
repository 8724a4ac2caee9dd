"""The float64 forward-backward oracle validates ITSELF (there is no reference implementation: parity unpinned,
oracle/fb_oracle.py header): brute-force path enumeration on tiny models, a log-sum-exp formulation on larger ones."""
import numpy as np
import pytest

from oracle import fb_oracle


def random_hmm(S, rng, zeros=False):
    A = rng.random((S, S)) ** 3
    if zeros:
        A[rng.random((S, S)) < 0.4] = 0
        A[:, -1] = np.maximum(A[:, -1], 0.05)          # the unvoiced state is reachable from every state
    A /= A.sum(1, keepdims=True)
    pi = rng.random(S) + 0.01
    return A, pi / pi.sum()


@pytest.mark.parametrize('S,T', [(2, 5), (3, 6), (4, 5), (5, 4)])
def test_against_bruteforce_enumeration(S, T):
    rng = np.random.default_rng(S * 10 + T)
    A, pi = random_hmm(S, rng, zeros=True)
    lik = rng.random((T, S)) * 3
    lik[rng.random((T, S)) < 0.3] = 0
    lik[:, -1] = np.maximum(lik[:, -1], 0.1)
    g, ll = fb_oracle.forward_backward_np(A, pi, lik)
    g2, ll2 = fb_oracle.forward_backward_bruteforce(A, pi, lik)
    assert np.allclose(g, g2, atol=1e-12) and np.isclose(ll, ll2, rtol=1e-12)
    assert np.allclose(g.sum(1), 1)


@pytest.mark.parametrize('S,T', [(30, 200), (97, 60)])
def test_against_logsumexp_formulation(S, T):
    rng = np.random.default_rng(S + T)
    A, pi = random_hmm(S, rng)
    lik = np.exp(3 * rng.standard_normal((T, S)))
    g, ll = fb_oracle.forward_backward_np(A, pi, lik)
    g2, ll2 = fb_oracle.forward_backward_logsumexp_np(A, pi, lik)
    assert np.allclose(g, g2, atol=1e-10) and np.isclose(ll, ll2, rtol=1e-10)


def test_batch_lengths():
    rng = np.random.default_rng(1)
    A, pi = random_hmm(6, rng)
    lik = rng.random((3, 9, 6)) + 0.1
    g, ll = fb_oracle.forward_backward_batch_np(A, pi, lik, [9, 0, 4])
    assert np.all(g[1] == 0) and ll[1] == 0 and np.all(g[2, 4:] == 0)
    g4, ll4 = fb_oracle.forward_backward_np(A, pi, lik[2, :4])
    assert np.array_equal(g[2, :4], g4) and ll[2] == ll4
