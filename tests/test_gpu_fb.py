"""Forward-backward kernel (vit_forward_backward_f32, through the C ABI) against the float64 oracle.

PARITY UNPINNED: the reference has no forward-backward code; the oracle (oracle/fb_oracle.py) is this repository's own
float64 restatement of the textbook recursion, self-validated by brute-force enumeration (tests/test_fb_oracle.py).
Tolerances are the north star's: 1e-4 absolute on gamma, 1e-5 relative on log L."""
import numpy as np
import pytest
import torch

from oracle import fb_oracle
from viterbi_spl_b200 import hmm_params, synth

pytestmark = pytest.mark.gpu
GAMMA_ATOL = 1e-4
LOGLIK_RTOL = 1e-5


@pytest.fixture(scope='module', params=['tc', 'simt'])
def FB(cuda_lib, request):
    """Both kernels behind vit_forward_backward_f32: the tcgen05 tensor-core one (the default where the shape fits) and the
    FFMA one (VIT_FB_IMPL=simt; also what S = 722 takes)."""
    import os
    assert torch.cuda.is_available()
    from viterbi_spl_b200 import ForwardBackward
    os.environ['VIT_FB_IMPL'] = request.param
    yield ForwardBackward
    os.environ.pop('VIT_FB_IMPL', None)


def random_hmm(S, rng, sparse=False):
    A = (rng.random((S, S)) ** 3).astype(np.float64)
    if sparse:
        A[rng.random((S, S)) < 0.6] = 0
        A[:, -1] = np.maximum(A[:, -1], 0.02)
    A /= A.sum(1, keepdims=True)
    pi = rng.random(S) + 0.01
    return A.astype(np.float32), (pi / pi.sum()).astype(np.float32)


def softmax_style_likelihoods(B, T, S, rng, scaled_by=None):
    """0-5 peaks per frame + the always-present unvoiced state, all other bins exactly 0
    (shape of SoftMaxViterbi.observation_probs_fn output, dcnet/softmax_viterbi.py:2530-2579)."""
    lik = np.zeros((B, T, S), np.float32)
    for b in range(B):
        for t in range(T):
            k = int(rng.integers(0, 6))
            idx = np.unique(np.append(rng.choice(S - 1, size=k, replace=False), S - 1))
            w = np.exp(2.0 * rng.standard_normal(len(idx)))
            w /= w.sum()
            if scaled_by is not None:
                w = w / scaled_by[idx]
            lik[b, t, idx] = w
    return lik


def check(FB, A, pi, lik, lengths=None):
    want_g, want_ll = fb_oracle.forward_backward_batch_np(A, pi, lik, lengths)
    g, ll = FB(A, pi).run_host(lik, lengths)
    assert g.dtype == np.float32 and g.shape == lik.shape
    err = np.abs(g - want_g).max()
    assert err <= GAMMA_ATOL, f'max |gamma error| {err}'
    assert np.allclose(ll, want_ll, rtol=LOGLIK_RTOL, atol=1e-5), (ll, want_ll)
    n = lik.shape[1] if lengths is None else None
    if n:
        assert np.allclose(g.sum(-1), 1, atol=1e-4)


@pytest.mark.parametrize('S,T,B', [(1, 1, 1), (2, 3, 1), (7, 5, 3), (32, 9, 15), (33, 40, 8), (97, 50, 5), (191, 20, 9),
                                   (192, 12, 6), (200, 64, 9), (321, 60, 17), (361, 80, 30), (383, 10, 3), (722, 24, 16)])
def test_dense_random_models(FB, S, T, B):
    rng = np.random.default_rng(S * 7 + T)
    A, pi = random_hmm(S, rng)
    lik = np.exp(2 * rng.standard_normal((B, T, S))).astype(np.float32)
    check(FB, A, pi, lik)


@pytest.mark.parametrize('state_set,scaled', [('dcnet', False), ('tonet', False), ('tonet', True)])
def test_real_state_sets_softmax_style_likelihoods_ragged(FB, state_set, scaled):
    A, pi = hmm_params.synthetic_hmm(state_set)
    S = len(pi)
    rng = np.random.default_rng(S)
    B, T = 20, 120
    lik = softmax_style_likelihoods(B, T, S, rng, scaled_by=pi if scaled else None)
    L = rng.integers(0, T + 1, size=B).astype(np.int32)
    L[:5] = [T, 1, 0, 2, T]
    check(FB, A.astype(np.float32), pi.astype(np.float32), lik, L)


def test_more_clips_than_one_wave_of_clusters(FB):
    """1500 ragged clips at S = 361: more than the 45 x 32 (tensor-core kernel) / 74 x 14 (FFMA kernel) clips that are
    co-resident, so clusters loop over several sub-batches."""
    A, pi = hmm_params.synthetic_hmm('tonet')
    rng = np.random.default_rng(11)
    B, T, S = 1500, 10, 361
    lik = np.exp(rng.standard_normal((B, T, S))).astype(np.float32)
    L = rng.integers(0, T + 1, size=B).astype(np.int32)
    L[::7] = T
    check(FB, A.astype(np.float32), pi.astype(np.float32), lik, L)


def test_long_clip_does_not_underflow(FB):
    A, pi = hmm_params.synthetic_hmm('tonet')
    rng = np.random.default_rng(3)
    lik = softmax_style_likelihoods(2, 3000, 361, rng) * np.float32(1e-3)      # log L ~ -3e4: unscaled fp32 would underflow
    check(FB, A.astype(np.float32), pi.astype(np.float32), lik)


def test_device_api_and_error_codes(FB, cuda_lib):
    import ctypes
    A, pi = random_hmm(50, np.random.default_rng(0), sparse=True)
    lik = torch.rand((4, 30, 50), device='cuda') + 0.05
    keep = lik.clone()
    fb = FB(A, pi)
    g, ll = fb.run_device(lik)
    assert g.is_cuda and ll.shape == (4,) and torch.equal(lik, keep)
    want_g, want_ll = fb_oracle.forward_backward_batch_np(A, pi, lik.cpu().numpy())
    assert np.abs(g.cpu().numpy() - want_g).max() <= GAMMA_ATOL
    n = ctypes.c_size_t(0)
    assert cuda_lib.vit_fb_workspace_bytes(4, 30, 50, ctypes.byref(n)) == 0 and n.value > 0
    assert cuda_lib.vit_fb_workspace_bytes(4, 30, 5000, ctypes.byref(n)) == -4
    assert cuda_lib.vit_fb_workspace_bytes(4, 0, 50, ctypes.byref(n)) == -1


def test_impossible_observation_sequence_gives_zero_gamma_not_nan(FB):
    """A clip whose likelihoods make every path impossible from frame 7 on (the library's own emission builder writes
    exact zeros for non-peak bins, and off-band transitions are exactly 0): the normaliser reaches 0.  gamma must be 0 for
    that clip (never NaN), log L = -inf, and the other clips of the batch are unaffected."""
    S, T = 40, 16
    rng = np.random.default_rng(5)
    A = np.zeros((S, S), np.float32)
    for d in (-1, 0, 1):
        i = np.arange(max(0, -d), min(S, S - d))
        A[i, i + d] = 1.0
    A /= A.sum(1, keepdims=True)                                  # band +-1: state 0 cannot reach state 30 in one step
    pi = np.full(S, 1.0 / S, np.float32)
    lik = (rng.random((3, T, S)) + 0.1).astype(np.float32)
    lik[1, 6] = 0
    lik[1, 6, 0] = 1.0                                            # frame 6: only state 0 ...
    lik[1, 7] = 0
    lik[1, 7, 30] = 1.0                                           # ... frame 7: only state 30 -> probability 0
    g, ll = FB(A, pi).run_host(lik)
    assert not np.isnan(g).any()
    assert np.all(g[1] == 0) and ll[1] == -np.inf
    want_g, want_ll = fb_oracle.forward_backward_batch_np(A, pi, lik[[0, 2]])
    assert np.abs(g[[0, 2]] - want_g).max() <= GAMMA_ATOL and np.allclose(ll[[0, 2]], want_ll, rtol=LOGLIK_RTOL)
