"""Banded forward-backward kernel (VIT_FB_BANDED behind vit_forward_backward_f32_ex, csrc/vit_fb_banded.cu) against the
float64 oracle, and against the dense kernels on the same inputs.

PARITY UNPINNED: the reference has no forward-backward code (oracle/fb_oracle.py defines the semantics).  The matrices
are the ones the reference's builders produce (dcnet/viterbi_transition_matrix.py:81-98: a band of +-d_max bins inside a
voiced/unvoiced switch, exact zeros elsewhere), which is the structure the kernel exploits.
Tolerances are the north star's: 1e-4 absolute on gamma, 1e-5 relative on log L."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import fb_oracle
from viterbi_spl_b200 import hmm_params

pytestmark = pytest.mark.gpu
GAMMA_ATOL = 1e-4
LOGLIK_RTOL = 1e-5


@pytest.fixture(scope='module')
def FB(cuda_lib):
    assert torch.cuda.is_available()
    os.environ.pop('VIT_FB_IMPL', None)
    from viterbi_spl_b200 import ForwardBackward
    return ForwardBackward


def banded_hmm(S, d, dense, rng, zero_fraction=0.0):
    """Row-stochastic [S, S]: band |i - j| <= d among the other states + one dense state (row and column), or no dense
    state (dense = None).  zero_fraction knocks out entries inside the band as well."""
    A = np.zeros((S, S))
    i, j = np.indices((S, S))
    band = np.abs(i - j) <= d
    A[band] = rng.random(band.sum()) ** 2 + 1e-3
    if zero_fraction:
        A[band & (rng.random((S, S)) < zero_fraction) & (i != j)] = 0
    if dense is not None:
        A[dense, :] = rng.random(S) * 0.01 + 1e-4
        A[:, dense] = rng.random(S) * 0.05 + 1e-3
    A /= A.sum(1, keepdims=True)
    pi = rng.random(S) + 0.01
    return A.astype(np.float32), (pi / pi.sum()).astype(np.float32)


def peaky_likelihoods(B, T, S, rng):
    """0-5 peaks per frame + the last state, other bins exactly 0 (SoftMaxViterbi.observation_probs_fn shape,
    dcnet/softmax_viterbi.py:2530-2579)."""
    lik = np.zeros((B, T, S), np.float32)
    for b in range(B):
        for t in range(T):
            k = int(rng.integers(0, 6))
            idx = np.unique(np.append(rng.choice(S - 1, size=min(k, S - 1), replace=False), S - 1))
            w = np.exp(2.0 * rng.standard_normal(len(idx)))
            lik[b, t, idx] = w / w.sum()
    return lik


def check(FB, A, pi, lik, lengths=None, impl='banded'):
    want_g, want_ll = fb_oracle.forward_backward_batch_np(A, pi, lik, lengths)
    fb = FB(A, pi, impl=impl)
    assert fb.structured
    g, ll = fb.run_host(lik, lengths)
    assert g.dtype == np.float32 and g.shape == lik.shape
    assert not np.isnan(g).any()
    err = np.abs(g - want_g).max()
    assert err <= GAMMA_ATOL, f'max |gamma error| {err}'
    assert np.allclose(ll, want_ll, rtol=LOGLIK_RTOL, atol=1e-5), (ll, want_ll)
    if lengths is None:
        assert np.allclose(g.sum(-1), 1, atol=1e-4)
    else:
        for b, n in enumerate(lengths):
            assert np.all(g[b, n:] == 0)
    return g, ll


@pytest.mark.parametrize('S,d,dense,T,B', [
    (3, 1, None, 3, 1), (5, 1, 4, 7, 3), (16, 2, 0, 9, 4), (33, 4, 32, 40, 9), (64, 3, None, 20, 8), (97, 5, 40, 50, 5),
    (127, 8, 126, 30, 7), (128, 7, 0, 12, 16), (200, 12, 199, 64, 9), (255, 11, 100, 10, 3), (321, 12, 320, 60, 17),
    (361, 14, 360, 80, 30), (383, 14, 382, 10, 3), (384, 13, None, 10, 5), (384, 14, 383, 6, 11)])
def test_random_banded_models(FB, S, d, dense, T, B):
    rng = np.random.default_rng(S * 31 + d)
    A, pi = banded_hmm(S, d, dense, rng, zero_fraction=0.2 if S > 60 else 0.0)
    lik = np.exp(2 * rng.standard_normal((B, T, S))).astype(np.float32)
    check(FB, A, pi, lik)


@pytest.mark.parametrize('state_set', ['dcnet', 'tonet'])
def test_reference_state_sets_ragged(FB, state_set):
    A, pi = hmm_params.synthetic_hmm(state_set)
    S = len(pi)
    rng = np.random.default_rng(S)
    B, T = 20, 120
    lik = peaky_likelihoods(B, T, S, rng)
    L = rng.integers(0, T + 1, size=B).astype(np.int32)
    L[:5] = [T, 1, 0, 2, T]
    check(FB, A.astype(np.float32), pi.astype(np.float32), lik, L)


def test_auto_takes_the_banded_kernel_and_agrees_with_the_dense_ones(FB, cuda_lib):
    from viterbi_spl_b200 import _lib
    A, pi = hmm_params.synthetic_hmm('tonet')
    A, pi = A.astype(np.float32), pi.astype(np.float32)
    rng = np.random.default_rng(1)
    lik = np.exp(rng.standard_normal((9, 50, 361))).astype(np.float32)
    n0 = _lib.launch_count()
    g_auto, ll_auto = FB(A, pi).run_host(lik)
    # form check + two convolution passes + their log L + two general banded passes and their log L (which return at once)
    assert _lib.launch_count() - n0 == 7
    g_b, ll_b = FB(A, pi, impl='banded').run_host(lik)
    assert np.array_equal(g_auto, g_b) and np.array_equal(ll_auto, ll_b)
    for impl in ('tc', 'simt'):
        g_d, ll_d = FB(A, pi, impl=impl).run_host(lik)
        assert np.abs(g_d - g_b).max() <= GAMMA_ATOL
        assert np.allclose(ll_d, ll_b, rtol=LOGLIK_RTOL)


def toeplitz_hmm(n_bins, d, dense, rng, floor=2):
    """The reference's recipe (dcnet/viterbi_transition_matrix.py:60-98): one jump histogram on every row of the band,
    rows normalised, embedded in a voiced/unvoiced switch; dense = 'last' | 'first' | None."""
    hist = np.maximum(np.floor(2e5 * np.exp(-np.abs(np.arange(-d, d + 1)) / 1.3) * (1 + 0.1 * rng.random(2 * d + 1))), floor)
    hist = hist / hist.sum()
    T = np.zeros((n_bins, n_bins), np.float32)
    for i in range(n_bins):
        for j in range(max(0, i - d), min(n_bins, i + d + 1)):
            T[i, j] = hist[j - i + d]
    T = T / np.sum(T, axis=1)[:, None]
    if dense is None:
        A = T
    else:
        sw = np.asarray([[0.9779, 0.0221], [0.0172, 0.9828]], np.float32)
        A = np.zeros((n_bins + 1, n_bins + 1), np.float32)
        v = slice(0, n_bins) if dense == 'last' else slice(1, n_bins + 1)
        u = n_bins if dense == 'last' else 0
        A[v, v] = T * sw[0, 0]
        A[v, u] = sw[0, 1]
        A[u, v] = sw[1, 0] / n_bins
        A[u, u] = sw[1, 1]
    S = A.shape[0]
    pi = rng.random(S) + 0.01
    return A.astype(np.float32), (pi / pi.sum()).astype(np.float32)


def run_both_forms(FB, A, pi, lik, lengths=None):
    """The same call with the convolution kernels allowed (default) and switched off (VIT_FB_CONV=0: general banded
    kernels); both against the oracle.  Returns whether the two runs differ in any bit (= different kernels ran)."""
    g1, l1 = check(FB, A, pi, lik, lengths)
    os.environ['VIT_FB_CONV'] = '0'
    try:
        g0, l0 = check(FB, A, pi, lik, lengths)
    finally:
        os.environ.pop('VIT_FB_CONV', None)
    assert np.abs(g1 - g0).max() <= 2e-5
    return not np.array_equal(g1, g0)


@pytest.mark.parametrize('n_bins,d,dense,T,B', [
    (360, 14, 'last', 90, 21), (320, 12, 'last', 70, 9), (360, 13, 'last', 33, 5), (383, 14, 'last', 20, 4),
    (384, 14, None, 20, 4), (200, 8, 'first', 40, 7), (120, 4, 'last', 50, 6), (61, 3, None, 25, 3), (40, 14, 'last', 30, 5)])
def test_scaled_toeplitz_matrices_take_the_convolution_kernels(FB, n_bins, d, dense, T, B):
    rng = np.random.default_rng(n_bins + d)
    A, pi = toeplitz_hmm(n_bins, d, dense, rng)
    S = len(pi)
    lik = peaky_likelihoods(B, T, S, rng) if d % 2 else np.exp(2 * rng.standard_normal((B, T, S))).astype(np.float32)
    L = rng.integers(0, T + 1, size=B).astype(np.int32)
    L[:3] = [T, 1, 0][:min(3, B)]
    assert run_both_forms(FB, A, pi, lik, L)


@pytest.mark.parametrize('n_bins,d,dense,T,B', [(721, 40, 'last', 24, 9), (721, 56, 'last', 18, 5), (721, 33, 'last', 12, 3),
                                                 (600, 20, None, 15, 4), (767, 28, 'last', 10, 3), (499, 40, 'first', 12, 4)])
def test_wide_bands_722_state_sets_take_the_wide_convolution_kernels(FB, n_bins, d, dense, T, B):
    """jdc (+-40 of 721 bins) and the imm HMM (+-56): two runs of 12 states per lane, taps in shared memory.  With the
    form check switched off the same call goes to the dense FFMA kernel."""
    rng = np.random.default_rng(n_bins + d)
    A, pi = toeplitz_hmm(n_bins, d, dense, rng)
    S = len(pi)
    lik = peaky_likelihoods(B, T, S, rng) if d == 40 else np.exp(2 * rng.standard_normal((B, T, S))).astype(np.float32)
    L = rng.integers(0, T + 1, size=B).astype(np.int32)
    L[:3] = [T, 1, 0]
    assert run_both_forms(FB, A, pi, lik, L)


def test_wide_general_band_falls_back_to_the_dense_kernel(FB):
    rng = np.random.default_rng(12)
    A, pi = banded_hmm(722, 30, 721, rng)
    lik = np.exp(rng.standard_normal((4, 12, 722))).astype(np.float32)
    assert not run_both_forms(FB, A, pi, lik)


def test_general_band_does_not_take_the_convolution_kernels(FB):
    rng = np.random.default_rng(9)
    A, pi = banded_hmm(200, 9, 199, rng)
    lik = np.exp(rng.standard_normal((5, 40, 200))).astype(np.float32)
    assert not run_both_forms(FB, A, pi, lik)


def test_convolution_form_long_clips_1024_clip_batch(FB):
    """More clips than are co-resident (148 SMs x 15 one-warp blocks), 400 frames, the tonet state set."""
    A, pi = hmm_params.synthetic_hmm('tonet')
    A, pi = A.astype(np.float32), pi.astype(np.float32)
    dev = torch.device('cuda')
    B, T, S = 2500, 400, 361
    g = torch.Generator(device=dev)
    g.manual_seed(6)
    lik = torch.softmax(2.0 * torch.randn((B, T, S), device=dev, generator=g), dim=-1)
    gamma, ll = FB(A, pi).run_device(lik)
    assert torch.allclose(gamma.sum(-1), torch.ones((B, T), device=dev), atol=1e-4)
    sub = [0, 1, 1234, 2499]
    want_g, want_ll = fb_oracle.forward_backward_batch_np(A, pi, lik[sub].cpu().numpy())
    assert np.abs(gamma[sub].cpu().numpy() - want_g).max() <= GAMMA_ATOL
    assert np.allclose(ll[sub].cpu().numpy(), want_ll, rtol=LOGLIK_RTOL)


def test_dense_matrix_is_refused_by_the_banded_kernel_and_auto_falls_back(FB):
    from viterbi_spl_b200._lib import VitError
    rng = np.random.default_rng(0)
    A = rng.random((50, 50)).astype(np.float32)
    A /= A.sum(1, keepdims=True)
    pi = np.full(50, 0.02, np.float32)
    lik = (rng.random((3, 8, 50)) + 0.1).astype(np.float32)
    fb = FB(A, pi, impl='banded')
    assert not fb.structured
    with pytest.raises(VitError):
        fb.run_host(lik)
    want_g, _ = fb_oracle.forward_backward_batch_np(A, pi, lik)
    g, _ = FB(A, pi).run_host(lik)
    assert np.abs(g - want_g).max() <= GAMMA_ATOL


def test_more_clips_than_one_pass_of_the_grid(FB):
    """1500 ragged clips at S = 361: more than the 148 x 8 co-resident ones, so the CTAs loop."""
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
    lik = peaky_likelihoods(2, 3000, 361, rng) * np.float32(1e-3)
    check(FB, A.astype(np.float32), pi.astype(np.float32), lik)


def test_impossible_observation_sequence_gives_zero_gamma_not_nan(FB):
    S, T = 40, 16
    rng = np.random.default_rng(5)
    A = np.zeros((S, S), np.float32)
    for d in (-1, 0, 1):
        i = np.arange(max(0, -d), min(S, S - d))
        A[i, i + d] = 1.0
    A /= A.sum(1, keepdims=True)
    pi = np.full(S, 1.0 / S, np.float32)
    lik = (rng.random((3, T, S)) + 0.1).astype(np.float32)
    lik[1, 6] = 0
    lik[1, 6, 0] = 1.0
    lik[1, 7] = 0
    lik[1, 7, 30] = 1.0
    fb = FB(A, pi, impl='banded')
    g, ll = fb.run_host(lik)
    assert not np.isnan(g).any()
    assert np.all(g[1] == 0) and ll[1] == -np.inf
    want_g, want_ll = fb_oracle.forward_backward_batch_np(A, pi, lik[[0, 2]])
    assert np.abs(g[[0, 2]] - want_g).max() <= GAMMA_ATOL and np.allclose(ll[[0, 2]], want_ll, rtol=LOGLIK_RTOL)


def test_c_abi_directly(cuda_lib):
    """vit_forward_backward_f32_ex through ctypes with an explicit vit_fb_opts; inputs are not modified."""
    from viterbi_spl_b200 import _lib
    rng = np.random.default_rng(2)
    A, pi = banded_hmm(100, 6, 99, rng)
    st = _lib.analyze_structure(A)
    assert st.kind == 1 and st.halfwidth == 6 and st.dense_index == 99 and st.background == 0.0
    B, T, S = 5, 33, 100
    lik = torch.rand((B, T, S), device='cuda') + 0.05
    keep = lik.clone()
    dA, dpi = torch.as_tensor(A).cuda(), torch.as_tensor(pi).cuda()
    n = ctypes.c_size_t(0)
    assert cuda_lib.vit_fb_workspace_bytes(B, T, S, ctypes.byref(n)) == 0
    ws = torch.empty(n.value, dtype=torch.uint8, device='cuda')
    gamma = torch.full((B, T, S), float('nan'), device='cuda')
    ll = torch.empty(B, device='cuda')
    opts = _lib.FbOpts()
    opts.impl = _lib.FB_BANDED
    opts.structure = ctypes.pointer(st)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = cuda_lib.vit_forward_backward_f32_ex(p(dA), p(dpi), p(lik), None, B, T, S, p(ws), n.value, p(gamma), p(ll),
                                               ctypes.byref(opts), None)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(lik, keep)
    want_g, want_ll = fb_oracle.forward_backward_batch_np(A, pi, lik.cpu().numpy())
    assert np.abs(gamma.cpu().numpy() - want_g).max() <= GAMMA_ATOL
    assert np.allclose(ll.cpu().numpy(), want_ll, rtol=LOGLIK_RTOL)
    # a structure with a non-zero background is refused
    st2 = _lib.Structure(1, 6, 99, -87.0, 0.0)
    opts.structure = ctypes.pointer(st2)
    assert cuda_lib.vit_forward_backward_f32_ex(p(dA), p(dpi), p(lik), None, B, T, S, p(ws), n.value, p(gamma), p(ll),
                                                ctypes.byref(opts), None) == -4


@pytest.mark.parametrize('case', ['conv_narrow', 'conv_wide', 'general', 'wide_fallback'])
def test_no_write_outside_gamma_loglik_and_workspace(cuda_lib, case):
    """compute-sanitizer is not available on the pool: canaries either side of every output buffer instead.  gamma, log L
    and the workspace are carved out of larger allocations filled with a sentinel; after the call the guards are intact."""
    from viterbi_spl_b200 import _lib
    rng = np.random.default_rng(31)
    if case == 'conv_narrow':
        A, pi = [x.astype(np.float32) for x in hmm_params.synthetic_hmm('tonet')]
    elif case == 'conv_wide':
        A, pi = [x.astype(np.float32) for x in hmm_params.synthetic_hmm('jdc')]
    elif case == 'general':
        A, pi = banded_hmm(361, 14, 360, rng)
    else:
        A, pi = banded_hmm(722, 30, 721, rng)
    S = len(pi)
    B, T = 7, 9
    st = _lib.analyze_structure(A)
    assert st.kind == 1 and st.background == 0.0
    lik_h = np.exp(rng.standard_normal((B, T, S))).astype(np.float32)
    L_h = np.asarray([T, 1, 0, 5, T, 2, 8], np.int32)
    G = 4096
    SENT = 12345.5
    n = ctypes.c_size_t(0)
    assert cuda_lib.vit_fb_workspace_bytes(B, T, S, ctypes.byref(n)) == 0
    ws_bytes = (n.value + 255) // 256 * 256
    big_g = torch.full((G + B * T * S + G,), SENT, device='cuda')
    big_l = torch.full((G + B + G,), SENT, device='cuda')
    big_w = torch.full((1024 + ws_bytes // 4 + G,), SENT, device='cuda')          # the workspace must be 256-byte aligned
    w_off = (-(big_w.data_ptr() + 1024 * 4) % 256) // 4 + 1024
    lik, L = torch.as_tensor(lik_h).cuda(), torch.as_tensor(L_h).cuda()
    dA, dpi = torch.as_tensor(A).cuda(), torch.as_tensor(pi).cuda()
    opts = _lib.FbOpts()
    opts.impl = _lib.FB_BANDED
    opts.structure = ctypes.pointer(st)
    vp = ctypes.c_void_p
    rc = cuda_lib.vit_forward_backward_f32_ex(vp(dA.data_ptr()), vp(dpi.data_ptr()), vp(lik.data_ptr()), vp(L.data_ptr()), B, T, S,
                                               vp(big_w.data_ptr() + 4 * w_off), ws_bytes, vp(big_g.data_ptr() + 4 * G),
                                               vp(big_l.data_ptr() + 4 * G), ctypes.byref(opts), None)
    assert rc == 0
    torch.cuda.synchronize()
    assert bool((big_g[:G] == SENT).all()) and bool((big_g[G + B * T * S:] == SENT).all())
    assert bool((big_l[:G] == SENT).all()) and bool((big_l[G + B:] == SENT).all())
    assert bool((big_w[:w_off] == SENT).all()) and bool((big_w[w_off + ws_bytes // 4:] == SENT).all())
    g = big_g[G:G + B * T * S].reshape(B, T, S).cpu().numpy()
    want_g, want_ll = fb_oracle.forward_backward_batch_np(A, pi, lik_h, L_h)
    assert np.abs(g - want_g).max() <= GAMMA_ATOL
    assert np.allclose(big_l[G:G + B].cpu().numpy(), want_ll, rtol=LOGLIK_RTOL, atol=1e-5)
