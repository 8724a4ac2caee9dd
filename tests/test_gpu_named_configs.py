"""BASELINE.json's configurations at their NAMED sizes against the CPU oracle:

  cfg 2  1024 clips x 3000 frames x 361 states    64 clips vs the C oracle, tensor-memory AND banded kernels
  cfg 3  722 states x 10,000 frames                8 clips vs the C oracle: streaming kernel (logA^T through the TMA ring),
                                                   wide banded kernel (jdc matrix), fully dense imm matrix
  cfg 4  forward-backward 1024 x 3000 x 361        gamma <= 1e-4 / log L <= 1e-5 on 8 clips vs the float64 oracle (tcgen05)
  cfg 5  one 1,000,000-frame sequence              path + score vs the C oracle (its long-sequence, multi-threaded form)

Oracle clips are generated on the host with NumPy (seeded) and planted into a batch whose other clips are generated on
the device, so the kernels run the full-size launch the benchmark times."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, fb_oracle
from viterbi_spl_b200 import hmm_params, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def Decoder(cuda_lib):
    assert torch.cuda.is_available(), 'gpu-marked tests need a CUDA device'
    from viterbi_spl_b200 import ViterbiDecoder
    return ViterbiDecoder


def planted_batch(B, T, S, where, seed, dev):
    """[B, T, S] device batch; clips `where` are host-generated (returned too, for the oracle)."""
    E = synth.device_dense_softmax(B, T, S, seed=seed, device=dev)
    host = np.stack([synth.dense_softmax(T, S, seed=10_000 + seed + b) if k % 2 == 0 else synth.sparse_peaks(T, S, seed=10_000 + seed + b)
                     for k, b in enumerate(where)])
    E[torch.as_tensor(where, device=dev)] = torch.as_tensor(host).to(dev)
    return E, host


def test_cfg2_1024x3000x361_64_clips_against_the_oracle(Decoder):
    B, T, S = 1024, 3000, 361
    A, pi = hmm_params.synthetic_hmm('tonet')
    logA_T, log_pi = hmm_params.log_params(A, pi)
    dev = torch.device('cuda')
    where = list(range(0, B, 16))                                  # 64 clips spread over all clusters / CTAs
    E, host = planted_batch(B, T, S, where, seed=7, dev=dev)
    want_p, want_s = c_oracle.decode_batch_c(logA_T, log_pi, host)
    for algo in ('tmem', 'banded'):
        p, s = Decoder(logA_T, log_pi, algo=algo).decode_device(E)
        assert np.array_equal(p[where].cpu().numpy(), want_p), algo
        assert np.array_equal(s[where].cpu().numpy(), want_s), algo
        del p, s
    torch.cuda.empty_cache()


@pytest.mark.parametrize('model,algos', [('jdc', ('stream', 'auto')), ('imm', ('stream',)), ('imm_hmm', ('auto',))])
def test_cfg3_722_states_10000_frames_8_clips_against_the_oracle(Decoder, model, algos):
    """S = 722, T = 10,000: the streaming kernel (what `auto` takes for a dense matrix once the batch fills its pass)
    and the wide banded kernel (what `auto` takes for the jdc +-40 and imm-HMM +-56 band matrices)."""
    from viterbi_spl_b200 import _lib
    B, T, S = 300, 10_000, 722
    A, pi = hmm_params.synthetic_hmm(model)
    logA_T, log_pi = hmm_params.log_params(A, pi, add_tiny=(model != 'imm'))
    dev = torch.device('cuda')
    where = [0, 13, 14, 99, 150, 200, 285, 299]                     # 8 clips, both pipelines of several CTAs
    E, host = planted_batch(B, T, S, where, seed=3, dev=dev)
    want_p, want_s = c_oracle.decode_batch_c(logA_T, log_pi, host)
    for algo in algos:
        dec = Decoder(logA_T, log_pi, algo=algo)
        if algo == 'auto':
            assert dec.structure.kind == 1                           # the banded fast path is what runs
        p, s = dec.decode_device(E)
        assert np.array_equal(p[where].cpu().numpy(), want_p), (model, algo)
        assert np.array_equal(s[where].cpu().numpy(), want_s), (model, algo)
        del dec, p, s
    assert _lib.load().vit_select_algo(4096, T, S) == _lib.ALGO_STREAM
    torch.cuda.empty_cache()


@pytest.mark.parametrize('impl', ['tc', 'banded'])
def test_cfg4_forward_backward_1024x3000x361_8_clips_against_the_float64_oracle(cuda_lib, impl):
    """Config 4 at its named size with the tcgen05 kernel (what the config names) and with the banded kernel (what
    `auto` takes for this matrix: band +-14 + the unvoiced state, exact zeros elsewhere)."""
    from viterbi_spl_b200 import ForwardBackward
    os.environ.pop('VIT_FB_IMPL', None)
    B, T, S = 1024, 3000, 361
    A, pi = hmm_params.synthetic_hmm('tonet')
    A, pi = A.astype(np.float32), pi.astype(np.float32)
    dev = torch.device('cuda')
    g = torch.Generator(device=dev)
    g.manual_seed(4)
    lik = torch.softmax(2.0 * torch.randn((B, T, S), device=dev, generator=g), dim=-1)
    where = [0, 15, 16, 31, 500, 777, 1000, 1023]
    rng = np.random.default_rng(8)
    host = np.zeros((len(where), T, S), np.float32)
    for k in range(len(where)):
        if k % 2 == 0:                                               # dense softmax likelihoods
            x = 2.0 * rng.standard_normal((T, S))
            host[k] = (np.exp(x - x.max(1, keepdims=True)) / np.exp(x - x.max(1, keepdims=True)).sum(1, keepdims=True)).astype(np.float32)
        else:                                                        # SoftMaxViterbi-style: a few peaks + unvoiced, divided by the prior
            idx = rng.integers(0, S - 1, size=(T, 4))
            w = np.exp(2.0 * rng.standard_normal((T, 5)))
            w /= w.sum(1, keepdims=True)
            np.put_along_axis(host[k], idx, (w[:, :4] / pi[idx]).astype(np.float32), axis=1)
            host[k][:, S - 1] = (w[:, 4] / pi[S - 1]).astype(np.float32)
    lik[torch.as_tensor(where, device=dev)] = torch.as_tensor(host).to(dev)
    fb = ForwardBackward(A, pi, impl=impl)
    assert fb.structured
    gamma, ll = fb.run_device(lik)
    got_g, got_l = gamma[where].cpu().numpy(), ll[where].cpu().numpy()
    want_g, want_l = fb_oracle.forward_backward_batch_np(A, pi, host)
    assert np.abs(got_g - want_g).max() <= 1e-4
    assert np.allclose(got_l, want_l, rtol=1e-5, atol=0)
    assert torch.allclose(gamma.sum(-1), torch.ones((B, T), device=dev), atol=1e-4)
    del gamma, lik
    torch.cuda.empty_cache()


def test_cfg5_single_1m_frame_sequence_against_the_oracle(Decoder):
    """B = 1, T = 1,000,000, S = 361 (the latency case of config 5) through the host API with `auto` (banded kernel,
    time-parallel backtrace over 7813 segments): path AND score bit-equal to the C oracle."""
    T, S = 1_000_000, 361
    A, pi = hmm_params.synthetic_hmm('tonet')
    logA_T, log_pi = hmm_params.log_params(A, pi)
    E = np.concatenate([synth.dense_softmax(250_000, S, seed=50 + k) for k in range(4)])
    E[300_000:400_000] = synth.sparse_peaks(100_000, S, seed=77)      # a stretch of exact log(tiny) entries
    paths, scores = Decoder(logA_T, log_pi, algo='auto').decode_host(E[None])
    want_p, want_s = c_oracle.viterbi_log_long_c(logA_T, log_pi, E)
    assert np.array_equal(paths[0], want_p) and scores[0] == want_s
