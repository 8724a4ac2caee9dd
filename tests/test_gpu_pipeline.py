"""The steps either side of the hot path on the GPU (vit_emissions_f32, vit_voiced_bins) against goldens produced by the
reference's own observation_probs_fn / __call__ (tests/golden/make_golden.py).  Peak patterns must be identical; values
within 1e-5 relative (GPU expf/logf vs NumPy's float32 routines -- this path is outside the bit-exact claim)."""
import os

import numpy as np
import pytest
import torch

from viterbi_spl_b200 import hmm_params

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
TINY = np.finfo(np.float32).tiny


@pytest.fixture(scope='module')
def pl(cuda_lib):
    assert torch.cuda.is_available()
    from viterbi_spl_b200 import pipeline
    return pipeline


def close(got, want):
    assert np.array_equal(got != 0, want != 0), 'peak pattern differs'
    assert np.allclose(got, want, rtol=1e-5, atol=1e-7), np.abs(got - want).max()


@pytest.mark.parametrize('scaled', [0, 1])
def test_softmax_model_matches_msnet_golden(pl, scaled):
    m = np.load(os.path.join(GOLD, 'msnet_softmax_viterbi.npz'))
    ml = np.load(os.path.join(GOLD, 'msnet_logdomain.npz'))
    logits = torch.as_tensor(m['logits'][None]).cuda()                      # [1, T, 1 + 320], column 0 = unvoiced
    prior = torch.as_tensor(np.roll(ml['ini_probs'], 1).copy()).cuda() if scaled else None
    p = pl.emissions_device(logits, 320, pl.SOFTMAX, prior, 5, 0.0, out_log=False)[0].cpu().numpy()
    close(p, m[f'prob_ts_{scaled}'])
    lp = pl.emissions_device(logits, 320, pl.SOFTMAX, prior, 5, 0.0, out_log=True)[0].cpu().numpy()
    assert np.allclose(lp, m[f'log_prob_ts_{scaled}'], rtol=0, atol=2e-5)
    # decode the GPU-built table: same (voiced, bins) as the reference's end-to-end __call__ on this recording
    from viterbi_spl_b200 import ViterbiDecoder
    st, _ = ViterbiDecoder(ml['logA_T'], ml['log_pi']).decode_device(torch.as_tensor(lp[None]).cuda())
    voiced, bins = pl.voiced_bins_device(st, 320)
    agree = np.mean((voiced[0].cpu().numpy() == m[f'voiced_{scaled}']) & (bins[0].cpu().numpy() == m[f'bins_{scaled}']))
    assert agree >= 0.99, agree


def test_shaun_model_matches_tonet_golden_and_full_pipeline(pl):
    b = np.load(os.path.join(GOLD, 'tonet_family_b.npz'))
    logits = torch.as_tensor(b['logits'][None]).cuda()                      # [1, T, 360]
    p = pl.emissions_device(logits, 360, pl.SHAUN, None, 5, 0.0, out_log=False)[0].cpu().numpy()      # threshold 0.5 -> logit 0
    close(p, b['probs_st'].T)
    assert np.allclose(p.sum(1), 1, atol=1e-5)
    mp = pl.MelodyPipeline(b['A'], b['pi'], model='shaun', voicing_threshold=0.5)
    voiced, bins = mp(b['logits'])
    agree = np.mean((voiced.cpu().numpy() == b['voiced']) & (bins.cpu().numpy() == b['bins']))
    assert agree >= 0.99, agree


@pytest.mark.parametrize('spw,n_bins', [(5, 64), (16, 100), (20, 721), (1, 3), (5, 7), (5, 13), (5, 360), (5, 383), (5, 384),
                                        (5, 385), (5, 6), (16, 721), (16, 756), (20, 757), (16, 19), (20, 23), (15, 360),
                                        (2, 9), (3, 50), (7, 700), (33, 400)])
def test_peak_picking_is_exact_with_ties_and_reflect_padding(pl, spw, n_bins):
    """Quantised logits (many equal neighbours) against the NumPy argmax-of-window rule of
    find_peaks_all_at_once_np_fn (dcnet/softmax_viterbi.py:2508-2528), restated here for arbitrary width."""
    rng = np.random.default_rng(spw * 1000 + n_bins)
    T = 50
    x = rng.integers(0, 4, size=(T, n_bins)).astype(np.float32)
    padded = np.pad(x, [(0, 0), (spw, spw)], mode='reflect')
    want = np.stack([np.argmax(padded[:, k:k + 2 * spw + 1], axis=1) == spw for k in range(n_bins)], axis=1)
    p = pl.emissions_device(torch.as_tensor(x[None]).cuda(), n_bins, pl.SHAUN, None, spw, 0.0, out_log=False)[0].cpu().numpy()
    assert np.array_equal(p[:, :n_bins] != 0, want)
    none = ~want.any(1)
    assert np.all(p[none, n_bins] == 1)                                     # no peak: E[unvoiced] = 1


@pytest.mark.parametrize('model,n_bins,scaled', [('softmax', 360, True), ('softmax', 320, False), ('shaun', 360, False),
                                                 ('shaun', 37, False), ('softmax', 384, True)])
def test_register_window_kernel_agrees_with_the_generic_kernel(pl, model, n_bins, scaled, monkeypatch):
    """spw = 5 takes the register-window kernel (12 bins per lane, cp.async ring, float4 rows at the alignment shift of
    every frame); VIT_EMIS_GENERIC keeps the generic one.  Same peaks exactly, values within the softmax-sum rounding."""
    g = torch.Generator(device='cuda'); g.manual_seed(n_bins)
    n_in = n_bins + 1 if model == 'softmax' else n_bins
    logits = 2 * torch.randn((3, 41, n_in), device='cuda', generator=g)
    logits[0, 3] = 0.0                                                      # a flat frame: no voiced peak at all
    logits[1, 5, n_in - 7:] = 9.0                                           # a plateau running into the reflected edge
    prior = None
    if scaled:
        prior = torch.rand((n_bins + 1,), device='cuda', generator=g) * 0.01 + 1e-3
    m = pl.SOFTMAX if model == 'softmax' else pl.SHAUN
    for out_log in (True, False):
        a = pl.emissions_device(logits, n_bins, m, prior, 5, 0.3, out_log=out_log)
        monkeypatch.setenv('VIT_EMIS_GENERIC', '1')
        b = pl.emissions_device(logits, n_bins, m, prior, 5, 0.3, out_log=out_log)
        monkeypatch.delenv('VIT_EMIS_GENERIC')
        zero = float(np.log(np.finfo(np.float32).tiny)) if out_log else 0.0
        assert torch.equal(a == zero, b == zero)
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-4 if out_log else 0.0)
        # an output table that is not 16-byte aligned cannot take float4 rows: the library falls back by itself
        flat = torch.empty((a.numel() + 1,), dtype=torch.float32, device='cuda')
        c = pl.emissions_device(logits, n_bins, m, prior, 5, 0.3, out_log=out_log, out=flat[1:].view(a.shape))
        assert torch.equal(c, b)
        # float4 rows at a per-frame alignment shift: nothing may land outside the table (canaries either side)
        big = torch.full((a.numel() + 8,), float('nan'), dtype=torch.float32, device='cuda')
        c2 = pl.emissions_device(logits, n_bins, m, prior, 5, 0.3, out_log=out_log, out=big[4:4 + a.numel()].view(a.shape))
        assert torch.equal(c2, a)
        assert bool(torch.isnan(big[:4]).all()) and bool(torch.isnan(big[-4:]).all())


@pytest.mark.parametrize('model,n_bins,spw,scaled', [('softmax', 721, 16, True), ('shaun', 721, 20, False), ('softmax', 721, 20, False),
                                                     ('shaun', 700, 16, False), ('softmax', 756, 16, True), ('shaun', 40, 20, False)])
def test_wide_register_window_kernel_agrees_with_the_generic_kernel(pl, model, n_bins, spw, scaled, monkeypatch):
    """Peak half-width 16 / 20 with up to 756 bins (jdc: 721 / 16, imm: 721 / 20) takes the wide register-window kernel (two
    passes of 12 bins per lane, window maxima by 3-input maxima in registers); VIT_EMIS_GENERIC keeps the generic one.
    Same peaks exactly, values within the softmax-sum rounding, nothing written outside the table."""
    g = torch.Generator(device='cuda'); g.manual_seed(n_bins + spw)
    n_in = n_bins + 1 if model == 'softmax' else n_bins
    logits = 2 * torch.randn((3, 37, n_in), device='cuda', generator=g)
    logits[0, 3] = 0.0                                                      # a flat frame: no voiced peak at all
    logits[1, 5, n_in - 9:] = 9.0                                           # a plateau running into the reflected edge
    logits[2, 7] = torch.round(logits[2, 7])                                # ties between neighbours
    prior = (torch.rand((n_bins + 1,), device='cuda', generator=g) * 0.01 + 1e-3) if scaled else None
    m = pl.SOFTMAX if model == 'softmax' else pl.SHAUN
    for out_log in (True, False):
        a = pl.emissions_device(logits, n_bins, m, prior, spw, 0.3, out_log=out_log)
        monkeypatch.setenv('VIT_EMIS_GENERIC', '1')
        b = pl.emissions_device(logits, n_bins, m, prior, spw, 0.3, out_log=out_log)
        monkeypatch.delenv('VIT_EMIS_GENERIC')
        zero = float(np.log(np.finfo(np.float32).tiny)) if out_log else 0.0
        assert torch.equal(a == zero, b == zero)
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-4 if out_log else 0.0)
        big = torch.full((a.numel() + 8,), float('nan'), dtype=torch.float32, device='cuda')
        c2 = pl.emissions_device(logits, n_bins, m, prior, spw, 0.3, out_log=out_log, out=big[4:4 + a.numel()].view(a.shape))
        assert torch.equal(c2, a)
        assert bool(torch.isnan(big[:4]).all()) and bool(torch.isnan(big[-4:]).all())
    # peaks against the NumPy window rule
    x = logits[..., 1:] if model == 'softmax' else logits
    x = x.cpu().numpy().reshape(-1, n_bins)
    padded = np.pad(x, [(0, 0), (spw, spw)], mode='reflect')
    want = np.lib.stride_tricks.sliding_window_view(padded, 2 * spw + 1, axis=1).argmax(-1) == spw
    got = pl.emissions_device(logits, n_bins, m, prior, spw, 0.3, out_log=False).cpu().numpy().reshape(-1, n_bins + 1)
    assert np.array_equal(got[:, :n_bins] != 0, want)


@pytest.mark.parametrize('case_name', ['jdc_SoftMaxViterbi_True', 'jdc_Viterbi_0p5', 'imm_Viterbi'])
def test_wide_kernel_against_goldens_made_by_the_reference_classes(pl, case_name):
    """Emission tables of the jdc / imm models against the reference's own observation_probs_fn (tests/golden/class_calls.npz)."""
    g = np.load(os.path.join(GOLD, 'class_calls.npz'))
    logits, want = g[f'{case_name}_logits'], g[f'{case_name}_probs']
    if case_name.startswith('jdc_SoftMax'):
        prior = torch.as_tensor(np.roll(g['pi_jdc'], 1).astype(np.float32).copy()).cuda()
        p = pl.emissions_device(torch.as_tensor(logits[None]).cuda(), 721, pl.SOFTMAX, prior, 16, 0.0, out_log=False)[0].cpu().numpy()
        close(p, want)
    elif case_name.startswith('jdc'):
        p = pl.emissions_device(torch.as_tensor(logits[None]).cuda(), 721, pl.SHAUN, None, 16, 0.0, out_log=False)[0].cpu().numpy()
        assert np.array_equal(p.T != 0, want != 0) and np.allclose(p.T, want, rtol=2e-5, atol=1e-7)
    else:
        x = np.ascontiguousarray(logits.T)                                     # imm hands logits over as [n_bins, T]
        p = pl.emissions_device(torch.as_tensor(x[None]).cuda(), 721, pl.SHAUN, None, 20, 2.442347, out_log=False)[0].cpu().numpy()
        assert np.array_equal(p.T != 0, want != 0) and np.allclose(p.T, want, rtol=2e-5, atol=1e-7)


def test_voiced_bins_and_batched_pipeline(pl):
    A, pi = hmm_params.synthetic_hmm('tonet')
    mp = pl.MelodyPipeline(A, pi, model='softmax', scaled=True)
    g = torch.Generator(device='cuda'); g.manual_seed(1)
    logits = 2 * torch.randn((5, 90, 361), device='cuda', generator=g)
    L = torch.as_tensor(np.asarray([90, 0, 17, 1, 90], np.int32)).cuda()
    voiced, bins = mp(logits, L)
    assert voiced.shape == (5, 90) and voiced.dtype == torch.bool and bins.dtype == torch.int64
    E = mp.emissions(logits)
    st, _ = mp.decoder.decode_device(E, L)
    st = st.cpu().numpy()
    assert np.array_equal(voiced.cpu().numpy(), (st >= 0) & (st < 360))
    assert np.array_equal(bins.cpu().numpy(), np.where(st < 0, -1, np.minimum(st, 359)))
    assert not voiced[1].any() and (bins[1] == -1).all()


# ---- post-decode statistics (vit_melody_stats_f32) vs the NumPy restatement of the reference's TF step -------------

def _stats_case(seed, T, n_bins=320):
    rng = np.random.default_rng(seed)
    logits = rng.normal(0, 3, (T, n_bins)).astype(np.float32)
    bins = rng.integers(0, n_bins, T)
    bins[:4] = np.asarray((0, n_bins - 1, 1, n_bins - 2))[:T]                                  # window clipped at both ends of the bin range
    voiced = rng.random(T) < 0.6
    ref = (23.6 + bins / 5. + rng.normal(0, 0.4, T)).astype(np.float32)        # around the decoded bin: hits and misses
    ref[rng.random(T) < 0.15] += 12.                                           # octave errors: chroma hit, pitch miss
    ref[rng.random(T) < 0.3] = 0.                                              # unvoiced reference frames
    return logits, ref, bins, voiced


def test_melody_stats_match_the_reference_restatement(pl):
    from oracle import post_oracle
    cases = [_stats_case(s, T) for s, T in ((1, 700), (2, 1), (3, 257), (4, 1200))]
    T_max = max(c[0].shape[0] for c in cases)
    B = len(cases)
    L = np.zeros((B, T_max, 320), np.float32)
    R = np.zeros((B, T_max), np.float32)
    Bn = np.full((B, T_max), -1, np.int64)
    V = np.zeros((B, T_max), bool)
    lengths = np.asarray([c[0].shape[0] for c in cases], np.int32)
    for b, (lg, ref, bins, voiced) in enumerate(cases):
        n = len(ref)
        L[b, :n], R[b, :n], Bn[b, :n], V[b, :n] = lg, ref, bins, voiced
    est, counters = pl.melody_stats_device(torch.as_tensor(L).cuda(), torch.as_tensor(R), torch.as_tensor(Bn),
                                           torch.as_tensor(V), lengths)
    est, counters = est.cpu().numpy(), counters.cpu().numpy()
    for b, (lg, ref, bins, voiced) in enumerate(cases):
        n = len(ref)
        want, c = post_oracle.melody_stats_np(ref, lg, bins, voiced)
        assert np.allclose(est[b, :n], want, rtol=1e-5, atol=1e-5), np.abs(est[b, :n] - want).max()
        assert np.all(est[b, n:] == 0)
        got = dict(zip(pl.COUNTER_NAMES, counters[b].tolist()))
        assert pl.COUNTER_NAMES == post_oracle.COUNTERS
        # a note within 1e-5 of a 0.5 threshold may fall on either side: allow that many frames of slack per counter
        diff = np.abs(np.abs(want) - ref)
        edge = int(np.sum(np.abs(diff - 0.5) < 2e-5) + np.sum(np.abs(np.abs(diff - np.round(diff / 12) * 12) - 0.5) < 2e-5))
        for k in pl.COUNTER_NAMES:
            assert abs(got[k] - c[k]) <= edge, (b, k, got[k], c[k])
        for k in pl.COUNTER_NAMES[:5]:
            assert got[k] == c[k], (b, k)
        assert got['gt_voiced'] + got['gt_unvoiced'] == n


def test_melody_stats_match_goldens_made_by_the_reference_function(pl):
    """vit_melody_stats_f32 against the outputs of the reference's own viterbi_update_states_tf_fn / est_notes_fn
    (tests/golden/melody_stats.npz, made by executing those functions on a NumPy-backed stand-in for their TensorFlow
    ops): notes to 1e-5, the five voicing counters exact, the four note-threshold counters up to the frames whose note
    lies within 2e-5 of a threshold."""
    g = np.load(os.path.join(GOLD, 'melody_stats.npz'))
    n_cases = int(g['n_cases'])
    T_max = max(len(g[f'c{k}_ref']) for k in range(n_cases))
    L = np.zeros((n_cases, T_max, 320), np.float32)
    R = np.zeros((n_cases, T_max), np.float32)
    Bn = np.full((n_cases, T_max), -1, np.int64)
    V = np.zeros((n_cases, T_max), bool)
    lengths = np.asarray([len(g[f'c{k}_ref']) for k in range(n_cases)], np.int32)
    for k in range(n_cases):
        n = lengths[k]
        L[k, :n], R[k, :n], Bn[k, :n], V[k, :n] = g[f'c{k}_logits'], g[f'c{k}_ref'], g[f'c{k}_bins'], g[f'c{k}_voiced']
    est, counters = pl.melody_stats_device(torch.as_tensor(L).cuda(), torch.as_tensor(R), torch.as_tensor(Bn),
                                           torch.as_tensor(V), lengths)
    est, counters = est.cpu().numpy(), counters.cpu().numpy()
    for k in range(n_cases):
        n = lengths[k]
        want = g[f'c{k}_est']
        assert np.allclose(est[k, :n], want, rtol=1e-5, atol=1e-5)
        diff = np.abs(np.abs(want) - g[f'c{k}_ref'])
        edge = int(np.sum(np.abs(diff - 0.5) < 2e-5) + np.sum(np.abs(np.abs(diff - np.round(diff / 12) * 12) - 0.5) < 2e-5))
        assert np.array_equal(counters[k, :5], g[f'c{k}_counters'][:5])
        assert np.all(np.abs(counters[k] - g[f'c{k}_counters']) <= edge)


def test_melody_stats_softmax_layout_and_pipeline_evaluate(pl):
    """Column 0 = unvoiced logit (offset 1), through MelodyPipeline.evaluate on the shipped msnet parameters."""
    from oracle import post_oracle
    ml = np.load(os.path.join(GOLD, 'msnet_logdomain.npz'))
    m = np.load(os.path.join(GOLD, 'msnet_softmax_viterbi.npz'))
    A = np.exp(ml['logA_T'].T.astype(np.float64)).astype(np.float32)
    mp = pl.MelodyPipeline(A, ml['ini_probs'], model='softmax', scaled=False)
    logits = m['logits']                                                        # [T, 321]
    T = logits.shape[0]
    rng = np.random.default_rng(5)
    ref = np.where(rng.random(T) < 0.5, 23.6 + rng.integers(0, 320, T) / 5., 0.).astype(np.float32)
    voiced, bins, est, counters = mp.evaluate(logits[None], ref[None])
    want, c = post_oracle.melody_stats_np(ref, logits[:, 1:], bins[0].cpu().numpy(), voiced[0].cpu().numpy())
    assert np.allclose(est[0].cpu().numpy(), want, rtol=1e-5, atol=1e-5)
    got = dict(zip(pl.COUNTER_NAMES, counters[0].tolist()))
    assert all(abs(got[k] - c[k]) <= 1 for k in pl.COUNTER_NAMES), (got, c)


@pytest.mark.parametrize('model,scaled', [('softmax', False), ('softmax', True), ('shaun', False)])
def test_pipeline_posteriors_match_the_float64_oracle_on_the_emission_likelihoods(pl, model, scaled):
    """logits -> observation_probs_fn likelihoods (GPU) -> forward-backward (GPU; the convolution kernels for this matrix)
    against the float64 oracle run on the same likelihoods.  Parity unpinned: the reference has no forward-backward."""
    from oracle import fb_oracle
    A, pi = hmm_params.synthetic_hmm('tonet')
    S = len(pi)
    rng = np.random.default_rng(17)
    B, T = 5, 60
    n_in = S if model == 'softmax' else S - 1
    logits = (2.0 * rng.standard_normal((B, T, n_in))).astype(np.float32)
    L = np.asarray([T, 1, 0, 33, T], np.int32)
    mp = pl.MelodyPipeline(A, pi, model=model, scaled=scaled)
    d_logits = torch.as_tensor(logits).cuda()
    lik = mp.emissions(d_logits, out_log=False).cpu().numpy()
    gamma, ll = mp.posteriors(logits, L)
    want_g, want_ll = fb_oracle.forward_backward_batch_np(A.astype(np.float32), pi.astype(np.float32), lik, L)
    assert np.abs(gamma.cpu().numpy() - want_g).max() <= 1e-4
    assert np.allclose(ll.cpu().numpy(), want_ll, rtol=1e-5, atol=1e-5)
    g1, l1 = mp.posteriors(logits[0])                                 # a single recording [T, *]
    assert g1.shape == (T, S) and np.abs(g1.cpu().numpy() - want_g[0]).max() <= 1e-4
