"""Host-side logic that needs no GPU: input validation of the reference entry points (same AssertionErrors as the
reference), sharding arithmetic, synthetic-input determinism, and the rule that the product never touches the oracle."""
import os
import re

import numpy as np
import pytest
import torch

from viterbi_spl_b200 import reference_api as api
from viterbi_spl_b200 import sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'viterbi_spl_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'libvit_oracle' not in text, f


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_decoder_fails_loudly_without_a_gpu():
    from viterbi_spl_b200 import ViterbiDecoder
    A, pi = synth.dyadic_hmm(5, seed=0)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        ViterbiDecoder(A, pi)


def test_log_domain_entry_point_asserts_like_the_reference():
    A, pi = synth.dyadic_hmm(6, seed=0)
    E = np.require(synth.dyadic((9, 6), 1).T, requirements=['C'])
    with pytest.raises(AssertionError):      # imm/tf_viterbi.py:78  C-contiguity
        api.viterbi_librosa_fn(log_transition_matrix_T=np.asfortranarray(A), log_prob_init=pi, log_probs_st=E)
    with pytest.raises(AssertionError):      # :79 dtype
        api.viterbi_librosa_fn(log_transition_matrix_T=A.astype(np.float64), log_prob_init=pi, log_probs_st=E)
    with pytest.raises(AssertionError):      # :85 shape
        api.viterbi_librosa_fn(log_transition_matrix_T=A, log_prob_init=pi, log_probs_st=E[:5])
    with pytest.raises(AssertionError):      # :82 len(log_prob_init)
        api.viterbi_librosa_fn(log_transition_matrix_T=A, log_prob_init=pi[:5], log_probs_st=E)
    with pytest.raises(TypeError):           # keyword-only, like the reference signature
        api.viterbi_librosa_fn(A, pi, E)
    with pytest.raises(AssertionError):      # imm/tf_viterbi.py:28-29 shape of the emissions
        api.tf_viterbi_librosa_fn(tf_log_transition_matrix_T=A, tf_log_prob_init=pi, tf_or_np_log_probs_st=E[:5])
    with pytest.raises(AssertionError):      # :26 len(prob_init)
        api.tf_viterbi_librosa_fn(tf_log_transition_matrix_T=A, tf_log_prob_init=pi[:5], tf_or_np_log_probs_st=E)
    with pytest.raises(TypeError):
        api.tf_viterbi_librosa_fn(A, pi, E)


def test_family_a_asserts_like_the_reference():
    rng = np.random.default_rng(0)
    A = rng.random((5, 5)).astype(np.float32)
    pi = np.full(5, 0.2, np.float32)
    probs = rng.random((5, 7)).astype(np.float32)
    with pytest.raises(AssertionError):      # rows must sum to 1 (dcnet/softmax_viterbi.py:2454-2455)
        api.Viterbi.viterbi_librosa_fn(transition_matrix=A, prob_init=pi, probs_st=probs)
    A = A / A.sum(1, keepdims=True)
    with pytest.raises(AssertionError):      # sum(prob_init) == 1 (:2457)
        api.viterbi_librosa_c_fn(transition_matrix=A, prob_init=pi * 2, probs_st=probs)
    with pytest.raises(AssertionError):      # probs shape (:2452)
        api.viterbi_numba_fn(transition_matrix=A, prob_init=pi, probs_st=probs[:4])


def test_shard_bounds_partition():
    for B in (0, 1, 7, 1024, 65536):
        for W in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(B, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_synthetic_inputs_are_deterministic_and_finite():
    a = synth.dense_softmax(50, 33, seed=4)
    b = synth.dense_softmax(50, 33, seed=4)
    assert np.array_equal(a, b) and a.dtype == np.float32 and np.isfinite(a).all()
    s = synth.sparse_peaks(40, 33, seed=5)
    assert np.isclose(s.min(), synth.LOG_TINY) and np.isfinite(s).all()
    t = synth.tie_stress((30, 9), seed=1)
    assert set(np.unique(t)).issubset({-0.25 * k for k in range(8)})
    assert np.array_equal(synth.batch('dyadic', 3, 8, 5, seed0=2)[1], synth.dyadic((8, 5), 3))


def test_voiced_and_bins():
    states = np.asarray([0, 5, 320, 319, 320], np.int64)
    voiced, bins = api.voiced_and_bins(states, 320)       # dcnet/softmax_viterbi.py:2427-2431
    assert voiced.tolist() == [True, True, False, True, False]
    assert bins.tolist() == [0, 5, 319, 319, 319]


def test_wave_planner():
    """viterbi_spl_b200.waves.plan_waves: consecutive cover, quantum-sized waves, ragged tail only at the end."""
    from viterbi_spl_b200.waves import plan_waves, wave_bytes_per_clip
    per_clip = wave_bytes_per_clip(3000, 361)
    assert per_clip == 3 * 3000 * 361 * 4 + 3000 * 8 + 4
    # config 5 on one 180 GB B200: 65,536 clips, quantum 1036
    waves = plan_waves(65536, per_clip, 150 * 10 ** 9, quantum=1036)
    assert waves[0][0] == 0 and waves[-1][1] == 65536
    assert all(a2 == b1 for (_, b1), (a2, _) in zip(waves, waves[1:]))
    assert all((b - a) % 1036 == 0 for a, b in waves[:-1])
    assert all((b - a) * per_clip <= 150 * 10 ** 9 for a, b in waves)
    assert len(waves) == 6
    # a budget below one quantum still makes progress; empty jobs give no waves; a clip larger than the budget raises
    assert plan_waves(10, 100, 350, quantum=8) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert plan_waves(0, 100, 1000) == []
    assert plan_waves(5, 100, 10 ** 6, quantum=4, max_wave_clips=2) == [(0, 2), (2, 4), (4, 5)]
    with pytest.raises(MemoryError):
        plan_waves(3, 1000, 999)


def test_host_paths_validate_clip_lengths():
    """decode_host / run_host / MelodyPipeline.evaluate still have the lengths on the host: a value outside [0, T] would
    make the kernels read and write past a clip's rows, so it is refused before anything is uploaded."""
    from viterbi_spl_b200.decoder import checked_lengths
    assert checked_lengths([0, 5, 3], 3, 5).dtype == np.int32
    for bad in ([0, 6, 3], [-1, 2, 2], [1, 2], np.zeros((3, 1), np.int32)):
        with pytest.raises(ValueError):
            checked_lengths(bad, 3, 5)
