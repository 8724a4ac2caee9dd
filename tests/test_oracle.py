"""CPU tests of the oracle itself (test infrastructure): the NumPy and C restatements of imm/tf_viterbi.py:75-109 must
reproduce the golden vectors made by executing the reference's own functions, and the live reference when present."""
import hashlib
import os

import numpy as np
import pytest

from oracle import c_oracle, np_oracle, ref_loader
from viterbi_spl_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    return np.load(os.path.join(GOLD, name))


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def exact_case(tag):
    g = load('exact_inputs.npz')
    S, T, coarse = (int(x) for x in g[tag + '_spec'])
    A, pi = synth.dyadic_hmm(S, seed=S + T, coarse=bool(coarse))
    E = synth.tie_stress((T, S), seed=7 * S + T) if coarse else synth.dyadic((T, S), seed=7 * S + T)
    assert sha(A, pi, E) == str(g[tag + '_sha']), 'seeded exact inputs are not reproducible on this machine'
    return A, pi, E, g[tag + '_states']


EXACT_TAGS = ['dyadic361', 'ties361', 'dyadic722', 'ties97', 't1', 't2']


@pytest.mark.parametrize('tag', EXACT_TAGS)
@pytest.mark.parametrize('impl', ['np', 'c'])
def test_exact_inputs_golden(tag, impl, oracle_c):
    A, pi, E, want = exact_case(tag)
    fn = np_oracle.viterbi_log_np if impl == 'np' else c_oracle.viterbi_log_c
    states, score = fn(A, pi, E)
    assert states.dtype == np.int64
    assert np.array_equal(states, want)


@pytest.mark.parametrize('impl', ['np', 'c'])
def test_msnet_logdomain_golden(impl, oracle_c):
    g = load('msnet_logdomain.npz')
    fn = np_oracle.viterbi_log_np if impl == 'np' else c_oracle.viterbi_log_c
    for kind in ('dense', 'sparse'):
        states, _ = fn(g['logA_T'], g['log_pi'], g['E_' + kind])
        assert np.array_equal(states, g['states_' + kind]), kind


def test_family_goldens_logdomain(oracle_c):
    g = load('family_a.npz')
    assert np.array_equal(c_oracle.viterbi_log_c(g['logA_T'], g['log_pi'], g['log_probs_ts'])[0], g['states'])
    g = load('tonet_family_b.npz')
    E = np.require(g['log_probs_st'].T, requirements=['C'])
    assert np.array_equal(np_oracle.viterbi_log_np(g['logA_T'], g['log_pi'], E)[0], g['states'])
    g = load('msnet_softmax_viterbi.npz')
    m = load('msnet_logdomain.npz')
    for scaled in (0, 1):
        st = c_oracle.viterbi_log_c(m['logA_T'], m['log_pi'], g[f'log_prob_ts_{scaled}'])[0]
        assert np.array_equal(st < 320, g[f'voiced_{scaled}'])
        assert np.array_equal(np.minimum(st, 319), g[f'bins_{scaled}'])


def test_family_a_np_matches_golden_when_libm_matches():
    g = load('family_a.npz')
    tiny = np.finfo(np.float32).tiny
    if not np.array_equal(np.log(g['probs_st'].T + tiny), g['log_probs_ts']):
        pytest.skip('this machine\'s float32 log differs from the one that made the golden (expected: not correctly rounded)')
    st = np_oracle.family_a_np(transition_matrix=g['A'], prob_init=g['pi'], probs_st=g['probs_st'])
    assert np.array_equal(st, g['states'])


def test_c_equals_np_with_tables(oracle_c):
    rng = np.random.default_rng(3)
    for S, T in [(1, 1), (2, 7), (33, 50), (130, 40)]:
        A, pi = synth.dyadic_hmm(S, seed=S, coarse=(S % 2 == 0))
        E = synth.tie_stress((T, S), seed=T) if S % 2 else synth.dyadic((T, S), seed=T)
        s1, sc1, T1a, T2a = np_oracle.viterbi_log_np(A, pi, E, return_tables=True)
        s2, sc2, T1b, T2b = c_oracle.viterbi_log_c(A, pi, E, return_tables=True)
        assert np.array_equal(s1, s2) and sc1 == sc2
        assert np.array_equal(T1a, T1b)
        assert np.array_equal(T2a[1:], T2b[1:])


def test_batch_ragged_and_threads(oracle_c):
    S, T, B = 47, 30, 11
    A, pi = synth.dyadic_hmm(S, seed=1)
    E = synth.batch('dyadic', B, T, S, seed0=5)
    L = np.asarray([30, 0, 1, 2, 17, 30, 29, 3, 0, 8, 30], np.int32)
    p1, s1 = np_oracle.decode_batch_np(A, pi, E, L)
    for nt in (1, 3, 0):
        p2, s2 = c_oracle.decode_batch_c(A, pi, E, L, nthreads=nt)
        assert np.array_equal(p1, p2) and np.array_equal(s1, s2)
    assert (p1[1] == -1).all() and s1[1] == -np.inf and (p1[2, 1:] == -1).all()


def test_minus_inf_and_all_equal_rows(oracle_c):
    S, T = 9, 6
    A = np.full((S, S), -np.inf, np.float32)
    A[np.arange(S), np.arange(S)] = 0
    A[0, :] = 0
    pi = np.zeros(S, np.float32)
    E = np.zeros((T, S), np.float32)
    s1, _ = np_oracle.viterbi_log_np(A, pi, E)
    s2, _ = c_oracle.viterbi_log_c(A, pi, E)
    assert np.array_equal(s1, s2) and (s1 == 0).all()       # every cell tied: first maximum wins everywhere


@pytest.mark.skipif(not ref_loader.available(), reason='/root/reference not present (GPU box)')
def test_live_reference_agrees():
    ref = ref_loader.log_domain_decode()
    for seed, (S, T) in enumerate([(61, 90), (321, 40)]):
        A, pi = synth.dyadic_hmm(S, seed=seed, coarse=bool(seed % 2))
        E = synth.dense_softmax(T, S, seed=seed)
        want = ref(log_transition_matrix_T=A, log_prob_init=pi, log_probs_st=np.require(E.T, requirements=['C']))
        assert np.array_equal(np_oracle.viterbi_log_np(A, pi, E)[0], want)
        assert np.array_equal(c_oracle.viterbi_log_c(A, pi, E)[0], want)
        assert np.array_equal(np_oracle.viterbi_log_st_np(log_transition_matrix_T=A, log_prob_init=pi,
                                                          log_probs_st=np.require(E.T, requirements=['C'])), want)


@pytest.mark.skipif(not ref_loader.available(), reason='/root/reference not present (GPU box)')
def test_live_reference_family_a_and_numba():
    g = load('family_a.npz')
    fa = ref_loader.family_a_decode()
    want = fa(transition_matrix=g['A'], prob_init=g['pi'], probs_st=np.asfortranarray(g['probs_st']))
    assert np.array_equal(np_oracle.family_a_np(transition_matrix=g['A'], prob_init=g['pi'], probs_st=g['probs_st']), want)
    nb = ref_loader.numba_core()
    got = nb(np.require(g['A'].T, requirements=['C']).copy(), g['pi'].copy(),
             np.require(g['probs_st'].T, requirements=['C']).copy())
    assert np.array_equal(got, want)


def test_post_decode_restatement_hand_checked():
    """oracle/post_oracle.py (restates TF code that cannot run here): hand-computed frames."""
    from oracle import post_oracle as po
    logits = np.zeros((3, 320), np.float32)             # sigmoid = 0.5 everywhere
    bins = np.asarray([0, 10, 319])
    voiced = np.asarray([True, False, True])
    # frame 0: bins {0, 1} -> (0*0.5 + 0.2*0.5) / 1.0 + 23.6 = 23.7; frame 1: bins 9..11 -> 2.0 + 23.6; frame 2: {318, 319}
    ref = np.asarray([23.7, 0.0, 23.6 + 63.7 + 12.0], np.float32)
    est, c = po.melody_stats_np(ref, logits, bins, voiced)
    assert np.allclose(est, [23.7, -25.6, 23.6 + 63.7], atol=1e-5)
    assert c == dict(gt_voiced=2, gt_unvoiced=1, correct_voiced=2, incorrect_voiced=0, correct_unvoiced=1,
                     correct_pitches_wide=1, correct_pitches_strict=1, correct_chromas_wide=2, correct_chromas_strict=2)


def test_long_sequence_oracle_equals_the_single_threaded_one(oracle_c):
    """vit_oracle_decode_long_f32 (targets of a step split over threads; what checks the 1,000,000-frame case) is
    bit-identical to the plain C restatement, for any thread count."""
    from viterbi_spl_b200 import hmm_params
    A, pi = hmm_params.synthetic_hmm('dcnet')
    logA_T, log_pi = hmm_params.log_params(A, pi)
    E = synth.sparse_peaks(700, 321, seed=9)
    want = c_oracle.viterbi_log_c(logA_T, log_pi, E)
    for n in (1, 3, 8, 0):
        got = c_oracle.viterbi_log_long_c(logA_T, log_pi, E, nthreads=n)
        assert np.array_equal(got[0], want[0]) and got[1] == want[1], n
    A, pi = synth.dyadic_hmm(5, seed=1, coarse=True)
    E = synth.tie_stress((40, 5), 2)
    assert np.array_equal(c_oracle.viterbi_log_long_c(A, pi, E, nthreads=8)[0], c_oracle.viterbi_log_c(A, pi, E)[0])


def test_post_oracle_equals_the_reference_statistics_step():
    """oracle/post_oracle.py against goldens made by EXECUTING the reference's own viterbi_update_states_tf_fn + est_notes_fn
    (dcnet/softmax_viterbi.py:1919-1958, 2923-2979) on a NumPy-backed stand-in for its TensorFlow ops
    (tests/golden/make_golden.py melody_stats), and against the live reference where the checkout is present."""
    from oracle import post_oracle as po
    g = load('melody_stats.npz')
    live = ref_loader.dcnet_melody_stats() if ref_loader.available() else None
    for k in range(int(g['n_cases'])):
        args = (g[f'c{k}_ref'], g[f'c{k}_logits'], g[f'c{k}_bins'], g[f'c{k}_voiced'])
        est, c = po.melody_stats_np(*args)
        assert np.allclose(est, g[f'c{k}_est'], rtol=1e-6, atol=1e-6)            # (np.exp may differ in the last ulp across machines)
        got = np.asarray([c[n] for n in po.COUNTERS], np.int64)
        diff = np.abs(np.abs(est) - args[0])
        edge = int(np.sum(np.abs(diff - 0.5) < 2e-5) + np.sum(np.abs(np.abs(diff - np.round(diff / 12) * 12) - 0.5) < 2e-5))
        assert np.all(np.abs(got - g[f'c{k}_counters']) <= edge) and np.array_equal(got[:5], g[f'c{k}_counters'][:5])
        if live is not None:
            est_l, c_l = live(*args)
            assert np.array_equal(est_l, est) and list(c_l.values()) == [c[n] for n in po.COUNTERS]
