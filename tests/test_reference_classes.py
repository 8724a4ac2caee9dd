"""CPU tests of the class-level drop-ins (viterbi_spl_b200.reference_classes): constructor signatures and parameter
loading, and the host-side emission builders -- bit-exact against the reference's OWN classes executed from
/root/reference (oracle/ref_loader.py; skipped where the checkout is absent) and against the committed goldens those
classes produced (tests/golden/class_calls.npz).  The decode itself needs the GPU: tests/test_gpu_reference_classes.py."""
import inspect
import os

import numpy as np
import pytest

from oracle import ref_loader as rl
import class_cases as cc
from viterbi_spl_b200 import hmm_params, reference_classes as rc

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
needs_reference = pytest.mark.skipif(not rl.available(), reason='reference checkout not present')


def golden_parameters(state_set):
    g = np.load(os.path.join(GOLD, 'class_calls.npz'))
    return g[f'A_{state_set}'], g[f'pi_{state_set}']


@pytest.fixture(scope='module')
def param_dirs(tmp_path_factory):
    dirs = {}
    for state_set in ('msnet_shipped', 'tonet', 'jdc', 'imm_hmm'):
        d = tmp_path_factory.mktemp(state_set)
        A, pi = golden_parameters(state_set)
        cc.write_dat(str(d), A, pi)
        dirs[state_set] = str(d)
    return dirs


def build(case, directory, **kw):
    ns, name, args, _, _ = case
    cls = getattr(getattr(rc, ns), name)
    return cls(*cc.ctor_args(args, cc.Var), directory=directory, **kw)


@pytest.mark.parametrize('case', cc.CASES, ids=cc.case_tag)
def test_emission_tables_match_the_goldens_made_by_the_reference_classes(case, param_dirs):
    g = np.load(os.path.join(GOLD, 'class_calls.npz'))
    tag = cc.case_tag(case)
    obj = build(case, param_dirs[case[3]])
    logits = g[f'{tag}_logits']
    peaks = obj.find_peaks_all_at_once_np_fn(np.require(logits.T, np.float32, ['C']) if case[0] == 'imm' else
                                             (np.pad(logits, [[0, 0], [1, 0]]) if case[:2] == ('dcnet', 'SoftMaxViterbi') else logits))
    want = g[f'{tag}_probs']
    got = obj.observation_probs_fn(logits.copy())
    assert got.dtype == np.float32 and got.shape == want.shape
    assert got.flags['F_CONTIGUOUS' if case[1].startswith('Viterbi') else 'C_CONTIGUOUS']
    # peaks are comparisons only: machine independent.  (shaun tables [S, T]: rows = states; SoftMax tables [T, S] with the
    # unvoiced state rolled to the end)
    if case[1].startswith('Viterbi'):
        assert np.array_equal(got[:-1] > 0, peaks.T)
        assert np.array_equal(want[:-1] > 0, peaks.T)
    else:
        assert np.array_equal(got[:, :-1] > 0, peaks[:, 1:])
        assert np.array_equal(want[:, :-1] > 0, peaks[:, 1:])
    # values: same ufuncs, same dtypes, same order -> same bits on the machine that made the goldens; elsewhere NumPy's
    # exp may differ in the last ulp
    if np.array_equal(np.exp(logits[:4].astype(np.float32)), g[f'{tag}_exp_probe']):
        assert np.array_equal(got, want)
    else:
        assert np.allclose(got, want, rtol=2e-6, atol=0)


@needs_reference
@pytest.mark.parametrize('case', cc.CASES, ids=cc.case_tag)
@pytest.mark.parametrize('seed', [1, 2])
def test_emission_tables_match_the_live_reference_classes(case, seed, param_dirs):
    ns, name, args, state_set, _ = case
    ref = rl.construct_in(param_dirs[state_set], rl.reference_class(ns, name), *cc.ctor_args(args, rl.FakeTF.Variable))
    obj = build(case, param_dirs[state_set])
    logits = cc.logits_for(case, 1000 * seed + len(cc.case_tag(case)), T=400)
    want = ref.observation_probs_fn(logits.copy())
    got = obj.observation_probs_fn(logits.copy())
    assert got.dtype == want.dtype and got.shape == want.shape and got.flags['F_CONTIGUOUS'] == want.flags['F_CONTIGUOUS']
    assert np.array_equal(got, want)
    # attributes the pipelines read
    assert obj.num_freq_bins == ref.num_freq_bins and obj.single_side_peak_width == ref.single_side_peak_width
    for attr in ('transition_matrix', 'ini_probs', 'log_transition_matrix_T', 'log_ini_probs', 'threshold', 'threshold_logit',
                 'scaled'):
        if hasattr(ref, attr):
            assert np.array_equal(np.asarray(getattr(obj, attr)), np.asarray(getattr(ref, attr))), attr


@needs_reference
@pytest.mark.parametrize('case', cc.CASES, ids=cc.case_tag)
def test_constructor_signatures_match_the_reference(case):
    ns, name, _, _, _ = case
    ref_params = list(inspect.signature(rl.reference_class(ns, name).__init__).parameters.values())
    mine = list(inspect.signature(getattr(getattr(rc, ns), name).__init__).parameters.values())
    positional = [p for p in mine if p.kind == p.POSITIONAL_OR_KEYWORD]
    assert [p.name for p in positional] == [p.name for p in ref_params]
    assert all(p.kind == p.KEYWORD_ONLY for p in mine[len(positional):])           # the extras never shift an argument


def test_constructors_read_the_working_directory_like_the_reference(param_dirs, monkeypatch):
    monkeypatch.chdir(param_dirs['tonet'])
    obj = rc.tonet.Viterbi(0.5)                                                     # tonet/softmax_priors.py:288
    A, pi = golden_parameters('tonet')
    want_A, want_pi = hmm_params.log_params(A, pi)
    assert np.array_equal(obj.log_transition_matrix_T, want_A) and np.array_equal(obj.log_ini_probs, want_pi)
    assert not obj.log_transition_matrix_T.flags['WRITEABLE']
    with pytest.raises(AssertionError):
        rc.tonet.Viterbi(1.5)
    with pytest.raises(AssertionError):
        rc.jdc.SoftMaxViterbi(True)                                                 # 361-state files, 722-state class
    monkeypatch.chdir(param_dirs['msnet_shipped'])
    v = rc.dcnet.Viterbi()
    assert v.threshold == np.log(0.31 / (1. - 0.31)) and v.transition_matrix.shape == (321, 321)
    s = rc.dcnet.SoftMaxViterbi(voicing_threspold_prob=cc.Var(0.31), scaled=True)
    assert s.scaled is True and s.log_ini_probs.dtype == np.float32
    # arrays instead of files
    s2 = rc.msnet.SoftMaxViterbi(False, transition_matrix=v.transition_matrix, ini_probs=v.ini_probs)
    assert np.array_equal(s2.log_transition_matrix_T, s.log_transition_matrix_T)


@needs_reference
def test_imm_hf0_processing_matches_the_live_reference():
    ref = rl.imm_viterbi_class()(20, 721)
    obj = rc.imm.HF0Viterbi(20, 721)
    assert np.array_equal(obj.log_transition_matrix_T, ref.log_transition_matrix_T)
    assert np.array_equal(obj.log_prob_init, ref.log_prob_init)
    for seed in (1, 2):
        HF0 = cc.hf0_case(seed)
        want, got = ref.process_HF0_fn(HF0), obj.process_HF0_fn(HF0)
        assert got.dtype == want.dtype == np.float32 and np.array_equal(got, want)
    tiny = np.full((721, 3), 1e-45, np.float32)                                   # below exp(-87): the floor of :80-81
    assert np.array_equal(obj.process_HF0_fn(tiny), ref.process_HF0_fn(tiny))


def test_imm_hf0_golden():
    g = np.load(os.path.join(GOLD, 'class_calls.npz'))
    obj = rc.imm.HF0Viterbi(20, 721)
    got = obj.process_HF0_fn(g['imm_HF0Viterbi_HF0'])
    want = g['imm_HF0Viterbi_log_HF0']
    assert got.shape == (722, want.shape[1]) and np.allclose(got, want, rtol=2e-6, atol=0)
    assert np.all(got[-1] == got.min())                                            # the padded unvoiced row
