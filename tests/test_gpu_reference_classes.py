"""GPU tests of the class-level drop-ins: ``voiced, bins = viterbi_ins(logits)`` -- the call the reference's pipelines
make (dcnet/softmax_viterbi.py:3039) -- for every copy of ``class Viterbi`` / ``class SoftMaxViterbi``, against goldens
produced by constructing and calling the reference's OWN classes (tests/golden/make_golden.py ``class_calls``), incl.
the jdc (721 bins, peak half-width 16) and imm (721 bins, half-width 20; HF0 decoder) shapes."""
import os

import numpy as np
import pytest

import class_cases as cc
from oracle import c_oracle
from viterbi_spl_b200 import reference_classes as rc

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
TINY = np.finfo(np.float32).tiny


@pytest.fixture(scope='module')
def gold():
    return np.load(os.path.join(GOLD, 'class_calls.npz'))


@pytest.fixture(scope='module')
def param_dirs(tmp_path_factory, gold, cuda_lib):
    dirs = {}
    for state_set in ('msnet_shipped', 'tonet', 'jdc', 'imm_hmm'):
        d = tmp_path_factory.mktemp(state_set)
        cc.write_dat(str(d), gold[f'A_{state_set}'], gold[f'pi_{state_set}'])
        dirs[state_set] = str(d)
    return dirs


def build(case, directory, **kw):
    ns, name, args, _, _ = case
    return getattr(getattr(rc, ns), name)(*cc.ctor_args(args, cc.Var), directory=directory, **kw)


@pytest.mark.parametrize('case', cc.CASES, ids=cc.case_tag)
def test_call_reproduces_the_reference_class(case, param_dirs, gold, monkeypatch):
    tag = cc.case_tag(case)
    monkeypatch.chdir(param_dirs[case[3]])                       # the constructors read their .dat files from cwd
    ns, name, args, _, _ = case
    obj = getattr(getattr(rc, ns), name)(*cc.ctor_args(args, cc.Var))
    logits = gold[f'{tag}_logits']
    probs = obj.observation_probs_fn(logits.copy())
    voiced, bins = obj(logits.copy())
    assert voiced.dtype == np.bool_ and bins.dtype == np.int64 and voiced.shape == bins.shape == (logits.shape[1 if ns == 'imm' else 0],)
    same_bits = np.array_equal(probs, gold[f'{tag}_probs']) and np.array_equal(np.log(probs + TINY), gold[f'{tag}_log_probs'])
    if same_bits:
        # same NumPy exp/log bits as the machine that ran the reference: the whole call must be bit-identical
        assert np.array_equal(voiced, gold[f'{tag}_voiced']) and np.array_equal(bins, gold[f'{tag}_bins'])
    else:
        # another libm: the emission table differs in the last ulp; check the decode of THIS table against the oracle
        assert np.allclose(probs, gold[f'{tag}_probs'], rtol=2e-6, atol=0)
        A, pi = gold[f'A_{case[3]}'], gold[f'pi_{case[3]}']
        logA_T = np.require(np.log(A + TINY).T, np.float32, ['C'])
        log_pi = np.log(pi + TINY).astype(np.float32)
        E = np.log(probs + TINY)
        E = np.require(E.T if name.startswith('Viterbi') else E, np.float32, ['C'])
        want, _ = c_oracle.viterbi_log_c(logA_T, log_pi, E)
        n = obj.num_freq_bins
        assert np.array_equal(voiced, want < n) and np.array_equal(bins, np.minimum(want, n - 1))
    # the decode of the GOLDEN log-domain table through the log-domain entry point: equal to the oracle's, and to the
    # golden result when this machine's np.log gives the golden machine's log matrix
    from viterbi_spl_b200 import reference_api
    A, pi = gold[f'A_{case[3]}'], gold[f'pi_{case[3]}']
    logA_T = np.require(np.log(A + TINY).T, np.float32, ['C'])
    log_pi = np.log(pi + TINY).astype(np.float32)
    E = gold[f'{tag}_log_probs']
    E = np.require(E.T if name.startswith('Viterbi') else E, np.float32, ['C'])
    st = reference_api.viterbi_librosa_fn(log_transition_matrix_T=logA_T, log_prob_init=log_pi,
                                          log_probs_st=np.require(E.T, np.float32, ['C']))
    want, _ = c_oracle.viterbi_log_c(logA_T, log_pi, E)
    assert np.array_equal(st, want)
    n = obj.num_freq_bins
    if same_bits:
        assert np.array_equal(st < n, gold[f'{tag}_voiced']) and np.array_equal(np.minimum(st, n - 1), gold[f'{tag}_bins'])


@pytest.mark.parametrize('case', [c for c in cc.CASES if c[0] in ('tonet', 'jdc', 'imm', 'msnet')], ids=cc.case_tag)
def test_device_emissions_option_agrees_on_almost_every_frame(case, param_dirs, gold):
    """device_emissions=True: emission table from vit_emissions_f32 (peaks exact, exp/log within 1e-5 of NumPy's), so a
    frame can differ only where two paths are within rounding of each other."""
    tag = cc.case_tag(case)
    host = build(case, param_dirs[case[3]])
    dev = build(case, param_dirs[case[3]], device_emissions=True)
    logits = gold[f'{tag}_logits']
    v0, b0 = host(logits.copy())
    v1, b1 = dev(logits.copy())
    assert v1.shape == v0.shape and b1.dtype == np.int64
    assert np.mean((v0 == v1) & (b0 == b1)) >= 0.97


def test_imm_hf0_decoder(gold):
    obj = rc.imm.HF0Viterbi(20, 721)                             # imm/tf_imm.py:166
    HF0 = gold['imm_HF0Viterbi_HF0']
    log_HF0 = obj.process_HF0_fn(HF0)
    states = obj(HF0)
    assert states.dtype == np.int64 and states.shape == (HF0.shape[1],)
    want, _ = c_oracle.viterbi_log_c(obj.log_transition_matrix_T, obj.log_prob_init, np.require(log_HF0.T, np.float32, ['C']))
    assert np.array_equal(states, want)
    import hashlib
    same_matrix = hashlib.sha256(np.ascontiguousarray(obj.log_transition_matrix_T).tobytes()).hexdigest() == str(gold['imm_HF0Viterbi_logA_T_sha'])
    if same_matrix and np.array_equal(log_HF0, gold['imm_HF0Viterbi_log_HF0']):
        assert np.array_equal(states, gold['imm_HF0Viterbi_states'])
    if same_matrix:
        assert np.array_equal(obj.viterbi_librosa_fn(gold['imm_HF0Viterbi_log_HF0']), gold['imm_HF0Viterbi_states'])
