"""HMM parameter builders and .dat I/O against the reference scripts and fixtures."""
import os

import numpy as np
import pytest

from oracle import ref_loader as rl
from viterbi_spl_b200 import hmm_params as hp

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
needs_ref = pytest.mark.skipif(not rl.available(), reason='/root/reference not present (GPU box)')


def synth_counts(n, rng, width):
    ti = np.zeros((n + 1, n + 1), np.int64)
    for _ in range(20000):
        i = rng.integers(0, n)
        j = int(np.clip(i + np.round(rng.laplace(0, width)), 0, n - 1))
        ti[i, j] += 1
    return ti


@needs_ref
def test_banded_transition_matches_dcnet_script():
    ti = synth_counts(320, np.random.default_rng(5), 2.5)
    ref = rl.run_ref_script('dcnet/viterbi_transition_matrix.py', {'transition_int': ti})['viterbi_transition_matrix']
    mine = hp.banded_transition_matrix(hp.jump_histogram(ti, 320, 12), 320, 6, hp.SWITCH_DCNET)
    assert mine.dtype == np.float32 and np.array_equal(ref, mine)


@needs_ref
def test_banded_transition_matches_tonet_script():
    ti = synth_counts(360, np.random.default_rng(6), 3.5)
    ref = rl.run_ref_script('tonet/viterbi_transition_post_processing.py', {'transition_int': ti})['viterbi_transition_matrix']
    d_max = hp.single_side_d_max(0.01, 60)
    assert d_max == 14
    mine = hp.banded_transition_matrix(hp.jump_histogram(ti, 360, d_max), 360, 2, hp.SWITCH_TONET)
    assert np.array_equal(ref, mine)


@needs_ref
def test_init_probs_match_scripts():
    ps = hp.synthetic_p_steady(360, seed=3)
    ref = rl.run_ref_script('tonet/p_steady_post_processing.py', {'p_steady': ps})['viterbi_init_probs']
    assert np.array_equal(ref, hp.floored_init_probs(ps))
    ps = hp.synthetic_p_steady(320, seed=4)
    ref = rl.run_ref_script('dcnet/viterbi_init_probs.py', {'p_steady': ps})['viterbi_init_probs']
    assert np.array_equal(ref, hp.floored_init_probs(ps, 3e-4))


@needs_ref
def test_dense_imm_matches_reference():
    g = rl.imm_gen_transition_matrix()
    assert np.array_equal(g(20, 721), hp.dense_imm_transition_matrix(20, 721))
    assert np.array_equal(g(5, 60), hp.dense_imm_transition_matrix(5, 60))


@needs_ref
def test_dat_reader_writer_match_reference(tmp_path):
    name, A = hp.load_dat(os.path.join(rl.REF_ROOT, 'msnet', 'viterbi_transition_matrix.dat'))
    name2, A2 = rl.dat_loader()(os.path.join(rl.REF_ROOT, 'msnet', 'viterbi_transition_matrix.dat'))
    assert name == name2 == 'viterbi_transition_matrix' and np.array_equal(A, A2) and A.shape == (321, 321)
    rng = np.random.default_rng(0)
    for arr in (np.asfortranarray(rng.random((5, 7)).astype(np.float32)), rng.random((4, 3)), rng.random(9).astype(np.float32),
                rng.integers(0, 9, (3, 2, 4))):
        a, b = str(tmp_path / 'a.dat'), str(tmp_path / 'b.dat')
        hp.save_dat(a, arr, 'rec')
        rl.dat_saver()(b, arr, 'rec')
        assert open(a, 'rb').read() == open(b, 'rb').read()
        n1, x1 = hp.load_dat(b)
        n2, x2 = rl.dat_loader()(a)
        assert n1 == n2 == 'rec' and np.array_equal(x1, arr) and np.array_equal(x2, arr)
        assert x1.flags['F_CONTIGUOUS'] == x2.flags['F_CONTIGUOUS']


def test_dat_roundtrip_without_reference(tmp_path):
    rng = np.random.default_rng(1)
    arr = rng.random((6, 4)).astype(np.float32)
    f = str(tmp_path / 'x.dat')
    hp.save_dat(f, arr, 'x')
    with open(f, 'rb') as fh:
        assert fh.readline() == b'x C float32 6 4\n'
    name, back = hp.load_dat(f)
    assert name == 'x' and np.array_equal(back, arr)
    # header variant without the C/F flag (the shipped msnet files use it)
    with open(f, 'wb') as fh:
        fh.write(b'viterbi_init_probs float32 24\n' + arr.tobytes())
    name, back = hp.load_dat(f)
    assert name == 'viterbi_init_probs' and np.array_equal(back, arr.reshape(24))


def test_synthetic_tonet_parameters_are_the_golden_ones():
    g = np.load(os.path.join(GOLD, 'tonet_family_b.npz'))
    A, pi = hp.synthetic_hmm('tonet', seed=0)
    assert np.array_equal(A, g['A']) and np.array_equal(pi, g['pi'])
    assert A.shape == (361, 361) and np.all(np.isclose(A.sum(1), 1)) and np.isclose(pi.sum(), 1)
    assert np.argmax(pi) == 360                                            # unvoiced is the last state
    band = np.abs(np.subtract.outer(np.arange(360), np.arange(360))) <= 14
    assert np.all(A[:360, :360][~band] == 0) and np.all(A[:360, :360][band] > 0)
    logA_T, log_pi = hp.log_params(A, pi)
    assert logA_T.flags['C_CONTIGUOUS'] and logA_T.dtype == np.float32
    assert np.isclose(logA_T.min(), -87.33655, atol=1e-4)                   # log(0 + tiny)


def test_state_sets():
    for name, S in [('dcnet', 321), ('tonet', 361), ('jdc', 722), ('imm', 722)]:
        A, pi = hp.synthetic_hmm(name)
        assert A.shape == (S, S) and pi.shape == (S,)
        assert np.allclose(A.sum(1), 1) and np.isclose(pi.sum(), 1)
    assert np.all(hp.synthetic_hmm('imm')[0] > 0)                           # imm/tf_imm.py:55
