"""Parity tests proper (run on the B200 with `-m gpu`): the CUDA decoders, called through the C ABI, against

  * the golden vectors produced by executing the reference's own functions (tests/golden/make_golden.py),
  * the CPU oracle on the same seeded inputs (sizes the oracle finishes in seconds),
  * size-independent properties at BASELINE.json's full size (1024 clips x 3000 frames x 361 states).

Bar: bit-exact state paths AND bit-exact fp32 path scores (the north star asks 1e-5 relative on scores; the
arithmetic is the same sequence of binary32 adds, so equality is what is asserted).
"""
import ctypes
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, np_oracle
from viterbi_spl_b200 import hmm_params, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), 'golden')
ALGOS = ['tmem', 'cluster', 'backpointer', 'stream']


def load(name):
    return np.load(os.path.join(GOLD, name))


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.fixture(scope='module')
def Decoder(cuda_lib):
    assert torch.cuda.is_available(), 'gpu-marked tests need a CUDA device'
    from viterbi_spl_b200 import ViterbiDecoder
    return ViterbiDecoder


def algos_for(S):
    from viterbi_spl_b200 import _lib
    if S <= 364:
        return ALGOS
    tmem_ok = _lib.load().vit_select_algo(1, 1, S) == _lib.ALGO_TMEM
    return ['tmem', 'backpointer', 'auto', 'stream'] if tmem_ok else ['backpointer', 'auto', 'stream']


# ---- golden vectors (made by the reference's own code) ------------------------------------------------------------

@pytest.mark.parametrize('tag', ['dyadic361', 'ties361', 'dyadic722', 'ties97', 't1', 't2'])
def test_golden_exact_inputs(Decoder, tag):
    g = load('exact_inputs.npz')
    S, T, coarse = (int(x) for x in g[tag + '_spec'])
    A, pi = synth.dyadic_hmm(S, seed=S + T, coarse=bool(coarse))
    E = synth.tie_stress((T, S), seed=7 * S + T) if coarse else synth.dyadic((T, S), seed=7 * S + T)
    assert sha(A, pi, E) == str(g[tag + '_sha'])
    for algo in algos_for(S):
        paths, scores = Decoder(A, pi, algo=algo).decode_host(E[None])
        assert paths.dtype == np.int64 and np.array_equal(paths[0], g[tag + '_states']), (tag, algo)


@pytest.mark.parametrize('algo', ALGOS)
def test_golden_msnet_shipped_parameters(Decoder, algo):
    g = load('msnet_logdomain.npz')
    dec = Decoder(g['logA_T'], g['log_pi'], algo=algo)
    E = np.stack([g['E_dense'], g['E_sparse']])
    paths, scores = dec.decode_host(E)
    assert np.array_equal(paths[0], g['states_dense'])
    assert np.array_equal(paths[1], g['states_sparse'])
    m = load('msnet_softmax_viterbi.npz')
    for scaled in (0, 1):
        st = dec.decode_host(m[f'log_prob_ts_{scaled}'][None])[0][0]
        assert np.array_equal(st < 320, m[f'voiced_{scaled}'])
        assert np.array_equal(np.minimum(st, 319), m[f'bins_{scaled}'])


def test_golden_imm_dense_matrix(Decoder):
    g = load('imm_dense.npz')
    A = hmm_params.dense_imm_transition_matrix(20, 721)
    logA_T, log_pi = hmm_params.log_params(A, np.full([722], 1. / 722), add_tiny=False)
    if sha(logA_T) != str(g['logA_T_sha']):
        pytest.skip('float64 log of this machine differs from the golden generator')
    E = np.require(g['log_HF0'].T, requirements=['C'])
    paths, _ = Decoder(logA_T, log_pi).decode_host(E[None])
    assert np.array_equal(paths[0], g['states'])


# ---- the reference's entry points (families A-D) --------------------------------------------------------------------

def test_reference_entry_points(Decoder):
    from viterbi_spl_b200 import reference_api as api
    tiny = np.finfo(np.float32).tiny
    g = load('family_a.npz')
    same_libm = np.array_equal(np.log(g['probs_st'].T + tiny), g['log_probs_ts'])
    want = g['states'] if same_libm else np_oracle.family_a_np(transition_matrix=g['A'], prob_init=g['pi'], probs_st=g['probs_st'])
    probs_f = np.asfortranarray(g['probs_st'])
    keep = probs_f.copy(order='F')
    for fn in (api.Viterbi.viterbi_librosa_fn, api.viterbi_librosa_c_fn):
        got = fn(transition_matrix=g['A'], prob_init=g['pi'], probs_st=probs_f)
        assert got.dtype == np.int64 and np.array_equal(got, want)
        assert np.array_equal(probs_f, keep)                               # Family A does not mutate its inputs
    assert np.array_equal(api.viterbi_tf_fn(g['A'], g['pi'], g['probs_st']), want.astype(np.int32))
    # numba path: the core logs its inputs in place; an F-ordered probs_st is reached through the view
    p2 = probs_f.copy(order='F')
    got = api.viterbi_numba_fn(transition_matrix=g['A'], prob_init=g['pi'], probs_st=p2)
    assert np.array_equal(got, want)
    # (the AOT core's `tinyp` is a float64 literal: numba logs in float64 and rounds once, dcnet/aot_viterbi_core.py:18-25)
    assert np.array_equal(p2, np.log(keep.astype(np.float64) + 1.1754944e-38).astype(np.float32))
    # log-domain function
    got = api.viterbi_librosa_fn(log_transition_matrix_T=g['logA_T'], log_prob_init=g['log_pi'],
                                 log_probs_st=np.require(g['log_probs_ts'].T, requirements=['C']))
    assert np.array_equal(got, g['states'])
    # float64 initial log-probabilities: the reference adds T1[0] in float64 and rounds once (imm/tf_viterbi.py:94)
    pi64 = g['log_pi'].astype(np.float64) + 1e-9
    E_st = np.require(g['log_probs_ts'].T, requirements=['C'])
    got = api.viterbi_librosa_fn(log_transition_matrix_T=g['logA_T'], log_prob_init=pi64, log_probs_st=E_st)
    assert np.array_equal(got, np_oracle.viterbi_log_np(g['logA_T'], np.zeros_like(g['log_pi']),
                                                        np.concatenate([(pi64 + g['log_probs_ts'][0]).astype(np.float32)[None],
                                                                        g['log_probs_ts'][1:]]))[0])
    with pytest.raises(AssertionError):
        api.viterbi_librosa_c_fn(transition_matrix=g['A'].astype(np.float64), prob_init=g['pi'], probs_st=probs_f)
    assert np.array_equal(api.viterbi_librosa_f64_fn(transition_matrix=g['A'], prob_init=g['pi'], probs_st=probs_f),
                          g['states_f64'] if same_libm else want)         # a7: float64-table variant, same paths here
    # eager-TF twin of the log-domain function (imm/tf_viterbi.py:8): same math, int32 result, any emission layout
    got = api.tf_viterbi_librosa_fn(tf_log_transition_matrix_T=g['logA_T'], tf_log_prob_init=g['log_pi'],
                                    tf_or_np_log_probs_st=g['log_probs_ts'].T)
    assert got.dtype == np.int32 and np.array_equal(got, g['states'])

    # Family B (tonet): F-ordered prob-domain [S, T], logged in place
    b = load('tonet_family_b.npz')
    vb = api.ViterbiB(0.5, transition_matrix=b['A'], ini_probs=b['pi'])     # tonet/softmax_priors.py:1693 Viterbi(voicing_threshold)
    assert np.array_equal(vb.log_transition_matrix_T, b['logA_T']) or not same_libm
    probs = np.asfortranarray(b['probs_st'])
    want_b = b['states'] if np.array_equal(np.log(b['probs_st'] + tiny), b['log_probs_st']) else \
        np_oracle.viterbi_log_np(vb.log_transition_matrix_T, vb.log_ini_probs, np.log(b['probs_st'].T + tiny))[0]
    got = vb.viterbi_librosa_fn(probs)
    assert np.array_equal(got, want_b)
    assert np.array_equal(probs, np.log(b['probs_st'] + tiny))             # mutated in place like the reference
    with pytest.raises(AssertionError):
        vb.viterbi_librosa_fn(np.ascontiguousarray(b['probs_st']))         # must be F-contiguous
    voiced, bins = api.voiced_and_bins(got, 360)
    if want_b is b['states']:
        assert np.array_equal(voiced, b['voiced']) and np.array_equal(bins, b['bins'])
        v2, b2 = vb(b['logits'].copy())                                    # the class-level call: logits -> (voiced, bins)
        assert np.array_equal(v2, b['voiced']) and np.array_equal(b2, b['bins'])

    # Family C (msnet shipped parameters): C-ordered prob-domain [T, S]
    m = load('msnet_softmax_viterbi.npz')
    ml = load('msnet_logdomain.npz')
    A_lin = np.load(os.path.join(GOLD, 'class_calls.npz'))['A_msnet_shipped']   # the shipped msnet/*.dat matrix
    sv = api.msnet.SoftMaxViterbi(True, transition_matrix=A_lin, ini_probs=ml['ini_probs'])
    assert np.array_equal(sv.log_transition_matrix_T, ml['logA_T']) or not same_libm
    sv.__dict__['_decoder_obj'] = Decoder(ml['logA_T'], ml['log_pi'])       # decode with the exact shipped log matrix
    for scaled in (0, 1):
        probs = m[f'prob_ts_{scaled}'].copy()
        if not np.array_equal(np.log(probs + tiny), m[f'log_prob_ts_{scaled}']):
            continue
        st = sv.viterbi_librosa_fn(probs)
        assert np.array_equal(st < 320, m[f'voiced_{scaled}']) and np.array_equal(np.minimum(st, 319), m[f'bins_{scaled}'])
        assert np.array_equal(probs, m[f'log_prob_ts_{scaled}'])

    # Family D (imm): dense matrix, log-domain [S, T]
    d = load('imm_dense.npz')
    imm = api.ImmViterbi(20, 721)
    if sha(imm.log_transition_matrix_T) == str(d['logA_T_sha']):
        assert np.array_equal(imm.viterbi_librosa_fn(d['log_HF0']), d['states'])


# ---- oracle on seeded inputs: shapes, ragged lengths, ties -----------------------------------------------------------

SHAPES = [(1, 1, 1), (2, 3, 1), (7, 5, 3), (31, 40, 2), (32, 9, 33), (33, 17, 5), (64, 21, 15), (96, 30, 4), (97, 50, 5),
          (128, 12, 3), (160, 14, 8), (192, 20, 6), (193, 25, 6), (200, 64, 9), (321, 100, 33), (361, 120, 70), (364, 30, 5),
          (365, 20, 3), (384, 18, 16), (385, 16, 9), (500, 15, 4), (722, 24, 17), (769, 9, 15),
          (97, 1000, 6), (361, 700, 5), (40, 2049, 3), (722, 600, 20)]     # long clips: many backtrace segments (speculate + fix up)


@pytest.mark.parametrize('S,T,B', SHAPES)
@pytest.mark.parametrize('kind', ['dyadic', 'tie_stress'])
def test_against_oracle_shapes(Decoder, S, T, B, kind):
    A, pi = synth.dyadic_hmm(S, seed=S, coarse=(kind == 'tie_stress'))
    E = synth.batch(kind, B, T, S, seed0=10 + S)
    want_p, want_s = c_oracle.decode_batch_c(A, pi, E)
    for algo in algos_for(S):
        p, s = Decoder(A, pi, algo=algo).decode_host(E)
        assert np.array_equal(p, want_p), (algo, 'paths')
        assert np.array_equal(s, want_s), (algo, 'scores')


@pytest.mark.parametrize('state_set', ['dcnet', 'tonet'])
@pytest.mark.parametrize('kind', ['dense_softmax', 'sparse_peaks'])
def test_against_oracle_ragged_real_state_sets(Decoder, state_set, kind):
    A, pi = hmm_params.synthetic_hmm(state_set)
    logA_T, log_pi = hmm_params.log_params(A, pi)
    S = len(pi)
    B, T = 45, 150
    E = synth.batch(kind, B, T, S, seed0=3)
    rng = np.random.default_rng(S)
    L = rng.integers(0, T + 1, size=B).astype(np.int32)
    L[:6] = [T, 1, 0, 2, T, 3]
    want_p, want_s = c_oracle.decode_batch_c(logA_T, log_pi, E, L)
    for algo in ALGOS:
        p, s = Decoder(logA_T, log_pi, algo=algo).decode_host(E, L)
        assert np.array_equal(p, want_p) and np.array_equal(s, want_s), algo
        assert (p[2] == -1).all() and s[2] == -np.inf and (p[1, 1:] == -1).all()


def test_more_clips_than_one_wave(Decoder):
    """2100 ragged clips: more than the 74 x 14 (tmem), 33 x 32 (cluster) or 148 x 8 (banded) clips that are co-resident, so
    the persistent CTAs loop over several sub-batches (re-armed delta buffers, barrier phases carried across)."""
    A, pi = hmm_params.synthetic_hmm('tonet')
    logA_T, log_pi = hmm_params.log_params(A, pi)
    B, T, S = 2100, 20, 361
    E = synth.batch('dyadic', B, T, S, seed0=4000)
    L = np.random.default_rng(5).integers(0, T + 1, size=B).astype(np.int32)
    L[::3] = T
    want_p, want_s = c_oracle.decode_batch_c(logA_T, log_pi, E, L)
    for algo in ('tmem', 'banded', 'cluster'):
        p, s = Decoder(logA_T, log_pi, algo=algo).decode_host(E, L)
        assert np.array_equal(p, want_p) and np.array_equal(s, want_s), algo


def test_jdc_and_imm_state_sets_722(Decoder):
    for name, add_tiny in (('jdc', True), ('imm', False), ('imm_hmm', True)):
        A, pi = hmm_params.synthetic_hmm(name)
        logA_T, log_pi = hmm_params.log_params(A, pi, add_tiny=add_tiny)
        E = synth.batch('dense_softmax', 5, 50, 722, seed0=5)
        want_p, want_s = c_oracle.decode_batch_c(logA_T, log_pi, E)
        # auto: the jdc (band +-40) and imm-HMM (+-56: 8 of its 29 band chunks come from shared memory) matrices take the
        # wide banded kernel, the dense imm matrix the tmem kernel
        for algo in ('auto', 'tmem', 'stream', 'backpointer') + (('banded',) if name != 'imm' else ()):
            p, s = Decoder(logA_T, log_pi, algo=algo).decode_host(E)
            assert np.array_equal(p, want_p) and np.array_equal(s, want_s), (name, algo)


def test_T1_T2_tables_match_the_reference_tables(Decoder):
    S, T, B = 61, 40, 3
    A, pi = synth.dyadic_hmm(S, seed=2, coarse=True)
    E = synth.batch('tie_stress', B, T, S, seed0=1)
    dec = Decoder(A, pi)
    dE = torch.as_tensor(E).cuda()
    paths, scores, T1, T2 = dec.decode_device(dE, want_tables=True)
    for b in range(B):
        st, sc, rT1, rT2 = np_oracle.viterbi_log_np(A, pi, E[b], return_tables=True)
        assert np.array_equal(T1[b].cpu().numpy(), rT1)
        assert np.array_equal(T2[b].cpu().numpy().astype(np.int64)[1:], rT2[1:])
        assert np.array_equal(paths[b].cpu().numpy(), st) and scores[b].item() == sc


def test_minus_inf_entries_and_all_ties(Decoder):
    S, T = 40, 25
    A = np.full((S, S), -np.inf, np.float32)
    A[np.arange(S), np.arange(S)] = 0
    A[0, :] = 0
    A[:, 3] = -1.5
    pi = np.zeros(S, np.float32)
    E = np.zeros((2, T, S), np.float32)
    E[1] = synth.tie_stress((T, S), 4)
    want_p, want_s = np_oracle.decode_batch_np(A, pi, E)
    for algo in ALGOS:
        p, s = Decoder(A, pi, algo=algo).decode_host(E)
        assert np.array_equal(p, want_p) and np.array_equal(s, want_s), algo


def banded_model(S, d, dense, seed, coarse, background=-87.33655):
    """logA^T with the structure of the reference's matrices: constant background, band |i-j| <= d of dyadic values,
    optionally one state that is a dense row and a dense column."""
    rng = np.random.default_rng(seed)
    A = np.full((S, S), np.float32(background), np.float32)
    q = 4.0 if coarse else 1024.0
    hi = 8 if coarse else 1 << 14
    for k in range(-d, d + 1):
        idx = np.arange(max(0, -k), min(S, S - k))
        A[idx, idx + k] = -(rng.integers(0, hi, size=len(idx)) / q).astype(np.float32)
    if dense is not None:
        A[dense, :] = -(rng.integers(0, hi, size=S) / q).astype(np.float32) - 3
        A[:, dense] = -(rng.integers(0, hi, size=S) / q).astype(np.float32) - 4
    pi = -(rng.integers(0, hi, size=S) / q).astype(np.float32)
    return A, pi


@pytest.mark.parametrize('S,d,dense,T,B', [(361, 14, 360, 90, 19), (321, 12, 320, 70, 9), (361, 14, 0, 40, 8), (97, 3, 50, 60, 5),
                                           (384, 14, 383, 25, 17), (5, 1, None, 30, 3), (64, 0, None, 20, 9), (200, 8, 17, 33, 11),
                                           (130, 13, None, 260, 4), (33, 4, 32, 1, 2),
                                           # wide bands / 722-state sets: the tensor-memory variant
                                           (722, 40, 721, 40, 9), (722, 33, 0, 25, 5), (500, 20, 250, 30, 12),
                                           (96, 28, None, 50, 4), (768, 40, 767, 12, 3), (384, 15, 383, 20, 10)])
@pytest.mark.parametrize('coarse', [False, True])
def test_banded_fast_path_is_bit_identical(Decoder, S, d, dense, T, B, coarse):
    """VIT_ALGO_BANDED (S (2d+2) cells per frame) against the oracle and the dense kernels, including tie-stress
    values, a -inf background, ragged lengths and dense states in arbitrary positions."""
    from viterbi_spl_b200 import _lib
    # (background -6 on the coarse models: close enough to the band values that background sources win or tie, which
    # sends the structured backtrace through its full-row fallback)
    for bg in (-87.33655, -np.inf) + ((-6.0,) if coarse else ()):
        A, pi = banded_model(S, d, dense, seed=S + d + T, coarse=coarse, background=bg)
        st = _lib.analyze_structure(A)
        assert st.kind == 1 and st.halfwidth <= d
        E = synth.batch('tie_stress' if coarse else 'dyadic', B, T, S, seed0=5 + S)
        L = np.random.default_rng(T).integers(0, T + 1, size=B).astype(np.int32)
        L[0] = T
        want_p, want_s = c_oracle.decode_batch_c(A, pi, E, L)
        dec = Decoder(A, pi, algo='banded')
        p, s = dec.decode_host(E, L)
        assert np.array_equal(p, want_p) and np.array_equal(s, want_s), bg
        pa, sa = Decoder(A, pi, algo='auto').decode_host(E, L)          # auto resolves to the banded kernel here
        assert np.array_equal(pa, want_p) and np.array_equal(sa, want_s)


@pytest.mark.parametrize('state_set', ['dcnet', 'tonet'])
def test_banded_fast_path_on_the_reference_state_sets(Decoder, state_set):
    A, pi = hmm_params.synthetic_hmm(state_set)
    logA_T, log_pi = hmm_params.log_params(A, pi)
    S = len(pi)
    for kind in ('dense_softmax', 'sparse_peaks'):
        E = synth.batch(kind, 21, 300, S, seed0=8)
        want_p, want_s = c_oracle.decode_batch_c(logA_T, log_pi, E)
        dE = torch.as_tensor(E).cuda()
        outs = {}
        for algo in ('banded', 'tmem'):
            p, s, = Decoder(logA_T, log_pi, algo=algo).decode_device(dE)
            outs[algo] = (p.cpu().numpy(), s.cpu().numpy())
            assert np.array_equal(outs[algo][0], want_p) and np.array_equal(outs[algo][1], want_s), (algo, kind)
    # a matrix without the structure is refused by the banded kernel and silently takes the dense one under auto
    Ad, pid = synth.dyadic_hmm(61, seed=3)
    with pytest.raises(Exception):
        Decoder(Ad, pid, algo='banded').decode_host(synth.batch('dyadic', 2, 5, 61, seed0=1))


@pytest.mark.parametrize('S,T,B,slab', [(361, 64, 40, 16), (361, 50, 9, 7), (97, 33, 6, 1), (321, 40, 15, 39), (722, 30, 16, 11)])
def test_frame_slabs_resume_from_the_delta_history(Decoder, S, T, B, slab):
    """The recursion run as consecutive frame ranges [t0, t1) (what decode_host does to overlap the upload of the next
    time slab with the decode of the current one) is bit-identical to one pass -- ragged lengths included."""
    A, pi = synth.dyadic_hmm(S, seed=S + 1, coarse=True)
    E = synth.batch('tie_stress', B, T, S, seed0=77)
    L = np.random.default_rng(S + T).integers(0, T + 1, size=B).astype(np.int32)
    L[0] = T
    want_p, want_s = c_oracle.decode_batch_c(A, pi, E, L)
    dec = Decoder(A, pi, algo='tmem')
    p, s = dec.decode_host(E, L, slab_frames=slab)
    assert np.array_equal(p, want_p) and np.array_equal(s, want_s)
    ps, ss = Decoder(A, pi, algo='stream').decode_host(E, L, slab_frames=slab)      # the streaming kernel resumes too
    assert np.array_equal(ps, want_p) and np.array_equal(ss, want_s)
    if S <= 384 or S % 2 == 0:
        # the banded kernels resume a frame range too (here on a banded matrix with tie-stress values; S = 722 takes the
        # tensor-memory variant)
        Ab, pib = banded_model(S, 6 if S <= 384 else 30, S - 1, seed=S, coarse=True)
        want_pb, want_sb = c_oracle.decode_batch_c(Ab, pib, E, L)
        for algo in ('banded', 'auto'):
            pb, sb = Decoder(Ab, pib, algo=algo).decode_host(E, L, slab_frames=slab)
            assert np.array_equal(pb, want_pb) and np.array_equal(sb, want_sb), algo
    dE, dL = torch.as_tensor(E).cuda(), torch.as_tensor(L).cuda()
    paths = scores = None
    for t0 in range(0, T, slab):
        last = t0 + slab >= T
        out = dec.decode_device(dE, dL, paths, scores, frame_range=(t0, min(T, t0 + slab)), backtrace=last)
        paths, scores = out
    assert np.array_equal(paths.cpu().numpy(), want_p) and np.array_equal(scores.cpu().numpy(), want_s)


@pytest.mark.parametrize('algo', ['tmem', 'banded', 'cluster'])
def test_pipelined_decoder_overlaps_backtrace_without_changing_results(Decoder, algo):
    from viterbi_spl_b200 import PipelinedDecoder
    A, pi = hmm_params.synthetic_hmm('dcnet')
    logA_T, log_pi = hmm_params.log_params(A, pi)
    S, T, B = 321, 400, 30
    batches = [synth.batch('dense_softmax', B, T, S, seed0=100 * k) for k in range(5)]
    dec = Decoder(logA_T, log_pi, algo=algo)
    pd = PipelinedDecoder(dec)
    outs = []
    for E in batches:
        outs.append(pd.submit(torch.as_tensor(E).cuda()))
    pd.finish()
    torch.cuda.synchronize()
    for E, (p, s) in zip(batches, outs):
        want_p, want_s = c_oracle.decode_batch_c(logA_T, log_pi, E)
        assert np.array_equal(p.cpu().numpy(), want_p) and np.array_equal(s.cpu().numpy(), want_s)


def test_device_api_lengths_and_untouched_inputs(Decoder):
    S, T, B = 97, 33, 6
    A, pi = synth.dyadic_hmm(S, seed=9)
    E = synth.batch('dyadic', B, T, S, seed0=2)
    L = np.asarray([33, 5, 0, 1, 20, 33], np.int32)
    dE, dL = torch.as_tensor(E).cuda(), torch.as_tensor(L).cuda()
    keep = dE.clone()
    from viterbi_spl_b200 import decode_batch, ViterbiDecoder
    out_p, out_s = torch.empty((B, T), dtype=torch.int64).pin_memory(), torch.empty((B,), dtype=torch.float32).pin_memory()
    hp, hs = ViterbiDecoder(A, pi).decode_host(E, L, out=(out_p, out_s))       # results land in the caller's pinned buffers
    assert np.shares_memory(hp, out_p.numpy())
    p, s = decode_batch(dE, A, pi, dL)
    assert np.array_equal(hp, p.cpu().numpy()) and np.array_equal(hs, s.cpu().numpy())
    assert p.is_cuda and p.dtype == torch.int64 and s.dtype == torch.float32
    want_p, want_s = np_oracle.decode_batch_np(A, pi, E, L)
    assert np.array_equal(p.cpu().numpy(), want_p) and np.array_equal(s.cpu().numpy(), want_s)
    assert torch.equal(dE, keep)


def test_c_abi_error_codes_on_device(cuda_lib):
    S, T, B = 16, 4, 2
    A, pi = synth.dyadic_hmm(S, seed=1)
    dA, dpi = torch.as_tensor(A).cuda(), torch.as_tensor(pi).cuda()
    dE = torch.zeros((B, T, S), device='cuda')
    paths = torch.empty((B, T), dtype=torch.int64, device='cuda')
    ws = torch.empty(1 << 20, dtype=torch.uint8, device='cuda')
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    assert cuda_lib.vit_decode_f32(P(dA), P(dpi), P(dE), None, B, T, S, P(ws), 64, P(paths), None, None) == -3
    assert cuda_lib.vit_decode_f32(P(dA), P(dpi), P(dE), None, B, T, S, ctypes.c_void_p(ws.data_ptr() + 8), 1 << 19,
                                   P(paths), None, None) == -6
    assert cuda_lib.vit_decode_f32(P(dA), P(dpi), P(dE), None, B, T, S, P(ws), ws.numel(), P(paths), None, None) == 0
    torch.cuda.synchronize()
    n0 = cuda_lib.vit_launch_count()
    assert cuda_lib.vit_decode_f32(P(dA), P(dpi), P(dE), None, 0, T, S, P(ws), ws.numel(), P(paths), None, None) == 0
    assert cuda_lib.vit_launch_count() == n0          # an empty batch launches nothing


# ---- full size: BASELINE.json configuration, size-independent properties ------------------------------------------

def test_full_size_properties(Decoder):
    """1024 clips x 3000 frames x 361 states: (1) the four independent CUDA implementations agree bit-exactly;
    (2) the returned score equals the fp32 score re-accumulated along the returned path with the reference's operation
    order, T1[t][s_t] = fl(fl(T1[t-1][s_{t-1}] + B[s_t, s_{t-1}]) + E[t][s_t]); (3) a subset equals the CPU oracle;
    (4) decoding is deterministic and invariant to the order of the clips in the batch."""
    B, T, S = 1024, 3000, 361
    A, pi = hmm_params.synthetic_hmm('tonet')
    logA_T, log_pi = hmm_params.log_params(A, pi)
    dev = torch.device('cuda')
    E = synth.device_dense_softmax(B, T, S, seed=99, device=dev)
    dec_c, dec_b = Decoder(logA_T, log_pi, algo='tmem'), Decoder(logA_T, log_pi, algo='backpointer')
    p1, s1 = dec_c.decode_device(E)
    p2, s2 = dec_b.decode_device(E)
    assert torch.equal(p1, p2) and torch.equal(s1, s2)
    del dec_b, p2, s2
    for other in ('cluster', 'banded'):
        dec_k = Decoder(logA_T, log_pi, algo=other)
        p2, s2 = dec_k.decode_device(E)
        assert torch.equal(p1, p2) and torch.equal(s1, s2), other
        del dec_k, p2, s2
    torch.cuda.empty_cache()
    # (2) score along the path, all clips at once, one fused pass per frame
    dA, dpi = torch.as_tensor(logA_T, device=dev), torch.as_tensor(log_pi, device=dev)
    ar = torch.arange(B, device=dev)
    acc = dpi[p1[:, 0]] + E[ar, 0, p1[:, 0]]
    for t in range(1, T):
        acc = (acc + dA[p1[:, t], p1[:, t - 1]]) + E[ar, t, p1[:, t]]
    assert torch.equal(acc, s1)
    assert int(p1.min()) >= 0 and int(p1.max()) < S
    # (3) oracle on a subset
    sub = [0, 511, 1023]
    want_p, want_s = c_oracle.decode_batch_c(logA_T, log_pi, E[sub].cpu().numpy())
    assert np.array_equal(p1[sub].cpu().numpy(), want_p) and np.array_equal(s1[sub].cpu().numpy(), want_s)
    # (4) determinism + permutation invariance on a slice (keeps memory bounded)
    perm = torch.randperm(256, device=dev)
    p3, s3 = dec_c.decode_device(E[:256][perm].contiguous())
    assert torch.equal(p3, p1[:256][perm]) and torch.equal(s3, s1[:256][perm])


# ---- jobs larger than HBM: waves (viterbi_spl_b200.waves) ----------------------------------------------------------

@pytest.mark.parametrize('algo', ['auto', 'tmem'])
def test_wave_decoder_matches_single_batch_and_oracle(Decoder, algo):
    from viterbi_spl_b200 import _lib
    from viterbi_spl_b200.waves import WaveDecoder, wave_bytes_per_clip
    A, pi = hmm_params.synthetic_hmm('tonet')
    logA_T, log_pi = hmm_params.log_params(A, pi)
    B, T, S = 23, 50, 361
    E = synth.batch('dense_softmax', B, T, S, seed0=300)
    lengths = np.asarray([T - (b % 5) * 7 for b in range(B)], np.int32)
    ref_paths, ref_scores = np_oracle.decode_batch_np(logA_T, log_pi, E, lengths)
    dec = Decoder(logA_T, log_pi, algo=algo)
    assert _lib.clips_in_flight(S, dec.algo, dec.structure) in (1184, 1036)      # 148 SMs x 8, 74 clusters x 14
    dE = torch.from_numpy(E).cuda()
    dL = torch.from_numpy(lengths).cuda()
    wd = WaveDecoder(dec, T, budget_bytes=5 * wave_bytes_per_clip(T, S))          # 5 clips per wave -> 5 waves
    got_p = np.full((B, T), -7, np.int64)
    got_s = np.zeros(B, np.float32)

    def fill(a, b, out):
        out.copy_(dE[a:b])
        return dL[a:b].contiguous()

    def sink(a, b, paths, scores):
        got_p[a:b] = paths.cpu().numpy()
        got_s[a:b] = scores.cpu().numpy()

    waves = wd.run(B, fill, sink)
    assert [b - a for a, b in waves] == [5, 5, 5, 5, 3]
    assert np.array_equal(got_p, ref_paths) and np.array_equal(got_s, ref_scores)


def test_stream_kernel_many_sub_batches_and_big_state_sets(Decoder):
    """VIT_ALGO_STREAM: more sub-batches than SMs (the producer runs ahead across sub-batch boundaries), ragged lengths,
    and state sets past what the tensor-memory kernel's clusters cover."""
    from viterbi_spl_b200 import _lib
    for S, B, T in ((33, 14 * 148 + 29, 12), (1100, 5, 9), (722, 40, 33)):
        A, pi = synth.dyadic_hmm(S, seed=S, coarse=True)
        E = synth.batch('tie_stress', min(B, 64), T, S, seed0=3 * S)
        E = np.ascontiguousarray(np.tile(E, ((B + len(E) - 1) // len(E), 1, 1))[:B])
        L = np.random.default_rng(S).integers(0, T + 1, size=B).astype(np.int32)
        L[0] = T
        want_p, want_s = c_oracle.decode_batch_c(A, pi, E, L)
        p, s = Decoder(A, pi, algo='stream').decode_host(E, L)
        assert np.array_equal(p, want_p) and np.array_equal(s, want_s), (S, B, T)
    # auto: a dense 722-state matrix takes the streaming kernel once the batch fills its pass, the tensor-memory kernel
    # for small batches
    L_ = _lib.load()
    assert L_.vit_select_algo(4096, 100, 722) == _lib.ALGO_STREAM
    assert L_.vit_select_algo(64, 100, 722) == _lib.ALGO_TMEM
    assert L_.vit_select_algo(4096, 100, 361) == _lib.ALGO_TMEM
    assert L_.vit_select_algo(8, 10, 1100) == _lib.ALGO_STREAM


def test_wave_decoder_single_emission_buffer(Decoder):
    """When two emission buffers leave less than one quantum per wave, WaveDecoder falls back to one buffer (fill and
    decode of consecutive waves then serialise on it)."""
    from viterbi_spl_b200.waves import WaveDecoder, wave_bytes_per_clip
    A, pi = synth.dyadic_hmm(97, seed=4)
    B, T, S = 19, 21, 97
    E = synth.batch('dyadic', B, T, S, seed0=40)
    want_p, want_s = c_oracle.decode_batch_c(A, pi, E)
    dec = Decoder(A, pi, algo='tmem')
    dE = torch.from_numpy(E).cuda()
    wd = WaveDecoder(dec, T, budget_bytes=5 * wave_bytes_per_clip(T, S, 1))
    wd.quantum, wd.emission_buffers, wd.bytes_per_clip = 5, 1, wave_bytes_per_clip(T, S, 1)
    got_p = np.zeros((B, T), np.int64)
    got_s = np.zeros(B, np.float32)

    def sink(a, b, paths, scores):
        got_p[a:b] = paths.cpu().numpy()
        got_s[a:b] = scores.cpu().numpy()

    def fill(a, b, out):
        out.copy_(dE[a:b])

    waves = wd.run(B, fill, sink)
    assert [b - a for a, b in waves] == [5, 5, 5, 4] and wd._emis[1] is None
    assert np.array_equal(got_p, want_p) and np.array_equal(got_s, want_s)
