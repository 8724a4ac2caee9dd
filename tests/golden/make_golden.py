"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN FUNCTIONS (AST-loaded from /root/reference; see
oracle/ref_loader.py) on seeded inputs.  Run in the build container (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Every file stores the LOG-DOMAIN inputs the decoder sees (so the expected paths do not depend on the libm of the
machine that replays them) and the reference outputs.  Inputs that are exact by construction (dyadic values) are
regenerated from their seed and pinned with a sha256 instead of being stored.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader as rl  # noqa: E402
from viterbi_spl_b200 import hmm_params, synth  # noqa: E402


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def save(name, **kw):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **kw)
    print(f'{name:42s} {os.path.getsize(path) / 1024:8.1f} KiB')


def main():
    assert rl.available(), 'needs the reference checkout at /root/reference'
    ref_log = rl.log_domain_decode()                      # imm/tf_viterbi.py:75-109, oracle of record
    ref_a = rl.family_a_decode()                          # dcnet/softmax_viterbi.py:2433-2485
    ref_c_fn = rl.dcnet_c_decode()                        # dcnet/tf_viterbi_decoding.py:156-207
    ref_f64 = rl.dcnet_f64_decode()                       # dcnet/tf_viterbi_decoding.py:209-263
    numba_core = rl.numba_core()                          # dcnet/tf_viterbi_decoding.py:75-116

    # 1. shipped msnet parameters (the only .dat fixtures in the reference), log-domain as the class ctor makes them
    sv = rl.msnet_softmax_viterbi(scaled=False)           # msnet/viterbi_softmax.py:1732-1910
    logA_T = np.array(sv.log_transition_matrix_T)
    log_pi = np.array(sv.log_ini_probs)
    T = 300
    E_dense = synth.dense_softmax(T, 321, seed=101)
    E_sparse = synth.sparse_peaks(T, 321, seed=102)
    st_dense = ref_log(log_transition_matrix_T=logA_T, log_prob_init=log_pi, log_probs_st=np.require(E_dense.T, requirements=['C']))
    st_sparse = ref_log(log_transition_matrix_T=logA_T, log_prob_init=log_pi, log_probs_st=np.require(E_sparse.T, requirements=['C']))
    save('msnet_logdomain.npz', logA_T=logA_T, log_pi=log_pi, ini_probs=np.array(sv.ini_probs),
         E_dense=E_dense, E_sparse=E_sparse,
         states_dense=st_dense, states_sparse=st_sparse)

    # 2. exact-by-construction inputs (no libm): dyadic and tie-stress, S = 361 and 722, plus T = 1 / T = 2
    out = {}
    for tag, S, T, coarse in [('dyadic361', 361, 250, False), ('ties361', 361, 250, True),
                              ('dyadic722', 722, 60, False), ('ties97', 97, 400, True),
                              ('t1', 33, 1, True), ('t2', 33, 2, True)]:
        A, pi = synth.dyadic_hmm(S, seed=S + T, coarse=coarse)
        E = synth.tie_stress((T, S), seed=7 * S + T) if coarse else synth.dyadic((T, S), seed=7 * S + T)
        st = ref_log(log_transition_matrix_T=A, log_prob_init=pi, log_probs_st=np.require(E.T, requirements=['C']))
        out[tag + '_states'] = st
        out[tag + '_sha'] = np.array(sha(A, pi, E))
        out[tag + '_spec'] = np.array([S, T, int(coarse)])
    save('exact_inputs.npz', **out)

    # 3. Family A (prob-domain in): static class method, the stand-alone "c" function, the float64-table variant and
    #    the numba core must all agree; the logs taken here are stored so a replay can detect a different libm
    rng = np.random.default_rng(5)
    S, T = 64, 200
    A = rng.random((S, S)).astype(np.float32) ** 4
    A[rng.random((S, S)) < 0.5] = 0
    A[np.arange(S), np.arange(S)] += 0.1
    A = (A / A.sum(1, keepdims=True)).astype(np.float32)
    pi = rng.random(S).astype(np.float32)
    pi = (pi / pi.sum()).astype(np.float32)
    probs_st = np.asfortranarray(np.exp(synth.sparse_peaks(T, S, seed=6)).T.astype(np.float32))
    probs_st[probs_st < 1e-30] = 0
    st_a = ref_a(transition_matrix=A, prob_init=pi, probs_st=probs_st)
    st_c = ref_c_fn(transition_matrix=A, prob_init=pi, probs_st=probs_st)
    st_64 = ref_f64(transition_matrix=A, prob_init=pi, probs_st=probs_st)
    st_nb = numba_core(np.require(A.T, requirements=['C']).copy(), pi.copy(), np.require(probs_st.T, requirements=['C']).copy())
    assert np.array_equal(st_a, st_c) and np.array_equal(st_a, st_nb)
    tiny = np.finfo(np.float32).tiny
    save('family_a.npz', A=A, pi=pi, probs_st=probs_st, states=st_a, states_f64=st_64,
         logA_T=np.require(np.log(A.T + tiny), requirements=['C']), log_pi=np.log(pi + tiny),
         log_probs_ts=np.require(np.log(probs_st.T + tiny), requirements=['C']))

    # 4. Family C end to end with the shipped parameters: logits -> observation_probs_fn -> decode -> (voiced, bins)
    T = 200
    logits = (1.5 * rng.standard_normal((T, 321))).astype(np.float32)
    tr = synth.pitch_track(T, 320, rng)
    for t in range(T):
        if tr[t] < 320:
            logits[t, 1 + tr[t]] += 5.0
            logits[t, 0] -= 1.0
        else:
            logits[t, 0] += 3.0
    res = {'logits': logits}
    for scaled in (False, True):
        m = rl.msnet_softmax_viterbi(scaled=scaled)
        prob_ts = m.observation_probs_fn(logits.copy())
        voiced, bins = m(logits.copy())
        res[f'prob_ts_{int(scaled)}'] = prob_ts
        res[f'voiced_{int(scaled)}'] = voiced
        res[f'bins_{int(scaled)}'] = bins
        logp = np.log(prob_ts + tiny)
        res[f'log_prob_ts_{int(scaled)}'] = logp
    save('msnet_softmax_viterbi.npz', **res)

    # 5. Family D: fully dense IMM matrix (S = 722), uniform pi, log-HF0 emissions
    cls = rl.imm_viterbi_class()
    imm = cls(20, 721)                                     # imm/tf_imm.py:51-68
    T = 40
    log_HF0 = np.require((-np.abs(3.0 * rng.standard_normal((722, T)))).astype(np.float32), requirements=['C'])
    st = imm.viterbi_librosa_fn(log_HF0)
    save('imm_dense.npz', log_HF0=log_HF0, states=st, logA_T_sha=np.array(sha(imm.log_transition_matrix_T)),
         log_pi=imm.log_prob_init)

    # 6. Family B: tonet class with S = 361 parameters built by the builder recipe (the reference ships none)
    import tempfile
    A361, pi361 = hmm_params.synthetic_hmm('tonet', seed=0)
    saver = rl.dat_saver()
    cls = rl.tonet_viterbi_class()
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        saver(os.path.join(d, 'viterbi_transition_matrix.dat'), A361, 'viterbi_transition_matrix')
        saver(os.path.join(d, 'viterbi_init_probs.dat'), pi361, 'viterbi_init_probs')
        os.chdir(d)
        try:
            tv = cls(0.5)                                  # tonet/softmax_priors.py:1693-1709
        finally:
            os.chdir(cwd)
    T = 150
    logits = (1.5 * rng.standard_normal((T, 360))).astype(np.float32)
    tr = synth.pitch_track(T, 360, rng)
    for t in range(T):
        if tr[t] < 360:
            logits[t, tr[t]] += 5.0
    probs_st = tv.observation_probs_fn(logits.copy())      # [S, T] F-order, prob-domain
    probs_keep = probs_st.copy(order='F')
    st = tv.viterbi_librosa_fn(probs_st)                   # logs in place
    voiced, bins = tv(logits.copy())
    save('tonet_family_b.npz', A=A361, pi=pi361, logits=logits, probs_st=probs_keep, log_probs_st=probs_st,
         states=st, voiced=voiced, bins=bins, logA_T=np.array(tv.log_transition_matrix_T),
         log_pi=np.array(tv.log_ini_probs))


def class_calls():
    """7. The class-level call the pipelines make (``voiced, bins = viterbi_ins(logits)``, dcnet/softmax_viterbi.py:3039)
    for every copy of ``class Viterbi`` / ``class SoftMaxViterbi`` (tests/class_cases.py lists them), by constructing the
    reference's own class in a directory holding the parameter files and calling its ``observation_probs_fn`` and
    ``__call__``; plus imm/tf_imm.py's HF0 decoder.  -> class_calls.npz"""
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import class_cases as cc
    tiny = np.finfo(np.float32).tiny
    out = {}
    dirs = {}
    keep = []
    for state_set in ('msnet_shipped', 'tonet', 'jdc', 'imm_hmm'):
        A, pi = cc.parameters(state_set, rl.REF_ROOT)
        d = tempfile.TemporaryDirectory()
        keep.append(d)
        cc.write_dat(d.name, A, pi)
        dirs[state_set] = d.name
        out[f'A_{state_set}'] = A
        out[f'pi_{state_set}'] = pi
    for case in cc.CASES:
        ns, name, args, state_set, _ = case
        tag = cc.case_tag(case)
        ref = rl.construct_in(dirs[state_set], rl.reference_class(ns, name), *cc.ctor_args(args, rl.FakeTF.Variable))
        logits = cc.logits_for(case, seed=sum(map(ord, tag)))
        probs = ref.observation_probs_fn(logits.copy())
        voiced, bins = ref(logits.copy())
        out[f'{tag}_logits'] = logits
        out[f'{tag}_probs'] = probs
        out[f'{tag}_log_probs'] = np.log(probs + tiny)                  # what the decoder sees (machine-independent replay)
        out[f'{tag}_voiced'] = voiced
        out[f'{tag}_bins'] = bins
        out[f'{tag}_exp_probe'] = np.exp(logits[:4].astype(np.float32))  # detects a different libm on replay
        print(f'  {tag:44s} T={len(bins):4d} voiced {voiced.mean():.2f}')
    imm = rl.imm_viterbi_class()(20, 721)
    HF0 = cc.hf0_case(5)
    log_HF0 = imm.process_HF0_fn(HF0)
    out['imm_HF0Viterbi_HF0'] = HF0
    out['imm_HF0Viterbi_log_HF0'] = log_HF0
    out['imm_HF0Viterbi_states'] = imm(HF0)
    out['imm_HF0Viterbi_logA_T_sha'] = np.array(sha(imm.log_transition_matrix_T))
    save('class_calls.npz', **out)


def melody_stats():
    """8. The statistics step after the decode: the reference's OWN MetricsInference.viterbi_update_states_tf_fn +
    MetricsBase.est_notes_fn / octave / count_nonzero_fn (dcnet/softmax_viterbi.py:1919-1958, 2923-2979), executed on
    the NumPy-backed stand-in for the TensorFlow ops they call (oracle/ref_loader.FakeTFStats).  -> melody_stats.npz"""
    fn = rl.dcnet_melody_stats()
    out = {}
    for k, (seed, T) in enumerate(((1, 300), (2, 1), (3, 129), (4, 400))):
        rng = np.random.default_rng(seed)
        logits = rng.normal(0, 3, (T, 320)).astype(np.float32)
        bins = rng.integers(0, 320, T).astype(np.int32)
        bins[:4] = np.asarray((0, 319, 1, 318))[:T]                              # window clipped at both ends of the bin range
        voiced = rng.random(T) < 0.6
        ref = (23.6 + bins / 5. + rng.normal(0, 0.4, T)).astype(np.float32)      # around the decoded bin: hits and misses
        ref[rng.random(T) < 0.15] += 12.                                         # octave errors: chroma hit, pitch miss
        ref[rng.random(T) < 0.3] = 0.                                            # unvoiced reference frames
        est, counters = fn(ref, logits, bins, voiced)
        out[f'c{k}_logits'], out[f'c{k}_bins'], out[f'c{k}_voiced'], out[f'c{k}_ref'] = logits, bins, voiced, ref
        out[f'c{k}_est'] = est
        out[f'c{k}_counters'] = np.asarray([counters[n] for n in ('gt_voiced', 'gt_unvoiced', 'voicing_correct_voiced',
                                                                  'voicing_incorrect_voiced', 'voicing_correct_unvoiced',
                                                                  'correct_pitches_wide', 'correct_pitches_strict',
                                                                  'correct_chromas_wide', 'correct_chromas_strict')], np.int64)
    out['n_cases'] = np.array(4)
    save('melody_stats.npz', **out)


if __name__ == '__main__':
    which = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if which in ('all', 'main'):
        main()
    if which in ('all', 'class_calls'):
        class_calls()
    if which in ('all', 'melody_stats'):
        melody_stats()
