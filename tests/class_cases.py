"""Shared by tests/test_reference_classes.py, tests/test_gpu_reference_classes.py and tests/golden/make_golden.py: the
parameter sets and logits the class-level drop-ins (viterbi_spl_b200.reference_classes) are exercised on, one case per
copy of the reference's ``class Viterbi`` / ``class SoftMaxViterbi``."""
import os

import numpy as np

from viterbi_spl_b200 import hmm_params, synth

# (namespace, class name, ctor args, state set whose parameters it loads, frames in the golden)
CASES = [
    ('dcnet', 'Viterbi', (), 'msnet_shipped', 150),
    ('dcnet', 'SoftMaxViterbi', ('VAR:0.31', True), 'msnet_shipped', 150),
    ('dcnet', 'SoftMaxViterbi', ('VAR:0.31', False), 'msnet_shipped', 150),
    ('msnet', 'Viterbi', (0.4,), 'msnet_shipped', 150),
    ('msnet', 'SoftMaxViterbi', (True,), 'msnet_shipped', 150),
    ('ftanet', 'Viterbi', (-0.3,), 'msnet_shipped', 150),
    ('ftanet', 'SoftMaxViterbi', (False,), 'msnet_shipped', 150),
    ('tonet', 'Viterbi', (0.5,), 'tonet', 150),
    ('tonet', 'ViterbiA', (0.35,), 'tonet', 150),
    ('tonet', 'SoftMaxViterbi', (True,), 'tonet', 150),
    ('tonet', 'SoftMaxViterbiAlwaysScaled', (), 'tonet', 150),
    ('tonet', 'SoftMaxViterbiWide', (), 'tonet', 150),
    ('jdc', 'Viterbi', (0.5,), 'jdc', 60),
    ('jdc', 'SoftMaxViterbi', (True,), 'jdc', 60),
    ('jdc', 'SoftMaxViterbi', (False,), 'jdc', 60),
    ('imm', 'Viterbi', (), 'imm_hmm', 60),
]


def case_tag(case):
    ns, name, args, _, _ = case
    return '_'.join([ns, name] + [str(a).replace('VAR:', 'v').replace('.', 'p').replace('-', 'm') for a in args])


def exact_unit_sum(pi):
    """Family-A ``init_probs_fn`` asserts ``np.sum(probs) == 1.`` exactly (dcnet/softmax_viterbi.py:2380): nudge the
    largest entry by ulps until the float32 sum is exactly 1."""
    pi = np.array(pi, np.float32)
    k = int(np.argmax(pi))
    for _ in range(64):
        s = np.sum(pi)
        if s == 1.:
            return pi
        pi[k] = np.nextafter(pi[k], np.float32(0 if s > 1 else 2), dtype=np.float32)
    raise AssertionError('could not make the initial distribution sum to exactly 1')


def parameters(state_set, ref_root='/root/reference'):
    """(A [S, S] float32 row-stochastic, pi [S] float32) of a state set.  'msnet_shipped' are the only parameter files
    the reference ships (msnet/*.dat, S = 321); the others come from the builder recipes on synthetic statistics."""
    if state_set == 'msnet_shipped':
        _, A = hmm_params.load_dat(os.path.join(ref_root, 'msnet', 'viterbi_transition_matrix.dat'))
        _, pi = hmm_params.load_dat(os.path.join(ref_root, 'msnet', 'viterbi_init_probs.dat'))
        return np.array(A), np.array(pi)
    A, pi = hmm_params.synthetic_hmm(state_set)
    return np.asarray(A, np.float32), exact_unit_sum(pi)


def write_dat(directory, A, pi):
    hmm_params.save_dat(os.path.join(directory, 'viterbi_transition_matrix.dat'), np.ascontiguousarray(A),
                        'viterbi_transition_matrix')
    hmm_params.save_dat(os.path.join(directory, 'viterbi_init_probs.dat'), np.ascontiguousarray(pi), 'viterbi_init_probs')


def logits_for(case, seed, T=None):
    """Acoustic-model-like logits for a case: noise + a strong ridge along a random-walk pitch track with unvoiced
    gaps (and frames with no voiced peak at all, and exact ties between neighbouring bins, so both special cases of the
    peak finders are exercised).  Layout as the class expects it."""
    ns, name, _, state_set, T0 = case
    T = T0 if T is None else T
    n_bins = {'msnet_shipped': 320, 'tonet': 360, 'jdc': 721, 'imm_hmm': 721}[state_set]
    rng = np.random.default_rng(seed)
    softmax = name.startswith('SoftMax')
    with_unvoiced = softmax and ns != 'dcnet'
    x = (1.5 * rng.standard_normal((T, n_bins))).astype(np.float32)
    tr = synth.pitch_track(T, n_bins, rng)
    voiced = tr < n_bins
    x[np.nonzero(voiced)[0], tr[voiced]] += np.float32(5.0)
    flat = rng.random(T) < 0.08                              # monotone frames: no interior peak (shaun: edge peaks only)
    x[flat] = np.linspace(-3, 3, n_bins, dtype=np.float32)[None] * rng.choice([-1, 1], size=(int(flat.sum()), 1)).astype(np.float32)
    x[rng.random(T) < 0.05] = np.float32(-4.0)               # constant frames: every window is tied -> first max wins
    ties = rng.random((T, n_bins)) < 0.02                    # exact ties between neighbours
    x[:, 1:][ties[:, 1:]] = x[:, :-1][ties[:, 1:]]
    x = np.round(x * 64) / 64                                # coarse grid: more exact ties
    x = x.astype(np.float32)
    if with_unvoiced:
        u = np.where(voiced, -1.0, 3.0) + 0.5 * rng.standard_normal(T)
        x = np.concatenate([u[:, None].astype(np.float32), x], axis=1)
    if ns == 'imm':
        x = np.ascontiguousarray(x.T)                        # imm/main_imm.py:186: [n_bins, T]
    return np.ascontiguousarray(x, np.float32)


def ctor_args(args, make_var):
    """'VAR:x' stands for the tf.Variable the dcnet SoftMaxViterbi constructor takes."""
    return tuple(make_var(float(a[4:])) if isinstance(a, str) and a.startswith('VAR:') else a for a in args)


class Var:
    """Anything with ``.numpy()`` is accepted where the reference takes a tf.Variable."""

    def __init__(self, v):
        self.v = np.float32(v)

    def numpy(self):
        return self.v


def hf0_case(seed, n_bins=721, T=50):
    """A source-filter-model-like HF0 activation matrix [n_bins, T] (imm/tf_imm.py:70): non-negative, many exact zeros."""
    rng = np.random.default_rng(seed)
    h = np.abs(rng.standard_normal((n_bins, T))) ** 3
    h[rng.random((n_bins, T)) < 0.3] = 0
    tr = synth.pitch_track(T, n_bins, rng)
    v = tr < n_bins
    h[tr[v], np.nonzero(v)[0]] += 30.0
    return h.astype(np.float32)
