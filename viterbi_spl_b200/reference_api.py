"""Drop-in replacements for the reference's decode entry points.

Same names, keyword-only signatures, array layouts, dtype of the result (``np.int64[T]``), in-place side effects and
``AssertionError`` behaviour as the reference functions they replace; the recursion + backtrace run on the GPU
through ``ViterbiDecoder`` (B = 1).  Where the reference takes ``log(x + tiny)`` of prob-domain inputs it does so
with NumPy on the host, and so do these wrappers -- same process, same libm, same bits -- because NumPy's float32 log
is not correctly rounded and a GPU ``logf`` could flip near-ties (hard part 1).

=========  ===========================================================  =======================================
family     reference entry point                                        here
=========  ===========================================================  =======================================
log / D    imm/tf_viterbi.py:75 ``viterbi_librosa_fn``                  ``viterbi_librosa_fn``
A          dcnet/softmax_viterbi.py:2433 ``Viterbi.viterbi_librosa_fn``  ``Viterbi.viterbi_librosa_fn`` (static),
           dcnet/tf_viterbi_decoding.py:156 ``viterbi_librosa_c_fn``     ``viterbi_librosa_c_fn``
numba      dcnet/tf_viterbi_decoding.py:119 ``viterbi_numba_fn``,        ``viterbi_numba_fn``, ``viterbi_numba.core``
           dcnet/aot_viterbi_core.py:8 ``viterbi_numba.core``
TF         dcnet/tf_viterbi_decoding.py:23 ``viterbi_tf_fn``,            ``viterbi_tf_fn``, ``tf_viterbi_librosa_fn``
           imm/tf_viterbi.py:8 ``tf_viterbi_librosa_fn``                 (NumPy in, int32 out)
f64 table  dcnet/tf_viterbi_decoding.py:209 ``viterbi_librosa_fn``        ``viterbi_librosa_f64_fn`` (same paths, fp32)
classes    ``Viterbi`` / ``SoftMaxViterbi`` of every directory           ``dcnet. msnet. ftanet. jdc. tonet. imm.``
           (ctor, observation_probs_fn, __call__ -> voiced, bins)        namespaces (reference_classes.py)
B          tonet/softmax_priors.py:1841 ``Viterbi.viterbi_librosa_fn``   ``tonet.Viterbi.viterbi_librosa_fn``
C          dcnet/softmax_viterbi.py:2636 ``SoftMaxViterbi...``           ``dcnet.SoftMaxViterbi.viterbi_librosa_fn``
D (class)  imm/tf_imm.py:90 ``Viterbi.viterbi_librosa_fn``               ``imm.HF0Viterbi.viterbi_librosa_fn``
=========  ===========================================================  =======================================
"""
import hashlib
import zlib
from collections import OrderedDict

import numpy as np

from .decoder import ViterbiDecoder

TINY = np.finfo(np.float32).tiny
_DECODERS = OrderedDict()
_MAX_CACHED = 8


def _content_key(x):
    """Content fingerprint of a float32 array at close to memory speed: CRC-32 of the bytes plus a digest of the
    per-row sums of the bit patterns (a cryptographic hash of the 2 MB matrix of a 722-state model costs 4 ms -- as much
    as decoding a recording)."""
    x = np.ascontiguousarray(x)
    rows = x.reshape(x.shape[0], -1).view(np.uint32).sum(axis=1, dtype=np.uint64)
    return (x.shape, zlib.crc32(memoryview(x).cast('B')), hashlib.blake2b(rows.tobytes(), digest_size=16).digest())


def _decoder_for(log_transition_matrix_T, log_prob_init):
    """Device-resident parameters are cached by content so per-call families do not re-upload the matrix."""
    A = np.require(log_transition_matrix_T, np.float32, ['C'])
    pi = np.require(log_prob_init, np.float32, ['C'])
    key = (_content_key(A), hashlib.blake2b(pi.tobytes(), digest_size=16).digest())
    dec = _DECODERS.get(key)
    if dec is None:
        dec = ViterbiDecoder(A, pi)
        _DECODERS[key] = dec
        while len(_DECODERS) > _MAX_CACHED:
            _DECODERS.popitem(last=False)
    else:
        _DECODERS.move_to_end(key)
    return dec


def _float32_parameters(B, prob_init, first_row):
    """The recursion the GPU runs is float32 throughout, like the reference's when its parameters are float32 (what its
    .dat files hold).  With a float64 ``prob_init`` the reference's ``T1[0] = prob_init + probs[0]`` (imm/tf_viterbi.py:94)
    adds in float64 and rounds ONCE into the float32 table: reproduced here by doing that one row on the host and handing
    the decoder a zero initial vector (0 + x is exact).  A float64 transition matrix would make every ``T1[t-1] + B`` of
    the reference a float64 add rounded to float32 -- not what this decoder computes, so it is refused rather than
    silently rounded twice."""
    assert B.dtype == np.float32, ('transition parameters must be float32 (a float64 matrix makes the reference add in '
                                   'float64 at every step; cast it to float32 first)')
    prob_init = np.asarray(prob_init)
    if prob_init.dtype == np.float32:
        return prob_init, None
    row0 = (prob_init + first_row).astype(np.float32)                # float64 add, one rounding
    return np.zeros(len(prob_init), np.float32), row0


def _decode_ts(log_transition_matrix_T, log_prob_init, log_probs_ts):
    """[T, S] log emissions -> int64[T] via the GPU decoder."""
    E = np.require(log_probs_ts, np.float32, ['C'])
    assert not np.isnan(E.min()), 'emissions contain NaN'            # (min propagates NaN: one pass, no temporary)
    log_prob_init, row0 = _float32_parameters(log_transition_matrix_T, log_prob_init, E[0] if len(E) else None)
    if row0 is not None and len(E):
        E = E.copy()
        E[0] = row0
    paths, _ = _decoder_for(log_transition_matrix_T, log_prob_init).decode_host(E[None])
    return paths[0]


def _decode_st(log_transition_matrix_T, log_prob_init, log_probs_st):
    """[S, T] log emissions, C-contiguous -> int64[T].  The transpose the reference does on the host
    (imm/tf_viterbi.py:89, a strided copy of the whole table) happens on the GPU after the upload instead."""
    E = log_probs_st
    if not E.flags['C_CONTIGUOUS'] or np.asarray(log_prob_init).dtype != np.float32:
        return _decode_ts(log_transition_matrix_T, log_prob_init, np.require(E.T, np.float32, ['C']))
    assert not np.isnan(E.min()), 'emissions contain NaN'
    paths, _ = _decoder_for(log_transition_matrix_T, log_prob_init).decode_host_st(E)
    return paths


# ---- log-domain (Family D) ----------------------------------------------------------------------------------------

def viterbi_librosa_fn(*, log_transition_matrix_T, log_prob_init, log_probs_st):
    """imm/tf_viterbi.py:75-109.  ``log_probs_st`` is ``[S, T]``; returns ``np.int64[T]``."""
    B = log_transition_matrix_T
    assert B.flags['C_CONTIGUOUS'] == True  # noqa: E712  (imm/tf_viterbi.py:78)
    assert B.dtype == np.float32
    S = len(B)
    assert len(log_prob_init) == S
    assert log_probs_st.dtype == np.float32
    assert log_probs_st.shape[0] == S
    return _decode_st(B, log_prob_init, log_probs_st)                # (:89's transpose runs on the GPU)


def tf_viterbi_librosa_fn(*, tf_log_transition_matrix_T, tf_log_prob_init, tf_or_np_log_probs_st):
    """imm/tf_viterbi.py:8-72 for NumPy inputs (TensorFlow is not a dependency here): the log-domain decode with the
    eager-TF argument names, shape asserts of :25-29 and the ``np.int32[T]`` result of :62.  ``tf.argmax`` documents no
    tie order; first-max-wins (NumPy, a1) is the pinned behaviour."""
    B = np.asarray(tf_log_transition_matrix_T, np.float32)
    probs = np.asarray(tf_or_np_log_probs_st, np.float32)            # (:21 convert_to_tensor(..., tf.float32))
    prob_init = np.asarray(tf_log_prob_init, np.float32)
    S = len(B)
    assert B.shape == (S, S)
    assert len(prob_init) == S
    T = probs.shape[1]
    assert probs.shape == (S, T)
    return _decode_st(np.require(B, np.float32, ['C']), prob_init, probs).astype(np.int32)


# ---- Family A: prob-domain in, logs taken on every call -------------------------------------------------------------

def _family_a(transition_matrix, prob_init, probs_st):
    B = transition_matrix
    probs = probs_st
    S = len(B)
    T = probs.shape[1]
    assert B.shape == (S, S)                                         # dcnet/softmax_viterbi.py:2450-2457
    assert probs.shape == (S, T)
    t = np.sum(B, axis=1)
    assert np.allclose(t, 1.)
    assert len(prob_init) == S
    assert np.isclose(np.sum(prob_init), 1.)
    tinyp = np.finfo(probs.dtype).tiny                               # :2459
    B = np.require(np.log(B.T + tinyp), requirements=['C'])          # :2461-2462  S <- S
    prob_init = np.log(prob_init + tinyp)                            # :2463
    probs = np.require(np.log(probs.T + tinyp), requirements=['C'])  # :2464-2465  T * S
    return _decode_ts(B, prob_init, probs)


def viterbi_librosa_c_fn(*, transition_matrix, prob_init, probs_st):
    """dcnet/tf_viterbi_decoding.py:156-207 (inputs are not mutated)."""
    return _family_a(transition_matrix, prob_init, probs_st)


def viterbi_librosa_f64_fn(*, transition_matrix, prob_init, probs_st):
    """dcnet/tf_viterbi_decoding.py:209-263 -- the module's plain ``viterbi_librosa_fn``, whose ``T1`` table is float64
    (``np.empty([T, S])``, :242), so ITS recursion adds in double.  Same signature and checks; the decode here is the
    float32 one of ``viterbi_librosa_c_fn`` (the parity target): the two agree wherever no
    two candidate paths are within float32 rounding of each other (equal on the golden inputs,
    tests/golden/family_a.npz ``states_f64``), but a near-tie can resolve differently than in the float64 table."""
    return _family_a(transition_matrix, prob_init, probs_st)


def viterbi_numba_fn(*, transition_matrix, prob_init, probs_st):
    """dcnet/tf_viterbi_decoding.py:119-153: validates, transposes, calls the compiled core."""
    B = transition_matrix
    probs = probs_st
    S = len(B)
    T = probs.shape[1]
    assert B.shape == (S, S)
    assert probs.shape == (S, T)
    assert np.allclose(np.sum(B, axis=1), 1.)
    assert len(prob_init) == S
    assert np.isclose(np.sum(prob_init), 1.)
    B = np.require(B.T, requirements=['C'])                          # :142
    probs = np.require(probs.T, requirements=['C'])                  # :143
    # like the reference, `probs` is a VIEW of an F-ordered probs_st, so the core's in-place log reaches the caller's
    # array (dcnet/aot_viterbi_core.py:25); only prob_init is copied (:147)
    return viterbi_numba.core(B, prob_init.copy(), probs)


class viterbi_numba:  # noqa: N801  (module name in the reference)
    """Stand-in for the numba AOT module ``viterbi_numba`` (dcnet/aot_viterbi_core.py:4-54)."""

    @staticmethod
    def core(B, prob_init, probs):
        """``i8[:](f4[:, ::1], f4[:], f4[:, ::1])``: B dst-major LINEAR probabilities, probs ``[T, S]``; like the AOT
        export it overwrites all three inputs with their logs (dcnet/aot_viterbi_core.py:23-25)."""
        assert B.dtype == np.float32 and prob_init.dtype == np.float32 and probs.dtype == np.float32
        assert B.flags['C_CONTIGUOUS'] and probs.flags['C_CONTIGUOUS']
        tinyp = 1.1754944e-38                                        # :18 -- a float64 literal: numba promotes B + tinyp
        S = B.shape[0]                                               # to float64, logs in float64 and rounds once to f4
        assert prob_init.shape[0] == S                               # :22
        B[:] = np.log(B.astype(np.float64) + tinyp)
        prob_init[:] = np.log(prob_init.astype(np.float64) + tinyp)
        probs[:] = np.log(probs.astype(np.float64) + tinyp)
        return _decode_ts(B, prob_init, probs)


def viterbi_tf_fn(transition_matrix, prob_init, probs_st):
    """dcnet/tf_viterbi_decoding.py:23-72 for NumPy inputs (TensorFlow is not a dependency here): positional
    arguments, S fixed by the matrix, int32 result like the tf.Variable `states` (:21)."""
    transition_matrix = np.asarray(transition_matrix, np.float32)
    prob_init = np.asarray(prob_init, np.float32)
    probs_st = np.asarray(probs_st, np.float32)
    assert np.allclose(np.sum(transition_matrix, axis=1), 1., atol=1e-6 * transition_matrix.shape[0])  # assert_near :45-48
    assert np.isclose(np.sum(prob_init), 1., atol=1e-5)
    B = np.require(np.log(transition_matrix.T + TINY), np.float32, ['C'])   # :53
    pi = np.log(prob_init + TINY)                                            # :54
    probs = np.require(np.log(probs_st.T + TINY), np.float32, ['C'])         # :55
    return _decode_ts(B, pi, probs).astype(np.int32)


# ---- decoder objects (Families A-D): the classes the pipelines construct and call -----------------------------------
#
# One namespace per experiment directory of the reference, each holding that directory's ``Viterbi`` / ``SoftMaxViterbi``
# with the reference's constructor signature, ``find_peaks_all_at_once_np_fn``, ``observation_probs_fn``,
# ``viterbi_librosa_fn`` and ``__call__(logits) -> (voiced, bins)`` (viterbi_spl_b200/reference_classes.py).  A pipeline
# switches with one import line, e.g. in dcnet/softmax_viterbi.py
#     from viterbi_spl_b200.reference_api import dcnet; Viterbi, SoftMaxViterbi = dcnet.Viterbi, dcnet.SoftMaxViterbi
# instead of the class definitions at :2273 and :2488.
from .reference_classes import HF0Viterbi, dcnet, ftanet, imm, jdc, msnet, tonet  # noqa: E402,F401

Viterbi = dcnet.Viterbi                   # dcnet/softmax_viterbi.py:2273  Viterbi()                 (Family A, static decode)
SoftMaxViterbi = dcnet.SoftMaxViterbi     # dcnet/softmax_viterbi.py:2488  SoftMaxViterbi(voicing_threspold_prob, scaled)  (C)
ViterbiB = tonet.Viterbi                  # tonet/softmax_priors.py:1691   Viterbi(voicing_threshold)                    (B)
ImmViterbi = HF0Viterbi                   # imm/tf_imm.py:48               Viterbi(bins_per_semitone, n_bins)            (D)


def voiced_and_bins(states, n_bins):
    """Post-decode step shared by every ``__call__``: unvoiced is the last state
    (dcnet/softmax_viterbi.py:2427-2431)."""
    voiced = states < n_bins
    bins = np.minimum(states, n_bins - 1)
    return voiced, bins
