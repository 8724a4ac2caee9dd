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
B          tonet/softmax_priors.py:1841 ``Viterbi.viterbi_librosa_fn``   ``ViterbiB.viterbi_librosa_fn``
C          dcnet/softmax_viterbi.py:2636 ``SoftMaxViterbi...``           ``SoftMaxViterbi.viterbi_librosa_fn``
D (class)  imm/tf_imm.py:90 ``Viterbi.viterbi_librosa_fn``               ``ImmViterbi.viterbi_librosa_fn``
=========  ===========================================================  =======================================
"""
import hashlib
import zlib
from collections import OrderedDict

import numpy as np

from . import hmm_params
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


def _decode_ts(log_transition_matrix_T, log_prob_init, log_probs_ts):
    """[T, S] log emissions -> int64[T] via the GPU decoder."""
    E = np.require(log_probs_ts, np.float32, ['C'])
    assert not np.isnan(E.min()), 'emissions contain NaN'            # (min propagates NaN: one pass, no temporary)
    paths, _ = _decoder_for(log_transition_matrix_T, log_prob_init).decode_host(E[None])
    return paths[0]


def _decode_st(log_transition_matrix_T, log_prob_init, log_probs_st):
    """[S, T] log emissions, C-contiguous -> int64[T].  The transpose the reference does on the host
    (imm/tf_viterbi.py:89, a strided copy of the whole table) happens on the GPU after the upload instead."""
    E = log_probs_st
    if not E.flags['C_CONTIGUOUS']:
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
        tinyp = np.float32(1.1754944e-38)                            # :18
        S = B.shape[0]
        assert prob_init.shape[0] == S                               # :22
        B[:] = np.log(B + tinyp)
        prob_init[:] = np.log(prob_init + tinyp)
        probs[:] = np.log(probs + tinyp)
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


# ---- decoder objects (Families A-D) -----------------------------------------------------------------------------------

class Viterbi:
    """Family A object (dcnet/softmax_viterbi.py:2273-2485): linear parameters, static decode."""

    def __init__(self, transition_matrix, ini_probs, num_freq_bins=None):
        t = np.sum(transition_matrix, axis=1)
        assert np.all(np.isclose(t, 1))                               # :2413-2414
        assert np.all(ini_probs > 0)                                  # :2381
        self.transition_matrix = transition_matrix
        self.ini_probs = ini_probs
        self.num_freq_bins = len(ini_probs) - 1 if num_freq_bins is None else num_freq_bins

    @staticmethod
    def viterbi_librosa_fn(*, transition_matrix, prob_init, probs_st):
        """dcnet/softmax_viterbi.py:2433-2485 (11 copies across the reference)."""
        return _family_a(transition_matrix, prob_init, probs_st)

    def decode_probs(self, observation_probs):
        """The tail of ``__call__`` (:2421-2431): decode, then voiced = bins < n_bins; bins = min(bins, n_bins-1)."""
        bins = Viterbi.viterbi_librosa_fn(transition_matrix=self.transition_matrix, prob_init=self.ini_probs,
                                          probs_st=observation_probs)
        return voiced_and_bins(bins, self.num_freq_bins)


class _LogParamDecoder:
    """Shared ctor of Families B/C: log(x + tiny), transpose, float32 C-contiguous, read-only
    (tonet/softmax_priors.py:1788-1823; dcnet/softmax_viterbi.py:2581-2618)."""

    def __init__(self, transition_matrix, ini_probs, num_freq_bins=None):
        U = len(ini_probs) - 1 if num_freq_bins is None else num_freq_bins
        self.num_freq_bins = U
        assert ini_probs.shape == (U + 1,)
        assert np.isclose(np.sum(ini_probs), 1)
        assert transition_matrix.shape == (U + 1, U + 1)
        assert np.all(np.isclose(np.sum(transition_matrix, axis=1), 1))
        self.ini_probs = ini_probs
        self.log_transition_matrix_T, self.log_ini_probs = hmm_params.log_params(transition_matrix, ini_probs)
        self.log_transition_matrix_T.flags['WRITEABLE'] = False
        self.log_ini_probs.flags['WRITEABLE'] = False
        self._decoder = ViterbiDecoder(self.log_transition_matrix_T, self.log_ini_probs)

    @classmethod
    def from_dat(cls, directory='.', **kw):
        """Load ``viterbi_transition_matrix.dat`` / ``viterbi_init_probs.dat`` like the reference ctors do."""
        import os
        name, A = hmm_params.load_dat(os.path.join(directory, 'viterbi_transition_matrix.dat'))
        assert name == 'viterbi_transition_matrix'
        name, pi = hmm_params.load_dat(os.path.join(directory, 'viterbi_init_probs.dat'))
        assert name == 'viterbi_init_probs'
        return cls(A, pi, **kw)

    def _decode(self, log_probs_ts):
        paths, _ = self._decoder.decode_host(np.require(log_probs_ts, np.float32, ['C'])[None])
        return paths[0]

    def decode_probs(self, probs):
        return voiced_and_bins(self.viterbi_librosa_fn(probs), self.num_freq_bins)


class ViterbiB(_LogParamDecoder):
    """Family B (tonet/softmax_priors.py:1691-1878 and the imm-HMM copies)."""

    def viterbi_librosa_fn(self, probs_st):
        """``probs_st [S, T]`` float32 F-contiguous, prob-domain; LOGGED IN PLACE (:1854-1857)."""
        S = self.num_freq_bins + 1
        assert probs_st.shape[0] == S
        assert probs_st.dtype == np.float32
        assert probs_st.flags['F_CONTIGUOUS'] == True  # noqa: E712
        np.add(probs_st, TINY, out=probs_st)
        np.log(probs_st, out=probs_st)
        probs = np.require(probs_st.T, np.float32, ['C'])
        return self._decode(probs)


class SoftMaxViterbi(_LogParamDecoder):
    """Family C (dcnet/softmax_viterbi.py:2488-2674 and 6 more copies): the north star's "posteriors [T, N]" layout."""

    def __init__(self, transition_matrix, ini_probs, scaled=False, num_freq_bins=None):
        super().__init__(transition_matrix, ini_probs, num_freq_bins)
        assert np.argmax(ini_probs) == self.num_freq_bins              # unvoiced (last) is the most likely start, :2589
        assert np.all(ini_probs > 0)
        self.scaled = scaled

    def viterbi_librosa_fn(self, probs_ts):
        """``probs_ts [T, S]`` float32 C-contiguous, prob-domain (may exceed 1 when scaled); LOGGED IN PLACE (:2650-2653)."""
        S = self.num_freq_bins + 1
        assert probs_ts.ndim == 2
        assert probs_ts.shape[1] == S
        assert probs_ts.dtype == np.float32
        assert probs_ts.flags['C_CONTIGUOUS']
        np.add(probs_ts, TINY, out=probs_ts)
        np.log(probs_ts, out=probs_ts)
        return self._decode(probs_ts)


class ImmViterbi:
    """Family D object (imm/tf_imm.py:48-135): fully dense transition matrix from the IMM recipe, uniform pi."""

    def __init__(self, bins_per_semitone, n_bins):
        self.b = bins_per_semitone
        self.n_bins = n_bins
        A = hmm_params.dense_imm_transition_matrix(bins_per_semitone, n_bins)      # :54
        assert np.all(A > 0)
        init = np.full([n_bins + 1], 1. / (n_bins + 1))                            # :62-64
        self.log_transition_matrix_T, self.log_prob_init = hmm_params.log_params(A, init, add_tiny=False)
        self._decoder = ViterbiDecoder(self.log_transition_matrix_T, self.log_prob_init)

    def viterbi_librosa_fn(self, log_HF0):
        """``log_HF0 [S, T]`` float32 log-domain (:90-127)."""
        S = self.n_bins + 1
        assert isinstance(log_HF0, np.ndarray)
        assert log_HF0.dtype == np.float32
        assert log_HF0.shape[0] == S
        probs = np.require(np.transpose(log_HF0), np.float32, ['C'])
        paths, _ = self._decoder.decode_host(probs[None])
        return paths[0]


def voiced_and_bins(states, n_bins):
    """Post-decode step shared by every ``__call__``: unvoiced is the last state
    (dcnet/softmax_viterbi.py:2427-2431)."""
    voiced = states < n_bins
    bins = np.minimum(states, n_bins - 1)
    return voiced, bins
