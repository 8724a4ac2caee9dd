"""Class-level drop-ins: the decoder objects the reference's pipelines construct and call (SURVEY.md section 8 rows
a10, a11, a12).

The call the reference makes per recording is ``voiced, bins = self.viterbi_ins(logits)``
(dcnet/softmax_viterbi.py:3039), i.e. ``Viterbi.__call__`` / ``SoftMaxViterbi.__call__``: emission table from the
acoustic model's logits (host NumPy), decode, ``voiced = bins < n_bins; bins = min(bins, n_bins - 1)``.  Every
experiment directory carries its own copy of those two classes with its own constants; they are mirrored here as one
namespace per directory with the reference's constructor signatures:

=========  ==================================================  =====================================================
namespace  ``Viterbi`` ("shaun" emission model)                 ``SoftMaxViterbi``
=========  ==================================================  =====================================================
dcnet      ``Viterbi()``  320 bins, th 0.31                      ``SoftMaxViterbi(voicing_threspold_prob, scaled)``
           dcnet/softmax_viterbi.py:2273-2485                    dcnet/softmax_viterbi.py:2488-2674 (pads the
           (main.py, lontano.py, viterbi_states_10ms_step.py)    threshold logit as column 0, :2545-2548)
msnet      ``Viterbi(voicing_threshold)``  320 bins              ``SoftMaxViterbi(scaled)``
           msnet/viterbi_softmax.py:1515-1729, hsieh_m2m3.py     msnet/viterbi_softmax.py:1732-1910
ftanet     ``Viterbi(voicing_threshold_logit)``  320 bins        ``SoftMaxViterbi(scaled)``
           ftanet/viterbi_performance.py:2569-2735               ftanet/viterbi_performance.py:2736-2916
jdc        ``Viterbi(voicing_threshold)``  721 bins, spw 16      ``SoftMaxViterbi(scaled)``
           jdc/viterbi_softmax.py:1902-2091                      jdc/viterbi_softmax.py:2094-2272
tonet      ``Viterbi(voicing_threshold)``  360 bins (Family B)   ``SoftMaxViterbi(scaled)`` :1881-2061; the ``SoftMaxViterbi()``
           tonet/softmax_priors.py:1691-1878                     of ablation.py (always scaled) / for_paper.py (spw 15, never)
imm        ``Viterbi()``  721 bins, spw 20 (Family B)            --   ``imm.HF0Viterbi(bins_per_semitone, n_bins)``
           imm/main_imm.py:141-325 (thresholding.py, ...)        is imm/tf_imm.py:48-135 (Family D, HF0 in)
=========  ==================================================  =====================================================

Like the reference, the constructors read ``viterbi_transition_matrix.dat`` / ``viterbi_init_probs.dat`` from the
working directory (``transition_matrix_fn`` / ``init_probs_fn``); keyword-only extras -- ``directory=``,
``transition_matrix=``, ``ini_probs=`` -- supply the parameters from elsewhere.  ``device_emissions=True`` builds the
emission table on the GPU (``vit_emissions_f32``: peaks exact, values within 1e-5 of NumPy's, so not bit-exact);
the default keeps it on the host in NumPy -- same process, same libm, same bits as the reference -- and only the
recursion + backtrace run on the GPU, which is what makes ``__call__`` reproduce the reference's output bit for bit.

The emission builders restate the reference's per-frame Python loops as NumPy array operations over all frames that
have the same number of peaks (the reference loops over frames one by one); every element goes through the same
ufunc with the same dtypes in the same order, so the tables are bit-identical (tests/test_reference_classes.py checks
this against the reference's own classes executed by oracle/ref_loader.py, and against goldens made by them).
"""
import os
import types

import numpy as np

from . import hmm_params
from .decoder import ViterbiDecoder

TINY = np.finfo(np.float32).tiny


# ---- pieces shared by all copies ----------------------------------------------------------------------------------------

def find_peaks(frames_logits, spw):
    """``[T, n]`` -> bool ``[T, n]``: bin k is a peak iff the FIRST maximum of the reflect-padded window
    ``[k - spw, k + spw]`` is its centre (``find_peaks_all_at_once_np_fn``, tonet/softmax_priors.py:1722-1739; the
    ``tf.argmax`` form dcnet/softmax_viterbi.py:2298-2314 documents no tie order, first-max is the pinned one)."""
    spw = int(spw)
    padded = np.pad(frames_logits, [(0, 0), (spw, spw)], mode='reflect')
    windows = np.lib.stride_tricks.sliding_window_view(padded, 2 * spw + 1, axis=1)
    return np.argmax(windows, axis=2) == spw


def _groups_by_peak_count(are_peaks):
    """Yields (frame indices [F], peak columns [F, n]) for every distinct number of peaks n >= 1 per frame."""
    counts = are_peaks.sum(axis=1)
    for n in np.unique(counts):
        if n == 0:
            continue
        rows = np.nonzero(counts == n)[0]
        cols = np.nonzero(are_peaks[rows])[1].reshape(len(rows), int(n))
        yield rows, cols


def _load_parameters(directory, transition_matrix, ini_probs):
    """``transition_matrix_fn`` / ``init_probs_fn`` without the per-family checks: the two ``.dat`` records the
    reference's constructors read from the working directory (dcnet/softmax_viterbi.py:2374-2417), unless given."""
    if transition_matrix is None:
        name, transition_matrix = hmm_params.load_dat(os.path.join(directory, 'viterbi_transition_matrix.dat'))
        assert name == 'viterbi_transition_matrix'
    if ini_probs is None:
        name, ini_probs = hmm_params.load_dat(os.path.join(directory, 'viterbi_init_probs.dat'))
        assert name == 'viterbi_init_probs'
    return np.asarray(transition_matrix), np.asarray(ini_probs)


def _read_only_log(x, transpose=False):
    t = np.log(x + TINY)                                                  # tonet/softmax_priors.py:1798-1803
    assert not np.any(np.isneginf(t))
    t = np.require(t.T if transpose else t, np.float32, ['C'])
    t.flags['WRITEABLE'] = False
    return t


def _lazy_decoder(obj, logA_T, log_pi):
    """The GPU decoder of a drop-in object, created at the first decode (the emission builders are host NumPy and need
    no device; decoding without the CUDA library or a GPU raises -- there is no CPU fallback)."""
    dec = obj.__dict__.get('_decoder_obj')
    if dec is None:
        dec = obj.__dict__['_decoder_obj'] = ViterbiDecoder(logA_T, log_pi, device=obj.__dict__.get('_device'))
    return dec


def _expit(s):
    """Element-wise form of ``Viterbi.expit`` (tonet/softmax_priors.py:1711-1720): the branch by sign kept."""
    pos = s > 0
    with np.errstate(over='ignore'):
        a = 1. / (1. + np.exp(-s))
        b = np.exp(s)
        b = b / (1. + b)
    return np.where(pos, a, b)


# ---- the "shaun" emission model + decode: class Viterbi ---------------------------------------------------------------

class _ShaunViterbi:
    """Common body of the ``class Viterbi`` copies.  Subclasses fix ``num_freq_bins``, ``single_side_peak_width``, the
    constructor signature, which parameter form the object keeps (``FAMILY`` 'A': linear ``transition_matrix`` /
    ``ini_probs`` and the static decode; 'B': ``log_transition_matrix_T`` / ``log_ini_probs`` and the in-place decode)
    and which of the two voicing-probability formulas the copy uses (``VOICING`` 'odds': exp(s) / (1 + exp(s)) with
    un-shifted exp of the peak logits, dcnet/softmax_viterbi.py:2344-2353; 'expit': the sign-split logistic with the
    peak logits shifted by their maximum, tonet/softmax_priors.py:1768-1778)."""

    FAMILY = 'A'
    VOICING = 'odds'
    LOGITS_ST = False            # imm/main_imm.py:186: logits arrive [n_bins, T] and are transposed

    def _setup(self, directory, transition_matrix, ini_probs, device, device_emissions):
        U = self.num_freq_bins
        A, pi = _load_parameters(directory, transition_matrix, ini_probs)
        assert A.shape == (U + 1, U + 1)
        assert np.all(np.isclose(np.sum(A, axis=1), 1))                   # transition_matrix_fn, e.g. :2413-2414
        assert pi.shape == (U + 1,)
        assert np.isclose(np.sum(pi), 1)
        self._device = device
        self._device_emissions = bool(device_emissions)
        self._pipeline = None
        if self.FAMILY == 'A':
            assert np.all(pi > 0)                                         # init_probs_fn :2381
            self.transition_matrix = A
            self.ini_probs = pi
        else:
            self.log_transition_matrix_T = _read_only_log(A, transpose=True)
            self.log_ini_probs = _read_only_log(pi)
            self._lin = (A, pi)

    @property
    def _decoder(self):
        return _lazy_decoder(self, self.log_transition_matrix_T, self.log_ini_probs)

    @staticmethod
    def expit(s):
        if s > 0:
            p = 1. / (1. + np.exp(-s))
        else:
            p = np.exp(s)
            p = p / (1. + p)
        return p

    def find_peaks_all_at_once_np_fn(self, frames_logits):
        assert frames_logits.ndim == 2
        assert frames_logits.shape[1] == self.num_freq_bins
        return find_peaks(frames_logits, self.single_side_peak_width)

    def _threshold_logit(self):
        return self.threshold

    def observation_probs_fn(self, logits):
        """logits ``[T, n_bins]`` float32 -> emission probabilities ``[n_bins + 1, T]`` float32 F-ordered, unvoiced
        last, columns summing to 1 (dcnet/softmax_viterbi.py:2316-2359; tonet/softmax_priors.py:1741-1786)."""
        assert isinstance(logits, np.ndarray)
        assert logits.dtype == np.float32
        if self.LOGITS_ST:
            logits = np.require(logits.T, np.float32, ['C'])
        n_frames, n_freq_bins = logits.shape
        assert n_freq_bins == self.num_freq_bins
        threshold = self._threshold_logit()
        p = 0.8
        offset = np.log(p / (1. - p))
        scale = 2.
        melodies_frames = np.zeros([n_freq_bins + 1, n_frames], np.float32, order='F')
        are_peaks = self.find_peaks_all_at_once_np_fn(logits)
        melodies_frames[-1, ~are_peaks.any(axis=1)] = 1                   # no peak: certainly unvoiced
        for rows, cols in _groups_by_peak_count(are_peaks):
            peak_logits = logits[rows[:, None], cols]                     # [F, n] float32 (a copy)
            g = np.max(peak_logits, axis=1)                               # the global peak's logit
            s = np.where(g >= threshold, scale * (g - threshold) + offset, scale * (g - threshold) - offset)
            if self.VOICING == 'odds':
                with np.errstate(over='ignore'):
                    p_voiced = np.exp(s)
                p_voiced = p_voiced / (1. + p_voiced)
            else:
                p_voiced = _expit(s)
                peak_logits -= g[:, None]
            np.exp(peak_logits, out=peak_logits)
            t = p_voiced / np.sum(peak_logits, axis=1)
            np.multiply(peak_logits, t[:, None], out=peak_logits)
            melodies_frames[cols, rows[:, None]] = peak_logits
            melodies_frames[-1, rows] = 1. - p_voiced
        t = np.sum(melodies_frames, axis=0)
        assert np.all(np.isclose(t, 1))
        return melodies_frames

    def _device_call(self, logits):
        from . import pipeline
        if self._pipeline is None:
            A, pi = (self.transition_matrix, self.ini_probs) if self.FAMILY == 'A' else self._lin
            self._pipeline = pipeline.MelodyPipeline(A, pi, model='shaun', single_side_peak_width=self.single_side_peak_width,
                                                     device=self._device)
            self._pipeline.threshold = float(self._threshold_logit())
        if self.LOGITS_ST:
            logits = np.require(logits.T, np.float32, ['C'])
        voiced, bins = self._pipeline(logits)
        return voiced.cpu().numpy(), bins.cpu().numpy()

    def __call__(self, logits):
        """logits ``[T, n_bins]`` -> ``(voiced bool [T], bins int64 [T])`` (dcnet/softmax_viterbi.py:2419-2431)."""
        if self._device_emissions:
            return self._device_call(logits)
        observation_probs = self.observation_probs_fn(logits)
        if self.FAMILY == 'A':
            bins = self.viterbi_librosa_fn(transition_matrix=self.transition_matrix, prob_init=self.ini_probs,
                                           probs_st=observation_probs)
        else:
            bins = self.viterbi_librosa_fn(observation_probs)
        n_bins = self.num_freq_bins
        voiced = bins < n_bins
        bins = np.minimum(bins, n_bins - 1)
        return voiced, bins


def _family_a_decode(*, transition_matrix, prob_init, probs_st):
    from . import reference_api
    return reference_api._family_a(transition_matrix, prob_init, probs_st)


def _family_b_decode(self, probs_st):
    """``probs_st [S, T]`` float32 F-contiguous, prob-domain; LOGGED IN PLACE (tonet/softmax_priors.py:1841-1878)."""
    S = self.num_freq_bins + 1
    assert probs_st.shape[0] == S
    assert probs_st.dtype == np.float32
    assert probs_st.flags['F_CONTIGUOUS'] == True  # noqa: E712
    np.add(probs_st, TINY, out=probs_st)
    np.log(probs_st, out=probs_st)
    probs = np.require(probs_st.T, np.float32, ['C'])
    paths, _ = self._decoder.decode_host(probs[None])
    return paths[0]


def _shaun_class(name, doc, *, bins, spw, family, voicing, ctor, logits_st=False):
    """Builds one directory's ``class Viterbi``.  ctor: 'none' -> ``Viterbi()`` with `fixed_threshold`; 'prob' ->
    ``Viterbi(voicing_threshold)``; 'logit' -> ``Viterbi(voicing_threshold_logit)``."""
    kind, fixed = ctor if isinstance(ctor, tuple) else (ctor, None)

    def _finish(self, directory, transition_matrix, ini_probs, device, device_emissions):
        self.num_freq_bins = bins
        self.single_side_peak_width = spw
        self._setup(directory, transition_matrix, ini_probs, device, device_emissions)

    if kind == 'none':
        def __init__(self, *, directory='.', transition_matrix=None, ini_probs=None, device=None, device_emissions=False):
            self.threshold = fixed() if callable(fixed) else fixed
            _finish(self, directory, transition_matrix, ini_probs, device, device_emissions)
    elif kind == 'prob':
        def __init__(self, voicing_threshold, *, directory='.', transition_matrix=None, ini_probs=None, device=None,
                     device_emissions=False):
            th = voicing_threshold
            assert th > 0
            assert th < 1
            self.threshold = np.log(th / (1. - th))
            _finish(self, directory, transition_matrix, ini_probs, device, device_emissions)
    else:
        def __init__(self, voicing_threshold_logit, *, directory='.', transition_matrix=None, ini_probs=None, device=None,
                     device_emissions=False):
            self.threshold_logit = voicing_threshold_logit
            _finish(self, directory, transition_matrix, ini_probs, device, device_emissions)

    body = {'__init__': __init__, '__doc__': doc, 'FAMILY': family, 'VOICING': voicing, 'LOGITS_ST': logits_st,
            '__qualname__': name}
    if kind == 'logit':
        body['_threshold_logit'] = lambda self: self.threshold_logit
    if family == 'A':
        body['viterbi_librosa_fn'] = staticmethod(_family_a_decode)
    else:
        body['viterbi_librosa_fn'] = _family_b_decode
    if kind == 'none' and fixed is not None and not callable(fixed):
        body['THRESHOLD'] = fixed
    return type('Viterbi', (_ShaunViterbi,), body)


# ---- the SoftMax emission model + decode: class SoftMaxViterbi --------------------------------------------------------

class _SoftMaxViterbi:
    """Common body of the ``class SoftMaxViterbi`` copies (Family C: ``probs_ts [T, S]`` C-ordered, logged in place)."""

    PADS_THRESHOLD = False       # dcnet: the model emits n_bins logits and the voicing-threshold logit is padded in front
    FIXED_SCALED = None          # the copies whose constructor takes no `scaled` (tonet/ablation.py: True, for_paper.py: False)

    def _is_scaled(self):
        return self.scaled if self.FIXED_SCALED is None else self.FIXED_SCALED

    def _setup(self, directory, transition_matrix, ini_probs, device, device_emissions):
        U = self.num_freq_bins
        A, pi = _load_parameters(directory, transition_matrix, ini_probs)
        assert A.shape == (U + 1, U + 1)                                  # transition_matrix_fn :2604-2609
        assert np.all(np.isclose(np.sum(A, axis=1), 1))
        assert pi.shape == (U + 1,)                                       # init_probs_fn :2588-2592
        assert np.argmax(pi) == U
        assert np.isclose(np.sum(pi), 1)
        assert np.all(pi > 0)
        self.log_transition_matrix_T = _read_only_log(A, transpose=True)
        self.ini_probs = pi
        self.log_ini_probs = _read_only_log(pi)
        self._lin = (A, pi)
        self._device = device
        self._device_emissions = bool(device_emissions)
        self._pipeline = None

    @property
    def _decoder(self):
        return _lazy_decoder(self, self.log_transition_matrix_T, self.log_ini_probs)

    def find_peaks_all_at_once_np_fn(self, logits):
        """logits ``[T, 1 + n_bins]`` -> bool ``[T, 1 + n_bins]``; column 0 (unvoiced) is always a peak
        (dcnet/softmax_viterbi.py:2508-2528)."""
        n_bins = self.num_freq_bins
        assert logits.ndim == 2
        assert logits.shape[1] == n_bins + 1
        are_peaks = np.empty([len(logits), n_bins + 1], np.bool_)
        are_peaks[:, 0] = True
        are_peaks[:, 1:] = find_peaks(logits[:, 1:], self.single_side_peak_width)
        return are_peaks

    def _with_unvoiced_column(self, logits):
        n_bins = self.num_freq_bins
        assert isinstance(logits, np.ndarray)
        assert logits.dtype == np.float32
        assert logits.ndim == 2
        if self.PADS_THRESHOLD:
            assert logits.shape[1] == n_bins
            v = self.voicing_threshold_prob_tf_var
            vth = v.numpy() if hasattr(v, 'numpy') else v
            vth_logits = np.log(vth / (1. - vth))                         # :2545-2546
            logits = np.pad(logits, [[0, 0], [1, 0]], mode='constant', constant_values=vth_logits)
        assert logits.shape[1] == n_bins + 1
        assert logits.flags['C_CONTIGUOUS']
        return logits

    def observation_probs_fn(self, logits):
        """logits ``[T, 1 + n_bins]`` (dcnet: ``[T, n_bins]``) float32 -> ``prob_ts [T, n_bins + 1]`` float32 C-ordered,
        unvoiced LAST: softmax over the peak logits, divided by the peak states' priors when ``scaled``
        (dcnet/softmax_viterbi.py:2530-2579)."""
        n_bins = self.num_freq_bins
        if self._is_scaled():
            ini_probs = self.ini_probs
            assert ini_probs.min() > 0.3 / (n_bins * 10)
            ini_probs = np.roll(ini_probs, 1)
            ini_probs = ini_probs.astype(np.float32)
        else:
            ini_probs = np.ones([n_bins + 1], dtype=np.float32)
        logits = self._with_unvoiced_column(logits)
        n_frames = len(logits)
        prob_ts = np.zeros([n_frames, 1 + n_bins], np.float32)
        are_peaks_ts = self.find_peaks_all_at_once_np_fn(logits)
        for rows, cols in _groups_by_peak_count(are_peaks_ts):
            if cols.shape[1] == 1:                                        # only the unvoiced column
                assert not cols.any()
                prob_ts[rows, 0] = 1. / ini_probs[0]
                continue
            peak_logits = logits[rows[:, None], cols]
            max_logit = np.max(peak_logits, axis=1)
            np.subtract(peak_logits, max_logit[:, None], out=peak_logits)
            np.exp(peak_logits, out=peak_logits)
            t = np.sum(peak_logits, axis=1)
            np.divide(peak_logits, t[:, None], out=peak_logits)
            np.divide(peak_logits, ini_probs[cols], out=peak_logits)
            prob_ts[rows[:, None], cols] = peak_logits
        return np.roll(prob_ts, shift=-1, axis=1)

    def viterbi_librosa_fn(self, probs_ts):
        """``probs_ts [T, S]`` float32 C-contiguous, prob-domain (may exceed 1 when scaled); LOGGED IN PLACE
        (dcnet/softmax_viterbi.py:2636-2674)."""
        S = self.num_freq_bins + 1
        assert probs_ts.ndim == 2
        assert probs_ts.shape[1] == S
        assert probs_ts.dtype == np.float32
        assert probs_ts.flags['C_CONTIGUOUS']
        np.add(probs_ts, TINY, out=probs_ts)
        np.log(probs_ts, out=probs_ts)
        paths, _ = self._decoder.decode_host(probs_ts[None])
        return paths[0]

    def _device_call(self, logits):
        from . import pipeline
        if self._pipeline is None:
            A, pi = self._lin
            self._pipeline = pipeline.MelodyPipeline(A, pi, model='softmax', scaled=bool(self._is_scaled()),
                                                     single_side_peak_width=self.single_side_peak_width, device=self._device)
        voiced, bins = self._pipeline(self._with_unvoiced_column(logits))
        return voiced.cpu().numpy(), bins.cpu().numpy()

    def __call__(self, logits):
        """logits -> ``(voiced bool [T], bins int64 [T])`` (dcnet/softmax_viterbi.py:2620-2634)."""
        if self._device_emissions:
            return self._device_call(logits)
        n_bins = self.num_freq_bins
        prob_ts = self.observation_probs_fn(logits)
        bins = self.viterbi_librosa_fn(prob_ts)
        voiced = bins < n_bins
        bins = np.minimum(bins, n_bins - 1)
        return voiced, bins


def _softmax_class(name, doc, *, bins, spw, ctor):
    """ctor: 'scaled' -> ``SoftMaxViterbi(scaled)``; 'dcnet' -> ``SoftMaxViterbi(voicing_threspold_prob, scaled)``;
    ('fixed', flag) -> ``SoftMaxViterbi()`` with the division by the priors always on (tonet/ablation.py) or always off
    (tonet/for_paper.py)."""
    fixed = None
    if isinstance(ctor, tuple):
        ctor, fixed = ctor

    def _finish(self, scaled, directory, transition_matrix, ini_probs, device, device_emissions):
        if fixed is None:
            self.scaled = scaled
        self.num_freq_bins = bins
        self.single_side_peak_width = spw
        self._setup(directory, transition_matrix, ini_probs, device, device_emissions)

    if ctor == 'scaled':
        def __init__(self, scaled, *, directory='.', transition_matrix=None, ini_probs=None, device=None,
                     device_emissions=False):
            _finish(self, scaled, directory, transition_matrix, ini_probs, device, device_emissions)
    elif ctor == 'fixed':
        def __init__(self, *, directory='.', transition_matrix=None, ini_probs=None, device=None, device_emissions=False):
            _finish(self, fixed, directory, transition_matrix, ini_probs, device, device_emissions)
    else:
        def __init__(self, voicing_threspold_prob, scaled, *, directory='.', transition_matrix=None, ini_probs=None,
                     device=None, device_emissions=False):
            # the reference asserts a tf.Variable (dcnet/softmax_viterbi.py:2497); anything with .numpy(), or a float
            self.voicing_threshold_prob_tf_var = voicing_threspold_prob
            _finish(self, scaled, directory, transition_matrix, ini_probs, device, device_emissions)
    return type('SoftMaxViterbi', (_SoftMaxViterbi,), {'__init__': __init__, '__doc__': doc, '__qualname__': name,
                                                       'PADS_THRESHOLD': ctor == 'dcnet', 'FIXED_SCALED': fixed})


# ---- imm/tf_imm.py: HF0 in, dense recipe matrix, uniform pi (Family D) ------------------------------------------------

class HF0Viterbi:
    """``imm/tf_imm.py:48-135 class Viterbi``: fully dense transition matrix from the IMM recipe
    (imm/transition_matrix.py:4-31), uniform initial distribution, emissions = log of the source-filter model's HF0."""

    def __init__(self, bins_per_semitone, n_bins, *, device=None):
        self.b = bins_per_semitone
        self.n_bins = n_bins
        transition_matrix = hmm_params.dense_imm_transition_matrix(bins_per_semitone, n_bins)   # :54
        assert np.all(transition_matrix > 0)
        transition_matrix = np.log(transition_matrix.T)                   # float64 logs, then the cast (:56-59)
        assert not np.any(np.isneginf(transition_matrix))
        self.log_transition_matrix_T = np.require(transition_matrix, np.float32, ['C'])
        init_probs = np.empty([n_bins + 1])                               # :61-66
        init_probs.fill(1. / (n_bins + 1))
        self.log_prob_init = np.log(init_probs).astype(np.float32)
        self._device = device

    @property
    def _decoder(self):
        return _lazy_decoder(self, self.log_transition_matrix_T, self.log_prob_init)

    def process_HF0_fn(self, HF0):
        """``HF0 [n_bins, T]`` (>= 0) -> ``log(HF0 + smallest positive entry)`` with an unvoiced row appended at the
        table's global minimum (imm/tf_imm.py:70-88).  Accepts anything with ``.numpy()`` where the reference accepts a
        tf.Tensor."""
        if not isinstance(HF0, np.ndarray):
            assert hasattr(HF0, 'numpy')
            HF0 = HF0.numpy()
        U = self.n_bins
        assert HF0.shape[0] == U
        t = HF0[HF0 > 0]
        t = t.min()
        if np.log(t) < -87:
            t = np.exp(-87)
        HF0 = HF0 + t
        HF0 = np.log(HF0)
        _min = np.min(HF0)
        HF0 = np.pad(HF0, [[0, 1], [0, 0]], mode='constant', constant_values=_min)
        return HF0

    def viterbi_librosa_fn(self, log_HF0):
        """``log_HF0 [S, T]`` float32 log-domain -> ``np.int64[T]`` (imm/tf_imm.py:90-127)."""
        S = self.n_bins + 1
        assert isinstance(log_HF0, np.ndarray)
        assert log_HF0.dtype == np.float32
        assert log_HF0.shape[0] == S
        if log_HF0.flags['C_CONTIGUOUS']:
            paths, _ = self._decoder.decode_host_st(log_HF0)              # (:101's transposing copy runs on the GPU)
            return paths
        probs = np.require(np.transpose(log_HF0), np.float32, ['C'])
        paths, _ = self._decoder.decode_host(probs[None])
        return paths[0]

    def __call__(self, HF0):
        """``HF0`` -> states ``np.int64[T]`` (imm/tf_imm.py:129-135)."""
        log_HF0 = self.process_HF0_fn(HF0)
        return self.viterbi_librosa_fn(log_HF0)


# ---- one namespace per experiment directory ---------------------------------------------------------------------------

def _logit(th):
    return np.log(th / (1. - th))


dcnet = types.SimpleNamespace(
    Viterbi=_shaun_class('dcnet.Viterbi', 'dcnet/softmax_viterbi.py:2273-2485 (main.py:2256, lontano.py:2256, '
                         'viterbi_states_10ms_step.py:1937): ``Viterbi()``, 320 bins, voicing threshold 0.31.',
                         bins=320, spw=5, family='A', voicing='odds', ctor=('none', lambda: _logit(0.31))),
    SoftMaxViterbi=_softmax_class('dcnet.SoftMaxViterbi', 'dcnet/softmax_viterbi.py:2488-2674: '
                                  '``SoftMaxViterbi(voicing_threspold_prob, scaled)``; logits [T, 320], the threshold '
                                  'logit is padded as the unvoiced column.', bins=320, spw=5, ctor='dcnet'))
msnet = types.SimpleNamespace(
    Viterbi=_shaun_class('msnet.Viterbi', 'msnet/viterbi_softmax.py:1515-1729 (hsieh_m2m3.py:1501): '
                         '``Viterbi(voicing_threshold)``, 320 bins.', bins=320, spw=5, family='A', voicing='odds', ctor='prob'),
    SoftMaxViterbi=_softmax_class('msnet.SoftMaxViterbi', 'msnet/viterbi_softmax.py:1732-1910: ``SoftMaxViterbi(scaled)``, '
                                  'logits [T, 321].', bins=320, spw=5, ctor='scaled'))
ftanet = types.SimpleNamespace(
    Viterbi=_shaun_class('ftanet.Viterbi', 'ftanet/viterbi_performance.py:2569-2735: ``Viterbi(voicing_threshold_logit)``, '
                         '320 bins.', bins=320, spw=5, family='A', voicing='odds', ctor='logit'),
    SoftMaxViterbi=_softmax_class('ftanet.SoftMaxViterbi', 'ftanet/viterbi_performance.py:2736-2916: '
                                  '``SoftMaxViterbi(scaled)``, logits [T, 321].', bins=320, spw=5, ctor='scaled'))
jdc = types.SimpleNamespace(
    Viterbi=_shaun_class('jdc.Viterbi', 'jdc/viterbi_softmax.py:1902-2091 (determine_threshold_kum_m2m3.py:1882): '
                         '``Viterbi(voicing_threshold)``, 721 bins, peak half-width 16.', bins=721, spw=16, family='A',
                         voicing='odds', ctor='prob'),
    SoftMaxViterbi=_softmax_class('jdc.SoftMaxViterbi', 'jdc/viterbi_softmax.py:2094-2272: ``SoftMaxViterbi(scaled)``, '
                                  'logits [T, 722], peak half-width 16.', bins=721, spw=16, ctor='scaled'))
tonet = types.SimpleNamespace(
    Viterbi=_shaun_class('tonet.Viterbi', 'tonet/softmax_priors.py:1691-1878 (for_paper.py, ablation.py; Family B): '
                         '``Viterbi(voicing_threshold)``, 360 bins.', bins=360, spw=5, family='B', voicing='expit', ctor='prob'),
    SoftMaxViterbi=_softmax_class('tonet.SoftMaxViterbi', 'tonet/softmax_priors.py:1881-2061: ``SoftMaxViterbi(scaled)``, '
                                  'logits [T, 361].', bins=360, spw=5, ctor='scaled'),
    SoftMaxViterbiAlwaysScaled=_softmax_class('tonet.SoftMaxViterbiAlwaysScaled', 'tonet/ablation.py:1874-2051: '
                                              '``SoftMaxViterbi()``, always divided by the priors.', bins=360, spw=5,
                                              ctor=('fixed', True)),
    SoftMaxViterbiWide=_softmax_class('tonet.SoftMaxViterbiWide', 'tonet/for_paper.py:1873-2039: ``SoftMaxViterbi()``, peak '
                                      'half-width 15, plain softmax over the peaks (never divided by the priors).', bins=360,
                                      spw=15, ctor=('fixed', False)))
# tonet/main_shaun.py:1664 and hyper_parameter_selection.py:1673 keep linear parameters and the static decode (Family A)
tonet.ViterbiA = _shaun_class('tonet.ViterbiA', 'tonet/main_shaun.py:1664-1858, hyper_parameter_selection.py:1673-1867: '
                              '``Viterbi(voicing_threshold)``, 360 bins, linear parameters + static decode.', bins=360,
                              spw=5, family='A', voicing='expit', ctor='prob')
imm = types.SimpleNamespace(
    Viterbi=_shaun_class('imm.Viterbi', 'imm/main_imm.py:141-325 (thresholding.py:642, original_adc04_performance.py:129; '
                         'Family B): ``Viterbi()``, 721 bins, peak half-width 20, logits given [n_bins, T].', bins=721,
                         spw=20, family='B', voicing='expit', ctor=('none', 2.442347), logits_st=True),
    HF0Viterbi=HF0Viterbi)
