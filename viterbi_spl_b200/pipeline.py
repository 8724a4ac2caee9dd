"""Whole post-processing chain on the GPU: acoustic-model logits -> emissions -> Viterbi -> (voiced, bins), batched over
recordings, nothing round-tripping the host.

It is what the reference's ``SoftMaxViterbi.__call__`` / ``Viterbi.__call__`` do per recording on the CPU
(dcnet/softmax_viterbi.py:2620-2634, tonet/softmax_priors.py:1825-1839):

    prob = self.observation_probs_fn(logits)       # python loop over frames          -> vit_emissions_f32
    bins = self.viterbi_librosa_fn(prob)           # log(prob + tiny), recursion      -> vit_decode_f32
    voiced = bins < n_bins; bins = np.minimum(bins, n_bins - 1)                       -> vit_voiced_bins

The emission values come from the GPU's expf/logf (1e-5 relative to NumPy's, not bit-identical), so unlike the
``reference_api`` wrappers -- which log on the host and are bit-exact -- a decoded path can differ from the reference
where two paths are within rounding of each other.  Use ``reference_api`` for bit-exact parity, this for throughput.
"""
import ctypes

import numpy as np
import torch

from . import _lib, hmm_params
from .decoder import ViterbiDecoder, checked_lengths

SOFTMAX, SHAUN = 0, 1


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def emissions_device(logits, n_bins, model, prior=None, single_side_peak_width=5, threshold=0.0, out_log=True, out=None):
    """logits: CUDA float32 [B, T, 1 + n_bins] (model SOFTMAX, column 0 = unvoiced) or [B, T, n_bins] (model SHAUN).
    prior: CUDA float32 [1 + n_bins] = roll(ini_probs, 1) for the "scaled" SoftMax model, else None.
    Returns the emission table [B, T, n_bins + 1] (unvoiced last), log(p + tiny) if out_log else p."""
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous() and logits.ndim == 3
    B, T, n_in = logits.shape
    assert n_in == (n_bins + 1 if model == SOFTMAX else n_bins)
    if prior is not None:
        assert model == SOFTMAX and prior.is_cuda and prior.dtype == torch.float32 and prior.shape == (n_bins + 1,)
    if out is None:
        out = torch.empty((B, T, n_bins + 1), dtype=torch.float32, device=logits.device)
    with torch.cuda.device(logits.device):
        st = torch.cuda.current_stream()
        _lib.check(_lib.load().vit_emissions_f32(_ptr(logits), _ptr(prior), B, T, n_bins, model, single_side_peak_width,
                                                 float(threshold), 1 if out_log else 0, _ptr(out),
                                                 ctypes.c_void_p(st.cuda_stream)))
    return out


def voiced_bins_device(states, n_bins):
    """states: CUDA int64 [...]; returns (voiced bool [...], bins int64 [...])."""
    assert states.is_cuda and states.dtype == torch.int64 and states.is_contiguous()
    voiced = torch.empty(states.shape, dtype=torch.uint8, device=states.device)
    bins = torch.empty_like(states)
    with torch.cuda.device(states.device):
        st = torch.cuda.current_stream()
        _lib.check(_lib.load().vit_voiced_bins(_ptr(states), states.numel(), n_bins, _ptr(voiced), _ptr(bins),
                                               ctypes.c_void_p(st.cuda_stream)))
    return voiced.bool(), bins


COUNTER_NAMES = ('gt_voiced', 'gt_unvoiced', 'correct_voiced', 'incorrect_voiced', 'correct_unvoiced',
                 'correct_pitches_wide', 'correct_pitches_strict', 'correct_chromas_wide', 'correct_chromas_strict')


def melody_stats_device(logits, ref_notes, bins, voiced, lengths=None, n_bins=None, logit_offset=0, note_min=23.6,
                        note_step=0.2):
    """The statistics step after the decode (MetricsInference.viterbi_update_states_tf_fn,
    dcnet/softmax_viterbi.py:2923-2979), batched: logits CUDA float32 [B, T, W] (bin k = column logit_offset + k),
    ref_notes float32 [B, T], bins int64 [B, T] and voiced bool/uint8 [B, T] as MelodyPipeline returns them ->
    (est_notes_with_voicing_info float32 [B, T], counters int64 [B, 9] in COUNTER_NAMES order)."""
    assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous() and logits.ndim == 3
    B, T, W = logits.shape
    n_bins = W - logit_offset if n_bins is None else int(n_bins)
    ref_notes = ref_notes.to(logits.device, torch.float32).contiguous()
    bins = bins.to(logits.device, torch.int64).contiguous()
    voiced = voiced.to(logits.device).to(torch.uint8).contiguous()
    assert ref_notes.shape == (B, T) and bins.shape == (B, T) and voiced.shape == (B, T)
    if lengths is not None:
        lengths = torch.as_tensor(lengths).to(logits.device, torch.int32).contiguous()
    est = torch.empty((B, T), dtype=torch.float32, device=logits.device)
    counters = torch.empty((B, len(COUNTER_NAMES)), dtype=torch.int64, device=logits.device)
    with torch.cuda.device(logits.device):
        st = torch.cuda.current_stream()
        _lib.check(_lib.load().vit_melody_stats_f32(_ptr(logits), W, int(logit_offset), _ptr(ref_notes), _ptr(bins),
                                                    _ptr(voiced), _ptr(lengths), B, T, n_bins, float(note_min),
                                                    float(note_step), _ptr(est), _ptr(counters),
                                                    ctypes.c_void_p(st.cuda_stream)))
    return est, counters


class MelodyPipeline:
    """logits [B, T, *] -> (voiced [B, T] bool, bins [B, T] int64) entirely on the GPU.

    model='softmax': logits [B, T, 1 + n_bins] with the unvoiced logit in column 0 (what the tonet/jdc/msnet/ftanet
    acoustic models emit, tonet/softmax_priors.py:2273-2274; dcnet pads the constant logit(voicing threshold),
    dcnet/softmax_viterbi.py:2546-2548); scaled=True divides by the prior of each peak state (:2571-2572).
    model='shaun': logits [B, T, n_bins] and a voicing threshold (probability)."""

    def __init__(self, transition_matrix, ini_probs, model='softmax', scaled=False, voicing_threshold=0.5,
                 single_side_peak_width=5, device=None, algo='auto'):
        self.n_bins = len(ini_probs) - 1
        self.model = {'softmax': SOFTMAX, 'shaun': SHAUN}[model]
        self.spw = int(single_side_peak_width)
        self.threshold = float(np.log(voicing_threshold / (1. - voicing_threshold)))         # tonet :1703-1706
        logA_T, log_pi = hmm_params.log_params(np.asarray(transition_matrix), np.asarray(ini_probs))
        self.decoder = ViterbiDecoder(logA_T, log_pi, device=device, algo=algo)
        self.device = self.decoder.device
        self.prior = None
        if self.model == SOFTMAX and scaled:
            self.prior = torch.as_tensor(np.roll(np.asarray(ini_probs, np.float32), 1).copy()).to(self.device)
        self._A = np.asarray(transition_matrix, np.float32)
        self._pi = np.asarray(ini_probs, np.float32)
        self._fb = None

    def emissions(self, logits, out_log=True):
        return emissions_device(logits, self.n_bins, self.model, self.prior, self.spw, self.threshold, out_log)

    def __call__(self, logits, lengths=None):
        logits = torch.as_tensor(logits)
        squeeze = logits.ndim == 2
        if squeeze:
            logits = logits[None]
        logits = logits.to(self.device, torch.float32).contiguous()
        E = self.emissions(logits, out_log=True)
        states, _ = self.decoder.decode_device(E, lengths)
        voiced, bins = voiced_bins_device(states, self.n_bins)
        return (voiced[0], bins[0]) if squeeze else (voiced, bins)

    def posteriors(self, logits, lengths=None):
        """logits [B, T, *] -> (gamma [B, T, n_bins + 1] posterior state marginals, unvoiced last; log L [B]): the emission
        LIKELIHOODS of observation_probs_fn (probability domain, what dcnet/softmax_viterbi.py:2530-2579 returns) go straight
        into the scaled forward-backward pass on the same transition matrix and initial distribution -- the sum-product
        counterpart of __call__ (the reference only has the max-product decode; semantics: oracle/fb_oracle.py)."""
        from .posterior import ForwardBackward
        logits = torch.as_tensor(logits)
        squeeze = logits.ndim == 2
        if squeeze:
            logits = logits[None]
        logits = logits.to(self.device, torch.float32).contiguous()
        if self._fb is None:
            self._fb = ForwardBackward(self._A, self._pi, device=self.device)
        if lengths is not None and not (torch.is_tensor(lengths) and lengths.is_cuda):
            lengths = torch.as_tensor(checked_lengths(np.asarray(lengths), logits.shape[0], logits.shape[1])).to(self.device)
        lik = self.emissions(logits, out_log=False)
        gamma, ll = self._fb.run_device(lik, lengths)
        return (gamma[0], ll[0]) if squeeze else (gamma, ll)

    def evaluate(self, logits, ref_notes, lengths=None, note_min=23.6, note_step=0.2):
        """logits [B, T, *] + reference notes [B, T] -> (voiced, bins, est_notes_with_voicing_info, counters [B, 9]):
        emissions, decode, voiced/bins and the frame statistics of dcnet/softmax_viterbi.py:3033-3046 in one go, all
        on the GPU."""
        logits = torch.as_tensor(logits).to(self.device, torch.float32).contiguous()
        assert logits.ndim == 3
        if lengths is not None:
            if not (torch.is_tensor(lengths) and lengths.is_cuda):
                lengths = torch.as_tensor(checked_lengths(torch.as_tensor(lengths).numpy() if torch.is_tensor(lengths) else lengths,
                                                          logits.shape[0], logits.shape[1]))
            lengths = lengths.to(self.device, torch.int32).contiguous()
        voiced, bins = self(logits, lengths)
        off = 1 if self.model == SOFTMAX else 0
        est, counters = melody_stats_device(logits, torch.as_tensor(ref_notes), bins, voiced, lengths, self.n_bins, off,
                                            note_min, note_step)
        return voiced, bins, est, counters
