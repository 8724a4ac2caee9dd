"""TEST INFRASTRUCTURE ONLY -- NumPy restatement of the reference's post-decode statistics step.

The reference's version is TensorFlow (`MetricsInference.viterbi_update_states_tf_fn`,
dcnet/softmax_viterbi.py:2923-2979, calling `MetricsBase.est_notes_fn`, dcnet/main.py:1911-1934) and TensorFlow is not
in this image; this file follows those lines one by one in float32.  PARITY PINNED (round 2): the reference's own
function text is executed on a NumPy-backed stand-in for the TensorFlow ops it calls (oracle/ref_loader.py
`dcnet_melody_stats`), its outputs are committed as tests/golden/melody_stats.npz, and tests/test_oracle.py checks this
restatement against them and against the live reference function (bit-equal here).  What stays outside the pin is
TensorFlow's own float32 `sigmoid` kernel (1 / (1 + exp(-x)) here), hence the 1e-5 tolerance on notes.  Only tests/ may
import this file.
"""
import numpy as np

NOTE_MIN = np.float32(23.6)                                                    # dcnet/softmax_viterbi.py:428
NOTE_RANGE = (np.arange(320) / 5. + 23.6).astype(np.float32)                    # :429-431
COUNTERS = ('gt_voiced', 'gt_unvoiced', 'correct_voiced', 'incorrect_voiced', 'correct_unvoiced',
            'correct_pitches_wide', 'correct_pitches_strict', 'correct_chromas_wide', 'correct_chromas_strict')


def sigmoid32(x):
    x = np.asarray(x, np.float32)
    return (np.float32(1) / (np.float32(1) + np.exp(-x))).astype(np.float32)


def est_notes_np(est_peak_indices, est_probs, note_range=NOTE_RANGE):
    """dcnet/main.py:1911-1934, float32.  est_peak_indices [T] int, est_probs [T, n_bins]."""
    note_offset = note_range[0]                                                # :1915
    rel = (note_range - note_offset).astype(np.float32)                         # :1917
    n_bins = est_probs.shape[1]
    frames = np.arange(n_bins, dtype=np.int64)                                  # :1920
    peak_masks = np.abs(np.asarray(est_peak_indices, np.int64)[:, None] - frames[None, :]) <= 1      # :1921-1923
    masked = np.where(peak_masks, est_probs, np.float32(0)).astype(np.float32)  # :1924
    norm = masked.sum(axis=1, dtype=np.float32)                                 # :1926
    est = (rel[None, :] * masked).sum(axis=1, dtype=np.float32)                 # :1929-1930
    est = est / np.maximum(norm, np.float32(1e-3))                              # :1932
    return (est + note_offset).astype(np.float32)                               # :1933


def melody_stats_np(ref_notes, logits, melody_bins, est_voicing, note_range=NOTE_RANGE):
    """dcnet/softmax_viterbi.py:2938-2977 for one recording -> (est_notes_with_voicing_info [T] f32, counters dict)."""
    ref_notes = np.asarray(ref_notes, np.float32)
    est_voicing = np.asarray(est_voicing, bool)
    ref_voicing = ref_notes > np.float32(.1)                                    # :2938
    est_notes = est_notes_np(melody_bins, sigmoid32(logits), note_range)        # :2943
    diff = np.abs(est_notes - ref_notes).astype(np.float32)                     # :2945
    c = {}
    c['gt_voiced'] = int(np.count_nonzero(ref_voicing))                         # :2948
    c['gt_unvoiced'] = int(ref_voicing.size - c['gt_voiced'])                   # :2949
    c['correct_voiced'] = int(np.count_nonzero(ref_voicing & est_voicing))      # :2950-2951
    c['incorrect_voiced'] = int(np.count_nonzero(~ref_voicing & est_voicing))   # :2952-2953
    c['correct_unvoiced'] = int(np.count_nonzero(~ref_voicing & ~est_voicing))  # :2954-2955
    wide = ref_voicing & (diff < np.float32(.5))                                # :2962-2963
    c['correct_pitches_wide'] = int(np.count_nonzero(wide))
    c['correct_pitches_strict'] = int(np.count_nonzero(wide & est_voicing))     # :2964
    octave = (np.floor(diff / np.float32(12.) + np.float32(.5)) * np.float32(12.)).astype(np.float32)   # dcnet/main.py:1937-1941
    cw = ref_voicing & (np.abs(diff - octave) < np.float32(.5))                 # :2970-2972
    c['correct_chromas_wide'] = int(np.count_nonzero(cw))
    c['correct_chromas_strict'] = int(np.count_nonzero(cw & est_voicing))       # :2973
    out = np.where(est_voicing, est_notes, -est_notes).astype(np.float32)       # :2979
    return out, c
