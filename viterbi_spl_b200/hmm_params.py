"""HMM parameter builders and ``.dat`` I/O for the pitch-bin HMMs.

These are the producers of the decoder's inputs: the dense ``[S, S]`` transition matrix (voiced bins + one
unvoiced state as the LAST index) and the ``[S]`` initial distribution.  Each function restates a reference
script as a pure function of its input arrays (the reference versions are top-level scripts that read
``transition_int.dat`` / ``p_steady.dat`` from the working directory):

* ``banded_transition_matrix``  -- dcnet/viterbi_transition_matrix.py:62-101,
                                   tonet/viterbi_transition_post_processing.py:44-88 (jdc/ftanet/imm copies)
* ``floored_init_probs``        -- dcnet/viterbi_init_probs.py:9-24, tonet/p_steady_post_processing.py:7-24
* ``dense_imm_transition_matrix`` -- imm/transition_matrix.py:4-31
* ``log_params``                -- the ctor-time ``log(x + tiny)`` + transpose of Family B/C decoders,
                                   tonet/softmax_priors.py:1788-1823
* ``load_dat`` / ``save_dat``   -- self_defined/load_np_array_from_file.py:4-27, save_np_array_to_file.py:4-39

Host-side NumPy only: this is offline parameter preparation, not the hot path.
"""
import os

import numpy as np

TINY = np.finfo(np.float32).tiny

# voiced/unvoiced switch matrices hard-coded by the reference ([[V->V, V->U], [U->V, U->U]])
SWITCH_DCNET = np.asarray([[0.98713454, 0.01286546], [0.01002112, 0.98997888]], np.float32)  # dcnet/viterbi_transition_matrix.py:78-79
SWITCH_TONET = np.asarray([[0.97790518, 0.02209482], [0.01720512, 0.98279488]], np.float32)  # tonet/viterbi_transition_post_processing.py:65-66


def single_side_d_max(h, bins_per_octave):
    """Largest pitch jump (bins per hop, one side) kept in the band: tonet/viterbi_transition_post_processing.py:11-18.
    h is the hop in seconds.  tonet/ftanet (h=0.01, B=60) -> 14; imm (B=240) -> 56."""
    return int(35.92 * h * bins_per_octave * 1.3 // 2)


def jump_histogram(transition_counts, n_bins, d_max):
    """Histogram of counted pitch jumps d = j - i clipped to [-d_max, d_max] (dcnet/viterbi_transition_matrix.py:62-74).
    transition_counts is the integer matrix `transition_int` (only its voiced [n_bins, n_bins] block is used)."""
    c = np.asarray(transition_counts)[:n_bins, :n_bins]
    i, j = np.nonzero(c)
    d = np.clip(j - i, -d_max, d_max) + d_max
    hist = np.zeros([2 * d_max + 1], np.int64)
    np.add.at(hist, d, c[i, j].astype(np.int64))
    return hist


def banded_transition_matrix(jump_hist, n_bins, count_floor, switch):
    """Dense row-stochastic A [n_bins+1, n_bins+1] float32: Toeplitz band of the floored, normalised jump
    histogram, rows renormalised (edges lose mass), then embedded in the 2x2 voiced/unvoiced switch.

    dcnet/viterbi_transition_matrix.py:75-98 (count_floor=6, d_max=12, SWITCH_DCNET);
    tonet/viterbi_transition_post_processing.py:56-86 (count_floor=2, d_max=14, SWITCH_TONET); jdc uses d_max=40,
    count_floor=6; imm-HMM d_max=56.  Entries outside the band are exactly 0 (-> log(tiny) in the decoder).
    """
    jump_hist = np.asarray(jump_hist, np.int64)
    d_max = (len(jump_hist) - 1) // 2
    assert len(jump_hist) == 2 * d_max + 1
    switch = np.asarray(switch, np.float32)
    d_trans = np.maximum(jump_hist, count_floor)
    d_trans = d_trans / np.sum(d_trans)                                   # float64 probabilities per jump
    idx = np.arange(n_bins)
    d = idx[None, :] - idx[:, None]                                       # j - i
    band = np.abs(d) <= d_max
    tm = np.zeros([n_bins, n_bins], np.float32)
    tm[band] = d_trans[(d + d_max)[band]]                                 # float64 -> float32 on assignment
    tm = tm / np.sum(tm, axis=1)[:, None]                                 # float32 row renormalisation
    assert np.all(np.isclose(np.sum(tm, axis=1), 1))
    tm = np.pad(tm, [(0, 1), (0, 1)])
    tm[:n_bins, :n_bins] *= switch[0, 0]
    tm[:n_bins, n_bins] = switch[0, 1]
    tm[n_bins, :n_bins] = switch[1, 0] / n_bins
    tm[n_bins, n_bins] = switch[1, 1]
    assert np.all(np.isclose(np.sum(tm, axis=1), 1))
    return tm


def floored_init_probs(p_steady, p_floor=None):
    """Initial distribution [S] float32 from steady-state occupancy (unvoiced last): voiced part floored at
    `p_floor`, renormalised to the voiced mass.  p_floor=3e-4 is dcnet/viterbi_init_probs.py:9;
    None -> 1/len(p_steady)/10 as in tonet/p_steady_post_processing.py:11."""
    p_steady = np.asarray(p_steady)
    if p_floor is None:
        p_floor = 1. / len(p_steady) / 10.
    p_unvoiced = p_steady[-1]
    p_voiced = 1. - p_unvoiced
    ps = np.maximum(p_steady[:-1], p_floor)
    ps = ps / np.sum(ps)
    ps = ps * p_voiced
    out = np.append(ps, p_unvoiced).astype(np.float32)
    assert np.isclose(np.sum(out), 1.)
    return out


def dense_imm_transition_matrix(bins_per_semitone, n_bins):
    """Fully dense (no zeros) A [n_bins+1, n_bins+1] float64 of the IMM decoder: exp(-floor(|i-j|/b)) with a cutoff at
    10 semitones; voiced->unvoiced 1e-90, unvoiced->voiced 1e-80, unvoiced->unvoiced 1e-100 (times the cutoff
    value) before row normalisation.  imm/transition_matrix.py:4-31."""
    p = np.exp(-(np.arange(n_bins) // bins_per_semitone).astype(np.float64))
    cutoff = 10 * bins_per_semitone
    p[cutoff:] = p[cutoff - 1]
    idx = np.arange(n_bins)
    A = np.empty([n_bins + 1, n_bins + 1], np.float64)
    A[:n_bins, :n_bins] = p[np.abs(idx[:, None] - idx[None, :])]
    cp = p[cutoff - 1]
    A[:n_bins, n_bins] = cp * 10 ** (-90)
    A[n_bins, :n_bins] = cp * 10 ** (-80)
    A[n_bins, n_bins] = cp * 10 ** (-100)
    A = A / np.sum(A, axis=1)[:, None]
    assert np.allclose(np.sum(A, axis=1), 1.)
    return A


def log_params(transition_matrix, init_probs, add_tiny=True):
    """(logA_T [S,S] f32 C-contiguous dst-major, log_pi [S] f32), exactly as the Family B/C constructors do:
    ``log(x + tiny)`` in the input dtype, transpose, ``np.require(float32, 'C')`` (tonet/softmax_priors.py:1788-1823).
    add_tiny=False is the imm ctor (imm/tf_imm.py:56-68: float64 log of an all-positive matrix, then cast)."""
    A = np.asarray(transition_matrix)
    pi = np.asarray(init_probs)
    if add_tiny:
        t = np.log(A + TINY)
        p = np.log(pi + TINY)
    else:
        assert np.all(A > 0)
        t = np.log(A)
        p = np.log(pi)
    assert not np.any(np.isneginf(t)) and not np.any(np.isneginf(p))
    logA_T = np.require(t.T, np.float32, ['C'])
    log_pi = np.require(p, np.float32)
    return logA_T, log_pi


# ---- .dat container: one ASCII header line `name [C|F] dtype d0 d1 ...\n` followed by raw bytes ----------------

def load_dat(file_name):
    """Reader for the reference's array container (self_defined/load_np_array_from_file.py:4-27), including the header
    variant without the C/F flag used by the shipped msnet/*.dat.  Returns (record_name, array)."""
    with open(file_name, 'rb') as fh:
        header = fh.readline().decode('utf-8').split()
        name = header[0]
        if header[1] in ('C', 'F'):
            order, dtype, dims = header[1], header[2], [int(x) for x in header[3:]]
        else:
            order, dtype, dims = 'C', header[1], [int(x) for x in header[2:]]
        out = np.frombuffer(fh.read(), dtype=dtype).reshape(*dims)   # payload is always stored C-ordered
        if len(dims) > 1 and order == 'F':
            out = np.require(out, requirements=['F'])
    return name, out


def save_dat(file_name, array, rec_name):
    """Writer matching self_defined/save_np_array_to_file.py:4-39 (header carries the C/F flag; payload C-ordered)."""
    assert isinstance(rec_name, str) and len(rec_name) and ' ' not in rec_name
    assert isinstance(array, np.ndarray) and array.ndim >= 1
    c_flag, f_flag = array.flags['C_CONTIGUOUS'], array.flags['F_CONTIGUOUS']
    if array.ndim == 1:
        order = 'C'
    else:
        assert (c_flag or f_flag) and not (c_flag and f_flag)
        order = 'C' if c_flag else 'F'
    header = ' '.join([rec_name, order, str(array.dtype)] + ['{:d}'.format(d) for d in array.shape]) + '\n'
    with open(file_name, 'wb') as fh:
        fh.write(header.encode('utf-8'))
        fh.write(np.ascontiguousarray(array).tobytes())
        fh.flush()
        os.fsync(fh.fileno())


# ---- synthetic statistics for the state sets whose .dat files the reference does not ship ---------------------

def synthetic_jump_counts(d_max, seed=0, scale=2.0e5):
    """Laplacian-ish pitch-jump histogram (counts) standing in for `transition_int` statistics (SURVEY.md 8(d) cfg 1)."""
    rng = np.random.default_rng(seed)
    d = np.arange(-d_max, d_max + 1)
    h = scale * np.exp(-np.abs(d) / 1.3) * (1.0 + 0.1 * rng.random(len(d)))
    return np.floor(h).astype(np.int64)


def synthetic_p_steady(n_bins, p_unvoiced=0.56, seed=0):
    rng = np.random.default_rng(seed)
    centre = rng.uniform(0.35, 0.65) * n_bins
    p = np.exp(-0.5 * ((np.arange(n_bins) - centre) / (0.12 * n_bins)) ** 2) * (0.5 + rng.random(n_bins))
    p = p / p.sum() * (1.0 - p_unvoiced)
    return np.append(p, p_unvoiced)


def synthetic_hmm(state_set, seed=0):
    """(A [S,S] f32/f64 row-stochastic, pi [S]) for a named state set:
    'dcnet' S=321 (d_max 12, floor 6), 'tonet' S=361 (d_max 14, floor 2), 'jdc' S=722 (d_max 40, floor 6),
    'imm' S=722 fully dense (imm/transition_matrix.py recipe with b=20, uniform pi as imm/tf_imm.py:62-67),
    'imm_hmm' S=722 banded (imm/viterbi_transition_post_processing.py:7-18: 240 bins per octave -> d_max 56, floor 2)."""
    if state_set == 'imm':
        n = 721
        return dense_imm_transition_matrix(20, n), np.full([n + 1], 1. / (n + 1))
    n, d_max, floor, switch, p_floor = {
        'dcnet': (320, 12, 6, SWITCH_DCNET, 3e-4),
        'tonet': (360, single_side_d_max(0.01, 60), 2, SWITCH_TONET, None),
        'jdc': (721, 40, 6, SWITCH_TONET, None),
        'imm_hmm': (721, single_side_d_max(0.01, 240), 2, SWITCH_DCNET, None),
    }[state_set]
    A = banded_transition_matrix(synthetic_jump_counts(d_max, seed), n, floor, switch)
    pi = floored_init_probs(synthetic_p_steady(n, seed=seed), p_floor)
    return A, pi
