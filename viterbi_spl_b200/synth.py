"""Seeded synthetic log-emissions of the shapes named in BASELINE.json.

There is no network for datasets or checkpoints, so the decoder is exercised on synthetic posteriors:

* ``dense_softmax``  logits ~ N(0, 2^2) + 4 on a random-walk pitch track with voiced/unvoiced segments -> softmax ->
                     ``log(p + tiny)`` (what a SoftMax-style emission model produces, dcnet/softmax_viterbi.py:2565-2572)
* ``sparse_peaks``   0-5 peaks per frame + the unvoiced state, all other bins exactly 0 -> ``log(tiny) = -87.33655``
                     (shape of SoftMaxViterbi.observation_probs_fn output, dcnet/softmax_viterbi.py:2530-2579)
* ``dyadic``         ``-(k / 2^10)``, k uniform in [0, 2^16): machine-independent values (no libm involved)
* ``tie_stress``     ``-(k / 4)``, k in [0, 8): ~44 % of (state, frame) cells are exactly tied, so any deviation from
                     NumPy's first-maximum-wins argmax shows up immediately

Host generators use NumPy (bit-reproducible from the seed); ``device_dense_softmax`` is the torch generator the
benchmark uses for full-size inputs (values differ from the NumPy one; parity is always checked on host-generated
clips that the oracle also sees).
"""
import numpy as np

TINY = np.finfo(np.float32).tiny
LOG_TINY = np.float32(np.log(TINY))  # -87.33655


def pitch_track(T, n_bins, rng):
    """Random-walk voiced pitch track with unvoiced gaps; returns state index per frame (unvoiced = n_bins)."""
    track = np.empty([T], np.int64)
    pos = rng.integers(0, n_bins)
    voiced = rng.random() < 0.5
    t = 0
    while t < T:
        seg = int(rng.integers(20, 200))
        for _ in range(min(seg, T - t)):
            if voiced:
                pos = int(np.clip(pos + rng.integers(-2, 3), 0, n_bins - 1))
                track[t] = pos
            else:
                track[t] = n_bins
            t += 1
        voiced = not voiced
        if voiced:
            pos = rng.integers(0, n_bins)
    return track


def dense_softmax(T, S, seed):
    """[T, S] float32 log-emissions, every state has positive probability."""
    rng = np.random.default_rng(seed)
    logits = (2.0 * rng.standard_normal((T, S))).astype(np.float32)
    tr = pitch_track(T, S - 1, rng)
    logits[np.arange(T), tr] += np.float32(4.0)
    logits -= logits.max(axis=1, keepdims=True)
    p = np.exp(logits)
    p /= p.sum(axis=1, keepdims=True)
    return np.log(p + TINY).astype(np.float32)


def sparse_peaks(T, S, seed, max_peaks=5):
    """[T, S] float32 log-emissions with at most `max_peaks` voiced peaks per frame + the unvoiced state (last)."""
    rng = np.random.default_rng(seed)
    probs = np.zeros([T, S], np.float32)
    tr = pitch_track(T, S - 1, rng)
    for t in range(T):
        k = int(rng.integers(0, max_peaks + 1))
        idx = rng.choice(S - 1, size=k, replace=False) if k else np.empty([0], np.int64)
        if tr[t] < S - 1 and k:
            idx[0] = tr[t]
        idx = np.unique(np.append(idx, S - 1))
        w = np.exp(2.0 * rng.standard_normal(len(idx))).astype(np.float32)
        if tr[t] < S - 1 and k:
            w[np.searchsorted(idx, tr[t])] *= np.float32(20.0)
        probs[t, idx] = w / w.sum()
    return np.log(probs + TINY).astype(np.float32)


def dyadic(shape, seed, bits=16, frac_bits=10):
    rng = np.random.default_rng(seed)
    return (-(rng.integers(0, 1 << bits, size=shape).astype(np.float64) / float(1 << frac_bits))).astype(np.float32)


def tie_stress(shape, seed):
    rng = np.random.default_rng(seed)
    return (-(rng.integers(0, 8, size=shape).astype(np.float64) / 4.0)).astype(np.float32)


def dyadic_hmm(S, seed, coarse=False):
    """Machine-independent (logA_T, log_pi): quantised negative values, no libm."""
    gen = tie_stress if coarse else dyadic
    return np.ascontiguousarray(gen((S, S), seed + 1000)), gen((S,), seed + 2000)


def batch(kind, B, T, S, seed0=0):
    """[B, T, S] float32, clip b generated with seed seed0 + b."""
    fn = {'dense_softmax': dense_softmax, 'sparse_peaks': sparse_peaks,
          'dyadic': lambda T_, S_, s: dyadic((T_, S_), s), 'tie_stress': lambda T_, S_, s: tie_stress((T_, S_), s)}[kind]
    out = np.empty([B, T, S], np.float32)
    for b in range(B):
        out[b] = fn(T, S, seed0 + b)
    return out


def device_dense_softmax(B, T, S, seed, device, out=None, chunk=64):
    """Full-size benchmark input generated on the GPU with torch: log_softmax of N(0, 2^2) logits with a +4 bump on a
    per-clip random-walk track.  Returns a [B, T, S] float32 CUDA tensor."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    if out is None:
        out = torch.empty((B, T, S), dtype=torch.float32, device=device)
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        n = b1 - b0
        logits = 2.0 * torch.randn((n, T, S), generator=g, device=device, dtype=torch.float32)
        steps = torch.randint(-2, 3, (n, T), generator=g, device=device)
        start = torch.randint(0, S - 1, (n, 1), generator=g, device=device)
        track = (start + torch.cumsum(steps, dim=1)).remainder(S - 1)
        unvoiced = (torch.rand((n, (T + 127) // 128), generator=g, device=device) < 0.4).repeat_interleave(128, dim=1)[:, :T]
        track = torch.where(unvoiced, torch.full_like(track, S - 1), track)
        logits.scatter_add_(2, track.unsqueeze(-1), torch.full((n, T, 1), 4.0, device=device))
        out[b0:b1] = torch.log_softmax(logits, dim=-1)
    return out
