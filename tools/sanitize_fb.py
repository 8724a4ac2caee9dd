"""Small-shape run of the structured forward-backward kernels (convolution narrow / wide, general band, dense FFMA
fall-back) for compute-sanitizer, one tool per call:
    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python tools/sanitize_fb.py
Checks gamma against the float64 oracle as it goes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import fb_oracle
from viterbi_spl_b200 import ForwardBackward, hmm_params

rng = np.random.default_rng(0)
ok = True


def band(S, d, dense):
    A = np.zeros((S, S))
    i, j = np.indices((S, S))
    m = np.abs(i - j) <= d
    A[m] = rng.random(m.sum()) + 0.01
    if dense is not None:
        A[dense, :] = rng.random(S) * 0.01 + 1e-4
        A[:, dense] = rng.random(S) * 0.05 + 1e-3
    A /= A.sum(1, keepdims=True)
    pi = rng.random(S) + 0.01
    return A.astype(np.float32), (pi / pi.sum()).astype(np.float32)


cases = [('conv narrow tonet', *[x.astype(np.float32) for x in hmm_params.synthetic_hmm('tonet')], 9, 13, None),
         ('conv narrow dcnet', *[x.astype(np.float32) for x in hmm_params.synthetic_hmm('dcnet')], 5, 7, None),
         ('conv wide jdc', *[x.astype(np.float32) for x in hmm_params.synthetic_hmm('jdc')], 4, 6, None),
         ('general band S=200', *band(200, 9, 199), 9, 11, None),
         ('general band S=384 no dense', *band(384, 14, None), 8, 5, None),
         ('wide general -> dense FFMA', *band(722, 30, 721), 3, 5, None),
         ('general band, conv off', *[x.astype(np.float32) for x in hmm_params.synthetic_hmm('tonet')], 9, 6, '0')]
for name, A, pi, B, T, conv in cases:
    S = len(pi)
    lik = np.exp(rng.standard_normal((B, T, S))).astype(np.float32)
    L = rng.integers(0, T + 1, size=B).astype(np.int32)
    L[:3] = [T, 1, 0]
    if conv is not None:
        os.environ['VIT_FB_CONV'] = conv
    g, ll = ForwardBackward(A, pi, impl='banded').run_host(lik, L)
    os.environ.pop('VIT_FB_CONV', None)
    wg, wl = fb_oracle.forward_backward_batch_np(A, pi, lik, L)
    good = bool(np.abs(g - wg).max() < 1e-4 and np.allclose(ll, wl, rtol=1e-5, atol=1e-5))
    ok &= good
    print(name, good)
print('all ok' if ok else 'MISMATCH')
