"""Parity probe + timing of VIT_ALGO_STREAM against the oracle / the tensor-memory kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import np_oracle
from viterbi_spl_b200 import ViterbiDecoder, hmm_params, synth

ok = True
for S, B, T in ((5, 3, 9), (97, 20, 30), (361, 33, 40), (722, 17, 25), (1000, 15, 12)):
    A, pi = synth.dyadic_hmm(S, seed=S)
    E = synth.batch('dyadic', B, T, S, seed0=S)
    L = np.random.default_rng(S).integers(0, T + 1, size=B).astype(np.int32); L[0] = T
    wp, ws = np_oracle.decode_batch_np(A, pi, E, L)
    p, s = ViterbiDecoder(A, pi, algo='stream').decode_host(E, L)
    good = np.array_equal(p, wp) and np.array_equal(s, ws)
    ok &= good
    print('parity', S, B, T, good, flush=True)
if len(sys.argv) > 1:
    for name, S, B, T in (('imm', 722, 4096, 500), ('tonet', 361, 4144, 1000)):
        A, pi = hmm_params.synthetic_hmm(name)
        logA_T, log_pi = hmm_params.log_params(A, pi, add_tiny=(name != 'imm'))
        E = synth.device_dense_softmax(B, T, S, seed=1, device='cuda')
        res = {}
        for algo in ('stream', 'tmem'):
            dec = ViterbiDecoder(logA_T, log_pi, algo=algo)
            fe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            for _ in range(2):
                p, s = dec.decode_device(E, forward_events=fe)
            torch.cuda.synchronize()
            ms = fe[0].elapsed_time(fe[1])
            res[algo] = (p.clone(), s.clone())
            cells = B * (T - 1) * S * S
            print(name, algo, 'forward ms %.2f' % ms, 'frac %.3f' % (cells / (ms * 1e-3) / (148 * 64 * 1.965e9)), flush=True)
        print('equal', torch.equal(res['stream'][0], res['tmem'][0]) and torch.equal(res['stream'][1], res['tmem'][1]))
print('ALL OK' if ok else 'PARITY FAILED')
