"""Quick on-GPU parity probe (development aid; the real tests are tests/test_gpu_*.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from viterbi_spl_b200 import ViterbiDecoder, hmm_params, synth
from oracle import c_oracle

def check(name, logA_T, log_pi, E, lengths=None, algos=('backpointer', 'cluster', 'tmem')):
    ref_p, ref_s = c_oracle.decode_batch_c(logA_T, log_pi, E, lengths)
    for algo in algos:
        try:
            dec = ViterbiDecoder(logA_T, log_pi, algo=algo)
            t0 = time.time()
            p, s = dec.decode_host(E, lengths)
            dt = time.time() - t0
        except Exception as ex:
            print(f'{name:28s} {algo:12s} ERROR {ex}'); continue
        okp = np.array_equal(p, ref_p); oks = np.array_equal(s, ref_s)
        nbad = int((p != ref_p).any(axis=1).sum())
        print(f'{name:28s} {algo:12s} paths_equal={okp} scores_equal={oks} bad_clips={nbad}/{len(E)} {dt*1e3:.1f} ms')

rng = np.random.default_rng(0)
for S, T, B in [(7, 5, 3), (97, 50, 5), (200, 64, 9), (321, 100, 33), (361, 120, 70), (384, 40, 3)]:
    A, pi = synth.dyadic_hmm(S, seed=S)
    E = synth.batch('dyadic', B, T, S, seed0=10)
    check(f'dyadic S={S} T={T} B={B}', A, pi, E)
    A, pi = synth.dyadic_hmm(S, seed=S, coarse=True)
    E = synth.batch('tie_stress', B, T, S, seed0=20)
    check(f'ties   S={S} T={T} B={B}', A, pi, E)
A, pi = hmm_params.synthetic_hmm('tonet'); logA_T, log_pi = hmm_params.log_params(A, pi)
E = synth.batch('dense_softmax', 40, 150, 361, seed0=3)
L = rng.integers(0, 151, size=40).astype(np.int32); L[0] = 150; L[1] = 1; L[2] = 0; L[3] = 2
check('tonet dense ragged', logA_T, log_pi, E, L)
E = synth.batch('sparse_peaks', 36, 150, 361, seed0=4)
check('tonet sparse', logA_T, log_pi, E)
A, pi = hmm_params.synthetic_hmm('imm'); logA_T, log_pi = hmm_params.log_params(A, pi, add_tiny=False)
E = synth.batch('dense_softmax', 4, 60, 722, seed0=5)
check('imm dense S=722', logA_T, log_pi, E, algos=('backpointer', 'tmem'))
