"""Small-shape run of every CUDA entry point for compute-sanitizer (memcheck / racecheck / synccheck), one tool per call:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
Checks results against the oracle as it goes, so a clean sanitizer log comes with a parity statement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import c_oracle, fb_oracle
from viterbi_spl_b200 import ViterbiDecoder, ForwardBackward, hmm_params, synth, pipeline

ok = True
A, pi = hmm_params.synthetic_hmm('tonet'); logA_T, log_pi = hmm_params.log_params(A, pi)
E = synth.batch('dense_softmax', 17, 140, 361, seed0=3)
L = np.asarray([140, 1, 0, 2, 139, 77] + [140] * 11, np.int32)
ref = c_oracle.decode_batch_c(logA_T, log_pi, E, L)
for algo in ('tmem', 'cluster', 'banded', 'backpointer'):
    p, s = ViterbiDecoder(logA_T, log_pi, algo=algo).decode_host(E, L)
    good = np.array_equal(p, ref[0]) and np.array_equal(s, ref[1]); ok &= good
    print(algo, 'S=361', good)
p, s = ViterbiDecoder(logA_T, log_pi, algo='tmem').decode_host(E, L, slab_frames=33)
good = np.array_equal(p, ref[0]); ok &= good; print('tmem slabs', good)
A7, pi7 = hmm_params.synthetic_hmm('jdc'); lA7, lp7 = hmm_params.log_params(A7, pi7)
E7 = synth.batch('dense_softmax', 15, 20, 722, seed0=4)
ref7 = c_oracle.decode_batch_c(lA7, lp7, E7)
p, s = ViterbiDecoder(lA7, lp7, algo='tmem').decode_host(E7)
good = np.array_equal(p, ref7[0]) and np.array_equal(s, ref7[1]); ok &= good; print('tmem S=722 (8-CTA clusters)', good)
Ad, pid = synth.dyadic_hmm(97, seed=1, coarse=True)
Ed = synth.batch('tie_stress', 9, 300, 97, seed0=2)
refd = c_oracle.decode_batch_c(Ad, pid, Ed)
p, s = ViterbiDecoder(Ad, pid, algo='tmem').decode_host(Ed)
good = np.array_equal(p, refd[0]); ok &= good; print('tmem S=97 ties, 3 backtrace segments', good)
lik = np.exp(2 * np.random.default_rng(0).standard_normal((16, 40, 361))).astype(np.float32)
g, ll = ForwardBackward(A.astype(np.float32), pi.astype(np.float32)).run_host(lik, np.asarray([40] * 14 + [0, 7], np.int32))
wg, wl = fb_oracle.forward_backward_batch_np(A, pi, lik, [40] * 14 + [0, 7])
good = np.abs(g - wg).max() < 1e-4; ok &= good; print('forward-backward', good)
mp = pipeline.MelodyPipeline(A, pi, model='softmax', scaled=True)
v, b = mp(2 * torch.randn((3, 50, 361), device='cuda'))
mp2 = pipeline.MelodyPipeline(A7, pi7, model='shaun', single_side_peak_width=16)
v2, b2 = mp2(2 * torch.randn((2, 30, 721), device='cuda'))
torch.cuda.synchronize()
print('pipeline ran', tuple(v.shape), tuple(v2.shape))
print('ALL OK' if ok else 'MISMATCH')
sys.exit(0 if ok else 1)
