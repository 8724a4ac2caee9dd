"""Small-shape run of the kernels either side of the decode (register-window emission builder, both models; generic
emission kernel; voiced/bins; tcgen05 forward-backward) for compute-sanitizer, one tool per call:
    compute-sanitizer --tool memcheck  python tools/sanitize_post.py
    compute-sanitizer --tool racecheck python tools/sanitize_post.py
(The round-1 GPU pool refuses compute-sanitizer; tests/test_gpu_pipeline.py guards the output table with canaries instead.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import fb_oracle
from viterbi_spl_b200 import ForwardBackward, hmm_params, pipeline

ok = True
A, pi = hmm_params.synthetic_hmm('tonet')
g = torch.Generator(device='cuda'); g.manual_seed(5)
for model, n_in in (('softmax', 361), ('shaun', 360)):
    mp = pipeline.MelodyPipeline(A, pi, model=model, scaled=(model == 'softmax'))
    logits = 2 * torch.randn((3, 37, n_in), device='cuda', generator=g)
    a = mp.emissions(logits)
    os.environ['VIT_EMIS_GENERIC'] = '1'
    b = mp.emissions(logits)
    del os.environ['VIT_EMIS_GENERIC']
    good = bool(torch.equal(a == a.min(), b == b.min())) and bool(torch.allclose(a, b, rtol=1e-5, atol=1e-4)); ok &= good
    print('emissions', model, good)
    v, bn = mp(logits)
torch.cuda.synchronize()
lik = np.exp(2 * np.random.default_rng(0).standard_normal((33, 24, 361))).astype(np.float32)
L = np.asarray([24] * 30 + [0, 7, 1], np.int32)
gm, ll = ForwardBackward(A.astype(np.float32), pi.astype(np.float32)).run_host(lik, L)
wg, wl = fb_oracle.forward_backward_batch_np(A, pi, lik, L)
good = np.abs(gm - wg).max() < 1e-4; ok &= good; print('forward-backward (2 clusters, ragged)', good)
print('ALL OK' if ok else 'MISMATCH')
sys.exit(0 if ok else 1)
