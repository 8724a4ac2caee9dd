"""BASELINE.json config 5, latency case: ONE sequence of 1,000,000 frames x 361 states (development tool, not a test).

    python tools/latency_1m.py [--frames 1000000] [--check]

Times decode_device (forward + backtrace) on a single clip and, with --check, compares the whole path and the score
with the C oracle (about a minute of CPU).  |delta| reaches ~1e7 after 1 M frames (ulp 1): ties are everywhere, so this
is also the harshest first-maximum-wins test there is.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from viterbi_spl_b200 import ViterbiDecoder, hmm_params, synth

ap = argparse.ArgumentParser()
ap.add_argument('--frames', type=int, default=1000000)
ap.add_argument('--states', type=int, default=361)
ap.add_argument('--check', action='store_true')
a = ap.parse_args()
T, S = a.frames, a.states
A, pi = hmm_params.synthetic_hmm('tonet' if S == 361 else 'dcnet')
logA_T, log_pi = hmm_params.log_params(A, pi)
E = synth.device_dense_softmax(1, T, S, seed=5, device=torch.device('cuda'))
out = {'frames': T, 'states': S}
for algo in ('tmem', 'auto'):
    dec = ViterbiDecoder(logA_T, log_pi, algo=algo)
    p, s = dec.decode_device(E)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    e0.record(); p, s = dec.decode_device(E, forward_events=ev); e1.record(); torch.cuda.synchronize()
    out[algo] = {'total_ms': e0.elapsed_time(e1), 'forward_ms': ev[0].elapsed_time(ev[1]),
                 'frames_per_s': T / (e0.elapsed_time(e1) * 1e-3), 'score': float(s[0])}
    if algo == 'tmem':
        p_t, s_t = p.clone(), s.clone()
    else:
        out['auto_is'] = 'banded' if dec.structure.kind == 1 else 'tmem'
        out['tmem_equals_auto'] = bool(torch.equal(p, p_t) and torch.equal(s, s_t))
    del dec
if a.check:
    from oracle import c_oracle
    t0 = time.time()
    rp, rs = c_oracle.decode_batch_c(logA_T, log_pi, E.cpu().numpy())
    out['oracle_s'] = time.time() - t0
    out['paths_equal_oracle'] = bool(np.array_equal(rp, p_t.cpu().numpy()))
    out['score_equal_oracle'] = bool(np.array_equal(rs, s_t.cpu().numpy()))
print(json.dumps(out))
