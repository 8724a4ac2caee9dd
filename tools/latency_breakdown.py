"""Where one single-recording call (B = 1, 3000 frames) spends its time: forward kernel, backtrace kernels, host side."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from viterbi_spl_b200 import ViterbiDecoder, hmm_params, synth

out = {}
for name, S in (('dcnet', 321), ('tonet', 361), ('jdc', 722)):
    A, pi = hmm_params.synthetic_hmm(name)
    logA_T, log_pi = hmm_params.log_params(A, pi)
    T = 3000
    E = synth.dense_softmax(T, S, seed=1)[None]
    dec = ViterbiDecoder(logA_T, log_pi)
    dE = torch.from_numpy(E).cuda()
    res = {}
    for algo in ('auto', 'tmem'):
        d = ViterbiDecoder(logA_T, log_pi, algo=algo) if algo != 'auto' else dec
        try:
            d.decode_device(dE)
        except Exception as ex:
            res[algo] = str(ex)
            continue
        torch.cuda.synchronize()
        fe = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fw, tot = [], []
        for _ in range(5):
            e0.record()
            d.decode_device(dE, forward_events=fe)
            e1.record()
            torch.cuda.synchronize()
            fw.append(fe[0].elapsed_time(fe[1])); tot.append(e0.elapsed_time(e1))
        res[algo] = {'forward_ms': float(np.median(fw)), 'decode_device_ms': float(np.median(tot))}
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        dec.decode_host(E)
        ts.append(time.perf_counter() - t0)
    res['decode_host_ms'] = 1e3 * float(np.median(ts))
    out[name] = res
print(json.dumps(out))
