"""BASELINE.json config 0: ONE clip of 3000 frames through the reference's own entry points (drop-ins in
viterbi_spl_b200.reference_api) vs the NumPy restatement of the reference on one host core.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import np_oracle
from viterbi_spl_b200 import hmm_params, synth, reference_api as api

out = {}
for name, S in (('dcnet', 321), ('tonet', 361), ('jdc', 722)):
    A, pi = hmm_params.synthetic_hmm(name)
    logA_T, log_pi = hmm_params.log_params(A, pi)
    T = 3000
    E = synth.dense_softmax(T, S, seed=1)                   # [T, S] log-domain
    E_st = np.require(E.T, requirements=['C'])
    # log-domain entry point (imm/tf_viterbi.py:75)
    api.viterbi_librosa_fn(log_transition_matrix_T=logA_T, log_prob_init=log_pi, log_probs_st=E_st)   # warm-up: upload + workspace
    ts = []
    for _ in range(10):
        t0 = time.perf_counter()
        st = api.viterbi_librosa_fn(log_transition_matrix_T=logA_T, log_prob_init=log_pi, log_probs_st=E_st)
        ts.append(time.perf_counter() - t0)
    # Family C object (dcnet/softmax_viterbi.py:2636): prob-domain [T, S], logged in place on the host
    ns = {321: api.msnet, 361: api.tonet, 722: api.jdc}.get(S)
    sv = ns.SoftMaxViterbi(False, transition_matrix=np.asarray(A, np.float32), ini_probs=np.asarray(pi, np.float32)) \
        if (ns is not None and np.argmax(pi) == S - 1) else None
    tc = []
    if sv is not None:
        P = np.exp(E)
        sv.viterbi_librosa_fn(P.copy())
        for _ in range(5):
            Pc = P.copy()
            t0 = time.perf_counter()
            st_c = sv.viterbi_librosa_fn(Pc)
            tc.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    ref, _ = np_oracle.viterbi_log_np(logA_T, log_pi, E)
    t_cpu = time.perf_counter() - t0
    out[name] = {'states': S, 'frames': T, 'gpu_log_domain_call_ms': 1e3 * float(np.median(ts)),
                 'gpu_family_c_call_ms': 1e3 * float(np.median(tc)) if tc else None,
                 'cpu_numpy_reference_ms': 1e3 * t_cpu, 'paths_equal': bool(np.array_equal(st, ref)),
                 'speedup_log_domain': t_cpu / float(np.median(ts))}
print(json.dumps(out))
