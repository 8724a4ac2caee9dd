"""Times the emission builder and the whole on-GPU post-processing chain (logits -> voiced/bins) at the benchmark shape
and reports the emission kernel against the HBM roofline (8 B per state-frame: 4 read + 4 written)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from viterbi_spl_b200 import hmm_params, pipeline

B, T, nb = 1024, 3000, 360
A, pi = hmm_params.synthetic_hmm('tonet')
dev = torch.device('cuda')
g = torch.Generator(device=dev); g.manual_seed(2)
out = {}
for model, scaled in (('softmax', True), ('shaun', False)):
    mp = pipeline.MelodyPipeline(A, pi, model=model, scaled=scaled)
    n_in = nb + 1 if model == 'softmax' else nb
    logits = 2 * torch.randn((B, T, n_in), device=dev, generator=g)
    E = mp.emissions(logits)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        E = pipeline.emissions_device(logits, nb, mp.model, mp.prior, mp.spw, mp.threshold, True, out=E)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    bytes_ = 4.0 * B * T * (n_in + nb + 1)
    del E
    mp(logits)                                  # warm-up of the whole chain
    torch.cuda.synchronize(); e0.record()
    for _ in range(3):
        v, b = mp(logits)
    e1.record(); torch.cuda.synchronize()
    chain_ms = e0.elapsed_time(e1) / 3
    del v, b
    mp.posteriors(logits)                       # logits -> likelihoods -> forward-backward posteriors
    torch.cuda.synchronize(); e0.record()
    for _ in range(3):
        gam, ll = mp.posteriors(logits)
    e1.record(); torch.cuda.synchronize()
    post_ms = e0.elapsed_time(e1) / 3
    del gam, ll
    v, b = mp(logits)
    out[model] = {'posterior_chain_ms': post_ms, 'posterior_chain_frames_per_s': B * T / (post_ms * 1e-3),
                  'emissions_ms': ms, 'emissions_GBps': bytes_ / (ms * 1e-3) / 1e9,
                  'chain_ms': chain_ms, 'chain_frames_per_s': B * T / (chain_ms * 1e-3),
                  'voiced_fraction': float(v.float().mean())}
    del mp, logits, v, b
    torch.cuda.empty_cache()
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
for m in out.values():
    m['emissions_frac_of_measured_hbm'] = m['emissions_GBps'] / peaks['hbm_gbs']
print(json.dumps(out))
