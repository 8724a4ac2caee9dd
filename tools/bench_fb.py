"""BASELINE.json config 4: scaled forward-backward posteriors, 1024 clips x 3000 frames x 361 states (development
bench; the contract bench is bench.py).  Prints one JSON line: frames/s, the FFMA roofline (one FFMA per cell, 2 S^2
cells per frame over both passes) and the HBM roofline (>= 20 S bytes per frame: b read twice, alpha~ written and read,
gamma written), plus the parity of a subset against the float64 oracle."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from viterbi_spl_b200 import ForwardBackward, hmm_params

ap = argparse.ArgumentParser()
ap.add_argument('--clips', type=int, default=1024)
ap.add_argument('--frames', type=int, default=3000)
ap.add_argument('--states', type=int, default=361)
ap.add_argument('--steps', type=int, default=5)
ap.add_argument('--warmup', type=int, default=2)
a = ap.parse_args()
B, T, S = a.clips, a.frames, a.states
A, pi = hmm_params.synthetic_hmm({321: 'dcnet', 361: 'tonet', 722: 'jdc'}[S])
A, pi = A.astype(np.float32), pi.astype(np.float32)
dev = torch.device('cuda')
g = torch.Generator(device=dev); g.manual_seed(4)
lik = torch.softmax(2.0 * torch.randn((B, T, S), device=dev, generator=g), dim=-1)       # dense softmax likelihoods
fb = ForwardBackward(A, pi)
gamma = torch.empty_like(lik); ll = torch.empty(B, device=dev)
for _ in range(a.warmup):
    fb.run_device(lik, None, gamma, ll)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(a.steps):
    fb.run_device(lik, None, gamma, ll)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
except Exception:
    pass
mhz, hbm = float(peaks.get('sm_max_mhz', 1965.0)), float(peaks.get('hbm_gbs', 6650.0))
cells = 2.0 * B * (T - 1) * S * S
ffma_peak = 148 * 128 * mhz * 1e6
bytes_ = 20.0 * B * T * S
from oracle import fb_oracle
sub = [0, B // 2, B - 1]
wg, wl = fb_oracle.forward_backward_batch_np(A, pi, lik[sub].cpu().numpy())
err = float(np.abs(gamma[sub].cpu().numpy() - wg).max())
rel = float(np.abs((ll[sub].cpu().numpy() - wl) / wl).max())
print(json.dumps({'metric': 'forward_backward_frames_per_sec', 'value': B * T / (ms * 1e-3), 'unit': 'frames/s',
                  'ms_per_step': ms, 'config': {'workload': f'scaled forward-backward {B} x {T} x {S}', 'dtype': 'f32'},
                  'roofline': {'bound': 'fp32_ffma', 'achieved': cells / (ms * 1e-3) / 1e12, 'peak': ffma_peak / 1e12,
                               'unit': 'Tcell/s', 'frac': cells / (ms * 1e-3) / ffma_peak},
                  'roofline_hbm': {'bound': 'hbm', 'achieved': bytes_ / (ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                                   'frac': bytes_ / (ms * 1e-3) / 1e9 / hbm},
                  'parity': {'max_abs_gamma_err': err, 'max_rel_loglik_err': rel, 'clips_checked': len(sub)}}))
