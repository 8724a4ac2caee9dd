"""BASELINE.json config 4: scaled forward-backward posteriors, 1024 clips x 3000 frames x 361 states (development
bench; the contract bench is bench.py, key `forward_backward`).  Prints one JSON line: frames/s, the HBM roofline (>= 20 S
bytes per frame: b read twice, alpha~ written and read, gamma written) and the tensor-pipe roofline (bf16 MMA flops the
kernel EXECUTES -- 4 products of 2 bf16 terms over the padded 128-row x K shards -- against the measured dense bf16
peak), plus the parity of a subset against the float64 oracle."""
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
ap.add_argument('--impl', default='auto')
a = ap.parse_args()
B, T, S = a.clips, a.frames, a.states
A, pi = hmm_params.synthetic_hmm({321: 'dcnet', 361: 'tonet', 722: 'jdc'}[S])
A, pi = A.astype(np.float32), pi.astype(np.float32)
dev = torch.device('cuda')
g = torch.Generator(device=dev); g.manual_seed(4)
lik = torch.softmax(2.0 * torch.randn((B, T, S), device=dev, generator=g), dim=-1)       # dense softmax likelihoods
fb = ForwardBackward(A, pi, impl=a.impl)
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
bytes_ = 20.0 * B * T * S
C = -(-S // 124)                                             # CTAs per cluster (124 state rows each), K padded to 128 per shard
kp = C * (-(-((-(-S // C)) + (-(-S // C) + 30) // 31) // 32) * 32)
mma_flops = 2.0 * B * (T - 1) * 2.0 * (C * 128) * kp * 4     # fwd + bwd, 2 flop, padded rows x padded K, 4 bf16 products
tensor_peak = float(peaks.get('bf16_tflops', 1671.0))
from oracle import fb_oracle
sub = [0, B // 2, B - 1]
wg, wl = fb_oracle.forward_backward_batch_np(A, pi, lik[sub].cpu().numpy())
err = float(np.abs(gamma[sub].cpu().numpy() - wg).max())
rel = float(np.abs((ll[sub].cpu().numpy() - wl) / wl).max())
uses_tc = a.impl == 'tc' or (a.impl == 'auto' and not fb.structured and S <= 372)
rec = {'metric': 'forward_backward_frames_per_sec', 'value': B * T / (ms * 1e-3), 'unit': 'frames/s',
                  'ms_per_step': ms, 'config': {'workload': f'scaled forward-backward {B} x {T} x {S}', 'dtype': 'f32', 'impl': a.impl},
                  'roofline_tensor': {'bound': 'tensor', 'achieved': mma_flops / (ms * 1e-3) / 1e12, 'peak': tensor_peak,
                                      'unit': 'TFLOP/s (bf16 MMA flops executed)', 'frac': mma_flops / (ms * 1e-3) / 1e12 / tensor_peak,
                                      'note': 'tcgen05 kernel only (S <= 372); the FFMA kernel executes no MMA'},
                  'roofline_hbm': {'bound': 'hbm', 'achieved': bytes_ / (ms * 1e-3) / 1e9, 'peak': hbm, 'unit': 'GB/s',
                                   'frac': bytes_ / (ms * 1e-3) / 1e9 / hbm},
                  'parity': {'max_abs_gamma_err': err, 'max_rel_loglik_err': rel, 'clips_checked': len(sub)}}
if not uses_tc:
    del rec['roofline_tensor']          # only the tcgen05 kernel executes MMAs
print(json.dumps(rec))
