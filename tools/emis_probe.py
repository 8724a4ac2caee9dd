"""Runs the emission kernel alone at the benchmark shape (for ncu captures and quick timing):
    python tools/emis_probe.py [softmax|shaun] [B] [T] [n_bins] [single_side_peak_width]
(n_bins 360, width 5: the dcnet / msnet / ftanet / tonet shape; 721 with 16 = jdc, 721 with 20 = imm)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viterbi_spl_b200 import pipeline

model = sys.argv[1] if len(sys.argv) > 1 else 'softmax'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
T = int(sys.argv[3]) if len(sys.argv) > 3 else 3000
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 360
spw = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dev = torch.device('cuda')
g = torch.Generator(device=dev); g.manual_seed(2)
n_in = nb + 1 if model == 'softmax' else nb
logits = 2 * torch.randn((B, T, n_in), device=dev, generator=g)
prior = (torch.rand((nb + 1,), device=dev, generator=g) * 0.01 + 1e-3) if model == 'softmax' else None
m = pipeline.SOFTMAX if model == 'softmax' else pipeline.SHAUN
E = pipeline.emissions_device(logits, nb, m, prior, spw, 0.0, True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(5):
    pipeline.emissions_device(logits, nb, m, prior, spw, 0.0, True, out=E)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({'model': model, 'B': B, 'T': T, 'n_bins': nb, 'single_side_peak_width': spw, 'ms': ms,
                  'GBps': 4.0 * B * T * (n_in + nb + 1) / (ms * 1e-3) / 1e9}))
